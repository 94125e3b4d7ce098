#!/bin/bash
# after a change to the block-sparse Schur / camera-block kernels: their parity tests (reduced and full size), then the cfg5 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "sparse or spchol or lockstep or product or eval" > gpurun_out/r2_check_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_check_tests.log
tail -n 3 gpurun_out/r2_check_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg5.log 2>&1; echo "cfg5 rc=$?"
python - gpurun_out/r02_bench_cfg5.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["final_cost"], d.get("parity_vs_n1"), d["roofline"]["frac"], d["cpu_baseline"]["value"])
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
