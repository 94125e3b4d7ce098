#!/bin/bash
# after a change to the block-sparse Schur kernels: their parity tests (reduced and full size) and the cfg5 / cfg4 bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "sparse or spchol or lockstep or product" > gpurun_out/r2_check_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_check_tests.log
tail -n 4 gpurun_out/r2_check_tests.log
for wl in cfg5 cfg4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r2_check_$wl.log 2>&1
  python - gpurun_out/r2_check_$wl.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["config"]["workload"][:5], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["final_cost"], d.get("parity_vs_n1"))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
done
