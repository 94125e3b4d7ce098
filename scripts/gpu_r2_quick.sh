#!/bin/bash
# quick check after a kernel change: sparse Cholesky parity tests + the cfg5 / cfg3 / cfg4 bench lines (no CPU baseline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "spchol or lockstep_exact or sparse" > gpurun_out/r2_quick_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_quick_tests.log
tail -4 gpurun_out/r2_quick_tests.log
for wl in cfg5 cfg3 cfg4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r2_quick_$wl.log 2>&1
  python - gpurun_out/r2_quick_$wl.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["config"]["workload"][:5], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["final_cost"])
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
    elif 'Error' in l or 'error' in l: print(l.strip()[:300])
PY
done
BA_SPCHOL_PROF=1 python profiles/profile_target.py 5 5 500 4 2>&1 | tail -3
BA_SPCHOL_PROF=1 python profiles/profile_target.py 3 5 500 4 2>&1 | tail -3
