#!/bin/bash
# full GPU suite + the bench lines a change to the camera-block kernels touches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests.log
tail -n 4 gpurun_out/r02_gputests.log
for wl in cfg5 cfg2 cfg3ref; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r2_check_$wl.log 2>&1
  python - gpurun_out/r2_check_$wl.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["config"]["workload"][:7], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["final_cost"], d.get("parity_vs_n1"))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
        if "sliding_sequence" in d: print(d["sliding_sequence"]["cpp_dropin"]["windows_per_s"])
PY
done
