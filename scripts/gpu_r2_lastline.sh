#!/bin/bash
# the default bench line at HEAD (cfg5, with the CPU baseline) and smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg5.log 2>&1; echo "cfg5 rc=$?"
python - gpurun_out/r02_bench_cfg5.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["final_cost"], d.get("parity_vs_n1")["ok"], d["roofline"]["kernel"][:30], d["roofline"]["frac"], d["cpu_baseline"]["value"])
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
