#!/bin/bash
# round 2: full GPU suite + bench of the sparse Cholesky path after the kernel changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spchol or out_of_range or auto_picks" > gpurun_out/r2_spchol_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_spchol_tests.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r2_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2_gputests.log
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r2_fullsize.log 2>&1; echo rc=$? >> gpurun_out/r2_fullsize.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg5_auto.log 2>&1
for wl in cfg3 cfg4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/r2_bench_${wl}_auto.log 2>&1
done
tail -c 300 gpurun_out/r2_spchol_tests.log; tail -c 600 gpurun_out/r2_gputests.log; tail -c 600 gpurun_out/r2_fullsize.log
for f in gpurun_out/r2_bench_*_auto.log; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["detail"]["solver"][:50], d["final_cost"])
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
        print("roofline", {k: (round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k not in ("timing","traffic_source")})
        print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("seconds"), "parity", d.get("parity_vs_n1"))
    elif 'Error' in l or 'error' in l: print(l.strip()[:300])
PY
done
