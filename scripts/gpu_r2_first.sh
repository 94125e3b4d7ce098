#!/bin/bash
# round 2, first GPU contact: the new sparse Cholesky + index guards, then the whole suite, then a first bench comparison
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt; free -g >> gpurun_out/r2_nproc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spchol or out_of_range or auto_picks" > gpurun_out/r2_spchol_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_spchol_tests.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py -x > gpurun_out/r2_gputests.log 2>&1; echo rc=$? >> gpurun_out/r2_gputests.log
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=30 > gpurun_out/r2_fullsize.log 2>&1; echo rc=$? >> gpurun_out/r2_fullsize.log
for sv in cholesky sparse; do
  timeout 600 python bench.py --steps 10 --warmup 3 --solver $sv --no-cpu-baseline > gpurun_out/r2_bench_cfg5_$sv.log 2>&1
done
for wl in cfg3 cfg4; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --solver cholesky --no-cpu-baseline > gpurun_out/r2_bench_${wl}_cholesky.log 2>&1
done
tail -c 600 gpurun_out/r2_spchol_tests.log; tail -c 400 gpurun_out/r2_gputests.log; tail -c 1500 gpurun_out/r2_fullsize.log
for f in gpurun_out/r2_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["solver"][:40], d["final_cost"], d["config"].get("pcg_iterations_per_lm")); print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()}); print(d.get("sparse_cholesky"))
PY
done
