#!/bin/bash
# device-resident store (N1): parity tests, then the cfg2 sliding schedule with and without the store
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_wrapper.py -m gpu -x -q > gpurun_out/r2_store_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_store_tests.log
tail -30 gpurun_out/r2_store_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/r2_store_sliding.log
import sys, json
sys.path.insert(0, '.')
import ba_b200
syn = ba_b200.synthetic
for growing in (False, True):
    for store in (False, True):
        for rep in range(2):
            seq = syn.make_config(2)
            r = ba_b200.hostlib.sliding_sequence(seq, 20, 10, max_num_iterations=10, fixed_iterations=True, device_store=store, growing_maps=growing)
        w = r["windows"]
        print("growing=%d store=%d: %d windows, %.1f windows/s | per window ms: " % (growing, store, w, 1e3 * w / r["ms"]["total"])
              + " ".join("%s %.3f" % (k, v / w) for k, v in r["ms"].items()))
PY
# band-aware blocked Cholesky (global REF problem): parity tests + the cfg3ref bench line, dense for comparison
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blocked_cholesky" > gpurun_out/r2_chol_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_chol_tests.log
tail -5 gpurun_out/r2_chol_tests.log
for env in "X=1" "BA_CHOL_DENSE=1"; do
  env $env timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg3ref --no-cpu-baseline > gpurun_out/r2_bench_cfg3ref_$env.log 2>&1
  python - "gpurun_out/r2_bench_cfg3ref_$env.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1][-14:], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["final_cost"], d["detail"]["solver"][:30])
    elif 'Error' in l or 'error' in l: print(l.strip()[:300])
PY
done
