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
