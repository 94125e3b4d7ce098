#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scripts/phase_probe.py 5 10 4 2>&1 | tail -1
python scripts/phase_probe.py 3 10 4 2>&1 | tail -1
python scripts/phase_probe.py 4 10 4 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py -m gpu -x -q -k "sparse or spchol or store" 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -6
import sys
sys.path.insert(0, '.')
import ba_b200
syn = ba_b200.synthetic
for growing in (False, True):
    for store in (False, True):
        for rep in range(2):
            seq = syn.make_config(2)
            r = ba_b200.hostlib.sliding_sequence(seq, 20, 10, max_num_iterations=10, fixed_iterations=True, device_store=store, growing_maps=growing)
        w = r["windows"]
        print("growing=%d store=%d: %.1f windows/s | " % (growing, store, 1e3 * w / r["ms"]["total"]) + " ".join("%s %.3f" % (k, v / w) for k, v in r["ms"].items()), flush=True)
PY
