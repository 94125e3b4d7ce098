#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
BA_STORE_PROF=1 python - <<'PY' 2>&1 | tail -12
import sys
sys.path.insert(0, '.')
import ba_b200
syn = ba_b200.synthetic
for growing in (False, True):
    seq = syn.make_config(2)
    r = ba_b200.hostlib.sliding_sequence(seq, 20, 10, max_num_iterations=10, fixed_iterations=True, device_store=True, growing_maps=growing)
    w = r["windows"]
    print("growing=%d: %.1f windows/s | " % (growing, 1e3 * w / r["ms"]["total"]) + " ".join("%s %.3f" % (k, v / w) for k, v in r["ms"].items()), flush=True)
PY
