#!/usr/bin/env python
"""Synthetic keyframe sequence for bench/ceres_baseline.cpp, and the comparison of two of its result files.

  python scripts/dump_sequence.py seq.bin [--cfg 2] [--scale 1.0]       write the sequence (BASELINE.json config 1 or 2)
  python scripts/dump_sequence.py --compare ref.bin gpu.bin             poses / landmarks / intrinsics of two runs

The comparison prints the largest translation and rotation difference over the keyframes (north-star bound after a fixed LM
iteration count: 1e-6 m / 1e-6 rad) and the largest landmark difference."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dump(path, cfg, scale):
    import ba_b200
    seq = ba_b200.synthetic.make_config(cfg, scale=scale)
    n_kf, n_obs, n_lm = int(seq.pose.shape[0]), int(len(seq.lm)), int(seq.pt.shape[0])
    with open(path, "wb") as f:
        f.write(struct.pack("<Iiii", 0xBA5E0001, n_kf, n_obs, n_lm))
        f.write(np.ascontiguousarray(seq.kf_ptr, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(seq.lm, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(seq.uv, dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(seq.depth, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(seq.pose, dtype="<f8").tobytes())
        f.write(np.arange(n_lm, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(seq.pt, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(seq.K, dtype="<f8").tobytes())
    print("%s: %d keyframes, %d observations, %d landmarks" % (path, n_kf, n_obs, n_lm))


def load_result(path):
    raw = open(path, "rb").read()
    magic, n_kf, n_lm, calls = struct.unpack_from("<Iiii", raw, 0)
    if magic != 0xBA5E0002:
        raise SystemExit("%s: not a result file" % path)
    off = 16
    pose = np.frombuffer(raw, "<f8", 7 * n_kf, off).reshape(n_kf, 7)
    off += 56 * n_kf
    pt = np.frombuffer(raw, "<f8", 3 * n_lm, off).reshape(n_lm, 3)
    off += 24 * n_lm
    return pose, pt, np.frombuffer(raw, "<f8", 4, off), calls


def compare(a, b):
    pa, la, ka, ca = load_result(a)
    pb, lb, kb, cb = load_result(b)
    if pa.shape != pb.shape or la.shape != lb.shape or ca != cb:
        raise SystemExit("different sequences / schedules: %s vs %s" % ((pa.shape, la.shape, ca), (pb.shape, lb.shape, cb)))
    dt = np.max(np.linalg.norm(pa[:, 4:] - pb[:, 4:], axis=1))
    qa, qb = pa[:, :4] / np.linalg.norm(pa[:, :4], axis=1, keepdims=True), pb[:, :4] / np.linalg.norm(pb[:, :4], axis=1, keepdims=True)
    dr = np.max(2.0 * np.arcsin(np.minimum(1.0, np.minimum(np.linalg.norm(qa - qb, axis=1), np.linalg.norm(qa + qb, axis=1)) / 2.0)))
    both = ~(np.isnan(la).any(axis=1) | np.isnan(lb).any(axis=1))
    dl = np.max(np.linalg.norm(la[both] - lb[both], axis=1)) if both.any() else 0.0
    dk = np.max(np.abs(ka - kb))
    ok = dt < 1e-6 and dr < 1e-6
    print("%d windowOptimize calls: max |dt| %.3e m, max rotation difference %.3e rad, max landmark difference %.3e m, "
          "max intrinsics difference %.3e -> %s (bound 1e-6 m / 1e-6 rad)" % (ca, dt, dr, dl, dk, "OK" if ok else "DIFFERENT"))
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--compare":
        sys.exit(compare(sys.argv[2], sys.argv[3]))
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    cfg = int(sys.argv[sys.argv.index("--cfg") + 1]) if "--cfg" in sys.argv else 2
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    dump(sys.argv[1], cfg, scale)
