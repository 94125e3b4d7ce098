#!/usr/bin/env python
"""Generates tests/golden/eval_golden.json by running the REFERENCE's own evaluation scripts
(rgb-d-toolset/evaluate_ate.py, evaluate_rpe.py, associate.py, imported from /root/reference with a
stub for the missing matplotlib) on small synthetic TUM-format trajectories.  Run in the build
container only (the GPU box has no /root/reference); the JSON is committed.

  python tests/golden/make_eval_golden.py
"""
import io
import json
import os
import sys
import tempfile
import types
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference/rgb-d-toolset"
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "matplotlib.patches"):
    m = types.ModuleType(name)
    m.use = lambda *a, **k: None
    sys.modules[name] = m
sys.path.insert(0, REF)
import associate  # noqa: E402
import evaluate_ate  # noqa: E402
import evaluate_rpe  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402


def make_case(seed, n, dt, noise_t, noise_r, drop):
    rng = np.random.default_rng(seed)
    t = 1305031100.0 + dt * np.arange(n)
    pos = np.stack([0.3 * np.sin(0.7 * np.arange(n) * dt), 0.2 * np.cos(0.5 * np.arange(n) * dt), 0.05 * np.arange(n) * dt], 1)
    rv = np.stack([0.1 * np.sin(0.3 * np.arange(n) * dt), 0.05 * np.arange(n) * dt, 0.08 * np.cos(0.4 * np.arange(n) * dt)], 1)
    q = Rotation.from_rotvec(rv).as_quat()
    gt = "# ground truth trajectory\n# file: synthetic\n# timestamp tx ty tz qx qy qz qw\n" + "".join(
        "%.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f\n" % (t[i], *pos[i], *q[i]) for i in range(n))
    # estimate: a rigidly moved, noisy, sub-sampled copy with slightly shifted stamps
    Rm = Rotation.from_rotvec([0.2, -0.1, 0.3])
    keep = [i for i in range(n) if i % drop != drop - 1]
    pe = Rm.apply(pos[keep]) + np.array([0.5, -0.2, 0.1]) + rng.normal(size=(len(keep), 3)) * noise_t
    qe = (Rm * Rotation.from_rotvec(rv[keep] + rng.normal(size=(len(keep), 3)) * noise_r)).as_quat()
    te = t[keep] + rng.uniform(-0.004, 0.004, size=len(keep))
    est = "".join("%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n" % (te[i], *pe[i], *qe[i]) for i in range(len(keep)))
    return gt, est


def run_reference(gt_text, est_text, delta_rot, horn, rpe_delta, rpe_unit):
    with tempfile.TemporaryDirectory() as d:
        g, e = os.path.join(d, "gt.txt"), os.path.join(d, "est.txt")
        open(g, "w").write(gt_text)
        open(e, "w").write(est_text)
        first, second = associate.read_file_list(g), associate.read_file_list(e)
        matches = associate.associate(first, second, 0.0, 0.02)
        fx = np.matrix([[float(v) for v in first[a][0:3]] for a, b in matches]).transpose()
        sx = np.matrix([[float(v) for v in second[b][0:3]] for a, b in matches]).transpose()
        rot, trans, err = evaluate_ate.align(sx, fx)
        fq = np.array([[float(v) for v in first[a][3:7]] for a, b in matches])
        sq = np.array([[float(v) for v in second[b][3:7]] for a, b in matches])
        fe = Rotation.from_quat(fq).as_euler("xyz", degrees=True)
        so = Rotation.from_quat(sq)
        if horn:
            so = Rotation.from_matrix(rot) * so
        se = so.as_euler("xyz", degrees=True)
        with redirect_stdout(io.StringIO()):
            roterr = {
                "AYE": evaluate_ate.calculateAYE(fe[:, 0], se[:, 0]), "APE": evaluate_ate.calculateAPE(fe[:, 1], se[:, 1]),
                "ARE": evaluate_ate.calculateARE(fe[:, 2], se[:, 2]),
                "RYE": evaluate_ate.calculateRYE(fe[:-delta_rot, 0], fe[delta_rot:, 0], se[:-delta_rot, 0], se[delta_rot:, 0]),
                "RPE": evaluate_ate.calculateRPE(fe[:-delta_rot, 1], fe[delta_rot:, 1], se[:-delta_rot, 1], se[delta_rot:, 1]),
                "RRE": evaluate_ate.calculateRRE(fe[:-delta_rot, 2], fe[delta_rot:, 2], se[:-delta_rot, 2], se[delta_rot:, 2]),
            }
        rows = evaluate_rpe.evaluate_trajectory(evaluate_rpe.read_trajectory(g), evaluate_rpe.read_trajectory(e), 0, True,
                                                rpe_delta, rpe_unit, 0.0, 1.0)
        return {"matches": [[a, b] for a, b in matches], "rot": np.asarray(rot).tolist(), "trans": np.asarray(trans).ravel().tolist(),
                "trans_error": np.asarray(err).tolist(), "rotation_errors": {k: float(v) for k, v in roterr.items()},
                "rpe_rows": np.asarray(rows).tolist()}


def main():
    cases = []
    for seed, n, dt, nt, nr, drop, delta_rot, horn, rd, ru in [
            (1, 60, 0.1, 0.01, 0.01, 7, 5, False, 1.0, "s"),
            (2, 120, 0.033, 0.003, 0.004, 4, 3, True, 5.0, "f"),
            (3, 40, 0.2, 0.02, 0.02, 9, 5, False, 0.3, "m")]:
        gt, est = make_case(seed, n, dt, nt, nr, drop)
        out = run_reference(gt, est, delta_rot, horn, rd, ru)
        cases.append({"gt": gt, "est": est, "delta_rot": delta_rot, "horn": horn, "rpe_delta": rd, "rpe_unit": ru, "expect": out})
    json.dump({"generator": "tests/golden/make_eval_golden.py (reference scripts: rgb-d-toolset/evaluate_ate.py, evaluate_rpe.py, "
                            "associate.py)", "cases": cases}, open(os.path.join(HERE, "eval_golden.json"), "w"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
