"""Shared helpers of the parity tests: product (C-ABI over CUDA) vs CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ba_b200  # noqa: E402
import oracle_lib as ora  # noqa: E402


def to_oracle(p):
    return ora.Problem(p.pose7, p.pt3, p.cam_idx, p.pt_idx, p.uv2, p.depth, p.intr, p.intr_prior, p.fixed_cam)


def mode_opts(mode, **kw):
    """(gpu options kwargs, oracle options kwargs) for REF / NS cost models."""
    d, k = {"REF": (1, 1), "NS": (0, 0), "DEPTH": (1, 0), "INTR": (0, 1)}[mode]
    g = dict(use_depth_prior=d, optimize_intrinsics=k)
    o = dict(use_depth_prior=d, optimize_intrinsics=k)
    for key, v in kw.items():
        if key == "solver":
            g["solver"] = v
            # oracle: 0 dense Schur (exact), 1 implicit PCG, 2 SPARSE_SCHUR-equivalent envelope Cholesky (exact)
            o["solver"] = 0 if v in (0, 1) else (2 if v == 4 else 1)
        elif key == "jacobian_store":
            g[key] = v
        elif key == "max_num_iterations":
            g[key] = v
            o[key] = v
        else:
            g[key] = v
    return g, o


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / max(s, 1e-300)) if a.size else 0.0


def pose_err(pa, pb):
    """max translation diff (m), max rotation angle diff (rad)."""
    dt = float(np.max(np.linalg.norm(pa[:, 4:] - pb[:, 4:], axis=1)))
    dr = float(np.max(ba_b200.se3.rot_angle(pa[:, :4], pb[:, :4])))
    return dt, dr
