"""Pins the CPU oracle against the 60-digit mpmath golden vectors (CPU-only)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ora

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "residual_jacobian_golden.json")))


def _close(a, b, rtol, scale=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return np.max(np.abs(a - b)) <= rtol * max(s, 1e-300)


@pytest.mark.parametrize("idx", range(len(GOLD["residual_jacobian"])))
def test_residuals_and_local_jacobians(idx):
    g = GOLD["residual_jacobian"][idx]
    pose, pt, intr, uv = g["pose"], g["pt"], g["intr"], g["uv"]
    r2, Jpose, Jpt, Jk = ora.reprojection(pose, pt, intr, uv, g["w_repr"])
    r1, Jdpose, Jdpt, Jdk = ora.depth_prior(pose, pt, intr, g["depth"], g["w_unpr"])
    Jplus = ora.se3_dx(pose)
    r = np.concatenate([r2, r1])
    Jloc = np.vstack([Jpose @ Jplus, Jdpose @ Jplus])
    # residuals are differences of O(300 px) quantities: tolerance relative to the
    # pixel magnitude times sqrt(weight) (cancellation), 1e-12 as in north_star
    sw = np.sqrt(g["w_repr"])
    assert _close(r[:2], g["r"][:2], 1e-12, scale=sw * 640.0)
    assert _close(r[2:], g["r"][2:], 1e-12, scale=np.sqrt(g["w_unpr"]) * max(1.0, g["depth"]))
    for row in range(3):
        assert _close(Jloc[row], g["Jpose"][row], 1e-12), (row, Jloc[row], g["Jpose"][row])
    assert _close(np.vstack([Jpt, Jdpt]), g["Jpt"], 1e-12)
    assert _close(np.vstack([Jk, Jdk]), g["Jintr"], 1e-12)


@pytest.mark.parametrize("idx", range(len(GOLD["se3"])))
def test_se3_ops(idx):
    g = GOLD["se3"][idx]
    e = ora.se3_exp(g["delta"])
    # Sophus' small-angle branch (theta < 1e-10, se3.hpp:736-738) uses V = R
    # instead of I + Omega/2: a deliberate O(theta*|upsilon|) deviation from the
    # exact exponential that the restatement must reproduce, not "fix".
    th = np.linalg.norm(g["delta"][3:])
    tol = 1e-14 if th >= 1e-10 else 2.0 * th
    assert _close(e, g["exp"], tol, scale=1.0 + np.max(np.abs(g["exp"])))
    assert _close(ora.se3_mul(g["a"], e), g["a_mul_exp"], tol, scale=4.0)
    assert _close(ora.se3_mul(g["a"], g["b"]), g["a_mul_b"], 1e-14, scale=4.0)
    assert _close(ora.se3_inverse(g["a"]), g["a_inv"], 1e-14, scale=4.0)
    assert _close(ora.se3_act(g["a"], g["p"]), g["a_act_p"], 1e-14, scale=8.0)
    # Sophus invariants: T * T^-1 = I, unit quaternion after product
    ident = ora.se3_mul(g["a"], ora.se3_inverse(g["a"]))
    assert np.max(np.abs(ident - np.array([0, 0, 0, 1, 0, 0, 0.0]))) < 1e-14
    q = ora.se3_mul(g["a"], g["b"])[:4]
    assert abs(np.dot(q, q) - 1.0) < 1e-14


def test_se3_exp_zero_is_identity():
    assert np.array_equal(ora.se3_exp(np.zeros(6)), np.array([0, 0, 0, 1, 0, 0, 0.0]))


@pytest.mark.parametrize("idx", range(len(GOLD["huber"])))
def test_huber(idx):
    g = GOLD["huber"][idx]
    rho = ora.huber(g["a"], g["s"])
    assert np.allclose(rho, g["rho"], rtol=1e-14, atol=0.0)
