"""The oracle's SPARSE_SCHUR-equivalent exact step (explicit S in envelope storage + sparse Cholesky: the reference's own
linear_solver_type, headers/BundleAdjustmentConfig.h:62) against its dense Schur step: same matrix, same right-hand side,
other storage and elimination order of the inner sums -> LM traces agree to round-off.  CPU only."""
import numpy as np
import pytest

from helpers import ba_b200, ora, pose_err, to_oracle

syn = ba_b200.synthetic


def _run(p, solver, **kw):
    op = to_oracle(p)
    rc, s, tr = ora.solve(op, ora.default_options(solver=solver, **kw))
    assert rc == 0
    return op, s, tr


@pytest.mark.parametrize("mode", [(1, 1), (0, 0), (1, 0), (0, 1)], ids=["REF", "NS", "DEPTH", "INTR"])
def test_sparse_equals_dense_window(mode):
    p = syn.make_config(1)
    kw = dict(use_depth_prior=mode[0], optimize_intrinsics=mode[1], max_num_iterations=8)
    a, sa, ta = _run(p, 0, **kw)
    b, sb, tb = _run(p, 2, **kw)
    assert sa.num_iterations == sb.num_iterations and sa.termination == sb.termination
    assert [t["step_is_successful"] for t in ta] == [t["step_is_successful"] for t in tb]
    for x, y in zip(ta, tb):
        assert abs(x["cost"] - y["cost"]) <= 1e-10 * abs(x["cost"])
    dt, dr = pose_err(a.pose7, b.pose7)
    assert dt < 1e-8 and dr < 1e-8
    assert np.max(np.abs(a.pt3 - b.pt3)) < 1e-7


@pytest.mark.parametrize("cfg,scale", [(3, 0.05), (4, 0.05)])
def test_sparse_equals_dense_banded(cfg, scale):
    """Sequential co-visibility (banded S): 40 keyframes / 86 loop cameras."""
    p = syn.make_config(cfg, scale=scale)
    kw = dict(use_depth_prior=0, optimize_intrinsics=0, max_num_iterations=6)
    a, sa, ta = _run(p, 0, **kw)
    b, sb, tb = _run(p, 2, **kw)
    assert sa.num_iterations == sb.num_iterations
    assert [t["step_is_successful"] for t in ta] == [t["step_is_successful"] for t in tb]
    assert abs(sa.final_cost - sb.final_cost) <= 1e-9 * sa.final_cost
    dt, dr = pose_err(a.pose7, b.pose7)
    assert dt < 1e-7 and dr < 1e-7


def test_sparse_fixed_camera_in_the_middle_and_none():
    p = syn.make_config(3, scale=0.03)
    for fixed in (5, -1):
        q = p.copy()
        q.fixed_cam = fixed
        kw = dict(use_depth_prior=0, optimize_intrinsics=0, max_num_iterations=4)
        a, sa, _ = _run(q, 0, **kw)
        b, sb, _ = _run(q, 2, **kw)
        assert abs(sa.final_cost - sb.final_cost) <= 1e-9 * sa.final_cost
        assert pose_err(a.pose7, b.pose7)[0] < 1e-7
