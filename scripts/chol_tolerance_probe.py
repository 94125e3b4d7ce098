#!/usr/bin/env python
"""Measured GPU-vs-oracle differences of the exact dense step (blocked Cholesky, REF cost) at the sizes of
tests/test_gpu_parity.py::test_solve_explicit_blocked_cholesky_ref: what tolerance the lock step really holds."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import ba_b200, mode_opts, ora, pose_err, to_oracle
syn = ba_b200.synthetic
for n_kf in (28, 33, 43, 60, 200):
    seq = syn.make_tum_sequence(n_kf, 30 * n_kf, 180 * n_kf, seed=21)
    p = syn.window_problem(seq, 0, n_kf - 1).problem
    g, o = mode_opts("REF", solver=0, max_num_iterations=6)
    s = ba_b200.GpuSolver(**g)
    s.upload(p)
    summ = s.solve()
    pose, pt, intr = s.download()
    tr = s.trace()
    s.close()
    op = to_oracle(p)
    rc, osum, otr = ora.solve(op, ora.default_options(num_threads=os.cpu_count(), **o))
    dtr = max(abs(a["cost"] - b["cost"]) / abs(b["cost"]) for a, b in zip(tr, otr))
    print("n_kf %d (n = %d): final cost rel diff %.2e, worst per-iteration cost rel diff %.2e, pose diff %.1e m / %.1e rad"
          % (n_kf, summ.reduced_dim, abs(summ.final_cost - osum.final_cost) / osum.final_cost, dtr, *pose_err(pose, op.pose7)))
