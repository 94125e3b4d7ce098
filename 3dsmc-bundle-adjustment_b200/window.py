"""Python mirror of the reference's optimiser entry points
(headers/OptimizationUtils.h:42, 55) over the C-ABI: same names, argument
meaning and in-place mutation contract, with the Ceres Problem/Solve calls
(src/OptimizationUtils.cpp:218-300) replaced by ba_gpu_upload / solve / download.
The compiled drop-in for main.cpp is host/OptimizationUtils_gpu.cpp.
"""
import time
from dataclasses import dataclass

import numpy as np

from . import synthetic
from .solver import GpuSolver, default_options


@dataclass
class CeresGlobalProblem:
    """headers/BundleAdjustmentConfig.h:44-69 (names kept)."""
    HUB_P_REPR: float = 1e-3
    WEIGHT_INTRINSICS: float = 1e-6
    WEIGHT_UNPR: float = 10.0
    HUB_P_UNPR: float = 1e-3
    frame_frequency: int = 10
    window_size: int = 0
    max_num_iterations: int = 75   # options.max_num_iterations (:64)
    eta: float = 1e-6              # options.eta (:65)

    def gpu_options(self, **kw):
        return default_options(HUB_P_REPR=self.HUB_P_REPR, WEIGHT_INTRINSICS=self.WEIGHT_INTRINSICS,
                               WEIGHT_UNPR=self.WEIGHT_UNPR, HUB_P_UNPR=self.HUB_P_UNPR,
                               max_num_iterations=self.max_num_iterations, eta=self.eta, **kw)


def count_constraints(seq, kf_i, kf_f):
    """countConstraints (src/OptimizationUtils.cpp:184-213): observations with depth > 1e-15."""
    a, b = int(seq.kf_ptr[kf_i]), int(seq.kf_ptr[kf_f + 1])
    return int(np.count_nonzero(seq.depth[a:b] > 1e-15))


def window_optimize(global_problem, kf_i, kf_f, seq, intrinsics_initial, intrinsics_optimized, solver=None,
                    return_summary=False, timing=None, **opt_overrides):
    """windowOptimize (src/OptimizationUtils.cpp:215-313).  Mutates seq.pose,
    seq.pt and intrinsics_optimized in place; returns True (as the reference
    always does) unless the GPU solve reports an error, in which case the inputs
    are left untouched and False is returned."""
    t0 = time.perf_counter()
    win = synthetic.window_problem(seq, kf_i, kf_f, intr=intrinsics_optimized, intr_prior=intrinsics_initial)
    own = solver is None
    if own:
        solver = GpuSolver(global_problem.gpu_options(**opt_overrides))
    try:
        t1 = time.perf_counter()
        solver.upload(win.problem)
        t2 = time.perf_counter()
        summary = solver.solve()
        t3 = time.perf_counter()
        pose, pt, intr = solver.download()
        t4 = time.perf_counter()
    except Exception:
        if own:
            solver.close()
        if return_summary:
            raise
        return False
    synthetic.write_back(seq, win, pose, pt)
    intrinsics_optimized[:] = intr
    if timing is not None:  # optional dict: seconds per phase, accumulated over calls
        t5 = time.perf_counter()
        for k, v in (("extract", t1 - t0), ("upload", t2 - t1), ("solve", t3 - t2), ("download", t4 - t3),
                     ("write_back", t5 - t4), ("solve_device", summary.solve_ms * 1e-3)):
            timing[k] = timing.get(k, 0.0) + v
    if own:
        solver.close()
    return (True, summary) if return_summary else True
