"""B200-native drop-in for the Ceres Problem/Solve path of the reference's
windowed bundle adjustment (src/OptimizationUtils.cpp:215-313).

The directory name is the repository's (not a Python identifier): import it
through the `ba_b200` shim at the repository root.
"""
from . import capi, evaluation, hostlib, se3, synthetic, trajectory  # noqa: F401
from .solver import BAError, BAProblem, GpuSolver, comm_unique_id, default_options  # noqa: F401
from .window import window_optimize, count_constraints, CeresGlobalProblem  # noqa: F401
