#!/usr/bin/env python
"""Profiling target: one fixed-size LM solve through the C-ABI, nothing else.
Run plain first, then under ncu with the same command line (B200_PROFILING.md).
  python profiles/profile_target.py [cfg] [lm_iterations] [max_pcg] [solver: 0 auto, 2 implicit, 3 block-sparse] [jacobian_store]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ba_b200  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_pcg = int(sys.argv[3]) if len(sys.argv) > 3 else 500
solver = int(sys.argv[4]) if len(sys.argv) > 4 else 0
store = int(sys.argv[5]) if len(sys.argv) > 5 else 0
p = ba_b200.synthetic.make_config(cfg)
if cfg == 2:
    p = ba_b200.synthetic.window_problem(p, 0, 19).problem
ns = p.depth is None
s = ba_b200.GpuSolver(max_num_iterations=iters, use_depth_prior=0 if ns else 1, optimize_intrinsics=0 if ns else 1,
                      max_linear_solver_iterations=max_pcg, function_tolerance=0.0, parameter_tolerance=0.0,
                      gradient_tolerance=0.0, solver=solver, jacobian_store=store)
s.upload(p)
summ = s.solve()
print("cfg%d: %d LM iterations, %d PCG iterations, %d launches, %.3f ms, cost %.6g -> %.6g"
      % (cfg, summ.num_iterations, summ.total_linear_iters, summ.kernel_launches, summ.solve_ms, summ.initial_cost,
         summ.final_cost))
s.close()
