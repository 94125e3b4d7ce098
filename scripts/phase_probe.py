#!/usr/bin/env python
"""Phase times (ms per LM iteration) of one fixed-iteration solve: python scripts/phase_probe.py [cfg] [iters] [solver]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ba_b200  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
solver = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = ba_b200.synthetic.make_config(cfg)
s = ba_b200.GpuSolver(max_num_iterations=iters, use_depth_prior=0, optimize_intrinsics=0, function_tolerance=0.0, parameter_tolerance=0.0,
                      gradient_tolerance=0.0, solver=solver)
for rep in range(2):
    s.upload(p)
    summ = s.solve()
ph = s.phase_times()
print("cfg%d solver %d: %.3f ms/it, cost %.12g | " % (cfg, summ.solver_used, summ.solve_ms / summ.num_iterations, summ.final_cost)
      + " ".join("%s %.3f" % (k[:14], v / summ.num_iterations) for k, v in ph.items()))
s.close()
