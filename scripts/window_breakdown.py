#!/usr/bin/env python
"""cfg 2 sliding windows through one reused context: wall time of upload / solve / download per window
(host extraction excluded), for K = 10 LM iterations per window as the BASELINE config states."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ba_b200  # noqa: E402

syn = ba_b200.synthetic
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
seq = syn.make_config(2)
s = ba_b200.GpuSolver(max_num_iterations=K, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
wins = [syn.window_problem(seq, n - 20, n - 1).problem for n in range(20, seq.pose.shape[0] + 1, 10)]
for rep in range(2):
    tu = ts = td = 0.0
    gpu_ms = 0.0
    for p in wins:
        t0 = time.perf_counter()
        s.upload(p)
        t1 = time.perf_counter()
        sm = s.solve()
        t2 = time.perf_counter()
        s.download()
        t3 = time.perf_counter()
        tu += t1 - t0
        ts += t2 - t1
        td += t3 - t2
        gpu_ms += sm.solve_ms
    n = len(wins)
    print("pass %d: %d windows x %d LM iterations: upload %.3f ms, solve %.3f ms (device %.3f ms), download %.3f ms per window; "
          "%.0f windows/s, %.0f LM it/s end to end" % (rep, n, K, 1e3 * tu / n, 1e3 * ts / n, gpu_ms / n, 1e3 * td / n,
                                                     n / (tu + ts + td), n * K / (tu + ts + td)), flush=True)
s.close()
