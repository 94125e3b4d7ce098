#!/usr/bin/env python
"""CPU arm of BASELINE.md section 6: the oracle (reference's SPARSE_SCHUR setting) on 1 thread and on all host cores, every
workload, a bounded number of LM iterations actually run.  python scripts/cpu_baseline_threads.py [workloads...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ba_b200  # noqa: E402

syn = ba_b200.synthetic
names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg3ref", "cfg4", "cfg5"]
for name in names:
    wl = bench.WORKLOADS[name]
    if name == "cfg3ref":
        c3 = syn.CONFIGS[3]
        seq = syn.make_tum_sequence(c3["n_kf"], c3["n_lm"], c3["n_obs"], syn.SEED_BASE + 3)
        pr = syn.window_problem(seq, 0, c3["n_kf"] - 1).problem
    else:
        pr = syn.make_config(wl["cfg"])
        if wl["cfg"] == 2:
            pr = syn.window_problem(pr, 0, 19).problem
    big = pr.n_obs > 100000
    for threads in (1, os.cpu_count() or 1):
        iters = (1 if pr.n_obs > 2000000 and threads == 1 else 3) if big else 20
        r = bench.cpu_reference_run(pr, wl["mode"], iters, threads, warmup=0 if big else 3)
        print(json.dumps({"workload": name, "threads": threads, "lm_iterations": r["lm_iterations"], "seconds": r["seconds"],
                          "lm_iterations_per_s": r["value"], "jacobian_obs_per_s": r["jacobian_obs_per_s"],
                          "seconds_linearize": r["seconds_linearize"], "seconds_linear_solve": r["seconds_linear_solve"]}), flush=True)
