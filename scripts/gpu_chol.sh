python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blocked or explicit" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python - > gpurun_out/chol_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0,'.')
import ba_b200
syn = ba_b200.synthetic
for n_kf, n_lm, n_obs in ((200, 15000, 100000), (400, 30000, 200000), (800, 60000, 400000)):
    seq = syn.make_tum_sequence(n_kf, n_lm, n_obs, seed=3)
    p = syn.window_problem(seq, 0, n_kf - 1).problem
    s = ba_b200.GpuSolver(max_num_iterations=5, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    t0 = time.time(); s.upload(p); tu = time.time() - t0
    summ = s.solve()
    print("REF global BA %d keyframes: n_red %d, upload %.2f s, %d LM its in %.1f ms (%.2f ms/it), cost %.6g -> %.6g, launches %d" % (
        n_kf, summ.reduced_dim, tu, summ.num_iterations, summ.solve_ms, summ.solve_ms / max(1, summ.num_iterations), summ.initial_cost, summ.final_cost, summ.kernel_launches), flush=True)
    s.close()
PY
