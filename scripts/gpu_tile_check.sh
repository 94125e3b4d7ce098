python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "matvec or implicit or tiled or full_size" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python - > gpurun_out/tile_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0,'.')
import ba_b200
from ba_b200 import capi
for cfg in (3,4,5):
    p = ba_b200.synthetic.make_config(cfg)
    for store in (2,3):
        s = ba_b200.GpuSolver(use_depth_prior=0, optimize_intrinsics=0, solver=2, jacobian_store=store, max_num_iterations=3,
                              function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
        s.upload(p)
        flush = cfg < 5
        mv = s.time_kernel(capi.BA_KERNEL_SCHUR_MATVEC, 3, 20, flush)
        p1 = s.time_kernel(capi.BA_KERNEL_SCHUR_PASS1, 3, 20, flush)
        p2 = s.time_kernel(capi.BA_KERNEL_SCHUR_PASS2, 3, 20, flush)
        summ = s.solve()
        print("cfg%d store=%d used=%d matvec %.4f ms pass1 %.4f pass2 %.4f | solve %.2f ms, %d LM its, %d PCG its, cost %.12g" % (cfg, store, s.jacobian_store_used(), mv, p1, p2, summ.solve_ms, summ.num_iterations, summ.total_linear_iters, summ.final_cost), flush=True)
        s.close()
PY
