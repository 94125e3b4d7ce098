python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python - > gpurun_out/sparse_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0,'.')
import ba_b200
from ba_b200 import capi
for cfg in (3,4,5):
    p = ba_b200.synthetic.make_config(cfg)
    for solver, pp in ((2,0),(3,0),(3,1)):
        s = ba_b200.GpuSolver(use_depth_prior=0, optimize_intrinsics=0, solver=solver, max_num_iterations=5, persistent_pcg=pp,
                              function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
        t0=time.time(); s.upload(p); t_up=time.time()-t0
        flush = cfg < 5
        mv = s.time_kernel(capi.BA_KERNEL_SCHUR_MATVEC, 3, 20, flush)
        summ = s.solve()
        tr = s.trace()
        print("pp=%d " % pp, end=""); print("cfg%d solver=%d upload %.3fs matvec %.4f ms | solve %.2f ms, %d LM its, %d PCG its, cost %.12g launches %d" % (cfg, solver, t_up, mv, summ.solve_ms, summ.num_iterations, summ.total_linear_iters, summ.final_cost, summ.kernel_launches), flush=True)
        print("   pcg per it", [t["linear_iters"] for t in tr], flush=True)
        s.close()
PY
