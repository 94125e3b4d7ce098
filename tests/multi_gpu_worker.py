"""torchrun worker: point-sharded solve on WORLD_SIZE GPUs vs the same solve on
one GPU (rank 0 checks).  Launched by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_worker.py [cfg] [scale] [iters] [solver: 0 auto, 2 implicit, 3 block-sparse PCG, 4 sparse Cholesky]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ba_b200  # noqa: E402


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    solver = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    persistent = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    full = ba_b200.synthetic.make_config(cfg, scale=scale)
    shard, ids = ba_b200.synthetic.shard_points(full, rank, world)
    opts = dict(use_depth_prior=0, optimize_intrinsics=0, solver=solver, max_num_iterations=iters, device=local, persistent_pcg=persistent)
    s = ba_b200.GpuSolver(n_obs_total=full.n_obs, **opts)
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idbuf.copy_(torch.frombuffer(bytearray(ba_b200.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    s.comm_init(idbuf.cpu().numpy().tobytes(), rank, world)
    s.upload(shard)
    if len(sys.argv) > 6 and sys.argv[6] == "distributed":
        info = s.spchol_info()
        assert info.get("distributed") == 1 and info["parts"] == world, "the distributed factorisation is not in force: %r" % (info,)
    summ = s.solve()
    pose, pt, _ = s.download()
    tr = s.trace()
    # every rank must hold the same poses / trace
    t = torch.from_numpy(pose.copy()).cuda()
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    assert torch.equal(tmax, tmin), "poses differ between ranks"
    # gather the sharded points on rank 0
    allpt = torch.zeros((full.n_pt, 3), dtype=torch.float64, device="cuda")
    allpt[ids] = torch.from_numpy(pt).cuda()
    dist.all_reduce(allpt)
    ok = True
    if rank == 0:
        s1 = ba_b200.GpuSolver(**opts)
        s1.upload(full)
        sum1 = s1.solve()
        pose1, pt1, _ = s1.download()
        tr1 = s1.trace()
        dcost = abs(summ.final_cost - sum1.final_cost) / sum1.final_cost
        dpose = float(np.max(np.abs(pose - pose1)))
        dpt = float(np.max(np.abs(allpt.cpu().numpy() - pt1)))
        print("world=%d solver=%d cfg%d scale=%g: LM its %d/%d, PCG its %d/%d, final cost %.12g vs %.12g (rel %.2e), max|dpose| %.2e, max|dpt| %.2e"
              % (world, solver, cfg, scale, summ.num_iterations, sum1.num_iterations, summ.total_linear_iters, sum1.total_linear_iters,
                 summ.final_cost, sum1.final_cost, dcost, dpose, dpt))
        # well-conditioned TUM-shaped problems (cfg3): strict; the ill-conditioned loop (cfg4/5) amplifies the
        # different summation order of the sharded reduction (see test_solve_implicit_pcg_ill_conditioned)
        # ... an exact step (sparse Cholesky, also what AUTO resolves to) has no such amplification: strict everywhere
        strict = cfg <= 3 or summ.solver_used == 4
        ok = (summ.num_iterations == sum1.num_iterations and dcost < (1e-8 if strict else 1e-5)
              and dpose < (1e-6 if strict else 3e-3) and dpt < (1e-5 if strict else 1e-1)
              and [t_["step_is_successful"] for t_ in tr] == [t_["step_is_successful"] for t_ in tr1])
        print("MULTI_GPU_PARITY", "OK" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
