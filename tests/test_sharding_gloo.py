"""World-size-2 CPU test (gloo) of the sharding logic of the multi-GPU path:
the reduced-system product is the all-reduced sum of every rank's partial
product over its own point range, computed here with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ba_b200, ora, to_oracle


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = ba_b200.synthetic.make_config(4, scale=0.01)
    shard, ids = ba_b200.synthetic.shard_points(full, rank, world)
    o = ora.default_options(use_depth_prior=0, optimize_intrinsics=0, solver=1)
    x = np.random.default_rng(5).normal(size=6 * full.n_cam)
    # partial product over this rank's points; the U + D^2 term is added once (rank 0)
    lo, hi = int(ids[0]), int(ids[-1]) + 1
    y = ora.schur_matvec(to_oracle(full), o, 1e4, x, pt_begin=lo, pt_end=hi, include_diag=1 if rank == 0 else 0)
    t = torch.from_numpy(y)
    dist.all_reduce(t)
    # scalar reductions of the LM controller: cost is a sum over shards with the GLOBAL weight 1/N
    o_sh = ora.default_options(use_depth_prior=0, optimize_intrinsics=0, n_obs_total=full.n_obs)
    c = torch.tensor([ora.evaluate(to_oracle(shard), o_sh, want_jac=False)["cost"]], dtype=torch.float64)
    dist.all_reduce(c)
    if rank == 0:
        yref = ora.schur_matvec(to_oracle(full), o, 1e4, x)
        cref = ora.evaluate(to_oracle(full), ora.default_options(use_depth_prior=0, optimize_intrinsics=0), want_jac=False)["cost"]
        q.put((float(np.max(np.abs(t.numpy() - yref)) / np.max(np.abs(yref))), abs(c.item() - cref) / cref))
    dist.barrier()
    dist.destroy_process_group()


def test_partial_products_sum_to_full_product():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29531 + (os.getpid() % 50)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err_y, err_c = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err_y < 1e-12 and err_c < 1e-13
