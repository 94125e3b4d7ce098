"""World-size-2 / 3 CPU test (gloo) of the DISTRIBUTED sparse Cholesky of the multi-GPU path: the host logic
-- subtree-to-rank partition (csrc/ba_sparse_symbolic.h through the C-ABI), the rule that decides which blocks of S are summed over
ranks, the exchange of the subtree roots' update matrices, the replicated top part and the assembly of the step -- executed by a
numpy emulation of the kernels that consumes the same symbolic tables.  Blocks a rank must never read are poisoned with NaN."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ba_b200
from spchol_emulation import NB, PARENT, Multifrontal, Symbolic, block_owner, covisibility_blocks


def _spd(n_cam, bi, bj, rng):
    A = np.zeros((n_cam * 6, n_cam * 6))
    blocks = np.zeros((len(bi), 6, 6))
    for b, (i, j) in enumerate(zip(bi, bj)):
        B = rng.normal(size=(6, 6))
        if i == j:
            B = B + B.T
        blocks[b] = B
        A[6 * i:6 * i + 6, 6 * j:6 * j + 6] = B
        A[6 * j:6 * j + 6, 6 * i:6 * i + 6] = B.T
    shift = np.abs(A).sum(axis=1).max() + 1.0
    return A + shift * np.eye(6 * n_cam), blocks, np.full((n_cam, 6), shift)


def _allreduce(a, op=dist.ReduceOp.SUM):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    dist.all_reduce(t, op=op)
    return t.numpy().reshape(a.shape)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ba_b200.synthetic.make_config(4, scale=0.1)
    bi, bj = covisibility_blocks(p.cam_idx, p.pt_idx, p.n_cam, p.fixed_cam)
    sym = Symbolic(p.n_cam, bi, bj, leaf=6, cap=770, max_own=24)   # (small leaves: a deep tree at this size)
    got, part, _ = sym.partitions[world]
    assert got == world, "the tree of the test problem must split into %d parts" % world
    A, blocks, dsq = _spd(p.n_cam, bi, bj, np.random.default_rng(11))      # same on every rank
    b = np.random.default_rng(12).normal(size=(p.n_cam, 6))
    # ---- what this rank's points contribute to S: camera ranges with an overlap (tracks cross the shard borders); the
    #      contributions of all ranks sum to `blocks`
    n = p.n_cam
    touch = np.zeros((world, len(bi)), dtype=bool)
    for r in range(world):
        lo, hi = r * n // world - 8, (r + 1) * n // world + 8
        touch[r] = (bi >= lo) & (bi < hi) & (bj >= lo) & (bj < hi)
    none = ~touch.any(axis=0)
    touch[np.minimum(bi[none] * world // n, world - 1), np.nonzero(none)[0]] = True
    local = np.where(touch[rank][:, None, None], blocks / touch.sum(axis=0)[:, None, None], 0.0)
    # ---- sparse exchange of S: blocks some rank contributes to without assembling them (build_spchol's rule)
    owner = block_owner(sym, part)
    mask = _allreduce((touch[rank] & (owner != rank)).astype(np.float64), dist.ReduceOp.MAX) != 0.0
    S = local.copy()
    S[mask] = _allreduce(local[mask])
    needed = (owner == rank) | (owner == -1)
    S[~needed] = np.nan                                   # never assembled by this rank
    assert np.allclose(S[needed], blocks[needed], rtol=0, atol=1e-12)
    # ---- phase A: own subtrees; exchange of the subtree roots' update matrices and right-hand-side updates
    mf = Multifrontal(sym, S, dsq, b)
    mf.factor(mf.nodes_bottom_up(lambda id_: part[id_] == rank))
    for id_, N in enumerate(sym.node):
        if part[id_] >= 0 and N[PARENT] >= 0 and part[N[PARENT]] == -1:
            nb = N[NB]
            own = part[id_] == rank
            mf.Us[id_] = _allreduce(mf.Us[id_] if own else np.zeros((nb, nb, 6, 6)))
            mf.rus[id_] = _allreduce(mf.rus[id_] if own else np.zeros((nb, 6)))
    # ---- phase B: top part on every rank, backward substitution of the top part and of the own subtrees
    top = mf.nodes_bottom_up(lambda id_: part[id_] == -1)
    mf.factor(top)
    mf.solve(top[::-1])
    mf.solve(mf.nodes_bottom_up(lambda id_: part[id_] == rank)[::-1])
    # ---- the step: own cameras everywhere, the top part's cameras from rank 0 only; one sum
    keep = np.zeros(n, dtype=bool)
    for id_, N in enumerate(sym.node):
        if part[id_] == rank or (part[id_] == -1 and rank == 0):
            keep[N[0]:N[0] + N[1]] = True
    assert not np.isnan(mf.ypos[keep]).any()
    mf.ypos[~keep] = 0.0
    mf.ypos = _allreduce(mf.ypos)
    y = mf.y()
    yref = np.linalg.solve(A, b.reshape(-1)).reshape(-1, 6)
    err = float(np.max(np.abs(y - yref)) / np.max(np.abs(yref)))
    spread = float(np.max(np.abs(_allreduce(y, dist.ReduceOp.MAX) - _allreduce(y, dist.ReduceOp.MIN))))
    if rank == 0:
        q.put((err, spread, int(mask.sum()), len(bi), int((part == -1).sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_factorisation_matches_dense_solve(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29631 + (os.getpid() % 50) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    err, spread, n_x, n_blk, n_top = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert err < 1e-9 and spread == 0.0
    assert 0 < n_x < n_blk // 2 and n_top >= 1     # the exchange really is sparse, the top part exists
