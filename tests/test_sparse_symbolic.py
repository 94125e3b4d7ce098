"""Symbolic phase of the sparse Cholesky of S (host code of libba_gpu.so, no GPU needed) + a numpy emulation of the
numeric phase that consumes the same structures: the supernodal multifrontal solve must equal a dense solve."""
import numpy as np
import pytest

from helpers import ba_b200
from spchol_emulation import K0, M, NB, PARENT, Symbolic, covisibility_blocks, factor_solve

syn = ba_b200.synthetic


def _random_spd(n_cam, bi, bj, rng):
    """Block matrix with the given upper pattern, made diagonally dominant."""
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
    dsq = np.full((n_cam, 6), shift)
    return A + shift * np.eye(6 * n_cam), blocks, dsq


def _check(n_cam, bi, bj, seed, **kw):
    rng = np.random.default_rng(seed)
    sym = Symbolic(n_cam, bi, bj, **kw)
    # structure invariants
    assert sorted(sym.perm.tolist()) == list(range(n_cam))
    assert int(sym.node[:, M].sum()) == n_cam
    for id_, N in enumerate(sym.node):
        assert (N[M] + N[NB]) * (N[M] + 1) <= kw.get("cap", 760)
        if N[PARENT] >= 0:
            assert N[PARENT] > id_ and sym.node[N[PARENT]][5] > N[5]
    A, blocks, dsq = _random_spd(n_cam, bi, bj, rng)
    b = rng.normal(size=(n_cam, 6))
    y = factor_solve(sym, blocks, dsq, b)
    yref = np.linalg.solve(A, b.reshape(-1)).reshape(-1, 6)
    assert np.max(np.abs(y - yref)) <= 1e-9 * np.max(np.abs(yref))
    return sym


@pytest.mark.parametrize("cfg,scale,leaf", [(3, 0.05, 8), (4, 0.05, 8), (4, 0.1, 16), (3, 0.1, 12)])
def test_banded_covisibility(cfg, scale, leaf):
    p = syn.make_config(cfg, scale=scale)
    bi, bj = covisibility_blocks(p.cam_idx, p.pt_idx, p.n_cam, p.fixed_cam)
    sym = _check(p.n_cam, bi, bj, 3, leaf=leaf, cap=300)
    assert sym.n_levels >= 3  # the dissection really branches


def test_small_capacity_splits_supernodes():
    p = syn.make_config(4, scale=0.05)
    bi, bj = covisibility_blocks(p.cam_idx, p.pt_idx, p.n_cam, p.fixed_cam)
    a = _check(p.n_cam, bi, bj, 4, leaf=16, cap=760)
    b = _check(p.n_cam, bi, bj, 4, leaf=16, cap=120, max_own=4)
    assert b.n_nodes > a.n_nodes


def test_loop_closure_and_disconnected_parts():
    """Not a band: a loop closure (first cameras see the last ones), an isolated camera, and two unconnected halves."""
    rng = np.random.default_rng(5)
    n = 60
    pairs = set((i, i) for i in range(n) if i != 17)          # camera 17 has no diagonal block (damping only)
    for i in range(n):
        for d in (1, 2, 3):
            if i + d < n and i != 17 and i + d != 17 and not (i < 30 <= i + d):
                pairs.add((i, i + d))
    pairs.add((0, 29)); pairs.add((1, 28)); pairs.add((31, 59))
    pairs = sorted(pairs)
    bi = np.array([a for a, _ in pairs], dtype=np.int32)
    bj = np.array([b for _, b in pairs], dtype=np.int32)
    _check(n, bi, bj, 6, leaf=6, cap=760)


def test_dense_window_is_one_chain():
    n = 7
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    bi = np.array([a for a, _ in pairs], dtype=np.int32)
    bj = np.array([b for _, b in pairs], dtype=np.int32)
    sym = _check(n, bi, bj, 7, leaf=32, cap=760)
    assert sym.n_nodes == 1 and sym.node[0][M] == 7 and sym.node[0][NB] == 0


def test_unsupported_when_a_front_cannot_fit():
    n = 40
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    bi = np.array([a for a, _ in pairs], dtype=np.int32)
    bj = np.array([b for _, b in pairs], dtype=np.int32)
    with pytest.raises(RuntimeError):
        Symbolic(n, bi, bj, leaf=4, cap=20, max_own=64)


@pytest.mark.parametrize("cfg,scale,leaf", [(5, 0.1, 32), (4, 0.5, 32), (3, 0.25, 16)])
def test_subtree_partition(cfg, scale, leaf):
    """Subtree-to-rank partition of the multi-GPU factorisation: parts are unions of whole subtrees, the top part is closed
    under `parent`, work is balanced on the sequential co-visibility of the synthetic configurations, and consecutive
    parts own consecutive stretches of the camera sequence."""
    p = syn.make_config(cfg, scale=scale)
    bi, bj = covisibility_blocks(p.cam_idx, p.pt_idx, p.n_cam, p.fixed_cam)
    sym = Symbolic(p.n_cam, bi, bj, leaf=leaf, cap=770, max_own=24)
    for parts, (got, part, work) in sym.partitions.items():
        assert got in (1, parts)
        if got == 1:
            assert (part == 0).all()
            continue
        assert part.min() == -1 and part.max() == parts - 1
        for id_, N in enumerate(sym.node):
            par = N[PARENT]
            if part[id_] == -1:
                assert par < 0 or part[par] == -1          # the top part is closed upwards
            elif par >= 0:
                assert part[par] in (-1, part[id_])        # subtrees are never split between ranks
        total, top, heaviest = work
        assert top <= 0.4 * total           # (opening stops once the top part passes 30 %)
        if sym.n_nodes >= 40 * parts:       # enough subtrees to balance
            assert heaviest <= 1.25 * (total - top) / parts
        # consecutive parts <-> consecutive cameras (median camera index of a part's own cameras is increasing)
        med = []
        for r in range(parts):
            cams = np.concatenate([sym.perm[N[K0]:N[K0] + N[M]] for id_, N in enumerate(sym.node) if part[id_] == r] or [np.zeros(0)])
            if len(cams):
                med.append(np.median(cams))
        assert med == sorted(med)


def _banded_blocks(n_cam, widths, rng, gaps=()):
    """Upper block pattern of a camera sequence whose co-visibility reaches `widths[i]` cameras ahead of camera i; cameras in
    `gaps` see nothing ahead (the sequence falls apart into independent pieces there)."""
    keys = set()
    for i in range(n_cam):
        keys.add((i, i))
        if i in gaps:
            continue
        for d in range(1, widths[i] + 1):
            if i + d < n_cam and not any(i < g_ < i + d for g_ in gaps) and rng.random() < 0.85:
                keys.add((i, i + d))
    k = sorted(keys)
    return np.array([a for a, _ in k], dtype=np.int32), np.array([b_ for _, b_ in k], dtype=np.int32)


@pytest.mark.parametrize("seed,n_cam,maxw,gaps", [(1, 90, 4, ()), (2, 140, 7, ()), (3, 120, 3, (40, 41, 85)), (4, 60, 10, ()),
                                                  (5, 200, 2, (99,))])
def test_distributed_protocol_on_varied_structures(seed, n_cam, maxw, gaps):
    """Partition + exchange rule + subtree / top split (simulated ranks, numpy emulation on the C-ABI's symbolic tables) on band
    widths that vary along the sequence and on sequences that fall apart into independent pieces (several tree roots)."""
    from spchol_emulation import distributed_solve_inprocess
    rng = np.random.default_rng(seed)
    widths = rng.integers(1, maxw + 1, size=n_cam)
    bi, bj = _banded_blocks(n_cam, widths, rng, set(gaps))
    sym = Symbolic(n_cam, bi, bj, leaf=5, cap=300, max_own=8)
    A, blocks, dsq = _random_spd(n_cam, bi, bj, rng)
    b = rng.normal(size=(n_cam, 6))
    yref = np.linalg.solve(A, b.reshape(-1)).reshape(-1, 6)
    done = 0
    for world, (got, part, _) in sym.partitions.items():
        if got != world:
            continue
        # every rank's points touch the blocks of its stretch of the sequence (with overlap); stragglers go to one rank
        touch = np.zeros((world, len(bi)), dtype=bool)
        for r in range(world):
            lo, hi = r * n_cam // world - maxw, (r + 1) * n_cam // world + maxw
            touch[r] = (bi >= lo) & (bi < hi) & (bj >= lo) & (bj < hi)
        none = ~touch.any(axis=0)
        touch[np.minimum(bi[none] * world // n_cam, world - 1), np.nonzero(none)[0]] = True
        y, n_x = distributed_solve_inprocess(sym, part, world, blocks, dsq, b, touch)
        assert np.max(np.abs(y - yref)) <= 1e-9 * np.max(np.abs(yref)), (world, n_x)
        done += 1
    assert done >= 1
