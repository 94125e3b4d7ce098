"""CPU-side checks of the boundary: the C-ABI library loads and exports every
symbol include/ba_gpu.h declares; create fails loudly without a GPU (no CPU
fallback); host-side SE3 mirror agrees with the oracle's Sophus restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ba_b200, ora


def _declared():
    src = open(ba_b200.capi.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_(?:gpu|sparse|store)_\w+)\s*\(", src)))


def test_header_symbols_exported():
    lib = ba_b200.capi.load()
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(ba_b200.capi.EXPORTS) == names


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the C compiler's layout of include/ba_gpu.h."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ba_gpu.h"\nint main(){printf("%zu %zu %zu\\n",'
                   'sizeof(ba_gpu_options),sizeof(ba_gpu_iter),sizeof(ba_gpu_summary));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", str(__import__("os").path.dirname(ba_b200.capi.HEADER_PATH)), str(src), "-o", str(exe)])
    so, si, ss = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert C.sizeof(ba_b200.capi.Options) == so
    assert C.sizeof(ba_b200.capi.IterRecord) == si
    assert C.sizeof(ba_b200.capi.Summary) == ss
    o = ba_b200.default_options()
    assert o.HUB_P_REPR == 1e-3 and o.WEIGHT_UNPR == 10.0 and o.max_num_iterations == 75
    assert o.initial_trust_region_radius == 1e4 and o.residual_reset_period == 10 and o.poll_interval == 10


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ba_b200.BAError) as e:
        ba_b200.GpuSolver()
    assert e.value.code == ba_b200.capi.BA_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_host_se3_mirror_matches_oracle():
    rng = np.random.default_rng(3)
    for _ in range(50):
        a = ba_b200.se3.exp(rng.normal(size=6))
        b = ba_b200.se3.exp(rng.normal(size=6) * 0.3)
        p = rng.normal(size=3)
        assert np.max(np.abs(ba_b200.se3.mul(a, b) - ora.se3_mul(a, b))) < 1e-15
        assert np.max(np.abs(ba_b200.se3.inverse(a) - ora.se3_inverse(a))) < 1e-15
        assert np.max(np.abs(ba_b200.se3.act(a, p) - ora.se3_act(a, p))) < 1e-15
        d = rng.normal(size=6) * 10.0 ** rng.uniform(-12, 0)
        assert np.max(np.abs(ba_b200.se3.exp(d) - ora.se3_exp(d))) < 1e-15


def test_synthetic_configs_shapes():
    p1 = ba_b200.synthetic.make_config(1)
    assert (p1.n_cam, p1.n_pt, p1.n_obs) == (7, 500, 3000)
    assert np.all(np.diff(p1.cam_idx) >= 0)
    # points are numbered by first appearance in the canonical order
    first = np.unique(p1.pt_idx, return_index=True)[1]
    assert np.all(np.diff(first) > 0)
    p4 = ba_b200.synthetic.make_config(4, scale=0.01)
    assert p4.depth is None and np.all(np.diff(p4.cam_idx) >= 0)
    # deterministic
    q1 = ba_b200.synthetic.make_config(1)
    assert np.array_equal(p1.uv2, q1.uv2) and np.array_equal(p1.pt_idx, q1.pt_idx)


def test_window_frame_change_roundtrip():
    seq = ba_b200.synthetic.make_tum_sequence(30, 600, 3600, seed=11)
    pose0, pt0 = seq.pose.copy(), seq.pt.copy()
    win = ba_b200.synthetic.window_problem(seq, 5, 24)
    assert np.max(np.abs(win.problem.pose7[0] - np.array([0, 0, 0, 1, 0, 0, 0.0]))) < 1e-15
    assert ba_b200.count_constraints(seq, 5, 24) == win.problem.n_obs
    ba_b200.synthetic.write_back(seq, win, win.problem.pose7, win.problem.pt3)
    assert np.max(np.abs(seq.pose - pose0)) < 1e-13 and np.max(np.abs(seq.pt - pt0)) < 1e-13


def test_shard_points_partition():
    p = ba_b200.synthetic.make_config(4, scale=0.02)
    tot = 0
    for r in range(4):
        q, ids = ba_b200.synthetic.shard_points(p, r, 4)
        tot += q.n_obs
        assert q.n_cam == p.n_cam and np.all(np.diff(q.cam_idx) >= 0)
        assert np.allclose(q.pt3, p.pt3[ids])
    assert tot == p.n_obs


def test_pdl_kernels_wait_before_touching_memory():
    """Kernels of the windowed LM iteration are launched with programmatic stream serialisation (ba_gpu.cu: LAUNCH with
    ctx->pdl), which is only correct if every one of them executes griddepcontrol.wait before its first memory access:
    the first statement of each kernel in ba_kernels.cuh / ba_kernels_chol.cuh must be gate_open() / pdl_wait() (or a
    d_k_* helper that starts with gate_open).  Upload-time index kernels are exempt: they are never launched with PDL."""
    import re
    src_dir = os.path.join(os.path.dirname(ba_b200.capi.LIB_PATH), "csrc")
    exempt = {"k_index_count", "k_exclusive_scan", "k_index_fill", "k_index_sort", "k_index_gather", "k_item_count",
              "k_item_fill", "k_iota", "k_fill", "k_fill_i32"}
    helpers_ok = True
    bad = []
    for f in ("ba_kernels.cuh", "ba_kernels_chol.cuh"):
        s = open(os.path.join(src_dir, f)).read()
        for m in re.finditer(r"__global__[^{;]*?\b(k\w+)\s*\(([^{;]*?)\)\s*\{", s, re.S):
            first = s[m.end():m.end() + 600].split(";")[0]
            if m.group(1) in exempt:
                continue
            if not ("gate_open" in first or "pdl_wait" in first or "d_k_" in first):
                bad.append((f, m.group(1)))
        for m in re.finditer(r"void (d_k_\w+)\s*\(([^{;]*?)\)\s*\{", s, re.S):
            helpers_ok &= "gate_open" in s[m.end():m.end() + 300].split(";")[0]
    assert not bad, bad
    assert helpers_ok
    k = open(os.path.join(src_dir, "ba_kernels.cuh")).read()
    body = k[k.index("bool gate_open("):]
    assert body.index("pdl_wait()") < body.index("st->done")
