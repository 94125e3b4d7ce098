"""Deterministic synthetic bundle-adjustment problems shaped like the inputs of
the reference's windowOptimize (BASELINE.json configs 1-5).

TUM-shaped sequences mirror what the reference's front end produces
(src/main.cpp:25-82, src/Map3D.cpp:29-97): K = (525,525,319.5,239.5) on 640x480
(Data/ros_default_intrinsics.txt:1, headers/VirtualSensor.h:44-51), pixels
rounded to float (cv::KeyPoint), depth quantised to 1/5000 m in float
(headers/VirtualSensor.h:97-102), landmark tracks that are runs of consecutive
keyframes, landmarks initialised by back-projection from the first observing
keyframe.  Observation order inside a keyframe is a seeded shuffle (the
reference iterates an unordered_map there, src/OptimizationUtils.cpp:257).
There is no network and no dataset in the image; every number here is synthetic.
"""
from dataclasses import dataclass

import numpy as np

from . import se3
from .solver import BAProblem

K_TUM = np.array([525.0, 525.0, 319.5, 239.5])
SEED_BASE = 0xBA0000


def rotmat_to_quat(R):
    """(…,3,3) proper rotations -> (…,4) quaternions (x,y,z,w), w >= 0."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    w = np.sqrt(np.maximum(0.0, 1.0 + m00 + m11 + m22)) / 2.0
    x = np.sqrt(np.maximum(0.0, 1.0 + m00 - m11 - m22)) / 2.0
    y = np.sqrt(np.maximum(0.0, 1.0 - m00 + m11 - m22)) / 2.0
    z = np.sqrt(np.maximum(0.0, 1.0 - m00 - m11 + m22)) / 2.0
    x = np.copysign(x, R[..., 2, 1] - R[..., 1, 2])
    y = np.copysign(y, R[..., 0, 2] - R[..., 2, 0])
    z = np.copysign(z, R[..., 1, 0] - R[..., 0, 1])
    q = np.stack([x, y, z, w], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def camera_point(pose7, p):
    """p_C = R^T (p - t) with the unit quaternion of pose7 (camera->world)."""
    qc = pose7[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    return se3.quat_rotate(qc, p - pose7[..., 4:])


def project(pose7, p, K):
    pc = camera_point(pose7, p)
    z = pc[..., 2]
    return np.stack([K[0] * pc[..., 0] / z + K[2], K[1] * pc[..., 1] / z + K[3]], axis=-1), z


def backproject(pose7, uv, depth, K):
    x = (uv[..., 0] - K[2]) / K[0] * depth
    y = (uv[..., 1] - K[3]) / K[1] * depth
    return se3.act(pose7, np.stack([x, y, depth], axis=-1))


def _exact_lengths(rng, n, total, lo, hi, mean_tail):
    """n integer track lengths in [lo,hi] that sum to `total` exactly."""
    L = lo + np.floor(rng.exponential(max(mean_tail, 1e-9), size=n)).astype(np.int64)
    L = np.clip(L, lo, hi)
    diff = int(total - L.sum())
    guard = 0
    while diff != 0 and guard < 200:
        step = 1 if diff > 0 else -1
        ok = np.flatnonzero((L + step >= lo) & (L + step <= hi))
        if ok.size == 0:
            raise ValueError("cannot reach the requested observation count")
        take = ok[rng.permutation(ok.size)[:abs(diff)]]
        L[take] += step
        diff = int(total - L.sum())
        guard += 1
    if diff != 0:
        raise ValueError("cannot reach the requested observation count")
    return L


def _camera_major(rng, cam, lm):
    """Camera-major order with a seeded shuffle inside each camera; landmarks
    relabelled by first appearance (src/OptimizationUtils.cpp:271-276)."""
    order = np.lexsort((rng.random(cam.shape[0]), cam))
    cam, lm = cam[order], lm[order]
    uniq, first = np.unique(lm, return_index=True)
    rank = np.empty(uniq.shape[0], dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(uniq.shape[0])
    new_of_old = np.full(int(lm.max()) + 1 if lm.size else 0, -1, dtype=np.int64)
    new_of_old[uniq] = rank
    return order, cam, new_of_old[lm], uniq[np.argsort(first, kind="stable")]


# --------------------------------------------------------------------------
# TUM-RGBD-shaped keyframe sequences (configs 1-3)
# --------------------------------------------------------------------------
@dataclass
class Sequence:
    K: np.ndarray
    truth_pose: np.ndarray   # [n_kf,7]
    pose: np.ndarray         # [n_kf,7] current estimate (mutated by window solves)
    truth_pt: np.ndarray     # [n_lm,3]
    pt: np.ndarray           # [n_lm,3] current estimate
    kf: np.ndarray           # [n_obs] keyframe of every observation, non-decreasing
    lm: np.ndarray           # [n_obs] landmark id
    uv: np.ndarray           # [n_obs,2] float-rounded pixels
    depth: np.ndarray        # [n_obs] quantised depth
    kf_ptr: np.ndarray       # [n_kf+1] CSR over keyframes


def tum_trajectory(n_kf):
    k = np.arange(n_kf, dtype=np.float64)
    t = 0.3 * np.stack([np.sin(0.10 * k), np.sin(0.07 * k + 1.0), np.sin(0.05 * k + 2.0)], axis=-1)
    rv = 0.05 * np.stack([np.sin(0.03 * k), np.sin(0.04 * k + 0.5), np.sin(0.02 * k + 1.0)], axis=-1)
    return se3.from_rotvec_t(rv, t)


def make_tum_sequence(n_kf, n_lm, n_obs, seed, pix_sigma=0.5, pose_sigma=(0.01, np.deg2rad(0.3)), K=K_TUM, noise=True):
    rng = np.random.default_rng(seed)
    truth = tum_trajectory(n_kf)
    hi = min(n_kf, 60)
    L = _exact_lengths(rng, n_lm, n_obs, min(2, n_kf), hi, n_obs / n_lm - min(2, n_kf))
    start = np.floor(rng.random(n_lm) * (n_kf - L + 1)).astype(np.int64)
    mid = start + L // 2
    uv0 = np.stack([rng.uniform(60, 580, n_lm), rng.uniform(60, 420, n_lm)], axis=-1)
    d0 = rng.uniform(1.0, 4.0, n_lm)
    truth_pt = backproject(truth[mid], uv0, d0, K)
    lm = np.repeat(np.arange(n_lm, dtype=np.int64), L)
    off = np.arange(L.sum(), dtype=np.int64) - np.repeat(np.cumsum(L) - L, L)
    kf = start[lm] + off
    order, kf, lm_new, old_of_new = _camera_major(rng, kf, lm)
    truth_pt = truth_pt[old_of_new]
    uv, z = project(truth[kf], truth_pt[lm_new], K)
    if np.min(z) <= 0.2:
        raise ValueError("synthetic landmark behind a camera")
    sig = pix_sigma if noise else 0.0
    uv = (uv + sig * rng.standard_normal(uv.shape)).astype(np.float32).astype(np.float64)
    dn = z + (0.003 * z * z * rng.standard_normal(z.shape) if noise else 0.0)
    depth = (np.round(dn * 5000.0) / 5000.0).astype(np.float32).astype(np.float64)
    # drifting initial trajectory: truth o exp(accumulated noise)
    if noise:
        step = np.concatenate([pose_sigma[0] * rng.standard_normal((n_kf, 3)), pose_sigma[1] * rng.standard_normal((n_kf, 3))], axis=-1)
        step[0] = 0.0
        drift = np.zeros((n_kf, 7)); drift[:, 3] = 1.0
        acc = np.array([0, 0, 0, 1.0, 0, 0, 0])
        for k in range(n_kf):
            acc = se3.mul(acc, se3.exp(step[k]))
            drift[k] = acc
        pose = se3.mul(truth, drift)
    else:
        pose = truth.copy()
    # landmark = back-projection from the first observing keyframe (src/Map3D.cpp:44, 89-91)
    first = np.unique(lm_new, return_index=True)[1]
    pt = backproject(pose[kf[first]], uv[first], depth[first], K)
    kf_ptr = np.searchsorted(kf, np.arange(n_kf + 1)).astype(np.int64)
    return Sequence(np.array(K, dtype=np.float64), truth, pose, truth_pt, pt, kf.astype(np.int32), lm_new.astype(np.int32),
                    uv, depth, kf_ptr)


@dataclass
class Window:
    problem: BAProblem
    kf_i: int
    kf_f: int
    lm_ids: np.ndarray      # landmark id of every problem point
    T0: np.ndarray          # pose of keyframe kf_i before the frame change


def window_problem(seq: Sequence, kf_i: int, kf_f: int, intr=None, intr_prior=None, use_depth=True) -> Window:
    """The problem windowOptimize hands to the solver for keyframes kf_i..kf_f:
    poses and landmarks moved into the frame of keyframe kf_i (:231-232, :248,
    :274), observations with depth > 1e-15 (:265), first pose fixed (:299)."""
    a, b = int(seq.kf_ptr[kf_i]), int(seq.kf_ptr[kf_f + 1])
    keep = seq.depth[a:b] > 1e-15
    kf = seq.kf[a:b][keep]
    lm = seq.lm[a:b][keep]
    uniq, first = np.unique(lm, return_index=True)
    order = np.argsort(first, kind="stable")
    lm_ids = uniq[order]
    new_of_old = np.full(int(seq.pt.shape[0]), -1, dtype=np.int64)
    new_of_old[lm_ids] = np.arange(lm_ids.shape[0])
    T0 = seq.pose[kf_i].copy()
    T0inv = se3.inverse(T0)
    pose = se3.mul(np.broadcast_to(T0inv, (kf_f - kf_i + 1, 7)), seq.pose[kf_i:kf_f + 1])
    pt = se3.act(T0inv, seq.pt[lm_ids])
    K = seq.K if intr is None else np.asarray(intr, dtype=np.float64)
    prob = BAProblem(pose, pt, (kf - kf_i).astype(np.int32), new_of_old[lm].astype(np.int32), seq.uv[a:b][keep],
                     seq.depth[a:b][keep] if use_depth else None, K, seq.K if intr_prior is None else intr_prior, 0,
                     {"kf_i": kf_i, "kf_f": kf_f})
    return Window(prob, kf_i, kf_f, lm_ids, T0)


def write_back(seq: Sequence, win: Window, pose7, pt3):
    """:303-310 -- poses and the touched landmarks back into the world frame."""
    seq.pose[win.kf_i:win.kf_f + 1] = se3.mul(np.broadcast_to(win.T0, pose7.shape), pose7)
    seq.pt[win.lm_ids] = se3.act(win.T0, pt3)


# --------------------------------------------------------------------------
# BAL-shaped street loop (configs 4-5): many cameras, points seen by runs of
# neighbouring cameras, reprojection only (NS mode)
# --------------------------------------------------------------------------
def make_loop_problem(n_cam, n_pt, n_obs, seed, spacing=0.3, K=K_TUM, pix_sigma=0.5, pose_sigma=(0.02, np.deg2rad(0.3)),
                      pt_sigma=0.05, max_track=40):
    rng = np.random.default_rng(seed)
    # street path: an arc with <= 1 degree of heading change per camera (a closed
    # loop once n_cam >= 360); tracks are runs of consecutive cameras
    dth = min(2.0 * np.pi / n_cam, np.deg2rad(1.0))
    th = dth * np.arange(n_cam)
    Rc = spacing / dth
    c, s = np.cos(th), np.sin(th)
    R = np.zeros((n_cam, 3, 3))
    R[:, 0, 0] = -s; R[:, 1, 0] = c            # x_c = tangent
    R[:, 2, 1] = 1.0                            # y_c = world z
    R[:, 0, 2] = c; R[:, 1, 2] = s             # z_c = outward radial (optical axis)
    t = np.stack([Rc * c, Rc * s, 0.2 * np.sin(7.0 * th)], axis=-1)
    truth = np.concatenate([rotmat_to_quat(R), t], axis=-1)
    L = _exact_lengths(rng, n_pt, n_obs, 2, min(max_track, n_cam), n_obs / n_pt - 2.0)
    home = np.floor(rng.random(n_pt) * (n_cam - L + 1)).astype(np.int64)
    mid = home + L // 2
    uv0 = np.stack([rng.uniform(0.25 * 640, 0.75 * 640, n_pt), rng.uniform(0.2 * 480, 0.8 * 480, n_pt)], axis=-1)
    dmin = np.maximum(5.0, 1.5 * L * spacing)
    d0 = dmin + rng.random(n_pt) * (60.0 - dmin)
    truth_pt = backproject(truth[mid], uv0, d0, K)
    lm = np.repeat(np.arange(n_pt, dtype=np.int64), L)
    off = np.arange(L.sum(), dtype=np.int64) - np.repeat(np.cumsum(L) - L, L)
    cam = home[lm] + off
    order, cam, lm_new, old_of_new = _camera_major(rng, cam, lm)
    truth_pt = truth_pt[old_of_new]
    uv, z = project(truth[cam], truth_pt[lm_new], K)
    if np.min(z) <= 0.5:
        raise ValueError("synthetic point too close to a camera")
    uv = uv + pix_sigma * rng.standard_normal(uv.shape)
    noise6 = np.concatenate([pose_sigma[0] * rng.standard_normal((n_cam, 3)), pose_sigma[1] * rng.standard_normal((n_cam, 3))], axis=-1)
    noise6[0] = 0.0
    pose = se3.mul(truth, se3.exp(noise6))
    pt = truth_pt + pt_sigma * rng.standard_normal(truth_pt.shape)
    return BAProblem(pose, pt, cam.astype(np.int32), lm_new.astype(np.int32), uv, None, np.array(K, dtype=np.float64), None, 0,
                     {"truth_pose": truth, "truth_pt": truth_pt})


# --------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------
CONFIGS = {
    1: dict(kind="tum_window", n_kf=7, n_lm=500, n_obs=3000, window=7),
    2: dict(kind="tum_sliding", n_kf=800, n_lm=80000, n_obs=480000, window=20, frame_frequency=10),
    3: dict(kind="tum_global", n_kf=800, n_lm=60000, n_obs=400000),
    4: dict(kind="loop", n_cam=1723, n_pt=156000, n_obs=680000),
    5: dict(kind="loop", n_cam=10000, n_pt=2000000, n_obs=8000000),
}


def make_config(cfg, scale=1.0, seed=None):
    """Returns a BAProblem (cfg 1, 3, 4, 5) or a Sequence (cfg 2).  `scale` < 1
    shrinks the counts proportionally (parity tests at oracle-friendly sizes)."""
    c = CONFIGS[cfg]
    seed = SEED_BASE + cfg if seed is None else seed
    sc = lambda v, lo=2: max(lo, int(round(v * scale)))
    if c["kind"] == "tum_window":
        seq = make_tum_sequence(c["n_kf"], sc(c["n_lm"], 8), sc(c["n_obs"], 24), seed)
        return window_problem(seq, 0, c["n_kf"] - 1).problem
    if c["kind"] == "tum_sliding":
        n_kf = sc(c["n_kf"], c["window"])
        return make_tum_sequence(n_kf, sc(c["n_lm"], 8), sc(c["n_obs"], 24), seed)
    if c["kind"] == "tum_global":
        n_kf = sc(c["n_kf"], 4)
        seq = make_tum_sequence(n_kf, sc(c["n_lm"], 8), sc(c["n_obs"], 24), seed)
        return window_problem(seq, 0, n_kf - 1, use_depth=False).problem
    if c["kind"] == "loop":
        return make_loop_problem(sc(c["n_cam"], 4), sc(c["n_pt"], 8), sc(c["n_obs"], 24), seed)
    raise KeyError(cfg)


def shard_points(p: BAProblem, rank: int, n_ranks: int):
    """Point-range sharding for the multi-GPU path: contiguous
    point ranges balanced by observation count; every rank keeps all cameras."""
    if n_ranks == 1:
        return p, np.arange(p.n_pt)
    cnt = np.bincount(p.pt_idx, minlength=p.n_pt)
    cum = np.cumsum(cnt)
    bounds = np.searchsorted(cum, np.arange(n_ranks + 1) * (cum[-1] / n_ranks), side="left")
    bounds[0], bounds[-1] = 0, p.n_pt
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (p.pt_idx >= lo) & (p.pt_idx < hi)
    q = BAProblem(p.pose7, p.pt3[lo:hi], p.cam_idx[keep], p.pt_idx[keep] - lo, p.uv2[keep],
                  None if p.depth is None else p.depth[keep], p.intr, p.intr_prior, p.fixed_cam,
                  {"pt_lo": lo, "pt_hi": hi, "n_obs_total": p.n_obs})
    return q, np.arange(lo, hi)
