"""Trajectory accuracy evaluation: ATE and RPE of an estimated
trajectory against ground truth, as the reference's TUM scripts compute them.

Follows rgb-d-toolset/ (the reference's copies of the TUM RGB-D tools, extended with
Euler-angle errors):
  read_file_list, associate         associate.py:49-101     greedy closest-stamp matching
  align                             evaluate_ate.py:51-83   Horn closed-form rigid alignment
  ate                               evaluate_ate.py:245-293 translational error statistics
  rotation_errors                   evaluate_ate.py:166-224 AYE/APE/ARE + RYE/RPE/RRE (xyz Euler, degrees)
  read_trajectory, relative_pose_error  evaluate_rpe.py:47-297  RPE over (fixed-delta) pose pairs
No plotting (matplotlib is not a dependency).  Written for arrays, not matrices; the
results are checked against the reference scripts' own output (tests/golden/eval_golden.json).
"""
import numpy as np
from scipy.spatial.transform import Rotation


# ---------------------------------------------------------------- files / association
def read_file_list(path):
    """{stamp: [fields...]} of a `stamp d1 d2 ...` text file; `#` lines and single-token lines skipped."""
    out = {}
    with open(path) as f:
        for line in f.read().replace(",", " ").replace("\t", " ").split("\n"):
            if not line or line[0] == "#":
                continue
            tok = line.split()
            if len(tok) > 1:
                out[float(tok[0])] = tok[1:]
    return out


def associate(first, second, offset=0.0, max_difference=0.02):
    """Greedy one-to-one matching of stamps by increasing |a - (b + offset)| < max_difference;
    ties resolved as the reference's sort of (diff, a, b) tuples does.  Returns sorted (a, b) pairs."""
    a = np.array(list(first.keys()), dtype=np.float64)
    b = np.array(list(second.keys()), dtype=np.float64)
    if a.size == 0 or b.size == 0:
        return []
    d = np.abs(a[:, None] - (b[None, :] + offset))
    ia, ib = np.nonzero(d < max_difference)
    order = np.lexsort((b[ib], a[ia], d[ia, ib]))
    used_a, used_b, matches = set(), set(), []
    for k in order:
        i, j = int(ia[k]), int(ib[k])
        if i in used_a or j in used_b:
            continue
        used_a.add(i)
        used_b.add(j)
        matches.append((float(a[i]), float(b[j])))
    matches.sort()
    return matches


# ---------------------------------------------------------------- ATE
def align(model, data):
    """Horn: rotation R (3x3) and translation t (3,) minimising sum |R model_i + t - data_i|^2,
    and the per-point residual norms.  model, data: (3, n)."""
    model = np.asarray(model, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64)
    mm, dm = model.mean(axis=1, keepdims=True), data.mean(axis=1, keepdims=True)
    W = (model - mm) @ (data - dm).T  # sum of outer(model_i, data_i)
    U, _, Vh = np.linalg.svd(W.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vh
    t = (dm - R @ mm)[:, 0]
    err = R @ model + t[:, None] - data
    return R, t, np.sqrt(np.sum(err * err, axis=0))


def error_stats(e):
    e = np.asarray(e, dtype=np.float64)
    return {"pairs": int(e.size), "rmse": float(np.sqrt(np.dot(e, e) / e.size)), "mean": float(np.mean(e)),
            "median": float(np.median(e)), "std": float(np.std(e)), "min": float(np.min(e)), "max": float(np.max(e))}


def _rms(x):
    x = np.asarray(x, dtype=np.float64)
    return float(np.sqrt(np.sum(np.square(x)) / x.size))


def rotation_errors(gt_quat, est_quat, delta=5, align_rotation=None):
    """Absolute yaw/pitch/roll errors (RMS of Euler-angle differences, degrees; yaw = the first
    'xyz' Euler angle as in the reference) and the relative ones over `delta` keyframes.
    Quaternions (n,4) as (qx,qy,qz,qw); align_rotation (3x3) pre-multiplies the estimate (--horn 1)."""
    g = Rotation.from_quat(np.asarray(gt_quat, dtype=np.float64)).as_euler("xyz", degrees=True)
    r = Rotation.from_quat(np.asarray(est_quat, dtype=np.float64))
    if align_rotation is not None:
        r = Rotation.from_matrix(align_rotation) * r
    e = r.as_euler("xyz", degrees=True)
    out = {}
    for k, (a, rel) in enumerate((("AYE", "RYE"), ("APE", "RPE"), ("ARE", "RRE"))):
        out[a] = _rms(g[:, k] - e[:, k])
        out[rel] = _rms((g[delta:, k] - g[:-delta, k]) - (e[delta:, k] - e[:-delta, k]))
    return out


def ate(gt_path, est_path, offset=0.0, scale=1.0, max_difference=0.02, delta=5, horn=False):
    """Absolute trajectory error of two TUM-format files: association, Horn alignment of the
    estimate onto the ground truth, translational statistics and the Euler-angle errors."""
    first, second = read_file_list(gt_path), read_file_list(est_path)
    matches = associate(first, second, float(offset), float(max_difference))
    if len(matches) < 2:
        raise ValueError("no matching timestamp pairs between ground truth and estimate")
    gxyz = np.array([[float(v) for v in first[a][0:3]] for a, _ in matches]).T
    exyz = np.array([[float(v) * float(scale) for v in second[b][0:3]] for _, b in matches]).T
    R, t, err = align(exyz, gxyz)
    gq = np.array([[float(v) for v in first[a][3:7]] for a, _ in matches])
    eq = np.array([[float(v) for v in second[b][3:7]] for _, b in matches])
    res = {"matches": matches, "rot": R, "trans": t, "trans_error": err, "stats": error_stats(err)}
    res.update(rotation_errors(gq, eq, int(delta), R if horn else None))
    return res


# ---------------------------------------------------------------- RPE
def transform44(t, q):
    """4x4 matrix of a position and a (not necessarily unit) quaternion (qx,qy,qz,qw); identity
    rotation for a vanishing quaternion."""
    q = np.array(q, dtype=np.float64)
    T = np.eye(4)
    T[:3, 3] = t
    nq = float(np.dot(q, q))
    if nq < np.finfo(float).eps * 4.0:
        return T
    q = q * np.sqrt(2.0 / nq)
    o = np.outer(q, q)
    T[:3, :3] = [[1.0 - o[1, 1] - o[2, 2], o[0, 1] - o[2, 3], o[0, 2] + o[1, 3]],
                 [o[0, 1] + o[2, 3], 1.0 - o[0, 0] - o[2, 2], o[1, 2] - o[0, 3]],
                 [o[0, 2] - o[1, 3], o[1, 2] + o[0, 3], 1.0 - o[0, 0] - o[1, 1]]]
    return T


def read_trajectory(path):
    """{stamp: 4x4} of a TUM trajectory file; all-zero quaternions and NaN rows skipped."""
    traj = {}
    with open(path) as f:
        for line in f.read().replace(",", " ").replace("\t", " ").split("\n"):
            if not line or line[0] == "#":
                continue
            v = [float(x) for x in line.split()]
            if v[4:8] == [0, 0, 0, 0] or any(np.isnan(x) for x in v):
                continue
            traj[v[0]] = transform44(v[1:4], v[4:8])
    return traj


def find_closest_index(L, t):
    """The reference's bisection (evaluate_rpe.py:124-152): returns the best index SEEN on the
    bisection path, which is what fixes the pair selection."""
    lo, hi, best = 0, len(L), 0
    diff = abs(L[0] - t)
    while lo < hi:
        mid = int((hi + lo) / 2)
        if abs(L[mid] - t) < diff:
            diff, best = abs(L[mid] - t), mid
        if t == L[mid]:
            return mid
        if L[mid] > t:
            hi = mid
        else:
            lo = mid + 1
    return best


def _angle(T):
    return float(np.arccos(min(1.0, max(-1.0, (np.trace(T[:3, :3]) - 1.0) / 2.0))))


def _cumulative(traj, keys, fn):
    out, s = [0], 0
    for a, b in zip(keys[1:], keys[:-1]):
        s += fn(np.linalg.inv(traj[a]) @ traj[b])
        out.append(s)
    return out


def relative_pose_error(traj_gt, traj_est, delta=1.0, delta_unit="s", offset=0.0, scale=1.0, fixed_delta=True):
    """Rows [stamp_est0, stamp_est1, stamp_gt0, stamp_gt1, trans_error, rot_error] over the pose pairs
    `delta` apart (fixed_delta) or over all pairs (small trajectories; the reference's random
    sub-sampling of > max_pairs pairs is not reproduced)."""
    sg, se = sorted(traj_gt.keys()), sorted(traj_est.keys())
    n = len(se)
    seen = []
    for t in se:
        t_gt = sg[find_closest_index(sg, t + offset)]
        t_back = se[find_closest_index(se, t_gt - offset)]
        if t_back not in seen:
            seen.append(t_back)
    if len(seen) < 2:
        raise ValueError("overlap of the time stamps is too small")
    if delta_unit == "s":
        index = list(se)
    elif delta_unit == "m":
        index = _cumulative(traj_est, se, lambda T: float(np.linalg.norm(T[:3, 3])))
    elif delta_unit == "rad":
        index = _cumulative(traj_est, se, _angle)
    elif delta_unit == "deg":
        index = _cumulative(traj_est, se, lambda T: _angle(T) * 180.0 / np.pi)
    elif delta_unit == "f":
        index = list(range(n))
    else:
        raise ValueError("unknown unit for delta: %r" % delta_unit)
    if fixed_delta:
        pairs = [(i, j) for i in range(n) for j in [find_closest_index(index, index[i] + delta)] if j != n - 1]
    else:
        pairs = [(i, j) for i in range(n) for j in range(n)]
    max_dt = 2.0 * float(np.median(np.diff(sg)))
    rows = []
    for i, j in pairs:
        e0, e1 = se[i], se[j]
        g0, g1 = sg[find_closest_index(sg, e0 + offset)], sg[find_closest_index(sg, e1 + offset)]
        if abs(g0 - (e0 + offset)) > max_dt or abs(g1 - (e1 + offset)) > max_dt:
            continue
        d_est = np.linalg.inv(traj_est[e1]) @ traj_est[e0]
        d_est[:3, 3] *= scale
        d_gt = np.linalg.inv(traj_gt[g1]) @ traj_gt[g0]
        E = np.linalg.inv(d_est) @ d_gt
        rows.append([e0, e1, g0, g1, float(np.linalg.norm(E[:3, 3])), _angle(E)])
    if len(rows) < 2:
        raise ValueError("no matching timestamp pairs between ground truth and estimate")
    return np.array(rows)


def rpe(gt_path, est_path, **kw):
    rows = relative_pose_error(read_trajectory(gt_path), read_trajectory(est_path), **kw)
    rot = error_stats(rows[:, 5])
    return {"rows": rows, "translational": error_stats(rows[:, 4]),
            "rotational_deg": {k: (v * 180.0 / np.pi if k != "pairs" else v) for k, v in rot.items()}}
