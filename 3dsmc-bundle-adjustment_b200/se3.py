"""Vectorised SE3 helpers with the semantics of the vendored Sophus
(headers/sophus/se3.hpp, so3.hpp) -- used on the HOST side only, for the frame
change windowOptimize performs before/after the solve
(src/OptimizationUtils.cpp:231-232, 248, 274, 303-310) and by the synthetic
generators.  Storage order (qx,qy,qz,qw,tx,ty,tz) (se3.hpp:356-365).
"""
import numpy as np

EPS = 1e-10  # headers/sophus/common.hpp:144


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_rotate(q, v):
    """Eigen _transformVector: v + w*uv + q.vec x uv, uv = 2 (q.vec x v)  (so3.hpp:322-324)."""
    qv = q[..., :3]
    uv = np.cross(qv, v)
    uv = uv + uv
    return v + q[..., 3:4] * uv + np.cross(qv, uv)


def quat_to_R(q):
    """Eigen toRotationMatrix (non-normalising), as the cost functors use it (:41)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x2, y2, z2 = 2 * x, 2 * y, 2 * z
    wx, wy, wz = x2 * w, y2 * w, z2 * w
    xx, xy, xz = x2 * x, y2 * x, z2 * x
    yy, yz, zz = y2 * y, z2 * y, z2 * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (yy + zz); R[..., 0, 1] = xy - wz; R[..., 0, 2] = xz + wy
    R[..., 1, 0] = xy + wz; R[..., 1, 1] = 1 - (xx + zz); R[..., 1, 2] = yz - wx
    R[..., 2, 0] = xz - wy; R[..., 2, 1] = yz + wx; R[..., 2, 2] = 1 - (xx + yy)
    return R


def mul(a, b):
    """SE3 product: t += R t2, then q *= q2 with the 2/(1+|q|^2) renormalisation
    (se3.hpp:317-321, so3.hpp:339-356)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    t = a[..., 4:] + quat_rotate(a[..., :4], b[..., 4:])
    q = quat_mul(a[..., :4], b[..., :4])
    n2 = np.sum(q * q, axis=-1, keepdims=True)
    q = np.where(n2 != 1.0, q * (2.0 / (1.0 + n2)), q)
    return np.concatenate([q, t], axis=-1)


def inverse(a):
    """SE3::inverse (se3.hpp:186-189): normalised conjugate, t' = R^-1 (-t)."""
    a = np.asarray(a, dtype=np.float64)
    q = a[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    q = q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))
    t = quat_rotate(q, a[..., 4:] * -1.0)
    return np.concatenate([q, t], axis=-1)


def act(a, p):
    """SE3 * point (se3.hpp:299-301)."""
    a = np.asarray(a, dtype=np.float64)
    return quat_rotate(a[..., :4], np.asarray(p, dtype=np.float64)) + a[..., 4:]


def exp(d):
    """SE3::exp (se3.hpp:725-746) with SO3::expAndTheta (so3.hpp:537-571)."""
    d = np.asarray(d, dtype=np.float64)
    ups, om = d[..., :3], d[..., 3:]
    th2 = np.sum(om * om, axis=-1)
    th = np.sqrt(th2)
    small = th < EPS
    ths = np.where(small, 1.0, th)
    th4 = th2 * th2
    imag = np.where(small, 0.5 - th2 / 48.0 + th4 / 3840.0, np.sin(0.5 * ths) / ths)
    real = np.where(small, 1.0 - th2 / 8.0 + th4 / 384.0, np.cos(0.5 * ths))
    q = np.concatenate([imag[..., None] * om, real[..., None]], axis=-1)
    Om = np.zeros(d.shape[:-1] + (3, 3))
    Om[..., 0, 1] = -om[..., 2]; Om[..., 0, 2] = om[..., 1]
    Om[..., 1, 0] = om[..., 2]; Om[..., 1, 2] = -om[..., 0]
    Om[..., 2, 0] = -om[..., 1]; Om[..., 2, 1] = om[..., 0]
    c1 = np.where(small, 0.0, (1.0 - np.cos(ths)) / np.where(small, 1.0, th2))
    c2 = np.where(small, 0.0, (ths - np.sin(ths)) / np.where(small, 1.0, th2 * ths))
    V = np.eye(3) + c1[..., None, None] * Om + c2[..., None, None] * (Om @ Om)
    V = np.where(small[..., None, None], quat_to_R(q), V)
    t = np.einsum("...ij,...j->...i", V, ups)
    return np.concatenate([q, t], axis=-1)


def from_rotvec_t(rv, t):
    """Pose with rotation exp(rv) and translation t (not the SE3 exponential)."""
    rv = np.asarray(rv, dtype=np.float64)
    z = np.zeros_like(rv)
    q = exp(np.concatenate([z, rv], axis=-1))[..., :4]
    return np.concatenate([q, np.asarray(t, dtype=np.float64)], axis=-1)


def rot_angle(qa, qb):
    """Angle (rad) of qa^-1 qb."""
    d = np.abs(np.sum(qa * qb, axis=-1)) / (np.linalg.norm(qa, axis=-1) * np.linalg.norm(qb, axis=-1))
    return 2.0 * np.arccos(np.clip(d, -1.0, 1.0))
