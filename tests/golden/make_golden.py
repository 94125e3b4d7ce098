#!/usr/bin/env python
"""Generates tests/golden/residual_jacobian_golden.json with mpmath (60 digits).

The reference has NO golden vectors for its optimiser, and
Ceres/Eigen cannot be built in this image, so these vectors are what pins the
oracle and the CUDA kernels:

  * residuals of ReprojectionConstraint / DepthPrior / IntrinsicsPrior
    (src/OptimizationUtils.cpp:25-49, 72-94, 116-125) evaluated from their
    mathematical definition at 60 digits;
  * the LOCAL pose Jacobian d r(T*exp(delta)) / d delta at delta=0 -- i.e. what
    Ceres obtains from  J_ambient * LocalParameterizationSE3::ComputeJacobian
    (headers/sophus/local_parameterization_se3.hpp:17-37) -- computed by
    60-digit central differences of the composite function (independent of any
    hand-derived formula), plus point and intrinsics Jacobians likewise;
  * SE3 exp / product / inverse / point action (headers/sophus/se3.hpp:725-746,
    317-321, 186-189, 299-301) incl. the small-angle branch;
  * ceres::HuberLoss values.

Run:  python tests/golden/make_golden.py     (deterministic; seed fixed)
"""
import json
import os
import random

from mpmath import mp, mpf, sqrt, sin, cos, matrix

mp.dps = 60
H = mpf(10) ** (-22)


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz)


def rot_eigen(q):
    """Eigen toRotationMatrix polynomial (non-normalising)."""
    x, y, z, w = q
    return matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def se3_exp(d):
    ups = matrix(d[:3])
    om = d[3:]
    th2 = om[0] ** 2 + om[1] ** 2 + om[2] ** 2
    th = sqrt(th2)
    Om = matrix([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    I = matrix([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    if th == 0:
        q = (mpf(0), mpf(0), mpf(0), mpf(1))
        V = I
    else:
        s = sin(th / 2) / th
        q = (s * om[0], s * om[1], s * om[2], cos(th / 2))
        V = I + (1 - cos(th)) / th2 * Om + (th - sin(th)) / (th2 * th) * Om * Om
    t = V * ups
    return list(q) + [t[0], t[1], t[2]]


def se3_mul(a, b):
    Ra = rot_eigen(a[:4])
    t = matrix(a[4:]) + Ra * matrix(b[4:])
    q = qmul(a[:4], b[:4])
    return list(q) + [t[0], t[1], t[2]]


def se3_inv(a):
    x, y, z, w = a[:4]
    n = sqrt(x * x + y * y + z * z + w * w)
    q = (-x / n, -y / n, -z / n, w / n)
    t = rot_eigen(q) * (-matrix(a[4:]))
    return list(q) + [t[0], t[1], t[2]]


def se3_act(a, p):
    r = rot_eigen(a[:4]) * matrix(p) + matrix(a[4:])
    return [r[0], r[1], r[2]]


def residuals(pose, pt, intr, uv, depth, w_repr, w_unpr):
    R = rot_eigen(pose[:4])
    pc = R.T * (matrix(pt) - matrix(pose[4:]))
    X, Y, Z = pc[0], pc[1], pc[2]
    u = intr[0] * X / Z + intr[2]
    v = intr[1] * Y / Z + intr[3]
    return [sqrt(w_repr) * (u - uv[0]), sqrt(w_repr) * (v - uv[1]), sqrt(w_unpr) * (depth - Z)]


def central(f, n):
    """Jacobian of f: R^n -> R^m at 0 by central differences (60 digits)."""
    cols = []
    for k in range(n):
        e = [mpf(0)] * n
        e[k] = H
        fp = f(e)
        e[k] = -H
        fm = f(e)
        cols.append([(a - b) / (2 * H) for a, b in zip(fp, fm)])
    m = len(cols[0])
    return [[cols[k][i] for k in range(n)] for i in range(m)]


def fl(x):
    return float(x)


def to_mp(v):
    return [mpf(float(x)) for x in v]


def rand_unit_quat(rng, small=False):
    if small:
        ax = [rng.gauss(0, 1) for _ in range(3)]
        n = sum(a * a for a in ax) ** 0.5
        ang = rng.uniform(0, 0.09)
        s = sin(mpf(ang) / 2)
        q = [mpf(a / n) * s for a in ax] + [cos(mpf(ang) / 2)]
    else:
        q = [mpf(rng.gauss(0, 1)) for _ in range(4)]
        n = sqrt(sum(a * a for a in q))
        q = [a / n for a in q]
    return [float(a) for a in q]  # rounded to double: |q| = 1 +- 1e-16


def main():
    rng = random.Random(0xBA60)
    samples = []
    for i in range(48):
        small = i % 2 == 0
        q = rand_unit_quat(rng, small)
        t = [rng.uniform(-0.5, 0.5) for _ in range(3)]
        pose = q + t
        # a point in front of the camera: choose camera-frame coords then map to world
        Z = rng.uniform(0.4, 6.0) if i % 7 else rng.uniform(20.0, 80.0)
        Xc = rng.uniform(-0.55, 0.55) * Z
        Yc = rng.uniform(-0.42, 0.42) * Z
        pw = se3_act(to_mp(pose), [mpf(Xc), mpf(Yc), mpf(Z)])
        pt = [float(a) for a in pw]
        intr = [525.0 + rng.uniform(-5, 5), 525.0 + rng.uniform(-5, 5), 319.5 + rng.uniform(-3, 3), 239.5 + rng.uniform(-3, 3)]
        # observed pixel: float32-rounded like cv::KeyPoint, a few px from the projection
        import struct
        u0 = 525.0 * Xc / Z + 319.5 + rng.gauss(0, 2.0 if i % 3 else 0.01)
        v0 = 525.0 * Yc / Z + 239.5 + rng.gauss(0, 2.0 if i % 3 else 0.01)
        uv = [struct.unpack("f", struct.pack("f", a))[0] for a in (u0, v0)]
        depth = struct.unpack("f", struct.pack("f", Z + rng.gauss(0, 0.01)))[0]
        n_obs = [3000, 12000, 400000, 8000000][i % 4]
        w_repr, w_unpr = 1.0 / n_obs, 10.0 / n_obs

        P, X, K = to_mp(pose), to_mp(pt), to_mp(intr)
        UV, D = to_mp(uv), mpf(depth)
        WR, WU = mpf(w_repr), mpf(w_unpr)
        r = residuals(P, X, K, UV, D, WR, WU)
        Jpose = central(lambda d: residuals(se3_mul(P, se3_exp(d)), X, K, UV, D, WR, WU), 6)
        Jpt = central(lambda d: residuals(P, [X[k] + d[k] for k in range(3)], K, UV, D, WR, WU), 3)
        Jk = central(lambda d: residuals(P, X, [K[k] + d[k] for k in range(4)], UV, D, WR, WU), 4)
        # cross-check against the closed form of SURVEY Appendix B
        R = rot_eigen(P[:4])
        pc = R.T * (matrix(X) - matrix(P[4:]))
        Xm, Ym, Zm = pc[0], pc[1], pc[2]
        iz = 1 / Zm
        sw, sd = sqrt(WR), sqrt(WU)
        cf = [[sw * K[0] * c for c in (-iz, 0, Xm * iz * iz, Xm * Ym * iz * iz, -(1 + Xm * Xm * iz * iz), Ym * iz)],
              [sw * K[1] * c for c in (0, -iz, Ym * iz * iz, 1 + Ym * Ym * iz * iz, -Xm * Ym * iz * iz, -Xm * iz)],
              [sd * c for c in (0, 0, 1, Ym, -Xm, 0)]]
        for a in range(3):
            for b in range(6):
                scale = max(abs(cf[a][b]), mpf(1e-30))
                # closed form assumes |q|=1; the double-rounded q is unit to ~1e-16
                assert abs(cf[a][b] - Jpose[a][b]) <= mpf(1e-13) * max(scale, sw), (i, a, b, cf[a][b], Jpose[a][b])
        samples.append(dict(pose=pose, pt=pt, intr=intr, uv=uv, depth=depth, w_repr=w_repr, w_unpr=w_unpr,
                            r=[fl(a) for a in r],
                            Jpose=[[fl(a) for a in row] for row in Jpose],
                            Jpt=[[fl(a) for a in row] for row in Jpt],
                            Jintr=[[fl(a) for a in row] for row in Jk]))

    se3 = []
    for i in range(24):
        if i % 4 == 0:
            d = [rng.uniform(-0.1, 0.1) for _ in range(3)] + [rng.uniform(-1e-11, 1e-11) for _ in range(3)]  # small-angle branch
        elif i % 4 == 1:
            d = [rng.uniform(-0.05, 0.05) for _ in range(3)] + [rng.uniform(-0.02, 0.02) for _ in range(3)]
        else:
            d = [rng.uniform(-1, 1) for _ in range(3)] + [rng.uniform(-1.5, 1.5) for _ in range(3)]
        a = rand_unit_quat(rng) + [rng.uniform(-2, 2) for _ in range(3)]
        b = rand_unit_quat(rng, True) + [rng.uniform(-2, 2) for _ in range(3)]
        p = [rng.uniform(-3, 3) for _ in range(3)]
        A, B, Dm, Pm = to_mp(a), to_mp(b), to_mp(d), to_mp(p)
        e = se3_exp(Dm)
        se3.append(dict(delta=d, a=a, b=b, p=p,
                        exp=[fl(x) for x in e],
                        a_mul_exp=[fl(x) for x in se3_mul(A, e)],
                        a_mul_b=[fl(x) for x in se3_mul(A, B)],
                        a_inv=[fl(x) for x in se3_inv(A)],
                        a_act_p=[fl(x) for x in se3_act(A, Pm)]))

    hub = []
    for a, s in [(1e-3, 1e-8), (1e-3, 1e-6), (1e-3, 1.0000001e-6), (1e-3, 4e-4), (1e-3, 2.5), (0.5, 0.2), (0.5, 0.3), (2.0, 100.0)]:
        A, S = mpf(a), mpf(s)
        if S > A * A:
            rt = sqrt(S)
            rho = [2 * A * rt - A * A, A / rt, -(A / rt) / (2 * S)]
        else:
            rho = [S, mpf(1), mpf(0)]
        hub.append(dict(a=a, s=s, rho=[fl(x) for x in rho]))

    out = dict(meta=dict(generator="tests/golden/make_golden.py", mp_dps=mp.dps, fd_step="1e-22",
                         note="values rounded to nearest double from 60-digit arithmetic"),
               residual_jacobian=samples, se3=se3, huber=hub)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "residual_jacobian_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
