"""ctypes binding of the CPU oracle (oracle/libba_oracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class OraProblem(C.Structure):
    _fields_ = [
        ("n_cam", C.c_int32), ("n_pt", C.c_int32), ("n_obs", C.c_int32), ("fixed_cam", C.c_int32),
        ("pose7", c_double_p), ("pt3", c_double_p),
        ("cam_idx", c_int32_p), ("pt_idx", c_int32_p),
        ("uv2", c_double_p), ("depth", c_double_p),
        ("intr", C.c_double * 4), ("intr_prior", C.c_double * 4),
    ]


class OraOptions(C.Structure):
    _fields_ = [
        ("huber_repr", C.c_double), ("huber_unpr", C.c_double),
        ("weight_unpr", C.c_double), ("weight_intrinsics", C.c_double),
        ("max_num_iterations", C.c_int32), ("eta", C.c_double),
        ("use_depth_prior", C.c_int32), ("optimize_intrinsics", C.c_int32),
        ("solver", C.c_int32), ("n_obs_total", C.c_int64),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("max_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
        ("max_pcg_iterations", C.c_int32), ("min_pcg_iterations", C.c_int32),
        ("residual_reset_period", C.c_int32), ("num_threads", C.c_int32),
    ]


class OraIter(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32),
        ("step_is_successful", C.c_int32), ("linear_iters", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
        ("step_norm", C.c_double), ("relative_decrease", C.c_double), ("radius", C.c_double),
        ("model_cost_change", C.c_double),
    ]


class OraSummary(C.Structure):
    _fields_ = [
        ("termination", C.c_int32), ("num_iterations", C.c_int32),
        ("num_successful", C.c_int32), ("num_unsuccessful", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("total_linear_iters", C.c_int64),
        ("seconds_total", C.c_double), ("seconds_linearize", C.c_double),
        ("seconds_linear_solve", C.c_double),
    ]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "libba_oracle.so")
        src = os.path.join(ORACLE_DIR, "ba_oracle.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            build_oracle()
        L = C.CDLL(path)
        L.ora_default_options.argtypes = [C.POINTER(OraOptions)]
        L.ora_evaluate.restype = C.c_int
        L.ora_solve.restype = C.c_int
        L.ora_schur_matvec.restype = C.c_int
        _LIB = L
    return _LIB


def default_options(**kw):
    o = OraOptions()
    lib().ora_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


class Problem:
    """Owns contiguous numpy buffers and the ctypes view of them."""

    def __init__(self, pose7, pt3, cam_idx, pt_idx, uv2, depth, intr, intr_prior=None, fixed_cam=0):
        self.pose7 = np.ascontiguousarray(pose7, dtype=np.float64).reshape(-1, 7).copy()
        self.pt3 = np.ascontiguousarray(pt3, dtype=np.float64).reshape(-1, 3).copy()
        self.cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int32).copy()
        self.pt_idx = np.ascontiguousarray(pt_idx, dtype=np.int32).copy()
        self.uv2 = np.ascontiguousarray(uv2, dtype=np.float64).reshape(-1, 2).copy()
        self.depth = None if depth is None else np.ascontiguousarray(depth, dtype=np.float64).copy()
        self.intr = np.ascontiguousarray(intr, dtype=np.float64).copy()
        self.intr_prior = self.intr.copy() if intr_prior is None else np.ascontiguousarray(intr_prior, dtype=np.float64).copy()
        self.fixed_cam = int(fixed_cam)

    @property
    def n_cam(self):
        return self.pose7.shape[0]

    @property
    def n_pt(self):
        return self.pt3.shape[0]

    @property
    def n_obs(self):
        return self.cam_idx.shape[0]

    def copy(self):
        return Problem(self.pose7, self.pt3, self.cam_idx, self.pt_idx, self.uv2, self.depth,
                       self.intr, self.intr_prior, self.fixed_cam)

    def c_struct(self):
        p = OraProblem()
        p.n_cam, p.n_pt, p.n_obs, p.fixed_cam = self.n_cam, self.n_pt, self.n_obs, self.fixed_cam
        p.pose7, p.pt3 = _dp(self.pose7), _dp(self.pt3)
        p.cam_idx, p.pt_idx = _ip(self.cam_idx), _ip(self.pt_idx)
        p.uv2, p.depth = _dp(self.uv2), _dp(self.depth)
        for i in range(4):
            p.intr[i] = self.intr[i]
            p.intr_prior[i] = self.intr_prior[i]
        return p

    def sync_back(self, cs):
        for i in range(4):
            self.intr[i] = cs.intr[i]


def evaluate(prob, opts, want_jac=True):
    R = 2 + (1 if opts.use_depth_prior else 0)
    n = prob.n_obs
    cost = C.c_double(0.0)
    out = {}
    r = np.zeros((n, R)); out["r"] = r
    Jc = np.zeros((n, R, 6)) if want_jac else None
    Jp = np.zeros((n, R, 3)) if want_jac else None
    Jk = np.zeros((n, 2, 4)) if want_jac else None
    gc = np.zeros((prob.n_cam, 6)); gp = np.zeros((prob.n_pt, 3)); gk = np.zeros(4)
    cs = prob.c_struct()
    rc = lib().ora_evaluate(C.byref(cs), C.byref(opts), C.byref(cost), _dp(r), _dp(Jc), _dp(Jp), _dp(Jk),
                            _dp(gc), _dp(gp), _dp(gk))
    out.update(rc=rc, cost=cost.value, Jc=Jc, Jp=Jp, Jk=Jk, g_c=gc, g_p=gp, g_k=gk)
    return out


def build_indices(prob):
    perm = np.zeros(prob.n_obs, dtype=np.int32)
    pt_rowptr = np.zeros(prob.n_pt + 1, dtype=np.int32)
    cam_rowptr = np.zeros(prob.n_cam + 1, dtype=np.int32)
    cs = prob.c_struct()
    lib().ora_build_indices(C.byref(cs), _ip(perm), _ip(pt_rowptr), _ip(cam_rowptr))
    return perm, pt_rowptr, cam_rowptr


def solve(prob, opts, trace_cap=256):
    """Runs the oracle LM in place on `prob`. Returns (rc, summary, trace list)."""
    cs = prob.c_struct()
    s = OraSummary()
    tr = (OraIter * trace_cap)()
    rc = lib().ora_solve(C.byref(cs), C.byref(opts), C.byref(s), tr, C.c_int32(trace_cap))
    prob.sync_back(cs)
    n = min(trace_cap, s.num_iterations + 1)
    trace = []
    for i in range(n):
        t = tr[i]
        if i > 0 and t.iteration == 0:
            break
        trace.append({f[0]: getattr(t, f[0]) for f in OraIter._fields_})
    return rc, s, trace


def schur_matvec(prob, opts, radius, x, pt_begin=0, pt_end=0, include_diag=1):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    cs = prob.c_struct()
    rc = lib().ora_schur_matvec(C.byref(cs), C.byref(opts), C.c_double(radius), _dp(x), _dp(y),
                                C.c_int32(pt_begin), C.c_int32(pt_end), C.c_int32(include_diag))
    if rc:
        raise RuntimeError("ora_schur_matvec rc=%d" % rc)
    return y


def se3_exp(d):
    d = np.ascontiguousarray(d, dtype=np.float64); o = np.zeros(7)
    lib().ora_se3_exp(_dp(d), _dp(o)); return o


def se3_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64); o = np.zeros(7)
    lib().ora_se3_mul(_dp(a), _dp(b), _dp(o)); return o


def se3_inverse(a):
    a = np.ascontiguousarray(a, dtype=np.float64); o = np.zeros(7)
    lib().ora_se3_inverse(_dp(a), _dp(o)); return o


def se3_act(a, p):
    a = np.ascontiguousarray(a, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64); o = np.zeros(3)
    lib().ora_se3_act(_dp(a), _dp(p), _dp(o)); return o


def backproject(uv, depth_img, intr, pose7=None):
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
    img = np.ascontiguousarray(depth_img, dtype=np.float32)
    k = np.ascontiguousarray(intr, dtype=np.float64)
    n = uv.shape[0]
    local, world = np.zeros((n, 3)), np.zeros((n, 3))
    fp = C.POINTER(C.c_float)
    p = np.ascontiguousarray(pose7, dtype=np.float64) if pose7 is not None else None
    rc = lib().ora_backproject(n, uv.ctypes.data_as(fp), img.ctypes.data_as(fp), img.shape[1], img.shape[0], _dp(k),
                               _dp(p) if p is not None else None, _dp(local), _dp(world) if p is not None else None)
    return rc, local, (world if p is not None else None)


def se3_dx(a):
    a = np.ascontiguousarray(a, dtype=np.float64); o = np.zeros((7, 6))
    lib().ora_se3_dx_this_mul_exp_x_at_0(_dp(a), _dp(o)); return o


def reprojection(pose7, pt, intr, uv, weight):
    pose7, pt, intr, uv = [np.ascontiguousarray(v, dtype=np.float64) for v in (pose7, pt, intr, uv)]
    r = np.zeros(2); Jpose = np.zeros((2, 7)); Jpt = np.zeros((2, 3)); Jk = np.zeros((2, 4))
    lib().ora_reprojection(_dp(pose7), _dp(pt), _dp(intr), _dp(uv), C.c_double(weight), _dp(r), _dp(Jpose), _dp(Jpt), _dp(Jk))
    return r, Jpose, Jpt, Jk


def depth_prior(pose7, pt, intr, depth, weight):
    pose7, pt, intr = [np.ascontiguousarray(v, dtype=np.float64) for v in (pose7, pt, intr)]
    r = np.zeros(1); Jpose = np.zeros((1, 7)); Jpt = np.zeros((1, 3)); Jk = np.zeros((1, 4))
    lib().ora_depth_prior(_dp(pose7), _dp(pt), _dp(intr), C.c_double(depth), C.c_double(weight), _dp(r), _dp(Jpose), _dp(Jpt), _dp(Jk))
    return r, Jpose, Jpt, Jk


def huber(a, s):
    rho = np.zeros(3)
    lib().ora_huber(C.c_double(a), C.c_double(s), _dp(rho))
    return rho
