/*
 * ba_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See ba_oracle.h for scope and parity status ("parity unpinned" at the Ceres
 * boundary; pinned by mpmath golden vectors in tests/golden/).
 *
 * Every function cites the reference lines (relative to /root/reference) or the
 * third-party algorithm (ceres-solver 2.0.0, eigen 3.4.0 -- conanfile.txt:2-4)
 * it restates.  Deliberately literal and unoptimised: residuals are evaluated
 * with forward-mode dual numbers exactly as ceres::AutoDiffCostFunction does,
 * the pose Jacobian is chained through Sophus' 7x6 plus-Jacobian, the Jacobian
 * is column-scaled in place as Ceres does, and the LM loop follows
 * trust_region_minimizer.cc step by step.
 */
#include "ba_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void ora_default_options(ora_options *o) {
  memset(o, 0, sizeof(*o));
  /* headers/BundleAdjustmentConfig.h:47-50 */
  o->huber_repr = 1e-3;
  o->weight_intrinsics = 1e-6;
  o->weight_unpr = 10.0;
  o->huber_unpr = 1e-3;
  /* headers/BundleAdjustmentConfig.h:64-65 */
  o->max_num_iterations = 75;
  o->eta = 1e-6;
  o->use_depth_prior = 1;
  o->optimize_intrinsics = 1;
  o->solver = ORA_SOLVER_DENSE_SCHUR;
  o->n_obs_total = 0;
  /* ceres 2.0.0 Solver::Options defaults */
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_radius = 1e4;
  o->max_radius = 1e16;
  o->min_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->max_pcg_iterations = 500;
  o->min_pcg_iterations = 0;
  o->residual_reset_period = 10;
  o->num_threads = 1;
}

/* ======================================================================
 * Dual numbers: ceres::Jet<double,14> (7 pose + 3 point + 4 intrinsics),
 * the type AutoDiffCostFunction<.,.,7,3,4> instantiates the functors with
 * (src/OptimizationUtils.cpp:53, 98).
 * ====================================================================== */
#define NJ 14
typedef struct {
  double a;
  double v[NJ];
} jet;

static inline jet j_const(double c) {
  jet r;
  r.a = c;
  for (int i = 0; i < NJ; ++i) r.v[i] = 0.0;
  return r;
}
static inline jet j_var(double c, int k) {
  jet r = j_const(c);
  r.v[k] = 1.0;
  return r;
}
static inline jet j_add(jet x, jet y) {
  jet r;
  r.a = x.a + y.a;
  for (int i = 0; i < NJ; ++i) r.v[i] = x.v[i] + y.v[i];
  return r;
}
static inline jet j_sub(jet x, jet y) {
  jet r;
  r.a = x.a - y.a;
  for (int i = 0; i < NJ; ++i) r.v[i] = x.v[i] - y.v[i];
  return r;
}
/* jet.h: Jet(f.a*g.a, f.a*g.v + f.v*g.a) */
static inline jet j_mul(jet x, jet y) {
  jet r;
  r.a = x.a * y.a;
  for (int i = 0; i < NJ; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a;
  return r;
}
/* jet.h: g_inv = 1/g.a; q = f.a*g_inv; Jet(q, (f.v - q*g.v)*g_inv) */
static inline jet j_div(jet x, jet y) {
  jet r;
  const double inv = 1.0 / y.a;
  const double q = x.a * inv;
  r.a = q;
  for (int i = 0; i < NJ; ++i) r.v[i] = (x.v[i] - q * y.v[i]) * inv;
  return r;
}

/* Eigen 3.4.0 QuaternionBase::toRotationMatrix (non-normalising), the routine
 * behind `q.matrix()` at src/OptimizationUtils.cpp:41, 85. R row-major. */
static void quat_to_R_jet(jet x, jet y, jet z, jet w, jet R[9]) {
  const jet two = j_const(2.0), one = j_const(1.0);
  const jet x2 = j_mul(two, x), y2 = j_mul(two, y), z2 = j_mul(two, z);
  const jet wx = j_mul(x2, w), wy = j_mul(y2, w), wz = j_mul(z2, w);
  const jet xx = j_mul(x2, x), xy = j_mul(y2, x), xz = j_mul(z2, x);
  const jet yy = j_mul(y2, y), yz = j_mul(z2, y), zz = j_mul(z2, z);
  R[0] = j_sub(one, j_add(yy, zz));
  R[1] = j_sub(xy, wz);
  R[2] = j_add(xz, wy);
  R[3] = j_add(xy, wz);
  R[4] = j_sub(one, j_add(xx, zz));
  R[5] = j_sub(yz, wx);
  R[6] = j_sub(xz, wy);
  R[7] = j_add(yz, wx);
  R[8] = j_sub(one, j_add(xx, yy));
}
static void quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = 2.0 * x, y2 = 2.0 * y, z2 = 2.0 * z;
  const double wx = x2 * w, wy = y2 * w, wz = z2 * w;
  const double xx = x2 * x, xy = y2 * x, xz = z2 * x;
  const double yy = y2 * y, yz = z2 * y, zz = z2 * z;
  R[0] = 1.0 - (yy + zz);
  R[1] = xy - wz;
  R[2] = xz + wy;
  R[3] = xy + wz;
  R[4] = 1.0 - (xx + zz);
  R[5] = yz - wx;
  R[6] = xz - wy;
  R[7] = yz + wx;
  R[8] = 1.0 - (xx + yy);
}

/* p_C = q.matrix().transpose() * (p_W - t)   src/OptimizationUtils.cpp:36-41 */
static void point_in_camera_jet(const jet pose[7], const jet xw[3], jet pc[3]) {
  jet R[9];
  quat_to_R_jet(pose[0], pose[1], pose[2], pose[3], R);
  const jet d0 = j_sub(xw[0], pose[4]), d1 = j_sub(xw[1], pose[5]),
            d2 = j_sub(xw[2], pose[6]);
  for (int i = 0; i < 3; ++i)
    pc[i] = j_add(j_add(j_mul(R[0 + i], d0), j_mul(R[3 + i], d1)),
                  j_mul(R[6 + i], d2));
}
static void point_in_camera(const double pose[7], const double xw[3],
                            double pc[3]) {
  double R[9];
  quat_to_R(pose, R);
  const double d0 = xw[0] - pose[4], d1 = xw[1] - pose[5], d2 = xw[2] - pose[6];
  for (int i = 0; i < 3; ++i)
    pc[i] = (R[0 + i] * d0 + R[3 + i] * d1) + R[6 + i] * d2;
}

static void make_jets(const double pose7[7], const double pt[3],
                      const double intr[4], jet pose[7], jet xw[3], jet k[4]) {
  for (int i = 0; i < 7; ++i) pose[i] = j_var(pose7[i], i);
  for (int i = 0; i < 3; ++i) xw[i] = j_var(pt[i], 7 + i);
  for (int i = 0; i < 4; ++i) k[i] = j_var(intr[i], 10 + i);
}

/* ReprojectionConstraint::operator()  src/OptimizationUtils.cpp:25-49.
 * (K * p_C / p_C.z)(0:1): row i of K*p_C, then the division; the zero entries
 * of the Identity-initialised K contribute exact zeros and are elided. */
static void reprojection_jet(const jet pose[7], const jet xw[3],
                             const jet k[4], const double uv[2], double weight,
                             jet res[2]) {
  jet pc[3];
  point_in_camera_jet(pose, xw, pc);
  const jet u = j_div(j_add(j_mul(k[0], pc[0]), j_mul(k[2], pc[2])), pc[2]);
  const jet v = j_div(j_add(j_mul(k[1], pc[1]), j_mul(k[3], pc[2])), pc[2]);
  const jet sw = j_const(sqrt(weight));
  res[0] = j_mul(sw, j_sub(u, j_const(uv[0])));
  res[1] = j_mul(sw, j_sub(v, j_const(uv[1])));
}
static void reprojection_val(const double pose[7], const double xw[3],
                             const double k[4], const double uv[2],
                             double weight, double res[2]) {
  double pc[3];
  point_in_camera(pose, xw, pc);
  const double u = (k[0] * pc[0] + k[2] * pc[2]) / pc[2];
  const double v = (k[1] * pc[1] + k[3] * pc[2]) / pc[2];
  const double sw = sqrt(weight);
  res[0] = sw * (u - uv[0]);
  res[1] = sw * (v - uv[1]);
}
/* DepthPrior::operator()  src/OptimizationUtils.cpp:72-94 */
static void depth_jet(const jet pose[7], const jet xw[3], double depth,
                      double weight, jet res[1]) {
  jet pc[3];
  point_in_camera_jet(pose, xw, pc);
  res[0] = j_mul(j_const(sqrt(weight)), j_sub(j_const(depth), pc[2]));
}
static void depth_val(const double pose[7], const double xw[3], double depth,
                      double weight, double res[1]) {
  double pc[3];
  point_in_camera(pose, xw, pc);
  res[0] = sqrt(weight) * (depth - pc[2]);
}

void ora_reprojection(const double pose7[7], const double pt[3],
                      const double intr[4], const double uv[2], double weight,
                      double r[2], double Jpose[14], double Jpt[6],
                      double Jintr[8]) {
  jet pose[7], xw[3], k[4], res[2];
  make_jets(pose7, pt, intr, pose, xw, k);
  reprojection_jet(pose, xw, k, uv, weight, res);
  for (int i = 0; i < 2; ++i) {
    if (r) r[i] = res[i].a;
    if (Jpose)
      for (int j = 0; j < 7; ++j) Jpose[i * 7 + j] = res[i].v[j];
    if (Jpt)
      for (int j = 0; j < 3; ++j) Jpt[i * 3 + j] = res[i].v[7 + j];
    if (Jintr)
      for (int j = 0; j < 4; ++j) Jintr[i * 4 + j] = res[i].v[10 + j];
  }
}
void ora_depth_prior(const double pose7[7], const double pt[3],
                     const double intr[4], double depth, double weight,
                     double r[1], double Jpose[7], double Jpt[3],
                     double Jintr[4]) {
  jet pose[7], xw[3], k[4], res[1];
  make_jets(pose7, pt, intr, pose, xw, k);
  depth_jet(pose, xw, depth, weight, res);
  if (r) r[0] = res[0].a;
  if (Jpose)
    for (int j = 0; j < 7; ++j) Jpose[j] = res[0].v[j];
  if (Jpt)
    for (int j = 0; j < 3; ++j) Jpt[j] = res[0].v[7 + j];
  if (Jintr)
    for (int j = 0; j < 4; ++j) Jintr[j] = res[0].v[10 + j];
}
/* IntrinsicsPrior::operator()  src/OptimizationUtils.cpp:116-125 */
void ora_intrinsics_prior(const double intr[4], const double prior[4],
                          double weight, double r[4], double J[16]) {
  const double sw = sqrt(weight);
  for (int i = 0; i < 4; ++i) {
    if (r) r[i] = sw * (prior[i] - intr[i]);
    if (J)
      for (int j = 0; j < 4; ++j) J[i * 4 + j] = (i == j) ? -sw : 0.0;
  }
}

/* ceres::HuberLoss::Evaluate (loss_function.cc, ceres 2.0.0) */
void ora_huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = fmax(DBL_MIN, a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s;
    rho[1] = 1.0;
    rho[2] = 0.0;
  }
}

/* ======================================================================
 * Sophus restatement (double only)
 * ====================================================================== */
#define SOPHUS_EPS 1e-10 /* headers/sophus/common.hpp:144 */

/* Eigen quaternion product (Hamilton), a*b; storage (x,y,z,w). */
static void quat_mul(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
/* Eigen QuaternionBase::_transformVector: uv = 2 (q.vec x v);
 * v + w*uv + q.vec x uv.   (so3.hpp:322-324) */
static void quat_rotate(const double q[4], const double v[3], double o[3]) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2],
                  q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2],
                       q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) o[i] = v[i] + q[3] * uv[i] + c[i];
}
/* SO3::operator*=  so3.hpp:339-356 (first-order renormalisation) */
static void so3_mul_inplace(double q[4], const double b[4]) {
  double o[4];
  quat_mul(q, b, o);
  const double n2 = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3];
  if (n2 != 1.0) {
    const double s = 2.0 / (1.0 + n2);
    for (int i = 0; i < 4; ++i) o[i] *= s;
  }
  memcpy(q, o, sizeof(o));
}
/* SE3::operator*  se3.hpp:285-289, 317-321: t += R*t2, then q *= q2 */
void ora_se3_mul(const double a[7], const double b[7], double out[7]) {
  double r[7], rt[3];
  memcpy(r, a, sizeof(r));
  quat_rotate(r, b + 4, rt);
  for (int i = 0; i < 3; ++i) r[4 + i] += rt[i];
  so3_mul_inplace(r, b);
  memcpy(out, r, sizeof(r));
}
/* SE3::inverse  se3.hpp:186-189; SO3::inverse so3.hpp:203-205 goes through the
 * normalising constructor (so3.hpp:434-441 -> normalize() :271-277). */
void ora_se3_inverse(const double a[7], double out[7]) {
  double q[4] = {-a[0], -a[1], -a[2], a[3]};
  const double len = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= len;
  const double nt[3] = {a[4] * -1.0, a[5] * -1.0, a[6] * -1.0};
  double t[3];
  quat_rotate(q, nt, t);
  memcpy(out, q, sizeof(q));
  memcpy(out + 4, t, sizeof(t));
}
/* SE3 * point  se3.hpp:299-301 */
void ora_se3_act(const double a[7], const double p[3], double out[3]) {
  double r[3];
  quat_rotate(a, p, r);
  for (int i = 0; i < 3; ++i) out[i] = r[i] + a[4 + i];
}
/* getLocalPoints3D (src/Map3D.cpp:76-97): camera-frame point of every key point from the depth image
 * (depth_frame.at<float>(trunc(v), trunc(u)), pixels are float), and, when pose7 is given, the world-frame
 * landmark of addNewLandmark (src/Map3D.cpp:44): T_w_c * p_local.  Returns -1 if a pixel is outside. */
int ora_backproject(int n, const float *uv2f, const float *depth_img, int width, int height, const double intr4[4],
                    const double *pose7, double *local3, double *world3) {
  for (int i = 0; i < n; ++i) {
    const double u = (double)uv2f[2 * i], v = (double)uv2f[2 * i + 1];
    const int col = (int)truncf(uv2f[2 * i]), row = (int)truncf(uv2f[2 * i + 1]);
    if (col < 0 || col >= width || row < 0 || row >= height) return -1;
    const double z = (double)depth_img[(size_t)row * width + col];
    double l[3];
    l[0] = z * (u - intr4[2]) / intr4[0];
    l[1] = z * (v - intr4[3]) / intr4[1];
    l[2] = z;
    if (local3) memcpy(local3 + 3 * (size_t)i, l, sizeof(l));
    if (world3 && pose7) ora_se3_act(pose7, l, world3 + 3 * (size_t)i);
  }
  return 0;
}
/* SE3::exp  se3.hpp:725-746 with SO3::expAndTheta so3.hpp:537-571 */
void ora_se3_exp(const double d[6], double out[7]) {
  const double *ups = d, *om = d + 3;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double th = sqrt(th2);
  const double half = 0.5 * th;
  double imag, real;
  if (th < SOPHUS_EPS) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    const double sh = sin(half);
    imag = sh / th;
    real = cos(half);
  }
  double q[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  const double Om[9] = {0.0, -om[2], om[1], om[2], 0.0, -om[0], -om[1], om[0], 0.0};
  double Om2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Om2[i * 3 + j] = (Om[i * 3 + 0] * Om[0 * 3 + j] + Om[i * 3 + 1] * Om[1 * 3 + j]) +
                       Om[i * 3 + 2] * Om[2 * 3 + j];
  double V[9];
  if (th < SOPHUS_EPS) {
    quat_to_R(q, V);
  } else {
    const double c1 = (1.0 - cos(th)) / th2;
    const double c2 = (th - sin(th)) / (th2 * th);
    for (int i = 0; i < 9; ++i)
      V[i] = ((i % 4 == 0 ? 1.0 : 0.0) + c1 * Om[i]) + c2 * Om2[i];
  }
  memcpy(out, q, sizeof(q));
  for (int i = 0; i < 3; ++i)
    out[4 + i] = (V[i * 3 + 0] * ups[0] + V[i * 3 + 1] * ups[1]) + V[i * 3 + 2] * ups[2];
}
/* SE3::Dx_this_mul_exp_x_at_0  se3.hpp:113-182. 7x6 row-major; rows follow the
 * storage order (qx,qy,qz,qw,tx,ty,tz), columns (upsilon, omega): quaternion
 * rows = 0.5 * q (x) (0,e_k), translation rows = R(q). */
void ora_se3_dx_this_mul_exp_x_at_0(const double a[7], double J[42]) {
  const double x = a[0], y = a[1], z = a[2], w = a[3];
  memset(J, 0, 42 * sizeof(double));
  const double hw = 0.5 * w, hx = 0.5 * x, hy = 0.5 * y, hz = 0.5 * z;
  /* d q / d omega */
  J[0 * 6 + 3] = hw;  J[0 * 6 + 4] = -hz; J[0 * 6 + 5] = hy;
  J[1 * 6 + 3] = hz;  J[1 * 6 + 4] = hw;  J[1 * 6 + 5] = -hx;
  J[2 * 6 + 3] = -hy; J[2 * 6 + 4] = hx;  J[2 * 6 + 5] = hw;
  J[3 * 6 + 3] = -hx; J[3 * 6 + 4] = -hy; J[3 * 6 + 5] = -hz;
  /* d t / d upsilon = R(q) written with squares as Sophus does */
  const double ww = w * w, xx = x * x, yy = y * y, zz = z * z;
  const double wz2 = 2.0 * w * z, xy2 = 2.0 * x * y, wy2 = 2.0 * w * y;
  const double xz2 = 2.0 * x * z, wx2 = 2.0 * w * x, yz2 = 2.0 * y * z;
  J[4 * 6 + 0] = -yy + -zz + ww + xx;
  J[4 * 6 + 1] = -wz2 + xy2;
  J[4 * 6 + 2] = wy2 + xz2;
  J[5 * 6 + 0] = wz2 + xy2;
  J[5 * 6 + 1] = -zz + (ww - xx) + yy;
  J[5 * 6 + 2] = -wx2 + yz2;
  J[6 * 6 + 0] = -wy2 + xz2;
  J[6 * 6 + 1] = wx2 + yz2;
  J[6 * 6 + 2] = -yy + zz + (ww - xx);
}

/* ======================================================================
 * Problem evaluation (ceres::ResidualBlock::Evaluate + ProgramEvaluator)
 * ====================================================================== */
typedef struct {
  int R;           /* residual rows per observation: 2 or 3 */
  int nk;          /* 4 if intrinsics are free else 0 */
  double w_repr, w_unpr;
  int *cam_rowptr; /* [n_cam+1] */
  int *pt_rowptr;  /* [n_pt+1] */
  int *perm;       /* [n_obs] point-major -> canonical index */
  int *cam_slot;   /* [n_cam] index among free cameras or -1 */
  int n_free;
} ora_layout;

static int build_layout(const ora_problem *p, const ora_options *o,
                        ora_layout *L) {
  L->R = 2 + (o->use_depth_prior ? 1 : 0);
  if (o->use_depth_prior && !p->depth) return -3;
  L->nk = o->optimize_intrinsics ? 4 : 0;
  const int64_t n = o->n_obs_total > 0 ? o->n_obs_total : (int64_t)p->n_obs;
  /* src/OptimizationUtils.cpp:280, 290: 1.0/admissible_obs, WEIGHT_UNPR/N */
  L->w_repr = 1.0 / (double)n;
  L->w_unpr = o->weight_unpr / (double)n;
  L->cam_rowptr = (int *)calloc((size_t)p->n_cam + 1, sizeof(int));
  L->pt_rowptr = (int *)calloc((size_t)p->n_pt + 1, sizeof(int));
  L->perm = (int *)malloc(sizeof(int) * (size_t)(p->n_obs > 0 ? p->n_obs : 1));
  L->cam_slot = (int *)malloc(sizeof(int) * (size_t)(p->n_cam > 0 ? p->n_cam : 1));
  for (int i = 0; i < p->n_obs; ++i) {
    const int c = p->cam_idx[i], q = p->pt_idx[i];
    if (c < 0 || c >= p->n_cam || q < 0 || q >= p->n_pt) return -2;
    if (i > 0 && c < p->cam_idx[i - 1]) return -2; /* must be camera-sorted */
    L->cam_rowptr[c + 1]++;
    L->pt_rowptr[q + 1]++;
  }
  for (int c = 0; c < p->n_cam; ++c) L->cam_rowptr[c + 1] += L->cam_rowptr[c];
  for (int q = 0; q < p->n_pt; ++q) L->pt_rowptr[q + 1] += L->pt_rowptr[q];
  int *cursor = (int *)malloc(sizeof(int) * (size_t)(p->n_pt > 0 ? p->n_pt : 1));
  for (int q = 0; q < p->n_pt; ++q) cursor[q] = L->pt_rowptr[q];
  for (int i = 0; i < p->n_obs; ++i) L->perm[cursor[p->pt_idx[i]]++] = i;
  free(cursor);
  L->n_free = 0;
  for (int c = 0; c < p->n_cam; ++c)
    L->cam_slot[c] = (c == p->fixed_cam) ? -1 : L->n_free++;
  return 0;
}
static void free_layout(ora_layout *L) {
  free(L->cam_rowptr);
  free(L->pt_rowptr);
  free(L->perm);
  free(L->cam_slot);
}

void ora_build_indices(const ora_problem *p, int32_t *perm, int32_t *pt_rowptr,
                       int32_t *cam_rowptr) {
  ora_options o;
  ora_default_options(&o);
  o.use_depth_prior = 0;
  ora_layout L;
  if (build_layout(p, &o, &L) == 0) {
    if (perm) memcpy(perm, L.perm, sizeof(int) * (size_t)p->n_obs);
    if (pt_rowptr) memcpy(pt_rowptr, L.pt_rowptr, sizeof(int) * ((size_t)p->n_pt + 1));
    if (cam_rowptr) memcpy(cam_rowptr, L.cam_rowptr, sizeof(int) * ((size_t)p->n_cam + 1));
  }
  free_layout(&L);
}

/* Jacobian store. Rows of one observation are contiguous. */
typedef struct {
  double *r;  /* [n_obs*R] */
  double *Jc; /* [n_obs*R*6] local pose Jacobian (zero rows never used for the fixed cam) */
  double *Jp; /* [n_obs*R*3] */
  double *Jk; /* [n_obs*2*4] reprojection rows only */
  double rk[4];
  double Jkk[4]; /* prior Jacobian = diag(Jkk); column-scaled with the rest of J */
  double *cam_cost; /* [n_cam] per-camera cost partials (deterministic sum) */
} ora_jac;

static void alloc_jac(ora_jac *J, const ora_problem *p, const ora_layout *L) {
  const size_t n = (size_t)(p->n_obs > 0 ? p->n_obs : 1);
  J->r = (double *)malloc(sizeof(double) * n * L->R);
  J->Jc = (double *)malloc(sizeof(double) * n * L->R * 6);
  J->Jp = (double *)malloc(sizeof(double) * n * L->R * 3);
  J->Jk = (double *)malloc(sizeof(double) * n * 8);
  J->cam_cost = (double *)malloc(sizeof(double) * (size_t)(p->n_cam > 0 ? p->n_cam : 1));
}
static void free_jac(ora_jac *J) {
  free(J->r);
  free(J->Jc);
  free(J->Jp);
  free(J->Jk);
  free(J->cam_cost);
}

/* Residuals + local Jacobians + loss correction for every residual block, in
 * the order the reference adds them (src/OptimizationUtils.cpp:236-294).
 * Returns cost via *cost; -1 if anything is non-finite. */
static int evaluate_full(const ora_problem *p, const ora_options *o,
                         const ora_layout *L, const double *pose7,
                         const double *pt3, const double intr[4], ora_jac *J,
                         double *cost) {
  const int R = L->R;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int c = 0; c < p->n_cam; ++c) {
    double Jplus[42];
    const double *pose = pose7 + 7 * c;
    ora_se3_dx_this_mul_exp_x_at_0(pose, Jplus);
    double ccost = 0.0;
    for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
      const double *pt = pt3 + 3 * p->pt_idx[i];
      jet jp[7], jx[3], jk[4], res[3];
      make_jets(pose, pt, intr, jp, jx, jk);
      reprojection_jet(jp, jx, jk, p->uv2 + 2 * i, L->w_repr, res);
      if (R == 3) depth_jet(jp, jx, p->depth[i], L->w_unpr, res + 2);
      for (int row = 0; row < R; ++row) {
        double *jc = J->Jc + ((size_t)i * R + row) * 6;
        double *jq = J->Jp + ((size_t)i * R + row) * 3;
        J->r[(size_t)i * R + row] = res[row].a;
        /* residual_block.cc: J_local = J_ambient(.x7) * PlusJacobian(7x6) */
        for (int k = 0; k < 6; ++k) {
          double acc = 0.0;
          for (int m = 0; m < 7; ++m) acc += res[row].v[m] * Jplus[m * 6 + k];
          jc[k] = acc;
        }
        for (int k = 0; k < 3; ++k) jq[k] = res[row].v[7 + k];
        if (row < 2)
          for (int k = 0; k < 4; ++k)
            J->Jk[((size_t)i * 2 + row) * 4 + k] = res[row].v[10 + k];
      }
      /* loss + Corrector (corrector.cc; rho2<=0 branch: scale by sqrt(rho1)) */
      {
        double *r = J->r + (size_t)i * R;
        double rho[3];
        const double s = r[0] * r[0] + r[1] * r[1];
        ora_huber(o->huber_repr, s, rho);
        ccost += 0.5 * rho[0];
        if (!(rho[2] == 0.0 && rho[1] == 1.0)) {
          const double sr = sqrt(rho[1]);
          for (int row = 0; row < 2; ++row) {
            for (int k = 0; k < 6; ++k) J->Jc[((size_t)i * R + row) * 6 + k] *= sr;
            for (int k = 0; k < 3; ++k) J->Jp[((size_t)i * R + row) * 3 + k] *= sr;
            for (int k = 0; k < 4; ++k) J->Jk[((size_t)i * 2 + row) * 4 + k] *= sr;
            r[row] *= sr;
          }
        }
        if (R == 3) {
          const double s2 = r[2] * r[2];
          ora_huber(o->huber_unpr, s2, rho);
          ccost += 0.5 * rho[0];
          if (!(rho[2] == 0.0 && rho[1] == 1.0)) {
            const double sr = sqrt(rho[1]);
            for (int k = 0; k < 6; ++k) J->Jc[((size_t)i * R + 2) * 6 + k] *= sr;
            for (int k = 0; k < 3; ++k) J->Jp[((size_t)i * R + 2) * 3 + k] *= sr;
            r[2] *= sr;
          }
        }
        for (int row = 0; row < R; ++row)
          if (!isfinite(r[row])) bad |= 1;
        for (int k = 0; k < R * 6; ++k)
          if (!isfinite(J->Jc[(size_t)i * R * 6 + k])) bad |= 1;
      }
    }
    J->cam_cost[c] = ccost;
  }
  double total = 0.0;
  if (L->nk) {
    /* IntrinsicsPrior, squared loss (nullptr)  src/OptimizationUtils.cpp:237-241 */
    double Jkk[16];
    ora_intrinsics_prior(intr, p->intr_prior, o->weight_intrinsics, J->rk, Jkk);
    for (int i = 0; i < 4; ++i) J->Jkk[i] = Jkk[i * 4 + i];
    double s = 0.0;
    for (int i = 0; i < 4; ++i) s += J->rk[i] * J->rk[i];
    total += 0.5 * s;
  } else {
    memset(J->Jkk, 0, sizeof(J->Jkk));
    memset(J->rk, 0, sizeof(J->rk));
  }
  for (int c = 0; c < p->n_cam; ++c) total += J->cam_cost[c];
  *cost = total;
  if (bad || !isfinite(total)) return -1;
  return 0;
}

/* Cost-only evaluation (T = double path of the functors). */
static int evaluate_cost(const ora_problem *p, const ora_options *o,
                         const ora_layout *L, const double *pose7,
                         const double *pt3, const double intr[4],
                         double *cam_cost, double *cost) {
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < p->n_cam; ++c) {
    const double *pose = pose7 + 7 * c;
    double ccost = 0.0;
    for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
      const double *pt = pt3 + 3 * p->pt_idx[i];
      double r[3], rho[3];
      reprojection_val(pose, pt, intr, p->uv2 + 2 * i, L->w_repr, r);
      ora_huber(o->huber_repr, r[0] * r[0] + r[1] * r[1], rho);
      ccost += 0.5 * rho[0];
      if (R == 3) {
        depth_val(pose, pt, p->depth[i], L->w_unpr, r + 2);
        ora_huber(o->huber_unpr, r[2] * r[2], rho);
        ccost += 0.5 * rho[0];
      }
    }
    cam_cost[c] = ccost;
  }
  double total = 0.0;
  if (L->nk) {
    double rk[4], s = 0.0;
    ora_intrinsics_prior(intr, p->intr_prior, o->weight_intrinsics, rk, NULL);
    for (int i = 0; i < 4; ++i) s += rk[i] * rk[i];
    total += 0.5 * s;
  }
  for (int c = 0; c < p->n_cam; ++c) total += cam_cost[c];
  *cost = total;
  return isfinite(total) ? 0 : -1;
}

/* gradient g = J^T r of the (unscaled or scaled) store */
static void compute_gradient(const ora_problem *p, const ora_layout *L,
                             const ora_jac *J, double *g_c, double *g_p,
                             double *g_k) {
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < p->n_cam; ++c) {
    double g[6] = {0, 0, 0, 0, 0, 0};
    if (L->cam_slot[c] >= 0)
      for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i)
        for (int row = 0; row < R; ++row) {
          const double r = J->r[(size_t)i * R + row];
          const double *jc = J->Jc + ((size_t)i * R + row) * 6;
          for (int k = 0; k < 6; ++k) g[k] += jc[k] * r;
        }
    memcpy(g_c + 6 * c, g, sizeof(g));
  }
#pragma omp parallel for schedule(static)
  for (int q = 0; q < p->n_pt; ++q) {
    double g[3] = {0, 0, 0};
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int i = L->perm[s];
      for (int row = 0; row < R; ++row) {
        const double r = J->r[(size_t)i * R + row];
        const double *jq = J->Jp + ((size_t)i * R + row) * 3;
        for (int k = 0; k < 3; ++k) g[k] += jq[k] * r;
      }
    }
    memcpy(g_p + 3 * q, g, sizeof(g));
  }
  if (g_k) {
    double g[4] = {0, 0, 0, 0};
    if (L->nk) {
      for (int i = 0; i < p->n_obs; ++i)
        for (int row = 0; row < 2; ++row) {
          const double r = J->r[(size_t)i * R + row];
          const double *jk = J->Jk + ((size_t)i * 2 + row) * 4;
          for (int k = 0; k < 4; ++k) g[k] += jk[k] * r;
        }
      for (int k = 0; k < 4; ++k) g[k] += J->Jkk[k] * J->rk[k];
    }
    memcpy(g_k, g, sizeof(g));
  }
}

int ora_evaluate(const ora_problem *p, const ora_options *o, double *cost,
                 double *r, double *Jc, double *Jp, double *Jk, double *g_c,
                 double *g_p, double *g_k) {
#ifdef _OPENMP
  omp_set_num_threads(o->num_threads > 0 ? o->num_threads : 1);
#endif
  ora_layout L;
  int rc = build_layout(p, o, &L);
  if (rc) {
    free_layout(&L);
    return rc;
  }
  ora_jac J;
  alloc_jac(&J, p, &L);
  double c = 0.0;
  rc = evaluate_full(p, o, &L, p->pose7, p->pt3, p->intr, &J, &c);
  const size_t n = (size_t)p->n_obs;
  if (cost) *cost = c;
  if (r) memcpy(r, J.r, sizeof(double) * n * L.R);
  if (Jc) memcpy(Jc, J.Jc, sizeof(double) * n * L.R * 6);
  if (Jp) memcpy(Jp, J.Jp, sizeof(double) * n * L.R * 3);
  if (Jk) memcpy(Jk, J.Jk, sizeof(double) * n * 8);
  if (g_c || g_p || g_k) {
    double *gc = g_c ? g_c : (double *)malloc(sizeof(double) * 6 * (size_t)(p->n_cam + 1));
    double *gp = g_p ? g_p : (double *)malloc(sizeof(double) * 3 * (size_t)(p->n_pt + 1));
    compute_gradient(p, &L, &J, gc, gp, g_k);
    if (!g_c) free(gc);
    if (!g_p) free(gp);
  }
  free_jac(&J);
  free_layout(&L);
  return rc;
}

/* ======================================================================
 * Linear algebra helpers
 * ====================================================================== */
/* In-place dense Cholesky A = L L^T (lower, row-major n x n). 0 ok, -1 not SPD */
static int chol_factor(double *A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !isfinite(d)) return -1;
    d = sqrt(d);
    A[(size_t)j * n + j] = d;
    const double inv = 1.0 / d;
#pragma omp parallel for schedule(static) if (n - j > 256)
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s * inv;
    }
  }
  return 0;
}
static void chol_solve(const double *L, int n, double *b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}
/* small SPD inverse via Cholesky: ceres InvertPSDMatrix (llt().solve(I)) */
static int spd_inverse(const double *A, int n, double *Ainv) {
  double L[36], col[6];
  memcpy(L, A, sizeof(double) * (size_t)n * n);
  if (chol_factor(L, n)) return -1;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    chol_solve(L, n, col);
    for (int i = 0; i < n; ++i) Ainv[i * n + j] = col[i];
  }
  return 0;
}

/* ======================================================================
 * LM workspace
 * ====================================================================== */
typedef struct {
  const ora_problem *p;
  const ora_options *o;
  ora_layout L;
  ora_jac J; /* scaled in place once jacobi scaling is known */
  /* per-column Jacobi scale */
  double *sc, *sp, sk[4];
  /* LM diagonal (clamped squared column norms of the scaled J) */
  double *dc, *dp, dk[4];
  /* gradient of the scaled system */
  double *gc, *gp, gk[4];
  /* step (scaled) */
  double *yc, *yp, yk[4];
  /* point blocks */
  double *Vinv; /* [n_pt*9] */
  int have_scale;
} ora_ws;

/* squared column norms of the current J store */
static void column_sqnorms(const ora_ws *w, double *nc, double *np, double nk[4]) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < p->n_cam; ++c) {
    double n[6] = {0, 0, 0, 0, 0, 0};
    if (L->cam_slot[c] >= 0)
      for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i)
        for (int row = 0; row < R; ++row) {
          const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
          for (int k = 0; k < 6; ++k) n[k] += jc[k] * jc[k];
        }
    memcpy(nc + 6 * c, n, sizeof(n));
  }
#pragma omp parallel for schedule(static)
  for (int q = 0; q < p->n_pt; ++q) {
    double n[3] = {0, 0, 0};
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int i = L->perm[s];
      for (int row = 0; row < R; ++row) {
        const double *jq = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int k = 0; k < 3; ++k) n[k] += jq[k] * jq[k];
      }
    }
    memcpy(np + 3 * q, n, sizeof(n));
  }
  for (int k = 0; k < 4; ++k) nk[k] = 0.0;
  if (L->nk) {
    for (int i = 0; i < p->n_obs; ++i)
      for (int row = 0; row < 2; ++row) {
        const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
        for (int k = 0; k < 4; ++k) nk[k] += jk[k] * jk[k];
      }
    for (int k = 0; k < 4; ++k) nk[k] += w->J.Jkk[k] * w->J.Jkk[k];
  }
}

/* jacobian->ScaleColumns(scale) */
static void scale_columns(ora_ws *w) {
  const ora_problem *p = w->p;
  const int R = w->L.R;
  if (w->L.nk)
    for (int k = 0; k < 4; ++k) w->J.Jkk[k] *= w->sk[k];
#pragma omp parallel for schedule(static)
  for (int i = 0; i < p->n_obs; ++i) {
    const double *sc = w->sc + 6 * p->cam_idx[i];
    const double *sp = w->sp + 3 * p->pt_idx[i];
    for (int row = 0; row < R; ++row) {
      double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
      double *jq = w->J.Jp + ((size_t)i * R + row) * 3;
      for (int k = 0; k < 6; ++k) jc[k] *= sc[k];
      for (int k = 0; k < 3; ++k) jq[k] *= sp[k];
      if (row < 2 && w->L.nk) {
        double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
        for (int k = 0; k < 4; ++k) jk[k] *= w->sk[k];
      }
    }
  }
}

/* Evaluate at x, build unscaled gradient + gradient max norm, then scale J.
 * trust_region_minimizer.cc: EvaluateGradientAndJacobian. */
static int evaluate_gradient_and_jacobian(ora_ws *w, const double *pose7,
                                          const double *pt3,
                                          const double intr[4], double *cost,
                                          double *gmax) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  if (evaluate_full(p, w->o, L, pose7, pt3, intr, &w->J, cost)) return -1;
  compute_gradient(p, L, &w->J, w->gc, w->gp, w->gk);
  /* |Plus(x, -g) - x|_inf in the ambient space */
  double m = 0.0;
  for (int c = 0; c < p->n_cam; ++c) {
    if (L->cam_slot[c] < 0) continue;
    double d[6], e[7], t[7];
    for (int k = 0; k < 6; ++k) d[k] = -w->gc[6 * c + k];
    ora_se3_exp(d, e);
    ora_se3_mul(pose7 + 7 * c, e, t);
    for (int k = 0; k < 7; ++k) m = fmax(m, fabs(pose7[7 * c + k] - t[k]));
  }
  for (int i = 0; i < 3 * p->n_pt; ++i) m = fmax(m, fabs(w->gp[i]));
  for (int i = 0; i < L->nk; ++i) m = fmax(m, fabs(w->gk[i]));
  *gmax = m;
  if (w->o->jacobi_scaling) {
    if (!w->have_scale) {
      column_sqnorms(w, w->sc, w->sp, w->sk);
      for (int i = 0; i < 6 * p->n_cam; ++i) w->sc[i] = 1.0 / (1.0 + sqrt(w->sc[i]));
      for (int i = 0; i < 3 * p->n_pt; ++i) w->sp[i] = 1.0 / (1.0 + sqrt(w->sp[i]));
      for (int i = 0; i < 4; ++i) w->sk[i] = 1.0 / (1.0 + sqrt(w->sk[i]));
      w->have_scale = 1;
    }
    scale_columns(w);
    /* gradient of the scaled system = scale .* g */
    for (int i = 0; i < 6 * p->n_cam; ++i) w->gc[i] *= w->sc[i];
    for (int i = 0; i < 3 * p->n_pt; ++i) w->gp[i] *= w->sp[i];
    for (int i = 0; i < 4; ++i) w->gk[i] *= w->sk[i];
  } else if (!w->have_scale) {
    for (int i = 0; i < 6 * p->n_cam; ++i) w->sc[i] = 1.0;
    for (int i = 0; i < 3 * p->n_pt; ++i) w->sp[i] = 1.0;
    for (int i = 0; i < 4; ++i) w->sk[i] = 1.0;
    w->have_scale = 1;
  }
  return 0;
}

/* levenberg_marquardt_strategy.cc: diagonal_ = clamp(SquaredColumnNorm) */
static void lm_diagonal(ora_ws *w) {
  const ora_options *o = w->o;
  column_sqnorms(w, w->dc, w->dp, w->dk);
  for (int i = 0; i < 6 * w->p->n_cam; ++i)
    w->dc[i] = fmin(fmax(w->dc[i], o->min_lm_diagonal), o->max_lm_diagonal);
  for (int i = 0; i < 3 * w->p->n_pt; ++i)
    w->dp[i] = fmin(fmax(w->dp[i], o->min_lm_diagonal), o->max_lm_diagonal);
  for (int i = 0; i < 4; ++i)
    w->dk[i] = fmin(fmax(w->dk[i], o->min_lm_diagonal), o->max_lm_diagonal);
}

/* per point: V = sum E^T E + D_p^2, Vinv (schur_eliminator_impl.h: ete) */
static int point_blocks(ora_ws *w, double radius) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int q = 0; q < p->n_pt; ++q) {
    double V[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int i = L->perm[s];
      for (int row = 0; row < R; ++row) {
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) V[a * 3 + b] += e[a] * e[b];
      }
    }
    for (int a = 0; a < 3; ++a) {
      const double D = sqrt(w->dp[3 * q + a] / radius);
      V[a * 3 + a] += D * D;
    }
    if (spd_inverse(V, 3, w->Vinv + 9 * (size_t)q)) bad |= 1;
  }
  return bad ? -1 : 0;
}

/* ---------- DENSE_SCHUR / SPARSE_SCHUR-equivalent exact step ---------- */
static int solve_dense_schur(ora_ws *w, double radius) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R, nk = L->nk;
  const int n = 6 * L->n_free + nk;
  const int koff = 6 * L->n_free;
  double *S = (double *)calloc((size_t)n * n + 1, sizeof(double));
  double *rhs = (double *)calloc((size_t)n + 1, sizeof(double));
  /* F^T F + D_f^2 and -F^T r */
  for (int i = 0; i < p->n_obs; ++i) {
    const int slot = L->cam_slot[p->cam_idx[i]];
    for (int row = 0; row < R; ++row) {
      double f[10];
      int col[10], m = 0;
      if (slot >= 0)
        for (int k = 0; k < 6; ++k) {
          f[m] = w->J.Jc[((size_t)i * R + row) * 6 + k];
          col[m++] = 6 * slot + k;
        }
      if (nk && row < 2)
        for (int k = 0; k < 4; ++k) {
          f[m] = w->J.Jk[((size_t)i * 2 + row) * 4 + k];
          col[m++] = koff + k;
        }
      const double r = w->J.r[(size_t)i * R + row];
      for (int a = 0; a < m; ++a) {
        rhs[col[a]] -= f[a] * r;
        for (int b = 0; b < m; ++b) S[(size_t)col[a] * n + col[b]] += f[a] * f[b];
      }
    }
  }
  if (nk)
    for (int k = 0; k < 4; ++k) {
      S[(size_t)(koff + k) * n + koff + k] += w->J.Jkk[k] * w->J.Jkk[k];
      rhs[koff + k] -= w->J.Jkk[k] * w->J.rk[k];
    }
  for (int c = 0; c < p->n_cam; ++c) {
    const int slot = L->cam_slot[c];
    if (slot < 0) continue;
    for (int k = 0; k < 6; ++k) {
      const double D = sqrt(w->dc[6 * c + k] / radius);
      S[(size_t)(6 * slot + k) * n + 6 * slot + k] += D * D;
    }
  }
  for (int k = 0; k < nk; ++k) {
    const double D = sqrt(w->dk[k] / radius);
    S[(size_t)(koff + k) * n + koff + k] += D * D;
  }
  if (point_blocks(w, radius)) {
    free(S);
    free(rhs);
    return -1;
  }
  /* eliminate every point (chunk) */
  int maxdeg = 1;
  for (int q = 0; q < p->n_pt; ++q) {
    const int d = L->pt_rowptr[q + 1] - L->pt_rowptr[q];
    if (d > maxdeg) maxdeg = d;
  }
  double *Wb = (double *)malloc(sizeof(double) * (size_t)(maxdeg + 1) * 30);
  for (int q = 0; q < p->n_pt; ++q) {
    const double *Vi = w->Vinv + 9 * (size_t)q;
    const int s0 = L->pt_rowptr[q], s1 = L->pt_rowptr[q + 1];
    const int deg = s1 - s0;
    /* W blocks: camera part 6x3 per obs, intrinsics part 4x3 summed */
    double Wk[12];
    memset(Wk, 0, sizeof(Wk));
    for (int s = s0; s < s1; ++s) {
      const int i = L->perm[s];
      double *W6 = Wb + (size_t)(s - s0) * 18;
      memset(W6, 0, sizeof(double) * 18);
      for (int row = 0; row < R; ++row) {
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b < 3; ++b) W6[a * 3 + b] += jc[a] * e[b];
        if (nk && row < 2) {
          const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
          for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 3; ++b) Wk[a * 3 + b] += jk[a] * e[b];
        }
      }
    }
    const double *gq = w->gp + 3 * q;
    double Vg[3];
    for (int a = 0; a < 3; ++a)
      Vg[a] = Vi[a * 3 + 0] * gq[0] + Vi[a * 3 + 1] * gq[1] + Vi[a * 3 + 2] * gq[2];
    /* blocks: deg camera blocks + optional intrinsics block */
    const int nb = deg + (nk ? 1 : 0);
    for (int a = 0; a < nb; ++a) {
      const double *Wa;
      int ra, offa;
      if (a < deg) {
        const int slot = L->cam_slot[p->cam_idx[L->perm[s0 + a]]];
        if (slot < 0) continue;
        Wa = Wb + (size_t)a * 18;
        ra = 6;
        offa = 6 * slot;
      } else {
        Wa = Wk;
        ra = 4;
        offa = koff;
      }
      double WV[18]; /* Wa * Vinv  (ra x 3) */
      for (int x = 0; x < ra; ++x)
        for (int y = 0; y < 3; ++y)
          WV[x * 3 + y] = Wa[x * 3 + 0] * Vi[0 * 3 + y] + Wa[x * 3 + 1] * Vi[1 * 3 + y] +
                          Wa[x * 3 + 2] * Vi[2 * 3 + y];
      /* rhs = -g_f + W Vinv g_e */
      for (int x = 0; x < ra; ++x)
        rhs[offa + x] += Wa[x * 3 + 0] * Vg[0] + Wa[x * 3 + 1] * Vg[1] + Wa[x * 3 + 2] * Vg[2];
      for (int b = 0; b < nb; ++b) {
        const double *Wc;
        int rb, offb;
        if (b < deg) {
          const int slot = L->cam_slot[p->cam_idx[L->perm[s0 + b]]];
          if (slot < 0) continue;
          Wc = Wb + (size_t)b * 18;
          rb = 6;
          offb = 6 * slot;
        } else {
          Wc = Wk;
          rb = 4;
          offb = koff;
        }
        for (int x = 0; x < ra; ++x)
          for (int y = 0; y < rb; ++y)
            S[(size_t)(offa + x) * n + offb + y] -=
                WV[x * 3 + 0] * Wc[y * 3 + 0] + WV[x * 3 + 1] * Wc[y * 3 + 1] +
                WV[x * 3 + 2] * Wc[y * 3 + 2];
      }
    }
  }
  /* rhs currently holds -g_f(from F^T r) ... note g of scaled system: the F^T r
   * accumulation above already used the scaled J, so nothing else to add. */
  int rc = 0;
  if (n > 0) {
    rc = chol_factor(S, n);
    if (!rc) chol_solve(S, n, rhs);
  }
  if (!rc) {
    for (int c = 0; c < p->n_cam; ++c) {
      const int slot = L->cam_slot[c];
      for (int k = 0; k < 6; ++k) w->yc[6 * c + k] = slot >= 0 ? rhs[6 * slot + k] : 0.0;
    }
    for (int k = 0; k < 4; ++k) w->yk[k] = nk ? rhs[koff + k] : 0.0;
  }
  free(Wb);
  free(S);
  free(rhs);
  return rc;
}

/* ---------- SPARSE_SCHUR-equivalent exact step: explicit S in envelope (skyline) storage + sparse Cholesky ----------
 * What the reference configures (headers/BundleAdjustmentConfig.h:62: linear_solver_type = SPARSE_SCHUR): Ceres forms the
 * reduced camera matrix S block-sparse (schur_eliminator_impl.h) and factorises it with a sparse Cholesky.  The sparse
 * backend's ordering does not change the solution (exact solve); this restatement keeps the natural camera order and stores
 * every scalar row from its first structural non-zero to the diagonal (envelope): Cholesky creates no fill outside it.
 * Rows of the 4 intrinsics (REF mode) are a dense border at the end.  Same arithmetic as solve_dense_schur, other storage. */
static int solve_sparse_schur(ora_ws *w, double radius) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R, nk = L->nk;
  const int nf = L->n_free;
  const int n = 6 * nf + nk;
  const int koff = 6 * nf;
  if (n == 0) return 0;
  if (point_blocks(w, radius)) return -1;
  /* envelope: first coupled camera slot of every camera slot */
  int *first = (int *)malloc(sizeof(int) * (size_t)(nf + 1));
  int *slot_cam = (int *)malloc(sizeof(int) * (size_t)(nf + 1));
  for (int c = 0; c < p->n_cam; ++c)
    if (L->cam_slot[c] >= 0) {
      first[L->cam_slot[c]] = L->cam_slot[c];
      slot_cam[L->cam_slot[c]] = c;
    }
  for (int q = 0; q < p->n_pt; ++q) {
    int lo = -1;
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int sl = L->cam_slot[p->cam_idx[L->perm[s]]];
      if (sl < 0) continue;
      if (lo < 0 || sl < lo) lo = sl;
    }
    if (lo < 0) continue;
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int sl = L->cam_slot[p->cam_idx[L->perm[s]]];
      if (sl >= 0 && lo < first[sl]) first[sl] = lo;
    }
  }
  size_t *rowstart = (size_t *)malloc(sizeof(size_t) * (size_t)(n + 1));
  int *rfirst = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  size_t tot = 0;
  for (int i = 0; i < n; ++i) {
    rfirst[i] = i < koff ? 6 * first[i / 6] : 0;
    rowstart[i] = tot;
    tot += (size_t)(i - rfirst[i] + 1);
  }
  rowstart[n] = tot;
  double *A = (double *)calloc(tot + 1, sizeof(double));
  double *rhs = (double *)calloc((size_t)n + 1, sizeof(double));
#define SKY(i, j) A[rowstart[i] + (size_t)((j) - rfirst[i])]
  /* W_o = Jc_o^T Jp_o (6x3) of every observation */
  double *W = (double *)malloc(sizeof(double) * ((size_t)p->n_obs + 1) * 18);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < p->n_obs; ++i) {
    double *W6 = W + (size_t)i * 18;
    for (int k = 0; k < 18; ++k) W6[k] = 0.0;
    for (int row = 0; row < R; ++row) {
      const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
      const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 3; ++b) W6[a * 3 + b] += jc[a] * e[b];
    }
  }
  /* camera rows: every thread owns the rows of its cameras (no write conflicts, fixed summation order) */
#pragma omp parallel for schedule(dynamic, 16)
  for (int sa = 0; sa < nf; ++sa) {
    const int c = slot_cam[sa];
    for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
      /* F^T F and -F^T r */
      for (int row = 0; row < R; ++row) {
        const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
        const double r = w->J.r[(size_t)i * R + row];
        for (int a = 0; a < 6; ++a) {
          rhs[6 * sa + a] -= jc[a] * r;
          for (int b = 0; b <= a; ++b) SKY(6 * sa + a, 6 * sa + b) += jc[a] * jc[b];
        }
      }
      /* - W_i V^-1 W_j^T for the observations j of the same point with slot_j <= slot_a */
      const int q = p->pt_idx[i];
      const double *Vi = w->Vinv + 9 * (size_t)q;
      const double *Wa = W + (size_t)i * 18;
      double WV[18];
      for (int x = 0; x < 6; ++x)
        for (int y = 0; y < 3; ++y)
          WV[x * 3 + y] = Wa[x * 3 + 0] * Vi[0 * 3 + y] + Wa[x * 3 + 1] * Vi[1 * 3 + y] + Wa[x * 3 + 2] * Vi[2 * 3 + y];
      const double *gq = w->gp + 3 * q;
      double Vg[3];
      for (int a = 0; a < 3; ++a) Vg[a] = Vi[a * 3 + 0] * gq[0] + Vi[a * 3 + 1] * gq[1] + Vi[a * 3 + 2] * gq[2];
      for (int x = 0; x < 6; ++x) rhs[6 * sa + x] += Wa[x * 3 + 0] * Vg[0] + Wa[x * 3 + 1] * Vg[1] + Wa[x * 3 + 2] * Vg[2];
      for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
        const int j = L->perm[s];
        const int sb = L->cam_slot[p->cam_idx[j]];
        if (sb < 0 || sb > sa) continue;
        const double *Wb = W + (size_t)j * 18;
        for (int x = 0; x < 6; ++x) {
          const int ymax = sb == sa ? x : 5;
          for (int y = 0; y <= ymax; ++y)
            SKY(6 * sa + x, 6 * sb + y) -= WV[x * 3 + 0] * Wb[y * 3 + 0] + WV[x * 3 + 1] * Wb[y * 3 + 1] + WV[x * 3 + 2] * Wb[y * 3 + 2];
        }
      }
    }
    for (int k = 0; k < 6; ++k) {
      const double D = sqrt(w->dc[6 * c + k] / radius);
      SKY(6 * sa + k, 6 * sa + k) += D * D;
    }
  }
  if (nk) {
    /* border rows of the 4 intrinsics columns (sequential: one owner) */
    for (int i = 0; i < p->n_obs; ++i) {
      const int sl = L->cam_slot[p->cam_idx[i]];
      for (int row = 0; row < 2; ++row) {
        const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
        const double r = w->J.r[(size_t)i * R + row];
        for (int a = 0; a < 4; ++a) {
          rhs[koff + a] -= jk[a] * r;
          for (int b = 0; b <= a; ++b) SKY(koff + a, koff + b) += jk[a] * jk[b];
          if (sl >= 0) {
            const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
            for (int b = 0; b < 6; ++b) SKY(koff + a, 6 * sl + b) += jk[a] * jc[b];
          }
        }
      }
    }
    for (int k = 0; k < 4; ++k) {
      SKY(koff + k, koff + k) += w->J.Jkk[k] * w->J.Jkk[k];
      rhs[koff + k] -= w->J.Jkk[k] * w->J.rk[k];
      const double D = sqrt(w->dk[k] / radius);
      SKY(koff + k, koff + k) += D * D;
    }
    for (int q = 0; q < p->n_pt; ++q) {
      const double *Vi = w->Vinv + 9 * (size_t)q;
      double Wk[12], WkV[12];
      memset(Wk, 0, sizeof(Wk));
      for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
        const int i = L->perm[s];
        for (int row = 0; row < 2; ++row) {
          const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
          const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
          for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 3; ++b) Wk[a * 3 + b] += jk[a] * e[b];
        }
      }
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 3; ++y)
          WkV[x * 3 + y] = Wk[x * 3 + 0] * Vi[0 * 3 + y] + Wk[x * 3 + 1] * Vi[1 * 3 + y] + Wk[x * 3 + 2] * Vi[2 * 3 + y];
      const double *gq = w->gp + 3 * q;
      double Vg[3];
      for (int a = 0; a < 3; ++a) Vg[a] = Vi[a * 3 + 0] * gq[0] + Vi[a * 3 + 1] * gq[1] + Vi[a * 3 + 2] * gq[2];
      for (int x = 0; x < 4; ++x) {
        rhs[koff + x] += Wk[x * 3 + 0] * Vg[0] + Wk[x * 3 + 1] * Vg[1] + Wk[x * 3 + 2] * Vg[2];
        for (int y = 0; y <= x; ++y)
          SKY(koff + x, koff + y) -= WkV[x * 3 + 0] * Wk[y * 3 + 0] + WkV[x * 3 + 1] * Wk[y * 3 + 1] + WkV[x * 3 + 2] * Wk[y * 3 + 2];
      }
      for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
        const int j = L->perm[s];
        const int sb = L->cam_slot[p->cam_idx[j]];
        if (sb < 0) continue;
        const double *Wb = W + (size_t)j * 18;
        for (int x = 0; x < 4; ++x)
          for (int y = 0; y < 6; ++y)
            SKY(koff + x, 6 * sb + y) -= WkV[x * 3 + 0] * Wb[y * 3 + 0] + WkV[x * 3 + 1] * Wb[y * 3 + 1] + WkV[x * 3 + 2] * Wb[y * 3 + 2];
      }
    }
  }
  free(W);
  /* envelope Cholesky, row by row: L_ij = (A_ij - sum_{k >= max(f_i, f_j)}^{j-1} L_ik L_jk) / L_jj */
  int rc = 0;
  for (int i = 0; i < n && !rc; ++i) {
    const int fi = rfirst[i];
    double *Li = A + rowstart[i];
    for (int j = fi; j <= i; ++j) {
      const int fj = rfirst[j];
      const int k0 = fi > fj ? fi : fj;
      const double *Lj = A + rowstart[j];
      double sacc = Li[j - fi];
      const double *a = Li + (k0 - fi), *b = Lj + (k0 - fj);
      const int len = j - k0;
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
      int k = 0;
      for (; k + 3 < len; k += 4) {
        d0 += a[k] * b[k];
        d1 += a[k + 1] * b[k + 1];
        d2 += a[k + 2] * b[k + 2];
        d3 += a[k + 3] * b[k + 3];
      }
      for (; k < len; ++k) d0 += a[k] * b[k];
      sacc -= (d0 + d1) + (d2 + d3);
      if (j < i) {
        Li[j - fi] = sacc / Lj[j - fj];
      } else {
        if (!(sacc > 0.0) || !isfinite(sacc)) {
          rc = -1;
          break;
        }
        Li[j - fi] = sqrt(sacc);
      }
    }
  }
  if (!rc) {
    for (int i = 0; i < n; ++i) {
      const int fi = rfirst[i];
      const double *Li = A + rowstart[i];
      double sacc = rhs[i];
      for (int k = fi; k < i; ++k) sacc -= Li[k - fi] * rhs[k];
      rhs[i] = sacc / Li[i - fi];
    }
    for (int i = n - 1; i >= 0; --i) {
      const int fi = rfirst[i];
      const double *Li = A + rowstart[i];
      const double v = rhs[i] / Li[i - fi];
      rhs[i] = v;
      for (int k = fi; k < i; ++k) rhs[k] -= Li[k - fi] * v;
    }
    for (int c = 0; c < p->n_cam; ++c) {
      const int slot = L->cam_slot[c];
      for (int k = 0; k < 6; ++k) w->yc[6 * c + k] = slot >= 0 ? rhs[6 * slot + k] : 0.0;
    }
    for (int k = 0; k < 4; ++k) w->yk[k] = nk ? rhs[koff + k] : 0.0;
  }
#undef SKY
  free(A);
  free(rhs);
  free(rowstart);
  free(rfirst);
  free(first);
  free(slot_cam);
  return rc;
}

/* back substitution: y_p = Vinv (-g_p - sum_o W_o^T y_c - Wk^T y_k) */
static void back_substitute(ora_ws *w) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int q = 0; q < p->n_pt; ++q) {
    double b[3] = {-w->gp[3 * q], -w->gp[3 * q + 1], -w->gp[3 * q + 2]};
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int i = L->perm[s];
      const int c = p->cam_idx[i];
      const int freec = L->cam_slot[c] >= 0;
      for (int row = 0; row < R; ++row) {
        double a = 0.0;
        if (freec) {
          const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
          for (int k = 0; k < 6; ++k) a += jc[k] * w->yc[6 * c + k];
        }
        if (L->nk && row < 2) {
          const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
          for (int k = 0; k < 4; ++k) a += jk[k] * w->yk[k];
        }
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int k = 0; k < 3; ++k) b[k] -= e[k] * a;
      }
    }
    const double *Vi = w->Vinv + 9 * (size_t)q;
    for (int a = 0; a < 3; ++a)
      w->yp[3 * q + a] = Vi[a * 3 + 0] * b[0] + Vi[a * 3 + 1] * b[1] + Vi[a * 3 + 2] * b[2];
  }
}

/* ---------- ITERATIVE_SCHUR-equivalent: implicit Schur + PCG ---------- */
/* t_p = Vinv * sum_{o in p} Jp^T (Jc x_c)  for p in [q0,q1) */
static void pass_points(const ora_ws *w, const double *x, double *t, int q0, int q1) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int q = q0; q < q1; ++q) {
    double b[3] = {0, 0, 0};
    for (int s = L->pt_rowptr[q]; s < L->pt_rowptr[q + 1]; ++s) {
      const int i = L->perm[s];
      const int c = p->cam_idx[i];
      if (L->cam_slot[c] < 0) continue;
      for (int row = 0; row < R; ++row) {
        const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
        double a = 0.0;
        for (int k = 0; k < 6; ++k) a += jc[k] * x[6 * c + k];
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int k = 0; k < 3; ++k) b[k] += e[k] * a;
      }
    }
    const double *Vi = w->Vinv + 9 * (size_t)q;
    for (int a = 0; a < 3; ++a)
      t[3 * q + a] = Vi[a * 3 + 0] * b[0] + Vi[a * 3 + 1] * b[1] + Vi[a * 3 + 2] * b[2];
  }
}
/* y_c = [diag] (U_c + D_c^2) x_c - sum_{o in c, p in [q0,q1)} Jc^T (Jp t_p) */
static void pass_cameras(const ora_ws *w, double radius, const double *x,
                         const double *t, double *y, int q0, int q1, int diag) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < p->n_cam; ++c) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (L->cam_slot[c] >= 0) {
      for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
        const int q = p->pt_idx[i];
        const int inshard = (q >= q0 && q < q1);
        for (int row = 0; row < R; ++row) {
          const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
          const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
          double a = 0.0;
          if (diag)
            for (int k = 0; k < 6; ++k) a += jc[k] * x[6 * c + k];
          if (inshard)
            a -= e[0] * t[3 * q] + e[1] * t[3 * q + 1] + e[2] * t[3 * q + 2];
          for (int k = 0; k < 6; ++k) acc[k] += jc[k] * a;
        }
      }
      if (diag)
        for (int k = 0; k < 6; ++k) {
          const double D = sqrt(w->dc[6 * c + k] / radius);
          acc[k] += D * D * x[6 * c + k];
        }
    } else if (diag) {
      /* fixed camera: not in the Ceres program; keep its rows inert */
      for (int k = 0; k < 6; ++k) acc[k] = x[6 * c + k];
    }
    memcpy(y + 6 * c, acc, sizeof(acc));
  }
}
static void schur_apply(const ora_ws *w, double radius, const double *x,
                        double *t, double *y) {
  pass_points(w, x, t, 0, w->p->n_pt);
  pass_cameras(w, radius, x, t, y, 0, w->p->n_pt, 1);
}
static double dotn(const double *a, const double *b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* SCHUR_JACOBI preconditioner: inverse of the 6x6 diagonal blocks of S */
static int schur_jacobi(const ora_ws *w, double radius, double *Minv) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int c = 0; c < p->n_cam; ++c) {
    double B[36];
    memset(B, 0, sizeof(B));
    if (L->cam_slot[c] < 0) {
      for (int k = 0; k < 6; ++k) B[k * 6 + k] = 1.0;
      memcpy(Minv + 36 * (size_t)c, B, sizeof(B));
      continue;
    }
    for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
      const double *Vi = w->Vinv + 9 * (size_t)p->pt_idx[i];
      double W[18];
      memset(W, 0, sizeof(W));
      for (int row = 0; row < R; ++row) {
        const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int a = 0; a < 6; ++a) {
          for (int b = 0; b < 6; ++b) B[a * 6 + b] += jc[a] * jc[b];
          for (int b = 0; b < 3; ++b) W[a * 3 + b] += jc[a] * e[b];
        }
      }
      double WV[18];
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 3; ++b)
          WV[a * 3 + b] = W[a * 3 + 0] * Vi[0 * 3 + b] + W[a * 3 + 1] * Vi[1 * 3 + b] +
                          W[a * 3 + 2] * Vi[2 * 3 + b];
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b)
          B[a * 6 + b] -= WV[a * 3 + 0] * W[b * 3 + 0] + WV[a * 3 + 1] * W[b * 3 + 1] +
                          WV[a * 3 + 2] * W[b * 3 + 2];
    }
    for (int k = 0; k < 6; ++k) {
      const double D = sqrt(w->dc[6 * c + k] / radius);
      B[k * 6 + k] += D * D;
    }
    if (spd_inverse(B, 6, Minv + 36 * (size_t)c)) bad |= 1;
  }
  return bad ? -1 : 0;
}

/* conjugate_gradients_solver.cc (ceres 2.0.0) on S y = rhs. Returns #iters or <0 */
static int solve_implicit_pcg(ora_ws *w, double radius, int *iters_out) {
  const ora_problem *p = w->p;
  const ora_options *o = w->o;
  const ora_layout *L = &w->L;
  const int n = 6 * p->n_cam;
  *iters_out = 0;
  if (L->nk) return -4; /* free intrinsics only with the dense solver */
  if (point_blocks(w, radius)) return -1;
  double *b = (double *)calloc((size_t)n + 1, sizeof(double));
  double *x = (double *)calloc((size_t)n + 1, sizeof(double));
  double *r = (double *)calloc((size_t)n + 1, sizeof(double));
  double *z = (double *)calloc((size_t)n + 1, sizeof(double));
  double *pp = (double *)calloc((size_t)n + 1, sizeof(double));
  double *tmp = (double *)calloc((size_t)n + 1, sizeof(double));
  double *t = (double *)calloc((size_t)3 * p->n_pt + 1, sizeof(double));
  double *Minv = (double *)calloc((size_t)36 * p->n_cam + 1, sizeof(double));
  int rc = 0, it = 0;
  /* rhs = -g_c + sum_o Jc^T Jp (Vinv g_p) */
  for (int q = 0; q < p->n_pt; ++q) {
    const double *Vi = w->Vinv + 9 * (size_t)q;
    const double *g = w->gp + 3 * q;
    for (int a = 0; a < 3; ++a)
      t[3 * q + a] = Vi[a * 3 + 0] * g[0] + Vi[a * 3 + 1] * g[1] + Vi[a * 3 + 2] * g[2];
  }
  pass_cameras(w, radius, x /*unused*/, t, b, 0, p->n_pt, 0);
  for (int c = 0; c < p->n_cam; ++c)
    for (int k = 0; k < 6; ++k)
      b[6 * c + k] = (L->cam_slot[c] >= 0) ? -w->gc[6 * c + k] - b[6 * c + k] : 0.0;
  /* note pass_cameras(diag=0) returns -sum Jc^T Jp t, hence the double minus */
  if (schur_jacobi(w, radius, Minv)) {
    rc = -1;
    goto done;
  }
  {
    const double norm_b = sqrt(dotn(b, b, n));
    if (norm_b == 0.0) goto finish; /* "Convergence. |b| = 0." */
    memcpy(r, b, sizeof(double) * (size_t)n); /* x = 0 */
    double rho = 1.0;
    double Q0 = -1.0 * dotn(x, b, n) - dotn(x, r, n);
    for (it = 1;; ++it) {
      for (int c = 0; c < p->n_cam; ++c)
        for (int a = 0; a < 6; ++a) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += Minv[36 * (size_t)c + a * 6 + k] * r[6 * c + k];
          z[6 * c + a] = s;
        }
      const double last_rho = rho;
      rho = dotn(r, z, n);
      if (rho == 0.0 || !isfinite(rho)) {
        rc = -1;
        break;
      }
      if (it == 1) {
        memcpy(pp, z, sizeof(double) * (size_t)n);
      } else {
        const double beta = rho / last_rho;
        if (beta == 0.0 || !isfinite(beta)) {
          rc = -1;
          break;
        }
        for (int i = 0; i < n; ++i) pp[i] = z[i] + beta * pp[i];
      }
      double *q = z;
      schur_apply(w, radius, pp, t, q);
      const double pq = dotn(pp, q, n);
      if (pq <= 0.0 || isinf(pq)) break; /* NO_CONVERGENCE, keep x */
      const double alpha = rho / pq;
      if (isinf(alpha)) {
        rc = -1;
        break;
      }
      for (int i = 0; i < n; ++i) x[i] = x[i] + alpha * pp[i];
      if (o->residual_reset_period > 0 && it % o->residual_reset_period == 0) {
        schur_apply(w, radius, x, t, tmp);
        for (int i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
      } else {
        for (int i = 0; i < n; ++i) r[i] = r[i] - alpha * q[i];
      }
      double xb = 0.0;
      for (int i = 0; i < n; ++i) xb += x[i] * (b[i] + r[i]);
      const double Q1 = -1.0 * xb;
      const double zeta = it * (Q1 - Q0) / Q1;
      if (zeta < o->eta && it >= o->min_pcg_iterations) break;
      Q0 = Q1;
      if (it >= o->max_pcg_iterations) break;
    }
  }
finish:
  *iters_out = it;
  if (!rc) {
    for (int i = 0; i < n; ++i)
      if (!isfinite(x[i])) rc = -1;
    memcpy(w->yc, x, sizeof(double) * (size_t)n);
    for (int c = 0; c < p->n_cam; ++c)
      if (L->cam_slot[c] < 0)
        for (int k = 0; k < 6; ++k) w->yc[6 * c + k] = 0.0;
    for (int k = 0; k < 4; ++k) w->yk[k] = 0.0;
  }
done:
  free(b); free(x); free(r); free(z); free(pp); free(tmp); free(t); free(Minv);
  return rc;
}

/* model_cost_change = -(J y).(r + J y / 2)   trust_region_minimizer.cc */
static double model_cost_change(const ora_ws *w) {
  const ora_problem *p = w->p;
  const ora_layout *L = &w->L;
  const int R = L->R;
  double *part = w->J.cam_cost;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < p->n_cam; ++c) {
    double acc = 0.0;
    const int freec = L->cam_slot[c] >= 0;
    for (int i = L->cam_rowptr[c]; i < L->cam_rowptr[c + 1]; ++i) {
      const int q = p->pt_idx[i];
      for (int row = 0; row < R; ++row) {
        double m = 0.0;
        if (freec) {
          const double *jc = w->J.Jc + ((size_t)i * R + row) * 6;
          for (int k = 0; k < 6; ++k) m += jc[k] * w->yc[6 * c + k];
        }
        const double *e = w->J.Jp + ((size_t)i * R + row) * 3;
        for (int k = 0; k < 3; ++k) m += e[k] * w->yp[3 * q + k];
        if (L->nk && row < 2) {
          const double *jk = w->J.Jk + ((size_t)i * 2 + row) * 4;
          for (int k = 0; k < 4; ++k) m += jk[k] * w->yk[k];
        }
        acc += m * (w->J.r[(size_t)i * R + row] + m / 2.0);
      }
    }
    part[c] = acc;
  }
  double total = 0.0;
  if (L->nk)
    for (int k = 0; k < 4; ++k) {
      const double m = w->J.Jkk[k] * w->yk[k];
      total += m * (w->J.rk[k] + m / 2.0);
    }
  for (int c = 0; c < p->n_cam; ++c) total += part[c];
  return -total;
}

static int ws_init(ora_ws *w, const ora_problem *p, const ora_options *o) {
  memset(w, 0, sizeof(*w));
  w->p = p;
  w->o = o;
  int rc = build_layout(p, o, &w->L);
  if (rc) return rc;
  alloc_jac(&w->J, p, &w->L);
  const size_t nc = (size_t)p->n_cam + 1, np = (size_t)p->n_pt + 1;
  w->sc = (double *)calloc(6 * nc, sizeof(double));
  w->dc = (double *)calloc(6 * nc, sizeof(double));
  w->gc = (double *)calloc(6 * nc, sizeof(double));
  w->yc = (double *)calloc(6 * nc, sizeof(double));
  w->sp = (double *)calloc(3 * np, sizeof(double));
  w->dp = (double *)calloc(3 * np, sizeof(double));
  w->gp = (double *)calloc(3 * np, sizeof(double));
  w->yp = (double *)calloc(3 * np, sizeof(double));
  w->Vinv = (double *)calloc(9 * np, sizeof(double));
  return 0;
}
static void ws_free(ora_ws *w) {
  free_layout(&w->L);
  if (w->J.r) free_jac(&w->J);
  free(w->sc); free(w->dc); free(w->gc); free(w->yc);
  free(w->sp); free(w->dp); free(w->gp); free(w->yp); free(w->Vinv);
}

int ora_schur_matvec(const ora_problem *p, const ora_options *o, double radius,
                     const double *x, double *y, int32_t q0, int32_t q1,
                     int32_t include_diag) {
#ifdef _OPENMP
  omp_set_num_threads(o->num_threads > 0 ? o->num_threads : 1);
#endif
  ora_ws w;
  int rc = ws_init(&w, p, o);
  if (rc) {
    ws_free(&w);
    return rc;
  }
  double cost, gmax;
  rc = evaluate_gradient_and_jacobian(&w, p->pose7, p->pt3, p->intr, &cost, &gmax);
  if (!rc) {
    lm_diagonal(&w);
    rc = point_blocks(&w, radius);
  }
  if (!rc) {
    const int n = 6 * p->n_cam;
    double *xs = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *t = (double *)calloc((size_t)3 * p->n_pt + 1, sizeof(double));
    for (int i = 0; i < n; ++i) xs[i] = x[i] / w.sc[i];
    if (q1 <= q0) {
      q0 = 0;
      q1 = p->n_pt;
    }
    pass_points(&w, xs, t, q0, q1);
    pass_cameras(&w, radius, xs, t, y, q0, q1, include_diag);
    for (int i = 0; i < n; ++i) y[i] = y[i] / w.sc[i];
    if (include_diag && p->fixed_cam >= 0) {
      /* unscaled-space convention of the C-ABI: the fixed camera carries only
       * its damping term min_lm_diagonal/radius (its Jacobian columns are 0) */
      for (int k = 0; k < 6; ++k)
        y[6 * p->fixed_cam + k] = (o->min_lm_diagonal / radius) * x[6 * p->fixed_cam + k];
    }
    free(xs);
    free(t);
  }
  ws_free(&w);
  return rc;
}

/* ======================================================================
 * ceres::Solve -- TrustRegionMinimizer + LevenbergMarquardtStrategy
 * ====================================================================== */
int ora_solve(ora_problem *p, const ora_options *o, ora_summary *sum,
              ora_iter *trace, int32_t trace_cap) {
#ifdef _OPENMP
  omp_set_num_threads(o->num_threads > 0 ? o->num_threads : 1);
#endif
  const double t_start = now_s();
  ora_summary S;
  memset(&S, 0, sizeof(S));
  ora_ws w;
  int rc = ws_init(&w, p, o);
  if (rc) {
    ws_free(&w);
    if (sum) *sum = S;
    return rc;
  }
  const ora_layout *L = &w.L;
  const size_t npose = 7 * (size_t)p->n_cam, npt = 3 * (size_t)p->n_pt;
  double *x_pose = p->pose7, *x_pt = p->pt3, *x_k = p->intr;
  double *c_pose = (double *)malloc(sizeof(double) * (npose + 1));
  double *c_pt = (double *)malloc(sizeof(double) * (npt + 1));
  double c_k[4];
  double *cam_cost = (double *)malloc(sizeof(double) * ((size_t)p->n_cam + 1));

  double x_cost = 0.0, gmax = 0.0;
  double radius = o->initial_radius, decrease_factor = 2.0;
  int reuse_diagonal = 0, invalid_run = 0;
  int ntrace = 0;
  ora_iter it;
  memset(&it, 0, sizeof(it));

  /* IterationZero */
  double t0 = now_s();
  if (evaluate_gradient_and_jacobian(&w, x_pose, x_pt, x_k, &x_cost, &gmax)) {
    S.termination = ORA_TERM_FAILURE;
    rc = -1;
    goto out;
  }
  S.seconds_linearize += now_s() - t0;
  S.initial_cost = x_cost;
  it.iteration = 0;
  it.cost = x_cost;
  it.gradient_max_norm = gmax;
  it.radius = radius;
  it.step_is_valid = 1; /* IterationZero marks itself valid+successful */
  it.step_is_successful = 1;
  if (trace && ntrace < trace_cap) trace[ntrace++] = it;
  S.termination = ORA_TERM_NO_CONVERGENCE;
  if (gmax <= o->gradient_tolerance) {
    S.termination = ORA_TERM_GRADIENT;
    goto out;
  }
  double x_norm;
  {
    double s = 0.0;
    for (int c = 0; c < p->n_cam; ++c)
      if (L->cam_slot[c] >= 0)
        for (int k = 0; k < 7; ++k) s += x_pose[7 * c + k] * x_pose[7 * c + k];
    for (size_t i = 0; i < npt; ++i) s += x_pt[i] * x_pt[i];
    for (int k = 0; k < L->nk; ++k) s += x_k[k] * x_k[k];
    x_norm = sqrt(s);
  }

  for (int iter = 1;; ++iter) {
    /* FinalizeIterationAndCheckIfMinimizerCanContinue of the previous iter */
    if (iter - 1 >= o->max_num_iterations) {
      S.termination = ORA_TERM_NO_CONVERGENCE;
      break;
    }
    if (it.step_is_successful && it.gradient_max_norm <= o->gradient_tolerance) {
      S.termination = ORA_TERM_GRADIENT;
      break;
    }
    if (radius <= o->min_radius) {
      S.termination = ORA_TERM_MIN_RADIUS;
      break;
    }
    memset(&it, 0, sizeof(it));
    it.iteration = iter;
    S.num_iterations = iter;

    /* ComputeTrustRegionStep */
    t0 = now_s();
    if (!reuse_diagonal) lm_diagonal(&w);
    int lin_rc, lin_iters = 0;
    if (o->solver == ORA_SOLVER_DENSE_SCHUR) {
      lin_rc = solve_dense_schur(&w, radius);
    } else if (o->solver == ORA_SOLVER_SPARSE_SCHUR) {
      lin_rc = solve_sparse_schur(&w, radius);
    } else {
      lin_rc = solve_implicit_pcg(&w, radius, &lin_iters);
    }
    if (lin_rc == -4) {
      rc = -4;
      S.termination = ORA_TERM_FAILURE;
      break;
    }
    reuse_diagonal = 1;
    it.linear_iters = lin_iters;
    S.total_linear_iters += lin_iters;
    double mcc = 0.0;
    int valid = 0;
    if (lin_rc == 0) {
      back_substitute(&w);
      mcc = model_cost_change(&w);
      valid = mcc > 0.0;
    }
    S.seconds_linear_solve += now_s() - t0;
    it.model_cost_change = mcc;
    it.step_is_valid = valid;
    if (!valid) {
      /* HandleInvalidStep */
      ++invalid_run;
      it.cost = x_cost;
      it.gradient_max_norm = gmax;
      if (invalid_run >= o->max_consecutive_invalid_steps) {
        S.termination = ORA_TERM_FAILURE;
        it.radius = radius;
        if (trace && ntrace < trace_cap) trace[ntrace++] = it;
        break;
      }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = 1;
      it.radius = radius;
      ++S.num_unsuccessful;
      if (trace && ntrace < trace_cap) trace[ntrace++] = it;
      continue;
    }
    invalid_run = 0;

    /* ComputeCandidatePointAndEvaluateCost: delta = y .* scale; x+ = Plus(x, delta) */
    memcpy(c_pose, x_pose, sizeof(double) * npose);
    for (int c = 0; c < p->n_cam; ++c) {
      if (L->cam_slot[c] < 0) continue;
      double d[6], e[7];
      for (int k = 0; k < 6; ++k) d[k] = w.yc[6 * c + k] * w.sc[6 * c + k];
      ora_se3_exp(d, e);
      ora_se3_mul(x_pose + 7 * c, e, c_pose + 7 * c);
    }
    for (size_t i = 0; i < npt; ++i) c_pt[i] = x_pt[i] + w.yp[i] * w.sp[i];
    for (int k = 0; k < 4; ++k) c_k[k] = x_k[k] + (L->nk ? w.yk[k] * w.sk[k] : 0.0);
    double cand_cost;
    t0 = now_s();
    if (evaluate_cost(p, o, L, c_pose, c_pt, c_k, cam_cost, &cand_cost))
      cand_cost = DBL_MAX;
    S.seconds_linearize += now_s() - t0;

    /* ParameterToleranceReached */
    {
      double s = 0.0;
      for (int c = 0; c < p->n_cam; ++c)
        if (L->cam_slot[c] >= 0)
          for (int k = 0; k < 7; ++k) {
            const double d = x_pose[7 * c + k] - c_pose[7 * c + k];
            s += d * d;
          }
      for (size_t i = 0; i < npt; ++i) {
        const double d = x_pt[i] - c_pt[i];
        s += d * d;
      }
      for (int k = 0; k < L->nk; ++k) {
        const double d = x_k[k] - c_k[k];
        s += d * d;
      }
      it.step_norm = sqrt(s);
    }
    it.radius = radius;
    it.cost = x_cost;
    it.gradient_max_norm = gmax;
    if (it.step_norm <= o->parameter_tolerance * (x_norm + o->parameter_tolerance)) {
      S.termination = ORA_TERM_PARAMETER;
      if (trace && ntrace < trace_cap) trace[ntrace++] = it;
      break;
    }
    /* FunctionToleranceReached */
    it.cost_change = x_cost - cand_cost;
    if (fabs(it.cost_change) <= o->function_tolerance * x_cost) {
      S.termination = ORA_TERM_FUNCTION;
      if (trace && ntrace < trace_cap) trace[ntrace++] = it;
      break;
    }
    /* IsStepSuccessful */
    it.relative_decrease = (x_cost - cand_cost) / mcc;
    if (it.relative_decrease > o->min_relative_decrease) {
      /* HandleSuccessfulStep */
      memcpy(x_pose, c_pose, sizeof(double) * npose);
      memcpy(x_pt, c_pt, sizeof(double) * npt);
      for (int k = 0; k < L->nk; ++k) x_k[k] = c_k[k];
      {
        double s = 0.0;
        for (int c = 0; c < p->n_cam; ++c)
          if (L->cam_slot[c] >= 0)
            for (int k = 0; k < 7; ++k) s += x_pose[7 * c + k] * x_pose[7 * c + k];
        for (size_t i = 0; i < npt; ++i) s += x_pt[i] * x_pt[i];
        for (int k = 0; k < L->nk; ++k) s += x_k[k] * x_k[k];
        x_norm = sqrt(s);
      }
      t0 = now_s();
      if (evaluate_gradient_and_jacobian(&w, x_pose, x_pt, x_k, &x_cost, &gmax)) {
        S.termination = ORA_TERM_FAILURE;
        rc = -1;
        break;
      }
      S.seconds_linearize += now_s() - t0;
      it.step_is_successful = 1;
      it.cost = x_cost;
      it.gradient_max_norm = gmax;
      /* LevenbergMarquardtStrategy::StepAccepted */
      radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = fmin(o->max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = 0;
      ++S.num_successful;
    } else {
      /* HandleUnsuccessfulStep -> StepRejected */
      it.step_is_successful = 0;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = 1;
      ++S.num_unsuccessful;
    }
    it.radius = radius;
    if (trace && ntrace < trace_cap) trace[ntrace++] = it;
  }
out:
  S.final_cost = x_cost;
  S.seconds_total = now_s() - t_start;
  if (sum) *sum = S;
  free(c_pose);
  free(c_pt);
  free(cam_cost);
  ws_free(&w);
  return rc;
}
