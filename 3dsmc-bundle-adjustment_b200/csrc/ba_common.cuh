// ba_common.cuh -- shared device-side types and helpers of the sm_100a BA solver.
//
// Data layout in HBM:
//   * observations are struct-of-arrays in the reference's canonical order
//     (camera-major, src/OptimizationUtils.cpp:244,257) and, as a second copy,
//     in point-major order (stable counting sort of pt_idx);
//   * the Jacobian store is "plane" SoA: one double2 array per column k holding
//     the (row0,row1) pair of the two reprojection rows, so every warp access is
//     a run of 32 consecutive 16-byte words (LDG.128 / STG.128, 512 B per
//     request).  The optional depth-prior row lives in plain double planes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define BA_THREADS 256
#define BA_WARPS (BA_THREADS / 32)
#define BA_FULL 0xffffffffu

// one camera-major work item: a run of <= BA_ITEM_OBS observations of ONE camera
#define BA_ITEM_OBS 256
// point-major tiles: BA_TILE_PTS consecutive points per CTA, observations are
// staged through shared memory in chunks of BA_TILE_OBS
#define BA_TILE_PTS 128
#define BA_TILE_OBS 512

struct BaItem {
  int32_t cam, begin, end, pad;
};

// Jacobian planes of one ordering (camera-major or point-major)
struct JPlanes {
  double2 *r;      // (r0, r1) robustified reprojection residual
  double2 *Jc[6];  // column k of the 2x6 local pose Jacobian (row0,row1)
  double2 *Jp[3];  // column k of the 2x3 point Jacobian
  double *r3;      // depth-prior row (REF mode only)
  double *Jc3[6];
  double *Jp3[3];
  double2 *Jk[2];  // intrinsics: Jk[0] = (d r0/d fx, d r1/d fy), Jk[1] = (d r0/d cx, d r1/d cy)
};

// problem-constant scalars, passed by value (kernel parameters live in the
// constant bank, so these are "intrinsics in constant memory" for NS mode)
struct CostParams {
  double sw_repr, sw_unpr;   // sqrt of the residual weights (:280, :290)
  double hub_repr, hub_unpr; // Huber deltas
  double sw_intr;            // sqrt(WEIGHT_INTRINSICS)
  int32_t fixed_cam;
  int32_t pad;
};

struct BaIterRec {  // == ba_gpu_iter (include/ba_gpu.h)
  int32_t iteration, step_is_valid, step_is_successful, linear_iters;
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius,
      model_cost_change;
};

// Levenberg-Marquardt controller state, device resident (Ceres 2.0.0
// trust_region_minimizer.cc / levenberg_marquardt_strategy.cc semantics)
struct LmState {
  double radius, decrease_factor;
  double x_cost, gmax, x_norm;
  double initial_cost;
  int32_t iter;          // iteration being executed (1-based), 0 = iteration zero
  int32_t invalid_run;
  int32_t done;          // LM finished
  int32_t termination;
  int32_t accepted;      // last step accepted -> relinearise
  int32_t last_successful;
  int32_t lin_fail;      // linear solver / point inverse failure in this iteration
  int32_t eval_fail;     // non-finite linearisation
  int32_t num_successful, num_unsuccessful;
  int32_t n_trace;
  int32_t have_scale;
  int64_t total_lin_iters;
  // PCG controller
  int32_t pcg_it, pcg_done, pcg_fail, pcg_break, pcg_iters_last, pcg_counter;
  double pcg_rho, pcg_beta, pcg_Q0;
  BaIterRec pending;     // record of an accepted step, completed after relinearisation
};

struct LmOptions {
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double max_radius, min_radius, min_relative_decrease;
  double min_lm_diagonal, max_lm_diagonal, eta;
  int32_t max_num_iterations, max_invalid, max_pcg, min_pcg, reset_period;
  int32_t trace_cap;
};


// ---------------------------------------------------------------- loads
__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ double ldg1(const double *p) { return __ldg(p); }
// streaming 128-bit load: Jacobian planes are read once per pass
__device__ __forceinline__ double2 lds2(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double lds1(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts2(double2 *p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y));
}
__device__ __forceinline__ void sts1(double *p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v));
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BA_FULL, v, o);
  return v;
}
// Warp sum of N per-lane values by recursive halving: at every level a lane keeps one half of its values and receives the
// partner's partial sums of that half, so 32 lanes end up with the N COMPLETE sums spread over them -- lane l holds entries
// idx .. idx + cnt - 1 in acc[0 .. cnt-1], cnt <= ceil(N / 32).  About N shuffled doubles instead of the 5 N of N butterflies
// (k_sp_schur's 36 butterflies were a quarter of its instructions).  Fixed tree: deterministic.
template <int N, int MASK>
__device__ __forceinline__ void warp_halve(double *acc, int lane, int &idx, int &cnt) {
  constexpr int H = (N + 1) / 2;
  const bool up = (lane & MASK) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const double lo = acc[i], hi = (i + H < N) ? acc[i + H] : 0.0;
    const double keep = up ? hi : lo, send = up ? lo : hi;
    acc[i] = keep + __shfl_xor_sync(BA_FULL, send, MASK);
  }
  if (up) {
    idx += H;
    cnt = cnt > H ? cnt - H : 0;
  } else {
    cnt = cnt < H ? cnt : H;
  }
  if constexpr (MASK > 1) warp_halve<H, MASK / 2>(acc, lane, idx, cnt);
}
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(double (&acc)[N], int lane, int &idx, int &cnt) {
  idx = 0;
  cnt = N;
  warp_halve<N, 16>(acc, lane, idx, cnt);
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(BA_FULL, v, o));
  return v;
}
// deterministic CTA sum; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem /*>=BA_WARPS*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += smem[i];
  }
  return s;
}
__device__ __forceinline__ double block_max(double v, double *smem) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s = fmax(s, smem[i]);
  }
  return s;
}
// every thread of the CTA gets sum(part[0..n)) -- fixed order => identical on
// every CTA that calls it (used to "all-reduce" per-CTA partials without atomics)
// (eight interleaved accumulators per thread: eight loads in flight instead of one dependent load per element -- the
// single-CTA controller kernels add up to 31 k per-CTA partials at config 5, 131 us with the sequential loop; the order
// is still fixed, and unchanged for n <= 8 blockDim.x ... only the grouping inside a thread's strided list differs)
__device__ __forceinline__ double block_sum_array(const double *part, int n, double *smem /*>=BA_WARPS+1*/) {
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int bd = blockDim.x;
  int i = threadIdx.x;
  for (; i + 7 * bd < n; i += 8 * bd) {
    double t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = part[i + k * bd];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += t[k];
  }
  for (int k = 0; i < n; i += bd, ++k) a[k & 7] += part[i];
  double v = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += smem[i];
    smem[BA_WARPS] = s;
  }
  __syncthreads();
  return smem[BA_WARPS];
}

// ---------------------------------------------------------------- SE3 (Sophus semantics)
// Eigen 3.4.0 QuaternionBase::toRotationMatrix (non-normalising); row-major R.
// Called by the cost functors at src/OptimizationUtils.cpp:41, 85.
__device__ __forceinline__ void quat_to_R(const double q[4], double R[9]) {
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
// Hamilton product, storage (x,y,z,w)
__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
// Eigen _transformVector (headers/sophus/so3.hpp:322-324)
__device__ __forceinline__ void quat_rotate(const double q[4], const double v[3], double o[3]) {
  double uv0 = q[1] * v[2] - q[2] * v[1], uv1 = q[2] * v[0] - q[0] * v[2], uv2 = q[0] * v[1] - q[1] * v[0];
  uv0 += uv0;
  uv1 += uv1;
  uv2 += uv2;
  const double c0 = q[1] * uv2 - q[2] * uv1, c1 = q[2] * uv0 - q[0] * uv2, c2 = q[0] * uv1 - q[1] * uv0;
  o[0] = v[0] + q[3] * uv0 + c0;
  o[1] = v[1] + q[3] * uv1 + c1;
  o[2] = v[2] + q[3] * uv2 + c2;
}
// SE3::exp (headers/sophus/se3.hpp:725-746) with SO3::expAndTheta
// (so3.hpp:537-571); epsilon 1e-10 (common.hpp:144)
__device__ __forceinline__ void se3_exp(const double d[6], double out[7]) {
  const double *ups = d, *om = d + 3;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double th = sqrt(th2);
  const double half = 0.5 * th;
  double imag, real;
  const bool small = th < 1e-10;
  if (small) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    imag = sin(half) / th;
    real = cos(half);
  }
  double q[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  double V[9];
  if (small) {
    quat_to_R(q, V);
  } else {
    const double Om[9] = {0.0, -om[2], om[1], om[2], 0.0, -om[0], -om[1], om[0], 0.0};
    const double c1 = (1.0 - cos(th)) / th2;
    const double c2 = (th - sin(th)) / (th2 * th);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double o2 = (Om[i * 3 + 0] * Om[0 * 3 + j] + Om[i * 3 + 1] * Om[1 * 3 + j]) + Om[i * 3 + 2] * Om[2 * 3 + j];
        V[i * 3 + j] = ((i == j ? 1.0 : 0.0) + c1 * Om[i * 3 + j]) + c2 * o2;
      }
  }
  out[0] = q[0];
  out[1] = q[1];
  out[2] = q[2];
  out[3] = q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) out[4 + i] = (V[i * 3 + 0] * ups[0] + V[i * 3 + 1] * ups[1]) + V[i * 3 + 2] * ups[2];
}
// SE3 product (se3.hpp:317-321: t += R*t2; then SO3::operator*= so3.hpp:339-356
// with the first-order renormalisation 2/(1+|q|^2))
__device__ __forceinline__ void se3_mul(const double a[7], const double b[7], double out[7]) {
  double rt[3], o[4];
  quat_rotate(a, b + 4, rt);
  quat_mul(a, b, o);
  const double n2 = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3];
  if (n2 != 1.0) {
    const double s = 2.0 / (1.0 + n2);
    o[0] *= s;
    o[1] *= s;
    o[2] *= s;
    o[3] *= s;
  }
  out[0] = o[0];
  out[1] = o[1];
  out[2] = o[2];
  out[3] = o[3];
  out[4] = a[4] + rt[0];
  out[5] = a[5] + rt[1];
  out[6] = a[6] + rt[2];
}
// LocalParameterizationSE3::Plus (headers/sophus/local_parameterization_se3.hpp:17-24)
__device__ __forceinline__ void se3_plus(const double T[7], const double d[6], double out[7]) {
  double e[7];
  se3_exp(d, e);
  se3_mul(T, e, out);
}

// ceres::HuberLoss + Corrector (rho2 <= 0 branch): returns sqrt(rho1) scaling
// and adds rho0 to *rho0.  a = delta, s = squared norm of the residual block.
__device__ __forceinline__ double huber_scale(double a, double s, double &rho0) {
  const double b = a * a;
  if (s > b) {
    const double rs = sqrt(s);
    rho0 = 2.0 * a * rs - b;
    const double rho1 = fmax(DBL_MIN, a / rs);
    return sqrt(rho1);
  }
  rho0 = s;
  return 1.0;
}

// 3x3 SPD inverse through Cholesky (ceres InvertPSDMatrix == llt().solve(I)),
// packed symmetric storage (00,01,02,11,12,22).  Returns false if not SPD.
__device__ __forceinline__ bool spd3_inverse(const double V[6], double Vi[6]) {
  const double a00 = V[0], a10 = V[1], a20 = V[2], a11 = V[3], a21 = V[4], a22 = V[5];
  if (!(a00 > 0.0)) return false;
  const double l00 = sqrt(a00);
  const double i00 = 1.0 / l00;
  const double l10 = a10 * i00, l20 = a20 * i00;
  const double d1 = a11 - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = sqrt(d1);
  const double i11 = 1.0 / l11;
  const double l21 = (a21 - l20 * l10) * i11;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = sqrt(d2);
  const double i22 = 1.0 / l22;
  // M = L^-1 (lower)
  const double m00 = i00, m11 = i11, m22 = i22;
  const double m10 = -l10 * m00 * i11;
  const double m21 = -l21 * m11 * i22;
  const double m20 = -(l20 * m00 + l21 * m10) * i22;
  // Vinv = M^T M
  Vi[0] = m00 * m00 + m10 * m10 + m20 * m20;
  Vi[1] = m10 * m11 + m20 * m21;
  Vi[2] = m20 * m22;
  Vi[3] = m11 * m11 + m21 * m21;
  Vi[4] = m21 * m22;
  Vi[5] = m22 * m22;
  return isfinite(Vi[0]) && isfinite(Vi[3]) && isfinite(Vi[5]);
}
__device__ __forceinline__ void sym3_mul(const double Vi[6], const double b[3], double o[3]) {
  o[0] = Vi[0] * b[0] + Vi[1] * b[1] + Vi[2] * b[2];
  o[1] = Vi[1] * b[0] + Vi[3] * b[1] + Vi[4] * b[2];
  o[2] = Vi[2] * b[0] + Vi[4] * b[1] + Vi[5] * b[2];
}
