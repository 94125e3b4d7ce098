// ba_kernels_store.cuh -- device-resident landmark / keyframe store for sliding windows.
//
// The reference keeps the tracking state in host containers -- std::vector<KeyFrame> (pose, key points, local depths,
// global_points_map) and Map3D = unordered_map<LandmarkId, Landmark> (headers/CommonTypes.h:15-43, filled by
// src/Map3D.cpp:7-74) -- and windowOptimize re-walks ALL of them for every window (src/OptimizationUtils.cpp:244-294).
// With the store, a keyframe's observation list (in its container order), the world poses and the world points stay in
// HBM; a window then costs the host only the keyframes that are new or grew since the last call, and the canonical
// enumeration of the window -- admissible observations in container order, point indices by first appearance
// (:257-276), the change into the frame of the window's first keyframe (:231-232, :248, :274) and the way back
// (:303-310) -- runs here.
//
// Bit-exactness: the full-upload path does the frame changes on the host with the Sophus formulas
// (host/compat/reference_types.h; x86-64 without FMA contraction).  These kernels use the same operation order with
// explicit round-to-nearest intrinsics (no contraction), so both paths hand the solver identical bits.
#pragma once
#include "ba_common.cuh"

// ---- SE3 with the host's operation order, no FMA contraction
__device__ __forceinline__ void rn_rotate(const double q[4], const double v[3], double o[3]) {
  double uv[3] = {__dsub_rn(__dmul_rn(q[1], v[2]), __dmul_rn(q[2], v[1])), __dsub_rn(__dmul_rn(q[2], v[0]), __dmul_rn(q[0], v[2])),
                  __dsub_rn(__dmul_rn(q[0], v[1]), __dmul_rn(q[1], v[0]))};
#pragma unroll
  for (int i = 0; i < 3; ++i) uv[i] = __dadd_rn(uv[i], uv[i]);
  const double c3[3] = {__dsub_rn(__dmul_rn(q[1], uv[2]), __dmul_rn(q[2], uv[1])), __dsub_rn(__dmul_rn(q[2], uv[0]), __dmul_rn(q[0], uv[2])),
                        __dsub_rn(__dmul_rn(q[0], uv[1]), __dmul_rn(q[1], uv[0]))};
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = __dadd_rn(__dadd_rn(v[i], __dmul_rn(q[3], uv[i])), c3[i]);
}
// SE3::inverse (se3.hpp:186-189): conjugate, re-normalised; t' = R^T (-t)
__device__ __forceinline__ void rn_se3_inverse(const double a[7], double o[7]) {
  double q[4] = {-a[0], -a[1], -a[2], a[3]};
  const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])), __dmul_rn(q[3], q[3])));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = __ddiv_rn(q[i], len);
  const double nt[3] = {__dmul_rn(a[4], -1.0), __dmul_rn(a[5], -1.0), __dmul_rn(a[6], -1.0)};
  rn_rotate(q, nt, o + 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = q[i];
}
// SE3 product (se3.hpp:317-321 + so3.hpp:339-356 with the 2 / (1 + |q|^2) renormalisation)
__device__ __forceinline__ void rn_se3_mul(const double a[7], const double b[7], double o[7]) {
  double rt[3];
  rn_rotate(a, b + 4, rt);
#pragma unroll
  for (int i = 0; i < 3; ++i) o[4 + i] = __dadd_rn(a[4 + i], rt[i]);
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  double q[4] = {
      __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(aw, bx), __dmul_rn(ax, bw)), __dmul_rn(ay, bz)), __dmul_rn(az, by)),
      __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(aw, by), __dmul_rn(ay, bw)), __dmul_rn(az, bx)), __dmul_rn(ax, bz)),
      __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(aw, bz), __dmul_rn(az, bw)), __dmul_rn(ax, by)), __dmul_rn(ay, bx)),
      __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(aw, bw), __dmul_rn(ax, bx)), __dmul_rn(ay, by)), __dmul_rn(az, bz))};
  const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])), __dmul_rn(q[3], q[3]));
  if (n2 != 1.0) {
    const double s = __ddiv_rn(2.0, __dadd_rn(1.0, n2));
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = __dmul_rn(q[i], s);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = q[i];
}
__device__ __forceinline__ void rn_se3_act(const double a[7], const double x[3], double o[3]) {
  double r[3];
  rn_rotate(a, x, r);
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = __dadd_rn(r[i], a[4 + i]);
}

// world points of landmarks into the table (by landmark id)
__global__ void __launch_bounds__(BA_THREADS)
ks_scatter_points(int n, const int32_t *__restrict__ id, const double *__restrict__ xyz, double *__restrict__ pt_w) {
  const int i = blockIdx.x * BA_THREADS + threadIdx.x;
  if (i >= n) return;
  const size_t d = 3 * (size_t)id[i];
  pt_w[d] = xyz[3 * (size_t)i];
  pt_w[d + 1] = xyz[3 * (size_t)i + 1];
  pt_w[d + 2] = xyz[3 * (size_t)i + 2];
}

// window position t -> (keyframe of the window, slot in the observation pool)
__device__ __forceinline__ void ks_locate(int t, int n_cam, const int32_t *__restrict__ win_off, const long long *__restrict__ win_seg,
                                          int &k, long long &src) {
  int lo = 0, hi = n_cam - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (win_off[mid] <= t)
      lo = mid;
    else
      hi = mid - 1;
  }
  k = lo;
  src = win_seg[lo] + (t - win_off[lo]);
}
// Enumeration of a window in two kernels and one scan.
//   ks_mark: admissible observations (local depth > 1e-15, src/OptimizationUtils.cpp:265-268) and, per landmark, the window
//   position of its first admissible observation: first[lm] = max over (epoch << 32 | ~t) -- an integer atomicMax whose
//   result does not depend on the execution order; entries of earlier windows carry a smaller epoch and lose, so the table
//   is never cleared.
//   ks_isfirst: packed[t] = admissible | is-first-appearance << 32; ONE 64-bit exclusive scan then yields both prefix sums:
//   low word = observation index (canonical order), high word = point index (order of first appearance, :271-276).
__global__ void __launch_bounds__(BA_THREADS)
ks_mark(int total, int n_cam, const int32_t *__restrict__ win_off, const long long *__restrict__ win_seg,
        const double *__restrict__ depth, const int32_t *__restrict__ lm, unsigned long long epoch, unsigned long long *first,
        unsigned long long *__restrict__ packed) {
  const int t = blockIdx.x * BA_THREADS + threadIdx.x;
  if (t > total) return;
  if (t == total) {
    packed[t] = 0;  // (the scan's last output holds the two counts)
    return;
  }
  int k;
  long long src;
  ks_locate(t, n_cam, win_off, win_seg, k, src);
  const bool ok = depth[src] > 1e-15;
  packed[t] = ok ? 1ull : 0ull;
  if (ok) atomicMax(first + lm[src], (epoch << 32) | (unsigned long long)(0xffffffffu - (unsigned)t));
}
__global__ void __launch_bounds__(BA_THREADS)
ks_isfirst(int total, int n_cam, const int32_t *__restrict__ win_off, const long long *__restrict__ win_seg,
           const int32_t *__restrict__ lm, unsigned long long epoch, const unsigned long long *__restrict__ first,
           unsigned long long *__restrict__ packed) {
  const int t = blockIdx.x * BA_THREADS + threadIdx.x;
  if (t >= total || !packed[t]) return;
  int k;
  long long src;
  ks_locate(t, n_cam, win_off, win_seg, k, src);
  if (first[lm[src]] == ((epoch << 32) | (unsigned long long)(0xffffffffu - (unsigned)t))) packed[t] = 1ull | (1ull << 32);
}
// the window's arrays in canonical order; points (first appearances) moved into the frame of the first keyframe
__global__ void __launch_bounds__(BA_THREADS)
ks_emit(int total, int n_cam, const int32_t *__restrict__ win_off, const long long *__restrict__ win_seg,
        const unsigned long long *__restrict__ packed, const unsigned long long *__restrict__ prefix, const int32_t *__restrict__ lm,
        const float2 *__restrict__ uvf, const double *__restrict__ depth, const unsigned long long *__restrict__ first,
        const double *__restrict__ pt_w, const double *__restrict__ T0inv, int32_t *__restrict__ w_cam, int32_t *__restrict__ w_pt,
        double2 *__restrict__ w_uv, double *__restrict__ w_depth, int32_t *__restrict__ lm_of_pt, double *__restrict__ pt3) {
  const int t = blockIdx.x * BA_THREADS + threadIdx.x;
  if (t >= total || !packed[t]) return;
  int k;
  long long src;
  ks_locate(t, n_cam, win_off, win_seg, k, src);
  const int l = lm[src], i = (int)(prefix[t] & 0xffffffffull);
  const int ft = (int)(0xffffffffu - (unsigned)(first[l] & 0xffffffffull));  // window position of the first appearance
  const int p = (int)(prefix[ft] >> 32);
  w_cam[i] = k;
  w_pt[i] = p;
  const float2 f = uvf[src];
  w_uv[i] = make_double2((double)f.x, (double)f.y);  // float -> double (:262)
  w_depth[i] = depth[src];
  if (ft == t) {
    lm_of_pt[p] = l;
    const double x[3] = {pt_w[3 * (size_t)l], pt_w[3 * (size_t)l + 1], pt_w[3 * (size_t)l + 2]};
    double o[3];
    rn_se3_act(T0inv, x, o);  // :274
    pt3[3 * (size_t)p] = o[0];
    pt3[3 * (size_t)p + 1] = o[1];
    pt3[3 * (size_t)p + 2] = o[2];
  }
}
// T0 (copy), T0^-1, and the window's poses in the frame of the first keyframe (:231-232, :248)
__global__ void ks_frame(int n_cam, const double *__restrict__ pose_w /* first keyframe of the window */, double *__restrict__ T0,
                         double *__restrict__ T0inv, double *__restrict__ w_pose) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double t0[7], ti[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) t0[j] = pose_w[j];
  rn_se3_inverse(t0, ti);
  if (k == 0) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      T0[j] = t0[j];
      T0inv[j] = ti[j];
    }
  }
  if (k >= n_cam) return;
  double p[7], o[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) p[j] = pose_w[7 * (size_t)k + j];
  rn_se3_mul(ti, p, o);
#pragma unroll
  for (int j = 0; j < 7; ++j) w_pose[7 * (size_t)k + j] = o[j];
}
// back to the world frame (:303-310): the store's tables are updated in place, the results packed for the host
__global__ void __launch_bounds__(BA_THREADS)
ks_writeback(int n_cam, int n_pt, const double *__restrict__ T0, const double *__restrict__ pose, const double *__restrict__ pt,
             const int32_t *__restrict__ lm_of_pt, double *__restrict__ pose_w, double *__restrict__ pt_w, double *__restrict__ pose_out,
             double *__restrict__ pt_out) {
  const int e = blockIdx.x * BA_THREADS + threadIdx.x;
  double t0[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) t0[j] = T0[j];
  if (e < n_cam) {
    double p[7], o[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) p[j] = pose[7 * (size_t)e + j];
    rn_se3_mul(t0, p, o);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      pose_w[7 * (size_t)e + j] = o[j];
      pose_out[7 * (size_t)e + j] = o[j];
    }
  } else if (e < n_cam + n_pt) {
    const int p = e - n_cam, l = lm_of_pt[p];
    const double x[3] = {pt[3 * (size_t)p], pt[3 * (size_t)p + 1], pt[3 * (size_t)p + 2]};
    double o[3];
    rn_se3_act(t0, x, o);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      pt_w[3 * (size_t)l + j] = o[j];
      pt_out[3 * (size_t)p + j] = o[j];
    }
  }
}
