// ba_kernels_chol.cuh -- blocked dense Cholesky for the explicit Schur complement of LARGE REF-mode
// problems (free intrinsics make the reduced system a dense-bordered matrix, so the reference's
// "global optimisation" call -- windowOptimize(cfg, 0, keyframes.size() - 1, ...), src/main.cpp:179-182
// -- needs an exact dense solve beyond the single-CTA kernel's 1024 unknowns).
//
// Right-looking, 64 x 64 fp64 tiles, lower triangle of the row-major n x n matrix, in place:
//   for k = 0 .. nt-1:   k_chol_potrf   diagonal tile (one CTA, shared memory, column Cholesky)
//                        k_chol_trsm    tiles (i, k), i > k:  X L_kk^T = A_ik   (one thread per row)
//                        k_chol_update  tiles (i, j), k < j <= i:  A_ij -= L_ik L_jk^T
//                                       (256 threads, 4 x 4 outputs each, both operand tiles in shared memory)
// then k_chol_solve: blocked forward / backward substitution and the scatter into y_c / y_k.
// Every sum has a fixed order (no atomics).  == the exact step of Ceres DENSE_SCHUR / SPARSE_SCHUR.
#pragma once
#include "ba_kernels.cuh"

#define CH_NB 64
#define CH_LD (CH_NB + 1)

__global__ void __launch_bounds__(CH_NB)
k_chol_potrf(int n, double *__restrict__ A, int k, LmState *st, int gate) {
  if (!gate_open(st, gate) || st->lin_fail) return;
  __shared__ double T[CH_NB * CH_LD];
  __shared__ double piv;
  __shared__ int fail;
  const int j0 = k * CH_NB, m = min(CH_NB, n - j0), i = threadIdx.x;
  if (i == 0) fail = 0;
  for (int c = 0; c < m; ++c)
    if (i < m) T[i * CH_LD + c] = c <= i ? A[(size_t)(j0 + i) * n + j0 + c] : 0.0;
  __syncthreads();
  for (int j = 0; j < m; ++j) {
    double s = 0.0;
    if (i >= j && i < m) {
      s = T[i * CH_LD + j];
      for (int q = 0; q < j; ++q) s -= T[i * CH_LD + q] * T[j * CH_LD + q];
      if (i == j) {
        if (!(s > 0.0) || !isfinite(s)) fail = 1;
        piv = 1.0 / sqrt(s);
        T[j * CH_LD + j] = sqrt(s);
      }
    }
    __syncthreads();
    if (fail) break;
    if (i > j && i < m) T[i * CH_LD + j] = s * piv;
    __syncthreads();
  }
  if (fail) {
    if (i == 0) st->lin_fail = 1;
    return;
  }
  for (int c = 0; c < m; ++c)
    if (i < m && c <= i) A[(size_t)(j0 + i) * n + j0 + c] = T[i * CH_LD + c];
}

// rows below the diagonal tile: one thread per row, L_kk broadcast from shared memory
__global__ void __launch_bounds__(CH_NB)
k_chol_trsm(int n, double *__restrict__ A, int k, const LmState *st, int gate) {
  if (!gate_open(st, gate) || st->lin_fail) return;
  extern __shared__ double sh[];
  double *L = sh, *X = sh + CH_NB * CH_LD;
  const int j0 = k * CH_NB, m = min(CH_NB, n - j0);
  const int r0 = (k + 1 + blockIdx.x) * CH_NB, t = threadIdx.x, row = r0 + t;
  for (int c = 0; c < m; ++c)
    if (t < m) L[t * CH_LD + c] = c <= t ? A[(size_t)(j0 + t) * n + j0 + c] : 0.0;
  // the row's 64 entries: coalesced across the CTA column by column would be strided; read row-wise per thread
  if (row < n)
    for (int c = 0; c < m; ++c) X[t * CH_LD + c] = A[(size_t)row * n + j0 + c];
  __syncthreads();
  if (row >= n) return;
  for (int j = 0; j < m; ++j) {
    double s = X[t * CH_LD + j];
    for (int q = 0; q < j; ++q) s -= X[t * CH_LD + q] * L[j * CH_LD + q];
    X[t * CH_LD + j] = s / L[j * CH_LD + j];
  }
  for (int c = 0; c < m; ++c) A[(size_t)row * n + j0 + c] = X[t * CH_LD + c];
}

// trailing update of the lower triangle: linear CTA index -> tile (i, j), j <= i, both > k
__global__ void __launch_bounds__(256)
k_chol_update(int n, double *__restrict__ A, int k, const LmState *st, int gate) {
  if (!gate_open(st, gate) || st->lin_fail) return;
  extern __shared__ double sh[];
  double *As = sh, *Bs = sh + CH_NB * CH_LD;
  // triangular decode: b = ti (ti + 1) / 2 + tj
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while (ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int i0 = (k + 1 + ti) * CH_NB, c0 = (k + 1 + tj) * CH_NB, j0 = k * CH_NB;
  const int kk = min(CH_NB, n - j0);
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    As[r * CH_LD + c] = (i0 + r < n && c < kk) ? A[(size_t)(i0 + r) * n + j0 + c] : 0.0;
    Bs[r * CH_LD + c] = (c0 + r < n && c < kk) ? A[(size_t)(c0 + r) * n + j0 + c] : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int q = 0; q < CH_NB; ++q) {
    double a[4], bb[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = As[(ty * 4 + r) * CH_LD + q];
#pragma unroll
    for (int c = 0; c < 4; ++c) bb[c] = Bs[(tx * 4 + c) * CH_LD + q];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * bb[c];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = i0 + ty * 4 + r;
    if (row >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = c0 + tx * 4 + c;
      if (col < n && col <= row) A[(size_t)row * n + col] -= acc[r][c];
    }
  }
}

// L z = b, L^T y = z (blocked, single CTA), then the scatter of k_cholesky_solve
__global__ void __launch_bounds__(1024)
k_chol_solve(int n, const double *__restrict__ A, const double *__restrict__ rhs, double *__restrict__ v /* scratch [n] */,
             int n_cam, int n_free, const int32_t *__restrict__ cam_slot, int nk, double *__restrict__ yc,
             double *__restrict__ yk, LmState *st, int gate) {
  if (!gate_open(st, gate) || st->lin_fail) return;
  __shared__ double zt[CH_NB];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nt = (n + CH_NB - 1) / CH_NB;
  for (int i = tid; i < n; i += nthr) v[i] = rhs[i];
  __syncthreads();
  // forward
  for (int k = 0; k < nt; ++k) {
    const int j0 = k * CH_NB, m = min(CH_NB, n - j0);
    if (tid < 32) {  // one warp solves the diagonal tile: lane owns rows lane, lane + 32
      for (int j = 0; j < m; ++j) {
        if ((j & 31) == tid) {
          double s = v[j0 + j];
          for (int q = 0; q < j; ++q) s -= A[(size_t)(j0 + j) * n + j0 + q] * zt[q];
          zt[j] = s / A[(size_t)(j0 + j) * n + j0 + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < m) v[j0 + tid] = zt[tid];
    for (int i = j0 + m + tid; i < n; i += nthr) {
      const double *Li = A + (size_t)i * n + j0;
      double s = v[i];
      for (int q = 0; q < m; ++q) s -= Li[q] * zt[q];
      v[i] = s;
    }
    __syncthreads();
  }
  // backward
  for (int k = nt - 1; k >= 0; --k) {
    const int j0 = k * CH_NB, m = min(CH_NB, n - j0);
    if (tid < 32) {
      for (int j = m - 1; j >= 0; --j) {
        if ((j & 31) == tid) {
          double s = v[j0 + j];
          for (int q = j + 1; q < m; ++q) s -= A[(size_t)(j0 + q) * n + j0 + j] * zt[q];
          zt[j] = s / A[(size_t)(j0 + j) * n + j0 + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < m) v[j0 + tid] = zt[tid];
    for (int i = tid; i < j0; i += nthr) {
      double s = v[i];
      for (int q = 0; q < m; ++q) s -= A[(size_t)(j0 + q) * n + i] * zt[q];
      v[i] = s;
    }
    __syncthreads();
  }
  for (int c = tid; c < n_cam; c += nthr) {
    const int slot = cam_slot[c];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double x = slot >= 0 ? v[6 * slot + q] : 0.0;
      if (!isfinite(x)) st->lin_fail = 1;
      yc[6 * (size_t)c + q] = x;
    }
  }
  if (tid < 4) yk[tid] = nk ? v[6 * n_free + tid] : 0.0;
}
