// ba_kernels_sparse.cuh -- EXPLICIT block-sparse Schur complement + PCG
// (BA_SOLVER_SPARSE_SCHUR_PCG): the reduced camera system of the large NS-mode
// problems formed once per LM iteration as a block-sparse symmetric matrix
//
//     S_ij = delta_ij U_i - sum_{p seen by i and j} W_ip V_p^-1 W_jp^T        (6x6 blocks)
//
// and multiplied by a block-CSR kernel inside PCG (== Ceres ITERATIVE_SCHUR with
// use_explicit_schur_complement, and the structure SPARSE_SCHUR factorises:
// src/OptimizationUtils.cpp:300 with linear_solver_type = SPARSE_SCHUR,
// headers/BundleAdjustmentConfig.h:63).  Sequential-SLAM problems have a very
// sparse S (config 5: 156k upper blocks = 45 MB for 10k cameras, L2-resident),
// so one PCG iteration reads tens of MB instead of streaming every observation.
//
// Structure (built once per upload, integer-only, on the device):
//   pairs   : for every point, all (a <= b) pairs of its free-camera observations
//             (point-major positions), keyed by (cam_a, cam_b); stable radix sort
//             by key => the pair list of every block, in point-major order
//             (the fixed summation order of the block)
//   blocks  : the distinct keys (upper triangle incl. diagonal), blk_ptr into pairs
//   rows    : per camera the entries (block, transposed?) in ascending column order
// Values: one warp per block walks its pair list; Jacobian entries are rebuilt from
// the factored point-major store (ba_kernels_fact.cuh).
#pragma once
#include "ba_kernels_fact.cuh"

// pairs per point: k (k + 1) / 2 over its free-camera observations
__global__ void __launch_bounds__(BA_THREADS)
k_sp_count(int n_pt, const int32_t *__restrict__ pt_rowptr, const int32_t *__restrict__ pm_cam, int fixed_cam,
           long long *__restrict__ cnt) {
  const int p = blockIdx.x * BA_THREADS + threadIdx.x;
  if (p > n_pt) return;
  long long k = 0, extra = 0;
  if (p < n_pt) {
    int prev = -1, run = 0;
    for (int s = pt_rowptr[p]; s < pt_rowptr[p + 1]; ++s) {
      const int c = pm_cam[s];
      if (c == fixed_cam) continue;
      ++k;
      run = (c == prev) ? run + 1 : 0;  // a camera seeing the point twice: both orders of the pair
      extra += run;
      prev = c;
    }
  }
  cnt[p] = k * (k + 1) / 2 + extra;  // cnt[n_pt] = 0: the scan's last entry is the total
}

__global__ void __launch_bounds__(BA_THREADS)
k_sp_emit(int n_pt, int n_cam, const int32_t *__restrict__ pt_rowptr, const int32_t *__restrict__ pm_cam, int fixed_cam,
          const long long *__restrict__ off, unsigned long long *__restrict__ keys, unsigned long long *__restrict__ vals) {
  const int p = blockIdx.x * BA_THREADS + threadIdx.x;
  if (p >= n_pt) return;
  long long w = off[p];
  const int b0 = pt_rowptr[p], e0 = pt_rowptr[p + 1];
  for (int a = b0; a < e0; ++a) {
    const int ca = pm_cam[a];
    if (ca == fixed_cam) continue;
    for (int b = a; b < e0; ++b) {
      const int cb = pm_cam[b];
      if (cb == fixed_cam) continue;
      // cameras ascend inside a point's run (stable sort of a camera-major list): ca <= cb
      keys[w] = (unsigned long long)ca * (unsigned long long)n_cam + (unsigned long long)cb;
      vals[w] = ((unsigned long long)(unsigned)a << 32) | (unsigned)b;
      ++w;
      if (ca == cb && a != b) {  // same camera twice: the diagonal block also needs W_b V^-1 W_a^T
        keys[w] = keys[w - 1];
        vals[w] = ((unsigned long long)(unsigned)b << 32) | (unsigned)a;
        ++w;
      }
    }
  }
}

// decode the distinct keys; per-row counts of upper and transposed entries
__global__ void __launch_bounds__(BA_THREADS)
k_sp_blocks(int n_blk, int n_cam, const unsigned long long *__restrict__ ukeys, int32_t *__restrict__ blk_i,
            int32_t *__restrict__ blk_j, int32_t *row_ucnt, int32_t *row_tcnt, unsigned long long *__restrict__ tkeys,
            int32_t *__restrict__ tvals) {
  const int b = blockIdx.x * BA_THREADS + threadIdx.x;
  if (b >= n_blk) return;
  const unsigned long long k = ukeys[b];
  const int i = (int)(k / (unsigned long long)n_cam), j = (int)(k % (unsigned long long)n_cam);
  blk_i[b] = i;
  blk_j[b] = j;
  atomicAdd(row_ucnt + i, 1);
  if (i != j) atomicAdd(row_tcnt + j, 1);
  // transposed list: sorted by (j, i); diagonal blocks get the largest key and are dropped
  tkeys[b] = i != j ? (unsigned long long)j * (unsigned long long)n_cam + (unsigned long long)i : ~0ull;
  tvals[b] = b;
}

// row r: transposed entries (columns < r, ascending) then upper entries (columns >= r, ascending)
// entry = block index, bit 31 set when the block is used transposed
__global__ void __launch_bounds__(BA_THREADS)
k_sp_entries(int n_cam, const int32_t *__restrict__ row_ustart, const int32_t *__restrict__ row_tstart,
             const int32_t *__restrict__ tvals_sorted, const int32_t *__restrict__ blk_i, const int32_t *__restrict__ blk_j,
             int32_t *__restrict__ ent_ptr, uint32_t *__restrict__ ent_blk, int32_t *__restrict__ ent_col) {
  const int r = blockIdx.x * BA_THREADS + threadIdx.x;
  if (r > n_cam) return;
  const int base = row_ustart[r] + row_tstart[r];
  ent_ptr[r] = base;
  if (r == n_cam) return;
  int w = base;
  for (int e = row_tstart[r]; e < row_tstart[r + 1]; ++e, ++w) {
    const int b = tvals_sorted[e];
    ent_blk[w] = (uint32_t)b | 0x80000000u;
    ent_col[w] = blk_i[b];
  }
  for (int b = row_ustart[r]; b < row_ustart[r + 1]; ++b, ++w) {
    ent_blk[w] = (uint32_t)b;
    ent_col[w] = blk_j[b];
  }
}

// ------------------------------------------------------------------ values
// One warp per block.  S_b = [i == j] U_i - diag(s_i) (sum_pairs W_a Vs_p W_b^T) diag(s_j),
// W_o = jr0_o (x) p0_o + jr1_o (x) p1_o (un-scaled rows rebuilt from the factored store).
// Through the 2x2 core M = P_a Vs P_b^T:  W_a Vs W_b^T = [jr0_a jr1_a] M [jr0_b jr1_b]^T.
__device__ __forceinline__ void sp_rows(const ObsGeo &o, const double R[9], double jr0[6], double jr1[6], double p0[3],
                                        double p1[3]) {
  jr0[0] = -o.wfx * o.iz; jr0[1] = 0.0; jr0[2] = o.wfx * o.iz * o.xz; jr0[3] = o.wfx * o.xz * o.yz;
  jr0[4] = -o.wfx * (1.0 + o.xz * o.xz); jr0[5] = o.wfx * o.yz;
  jr1[0] = 0.0; jr1[1] = -o.wfy * o.iz; jr1[2] = o.wfy * o.iz * o.yz; jr1[3] = o.wfy * (1.0 + o.yz * o.yz);
  jr1[4] = -o.wfy * o.xz * o.yz; jr1[5] = -o.wfy * o.xz;
  const double a = o.wfx * o.iz, b = o.wfy * o.iz;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p0[k] = a * (R[3 * k] - o.xz * R[3 * k + 2]);
    p1[k] = b * (R[3 * k + 1] - o.yz * R[3 * k + 2]);
  }
}

__global__ void __launch_bounds__(BA_THREADS)
k_sp_schur(int n_blk, const int32_t *__restrict__ blk_ptr, const int32_t *__restrict__ blk_i, const int32_t *__restrict__ blk_j,
           const unsigned long long *__restrict__ pairs, const int32_t *__restrict__ pm_pt, FPlanes F,
           const double *__restrict__ geo, const double *__restrict__ intr, const double *__restrict__ Vs,
           const double *__restrict__ U, double *__restrict__ S, const LmState *st, int gate) {
  if (!gate_open(st, gate)) return;
  const int b = (blockIdx.x * BA_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= n_blk) return;
  const int ci = blk_i[b], cj = blk_j[b];
  const double fx = ldg1(intr), fy = ldg1(intr + 1);
  CamRec ri, rj;
  load_camrec(geo, ci, ri);
  load_camrec(geo, cj, rj);
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  for (int e = blk_ptr[b] + lane; e < blk_ptr[b + 1]; e += 32) {
    const unsigned long long pr = pairs[e];
    const int oa = (int)(pr >> 32), ob = (int)(pr & 0xffffffffu);
    const int p = __ldg(pm_pt + oa);
    double vs[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) vs[k] = ldg1(Vs + 6 * (size_t)p + k);
    const ObsGeo ga = load_geo(F, oa, fx, fy), gb = load_geo(F, ob, fx, fy);
    double a0[6], a1[6], pa0[3], pa1[3], b0[6], b1[6], pb0[3], pb1[3];
    sp_rows(ga, ri.R, a0, a1, pa0, pa1);
    sp_rows(gb, rj.R, b0, b1, pb0, pb1);
    double v0[3], v1[3];
    sym3_mul(vs, pa0, v0);
    sym3_mul(vs, pa1, v1);
    const double m00 = v0[0] * pb0[0] + v0[1] * pb0[1] + v0[2] * pb0[2];
    const double m01 = v0[0] * pb1[0] + v0[1] * pb1[1] + v0[2] * pb1[2];
    const double m10 = v1[0] * pb0[0] + v1[1] * pb0[1] + v1[2] * pb0[2];
    const double m11 = v1[0] * pb1[0] + v1[1] * pb1[1] + v1[2] * pb1[2];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double h0 = m00 * b0[c] + m01 * b1[c], h1 = m10 * b0[c] + m11 * b1[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[r * 6 + c] += a0[r] * h0 + a1[r] * h1;
    }
  }
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = warp_sum(acc[k]);
  // lanes 0..35 -> lane k writes entry k (two rounds)
  double *Sb = S + 36 * (size_t)b;
#pragma unroll
  for (int k = 0; k < 36; ++k) {
    if (lane == (k & 31)) {
      const int r = k / 6, c = k - 6 * (k / 6);
      const double u = ci == cj ? U[36 * (size_t)ci + k] : 0.0;
      Sb[k] = u - acc[k] * (ri.s[r] * rj.s[c]);
    }
  }
}

// ------------------------------------------------------------------ y = S x  (without the LM damping)
// One warp per camera row; a lane multiplies whole 6x6 blocks (256-bit loads), the
// lane partials are combined by a fixed butterfly.
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__global__ void __launch_bounds__(BA_THREADS)
k_bsr_spmv(int n_cam, const int32_t *__restrict__ ent_ptr, const uint32_t *__restrict__ ent_blk,
           const int32_t *__restrict__ ent_col, const double *__restrict__ S, const double *__restrict__ x,
           double *__restrict__ y, const LmState *st, int gate) {
  if (!gate_open(st, gate)) return;
  const int r = (blockIdx.x * BA_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n_cam) return;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int e = ent_ptr[r] + lane; e < ent_ptr[r + 1]; e += 32) {
    const uint32_t eb = __ldg(ent_blk + e);
    const int col = __ldg(ent_col + e);
    const double *B = S + 36 * (size_t)(eb & 0x7fffffffu);
    double m[36], xv[6];
#pragma unroll
    for (int k = 0; k < 9; ++k) ld256(B + 4 * k, m[4 * k], m[4 * k + 1], m[4 * k + 2], m[4 * k + 3]);
    load6(x + 6 * (size_t)col, xv);
    if (eb & 0x80000000u) {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[a] += m[k * 6 + a] * xv[k];
    } else {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[a] += m[a * 6 + k] * xv[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) store6(y + 6 * (size_t)r, acc);
}
