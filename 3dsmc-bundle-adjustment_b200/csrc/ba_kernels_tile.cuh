// ba_kernels_tile.cuh -- single-pass ("tile-fused") implicit Schur product.
//
//   y_c = sum_{o in c} Jc_o^T ( Jc_o x_c - Jp_o V_p^-1 sum_{o' in p} Jp_o'^T Jc_o' x_c' )
//
// The two-pass form (ba_kernels_fact.cuh) streams the factored store twice per
// PCG iteration: point-major for t_p, camera-major for y_c, plus a 32 B/obs
// gather of t_p.  Sequential-SLAM problems have strong index locality: points
// are numbered by first appearance (src/OptimizationUtils.cpp:271-276) and
// tracks are runs of nearby keyframes, so the observations of BA_TILE_PTS
// consecutive points touch only a short run of cameras.  The fused kernel gives
// one CTA one such point tile and does BOTH reductions inside the CTA:
//
//   phase 0  load the tile's observations ONCE (32 B geometry + 4 B packed
//            indices each, kept in registers), stage the cameras' (s.*x, R)
//            records and the points' pre-scaled V^-1 in shared memory
//   phase 1  v_o = Jp_o^T (Jc_o x_c)            -> smem, point-major slot
//   phase 2  t_p = Vs_p sum_o v_o               (one thread per point, fixed order)
//   phase 3  c_o = Jc_o^T (Jc_o x_c - Jp_o t_p) -> smem, camera-major slot
//   phase 4  per (camera, component) segment sums -> one 6-vector per
//            (tile, camera), written where the camera's partial list expects it
//
// The tile's observation stream is stored in TILE-LOCAL CAMERA-MAJOR order (so a
// warp's lanes mostly share one camera record: shared-memory broadcast), and
// every observation carries rank (its point-major position in the tile, 16 b),
// local point (8 b) and camera slot (8 b) in one 32-bit word.  All sums run in a
// fixed order; no floating-point atomics.  A final per-camera kernel adds the
// camera's tile partials in tile order (k_pcg_q / k_pcg_reset, unchanged).
//
// Algorithmic traffic: 36 B/obs + 48 B/point (+ 48 B per (tile, camera) partial),
// against 104 B/obs + 104 B/point for the two-pass factored product and
// 320 B/obs for the materialised one.
//
// Inputs whose tiles span more than BA_TILE_MAXSPAN cameras or hold more than
// BA_TILE_MAXCAP observations use the two-pass kernels instead.
#pragma once
#include "ba_kernels_fact.cuh"

#define BA_TILE_MAXSPAN 64
#define BA_TILE_MAXCAP 2048
#define BA_TILE_REC 18

struct TileMeta {
  int32_t o0, n, lo, span;  // first observation, count, first camera, cameras spanned
};

// ------------------------------------------------------------------ setup (integer-only)
__global__ void __launch_bounds__(BA_THREADS)
kt_tile_meta(int n_pt, const int32_t *__restrict__ pt_rowptr, const int32_t *__restrict__ pm_cam, TileMeta *__restrict__ meta,
             int32_t *maxima /* [0] span, [1] obs */) {
  __shared__ int lo_s[BA_WARPS], hi_s[BA_WARPS];
  const int p0 = blockIdx.x * BA_TILE_PTS, p1 = min(n_pt, p0 + BA_TILE_PTS);
  const int o0 = pt_rowptr[p0], o1 = pt_rowptr[p1];
  int lo = 0x7fffffff, hi = -1;
  for (int s = o0 + threadIdx.x; s < o1; s += BA_THREADS) {
    const int c = pm_cam[s];
    lo = min(lo, c);
    hi = max(hi, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(BA_FULL, lo, o));
    hi = max(hi, __shfl_xor_sync(BA_FULL, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    lo_s[threadIdx.x >> 5] = lo;
    hi_s[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < BA_WARPS; ++w) {
      lo = min(lo, lo_s[w]);
      hi = max(hi, hi_s[w]);
    }
    TileMeta m;
    m.o0 = o0;
    m.n = o1 - o0;
    m.lo = hi >= lo ? lo : 0;
    m.span = hi >= lo ? hi - lo + 1 : 0;
    meta[blockIdx.x] = m;
    atomicMax(maxima, m.span);
    atomicMax(maxima + 1, m.n);
  }
}

__global__ void k_fill_i32(int n, int32_t *p, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// per tile: stable counting sort of its observations by camera, the permuted
// (tile-camera-major) input arrays of the linearisation, the packed per-observation
// word and the camera segment table.  Dynamic smem: int place[cap].
__global__ void __launch_bounds__(BA_THREADS)
kt_tile_build(const TileMeta *__restrict__ meta, const int32_t *__restrict__ pm_cam, const int32_t *__restrict__ pm_pt,
              const double2 *__restrict__ pm_uv, int32_t *__restrict__ tm_cam, int32_t *__restrict__ tm_pt,
              double2 *__restrict__ tm_uv, uint32_t *__restrict__ aux, uint16_t *__restrict__ tile_seg, int32_t *cam_tmin,
              int32_t *cam_tmax) {
  extern __shared__ int place[];
  __shared__ int cnt[BA_TILE_MAXSPAN + 1], cur[BA_TILE_MAXSPAN];
  const int t = blockIdx.x;
  const TileMeta m = meta[t];
  for (int i = threadIdx.x; i <= BA_TILE_MAXSPAN; i += BA_THREADS) cnt[i] = 0;
  for (int i = threadIdx.x; i < BA_TILE_MAXSPAN; i += BA_THREADS) cur[i] = 0;
  __syncthreads();
  for (int l = threadIdx.x; l < m.n; l += BA_THREADS) atomicAdd(&cnt[pm_cam[m.o0 + l] - m.lo + 1], 1);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < m.span; ++i) cnt[i + 1] += cnt[i];  // cnt[slot] = segment start
  __syncthreads();
  for (int l = threadIdx.x; l < m.n; l += BA_THREADS) {
    const int slot = pm_cam[m.o0 + l] - m.lo;
    place[cnt[slot] + atomicAdd(&cur[slot], 1)] = l;
  }
  __syncthreads();
  // ascending point-major rank inside every segment == the stable sort
  if (threadIdx.x < m.span) {
    const int b = cnt[threadIdx.x], e = cnt[threadIdx.x + 1];
    for (int i = b + 1; i < e; ++i) {
      const int v = place[i];
      int j = i - 1;
      while (j >= b && place[j] > v) {
        place[j + 1] = place[j];
        --j;
      }
      place[j + 1] = v;
    }
    if (e > b) {
      atomicMin(cam_tmin + m.lo + threadIdx.x, t);
      atomicMax(cam_tmax + m.lo + threadIdx.x, t);
    }
  }
  __syncthreads();
  const int p0 = t * BA_TILE_PTS;
  for (int j = threadIdx.x; j < m.n; j += BA_THREADS) {
    const int l = place[j], s = m.o0 + l;
    const int c = pm_cam[s], p = pm_pt[s];
    tm_cam[m.o0 + j] = c;
    tm_pt[m.o0 + j] = p;
    tm_uv[m.o0 + j] = pm_uv[s];
    aux[m.o0 + j] = (uint32_t)l | ((uint32_t)(p - p0) << 16) | ((uint32_t)(c - m.lo) << 24);
  }
  for (int i = threadIdx.x; i <= BA_TILE_MAXSPAN; i += BA_THREADS)
    tile_seg[(size_t)t * (BA_TILE_MAXSPAN + 1) + i] = (uint16_t)(i <= m.span ? cnt[i] : cnt[m.span]);
}

// per camera: its tiles in ascending order -> number of partials (ASSIGN = 0) or
// the output slot of every (tile, camera) pair (ASSIGN = 1)
template <int ASSIGN>
__global__ void __launch_bounds__(BA_THREADS)
kt_cam_tiles(int n_cam, const TileMeta *__restrict__ meta, const uint16_t *__restrict__ tile_seg, const int32_t *__restrict__ cam_tmin,
             const int32_t *__restrict__ cam_tmax, const int32_t *__restrict__ tpart_ptr, int32_t *__restrict__ tcnt,
             int32_t *__restrict__ tile_out) {
  const int c = blockIdx.x * BA_THREADS + threadIdx.x;
  if (c >= n_cam) return;
  int k = 0;
  const int base = ASSIGN ? tpart_ptr[c] : 0;
  for (int t = cam_tmin[c]; t <= cam_tmax[c]; ++t) {
    const TileMeta m = meta[t];
    const int slot = c - m.lo;
    if (slot < 0 || slot >= m.span) continue;
    const uint16_t *sg = tile_seg + (size_t)t * (BA_TILE_MAXSPAN + 1) + slot;
    if (sg[1] == sg[0]) continue;
    if (ASSIGN) tile_out[(size_t)t * BA_TILE_MAXSPAN + slot] = base + k;
    ++k;
  }
  if (!ASSIGN) tcnt[c] = k;
}

// ------------------------------------------------------------------ linearise into the tile store (g only)
__global__ void __launch_bounds__(BA_THREADS)
kt_linearize(int n_obs, const int32_t *__restrict__ cam_idx, const int32_t *__restrict__ pt_idx, const double2 *__restrict__ uv,
             const double *__restrict__ pose, const double *__restrict__ pt, const double *__restrict__ intr, CostParams cp,
             double2 *__restrict__ g0, double2 *__restrict__ g1, const LmState *st, int gate) {
  if (!gate_open(st, gate)) return;
  const int i = blockIdx.x * BA_THREADS + threadIdx.x;
  if (i >= n_obs) return;
  const int c = cam_idx[i], p = pt_idx[i];
  const double *P = pose + 7 * (size_t)c;
  const double q[4] = {ldg1(P), ldg1(P + 1), ldg1(P + 2), ldg1(P + 3)};
  double R[9];
  quat_to_R(q, R);
  const double *X3 = pt + 3 * (size_t)p;
  const double d0 = ldg1(X3) - ldg1(P + 4), d1 = ldg1(X3 + 1) - ldg1(P + 5), d2 = ldg1(X3 + 2) - ldg1(P + 6);
  const double X = (R[0] * d0 + R[3] * d1) + R[6] * d2;
  const double Y = (R[1] * d0 + R[4] * d1) + R[7] * d2;
  const double Z = (R[2] * d0 + R[5] * d1) + R[8] * d2;
  const double iz = 1.0 / Z;
  const double xz = X * iz, yz = Y * iz;
  const double2 m = lds2(uv + i);
  const double r0 = cp.sw_repr * ((ldg1(intr) * xz + ldg1(intr + 2)) - m.x);
  const double r1 = cp.sw_repr * ((ldg1(intr + 1) * yz + ldg1(intr + 3)) - m.y);
  double rho0;
  const double hs = huber_scale(cp.hub_repr, r0 * r0 + r1 * r1, rho0);
  sts2(g0 + i, make_double2(xz, yz));
  sts2(g1 + i, make_double2(iz, cp.sw_repr * hs));
}

// ------------------------------------------------------------------ the fused product
// Shared-memory budget is what bounds this kernel (every observation moves its
// camera record, 3 + 3 + 6 + 6 doubles through the banks), so:
//   * the camera record is (s.*x (6), q (4)) = 5 LDS.128 instead of (s.*x, R) = 8:
//     the rotation is applied from the quaternion (R v = v + 2w (u x v) + 2 u x (u x v),
//     identical to Eigen's toRotationMatrix for any q) -- flops are free here;
//   * the column stride of the staging buffer is CAP + 8 doubles and phase 4 gives
//     8 lanes to every (camera, component): a half-warp reads two columns 8 banks
//     apart, conflict-free.
__device__ __forceinline__ void quat_rot(const double q[4], const double v[3], double o[3]) {  // o = R(q) v
  const double t0 = 2.0 * (q[1] * v[2] - q[2] * v[1]);
  const double t1 = 2.0 * (q[2] * v[0] - q[0] * v[2]);
  const double t2 = 2.0 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * t0 + (q[1] * t2 - q[2] * t1);
  o[1] = v[1] + q[3] * t1 + (q[2] * t0 - q[0] * t2);
  o[2] = v[2] + q[3] * t2 + (q[0] * t1 - q[1] * t0);
}
__device__ __forceinline__ void quat_rot_t(const double q[4], const double v[3], double o[3]) {  // o = R(q)^T v
  const double t0 = 2.0 * (q[1] * v[2] - q[2] * v[1]);
  const double t1 = 2.0 * (q[2] * v[0] - q[0] * v[2]);
  const double t2 = 2.0 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] - q[3] * t0 + (q[1] * t2 - q[2] * t1);
  o[1] = v[1] - q[3] * t1 + (q[2] * t0 - q[0] * t2);
  o[2] = v[2] - q[3] * t2 + (q[0] * t1 - q[1] * t0);
}

#define BA_TILE_QREC 10
template <int NPT>
__global__ void __launch_bounds__(BA_THREADS, (NPT <= 3 ? 3 : (NPT <= 4 ? 2 : 1)))
kt_schur_fused(int n_pt, const int32_t *__restrict__ pt_rowptr, const TileMeta *__restrict__ meta,
               const uint16_t *__restrict__ tile_seg, const int32_t *__restrict__ tile_out, const uint32_t *__restrict__ aux,
               const double2 *__restrict__ g0, const double2 *__restrict__ g1, const double *__restrict__ geo,
               const double *__restrict__ pose, const double *__restrict__ x, const double *__restrict__ intr,
               const double *__restrict__ Vs, double *__restrict__ part, const LmState *st, int gate) {
  if (!gate_open(st, gate)) return;
  constexpr int CAP = NPT * BA_THREADS;
  constexpr int CS = CAP + 8;  // column stride (doubles)
  extern __shared__ __align__(16) double smem[];
  double *buf = smem;                                  // [6][CS]  (phases 1-2 use [3][CS])
  double *tp = buf + 6 * CS;                           // [3][BA_TILE_PTS]
  double *rec = tp + 3 * BA_TILE_PTS;                  // [MAXSPAN][QREC]: s.*x (6), q (4)
  double *scs = rec + BA_TILE_MAXSPAN * BA_TILE_QREC;  // [MAXSPAN][6] camera column scale
  __shared__ int seg_s[BA_TILE_MAXSPAN + 1], out_s[BA_TILE_MAXSPAN];
  const int tid = threadIdx.x, t = blockIdx.x;
  const TileMeta m = meta[t];
  const double fx = ldg1(intr), fy = ldg1(intr + 1);

  // ---- phase 0: every global load of the tile, issued back to back
  ObsGeo o[NPT];
  uint32_t a[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int l = tid + j * BA_THREADS;
    if (l < m.n) {
      const double2 u = lds2(g0 + m.o0 + l), w = lds2(g1 + m.o0 + l);
      a[j] = __ldg(aux + m.o0 + l);
      o[j].xz = u.x;
      o[j].yz = u.y;
      o[j].iz = w.x;
      o[j].wfx = w.y * fx;
      o[j].wfy = w.y * fy;
    } else {
      a[j] = (uint32_t)CAP;  // rank CAP = the padding of the staging columns: a harmless dummy slot
      o[j].xz = o[j].yz = o[j].iz = o[j].wfx = o[j].wfy = 0.0;
    }
  }
  const int p = t * BA_TILE_PTS + tid;
  const bool own = tid < BA_TILE_PTS && p < n_pt;
  int pb = 0, pe = 0;
  double vs[6] = {0, 0, 0, 0, 0, 0};
  if (own) {
    pb = pt_rowptr[p] - m.o0;
    pe = pt_rowptr[p + 1] - m.o0;
#pragma unroll
    for (int k = 0; k < 6; ++k) vs[k] = ldg1(Vs + 6 * (size_t)p + k);
  }
  for (int idx = tid; idx < m.span * 16; idx += BA_THREADS) {
    const int rc = idx >> 4, k = idx & 15;
    const size_t c = (size_t)(m.lo + rc);
    if (k < 6) {
      const double s = geo[c * BA_CAMREC + 9 + k];
      scs[rc * 6 + k] = s;
      rec[rc * BA_TILE_QREC + k] = s * x[c * 6 + k];
    } else if (k < 10)
      rec[rc * BA_TILE_QREC + k] = pose[7 * c + (k - 6)];
  }
  if (tid <= m.span) seg_s[tid] = tile_seg[(size_t)t * (BA_TILE_MAXSPAN + 1) + tid];
  if (tid < m.span) out_s[tid] = tile_out[(size_t)t * BA_TILE_MAXSPAN + tid];
  __syncthreads();

  // ---- phase 1: v = Jpu^T (Jcu (s.*x)) into the point-major slot
  // Whole warps beyond the tile's observations skip a round (warp-uniform branch); inside a partially filled
  // warp the idle lanes compute on zero weights into the dummy slot -- no per-lane divergence.
  double a0[NPT], a1[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    a0[j] = a1[j] = 0.0;
    if ((tid & ~31) + j * BA_THREADS < m.n) {
      const int slot = a[j] >> 24, rank = a[j] & 0xffff;
      const double2 *cx = reinterpret_cast<const double2 *>(rec + slot * BA_TILE_QREC);
      const double2 c0 = cx[0], c1 = cx[1], c2 = cx[2], c3 = cx[3], c4 = cx[4];
      const double xx[6] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
      const double q[4] = {c3.x, c3.y, c4.x, c4.y};
      jc_dot(o[j], xx, a0[j], a1[j]);
      // Jpu^T a = R (al0, al1, -(xz al0 + yz al1))
      const double al0 = (o[j].wfx * o[j].iz) * a0[j], al1 = (o[j].wfy * o[j].iz) * a1[j];
      const double al[3] = {al0, al1, -(o[j].xz * al0 + o[j].yz * al1)};
      double v[3];
      quat_rot(q, al, v);
      buf[rank] = v[0];
      buf[CS + rank] = v[1];
      buf[2 * CS + rank] = v[2];
    }
  }
  __syncthreads();

  // ---- phase 2: t_p = Vs_p * (sum of the point's run, in point-major order)
  if (own) {
    double b[3] = {0, 0, 0}, tt[3];
    for (int l = pb; l < pe; ++l) {
      b[0] += buf[l];
      b[1] += buf[CS + l];
      b[2] += buf[2 * CS + l];
    }
    sym3_mul(vs, b, tt);
    tp[tid] = tt[0];
    tp[BA_TILE_PTS + tid] = tt[1];
    tp[2 * BA_TILE_PTS + tid] = tt[2];
  }
  __syncthreads();

  // ---- phase 3: c = Jcu^T (Jcu (s.*x) - Jpu t_p) into the camera-major slot (= load order)
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int l = tid + j * BA_THREADS;
    if ((tid & ~31) + j * BA_THREADS < m.n) {
      const int slot = a[j] >> 24, lp = (a[j] >> 16) & 0xff;
      const double2 *cx = reinterpret_cast<const double2 *>(rec + slot * BA_TILE_QREC);
      const double2 c3 = cx[3], c4 = cx[4];
      const double q[4] = {c3.x, c3.y, c4.x, c4.y};
      const double tt[3] = {tp[lp], tp[BA_TILE_PTS + lp], tp[2 * BA_TILE_PTS + lp]};
      double u[3];
      quat_rot_t(q, tt, u);
      const double b0 = (o[j].wfx * o[j].iz) * (u[0] - o[j].xz * u[2]);
      const double b1 = (o[j].wfy * o[j].iz) * (u[1] - o[j].yz * u[2]);
      // c = Jcu^T (e0, e1), written out (no accumulator to clear)
      const double e0 = o[j].wfx * (a0[j] - b0), e1 = o[j].wfy * (a1[j] - b1);
      buf[l] = -(o[j].iz * e0);
      buf[CS + l] = -(o[j].iz * e1);
      buf[2 * CS + l] = o[j].iz * (o[j].xz * e0 + o[j].yz * e1);
      buf[3 * CS + l] = (o[j].xz * o[j].yz) * e0 + (1.0 + o[j].yz * o[j].yz) * e1;
      buf[4 * CS + l] = -((1.0 + o[j].xz * o[j].xz) * e0 + (o[j].xz * o[j].yz) * e1);
      buf[5 * CS + l] = o[j].yz * e0 - o[j].xz * e1;
    }
  }
  __syncthreads();

  // ---- phase 4: segment sums, 8 lanes per (camera, component), combined in a fixed tree
  const int n_work = m.span * 48;
  for (int base = 0; base < n_work; base += BA_THREADS) {
    const int idx = base + tid;
    const bool act = idx < n_work;
    const int sub = idx & 7, pair = idx >> 3;
    const int slot = act ? pair / 6 : 0, k = act ? pair - 6 * (pair / 6) : 0;
    const int lo = seg_s[slot], hi = act ? seg_s[slot + 1] : lo;
    double s = 0.0;
    const double *col = buf + k * CS;
    for (int l = lo + sub; l < hi; l += 8) s += col[l];
    s += __shfl_xor_sync(BA_FULL, s, 1);
    s += __shfl_xor_sync(BA_FULL, s, 2);
    s += __shfl_xor_sync(BA_FULL, s, 4);
    if (act && sub == 0 && hi > lo) part[6 * (size_t)out_s[slot] + k] = s * scs[slot * 6 + k];
  }
}
