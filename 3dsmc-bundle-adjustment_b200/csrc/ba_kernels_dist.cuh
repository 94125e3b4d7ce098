// ba_kernels_dist.cuh -- the persistent PCG of the block-sparse solver with the PRODUCT row-sharded
// over the GPUs of one node and the exchange fused into the kernel over NVLink peer memory.
//
// One process per GPU; S, b, M^-1 are complete and identical on every rank (S is all-reduced once
// per LM iteration).  Per PCG iteration the block-CSR product q = S p + D^2 p is the only heavy
// phase (87 MB of L2 traffic at config 5), so that is what is sharded: rank k owns the camera blocks
// wb = k, k + N, ... (32 consecutive cameras each) and computes q and p.q for ITS rows only.  The lane
// that holds a result stores it straight into every rank's memory (cudaIpc-mapped exchange buffers)
// as a 16-byte FLAG-IN-DATA slot of two tagged 8-byte words {32 value bits | 32-bit round tag} (one st.v2.u64; each
// 8-byte word is single-copy atomic, the pair need not be): the consumer polls until both words carry the current tag.  No fence, no separate flag, no
// cross-GPU barrier -- measured on 2 x B200 (profiles/r01_nvlink_latency.txt): 1.14 us one way for a
// flag-in-data store against 4.4-6.2 us for "stores + fence.sys + flag" (a first version of this kernel
// with two fence + flag barriers per iteration ran at 48 us per PCG iteration against 29 us on one GPU).
// The vector phase (alpha, x, r, z = M^-1 r, controller) is cheap and runs REPLICATED on every rank from
// the complete q, so one exchange per PCG iteration suffices.
//
//   phase I    own rows: q_r, p.q_r  -> slots on every rank        (p_new kept complete locally)
//   sum        64-row slices of p.q (fixed granularity), then the slice partials: alpha
//   phase II   ALL camera blocks (replicated): poll q, x += alpha p, r -= alpha q, z = M^-1 r, partials
//   grid barrier (local), controller (every CTA of every rank: same numbers, same order)
//
// Slots are double-buffered by the exchange round (epoch & 1): a rank can run at most one round ahead of
// the slowest one, because finishing a round needs everybody's data of that round.  Epochs increase
// monotonically over the life of the context, so stale slots never match.  Every rank computes the same
// sums in the same order: all ranks hold bit-identical iterates.  Polls time out (4 s) into an abort
// flag, so a dead peer cannot hang the other GPUs.
#pragma once
#include "ba_kernels_sparse.cuh"

#define BA_MAX_RANKS 8
#define BA_PQ_SLICE 64

struct __align__(16) LLSlot {
  double v;
  unsigned long long e;
};
struct PcgFan {
  int n_ranks, rank;
  LLSlot *q[BA_MAX_RANKS];    // [2][6 n_cam]  q of every camera, written by the row owners
  LLSlot *pq[BA_MAX_RANKS];   // [2][n_cam]    p.q of every row
  LLSlot *slice;              // [2][n_slices] local: 64-row partial sums of p.q
  unsigned long long *epoch;  // own running round counter (survives launches)
  int *abort_flag;            // set on a poll time-out
};

// Slot encoding (NCCL-LL style): the 16-byte slot is TWO independent 8-byte words {32 data bits | 32-bit tag << 32}.
// The PTX memory model only guarantees single-copy atomicity per aligned scalar of a vector access, so each 8-byte word
// carries its own tag and the consumer accepts a slot only when BOTH words show the tag of the current round: a torn
// delivery (one word new, one old) can never be read as a value.  tag = (round & 0x7fffffff) | 0x80000000: never zero
// (fresh buffers are zeroed), a stale slot could only match 2^31 rounds later.
__device__ __forceinline__ unsigned long long ll_tag(unsigned long long e) { return ((e & 0x7fffffffull) | 0x80000000ull) << 32; }
__device__ __forceinline__ void ll_store(LLSlot *p, double v, unsigned long long e) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v), tag = ll_tag(e);
  const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// spins until both words of the slot carry the tag of round e; on time-out / abort returns 0.0
__device__ __forceinline__ double ll_wait(const LLSlot *p, unsigned long long e, int *abort_flag) {
  unsigned long long a, b;
  const unsigned long long tag = ll_tag(e), hi = 0xffffffff00000000ull;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if ((a & hi) != tag || (b & hi) != tag) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int spin = 0;; ++spin) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
      if ((a & hi) == tag && (b & hi) == tag) break;
      if ((spin & 63) == 63) {
        if (*((volatile int *)abort_flag)) return 0.0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {
          *((volatile int *)abort_flag) = 1;
          return 0.0;
        }
      }
    }
  }
  return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}
// local grid barrier that gives up when the abort flag is raised
__device__ __forceinline__ void grid_barrier_abortable(unsigned int *bar, unsigned int &epoch, int *abort_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    int spin = 0;
    while (*((volatile unsigned int *)bar) < epoch) {
      if ((++spin & 255) == 0 && *((volatile int *)abort_flag)) break;
    }
    __threadfence();
  }
  __syncthreads();
}

// q of one row (lanes 0..5 hold a component each after the butterfly) and v.q -> the slots of every rank
template <int WITH_PQ>
__device__ __forceinline__ void dist_row_publish(const PcgFan &f, unsigned long long e, size_t qoff, size_t pqoff, int row, int lane,
                                                 const double acc[6], double pv, double dk) {
  const double qv = pick6(acc, lane) + dk * pv;
  if (lane < 6)
    for (int k = 0; k < f.n_ranks; ++k) ll_store(f.q[k] + qoff + 6 * (size_t)row + lane, qv, e);
  if (WITH_PQ) {
    const double t = pv * qv;
    double s = __shfl_sync(BA_FULL, t, 0);
#pragma unroll
    for (int k = 1; k < 6; ++k) s += __shfl_sync(BA_FULL, t, k);
    if (lane < f.n_ranks) ll_store(f.pq[lane] + pqoff + row, s, e);  // lane k serves rank k
  }
}

// the warp's first rows from its shared-memory entry cache (as bsr_rows_cached): one dependent L2 round trip per row
template <int WITH_PQ>
__device__ __forceinline__ void dist_rows_cached(const PcgFan &f, unsigned long long e, int n_cam, int lane, int n_cached,
                                                 const int4 *rowinfo, const int2 *ecache, const double *__restrict__ S,
                                                 const double *__restrict__ dsq, const double *za, const double *pb, double beta,
                                                 bool use_pb) {
  const size_t qoff = (size_t)(e & 1) * 6 * n_cam, pqoff = (size_t)(e & 1) * n_cam;
  int slot = 0;
  for (int i = 0; i < n_cached; ++i) {
    const int4 be = rowinfo[i];  // (first entry, end, row, -)
    const int row = be.z;
    double zk = 0.0, pk = 0.0, dk = 0.0;
    if (lane < 6) {
      zk = __ldcg(za + 6 * (size_t)row + lane);
      if (use_pb) pk = __ldcg(pb + 6 * (size_t)row + lane);
      dk = __ldg(dsq + 6 * (size_t)row + lane);
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int en = be.x; en < be.y; en += 32, ++slot) {
      const int2 ent = ecache[slot * 32 + lane];
      if (en + lane < be.y) bsr_entry(ent, S, za, pb, beta, use_pb, acc);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
    dist_row_publish<WITH_PQ>(f, e, qoff, pqoff, row, lane, acc, use_pb ? zk + beta * pk : zk, dk);
  }
}

// rows order[gw], order[gw + nw], ... (< n_rows) of q = S v + dsq .* v, v = za (+ beta pb): the six
// components and (WITH_PQ) v.q of the row go to every rank's slots of round e
template <int WITH_PQ>
__device__ __forceinline__ void dist_rows(const PcgFan &f, unsigned long long e, int n_cam, int n_rows, int gw, int nw, int lane,
                                          const int32_t *__restrict__ order, const int32_t *__restrict__ ent_ptr,
                                          const int2 *__restrict__ ent, const double *__restrict__ S,
                                          const double *__restrict__ dsq, const double *za, const double *pb, double beta,
                                          bool use_pb) {
  const int2 none = make_int2(0, 0);
  const size_t qoff = (size_t)(e & 1) * 6 * n_cam, pqoff = (size_t)(e & 1) * n_cam;
  int idx = gw;
  if (idx >= n_rows) return;
  int row = __ldg(order + idx);
  int row1 = idx + nw < n_rows ? __ldg(order + idx + nw) : -1;
  int b0 = __ldg(ent_ptr + row), e0 = __ldg(ent_ptr + row + 1);
  int b1 = 0, e1 = 0;
  if (row1 >= 0) {
    b1 = __ldg(ent_ptr + row1);
    e1 = __ldg(ent_ptr + row1 + 1);
  }
  int2 en0 = b0 + lane < e0 ? __ldg(ent + b0 + lane) : none;
  int2 en0b = b0 + lane + 32 < e0 ? __ldg(ent + b0 + lane + 32) : none;
  for (; idx < n_rows; idx += nw) {
    const int2 en1 = b1 + lane < e1 ? __ldg(ent + b1 + lane) : none;
    const int2 en1b = b1 + lane + 32 < e1 ? __ldg(ent + b1 + lane + 32) : none;
    int b2 = 0, e2 = 0, row2 = -1;
    if (idx + 2 * nw < n_rows) {
      row2 = __ldg(order + idx + 2 * nw);
      b2 = __ldg(ent_ptr + row2);
      e2 = __ldg(ent_ptr + row2 + 1);
    }
    double zk = 0.0, pk = 0.0, dk = 0.0;
    if (lane < 6) {
      zk = __ldcg(za + 6 * (size_t)row + lane);
      if (use_pb) pk = __ldcg(pb + 6 * (size_t)row + lane);
      dk = __ldg(dsq + 6 * (size_t)row + lane);
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (b0 + lane < e0) bsr_entry(en0, S, za, pb, beta, use_pb, acc);
    if (b0 + lane + 32 < e0) bsr_entry(en0b, S, za, pb, beta, use_pb, acc);
    for (int en = b0 + lane + 64; en < e0; en += 32) bsr_entry(__ldg(ent + en), S, za, pb, beta, use_pb, acc);
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
    dist_row_publish<WITH_PQ>(f, e, qoff, pqoff, row, lane, acc, use_pb ? zk + beta * pk : zk, dk);
    row = row1; row1 = row2;
    b0 = b1; e0 = e1; en0 = en1; en0b = en1b;
    b1 = b2; e1 = e2;
  }
}

// sum of p.q over all rows (every CTA gets it): 64-row slices by the CTAs of this GPU, then the slices
__device__ __forceinline__ double ll_sum_pq(const PcgFan &f, unsigned long long e, int n_cam, double *red) {
  const int tid = threadIdx.x, n_slices = (n_cam + BA_PQ_SLICE - 1) / BA_PQ_SLICE;
  const LLSlot *pq = f.pq[f.rank] + (size_t)(e & 1) * n_cam;
  LLSlot *sl = f.slice + (size_t)(e & 1) * n_slices;
  // stage 1: four slices per CTA pass (64 threads each), fixed order inside a slice
  for (int s0 = blockIdx.x * 4; s0 < n_slices; s0 += gridDim.x * 4) {
    const int s = s0 + (tid >> 6), row = s * BA_PQ_SLICE + (tid & 63);
    double v = 0.0;
    if (s < n_slices && row < n_cam) v = ll_wait(pq + row, e, f.abort_flag);
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if ((tid & 63) == 0 && s < n_slices) ll_store(sl + s, red[tid >> 5] + red[(tid >> 5) + 1], e);
  }
  // stage 2: every CTA adds all the slice partials
  double v = 0.0;
  for (int i = tid; i < n_slices; i += BA_THREADS) v += ll_wait(sl + i, e, f.abort_flag);
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < BA_WARPS; ++k) s += red[k];
    red[BA_WARPS] = s;
  }
  __syncthreads();
  return red[BA_WARPS];
}

// phase II for camera block wb (lane = camera), replicated on every rank; q comes from the slots
template <int RESET, int CACHED>
__device__ __forceinline__ void ll_update_warp(const PcgFan &f, unsigned long long e, int n_cam, int wb, int lane, double alpha,
                                               bool skip_r, const double *__restrict__ b, const double *__restrict__ Minv,
                                               const double *minv_s, const double breg[6], double *x, double *r, double *z,
                                               const double *pnew, double *part_rho, double *part_Q) {
  const int c = wb * 32 + lane;
  double rz = 0.0, xq = 0.0;
  if (c < n_cam) {
    double xv[6], rv[6], qv[6], bv[6], zv[6];
    load6cg(x + 6 * (size_t)c, xv);
    if (!RESET) {
      double pv[6];
      load6cg(pnew + 6 * (size_t)c, pv);
#pragma unroll
      for (int k = 0; k < 6; ++k) xv[k] = xv[k] + alpha * pv[k];
      store6(x + 6 * (size_t)c, xv);
    }
    if (!skip_r) {
      const LLSlot *qs = f.q[f.rank] + (size_t)(e & 1) * 6 * n_cam + 6 * (size_t)c;
#pragma unroll
      for (int k = 0; k < 6; ++k) qv[k] = ll_wait(qs + k, e, f.abort_flag);
      if (CACHED) {
#pragma unroll
        for (int k = 0; k < 6; ++k) bv[k] = breg[k];
      } else
        load6(b + 6 * (size_t)c, bv);
      if (RESET) {
#pragma unroll
        for (int k = 0; k < 6; ++k) rv[k] = bv[k] - qv[k];
      } else {
        load6cg(r + 6 * (size_t)c, rv);
#pragma unroll
        for (int k = 0; k < 6; ++k) rv[k] = rv[k] - alpha * qv[k];
      }
      store6(r + 6 * (size_t)c, rv);
      if (CACHED) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sacc += minv_s[(a * 6 + k) * 32 + lane] * rv[k];
          zv[a] = sacc;
        }
      } else
        minv_mul(Minv, c, rv, zv);
      store6(z + 6 * (size_t)c, zv);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        rz += rv[k] * zv[k];
        xq += xv[k] * (bv[k] + rv[k]);
      }
    }
  }
  if (!skip_r) {
    rz = warp_sum(rz);
    xq = warp_sum(xq);
    if (lane == 0) {
      part_rho[wb] = rz;
      part_Q[wb] = xq;
    }
  }
}

__global__ void __launch_bounds__(BA_THREADS, 1)
k_pcg_sparse_dist(PcgFan f, int n_cam, int n_my_rows, const int32_t *__restrict__ my_rows, const int32_t *__restrict__ ent_ptr,
                  const int2 *__restrict__ ent, const double *__restrict__ S, const double *__restrict__ dsq,
                  const double *__restrict__ b, const double *__restrict__ Minv, double *x, double *r, double *z, double *pbuf0,
                  double *pbuf1, double *part_rho, double *part_Q, unsigned int *bar, LmOptions lo, LmState *st,
                  unsigned long long *prof) {
  if (st->done || st->pcg_done) return;  // identical on every CTA and every rank
  __shared__ double red[2 * BA_WARPS + 4];
  extern __shared__ double minv_all[];  // per warp: M^-1 [36][32] of its first camera block, entry cache, row records
  const int tid = threadIdx.x, lane = tid & 31;
  const int gtid = blockIdx.x * BA_THREADS + tid, nthreads = gridDim.x * BA_THREADS;
  const int gw = gtid >> 5, nw = nthreads >> 5;
  const int n_wb = (n_cam + 31) / 32;
  char *wsm = reinterpret_cast<char *>(minv_all) + (size_t)(tid >> 5) * BA_PCG_SMEM_PER_WARP;
  double *minv_s = reinterpret_cast<double *>(wsm);
  int2 *ecache = reinterpret_cast<int2 *>(wsm + 36 * 32 * 8);
  int4 *rowinfo = reinterpret_cast<int4 *>(ecache + BA_PCG_ENT_SLOTS * 32);
  int n_cached = 0;
  {
    int slot = 0;
    for (int idx = gw; idx < n_my_rows && n_cached < BA_PCG_ROWS; idx += nw) {
      const int row = my_rows[idx];
      const int b0 = ent_ptr[row], e0 = ent_ptr[row + 1];
      const int trips = (e0 - b0 + 31) >> 5;
      if (slot + trips > BA_PCG_ENT_SLOTS) break;
      if (lane == 0) rowinfo[n_cached] = make_int4(b0, e0, row, 0);
      for (int en = b0; en < e0; en += 32, ++slot) ecache[slot * 32 + lane] = en + lane < e0 ? ent[en + lane] : make_int2(0, 0);
      ++n_cached;
    }
  }
  double breg[6] = {0, 0, 0, 0, 0, 0};
  if (gw < n_wb) {
    const int c = gw * 32 + lane;
    if (c < n_cam) {
#pragma unroll
      for (int k = 0; k < 36; ++k) minv_s[k * 32 + lane] = Minv[36 * (size_t)c + k];
      load6(b + 6 * (size_t)c, breg);
    }
  }
  __syncwarp();
  int it = st->pcg_it;
  double rho = st->pcg_rho, beta = st->pcg_beta, Q0 = st->pcg_Q0;
  int fail = 0, brk = 0, iters_last = 0;
  unsigned int local_epoch = 0;
  unsigned long long e = *f.epoch;  // same on every CTA and rank: written back only at the very end
  double *pold = pbuf0, *pnew = pbuf1;
  unsigned long long t0 = 0, tacc[6] = {0, 0, 0, 0, 0, 0};
#define DPROF_TICK(slot)                                            \
  if (prof && blockIdx.x == 0 && tid == 0) {                        \
    unsigned long long t1_;                                         \
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1_));          \
    tacc[slot] += t1_ - t0;                                         \
    t0 = t1_;                                                       \
  }
  if (prof && blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));

  for (;;) {
    // ---- phase I: own rows of q = S p + D^2 p (p = z + beta p_old on the fly) -> slots of round e on every rank;
    //      p is kept complete locally (every rank updates all of it)
    ++e;
    dist_rows_cached<1>(f, e, n_cam, lane, n_cached, rowinfo, ecache, S, dsq, z, pold, beta, it > 1);
    dist_rows<1>(f, e, n_cam, n_my_rows, gw + n_cached * nw, nw, lane, my_rows, ent_ptr, ent, S, dsq, z, pold, beta, it > 1);
    for (int i = gtid; i < 6 * n_cam; i += nthreads) {
      const double zv = __ldcg(z + i);
      pnew[i] = it > 1 ? zv + beta * __ldcg(pold + i) : zv;
    }
    DPROF_TICK(0)
    const double pq = ll_sum_pq(f, e, n_cam, red);
    DPROF_TICK(1)
    if (*((volatile int *)f.abort_flag)) {
      fail = 1;
      iters_last = it;
      break;
    }
    if (pq <= 0.0 || isinf(pq) || isnan(pq)) {
      iters_last = it;
      brk = 1;
      break;
    }
    const double alpha = rho / pq;
    if (isinf(alpha)) {
      iters_last = it;
      fail = 1;
      break;
    }
    const bool reset = lo.reset_period > 0 && (it % lo.reset_period) == 0;
    // ---- phase II, replicated: all camera blocks
    if (gw < n_wb) ll_update_warp<0, 1>(f, e, n_cam, gw, lane, alpha, reset, b, Minv, minv_s, breg, x, r, z, pnew, part_rho, part_Q);
    for (int wb = gw + nw; wb < n_wb; wb += nw)
      ll_update_warp<0, 0>(f, e, n_cam, wb, lane, alpha, reset, b, Minv, minv_s, breg, x, r, z, pnew, part_rho, part_Q);
    DPROF_TICK(2)
    grid_barrier_abortable(bar, local_epoch, f.abort_flag);
    DPROF_TICK(3)
    if (reset) {
      // ---- residual reset: own rows of q = S x + D^2 x -> slots of the next round; r = b - q, z = M^-1 r everywhere
      ++e;
      dist_rows_cached<0>(f, e, n_cam, lane, n_cached, rowinfo, ecache, S, dsq, x, x, 0.0, false);
      dist_rows<0>(f, e, n_cam, n_my_rows, gw + n_cached * nw, nw, lane, my_rows, ent_ptr, ent, S, dsq, x, x, 0.0, false);
      if (gw < n_wb) ll_update_warp<1, 1>(f, e, n_cam, gw, lane, 0.0, false, b, Minv, minv_s, breg, x, r, z, pnew, part_rho, part_Q);
      for (int wb = gw + nw; wb < n_wb; wb += nw)
        ll_update_warp<1, 0>(f, e, n_cam, wb, lane, 0.0, false, b, Minv, minv_s, breg, x, r, z, pnew, part_rho, part_Q);
      grid_barrier_abortable(bar, local_epoch, f.abort_flag);
    }
    DPROF_TICK(4)

    // ---- controller: every CTA of every rank, same arrays, same order
    double rho_new, xq;
    block_sum2_cg(part_rho, part_Q, n_wb, red, rho_new, xq);
    iters_last = it;
    const double Q1 = -1.0 * xq;
    const double zeta = it * (Q1 - Q0) / Q1;
    if (zeta < lo.eta && it >= lo.min_pcg) break;
    Q0 = Q1;
    if (it >= lo.max_pcg) break;
    const double beta_new = rho_new / rho;
    if (rho_new == 0.0 || !isfinite(rho_new) || beta_new == 0.0 || !isfinite(beta_new)) {
      iters_last = it + 1;
      fail = 1;
      break;
    }
    rho = rho_new;
    beta = beta_new;
    ++it;
    double *t = pold;
    pold = pnew;
    pnew = t;
    DPROF_TICK(5)
  }
  if (prof && blockIdx.x == 0 && tid == 0)
    for (int k = 0; k < 6; ++k) prof[k] += tacc[k];
  // all CTAs of this GPU read *f.epoch at their start, before their first grid barrier; nobody gets here earlier
  if (blockIdx.x == 0 && tid == 0) {
    *f.epoch = e;
    st->pcg_it = it;
    st->pcg_rho = rho;
    st->pcg_beta = beta;
    st->pcg_Q0 = Q0;
    st->pcg_iters_last = iters_last;
    st->pcg_break = brk;
    if (fail || *((volatile int *)f.abort_flag)) st->lin_fail = 1;
    st->pcg_done = 1;
  }
}

// rows owned by `rank` (camera blocks wb = rank, rank + N, ...) flagged in the global sorted row order
__global__ void __launch_bounds__(BA_THREADS)
k_dist_flag_rows(int n_cam, const int32_t *__restrict__ row_order, int rank, int n_ranks, int32_t *__restrict__ flag) {
  const int i = blockIdx.x * BA_THREADS + threadIdx.x;
  if (i >= n_cam) return;
  flag[i] = ((row_order[i] >> 5) % n_ranks) == rank ? 1 : 0;
}
__global__ void __launch_bounds__(BA_THREADS)
k_dist_pick_rows(int n_cam, const int32_t *__restrict__ row_order, const int32_t *__restrict__ flag,
                 const int32_t *__restrict__ pos, int32_t *__restrict__ my_rows) {
  const int i = blockIdx.x * BA_THREADS + threadIdx.x;
  if (i >= n_cam) return;
  if (flag[i]) my_rows[pos[i]] = row_order[i];
}
