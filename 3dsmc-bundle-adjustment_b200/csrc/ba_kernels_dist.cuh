// ba_kernels_dist.cuh -- the persistent PCG of the block-sparse solver, ROW-SHARDED over the
// GPUs of one node, with the exchange fused into the kernel over NVLink peer memory.
//
// One process per GPU; S, b, M^-1 are complete and identical on every rank (S is all-reduced once
// per LM iteration).  Each rank owns the camera blocks wb = rank, rank + N, ... (32 consecutive
// cameras each, the granularity of the single-GPU partial sums) and runs the same persistent loop
// as k_pcg_sparse_persistent on ITS rows only.  What the others need is written straight into
// their memory (cudaIpc-mapped exchange buffers, plain st.global over NVLink), by the thread that
// produced it:
//     phase I   q_r, p.q of the row           -> row_pq[row]         on every rank
//     phase II  z_c, partials of r.z, x.(b+r) -> z[c], wb_rho/wb_Q   on every rank
//               (x_c as well on residual-reset iterations and at the end)
// followed by a cross-GPU barrier: the local grid barrier, then the last CTA to arrive publishes
// an epoch number into every peer's flag slot (st.release.sys) and every CTA waits for all slots.
// No NCCL call, no host round trip inside the solve: two NVLink barriers per PCG iteration.
// Every rank sums the same complete arrays in the same order, so the PCG scalars -- and with them
// every iterate -- are BIT-IDENTICAL to the single-GPU solve and identical across ranks.
// Spin waits carry a time-out (a rank that died must not hang the others' GPUs).
//
// STATUS (round 1, 2 x B200, cfg 5): correct (tests/test_gpu_multi.py) but NOT faster -- per PCG iteration
// product 12.2 us (18.7 on one GPU) + 2 x 12.5 us NVLink barrier (fence.sys drain + flag round trip) against
// 2 x 1.7 us for the local grid barrier: 48 us vs 29 us.  Selected only by persistent_pcg = 2; the default
// multi-GPU mode runs the PCG replicated.  Next step: flag-in-data (LL-style) slots instead of fence + flag.
#pragma once
#include "ba_kernels_sparse.cuh"

#define BA_MAX_RANKS 8
struct PcgFan {
  int n_ranks, rank;
  double *z[BA_MAX_RANKS], *x[BA_MAX_RANKS], *row_pq[BA_MAX_RANKS], *wb_rho[BA_MAX_RANKS], *wb_Q[BA_MAX_RANKS];
  unsigned long long *flags[BA_MAX_RANKS];  // flags[k]: rank k's slots; this rank writes flags[k][rank]
  unsigned long long *epoch;                // own running epoch counter (survives launches)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// local grid barrier + cross-GPU epoch exchange.  Returns false on time-out.
__device__ __forceinline__ bool cross_barrier(const PcgFan &f, unsigned int *bar, unsigned int &local_epoch,
                                              unsigned long long &epoch) {
  __shared__ int ok_s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int ok = 1;
    local_epoch += gridDim.x;
    epoch += 1;
    __threadfence_system();  // this CTA's peer stores are visible system-wide before it arrives
    const unsigned int old = atomicAdd(bar, 1u);
    if (old == local_epoch - 1u) {  // last local CTA: tell every rank (this one included)
      __threadfence_system();
      for (int k = 0; k < f.n_ranks; ++k) st_release_sys(f.flags[k] + f.rank, epoch);
    }
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int k = 0; k < f.n_ranks && ok; ++k) {
      while (ld_acquire_sys(f.flags[f.rank] + k) < epoch) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {  // 4 s: a peer is gone
          ok = 0;
          break;
        }
      }
    }
    __threadfence_system();
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}

// rows order[gw], order[gw + nw], ... (< n_rows) of out = S v + dsq .* v, v = za (+ beta pb);
// MODE 0: p_new for the row, p.q of the row to every rank.  MODE 1 (residual reset): out only.
template <int MODE>
__device__ __forceinline__ void dist_rows(const PcgFan &f, int n_rows, int gw, int nw, int lane,
                                          const int32_t *__restrict__ order, const int32_t *__restrict__ ent_ptr,
                                          const int2 *__restrict__ ent, const double *__restrict__ S,
                                          const double *__restrict__ dsq, const double *za, const double *pb, double beta,
                                          bool use_pb, double *pnew, double *out) {
  const int2 none = make_int2(0, 0);
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
    for (int e = b0 + lane + 64; e < e0; e += 32) bsr_entry(__ldg(ent + e), S, za, pb, beta, use_pb, acc);
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
    const double pv = use_pb ? zk + beta * pk : zk;
    const double qv = pick6(acc, lane) + dk * pv;
    if (lane < 6) {
      if (MODE == 0) pnew[6 * (size_t)row + lane] = pv;
      out[6 * (size_t)row + lane] = qv;
    }
    if (MODE == 0) {
      const double t = pv * qv;
      double s = __shfl_sync(BA_FULL, t, 0);
#pragma unroll
      for (int k = 1; k < 6; ++k) s += __shfl_sync(BA_FULL, t, k);
      if (lane < f.n_ranks) f.row_pq[lane][row] = s;  // lane k writes rank k's copy
    }
    row = row1; row1 = row2;
    b0 = b1; e0 = e1; en0 = en1; en0b = en1b;
    b1 = b2; e1 = e2;
  }
}

// phase II for the owned warp-block wb (lane = camera); z and the two partials go to every rank,
// x too when push_x (residual-reset iterations need the complete x for the product S x).
template <int RESET>
__device__ __forceinline__ void dist_update(const PcgFan &f, int n_cam, int wb, int lane, double alpha, bool skip_r, bool push_x,
                                            const double *__restrict__ b, const double *__restrict__ Minv, double *r,
                                            const double *pnew, const double *q) {
  const int c = wb * 32 + lane;
  double rz = 0.0, xq = 0.0;
  double *x = f.x[f.rank];
  if (c < n_cam) {
    double xv[6], rv[6], qv[6], bv[6], zv[6];
    load6cg(x + 6 * (size_t)c, xv);
    if (!RESET) {
      double pv[6];
      load6cg(pnew + 6 * (size_t)c, pv);
#pragma unroll
      for (int k = 0; k < 6; ++k) xv[k] = xv[k] + alpha * pv[k];
      store6(x + 6 * (size_t)c, xv);
      if (push_x)
        for (int k = 0; k < f.n_ranks; ++k)
          if (k != f.rank) store6(f.x[k] + 6 * (size_t)c, xv);
    }
    if (!skip_r) {
      load6cg(q + 6 * (size_t)c, qv);
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
      minv_mul(Minv, c, rv, zv);
      for (int k = 0; k < f.n_ranks; ++k) store6(f.z[k] + 6 * (size_t)c, zv);
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
    if (lane < f.n_ranks) {
      f.wb_rho[lane][wb] = rz;
      f.wb_Q[lane][wb] = xq;
    }
  }
}

__global__ void __launch_bounds__(BA_THREADS, 1)
k_pcg_sparse_dist(PcgFan f, int n_cam, int n_my_rows, const int32_t *__restrict__ my_rows, const int32_t *__restrict__ ent_ptr,
                  const int2 *__restrict__ ent, const double *__restrict__ S, const double *__restrict__ dsq,
                  const double *__restrict__ b, const double *__restrict__ Minv, double *r, double *pbuf0, double *pbuf1, double *q,
                  unsigned int *bar, LmOptions lo, LmState *st, int *comm_fail, unsigned long long *prof) {
  if (st->done || st->pcg_done) return;  // identical on every CTA and every rank
  __shared__ double red[2 * BA_WARPS + 4];
  const int tid = threadIdx.x, lane = tid & 31;
  const int gtid = blockIdx.x * BA_THREADS + tid, nthreads = gridDim.x * BA_THREADS;
  const int gw = gtid >> 5, nw = nthreads >> 5;
  const int n_wb = (n_cam + 31) / 32;
  const int N = f.n_ranks, me = f.rank;
  double *z = f.z[me], *x = f.x[me], *row_pq = f.row_pq[me], *part_rho = f.wb_rho[me], *part_Q = f.wb_Q[me];
  int it = st->pcg_it;
  double rho = st->pcg_rho, beta = st->pcg_beta, Q0 = st->pcg_Q0;
  int fail = 0, brk = 0, iters_last = 0, dead = 0;
  unsigned int local_epoch = 0;
  unsigned long long epoch = *f.epoch;  // same value on every CTA: written back only after the last barrier
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
    // ---- phase I: own rows of q = S p + D^2 p (p = z + beta p_old on the fly), p.q per row to every rank;
    //      p itself is kept complete locally (every rank updates all of it)
    dist_rows<0>(f, n_my_rows, gw, nw, lane, my_rows, ent_ptr, ent, S, dsq, z, pold, beta, it > 1, pnew, q);
    for (int i = gtid; i < 6 * n_cam; i += nthreads) {
      const double zv = __ldcg(z + i);
      pnew[i] = it > 1 ? zv + beta * __ldcg(pold + i) : zv;
    }
    DPROF_TICK(0)
    if (!cross_barrier(f, bar, local_epoch, epoch)) {
      dead = 1;
      break;
    }
    DPROF_TICK(1)

    // ---- phase II on the owned camera blocks
    const double pq = block_sum_wide_cg(row_pq, n_cam, red);
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
    DPROF_TICK(2)
    for (int w = gw; me + N * w < n_wb; w += nw)
      dist_update<0>(f, n_cam, me + N * w, lane, alpha, reset, reset, b, Minv, r, pnew, q);
    DPROF_TICK(3)
    if (!cross_barrier(f, bar, local_epoch, epoch)) {
      dead = 1;
      break;
    }
    DPROF_TICK(4)
    if (reset) {
      dist_rows<1>(f, n_my_rows, gw, nw, lane, my_rows, ent_ptr, ent, S, dsq, x, x, 0.0, false, nullptr, q);
      // q of the owned rows is consumed by the owner only: a local grid barrier would do, the cross barrier
      // keeps one code path (reset iterations are 1 in 10)
      if (!cross_barrier(f, bar, local_epoch, epoch)) {
        dead = 1;
        break;
      }
      for (int w = gw; me + N * w < n_wb; w += nw)
        dist_update<1>(f, n_cam, me + N * w, lane, 0.0, false, false, b, Minv, r, pnew, q);
      if (!cross_barrier(f, bar, local_epoch, epoch)) {
        dead = 1;
        break;
      }
    }

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
  // ---- the solution: every rank needs all of x (back-substitution is per point shard)
  if (!dead) {
    for (int w = gw; me + N * w < n_wb; w += nw) {
      const int c = (me + N * w) * 32 + lane;
      if (c < n_cam) {
        double xv[6];
        load6cg(x + 6 * (size_t)c, xv);
        for (int k = 0; k < N; ++k)
          if (k != me) store6(f.x[k] + 6 * (size_t)c, xv);
      }
    }
    if (!cross_barrier(f, bar, local_epoch, epoch)) dead = 1;
  }
  if (blockIdx.x == 0 && tid == 0) {
    *f.epoch = epoch;
    st->pcg_it = it;
    st->pcg_rho = rho;
    st->pcg_beta = beta;
    st->pcg_Q0 = Q0;
    st->pcg_iters_last = iters_last;
    st->pcg_break = brk;
    if (fail || dead) st->lin_fail = 1;
    if (dead) *comm_fail = 1;
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
