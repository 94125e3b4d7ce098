// Micro-benchmark (development aid, not product): one-way latency of a flag-in-data 16-byte store
// over NVLink peer memory, and of "data stores + fence.sys + flag", between two GPUs of one process.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_pingpong nvlink_pingpong.cu && ./nvlink_pingpong
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct __align__(16) Slot { double v; unsigned long long f; };

__device__ __forceinline__ void st_slot(Slot *p, double v, unsigned long long f) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(f) : "memory");
}
__device__ __forceinline__ bool ld_slot(const Slot *p, unsigned long long want, double &v) {
  unsigned long long a, b;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  v = __longlong_as_double(a);
  return b == want;
}

// role 0 sends epoch e to the peer slot then waits for the echo in its own slot; role 1 echoes
__global__ void pingpong_ll(int role, Slot *mine, Slot *peer, int iters, unsigned long long *ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  double v;
  for (int e = 1; e <= iters; ++e) {
    if (role == 0) {
      st_slot(peer, 1.0 * e, e);
      while (!ld_slot(mine, e, v)) {}
    } else {
      while (!ld_slot(mine, e, v)) {}
      st_slot(peer, v, e);
    }
  }
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  *ns = t1 - t0;
}
// same, but "payload store(s) + __threadfence_system + separate flag"
__global__ void pingpong_fence(int role, double *mine_d, double *peer_d, unsigned long long *mine_f, unsigned long long *peer_f,
                               int iters, int payload, unsigned long long *ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int e = 1; e <= iters; ++e) {
    if (role == 0) {
      for (int k = 0; k < payload; ++k) peer_d[k] = e + k;
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_f), "l"((unsigned long long)e) : "memory");
      unsigned long long f;
      do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine_f) : "memory"); } while (f < (unsigned long long)e);
    } else {
      unsigned long long f;
      do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine_f) : "memory"); } while (f < (unsigned long long)e);
      for (int k = 0; k < payload; ++k) peer_d[k] = mine_d[k];
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_f), "l"((unsigned long long)e) : "memory");
    }
  }
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  *ns = t1 - t0;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0->1: %d\n", can);
  void *buf[2]; unsigned long long *ns[2]; cudaStream_t st[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&buf[d], 1 << 16));
    CK(cudaMemset(buf[d], 0, 1 << 16));
    CK(cudaMallocHost(&ns[d], 8));
    CK(cudaStreamCreate(&st[d]));
  }
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(buf[d], 0, 1 << 16)); CK(cudaDeviceSynchronize()); }
    for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d));
      pingpong_ll<<<1, 1, 0, st[d]>>>(d, (Slot *)buf[d], (Slot *)buf[1 - d], iters, ns[d]);
    }
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
    printf("flag-in-data 16 B store: round trip %.3f us (one way %.3f us)\n", *ns[0] / 1e3 / iters, *ns[0] / 2e3 / iters);
  }
  for (int payload : {1, 8, 64}) {
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(buf[d], 0, 1 << 16)); CK(cudaDeviceSynchronize()); }
    for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d));
      double *md = (double *)buf[d] + 64, *pd = (double *)buf[1 - d] + 64;
      pingpong_fence<<<1, 1, 0, st[d]>>>(d, md, pd, (unsigned long long *)buf[d], (unsigned long long *)buf[1 - d], iters, payload, ns[d]);
    }
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
    printf("payload %2d doubles + fence.sys + flag: round trip %.3f us (one way %.3f us)\n", payload, *ns[0] / 1e3 / iters, *ns[0] / 2e3 / iters);
  }
  return 0;
}
