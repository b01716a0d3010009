// Micro-benchmark: latency of a 2-GPU scalar exchange through peer-mapped mailboxes.
// Single process, 2 devices, peer access enabled.  nvcc -arch=sm_100a -O3 -o xchg_bench xchg_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Slot { unsigned long long a, b; };
__device__ __forceinline__ void st_rel(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// one exchange, protocol V: 1 = value, fence, release flag / acquire poll; 2 = LL (two 8-byte packets); 3 = fence.sys + LL
template <int V>
__device__ __forceinline__ double xchg(Slot *mine, Slot *peer, double v, unsigned seq) {
  const int par = seq & 1;
  if (V == 1) {
    st_rlx(&peer[par].a, (unsigned long long)__double_as_longlong(v));
    __threadfence_system();
    st_rel(&peer[par].b, seq);
    while (ld_acq(&mine[par].b) != seq) {}
    return __longlong_as_double((long long)ld_rlx(&mine[par].a));
  } else {
    if (V == 3) __threadfence_system();
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_rlx(&peer[par].a, ((unsigned long long)seq << 32) | (bits & 0xffffffffull));
    st_rlx(&peer[par].b, ((unsigned long long)seq << 32) | (bits >> 32));
    unsigned long long a, b;
    do { a = ld_rlx(&mine[par].a); } while ((unsigned)(a >> 32) != seq);
    do { b = ld_rlx(&mine[par].b); } while ((unsigned)(b >> 32) != seq);
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
  }
}

template <int V>
__global__ void k_loop(Slot *mine, Slot *peer, unsigned seq0, int n, double *out) {
  double acc = 0;
  for (int i = 0; i < n; i++) acc += xchg<V>(mine, peer, 1.0 + i, seq0 + i + 1);
  *out = acc;
}
// one exchange per kernel; seq kept in device memory; optional dummy peer "halo" stores by many blocks first
template <int V>
__global__ void k_one(Slot *mine, Slot *peer, unsigned *seq, double *out, double *peer_halo, int nh, unsigned *ticket) {
  if (peer_halo) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += gridDim.x * blockDim.x) peer_halo[i] = (double)i;
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    __threadfence();
    *ticket = 0;
    const unsigned s = *seq + 1; *seq = s;
    *out = xchg<V>(mine, peer, 2.0, s);
  }
}
__global__ void k_empty(unsigned *ticket) {
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); last = atomicAdd(ticket, 1u) == gridDim.x - 1; if (last) *ticket = 0; }
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  Slot *mail[2]; double *out[2], *halo[2]; unsigned *seq[2], *ticket[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
  const int NH = 8192;
  for (int d = 0; d < 2; d++) {
    CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&mail[d], 2 * sizeof(Slot))); CK(cudaMemset(mail[d], 0, 2 * sizeof(Slot)));
    CK(cudaMalloc(&out[d], 8)); CK(cudaMalloc(&halo[d], NH * 8)); CK(cudaMalloc(&seq[d], 4)); CK(cudaMalloc(&ticket[d], 4));
    CK(cudaMemset(seq[d], 0, 4)); CK(cudaMemset(ticket[d], 0, 4));
    CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
  }
  unsigned seq0 = 0;
  const int N = 5000;
  auto run_loop = [&](int V) -> int {
    for (int rep = 0; rep < 2; rep++) {
      for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d]));
        if (V == 1) k_loop<1><<<1, 1, 0, st[d]>>>(mail[d], mail[1 - d], seq0, N, out[d]);
        if (V == 2) k_loop<2><<<1, 1, 0, st[d]>>>(mail[d], mail[1 - d], seq0, N, out[d]);
        if (V == 3) k_loop<3><<<1, 1, 0, st[d]>>>(mail[d], mail[1 - d], seq0, N, out[d]);
        CK(cudaEventRecord(e1[d], st[d]));
      }
      seq0 += N;
      float ms[2];
      for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); CK(cudaEventElapsedTime(&ms[d], e0[d], e1[d])); }
      if (rep) printf("in-kernel loop  V%d: %.2f us / exchange (gpu0 %.2f, gpu1 %.2f)\n", V, 1e3 * (ms[0] > ms[1] ? ms[0] : ms[1]) / N, 1e3 * ms[0] / N, 1e3 * ms[1] / N);
    }
    return 0;
  };
  if (run_loop(1) || run_loop(2) || run_loop(3)) return 1;
  // device seq counters continue from seq0
  for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaMemcpy(seq[d], &seq0, 4, cudaMemcpyHostToDevice)); }
  // kernel-per-exchange through graphs
  const int K = 200;
  auto run_graph = [&](int V, int grid, bool with_halo, const char *tag) -> int {
    cudaGraphExec_t ex[2];
    for (int d = 0; d < 2; d++) {
      CK(cudaSetDevice(d));
      cudaGraph_t g;
      CK(cudaStreamBeginCapture(st[d], cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < K; i++) {
        double *ph = with_halo ? halo[1 - d] : nullptr;
        if (V == 0) k_empty<<<grid, 256, 0, st[d]>>>(ticket[d]);
        if (V == 1) k_one<1><<<grid, 256, 0, st[d]>>>(mail[d], mail[1 - d], seq[d], out[d], ph, NH, ticket[d]);
        if (V == 2) k_one<2><<<grid, 256, 0, st[d]>>>(mail[d], mail[1 - d], seq[d], out[d], ph, NH, ticket[d]);
        if (V == 3) k_one<3><<<grid, 256, 0, st[d]>>>(mail[d], mail[1 - d], seq[d], out[d], ph, NH, ticket[d]);
      }
      CK(cudaStreamEndCapture(st[d], &g));
      CK(cudaGraphInstantiate(&ex[d], g, 0));
      CK(cudaGraphDestroy(g));
    }
    for (int rep = 0; rep < 3; rep++) {
      for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); CK(cudaGraphLaunch(ex[d], st[d])); CK(cudaEventRecord(e1[d], st[d])); }
      float ms[2];
      for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); CK(cudaEventElapsedTime(&ms[d], e0[d], e1[d])); }
      if (rep == 2) printf("graph, 1 kernel per exchange  %-28s grid=%4d: %.2f us / kernel\n", tag, grid, 1e3 * (ms[0] > ms[1] ? ms[0] : ms[1]) / K);
    }
    for (int d = 0; d < 2; d++) cudaGraphExecDestroy(ex[d]);
    return 0;
  };
  for (int grid : {1, 444, 2048}) {
    if (run_graph(0, grid, false, "empty (ticket only)")) return 1;
    if (run_graph(1, grid, false, "V1 fence+release")) return 1;
    if (run_graph(2, grid, false, "V2 LL")) return 1;
    if (run_graph(3, grid, false, "V3 fence.sys + LL")) return 1;
    if (run_graph(3, grid, true, "V3 + 64 KB peer halo stores")) return 1;
    if (run_graph(1, grid, true, "V1 + 64 KB peer halo stores")) return 1;
  }
  return 0;
}
