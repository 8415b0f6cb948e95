// tools/p2p_probe.cu — NVLink peer-access characteristics that size the sharded swap pass (single process, 2 GPUs).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/p2p_probe.cu -o tools/_build/p2p_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__global__ void gather_rows(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ idx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4* s = reinterpret_cast<const double4*>(src + (size_t)idx[i] * 8);
  double4 a = s[0], b = s[1];
  double4* d = reinterpret_cast<double4*>(dst + (size_t)i * 8);
  d[0] = a; d[1] = b;
}
__global__ void chase(const int* __restrict__ next, int* out, int steps) {
  int p = 0;
  for (int i = 0; i < steps; ++i) p = next[p];
  *out = p;
}
__global__ void store_rows(double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<double2*>(dst)[i] = make_double2(1.0, 2.0);
}
__global__ void flag_set(volatile unsigned long long* f, unsigned long long v) { *f = v; }
__global__ void flag_wait(volatile unsigned long long* f, unsigned long long v) { while (*f < v) {} }

static float time_kernel(cudaStream_t s, int reps, void (*launch)(cudaStream_t, void*), void* ctx) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(s, ctx); CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(a, s));
  for (int i = 0; i < reps; ++i) launch(s, ctx);
  CK(cudaEventRecord(b, s)); CK(cudaStreamSynchronize(s));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

struct G { const double* src; double* dst; const int* idx; int n; };
static void launch_gather(cudaStream_t s, void* c) { G* g = (G*)c; gather_rows<<<(g->n + 255) / 256, 256, 0, s>>>(g->src, g->dst, g->idx, g->n); }
struct C2 { const int* next; int* out; int steps; };
static void launch_chase(cudaStream_t s, void* c) { C2* g = (C2*)c; chase<<<1, 1, 0, s>>>(g->next, g->out, g->steps); }
struct S { double* dst; int n; };
static void launch_store(cudaStream_t s, void* c) { S* g = (S*)c; store_rows<<<(g->n + 255) / 256, 256, 0, s>>>(g->dst, g->n); }

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { std::printf("need 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  const int NR = 65536;
  double *loc, *rem, *dst; int *idx, *nl, *nr, *out;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, (size_t)NR * 64)); CK(cudaMemset(rem, 0, (size_t)NR * 64));
  int* h = (int*)std::malloc(NR * 4);
  for (int i = 0; i < NR; ++i) h[i] = (int)(((long long)i * 40503) % NR);
  CK(cudaMalloc(&nr, NR * 4)); CK(cudaMemcpy(nr, h, NR * 4, cudaMemcpyHostToDevice));
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&loc, (size_t)NR * 64)); CK(cudaMalloc(&dst, (size_t)NR * 64)); CK(cudaMalloc(&idx, NR * 4)); CK(cudaMalloc(&nl, NR * 4)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(loc, 0, (size_t)NR * 64));
  CK(cudaMemcpy(idx, h, NR * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(nl, h, NR * 4, cudaMemcpyHostToDevice));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  C2 cl{nl, out, 2000}, cr{nr, out, 2000};
  std::printf("dependent-load latency  local: %.0f ns   peer (NVLink): %.0f ns\n", time_kernel(s, 3, launch_chase, &cl) * 1000 / 2000,
              time_kernel(s, 3, launch_chase, &cr) * 1000 / 2000);
  for (int n : {1024, 4096, 16384, 65536}) {
    G gl{loc, dst, idx, n}, gr{rem, dst, idx, n};
    std::printf("gather %6d random 64-B rows: local %.2f us   peer %.2f us (%.1f GB/s)\n", n, time_kernel(s, 20, launch_gather, &gl),
                time_kernel(s, 20, launch_gather, &gr), n * 64.0 / time_kernel(s, 20, launch_gather, &gr) / 1e3);
  }
  for (int n : {32768, 262144}) {
    S sl{loc, n}, sr{rem, n};
    std::printf("coalesced 16-B stores x %6d: local %.2f us   peer %.2f us (%.1f GB/s)\n", n, time_kernel(s, 20, launch_store, &sl),
                time_kernel(s, 20, launch_store, &sr), n * 16.0 / time_kernel(s, 20, launch_store, &sr) / 1e3);
  }
  return 0;
}
