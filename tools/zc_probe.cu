// Zero-copy (SM-driven) PCIe transfers vs the copy engines: kernels that read / write page-locked host memory directly.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/zc_probe tools/zc_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_cv(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <int UNROLL, bool CV>
__global__ void __launch_bounds__(256) copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = CV ? ld_cv(src + i + u * stride) : src[i + u * stride];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n16; i += stride) dst[i] = CV ? ld_cv(src + i) : src[i];
}


// ---- the same transfers issued by the TMA unit: one thread per CTA moves 16 KiB chunks global -> shared -> global with
//      cp.async.bulk (S stages in flight); either side may be mapped host memory
constexpr int TMA_S = 4, TMA_CH = 16384;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) tma_copy_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long mbar[TMA_S];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < TMA_S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const size_t nchunks = (bytes + TMA_CH - 1) / TMA_CH;
  const size_t first = blockIdx.x, step = gridDim.x;
  const size_t n_my = first < nchunks ? (nchunks - first + step - 1) / step : 0;
  auto load = [&](size_t k) {
    const int s = (int)(k % TMA_S);
    const size_t c = first + k * step;
    const uint32_t n = (uint32_t)((c + 1) * TMA_CH <= bytes ? TMA_CH : bytes - c * TMA_CH);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[s])), "r"(n) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sm + (size_t)s * TMA_CH)), "l"(src + c * TMA_CH), "r"(n), "r"(smem_u32(&mbar[s])) : "memory");
  };
  for (size_t k = 0; k < n_my && k < (size_t)TMA_S; ++k) load(k);
  uint32_t phase = 0u;   // bit s = parity to wait for on stage s
  for (size_t k = 0; k < n_my; ++k) {
    const int s = (int)(k % TMA_S);
    const size_t c = first + k * step;
    const uint32_t n = (uint32_t)((c + 1) * TMA_CH <= bytes ? TMA_CH : bytes - c * TMA_CH);
    uint32_t done = 0u;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&mbar[s])), "r"((phase >> s) & 1u) : "memory");
    phase ^= 1u << s;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * TMA_CH), "r"(smem_u32(sm + (size_t)s * TMA_CH)), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (k + TMA_S < n_my) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read its stage: reuse it
      load(k + TMA_S);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const size_t bytes = (size_t)(5.25 * (1 << 20)), n16 = bytes / 16;
  uint4 *h_in, *h_out, *d_a, *d_b;
  CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d_a, bytes));
  CK(cudaMalloc(&d_b, bytes));
  memset(h_in, 1, bytes);
  CK(cudaMemset(d_b, 2, bytes));
  cudaStream_t s[8];
  for (int i = 0; i < 8; ++i) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
  cudaEvent_t e0, e1, ef, ej[8];
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventCreateWithFlags(&ef, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&ej[i], cudaEventDisableTiming));
  auto timed = [&](const char* name, auto fn) {
    for (int i = 0; i < 5; ++i) fn();
    CK(cudaDeviceSynchronize());
    const int reps = 50;
    CK(cudaEventRecord(e0, s[0]));
    for (int i = 0; i < reps; ++i) fn();
    CK(cudaEventRecord(e1, s[0]));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-58s %8.1f us  (%5.1f GB/s per direction)\n", name, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
  };
  timed("DMA H2D", [&] { CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s[0])); });
  timed("DMA D2H", [&] { CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s[0])); });
  for (int grid : {16, 32, 64, 148, 296}) {
    char nm[128];
    snprintf(nm, sizeof nm, "kernel H2D (host read, ld.cv, unroll 4), %d CTAs", grid);
    timed(nm, [&] { copy_kernel<4, true><<<grid, 256, 0, s[0]>>>(d_a, h_in, n16); });
    snprintf(nm, sizeof nm, "kernel H2D (host read, plain ld, unroll 4), %d CTAs", grid);
    timed(nm, [&] { copy_kernel<4, false><<<grid, 256, 0, s[0]>>>(d_a, h_in, n16); });
    snprintf(nm, sizeof nm, "kernel H2D (host read, ld.cv, unroll 8), %d CTAs", grid);
    timed(nm, [&] { copy_kernel<8, true><<<grid, 256, 0, s[0]>>>(d_a, h_in, n16); });
    snprintf(nm, sizeof nm, "kernel D2H (host write, unroll 4), %d CTAs", grid);
    timed(nm, [&] { copy_kernel<4, false><<<grid, 256, 0, s[0]>>>(h_out, d_b, n16); });
  }
  CK(cudaFuncSetAttribute(tma_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_S * TMA_CH));
  for (int grid : {8, 16, 32, 64}) {
    char nm[128];
    snprintf(nm, sizeof nm, "TMA bulk H2D (host -> smem -> device), %d CTAs", grid);
    timed(nm, [&] { tma_copy_kernel<<<grid, 32, TMA_S * TMA_CH, s[0]>>>((unsigned char*)d_a, (const unsigned char*)h_in, bytes); });
    snprintf(nm, sizeof nm, "TMA bulk D2H (device -> smem -> host), %d CTAs", grid);
    timed(nm, [&] { tma_copy_kernel<<<grid, 32, TMA_S * TMA_CH, s[0]>>>((unsigned char*)h_out, (const unsigned char*)d_b, bytes); });
  }
  auto fork = [&](int k) { CK(cudaEventRecord(ef, s[0])); for (int i = 1; i <= k; ++i) CK(cudaStreamWaitEvent(s[i], ef, 0)); };
  auto join = [&](int k) { for (int i = 1; i <= k; ++i) { CK(cudaEventRecord(ej[i], s[i])); CK(cudaStreamWaitEvent(s[0], ej[i], 0)); } };
  timed("kernel H2D + kernel D2H concurrently (64 CTAs each)", [&] {
    fork(2);
    copy_kernel<4, true><<<64, 256, 0, s[1]>>>(d_a, h_in, n16);
    copy_kernel<4, false><<<64, 256, 0, s[2]>>>(h_out, d_b, n16);
    join(2);
  });
  timed("TMA bulk H2D + TMA bulk D2H concurrently (32 CTAs each)", [&] {
    fork(2);
    tma_copy_kernel<<<32, 32, TMA_S * TMA_CH, s[1]>>>((unsigned char*)d_a, (const unsigned char*)h_in, bytes);
    tma_copy_kernel<<<32, 32, TMA_S * TMA_CH, s[2]>>>((unsigned char*)h_out, (const unsigned char*)d_b, bytes);
    join(2);
  });
  timed("TMA bulk H2D + kernel D2H concurrently", [&] {
    fork(2);
    tma_copy_kernel<<<32, 32, TMA_S * TMA_CH, s[1]>>>((unsigned char*)d_a, (const unsigned char*)h_in, bytes);
    copy_kernel<4, false><<<64, 256, 0, s[2]>>>(h_out, d_b, n16);
    join(2);
  });
  timed("kernel H2D + TMA bulk D2H concurrently", [&] {
    fork(2);
    copy_kernel<4, true><<<64, 256, 0, s[1]>>>(d_a, h_in, n16);
    tma_copy_kernel<<<32, 32, TMA_S * TMA_CH, s[2]>>>((unsigned char*)h_out, (const unsigned char*)d_b, bytes);
    join(2);
  });
  timed("DMA H2D + DMA D2H concurrently", [&] {
    fork(2);
    CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s[1]));
    CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s[2]));
    join(2);
  });
  timed("DMA H2D + kernel D2H concurrently", [&] {
    fork(2);
    CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s[1]));
    copy_kernel<4, false><<<64, 256, 0, s[2]>>>(h_out, d_b, n16);
    join(2);
  });
  for (int k : {8, 16}) {
    char nm[128];
    const size_t c16 = n16 / k;
    snprintf(nm, sizeof nm, "DMA H2D in %d chunks, one stream", k);
    timed(nm, [&] { for (int i = 0; i < k; ++i) CK(cudaMemcpyAsync(d_a + i * c16, h_in + i * c16, c16 * 16, cudaMemcpyHostToDevice, s[0])); });
    snprintf(nm, sizeof nm, "DMA H2D in %d chunks, round-robin over 4 streams", k);
    timed(nm, [&] {
      fork(4);
      for (int i = 0; i < k; ++i) CK(cudaMemcpyAsync(d_a + i * c16, h_in + i * c16, c16 * 16, cudaMemcpyHostToDevice, s[1 + i % 4]));
      join(4);
    });
    snprintf(nm, sizeof nm, "kernel H2D in %d chunks (64 CTAs), one stream", k);
    timed(nm, [&] { for (int i = 0; i < k; ++i) copy_kernel<4, true><<<64, 256, 0, s[0]>>>(d_a + i * c16, h_in + i * c16, c16); });
    snprintf(nm, sizeof nm, "kernel D2H in %d chunks (64 CTAs), one stream", k);
    timed(nm, [&] { for (int i = 0; i < k; ++i) copy_kernel<4, false><<<64, 256, 0, s[0]>>>(h_out + i * c16, d_b + i * c16, c16); });
  }
  // correctness of the last transfers
  CK(cudaDeviceSynchronize());
  std::vector<unsigned char> chk(bytes);
  CK(cudaMemcpy(chk.data(), d_a, bytes, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < bytes; ++i) bad += chk[i] != 1;
  const unsigned char* ho = (const unsigned char*)h_out;
  for (size_t i = 0; i < bytes; ++i) bad += ho[i] != 2;
  printf("mismatching bytes: %zu\n", bad);
  return 0;
}
