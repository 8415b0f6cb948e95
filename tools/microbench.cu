// tools/microbench.cu — per-launch and per-phase timing of the hot-path kernels on the C2 shape.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tools/microbench.cu \
//        -L eryn_b200/lib -leryn_b200_prof -o tools/microbench        (tools/build_microbench.sh)
//
// Prints: launch overhead of empty kernels (plain / cluster of 8) inside a CUDA graph, the per-launch
// time of eb_stretch_step and eb_pt_swap in a graph of N sequential launches, and — with the profiling
// build of the library (-DEB_PHASE_TIMERS) — the SM-cycle marks of CTA (0,0).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "eryn_b200.h"

extern "C" int eb_debug_marks_stretch(long long* out_host);
extern "C" int eb_debug_marks_swap(long long* out_host);

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(1);                                                                   \
    }                                                                                 \
  } while (0)
#define EB(x)                                                              \
  do {                                                                     \
    int r_ = (x);                                                          \
    if (r_) {                                                              \
      std::fprintf(stderr, "%s -> %d: %s\n", #x, r_, eb_last_error());     \
      std::exit(1);                                                        \
    }                                                                      \
  } while (0)

__global__ void empty_kernel(int* p) {
  if (p && threadIdx.x == 9999) *p = 1;
}
__global__ void cluster_sync_kernel(int* p) {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (p && threadIdx.x == 9999) *p = 1;
}

static float time_graph(cudaStream_t s, int n, const std::function<void()>& body, int reps = 5) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  body();  // warm
  CK(cudaStreamSynchronize(s));
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < n; ++i) body();
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a, s));
    CK(cudaGraphLaunch(ge, s));
    CK(cudaEventRecord(b, s));
    CK(cudaStreamSynchronize(s));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGraphExecDestroy(ge));
  CK(cudaGraphDestroy(g));
  return best * 1000.f / n;  // us per launch
}

static void launch_empty(cudaStream_t s, int gx, int gy, int threads, int cluster, bool sync) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(gx, gy, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  if (cluster > 1) {
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  int* np = nullptr;
  if (sync) CK(cudaLaunchKernelEx(&cfg, cluster_sync_kernel, np));
  else CK(cudaLaunchKernelEx(&cfg, empty_kernel, np));
}

int main(int argc, char** argv) {
  const int T = argc > 1 ? std::atoi(argv[1]) : 16, W = argc > 2 ? std::atoi(argv[2]) : 4096, D = argc > 3 ? std::atoi(argv[3]) : 8;
  const int N = 50;
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::printf("shape T=%d W=%d D=%d\n", T, W, D);
  std::printf("empty plain   grid 128x256      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 128, 1, 256, 1, false); }));
  std::printf("empty plain   grid (8,16)x256   : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 1, false); }));
  std::printf("empty cluster8 grid (8,16)x256  : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 8, false); }));
  std::printf("cluster8 + cluster barrier      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 8, true); }));
  std::printf("empty plain   grid 256x128      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 256, 1, 128, 1, false); }));

  // ---- state -------------------------------------------------------------------------------------
  const size_t n = (size_t)T * W;
  std::vector<double> hc(n * D), hb(T), hpr(3 * D), hlk(D + D * D, 0.0);
  srand(1);
  for (auto& v : hc) v = 6.0 * rand() / RAND_MAX - 3.0;
  for (int t = 0; t < T; ++t) hb[t] = std::pow(1.5, -t);
  for (int d = 0; d < D; ++d) { hpr[d] = -10; hpr[D + d] = 10; hpr[2 * D + d] = std::log(1.0 / 20.0); hlk[D + d * D + d] = 1.0; }
  double *coords, *logl, *logp, *betas, *pr, *lk;
  uint8_t* acc; uint32_t* cnt; eb_ctrl* ctrl;
  CK(cudaMalloc(&coords, n * D * 8)); CK(cudaMalloc(&logl, n * 8)); CK(cudaMalloc(&logp, n * 8)); CK(cudaMalloc(&betas, T * 8));
  CK(cudaMalloc(&pr, 3 * D * 8)); CK(cudaMalloc(&lk, (D + D * D) * 8)); CK(cudaMalloc(&acc, n)); CK(cudaMalloc(&cnt, n * 4));
  CK(cudaMalloc(&ctrl, sizeof(eb_ctrl)));
  CK(cudaMemcpy(coords, hc.data(), n * D * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(betas, hb.data(), T * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pr, hpr.data(), 3 * D * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(lk, hlk.data(), (D + D * D) * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(ctrl, 0, sizeof(eb_ctrl))); CK(cudaMemset(cnt, 0, n * 4));
  eb_state st; std::memset(&st, 0, sizeof(st));
  st.ntemps = T; st.nwalkers = W; st.nleaves = 1; st.ndim = D; st.coords = coords; st.logl = logl; st.logp = logp; st.betas = betas;
  eb_prior prior{pr, pr + D, pr + 2 * D};
  eb_like like{EB_LIKE_GAUSSIAN, 0, D + D * D, 0, lk};
  EB(eb_eval_state(&st, &prior, &like, s));
  eb_stretch_rng sr; std::memset(&sr, 0, sizeof(sr));
  sr.mode = EB_RNG_PHILOX; sr.randomize_split = 1; sr.seed = 7; sr.iter_dev = &ctrl->iter;
  eb_swap_rng wr; std::memset(&wr, 0, sizeof(wr));
  wr.mode = EB_RNG_PHILOX; wr.permute = 1; wr.seed = 7; wr.iter_dev = &ctrl->iter;
  eb_adapt ad{1, -1, 10000.0, 100.0};
  eb_gauss_rng gr; std::memset(&gr, 0, sizeof(gr));
  gr.mode = EB_RNG_PHILOX; gr.cov_kind = 0; gr.scale = 0.1; gr.seed = 7; gr.iter_dev = &ctrl->iter;

  std::printf("eb_stretch_step (both halves)   : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s)); }));
  eb_stretch_rng sr0 = sr; sr0.randomize_split = 0;
  std::printf("eb_stretch_step no-randomize    : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr0, acc, cnt, s)); }));
  std::printf("eb_gaussian_step                : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_gaussian_step(&st, &prior, &like, &gr, acc, cnt, s)); }));
  std::printf("eb_pt_swap                      : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_pt_swap(&st, &wr, &ad, ctrl, s)); }));
  std::printf("eb_eval_state                   : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_eval_state(&st, &prior, &like, s)); }));
  std::printf("iteration (stretch + swap)      : %.2f us\n", time_graph(s, N, [&] {
    EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s));
    EB(eb_pt_swap(&st, &wr, &ad, ctrl, s));
  }));

  // ---- phase marks (profiling build) ---------------------------------------------------------------
  long long m[64];
  EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s));
  EB(eb_pt_swap(&st, &wr, &ad, ctrl, s));
  CK(cudaStreamSynchronize(s));
  if (eb_debug_marks_stretch(m) == 0) {
    const char* sn[] = {"start", "rng-init", "prepare x2", "stage params", "finish half 0", "(loop)", "cluster barrier", "finish half 1"};
    std::printf("stretch CTA(0,0) cycles: ");
    for (int i = 1; i < 8; ++i) std::printf("%s=%lld ", sn[i], m[i] - m[i - 1]);
    std::printf(" total=%lld\n", m[7] - m[0]);
    EB(eb_debug_marks_swap(m));
    const char* wn[] = {"start", "phase0 keys", "phase1 pos/logl/logu", "phase2 cascade", "counts", "phase3 rows", "tail-last"};
    std::printf("swap    CTA 0 cycles   : ");
    for (int i = 17; i <= 22; ++i) std::printf("%s=%lld ", wn[i - 16], m[i] - m[i - 1]);
    std::printf(" total=%lld\n", m[22] - m[16]);
  }
  return 0;
}
