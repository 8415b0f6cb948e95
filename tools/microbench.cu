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

#ifndef NO_MARKS
extern "C" int eb_debug_marks_stretch(long long* out_host);
extern "C" int eb_debug_marks_swap(long long* out_host);
extern "C" int eb_debug_marks_swap_global(unsigned long long* mn, unsigned long long* mx, int reset);
extern "C" int eb_debug_marks_stretch_global(unsigned long long* mn, unsigned long long* mx, int reset);
extern "C" int eb_debug_marks_swap_cta(unsigned long long* out8x1024);
#endif

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
  const int LK = argc > 4 ? std::atoi(argv[4]) : 0;   // 0 Gaussian, 1 Rosenbrock, 2 mixture of 4 Gaussians
  const int N = 50;
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::printf("shape T=%d W=%d D=%d like=%d\n", T, W, D, LK);
  std::printf("empty plain   grid 128x256      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 128, 1, 256, 1, false); }));
  std::printf("empty plain   grid (8,16)x256   : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 1, false); }));
  std::printf("empty cluster8 grid (8,16)x256  : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 8, false); }));
  std::printf("cluster8 + cluster barrier      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 8, 16, 256, 8, true); }));
  std::printf("empty plain   grid 256x128      : %.2f us/launch\n", time_graph(s, N, [&] { launch_empty(s, 256, 1, 128, 1, false); }));

  // ---- state -------------------------------------------------------------------------------------
  const size_t n = (size_t)T * W;
  const int KM = 4;
  const int nlk = LK == 0 ? D + D * D : LK == 1 ? 0 : KM * (2 + D);
  std::vector<double> hc(n * D), hb(T), hpr(3 * D), hlk(nlk + 1, 0.0);
  srand(1);
  for (auto& v : hc) v = 6.0 * rand() / RAND_MAX - 3.0;
  for (int t = 0; t < T; ++t) hb[t] = std::pow(1.5, -t);
  for (int d = 0; d < D; ++d) { hpr[d] = -10; hpr[D + d] = 10; hpr[2 * D + d] = std::log(1.0 / 20.0); if (LK == 0) hlk[D + d * D + d] = 1.0; }
  if (LK == 2) for (int k = 0; k < KM; ++k) { hlk[k] = -1.0 - 0.1 * k; hlk[KM + k] = 0.5 / (1.0 + 0.2 * k); for (int d = 0; d < D; ++d) hlk[2 * KM + k * D + d] = 4.0 * rand() / RAND_MAX - 2.0; }
  double *coords, *logl, *logp, *betas, *pr, *lk;
  uint8_t* acc; uint32_t* cnt; eb_ctrl* ctrl;
  CK(cudaMalloc(&coords, n * D * 8)); CK(cudaMalloc(&logl, n * 8)); CK(cudaMalloc(&logp, n * 8)); CK(cudaMalloc(&betas, T * 8));
  CK(cudaMalloc(&pr, 3 * D * 8)); CK(cudaMalloc(&lk, (nlk + 1) * 8)); CK(cudaMalloc(&acc, n)); CK(cudaMalloc(&cnt, n * 4));
  CK(cudaMalloc(&ctrl, sizeof(eb_ctrl)));
  CK(cudaMemcpy(coords, hc.data(), n * D * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(betas, hb.data(), T * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pr, hpr.data(), 3 * D * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(lk, hlk.data(), (nlk + 1) * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(ctrl, 0, sizeof(eb_ctrl))); CK(cudaMemset(cnt, 0, n * 4));
  eb_state st; std::memset(&st, 0, sizeof(st));
  st.ntemps = T; st.nwalkers = W; st.nleaves = 1; st.ndim = D; st.coords = coords; st.logl = logl; st.logp = logp; st.betas = betas;
  eb_prior prior{pr, pr + D, pr + 2 * D, nullptr};
  eb_like like{LK, LK == 2 ? KM : 0, nlk, 0, lk};
  EB(eb_eval_state(&st, &prior, &like, s));
  eb_stretch_rng sr; std::memset(&sr, 0, sizeof(sr));
  sr.mode = EB_RNG_PHILOX; sr.randomize_split = 1; sr.seed = 7; sr.iter_dev = &ctrl->iter_next; sr.pdl_chain = 1;
  eb_swap_rng wr; std::memset(&wr, 0, sizeof(wr));
  wr.mode = EB_RNG_PHILOX; wr.permute = 1; wr.seed = 7; wr.iter_dev = &ctrl->iter;
  {   // staging buffers of the shapes whose rows do not move through registers (long ladders / long rows)
    double *rs, *ls;
    CK(cudaMalloc(&rs, (size_t)T * W * D * 8)); CK(cudaMalloc(&ls, (size_t)T * W * 8));
    wr.row_scratch = rs; wr.logp_scratch = ls;
  }
  eb_adapt ad{1, -1, 10000.0, 100.0};
  eb_gauss_rng gr; std::memset(&gr, 0, sizeof(gr));
  gr.mode = EB_RNG_PHILOX; gr.cov_kind = 0; gr.scale = 0.1; gr.seed = 7; gr.iter_dev = &ctrl->iter;

  std::printf("eb_stretch_step (both halves)   : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s)); }));
  std::printf("eb_stretch_step, no count buffer: %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, nullptr, s)); }));
  eb_stretch_rng sr0 = sr; sr0.randomize_split = 0;
  std::printf("eb_stretch_step no-randomize    : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr0, acc, cnt, s)); }));
  std::printf("eb_gaussian_step                : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_gaussian_step(&st, &prior, &like, &gr, acc, cnt, s)); }));
  std::printf("eb_pt_swap                      : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_pt_swap(&st, &wr, &ad, ctrl, s)); }));
  std::printf("eb_eval_state                   : %.2f us/launch\n", time_graph(s, N, [&] { EB(eb_eval_state(&st, &prior, &like, s)); }));
  std::printf("iteration (stretch + swap)      : %.2f us\n", time_graph(s, N, [&] {
    EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s));
    EB(eb_pt_swap(&st, &wr, &ad, ctrl, s));
  }));

  // ---- the sharded pass on ONE GPU (world = 1: publish to self, every row rewritten into the alternate buffers) -------
  {
    double *coords2, *logl2, *logp2, *logl_all, *betas_all;
    unsigned long long* flags;
    CK(cudaMalloc(&coords2, n * D * 8)); CK(cudaMalloc(&logl2, n * 8)); CK(cudaMalloc(&logp2, n * 8));
    CK(cudaMalloc(&logl_all, n * 8)); CK(cudaMalloc(&betas_all, T * 8)); CK(cudaMalloc(&flags, 16 * 8));
    CK(cudaMemset(flags, 0, 16 * 8));
    CK(cudaMemcpy(betas_all, hb.data(), T * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctrl, 0, sizeof(eb_ctrl)));
    eb_state cur = st, alt = st;
    alt.coords = coords2; alt.logl = logl2; alt.logp = logp2;
    cur.betas = betas_all; alt.betas = betas_all;
    eb_swap_rng wr2 = wr;
    auto one = [&](bool with_publish) {
      eb_publish pb; std::memset(&pb, 0, sizeof(pb));
      pb.rank = 0; pb.world = 1; pb.ntemps_total = T; pb.nwalkers = W; pb.temp_begin[0] = 0; pb.temp_begin[1] = T;
      pb.logl_local = cur.logl; pb.logl_all_peer[0] = logl_all; pb.flags_peer[0] = (uint64_t*)flags;
      if (with_publish) EB(eb_publish_logl(&pb, ctrl, s));
      eb_shard sh; std::memset(&sh, 0, sizeof(sh));
      sh.rank = 0; sh.world = 1; sh.ntemps_total = T; sh.temp_begin[0] = 0; sh.temp_begin[1] = T;
      sh.coords_src[0] = cur.coords; sh.logp_src[0] = cur.logp; sh.logl_all = with_publish ? logl_all : cur.logl;
      sh.betas_all = betas_all; sh.flags = with_publish ? (const uint64_t*)flags : nullptr;
      EB(eb_pt_swap_sharded(&sh, &alt, &wr2, &ad, ctrl, s));
      std::swap(cur, alt);
    };
    // an even number of iterations per graph so that the buffers are back in place at every replay
    std::printf("sharded swap, world=1, no publish : %.2f us/launch\n", time_graph(s, N, [&] { one(false); }));
    std::printf("publish alone                     : %.2f us/launch\n", time_graph(s, N, [&] {
      eb_publish pb; std::memset(&pb, 0, sizeof(pb));
      pb.rank = 0; pb.world = 1; pb.ntemps_total = T; pb.nwalkers = W; pb.temp_begin[0] = 0; pb.temp_begin[1] = T;
      pb.logl_local = cur.logl; pb.logl_all_peer[0] = logl_all; pb.flags_peer[0] = (uint64_t*)flags;
      EB(eb_publish_logl(&pb, ctrl, s));
    }));
    CK(cudaStreamSynchronize(s));
    CK(cudaMemset(ctrl, 0, sizeof(eb_ctrl))); CK(cudaMemset(flags, 0, 16 * 8));   // flag words count publish CTAs since iteration 0
    std::printf("publish + sharded swap, world=1   : %.2f us/pair\n", time_graph(s, N, [&] { one(true); }));
  }
#ifndef NO_MARKS
  // ---- globaltimer spread of the marks over all CTAs of ONE launch (ns relative to the first CTA's first mark) ------
  {
    unsigned long long mn[64], mx[64];
    CK(cudaStreamSynchronize(s));
    eb_debug_marks_swap_global(mn, mx, 1);
    EB(eb_pt_swap(&st, &wr, &ad, ctrl, s));
    CK(cudaStreamSynchronize(s));
    if (eb_debug_marks_swap_global(mn, mx, 0) == 0) {
      std::printf("swap marks, ns since CTA 0 start [CTA 0 .. adapt CTA]:");
      for (int i = 16; i <= 25; ++i) if (i != 22) std::printf(" m%d=%llu", i, mn[i] - mn[16]);
      std::printf(" | adapting warp: start=%llu end=%llu", mx[28] - mn[16], mx[23] - mn[16]);
      std::printf(" | CTA 0: counts=%llu rows=%llu", mn[20] - mn[16], mn[21] - mn[16]);
      static unsigned long long cta[8 * 1024];
      if (eb_debug_marks_swap_cta(cta) == 0) {
        const int nreal = (W + 15) / 16 < 1024 ? (W + 15) / 16 : 1024;   // chain CTAs of the 16-rung shape (16 chains per CTA)
        unsigned long long c4 = 0, c5 = 0;
        for (int b = 0; b < nreal; ++b) { if (cta[4 * 1024 + b] > c4) c4 = cta[4 * 1024 + b]; if (cta[5 * 1024 + b] > c5) c5 = cta[5 * 1024 + b]; }
        std::printf(" | latest over the first %d CTAs: counts=%llu rows=%llu", nreal, c4 - mn[16], c5 - mn[16]);
      }
      std::printf("\n");
    }
    eb_debug_marks_stretch_global(mn, mx, 1);
    EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s));
    CK(cudaStreamSynchronize(s));
    if (eb_debug_marks_stretch_global(mn, mx, 0) == 0) {
      std::printf("stretch marks (both launches), ns since first CTA start [first .. last]:");
      for (int i = 0; i <= 7; ++i) if (i != 5) std::printf(" m%d=[%llu..%llu]", i, mn[i] - mn[0], mx[i] - mn[0]);
      std::printf("\n");
    }
  }
  // ---- phase marks (profiling build) ---------------------------------------------------------------
  long long m[64];
  EB(eb_stretch_step(&st, &prior, &like, 2.0, &sr, acc, cnt, s));
  EB(eb_pt_swap(&st, &wr, &ad, ctrl, s));
  CK(cudaStreamSynchronize(s));
  if (eb_debug_marks_stretch(m) == 0) {
    std::printf("stretch CTA(0,0) thread 0 (split 0) cycles: rng-init=%lld prepare=%lld stage params=%lld finish=%lld barrier=%lld total=%lld\n",
                m[1] - m[0], m[2] - m[1], m[3] - m[2], m[6] - m[4], m[7] - m[6], m[7] - m[0]);
    EB(eb_debug_marks_swap(m));
    const char* wn[] = {"start", "phase0 keys", "phase1 pos/logl/logu", "phase2 cascade", "counts", "phase3 rows", "tail-last"};
    std::printf("swap    CTA 0 cycles   : ");
    for (int i = 17; i <= 22; ++i) std::printf("%s=%lld ", wn[i - 16], m[i] - m[i - 1]);
    std::printf(" total=%lld\n", m[22] - m[16]);
  }
#endif
  return 0;
}
