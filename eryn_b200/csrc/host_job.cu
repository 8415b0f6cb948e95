// eb_run_host — the reference-facing entry point with HOST buffers.
//
// What a caller of the reference does per iteration is `move.propose(model, state)` on NumPy
// arrays (ensemble.py:974) followed by `temper_comps` (red_blue.py:330-331).  This entry point is
// the same contract over the C ABI: host arrays in, `niter` iterations of (move + swap pass) on
// the device in philox mode, host arrays out.  Device scratch is cached between calls.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/eryn_b200.h"

namespace {

struct Pool {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return 1;
    cap = bytes;
    return 0;
  }
};

// page-locked host staging for the small per-call blocks (a cudaMemcpyAsync from pageable memory is staged by the driver
// and blocks the caller for every such copy)
struct PinnedPool {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return 1;
    cap = bytes;
    return 0;
  }
};

struct HostCtx {
  Pool coords, logl, logp, betas, small, acc, acc_cnt, row_scratch, logp_scratch;
  PinnedPool pin_in, pin_out;
  cudaStream_t stream = nullptr;
  std::mutex mu;
};
HostCtx g_ctx;

#define HJ_CUDA(call)                                                      \
  do {                                                                     \
    cudaError_t e_ = (call);                                               \
    if (e_ != cudaSuccess) {                                               \
      std::fprintf(stderr, "eb_run_host: %s: %s\n", #call, cudaGetErrorString(e_)); \
      return e_ == cudaErrorNoDevice ? EB_ERR_NODEVICE : EB_ERR_CUDA;      \
    }                                                                      \
  } while (0)

}  // namespace

extern "C" int eb_run_host(eb_host_job* job, int32_t niter) {
  if (!job || niter < 0) return EB_ERR_INVALID;
  if (eb_device_count() < 1) return EB_ERR_NODEVICE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  HostCtx& cx = g_ctx;
  const int T = job->ntemps, W = job->nwalkers, L = job->nleaves, D = job->ndim;
  if (T < 1 || W < 2 || L != 1 || D < 1 || !job->coords_host || !job->logl_host || !job->logp_host ||
      !job->prior_lo_host || !job->prior_hi_host)
    return EB_ERR_INVALID;
  if (!cx.stream) HJ_CUDA(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
  cudaStream_t s = cx.stream;
  const size_t n = (size_t)T * W;
  const size_t bc = n * L * D * sizeof(double), bs = n * sizeof(double);
  // the small inputs travel as ONE block: [eb_ctrl | prior lo, hi, logpdf (3 D) | likelihood parameters], built in pinned memory
  const size_t off_prior = (sizeof(eb_ctrl) + 15) & ~(size_t)15;
  const size_t off_like = off_prior + 3 * (size_t)D * sizeof(double);
  const size_t small_bytes = off_like + ((size_t)job->like_nparams + 1) * sizeof(double);
  if (cx.coords.ensure(bc) || cx.logl.ensure(bs) || cx.logp.ensure(bs) || cx.betas.ensure(T * sizeof(double)) ||
      cx.small.ensure(small_bytes) || cx.acc.ensure(n) || cx.acc_cnt.ensure(n * sizeof(uint32_t)) ||
      cx.pin_in.ensure(small_bytes) || cx.pin_out.ensure(sizeof(eb_ctrl)) ||
      (T > 32 && (cx.row_scratch.ensure(bc) || cx.logp_scratch.ensure(bs))))
    return EB_ERR_CUDA;

  // ---- host -> device -------------------------------------------------------------------------
  HJ_CUDA(cudaMemcpyAsync(cx.coords.p, job->coords_host, bc, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemcpyAsync(cx.logl.p, job->logl_host, bs, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemcpyAsync(cx.logp.p, job->logp_host, bs, cudaMemcpyHostToDevice, s));
  if (job->betas_host)
    HJ_CUDA(cudaMemcpyAsync(cx.betas.p, job->betas_host, T * sizeof(double), cudaMemcpyHostToDevice, s));
  {
    unsigned char* blk = (unsigned char*)cx.pin_in.p;
    eb_ctrl* hc_in = (eb_ctrl*)blk;
    std::memset(hc_in, 0, sizeof(eb_ctrl));
    hc_in->iter = job->iter0;
    hc_in->iter_next = job->iter0;
    hc_in->time = job->adapt_time0;
    double* pr = (double*)(blk + off_prior);
    for (int d = 0; d < D; ++d) {
      double lo = job->prior_lo_host[d], hi = job->prior_hi_host[d];
      if (lo > hi) { double t = lo; lo = hi; hi = t; }   // prior.py:29-32
      if (lo == hi) return EB_ERR_INVALID;               // prior.py:33-34
      pr[d] = lo; pr[D + d] = hi; pr[2 * D + d] = std::log(1.0 / (hi - lo));  // prior.py:40-41
    }
    if (job->like_nparams > 0)
      std::memcpy(blk + off_like, job->like_params_host, (size_t)job->like_nparams * sizeof(double));
    HJ_CUDA(cudaMemcpyAsync(cx.small.p, blk, small_bytes, cudaMemcpyHostToDevice, s));
  }
  HJ_CUDA(cudaMemsetAsync(cx.acc_cnt.p, 0, n * sizeof(uint32_t), s));
  unsigned char* dsmall = (unsigned char*)cx.small.p;
  const double* dprior = (const double*)(dsmall + off_prior);
  const double* dlike = (const double*)(dsmall + off_like);

  // ---- iterations -----------------------------------------------------------------------------
  eb_state st;
  st.ntemps = T; st.nwalkers = W; st.nleaves = L; st.ndim = D; st.temp_offset = 0; st.inds_stride = 0;
  st.coords = (double*)cx.coords.p; st.logl = (double*)cx.logl.p; st.logp = (double*)cx.logp.p;
  st.inds = nullptr; st.betas = job->betas_host ? (double*)cx.betas.p : nullptr;
  eb_prior prior{dprior, dprior + D, dprior + 2 * D, nullptr};
  eb_like like{job->like_kind, job->like_ncomp, job->like_nparams, 0, dlike};
  eb_ctrl* dctrl = (eb_ctrl*)dsmall;
  eb_stretch_rng srng;
  std::memset(&srng, 0, sizeof(srng));
  srng.mode = EB_RNG_PHILOX; srng.randomize_split = job->randomize_split; srng.seed = job->seed;
  srng.iter_dev = &dctrl->iter_next;   // published early by the swap pass: the next move may start its draws
  srng.pdl_chain = 1;
  eb_gauss_rng grng;
  std::memset(&grng, 0, sizeof(grng));
  grng.mode = EB_RNG_PHILOX; grng.cov_kind = 0; grng.scale = job->gauss_scale; grng.seed = job->seed;
  grng.iter_dev = &dctrl->iter;
  eb_swap_rng wrng;
  std::memset(&wrng, 0, sizeof(wrng));
  wrng.mode = EB_RNG_PHILOX; wrng.permute = job->permute; wrng.seed = job->seed; wrng.iter_dev = &dctrl->iter;
  if (T > 32) { wrng.row_scratch = (double*)cx.row_scratch.p; wrng.logp_scratch = (double*)cx.logp_scratch.p; }
  for (int it = 0; it < niter; ++it) {
    const int mv = job->move_schedule_host ? job->move_schedule_host[it] : 0;
    int rc;
    if (mv == 0) {
      rc = eb_stretch_step(&st, &prior, &like, job->stretch_a, &srng, (uint8_t*)cx.acc.p, (uint32_t*)cx.acc_cnt.p, s);
    } else {
      rc = eb_gaussian_step(&st, &prior, &like, &grng, (uint8_t*)cx.acc.p, (uint32_t*)cx.acc_cnt.p, s);
    }
    if (rc) return rc;
    rc = st.betas ? eb_pt_swap(&st, &wrng, &job->adapt, dctrl, s) : eb_advance_iter(dctrl, s);
    if (rc) return rc;
  }

  // ---- device -> host -------------------------------------------------------------------------
  HJ_CUDA(cudaMemcpyAsync(job->coords_host, cx.coords.p, bc, cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(job->logl_host, cx.logl.p, bs, cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(job->logp_host, cx.logp.p, bs, cudaMemcpyDeviceToHost, s));
  if (job->betas_host)
    HJ_CUDA(cudaMemcpyAsync(job->betas_host, cx.betas.p, T * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (job->accepted_count_host)
    HJ_CUDA(cudaMemcpyAsync(job->accepted_count_host, cx.acc_cnt.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  // of the control block only the head (iter, time, error) and the swap counts come back, into pinned memory
  eb_ctrl* hc = (eb_ctrl*)cx.pin_out.p;
  HJ_CUDA(cudaMemcpyAsync(hc, dctrl, offsetof(eb_ctrl, swaps_work), cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(hc->swaps_accepted, dctrl->swaps_accepted, sizeof(int32_t) * (size_t)(T > 1 ? T - 1 : 1),
                          cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaStreamSynchronize(s));
  if (hc->error) {
    std::fprintf(stderr, "eb_run_host: device error %u (a bounded in-kernel wait ran out)\n", hc->error);
    return EB_ERR_CUDA;
  }
  if (job->swaps_accepted_host)
    for (int i = 0; i + 1 < T; ++i) job->swaps_accepted_host[i] = hc->swaps_accepted[i];
  job->iter0 = hc->iter;
  job->adapt_time0 = hc->time;
  return EB_OK;
}
