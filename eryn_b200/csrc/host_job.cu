// eb_run_host — the reference-facing entry point with HOST buffers.
//
// What a caller of the reference does per iteration is `move.propose(model, state)` on NumPy
// arrays (ensemble.py:974) followed by `temper_comps` (red_blue.py:330-331).  This entry point is
// the same contract over the C ABI: host arrays in, `niter` iterations of (move + swap pass) on
// the device in philox mode, host arrays out.  Device scratch is cached between calls.
//
// Two schedules, same results:
//   plain      upload everything -> niter x (move, swap pass) -> download everything: copy engines, one stream;
//   wavefront  (one tempered iteration on page-locked host arrays, the drop-in call of INTEGRATION.md)
//              The ladder walks hot -> cold (tempering.py:515) and rung i is final once the swap (i, i-1) is decided, so
//              the temperatures travel hottest first in G groups on four streams:
//                  s_in   upload of group g (coords, logl, logp)
//                  s      move kernels on the temperatures of group g, as soon as the group has landed
//                  s_sw   eb_pt_swap_range over the group's rungs and the boundary rung of the hotter group; after the
//                         last group eb_pt_swap_finish (counts, ladder adaptation)
//                  s_out  download of the rungs the range just made final
//              so that uploads, kernels and downloads overlap (PCIe is full duplex).  The transfers are kernels that
//              read / write the mapped host arrays themselves (zc_copy_kernel below): on this platform a copy-engine
//              operation between two kernels costs more than a group takes to cross the link.  The schedule is
//              captured into a CUDA graph the second time a job with the same buffers and parameters is seen and
//              replayed afterwards (one cudaGraphLaunch instead of ~10 driver calls per group).
//              Measured on config 2 (16 x 4096 x 8-d, 2 x 5.25 MB per call): plain 262 us, wavefront 222 us; the link
//              carries ~55 GB/s summed over both directions while both are busy (profiles/r02_e2e_probe*.txt).
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
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

constexpr int MAX_GROUPS = 16;

// everything a captured wavefront graph has baked in: a job that differs in any of these is captured anew
struct PipeKey {
  int32_t T, W, D, G, mv, like_kind, like_ncomp, like_nparams, permute, randomize_split;
  int32_t b[MAX_GROUPS + 1];
  const void *coords_host, *logl_host, *logp_host, *acc_cnt_host;
  const void *d_coords, *d_logl, *d_logp, *d_small, *d_acc, *d_acc_cnt, *pin_in, *pin_out;
  double stretch_a, gauss_scale;
  uint64_t seed;
  eb_adapt adapt;
};

struct PipeGraph {
  PipeKey key;
  int seen = 0;                       // calls with this key so far
  bool no_graph = false;              // capture failed for this key: issue the schedule directly
  cudaGraphExec_t exec = nullptr;
};

struct HostCtx {
  Pool coords, logl, logp, small, acc, acc_cnt, row_scratch, logp_scratch;
  PinnedPool pin_in, pin_out;
  cudaStream_t stream = nullptr, s_in = nullptr, s_sw = nullptr, s_out = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join_sw = nullptr, ev_join_out = nullptr;
  cudaEvent_t ev_in[MAX_GROUPS], ev_k1[MAX_GROUPS], ev_out[MAX_GROUPS];
  PipeGraph graphs[2];                // one per move kind
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

// The wavefront schedule moves its bytes with the SMs: kernels that read / write the page-locked host arrays directly
// (mapped host memory, ~49 GB/s either way on this PCIe 5 link).  A copy-engine operation that depends on a kernel, or
// a kernel that depends on one, costs 15-30 us of hand-over on top of ~3.7 us per operation (measured:
// profiles/r02_zc_probe.txt, r02_e2e_probe_stamps.txt), which is more than a group of temperatures takes to cross the
// link; kernel-to-kernel edges cost 2-3 us.  Host reads bypass the caches (ld.cv: the host rewrites these arrays between
// calls).  One launch moves up to three segments (coords, logl, logp of a group of temperatures).
struct CopySeg { void* dst; const void* src; size_t bytes; };
struct CopyArgs { CopySeg seg[3]; int nseg; int pdl; };

__device__ __forceinline__ uint4 ld_cv_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_cv_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
template <bool FROM_HOST, typename V>
__device__ __forceinline__ V zc_load(const V* p);
template <> __device__ __forceinline__ uint4 zc_load<true, uint4>(const uint4* p) { return ld_cv_u4(p); }
template <> __device__ __forceinline__ uint4 zc_load<false, uint4>(const uint4* p) { return *p; }
template <> __device__ __forceinline__ unsigned long long zc_load<true, unsigned long long>(const unsigned long long* p) { return ld_cv_u64(p); }
template <> __device__ __forceinline__ unsigned long long zc_load<false, unsigned long long>(const unsigned long long* p) { return *p; }

template <bool FROM_HOST, typename V>
__device__ __forceinline__ void zc_copy_segment(V* __restrict__ dst, const V* __restrict__ src, size_t n) {
  constexpr int U = 4;                       // loads in flight per thread
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = zc_load<FROM_HOST, V>(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) dst[i] = zc_load<FROM_HOST, V>(src + i);
}

template <bool FROM_HOST>
__global__ void __launch_bounds__(256) zc_copy_kernel(const __grid_constant__ CopyArgs a) {
  // the next transfer of the same direction may start at once (programmatic dependent launch): the link never drains
  // between two groups
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int k = 0; k < a.nseg; ++k) {
    const size_t addr = reinterpret_cast<size_t>(a.seg[k].dst) | reinterpret_cast<size_t>(a.seg[k].src) | a.seg[k].bytes;
    if ((addr & 15) == 0)
      zc_copy_segment<FROM_HOST, uint4>((uint4*)a.seg[k].dst, (const uint4*)a.seg[k].src, a.seg[k].bytes / 16);
    else
      zc_copy_segment<FROM_HOST, unsigned long long>((unsigned long long*)a.seg[k].dst, (const unsigned long long*)a.seg[k].src,
                                                     a.seg[k].bytes / 8);
  }
}

// bytes must be multiples of 8 (they are: doubles, and the small blocks are padded)
int zc_copy(bool from_host, const CopyArgs& a, int max_ctas, cudaStream_t s) {
  size_t total = 0;
  for (int k = 0; k < a.nseg; ++k) total += a.seg[k].bytes;
  if (total == 0) return EB_OK;
  size_t grid = (total / 16 + 1023) / 1024;            // four 16-byte units per thread and trip
  if (grid < 1) grid = 1;
  if (grid > (size_t)max_ctas) grid = (size_t)max_ctas;
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  if (a.pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  if (from_host) HJ_CUDA(cudaLaunchKernelEx(&cfg, zc_copy_kernel<true>, a));
  else HJ_CUDA(cudaLaunchKernelEx(&cfg, zc_copy_kernel<false>, a));
  return EB_OK;
}

// EB_HOST_STAMPS=1 (diagnostic): one-thread kernels note %globaltimer behind every step of the wavefront schedule and the
// call prints the timeline to stderr
constexpr int MAX_STAMPS = 8 * MAX_GROUPS + 8;
__global__ void stamp_kernel(unsigned long long* out) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *out = t;
}
struct Stamps {
  unsigned long long* dev = nullptr;
  const char* what[MAX_STAMPS];
  int group[MAX_STAMPS];
  int n = 0;
  bool on = false;
  void mark(const char* w, int g, cudaStream_t s) {
    if (!on || n >= MAX_STAMPS) return;
    what[n] = w; group[n] = g;
    stamp_kernel<<<1, 1, 0, s>>>(dev + n);
    ++n;
  }
};
Stamps g_stamps;

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  // page-locked (cudaHostAlloc / cudaHostRegister) AND addressable by the kernels under the same pointer
  return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

int ensure_pipe_objects(HostCtx& cx) {
  if (cx.s_in) return EB_OK;
  HJ_CUDA(cudaStreamCreateWithFlags(&cx.s_in, cudaStreamNonBlocking));
  HJ_CUDA(cudaStreamCreateWithFlags(&cx.s_sw, cudaStreamNonBlocking));
  HJ_CUDA(cudaStreamCreateWithFlags(&cx.s_out, cudaStreamNonBlocking));
  HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_fork, cudaEventDisableTiming));
  HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_join_sw, cudaEventDisableTiming));
  HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_join_out, cudaEventDisableTiming));
  for (int g = 0; g < MAX_GROUPS; ++g) {
    HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_in[g], cudaEventDisableTiming));
    HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_k1[g], cudaEventDisableTiming));
    HJ_CUDA(cudaEventCreateWithFlags(&cx.ev_out[g], cudaEventDisableTiming));
  }
  return EB_OK;
}

// layout of the small device block: [eb_ctrl | betas T | prior lo, hi, logpdf (3 D) | likelihood parameters]; the head
// [eb_ctrl | betas] is what comes back
struct SmallLayout {
  size_t off_betas, off_prior, off_like, bytes;
};
SmallLayout small_layout(int T, int D, int like_nparams) {
  SmallLayout l;
  l.off_betas = (sizeof(eb_ctrl) + 15) & ~(size_t)15;
  l.off_prior = l.off_betas + (((size_t)T * sizeof(double) + 15) & ~(size_t)15);
  l.off_like = l.off_prior + 3 * (size_t)D * sizeof(double);
  l.bytes = l.off_like + ((size_t)like_nparams + 1) * sizeof(double);
  return l;
}

struct Plan {
  eb_state st;
  eb_prior prior;
  eb_like like;
  eb_ctrl* dctrl;
  eb_stretch_rng srng;
  eb_gauss_rng grng;
  eb_swap_rng wrng;
  uint8_t* acc;
  uint32_t* acc_cnt;
};

int run_move(const Plan& pl, const eb_host_job* job, int mv, const eb_state& st, size_t slot0, cudaStream_t s) {
  if (mv == 0)
    return eb_stretch_step(&st, &pl.prior, &pl.like, job->stretch_a, &pl.srng, pl.acc + slot0, pl.acc_cnt + slot0, s);
  return eb_gaussian_step(&st, &pl.prior, &pl.like, &pl.grng, pl.acc + slot0, pl.acc_cnt + slot0, s);
}

// the small results: head of the control block (iter, time, error), the swap counts and the ladder, into pinned memory
int download_small(HostCtx& cx, const Plan& pl, const SmallLayout& lay, int T, bool tempered, cudaStream_t s) {
  unsigned char* out = (unsigned char*)cx.pin_out.p;
  const unsigned char* dsmall = (const unsigned char*)cx.small.p;
  eb_ctrl* hc = (eb_ctrl*)out;
  HJ_CUDA(cudaMemcpyAsync(hc, pl.dctrl, offsetof(eb_ctrl, swaps_work), cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(hc->swaps_accepted, pl.dctrl->swaps_accepted, sizeof(int32_t) * (size_t)(T > 1 ? T - 1 : 1),
                          cudaMemcpyDeviceToHost, s));
  if (tempered)
    HJ_CUDA(cudaMemcpyAsync(out + lay.off_betas, dsmall + lay.off_betas, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, s));
  return EB_OK;
}

// ---- plain schedule ------------------------------------------------------------------------------------------------
int issue_plain(HostCtx& cx, eb_host_job* job, int32_t niter, Plan& pl, const SmallLayout& lay) {
  cudaStream_t s = cx.stream;
  const int T = job->ntemps, W = job->nwalkers;
  const size_t n = (size_t)T * W;
  const size_t bc = n * job->nleaves * job->ndim * sizeof(double), bs = n * sizeof(double);
  HJ_CUDA(cudaMemcpyAsync(cx.coords.p, job->coords_host, bc, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemcpyAsync(cx.logl.p, job->logl_host, bs, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemcpyAsync(cx.logp.p, job->logp_host, bs, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemcpyAsync(cx.small.p, cx.pin_in.p, lay.bytes, cudaMemcpyHostToDevice, s));
  HJ_CUDA(cudaMemsetAsync(cx.acc_cnt.p, 0, n * sizeof(uint32_t), s));
  pl.srng.iter_dev = &pl.dctrl->iter_next;   // published early by the swap pass: the next move may start its draws
  pl.srng.pdl_chain = 1;
  for (int it = 0; it < niter; ++it) {
    const int mv = job->move_schedule_host ? job->move_schedule_host[it] : 0;
    int rc = run_move(pl, job, mv, pl.st, 0, s);
    if (rc) return rc;
    rc = pl.st.betas ? eb_pt_swap(&pl.st, &pl.wrng, &job->adapt, pl.dctrl, s) : eb_advance_iter(pl.dctrl, s);
    if (rc) return rc;
  }
  HJ_CUDA(cudaMemcpyAsync(job->coords_host, cx.coords.p, bc, cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(job->logl_host, cx.logl.p, bs, cudaMemcpyDeviceToHost, s));
  HJ_CUDA(cudaMemcpyAsync(job->logp_host, cx.logp.p, bs, cudaMemcpyDeviceToHost, s));
  if (job->accepted_count_host)
    HJ_CUDA(cudaMemcpyAsync(job->accepted_count_host, cx.acc_cnt.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  return download_small(cx, pl, lay, T, pl.st.betas != nullptr, s);
}

// ---- wavefront schedule (one tempered iteration) ------------------------------------------------------------------------
// group g covers the temperatures [b[g+1], b[g]), b[0] = T > b[1] > ... > b[G] = 0
int issue_wavefront(HostCtx& cx, eb_host_job* job, int mv, int G, const int* b, Plan& pl, const SmallLayout& lay) {
  cudaStream_t s = cx.stream, s_in = cx.s_in, s_sw = cx.s_sw, s_out = cx.s_out;
  const int T = job->ntemps, W = job->nwalkers, LD = job->nleaves * job->ndim;
  double* dco = (double*)cx.coords.p; double* dll = (double*)cx.logl.p; double* dlp = (double*)cx.logp.p;
  const int pdl = env_int("EB_HOST_PDL", 1), ctas = env_int("EB_HOST_COPY_CTAS", 32);
  int rc;

  g_stamps.n = 0;
  g_stamps.mark("start", -1, s);
  HJ_CUDA(cudaEventRecord(cx.ev_fork, s));
  HJ_CUDA(cudaStreamWaitEvent(s_in, cx.ev_fork, 0));
  {
    CopyArgs a{};   // control block, ladder, prior and likelihood parameters: one small block
    a.seg[0] = CopySeg{cx.small.p, cx.pin_in.p, (lay.bytes + 7) & ~(size_t)7};
    a.nseg = 1;
    rc = zc_copy(true, a, 4, s);
    if (rc) return rc;
  }
  HJ_CUDA(cudaMemsetAsync(cx.acc_cnt.p, 0, (size_t)T * W * sizeof(uint32_t), s));
  // uploads, hottest group first
  for (int g = 0; g < G; ++g) {
    const size_t lo = (size_t)b[g + 1] * W, cnt = (size_t)(b[g] - b[g + 1]) * W;
    CopyArgs a{};
    a.seg[0] = CopySeg{dco + lo * LD, job->coords_host + lo * LD, cnt * LD * sizeof(double)};
    a.seg[1] = CopySeg{dll + lo, job->logl_host + lo, cnt * sizeof(double)};
    a.seg[2] = CopySeg{dlp + lo, job->logp_host + lo, cnt * sizeof(double)};
    a.nseg = 3;
    a.pdl = pdl && g > 0;
    rc = zc_copy(true, a, ctas, s_in);
    if (rc) return rc;
    g_stamps.mark("in", g, s_in);
    HJ_CUDA(cudaEventRecord(cx.ev_in[g], s_in));
  }
  pl.srng.iter_dev = &pl.dctrl->iter;   // no pass in front of the move: plain launches keyed by the uploaded counter
  pl.srng.pdl_chain = 0;
  for (int g = 0; g < G; ++g) {
    const int t_lo = b[g + 1], t_hi = b[g];
    const size_t lo = (size_t)t_lo * W;
    // the move on the temperatures of this group (random streams keyed by the global temperature: temp_offset)
    HJ_CUDA(cudaStreamWaitEvent(s, cx.ev_in[g], 0));
    eb_state sg = pl.st;
    sg.ntemps = t_hi - t_lo; sg.temp_offset = t_lo;
    sg.coords = dco + lo * LD; sg.logl = dll + lo; sg.logp = dlp + lo; sg.betas = pl.st.betas + t_lo;
    rc = run_move(pl, job, mv, sg, lo, s);
    if (rc) return rc;
    g_stamps.mark("move done", g, s);
    HJ_CUDA(cudaEventRecord(cx.ev_k1[g], s));
    // the rungs of this group, starting at the boundary rung of the hotter group
    const int r_hi = g == 0 ? T - 1 : t_hi, r_lo = t_lo;
    HJ_CUDA(cudaStreamWaitEvent(s_sw, cx.ev_k1[g], 0));
    rc = eb_pt_swap_range(&pl.st, &pl.wrng, pl.dctrl, r_hi, r_lo, s_sw);
    if (rc) return rc;
    g_stamps.mark("rungs done", g, s_sw);
    HJ_CUDA(cudaEventRecord(cx.ev_out[g], s_sw));
    if (g == G - 1) {   // off the path of the row downloads: only the small results wait for it
      rc = eb_pt_swap_finish(&pl.st, &pl.wrng, &job->adapt, pl.dctrl, s_sw);
      if (rc) return rc;
    }
    // rungs r_lo+1 .. r_hi are final (the last group also finishes rung 0)
    const int f_lo = g == G - 1 ? 0 : r_lo + 1;
    const size_t flo = (size_t)f_lo * W, fcnt = (size_t)(r_hi - f_lo + 1) * W;
    HJ_CUDA(cudaStreamWaitEvent(s_out, cx.ev_out[g], 0));
    if (r_hi < f_lo) continue;
    CopyArgs a{};
    a.seg[0] = CopySeg{job->coords_host + flo * LD, dco + flo * LD, fcnt * LD * sizeof(double)};
    a.seg[1] = CopySeg{job->logl_host + flo, dll + flo, fcnt * sizeof(double)};
    a.seg[2] = CopySeg{job->logp_host + flo, dlp + flo, fcnt * sizeof(double)};
    a.nseg = 3;
    rc = zc_copy(false, a, ctas, s_out);
    if (rc) return rc;
    g_stamps.mark("out", g, s_out);
  }
  if (job->accepted_count_host) {
    const size_t bytes = (size_t)T * W * sizeof(uint32_t);
    if (bytes & 7) {   // an odd number of counters: not a whole number of 8-byte units
      HJ_CUDA(cudaMemcpyAsync(job->accepted_count_host, cx.acc_cnt.p, bytes, cudaMemcpyDeviceToHost, s));
    } else {
      CopyArgs a{};
      a.seg[0] = CopySeg{job->accepted_count_host, cx.acc_cnt.p, bytes};
      a.nseg = 1;
      rc = zc_copy(false, a, 8, s);
      if (rc) return rc;
    }
  }
  {
    // the small results: head of the control block (iter, time, error), the swap counts and the ladder
    unsigned char* out = (unsigned char*)cx.pin_out.p;
    const unsigned char* dsmall = (const unsigned char*)cx.small.p;
    eb_ctrl* hc = (eb_ctrl*)out;
    CopyArgs a{};
    a.seg[0] = CopySeg{hc, pl.dctrl, offsetof(eb_ctrl, swaps_work) & ~(size_t)7};
    a.seg[1] = CopySeg{hc->swaps_accepted, pl.dctrl->swaps_accepted, (sizeof(int32_t) * (size_t)(T > 1 ? T - 1 : 1) + 7) & ~(size_t)7};
    a.seg[2] = CopySeg{out + lay.off_betas, dsmall + lay.off_betas, (size_t)T * sizeof(double)};
    a.nseg = 3;
    rc = zc_copy(false, a, 1, s_sw);
    if (rc) return rc;
  }
  HJ_CUDA(cudaEventRecord(cx.ev_join_sw, s_sw));
  HJ_CUDA(cudaEventRecord(cx.ev_join_out, s_out));
  HJ_CUDA(cudaStreamWaitEvent(s, cx.ev_join_sw, 0));
  HJ_CUDA(cudaStreamWaitEvent(s, cx.ev_join_out, 0));
  g_stamps.mark("joined", -1, s);
  return EB_OK;
}

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
  const bool tempered = job->betas_host != nullptr;
  const SmallLayout lay = small_layout(T, D, job->like_nparams);
  if (cx.coords.ensure(bc) || cx.logl.ensure(bs) || cx.logp.ensure(bs) || cx.small.ensure(lay.bytes) || cx.acc.ensure(n) ||
      cx.acc_cnt.ensure(n * sizeof(uint32_t)) || cx.pin_in.ensure(lay.bytes) || cx.pin_out.ensure(lay.off_prior) ||
      (T > 32 && (cx.row_scratch.ensure(bc) || cx.logp_scratch.ensure(bs))))
    return EB_ERR_CUDA;

  // ---- the small inputs travel as ONE block, built in pinned memory ---------------------------------------------------
  {
    unsigned char* blk = (unsigned char*)cx.pin_in.p;
    eb_ctrl* hc_in = (eb_ctrl*)blk;
    std::memset(hc_in, 0, sizeof(eb_ctrl));
    hc_in->iter = job->iter0;
    hc_in->iter_next = job->iter0;
    hc_in->time = job->adapt_time0;
    if (tempered) std::memcpy(blk + lay.off_betas, job->betas_host, (size_t)T * sizeof(double));
    double* pr = (double*)(blk + lay.off_prior);
    for (int d = 0; d < D; ++d) {
      double lo = job->prior_lo_host[d], hi = job->prior_hi_host[d];
      if (lo > hi) { double t = lo; lo = hi; hi = t; }   // prior.py:29-32
      if (lo == hi) return EB_ERR_INVALID;               // prior.py:33-34
      pr[d] = lo; pr[D + d] = hi; pr[2 * D + d] = std::log(1.0 / (hi - lo));  // prior.py:40-41
    }
    if (job->like_nparams > 0)
      std::memcpy(blk + lay.off_like, job->like_params_host, (size_t)job->like_nparams * sizeof(double));
  }
  unsigned char* dsmall = (unsigned char*)cx.small.p;
  const double* dprior = (const double*)(dsmall + lay.off_prior);

  Plan pl;
  std::memset(&pl, 0, sizeof(pl));
  pl.st.ntemps = T; pl.st.nwalkers = W; pl.st.nleaves = L; pl.st.ndim = D; pl.st.temp_offset = 0; pl.st.inds_stride = 0;
  pl.st.coords = (double*)cx.coords.p; pl.st.logl = (double*)cx.logl.p; pl.st.logp = (double*)cx.logp.p;
  pl.st.inds = nullptr; pl.st.betas = tempered ? (double*)(dsmall + lay.off_betas) : nullptr;
  pl.prior = eb_prior{dprior, dprior + D, dprior + 2 * D, nullptr};
  pl.like = eb_like{job->like_kind, job->like_ncomp, job->like_nparams, 0, (const double*)(dsmall + lay.off_like)};
  pl.dctrl = (eb_ctrl*)dsmall;
  pl.acc = (uint8_t*)cx.acc.p; pl.acc_cnt = (uint32_t*)cx.acc_cnt.p;
  pl.srng.mode = EB_RNG_PHILOX; pl.srng.randomize_split = job->randomize_split; pl.srng.seed = job->seed;
  pl.grng.mode = EB_RNG_PHILOX; pl.grng.cov_kind = 0; pl.grng.scale = job->gauss_scale; pl.grng.seed = job->seed;
  pl.grng.iter_dev = &pl.dctrl->iter;
  pl.wrng.mode = EB_RNG_PHILOX; pl.wrng.permute = job->permute; pl.wrng.seed = job->seed; pl.wrng.iter_dev = &pl.dctrl->iter;
  if (T > 32) { pl.wrng.row_scratch = (double*)cx.row_scratch.p; pl.wrng.logp_scratch = (double*)cx.logp_scratch.p; }

  // ---- which schedule ---------------------------------------------------------------------------------------------------
  // EB_HOST_PIPE: 0 = always plain, 1 (default) = wavefront when it pays (>= 1 MiB of state), 2 = wavefront whenever legal
  const int pipe_mode = env_int("EB_HOST_PIPE", 1);
  const int groups_env = env_int("EB_HOST_GROUPS", 0);
  const int graph_mode = env_int("EB_HOST_GRAPH", 1);
  const size_t state_bytes = bc + 2 * bs;
  bool pipe = pipe_mode > 0 && niter == 1 && tempered && T >= 2 && T <= 128 && D * L <= 32 &&
              (pipe_mode > 1 || state_bytes >= ((size_t)1 << 20));
  if (pipe) pipe = is_pinned(job->coords_host) && is_pinned(job->logl_host) && is_pinned(job->logp_host) &&
                   (!job->accepted_count_host || is_pinned(job->accepted_count_host));
  int rc = EB_OK;
  if (!pipe) {
    rc = issue_plain(cx, job, niter, pl, lay);
    if (rc) return rc;
  } else {
    rc = ensure_pipe_objects(cx);
    if (rc) return rc;
    g_stamps.on = env_int("EB_HOST_STAMPS", 0) != 0;
    if (g_stamps.on && !g_stamps.dev) HJ_CUDA(cudaMalloc(&g_stamps.dev, sizeof(unsigned long long) * MAX_STAMPS));
    int G = groups_env > 0 ? groups_env : (int)(state_bytes >> 20);   // 1 MiB or more per group, at most four
    if (G > 4 && groups_env <= 0) G = 4;
    if (G > T) G = T;
    if (G > MAX_GROUPS) G = MAX_GROUPS;
    if (G < 1) G = 1;
    // group boundaries, hot -> cold: equal sizes, or EB_HOST_SPLIT="n0,n1,..." temperatures per group (experiments)
    int b[MAX_GROUPS + 1];
    for (int g = 0; g <= G; ++g) b[g] = T - (int)(((long long)T * g) / G);
    if (const char* sp = std::getenv("EB_HOST_SPLIT")) {
      int sizes[MAX_GROUPS], ns = 0, sum = 0;
      for (const char* q = sp; *q && ns < MAX_GROUPS;) {
        const int v = std::atoi(q);
        if (v < 1) { ns = 0; break; }
        sizes[ns++] = v; sum += v;
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
      if (ns >= 1 && sum == T) {
        G = ns;
        b[0] = T;
        for (int g = 0; g < G; ++g) b[g + 1] = b[g] - sizes[g];
      }
    }
    const int mv = job->move_schedule_host ? (job->move_schedule_host[0] ? 1 : 0) : 0;
    PipeKey key;
    std::memset(&key, 0, sizeof(key));
    for (int g = 0; g <= G; ++g) key.b[g] = b[g];
    key.T = T; key.W = W; key.D = D; key.G = G; key.mv = mv; key.like_kind = job->like_kind; key.like_ncomp = job->like_ncomp;
    key.like_nparams = job->like_nparams; key.permute = job->permute; key.randomize_split = job->randomize_split;
    key.coords_host = job->coords_host; key.logl_host = job->logl_host; key.logp_host = job->logp_host;
    key.acc_cnt_host = job->accepted_count_host;
    key.d_coords = cx.coords.p; key.d_logl = cx.logl.p; key.d_logp = cx.logp.p; key.d_small = cx.small.p;
    key.d_acc = cx.acc.p; key.d_acc_cnt = cx.acc_cnt.p; key.pin_in = cx.pin_in.p; key.pin_out = cx.pin_out.p;
    key.stretch_a = job->stretch_a; key.gauss_scale = job->gauss_scale; key.seed = job->seed; key.adapt = job->adapt;
    PipeGraph& pg = cx.graphs[mv];
    const bool same = pg.seen > 0 && std::memcmp(&pg.key, &key, sizeof(key)) == 0;
    if (!same) {
      if (pg.exec) cudaGraphExecDestroy(pg.exec);
      pg.exec = nullptr;
      pg.key = key;
      pg.seen = 0;
      pg.no_graph = false;
    }
    ++pg.seen;
    if (pg.exec) {
      HJ_CUDA(cudaGraphLaunch(pg.exec, s));
    } else if (graph_mode && !pg.no_graph && pg.seen >= 2) {
      // second sighting of this job: capture the schedule, replay it from now on
      cudaGraph_t graph = nullptr;
      HJ_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      rc = issue_wavefront(cx, job, mv, G, b, pl, lay);
      const cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        std::fprintf(stderr, "eb_run_host: capture of the wavefront schedule failed (%s); issuing it directly\n",
                     ce != cudaSuccess ? cudaGetErrorString(ce) : "launch error");
        pg.no_graph = true;
        rc = issue_wavefront(cx, job, mv, G, b, pl, lay);
        if (rc) return rc;
      } else {
        const cudaError_t ie = cudaGraphInstantiate(&pg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
          pg.exec = nullptr;
          cudaGetLastError();
          pg.no_graph = true;
          rc = issue_wavefront(cx, job, mv, G, b, pl, lay);
          if (rc) return rc;
        } else {
          HJ_CUDA(cudaGraphLaunch(pg.exec, s));
        }
      }
    } else {
      rc = issue_wavefront(cx, job, mv, G, b, pl, lay);
      if (rc) return rc;
    }
  }

  HJ_CUDA(cudaStreamSynchronize(s));
  if (pipe && g_stamps.on && g_stamps.n > 0) {
    unsigned long long t[MAX_STAMPS];
    HJ_CUDA(cudaMemcpy(t, g_stamps.dev, sizeof(unsigned long long) * g_stamps.n, cudaMemcpyDeviceToHost));
    std::fprintf(stderr, "eb_run_host timeline (us after the first kernel):");
    for (int i = 0; i < g_stamps.n; ++i)
      std::fprintf(stderr, " [%s %d: %.1f]", g_stamps.what[i], g_stamps.group[i], (double)(long long)(t[i] - t[0]) * 1e-3);
    std::fprintf(stderr, "\n");
  }
  const unsigned char* out = (const unsigned char*)cx.pin_out.p;
  const eb_ctrl* hc = (const eb_ctrl*)out;
  if (hc->error) {
    std::fprintf(stderr, "eb_run_host: device error %u (a bounded in-kernel wait ran out)\n", hc->error);
    return EB_ERR_CUDA;
  }
  if (tempered) std::memcpy(job->betas_host, out + lay.off_betas, (size_t)T * sizeof(double));
  if (job->swaps_accepted_host)
    for (int i = 0; i + 1 < T; ++i) job->swaps_accepted_host[i] = hc->swaps_accepted[i];
  job->iter0 = hc->iter;
  job->adapt_time0 = hc->time;
  return EB_OK;
}
