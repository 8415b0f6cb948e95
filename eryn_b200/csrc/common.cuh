// eryn_b200 — hand-written sm_100a kernels of the walker-parallel sampling hot path + C ABI.
//
// Kernels:
//   K0  eval_state_kernel         log-prior + log-like of the whole state            (abi_core.cu)
//   K1  stretch_step_kernel       fused, one launch per red/blue half chained by programmatic dependent launch:
//                                 draw -> gather complement -> stretch -> prior/like ->
//                                 tempered Metropolis test -> in-place update        (k_stretch.cu, hot kernel)
//       stretch_lanes_kernel      the same step with a walker spread over 2 / 4 lanes (HBM-sized shapes, stretch_lanes.cuh)
//   K2  gaussian_step_kernel      fused Gaussian Metropolis step over all walkers    (k_gauss.cu)
//   K3  pt_swap_kernel            chain-parallel swap ladder (decide on logl, then move only the rows
//                                 that changed rung, in place) + ladder adaptation: by the warp that draws the last
//                                 ticket, by an extra CTA (sharded passes), or deferred to the next move kernel
//                                 (lazy_adapt_apply below / adapt_flush_kernel)                      (k_swap.cu)
//   K3w pt_swap_range_kernel /    the ladder a few rungs at a time + its tail: what the wavefront schedule of
//       pt_swap_finish_kernel     eb_run_host (host_job.cu, with the zero-copy transfer kernel zc_copy_kernel) runs
//   K3r pt_pairmap_kernel         replay mode: host permutations -> per-position pair map
//   K4  stretch_propose_kernel / accept_update_kernel / box_prior_kernel   (split path)
//   K3s pt_swap_split_kernel      the swap pass of a temperature-sharded run, chains split over the ranks (k_swap_split.cu)
//   K6-K9 multi-branch kernels    reversible jump + group stretch (k_rj.cu)
//   K10 stage_pack_kernel         snapshot of a stored sample for the staged Backend.save_step (k_stage.cu)
//   K11 mt_distgen_kernel         multiple-try Metropolis with an independent proposal (k_mt.cu)
//   K12 resident_kernel           whole iterations in one launch, state resident in shared memory, a cluster per
//                                 temperature (resident.cuh; measured slower than K1 + K3 chained by PDL: opt-in)
//
// Built with --fmad=false so that +,-,*,/ round exactly like the NumPy reference.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/eryn_b200.h"
#include "likelihoods.cuh"
#include "rng.cuh"

namespace eb {

int fail(int code, const char* fmt, ...);   // sets eb_last_error(), returns code
int check_launch(const char* what);

#define EB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? EB_ERR_NODEVICE : EB_ERR_CUDA, \
                  "%s: %s", #call, cudaGetErrorString(e_));                                    \
  } while (0)

constexpr int BLOCK = 128;

// Phase timers of the profiling build (-DEB_PHASE_TIMERS, tools/microbench.cu): CTA (0,0) thread 0 records the
// SM cycle counter at named points; the production library compiles them away.
#ifdef EB_PHASE_TIMERS
#define EB_DBG_SKIP(bit) ((p.dbg_skip & (bit)) != 0)   // profiling build: switch phases off to time the rest
static __device__ long long eb_dbg_marks[64];   // one copy per translation unit (no relocatable device code)
#define EB_DEFINE_MARK_READER(NAME)                                                                   \
  extern "C" __attribute__((visibility("default"))) int NAME(long long* out_host) {                  \
    return cudaMemcpyFromSymbol(out_host, eb_dbg_marks, sizeof(long long) * 64) == cudaSuccess ? 0 : 3; \
  }                                                                                                   \
  extern "C" __attribute__((visibility("default"))) int NAME##_global(unsigned long long* mn, unsigned long long* mx, int reset) { \
    if (reset) {                                                                                      \
      unsigned long long big[64], zero[64];                                                           \
      for (int i = 0; i < 64; ++i) { big[i] = ~0ull; zero[i] = 0ull; }                                \
      cudaMemcpyToSymbol(eb_dbg_gmin, big, sizeof(big));                                              \
      cudaMemcpyToSymbol(eb_dbg_gmax, zero, sizeof(zero));                                            \
      return 0;                                                                                       \
    }                                                                                                 \
    cudaMemcpyFromSymbol(mn, eb_dbg_gmin, sizeof(unsigned long long) * 64);                           \
    return cudaMemcpyFromSymbol(mx, eb_dbg_gmax, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : 3; \
  }                                                                                                   \
  extern "C" __attribute__((visibility("default"))) int NAME##_cta(unsigned long long* out8x1024) {  \
    return cudaMemcpyFromSymbol(out8x1024, eb_dbg_cta, sizeof(unsigned long long) * 8 * 1024) == cudaSuccess ? 0 : 3; \
  }
static __device__ unsigned long long eb_dbg_gmin[64], eb_dbg_gmax[64];   // globaltimer (ns): first CTA / last CTA
static __device__ unsigned long long eb_dbg_cta[8][1024];                // marks >= 16: globaltimer of every CTA, slot i & 7
#define EB_MARK(i)                                                                     \
  do {                                                                                 \
    if (threadIdx.x == 0 && blockIdx.y == 0) {                                         \
      unsigned long long gt_;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                          \
      if (blockIdx.x == 0) { eb_dbg_gmin[i] = gt_; eb_dbg_marks[i] = clock64(); }      \
      if (blockIdx.x == gridDim.x - 1) eb_dbg_gmax[i] = gt_;                           \
      if ((i) >= 16 && blockIdx.x < 1024) eb_dbg_cta[(i) & 7][blockIdx.x] = gt_;       \
    }                                                                                  \
  } while (0)
// the same from whichever CTA gets there (e.g. the CTA that adapts the ladder): recorded as the "last CTA" value
#define EB_MARK_ANY(i)                                                                 \
  do {                                                                                 \
    if ((threadIdx.x & 31) == 0) {                                                     \
      unsigned long long gt_;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                          \
      eb_dbg_gmax[i] = gt_;                                                            \
    }                                                                                  \
  } while (0)
#else
#define EB_MARK_ANY(i) do { } while (0)
#define EB_MARK(i) do { } while (0)
#define EB_DEFINE_MARK_READER(NAME)
#define EB_DBG_SKIP(bit) false
#endif

// volatile 64-bit global load as inline PTX: the compiler can neither cache it nor fold it with a parameter read
__device__ __forceinline__ unsigned long long ld_volatile_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ---- lazy ladder adaptation (eb_swap_rng.defer_adapt) -------------------------------------------------------------------
// A deferred pass leaves its accepted-swap counts in the rows [(it & 1) * LAZY_SLOTS, +LAZY_SLOTS) of eb_ctrl.swaps_work and
// a snapshot of the ladder, the adaptation clock and the adaptation parameters.  Whoever needs the adapted ladder first —
// every CTA of the next stretch kernel, in its prologue, or the one CTA of eb_adapt_flush — folds the counts and runs
// adapt_temps (tempering.py:563-596; same operands, same operations, same order as pt_swap_adapt) on shared memory.
// `writer` (one CTA) also does the bookkeeping of temper_comps: betas, swaps_accepted / swaps_total, the clock, and it
// zeroes the count rows of the OTHER parity (last read one iteration ago, next written by the next pass).
constexpr int LAZY_SLOTS = 8;
struct LazyShared { double b[EB_MAX_TEMPS]; double d[EB_MAX_TEMPS]; int c[EB_MAX_TEMPS]; };

// all threads of the CTA call this (block barriers inside); returns true if a pass was pending for iteration `it`, and
// then sh.b holds the adapted ladder
// (the work itself is kept out of line: the stretch kernels run at their register limit and only look at the flag)
static __device__ __noinline__ void lazy_adapt_work(eb_ctrl* ctrl, unsigned long long pend, double* betas_global, bool writer,
                                                    bool zero_own, LazyShared& sh, int cnt_mine, double beta_mine) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int T = ctrl->pend_T, W = ctrl->pend_W;
  const unsigned long long itp = pend - 1ull;
  const int row0 = (int)(itp & 1ull) * LAZY_SLOTS;
  if (tid < T) { sh.b[tid] = beta_mine; sh.c[tid] = tid < T - 1 ? cnt_mine : 0; }   // requested ahead of the flag
  for (int r = tid + nt; r < T; r += nt) {                                          // CTAs narrower than the ladder
    sh.b[r] = ctrl->pend_betas[r];
    int v = 0;
    if (r < T - 1) {
      int w8[LAZY_SLOTS];
#pragma unroll
      for (int sl = 0; sl < LAZY_SLOTS; ++sl) w8[sl] = *reinterpret_cast<volatile int*>(&ctrl->swaps_work[row0 + sl][r]);
#pragma unroll
      for (int sl = 0; sl < LAZY_SLOTS; ++sl) v += w8[sl];
    }
    sh.c[r] = v;
  }
  const long long time_now = ctrl->pend_time;
  const bool adapting = ctrl->pend_adapt_on && ctrl->pend_adaptive && T > 1;                     // tempering.py:632-633
  const bool moving = adapting && (ctrl->pend_stop < 0 || time_now < (long long)ctrl->pend_stop);   // :590
  __syncthreads();
  if (moving) {
    const double decay = ctrl->pend_lag / ((double)time_now + ctrl->pend_lag);                   // :571
    const double kappa = decay / ctrl->pend_t0;                                                   // :572
    const double nw = (double)W;
    for (int j = tid; j + 2 < T; j += nt) {
      const double r0 = (double)sh.c[j] / nw, r1 = (double)sh.c[j + 1] / nw;                      // :587
      const double dS = kappa * (r0 - r1);                                                        // :575
      const double dT = 1.0 / sh.b[j + 1] - 1.0 / sh.b[j];                                        // :578
      sh.d[j] = dT * exp(dS);                                                                     // :579
    }
    __syncthreads();
    if (tid == 0) {                                                                               // np.cumsum: sequential adds, in order
      double cum = 0.0;
      for (int j = 0; j + 2 < T; ++j) { cum = cum + sh.d[j]; sh.d[j] = cum; }
    }
    __syncthreads();
    const double inv_b0 = 1.0 / sh.b[0];
    for (int j = tid; j + 2 < T; j += nt) {          // every thread touches its own rungs only (and rung 0, never written)
      const double bold = sh.b[j + 1];
      const double bnew = 1.0 / (sh.d[j] + inv_b0);                                               // :580
      sh.b[j + 1] = bold + (bnew - bold);                                                         // :583, :593
    }
    __syncthreads();
  }
  // the bookkeeping happens once per deferred pass: a second stretch kernel of the same iteration (no pass in between)
  // recomputes the same ladder from the snapshot and leaves the counters alone.  Only writer CTAs read or write
  // adapt_applied, and they run in stream order.
  if (writer && ld_volatile_u64(&ctrl->adapt_applied) != pend) {
    for (int r = tid; r < T; r += nt) betas_global[r] = sh.b[r];
    for (int r = tid; r < T - 1; r += nt) {
      ctrl->swaps_accepted[r] = sh.c[r];
      atomicAdd(reinterpret_cast<unsigned long long*>(&ctrl->swaps_total[r]), (unsigned long long)sh.c[r]);
      const int other = (row0 ^ LAZY_SLOTS);
#pragma unroll
      for (int sl = 0; sl < LAZY_SLOTS; ++sl) ctrl->swaps_work[other + sl][r] = 0;
      if (zero_own) {      // nobody else is reading them (eb_adapt_flush: one CTA)
#pragma unroll
        for (int sl = 0; sl < LAZY_SLOTS; ++sl) ctrl->swaps_work[row0 + sl][r] = 0;
      }
    }
    __syncthreads();   // every thread of the writer has read adapt_applied
    if (tid == 0) {
      if (adapting) ctrl->time = time_now + 1;                                                    // :596
      ctrl->adapt_applied = pend;
    }
  }
}

__device__ __forceinline__ bool lazy_adapt_apply(eb_ctrl* ctrl, unsigned long long it, double* betas_global, bool writer,
                                                 bool zero_own, LazyShared& sh) {
  // the counts and the snapshot ladder of this thread's rung are requested together with the flag (their addresses do not
  // depend on it: a pass pending for iteration `it` has iteration number it - 1), so the fold costs one round trip, not two
  const int tid = threadIdx.x;
  int cnt_mine = 0;
  double beta_mine = 0.0;
  if (tid < EB_MAX_TEMPS) {
    const int row0 = (int)((it - 1ull) & 1ull) * LAZY_SLOTS;
    int w8[LAZY_SLOTS];
#pragma unroll
    for (int sl = 0; sl < LAZY_SLOTS; ++sl) w8[sl] = *reinterpret_cast<volatile int*>(&ctrl->swaps_work[row0 + sl][tid]);
    beta_mine = *reinterpret_cast<volatile double*>(&ctrl->pend_betas[tid]);
#pragma unroll
    for (int sl = 0; sl < LAZY_SLOTS; ++sl) cnt_mine += w8[sl];
  }
  const unsigned long long pend = ld_volatile_u64(&ctrl->adapt_pending);
  if (pend == 0ull || pend != it) return false;        // uniform over the grid: written by a kernel that has completed
  lazy_adapt_work(ctrl, pend, betas_global, writer, zero_own, sh, cnt_mine, beta_mine);
  return true;
}

struct Common {
  double* coords; double* logl; double* logp; uint8_t* inds; double* betas;
  int T, W, L, D, LD;
  int Lb;  // bytes per walker of `inds` (leaf flags, or flags + friend table of a multi-branch state)
  int t0;  // global index of local temperature 0 (random-stream keying)
  const double* lo; const double* hi; const double* lpdf; const double* per;
  const double* like_params; int like_nparams, like_ncomp, like_kind;
};

// stage [lo D][hi D][lpdf D][like params] into shared memory.  The Gaussian functor's precision matrix is staged as
// the packed upper triangle S_ii = P_ii, S_ij = P_ij + P_ji (x^T P x needs D(D+1)/2 products instead of D^2).
// Two steps so that the global-load latency hides behind other work: `stage_load` issues the loads of this thread's
// (first) element into registers, `stage_store` writes them to shared memory and copies whatever exceeds one element
// per thread; the caller synchronises the block afterwards.
struct Staged { double lo, hi, lpdf, per, p0, p1, p2; };
constexpr int PRIOR_ROWS = 4;   // shared-memory rows of D doubles in front of the likelihood block: lo, hi, lpdf, period

__device__ __forceinline__ void stage_store_gauss(double* sp, int D, int e, double pij, double pji) {
  const int i = e / D, j = e - i * D;
  if (j == i) sp[D + sym_row_offset(i, D)] = pij;
  else if (j > i) sp[D + sym_row_offset(i, D) + (j - i)] = pij + pji;
}

__device__ __forceinline__ void stage_load(const Common& c, Staged& r) {
  const int D = c.D, tid = threadIdx.x;
  if (tid < D) { r.lo = c.lo[tid]; r.hi = c.hi[tid]; r.lpdf = c.lpdf[tid]; r.per = c.per ? c.per[tid] : 0.0; }
  if (c.like_kind == EB_LIKE_GAUSSIAN) {
    if (tid < D) r.p0 = c.like_params[tid];
    if (tid < D * D) {
      const int i = tid / D, j = tid - i * D;
      const double* P = c.like_params + D;
      r.p1 = P[tid];
      r.p2 = (j > i) ? P[j * D + i] : 0.0;
    }
  } else if (tid < c.like_nparams) {
    r.p0 = c.like_params[tid];
  }
}

__device__ __forceinline__ void stage_store(const Common& c, const Staged& r, double* sm) {
  const int D = c.D, tid = threadIdx.x, nt = blockDim.x;
  if (tid < D) { sm[tid] = r.lo; sm[D + tid] = r.hi; sm[2 * D + tid] = r.lpdf; sm[3 * D + tid] = r.per; }
  for (int i = tid + nt; i < D; i += nt) {
    sm[i] = c.lo[i]; sm[D + i] = c.hi[i]; sm[2 * D + i] = c.lpdf[i]; sm[3 * D + i] = c.per ? c.per[i] : 0.0;
  }
  double* sp = sm + PRIOR_ROWS * D;
  if (c.like_kind == EB_LIKE_GAUSSIAN) {
    const double* P = c.like_params + D;
    if (tid < D) sp[tid] = r.p0;
    for (int i = tid + nt; i < D; i += nt) sp[i] = c.like_params[i];
    if (tid < D * D) stage_store_gauss(sp, D, tid, r.p1, r.p2);
    for (int e = tid + nt; e < D * D; e += nt) {
      const int i = e / D, j = e - i * D;
      stage_store_gauss(sp, D, e, P[e], P[j * D + i]);
    }
  } else {
    if (tid < c.like_nparams) sp[tid] = r.p0;
    for (int i = tid + nt; i < c.like_nparams; i += nt) sp[i] = c.like_params[i];
  }
}

__device__ __forceinline__ void stage_params(const Common& c, double* sm) {
  Staged r;
  stage_load(c, r);
  stage_store(c, r, sm);
  __syncthreads();
}

// log-prior (single leaf, L == 1) and gated log-like of a proposed point
template <int DMAX, int LIKE, bool EXACT>
__device__ __forceinline__ void eval_point(const double (&q)[DMAX], const Common& c, const double* sm, bool leaf_active,
                                           double& lp, double& ll) {
  const int D = EXACT ? DMAX : c.D;
  lp = leaf_active ? box_logpdf_leaf<DMAX, EXACT>(q, D, sm, sm + D, sm + 2 * D) : 0.0;
  if (isinf(lp) || !leaf_active) {
    ll = FILL_LOGL;  // ensemble.py:1279-1282, :1486 / fill_zero_leaves_val :1499
  } else {
    ll = Like<LIKE>::template eval<DMAX, EXACT>(q, D, sm + PRIOR_ROWS * D, c.like_ncomp);
    if (ll != ll) ll = FILL_LOGL;  // red_blue.py:279-281
  }
}

// NumPy's float `%` (npy_divmod): fmod, then the sign of the divisor
__device__ __forceinline__ double np_mod(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) {
    if ((b < 0.0) != (m < 0.0)) m += b;
  } else {
    m = copysign(0.0, b);
  }
  return m;
}

int fill_common(Common& c, const eb_state* st, const eb_prior* prior, const eb_like* like, bool need_fused);

__host__ __device__ inline size_t smem_bytes(const Common& c, int extra = 0) {
  return sizeof(double) * (size_t)(PRIOR_ROWS * c.D + c.like_nparams + extra);
}

static inline int bucket(int LD) { return LD <= 8 ? 8 : LD <= 16 ? 16 : LD <= 24 ? 24 : 32; }

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 200 * 1024) return fail(EB_ERR_UNSUPPORTED, "parameter block of %zu bytes does not fit shared memory", bytes);
    EB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return EB_OK;
}

// MACRO(DMAX, EXACT): exact-length kernels for the BASELINE row lengths (8, 20), padded buckets otherwise
#define EB_DISPATCH_DMAX(LD, MACRO)              \
  switch (LD) {                                  \
    case 8: MACRO(8, true); break;               \
    case 20: MACRO(20, true); break;             \
    default:                                     \
      switch (bucket(LD)) {                      \
        case 8: MACRO(8, false); break;          \
        case 16: MACRO(16, false); break;        \
        case 24: MACRO(24, false); break;        \
        default: MACRO(32, false); break;        \
      }                                          \
  }
// padded buckets only (kernels off the hot path)
#define EB_DISPATCH_DMAX_GENERIC(LD, MACRO)      \
  switch (bucket(LD)) {                          \
    case 8: MACRO(8, false); break;              \
    case 16: MACRO(16, false); break;            \
    case 24: MACRO(24, false); break;            \
    default: MACRO(32, false); break;            \
  }
#define EB_DISPATCH_LIKE(KIND, MACRO2)       \
  switch (KIND) {                            \
    case 0: MACRO2(0); break;                \
    case 1: MACRO2(1); break;                \
    default: MACRO2(2); break;               \
  }

}  // namespace eb
