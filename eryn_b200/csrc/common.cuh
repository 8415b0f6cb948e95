// eryn_b200 — hand-written sm_100a kernels of the walker-parallel sampling hot path + C ABI.
//
// Kernels:
//   K0  eval_state_kernel         log-prior + log-like of the whole state            (abi_core.cu)
//   K1  stretch_step_kernel       fused, both red/blue halves per launch (cluster per temperature):
//                                 draw -> gather complement -> stretch -> prior/like ->
//                                 tempered Metropolis test -> in-place update        (k_stretch.cu, hot kernel)
//   K2  gaussian_step_kernel      fused Gaussian Metropolis step over all walkers    (k_gauss.cu)
//   K3  pt_swap_kernel            chain-parallel swap ladder (decide on logl, then move only the rows
//                                 that changed rung, in place) + last-block ladder adaptation (k_swap.cu)
//   K3r pt_pairmap_kernel         replay mode: host permutations -> per-position pair map
//   K4  stretch_propose_kernel / accept_update_kernel / box_prior_kernel   (split path)
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
static __device__ long long eb_dbg_marks[64];   // one copy per translation unit (no relocatable device code)
#define EB_DEFINE_MARK_READER(NAME)                                                                   \
  extern "C" __attribute__((visibility("default"))) int NAME(long long* out_host) {                  \
    return cudaMemcpyFromSymbol(out_host, eb_dbg_marks, sizeof(long long) * 64) == cudaSuccess ? 0 : 3; \
  }
#define EB_MARK(i)                                                                     \
  do {                                                                                 \
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) eb_dbg_marks[i] = clock64(); \
  } while (0)
#else
#define EB_MARK(i) do { } while (0)
#define EB_DEFINE_MARK_READER(NAME)
#endif

struct Common {
  double* coords; double* logl; double* logp; uint8_t* inds; double* betas;
  int T, W, L, D, LD;
  int t0;  // global index of local temperature 0 (random-stream keying)
  const double* lo; const double* hi; const double* lpdf;
  const double* like_params; int like_nparams, like_ncomp;
};

// stage [lo D][hi D][lpdf D][like params] into shared memory
__device__ __forceinline__ void stage_params(const Common& c, double* sm) {
  const int D = c.D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    sm[i] = c.lo[i]; sm[D + i] = c.hi[i]; sm[2 * D + i] = c.lpdf[i];
  }
  for (int i = threadIdx.x; i < c.like_nparams; i += blockDim.x) sm[3 * D + i] = c.like_params[i];
  __syncthreads();
}

// log-prior (single leaf, L == 1) and gated log-like of a proposed point
template <int DMAX, int LIKE>
__device__ __forceinline__ void eval_point(const double (&q)[DMAX], const Common& c, const double* sm, bool leaf_active,
                                           double& lp, double& ll) {
  const int D = c.D;
  lp = leaf_active ? box_logpdf_leaf<DMAX>(q, 0, D, sm, sm + D, sm + 2 * D) : 0.0;
  if (isinf(lp) || !leaf_active) {
    ll = FILL_LOGL;  // ensemble.py:1279-1282, :1486 / fill_zero_leaves_val :1499
  } else {
    ll = Like<LIKE>::template eval<DMAX>(q, D, sm + 3 * D, c.like_ncomp);
    if (ll != ll) ll = FILL_LOGL;  // red_blue.py:279-281
  }
}

int fill_common(Common& c, const eb_state* st, const eb_prior* prior, const eb_like* like, bool need_fused);

static inline size_t smem_bytes(const Common& c, int extra = 0) {
  return sizeof(double) * (size_t)(3 * c.D + c.like_nparams + extra);
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

#define EB_DISPATCH_DMAX(LD, MACRO)          \
  switch (bucket(LD)) {                      \
    case 8: MACRO(8); break;                 \
    case 16: MACRO(16); break;               \
    case 24: MACRO(24); break;               \
    default: MACRO(32); break;               \
  }
#define EB_DISPATCH_LIKE(KIND, MACRO2)       \
  switch (KIND) {                            \
    case 0: MACRO2(0); break;                \
    case 1: MACRO2(1); break;                \
    default: MACRO2(2); break;               \
  }

}  // namespace eb
