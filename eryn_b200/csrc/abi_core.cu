// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include "common.cuh"

namespace eb {

static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(EB_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return EB_OK;
}


// ================================================================================================
// K0: evaluate the whole state
// ================================================================================================
template <int DMAX, int LIKE>
__global__ void __launch_bounds__(BLOCK) eval_state_kernel(const Common c) {
  extern __shared__ double sm[];
  stage_params(c, sm);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * c.W) return;
  double x[DMAX];
  load_row<DMAX>(c.coords + (size_t)tid * c.LD, c.LD, x);
  const bool active = c.inds ? (c.inds[tid] != 0) : true;
  double lp, ll;
  eval_point<DMAX, LIKE, false>(x, c, sm, active, lp, ll);
  c.logp[tid] = lp;
  c.logl[tid] = ll;
}

__global__ void __launch_bounds__(BLOCK) box_prior_kernel(const double* __restrict__ q, const uint8_t* __restrict__ inds,
                                                          int nrows, int L, int D, const double* __restrict__ lo,
                                                          const double* __restrict__ hi,
                                                          const double* __restrict__ lpdf, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double tot = 0.0;
  for (int l = 0; l < L; ++l) {
    if (inds && !inds[(size_t)r * L + l]) continue;
    const double* x = q + ((size_t)r * L + l) * D;
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double v = x[d];
      double t = 0.0;
      if (v >= lo[d] && v <= hi[d]) t = lpdf[d];
      if (v < lo[d] || v > hi[d]) t = neg_inf();
      s += t;
    }
    tot += s;
  }
  out[r] = tot;
}

// A stretch kernel launched as a programmatic dependent reads iter_next BEFORE its grid-dependency wait
// (eb_stretch_rng.pdl_chain), so the writer publishes it like the swap pass does: store, device-scope fence, trigger.
__global__ void advance_iter_kernel(eb_ctrl* ctrl) {
  const unsigned long long next = ctrl->iter + 1ull;
  ctrl->iter = next;
  *reinterpret_cast<volatile unsigned long long*>(&ctrl->iter_next) = next;
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

int fill_common(Common& c, const eb_state* st, const eb_prior* prior, const eb_like* like, bool need_fused) {
  if (!st) return fail(EB_ERR_INVALID, "state is NULL");
  if (st->ntemps < 1 || st->nwalkers < 1 || st->nleaves < 1 || st->ndim < 1)
    return fail(EB_ERR_INVALID, "incompatible input dimensions (T=%d W=%d L=%d D=%d)", st->ntemps, st->nwalkers,
                st->nleaves, st->ndim);
  if (st->ntemps > EB_MAX_TEMPS) return fail(EB_ERR_UNSUPPORTED, "ntemps %d > %d", st->ntemps, EB_MAX_TEMPS);
  if (!st->coords || !st->logl || !st->logp) return fail(EB_ERR_INVALID, "coords/logl/logp must be device pointers");
  c.coords = st->coords; c.logl = st->logl; c.logp = st->logp;
  c.inds = const_cast<uint8_t*>(st->inds); c.betas = st->betas;
  c.T = st->ntemps; c.W = st->nwalkers; c.L = st->nleaves; c.D = st->ndim; c.LD = st->nleaves * st->ndim;
  c.t0 = st->temp_offset;
  c.Lb = st->inds_stride > 0 ? st->inds_stride : st->nleaves;
  c.lo = c.hi = c.lpdf = nullptr; c.per = nullptr; c.like_params = nullptr; c.like_nparams = 0; c.like_ncomp = 0;
  if (prior) { c.lo = prior->lo; c.hi = prior->hi; c.lpdf = prior->logpdf; c.per = prior->period; }
  c.like_kind = -1;
  if (like) { c.like_params = like->params; c.like_nparams = like->nparams; c.like_ncomp = like->ncomp; c.like_kind = like->kind; }
  if (need_fused) {
    if (!prior || !prior->lo || !prior->hi || !prior->logpdf) return fail(EB_ERR_INVALID, "prior is required");
    if (!like) return fail(EB_ERR_INVALID, "likelihood functor is required");
    if (st->nleaves != 1)
      return fail(EB_ERR_UNSUPPORTED, "fused kernels need nleaves == 1 (got %d); use the split path", st->nleaves);
    if (c.LD > EB_MAX_ROW)
      return fail(EB_ERR_UNSUPPORTED, "ndim %d > %d not covered by the fused kernels; use the split path", c.LD,
                  EB_MAX_ROW);
    const int D = c.D;
    int need = 0;
    switch (like->kind) {
      case EB_LIKE_GAUSSIAN: need = D + D * D; break;
      case EB_LIKE_ROSENBROCK: need = 0; break;
      case EB_LIKE_GMIX:
        if (like->ncomp < 1) return fail(EB_ERR_INVALID, "EB_LIKE_GMIX needs ncomp >= 1");
        need = like->ncomp * (2 + D);
        break;
      default: return fail(EB_ERR_INVALID, "unknown likelihood kind %d", like->kind);
    }
    if (like->nparams != need) return fail(EB_ERR_INVALID, "likelihood kind %d expects %d params, got %d", like->kind, need, like->nparams);
    if (need > 0 && !like->params) return fail(EB_ERR_INVALID, "likelihood params pointer is NULL");
  }
  return EB_OK;
}

template <int DMAX, int LIKE>
static int launch_eval(const Common& c, cudaStream_t s) {
  const int n = c.T * c.W;
  const size_t sb = smem_bytes(c);
  int rc = set_smem(eval_state_kernel<DMAX, LIKE>, sb);
  if (rc) return rc;
  eval_state_kernel<DMAX, LIKE><<<(n + BLOCK - 1) / BLOCK, BLOCK, sb, s>>>(c);
  return EB_OK;
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_abi_version(void) { return EB_ABI_VERSION; }
const char* eb_last_error(void) { return g_err; }
size_t eb_ctrl_size(void) { return sizeof(eb_ctrl); }
size_t eb_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(eb_state);
    case 1: return sizeof(eb_prior);
    case 2: return sizeof(eb_like);
    case 3: return sizeof(eb_stretch_rng);
    case 4: return sizeof(eb_gauss_rng);
    case 5: return sizeof(eb_swap_rng);
    case 6: return sizeof(eb_ctrl);
    case 7: return sizeof(eb_adapt);
    case 8: return sizeof(eb_host_job);
    case 9: return sizeof(eb_shard);
    case 10: return sizeof(eb_publish);
    case 11: return sizeof(eb_mb_layout);
    case 12: return sizeof(eb_mb_state);
    case 13: return sizeof(eb_pulse_data);
    case 14: return sizeof(eb_mb_friends);
    case 15: return sizeof(eb_mb_group_rng);
    case 16: return sizeof(eb_mb_rj_rng);
    case 17: return sizeof(eb_split);
    case 18: return sizeof(eb_stage);
    case 19: return sizeof(eb_mt_rng);
    default: return 0;
  }
}

int eb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int eb_eval_state(const eb_state* st, const eb_prior* prior, const eb_like* like, void* stream) {
  Common c;
  int rc = fill_common(c, st, prior, like, true);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
#define L2_(K) rc = launch_eval<DM_, K>(c, s)
#define L1_(DM, EX)                          \
  {                                          \
    constexpr int DM_ = DM;                  \
    EB_DISPATCH_LIKE(like->kind, L2_)        \
  }
  EB_DISPATCH_DMAX_GENERIC(c.LD, L1_)
#undef L1_
#undef L2_
  if (rc) return rc;
  return check_launch("eval_state");
}

int eb_advance_iter(eb_ctrl* ctrl, void* stream) {
  if (!ctrl) return fail(EB_ERR_INVALID, "ctrl is NULL");
  advance_iter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctrl);
  return check_launch("advance_iter");
}

int eb_box_log_prior(const double* q, const uint8_t* inds_sub, int32_t nrows, int32_t nleaves, int32_t ndim,
                     const eb_prior* prior, double* logp_out, void* stream) {
  if (!q || !prior || !logp_out || nrows < 1) return fail(EB_ERR_INVALID, "NULL/empty argument");
  box_prior_kernel<<<(nrows + BLOCK - 1) / BLOCK, BLOCK, 0, (cudaStream_t)stream>>>(q, inds_sub, nrows, nleaves, ndim,
                                                                                     prior->lo, prior->hi, prior->logpdf,
                                                                                     logp_out);
  return check_launch("box_prior");
}

}  // extern "C"
