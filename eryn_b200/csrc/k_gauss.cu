// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include "common.cuh"

namespace eb {

// ================================================================================================
// K2: fused Gaussian Metropolis step (all walkers at once, mh.py:56-193)
// ================================================================================================
struct GaussArgs {
  Common c;
  int cov_kind; double scale; const double* chol; const double* delta; const double* u_acc;
  int philox;
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  uint8_t* accepted; uint32_t* accepted_count;
  uint32_t gmask; int gidx;   // Gibbs split (mh.py:77-183): parameters that change (0 = all), index of the split
  int dim_mode; double log_factor;   // gaussian.py:134-181: 1 = one random dimension per walker; scale factor exp(U(-lf, lf))
  eb_ctrl* lazy_ctrl;   // a pass that deferred its ladder adaptation (common.cuh:lazy_adapt_apply)
};

template <int DMAX, int LIKE, bool PHILOX, bool EXACT>
__global__ void __launch_bounds__(BLOCK) gaussian_step_kernel(const GaussArgs p) {
  extern __shared__ __align__(16) double sm[];
  const Common& c = p.c;
  stage_params(c, sm);
  const int D = EXACT ? DMAX : c.D;
  double* s_chol = sm + PRIOR_ROWS * D + c.like_nparams;
  if (PHILOX && p.cov_kind == 1) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) s_chol[i] = p.chol[i];
    __syncthreads();
  }
  // a pass that deferred its ladder adaptation: every CTA folds the counts and adapts the ladder itself, CTA 0 does the
  // bookkeeping (all threads take part: block barriers inside)
  __shared__ LazyShared lazy_sh;
  bool lazy = false;
  if (PHILOX && p.lazy_ctrl)
    lazy = lazy_adapt_apply(p.lazy_ctrl, p.iter_dev ? ld_volatile_u64(p.iter_dev) : p.iter, c.betas, blockIdx.x == 0, false, lazy_sh);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * c.W) return;
  const int t = tid / c.W;
  double q[DMAX];
  load_row<DMAX>(c.coords + (size_t)tid * D, D, q);
  const bool active = c.inds ? (c.inds[tid] != 0) : true;
  double u_acc;
  double factors = 0.0;                                                      // gaussian.py:131
  if (p.cov_kind == 2 && active)                                             // distgen.py:96: + log q(old)
    factors += +1.0 * box_logpdf_leaf<DMAX, EXACT>(q, D, sm, sm + D, sm + 2 * D);
  if (PHILOX && p.cov_kind == 2) {
    // DistributionGenerate: every active leaf is redrawn from the (uniform) priors, distgen.py:99 / prior.py:66
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
#pragma unroll
    for (int j = 0; j < DMAX; j += 2) {
      if (EXACT || j < D) {
        const uint4 r = stream(key, TAG_GAUSS, (uint32_t)(tid + c.t0 * c.W), (uint32_t)(j >> 1));
        if (active) {
          q[j] = u01_52(r.x, r.y) * (sm[D + j] - sm[j]) + sm[j];
          if (j + 1 < DMAX && (EXACT || j + 1 < D)) q[j + 1] = u01_52(r.z, r.w) * (sm[D + j + 1] - sm[j + 1]) + sm[j + 1];
        }
      }
    }
    const uint4 ra = stream(key, TAG_ACCEPT, (uint32_t)(tid + c.t0 * c.W), 0u);
    u_acc = u01_52(ra.x, ra.y);
  } else if (PHILOX) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
    double z[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; j += 2) {
      if (EXACT || j < D) {
        const uint4 r = stream(key, TAG_GAUSS, (uint32_t)(tid + c.t0 * c.W), (uint32_t)(j >> 1) | ((uint32_t)p.gidx << 16));
        const double rad = sqrt(-2.0 * log(u01_52(r.x, r.y)));
        double sn, cs;
        sincos(6.283185307179586 * u01_52(r.z, r.w), &sn, &cs);
        z[j] = rad * cs;
        z[j + 1] = rad * sn;
      } else {
        z[j] = 0.0; z[j + 1] = 0.0;
      }
    }
    const uint4 ra = stream(key, TAG_ACCEPT, (uint32_t)(tid + c.t0 * c.W), (uint32_t)p.gidx << 16);
    u_acc = u01_52(ra.x, ra.y);
    double f = 1.0;
    if (p.log_factor > 0.0) {   // get_factor (gaussian.py:161-164): one draw per call, the same for every walker
      const uint4 rf = stream(key, TAG_GAUSS, 0xFFFFFFFFu, 0xFFFFu | ((uint32_t)p.gidx << 16));
      f = exp(-p.log_factor + (p.log_factor - (-p.log_factor)) * u01_52(rf.x, rf.y));   // rng.uniform(-lf, lf)
    }
    if (active) {
      if (p.cov_kind == 0) {
        const double fs = f * p.scale;
        const int dsel = p.dim_mode == 1 ? (int)__umulhi(ra.z, (uint32_t)D) : -1;       // gaussian.py:172-173
#pragma unroll
        for (int j = 0; j < DMAX; ++j)
          if ((EXACT || j < D) && (dsel < 0 || j == dsel)) q[j] = q[j] + fs * z[j];     // gaussian.py:166-167
      } else {
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
          if (i < D) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < DMAX; ++j)
              if (j <= i && j < D) acc += s_chol[i * D + j] * z[j];
            q[i] = q[i] + f * acc;                                           // gaussian.py:192-195
          }
      }
    }
  } else {
    const double* dl = p.delta + (size_t)tid * D;
    if (active) {
#pragma unroll
      for (int j = 0; j < DMAX; ++j)
        if (EXACT || j < D) q[j] = p.cov_kind == 2 ? dl[j] : q[j] + dl[j];   // distgen: dl is the new point itself
    }
    u_acc = p.u_acc[tid];
  }
  if (c.per && active && p.cov_kind != 2) {                                  // gaussian.py:111-129 (distgen.py does not wrap)
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if ((EXACT || j < D) && sm[3 * D + j] > 0.0) q[j] = np_mod(q[j], sm[3 * D + j]);
  }
  if (p.gmask) {   // Gibbs split: the other parameters keep their values (cleanup_proposals_gibbs, move.py:302-307)
    const double* old = c.coords + (size_t)tid * D;
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if ((EXACT || j < D) && !((p.gmask >> j) & 1u)) q[j] = old[j];
  }
  const double ll0 = c.logl[tid], lp0 = c.logp[tid];
  const bool tempered = c.betas != nullptr;
  const double beta = tempered ? (lazy ? lazy_sh.b[c.t0 + t] : c.betas[t]) : 1.0;
  double lp, ll;
  eval_point<DMAX, LIKE, EXACT>(q, c, sm, active, lp, ll);                          // mh.py:134-148
  const double logP = log_posterior(ll, lp, beta, tempered);
  const double prevP = log_posterior(ll0, lp0, beta, tempered);
  if (p.cov_kind == 2 && active)                                             // distgen.py:102: - log q(new)
    factors += -1.0 * box_logpdf_leaf<DMAX, EXACT>(q, D, sm, sm + D, sm + 2 * D);
  const bool keep = (factors + logP - prevP) > log(u_acc);                   // mh.py:168-171
  if (keep) {
    store_row<DMAX>(c.coords + (size_t)tid * D, D, q);
    c.logl[tid] = ll;
    c.logp[tid] = isinf(lp) ? 0.0 : lp;
  }
  p.accepted[tid] = keep ? 1 : 0;
  if (p.accepted_count && keep) p.accepted_count[tid] += 1u;
}

template <int DMAX, int LIKE, bool EXACT>
static int launch_gauss(const GaussArgs& a, cudaStream_t s) {
  const int n = a.c.T * a.c.W;
  const size_t sb = smem_bytes(a.c, a.c.D * a.c.D);
  if (a.philox) {
    int rc = set_smem(gaussian_step_kernel<DMAX, LIKE, true, EXACT>, sb);
    if (rc) return rc;
    gaussian_step_kernel<DMAX, LIKE, true, EXACT><<<(n + BLOCK - 1) / BLOCK, BLOCK, sb, s>>>(a);
  } else {
    int rc = set_smem(gaussian_step_kernel<DMAX, LIKE, false, EXACT>, sb);
    if (rc) return rc;
    gaussian_step_kernel<DMAX, LIKE, false, EXACT><<<(n + BLOCK - 1) / BLOCK, BLOCK, sb, s>>>(a);
  }
  return EB_OK;
}

// one object file per likelihood kind (build.py compiles this source with -DEB_ONLY_LIKE=k)
template <int LIKE>
int launch_gauss_like(const GaussArgs& a, cudaStream_t s);
#ifndef EB_ONLY_LIKE
#define EB_ONLY_LIKE -1
#endif
#define EB_GAUSS_LIKE_DEF(K)                                      \
  template <>                                                     \
  int launch_gauss_like<K>(const GaussArgs& a, cudaStream_t s) {  \
    int rc = EB_OK;                                               \
    EB_DISPATCH_DMAX(a.c.LD, EB_GAUSS_L1_##K)                     \
    return rc;                                                    \
  }
#define EB_GAUSS_L1_0(DM, EX) rc = launch_gauss<DM, 0, EX>(a, s)
#define EB_GAUSS_L1_1(DM, EX) rc = launch_gauss<DM, 1, EX>(a, s)
#define EB_GAUSS_L1_2(DM, EX) rc = launch_gauss<DM, 2, EX>(a, s)
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
EB_GAUSS_LIKE_DEF(0)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 1
EB_GAUSS_LIKE_DEF(1)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 2
EB_GAUSS_LIKE_DEF(2)
#endif

}  // namespace eb

#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
using namespace eb;

extern "C" {

int eb_gaussian_step(const eb_state* st, const eb_prior* prior, const eb_like* like, const eb_gauss_rng* rng,
                     uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  GaussArgs args;
  int rc = fill_common(args.c, st, prior, like, true);
  if (rc) return rc;
  if (!rng) return fail(EB_ERR_INVALID, "rng is NULL");
  if (!accepted) return fail(EB_ERR_INVALID, "accepted is NULL");
  args.cov_kind = rng->cov_kind; args.scale = rng->scale; args.chol = rng->chol; args.delta = rng->delta;
  args.u_acc = rng->u_acc; args.philox = rng->mode == EB_RNG_PHILOX;
  args.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); args.seed_hi = (uint32_t)(rng->seed >> 32);
  args.iter_dev = (const unsigned long long*)rng->iter_dev; args.iter = rng->iter;
  args.accepted = accepted; args.accepted_count = accepted_count;
  args.gmask = rng->gibbs_mask; args.gidx = rng->gibbs_index;
  args.dim_mode = rng->dim_mode; args.log_factor = rng->log_factor;
  args.lazy_ctrl = (args.philox && st->temp_offset == 0 && st->betas) ? (eb_ctrl*)rng->lazy_ctrl : nullptr;
  if (args.dim_mode != 0 && (args.dim_mode != 1 || rng->cov_kind != 0))
    return fail(EB_ERR_INVALID, "dim_mode must be 0 (vector) or, for scalar proposals, 1 (random)");
  if (args.log_factor < 0.0) return fail(EB_ERR_INVALID, "'factor' must be >= 1.0");
  if (args.gmask && rng->cov_kind == 2) return fail(EB_ERR_UNSUPPORTED, "DistributionGenerate has no Gibbs splits on the device");
  if (args.gmask && st->nleaves != 1) return fail(EB_ERR_UNSUPPORTED, "Gibbs splits address the parameters of one leaf");
  if (args.gidx < 0 || args.gidx > 0xFFFF) return fail(EB_ERR_INVALID, "gibbs_index out of range");
  if (args.philox) {
    if (rng->cov_kind == 1 && !rng->chol) return fail(EB_ERR_INVALID, "matrix proposal needs the Cholesky factor");
    if (rng->cov_kind < 0 || rng->cov_kind > 2) return fail(EB_ERR_INVALID, "Invalid proposal scale dimensions");
  } else if (rng->mode == EB_RNG_REPLAY) {
    if (!rng->delta || !rng->u_acc) return fail(EB_ERR_INVALID, "replay mode needs delta and u_acc");
  } else {
    return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (like->kind) {
    case 0: rc = launch_gauss_like<0>(args, s); break;
    case 1: rc = launch_gauss_like<1>(args, s); break;
    default: rc = launch_gauss_like<2>(args, s); break;
  }
  if (rc) return rc;
  return check_launch("gaussian_step");
}

}  // extern "C"
#endif  // EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
