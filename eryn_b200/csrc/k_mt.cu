// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
//
// K11: multiple-try Metropolis with an independent proposal — MTDistGenMove(generate_dist = priors, num_try,
// independent=True): moves/multipletry.py:238-514 + moves/mtdistgen.py:8-133 inside MHMove.propose (mh.py:56-193).
//
// One thread owns one walker.  Per walker: num_try points are drawn from the (uniform box) priors; every try gets its
// log-prior, log-likelihood (NOT gated by the prior: the reference passes no logp, mtdistgen.py:117-121) and importance
// weight log w_j = beta logL_j + logp_j - log q(y_j); one try is picked with probability w_j / sum w by inverting the
// cumulative sum against ONE uniform (multipletry.py:49-53); the auxiliary set of an independent proposal is the same
// tries with the current point in place of the picked one (:383-416); the detailed-balance factor is written exactly as
// the reference writes it (:467-471) and the move finishes as a Metropolis step on the stored (mt_ll, mt_lp) values
// (mh.py:146-183).  Only log w_j is kept per try (a thread-local array); the picked point is regenerated from its
// counter (philox) or re-read (replay).  The sums of the log-sum-exp run in index order (NumPy's pairwise blocking for
// more than 8 tries differs in the last bits: tolerance 1e-10).
#include "common.cuh"

namespace eb {

constexpr uint32_t TAG_MT = 9;
constexpr int MT_MAX_TRY = 128;

struct MTArgs {
  Common c;
  int num_try, philox;
  const double* tries; const double* u_sel; const double* u_acc;   // replay [T][W][NT][D], [T][W], [T][W]
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  uint8_t* accepted; uint32_t* accepted_count;
};

// try j of flat walker `fw` into y[]
template <int DMAX, bool PHILOX>
__device__ __forceinline__ void mt_point(const MTArgs& p, const RngKey& key, const double* sm, int D, uint32_t fw, size_t tid,
                                         int j, double (&y)[DMAX]) {
  if (PHILOX) {
#pragma unroll
    for (int d = 0; d < DMAX; d += 2) {
      if (d < D) {
        const uint4 r = stream(key, TAG_MT, fw, (uint32_t)(j * 16 + (d >> 1)));
        y[d] = u01_52(r.x, r.y) * (sm[D + d] - sm[d]) + sm[d];                       // prior.py:66
        if (d + 1 < DMAX && d + 1 < D) y[d + 1] = u01_52(r.z, r.w) * (sm[D + d + 1] - sm[d + 1]) + sm[d + 1];
      }
    }
  } else {
    const double* src = p.tries + (tid * p.num_try + j) * D;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (d < D) y[d] = src[d];
  }
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d >= D) y[d] = 0.0;
}

template <int DMAX, int LIKE, bool PHILOX>
__global__ void __launch_bounds__(BLOCK) mt_distgen_kernel(const MTArgs p) {
  extern __shared__ __align__(16) double sm[];
  const Common& c = p.c;
  stage_params(c, sm);
  const int D = c.D, NT = p.num_try;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)c.T * c.W) return;
  const int t = (int)(tid / c.W);
  const uint32_t fw = (uint32_t)(tid + (size_t)c.t0 * c.W);
  RngKey key;
  double u_sel, u_acc;
  if (PHILOX) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    const uint4 rs = stream(key, TAG_MT, fw, 0xFFFFu);
    u_sel = u01_52(rs.x, rs.y);
    const uint4 ra = stream(key, TAG_ACCEPT, fw, 0u);
    u_acc = u01_52(ra.x, ra.y);
  } else {
    u_sel = p.u_sel[tid];
    u_acc = p.u_acc[tid];
  }
  const bool tempered = c.betas != nullptr;
  const double beta = tempered ? c.betas[t] : 1.0;
  const double* sp = sm + PRIOR_ROWS * D;
  double lw[MT_MAX_TRY];
  double y[DMAX];
  // ---- every try: prior, likelihood, importance weight (multipletry.py:319-356)
  for (int j = 0; j < NT; ++j) {
    mt_point<DMAX, PHILOX>(p, key, sm, D, fw, tid, j, y);
    const double lpp = box_logpdf_leaf<DMAX, false>(y, D, sm, sm + D, sm + 2 * D);   // log q(y_j): the generating density
    double ll = Like<LIKE>::template eval<DMAX, false>(y, D, sp, c.like_ncomp);
    if (ll != ll) ll = FILL_LOGL;                                                    // :339-341
    const double lp = lpp;                                                           // generate_dist is the prior
    lw[j] = (beta * ll + lp) - lpp;                                                  // :354, get_mt_computations :41
  }
  double mx = lw[0];
  for (int j = 1; j < NT; ++j) mx = lw[j] > mx ? lw[j] : mx;
  double se = 0.0;
  for (int j = 0; j < NT; ++j) se += exp(lw[j] - mx);
  const double lsw = mx + log(se);                                                   // logsumexp, :25-31
  int pick = 0;
  {
    double cum = 0.0;
    bool found = false;
    for (int j = 0; j < NT; ++j) {
      cum += exp(lw[j] - lsw);                                                       // probs.cumsum(1), :49-53
      if (!found && cum > u_sel) { pick = j; found = true; }
    }
  }
  // ---- the picked try (regenerated) and the current point
  mt_point<DMAX, PHILOX>(p, key, sm, D, fw, tid, pick, y);
  const double lpp_out = box_logpdf_leaf<DMAX, false>(y, D, sm, sm + D, sm + 2 * D);
  double ll_out = Like<LIKE>::template eval<DMAX, false>(y, D, sp, c.like_ncomp);
  if (ll_out != ll_out) ll_out = FILL_LOGL;
  const double lp_out = lpp_out;
  const double logP_out = beta * ll_out + lp_out;
  double x0[DMAX];
  load_row<DMAX>(c.coords + tid * D, D, x0);
  const double ll0 = c.logl[tid], lp0 = c.logp[tid];
  const double aux_lpp_out = box_logpdf_leaf<DMAX, false>(x0, D, sm, sm + D, sm + 2 * D);   // special_generate_logpdf(coords), :390
  const double aux_logP_out = beta * ll0 + lp0;                                             // :407
  // auxiliary weights: the tries with the current point in place of the picked one (:383-416)
  lw[pick] = aux_logP_out - aux_lpp_out;
  mx = lw[0];
  for (int j = 1; j < NT; ++j) mx = lw[j] > mx ? lw[j] : mx;
  se = 0.0;
  for (int j = 0; j < NT; ++j) se += exp(lw[j] - mx);
  const double aux_lsw = mx + log(se);
  const double factors = ((aux_logP_out - aux_lsw) - aux_lpp_out + aux_lpp_out)
                         - ((logP_out - lsw) - lpp_out + lpp_out);                          // :467-471
  // ---- Metropolis step on the stored values (mh.py:146-183)
  const double logP = log_posterior(ll_out, lp_out, beta, tempered);
  const double prevP = log_posterior(ll0, lp0, beta, tempered);
  const bool keep = (factors + logP - prevP) > log(u_acc);
  if (keep) {
    store_row<DMAX>(c.coords + tid * D, D, y);
    c.logl[tid] = ll_out;
    c.logp[tid] = isinf(lp_out) ? 0.0 : lp_out;
    if (p.accepted_count) p.accepted_count[tid] += 1u;
  }
  p.accepted[tid] = keep ? 1 : 0;
}

template <int DMAX, int LIKE>
static int launch_mt(const MTArgs& a, cudaStream_t s) {
  const size_t n = (size_t)a.c.T * a.c.W;
  const size_t sb = smem_bytes(a.c);
  const unsigned grid = (unsigned)((n + BLOCK - 1) / BLOCK);
  if (a.philox) {
    int rc = set_smem(mt_distgen_kernel<DMAX, LIKE, true>, sb);
    if (rc) return rc;
    mt_distgen_kernel<DMAX, LIKE, true><<<grid, BLOCK, sb, s>>>(a);
  } else {
    int rc = set_smem(mt_distgen_kernel<DMAX, LIKE, false>, sb);
    if (rc) return rc;
    mt_distgen_kernel<DMAX, LIKE, false><<<grid, BLOCK, sb, s>>>(a);
  }
  return EB_OK;
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_mt_distgen_step(const eb_state* st, const eb_prior* prior, const eb_like* like, const eb_mt_rng* rng,
                       uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  MTArgs a;
  int rc = fill_common(a.c, st, prior, like, true);
  if (rc) return rc;
  if (!rng || !accepted) return fail(EB_ERR_INVALID, "rng/accepted is NULL");
  if (st->inds) return fail(EB_ERR_UNSUPPORTED, "multiple try works on one present leaf per walker (multipletry.py:548)");
  if (rng->num_try < 1 || rng->num_try > MT_MAX_TRY) return fail(EB_ERR_INVALID, "num_try must be 1..%d", MT_MAX_TRY);
  if (a.c.per) return fail(EB_ERR_UNSUPPORTED, "the multiple-try move draws from the priors: no periodic wrap");
  a.num_try = rng->num_try;
  a.philox = rng->mode == EB_RNG_PHILOX;
  a.tries = rng->tries; a.u_sel = rng->u_sel; a.u_acc = rng->u_acc;
  a.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(rng->seed >> 32);
  a.iter_dev = (const unsigned long long*)rng->iter_dev; a.iter = rng->iter;
  a.accepted = accepted; a.accepted_count = accepted_count;
  if (!a.philox) {
    if (rng->mode != EB_RNG_REPLAY) return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
    if (!rng->tries || !rng->u_sel || !rng->u_acc) return fail(EB_ERR_INVALID, "replay mode needs tries, u_sel, u_acc");
  }
  cudaStream_t s = (cudaStream_t)stream;
#define L2_(K) rc = launch_mt<DM_, K>(a, s)
#define L1_(DM, EX)                          \
  {                                          \
    constexpr int DM_ = DM;                  \
    EB_DISPATCH_LIKE(like->kind, L2_)        \
  }
  EB_DISPATCH_DMAX_GENERIC(a.c.LD, L1_)
#undef L1_
#undef L2_
  if (rc) return rc;
  return check_launch("mt_distgen");
}

}  // extern "C"
