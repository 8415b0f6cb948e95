// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include "common.cuh"

namespace eb {

// ================================================================================================
// K1: fused stretch half step
// ================================================================================================
struct StretchArgs {
  Common c;
  double a;
  int split, Ns, Nc;
  const int32_t* sub_idx; const int32_t* comp_idx; const long long* rint; const double* u_z; const double* u_acc;
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter; int randomize;
  uint8_t* accepted; uint32_t* accepted_count;
  // split path outputs
  double* q_out; double* factors_out; int32_t* sub_out;
};

// Select the moving walker w, its complement partner wc and the stretch/accept uniforms.
template <bool PHILOX>
__device__ __forceinline__ void stretch_draw(const StretchArgs& p, int t, int k, int& w, int& wc, double& u_z,
                                             double& u_acc) {
  if (PHILOX) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
    Feistel sig;
    if (p.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)t, (uint32_t)p.c.W);
    const uint32_t s = (uint32_t)p.split;
    const uint4 r = stream(key, TAG_STRETCH, (uint32_t)k, (uint32_t)(2 * t) + s);
    const uint32_t rint = __umulhi(r.x, (uint32_t)p.Nc);
    u_z = u01_52(r.z, r.w);
    uint32_t ws = 2u * (uint32_t)k + s, wcs = 2u * rint + (1u - s);
    if (p.randomize) { ws = sig(ws); wcs = sig(wcs); }
    w = (int)ws; wc = (int)wcs;
    const uint4 ra = stream(key, TAG_ACCEPT, (uint32_t)(t * p.c.W + w), s);
    u_acc = u01_52(ra.x, ra.y);
  } else {
    const size_t i = (size_t)t * p.Ns + k;
    w = p.sub_idx[i];
    wc = p.comp_idx[(size_t)t * p.Nc + (int)p.rint[i]];
    u_z = p.u_z[i];
    u_acc = p.u_acc ? p.u_acc[i] : 0.5;
  }
}

template <int DMAX, int LIKE, bool PHILOX>
__global__ void __launch_bounds__(BLOCK) stretch_half_step_kernel(const StretchArgs p) {
  extern __shared__ double sm[];
  const Common& c = p.c;
  stage_params(c, sm);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * p.Ns) return;
  const int t = tid / p.Ns, k = tid - t * p.Ns;
  int w, wc;
  double u_z, u_acc;
  stretch_draw<PHILOX>(p, t, k, w, wc, u_z, u_acc);

  const size_t slot = (size_t)t * c.W + w;
  double q[DMAX], cc[DMAX];
  load_row<DMAX>(c.coords + slot * c.LD, c.LD, q);                          // s  (red_blue.py:173-179)
  load_row<DMAX>(c.coords + ((size_t)t * c.W + wc) * c.LD, c.LD, cc);       // c_temp (stretch.py:100)
  const double ll0 = c.logl[slot], lp0 = c.logp[slot];
  const bool active = c.inds ? (c.inds[slot] != 0) : true;
  const bool tempered = c.betas != nullptr;
  const double beta = tempered ? c.betas[t] : 1.0;

  double zz = (p.a - 1.0) * u_z + 1.0;                                       // stretch.py:129-132
  zz = zz * zz / p.a;
#pragma unroll
  for (int j = 0; j < DMAX; ++j)
    if (j < c.LD) q[j] = cc[j] - (cc[j] - q[j]) * zz;                        // stretch.py:143-145
  const double factors = ((double)c.LD - 1.0) * log(zz);                     // stretch.py:223

  double lp, ll;
  eval_point<DMAX, LIKE>(q, c, sm, active, lp, ll);                          // red_blue.py:260,270
  const double logP = log_posterior(ll, lp, beta, tempered);                 // red_blue.py:283
  const double prevP = log_posterior(ll0, lp0, beta, tempered);              // red_blue.py:285-290
  const double lnpdiff = factors + logP - prevP;                             // red_blue.py:292
  const bool keep = lnpdiff > log(u_acc);                                    // red_blue.py:294

  if (keep) {                                                                // move.py:472-703
    store_row<DMAX>(c.coords + slot * c.LD, c.LD, q);
    c.logl[slot] = ll;
    c.logp[slot] = isinf(lp) ? 0.0 : lp;                                     // move.py:526
  }
  p.accepted[slot] = keep ? 1 : 0;
  if (p.accepted_count && keep) p.accepted_count[slot] += 1u;
}

// split path: proposal only.  Generic in L and D (rows streamed, nothing kept in registers).
template <bool PHILOX>
__global__ void __launch_bounds__(BLOCK) stretch_propose_kernel(const StretchArgs p) {
  const Common& c = p.c;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * p.Ns) return;
  const int t = tid / p.Ns, k = tid - t * p.Ns;
  int w, wc;
  double u_z, u_acc;
  stretch_draw<PHILOX>(p, t, k, w, wc, u_z, u_acc);
  double zz = (p.a - 1.0) * u_z + 1.0;
  zz = zz * zz / p.a;
  const double* s = c.coords + ((size_t)t * c.W + w) * c.LD;
  const double* cr = c.coords + ((size_t)t * c.W + wc) * c.LD;
  double* q = p.q_out + (size_t)tid * c.LD;
  for (int j = 0; j < c.LD; ++j) q[j] = cr[j] - (cr[j] - s[j]) * zz;
  p.factors_out[tid] = ((double)c.LD - 1.0) * log(zz);
  p.sub_out[tid] = w;
}

struct AcceptArgs {
  Common c;
  const int32_t* sub; int nsub;
  const double* q; const double* factors; const double* logl_new; const double* logp_new; const double* u_acc;
  int philox, slot;
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  uint8_t* accepted; uint32_t* accepted_count;
};

__global__ void __launch_bounds__(BLOCK) accept_update_kernel(const AcceptArgs p) {
  const Common& c = p.c;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * p.nsub) return;
  const int t = tid / p.nsub;
  const int w = p.sub[tid];
  const size_t slot = (size_t)t * c.W + w;
  double u;
  if (p.philox) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
    const uint4 ra = stream(key, TAG_ACCEPT, (uint32_t)slot, (uint32_t)p.slot);
    u = u01_52(ra.x, ra.y);
  } else {
    u = p.u_acc[tid];
  }
  const bool tempered = c.betas != nullptr;
  const double beta = tempered ? c.betas[t] : 1.0;
  double ll = p.logl_new[tid];
  const double lp = p.logp_new[tid];
  if (ll != ll) ll = FILL_LOGL;
  const double logP = log_posterior(ll, lp, beta, tempered);
  const double prevP = log_posterior(c.logl[slot], c.logp[slot], beta, tempered);
  const double f = p.factors ? p.factors[tid] : 0.0;
  const bool keep = (f + logP - prevP) > log(u);
  if (keep) {
    const double* q = p.q + (size_t)tid * c.LD;
    double* dst = c.coords + slot * c.LD;
    for (int j = 0; j < c.LD; ++j) dst[j] = q[j];
    c.logl[slot] = ll;
    c.logp[slot] = isinf(lp) ? 0.0 : lp;
  }
  p.accepted[slot] = keep ? 1 : 0;
  if (p.accepted_count && keep) p.accepted_count[slot] += 1u;
}

// box prior of arbitrary rows q[nrows][L][D] (ensemble.py:1192-1212)
template <int DMAX, int LIKE>
static int launch_stretch(const StretchArgs& a, bool philox, cudaStream_t s) {
  const int n = a.c.T * a.Ns;
  const size_t sb = smem_bytes(a.c);
  if (philox) {
    int rc = set_smem(stretch_half_step_kernel<DMAX, LIKE, true>, sb);
    if (rc) return rc;
    stretch_half_step_kernel<DMAX, LIKE, true><<<(n + BLOCK - 1) / BLOCK, BLOCK, sb, s>>>(a);
  } else {
    int rc = set_smem(stretch_half_step_kernel<DMAX, LIKE, false>, sb);
    if (rc) return rc;
    stretch_half_step_kernel<DMAX, LIKE, false><<<(n + BLOCK - 1) / BLOCK, BLOCK, sb, s>>>(a);
  }
  return EB_OK;
}

static int fill_stretch_args(StretchArgs& a, const eb_state* st, double stretch_a, int split, const eb_stretch_rng* rng) {
  if (!rng) return fail(EB_ERR_INVALID, "rng is NULL");
  if (split != 0 && split != 1) return fail(EB_ERR_INVALID, "split must be 0 or 1 (nsplits == 2)");
  if (!(stretch_a > 1.0)) return fail(EB_ERR_INVALID, "stretch scale a must be > 1");
  const int W = st->nwalkers;
  a.a = stretch_a; a.split = split;
  a.Ns = split == 0 ? (W + 1) / 2 : W / 2;      // red_blue.py:121: labels = arange(W) % 2
  a.Nc = W - a.Ns;
  if (a.Ns < 1 || a.Nc < 1) return fail(EB_ERR_INVALID, "nwalkers must be >= 2 for a red-blue move");
  a.sub_idx = rng->sub_idx; a.comp_idx = rng->comp_idx; a.rint = (const long long*)rng->rint;
  a.u_z = rng->u_z; a.u_acc = rng->u_acc;
  a.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(rng->seed >> 32);
  a.iter_dev = (const unsigned long long*)rng->iter_dev; a.iter = rng->iter; a.randomize = rng->randomize_split;
  a.q_out = nullptr; a.factors_out = nullptr; a.sub_out = nullptr;
  a.accepted = nullptr; a.accepted_count = nullptr;
  if (rng->mode == EB_RNG_REPLAY) {
    if (!rng->sub_idx || !rng->comp_idx || !rng->rint || !rng->u_z)
      return fail(EB_ERR_INVALID, "replay mode needs sub_idx, comp_idx, rint, u_z");
  } else if (rng->mode != EB_RNG_PHILOX) {
    return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  }
  return EB_OK;
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_stretch_half_step(const eb_state* st, const eb_prior* prior, const eb_like* like, double a, int32_t split,
                         const eb_stretch_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  StretchArgs args;
  int rc = fill_common(args.c, st, prior, like, true);
  if (rc) return rc;
  rc = fill_stretch_args(args, st, a, split, rng);
  if (rc) return rc;
  if (rng->mode == EB_RNG_REPLAY && !rng->u_acc) return fail(EB_ERR_INVALID, "replay mode needs u_acc");
  if (!accepted) return fail(EB_ERR_INVALID, "accepted is NULL");
  args.accepted = accepted; args.accepted_count = accepted_count;
  const bool philox = rng->mode == EB_RNG_PHILOX;
  cudaStream_t s = (cudaStream_t)stream;
#define L2_(K) rc = launch_stretch<DM_, K>(args, philox, s)
#define L1_(DM)                              \
  {                                          \
    constexpr int DM_ = DM;                  \
    EB_DISPATCH_LIKE(like->kind, L2_)        \
  }
  EB_DISPATCH_DMAX(args.c.LD, L1_)
#undef L1_
#undef L2_
  if (rc) return rc;
  return check_launch("stretch_half_step");
}

int eb_stretch_propose(const eb_state* st, double a, int32_t split, const eb_stretch_rng* rng, double* q,
                       double* factors, int32_t* sub_out, void* stream) {
  StretchArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_stretch_args(args, st, a, split, rng);
  if (rc) return rc;
  if (!q || !factors || !sub_out) return fail(EB_ERR_INVALID, "q/factors/sub_out is NULL");
  args.q_out = q; args.factors_out = factors; args.sub_out = sub_out;
  const int n = args.c.T * args.Ns;
  cudaStream_t s = (cudaStream_t)stream;
  if (rng->mode == EB_RNG_PHILOX) stretch_propose_kernel<true><<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(args);
  else stretch_propose_kernel<false><<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(args);
  return check_launch("stretch_propose");
}

int eb_accept_update(const eb_state* st, const int32_t* sub, int32_t nsub, const double* q, const double* factors,
                     const double* logl_new, const double* logp_new, const double* u_acc, int32_t slot,
                     const eb_stretch_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  AcceptArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  if (!sub || !q || !logl_new || !logp_new || !accepted) return fail(EB_ERR_INVALID, "NULL argument");
  if (nsub < 1 || nsub > st->nwalkers) return fail(EB_ERR_INVALID, "nsub out of range");
  args.sub = sub; args.nsub = nsub; args.q = q; args.factors = factors; args.logl_new = logl_new;
  args.logp_new = logp_new; args.u_acc = u_acc; args.slot = slot;
  args.philox = rng && rng->mode == EB_RNG_PHILOX;
  if (!args.philox && !u_acc) return fail(EB_ERR_INVALID, "replay mode needs u_acc");
  args.seed_lo = rng ? (uint32_t)(rng->seed & 0xFFFFFFFFull) : 0; args.seed_hi = rng ? (uint32_t)(rng->seed >> 32) : 0;
  args.iter_dev = rng ? (const unsigned long long*)rng->iter_dev : nullptr; args.iter = rng ? rng->iter : 0;
  args.accepted = accepted; args.accepted_count = accepted_count;
  const int n = args.c.T * nsub;
  accept_update_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, (cudaStream_t)stream>>>(args);
  return check_launch("accept_update");
}

}  // extern "C"
