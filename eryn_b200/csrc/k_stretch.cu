// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include <cooperative_groups.h>

#include "common.cuh"

namespace eb {

// ================================================================================================
// K1: fused StretchMove step (both red/blue halves in one launch)
// ================================================================================================
// Within a half, thread k of a temperature moves the k-th walker of the active split (dense warps:
// every lane does a proposal).  Half 1 reads what half 0 wrote within the same temperature only
// (red_blue.py:183-197 gathers along the walker axis), so the barrier between the halves is a
// thread-block-cluster barrier over the CTAs that own that temperature, not a grid barrier.
// Production (philox) mode, per walker: ONE Philox block gives the partner index + stretch uniform
// (split_draw) and the accept uniform; the random red/blue split of the temperature is a keyed
// bijection sigma_t of [0, W) (even positions = split 0), so the moving walker is sigma_t(2k+s) and
// its partner sigma_t(2*rint+1-s): no index lists in memory.
struct StretchArgs {
  Common c;
  double a;
  int philox, randomize;
  int both;        // 1: both halves in this launch (cluster barrier in between); 0: only `split`
  int split;
  int cpt;         // CTAs per temperature (= cluster size when both == 1)
  int Ns[2];
  // replay
  const int32_t* list[2]; const long long* rint[2]; const double* u_z[2]; const double* u_acc[2];
  // philox
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  uint8_t* accepted; uint32_t* accepted_count;
  // split path outputs
  double* q_out; double* factors_out; int32_t* sub_out;
};

constexpr int STRETCH_THREADS = 256;

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The k-th walker of split s at local temperature t: its id, its partner and the two uniforms.
template <bool PHILOX>
__device__ __forceinline__ void stretch_draw(const StretchArgs& p, const RngKey& key, const Feistel& sig, int t, int k,
                                             int s, int& w, int& wc, double& u_z, double& u_acc) {
  if (PHILOX) {
    const uint32_t pos = 2u * (uint32_t)k + (uint32_t)s;                     // red_blue.py:121-124
    const uint4 r = stream(key, TAG_STRETCH, pos, (uint32_t)(p.c.t0 + t));
    uint32_t rint;
    split_draw(r.x, r.y, (uint32_t)p.Ns[1 - s], rint, u_z);                  // stretch.py:93, :131
    u_acc = u01_52(r.z, r.w);                                                // red_blue.py:294
    const uint32_t ppos = 2u * rint + (uint32_t)(1 - s);
    w = (int)(p.randomize ? sig(pos) : pos);
    wc = (int)(p.randomize ? sig(ppos) : ppos);                              // stretch.py:100
  } else {
    const size_t i = (size_t)t * p.Ns[s] + k;
    w = p.list[s][i];                                                        // red_blue.py:150-154
    wc = p.list[1 - s][(size_t)t * p.Ns[1 - s] + (int)p.rint[s][i]];        // stretch.py:100
    u_z = p.u_z[s][i];
    u_acc = p.u_acc[s] ? p.u_acc[s][i] : 0.5;
  }
}

// One proposal = two stages.  `prepare` does everything that does not depend on the other split: the draws,
// the own row / logl / logp loads and both logarithms.  `finish` gathers the partner row (which, in half 1,
// the other split may have just rewritten), evaluates and applies the Metropolis test.  The kernel prepares
// BOTH halves of a thread up front, so that after the barrier between the halves only `finish` remains.
template <int DMAX>
struct WalkerJob {
  double q[DMAX];      // own coordinates (s), then the proposal
  double ll0, lp0, zz, factors, log_u;
  int w, wc;
  bool live, active;
};

template <int DMAX, bool PHILOX>
__device__ __forceinline__ void job_prepare(const StretchArgs& p, const RngKey& key, const Feistel& sig, int t, int k,
                                            int s, WalkerJob<DMAX>& j) {
  const Common& c = p.c;
  j.live = k < p.Ns[s];
  if (!j.live) return;
  double u_z, u_acc;
  stretch_draw<PHILOX>(p, key, sig, t, k, s, j.w, j.wc, u_z, u_acc);
  const size_t slot = (size_t)t * c.W + j.w;
  load_row<DMAX>(c.coords + slot * c.LD, c.LD, j.q);                         // s  (red_blue.py:173-179)
  j.ll0 = c.logl[slot];
  j.lp0 = c.logp[slot];
  j.active = c.inds ? (c.inds[slot] != 0) : true;
  double zz = (p.a - 1.0) * u_z + 1.0;                                       // stretch.py:129-132
  zz = zz * zz / p.a;
  j.zz = zz;
  j.factors = ((double)c.LD - 1.0) * log(zz);                                // stretch.py:223
  j.log_u = log(u_acc);                                                      // red_blue.py:294
}

template <int DMAX, int LIKE>
__device__ __forceinline__ void job_finish(const StretchArgs& p, const double* sm, int t, WalkerJob<DMAX>& j) {
  if (!j.live) return;
  const Common& c = p.c;
  const size_t slot = (size_t)t * c.W + j.w;
  double cc[DMAX];
  load_row<DMAX>(c.coords + ((size_t)t * c.W + j.wc) * c.LD, c.LD, cc);     // c_temp (stretch.py:100)
  const bool tempered = c.betas != nullptr;
  const double beta = tempered ? c.betas[t] : 1.0;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < c.LD) j.q[d] = cc[d] - (cc[d] - j.q[d]) * j.zz;                  // stretch.py:143-145
  double lp, ll;
  eval_point<DMAX, LIKE>(j.q, c, sm, j.active, lp, ll);                      // red_blue.py:260,270
  const double logP = log_posterior(ll, lp, beta, tempered);                 // red_blue.py:283
  const double prevP = log_posterior(j.ll0, j.lp0, beta, tempered);          // red_blue.py:285-290
  const double lnpdiff = j.factors + logP - prevP;                           // red_blue.py:292
  const bool keep = lnpdiff > j.log_u;                                       // red_blue.py:294
  if (keep) {                                                                // move.py:472-703
    store_row<DMAX>(c.coords + slot * c.LD, c.LD, j.q);
    c.logl[slot] = ll;
    c.logp[slot] = isinf(lp) ? 0.0 : lp;                                     // move.py:526
    if (p.accepted_count) p.accepted_count[slot] += 1u;
  }
  p.accepted[slot] = keep ? 1 : 0;
}

// grid = (cpt, T); with p.both the launch carries cluster dimension (cpt, 1, 1)
template <int DMAX, int LIKE, bool PHILOX>
__global__ void __launch_bounds__(STRETCH_THREADS, DMAX <= 8 ? 2 : 1) stretch_step_kernel(const StretchArgs p) {
  extern __shared__ double sm[];
  const Common& c = p.c;
  const int t = blockIdx.y;
  EB_MARK(0);
  RngKey key;
  Feistel sig;
  if (PHILOX) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    // every thread derives the split bijection of its temperature itself: one more Philox block per
    // thread, but no block barrier in front of the draws
    if (p.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)(c.t0 + t), (uint32_t)c.W);
  }
  const int stride = p.cpt * blockDim.x;
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int s_first = p.both ? 0 : p.split, s_last = p.both ? 1 : p.split;
  constexpr bool PRE = DMAX <= 16;       // both jobs live in registers at once
  WalkerJob<DMAX> ja, jb;
  EB_MARK(1);
  job_prepare<DMAX, PHILOX>(p, key, sig, t, k0, s_first, ja);
  if (PRE && p.both) job_prepare<DMAX, PHILOX>(p, key, sig, t, k0, 1, jb);
  EB_MARK(2);
  stage_params(c, sm);  // ends with __syncthreads()
  EB_MARK(3);
  job_finish<DMAX, LIKE>(p, sm, t, ja);
  EB_MARK(4);
  for (int k = k0 + stride; k < p.Ns[s_first]; k += stride) {
    job_prepare<DMAX, PHILOX>(p, key, sig, t, k, s_first, ja);
    job_finish<DMAX, LIKE>(p, sm, t, ja);
  }
  if (s_last != s_first) {
    if (!PRE) job_prepare<DMAX, PHILOX>(p, key, sig, t, k0, 1, jb);   // own row: not touched by half 0
    // half 1 gathers what half 0 wrote, inside this temperature only
    EB_MARK(5);
    if (p.cpt > 1) cluster_barrier();
    else __syncthreads();
    EB_MARK(6);
    job_finish<DMAX, LIKE>(p, sm, t, jb);
    EB_MARK(7);
    for (int k = k0 + stride; k < p.Ns[1]; k += stride) {
      job_prepare<DMAX, PHILOX>(p, key, sig, t, k, 1, jb);
      job_finish<DMAX, LIKE>(p, sm, t, jb);
    }
  }
}

// split path: proposal only, thread per (t, k) of split s.  Generic in L and D (rows streamed).
template <bool PHILOX>
__global__ void __launch_bounds__(BLOCK) stretch_propose_kernel(const StretchArgs p) {
  const Common& c = p.c;
  const int s = p.split, Ns = p.Ns[s];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= c.T * Ns) return;
  const int t = tid / Ns, k = tid - t * Ns;
  RngKey key;
  Feistel sig;
  if (PHILOX) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    if (p.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)(c.t0 + t), (uint32_t)c.W);
  }
  int w, wc;
  double u_z, u_acc;
  stretch_draw<PHILOX>(p, key, sig, t, k, s, w, wc, u_z, u_acc);
  double zz = (p.a - 1.0) * u_z + 1.0;
  zz = zz * zz / p.a;
  const double* sr = c.coords + ((size_t)t * c.W + w) * c.LD;
  const double* cr = c.coords + ((size_t)t * c.W + wc) * c.LD;
  double* q = p.q_out + (size_t)tid * c.LD;
  for (int j = 0; j < c.LD; ++j) q[j] = cr[j] - (cr[j] - sr[j]) * zz;
  p.factors_out[tid] = ((double)c.LD - 1.0) * log(zz);
  p.sub_out[tid] = w;
}

struct AcceptArgs {
  Common c;
  const int32_t* sub; int nsub;
  const double* q; const double* factors; const double* logl_new; const double* logp_new; const double* u_acc;
  int philox, split;
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
    const int k = tid - t * p.nsub;      // rows of q are in split-rank order (eb_stretch_propose)
    const uint4 ra = stream(key, TAG_STRETCH, (uint32_t)(2 * k + p.split), (uint32_t)(c.t0 + t));
    u = u01_52(ra.z, ra.w);
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

template <typename K>
static int launch_stretch_kernel(K kernel, const StretchArgs& a, size_t sb, cudaStream_t s) {
  int rc = set_smem(kernel, sb);
  if (rc) return rc;
  int threads = (a.Ns[0] + a.cpt - 1) / a.cpt;   // one walker of the active split per thread
  threads = ((threads + 31) / 32) * 32;
  if (threads > STRETCH_THREADS) threads = STRETCH_THREADS;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)a.cpt, (unsigned)a.c.T, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = sb;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  if (a.both && a.cpt > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.cpt;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  EB_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return EB_OK;
}

template <int DMAX, int LIKE>
static int launch_stretch(const StretchArgs& a, cudaStream_t s) {
  const size_t sb = smem_bytes(a.c);
  if (a.philox) return launch_stretch_kernel(stretch_step_kernel<DMAX, LIKE, true>, a, sb, s);
  return launch_stretch_kernel(stretch_step_kernel<DMAX, LIKE, false>, a, sb, s);
}

static int fill_stretch_args(StretchArgs& a, const eb_state* st, double stretch_a, const eb_stretch_rng* rng,
                             bool need_u_acc) {
  if (!rng) return fail(EB_ERR_INVALID, "rng is NULL");
  if (!(stretch_a > 1.0)) return fail(EB_ERR_INVALID, "stretch scale a must be > 1");
  const int W = st->nwalkers;
  a.a = stretch_a;
  a.Ns[0] = (W + 1) / 2;                         // red_blue.py:121: labels = arange(W) % 2
  a.Ns[1] = W / 2;
  if (a.Ns[1] < 1) return fail(EB_ERR_INVALID, "nwalkers must be >= 2 for a red-blue move");
  a.philox = rng->mode == EB_RNG_PHILOX;
  a.randomize = rng->randomize_split;
  for (int s = 0; s < 2; ++s) {
    a.list[s] = rng->list[s]; a.rint[s] = (const long long*)rng->rint[s];
    a.u_z[s] = rng->u_z[s]; a.u_acc[s] = rng->u_acc[s];
  }
  a.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(rng->seed >> 32);
  a.iter_dev = (const unsigned long long*)rng->iter_dev; a.iter = rng->iter;
  a.q_out = nullptr; a.factors_out = nullptr; a.sub_out = nullptr;
  a.accepted = nullptr; a.accepted_count = nullptr;
  a.both = 0; a.split = 0; a.cpt = 1;
  if (rng->mode == EB_RNG_REPLAY) {
    for (int s = 0; s < 2; ++s)
      if (!rng->list[s] || !rng->rint[s] || !rng->u_z[s] || (need_u_acc && !rng->u_acc[s]))
        return fail(EB_ERR_INVALID, "replay mode needs list, rint, u_z%s for both splits", need_u_acc ? ", u_acc" : "");
  } else if (rng->mode != EB_RNG_PHILOX) {
    return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  }
  return EB_OK;
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_stretch_step(const eb_state* st, const eb_prior* prior, const eb_like* like, double a,
                    const eb_stretch_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  StretchArgs args;
  int rc = fill_common(args.c, st, prior, like, true);
  if (rc) return rc;
  rc = fill_stretch_args(args, st, a, rng, true);
  if (rc) return rc;
  if (!accepted) return fail(EB_ERR_INVALID, "accepted is NULL");
  args.accepted = accepted; args.accepted_count = accepted_count;
  cudaStream_t s = (cudaStream_t)stream;
  // One cluster (<= 8 CTAs, portable size) per temperature runs both halves when that still fills the
  // chip or the temperature is small; otherwise one grid-wide launch per half.
  const int T = args.c.T, half = args.Ns[0];
  int cpt = 1;
  while (cpt < 8 && cpt * STRETCH_THREADS < half) cpt <<= 1;
  const bool fused = (half <= cpt * STRETCH_THREADS * 4) || (T * cpt >= 148);
  int nlaunch = 1;
  if (fused) {
    args.both = 1; args.cpt = cpt;
  } else {
    args.both = 0; args.cpt = (half + STRETCH_THREADS - 1) / STRETCH_THREADS;
    nlaunch = 2;
  }
  for (int l = 0; l < nlaunch; ++l) {
    args.split = l;
#define L2_(K) rc = launch_stretch<DM_, K>(args, s)
#define L1_(DM)                              \
  {                                          \
    constexpr int DM_ = DM;                  \
    EB_DISPATCH_LIKE(like->kind, L2_)        \
  }
    EB_DISPATCH_DMAX(args.c.LD, L1_)
#undef L1_
#undef L2_
    if (rc) return rc;
  }
  return check_launch("stretch_step");
}

int eb_stretch_propose(const eb_state* st, double a, int32_t split, const eb_stretch_rng* rng, double* q,
                       double* factors, int32_t* sub_out, void* stream) {
  StretchArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_stretch_args(args, st, a, rng, false);
  if (rc) return rc;
  if (split != 0 && split != 1) return fail(EB_ERR_INVALID, "split must be 0 or 1 (nsplits == 2)");
  if (!q || !factors || !sub_out) return fail(EB_ERR_INVALID, "q/factors/sub_out is NULL");
  args.split = split;
  args.q_out = q; args.factors_out = factors; args.sub_out = sub_out;
  const int n = args.c.T * args.Ns[split];
  cudaStream_t s = (cudaStream_t)stream;
  if (rng->mode == EB_RNG_PHILOX) stretch_propose_kernel<true><<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(args);
  else stretch_propose_kernel<false><<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(args);
  return check_launch("stretch_propose");
}

int eb_accept_update(const eb_state* st, const int32_t* sub, int32_t nsub, const double* q, const double* factors,
                     const double* logl_new, const double* logp_new, int32_t split, const eb_stretch_rng* rng,
                     uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  AcceptArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  if (!sub || !q || !logl_new || !logp_new || !accepted || !rng) return fail(EB_ERR_INVALID, "NULL argument");
  if (nsub < 1 || nsub > st->nwalkers) return fail(EB_ERR_INVALID, "nsub out of range");
  if (split != 0 && split != 1) return fail(EB_ERR_INVALID, "split must be 0 or 1 (nsplits == 2)");
  args.sub = sub; args.nsub = nsub; args.q = q; args.factors = factors; args.logl_new = logl_new;
  args.logp_new = logp_new; args.split = split;
  args.philox = rng->mode == EB_RNG_PHILOX;
  args.u_acc = args.philox ? nullptr : rng->u_acc[split];
  if (!args.philox && !args.u_acc) return fail(EB_ERR_INVALID, "replay mode needs u_acc");
  args.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); args.seed_hi = (uint32_t)(rng->seed >> 32);
  args.iter_dev = (const unsigned long long*)rng->iter_dev; args.iter = rng->iter;
  args.accepted = accepted; args.accepted_count = accepted_count;
  const int n = args.c.T * nsub;
  accept_update_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, (cudaStream_t)stream>>>(args);
  return check_launch("accept_update");
}

}  // extern "C"

EB_DEFINE_MARK_READER(eb_debug_marks_stretch)
