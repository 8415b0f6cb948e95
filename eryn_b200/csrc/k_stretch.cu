// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace eb {

// ================================================================================================
// K1: fused StretchMove step — one launch per red/blue half, chained by programmatic dependent launch
// ================================================================================================
// Within a half, thread k of a temperature moves the k-th walker of the active split (dense warps: every lane does a
// proposal).  Half 1 reads what half 0 wrote, inside the same temperature only (red_blue.py:183-197 gathers along the
// walker axis): it is launched as a programmatic dependent of half 0 — its draws (and the loads of its own rows, which
// half 0 does not touch) run while half 0 still evaluates, only its `finish` stage waits (griddepcontrol.wait).  Small
// ensembles (Ns <= 128) run both halves in one CTA per temperature with a block barrier in between.  (A thread-block
// cluster per temperature with a cluster barrier between the halves was measured and dropped)
// Production (philox) mode, per walker: ONE Philox block gives the partner index + stretch uniform (split_draw) and the
// accept uniform; the random red/blue split of the temperature is a keyed bijection sigma_t of [0, W) (even positions =
// split 0), so the moving walker is sigma_t(2k+s) and its partner sigma_t(2*rint+1-s): no index lists in memory.
// HBM-sized shapes take the lane-split variant of this kernel (stretch_lanes.cuh).
struct StretchArgs {
  Common c;
  double a;
  int philox, randomize;
  int both;        // 1: both halves in this launch (one CTA per temperature); 0: only `split`
  int split;
  int pdl;         // this launch carries the programmatic-stream-serialization attribute
  uint32_t gmask;  // Gibbs split: parameters that move (0 = all), number of them, index of the split in this propose call
  int gndim, gidx;
  int Ns[2];
  // replay
  const int32_t* list[2]; const long long* rint[2]; const double* u_z[2]; const double* u_acc[2];
  // philox
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  uint8_t* accepted; uint32_t* accepted_count;
  eb_ctrl* lazy_ctrl;   // first launch of a step after a pass that deferred its ladder adaptation (common.cuh:lazy_adapt_apply)
  // split path outputs
  double* q_out; double* factors_out; int32_t* sub_out;
};

constexpr int STRETCH_HALF_THREADS = 256;   // threads per CTA of a per-half launch

// The k-th walker of split s at local temperature t: its id, its partner and the two uniforms.
template <bool PHILOX>
__device__ __forceinline__ void stretch_draw(const StretchArgs& p, const RngKey& key, const Feistel& sig, int t, int k,
                                             int s, int& w, int& wc, double& u_z, double& u_acc) {
  if (PHILOX) {
    const uint32_t pos = 2u * (uint32_t)k + (uint32_t)s;                     // red_blue.py:121-124
    const uint4 r = stream(key, TAG_STRETCH, pos, (uint32_t)(p.c.t0 + t) | ((uint32_t)p.gidx << 16));
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

// One proposal = three stages.  `draw`: the random draws and both logarithms — no state is read, so it can run
// before the previous kernel has finished (programmatic dependent launch).  `load_own`: the walker's own row, logl,
// logp.  `finish`: gather the partner row, evaluate, Metropolis test, in-place update.
template <int DMAX>
struct WalkerJob {
  double q[DMAX];      // own coordinates (s), then the proposal
  double ll0, lp0, zz, factors, log_u, beta;
  int w, wc;
  bool live, active;
};

template <int DMAX, bool PHILOX, bool EXACT>
__device__ __forceinline__ void job_draw(const StretchArgs& p, const RngKey& key, const Feistel& sig, int t, int k,
                                         int s, WalkerJob<DMAX>& j) {
  const int LD = EXACT ? DMAX : p.c.LD;
  j.live = k < p.Ns[s];
  if (!j.live) return;
  double u_z, u_acc;
  stretch_draw<PHILOX>(p, key, sig, t, k, s, j.w, j.wc, u_z, u_acc);
  double zz = (p.a - 1.0) * u_z + 1.0;                                       // stretch.py:129-132
  zz = zz * zz / p.a;
  j.zz = zz;
  j.factors = ((double)LD - 1.0) * log(zz);                                  // stretch.py:223
  if (p.gmask && p.gndim != LD)                                              // stretch.py:55-72 adjust_factors (Gibbs split)
    j.factors = j.factors / ((double)LD - 1.0) * ((double)p.gndim - 1.0);
  j.log_u = log(u_acc);                                                      // red_blue.py:294
}

template <int DMAX, bool EXACT>
__device__ __forceinline__ void job_load_own(const StretchArgs& p, int t, WalkerJob<DMAX>& j) {
  if (!j.live) return;
  const Common& c = p.c;
  const int LD = EXACT ? DMAX : c.LD;
  const size_t slot = (size_t)t * c.W + j.w;
  load_row<DMAX>(c.coords + slot * LD, LD, j.q);                             // s  (red_blue.py:173-179)
  j.ll0 = c.logl[slot];
  j.lp0 = c.logp[slot];
  j.active = c.inds ? (c.inds[slot] != 0) : true;
}

// the partner row c_temp (stretch.py:100): its own stage so that the request can go out ahead of other work
template <int DMAX, bool EXACT>
__device__ __forceinline__ void job_load_partner(const StretchArgs& p, int t, const WalkerJob<DMAX>& j, double (&cc)[DMAX]) {
  if (!j.live) return;
  const Common& c = p.c;
  const int LD = EXACT ? DMAX : c.LD;
  load_row<DMAX>(c.coords + ((size_t)t * c.W + j.wc) * LD, LD, cc);
}

template <int DMAX, int LIKE, bool EXACT>
__device__ __forceinline__ void job_finish(const StretchArgs& p, const double* sm, int t, WalkerJob<DMAX>& j,
                                           double (&cc)[DMAX]) {
  if (!j.live) return;
  const Common& c = p.c;
  const int LD = EXACT ? DMAX : c.LD;
  const size_t slot = (size_t)t * c.W + j.w;
  const bool tempered = c.betas != nullptr;
  if (c.per) {
    // periodic parameters (single leaf: LD == D): the distance from s to c goes through the boundary when that is
    // shorter (utils/periodic.py:49-117), the proposal is wrapped into [0, period) (:119-151)
    const double* per = sm + 3 * LD;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (EXACT || d < LD) {
        const double P = per[d], s0 = j.q[d];
        double diff = cc[d] - s0;
        if (P > 0.0 && fabs(diff) > P / 2.0) {
          const double new_s = diff < 0.0 ? -(P - s0) : (P + s0);
          diff = cc[d] - new_s;
        }
        double v = cc[d] - diff * j.zz;                                       // stretch.py:145
        if (P > 0.0) v = np_mod(v, P);
        if (p.gmask && !((p.gmask >> d) & 1u)) v = s0;                        // move.py:302-307 (Gibbs split)
        j.q[d] = v;
      }
  } else {
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (EXACT || d < LD) {
        const double v = cc[d] - (cc[d] - j.q[d]) * j.zz;                     // stretch.py:143-145
        j.q[d] = (p.gmask && !((p.gmask >> d) & 1u)) ? j.q[d] : v;            // move.py:302-307 (Gibbs split)
      }
  }
  double lp, ll;
  eval_point<DMAX, LIKE, EXACT>(j.q, c, sm, j.active, lp, ll);               // red_blue.py:260,270
  const double logP = log_posterior(ll, lp, j.beta, tempered);               // red_blue.py:283
  const double prevP = log_posterior(j.ll0, j.lp0, j.beta, tempered);        // red_blue.py:285-290
  const double lnpdiff = j.factors + logP - prevP;                           // red_blue.py:292
  const bool keep = lnpdiff > j.log_u;                                       // red_blue.py:294
  if (keep) {                                                                // move.py:472-703
    store_row<DMAX>(c.coords + slot * LD, LD, j.q);
    c.logl[slot] = ll;
    c.logp[slot] = isinf(lp) ? 0.0 : lp;
  }
  // from the second Gibbs split of a propose call on, the reported mask is the running OR over the splits and the counter
  // grows by that OR (red_blue.py:296-309, :325)
  const bool rep = keep || (p.gidx > 0 && p.accepted[slot] != 0);
  if (p.accepted_count && rep) p.accepted_count[slot] += 1u;
  p.accepted[slot] = rep ? 1 : 0;
}

// programmatic dependent launch (PDL): the next kernel in the stream may start its state-independent prologue once
// every CTA of this grid has executed launch_dependents; it blocks in grid_dependency_wait until this grid has
// completed and its writes are visible.  Both are no-ops for launches without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Per-half launch (!p.both): grid = (ceil(Ns[split] / blockDim), T), every thread owns one walker of `split`; half 1
// is launched as a programmatic dependent of half 0, so its draws overlap half 0 and only its `finish` waits.
// Small ensembles (p.both: Ns[0] <= blockDim/2, grid = (1, T)): both halves in one CTA per temperature, the first
// half of the threads owns split 0, the second half split 1, a block barrier in between.
template <int DMAX, int LIKE, bool PHILOX, bool EXACT>
__global__ void __launch_bounds__(STRETCH_HALF_THREADS, 2) stretch_step_kernel(const StretchArgs p) {
  extern __shared__ __align__(16) double sm[];
  const Common& c = p.c;
  const int t = blockIdx.y;
  EB_MARK(0);
  unsigned long long it = p.iter;
  if (PHILOX && p.iter_dev) it = *reinterpret_cast<const volatile unsigned long long*>(p.iter_dev);
  // half 1 lets the swap pass start its own prologue right away; half 0 releases half 1 only after its wait, so that
  // half 1 never runs ahead of the kernel BEFORE half 0 (it loads its own rows before waiting)
  if (!p.both && p.split == 1) pdl_launch_dependents();
  Staged staged;
  stage_load(c, staged);                // constant parameters: global loads in flight while the draws are computed
  RngKey key;
  Feistel sig;
  if (PHILOX) {
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    // every thread derives the split bijection of its temperature itself: one more Philox block per
    // thread, but no block barrier in front of the draws
    if (p.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)(c.t0 + t), (uint32_t)c.W);
  }
  int s, k;
  if (p.both) {
    const int ht = blockDim.x >> 1;
    s = threadIdx.x >= ht ? 1 : 0;
    k = threadIdx.x - s * ht;
  } else {
    s = p.split;
    k = blockIdx.x * blockDim.x + threadIdx.x;
  }
  WalkerJob<DMAX> job;
  EB_MARK(1);
  job_draw<DMAX, PHILOX, EXACT>(p, key, sig, t, k, s, job);
  job.beta = 1.0;
  EB_MARK(2);
  const bool own_early = !p.both && s == 1;   // split-1 rows are not touched by half 0
  if (own_early) job_load_own<DMAX, EXACT>(p, t, job);
  stage_store(c, staged, sm);                 // constant parameters (no kernel writes them): in place before the wait, so that
  __syncthreads();                            // the block barrier is not on the path behind it
  pdl_wait();
  if (p.both || s == 0) pdl_launch_dependents();
  if (!own_early) job_load_own<DMAX, EXACT>(p, t, job);
  // the partner row is requested now unless this is the second half of a fused CTA (which must wait for the first)
  double cc[DMAX];
  const bool partner_early = !(p.both && s == 1);
  if (partner_early) job_load_partner<DMAX, EXACT>(p, t, job, cc);
  // a pass that deferred its ladder adaptation: every CTA folds the counts and adapts the ladder itself, under the row
  // loads just issued (the ladder is needed at the Metropolis test only); CTA (0,0) also does the bookkeeping
  __shared__ LazyShared lazy_sh;
  bool lazy = false;
  if (PHILOX && p.lazy_ctrl && (p.both || s == 0))
    lazy = lazy_adapt_apply(p.lazy_ctrl, it, c.betas, blockIdx.x == 0 && blockIdx.y == 0, false, lazy_sh);
  if (c.betas) job.beta = lazy ? lazy_sh.b[c.t0 + t] : c.betas[t];   // adapted by the swap pass: read after the wait
  EB_MARK(3);
  // Half 1 gathers what half 0 wrote, inside this temperature only.  In a fused CTA the split-1 warps arrive at the
  // barrier first and wait there for the split-0 warps, which arrive after their `finish`.
  const bool wait_first = p.both && s == 1;
  if (wait_first) asm volatile("barrier.sync.aligned 0;" ::: "memory");
  EB_MARK(4);
  if (!partner_early) job_load_partner<DMAX, EXACT>(p, t, job, cc);
  job_finish<DMAX, LIKE, EXACT>(p, sm, t, job, cc);
  EB_MARK(6);
  if (p.both && s == 0) asm volatile("barrier.sync.aligned 0;" ::: "memory");
  EB_MARK(7);
}

}  // namespace eb
#include "stretch_lanes.cuh"
namespace eb {

#if !defined(EB_ONLY_LIKE) || EB_ONLY_LIKE == 0
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
  for (int j = 0; j < c.LD; ++j) {
    const double P = c.per ? c.per[j % c.D] : 0.0;
    double diff = cr[j] - sr[j];
    if (P > 0.0 && fabs(diff) > P / 2.0) diff = cr[j] - (diff < 0.0 ? -(P - sr[j]) : (P + sr[j]));
    double v = cr[j] - diff * zz;
    if (P > 0.0) v = np_mod(v, P);
    q[j] = v;
  }
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

#endif  // split path kernels (object 0 only)

template <typename K>
static int launch_stretch_kernel(K kernel, const StretchArgs& a, size_t sb, cudaStream_t s) {
  int rc = set_smem(kernel, sb);
  if (rc) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  if (a.both) {
    const int ht = ((a.Ns[0] + 31) / 32) * 32;
    cfg.gridDim = dim3(1, (unsigned)a.c.T, 1);
    cfg.blockDim = dim3((unsigned)(2 * ht), 1, 1);
    if (a.pdl) {
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    }
  } else {
    const int n = a.Ns[a.split];
    const int threads = n < STRETCH_HALF_THREADS ? ((n + 31) / 32) * 32 : STRETCH_HALF_THREADS;
    cfg.gridDim = dim3((unsigned)((n + threads - 1) / threads), (unsigned)a.c.T, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    if (a.pdl) {
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    }
  }
  cfg.dynamicSmemBytes = sb;
  cfg.stream = s;
  EB_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return EB_OK;
}

// Which kernel runs a per-half launch: 0 = one thread per walker with the row in registers (stretch_step_kernel: the
// latency-optimised kernel of the shapes that do not fill the GPU), 2 / 4 = lanes per walker (stretch_lanes.cuh).  Below
// about one wave of CTAs the step is bound by the latency of one walker's dependent chain, which the rounds of the
// lane-split kernel lengthen; at HBM-sized shapes occupancy decides (numbers in profiles/README.md).  EB_K1_LPW = 1 / 2 / 4
// forces a variant (tests, profiling); it is read at every launch.
static int stretch_variant_for(const StretchArgs& a) {
  const char* env = getenv("EB_K1_LPW");
  const int forced = env ? atoi(env) : 0;
  if (a.both || a.gmask || a.c.L != 1 || (a.c.LD != 8 && a.c.LD != 20)) return 0;
  if (forced == 1) return 0;
  if (forced == 2 && a.c.LD == 8) return 2;
  if (forced == 2 || forced == 4) return 4;
  static const long long min_walkers = getenv("EB_K1_LANES_MIN") ? atoll(getenv("EB_K1_LANES_MIN")) : 148ll * 2 * 256;
  return (long long)a.c.T * a.Ns[a.split] >= min_walkers ? 4 : 0;
}

template <int DMAX, int LIKE, bool EXACT>
static int launch_stretch(const StretchArgs& a, cudaStream_t s) {
  const size_t sb = smem_bytes(a.c);
  if constexpr (EXACT) {
    const int v = stretch_variant_for(a);
    if (v == 4) {
      if (a.philox) return launch_stretch_kernel(stretch_lanes_kernel<DMAX, 4, LIKE, true>, a, sb, s);
      return launch_stretch_kernel(stretch_lanes_kernel<DMAX, 4, LIKE, false>, a, sb, s);
    }
    if constexpr (DMAX == 8) if (v == 2) {
      if (a.philox) return launch_stretch_kernel(stretch_lanes_kernel<8, 2, LIKE, true>, a, sb, s);
      return launch_stretch_kernel(stretch_lanes_kernel<8, 2, LIKE, false>, a, sb, s);
    }
  }
  if (a.philox) return launch_stretch_kernel(stretch_step_kernel<DMAX, LIKE, true, EXACT>, a, sb, s);
  return launch_stretch_kernel(stretch_step_kernel<DMAX, LIKE, false, EXACT>, a, sb, s);
}

// one object file per likelihood kind (build.py compiles this source with -DEB_ONLY_LIKE=k): the kernels of kind k
template <int LIKE>
int launch_stretch_like(const StretchArgs& a, cudaStream_t s);
#ifndef EB_ONLY_LIKE
#define EB_ONLY_LIKE -1   // single-object build: everything
#endif
#define EB_STRETCH_LIKE_DEF(K)                                        \
  template <>                                                         \
  int launch_stretch_like<K>(const StretchArgs& a, cudaStream_t s) {  \
    int rc = EB_OK;                                                   \
    EB_DISPATCH_DMAX(a.c.LD, EB_STRETCH_L1_##K)                       \
    return rc;                                                        \
  }
#define EB_STRETCH_L1_0(DM, EX) rc = launch_stretch<DM, 0, EX>(a, s)
#define EB_STRETCH_L1_1(DM, EX) rc = launch_stretch<DM, 1, EX>(a, s)
#define EB_STRETCH_L1_2(DM, EX) rc = launch_stretch<DM, 2, EX>(a, s)
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
EB_STRETCH_LIKE_DEF(0)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 1
EB_STRETCH_LIKE_DEF(1)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 2
EB_STRETCH_LIKE_DEF(2)
#endif

}  // namespace eb
#include "resident.cuh"
namespace eb {

#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
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
  a.both = 0; a.split = 0; a.pdl = 0;
  a.lazy_ctrl = nullptr;
  a.gmask = rng->gibbs_mask; a.gndim = rng->gibbs_ndim; a.gidx = rng->gibbs_index;
  if (a.gmask) {
    const int LD = st->nleaves * st->ndim;
    if (st->nleaves != 1) return fail(EB_ERR_UNSUPPORTED, "Gibbs splits of the fused stretch kernel address the parameters of one leaf");
    if (LD < 32 && (a.gmask >> LD)) return fail(EB_ERR_INVALID, "gibbs_mask selects parameters beyond ndim");
    if (a.gndim != __builtin_popcount(a.gmask)) return fail(EB_ERR_INVALID, "gibbs_ndim must count the bits of gibbs_mask");
  }
  if (a.gidx < 0 || a.gidx > 0xFFFF) return fail(EB_ERR_INVALID, "gibbs_index out of range");
  if (rng->pdl_chain && rng->mode != EB_RNG_PHILOX) return fail(EB_ERR_INVALID, "pdl_chain is a philox-mode option");
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
  // One launch per half; half 1 is a programmatic dependent of half 0 (its draws overlap half 0's evaluation).  Small
  // ensembles run both halves in one CTA per temperature.  (A cluster-per-temperature variant with a cluster barrier
  // between the halves measured slower: 16 clusters of 8 CTAs do not all fit the GPCs at once — DESIGN.md §4.1.)
  static const int pdl_mask = getenv("EB_PDL_MASK") ? atoi(getenv("EB_PDL_MASK")) : 7;   // 1: half 1, 4: half 0 after a swap pass
  const int half = args.Ns[0];
  const bool fused = half <= STRETCH_HALF_THREADS / 2;
  const int nlaunch = fused ? 1 : 2;
  args.both = fused ? 1 : 0;
  for (int l = 0; l < nlaunch; ++l) {
    args.split = l;
    args.pdl = ((l == 1 && (pdl_mask & 1)) || (l == 0 && rng->pdl_chain && (pdl_mask & 4))) ? 1 : 0;
    args.lazy_ctrl = (l == 0 && args.philox && st->temp_offset == 0) ? (eb_ctrl*)rng->lazy_ctrl : nullptr;
    switch (like->kind) {
      case 0: rc = launch_stretch_like<0>(args, s); break;
      case 1: rc = launch_stretch_like<1>(args, s); break;
      default: rc = launch_stretch_like<2>(args, s); break;
    }
    if (rc) return rc;
  }
  return check_launch("stretch_step");
}

size_t eb_resident_scratch_bytes(const eb_state* st) {
  if (!st || st->ntemps < 1 || st->nwalkers < 1 || st->nleaves < 1 || st->ndim < 1) return 0;
  const size_t n = (size_t)st->ntemps * st->nwalkers;
  const size_t RS = ((size_t)st->nleaves * st->ndim + 2 + 1) & ~(size_t)1;
  return RES_HEADER_BYTES + ((n * sizeof(int32_t) + 15) & ~(size_t)15) + 2 * n * RS * sizeof(double);
}

int eb_resident_run(const eb_state* st, const eb_prior* prior, const eb_like* like, double a, const eb_stretch_rng* srng,
                    const eb_swap_rng* wrng, const eb_adapt* adapt, eb_ctrl* ctrl, int32_t niter, uint8_t* accepted,
                    uint32_t* accepted_count, void* scratch, size_t scratch_bytes, void* stream) {
  ResidentArgs args;
  memset(&args, 0, sizeof(args));
  int rc = fill_common(args.sa.c, st, prior, like, true);
  if (rc) return rc;
  rc = fill_stretch_args(args.sa, st, a, srng, true);
  if (rc) return rc;
  const Common& c = args.sa.c;
  if (!wrng || !ctrl) return fail(EB_ERR_INVALID, "swap rng / ctrl is NULL");
  if (niter < 0) return fail(EB_ERR_INVALID, "niter < 0");
  // what the resident kernel covers; everything else runs as eb_stretch_step + eb_pt_swap
  if (!args.sa.philox || wrng->mode != EB_RNG_PHILOX) return fail(EB_ERR_UNSUPPORTED, "resident kernel: philox mode only");
  if (!st->betas) return fail(EB_ERR_UNSUPPORTED, "resident kernel: tempered samplers only");
  if (st->inds || c.per || args.sa.gmask) return fail(EB_ERR_UNSUPPORTED, "resident kernel: no leaf flags, periodic parameters or Gibbs splits");
  if (c.T < 2 || c.T > RES_MAX_T) return fail(EB_ERR_UNSUPPORTED, "resident kernel: 2 <= ntemps <= %d", RES_MAX_T);
  if (c.LD > 16) return fail(EB_ERR_UNSUPPORTED, "resident kernel: rows of up to 16 doubles");
  if (st->temp_offset != 0) return fail(EB_ERR_UNSUPPORTED, "resident kernel: single-GPU states");
  if (srng->iter_dev != wrng->iter_dev || (!srng->iter_dev && srng->iter != wrng->iter))
    return fail(EB_ERR_INVALID, "resident kernel: move and pass must share the iteration counter");
  args.niter = niter;
  args.permute = wrng->permute;
  args.wseed_lo = (uint32_t)(wrng->seed & 0xFFFFFFFFull); args.wseed_hi = (uint32_t)(wrng->seed >> 32);
  args.adapt_on = adapt != nullptr;
  args.adaptive = adapt ? adapt->adaptive : 0;
  args.stop_adaptation = adapt ? adapt->stop_adaptation : -1;
  args.lag = adapt ? adapt->adaptation_lag : 10000.0;
  args.t0 = adapt ? adapt->adaptation_time : 100.0;
  args.ctrl = ctrl;
  args.sa.accepted = accepted; args.sa.accepted_count = accepted_count;
  const bool launch = niter > 0;
  if (launch) {
    if (!accepted) return fail(EB_ERR_INVALID, "accepted is NULL");
    if (!scratch || scratch_bytes < eb_resident_scratch_bytes(st))
      return fail(EB_ERR_INVALID, "resident kernel: scratch of %zu bytes needed", eb_resident_scratch_bytes(st));
    unsigned char* sc = (unsigned char*)scratch;
    const size_t n = (size_t)c.T * c.W;
    args.g_bar = (unsigned*)sc;
    args.g_counts = (int*)(sc + 64);
    args.g_src = (int32_t*)(sc + RES_HEADER_BYTES);
    args.g_rec = (double*)(sc + RES_HEADER_BYTES + ((n * sizeof(int32_t) + 15) & ~(size_t)15));
    EB_CUDA(cudaMemsetAsync(sc, 0, RES_HEADER_BYTES, (cudaStream_t)stream));
  }
  ResidentPlan plan;
  switch (like->kind) {
    case 0: rc = resident_like<0>(args, plan, launch, (cudaStream_t)stream); break;
    case 1: rc = resident_like<1>(args, plan, launch, (cudaStream_t)stream); break;
    default: rc = resident_like<2>(args, plan, launch, (cudaStream_t)stream); break;
  }
  if (rc) return rc;
  return launch ? check_launch("resident") : EB_OK;
}

int eb_stretch_propose(const eb_state* st, double a, int32_t split, const eb_stretch_rng* rng, const double* period,
                       double* q, double* factors, int32_t* sub_out, void* stream) {
  StretchArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  args.c.per = period;
  rc = fill_stretch_args(args, st, a, rng, false);
  if (rc) return rc;
  if (split != 0 && split != 1) return fail(EB_ERR_INVALID, "split must be 0 or 1 (nsplits == 2)");
  if (rng->gibbs_mask) return fail(EB_ERR_UNSUPPORTED, "Gibbs splits run in the fused kernels");
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
#else
}  // namespace eb
#endif  // EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
