// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include "common.cuh"

namespace eb {

// ================================================================================================
// K3: chain-parallel parallel-tempering swap pass
// ================================================================================================
// The reference walks the ladder hot -> cold; at rung i it pairs (i, iperm[k]) with
// (i-1, i1perm[k]) (tempering.py:515-559).  Both permutations are bijections, so the slots touched
// by successive rungs form W disjoint chains  p_{T-1} -> p_{T-2} -> ... -> p_0  with
// p_{i-1} = i1perm_i[iperm_i^{-1}[p_i]], and every exchange stays inside one chain.  One group
// of 8 lanes owns a chain: it carries the walker that is bubbling down in registers, decides each
// rung from logl alone (:538) and rewrites only the slots whose content changed, in place.
struct SwapArgs {
  Common c;
  int philox, permute;
  const int32_t* next_pos; const double* u_at;  // replay pair map [T][W]
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  eb_ctrl* ctrl;
  int adapt_on, adaptive, stop_adaptation; double lag, t0;
};

struct PairSrc {  // per-rung description of "who is my partner one rung colder"
  const SwapArgs* p;
  RngKey key;
};

template <bool PHILOX>
__device__ __forceinline__ void pair_lookup(const SwapArgs& p, const RngKey& key, int rung, int pos, int& pos_next,
                                            double& u) {
  if (PHILOX) {
    if (p.permute) {
      Feistel sig;
      sig.init(key, TAG_SWAP_KEY, (uint32_t)rung, (uint32_t)p.c.W);
      pos_next = (int)sig((uint32_t)pos);
    } else {
      pos_next = pos;
    }
    const uint4 r = stream(key, TAG_SWAP_U, (uint32_t)pos, (uint32_t)rung);
    u = u01_52(r.x, r.y);
  } else {
    const size_t i = (size_t)rung * p.c.W + pos;
    pos_next = p.next_pos[i];
    u = p.u_at[i];
  }
}

constexpr int CHAIN_LANES = 8;

template <int NPL>
struct WalkerRegs {
  double x[NPL];
  double ll, lp;
  uint32_t inds;  // this lane's leaf flags (bit j = leaf lane + 8 j)
};

template <int NPL>
__device__ __forceinline__ void load_walker(const Common& c, int rung, int pos, int lane, WalkerRegs<NPL>& r) {
  const size_t slot = (size_t)rung * c.W + pos;
  const double* row = c.coords + slot * c.LD;
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int e = lane + CHAIN_LANES * j;
    r.x[j] = (e < c.LD) ? row[e] : 0.0;
  }
  r.ll = c.logl[slot];
  r.lp = c.logp[slot];
  r.inds = 0;
  if (c.inds) {
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int l = lane + CHAIN_LANES * j;
      if (l < c.L && c.inds[slot * c.L + l]) r.inds |= (1u << j);
    }
  }
}

template <int NPL>
__device__ __forceinline__ void store_walker(const Common& c, int rung, int pos, int lane, const WalkerRegs<NPL>& r) {
  const size_t slot = (size_t)rung * c.W + pos;
  double* row = c.coords + slot * c.LD;
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int e = lane + CHAIN_LANES * j;
    if (e < c.LD) row[e] = r.x[j];
  }
  if (lane == 0) {
    c.logl[slot] = r.ll;
    c.logp[slot] = r.lp;
  }
  if (c.inds) {
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int l = lane + CHAIN_LANES * j;
      if (l < c.L) c.inds[slot * c.L + l] = (r.inds >> j) & 1u;
    }
  }
}

// tempering.py:563-596 on one thread (T <= 256)
__device__ void adapt_ladder(const SwapArgs& p, eb_ctrl* ctrl) {
  const int T = p.c.T;
  double* betas = p.c.betas;
  const double time = (double)ctrl->time;
  if (p.stop_adaptation < 0 || ctrl->time < (long long)p.stop_adaptation) {
    const double decay = p.lag / (time + p.lag);                             // :571
    const double kappa = decay / p.t0;                                       // :572
    const double nw = (double)p.c.W;
    const double inv_b0 = 1.0 / betas[0];
    double cum = 0.0;
    double b_prev_old = betas[0];
    // deltaTs[j] = (1/betas[j+1] - 1/betas[j]) * exp(kappa*(ratios[j]-ratios[j+1])),  j = 0..T-3
    for (int j = 0; j + 2 < T; ++j) {
      const double bj1_old = betas[j + 1];
      const double r0 = (double)ctrl->swaps_accepted[j] / nw;                // :587
      const double r1 = (double)ctrl->swaps_accepted[j + 1] / nw;
      const double dS = kappa * (r0 - r1);                                   // :575
      double dT = 1.0 / bj1_old - 1.0 / b_prev_old;                          // :578
      dT = dT * exp(dS);                                                     // :579
      cum = cum + dT;                                                        // np.cumsum
      const double bnew = 1.0 / (cum + inv_b0);                              // :580
      betas[j + 1] = bj1_old + (bnew - bj1_old);                             // :583, :593
      b_prev_old = bj1_old;
    }
  }
  ctrl->time += 1;                                                           // :596
}

template <int NPL, bool PHILOX>
__global__ void __launch_bounds__(BLOCK) pt_swap_kernel(const SwapArgs p) {
  const Common& c = p.c;
  __shared__ int s_cnt[EB_MAX_TEMPS];
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < c.T; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();

  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = gtid / CHAIN_LANES;
  const int lane = gtid % CHAIN_LANES;
  if (chain < c.W) {
    const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
    const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
    const int T = c.T;
    int pos = chain;                       // p_{T-1}
    int origin = T - 1;
    WalkerRegs<NPL> carry, nxt, nn;
    load_walker<NPL>(c, T - 1, pos, lane, carry);
    int pos_n; double u_i;
    pair_lookup<PHILOX>(p, key, T - 1, pos, pos_n, u_i);
    load_walker<NPL>(c, T - 2, pos_n, lane, nxt);
    for (int i = T - 1; i >= 1; --i) {
      int pos_nn = 0; double u_n = 0.5;
      if (i >= 2) {                        // prefetch the partner of the next rung before any store
        pair_lookup<PHILOX>(p, key, i - 1, pos_n, pos_nn, u_n);
        load_walker<NPL>(c, i - 2, pos_nn, lane, nn);
      }
      const double dbeta = c.betas[i - 1] - c.betas[i];                      // tempering.py:518-522
      const double paccept = dbeta * (carry.ll - nxt.ll);                    // :538
      const bool sel = paccept > log(u_i);                                   // :535, :541
      if (sel) {
        store_walker<NPL>(c, i, pos, lane, nxt);                             // (i-1) walker moves up
        if (lane == 0) atomicAdd(&s_cnt[i - 1], 1);                          // :542
      } else {
        if (origin != i) store_walker<NPL>(c, i, pos, lane, carry);          // carried walker settles here
        carry = nxt;
        origin = i - 1;
      }
      pos = pos_n; pos_n = pos_nn; u_i = u_n; nxt = nn;
    }
    if (origin != 0) store_walker<NPL>(c, 0, pos, lane, carry);
  }
  __syncthreads();
  eb_ctrl* ctrl = p.ctrl;
  for (int i = threadIdx.x; i < c.T - 1; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&ctrl->swaps_work[i], s_cnt[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(&ctrl->ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    for (int i = 0; i < c.T - 1; ++i) {
      const int v = atomicExch(&ctrl->swaps_work[i], 0);
      ctrl->swaps_accepted[i] = v;
      ctrl->swaps_total[i] += (unsigned long long)v;
    }
    if (p.adapt_on && p.adaptive && c.T > 1) adapt_ladder(p, ctrl);          // tempering.py:632-633
    ctrl->iter += 1ull;
    ctrl->ticket = 0u;
  }
}

// K3r: replay mode — turn the host permutations of every rung into a per-position pair map.
__global__ void __launch_bounds__(BLOCK) pt_pairmap_kernel(const int32_t* __restrict__ iperm,
                                                           const int32_t* __restrict__ i1perm,
                                                           const double* __restrict__ u, int T, int W,
                                                           int32_t* __restrict__ next_pos, double* __restrict__ u_at) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= T * W || tid < W) return;  // row 0 unused
  const int rung = tid / W;
  const int p = iperm[tid];
  next_pos[(size_t)rung * W + p] = i1perm[tid];
  u_at[(size_t)rung * W + p] = u[tid];
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_pt_swap(const eb_state* st, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl, void* stream) {
  SwapArgs args;
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  if (!rng || !ctrl) return fail(EB_ERR_INVALID, "rng/ctrl is NULL");
  if (!st->betas) return fail(EB_ERR_INVALID, "swap pass needs betas");
  cudaStream_t s = (cudaStream_t)stream;
  const int T = args.c.T, W = args.c.W;
  if (T < 2) return eb_advance_iter(ctrl, stream);  // range(ntemps-1, 0, -1) is empty
  if (args.c.LD > CHAIN_LANES * 4 || args.c.L > CHAIN_LANES * 4)
    return fail(EB_ERR_UNSUPPORTED, "swap kernel covers nleaves*ndim <= %d (got %d)", CHAIN_LANES * 4, args.c.LD);
  args.philox = rng->mode == EB_RNG_PHILOX; args.permute = rng->permute;
  args.next_pos = rng->next_pos; args.u_at = rng->u_at;
  args.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); args.seed_hi = (uint32_t)(rng->seed >> 32);
  args.iter_dev = (const unsigned long long*)rng->iter_dev; args.iter = rng->iter;
  args.ctrl = ctrl;
  args.adapt_on = adapt != nullptr;
  args.adaptive = adapt ? adapt->adaptive : 0;
  args.stop_adaptation = adapt ? adapt->stop_adaptation : -1;
  args.lag = adapt ? adapt->adaptation_lag : 10000.0;
  args.t0 = adapt ? adapt->adaptation_time : 100.0;
  if (!args.philox) {
    if (rng->mode != EB_RNG_REPLAY) return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
    if (!rng->iperm || !rng->i1perm || !rng->u || !rng->next_pos || !rng->u_at)
      return fail(EB_ERR_INVALID, "replay mode needs iperm, i1perm, u and the next_pos/u_at scratch");
    const int n = T * W;
    pt_pairmap_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(rng->iperm, rng->i1perm, rng->u, T, W, rng->next_pos,
                                                               rng->u_at);
    rc = check_launch("pt_pairmap");
    if (rc) return rc;
  }
  const int npl = (max(args.c.LD, args.c.L) + CHAIN_LANES - 1) / CHAIN_LANES;
  const int nthreads = W * CHAIN_LANES;
  const int grid = (nthreads + BLOCK - 1) / BLOCK;
#define SW_(N)                                                                     \
  if (args.philox) pt_swap_kernel<N, true><<<grid, BLOCK, 0, s>>>(args);           \
  else pt_swap_kernel<N, false><<<grid, BLOCK, 0, s>>>(args)
  switch (npl) {
    case 1: SW_(1); break;
    case 2: SW_(2); break;
    case 3: SW_(3); break;
    default: SW_(4); break;
  }
#undef SW_
  return check_launch("pt_swap");
}

}  // extern "C"
