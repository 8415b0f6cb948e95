// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
//
// K1L: the stretch step with one walker spread over LPW lanes (2 or 4) — the kernel of the HBM-sized shapes.
//
// Why: with one thread per walker a 20-double row, its partner and the proposal live in registers (128 per thread at
// config 4: two CTAs per SM, 22 % warps active, DRAM at 28 % of the peak — profiles/r01_ncu_full_c4_k1_summary.txt).
// Here a lane keeps D/LPW elements of each row, so the footprint per thread drops below 64 registers and four or more
// CTAs are resident per SM; the dependent chains of the likelihood shorten by the same factor.
//
//   draw    one walker per LANE, exactly as in stretch_step_kernel (one Philox block: partner index, stretch uniform,
//           accept uniform; keyed bijection of the red/blue split; both logarithms) — no redundant random work;
//   rounds  the 32 walkers of a warp are finished in LPW rounds of 32/LPW walkers: the group of LPW lanes g takes the
//           draws of walker r*(32/LPW)+g by shuffle, lane `sub` of the group owns elements sub, sub+LPW, sub+2*LPW, ...
//           of the row, so that one load instruction of the group reads LPW consecutive doubles (one 32-byte sector at
//           LPW = 4): every sector of a row is requested exactly once;
//   eval    box prior: in-box count reduced over the group; likelihood partial sums reduced with xor-shuffles (the
//           butterfly leaves bit-identical totals on all lanes of the group, so they agree on the Metropolis test);
//   update  every lane stores its elements of an accepted proposal, lane 0 of the group the scalars.
// Reference semantics per line are those cited in stretch_step_kernel (red_blue.py:148-323, stretch.py:74-231,
// move.py:472-703); the reductions change the summation order of the likelihood only (tolerance 1e-10).
#pragma once

#ifndef EB_LANES_MINB
#define EB_LANES_MINB 4   // 64 registers per thread: four CTAs of 256 threads per SM
#endif
#ifndef EB_LANES_UNROLL
#define EB_LANES_UNROLL 2   // two rounds per loop trip: the second round's rows are requested under the first one's tail
#endif

namespace eb {

template <int LPW>
__device__ __forceinline__ double group_sum(unsigned gmask, double v) {
#pragma unroll
  for (int m = 1; m < LPW; m <<= 1) v += __shfl_xor_sync(gmask, v, m);
  return v;
}
template <int LPW>
__device__ __forceinline__ int group_sum_int(unsigned gmask, int v) {
#pragma unroll
  for (int m = 1; m < LPW; m <<= 1) v += __shfl_xor_sync(gmask, v, m);
  return v;
}
template <int LPW>
__device__ __forceinline__ double group_max(unsigned gmask, double v) {
#pragma unroll
  for (int m = 1; m < LPW; m <<= 1) {
    const double o = __shfl_xor_sync(gmask, v, m);
    v = o > v ? o : v;
  }
  return v;
}

// likelihood of a point held as x[e] = element e*LPW + sub; `sp` = the staged functor parameters (stage_store)
template <int KIND>
struct LikeLanes;

template <>
struct LikeLanes<0> {  // EB_LIKE_GAUSSIAN: mu[D], packed S (Like<0>)
  template <int D, int LPW>
  static __device__ __forceinline__ double eval(const double (&x)[D / LPW], int sub, int base, unsigned gmask,
                                                const double* __restrict__ sp, int) {
    constexpr int EPL = D / LPW;
    const double* mu = sp;
    const double* S = sp + D;
    double dfull[D], de[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      de[e] = x[e] - mu[e * LPW + sub];
#pragma unroll
      for (int o = 0; o < LPW; ++o) dfull[e * LPW + o] = __shfl_sync(gmask, de[e], base + o);
    }
    // x^T P x = sum_i d_i (S_ii d_i + sum_{j>i} S_ij d_j); this lane takes the rows i = e*LPW + sub
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int i = e * LPW + sub;
      const double* row = S + sym_row_offset(i, D) - i;   // row[j] = S_ij for j >= i
      double r0 = 0.0, r1 = 0.0;
#pragma unroll
      for (int j = e * LPW; j < D; ++j) {   // j >= e*LPW covers every j >= i; the first `sub` of them lie below the diagonal
        const double sij = j >= i ? row[j] : 0.0;
        if ((j & 1) == 0) r0 = fma(sij, dfull[j], r0);
        else r1 = fma(sij, dfull[j], r1);
      }
      acc = fma(de[e], r0 + r1, acc);
    }
    return -0.5 * group_sum<LPW>(gmask, acc);
  }
};

template <>
struct LikeLanes<1> {  // EB_LIKE_ROSENBROCK
  template <int D, int LPW>
  static __device__ __forceinline__ double eval(const double (&x)[D / LPW], int sub, int base, unsigned gmask,
                                                const double* __restrict__, int) {
    constexpr int EPL = D / LPW;
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      // x[j+1]: the same e of the next lane, or element (e+1)*LPW of lane 0 of the group
      const double a_next = __shfl_sync(gmask, x[e], base + ((sub + 1) % LPW));
      const double b_next = __shfl_sync(gmask, x[e + 1 < EPL ? e + 1 : e], base);
      const double xn = sub == LPW - 1 ? b_next : a_next;
      const int j = e * LPW + sub;
      if (j < D - 1) {
        const double a = xn - x[e] * x[e];
        const double b = 1.0 - x[e];
        acc += 100.0 * (a * a) + b * b;
      }
    }
    return -group_sum<LPW>(gmask, acc);
  }
};

template <>
struct LikeLanes<2> {  // EB_LIKE_GMIX: logc[K], hinv[K], mu[K*D]
  template <int D, int LPW>
  static __device__ __forceinline__ double partial(const double (&x)[D / LPW], int sub, const double* __restrict__ mu) {
    constexpr int EPL = D / LPW;
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const double dd = x[e] - mu[e * LPW + sub];
      if ((e & 1) == 0) r0 = fma(dd, dd, r0);
      else r1 = fma(dd, dd, r1);
    }
    return r0 + r1;
  }
  template <int D, int LPW>
  static __device__ __forceinline__ double eval(const double (&x)[D / LPW], int sub, int base, unsigned gmask,
                                                const double* __restrict__ sp, int K) {
    const double* logc = sp;
    const double* hinv = sp + K;
    const double* mu = sp + 2 * K;
    if (K == LPW) {
      // one component per lane: after the reduce-scatter lane `sub` holds |x - mu_sub|^2, so the group evaluates K
      // exponentials in parallel (two-pass log-sum-exp, max and sum by butterfly)
      double part[LPW];
#pragma unroll
      for (int k = 0; k < LPW; ++k) part[k] = partial<D, LPW>(x, sub, mu + k * D);
      double mine;
      if (LPW == 4) {
        // step 1 (xor 2): lanes with bit 1 clear keep components {0,1}, the others {2,3}
        const bool hi2 = (sub & 2) != 0;
        const double s0 = hi2 ? part[0] : part[2], s1 = hi2 ? part[1] : part[3];
        const double k0 = (hi2 ? part[2] : part[0]) + __shfl_xor_sync(gmask, s0, 2);
        const double k1 = (hi2 ? part[3] : part[1]) + __shfl_xor_sync(gmask, s1, 2);
        // step 2 (xor 1): even lanes keep the first of their pair, odd lanes the second
        const bool hi1 = (sub & 1) != 0;
        mine = (hi1 ? k1 : k0) + __shfl_xor_sync(gmask, hi1 ? k0 : k1, 1);
      } else {  // LPW == 2
        const bool hi1 = (sub & 1) != 0;
        mine = (hi1 ? part[1] : part[0]) + __shfl_xor_sync(gmask, hi1 ? part[0] : part[1], 1);
      }
      const double e = logc[sub] - mine * hinv[sub];
      const double m = group_max<LPW>(gmask, e);
      const double s = group_sum<LPW>(gmask, exp(e - m));
      return m + log(s);
    }
    double m = neg_inf(), s = 0.0;
    for (int k = 0; k < K; ++k) {   // any K: totals on every lane, online log-sum-exp as in Like<2>
      const double tot = group_sum<LPW>(gmask, partial<D, LPW>(x, sub, mu + k * D));
      const double e = logc[k] - tot * hinv[k];
      if (e > m) {
        s = s * exp(m - e) + 1.0;
        m = e;
      } else {
        s += exp(e - m);
      }
    }
    return m + log(s);
  }
};

template <int D, int LPW, int LIKE, bool PHILOX>
__global__ void __launch_bounds__(STRETCH_HALF_THREADS, EB_LANES_MINB) stretch_lanes_kernel(const StretchArgs p) {
  constexpr int EPL = D / LPW;        // elements per lane
  constexpr int WPR = 32 / LPW;       // walkers per round
  static_assert(D % LPW == 0 && (LPW == 2 || LPW == 4), "row length must split evenly over 2 or 4 lanes");
  extern __shared__ __align__(16) double sm[];
  const Common& c = p.c;
  const int t = blockIdx.y;
  const int s = p.split;
  unsigned long long it = p.iter;
  if (PHILOX && p.iter_dev) it = *reinterpret_cast<const volatile unsigned long long*>(p.iter_dev);
  if (s == 1) pdl_launch_dependents();   // same chaining protocol as stretch_step_kernel
  stage_params(c, sm);   // constant parameters (not written by any kernel): staged before the draws, off the registers
  RngKey key;
  Feistel sig;
  if (PHILOX) {
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    if (p.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)(c.t0 + t), (uint32_t)c.W);
  }
  // ---- draw: one walker per lane ------------------------------------------------------------------------------
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < p.Ns[s];
  int w = 0, wc = 0;
  double zz = 1.0, factors = 0.0, log_u = 0.0;
  if (live) {
    double u_z, u_acc;
    stretch_draw<PHILOX>(p, key, sig, t, k, s, w, wc, u_z, u_acc);
    zz = (p.a - 1.0) * u_z + 1.0;                                              // stretch.py:129-132
    zz = zz * zz / p.a;
    factors = ((double)D - 1.0) * log(zz);                                     // stretch.py:223
    log_u = log(u_acc);                                                        // red_blue.py:294
  }
  pdl_wait();
  if (s == 0) pdl_launch_dependents();
  const bool tempered = c.betas != nullptr;
  __shared__ LazyShared lazy_sh;                     // a pass that deferred its ladder adaptation (stretch_step_kernel)
  bool lazy = false;
  if (PHILOX && p.lazy_ctrl && s == 0)
    lazy = lazy_adapt_apply(p.lazy_ctrl, it, c.betas, blockIdx.x == 0 && blockIdx.y == 0, false, lazy_sh);
  const double beta = tempered ? (lazy ? lazy_sh.b[c.t0 + t] : c.betas[t]) : 1.0;   // adapted by the swap pass: read after the wait
  const double* lo = sm;
  const double* hi = sm + D;
  const double* lpdf = sm + 2 * D;
  const double* per = sm + 3 * D;
  const double* sp = sm + PRIOR_ROWS * D;
  double lp_inside = 0.0;                            // every parameter inside its box: the in-order sum (prior.py:369-385)
#pragma unroll
  for (int j = 0; j < D; ++j) lp_inside += lpdf[j];

  const int lane = threadIdx.x & 31, sub = lane % LPW, grp = lane / LPW, base = lane - sub;
  constexpr unsigned gmask = 0xffffffffu;   // control flow is warp-uniform throughout: full-mask shuffles, no WARPSYNC
  const bool periodic = c.per != nullptr;
  const size_t trow = (size_t)t * c.W;

  constexpr int ROUND_UNROLL = EB_LANES_UNROLL;
#pragma unroll ROUND_UNROLL
  for (int r = 0; r < LPW; ++r) {
    const int src = r * WPR + grp;
    const int w_j = __shfl_sync(0xffffffffu, w, src);
    const int wc_j = __shfl_sync(0xffffffffu, wc, src);
    const double zz_j = __shfl_sync(0xffffffffu, zz, src);
    const double f_j = __shfl_sync(0xffffffffu, factors, src);
    const double lu_j = __shfl_sync(0xffffffffu, log_u, src);
    const bool live_j = __shfl_sync(0xffffffffu, (int)live, src) != 0;
    // dead groups (tail of the last CTA) run on walker 0 and store nothing
    const size_t slot = trow + (size_t)w_j;
    const double* own = c.coords + slot * D + sub;
    const double* par = c.coords + (trow + (size_t)wc_j) * D + sub;
    double q[EPL], cc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) q[e] = own[e * LPW];                         // s  (red_blue.py:173-179)
#pragma unroll
    for (int e = 0; e < EPL; ++e) cc[e] = par[e * LPW];                        // c_temp (stretch.py:100)
    const double ll0 = c.logl[slot], lp0 = c.logp[slot];
    const bool active = c.inds ? (c.inds[slot] != 0) : true;
    if (periodic) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {                                          // utils/periodic.py:49-151
        const double P = per[e * LPW + sub], s0 = q[e];
        double diff = cc[e] - s0;
        if (P > 0.0 && fabs(diff) > P / 2.0) {
          const double new_s = diff < 0.0 ? -(P - s0) : (P + s0);
          diff = cc[e] - new_s;
        }
        double v = cc[e] - diff * zz_j;                                        // stretch.py:145
        if (P > 0.0) v = np_mod(v, P);
        q[e] = v;
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) q[e] = cc[e] - (cc[e] - q[e]) * zz_j;      // stretch.py:143-145
    }
    // ---- box prior (prior.py:80-88, :369-385) ----
    int nin = 0, nnan = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const double v = q[e];
      nin += (int)((v >= lo[e * LPW + sub]) & (v <= hi[e * LPW + sub]));
      nnan += (int)(v != v);
    }
    nin = group_sum_int<LPW>(gmask, nin | (nnan << 16));
    double lp = (nin & 0xFFFF) == D ? lp_inside : neg_inf();
    if (__any_sync(0xffffffffu, (nin >> 16) != 0)) {
      // NaN coordinate somewhere in the warp (rare): that parameter contributes 0, the others as usual, added in index
      // order; every group runs the ordered sum so that the shuffles stay warp-uniform
      double lpn = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const int e = j / LPW;
        const double v = q[e];
        double tj = 0.0;
        if (v >= lo[e * LPW + sub] && v <= hi[e * LPW + sub]) tj = lpdf[e * LPW + sub];
        if (v < lo[e * LPW + sub] || v > hi[e * LPW + sub]) tj = neg_inf();
        lpn += __shfl_sync(0xffffffffu, tj, base + (j % LPW));
      }
      if ((nin >> 16) != 0) lp = lpn;
    }
    if (!active) lp = 0.0;                                                     // ensemble.py:1207
    // every group evaluates the likelihood (warp-uniform shuffles); proposals outside the prior discard the value
    double ll = LikeLanes<LIKE>::template eval<D, LPW>(q, sub, base, gmask, sp, c.like_ncomp);
    if (ll != ll) ll = FILL_LOGL;                                              // red_blue.py:279-281
    if (isinf(lp) || !active) ll = FILL_LOGL;                                  // ensemble.py:1279-1282, :1486
    const double logP = log_posterior(ll, lp, beta, tempered);                 // red_blue.py:283
    const double prevP = log_posterior(ll0, lp0, beta, tempered);              // red_blue.py:285-290
    const double lnpdiff = f_j + logP - prevP;                                 // red_blue.py:292
    const bool keep = live_j && (lnpdiff > lu_j);                              // red_blue.py:294
    if (keep) {                                                                // move.py:472-703
      double* dst = c.coords + slot * D + sub;
#pragma unroll
      for (int e = 0; e < EPL; ++e) dst[e * LPW] = q[e];
      if (sub == 0) {
        c.logl[slot] = ll;
        c.logp[slot] = isinf(lp) ? 0.0 : lp;
        if (p.accepted_count) p.accepted_count[slot] += 1u;
      }
    }
    if (live_j && sub == 0) p.accepted[slot] = keep ? 1 : 0;
  }
}

}  // namespace eb
