// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include <cstdlib>

#include "common.cuh"

namespace eb {

// ================================================================================================
// K3: chain-parallel parallel-tempering swap pass
// ================================================================================================
// The reference walks the ladder hot -> cold; at rung i it pairs (i, iperm[k]) with
// (i-1, i1perm[k]) (tempering.py:515-559).  Both permutations are bijections, so the slots touched
// by successive rungs form W disjoint chains  p_{T-1} -> p_{T-2} -> ... -> p_0  and every exchange
// stays inside one chain.  The accept test needs logl only (:538), so a chain is resolved by a group of
// CL lanes (8, 16 or 32: one or two rungs per lane):
//   0. (before the grid-dependency wait, overlapping the move kernel) positions p_r of the chain on every rung and
//      log(u) of every rung (philox mode: p_r = sigma_r(chain) with one keyed bijection per rung, so the pairing of
//      rung i is sigma_{i-1} o sigma_i^{-1}, a uniform random bijection exactly like i1perm o iperm^{-1});
//   1. gather logl[r][p_r] (replay mode: lane 0 first follows the host pair map);
//   2. the hot -> cold cascade on those T numbers, run by every lane of the chain (broadcast shared-memory reads),
//      leaving the accept bits in registers; src[r] = rung whose walker ends on rung r follows from the bits;
//   3. only the rows with src[r] != r move, in place: every lane gathers the source rows of its rungs into
//      registers, one warp-level sync, then writes (long rows / leaf flags: through a global staging buffer).
// The per-rung swap counts are folded and the ladder adapted (tempering.py:563-596) by the warp that draws the last
// ticket, by an extra CTA (sharded passes) or by the next move kernel (deferred adaptation): see the count publication
// below and common.cuh:lazy_adapt_apply.
struct SwapArgs {
  Common c;                                       // local state (sharded: the DESTINATION buffers of this rank)
  int T;                                          // rungs of the full ladder
  const double* logl_in;                          // [T][W] log-likelihoods the cascade reads
  double* betas;                                  // [T] full ladder (adapted in place)
  // temperature-sharded run (eb_pt_swap_sharded): this rank writes rungs [t_lo, t_hi) into c.*, reading
  // the source rows from the rank that owns them (peer-mapped pointers)
  int sharded, world, t_lo, t_hi;
  int temp_begin[EB_MAX_RANKS + 1];
  const double* coords_src[EB_MAX_RANKS]; const double* logp_src[EB_MAX_RANKS]; const uint8_t* inds_src[EB_MAX_RANKS];
  const unsigned long long* flags;                // sharded: local flag words raised by every rank's publish kernel
  // fused publish (eb_shard.pub_*): the pass itself all-gathers logl as self-validating 16-byte units
  int rank, pdl;
  const double* pub_src; uint4* pub_dst[EB_MAX_RANKS]; const uint4* ll_in;
  // row mail (eb_shard.mail_*): rows that change rank are PUSHED by the rank that owns the source rung
  uint4* mail_dst[EB_MAX_RANKS]; const uint4* mail_in;
  int philox, permute, cpb;                       // cpb = chains per block
  double* scratch_coords; double* scratch_logp; uint8_t* scratch_inds;   // staging of moved rows (RR == 0 path)
  const int32_t* next_pos; const double* u_at;  // replay pair map [T][W]
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  eb_ctrl* ctrl;
  int adapt_on, adaptive, stop_adaptation; double lag, t0;
  int dbg_skip;
  int defer;                                      // single-GPU pass: leave counts + snapshot for a lazy adaptation
};

// CTAs of the publish kernel for `ndoubles` of logl (must match k_shard.cu: a flag word counts publishing CTAs)
__host__ __device__ inline int publish_grid(size_t ndoubles) {
  size_t g = (ndoubles / 2 + 255) / 256;
  return g < 1 ? 1 : g > 148 ? 148 : (int)g;
}

// CTAs of a swap pass that take part in the fused publish of `ndoubles` of logl (grid of `nreal` chain CTAs)
__host__ __device__ inline int publish_ctas(size_t ndoubles, int nreal) {
  const int g = publish_grid(ndoubles) > 128 ? 128 : publish_grid(ndoubles);
  return g < nreal ? g : nreal;
}

struct SwapLayout {  // byte offsets into dynamic shared memory
  size_t betas, dts, ll, lu, keys, pos, cnt, band, rej, on, total;
};
__host__ __device__ inline SwapLayout swap_layout(int T, int cpb) {
  SwapLayout s;
  size_t o = 0;
  s.betas = o; o += sizeof(double) * T;
  s.dts = o; o += sizeof(double) * T;
  s.ll = o; o += sizeof(double) * T * cpb;
  s.lu = o; o += sizeof(double) * T * cpb;
  s.keys = o; o += sizeof(uint32_t) * FEISTEL_ROUNDS * T;
  s.pos = o; o += sizeof(int) * T * cpb;
  s.cnt = o; o += sizeof(int) * T;
  s.band = o; o += (size_t)T * cpb;
  s.rej = o; o += (size_t)T * cpb;
  s.on = o; o += (size_t)T * cpb;
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// release/acquire fence at device scope (__threadfence() is the sequentially consistent one: MEMBAR.SC + L1 invalidate)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// bounded spins count SM cycles (reading %globaltimer is slow): ~2 s at 2 GHz
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000ll;

__device__ __forceinline__ bool sel_bit(unsigned long long lo, unsigned long long hi, int i) {
  return i < 64 ? ((lo >> i) & 1ull) != 0ull : ((hi >> (i - 64)) & 1ull) != 0ull;
}

// source rung of the walker that ends on rung r, from the accept bits of the cascade (bit i = swap accepted at rung i):
// an accepted swap at r brings up the walker of rung r-1; otherwise rung r receives the walker that was being carried
// down, which started at the top of the run of accepted swaps directly above r
__device__ __forceinline__ int swap_source(unsigned long long lo, unsigned long long hi, int r, int T) {
  if (r >= 1 && sel_bit(lo, hi, r)) return r - 1;
  int o = r;
  while (o + 1 < T && sel_bit(lo, hi, o + 1)) ++o;
  return o;
}

// the per-walker byte payload that travels with a row (leaf flags, friend table): word copies when aligned
__device__ __forceinline__ void copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int n) {
  if (((n | (int)(reinterpret_cast<size_t>(dst) | reinterpret_cast<size_t>(src))) & 3) == 0) {
    for (int i = 0; i < n; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = *reinterpret_cast<const uint32_t*>(src + i);
  } else {
    for (int i = 0; i < n; ++i) dst[i] = src[i];
  }
}

// copy one row between distinct slots through registers; LD is a runtime value
__device__ __forceinline__ void copy_row(double* __restrict__ dst, const double* __restrict__ src, int LD) {
  if ((LD & 3) == 0) {
    for (int e = 0; e < LD; e += 4) {
      double a, b, c, d;
      ld256(src + e, a, b, c, d);
      st256(dst + e, a, b, c, d);
    }
  } else if ((LD & 1) == 0) {
    for (int e = 0; e < LD; e += 2) *reinterpret_cast<double2*>(dst + e) = *reinterpret_cast<const double2*>(src + e);
  } else {
    for (int e = 0; e < LD; ++e) dst[e] = src[e];
  }
}

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4* p, const uint4 u) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void ld_volatile_d2(const double* p, double& a, double& b) {
  asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}

// Fused publish of a sharded pass (eb_shard.pub_*): the first CTAs of the grid write this rank's logl rows into the LL
// buffer of EVERY rank (own rank included) as 16-byte units {lo32, tag, hi32, tag}, tag = iter+1, with coalesced 16-byte
// NVLink peer stores.  A unit validates itself (each aligned 8-byte half is written atomically and carries the tag), so
// there is no release fence, no flag and no wait for store acknowledgements: the latency of the all-gather is one
// one-way NVLink trip.  Buffers alternate with the iteration parity, so a tag can only be confused with the one written
// two iterations earlier, which differs.
__device__ __forceinline__ void publish_ll(const SwapArgs& p, unsigned long long it, int nreal) {
  const size_t n = (size_t)(p.t_hi - p.t_lo) * p.c.W, off = (size_t)p.t_lo * p.c.W;
  const int npub = publish_ctas(n, nreal);
  if ((int)blockIdx.x >= npub) return;
  const uint32_t tag = (uint32_t)(it + 1ull);
  const size_t stride = (size_t)npub * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = p.pub_src[i];
    const uint4 u = make_uint4((uint32_t)__double2loint(v), tag, (uint32_t)__double2hiint(v), tag);
    for (int gr = 0; gr < p.world; ++gr) st_volatile_u4(p.pub_dst[gr] + off + i, u);
  }
}

// slots the CTAs spread their partial swap counts and arrivals over (same-address atomics serialise in L2, but the
// adapt CTA reads slots x (T-1) words): fewer slots for long ladders
__device__ __forceinline__ int swap_slots(int T) { return 8; }

// The adapt CTA of the swap pass: wait for the counts of all `nreal` chain CTAs, fold them into swaps_accepted and
// apply adapt_temps (tempering.py:563-596); ticks the iteration counter.
__device__ __forceinline__ void pt_swap_adapt(const SwapArgs& p, int T, int W, int nreal, unsigned long long it,
                                              long long time_now_t0, double* s_betas, double* s_dts, int* s_cnt) {
  eb_ctrl* ctrl = p.ctrl;
  const int tid = threadIdx.x;
  __shared__ long long s_time;
  __shared__ int s_ok2;
  if (tid == 0) { s_time = time_now_t0; s_ok2 = 1; }
  for (int r = tid; r < T; r += blockDim.x) s_cnt[r] = 0;
  __syncthreads();
  const int NS = swap_slots(T);
  if (tid < NS) {
    const unsigned expected = (unsigned)(nreal / NS + (tid < nreal % NS ? 1 : 0));
    const volatile unsigned* a = &ctrl->arrive[tid];
    const long long t_start = clock64();
    while (*a < expected)
      if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) { s_ok2 = 0; break; }
    fence_acq_rel_gpu();
    ctrl->arrive[tid] = 0u;
  }
  __syncthreads();
  EB_MARK(28);
  if (!s_ok2) {
    // A CTA never arrived (bounded wait ran out: a debugger / profiler replay, a wedged SM).  The error word is sticky —
    // the host raises at its next look at the control block (DeviceContext.download, the sampler's yield points, the
    // staged stores) — and the control block is left CONSISTENT: counts and arrivals zeroed, iter == iter_next, no
    // adaptation from partial counts, so later kernels key their draws alike.
    if (tid == 0) atomicExch(&ctrl->error, EB_DEVERR_SWAP_TIMEOUT);
    for (int e = tid; e < EB_SWAP_SLOTS * (EB_MAX_TEMPS); e += blockDim.x) ctrl->swaps_work[e / EB_MAX_TEMPS][e % EB_MAX_TEMPS] = 0;
    if (tid < EB_SWAP_SLOTS) ctrl->arrive[tid] = 0u;
    for (int r = tid; r < T - 1; r += blockDim.x) ctrl->swaps_accepted[r] = 0;
    if (tid == 0) ctrl->iter = it + 1ull;
    return;
  }
  // fold the slot counts (independent loads first, the dependent bookkeeping stores last: the ladder is what the next
  // kernel waits for)
  for (int e = tid; e < NS * (T - 1); e += blockDim.x) {
    const int v = *reinterpret_cast<volatile int*>(&ctrl->swaps_work[e / (T - 1)][e % (T - 1)]);
    if (v) atomicAdd(&s_cnt[e % (T - 1)], v);
  }
  __syncthreads();
  EB_MARK(29);
  const long long time_now = s_time;
  if (p.adapt_on && p.adaptive && T > 1) {                                     // tempering.py:632-633
    if (p.stop_adaptation < 0 || time_now < (long long)p.stop_adaptation) {   // :590
      const double decay = p.lag / ((double)time_now + p.lag);                 // :571
      const double kappa = decay / p.t0;                                       // :572
      const double nw = (double)W;
      // deltaTs[j] = (1/betas[j+1] - 1/betas[j]) * exp(kappa*(ratios[j]-ratios[j+1])),  j = 0..T-3
      for (int j = tid; j + 2 < T; j += blockDim.x) {
        const double r0 = (double)s_cnt[j] / nw, r1 = (double)s_cnt[j + 1] / nw;   // :587
        const double dS = kappa * (r0 - r1);                                   // :575
        double dT = 1.0 / s_betas[j + 1] - 1.0 / s_betas[j];                   // :578
        s_dts[j] = dT * exp(dS);                                               // :579
      }
      __syncthreads();
      if (tid == 0) {                                                          // np.cumsum: sequential adds, in order
        double cum = 0.0;
        for (int j0 = 0; j0 + 2 < T; j0 += 8) {                                // operands loaded ahead of the dependent adds
          double v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = j0 + k + 2 < T ? s_dts[j0 + k] : 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (j0 + k + 2 < T) { cum = cum + v[k]; s_dts[j0 + k] = cum; }
        }
      }
      __syncthreads();
      const double inv_b0 = 1.0 / s_betas[0];
      for (int j = tid; j + 2 < T; j += blockDim.x) {
        const double bold = s_betas[j + 1];
        const double bnew = 1.0 / (s_dts[j] + inv_b0);                         // :580
        p.betas[j + 1] = bold + (bnew - bold);                                 // :583, :593
      }
    }
    if (tid == 0) ctrl->time = time_now + 1;                                   // :596
  }
  EB_MARK(30);
  for (int e = tid; e < NS * (T - 1); e += blockDim.x) ctrl->swaps_work[e / (T - 1)][e % (T - 1)] = 0;
  for (int r = tid; r < T - 1; r += blockDim.x) {
    const int v = s_cnt[r];
    ctrl->swaps_accepted[r] = v;
    ctrl->swaps_total[r] += (unsigned long long)v;
  }
  if (tid == 0) ctrl->iter = it + 1ull;
}

// The same fold + adapt_temps done by ONE WARP: the warp of the chain CTA that published its counts last (ticket
// election) runs it straight away — no extra CTA polling arrival words, no block barriers — while the other warps of that
// CTA and all other CTAs go on moving rows.  Every operand and every operation is the one of pt_swap_adapt, so the ladder
// is bit-identical.  `s_betas` / `s_dts` / `s_cnt` are the CTA's shared arrays; no other warp touches them after the
// block barrier that precedes the count publication.
__device__ __forceinline__ void pt_swap_adapt_warp(const SwapArgs& p, int T, int W, unsigned long long it, double* s_betas,
                                                   double* s_dts, int* s_cnt) {
  eb_ctrl* ctrl = p.ctrl;
  const int lane = threadIdx.x & 31;
  const int row0 = (int)(it & 1ull) * LAZY_SLOTS;   // single-GPU passes alternate between two sets of count rows (common.cuh)
  // fold the slot counts: all loads of a lane in flight together, then the sums; the rows are zeroed for the passes to
  // come — both parities: the other one may hold the counts of a deferred pass that has been applied since
  const long long time_now = *reinterpret_cast<const volatile long long*>(&ctrl->time);
  for (int r = lane; r < T - 1; r += 32) {
    int w8[LAZY_SLOTS];
#pragma unroll
    for (int sl = 0; sl < LAZY_SLOTS; ++sl) w8[sl] = *reinterpret_cast<volatile int*>(&ctrl->swaps_work[row0 + sl][r]);
    int v = 0;
#pragma unroll
    for (int sl = 0; sl < LAZY_SLOTS; ++sl) v += w8[sl];
    s_cnt[r] = v;
#pragma unroll
    for (int sl = 0; sl < 2 * LAZY_SLOTS; ++sl) ctrl->swaps_work[sl][r] = 0;
  }
  __syncwarp();
  if (p.adapt_on && p.adaptive && T > 1) {                                     // tempering.py:632-633
    if (p.stop_adaptation < 0 || time_now < (long long)p.stop_adaptation) {   // :590
      const double decay = p.lag / ((double)time_now + p.lag);                 // :571
      const double kappa = decay / p.t0;                                       // :572
      const double nw = (double)W;
      for (int j = lane; j + 2 < T; j += 32) {
        const double r0 = (double)s_cnt[j] / nw, r1 = (double)s_cnt[j + 1] / nw;   // :587
        const double dS = kappa * (r0 - r1);                                   // :575
        const double dT = 1.0 / s_betas[j + 1] - 1.0 / s_betas[j];             // :578
        s_dts[j] = dT * exp(dS);                                               // :579
      }
      __syncwarp();
      if (lane == 0) {                                                         // np.cumsum: sequential adds, in order
        double cum = 0.0;
        for (int j0 = 0; j0 + 2 < T; j0 += 8) {                                // operands loaded ahead of the dependent adds
          double v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = j0 + k + 2 < T ? s_dts[j0 + k] : 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (j0 + k + 2 < T) { cum = cum + v[k]; s_dts[j0 + k] = cum; }
        }
      }
      __syncwarp();
      const double inv_b0 = 1.0 / s_betas[0];
      for (int j = lane; j + 2 < T; j += 32) {
        const double bold = s_betas[j + 1];
        const double bnew = 1.0 / (s_dts[j] + inv_b0);                         // :580
        p.betas[j + 1] = bold + (bnew - bold);                                 // :583, :593
      }
    }
    if (lane == 0) ctrl->time = time_now + 1;                                  // :596
  }
  for (int r = lane; r < T - 1; r += 32) {
    const int v = s_cnt[r];
    ctrl->swaps_accepted[r] = v;
    atomicAdd(reinterpret_cast<unsigned long long*>(&ctrl->swaps_total[r]), (unsigned long long)v);   // no round trip
  }
  if (lane == 0) {
    ctrl->ticket = 0u;
    ctrl->iter = it + 1ull;
  }
}

// One row copied by a whole warp: coalesced 16 / 8-byte (doubles) or 4 / 1-byte (payload) accesses, every load of the row
// in flight before the first store.  Rows of the staging path are long (config 5: 60 doubles + 1.3 KB of leaf flags and
// friend table per walker); one lane per rung copying them element by element was 228 us per pass there.
__device__ __forceinline__ void warp_copy_doubles(double* __restrict__ dst, const double* __restrict__ src, int n, int wl) {
  if ((n & 1) == 0 && ((reinterpret_cast<size_t>(dst) | reinterpret_cast<size_t>(src)) & 15) == 0) {
    const int n2 = n >> 1;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (int i0 = 0; i0 < n2; i0 += 4 * 32) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + wl; if (i < n2) v[u] = s2[i]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + wl; if (i < n2) d2[i] = v[u]; }
    }
  } else {
    for (int i0 = 0; i0 < n; i0 += 4 * 32) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + wl; if (i < n) v[u] = src[i]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + wl; if (i < n) dst[i] = v[u]; }
    }
  }
}
__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int n, int wl) {
  if (((n | (int)(reinterpret_cast<size_t>(dst) | reinterpret_cast<size_t>(src))) & 3) == 0) {
    const int n4 = n >> 2;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    for (int i0 = 0; i0 < n4; i0 += 12 * 32) {
      uint32_t v[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) { const int i = i0 + u * 32 + wl; if (i < n4) v[u] = s4[i]; }
#pragma unroll
      for (int u = 0; u < 12; ++u) { const int i = i0 + u * 32 + wl; if (i < n4) d4[i] = v[u]; }
    }
  } else {
    for (int i = wl; i < n; i += 32) dst[i] = src[i];
  }
}

constexpr int SWAP_THREADS = 256;
constexpr int SWAP_AGES = 8;      // tests per walker evaluated ahead of the cascade walk (bits of one band byte)

// CL lanes resolve one chain; lane l owns rungs l, l+CL, l+2CL, ...  (CL = 8, 16 or 32).
// RR > 0: rows of up to RR doubles move through registers (needs T <= RPL*CL and no leaf flags); RR == 0: through the
// global staging buffers.
template <bool PHILOX, bool SHARDED, int CL, int RR, int RPL>
__global__ void __launch_bounds__(SWAP_THREADS) pt_swap_kernel(const __grid_constant__ SwapArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const Common& c = p.c;
  const int T = p.T, W = c.W, LD = c.LD, L = c.Lb, cpb = p.cpb;
  const SwapLayout lay = swap_layout(T, cpb);
  double* s_betas = reinterpret_cast<double*>(smraw + lay.betas);
  double* s_dts = reinterpret_cast<double*>(smraw + lay.dts);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smraw + lay.keys);
  int* s_cnt = reinterpret_cast<int*>(smraw + lay.cnt);

  // The grid carries one extra CTA (the last) without chains: it waits until every other CTA has published its swap
  // counts, then folds them and adapts the ladder WHILE the other CTAs are still moving rows.
  const bool adapt_cta = blockIdx.x == gridDim.x - 1;
  const int nreal = (int)gridDim.x - 1;
  const int tid = threadIdx.x;
  const int g = tid / CL, lane = tid % CL;
  const int chain = blockIdx.x * cpb + g;
  bool valid = !adapt_cta && g < cpb && chain < W;
  eb_ctrl* ctrl = p.ctrl;
  // ---- prologue: nothing here reads the walker state, so under programmatic dependent launch it overlaps the move
  //      kernel that precedes this pass.  ctrl->iter is written only by this kernel's own tail.
  if (p.pdl == 2) pdl_wait();
  unsigned long long it = p.iter;   // (not a ?: of the two: with a __grid_constant__ parameter block nvcc merges the
  // arms into ONE global load whose address may then point into parameter space)
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
  if (adapt_cta && tid == 0) {
    // the next move kernel may start its draws while this pass still runs: it keys them by iter_next.  The adapt CTA
    // does it (it has nothing else to do yet): on a chain CTA the fence would delay that CTA's chains — and, in a
    // sharded pass, its share of the publish — by a microsecond, and every rank waits for the slowest chain.
    *reinterpret_cast<volatile unsigned long long*>(&ctrl->iter_next) = it + 1ull;
    fence_acq_rel_gpu();
  }
  const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);

  EB_MARK(16);
  // Ordinary launch (or wait-first dependent): the move kernel has completed, so the rows go out first and the
  // prologue below (positions, log u) runs while they cross NVLink.
  if (SHARDED && p.pub_src && p.pdl != 1) publish_ll(p, it, nreal);
  for (int r = tid; r < T; r += blockDim.x) {
    s_cnt[r] = 0;
    if (PHILOX && p.permute) Feistel::make_keys(key, TAG_SWAP_KEY, (uint32_t)r, s_keys + FEISTEL_ROUNDS * r);
  }
  __syncthreads();

  const int gg = valid ? g : 0;
  double* ll = reinterpret_cast<double*>(smraw + lay.ll) + (size_t)gg * T;
  double* lu = reinterpret_cast<double*>(smraw + lay.lu) + (size_t)gg * T;
  int* pos = reinterpret_cast<int*>(smraw + lay.pos) + (size_t)gg * T;

  EB_MARK(17);
  if (PHILOX && valid && !EB_DBG_SKIP(16)) {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    for (int r = lane; r < T; r += CL) {
      int pz = chain;
      if (p.permute) {
        Feistel sig;
        sig.init_from(s_keys + FEISTEL_ROUNDS * r, (uint32_t)W);
        pz = (int)sig((uint32_t)chain);
      }
      pos[r] = pz;
      // one Philox block serves rungs r and r + 8 of a chain: counter (chain, (r & 7) | ((r >> 4) << 3)),
      // word pair (r >> 3) & 1
      const bool second = ((r >> 3) & 1) != 0;
      if (CL != 8 || !second) q = stream(key, TAG_SWAP_U, (uint32_t)chain, (uint32_t)((r & 7) | ((r >> 4) << 3)));
      const double u = second ? u01_52(q.z, q.w) : u01_52(q.x, q.y);
      lu[r] = log(u);                                                          // tempering.py:535
    }
  }
  pdl_wait();                 // the move kernel has completed; its writes are visible
  pdl_launch_dependents();    // the next move kernel may begin its draws
  if (SHARDED && p.pub_src && p.pdl == 1) publish_ll(p, it, nreal);
  EB_MARK(27);
  if (SHARDED && p.flags) {
    // separate publish kernel (eb_publish_logl): every rank's logl rows of THIS iteration must have landed in logl_in:
    // bounded spin on the local flag words, one thread per CTA
    __shared__ bool s_ok;
    if (tid == 0) {
      bool ok = *reinterpret_cast<volatile unsigned int*>(&ctrl->error) == 0u;
      const long long t_start = clock64();
      for (int gr = 0; ok && gr < p.world; ++gr) {
        const size_t rows = (size_t)(p.temp_begin[gr + 1] - p.temp_begin[gr]) * W;
        // flag word g counts the CTAs of rank g's publish kernels so far (k_shard.cu)
        const unsigned long long target = (it + 1ull) * (unsigned long long)publish_grid(rows);
        const volatile unsigned long long* f = p.flags + gr;
        while (*f < target)
          if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) { ok = false; break; }
      }
      if (!ok) atomicExch(&ctrl->error, EB_DEVERR_PEER_TIMEOUT);
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) valid = false;   // the error is sticky; the CTA still draws its ticket so that the pass ends consistently
  }
  EB_MARK(26);
  for (int r = tid; r < T; r += blockDim.x) {   // the ladder is adapted by the previous pass: read after the wait
    const double b = p.betas[r];
    s_betas[r] = b;
    s_dts[r] = r >= 1 ? p.betas[r - 1] - b : 0.0;                              // :518-522
  }
  if (!PHILOX) {
    if (valid && lane == 0) {            // replay: p_{i-1} = i1perm_i[iperm_i^{-1}[p_i]] (pt_pairmap_kernel)
      int pz = chain;
      pos[T - 1] = pz;
      for (int i = T - 1; i >= 1; --i) {
        pz = p.next_pos[(size_t)i * W + pz];
        pos[i - 1] = pz;
      }
    }
    __syncwarp();
    if (valid)
      for (int r = lane; r < T; r += CL) lu[r] = log((r >= 1) ? p.u_at[(size_t)r * W + pos[r]] : 0.5);
  }
  if (SHARDED && p.ll_in) {
    // fused publish: every logl arrives as a 16-byte unit {lo32, tag, hi32, tag} with tag = iter+1 (publish_ll); a lane
    // polls the units of its rungs (at most four: T <= 128, CL = 32 beyond 16 rungs) until both tags match.  No fence and
    // no flag: each 8-byte half is written atomically and validates itself.
    if (valid) {
      const uint32_t tag = (uint32_t)(it + 1ull);
      bool pend[4];
      uint4 v[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) pend[m] = lane + m * CL < T;
      bool ok = *reinterpret_cast<volatile unsigned int*>(&ctrl->error) == 0u;   // an earlier timeout: do not spin again
      const long long t_start = clock64();
      for (;;) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (pend[m]) {
            const int r = lane + m * CL;
            v[m] = ld_volatile_u4(p.ll_in + (size_t)r * W + pos[r]);
          }
        bool any = false;
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (pend[m]) {
            if ((v[m].y == tag && v[m].w == tag) || !ok) {
              ll[lane + m * CL] = __hiloint2double((int)v[m].z, (int)v[m].x);
              pend[m] = false;
            } else {
              any = true;
            }
          }
        if (!any) break;
        if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) {
          atomicExch(&ctrl->error, EB_DEVERR_PEER_TIMEOUT);
          ok = false;
        }
      }
    }
  } else if (valid) {
    for (int r = lane; r < T; r += CL) ll[r] = p.logl_in[(size_t)r * W + pos[r]];
  }
  __syncthreads();
  // sharded: the rows most likely to leave this rank as mail (my top rung's walker if the swap above it is accepted, my
  // bottom rung's walker if it gets carried down) are requested now, one element per lane, so that their L2 latency runs
  // under the cascade; needs the LD+1 units of a mail to fit the lanes of the chain
  const bool mail_pre = SHARDED && p.mail_in && valid && LD + 1 <= CL;
  double mail_up = 0.0, mail_dn = 0.0;
  if (mail_pre && lane <= LD) {
    if (p.t_hi < T) {
      const size_t sslot = (size_t)(p.t_hi - 1 - p.t_lo) * W + pos[p.t_hi - 1];
      mail_up = lane < LD ? p.coords_src[p.rank][sslot * LD + lane] : p.logp_src[p.rank][sslot];
    }
    if (p.t_lo >= 1) {
      const size_t sslot = (size_t)pos[p.t_lo];
      mail_dn = lane < LD ? p.coords_src[p.rank][sslot * LD + lane] : p.logp_src[p.rank][sslot];
    }
  }
  EB_MARK(18);
  // ---- the cascade, hot -> cold (tempering.py:515-559 restricted to this chain).
  // The carried log-likelihood is always an ORIGINAL value ll[j]: the walker that starts on rung j is tested at rungs
  // j, j-1, ... (test_i(x) = dts[i] * (x - ll[i-1]) > lu[i], :538/:541, s_dts[i] = betas[i-1]-betas[i]) until a swap is
  // rejected at rung s = j - run; it settles there and the walker of rung s-1 is carried on.  None of these tests depends
  // on what the cascade decides, only on WHICH walkers get carried.  Sharded passes and long ladders therefore resolve
  // the chain by walking over the CARRIED walkers only (below), with their first tests evaluated ahead for all walkers at
  // once and long runs extended 32 rungs per step (same expression and rounding as the reference throughout).
  unsigned long long sel_lo = 0ull, sel_hi = 0ull;
  if (RR > 0 && !SHARDED) {
    // single-GPU passes with rows in registers (T <= CL * RPL <= 64): the plain sequential cascade with a compile-time
    // trip count, so that the operand loads of all rungs are hoisted above the dependent chain; every lane runs it on
    // broadcast operands (0.8 us at 16 rungs, 3.9 us at 64).
    if (valid && !EB_DBG_SKIP(8)) {
      double carry = ll[T - 1];
#pragma unroll
      for (int i = CL * RPL - 1; i >= 1; --i) {
        if (i < T) {
          const double lower = ll[i - 1];
          const bool sel = s_dts[i] * (carry - lower) > lu[i];                 // :538, :541  (s_dts[i] = betas[i-1]-betas[i])
          if (sel) {
            if (i < 64) sel_lo |= 1ull << i;
            else sel_hi |= 1ull << (i - 64);
          } else {
            carry = lower;               // the carried walker settles on rung i, rung i-1's walker is carried on
          }
        }
      }
    }
  } else {
    constexpr int NB = CL == 32 ? 4 : 1;              // T <= 128 with 32 lanes, T <= CL otherwise
    unsigned char* sband = smraw + lay.band + (size_t)gg * T;
    unsigned char* s_rej = smraw + lay.rej + (size_t)gg * T;
    // (a) band of walker j: bit a = test at rung j-a with x = ll[j], a < SWAP_AGES — all walkers at once
    if (valid) {
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        const int j = lane + m * CL;
        if (j < T) s_rej[j] = 0;
        if (j >= 1 && j < T) {
          const double x = ll[j];
          unsigned b = 0u;
#pragma unroll
          for (int a = 0; a < SWAP_AGES; ++a) {
            const int i = j - a;
            if (i >= 1) b |= (unsigned)(s_dts[i] * (x - ll[i - 1]) > lu[i]) << a;
          }
          sband[j] = (unsigned char)b;
        }
      }
    }
    __syncwarp();
    // (b) the walk over the CARRIED walkers only: j -> j - run - 1.  Every rung sees exactly one test and the rejected
    //     ones are the rungs where a carried walker settles, so the walk only marks those; the accept bits are the
    //     complement, collected by ballot below.  A run that outlasts the band (hot end of a long ladder: beta ~ 0, every
    //     swap accepted) is extended 32 rungs per step — the tests of one carried value are independent of each other,
    //     so the lanes of the warp evaluate one rung each and vote (CL == 32: one chain per warp, uniform control flow);
    //     groups of 8 / 16 lanes (T <= 16) extend it with the plain sequential tests.
    if (CL == 32) {
      int j = T - 1;
      while (j >= 1) {                                // uniform over the warp
        int run = 0;
        if (valid && !EB_DBG_SKIP(8)) {
          run = __ffs((int)~(unsigned)sband[j]) - 1;
          if (run == SWAP_AGES) {
            const double x = ll[j];
            int base = j - SWAP_AGES;                 // next untested rung
            while (base >= 1) {
              const int i = base - lane;
              const bool f = i >= 1 && s_dts[i] * (x - ll[i - 1]) > lu[i];
              const int n = __ffs((int)~__ballot_sync(0xffffffffu, f)) - 1;   // wins in a row from rung `base` downwards
              if (n < 0) { run += 32; base -= 32; continue; }                 // all 32 won
              run += n;
              break;
            }
          }
          s_rej[j - run] = 1;                         // settles here: this swap was rejected (or rung 0)
        }
        j -= run + 1;                                 // the walker below is carried on
      }
    } else if (valid && !EB_DBG_SKIP(8)) {
      int j = T - 1;
      while (j >= 1) {
        int run = __ffs((int)~(unsigned)sband[j]) - 1;
        if (run == SWAP_AGES) {
          const double x = ll[j];
          int i = j - SWAP_AGES;
          while (i >= 1 && s_dts[i] * (x - ll[i - 1]) > lu[i]) { ++run; --i; }
        }
        s_rej[j - run] = 1;
        j -= run + 1;
      }
    }
    __syncwarp();
    {
      const int sh = (tid & 31) - lane;               // first lane of this chain's group within the warp
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        if (m * CL < T) {                             // uniform
          const int r = lane + m * CL;
          const bool acc = valid && r >= 1 && r < T && s_rej[r] == 0;
          const unsigned v = __ballot_sync(0xffffffffu, acc);
          if (CL == 32) {                             // one chain per warp: ballot m holds rungs 32m .. 32m+31
            if (m == 0) sel_lo |= (unsigned long long)v;
            if (m == 1) sel_lo |= (unsigned long long)v << 32;
            if (m == 2) sel_hi |= (unsigned long long)v;
            if (m == 3) sel_hi |= (unsigned long long)v << 32;
          } else {                                    // several chains per warp: this chain's CL bits
            sel_lo = (unsigned long long)((v >> sh) & ((1u << (CL & 31)) - 1u));
          }
        }
      }
    }
  }

  EB_MARK(19);
  // ---- sharded: rows that change rank leave first, as mail pushed by the rank that owns the source rung (every rank has
  //      resolved the whole chain, so sender and receiver agree without talking): a one-way NVLink trip that runs under
  //      the count publication and the local row copies, instead of the round trip of a pull.  A mail is LD+1
  //      self-validating units (row, then logp) in the receiver's mailbox, slot [direction][chain]; per chain and rank at
  //      most one walker arrives from below (into rung t_lo, when the swap at t_lo is accepted) and at most one from above
  //      (the carried walker, where it settles).  The lanes of the chain share the units of a mail (coalesced loads and
  //      peer stores); the two usual source rows were requested before the cascade (mail_up / mail_dn), so sending them
  //      costs stores only.  This is the head of the longest dependence between ranks (my publish -> the peer's cascade
  //      -> its mail -> my rows), which is why it comes before the counts.
  if (SHARDED && p.mail_in && valid) {
    const int MU = LD + 1;
    const uint32_t tag = (uint32_t)(it + 1ull);
    // up: the walker of my top rung moves up to rung t_hi
    if (p.t_hi < T && sel_bit(sel_lo, sel_hi, p.t_hi)) {
      const int sr = p.t_hi - 1;
      int gd = p.rank;
      while (gd + 1 < p.world && p.t_hi >= p.temp_begin[gd + 1]) ++gd;
      const size_t sslot = (size_t)(sr - p.t_lo) * W + pos[sr];
      uint4* box = p.mail_dst[gd] + ((size_t)0 * W + chain) * MU;
      for (int e = lane; e < MU; e += CL) {
        const double v = mail_pre ? mail_up : e < LD ? p.coords_src[p.rank][sslot * LD + e] : p.logp_src[p.rank][sslot];
        st_volatile_u4(box + e, make_uint4((uint32_t)__double2loint(v), tag, (uint32_t)__double2hiint(v), tag));
      }
    }
    // down: the walker carried across my lower boundary, if it started on one of my rungs
    if (p.t_lo >= 1 && sel_bit(sel_lo, sel_hi, p.t_lo)) {
      int o = p.t_lo;
      while (o + 1 < T && sel_bit(sel_lo, sel_hi, o + 1)) ++o;      // rung the carried walker started on
      if (o < p.t_hi) {
        int d = p.t_lo - 1;
        while (d >= 1 && sel_bit(sel_lo, sel_hi, d)) --d;           // rung it settles on
        int gd = 0;
        while (gd + 1 < p.world && d >= p.temp_begin[gd + 1]) ++gd;
        const size_t sslot = (size_t)(o - p.t_lo) * W + pos[o];
        uint4* box = p.mail_dst[gd] + ((size_t)1 * W + chain) * MU;
        for (int e = lane; e < MU; e += CL) {
          const double v = (mail_pre && o == p.t_lo) ? mail_dn
                           : e < LD ? p.coords_src[p.rank][sslot * LD + e] : p.logp_src[p.rank][sslot];
          st_volatile_u4(box + e, make_uint4((uint32_t)__double2loint(v), tag, (uint32_t)__double2hiint(v), tag));
        }
      }
    }
  }
  EB_MARK(31);

  // ---- single-GPU passes with rows in registers: the source rows of the rungs that change are requested NOW, so that
  //      their L2 latency runs under the count publication below (two block barriers and the global reductions); the
  //      stores follow after it (all reads of a chain before any of its writes: the block barriers order them)
  constexpr int RRX = RR > 0 ? RR : 1;
  double rowv[RPL][RRX], lpv[RPL], llv[RPL];
  long long dsl[RPL];
#pragma unroll
  for (int m = 0; m < RPL; ++m) dsl[m] = -1;
  if (RR > 0 && !SHARDED && !EB_DBG_SKIP(1)) {
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      const int r = lane + m * CL;
      dsl[m] = -1;
      if (valid && r < T) {
        const int s = swap_source(sel_lo, sel_hi, r, T);
        if (s != r) {
          const size_t sslot = (size_t)s * W + pos[s];
          dsl[m] = (long long)r * W + pos[r];
          if (EB_DBG_SKIP(32)) {
#pragma unroll
            for (int e = 0; e < RRX; ++e) rowv[m][e] = 1.0;
            lpv[m] = 0.0; llv[m] = ll[s];
          } else
          if ((LD & 3) == 0) {
#pragma unroll
            for (int e = 0; e < RRX; e += 4)
              if (e < LD) ld256(c.coords + sslot * LD + e, rowv[m][e], rowv[m][e + 1], rowv[m][e + 2], rowv[m][e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < RRX; ++e)
              if (e < LD) rowv[m][e] = c.coords[sslot * LD + e];
          }
          if (!EB_DBG_SKIP(32)) lpv[m] = c.logp[sslot];
          llv[m] = ll[s];
        }
      }
    }
  }

  // ---- swap counts: swaps_accepted[r-1] counts accepted swaps at rung r (:542): ballot over the chains of the warp,
  //      shared-memory atomics over the block, global atomics over the grid
  {
    constexpr int CPW = 32 / CL;                     // chains per warp
    const int wl = tid & 31;
    for (int r0 = 0; r0 < T; r0 += CL) {             // uniform trip count: the loop body votes
      const int r = r0 + lane;
      const bool b = valid && r >= 1 && r < T && sel_bit(sel_lo, sel_hi, r);
      const unsigned v = __ballot_sync(0xffffffffu, b);
      if (wl < CL) {
        unsigned m = 0u;
#pragma unroll
        for (int q = 0; q < CPW; ++q) m |= 1u << (wl + q * CL);
        const int n = __popc(v & m);
        if (n) atomicAdd(&s_cnt[r - 1], n);
      }
    }
  }

  unsigned defer_ticket = 0xffffffffu;
  // The counts are published here, before the rows move: fire-and-forget reductions spread over swap_slots() addresses
  // (same-address atomics serialise in L2).
  __syncthreads();
  if (!SHARDED) {
    // Single-GPU pass: warp 0 publishes and draws ONE ticket per CTA with release / acquire semantics.  The CTA that draws
    // the last ticket has every count in sight: its warp 0 folds them and adapts the ladder at once (pt_swap_adapt_warp)
    // while everybody else — its own other warps included — moves rows.  No CTA polls, no block barrier in the tail.
    if (adapt_cta) {              // the extra CTA published iter_next at its start
      if (p.defer) {
        // deferred adaptation: the snapshot the next stretch kernel (or eb_adapt_flush) adapts from.  Taken here, after the
        // grid-dependency wait: the ladder and the clock may have been written by the kernel before this pass.
        for (int r = tid; r < T; r += blockDim.x) ctrl->pend_betas[r] = p.betas[r];
        if (tid == 0) {
          ctrl->pend_time = *reinterpret_cast<const volatile long long*>(&ctrl->time);
          ctrl->pend_adapt_on = p.adapt_on; ctrl->pend_adaptive = p.adaptive; ctrl->pend_stop = p.stop_adaptation;
          ctrl->pend_T = T; ctrl->pend_W = W; ctrl->pend_lag = p.lag; ctrl->pend_t0 = p.t0;
          ctrl->adapt_pending = it + 1ull;
        }
      }
      return;
    }
    if (!EB_DBG_SKIP(2) && tid < 32) {
      const int row = (int)(it & 1ull) * LAZY_SLOTS + (int)(blockIdx.x % LAZY_SLOTS);
      for (int r = tid; r < T - 1; r += 32)
        if (s_cnt[r]) atomicAdd(&ctrl->swaps_work[row][r], s_cnt[r]);
      if (!p.defer) {
        __syncwarp();
        unsigned prev = 0u;
        if (tid == 0) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&ctrl->ticket) : "memory");
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == (unsigned)(nreal - 1) && !EB_DBG_SKIP(4)) {
          EB_MARK_ANY(28);
          pt_swap_adapt_warp(p, T, W, it, s_betas, s_dts, s_cnt);
          EB_MARK_ANY(23);
        }
      } else if (tid == 0) {
        // deferred: nobody folds here, so nothing has to be released — a relaxed ticket only finds the CTA that ticks the
        // iteration counter once every CTA has read it (its value is looked at after the rows have been issued)
        asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(defer_ticket) : "l"(&ctrl->ticket) : "memory");
      }
    }
  } else {
    // Sharded pass: the rows of a chain wait for mail from the peers, so no chain warp can spare the time: an extra CTA
    // (the last block) spins on the arrival slots, folds the counts and adapts the ladder WHILE the others move rows.
    if (EB_DBG_SKIP(2)) {
      if (adapt_cta) return;
    } else if (!adapt_cta) {
      for (int r = tid; r < T - 1; r += blockDim.x)
        if (s_cnt[r]) atomicAdd(&ctrl->swaps_work[blockIdx.x % swap_slots(T)][r], s_cnt[r]);
      __syncthreads();
      if (tid == 0) {        // block barrier + one device-scope release by the signalling thread (cumulative)
        fence_acq_rel_gpu();
        atomicAdd(&ctrl->arrive[blockIdx.x % swap_slots(T)], 1u);
      }
    } else {
      if (EB_DBG_SKIP(4)) { if (tid < EB_SWAP_SLOTS) ctrl->arrive[tid] = 0u; return; }
      const long long time_now = *reinterpret_cast<const volatile long long*>(&ctrl->time);
      pt_swap_adapt(p, T, W, nreal, it, time_now, s_betas, s_dts, s_cnt);
      EB_MARK(23);
      return;
    }
  }

  EB_MARK(20);
  // ---- move the rows that changed rung (do_swaps_indexing, tempering.py:351-482): every lane gathers the source rows
  //      of its rungs, the lanes of the chain synchronise (all reads before any write), then write
  if (EB_DBG_SKIP(1)) return;
  if (!SHARDED) {
    if (RR > 0) {
      EB_MARK(24);
      if (!EB_DBG_SKIP(128)) __syncwarp();
      EB_MARK(25);
      if (EB_DBG_SKIP(64)) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < RPL; ++m)
          if (dsl[m] >= 0) {
#pragma unroll
            for (int e = 0; e < RRX; ++e) acc += rowv[m][e];
            acc += lpv[m] + llv[m];
          }
        if (acc == 1.2345e300) c.logl[0] = acc;
        EB_MARK(21);
        return;
      }
#pragma unroll
      for (int m = 0; m < RPL; ++m)
        if (dsl[m] >= 0) {
          const size_t dslot = (size_t)dsl[m];
          if ((LD & 3) == 0) {
#pragma unroll
            for (int e = 0; e < RRX; e += 4)
              if (e < LD) st256(c.coords + dslot * LD + e, rowv[m][e], rowv[m][e + 1], rowv[m][e + 2], rowv[m][e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < RRX; ++e)
              if (e < LD) c.coords[dslot * LD + e] = rowv[m][e];
          }
          c.logp[dslot] = lpv[m];
          c.logl[dslot] = llv[m];
        }
    } else {
      // long rows / leaf flags / long ladders: the moved rows rest in the global staging buffers between the gather
      // and the scatter (indexed by destination slot, so chains never collide).  The WARP walks over the chains of its
      // lane groups and copies every moved row cooperatively (warp_copy_*); the source rung of every rung comes out of
      // one downward pass over the accept bits.
      constexpr int CPW = 32 / CL;
      const int wl = tid & 31, warp = tid >> 5;
      const bool long_rows = LD * (int)sizeof(double) + (c.inds ? L : 0) >= 256;
      if (!long_rows) {
        // short rows on a long ladder: one lane per rung keeps 32 rows of the warp in flight at once
        for (int r = lane; valid && r < T; r += CL) {
          const int sr = swap_source(sel_lo, sel_hi, r, T);
          if (sr == r) continue;
          const size_t sslot = (size_t)sr * W + pos[sr], dslot = (size_t)r * W + pos[r];
          copy_row(p.scratch_coords + dslot * LD, c.coords + sslot * LD, LD);
          p.scratch_logp[dslot] = c.logp[sslot];
          if (c.inds) copy_bytes(p.scratch_inds + dslot * L, c.inds + sslot * L, L);
        }
        __syncwarp();
        for (int r = lane; valid && r < T; r += CL) {
          const int sr = swap_source(sel_lo, sel_hi, r, T);
          if (sr == r) continue;
          const size_t dslot = (size_t)r * W + pos[r];
          copy_row(c.coords + dslot * LD, p.scratch_coords + dslot * LD, LD);
          c.logp[dslot] = p.scratch_logp[dslot];
          c.logl[dslot] = ll[sr];
          if (c.inds) copy_bytes(c.inds + dslot * L, p.scratch_inds + dslot * L, L);
        }
      } else
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
        for (int q = 0; q < CPW; ++q) {
          const int gl = q * CL;                                   // first lane of the chain group
          const unsigned long long qlo = __shfl_sync(0xffffffffu, sel_lo, gl), qhi = __shfl_sync(0xffffffffu, sel_hi, gl);
          if (!__shfl_sync(0xffffffffu, (int)valid, gl)) continue; // uniform over the warp
          const int gq = warp * CPW + q;                           // the group's index within the CTA
          const int* qpos = reinterpret_cast<const int*>(smraw + lay.pos) + (size_t)gq * T;
          const double* qll = reinterpret_cast<const double*>(smraw + lay.ll) + (size_t)gq * T;
          int top = T - 1;                                         // rung the walker carried past rung r started on
          for (int r = T - 1; r >= 0; --r) {
            if (!(r + 1 < T && sel_bit(qlo, qhi, r + 1))) top = r;
            const int sr = (r >= 1 && sel_bit(qlo, qhi, r)) ? r - 1 : top;   // = swap_source(qlo, qhi, r, T)
            if (sr == r) continue;
            const size_t sslot = (size_t)sr * W + qpos[sr], dslot = (size_t)r * W + qpos[r];
            if (phase == 0) {
              warp_copy_doubles(p.scratch_coords + dslot * LD, c.coords + sslot * LD, LD, wl);
              if (wl == 0) p.scratch_logp[dslot] = c.logp[sslot];
              if (c.inds) warp_copy_bytes(p.scratch_inds + dslot * L, c.inds + sslot * L, L, wl);
            } else {
              warp_copy_doubles(c.coords + dslot * LD, p.scratch_coords + dslot * LD, LD, wl);
              if (wl == 0) {
                c.logp[dslot] = p.scratch_logp[dslot];
                c.logl[dslot] = qll[sr];
              }
              if (c.inds) warp_copy_bytes(c.inds + dslot * L, p.scratch_inds + dslot * L, L, wl);
            }
          }
        }
        __syncwarp();      // every gather of the warp's chains before any scatter (chains never leave their warp)
      }
    }
  } else if (valid) {
    // Sharded: this rank owns rungs [t_lo, t_hi).  Every owned slot is (re)written into the destination buffers
    // from the CURRENT buffers of whichever rank holds the source rung (NVLink peer loads); lanes work on different
    // rungs, so the loads of all owned rungs of the chain are in flight together.  Source and destination buffers
    // are distinct: no staging.
    const bool mail = p.mail_in != nullptr;
    for (int r = p.t_lo + lane; r < p.t_hi; r += CL) {
      const int s = swap_source(sel_lo, sel_hi, r, T);
      int gsrc = 0;
      while (gsrc + 1 < p.world && s >= p.temp_begin[gsrc + 1]) ++gsrc;
      const size_t dslot = (size_t)(r - p.t_lo) * W + pos[r];
      c.logl[dslot] = ll[s];
      if (mail && gsrc != p.rank) continue;             // arrives by mail (below)
      const size_t sslot = (size_t)(s - p.temp_begin[gsrc]) * W + pos[s];
      copy_row(c.coords + dslot * LD, p.coords_src[gsrc] + sslot * LD, LD);
      c.logp[dslot] = p.logp_src[gsrc][sslot];
      if (c.inds) copy_bytes(c.inds + dslot * L, p.inds_src[gsrc] + sslot * L, L);
    }
    if (mail) {
      // the mail of this chain, after the local copies were issued: [0] from below into rung t_lo, [1] from above into
      // the rung where the carried walker settles (if that is one of mine); every lane polls its units and writes them
      const int MU = LD + 1;
      const uint32_t tag = (uint32_t)(it + 1ull);
      int dest[2] = {-1, -1};
      if (p.t_lo >= 1 && sel_bit(sel_lo, sel_hi, p.t_lo)) dest[0] = p.t_lo;
      if (p.t_hi < T && sel_bit(sel_lo, sel_hi, p.t_hi)) {
        int d = p.t_hi - 1;
        while (d >= 1 && sel_bit(sel_lo, sel_hi, d)) --d;
        if (d >= p.t_lo) dest[1] = d;
      }
      bool ok = *reinterpret_cast<volatile unsigned int*>(&ctrl->error) == 0u;
      const long long t_start = clock64();
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        if (dest[dir] < 0) continue;
        const size_t dslot = (size_t)(dest[dir] - p.t_lo) * W + pos[dest[dir]];
        const uint4* box = p.mail_in + ((size_t)dir * W + chain) * MU;
        for (int e = lane; e < MU; e += CL) {
          uint4 v = ld_volatile_u4(box + e);
          while ((v.y != tag || v.w != tag) && ok) {
            if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) {
              atomicExch(&ctrl->error, EB_DEVERR_PEER_TIMEOUT);
              ok = false;
            }
            v = ld_volatile_u4(box + e);
          }
          const double x = __hiloint2double((int)v.z, (int)v.x);
          if (e < LD) c.coords[dslot * LD + e] = x;
          else c.logp[dslot] = x;
        }
      }
    }
  }

  EB_MARK(21);
  if (!SHARDED && p.defer && tid == 0 && defer_ticket == (unsigned)(nreal - 1)) {
    ctrl->ticket = 0u;
    ctrl->iter = it + 1ull;      // every CTA has read the counter: this was the last ticket
  }
}

// ================================================================================================
// K3w: the ladder in rung RANGES (wavefront form of the pass, eb_pt_swap_range / eb_pt_swap_finish)
// ================================================================================================
// eb_run_host streams the state in hottest temperature first.  Rung i is final as soon as the swap (i, i-1) has been
// decided (tempering.py:515 walks i = T-1 .. 1), so the ladder can be resolved a few rungs at a time while colder
// temperatures are still on their way in and finished rungs are already on their way out.  One thread owns one chain
// (the same chains, positions sigma_r(chain) and uniforms as pt_swap_kernel: results are bit-identical) and walks the
// rungs r_hi .. r_lo+1 in place: the walker being carried down stays in registers, the row of the rung below is
// requested one rung ahead, and a slot is written only when its content changes.
struct SwapRangeArgs {
  Common c;                     // the FULL state (all T temperatures), betas = full ladder
  int r_hi, r_lo;               // swaps (r, r-1) for r = r_hi .. r_lo+1
  int permute;
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  eb_ctrl* ctrl;
};

constexpr int RANGE_THREADS = 128;

template <int RR>
struct CarriedRow { double v[RR]; double ll, lp; };

template <int RR>
__device__ __forceinline__ void range_load(const Common& c, size_t slot, int LD, CarriedRow<RR>& w) {
  if ((LD & 3) == 0) {
#pragma unroll
    for (int e = 0; e < RR; e += 4)
      if (e < LD) ld256(c.coords + slot * LD + e, w.v[e], w.v[e + 1], w.v[e + 2], w.v[e + 3]);
  } else {
#pragma unroll
    for (int e = 0; e < RR; ++e)
      if (e < LD) w.v[e] = c.coords[slot * LD + e];
  }
  w.ll = c.logl[slot];
  w.lp = c.logp[slot];
}

template <int RR>
__device__ __forceinline__ void range_store(const Common& c, size_t slot, int LD, const CarriedRow<RR>& w) {
  if ((LD & 3) == 0) {
#pragma unroll
    for (int e = 0; e < RR; e += 4)
      if (e < LD) st256(c.coords + slot * LD + e, w.v[e], w.v[e + 1], w.v[e + 2], w.v[e + 3]);
  } else {
#pragma unroll
    for (int e = 0; e < RR; ++e)
      if (e < LD) c.coords[slot * LD + e] = w.v[e];
  }
  c.logl[slot] = w.ll;
  c.logp[slot] = w.lp;
}

template <int RR>
__global__ void __launch_bounds__(RANGE_THREADS) pt_swap_range_kernel(const __grid_constant__ SwapRangeArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const Common& c = p.c;
  const int W = c.W, LD = c.LD, nr = p.r_hi - p.r_lo + 1, tid = threadIdx.x;
  double* s_dts = reinterpret_cast<double*>(smraw);                                  // [nr] betas[r-1] - betas[r]
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smraw + sizeof(double) * nr);       // [nr][FEISTEL_ROUNDS]
  int* s_cnt = reinterpret_cast<int*>(smraw + (sizeof(double) + sizeof(uint32_t) * FEISTEL_ROUNDS) * nr);   // [nr]
  unsigned long long it = p.iter;
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
  const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
  for (int i = tid; i < nr; i += blockDim.x) {
    const int r = p.r_lo + i;
    s_cnt[i] = 0;
    s_dts[i] = r >= 1 ? c.betas[r - 1] - c.betas[r] : 0.0;                           // tempering.py:518-522
    if (p.permute) Feistel::make_keys(key, TAG_SWAP_KEY, (uint32_t)r, s_keys + FEISTEL_ROUNDS * i);
  }
  __syncthreads();
  const int chain = blockIdx.x * blockDim.x + tid;
  const bool valid = chain < W;
  auto position = [&](int r) -> size_t {
    int pz = chain;
    if (p.permute) {
      Feistel sig;
      sig.init_from(s_keys + FEISTEL_ROUNDS * (r - p.r_lo), (uint32_t)W);
      pz = (int)sig((uint32_t)chain);
    }
    return (size_t)r * W + pz;
  };
  CarriedRow<RR> carry, low, nxt;
  size_t slot_r = 0, slot_low = 0, slot_nxt = 0;
  bool dirty = false;            // the carried walker differs from what slot_r holds
  if (valid) {
    slot_r = position(p.r_hi);
    range_load<RR>(c, slot_r, LD, carry);
    slot_nxt = position(p.r_hi - 1);
    range_load<RR>(c, slot_nxt, LD, nxt);
  }
  for (int r = p.r_hi; r > p.r_lo; --r) {                                             // uniform trip count: the body votes
    bool sel = false;
    if (valid) {
      low = nxt;
      slot_low = slot_nxt;
      if (r - 1 > p.r_lo) {                                                           // the rung after this one, requested now
        slot_nxt = position(r - 2);
        range_load<RR>(c, slot_nxt, LD, nxt);
      }
      // the uniforms of pt_swap_kernel: one Philox block serves rungs r and r + 8 of a chain
      const uint4 q = stream(key, TAG_SWAP_U, (uint32_t)chain, (uint32_t)((r & 7) | ((r >> 4) << 3)));
      const double u = ((r >> 3) & 1) ? u01_52(q.z, q.w) : u01_52(q.x, q.y);
      sel = s_dts[r - p.r_lo] * (carry.ll - low.ll) > log(u);                         // tempering.py:535-541
      if (sel) {
        range_store<RR>(c, slot_r, LD, low);       // the walker of rung r-1 moves up; the carried one goes on down
        dirty = true;
      } else {
        if (dirty) range_store<RR>(c, slot_r, LD, carry);   // the carried walker settles on rung r
        carry = low;
        dirty = false;
      }
      slot_r = slot_low;
    }
    const int n = __popc(__ballot_sync(0xffffffffu, sel));
    if ((tid & 31) == 0 && n) atomicAdd(&s_cnt[r - p.r_lo], n);
  }
  if (valid && dirty) range_store<RR>(c, slot_r, LD, carry);
  __syncthreads();
  for (int i = tid + 1; i < nr; i += blockDim.x)                                       // swaps_accepted[r-1] = rung r (:542)
    if (s_cnt[i]) atomicAdd(&p.ctrl->swaps_work[blockIdx.x % swap_slots(c.T)][p.r_lo + i - 1], s_cnt[i]);
}

// the tail of a pass resolved by range kernels: fold the counts, adapt the ladder, tick the iteration counter
__global__ void __launch_bounds__(SWAP_THREADS) pt_swap_finish_kernel(const __grid_constant__ SwapArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int T = p.T, tid = threadIdx.x;
  double* s_betas = reinterpret_cast<double*>(smraw);
  double* s_dts = s_betas + T;
  int* s_cnt = reinterpret_cast<int*>(s_dts + T);
  unsigned long long it = p.iter;
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
  for (int r = tid; r < T; r += blockDim.x) s_betas[r] = p.betas[r];
  const long long time_now = *reinterpret_cast<const volatile long long*>(&p.ctrl->time);
  __syncthreads();
  pt_swap_adapt(p, T, p.c.W, 0, it, time_now, s_betas, s_dts, s_cnt);
  if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(&p.ctrl->iter_next) = it + 1ull;
}

// eb_adapt_flush: a deferred adaptation applied by one CTA (whoever needs the ladder before the next stretch kernel)
__global__ void __launch_bounds__(128) adapt_flush_kernel(eb_ctrl* ctrl, double* betas) {
  __shared__ LazyShared sh;
  const unsigned long long pend = ld_volatile_u64(&ctrl->adapt_pending);
  if (pend == 0ull || pend == ld_volatile_u64(&ctrl->adapt_applied)) return;   // nothing deferred, or applied already
  lazy_adapt_apply(ctrl, pend, betas, true, true, sh);
  // a stretch kernel looks at adapt_pending only (its CTAs must all take the same decision while CTA (0,0) does the
  // bookkeeping): after a flush there is nothing left for it
  __syncthreads();
  if (threadIdx.x == 0) ctrl->adapt_pending = 0ull;
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

namespace eb {

static int fill_swap_common(SwapArgs& args, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl) {
  if (!rng || !ctrl) return fail(EB_ERR_INVALID, "rng/ctrl is NULL");
  args.philox = rng->mode == EB_RNG_PHILOX; args.permute = rng->permute;
  args.next_pos = rng->next_pos; args.u_at = rng->u_at;
  args.scratch_coords = rng->row_scratch; args.scratch_logp = rng->logp_scratch; args.scratch_inds = rng->inds_scratch;
  args.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); args.seed_hi = (uint32_t)(rng->seed >> 32);
  args.iter_dev = (const unsigned long long*)rng->iter_dev; args.iter = rng->iter;
  args.ctrl = ctrl;
  args.dbg_skip = getenv("EB_SWAP_SKIP") ? atoi(getenv("EB_SWAP_SKIP")) : 0;
  args.adapt_on = adapt != nullptr;
  args.adaptive = adapt ? adapt->adaptive : 0;
  args.stop_adaptation = adapt ? adapt->stop_adaptation : -1;
  args.lag = adapt ? adapt->adaptation_lag : 10000.0;
  args.t0 = adapt ? adapt->adaptation_time : 100.0;
  if (!args.philox && rng->mode != EB_RNG_REPLAY) return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  args.defer = rng->defer_adapt ? 1 : 0;
  if (args.defer && !args.philox) return fail(EB_ERR_INVALID, "defer_adapt is a philox-mode option");
  return EB_OK;
}

template <bool PHILOX, bool SHARDED, int CL, int RR, int RPL>
static int launch_swap_kernel(SwapArgs& args, cudaStream_t s) {
  auto kernel = pt_swap_kernel<PHILOX, SHARDED, CL, RR, RPL>;
  const int W = args.c.W;
  const int cpb = SWAP_THREADS / CL;
  args.cpb = cpb;
  const size_t sb = swap_layout(args.T, cpb).total;
  int rc = set_smem(kernel, sb);
  if (rc) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((W + cpb - 1) / cpb) + 1u, 1, 1);   // + the adapt CTA
  cfg.blockDim = dim3(SWAP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = sb;
  cfg.stream = s;
  // programmatic dependent of the move kernel: positions and log(u) are computed while the move still runs
  cudaLaunchAttribute attr[1];
  static const int pdl_mask = getenv("EB_PDL_MASK") ? atoi(getenv("EB_PDL_MASK")) : 5;
  // bit 1: programmatic dependent of the move kernel, prologue before the grid-dependency wait; bit 3: programmatic
  // dependent that waits FIRST (its CTAs are resident when the move kernel retires, no launch gap, and they do not
  // compete with the move kernel for issue slots)
  args.pdl = (pdl_mask & 8) ? 2 : (pdl_mask & 2) ? 1 : 0;
  if (pdl_mask & 10) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  EB_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
  return check_launch("pt_swap");
}

template <bool PHILOX, bool SHARDED>
static int launch_swap_shape(SwapArgs& args, cudaStream_t s) {
  const int T = args.T, LD = args.c.LD;
  // sharded passes copy every owned row between distinct buffers (no staging): RR is irrelevant there
  const bool regs_ok = SHARDED || (!args.c.inds && LD <= 32);
  const int cl = T <= 8 ? 8 : T <= 16 ? 16 : 32;
  const int rpl = (T + cl - 1) / cl;
  int rr = 0;
  if (regs_ok && rpl == 1) rr = LD <= 8 ? 8 : 32;
  else if (regs_ok && rpl == 2 && LD <= 8) rr = 8;
  if (SHARDED) rr = 8;
  if (rr == 0 && !SHARDED && (!args.scratch_coords || !args.scratch_logp || (args.c.inds && !args.scratch_inds)))
    return fail(EB_ERR_INVALID,
                "swap pass: this shape (T=%d, row of %d doubles%s) moves rows through staging buffers: set "
                "eb_swap_rng.row_scratch / logp_scratch%s", T, LD, args.c.inds ? ", leaf flags" : "",
                args.c.inds ? " / inds_scratch" : "");
#define EB_SWAP_GO(CL_, RR_, RPL_) return launch_swap_kernel<PHILOX, SHARDED, CL_, RR_, RPL_>(args, s)
  if (cl == 8) {
    if (rr == 8) EB_SWAP_GO(8, 8, 1);
    if (rr == 32) EB_SWAP_GO(8, 32, 1);
    EB_SWAP_GO(8, 0, 1);
  }
  if (cl == 16) {
    if (rr == 8) EB_SWAP_GO(16, 8, 1);
    if (rr == 32) EB_SWAP_GO(16, 32, 1);
    EB_SWAP_GO(16, 0, 1);
  }
  if (rr == 8 && rpl == 1) EB_SWAP_GO(32, 8, 1);
  if (rr == 32 && rpl == 1) EB_SWAP_GO(32, 32, 1);
  if (rr == 8 && rpl == 2) EB_SWAP_GO(32, 8, 2);
  EB_SWAP_GO(32, 0, 1);
#undef EB_SWAP_GO
}

template <bool SHARDED>
static int launch_swap(SwapArgs& args, cudaStream_t s) {
  if (args.T > 128) return fail(EB_ERR_UNSUPPORTED, "swap pass supports ladders of up to 128 temperatures (got %d)", args.T);
  if (args.philox) return launch_swap_shape<true, SHARDED>(args, s);
  return launch_swap_shape<false, SHARDED>(args, s);
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_pt_swap(const eb_state* st, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl, void* stream) {
  SwapArgs args;
  memset(&args, 0, sizeof(args));
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_swap_common(args, rng, adapt, ctrl);
  if (rc) return rc;
  if (!st->betas) return fail(EB_ERR_INVALID, "swap pass needs betas");
  cudaStream_t s = (cudaStream_t)stream;
  const int T = args.c.T, W = args.c.W;
  if (T < 2) return eb_advance_iter(ctrl, stream);  // range(ntemps-1, 0, -1) is empty
  args.T = T; args.logl_in = args.c.logl; args.betas = args.c.betas;
  if (!args.philox) {
    if (!rng->iperm || !rng->i1perm || !rng->u || !rng->next_pos || !rng->u_at)
      return fail(EB_ERR_INVALID, "replay mode needs iperm, i1perm, u and the next_pos/u_at scratch");
    const int n = T * W;
    pt_pairmap_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(rng->iperm, rng->i1perm, rng->u, T, W, rng->next_pos,
                                                               rng->u_at);
    rc = check_launch("pt_pairmap");
    if (rc) return rc;
  }
  return launch_swap<false>(args, s);
}

int eb_pt_swap_sharded(const eb_shard* sh, const eb_state* dst, const eb_swap_rng* rng, const eb_adapt* adapt,
                       eb_ctrl* ctrl, void* stream) {
  SwapArgs args;
  memset(&args, 0, sizeof(args));
  if (!sh) return fail(EB_ERR_INVALID, "shard description is NULL");
  int rc = fill_common(args.c, dst, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_swap_common(args, rng, adapt, ctrl);
  if (rc) return rc;
  if (!args.philox) return fail(EB_ERR_UNSUPPORTED, "temperature-sharded swaps run in philox mode only");
  if (args.defer) return fail(EB_ERR_UNSUPPORTED, "defer_adapt: single-GPU passes only");
  if (sh->world < 1 || sh->world > EB_MAX_RANKS || sh->rank < 0 || sh->rank >= sh->world)
    return fail(EB_ERR_INVALID, "bad rank/world %d/%d", sh->rank, sh->world);
  const int T = sh->ntemps_total;
  if (T < 1 || T > EB_MAX_TEMPS) return fail(EB_ERR_INVALID, "ntemps_total %d out of range", T);
  if (!sh->logl_all || !sh->betas_all) return fail(EB_ERR_INVALID, "logl_all/betas_all is NULL");
  if (sh->temp_begin[0] != 0 || sh->temp_begin[sh->world] != T)
    return fail(EB_ERR_INVALID, "temp_begin must run from 0 to ntemps_total");
  for (int g = 0; g < sh->world; ++g) {
    if (sh->temp_begin[g + 1] < sh->temp_begin[g]) return fail(EB_ERR_INVALID, "temp_begin must be non-decreasing");
    const bool owns = sh->temp_begin[g + 1] > sh->temp_begin[g];
    if (owns && (!sh->coords_src[g] || !sh->logp_src[g] || (dst->inds && !sh->inds_src[g])))
      return fail(EB_ERR_INVALID, "source pointers of rank %d are NULL", g);
    args.coords_src[g] = sh->coords_src[g]; args.logp_src[g] = sh->logp_src[g]; args.inds_src[g] = sh->inds_src[g];
  }
  for (int g = 0; g <= sh->world; ++g) args.temp_begin[g] = sh->temp_begin[g];
  args.sharded = 1; args.world = sh->world;
  args.t_lo = sh->temp_begin[sh->rank]; args.t_hi = sh->temp_begin[sh->rank + 1];
  if (dst->ntemps != args.t_hi - args.t_lo || dst->temp_offset != args.t_lo)
    return fail(EB_ERR_INVALID, "destination state must hold this rank's temperatures [%d, %d)", args.t_lo, args.t_hi);
  args.T = T; args.logl_in = sh->logl_all; args.betas = sh->betas_all;
  args.flags = (const unsigned long long*)sh->flags;
  args.rank = sh->rank;
  if (sh->pub_src) {
    if (sh->flags) return fail(EB_ERR_INVALID, "fused publish validates the data itself: pass flags = NULL");
    if (!sh->ll_in) return fail(EB_ERR_INVALID, "fused publish needs ll_in");
    for (int g = 0; g < sh->world; ++g) {
      if (!sh->pub_ll[g]) return fail(EB_ERR_INVALID, "LL buffer of rank %d is NULL", g);
      args.pub_dst[g] = (uint4*)sh->pub_ll[g];
    }
    args.pub_src = sh->pub_src;
    args.ll_in = (const uint4*)sh->ll_in;
  }
  if (sh->mail_in) {
    if (!sh->pub_src) return fail(EB_ERR_INVALID, "row mail rides on the fused publish (set pub_src)");
    if (dst->inds) return fail(EB_ERR_UNSUPPORTED, "row mail carries coords and logp only (no leaf flags)");
    for (int g = 0; g < sh->world; ++g) {
      if (!sh->mail_peer[g]) return fail(EB_ERR_INVALID, "mailbox of rank %d is NULL", g);
      args.mail_dst[g] = (uint4*)sh->mail_peer[g];
    }
    args.mail_in = (const uint4*)sh->mail_in;
  }
  if (T < 2) return eb_advance_iter(ctrl, stream);
  return launch_swap<true>(args, (cudaStream_t)stream);
}

int eb_adapt_flush(eb_ctrl* ctrl, double* betas, void* stream) {
  if (!ctrl || !betas) return fail(EB_ERR_INVALID, "ctrl/betas is NULL");
  adapt_flush_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(ctrl, betas);
  return check_launch("adapt_flush");
}

int eb_pt_swap_range(const eb_state* st, const eb_swap_rng* rng, eb_ctrl* ctrl, int32_t r_hi, int32_t r_lo, void* stream) {
  SwapRangeArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_common(a.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  if (!rng || !ctrl) return fail(EB_ERR_INVALID, "rng/ctrl is NULL");
  if (rng->mode != EB_RNG_PHILOX) return fail(EB_ERR_UNSUPPORTED, "rung ranges run in philox mode only");
  if (!st->betas) return fail(EB_ERR_INVALID, "swap pass needs betas");
  if (st->inds || a.c.LD > 32) return fail(EB_ERR_UNSUPPORTED, "rung ranges move rows of up to 32 doubles without leaf flags");
  if (r_lo < 0 || r_hi >= a.c.T || r_hi < r_lo) return fail(EB_ERR_INVALID, "bad rung range [%d, %d] of %d", r_lo, r_hi, a.c.T);
  if (r_hi == r_lo) return EB_OK;
  a.r_hi = r_hi; a.r_lo = r_lo; a.permute = rng->permute;
  a.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(rng->seed >> 32);
  a.iter_dev = (const unsigned long long*)rng->iter_dev; a.iter = rng->iter;
  a.ctrl = ctrl;
  const int nr = r_hi - r_lo + 1;
  const size_t sb = (sizeof(double) + sizeof(uint32_t) * FEISTEL_ROUNDS + sizeof(int)) * (size_t)nr;
  const unsigned grid = (unsigned)((a.c.W + RANGE_THREADS - 1) / RANGE_THREADS);
  if (a.c.LD <= 8) pt_swap_range_kernel<8><<<grid, RANGE_THREADS, sb, (cudaStream_t)stream>>>(a);
  else if (a.c.LD <= 20) pt_swap_range_kernel<20><<<grid, RANGE_THREADS, sb, (cudaStream_t)stream>>>(a);
  else pt_swap_range_kernel<32><<<grid, RANGE_THREADS, sb, (cudaStream_t)stream>>>(a);
  return check_launch("pt_swap_range");
}

int eb_pt_swap_finish(const eb_state* st, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl, void* stream) {
  SwapArgs args;
  memset(&args, 0, sizeof(args));
  int rc = fill_common(args.c, st, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_swap_common(args, rng, adapt, ctrl);
  if (rc) return rc;
  if (!st->betas) return fail(EB_ERR_INVALID, "swap pass needs betas");
  const int T = args.c.T;
  if (T > 128) return fail(EB_ERR_UNSUPPORTED, "swap pass supports ladders of up to 128 temperatures (got %d)", T);
  args.T = T; args.betas = args.c.betas;
  const size_t sb = (2 * sizeof(double) + sizeof(int)) * (size_t)T;
  pt_swap_finish_kernel<<<1, SWAP_THREADS, sb, (cudaStream_t)stream>>>(args);
  return check_launch("pt_swap_finish");
}

}  // extern "C"

EB_DEFINE_MARK_READER(eb_debug_marks_swap)
