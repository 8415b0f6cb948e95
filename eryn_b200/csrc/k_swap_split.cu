// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
//
// K3s: the swap pass of a temperature-sharded run with the CHAINS split over the ranks.
//
// Validated on 2 GPUs against the unsharded oracle (tests/test_mgpu.py: uneven partitions, a Stretch/Gaussian mix, W = 16384
// chains where a CTA takes several chain groups, 128 rungs, config 4 at full size) and on 8 GPUs by the parity check bench.py
// runs after its timed region; `ShardedRun(comm="auto")` selects it beyond 2 ranks (measurements: profiles/README.md).
//
// Why: the fused sharded pass (k_swap.cu) resolves the WHOLE ladder on every rank — every rank draws positions and
// log u for all T rungs of all W chains and receives the logl of every other rank, so its cost grows with the number
// of ranks although the move work per rank is constant (36 / 43 / 69 us per iteration at 2 / 4 / 8 GPUs).  Here rank h
// resolves only the chains c with c % world == h, in two one-way NVLink hops:
//   A. every rank, for its OWN rungs r and every chain c: position p = sigma_r(c), and logl[r][p] goes to the rank that
//      resolves c as a self-validating 16-byte unit (k_swap.cu:publish_ll format), slot [r][c / world];
//   B. the resolving warp of chain c draws log u of every rung, polls the T units of the chain, runs the cascade (walk
//      over the carried walkers, as in k_swap.cu) and sends the accept bits of the chain (two units) to EVERY rank; its
//      accepted swaps are counted into the rank's partial swap counts;
//   C. every rank, every chain: poll the accept bits, then exactly as in the fused pass — rows that change rank leave as
//      mail pushed by the rank that owns the source rung (coords, logp AND logl here: a rank no longer knows the logl
//      of foreign rungs), local rows are copied into the alternate buffers, mail is polled;
//   D. the adapt CTA folds the partial counts of its rank, exchanges them with all ranks (units again) and adapts the
//      ladder — identically on every rank — while the rows move.
// Per rank and iteration: T_rank x W positions, T x W / world log u draws and cascade steps, 16 B x T_rank x W sent and
// received — the single-GPU amounts, at any number of ranks.  Random streams and arithmetic are those of k_swap.cu, so
// the chain is the same chain bit for bit.
#include <cstdlib>

#include "common.cuh"

namespace eb {
namespace split {

constexpr int THREADS = 256;
constexpr int CPB = 8;            // chains per CTA: one warp each
constexpr int AGES = 8;           // tests per walker evaluated ahead of the cascade walk
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000ll;

struct Args {
  Common c;                       // destination (alternate) buffers of this rank: rungs [t_lo, t_hi)
  int T, world, rank, t_lo, t_hi, permute, Wr;   // Wr = chains a rank resolves at most = ceil(W / world)
  int gpc;                        // chain groups per CTA (shared-memory layout)
  int nl;                         // lanes per chain in the phases A / C1 / C2 (4..32): a warp takes 32 / nl chains at a time
  int temp_begin[EB_MAX_RANKS + 1];
  const double* coords_cur; const double* logl_cur; const double* logp_cur;   // this rank's CURRENT buffers
  double* betas;                  // [T] local copy of the full ladder, adapted identically on every rank
  uint4* llc_dst[EB_MAX_RANKS]; const uint4* llc_in;     // hop A: [T][Wr] units on the resolving rank
  uint4* bits_dst[EB_MAX_RANKS]; const uint4* bits_in;   // hop B: [W][2] units on every rank
  uint4* cnt_dst[EB_MAX_RANKS]; const uint4* cnt_in;     // partial swap counts: [world][T] units on every rank
  uint4* mail_dst[EB_MAX_RANKS]; const uint4* mail_in;   // rows: [2][W][LD + 2] units
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  eb_ctrl* ctrl;
  int adapt_on, adaptive, stop_adaptation; double lag, t0;
};

struct Layout {  // byte offsets into dynamic shared memory
  size_t betas, dts, ll, lu, keys, pos, sel, cnt, band, rej, total;
};
// gpc = chain groups per CTA (the grid is capped at what is resident; a CTA loops over its groups phase by phase)
__host__ __device__ inline Layout layout(int T, int nown, int gpc, int cpw) {
  Layout s;
  size_t o = 0;
  s.betas = o; o += sizeof(double) * T;
  s.dts = o; o += sizeof(double) * T;
  s.ll = o; o += sizeof(double) * T * CPB;
  s.lu = o; o += sizeof(double) * T * CPB;
  s.sel = o; o += sizeof(unsigned long long) * 2 * CPB * cpw * gpc;
  s.keys = o; o += sizeof(uint32_t) * FEISTEL_ROUNDS * nown;
  s.pos = o; o += sizeof(int) * nown * CPB * cpw * gpc;
  s.cnt = o; o += sizeof(int) * T;
  s.band = o; o += (size_t)T * CPB;
  s.rej = o; o += (size_t)T * CPB;
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4* p, const uint4 u) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
// a double / two 32-bit words as a self-validating unit {lo32, tag, hi32, tag} (each aligned 8-byte half is written atomically)
__device__ __forceinline__ uint4 unit_of(double v, uint32_t tag) {
  return make_uint4((uint32_t)__double2loint(v), tag, (uint32_t)__double2hiint(v), tag);
}
__device__ __forceinline__ uint4 unit_of(uint32_t lo, uint32_t hi, uint32_t tag) { return make_uint4(lo, tag, hi, tag); }
__device__ __forceinline__ double unit_double(const uint4 v) { return __hiloint2double((int)v.z, (int)v.x); }

// poll one unit until both tags match (bounded; `ok` false: an earlier timeout, do not spin again)
__device__ __forceinline__ uint4 poll_unit(const uint4* p, uint32_t tag, bool& ok, long long t_start, eb_ctrl* ctrl) {
  uint4 v = ld_volatile_u4(p);
  while ((v.y != tag || v.w != tag) && ok) {
    if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) {
      atomicExch(&ctrl->error, EB_DEVERR_PEER_TIMEOUT);
      ok = false;
    }
    v = ld_volatile_u4(p);
  }
  return v;
}

__device__ __forceinline__ bool sel_bit(unsigned long long lo, unsigned long long hi, int i) {
  return i < 64 ? ((lo >> i) & 1ull) != 0ull : ((hi >> (i - 64)) & 1ull) != 0ull;
}
// source rung of the walker that ends on rung r (k_swap.cu:swap_source)
__device__ __forceinline__ int swap_source(unsigned long long lo, unsigned long long hi, int r, int T) {
  if (r >= 1 && sel_bit(lo, hi, r)) return r - 1;
  int o = r;
  while (o + 1 < T && sel_bit(lo, hi, o + 1)) ++o;
  return o;
}
__device__ __forceinline__ int owner_of(const Args& p, int r) {
  int g = 0;
  while (g + 1 < p.world && r >= p.temp_begin[g + 1]) ++g;
  return g;
}
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

// The adapt CTA (D): local counts -> exchange -> adapt_temps (tempering.py:563-596) -> bookkeeping.
__device__ __forceinline__ void adapt_cta_work(const Args& p, int nreal, unsigned long long it, long long time_t0,
                                               double* s_betas, double* s_dts, int* s_cnt) {
  eb_ctrl* ctrl = p.ctrl;
  const int tid = threadIdx.x, T = p.T, W = p.c.W;
  const uint32_t tag = (uint32_t)(it + 1ull);
  __shared__ int s_ok;
  __shared__ long long s_time;    // TemperatureControl.time, read by thread 0 at kernel start
  if (tid == 0) { s_ok = 1; s_time = time_t0; }
  for (int r = tid; r < T; r += blockDim.x) s_cnt[r] = 0;
  __syncthreads();
  constexpr int NS = 8;           // arrival / count slots (k_swap.cu:swap_slots for long ladders)
  if (tid < NS) {
    const unsigned expected = (unsigned)(nreal / NS + (tid < nreal % NS ? 1 : 0));
    const volatile unsigned* a = &ctrl->arrive[tid];
    const long long t_start = clock64();
    while (*a < expected)
      if (clock64() - t_start > SPIN_TIMEOUT_CYCLES) { s_ok = 0; break; }
    fence_acq_rel_gpu();
    ctrl->arrive[tid] = 0u;
  }
  __syncthreads();
  if (!s_ok) {   // sticky error, control block left consistent (k_swap.cu:pt_swap_adapt)
    if (tid == 0) atomicExch(&ctrl->error, EB_DEVERR_SWAP_TIMEOUT);
    for (int e = tid; e < EB_SWAP_SLOTS * (EB_MAX_TEMPS); e += blockDim.x) ctrl->swaps_work[e / EB_MAX_TEMPS][e % EB_MAX_TEMPS] = 0;
    if (tid < EB_SWAP_SLOTS) ctrl->arrive[tid] = 0u;
    for (int r = tid; r < T - 1; r += blockDim.x) ctrl->swaps_accepted[r] = 0;
    if (tid == 0) ctrl->iter = it + 1ull;
    return;
  }
  for (int e = tid; e < NS * (T - 1); e += blockDim.x) {
    const int v = *reinterpret_cast<volatile int*>(&ctrl->swaps_work[e / (T - 1)][e % (T - 1)]);
    if (v) atomicAdd(&s_cnt[e % (T - 1)], v);
  }
  __syncthreads();
  // this rank's partial counts go to every rank; the totals are the sums over the ranks in rank order (integers)
  for (int r = tid; r < T - 1; r += blockDim.x) {
    const uint4 u = unit_of((uint32_t)s_cnt[r], 0u, tag);
    for (int g = 0; g < p.world; ++g) st_volatile_u4(p.cnt_dst[g] + (size_t)p.rank * T + r, u);
  }
  bool ok = *reinterpret_cast<volatile unsigned int*>(&ctrl->error) == 0u;
  const long long t_start = clock64();
  for (int r = tid; r < T - 1; r += blockDim.x) {
    int tot = 0;
    for (int g = 0; g < p.world; ++g) tot += (int)poll_unit(p.cnt_in + (size_t)g * T + r, tag, ok, t_start, ctrl).x;
    s_cnt[r] = tot;
  }
  __syncthreads();
  const long long time_now = s_time;
  if (p.adapt_on && p.adaptive && T > 1) {                                     // tempering.py:632-633
    if (p.stop_adaptation < 0 || time_now < (long long)p.stop_adaptation) {   // :590
      const double decay = p.lag / ((double)time_now + p.lag);                 // :571
      const double kappa = decay / p.t0;                                       // :572
      const double nw = (double)W;
      for (int j = tid; j + 2 < T; j += blockDim.x) {
        const double r0 = (double)s_cnt[j] / nw, r1 = (double)s_cnt[j + 1] / nw;   // :587
        const double dS = kappa * (r0 - r1);                                   // :575
        const double dT = 1.0 / s_betas[j + 1] - 1.0 / s_betas[j];             // :578
        s_dts[j] = dT * exp(dS);                                               // :579
      }
      __syncthreads();
      if (tid == 0) {                                                          // np.cumsum: sequential adds, in order
        double cum = 0.0;
        for (int j = 0; j + 2 < T; ++j) { cum = cum + s_dts[j]; s_dts[j] = cum; }
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
  for (int e = tid; e < NS * (T - 1); e += blockDim.x) ctrl->swaps_work[e / (T - 1)][e % (T - 1)] = 0;
  for (int r = tid; r < T - 1; r += blockDim.x) {
    const int v = s_cnt[r];
    ctrl->swaps_accepted[r] = v;
    ctrl->swaps_total[r] += (unsigned long long)v;
  }
  if (tid == 0) ctrl->iter = it + 1ull;
}

// One chain's share of phase A: positions on this rank's rungs, and the log-likelihood found there goes to the chain's
// resolver.  `pos` (shared memory, nown ints) is kept for phases C1 / C2.
__device__ __forceinline__ void phase_a(const Args& p, const uint32_t* s_keys, int chain, int sub, int nl, int nown,
                                        uint32_t tag, int* pos) {
  const int W = p.c.W;
  const int h = chain % p.world;
  const size_t cslot = (size_t)(chain / p.world);
  for (int k = sub; k < nown; k += nl) {
    int pz = chain;
    if (p.permute) {
      Feistel sig;
      sig.init_from(s_keys + FEISTEL_ROUNDS * k, (uint32_t)W);
      pz = (int)sig((uint32_t)chain);
    }
    pos[k] = pz;
    const double v = p.logl_cur[(size_t)k * W + pz];
    st_volatile_u4(p.llc_dst[h] + (size_t)(p.t_lo + k) * p.Wr + cslot, unit_of(v, tag));
  }
}

// Phase B for one chain this rank resolves: log u, the T units, the cascade (walk over the carried walkers, k_swap.cu),
// the accept bits to every rank, the accepted swaps into the CTA's partial counts.
__device__ __forceinline__ void phase_b(const Args& p, const RngKey& key, int chain, int lane, uint32_t tag, bool& ok,
                                        long long t_start, const double* s_dts, double* ll, double* lu,
                                        unsigned char* sband, unsigned char* s_rej, int* s_cnt) {
  const int T = p.T;
  eb_ctrl* ctrl = p.ctrl;
  for (int r = lane; r < T; r += 32) {
    // one Philox block serves rungs r and r + 8 of a chain (k_swap.cu): counter (chain, (r & 7) | ((r >> 4) << 3)),
    // word pair (r >> 3) & 1
    const bool second = ((r >> 3) & 1) != 0;
    const uint4 q = stream(key, TAG_SWAP_U, (uint32_t)chain, (uint32_t)((r & 7) | ((r >> 4) << 3)));
    const double u = second ? u01_52(q.z, q.w) : u01_52(q.x, q.y);
    lu[r] = log(u);                                                          // tempering.py:535
    s_rej[r] = 0;
  }
  const uint4* src = p.llc_in + (size_t)(chain / p.world);
  {   // the (up to four) units of a lane are requested together, then re-polled until their tags match
    uint4 v[4];
    bool pend[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) pend[m] = lane + 32 * m < T;
    for (;;) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (pend[m]) v[m] = ld_volatile_u4(src + (size_t)(lane + 32 * m) * p.Wr);
      bool any = false;
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (pend[m]) {
          if ((v[m].y == tag && v[m].w == tag) || !ok) {
            ll[lane + 32 * m] = unit_double(v[m]);
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
  __syncwarp();
  EB_MARK(18);
  // band of walker j: bit a = test at rung j-a with x = ll[j] (tempering.py:538, :541), all walkers at once
  for (int j = lane; j < T; j += 32) {
    if (j >= 1) {
      const double x = ll[j];
      unsigned b = 0u;
#pragma unroll
      for (int a = 0; a < AGES; ++a) {
        const int i = j - a;
        if (i >= 1) b |= (unsigned)(s_dts[i] * (x - ll[i - 1]) > lu[i]) << a;
      }
      sband[j] = (unsigned char)b;
    }
  }
  __syncwarp();
  // walk over the carried walkers (k_swap.cu): mark the rung where each settles; long runs extended by a vote
  int j = T - 1;
  while (j >= 1) {
    int run = __ffs((int)~(unsigned)sband[j]) - 1;
    if (run == AGES) {
      const double x = ll[j];
      int base = j - AGES;
      while (base >= 1) {
        const int i = base - lane;
        const bool f = i >= 1 && s_dts[i] * (x - ll[i - 1]) > lu[i];
        const int n = __ffs((int)~__ballot_sync(0xffffffffu, f)) - 1;
        if (n < 0) { run += 32; base -= 32; continue; }
        run += n;
        break;
      }
    }
    s_rej[j - run] = 1;
    j -= run + 1;
  }
  __syncwarp();
  unsigned long long sel_lo = 0ull, sel_hi = 0ull;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    if (m * 32 < T) {
      const int r = lane + m * 32;
      const bool acc = r >= 1 && r < T && s_rej[r] == 0;
      const unsigned v = __ballot_sync(0xffffffffu, acc);
      if (acc) atomicAdd(&s_cnt[r - 1], 1);          // swaps_accepted[r-1] counts accepted swaps at rung r (:542)
      if (m == 0) sel_lo |= (unsigned long long)v;
      if (m == 1) sel_lo |= (unsigned long long)v << 32;
      if (m == 2) sel_hi |= (unsigned long long)v;
      if (m == 3) sel_hi |= (unsigned long long)v << 32;
    }
  }
  if (lane < p.world) {
    st_volatile_u4(p.bits_dst[lane] + 2 * (size_t)chain, unit_of((uint32_t)sel_lo, (uint32_t)(sel_lo >> 32), tag));
    st_volatile_u4(p.bits_dst[lane] + 2 * (size_t)chain + 1, unit_of((uint32_t)sel_hi, (uint32_t)(sel_hi >> 32), tag));
  }
  __syncwarp();   // ll / lu / band of this warp are reused by its next chain
}

// The grid is sized to be RESIDENT (every CTA both feeds remote resolvers and waits for remote ones); a CTA owns the
// chain groups blockIdx.x, blockIdx.x + nctas, ... and runs every phase over all of its groups before the next phase
// starts, so no CTA ever waits for a CTA that has not been scheduled, on this rank or on a peer:
//   A  (never waits)            all groups: own-rung positions, logl units to the resolvers
//   B  (waits for A of peers)   the chains this rank resolves: cascade, accept bits to every rank, partial counts
//      counts published for the adapt CTA (it exchanges them and adapts the ladder under C1 / C2)
//   C1 (waits for B of peers)   all groups: accept bits, mail pushed, local rows copied into the alternate buffers
//   C2 (waits for C1 of peers)  all groups: mail polled and stored
// Groups are visited in the same order on every rank (same grid), so the waits of a phase are always served by an
// earlier or equal step of the peer.
__global__ void __launch_bounds__(THREADS) pt_swap_split_kernel(const __grid_constant__ Args p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const Common& c = p.c;
  const int T = p.T, W = c.W, LD = c.LD;
  const int nown = p.t_hi - p.t_lo;
  const int NL = p.nl, CPW = 32 / NL, GC = CPB * CPW;   // lanes per chain, chains per warp, chains per group
  const Layout lay = layout(T, nown, p.gpc, CPW);
  double* s_betas = reinterpret_cast<double*>(smraw + lay.betas);
  double* s_dts = reinterpret_cast<double*>(smraw + lay.dts);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smraw + lay.keys);
  int* s_cnt = reinterpret_cast<int*>(smraw + lay.cnt);

  const bool adapt_cta = blockIdx.x == gridDim.x - 1;   // one extra CTA without chains (D)
  const int nctas = (int)gridDim.x - 1;
  const int ngroups = (W + GC - 1) / GC;
  const int sub = (threadIdx.x & 31) % NL, cw = (threadIdx.x & 31) / NL;   // lane within its chain, chain within the warp
  const int tid = threadIdx.x, g = tid >> 5, lane = tid & 31;
  eb_ctrl* ctrl = p.ctrl;
  long long time_now = 0;
  if (adapt_cta && tid == 0) time_now = *reinterpret_cast<const volatile long long*>(&ctrl->time);
  unsigned long long it = p.iter;
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
  if (adapt_cta && tid == 0) {   // the next move kernel keys its draws by iter_next (eb_stretch_rng.pdl_chain)
    *reinterpret_cast<volatile unsigned long long*>(&ctrl->iter_next) = it + 1ull;
    fence_acq_rel_gpu();
  }
  const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);
  const uint32_t tag = (uint32_t)(it + 1ull);
  EB_MARK(16);

  for (int r = tid; r < T; r += blockDim.x) s_cnt[r] = 0;
  for (int r = tid; r < nown; r += blockDim.x)
    if (p.permute) Feistel::make_keys(key, TAG_SWAP_KEY, (uint32_t)(p.t_lo + r), s_keys + FEISTEL_ROUNDS * r);
  for (int r = tid; r < T; r += blockDim.x) {            // the ladder was adapted by the previous pass (completed)
    const double b = p.betas[r];
    s_betas[r] = b;
    s_dts[r] = r >= 1 ? p.betas[r - 1] - b : 0.0;                              // tempering.py:518-522
  }
  __syncthreads();

  double* ll = reinterpret_cast<double*>(smraw + lay.ll) + (size_t)g * T;
  double* lu = reinterpret_cast<double*>(smraw + lay.lu) + (size_t)g * T;
  unsigned char* sband = smraw + lay.band + (size_t)g * T;
  unsigned char* s_rej = smraw + lay.rej + (size_t)g * T;
  int* pos_all = reinterpret_cast<int*>(smraw + lay.pos);                            // [gpc][GC][nown]
  unsigned long long* sel_all = reinterpret_cast<unsigned long long*>(smraw + lay.sel);   // [gpc][GC][2]
  bool ok = *reinterpret_cast<volatile unsigned int*>(&ctrl->error) == 0u;
  const long long t_start = clock64();

  // ---- A ----
  if (!adapt_cta) {
    int gi = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += nctas, ++gi) {
      const int ci = g * CPW + cw, chain = grp * GC + ci;
      if (chain < W) phase_a(p, s_keys, chain, sub, NL, nown, tag, pos_all + ((size_t)gi * GC + ci) * nown);
    }
  }
  __syncwarp();               // pos[] of the chain is read by all lanes below
  pdl_launch_dependents();    // the next move kernel may begin its draws
  EB_MARK(17);

  // ---- B ----
  if (!adapt_cta) {
    for (int grp = blockIdx.x; grp < ngroups; grp += nctas)
      for (int c2 = 0; c2 < CPW; ++c2) {               // the whole warp resolves one chain at a time
        const int chain = grp * GC + g * CPW + c2;
        if (chain < W && (chain % p.world) == p.rank)  // uniform over the warp
          phase_b(p, key, chain, lane, tag, ok, t_start, s_dts, ll, lu, sband, s_rej, s_cnt);
      }
  }
  EB_MARK(19);

  // ---- the rank's partial swap counts (resolved chains only) for the adapt CTA
  __syncthreads();
  if (!adapt_cta) {
    for (int r = tid; r < T - 1; r += blockDim.x)
      if (s_cnt[r]) atomicAdd(&ctrl->swaps_work[blockIdx.x % 8][r], s_cnt[r]);
    __syncthreads();
    if (tid == 0) {
      fence_acq_rel_gpu();
      atomicAdd(&ctrl->arrive[blockIdx.x % 8], 1u);
    }
  } else {
    adapt_cta_work(p, nctas, it, time_now, s_betas, s_dts, s_cnt);
    EB_MARK(23);
    return;
  }
  EB_MARK(20);

  // ---- C1: accept bits of every chain; rows that change rank leave as mail pushed by the rank that owns the source rung
  //      (coords, logp, logl); rows that stay on this rank are copied into the alternate buffers
  const int MU = LD + 2;
  {
    int gi = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += nctas, ++gi) {
      const int ci = g * CPW + cw, chain = grp * GC + ci;
      if (chain >= W) continue;
      const int* pos = pos_all + ((size_t)gi * GC + ci) * nown;
      const uint4 a = poll_unit(p.bits_in + 2 * (size_t)chain, tag, ok, t_start, ctrl);
      const uint4 b = poll_unit(p.bits_in + 2 * (size_t)chain + 1, tag, ok, t_start, ctrl);
      const unsigned long long sel_lo = ((unsigned long long)a.z << 32) | a.x;
      const unsigned long long sel_hi = ((unsigned long long)b.z << 32) | b.x;
      if (sub == 0) {
        sel_all[((size_t)gi * GC + ci) * 2] = sel_lo;
        sel_all[((size_t)gi * GC + ci) * 2 + 1] = sel_hi;
      }
      if (p.t_hi < T && sel_bit(sel_lo, sel_hi, p.t_hi)) {          // up: my top rung's walker moves to rung t_hi
        const int gd = owner_of(p, p.t_hi);
        const size_t sslot = (size_t)(nown - 1) * W + pos[nown - 1];
        uint4* box = p.mail_dst[gd] + ((size_t)0 * W + chain) * MU;
        for (int e = sub; e < MU; e += NL) {
          const double v = e < LD ? p.coords_cur[sslot * LD + e] : e == LD ? p.logp_cur[sslot] : p.logl_cur[sslot];
          st_volatile_u4(box + e, unit_of(v, tag));
        }
      }
      if (p.t_lo >= 1 && sel_bit(sel_lo, sel_hi, p.t_lo)) {         // down: the walker carried across my lower boundary
        int o = p.t_lo;
        while (o + 1 < T && sel_bit(sel_lo, sel_hi, o + 1)) ++o;    // rung it started on
        if (o < p.t_hi) {                                           // ... one of mine
          int d = p.t_lo - 1;
          while (d >= 1 && sel_bit(sel_lo, sel_hi, d)) --d;         // rung it settles on
          const int gd = owner_of(p, d);
          const size_t sslot = (size_t)(o - p.t_lo) * W + pos[o - p.t_lo];
          uint4* box = p.mail_dst[gd] + ((size_t)1 * W + chain) * MU;
          for (int e = sub; e < MU; e += NL) {
            const double v = e < LD ? p.coords_cur[sslot * LD + e] : e == LD ? p.logp_cur[sslot] : p.logl_cur[sslot];
            st_volatile_u4(box + e, unit_of(v, tag));
          }
        }
      }
      // every owned slot is rewritten into the alternate buffers; sources on this rank are copied here
      for (int r = p.t_lo + sub; r < p.t_hi; r += NL) {
        const int s = swap_source(sel_lo, sel_hi, r, T);
        if (s < p.t_lo || s >= p.t_hi) continue;
        const size_t sslot = (size_t)(s - p.t_lo) * W + pos[s - p.t_lo];
        const size_t dslot = (size_t)(r - p.t_lo) * W + pos[r - p.t_lo];
        copy_row(c.coords + dslot * LD, p.coords_cur + sslot * LD, LD);
        c.logp[dslot] = p.logp_cur[sslot];
        c.logl[dslot] = p.logl_cur[sslot];
      }
    }
  }
  __syncwarp();
  EB_MARK(31);

  // ---- C2: the others arrive by mail: [0] from below into rung t_lo, [1] from above into the rung where the carried
  //      walker settles
  {
    int gi = 0;
    for (int grp = blockIdx.x; grp < ngroups; grp += nctas, ++gi) {
      const int ci = g * CPW + cw, chain = grp * GC + ci;
      if (chain >= W) continue;
      const int* pos = pos_all + ((size_t)gi * GC + ci) * nown;
      const unsigned long long sel_lo = sel_all[((size_t)gi * GC + ci) * 2];
      const unsigned long long sel_hi = sel_all[((size_t)gi * GC + ci) * 2 + 1];
      int dest[2] = {-1, -1};
      if (p.t_lo >= 1 && sel_bit(sel_lo, sel_hi, p.t_lo)) dest[0] = p.t_lo;
      if (p.t_hi < T && sel_bit(sel_lo, sel_hi, p.t_hi)) {
        int d = p.t_hi - 1;
        while (d >= 1 && sel_bit(sel_lo, sel_hi, d)) --d;
        if (d >= p.t_lo) dest[1] = d;
      }
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        if (dest[dir] < 0) continue;
        const size_t dslot = (size_t)(dest[dir] - p.t_lo) * W + pos[dest[dir] - p.t_lo];
        const uint4* box = p.mail_in + ((size_t)dir * W + chain) * MU;
        for (int e = sub; e < MU; e += NL) {
          const double x = unit_double(poll_unit(box + e, tag, ok, t_start, ctrl));
          if (e < LD) c.coords[dslot * LD + e] = x;
          else if (e == LD) c.logp[dslot] = x;
          else c.logl[dslot] = x;
        }
      }
    }
  }
  EB_MARK(21);
}

}  // namespace split
}  // namespace eb

using namespace eb;

extern "C" {

int eb_pt_swap_split(const eb_split* sp, const eb_state* dst, const eb_swap_rng* rng, const eb_adapt* adapt,
                     eb_ctrl* ctrl, void* stream) {
  split::Args a;
  memset(&a, 0, sizeof(a));
  if (!sp || !rng || !ctrl) return fail(EB_ERR_INVALID, "split description / rng / ctrl is NULL");
  int rc = fill_common(a.c, dst, nullptr, nullptr, false);
  if (rc) return rc;
  if (rng->mode != EB_RNG_PHILOX) return fail(EB_ERR_UNSUPPORTED, "temperature-sharded swaps run in philox mode only");
  if (dst->inds) return fail(EB_ERR_UNSUPPORTED, "the chain-split pass moves coords, logp and logl only (no leaf flags)");
  if (sp->world < 1 || sp->world > EB_MAX_RANKS || sp->rank < 0 || sp->rank >= sp->world)
    return fail(EB_ERR_INVALID, "bad rank/world %d/%d", sp->rank, sp->world);
  const int T = sp->ntemps_total;
  if (T < 2 || T > EB_MAX_TEMPS) return fail(EB_ERR_INVALID, "ntemps_total %d out of range [2, %d]", T, EB_MAX_TEMPS);
  if (sp->temp_begin[0] != 0 || sp->temp_begin[sp->world] != T)
    return fail(EB_ERR_INVALID, "temp_begin must run from 0 to ntemps_total");
  for (int g = 0; g < sp->world; ++g) {
    if (sp->temp_begin[g + 1] <= sp->temp_begin[g]) return fail(EB_ERR_INVALID, "every rank must own a temperature");
    if (!sp->llc_peer[g] || !sp->bits_peer[g] || !sp->cnt_peer[g] || !sp->mail_peer[g])
      return fail(EB_ERR_INVALID, "exchange buffers of rank %d are NULL", g);
    a.llc_dst[g] = (uint4*)sp->llc_peer[g]; a.bits_dst[g] = (uint4*)sp->bits_peer[g];
    a.cnt_dst[g] = (uint4*)sp->cnt_peer[g]; a.mail_dst[g] = (uint4*)sp->mail_peer[g];
  }
  for (int g = 0; g <= sp->world; ++g) a.temp_begin[g] = sp->temp_begin[g];
  if (!sp->llc_in || !sp->bits_in || !sp->cnt_in || !sp->mail_in || !sp->coords_cur || !sp->logl_cur || !sp->logp_cur ||
      !sp->betas_all)
    return fail(EB_ERR_INVALID, "local buffers of the chain-split pass are NULL");
  a.llc_in = (const uint4*)sp->llc_in; a.bits_in = (const uint4*)sp->bits_in;
  a.cnt_in = (const uint4*)sp->cnt_in; a.mail_in = (const uint4*)sp->mail_in;
  a.coords_cur = sp->coords_cur; a.logl_cur = sp->logl_cur; a.logp_cur = sp->logp_cur; a.betas = sp->betas_all;
  a.T = T; a.world = sp->world; a.rank = sp->rank;
  a.t_lo = sp->temp_begin[sp->rank]; a.t_hi = sp->temp_begin[sp->rank + 1];
  if (dst->ntemps != a.t_hi - a.t_lo || dst->temp_offset != a.t_lo)
    return fail(EB_ERR_INVALID, "destination state must hold this rank's temperatures [%d, %d)", a.t_lo, a.t_hi);
  a.Wr = (a.c.W + sp->world - 1) / sp->world;
  a.permute = rng->permute;
  a.seed_lo = (uint32_t)(rng->seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(rng->seed >> 32);
  a.iter_dev = (const unsigned long long*)rng->iter_dev; a.iter = rng->iter;
  a.ctrl = ctrl;
  a.adapt_on = adapt != nullptr;
  a.adaptive = adapt ? adapt->adaptive : 0;
  a.stop_adaptation = adapt ? adapt->stop_adaptation : -1;
  a.lag = adapt ? adapt->adaptation_lag : 10000.0;
  a.t0 = adapt ? adapt->adaptation_time : 100.0;
  // Every CTA both feeds remote resolvers (A) and waits for remote ones (B, C): the grid must be resident at once, or
  // ranks would wait for each other's unscheduled CTAs.  A CTA takes gpc chain groups; gpc is the smallest count whose
  // grid fits (the shared-memory footprint grows with gpc, hence the loop).  Every rank computes the same grid.
  const int nown = a.t_hi - a.t_lo;
  a.nl = nown <= 4 ? 4 : nown <= 8 ? 8 : nown <= 16 ? 16 : 32;   // lanes per chain where a lane works on one of the rank's rungs
  const int cpw = 32 / a.nl;
  const int ngroups = (a.c.W + split::CPB * cpw - 1) / (split::CPB * cpw);
  int dev = 0, sms = 0;
  EB_CUDA(cudaGetDevice(&dev));
  EB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t sb = 0;
  unsigned grid = 0;
  for (int gpc = 1;; ++gpc) {
    sb = split::layout(T, nown, gpc, cpw).total;
    if (sb > 200 * 1024) return fail(EB_ERR_UNSUPPORTED, "chain-split pass: nwalkers %d does not fit a resident grid", a.c.W);
    rc = set_smem(split::pt_swap_split_kernel, sb);
    if (rc) return rc;
    int per_sm = 0;
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, split::pt_swap_split_kernel, split::THREADS, sb));
    const int nctas = (ngroups + gpc - 1) / gpc;
    // keep one SM's worth of slack: the move kernel that follows as a programmatic dependent needs room to start
    if ((long long)nctas + 1 <= (long long)per_sm * sms - per_sm || gpc >= 64) {
      a.gpc = gpc;
      grid = (unsigned)nctas + 1u;   // + the adapt CTA
      break;
    }
  }
  split::pt_swap_split_kernel<<<grid, split::THREADS, sb, (cudaStream_t)stream>>>(a);
  return check_launch("pt_swap_split");
}

}  // extern "C"

EB_DEFINE_MARK_READER(eb_debug_marks_split)
