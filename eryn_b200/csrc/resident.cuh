// Part of eryn_b200; included at the end of k_stretch.cu (same translation unit: StretchArgs, job_draw, eval_point).
//
// K12 resident_kernel — `niter` whole iterations (StretchMove + temperature swaps + ladder adaptation) in ONE launch with
// the walker state RESIDENT IN SHARED MEMORY (ensemble.py:965-1045 is the loop, red_blue.py:89-333 + stretch.py:74-231 the
// move, tempering.py:484-649 the pass).
//
// For ensembles that fit the SMs' shared memory (config 2: 16 x 4096 x 8-d = 5.2 MB of 33 MB) the per-launch kernels are
// bound by latency, not bandwidth: three launch boundaries and five or six dependent L2 / HBM round trips per iteration
// (DESIGN.md §5).  Here a thread-block CLUSTER owns one temperature: its CTAs hold the temperature's walkers as records
// {row, logl, logp} in shared memory, a half step gathers the moving walker and its partner through distributed shared
// memory and updates the record in place, and the two halves are separated by a cluster barrier instead of a kernel
// boundary.  Only the swap pass crosses temperatures, i.e. clusters, and goes through global memory (L2-resident):
//     publish   every CTA writes its records to a global record buffer (coalesced), then ARRIVES at grid barrier 1 and,
//               before waiting, computes what does not depend on the state: the positions sigma_r(chain) and log u of
//               the chains it resolves (same chains, bijections and uniforms as pt_swap_kernel);
//     resolve   chains are dealt round-robin over ALL CTAs; a group of CL lanes gathers the chain's logl from the record
//               buffer, runs the hot -> cold cascade, writes for every rung whose walker changes the SOURCE slot into a
//               global order table and adds the accepted swaps to global per-rung counters; grid barrier 2;
//     fetch     every CTA looks up the orders of its own slots and pulls the moved records into shared memory; every CTA
//               folds the counters and adapts the ladder redundantly (identical arithmetic: adapt_temps, sequential cumsum).
// Two grid barriers and three cluster barriers per iteration; the record buffers alternate with the iteration parity so
// that a CTA publishing iteration i+1 never overwrites records a slower CTA still fetches for iteration i.
// Every draw is the one the per-launch kernels make (same Philox counters, same Feistel bijections) and the arithmetic is
// the same code (job_draw, eval_point, log_posterior): the chain is bit-identical to eb_stretch_step + eb_pt_swap
// (tests/test_gpu_resident.py).
#pragma once
#include <cooperative_groups.h>

namespace eb {
namespace cg = cooperative_groups;

constexpr int RES_THREADS = 512;
constexpr int RES_MAX_T = 32;            // one rung per lane of a chain group
constexpr size_t RES_HEADER_BYTES = 2048;   // scratch header: grid-barrier word, error word, swap counters [2][EB_MAX_TEMPS]

// -DEB_RES_MARKS (tools/build_variant.sh): thread 0 of the first and the last CTA note %globaltimer at the phase
// boundaries of the last iteration in the scratch header (tools/res_probe.py prints them)
#ifdef EB_RES_MARKS
#define RES_MARK(k)                                                                                   \
  do {                                                                                                \
    if (tid == 0 && (cta == 0 || cta == G - 1) && i == p.niter - 1) {                                 \
      unsigned long long gt_;                                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                         \
      reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(p.g_bar) + 1152)[(cta == 0 ? 0 : 16) + (k)] = gt_; \
    }                                                                                                 \
  } while (0)
#else
#define RES_MARK(k) do { } while (0)
#endif

struct ResidentArgs {
  StretchArgs sa;        // c = the full state in global memory; a, Ns, randomize, seed, iter_dev / iter
  int niter;
  int CS;                // CTAs per temperature = cluster size
  int Wc;                // walker slots per CTA = ceil(W / CS)
  int RS;                // doubles per record: row (LD), logl, logp, padded to an even count
  int CL;                // lanes per chain in the resolve phase (8, 16 or 32; >= T)
  int cpc;               // chains per CTA = ceil(W / (T * CS))
  int permute;
  int adapt_on, adaptive, stop_adaptation;
  double lag, t0;
  uint32_t wseed_lo, wseed_hi;   // seed of the swap streams (eb_swap_rng.seed)
  eb_ctrl* ctrl;
  double* g_rec;         // [2][T * W * RS] published records
  int32_t* g_src;        // [T * W] source slot of the record that moves into a slot, -1 = stays
  int* g_counts;         // [2][EB_MAX_TEMPS] accepted swaps per rung
  unsigned* g_bar;       // grid-barrier counter, zeroed by the host before the launch
};

struct ResidentLayout {  // byte offsets into dynamic shared memory
  size_t params, rec, acc_cnt, acc_flag, betas, dts, adt, swcnt, cnt2, keys, ll, lu, pos, total;
};
__host__ __device__ inline ResidentLayout resident_layout(const Common& c, int Wc, int RS, int cpc) {
  ResidentLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return at; };
  l.params = take(smem_bytes(c));
  l.rec = take(sizeof(double) * (size_t)Wc * RS);
  l.acc_cnt = take(sizeof(uint32_t) * (size_t)Wc);
  l.acc_flag = take((size_t)Wc);
  l.betas = take(sizeof(double) * c.T);
  l.dts = take(sizeof(double) * c.T);
  l.adt = take(sizeof(double) * c.T);
  l.swcnt = take(sizeof(int) * c.T);
  l.cnt2 = take(sizeof(int) * c.T);
  l.keys = take(sizeof(uint32_t) * FEISTEL_ROUNDS * c.T);
  l.ll = take(sizeof(double) * (size_t)cpc * c.T);
  l.lu = take(sizeof(double) * (size_t)cpc * c.T);
  l.pos = take(sizeof(int) * (size_t)cpc * c.T);
  l.total = o;
  return l;
}

// ---- grid barrier: arrive (block barrier + one release by thread 0) and wait (bounded spin by thread 0 + block barrier)
__device__ __forceinline__ void res_gbar_arrive(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void res_gbar_wait(unsigned* bar, unsigned target, eb_ctrl* ctrl, int* s_dead) {
  if (threadIdx.x == 0 && !*s_dead) {
    const long long t_start = clock64();
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t_start > 4000000000ll) {   // ~2 s: a CTA of the grid is not running (the grid must be co-resident)
        atomicExch(&ctrl->error, (unsigned)EB_DEVERR_SWAP_TIMEOUT);
        *s_dead = 1;
        break;
      }
    }
  }
  __syncthreads();
}

template <int DMAX>
__device__ __forceinline__ void res_load_row(const double* r, int LD, double (&x)[DMAX]) {
  if ((LD & 1) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 2)
      if (j < LD) {
        const double2 v = *reinterpret_cast<const double2*>(r + j);
        x[j] = v.x; x[j + 1] = v.y;
      }
  } else {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (j < LD) x[j] = r[j];
  }
}
template <int DMAX>
__device__ __forceinline__ void res_store_row(double* r, int LD, const double (&x)[DMAX]) {
  if ((LD & 1) == 0) {
#pragma unroll
    for (int j = 0; j < DMAX; j += 2)
      if (j < LD) *reinterpret_cast<double2*>(r + j) = make_double2(x[j], x[j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
      if (j < LD) r[j] = x[j];
  }
}

// source rung of the walker that ends on rung r (bit i of `bits` = swap accepted at rung i); as in k_swap.cu
__device__ __forceinline__ int res_swap_source(unsigned long long bits, int r, int T) {
  if (r >= 1 && ((bits >> r) & 1ull)) return r - 1;
  int o = r;
  while (o + 1 < T && ((bits >> (o + 1)) & 1ull)) ++o;
  return o;
}

template <int DMAX, int LIKE, bool EXACT>
__global__ void __launch_bounds__(RES_THREADS, 1) resident_kernel(const __grid_constant__ ResidentArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cg::cluster_group cluster = cg::this_cluster();
  const StretchArgs& sa = p.sa;
  const Common& c = sa.c;
  const int T = c.T, W = c.W, LD = EXACT ? DMAX : c.LD, RS = p.RS, Wc = p.Wc, CS = p.CS;
  const int tid = threadIdx.x, NT = blockDim.x, G = gridDim.x, cta = blockIdx.x;
  const int t = cta / CS;                       // the cluster's temperature
  const int cr = (int)cluster.block_rank();     // this CTA's share of its walkers
  const ResidentLayout lay = resident_layout(c, Wc, RS, p.cpc);
  double* sm_params = reinterpret_cast<double*>(smraw + lay.params);
  double* rec = reinterpret_cast<double*>(smraw + lay.rec);
  uint32_t* s_acc_cnt = reinterpret_cast<uint32_t*>(smraw + lay.acc_cnt);
  uint8_t* s_acc_flag = smraw + lay.acc_flag;
  double* s_betas = reinterpret_cast<double*>(smraw + lay.betas);
  double* s_dts = reinterpret_cast<double*>(smraw + lay.dts);
  double* s_adt = reinterpret_cast<double*>(smraw + lay.adt);
  int* s_swcnt = reinterpret_cast<int*>(smraw + lay.swcnt);
  int* s_cnt2 = reinterpret_cast<int*>(smraw + lay.cnt2);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smraw + lay.keys);
  double* s_ll = reinterpret_cast<double*>(smraw + lay.ll);
  double* s_lu = reinterpret_cast<double*>(smraw + lay.lu);
  int* s_pos = reinterpret_cast<int*>(smraw + lay.pos);
  __shared__ int s_dead;
  __shared__ unsigned long long s_total[RES_MAX_T];   // CTA 0: accepted swaps over the launch
  __shared__ int s_last[RES_MAX_T];

  // ---- load: parameters, this CTA's slice of the state, the ladder, the counters
  stage_params(c, sm_params);
  const int w0 = cr * Wc;
  const int nw = W - w0 < Wc ? (W - w0 > 0 ? W - w0 : 0) : Wc;
  for (int j = tid; j < nw; j += NT) {
    const size_t slot = (size_t)t * W + w0 + j;
    double x[DMAX];
    load_row<DMAX>(c.coords + slot * LD, LD, x);
    double* r = rec + (size_t)j * RS;
    res_store_row<DMAX>(r, LD, x);
    r[LD] = c.logl[slot];
    r[LD + 1] = c.logp[slot];
    s_acc_cnt[j] = 0u;
    s_acc_flag[j] = 0;
  }
  for (int r = tid; r < T; r += NT) { s_betas[r] = c.betas[r]; s_total[r] = 0ull; s_last[r] = 0; }
  if (tid == 0) s_dead = 0;
  unsigned long long it0 = sa.iter;
  if (sa.iter_dev) it0 = ld_volatile_u64(sa.iter_dev);
  long long time_now = *reinterpret_cast<const volatile long long*>(&p.ctrl->time);
  unsigned bar_target = 0u;
  cluster.sync();

  auto rec_of = [&](int w) -> double* {         // record of walker slot w of this temperature (any CTA of the cluster)
    const int owner = w / Wc;
    return cluster.map_shared_rank(rec, owner) + (size_t)(w - owner * Wc) * RS;
  };

  for (int i = 0; i < p.niter; ++i) {
    const unsigned long long it = it0 + (unsigned long long)i;
    const int par = (int)(it & 1ull);
    RES_MARK(0);
    // ================= the move: both red/blue halves, a cluster barrier after each =================
    {
      const RngKey key = make_rng_key(sa.seed_lo, sa.seed_hi, it);
      Feistel sig;
      if (sa.randomize) sig.init(key, TAG_SPLIT_KEY, (uint32_t)(c.t0 + t), (uint32_t)W);
      const double beta = s_betas[t];
      for (int s = 0; s < 2; ++s) {
        const int KPC = (sa.Ns[s] + CS - 1) / CS;
        for (int j = tid; j < KPC; j += NT) {
          const int k = cr * KPC + j;
          WalkerJob<DMAX> job;
          job_draw<DMAX, true, EXACT>(sa, key, sig, t, k, s, job);
          if (!job.live) continue;
          double* own = rec_of(job.w);
          const double* partner = rec_of(job.wc);
          double cc[DMAX];
          res_load_row<DMAX>(own, LD, job.q);                                   // s  (red_blue.py:173-179)
          res_load_row<DMAX>(partner, LD, cc);                                  // c_temp (stretch.py:100)
          const double ll0 = own[LD], lp0 = own[LD + 1];
#pragma unroll
          for (int d = 0; d < DMAX; ++d)
            if (EXACT || d < LD) job.q[d] = cc[d] - (cc[d] - job.q[d]) * job.zz;   // stretch.py:143-145
          double lp, ll;
          eval_point<DMAX, LIKE, EXACT>(job.q, c, sm_params, true, lp, ll);     // red_blue.py:260,270
          const double logP = log_posterior(ll, lp, beta, true);                // red_blue.py:283
          const double prevP = log_posterior(ll0, lp0, beta, true);             // red_blue.py:285-290
          const bool keep = (job.factors + logP - prevP) > job.log_u;           // red_blue.py:292-294
          if (keep) {                                                           // move.py:472-703
            res_store_row<DMAX>(own, LD, job.q);
            own[LD] = ll;
            own[LD + 1] = isinf(lp) ? 0.0 : lp;
          }
          const int owner = job.w / Wc, lj = job.w - owner * Wc;
          uint32_t* ac = cluster.map_shared_rank(s_acc_cnt, owner);
          uint8_t* af = cluster.map_shared_rank(s_acc_flag, owner);
          if (keep) ac[lj] += 1u;
          af[lj] = keep ? 1 : 0;
        }
        RES_MARK(1 + 2 * s);
        cluster.sync();
        RES_MARK(2 + 2 * s);
      }
    }
    if (T < 2) continue;   // range(ntemps-1, 0, -1) is empty: no pass (time and counts untouched, as eb_pt_swap)

    // ================= the pass =================
    const RngKey wkey = make_rng_key(p.wseed_lo, p.wseed_hi, it);
    double* g_rec = p.g_rec + (size_t)par * T * W * RS;
    int* g_counts = p.g_counts + par * EB_MAX_TEMPS;
    for (int r = tid; r < T; r += NT) {
      s_swcnt[r] = 0;
      s_dts[r] = r >= 1 ? s_betas[r - 1] - s_betas[r] : 0.0;                    // tempering.py:518-522
      if (p.permute) Feistel::make_keys(wkey, TAG_SWAP_KEY, (uint32_t)r, s_keys + FEISTEL_ROUNDS * r);
    }
    // publish this CTA's records
    {
      const double2* src = reinterpret_cast<const double2*>(rec);
      double2* dst = reinterpret_cast<double2*>(g_rec + ((size_t)t * W + w0) * RS);
      const int n2 = nw * RS / 2;
      for (int e = tid; e < n2; e += NT) dst[e] = src[e];
    }
    RES_MARK(5);
    res_gbar_arrive(p.g_bar);
    bar_target += (unsigned)G;
    // state-independent prologue of the chains this CTA resolves: chain = cta + q * G
    const int CL = p.CL, ntask = p.cpc * CL;
    for (int task = tid; task < ntask; task += NT) {
      const int q = task / CL, lane = task - q * CL;
      const int chain = cta + q * G;
      if (chain < W && lane < T) {
        const int r = lane;
        int pz = chain;
        if (p.permute) {
          Feistel sg;
          sg.init_from(s_keys + FEISTEL_ROUNDS * r, (uint32_t)W);
          pz = (int)sg((uint32_t)chain);
        }
        s_pos[q * T + r] = pz;
        // the uniforms of pt_swap_kernel: one Philox block serves rungs r and r + 8 of a chain
        const uint4 u4 = stream(wkey, TAG_SWAP_U, (uint32_t)chain, (uint32_t)((r & 7) | ((r >> 4) << 3)));
        const double u = ((r >> 3) & 1) ? u01_52(u4.z, u4.w) : u01_52(u4.x, u4.y);
        s_lu[q * T + r] = log(u);                                               // tempering.py:535
      }
    }
    RES_MARK(6);
    res_gbar_wait(p.g_bar, bar_target, p.ctrl, &s_dead);
    RES_MARK(7);
    // the counters of the NEXT pass: last read after barrier 2 of the previous pass, next written after barrier 1 of
    // the next one
    if (cta == 0)
      for (int r = tid; r < T; r += NT) p.g_counts[(par ^ 1) * EB_MAX_TEMPS + r] = 0;
    // resolve
    for (int task0 = 0; task0 < ntask; task0 += NT) {      // uniform trip count: the body votes
      const int task = task0 + tid;
      const int q = task / CL, lane = task - q * CL;
      const int chain = cta + q * G;
      const bool valid = task < ntask && chain < W && lane < T;
      if (valid) s_ll[q * T + lane] = __ldcg(g_rec + ((size_t)lane * W + s_pos[q * T + lane]) * RS + LD);
      __syncwarp();
      unsigned long long bits = 0ull;
      if (valid) {
        const double* ll = s_ll + q * T;
        const double* lu = s_lu + q * T;
        double carry = ll[T - 1];
        for (int r = T - 1; r >= 1; --r) {                                      // tempering.py:515-559 on this chain
          const double lower = ll[r - 1];
          const bool sel = s_dts[r] * (carry - lower) > lu[r];                  // :538, :541
          if (sel) bits |= 1ull << r;
          else carry = lower;           // the carried walker settles on rung r, rung r-1's walker is carried on
        }
        const int src = res_swap_source(bits, lane, T);
        p.g_src[(size_t)lane * W + s_pos[q * T + lane]] = src != lane ? src * W + s_pos[q * T + src] : -1;
      }
      // swaps_accepted[r-1] counts the accepted swaps at rung r (:542): ballot over the chains of the warp
      const bool b = valid && lane >= 1 && ((bits >> lane) & 1ull);
      const unsigned v = __ballot_sync(0xffffffffu, b);
      const int wl = tid & 31;
      if (wl < CL && wl >= 1 && wl < T) {
        unsigned m = 0u;
        for (int g2 = 0; g2 < 32 / CL; ++g2) m |= 1u << (wl + g2 * CL);
        const int n = __popc(v & m);
        if (n) atomicAdd(&s_swcnt[wl - 1], n);
      }
    }
    __syncthreads();
    for (int r = tid; r < T - 1; r += NT)
      if (s_swcnt[r]) atomicAdd(&g_counts[r], s_swcnt[r]);
    RES_MARK(8);
    res_gbar_arrive(p.g_bar);
    bar_target += (unsigned)G;
    res_gbar_wait(p.g_bar, bar_target, p.ctrl, &s_dead);
    RES_MARK(9);
    // fetch the records that moved into this CTA's slots (do_swaps_indexing, tempering.py:351-482)
    for (int j = tid; j < nw; j += NT) {
      const int o = __ldcg(p.g_src + (size_t)t * W + w0 + j);
      if (o >= 0) {
        const double2* src = reinterpret_cast<const double2*>(g_rec + (size_t)o * RS);
        double2* dst = reinterpret_cast<double2*>(rec + (size_t)j * RS);
        for (int e = 0; e < RS / 2; ++e) dst[e] = __ldcg(src + e);
      }
    }
    RES_MARK(10);
    // counts and ladder adaptation (adapt_temps, tempering.py:563-596), redundantly in every CTA
    for (int r = tid; r < T - 1; r += NT) s_cnt2[r] = __ldcg(g_counts + r);
    __syncthreads();
    if (cta == 0)
      for (int r = tid; r < T - 1; r += NT) { s_last[r] = s_cnt2[r]; s_total[r] += (unsigned long long)s_cnt2[r]; }
    if (p.adapt_on && p.adaptive) {                                              // tempering.py:632-633
      if (p.stop_adaptation < 0 || time_now < (long long)p.stop_adaptation) {   // :590
        const double decay = p.lag / ((double)time_now + p.lag);                 // :571
        const double kappa = decay / p.t0;                                       // :572
        const double nwalk = (double)W;
        for (int j = tid; j + 2 < T; j += NT) {
          const double r0 = (double)s_cnt2[j] / nwalk, r1 = (double)s_cnt2[j + 1] / nwalk;   // :587
          const double dS = kappa * (r0 - r1);                                   // :575
          const double dT = 1.0 / s_betas[j + 1] - 1.0 / s_betas[j];             // :578
          s_adt[j] = dT * exp(dS);                                               // :579
        }
        __syncthreads();
        if (tid == 0) {                                                          // np.cumsum: sequential adds, in order
          double cum = 0.0;
          for (int j = 0; j + 2 < T; ++j) { cum = cum + s_adt[j]; s_adt[j] = cum; }
        }
        __syncthreads();
        const double inv_b0 = 1.0 / s_betas[0];
        double bnew_mine = 0.0;
        const bool mine = tid + 2 < T;
        if (mine) {
          const double bold = s_betas[tid + 1];
          const double bnew = 1.0 / (s_adt[tid] + inv_b0);                       // :580
          bnew_mine = bold + (bnew - bold);                                      // :583, :593
        }
        __syncthreads();
        if (mine) s_betas[tid + 1] = bnew_mine;
      }
      time_now += 1;                                                             // :596
    }
    RES_MARK(11);
    cluster.sync();     // fetched records and the new ladder are in place before the next half step reads them
    RES_MARK(12);
  }

  // ---- store: the state slice, the accept mask of the last iteration, the counters; CTA 0: ladder and control block
  for (int j = tid; j < nw; j += NT) {
    const size_t slot = (size_t)t * W + w0 + j;
    const double* r = rec + (size_t)j * RS;
    double x[DMAX];
    res_load_row<DMAX>(r, LD, x);
    store_row<DMAX>(c.coords + slot * LD, LD, x);
    c.logl[slot] = r[LD];
    c.logp[slot] = r[LD + 1];
    sa.accepted[slot] = s_acc_flag[j];
    if (sa.accepted_count) sa.accepted_count[slot] += s_acc_cnt[j];
  }
  if (cta == 0) {
    for (int r = tid; r < T; r += NT) c.betas[r] = s_betas[r];
    if (T >= 2) {
      for (int r = tid; r < T - 1; r += NT) {
        p.ctrl->swaps_accepted[r] = s_last[r];
        p.ctrl->swaps_total[r] += s_total[r];
      }
    }
    if (tid == 0) {
      p.ctrl->time = time_now;
      p.ctrl->iter = it0 + (unsigned long long)p.niter;
      *reinterpret_cast<volatile unsigned long long*>(&p.ctrl->iter_next) = it0 + (unsigned long long)p.niter;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct ResidentPlan { int CS, Wc, RS, CL, cpc, threads; size_t smem; };

template <int DMAX, int LIKE, bool EXACT>
static int resident_try(const ResidentArgs& base, ResidentPlan& plan, bool launch, cudaStream_t s) {
  auto kernel = resident_kernel<DMAX, LIKE, EXACT>;
  const Common& c = base.sa.c;
  int dev = 0, nsm = 0;
  size_t smem_max = 0;
  {
    int v = 0;
    EB_CUDA(cudaGetDevice(&dev));
    EB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    EB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    smem_max = (size_t)v;
  }
  const int RS = (c.LD + 2 + 1) & ~1;
  const int CL = c.T <= 8 ? 8 : c.T <= 16 ? 16 : 32;
  static const int cs_env = getenv("EB_RESIDENT_CS") ? atoi(getenv("EB_RESIDENT_CS")) : 0;
  for (int cs : {4, 2, 1}) {
    if (cs_env && cs != cs_env) continue;
    const int Wc = (c.W + cs - 1) / cs;
    if (cs > 1 && Wc < 64) continue;                       // small ensembles: fewer, fuller CTAs
    if (c.T * cs > nsm) continue;
    const int cpc = (c.W + c.T * cs - 1) / (c.T * cs);
    const size_t smem = resident_layout(c, Wc, RS, cpc).total;
    if (smem > smem_max) continue;
    const int kpc = (base.sa.Ns[0] + cs - 1) / cs;
    int threads = kpc >= RES_THREADS ? RES_THREADS : ((kpc + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    EB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(c.T * cs), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
    if (nclusters < c.T) continue;                          // every CTA takes part in the grid barriers: all co-resident
    plan = ResidentPlan{cs, Wc, RS, CL, cpc, threads, smem};
    if (launch) {
      ResidentArgs a = base;
      a.CS = cs; a.Wc = Wc; a.RS = RS; a.CL = CL; a.cpc = cpc;
      EB_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
    }
    return EB_OK;
  }
  return fail(EB_ERR_UNSUPPORTED, "resident kernel: the ensemble (T=%d W=%d row=%d) does not fit the SMs' shared memory",
              c.T, c.W, c.LD);
}

template <int LIKE>
int resident_like(const ResidentArgs& a, ResidentPlan& plan, bool launch, cudaStream_t s);
#define EB_RESIDENT_LIKE_DEF(K)                                                              \
  template <>                                                                                \
  int resident_like<K>(const ResidentArgs& a, ResidentPlan& plan, bool launch, cudaStream_t s) { \
    const int LD = a.sa.c.LD;                                                                \
    if (LD == 8) return resident_try<8, K, true>(a, plan, launch, s);                        \
    if (LD < 8) return resident_try<8, K, false>(a, plan, launch, s);                        \
    return resident_try<16, K, false>(a, plan, launch, s);                                   \
  }
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 0
EB_RESIDENT_LIKE_DEF(0)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 1
EB_RESIDENT_LIKE_DEF(1)
#endif
#if EB_ONLY_LIKE == -1 || EB_ONLY_LIKE == 2
EB_RESIDENT_LIKE_DEF(2)
#endif

}  // namespace eb
