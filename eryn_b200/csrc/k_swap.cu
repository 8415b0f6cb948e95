// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
#include "common.cuh"

namespace eb {

// ================================================================================================
// K3: chain-parallel parallel-tempering swap pass
// ================================================================================================
// The reference walks the ladder hot -> cold; at rung i it pairs (i, iperm[k]) with
// (i-1, i1perm[k]) (tempering.py:515-559).  Both permutations are bijections, so the slots touched
// by successive rungs form W disjoint chains  p_{T-1} -> p_{T-2} -> ... -> p_0  and every exchange
// stays inside one chain.  The accept test needs logl only (:538), so a chain is resolved in three
// phases by a group of 8 lanes, all staging in shared memory:
//   1. positions p_r of the chain on every rung, logl[r][p_r] and log(u) of every rung — independent
//      across rungs, lanes work on different rungs (philox mode: p_r = sigma_r(chain) with one keyed
//      bijection per rung, so the pairing of rung i is sigma_{i-1} o sigma_i^{-1}, a uniform random
//      bijection exactly like i1perm o iperm^{-1}; replay mode: lane 0 follows the host pair map);
//   2. the hot -> cold cascade on those T numbers (one lane, registers + shared memory) giving
//      src[r] = rung whose walker ends on rung r;
//   3. only the rows with src[r] != r move: cp.async global -> shared for all of them at once, one
//      warp-level sync, then shared -> global.  In place, no second state buffer.
// The last block to finish folds the per-rung swap counts and adapts the ladder (tempering.py:563-596).
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
  int philox, permute, cpb;                       // cpb = chains per block
  int spec;                                       // prefetch every row of the chain before the decisions (small, L2-resident states)
  const int32_t* next_pos; const double* u_at;  // replay pair map [T][W]
  uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  eb_ctrl* ctrl;
  int adapt_on, adaptive, stop_adaptation; double lag, t0;
};

constexpr int CHAIN_LANES = 8;

struct SwapLayout {  // byte offsets into dynamic shared memory
  size_t betas, dts, ll, lu, rows, keys, pos, src, cnt, inds, total;
};
__host__ __device__ inline SwapLayout swap_layout(int T, int LD, int L, int cpb, bool has_inds, bool stage_rows) {
  SwapLayout s;
  const int RS = (LD + 2) & ~1;                            // row stride in doubles (coords + logp), 16-byte rows
  size_t o = 0;
  s.rows = o; o += stage_rows ? sizeof(double) * T * RS * cpb : 0;   // [chain][rung][RS]
  s.betas = o; o += sizeof(double) * T;
  s.dts = o; o += sizeof(double) * T;
  s.ll = o; o += sizeof(double) * T * cpb;
  s.lu = o; o += sizeof(double) * T * cpb;
  s.keys = o; o += sizeof(uint32_t) * FEISTEL_ROUNDS * T;
  s.pos = o; o += sizeof(int) * T * cpb;
  s.src = o; o += sizeof(int) * T * cpb;
  s.cnt = o; o += sizeof(int) * T;
  s.inds = o; o += (has_inds && stage_rows) ? (size_t)T * L * cpb : 0;
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Lane l of a chain's 8-lane group owns rungs l, l+8, l+16, ... in every phase.
template <bool PHILOX, bool SHARDED>
__global__ void __launch_bounds__(128) pt_swap_kernel(const SwapArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const Common& c = p.c;
  const int T = p.T, W = c.W, LD = c.LD, L = c.L, cpb = p.cpb;
  const int RS = (LD + 2) & ~1;
  const SwapLayout lay = swap_layout(T, LD, L, cpb, c.inds != nullptr, !SHARDED);
  double* s_betas = reinterpret_cast<double*>(smraw + lay.betas);
  double* s_dts = reinterpret_cast<double*>(smraw + lay.dts);
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smraw + lay.keys);
  int* s_cnt = reinterpret_cast<int*>(smraw + lay.cnt);
  __shared__ bool s_last;

  const int tid = threadIdx.x;
  const int g = tid / CHAIN_LANES, lane = tid % CHAIN_LANES;
  const int chain = blockIdx.x * cpb + g;
  const bool valid = g < cpb && chain < W;
  const unsigned long long it = p.iter_dev ? *p.iter_dev : p.iter;
  const RngKey key = make_rng_key(p.seed_lo, p.seed_hi, it);

  EB_MARK(16);
  if (SHARDED && p.flags) {
    // every rank's logl rows of THIS iteration must have landed in logl_in (eb_publish_logl): bounded spin on
    // the local flag words, one thread per CTA
    __shared__ bool s_ok;
    if (tid == 0) {
      bool ok = *reinterpret_cast<volatile unsigned int*>(&p.ctrl->error) == 0u;
      const unsigned long long target = p.ctrl->iter + 1ull;
      unsigned long long t_start;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
      for (int gr = 0; ok && gr < p.world; ++gr) {
        const volatile unsigned long long* f = p.flags + gr;
        while (*f < target) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (now - t_start > EB_PEER_TIMEOUT_NS) { ok = false; break; }
        }
      }
      if (!ok) atomicExch(&p.ctrl->error, EB_DEVERR_PEER_TIMEOUT);
      __threadfence_system();
      s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
  }
  // ---- phase 0: ladder and per-rung bijection keys, once per block ------------------------------
  for (int r = tid; r < T; r += blockDim.x) {
    s_betas[r] = p.betas[r];
    s_cnt[r] = 0;
    if (PHILOX && p.permute) Feistel::make_keys(key, TAG_SWAP_KEY, (uint32_t)r, s_keys + FEISTEL_ROUNDS * r);
  }
  __syncthreads();

  const int gg = valid ? g : 0;
  double* ll = reinterpret_cast<double*>(smraw + lay.ll) + (size_t)gg * T;
  double* lu = reinterpret_cast<double*>(smraw + lay.lu) + (size_t)gg * T;
  double* rows = reinterpret_cast<double*>(smraw + lay.rows) + (size_t)gg * T * RS;
  int* pos = reinterpret_cast<int*>(smraw + lay.pos) + (size_t)gg * T;
  int* src = reinterpret_cast<int*>(smraw + lay.src) + (size_t)gg * T;
  uint8_t* sinds = smraw + lay.inds + (size_t)gg * T * L;

  EB_MARK(17);
  // ---- phase 1: chain positions, logl and log(u) of the owned rungs ----------------------------
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
  }
  if (valid) {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    for (int m = 0, r = lane; r < T; ++m, r += CHAIN_LANES) {
      int pz;
      if (PHILOX) {
        pz = chain;
        if (p.permute) {
          Feistel sig;
          sig.init_from(s_keys + FEISTEL_ROUNDS * r, (uint32_t)W);
          pz = (int)sig((uint32_t)chain);
        }
        pos[r] = pz;
      } else {
        pz = pos[r];
      }
      ll[r] = p.logl_in[(size_t)r * W + pz];
      if (!SHARDED && p.spec) {          // the rows can start moving before the decisions are known
        const size_t slot = (size_t)r * W + pz;
        const double* grow = c.coords + slot * LD;
        double* srow = rows + (size_t)r * RS;
        if ((LD & 1) == 0) {
          for (int e = 0; e < LD; e += 2) cp_async16(srow + e, grow + e);
        } else {
          for (int e = 0; e < LD; ++e) cp_async8(srow + e, grow + e);
        }
        cp_async8(srow + LD, c.logp + slot);
        if (c.inds)
          for (int l = 0; l < L; ++l) sinds[r * L + l] = c.inds[slot * L + l];
      }
      double u;
      if (PHILOX) {                      // one Philox block serves two owned rungs (r and r + 8)
        if ((m & 1) == 0) q = stream(key, TAG_SWAP_U, (uint32_t)chain, (uint32_t)(lane + CHAIN_LANES * (m >> 1)));
        u = (m & 1) ? u01_52(q.z, q.w) : u01_52(q.x, q.y);
      } else {
        u = (r >= 1) ? p.u_at[(size_t)r * W + pz] : 0.5;
      }
      lu[r] = log(u);                                                          // tempering.py:535
    }
  }
  __syncwarp();

  EB_MARK(18);
  // ---- phase 2: the cascade, hot -> cold (tempering.py:515-559 restricted to this chain) --------
  if (valid && lane == 0) {
    double carry = ll[T - 1];
    int origin = T - 1;
    for (int i = T - 1; i >= 1; --i) {
      const double dbeta = s_betas[i - 1] - s_betas[i];                        // :518-522
      const double lower = ll[i - 1];
      const bool sel = dbeta * (carry - lower) > lu[i];                        // :538, :541
      if (sel) {
        src[i] = i - 1;                  // the colder walker moves up, the carried one keeps falling
      } else {
        src[i] = origin;                 // the carried walker settles on rung i
        carry = lower;
        origin = i - 1;
      }
    }
    src[0] = origin;
  }
  __syncwarp();

  EB_MARK(19);
  // ---- swap counts: swaps_accepted[r-1] counts src[r] == r-1 (:542), summed over the 4 chains of the
  //      warp by shuffles, over the block in shared memory, over the grid by atomics.  The ticket is taken
  //      here, before the rows move, so that its round trip overlaps phase 3.
  for (int r0 = 0; r0 < T; r0 += CHAIN_LANES) {   // uniform trip count: the loop body shuffles
    const int r = r0 + lane;
    int sel = (valid && r >= 1 && r < T && src[r] == r - 1) ? 1 : 0;
    sel += __shfl_xor_sync(0xffffffffu, sel, 8);
    sel += __shfl_xor_sync(0xffffffffu, sel, 16);
    if (sel && (tid & 31) < CHAIN_LANES) atomicAdd(&s_cnt[r - 1], sel);
  }
  __syncthreads();
  eb_ctrl* ctrl = p.ctrl;
  for (int r = tid; r < T - 1; r += blockDim.x)
    if (s_cnt[r]) atomicAdd(&ctrl->swaps_work[r], s_cnt[r]);
  __threadfence();
  __syncthreads();
  unsigned int ticket = 0u;
  if (tid == 0) ticket = atomicAdd(&ctrl->ticket, 1u);

  EB_MARK(20);
  // ---- phase 3: move the rows that changed rung (do_swaps_indexing, tempering.py:351-482) ------
  if (!SHARDED) {
    // staging row index: by source rung when prefetched (phase 1), by destination rung otherwise
    if (!p.spec && valid) {
      for (int r = lane; r < T; r += CHAIN_LANES) {
        const int s = src[r];
        if (s == r) continue;
        const size_t sslot = (size_t)s * W + pos[s];
        const double* grow = c.coords + sslot * LD;
        double* srow = rows + (size_t)r * RS;
        if ((LD & 1) == 0) {
          for (int e = 0; e < LD; e += 2) cp_async16(srow + e, grow + e);
        } else {
          for (int e = 0; e < LD; ++e) cp_async8(srow + e, grow + e);
        }
        cp_async8(srow + LD, c.logp + sslot);
        if (c.inds)
          for (int l = 0; l < L; ++l) sinds[r * L + l] = c.inds[sslot * L + l];
      }
    }
    cp_async_wait_all();
    __syncwarp();                        // every lane has read its sources before any lane writes
    if (valid) {
      for (int r = lane; r < T; r += CHAIN_LANES) {
        const int s = src[r];
        if (s == r) continue;
        const int st = p.spec ? s : r;
        const size_t dslot = (size_t)r * W + pos[r];
        double* grow = c.coords + dslot * LD;
        const double* srow = rows + (size_t)st * RS;
        if ((LD & 1) == 0) {
          for (int e = 0; e < LD; e += 2) *reinterpret_cast<double2*>(grow + e) = *reinterpret_cast<const double2*>(srow + e);
        } else {
          for (int e = 0; e < LD; ++e) grow[e] = srow[e];
        }
        c.logp[dslot] = srow[LD];
        c.logl[dslot] = ll[s];
        if (c.inds)
          for (int l = 0; l < L; ++l) c.inds[dslot * L + l] = sinds[st * L + l];
      }
    }
  } else if (valid) {
    // Sharded: this rank owns rungs [t_lo, t_hi).  Every owned slot is (re)written into the destination
    // buffers from the CURRENT buffers of whichever rank holds the source rung (NVLink peer loads), the 8
    // lanes of the chain sharing each row.  No staging: source and destination buffers are distinct.
    for (int r = p.t_lo; r < p.t_hi; ++r) {
      const int s = src[r];
      int gsrc = 0;
      while (gsrc + 1 < p.world && s >= p.temp_begin[gsrc + 1]) ++gsrc;
      const size_t sslot = (size_t)(s - p.temp_begin[gsrc]) * W + pos[s];
      const size_t dslot = (size_t)(r - p.t_lo) * W + pos[r];
      const double* grow = p.coords_src[gsrc] + sslot * LD;
      double* drow = c.coords + dslot * LD;
      for (int e = lane; e < LD; e += CHAIN_LANES) drow[e] = grow[e];
      if (lane == (LD & (CHAIN_LANES - 1))) {
        c.logp[dslot] = p.logp_src[gsrc][sslot];
        c.logl[dslot] = ll[s];
      }
      if (c.inds)
        for (int l = lane; l < L; l += CHAIN_LANES) c.inds[dslot * L + l] = p.inds_src[gsrc][sslot * L + l];
    }
  }

  EB_MARK(21);
  // ---- the block that drew the last ticket folds the counts and adapts the ladder -----------------
  if (tid == 0) s_last = (ticket == gridDim.x - 1);
  __syncthreads();
  EB_MARK(22);
  if (!s_last) return;
  __threadfence();
  for (int r = tid; r < T - 1; r += blockDim.x) {
    const int v = atomicExch(&ctrl->swaps_work[r], 0);
    ctrl->swaps_accepted[r] = v;
    ctrl->swaps_total[r] += (unsigned long long)v;
    s_cnt[r] = v;
  }
  const long long time_now = ctrl->time;
  __syncthreads();
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
      if (tid == 0) {                                                          // np.cumsum: sequential adds
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
  if (tid == 0) {
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

namespace eb {

static int fill_swap_common(SwapArgs& args, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl) {
  if (!rng || !ctrl) return fail(EB_ERR_INVALID, "rng/ctrl is NULL");
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
  if (!args.philox && rng->mode != EB_RNG_REPLAY) return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  return EB_OK;
}

template <bool SHARDED>
static int launch_swap(SwapArgs& args, cudaStream_t s) {
  // chains per block: as many as fit shared memory, at most 16 (128 threads)
  const int T = args.T, W = args.c.W;
  const bool has_inds = args.c.inds != nullptr;
  int cpb = 16;
  while (cpb > 1 && swap_layout(T, args.c.LD, args.c.L, cpb, has_inds, !SHARDED).total > 96 * 1024) cpb >>= 1;
  const size_t sb = swap_layout(T, args.c.LD, args.c.L, cpb, has_inds, !SHARDED).total;
  if (sb > 200 * 1024)
    return fail(EB_ERR_UNSUPPORTED, "swap pass: one chain of %d rungs x %d doubles does not fit shared memory", T,
                args.c.LD);
  args.cpb = cpb;
  args.spec = (!SHARDED && (size_t)T * W * (args.c.LD + 2) * sizeof(double) <= (size_t)48 << 20) ? 1 : 0;
  const int threads = max(32, cpb * CHAIN_LANES);
  const int grid = (W + cpb - 1) / cpb;
  int rc;
  if (args.philox) {
    rc = set_smem(pt_swap_kernel<true, SHARDED>, sb);
    if (rc) return rc;
    pt_swap_kernel<true, SHARDED><<<grid, threads, sb, s>>>(args);
  } else {
    rc = set_smem(pt_swap_kernel<false, SHARDED>, sb);
    if (rc) return rc;
    pt_swap_kernel<false, SHARDED><<<grid, threads, sb, s>>>(args);
  }
  return check_launch("pt_swap");
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
  if (T < 2) return eb_advance_iter(ctrl, stream);
  return launch_swap<true>(args, (cudaStream_t)stream);
}

}  // extern "C"

EB_DEFINE_MARK_READER(eb_debug_marks_swap)
