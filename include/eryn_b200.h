/* eryn_b200 — C ABI of the B200-native walker-parallel sampling hot path of Eryn.
 *
 * The reference (mikekatz04/Eryn, /root/reference/src/eryn) is pure Python/NumPy: there is no
 * FFI today.  The boundary this library replaces is Eryn's Move protocol
 * (`Move.propose(model, state)`, ensemble.py:974) and `TemperatureControl.temper_comps`
 * (tempering.py:598).  Each entry point below names the reference routine it stands in for.
 * INTEGRATION.md shows the ctypes binding a maintainer of Eryn would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every `double*`/`uint8_t*`/`int32_t*` inside the structs
 *     is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); nothing here
 *     synchronises except the `*_host` entry points;
 *   - all floating data is IEEE float64, C-contiguous, laid out exactly like the reference's
 *     arrays: coords[ntemps][nwalkers][nleaves][ndim], logl/logp[ntemps][nwalkers],
 *     inds[ntemps][nwalkers][nleaves] (uint8, 0/1), betas[ntemps];
 *   - return value 0 = EB_OK, otherwise an `eb_status`; `eb_last_error()` gives the text.
 */
#ifndef ERYN_B200_H
#define ERYN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EB_API __attribute__((visibility("default")))
#else
#define EB_API
#endif

#define EB_ABI_VERSION 8
#define EB_MAX_TEMPS 128
#define EB_MAX_ROW 32 /* nleaves*ndim supported by the fused kernels */
#define EB_MAX_RANKS 16 /* GPUs of one temperature-sharded run */
#define EB_SWAP_SLOTS 32

typedef enum {
  EB_OK = 0,
  EB_ERR_INVALID = 1,     /* bad argument (ValueError in the reference) */
  EB_ERR_UNSUPPORTED = 2, /* shape / option outside what the fused kernels cover */
  EB_ERR_CUDA = 3,        /* a CUDA runtime call failed */
  EB_ERR_NODEVICE = 4     /* no CUDA device: there is NO CPU fallback */
} eb_status;

typedef enum { EB_RNG_REPLAY = 0, EB_RNG_PHILOX = 1 } eb_rng_mode;

#define EB_DEVERR_PEER_TIMEOUT 1u /* eb_ctrl.error: a peer's flag did not arrive within EB_PEER_TIMEOUT_NS */
#define EB_DEVERR_SWAP_TIMEOUT 2u /* eb_ctrl.error: the swap pass never saw the counts of all its CTAs */
#define EB_PEER_TIMEOUT_NS 2000000000ull /* nominal; the kernels count 4e9 SM cycles */

/* device-side log-likelihood functors (SURVEY.md §8d synthetic targets) */
typedef enum {
  EB_LIKE_GAUSSIAN = 0,   /* params: mu[D], P[D*D]           logL = -1/2 (x-mu)^T P (x-mu)          */
  EB_LIKE_ROSENBROCK = 1, /* params: none                    logL = -sum 100(x_{i+1}-x_i^2)^2+(1-x_i)^2 */
  EB_LIKE_GMIX = 2        /* params: logc[K], hinv[K], mu[K*D]  logL = logsumexp_k(logc_k - hinv_k |x-mu_k|^2) */
} eb_like_kind;

/* Walker state of one branch — state.py:387 (State) / :330 (Branch). */
typedef struct {
  int32_t ntemps, nwalkers, nleaves, ndim;
  int32_t temp_offset; /* index of local temperature 0 in the full ladder (temperature-sharded runs; else 0):
                          the random streams are keyed by the GLOBAL temperature, so a sharded run
                          reproduces the single-GPU chain */
  int32_t inds_stride; /* bytes per walker of `inds` when it carries more than the leaf flags (multi-branch states:
                          flags + friend table, eb_mb_state.aux); 0 = nleaves.  Only the swap pass looks at it. */
  double* coords;      /* [T][W][L][D] */
  double* logl;        /* [T][W]  State.log_like  */
  double* logp;        /* [T][W]  State.log_prior */
  uint8_t* inds;       /* [T][W][L] or NULL = all leaves active (moved by the swap pass) */
  double* betas;       /* [T] or NULL = no TemperatureControl (move.py:443 basic posterior) */
} eb_state;

/* Independent uniform priors — prior.py:12-91 + ProbDistContainer.logpdf prior.py:337-392. */
typedef struct {
  const double* lo;     /* [D] */
  const double* hi;     /* [D] */
  const double* logpdf; /* [D] log(1/(hi-lo)) as the host computed it (prior.py:40-41) */
  const double* period; /* [D] period of each parameter, 0 = not periodic; NULL = no periodic parameters.  The stretch
                           proposal then takes the distance through the boundary when that is shorter and wraps the
                           proposal into [0, period) (utils/periodic.py:49-151, stretch.py:136-153); the Gaussian
                           proposal wraps (gaussian.py:111-129).  Single-branch kernels only. */
} eb_prior;

typedef struct {
  int32_t kind;         /* eb_like_kind */
  int32_t ncomp;        /* K for EB_LIKE_GMIX, else 0 */
  int32_t nparams;      /* doubles in params */
  int32_t _pad;
  const double* params; /* device */
} eb_like;

/* Random inputs of one StretchMove step (both red/blue halves).
 * replay: the host NumPy draws in reference order.  `list[s]` holds the ascending walker ids of split s
 *         after the label shuffle (red_blue.py:121-124, :150-154) and, per split, the three draws of
 *         stretch.py:93 (rint), stretch.py:131 (u_z) and red_blue.py:294 (u_acc); everything is
 *         [T][Ns_s] with Ns_0 = ceil(W/2), Ns_1 = floor(W/2).  The device composes them exactly as the
 *         reference does: the k-th walker of split s is list[s][t][k] and its partner is
 *         list[1-s][t][rint[s][t][k]] (stretch.py:100).
 * philox: everything is generated in-kernel from (seed, *iter_dev, purpose tag, global temperature,
 *         position in the split). */
typedef struct {
  int32_t mode; /* eb_rng_mode */
  int32_t randomize_split; /* red_blue.py:123 (philox mode; replay encodes it in list) */
  int32_t pdl_chain;       /* philox: the previous kernel on the stream is eb_pt_swap[_sharded] of the same eb_ctrl and
                              iter_dev points at eb_ctrl.iter_next: the first half may then start its draws while that
                              pass is still moving rows (programmatic dependent launch) */
  int32_t _pad;
  const int32_t* list[2];  /* replay [T][Ns_s] */
  const int64_t* rint[2];  /* replay [T][Ns_s] */
  const double* u_z[2];    /* replay [T][Ns_s] */
  const double* u_acc[2];  /* replay [T][Ns_s]; may be NULL for eb_stretch_propose */
  uint64_t seed;           /* philox */
  const uint64_t* iter_dev; /* philox: device iteration counter (eb_ctrl.iter) or NULL */
  uint64_t iter;           /* philox: used when iter_dev == NULL */
  /* Gibbs split of this call (Move.gibbs_sampling_setup, moves/move.py:113-402; red_blue.py:127-325):
   * gibbs_mask  bit j set = parameter j of the (single) leaf moves, the others keep their values
   *             (cleanup_proposals_gibbs, move.py:302-307); 0 = no Gibbs split, every parameter moves
   * gibbs_ndim  number of selected parameters: factors = (gibbs_ndim - 1) log zz, computed the way adjust_factors does
   *             (stretch.py:55-72)
   * gibbs_index index of the split inside one propose() call: the philox streams are keyed by it; from the second
   *             split on `accepted` is the running OR over the splits and accepted_count is incremented by that OR,
   *             which is what the reference adds to Move.accepted per split (red_blue.py:296-309, :325) */
  uint32_t gibbs_mask;
  int32_t gibbs_ndim, gibbs_index, _pad2;
  void* lazy_ctrl;         /* philox, first launch of a step: the eb_ctrl of a pass that may have deferred its ladder
                              adaptation (eb_swap_rng.defer_adapt); the kernel folds the counts and adapts the ladder in
                              its prologue if that pass belongs to the iteration before this one.  NULL = nothing deferred */
} eb_stretch_rng;

/* Random inputs of one Gaussian Metropolis step (gaussian.py:68-195, mh.py:171). */
typedef struct {
  int32_t mode;
  int32_t cov_kind;       /* 0 scalar (gaussian.py:166), 1 full matrix via Cholesky factor (gaussian.py:192),
                             2 DistributionGenerate: every active leaf redrawn from the priors, factors =
                             log q(old) - log q(new) (distgen.py:34-104); replay: `delta` holds the new points */
  double scale;           /* sqrt(cov) for cov_kind 0 */
  const double* chol;     /* [D][D] lower Cholesky factor for cov_kind 1 (philox mode) */
  const double* delta;    /* replay [T][W][L][D] proposal increment drawn on the host */
  const double* u_acc;    /* replay [T][W] */
  uint64_t seed;
  const uint64_t* iter_dev;
  uint64_t iter;
  /* Gibbs split (mh.py:77-183): only the parameters in gibbs_mask change (0 = all); gibbs_index keys the philox streams
   * of the split.  `accepted` is the mask of this split (mh.py:171), accepted_count its sum over splits (:187). */
  uint32_t gibbs_mask;
  int32_t gibbs_index;
  /* philox mode, scalar proposals (gaussian.py:134-181): dim_mode 0 = "vector" (all dimensions), 1 = "random" (one
   * uniformly drawn dimension per walker; "sequential" is gibbs_mask with one bit, advanced by the host);
   * log_factor > 0: the scale is multiplied by exp(U(-log_factor, log_factor)), one draw per call (get_factor). */
  int32_t dim_mode, _pad3;
  double log_factor;
  void* lazy_ctrl;        /* philox: as eb_stretch_rng.lazy_ctrl (a pass that deferred its ladder adaptation) */
} eb_gauss_rng;

/* Random inputs of one swap pass (tempering.py:525-535). */
typedef struct {
  int32_t mode;
  int32_t permute;        /* tempering.py:525 */
  const int32_t* iperm;   /* replay [T][W], row i = permutation used at rung i (row 0 unused) */
  const int32_t* i1perm;  /* replay [T][W] */
  const double* u;        /* replay [T][W] uniforms (log taken on device, :535) */
  int32_t* next_pos;      /* replay scratch [T][W] int32 */
  double* u_at;           /* replay scratch [T][W] */
  double* row_scratch;    /* [T][W][L][D] staging of moved rows; needed when rows are longer than 32 doubles, the state
                             has leaf flags, T > 64, or T > 32 with rows longer than 8 doubles (else may be NULL) */
  double* logp_scratch;   /* [T][W], with row_scratch */
  uint8_t* inds_scratch;  /* [T][W][L], with row_scratch when the state has leaf flags */
  uint64_t seed;
  const uint64_t* iter_dev;
  uint64_t iter;
  int32_t defer_adapt;    /* single-GPU philox passes: do not fold the counts / adapt the ladder in this kernel; leave them
                             for the next stretch kernel (eb_stretch_rng.lazy_ctrl) or eb_adapt_flush.  The pass then ends
                             with its row moves.  Whoever reads betas / swaps_accepted / time before the
                             next stretch step must call eb_adapt_flush first. */
  int32_t _pad_defer;
} eb_swap_rng;

/* Device control block: iteration counter, ladder-adaptation clock and swap statistics
 * (TemperatureControl.time / .swaps_accepted, tempering.py:280,500,596). */
typedef struct {
  uint64_t iter;                        /* incremented by eb_pt_swap (or eb_advance_iter) */
  int64_t time;                         /* TemperatureControl.time */
  uint32_t ticket;                      /* last-block election */
  uint32_t error;                       /* 0, or EB_DEVERR_* set by a kernel (e.g. a peer never signalled) */
  int32_t swaps_work[EB_SWAP_SLOTS][EB_MAX_TEMPS]; /* scratch, zero between passes; CTAs spread their partial counts
                                                      over the slots (same-address atomics serialise in L2) */
  int32_t swaps_accepted[EB_MAX_TEMPS]; /* result of the last pass, entry i-1 = rung i */
  uint64_t swaps_total[EB_MAX_TEMPS];   /* running sum */
  uint32_t arrive[EB_SWAP_SLOTS];       /* swap pass: CTAs that have published their counts (spread like swaps_work) */
  uint64_t iter_next;                   /* = iter between iterations; a swap pass sets it to iter+1 BEFORE it releases
                                           its programmatic dependents, so the next move kernel can key its draws
                                           while the pass still runs (eb_stretch_rng.pdl_chain) */
  /* lazy ladder adaptation (eb_swap_rng.defer_adapt): the pass leaves its swap counts in swaps_work and a snapshot of
   * what adapt_temps needs; the next stretch kernel (eb_stretch_rng.lazy_ctrl) or eb_adapt_flush folds and adapts */
  uint64_t adapt_pending;               /* 0, or iter+1 of the deferred pass */
  uint64_t adapt_applied;               /* value of adapt_pending that has been applied */
  int64_t pend_time;                    /* TemperatureControl.time at the deferred pass */
  int32_t pend_adapt_on, pend_adaptive, pend_stop, pend_T, pend_W, _pad_lazy;
  double pend_lag, pend_t0;
  double pend_betas[EB_MAX_TEMPS];      /* the ladder at the deferred pass */
} eb_ctrl;

typedef struct {
  int32_t adaptive;         /* tempering.py:632 */
  int32_t stop_adaptation;  /* tempering.py:590 */
  double adaptation_lag;    /* tempering.py:571 */
  double adaptation_time;   /* tempering.py:572 */
} eb_adapt;

/* ---- library ---------------------------------------------------------------------- */
EB_API int eb_abi_version(void);
EB_API const char* eb_last_error(void);
EB_API int eb_device_count(void);
EB_API size_t eb_ctrl_size(void);
/* sizeof() of the ABI structs, for binding self-checks: 0 eb_state, 1 eb_prior, 2 eb_like,
 * 3 eb_stretch_rng, 4 eb_gauss_rng, 5 eb_swap_rng, 6 eb_ctrl, 7 eb_adapt, 8 eb_host_job, 9 eb_shard,
 * 10 eb_publish, 11 eb_mb_layout, 12 eb_mb_state, 13 eb_pulse_data, 14 eb_mb_friends, 15 eb_mb_group_rng,
 * 16 eb_mb_rj_rng, 17 eb_split, 18 eb_stage, 19 eb_mt_rng */
EB_API size_t eb_struct_size(int which);

/* ---- probability evaluation:  EnsembleSampler.compute_log_prior (ensemble.py:1127) and
 *      compute_log_like (ensemble.py:1219) on the whole state; fills st->logp, st->logl.
 *      Walkers with logp = -inf are not evaluated and get -1e300 (ensemble.py:1279-1282,1486). */
EB_API int eb_eval_state(const eb_state* st, const eb_prior* prior, const eb_like* like, void* stream);

/* ---- StretchMove.propose without the tempering tail: BOTH red/blue half steps of
 *      red_blue.py:148-323 + stretch.py:74-231 + Move.update (move.py:472-703): one launch per half, the second a
 *      programmatic dependent of the first (its draws overlap the first half's evaluation); small ensembles
 *      (ceil(W/2) <= 128) run both halves in one CTA per temperature.  Shapes that fill the GPU (exact row lengths 8 and
 *      20) take the lane-split kernel (csrc/stretch_lanes.cuh): same draws, same results up to the summation order of
 *      the likelihood.
 *      `accepted` [T][W] uint8 gets the accept flag of every walker; `accepted_count` (nullable,
 *      [T][W] uint32) is incremented (Move.accepted, move.py:404). */
EB_API int eb_stretch_step(const eb_state* st, const eb_prior* prior, const eb_like* like, double a,
                    const eb_stretch_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream);

/* ---- `niter` whole iterations — StretchMove.propose (red_blue.py:89-333, stretch.py:74-231) followed by
 *      TemperatureControl.temper_comps (tempering.py:598-649) — in ONE launch with the walker state resident in shared
 *      memory (the loop of ensemble.py:965-1045 for ensembles that fit the SMs: a thread-block cluster per temperature,
 *      half steps separated by cluster barriers, the pass through an L2-resident record buffer and two grid barriers;
 *      csrc/resident.cuh).  Same draws and arithmetic as eb_stretch_step + eb_pt_swap: the chain is bit-identical.
 *      Philox mode, tempered, single leaf without leaf flags / periodic parameters / Gibbs splits, 2 <= ntemps <= 32,
 *      rows of up to 16 doubles, and the state must fit shared memory: EB_ERR_UNSUPPORTED otherwise (niter = 0 only
 *      answers that question, nothing is launched) and the caller runs the per-launch kernels.  `scratch`: device memory
 *      of eb_resident_scratch_bytes(st) bytes, owned by the caller; both rng structs must name the same iteration counter
 *      (eb_ctrl.iter), which is advanced by niter; `accepted` receives the mask of the last iteration, `accepted_count`
 *      (nullable) is incremented per accepted proposal.  The launch needs the GPU to itself (its CTAs meet at grid
 *      barriers; a bounded wait sets eb_ctrl.error = EB_DEVERR_SWAP_TIMEOUT). */
EB_API size_t eb_resident_scratch_bytes(const eb_state* st);
EB_API int eb_resident_run(const eb_state* st, const eb_prior* prior, const eb_like* like, double a,
                    const eb_stretch_rng* srng, const eb_swap_rng* wrng, const eb_adapt* adapt, eb_ctrl* ctrl,
                    int32_t niter, uint8_t* accepted, uint32_t* accepted_count, void* scratch, size_t scratch_bytes,
                    void* stream);

/* ---- GaussianMove: one Metropolis step over all walkers = mh.py:56-193 + gaussian.py:68-195. */
EB_API int eb_gaussian_step(const eb_state* st, const eb_prior* prior, const eb_like* like,
                     const eb_gauss_rng* rng, uint8_t* accepted, uint32_t* accepted_count,
                     void* stream);

/* ---- TemperatureControl.temper_comps (tempering.py:598-649): swap ladder
 *      (temperature_swaps :484-561, do_swaps_indexing :351-482) + adapt_temps (:585-596).
 *      Resolved chain-parallel; results identical to the sequential ladder.
 *      Increments ctrl->iter.  `adapt` may be NULL (no adaptation, rj.py:381-382). */
EB_API int eb_pt_swap(const eb_state* st, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl,
               void* stream);

/* ---- apply a deferred ladder adaptation now (eb_swap_rng.defer_adapt; adapt_temps tempering.py:563-596 and the
 *      bookkeeping of temper_comps :598-649): no-op if nothing is pending.  `betas` = the ladder the pass ran on. */
EB_API int eb_adapt_flush(eb_ctrl* ctrl, double* betas, void* stream);

/* ---- the same pass a few rungs at a time (the wavefront form eb_run_host pipelines against its host copies):
 *      tempering.py:515 walks i = T-1 .. 1, and rung i is final once the swap (i, i-1) is decided, so
 *        eb_pt_swap_range(r_hi, r_lo)   decides and applies the swaps (r, r-1) for r = r_hi .. r_lo+1 in place
 *                                       (tempering.py:525-559 + do_swaps_indexing :351-482; philox mode, `st` is the
 *                                       FULL state, rows of up to 32 doubles, no leaf flags) and adds the accepted
 *                                       counts to ctrl->swaps_work;
 *        eb_pt_swap_finish              folds the counts into swaps_accepted / swaps_total, applies adapt_temps
 *                                       (:563-596) and increments ctrl->iter (and iter_next).
 *      Ranges must be issued hot -> cold, adjoining (r_lo of one = r_hi of the next), from T-1 down to 0, followed by
 *      one finish: the result is bit-identical to one eb_pt_swap with the same eb_swap_rng (same chains, positions and
 *      uniforms). */
EB_API int eb_pt_swap_range(const eb_state* st, const eb_swap_rng* rng, eb_ctrl* ctrl, int32_t r_hi, int32_t r_lo,
                     void* stream);
EB_API int eb_pt_swap_finish(const eb_state* st, const eb_swap_rng* rng, const eb_adapt* adapt, eb_ctrl* ctrl,
                      void* stream);

/* ---- the same pass when the ladder is sharded over GPUs by temperature (one process per GPU; rank g
 *      owns temperatures [temp_begin[g], temp_begin[g+1]) of all walkers).  The moves need
 *      no communication (red_blue.py:183-197 gathers along the walker axis only).  One iteration is
 *        move kernel (local)  ->  eb_publish_logl  ->  eb_pt_swap_sharded  ->  flip current/alternate
 *      eb_publish_logl is the all-gather of logl written as peer stores: every CTA copies its slice of the
 *      rank's rows into the `logl_all` buffer of EVERY rank over NVLink and then adds 1 to the rank's flag
 *      word on every rank (release at system scope).  eb_pt_swap_sharded first waits (bounded spin on LOCAL
 *      memory) until every flag word g reached (iter+1) * (CTAs of rank g's publish kernel), i.e. all of
 *      iteration iter's rows have landed — ctrl->iter must be 0 when the flag words are zeroed —, then resolves the whole ladder redundantly from `logl_all` — decisions depend on
 *      logl only (tempering.py:538), so all ranks agree bit for bit on swap counts and on the adapted
 *      ladder — and writes ITS rungs into `dst` (its alternate buffers), pulling each source row from the
 *      CURRENT buffers of the rank that owns it (`*_src[g]` are peer-mapped device pointers; entry `rank`
 *      is local).  The caller then flips current/alternate AND the logl_all parity buffer: a rank may run
 *      one publish ahead of a slow peer, never two (its next swap waits for that peer's next flag).
 *      Philox mode only. */
typedef struct {
  int32_t rank, world;
  int32_t ntemps_total;                   /* T of the full ladder */
  int32_t temp_begin[EB_MAX_RANKS + 1];
  const double* coords_src[EB_MAX_RANKS];  /* [T_g][W][L][D] current coords of rank g */
  const double* logp_src[EB_MAX_RANKS];    /* [T_g][W] */
  const uint8_t* inds_src[EB_MAX_RANKS];   /* [T_g][W][L] or NULL */
  const double* logl_all;                  /* [T][W] local copy of every rank's logl (this parity) */
  double* betas_all;                       /* [T] local copy of the full ladder; adapted in place */
  const uint64_t* flags;                   /* [EB_MAX_RANKS] local flag words (NULL: caller ordered the ranks itself,
                                              e.g. with an NCCL all-gather of logl) */
  /* Fused publish (ABI v4; pub_src NULL = the rows were published by eb_publish_logl or by NCCL).  With pub_src set the
   * swap kernel itself performs the all-gather: its first CTAs write this rank's logl rows into the LL buffer of EVERY
   * rank (own rank included) as self-validating 16-byte units {lo32, tag, hi32, tag}, tag = (uint32_t)(iter+1), with
   * coalesced NVLink peer stores, and the cascade polls the units it needs in `ll_in` until both tags match.  Each
   * aligned 8-byte half is written atomically and carries the tag, so the exchange needs no release fence, no flag word
   * and no wait for store acknowledgements (one one-way NVLink trip instead of three); `flags` must be NULL.  The LL
   * buffers alternate with the iteration parity like logl_all, and must be zero when ctrl->iter is 0. */
  const double* pub_src;                   /* [T_rank][W] this rank's current logl rows */
  void* pub_ll[EB_MAX_RANKS];              /* peer-mapped LL buffers (this parity) of every rank: [T][W] 16-byte units */
  const void* ll_in;                       /* local LL buffer (this parity) */
  /* Row mail (optional, needs pub_src; NULL = rows that change rank are pulled from *_src of the owning rank).  Every
   * rank resolves the whole chain, so the rank that owns the SOURCE rung of a walker that changes rank pushes the row
   * (coords, then logp; logl is known from the exchange) into the mailbox of the receiving rank as nleaves*ndim+1
   * self-validating 16-byte units, slot [direction][walker chain]: direction 0 = arrives from the colder neighbour into
   * the receiver's first rung, 1 = the carried walker arriving from hotter rungs.  One one-way NVLink trip, issued
   * before the swap counts are published, instead of the round trip of a pull.  Mailboxes alternate with the iteration
   * parity and must be zero when ctrl->iter is 0; [2][nwalkers][nleaves*ndim+1] units each. */
  void* mail_peer[EB_MAX_RANKS];           /* peer-mapped mailboxes (this parity) of every rank */
  const void* mail_in;                     /* local mailbox (this parity) */
} eb_shard;
EB_API int eb_pt_swap_sharded(const eb_shard* sh, const eb_state* dst, const eb_swap_rng* rng,
                       const eb_adapt* adapt, eb_ctrl* ctrl, void* stream);

/* ---- EXPERIMENTAL (one single-rank GPU run so far): the sharded pass with the CHAINS split over the ranks.  Rank h
 *      resolves the chains c with c % world == h: (A) every rank sends logl[r][sigma_r(c)] of its own rungs r to the
 *      resolver of c (`llc`, [T][ceil(W/world)] 16-byte self-validating units as in eb_shard.pub_ll), (B) the resolver
 *      runs the cascade and sends the accept bits of the chain to every rank (`bits`, [W][2] units), (C) every rank moves
 *      its rows — rows that change rank as mail pushed by the source rank (`mail`, [2][W][nleaves*ndim + 2] units: row,
 *      logp, logl) — and (D) the partial swap counts are exchanged (`cnt`, [world][T] units) and the ladder adapted
 *      identically everywhere.  Two one-way NVLink hops, per-rank work and traffic independent of the number of ranks.
 *      All exchange buffers alternate with the iteration parity and must be zero when ctrl->iter is 0.  The grid
 *      (nwalkers/8 + 1 CTAs) must be resident at once (EB_ERR_UNSUPPORTED otherwise).  Philox mode only. */
typedef struct {
  int32_t rank, world;
  int32_t ntemps_total, _pad;
  int32_t temp_begin[EB_MAX_RANKS + 1];
  const double* coords_cur;                /* this rank's CURRENT buffers [T_rank][W][L][D], [T_rank][W], [T_rank][W] */
  const double* logl_cur;
  const double* logp_cur;
  double* betas_all;                       /* [T] local copy of the full ladder; adapted in place */
  void* llc_peer[EB_MAX_RANKS];  const void* llc_in;
  void* bits_peer[EB_MAX_RANKS]; const void* bits_in;
  void* cnt_peer[EB_MAX_RANKS];  const void* cnt_in;
  void* mail_peer[EB_MAX_RANKS]; const void* mail_in;
} eb_split;
EB_API int eb_pt_swap_split(const eb_split* sp, const eb_state* dst, const eb_swap_rng* rng,
                     const eb_adapt* adapt, eb_ctrl* ctrl, void* stream);

typedef struct {
  int32_t rank, world;
  int32_t ntemps_total, nwalkers;
  int32_t temp_begin[EB_MAX_RANKS + 1];
  const double* logl_local;               /* [T_rank][W] */
  double* logl_all_peer[EB_MAX_RANKS];    /* peer-mapped logl_all (this parity) of every rank; entry `rank` local */
  uint64_t* flags_peer[EB_MAX_RANKS];     /* peer-mapped flag arrays of every rank */
} eb_publish;
EB_API int eb_publish_logl(const eb_publish* pub, eb_ctrl* ctrl, void* stream);

/* ---- MTDistGenMove(generate_dist = priors, num_try, independent=True): multiple-try Metropolis, one launch over all
 *      walkers = moves/multipletry.py:238-514 + moves/mtdistgen.py:8-133 inside MHMove.propose (mh.py:56-193).
 *      replay: `tries` [T][W][num_try][D] are the host draws of mtdistgen.py:58 (one global rand(n, num_try) per
 *      parameter), `u_sel` [T][W] the global uniform that picks a try (multipletry.py:51), `u_acc` [T][W] the private
 *      Metropolis uniform (mh.py:171).  philox: everything from (seed, *iter_dev, walker, try). */
typedef struct {
  int32_t mode, num_try;
  const double* tries; const double* u_sel; const double* u_acc;
  uint64_t seed; const uint64_t* iter_dev; uint64_t iter;
} eb_mt_rng;
EB_API int eb_mt_distgen_step(const eb_state* st, const eb_prior* prior, const eb_like* like, const eb_mt_rng* rng,
                       uint8_t* accepted, uint32_t* accepted_count, void* stream);

/* ---- Backend.save_step staging (backends/backend.py:1014-1091; call site ensemble.py:1013-1028).
 *      ONE kernel gathers every array of a stored step into one contiguous device staging slot (a snapshot: the
 *      sampler runs on while a side stream copies the slot to pinned host memory): segment i = nbytes[i] bytes from
 *      src[i] (8-byte aligned device pointers), packed back to back, each padded to a multiple of 8 bytes.  Segment 0
 *      may be the coordinates [rows][mask_nleaves][mask_ndim] with `mask_inds` = the leaf flags [rows][mask_nleaves]:
 *      coordinates of inactive leaves are then stored as `fill` (NaN in the reference, backend.py:1053-1059). */
#define EB_STAGE_MAX_SEGMENTS 16
typedef struct {
  int32_t nseg, mask_nleaves, mask_ndim, _pad;
  const void* src[EB_STAGE_MAX_SEGMENTS];
  uint64_t nbytes[EB_STAGE_MAX_SEGMENTS];
  const uint8_t* mask_inds;   /* nullable */
  double fill;
  void* dst; uint64_t dst_bytes;
} eb_stage;
EB_API int eb_stage_pack(const eb_stage* sg, void* stream);

/* device memory that can be shared between the processes of one box (cudaMalloc + CUDA IPC):
 * the handle is the 64-byte cudaIpcMemHandle_t; eb_ipc_open maps a peer's allocation (peer access is
 * enabled lazily by the driver) and eb_ipc_close unmaps it. */
#define EB_IPC_HANDLE_BYTES 64
EB_API int eb_dev_malloc(size_t bytes, void** out);
EB_API int eb_dev_free(void* p);
EB_API int eb_ipc_export(const void* dev_ptr, uint8_t* handle64);
EB_API int eb_ipc_open(const uint8_t* handle64, void** out);
EB_API int eb_ipc_close(void* p);


/* ======================================================================================================
 * Reversible jump + group stretch over several branches (BASELINE config 5).
 *
 * A walker owns, per branch b, up to nleaves[b] leaves of ndim[b] parameters (state.py:330 Branch); `inds` says
 * which leaves exist.  On the device the branches of a walker are stored back to back:
 *   coords [T][W][row]      row = sum_b nleaves[b]*ndim[b]; branch b starts at sum_{b'<b} nleaves[b']*ndim[b']
 *   aux    [T][W][stride]   leaf flags (sum_b nleaves[b] bytes, padded to a multiple of 4), then the friend table
 *                           int32 [sum_b nleaves[b]][nfriends] of the group move (the reference keeps it in a
 *                           BranchSupplemental, which the swap pass moves with the walker, tempering.py:351-482)
 * and the swap pass is eb_pt_swap on eb_state{nleaves = 1, ndim = row, inds = aux, inds_stride = stride}.
 * ====================================================================================================== */
#define EB_MAX_BRANCHES 4
#define EB_MB_MAX_ROW 128    /* doubles per walker */
#define EB_MB_MAX_LEAVES 64  /* leaves per walker, all branches */

typedef struct {
  int32_t nbranches;
  int32_t nfriends;                       /* width of the friend table, 0 = none */
  int32_t nleaves[EB_MAX_BRANCHES];       /* = nleaves_max */
  int32_t ndim[EB_MAX_BRANCHES];
  int32_t nleaves_min[EB_MAX_BRANCHES];   /* rj.py:33-58 */
  int32_t kind[EB_MAX_BRANCHES];          /* eb_pulse_kind of the branch's template */
  int32_t friend_key[EB_MAX_BRANCHES];    /* parameter the friends are sorted by (tests/test_eryn.py:822: index 1) */
} eb_mb_layout;

typedef struct {
  int32_t ntemps, nwalkers, temp_offset, _pad;
  double* coords;  /* [T][W][row] */
  double* logl;    /* [T][W] */
  double* logp;    /* [T][W] */
  uint8_t* aux;    /* [T][W][stride] */
  double* betas;   /* [T] or NULL */
} eb_mb_state;

/* log L = -1/2 sum_i ((template(t_i) - y_i)/sigma)^2, template = sum of the active leaves' pulses — the reference
 * test's likelihood (tests/test_eryn.py:38-92).  Walkers without any active leaf or with logp = -inf are not
 * evaluated and get -1e300 (ensemble.py:1279-1282, :1486-1513). */
typedef enum { EB_PULSE_GAUSS = 0 /* a exp(-(t-b)^2/(2c^2)) */, EB_PULSE_SINE = 1 /* a sin(2 pi b t + c) */ } eb_pulse_kind;
typedef struct {
  int32_t nt, _pad;
  double sigma;
  const double* t;  /* [nt] device */
  const double* y;  /* [nt] device */
} eb_pulse_data;

/* stationary friends of the group move (group.py:50-95; rule of tests/test_eryn.py:813-907): per branch the cold
 * chain's active leaves sorted by their key parameter, built on the host every n_iter_update iterations */
typedef struct {
  int32_t nfr[EB_MAX_BRANCHES];
  const double* coords[EB_MAX_BRANCHES]; /* [nfr][ndim] device */
  const double* keys[EB_MAX_BRANCHES];   /* [nfr] device, ascending */
} eb_mb_friends;

typedef struct {
  int32_t mode, _pad;      /* eb_rng_mode */
  const int32_t* pick;     /* replay [T][W][sum nleaves]: column of the friend table (fixture's randint) */
  const double* u_z;       /* replay [T][W]  stretch.py:131 */
  const double* u_acc;     /* replay [T][W]  group.py:254 */
  uint64_t seed; const uint64_t* iter_dev; uint64_t iter;
} eb_mb_group_rng;

typedef struct {
  int32_t mode;
  uint32_t branch_mask;    /* Gibbs split over branches (rj.py:168-343 with gibbs_sampling_setup = branch names, as
                              rj_moves="iterate_branches" / "separate_branches" set it up, ensemble.py:434-470): bit b set =
                              branch b gets a birth/death proposal in this call; 0 = all branches.  With a mask the prior of
                              a proposal that leaves the selected branches empty is fixed up as fix_logp_gibbs does
                              (move.py:369-402). */
  const int32_t* change;   /* replay [B][T][W] +1 / -1 / 0 after the edge fix-up (distgenrj.py:61-71) */
  const int32_t* leaf;     /* replay [B][T][W] leaf that is born or dies (distgenrj.py:97,111) */
  const double* birth[EB_MAX_BRANCHES]; /* replay [T][W][ndim_b] prior draws for the births (prior.py:56-71) */
  const double* u_acc;     /* replay [T][W]  rj.py:332 */
  uint64_t seed; const uint64_t* iter_dev; uint64_t iter;
  int32_t gibbs_index, _pad2; /* index of the split inside one propose() call: keys the philox streams */
} eb_mb_rj_rng;

/* compute_log_prior + compute_log_like of the whole state (ensemble.py:898-912) */
EB_API int eb_mb_eval_state(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior,
                            const eb_pulse_data* data, void* stream);
/* friend table: mode 0 = every active leaf gets its nfriends nearest friends, inactive leaves -1 (fixture
 * setup_friends); mode 1 = only active leaves whose row is all -1 (fix_friends) */
EB_API int eb_mb_friends_update(const eb_mb_layout* lay, const eb_mb_state* st, const eb_mb_friends* fr, int32_t mode,
                                void* stream);
/* GroupStretchMove.propose without the tempering tail: group.py:122-270 + groupstretch.py:34-120 */
EB_API int eb_mb_group_stretch(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior,
                               const eb_pulse_data* data, const eb_mb_friends* fr, double a, const eb_mb_group_rng* rng,
                               uint8_t* accepted, uint32_t* accepted_count, void* stream);
/* DistributionGenerateRJ.propose without the tempering tail: rj.py:145-343 + distgenrj.py:35-222 */
EB_API int eb_mb_rj_step(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior, const eb_pulse_data* data,
                         const eb_mb_rj_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream);
/* bytes per walker of the aux array for a layout */
EB_API int32_t eb_mb_aux_stride(const eb_mb_layout* lay);

/* iteration counter tick for untempered runs (no swap pass) */
EB_API int eb_advance_iter(eb_ctrl* ctrl, void* stream);

/* ---- split path for likelihoods that are not device functors (user callables on device
 *      tensors): proposal only (stretch.py:160-231) and accept+update only
 *      (red_blue.py:283-323, move.py:472).  q [T][Ns][L][D], factors [T][Ns],
 *      sub_out [T][Ns] int32 = walker ids the rows of q belong to. */
EB_API int eb_stretch_propose(const eb_state* st, double a, int32_t split, const eb_stretch_rng* rng,
                       const double* period /* [D] as eb_prior.period, or NULL */,
                       double* q, double* factors, int32_t* sub_out, void* stream);
EB_API int eb_accept_update(const eb_state* st, const int32_t* sub, int32_t nsub, const double* q,
                     const double* factors, const double* logl_new, const double* logp_new,
                     int32_t split, const eb_stretch_rng* rng,
                     uint8_t* accepted, uint32_t* accepted_count, void* stream);
/* box prior of proposed points q [T][Ns][L][D] -> logp_out [T][Ns] (ensemble.py:1192-1212) */
EB_API int eb_box_log_prior(const double* q, const uint8_t* inds_sub, int32_t nrows, int32_t nleaves,
                     int32_t ndim, const eb_prior* prior, double* logp_out, void* stream);

/* ---- reference-facing entry with HOST buffers: uploads the state, runs `niter` full
 *      iterations (move + swap pass) in philox mode, downloads the state.  This is the call
 *      bench.py's `e2e` leg times.  move_schedule_host[niter]: 0 = stretch, 1 = gaussian
 *      (the reference's per-iteration `random.choice(moves)`, ensemble.py:971); NULL = stretch.
 *      A single tempered iteration on page-locked host arrays (niter == 1, cudaHostAlloc / cudaHostRegister memory) runs
 *      as a wavefront: temperatures are uploaded hottest first in groups, every group is moved as soon as it has landed,
 *      the ladder is resolved a rung range at a time (eb_pt_swap_range) and finished rungs are downloaded while colder
 *      groups are still arriving, so the two PCIe directions and the kernels overlap; results are identical to the
 *      plain sequence.  EB_HOST_PIPE=0 selects the plain sequence, EB_HOST_GROUPS the number of groups. */
typedef struct {
  int32_t ntemps, nwalkers, nleaves, ndim;
  double* coords_host; double* logl_host; double* logp_host; double* betas_host; /* betas NULL = untempered */
  const double* prior_lo_host; const double* prior_hi_host;
  int32_t like_kind, like_ncomp, like_nparams, _pad;
  const double* like_params_host;
  double stretch_a;
  double gauss_scale;
  uint64_t seed, iter0;
  eb_adapt adapt;
  int64_t adapt_time0;
  int32_t permute, randomize_split;
  const uint8_t* move_schedule_host;
  int32_t* swaps_accepted_host;   /* [T-1] of the last iteration, nullable */
  uint32_t* accepted_count_host;  /* [T][W] accumulated over niter, nullable */
} eb_host_job;
EB_API int eb_run_host(eb_host_job* job, int32_t niter);

#ifdef __cplusplus
}
#endif
#endif /* ERYN_B200_H */
