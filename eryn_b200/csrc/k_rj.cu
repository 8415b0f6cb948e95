// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
//
// K6-K9: reversible jump + group stretch over several branches (BASELINE config 5).  One warp owns one walker: the
// proposal row and the leaf flags are staged in shared memory, the prior is evaluated leaf-parallel, the likelihood
// (sum of the active leaves' pulses against a data vector) time-point-parallel with a warp reduction.
//   K6 mb_eval_kernel           log-prior + log-like of the whole state               (ensemble.py:898-912)
//   K7 mb_friends_kernel        nearest-friends table of the group move               (tests/test_eryn.py:813-907)
//   K8 mb_group_stretch_kernel  GroupStretchMove.propose                              (group.py:122-270)
//   K9 mb_rj_kernel             DistributionGenerateRJ.propose                        (rj.py:145-343, distgenrj.py:35-222)
#include "common.cuh"

namespace eb {

constexpr int MB_THREADS = 128;          // 4 walkers per CTA
constexpr int MB_WARPS = MB_THREADS / 32;
constexpr uint32_t TAG_GROUP = 8;

struct MBArgs {
  int nb, nfriends, row, Ltot, flags_pad, aux_stride;
  int L[EB_MAX_BRANCHES], D[EB_MAX_BRANCHES], coff[EB_MAX_BRANCHES], loff[EB_MAX_BRANCHES], poff[EB_MAX_BRANCHES];
  int kind[EB_MAX_BRANCHES], nmin[EB_MAX_BRANCHES], key[EB_MAX_BRANCHES];
  int T, W, t0;
  double* coords; double* logl; double* logp; uint8_t* aux; double* betas;
  const double* lo; const double* hi; const double* lpdf;
  int nt; double sigma; const double* tt; const double* yy;
  // friends
  int nfr[EB_MAX_BRANCHES]; const double* fcoords[EB_MAX_BRANCHES]; const double* fkeys[EB_MAX_BRANCHES]; int fmode;
  // rng
  int philox; uint32_t seed_lo, seed_hi; const unsigned long long* iter_dev; unsigned long long iter;
  double a;
  const int32_t* pick; const double* u_z; const double* u_acc;
  const int32_t* change; const int32_t* leaf; const double* birth[EB_MAX_BRANCHES];
  uint8_t* accepted; uint32_t* accepted_count;
  uint32_t bmask; int gidx;   // reversible jump: Gibbs split over branches (0 = all), index of the split (stream key)
};

struct WarpSm {
  double q[EB_MB_MAX_ROW];
  double leafp[EB_MB_MAX_LEAVES];
  uint8_t flags[EB_MB_MAX_LEAVES];
};

__device__ __forceinline__ int branch_of(const MBArgs& p, int j) {
  int b = 0;
  while (b + 1 < p.nb && j >= p.loff[b + 1]) ++b;
  return b;
}

// log-prior and gated log-like of the row in ws.q with the leaf flags in ws.flags; every lane gets the results
__device__ __forceinline__ void mb_eval(const MBArgs& p, WarpSm& ws, int lane, double& lp, double& ll, uint32_t bmask = 0u) {
  // ---- prior: one leaf per lane; parameters added in index order starting from 0.0 (prior.py:369-385)
  for (int j = lane; j < p.Ltot; j += 32) {
    const int b = branch_of(p, j);
    double s = 0.0;
    if (ws.flags[j]) {
      const double* x = ws.q + p.coff[b] + (j - p.loff[b]) * p.D[b];
      for (int d = 0; d < p.D[b]; ++d) {
        const double v = x[d], l = p.lo[p.poff[b] + d], h = p.hi[p.poff[b] + d];
        double t = 0.0;
        if (v >= l && v <= h) t = p.lpdf[p.poff[b] + d];
        if (v < l || v > h) t = neg_inf();
        s += t;
      }
    }
    ws.leafp[j] = s;                                   // inactive leaves contribute 0 (ensemble.py:1207)
  }
  __syncwarp();
  lp = 0.0;
  bool any_leaf = false;
  for (int b = 0; b < p.nb; ++b) {                     // leaves summed per branch, branches added in order (:1210)
    double sb = 0.0;
    for (int l = 0; l < p.L[b]; ++l) {
      sb += ws.leafp[p.loff[b] + l];
      any_leaf |= ws.flags[p.loff[b] + l] != 0;
    }
    lp += sb;
  }
  if (bmask) {   // fix_logp_gibbs (move.py:369-402): leaves of the branches of this Gibbs split vs all leaves
    bool here = false;
    for (int b = 0; b < p.nb; ++b)
      if ((bmask >> b) & 1u)
        for (int l = 0; l < p.L[b]; ++l) here |= ws.flags[p.loff[b] + l] != 0;
    if (any_leaf && !here) lp = neg_inf();    // no use in running because no change
    if (!any_leaf && !here) lp = 0.0;         // there is nothing in the model currently
  }
  // ---- likelihood: not evaluated outside the prior or without leaves (ensemble.py:1279-1282, :1486-1513)
  if (isinf(lp) || !any_leaf) {
    ll = FILL_LOGL;
    return;
  }
  double part = 0.0;
  for (int i = lane; i < p.nt; i += 32) {
    const double t = p.tt[i];
    double tmpl = 0.0;
    for (int b = 0; b < p.nb; ++b) {
      for (int l = 0; l < p.L[b]; ++l) {
        if (!ws.flags[p.loff[b] + l]) continue;
        const double* x = ws.q + p.coff[b] + l * p.D[b];
        const double a = x[0], bb = x[1], c = x[2];
        if (p.kind[b] == EB_PULSE_GAUSS) {
          const double dt = t - bb;
          tmpl += a * exp(-(dt * dt) / (2.0 * (c * c)));              // tests/test_eryn.py:38-40
        } else {
          tmpl += a * sin(6.283185307179586 * bb * t + c);             // tests/test_eryn.py:69-71
        }
      }
    }
    const double r = (tmpl - p.yy[i]) / p.sigma;
    part += r * r;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  ll = -0.5 * part;
  if (ll != ll) ll = FILL_LOGL;
}

__device__ __forceinline__ void mb_load_walker(const MBArgs& p, WarpSm& ws, size_t slot, int lane) {
  for (int e = lane; e < p.row; e += 32) ws.q[e] = p.coords[slot * p.row + e];
  for (int j = lane; j < p.Ltot; j += 32) ws.flags[j] = p.aux[slot * p.aux_stride + j];
  __syncwarp();
}

// Metropolis test + Move.update (move.py:472-703) for the proposal staged in ws
__device__ __forceinline__ void mb_accept(const MBArgs& p, WarpSm& ws, size_t slot, int t, int lane, double factors,
                                          double u_acc, bool flags_changed, uint32_t bmask = 0u) {
  double lp, ll;
  mb_eval(p, ws, lane, lp, ll, bmask);
  const bool tempered = p.betas != nullptr;
  const double beta = tempered ? p.betas[t] : 1.0;
  const double logP = log_posterior(ll, lp, beta, tempered);
  const double prevP = log_posterior(p.logl[slot], p.logp[slot], beta, tempered);
  const bool keep = (factors + logP - prevP) > log(u_acc);
  if (keep) {
    for (int e = lane; e < p.row; e += 32) p.coords[slot * p.row + e] = ws.q[e];
    if (flags_changed)
      for (int j = lane; j < p.Ltot; j += 32) p.aux[slot * p.aux_stride + j] = ws.flags[j];
    if (lane == 0) {
      p.logl[slot] = ll;
      p.logp[slot] = isinf(lp) ? 0.0 : lp;                               // move.py:526
      if (p.accepted_count) p.accepted_count[slot] += 1u;
    }
  }
  if (lane == 0) p.accepted[slot] = keep ? 1 : 0;
}

__global__ void __launch_bounds__(MB_THREADS) mb_eval_kernel(const __grid_constant__ MBArgs p) {
  __shared__ WarpSm sm[MB_WARPS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slot = (size_t)blockIdx.x * MB_WARPS + wid;
  if (slot >= (size_t)p.T * p.W) return;
  WarpSm& ws = sm[wid];
  mb_load_walker(p, ws, slot, lane);
  double lp, ll;
  mb_eval(p, ws, lane, lp, ll);
  if (lane == 0) { p.logp[slot] = lp; p.logl[slot] = ll; }
}

// nearest friends of value v among the ascending keys[0..n): indices in order of increasing distance
__device__ __forceinline__ void nearest_friends(const double* __restrict__ keys, int n, double v, int nf, int32_t* out) {
  int lo = 0, hi = n;                    // first index with keys[idx] >= v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  int l = lo - 1, r = lo;
  for (int k = 0; k < nf; ++k) {
    const bool has_l = l >= 0, has_r = r < n;
    if (!has_l && !has_r) { out[k] = -1; continue; }
    bool take_l;
    if (has_l && has_r) take_l = fabs(v - keys[l]) <= fabs(v - keys[r]);
    else take_l = has_l;
    if (take_l) out[k] = l--; else out[k] = r++;
  }
}

__global__ void __launch_bounds__(MB_THREADS) mb_friends_kernel(const __grid_constant__ MBArgs p) {
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)p.T * p.W * p.Ltot;
  if (id >= n) return;
  const size_t slot = id / p.Ltot;
  const int j = (int)(id - slot * p.Ltot);
  const int b = branch_of(p, j);
  const bool active = p.aux[slot * p.aux_stride + j] != 0;
  int32_t* tab = reinterpret_cast<int32_t*>(p.aux + slot * p.aux_stride + p.flags_pad) + (size_t)j * p.nfriends;
  if (p.fmode == 0) {
    if (!active) {
      for (int k = 0; k < p.nfriends; ++k) tab[k] = -1;
      return;
    }
  } else {
    if (!active) return;
    for (int k = 0; k < p.nfriends; ++k)
      if (tab[k] != -1) return;          // only leaves without an assigned row (born through RJ)
  }
  const double v = p.coords[slot * p.row + p.coff[b] + (j - p.loff[b]) * p.D[b] + p.key[b]];
  nearest_friends(p.fkeys[b], p.nfr[b], v, p.nfriends, tab);
}

__global__ void __launch_bounds__(MB_THREADS) mb_group_stretch_kernel(const __grid_constant__ MBArgs p) {
  __shared__ WarpSm sm[MB_WARPS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slot = (size_t)blockIdx.x * MB_WARPS + wid;
  if (slot >= (size_t)p.T * p.W) return;
  const int t = (int)(slot / p.W);
  WarpSm& ws = sm[wid];
  mb_load_walker(p, ws, slot, lane);
  const uint32_t fw = (uint32_t)(slot + (size_t)p.t0 * p.W);
  RngKey key;
  double u_z, u_acc;
  if (p.philox) {
    unsigned long long it = p.iter;   // (not a ?: of the two: with a __grid_constant__ parameter block nvcc merges the
  // arms into ONE global load whose address may then point into parameter space)
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    const uint4 r = stream(key, TAG_GROUP, fw, 0u);
    u_z = u01_52(r.x, r.y);
    u_acc = u01_52(r.z, r.w);
  } else {
    u_z = p.u_z[slot];
    u_acc = p.u_acc[slot];
  }
  double zz = (p.a - 1.0) * u_z + 1.0;                                     // stretch.py:129-132
  zz = zz * zz / p.a;
  const double factors = ((double)p.row - 1.0) * log(zz);                  // groupstretch.py:112
  const int32_t* tab = reinterpret_cast<const int32_t*>(p.aux + slot * p.aux_stride + p.flags_pad);
  for (int j = lane; j < p.Ltot; j += 32) {
    const int b = branch_of(p, j);
    double* x = ws.q + p.coff[b] + (j - p.loff[b]) * p.D[b];
    const double* fr = nullptr;
    if (ws.flags[j]) {                                                     // fixture find_friends
      int pk;
      if (p.philox) {
        const uint4 r = stream(key, TAG_GROUP, fw, (uint32_t)(1 + (j >> 2)));
        const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
        pk = (int)__umulhi(w, (uint32_t)p.nfriends);
      } else {
        pk = p.pick[slot * p.Ltot + j];
      }
      int idx = tab[(size_t)j * p.nfriends + pk];
      if (idx < 0) idx += p.nfr[b];                                        // numpy negative index
      fr = p.fcoords[b] + (size_t)idx * p.D[b];
    }
    for (int d = 0; d < p.D[b]; ++d) {
      const double c = fr ? fr[d] : 0.0;                                   // friends = zeros_like(s) for inactive leaves
      x[d] = c - (c - x[d]) * zz;                                          // stretch.py:143-145
    }
  }
  __syncwarp();
  mb_accept(p, ws, slot, t, lane, factors, u_acc, false);
}

__device__ __forceinline__ double leaf_logpdf(const MBArgs& p, int b, const double* x) {
  double s = 0.0;
  for (int d = 0; d < p.D[b]; ++d) {
    const double v = x[d], l = p.lo[p.poff[b] + d], h = p.hi[p.poff[b] + d];
    double t = 0.0;
    if (v >= l && v <= h) t = p.lpdf[p.poff[b] + d];
    if (v < l || v > h) t = neg_inf();
    s += t;
  }
  return s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_rj_kernel(const __grid_constant__ MBArgs p) {
  __shared__ WarpSm sm[MB_WARPS];
  __shared__ double s_factors[MB_WARPS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slot = (size_t)blockIdx.x * MB_WARPS + wid;
  if (slot >= (size_t)p.T * p.W) return;
  const int t = (int)(slot / p.W);
  const size_t TW = (size_t)p.T * p.W;
  WarpSm& ws = sm[wid];
  mb_load_walker(p, ws, slot, lane);
  const uint32_t fw = (uint32_t)(slot + (size_t)p.t0 * p.W);
  RngKey key;
  double u_acc = 0.5;
  if (p.philox) {
    unsigned long long it = p.iter;   // (not a ?: of the two: with a __grid_constant__ parameter block nvcc merges the
  // arms into ONE global load whose address may then point into parameter space)
  if (p.iter_dev) it = ld_volatile_u64(p.iter_dev);
    key = make_rng_key(p.seed_lo, p.seed_hi, it);
    const uint4 r = stream(key, TAG_RJ, fw, (uint32_t)p.gidx << 8);
    u_acc = u01_52(r.z, r.w);
  } else {
    u_acc = p.u_acc[slot];
  }
  if (lane == 0) {                       // the bookkeeping of <= 4 branches x a few leaves is serial
    double factors = 0.0, edge = 0.0;
    const double lhalf = log(1.0 / 2.0);
    for (int b = 0; b < p.nb; ++b) {
      const int L = p.L[b], nmin = p.nmin[b], nmax = L;
      if (nmin == nmax) continue;                                          // distgenrj.py:166-167
      if (p.bmask && !((p.bmask >> b) & 1u)) continue;                     // not in this Gibbs split (rj.py:168-176)
      int nl = 0;
      for (int l = 0; l < L; ++l) nl += ws.flags[p.loff[b] + l] ? 1 : 0;
      int change, lf;
      if (p.philox) {
        const uint4 r = stream(key, TAG_RJ, fw, (uint32_t)(8 * b) | ((uint32_t)p.gidx << 8));
        change = (r.x & 1u) ? +1 : -1;                                     // distgenrj.py:62
        if (nl == nmin) change = +1;                                       // :67-71
        if (nl == nmax) change = -1;
        const int ncand = change == +1 ? L - nl : nl;
        int k = (int)__umulhi(r.y, (uint32_t)ncand);                       // :97, :111: uniform among the candidates
        lf = -1;
        for (int l = 0; l < L; ++l) {
          const bool cand = (ws.flags[p.loff[b] + l] != 0) == (change == -1);
          if (cand && k-- == 0) { lf = l; break; }
        }
      } else {
        change = p.change[(size_t)b * TW + slot];
        lf = p.leaf[(size_t)b * TW + slot];
      }
      double* x = ws.q + p.coff[b] + lf * p.D[b];
      if (change == -1) {
        ws.flags[p.loff[b] + lf] = 0;
        factors += +1.0 * leaf_logpdf(p, b, x);                            // distgenrj.py:201-203
      } else if (change == +1) {
        ws.flags[p.loff[b] + lf] = 1;
        for (int d = 0; d < p.D[b]; ++d) {
          if (p.philox) {
            const uint4 r = stream(key, TAG_RJ, fw, (uint32_t)(8 * b + 1 + (d >> 1)) | ((uint32_t)p.gidx << 8));
            const double u = (d & 1) ? u01_52(r.z, r.w) : u01_52(r.x, r.y);
            const double l = p.lo[p.poff[b] + d], h = p.hi[p.poff[b] + d];
            x[d] = u * (h - l) + l;                                        // prior.py:66
          } else {
            x[d] = p.birth[b][slot * p.D[b] + d];
          }
        }
        factors += -1.0 * leaf_logpdf(p, b, x);                            // distgenrj.py:217-219
      }
      if (nmin + 1 != nmax) {                                              // rj.py:241-271
        const int nnew = nl + change;
        if (nl == nmin) edge += lhalf;
        if (nl == nmax) edge += lhalf;
        if (nnew == nmin) edge -= lhalf;
        if (nnew == nmax) edge -= lhalf;
      }
    }
    s_factors[wid] = factors + edge;                                       // rj.py:273
  }
  __syncwarp();
  mb_accept(p, ws, slot, t, lane, s_factors[wid], u_acc, true, p.bmask);
}

static int fill_mb(MBArgs& a, const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior,
                   const eb_pulse_data* data, bool need_prob) {
  if (!lay || !st) return fail(EB_ERR_INVALID, "layout/state is NULL");
  memset(&a, 0, sizeof(a));
  if (lay->nbranches < 1 || lay->nbranches > EB_MAX_BRANCHES)
    return fail(EB_ERR_INVALID, "nbranches must be 1..%d", EB_MAX_BRANCHES);
  a.nb = lay->nbranches;
  a.nfriends = lay->nfriends;
  int co = 0, lo = 0, po = 0;
  for (int b = 0; b < a.nb; ++b) {
    if (lay->nleaves[b] < 1 || lay->ndim[b] < 1) return fail(EB_ERR_INVALID, "incompatible input dimensions (branch %d)", b);
    if (lay->nleaves_min[b] < 0 || lay->nleaves_min[b] > lay->nleaves[b])
      return fail(EB_ERR_INVALID, "nleaves_min cannot be greater than nleaves_max.");
    if (need_prob && lay->ndim[b] != 3) return fail(EB_ERR_UNSUPPORTED, "the pulse likelihood has 3 parameters per leaf");
    a.L[b] = lay->nleaves[b]; a.D[b] = lay->ndim[b]; a.coff[b] = co; a.loff[b] = lo; a.poff[b] = po;
    a.kind[b] = lay->kind[b]; a.nmin[b] = lay->nleaves_min[b]; a.key[b] = lay->friend_key[b];
    if (a.key[b] < 0 || a.key[b] >= a.D[b]) return fail(EB_ERR_INVALID, "friend_key out of range");
    co += a.L[b] * a.D[b]; lo += a.L[b]; po += a.D[b];
  }
  a.row = co; a.Ltot = lo;
  if (a.row > EB_MB_MAX_ROW || a.Ltot > EB_MB_MAX_LEAVES)
    return fail(EB_ERR_UNSUPPORTED, "walker of %d doubles / %d leaves exceeds %d / %d", a.row, a.Ltot, EB_MB_MAX_ROW,
                EB_MB_MAX_LEAVES);
  a.flags_pad = (a.Ltot + 3) & ~3;
  a.aux_stride = a.flags_pad + 4 * a.Ltot * (a.nfriends > 0 ? a.nfriends : 0);
  if (st->ntemps < 1 || st->nwalkers < 1) return fail(EB_ERR_INVALID, "incompatible input dimensions");
  if (!st->coords || !st->logl || !st->logp || !st->aux) return fail(EB_ERR_INVALID, "coords/logl/logp/aux must be device pointers");
  a.T = st->ntemps; a.W = st->nwalkers; a.t0 = st->temp_offset;
  a.coords = st->coords; a.logl = st->logl; a.logp = st->logp; a.aux = st->aux; a.betas = st->betas;
  if (need_prob) {
    if (!prior || !prior->lo || !prior->hi || !prior->logpdf) return fail(EB_ERR_INVALID, "prior is required");
    if (!data || !data->t || !data->y || data->nt < 1 || !(data->sigma > 0.0)) return fail(EB_ERR_INVALID, "pulse data is required");
    a.lo = prior->lo; a.hi = prior->hi; a.lpdf = prior->logpdf;
    a.nt = data->nt; a.sigma = data->sigma; a.tt = data->t; a.yy = data->y;
  }
  return EB_OK;
}

static int fill_friends(MBArgs& a, const eb_mb_friends* fr) {
  if (!fr) return fail(EB_ERR_INVALID, "friends are NULL");
  if (a.nfriends < 1) return fail(EB_ERR_INVALID, "the layout has no friend table (nfriends = 0)");
  for (int b = 0; b < a.nb; ++b) {
    if (fr->nfr[b] < 1 || !fr->coords[b] || !fr->keys[b]) return fail(EB_ERR_INVALID, "branch %d has no friends", b);
    a.nfr[b] = fr->nfr[b]; a.fcoords[b] = fr->coords[b]; a.fkeys[b] = fr->keys[b];
  }
  return EB_OK;
}

static void fill_philox(MBArgs& a, uint64_t seed, const uint64_t* iter_dev, uint64_t iter) {
  a.philox = 1;
  a.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull); a.seed_hi = (uint32_t)(seed >> 32);
  a.iter_dev = (const unsigned long long*)iter_dev; a.iter = iter;
}

static inline int walker_grid(const MBArgs& a) { return (int)(((size_t)a.T * a.W + MB_WARPS - 1) / MB_WARPS); }

}  // namespace eb

using namespace eb;

extern "C" {

int32_t eb_mb_aux_stride(const eb_mb_layout* lay) {
  if (!lay) return -1;
  int lt = 0;
  for (int b = 0; b < lay->nbranches && b < EB_MAX_BRANCHES; ++b) lt += lay->nleaves[b];
  return ((lt + 3) & ~3) + 4 * lt * (lay->nfriends > 0 ? lay->nfriends : 0);
}

int eb_mb_eval_state(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior, const eb_pulse_data* data,
                     void* stream) {
  MBArgs a;
  int rc = fill_mb(a, lay, st, prior, data, true);
  if (rc) return rc;
  mb_eval_kernel<<<walker_grid(a), MB_THREADS, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mb_eval");
}

int eb_mb_friends_update(const eb_mb_layout* lay, const eb_mb_state* st, const eb_mb_friends* fr, int32_t mode,
                         void* stream) {
  MBArgs a;
  int rc = fill_mb(a, lay, st, nullptr, nullptr, false);
  if (rc) return rc;
  rc = fill_friends(a, fr);
  if (rc) return rc;
  if (mode != 0 && mode != 1) return fail(EB_ERR_INVALID, "mode must be 0 (setup) or 1 (fix)");
  a.fmode = mode;
  const size_t n = (size_t)a.T * a.W * a.Ltot;
  mb_friends_kernel<<<(unsigned)((n + MB_THREADS - 1) / MB_THREADS), MB_THREADS, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mb_friends");
}

int eb_mb_group_stretch(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior, const eb_pulse_data* data,
                        const eb_mb_friends* fr, double stretch_a, const eb_mb_group_rng* rng, uint8_t* accepted,
                        uint32_t* accepted_count, void* stream) {
  MBArgs a;
  int rc = fill_mb(a, lay, st, prior, data, true);
  if (rc) return rc;
  rc = fill_friends(a, fr);
  if (rc) return rc;
  if (!rng || !accepted) return fail(EB_ERR_INVALID, "rng/accepted is NULL");
  if (!(stretch_a > 1.0)) return fail(EB_ERR_INVALID, "stretch scale a must be > 1");
  a.a = stretch_a;
  if (rng->mode == EB_RNG_PHILOX) {
    fill_philox(a, rng->seed, rng->iter_dev, rng->iter);
  } else if (rng->mode == EB_RNG_REPLAY) {
    if (!rng->pick || !rng->u_z || !rng->u_acc) return fail(EB_ERR_INVALID, "replay mode needs pick, u_z, u_acc");
    a.pick = rng->pick; a.u_z = rng->u_z; a.u_acc = rng->u_acc;
  } else {
    return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  }
  a.accepted = accepted; a.accepted_count = accepted_count;
  mb_group_stretch_kernel<<<walker_grid(a), MB_THREADS, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mb_group_stretch");
}

int eb_mb_rj_step(const eb_mb_layout* lay, const eb_mb_state* st, const eb_prior* prior, const eb_pulse_data* data,
                  const eb_mb_rj_rng* rng, uint8_t* accepted, uint32_t* accepted_count, void* stream) {
  MBArgs a;
  int rc = fill_mb(a, lay, st, prior, data, true);
  if (rc) return rc;
  if (!rng || !accepted) return fail(EB_ERR_INVALID, "rng/accepted is NULL");
  const uint32_t bm = rng->branch_mask;
  if (rng->mode == EB_RNG_PHILOX) {
    fill_philox(a, rng->seed, rng->iter_dev, rng->iter);
  } else if (rng->mode == EB_RNG_REPLAY) {
    if (!rng->change || !rng->leaf || !rng->u_acc) return fail(EB_ERR_INVALID, "replay mode needs change, leaf, u_acc");
    a.change = rng->change; a.leaf = rng->leaf; a.u_acc = rng->u_acc;
    for (int b = 0; b < a.nb; ++b) {
      if (a.nmin[b] != a.L[b] && (!bm || ((bm >> b) & 1u)) && !rng->birth[b])
        return fail(EB_ERR_INVALID, "replay mode needs the birth draws of branch %d", b);
      a.birth[b] = rng->birth[b];
    }
  } else {
    return fail(EB_ERR_INVALID, "unknown rng mode %d", rng->mode);
  }
  a.bmask = rng->branch_mask; a.gidx = rng->gibbs_index;
  if (a.nb < 32 && (a.bmask >> a.nb)) return fail(EB_ERR_INVALID, "branch_mask selects branches the layout does not have");
  if (a.gidx < 0 || a.gidx > 0xFFFF) return fail(EB_ERR_INVALID, "gibbs_index out of range");
  bool any = false;
  for (int b = 0; b < a.nb; ++b) any |= a.nmin[b] != a.L[b] && (!a.bmask || ((a.bmask >> b) & 1u));
  if (!any)
    return fail(EB_ERR_INVALID, "Right now, no models are getting a reversible jump proposal. Check nleaves_min and "
                                "nleaves_max or do not use rj proposal.");
  a.accepted = accepted; a.accepted_count = accepted_count;
  mb_rj_kernel<<<walker_grid(a), MB_THREADS, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mb_rj");
}

}  // extern "C"
