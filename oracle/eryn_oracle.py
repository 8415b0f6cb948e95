"""ORACLE — CPU (NumPy) restatement of Eryn's walker-parallel sampling hot path.

THIS IS TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and the
CPU-baseline legs of `bench.py` may import it; the product package
(`eryn_b200/`) never does and fails loudly without its CUDA library.

What is restated (file:line into /root/reference/src/eryn):
  box_log_prior            prior.py:80-88 (UniformDistribution.logpdf), prior.py:337-392
                           (ProbDistContainer.logpdf), ensemble.py:1192-1212 (rectangular route)
  log_like                 ensemble.py:1219-1545 (mask logp=-inf, -1e300 fill)
  tempered_log_posterior   tempering.py:284-349
  stretch_half_step        red_blue.py:148-323, stretch.py:74-231
  gaussian_step            mh.py:56-193, gaussian.py:68-195
  update_subset            move.py:472-703 (coords/logl/logp/inds part)
  temperature_swaps        tempering.py:484-561 with do_swaps_indexing :351-482
  adapt_temps              tempering.py:563-596
  OracleSampler.iterate    ensemble.py:965-984 (+ red_blue.py:119-124, :326-331)

Pinning: `tests/golden/make_golden.py` runs the *unmodified reference* (imported from
/root/reference through the stub shim of SURVEY.md App. A) under fixed seeds and stores
per-iteration outputs in `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays the
same seeds through this file and requires bit-equal accept masks / swap counts and
float agreement to 1e-13.  Known answers KAT-1 / KAT-2 of SURVEY.md App. C are checked too.

Random sources.  `NumpyStreams` consumes the two MT19937 streams in exactly the
reference's order.  `PhiloxStreams` is the production-mode
source of the CUDA path (oracle/philox_np.py); the *use* of each draw is the same.
"""
import numpy as np

from . import philox_np as px

FILL = -1e300  # ensemble.py:1486,1499


# ----------------------------------------------------------------------------------
# target densities used by the BASELINE.json configs
# ----------------------------------------------------------------------------------
class GaussianLike:
    """log L = -1/2 (x-mu)^T P (x-mu)   (tests/test_eryn.py:33-35 vectorised)."""

    kind = 0

    def __init__(self, mu, prec):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.prec = np.ascontiguousarray(prec, dtype=np.float64)

    def __call__(self, x):
        d = x - self.mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, self.prec, d)

    def params(self):
        return np.concatenate([self.mu, self.prec.ravel()])


class RosenbrockLike:
    """log L = -sum_{i<d-1} [100 (x_{i+1}-x_i^2)^2 + (1-x_i)^2]   (SURVEY.md §8d C3)."""

    kind = 1

    def __call__(self, x):
        return -np.sum(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2, axis=1)

    def params(self):
        return np.zeros(0)


class GaussianMixtureLike:
    """log L = log sum_k w_k N(x; mu_k, sigma_k^2 I)   (SURVEY.md §8d C4)."""

    kind = 2

    def __init__(self, mus, sigmas, weights):
        self.mus = np.ascontiguousarray(mus, dtype=np.float64)  # [K, D]
        self.sigmas = np.asarray(sigmas, dtype=np.float64)  # [K]
        self.weights = np.asarray(weights, dtype=np.float64)  # [K]
        K, D = self.mus.shape
        # per-component constant log w_k - D log sigma_k - D/2 log 2pi
        self.logc = np.log(self.weights) - D * np.log(self.sigmas) - 0.5 * D * np.log(2.0 * np.pi)
        self.hinv = 0.5 / self.sigmas**2

    def __call__(self, x):
        r2 = ((x[:, None, :] - self.mus[None, :, :]) ** 2).sum(-1)  # [N, K]
        e = self.logc[None, :] - r2 * self.hinv[None, :]
        m = e.max(axis=1)
        return m + np.log(np.exp(e - m[:, None]).sum(axis=1))

    def params(self):
        return np.concatenate([self.logc, self.hinv, self.mus.ravel()])


# ----------------------------------------------------------------------------------
# probability evaluation
# ----------------------------------------------------------------------------------
class BoxPrior:
    """Independent uniform priors, one per parameter (prior.py:12-91, 219-392)."""

    def __init__(self, lo, hi):
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        swap = lo > hi  # prior.py:29-32
        lo, hi = np.where(swap, hi, lo), np.where(swap, lo, hi)
        self.lo, self.hi = lo, hi
        self.logpdf_val = np.log(1.0 / (hi - lo))  # prior.py:40-41

    def logpdf(self, x):
        """x [N, D] -> [N]; sum over parameters in index order starting from 0.0 (prior.py:369-385)."""
        out = np.zeros(x.shape[0])
        for d in range(x.shape[1]):
            v = x[:, d]
            t = np.zeros_like(v)
            t[(v >= self.lo[d]) & (v <= self.hi[d])] = self.logpdf_val[d]
            t[(v < self.lo[d]) | (v > self.hi[d])] = -np.inf
            out += t
        return out

    def rvs(self, size, random):
        """prior.py:432-497 with UniformDistribution.rvs :56-71: one rand(*size) per parameter."""
        size = (size,) if isinstance(size, int) else tuple(size)
        out = np.zeros(size + (len(self.lo),))
        for d in range(len(self.lo)):
            out[..., d] = random.rand(*size) * (self.hi[d] - self.lo[d]) + self.lo[d]
        return out


def box_log_prior(prior, coords, inds):
    """ensemble.py:1192-1212: logpdf per leaf, inactive leaves -> 0, sum over leaves."""
    T, W, L, D = coords.shape
    v = prior.logpdf(coords.reshape(-1, D)).reshape(T, W, L)
    v[~inds] = 0.0
    out = np.zeros((T, W))
    out += v.sum(axis=-1)
    return out


def log_like(like, coords, inds, logp):
    """ensemble.py:1219-1545 for one branch with nleaves_max == 1 (vectorised route).

    Walkers whose logp is -inf are not evaluated and get -1e300 (:1279-1282, :1486)."""
    T, W, L, D = coords.shape
    assert L == 1, "oracle.log_like: multi-leaf likelihoods go through rj_oracle"
    if np.all(np.isinf(logp)):  # :1272-1276
        return np.full_like(logp, FILL)
    keep = inds[:, :, 0] & ~np.isinf(logp)
    ll = np.full(T * W, FILL)
    flat = keep.reshape(-1)
    if flat.any():
        ll[flat] = like(coords.reshape(T * W, D)[flat])
    return ll.reshape(T, W)


def tempered_log_posterior(logl, logp, betas):
    """tempering.py:284-349 (betas None -> move.py:443 basic posterior)."""
    if betas is None:
        return logl + logp
    with np.errstate(invalid="ignore"):
        loglT = logl * betas[:, None]
    loglT[np.isnan(loglT)] = -np.inf
    return loglT + logp


# ----------------------------------------------------------------------------------
# state
# ----------------------------------------------------------------------------------
class OState:
    """coords [T,W,L,D] f64, inds [T,W,L] bool, logl/logp [T,W] f64 (state.py:387-470)."""

    def __init__(self, coords, inds=None, logl=None, logp=None):
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim == 2:
            coords = coords[None, :, None, :]
        elif coords.ndim == 3:
            coords = coords[:, :, None, :]
        self.coords = coords.copy()
        self.inds = np.ones(coords.shape[:3], dtype=bool) if inds is None else inds.copy()
        self.logl = None if logl is None else logl.copy()
        self.logp = None if logp is None else logp.copy()

    def copy(self):
        return OState(self.coords, self.inds, self.logl, self.logp)


def update_subset(state, sub, q, new_logl, new_logp, keep):
    """move.py:472-703 restricted to coords/logl/logp (inds unchanged by in-model moves).

    sub [T,Ns] walker ids, q [T,Ns,L,D]; the boolean-multiply blend of the reference is kept
    literally so that +-0.0 and NaN propagation match."""
    old_ll = np.take_along_axis(state.logl, sub, axis=1)
    np.put_along_axis(state.logl, sub, new_logl * keep + old_ll * (~keep), axis=1)
    old_lp = np.take_along_axis(state.logp, sub, axis=1)
    nlp = new_logp.copy()
    nlp[np.isinf(nlp)] = 0.0  # move.py:526
    np.put_along_axis(state.logp, sub, nlp * keep + old_lp * (~keep), axis=1)
    old_c = np.take_along_axis(state.coords, sub[:, :, None, None], axis=1)
    tmp = old_c.copy()
    tmp[keep] = q[keep]  # move.py:666-667
    np.put_along_axis(state.coords, sub[:, :, None, None], tmp, axis=1)


# ----------------------------------------------------------------------------------
# moves
# ----------------------------------------------------------------------------------
def periodic_distance(s, c, periods):
    """utils/periodic.py:49-117 with p1 = s, p2 = c (stretch.py:136-141): c - s, through the boundary when that is
    shorter.  periods [D]: period of each parameter, 0 = not periodic."""
    diff = c - s
    for d in np.nonzero(periods)[0]:
        P = periods[d]
        dp = diff[..., d].copy()
        fix = np.abs(dp) > P / 2.0
        new_s = -(P - s[..., d]) * (dp < 0.0) + (P + s[..., d]) * (dp >= 0.0)
        dp[fix] = c[..., d][fix] - new_s[fix]
        diff[..., d] = dp
    return diff


def periodic_wrap(q, periods):
    """utils/periodic.py:119-151: q % period on the periodic parameters (in place)."""
    for d in np.nonzero(periods)[0]:
        q[..., d] = q[..., d] % periods[d]
    return q


def stretch_half_step(state, sub, comp, rint, u_z, u_acc, a, betas, prior, like, periods=None, gibbs_mask=None):
    """One red/blue half step (red_blue.py:148-323 + stretch.py:74-231).

    sub [T,Ns] / comp [T,Nc]: walker ids of the moving subset / the complement, in the order
    the random draws are indexed by.  Returns keep [T,Ns] bool and the intermediate values."""
    T, W, L, D = state.coords.shape
    s = np.take_along_axis(state.coords, sub[:, :, None, None], axis=1)
    c = np.take_along_axis(state.coords, comp[:, :, None, None], axis=1)
    c_temp = np.take_along_axis(c, rint[:, :, None, None], axis=1)  # stretch.py:100
    zz = ((a - 1.0) * u_z + 1) ** 2.0 / a  # stretch.py:129-132
    if periods is not None:
        q = periodic_wrap(c_temp - periodic_distance(s, c_temp, periods) * zz[:, :, None, None], periods)  # :136-153
    else:
        q = c_temp - (c_temp - s) * zz[:, :, None, None]  # stretch.py:143-145
    factors = (L * D - 1.0) * np.log(zz)  # stretch.py:223
    if gibbs_mask is not None:
        # Gibbs split (move.py:113-402): only the selected parameters move (cleanup_proposals_gibbs, move.py:302-307) and
        # the detailed-balance factor counts them only (red_blue.py:196-207, stretch.py:55-72 adjust_factors)
        q[:, :, ~gibbs_mask] = s[:, :, ~gibbs_mask]
        g = float(gibbs_mask.sum())
        if g != L * D:
            factors = factors / (L * D - 1.0) * (g - 1.0)
    new_inds = np.take_along_axis(state.inds, sub[:, :, None], axis=1)
    logp = box_log_prior(prior, q, new_inds)  # red_blue.py:260
    logl = log_like(like, q, new_inds, logp)  # red_blue.py:270
    logl[np.isnan(logl)] = FILL  # red_blue.py:279-281
    logP = tempered_log_posterior(logl, logp, betas)
    prev_logl = np.take_along_axis(state.logl, sub, axis=1)
    prev_logp = np.take_along_axis(state.logp, sub, axis=1)
    prev_logP = tempered_log_posterior(prev_logl, prev_logp, betas)
    lnpdiff = factors + logP - prev_logP  # red_blue.py:292
    keep = lnpdiff > np.log(u_acc)  # red_blue.py:294
    update_subset(state, sub, q, logl, logp, keep)
    return keep, dict(q=q, logl=logl, logp=logp, zz=zz, lnpdiff=lnpdiff)


def gaussian_step(state, delta, u_acc, betas, prior, like, periods=None, gibbs_mask=None):
    """MH step with an additive proposal (mh.py:56-193, gaussian.py:68-131).

    delta [T,W,L,D] is the proposal increment (scale*randn or multivariate_normal draw) for
    the active leaves and 0 elsewhere; factors are zero (gaussian.py:131)."""
    T, W, L, D = state.coords.shape
    q = state.coords.copy()
    q[state.inds] = (state.coords + delta)[state.inds]  # gaussian.py:99-108
    if periods is not None:
        periodic_wrap(q, periods)  # gaussian.py:111-129
    if gibbs_mask is not None:  # cleanup_proposals_gibbs (move.py:302-307): the other parameters keep their values
        q[:, :, ~gibbs_mask] = state.coords[:, :, ~gibbs_mask]
    logp = box_log_prior(prior, q, state.inds)
    logl = log_like(like, q, state.inds, logp)
    logP = tempered_log_posterior(logl, logp, betas)
    prev_logP = tempered_log_posterior(state.logl, state.logp, betas)
    lnpdiff = np.zeros((T, W)) + logP - prev_logP  # mh.py:168
    keep = lnpdiff > np.log(u_acc)  # mh.py:171
    sub = np.tile(np.arange(W), (T, 1))
    update_subset(state, sub, q, logl, logp, keep)
    return keep, dict(q=q, logl=logl, logp=logp, lnpdiff=lnpdiff)


def distgen_step(state, new_points, u_acc, betas, prior, like):
    """MH step that redraws every active leaf from `generate_dist` = the priors (mh.py:56-193 with
    distgen.py:34-104): factors = +log q(old) - log q(new).  new_points [T,W,L,D]: the draws (0 where inactive)."""
    T, W, L, D = state.coords.shape
    q = state.coords.copy()
    factors = np.zeros((T, W))
    old = state.coords[state.inds]
    tw = np.where(state.inds)[:2]
    factors[tw] += +1 * prior.logpdf(old)  # distgen.py:96
    q[state.inds] = new_points[state.inds]
    factors[tw] += -1 * prior.logpdf(q[state.inds])  # distgen.py:102
    logp = box_log_prior(prior, q, state.inds)
    logl = log_like(like, q, state.inds, logp)
    logP = tempered_log_posterior(logl, logp, betas)
    prev_logP = tempered_log_posterior(state.logl, state.logp, betas)
    lnpdiff = factors + logP - prev_logP  # mh.py:168
    keep = lnpdiff > np.log(u_acc)  # mh.py:171
    sub = np.tile(np.arange(W), (T, 1))
    update_subset(state, sub, q, logl, logp, keep)
    return keep, dict(q=q, logl=logl, logp=logp, lnpdiff=lnpdiff)


def _mt_logsumexp(a):
    """multipletry.py:25-31: max, exp of the differences summed, log"""
    mx = np.max(a, axis=-1)
    return mx + np.log(np.exp(a - mx[:, None]).sum(axis=-1))


def mt_distgen_step(state, tries, u_sel, u_acc, betas, prior, like):
    """MTDistGenMove(generate_dist = priors, num_try, independent=True): multiple-try Metropolis
    (multipletry.py:238-514, mtdistgen.py:8-133) inside MHMove.propose (mh.py:56-193).

    tries [T,W,NT,D]: the draws from the priors (mtdistgen.py:58), u_sel [T,W]: the uniform that picks a try by
    importance weight (multipletry.py:51, GLOBAL stream), u_acc [T,W]: the Metropolis uniform (mh.py:171)."""
    T, W, L, D = state.coords.shape
    assert L == 1 and bool(np.all(state.inds)), "multiple try works on one (present) leaf per walker (multipletry.py:548)"
    NT = tries.shape[2]
    n = T * W
    pts = tries.reshape(n, NT, D)
    cur = state.coords.reshape(n, D)
    b = np.ones(T) if betas is None else betas
    beta_w = np.repeat(b, W)                                                  # multipletry.py:553-556
    lpp = prior.logpdf(pts.reshape(n * NT, D)).reshape(n, NT)                 # mtdistgen.py:62-64
    ll = like(pts.reshape(n * NT, D)).reshape(n, NT)                          # :334 (no prior gate: no logp is passed)
    ll[np.isnan(ll)] = FILL                                                   # :339-341
    lp = prior.logpdf(pts.reshape(n * NT, D)).reshape(n, NT)                  # :344
    logP = beta_w[:, None] * ll + lp                                          # :354, get_mt_log_posterior :205-236
    log_w = logP - lpp                                                        # get_mt_computations :34-55 (not symmetric)
    lsw = _mt_logsumexp(log_w)
    probs = np.exp(log_w - lsw[:, None])
    keep_idx = (probs.cumsum(1) > u_sel.reshape(n)[:, None]).argmax(1)
    it = (np.arange(n), keep_idx)
    lp_out, ll_out, logP_out = lp[it], ll[it], logP[it]
    q = pts[it].copy()
    lpp_out = lpp[it]
    # independent proposal: the auxiliary set repeats the tries, with the current point in place of the chosen one (:383-416)
    aux_ll, aux_lp, aux_lpp = ll.copy(), lp.copy(), lpp.copy()
    aux_ll[it] = state.logl.reshape(n)
    aux_lp[it] = state.logp.reshape(n)
    aux_lpp[it] = prior.logpdf(cur)
    aux_logP = beta_w[:, None] * aux_ll + aux_lp
    aux_lsw = _mt_logsumexp(aux_logP - aux_lpp)
    aux_logP_out, aux_lpp_out = aux_logP[it], aux_lpp[it]
    factors = ((aux_logP_out - aux_lsw) - aux_lpp_out + aux_lpp_out) - ((logP_out - lsw) - lpp_out + lpp_out)  # :467-471
    # mh.py:146-171 with the stored multiple-try values (mt_ll, mt_lp)
    logl, logp = ll_out.reshape(T, W), lp_out.reshape(T, W)
    logP2 = tempered_log_posterior(logl, logp, betas)
    prev_logP = tempered_log_posterior(state.logl, state.logp, betas)
    lnpdiff = factors.reshape(T, W) + logP2 - prev_logP
    keep = lnpdiff > np.log(u_acc)
    sub = np.tile(np.arange(W), (T, 1))
    update_subset(state, sub, q.reshape(T, W, 1, D), logl, logp, keep)
    return keep, dict(q=q, factors=factors, chosen=keep_idx.reshape(T, W), lnpdiff=lnpdiff)


# ----------------------------------------------------------------------------------
# parallel tempering
# ----------------------------------------------------------------------------------
def temperature_swaps(state, betas, iperms, i1perms, us):
    """tempering.py:484-561: sequential ladder hot -> cold.

    iperms/i1perms/us are indexed by rung i (entry 0 unused).  Returns swaps_accepted [T-1]."""
    T, W = state.logl.shape
    swaps_accepted = np.empty(T - 1)
    for i in range(T - 1, 0, -1):
        dbeta = betas[i - 1] - betas[i]  # :518-522
        iperm, i1perm = iperms[i], i1perms[i]
        raccept = np.log(us[i])  # :535
        paccept = dbeta * (state.logl[i, iperm] - state.logl[i - 1, i1perm])  # :538
        sel = paccept > raccept
        swaps_accepted[i - 1] = np.sum(sel)
        a, b = iperm[sel], i1perm[sel]
        for arr in (state.coords, state.inds, state.logl, state.logp):  # :351-482
            tmp = arr[i, a].copy()
            arr[i, a] = arr[i - 1, b]
            arr[i - 1, b] = tmp
    return swaps_accepted


def resolve_ladder(logl, betas, iperms, i1perms, us):
    """The swap pass needs logl only (tempering.py:538): run it on slot ids.  Returns (src [T,W] flat slot
    t*W+w whose walker ends at each slot, swaps_accepted [T-1])."""
    T, W = logl.shape
    ids = OState(np.arange(T * W, dtype=np.float64).reshape(T, W, 1, 1), logl=logl, logp=np.zeros((T, W)))
    sw = temperature_swaps(ids, betas, iperms, i1perms, us)
    return ids.coords[:, :, 0, 0].astype(np.int64), sw


def adapt_temps(betas, swaps_accepted, nwalkers, time, adaptation_lag=10000, adaptation_time=100):
    """tempering.py:563-596: returns the new ladder (betas0 + (new - betas0), as the reference)."""
    ratios = swaps_accepted / np.full(len(swaps_accepted), nwalkers)  # :587, :282
    b = betas.copy()
    decay = adaptation_lag / (time + adaptation_lag)
    kappa = decay / adaptation_time
    dSs = kappa * (ratios[:-1] - ratios[1:])
    deltaTs = np.diff(1 / b[:-1])
    deltaTs *= np.exp(dSs)
    b[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / b[0])
    return betas + (b - betas)  # :583, :593


def make_ladder_default(ndim, ntemps):
    """tempering.py:10-197 for the (ntemps given, Tmax None) branch used by the configs."""
    tstep_tab = _TSTEP
    tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim) if ndim > len(tstep_tab) else tstep_tab[ndim - 1]
    Tmax = tstep ** (ntemps - 1)
    return np.logspace(0, -np.log10(Tmax), ntemps)


# ----------------------------------------------------------------------------------
# random sources
# ----------------------------------------------------------------------------------
class NumpyStreams:
    """The reference's two MT19937 streams in the reference's call order."""

    mode = "numpy"

    def __init__(self, private, glob):
        self.private = private  # ensemble.py:651-652
        self.glob = glob  # module-level np.random (red_blue.py:124, tempering.py:526-535)

    def move_choice(self, it, weights):
        return int(self.private.choice(len(weights), p=weights))  # ensemble.py:971

    def split_lists(self, it, T, W, randomize=True):
        labels = np.tile(np.arange(W), (T, 1)) % 2  # red_blue.py:121-122
        if randomize:  # red_blue.py:123 (StretchMove(randomize_split=...))
            for row in labels:
                self.glob.shuffle(row)  # :124
        ids = np.tile(np.arange(W), (T, 1))
        return [ids[labels == s].reshape(T, -1) for s in (0, 1)]  # :150-154 ascending ids

    def stretch(self, it, split, T, Ns, Nc, sub, gidx=0):
        rint = self.private.randint(Nc, size=(T, Ns))  # stretch.py:93
        u_z = self.private.rand(T, Ns)  # stretch.py:131
        u_acc = self.private.rand(T, Ns)  # red_blue.py:294
        return rint, u_z, u_acc

    def gauss_increment(self, it, inds, D, proposal, gidx=0):
        n = int(inds.sum())
        f = 1.0
        if proposal.get("factor") is not None:  # gaussian.py:161-164: ONE scale factor per call, drawn first
            lf = np.log(proposal["factor"])
            f = np.exp(self.private.uniform(-lf, lf))
        if proposal["kind"] == "scalar":  # gaussian.py:166-167
            d = f * proposal["scale"] * self.private.randn(n, D)
        else:  # gaussian.py:192-195
            d = f * self.private.multivariate_normal(np.zeros(D), proposal["cov"], size=n)
        mode = proposal.get("mode", "vector")
        if mode == "random":  # gaussian.py:172-173: one random dimension per walker
            m = self.private.randint(D, size=n)
            keep = np.zeros((n, D), dtype=bool)
            keep[np.arange(n), m] = True
            d = np.where(keep, d, 0.0)
        elif mode == "sequential":  # gaussian.py:174-176: the same, next dimension for everybody
            idx = proposal.get("_index", 0)
            keep = np.zeros((n, D), dtype=bool)
            keep[:, idx % D] = True
            d = np.where(keep, d, 0.0)
            proposal["_index"] = (idx + 1) % D
        delta = np.zeros(inds.shape + (D,))
        delta[inds] = d
        return delta

    def prior_draws(self, it, inds, prior):
        """distgen.py:99: generate_dist.rvs(size=n) -> one global rand(n) per parameter (prior.py:56-71, :432-497)"""
        out = np.zeros(inds.shape + (len(prior.lo),))
        out[inds] = prior.rvs(int(inds.sum()), self.glob)
        return out

    def mt_draws(self, it, T, W, NT, prior):
        """mtdistgen.py:58: generate_dist.rvs(size=(n, num_try)) — one global rand(n, num_try) per parameter — then the
        global rand(n) that picks the try (multipletry.py:51)"""
        tries = prior.rvs((T * W, NT), self.glob).reshape(T, W, NT, len(prior.lo))
        return tries, self.glob.rand(T * W).reshape(T, W)

    def accept_uniforms(self, it, slot, T, W, gidx=0):
        return self.private.rand(T, W)  # mh.py:171

    def swap_draws(self, it, T, W, permute=True):
        iperms, i1perms, us = [None] * T, [None] * T, [None] * T
        for i in range(T - 1, 0, -1):
            if permute:
                iperms[i] = self.glob.permutation(W)  # tempering.py:526-527
                i1perms[i] = self.glob.permutation(W)
            else:
                iperms[i] = np.arange(W)
                i1perms[i] = np.arange(W)
            us[i] = self.glob.uniform(size=W)  # :535
        return iperms, i1perms, us


class PhiloxStreams:
    """Production-mode source of the CUDA path (counter based; see oracle/philox_np.py)."""

    mode = "philox"

    def __init__(self, seed, schedule_random=None, t0=0):
        self.seed = int(seed)
        self.sched = schedule_random if schedule_random is not None else np.random.RandomState(self.seed & 0x7FFFFFFF)
        self.t0 = int(t0)  # global index of local temperature 0 (temperature-sharded runs key the streams globally)

    def move_choice(self, it, weights):
        return int(self.sched.choice(len(weights), p=weights))  # host-side schedule, as the reference

    def split_lists(self, it, T, W, randomize=True):
        n0 = (W + 1) // 2
        n1 = W // 2
        subs0 = np.empty((T, n0), dtype=np.int64)
        subs1 = np.empty((T, n1), dtype=np.int64)
        for t in range(T):
            # randomize_split=False: the identity instead of the keyed bijection (csrc/k_stretch.cu:stretch_draw)
            sig = px.split_perm(it, self.seed, self.t0 + t, W) if randomize else np.arange(W)
            subs0[t] = sig[0::2][:n0]
            subs1[t] = sig[1::2][:n1]
        return [subs0, subs1]

    def stretch(self, it, split, T, Ns, Nc, sub, gidx=0):
        # one Philox block per walker, keyed by (position in the split permutation, temperature | Gibbs split << 16)
        pos = (2 * np.arange(Ns) + split)[None, :].repeat(T, axis=0)
        t = self.t0 + np.arange(T)[:, None].repeat(Ns, axis=1)
        return px.stretch_draws(it, self.seed, t | (int(gidx) << 16), pos, Nc)

    def accept_for(self, it, slot, flat_walker):
        return px.accept_draws(it, self.seed, flat_walker, slot)

    def gauss_increment(self, it, inds, D, proposal, gidx=0):
        T, W, L = inds.shape
        flat = np.arange(T * W * L, dtype=np.uint32) + np.uint32(self.t0 * W * L)
        z = px.gauss_draws(it, self.seed, flat, D, gidx=gidx)
        f = 1.0
        if proposal.get("factor") is not None:  # one factor per call: TAG_GAUSS block (0xFFFFFFFF, 0xFFFF | gidx << 16)
            lf = np.log(proposal["factor"])
            r = px._stream(px.TAG_GAUSS, it, self.seed, np.uint32(0xFFFFFFFF), np.uint32(0xFFFF | (int(gidx) << 16)))
            f = np.exp(-lf + (lf - (-lf)) * float(px.u01_52(r[0], r[1])))
        if proposal["kind"] == "scalar":
            d = (f * proposal["scale"]) * z
        else:
            d = f * (z @ proposal["chol"].T)
        mode = proposal.get("mode", "vector")
        if mode == "random":  # dimension = high word of (third word of the walker's accept block) x D
            ra = px._stream(px.TAG_ACCEPT, it, self.seed, flat, np.uint32(int(gidx) << 16))
            m = ((ra[2].astype(np.uint64) * np.uint64(D)) >> np.uint64(32)).astype(np.int64)
            keep = np.zeros(d.shape, dtype=bool)
            keep[np.arange(d.shape[0]), m] = True
            d = np.where(keep, d, 0.0)
        elif mode == "sequential":
            idx = proposal.get("_index", 0)
            keep = np.zeros(d.shape, dtype=bool)
            keep[:, idx % D] = True
            d = np.where(keep, d, 0.0)
            proposal["_index"] = (idx + 1) % D
        delta = d.reshape(T, W, L, D)
        delta[~inds] = 0.0
        return delta

    def prior_draws(self, it, inds, prior):
        """uniform u of (flat leaf, parameter d): TAG_GAUSS block (flat, d // 2), word pair d % 2 (k_gauss.cu)"""
        T, W, L = inds.shape
        D = len(prior.lo)
        flat = np.arange(T * W * L, dtype=np.uint32) + np.uint32(self.t0 * W * L)
        out = np.zeros((T * W * L, D))
        for d in range(D):
            r = px._stream(px.TAG_GAUSS, it, self.seed, flat, np.uint32(d // 2))
            u = px.u01_52(r[2], r[3]) if d % 2 else px.u01_52(r[0], r[1])
            out[:, d] = u * (prior.hi[d] - prior.lo[d]) + prior.lo[d]
        out = out.reshape(T, W, L, D)
        out[~inds] = 0.0
        return out

    def mt_draws(self, it, T, W, NT, prior):
        """try j, parameter d of flat walker f: TAG_MT block (f, j * 16 + d // 2), word pair d % 2; the selection uniform:
        block (f, 0xFFFF), first pair (csrc/k_mt.cu)"""
        D = len(prior.lo)
        flat = np.arange(T * W, dtype=np.uint32) + np.uint32(self.t0 * W)
        tries = np.zeros((T * W, NT, D))
        for j in range(NT):
            for d in range(D):
                r = px._stream(px.TAG_MT, it, self.seed, flat, np.uint32(j * 16 + d // 2))
                u = px.u01_52(r[2], r[3]) if d % 2 else px.u01_52(r[0], r[1])
                tries[:, j, d] = u * (prior.hi[d] - prior.lo[d]) + prior.lo[d]
        r = px._stream(px.TAG_MT, it, self.seed, flat, np.uint32(0xFFFF))
        return tries.reshape(T, W, NT, D), px.u01_52(r[0], r[1]).reshape(T, W)

    def accept_uniforms(self, it, slot, T, W, gidx=0):
        flat = np.arange(T * W, dtype=np.uint32) + np.uint32(self.t0 * W)
        return px.accept_draws(it, self.seed, flat, slot | (int(gidx) << 16)).reshape(T, W)

    def swap_draws(self, it, T, W, permute=True):
        iperms, i1perms, us = [None] * T, [None] * T, [None] * T
        # one keyed bijection sigma_r per rung; pair k of rung i is (i, sigma_i(k)) with (i-1, sigma_{i-1}(k)),
        # i.e. iperm = sigma_i, i1perm = sigma_{i-1} in tempering.py:526-527; uniform k belongs to pair k
        sig = [px.swap_perm(it, self.seed, r, W) if permute else np.arange(W) for r in range(T)]
        for i in range(T - 1, 0, -1):
            iperms[i] = sig[i]
            i1perms[i] = sig[i - 1]
            us[i] = px.swap_uniforms(it, self.seed, i, W)
        return iperms, i1perms, us


# ----------------------------------------------------------------------------------
# driver
# ----------------------------------------------------------------------------------
class OracleSampler:
    """ensemble.py:965-984 + the per-move tails (red_blue.py:326-331, mh.py:186-191).

    moves: list of dicts {"kind": "stretch", "a": 2.0} or
           {"kind": "gaussian", "proposal": {"kind": "scalar", "scale": s}}.
    betas None means no TemperatureControl (tempering_kwargs == {})."""

    def __init__(self, prior, like, moves, weights, streams, betas=None, adaptive=True,
                 adaptation_lag=10000, adaptation_time=100, stop_adaptation=-1, permute=True, periods=None):
        self.periods = None if periods is None else np.asarray(periods, dtype=np.float64)
        self.prior, self.like = prior, like
        import copy
        self.moves = copy.deepcopy(moves)  # a sequential-mode Gaussian proposal keeps its dimension counter in its dict
        w = np.atleast_1d(np.asarray(weights, dtype=float))
        self.weights = w / w.sum()  # ensemble.py:376-377
        self.streams = streams
        self.betas = None if betas is None else np.asarray(betas, dtype=np.float64).copy()
        self.adaptive, self.lag, self.t0 = adaptive, adaptation_lag, adaptation_time
        self.stop_adaptation, self.permute = stop_adaptation, permute
        self.time = 0
        self.iteration = 0
        self.swaps_accepted = None
        self.last_move = None

    def initialise(self, state):
        if state.logp is None:
            state.logp = box_log_prior(self.prior, state.coords, state.inds)  # ensemble.py:898-901
        if state.logl is None:
            state.logl = log_like(self.like, state.coords, state.inds, state.logp)  # :903-912
        return state

    def iterate(self, state):
        """one sampler iteration (ensemble.py:965-985): choose a move, propose; a {"kind": "combine", "moves": [...]}
        move runs its sub-moves in order, each with its own tempering tail (combine.py:99-135), and returns the accept
        COUNT of every walker"""
        mi = self.streams.move_choice(self.iteration, self.weights)
        move = self.moves[mi]
        self.last_move = mi
        if move["kind"] == "combine":
            accepted = None
            for sub in move["moves"]:
                acc = self._propose(state, sub)
                accepted = acc.astype(np.int64) if accepted is None else accepted + acc
            return accepted
        return self._propose(state, move)

    def _propose(self, state, move):
        """move.propose(model, state): proposal + Metropolis step + temper_comps; ticks the iteration counter that keys
        the counter-based streams (the device ticks it in the swap pass)"""
        T, W, L, D = state.coords.shape
        it = self.iteration
        st = self.streams
        accepted = np.zeros((T, W), dtype=bool)
        # Gibbs splits (move.py:113-402): {"gibbs": [mask [L, D] bool | None, ...]} — every split has its own draws,
        # Metropolis test and update inside this one propose call; the tempering tail runs once, after the last split
        gibbs = move.get("gibbs") or [None]
        self.last_nsplits = len(gibbs)
        self.last_accept_sum = np.zeros((T, W), dtype=np.int64)  # what the reference adds to move.accepted in this call
        if move["kind"] == "stretch":
            # red_blue.py:118-124: one red/blue labelling for all Gibbs splits
            lists = st.split_lists(it, T, W, randomize=move.get("randomize_split", True))
            for gi, gmask in enumerate(gibbs):
                if gmask is not None and not np.any(state.inds[:, :, gmask.any(axis=-1)]):
                    continue  # setup_proposals: no leaf to propose for (red_blue.py:142-143)
                for split in (0, 1):
                    sub, comp = lists[split], lists[1 - split]
                    Ns, Nc = sub.shape[1], comp.shape[1]
                    rint, u_z, u_acc = st.stretch(it, split, T, Ns, Nc, sub, gidx=gi)
                    if u_acc is None:
                        flat = (np.arange(T)[:, None] * W + sub).astype(np.uint32)
                        u_acc = st.accept_for(it, split, flat)
                    keep, _ = stretch_half_step(state, sub, comp, rint, u_z, u_acc, move.get("a", 2.0),
                                                self.betas, self.prior, self.like, self.periods, gibbs_mask=gmask)
                    # red_blue.py:296-309: `accepted` ORs the halves AND the Gibbs splits of this call
                    cur = np.take_along_axis(accepted, sub, axis=1)
                    np.put_along_axis(accepted, sub, cur | keep, axis=1)
                self.last_accept_sum += accepted  # red_blue.py:325: self.accepted += accepted (the running OR), per split
        elif move["kind"] == "gaussian":
            for gi, gmask in enumerate(gibbs):
                inds_go = state.inds if gmask is None else state.inds & gmask.any(axis=-1)[None, None, :]
                if not np.any(inds_go):
                    continue
                delta = st.gauss_increment(it, inds_go, D, move["proposal"], gidx=gi)
                u_acc = st.accept_uniforms(it, 0, T, W, gidx=gi)
                keep, _ = gaussian_step(state, delta, u_acc, self.betas, self.prior, self.like, self.periods,
                                        gibbs_mask=gmask)
                accepted = keep  # mh.py:171: the mask of the LAST split is what propose returns
                self.last_accept_sum += keep  # mh.py:187
        elif move["kind"] == "mt":  # MTDistGenMove(priors, num_try, independent=True)
            tries, u_sel = st.mt_draws(it, T, W, int(move["num_try"]), self.prior)
            u_acc = st.accept_uniforms(it, 0, T, W)
            keep, _ = mt_distgen_step(state, tries, u_sel, u_acc, self.betas, self.prior, self.like)
            accepted = keep
        elif move["kind"] == "distgen":
            new_points = st.prior_draws(it, state.inds, self.prior)
            u_acc = st.accept_uniforms(it, 0, T, W)
            keep, _ = distgen_step(state, new_points, u_acc, self.betas, self.prior, self.like)
            accepted = keep
        else:
            raise ValueError(move["kind"])
        if self.betas is not None:  # temper_comps, tempering.py:598-649
            iperms, i1perms, us = st.swap_draws(it, T, W, self.permute)
            self.swaps_accepted = temperature_swaps(state, self.betas, iperms, i1perms, us)
            if self.adaptive and T > 1:  # :632-633, :585-596
                if self.stop_adaptation < 0 or self.time < self.stop_adaptation:
                    self.betas = adapt_temps(self.betas, self.swaps_accepted, W, self.time, self.lag, self.t0)
                self.time += 1
        self.iteration += 1
        return accepted


_TSTEP = np.array([
    25.2741, 7.0, 4.47502, 3.5236, 3.0232, 2.71225, 2.49879, 2.34226, 2.22198, 2.12628,
    2.04807, 1.98276, 1.92728, 1.87946, 1.83774, 1.80096, 1.76826, 1.73895, 1.7125, 1.68849,
    1.66657, 1.64647, 1.62795, 1.61083, 1.59494, 1.58014, 1.56632, 1.55338, 1.54123, 1.5298,
    1.51901, 1.50881, 1.49916, 1.49, 1.4813, 1.47302, 1.46512, 1.45759, 1.45039, 1.4435,
    1.4369, 1.43056, 1.42448, 1.41864, 1.41302, 1.40761, 1.40239, 1.39736, 1.3925, 1.38781,
    1.38327, 1.37888, 1.37463, 1.37051, 1.36652, 1.36265, 1.35889, 1.35524, 1.3517, 1.34825,
    1.3449, 1.34164, 1.33847, 1.33538, 1.33236, 1.32943, 1.32656, 1.32377, 1.32104, 1.31838,
    1.31578, 1.31325, 1.31076, 1.30834, 1.30596, 1.30364, 1.30137, 1.29915, 1.29697, 1.29484,
    1.29275, 1.29071, 1.2887, 1.28673, 1.2848, 1.28291, 1.28106, 1.27923, 1.27745, 1.27569,
    1.27397, 1.27227, 1.27061, 1.26898, 1.26737, 1.26579, 1.26424, 1.26271, 1.26121, 1.25973,
])
