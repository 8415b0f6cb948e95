"""ORACLE (test infrastructure only) — CPU restatement of the reversible-jump / group-stretch part of Eryn's hot path
(BASELINE config 5).  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

What is restated (file:line into /root/reference/src/eryn):
  mb_log_prior        ensemble.py:1192-1212 (per branch logpdf of every leaf, inactive leaves -> 0, leaves summed,
                      branches added in order) with prior.py:80-88, :337-392
  mb_log_like         ensemble.py:1219-1545, non-vectorised route: walkers with logp = -inf or without any active leaf are
                      not evaluated (-1e300, :1279-1282, :1486-1513); the others get the user likelihood on their active
                      leaves in leaf order — here the reference test's log_like_fn_gauss_and_sine (tests/test_eryn.py:38-92)
  group_stretch_step  group.py:122-281 (GroupMove.propose) + groupstretch.py:34-120 + stretch.py:103-158
  friends rule        the reference test fixture MeanGaussianGroupMove (tests/test_eryn.py:813-907), any branch
  rj_step             rj.py:145-388 (ReversibleJumpMove.propose: edge factors :228-271, accept :330-332, swaps without
                      adaptation :381-382) + distgenrj.py:35-222 (DistributionGenerateRJ)
  swaps               tempering.py:484-561 on every array of the state, branch supplementals included (:351-482)

Pinned by tests/golden/c5_*.npz, recorded from the unmodified reference by tests/golden/make_golden_rj.py.
"""
import numpy as np

from . import eryn_oracle as orc
from . import philox_np as px

FILL = -1e300
TAG_GROUP = 8


# ----------------------------------------------------------------------------------
# likelihood of the reference test: sum of Gaussian pulses + sines against a data vector
# ----------------------------------------------------------------------------------
class PulseLike:
    """kinds[b] = 0: a*exp(-(t-b)^2/(2c^2)), 1: a*sin(2 pi b t + c); log L = -1/2 sum(((template - y)/sigma)^2)."""

    def __init__(self, t, y, sigma, kinds):
        self.t, self.y, self.sigma, self.kinds = np.asarray(t, float), np.asarray(y, float), float(sigma), list(kinds)

    def __call__(self, params_by_branch):
        template = np.zeros_like(self.t)
        for kind, params in zip(self.kinds, params_by_branch):
            if params is None:
                continue
            for a, b, c in params:
                if kind == 0:
                    template += a * np.exp(-((self.t - b) ** 2) / (2 * c**2))
                else:
                    template += a * np.sin(2 * np.pi * b * self.t + c)
        return -0.5 * np.sum(((template - self.y) / self.sigma) ** 2, axis=-1)


class MBState:
    """coords[b] [T,W,L_b,D_b], inds[b] [T,W,L_b] bool, logl/logp [T,W], closest[b] [T,W,L_b,nfriends] int (or None)."""

    def __init__(self, coords, inds, logl=None, logp=None, closest=None):
        self.coords = [np.array(c, dtype=np.float64) for c in coords]
        self.inds = [np.array(i, dtype=bool) for i in inds]
        self.logl = None if logl is None else np.array(logl, dtype=np.float64)
        self.logp = None if logp is None else np.array(logp, dtype=np.float64)
        self.closest = closest

    @property
    def shape(self):
        return self.coords[0].shape[:2]


def mb_log_prior(priors, coords, inds):
    T, W = coords[0].shape[:2]
    out = np.zeros((T, W))
    for pr, c, i in zip(priors, coords, inds):
        L, D = c.shape[2:]
        v = pr.logpdf(c.reshape(-1, D)).reshape(T, W, L)
        v[~i] = 0.0
        out += v.sum(axis=-1)
    return out


def mb_log_like(like, coords, inds, logp):
    T, W = logp.shape
    if np.all(np.isinf(logp)):
        return np.full_like(logp, FILL)
    ll = np.full((T, W), FILL)
    for t in range(T):
        for w in range(W):
            if np.isinf(logp[t, w]):
                continue
            if not any(i[t, w].any() for i in inds):
                continue  # no leaf at all: not a group, fill_zero_leaves_val (ensemble.py:1486-1513)
            ll[t, w] = like([c[t, w][i[t, w]] if i[t, w].any() else None for c, i in zip(coords, inds)])
    return ll


# ----------------------------------------------------------------------------------
# friends (tests/test_eryn.py:813-907)
# ----------------------------------------------------------------------------------
class Friends:
    def __init__(self, nfriends, key_index=1):
        self.nfriends, self.key = int(nfriends), int(key_index)
        self.friends, self.means = [], []

    def _closest(self, b, vals):
        dist = np.abs(vals[:, None] - self.means[b][None, :])
        return np.argsort(dist, axis=1)[:, : self.nfriends]

    def setup(self, state):
        nb = len(state.coords)
        self.friends, self.means = [None] * nb, [None] * nb
        if state.closest is None:
            state.closest = [np.zeros(i.shape + (self.nfriends,), dtype=np.int64) for i in state.inds]
        for b in range(nb):
            c, i = state.coords[b], state.inds[b]
            fr = c[0, i[0]]
            means, uni = np.unique(fr[:, self.key].copy(), return_index=True)
            self.friends[b], self.means[b] = fr[uni], means  # np.unique returns them sorted already
            state.closest[b][i] = self._closest(b, c[i, self.key])
            state.closest[b][~i] = -1

    def fix(self, state):
        for b in range(len(state.coords)):
            i = state.inds[b]
            fix = i & np.all(state.closest[b] == -1, axis=-1)
            if fix.any():
                state.closest[b][fix] = self._closest(b, state.coords[b][fix, self.key])


# ----------------------------------------------------------------------------------
# random sources
# ----------------------------------------------------------------------------------
class NumpyStreamsMB(orc.NumpyStreams):
    """the reference's draw order for the config-5 moves (SURVEY.md App. B + the fixture's global randint)"""

    def group_draws(self, it, state, nfriends):
        T, W = state.shape
        picks, u_z = [], None
        for b, i in enumerate(state.inds):
            n = int(i.sum())
            r = self.glob.randint(nfriends, size=(n,))  # fixture find_friends (global stream)
            p = np.zeros(i.shape, dtype=np.int64)
            p[i] = r
            picks.append(p)
            if b == 0:
                u_z = self.private.rand(T, W)  # stretch.py:131, first branch only
        return picks, u_z, None  # u_acc is drawn after the likelihood: accept_uniforms()

    def rj_draws(self, it, inds, nmin, nmax, priors, branches=None, gidx=0):
        """distgenrj.py:35-122 for every branch of the Gibbs split (`branches`, default all), then the births from the
        priors (prior.py:56-71, global stream)"""
        T, W = inds[0].shape[:2]
        changes = []
        for b, i in enumerate(inds):
            if nmin[b] == nmax[b] or (branches is not None and b not in branches):
                changes.append(None)
                continue
            nleaves = i.sum(axis=-1)
            change = self.private.choice([-1, +1], size=nleaves.shape)
            change = change * ((nleaves != nmin[b]) & (nleaves != nmax[b])) + (+1) * (nleaves == nmin[b]) \
                + (-1) * (nleaves == nmax[b])
            leaf = np.full((T, W), -1, dtype=np.int64)
            for t in range(T):
                for w in range(W):
                    if change[t, w] == +1:
                        leaf[t, w] = self.private.choice(np.where(~i[t, w])[0])
                    elif change[t, w] == -1:
                        leaf[t, w] = self.private.choice(np.where(i[t, w])[0])
            changes.append((change, leaf))
        births = []
        for b, ch in enumerate(changes):
            if ch is None:
                births.append(None)
                continue
            n = int((ch[0] == +1).sum())
            D = len(priors[b].lo)
            vals = np.zeros((n, D))
            for d in range(D):
                vals[:, d] = self.glob.rand(n) * (priors[b].hi[d] - priors[b].lo[d]) + priors[b].lo[d]
            full = np.zeros((T, W, D))
            full[ch[0] == +1] = vals  # row-major (t, w) order = order of inds_for_change["+1"]
            births.append(full)
        return changes, births


class PhiloxStreamsMB(orc.PhiloxStreams):
    """production-mode draws of the config-5 kernels (csrc/k_rj.cu); walker id fw = (t0 + t) * W + w"""

    def _fw(self, T, W):
        return (np.arange(T * W, dtype=np.uint32) + np.uint32(self.t0 * W)).reshape(T, W)

    def group_draws(self, it, state, nfriends):
        T, W = state.shape
        fw = self._fw(T, W)
        r0, r1, r2, r3 = px._stream(TAG_GROUP, it, self.seed, fw, np.uint32(0))
        u_z, u_acc = px.u01_52(r0, r1), px.u01_52(r2, r3)
        picks, j0 = [], 0
        for i in state.inds:
            L = i.shape[2]
            p = np.zeros(i.shape, dtype=np.int64)
            for l in range(L):
                j = j0 + l
                w = px._stream(TAG_GROUP, it, self.seed, fw, np.uint32(1 + j // 4))[j % 4]
                p[:, :, l] = (w.astype(np.uint64) * np.uint64(nfriends)) >> np.uint64(32)
            picks.append(p)
            j0 += L
        return picks, u_z, u_acc

    def rj_draws(self, it, inds, nmin, nmax, priors, branches=None, gidx=0):
        T, W = inds[0].shape[:2]
        fw = self._fw(T, W)
        g = np.uint32(int(gidx) << 8)   # the Gibbs split of the propose call keys the streams (csrc/k_rj.cu)
        changes, births = [], []
        for b, i in enumerate(inds):
            if nmin[b] == nmax[b] or (branches is not None and b not in branches):
                changes.append(None)
                births.append(None)
                continue
            r0, r1, _, _ = px._stream(px.TAG_RJ, it, self.seed, fw, np.uint32(8 * b) | g)
            nleaves = i.sum(axis=-1)
            change = np.where(r0 & np.uint32(1), 1, -1)
            change = change * ((nleaves != nmin[b]) & (nleaves != nmax[b])) + (+1) * (nleaves == nmin[b]) \
                + (-1) * (nleaves == nmax[b])
            L = i.shape[2]
            ncand = np.where(change == 1, L - nleaves, nleaves)
            k = (r1.astype(np.uint64) * ncand.astype(np.uint64)) >> np.uint64(32)
            leaf = np.full((T, W), -1, dtype=np.int64)
            for t in range(T):
                for w in range(W):
                    cand = np.where(~i[t, w])[0] if change[t, w] == 1 else np.where(i[t, w])[0]
                    leaf[t, w] = cand[int(k[t, w])]
            D = len(priors[b].lo)
            full = np.zeros((T, W, D))
            for d in range(D):
                q = px._stream(px.TAG_RJ, it, self.seed, fw, np.uint32(8 * b + 1 + d // 2) | g)
                u = px.u01_52(q[2], q[3]) if d % 2 else px.u01_52(q[0], q[1])
                full[:, :, d] = u * (priors[b].hi[d] - priors[b].lo[d]) + priors[b].lo[d]
            full[change != 1] = 0.0
            changes.append((change, leaf))
            births.append(full)
        return changes, births

    def rj_accept(self, it, T, W, gidx=0):
        _, _, r2, r3 = px._stream(px.TAG_RJ, it, self.seed, self._fw(T, W), np.uint32(int(gidx) << 8))
        return px.u01_52(r2, r3)


# ----------------------------------------------------------------------------------
# the sampler
# ----------------------------------------------------------------------------------
def swap_everything(state, betas, iperms, i1perms, us):
    """tempering.py:484-561 on all arrays of the state (coords, inds, branch supplementals, logl, logp)."""
    T, W = state.logl.shape
    arrays = list(state.coords) + list(state.inds) + ([] if state.closest is None else list(state.closest)) \
        + [state.logl, state.logp]
    swaps = np.empty(T - 1)
    for i in range(T - 1, 0, -1):
        dbeta = betas[i - 1] - betas[i]
        iperm, i1perm = iperms[i], i1perms[i]
        sel = dbeta * (state.logl[i, iperm] - state.logl[i - 1, i1perm]) > np.log(us[i])
        swaps[i - 1] = np.sum(sel)
        a, b = iperm[sel], i1perm[sel]
        for arr in arrays:
            tmp = arr[i, a].copy()
            arr[i, a] = arr[i - 1, b]
            arr[i - 1, b] = tmp
    return swaps


class OracleSamplerMB:
    """One iteration = GroupStretch move (+ swaps + adaptation) then the RJ move (+ swaps, no adaptation):
    ensemble.py:965-1006 with group.py:122-281 and rj.py:145-388."""

    def __init__(self, priors, like, nmin, nmax, streams, betas, nfriends, n_iter_update, a=2.0, key_index=1,
                 rj_mode="together"):
        # rj_mode (ensemble.py:410-470): "together" = one RJ move over all branches; "iterate_branches" = one RJ move whose
        # Gibbs splits are the branches, one after the other; "separate_branches" = one RJ move per branch, one of them
        # chosen per iteration
        self.rj_mode = rj_mode
        self.last_rj_move = 0
        self.priors, self.like, self.nmin, self.nmax = priors, like, list(nmin), list(nmax)
        self.streams, self.betas = streams, np.asarray(betas, dtype=np.float64).copy()
        self.a, self.n_iter_update = a, n_iter_update
        self.friends = Friends(nfriends, key_index)
        self.iter = 0      # GroupMove.iter
        self.iteration = 0  # stream position
        self.time = 0
        self.swaps_accepted = None

    def _post(self, logl, logp):
        return orc.tempered_log_posterior(logl, logp, self.betas)

    def _accept(self, state, q, new_inds, factors, u_acc, branches_run=None):
        logp = mb_log_prior(self.priors, q, new_inds)
        if branches_run is not None:  # fix_logp_gibbs, move.py:369-402
            here = sum(new_inds[b].sum(axis=-1) for b in branches_run)
            total = sum(i.sum(axis=-1) for i in new_inds)
            logp[(total != 0) & (here == 0)] = -np.inf  # no use in running because no change
            logp[(total == 0) & (here == 0)] = 0.0      # there is nothing in the model currently
        logl = mb_log_like(self.like, q, new_inds, logp)
        lnpdiff = factors + self._post(logl, logp) - self._post(state.logl, state.logp)
        keep = lnpdiff > np.log(u_acc)
        # Move.update, move.py:472-703
        state.logl = logl * keep + state.logl * (~keep)
        nlp = logp.copy()
        nlp[np.isinf(nlp)] = 0.0
        state.logp = nlp * keep + state.logp * (~keep)
        for b in range(len(q)):
            state.inds[b] = new_inds[b] * keep[:, :, None] + state.inds[b] * (~keep[:, :, None])
            state.coords[b][keep] = q[b][keep]
        return keep

    def _swaps(self, state, it, adapt):
        T, W = state.shape
        if T < 2:
            return
        iperms, i1perms, us = self.streams.swap_draws(it, T, W, True)
        self.swaps_accepted = swap_everything(state, self.betas, iperms, i1perms, us)
        if adapt:
            self.betas = orc.adapt_temps(self.betas, self.swaps_accepted, W, self.time)
            self.time += 1

    def group_step(self, state):
        it = self.iteration
        st = self.streams
        T, W = state.shape
        if hasattr(st, "private"):
            st.private.choice(1, p=[1.0])  # ensemble.py:971, one in-model move
        if self.iter == 0 or self.iter % self.n_iter_update == 0:
            self.friends.setup(state)
        elif self.iter != 0:
            self.friends.fix(state)
        picks, u_z, u_acc = st.group_draws(it, state, self.friends.nfriends)
        zz = ((self.a - 1.0) * u_z + 1) ** 2.0 / self.a
        q, ndim = [], 0
        for b, (c, i) in enumerate(zip(state.coords, state.inds)):
            fr = np.zeros_like(c)
            idx = np.take_along_axis(state.closest[b], picks[b][..., None], axis=-1)[..., 0]
            fr[i] = self.friends.friends[b][idx[i]]
            q.append(fr - (fr - c) * zz[:, :, None, None])
            ndim += c.shape[2] * c.shape[3]
        factors = (ndim - 1.0) * np.log(zz)
        if u_acc is None:
            u_acc = st.accept_uniforms(it, 0, T, W)
        keep = self._accept(state, q, state.inds, factors, u_acc)
        self._swaps(state, it, adapt=True)
        self.iter += 1
        return keep

    def rj_step(self, state):
        it = self.iteration
        st = self.streams
        T, W = state.shape
        nb = len(state.coords)
        if self.rj_mode == "separate_branches":  # ensemble.py:990: one of the per-branch moves
            if hasattr(st, "private"):
                self.last_rj_move = int(st.private.choice(nb, p=np.ones(nb) / nb))
            else:
                self.last_rj_move = int(st.sched.choice(nb, p=np.ones(nb) / nb))
            splits = [[self.last_rj_move]]
        else:
            if hasattr(st, "private"):
                st.private.choice(1, p=[1.0])  # ensemble.py:990
            splits = [None] if self.rj_mode == "together" else [[b] for b in range(nb)]
        keep = None
        for gi, branches in enumerate(splits):  # rj.py:168-343, one pass per Gibbs split
            changes, births = st.rj_draws(it, state.inds, self.nmin, self.nmax, self.priors, branches=branches, gidx=gi)
            q = [c.copy() for c in state.coords]
            new_inds = [i.copy() for i in state.inds]
            factors = np.zeros((T, W))
            tt, ww = np.meshgrid(np.arange(T), np.arange(W), indexing="ij")
            for b, ch in enumerate(changes):
                if ch is None:
                    continue
                change, leaf = ch
                dm = change == -1
                new_inds[b][tt[dm], ww[dm], leaf[dm]] = False
                factors[dm] += +1 * self.priors[b].logpdf(q[b][tt[dm], ww[dm], leaf[dm]])
                bm = change == +1
                new_inds[b][tt[bm], ww[bm], leaf[bm]] = True
                q[b][tt[bm], ww[bm], leaf[bm]] = births[b][bm]
                factors[bm] += -1 * self.priors[b].logpdf(q[b][tt[bm], ww[bm], leaf[bm]])
            edge = np.zeros((T, W))
            for b in range(len(q)):
                if branches is not None and b not in branches:  # rj.py:236-237
                    continue
                if self.nmin[b] == self.nmax[b] or self.nmin[b] + 1 == self.nmax[b]:
                    continue
                old_n, new_n = state.inds[b].sum(axis=-1), new_inds[b].sum(axis=-1)
                edge[old_n == self.nmin[b]] += np.log(1 / 2.0)
                edge[old_n == self.nmax[b]] += np.log(1 / 2.0)
                edge[new_n == self.nmin[b]] -= np.log(1 / 2.0)
                edge[new_n == self.nmax[b]] -= np.log(1 / 2.0)
            factors += edge
            u_acc = st.rj_accept(it, T, W, gidx=gi) if hasattr(st, "rj_accept") else st.accept_uniforms(it, 1, T, W)
            keep = self._accept(state, q, new_inds, factors, u_acc, branches_run=branches)
        self._swaps(state, it, adapt=False)  # rj.py:381-382, once, after the last split
        return keep  # rj.py:385: the accepts of the LAST split are what the move counts

    def initialise(self, state):
        if state.logp is None:
            state.logp = mb_log_prior(self.priors, state.coords, state.inds)
        if state.logl is None:
            state.logl = mb_log_like(self.like, state.coords, state.inds, state.logp)
        return state

    def iterate(self, state):
        """returns (in-model accept mask, rj accept mask); the stream position advances by 2 (one per swap pass)"""
        acc = self.group_step(state)
        swaps_in_model = None if self.swaps_accepted is None else self.swaps_accepted.copy()
        self.iteration += 1
        racc = self.rj_step(state)
        self.iteration += 1
        self.swaps_in_model = swaps_in_model
        return acc, racc
