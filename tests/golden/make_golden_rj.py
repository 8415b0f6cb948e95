"""Golden vectors for BASELINE config 5 (reversible jump + group stretch), from the UNMODIFIED reference.

Run in the build container only:   python tests/golden/make_golden_rj.py

Two branches ("gauss": a*exp(-(t-b)^2/(2c^2)) pulses, "sine": a*sin(2 pi b t + c)), nleaves_min 0, the reference's
`log_like_fn_gauss_and_sine` (tests/test_eryn.py:69-92), uniform priors (tests/test_eryn.py:416-427), in-model move = a
GroupStretchMove subclass whose friend rule is the one of the reference's test fixture `MeanGaussianGroupMove`
(tests/test_eryn.py:813-907), applied to both branches: friends = the cold chain's active leaves sorted by their
second parameter, every active leaf keeps the indices of its `nfriends` nearest friends (refreshed every
`n_iter_update` iterations, leaves born in between are fixed up), and a proposal picks one of them at random.
`rj_moves=True` (DistributionGenerateRJ from the priors, all branches together).

Recorded after every iteration: coords / inds per branch, log_like, log_prior, betas, in-model accept mask, rj accept
mask, swap counts of the in-model move.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import _install_shim  # noqa: E402  (installs the stub modules + reference path)

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
from eryn.ensemble import EnsembleSampler  # noqa: E402
from eryn.moves import GroupStretchMove  # noqa: E402
from eryn.prior import uniform_dist  # noqa: E402
from eryn.state import BranchSupplemental, State  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- the reference test's likelihood (tests/test_eryn.py:38-92) ---------------------------------
def gaussian_pulse(x, a, b, c):
    return a * np.exp(-((x - b) ** 2) / (2 * c**2))


def combine_gaussians(t, params):
    template = np.zeros_like(t)
    for param in params:
        template += gaussian_pulse(t, *param)
    return template


def sine(x, a, b, c):
    return a * np.sin(2 * np.pi * b * x + c)


def combine_sine(t, params):
    template = np.zeros_like(t)
    for param in params:
        template += sine(t, *param)
    return template


def log_like_fn_gauss_and_sine(params_both, t, data, sigma):
    params_gauss, params_sine = params_both
    template = np.zeros_like(t)
    if params_gauss is not None:
        template += combine_gaussians(t, params_gauss)
    if params_sine is not None:
        template += combine_sine(t, params_sine)
    return -0.5 * np.sum(((template - data) / sigma) ** 2, axis=-1)


class NearestFriendsGroupMove(GroupStretchMove):
    """tests/test_eryn.py:813-907 (MeanGaussianGroupMove), for every branch, friend parameter = index 1."""

    def __init__(self, key_index=1, **kwargs):
        GroupStretchMove.__init__(self, **kwargs)
        self.key_index = key_index
        self.friends, self.means = {}, {}

    def _closest(self, name, vals):
        dist = np.abs(vals[:, None] - self.means[name][None, :])
        return np.argsort(dist, axis=1)[:, : self.nfriends]

    def setup_friends(self, branches):
        for name, br in branches.items():
            friends = br.coords[0, br.inds[0]]
            means = friends[:, self.key_index].copy()
            self.means[name], uni = np.unique(means, return_index=True)
            self.friends[name] = friends[uni]
            srt = np.argsort(self.means[name])
            self.friends[name][:] = self.friends[name][srt]
            self.means[name][:] = self.means[name][srt]
            cur = br.coords[br.inds, self.key_index]
            br.branch_supplemental[br.inds] = {"inds_closest": self._closest(name, cur)}
            T, W, L = br.inds.shape
            br.branch_supplemental[~br.inds] = {"inds_closest": -np.ones((T, W, L, self.nfriends), dtype=int)[~br.inds]}

    def fix_friends(self, branches):
        for name, br in branches.items():
            fix = br.inds & np.all(br.branch_supplemental[:]["inds_closest"] == -1, axis=-1)
            if not np.any(fix):
                continue
            cur = br.coords[fix, self.key_index]
            br.branch_supplemental[fix] = {"inds_closest": self._closest(name, cur)}

    def find_friends(self, name, s, s_inds=None, branch_supps=None):
        friends = np.zeros_like(s)
        here = branch_supps[name][s_inds]["inds_closest"]
        pick = here[np.arange(here.shape[0]), np.random.randint(self.nfriends, size=(here.shape[0],))]
        friends[s_inds] = self.friends[name][pick]
        return friends


def run_case(name, seed, T, W, L, nt, nfriends, n_iter_update, nits, sigma, rj_moves=True):
    np.random.seed(seed)
    branch_names = ["gauss", "sine"]
    ndims = {"gauss": 3, "sine": 3}
    nleaves_max = dict(L)
    nleaves_min = {"gauss": 0, "sine": 0}
    t = np.linspace(-1, 1, nt)
    ginj = np.array([[3.3, -0.2, 0.1], [2.6, -0.1, 0.1], [3.4, 0.0, 0.1], [2.9, 0.3, 0.1]])[: min(4, L["gauss"] - 1)]
    sinj = np.array([[1.3, 10.1, 1.0], [0.8, 4.6, 1.2]])[: min(2, L["sine"] - 1)]
    y = combine_gaussians(t, ginj) + combine_sine(t, sinj) + sigma * np.random.randn(nt)
    coords = {k: np.zeros((T, W, nleaves_max[k], 3)) for k in branch_names}
    inds = {k: np.zeros((T, W, nleaves_max[k]), dtype=bool) for k in branch_names}
    for k, inj in (("gauss", ginj), ("sine", sinj)):
        for nn in range(len(inj)):
            coords[k][:, :, nn] = np.random.multivariate_normal(inj[nn], np.diag(np.ones(3) * 1e-4), size=(T, W))
            inds[k][:, :, nn] = True
    priors = {
        "gauss": {0: uniform_dist(2.5, 3.5), 1: uniform_dist(t.min(), t.max()), 2: uniform_dist(0.01, 0.21)},
        "sine": {0: uniform_dist(0.5, 1.5), 1: uniform_dist(1.0, 20.0), 2: uniform_dist(0.0, 2 * np.pi)},
    }
    move = NearestFriendsGroupMove(nfriends=nfriends, n_iter_update=n_iter_update)
    sampler = EnsembleSampler(W, ndims, log_like_fn_gauss_and_sine, priors, args=[t, y, sigma],
                              tempering_kwargs=dict(ntemps=T), nbranches=2, branch_names=branch_names,
                              nleaves_max=nleaves_max, nleaves_min=nleaves_min, moves=move, rj_moves=rj_moves)
    lp = sampler.compute_log_prior(coords, inds=inds)
    ll = sampler.compute_log_like(coords, inds=inds, logp=lp)[0]
    supp = {k: BranchSupplemental({"inds_closest": np.zeros(inds[k].shape + (nfriends,), dtype=int)},
                                  base_shape=inds[k].shape) for k in branch_names}
    state0 = State({k: v.copy() for k, v in coords.items()}, log_like=ll.copy(), log_prior=lp.copy(),
                   inds={k: v.copy() for k, v in inds.items()}, branch_supplemental=supp)
    rec = {k: [] for k in ("cg", "cs", "ig", "is", "logl", "logp", "betas", "acc", "rjacc", "swaps")}
    prev_acc = np.zeros((T, W))
    prev_rj = np.zeros((T, W))
    rjms = list(sampler.rj_moves)
    rec["rjmove"] = []
    prev_np = [0 for _ in rjms]
    for state in sampler.sample(state0, iterations=nits, store=False, skip_initial_state_check=True):
        rec["cg"].append(state.branches["gauss"].coords.copy())
        rec["cs"].append(state.branches["sine"].coords.copy())
        rec["ig"].append(state.branches["gauss"].inds.copy())
        rec["is"].append(state.branches["sine"].inds.copy())
        rec["logl"].append(state.log_like.copy())
        rec["logp"].append(state.log_prior.copy())
        rec["betas"].append(sampler.temperature_control.betas.copy())
        rec["acc"].append((move.accepted - prev_acc).astype(bool))
        rj_total = sum(m.accepted for m in rjms)
        rec["rjacc"].append((rj_total - prev_rj).astype(bool))
        rec["rjmove"].append(np.int64([k for k, m in enumerate(rjms) if m.num_proposals != prev_np[k]][0]))
        prev_np = [m.num_proposals for m in rjms]
        prev_acc, prev_rj = move.accepted.copy(), rj_total.copy()
        rec["swaps"].append(np.asarray(sampler.temperature_control.swaps_accepted).copy())
    out = dict(seed=seed, T=T, W=W, Lg=L["gauss"], Ls=L["sine"], nt=nt, nfriends=nfriends, n_iter_update=n_iter_update,
               nits=nits, sigma=sigma, t=t, y=y, cg0=coords["gauss"], cs0=coords["sine"], ig0=inds["gauss"],
               is0=inds["sine"], logl0=ll, logp0=lp, rj_mode=str(rj_moves))
    for k, v in rec.items():
        a = np.stack(v)
        out[k] = np.packbits(a, axis=-1) if a.dtype == bool else a
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    nl = np.stack(rec["ig"]).sum(-1)[-1, 0], np.stack(rec["is"]).sum(-1)[-1, 0]
    print(f"{name}: acc={np.stack(rec['acc']).sum()} rjacc={np.stack(rec['rjacc']).sum()} "
          f"logl.sum={rec['logl'][-1].sum():.12e} cold nleaves gauss={nl[0].tolist()} sine={nl[1].tolist()} "
          f"size={os.path.getsize(os.path.join(HERE, name + '.npz'))}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gibbs":
        # rj_moves="iterate_branches" / "separate_branches" (ensemble.py:434-470): the branches as Gibbs splits of the RJ move
        run_case("c5_iter", 31, 3, 12, {"gauss": 4, "sine": 3}, 48, 5, 4, 20, 2.0, rj_moves="iterate_branches")
        run_case("c5_sep", 32, 2, 16, {"gauss": 5, "sine": 3}, 32, 6, 3, 20, 3.0, rj_moves="separate_branches")
        sys.exit(0)
    run_case("c5_small", 2024, 3, 12, {"gauss": 4, "sine": 3}, 48, 5, 4, 24, 2.0)
    run_case("c5_wide", 7, 2, 16, {"gauss": 6, "sine": 2}, 32, 8, 3, 16, 3.0)
