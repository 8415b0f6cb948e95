"""GPU parity of the reversible-jump / group-stretch kernels (BASELINE config 5).  Everything goes through the C ABI.

  * replay mode vs golden vectors recorded from the unmodified reference (tests/golden/make_golden_rj.py): accept masks
    (in-model and rj), leaf flags and swap counts bit-equal, floats to 1e-10 relative;
  * philox mode vs the oracle with the same counter-based streams, up to the full config-5 size (8 x 2048 walkers,
    2 branches x 10 leaves)."""
import numpy as np
import pytest

from oracle import eryn_oracle as orc
from oracle import rj_oracle as rjo
from tests import cases_rj

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def close(a, b, what):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-300, err_msg=what)


def make_sampler(t, y, sigma, T, W, Lg, Ls, nfriends, n_iter_update, rng, seed=None, rj_moves=True):
    from eryn_b200 import EnsembleSampler
    from eryn_b200.moves import GroupStretchMove
    from eryn_b200.multibranch import PulseLikelihood
    from eryn_b200.prior import uniform_dist
    gl, gh = cases_rj.PRIOR_BOUNDS["gauss"](t)
    sl, sh = cases_rj.PRIOR_BOUNDS["sine"](t)
    priors = {"gauss": {i: uniform_dist(gl[i], gh[i]) for i in range(3)},
              "sine": {i: uniform_dist(sl[i], sh[i]) for i in range(3)}}
    like = PulseLikelihood(t, y, sigma, {"gauss": "gauss", "sine": "sine"})
    move = GroupStretchMove(nfriends=nfriends, n_iter_update=n_iter_update)
    return EnsembleSampler(W, {"gauss": 3, "sine": 3}, like, priors, tempering_kwargs=dict(ntemps=T), nbranches=2,
                           branch_names=["gauss", "sine"], nleaves_max={"gauss": Lg, "sine": Ls},
                           nleaves_min={"gauss": 0, "sine": 0}, moves=move, rj_moves=rj_moves, rng=rng, seed=seed), move


@pytest.mark.parametrize("name", cases_rj.NAMES)
def test_rj_replay_matches_reference_golden(name):
    from eryn_b200.state import State
    g = cases_rj.load(name)
    t, coords, inds, _, _ = cases_rj.replay_setup(g)  # leaves the GLOBAL stream where the reference's sampler started
    T, W, Lg, Ls = int(g["T"]), int(g["W"]), int(g["Lg"]), int(g["Ls"])
    mode = cases_rj.rj_mode(g)   # "together", "iterate_branches" or "separate_branches" (ensemble.py:410-470)
    smp, move = make_sampler(t, g["y"], float(g["sigma"]), T, W, Lg, Ls, int(g["nfriends"]), int(g["n_iter_update"]),
                             "numpy-replay", rj_moves=True if mode == "together" else mode)
    cd = {"gauss": coords[0], "sine": coords[1]}
    idd = {"gauss": inds[0], "sine": inds[1]}
    lp = smp.compute_log_prior(cd, inds=idd)
    ll = smp.compute_log_like(cd, inds=idd, logp=lp)[0]
    np.testing.assert_array_equal(lp, g["logp0"])
    close(ll, g["logl0"], "initial logl")
    st0 = State(cd, inds=idd, log_like=ll, log_prior=lp)
    pa, pr = np.zeros((T, W)), np.zeros((T, W))
    for it, state in enumerate(smp.sample(st0, iterations=int(g["nits"]), store=False)):
        a, r = move.accepted, sum(m.accepted for m in smp.rj_moves)
        assert np.array_equal((a - pa).astype(bool), g["acc"][it]), f"in-model accept mask differs at iteration {it}"
        assert np.array_equal((r - pr).astype(bool), g["rjacc"][it]), f"rj accept mask differs at iteration {it}"
        pa, pr = a.copy(), r.copy()
        assert np.array_equal(state.branches["gauss"].inds, g["ig"][it]), f"gauss inds it {it}"
        assert np.array_equal(state.branches["sine"].inds, g["is"][it]), f"sine inds it {it}"
        close(state.branches["gauss"].coords, g["cg"][it], f"gauss coords it {it}")
        close(state.branches["sine"].coords, g["cs"][it], f"sine coords it {it}")
        close(state.log_like, g["logl"][it], f"logl it {it}")
        close(state.log_prior, g["logp"][it], f"logp it {it}")
        close(state.betas, g["betas"][it], f"betas it {it}")
        assert np.array_equal(smp.temperature_control.swaps_accepted, g["swaps"][it]), f"swaps it {it}"


class _RJSched(object):
    """the device sampler's private MT19937 stream in philox mode: one draw for the in-model move (ensemble.py:971), one
    for the rj move (:990) per iteration; the oracle asks only for the rj choice of 'separate_branches'."""

    def __init__(self, rs, mode):
        self.rs, self.mode = rs, mode

    def choice(self, n, p):
        self.rs.choice(1, p=[1.0])           # the in-model move of this iteration
        return self.rs.choice(n, p=p)        # the rj move


def random_start(T, W, Lg, Ls, t, seed):
    r = np.random.RandomState(seed)
    coords, inds = [], []
    for L, (lo, hi) in ((Lg, cases_rj.PRIOR_BOUNDS["gauss"](t)), (Ls, cases_rj.PRIOR_BOUNDS["sine"](t))):
        lo, hi = np.asarray(lo), np.asarray(hi)
        coords.append(r.uniform(lo, hi, size=(T, W, L, 3)))
        n = r.randint(1, L, size=(T, W))
        inds.append(np.arange(L)[None, None, :] < n[:, :, None])
    return coords, inds


@pytest.mark.parametrize("T,W,Lg,Ls,nt,nfriends,nup,nits,mode", [
    (3, 32, 4, 3, 48, 6, 3, 7, "together"), (8, 2048, 10, 10, 64, 16, 2, 2, "together"),
    (3, 32, 4, 3, 48, 6, 3, 7, "iterate_branches"), (3, 32, 4, 3, 48, 6, 3, 9, "separate_branches"),
    (4, 512, 10, 10, 64, 16, 2, 3, "iterate_branches")])
def test_rj_philox_matches_oracle(T, W, Lg, Ls, nt, nfriends, nup, nits, mode):
    from eryn_b200.state import State
    seed = 31337
    t = np.linspace(-1, 1, nt)
    y = 3.0 * np.exp(-((t + 0.2) ** 2) / 0.02) + np.sin(2 * np.pi * 4.6 * t + 1.2) + np.random.RandomState(5).randn(nt)
    coords, inds = random_start(T, W, Lg, Ls, t, 11)
    like = rjo.PulseLike(t, y, 2.0, [0, 1])
    np.random.seed(77)  # the sampler's private stream (choice among the per-branch rj moves) = copy of the global state
    sched = np.random.RandomState(77)
    osmp = rjo.OracleSamplerMB(cases_rj.priors_for(t), like, [0, 0], [Lg, Ls],
                               rjo.PhiloxStreamsMB(seed, schedule_random=_RJSched(sched, mode)),
                               betas=orc.make_ladder_default(3 * (Lg + Ls), T), nfriends=nfriends, n_iter_update=nup,
                               rj_mode=mode)
    ost = osmp.initialise(rjo.MBState(coords, inds))
    smp, move = make_sampler(t, y, 2.0, T, W, Lg, Ls, nfriends, nup, "philox", seed=seed,
                             rj_moves=True if mode == "together" else mode)
    st0 = State({"gauss": coords[0], "sine": coords[1]}, inds={"gauss": inds[0], "sine": inds[1]})
    pa, pr = np.zeros((T, W)), np.zeros((T, W))
    for it, state in enumerate(smp.sample(st0, iterations=nits, store=False)):
        if it == 0:
            pass
        oa, orj = osmp.iterate(ost)
        a, r = move.accepted, sum(m.accepted for m in smp.rj_moves)
        assert np.array_equal((a - pa).astype(bool), oa), f"in-model accept mask differs at iteration {it}"
        assert np.array_equal((r - pr).astype(bool), orj), f"rj accept mask differs at iteration {it}"
        pa, pr = a.copy(), r.copy()
        assert np.array_equal(state.branches["gauss"].inds, ost.inds[0]) and np.array_equal(state.branches["sine"].inds, ost.inds[1])
        close(state.branches["gauss"].coords, ost.coords[0], f"gauss coords it {it}")
        close(state.branches["sine"].coords, ost.coords[1], f"sine coords it {it}")
        close(state.log_like, ost.logl, f"logl it {it}")
        close(state.log_prior, ost.logp, f"logp it {it}")
        close(state.betas, osmp.betas, f"betas it {it}")
        assert np.array_equal(smp.temperature_control.swaps_accepted, osmp.swaps_accepted), f"swaps it {it}"
        tab = smp.ctx.friend_table_host(smp._dstate_last) if hasattr(smp, "_dstate_last") else None
    assert pa.sum() > 0 and pr.sum() > 0


def test_rj_backend_counts_and_errors():
    from eryn_b200.moves import GroupStretchMove
    from eryn_b200.state import State
    t = np.linspace(-1, 1, 32)
    y = np.random.RandomState(1).randn(32)
    coords, inds = random_start(2, 16, 3, 3, t, 3)
    smp, move = make_sampler(t, y, 2.0, 2, 16, 3, 3, 4, 5, "philox", seed=2)
    st0 = State({"gauss": coords[0], "sine": coords[1]}, inds={"gauss": inds[0], "sine": inds[1]})
    smp.run_mcmc(st0, 6)
    assert smp.backend.iteration == 6 and smp.backend.rj_accepted.shape == (2, 16)
    nl = smp.get_nleaves()
    assert nl["gauss"].shape == (6, 2, 16) and nl["gauss"].max() <= 3 and nl["sine"].min() >= 0
    with pytest.raises(ValueError):
        GroupStretchMove(nfriends=4, n_iter_update=1)  # group.py:46-47
    with pytest.raises(ValueError):
        smp.run_mcmc(State({"gauss": coords[0][:, :8], "sine": coords[1][:, :8]}), 1)  # incompatible input dimensions
