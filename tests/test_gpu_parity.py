"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI.

  * replay mode vs the golden vectors recorded from the unmodified reference: accept masks, swap counts and
    move choices bit-equal; coords / logl / logp / betas to 1e-10 relative (BASELINE.json tolerance);
  * philox (production) mode vs the oracle fed with the same counter-based streams, up to the full
    BASELINE config sizes;
  * size-independent properties of the swap pass and of the sampler at full size.
"""
import ctypes

import numpy as np
import pytest

from oracle import eryn_oracle as orc
from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # north_star: "within 1e-10 relative for log-prob/coord floats"


def dev_like(olike):
    from eryn_b200 import likelihood as lk
    if isinstance(olike, orc.GaussianLike):
        return lk.GaussianLikelihood(olike.mu, olike.prec)
    if isinstance(olike, orc.RosenbrockLike):
        return lk.RosenbrockLikelihood()
    return lk.GaussianMixtureLikelihood(olike.mus, olike.sigmas, olike.weights)


def _nprop(mv):  # a CombineMove counts proposals in its sub-moves (combine.py:126)
    return sum(m.num_proposals for m in mv.moves) if hasattr(mv, "moves") else mv.num_proposals


def _acc(mv):  # CombineMove.accepted is the list of its sub-moves' counters (combine.py:32-41)
    return np.sum(mv.accepted, axis=0) if hasattr(mv, "moves") else mv.accepted


def dev_moves(moves, priors=None):
    from eryn_b200.moves import CombineMove, DistributionGenerate, GaussianMove, MTDistGenMove, StretchMove
    out = []
    for m in moves:
        kw = {}
        if m.get("gibbs") is not None:  # the reference's input form: a list of (branch, mask) tuples / branch names
            kw["gibbs_sampling_setup"] = [("model_0", g) if g is not None else "model_0" for g in m["gibbs"]]
        if m["kind"] == "combine":
            out.append(CombineMove(dev_moves(m["moves"], priors)))
        elif m["kind"] == "stretch":
            out.append(StretchMove(a=m.get("a", 2.0), randomize_split=m.get("randomize_split", True),
                                   live_dangerously=m.get("live_dangerously", False), **kw))
        elif m["kind"] == "distgen":
            out.append(DistributionGenerate({"model_0": priors}))
        elif m["kind"] == "mt":  # generate_dist as a bare ProbDistContainer, as the reference's test passes it
            out.append(MTDistGenMove(priors, num_try=m["num_try"], independent=True))
        else:
            p = m["proposal"]
            if p.get("mode") is not None:
                kw["mode"] = p["mode"]
            if p.get("factor") is not None:
                kw["factor"] = p["factor"]
            out.append(GaussianMove({"model_0": p["scale"] ** 2 if p["kind"] == "scalar" else p["cov"]}, **kw))
    return out


def close(a, b, what):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-300, err_msg=what)


# ----------------------------------------------------------------------------------------------------
# 1. replay mode against the reference's golden vectors
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.CASES))
def test_replay_matches_reference_golden(name):
    from eryn_b200 import EnsembleSampler
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    g = cases.load(name)
    c = cases.CASES[name]
    d, W, T = int(g["ndim"]), int(g["nwalkers"]), int(g["ntemps"])
    np.random.seed(int(g["seed"]))  # same call order as tests/golden/make_golden.py
    priors = ProbDistContainer({i: uniform_dist(float(g["lo"]), float(g["hi"])) for i in range(d)})
    tk = {}
    if bool(g["tempered"]):
        tk = dict(ntemps=T, adaptive=c.get("adaptive", True), permute=c.get("permute", True))
        tk.update(c.get("tempering", {}))
    moves = dev_moves(c["moves"], priors)
    w = c.get("weights", [1.0] * len(moves))
    periodic = None
    if c.get("periods") is not None:
        periodic = {"model_0": {i: float(p) for i, p in enumerate(c["periods"]) if p > 0}}
    sampler = EnsembleSampler(W, d, dev_like(c["like"](d)), priors, tempering_kwargs=tk,
                              moves=list(zip(moves, w)), rng="numpy-replay", periodic=periodic)
    x0 = priors.rvs(size=(T, W))
    assert np.array_equal(x0, g["x0"])
    prev = [np.zeros((T, W)) for _ in moves]
    prev_n = [0 for _ in moves]
    for it, state in enumerate(sampler.sample(x0, iterations=int(g["nits"]), store=False)):
        which, acc = None, None
        for k, mv in enumerate(sampler.moves):
            if _nprop(mv) != prev_n[k]:
                a = _acc(mv)
                which, acc = k, (a - prev[k]).astype(g["accepted"].dtype)  # masks; accept counts for combined moves
                prev[k], prev_n[k] = a.copy(), _nprop(mv)
        assert which == g["move"][it], f"move choice differs at iteration {it}"
        assert np.array_equal(acc, g["accepted"][it]), f"accept mask differs at iteration {it}"
        close(state.branches_coords["model_0"][:, :, 0, :], g["coords"][it], f"coords it {it}")
        close(state.log_like, g["logl"][it], f"logl it {it}")
        close(state.log_prior, g["logp"][it], f"logp it {it}")
        if T > 1:
            assert np.array_equal(sampler.temperature_control.swaps_accepted, g["swaps"][it]), f"swaps it {it}"
            close(state.betas, g["betas"][it], f"betas it {it}")


def test_known_answers_through_run_mcmc():
    """KAT-1 / KAT-2 of SURVEY.md App. C through the public run_mcmc + backend path."""
    from eryn_b200 import EnsembleSampler
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    np.random.seed(42)
    pri = ProbDistContainer({i: uniform_dist(-5.0, 5.0) for i in range(5)})
    s = EnsembleSampler(32, 5, GaussianLikelihood(np.zeros(5), np.eye(5)), pri, rng="numpy-replay")
    x0 = pri.rvs(size=(32,))
    s.run_mcmc(x0, 100)
    last = s.get_last_sample()
    assert abs(last.branches_coords["model_0"].sum() - (-6.863121888547821e-01)) < 1e-9
    assert abs(last.log_like.sum() - (-6.935659199141485e01)) < 1e-8
    assert s.backend.accepted.sum() == 1757
    np.random.seed(42)
    pri = ProbDistContainer({i: uniform_dist(-5.0, 5.0) for i in range(3)})
    s = EnsembleSampler(16, 3, GaussianLikelihood(np.zeros(3), np.eye(3)), pri, tempering_kwargs=dict(ntemps=4),
                        rng="numpy-replay")
    x0 = pri.rvs(size=(4, 16))
    s.run_mcmc(x0, 50)
    assert s.backend.accepted.sum() == 1859
    assert list(s.backend.swaps_accepted) == [228, 436, 677]
    np.testing.assert_allclose(s.get_last_sample().betas,
                               [1, 0.24676878824520287, 0.05735976075520548, 0.01115873741850507], rtol=1e-10)
    assert abs(s.get_last_sample().branches_coords["model_0"].sum() - 2.892038585097532e01) < 1e-8


# ----------------------------------------------------------------------------------------------------
# 2. philox (production) mode against the oracle with identical streams
# ----------------------------------------------------------------------------------------------------
def c2_like(d):
    return orc.GaussianLike(np.zeros(d), cases.corr_prec(d))


def gmix_like(d, K=4, seed=5):
    r = np.random.RandomState(seed)
    return orc.GaussianMixtureLike(r.uniform(-5, 5, size=(K, d)), r.uniform(0.5, 1.5, size=K), np.full(K, 1.0 / K))


PHILOX_CASES = {
    # name: (T, W, d, like factory, moves, weights, nits, lo, hi)
    "c1": (1, 32, 5, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 30, -5, 5),
    "odd": (3, 99, 5, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=1.7)], [1.0], 15, -5, 5),
    "tight": (4, 64, 8, c2_like, [dict(kind="stretch", a=2.0)], [1.0], 15, -1.5, 1.5),
    "c2_full": (16, 4096, 8, c2_like, [dict(kind="stretch", a=2.0)], [1.0], 6, -10, 10),
    "c3_full": (16, 4096, 8, lambda d: orc.RosenbrockLike(),
                [dict(kind="stretch", a=2.0), dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.1))],
                [0.5, 0.5], 8, -10, 10),
    "gauss_matrix": (3, 40, 3, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                     [dict(kind="gaussian", proposal=dict(kind="matrix", cov=cases.COV3,
                                                          chol=np.linalg.cholesky(cases.COV3)))], [1.0], 15, -5, 5),
    "c4_slice": (8, 1024, 20, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 5, -10, 10),
    "distgen": (3, 64, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                [dict(kind="stretch", a=2.0), dict(kind="distgen")], [0.5, 0.5], 12, -2, 2),
    "combine": (3, 64, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                [dict(kind="combine", moves=[dict(kind="stretch", a=2.0),
                                             dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.4))])], [1.0], 10, -5, 5),
    "distgen_d8": (2, 128, 8, c2_like, [dict(kind="distgen")], [1.0], 6, -1, 1),
    "nosplit": (3, 99, 5, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                [dict(kind="stretch", a=2.0, randomize_split=False)], [1.0], 10, -5, 5),
    "d13": (2, 64, 13, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 8, -5, 5),
    "d18": (2, 64, 18, lambda d: orc.RosenbrockLike(), [dict(kind="stretch", a=2.0)], [1.0], 8, -5, 5),
    "d30": (2, 128, 30, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 6, -10, 10),
    # ladder lengths that exercise every swap-kernel shape: 8/16/32 lanes per chain, two rungs per lane (T = 40),
    # the staging-buffer path (T = 70, and rows longer than 32 doubles cannot occur in the fused kernels), T = 128
    "T8": (8, 64, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 6, -5, 5),
    "T24": (24, 64, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 6, -5, 5),
    "T40": (40, 48, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 6, -5, 5),
    "T40_d20": (40, 48, 20, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 4, -10, 10),
    "T70": (70, 32, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 5, -5, 5),
    "T128": (128, 32, 8, c2_like, [dict(kind="stretch", a=2.0)], [1.0], 4, -10, 10),
    "gmix_d8": (4, 256, 8, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 6, -10, 10),
    "gmix_k6_d20": (3, 128, 20, lambda d: gmix_like(d, K=6, seed=8), [dict(kind="stretch", a=2.0)], [1.0], 5, -10, 10),
    "gmix_tight_d20": (3, 128, 20, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 6, -3.2, 3.2),
    # GaussianMove modes and factor (gaussian.py:134-181)
    "gauss_modes_d8": (3, 64, 8, c2_like,
                       [dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.4, mode="random", factor=3.0)),
                        dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.5, mode="sequential")),
                        dict(kind="gaussian", proposal=dict(kind="matrix", cov=np.eye(8) * 0.01, chol=np.eye(8) * 0.1,
                                                            factor=1.5))], [0.4, 0.4, 0.2], 12, -10, 10),
    # multiple-try Metropolis (multipletry.py + mtdistgen.py): few and many tries, every likelihood functor
    "mt_d4": (3, 64, 4, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d) / 0.25),
              [dict(kind="stretch", a=2.0), dict(kind="mt", num_try=5)], [0.5, 0.5], 12, -2, 2),
    "mt_25_d8": (4, 128, 8, c2_like, [dict(kind="mt", num_try=25)], [1.0], 6, -3, 3),
    "mt_gmix_d20": (2, 64, 20, gmix_like, [dict(kind="mt", num_try=40), dict(kind="stretch", a=2.0)], [0.5, 0.5], 6, -6, 6),
    "mt_rosen_d5": (2, 64, 5, lambda d: orc.RosenbrockLike(), [dict(kind="mt", num_try=9)], [1.0], 8, -2, 2),
    # Gibbs splits at the parameter level (move.py:113-402) in the exact-length (D = 8) and the padded (D = 5) kernels
    "gibbs_d8": (3, 64, 8, c2_like,
                 [dict(kind="stretch", a=2.0, gibbs=[cases.gmask(8, 0, 1, 2), cases.gmask(8, 3, 4, 5, 6, 7)]),
                  dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.3),
                       gibbs=[cases.gmask(8, 7), None, cases.gmask(8, 0, 2, 4, 6)])], [0.5, 0.5], 10, -10, 10),
    "gibbs_d5": (2, 32, 5, lambda d: orc.RosenbrockLike(),
                 [dict(kind="stretch", a=2.0, gibbs=[cases.gmask(5, 4), cases.gmask(5, 0, 1, 2, 3), None])], [1.0], 8, -5, 5),
    "W_small": (3, 16, 3, lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0], 10, -5, 5),
    "W_257": (2, 257, 5, lambda d: orc.RosenbrockLike(), [dict(kind="stretch", a=2.0)], [1.0], 6, -5, 5),
}


def run_philox_case(T, W, d, like_f, moves, weights, nits, lo, hi, seed=2024, untempered=False, check_every=1,
                    periods=None):
    from eryn_b200 import EnsembleSampler
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    olike = like_f(d)
    prior = orc.BoxPrior(np.full(d, float(lo)), np.full(d, float(hi)))
    betas = None if untempered else (orc.make_ladder_default(d, T) if T > 1 else np.array([1.0]))
    sched = np.random.RandomState(7)
    osmp = orc.OracleSampler(prior, olike, moves, weights, orc.PhiloxStreams(seed, schedule_random=sched), betas=betas,
                             periods=periods)
    x0 = np.random.RandomState(1).uniform(max(lo, -3), min(hi, 3), size=(T, W, d))
    ost = osmp.initialise(orc.OState(x0))

    np.random.seed(7)  # the sampler's private stream (move schedule) = copy of the global state
    priors = ProbDistContainer({i: uniform_dist(float(lo), float(hi)) for i in range(d)})
    tk = {} if untempered else dict(ntemps=T)
    dm = dev_moves(moves, priors)
    periodic = None if periods is None else {"model_0": {i: float(p) for i, p in enumerate(periods) if p > 0}}
    smp = EnsembleSampler(W, d, dev_like(olike), priors, tempering_kwargs=tk, moves=list(zip(dm, weights)),
                          rng="philox", seed=seed, periodic=periodic)
    n_acc = 0
    for it, state in enumerate(smp.sample(x0, iterations=nits, store=False)):
        acc_o = osmp.iterate(ost)
        if it % check_every and it != nits - 1:
            continue
        mv = smp.moves[osmp.last_move]
        close(state.branches_coords["model_0"], ost.coords, f"coords it {it}")
        close(state.log_like, ost.logl, f"logl it {it}")
        close(state.log_prior, ost.logp, f"logp it {it}")
        n_acc += int(acc_o.sum())
        if not untempered and T > 1:
            assert np.array_equal(smp.temperature_control.swaps_accepted, osmp.swaps_accepted), f"swaps it {it}"
            close(state.betas, osmp.betas, f"betas it {it}")
    total = sum(_acc(m).sum() for m in smp.moves)
    assert total > 0
    return smp, osmp


@pytest.mark.parametrize("name", list(PHILOX_CASES))
def test_philox_matches_oracle(name):
    run_philox_case(*PHILOX_CASES[name])


def test_philox_periodic_matches_oracle():
    """periodic parameters (utils/periodic.py) in the fused kernels, production streams, D = 8 (exact-length kernel) and
    D = 5 (padded kernel); walkers start uniform over the period so distances through the boundary are common"""
    for d, T, W in ((8, 4, 256), (5, 3, 64)):
        periods = np.zeros(d)
        periods[[0, 2, d - 1]] = [3.0, 2 * np.pi, 1.5]
        mu = np.linspace(0.1, 1.4, d)
        run_philox_case(T, W, d, lambda dd: orc.GaussianLike(mu, np.eye(dd) / 0.3),
                        [dict(kind="stretch", a=2.0), dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.3))],
                        [0.5, 0.5], 12, 0.0, 7.0, periods=periods)


def test_philox_untempered_matches_oracle():
    run_philox_case(1, 64, 4, lambda d: orc.RosenbrockLike(), [dict(kind="stretch", a=2.0)], [1.0], 20, -5, 5,
                    untempered=True)


def test_philox_accept_masks_bit_exact_c2():
    """accept masks / swap counts bit-equal at the full BASELINE config-2 size (16 x 4096 x 8-d)."""
    T, W, d = 16, 4096, 8
    smp, osmp = run_philox_case(T, W, d, c2_like, [dict(kind="stretch", a=2.0)], [1.0], 3, -10, 10, seed=99)
    # move.accepted accumulated on the device equals what the oracle's masks add up to
    from eryn_b200 import EnsembleSampler  # noqa: F401
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    o2 = orc.OracleSampler(prior, c2_like(d), [dict(kind="stretch", a=2.0)], [1.0],
                           orc.PhiloxStreams(99, schedule_random=np.random.RandomState(7)),
                           betas=orc.make_ladder_default(d, T))
    x0 = np.random.RandomState(1).uniform(-3, 3, size=(T, W, d))
    st = o2.initialise(orc.OState(x0))
    tot = np.zeros((T, W))
    for _ in range(3):
        tot += o2.iterate(st)
    assert np.array_equal(smp.moves[0].accepted, tot)


@pytest.fixture
def k1_lanes(monkeypatch):
    """force a variant of the stretch kernel (EB_K1_LPW is read at every launch): 1 = one thread per walker,
    2 / 4 = the lane-split kernel (csrc/stretch_lanes.cuh) that the library picks by itself only for HBM-sized shapes"""
    def force(lpw):
        monkeypatch.setenv("EB_K1_LPW", str(lpw))
    return force


@pytest.mark.parametrize("lpw,name", [(4, "c2_full"), (2, "c2_full"), (4, "c3_full"), (2, "c3_full"), (4, "c4_slice"),
                                      (4, "tight"), (2, "tight"), (4, "T40_d20"), (4, "gmix_d8"), (4, "gmix_k6_d20"),
                                      (4, "gmix_tight_d20"), (2, "gmix_d8")])
def test_lane_split_stretch_kernel_matches_oracle(k1_lanes, lpw, name):
    """the lane-split stretch kernel (2 / 4 lanes per walker) against the oracle: every likelihood functor, priors tight
    enough that many proposals leave the box, more mixture components than lanes, exact row lengths 8 and 20"""
    k1_lanes(lpw)
    run_philox_case(*PHILOX_CASES[name])


@pytest.mark.parametrize("lpw", [2, 4])
def test_lane_split_stretch_kernel_periodic_and_replay(k1_lanes, lpw):
    k1_lanes(lpw)
    d, T, W = 8, 4, 256
    periods = np.zeros(d)
    periods[[0, 2, d - 1]] = [3.0, 2 * np.pi, 1.5]
    mu = np.linspace(0.1, 1.4, d)
    run_philox_case(T, W, d, lambda dd: orc.GaussianLike(mu, np.eye(dd) / 0.3), [dict(kind="stretch", a=2.0)], [1.0], 10,
                    0.0, 7.0, periods=periods)
    for name in ("c2_small", "c2_tightprior"):   # replay mode against the reference's golden vectors (8-d)
        test_replay_matches_reference_golden(name)


def test_philox_config4_full_size():
    """BASELINE config 4 at full size on one GPU: 32 temperatures x 16384 walkers x 20-d mixture of 4 Gaussians,
    StretchMove + PT swaps + adaptation, production streams against the oracle (the lane-split stretch kernel is what
    the library launches at this size; the swap pass moves 20-double rows over 32 rungs)."""
    smp, osmp = run_philox_case(32, 16384, 20, gmix_like, [dict(kind="stretch", a=2.0)], [1.0], 2, -10, 10, seed=31)
    assert osmp.swaps_accepted.sum() > 0


# ----------------------------------------------------------------------------------------------------
# 3. the split path (callable on CUDA tensors) gives the same chain as the fused functor
# ----------------------------------------------------------------------------------------------------
def test_split_path_equals_fused():
    import torch
    from eryn_b200 import EnsembleSampler
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    d, T, W = 6, 4, 64
    P = cases.corr_prec(d)
    Pt = torch.from_numpy(P).cuda()

    def torch_like(x):  # x [N, 1, d]
        v = x[:, 0, :]
        return -0.5 * torch.einsum("ni,ij,nj->n", v, Pt, v)

    outs = []
    for like in (GaussianLikelihood(np.zeros(d), P), torch_like):
        np.random.seed(3)
        pri = ProbDistContainer({i: uniform_dist(-4.0, 4.0) for i in range(d)})
        s = EnsembleSampler(W, d, like, pri, tempering_kwargs=dict(ntemps=T), rng="philox", seed=5)
        x0 = np.random.RandomState(2).uniform(-2, 2, size=(T, W, d))
        st = s.run_mcmc(x0, 12)
        outs.append((st.branches_coords["model_0"].copy(), st.log_like.copy(), s.backend.accepted.copy(),
                     s.backend.swaps_accepted.copy()))
    close(outs[0][0], outs[1][0], "coords")
    close(outs[0][1], outs[1][1], "logl")
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])


# ----------------------------------------------------------------------------------------------------
# 4. properties at full size
# ----------------------------------------------------------------------------------------------------
def test_swap_pass_is_a_permutation_and_consistent():
    """A swap pass only permutes walkers inside each walker column family: the multiset of (logl, logp, coords)
    rows is conserved, every row stays self-consistent, and swaps_accepted matches the number of moved rows."""
    import torch
    from eryn_b200.device import DeviceContext
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.moves import TemperatureControl
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    from eryn_b200.state import State
    T, W, d = 16, 4096, 8
    P = cases.corr_prec(d)
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    ctx = DeviceContext(pri, GaussianLikelihood(np.zeros(d), P), rng="philox", seed=11)
    tc = TemperatureControl(d, W, ntemps=T)
    tc.bind(ctx)
    x0 = np.random.RandomState(0).uniform(-3, 3, size=(T, W, 1, d))
    ds = ctx.upload(State(x0), betas=tc.betas_dev)
    ctx.eval_state(ds)
    before = ctx.download(ds)
    b0 = tc.betas.copy()
    tc.temper_comps(ds)
    after = ctx.download(ds)
    key_b = np.sort(before.log_like.ravel())
    key_a = np.sort(after.log_like.ravel())
    assert np.array_equal(key_b, key_a)
    assert np.array_equal(np.sort(before.branches_coords["model_0"].reshape(-1, d), axis=0),
                          np.sort(after.branches_coords["model_0"].reshape(-1, d), axis=0))
    # rows still consistent: recomputing logl from the moved coords reproduces the moved logl
    chk = ctx.upload(State(after.branches_coords["model_0"].copy()))
    ctx.eval_state(chk)
    assert np.array_equal(chk.logl.cpu().numpy(), after.log_like)
    sw = tc.swaps_accepted
    assert sw.shape == (T - 1,) and np.all(sw > 0) and np.all(sw <= W)
    # a walker moves up one rung per accepted swap: rows that changed at rung i>=1 >= swaps at that rung
    changed = (before.log_like != after.log_like).sum(axis=1)
    assert changed.sum() >= sw.sum()
    assert tc.time == 1 and not np.array_equal(tc.betas, b0) and tc.betas[0] == 1.0 and tc.betas[-1] == b0[-1]


def test_sampler_recovers_target_moments_c2():
    """Full config 2 in production mode: the beta=1 chain samples the correlated Gaussian (mean 0, cov Sigma)."""
    from eryn_b200 import EnsembleSampler
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    T, W, d = 16, 4096, 8
    P = cases.corr_prec(d)
    cov = np.linalg.inv(P)
    np.random.seed(1)
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    s = EnsembleSampler(W, d, GaussianLikelihood(np.zeros(d), P), pri, tempering_kwargs=dict(ntemps=T), rng="philox",
                        seed=321)
    x0 = np.random.RandomState(4).uniform(-3, 3, size=(T, W, d))
    st = s.run_mcmc(x0, 4, burn=600, thin_by=25)
    x = s.get_chain()["model_0"][:, 0, :, 0, :].reshape(-1, d)
    assert np.all(np.abs(x.mean(0)) < 0.05)
    np.testing.assert_allclose(np.cov(x.T), cov, atol=0.08)
    assert 0.2 < s.backend.accepted.mean() / s.backend.iteration <= 1.0
    assert np.all(np.diff(st.betas) < 0)


# ----------------------------------------------------------------------------------------------------
# 5. reference-facing host-buffer entry point and error behaviour
# ----------------------------------------------------------------------------------------------------
def test_run_host_matches_oracle():
    from eryn_b200 import _lib
    lib = _lib.require_device()
    T, W, d, nit, seed = 4, 256, 8, 5, 17
    olike = c2_like(d)
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    osmp = orc.OracleSampler(prior, olike, [dict(kind="stretch", a=2.0)], [1.0], orc.PhiloxStreams(seed),
                             betas=orc.make_ladder_default(d, T))
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(T, W, d))
    ost = osmp.initialise(orc.OState(x0))
    coords = np.ascontiguousarray(ost.coords.copy())
    logl, logp = ost.logl.copy(), ost.logp.copy()
    betas = osmp.betas.copy()
    for _ in range(nit):
        osmp.iterate(ost)
    par = np.ascontiguousarray(olike.params())
    lo, hi = prior.lo.copy(), prior.hi.copy()
    swaps = np.zeros(T - 1, dtype=np.int32)
    cnt = np.zeros((T, W), dtype=np.uint32)
    job = _lib.eb_host_job()
    job.ntemps, job.nwalkers, job.nleaves, job.ndim = T, W, 1, d
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    job.coords_host, job.logl_host, job.logp_host, job.betas_host = vp(coords), vp(logl), vp(logp), vp(betas)
    job.prior_lo_host, job.prior_hi_host = vp(lo), vp(hi)
    job.like_kind, job.like_ncomp, job.like_nparams, job.like_params_host = 0, 0, par.size, vp(par)
    job.stretch_a, job.gauss_scale, job.seed, job.iter0 = 2.0, 0.1, seed, 0
    job.adapt = _lib.eb_adapt(1, -1, 10000.0, 100.0)
    job.adapt_time0, job.permute, job.randomize_split = 0, 1, 1
    job.swaps_accepted_host, job.accepted_count_host = vp(swaps), vp(cnt)
    _lib.check(lib.eb_run_host(ctypes.byref(job), nit), "eb_run_host")
    close(coords, ost.coords, "coords")
    close(logl, ost.logl, "logl")
    close(betas, osmp.betas, "betas")
    assert np.array_equal(swaps, osmp.swaps_accepted)
    assert job.iter0 == nit and job.adapt_time0 == nit and cnt.sum() > 0


def test_error_behaviour_matches_reference():
    from eryn_b200 import EnsembleSampler
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    d = 5
    pri = ProbDistContainer({i: uniform_dist(-5.0, 5.0) for i in range(d)})
    like = GaussianLikelihood(np.zeros(d), np.eye(d))
    s = EnsembleSampler(6, d, like, pri)  # fewer walkers than 2*ndim: red_blue.py:103-114
    with pytest.raises(RuntimeError):
        s.run_mcmc(np.random.RandomState(0).uniform(-1, 1, size=(1, 6, d)), 1)
    s = EnsembleSampler(32, d, like, pri)
    with pytest.raises(ValueError):  # ensemble.py:877-885 incompatible input dimensions
        s.run_mcmc(np.zeros((1, 31, d)), 1)
    with pytest.raises(ValueError):  # start outside the prior: "The initial log_prior was +/- infinite"
        s.run_mcmc(np.full((1, 32, d), 7.0), 1)
    with pytest.raises(ValueError):
        s.run_mcmc(np.random.RandomState(0).uniform(-1, 1, size=(1, 32, d)), 1, thin_by=0)
