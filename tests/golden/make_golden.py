"""Generate golden vectors by running the UNMODIFIED reference (Eryn @ /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

For each case the reference's EnsembleSampler is run under a fixed seed and, after every
iteration, coords / log_like / log_prior / accepted / swaps_accepted / betas are recorded.
The cases mirror BASELINE.json's configs at sizes small enough to commit (<300 KB total):

  c1_kat1        1 temp x 32 walkers x 5-d iso Gaussian, StretchMove, 100 its
  pt_kat2        4 temps x 16 walkers x 3-d, StretchMove + PT, 50 its               (KAT-2)
  c2_small       4 temps x 64 walkers x 8-d correlated Gaussian, Stretch + PT, vectorize=True
  c2_tightprior  same but prior U(-1.5,1.5): many proposals leave the prior (-inf / -1e300 route)
  c3_small       3 temps x 32 walkers x 4-d Rosenbrock, Stretch/Gaussian(scalar) 50/50 + PT
  gauss_matrix   3 temps x 24 walkers x 3-d, GaussianMove(full covariance) + PT
  odd_walkers    1 temp x 99 walkers x 5-d (tests/test_eryn.py:96 test_base shape), a=1.5
  noadapt_noperm 4 temps x 32 walkers, adaptive=False, permute=False
  tiny_live      3 temps x 6 walkers x 5-d, StretchMove(live_dangerously=True)   (argv: tiny)
  one_temp       tempering_kwargs=dict(ntemps=1): 1 temp x 20 walkers              (argv: tiny)
  nosplit        3 temps x 30 walkers, StretchMove(randomize_split=False)   (argv: nosplit)
  stop_adapt     5 temps x 40 walkers, adaptation_lag=30, adaptation_time=4, stop_adaptation=6   (argv: stop_adapt)
  gauss_modes    (`gauss_modes`) GaussianMove modes random / sequential / vector and `factor`
  mt_mix         (`mt`) 3 temps x 16 walkers x 3-d, MTDistGenMove(num_try=6, independent) mixed with StretchMove
  gibbs_mix      (`gibbs`) 3 temps x 24 walkers x 4-d, Stretch and Gaussian moves with parameter-level Gibbs splits
"""
import os
import sys
import types

import numpy as np


def _install_shim():
    def _stub(name, attrs=()):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        sys.modules[name] = m
        return m

    mpl = _stub("matplotlib")
    mpl.rcParams = {}
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("matplotlib.patches", ("Ellipse", "Rectangle"))
    _stub("matplotlib.colors").to_rgba = lambda *a, **k: None
    _stub("corner")
    _stub("seaborn")
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/src")


_install_shim()
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
from eryn.ensemble import EnsembleSampler  # noqa: E402
from eryn.moves import CombineMove, DistributionGenerate, GaussianMove, MTDistGenMove, StretchMove  # noqa: E402
from eryn.prior import ProbDistContainer, uniform_dist  # noqa: E402
from eryn.state import State  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- likelihoods (the reference calls these; definitions follow SURVEY.md §8d) ------------
def ll_single(x, mu, invcov):  # tests/test_eryn.py:33-35
    diff = x - mu
    return -0.5 * (diff * np.dot(invcov, diff.T).T).sum()


def ll_gauss_vec(x, mu, prec):
    d = x - mu
    return -0.5 * np.einsum("ni,ij,nj->n", d, prec, d)


def ll_rosen_vec(x):
    return -np.sum(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2, axis=1)


def corr_prec(d, seed=99):
    A = np.random.RandomState(seed).randn(d, d)
    cov = A @ A.T / d + np.eye(d)
    return np.linalg.inv(cov)


def run_case(name, seed, ndim, nwalkers, ntemps, nits, like, like_args, vectorize, lo, hi,
             moves_factory=None, tempering_kwargs=None, periodic=None):
    np.random.seed(seed)
    priors = ProbDistContainer({i: uniform_dist(lo, hi) for i in range(ndim)})
    run_case.priors = priors
    tk = {} if ntemps is None else dict(ntemps=ntemps)
    if tempering_kwargs:
        tk.update(tempering_kwargs)
    moves = moves_factory() if moves_factory else None
    sampler = EnsembleSampler(nwalkers, ndim, like, priors, args=like_args, tempering_kwargs=tk,
                              moves=moves, vectorize=vectorize, periodic=periodic)
    T = sampler.ntemps
    x0 = priors.rvs(size=(T, nwalkers))
    rec = dict(coords=[], logl=[], logp=[], accepted=[], swaps=[], betas=[], move=[])
    prev_acc = [np.zeros((T, nwalkers)) for _ in sampler.moves]
    prev_np = [0 for _ in sampler.moves]

    def nprop(mv):  # a CombineMove counts proposals in its sub-moves only (combine.py:126)
        return sum(m.num_proposals for m in mv.moves) if isinstance(mv, CombineMove) else mv.num_proposals

    def acc_of(mv):  # the sub-moves hold the counters (the reference's own CombineMove.accepted getter reads an attribute
        # its setter never creates, combine.py:35,44-48)
        return np.sum([m.accepted for m in mv.moves], axis=0) if isinstance(mv, CombineMove) else mv.accepted
    state0 = State(x0[:, :, None, :].copy())
    first = True
    for state in sampler.sample(state0, iterations=nits, store=False, skip_initial_state_check=True):
        if first:
            first = False
        which = None
        acc = None
        for k, mv in enumerate(sampler.moves):
            if nprop(mv) != prev_np[k]:
                which = k
                acc = acc_of(mv) - prev_acc[k]
                prev_acc[k] = acc_of(mv).copy()
                prev_np[k] = nprop(mv)
        rec["move"].append(which)
        rec.setdefault("acc_count", []).append(acc.astype(np.uint8))
        rec["accepted"].append(acc.astype(bool))
        rec["coords"].append(state.branches_coords["model_0"].copy())
        rec["logl"].append(state.log_like.copy())
        rec["logp"].append(state.log_prior.copy())
        if sampler.temperature_control is not None and T > 1:
            rec["swaps"].append(np.asarray(sampler.temperature_control.swaps_accepted).copy())
            rec["betas"].append(sampler.temperature_control.betas.copy())
        else:
            rec["swaps"].append(np.zeros(0))
            rec["betas"].append(np.ones(T))
    out = dict(
        seed=seed, ndim=ndim, nwalkers=nwalkers, ntemps=T, nits=nits, lo=lo, hi=hi,
        tempered=sampler.temperature_control is not None,
        x0=x0,
        coords=np.stack(rec["coords"])[:, :, :, 0, :],
        logl=np.stack(rec["logl"]), logp=np.stack(rec["logp"]),
        accepted=np.packbits(np.stack(rec["accepted"]).astype(np.uint8), axis=-1),
        swaps=np.stack(rec["swaps"]), betas=np.stack(rec["betas"]),
        move=np.asarray(rec["move"], dtype=np.int64),
    )
    if np.stack(rec["acc_count"]).max() > 1:  # combined moves: several proposals per walker and iteration
        out["acc_count"] = np.stack(rec["acc_count"])
    # initial logl/logp the sampler computed
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: T={T} W={nwalkers} d={ndim} its={nits}  coords.sum={out['coords'][-1].sum():.15e} "
          f"acc={np.stack(rec['accepted']).sum()}  size={os.path.getsize(os.path.join(HERE, name + '.npz'))}")
    return sampler, out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "periodic":
        mu3 = np.array([0.2, 6.1, 3.0])
        per = {"model_0": {0: 2 * np.pi, 1: 2 * np.pi}}
        # walkers start uniform over [0, 2 pi): the periodic distance (|c - s| > pi) and the wrap are exercised constantly
        run_case("periodic_mix", 21, 3, 24, 3, 40, ll_gauss_vec, [mu3, np.eye(3) / 0.49], True, 0.0, 2 * np.pi,
                 moves_factory=lambda: [(StretchMove(), 0.5), (GaussianMove({"model_0": 0.25}), 0.5)], periodic=per)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stop_adapt":
        # a fast-moving ladder (adaptation_lag = 30, adaptation_time = 4) that freezes once the adaptation clock reaches
        # stop_adaptation = 6 (tempering.py:571-572, :590)
        run_case("stop_adapt", 19, 3, 40, 5, 15, ll_gauss_vec, [np.zeros(3), np.eye(3)], True, -5.0, 5.0,
                 tempering_kwargs=dict(stop_adaptation=6, adaptation_lag=30, adaptation_time=4))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nosplit":
        # StretchMove(randomize_split=False) (red_blue.py:123): even / odd walkers, no shuffle drawn from the global stream
        run_case("nosplit", 23, 4, 30, 3, 20, ll_gauss_vec, [np.zeros(4), np.eye(4)], True, -5.0, 5.0,
                 moves_factory=lambda: StretchMove(randomize_split=False))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tiny":
        # edge cases of the shapes: fewer walkers than 2 ndim (live_dangerously=True, red_blue.py:103-114), 3 walkers per
        # half; and a tempered sampler with ONE temperature (no swap, no adaptation: tempering.py:515, :632)
        run_case("tiny_live", 31, 5, 6, 3, 25, ll_gauss_vec, [np.zeros(5), np.eye(5)], True, -5.0, 5.0,
                 moves_factory=lambda: StretchMove(live_dangerously=True))
        run_case("one_temp", 37, 4, 20, 1, 20, ll_gauss_vec, [np.zeros(4), np.eye(4)], True, -5.0, 5.0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "distgen":
        # prior-draw Metropolis move (distgen.py) mixed with the stretch move; narrow box so that prior draws get accepted
        run_case("distgen_mix", 33, 3, 24, 3, 40, ll_gauss_vec, [np.zeros(3), np.eye(3)], True, -2.0, 2.0,
                 moves_factory=lambda: [(StretchMove(), 0.5),
                                        (DistributionGenerate({"model_0": run_case.priors}), 0.5)])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "combine":
        # CombineMove (combine.py): a stretch move then a Gaussian move per iteration, each with its own tempering tail
        run_case("combine_sg", 57, 3, 24, 3, 30, ll_gauss_vec, [np.zeros(3), np.eye(3)], True, -5.0, 5.0,
                 moves_factory=lambda: [(CombineMove([StretchMove(), GaussianMove({"model_0": 0.25})]), 1.0)])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gauss_modes":
        # GaussianMove(mode="random" | "sequential", factor) (gaussian.py:134-181): one random / the next dimension per
        # call, proposal scale multiplied by exp(U(-log factor, log factor))
        run_case("gauss_modes", 41, 4, 24, 3, 40, ll_gauss_vec, [np.zeros(4), corr_prec(4)], True, -5.0, 5.0,
                 moves_factory=lambda: [(GaussianMove({"model_0": 0.16}, mode="random", factor=2.0), 0.5),
                                        (GaussianMove({"model_0": 0.25}, mode="sequential"), 0.3),
                                        (GaussianMove({"model_0": 0.04}, mode="vector"), 0.2)])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mt":
        # multiple-try Metropolis with an independent proposal (multipletry.py:238-514 + mtdistgen.py): num_try draws from
        # the priors per walker, one chosen by importance weight, balanced against the auxiliary set; mixed with stretch
        run_case("mt_mix", 23, 3, 16, 3, 30, ll_gauss_vec, [np.zeros(3), np.eye(3) / 0.25], True, -2.0, 2.0,
                 moves_factory=lambda: [(StretchMove(), 0.4), (MTDistGenMove(run_case.priors, num_try=6, independent=True), 0.6)])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gibbs":
        # Gibbs splits at the parameter level (moves/move.py:113-402): a stretch move that updates parameters {0,1} then
        # {2,3}, and a Gaussian move over {0}, {1,2} and the whole leaf — the splits of one move run inside ONE propose
        # call, each with its own draws, Metropolis test and update, and one tempering tail at the end
        def masks(*sets):
            out = []
            for s_ in sets:
                m = np.zeros((1, 4), dtype=bool)
                m[0, list(s_)] = True
                out.append(("model_0", m))
            return out
        run_case("gibbs_mix", 19, 4, 24, 3, 30, ll_gauss_vec, [np.zeros(4), corr_prec(4)], True, -5.0, 5.0,
                 moves_factory=lambda: [(StretchMove(gibbs_sampling_setup=masks((0, 1), (2, 3))), 0.5),
                                        (GaussianMove({"model_0": 0.09},
                                                      gibbs_sampling_setup=masks((0,), (1, 2)) + ["model_0"]), 0.5)])
        sys.exit(0)
    run_case("c1_kat1", 42, 5, 32, None, 100, ll_single, [np.zeros(5), np.eye(5)], False, -5.0, 5.0)
    run_case("pt_kat2", 42, 3, 16, 4, 50, ll_single, [np.zeros(3), np.eye(3)], False, -5.0, 5.0)
    P8 = corr_prec(8)
    run_case("c2_small", 1234, 8, 64, 4, 30, ll_gauss_vec, [np.zeros(8), P8], True, -10.0, 10.0)
    run_case("c2_tightprior", 77, 8, 64, 4, 30, ll_gauss_vec, [np.zeros(8), P8], True, -1.5, 1.5)
    run_case("c3_small", 5, 4, 32, 3, 40, ll_rosen_vec, None, True, -10.0, 10.0,
             moves_factory=lambda: [(StretchMove(), 0.5), (GaussianMove({"model_0": 0.01}), 0.5)])
    cov3 = np.array([[0.04, 0.01, 0.0], [0.01, 0.09, -0.02], [0.0, -0.02, 0.01]])
    run_case("gauss_matrix", 11, 3, 24, 3, 30, ll_gauss_vec, [np.zeros(3), np.eye(3)], True, -5.0, 5.0,
             moves_factory=lambda: GaussianMove({"model_0": cov3}))
    run_case("odd_walkers", 3, 5, 99, None, 20, ll_gauss_vec, [np.zeros(5), np.eye(5)], True, -5.0, 5.0,
             moves_factory=lambda: StretchMove(a=1.5))
    run_case("noadapt_noperm", 8, 3, 32, 4, 25, ll_gauss_vec, [np.zeros(3), np.eye(3)], True, -5.0, 5.0,
             tempering_kwargs=dict(adaptive=False, permute=False))
