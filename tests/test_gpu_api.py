"""The public sampler API on the device-resident fast path (run on the B200 box: `pytest -m gpu`).

`EnsembleSampler.sample` / `run_mcmc` in production (philox) mode replay captured CUDA graphs between yields and store
through the staging ring (eryn_b200/staging.py: pack kernel + asynchronous copy into pinned memory, lazy host State,
deferred Backend.save_step).  Every observable — the stored chain, log_like / log_prior / betas histories, accepted and
swap counters, move acceptance fractions, the returned State — must equal, bit for bit, what the eager path (one launch
sequence per iteration, synchronous download and save_step at every stored step: the path the parity suite checks
against the oracle) produces."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def make_sampler(T, W, d, moves_f, seed=77, like="gauss", inds=False, **kw):
    from eryn_b200 import EnsembleSampler
    from eryn_b200 import likelihood as lk
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    np.random.seed(5)
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    L = lk.GaussianLikelihood(np.zeros(d), cases.corr_prec(d)) if like == "gauss" else lk.RosenbrockLikelihood()
    tk = dict(ntemps=T) if T > 1 else {}
    return EnsembleSampler(W, d, L, pri, tempering_kwargs=tk, moves=moves_f(pri), rng="philox", seed=seed, **kw), pri


def stretch_only(pri):
    from eryn_b200.moves import StretchMove
    return [(StretchMove(a=2.0), 1.0)]


def mix(pri):
    from eryn_b200.moves import GaussianMove, StretchMove
    return [(StretchMove(a=2.0), 0.6), (GaussianMove({"model_0": 0.04}), 0.4)]


def combine(pri):
    from eryn_b200.moves import CombineMove, DistributionGenerate, GaussianMove, StretchMove
    return [(CombineMove([StretchMove(a=2.0), GaussianMove({"model_0": 0.04})]), 0.7),
            (DistributionGenerate({"model_0": pri}), 0.3)]


def run_both(T, W, d, moves_f, nsteps, thin_by, burn=None, **kw):
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(T, W, d))
    out = []
    for eager in (True, False):
        smp, _ = make_sampler(T, W, d, moves_f, **kw)
        smp.force_eager = eager
        last = smp.run_mcmc(x0, nsteps, thin_by=thin_by, burn=burn)
        out.append((smp, last))
    return out


def same_backend(a, b):
    assert a.backend.iteration == b.backend.iteration
    for n in a.branch_names:
        np.testing.assert_array_equal(a.get_chain()[n], b.get_chain()[n])
    np.testing.assert_array_equal(a.get_log_like(), b.get_log_like())
    np.testing.assert_array_equal(a.get_log_prior(), b.get_log_prior())
    np.testing.assert_array_equal(a.get_betas(), b.get_betas())
    np.testing.assert_array_equal(a.backend.accepted, b.backend.accepted)
    np.testing.assert_array_equal(a.backend.swaps_accepted, b.backend.swaps_accepted)
    for k in a.backend.move_info:
        np.testing.assert_array_equal(a.backend.move_info[k]["acceptance_fraction"],
                                      b.backend.move_info[k]["acceptance_fraction"])


@pytest.mark.parametrize("moves_f,thin_by,burn", [(stretch_only, 1, None), (stretch_only, 7, 5), (stretch_only, 40, None),
                                                  (mix, 3, 4), (combine, 2, None)])
def test_resident_path_equals_eager_path(moves_f, thin_by, burn):
    (se, le), (sf, lf) = run_both(4, 128, 8, moves_f, nsteps=6, thin_by=thin_by, burn=burn)
    same_backend(se, sf)
    np.testing.assert_array_equal(le.branches_coords["model_0"], lf.branches_coords["model_0"])
    np.testing.assert_array_equal(le.log_like, lf.log_like)
    np.testing.assert_array_equal(le.betas, lf.betas)
    for me, mf in zip(se.moves, sf.moves):
        np.testing.assert_array_equal(np.asarray(me.accepted), np.asarray(mf.accepted))
    assert type(lf).__name__ == "LazyState"


def test_resident_path_untempered_and_rosenbrock():
    (se, le), (sf, lf) = run_both(1, 64, 8, mix, nsteps=8, thin_by=3, like="rosen")
    same_backend(se, sf)
    np.testing.assert_array_equal(le.log_like, lf.log_like)


def test_staged_store_survives_ring_wraparound_and_late_reads():
    """more stored steps than ring slots, states kept by the consumer and read only after the run"""
    smp, _ = make_sampler(3, 64, 8, stretch_only)
    ref, _ = make_sampler(3, 64, 8, stretch_only)
    ref.force_eager = True
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(3, 64, 8))
    kept = list(smp.sample(x0, iterations=11, thin_by=2))
    kept_ref = list(ref.sample(x0, iterations=11, thin_by=2))
    same_backend(ref, smp)
    for a, b in zip(kept_ref, kept):   # the eager path yields snapshots as well (a new State per yield)
        np.testing.assert_array_equal(a.branches_coords["model_0"], b.branches_coords["model_0"])
        np.testing.assert_array_equal(a.log_like, b.log_like)
        np.testing.assert_array_equal(a.betas, b.betas)


def test_update_fn_every_inner_iteration_and_in_place_edits():
    """ensemble.py:1030-1036: update_fn runs when (i+1) % update_iterations == 0 for the INNER iteration counter i, also
    inside thin_by blocks, gets i, and its in-place edits of the state carry into the chain"""
    calls = {}

    def make(eager):
        seen = []

        def update(i, state, sampler):
            seen.append(i)
            state.branches["model_0"].coords[:, :3] *= 0.5          # in-place edit of the walkers
            state.log_like[:, :3] = -1.0
            b = state.betas.copy()
            b[-1] *= 0.9
            state.betas = b                                            # and of the ladder
        smp, _ = make_sampler(3, 64, 8, stretch_only, update_fn=update, update_iterations=5)
        smp.force_eager = eager
        x0 = np.random.RandomState(3).uniform(-3, 3, size=(3, 64, 8))
        last = smp.run_mcmc(x0, 6, thin_by=4)
        calls[eager] = seen
        return smp, last
    se, le = make(True)
    sf, lf = make(False)
    assert calls[True] == calls[False] == [4, 9, 14, 19]
    same_backend(se, sf)
    np.testing.assert_array_equal(le.branches_coords["model_0"], lf.branches_coords["model_0"])


def test_loop_body_edits_of_the_yielded_state_carry_on():
    outs = []
    for eager in (True, False):
        smp, _ = make_sampler(2, 64, 8, stretch_only)
        smp.force_eager = eager
        x0 = np.random.RandomState(3).uniform(-3, 3, size=(2, 64, 8))
        for k, st in enumerate(smp.sample(x0, iterations=5, thin_by=2)):
            if k == 2:
                st.branches["model_0"].coords[0, :5] = 0.25
        outs.append(smp)
    same_backend(*outs)


def test_supplied_log_like_is_kept_when_only_the_prior_is_missing():
    """ensemble.py:898-912 evaluates only the missing quantity"""
    from eryn_b200.state import State
    smp, _ = make_sampler(2, 64, 8, stretch_only)
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(2, 64, 8))
    ll = np.full((2, 64), -3.25)
    st = next(smp.sample(State(x0, log_like=ll), iterations=1, store=False, thin_by=1))
    # one iteration later the walkers that did not move still carry the supplied value
    assert np.any(st.log_like == -3.25)


def test_c2_run_mcmc_throughput_path_is_graph_replay():
    """full config-2 size through run_mcmc(thin_by=25): the launches per iteration are the three hot-path kernels (plus
    one pack kernel per stored step) — no torch elementwise kernels, no per-iteration host work"""
    import torch
    smp, _ = make_sampler(16, 4096, 8, stretch_only)
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(16, 4096, 8))
    smp.run_mcmc(x0, 2, thin_by=25)          # warm-up + capture
    l0 = smp.ctx.launches
    torch.cuda.synchronize()
    smp.run_mcmc(None, 4, thin_by=25)
    per_it = (smp.ctx.launches - l0) / 100.0
    assert per_it <= 3.1, per_it
    assert smp.backend.iteration == 6
