"""Pin the oracle: replay the seeds of tests/golden/*.npz (made by the unmodified reference)
through oracle/eryn_oracle.py.  Accept masks, swap counts and move choices must be bit-equal;
floats to 1e-13 relative (they are in fact bit-equal on the build box)."""
import hashlib

import numpy as np
import pytest

from oracle import eryn_oracle as orc
from tests import cases


def make_oracle(name, g, streams):
    c = cases.CASES[name]
    d = int(g["ndim"])
    prior = orc.BoxPrior(np.full(d, float(g["lo"])), np.full(d, float(g["hi"])))
    T = int(g["ntemps"])
    betas = None
    if bool(g["tempered"]):
        betas = orc.make_ladder_default(d, T) if T > 1 else np.array([1.0])
    return orc.OracleSampler(prior, c["like"](d), c["moves"], c.get("weights", [1.0]), streams, betas=betas,
                             adaptive=c.get("adaptive", True), permute=c.get("permute", True),
                             periods=c.get("periods"), **c.get("tempering", {})), prior


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_reference(name):
    g = cases.load(name)
    private, glob = cases.seeded_streams(int(g["seed"]))
    smp, prior = make_oracle(name, g, orc.NumpyStreams(private, glob))
    T, W = int(g["ntemps"]), int(g["nwalkers"])
    x0 = prior.rvs((T, W), glob)  # prior.py:64 draws from the global stream
    assert np.array_equal(x0, g["x0"])
    state = smp.initialise(orc.OState(x0))
    for it in range(int(g["nits"])):
        acc = smp.iterate(state)
        assert smp.last_move == g["move"][it], f"move choice differs at it {it}"
        if "gibbs" in smp.moves[smp.last_move]:  # what the reference added to move.accepted over the Gibbs splits
            acc = smp.last_accept_sum
        assert np.array_equal(acc, g["accepted"][it]), f"accept mask differs at it {it}"
        np.testing.assert_allclose(state.coords[:, :, 0, :], g["coords"][it], rtol=1e-13, atol=0)
        np.testing.assert_allclose(state.logl, g["logl"][it], rtol=1e-13, atol=0)
        np.testing.assert_allclose(state.logp, g["logp"][it], rtol=1e-13, atol=0)
        if T > 1:
            assert np.array_equal(smp.swaps_accepted, g["swaps"][it]), f"swap counts differ at it {it}"
            np.testing.assert_allclose(smp.betas, g["betas"][it], rtol=1e-13, atol=0)


def test_known_answers_appendix_c():
    """SURVEY.md Appendix C (generated from the reference in the survey session)."""
    g = cases.load("c1_kat1")
    c = g["coords"][-1][:, :, None, :]
    assert abs(c.sum() - (-6.863121888547821e-01)) < 1e-12
    assert abs(g["logl"][-1].sum() - (-6.935659199141485e01)) < 1e-10
    assert g["accepted"].sum() == 1757
    assert hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:16] == "f23ffcb2aeceb718"
    g = cases.load("pt_kat2")
    assert abs(g["coords"][-1].sum() - 2.892038585097532e01) < 1e-11
    assert g["accepted"].sum() == 1859
    assert np.array_equal(g["swaps"].sum(0), [228, 436, 677])
    np.testing.assert_allclose(g["betas"][-1], [1, 0.24676878824520287, 0.05735976075520548, 0.01115873741850507],
                               rtol=1e-14)


def test_ladders_match_reference():
    """tests/golden/ladders.npz (make_golden_ladders.py ran the unmodified reference): make_ladder on a grid of
    (ndim, ntemps, Tmax) for the host mirror and the oracle, and adapt_temps driven by recorded swap counts."""
    import os
    from eryn_b200.moves import make_ladder
    g = np.load(os.path.join(cases.GOLDEN, "ladders.npz"))
    for k, (ndim, ntemps, tmax) in enumerate(g["grid"]):
        ref = g[f"ladder_{k}"]
        got = make_ladder(int(ndim), ntemps=int(ntemps), Tmax=None if tmax < 0 else float(tmax))
        np.testing.assert_allclose(got, ref, rtol=1e-14, atol=0, err_msg=f"make_ladder({ndim}, {ntemps}, {tmax})")
        if tmax < 0:
            np.testing.assert_allclose(orc.make_ladder_default(int(ndim), int(ntemps)), ref, rtol=1e-14, atol=0)
    for k in range(3):
        T, W, lag, t0 = g[f"adapt_{k}_cfg"]
        betas = g[f"adapt_{k}_betas0"].copy()
        for step, counts in enumerate(g[f"adapt_{k}_counts"]):
            betas = orc.adapt_temps(betas, counts.astype(float), int(W), step, adaptation_lag=lag, adaptation_time=t0)
            np.testing.assert_allclose(betas, g[f"adapt_{k}_hist"][step], rtol=1e-13, atol=0,
                                       err_msg=f"adapt_temps cfg {k} step {step}")
