"""The config-5 oracle (oracle/rj_oracle.py) against golden vectors recorded from the unmodified reference."""
import numpy as np
import pytest

from oracle import rj_oracle as rjo
from tests import cases_rj


@pytest.mark.parametrize("name", cases_rj.NAMES)
def test_rj_oracle_reproduces_reference(name):
    g = cases_rj.load(name)
    t, coords, inds, private, glob = cases_rj.replay_setup(g)
    smp = cases_rj.oracle_sampler(g, rjo.NumpyStreamsMB(private, glob), t)
    st = smp.initialise(rjo.MBState(coords, inds))
    np.testing.assert_allclose(st.logl, g["logl0"], rtol=1e-13)
    np.testing.assert_array_equal(st.logp, g["logp0"])
    for it in range(int(g["nits"])):
        acc, racc = smp.iterate(st)
        assert np.array_equal(acc, g["acc"][it]), f"in-model accept mask differs at iteration {it}"
        assert np.array_equal(racc, g["rjacc"][it]), f"rj accept mask differs at iteration {it}"
        if "rjmove" in g:
            assert smp.last_rj_move == int(g["rjmove"][it]), f"rj move choice differs at iteration {it}"
        assert np.array_equal(st.inds[0], g["ig"][it]) and np.array_equal(st.inds[1], g["is"][it]), f"inds it {it}"
        np.testing.assert_allclose(st.coords[0], g["cg"][it], rtol=1e-13, atol=1e-300, err_msg=f"gauss coords it {it}")
        np.testing.assert_allclose(st.coords[1], g["cs"][it], rtol=1e-13, atol=1e-300, err_msg=f"sine coords it {it}")
        np.testing.assert_allclose(st.logl, g["logl"][it], rtol=1e-13, err_msg=f"logl it {it}")
        np.testing.assert_allclose(st.logp, g["logp"][it], rtol=1e-13, err_msg=f"logp it {it}")
        np.testing.assert_allclose(smp.betas, g["betas"][it], rtol=1e-13, err_msg=f"betas it {it}")
        # the reference reports the swap counts of the LAST pass (the rj move's): tempering.py:500
        assert np.array_equal(smp.swaps_accepted, g["swaps"][it]), f"swaps it {it}"


def test_philox_mb_streams_are_valid():
    """production-mode draws of the config-5 kernels: picks in range, slots valid, births inside the prior"""
    g = cases_rj.load("c5_small")
    t, coords, inds, _, _ = cases_rj.replay_setup(g)
    smp = cases_rj.oracle_sampler(g, rjo.PhiloxStreamsMB(99), t)
    st = smp.initialise(rjo.MBState(coords, inds))
    nl0 = [i.sum() for i in st.inds]
    tot = 0
    for it in range(12):
        acc, racc = smp.iterate(st)
        tot += acc.sum() + racc.sum()
        for b in range(2):
            n = st.inds[b].sum(axis=-1)
            assert n.min() >= 0 and n.max() <= st.inds[b].shape[2]
            c = st.coords[b][st.inds[b]]
            assert np.all(c >= smp.priors[b].lo) and np.all(c <= smp.priors[b].hi)
    assert tot > 0 and [i.sum() for i in st.inds] != nl0
