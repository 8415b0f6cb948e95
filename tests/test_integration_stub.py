"""INTEGRATION.md §2 is executable documentation: the ctypes stub a maintainer would add to Eryn is cut out of the
markdown and run as written — only its two environment lines are rewritten (the library path, and `Move` = a stand-in
with the attributes the stub touches, because the reference itself cannot travel to the GPU box) — against the oracle:
six `propose()` calls on NumPy state arrays must give the oracle's chain (production streams)."""
import os
import re

import numpy as np
import pytest

from oracle import eryn_oracle as orc
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    src = [b for b in blocks if "class B200StretchMove" in b]
    assert len(src) == 1
    return src[0]


def test_stub_is_present_and_declares_the_abi_struct():
    src = stub_source()
    assert "eb_run_host" in src and "_HostJob" in src and "eb_struct_size(8)" in src


@pytest.mark.gpu
def test_integration_stub_runs_and_matches_oracle():
    from eryn_b200 import _lib
    _lib.require_device()
    src = stub_source()
    src = src.replace('from .move import Move          # eryn.moves.move.Move', "")
    src = src.replace('C.CDLL("liberyn_b200.so")', f'C.CDLL({_lib.LIB_PATH!r})')

    class Move(object):  # what eryn.moves.move.Move gives the stub: counters and the temperature control slot
        def __init__(self, temperature_control=None, **kwargs):
            self.temperature_control = temperature_control
            self.accepted = None
            self.num_proposals = 0

    ns = {"Move": Move}
    exec(compile(src, "INTEGRATION.md:stub", "exec"), ns)

    class TC(object):  # the attributes of eryn's TemperatureControl the stub reads / writes (tempering.py:200-283)
        def __init__(self, betas):
            self.betas, self.adaptive, self.stop_adaptation = betas, True, -1
            self.adaptation_lag, self.adaptation_time, self.time, self.permute = 10000, 100, 0, True
            self.swaps_accepted = None

    class Branch(object):
        def __init__(self, coords):
            self.coords = coords

    class HostState(object):
        def __init__(self, coords, logl, logp):
            self.branches = {"model_0": Branch(coords)}
            self.log_like, self.log_prior = logl, logp

    T, W, d, nit, seed = 4, 256, 8, 6, 17
    olike = orc.GaussianLike(np.zeros(d), cases.corr_prec(d))
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    osmp = orc.OracleSampler(prior, olike, [dict(kind="stretch", a=2.0)], [1.0], orc.PhiloxStreams(seed),
                             betas=orc.make_ladder_default(d, T))
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(T, W, d))
    ost = osmp.initialise(orc.OState(x0))
    tc = TC(osmp.betas.copy())
    mv = ns["B200StretchMove"](prior.lo, prior.hi, 0, olike.params(), a=2.0, seed=seed, temperature_control=tc)
    mv.accepted = np.zeros((T, W))
    state = HostState(ost.coords.copy(), ost.logl.copy(), ost.logp.copy())
    total = np.zeros((T, W))
    for _ in range(nit):
        total += osmp.iterate(ost)
        state, acc = mv.propose(None, state)
        assert acc.shape == (T, W)
        assert np.array_equal(tc.swaps_accepted, osmp.swaps_accepted)
    np.testing.assert_allclose(state.branches["model_0"].coords, ost.coords, rtol=1e-10)
    np.testing.assert_allclose(state.log_like, ost.logl, rtol=1e-10)
    np.testing.assert_allclose(tc.betas, osmp.betas, rtol=1e-10)
    assert np.array_equal(mv.accepted, total) and mv.num_proposals == nit and tc.time == nit
