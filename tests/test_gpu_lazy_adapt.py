"""Lazy ladder adaptation (eb_swap_rng.defer_adapt / eb_stretch_rng.lazy_ctrl / eb_adapt_flush): a swap pass that leaves
adapt_temps (tempering.py:563-596) and its bookkeeping (:598-649) to the next stretch kernel must give the same chain, the
same ladder and the same counters, bit for bit, as the pass that adapts in its own kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _ctx(T, W, d, kind, seed=21):
    from eryn_b200.device import DeviceContext
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    from eryn_b200.state import State
    from tests.test_gpu_resident import _like
    r = np.random.RandomState(5)
    ctx = DeviceContext(ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)}), _like(kind, d, r),
                        rng="philox", seed=seed)
    x0 = r.uniform(-3, 3, size=(T, W, 1, d))
    ds = ctx.upload(State({"model_0": x0}), betas=torch.from_numpy(np.geomspace(1.0, 1e-2, T)).to(ctx.device))
    ctx.eval_state(ds)
    return ctx, ds


def _snap(ctx, ds, cnt, T):
    torch.cuda.synchronize()
    ctx.check_error()
    c = ctx.read_ctrl()   # (flushes a deferred adaptation)
    return dict(coords=ds.coords.cpu().numpy().copy(), logl=ds.logl.cpu().numpy().copy(), logp=ds.logp.cpu().numpy().copy(),
                betas=ds.betas.cpu().numpy().copy(), cnt=cnt.cpu().numpy().copy(), iter=int(c.iter), inext=int(c.iter_next),
                time=int(c.time), swaps=np.array(c.swaps_accepted[:T - 1]), total=np.array(c.swaps_total[:T - 1]),
                work=np.array([list(row) for row in c.swaps_work]).sum(), ticket=int(c.ticket))


def _run(T, W, d, kind, plan, lazy, adapt):
    """plan: list of ('it', n) = n iterations of stretch + pass, ('snap',) = look at everything, ('gauss',) = one Gaussian
    step + pass in between"""
    ctx, ds = _ctx(T, W, d, kind)
    ctx.lazy_adapt = lazy
    cnt = torch.zeros((T, W), dtype=torch.int32, device=ctx.device)
    out = []
    for step in plan:
        if step[0] == "it":
            for _ in range(step[1]):
                ctx.stretch_step(ds, 2.0, accepted_count=cnt)
                ctx.pt_swap(ds, adapt=adapt)
        elif step[0] == "gauss":
            ctx.gaussian_step(ds, dict(kind="scalar", scale=0.05), accepted_count=cnt)
            ctx.pt_swap(ds, adapt=adapt)
        else:
            out.append(_snap(ctx, ds, cnt, T))
    out.append(_snap(ctx, ds, cnt, T))
    return out


ADAPT = dict(adaptive=True, stop_adaptation=-1, adaptation_lag=50.0, adaptation_time=10.0)


@pytest.mark.parametrize("T,W,d,kind", [
    (16, 4096, 8, 0),        # config 2: one launch per half
    (4, 64, 5, 1),           # both halves in one CTA per temperature
    (5, 300, 8, 0),
    (2, 64, 3, 0), (3, 96, 8, 2),
    (32, 16384, 20, 2),      # config 4: lane-split kernel, more CTAs than fit at once
    (64, 512, 8, 0),         # long ladder
])
def test_lazy_adaptation_equals_adaptation_in_the_pass(T, W, d, kind):
    plan = [("it", 1), ("snap",), ("it", 3), ("snap",), ("it", 2)]
    ref = _run(T, W, d, kind, plan, False, ADAPT)
    lazy = _run(T, W, d, kind, plan, True, ADAPT)
    for i, (a, b) in enumerate(zip(ref, lazy)):
        for k in a:
            assert np.array_equal(a[k], b[k]), f"{k} differs at snapshot {i}"
    assert ref[-1]["iter"] == 6 and ref[-1]["total"].sum() > 0 and ref[-1]["work"] == 0
    assert T < 3 or not np.array_equal(ref[-1]["betas"], ref[0]["betas"])


@pytest.mark.parametrize("adapt", [None, dict(adaptive=False, stop_adaptation=-1, adaptation_lag=50.0, adaptation_time=10.0),
                                   dict(adaptive=True, stop_adaptation=2, adaptation_lag=50.0, adaptation_time=10.0)])
def test_lazy_adaptation_options_and_other_moves_in_between(adapt):
    plan = [("it", 2), ("gauss",), ("it", 2), ("snap",), ("gauss",), ("gauss",), ("it", 3)]
    ref = _run(8, 512, 8, 0, plan, False, adapt)
    lazy = _run(8, 512, 8, 0, plan, True, adapt)
    for i, (a, b) in enumerate(zip(ref, lazy)):
        for k in a:
            assert np.array_equal(a[k], b[k]), f"{k} differs at snapshot {i}"


def test_sampler_takes_lazy_adaptation_and_matches_eager(monkeypatch):
    """run_mcmc on a plain stretch sampler: the resident path runs with lazy adaptation (every pass deferred) and equals
    the eager path and the resident path without it"""
    from tests.test_gpu_api import make_sampler, same_backend, stretch_only
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(8, 256, 8))
    runs = {}
    for mode in ("eager", "resident", "lazy"):
        monkeypatch.setenv("EB_LAZY_ADAPT", "1" if mode == "lazy" else "0")
        smp, _ = make_sampler(8, 256, 8, stretch_only)
        smp.force_eager = mode == "eager"
        seen = []
        for st in smp.sample(x0, iterations=6, thin_by=7):
            seen.append((st.betas.copy(), smp.temperature_control.swaps_accepted.copy(), smp.temperature_control.time))
        runs[mode] = (smp, seen)
    for mode in ("resident", "lazy"):
        same_backend(runs["eager"][0], runs[mode][0])
        for a, b in zip(runs["eager"][1], runs[mode][1]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


@pytest.mark.parametrize("moves_name", ["stretch_only", "mix"])
def test_lazy_adaptation_across_calls_that_only_replay_graphs(moves_name, monkeypatch):
    """a second sample() call whose iterations are all graph replays (no pass goes through DeviceContext.pt_swap) still
    flushes at every yield: stored ladders, clocks and swap counts equal the eager path"""
    from tests import test_gpu_api as api
    moves_f = getattr(api, moves_name)
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(4, 128, 8))
    runs = {}
    for mode in ("eager", "lazy"):
        monkeypatch.setenv("EB_LAZY_ADAPT", "1" if mode == "lazy" else "0")
        smp, _ = api.make_sampler(4, 128, 8, moves_f)
        smp.force_eager = mode == "eager"
        smp.run_mcmc(x0, 4, thin_by=3, burn=3)
        seen = []
        for st in smp.sample(smp.get_last_sample(), iterations=5, thin_by=3):       # graphs of this block size exist already
            seen.append((st.betas.copy(), smp.temperature_control.time, smp.temperature_control.swaps_accepted.copy()))
        runs[mode] = (smp, seen)
    api.same_backend(runs["eager"][0], runs["lazy"][0])
    for a, b in zip(runs["eager"][1], runs["lazy"][1]):
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2])
