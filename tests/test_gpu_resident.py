"""K12 resident_kernel (csrc/resident.cuh): whole iterations in one launch with the state in shared memory must equal
the per-launch kernels (eb_stretch_step + eb_pt_swap, which the parity suite checks against the oracle) bit for bit.
Reference loop: ensemble.py:965-1045; move red_blue.py:89-333; pass tempering.py:484-649."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _like(kind, d, r):
    from eryn_b200.likelihood import GaussianLikelihood, GaussianMixtureLikelihood, RosenbrockLikelihood
    if kind == 0:
        A = r.randn(d, d)
        return GaussianLikelihood(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
    if kind == 1:
        return RosenbrockLikelihood()
    return GaussianMixtureLikelihood(r.uniform(-5, 5, size=(4, d)), r.uniform(0.5, 1.5, size=4), np.full(4, 0.25))


def _run(T, W, d, kind, blocks, resident, permute=True, adaptive=True, seed=11):
    from eryn_b200.device import DeviceContext
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    from eryn_b200.state import State
    r = np.random.RandomState(5)
    ctx = DeviceContext(ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)}), _like(kind, d, r),
                        rng="philox", seed=seed)
    x0 = r.uniform(-3, 3, size=(T, W, 1, d))
    betas = torch.from_numpy(np.geomspace(1.0, 1e-2, T)).to(ctx.device)
    ds = ctx.upload(State({"model_0": x0}), betas=betas)
    ctx.eval_state(ds)
    cnt = torch.zeros((T, W), dtype=torch.int32, device=ctx.device)
    adapt = dict(adaptive=adaptive, stop_adaptation=-1, adaptation_lag=50.0, adaptation_time=10.0)
    hist = []
    for n in blocks:
        if resident:
            acc = ctx.resident_run(ds, 2.0, n, permute=permute, adapt=adapt, accepted_count=cnt)
            assert acc is not False, "shape expected to be covered by the resident kernel"
        else:
            for _ in range(n):
                acc = ctx.stretch_step(ds, 2.0, accepted_count=cnt)
                ctx.pt_swap(ds, permute=permute, adapt=adapt)
        torch.cuda.synchronize()
        ctx.check_error()
        c = ctx.read_ctrl()
        hist.append(dict(coords=ds.coords.cpu().numpy().copy(), logl=ds.logl.cpu().numpy().copy(),
                         logp=ds.logp.cpu().numpy().copy(), betas=ds.betas.cpu().numpy().copy(),
                         acc=acc.cpu().numpy().copy(), cnt=cnt.cpu().numpy().copy(), iter=int(c.iter), time=int(c.time),
                         inext=int(c.iter_next), swaps=np.array(c.swaps_accepted[:T - 1]), total=np.array(c.swaps_total[:T - 1])))
    return hist


@pytest.mark.parametrize("T,W,d,kind,blocks", [
    (16, 4096, 8, 0, [1, 2, 5]),      # config 2: cluster of 4 CTAs per temperature
    (4, 512, 8, 0, [3, 1]),
    (5, 300, 5, 1, [2, 2]),           # padded rows, odd sizes
    (3, 37, 3, 0, [4]),               # one CTA per temperature
    (8, 1000, 12, 2, [2, 1]),         # 16-double bucket, mixture likelihood
    (32, 1024, 8, 0, [2]),            # 32 rungs: one chain per warp
    (2, 64, 8, 1, [3]),
])
def test_resident_equals_per_launch_kernels(T, W, d, kind, blocks):
    ref = _run(T, W, d, kind, blocks, resident=False)
    res = _run(T, W, d, kind, blocks, resident=True)
    for b, (a, x) in enumerate(zip(ref, res)):
        for k in a:
            assert np.array_equal(a[k], x[k]), f"{k} differs after block {b}"
    assert ref[-1]["iter"] == sum(blocks) and ref[-1]["cnt"].sum() > 0 and ref[-1]["total"].sum() > 0
    assert not np.array_equal(ref[-1]["betas"], ref[0]["betas"]) or len(blocks) == 1 or T < 3


@pytest.mark.parametrize("permute,adaptive", [(False, True), (True, False)])
def test_resident_options(permute, adaptive):
    ref = _run(8, 512, 8, 0, [3], resident=False, permute=permute, adaptive=adaptive)
    res = _run(8, 512, 8, 0, [3], resident=True, permute=permute, adaptive=adaptive)
    for k in ref[0]:
        assert np.array_equal(ref[0][k], res[0][k]), k


def test_resident_declines_what_it_does_not_cover():
    from eryn_b200.device import DeviceContext
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    from eryn_b200.state import State
    r = np.random.RandomState(0)
    d = 20
    ctx = DeviceContext(ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)}), _like(2, d, r), rng="philox", seed=1)
    T, W = 32, 16384           # config 4: 92 MB of state
    ds = ctx.upload(State({"model_0": r.uniform(-3, 3, size=(T, W, 1, d))}), betas=torch.from_numpy(np.geomspace(1, 1e-3, T)).to(ctx.device))
    assert ctx.resident_run(ds, 2.0, 0) is False


def test_sampler_uses_the_resident_kernel_for_small_ensembles(monkeypatch):
    """run_mcmc(thin_by=20) on 8 x 256 walkers: blocks of >= 8 iterations go through K12 (one launch per block); the chain
    equals the eager path and the replayed per-launch kernels bit for bit"""
    from tests.test_gpu_api import make_sampler, same_backend, stretch_only
    x0 = np.random.RandomState(3).uniform(-3, 3, size=(8, 256, 8))
    runs = {}
    for mode in ("eager", "graphs", "k12"):
        monkeypatch.setenv("EB_RESIDENT_MAX_WALKERS", "0" if mode == "graphs" else "16384")
        smp, _ = make_sampler(8, 256, 8, stretch_only)
        smp.force_eager = mode == "eager"
        last = smp.run_mcmc(x0, 5, thin_by=20, burn=30)
        runs[mode] = (smp, last, smp.ctx.launches)
    for mode in ("graphs", "k12"):
        same_backend(runs["eager"][0], runs[mode][0])
        for n in ("log_like", "log_prior", "betas"):
            np.testing.assert_array_equal(getattr(runs["eager"][1], n), getattr(runs[mode][1], n))
        np.testing.assert_array_equal(runs["eager"][1].branches_coords["model_0"], runs[mode][1].branches_coords["model_0"])
    assert runs["k12"][0]._k12_ok is True and runs["graphs"][0]._k12_ok in (False, None)
    assert runs["k12"][2] < runs["graphs"][2] / 4      # one launch per block instead of three per iteration
