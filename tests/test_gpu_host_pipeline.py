"""The wavefront schedule of eb_run_host (csrc/host_job.cu) and the rung-range form of the swap pass
(eb_pt_swap_range / eb_pt_swap_finish, csrc/k_swap.cu) against the plain sequence: bit-identical results.
Reference semantics: tempering.py:484-561 (hot -> cold ladder walk), :563-596 (adapt_temps)."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _swap_setup(T, W, LD, seed):
    from eryn_b200 import _lib
    lib = _lib.require_device()
    r = np.random.RandomState(seed)
    dev = torch.device("cuda", 0)
    coords = torch.from_numpy(r.randn(T, W, 1, LD)).to(dev)
    logl = torch.from_numpy(-5.0 * r.rand(T, W) * np.arange(1, T + 1)[:, None]).to(dev)
    logp = torch.from_numpy(r.randn(T, W)).to(dev)
    betas = torch.from_numpy(np.geomspace(1.0, 1e-3, T)).to(dev)
    return lib, _lib, dev, coords, logl, logp, betas


@pytest.mark.parametrize("T,W,LD,cuts", [
    (2, 77, 3, []), (5, 300, 8, [3]), (16, 1000, 8, [14, 12, 10, 8, 6, 4, 2]), (16, 513, 20, [15, 7, 6]),
    (33, 129, 5, [20, 19, 1]), (7, 64, 32, [4]), (12, 4096, 8, [9, 6, 3]),
])
@pytest.mark.parametrize("permute", [1, 0])
def test_rung_ranges_equal_one_pass(T, W, LD, cuts, permute):
    """eb_pt_swap_range over adjoining ranges + eb_pt_swap_finish == one eb_pt_swap (same chains, positions, uniforms)"""
    lib, _lib, dev, coords, logl, logp, betas = _swap_setup(T, W, LD, 100 + T)
    adapt = _lib.eb_adapt(1, -1, 10.0, 5.0)   # strong adaptation: the ladder visibly moves every pass
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(ranged):
        c, ll, lp, b = coords.clone(), logl.clone(), logp.clone(), betas.clone()
        ctrl = torch.zeros(lib.eb_ctrl_size(), dtype=torch.uint8, device=dev)
        st = _lib.eb_state(T, W, 1, LD, 0, 0, _ptr(c), _ptr(ll), _ptr(lp), None, _ptr(b))
        scr_c, scr_p = torch.zeros_like(c), torch.zeros_like(lp)
        rng = _lib.eb_swap_rng()
        rng.mode, rng.permute, rng.seed, rng.iter_dev = 1, permute, 991, ctypes.c_void_p(ctrl.data_ptr())
        rng.row_scratch, rng.logp_scratch = _ptr(scr_c), _ptr(scr_p)
        hist = []
        for it in range(3):
            if not ranged:
                _lib.check(lib.eb_pt_swap(ctypes.byref(st), ctypes.byref(rng), ctypes.byref(adapt), _ptr(ctrl), stream), "swap")
            else:
                edges = [T - 1] + list(cuts) + [0]
                for hi, lo in zip(edges[:-1], edges[1:]):
                    _lib.check(lib.eb_pt_swap_range(ctypes.byref(st), ctypes.byref(rng), _ptr(ctrl), hi, lo, stream), "range")
                _lib.check(lib.eb_pt_swap_finish(ctypes.byref(st), ctypes.byref(rng), ctypes.byref(adapt), _ptr(ctrl), stream), "finish")
            torch.cuda.synchronize()
            hist.append((c.cpu().numpy().copy(), ll.cpu().numpy().copy(), lp.cpu().numpy().copy(), b.cpu().numpy().copy(),
                         ctrl.cpu().numpy().copy()))
        return hist

    one, rng_ = run(False), run(True)
    for it, (a, bb) in enumerate(zip(one, rng_)):
        for k, name in enumerate(("coords", "logl", "logp", "betas", "ctrl")):
            assert np.array_equal(a[k], bb[k]), f"{name} differs after pass {it}"
    assert not np.array_equal(one[0][0], coords.cpu().numpy())   # something did move
    assert not np.array_equal(one[-1][3], betas.cpu().numpy()) or T < 3   # and the ladder adapted


def _host_job(_lib, T, W, d, like_kind, par, ncomp, arrays, swaps, cnt, seed):
    coords, logl, logp, betas, lo, hi = arrays
    job = _lib.eb_host_job()
    job.ntemps, job.nwalkers, job.nleaves, job.ndim = T, W, 1, d
    job.coords_host, job.logl_host, job.logp_host, job.betas_host = [_ptr(t) for t in (coords, logl, logp, betas)]
    job.prior_lo_host, job.prior_hi_host = ctypes.c_void_p(lo.ctypes.data), ctypes.c_void_p(hi.ctypes.data)
    job.like_kind, job.like_ncomp, job.like_nparams = like_kind, ncomp, par.size
    job.like_params_host = ctypes.c_void_p(par.ctypes.data) if par.size else None
    job.stretch_a, job.gauss_scale, job.seed, job.iter0 = 2.0, 0.1, seed, 0
    job.adapt = _lib.eb_adapt(1, -1, 50.0, 10.0)
    job.adapt_time0, job.permute, job.randomize_split = 0, 1, 1
    job.swaps_accepted_host, job.accepted_count_host = _ptr(swaps), _ptr(cnt)
    return job


@pytest.mark.parametrize("T,W,d,kind,groups", [
    (16, 512, 8, 0, 0), (16, 512, 8, 0, 16), (5, 130, 8, 0, 3), (8, 256, 20, 2, 4), (6, 200, 5, 1, 6), (2, 64, 3, 0, 0),
])
def test_wavefront_schedule_equals_plain(T, W, d, kind, groups, monkeypatch):
    """eb_run_host, one iteration per call on pinned arrays: wavefront (direct, then captured graph, stretch and
    Gaussian moves alternating) == plain sequence, bit for bit, over several calls"""
    from eryn_b200 import _lib
    from eryn_b200.likelihood import GaussianLikelihood, GaussianMixtureLikelihood, RosenbrockLikelihood
    lib = _lib.require_device()
    r = np.random.RandomState(7)
    if kind == 0:
        A = r.randn(d, d)
        lk = GaussianLikelihood(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
    elif kind == 1:
        lk = RosenbrockLikelihood()
    else:
        lk = GaussianMixtureLikelihood(r.uniform(-5, 5, size=(4, d)), r.uniform(0.5, 1.5, size=4), np.full(4, 0.25))
    par = np.ascontiguousarray(lk.params(), dtype=np.float64)
    lo, hi = np.full(d, -10.0), np.full(d, 10.0)
    x0 = r.uniform(-3, 3, size=(T, W, 1, d))
    betas0 = np.geomspace(1.0, 1e-2, T)
    sched = np.array([0, 0, 1, 0, 1, 1, 0, 0], dtype=np.uint8)

    def run(mode):
        monkeypatch.setenv("EB_HOST_PIPE", mode)
        monkeypatch.setenv("EB_HOST_GROUPS", str(groups))
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        coords, betas = pin(x0.copy()), pin(betas0.copy())
        logl, logp = pin(np.zeros((T, W))), pin(np.zeros((T, W)))
        swaps, cnt = pin(np.zeros(T - 1, dtype=np.int32)), pin(np.zeros((T, W), dtype=np.uint32))
        # initial logl / logp from the device (eb_eval_state), so both runs start from the same numbers
        dev = torch.device("cuda", 0)
        dc, dl, dp = coords.to(dev), logl.to(dev), logp.to(dev)
        dpr = torch.from_numpy(np.stack([lo, hi, np.log(1.0 / (hi - lo))])).to(dev)
        dlk = torch.from_numpy(par if par.size else np.zeros(1)).to(dev)
        st = _lib.eb_state(T, W, 1, d, 0, 0, _ptr(dc), _ptr(dl), _ptr(dp), None, None)
        pr = _lib.eb_prior(_ptr(dpr[0]), _ptr(dpr[1]), _ptr(dpr[2]), None)
        lkc = _lib.eb_like(int(lk.kind), int(lk.ncomp), int(par.size), 0, _ptr(dlk))
        _lib.check(lib.eb_eval_state(ctypes.byref(st), ctypes.byref(pr), ctypes.byref(lkc),
                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "eval")
        torch.cuda.synchronize()
        logl.copy_(dl.cpu()); logp.copy_(dp.cpu())
        job = _host_job(_lib, T, W, d, int(lk.kind), par, int(lk.ncomp), (coords, logl, logp, betas, lo, hi), swaps, cnt, 4242)
        hist = []
        for i in range(len(sched)):
            job.move_schedule_host = ctypes.c_void_p(sched.ctypes.data + i)
            _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
            hist.append([t.numpy().copy() for t in (coords, logl, logp, betas, swaps, cnt)] + [job.iter0, job.adapt_time0])
        return hist

    plain, wave = run("0"), run("2")
    names = ("coords", "logl", "logp", "betas", "swaps", "accepted_count", "iter", "time")
    for i, (a, b) in enumerate(zip(plain, wave)):
        for k, name in enumerate(names):
            assert np.array_equal(a[k], b[k]), f"{name} differs after call {i}"
    assert plain[-1][6] == len(sched) and plain[-1][5].sum() > 0 and plain[-1][4].sum() > 0


def test_wavefront_needs_pinned_arrays_else_plain(monkeypatch):
    """pageable host arrays silently take the plain schedule (same results)"""
    from eryn_b200 import _lib
    lib = _lib.require_device()
    monkeypatch.setenv("EB_HOST_PIPE", "2")
    T, W, d = 4, 128, 8
    r = np.random.RandomState(1)
    par = np.concatenate([np.zeros(d), np.eye(d).ravel()])
    lo, hi = np.full(d, -10.0), np.full(d, 10.0)
    x0 = r.uniform(-3, 3, size=(T, W, 1, d))
    outs = []
    for pinned in (False, True):
        mk = (lambda a: torch.from_numpy(a).pin_memory()) if pinned else (lambda a: torch.from_numpy(a))
        coords, betas = mk(x0.copy()), mk(np.geomspace(1.0, 0.1, T))
        logl, logp = mk(-0.5 * (x0[:, :, 0] ** 2).sum(-1)), mk(np.full((T, W), d * np.log(1 / 20.0)))
        swaps, cnt = mk(np.zeros(T - 1, dtype=np.int32)), mk(np.zeros((T, W), dtype=np.uint32))
        job = _host_job(_lib, T, W, d, 0, par, 0, (coords, logl, logp, betas, lo, hi), swaps, cnt, 5)
        for _ in range(3):
            _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
        outs.append([t.numpy().copy() for t in (coords, logl, logp, betas, swaps, cnt)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
