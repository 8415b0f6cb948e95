#!/usr/bin/env python
"""bench.py — walker-updates/sec of the sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl reference]

One "step" = one sampler iteration of the workload: one in-model move (every walker gets one proposal +
Metropolis test) + one parallel-tempering swap pass + ladder adaptation.  N=1 runs BASELINE config 2
(16 temps x 4096 walkers x 8-d correlated Gaussian, StretchMove + PT).  Prints ONE JSON line on rank 0.

  value     device-resident throughput (state already in HBM), production (philox) mode, each step a
            replay of the captured iteration graph, timed by its own CUDA-event pair with an
            L2 flush between steps (outside the pairs);
  e2e       the same step through the C-ABI entry with HOST buffers (eb_run_host): pinned host state ->
            H2D -> kernels -> D2H -> host state, every step;
  roofline  the fused stretch step kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak;
            extra.roofline_swap: the swap pass; extra.roofline_c4: the stretch kernel on the HBM-sized config 4;
  cpu_baseline  the NumPy oracle port of the reference path on this box's host cores, same workload;
  extra.api the same workload through the public sampler API (EnsembleSampler.run_mcmc, thin_by=25, stored);
  extra.c3 / extra.c4 / extra.c5   the other BASELINE configurations (one GPU);
  N > 1     the ladder sharded by temperature, weak scaling of config 2 (16 temperatures x 4096 walkers per GPU);
            extra.c4_strong = config 4 (32 temperatures fixed) sharded over the N GPUs; `parity` = a sharded chain
            against the same chain on one GPU (the run exits non-zero on a mismatch).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "walker-updates/sec (ntemps*nwalkers/iter)"
UNIT = "walker-updates/s"


# ------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------
def corr_prec(d, seed=99):
    A = np.random.RandomState(seed).randn(d, d)
    return np.linalg.inv(A @ A.T / d + np.eye(d))


def workload(name, ngpus=1):
    """returns dict(T, W, d, lo, hi, like=('gauss'|'rosen'|'gmix', params...), moves, weights, label)"""
    if name == "c2":
        # N > 1: weak scaling along the sharded axis — every GPU owns 16 temperatures of 4096 walkers
        T, W, d = 16 * ngpus, 4096, 8
        return dict(T=T, W=W, d=d, lo=-10.0, hi=10.0, like=("gauss", np.zeros(d), corr_prec(d)),
                    moves=[dict(kind="stretch", a=2.0)], weights=[1.0],
                    label=f"C2: {T} temps x {W} walkers x {d}-d correlated Gaussian, StretchMove + PT swaps")
    if name == "c3":
        T, W, d = 16 * ngpus, 4096, 8
        return dict(T=T, W=W, d=d, lo=-10.0, hi=10.0, like=("rosen",),
                    moves=[dict(kind="stretch", a=2.0), dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.1))],
                    weights=[0.5, 0.5],
                    label=f"C3: {T} temps x {W} walkers x {d}-d Rosenbrock, StretchMove/GaussianMove 50/50 + PT")
    if name == "c4":
        T, W, d, K = 32, 16384, 20, 4
        r = np.random.RandomState(5)
        return dict(T=T, W=W, d=d, lo=-10.0, hi=10.0,
                    like=("gmix", r.uniform(-5, 5, size=(K, d)), r.uniform(0.5, 1.5, size=K), np.full(K, 1.0 / K)),
                    moves=[dict(kind="stretch", a=2.0)], weights=[1.0],
                    label=f"C4: {T} temps x {W} walkers x {d}-d {K}-component Gaussian mixture, StretchMove + PT")
    raise SystemExit(f"unknown workload {name}")


def config_of(wl):
    """the `config` object of the JSON line: the same for the GPU arm and the reference arm"""
    return dict(workload=wl["label"], ntemps=wl["T"], nwalkers=wl["W"], ndim=wl["d"])


def oracle_objects(wl):
    from oracle import eryn_oracle as orc
    k = wl["like"]
    like = orc.GaussianLike(k[1], k[2]) if k[0] == "gauss" else orc.RosenbrockLike() if k[0] == "rosen" \
        else orc.GaussianMixtureLike(k[1], k[2], k[3])
    prior = orc.BoxPrior(np.full(wl["d"], wl["lo"]), np.full(wl["d"], wl["hi"]))
    return orc, like, prior


def device_like(wl):
    from eryn_b200 import likelihood as lk
    k = wl["like"]
    return lk.GaussianLikelihood(k[1], k[2]) if k[0] == "gauss" else lk.RosenbrockLikelihood() if k[0] == "rosen" \
        else lk.GaussianMixtureLikelihood(k[1], k[2], k[3])


def initial_coords(wl, seed=1234):
    return np.random.RandomState(seed).uniform(-3.0, 3.0, size=(wl["T"], wl["W"], 1, wl["d"]))


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (NumPy, reference call order, MT19937 streams)
# ------------------------------------------------------------------------------------------------------
def run_cpu(wl, steps, warmup, budget_s=None):
    orc, like, prior = oracle_objects(wl)
    glob = np.random.RandomState(1234)
    private = np.random.RandomState()
    private.set_state(glob.get_state())
    smp = orc.OracleSampler(prior, like, wl["moves"], wl["weights"], orc.NumpyStreams(private, glob),
                            betas=orc.make_ladder_default(wl["d"], wl["T"]))
    st = smp.initialise(orc.OState(initial_coords(wl)))
    for _ in range(warmup):
        smp.iterate(st)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        smp.iterate(st)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return wl["T"] * wl["W"] * done / dt, dt / done, done


_POOL_LIKE = None


def _pool_init(kind, params):
    global _POOL_LIKE
    from oracle import eryn_oracle as orc
    _POOL_LIKE = orc.GaussianLike(*params) if kind == "gauss" else orc.RosenbrockLike() if kind == "rosen" \
        else orc.GaussianMixtureLike(*params)


def _pool_call(x):
    return float(_POOL_LIKE(x[None, :])[0])


def run_cpu_pool(wl, budget_s=20.0):
    """The reference's only multi-core facility (ensemble.py:1474-1481): vectorize=False, the likelihood called once per
    walker through multiprocessing.Pool(ncores).map.  Timed: the likelihood leg of ONE iteration (every walker evaluated
    once) on a bounded number of walkers, on top of the vectorised rest of the iteration."""
    import multiprocessing as mp
    k = wl["like"]
    ncores = os.cpu_count() or 1
    x = initial_coords(wl).reshape(-1, wl["d"])
    n = x.shape[0]
    with mp.get_context("fork").Pool(ncores, initializer=_pool_init, initargs=(k[0], tuple(k[1:]))) as pool:
        pool.map(_pool_call, list(x[:1024]))
        t0 = time.perf_counter()
        m = 0
        chunk = 16384
        while m < n and time.perf_counter() - t0 < budget_s:
            pool.map(_pool_call, list(x[m:m + chunk]))
            m += min(chunk, n - m)
        dt = time.perf_counter() - t0
    return m / dt, ncores, m


def cpu_threads():
    """threads the NumPy port actually computes with: its work is elementwise / einsum-without-BLAS / fancy indexing,
    all single-threaded in NumPy (the BLAS pool, if any, stays idle on this path)"""
    return 1


def blas_pool():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return 1


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields via NVML) — runs during the timed region and a sustained replay of the same step
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        self.stop_flag = True
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples),
                    window="the timed steps plus >= 0.3 s of the same step replayed back to back")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
# one workload on one GPU: context, moves, one captured graph per move kind
# ------------------------------------------------------------------------------------------------------
class SingleGpu(object):
    def __init__(self, wl, dev, seed=20261017):
        import torch
        from eryn_b200.device import DeviceContext
        from eryn_b200.moves import GaussianMove, StretchMove, TemperatureControl
        from eryn_b200.prior import ProbDistContainer, uniform_dist
        from eryn_b200.state import State
        self.torch, self.wl, self.dev = torch, wl, dev
        T, W, d = wl["T"], wl["W"], wl["d"]
        self.pri = ProbDistContainer({i: uniform_dist(wl["lo"], wl["hi"]) for i in range(d)})
        self.ctx = DeviceContext(self.pri, device_like(wl), device=dev, rng="philox", seed=seed)
        self.tc = TemperatureControl(d, W, ntemps=T)
        self.tc.bind(self.ctx)
        # loops of plain stretch proposals: the swap pass leaves its ladder adaptation to the next stretch kernel
        # (DeviceContext.lazy_adapt), as EnsembleSampler's resident path does
        self.ctx.lazy_adapt = (all(m["kind"] in ("stretch", "gaussian") for m in wl["moves"]) and T * W <= 131072
                               and os.environ.get("EB_LAZY_ADAPT", "1") != "0")
        self.moves = []
        for m in wl["moves"]:
            mv = StretchMove(a=m["a"]) if m["kind"] == "stretch" else GaussianMove({"model_0": m["proposal"]["scale"] ** 2})
            mv.temperature_control = self.tc
            mv.bind(self.ctx)
            mv.accepted = np.zeros((T, W))
            self.moves.append(mv)
        self.ds = self.ctx.upload(State(initial_coords(wl)), betas=self.tc.betas_dev)
        self.ctx.eval_state(self.ds)
        torch.cuda.synchronize()
        self.stream = torch.cuda.Stream(device=dev)
        self.graphs = []
        with torch.cuda.stream(self.stream):
            for mv in self.moves:  # warm-up launches outside capture (module load, attribute setup)
                mv.propose(None, self.ds)
            torch.cuda.synchronize()
            for mv in self.moves:
                g = torch.cuda.CUDAGraph()
                l0 = self.ctx.launches
                with torch.cuda.graph(g, stream=self.stream):
                    mv.propose(None, self.ds)
                self.graphs.append((g, self.ctx.launches - l0))
        torch.cuda.synchronize()
        self.stretch = [m for m in self.moves if isinstance(m, StretchMove)]

    def schedule(self, n):
        w = np.asarray(self.wl["weights"], dtype=float)
        return np.random.RandomState(7).choice(len(self.moves), p=w / w.sum(), size=n)

    def timed_steps(self, steps, warmup, flush=None, clk=None):
        """per-step CUDA-event pairs, L2 flushed between steps (outside the pairs); returns (total_ms, launches)"""
        torch = self.torch
        sched = self.schedule(warmup + steps)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(self.stream):
            for i in range(warmup):
                self.graphs[sched[i]][0].replay()
            torch.cuda.synchronize()
            if clk is not None:
                clk.start()
            launches = 0
            for i in range(steps):
                if flush is not None:
                    flush.fill_(i & 0xFF)  # L2 flush between timed iterations, outside the event pair
                g, nl = self.graphs[sched[warmup + i]]
                evs[i][0].record(self.stream)
                g.replay()
                evs[i][1].record(self.stream)
                launches += nl
            torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in evs)), launches

    def resident_ms(self, steps, min_seconds=0.0):
        """the same steps back to back without flush (state stays L2-resident, as in a real run)"""
        torch = self.torch
        sched = self.schedule(steps)
        with torch.cuda.stream(self.stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for i in range(steps):
                self.graphs[sched[i]][0].replay()
            e1.record(self.stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if min_seconds > 0:  # sustained load for the clock sampler
                n = int(min_seconds / max(ms * 1e-3, 1e-7))
                for i in range(n):
                    self.graphs[sched[i % steps]][0].replay()
                torch.cuda.synchronize()
        return ms

    def graph_us(self, body, nrep):
        """average duration of `body` (kernel launches) over nrep back-to-back repetitions inside one graph"""
        torch = self.torch
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.stream(self.stream):
            body()
            torch.cuda.synchronize()
            with torch.cuda.graph(gk, stream=self.stream):
                for _ in range(nrep):
                    body()
            gk.replay()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(self.stream)
            gk.replay()
            k1.record(self.stream)
            torch.cuda.synchronize()
        return k0.elapsed_time(k1) * 1e3 / nrep

    def stretch_us(self, nrep):
        self.ctx.flush_adapt()   # back-to-back stretch steps without a pass in between: nothing may be pending
        cnt = self.stretch[0]._count_buffer(self.ctx, self.wl["T"], self.wl["W"])
        return self.graph_us(lambda: self.ctx.stretch_step(self.ds, 2.0, accepted_count=cnt), nrep)

    def swap_us(self, nrep):
        """average duration of the swap pass alone (back-to-back launches in one graph).  With lazy adaptation the passes
        of such a graph pile their counts up unapplied, so the ladder and the control block are put back afterwards."""
        from eryn_b200 import _lib
        torch = self.torch
        self.ctx.flush_adapt()
        torch.cuda.synchronize()
        betas0, ctrl0 = self.ds.betas.clone(), self.ctx.ctrl.clone()
        us = self.graph_us(lambda: self.tc.temper_comps(self.ds), nrep)
        torch.cuda.synchronize()
        it_now = self.ctx.ctrl[:8].clone()
        self.ctx.ctrl.copy_(ctrl0)
        self.ctx.ctrl[:8].copy_(it_now)                              # the iteration counter keeps running
        o = _lib.eb_ctrl.iter_next.offset
        self.ctx.ctrl[o:o + 8].copy_(it_now)
        self.ds.betas.copy_(betas0)
        torch.cuda.synchronize()
        return us

    def moved_fraction(self):
        """fraction of walkers whose row changes rung in one swap pass (measured on the current state)"""
        torch = self.torch
        before = self.ds.logl.clone()
        with torch.cuda.stream(self.stream):
            self.tc.temper_comps(self.ds)
            torch.cuda.synchronize()
        return float((self.ds.logl != before).double().mean().item())


def roofline_entry(kernel, alg_bytes, us, peak, peak_src, traffic=None, **extra):
    a = alg_bytes / (us * 1e-6) / 1e9
    out = dict(bound="hbm", kernel=kernel, achieved=round(a, 1), peak=peak, unit="GB/s", frac=round(a / peak, 4),
               traffic=traffic, peak_source=peak_src, algorithmic_bytes_per_launch=int(alg_bytes), avg_launch_us=round(us, 3))
    out.update(extra)
    return out


def bench_api(wl, dev, nsteps=40, thin_by=25, store=True):
    """the public sampler API: EnsembleSampler.run_mcmc(x0, nsteps, thin_by) in production mode, stored steps included"""
    import torch
    from eryn_b200 import EnsembleSampler
    from eryn_b200.moves import GaussianMove, StretchMove
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    T, W, d = wl["T"], wl["W"], wl["d"]
    np.random.seed(11)
    pri = ProbDistContainer({i: uniform_dist(wl["lo"], wl["hi"]) for i in range(d)})
    moves = [(StretchMove(a=m["a"]) if m["kind"] == "stretch" else GaussianMove({"model_0": m["proposal"]["scale"] ** 2}), w)
             for m, w in zip(wl["moves"], wl["weights"])]
    smp = EnsembleSampler(W, d, device_like(wl), pri, tempering_kwargs=dict(ntemps=T), moves=moves, rng="philox",
                          seed=20261017, device=dev)
    x0 = initial_coords(wl)[:, :, 0, :]
    smp.run_mcmc(x0, 3, thin_by=thin_by, store=store)  # warm-up: first launches, graph capture
    torch.cuda.synchronize()
    l0 = smp.ctx.launches
    t0 = time.perf_counter()
    smp.run_mcmc(None, nsteps, thin_by=thin_by, store=store)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nit = nsteps * thin_by
    return dict(value=T * W * nit / dt, ms_per_step=dt / nit * 1e3, iterations=nit, thin_by=thin_by, store=store,
                stored_steps=int(smp.backend.iteration), launches_per_iteration=round((smp.ctx.launches - l0) / nit, 3),
                api="EnsembleSampler.run_mcmc(None, %d, thin_by=%d, store=%s): wall clock around the call + a final "
                    "synchronize; CUDA-graph replay between yields, stored steps through the staging ring" % (nsteps, thin_by, store))


def bench_resident(dev, n=25, reps=10):
    """K12 resident_kernel (whole iterations in one launch, state in shared memory; csrc/resident.cuh) against the replayed
    per-launch kernels, us per iteration of blocks of n iterations, device-resident, CUDA events"""
    import torch
    out = {}
    for key, T, W in (("c2", 16, 4096), ("small_8x256", 8, 256)):
        wl = dict(workload("c2"))
        wl["T"], wl["W"] = T, W
        sg = SingleGpu(wl, dev, seed=3)
        mv, tc = sg.moves[0], sg.tc
        cnt = mv._count_buffer(sg.ctx, T, W)
        ad = dict(adaptive=True, stop_adaptation=tc.stop_adaptation, adaptation_lag=tc.adaptation_lag,
                  adaptation_time=tc.adaptation_time)
        per_launch_us = sg.graph_us(lambda: mv.propose(None, sg.ds), n)
        with torch.cuda.stream(sg.stream):
            run = lambda: sg.ctx.resident_run(sg.ds, mv.a, n, randomize_split=mv.randomize_split, permute=tc.permute,
                                              adapt=ad, accepted_count=cnt)
            if run() is False:
                out[key] = dict(supported=False)
                continue
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sg.stream)
            for _ in range(reps):
                run()
            e1.record(sg.stream)
            torch.cuda.synchronize()
        k12_us = e0.elapsed_time(e1) * 1e3 / (reps * n)
        sg.ctx.check_error()
        out[key] = dict(ntemps=T, nwalkers=W, ndim=wl["d"], iterations_per_launch=n, resident_kernel_us_per_iteration=round(k12_us, 2),
                        per_launch_kernels_us_per_iteration=round(per_launch_us, 2),
                        value_resident_kernel=T * W / (k12_us * 1e-6))
        del sg
    out["note"] = ("negative result, kept measured: against a replayed graph of the same block (launches chained by "
                   "programmatic dependent launch) the resident kernel does not win — 64 of 148 SMs, distributed-shared-"
                   "memory gathers at ~20 B/clk per SM, two grid barriers per pass; the sampler uses it only when "
                   "EB_RESIDENT_MAX_WALKERS opts in")
    return out


# ---- config 5: reversible jump + group stretch over two branches (tests/test_eryn.py:38-92, :416-427, :813-907) -----
C5_GINJ = np.array([[3.3, -0.2, 0.1], [2.6, -0.1, 0.1], [3.4, 0.0, 0.1], [2.9, 0.3, 0.1]])


def bench_c5(dev, iters=40, nt=500):
    import torch
    from eryn_b200 import EnsembleSampler
    from eryn_b200.moves import GroupStretchMove
    from eryn_b200.multibranch import PulseLikelihood
    from eryn_b200.prior import uniform_dist
    from eryn_b200.state import State
    T, W, L, nfriends = 8, 2048, 10, 16
    t = np.linspace(-1, 1, nt)
    y = C5_GINJ[:, 0:1].T @ np.exp(-((t[None, :] - C5_GINJ[:, 1:2]) ** 2) / (2 * C5_GINJ[:, 2:3] ** 2))
    y = y[0] + 2.0 * np.random.RandomState(0).randn(nt)
    bounds = {"gauss": ([2.5, t.min(), 0.01], [3.5, t.max(), 0.21]), "sine": ([0.5, 1.0, 0.0], [1.5, 20.0, 2 * np.pi])}
    priors = {n: {i: uniform_dist(lo[i], hi[i]) for i in range(3)} for n, (lo, hi) in bounds.items()}
    r = np.random.RandomState(11)
    coords, inds = {}, {}
    for n, (lo, hi) in bounds.items():
        coords[n] = r.uniform(np.asarray(lo), np.asarray(hi), size=(T, W, L, 3))
        inds[n] = r.rand(T, W, L) < 0.4
        inds[n][:, :, 0] = True
    np.random.seed(3)
    smp = EnsembleSampler(W, {"gauss": 3, "sine": 3}, PulseLikelihood(t, y, 2.0, {"gauss": "gauss", "sine": "sine"}), priors,
                          tempering_kwargs=dict(ntemps=T), nbranches=2, branch_names=["gauss", "sine"],
                          nleaves_max={"gauss": L, "sine": L}, nleaves_min={"gauss": 0, "sine": 0},
                          moves=GroupStretchMove(nfriends=nfriends, n_iter_update=100), rj_moves=True, rng="philox", seed=1,
                          device=dev)
    st0 = State(coords, inds=inds)
    smp.run_mcmc(st0, 0, burn=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp.run_mcmc(st0, 1, burn=iters - 1, thin_by=1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    return dict(value=T * W / dt, ms_per_step=dt * 1e3, iterations=iters,
                workload=f"C5: {T} temps x {W} walkers, branches gauss + sine (3-d leaves, nleaves_max={L}), {nt}-point data, "
                         f"GroupStretchMove(nfriends={nfriends}) + DistributionGenerateRJ birth/death, 2 swap passes per iteration",
                api="EnsembleSampler.run_mcmc (burn + 1 stored step), wall clock")


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from eryn_b200 import _lib

    wl = workload(args.workload, world)
    T, W, d = wl["T"], wl["W"], wl["d"]
    dev = torch.device("cuda", local)
    peak, peak_src = measured_peak_gbs()
    if world > 1:
        run_multi(args, wl, rank, world, local, dev, peak, peak_src)
        return

    sg = SingleGpu(wl, dev)
    ctx, tc = sg.ctx, sg.tc
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    clk = ClockSampler(local)
    t_wall0 = time.perf_counter()
    total_ms, launches = sg.timed_steps(args.steps, args.warmup, flush=flush, clk=clk)
    t_wall = time.perf_counter() - t_wall0
    value = T * W * args.steps / (total_ms * 1e-3)
    resident_ms = sg.resident_ms(args.steps, min_seconds=0.0 if args.profile else 0.35)
    clocks = clk.summary()

    # ---- roofline of the dominant kernel: the fused stretch step (two launches: red half, blue half) ------------------
    k_us = sg.stretch_us(100)
    alg_bytes = (24 * d + 41) * T * W  # 24*D+41 B per walker-update; one step updates all T*W walkers
    traffic = None  # DRAM bytes per stretch step from the committed ncu --set full capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[args.workload]
        traffic = tr["launches_per_step"] * (tr["dram_bytes_read_per_launch"] + tr["dram_bytes_write_per_launch"])
    except Exception:
        pass
    roofline = roofline_entry("stretch_step_kernel (both red/blue launches of one step)", alg_bytes, k_us, peak, peak_src, traffic,
                              note="per stretch step = 2 launches; the state (4.7 MB at C2) is L2-resident across launches, so this "
                                   "is an algorithmic-bytes rate against the HBM peak, limited by launch/dependency latency at this "
                                   "size; extra.roofline_c4 is the same step on the HBM-sized config-4 working set")
    # ---- the swap pass: 12 B per walker (logl read + slot bookkeeping) + (16 D + 34) B per walker whose row moves -------
    roofline_swap = None
    extra = {}
    if not args.profile:
        mf = sg.moved_fraction()
        s_us = sg.swap_us(100)
        sb = (12 + (16 * d + 34) * mf) * T * W
        roofline_swap = roofline_entry("pt_swap_kernel (ladder + row moves, one launch; the ladder adaptation is left to the next "
                                       "stretch kernel)" if sg.ctx.lazy_adapt else
                                       "pt_swap_kernel (ladder + row moves + adaptation, one launch)", sb, s_us, peak, peak_src,
                                       moved_fraction=round(mf, 4),
                                       note="algorithmic bytes = 12 + (16 D + 34) x moved_fraction per walker; latency-bound at "
                                            "this size (the ladder is a dependent chain over the rungs)")
        # whole iteration against its algorithmic bytes (~ 40 D + 90 B per walker-update)
        extra["roofline_iteration"] = roofline_entry("one iteration = stretch step + swap pass (graph replay, L2 flushed)",
                                                     (40 * d + 90) * T * W, total_ms / args.steps * 1e3, peak, peak_src)

    # ---- the other BASELINE configurations on this GPU ------------------------------------------------------------------
    roofline_c4 = None
    if not args.profile:
        del sg.graphs
        wl4 = workload("c4")
        sg4 = SingleGpu(wl4, dev, seed=7)
        k4_us = sg4.stretch_us(20)
        b4 = (24 * wl4["d"] + 41) * wl4["T"] * wl4["W"]
        roofline_c4 = roofline_entry("stretch step (lane-split kernel, both launches)", b4, k4_us, peak, peak_src,
                                     workload=wl4["label"], walker_updates_per_s=wl4["T"] * wl4["W"] / (k4_us * 1e-6))
        tot4, _ = sg4.timed_steps(30, 5, flush=flush)
        res4 = sg4.resident_ms(30)
        n4 = wl4["T"] * wl4["W"]
        extra["c4"] = dict(value=n4 * 30 / (tot4 * 1e-3), ms_per_step=tot4 / 30, ms_per_step_resident_no_flush=res4,
                           swap_us=round(sg4.swap_us(20), 2), workload=wl4["label"],
                           roofline_iteration=roofline_entry("one iteration = stretch step + swap pass", (40 * wl4["d"] + 90) * n4,
                                                             tot4 / 30 * 1e3, peak, peak_src))
        del sg4
        wl3 = workload("c3")
        sg3 = SingleGpu(wl3, dev, seed=9)
        tot3, _ = sg3.timed_steps(100, 10, flush=flush)
        extra["c3"] = dict(value=wl3["T"] * wl3["W"] * 100 / (tot3 * 1e-3), ms_per_step=tot3 / 100,
                           ms_per_step_resident_no_flush=sg3.resident_ms(100), workload=wl3["label"])
        del sg3
        try:
            extra["c5"] = bench_c5(dev)
        except Exception as e:  # reported, never silently dropped
            extra["c5"] = dict(error=f"{type(e).__name__}: {e}")
        try:
            extra["resident_kernel"] = bench_resident(dev)
        except Exception as e:
            extra["resident_kernel"] = dict(error=f"{type(e).__name__}: {e}")
        extra["api"] = bench_api(wl, dev, store=True)
        extra["api_thin100"] = bench_api(wl, dev, nsteps=10, thin_by=100, store=True)
        extra["api_not_stored"] = bench_api(wl, dev, store=False)
        extra["api"]["note"] = ("stored: every 25th iteration leaves as a 5.3 MB sample; the GPU side is a pack kernel + an "
                                "asynchronous copy, the host side is Backend.save_step's copy into freshly grown chain memory "
                                "(first-touch page faults, ~1 us per 4 KiB page on one host thread) — the cost the reference's "
                                "own save_step pays; api_thin100 stores 4x less often, api_not_stored not at all")

    # ---- e2e: C-ABI call with HOST buffers, one iteration per call ------------------------------------------
    lib = _lib.load()
    host = ctx.download(sg.ds)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_coords, h_logl, h_logp, h_betas = pin(host.branches_coords["model_0"]), pin(host.log_like), pin(host.log_prior), pin(host.betas)
    lo, hi, _ = sg.pri.arrays()
    lk = device_like(wl)
    par = np.ascontiguousarray(lk.params())
    schedule = sg.schedule(args.warmup + args.steps)
    h_sched = np.ascontiguousarray(schedule.astype(np.uint8))
    job = _lib.eb_host_job()
    job.ntemps, job.nwalkers, job.nleaves, job.ndim = T, W, 1, d
    job.coords_host, job.logl_host, job.logp_host, job.betas_host = [ctypes.c_void_p(t.data_ptr()) for t in (h_coords, h_logl, h_logp, h_betas)]
    job.prior_lo_host, job.prior_hi_host = ctypes.c_void_p(lo.ctypes.data), ctypes.c_void_p(hi.ctypes.data)
    job.like_kind, job.like_ncomp, job.like_nparams = int(lk.kind), int(lk.ncomp), int(par.size)
    job.like_params_host = ctypes.c_void_p(par.ctypes.data) if par.size else None
    job.stretch_a, job.gauss_scale, job.seed, job.iter0 = 2.0, 0.1, 20261017, 0
    job.adapt = _lib.eb_adapt(1, -1, 10000.0, 100.0)
    job.adapt_time0, job.permute, job.randomize_split = 0, 1, 1
    ne2e = max(20, min(args.steps, 200))
    for i in range(3):
        job.move_schedule_host = ctypes.c_void_p(h_sched.ctypes.data + i)
        _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
    t0 = time.perf_counter()
    for i in range(ne2e):
        job.move_schedule_host = ctypes.c_void_p(h_sched.ctypes.data + (i % len(h_sched)))
        _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
    e2e_s = (time.perf_counter() - t0) / ne2e
    state_bytes = h_coords.numel() * 8 + h_logl.numel() * 8 + h_logp.numel() * 8 + h_betas.numel() * 8
    e2e = dict(value=T * W / e2e_s, unit=UNIT,
               h2d_bytes_per_step=int(state_bytes + par.size * 8 + 3 * d * 8 + lib.eb_ctrl_size()),
               d2h_bytes_per_step=int(state_bytes + 16 + 4 * (T - 1)), ms_per_step=e2e_s * 1e3,
               api="eb_run_host(job, niter=1), every step: pinned host State in, host State out; wavefront schedule "
                   "(temperature groups uploaded hottest first, moved as they land, ladder resolved in rung ranges, "
                   "finished rungs downloaded while colder groups arrive; transfers by kernels on the mapped host "
                   "arrays, one captured CUDA graph per call)")
    # the same call with the plain schedule (upload all -> 3 kernels -> download all through the copy engines)
    os.environ["EB_HOST_PIPE"] = "0"
    for i in range(3):
        _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
    t0 = time.perf_counter()
    for i in range(ne2e):
        _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
    e2e_plain_s = (time.perf_counter() - t0) / ne2e
    os.environ.pop("EB_HOST_PIPE")
    extra["e2e_plain_schedule"] = dict(value=T * W / e2e_plain_s, ms_per_step=e2e_plain_s * 1e3,
                                       api="EB_HOST_PIPE=0: H2D -> 3 kernels -> D2H through the copy engines")

    # ---- CPU baseline (oracle port), bounded sample -------------------------------------------------------------
    cpu_v, cpu_s, cpu_n = run_cpu(wl, steps=10 ** 6, warmup=1, budget_s=0.5 if args.profile else 12.0)
    cpu = dict(value=cpu_v, unit=UNIT, cores=cpu_threads(), kind="port",
               sample=f"{cpu_n} iterations of the same workload ({cpu_s * 1e3:.1f} ms/iteration), NumPy oracle port, "
                      f"os.cpu_count()={os.cpu_count()}, BLAS pool {blas_pool()} (idle: the path is single-threaded NumPy, like the reference)")

    extra.update(ms_per_step_resident_no_flush=resident_ms, value_resident_no_flush=T * W / (resident_ms * 1e-3),
                 roofline_c4=roofline_c4, roofline_swap=roofline_swap, wall_s_timed_region=t_wall,
                 betas_cold_hot=[float(tc.betas[0]), float(tc.betas[-1])], swaps_accepted_last=tc.swaps_accepted.tolist()[:4],
                 notes=dict(rng="philox (counter-based, in-kernel)",
                            l2="flushed between timed steps (256 MiB fill outside the per-step CUDA-event pairs)",
                            step="one iteration = stretch step (red launch + blue launch chained by programmatic dependent "
                                 "launch) + 1 swap kernel (CUDA graph replay); the swap kernel leaves the ladder adaptation "
                                 "of its pass to the prologue of the next stretch kernel (lazy adaptation)"))
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup,
               ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f64", data="synthetic", config=config_of(wl),
               clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, extra=extra)
    print(json.dumps(out))


def run_multi(args, wl, rank, world, local, dev, peak, peak_src):
    """N > 1: the ladder sharded by temperature (eryn_b200/dist.py)"""
    import torch
    import torch.distributed as dist
    from eryn_b200.dist import run_sharded_bench, sharded_parity_check
    T, W, d = wl["T"], wl["W"], wl["d"]
    wl["device_like"] = device_like(wl)
    wl["x0"] = initial_coords(wl)
    res = run_sharded_bench(args, wl, rank, world, local, clock_sampler_cls=ClockSampler, comm=args.comm, sustain_s=0.35)
    # ---- BASELINE config 4 (32 temperatures, fixed) sharded over the same ranks: strong scaling -------------------------
    wl4 = workload("c4")
    c4 = None
    if wl4["T"] >= world and not args.profile:
        wl4["device_like"] = device_like(wl4)
        wl4["x0"] = initial_coords(wl4)
        a4 = argparse.Namespace(steps=min(args.steps, 50), warmup=max(3, min(args.warmup, 5)))
        r4 = run_sharded_bench(a4, wl4, rank, world, local, comm=args.comm, with_e2e=False, with_k1=False)
        one = None
        if rank == 0:  # the same configuration on ONE GPU (rank 0's), same timing method
            sg4 = SingleGpu(wl4, dev, seed=7)
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
            tot4, _ = sg4.timed_steps(a4.steps, a4.warmup, flush=flush)
            one = tot4 / a4.steps
            del sg4, flush
        dist.barrier()
        if rank == 0:
            c4 = dict(workload=wl4["label"], value=r4["value"], ms_per_step=r4["ms_per_step"],
                      ms_per_step_resident_no_flush=r4["resident_ms"], ms_per_step_1gpu=one,
                      speedup_vs_1gpu=one / r4["ms_per_step"], temps_per_gpu=wl4["T"] // world,
                      scaling="strong: the configuration is fixed, its 32 temperatures are spread over the GPUs")
    # ---- parity of the sharded path, in the same run ---------------------------------------------------------------------
    parity = sharded_parity_check(rank, world, local, comm=args.comm)
    if rank == 0:
        roof = None
        if res.get("k1_us"):
            b = (24 * d + 41) * res["local_temps"] * W
            roof = roofline_entry("stretch_step_kernel (rank 0's temperatures, both launches of one step)", b, res["k1_us"], peak,
                                  peak_src)
        comm_txt = {"fused": "every rank resolves the whole ladder: logl all-gather as NVLink peer stores of self-validating units "
                             "inside the swap kernel, rows pushed as mail (no NCCL on the data path)",
                    "split": "chain-split pass: every rank resolves 1/N of the chains; logl units to the resolver, accept bits to "
                             "every rank, rows pushed as mail — three one-way NVLink hops of self-validating units inside ONE "
                             "kernel (k_swap_split.cu), no NCCL on the data path",
                    "p2p": "publish kernel: NVLink peer stores + flag words, no NCCL on the data path",
                    "nccl": "NCCL all_gather of logl + NVLink peer row pulls"}[res["comm"]]
        out = dict(metric=METRIC, value=res["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=res["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="f64", data="synthetic", config=config_of(wl),
                   clocks=res["clocks"], e2e=dict(unit=UNIT, **res["e2e"]), gpu_launches=res["launches"],
                   roofline=roof, cpu_baseline=None, parity=parity,
                   extra=dict(ms_per_step_resident_no_flush=res["resident_ms"],
                              value_resident_no_flush=T * W / (res["resident_ms"] * 1e-3), c4_strong=c4,
                              betas_cold_hot=res["betas"], swaps_accepted_last=res["swaps"],
                              timing="per-step CUDA events on each rank, summed, MAX over ranks",
                              notes=dict(rng="philox (counter-based, in-kernel)",
                                         parallelism=f"temperature-sharded x{world} (temp_begin={res['temp_begin']}), weak "
                                                     f"scaling: 16 temperatures x 4096 walkers per GPU",
                                         comm=comm_txt,
                                         l2="flushed between timed steps (256 MiB fill outside the per-step CUDA-event pairs)",
                                         step="one iteration = 2 stretch launches + 1 sharded swap/adapt kernel, chained by "
                                              "programmatic dependent launch" + (" (CUDA graph replay)" if res["graph"] else ""))))
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit("sharded parity check FAILED: " + json.dumps(parity))


def run_reference(args):
    """Reference arm: the reference's CPU algorithm for the path (oracle port; the Python reference cannot travel
    to the GPU box), same workload/metric, all host threads NumPy will use.  Rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    wl = workload(args.workload, max(1, args.gpus))
    v, s, n = run_cpu(wl, steps=args.steps, warmup=args.warmup, budget_s=150.0)
    cpu = dict(value=v, unit=UNIT, cores=cpu_threads(), kind="port",
               sample=f"{n} iterations, {s * 1e3:.1f} ms/iteration, os.cpu_count()={os.cpu_count()}, BLAS pool {blas_pool()} "
                      f"(idle: the reference path is single-process, single-threaded NumPy; its only multi-core "
                      f"facility is pool.map over per-walker likelihood calls, slower than vectorising)")
    extra = dict(rng="numpy MT19937 (reference order)")
    if args.gpus <= 1:
        try:  # the secondary baseline BASELINE.md §3 names: vectorize=False + multiprocessing.Pool (ensemble.py:1474-1481)
            rate, ncores, m = run_cpu_pool(wl)
            per_iter = wl["T"] * wl["W"] / rate
            extra["pool_baseline"] = dict(
                likelihood_calls_per_s=rate, cores=ncores, sample=f"{m} per-walker likelihood calls through Pool({ncores}).map",
                value=wl["T"] * wl["W"] / (per_iter + s), unit=UNIT,
                note="walker-updates/s of an iteration whose likelihood leg goes through the pool (one call per walker) and "
                     "whose remaining legs take the vectorised port's time: slower than vectorize=True, as SURVEY.md §8d expects")
        except Exception as e:
            extra["pool_baseline"] = dict(error=f"{type(e).__name__}: {e}")
    print(json.dumps(dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=n,
                          warmup=args.warmup, ms_per_step=s * 1e3, higher_is_better=True, scaling="weak",
                          vs_baseline=None, dtype="f64", data="synthetic", config=config_of(wl), cpu_baseline=cpu,
                          e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), extra=extra)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--comm", default="auto", choices=["auto", "fused", "split", "p2p", "nccl"],
                    help="multi-GPU swap pass (N > 1); auto = fused at 2 GPUs, chain-split beyond")
    ap.add_argument("--profile", action="store_true", help="shorten the CPU-baseline leg and skip the extras (for runs under ncu)")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
