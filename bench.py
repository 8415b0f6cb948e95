#!/usr/bin/env python
"""bench.py — walker-updates/sec of the sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl reference]

One "step" = one sampler iteration of the workload: one in-model move (every walker gets one proposal +
Metropolis test) + one parallel-tempering swap pass + ladder adaptation.  N=1 runs BASELINE config 2
(16 temps x 4096 walkers x 8-d correlated Gaussian, StretchMove + PT).  Prints ONE JSON line on rank 0.

  value     device-resident throughput (state already in HBM), production (philox) mode, each step a
            replay of the captured 2-kernel iteration graph, timed by its own CUDA-event pair with an
            L2 flush between steps (outside the pairs);
  e2e       the same step through the C-ABI entry with HOST buffers (eb_run_host): pinned host state ->
            H2D -> kernels -> D2H -> host state, every step;
  roofline  the fused stretch half-step kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak;
  cpu_baseline  the NumPy oracle port of the reference path on this box's host cores, same workload.
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
# clocks sampler (nvidia-smi fields via NVML) — runs during the timed region
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
            time.sleep(0.02)

    def summary(self):
        self.stop_flag = True
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
    from eryn_b200.device import DeviceContext
    from eryn_b200.moves import StretchMove, GaussianMove, TemperatureControl
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    from eryn_b200.state import State

    wl = workload(args.workload, world)
    T, W, d = wl["T"], wl["W"], wl["d"]
    if world > 1:
        from eryn_b200.dist import run_sharded_bench
        wl["device_like"] = device_like(wl)
        wl["x0"] = initial_coords(wl)
        res = run_sharded_bench(args, wl, rank, world, local, clock_sampler_cls=ClockSampler, comm=args.comm)
        if rank == 0:
            peak, peak_src = measured_peak_gbs()
            roof = None
            if res.get("k1_us"):
                b = (24 * d + 41) * res["local_temps"] * W
                a = b / (res["k1_us"] * 1e-6) / 1e9
                roof = dict(bound="hbm", kernel="stretch_step_kernel (rank 0's temperatures, both launches of one step)",
                            achieved=round(a, 1), peak=peak, unit="GB/s", frac=round(a / peak, 4), traffic=None,
                            peak_source=peak_src, algorithmic_bytes_per_launch=b, avg_launch_us=round(res["k1_us"], 3))
            out = dict(metric=METRIC, value=res["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                       ms_per_step=res["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                       dtype="f64", data="synthetic",
                       config=dict(workload=wl["label"], ntemps=T, nwalkers=W, ndim=d, rng="philox",
                                   parallelism=f"temperature-sharded x{world} (temp_begin={res['temp_begin']}), "
                                               f"weak scaling: 16 temperatures x 4096 walkers per GPU",
                                   comm={"fused": "logl all-gather as NVLink peer stores + iteration flags inside the swap kernel "
                                                  "(no separate publish launch, no NCCL on the data path)",
                                         "split": "EXPERIMENTAL chain-split pass: every rank resolves 1/N of the chains, two one-way "
                                                  "NVLink hops of self-validating units (k_swap_split.cu)",
                                         "p2p": "publish kernel: NVLink peer stores + flag words, no NCCL on the data path",
                                         "nccl": "NCCL all_gather of logl + NVLink peer row pulls"}[res["comm"]],
                                   l2="flushed between timed steps (256 MiB fill outside the per-step CUDA-event pairs)",
                                   step=("one iteration = 2 stretch launches + 1 sharded publish/swap/adapt kernel, chained by "
                                         "programmatic dependent launch" if res["comm"] in ("fused", "split") else
                                         "one iteration = move kernel + publish kernel + sharded swap/adapt kernel")
                                        + (" (CUDA graph replay)" if res["graph"] else "")),
                       clocks=res["clocks"], e2e=dict(unit=UNIT, **res["e2e"]), gpu_launches=res["launches"],
                       roofline=roof, cpu_baseline=None,
                       extra=dict(ms_per_step_resident_no_flush=res["resident_ms"],
                                  value_resident_no_flush=T * W / (res["resident_ms"] * 1e-3),
                                  betas_cold_hot=res["betas"], swaps_accepted_last=res["swaps"],
                                  timing="per-step CUDA events on each rank, summed, MAX over ranks"))
            print(json.dumps(out))
        dist.barrier()
        dist.destroy_process_group()
        return

    dev = torch.device("cuda", local)
    pri = ProbDistContainer({i: uniform_dist(wl["lo"], wl["hi"]) for i in range(d)})
    ctx = DeviceContext(pri, device_like(wl), device=dev, rng="philox", seed=20261017)
    tc = TemperatureControl(d, W, ntemps=T)
    tc.bind(ctx)
    moves = []
    for m in wl["moves"]:
        mv = StretchMove(a=m["a"]) if m["kind"] == "stretch" else GaussianMove({"model_0": m["proposal"]["scale"] ** 2})
        mv.temperature_control = tc
        mv.bind(ctx)
        mv.accepted = np.zeros((T, W))
        moves.append(mv)
    x0 = initial_coords(wl)
    ds = ctx.upload(State(x0), betas=tc.betas_dev)
    ctx.eval_state(ds)
    torch.cuda.synchronize()

    # host-side move schedule (the reference's per-iteration random.choice, ensemble.py:971)
    sched_rng = np.random.RandomState(7)
    nmoves = len(moves)

    # ---- capture one iteration per move kind into a CUDA graph -------------------------------------------
    stream = torch.cuda.Stream(device=dev)
    graphs = []
    with torch.cuda.stream(stream):
        for mv in moves:  # warm-up launches outside capture (module load, attribute setup)
            mv.propose(None, ds)
        torch.cuda.synchronize()
        for mv in moves:
            g = torch.cuda.CUDAGraph()
            l0 = ctx.launches
            with torch.cuda.graph(g, stream=stream):
                mv.propose(None, ds)
            graphs.append((g, ctx.launches - l0))
    torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    schedule = sched_rng.choice(nmoves, p=np.asarray(wl["weights"]) / np.sum(wl["weights"]), size=args.warmup + args.steps)

    clk = ClockSampler(local)
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            graphs[schedule[i]][0].replay()
        torch.cuda.synchronize()
        clk.start()
        launches = 0
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations, outside the event pair
            g, nl = graphs[schedule[args.warmup + i]]
            evs[i][0].record(stream)
            g.replay()
            evs[i][1].record(stream)
            launches += nl
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        # the same K steps back to back without flush (state stays L2-resident, as in a real run)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            graphs[schedule[args.warmup + i]][0].replay()
        e1.record(stream)
        torch.cuda.synchronize()
    clocks = clk.summary()
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    value = T * W * args.steps / (total_ms * 1e-3)
    resident_ms = e0.elapsed_time(e1) / args.steps

    # ---- roofline of the dominant kernel: the fused stretch step (two launches: red half, blue half) ------------------
    def stretch_us(ctx_, ds_, cnt_, nrep):
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            ctx_.stretch_step(ds_, 2.0, accepted_count=cnt_)
            torch.cuda.synchronize()
            with torch.cuda.graph(gk, stream=stream):
                for r in range(nrep):
                    ctx_.stretch_step(ds_, 2.0, accepted_count=cnt_)
            gk.replay()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            gk.replay()
            k1.record(stream)
            torch.cuda.synchronize()
        return k0.elapsed_time(k1) * 1e3 / nrep

    st_move = [m for m in moves if isinstance(m, StretchMove)][0]
    cnt = st_move._count_buffer(ctx, T, W)
    k_us = stretch_us(ctx, ds, cnt, 100)
    D = d
    alg_bytes = (24 * D + 41) * T * W  # 24*D+41 B per walker-update; one step updates all T*W walkers
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (k_us * 1e-6) / 1e9
    traffic = None  # DRAM bytes per stretch step from the committed ncu --set full capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[args.workload]
        traffic = tr["launches_per_step"] * (tr["dram_bytes_read_per_launch"] + tr["dram_bytes_write_per_launch"])
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel="stretch_step_kernel (both red/blue launches of one step)", achieved=round(achieved, 1),
                    peak=peak, unit="GB/s", frac=round(achieved / peak, 4), traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg_bytes, avg_launch_us=round(k_us, 3),
                    note="per stretch step = 2 launches; state (4.7 MB at C2) is L2-resident across launches, so this is an "
                         "algorithmic-bytes rate against the HBM peak, limited by launch/dependency latency at this size; "
                         "extra.roofline_c4 is the same kernel on the HBM-sized config-4 working set")
    # the same kernel where the working set exceeds what one launch can hide behind latency: BASELINE config 4 on one GPU
    roofline_c4 = None
    if not args.profile:
        wl4 = workload("c4")
        pri4 = ProbDistContainer({i: uniform_dist(wl4["lo"], wl4["hi"]) for i in range(wl4["d"])})
        ctx4 = DeviceContext(pri4, device_like(wl4), device=dev, rng="philox", seed=7)
        tc4 = TemperatureControl(wl4["d"], wl4["W"], ntemps=wl4["T"])
        tc4.bind(ctx4)
        ds4 = ctx4.upload(State(initial_coords(wl4)), betas=tc4.betas_dev)
        ctx4.eval_state(ds4)
        cnt4 = torch.zeros((wl4["T"], wl4["W"]), dtype=torch.int32, device=dev)
        k4_us = stretch_us(ctx4, ds4, cnt4, 20)
        b4 = (24 * wl4["d"] + 41) * wl4["T"] * wl4["W"]
        a4 = b4 / (k4_us * 1e-6) / 1e9
        roofline_c4 = dict(workload=wl4["label"], kernel="stretch_step_kernel", achieved=round(a4, 1), peak=peak, unit="GB/s",
                           frac=round(a4 / peak, 4), algorithmic_bytes_per_launch=b4, avg_launch_us=round(k4_us, 2),
                           walker_updates_per_s=wl4["T"] * wl4["W"] / (k4_us * 1e-6))
        del ds4, ctx4, cnt4

    # ---- e2e: C-ABI call with HOST buffers, one iteration per call ------------------------------------------
    lib = _lib.load()
    host = ctx.download(ds)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_coords, h_logl, h_logp, h_betas = pin(host.branches_coords["model_0"]), pin(host.log_like), pin(host.log_prior), pin(host.betas)
    lo, hi, _ = pri.arrays()
    par = np.ascontiguousarray(device_like(wl).params())
    h_sched = np.ascontiguousarray(schedule.astype(np.uint8))
    job = _lib.eb_host_job()
    job.ntemps, job.nwalkers, job.nleaves, job.ndim = T, W, 1, d
    job.coords_host, job.logl_host, job.logp_host, job.betas_host = [ctypes.c_void_p(t.data_ptr()) for t in (h_coords, h_logl, h_logp, h_betas)]
    job.prior_lo_host, job.prior_hi_host = ctypes.c_void_p(lo.ctypes.data), ctypes.c_void_p(hi.ctypes.data)
    lk = device_like(wl)
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
    e2e = dict(value=T * W / e2e_s, unit=UNIT, h2d_bytes_per_step=int(state_bytes + par.size * 8 + 3 * d * 8 + 4120),
               d2h_bytes_per_step=int(state_bytes + 4120), ms_per_step=e2e_s * 1e3,
               api="eb_run_host(job, niter=1): pinned host State -> H2D -> 2 kernels -> D2H, every step")

    # ---- CPU baseline (oracle port), bounded sample -------------------------------------------------------------
    cpu_v, cpu_s, cpu_n = run_cpu(wl, steps=10 ** 6, warmup=1, budget_s=0.5 if args.profile else 12.0)
    cpu = dict(value=cpu_v, unit=UNIT, cores=cpu_threads(), kind="port",
               sample=f"{cpu_n} iterations of the same workload ({cpu_s * 1e3:.1f} ms/iteration), NumPy oracle port, "
                      f"os.cpu_count()={os.cpu_count()}, BLAS pool {blas_pool()} (idle: the path is single-threaded NumPy, like the reference)")

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup,
               ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f64", data="synthetic",
               config=dict(workload=wl["label"], ntemps=T, nwalkers=W, ndim=d, rng="philox",
                           l2="flushed between timed steps (256 MiB fill outside the per-step CUDA-event pairs)",
                           step="one iteration = stretch step (red launch + blue launch chained by programmatic dependent launch) + 1 swap/adapt kernel (CUDA graph replay)"),
               clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu,
               extra=dict(ms_per_step_resident_no_flush=resident_ms,
                          value_resident_no_flush=T * W / (resident_ms * 1e-3), roofline_c4=roofline_c4,
                          wall_s_timed_region=t_wall, betas_cold_hot=[float(tc.betas[0]), float(tc.betas[-1])],
                          swaps_accepted_last=tc.swaps_accepted.tolist()[:4]))
    print(json.dumps(out))


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
    print(json.dumps(dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=n,
                          warmup=args.warmup, ms_per_step=s * 1e3, higher_is_better=True, scaling="weak",
                          vs_baseline=None, dtype="f64", data="synthetic",
                          config=dict(workload=wl["label"], ntemps=wl["T"], nwalkers=wl["W"], ndim=wl["d"],
                                      rng="numpy MT19937 (reference order)"),
                          cpu_baseline=cpu,
                          e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--comm", default="fused", choices=["fused", "split", "p2p", "nccl"], help="multi-GPU logl exchange (N > 1)")
    ap.add_argument("--profile", action="store_true", help="shorten the CPU-baseline leg (for runs under ncu)")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
