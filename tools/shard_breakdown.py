"""Where the time of a temperature-sharded iteration goes (run under torch.distributed.run, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py

For every comm mode the iteration is captured as a CUDA graph per buffer parity (as bench.py does) and replayed back to
back, so the ranks pace each other on the device; the per-iteration time is the CUDA-event time of the whole run / n.
With the profiling build (ERYN_B200_LIB=tools/_build/liberyn_b200_prof.so) the in-kernel phase marks of the LAST
iteration are printed as a timeline in ns (globaltimer of that GPU) relative to the start of the second stretch launch:
  K1b start/end (CTA 0) | swap: start, keys, [wait for K1b], publish issued, flags seen, logl gathered, cascade done,
  counts published, rows moved (CTA 0), adapt CTA done, last CTA rows moved."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    from eryn_b200 import dist as ed
    from eryn_b200.device import DeviceContext
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.moves import StretchMove
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    T, W, d = 16 * world, 4096, 8
    shape = os.environ.get("EB_BREAKDOWN_SHAPE")   # "c4": BASELINE config 4 (32 x 16384 x 20-d mixture), strong scaling
    A = np.random.RandomState(99).randn(d, d)
    P = np.linalg.inv(A @ A.T / d + np.eye(d))
    like = GaussianLikelihood(np.zeros(d), P)
    if shape == "c4":
        from eryn_b200.likelihood import GaussianMixtureLikelihood
        T, W, d = 32, 16384, 20
        r = np.random.RandomState(5)
        like = GaussianMixtureLikelihood(r.uniform(-5, 5, size=(4, d)), r.uniform(0.5, 1.5, size=4), np.full(4, 0.25))
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    prof = bool(os.environ.get("ERYN_B200_LIB"))
    modes = os.environ.get("EB_BREAKDOWN_MODES", "split,fused").split(",")
    for comm in modes:
        ctx = DeviceContext(pri, like, rng="philox", seed=1)
        run = ed.ShardedRun(ctx, T, W, comm=comm)
        tc = ed.ShardedTemperatureControl(run, d, W)
        mv = StretchMove(a=2.0)
        mv.temperature_control = tc
        mv.bind(ctx)
        mv.accepted = np.zeros((run.t_hi - run.t_lo, W))
        run.load(np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d)), tc._betas_host)
        stream = torch.cuda.Stream()
        graphs = []
        with torch.cuda.stream(stream):
            for _ in range(4):
                mv.propose(None, run.current)
            torch.cuda.synchronize()
            dist.barrier()
            for par in range(2):
                run.parity = par
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    mv.propose(None, run.current)
                graphs.append(g)
            run.parity = 0
            torch.cuda.synchronize()
            dist.barrier()
            n = 400
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(n):
                    graphs[i & 1].replay()
                e1.record(stream)
                torch.cuda.synchronize()
                dist.barrier()
            us = e0.elapsed_time(e1) * 1e3 / n
        run.check()
        print(f"[{comm}] rank {rank}: {us:.1f} us per iteration (graph replays back to back, T={T}, W={W})", flush=True)
        if prof:
            mn, mx = (ctypes.c_uint64 * 64)(), (ctypes.c_uint64 * 64)()
            sn, sx = (ctypes.c_uint64 * 64)(), (ctypes.c_uint64 * 64)()
            rd = "eb_debug_marks_split" if comm == "split" else "eb_debug_marks_swap"
            if getattr(run.lib, rd + "_global")(mn, mx, 0) == 0 and run.lib.eb_debug_marks_stretch_global(sn, sx, 0) == 0:
                base = sn[0]
                f = lambda v: int(v) - int(base)
                print(f"[{comm}] rank {rank} timeline ns: K1b start 0 end {f(sn[7])} lastCTA-end {f(sx[7])} | swap start {f(mn[16])} "
                      f"keys {f(mn[17])} publish-issued {f(mn[27])} flags-seen {f(mn[26])} gathered {f(mn[18])} cascade {f(mn[19])} "
                      f"counts {f(mn[20])} mail-pushed {f(mn[31])} rows {f(mn[21])} | adaptCTA start {f(mx[16])} gathered {f(mx[18])} all-arrived {f(mx[28])} "
                      f"folded {f(mx[29])} betas {f(mx[30])} done {f(mx[23])}",
                      flush=True)
            cta = (ctypes.c_uint64 * (8 * 1024))()
            if hasattr(run.lib, rd + "_cta") and getattr(run.lib, rd + "_cta")(cta) == 0:
                a = np.frombuffer(cta, dtype=np.uint64).reshape(8, 1024).astype(np.int64)
                nreal = (W + 7) // 8 if (T > 16 or comm == "split") else (W + 15) // 16
                if comm == "split":  # chain groups of 8 warps x (32 / nl) chains, nl = lanes per chain = pow2 >= owned rungs
                    nown = run.t_hi - run.t_lo
                    nl = 4 if nown <= 4 else 8 if nown <= 8 else 16 if nown <= 16 else 32
                    nreal = min(1024, (W + 8 * (32 // nl) - 1) // (8 * (32 // nl)))
                names = {0: "start", 1: "keys", 2: "gathered", 3: "cascade", 4: "counts", 7: "mail-pushed", 5: "rows"}
                txt = []
                for slot in (0, 1, 2, 3, 4, 7, 5):
                    v = a[slot, :nreal] - int(base)
                    txt.append(f"{names[slot]} min {v.min()} med {int(np.median(v))} max {v.max()} (argmax CTA {int(v.argmax())})")
                print(f"[{comm}] rank {rank} over the {nreal} chain CTAs, ns: " + " | ".join(txt), flush=True)
        dist.barrier()
        run.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
