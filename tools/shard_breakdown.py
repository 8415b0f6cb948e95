"""Per-stage device time of the temperature-sharded iteration (run under torch.distributed.run, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_breakdown.py

Stages are bracketed by CUDA events on every rank (no graph): move kernels, publish kernel, sharded swap kernel.
The swap stage contains the in-kernel wait for the peers' flags, so rank skew shows up there."""
import ctypes as C
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
    from eryn_b200 import _lib
    from eryn_b200 import dist as ed
    from eryn_b200.device import DeviceContext
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.moves import StretchMove
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    T, W, d = 16 * world, 4096, 8
    A = np.random.RandomState(99).randn(d, d)
    P = np.linalg.inv(A @ A.T / d + np.eye(d))
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    for comm in ("p2p", "nccl"):
        ctx = DeviceContext(pri, GaussianLikelihood(np.zeros(d), P), rng="philox", seed=1)
        run = ed.ShardedRun(ctx, T, W, comm=comm)
        tc = ed.ShardedTemperatureControl(run, d, W)
        mv = StretchMove(a=2.0)
        mv.temperature_control = tc
        mv.bind(ctx)
        cnt = mv._count_buffer(ctx, run.t_hi - run.t_lo, W)
        run.load(np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d)), tc._betas_host)
        n = 200
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
        ad = dict(adaptive=True, stop_adaptation=-1, adaptation_lag=10000, adaptation_time=100)
        for it in range(20):
            ctx.stretch_step(run.current, 2.0, accepted_count=cnt)
            run.swap(adapt=ad)
        torch.cuda.synchronize()
        dist.barrier()
        for it in range(n):
            ev[it][0].record()
            ctx.stretch_step(run.current, 2.0, accepted_count=cnt)
            ev[it][1].record()
            p = run.parity
            if comm == "p2p":
                _lib.check(run.lib.eb_publish_logl(C.byref(run._pub[p]), C.c_void_p(ctx.ctrl.data_ptr()), ctx.stream()), "pub")
            else:
                dist.all_gather_into_tensor(run.logl_all[p], run.states[p].logl)
            ev[it][2].record()
            # swap without the publish: call the second half of run.swap by hand
            r = _lib.eb_swap_rng()
            r.mode, r.permute, r.seed, r.iter_dev = _lib.EB_RNG_PHILOX, 1, ctx.seed, ctx.iter_ptr
            a = _lib.eb_adapt(1, -1, 10000.0, 100.0)
            dst = run.states[1 - p].c_struct()
            _lib.check(run.lib.eb_pt_swap_sharded(C.byref(run._shard[p]), C.byref(dst), C.byref(r), C.byref(a),
                                                  C.c_void_p(ctx.ctrl.data_ptr()), ctx.stream()), "swap")
            run.parity = 1 - p
            ev[it][3].record()
        torch.cuda.synchronize()
        run.check()
        seg = np.array([[e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(3)] for e in ev])
        tot = np.array([ev[i][0].elapsed_time(ev[i + 1][0]) * 1e3 for i in range(n - 1)])
        print(f"[{comm}] rank {rank}: move {np.median(seg[:, 0]):.1f} us, publish/all-gather {np.median(seg[:, 1]):.1f} us, "
              f"sharded swap (incl. wait for peers) {np.median(seg[:, 2]):.1f} us; iteration-to-iteration {np.median(tot):.1f} us "
              f"(host-paced, no graph)", flush=True)
        if os.environ.get("ERYN_B200_LIB"):
            import ctypes
            mn = (ctypes.c_uint64 * 64)()
            mx = (ctypes.c_uint64 * 64)()
            if run.lib.eb_debug_marks_swap_global(mn, mx, 0) == 0:
                base = mn[16]
                print(f"[{comm}] rank {rank} last swap kernel, CTA 0, ns since its start: " +
                      " ".join(f"m{i}={mn[i] - base}" for i in (17, 26, 18, 19, 20, 24, 25, 21)) +
                      f" adapt-CTA-end={mx[23] - base}", flush=True)
        dist.barrier()
        run.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
