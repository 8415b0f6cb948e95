"""Time the resident kernel (eb_resident_run) on config 2 for several iterations per launch (CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eryn_b200.device import DeviceContext  # noqa: E402
from eryn_b200.likelihood import GaussianLikelihood  # noqa: E402
from eryn_b200.prior import ProbDistContainer, uniform_dist  # noqa: E402
from eryn_b200.state import State  # noqa: E402


def main():
    T, W, d = [int(x) for x in os.environ.get("EB_PROBE_SHAPE", "16,4096,8").split(",")]
    r = np.random.RandomState(0)
    A = r.randn(d, d)
    ctx = DeviceContext(ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)}),
                        GaussianLikelihood(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d))), rng="philox", seed=3)
    ds = ctx.upload(State({"model_0": r.uniform(-3, 3, size=(T, W, 1, d))}),
                    betas=torch.from_numpy(np.geomspace(1.0, 1e-3, T)).to(ctx.device))
    ctx.eval_state(ds)
    adapt = dict(adaptive=True, stop_adaptation=-1, adaptation_lag=10000.0, adaptation_time=100.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps, do_flush):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            if do_flush:
                flush.fill_(1)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps * 1e3

    def per_launch(n):
        def f():
            for _ in range(n):
                ctx.stretch_step(ds, 2.0)
                ctx.pt_swap(ds, adapt=adapt)
        return f
    g = torch.cuda.CUDAGraph()
    per_launch(1)()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        per_launch(1)()
    print(f"per-launch kernels, graph of 1 iteration : {timed(g.replay, 50, True):8.1f} us cold   {timed(g.replay, 50, False):8.1f} us warm")
    for n in (1, 2, 5, 25, 100):
        f = lambda: ctx.resident_run(ds, 2.0, n, adapt=adapt)
        tc, tw = timed(f, 30, True), timed(f, 30, False)
        print(f"resident kernel, {n:3d} iterations / launch : {tc:8.1f} us cold   {tw:8.1f} us warm   = {tw / n:6.2f} us/iteration warm, "
              f"{T * W * n / (tw * 1e-6):.3e} wu/s")
    ctx.check_error()
    if os.environ.get("ERYN_B200_LIB"):   # marks build (tools/build_variant.sh resmarks -DEB_RES_MARKS)
        names = ["start", "half0 done", "csync0", "half1 done", "csync1", "published", "prologue", "gbar1", "resolved",
                 "gbar2", "fetched", "adapted", "csync2"]
        for n in (1, 25):
            ctx.resident_run(ds, 2.0, n, adapt=adapt)
            torch.cuda.synchronize()
            scr = ctx.scratch("resident", (int(ctx.lib.eb_resident_scratch_bytes(__import__("ctypes").byref(ds.c_struct()))),), torch.uint8)
            m = scr[1152:1152 + 256].cpu().numpy().view(np.uint64).astype(np.int64).reshape(2, 16)
            for row, who in zip(m, ("first CTA", "last CTA")):
                print(f"niter={n} {who}: " + "  ".join(f"{nm} {(row[k] - m[0][0]) * 1e-3:.2f}" for k, nm in enumerate(names)))


if __name__ == "__main__":
    main()
