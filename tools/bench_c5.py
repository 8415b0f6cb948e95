"""Config 5 (reversible jump + group stretch, 8 temps x 2048 walkers, 2 branches x 10 leaves) timing: device vs the oracle.

    python tools/bench_c5.py [--nt 500] [--iters 50]

One iteration = group-stretch move + swap pass (+adaptation) + rj move + swap pass.  Prints walker-updates/s
(ntemps x nwalkers per iteration, as BASELINE.json's metric) for the device (philox mode, device-resident) and for the
NumPy oracle port of the reference path (a few iterations)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=500)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--cpu-iters", type=int, default=1)
    a = ap.parse_args()
    import torch
    from oracle import eryn_oracle as orc
    from oracle import rj_oracle as rjo
    from tests import cases_rj
    from tests.test_gpu_parity_rj import make_sampler, random_start
    from eryn_b200.state import State
    T, W, L, nfriends = 8, 2048, 10, 16
    t = np.linspace(-1, 1, a.nt)
    y = cases_rj.GINJ[:, 0:1].T @ np.exp(-((t[None, :] - cases_rj.GINJ[:, 1:2]) ** 2) / (2 * cases_rj.GINJ[:, 2:3] ** 2))
    y = y[0] + 2.0 * np.random.RandomState(0).randn(a.nt)
    coords, inds = random_start(T, W, L, L, t, 11)
    smp, move = make_sampler(t, y, 2.0, T, W, L, L, nfriends, 100, "philox", seed=1)
    st0 = State({"gauss": coords[0], "sine": coords[1]}, inds={"gauss": inds[0], "sine": inds[1]})
    smp.run_mcmc(st0, 0, burn=5)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp.run_mcmc(st0, 1, burn=a.iters - 1, thin_by=1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.iters
    print(f"device : {dt * 1e3:8.3f} ms/iteration  {T * W / dt:.3e} walker-updates/s  (nt={a.nt}, {torch.cuda.get_device_name(0)})")
    if a.cpu_iters <= 0:
        return
    like = rjo.PulseLike(t, y, 2.0, [0, 1])
    osmp = rjo.OracleSamplerMB(cases_rj.priors_for(t), like, [0, 0], [L, L], rjo.PhiloxStreamsMB(1),
                               betas=orc.make_ladder_default(6 * L, T), nfriends=nfriends, n_iter_update=100)
    ost = osmp.initialise(rjo.MBState(coords, inds))
    t0 = time.perf_counter()
    for _ in range(a.cpu_iters):
        osmp.iterate(ost)
    dc = (time.perf_counter() - t0) / a.cpu_iters
    print(f"oracle : {dc * 1e3:8.1f} ms/iteration  {T * W / dc:.3e} walker-updates/s  (NumPy port of the reference path, "
          f"{os.cpu_count()} host cores visible)   ratio {dc / dt:.0f}x")


if __name__ == "__main__":
    main()
