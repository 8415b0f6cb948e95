"""Worker of tests/test_mgpu.py: one rank of a temperature-sharded run (launched by torch.distributed.run)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--comm", default="fused")
    ap.add_argument("--ntemps", type=int, default=4)
    ap.add_argument("--nwalkers", type=int, default=256)
    ap.add_argument("--ndim", type=int, default=8)
    ap.add_argument("--nit", type=int, default=6)
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--mix", type=int, default=0, help="1: stretch/gaussian schedule")
    ap.add_argument("--like", default="gauss", help="gauss: correlated Gaussian; gmix: BASELINE config 4's mixture of 4 Gaussians")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from eryn_b200 import dist as ed
    from eryn_b200.device import DeviceContext
    from eryn_b200.likelihood import GaussianLikelihood, GaussianMixtureLikelihood
    from eryn_b200.moves import GaussianMove, StretchMove
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    T, W, d = a.ntemps, a.nwalkers, a.ndim
    A = np.random.RandomState(99).randn(d, d)
    P = np.linalg.inv(A @ A.T / d + np.eye(d))
    pri = ProbDistContainer({i: uniform_dist(-10.0, 10.0) for i in range(d)})
    like = GaussianLikelihood(np.zeros(d), P)
    if a.like == "gmix":
        r = np.random.RandomState(5)
        like = GaussianMixtureLikelihood(r.uniform(-5, 5, size=(4, d)), r.uniform(0.5, 1.5, size=4), np.full(4, 0.25))
    ctx = DeviceContext(pri, like, rng="philox", seed=a.seed)
    run = ed.ShardedRun(ctx, T, W, comm=a.comm)
    tc = ed.ShardedTemperatureControl(run, d, W)
    moves = [StretchMove(a=2.0), GaussianMove({"model_0": 0.01})]
    for mv in moves:
        mv.temperature_control = tc
        mv.bind(ctx)
        mv.accepted = np.zeros((run.t_hi - run.t_lo, W))
    x0 = np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d))
    run.load(x0, tc._betas_host)
    sched = np.random.RandomState(7)
    swaps = None
    for it in range(a.nit):
        mi = int(sched.choice(2, p=[0.5, 0.5])) if a.mix else 0
        state, acc = moves[mi].propose(None, run.current)
        assert state is run.current
    full = run.gather()
    swaps = tc.swaps_accepted
    acc_all = ed.gather_rows(sum(m.accepted for m in moves), run.temp_begin)
    if dist.get_rank() == 0:
        np.savez(os.path.join(a.out, f"mgpu_{a.comm}.npz"), swaps=swaps, accepted=acc_all, time=tc.time, **full)
    run.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
