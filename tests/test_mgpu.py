"""2-GPU parity of the temperature-sharded run (run with `gpurun --gpus 2 -- python -m pytest tests/test_mgpu.py -m gpu`).

The sharded run (one process per GPU, NVLink peer stores + flags) must reproduce the unsharded
oracle chain: swap counts / accept counts bit-equal, floats to 1e-10 relative."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import eryn_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle(T, W, d, nit, seed, mix, like_kind="gauss"):
    A = np.random.RandomState(99).randn(d, d)
    like = orc.GaussianLike(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
    if like_kind == "gmix":
        r = np.random.RandomState(5)
        like = orc.GaussianMixtureLike(r.uniform(-5, 5, size=(4, d)), r.uniform(0.5, 1.5, size=4), np.full(4, 0.25))
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    moves = [dict(kind="stretch", a=2.0), dict(kind="gaussian", proposal=dict(kind="scalar", scale=0.1))]
    weights = [0.5, 0.5] if mix else [1.0, 0.0]
    sched = np.random.RandomState(7)

    class Sched:  # the worker draws choice(2, p=[.5,.5]) only when mixing
        def choice(self, n, p):
            return sched.choice(n, p=[0.5, 0.5]) if mix else 0

    smp = orc.OracleSampler(prior, like, moves, weights, orc.PhiloxStreams(seed, schedule_random=Sched()),
                            betas=orc.make_ladder_default(d, T))
    st = smp.initialise(orc.OState(np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d))))
    acc = np.zeros((T, W))
    for _ in range(nit):
        acc += smp.iterate(st)
    return smp, st, acc


@pytest.mark.parametrize("comm,T,W,mix", [("fused", 4, 256, 0), ("p2p", 4, 256, 0), ("nccl", 4, 256, 0), ("fused", 5, 99, 1),
                                                ("fused", 16, 4096, 0), ("p2p", 16, 4096, 0), ("fused", 72, 64, 0),
                                                ("fused", 128, 48, 0), ("fused", 32, 16384, 0),
                                                ("split", 4, 256, 0), ("split", 5, 99, 1), ("split", 16, 4096, 0),
                                                ("split", 128, 48, 0), ("split", 32, 16384, 0), ("split", 7, 1000, 1),
                                                ("split", 8, 16384, 0), ("split", 24, 333, 1), ("auto", 6, 128, 0)])
def test_sharded_run_matches_unsharded_oracle(tmp_path, comm, T, W, mix):
    _run_and_compare(tmp_path, comm, T, W, mix, nproc=2)


@pytest.mark.parametrize("T,W", [(4, 256), (16, 4096), (128, 48)])
def test_split_pass_single_rank_matches_oracle(tmp_path, T, W):
    """the chain-split pass with world = 1 (one GPU is enough): positions, units, cascade, counts exchange with itself"""
    _run_and_compare(tmp_path, "split", T, W, 0, nproc=1)


@pytest.mark.parametrize("comm", ["fused", "split"])
def test_config4_full_size_sharded(tmp_path, comm):
    """BASELINE config 4 at full size (32 temperatures x 16384 walkers x 20-d mixture of 4 Gaussians), the ladder sharded
    over 2 GPUs, against the unsharded oracle"""
    _run_and_compare(tmp_path, comm, 32, 16384, 0, nproc=2, d=20, nit=2, like_kind="gmix")


def _run_and_compare(tmp_path, comm, T, W, mix, nproc, d=8, nit=6, like_kind="gauss"):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    seed = 4242
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--out",
           str(tmp_path), "--comm", comm, "--ntemps", str(T), "--nwalkers", str(W), "--ndim", str(d), "--nit", str(nit), "--seed",
           str(seed), "--mix", str(mix), "--like", like_kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = np.load(tmp_path / f"mgpu_{comm}.npz")
    smp, st, acc = _oracle(T, W, d, nit, seed, mix, like_kind)
    np.testing.assert_allclose(got["coords"], st.coords, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(got["logl"], st.logl, rtol=1e-10)
    np.testing.assert_allclose(got["logp"], st.logp, rtol=1e-10)
    np.testing.assert_allclose(got["betas"], smp.betas, rtol=1e-10)
    assert np.array_equal(got["swaps"], smp.swaps_accepted)
    assert np.array_equal(got["accepted"], acc)
    assert int(got["time"]) == nit
