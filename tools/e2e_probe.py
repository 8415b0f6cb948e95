"""Time eb_run_host(job, 1) on config 2 (16 x 4096 x 8-d Gaussian) for the plain and the wavefront schedule
(EB_HOST_PIPE / EB_HOST_GROUPS / EB_HOST_GRAPH).  Wall clock around the calls, pinned host arrays."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eryn_b200 import _lib  # noqa: E402


def main():
    lib = _lib.require_device()
    T, W, d = [int(x) for x in os.environ.get("EB_PROBE_SHAPE", "16,4096,8").split(",")]
    r = np.random.RandomState(0)
    A = r.randn(d, d)
    P = np.linalg.inv(A @ A.T / d + np.eye(d))
    par = np.concatenate([np.zeros(d), P.ravel()])
    lo, hi = np.full(d, -10.0), np.full(d, 10.0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    x0 = r.uniform(-3, 3, size=(T, W, 1, d))
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    n = int(os.environ.get("EB_PROBE_N", "200"))
    only = os.environ.get("EB_PROBE_ONLY")
    for label, env in [("plain", dict(EB_HOST_PIPE="0")),
                       ("wave G=1 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="1")),
                       ("wave G=2 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="2")),
                       ("wave G=3 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="3")),
                       ("wave G=4 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="4")),
                       ("wave G=5 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="5")),
                       ("wave G=6 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="6")),
                       ("wave G=8 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="8")),
                       ("wave G=16 graph", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="16")),
                       ("wave G=4 direct", dict(EB_HOST_PIPE="2", EB_HOST_GROUPS="4", EB_HOST_GRAPH="0")),
                       ("wave 3,5,5,3", dict(EB_HOST_PIPE="2", EB_HOST_SPLIT="3,5,5,3")),
                       ("wave 2,4,5,3,2", dict(EB_HOST_PIPE="2", EB_HOST_SPLIT="2,4,5,3,2")),
                       ("wave 2,6,6,2", dict(EB_HOST_PIPE="2", EB_HOST_SPLIT="2,6,6,2")),
                       ("wave 1,3,4,4,3,1", dict(EB_HOST_PIPE="2", EB_HOST_SPLIT="1,3,4,4,3,1")),
                       ("wave 4,8,4", dict(EB_HOST_PIPE="2", EB_HOST_SPLIT="4,8,4")),
                       ("wave default", dict(EB_HOST_PIPE="1"))]:
        if only and label not in only.split(";"):
            continue
        for k in ("EB_HOST_PIPE", "EB_HOST_GROUPS", "EB_HOST_GRAPH", "EB_HOST_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        coords, betas = pin(x0.copy()), pin(np.geomspace(1.0, 1e-3, T))
        xx = x0[:, :, 0]
        logl, logp = pin(-0.5 * np.einsum("twi,ij,twj->tw", xx, P, xx)), pin(np.full((T, W), d * np.log(1 / 20.0)))
        job = _lib.eb_host_job()
        job.ntemps, job.nwalkers, job.nleaves, job.ndim = T, W, 1, d
        job.coords_host, job.logl_host, job.logp_host, job.betas_host = ptr(coords), ptr(logl), ptr(logp), ptr(betas)
        job.prior_lo_host, job.prior_hi_host = ctypes.c_void_p(lo.ctypes.data), ctypes.c_void_p(hi.ctypes.data)
        job.like_kind, job.like_ncomp, job.like_nparams = 0, 0, par.size
        job.like_params_host = ctypes.c_void_p(par.ctypes.data)
        job.stretch_a, job.gauss_scale, job.seed, job.iter0 = 2.0, 0.1, 20261017, 0
        job.adapt = _lib.eb_adapt(1, -1, 10000.0, 100.0)
        job.adapt_time0, job.permute, job.randomize_split = 0, 1, 1
        if os.environ.get("EB_PROBE_STAMPS") and env.get("EB_HOST_PIPE") != "0":
            # timeline of the captured schedule: stamps on from the first call, so the graph contains them
            os.environ["EB_HOST_STAMPS"] = "1"
            for _ in range(4):
                _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
            os.environ.pop("EB_HOST_STAMPS")
            print(label, "(timelines above: direct, capture, replay, replay)", flush=True)
            continue
        for _ in range(5):
            _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
        t0 = time.perf_counter()
        for _ in range(n):
            _lib.check(lib.eb_run_host(ctypes.byref(job), 1), "eb_run_host")
        dt = (time.perf_counter() - t0) / n
        chk = float(coords.numpy().sum())
        print(f"{label:18s} {dt * 1e6:8.1f} us/call  {T * W / dt:.3e} wu/s  checksum {chk:.12e}", flush=True)


if __name__ == "__main__":
    main()
