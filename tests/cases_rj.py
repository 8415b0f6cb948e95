"""Shared definitions of the config-5 golden cases (what tests/golden/make_golden_rj.py ran the reference on)."""
import os

import numpy as np

from oracle import eryn_oracle as orc
from oracle import rj_oracle as rjo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["c5_small", "c5_wide", "c5_iter", "c5_sep"]   # the last two: rj_moves="iterate_branches" / "separate_branches"
PRIOR_BOUNDS = {  # tests/test_eryn.py:416-427
    "gauss": lambda t: ([2.5, t.min(), 0.01], [3.5, t.max(), 0.21]),
    "sine": lambda t: ([0.5, 1.0, 0.0], [1.5, 20.0, 2 * np.pi]),
}
GINJ = np.array([[3.3, -0.2, 0.1], [2.6, -0.1, 0.1], [3.4, 0.0, 0.1], [2.9, 0.3, 0.1]])
SINJ = np.array([[1.3, 10.1, 1.0], [0.8, 4.6, 1.2]])


def rj_mode(g):
    m = str(g["rj_mode"]) if "rj_mode" in g else "True"
    return "together" if m == "True" else m


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    W = int(g["W"])
    for k, n in (("ig", int(g["Lg"])), ("is", int(g["Ls"])), ("acc", W), ("rjacc", W)):
        g[k] = np.unpackbits(g[k], axis=-1)[..., :n].astype(bool)
    return g


def replay_setup(g):
    """Re-draw what make_golden_rj.py drew from the global stream before the sampler existed; returns the initial
    coords/inds and the (private, global) streams in the state the reference's sampler started from."""
    T, W, Lg, Ls, nt = int(g["T"]), int(g["W"]), int(g["Lg"]), int(g["Ls"]), int(g["nt"])
    np.random.seed(int(g["seed"]))
    t = np.linspace(-1, 1, nt)
    ginj, sinj = GINJ[: min(4, Lg - 1)], SINJ[: min(2, Ls - 1)]
    noise = float(g["sigma"]) * np.random.randn(nt)
    coords = [np.zeros((T, W, Lg, 3)), np.zeros((T, W, Ls, 3))]
    inds = [np.zeros((T, W, Lg), dtype=bool), np.zeros((T, W, Ls), dtype=bool)]
    for b, inj in enumerate((ginj, sinj)):
        for nn in range(len(inj)):
            coords[b][:, :, nn] = np.random.multivariate_normal(inj[nn], np.diag(np.ones(3) * 1e-4), size=(T, W))
            inds[b][:, :, nn] = True
    assert np.array_equal(coords[0], g["cg0"]) and np.array_equal(coords[1], g["cs0"])
    glob = np.random.RandomState()
    glob.set_state(np.random.get_state())
    private = np.random.RandomState()
    private.set_state(glob.get_state())
    return t, coords, inds, private, glob


def priors_for(t):
    return [orc.BoxPrior(*PRIOR_BOUNDS["gauss"](t)), orc.BoxPrior(*PRIOR_BOUNDS["sine"](t))]


def oracle_sampler(g, streams, t=None):
    t = np.asarray(g["t"]) if t is None else t
    like = rjo.PulseLike(t, g["y"], float(g["sigma"]), [0, 1])
    T = int(g["T"])
    ndim_total = 3 * (int(g["Lg"]) + int(g["Ls"]))
    # ensemble.py:321-334: the ladder is built for the total dimension over all branches
    return rjo.OracleSamplerMB(priors_for(t), like, [0, 0], [int(g["Lg"]), int(g["Ls"])], streams,
                               betas=orc.make_ladder_default(ndim_total, T), nfriends=int(g["nfriends"]),
                               n_iter_update=int(g["n_iter_update"]), rj_mode=rj_mode(g))
