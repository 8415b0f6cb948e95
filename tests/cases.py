"""Shared definitions of the golden cases (what tests/golden/make_golden.py ran the reference on)."""
import os

import numpy as np

from oracle import eryn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def corr_prec(d, seed=99):
    A = np.random.RandomState(seed).randn(d, d)
    cov = A @ A.T / d + np.eye(d)
    return np.linalg.inv(cov)


def gmask(d, *params):
    """Gibbs split of a single-leaf branch: boolean [nleaves_max = 1, ndim] with the given parameters selected"""
    m = np.zeros((1, d), dtype=bool)
    m[0, list(params)] = True
    return m


COV3 = np.array([[0.04, 0.01, 0.0], [0.01, 0.09, -0.02], [0.0, -0.02, 0.01]])

# name -> (likelihood factory(ndim), moves, weights, tempering extras)
CASES = {
    "c1_kat1": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), moves=[dict(kind="stretch", a=2.0)]),
    "pt_kat2": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), moves=[dict(kind="stretch", a=2.0)]),
    "c2_small": dict(like=lambda d: orc.GaussianLike(np.zeros(d), corr_prec(d)), moves=[dict(kind="stretch", a=2.0)]),
    "c2_tightprior": dict(like=lambda d: orc.GaussianLike(np.zeros(d), corr_prec(d)),
                          moves=[dict(kind="stretch", a=2.0)]),
    "c3_small": dict(like=lambda d: orc.RosenbrockLike(),
                     moves=[dict(kind="stretch", a=2.0),
                            dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.01)))],
                     weights=[0.5, 0.5]),
    "gauss_matrix": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                         moves=[dict(kind="gaussian", proposal=dict(kind="matrix", cov=COV3))]),
    "odd_walkers": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), moves=[dict(kind="stretch", a=1.5)]),
    "periodic_mix": dict(like=lambda d: orc.GaussianLike(np.array([0.2, 6.1, 3.0]), np.eye(3) / 0.49),
                         moves=[dict(kind="stretch", a=2.0),
                                dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.25)))],
                         weights=[0.5, 0.5], periods=np.array([2 * np.pi, 2 * np.pi, 0.0])),
    "distgen_mix": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                        moves=[dict(kind="stretch", a=2.0), dict(kind="distgen")], weights=[0.5, 0.5]),
    "combine_sg": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                       moves=[dict(kind="combine", moves=[dict(kind="stretch", a=2.0),
                                                          dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.25)))])]),
    "gibbs_mix": dict(like=lambda d: orc.GaussianLike(np.zeros(d), corr_prec(d)),
                      moves=[dict(kind="stretch", a=2.0, gibbs=[gmask(4, 0, 1), gmask(4, 2, 3)]),
                             dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.09)),
                                  gibbs=[gmask(4, 0), gmask(4, 1, 2), None])],
                      weights=[0.5, 0.5]),
    "gauss_modes": dict(like=lambda d: orc.GaussianLike(np.zeros(d), corr_prec(d)),
                        moves=[dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.16), mode="random", factor=2.0)),
                               dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.25), mode="sequential")),
                               dict(kind="gaussian", proposal=dict(kind="scalar", scale=np.sqrt(0.04), mode="vector"))],
                        weights=[0.5, 0.3, 0.2]),
    "mt_mix": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d) / 0.25),
                   moves=[dict(kind="stretch", a=2.0), dict(kind="mt", num_try=6)], weights=[0.4, 0.6]),
    "tiny_live": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                      moves=[dict(kind="stretch", a=2.0, live_dangerously=True)]),
    "one_temp": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), moves=[dict(kind="stretch", a=2.0)]),
    "nosplit": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                    moves=[dict(kind="stretch", a=2.0, randomize_split=False)]),
    "stop_adapt": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)), moves=[dict(kind="stretch", a=2.0)],
                       tempering=dict(stop_adaptation=6, adaptation_lag=30, adaptation_time=4)),
    "noadapt_noperm": dict(like=lambda d: orc.GaussianLike(np.zeros(d), np.eye(d)),
                           moves=[dict(kind="stretch", a=2.0)], adaptive=False, permute=False),
}


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    T, W = int(g["ntemps"]), int(g["nwalkers"])
    g["accepted"] = np.unpackbits(g["accepted"], axis=-1)[..., :W].astype(bool)
    if "acc_count" in g:  # combined moves: accept COUNT of every walker per iteration
        g["accepted"] = g["acc_count"].astype(np.int64)
    return g


def seeded_streams(seed):
    """The reference's two streams after `np.random.seed(seed)` + sampler construction
    (ensemble.py:604,651-652: the private stream is a copy of the global state)."""
    glob = np.random.RandomState(seed)
    private = np.random.RandomState()
    private.set_state(glob.get_state())
    return private, glob
