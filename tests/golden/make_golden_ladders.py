"""Golden ladders from the UNMODIFIED reference (build container only, like make_golden.py):

    python tests/golden/make_golden_ladders.py

* make_ladder (tempering.py:10-197) on a grid of (ndim, ntemps, Tmax);
* TemperatureControl.adapt_temps (tempering.py:563-596) driven by synthetic swap counts for 300 steps from time 0, the
  default adaptation lag / time and one non-default pair — the ladder after every step is recorded.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: F401,E402  (installs the import shim of SURVEY.md App. A)
from eryn.moves.tempering import TemperatureControl, make_ladder  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = {}
    grid = []
    for ndim in (1, 2, 3, 5, 8, 20, 60, 100, 101, 150):
        for ntemps in (1, 2, 4, 16, 128):
            for tmax in (None, np.inf, 50.0):
                try:
                    b = make_ladder(ndim, ntemps=ntemps, Tmax=tmax)
                except Exception as e:  # record which combinations the reference rejects
                    b = np.array([np.nan])
                    print("reference raises for", ndim, ntemps, tmax, type(e).__name__)
                grid.append((ndim, ntemps, -1.0 if tmax is None else tmax))
                out[f"ladder_{len(grid) - 1}"] = np.asarray(b, dtype=np.float64)
    out["grid"] = np.asarray(grid, dtype=np.float64)
    rng = np.random.RandomState(4)
    for k, (T, W, lag, t0) in enumerate([(4, 16, 10000, 100), (16, 4096, 10000, 100), (8, 64, 50, 5)]):
        tc = TemperatureControl(3, W, ntemps=T, adaptive=True, adaptation_lag=lag, adaptation_time=t0)
        counts = rng.binomial(W, rng.uniform(0.1, 0.9, size=T - 1), size=(300, T - 1))
        hist = []
        for step in range(300):
            tc.swaps_accepted = counts[step].astype(float)
            tc.adapt_temps()
            hist.append(tc.betas.copy())
        out[f"adapt_{k}_cfg"] = np.array([T, W, lag, t0], dtype=np.float64)
        out[f"adapt_{k}_counts"] = counts
        out[f"adapt_{k}_betas0"] = make_ladder(3, ntemps=T)
        out[f"adapt_{k}_hist"] = np.stack(hist)
    np.savez_compressed(os.path.join(HERE, "ladders.npz"), **out)
    print("ladders.npz", os.path.getsize(os.path.join(HERE, "ladders.npz")), "bytes;", len(grid), "ladders")
