"""Priors the device path understands: independent uniform boxes (prior.py:12-91, 219-497).

`logpdf` is evaluated on the device (csrc/likelihoods.cuh: box_logpdf_leaf); `rvs` stays on the
host and consumes the global NumPy stream exactly like the reference (prior.py:56-71, 432-497)."""
import numpy as np

__all__ = ["UniformDistribution", "uniform_dist", "ProbDistContainer"]


class UniformDistribution(object):
    def __init__(self, min_val, max_val):
        if min_val > max_val:
            min_val, max_val = max_val, min_val
        elif min_val == max_val:
            raise ValueError("Min and max values are the same.")
        self.min_val, self.max_val = min_val, max_val
        self.diff = max_val - min_val
        self.pdf_val = 1 / self.diff
        self.logpdf_val = np.log(self.pdf_val)

    def rvs(self, size=1):
        if not isinstance(size, int) and not isinstance(size, tuple):
            raise ValueError("size must be an integer or tuple of ints.")
        if isinstance(size, int):
            size = (size,)
        return np.random.rand(*size) * self.diff + self.min_val


def uniform_dist(min, max):
    return UniformDistribution(min, max)


class ProbDistContainer:
    """{int index: UniformDistribution}.  Other distributions need the split path with a user
    prior evaluated on device tensors (not part of this build)."""

    def __init__(self, priors_in):
        self.priors_in = dict(priors_in)
        keys = list(self.priors_in.keys())
        if not all(isinstance(k, (int, np.integer)) for k in keys):
            raise ValueError("Keys for prior dictionary must be integers on the device path.")
        if sorted(keys) != list(range(len(keys))):
            raise ValueError("Please ensure all sampled parameters are included in priors.")
        for d in self.priors_in.values():
            if not isinstance(d, UniformDistribution):
                raise NotImplementedError("device priors are uniform boxes (uniform_dist)")
        self.ndim = len(keys)
        self.priors = [[np.array([k]), self.priors_in[k]] for k in keys]  # insertion order, as the reference
        self.key_order = list(range(self.ndim))

    def arrays(self):
        """(lo, hi, logpdf_val) in parameter-index order."""
        lo = np.array([self.priors_in[i].min_val for i in range(self.ndim)], dtype=np.float64)
        hi = np.array([self.priors_in[i].max_val for i in range(self.ndim)], dtype=np.float64)
        lp = np.array([self.priors_in[i].logpdf_val for i in range(self.ndim)], dtype=np.float64)
        return lo, hi, lp

    def rvs(self, size=1):
        if isinstance(size, int):
            size = (size,)
        elif not isinstance(size, tuple):
            raise ValueError("Size must be int or tuple of ints.")
        out = np.zeros(size + (self.ndim,))
        for inds, prior_i in self.priors:
            out[..., inds[0]] = prior_i.rvs(size=size)
        return out
