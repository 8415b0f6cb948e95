"""Device log-likelihood functors (selected by enum; evaluated inside the fused kernels).

An arbitrary Python callable cannot be fused into a CUDA kernel; the built-in functors cover
BASELINE.json's synthetic targets.  A callable acting on CUDA tensors
(`TorchLikelihood`) goes through the split path: propose kernel -> callable -> accept kernel."""
import numpy as np

from . import _lib

__all__ = ["GaussianLikelihood", "RosenbrockLikelihood", "GaussianMixtureLikelihood", "TorchLikelihood"]


class DeviceLikelihood(object):
    kind = None
    ncomp = 0

    def params(self):
        return np.zeros(0)


class GaussianLikelihood(DeviceLikelihood):
    """log L = -1/2 (x - mu)^T P (x - mu);  P is the precision (inverse covariance) matrix."""

    kind = _lib.EB_LIKE_GAUSSIAN

    def __init__(self, mu, invcov):
        self.mu = np.asarray(mu, dtype=np.float64).ravel()
        self.invcov = np.ascontiguousarray(invcov, dtype=np.float64)
        d = self.mu.shape[0]
        if self.invcov.shape != (d, d):
            raise ValueError("invcov must be (ndim, ndim)")

    def params(self):
        return np.concatenate([self.mu, self.invcov.ravel()])


class RosenbrockLikelihood(DeviceLikelihood):
    """log L = -sum_{i<d-1} [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2]."""

    kind = _lib.EB_LIKE_ROSENBROCK


class GaussianMixtureLikelihood(DeviceLikelihood):
    """log L = log sum_k w_k N(x; mu_k, sigma_k^2 I)."""

    kind = _lib.EB_LIKE_GMIX

    def __init__(self, mus, sigmas, weights=None):
        self.mus = np.ascontiguousarray(mus, dtype=np.float64)
        K, D = self.mus.shape
        self.sigmas = np.asarray(sigmas, dtype=np.float64).ravel()
        self.weights = np.full(K, 1.0 / K) if weights is None else np.asarray(weights, dtype=np.float64).ravel()
        if self.sigmas.shape != (K,) or self.weights.shape != (K,):
            raise ValueError("sigmas/weights must have one entry per component")
        self.ncomp = K
        self.logc = np.log(self.weights) - D * np.log(self.sigmas) - 0.5 * D * np.log(2.0 * np.pi)
        self.hinv = 0.5 / self.sigmas ** 2

    def params(self):
        return np.concatenate([self.logc, self.hinv, self.mus.ravel()])


class TorchLikelihood(object):
    """Wrap `fn(x: cuda float64 tensor [N, nleaves, ndim]) -> tensor [N]` for the split path."""

    kind = None

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x):
        return self.fn(x)
