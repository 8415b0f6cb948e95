"""TemperatureControl on the device (reference: moves/tempering.py)."""
import numpy as np
import torch

from ..device import DeviceState

__all__ = ["TemperatureControl", "make_ladder"]

# temperature step for a 25% swap acceptance on a Gaussian posterior, per dimension (ptemcee's table,
# tempering.py:57-158)
_TSTEP = np.array([
    25.2741, 7.0, 4.47502, 3.5236, 3.0232, 2.71225, 2.49879, 2.34226, 2.22198, 2.12628,
    2.04807, 1.98276, 1.92728, 1.87946, 1.83774, 1.80096, 1.76826, 1.73895, 1.7125, 1.68849,
    1.66657, 1.64647, 1.62795, 1.61083, 1.59494, 1.58014, 1.56632, 1.55338, 1.54123, 1.5298,
    1.51901, 1.50881, 1.49916, 1.49, 1.4813, 1.47302, 1.46512, 1.45759, 1.45039, 1.4435,
    1.4369, 1.43056, 1.42448, 1.41864, 1.41302, 1.40761, 1.40239, 1.39736, 1.3925, 1.38781,
    1.38327, 1.37888, 1.37463, 1.37051, 1.36652, 1.36265, 1.35889, 1.35524, 1.3517, 1.34825,
    1.3449, 1.34164, 1.33847, 1.33538, 1.33236, 1.32943, 1.32656, 1.32377, 1.32104, 1.31838,
    1.31578, 1.31325, 1.31076, 1.30834, 1.30596, 1.30364, 1.30137, 1.29915, 1.29697, 1.29484,
    1.29275, 1.29071, 1.2887, 1.28673, 1.2848, 1.28291, 1.28106, 1.27923, 1.27745, 1.27569,
    1.27397, 1.27227, 1.27061, 1.26898, 1.26737, 1.26579, 1.26424, 1.26271, 1.26121, 1.25973,
])


def make_ladder(ndim, ntemps=None, Tmax=None):
    """Geometric ladder of betas (tempering.py:10-197): same branches, same errors."""
    if type(ndim) != int or ndim < 1:
        raise ValueError("Invalid number of dimensions specified.")
    if ntemps is None and Tmax is None:
        raise ValueError("Must specify one of ``ntemps`` and ``Tmax``.")
    if Tmax is not None and Tmax <= 1:
        raise ValueError("``Tmax`` must be greater than 1.")
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError("Invalid number of temperatures specified.")
    tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim) if ndim > _TSTEP.shape[0] else _TSTEP[ndim - 1]
    append_inf = False
    if Tmax == np.inf:
        append_inf, Tmax, ntemps = True, None, ntemps - 1
    if ntemps is not None:
        if Tmax is None:
            Tmax = tstep ** (ntemps - 1)
    else:
        if Tmax is None:
            raise ValueError("Must specify at least one of ``ntemps and finite ``Tmax``.")
        ntemps = int(np.log(Tmax) / np.log(tstep) + 2)
    betas = np.logspace(0, -np.log10(Tmax), ntemps)
    if append_inf:
        betas = np.concatenate((betas, [0]))
    return betas


class TemperatureControl(object):
    """Tempered posterior, swap ladder and ladder adaptation (tempering.py:200-649).

    The ladder lives on the device (`betas_dev`); `.betas`, `.swaps_accepted` and `.time` read it back."""

    def __init__(self, effective_ndim, nwalkers, ntemps=1, betas=None, Tmax=None, adaptive=True,
                 adaptation_lag=10000, adaptation_time=100, stop_adaptation=-1, permute=True,
                 skip_swap_supp_names=[]):
        if betas is None:
            betas = np.array([1.0]) if ntemps == 1 else make_ladder(effective_ndim, ntemps=ntemps, Tmax=Tmax)
        self.nwalkers = nwalkers
        self._betas_host = np.asarray(betas, dtype=np.float64).copy()
        self.ntemps = len(self._betas_host)
        self.permute = permute
        self.adaptive = adaptive
        self.adaptation_time, self.adaptation_lag = adaptation_time, adaptation_lag
        self.stop_adaptation = stop_adaptation
        self.swaps_proposed = np.full(self.ntemps - 1, self.nwalkers)
        self.ctx = None
        self._betas_dev = None
        self._time0 = 0

    def bind(self, ctx):
        self.ctx = ctx
        self._betas_dev = torch.from_numpy(self._betas_host.copy()).to(ctx.device)
        ctx.write_ctrl(time=self._time0)

    # ---- mirrors of the reference attributes ---------------------------------------------------------
    @property
    def betas_dev(self):
        if self._betas_dev is None:
            raise RuntimeError("TemperatureControl is not bound to a DeviceContext")
        return self._betas_dev

    @property
    def betas(self):
        if self._betas_dev is None:
            return self._betas_host
        self.ctx.flush_adapt()   # a pass may have left its ladder adaptation to the next stretch kernel
        return self._betas_dev.cpu().numpy()

    @betas.setter
    def betas(self, b):
        b = np.asarray(b, dtype=np.float64)
        if b.shape != (self.ntemps,):
            raise ValueError("betas has the wrong number of temperatures")
        self._betas_host = b.copy()
        if self._betas_dev is not None:
            self.ctx.flush_adapt()
            self._betas_dev.copy_(torch.from_numpy(self._betas_host))

    @property
    def time(self):
        return self._time0 if self.ctx is None else int(self.ctx.read_ctrl().time)

    @time.setter
    def time(self, t):
        self._time0 = int(t)
        if self.ctx is not None:
            self.ctx.write_ctrl(time=t)

    @property
    def swaps_accepted(self):
        c = self.ctx.read_ctrl()
        return np.array(c.swaps_accepted[: self.ntemps - 1], dtype=np.float64)

    def compute_log_posterior_tempered(self, logl, logp, betas=None):
        """tempering.py:284-349 for host arrays (diagnostics; the kernels apply the same rule)."""
        assert logl.shape == logp.shape
        b = self.betas if betas is None else betas
        with np.errstate(invalid="ignore"):
            loglT = logl * (b if logl.ndim == 1 else b[:, None])
        loglT[np.isnan(loglT)] = -np.inf
        return loglT + logp

    # ---- the swap pass ---------------------------------------------------------------------------------
    def temper_comps(self, state, adapt=True):
        """Swap ladder + adaptation (tempering.py:598-649).  DeviceState in -> same DeviceState out."""
        ctx = self.ctx
        host_state = None
        d = state
        if not isinstance(state, DeviceState):
            host_state = state
            d = ctx.upload(state, betas=self.betas_dev)
        T, W, L, D = d.shape
        ad = None
        if adapt and self.adaptive and T > 1:
            ad = dict(adaptive=True, stop_adaptation=self.stop_adaptation, adaptation_lag=self.adaptation_lag,
                      adaptation_time=self.adaptation_time)
        replay = None
        if ctx.rng == "numpy-replay" and T > 1:
            iperm = np.zeros((T, W), dtype=np.int32)
            i1perm = np.zeros((T, W), dtype=np.int32)
            u = np.ones((T, W))
            for i in range(T - 1, 0, -1):  # tempering.py:515-535, global NumPy stream
                if self.permute:
                    iperm[i] = np.random.permutation(W)
                    i1perm[i] = np.random.permutation(W)
                else:
                    iperm[i] = np.arange(W)
                    i1perm[i] = np.arange(W)
                u[i] = np.random.uniform(size=W)
            replay = (iperm, i1perm, u)
        ctx.pt_swap(d, permute=self.permute, adapt=ad, replay=replay)
        if host_state is not None:
            return ctx.download(d, into=host_state)
        return d
