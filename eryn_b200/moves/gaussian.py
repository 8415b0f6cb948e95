"""GaussianMove on the device (reference: moves/mh.py:56-193 + moves/gaussian.py:68-195)."""
import numpy as np

from .move import Move

__all__ = ["GaussianMove"]


class GaussianMove(Move):
    """Metropolis step with a Gaussian proposal: scalar (isotropic) or full covariance, mode 'vector'.

    The reference's vector-covariance form is broken upstream (gaussian.py:144 raises LinAlgError);
    it is rejected here with the same error class."""

    def __init__(self, cov_all, mode="vector", factor=None, **kwargs):
        if mode != "vector" or factor is not None:
            raise NotImplementedError("device GaussianMove implements mode='vector', factor=None")
        self.all_proposal = {}
        for name, cov in cov_all.items():
            try:
                float(cov)
            except TypeError:
                cov = np.atleast_1d(cov)
                if len(cov.shape) == 1:
                    raise np.linalg.LinAlgError("diagonal proposals are not usable in the reference (gaussian.py:144)")
                elif len(cov.shape) == 2 and cov.shape[0] == cov.shape[1]:
                    self.all_proposal[name] = dict(kind="matrix", cov=np.asarray(cov, dtype=np.float64),
                                                   chol=np.linalg.cholesky(cov))
                else:
                    raise ValueError("Invalid proposal scale dimensions")
            else:
                self.all_proposal[name] = dict(kind="scalar", scale=np.sqrt(cov))
        super().__init__(**kwargs)

    graphable = True

    def propose(self, model, state):
        ctx, d, host_state = self._enter(state)
        if not ctx.fused:
            raise NotImplementedError("GaussianMove needs a DeviceLikelihood functor in this build")
        T, W, L, D = d.shape
        prop = self.all_proposal[d.branch_name]
        cnt = self._count_buffer(ctx, T, W)
        splits = self._single_branch_splits(d.branch_name, L, D)
        self._nsplits_run = len(splits)
        acc = None
        for gi, bits, gnd in splits:  # mh.py:77-183: one Metropolis step per split
            kw = {} if (bits == 0 and gi == 0) else dict(gibbs=(bits, gi))
            if ctx.rng == "numpy-replay":
                n = T * W * L if d.inds is None else int(d.inds.sum().item())
                if prop["kind"] == "scalar":  # gaussian.py:166-167
                    inc = 1.0 * prop["scale"] * model.random.randn(n, D)
                else:  # gaussian.py:192-195
                    inc = 1.0 * model.random.multivariate_normal(np.zeros(D), prop["cov"], size=n)
                if d.inds is None:
                    delta = inc.reshape(T, W, L, D)
                else:
                    delta = np.zeros((T, W, L, D))
                    delta[d.inds.cpu().numpy().astype(bool)] = inc
                u_acc = model.random.rand(T, W)  # mh.py:171
                acc = ctx.gaussian_step(d, prop, replay=(delta, u_acc), accepted_count=cnt, **kw)
            else:
                acc = ctx.gaussian_step(d, prop, accepted_count=cnt, **kw)
            self.num_proposals += 1  # mh.py:188: per Gibbs split
        if acc is None:
            acc = ctx.accepted_mask(T, W)
            acc.zero_()
        return self._exit(ctx, d, host_state, acc)

    def _host_tick(self, n=1):
        self.num_proposals += n * getattr(self, "_nsplits_run", 1)
