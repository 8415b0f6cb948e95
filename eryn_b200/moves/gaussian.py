"""GaussianMove on the device (reference: moves/mh.py:56-193 + moves/gaussian.py:68-195)."""
import numpy as np

from .move import Move

__all__ = ["GaussianMove"]


class GaussianMove(Move):
    """Metropolis step with a Gaussian proposal (gaussian.py:8-66): scalar (isotropic) or full covariance.
    Scalar proposals take mode "vector" (all dimensions), "random" (one random dimension per walker) or "sequential"
    (the next dimension for everybody) and `factor` (scale multiplied by exp(U(-log factor, log factor)), one draw per
    call); matrix proposals take mode "vector" only, as in the reference (gaussian.py:184-185).

    The reference's vector-covariance (diagonal) form is broken upstream (gaussian.py:144 raises LinAlgError); it is
    rejected here with the same error class."""

    def __init__(self, cov_all, mode="vector", factor=None, **kwargs):
        if factor is not None and factor < 1.0:
            raise ValueError("'factor' must be >= 1.0")  # gaussian.py:152-153
        self.mode, self.factor = mode, factor
        self._seq_index = 0  # gaussian.py:137, :176
        self.all_proposal = {}
        for name, cov in cov_all.items():
            try:
                float(cov)
            except TypeError:
                cov = np.atleast_1d(cov)
                if len(cov.shape) == 1:
                    raise np.linalg.LinAlgError("diagonal proposals are not usable in the reference (gaussian.py:144)")
                elif len(cov.shape) == 2 and cov.shape[0] == cov.shape[1]:
                    if mode not in ["vector"]:
                        raise ValueError("'{0}' is not a recognized mode. Please select from: {1}".format(mode, ["vector"]))
                    self.all_proposal[name] = dict(kind="matrix", cov=np.asarray(cov, dtype=np.float64),
                                                   chol=np.linalg.cholesky(cov))
                else:
                    raise ValueError("Invalid proposal scale dimensions")
            else:
                if mode not in ["vector", "random", "sequential"]:
                    raise ValueError("'{0}' is not a recognized mode. Please select from: {1}".format(
                        mode, ["vector", "random", "sequential"]))
                self.all_proposal[name] = dict(kind="scalar", scale=np.sqrt(cov))
        super().__init__(**kwargs)
        if self.gibbs_sampling_setup is not None and mode != "vector":
            raise NotImplementedError("Gibbs splits combine with mode='vector' on the device path")

    @property
    def graphable(self):
        return self.mode != "sequential"  # the dimension counter advances on the host

    def propose(self, model, state):
        ctx, d, host_state = self._enter(state)
        if not ctx.fused:
            raise NotImplementedError("GaussianMove needs a DeviceLikelihood functor in this build")
        T, W, L, D = d.shape
        prop = self.all_proposal[d.branch_name]
        cnt = self._count_buffer(ctx, T, W)
        splits = self._single_branch_splits(d.branch_name, L, D)
        self._nsplits_run = len(splits)
        lf = 0.0 if self.factor is None else float(np.log(self.factor))
        acc = None
        for gi, bits, gnd in splits:  # mh.py:77-183: one Metropolis step per split
            gibbs = None if (bits == 0 and gi == 0) else (bits, gi)
            seq_dim = None
            if self.mode == "sequential":  # gaussian.py:174-176: the next dimension, for every walker
                seq_dim = self._seq_index % D
                self._seq_index = (self._seq_index + 1) % D
            if ctx.rng == "numpy-replay":
                n = T * W * L if d.inds is None else int(d.inds.sum().item())
                f = 1.0 if self.factor is None else np.exp(model.random.uniform(-lf, lf))  # gaussian.py:161-164, drawn first
                if prop["kind"] == "scalar":  # gaussian.py:166-167
                    inc = f * prop["scale"] * model.random.randn(n, D)
                else:  # gaussian.py:192-195
                    inc = f * model.random.multivariate_normal(np.zeros(D), prop["cov"], size=n)
                if self.mode == "random":  # gaussian.py:172-173
                    m = model.random.randint(D, size=n)
                    only = np.zeros((n, D), dtype=bool)
                    only[np.arange(n), m] = True
                    inc = np.where(only, inc, 0.0)
                elif seq_dim is not None:
                    only = np.zeros((n, D), dtype=bool)
                    only[:, seq_dim] = True
                    inc = np.where(only, inc, 0.0)
                if d.inds is None:
                    delta = inc.reshape(T, W, L, D)
                else:
                    delta = np.zeros((T, W, L, D))
                    delta[d.inds.cpu().numpy().astype(bool)] = inc
                u_acc = model.random.rand(T, W)  # mh.py:171
                acc = ctx.gaussian_step(d, prop, replay=(delta, u_acc), accepted_count=cnt, gibbs=gibbs)
            else:
                if seq_dim is not None:
                    gibbs = (1 << seq_dim, gi)  # one dimension moves: the kernel's parameter mask
                acc = ctx.gaussian_step(d, prop, accepted_count=cnt, gibbs=gibbs,
                                        dim_mode=1 if self.mode == "random" else 0, log_factor=lf)
            self.num_proposals += 1  # mh.py:188: per Gibbs split
        if acc is None:
            acc = ctx.accepted_mask(T, W)
            acc.zero_()
        return self._exit(ctx, d, host_state, acc)

    def _host_tick(self, n=1):
        self.num_proposals += n * getattr(self, "_nsplits_run", 1)
