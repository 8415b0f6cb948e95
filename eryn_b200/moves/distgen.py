"""DistributionGenerate on the device (reference: moves/distgen.py:34-104 over moves/mh.py:56-193): a Metropolis step
whose proposal redraws every active leaf from `generate_dist`; factors = log q(old) - log q(new)."""
import numpy as np

from .move import Move

__all__ = ["DistributionGenerate"]


class DistributionGenerate(Move):
    def __init__(self, generate_dist, *args, **kwargs):
        from ..prior import ProbDistContainer
        if not isinstance(generate_dist, dict):
            raise ValueError("When entering directly into the DistributionGenerate class, generate_dist must be a "
                             "dictionary. The keys are branch names and the items are ProbDistContainer objects.")
        for key in generate_dist:
            if not isinstance(generate_dist[key], ProbDistContainer):
                raise ValueError("Distributions need to be eryn.prior.ProbDistContainer object.")
        self.generate_dist = generate_dist
        super().__init__(*args, **kwargs)

    graphable = True

    def propose(self, model, state):
        ctx, d, host_state = self._enter(state)
        if not ctx.fused:
            raise NotImplementedError("DistributionGenerate needs a DeviceLikelihood functor in this build")
        if self.generate_dist[d.branch_name] is not ctx.priors:
            raise NotImplementedError("the device move generates from the sampler's (uniform) priors")
        T, W, L, D = d.shape
        cnt = self._count_buffer(ctx, T, W)
        prop = dict(kind="prior")
        if ctx.rng == "numpy-replay":
            n = T * W * L if d.inds is None else int(d.inds.sum().item())
            pts = self.generate_dist[d.branch_name].rvs(size=n)           # distgen.py:99: GLOBAL stream (prior.py:64)
            if d.inds is None:
                new = pts.reshape(T, W, L, D)
            else:
                new = np.zeros((T, W, L, D))
                new[d.inds.cpu().numpy().astype(bool)] = pts
            u_acc = model.random.rand(T, W)                                # mh.py:171
            acc = ctx.gaussian_step(d, prop, replay=(new, u_acc), accepted_count=cnt)
        else:
            acc = ctx.gaussian_step(d, prop, accepted_count=cnt)
        self.num_proposals += 1
        return self._exit(ctx, d, host_state, acc)
