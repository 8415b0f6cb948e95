"""MTDistGenMove on the device (reference: moves/multipletry.py:62-514 + moves/mtdistgen.py:8-133 over moves/mh.py:56-193):
multiple-try Metropolis with an independent proposal drawn from `generate_dist`."""
import numpy as np

from .move import Move

__all__ = ["MTDistGenMove"]


class MTDistGenMove(Move):
    """Same constructor surface as eryn.moves.MTDistGenMove: `generate_dist` (a ProbDistContainer, as the reference's
    test passes it, or {branch: ProbDistContainer}), num_try, independent, symmetric, rj.  The device kernel implements
    the independent form (`independent=True`: the auxiliary set reuses the tries, multipletry.py:383-416) with the
    sampler's uniform priors as the generating distribution."""

    graphable = True

    def __init__(self, generate_dist, num_try=1, independent=False, symmetric=False, rj=False, **kwargs):
        from ..prior import ProbDistContainer
        if rj:
            if symmetric or independent:
                raise ValueError("If rj==True, symmetric and independt must both be False.")  # multipletry.py:93-97
            raise NotImplementedError("nested reversible-jump multiple try (MTDistGenMoveRJ) is not part of the device path")
        if not independent or symmetric:
            raise NotImplementedError("the device multiple-try move draws from a distribution that does not depend on the "
                                      "current point: independent=True, symmetric=False (mtdistgen.py:10-12)")
        if isinstance(generate_dist, dict):
            for key in generate_dist:
                if not isinstance(generate_dist[key], ProbDistContainer):
                    raise ValueError("Distributions need to be eryn.prior.ProbDistContainer object.")
        elif not isinstance(generate_dist, ProbDistContainer):
            raise ValueError("Distributions need to be eryn.prior.ProbDistContainer object.")
        self.generate_dist = generate_dist
        self.num_try = int(num_try)
        self.independent, self.symmetric, self.rj = independent, symmetric, rj
        if self.num_try < 1:
            raise ValueError("num_try must be >= 1")
        super().__init__(**kwargs)
        if self.gibbs_sampling_setup is not None:
            raise NotImplementedError("Gibbs splits of the multiple-try move are not part of the device path")

    def propose(self, model, state):
        ctx, d, host_state = self._enter(state)
        if not ctx.fused:
            raise NotImplementedError("MTDistGenMove needs a DeviceLikelihood functor in this build")
        gd = self.generate_dist[d.branch_name] if isinstance(self.generate_dist, dict) else self.generate_dist
        if gd is not ctx.priors:
            raise NotImplementedError("the device move generates from the sampler's (uniform) priors")
        T, W, L, D = d.shape
        if L != 1 or d.inds is not None:
            raise ValueError("multiple try proposes one present leaf per walker (multipletry.py:545-549)")
        cnt = self._count_buffer(ctx, T, W)
        if ctx.rng == "numpy-replay":
            tries = gd.rvs(size=(T * W, self.num_try))      # mtdistgen.py:58: GLOBAL stream, one rand(n, num_try) per parameter
            u_sel = np.random.rand(T * W)                   # multipletry.py:51: GLOBAL stream
            u_acc = model.random.rand(T, W)                 # mh.py:171
            acc = ctx.mt_distgen_step(d, self.num_try, replay=(tries, u_sel, u_acc), accepted_count=cnt)
        else:
            acc = ctx.mt_distgen_step(d, self.num_try, accepted_count=cnt)
        self.num_proposals += 1
        return self._exit(ctx, d, host_state, acc)
