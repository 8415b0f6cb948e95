"""Reversible-jump birth/death move on the device (reference: moves/rj.py:145-388 + moves/distgenrj.py:35-222)."""
import numpy as np

from .move import Move

__all__ = ["ReversibleJumpMove", "DistributionGenerateRJ"]


class ReversibleJumpMove(Move):
    def __init__(self, nleaves_max=None, nleaves_min=None, dr=None, dr_max_iter=5, tune=False, fix_change=None, **kwargs):
        super().__init__(**kwargs)
        if nleaves_max is None or nleaves_min is None:
            raise ValueError("Must provide nleaves_min and nleaves_max keyword arguments for RJ.")
        if not isinstance(nleaves_max, dict) or not isinstance(nleaves_min, dict):
            raise ValueError("nleaves_min and nleaves_max must be provided as dictionaries with keys as branch names and "
                             "values as the max or min leaf count.")
        if dr:
            raise NotImplementedError("Delayed Rejection will be implemented soon. Check for updated versions.")  # rj.py:350
        if fix_change not in [None]:
            raise NotImplementedError("fix_change is not part of the device move")
        self.nleaves_max, self.nleaves_min = nleaves_max, nleaves_min
        self.fix_change = None


class DistributionGenerateRJ(ReversibleJumpMove):
    """Birth from `generate_dist` (uniform priors), death of a random leaf; one leaf per branch and walker."""

    def __init__(self, generate_dist, *args, **kwargs):
        from ..prior import ProbDistContainer
        for key in generate_dist:
            if not isinstance(generate_dist[key], ProbDistContainer):
                raise ValueError("Distributions need to be eryn.prior.ProbDistContiner object.")
        self.generate_dist = generate_dist
        super().__init__(*args, **kwargs)

    def propose(self, model, state):
        from ..multibranch import MBContext, MBDeviceState
        ctx = self._context()
        if not isinstance(ctx, MBContext):
            raise RuntimeError("reversible jump runs on a multi-branch sampler")
        lay = ctx.layout
        for n in lay.branch_names:
            if self.generate_dist[n] is not ctx.priors[n]:
                raise NotImplementedError("the device move generates births from the sampler's priors")
            if self.nleaves_max[n] != lay.nleaves_max[n] or self.nleaves_min[n] != lay.nleaves_min[n]:
                raise ValueError("nleaves_min / nleaves_max differ from the sampler's")
        host_state = None
        d = state
        if not isinstance(state, MBDeviceState):
            host_state = state
            tc = self.temperature_control
            d = ctx.upload(state, betas=None if tc is None else tc.betas_dev)
        T, W = d.shape[:2]
        cnt = self._count_buffer(ctx, T, W)
        # Gibbs splits over branches (rj.py:168-343; rj_moves="iterate_branches" / "separate_branches" set them up,
        # ensemble.py:434-470): each split is a full birth/death proposal + Metropolis step restricted to its branches
        splits = []
        for split in self.gibbs_splits:
            if split is None:
                splits.append(list(lay.branch_names))
                continue
            for n, m in split.items():
                if n not in lay.branch_names:
                    raise KeyError(f"gibbs_sampling_setup names branch {n!r}; the sampler has {lay.branch_names}")
                if m is not None:
                    raise ValueError("inputting gibbs indexing at the leaf/parameter level is not allowed with an RJ "
                                     "proposal. Only branch names.")
            splits.append(list(split.keys()))
        acc = None
        for gi, names_run in enumerate(splits):
            mask = 0 if len(names_run) == len(lay.branch_names) and gi == 0 and len(splits) == 1 else \
                sum(1 << lay.branch_names.index(n) for n in names_run)
            last = gi == len(splits) - 1
            # rj.py:385-386: the move counts the accepts of the LAST split only
            kw = dict(accepted_count=cnt if last else None, branch_mask=mask, gibbs_index=gi)
            if ctx.rng == "numpy-replay":
                flags = d.flags_host()
                nb = len(lay.branch_names)
                change_all = np.zeros((nb, T, W), dtype=np.int32)
                leaf_all = np.zeros((nb, T, W), dtype=np.int32)
                for n in names_run:                                              # distgenrj.py:160-173
                    b = lay.branch_names.index(n)
                    nmin, nmax = lay.nleaves_min[n], lay.nleaves_max[n]
                    if nmin == nmax:
                        continue
                    f = flags[n]
                    nleaves = f.sum(axis=-1)
                    change = model.random.choice([-1, +1], size=nleaves.shape)   # :62
                    change = change * ((nleaves != nmin) & (nleaves != nmax)) + (+1) * (nleaves == nmin) \
                        + (-1) * (nleaves == nmax)                               # :67-71
                    for t in range(T):                                           # :85-121, same draw order
                        for w in range(W):
                            if change[t, w] == +1:
                                leaf_all[b, t, w] = model.random.choice(np.where(~f[t, w])[0])
                            elif change[t, w] == -1:
                                leaf_all[b, t, w] = model.random.choice(np.where(f[t, w])[0])
                    change_all[b] = change
                births = [None] * nb
                for n in names_run:                                              # distgenrj.py:176-219
                    b = lay.branch_names.index(n)
                    nmin, nmax = lay.nleaves_min[n], lay.nleaves_max[n]
                    if nmin == nmax:
                        continue
                    sel = change_all[b] == +1
                    full = np.zeros((T, W, lay.ndims[n]))
                    full[sel] = self.generate_dist[n].rvs(size=int(sel.sum()))   # prior.py:56-71: GLOBAL stream
                    births[b] = full
                u_acc = model.random.rand(T, W)                                  # rj.py:332
                acc = ctx.rj_step(d, replay=(change_all, leaf_all, births, u_acc), **kw)
            else:
                acc = ctx.rj_step(d, **kw)
        self.num_proposals += 1
        tc = self.temperature_control
        if tc is not None:
            d = tc.temper_comps(d, adapt=False)                              # rj.py:381-382
        else:
            ctx.advance_iter()
        if host_state is None:
            return d, acc
        return ctx.download(d, into=host_state), acc.cpu().numpy().astype(bool)
