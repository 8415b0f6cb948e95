"""GroupStretchMove on the device (reference: moves/group.py:122-281 + moves/groupstretch.py:34-120).

In the reference GroupMove is abstract: the user supplies `setup_friends` / `fix_friends` / `find_friends` in Python
(group.py:50-95).  Python callbacks cannot run inside a kernel, so the device move ships ONE concrete friend rule —
the one of the reference's own test fixture `MeanGaussianGroupMove` (tests/test_eryn.py:813-907), for every branch:

  * every `n_iter_update` iterations the stationary friends of a branch are the cold chain's active leaves, made unique
    and sorted by one parameter (`friend_key`, default index 1); every active leaf stores the indices of its `nfriends`
    nearest friends along that parameter (host: np.unique + sort of <= nwalkers*nleaves values; device: nearest-K);
  * leaves born through reversible jump in between get their row on the next call (`fix_friends`);
  * a proposal picks one stored friend per active leaf at random and stretches the whole walker with ONE z
    (stretch.py:129-145); inactive leaves stretch against 0 like the reference's zero-filled buffer.

The friend table travels with the walker through the swap pass, as the reference's BranchSupplemental does
(tempering.py:351-482)."""
import numpy as np

from .move import Move

__all__ = ["GroupStretchMove"]


class GroupStretchMove(Move):
    def __init__(self, nfriends=None, n_iter_update=100, a=2.0, friend_key=1, live_dangerously=False, **kwargs):
        super().__init__(**kwargs)
        if self.gibbs_sampling_setup is not None:
            raise NotImplementedError("Gibbs splits of the group move are not part of the device path (DESIGN.md)")
        if nfriends is None:
            raise TypeError("int() argument must be a string, a bytes-like object or a real number, not 'NoneType'")  # group.py:43
        self.nfriends = int(nfriends)
        self.n_iter_update = n_iter_update
        if self.n_iter_update <= 1 and not live_dangerously:
            raise ValueError("n_iter_update must be greather than or equal to 2.")  # group.py:46-47
        self.a = a
        self.friend_key = friend_key
        self.iter = 0
        self._friends = None  # (eb_mb_friends, keep-alive tensors)
        self.friends, self.means = {}, {}

    # ---- fixture rule, host part: unique + sorted cold-chain leaves (tests/test_eryn.py:818-831) ----------
    def setup_friends(self, ctx, d):
        lay = ctx.layout
        host = ctx.download(d)
        for n in lay.branch_names:
            br = host.branches[n]
            fr = br.coords[0, br.inds[0]]
            if fr.shape[0] < self.nfriends:
                raise ValueError(f"branch {n!r}: {fr.shape[0]} cold-chain leaves cannot provide {self.nfriends} friends")
            means, uni = np.unique(fr[:, lay.friend_key[n]].copy(), return_index=True)
            self.friends[n], self.means[n] = fr[uni], means
        self._friends = ctx.make_friends(self.friends, self.means)
        ctx.friends_update(d, self._friends[0], 0)

    def propose(self, model, state):
        from ..multibranch import MBContext, MBDeviceState
        ctx = self._context()
        if not isinstance(ctx, MBContext):
            raise RuntimeError("GroupStretchMove runs on a multi-branch sampler (EnsembleSampler with nleaves_max / several branches)")
        if ctx.layout.nfriends != self.nfriends:
            raise ValueError("the sampler's friend table width differs from this move's nfriends")
        host_state = None
        d = state
        if not isinstance(state, MBDeviceState):
            host_state = state
            tc = self.temperature_control
            d = ctx.upload(state, betas=None if tc is None else tc.betas_dev)
        T, W = d.shape[:2]
        lay = ctx.layout
        if self.iter == 0 or self.iter % self.n_iter_update == 0:    # group.py:148-149
            self.setup_friends(ctx, d)
        else:                                                        # group.py:156-157
            ctx.friends_update(d, self._friends[0], 1)
        cnt = self._count_buffer(ctx, T, W)
        if ctx.rng == "numpy-replay":
            flags = d.flags_host()
            pick = np.zeros((T, W, lay.ltot), dtype=np.int32)
            u_z = None
            for b, n in enumerate(lay.branch_names):
                f = flags[n]
                r = np.random.randint(self.nfriends, size=(int(f.sum()),))   # fixture find_friends: GLOBAL stream
                sub = np.zeros(f.shape, dtype=np.int32)
                sub[f] = r
                pick[:, :, lay.loff[n]:lay.loff[n] + lay.nleaves_max[n]] = sub
                if b == 0:
                    u_z = model.random.rand(T, W)                            # stretch.py:131
            u_acc = model.random.rand(T, W)                                  # group.py:254
            acc = ctx.group_stretch(d, self.a, self._friends[0], replay=(pick, u_z, u_acc), accepted_count=cnt)
        else:
            acc = ctx.group_stretch(d, self.a, self._friends[0], accepted_count=cnt)
        self.num_proposals += 1
        tc = self.temperature_control
        if tc is not None:
            d = tc.temper_comps(d)                                           # group.py:272-273
        else:
            ctx.advance_iter()
        # group.py:275-278 re-runs setup_friends on a deep copy of the pre-move state: same friends, and the table it
        # fills belongs to the copy — nothing to do here
        self.iter += 1
        if host_state is None:
            return d, acc
        return ctx.download(d, into=host_state), acc.cpu().numpy().astype(bool)
