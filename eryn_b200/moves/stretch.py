"""StretchMove on the device (reference: moves/red_blue.py:89-333 + moves/stretch.py)."""
import numpy as np

from .move import Move

__all__ = ["StretchMove"]


class StretchMove(Move):
    """Affine-invariant stretch move, red/blue parallelisation (Goodman & Weare 2010).

    Same constructor surface as eryn.moves.StretchMove for what the device path covers:
    a, nsplits (must be 2), randomize_split, live_dangerously, temperature_control."""

    def __init__(self, a=2.0, nsplits=2, randomize_split=True, live_dangerously=False, **kwargs):
        if int(nsplits) != 2:
            raise NotImplementedError("the device stretch kernel implements nsplits == 2 (the reference default)")
        self.a = a
        self.nsplits = 2
        self.randomize_split = randomize_split
        self.live_dangerously = live_dangerously
        super().__init__(**kwargs)

    graphable = True

    def propose(self, model, state):
        """(state, accepted) — ensemble.py:974.  `state` may be a host State or a DeviceState."""
        ctx, d, host_state = self._enter(state)
        T, W, L, D = d.shape
        if W < 2 * L * D and not self.live_dangerously:  # red_blue.py:103-114
            raise RuntimeError(
                "It is unadvisable to use a red-blue move with fewer walkers than twice the number of "
                "dimensions. If you would like to do this, please set live_dangerously to True.")
        cnt = self._count_buffer(ctx, T, W)
        step = ctx.stretch_step if ctx.fused else ctx.stretch_step_split
        splits = self._single_branch_splits(d.branch_name, L, D)
        if len(splits) != 1 or splits[0][1]:
            if not ctx.fused:
                raise NotImplementedError("Gibbs splits run in the fused kernels (DeviceLikelihood functor)")
        lists = None
        if ctx.rng == "numpy-replay":
            # same draws, same order as the reference: global shuffle once per propose (red_blue.py:124), then per Gibbs
            # split and red/blue half private randint / rand / rand (stretch.py:93, :131, red_blue.py:294)
            ids = np.tile(np.arange(W), (T, 1))
            labels = ids % self.nsplits
            if self.randomize_split:
                [np.random.shuffle(x) for x in labels]
            lists = [ids[labels == s].reshape(T, -1) for s in range(2)]
        self._nsplits_run = len(splits)
        acc = None
        for gi, bits, gnd in splits:
            gibbs = None if (bits == 0 and gi == 0) else (bits, gnd, gi)
            kw = {} if gibbs is None else dict(gibbs=gibbs)
            if lists is not None:
                rint, u_z, u_acc = [], [], []
                for split in range(2):
                    Ns, Nc = lists[split].shape[1], lists[1 - split].shape[1]
                    rint.append(model.random.randint(Nc, size=(T, Ns)))
                    u_z.append(model.random.rand(T, Ns))
                    u_acc.append(model.random.rand(T, Ns))
                acc = step(d, self.a, replay=dict(lists=lists, rint=rint, u_z=u_z, u_acc=u_acc), accepted_count=cnt, **kw)
            else:
                acc = step(d, self.a, randomize_split=self.randomize_split, accepted_count=cnt, **kw)
            self.num_proposals += 1  # red_blue.py:326: per Gibbs split
        if acc is None:  # every split was empty: nothing proposed (red_blue.py:142-143)
            acc = ctx.accepted_mask(T, W)
            acc.zero_()
        return self._exit(ctx, d, host_state, acc)

    def _host_tick(self, n=1):
        self.num_proposals += n * getattr(self, "_nsplits_run", 1)
