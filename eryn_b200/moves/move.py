"""Move base class: the attributes EnsembleSampler touches on a move (move.py:404-470)."""
import numpy as np

from ..device import DeviceContext, DeviceState

__all__ = ["Move"]


class Move(object):
    def __init__(self, temperature_control=None, periodic=None, ctx=None, gibbs_sampling_setup=None, **kwargs):
        if kwargs:
            raise NotImplementedError(f"unsupported Move kwargs on the device path: {sorted(kwargs)}")
        self._initialize_branch_setup(gibbs_sampling_setup)
        self.periodic = periodic  # applied inside the kernels through the DeviceContext (utils/periodic.py)
        self._accepted = None
        self._accepted_dev = None
        self.num_proposals = 0
        self.temperature_control = temperature_control
        self.ctx = ctx

    # ---- Gibbs sampling setup (move.py:113-240) ------------------------------------------------------------------
    def _initialize_branch_setup(self, gibbs_sampling_setup):
        """Same input forms and checks as the reference: None | branch name | (branch, mask or None) | {branch: mask or
        None} | list of those.  A mask is a 2-d boolean array [nleaves_max, ndim].  Stored as `gibbs_splits`: a list of
        {branch: mask | None} dicts, one per split, or [None] (no Gibbs sampling: all branches, all parameters)."""
        self.gibbs_sampling_setup = gibbs_sampling_setup
        if gibbs_sampling_setup is None:
            self.gibbs_splits = [None]
            return
        if type(gibbs_sampling_setup) not in [str, tuple, list, dict]:
            raise ValueError("gibbs_sampling_setup must be string, dict, tuple, or list.")
        items = gibbs_sampling_setup if isinstance(gibbs_sampling_setup, list) else [gibbs_sampling_setup]
        msg = ("When inputing gibbs indexing and using a 2-tuple, second item must be None or 2D np.ndarray of shape "
               "(nleaves_max, ndim).")

        def check(v):
            if (not isinstance(v, np.ndarray) and v is not None) or (isinstance(v, np.ndarray) and v.ndim != 2):
                raise ValueError(msg)
            return None if v is None else np.asarray(v, dtype=bool)
        splits = []
        for item in items:
            if isinstance(item, str):
                splits.append({item: None})
            elif isinstance(item, tuple):
                assert len(item) == 2
                splits.append({item[0]: check(item[1])})
            elif isinstance(item, dict):
                splits.append({k: check(v) for k, v in item.items()})
            else:
                raise ValueError("If providing a list for gibbs_sampling_setup, each item needs to be a string, tuple, or dict.")
        self.gibbs_splits = splits

    def _single_branch_splits(self, branch_name, L, D):
        """the Gibbs splits of a single-branch, single-leaf device state as (index, mask bits, number of selected
        parameters); mask bits 0 = every parameter.  Splits without a selected parameter are skipped, as the reference
        skips them (`at_least_one_proposal`, move.py:283-300)."""
        out = []
        for gi, split in enumerate(self.gibbs_splits):
            if split is None:
                out.append((gi, 0, L * D))
                continue
            if list(split.keys()) != [branch_name]:
                raise KeyError(f"gibbs_sampling_setup names branches {sorted(split)}; this sampler has one branch, {branch_name!r}")
            m = split[branch_name]
            if m is None:
                out.append((gi, 0, L * D))
                continue
            if m.shape != (L, D):
                raise ValueError(f"Gibbs mask must have shape (nleaves_max, ndim) = ({L}, {D})")
            if L != 1:
                raise NotImplementedError("parameter-level Gibbs splits run in the single-leaf kernels (DESIGN.md)")
            if not m.any():
                continue
            bits = 0
            for j in np.nonzero(m[0])[0]:
                bits |= 1 << int(j)
            out.append((gi, bits, int(m.sum())))
        return out

    # ---- accepted counters (move.py:404-421); the device accumulator is merged lazily --------------
    @property
    def accepted(self):
        if self._accepted is None:
            raise ValueError("accepted must be inititalized with the init_accepted function if you want to use it.")
        if self._accepted_dev is not None:
            return self._accepted + self._accepted_dev.cpu().numpy().astype(np.float64)
        return self._accepted

    @accepted.setter
    def accepted(self, accepted):
        assert isinstance(accepted, np.ndarray)
        self._accepted = accepted
        if self._accepted_dev is not None:
            self._accepted_dev.zero_()

    @property
    def acceptance_fraction(self):
        return self.accepted / self.num_proposals

    @property
    def temperature_control(self):
        return self._temperature_control

    @temperature_control.setter
    def temperature_control(self, temperature_control):
        self._temperature_control = temperature_control
        if temperature_control is not None:
            self.ntemps = temperature_control.ntemps

    def tune(self, state, accepted):
        pass

    # ---- CUDA-graph replay of the sampler loop (ensemble.py: EnsembleSampler._advance_resident) ------------------
    graphable = False  # True: propose() on a DeviceState in philox mode only launches kernels (no host draws / syncs)

    def _host_tick(self, n=1):
        """the host-side bookkeeping of n proposals whose launches are replayed from a captured graph"""
        self.num_proposals += n

    # ---- helpers ----------------------------------------------------------------------------------
    def bind(self, ctx):
        self.ctx = ctx
        if self.temperature_control is not None and self.temperature_control.ctx is None:
            self.temperature_control.bind(ctx)

    def _context(self):
        if not isinstance(self.ctx, DeviceContext):
            raise RuntimeError("move is not bound to a DeviceContext (EnsembleSampler does this; standalone use: "
                               "pass ctx=DeviceContext(priors, log_like_fn))")
        return self.ctx

    def _count_buffer(self, ctx, T, W):
        import torch
        if self._accepted_dev is None or tuple(self._accepted_dev.shape) != (T, W):
            self._accepted_dev = torch.zeros((T, W), dtype=torch.int32, device=ctx.device)
        if self._accepted is None:
            self._accepted = np.zeros((T, W))
        return self._accepted_dev

    def _enter(self, state):
        """host State -> DeviceState (drop-in use with NumPy states); DeviceState passes through."""
        ctx = self._context()
        if isinstance(state, DeviceState):
            return ctx, state, None
        tc = self.temperature_control
        betas = None
        if tc is not None:
            if state.betas is not None:
                tc.betas = state.betas
            betas = tc.betas_dev
        return ctx, ctx.upload(state, betas=betas), state

    def _exit(self, ctx, d, host_state, acc):
        tc = self.temperature_control
        if tc is not None:
            d = tc.temper_comps(d)
        else:
            ctx.advance_iter()
        if host_state is None:
            return d, acc
        accepted = acc.cpu().numpy().astype(bool)
        state = ctx.download(d, into=host_state)
        return state, accepted
