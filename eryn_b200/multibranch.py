"""Several branches with a variable number of leaves on the device (BASELINE config 5: reversible jump + group stretch).

The reference keeps one coords / inds array per branch (state.py:330-384).  On the device the branches of a walker
are stored back to back in one row, and the leaf flags together with the group move's friend table form the walker's
byte payload (`aux`), so the swap pass moves everything that belongs to a walker in one go — include/eryn_b200.h,
section "Reversible jump + group stretch".  Host `State` objects keep the reference layout; `MBContext.upload` /
`download` convert."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import DeviceContext, DeviceState, _ptr
from .state import State

__all__ = ["PulseLikelihood", "MBLayout", "MBDeviceState", "MBContext"]


class PulseLikelihood(object):
    """log L = -1/2 sum(((template(t) - y)/sigma)^2) with template = sum over active leaves of
    'gauss': a exp(-(t-b)^2/(2c^2))  or  'sine': a sin(2 pi b t + c)   (the reference test's likelihood,
    tests/test_eryn.py:38-92).  kinds: {branch_name: 'gauss' | 'sine'}."""

    KINDS = {"gauss": _lib.EB_PULSE_GAUSS, "sine": _lib.EB_PULSE_SINE}

    def __init__(self, t, y, sigma, kinds):
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        if self.t.shape != self.y.shape or self.t.ndim != 1:
            raise ValueError("t and y must be 1-d arrays of the same length")
        self.sigma = float(sigma)
        for k in kinds.values():
            if k not in self.KINDS:
                raise ValueError(f"unknown pulse kind {k!r}")
        self.kinds = dict(kinds)


class MBLayout(object):
    def __init__(self, branch_names, ndims, nleaves_max, nleaves_min, kinds, nfriends=0, friend_key=1):
        self.branch_names = list(branch_names)
        if not 1 <= len(self.branch_names) <= _lib.EB_MAX_BRANCHES:
            raise ValueError(f"1..{_lib.EB_MAX_BRANCHES} branches")
        self.ndims = {n: int(ndims[n]) for n in self.branch_names}
        self.nleaves_max = {n: int(nleaves_max[n]) for n in self.branch_names}
        self.nleaves_min = {n: int(nleaves_min[n]) for n in self.branch_names}
        self.nfriends = int(nfriends)
        fk = friend_key if isinstance(friend_key, dict) else {n: friend_key for n in self.branch_names}
        self.friend_key = {n: int(fk[n]) for n in self.branch_names}
        c = _lib.eb_mb_layout()
        c.nbranches, c.nfriends = len(self.branch_names), self.nfriends
        self.coff, self.loff, self.poff = {}, {}, {}
        co = lo = po = 0
        for b, n in enumerate(self.branch_names):
            if self.nleaves_min[n] > self.nleaves_max[n]:
                raise ValueError("nleaves_min cannot be greater than nleaves_max.")
            c.nleaves[b], c.ndim[b], c.nleaves_min[b] = self.nleaves_max[n], self.ndims[n], self.nleaves_min[n]
            c.kind[b] = PulseLikelihood.KINDS[kinds[n]]
            c.friend_key[b] = self.friend_key[n]
            self.coff[n], self.loff[n], self.poff[n] = co, lo, po
            co += self.nleaves_max[n] * self.ndims[n]
            lo += self.nleaves_max[n]
            po += self.ndims[n]
        self.row, self.ltot = co, lo
        self.c = c
        self.flags_pad = (lo + 3) // 4 * 4
        self.aux_stride = self.flags_pad + 4 * lo * max(self.nfriends, 0)


class MBDeviceState(DeviceState):
    """coords [T,W,1,row] f64, logl/logp [T,W], inds = aux [T,W,aux_stride] u8 (flags, then the friend table)."""

    def __init__(self, layout, coords, logl, logp, aux, betas=None, temp_offset=0):
        super().__init__(coords, logl, logp, aux, betas, "+".join(layout.branch_names), temp_offset)
        self.layout = layout
        self.inds_stride = layout.aux_stride

    @property
    def aux(self):
        return self.inds

    def mb_struct(self):
        T, W = self.shape[:2]
        return _lib.eb_mb_state(T, W, self.temp_offset, 0, _ptr(self.coords), _ptr(self.logl), _ptr(self.logp),
                                _ptr(self.inds), _ptr(self.betas))

    def flags_host(self):
        """{branch: bool [T,W,L]} — the replay-mode host draws need them (distgenrj.py:58, fixture find_friends)"""
        lay = self.layout
        f = self.inds[:, :, :lay.ltot].cpu().numpy().astype(bool)
        return {n: f[:, :, lay.loff[n]:lay.loff[n] + lay.nleaves_max[n]] for n in lay.branch_names}


class MBContext(DeviceContext):
    """priors: {branch: ProbDistContainer}; like: PulseLikelihood; the rest as DeviceContext."""

    def __init__(self, priors, like, branch_names, ndims, nleaves_max, nleaves_min, nfriends=0, friend_key=1, device=None,
                 rng="philox", seed=0):
        self._init_common(device, rng, seed)
        if not isinstance(like, PulseLikelihood):
            raise NotImplementedError("several branches / variable leaf counts run with the built-in PulseLikelihood "
                                      "functor on the device (DESIGN.md §9)")
        self.like = like
        self.fused = True
        self.layout = MBLayout(branch_names, ndims, nleaves_max, nleaves_min, like.kinds, nfriends, friend_key)
        self.priors = {n: priors[n] for n in branch_names}
        lo, hi, lp = [np.concatenate([self.priors[n].arrays()[k] for n in branch_names]) for k in range(3)]
        self._prior_dev = torch.from_numpy(np.stack([lo, hi, lp])).to(self.device)
        self._prior_c = _lib.eb_prior(_ptr(self._prior_dev[0]), _ptr(self._prior_dev[1]), _ptr(self._prior_dev[2]), None)
        self.periods = None
        self._t_dev, self._y_dev = self.to_dev(like.t), self.to_dev(like.y)
        self._data_c = _lib.eb_pulse_data(len(like.t), 0, like.sigma, _ptr(self._t_dev), _ptr(self._y_dev))
        if self.lib.eb_mb_aux_stride(C.byref(self.layout.c)) != self.layout.aux_stride:
            raise _lib.ErynB200Error("aux stride mismatch between the host layout and the library")
        self.ndim = self.layout.row

    # ---- state movement -------------------------------------------------------------------------
    def upload(self, state, betas=None):
        lay = self.layout
        names = list(state.branches.keys())
        if names != lay.branch_names:
            raise ValueError(f"state branches {names} differ from the sampler's {lay.branch_names}")
        T, W = state.branches[names[0]].shape[:2]
        row = np.empty((T, W, 1, lay.row))
        aux = np.zeros((T, W, lay.aux_stride), dtype=np.uint8)
        for n in names:
            br = state.branches[n]
            if tuple(br.shape) != (T, W, lay.nleaves_max[n], lay.ndims[n]):
                raise ValueError("incompatible input dimensions")
            row[:, :, 0, lay.coff[n]:lay.coff[n] + lay.nleaves_max[n] * lay.ndims[n]] = br.coords.reshape(T, W, -1)
            aux[:, :, lay.loff[n]:lay.loff[n] + lay.nleaves_max[n]] = br.inds
        if lay.nfriends > 0:
            aux[:, :, lay.flags_pad:] = 0xFF  # friend table = -1 everywhere until the group move sets it up
        dev = self.device
        logl = self.to_dev(state.log_like, np.float64) if state.log_like is not None else torch.empty((T, W), dtype=torch.float64, device=dev)
        logp = self.to_dev(state.log_prior, np.float64) if state.log_prior is not None else torch.empty((T, W), dtype=torch.float64, device=dev)
        return MBDeviceState(lay, self.to_dev(row), logl, logp, self.to_dev(aux), betas)

    def download(self, d, into=None, random_state=None):
        lay = self.layout
        T, W = d.shape[:2]
        row = d.coords.cpu().numpy().reshape(T, W, lay.row)
        flags = d.inds[:, :, :lay.ltot].cpu().numpy().astype(bool)
        coords = {n: row[:, :, lay.coff[n]:lay.coff[n] + lay.nleaves_max[n] * lay.ndims[n]].reshape(
            T, W, lay.nleaves_max[n], lay.ndims[n]).copy() for n in lay.branch_names}
        inds = {n: flags[:, :, lay.loff[n]:lay.loff[n] + lay.nleaves_max[n]].copy() for n in lay.branch_names}
        logl, logp = d.logl.cpu().numpy(), d.logp.cpu().numpy()
        betas = None if d.betas is None else d.betas.cpu().numpy()
        if into is None:
            return State(coords, inds=inds, log_like=logl, log_prior=logp, betas=betas, random_state=random_state)
        for n in lay.branch_names:
            into.branches[n].coords[...] = coords[n]
            into.branches[n].inds[...] = inds[n]
        into.log_like, into.log_prior = logl, logp
        if betas is not None:
            into.betas = betas
        return into

    def friend_table_host(self, d):
        lay = self.layout
        T, W = d.shape[:2]
        tab = d.inds[:, :, lay.flags_pad:].contiguous().cpu().numpy().view(np.int32).reshape(T, W, lay.ltot, lay.nfriends)
        return {n: tab[:, :, lay.loff[n]:lay.loff[n] + lay.nleaves_max[n]] for n in lay.branch_names}

    # ---- kernels ------------------------------------------------------------------------------
    def eval_state(self, d):
        st = d.mb_struct()
        _lib.check(self.lib.eb_mb_eval_state(C.byref(self.layout.c), C.byref(st), C.byref(self._prior_c),
                                             C.byref(self._data_c), self.stream()), "eb_mb_eval_state")
        self.launches += 1

    def make_friends(self, coords_by_branch, keys_by_branch):
        """device copy of the stationary friends: per branch [nfr, D] coords sorted by their key, and the keys"""
        fr = _lib.eb_mb_friends()
        keep = []
        for b, n in enumerate(self.layout.branch_names):
            c, k = self.to_dev(coords_by_branch[n], np.float64), self.to_dev(keys_by_branch[n], np.float64)
            keep += [c, k]
            fr.nfr[b], fr.coords[b], fr.keys[b] = c.shape[0], c.data_ptr(), k.data_ptr()
        return fr, keep

    def friends_update(self, d, friends, mode):
        st = d.mb_struct()
        _lib.check(self.lib.eb_mb_friends_update(C.byref(self.layout.c), C.byref(st), C.byref(friends), int(mode),
                                                 self.stream()), "eb_mb_friends_update")
        self.launches += 1

    def group_stretch(self, d, a, friends, replay=None, accepted_count=None):
        T, W = d.shape[:2]
        st = d.mb_struct()
        r = _lib.eb_mb_group_rng()
        keep = None
        if replay is None:
            r.mode, r.seed, r.iter_dev = _lib.EB_RNG_PHILOX, self.seed, self.iter_ptr
        else:
            pick, u_z, u_acc = replay
            keep = (self.to_dev(pick, np.int32), self.to_dev(u_z, np.float64), self.to_dev(u_acc, np.float64))
            r.mode = _lib.EB_RNG_REPLAY
            r.pick, r.u_z, r.u_acc = [_ptr(x) for x in keep]
        acc = self.accepted_mask(T, W)
        _lib.check(self.lib.eb_mb_group_stretch(C.byref(self.layout.c), C.byref(st), C.byref(self._prior_c),
                                                C.byref(self._data_c), C.byref(friends), float(a), C.byref(r), _ptr(acc),
                                                _ptr(accepted_count), self.stream()), "eb_mb_group_stretch")
        self.launches += 1
        return acc

    def rj_step(self, d, replay=None, accepted_count=None, branch_mask=0, gibbs_index=0):
        """DistributionGenerateRJ step; branch_mask / gibbs_index: the Gibbs split over branches of this call (rj.py:168)"""
        T, W = d.shape[:2]
        st = d.mb_struct()
        r = _lib.eb_mb_rj_rng()
        r.branch_mask, r.gibbs_index = int(branch_mask), int(gibbs_index)
        keep = None
        if replay is None:
            r.mode, r.seed, r.iter_dev = _lib.EB_RNG_PHILOX, self.seed, self.iter_ptr
        else:
            change, leaf, births, u_acc = replay
            keep = [self.to_dev(change, np.int32), self.to_dev(leaf, np.int32), self.to_dev(u_acc, np.float64)]
            r.mode = _lib.EB_RNG_REPLAY
            r.change, r.leaf, r.u_acc = [_ptr(x) for x in keep]
            for b, bt in enumerate(births):
                if bt is not None:
                    t = self.to_dev(bt, np.float64)
                    keep.append(t)
                    r.birth[b] = t.data_ptr()
        acc = self.scratch("rj_accepted", (T, W), torch.uint8)
        _lib.check(self.lib.eb_mb_rj_step(C.byref(self.layout.c), C.byref(st), C.byref(self._prior_c), C.byref(self._data_c),
                                          C.byref(r), _ptr(acc), _ptr(accepted_count), self.stream()), "eb_mb_rj_step")
        self.launches += 1
        return acc
