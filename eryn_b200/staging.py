"""Staging of stored samples: Backend.save_step without stalling the sampler.

Reference: `Backend.save_step` (backends/backend.py:1014-1091) is called from the sampler loop at every
`thin_by`-th iteration (ensemble.py:1013-1028) with the host State; it NaN-fills the coordinates of inactive
leaves (backend.py:1053-1059) and copies everything into the chain arrays.

Here the walkers live on the device.  A stored step is
  1. one pack kernel on the SAMPLER's stream (`eb_stage_pack`, csrc/k_stage.cu): coords (NaN mask applied), logl,
     logp, betas, leaf flags, the accept mask, the per-move accept counters and the control block (swap counts,
     adaptation clock) gathered into one contiguous device slot — a snapshot, after which the sampler runs on;
  2. one device-to-host copy of the slot into pinned memory on a SIDE stream (ordered after the pack by an event);
  3. the host arrays are cut out of the pinned slot lazily: `LazyState` materialises on first attribute access,
     the backend drains finished slots in order whenever it is asked for something.
The ring has `nslots` slots; pushing into a slot that is still pending drains it first (the only point where the
host ever waits, and only for a copy that was issued `nslots` yields earlier).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .state import State

__all__ = ["StoreRing", "LazyState"]


def _pad8(n):
    return (int(n) + 7) & ~7


class _Slot(object):
    __slots__ = ("dev", "host", "packed", "copied", "busy", "copying", "meta")

    def __init__(self, nbytes, device):
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.packed = torch.cuda.Event()
        self.copied = torch.cuda.Event()
        self.busy = False
        self.copying = False
        self.meta = None


class StoreRing(object):
    """Ring of device + pinned-host staging slots for the samples of one DeviceState."""

    def __init__(self, ctx, d, count_buffers=(), nslots=4, fill=np.nan):
        self.ctx, self.lib = ctx, ctx.lib
        self.device = d.device
        self.d = d
        self.fill = float(fill)
        T, W, L, D = d.shape
        self.shape = (T, W, L, D)
        self.count_buffers = list(count_buffers)
        # (name, tensor, numpy dtype, shape)
        segs = [("coords", d.coords, np.float64, (T, W, L, D)), ("logl", d.logl, np.float64, (T, W)),
                ("logp", d.logp, np.float64, (T, W)), ("ctrl", ctx.ctrl, np.uint8, (ctx.ctrl.numel(),)),
                ("accepted", ctx.accepted_mask(T, W), np.uint8, (T, W))]
        if d.betas is not None:
            segs.append(("betas", d.betas, np.float64, (T,)))
        if d.inds is not None:
            # segment 0 is stored with the NaN mask applied (what the backend keeps); the yielded State needs the raw values
            segs.append(("inds", d.inds, np.uint8, (T, W, L)))
            segs.append(("coords_raw", d.coords, np.float64, (T, W, L, D)))
        for i, cb in enumerate(self.count_buffers):
            segs.append((f"count{i}", cb, np.int32, (T, W)))
        if len(segs) > _lib.EB_STAGE_MAX_SEGMENTS:
            raise ValueError("too many per-move counters for one staging slot")
        self.segs = segs
        self.offsets = {}
        o = 0
        for name, t, dt, shp in segs:
            nb = t.numel() * t.element_size()
            self.offsets[name] = (o, nb, dt, shp)
            o += _pad8(nb)
        self.nbytes = o
        self.slots = [_Slot(o, self.device) for _ in range(int(nslots))]
        self.side = torch.cuda.Stream(device=self.device)
        self.next = 0
        sg = _lib.eb_stage()
        sg.nseg = len(segs)
        for i, (name, t, dt, shp) in enumerate(segs):
            sg.src[i] = t.data_ptr()
            sg.nbytes[i] = t.numel() * t.element_size()
        if d.inds is not None:
            sg.mask_inds = d.inds.data_ptr()
            sg.mask_nleaves, sg.mask_ndim = L, D
        sg.fill = self.fill
        sg.dst_bytes = o
        self._sg = sg
        self.on_evict = None  # called with a slot's ticket before the slot is reused (the sampler drains its store queue)

    def push(self, meta=None, copy=True):
        """snapshot the state now (in stream order); returns a ticket.  `copy=False` takes the device snapshot only: the
        copy to the host is issued when (if) somebody reads the sample.  Never synchronises unless the ring is full."""
        idx = self.next
        self.next = (idx + 1) % len(self.slots)
        s = self.slots[idx]
        if s.busy:
            if self.on_evict is not None:
                self.on_evict(s.meta)   # the sampler stores every pending sample up to this one
            s.meta.evict()              # a yielded state that is still referenced keeps its data
            if s.copying:
                s.copied.synchronize()  # the side stream may still be reading s.dev
        self.ctx.flush_adapt()   # the snapshot holds the ladder and the swap counts: a deferred adaptation goes first
        cur = torch.cuda.current_stream(self.device)
        self._sg.dst = s.dev.data_ptr()
        _lib.check(self.lib.eb_stage_pack(C.byref(self._sg), C.c_void_p(cur.cuda_stream)), "eb_stage_pack")
        self.ctx.launches += 1
        s.packed.record(cur)
        s.busy, s.copying = True, False
        ticket = _Ticket(self, idx, meta)
        s.meta = ticket
        if copy:
            self.start_copy(idx)
        return ticket

    def start_copy(self, idx):
        s = self.slots[idx]
        if s.copying:
            return
        self.side.wait_event(s.packed)
        with torch.cuda.stream(self.side):
            s.host.copy_(s.dev, non_blocking=True)
            s.copied.record(self.side)
        s.copying = True

    def arrays(self, idx):
        """host views into the pinned slot (valid until the slot is reused); waits for the copy"""
        s = self.slots[idx]
        self.start_copy(idx)
        s.copied.synchronize()
        raw = s.host.numpy()
        out = {}
        for name, (o, nb, dt, shp) in self.offsets.items():
            out[name] = raw[o:o + nb].view(dt).reshape(shp)
        return out


class _Ticket(object):
    """one pushed sample.  `views()` = arrays in the pinned slot (no copy; valid until the slot is reused), `get()` =
    private host copies (taken once; what a materialised LazyState keeps)."""

    def __init__(self, ring, idx, meta):
        self.ring, self.idx, self.meta = ring, idx, meta
        self._data = None
        self.state_ref = None   # weakref to the LazyState that was yielded for this sample
        self.stored = False

    def _in_slot(self):
        return self.ring.slots[self.idx].meta is self

    def ready(self):
        s = self.ring.slots[self.idx]
        return self._data is not None or (s.copying and s.copied.query())

    def views(self):
        if self._data is not None:
            return self._data
        if not self._in_slot():
            raise RuntimeError("staged sample was evicted before it was read")
        return self.ring.arrays(self.idx)

    def get(self):
        if self._data is None:
            self._data = {k: v.copy() for k, v in self.views().items()}
        return self._data

    def release(self):
        """the backend has stored this sample"""
        self.stored = True

    def evict(self):
        """the slot is about to be reused: a yielded state that is still referenced keeps its data"""
        ref = self.state_ref() if self.state_ref is not None else None
        if ref is not None and self._data is None:
            self.get()


class LazyState(State):
    """The State yielded by the device sampler: its arrays are cut out of the pinned staging slot on first access.
    Untouched states cost the host nothing."""

    _LAZY = ("branches", "log_like", "log_prior", "betas", "blobs", "supplemental")

    def __init__(self, ticket, branch_name, random_state=None):  # noqa: deliberately not calling State.__init__
        object.__setattr__(self, "_ticket", ticket)
        object.__setattr__(self, "_branch_name", branch_name)
        object.__setattr__(self, "_pristine", None)
        self.random_state = random_state

    @property
    def materialized(self):
        return "branches" in self.__dict__

    def _materialize(self):
        a = self._ticket.get()
        inds = a.get("inds")
        coords = a["coords_raw" if inds is not None else "coords"].copy()
        st = State({self._branch_name: coords}, inds=None if inds is None else {self._branch_name: inds.astype(bool)},
                   log_like=a["logl"].copy(), log_prior=a["logp"].copy(),
                   betas=None if "betas" not in a else a["betas"].copy())
        d = self.__dict__
        d["branches"], d["log_like"], d["log_prior"], d["betas"] = st.branches, st.log_like, st.log_prior, st.betas
        d["blobs"], d["supplemental"] = None, None
        object.__setattr__(self, "_pristine", a)

    def __getattr__(self, name):  # only reached when normal lookup fails
        if name in LazyState._LAZY:
            self._materialize()
            return self.__dict__[name]
        raise AttributeError(name)

    def modified(self):
        """which arrays a consumer changed after the state was yielded (compared with the staged copy)"""
        if not self.materialized:
            return {}
        a = self._pristine
        out = {}
        br = self.branches[self._branch_name]
        if not np.array_equal(br.coords, a["coords_raw" if "inds" in a else "coords"], equal_nan=True):
            out["coords"] = br.coords
        if not np.array_equal(self.log_like, a["logl"]):
            out["logl"] = self.log_like
        if not np.array_equal(self.log_prior, a["logp"]):
            out["logp"] = self.log_prior
        if "inds" in a and not np.array_equal(br.inds, a["inds"].astype(bool)):
            out["inds"] = br.inds
        if "betas" in a and self.betas is not None and not np.array_equal(self.betas, a["betas"]):
            out["betas"] = self.betas
        return out
