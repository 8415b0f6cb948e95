"""Device-resident walker state and the thin launch layer over the C ABI.

PyTorch is used for device memory, streams and (in `dist.py`) `torch.distributed`; every kernel
is launched through `liberyn_b200.so` on torch's current stream with raw device pointers."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .likelihood import DeviceLikelihood, TorchLikelihood
from .state import State

__all__ = ["DeviceState", "DeviceContext"]

_CTRL_ITER_OFF = 0


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class DeviceState(object):
    """Walker state of one branch as CUDA tensors, laid out like the reference's arrays:
    coords [T,W,L,D] f64, logl/logp [T,W] f64, inds [T,W,L] u8 (or None), betas [T] f64 (or None)."""

    def __init__(self, coords, logl, logp, inds=None, betas=None, branch_name="model_0", temp_offset=0):
        assert coords.is_cuda and coords.dtype == torch.float64 and coords.is_contiguous() and coords.dim() == 4
        self.coords, self.logl, self.logp, self.inds, self.betas = coords, logl, logp, inds, betas
        self.branch_name = branch_name
        self.temp_offset = int(temp_offset)  # temperature-sharded runs: global index of local temperature 0
        self.shape = tuple(coords.shape)
        self._c = None

    @property
    def device(self):
        return self.coords.device

    def c_struct(self):
        T, W, L, D = self.shape
        st = _lib.eb_state(T, W, L, D, self.temp_offset, getattr(self, "inds_stride", 0), _ptr(self.coords), _ptr(self.logl), _ptr(self.logp),
                           _ptr(self.inds), _ptr(self.betas))
        return st

    def clone(self):
        return DeviceState(self.coords.clone(), self.logl.clone(), self.logp.clone(),
                           None if self.inds is None else self.inds.clone(),
                           None if self.betas is None else self.betas.clone(), self.branch_name, self.temp_offset)


class DeviceContext(object):
    """Everything a move needs on the device: priors, likelihood functor, RNG mode and the control block.

    rng = "philox"       production: all randoms generated in-kernel (seed, epoch counter, tags)
    rng = "numpy-replay" parity: the host draws NumPy randoms in the reference's order and the kernels
                          consume them (bit-identical accept masks to the reference)."""

    def __init__(self, priors, log_like_fn, device=None, rng="philox", seed=0, branch_name="model_0", periodic=None):
        self._init_common(device, rng, seed)
        self.branch_name = branch_name
        if isinstance(priors, dict):
            priors = priors[branch_name]
        self.priors = priors
        lo, hi, lp = priors.arrays()
        self.ndim = len(lo)
        # periodic parameters: {index: period} (utils/periodic.py:16-47), a dense [ndim] array on the device (0 = none)
        per = np.zeros(self.ndim)
        if periodic:
            for idx, period in periodic.items():
                if not isinstance(idx, (int, np.integer)):
                    raise ValueError("If providing str values for the variable names, must provide key_order argument.")
                per[int(idx)] = float(period)
        self.periods = per if per.any() else None
        self._prior_dev = torch.from_numpy(np.stack([lo, hi, lp, per])).to(self.device)
        self._prior_c = _lib.eb_prior(_ptr(self._prior_dev[0]), _ptr(self._prior_dev[1]), _ptr(self._prior_dev[2]),
                                      _ptr(self._prior_dev[3]) if self.periods is not None else None)
        self.like = log_like_fn
        if isinstance(log_like_fn, DeviceLikelihood):
            p = np.ascontiguousarray(log_like_fn.params(), dtype=np.float64)
            self._like_dev = torch.from_numpy(p if p.size else np.zeros(1)).to(self.device)
            self._like_c = _lib.eb_like(int(log_like_fn.kind), int(log_like_fn.ncomp), int(p.size), 0,
                                        _ptr(self._like_dev))
            self.fused = True
        elif callable(log_like_fn):
            if not isinstance(log_like_fn, TorchLikelihood):
                self.like = TorchLikelihood(log_like_fn)
            self._like_c = None
            self.fused = False
        else:
            raise ValueError("log_like_fn must be a DeviceLikelihood or a callable on CUDA tensors")

    def _init_common(self, device, rng, seed):
        self.lib = _lib.require_device()
        if rng not in ("philox", "numpy-replay"):
            raise ValueError("rng must be 'philox' or 'numpy-replay'")
        self.rng = rng
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctrl = torch.zeros(self.lib.eb_ctrl_size(), dtype=torch.uint8, device=self.device)
        self._scratch = {}
        self.launches = 0  # kernels launched through this context (bench.py reports it)
        # Lazy ladder adaptation (include/eryn_b200.h: eb_swap_rng.defer_adapt): while it is on, a swap pass ends with its
        # row moves and the NEXT stretch kernel folds the counts and adapts the ladder in its prologue — 3 us less per
        # iteration at config 2.  Everything else that looks at betas / swap counts / the clock goes through
        # flush_adapt() first (a one-CTA kernel, no-op when nothing is pending).  Only loops that consist of plain
        # StretchMove proposals switch it on (EnsembleSampler._sample_resident, bench.py).
        self.lazy_adapt = False
        self._lazy_betas = None

    # ---- helpers ------------------------------------------------------------------------------
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def scratch(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def accepted_mask(self, T, W):
        return self.scratch("accepted", (T, W), torch.uint8)

    @property
    def iter_ptr(self):
        return C.c_void_p(self.ctrl.data_ptr() + _CTRL_ITER_OFF)

    @property
    def iter_next_ptr(self):
        """eb_ctrl.iter_next: equals iter between iterations; the swap pass publishes iter+1 there before it releases
        its programmatic dependents, so the next stretch kernel can start its draws early (pdl_chain)."""
        return C.c_void_p(self.ctrl.data_ptr() + _lib.eb_ctrl.iter_next.offset)

    def lazy_begin(self, betas):
        """switch lazy ladder adaptation on for a loop of stretch / Gaussian proposals on the ladder `betas` (device tensor)"""
        self.lazy_adapt = True
        self._lazy_betas = betas

    def lazy_end(self):
        """apply what is pending and go back to passes that adapt in their own kernel"""
        self.flush_adapt()
        self.lazy_adapt = False
        self._lazy_betas = None

    def flush_adapt(self):
        """apply a deferred ladder adaptation now (no-op on the device if nothing is pending).  The ladder tensor is
        remembered from lazy_begin / the first deferred pass: replayed graphs defer without passing through pt_swap()."""
        if self._lazy_betas is not None:
            _lib.check(self.lib.eb_adapt_flush(_ptr(self.ctrl), _ptr(self._lazy_betas), self.stream()), "eb_adapt_flush")
            self.launches += 1

    def read_ctrl(self):
        self.flush_adapt()
        return _lib.eb_ctrl.from_buffer_copy(self.ctrl.cpu().numpy().tobytes())

    def write_ctrl(self, iter=None, time=None):
        c = self.read_ctrl()   # (flushes a deferred adaptation first)
        if iter is not None:
            c.iter = int(iter)
            c.iter_next = int(iter)
        if time is not None:
            c.time = int(time)
        self.ctrl.copy_(torch.frombuffer(bytearray(bytes(c)), dtype=torch.uint8))

    def to_dev(self, arr, dtype=None):
        a = np.ascontiguousarray(arr, dtype=dtype)
        return torch.from_numpy(a).to(self.device)

    # ---- state movement -------------------------------------------------------------------------
    def upload(self, state, betas=None):
        """Host State -> DeviceState (single branch)."""
        if len(state.branches) != 1:
            raise NotImplementedError("the device hot path covers one branch per sampler in this build")
        name = list(state.branches.keys())[0]
        br = state.branches[name]
        coords = self.to_dev(br.coords, np.float64)
        T, W, L, D = coords.shape
        inds = None if bool(np.all(br.inds)) else self.to_dev(br.inds.astype(np.uint8))
        logl = self.to_dev(state.log_like, np.float64) if state.log_like is not None else \
            torch.empty((T, W), dtype=torch.float64, device=self.device)
        logp = self.to_dev(state.log_prior, np.float64) if state.log_prior is not None else \
            torch.empty((T, W), dtype=torch.float64, device=self.device)
        return DeviceState(coords, logl, logp, inds, betas, name)

    def check_error(self):
        """raise if a kernel set eb_ctrl.error (a bounded in-kernel wait ran out): the chain is invalid from there on"""
        err = int(_lib.eb_ctrl.from_buffer_copy(self.ctrl.cpu().numpy().tobytes()).error)
        if err:
            raise _lib.ErynB200Error(f"device error {err} (EB_DEVERR_*): a bounded wait inside the swap pass timed out "
                                     "(peer never published / a CTA never arrived); results after that pass are invalid")

    def download(self, d, into=None, random_state=None):
        """DeviceState -> host State (new, or refreshing the arrays of `into` in place)."""
        self.flush_adapt()
        self.check_error()
        coords = d.coords.cpu().numpy()
        logl = d.logl.cpu().numpy()
        logp = d.logp.cpu().numpy()
        betas = None if d.betas is None else d.betas.cpu().numpy()
        inds = None if d.inds is None else d.inds.cpu().numpy().astype(bool)
        if into is None:
            st = State({d.branch_name: coords}, inds=None if inds is None else {d.branch_name: inds},
                       log_like=logl, log_prior=logp, betas=betas, random_state=random_state)
            return st
        br = into.branches[d.branch_name]
        br.coords[...] = coords
        if inds is not None:
            br.inds[...] = inds
        if into.log_like is None:
            into.log_like = logl
        else:
            into.log_like[...] = logl
        if into.log_prior is None:
            into.log_prior = logp
        else:
            into.log_prior[...] = logp
        if betas is not None:
            into.betas = betas
        return into

    # ---- kernels ------------------------------------------------------------------------------
    def _require_fused(self):
        if not self.fused:
            raise _lib.ErynB200Error("fused kernels need a DeviceLikelihood functor")

    def eval_state(self, d):
        """compute_log_prior + compute_log_like of the whole state (ensemble.py:1127, :1219)."""
        if self.fused:
            st = d.c_struct()
            _lib.check(self.lib.eb_eval_state(C.byref(st), C.byref(self._prior_c), C.byref(self._like_c),
                                              self.stream()), "eb_eval_state")
            self.launches += 1
        else:
            T, W, L, D = d.shape
            lp = self.box_log_prior(d.coords.view(T * W, L, D), None if d.inds is None else d.inds.view(T * W, L))
            d.logp.copy_(lp.view(T, W))
            d.logl.copy_(self.user_log_like(d.coords.view(T * W, L, D), lp).view(T, W))

    def box_log_prior(self, q, inds=None):
        n, L, D = q.shape
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.eb_box_log_prior(_ptr(q), _ptr(inds), n, L, D, C.byref(self._prior_c), _ptr(out),
                                             self.stream()), "eb_box_log_prior")
        self.launches += 1
        return out

    def user_log_like(self, q, logp):
        """ensemble.py:1219-1545 for a callable on device tensors: skip logp = -inf, fill -1e300."""
        ll = torch.full_like(logp, -1e300)
        ok = ~torch.isinf(logp)
        if bool(ok.any()):
            ll[ok] = self.like(q[ok]).to(torch.float64)
        return ll

    def _stretch_rng(self, randomize_split, replay):
        """replay = dict(lists [2][T,Ns_s], rint [2], u_z [2], u_acc [2]) of host NumPy draws."""
        r = _lib.eb_stretch_rng()
        r.randomize_split = int(bool(randomize_split))
        keep = None
        if replay is None:
            r.mode = _lib.EB_RNG_PHILOX
            r.seed = self.seed
            r.iter_dev = self.iter_next_ptr
            r.pdl_chain = 1
            if self.lazy_adapt:
                r.lazy_ctrl = _ptr(self.ctrl)
        else:
            keep = dict(list=[self.to_dev(x, np.int32) for x in replay["lists"]],
                        rint=[self.to_dev(x, np.int64) for x in replay["rint"]],
                        u_z=[self.to_dev(x, np.float64) for x in replay["u_z"]],
                        u_acc=[None if x is None else self.to_dev(x, np.float64) for x in replay["u_acc"]])
            r.mode = _lib.EB_RNG_REPLAY
            for s in range(2):
                r.list[s], r.rint[s] = _ptr(keep["list"][s]), _ptr(keep["rint"][s])
                r.u_z[s], r.u_acc[s] = _ptr(keep["u_z"][s]), _ptr(keep["u_acc"][s])
        return r, keep

    def stretch_step(self, d, a, randomize_split=True, replay=None, accepted_count=None, gibbs=None):
        """StretchMove step, both halves fused (red_blue.py:148-323 + stretch.py:74-231 + move.py:472).
        gibbs = (parameter mask bits, number of selected parameters, index of the split in this propose call)."""
        self._require_fused()
        T, W, L, D = d.shape
        st = d.c_struct()
        r, keep = self._stretch_rng(randomize_split, replay)
        if gibbs is not None:
            r.gibbs_mask, r.gibbs_ndim, r.gibbs_index = int(gibbs[0]), int(gibbs[1]), int(gibbs[2])
            r.pdl_chain = 0 if gibbs[2] > 0 else r.pdl_chain  # only the first split follows the swap pass
            if gibbs[2] > 0:
                r.lazy_ctrl = None
        acc = self.accepted_mask(T, W)
        _lib.check(self.lib.eb_stretch_step(C.byref(st), C.byref(self._prior_c), C.byref(self._like_c),
                                            float(a), C.byref(r), _ptr(acc), _ptr(accepted_count),
                                            self.stream()), "eb_stretch_step")
        self.launches += 1 if (W + 1) // 2 <= 128 else 2  # one launch per red/blue half (small ensembles: one CTA per temperature)
        return acc

    def stretch_step_split(self, d, a, randomize_split=True, replay=None, accepted_count=None):
        """Split path for callables, per half: propose kernel -> prior kernel -> user likelihood -> accept kernel."""
        self.flush_adapt()
        T, W, L, D = d.shape
        st = d.c_struct()
        r, keep = self._stretch_rng(randomize_split, replay)
        acc = self.accepted_mask(T, W)
        for split in range(2):
            Ns = (W + 1) // 2 if split == 0 else W // 2
            q = torch.empty((T * Ns, L, D), dtype=torch.float64, device=self.device)
            factors = torch.empty(T * Ns, dtype=torch.float64, device=self.device)
            sub = torch.empty(T * Ns, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.eb_stretch_propose(C.byref(st), float(a), int(split), C.byref(r),
                                                   _ptr(self._prior_dev[3]) if self.periods is not None else None,
                                                   _ptr(q), _ptr(factors), _ptr(sub), self.stream()),
                       "eb_stretch_propose")
            inds_sub = None
            if d.inds is not None:
                inds_sub = torch.gather(d.inds, 1, sub.view(T, Ns, 1).long().expand(T, Ns, L)).reshape(T * Ns, L).contiguous()
            lp = self.box_log_prior(q, inds_sub)
            ll = self.user_log_like(q, lp)
            _lib.check(self.lib.eb_accept_update(C.byref(st), _ptr(sub), Ns, _ptr(q), _ptr(factors), _ptr(ll), _ptr(lp),
                                                 int(split), C.byref(r), _ptr(acc), _ptr(accepted_count),
                                                 self.stream()), "eb_accept_update")
            self.launches += 3
        return acc

    def gaussian_step(self, d, proposal, replay=None, accepted_count=None, gibbs=None, dim_mode=0, log_factor=0.0):
        """GaussianMove step, fused (mh.py:56-193 + gaussian.py:68-195).  gibbs = (parameter mask bits, split index);
        dim_mode 1 = one random dimension per walker, log_factor = log of GaussianMove's `factor` (philox mode)."""
        self._require_fused()
        lazy = self.lazy_adapt and replay is None and (gibbs is None or gibbs[1] == 0)
        if not lazy:
            self.flush_adapt()
        T, W, L, D = d.shape
        st = d.c_struct()
        r = _lib.eb_gauss_rng()
        if lazy:
            r.lazy_ctrl = _ptr(self.ctrl)   # the kernel applies a deferred ladder adaptation in its prologue
        r.dim_mode, r.log_factor = int(dim_mode), float(log_factor)
        if gibbs is not None:
            r.gibbs_mask, r.gibbs_index = int(gibbs[0]), int(gibbs[1])
        keep = None
        if replay is None:
            r.mode = _lib.EB_RNG_PHILOX
            r.seed = self.seed
            r.iter_dev = self.iter_ptr
            if proposal["kind"] == "scalar":
                r.cov_kind, r.scale = 0, float(proposal["scale"])
            elif proposal["kind"] == "prior":
                r.cov_kind = 2
            else:
                keep = proposal.get("_chol_dev")
                if keep is None or keep.device != self.device:  # uploaded once: launches must be graph-capturable
                    keep = proposal["_chol_dev"] = self.to_dev(proposal["chol"], np.float64)
                r.cov_kind, r.chol = 1, _ptr(keep)
        else:
            delta, u_acc = replay
            keep = (self.to_dev(delta, np.float64), self.to_dev(u_acc, np.float64))
            r.mode = _lib.EB_RNG_REPLAY
            r.cov_kind = 2 if proposal["kind"] == "prior" else 0
            r.delta, r.u_acc = _ptr(keep[0]), _ptr(keep[1])
        acc = self.accepted_mask(T, W)
        _lib.check(self.lib.eb_gaussian_step(C.byref(st), C.byref(self._prior_c), C.byref(self._like_c), C.byref(r),
                                             _ptr(acc), _ptr(accepted_count), self.stream()), "eb_gaussian_step")
        self.launches += 1
        return acc

    def mt_distgen_step(self, d, num_try, replay=None, accepted_count=None):
        """MTDistGenMove step (multipletry.py:238-514 + mtdistgen.py inside mh.py:56-193), all walkers in one launch.
        replay = (tries [T*W, num_try, D], u_sel [T*W], u_acc [T, W]) drawn on the host in the reference's order."""
        self._require_fused()
        self.flush_adapt()
        T, W, L, D = d.shape
        st = d.c_struct()
        r = _lib.eb_mt_rng()
        r.num_try = int(num_try)
        keep = None
        if replay is None:
            r.mode, r.seed, r.iter_dev = _lib.EB_RNG_PHILOX, self.seed, self.iter_ptr
        else:
            keep = [self.to_dev(x, np.float64) for x in replay]
            r.mode = _lib.EB_RNG_REPLAY
            r.tries, r.u_sel, r.u_acc = [_ptr(x) for x in keep]
        acc = self.accepted_mask(T, W)
        _lib.check(self.lib.eb_mt_distgen_step(C.byref(st), C.byref(self._prior_c), C.byref(self._like_c), C.byref(r),
                                               _ptr(acc), _ptr(accepted_count), self.stream()), "eb_mt_distgen_step")
        self.launches += 1
        return acc

    def pt_swap(self, d, permute=True, adapt=None, replay=None):
        """temper_comps: swap ladder + adaptation (tempering.py:484-649)."""
        T, W, L, D = d.shape
        st = d.c_struct()
        r = _lib.eb_swap_rng()
        r.permute = int(bool(permute))
        keep = None
        if replay is None:
            r.mode = _lib.EB_RNG_PHILOX
            r.seed = self.seed
            r.iter_dev = self.iter_ptr
            if self.lazy_adapt and T > 1 and d.temp_offset == 0:
                r.defer_adapt = 1
                self._lazy_betas = d.betas
        else:
            iperm, i1perm, u = replay
            keep = (self.to_dev(iperm, np.int32), self.to_dev(i1perm, np.int32), self.to_dev(u, np.float64),
                    self.scratch("next_pos", (T, W), torch.int32), self.scratch("u_at", (T, W), torch.float64))
            r.mode = _lib.EB_RNG_REPLAY
            r.iperm, r.i1perm, r.u, r.next_pos, r.u_at = [_ptr(t) for t in keep]
            self.launches += 1 if T > 1 else 0
        if d.inds is not None or L * D > 32 or T > 32:  # shapes whose rows move through staging buffers (k_swap.cu)
            r.row_scratch = _ptr(self.scratch("swap_rows", (T, W, L, D), torch.float64))
            r.logp_scratch = _ptr(self.scratch("swap_logp", (T, W), torch.float64))
            if d.inds is not None:
                r.inds_scratch = _ptr(self.scratch("swap_inds", (T, W, getattr(d, "inds_stride", 0) or L), torch.uint8))
        ad = None
        if adapt is not None:
            ad = _lib.eb_adapt(int(adapt["adaptive"]), int(adapt["stop_adaptation"]), float(adapt["adaptation_lag"]),
                               float(adapt["adaptation_time"]))
        _lib.check(self.lib.eb_pt_swap(C.byref(st), C.byref(r), C.byref(ad) if ad is not None else None,
                                       _ptr(self.ctrl), self.stream()), "eb_pt_swap")
        self.launches += 1

    def resident_run(self, d, a, niter, randomize_split=True, permute=True, adapt=None, accepted_count=None):
        """`niter` whole iterations (StretchMove + temper_comps) in ONE launch with the state resident in shared memory
        (csrc/resident.cuh; ensemble.py:965-1045).  niter = 0 only asks whether this sampler is covered: returns False
        if not (the caller then runs stretch_step + pt_swap), else True / the accept mask of the last iteration."""
        if not self.fused or self.rng != "philox" or d.betas is None or self.periods is not None:
            return False
        self.flush_adapt()
        T, W, L, D = d.shape
        st = d.c_struct()
        sr = _lib.eb_stretch_rng()
        sr.mode, sr.seed, sr.iter_dev, sr.randomize_split = _lib.EB_RNG_PHILOX, self.seed, self.iter_ptr, int(bool(randomize_split))
        wr = _lib.eb_swap_rng()
        wr.mode, wr.seed, wr.iter_dev, wr.permute = _lib.EB_RNG_PHILOX, self.seed, self.iter_ptr, int(bool(permute))
        ad = None
        if adapt is not None:
            ad = _lib.eb_adapt(int(adapt["adaptive"]), int(adapt["stop_adaptation"]), float(adapt["adaptation_lag"]),
                               float(adapt["adaptation_time"]))
        nbytes = int(self.lib.eb_resident_scratch_bytes(C.byref(st)))
        scr = self.scratch("resident", (nbytes,), torch.uint8)
        acc = self.accepted_mask(T, W)
        rc = self.lib.eb_resident_run(C.byref(st), C.byref(self._prior_c), C.byref(self._like_c), float(a), C.byref(sr),
                                      C.byref(wr), C.byref(ad) if ad is not None else None, _ptr(self.ctrl), int(niter),
                                      _ptr(acc), _ptr(accepted_count), _ptr(scr), nbytes, self.stream())
        if rc == 2:   # EB_ERR_UNSUPPORTED: not a shape / configuration of the resident kernel
            return False
        _lib.check(rc, "eb_resident_run")
        if niter == 0:
            return True
        self.launches += 1
        return acc

    def advance_iter(self):
        _lib.check(self.lib.eb_advance_iter(_ptr(self.ctrl), self.stream()), "eb_advance_iter")
        self.launches += 1
