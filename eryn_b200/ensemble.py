"""EnsembleSampler: the reference's driver surface (ensemble.py) over the device hot path.

The host keeps what the reference keeps on the host — move schedule, State bookkeeping, the chain
store — and the walkers live on the GPU between yields: `thin_by` iterations run back to back with
no host synchronisation, and a host `State` is materialised only at yield/store points
(ensemble.py:1013, :1045)."""
import gc
import os
import warnings
import weakref
from collections.abc import Iterable

import numpy as np
import torch

from . import _lib
from .backend import Backend
from .device import DeviceContext, DeviceState
from .model import Model
from .moves import StretchMove, TemperatureControl
from .prior import ProbDistContainer
from .staging import LazyState
from .state import State

__all__ = ["EnsembleSampler", "walkers_independent"]


def walkers_independent(coords_in):
    """ensemble.py:1670-1700 (emcee's linear-independence check)."""
    assert coords_in.ndim == 4
    ntemps, nwalkers, nleaves_max, ndim = coords_in.shape
    coords = coords_in.reshape(ntemps * nwalkers, nleaves_max * ndim)
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C /= C_colmax
    C /= np.sqrt(np.sum(C ** 2, axis=0))
    return np.linalg.cond(C.astype(float)) <= 1e8


class EnsembleSampler(object):
    """Same call surface as eryn.ensemble.EnsembleSampler for the hot path (single branch, in-model moves
    StretchMove / GaussianMove, parallel tempering).  Extra keyword arguments:

      rng     "philox" (default; randoms generated in-kernel) or "numpy-replay" (host NumPy draws in the
              reference's order -> chains identical to the reference under the same np.random.seed)
      seed    Philox seed (default: taken from the NumPy global state at construction)
      device  CUDA device (default: current)

    `log_like_fn` is a device functor from eryn_b200.likelihood (fused kernels) or a callable on CUDA tensors
    (split path).  There is no CPU fallback."""

    def __init__(self, nwalkers, ndims, log_like_fn, priors, provide_groups=False, provide_supplemental=False,
                 tempering_kwargs={}, branch_names=None, nbranches=1, nleaves_max=1, nleaves_min=0, pool=None,
                 moves=None, rj_moves=None, args=None, kwargs=None, backend=None, vectorize=True, periodic=None,
                 update_fn=None, update_iterations=-1, stopping_fn=None, stopping_iterations=-1,
                 fill_zero_leaves_val=-1e300, num_repeats_in_model=1, track_moves=True, info={},
                 rng="philox", seed=None, device=None):
        for name, val in (("provide_groups", provide_groups), ("provide_supplemental", provide_supplemental),
                          ("pool", pool), ("args", args), ("kwargs", kwargs)):
            if val:
                raise NotImplementedError(f"{name} is outside the device hot path built so far (DESIGN.md §7)")
        if fill_zero_leaves_val != -1e300:
            raise NotImplementedError("fill_zero_leaves_val is fixed at -1e300 on the device path")
        if isinstance(ndims, dict):
            branch_names = list(ndims.keys()) if branch_names is None else branch_names
        if branch_names is None:
            branch_names = ["model_0"] if nbranches == 1 else [f"model_{i}" for i in range(nbranches)]
        self.branch_names = list(branch_names)
        nb = len(self.branch_names)
        as_dict = lambda v: dict(v) if isinstance(v, dict) else {n: int(v) for n in self.branch_names}
        self.ndims = as_dict(ndims)
        self.nleaves_max = as_dict(nleaves_max)
        self.nleaves_min = as_dict(nleaves_min if nleaves_min is not None else 0)
        self.nbranches = nb
        self.has_reversible_jump = bool(rj_moves)
        # several branches, several leaves or reversible jump: the multi-branch device path
        self._mb = nb > 1 or max(self.nleaves_max.values()) > 1 or self.has_reversible_jump
        if isinstance(priors, ProbDistContainer):
            priors = {self.branch_names[0]: priors}
        elif isinstance(priors, dict) and not all(n in priors for n in self.branch_names):
            priors = {self.branch_names[0]: ProbDistContainer(priors)}
        priors = {n: (p if isinstance(p, ProbDistContainer) else ProbDistContainer(p)) for n, p in priors.items()}
        name = self.branch_names[0]
        self.priors = priors
        self.key_order = {n: priors[n].key_order for n in self.branch_names}
        self.nwalkers = int(nwalkers)
        self.num_repeats_in_model = int(num_repeats_in_model)
        self.num_repeats_rj = 1
        self.track_moves = track_moves
        self.update_fn, self.update_iterations = update_fn, update_iterations
        self.stopping_fn, self.stopping_iterations = stopping_fn, stopping_iterations

        # tempering (ensemble.py:321-334): effective dimension = sum over branches of nleaves_max * ndim
        if tempering_kwargs == {}:
            self.ntemps = 1
            self.temperature_control = None
        else:
            total_ndim = sum(self.nleaves_max[n] * self.ndims[n] for n in self.branch_names)
            self.temperature_control = TemperatureControl(total_ndim, nwalkers, **tempering_kwargs)
            self.ntemps = self.temperature_control.ntemps

        # move schedule (ensemble.py:349-378)
        if moves is None:
            if self._mb:
                raise NotImplementedError("several branches / leaves need an in-model move that handles them on the device: "
                                          "pass moves=GroupStretchMove(nfriends=...)")
            self.moves = [StretchMove(temperature_control=self.temperature_control, a=2.0)]
            self.weights = [1.0]
        elif isinstance(moves, Iterable):
            try:
                self.moves, self.weights = [list(tmp) for tmp in zip(*moves)]
            except TypeError:
                self.moves = moves
                self.weights = np.ones(len(moves))
        else:
            self.moves = [moves]
            self.weights = [1.0]
        self.weights = np.atleast_1d(self.weights).astype(float)
        self.weights /= np.sum(self.weights)

        # random streams (ensemble.py:604, :651-652): the private stream is a copy of the global state
        state = np.random.get_state()
        self._random = np.random.mtrand.RandomState()
        self._random.set_state(state)
        if seed is None:
            seed = int(state[1][0]) | (int(state[1][1]) << 32)

        if self._mb and periodic is not None:
            raise NotImplementedError("periodic parameters are handled by the single-branch kernels (DESIGN.md §7)")
        if self._mb:
            from .moves import DistributionGenerateRJ, GroupStretchMove
            from .multibranch import MBContext
            groups = [m for m in self.moves if isinstance(m, GroupStretchMove)]
            if len(groups) != len(self.moves):
                raise NotImplementedError("in-model moves of a multi-branch sampler: GroupStretchMove")
            if len({(m.nfriends, str(m.friend_key)) for m in groups}) != 1:
                raise ValueError("all GroupStretchMoves of one sampler share one friend table (nfriends, friend_key)")
            self.ctx = MBContext(priors, log_like_fn, self.branch_names, self.ndims, self.nleaves_max, self.nleaves_min,
                                 nfriends=groups[0].nfriends, friend_key=groups[0].friend_key, device=device, rng=rng,
                                 seed=seed)
            # rj move schedule (ensemble.py:380-507): True / "together" = one DistributionGenerateRJ over all branches
            self.rj_moves, self.rj_weights = None, None
            if self.has_reversible_jump:
                if rj_moves is True or rj_moves == "together":
                    self.rj_moves = [DistributionGenerateRJ(priors, nleaves_max=self.nleaves_max,
                                                            nleaves_min=self.nleaves_min,
                                                            temperature_control=self.temperature_control)]
                elif rj_moves == "iterate_branches":  # ensemble.py:434-449: the branches one after the other in ONE move
                    self.rj_moves = [DistributionGenerateRJ(priors, nleaves_max=self.nleaves_max,
                                                            nleaves_min=self.nleaves_min,
                                                            temperature_control=self.temperature_control,
                                                            gibbs_sampling_setup=list(self.branch_names))]
                elif rj_moves == "separate_branches":  # ensemble.py:451-472: one move per branch, one chosen per iteration
                    self.rj_moves = [DistributionGenerateRJ(priors, nleaves_max=self.nleaves_max,
                                                            nleaves_min=self.nleaves_min,
                                                            temperature_control=self.temperature_control,
                                                            gibbs_sampling_setup=[bn]) for bn in self.branch_names]
                elif isinstance(rj_moves, str):
                    raise ValueError("When providing a str for rj_moves, must be 'together', 'iterate_branches', or "
                                     f"'separate_branches'. Input is {rj_moves}")
                else:
                    self.rj_moves = list(rj_moves) if isinstance(rj_moves, Iterable) else [rj_moves]
                self.rj_weights = np.ones(len(self.rj_moves)) / len(self.rj_moves)
        else:
            self.rj_moves, self.rj_weights = None, None
            per = None
            if periodic is not None:  # ensemble.py:338-347
                if not isinstance(periodic, dict):
                    raise ValueError("periodic must be PeriodicContainer or dict if not None.")
                per = periodic.get(name)
            self.ctx = DeviceContext(priors[name], log_like_fn, device=device, rng=rng, seed=seed, branch_name=name,
                                     periodic=per)
        self.log_like_fn = self.ctx.like
        if self.temperature_control is not None:
            self.temperature_control.bind(self.ctx)
        for move in self.moves + (self.rj_moves or []):
            if self.temperature_control is not None and move.temperature_control is None:
                move.temperature_control = self.temperature_control  # ensemble.py:516-525
            move.bind(self.ctx)
            if periodic is not None and move.periodic is None:
                move.periodic = periodic  # ensemble.py:528-536
            move.accepted = np.zeros((self.ntemps, self.nwalkers))  # ensemble.py:539-540

        self.backend = Backend() if backend is None else backend
        self.all_moves = {}
        if self.track_moves:
            counts = {}
            for move in self.moves + (self.rj_moves or []):
                mn = move.__class__.__name__
                counts[mn] = counts.get(mn, -1) + 1
                self.all_moves[f"{mn}_{counts[mn]}"] = move  # ensemble.py:563-583
            self.move_keys = list(self.all_moves.keys())
        else:
            self.move_keys = None
        self._previous_state = None
        if not self.backend.initialized:
            self.backend.reset(self.nwalkers, self.ndims, nleaves_max=self.nleaves_max, ntemps=self.ntemps,
                               branch_names=self.branch_names, rj=self.has_reversible_jump, moves=self.move_keys,
                               key_order=self.key_order, **info)
        self._dstate = None
        self._graphs, self._warm, self._ring, self._pending = {}, set(), None, []
        self.force_eager = False  # True: launch every iteration eagerly and store synchronously (the checked baseline)
        _weak_flush = weakref.WeakMethod(self._flush_store)  # no reference cycle sampler <-> backend
        self.backend._flush_cb = lambda: (_weak_flush() or (lambda: None))()  # getters see every step that was yielded

    # ---- small mirrors ---------------------------------------------------------------------------------
    @property
    def random_state(self):
        return self._random.get_state()

    @property
    def iteration(self):
        return self.backend.iteration

    @property
    def acceptance_fraction(self):
        return self.backend.accepted / float(self.backend.iteration)

    def get_chain(self, **kwargs):
        return self.backend.get_chain(**kwargs)

    def get_log_like(self, **kwargs):
        return self.backend.get_log_like(**kwargs)

    def get_nleaves(self, **kwargs):
        return self.backend.get_nleaves(**kwargs)

    def get_inds(self, **kwargs):
        return self.backend.get_inds(**kwargs)

    def get_log_prior(self, **kwargs):
        return self.backend.get_log_prior(**kwargs)

    def get_betas(self, **kwargs):
        return self.backend.get_betas(**kwargs)

    def get_last_sample(self):
        return self.backend.get_last_sample()

    def get_model(self):
        """ensemble.py:780-806"""
        return Model(self.log_like_fn, self.compute_log_like, self.compute_log_prior, self.temperature_control, map,
                     self._random)

    def _mb_eval(self, coords, inds):
        st = State({n: np.asarray(coords[n], dtype=np.float64) for n in self.branch_names},
                   inds=None if inds is None else {n: inds[n] for n in self.branch_names})
        d = self.ctx.upload(st)
        self.ctx.eval_state(d)
        return d.logp.cpu().numpy(), d.logl.cpu().numpy()

    def compute_log_prior(self, coords, inds=None, supps=None, branch_supps=None):
        """ensemble.py:1127 — evaluated on the device; returns an ndarray [ntemps, nwalkers]."""
        if self._mb:
            return self._mb_eval(coords, inds)[0]
        name = self.branch_names[0]
        c = coords[name] if isinstance(coords, dict) else coords
        T, W, L, D = c.shape
        q = self.ctx.to_dev(c, np.float64).view(T * W, L, D)
        ii = None
        if inds is not None:
            ii = self.ctx.to_dev((inds[name] if isinstance(inds, dict) else inds).astype(np.uint8)).view(T * W, L)
        out = self.ctx.box_log_prior(q, ii).cpu().numpy().reshape(T, W)
        if np.any(np.isnan(out)):
            raise ValueError("The prior function is returning Nan.")
        return out

    def compute_log_like(self, coords, inds=None, logp=None, supps=None, branch_supps=None):
        """ensemble.py:1219 — evaluated on the device; returns (ndarray [ntemps, nwalkers], None)."""
        if self._mb:
            return self._mb_eval(coords, inds)[1], None
        name = self.branch_names[0]
        c = coords[name] if isinstance(coords, dict) else coords
        if np.any(np.isinf(c)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(c)):
            raise ValueError("At least one parameter value was NaN")
        st = State({name: c}, inds=None if inds is None else {name: (inds[name] if isinstance(inds, dict) else inds)})
        d = self.ctx.upload(st)
        self.ctx.eval_state(d)
        ll = d.logl.cpu().numpy()
        if np.any(np.isnan(ll)):
            raise ValueError("The likelihood function is returning Nan.")
        return ll, None

    # ---- the loop -------------------------------------------------------------------------------------
    def sample(self, initial_state, iterations=1, tune=False, skip_initial_state_check=True, thin_by=1, store=True,
               progress=False):
        """Advance the chain as a generator (ensemble.py:808-1045).

        Production (philox) runs of the single-branch samplers take the device-resident fast path: the launches of
        `thin_by` iterations are captured once into CUDA graphs and replayed, stored steps leave through the staging
        ring (staging.py) and the yielded State materialises on the host only if the consumer looks at it."""
        if iterations is None and store:
            raise ValueError("'store' must be False when 'iterations' is None")
        state = State(initial_state, copy=True)
        name = self.branch_names[0]
        for nm, branch in state.branches.items():
            if tuple(branch.shape) != (self.ntemps, self.nwalkers, self.nleaves_max[nm], self.ndims[nm]):
                raise ValueError("incompatible input dimensions")
        if (not skip_initial_state_check) and (not walkers_independent(state.branches[name].coords)):
            raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                             "linearly independent for the best performance")
        tc = self.temperature_control
        if tc is not None:
            if state.betas is not None:
                if state.betas.shape[0] != self.ntemps:
                    raise ValueError("Input state has inverse temperatures (betas), but not the correct number of "
                                     "temperatures according to sampler inputs.")
                tc.betas = state.betas.copy()  # ensemble.py:915-921
            else:
                state.betas = tc.betas.copy()
        d = self._upload_resident(state, None if tc is None else tc.betas_dev)
        if state.log_prior is None or state.log_like is None:  # ensemble.py:898-912: only what is missing is evaluated
            keep_lp = None if state.log_prior is None else d.logp.clone()
            keep_ll = None if state.log_like is None else d.logl.clone()
            self.ctx.eval_state(d)
            if keep_lp is not None:
                d.logp.copy_(keep_lp)
            if keep_ll is not None:
                d.logl.copy_(keep_ll)
        logl0, logp0 = d.logl.cpu().numpy(), d.logp.cpu().numpy()
        if np.shape(logl0) != (self.ntemps, self.nwalkers) or np.shape(logp0) != (self.ntemps, self.nwalkers):
            raise ValueError("incompatible input dimensions")
        if np.any(np.isnan(logl0)):
            raise ValueError("The initial log_like was NaN")
        if np.any(np.isinf(logl0)):
            raise ValueError("The initial log_like was +/- infinite")
        if np.any(np.isnan(logp0)):
            raise ValueError("The initial log_prior was NaN")
        if np.any(np.isinf(logp0)):
            raise ValueError("The initial log_prior was +/- infinite")
        thin_by = int(thin_by)
        if thin_by <= 0:
            raise ValueError("Invalid thinning argument")
        if store:
            self.backend.grow(iterations, None)
        model = self.get_model()
        self._dstate = d
        fast = (self.ctx.rng == "philox" and not self._mb and not tune and self.ctx.fused and not self.force_eager
                and all(getattr(m, "graphable", False) for m in self.moves))
        gen = self._sample_resident if fast else self._sample_eager
        try:
            yield from gen(model, d, iterations, thin_by, store, tune)
        finally:
            self._flush_store()
        self._check_device_error()

    def _upload_resident(self, state, betas_dev):
        """the device state of this sampler keeps its buffers between calls (captured graphs stay valid)"""
        d = self._dstate
        br = state.branches[self.branch_names[0]]
        if (self._mb or d is None or d.shape != tuple(br.coords.shape) or (d.inds is None) != bool(np.all(br.inds))
                or d.betas is not betas_dev):
            self._graphs, self._warm, self._ring, self._k12_ok = {}, set(), None, None
            return self.ctx.upload(state, betas=betas_dev)
        d.coords.copy_(torch.from_numpy(np.ascontiguousarray(br.coords, dtype=np.float64)))
        if d.inds is not None:
            d.inds.copy_(torch.from_numpy(np.ascontiguousarray(br.inds.astype(np.uint8))))
        if state.log_like is not None:
            d.logl.copy_(torch.from_numpy(np.ascontiguousarray(state.log_like, dtype=np.float64)))
        if state.log_prior is not None:
            d.logp.copy_(torch.from_numpy(np.ascontiguousarray(state.log_prior, dtype=np.float64)))
        return d

    def _check_device_error(self, ctrl_bytes=None):
        """eb_ctrl.error is set by a kernel whose bounded wait ran out (a swap pass that never saw all its CTAs): the
        chain is not trustworthy from that iteration on"""
        err = int(self.ctx.read_ctrl().error) if ctrl_bytes is None else \
            int(_lib.eb_ctrl.from_buffer_copy(ctrl_bytes.tobytes()).error)
        if err:
            raise _lib.ErynB200Error(f"device error {err} (EB_DEVERR_*) raised inside a kernel: a bounded wait of the "
                                     "swap pass timed out; the chain after that iteration is invalid")

    def _apply_host_edits(self, d, host, before=None):
        """A consumer (update_fn, or the body of a `for state in sampler.sample(...)` loop) may edit the yielded state
        in place, as with the reference, where the sampler carries the same object on (ensemble.py:1030-1045): bring
        such edits back to the device.  `before` = pristine copies for states that were downloaded eagerly."""
        tc = self.temperature_control
        if isinstance(host, LazyState):
            mod = host.modified()
        else:
            mod = {}
            br = host.branches[self.branch_names[0]]
            if not np.array_equal(br.coords, before["coords"], equal_nan=True):
                mod["coords"] = br.coords
            if not np.array_equal(host.log_like, before["logl"]):
                mod["logl"] = host.log_like
            if not np.array_equal(host.log_prior, before["logp"]):
                mod["logp"] = host.log_prior
            if before.get("inds") is not None and not np.array_equal(br.inds, before["inds"]):
                mod["inds"] = br.inds
            if host.betas is not None and before.get("betas") is not None and not np.array_equal(host.betas, before["betas"]):
                mod["betas"] = host.betas
        if not mod:
            return
        if "coords" in mod:
            d.coords.copy_(torch.from_numpy(np.ascontiguousarray(mod["coords"], dtype=np.float64)).view_as(d.coords))
        if "logl" in mod:
            d.logl.copy_(torch.from_numpy(np.ascontiguousarray(mod["logl"], dtype=np.float64)))
        if "logp" in mod:
            d.logp.copy_(torch.from_numpy(np.ascontiguousarray(mod["logp"], dtype=np.float64)))
        if "inds" in mod:
            if d.inds is None:
                raise NotImplementedError("the yielded state's leaf flags were edited, but the device state was uploaded "
                                          "without leaf flags (all leaves active)")
            d.inds.copy_(torch.from_numpy(np.ascontiguousarray(mod["inds"].astype(np.uint8))).view_as(d.inds))
        if "betas" in mod and tc is not None:
            tc.betas = mod["betas"]

    # ---- eager path: replay mode, several branches / reversible jump, tuning ----------------------------------------
    def _sample_eager(self, model, d, iterations, thin_by, store, tune):
        tc = self.temperature_control
        acc_total = torch.zeros((self.ntemps, self.nwalkers), dtype=torch.int32, device=self.ctx.device)
        rj_total = torch.zeros((self.ntemps, self.nwalkers), dtype=torch.int32, device=self.ctx.device)
        in_model_swaps = None
        i = 0
        it_range = iter(int, 1) if iterations is None else range(iterations)
        for _ in it_range:
            for inner in range(thin_by):
                acc_total.zero_()  # ensemble.py:968: `accepted` restarts every inner iteration
                for repeat in range(self.num_repeats_in_model):
                    mi = self._random.choice(len(self.moves), p=self.weights)  # ensemble.py:971 (1 uniform)
                    move = self.moves[mi]
                    d, acc = move.propose(model, d)  # device resident: no host sync
                    acc_total += acc
                    if tune:
                        move.tune(d, acc)
                if self.has_reversible_jump:  # ensemble.py:986-1006
                    if store and inner == thin_by - 1 and tc is not None and self.ntemps > 1:
                        in_model_swaps = tc.swaps_accepted  # ensemble.py:977: read before the rj move swaps again
                    rj_total.zero_()
                    for repeat in range(self.num_repeats_rj):
                        ri = self._random.choice(len(self.rj_moves), p=self.rj_weights)  # ensemble.py:990
                        d, racc = self.rj_moves[ri].propose(model, d)
                        rj_total += racc
                if store and inner == thin_by - 1:  # ensemble.py:1013-1028
                    host = self.ctx.download(d, random_state=self.random_state)
                    maf = {k: m.acceptance_fraction for k, m in self.all_moves.items()} if self.track_moves else None
                    swaps = tc.swaps_accepted if (tc is not None and self.ntemps > 1) else None
                    if self.has_reversible_jump:
                        swaps = in_model_swaps
                    self.backend.save_step(host, acc_total.cpu().numpy(), swaps_accepted=swaps,
                                           rj_accepted=rj_total.cpu().numpy() if self.has_reversible_jump else None,
                                           moves_accepted_fraction=maf)
                else:
                    host = None
                if (self.update_iterations > 0 and self.update_fn is not None
                        and (i + 1) % self.update_iterations == 0):  # ensemble.py:1030-1036, every inner iteration
                    if host is None:
                        host = self.ctx.download(d, random_state=self.random_state)
                    d = self._call_with_edits(d, host, lambda h: self.update_fn(i, h, self))
                    host = None if inner != thin_by - 1 else host
                i += 1
            self._check_device_error()
            if host is None:
                host = self.ctx.download(d, random_state=self.random_state)
            before = self._pristine_of(host)
            yield host
            d = self._sync_edits(d, host, before)

    def _pristine_of(self, host):
        if self._mb:
            return dict(coords={n: b.coords.copy() for n, b in host.branches.items()},
                        inds={n: b.inds.copy() for n, b in host.branches.items()}, logl=host.log_like.copy(),
                        logp=host.log_prior.copy(), betas=None if host.betas is None else host.betas.copy())
        br = host.branches[self.branch_names[0]]
        return dict(coords=br.coords.copy(), inds=br.inds.copy(), logl=host.log_like.copy(), logp=host.log_prior.copy(),
                    betas=None if host.betas is None else host.betas.copy())

    def _sync_edits(self, d, host, before):
        """eager path: after a consumer had the host state, carry its edits to the device"""
        tc = self.temperature_control
        if self._mb:
            same = all(np.array_equal(host.branches[n].coords, before["coords"][n], equal_nan=True)
                       and np.array_equal(host.branches[n].inds, before["inds"][n]) for n in host.branches)
            same = same and np.array_equal(host.log_like, before["logl"]) and np.array_equal(host.log_prior, before["logp"])
            if not same:
                d = self.ctx.upload(host, betas=None if tc is None else tc.betas_dev)
            if host.betas is not None and tc is not None and not np.array_equal(host.betas, before["betas"]):
                tc.betas = host.betas
            return d
        if d.inds is None and not bool(np.all(host.branches[self.branch_names[0]].inds)):
            raise NotImplementedError("the yielded state's leaf flags were edited, but the device state was uploaded "
                                      "without leaf flags (all leaves active)")
        self._apply_host_edits(d, host, before)
        return d

    def _call_with_edits(self, d, host, fn):
        before = self._pristine_of(host)
        fn(host)
        return self._sync_edits(d, host, before)

    # ---- resident path: CUDA-graph replay + staged stores (production) ----------------------------------------------
    _CHUNK = 32  # iterations per captured graph of a single-move sampler (bounds capture time and graph size)

    def _graph(self, key, body):
        """capture `body` (kernel launches only) once per key, then replay"""
        g = self._graphs.get(key)
        if g is None:
            cap = getattr(self, "_capture_stream", None)
            if cap is None:
                cap = self._capture_stream = torch.cuda.Stream(device=self.ctx.device)
            cap.wait_stream(torch.cuda.current_stream(self.ctx.device))
            graph = torch.cuda.CUDAGraph()
            l0 = self.ctx.launches
            # Graphs of a dead sampler must not be destroyed while this capture is open (their reset is a CUDA call that
            # invalidates it): collect garbage now and keep the collector off for the few launches of the capture.
            gc.collect()
            gc_was = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph, stream=cap):
                    body()
            finally:
                if gc_was:
                    gc.enable()
            g = self._graphs[key] = (graph, self.ctx.launches - l0)
            self.ctx.launches = l0
        g[0].replay()
        self.ctx.launches += g[1]

    _K12_MIN_ITERS = 8       # below this a block runs as replayed per-launch kernels (K12 pays ~25 us per launch)
    # ntemps * nwalkers up to which blocks go through K12.  0 = never (the default): measured against a replayed graph
    # of the SAME block, whose launches are chained by programmatic dependent launch, K12 does not win at any size
    # (8 x 256: 12.2 vs 11.7 us per iteration, config 2: 32.4 vs 19.5; bench.py extra.resident_kernel).
    # EB_RESIDENT_MAX_WALKERS opts in.
    _K12_MAX_WALKERS = 0

    def _k12_applies(self, mv, d):
        """a plain StretchMove of a tempered single-leaf sampler on an ensemble small enough for the resident kernel"""
        ok = getattr(self, "_k12_ok", None)
        if ok is None:
            from .moves.stretch import StretchMove
            T, W, L, D = d.shape
            limit = int(os.environ.get("EB_RESIDENT_MAX_WALKERS", self._K12_MAX_WALKERS))
            ok = (type(mv) is StretchMove and mv.temperature_control is not None and T >= 2 and T * W <= limit
                  and d.inds is None and mv._single_branch_splits(d.branch_name, L, D) == [(0, 0, L * D)]
                  and self.ctx.resident_run(d, mv.a, 0) is True)
            self._k12_ok = ok
        return ok

    def _advance_resident(self, model, d, n):
        """n inner iterations, device resident, no host synchronisation"""
        R = self.num_repeats_in_model
        mis = self._random.choice(len(self.moves), p=self.weights, size=n * R)  # ensemble.py:971, one uniform per move
        if len(self.moves) == 1:
            mv = self.moves[0]
            todo = n * R
            if 0 not in self._warm:  # first proposal eagerly: module load, scratch buffers, attribute setup
                mv.propose(model, d)
                self._warm.add(0)
                todo -= 1
            if todo >= self._K12_MIN_ITERS and self._k12_applies(mv, d):
                # small ensembles: the whole block in ONE launch with the state resident in shared memory (K12,
                # csrc/resident.cuh) — same chain bit for bit, no kernel boundaries between iterations
                tc = mv.temperature_control
                ad = None
                if tc.adaptive:
                    ad = dict(adaptive=True, stop_adaptation=tc.stop_adaptation, adaptation_lag=tc.adaptation_lag,
                              adaptation_time=tc.adaptation_time)
                T, W = d.shape[0], d.shape[1]
                self.ctx.resident_run(d, mv.a, todo, randomize_split=mv.randomize_split, permute=tc.permute, adapt=ad,
                                      accepted_count=mv._count_buffer(self.ctx, T, W))
                mv._host_tick(todo)
                return
            while todo > 0:
                k = min(todo, self._CHUNK)

                def body(k=k):
                    for _ in range(k):
                        mv.propose(model, d)
                if (0, k) in self._graphs:
                    mv._host_tick(k)
                self._graph((0, k), body)
                todo -= k
            return
        for mi in mis:
            mi = int(mi)
            mv = self.moves[mi]
            if mi not in self._warm:
                mv.propose(model, d)
                self._warm.add(mi)
                continue
            if (mi, 1) in self._graphs:
                mv._host_tick(1)
            self._graph((mi, 1), lambda mv=mv: mv.propose(model, d))

    def _lazy_adapt_applies(self, d):
        """the loop consists of plain StretchMove / GaussianMove proposals on a tempered ensemble: the swap pass may leave its
        ladder adaptation to the next move kernel (DeviceContext.lazy_adapt)"""
        if os.environ.get("EB_LAZY_ADAPT", "1") == "0" or self.num_repeats_in_model != 1:
            return False
        from .moves import GaussianMove
        T, W, L, D = d.shape
        # worth it while the pass is latency-bound (config 2: 10.2 -> 7.6 us); on grids of thousands of CTAs (config 4) the
        # per-CTA fold in the move kernel costs more than the pass saves
        if T <= 1 or d.inds is not None or T * W > self._LAZY_MAX_WALKERS:
            return False
        for mv in self.moves:   # the kernels that apply a deferred adaptation themselves: stretch and Gaussian steps
            if type(mv) not in (StretchMove, GaussianMove) or mv.temperature_control is None:
                return False
            if mv._single_branch_splits(d.branch_name, L, D) != [(0, 0, L * D)]:
                return False
        return True

    _LAZY_MAX_WALKERS = 131072

    def _sample_resident(self, model, d, iterations, thin_by, store, tune):
        if self._lazy_adapt_applies(d):
            self.ctx.lazy_begin(d.betas)
        try:
            yield from self._sample_resident_loop(model, d, iterations, thin_by, store, tune)
        finally:
            self.ctx.lazy_end()

    def _sample_resident_loop(self, model, d, iterations, thin_by, store, tune):
        from .staging import StoreRing
        tc = self.temperature_control
        T, W = self.ntemps, self.nwalkers
        name = self.branch_names[0]
        leaf_moves = []
        for m in self.moves:
            leaf_moves += list(getattr(m, "moves", [m]))
        counts = [m._count_buffer(self.ctx, T, W) for m in leaf_moves]
        if self._ring is None:
            self._ring = StoreRing(self.ctx, d, count_buffers=counts, nslots=4, fill=self.backend.store_missing_leaves)
            self._ring.on_evict = self._drain_ticket
        ring = self._ring
        # `accepted` of a stored step = the accepts of its LAST inner iteration (ensemble.py:968).  With one mask-writing
        # kernel per iteration that is the accept mask left in the scratch buffer; otherwise (CombineMove, repeats) the
        # difference of the in-kernel counters around that iteration.
        simple = self.num_repeats_in_model == 1 and all(not hasattr(m, "moves") for m in self.moves)
        upd = self.update_iterations if (self.update_iterations > 0 and self.update_fn is not None) else 0
        i = 0
        it_range = iter(int, 1) if iterations is None else range(iterations)
        def snapshot(copy):
            ticket = ring.push(copy=copy)
            lazy = LazyState(ticket, name, random_state=self.random_state)
            ticket.state_ref = weakref.ref(lazy)
            return ticket, lazy

        for _ in it_range:
            done = 0
            pre = None
            lazy = None
            while done < thin_by:
                seg = thin_by - done
                if not simple and seg > 1:
                    seg -= 1  # stop before the last inner iteration to take the counter snapshot
                if upd:
                    seg = min(seg, upd - (i % upd))
                if not simple and done + seg == thin_by:
                    pre = torch.stack(counts).sum(0, dtype=torch.int32)
                self._advance_resident(model, d, seg)
                i += seg
                done += seg
                if done == thin_by:
                    # ---- stored step / yield point (ensemble.py:1013-1045): snapshot, async copy, lazy host State.  It
                    #      comes BEFORE update_fn, as in the reference: the stored sample is the un-edited one
                    acc_delta = None
                    if not simple:
                        acc_delta = torch.stack(counts).sum(0, dtype=torch.int32) - pre
                    ticket, lazy = snapshot(copy=store)  # not stored: device snapshot only, copied if somebody looks
                    if store:
                        ticket.meta = dict(acc_delta=acc_delta, nprop=[m.num_proposals for m in leaf_moves],
                                           base=[m._accepted.copy() for m in leaf_moves], random_state=lazy.random_state)
                        self._pending.append(ticket)
                        self._drain_ready()
                if upd and i % upd == 0:  # ensemble.py:1030-1036 (every inner iteration, index of that iteration)
                    st = lazy if done == thin_by else snapshot(copy=True)[1]
                    self.update_fn(i - 1, st, self)
                    self._apply_host_edits(d, st)
            yield lazy
            self._apply_host_edits(d, lazy)

    # ---- deferred Backend.save_step -----------------------------------------------------------------------------------
    def _drain_ticket(self, ticket):
        """store every pending sample up to and including `ticket` (in order)"""
        while self._pending and ticket in self._pending:
            self._store_ticket(self._pending.pop(0))

    def _drain_ready(self):
        while self._pending and self._pending[0].ready():
            self._store_ticket(self._pending.pop(0))

    def _flush_store(self):
        while self._pending:
            self._store_ticket(self._pending.pop(0))

    def _store_ticket(self, ticket):
        a = ticket.views()
        meta = ticket.meta
        ctrl = _lib.eb_ctrl.from_buffer_copy(a["ctrl"].tobytes())
        if ctrl.error:
            self._check_device_error(a["ctrl"])
        T, W = self.ntemps, self.nwalkers
        name = self.branch_names[0]
        acc = a["accepted"].astype(np.int64) if meta["acc_delta"] is None else meta["acc_delta"].cpu().numpy()
        swaps = None
        if self.temperature_control is not None and T > 1:
            swaps = np.array(ctrl.swaps_accepted[: T - 1], dtype=np.float64)
        maf = None
        if self.track_moves:
            per_leaf = {}
            k = 0
            for mv in self.moves:
                for leaf in getattr(mv, "moves", [mv]):
                    per_leaf[id(leaf)] = (meta["base"][k] + a[f"count{k}"]) / max(meta["nprop"][k], 1)
                    k += 1
            maf = {}
            for key, mv in self.all_moves.items():
                if hasattr(mv, "moves"):
                    maf[key] = np.mean([per_leaf[id(leaf)] for leaf in mv.moves], axis=0)
                else:
                    maf[key] = per_leaf[id(mv)]
        inds = a["inds"].astype(bool) if "inds" in a else None
        self.backend.save_arrays({name: a["coords"]}, None if inds is None else {name: inds}, a["logl"], a["logp"],
                                 a.get("betas"), acc, swaps_accepted=swaps, moves_accepted_fraction=maf,
                                 random_state=meta["random_state"])
        ticket.release()

    def run_mcmc(self, initial_state, nsteps, burn=None, post_burn_update=False, **kwargs):
        """ensemble.py:1047-1125"""
        if initial_state is None:
            if self._previous_state is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
            initial_state = self._previous_state
        if burn is not None and burn != 0:
            bk = dict(kwargs)
            bk["store"] = False
            # the reference iterates `burn` times with thin_by = 1 and drops every state but the last; one yield after
            # `burn` inner iterations is the same chain (update_fn is checked per inner iteration either way) and keeps
            # the walkers on the device throughout
            bk["thin_by"] = int(burn)
            i = int(burn)
            for results in self.sample(initial_state, iterations=1, **bk):
                pass
            if post_burn_update and self.update_fn is not None:
                self.update_fn(i, results, self)
            initial_state = results
        if nsteps == 0:
            return initial_state
        results = None
        i = 0
        for results in self.sample(initial_state, iterations=nsteps, **kwargs):
            if self.stopping_iterations > 0 and self.stopping_fn is not None and (i + 1) % self.stopping_iterations == 0:
                if self.stopping_fn(i, results, self):
                    break
            i += 1
        self._previous_state = results
        return results
