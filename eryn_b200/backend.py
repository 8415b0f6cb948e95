"""In-memory chain store with the getters of eryn.backends.Backend that the hot path's callers use
(backend.py:616-1091).  HDF5 storage is host I/O outside this build's scope."""
import ctypes
import mmap
import os
import threading

import numpy as np

from .state import State

__all__ = ["Backend"]

# ---- populating freshly grown chain memory off the sampler's thread ----------------------------------------------------
# A stored sample is copied into chain memory that np.empty has only reserved: every 4 KiB page faults on first touch
# (~1 us each; 1.3 ms for a 5.3 MB config-2 sample — twice the GPU time of the 25 iterations between two stored samples).
# grow() therefore asks the kernel to populate the new pages from a helper thread, oldest sample slot first, with
# MADV_POPULATE_WRITE (Linux >= 5.14): unlike a touching loop it cannot race with the stores (contents are never
# changed) and the call runs without the GIL.  Transparent huge pages are requested first where the system allows them.
_MADV_HUGEPAGE, _MADV_POPULATE_WRITE = 14, 23
_PREFAULT_MIN_BYTES = 8 << 20
_PREFAULT_MAX_BYTES = 1 << 30
_PREFAULT_CHUNK = 4 << 20
_PREFAULT_THREADS = 4     # page zeroing is per-core work: one thread populates ~4-5 GB/s


# ---- the copy of a stored sample into the chain, split over a few threads (one host thread copies ~4 GB/s here) ------
_COPY_MIN_BYTES = 2 << 20
_COPY_THREADS = 4
_copy_pool = None


def _store_copy(dst, src):
    """dst[...] = src for one sample; large samples are split along the first axis over a small thread pool (NumPy copies
    release the GIL)"""
    global _copy_pool
    src = np.asarray(src)
    if dst.nbytes < _COPY_MIN_BYTES or src.shape != dst.shape or dst.shape[0] < 2:
        dst[...] = src
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="eryn_b200-store")
    n = min(_COPY_THREADS, dst.shape[0])
    edges = [dst.shape[0] * k // n for k in range(n + 1)]
    futs = [_copy_pool.submit(np.copyto, dst[a:b], src[a:b]) for a, b in zip(edges[1:-1], edges[2:])]
    np.copyto(dst[edges[0]:edges[1]], src[edges[0]:edges[1]])     # this thread takes the first part
    for f in futs:
        f.result()


def _prefault_async(arr, first_row):
    """populate the pages of arr[first_row:] in the background; returns the threads (or None if there is nothing to do)"""
    if arr.size == 0 or first_row >= len(arr) or not arr.flags.c_contiguous:
        return None
    row_bytes = arr.strides[0]
    page = mmap.PAGESIZE
    lo = arr.ctypes.data + first_row * row_bytes
    hi = arr.ctypes.data + len(arr) * row_bytes
    lo = (lo + page - 1) // page * page
    hi = hi // page * page
    if hi - lo < _PREFAULT_MIN_BYTES:
        return None
    hi = min(hi, lo + _PREFAULT_MAX_BYTES)   # a long run reserves far more than it may ever touch: populate the head only
    try:
        madvise = ctypes.CDLL(None, use_errno=True).madvise
    except (OSError, AttributeError):
        return None
    madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    madvise.restype = ctypes.c_int

    madvise(lo, hi - lo, _MADV_HUGEPAGE)                           # advisory; ignored where THP is off
    nthreads = max(1, min(_PREFAULT_THREADS, (os.cpu_count() or 2) // 2))

    def work(j, keep=arr):   # `keep` pins the array: its memory cannot be unmapped while the thread runs
        # thread j takes chunks j, j + nthreads, ...: all of them advance from the front, where the next store lands
        at = lo + j * _PREFAULT_CHUNK
        while at < hi:
            n = min(_PREFAULT_CHUNK, hi - at)
            if madvise(at, n, _MADV_POPULATE_WRITE) != 0:          # old kernel / no memory: first touch will do it
                break
            at += nthreads * _PREFAULT_CHUNK
    threads = [threading.Thread(target=work, args=(j,), name="eryn_b200-prefault", daemon=True) for j in range(nthreads)]
    for th in threads:
        th.start()
    return threads


class Backend(object):
    def __init__(self, store_missing_leaves=np.nan):
        self.initialized = False
        self.store_missing_leaves = store_missing_leaves
        self._flush_cb = None  # the sampler's deferred stores (staging.py): every getter drains them first

    def _flush(self):
        if self._flush_cb is not None:
            self._flush_cb()

    def reset(self, nwalkers, ndims, nleaves_max=1, ntemps=1, branch_names=None, nbranches=1, rj=False,
              moves=None, key_order=None, **info):
        self.nwalkers, self.ntemps = int(nwalkers), int(ntemps)
        self.branch_names = list(branch_names) if branch_names is not None else ["model_0"]
        self.ndims = ndims if isinstance(ndims, dict) else {n: ndims for n in self.branch_names}
        self.nleaves_max = nleaves_max if isinstance(nleaves_max, dict) else {n: nleaves_max for n in self.branch_names}
        self.rj = rj
        self.move_keys = list(moves) if moves is not None else None
        self.key_order = key_order
        self.iteration = 0
        self.accepted = np.zeros((self.ntemps, self.nwalkers), dtype=int)
        self.swaps_accepted = np.zeros((self.ntemps - 1,), dtype=int)
        self.rj_accepted = np.zeros((self.ntemps, self.nwalkers), dtype=int) if rj else None
        self.chain = {n: np.empty((0, self.ntemps, self.nwalkers, self.nleaves_max[n], self.ndims[n]))
                      for n in self.branch_names}
        self.inds = {n: np.empty((0, self.ntemps, self.nwalkers, self.nleaves_max[n]), dtype=bool)
                     for n in self.branch_names}
        self.log_like = np.empty((0, self.ntemps, self.nwalkers))
        self.log_prior = np.empty((0, self.ntemps, self.nwalkers))
        self.betas = np.empty((0, self.ntemps))
        self.move_info = {k: dict(acceptance_fraction=np.zeros((self.ntemps, self.nwalkers)))
                          for k in (self.move_keys or [])}
        self.random_state = None
        self.initialized = True

    @property
    def shape(self):
        return {n: (self.ntemps, self.nwalkers, self.nleaves_max[n], self.ndims[n]) for n in self.branch_names}

    def grow(self, ngrow, blobs=None):
        """backend.py:993-1012.  The new part of every array is left untouched (np.empty, no concatenate with a dummy
        block); the pages of the chain arrays are populated by a helper thread (_prefault_async) ahead of the stores."""
        i = ngrow - (len(self.log_like) - self.iteration)
        if i <= 0:
            return

        def grown(arr, prefault=False):
            new = np.empty((len(arr) + i,) + arr.shape[1:], dtype=arr.dtype)
            new[:len(arr)] = arr
            if prefault:
                _prefault_async(new, len(arr))
            return new
        for n in self.branch_names:
            self.chain[n] = grown(self.chain[n], prefault=True)
            self.inds[n] = grown(self.inds[n])
        self.log_like = grown(self.log_like)
        self.log_prior = grown(self.log_prior)
        self.betas = grown(self.betas)

    def save_step(self, state, accepted, rj_accepted=None, swaps_accepted=None, moves_accepted_fraction=None):
        it = self.iteration
        if it >= len(self.log_like):
            raise ValueError("backend is full: call grow() first")
        for n, br in state.branches.items():
            self.inds[n][it] = br.inds
            c = br.coords.copy()
            c[~br.inds] = self.store_missing_leaves  # backend.py:1053-1059
            self.chain[n][it] = c
        self.log_like[it] = state.log_like
        self.log_prior[it] = state.log_prior
        if state.betas is not None:
            self.betas[it] = state.betas
        self.accepted += np.asarray(accepted).astype(int)
        if swaps_accepted is not None:
            self.swaps_accepted += np.asarray(swaps_accepted).astype(int)
        if rj_accepted is not None and self.rj_accepted is not None:
            self.rj_accepted += np.asarray(rj_accepted).astype(int)
        if moves_accepted_fraction is not None:
            for k, v in moves_accepted_fraction.items():
                self.move_info[k]["acceptance_fraction"][:] = v
        self.random_state = state.random_state
        self.iteration += 1

    def save_arrays(self, coords, inds, log_like, log_prior, betas, accepted, rj_accepted=None, swaps_accepted=None,
                    moves_accepted_fraction=None, random_state=None):
        """save_step (backend.py:1014-1091) for a sample that arrives as plain arrays out of the staging ring: `coords`
        already carry `store_missing_leaves` for inactive leaves (masked on the device, csrc/k_stage.cu)"""
        it = self.iteration
        if it >= len(self.log_like):
            raise ValueError("backend is full: call grow() first")
        for n in self.branch_names:
            _store_copy(self.chain[n][it], coords[n])
            self.inds[n][it] = True if inds is None else inds[n]
        self.log_like[it] = log_like
        self.log_prior[it] = log_prior
        if betas is not None:
            self.betas[it] = betas
        self.accepted += np.asarray(accepted).astype(int)
        if swaps_accepted is not None:
            self.swaps_accepted += np.asarray(swaps_accepted).astype(int)
        if rj_accepted is not None and self.rj_accepted is not None:
            self.rj_accepted += np.asarray(rj_accepted).astype(int)
        if moves_accepted_fraction is not None:
            for k, v in moves_accepted_fraction.items():
                self.move_info[k]["acceptance_fraction"][:] = v
        self.random_state = random_state
        self.iteration += 1

    def _get(self, arr, thin=1, discard=0):
        self._flush()
        return arr[discard + thin - 1:self.iteration:thin]

    def get_chain(self, thin=1, discard=0):
        return {n: self._get(self.chain[n], thin, discard) for n in self.branch_names}

    def get_inds(self, thin=1, discard=0):
        return {n: self._get(self.inds[n], thin, discard) for n in self.branch_names}

    def get_nleaves(self, thin=1, discard=0):
        """backend.py:590-614: number of active leaves per walker"""
        return {n: v.sum(axis=-1) for n, v in self.get_inds(thin, discard).items()}

    def get_log_like(self, thin=1, discard=0):
        return self._get(self.log_like, thin, discard)

    def get_log_prior(self, thin=1, discard=0):
        return self._get(self.log_prior, thin, discard)

    def get_betas(self, thin=1, discard=0):
        return self._get(self.betas, thin, discard)

    def get_last_sample(self):
        self._flush()
        if (not self.initialized) or self.iteration <= 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        it = self.iteration - 1
        coords = {n: self.chain[n][it].copy() for n in self.branch_names}
        inds = {n: self.inds[n][it].copy() for n in self.branch_names}
        return State(coords, inds=inds, log_like=self.log_like[it].copy(), log_prior=self.log_prior[it].copy(),
                     betas=self.betas[it].copy(), random_state=self.random_state)

    @property
    def acceptance_fraction(self):
        self._flush()
        return self.accepted / float(max(self.iteration, 1))
