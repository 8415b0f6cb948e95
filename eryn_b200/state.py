"""Host-side walker containers (bookkeeping stays on the host, as in the reference).

`State` / `Branch` follow eryn.state (state.py:330-562) for what the hot path touches:
coords dict keyed by branch name, inds, log_like, log_prior, betas, random_state.
`BranchSupplemental`, blobs and multi-branch device kernels are outside this build's scope."""
from copy import deepcopy

import numpy as np

__all__ = ["State", "Branch"]


class Branch(object):
    """One branch (model): coords [ntemps, nwalkers, nleaves_max, ndim] + inds [.., nleaves_max]."""

    def __init__(self, coords, inds=None, branch_supplemental=None):
        self.coords = coords
        self.ntemps, self.ntrees, self.nleaves_max, self.ndim = coords.shape
        self.shape = coords.shape
        if inds is None:
            self.inds = np.full((self.ntemps, self.ntrees, self.nleaves_max), True)
        elif not isinstance(inds, np.ndarray):
            raise ValueError("inds must be np.ndarray in Branch.")
        elif inds.shape != (self.ntemps, self.ntrees, self.nleaves_max):
            raise ValueError("inds has wrong shape.")
        else:
            self.inds = inds
        if branch_supplemental is not None:
            raise NotImplementedError("branch_supplemental is host bookkeeping outside the device hot path")
        self.branch_supplemental = None

    @property
    def nleaves(self):
        return np.sum(self.inds, axis=-1)


class State(object):
    """The state of the ensemble (state.py:387).  Accepts an ndarray, a dict or another State."""

    def __init__(self, coords, inds=None, branch_supplemental=None, supplemental=None, log_like=None,
                 log_prior=None, betas=None, blobs=None, random_state=None, copy=False):
        dc = deepcopy if copy else (lambda x: x)
        if hasattr(coords, "branches"):
            self.branches = dc(coords.branches)
            self.log_like = dc(coords.log_like)
            self.log_prior = dc(coords.log_prior)
            self.blobs = dc(getattr(coords, "blobs", None))
            self.betas = dc(coords.betas)
            self.supplemental = dc(getattr(coords, "supplemental", None))
            self.random_state = dc(coords.random_state)
            return
        if isinstance(coords, np.ndarray):
            coords = {"model_0": coords}
        elif not isinstance(coords, dict):
            raise ValueError("Input coords need to be np.ndarray, dict, or State object.")
        coords = dict(coords)
        for name in coords:
            if coords[name].ndim == 2:
                coords[name] = coords[name][None, :, None, :]
            if coords[name].ndim == 3:
                coords[name] = coords[name][:, :, None, :]
            elif coords[name].ndim < 2 or coords[name].ndim > 4:
                raise ValueError("Dimension off coordinates must be between 2 and 4.")
        if inds is None:
            inds = {key: None for key in coords}
        elif not isinstance(inds, dict):
            raise ValueError("inds must be None or dict.")
        if branch_supplemental is not None or supplemental is not None or blobs is not None:
            raise NotImplementedError("supplementals / blobs are host bookkeeping outside the device hot path")
        self.branches = {key: Branch(dc(c), inds=inds[key]) for key, c in coords.items()}
        self.log_like = dc(np.atleast_2d(log_like)) if log_like is not None else None
        self.log_prior = dc(np.atleast_2d(log_prior)) if log_prior is not None else None
        self.blobs = None
        self.betas = dc(np.atleast_1d(betas)) if betas is not None else None
        self.supplemental = None
        self.random_state = dc(random_state)

    @property
    def branches_inds(self):
        return {name: b.inds for name, b in self.branches.items()}

    @property
    def branches_coords(self):
        return {name: b.coords for name, b in self.branches.items()}

    @property
    def branches_supplemental(self):
        return {name: None for name in self.branches}

    @property
    def branch_names(self):
        return list(self.branches.keys())

    def get_log_posterior(self, temper=False):
        betas = self.betas if temper else np.ones_like(self.betas)
        return betas[:, None] * self.log_like + self.log_prior
