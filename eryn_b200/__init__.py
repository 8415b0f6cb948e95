"""eryn_b200 — B200-native walker-parallel sampling hot path with Eryn's EnsembleSampler / Move / State surface.

Importing the package does not touch CUDA; the first compute call loads `lib/liberyn_b200.so` and
raises if it is missing or no GPU is visible (there is no CPU fallback)."""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: `import eryn_b200` must work on a machine without torch/CUDA (symbol checks, docs)
    if name in ("EnsembleSampler", "walkers_independent"):
        from . import ensemble
        return getattr(ensemble, name)
    if name in ("State", "Branch"):
        from . import state
        return getattr(state, name)
    if name == "Backend":
        from .backend import Backend
        return Backend
    if name in ("DeviceContext", "DeviceState"):
        from . import device
        return getattr(device, name)
    raise AttributeError(name)
