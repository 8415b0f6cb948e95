"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/eryn_b200.h declares, the
ctypes struct layouts match the C ones, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "eryn_b200.h")).read()
    return sorted(set(re.findall(r"EB_API\s+[\w\s\*]+?\b(eb_\w+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from eryn_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding table and header disagree"
    for n in names:
        assert hasattr(lib, n), f"{n} not exported"
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (eb_\w+)", out)))
    assert exported == names, "library exports differ from the header"


def test_struct_layouts_match():
    from eryn_b200 import _lib
    lib = _lib.load()
    assert lib.eb_abi_version() == 8
    for i, st in enumerate(_lib.STRUCTS):
        assert lib.eb_struct_size(i) == ctypes.sizeof(st), st.__name__
    assert lib.eb_ctrl_size() == ctypes.sizeof(_lib.eb_ctrl)


def test_no_cpu_fallback():
    """Without a GPU the product path must raise, not fall back (run only where no device is visible)."""
    from eryn_b200 import _lib
    lib = _lib.load()
    if lib.eb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from eryn_b200.device import DeviceContext
    from eryn_b200.likelihood import GaussianLikelihood
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    pri = ProbDistContainer({i: uniform_dist(-1.0, 1.0) for i in range(2)})
    with pytest.raises(_lib.ErynB200Error):
        DeviceContext(pri, GaussianLikelihood(np.zeros(2), np.eye(2)))
    job = _lib.eb_host_job()
    assert lib.eb_run_host(ctypes.byref(job), 1) != 0


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under eryn_b200/ may import it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "eryn_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
