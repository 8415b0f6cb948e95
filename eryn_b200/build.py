"""Build `eryn_b200/lib/liberyn_b200.so` (sm_100a) with nvcc.  Cross-compiles without a GPU.

    python -m eryn_b200.build [--force]

The library is a plain C-ABI shared object (include/eryn_b200.h); it is built in-tree so that it
travels with the repository snapshot to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "liberyn_b200.so")
# (source, extra defines, object name): the two move kernels are compiled once per likelihood kind so the template
# instantiations build in parallel
SOURCES = [("abi_core.cu", [], "abi_core.o"), ("k_swap.cu", [], "k_swap.o"), ("k_shard.cu", [], "k_shard.o"), ("k_swap_split.cu", [], "k_swap_split.o"), ("k_rj.cu", [], "k_rj.o"),
           ("host_job.cu", [], "host_job.o"), ("k_stage.cu", [], "k_stage.o"), ("k_mt.cu", [], "k_mt.o")]
for _k in range(3):
    SOURCES.append(("k_stretch.cu", [f"-DEB_ONLY_LIKE={_k}"], f"k_stretch_{_k}.o"))
    SOURCES.append(("k_gauss.cu", [f"-DEB_ONLY_LIKE={_k}"], f"k_gauss_{_k}.o"))
HEADERS = ["common.cuh", "likelihoods.cuh", "rng.cuh", "stretch_lanes.cuh", "resident.cuh", os.path.join("..", "..", "include", "eryn_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: eryn_b200 needs the CUDA toolkit to build its kernels")


def _digest():
    h = hashlib.sha256()
    for f in sorted({x[0] for x in SOURCES}) + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(SOURCES).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "liberyn_b200.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(item):
        src, defs, objname = item
        obj = os.path.join(OBJDIR, objname)
        cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} {defs}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
