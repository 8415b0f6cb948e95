"""ORACLE (test infrastructure only) — counter-based random streams in NumPy.

This file restates, on the CPU, the *production-mode* random streams of the
CUDA path (`eryn_b200/csrc/rng.cuh`).  It is imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs; nothing in the
product package imports it.

The reference (`/root/reference/src/eryn`) draws its randoms from two NumPy
MT19937 streams (`ensemble.py:651-652`, `red_blue.py:124`,
`tempering.py:526-535`).  Those cannot be generated on the device, so the
device path has two modes: *replay* (host NumPy draws, reference order,
bit-parity with the reference) and *philox* (this file's streams).  The
algorithmic use of each draw is identical in both modes and is restated in
`oracle/eryn_oracle.py`.

Contents
  philox4x32_10      Random123 / cuRAND Philox4x32-10 block function
  u01_52             two 32-bit words -> double in the open interval (0, 1)
  feistel_perm       keyed bijection on [0, n) (4-round Feistel + cycle walk)
  draw_*             the named streams (purpose tags) the kernels consume
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)

# purpose tags (upper 8 bits of counter word 3) — must match rng.cuh
TAG_SPLIT_KEY = 1
TAG_STRETCH = 2
TAG_GAUSS = 3
TAG_ACCEPT = 4
TAG_SWAP_KEY = 5
TAG_SWAP_U = 6
TAG_RJ = 7
TAG_MT = 9   # (8 = the group move's tag, oracle/rj_oracle.py)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10.  All arguments broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01_52(lo, hi):
    """(lo, hi) uint32 -> double in (0,1): ((hi:lo) >> 12) + 0.5) * 2^-52."""
    x = (np.asarray(hi, dtype=np.uint64) << np.uint64(32)) | np.asarray(lo, dtype=np.uint64)
    return ((x >> np.uint64(12)).astype(np.float64) + 0.5) * (2.0 ** -52)


def _ctr3(tag, it):
    it = int(it)
    return np.uint32(((tag & 0xFF) << 24) | ((it >> 32) & 0xFFFFFF)), np.uint32(it & 0xFFFFFFFF)


def _stream(tag, it, seed, c0, c1):
    c3, c2 = _ctr3(tag, it)
    seed = int(seed)
    return philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def _fmix32(h):
    h = np.asarray(h, dtype=np.uint32)
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint32(16))
        h = h * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(13))
        h = h * np.uint32(0xC2B2AE35)
        h = h ^ (h >> np.uint32(16))
    return h


FEISTEL_ROUNDS = 4


def feistel_keys(tag, it, seed, idx):
    """Four round keys (one Philox block) for the bijection number `idx` of stream `tag` at iteration `it`."""
    w0, w1, w2, w3 = _stream(tag, it, seed, np.uint32(idx), np.uint32(0))
    return [np.uint32(w) for w in (w0, w1, w2, w3)]


def half_bits(n):
    bits = max(int(n - 1).bit_length(), 1)
    return (bits + 1) // 2


def feistel_perm(x, n, keys):
    """Keyed bijection of [0, n) applied elementwise to the uint32 array x."""
    x = np.array(x, dtype=np.uint32, copy=True)
    if n <= 1:
        return np.zeros_like(x)
    hb = np.uint32(half_bits(n))
    mask = np.uint32((1 << int(hb)) - 1)
    todo = np.ones(x.shape, dtype=bool)
    while np.any(todo):
        v = x[todo]
        L = v >> hb
        R = v & mask
        for r in range(8 if int(hb) <= 3 else FEISTEL_ROUNDS):  # small domains: eight rounds (rng.cuh)
            k = np.uint32((int(keys[r & 3]) + 0x9E3779B9 * (r >> 2)) & 0xFFFFFFFF)
            L, R = R, L ^ (_fmix32(R ^ k) & mask)
        v = (L << hb) | R
        x[todo] = v
        todo = x >= np.uint32(n)
    return x


# ----------------------------------------------------------------------------------
# named streams
# ----------------------------------------------------------------------------------
def split_perm(it, seed, t, W):
    """sigma_t : [0,W) -> [0,W), the red/blue assignment of temperature t at iteration it."""
    keys = feistel_keys(TAG_SPLIT_KEY, it, seed, t)
    return feistel_perm(np.arange(W, dtype=np.uint32), W, keys).astype(np.int64)


def split_draw(lo, hi, n):
    """(integer, fraction) of x*n/2^64 for the 64-bit uniform x = hi:lo — rng.cuh split_draw."""
    lo = np.asarray(lo, dtype=np.uint64)
    hi = np.asarray(hi, dtype=np.uint64)
    n = int(n)
    # x*n as a 96-bit product from 32-bit halves (uint64 arithmetic wraps, which is what we want for the low word)
    with np.errstate(over="ignore"):
        p_lo = lo * np.uint64(n)                      # < 2^64 (n < 2^32)
        p_hi = hi * np.uint64(n)                      # weight 2^32
        mid = (p_lo >> np.uint64(32)) + (p_hi & _MASK32)
        ipart = (p_hi >> np.uint64(32)) + (mid >> np.uint64(32))
        f = ((mid & _MASK32) << np.uint64(32)) | (p_lo & _MASK32)
    frac = ((f >> np.uint64(12)).astype(np.float64) + 0.5) * (2.0 ** -52)
    return ipart.astype(np.int64), frac


def stretch_draws(it, seed, t_global, pos, Nc):
    """(rint, u_z, u_acc) of the walkers at split positions `pos` of global temperature `t_global`;
    one Philox block per walker, counter (pos, t_global)."""
    r0, r1, r2, r3 = _stream(TAG_STRETCH, it, seed, np.asarray(pos, dtype=np.uint32),
                             np.asarray(t_global, dtype=np.uint32))
    rint, u_z = split_draw(r0, r1, Nc)
    return rint, u_z, u01_52(r2, r3)


def accept_draws(it, seed, flat_walker, slot):
    """u_acc for flat walker ids (t*W+w); counter (flat, slot)."""
    r0, r1, _, _ = _stream(TAG_ACCEPT, it, seed, np.asarray(flat_walker, dtype=np.uint32), np.uint32(slot))
    return u01_52(r0, r1)


def gauss_draws(it, seed, flat_leaf, D, gidx=0):
    """Standard normals [N, D] for flat leaf ids; counter (flat_leaf, pair j | Gibbs split << 16); Box–Muller."""
    flat_leaf = np.asarray(flat_leaf, dtype=np.uint32)
    npair = (D + 1) // 2
    j = np.arange(npair, dtype=np.uint32)[None, :] | np.uint32(int(gidx) << 16)
    r0, r1, r2, r3 = _stream(TAG_GAUSS, it, seed, flat_leaf[:, None], j)
    u1 = u01_52(r0, r1)
    u2 = u01_52(r2, r3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    z = np.empty((flat_leaf.shape[0], 2 * npair))
    z[:, 0::2] = rad * np.cos(ang)
    z[:, 1::2] = rad * np.sin(ang)
    return z[:, :D]


def swap_perm(it, seed, rung, W):
    keys = feistel_keys(TAG_SWAP_KEY, it, seed, rung)
    return feistel_perm(np.arange(W, dtype=np.uint32), W, keys).astype(np.int64)


def swap_uniforms(it, seed, rung, W):
    """u of pair (chain) k at rung `rung`: one Philox block serves rungs r and r+8 of a chain
    (k_swap.cu: lane r%8 owns rungs r, r+8, ...): counter (k, (r&7) | ((r>>4)<<3)), word pair (r>>3)&1."""
    r = int(rung)
    j = (r & 7) | ((r >> 4) << 3)
    r0, r1, r2, r3 = _stream(TAG_SWAP_U, it, seed, np.arange(W, dtype=np.uint32), np.uint32(j))
    return u01_52(r2, r3) if (r >> 3) & 1 else u01_52(r0, r1)
