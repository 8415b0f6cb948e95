"""Temperature-sharded runs over the GPUs of one box: one process per GPU (torch.distributed for the plumbing).

Why by temperature: the moves couple walkers only inside one temperature
(red_blue.py:183-197 gathers along the walker axis), so a rank that owns whole temperatures runs its
moves with no communication at all; the only coupling between temperatures is the swap ladder
(tempering.py:484-561), whose decisions depend on logl alone (:538).  Per iteration a rank

  1. runs the move kernel on its temperatures                                     (local),
  2. publishes its logl rows into every rank's `logl_all` buffer with NVLink peer stores and raises its
     flag word on every rank                                                      (eb_publish_logl),
  3. waits (in-kernel, on local memory) for all flags, resolves the WHOLE ladder redundantly — every rank
     gets bit-identical swap counts and the same adapted ladder — and writes its own rungs into its
     alternate buffers, pulling each source row from the current buffers of the rank that owns it
     through peer-mapped pointers                                                 (eb_pt_swap_sharded),
  4. flips current/alternate.

There is no NCCL call on the data path (`comm="p2p"`); `comm="nccl"` replaces step 2 by an
`all_gather_into_tensor` of logl (the baseline the peer-store version is measured against).  The random
streams are keyed by GLOBAL temperature and chain index, so a sharded run reproduces the single-GPU chain
bit for bit (tests/test_mgpu.py).

Host-side pieces that do not need a GPU (partition, arena layout, object exchange, state scatter/gather)
are exercised with the gloo backend, world_size 2, in tests/test_dist_cpu.py.
"""
import ctypes as C

import numpy as np

__all__ = ["temperature_partition", "owner_of", "ArenaLayout", "exchange", "scatter_rows", "gather_rows",
           "ShardedRun", "ShardedTemperatureControl", "run_sharded_bench", "sharded_parity_check"]

ALIGN = 256


# ------------------------------------------------------------------------------------------------------
# host logic (no GPU needed)
# ------------------------------------------------------------------------------------------------------
def temperature_partition(ntemps, world):
    """Contiguous, balanced split of the ladder: temp_begin[g] .. temp_begin[g+1] belong to rank g.
    Every rank must own at least one temperature (otherwise shard by replicas instead)."""
    ntemps, world = int(ntemps), int(world)
    if world < 1 or ntemps < 1:
        raise ValueError("ntemps and world must be >= 1")
    if world > ntemps:
        raise ValueError(f"cannot shard {ntemps} temperatures over {world} ranks: every rank needs >= 1 temperature")
    base, extra = divmod(ntemps, world)
    tb = [0]
    for g in range(world):
        tb.append(tb[-1] + base + (1 if g < extra else 0))
    return tb


def owner_of(temp_begin, t):
    """rank that owns global temperature t"""
    g = 0
    while g + 2 < len(temp_begin) and t >= temp_begin[g + 1]:
        g += 1
    return g


class ArenaLayout(object):
    """Byte offsets of one rank's IPC-shared arena.  Deterministic in (ntemps_total, nwalkers, nleaves, ndim, rank,
    world), so every rank can address every other rank's buffers from the arena base pointer alone."""

    def __init__(self, temp_begin, rank, nwalkers, nleaves, ndim):
        T = temp_begin[-1]
        Tg = temp_begin[rank + 1] - temp_begin[rank]
        self.Tg, self.T, self.W, self.L, self.D = Tg, T, int(nwalkers), int(nleaves), int(ndim)
        o = 0

        def take(nbytes):
            nonlocal o
            at = o
            o += (int(nbytes) + ALIGN - 1) // ALIGN * ALIGN
            return at

        n = Tg * self.W
        self.coords = [take(n * self.L * self.D * 8) for _ in range(2)]
        self.logl = [take(n * 8) for _ in range(2)]
        self.logp = [take(n * 8) for _ in range(2)]
        self.logl_all = [take(T * self.W * 8) for _ in range(2)]
        self.betas_all = take(T * 8)
        self.flags = take(16 * 8)
        # fused publish: logl of every rank as self-validating 16-byte units (eb_shard.pub_ll / ll_in), per parity
        self.logl_ll = [take(T * self.W * 16) for _ in range(2)]
        # row mail: [direction][walker chain][L*D + 1] units per parity (eb_shard.mail_peer / mail_in)
        self.mail = [take(2 * self.W * (self.L * self.D + 1) * 16) for _ in range(2)]
        # chain-split pass (comm="split", eb_split; EXPERIMENTAL): units per parity — logl of the chains this rank resolves
        # [T][ceil(W / world)], accept bits of every chain [W][2], partial swap counts [world][T], row mail with logl
        world = len(temp_begin) - 1
        self.Wr = (self.W + world - 1) // world
        self.llc = [take(T * self.Wr * 16) for _ in range(2)]
        self.bits = [take(self.W * 2 * 16) for _ in range(2)]
        self.cnt = [take(world * T * 16) for _ in range(2)]
        self.mail2 = [take(2 * self.W * (self.L * self.D + 2) * 16) for _ in range(2)]
        self.total = o


def exchange(obj, group=None):
    """all-gather of a picklable object: list indexed by rank (torch.distributed, any backend)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def scatter_rows(full, temp_begin, rank):
    """this rank's temperature rows of a full [T, ...] host array (a copy)"""
    return np.ascontiguousarray(full[temp_begin[rank]:temp_begin[rank + 1]])


def gather_rows(local, temp_begin, group=None):
    """all ranks' [T_g, ...] host arrays -> the full [T, ...] array on every rank"""
    parts = exchange(np.ascontiguousarray(local), group)
    for g, p in enumerate(parts):
        if p.shape[0] != temp_begin[g + 1] - temp_begin[g]:
            raise ValueError(f"rank {g} sent {p.shape[0]} rows, its partition has {temp_begin[g + 1] - temp_begin[g]}")
    return np.concatenate(parts, axis=0)


# ------------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------------
class _RawCuda(object):
    """a raw device pointer dressed up for torch.as_tensor (no ownership)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr,
                                             data=(int(ptr), False), version=2, strides=None)


def _tensor_at(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)


class ShardedRun(object):
    """One rank of a temperature-sharded run: the shared arena, the peer mappings and the two
    (current, alternate) DeviceStates."""

    def __init__(self, ctx, ntemps, nwalkers, nleaves=1, group=None, comm="auto", mail=True):
        import torch
        import torch.distributed as dist
        from . import _lib
        from .device import DeviceState
        if comm not in ("auto", "fused", "split", "p2p", "nccl"):
            raise ValueError("comm must be 'auto', 'fused', 'split', 'p2p' or 'nccl'")
        if comm == "auto":
            # two ranks: every rank resolves the whole ladder (two NVLink hops, the redundant half of the ladder is cheap);
            # more ranks: every rank resolves 1/N of the chains (three hops, but work and inbound volume per rank do not
            # grow with N) — measured crossover between N = 2 and N = 4 (profiles/README.md)
            comm = "fused" if dist.get_world_size(group) <= 2 else "split"
        if ctx.rng != "philox":
            raise ValueError("temperature-sharded runs use the philox streams (replay mode is single-GPU)")
        self.ctx, self.group, self.comm = ctx, group, comm
        self.lib = ctx.lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.EB_MAX_RANKS:
            raise ValueError(f"at most {_lib.EB_MAX_RANKS} ranks")
        self.temp_begin = temperature_partition(ntemps, self.world)
        self.T, self.W, self.L, self.D = int(ntemps), int(nwalkers), int(nleaves), ctx.ndim
        self.t_lo, self.t_hi = self.temp_begin[self.rank], self.temp_begin[self.rank + 1]
        if comm == "nccl" and len({self.temp_begin[g + 1] - self.temp_begin[g] for g in range(self.world)}) != 1:
            raise ValueError("comm='nccl' needs ntemps divisible by the number of ranks (all_gather_into_tensor)")
        self.layouts = [ArenaLayout(self.temp_begin, g, self.W, self.L, self.D) for g in range(self.world)]
        lay = self.layouts[self.rank]
        base = C.c_void_p()
        _lib.check(self.lib.eb_dev_malloc(lay.total, C.byref(base)), "eb_dev_malloc")
        self.base = base.value
        self._peer_open = []
        handle = (C.c_uint8 * _lib.EB_IPC_HANDLE_BYTES)()
        _lib.check(self.lib.eb_ipc_export(C.c_void_p(self.base), handle), "eb_ipc_export")
        handles = exchange(bytes(handle), group)
        self.bases = []
        for g, h in enumerate(handles):
            if g == self.rank:
                self.bases.append(self.base)
                continue
            buf = (C.c_uint8 * _lib.EB_IPC_HANDLE_BYTES).from_buffer_copy(h)
            p = C.c_void_p()
            _lib.check(self.lib.eb_ipc_open(buf, C.byref(p)), f"eb_ipc_open(rank {g})")
            self._peer_open.append(p.value)
            self.bases.append(p.value)
        dev = ctx.device
        Tg = self.t_hi - self.t_lo
        self.betas_all = _tensor_at(self.base + lay.betas_all, (self.T,), "<f8", dev)
        self.flags = _tensor_at(self.base + lay.flags, (16,), "<i8", dev)
        self.logl_all = [_tensor_at(self.base + lay.logl_all[p], (self.T, self.W), "<f8", dev) for p in range(2)]
        self.states = []
        for p in range(2):
            coords = _tensor_at(self.base + lay.coords[p], (Tg, self.W, self.L, self.D), "<f8", dev)
            logl = _tensor_at(self.base + lay.logl[p], (Tg, self.W), "<f8", dev)
            logp = _tensor_at(self.base + lay.logp[p], (Tg, self.W), "<f8", dev)
            self.states.append(DeviceState(coords, logl, logp, None, self.betas_all[self.t_lo:self.t_hi],
                                           ctx.branch_name, temp_offset=self.t_lo))
        self.parity = 0
        # per-parity C descriptions
        self._shard, self._pub = [], []
        for p in range(2):
            sh = _lib.eb_shard()
            sh.rank, sh.world, sh.ntemps_total = self.rank, self.world, self.T
            pb = _lib.eb_publish()
            pb.rank, pb.world, pb.ntemps_total, pb.nwalkers = self.rank, self.world, self.T, self.W
            for g in range(self.world + 1):
                sh.temp_begin[g] = self.temp_begin[g]
                pb.temp_begin[g] = self.temp_begin[g]
            for g in range(self.world):
                lg = self.layouts[g]
                sh.coords_src[g] = self.bases[g] + lg.coords[p]
                sh.logp_src[g] = self.bases[g] + lg.logp[p]
                sh.inds_src[g] = None
                pb.logl_all_peer[g] = self.bases[g] + lg.logl_all[p]
                pb.flags_peer[g] = self.bases[g] + lg.flags
            sh.logl_all = self.base + lay.logl_all[p]
            sh.betas_all = self.base + lay.betas_all
            sh.flags = (self.base + lay.flags) if comm == "p2p" else None
            if comm == "fused":  # the swap kernel publishes itself (eb_shard.pub_*): self-validating units, no flags
                sh.flags = None
                sh.pub_src = self.base + lay.logl[p]
                sh.ll_in = self.base + lay.logl_ll[p]
                for g in range(self.world):
                    sh.pub_ll[g] = self.bases[g] + self.layouts[g].logl_ll[p]
                if mail:
                    sh.mail_in = self.base + lay.mail[p]
                    for g in range(self.world):
                        sh.mail_peer[g] = self.bases[g] + self.layouts[g].mail[p]
            pb.logl_local = self.base + lay.logl[p]
            self._shard.append(sh)
            self._pub.append(pb)
        self._split = []
        if comm == "split":  # EXPERIMENTAL: every rank resolves 1/world of the chains (csrc/k_swap_split.cu)
            for p in range(2):
                sp = _lib.eb_split()
                sp.rank, sp.world, sp.ntemps_total = self.rank, self.world, self.T
                for g in range(self.world + 1):
                    sp.temp_begin[g] = self.temp_begin[g]
                sp.coords_cur = self.base + lay.coords[p]
                sp.logl_cur = self.base + lay.logl[p]
                sp.logp_cur = self.base + lay.logp[p]
                sp.betas_all = self.base + lay.betas_all
                sp.llc_in, sp.bits_in = self.base + lay.llc[p], self.base + lay.bits[p]
                sp.cnt_in, sp.mail_in = self.base + lay.cnt[p], self.base + lay.mail2[p]
                for g in range(self.world):
                    lg = self.layouts[g]
                    sp.llc_peer[g], sp.bits_peer[g] = self.bases[g] + lg.llc[p], self.bases[g] + lg.bits[p]
                    sp.cnt_peer[g], sp.mail_peer[g] = self.bases[g] + lg.cnt[p], self.bases[g] + lg.mail2[p]
                self._split.append(sp)
        ctx.write_ctrl(iter=0)  # the flag words count publish CTAs since iteration 0
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every arena is mapped and zeroed (flags = 0) before anyone publishes

    # ---- state ----------------------------------------------------------------------------------------
    @property
    def current(self):
        return self.states[self.parity]

    def load(self, coords_full, betas_full, logl_full=None, logp_full=None):
        """take this rank's temperatures of a full host state [T,W,L,D]; evaluate logl/logp if not given"""
        import torch
        d = self.current
        d.coords.copy_(torch.from_numpy(scatter_rows(np.asarray(coords_full, dtype=np.float64).reshape(
            self.T, self.W, self.L, self.D), self.temp_begin, self.rank)))
        self.betas_all.copy_(torch.from_numpy(np.asarray(betas_full, dtype=np.float64)))
        if logl_full is None or logp_full is None:
            self.ctx.eval_state(d)
        else:
            d.logl.copy_(torch.from_numpy(scatter_rows(np.asarray(logl_full), self.temp_begin, self.rank)))
            d.logp.copy_(torch.from_numpy(scatter_rows(np.asarray(logp_full), self.temp_begin, self.rank)))
        return d

    def gather(self):
        """the full host state on every rank: dict(coords, logl, logp, betas)"""
        d = self.current
        self.check()
        return dict(coords=gather_rows(d.coords.cpu().numpy(), self.temp_begin, self.group),
                    logl=gather_rows(d.logl.cpu().numpy(), self.temp_begin, self.group),
                    logp=gather_rows(d.logp.cpu().numpy(), self.temp_begin, self.group),
                    betas=self.betas_all.cpu().numpy())

    def check(self):
        err = int(self.ctx.read_ctrl().error)
        if err:
            from . import _lib
            raise _lib.ErynB200Error(f"device error {err} on rank {self.rank}: a peer never published its logl "
                                     "(EB_DEVERR_PEER_TIMEOUT) — a rank died or the ranks are out of step")

    # ---- the swap pass --------------------------------------------------------------------------------
    def swap(self, permute=True, adapt=None):
        """steps 2-4 of the module docstring; returns the new current DeviceState"""
        import torch.distributed as dist
        from . import _lib
        ctx, p = self.ctx, self.parity
        if self.comm == "p2p":
            _lib.check(self.lib.eb_publish_logl(C.byref(self._pub[p]), C.c_void_p(ctx.ctrl.data_ptr()), ctx.stream()),
                       "eb_publish_logl")
            ctx.launches += 1
        elif self.comm == "nccl":
            dist.all_gather_into_tensor(self.logl_all[p], self.states[p].logl, group=self.group)
        r = _lib.eb_swap_rng()
        r.mode, r.permute, r.seed, r.iter_dev = _lib.EB_RNG_PHILOX, int(bool(permute)), ctx.seed, ctx.iter_ptr
        ad = None
        if adapt is not None:
            ad = _lib.eb_adapt(int(adapt["adaptive"]), int(adapt["stop_adaptation"]), float(adapt["adaptation_lag"]),
                               float(adapt["adaptation_time"]))
        dst = self.states[1 - p].c_struct()
        if self.comm == "split":
            _lib.check(self.lib.eb_pt_swap_split(C.byref(self._split[p]), C.byref(dst), C.byref(r),
                                                 C.byref(ad) if ad is not None else None,
                                                 C.c_void_p(ctx.ctrl.data_ptr()), ctx.stream()), "eb_pt_swap_split")
        else:
            _lib.check(self.lib.eb_pt_swap_sharded(C.byref(self._shard[p]), C.byref(dst), C.byref(r),
                                                   C.byref(ad) if ad is not None else None,
                                                   C.c_void_p(ctx.ctrl.data_ptr()), ctx.stream()), "eb_pt_swap_sharded")
        ctx.launches += 1
        self.parity = 1 - p
        return self.states[self.parity]

    def close(self):
        import torch
        import torch.distributed as dist
        if self.base is None:
            return
        torch.cuda.synchronize(self.ctx.device)
        dist.barrier(self.group)  # nobody still reads my arena
        for p in self._peer_open:
            self.lib.eb_ipc_close(C.c_void_p(p))
        self._peer_open = []
        self.states, self.logl_all, self.betas_all, self.flags = [], [], None, None
        self.lib.eb_dev_free(C.c_void_p(self.base))
        self.base = None


def _make_sharded_tc():
    from .moves.tempering import TemperatureControl

    class ShardedTemperatureControl(TemperatureControl):
        """TemperatureControl of a temperature-sharded run: `temper_comps` is the cross-GPU swap pass.
        `.betas` is the FULL ladder (identical on every rank), `.swaps_accepted` the full [T-1] counts."""

        def __init__(self, run, effective_ndim, nwalkers, **kwargs):
            super().__init__(effective_ndim, nwalkers, ntemps=run.T, **kwargs)
            self.run = run
            self.ctx = run.ctx
            self._betas_dev = run.betas_all
            import torch
            self._betas_dev.copy_(torch.from_numpy(self._betas_host))
            run.ctx.write_ctrl(time=self._time0)

        def bind(self, ctx):  # already bound to the run's context
            if ctx is not self.run.ctx:
                raise ValueError("a ShardedTemperatureControl belongs to its ShardedRun's DeviceContext")

        def temper_comps(self, state, adapt=True):
            run = self.run
            if state is not run.current:
                raise ValueError("temper_comps of a sharded run works on the run's current DeviceState")
            ad = None
            if adapt and self.adaptive and run.T > 1:
                ad = dict(adaptive=True, stop_adaptation=self.stop_adaptation, adaptation_lag=self.adaptation_lag,
                          adaptation_time=self.adaptation_time)
            return run.swap(permute=self.permute, adapt=ad)

    return ShardedTemperatureControl


def __getattr__(name):
    if name == "ShardedTemperatureControl":
        return _make_sharded_tc()
    raise AttributeError(name)


# ------------------------------------------------------------------------------------------------------
# parity self-check of a sharded run (bench.py --gpus N prints it; tests/test_mgpu.py checks it against the oracle)
# ------------------------------------------------------------------------------------------------------
def sharded_parity_check(rank, world, local, comm="auto", ntemps=None, nwalkers=256, ndim=8, nit=6, seed=4242,
                         like=None, lo=-10.0, hi=10.0):
    """Run a short temperature-sharded chain over all ranks and the SAME chain unsharded on rank 0's GPU (same seed,
    same counter-based streams keyed by global temperature / chain) and compare every array of the final state.
    Returns dict(ok, max_rel, swaps_equal, accepted_equal, ...) on rank 0, None elsewhere.  Results must be bit-identical
    (max_rel == 0): both runs execute the same arithmetic in the same order."""
    import torch
    import torch.distributed as dist
    from .device import DeviceContext
    from .likelihood import GaussianLikelihood
    from .moves import StretchMove, TemperatureControl
    from .prior import ProbDistContainer, uniform_dist
    from .state import State
    T = 16 * world if ntemps is None else int(ntemps)
    W, d = int(nwalkers), int(ndim)
    dev = torch.device("cuda", local)
    if like is None:
        A = np.random.RandomState(99).randn(d, d)
        like = GaussianLikelihood(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
    pri = ProbDistContainer({i: uniform_dist(lo, hi) for i in range(d)})
    x0 = np.random.RandomState(1).uniform(-3.0, 3.0, size=(T, W, 1, d))
    ctx = DeviceContext(pri, like, device=dev, rng="philox", seed=seed)
    run = ShardedRun(ctx, T, W, comm=comm)
    tc = _make_sharded_tc()(run, d, W)
    mv = StretchMove(a=2.0)
    mv.temperature_control = tc
    mv.bind(ctx)
    mv.accepted = np.zeros((run.t_hi - run.t_lo, W))
    run.load(x0, tc._betas_host)
    for _ in range(nit):
        mv.propose(None, run.current)
    full = run.gather()
    swaps = tc.swaps_accepted
    acc = gather_rows(mv.accepted, run.temp_begin)
    comm_used = run.comm
    run.close()
    out = None
    if rank == 0:
        ctx1 = DeviceContext(pri, like, device=dev, rng="philox", seed=seed)
        tc1 = TemperatureControl(d, W, ntemps=T)
        tc1.bind(ctx1)
        mv1 = StretchMove(a=2.0)
        mv1.temperature_control = tc1
        mv1.bind(ctx1)
        mv1.accepted = np.zeros((T, W))
        ds = ctx1.upload(State(x0), betas=tc1.betas_dev)
        ctx1.eval_state(ds)
        for _ in range(nit):
            mv1.propose(None, ds)
        ref = ctx1.download(ds)

        def rel(a, b):
            a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
            den = np.maximum(np.abs(b), 1e-300)
            return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
        max_rel = max(rel(full["coords"], ref.branches_coords[ds.branch_name]), rel(full["logl"], ref.log_like),
                      rel(full["logp"], ref.log_prior), rel(full["betas"], ref.betas))
        swaps_equal = bool(np.array_equal(swaps, tc1.swaps_accepted))
        accepted_equal = bool(np.array_equal(acc, mv1.accepted))
        out = dict(ok=bool(max_rel <= 1e-10 and swaps_equal and accepted_equal), max_rel=max_rel, swaps_equal=swaps_equal,
                   accepted_equal=accepted_equal, ntemps=T, nwalkers=W, ndim=d, iterations=nit, comm=comm_used,
                   against="the same chain run unsharded on rank 0's GPU (coords, logl, logp, betas, swap counts, accept counts)",
                   swaps_accepted_head=[int(v) for v in swaps[:4]])
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1)
# ------------------------------------------------------------------------------------------------------
def run_sharded_bench(args, wl, rank, world, local, clock_sampler_cls=None, comm="auto", with_e2e=True, with_k1=True,
                      sustain_s=0.0):
    """One workload of bench.py on all ranks: the ladder sharded by temperature, one captured graph per (move kind, buffer
    parity) replayed per step.  Per-step CUDA events with an L2 flush between steps (outside the event pairs), summed,
    MAX over ranks.  Returns the JSON dict on rank 0 (None elsewhere)."""
    import time
    import torch
    import torch.distributed as dist
    from .device import DeviceContext
    from .moves import GaussianMove, StretchMove
    from .prior import ProbDistContainer, uniform_dist

    T, W, d = wl["T"], wl["W"], wl["d"]
    dev = torch.device("cuda", local)
    pri = ProbDistContainer({i: uniform_dist(wl["lo"], wl["hi"]) for i in range(d)})
    ctx = DeviceContext(pri, wl["device_like"], device=dev, rng="philox", seed=20261017)
    run = ShardedRun(ctx, T, W, comm=comm)
    STC = _make_sharded_tc()
    tc = STC(run, d, W)
    moves = []
    for m in wl["moves"]:
        mv = StretchMove(a=m["a"]) if m["kind"] == "stretch" else GaussianMove({"model_0": m["proposal"]["scale"] ** 2})
        mv.temperature_control = tc
        mv.bind(ctx)
        mv.accepted = np.zeros((run.t_hi - run.t_lo, W))
        moves.append(mv)
    x0 = wl["x0"]
    run.load(x0, tc._betas_host)
    torch.cuda.synchronize()
    dist.barrier()

    sched_rng = np.random.RandomState(7)  # same schedule on every rank
    nmoves = len(moves)
    schedule = sched_rng.choice(nmoves, p=np.asarray(wl["weights"]) / np.sum(wl["weights"]), size=args.warmup + args.steps)
    stream = torch.cuda.Stream(device=dev)

    # one iteration = move + (publish +) swap; the buffers flip every iteration, so a graph is captured per
    # (move kind, parity) and replayed according to the schedule
    graphs = {}
    use_graph = comm != "nccl"
    with torch.cuda.stream(stream):
        for mv in moves:  # warm-up outside capture, an even number of iterations (parity returns to 0)
            for _ in range(2):
                mv.propose(None, run.current)
        torch.cuda.synchronize()
        dist.barrier()
        if use_graph:
            for par in range(2):
                for mi, mv in enumerate(moves):
                    run.parity = par
                    g = torch.cuda.CUDAGraph()
                    l0 = ctx.launches
                    with torch.cuda.graph(g, stream=stream):
                        mv.propose(None, run.current)
                    graphs[(mi, par)] = (g, ctx.launches - l0)
            run.parity = 0
    torch.cuda.synchronize()
    dist.barrier()

    def one_step(i):
        mi = int(schedule[i])
        if use_graph:
            g, nl = graphs[(mi, run.parity)]
            g.replay()
            run.parity ^= 1
            return nl
        l0 = ctx.launches
        moves[mi].propose(None, run.current)
        return ctx.launches - l0

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clk = clock_sampler_cls(local) if clock_sampler_cls is not None else None
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            one_step(i)
        torch.cuda.synchronize()
        dist.barrier()
        if clk is not None:
            clk.start()
        launches = 0
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            evs[i][0].record(stream)
            launches += one_step(args.warmup + i)
            evs[i][1].record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        # back to back without flush
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            one_step(args.warmup + i)
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        if sustain_s > 0:  # keep the GPUs under this load long enough for the clock sampler (same count on every rank)
            per = max(e0.elapsed_time(e1) / args.steps * 1e-3, 1e-6)
            tt = torch.tensor([per], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            n_sus = int(sustain_s / float(tt[0])) // 2 * 2
            for i in range(n_sus):
                one_step(args.warmup + (i % args.steps))
            torch.cuda.synchronize()
            dist.barrier()
    run.check()
    # the dominant kernel alone (this rank's temperatures): back-to-back stretch steps in one graph
    st_moves = [m for m in moves if isinstance(m, StretchMove)]
    k1_us = None
    if st_moves and with_k1:
        nrep = 50
        cnt = st_moves[0]._count_buffer(ctx, run.t_hi - run.t_lo, W)
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(gk, stream=stream):
                for _ in range(nrep):
                    ctx.stretch_step(run.current, 2.0, accepted_count=cnt)
            gk.replay()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            gk.replay()
            k1.record(stream)
            torch.cuda.synchronize()
        k1_us = k0.elapsed_time(k1) * 1e3 / nrep
    clocks = clk.summary() if clk is not None else None
    total_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    resident_ms = e0.elapsed_time(e1)
    tt = torch.tensor([total_ms, resident_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # max over ranks
    total_ms, resident_ms = float(tt[0]), float(tt[1]) / args.steps

    # ---- e2e: host buffers in, one iteration, host buffers out, every step (per rank: its own temperatures) ----
    e2e = None
    if with_e2e:
        cur = run.current
        h = dict(coords=cur.coords.cpu().pin_memory(), logl=cur.logl.cpu().pin_memory(), logp=cur.logp.cpu().pin_memory())
        ne2e = max(20, min(args.steps, 100))
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.stream(stream):
            t0 = time.perf_counter()
            for i in range(ne2e):
                cur = run.current
                cur.coords.copy_(h["coords"], non_blocking=True)
                cur.logl.copy_(h["logl"], non_blocking=True)
                cur.logp.copy_(h["logp"], non_blocking=True)
                one_step(args.warmup + (i % args.steps))
                cur = run.current
                h["coords"].copy_(cur.coords, non_blocking=True)
                h["logl"].copy_(cur.logl, non_blocking=True)
                h["logp"].copy_(cur.logp, non_blocking=True)
                stream.synchronize()
            e2e_s = (time.perf_counter() - t0) / ne2e
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        run.check()
        shard_bytes = sum(t.numel() * 8 for t in h.values())
        e2e = dict(value=T * W / e2e_s, ms_per_step=e2e_s * 1e3, h2d_bytes_per_step=int(shard_bytes * world),
                   d2h_bytes_per_step=int(shard_bytes * world),
                   api="per rank: pinned host shard -> H2D -> move + sharded swap pass -> D2H, every step")
    swaps = tc.swaps_accepted.tolist()[:4]
    betas = run.betas_all.cpu().numpy()
    out = None
    if rank == 0:
        out = dict(value=T * W * args.steps / (total_ms * 1e-3), ms_per_step=total_ms / args.steps,
                   resident_ms=resident_ms, launches=int(launches), clocks=clocks, e2e=e2e,
                   swaps=swaps, betas=[float(betas[0]), float(betas[-1])], temp_begin=run.temp_begin, comm=comm,
                   graph=use_graph, k1_us=k1_us, local_temps=run.t_hi - run.t_lo)
        out["comm"] = run.comm
    run.close()
    del flush
    return out
