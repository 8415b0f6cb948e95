"""Host-side logic of the temperature-sharded run, exercised with torch.distributed (gloo, world_size 2) on CPU.

The kernels are not involved: every rank runs the ORACLE's move on its temperatures with the production
(philox) streams keyed by global temperature, exchanges logl through `eryn_b200.dist` helpers, resolves the
whole ladder redundantly and pulls the rows of its rungs — the algorithm of DESIGN.md §6 — and the result must
equal the unsharded oracle chain bit for bit.  tests/test_mgpu.py checks the CUDA implementation of the same
scheme on 2 GPUs.
"""
import os
import socket

import numpy as np
import pytest

from oracle import eryn_oracle as orc


def test_temperature_partition():
    from eryn_b200.dist import owner_of, temperature_partition
    assert temperature_partition(16, 8) == [0, 2, 4, 6, 8, 10, 12, 14, 16]
    assert temperature_partition(16, 1) == [0, 16]
    assert temperature_partition(5, 2) == [0, 3, 5]
    assert temperature_partition(7, 3) == [0, 3, 5, 7]
    tb = temperature_partition(32, 8)
    assert [owner_of(tb, t) for t in (0, 3, 4, 31)] == [0, 0, 1, 7]
    with pytest.raises(ValueError):
        temperature_partition(4, 8)
    with pytest.raises(ValueError):
        temperature_partition(0, 1)


def test_arena_layout_is_deterministic_and_disjoint():
    from eryn_b200.dist import ALIGN, ArenaLayout, temperature_partition
    tb = temperature_partition(5, 2)
    for rank in range(2):
        a, b = ArenaLayout(tb, rank, 100, 1, 8), ArenaLayout(tb, rank, 100, 1, 8)
        assert a.__dict__ == b.__dict__
        offs = a.coords + a.logl + a.logp + a.logl_all + [a.betas_all, a.flags] + a.logl_ll + a.mail + [a.total]
        assert offs == sorted(offs) and len(set(offs)) == len(offs)
        assert all(o % ALIGN == 0 for o in offs)
        assert a.coords[1] - a.coords[0] >= a.Tg * 100 * 8 * 8
        assert a.logl_all[1] - a.logl_all[0] >= 5 * 100 * 8
        # fused publish: one 16-byte self-validating unit per (temperature, walker) of the FULL ladder, per parity;
        # row mail: [direction][walker chain][L*D + 1] units per parity
        assert a.logl_ll[1] - a.logl_ll[0] >= 5 * 100 * 16
        assert a.mail[1] - a.mail[0] >= 2 * 100 * (8 + 1) * 16 and a.total - a.mail[1] >= 2 * 100 * (8 + 1) * 16
    assert ArenaLayout(tb, 0, 100, 1, 8).Tg == 3 and ArenaLayout(tb, 1, 100, 1, 8).Tg == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_oracle_worker(rank, world, port, T, W, d, nit, seed, out_dir):
    import torch.distributed as dist
    from eryn_b200.dist import exchange, gather_rows, owner_of, scatter_rows, temperature_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tb = temperature_partition(T, world)
        t_lo, t_hi = tb[rank], tb[rank + 1]
        assert exchange(dict(rank=rank, tb=tb)) == [dict(rank=g, tb=tb) for g in range(world)]
        A = np.random.RandomState(99).randn(d, d)
        like = orc.GaussianLike(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
        prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
        x0 = np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d))
        betas = orc.make_ladder_default(d, T)
        st = orc.OState(scatter_rows(x0, tb, rank))
        st.logp = orc.box_log_prior(prior, st.coords, st.inds)
        st.logl = orc.log_like(like, st.coords, st.inds, st.logp)
        streams = orc.PhiloxStreams(seed, t0=t_lo)          # local temperature 0 = global t_lo
        full_streams = orc.PhiloxStreams(seed)              # the ladder is resolved on global indices
        time = 0
        for it in range(nit):
            # 1. the move, local temperatures only
            lists = streams.split_lists(it, t_hi - t_lo, W)
            for split in (0, 1):
                sub, comp = lists[split], lists[1 - split]
                rint, u_z, u_acc = streams.stretch(it, split, t_hi - t_lo, sub.shape[1], comp.shape[1], sub)
                orc.stretch_half_step(st, sub, comp, rint, u_z, u_acc, 2.0, betas[t_lo:t_hi], prior, like)
            # 2. all-gather of logl
            logl_all = gather_rows(st.logl, tb)
            # 3. every rank resolves the whole ladder, then pulls the rows of its rungs from their owners
            iperms, i1perms, us = full_streams.swap_draws(it, T, W, True)
            src, swaps = orc.resolve_ladder(logl_all, betas, iperms, i1perms, us)
            coords_by_rank = exchange(st.coords)
            logp_by_rank = exchange(st.logp)
            new_c, new_lp, new_ll = st.coords.copy(), st.logp.copy(), st.logl.copy()
            for t in range(t_lo, t_hi):
                for w in range(W):
                    s_t, s_w = divmod(int(src[t, w]), W)
                    g = owner_of(tb, s_t)
                    new_c[t - t_lo, w] = coords_by_rank[g][s_t - tb[g], s_w]
                    new_lp[t - t_lo, w] = logp_by_rank[g][s_t - tb[g], s_w]
                    new_ll[t - t_lo, w] = logl_all[s_t, s_w]
            st.coords, st.logp, st.logl = new_c, new_lp, new_ll
            betas = orc.adapt_temps(betas, swaps, W, time)
            time += 1
        full = dict(coords=gather_rows(st.coords, tb), logl=gather_rows(st.logl, tb), logp=gather_rows(st.logp, tb),
                    betas=betas, swaps=swaps)
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), **full)
        with pytest.raises(ValueError):
            gather_rows(np.zeros((t_hi - t_lo + 1, 2)), tb)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("T,world", [(4, 2), (5, 2)])
def test_sharded_oracle_equals_unsharded_gloo(tmp_path, T, world):
    import torch.multiprocessing as mp
    W, d, nit, seed = 24, 4, 6, 4242
    mp.spawn(_sharded_oracle_worker, args=(world, _free_port(), T, W, d, nit, seed, str(tmp_path)), nprocs=world,
             join=True)
    got = np.load(tmp_path / "sharded.npz")
    A = np.random.RandomState(99).randn(d, d)
    like = orc.GaussianLike(np.zeros(d), np.linalg.inv(A @ A.T / d + np.eye(d)))
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    smp = orc.OracleSampler(prior, like, [dict(kind="stretch", a=2.0)], [1.0], orc.PhiloxStreams(seed),
                            betas=orc.make_ladder_default(d, T))
    st = smp.initialise(orc.OState(np.random.RandomState(1).uniform(-3, 3, size=(T, W, 1, d))))
    for _ in range(nit):
        smp.iterate(st)
    assert np.array_equal(got["coords"], st.coords)
    assert np.array_equal(got["logl"], st.logl)
    assert np.array_equal(got["logp"], st.logp)
    assert np.array_equal(got["betas"], smp.betas)
    assert np.array_equal(got["swaps"], smp.swaps_accepted)


def _swap_source(sel, r, T):
    """k_swap.cu:swap_source — rung whose walker ends on rung r (sel[i] = swap accepted at rung i)"""
    if r >= 1 and sel[r]:
        return r - 1
    o = r
    while o + 1 < T and sel[o + 1]:
        o += 1
    return o


def test_row_mail_routing_senders_and_receivers_agree():
    """The sharded swap pass pushes rows that change rank (eb_shard.mail_*): the rank owning the SOURCE rung sends, the
    rank owning the destination rung polls its mailbox slot [direction] — both derive everything from the accept bits.
    Restated here for random bit patterns and partitions (walkers carried across several ranks included): every owned
    rung with a remote source is served by exactly one mail, addressed to the right rank and direction."""
    from eryn_b200.dist import owner_of, temperature_partition
    rng = np.random.RandomState(11)
    for _ in range(2000):
        world = int(rng.choice([2, 3, 4, 8]))
        T = int(rng.randint(world, 40))
        tb = temperature_partition(T, world)
        sel = np.zeros(T + 1, bool)
        sel[1:T] = rng.rand(T - 1) < rng.choice([0.2, 0.5, 0.9, 1.0])
        sent = {}  # (dest rank, direction) -> source rung
        for g in range(world):
            t_lo, t_hi = tb[g], tb[g + 1]
            if t_hi < T and sel[t_hi]:  # up: my top walker moves to rung t_hi
                key = (owner_of(tb, t_hi), 0)
                assert key not in sent
                sent[key] = t_hi - 1
            if t_lo >= 1 and sel[t_lo]:  # down: the walker carried across my lower boundary, if it started here
                o = t_lo
                while o + 1 < T and sel[o + 1]:
                    o += 1
                if o < t_hi:
                    d = t_lo - 1
                    while d >= 1 and sel[d]:
                        d -= 1
                    key = (owner_of(tb, d), 1)
                    assert key not in sent
                    sent[key] = o
        expected = {}
        for g in range(world):
            t_lo, t_hi = tb[g], tb[g + 1]
            dest = [-1, -1]
            if t_lo >= 1 and sel[t_lo]:
                dest[0] = t_lo
            if t_hi < T and sel[t_hi]:
                d = t_hi - 1
                while d >= 1 and sel[d]:
                    d -= 1
                if d >= t_lo:
                    dest[1] = d
            remote = {r: _swap_source(sel, r, T) for r in range(t_lo, t_hi) if owner_of(tb, _swap_source(sel, r, T)) != g}
            assert sorted(remote) == sorted(x for x in dest if x >= 0)
            for direction, r in enumerate(dest):
                if r >= 0:
                    assert (remote[r] == r - 1) == (direction == 0)
                    expected[(g, direction)] = remote[r]
        assert sent == expected


def test_chain_split_pass_protocol_model():
    """NumPy model of the chain-split sharded pass (csrc/k_swap_split.cu, phases A-D) with the kernel's buffer indexing:
    every rank sends logl of its own rungs to the chain's resolver (slot [r][c // world] on rank c % world), the resolver
    cascades and broadcasts accept bits, rows move locally or as mail (coords, logp, logl), partial counts are summed
    over the ranks.  The gathered result must equal the sequential reference ladder (oracle temperature_swaps) on the
    unsharded state — counts included — for uneven partitions and walkers carried across several ranks."""
    from eryn_b200.dist import owner_of, temperature_partition
    from oracle import philox_np as px
    rng = np.random.RandomState(5)
    for world, T, W, hot in ((2, 5, 12, 0.0), (3, 9, 10, 1.0), (4, 16, 9, 30.0), (8, 24, 16, 100.0), (1, 6, 8, 1.0)):
        D, it, seed = 3, 7, 99
        tb = temperature_partition(T, world)
        coords = rng.randn(T, W, 1, D)
        logl = rng.randn(T, W) * 3.0
        logp = rng.randn(T, W)
        betas = np.sort(rng.rand(T))[::-1].copy()
        betas[0] = 1.0
        betas[T // 2:] *= 1.0 / (1.0 + hot)  # a hot end where nearly every swap is accepted (long runs)
        sig = [px.swap_perm(it, seed, r, W) for r in range(T)]
        us = [None] + [px.swap_uniforms(it, seed, i, W) for i in range(1, T)]
        ref = orc.OState(coords.copy(), logl=logl.copy(), logp=logp.copy())
        ref_counts = orc.temperature_swaps(ref, betas, [None] + sig[1:], [None] + sig[:-1], us)

        Wr = (W + world - 1) // world
        llc = [np.full((T, Wr), np.nan) for _ in range(world)]
        bits = [np.zeros((W, T), bool) for _ in range(world)]
        got_bits = [np.zeros(W, bool) for _ in range(world)]
        partial = np.zeros((world, T - 1), int)
        mail = [[dict(), dict()] for _ in range(world)]  # [rank][direction][chain] -> (row, logp, logl)
        # A: own rungs of every chain -> the chain's resolver
        for g in range(world):
            for r in range(tb[g], tb[g + 1]):
                for c in range(W):
                    llc[c % world][r, c // world] = logl[r, sig[r][c]]
        # B: resolvers
        for h in range(world):
            for c in range(h, W, world):
                ll = llc[h][:, c // world]
                assert not np.isnan(ll).any()
                sel = np.zeros(T, bool)
                carry = ll[T - 1]
                for i in range(T - 1, 0, -1):
                    if (betas[i - 1] - betas[i]) * (carry - ll[i - 1]) > np.log(us[i][c]):
                        sel[i] = True
                    else:
                        carry = ll[i - 1]
                partial[h] += sel[1:]
                for g in range(world):
                    bits[g][c] = sel
                    got_bits[g][c] = True
        assert all(b.all() for b in got_bits)
        # C: mail pushes, then rows
        new = [dict(coords=np.full((tb[g + 1] - tb[g], W, 1, D), np.nan), logl=np.full((tb[g + 1] - tb[g], W), np.nan),
                    logp=np.full((tb[g + 1] - tb[g], W), np.nan)) for g in range(world)]
        for g in range(world):
            t_lo, t_hi = tb[g], tb[g + 1]
            for c in range(W):
                sel = bits[g][c]
                if t_hi < T and sel[t_hi]:
                    s = t_hi - 1
                    mail[owner_of(tb, t_hi)][0][c] = (coords[s, sig[s][c]], logp[s, sig[s][c]], logl[s, sig[s][c]])
                if t_lo >= 1 and sel[t_lo]:
                    o = t_lo
                    while o + 1 < T and sel[o + 1]:
                        o += 1
                    if o < t_hi:
                        d = t_lo - 1
                        while d >= 1 and sel[d]:
                            d -= 1
                        mail[owner_of(tb, d)][1][c] = (coords[o, sig[o][c]], logp[o, sig[o][c]], logl[o, sig[o][c]])
        for g in range(world):
            t_lo, t_hi = tb[g], tb[g + 1]
            for c in range(W):
                sel = np.append(bits[g][c], False)
                dest = [-1, -1]
                if t_lo >= 1 and sel[t_lo]:
                    dest[0] = t_lo
                if t_hi < T and sel[t_hi]:
                    d = t_hi - 1
                    while d >= 1 and sel[d]:
                        d -= 1
                    if d >= t_lo:
                        dest[1] = d
                for r in range(t_lo, t_hi):
                    s = _swap_source(sel, r, T)
                    if t_lo <= s < t_hi:
                        row = (coords[s, sig[s][c]], logp[s, sig[s][c]], logl[s, sig[s][c]])
                    else:
                        row = mail[g][dest.index(r)][c]
                    new[g]["coords"][r - t_lo, sig[r][c]], new[g]["logp"][r - t_lo, sig[r][c]] = row[0], row[1]
                    new[g]["logl"][r - t_lo, sig[r][c]] = row[2]
        out = {k: np.concatenate([n[k] for n in new], axis=0) for k in ("coords", "logl", "logp")}
        assert np.array_equal(out["coords"], ref.coords) and np.array_equal(out["logl"], ref.logl)
        assert np.array_equal(out["logp"], ref.logp)
        assert np.array_equal(partial.sum(0), ref_counts.astype(int))  # D: the counts exchange
