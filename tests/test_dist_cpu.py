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
