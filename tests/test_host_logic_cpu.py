"""Host-side mirror of the reference interface (no GPU needed)."""
import numpy as np
import pytest

from oracle import eryn_oracle as orc
from oracle import philox_np as px
from tests import cases


def test_make_ladder_matches_oracle_and_reference_values():
    from eryn_b200.moves import make_ladder
    for nd, nt in [(3, 4), (8, 16), (20, 32), (150, 5)]:
        np.testing.assert_array_equal(make_ladder(nd, ntemps=nt), orc.make_ladder_default(nd, nt))
    # golden: the reference's ladder after 0 adaptations is betas[0] of the first iteration's parent
    b = make_ladder(3, ntemps=4, Tmax=np.inf)
    assert b[-1] == 0 and len(b) == 4
    with pytest.raises(ValueError):
        make_ladder(0, ntemps=3)
    with pytest.raises(ValueError):
        make_ladder(3)
    with pytest.raises(ValueError):
        make_ladder(3, ntemps=3, Tmax=0.5)


def test_prior_rvs_consumes_global_stream_like_reference():
    from eryn_b200.prior import ProbDistContainer, uniform_dist
    g = cases.load("c2_small")
    np.random.seed(int(g["seed"]))
    pri = ProbDistContainer({i: uniform_dist(float(g["lo"]), float(g["hi"])) for i in range(int(g["ndim"]))})
    x0 = pri.rvs(size=(int(g["ntemps"]), int(g["nwalkers"])))
    np.testing.assert_array_equal(x0, g["x0"])
    lo, hi, lp = pri.arrays()
    np.testing.assert_array_equal(lp, orc.BoxPrior(lo, hi).logpdf_val)
    with pytest.raises(ValueError):
        uniform_dist(1.0, 1.0)
    assert uniform_dist(2.0, -2.0).min_val == -2.0  # swapped like prior.py:29-32


def test_state_shapes():
    from eryn_b200.state import State
    s = State(np.zeros((6, 3)))
    assert s.branches["model_0"].shape == (1, 6, 1, 3)
    s = State(np.zeros((2, 6, 3)))
    assert s.branches["model_0"].shape == (2, 6, 1, 3) and s.branches_inds["model_0"].all()
    with pytest.raises(ValueError):
        State(np.zeros(3))
    s2 = State(s, copy=True)
    s2.branches["model_0"].coords[0, 0, 0, 0] = 1.0
    assert s.branches["model_0"].coords[0, 0, 0, 0] == 0.0


def test_backend_roundtrip():
    from eryn_b200.backend import Backend
    from eryn_b200.state import State
    b = Backend()
    b.reset(4, 2, ntemps=3, branch_names=["model_0"], moves=["StretchMove_0"])
    b.grow(2)
    st = State(np.arange(24.0).reshape(3, 4, 1, 2), log_like=np.ones((3, 4)), log_prior=np.zeros((3, 4)),
               betas=np.array([1.0, 0.5, 0.1]))
    b.save_step(st, np.ones((3, 4)), swaps_accepted=np.array([1.0, 2.0]))
    b.save_step(st, np.ones((3, 4)), swaps_accepted=np.array([1.0, 2.0]))
    assert b.iteration == 2 and b.get_chain()["model_0"].shape == (2, 3, 4, 1, 2)
    assert b.accepted.sum() == 24 and list(b.swaps_accepted) == [2, 4]
    last = b.get_last_sample()
    np.testing.assert_array_equal(last.branches["model_0"].coords, st.branches["model_0"].coords)


def test_philox_oracle_streams():
    # Random123 known-answer vectors for Philox4x32-10
    def h(v):
        return [int(x) for x in v]
    assert h(px.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    m = 0xFFFFFFFF
    assert h(px.philox4x32_10(m, m, m, m, m, m)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(px.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    for n in (1, 2, 3, 7, 32, 99, 4096, 5000):
        p = px.split_perm(5, 77, 2, n)
        assert sorted(p.tolist()) == list(range(n))
    u = px.swap_uniforms(3, 9, 2, 1000)
    assert u.min() > 0.0 and u.max() < 1.0
    z = px.gauss_draws(0, 1, np.arange(20000), 3)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02


def test_oracle_philox_sampler_is_a_valid_sampler():
    """Production-mode streams through the oracle: a unit Gaussian is recovered (mean 0, var 1)."""
    d, W = 2, 64
    prior = orc.BoxPrior(np.full(d, -10.0), np.full(d, 10.0))
    smp = orc.OracleSampler(prior, orc.GaussianLike(np.zeros(d), np.eye(d)), [dict(kind="stretch", a=2.0)], [1.0],
                            orc.PhiloxStreams(123), betas=orc.make_ladder_default(d, 3))
    st = smp.initialise(orc.OState(np.random.RandomState(0).uniform(-1, 1, size=(3, W, d))))
    xs = []
    for it in range(600):
        smp.iterate(st)
        if it >= 100:
            xs.append(st.coords[0, :, 0, :].copy())
    xs = np.concatenate(xs)
    assert np.all(np.abs(xs.mean(0)) < 0.15)
    assert np.all(np.abs(xs.var(0) - 1.0) < 0.2)
    assert smp.swaps_accepted.sum() > 0


def _cascade_sequential(T, ll, lu, dts):
    """tempering.py:515-559 restricted to one chain (the order the reference resolves the ladder in)"""
    sel = np.zeros(T, bool)
    carry = ll[T - 1]
    for i in range(T - 1, 0, -1):
        lower = ll[i - 1]
        if dts[i] * (carry - lower) > lu[i]:
            sel[i] = True
        else:
            carry = lower
    return sel


def _cascade_band_walk(T, ll, lu, dts, ages=8, lanes=32):
    """the resolution of k_swap.cu for long ladders / sharded passes: a band of `ages` tests per walker evaluated ahead,
    a walk over the carried walkers only (j -> j - run - 1) that marks the rung where each settles, runs that outlast
    the band extended `lanes` rungs per step by a vote; accepted rungs = the complement of the marked ones"""
    band = np.zeros(T, int)
    for j in range(1, T):
        for a in range(ages):
            i = j - a
            if i >= 1 and dts[i] * (ll[j] - ll[i - 1]) > lu[i]:
                band[j] |= 1 << a
    rej = np.zeros(T, bool)
    j = T - 1
    while j >= 1:
        run = 0
        while run < ages and (band[j] >> run) & 1:
            run += 1
        if run == ages:
            base = j - ages
            while base >= 1:
                votes = [(base - l >= 1) and bool(dts[base - l] * (ll[j] - ll[base - l - 1]) > lu[base - l]) for l in range(lanes)]
                n = votes.index(False) if False in votes else -1
                if n < 0:
                    run, base = run + lanes, base - lanes
                    continue
                run += n
                break
        rej[j - run] = True
        j -= run + 1
    sel = ~rej
    sel[0] = False
    return sel


def test_band_walk_cascade_equals_sequential_ladder():
    rng = np.random.RandomState(3)
    for _ in range(600):
        T = int(rng.choice([2, 3, 5, 8, 16, 24, 32, 70, 128]))
        ll = rng.randn(T) * rng.choice([0.1, 1, 10])
        lu = np.log(rng.rand(T)) * rng.choice([0.01, 1])
        dts = np.abs(rng.randn(T)) * rng.choice([0.0, 0.01, 1, 100])   # 0.0: the hot end, every swap accepted
        assert np.array_equal(_cascade_sequential(T, ll, lu, dts), _cascade_band_walk(T, ll, lu, dts))


def test_combine_move_surface():
    """CombineMove attribute plumbing (combine.py:32-97): counters, tempering object and periodic info go to the sub-moves"""
    from eryn_b200.moves import CombineMove, GaussianMove, StretchMove, TemperatureControl
    subs = [StretchMove(a=2.0), (GaussianMove({"model_0": 0.25}), 0.3)]   # weights in tuples are ignored (combine.py:16-18)
    mv = CombineMove(subs)
    assert [type(m).__name__ for m in mv.moves] == ["StretchMove", "GaussianMove"]
    mv.accepted = np.zeros((2, 6))
    assert len(mv.accepted) == 2 and all(a.shape == (2, 6) for a in mv.accepted)
    tc = TemperatureControl(3, 6, ntemps=2)
    mv.temperature_control = tc
    assert all(m.temperature_control is tc for m in mv.moves) and mv.ntemps == 2
    mv.periodic = {"model_0": {0: 1.0}}
    assert all(m.periodic == {"model_0": {0: 1.0}} for m in mv.moves)
    for m, n in zip(mv.moves, (4, 2)):
        m.num_proposals = n
        m.accepted = np.full((2, 6), 2.0)
    np.testing.assert_allclose(mv.acceptance_fraction, np.full((2, 6), 0.75))   # mean of 2/4 and 2/2
    assert len(mv.acceptance_fraction_separate) == 2
    with pytest.raises(ValueError):
        CombineMove([])


def test_swap_and_split_bijections_have_uniform_pair_frequencies():
    """The production streams pair walkers through keyed bijections (4-round Feistel + cycle walking, csrc/rng.cuh)
    instead of NumPy permutations.  A 4-round Feistel network on <= 14 bits does not reach every permutation, and MCMC
    validity does not need it: the pairing only has to be independent of the state and fair.  What the sampler relies on
    is checked here as frequencies over many iterations (chi-square against the uniform law, 5-sigma bounds):
      * swap pass: walker a of rung i meets walker b of rung i-1 with probability 1/W for every (a, b)
        (tempering.py:526-527 pairs iperm[k] with i1perm[k]);
      * red/blue split: every walker lands in split 0 with probability Ns0/W, and an ordered pair of walkers is
        (moving walker, one of its possible partners) equally often (red_blue.py:121-124, stretch.py:93)."""
    from oracle import philox_np as px
    for W in (7, 16, 50):
        nit = 400 * W
        pair = np.zeros((W, W))
        in0 = np.zeros(W)
        pos = np.zeros((W, W))
        for it in range(nit):
            s_hi, s_lo = px.swap_perm(it, 12345, 3, W), px.swap_perm(it, 12345, 2, W)
            pair[s_hi, s_lo] += 1          # chain k: slot sigma_3(k) of rung 3 with slot sigma_2(k) of rung 2
            sp = px.split_perm(it, 12345, 1, W)
            in0[sp[0::2]] += 1
            pos[np.arange(W), sp] += 1     # position -> walker
        for counts, cells in ((pair, W * W), (pos, W * W)):
            exp = nit / W
            chi2 = ((counts - exp) ** 2 / exp).sum()
            dof = (W - 1) ** 2             # doubly stochastic table
            assert abs(chi2 - dof) < 5.0 * np.sqrt(2.0 * dof) + 0.05 * dof, (W, chi2, dof)
        n0 = (W + 1) // 2
        exp0 = nit * n0 / W
        z = (in0 - exp0) / np.sqrt(nit * (n0 / W) * (1 - n0 / W))
        assert np.abs(z).max() < 5.0, (W, z)


def test_backend_grow_prefaults_without_touching_stored_samples():
    """Backend.grow populates new chain pages from a helper thread (MADV_POPULATE_WRITE): stores that race with it keep
    their contents, and the helper is optional (errors are swallowed)"""
    import threading
    from eryn_b200 import backend as bk
    from eryn_b200.backend import Backend
    T, W, D, n = 4, 4096, 8, 40          # 1 MB per sample, 40 MB of chain: above the prefault threshold
    b = Backend()
    b.reset(W, D, ntemps=T, branch_names=["model_0"])
    b.grow(n)
    r = np.random.RandomState(0)
    samples = []
    for it in range(n):
        c = r.randn(T, W, 1, D)
        samples.append(c)
        b.save_arrays({"model_0": c}, None, c[..., 0, 0], c[..., 0, 1], np.ones(T), np.zeros((T, W)))
    for t in threading.enumerate():
        if t.name == "eryn_b200-prefault":
            t.join(10.0)
    chain = b.get_chain()["model_0"]
    for it in range(n):
        assert np.array_equal(chain[it], samples[it])
    # growing again keeps what is stored
    b.grow(n)
    assert np.array_equal(b.get_chain()["model_0"][n - 1], samples[-1])
    assert bk._prefault_async(np.empty((3, 4)), 0) is None      # small arrays: nothing to do


def test_lazy_adaptation_protocol_model():
    """NumPy model of the deferred ladder adaptation (csrc/k_swap.cu count publication, common.cuh:lazy_adapt_apply,
    adapt_flush_kernel, DeviceContext.flush_adapt): two sets of count rows alternating with the iteration parity, a
    snapshot of ladder and clock, the pending / applied markers and the writer guard.  Random schedules of deferred and
    self-adapting passes, move kernels that apply (possibly twice per iteration), kernels that need a host flush, and host
    reads must (a) give the ladder, clock and totals of a pass that adapts every time (tempering.py:563-596) and (b) find
    the count rows of every pass zero when it starts."""
    T, W, lag, t0 = 6, 64, 50.0, 10.0

    def adapt(betas, counts, time):   # adapt_temps, the arithmetic of pt_swap_adapt
        b = betas.copy()
        kappa = (lag / (time + lag)) / t0
        ratios = counts / float(W)
        dS = kappa * (ratios[:-1] - ratios[1:])
        dT = np.diff(1.0 / b[:-1]) * np.exp(dS)
        b[1:-1] = b[1:-1] + (1.0 / (np.cumsum(dT) + 1.0 / b[0]) - b[1:-1])
        return b

    class Dev(object):
        def __init__(self, betas):
            self.betas, self.time, self.iter = betas.copy(), 0, 0
            self.work = np.zeros((2, T - 1), dtype=np.int64)
            self.total = np.zeros(T - 1, dtype=np.int64)
            self.last = np.zeros(T - 1, dtype=np.int64)
            self.pending = self.applied = 0
            self.snap = None

        def swap_pass(self, counts, defer):
            it = self.iter
            assert not self.work[it & 1].any(), "count rows of this parity must be zero when a pass starts"
            self.work[it & 1] += counts
            if defer:
                self.snap = (self.betas.copy(), self.time)
                self.pending = it + 1
            else:
                assert self.pending in (0, self.applied), "the host flushes before a pass that adapts itself"
                c = self.work[it & 1].copy()
                self.betas = adapt(self.betas, c, self.time)
                self.work[:] = 0
                self.last, self.total, self.time = c, self.total + c, self.time + 1
            self.iter = it + 1

        def apply(self, it, writer_zero_own=False):
            """what every CTA of a move kernel of iteration `it` (or the flush kernel) computes; returns the ladder it uses"""
            if self.pending == 0 or self.pending != it:
                return self.betas
            row = (it - 1) & 1
            c = self.work[row].copy()
            ladder = adapt(self.snap[0], c, self.snap[1])
            if self.applied != self.pending:            # the writer CTA, once per deferred pass
                self.betas = ladder.copy()
                self.last, self.total, self.time = c, self.total + c, self.snap[1] + 1
                self.work[row ^ 1] = 0
                if writer_zero_own:
                    self.work[row] = 0
                self.applied = self.pending
            return ladder

        def flush(self):
            if self.pending and self.pending != self.applied:
                self.apply(self.pending, writer_zero_own=True)
                self.pending = 0

    rs = np.random.RandomState(11)
    for trial in range(200):
        betas0 = np.geomspace(1.0, 1e-2, T)
        dev = Dev(betas0)
        ref_b, ref_time, ref_total = betas0.copy(), 0, np.zeros(T - 1, dtype=np.int64)
        lazy = bool(rs.randint(2))
        for it in range(rs.randint(1, 30)):
            if rs.rand() < 0.15:                          # the loop changes character: lazy_begin / lazy_end
                if lazy:
                    dev.flush()
                lazy = not lazy
            kind = rs.randint(3)                          # 0 / 1: kernels that apply themselves, 2: needs the host flush
            if kind == 2 or not lazy:
                dev.flush()
                used = dev.betas
            else:
                used = dev.apply(dev.iter)
                if rs.rand() < 0.3:
                    assert np.array_equal(dev.apply(dev.iter), used)   # a second kernel of the same iteration: same ladder
            assert np.array_equal(used, ref_b), "a move kernel must see the ladder adapted by every earlier pass"
            counts = rs.randint(0, W + 1, size=T - 1)
            dev.swap_pass(counts, defer=lazy)
            ref_b, ref_time, ref_total = adapt(ref_b, counts, ref_time), ref_time + 1, ref_total + counts
            if rs.rand() < 0.3:                           # a host read: download / read_ctrl / staging snapshot
                dev.flush()
                assert np.array_equal(dev.betas, ref_b) and dev.time == ref_time and np.array_equal(dev.total, ref_total)
                assert np.array_equal(dev.last, counts)
        dev.flush()
        assert np.array_equal(dev.betas, ref_b) and dev.time == ref_time and np.array_equal(dev.total, ref_total)


def test_wavefront_groups_cover_every_rung_once():
    """the group arithmetic of eb_run_host's wavefront schedule (csrc/host_job.cu:issue_wavefront): temperatures are
    uploaded in G groups hottest first, the rung ranges adjoin from T-1 down to 0, and every rung is downloaded exactly
    once, never before the range that makes it final (tempering.py:515: rung i is final after the swap (i, i-1))"""
    for T in range(2, 40):
        for G in range(1, min(T, 16) + 1):
            b = [T - (T * g) // G for g in range(G + 1)]          # group g = temperatures [b[g+1], b[g])
            assert b[0] == T and b[-1] == 0 and all(b[g] > b[g + 1] for g in range(G))
            uploaded, downloaded, resolved = set(), [], set()
            prev_lo = None
            for g in range(G):
                t_lo, t_hi = b[g + 1], b[g]
                uploaded |= set(range(t_lo, t_hi))
                r_hi, r_lo = (T - 1 if g == 0 else t_hi), t_lo
                assert prev_lo is None or r_hi == prev_lo          # ranges adjoin
                prev_lo = r_lo
                assert set(range(r_lo, r_hi + 1)) <= uploaded      # a range only touches temperatures that have landed
                resolved |= set(range(r_lo + 1, r_hi + 1))         # swaps (r, r-1) for r = r_hi .. r_lo+1
                f_lo = 0 if g == G - 1 else r_lo + 1
                final = list(range(f_lo, r_hi + 1))
                assert all(r in resolved or r == 0 for r in final)
                downloaded += final
            assert sorted(downloaded) == list(range(T)) and resolved == set(range(1, T))
