"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  The CUDA path is driven through the C ABI
(ctypes, include/g2048.h) and compared with the CPU oracle -- bit-exact for boards, flags, scores, aux
state, the float64 shaped reward and (deterministic mode) the float32 Q-table; atomic-mode Q values
within the float32 tolerance stated in the test.  Golden vectors recorded from the reference itself are
replayed as well.  Nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

import oracle
from test_oracle_golden import replay_config2, replay_env

pytestmark = pytest.mark.gpu

Q_ATOL, Q_RTOL = 1e-5, 1e-5   # float32 atomic-sum-order tolerance (1e-6 is typical)


def vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def L():
    import g2048
    g2048.init(0)
    return g2048.lib()


@pytest.fixture(scope="module")
def ctx(L):
    handle = L.g2048_ctx_create(0, 1 << 20, 1 << 22)
    assert handle, L.g2048_last_error()
    yield handle
    L.g2048_ctx_destroy(handle)


def ok(L, rc):
    assert rc == 0, L.g2048_last_error().decode()


def random_boards(rng, n, lmax=15, p_zero=0.3):
    lv = rng.randint(1, lmax + 1, size=(n, 16))
    lv = np.where(rng.random_sample((n, 16)) < p_zero, 0, lv).astype(np.uint64)
    b = np.zeros(n, np.uint64)
    for j in range(16):
        b |= lv[:, j] << np.uint64(4 * j)
    b[b == 0] = 1
    return b


def fresh_envs(n, seed=5, episode=0, base=0):
    boards = np.zeros(n, np.uint64)
    oracle.env_reset(boards, None, None, None, seed=seed, episode_idx=episode, env_id_base=base)
    return boards, np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)


# ------------------------------------------------------------------------------------------ moves
def test_row_lut_all_65536_rows(L, ctx):
    """Every row through move(0) on the GPU == the oracle's restatement of move_left."""
    res, mg = oracle.row_table()
    rows = np.arange(65536, dtype=np.uint64)
    for shift in (0, 16, 32, 48):   # the row in each of the four board rows
        boards = rows << np.uint64(shift)
        boards[0] = 1
        out, moved, sc = np.zeros_like(boards), np.zeros(65536, np.uint8), np.zeros(65536, np.int32)
        ok(L, L.g2048_ctx_move_trial(ctx, vp(boards), vp(np.zeros(65536, np.uint8)), vp(out), vp(moved), vp(sc), 65536))
        want = res.astype(np.uint64) << np.uint64(shift)
        score = np.array([(1 << (m >> 4) if m >> 4 else 0) + (1 << (m & 15) if m & 15 else 0) for m in mg.tolist()], np.int32)
        assert np.array_equal(out[1:], want[1:]) and np.array_equal(sc[1:], score[1:])
        assert np.array_equal(moved[1:] != 0, (want != boards)[1:])


@pytest.mark.parametrize("lmax,p_zero", [(15, 0.3), (15, 0.0), (3, 0.1), (11, 0.6), (2, 0.0)])
def test_moves_and_legal_mask_random_boards(L, ctx, lmax, p_zero):
    rng = np.random.RandomState(lmax * 7 + int(p_zero * 10))
    n = 200_000
    boards = random_boards(rng, n, lmax, p_zero)
    actions = rng.randint(0, 4, size=n).astype(np.uint8)
    out, moved, sc = np.zeros_like(boards), np.zeros(n, np.uint8), np.zeros(n, np.int32)
    ok(L, L.g2048_ctx_move_trial(ctx, vp(boards), vp(actions), vp(out), vp(moved), vp(sc), n))
    wb, wm, ws = oracle.move(boards, actions)
    assert np.array_equal(out, wb) and np.array_equal(moved, wm) and np.array_equal(sc, ws)
    lm = np.zeros(n, np.uint8)
    ok(L, L.g2048_ctx_legal_mask(ctx, vp(boards), vp(lm), n))
    assert np.array_equal(lm, oracle.legal_mask(boards))
    assert np.array_equal((lm == 0) & ((boards != 0)), oracle.dead(boards).astype(bool))


# ------------------------------------------------------------------------------------------ golden replay
def make_ctx_step(L, ctx):
    def step(boards, aux, score, actions, draws, flavour):
        n = len(boards)
        r, f, m, ms = np.zeros(n), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
        ok(L, L.g2048_ctx_env_step(ctx, vp(boards), vp(aux), vp(score), vp(actions), vp(draws), vp(r), vp(f), vp(m),
                                   vp(ms), n, flavour, 0, 0, 0))
        return r, f, m, ms
    return step


@pytest.mark.parametrize("name,flavour", [("env_penalty", 0), ("env_nopenalty", 1)])
def test_env_step_replays_reference_golden(L, ctx, golden, name, flavour):
    """Bit-exact replay of action/spawn trajectories recorded from the reference (BASELINE config 2 form)."""
    replay_env(golden(name), flavour, make_ctx_step(L, ctx))


@pytest.mark.parametrize("name,flavour", [("config2_penalty", 0), ("config2_nopenalty", 1)])
def test_config2_4096_envs_512_steps_replay_reference_recorded_draws(L, ctx, golden, name, flavour):
    """BASELINE config 2 as written: 4,096 envs x 512 steps, the actions and spawn draws the REFERENCE made
    (np.random.seed(1000 + i), actions RandomState(2000 + i)), replayed through g2048_ctx_env_step / _env_reset: every
    env's digest over all its steps' outputs, final board, score and aux equal the reference's."""
    def gpu_step(boards, aux, score, actions, draws, fl):
        n = len(boards)
        reward, flags = np.zeros(n, np.float64), np.zeros(n, np.uint8)
        maxlvl, ms = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        ok(L, L.g2048_ctx_env_step(ctx, vp(boards), vp(aux), vp(score), vp(actions), vp(np.ascontiguousarray(draws)),
                                   vp(reward), vp(flags), vp(maxlvl), vp(ms), n, fl, 0, 0, 0))
        return reward, flags, maxlvl, ms

    def gpu_reset(boards, score, mask, draws):
        ok(L, L.g2048_ctx_env_reset(ctx, vp(boards), vp(score), vp(mask), vp(np.ascontiguousarray(draws)), len(boards), 0, 0, 0))

    steps, resets = replay_config2(golden(name), flavour, gpu_step, gpu_reset)
    assert steps == 4096 * 512 and resets > 4096


@pytest.mark.parametrize("name", ["env_penalty", "env_nopenalty"])
def test_env_reset_replays_reference_golden(L, ctx, golden, name):
    g = golden(name)
    n = len(g["reset_board"])
    boards, score = np.zeros(n, np.uint64), np.ones(n, np.int32)
    ok(L, L.g2048_ctx_env_reset(ctx, vp(boards), vp(score), None, vp(np.ascontiguousarray(g["reset_draws"])), n, 0, 0, 0))
    assert np.array_equal(boards, g["reset_board"]) and not score.any()


@pytest.mark.parametrize("flavour", [0, 1])
def test_env_step_philox_4096_envs_512_steps(L, ctx, flavour):
    """BASELINE config 2 size: 4096 envs x 512 steps, every output of every step against the oracle."""
    n, T, seed, base = 4096, 512, 0xABCDEF, 1 << 33
    rng = np.random.RandomState(flavour)
    gb, ga, gs = fresh_envs(n, seed, 0, base)
    cb, ca, cs = gb.copy(), ga.copy(), gs.copy()
    step = None
    for t in range(T):
        actions = rng.randint(0, 4, size=n).astype(np.uint8)
        r, f, m, ms = np.zeros(n), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
        ok(L, L.g2048_ctx_env_step(ctx, vp(gb), vp(ga), vp(gs), vp(actions), None, vp(r), vp(f), vp(m), vp(ms), n,
                                   flavour, seed, t, base))
        wr, wf, wm, wms = oracle.env_step(cb, ca, cs, actions, None, flavour, seed, t, base)
        assert np.array_equal(gb, cb), t
        assert np.array_equal(f, wf) and np.array_equal(m, wm) and np.array_equal(ms, wms), t
        assert np.array_equal(r.view(np.uint64), wr.view(np.uint64)), t
        assert np.array_equal(gs, cs), t
        if flavour == 0:
            assert np.array_equal(ga, ca), t
        done = ((f >> 2) & 1).astype(np.uint8)
        if done.any():
            ok(L, L.g2048_ctx_env_reset(ctx, vp(gb), vp(gs), vp(done), None, n, seed, t + 1, base))
            oracle.env_reset(cb, cs, done, None, seed, t + 1, base)
            assert np.array_equal(gb, cb)
    assert step is None


# ------------------------------------------------------------------------------------------ fused rollouts
@pytest.mark.parametrize("flavour", [0, 1])
@pytest.mark.parametrize("n,k", [(20000, 300), (1000, 700), (1, 2000)])
def test_rollout_random_matches_oracle(L, ctx, flavour, n, k):
    """Fused K-step rollout (boards in registers, smem LUT for n >= 16384) == oracle, bit for bit."""
    seed, base, t0 = 99 + n, 12345, 7
    gb, ga, gs = fresh_envs(n, seed, 0, base)
    cb, ca, cs = gb.copy(), ga.copy(), gs.copy()
    gc = np.zeros(16, np.int64)
    ok(L, L.g2048_ctx_rollout_random(ctx, vp(gb), vp(ga), vp(gs), n, k, flavour, seed, t0, base, vp(gc)))
    cc = oracle.rollout_random(cb, ca, cs, k, flavour, seed, t0, base, threads=8)
    assert np.array_equal(gb, cb) and np.array_equal(gs, cs)
    if flavour == 0:
        assert np.array_equal(ga, ca)
    assert np.array_equal(gc, cc), (gc, cc)
    assert gc[0] == n * k


def test_rollout_random_is_sharding_invariant_at_1M_envs(L, ctx):
    """BASELINE config 3 size: 2^20 envs; one launch == two half-size launches with shifted env ids."""
    n, k, seed = 1 << 20, 48, 0x2048
    b, a, s = fresh_envs(n, seed)
    b1, a1, s1, c1 = b.copy(), a.copy(), s.copy(), np.zeros(16, np.int64)
    ok(L, L.g2048_ctx_rollout_random(ctx, vp(b1), vp(a1), vp(s1), n, k, 0, seed, 0, 0, vp(c1)))
    h = n // 2
    c2 = np.zeros(16, np.int64)
    for lo in (0, h):
        bb, aa, ss, cc = b[lo:lo + h].copy(), a[lo:lo + h].copy(), s[lo:lo + h].copy(), np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_random(ctx, vp(bb), vp(aa), vp(ss), h, k, 0, seed, 0, lo, vp(cc)))
        assert np.array_equal(bb, b1[lo:lo + h]) and np.array_equal(aa, a1[lo:lo + h]) and np.array_equal(ss, s1[lo:lo + h])
        c2 += cc
        c2[4] = max(c2[4], cc[4]) if lo else cc[4]
    c2[4] = c1[4]
    assert np.array_equal(c1, c2) and c1[0] == n * k
    # tile-sum conservation: sum of tiles == 2 * (#2-spawns) + 4 * (#4-spawns) is not observable here, but every
    # board must stay a legal non-empty position
    assert (b1 != 0).all()
    # a 65,536-env sample of the same launch against the oracle
    m = 1 << 16
    cb, ca, cs = b[:m].copy(), a[:m].copy(), s[:m].copy()
    oracle.rollout_random(cb, ca, cs, k, 0, seed, 0, 0, threads=8)
    assert np.array_equal(cb, b1[:m]) and np.array_equal(ca, a1[:m]) and np.array_equal(cs, s1[:m])


def test_rollout_random_8M_envs_as_one_launch_or_eight_shards(L):
    """BASELINE config 4 size: 2^23 envs (device-pointer API).  One launch == eight launches of 2^20 envs with shifted
    env ids (what eight GPUs run), for both flavours; a sample from the LAST shard is checked against the oracle, so the
    high env ids go through the same Philox counter path on both sides."""
    import torch
    n, G, k, seed = 1 << 23, 8, 24, 0x2048
    h = n // G
    st = torch.cuda.current_stream().cuda_stream
    for flavour in (0, 1):
        def fresh():
            b = torch.zeros(n, dtype=torch.int64, device="cuda")
            a = torch.full((n,), oracle.AUX_INIT, dtype=torch.int64, device="cuda")
            s = torch.zeros(n, dtype=torch.int32, device="cuda")
            ok(L, L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, seed, 0, 0, st))
            return b, a, s
        b1, a1, s1 = fresh()
        start = b1[n - h:n - h + 4096].cpu().numpy().view(np.uint64).copy()
        c1 = torch.zeros(16, dtype=torch.int64, device="cuda")
        ok(L, L.g2048_rollout_random(b1.data_ptr(), a1.data_ptr(), s1.data_ptr(), n, k, flavour, seed, 0, 0, c1.data_ptr(), st))
        b2, a2, s2 = fresh()
        c2 = torch.zeros(16, dtype=torch.int64, device="cuda")
        for r in range(G):
            lo = r * h
            ok(L, L.g2048_rollout_random(b2[lo:].data_ptr(), a2[lo:].data_ptr(), s2[lo:].data_ptr(), h, k, flavour, seed, 0,
                                         lo, c2.data_ptr(), st))
        assert torch.equal(b1, b2) and torch.equal(a1, a2) and torch.equal(s1, s2)
        assert torch.equal(c1, c2) and int(c1[0]) == n * k
        cb, ca, cs = start, np.full(4096, oracle.AUX_INIT, np.uint64), np.zeros(4096, np.int32)
        oracle.rollout_random(cb, ca, cs, k, flavour, seed, 0, n - h, threads=4)
        assert np.array_equal(cb, b1[n - h:n - h + 4096].cpu().numpy().view(np.uint64))
        assert np.array_equal(cs, s1[n - h:n - h + 4096].cpu().numpy())
        if flavour == 0:
            assert np.array_equal(ca, a1[n - h:n - h + 4096].cpu().numpy().view(np.uint64))


# ------------------------------------------------------------------------------------------ Q-table
def export_ctx_table(L, ctx):
    n = L.g2048_ctx_qtable_size(ctx)
    assert n >= 0
    keys, rows = np.zeros(max(n, 1), np.uint64), np.zeros((max(n, 1), 4), np.float32)
    assert L.g2048_ctx_qtable_export(ctx, vp(keys), vp(rows), n) == n
    order = np.argsort(keys[:n])
    return keys[:n][order], rows[:n][order]


def test_qtable_update_n1_teacher_forced_reference_transitions(L, ctx, golden):
    """update_q_value on the reference's own transitions, one at a time (N = 1 == the reference's order):
    float32 bit-exact against the oracle, and within tolerance of the reference's float64 table."""
    g = golden("qlearn_ref")
    n = 3000
    ok(L, L.g2048_ctx_qtable_clear(ctx))
    r32 = g["r"].astype(np.float32)
    tab = oracle.QTable(1 << 15, f32=True)
    for i in range(n):
        sl = slice(i, i + 1)
        ok(L, L.g2048_ctx_qtable_update(ctx, vp(g["s"][sl]), vp(g["a"][sl]), vp(r32[sl]), vp(g["s2"][sl]),
                                        vp(g["done"][sl]), 1, 0.1, 0.99, i % 2))
        tab.update_batch_f32(g["s"][sl], g["a"][sl], r32[sl], g["s2"][sl], g["done"][sl], 0.1, 0.99)
    keys, rows = export_ctx_table(L, ctx)
    wk, wr = tab.export()
    assert np.array_equal(keys, wk)
    assert np.array_equal(rows, wr.astype(np.float32))
    t64 = oracle.QTable(1 << 15, f32=False)
    t64.update_seq_f64(g["s"][:n], g["a"][:n], g["r"][:n], g["s2"][:n], g["done"][:n], 0.1, 0.99)
    k64, r64 = t64.export()
    assert np.array_equal(keys, k64)
    np.testing.assert_allclose(rows, r64, rtol=Q_RTOL, atol=Q_ATOL)


@pytest.mark.parametrize("mode", [0, 1])
def test_qtable_update_batched_synchronous(L, ctx, mode):
    """One synchronous batch with heavy (state, action) collisions.  Deterministic mode: bit-exact against the
    oracle (duplicates applied in ascending index order).  Atomic mode: duplicates are applied in an unspecified
    order, so (a) collision-free batches must agree within float32 rounding, (b) with collisions every Q value must
    lie in the hull of {old value, its targets} (each update is a convex step towards a target)."""
    rng = np.random.RandomState(3)
    pool = random_boards(rng, 3000, 8, 0.3)
    ok(L, L.g2048_ctx_qtable_clear(ctx))
    tab = oracle.QTable(1 << 14, f32=True)
    for rep in range(6):
        n = 100_000
        s, s2 = pool[rng.randint(0, len(pool), n)], pool[rng.randint(0, len(pool), n)]
        a = rng.randint(0, 4, n).astype(np.uint8)
        r = (rng.standard_normal(n) * 3).astype(np.float32)
        d = (rng.random_sample(n) < 0.05).astype(np.uint8)
        if mode == 0:
            before = dict(zip(*[x.tolist() if i == 0 else list(x) for i, x in enumerate(tab.export())]))
            # bootstrap values on the snapshot -> per-(s,a) target hull
            k0, r0 = tab.export()
            lut = {int(k): row for k, row in zip(k0, r0)}
            best = np.array([lut[int(x)].max() if int(x) in lut else 0.0 for x in s2], np.float32)
            target = r + np.where(d != 0, np.float32(0), np.float32(0.99) * best)
        ok(L, L.g2048_ctx_qtable_update(ctx, vp(s), vp(a), vp(r), vp(s2), vp(d), n, 0.1, 0.99, mode))
        tab.update_batch_f32(s, a, r, s2, d, 0.1, 0.99)
        keys, rows = export_ctx_table(L, ctx)
        wk, wr = tab.export()
        assert np.array_equal(keys, wk)
        if mode == 1:
            assert np.array_equal(rows, wr.astype(np.float32)), rep
        else:
            pos = {int(k): i for i, k in enumerate(keys)}
            lo = np.array([[lut[int(k)][c] if int(k) in lut else 0.0 for c in range(4)] for k in keys], np.float64)
            hi = lo.copy()
            idx = np.array([pos[int(x)] for x in s])
            np.minimum.at(lo, (idx, a), target)
            np.maximum.at(hi, (idx, a), target)
            assert (rows >= lo - 1e-4).all() and (rows <= hi + 1e-4).all()
            ok(L, L.g2048_ctx_qtable_clear(ctx))   # restart both from the same (empty) table
            tab = oracle.QTable(1 << 14, f32=True)
    # collision-free batch: atomic == deterministic == oracle up to float32 rounding of one update
    ok(L, L.g2048_ctx_qtable_clear(ctx))
    tab = oracle.QTable(1 << 14, f32=True)
    s = pool[:2000].copy(); s2 = pool[1000:3000].copy()
    s, uniq = np.unique(s, return_index=True); s2 = s2[uniq]
    n = len(s)
    a = rng.randint(0, 4, n).astype(np.uint8); r = rng.standard_normal(n).astype(np.float32); d = np.zeros(n, np.uint8)
    for _ in range(3):
        ok(L, L.g2048_ctx_qtable_update(ctx, vp(s), vp(a), vp(r), vp(s2), vp(d), n, 0.1, 0.99, mode))
        tab.update_batch_f32(s, a, r, s2, d, 0.1, 0.99)
    keys, rows = export_ctx_table(L, ctx)
    wk, wr = tab.export()
    assert np.array_equal(keys, wk)
    np.testing.assert_allclose(rows, wr, rtol=Q_RTOL, atol=Q_ATOL)


def test_qtable_lookup_and_choose_action(L, ctx):
    rng = np.random.RandomState(11)
    pool = random_boards(rng, 5000, 10, 0.3)
    ok(L, L.g2048_ctx_qtable_clear(ctx))
    tab = oracle.QTable(1 << 14, f32=True)
    n = 50_000
    s, s2 = pool[rng.randint(0, len(pool), n)], pool[rng.randint(0, len(pool), n)]
    a, r = rng.randint(0, 4, n).astype(np.uint8), rng.standard_normal(n).astype(np.float32)
    d = np.zeros(n, np.uint8)
    ok(L, L.g2048_ctx_qtable_update(ctx, vp(s), vp(a), vp(r), vp(s2), vp(d), n, 0.1, 0.9, 1))
    tab.update_batch_f32(s, a, r, s2, d, 0.1, 0.9)
    probe = np.concatenate([pool[:2000], random_boards(rng, 2000, 12, 0.5)])
    rows, found = np.zeros((len(probe), 4), np.float32), np.zeros(len(probe), np.uint8)
    ok(L, L.g2048_ctx_qtable_lookup(ctx, vp(probe), len(probe), vp(rows), vp(found), 0))
    for i in range(0, len(probe), 97):
        wr, wf = tab.get(probe[i])
        assert bool(found[i]) == wf and np.array_equal(rows[i], wr.astype(np.float32))
    size0 = L.g2048_ctx_qtable_size(ctx)
    assert size0 == len(tab)
    for eps in (0.0, 0.3, 1.0):
        acts = np.zeros(len(probe), np.uint8)
        ok(L, L.g2048_ctx_choose_action(ctx, vp(probe), vp(acts), len(probe), eps, 77, 5, 1000))
        want = tab.choose_action(probe, oracle.eps_threshold(eps), 77, 5, 1000)
        assert np.array_equal(acts, want)
    assert L.g2048_ctx_qtable_size(ctx) == len(tab)   # choose_action inserts like the defaultdict


def test_rollout_qlearn_single_env_is_the_reference_order(L, ctx):
    """Fused asynchronous rollout with N = 1 == sequential Q-learning (main.py:91-101) in float32, bit for bit."""
    for flavour, eps in ((0, 0.2), (1, 0.5)):
        seed, base, k = 31337, 3, 4000
        ok(L, L.g2048_ctx_qtable_clear(ctx))
        gb, ga, gs = fresh_envs(1, seed, 0, base)
        cb, ca, cs = gb.copy(), ga.copy(), gs.copy()
        gc = np.zeros(16, np.int64)
        tab = oracle.QTable(1 << 15, f32=True)
        for part in range(2):   # two launches: the carried state must survive the boundary
            ok(L, L.g2048_ctx_rollout_qlearn(ctx, vp(gb), vp(ga), vp(gs), 1, k, flavour, 0.1, 0.99, eps, seed, part * k,
                                             base, vp(gc)))
        cc = oracle.rollout_qlearn_seq(cb, ca, cs, tab, 2 * k, 0.1, 0.99, eps, flavour, seed, 0, base)
        assert np.array_equal(gb, cb) and np.array_equal(ga, ca) and np.array_equal(gs, cs)
        keys, rows = export_ctx_table(L, ctx)
        wk, wr = tab.export()
        assert np.array_equal(keys, wk)
        assert np.array_equal(rows, wr.astype(np.float32))
        assert gc[0] == k and cc[0] == 2 * k


def test_rollout_qlearn_1M_envs_properties(L, ctx):
    """BASELINE config 3 size (2^20 envs): step count, no dropped inserts, table size == inserts, EVERY update applied
    (LOST == 0 while thousands of envs update the same start states concurrently: RETRIED > 0), finite Q bounded by
    |r|max / (1 - gamma) (the update is a contraction, not a sum of stale deltas), boards stay legal."""
    n, k, seed = 1 << 20, 48, 0x2048
    big = L.g2048_ctx_create(0, n, 1 << 27)
    assert big, L.g2048_last_error()
    try:
        b, a, s = fresh_envs(n, seed)
        c = np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_qlearn(big, vp(b), vp(a), vp(s), n, k, 0, 0.1, 0.99, 0.1, seed, 0, 0, vp(c)))
        assert c[0] == n * k and c[7] == 0
        assert c[8] == 0 and c[9] > 0          # LOST, RETRIED (include/g2048.h)
        size = L.g2048_ctx_qtable_size(big)
        assert size == c[6]
        keys, rows = np.zeros(size, np.uint64), np.zeros((size, 4), np.float32)
        assert L.g2048_ctx_qtable_export(big, vp(keys), vp(rows), size) == size
        assert len(np.unique(keys)) == size and (keys != 0).all()
        assert np.isfinite(rows).all() and np.abs(rows).max() <= 20 / (1 - 0.99)
        assert (b != 0).all()
    finally:
        L.g2048_ctx_destroy(big)


def test_fused_rollout_all_exploring_is_the_random_rollout_and_stores_every_visited_state(L):
    """BASELINE config 3 size, epsilon = 1: the actions do not depend on Q, so the fused Q-learning rollout must leave
    boards / aux / score bit-identical to the oracle's random-policy rollout (same Philox action x3 >> 30), and --
    defaultdict semantics, main.py:16 -- its table must hold exactly the states the oracle visits: the reset boards
    and the board after each of the 16 steps (no game ends that early), nothing more, nothing twice."""
    n, k, seed = 1 << 20, 16, 0x2048
    big = L.g2048_ctx_create(0, n, 1 << 26)
    assert big, L.g2048_last_error()
    try:
        b, a, s = fresh_envs(n, seed)
        cb, ca, cs = b.copy(), a.copy(), s.copy()
        c = np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_qlearn(big, vp(b), vp(a), vp(s), n, k, 0, 0.1, 0.99, 1.0, seed, 0, 0, vp(c)))
        visited = [cb.copy()]
        total = np.zeros(oracle.N_COUNTERS, np.int64)
        for t in range(k):
            ct = oracle.rollout_random(cb, ca, cs, 1, 0, seed, t, 0, threads=8)
            mx = max(total[4], ct[4])
            total += ct
            total[4] = mx                                       # the max level is a maximum, not a sum
            visited.append(cb.copy())
        assert total[2] == 0                                   # no episode ended: every visited state is in `visited`
        assert np.array_equal(b, cb) and np.array_equal(a, ca) and np.array_equal(s, cs)
        assert np.array_equal(c[:6], total[:6])               # steps, valid, episodes, score, max level, reward checksum
        assert c[7] == 0 and c[8] == 0
        want = np.unique(np.concatenate(visited))
        size = L.g2048_ctx_qtable_size(big)
        assert size == len(want) == c[6]
        keys, rows = np.zeros(size, np.uint64), np.zeros((size, 4), np.float32)
        assert L.g2048_ctx_qtable_export(big, vp(keys), vp(rows), size) == size
        assert np.array_equal(np.sort(keys), want)
        assert np.isfinite(rows).all()
    finally:
        L.g2048_ctx_destroy(big)


def late_game_boards(rng, n):
    """n boards with 7 tiles of levels 5..12 on random cells: nine cells stay empty (no game ends within a few moves)
    and two such boards -- or any of their successors -- coincide with negligible probability."""
    b = np.zeros(n, np.uint64)
    for i in range(n):
        cells = rng.choice(16, 7, replace=False)
        lv = rng.randint(5, 13, size=7)
        v = 0
        for cell, l in zip(cells, lv):
            v |= int(l) << (4 * int(cell))
        b[i] = v
    return b


def test_fused_rollout_collision_free_envs_are_their_own_sequential_runs(L):
    """20,000 envs teleported to different late-game boards: an env that shares no state with any other env for the
    next steps must come out of the asynchronous fused rollout (big-launch path: 128-bit insert + update, deferred
    list, grouped apply) exactly as out of sequential Q-learning on its own (main.py:91-101) -- boards, aux, score and
    every Q row it touched, bit for bit.  The oracle's trajectories say which envs those are (all but a handful)."""
    n, k, seed = 20000, 6, 99
    big = L.g2048_ctx_create(0, n, 1 << 22)
    assert big, L.g2048_last_error()
    try:
        b = late_game_boards(np.random.RandomState(5), n)
        assert len(np.unique(b)) == n
        a, s = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
        cb, ca, cs = b.copy(), a.copy(), s.copy()
        c = np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_qlearn(big, vp(b), vp(a), vp(s), n, k, 0, 0.1, 0.99, 0.3, seed, 0, 0, vp(c)))
        assert c[0] == n * k and c[7] == 0 and c[8] == 0
        tab = oracle.QTable(1 << 20, f32=True)
        visited = [cb.copy()]
        for t in range(k):
            cc = oracle.rollout_qlearn_seq(cb, ca, cs, tab, 1, 0.1, 0.99, 0.3, 0, seed, t, 0)
            assert cc[2] == 0                                  # no game ended: no common reset boards
            visited.append(cb.copy())
        # states seen by more than one env; an env that ever stood on one is not on its own any more
        per_env = np.stack(visited, 1)                         # [n][k + 1]
        pairs = np.unique(np.stack([per_env.ravel(), np.repeat(np.arange(n, dtype=np.uint64), k + 1)], 1), axis=0)
        keys_seen, owners = np.unique(pairs[:, 0], return_counts=True)
        shared = keys_seen[owners > 1]
        clean = ~np.isin(per_env, shared).any(1)
        assert clean.mean() > 0.99
        assert np.array_equal(b[clean], cb[clean]) and np.array_equal(a[clean], ca[clean]) and np.array_equal(s[clean], cs[clean])
        size = L.g2048_ctx_qtable_size(big)
        keys, rows = np.zeros(size, np.uint64), np.zeros((size, 4), np.float32)
        assert L.g2048_ctx_qtable_export(big, vp(keys), vp(rows), size) == size
        order = np.argsort(keys)
        keys, rows = keys[order], rows[order]
        wk, wr = tab.export()
        own = np.unique(per_env[clean])                        # touched by clean envs only (shared states are excluded)
        gi, wi = np.searchsorted(keys, own), np.searchsorted(wk, own)
        assert np.array_equal(keys[gi], own) and np.array_equal(wk[wi], own)
        assert np.array_equal(rows[gi], wr[wi].astype(np.float32))
        assert len(own) > 2 * n
    finally:
        L.g2048_ctx_destroy(big)


def test_fused_rollout_applies_every_update_when_all_envs_update_one_value(L):
    """The worst collision there is: 32,768 envs on the SAME board take the same greedy action (zero row: action 0) in
    one step, i.e. 32,768 simultaneous updates of one Q value, all with the same target (their next states differ by
    the spawned tile and are new: max Q = 0).  Sequential Q-learning -- in any order -- gives q_N = T (1 - (1 - lr)^N);
    every update dropped would lower it by lr e^-2 T = 8e-6 T.  The kernel that skipped an update whenever its
    compare-and-swap lost (round 1) keeps a handful of them; here LOST must be 0 and the value within 1e-3 of the chain
    (the long group is folded segment-wise: float32 rounding of the composition, see k_defer_apply)."""
    n, seed, lr = 1 << 15, 7, 2.0 ** -14
    big = L.g2048_ctx_create(0, n, 1 << 20)
    assert big, L.g2048_last_error()
    try:
        b0 = late_game_boards(np.random.RandomState(11), 1)[0]
        assert oracle.move(np.array([b0], np.uint64), np.zeros(1, np.uint8))[1][0]   # action 0 moves this board
        b = np.full(n, b0, np.uint64)
        a, s = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
        c = np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_qlearn(big, vp(b), vp(a), vp(s), n, 1, 0, lr, 0.99, 0.0, seed, 0, 0, vp(c)))
        assert c[0] == n and c[8] == 0 and c[9] > n // 2       # nearly every first attempt loses its race
        one = np.array([b0], np.uint64)
        reward, flags, _, _ = oracle.env_step(one.copy(), np.array([oracle.AUX_INIT], np.uint64), np.zeros(1, np.int32),
                                              np.zeros(1, np.uint8), None, 0, seed, 0, 0)
        target = np.float32(reward[0])
        q = np.float32(0)
        for _ in range(n):
            q = np.float32(q + np.float32(np.float32(lr) * np.float32(target - q)))
        rows, found = np.zeros((1, 4), np.float32), np.zeros(1, np.uint8)
        ok(L, L.g2048_ctx_qtable_lookup(big, vp(one), 1, vp(rows), vp(found), 0))
        assert found[0] and abs(float(target)) > 0.01
        assert abs(rows[0, 0] - q) <= 1e-3 * abs(q), (rows[0], q, target)
        assert (rows[0, 1:] == 0).all()
        assert L.g2048_ctx_qtable_size(big) == c[6] == 1 + len(np.unique(b))
    finally:
        L.g2048_ctx_destroy(big)


# ------------------------------------------------------------------------------------------ distributions
def test_philox_mode_matches_the_reference_statistics(L, ctx):
    """Philox cannot follow MT19937 draw by draw, so the Philox mode is checked statistically against the reference's
    own random-policy numbers: 12,000 episodes per flavour of the unmodified envs (oracle/ref_stats.py):
    penalty env  : episode length 142.064 +- 0.427 (standard error), game score 1099.69 +- 4.87, invalid moves 0.16395
    nopenalty env: episode length 133.392 +- 0.390, game score 1017.15 +- 4.50, invalid moves 0.15393
    The GPU plays ~280 episodes on each of 65,536 envs (its own standard error is negligible), so the tolerances are
    3.5 standard errors of the REFERENCE's means: 1.1 % on the length, 1.6 % on the score, 0.003 on the invalid fraction
    (was 4 % / 5 % / 0.02 against 1,000 reference episodes).  The unfinished last game of every env is accounted for
    by its expected share (renewal theory: half an episode plus the variance term)."""
    n, k = 1 << 16, 40000
    for flavour, length, l_sem, score, s_sem, invalid in ((0, 142.064, 0.427, 1099.69, 4.87, 0.16395),
                                                          (1, 133.392, 0.390, 1017.15, 4.50, 0.15393)):
        b, a, s = fresh_envs(n, seed=77 + flavour)
        c = np.zeros(16, np.int64)
        ok(L, L.g2048_ctx_rollout_random(ctx, vp(b), vp(a), vp(s), n, k, flavour, 77 + flavour, 0, 0, vp(c)))
        steps, valid, episodes, total_score = (float(x) for x in c[:4])
        # the running game of every env has played E[L^2] / (2 E[L]) ~ 0.55 episodes' worth of steps and ~ 0.3 of a score
        est_len = steps / (episodes + 0.55 * n)
        est_score = total_score / (episodes + 0.30 * n)
        assert abs(est_len - length) < 3.5 * l_sem, (est_len, length)
        assert abs((1 - valid / steps) - invalid) < 0.003, (1 - valid / steps, invalid)
        assert abs(est_score - score) < 3.5 * s_sem, (est_score, score)
    m = 1 << 20
    boards = np.zeros(m, np.uint64)
    ok(L, L.g2048_ctx_env_reset(ctx, vp(boards), None, None, None, m, 4242, 9, 0))
    cells = oracle.unpack_i64(boards).reshape(m, 16)
    assert ((cells != 0).sum(1) == 2).all()
    fours = (cells == 4).sum() / (2.0 * m)
    assert abs(fours - 0.1) < 0.002
    occupancy = (cells != 0).mean(0)
    assert np.abs(occupancy - 2 / 16).max() < 0.003                       # uniform over the 16 cells
