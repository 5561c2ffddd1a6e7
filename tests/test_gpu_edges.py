"""Edge cases of the C ABI on the GPU: empty and ragged batches, optional (NULL) outputs, argument errors,
huge env ids / step indices, a table that runs full, malformed replay input."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def L():
    import g2048
    g2048.init(0)
    return g2048.lib()


@pytest.fixture(scope="module")
def ctx(L):
    h = L.g2048_ctx_create(0, 1 << 16, 1 << 12)
    assert h
    yield h
    L.g2048_ctx_destroy(h)


def envs(n, seed=3, base=0):
    b = np.zeros(n, np.uint64)
    oracle.env_reset(b, None, None, None, seed=seed, env_id_base=base)
    return b, np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)


def test_empty_batches_are_no_ops(L, ctx):
    z8, z4, z1 = np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.uint8)
    c = np.zeros(16, np.int64)
    assert L.g2048_ctx_env_step(ctx, vp(z8), vp(z8), vp(z4), vp(z1), None, None, None, None, None, 0, 0, 1, 2, 3) == 0
    assert L.g2048_ctx_env_reset(ctx, vp(z8), None, None, None, 0, 1, 2, 3) == 0
    assert L.g2048_ctx_rollout_random(ctx, vp(z8), vp(z8), vp(z4), 0, 10, 0, 1, 2, 3, vp(c)) == 0
    assert L.g2048_ctx_rollout_qlearn(ctx, vp(z8), vp(z8), vp(z4), 0, 10, 0, 0.1, 0.9, 0.1, 1, 2, 3, vp(c)) == 0
    assert L.g2048_ctx_qtable_update(ctx, vp(z8), vp(z1), vp(np.zeros(0, np.float32)), vp(z8), vp(z1), 0, 0.1, 0.9, 1) == 0
    assert not c.any()
    b, a, s = envs(5)
    b0 = b.copy()
    assert L.g2048_ctx_rollout_random(ctx, vp(b), vp(a), vp(s), 5, 0, 0, 1, 2, 3, vp(c)) == 0   # zero steps
    assert np.array_equal(b, b0)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 255, 1025, 16385, 40001])
@pytest.mark.parametrize("flavour", [0, 1])
def test_ragged_batch_sizes(L, ctx, n, flavour):
    """Sizes that are not multiples of the warp / block / grid, on both LUT paths (global below 16,384 envs,
    shared memory above), with huge env ids and step indices (Philox counter words 1 and 3)."""
    seed, base, t0, k = 5, (1 << 40) + 12345, (1 << 33) + 7, 37
    b, a, s = envs(n, seed, base)
    cb, ca, cs = b.copy(), a.copy(), s.copy()
    c = np.zeros(16, np.int64)
    assert L.g2048_ctx_rollout_random(ctx, vp(b), vp(a), vp(s), n, k, flavour, seed, t0, base, vp(c)) == 0
    cc = oracle.rollout_random(cb, ca, cs, k, flavour, seed, t0, base, threads=4)
    assert np.array_equal(b, cb) and np.array_equal(s, cs) and np.array_equal(c, cc)
    # one single step, every output
    actions = (np.arange(n) % 4).astype(np.uint8)
    r, f, m, ms = np.zeros(n), np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32)
    assert L.g2048_ctx_env_step(ctx, vp(b), vp(a), vp(s), vp(actions), None, vp(r), vp(f), vp(m), vp(ms), n, flavour,
                                seed, t0 + k, base) == 0
    wr, wf, wm, wms = oracle.env_step(cb, ca, cs, actions, None, flavour, seed, t0 + k, base)
    assert np.array_equal(b, cb) and np.array_equal(f, wf) and np.array_equal(m, wm) and np.array_equal(ms, wms)
    assert np.array_equal(r.view(np.uint64), wr.view(np.uint64))


def test_optional_outputs_may_be_null(L, ctx):
    n = 1000
    b, a, s = envs(n)
    cb, ca, cs = b.copy(), a.copy(), s.copy()
    actions = np.full(n, 2, np.uint8)
    assert L.g2048_ctx_env_step(ctx, vp(b), None, None, vp(actions), None, None, None, None, None, n, 1, 9, 0, 0) == 0
    oracle.env_step(cb, None, None, actions, None, 1, 9, 0, 0)
    assert np.array_equal(b, cb)
    assert L.g2048_ctx_rollout_random(ctx, vp(b), None, None, n, 5, 1, 9, 1, 0, None) == 0
    oracle.rollout_random(cb, None, None, 5, 1, 9, 1, 0)
    assert np.array_equal(b, cb)


def test_argument_errors_are_reported_not_crashed(L, ctx):
    b, a, s = envs(8)
    act = np.zeros(8, np.uint8)
    assert L.g2048_ctx_env_step(ctx, None, vp(a), vp(s), vp(act), None, None, None, None, None, 8, 0, 0, 0, 0) != 0
    assert b"null" in L.g2048_last_error()
    assert L.g2048_ctx_env_step(ctx, vp(b), vp(a), vp(s), vp(act), None, None, None, None, None, 8, 7, 0, 0, 0) != 0   # flavour
    assert L.g2048_ctx_env_step(ctx, vp(b), vp(a), vp(s), vp(act), None, None, None, None, None, 1 << 20, 0, 0, 0, 0) != 0  # > max_envs
    assert L.g2048_ctx_env_step(None, vp(b), vp(a), vp(s), vp(act), None, None, None, None, None, 8, 0, 0, 0, 0) != 0
    assert not L.g2048_ctx_create(0, 16, 1000)            # capacity not a power of two
    assert not L.g2048_ctx_create(0, 0, 0)
    no_table = L.g2048_ctx_create(0, 16, 0)
    assert no_table
    assert L.g2048_ctx_rollout_qlearn(no_table, vp(b), vp(a), vp(s), 8, 1, 0, 0.1, 0.9, 0.1, 0, 0, 0, None) != 0
    assert L.g2048_ctx_qtable_size(no_table) == -1
    L.g2048_ctx_destroy(no_table)
    import torch
    t = torch.zeros(64, dtype=torch.int64, device="cuda")
    assert L.g2048_qtable_lookup(t.data_ptr(), 3, t.data_ptr(), 1, t.data_ptr(), None, 0, None) != 0   # capacity 3


def test_malformed_replay_draws_stay_in_range(L, ctx):
    """Replay draws are untrusted input: a cell index beyond the number of empty cells must not corrupt the board."""
    n = 4096
    rng = np.random.RandomState(0)
    b, a, s = envs(n)
    draws = rng.randint(0, 256, size=(n, 4)).astype(np.uint8)
    actions = rng.randint(0, 4, n).astype(np.uint8)
    before = b.copy()
    assert L.g2048_ctx_env_step(ctx, vp(b), vp(a), vp(s), vp(actions), vp(draws), None, None, None, None, n, 0, 0, 0, 0) == 0
    tiles_before, tiles_after = oracle.unpack_i64(before).sum((1, 2)), oracle.unpack_i64(b).sum((1, 2))
    grew = tiles_after - tiles_before
    assert set(np.unique(grew)).issubset({0, 2, 4})       # exactly one spawned 2/4 or nothing
    assert L.g2048_ctx_env_reset(ctx, vp(b), None, None, vp(draws), n, 0, 0, 0) == 0
    cells = oracle.unpack_i64(b)
    assert ((cells != 0).sum((1, 2)) == 2).all() and np.isin(cells, (0, 2, 4)).all()


def test_table_that_runs_full_drops_and_survives(L):
    """Capacity 4,096 slots against ~100k distinct states: lookups that hit the probe limit are counted as dropped,
    the rollout still finishes, the table never reports more states than slots, keys stay unique."""
    small = L.g2048_ctx_create(0, 1 << 15, 1 << 12)
    try:
        n = 20000
        b, a, s = envs(n)
        c = np.zeros(16, np.int64)
        assert L.g2048_ctx_rollout_qlearn(small, vp(b), vp(a), vp(s), n, 40, 0, 0.1, 0.99, 0.5, 1, 0, 0, vp(c)) == 0
        assert c[0] == n * 40 and c[7] > 0
        size = L.g2048_ctx_qtable_size(small)
        assert 0 < size <= 1 << 12 and size == c[6]
        keys, rows = np.zeros(size, np.uint64), np.zeros((size, 4), np.float32)
        assert L.g2048_ctx_qtable_export(small, vp(keys), vp(rows), size) == size
        assert len(np.unique(keys)) == size and np.isfinite(rows).all()
        # the synchronous path on the same full table
        sc = b.copy()
        act = np.zeros(n, np.uint8); r = np.ones(n, np.float32); d = np.zeros(n, np.uint8)
        assert L.g2048_ctx_qtable_update(small, vp(sc), vp(act), vp(r), vp(b), vp(d), n, 0.1, 0.9, 1) == 0
        assert L.g2048_ctx_qtable_size(small) <= 1 << 12
    finally:
        L.g2048_ctx_destroy(small)


def test_concurrent_inserts_of_the_same_keys_never_duplicate(L):
    """A million threads look up (with insert) only 1,000 distinct states at once, then 2^20 distinct ones: every key
    ends up in exactly one slot and every later lookup finds it (races on the key CAS are resolved correctly)."""
    import torch
    cap = 1 << 22
    table = torch.zeros(cap * 4, dtype=torch.int64, device="cuda")
    rng = np.random.RandomState(5)
    pool = (rng.randint(1, 1 << 62, size=1000, dtype=np.int64) | 1)
    n = 1 << 20
    keys = torch.from_numpy(pool[rng.randint(0, 1000, n)]).cuda()
    rows = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    found = torch.empty(n, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.g2048_qtable_lookup(table.data_ptr(), cap, keys.data_ptr(), n, rows.data_ptr(), found.data_ptr(), 1, st) == 0
    assert L.g2048_qtable_size(table.data_ptr(), cap, cnt.data_ptr(), st) == 0
    assert int(cnt.item()) == 1000
    assert L.g2048_qtable_lookup(table.data_ptr(), cap, keys.data_ptr(), n, rows.data_ptr(), found.data_ptr(), 0, st) == 0
    assert bool(found.bool().all()) and float(rows.abs().sum()) == 0.0
    distinct = torch.from_numpy(np.unique(rng.randint(1, 1 << 62, size=n, dtype=np.int64))).cuda()
    m = distinct.numel()
    assert L.g2048_qtable_lookup(table.data_ptr(), cap, distinct.data_ptr(), m, rows.data_ptr(), found.data_ptr(), 1, st) == 0
    assert L.g2048_qtable_size(table.data_ptr(), cap, cnt.data_ptr(), st) == 0
    extra = int((~torch.isin(distinct, torch.from_numpy(pool).cuda())).sum().item())
    assert int(cnt.item()) == 1000 + extra
    ek = torch.empty(cap, dtype=torch.int64, device="cuda")
    er = torch.empty((cap, 4), dtype=torch.float32, device="cuda")
    cnt.zero_()
    assert L.g2048_qtable_export(table.data_ptr(), cap, ek.data_ptr(), er.data_ptr(), cap, cnt.data_ptr(), st) == 0
    got = ek[: int(cnt.item())]
    assert got.unique().numel() == got.numel()


def test_c_program_drives_the_abi(L):
    """examples/c_api_demo.c: reset / choose_action / step / update_q_value / fused rollout from plain C."""
    import os
    import subprocess
    import tempfile
    import g2048
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so_dir = os.path.dirname(g2048.build())
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "demo")
        subprocess.run(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_api_demo.c"),
                        "-L" + so_dir, "-lg2048", "-Wl,-rpath," + so_dir, "-o", exe], check=True)
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode == 0, run.stdout + run.stderr
        assert "fused rollout: 1048576 steps" in run.stdout


@pytest.mark.parametrize("replay", [False, True])
@pytest.mark.parametrize("flavour", [0, 1])
def test_env_step_large_batch_path_equals_small_batch_path(flavour, replay):
    """From 2^19 envs on g2048_env_step runs as persistent CTAs with the LUT in shared memory; the same envs stepped in
    chunks of 2^17 (LUT through L1) must give identical boards, aux, scores, rewards, flags and max tiles."""
    import torch
    import g2048
    g2048.init(0)
    L = g2048.lib()
    n, chunk, seed = (1 << 19) + 1234, 1 << 17, 31
    st = torch.cuda.current_stream().cuda_stream

    def fresh():
        b = torch.zeros(n, dtype=torch.int64, device="cuda")
        a = torch.full((n,), 0xFF01, dtype=torch.int64, device="cuda")
        s = torch.zeros(n, dtype=torch.int32, device="cuda")
        assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, seed, 0, 0, st) == 0
        return b, a, s

    outs = []
    for big in (True, False):
        b, a, s = fresh()
        r = torch.zeros(n, dtype=torch.float64, device="cuda")
        f = torch.zeros(n, dtype=torch.uint8, device="cuda")
        m = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ms = torch.zeros(n, dtype=torch.int32, device="cuda")
        gen = torch.Generator(device="cuda").manual_seed(5)
        for t in range(40):
            act = ((torch.arange(n, device="cuda") * 7 + t * 3) % 4).to(torch.uint8)
            # recorded draws (spawn cell, is-4, and the nopenalty full-board pair): 4 bytes per env, same for both paths
            draws = torch.randint(0, 16, (n, 4), device="cuda", generator=gen).to(torch.uint8)
            draws[:, 1] = (draws[:, 1] == 0).to(torch.uint8)
            draws[:, 3] = (draws[:, 3] == 0).to(torch.uint8)
            spans = [(0, n)] if big else [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
            for lo, hi in spans:
                rc = L.g2048_env_step(b[lo:].data_ptr(), a[lo:].data_ptr(), s[lo:].data_ptr(), act[lo:].data_ptr(),
                                      draws[lo:].data_ptr() if replay else None,
                                      r[lo:].data_ptr(), None, f[lo:].data_ptr(), m[lo:].data_ptr(), ms[lo:].data_ptr(),
                                      hi - lo, flavour, seed, t, lo, st)
                assert rc == 0, L.g2048_last_error()
        outs.append((b, a, s, r, f, m, ms))
    for name, x, y in zip(("boards", "aux", "score", "reward", "flags", "maxlvl", "move_score"), *outs):
        bad = (x != y).nonzero().flatten()
        assert bad.numel() == 0, (name, bad.numel(), bad[:5].tolist(), x[bad[:5]].tolist(), y[bad[:5]].tolist())
    assert int(outs[0][6].sum()) > 0 and int((outs[0][4] & 1).sum()) > n // 2      # merges and valid moves happened


def test_move_properties_at_1M_boards_on_the_gpu():
    """BASELINE size, size-independent properties checked on the GPU's own outputs (no oracle in the loop): tile-sum
    conservation, score > 0 iff a tile disappeared, the four directions as mirror images / transposes of one another,
    legal mask == the four trial moves, all on 2^20 random boards of every density."""
    import torch
    import g2048
    g2048.init(0)
    L = g2048.lib()
    n = 1 << 20
    st = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda").manual_seed(7)
    lv = torch.randint(1, 12, (n, 16), device="cuda", generator=gen)
    p_zero = torch.rand((n, 1), device="cuda", generator=gen)
    lv = torch.where(torch.rand((n, 16), device="cuda", generator=gen) < p_zero, torch.zeros_like(lv), lv)
    lv[:, 0] = torch.where(lv.sum(1) == 0, torch.ones_like(lv[:, 0]), lv[:, 0])

    def pack(levels):
        b = torch.zeros(levels.shape[0], dtype=torch.int64, device="cuda")
        for j in range(16):
            b |= levels[:, j] << (4 * j)
        return b

    def unpack(b):
        return torch.stack([(b >> (4 * j)) & 15 for j in range(16)], dim=1)

    def mirror(b):
        return pack(unpack(b).view(-1, 4, 4).flip(2).reshape(-1, 16))

    def transpose(b):
        return pack(unpack(b).view(-1, 4, 4).transpose(1, 2).reshape(-1, 16))

    def tile_sum(b):
        u = unpack(b)
        return torch.where(u > 0, torch.ones_like(u) << u, torch.zeros_like(u)).sum(1)

    def move(b, action):
        a = torch.full((n,), action, dtype=torch.uint8, device="cuda")
        out = torch.empty_like(b)
        moved = torch.empty(n, dtype=torch.uint8, device="cuda")
        score = torch.empty(n, dtype=torch.int32, device="cuda")
        assert L.g2048_move_trial(b.data_ptr(), a.data_ptr(), out.data_ptr(), moved.data_ptr(), score.data_ptr(), n, st) == 0
        return out, moved, score

    boards = pack(lv)
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert L.g2048_legal_mask(boards.data_ptr(), mask.data_ptr(), n, st) == 0
    left, mv_l, sc_l = move(boards, 0)
    for action in range(4):
        out, moved, score = move(boards, action)
        assert torch.equal(tile_sum(out), tile_sum(boards))
        fewer = (unpack(out) > 0).sum(1) < (unpack(boards) > 0).sum(1)
        assert torch.equal(score > 0, fewer) and torch.equal(moved != 0, out != boards)
        assert torch.equal((mask >> action) & 1, moved)
    right, mv_r, sc_r = move(mirror(boards), 2)
    assert torch.equal(mirror(right), left) and torch.equal(mv_r, mv_l) and torch.equal(sc_r, sc_l)
    up, mv_u, sc_u = move(transpose(boards), 1)
    assert torch.equal(transpose(up), left) and torch.equal(mv_u, mv_l) and torch.equal(sc_u, sc_l)
    down, mv_d, sc_d = move(transpose(mirror(boards)), 3)
    assert torch.equal(mirror(transpose(down)), left) and torch.equal(mv_d, mv_l) and torch.equal(sc_d, sc_l)
    full = (unpack(boards) > 0).all(1)
    assert int((full & (mask == 0)).sum()) > 0 and int((mask == 15).sum()) > 0
