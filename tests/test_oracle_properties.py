"""Size-independent properties of the 2048 move on the CPU oracle (which the GPU path must equal bit for bit): tile-sum
conservation, the four directions as mirror images / transposes of one another, legal mask == "the trial move changes
the board" (mainDQL_CNN_step2.py:169-174), dead == full board without a legal move (Game2048_env.py:65-75), pack/unpack
round trip, and the row table as the single-row case of the move.  Random boards of every density, 200,000 of them."""
import numpy as np
import pytest

import oracle

N = 200_000


def boards_from_levels(lv):
    b = np.zeros(len(lv), np.uint64)
    for j in range(16):
        b |= lv[:, j].astype(np.uint64) << np.uint64(4 * j)
    return b


def levels(b):
    return np.stack([(b >> np.uint64(4 * j)) & np.uint64(15) for j in range(16)], axis=1).astype(np.int64)


@pytest.fixture(scope="module")
def boards():
    rng = np.random.RandomState(2048)
    lv = rng.randint(1, 12, size=(N, 16))
    p_zero = rng.random_sample((N, 1))                     # every density from empty-ish to full
    lv = np.where(rng.random_sample((N, 16)) < p_zero, 0, lv)
    lv[:, 0] = np.where(lv.sum(1) == 0, 1, lv[:, 0])       # never the all-empty board
    return boards_from_levels(lv)


def mirror(b):      # reverse every row: cell (r, c) -> (r, 3 - c)
    lv = levels(b).reshape(-1, 4, 4)[:, :, ::-1].reshape(-1, 16)
    return boards_from_levels(lv)


def transpose(b):   # cell (r, c) -> (c, r)
    lv = levels(b).reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)
    return boards_from_levels(lv)


def tile_sum(b):
    lv = levels(b)
    return np.where(lv > 0, 1 << lv, 0).sum(1)


@pytest.mark.parametrize("action", [0, 1, 2, 3])
def test_move_conserves_the_tile_sum_and_scores_the_merges(boards, action):
    a = np.full(N, action, np.uint8)
    out, moved, score = oracle.move(boards, a)
    assert np.array_equal(tile_sum(out), tile_sum(boards))
    n_before, n_after = (levels(boards) > 0).sum(1), (levels(out) > 0).sum(1)
    assert np.all(n_after <= n_before) and np.all((score > 0) == (n_after < n_before))     # every merge removes a tile
    assert np.all(score % 4 == 0)                                                           # merged tiles are >= 4
    assert np.array_equal(moved.astype(bool), out != boards)


def test_directions_are_mirror_images_and_transposes(boards):
    left, mv_l, sc_l = oracle.move(boards, np.zeros(N, np.uint8))
    right, mv_r, sc_r = oracle.move(mirror(boards), np.full(N, 2, np.uint8))
    assert np.array_equal(mirror(right), left) and np.array_equal(mv_l, mv_r) and np.array_equal(sc_l, sc_r)
    up, mv_u, sc_u = oracle.move(transpose(boards), np.full(N, 1, np.uint8))
    assert np.array_equal(transpose(up), left) and np.array_equal(mv_l, mv_u) and np.array_equal(sc_l, sc_u)
    down, mv_d, sc_d = oracle.move(transpose(mirror(boards)), np.full(N, 3, np.uint8))
    assert np.array_equal(mirror(transpose(down)), left) and np.array_equal(mv_l, mv_d) and np.array_equal(sc_l, sc_d)


def test_legal_mask_and_dead_follow_from_the_trial_moves(boards):
    mask = oracle.legal_mask(boards)
    for action in range(4):
        _, moved, _ = oracle.move(boards, np.full(N, action, np.uint8))
        assert np.array_equal((mask >> action) & 1, moved)
    full = (levels(boards) > 0).all(1)
    assert np.array_equal(oracle.dead(boards).astype(bool), full & (mask == 0))
    assert (full & (mask == 0)).sum() > 0 and (mask == 15).sum() > 0           # both kinds occur in the sample


def test_a_move_is_not_undone_by_moving_again_without_merges(boards):
    """Sliding is idempotent: after a left move, a second left move can only merge, never slide -- so if it scores
    nothing it changes nothing."""
    a = np.zeros(N, np.uint8)
    once, _, _ = oracle.move(boards, a)
    twice, moved, score = oracle.move(once, a)
    assert np.array_equal(moved.astype(bool), score > 0)
    assert np.array_equal(twice[score == 0], once[score == 0])


def test_pack_unpack_round_trip_and_row_table(boards):
    tiles = oracle.unpack_i64(boards)
    assert np.array_equal(np.where(tiles > 0, np.log2(np.maximum(tiles, 1)).astype(np.int64), 0).reshape(-1, 16), levels(boards))
    packed, bad = oracle.pack_i64(tiles)
    assert bad == 0 and np.array_equal(packed, boards)
    # the 65,536-row table is the move on a board whose only non-empty row is row 0
    res, _ = oracle.row_table()
    rows = np.arange(65536, dtype=np.uint64)
    rows = rows[rows != 0]
    out, _, _ = oracle.move(rows, np.zeros(len(rows), np.uint8))
    assert np.array_equal(out, res[rows.astype(np.int64)].astype(np.uint64))


# ---------------------------------------------------------------- the batched Q-update, restated independently
def test_batched_update_is_the_reference_rule_applied_in_batch_order():
    """update_q_value (main.py:40-43) for a batch, as this build defines it: all targets from the
    table as it is BEFORE the batch, then every (state, action) receives its targets one after another in batch order,
    q <- q + lr (target - q), every operation rounded to float32.  A literal numpy/dict restatement must give the
    oracle's table bit for bit -- with heavy collisions (a pool of 40 states, 5,000 transitions)."""
    rng = np.random.RandomState(11)
    pool = rng.randint(1, 1 << 40, size=40).astype(np.uint64)
    lr, gamma = np.float32(0.1), np.float32(0.99)
    tab = oracle.QTable(1 << 12, f32=True)
    ref = {}                                               # key -> float32[4]
    for batch in range(6):
        n = 5000
        s, s2 = pool[rng.randint(0, 40, n)], pool[rng.randint(0, 40, n)]
        a = rng.randint(0, 4, n).astype(np.uint8)
        r = rng.standard_normal(n).astype(np.float32)
        done = (rng.random_sample(n) < 0.1).astype(np.uint8)
        tab.update_batch_f32(s, a, r, s2, done, float(lr), float(gamma))
        snap = {k: v.copy() for k, v in ref.items()}
        zero = np.zeros(4, np.float32)
        target = np.empty(n, np.float32)
        for i in range(n):
            best = np.float32(snap.get(int(s2[i]), zero).max())
            g = np.float32(gamma * best)
            target[i] = np.float32(r[i] + (np.float32(0) if done[i] else g))
        for i in range(n):
            row = ref.setdefault(int(s[i]), np.zeros(4, np.float32))
            ref.setdefault(int(s2[i]), np.zeros(4, np.float32))            # reading a state creates its zero row
            q = row[a[i]]
            row[a[i]] = np.float32(q + np.float32(lr * np.float32(target[i] - q)))
    keys, rows = tab.export()
    want_keys = np.array(sorted(ref), np.uint64)
    assert np.array_equal(keys, want_keys)
    want = np.stack([ref[int(k)] for k in want_keys])
    assert np.array_equal(rows.astype(np.float32), want)
    assert np.isfinite(want).all() and np.abs(want).max() < 10      # a contraction: no blow-up under 125 hits per value


def test_apply_targets_is_order_sensitive_and_sequential():
    """apply_targets: the records of one (state, action) are applied in the order given -- reversing them changes the
    float32 result, and the result always lies between the old value and the extreme targets."""
    key = np.full(64, 12345, np.uint64)
    a = np.zeros(64, np.uint8)
    t = np.linspace(-3, 5, 64).astype(np.float32)
    one, two = oracle.QTable(1 << 8, f32=True), oracle.QTable(1 << 8, f32=True)
    one.apply_targets_f32(key, a, t, 0.1)
    two.apply_targets_f32(key, a, t[::-1].copy(), 0.1)
    q1, q2 = one.export()[1][0, 0], two.export()[1][0, 0]
    assert q1 != q2 and -3 <= min(q1, q2) and max(q1, q2) <= 5
    q = np.float32(0)
    for x in t:
        q = np.float32(q + np.float32(np.float32(0.1) * np.float32(x - q)))
    assert q1 == q


@pytest.mark.parametrize("flavour", [oracle.FLAVOUR_PENALTY, oracle.FLAVOUR_NOPENALTY])
def test_synchronous_step_is_choose_then_step_then_batch_update(flavour):
    """orc_qlearn_step_sync (what the GPU's synchronous modes are compared with) == the three reference calls of the
    loop main.py:91-101 composed from the separately pinned pieces: choose_action for every env on the table as it is,
    env.step with the same Philox draws, then update_q_value for the whole batch (targets from the pre-step table,
    applied in env order).  Fresh boards, 25 steps (no game ends that early, so no reset is involved)."""
    n, seed, base, eps, lr, gamma = 700, 31, 1000, 0.35, 0.1, 0.99
    b1 = np.zeros(n, np.uint64)
    oracle.env_reset(b1, None, None, None, seed=seed, episode_idx=0, env_id_base=base)
    a1, s1 = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
    b2, a2, s2 = b1.copy(), a1.copy(), s1.copy()
    t1, t2 = oracle.QTable(1 << 16, f32=True), oracle.QTable(1 << 16, f32=True)
    for t in range(25):
        oracle.qlearn_step_sync(b1, a1, s1, t1, lr, gamma, eps, flavour, seed, t, base)
        state = b2.copy()
        actions = t2.choose_action(state, oracle.eps_threshold(eps), seed, t, base)
        reward, flags, _, _ = oracle.env_step(b2, a2, s2, actions, None, flavour, seed, t, base)
        done = (flags >> 2) & 1
        assert done.sum() == 0
        t2.update_batch_f32(state, actions, reward.astype(np.float32), b2.copy(), done.astype(np.uint8), lr, gamma)
        assert np.array_equal(b1, b2) and np.array_equal(a1, a2) and np.array_equal(s1, s2), t
    (k1, r1), (k2, r2) = t1.export(), t2.export()
    assert np.array_equal(k1, k2) and np.array_equal(r1, r2)
    assert len(k1) > 2000 and np.abs(r1).sum() > 0


@pytest.mark.parametrize("flavour", [oracle.FLAVOUR_PENALTY, oracle.FLAVOUR_NOPENALTY])
def test_sequential_rollout_of_one_env_is_the_reference_loop_step_by_step(flavour):
    """orc_rollout_qlearn_seq with one env (the oracle behind smoke() and the N = 1 GPU parity test) == the loop of
    main.py:91-101 made of single calls: choose_action, env.step, update_q_value, one transition at a time."""
    seed, base, eps, lr, gamma, steps = 9, 5, 0.3, 0.1, 0.99, 120
    b1 = np.zeros(1, np.uint64)
    oracle.env_reset(b1, None, None, None, seed=seed, episode_idx=0, env_id_base=base)
    a1, s1 = np.full(1, oracle.AUX_INIT, np.uint64), np.zeros(1, np.int32)
    b2, a2, s2 = b1.copy(), a1.copy(), s1.copy()
    t1, t2 = oracle.QTable(1 << 12, f32=True), oracle.QTable(1 << 12, f32=True)
    oracle.rollout_qlearn_seq(b1, a1, s1, t1, steps, lr, gamma, eps, flavour, seed, 0, base)
    for t in range(steps):
        state = b2.copy()
        action = t2.choose_action(state, oracle.eps_threshold(eps), seed, t, base)
        reward, flags, _, _ = oracle.env_step(b2, a2, s2, action, None, flavour, seed, t, base)
        assert not (flags[0] >> 2) & 1, "the game must not end inside this test (no reset path in the composition)"
        t2.update_batch_f32(state, action, reward.astype(np.float32), b2.copy(), np.zeros(1, np.uint8), lr, gamma)
    assert np.array_equal(b1, b2) and np.array_equal(a1, a2) and np.array_equal(s1, s2)
    (k1, r1), (k2, r2) = t1.export(), t2.export()
    assert np.array_equal(k1, k2) and np.array_equal(r1, r2) and len(k1) > 50


def test_apply_commutes_across_different_state_actions_only():
    """The invariant behind the sort-based deterministic apply and the owner-computes exchange: applying records is
    invariant under ANY reordering that keeps the relative order of the records of one (state, action) -- e.g. grouping
    by (state, action), or partitioning by the owner of the state and ordering each part by global env index -- and it
    is NOT invariant under reorderings inside one (state, action)."""
    rng = np.random.RandomState(3)
    n, owners = 20_000, 8
    pool = rng.randint(1, 1 << 40, size=300).astype(np.uint64)
    keys = pool[rng.randint(0, 300, n)]
    acts = rng.randint(0, 4, n).astype(np.uint8)
    tg = rng.standard_normal(n).astype(np.float32)

    def table_after(order):
        t = oracle.QTable(1 << 12, f32=True)
        t.apply_targets_f32(keys[order].copy(), acts[order].copy(), tg[order].copy(), 0.1)
        return t.export()

    k0, r0 = table_after(np.arange(n))
    # (a) stable grouping by (state, action): what the radix sort + segment apply does
    grouped = np.lexsort((np.arange(n), acts, keys))
    # (b) owner-computes: records arrive in arbitrary order, are split by owner and sorted by (slot, action, global index)
    arrival = rng.permutation(n)
    owner = (keys[arrival] % np.uint64(owners)).astype(np.int64)
    parts = []
    for j in rng.permutation(owners):                      # the owners work independently, in any order
        mine = arrival[owner == j]
        parts.append(mine[np.lexsort((mine, acts[mine], keys[mine]))])
    owned = np.concatenate(parts)
    for order in (grouped, owned):
        k, r = table_after(order)
        assert np.array_equal(k, k0) and np.array_equal(r, r0)
    # (c) reversing the records inside the (state, action) groups changes the float32 result
    rev = np.lexsort((-np.arange(n), acts, keys))
    k, r = table_after(rev)
    assert np.array_equal(k, k0) and not np.array_equal(r, r0)
