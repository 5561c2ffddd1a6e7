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
