"""CPU-side checks of what the library uploads to the GPU (g2048_host_tables needs no device): the row LUT and
the float64 reward tables against the oracle -- the tables are what makes the shaped reward bit-exact."""
import ctypes as C

import numpy as np

import g2048
import oracle


def tables():
    L = g2048.lib()
    row, mg, ms = np.zeros(65536, np.uint16), np.zeros(65536, np.uint8), np.zeros(256, np.uint32)
    rv, ri, pen = np.zeros(16 * 16 * 256), np.zeros(2 * 16 * 16), np.zeros(32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.g2048_host_tables(p(row), p(mg), p(ms), p(rv), p(ri), p(pen))
    return row, mg, ms, rv, ri, pen


def test_row_lut_equals_oracle_and_literal_move_left():
    row, mg, ms, *_ = tables()
    orow, omg = oracle.row_table()
    assert np.array_equal(row, orow) and np.array_equal(mg, omg)
    for m in range(256):
        hi, lo = m >> 4, m & 15
        assert ms[m] == ((1 << hi if hi else 0) + (1 << lo if lo else 0)) | (hi << 24)
    # a literal restatement of move_left (Game2048_env.py:22-46) on raw tile values, sampled rows
    rng = np.random.RandomState(0)
    for r in rng.randint(0, 65536, 3000).tolist():
        tiles = [(1 << ((r >> (4 * c)) & 15)) if (r >> (4 * c)) & 15 else 0 for c in range(4)]
        if tiles.count(32768) >= 2:
            continue
        nz = [t for t in tiles if t]
        out, skip = [], False
        for i in range(len(nz)):
            if skip:
                skip = False
                continue
            if i + 1 < len(nz) and nz[i] == nz[i + 1]:
                out.append(nz[i] * 2)
                skip = True
            else:
                out.append(nz[i])
        out += [0] * (4 - len(out))
        want = sum((v.bit_length() - 1 if v else 0) << (4 * c) for c, v in enumerate(out))
        assert row[r] == want, hex(r)


def test_reward_tables_are_the_reference_expressions_bit_for_bit():
    """Every table entry == calculate_reward + update_and_normalize (Game2048_env.py:136-205) evaluated by the
    oracle (pinned on reference goldens) -- compared as float64 bit patterns."""
    *_, rv, ri, pen = tables()
    lib = oracle.load()
    for lvl in range(1, 16):
        for d in range(0, lvl):
            for over in (0, 1):
                prev = C.c_int(lvl - d)
                got = lib.orc_calculate_reward(0, 0, over, lvl, C.byref(prev))
                assert np.float64(got).view(np.uint64) == ri[over * 256 + lvl * 16 + d].view(np.uint64), (lvl, d, over)
                assert prev.value == lvl
            for s4 in range(256):
                prev = C.c_int(lvl - d)
                got = lib.orc_calculate_reward(4 * s4, 1, 0, lvl, C.byref(prev))
                assert np.float64(got).view(np.uint64) == rv[(lvl * 16 + d) * 256 + s4].view(np.uint64), (lvl, d, s4)
    # scores >= 1024 normalise to exactly 10 (the device skips the table there)
    for lvl in range(1, 16):
        for s in (1024, 1028, 4096, 262144):
            prev = C.c_int(lvl)
            assert lib.orc_calculate_reward(s, 1, 0, lvl, C.byref(prev)) == 10.0
    assert [pen[k] for k in range(32)] == [-1.0] + [lib.orc_stall_penalty(k) for k in range(1, 32)]


def test_oracle_move_properties():
    """valid <=> board changed; moves conserve the tile sum; legal mask == OR of the four trial moves; dead <=> full
    and no legal move (test plan item 3) on random boards incl. level-15 tiles."""
    rng = np.random.RandomState(1)
    n = 20000
    lv = rng.randint(0, 16, size=(n, 16)) * (rng.random_sample((n, 16)) < 0.7)
    boards = np.zeros(n, np.uint64)
    for j in range(16):
        boards |= lv[:, j].astype(np.uint64) << np.uint64(4 * j)
    boards[boards == 0] = 1
    tiles = oracle.unpack_i64(boards).reshape(n, 16).sum(1)
    lm = np.zeros(n, np.uint8)
    for a in range(4):
        out, moved, score = oracle.move(boards, np.full(n, a, np.uint8))
        assert np.array_equal(moved.astype(bool), out != boards)
        assert np.array_equal(oracle.unpack_i64(out).reshape(n, 16).sum(1), tiles)
        assert ((score % 4) == 0).all() and (score[~moved.astype(bool)] == 0).all()
        lm |= (moved << a).astype(np.uint8)
    assert np.array_equal(lm, oracle.legal_mask(boards))
    full = (oracle.unpack_i64(boards).reshape(n, 16) != 0).all(1)
    assert np.array_equal(oracle.dead(boards).astype(bool), full & (lm == 0))
    # Philox known answers (Random123 kat vectors, philox4x32-10)
    assert oracle.philox(0, 0, 0, 0).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
