"""CPU oracle for the 2048 / Q-learning hot path -- TEST INFRASTRUCTURE ONLY.

`oracle.load()` builds (gcc) and loads oracle/g2048_oracle.c, a plain-C
restatement of the reference algorithm pinned against golden vectors recorded
from the reference itself (tests/golden/, oracle/make_golden.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libg2048_oracle.so")
_SRC = os.path.join(_HERE, "g2048_oracle.c")

FLAVOUR_PENALTY, FLAVOUR_NOPENALTY = 0, 1
AUX_INIT = 0x000000000000FF01
PEN_SAT = 25
N_COUNTERS = 16


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True, capture_output=True)
    return _SO


_lib = None


def _p(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        vp, i64, u64, i32, u32, f32, f64 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_uint32, C.c_float, C.c_double
        sig = {
            "orc_philox": (None, [u64, u64, u64, u32, vp]),
            "orc_env_step": (None, [vp] * 9 + [i64, i32, u64, u64, u64]),
            "orc_env_reset": (None, [vp] * 4 + [i64, u64, u64, u64]),
            "orc_legal_mask": (None, [vp, vp, i64]),
            "orc_dead": (None, [vp, vp, i64]),
            "orc_move": (None, [vp, vp, vp, vp, i64]),
            "orc_row_table": (None, [vp, vp]),
            "orc_calculate_reward": (f64, [i64, i32, i32, i32, vp]),
            "orc_stall_penalty": (f64, [i32]),
            "orc_pack_i64": (i64, [vp, vp, i64]),
            "orc_unpack_i64": (None, [vp, vp, i64]),
            "orc_encode_onehot": (None, [vp, vp, i64]),
            "orc_qtab_new": (vp, [u64, i32]),
            "orc_qtab_free": (None, [vp]),
            "orc_qtab_size": (u64, [vp]),
            "orc_qtab_export": (i64, [vp, vp, vp, i64]),
            "orc_qtab_get": (i32, [vp, u64, vp]),
            "orc_q_update_seq_f64": (None, [vp] * 6 + [i64, f64, f64]),
            "orc_q_replay_agent_f64": (None, [vp] * 8 + [i64, f64, f64]),
            "orc_q_update_batch_f32": (None, [vp] * 6 + [i64, f32, f32]),
            "orc_q_apply_targets_f32": (None, [vp] * 4 + [f32, i64]),
            "orc_rollout_random": (None, [vp, vp, vp, i64, i64, i32, u64, u64, u64, vp]),
            "orc_choose_action": (None, [vp, vp, i64, u64, u64, u64, u64, vp]),
            "orc_rollout_qlearn_seq": (None, [vp, vp, vp, vp, i64, i64, i32, f32, f32, u64, u64, u64, u64, vp]),
            "orc_qlearn_step_sync": (None, [vp, vp, vp, vp, i64, i32, f32, f32, u64, u64, u64, u64, vp, vp, vp, vp, i32]),
            "orc_rollout_random_mt": (None, [vp, vp, vp, i64, i64, i32, u64, u64, u64, vp, i32]),
            "orc_rollout_qlearn_mt": (None, [vp, vp, vp, i64, i64, i32, f32, f32, u64, u64, u64, u64, vp, u64, i32]),
            "orc_rollout_qlearn_mt_tables": (None, [vp, vp, vp, i64, i64, i32, f32, f32, u64, u64, u64, u64, vp, vp, i32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


# --------------------------------------------------------------------------- numpy-level helpers
def philox(seed, env_id, step, stream):
    out = np.zeros(4, np.uint32)
    load().orc_philox(seed, env_id, step, stream, _p(out, np.uint32))
    return out


def env_step(boards, aux, score, actions, draws=None, flavour=FLAVOUR_PENALTY, seed=0, step_idx=0, env_id_base=0):
    """In-place on boards/aux/score; returns (reward f64, flags u8, maxlvl u8, move_score i32)."""
    n = len(boards)
    reward, flags = np.zeros(n, np.float64), np.zeros(n, np.uint8)
    maxlvl, ms = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    load().orc_env_step(_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), _p(actions, np.uint8),
                        _p(draws, np.uint8), _p(reward, np.float64), _p(flags, np.uint8), _p(maxlvl, np.uint8),
                        _p(ms, np.int32), n, flavour, seed, step_idx, env_id_base)
    return reward, flags, maxlvl, ms


def env_reset(boards, score=None, mask=None, draws=None, seed=0, episode_idx=0, env_id_base=0):
    load().orc_env_reset(_p(boards, np.uint64), _p(score, np.int32), _p(mask, np.uint8), _p(draws, np.uint8),
                         len(boards), seed, episode_idx, env_id_base)


def legal_mask(boards):
    out = np.zeros(len(boards), np.uint8)
    load().orc_legal_mask(_p(boards, np.uint64), _p(out, np.uint8), len(boards))
    return out


def dead(boards):
    out = np.zeros(len(boards), np.uint8)
    load().orc_dead(_p(boards, np.uint64), _p(out, np.uint8), len(boards))
    return out


def move(boards, actions):
    """Returns (new boards, moved u8, score i32); no spawn."""
    b = boards.copy()
    moved, score = np.zeros(len(b), np.uint8), np.zeros(len(b), np.int32)
    load().orc_move(_p(b, np.uint64), _p(actions, np.uint8), _p(moved, np.uint8), _p(score, np.int32), len(b))
    return b, moved, score


def row_table():
    res, mg = np.zeros(65536, np.uint16), np.zeros(65536, np.uint8)
    load().orc_row_table(_p(res, np.uint16), _p(mg, np.uint8))
    return res, mg


def pack_i64(tiles):
    tiles = np.ascontiguousarray(tiles, np.int64).reshape(-1, 16)
    out = np.zeros(len(tiles), np.uint64)
    bad = load().orc_pack_i64(_p(tiles, np.int64), _p(out, np.uint64), len(tiles))
    return out, bad


def unpack_i64(boards):
    out = np.zeros((len(boards), 4, 4), np.int64)
    load().orc_unpack_i64(_p(boards, np.uint64), _p(out, np.int64), len(boards))
    return out


def encode_onehot(boards):
    out = np.zeros((len(boards), 16, 4, 4), np.float32)
    load().orc_encode_onehot(_p(boards, np.uint64), _p(out, np.float32), len(boards))
    return out


class QTable:
    """CPU open-addressing Q-table (float64 rows = reference arithmetic, float32 rows = framework arithmetic)."""

    def __init__(self, capacity: int, f32: bool):
        assert capacity & (capacity - 1) == 0
        self.f32 = f32
        self.h = load().orc_qtab_new(capacity, int(f32))

    def __del__(self):
        if getattr(self, "h", None):
            load().orc_qtab_free(self.h)
            self.h = None

    def __len__(self):
        return int(load().orc_qtab_size(self.h))

    def export(self):
        n = len(self)
        keys, rows = np.zeros(n, np.uint64), np.zeros((n, 4), np.float64)
        load().orc_qtab_export(self.h, _p(keys, np.uint64), _p(rows, np.float64), n)
        order = np.argsort(keys)
        return keys[order], rows[order]

    def get(self, key):
        row = np.zeros(4, np.float64)
        found = load().orc_qtab_get(self.h, int(key), _p(row, np.float64))
        return row, bool(found)

    def update_seq_f64(self, s, a, r, s2, done, lr, gamma):
        assert not self.f32
        load().orc_q_update_seq_f64(self.h, _p(s, np.uint64), _p(a, np.uint8), _p(r, np.float64), _p(s2, np.uint64),
                                    _p(done, np.uint8), len(s), lr, gamma)

    def replay_agent_f64(self, s, explore, rand_action, r, s2, done, lr, gamma):
        """choose_action + update_q_value per transition; returns the actions the agent takes."""
        assert not self.f32
        actions = np.zeros(len(s), np.uint8)
        load().orc_q_replay_agent_f64(self.h, _p(s, np.uint64), _p(explore, np.uint8), _p(rand_action, np.uint8),
                                      _p(r, np.float64), _p(s2, np.uint64), _p(done, np.uint8), _p(actions, np.uint8),
                                      len(s), lr, gamma)
        return actions

    def update_batch_f32(self, s, a, r, s2, done, lr, gamma):
        assert self.f32
        load().orc_q_update_batch_f32(self.h, _p(s, np.uint64), _p(a, np.uint8), _p(r, np.float32), _p(s2, np.uint64),
                                      _p(done, np.uint8), len(s), lr, gamma)

    def apply_targets_f32(self, keys, a, target, lr):
        assert self.f32
        load().orc_q_apply_targets_f32(self.h, _p(keys, np.uint64), _p(a, np.uint8), _p(target, np.float32), lr,
                                       len(keys))

    def choose_action(self, boards, eps_thresh, seed, step_idx, env_id_base=0):
        assert self.f32
        actions = np.zeros(len(boards), np.uint8)
        load().orc_choose_action(self.h, _p(boards, np.uint64), len(boards), eps_thresh, seed, step_idx, env_id_base,
                                 _p(actions, np.uint8))
        return actions


def eps_threshold(eps: float) -> int:
    """explore iff x < floor(eps * 2^32) as a 64-bit compare."""
    return int(min(max(eps, 0.0), 1.0) * 4294967296.0)


def rollout_random(boards, aux, score, k_steps, flavour=FLAVOUR_PENALTY, seed=0, step_base=0, env_id_base=0,
                   threads=1):
    counters = np.zeros(N_COUNTERS, np.int64)
    args = [_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), len(boards), k_steps, flavour, seed,
            step_base, env_id_base, _p(counters, np.int64)]
    if threads > 1:
        load().orc_rollout_random_mt(*args, threads)
    else:
        load().orc_rollout_random(*args)
    return counters


def rollout_qlearn_seq(boards, aux, score, table: QTable, k_steps, lr, gamma, eps, flavour=FLAVOUR_PENALTY, seed=0,
                       step_base=0, env_id_base=0):
    counters = np.zeros(N_COUNTERS, np.int64)
    load().orc_rollout_qlearn_seq(_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), table.h,
                                  len(boards), k_steps, flavour, lr, gamma, eps_threshold(eps), seed, step_base,
                                  env_id_base, _p(counters, np.int64))
    return counters


def rollout_qlearn_mt(boards, aux, score, k_steps, lr, gamma, eps, table_capacity, threads, flavour=FLAVOUR_PENALTY,
                      seed=0, step_base=0, env_id_base=0):
    """CPU baseline: `threads` independent shards, each with its own table (sequential semantics per shard)."""
    counters = np.zeros(N_COUNTERS, np.int64)
    load().orc_rollout_qlearn_mt(_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), len(boards), k_steps,
                                 flavour, lr, gamma, eps_threshold(eps), seed, step_base, env_id_base,
                                 _p(counters, np.int64), table_capacity, threads)
    return counters


def rollout_qlearn_mt_tables(boards, aux, score, k_steps, lr, gamma, eps, tables, flavour=FLAVOUR_PENALTY, seed=0,
                             step_base=0, env_id_base=0):
    """CPU baseline with persistent tables: thread w plays its env shard against tables[w] (QTable, f32) across calls."""
    counters = np.zeros(N_COUNTERS, np.int64)
    arr = (C.c_void_p * len(tables))(*[t.h for t in tables])
    load().orc_rollout_qlearn_mt_tables(_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), len(boards),
                                        k_steps, flavour, lr, gamma, eps_threshold(eps), seed, step_base, env_id_base,
                                        _p(counters, np.int64), arr, len(tables))
    return counters


def qlearn_step_sync(boards, aux, score, table: QTable, lr, gamma, eps, flavour=FLAVOUR_PENALTY, seed=0, step=0,
                     env_id_base=0, records=False, apply=True):
    n = len(boards)
    counters = np.zeros(N_COUNTERS, np.int64)
    rk = np.zeros(n, np.uint64) if records else None
    ra = np.zeros(n, np.uint8) if records else None
    rd = np.zeros(n, np.float32) if records else None
    load().orc_qlearn_step_sync(_p(boards, np.uint64), _p(aux, np.uint64), _p(score, np.int32), table.h, n, flavour,
                                lr, gamma, eps_threshold(eps), seed, step, env_id_base, _p(counters, np.int64),
                                _p(rk, np.uint64), _p(ra, np.uint8), _p(rd, np.float32), int(apply))
    return counters, (rk, ra, rd)


def decay_exploration_schedule(total_epochs, eps0=1.0, eps_min=0.01):
    """QLearningAgent.__init__/decay_exploration (main.py:15-32, 45-57): epsilon before each epoch, plus the final one."""
    first, second, third = total_epochs * 0.30, total_epochs * 0.60, total_epochs * 0.80
    slow1 = (eps0 - (eps_min * 1.5)) / first
    fast = ((eps0 - eps_min) - (eps_min * 1.5)) / (second - first)
    slow2 = (eps_min * 1.1 - eps_min) / (third - second)
    eps, out = eps0, []
    for epoch in range(total_epochs):
        out.append(eps)
        if epoch < first:
            eps = max(eps_min * 1.5, eps - slow1)
        elif epoch < second:
            eps = max(eps_min * 1.1, eps - fast)
        elif epoch < third:
            eps = max(eps_min, eps - slow2)
        else:
            eps = eps_min
    out.append(eps)
    return np.array(out, np.float64)
