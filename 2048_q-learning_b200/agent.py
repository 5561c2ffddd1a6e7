"""Tabular Q-learning on an open-addressing hash table in HBM (device-pointer API of libg2048.so).

`BatchedQLearningAgent` is the N-env form of the reference's `QLearningAgent`
(QLearningBase/Agent/main.py:14-57): q_table -> 32-byte slots keyed by the packed board,
choose_action -> epsilon-greedy with Philox draws, update_q_value -> batched synchronous update
(atomic or deterministic) or the fully fused asynchronous rollout.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check
from .common import MODES, epsilon_schedule_step, init_schedule
from .env import BatchedGame2048Env, _ptr, _stream, _u8


class BatchedQLearningAgent:
    def __init__(self, total_epochs: int, action_space: int = 4, learning_rate: float = 0.1,
                 discount_factor: float = 0.9, exploration_rate: float = 1.0, exploration_min: float = 0.01, *,
                 capacity: int = 1 << 24, device: int | torch.device = 0, seed: int = 0x2048):
        if action_space != 4:
            raise ValueError("the 2048 Q-table has 4 actions per state")
        if capacity & (capacity - 1) or capacity > (1 << 31):
            raise ValueError("capacity must be a power of two <= 2^31")
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise _lib.G2048Error("BatchedQLearningAgent needs a CUDA device (no CPU fallback)")
        self.device, self.capacity, self.seed = dev, int(capacity), int(seed)
        self.lr, self.gamma, self.action_space = learning_rate, discount_factor, action_space
        init_schedule(self, total_epochs, exploration_rate, exploration_min)
        _lib.init(dev.index or 0)
        self.lib = _lib.lib()
        with torch.cuda.device(dev):
            self.table = torch.zeros(self.capacity * 4, dtype=torch.int64, device=dev)  # 32-byte slots
            self._scratch = None
            self._count = torch.zeros(1, dtype=torch.int64, device=dev)

    # ---- reference API, batched ---------------------------------------------------------------
    def choose_action(self, boards: torch.Tensor, step_idx: int, env_id_base: int = 0) -> torch.Tensor:
        """choose_action (main.py:34-38) for N states; draws = Philox(seed, env id, step_idx)."""
        n = boards.numel()
        actions = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_choose_action(_ptr(self.table), self.capacity, _ptr(boards), _ptr(actions), n,
                                               float(self.epsilon), self.seed, step_idx, env_id_base, _stream()),
                  "g2048_choose_action")
        return actions

    def _scratch_for(self, n):
        need = int(self.lib.g2048_qlearn_scratch_bytes(n))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._scratch

    def update_q_value(self, state, action, reward, next_state, done, mode: str = "deterministic") -> None:
        """update_q_value (main.py:40-43) for N transitions as one synchronous batch (float32).  "deterministic"
        applies the targets of one (state, action) in batch order (any number of collisions, long runs go to a
        warp-cooperative kernel); "atomic" skips the sort -- faster for batches with few collisions (7.8 vs 5.6 G
        updates/s), but a value that thousands of transitions hit at once is then updated at one L2 round trip per
        transition."""
        n = state.numel()
        with torch.cuda.device(self.device):
            a, d = _u8(action, self.device), _u8(done, self.device)
            r = reward.to(device=self.device, dtype=torch.float32).contiguous()
            sc = self._scratch_for(n)
            check(self.lib.g2048_qtable_update(_ptr(self.table), self.capacity, _ptr(state), _ptr(a), _ptr(r),
                                               _ptr(next_state), _ptr(d), n, self.lr, self.gamma, MODES[mode], _ptr(sc),
                                               sc.numel(), _stream()), "g2048_qtable_update")

    def apply_targets(self, keys, actions, targets, mode: str = "deterministic", lr: float | None = None) -> None:
        """Q[key][a] <- Q + lr (target - Q) for (state, action, target) records, e.g. gathered from all ranks."""
        n = keys.numel()
        with torch.cuda.device(self.device):
            sc = self._scratch_for(n)
            check(self.lib.g2048_qtable_apply_targets(_ptr(self.table), self.capacity, _ptr(keys), _ptr(actions),
                                                      _ptr(targets), n, self.lr if lr is None else lr, MODES[mode],
                                                      _ptr(sc), sc.numel(), _stream()), "g2048_qtable_apply_targets")

    def decay_exploration(self, current_epoch: int) -> float:
        return epsilon_schedule_step(self, current_epoch)

    def q_values(self, boards: torch.Tensor, insert: bool = False):
        """q_table[state] for N states -> (rows float32[N,4], found bool[N])."""
        n = boards.numel()
        rows = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        found = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_qtable_lookup(_ptr(self.table), self.capacity, _ptr(boards), n, _ptr(rows), _ptr(found),
                                               int(insert), _stream()), "g2048_qtable_lookup")
        return rows, found != 0

    # ---- fused / synchronous training steps ------------------------------------------------------
    def rollout(self, env: BatchedGame2048Env, k_steps: int) -> torch.Tensor:
        """k_steps of the loop main.py:91-101 for every env in one kernel (asynchronous atomic updates)."""
        with torch.cuda.device(self.device):
            env.counters.zero_()
            check(self.lib.g2048_rollout_qlearn(_ptr(env.boards), _ptr(env.aux), _ptr(env.score), _ptr(self.table),
                                                self.capacity, env.n, k_steps, env.flavour, self.lr, self.gamma,
                                                float(self.epsilon), env.seed, env.step_idx, env.env_id_base,
                                                _ptr(env.counters), _stream()), "g2048_rollout_qlearn")
        env.step_idx += k_steps
        return env.counters

    def step_sync(self, env: BatchedGame2048Env, mode: str = "deterministic", apply: bool = True, records: bool = False):
        """One synchronous batched step; optionally returns the (key, action, target) records."""
        n = env.n
        with torch.cuda.device(self.device):
            rk = torch.empty(n, dtype=torch.int64, device=self.device) if records else None
            ra = torch.empty(n, dtype=torch.uint8, device=self.device) if records else None
            rd = torch.empty(n, dtype=torch.float32, device=self.device) if records else None
            sc = self._scratch_for(n)
            check(self.lib.g2048_qlearn_step(_ptr(env.boards), _ptr(env.aux), _ptr(env.score), _ptr(self.table),
                                             self.capacity, n, env.flavour, self.lr, self.gamma, float(self.epsilon),
                                             MODES[mode], int(apply), env.seed, env.step_idx, env.env_id_base,
                                             _ptr(env.counters), _ptr(rk), _ptr(ra), _ptr(rd), _ptr(sc), sc.numel(),
                                             _stream()), "g2048_qlearn_step")
        env.step_idx += 1
        return (rk, ra, rd) if records else None

    def emit_records(self, env: BatchedGame2048Env, records) -> None:
        """Synchronous step, exchange form: advance every env and write its packed 16-byte (state, action, target)
        record to `records` (an int64[n, 2] tensor or a raw device pointer, e.g. NVLink peer memory); the table's
        values stay untouched until `apply_records`."""
        ptr = records if isinstance(records, int) else _ptr(records)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_qlearn_emit(_ptr(env.boards), _ptr(env.aux), _ptr(env.score), _ptr(self.table),
                                             self.capacity, env.n, env.flavour, self.gamma, float(self.epsilon), env.seed,
                                             env.step_idx, env.env_id_base, _ptr(env.counters), ptr, _stream()),
                  "g2048_qlearn_emit")
        env.step_idx += 1

    def apply_records(self, lists, counts, mode: str = "deterministic", lr: float | None = None) -> None:
        """Apply record lists in order (tensors or raw device pointers, local or peer memory): the gather of the
        exchange step is fused into the kernel that looks the states up."""
        import ctypes
        ptrs = [(x if isinstance(x, int) else _ptr(x)) or 0 for x in lists]
        m = len(ptrs)
        arr_p = (ctypes.c_void_p * m)(*ptrs)
        arr_n = (ctypes.c_int64 * m)(*[int(c) for c in counts])
        n = int(sum(counts))
        with torch.cuda.device(self.device):
            sc = self._scratch_for(max(n, 1))
            check(self.lib.g2048_qtable_apply_records(_ptr(self.table), self.capacity, arr_p, arr_n, m,
                                                      self.lr if lr is None else lr, MODES[mode], _ptr(sc), sc.numel(),
                                                      _stream()), "g2048_qtable_apply_records")

    # ---- table management ------------------------------------------------------------------------
    def clear(self):
        self.table.zero_()

    def __len__(self) -> int:
        with torch.cuda.device(self.device):
            check(self.lib.g2048_qtable_size(_ptr(self.table), self.capacity, _ptr(self._count), _stream()),
                  "g2048_qtable_size")
        return int(self._count.item())

    def probe_stats(self) -> dict:
        """{"states", "mean_probe_length", "max_probe_length", "load_factor"} of the table as it is now."""
        st = torch.zeros(3, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_qtable_probe_stats(_ptr(self.table), self.capacity, _ptr(st), _stream()),
                  "g2048_qtable_probe_stats")
        n, total, mx = (int(x) for x in st.tolist())
        return {"states": n, "mean_probe_length": 1 + total / max(n, 1), "max_probe_length": 1 + mx,
                "load_factor": n / self.capacity}

    def export(self):
        """(keys uint64[n] as int64 tensor, rows float32[n,4]) sorted by key."""
        n = len(self)
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        rows = torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._count.zero_()
            check(self.lib.g2048_qtable_export(_ptr(self.table), self.capacity, _ptr(keys), _ptr(rows), n,
                                               _ptr(self._count), _stream()), "g2048_qtable_export")
        k = keys[:n].cpu().numpy().view(np.uint64)
        order = np.argsort(k)
        return k[order], rows[:n].cpu().numpy()[order]

    def to_dict(self) -> dict:
        """The reference's q_table form: {tuple-of-tuples of raw tile values: np.ndarray[4]} (main.py:16, :82)."""
        keys, rows = self.export()
        out = {}
        for k, row in zip(keys.tolist(), rows):
            cells = [(k >> (4 * j)) & 15 for j in range(16)]
            tiles = [(1 << c) if c else 0 for c in cells]
            out[tuple(tuple(tiles[4 * r:4 * r + 4]) for r in range(4))] = row.astype(np.float64)
        return out

    def save(self, path: str) -> None:
        keys, rows = self.export()
        torch.save({"keys": torch.from_numpy(keys.view(np.int64)), "rows": torch.from_numpy(rows),
                    "capacity": self.capacity, "epsilon": self.epsilon, "lr": self.lr, "gamma": self.gamma}, path)

    def load(self, path: str) -> None:
        blob = torch.load(path)
        self.clear()
        keys = blob["keys"].to(self.device)
        rows = blob["rows"].to(self.device)
        self.epsilon = blob["epsilon"]
        for a in range(4):  # lr = 1 on an empty table: Q[key][a] <- row[a]
            self.apply_targets(keys, torch.full((keys.numel(),), a, dtype=torch.uint8, device=self.device),
                               rows[:, a].contiguous(), mode="atomic", lr=1.0)
