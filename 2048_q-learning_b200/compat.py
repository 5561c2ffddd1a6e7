"""N = 1 adapters with the reference's exact duck-typed API, running on the GPU library.

    from g2048 import Game2048_env, QLearningAgent     # instead of the reference's modules
    env = Game2048_env(); agent = QLearningAgent(num_episodes, action_space=env.action_space.n)
    ... the loop of QLearningBase/Agent/main.py:80-109 runs unchanged ...

Interfaces mirrored:
  Game2048 / Game2048_env   QLearningBase/environment/Game2048_env.py:10-205   (flavour="penalty")
                            Deep_QLearning/environment/Game2048_nopenalty_env.py:10-150 (flavour="nopenalty")
  QLearningAgent            QLearningBase/Agent/main.py:14-57

RNG: with rng="numpy" (default) the adapters draw from the global `np.random` / `random` generators with
the same calls in the same order as the reference (including the phantom spawn draws inside
is_game_over, Game2048_env.py:69-74) and hand the draws to the kernels as replay input -- so under the
same `np.random.seed` / `random.seed` they reproduce the reference's trajectories bit for bit.
rng="philox" uses the library's counter-based generator instead.

These adapters exist for drop-in compatibility and parity checks; each step is a few tiny kernel
launches.  Throughput comes from the batched classes (env.py / agent.py).  Host code here never
computes game logic: moves, merges, rewards and Q arithmetic all run in libg2048.so.
"""
from __future__ import annotations

import collections
import collections.abc
import ctypes as C
import random as _pyrandom

import numpy as np

from . import _lib
from ._lib import check
from .common import AUX_INIT, FLAVOURS, ActionSpace, ObservationSpace, epsilon_schedule_step, init_schedule

_PEN = [-1.0]
for _ in range(25):  # Game2048_env.py:124: penalty = max(last * 1.1, -10)
    _PEN.append(max(_PEN[-1] * 1.1, -10))


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_tiles(tiles) -> int:
    """(4,4) raw tile values -> packed uint64 (cell (r,c) = nibble 4r+c = log2(tile))."""
    b = 0
    for j, v in enumerate(np.asarray(tiles).reshape(16).tolist()):
        v = int(v)
        lvl = v.bit_length() - 1 if v else 0
        if v and ((1 << lvl) != v or not 1 <= lvl <= 15):
            raise ValueError(f"tile {v} is not a power of two in 2..32768")
        b |= lvl << (4 * j)
    return b


def unpack_tiles(b: int) -> np.ndarray:
    cells = [(int(b) >> (4 * j)) & 15 for j in range(16)]
    return np.array([(1 << c) if c else 0 for c in cells], dtype=np.int64).reshape(4, 4)


class _Backend:
    """One shared host-buffer context per (device, table capacity)."""

    _cache: dict = {}

    def __init__(self, device: int, capacity: int):
        _lib.init(device)
        self.lib = _lib.lib()
        self.ctx = self.lib.g2048_ctx_create(device, 16, capacity)
        if not self.ctx:
            raise _lib.G2048Error("g2048_ctx_create failed: " + self.lib.g2048_last_error().decode())

    @classmethod
    def get(cls, device: int, capacity: int = 0) -> "_Backend":
        key = (device, capacity)
        if capacity or key not in cls._cache:
            be = cls(device, capacity)
            if capacity:
                return be
            cls._cache[key] = be
        return cls._cache[key]

    # thin wrappers on single-element numpy buffers
    def move_trial(self, board: int, action: int):
        b, a = np.array([board], np.uint64), np.array([action], np.uint8)
        out, moved, sc = np.zeros(1, np.uint64), np.zeros(1, np.uint8), np.zeros(1, np.int32)
        check(self.lib.g2048_ctx_move_trial(self.ctx, _vp(b), _vp(a), _vp(out), _vp(moved), _vp(sc), 1), "move_trial")
        return int(out[0]), bool(moved[0]), int(sc[0])

    def legal_mask(self, board: int) -> int:
        b, out = np.array([board], np.uint64), np.zeros(1, np.uint8)
        check(self.lib.g2048_ctx_legal_mask(self.ctx, _vp(b), _vp(out), 1), "legal_mask")
        return int(out[0])

    def env_step(self, board, aux, score, action, draws, flavour, seed, step_idx):
        b, x, s = np.array([board], np.uint64), np.array([aux], np.uint64), np.array([score], np.int32)
        a = np.array([action], np.uint8)
        d = None if draws is None else np.array([draws], np.uint8)
        r, f, m, ms = np.zeros(1, np.float64), np.zeros(1, np.uint8), np.zeros(1, np.uint8), np.zeros(1, np.int32)
        check(self.lib.g2048_ctx_env_step(self.ctx, _vp(b), _vp(x), _vp(s), _vp(a), _vp(d), _vp(r), _vp(f), _vp(m),
                                          _vp(ms), 1, flavour, seed, step_idx, 0), "env_step")
        return int(b[0]), int(x[0]), int(s[0]), float(r[0]), int(f[0]), int(m[0]), int(ms[0])

    def env_reset(self, draws, seed, episode_idx):
        b = np.zeros(1, np.uint64)
        d = None if draws is None else np.array([draws], np.uint8)
        check(self.lib.g2048_ctx_env_reset(self.ctx, _vp(b), None, None, _vp(d), 1, seed, episode_idx, 0), "env_reset")
        return int(b[0])


def _n_empty(board: int) -> int:
    return sum(1 for j in range(16) if not (board >> (4 * j)) & 15)


class Game2048:
    """Game core adapter (Game2048_env.py:10-75 / Game2048_nopenalty_env.py:10-78): `.board` is a (4,4) int64
    array of raw tile values, readable and assignable; `move`, `add_number`, `is_game_over` as in the reference."""

    def __init__(self, flavour="penalty", rng="numpy", device=0, seed=0x2048, _episode=0):
        self._flavour, self._rng, self._seed = flavour, rng, seed
        self._be = _Backend.get(device)
        self._draw_idx = _episode << 20
        if rng == "numpy":
            draws = self._spawn_draw(16) + self._spawn_draw(15)
            packed = self._be.env_reset(draws, 0, 0)
        else:
            packed = self._be.env_reset(None, seed, _episode)
        self.board = unpack_tiles(packed)
        if flavour == "nopenalty":
            self.moved_board = np.zeros((4, 4), dtype=np.int64)

    def _spawn_draw(self, n_empty):
        """The two draws of add_number (Game2048_env.py:19-20) from the global np.random stream."""
        k = int(np.random.randint(0, n_empty))
        return [k, 0 if np.random.random() < 0.9 else 1]

    def _spawn_into(self, packed: int) -> int:
        n = _n_empty(packed)
        if n == 0:
            return packed
        if self._rng == "numpy":
            k, is4 = self._spawn_draw(n)
        else:
            x = np.random.Generator(np.random.Philox(key=self._seed, counter=self._draw_idx)).integers(0, 1 << 32, 2)
            self._draw_idx += 1
            k, is4 = int(x[0]) * n >> 32, int(int(x[1]) >= 0xE6666666)
        empties = [j for j in range(16) if not (packed >> (4 * j)) & 15]
        return packed | ((2 if is4 else 1) << (4 * empties[k]))

    def add_number(self, board=None):
        target = self.board if board is None else board
        target[...] = unpack_tiles(self._spawn_into(pack_tiles(target)))

    def move(self, action, trial=False):
        """(moved, score); penalty flavour mutates `.board`, nopenalty leaves the result in `.moved_board`."""
        out, moved, score = self._be.move_trial(pack_tiles(self.board), int(action))
        if self._flavour == "penalty":
            if moved:
                out = self._spawn_into(out)
            self.board = unpack_tiles(out)
        else:
            if moved and not trial:
                out = self._spawn_into(out)
            # trial moves leave moved_board as the un-moved copy (move_left only writes when not trial, :46-47)
            self.moved_board = unpack_tiles(out if not trial else pack_tiles(self.board))
        return moved, np.int64(score)

    def is_game_over(self):
        packed = pack_tiles(self.board)
        lm = self._be.legal_mask(packed)
        if _n_empty(packed):
            return False
        if lm == 0:
            if self._flavour == "nopenalty":
                self.moved_board = self.board.copy()
            return True
        # the reference performs the first legal move with a real spawn and restores the board (:69-74);
        # only the RNG stream (and, nopenalty, moved_board) keeps a trace of it
        a = (lm & -lm).bit_length() - 1
        out, _, _ = self._be.move_trial(packed, a)
        out = self._spawn_into(out)
        if self._flavour == "nopenalty":
            self.moved_board = unpack_tiles(out)
        return False


class Game2048_env:
    """Gym-style env adapter: reset() -> board; step(action) -> (board, reward, done, max_number)
    (old-gym 4-tuple, Game2048_env.py:129).  All game logic, the shaped reward and the done rule run in
    k_env_step; this class only moves 8-byte boards and the reference's RNG draws across the boundary."""

    rewards_buffer = collections.deque()
    iter = 0

    def __init__(self, flavour="penalty", rng="numpy", device=0, seed=0x2048):
        if flavour not in FLAVOURS:
            raise ValueError(flavour)
        self._flavour, self._rng, self._device, self._seed = flavour, rng, device, seed
        self._be = _Backend.get(device)
        self._episode = 0
        self._step_idx = 0
        self._aux = AUX_INIT
        self.action_space = ActionSpace()
        self.observation_space = ObservationSpace()
        self.game = Game2048(flavour, rng, device, seed, self._episode)
        self.score = 0
        self.move_score = 0
        self.penalty = 10
        self.scaling_factor = 1.2
        self.max_consecutive_actions = 10
        self._publish_aux()
        if flavour == "nopenalty":
            self.prev_max_tile = 2
            self.max_number = 0

    def _publish_aux(self):
        aux = self._aux
        self.previous_max = 1 << (aux & 0xFF)
        ca = (aux >> 8) & 0xFF
        self.consecutive_action = None if ca == 0xFF else ca
        self.last_consecutive_penalty = _PEN[(aux >> 16) & 0xFF]
        self.consecutive_count = aux >> 32

    def reset(self):
        self._episode += 1
        self.game = Game2048(self._flavour, self._rng, self._device, self._seed, self._episode)
        self.score = 0
        if self._flavour == "nopenalty":
            self.prev_max_tile = 2
        return self.game.board

    def step(self, action):
        action = int(action)
        g, be = self.game, self._be
        S = pack_tiles(g.board)
        draws = None
        if self._rng == "numpy":  # reproduce the reference's np.random call sequence, then replay it on the GPU
            draws = [255, 255, 255, 255]
            out, moved, _ = be.move_trial(S, action)
            if moved:
                draws[0], draws[1] = g._spawn_draw(_n_empty(out))
            if self._flavour == "nopenalty" and _n_empty(S) == 0:
                lm = be.legal_mask(S)
                if lm:
                    out2, _, _ = be.move_trial(S, (lm & -lm).bit_length() - 1)
                    draws[2], draws[3] = g._spawn_draw(_n_empty(out2))
        code = FLAVOURS[self._flavour]
        b, aux, score, reward, flags, maxlvl, ms = be.env_step(S, self._aux, self.score, action, draws, code, self._seed,
                                                               self._step_idx)
        self._step_idx += 1
        self._aux, self.score, self.move_score = aux, score, ms
        done = bool(flags & 4)
        if self._flavour == "penalty":
            g.board = unpack_tiles(b)
            self._publish_aux()
            if self._rng == "numpy" and _n_empty(b) == 0 and not (flags & 2):
                lm = flags >> 4  # phantom spawn draws of is_game_over (Game2048_env.py:69-74)
                out3, _, _ = be.move_trial(b, (lm & -lm).bit_length() - 1)
                g._spawn_draw(_n_empty(out3))
            return g.board, reward, done, np.int64(1 << maxlvl)
        g.moved_board = unpack_tiles(b)  # the caller commits: env.game.board = next_state (mainDQL_CNN_step2.py:237)
        return g.moved_board, int(reward), done, np.int64(1 << maxlvl)

    def showMatrix(self):
        print(self.score)
        print(self.game.board)


class _QTableView(collections.abc.Mapping):
    """`agent.q_table` facade: state (tuple of tuples of raw tiles) -> np.ndarray[4]; reading inserts a zero
    row like the reference's defaultdict (main.py:16)."""

    def __init__(self, agent):
        self._a = agent

    def __getitem__(self, state):
        return self._a._row(pack_tiles(state), insert=True)

    def __len__(self):
        return int(self._a._be.lib.g2048_ctx_qtable_size(self._a._be.ctx))

    def __iter__(self):
        return iter(self._a.to_dict())

    def items(self):
        return self._a.to_dict().items()


class QLearningAgent:
    """QLearningAgent adapter (main.py:14-57); the Q-table lives in HBM, arithmetic is float32 on the GPU."""

    def __init__(self, total_epochs, action_space, learning_rate=0.1, discount_factor=0.9, exploration_rate=1.0,
                 exploration_min=0.01, *, capacity=1 << 22, device=0, rng="python", seed=0x2048):
        if action_space != 4:
            raise ValueError("the 2048 Q-table has 4 actions per state")
        self.lr, self.gamma, self.action_space = learning_rate, discount_factor, action_space
        init_schedule(self, total_epochs, exploration_rate, exploration_min)
        self._be = _Backend.get(device, capacity)
        self._rng, self._seed, self._step = rng, seed, 0
        self.q_table = _QTableView(self)

    def _row(self, key: int, insert: bool) -> np.ndarray:
        k, rows, found = np.array([key], np.uint64), np.zeros((1, 4), np.float32), np.zeros(1, np.uint8)
        check(self._be.lib.g2048_ctx_qtable_lookup(self._be.ctx, _vp(k), 1, _vp(rows), _vp(found), int(insert)), "lookup")
        return rows[0].astype(np.float64)

    def choose_action(self, state):
        if self._rng == "python":  # same draws, same order as main.py:35-36
            if _pyrandom.random() < self.epsilon:
                return _pyrandom.randint(0, self.action_space - 1)
            b, a = np.array([pack_tiles(state)], np.uint64), np.zeros(1, np.uint8)
            check(self._be.lib.g2048_ctx_choose_action(self._be.ctx, _vp(b), _vp(a), 1, 0.0, 0, 0, 0), "choose_action")
            return int(a[0])
        b, a = np.array([pack_tiles(state)], np.uint64), np.zeros(1, np.uint8)
        check(self._be.lib.g2048_ctx_choose_action(self._be.ctx, _vp(b), _vp(a), 1, float(self.epsilon), self._seed,
                                                   self._step, 0), "choose_action")
        self._step += 1
        return int(a[0])

    def update_q_value(self, state, action, reward, next_state, done):
        s, s2 = np.array([pack_tiles(state)], np.uint64), np.array([pack_tiles(next_state)], np.uint64)
        a, d = np.array([int(action)], np.uint8), np.array([1 if done else 0], np.uint8)
        r = np.array([reward], np.float32)
        # one record: the atomic path is the same arithmetic as the deterministic one, without the sort launches
        check(self._be.lib.g2048_ctx_qtable_update(self._be.ctx, _vp(s), _vp(a), _vp(r), _vp(s2), _vp(d), 1,
                                                   float(self.lr), float(self.gamma), 0), "qtable_update")

    def decay_exploration(self, current_epoch):
        epsilon_schedule_step(self, current_epoch)

    def export(self):
        n = int(self._be.lib.g2048_ctx_qtable_size(self._be.ctx))
        keys, rows = np.zeros(max(n, 1), np.uint64), np.zeros((max(n, 1), 4), np.float32)
        got = self._be.lib.g2048_ctx_qtable_export(self._be.ctx, _vp(keys), _vp(rows), n)
        if got < 0:
            raise _lib.G2048Error("qtable_export failed")
        order = np.argsort(keys[:n])
        return keys[:n][order], rows[:n][order]

    def to_dict(self):
        keys, rows = self.export()
        return {tuple(map(tuple, unpack_tiles(int(k)).tolist())): row.astype(np.float64) for k, row in zip(keys, rows)}
