"""Batched 2048 environment on the GPU (device-pointer API of libg2048.so).

`BatchedGame2048Env` is the N-env form of the reference's `Game2048_env`
(QLearningBase/environment/Game2048_env.py:78-205, flavour "penalty";
Deep_QLearning/environment/Game2048_nopenalty_env.py:81-150 with the caller-commit protocol of
mainDQL_CNN_step2.py:163-237, flavour "nopenalty").  PyTorch is only the owner of device memory and of
the CUDA stream; every computation is a kernel of libg2048.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check

from .common import AUX_INIT, COUNTER_NAMES, FLAVOURS, MODES, N_COUNTERS, ActionSpace, ObservationSpace  # noqa: F401


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _u8(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, np.uint8))
    return x.to(device=device, dtype=torch.uint8).contiguous()


def boards_to_numpy(boards: torch.Tensor) -> np.ndarray:
    """int64-typed device tensor of packed boards -> np.uint64 array."""
    return boards.detach().cpu().numpy().view(np.uint64)


def boards_from_numpy(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, np.uint64).view(np.int64)).to(device)


class BatchedGame2048Env:
    """N independent 2048 games, one packed uint64 board each, resident in HBM.

    reset() -> boards;  step(actions) -> (boards, reward, done, max_number)  -- the reference's 4-tuple
    (Game2048_env.py:129), as tensors of length N.  Boards are int64-typed tensors holding the packed
    uint64 (cell (r,c) = nibble 4r+c = log2(tile)); `tiles()` gives the reference's (N,4,4) int64 arrays.
    Random draws come from Philox keyed by (seed, global env id, step), so results do not depend on how
    envs are sharded over GPUs.
    """

    action_space = ActionSpace()
    observation_space = ObservationSpace()

    def __init__(self, n_envs: int, flavour: str = "penalty", device: int | torch.device = 0, seed: int = 0x2048,
                 env_id_base: int = 0, auto_reset: bool = False):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise _lib.G2048Error("BatchedGame2048Env needs a CUDA device (no CPU fallback)")
        self.device, self.n = dev, int(n_envs)
        self.flavour = FLAVOURS[flavour]
        self.seed, self.env_id_base, self.auto_reset = int(seed), int(env_id_base), auto_reset
        _lib.init(dev.index or 0)
        self.lib = _lib.lib()
        with torch.cuda.device(dev):
            self.boards = torch.zeros(self.n, dtype=torch.int64, device=dev)
            self.aux = torch.full((self.n,), AUX_INIT, dtype=torch.int64, device=dev)
            self.score = torch.zeros(self.n, dtype=torch.int32, device=dev)
            self.reward = torch.zeros(self.n, dtype=torch.float64, device=dev)
            self.flags = torch.zeros(self.n, dtype=torch.uint8, device=dev)
            self.maxlvl = torch.zeros(self.n, dtype=torch.uint8, device=dev)
            self.move_score = torch.zeros(self.n, dtype=torch.int32, device=dev)
            self.counters = torch.zeros(N_COUNTERS, dtype=torch.int64, device=dev)
        self.step_idx = 0
        self.episode_idx = 0

    # ---- reference API, batched ---------------------------------------------------------------
    def reset(self, mask: torch.Tensor | None = None, replay_draws=None) -> torch.Tensor:
        """Game2048_env.reset (Game2048_env.py:187-191) for all envs, or those where mask != 0."""
        with torch.cuda.device(self.device):
            m = None if mask is None else _u8(mask, self.device)
            d = None if replay_draws is None else _u8(replay_draws, self.device)
            check(self.lib.g2048_env_reset(_ptr(self.boards), _ptr(self.score), _ptr(m), _ptr(d), self.n, self.seed,
                                           self.episode_idx, self.env_id_base, _stream()), "g2048_env_reset")
        self.episode_idx += 1
        return self.boards

    def step(self, actions, replay_draws=None):
        """Game2048_env.step (Game2048_env.py:97-129 / Game2048_nopenalty_env.py:106-120).

        Returns (boards, reward float64[N], done bool[N], max_number int64[N]); `self.flags` additionally
        holds valid / game_over / done bits and the legal-move mask of the new boards."""
        with torch.cuda.device(self.device):
            a = _u8(actions, self.device)
            d = None if replay_draws is None else _u8(replay_draws, self.device)
            check(self.lib.g2048_env_step(_ptr(self.boards), _ptr(self.aux), _ptr(self.score), _ptr(a), _ptr(d),
                                          _ptr(self.reward), None, _ptr(self.flags), _ptr(self.maxlvl),
                                          _ptr(self.move_score), self.n, self.flavour, self.seed, self.step_idx,
                                          self.env_id_base, _stream()), "g2048_env_step")
            self.step_idx += 1
            done = (self.flags & 4) != 0
            max_number = torch.ones_like(self.boards) << self.maxlvl.to(torch.int64)
            out = (self.boards, self.reward, done, max_number)
            if self.auto_reset and replay_draws is None:
                out = (self.boards.clone(), self.reward, done, max_number)
                self.reset(mask=done)
        return out

    @property
    def valid(self):
        return (self.flags & 1) != 0

    @property
    def game_over(self):
        return (self.flags & 2) != 0

    def legal_mask(self, boards: torch.Tensor | None = None) -> torch.Tensor:
        """4-bit mask per env: bit a set iff game.move(a, trial=True) moves (mainDQL_CNN_step2.py:169-174)."""
        b = self.boards if boards is None else boards
        out = torch.empty(b.numel(), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_legal_mask(_ptr(b), _ptr(out), b.numel(), _stream()), "g2048_legal_mask")
        return out

    def move_trial(self, actions, boards: torch.Tensor | None = None):
        """game.move(a, trial=True) -> (moved boards, moved bool, score)."""
        b = self.boards if boards is None else boards
        n = b.numel()
        with torch.cuda.device(self.device):
            a = _u8(actions, self.device)
            out = torch.empty_like(b)
            moved = torch.empty(n, dtype=torch.uint8, device=self.device)
            sc = torch.empty(n, dtype=torch.int32, device=self.device)
            check(self.lib.g2048_move_trial(_ptr(b), _ptr(a), _ptr(out), _ptr(moved), _ptr(sc), n, _stream()),
                  "g2048_move_trial")
        return out, moved != 0, sc

    def tiles(self, boards: torch.Tensor | None = None) -> torch.Tensor:
        """(N,4,4) int64 raw tile values -- the reference's board arrays."""
        b = self.boards if boards is None else boards
        out = torch.empty((b.numel(), 4, 4), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_unpack_i64(_ptr(b), _ptr(out), b.numel(), _stream()), "g2048_unpack_i64")
        return out

    def set_tiles(self, tiles: torch.Tensor) -> None:
        """env.game.board = tiles (the assignable board of main.py:85 / mainDQL_CNN_step2.py:237)."""
        t = tiles.to(device=self.device, dtype=torch.int64).contiguous().view(-1, 16)
        bad = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_pack_i64(_ptr(t), _ptr(self.boards), t.shape[0], _ptr(bad), _stream()), "g2048_pack_i64")
        if int(bad.item()):
            raise ValueError(f"{int(bad.item())} cells are not 0 or a power of two in 2..32768")

    def encode_onehot(self, boards: torch.Tensor | None = None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """DQNAgent.encode_state (Dqn8TestNOPERCNN.py:271-277): (N,16,4,4) [batch, level, row, col]."""
        b = self.boards if boards is None else boards
        code = {torch.float32: 0, torch.bfloat16: 1}[dtype]
        out = torch.empty((b.numel(), 16, 4, 4), dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_encode_onehot(_ptr(b), _ptr(out), b.numel(), code, _stream()), "g2048_encode_onehot")
        return out

    # ---- fused rollouts -----------------------------------------------------------------------
    def rollout_random(self, k_steps: int) -> dict:
        """k_steps uniform-random-policy steps per env with in-kernel reset; boards stay in registers."""
        with torch.cuda.device(self.device):
            self.counters.zero_()
            check(self.lib.g2048_rollout_random(_ptr(self.boards), _ptr(self.aux), _ptr(self.score), self.n, k_steps,
                                                self.flavour, self.seed, self.step_idx, self.env_id_base,
                                                _ptr(self.counters), _stream()), "g2048_rollout_random")
        self.step_idx += k_steps
        return self.counters

    def counters_dict(self, counters: torch.Tensor | None = None) -> dict:
        c = (self.counters if counters is None else counters).cpu().tolist()
        return dict(zip(COUNTER_NAMES, c))

    def showMatrix(self, i: int = 0):
        print(int(self.score[i]))
        print(self.tiles(self.boards[i:i + 1])[0].cpu().numpy())
