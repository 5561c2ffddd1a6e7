"""DQN agent fed by the batched GPU env (SURVEY.md section 8f "next" #1, BASELINE config 5).

PyTorch restatement of Deep_QLearning/main_dir/Dqn8TestNOPERCNN.py: the network stays on PyTorch/cuBLAS/cuDNN
(north_star); what is B200-native here is everything around it -- boards stay packed (8 bytes) in the replay
memory and are one-hot encoded on the fly by `k_onehot_*`, actions for all envs come from `k_select_action`
(act / act_ripetitive with the legal-move mask the env step already produced), and transitions never leave HBM.

Reference pieces mirrored
  DQNModel            :202-246  three blocks of four parallel Conv2D(512, k in 1..4, 'same') + concat + ReLU,
                                Flatten, Dense 1024 ReLU, Dropout 0.5, Dense 4; Adam 5e-5; 197,204,996 parameters.
                                Keras reads Input(16,4,4) channels-last: H = level, W = row, C = col.
  DQNAgent            :248-400  same constructor arguments; encode_state :271-277, act :312-324,
                                act_ripetitive :326-336, update_epsilon :341-343, remember, replay :351-390
                                (loss = mean over B x 4 of (target - q)^2 with target == q except at the taken
                                action; terminal target = reward), update_target_model :338-339.
Differences from the reference (documented; DESIGN.md "Known divergences"):
  * the replay memory stores next_state explicitly instead of keras-rl's "next stored observation" (:48-66), which for
    one env is the same board except for the newest entry (all zeros there until the next append);
  * by default the drivers below restrict every action to the legal moves and store every transition.  The
    reference's own rule -- act() unrestricted, so invalid moves ARE played (reward -10) and stored;
    act_ripetitive() only when the previous transition was not stored; remember() drops a transition that repeats
    the env's previously stored (state, next_state) unless the game ended (:279-297, mainDQL_CNN_step2.py:176-185,
    :220) -- is `dqn_step(..., reference_driver=True)`: the duplicate filter and the `memory_saved` bit are kept per
    env (the reference has one env, so "the previous entry of the memory" is that env's);
  * prioritisation is alpha = 0 (uniform) as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import check
from .env import BatchedGame2048Env, _ptr, _stream


class ConvBlock(nn.Module):
    """conv_block (:231-246): four parallel 'same' convolutions with kernel sizes 1..4, concatenated, ReLU."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        d = c_out // 4
        self.convs = nn.ModuleList([nn.Conv2d(c_in, d, kernel_size=k, padding="same") for k in (1, 2, 3, 4)])

    def forward(self, x):
        return F.relu(torch.cat([c(x) for c in self.convs], dim=1))


class DQNModel(nn.Module):
    """_build_model (:209-229).  Input: (N, 16, 4, 4) one-hot [batch, level, row, col] as produced by encode_state."""

    def __init__(self, action_space: int = 4, width: int = 2048, hidden: int = 1024):
        super().__init__()
        self.blocks = nn.Sequential(ConvBlock(4, width), ConvBlock(width, width), ConvBlock(width, width))
        self.fc1 = nn.Linear(16 * 4 * width, hidden)
        self.drop = nn.Dropout(0.5)
        self.fc2 = nn.Linear(hidden, action_space)

    def forward(self, x):
        # Keras channels-last reading of (16,4,4): H = level, W = row, C = col  ->  torch (N, C=col, H=level, W=row)
        x = x.permute(0, 3, 1, 2)
        x = self.blocks(x)
        x = x.permute(0, 2, 3, 1).flatten(1)          # Keras Flatten order (h, w, c)
        return self.fc2(self.drop(F.relu(self.fc1(x))))


class BatchedDQNAgent:
    """DQNAgent (:248-400) for N envs at once."""

    def __init__(self, state_shape=(16, 4, 4), action_space=4, gamma=0.99, decay_episodes=200, epsilon=0.9,
                 epsilon_min=0.001, epsilon_decay=0.9999, batch_size=64, memory_size=50000, alpha=0.0, beta=1.0,
                 beta_increment=1e-5, *, device=0, width=2048, hidden=1024, dtype=torch.float32, seed=0x2048,
                 learning_rate=5e-5):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise _lib.G2048Error("BatchedDQNAgent needs a CUDA device (no CPU fallback)")
        _lib.init(dev.index or 0)
        self.lib, self.device, self.dtype, self.seed = _lib.lib(), dev, dtype, int(seed)
        self.state_shape, self.action_space, self.gamma = state_shape, action_space, gamma
        self.epsilon = self.epsilon_start = epsilon
        self.epsilon_min, self.epsilon_decay, self.decay_episodes = epsilon_min, epsilon_decay, decay_episodes
        self.batch_size, self.memory_size, self.alpha, self.beta, self.beta_increment = (batch_size, memory_size, alpha,
                                                                                         beta, beta_increment)
        self.model = DQNModel(action_space, width, hidden).to(dev, dtype)
        self.target_model = DQNModel(action_space, width, hidden).to(dev, dtype)
        self.update_target_model()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=learning_rate)
        self.step_counter = 0
        self.grad_sync = None
        self.loss_history: list[float] = []
        # replay memory: packed boards, never one-hot (16 B per transition instead of 2 KB)
        m = memory_size
        self.mem_state = torch.zeros(m, dtype=torch.int64, device=dev)
        self.mem_next = torch.zeros(m, dtype=torch.int64, device=dev)
        self.mem_action = torch.zeros(m, dtype=torch.int64, device=dev)
        self.mem_reward = torch.zeros(m, dtype=torch.float32, device=dev)
        self.mem_done = torch.zeros(m, dtype=torch.bool, device=dev)
        self.nb_entries, self._head = 0, 0
        # reference driver (dqn_step(reference_driver=True)): per env the last stored (state, next_state) and whether the
        # previous transition was stored (`memory_saved`, mainDQL_CNN_step2.py:183-185, :220)
        self._last_state = self._last_next = self.memory_saved = None
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(self.seed)

    # ---- reference API ---------------------------------------------------------------------------------------
    def encode_state(self, boards: torch.Tensor) -> torch.Tensor:
        """encode_state (:271-277) on packed boards: (N, 16, 4, 4) one-hot in the model's dtype."""
        code = {torch.float32: 0, torch.bfloat16: 1}[self.dtype]
        n = boards.numel()
        out = torch.empty((n, 16, 4, 4), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_encode_onehot(_ptr(boards), _ptr(out), n, code, _stream()), "g2048_encode_onehot")
        return out

    def update_epsilon(self):
        self.epsilon = max(self.epsilon_min, self.epsilon_start * (self.epsilon_decay ** self.step_counter))

    @torch.no_grad()
    def q_values(self, boards: torch.Tensor) -> torch.Tensor:
        self.model.eval()
        return self.model(self.encode_state(boards)).float().contiguous()

    def _select(self, boards, legal_mask, env_id_base):
        self.update_epsilon()
        self.step_counter += 1
        q = self.q_values(boards)
        n = boards.numel()
        actions = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.g2048_select_action(_ptr(q), _ptr(legal_mask), _ptr(actions), n, float(self.epsilon),
                                               self.seed, self.step_counter, env_id_base, _stream()),
                  "g2048_select_action")
        return actions

    def act(self, boards: torch.Tensor, env_id_base: int = 0) -> torch.Tensor:
        """act (:312-324): epsilon-greedy over the four actions."""
        return self._select(boards, None, env_id_base)

    def act_ripetitive(self, boards: torch.Tensor, legal_mask: torch.Tensor, env_id_base: int = 0) -> torch.Tensor:
        """act_ripetitive (:326-336): epsilon-greedy restricted to the legal moves (4-bit mask per env)."""
        return self._select(boards, legal_mask.to(torch.uint8).contiguous(), env_id_base)

    def remember(self, state, action, reward, done, next_state, filter_duplicates: bool = False):
        """remember (:279-297) for a batch of transitions (packed boards), ring-buffer append.  With
        filter_duplicates (the reference's rule) a transition whose (state, next_state) equals the env's previously
        stored one is dropped unless the game ended; returns the per-env `memory_saved` mask (all True otherwise)."""
        n = state.numel()
        keep = torch.ones(n, dtype=torch.bool, device=self.device)
        if filter_duplicates:
            if self._last_state is None or self._last_state.numel() != n:
                self._last_state = torch.full((n,), -1, dtype=torch.int64, device=self.device)
                self._last_next = torch.full((n,), -1, dtype=torch.int64, device=self.device)
            keep = done.to(torch.bool) | (state != self._last_state) | (next_state != self._last_next)
            self._last_state = torch.where(keep, state, self._last_state)
            self._last_next = torch.where(keep, next_state, self._last_next)
            state, next_state, action, reward, done = (state[keep], next_state[keep], action[keep], reward[keep], done[keep])
            n = state.numel()
        if n > self.memory_size:                     # more transitions than slots: only the newest fit (no index repeats)
            state, next_state, action, reward, done = (t[-self.memory_size:] for t in (state, next_state, action, reward, done))
            n = self.memory_size
        if n:
            idx = (torch.arange(n, device=self.device) + self._head) % self.memory_size
            self.mem_state[idx] = state
            self.mem_next[idx] = next_state
            self.mem_action[idx] = action.to(torch.int64)
            self.mem_reward[idx] = reward.to(torch.float32)
            self.mem_done[idx] = done.to(torch.bool)
            self._head = (self._head + n) % self.memory_size
            self.nb_entries = min(self.nb_entries + n, self.memory_size)
        return keep

    def replay(self, episode=None):
        """replay (:351-390): one Adam step on a uniformly sampled minibatch."""
        if self.nb_entries < self.batch_size or self.epsilon >= 1:
            return None
        idx = torch.randint(0, self.nb_entries, (self.batch_size,), device=self.device, generator=self._gen)
        actions, rewards, done = self.mem_action[idx], self.mem_reward[idx], self.mem_done[idx]
        self.model.train()
        self.target_model.eval()
        q = self.model(self.encode_state(self.mem_state[idx])).float()
        with torch.no_grad():
            next_q = self.target_model(self.encode_state(self.mem_next[idx])).float()
            target_a = torch.where(done, rewards, rewards + self.gamma * next_q.max(dim=1).values)   # :368-372
            targets = q.detach().clone()
            targets.scatter_(1, actions.unsqueeze(1), target_a.unsqueeze(1))
        loss = torch.mean((targets - q) ** 2)                                                         # :376
        if self.grad_sync is not None:              # data parallel: gradients live in one flat all-reduce buffer
            self.grad_sync.zero_grad()
            loss.backward()
            self.grad_sync()
        else:
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
        self.optimizer.step()
        self.loss_history.append(float(loss.detach()))
        return self.loss_history[-1]

    def update_target_model(self):
        self.target_model.load_state_dict(self.model.state_dict())

    def data_parallel(self, group=None):
        """Train this agent data-parallel over the ranks of `group` (one process per GPU, envs sharded): weights
        start from rank 0's, every replay() averages the gradients with one NCCL all-reduce (dist.GradientAllReduce)."""
        from .dist import GradientAllReduce
        self.grad_sync = GradientAllReduce(self.model, group)
        self.grad_sync.sync_parameters()
        self.update_target_model()
        return self

    def change_lr_function(self, reached_1024: bool = False):
        """:299-310: lr <- max(lr * 0.98, 1e-6) whenever an episode ended with a 1024 tile."""
        lr = self.optimizer.param_groups[0]["lr"]
        if reached_1024:
            lr = max(lr * 0.98, 1e-6)
            for g in self.optimizer.param_groups:
                g["lr"] = lr
        return lr

    def save_agent_state(self, path: str):
        torch.save({"model": self.model.state_dict(), "target": self.target_model.state_dict(),
                    "opt": self.optimizer.state_dict(), "step_counter": self.step_counter, "epsilon": self.epsilon,
                    "memory": (self.mem_state, self.mem_next, self.mem_action, self.mem_reward, self.mem_done,
                               self.nb_entries, self._head), "loss_history": self.loss_history}, path)

    def load_agent_state(self, path: str):
        blob = torch.load(path, map_location=self.device)
        self.model.load_state_dict(blob["model"])
        self.target_model.load_state_dict(blob["target"])
        self.optimizer.load_state_dict(blob["opt"])
        self.step_counter, self.epsilon, self.loss_history = blob["step_counter"], blob["epsilon"], blob["loss_history"]
        (self.mem_state, self.mem_next, self.mem_action, self.mem_reward, self.mem_done, self.nb_entries,
         self._head) = blob["memory"]


def terminal_bonus(boards: torch.Tensor, done: torch.Tensor) -> torch.Tensor:
    """The driver's terminal bonus (mainDQL_CNN_step2.py:202-213): +100 if the final board holds a 2048 tile or
    more, +50 if it holds two tiles >= 1024, on packed boards."""
    lv = torch.stack([(boards >> (4 * j)) & 15 for j in range(16)], dim=1)
    top2 = lv.topk(2, dim=1).values
    bonus = torch.where(top2[:, 0] >= 11, 100.0, torch.where((top2[:, 0] >= 10) & (top2[:, 1] >= 10), 50.0, 0.0))
    return torch.where(done, bonus, torch.zeros_like(bonus))


def dqn_step(env: BatchedGame2048Env, agent: BatchedDQNAgent, train: bool = True, reference_driver: bool = False):
    """One step of the driver loop mainDQL_CNN_step2.py:163-237 for all envs: legal-move mask, action, env.step
    (nopenalty flavour; the commit of :237 is folded into the batched env), terminal bonus, remember, and a reset of
    the finished games.  Unfused form (one library call per piece); `FusedDQNFeed` does the env side in one launch.

    Default: every action is restricted to the legal moves (act_ripetitive) and every transition is stored.
    reference_driver=True follows the reference literally: act() -- unrestricted, invalid moves are played, cost -10
    and are stored -- unless the env's previous transition was not stored (:183-185), then act_ripetitive(); and
    remember() drops repeats of the env's previously stored transition (:283-297)."""
    state = env.boards.clone()
    legal = env.legal_mask()
    if reference_driver:
        actions = agent.act(state, env.env_id_base)                                    # :176
        saved = agent.memory_saved
        if saved is None or saved.numel() != env.n:
            saved = torch.zeros(env.n, dtype=torch.bool, device=env.device)            # memory_saved starts False
        if not bool(saved.all()):
            actions = torch.where(saved, actions, agent.act_ripetitive(state, legal, env.env_id_base))   # :183-185
    else:
        actions = agent.act_ripetitive(state, legal, env.env_id_base)
    next_state, reward, done, _ = env.step(actions)
    reward = reward.to(torch.float32) + terminal_bonus(next_state, done)
    if train:
        kept = agent.remember(state, actions, reward, done, next_state.clone(), filter_duplicates=reference_driver)
        if reference_driver:
            agent.memory_saved = torch.where(done, torch.zeros_like(kept), kept)       # a new game starts unsaved (:153-158)
    if bool(done.any()):
        env.reset(mask=done)
    return reward, done


class FusedDQNFeed:
    """The env side of the DQN loop in ONE kernel launch per step (`g2048_dqn_env_step`): act_ripetitive on the
    network's outputs, env step, terminal bonus, reset of finished games, legal-move mask and one-hot encoding of the
    boards the envs continue from.  Per step the host does: q = model(onehot) -> step(q) -> remember."""

    def __init__(self, env: BatchedGame2048Env, agent: BatchedDQNAgent, terminal_bonus: bool = True):
        if env.flavour != 1:
            raise ValueError("the DQN driver uses the nopenalty env")
        self.env, self.agent = env, agent
        dev, n = env.device, env.n
        self.opts = (1 if terminal_bonus else 0) | 2
        self.code = {torch.float32: 0, torch.bfloat16: 1}[agent.dtype]
        self.onehot = agent.encode_state(env.boards)
        self.legal = env.legal_mask()
        self.actions = torch.empty(n, dtype=torch.uint8, device=dev)
        self.state = torch.empty(n, dtype=torch.int64, device=dev)
        self.next_state = torch.empty(n, dtype=torch.int64, device=dev)
        self.reward = torch.empty(n, dtype=torch.float32, device=dev)
        self.done = torch.empty(n, dtype=torch.uint8, device=dev)

    def step(self, qvalues: torch.Tensor | None = None, train: bool = True):
        env, agent = self.env, self.agent
        if qvalues is None:
            with torch.no_grad():
                agent.model.eval()
                qvalues = agent.model(self.onehot).float().contiguous()
        agent.update_epsilon()
        agent.step_counter += 1
        with torch.cuda.device(env.device):
            check(agent.lib.g2048_dqn_env_step(_ptr(env.boards), _ptr(env.score), _ptr(qvalues), _ptr(self.legal),
                                               _ptr(self.actions), _ptr(self.state), _ptr(self.next_state),
                                               _ptr(self.reward), _ptr(self.done), _ptr(self.legal), _ptr(self.onehot),
                                               self.code, env.n, float(agent.epsilon), self.opts, env.seed, env.step_idx,
                                               env.episode_idx, env.env_id_base, _stream()), "g2048_dqn_env_step")
        env.step_idx += 1
        env.episode_idx += 1
        done = self.done != 0
        if train:
            agent.remember(self.state, self.actions, self.reward, done, self.next_state)
        return self.reward, done
