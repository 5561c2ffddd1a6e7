"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

TEST INFRASTRUCTURE.  The reference has no tests or fixtures of its own
(SURVEY.md section 4), so parity is pinned against the reference itself: this
script imports it (oracle/ref_shim.py), records every RNG draw it makes by
wrapping `np.random.randint` / `np.random.random` (the envs look them up through
the `np.random` namespace at call time, Game2048_env.py:19-20) and
`random.random` / `random.randint` (tabular agent, main.py:35-36), and stores
inputs + outputs as small .npz files.  The committed files are what
tests/test_oracle_golden.py (CPU) and tests/test_gpu_parity.py (GPU) replay.

Files
  env_penalty.npz    QLearningBase/environment/Game2048_env.py        (penalty flavour)
  env_nopenalty.npz  Deep_QLearning/environment/Game2048_nopenalty_env.py under the
                     caller protocol of mainDQL_CNN_step2.py:163-237 (caller commits the board)
  qlearn_ref.npz     QLearningBase/Agent/main.py QLearningAgent driven by the loop main.py:80-109
  compat_seeded.npz  both envs and the full tabular loop under fixed np.random / random seeds
"""
from __future__ import annotations

import os
import random as pyrandom
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
N_ENVS = 96
N_STEPS = 384
PEN_SAT = 25
SEED_OFFSET = 0      # the committed goldens use 0; tests/test_oracle_live_reference.py records fresh trajectories with others


def pack_board(tiles) -> int:
    b = 0
    for j, v in enumerate(np.asarray(tiles).reshape(16)):
        v = int(v)
        lvl = v.bit_length() - 1 if v else 0
        assert v == 0 or (1 << lvl) == v and 1 <= lvl <= 15, v
        b |= lvl << (4 * j)
    return b


class DrawRecorder:
    """Wraps np.random.randint / np.random.random; every call is logged."""

    def __init__(self):
        self.log = []
        self._ri, self._rr = np.random.randint, np.random.random

    def __enter__(self):
        def randint(low, high=None, *a, **k):
            v = self._ri(low, high, *a, **k)
            self.log.append(("i", int(v), int(high)))
            return v

        def rnd(*a, **k):
            v = self._rr(*a, **k)
            self.log.append(("f", float(v)))
            return v

        np.random.randint, np.random.random = randint, rnd
        return self

    def __exit__(self, *exc):
        np.random.randint, np.random.random = self._ri, self._rr

    def take(self):
        out, self.log = self.log, []
        return out


def spawn_pairs(log):
    """[(k, is4), ...] from a draw log that must be (randint, random) pairs."""
    assert len(log) % 2 == 0, log
    pairs = []
    for i in range(0, len(log), 2):
        assert log[i][0] == "i" and log[i + 1][0] == "f", log
        pairs.append((log[i][1], 0 if log[i + 1][1] < 0.9 else 1))
    return pairs


def teleport_board(rs: np.random.RandomState, lmax: int) -> np.ndarray:
    """A random mid/late-game board (raw tile values) with max level <= lmax."""
    p_zero = rs.choice([0.0, 0.0, 0.1, 0.3])
    lv = rs.randint(1, min(lmax, 14) + 1, size=16)  # at most one level-15 tile: 65536 is unrepresentable
    lv[rs.randint(0, 16)] = lmax
    lv = np.where(rs.random_sample(16) < p_zero, 0, lv)
    if rs.random_sample() < 0.35:  # checkerboard-ish dead board
        cap = min(lmax, 14)
        a, b = rs.randint(1, cap + 1), rs.randint(1, cap + 1)
        if a == b:
            b = a % cap + 1 if cap > 1 else a
        if a != b:
            lv = np.array([[a, b, a, b], [b, a, b, a], [a, b, a, b], [b, a, b, a]]).reshape(16)
            for _ in range(rs.randint(0, 4)):
                lv[rs.randint(0, 16)] = rs.randint(1, cap + 1)
    board = np.where(lv > 0, 2 ** lv.astype(np.int64), 0).reshape(4, 4)
    if not board.any():
        board[0, 0] = 2
    return board.astype(np.int64)


def policy_action(kind: str, rs: np.random.RandomState, last: int | None) -> int:
    if kind == "uniform" or last is None:
        return int(rs.randint(0, 4))
    if kind == "sticky":  # long same-action runs: stall penalty and >100 termination
        return last if rs.random_sample() < 0.985 else int(rs.randint(0, 4))
    if kind == "corner":  # mostly left/up, reaches higher tiles
        return int(rs.choice([0, 1, 0, 1, 0, 1, 2, 3]))
    raise ValueError(kind)


def env_kind(i: int) -> tuple[str, bool]:
    """(policy, teleport) of golden env i."""
    if i < 40:
        return "uniform", False
    if i < 56:
        return "sticky", False
    if i < 64:
        return "corner", False
    if i < 88:
        return "uniform", True
    return "sticky", True


def record_env(flavour: str):
    mod = ref_shim.load_penalty_env() if flavour == "penalty" else ref_shim.load_nopenalty_env()
    E, T = N_ENVS, N_STEPS
    g = {
        "board_in": np.zeros((E, T), np.uint64), "reload": np.zeros((E, T), np.uint8),
        "action": np.zeros((E, T), np.uint8), "draws": np.full((E, T, 4), 255, np.uint8),
        "board_out": np.zeros((E, T), np.uint64), "reward": np.zeros((E, T), np.float64),
        "flags": np.zeros((E, T), np.uint8), "maxlvl": np.zeros((E, T), np.uint8),
        "move_score": np.zeros((E, T), np.int32), "env_score": np.zeros((E, T), np.int32),
        "aux_out": np.zeros((E, T), np.uint64), "last_pen": np.zeros((E, T), np.float64),
    }
    reset_draws, reset_boards = [], []
    pen_table = [-1.0]
    for _ in range(64):
        pen_table.append(max(pen_table[-1] * 1.1, -10))

    # capture (moved, score) of every Game2048.move and the result of is_game_over
    calls = {}
    orig_move, orig_over = mod.Game2048.move, mod.Game2048.is_game_over

    def move_wrap(self, *a, **k):
        res = orig_move(self, *a, **k)
        calls.setdefault("move", []).append((a, k, (bool(res[0]), int(res[1]))))
        return res

    def over_wrap(self):
        calls["in_over"] = True
        n0 = len(calls.get("move", []))
        res = orig_over(self)
        calls["over"] = bool(res)
        calls["over_moves"] = calls.get("move", [])[n0:]
        del calls["move"][n0:]
        calls["in_over"] = False
        return res

    mod.Game2048.move, mod.Game2048.is_game_over = move_wrap, over_wrap
    try:
        with DrawRecorder() as rec:
            for i in range(E):
                kind, tele = env_kind(i)
                np.random.seed(1000 + SEED_OFFSET + i)
                rs = np.random.RandomState(2000 + SEED_OFFSET + i)
                env = mod.Game2048_env()
                rec.take()
                episode = 0
                fresh = True
                last = None
                for t in range(T):
                    if fresh and tele:
                        lmax = int(min(15, 2 + episode // 2 + rs.randint(0, 3)))
                        env.game.board = teleport_board(rs, lmax)
                    b_in = pack_board(env.game.board)
                    g["board_in"][i, t] = b_in
                    g["reload"][i, t] = 1 if fresh else 0
                    fresh = False
                    full_before = not (np.asarray(env.game.board) == 0).any()
                    a = policy_action(kind, rs, last)
                    last = a
                    calls.clear()
                    board, reward, done, max_number = env.step(a)
                    log = rec.take()
                    (margs, mkw, (valid, mscore)) = calls["move"][0]
                    assert margs[0] == a
                    game_over = calls["over"]
                    pairs = spawn_pairs(log)
                    d = [255, 255, 255, 255]
                    if flavour == "penalty":
                        # [spawn of the move if valid] + [phantom spawn if the new board is full and alive]
                        if valid:
                            d[0], d[1] = pairs[0]
                        assert len(pairs) == int(valid) + int(len(calls["over_moves"]) > 0 and not game_over)
                    else:
                        # [spawn of the move if valid] + [quirk spawn if S was full and alive]
                        if valid:
                            d[0], d[1] = pairs[0]
                        if full_before and not game_over:
                            d[2], d[3] = pairs[-1]
                        assert len(pairs) == int(valid) + int(full_before and not game_over)
                        env.game.board = board  # caller commit, mainDQL_CNN_step2.py:237
                    g["action"][i, t] = a
                    g["draws"][i, t] = d
                    g["board_out"][i, t] = pack_board(board)
                    g["reward"][i, t] = float(reward)
                    g["flags"][i, t] = int(valid) | (int(game_over) << 1) | (int(bool(done)) << 2)
                    g["maxlvl"][i, t] = int(max_number).bit_length() - 1
                    g["move_score"][i, t] = mscore
                    g["env_score"][i, t] = int(env.score)
                    if flavour == "penalty":
                        ca = 255 if env.consecutive_action is None else int(env.consecutive_action)
                        lp = float(env.last_consecutive_penalty)
                        pen_idx = min(pen_table.index(lp), PEN_SAT)
                        prev_level = int(env.previous_max).bit_length() - 1
                        g["aux_out"][i, t] = prev_level | (ca << 8) | (pen_idx << 16) | (int(env.consecutive_count) << 32)
                        g["last_pen"][i, t] = lp
                    if done:
                        env.reset()
                        rp = spawn_pairs(rec.take())
                        assert len(rp) == 2
                        reset_draws.append([rp[0][0], rp[0][1], rp[1][0], rp[1][1]])
                        reset_boards.append(pack_board(env.game.board))
                        episode += 1
                        fresh = True
    finally:
        mod.Game2048.move, mod.Game2048.is_game_over = orig_move, orig_over
    g["reset_draws"] = np.array(reset_draws, np.uint8).reshape(-1, 4)
    g["reset_board"] = np.array(reset_boards, np.uint64)
    return g


def record_qlearn(episodes: int = 150):
    """The loop of main.py:80-109 on the reference agent + penalty env."""
    penv = ref_shim.load_penalty_env()
    agent_mod = ref_shim.load_tabular_agent()
    np.random.seed(SEED_OFFSET)
    pyrandom.seed(SEED_OFFSET)
    env = penv.Game2048_env()
    agent = agent_mod.QLearningAgent(episodes, action_space=env.action_space.n, learning_rate=0.1,
                                     discount_factor=0.99, exploration_rate=0.95)
    agent_draws = []
    orig_r, orig_i = pyrandom.random, pyrandom.randint

    def r_wrap():
        v = orig_r()
        agent_draws.append(("f", v))
        return v

    def i_wrap(a, b):
        v = orig_i(a, b)
        agent_draws.append(("i", v))
        return v

    pyrandom.random, pyrandom.randint = r_wrap, i_wrap
    S, A, R, S2, D, EXP, RA, EP = [], [], [], [], [], [], [], []
    eps_hist = []
    try:
        for episode in range(episodes):
            eps_hist.append(agent.epsilon)
            state = tuple(map(tuple, env.reset()))
            done = False
            while not done:
                del agent_draws[:]
                action = agent.choose_action(state)
                explore = agent_draws[0][1] < agent.epsilon
                ra = agent_draws[1][1] if explore else 255
                next_state, reward, done, info = env.step(action)
                next_state = tuple(map(tuple, next_state))
                _ = agent.q_table[state]  # main.py:96 (insert side effect)
                agent.update_q_value(state, action, reward, next_state, done)
                S.append(pack_board(state)); A.append(int(action)); R.append(float(reward))
                S2.append(pack_board(next_state)); D.append(int(bool(done)))
                EXP.append(int(explore)); RA.append(int(ra)); EP.append(episode)
                state = next_state
            agent.decay_exploration(episode)
    finally:
        pyrandom.random, pyrandom.randint = orig_r, orig_i
    eps_hist.append(agent.epsilon)
    keys = np.array([pack_board(k) for k in agent.q_table.keys()], np.uint64)
    rows = np.array([np.asarray(v, np.float64) for v in agent.q_table.values()], np.float64).reshape(-1, 4)
    order = np.argsort(keys)
    return {
        "s": np.array(S, np.uint64), "a": np.array(A, np.uint8), "r": np.array(R, np.float64),
        "s2": np.array(S2, np.uint64), "done": np.array(D, np.uint8), "explore": np.array(EXP, np.uint8),
        "rand_action": np.array(RA, np.uint8), "episode": np.array(EP, np.int32),
        "eps": np.array(eps_hist, np.float64), "q_keys": keys[order], "q_rows": rows[order],
        "params": np.array([episodes, 0.1, 0.99, 0.95, 0.01], np.float64),
    }


def record_seeded(steps: int = 400, episodes: int = 20):
    """What a user of the reference sees under fixed seeds (for the drop-in adapters, rng="numpy"/"python"):
    (a) both envs driven by a fixed action sequence under np.random.seed(4242);
    (b) the full tabular loop main.py:80-109 under np.random.seed(1) / random.seed(1)."""
    out = {}
    for flavour in ("penalty", "nopenalty"):
        mod = ref_shim.load_penalty_env() if flavour == "penalty" else ref_shim.load_nopenalty_env()
        np.random.seed(4242)
        actions = np.random.RandomState(7).randint(0, 4, size=steps)
        env = mod.Game2048_env()
        boards, rewards, dones, maxes, scores, first = [], [], [], [], [], pack_board(env.game.board)
        for t in range(steps):
            board, reward, done, max_number = env.step(int(actions[t]))
            if flavour == "nopenalty":
                env.game.board = board
            boards.append(pack_board(board)); rewards.append(float(reward)); dones.append(int(bool(done)))
            maxes.append(int(max_number)); scores.append(int(env.score))
            if done:
                env.reset()
                boards[-1] = pack_board(env.game.board)  # what the next step starts from
        out.update({f"{flavour}_actions": actions.astype(np.uint8), f"{flavour}_first": np.uint64(first),
                    f"{flavour}_boards": np.array(boards, np.uint64), f"{flavour}_rewards": np.array(rewards),
                    f"{flavour}_dones": np.array(dones, np.uint8), f"{flavour}_max": np.array(maxes, np.int64),
                    f"{flavour}_scores": np.array(scores, np.int64)})
    penv = ref_shim.load_penalty_env()
    agent_mod = ref_shim.load_tabular_agent()
    np.random.seed(1)
    pyrandom.seed(1)
    env = penv.Game2048_env()
    agent = agent_mod.QLearningAgent(episodes, action_space=env.action_space.n, learning_rate=0.1,
                                     discount_factor=0.99, exploration_rate=0.95)
    A, R, B, totals = [], [], [], []
    for episode in range(episodes):
        state = tuple(map(tuple, env.reset()))
        done, total = False, 0
        while not done:
            action = agent.choose_action(state)
            next_state, reward, done, info = env.step(action)
            next_state = tuple(map(tuple, next_state))
            _ = agent.q_table[state]
            agent.update_q_value(state, action, reward, next_state, done)
            state = next_state
            total += reward
            A.append(int(action)); R.append(float(reward)); B.append(pack_board(next_state))
        totals.append(total)
        agent.decay_exploration(episode)
    keys = np.array([pack_board(k) for k in agent.q_table.keys()], np.uint64)
    rows = np.array([np.asarray(v, np.float64) for v in agent.q_table.values()]).reshape(-1, 4)
    order = np.argsort(keys)
    out.update({"loop_actions": np.array(A, np.uint8), "loop_rewards": np.array(R), "loop_boards": np.array(B, np.uint64),
                "loop_totals": np.array(totals), "loop_q_keys": keys[order], "loop_q_rows": rows[order],
                "loop_params": np.array([episodes, 0.1, 0.99, 0.95], np.float64), "loop_eps_final": np.float64(agent.epsilon)})
    return out


def record_dqn(n_boards: int = 512, n_cases: int = 3000):
    """The action-selection path of the reference's DQN agent, executed from its own source (ref_shim.load_dqn_agent):
    encode_state (Dqn8TestNOPERCNN.py:271-277) on random boards, and act / act_ripetitive (:312-336) + update_epsilon
    (:341-343) on a stand-in object whose `model.predict` returns prescribed Q-values.  epsilon_start = 0 puts every
    call on the greedy branch, which is the deterministic part (argmax, ties, restriction to the legal moves, the
    fallback without legal moves); the exploring branch draws from MT19937 and is checked statistically elsewhere."""
    mod = ref_shim.load_dqn_agent()
    A = mod.DQNAgent
    rs = np.random.RandomState(77)
    lv = rs.randint(0, 16, size=(n_boards, 16))
    lv = np.where(rs.random_sample((n_boards, 16)) < rs.random_sample((n_boards, 1)), 0, lv)
    tiles = np.where(lv > 0, 1 << lv, 0).astype(np.int64).reshape(n_boards, 4, 4)
    enc = np.concatenate([np.asarray(A.encode_state(None, b)) for b in tiles]).astype(np.float32)

    class Model:
        q = None

        def predict(self, x, verbose=0):
            return Model.q[None, :]

    class Stand:
        pass
    me = Stand()
    me.epsilon_start, me.epsilon_min, me.epsilon_decay, me.epsilon, me.step_counter, me.action_space = 0.0, 0.0, 0.9999, 0.0, 0, 4
    me.model = Model()
    me.encode_state = lambda b: A.encode_state(me, b)
    me.update_epsilon = lambda: A.update_epsilon(me)
    q = rs.standard_normal((n_cases, 4)).astype(np.float32)
    q[::3, 1] = q[::3, 3]          # ties
    q[::5, 0] = q[::5, 2]
    q[::11] = q[::11, :1]          # all four equal
    legal = rs.randint(0, 16, n_cases).astype(np.uint8)
    legal[::13] = 0                # no legal move at all
    act, act_rip = np.zeros(n_cases, np.uint8), np.zeros(n_cases, np.uint8)
    np.random.seed(5)              # the draws only decide "explore?": rand() <= 0 never holds
    for i in range(n_cases):
        Model.q = q[i]
        board = tiles[i % n_boards]
        act[i] = int(A.act(me, board))
        moves = [a for a in range(4) if (int(legal[i]) >> a) & 1]
        act_rip[i] = int(A.act_ripetitive(me, board, moves))
    assert me.step_counter == n_cases                         # act() counts steps (:316), act_ripetitive does not
    # update_epsilon with the reference's default constants (:249): epsilon(step) for a few step counts
    me.epsilon_start, me.epsilon_min = 0.9, 0.001
    steps = np.array([0, 1, 10, 1000, 10_000, 50_000, 67_000, 68_100, 100_000], np.int64)
    eps = []
    for st in steps:
        me.step_counter = int(st)
        A.update_epsilon(me)
        eps.append(me.epsilon)
    return {"tiles": tiles, "onehot": enc, "q": q, "legal": legal, "act": act, "act_ripetitive": act_rip,
            "eps_steps": steps, "eps": np.array(eps, np.float64)}


# --------------------------------------------------------------------------- BASELINE config 2 at its stated size
C2_ENVS, C2_STEPS = 4096, 512
FNV = np.uint64(0x100000001B3)


def c2_fold(h, *words):
    """Per-env running digest of everything a step returns (uint64 arithmetic wraps): the replay recomputes it."""
    for w in words:
        h = h * FNV + np.asarray(w).astype(np.uint64)
    return h


def record_config2(flavour: str, n_envs: int = C2_ENVS, n_steps: int = C2_STEPS, first_env: int = 0):
    """SURVEY.md 8d "C2": env i is seeded np.random.seed(1000 + i), plays actions
    np.random.RandomState(2000 + i).randint(0, 4) for 512 steps with reset-on-done (nopenalty flavour: caller commits
    the board).  Stored: the INPUTS of a bit-exact replay, packed -- actions (2 bits), the spawn of the move
    (k | is4 << 4, 255 = none), the full-board quirk spawn, the reset spawns -- and per env a 64-bit digest over every
    step's outputs (board, reward bits, flags, max level, move score, env.score, aux) plus the final board/score/aux."""
    mod = ref_shim.load_penalty_env() if flavour == "penalty" else ref_shim.load_nopenalty_env()
    E, T = n_envs, n_steps
    action = np.zeros((E, T), np.uint8)
    spawn = np.full((E, T), 255, np.uint8)
    quirk = np.full((E, T), 255, np.uint8)
    digest = np.zeros(E, np.uint64)
    final_board, final_score, final_aux = np.zeros(E, np.uint64), np.zeros(E, np.int32), np.zeros(E, np.uint64)
    start_draws = np.zeros((E, 4), np.uint8)
    resets = []                                    # (env, step after which the reset happened, k_a, is4_a, k_b, is4_b)
    pen_table = [-1.0]
    for _ in range(64):
        pen_table.append(max(pen_table[-1] * 1.1, -10))
    calls = {}
    orig_move, orig_over = mod.Game2048.move, mod.Game2048.is_game_over

    def move_wrap(self, *a, **k):
        res = orig_move(self, *a, **k)
        if not calls.get("in_over"):
            calls.setdefault("move", []).append((bool(res[0]), int(res[1])))
        return res

    def over_wrap(self):
        calls["in_over"] = True
        res = orig_over(self)
        calls["in_over"] = False
        calls["over"] = bool(res)
        return res

    mod.Game2048.move, mod.Game2048.is_game_over = move_wrap, over_wrap
    try:
        with DrawRecorder() as rec:
            for i in range(E):
                np.random.seed(1000 + first_env + i)
                rs = np.random.RandomState(2000 + first_env + i)
                env = mod.Game2048_env()
                sp = spawn_pairs(rec.take())
                start_draws[i] = [sp[0][0], sp[0][1], sp[1][0], sp[1][1]]
                h = np.uint64(0)
                with np.errstate(over="ignore"):
                    for t in range(T):
                        full_before = not (np.asarray(env.game.board) == 0).any()
                        a = int(rs.randint(0, 4))
                        calls.clear()
                        board, reward, done, max_number = env.step(a)
                        valid, mscore = calls["move"][0]
                        game_over = calls["over"]
                        pairs = spawn_pairs(rec.take())
                        if valid:
                            spawn[i, t] = pairs[0][0] | (pairs[0][1] << 4)
                        if flavour != "penalty":
                            if full_before and not game_over:
                                quirk[i, t] = pairs[-1][0] | (pairs[-1][1] << 4)
                            env.game.board = board         # caller commit, mainDQL_CNN_step2.py:237
                        action[i, t] = a
                        aux = 0
                        if flavour == "penalty":
                            ca = 255 if env.consecutive_action is None else int(env.consecutive_action)
                            pen_idx = min(pen_table.index(float(env.last_consecutive_penalty)), PEN_SAT)
                            aux = (int(env.previous_max).bit_length() - 1) | (ca << 8) | (pen_idx << 16) | \
                                  (int(env.consecutive_count) << 32)
                        fl = int(valid) | (int(game_over) << 1) | (int(bool(done)) << 2)
                        words = fl | ((int(max_number).bit_length() - 1) << 8) | (mscore << 16)
                        h = c2_fold(h, np.uint64(pack_board(board)), np.float64(reward).view(np.uint64), np.uint64(words),
                                    np.uint64(int(env.score)), np.uint64(aux))
                        if done:
                            env.reset()
                            rp = spawn_pairs(rec.take())
                            resets.append((i, t, rp[0][0], rp[0][1], rp[1][0], rp[1][1]))
                digest[i] = h
                final_board[i] = pack_board(env.game.board)
                final_score[i] = int(env.score)
                if flavour == "penalty":
                    ca = 255 if env.consecutive_action is None else int(env.consecutive_action)
                    pen_idx = min(pen_table.index(float(env.last_consecutive_penalty)), PEN_SAT)
                    final_aux[i] = (int(env.previous_max).bit_length() - 1) | (ca << 8) | (pen_idx << 16) | \
                                   (int(env.consecutive_count) << 32)
    finally:
        mod.Game2048.move, mod.Game2048.is_game_over = orig_move, orig_over
    packed_actions = (action[:, 0::4] | (action[:, 1::4] << 2) | (action[:, 2::4] << 4) | (action[:, 3::4] << 6)).astype(np.uint8)
    return {"actions4": packed_actions, "spawn": spawn, "quirk": quirk, "start_draws": start_draws,
            "resets": np.array(resets, np.int32).reshape(-1, 6), "digest": digest, "final_board": final_board,
            "final_score": final_score, "final_aux": final_aux, "shape": np.array([E, T, first_env], np.int64)}


def main():
    if not ref_shim.available():
        raise SystemExit(f"reference not found under {ref_shim.REF_ROOT}")
    os.makedirs(OUT_DIR, exist_ok=True)
    for flavour in ("penalty", "nopenalty"):
        g = record_env(flavour)
        path = os.path.join(OUT_DIR, f"env_{flavour}.npz")
        np.savez_compressed(path, **g)
        fl = g["flags"]
        print(f"{path}: {fl.size} steps, valid {np.mean(fl & 1):.3f}, game_over {int(np.sum((fl >> 1) & 1))}, "
              f"done {int(np.sum((fl >> 2) & 1))}, resets {len(g['reset_board'])}, max level {int(g['maxlvl'].max())}, "
              f"{os.path.getsize(path) / 1e3:.0f} kB")
    sd = record_seeded()
    path = os.path.join(OUT_DIR, "compat_seeded.npz")
    np.savez_compressed(path, **sd)
    print(f"{path}: {len(sd['loop_actions'])} loop steps, {os.path.getsize(path) / 1e3:.0f} kB")
    dq = record_dqn()
    path = os.path.join(OUT_DIR, "dqn_agent.npz")
    np.savez_compressed(path, **dq)
    print(f"{path}: {len(dq['tiles'])} encoded boards, {len(dq['q'])} action selections, {os.path.getsize(path) / 1e3:.0f} kB")
    q = record_qlearn()
    path = os.path.join(OUT_DIR, "qlearn_ref.npz")
    np.savez_compressed(path, **q)
    print(f"{path}: {len(q['s'])} transitions, {len(q['q_keys'])} states, {os.path.getsize(path) / 1e3:.0f} kB")


def main_config2():
    """python oracle/make_golden.py config2 penalty|nopenalty  (about ten minutes each)"""
    flavour = sys.argv[2]
    g = record_config2(flavour)
    path = os.path.join(OUT_DIR, f"config2_{flavour}.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes;", len(g["resets"]), "resets")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "config2":
    main_config2()
    sys.exit(0)
if __name__ == "__main__":
    main()
