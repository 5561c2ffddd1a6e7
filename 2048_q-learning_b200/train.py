"""Episode-level drivers (SURVEY.md section 8f "next" #2): the reference's tabular training loop, once verbatim
on the N = 1 adapters (or the reference's own classes) and once in batched form on the GPU classes.

Reference: QLearningBase/Agent/main.py:59-115 (loop, CSV debug log `Episode,Action,Q-Values,Reward,Total-Reward,
Max Value`, epsilon decay per episode).  Driver quirks that are not reproduced: `num_episodes = 3` as committed
(:67) and the "extend until 1024" hack (:88-89), which cannot extend an already evaluated `range`.
"""
from __future__ import annotations

import csv

import numpy as np

CSV_HEADER = ["Episode", "Action", "Q-Values", "Reward", "Total-Reward", "Max Value"]


def log_debug_info(file_path, episode, action, q_values, reward, total_reward, max_value):
    """main.py:59-62"""
    with open(file_path, mode="a", newline="") as f:
        csv.writer(f).writerow([episode, action, q_values, reward, total_reward, max_value])


def train_tabular(env, agent, num_episodes: int, log_file: str | None = None, on_episode=None):
    """The loop of main.py:80-109 on any (env, agent) pair with the reference's duck-typed API -- the N = 1 adapters
    `Game2048_env` / `QLearningAgent` of this package, or the reference's own objects.  Returns per-episode
    (total_reward, max_tile, steps)."""
    if log_file:
        with open(log_file, mode="w", newline="") as f:
            csv.writer(f).writerow(CSV_HEADER)
    history = []
    for episode in range(num_episodes):
        state = tuple(map(tuple, env.reset()))
        done, total_reward, steps = False, 0, 0
        while not done:
            action = agent.choose_action(state)
            next_state, reward, done, info = env.step(action)
            next_state = tuple(map(tuple, next_state))
            q_values = agent.q_table[state]
            max_value = np.max(next_state)
            agent.update_q_value(state, action, reward, next_state, done)
            state = next_state
            total_reward += reward
            steps += 1
            if done and log_file:
                log_debug_info(log_file, episode, action, q_values, reward, total_reward, max_value)
        agent.decay_exploration(episode)
        history.append((total_reward, int(max_value), steps))
        if on_episode:
            on_episode(episode, history[-1])
    return history


BATCHED_CSV_HEADER = ["Epoch", "Epsilon", "Env-Steps", "Episodes", "Valid-Fraction", "Mean-Episode-Score", "Max Value",
                      "States", "Lost-Updates"]


def train_tabular_batched(env, agent, total_epochs: int, steps_per_epoch: int = 64, log_file: str | None = None,
                          on_epoch=None):
    """Batched form: an "epoch" is `steps_per_epoch` fused env steps of every env (agent.rollout) followed by one
    step of the reference's epsilon schedule (decay_exploration, main.py:45-57).  One CSV row per epoch."""
    if log_file:
        with open(log_file, mode="w", newline="") as f:
            csv.writer(f).writerow(BATCHED_CSV_HEADER)
    env.reset()
    history = []
    for epoch in range(total_epochs):
        c = env.counters_dict(agent.rollout(env, steps_per_epoch))
        row = [epoch, agent.epsilon, c["steps"], c["episodes"], c["valid"] / max(c["steps"], 1),
               c["score"] / max(c["episodes"], 1), 1 << c["maxlvl"], len(agent), c["lost"]]
        history.append(row)
        if log_file:
            with open(log_file, mode="a", newline="") as f:
                csv.writer(f).writerow(row)
        if on_epoch:
            on_epoch(epoch, row)
        agent.decay_exploration(epoch)
    return history


DQN_CSV_HEADER = ["Step", "Episodes", "Epsilon", "Mean Reward", "Done", "Replays", "Loss", "LR", "Best Tile", "Buffer"]


def train_dqn(env, agent, total_steps: int, *, replays_per_episode: int = 100, max_replays_per_step: int = 100,
              target_sync_episodes: int = 20, save_every_episodes: int = 100, save_dir: str | None = None,
              log_file: str | None = None, on_step=None):
    """The DQN driver loop (mainDQL_CNN_step2.py:151-333) for N envs at once on the batched nopenalty env.

    Per step (one launch of the fused env side, `FusedDQNFeed`): legal-move mask -> act_ripetitive (:169-185) ->
    env.step (:200) -> terminal bonus (:202-213) -> remember (:220) -> board commit (:237) -> reset of finished games.
    Episode-level cadence as in the reference, counted in finished episodes of ANY env: `replays_per_episode` replay()
    calls per finished game (:223-226, capped per step), learning-rate cut when a game ended on a board holding a
    1024 tile (Dqn8TestNOPERCNN.py:283-284, :299-310), target sync every 20 episodes (:275-277), full agent save every
    100 episodes (:322-328).  Not reproduced: the plotting (:271), the commented-out rollback block, and
    clean_low_score_episodes (:317-319; keras-rl episode bookkeeping that a flat transition ring does not keep).
    Every action here is restricted to the legal moves and every transition is stored; the reference's own rule
    (act() unrestricted with invalid moves played and stored, act_ripetitive() only after a dropped transition,
    duplicate filter in remember()) is `dqn.dqn_step(env, agent, reference_driver=True)`, the unfused per-step form.
    Returns {"episodes", "steps", "max_tile_list", "score_list", "loss_history", "best_tile"} like the arrays the
    reference keeps (:100-103)."""
    import os

    import torch

    from .dqn import FusedDQNFeed
    if log_file:
        with open(log_file, mode="w", newline="") as f:
            csv.writer(f).writerow(DQN_CSV_HEADER)
    if env.step_idx == 0 and int(env.boards.ne(0).sum()) == 0:
        env.reset()
    feed = FusedDQNFeed(env, agent)
    episodes, best_tile = 0, 0
    max_tile_list, score_list, loss_history = [], [], []
    next_sync, next_save = target_sync_episodes, save_every_episodes
    for step in range(total_steps):
        prev_score = env.score.clone()                       # env.score of a finished game is reset inside the launch
        reward, done = feed.step()
        n_done = int(done.sum())
        replays, loss = 0, None
        if n_done:
            final = feed.next_state[done]
            lv = torch.stack([(final >> (4 * j)) & 15 for j in range(16)], dim=1).max(dim=1).values
            tiles = (1 << lv.to(torch.int64)).tolist()
            max_tile_list += tiles
            score_list += prev_score[done].tolist()
            best_tile = max(best_tile, max(tiles))
            start = feed.state[done]
            lv0 = torch.stack([(start >> (4 * j)) & 15 for j in range(16)], dim=1).max(dim=1).values
            if bool((lv0 >= 10).any()):
                agent.change_lr_function(True)
            episodes += n_done
            replays = min(replays_per_episode * n_done, max_replays_per_step)
            for _ in range(replays):
                out = agent.replay(episodes)
                loss = out if out is not None else loss
            if loss is not None:
                loss_history.append(loss)
            if target_sync_episodes and episodes >= next_sync:
                agent.update_target_model()
                next_sync = (episodes // target_sync_episodes + 1) * target_sync_episodes
            if save_dir and save_every_episodes and episodes >= next_save:
                os.makedirs(save_dir, exist_ok=True)
                agent.save_agent_state(os.path.join(save_dir, f"agent_episode_{episodes}.pt"))
                next_save = (episodes // save_every_episodes + 1) * save_every_episodes
        row = [step, episodes, agent.epsilon, float(reward.mean()), n_done, replays, loss,
               agent.optimizer.param_groups[0]["lr"], best_tile, agent.nb_entries]
        if log_file:
            with open(log_file, mode="a", newline="") as f:
                csv.writer(f).writerow(row)
        if on_step:
            on_step(step, row)
    return {"episodes": episodes, "steps": total_steps, "max_tile_list": max_tile_list, "score_list": score_list,
            "loss_history": loss_history, "best_tile": best_tile}


def evaluate_random(env, episodes: int = 10, rng=None):
    """The demo's "auto play" mode (GameDemo.py:272-286), headless: uniformly random actions on the N = 1 adapters or
    the reference's env objects until the game ends.  Returns per-episode (game score, max tile, steps)."""
    import numpy as np
    rng = rng or np.random
    out = []
    for _ in range(episodes):
        env.reset()
        done, steps, max_tile = False, 0, 0
        while not done:
            step = env.step(int(rng.randint(0, 4)))
            board, done, max_tile = step[0], step[2], step[3]
            if hasattr(env.game, "moved_board"):          # nopenalty flavour: the caller commits the board (:279)
                env.game.board = board
            steps += 1
        out.append((int(env.score), int(max_tile), steps))
    return out


def evaluate_dqn(env, agent, episodes: int = 1, max_steps: int = 100000):
    """The demo's "model play" mode (GameDemo.py:288-316) for ALL envs of a batched nopenalty env at once: greedy on
    the network's Q values restricted to the legal moves (epsilon = 0), no learning, until `episodes` games per env
    have ended.  Returns {"scores": [...], "max_tiles": [...], "steps": n}."""
    import torch

    from .dqn import FusedDQNFeed
    saved = (agent.epsilon_start, agent.epsilon_min, agent.step_counter)
    agent.epsilon_start = agent.epsilon_min = 0.0
    scores, tiles, steps = [], [], 0
    try:
        env.reset()
        feed = FusedDQNFeed(env, agent)
        left = torch.full((env.n,), episodes, dtype=torch.int64, device=env.device)
        while steps < max_steps and bool((left > 0).any()):
            prev_score = env.score.clone()
            _, done = feed.step(train=False)
            steps += 1
            count = done & (left > 0)
            if bool(count.any()):
                final = feed.next_state[count]
                lv = torch.stack([(final >> (4 * j)) & 15 for j in range(16)], dim=1).max(dim=1).values
                tiles += (1 << lv.to(torch.int64)).tolist()
                scores += prev_score[count].tolist()
                left -= count.to(torch.int64)
    finally:
        agent.epsilon_start, agent.epsilon_min, agent.step_counter = saved
    return {"scores": scores, "max_tiles": tiles, "steps": steps}


def evaluate_tabular(env, agent, episodes: int = 10):
    """Greedy play (epsilon = 0) of the N = 1 adapters or the reference's objects -- the headless form of the
    demo's "model play" mode (GameDemo.py:258-316).  Returns per-episode (game score, max tile, steps)."""
    saved, agent.epsilon = agent.epsilon, 0.0
    out = []
    try:
        for _ in range(episodes):
            state = tuple(map(tuple, env.reset()))
            done, steps, max_tile = False, 0, 0
            while not done:
                action = agent.choose_action(state)
                board, reward, done, max_tile = env.step(action)
                if hasattr(env.game, "moved_board"):      # nopenalty flavour: the caller commits the board
                    env.game.board = board
                state = tuple(map(tuple, board))
                steps += 1
            out.append((int(env.score), int(max_tile), steps))
    finally:
        agent.epsilon = saved
    return out
