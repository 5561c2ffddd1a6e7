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
