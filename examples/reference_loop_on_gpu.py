"""The reference's tabular training script (QLearningBase/Agent/main.py:65-115) with two imports changed.

    python examples/reference_loop_on_gpu.py [episodes]

Every env.step / choose_action / update_q_value below runs in libg2048.so on the GPU; under the same
np.random / random seeds the trajectory is the reference's own (tests/test_gpu_agent.py)."""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from g2048 import Game2048_env, QLearningAgent, train_tabular  # noqa: E402  (instead of the reference's modules)

if __name__ == "__main__":
    episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    np.random.seed(0)
    random.seed(0)
    env = Game2048_env()
    agent = QLearningAgent(episodes, action_space=env.action_space.n, learning_rate=0.1, discount_factor=0.99,
                           exploration_rate=0.95)
    import time
    t0 = time.perf_counter()
    history = train_tabular(env, agent, episodes, log_file="debug_log.csv",
                            on_episode=lambda ep, h: print(f"episode {ep}: total reward {h[0]:.3f}, max tile {h[1]}, {h[2]} steps"))
    dt = time.perf_counter() - t0
    print(f"{len(agent.q_table)} states in the Q-table, epsilon {agent.epsilon:.4f}; "
          f"{sum(h[2] for h in history) / dt:.0f} env steps/s through the N = 1 adapters (a compatibility route: one env "
          f"cannot fill a GPU -- the batched classes are the throughput path)")
