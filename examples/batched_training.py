"""Tabular Q-learning at GPU scale: 2^20 environments, fused epsilon-greedy rollouts, the reference's epsilon schedule.

    python examples/batched_training.py [epochs]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from g2048 import BatchedGame2048Env, BatchedQLearningAgent, train_tabular_batched  # noqa: E402

if __name__ == "__main__":
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    env = BatchedGame2048Env(1 << 20, flavour="penalty", seed=0x2048)
    agent = BatchedQLearningAgent(epochs, 4, learning_rate=0.1, discount_factor=0.99, exploration_rate=0.95,
                                  capacity=1 << 30)
    t0 = time.perf_counter()
    hist = train_tabular_batched(env, agent, epochs, steps_per_epoch=32, log_file="batched_log.csv",
                                 on_epoch=lambda ep, row: print(dict(zip(("epoch", "eps", "steps", "episodes", "valid",
                                                                          "score/episode", "max tile", "states", "lost"), row))))
    dt = time.perf_counter() - t0
    print(f"{sum(r[2] for r in hist) / dt / 1e9:.2f} G env-steps/s including logging; table: {len(agent)} states")
    agent.save("q_table.pt")
