"""Where the time of one exact synchronous step goes: G virtual ranks of 2^20 envs on one GPU (emit per rank, then one
apply of all G record lists, as every replica does).  Run under `ncu --metrics gpu__time_duration.sum` for the split."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g2048  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, steps = 1 << 20, int(sys.argv[2]) if len(sys.argv) > 2 else 6
envs = [g2048.BatchedGame2048Env(n, "penalty", seed=7, env_id_base=r * n) for r in range(G)]
agent = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.1, capacity=1 << 28, seed=7)
recs = [torch.zeros((n, 2), dtype=torch.int64, device="cuda") for _ in range(G)]
for e in envs:
    e.reset()
for t in range(steps + 2):
    if t == 2:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    for r in range(G):
        agent.emit_records(envs[r], recs[r])
    agent.apply_records(recs, [n] * G)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print(f"G={G}: {dt * 1e3:.3f} ms per step of {G * n} records ({G * n / dt / 1e9:.2f} G records/s)")
