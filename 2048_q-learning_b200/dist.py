"""Env shards over the GPUs of one box + the Q-target exchange (one process per GPU, torch.distributed).

Envs are independent (the reference has exactly one, main.py:66), so rank r owns the contiguous global env
ids [lo, hi) and its Philox draws are keyed by the GLOBAL env id: results do not depend on the sharding.
The Q-table is replicated; the only exchange step is the list of (state key, action, target) records of one
synchronous step (16 B per transition, fixed size per rank): `all_gather` in rank order = ascending global
env id, then every replica applies the whole list with the same deterministic kernel (sort by
(state, action) + segmented sum), so all replicas stay identical and equal to the 1-GPU result.
NCCL (NVLink 5 / NVSwitch) on GPUs; the same code runs on gloo for the CPU tests with an oracle-backed engine.

The fused asynchronous rollout (agent.rollout) has no exchange step: across GPUs it runs as independent
replicas (DESIGN.md "Multi-GPU": a replicated table cannot scale an exact per-step exchange, every replica
must apply every rank's updates).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous global env-id range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TorchEngine:
    """The GPU engine: BatchedGame2048Env + BatchedQLearningAgent of this rank."""

    def __init__(self, env, agent):
        self.env, self.agent = env, agent

    def emit(self):
        return self.agent.step_sync(self.env, mode="deterministic", apply=False, records=True)

    def apply(self, keys, actions, targets):
        self.agent.apply_targets(keys, actions, targets, mode="deterministic")


class ShardedQLearning:
    """Synchronous data-parallel tabular Q-learning: step() = local emit -> all_gather -> apply everywhere."""

    def __init__(self, engine, n_total: int, group=None):
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.sizes = [hi - lo for lo, hi in (shard_range(n_total, r, self.world) for r in range(self.world))]
        self.pad = max(self.sizes)

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t
        if t.numel() < self.pad:  # all_gather needs equal sizes: pad, then cut each rank's tail
            t = torch.cat([t, t.new_zeros(self.pad - t.numel())])
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return torch.cat([o[:n] for o, n in zip(out, self.sizes)])

    def step(self):
        keys, actions, targets = self.engine.emit()
        self.engine.apply(self._gather(keys), self._gather(actions), self._gather(targets))
