"""The three ways to run tabular Q-learning on several GPUs of one box (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 \\
        examples/multi_gpu_modes.py

  replicas   every GPU learns its own table (agent.rollout), nothing is exchanged
  exact      synchronous steps; every replica applies every rank's (state, action, target) records, which it reads
             in place from the owner's HBM over NVLink peer memory -> replicas identical to the 1-GPU result
  shared     ONE table sharded over the GPUs' HBM; the fused rollout reads and updates remote slots over NVLink
  owner      exact synchronous steps on that shared table: records are routed to the GPU that owns the state's slot,
             every GPU sorts and applies only its share -> same table as "exact", 1/G of the apply work per GPU
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g2048  # noqa: E402
from g2048 import dist as gdist  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_total = 1 << 18
    lo, hi = gdist.shard_range(n_total, rank, world)

    def fresh():
        env = g2048.BatchedGame2048Env(hi - lo, "penalty", device=local, seed=1, env_id_base=lo)
        env.reset()
        return env

    # 1. independent replicas
    env, agent = fresh(), g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.1, capacity=1 << 24, device=local, seed=1)
    c = env.counters_dict(agent.rollout(env, 64))
    print(f"[rank {rank}] replicas: {c['steps']} steps, {len(agent)} states in this GPU's table")

    # 2. exact synchronous exchange
    env, agent = fresh(), g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.1, capacity=1 << 24, device=local, seed=1)
    sh = gdist.ShardedQLearning(gdist.TorchEngine(env, agent), n_total, transport="peer" if world > 1 else "nccl")
    for _ in range(16):
        sh.step()
    keys, rows = agent.export()
    print(f"[rank {rank}] exact: table digest {float(rows.astype('float64').sum()):.6f} (equal on every rank)")
    sh.close()

    # 3. one table for the whole box
    env = fresh()
    if world > 1 and world & (world - 1) == 0:
        shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 24) // world)
    else:
        shared = gdist.SharedQTable(g2048.lib(), dev, 1 << 24, shards=[torch.zeros(4 << 24, dtype=torch.int64, device=dev)])
    shared.rollout(env, 64, 0.1, 0.99, 0.1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    print(f"[rank {rank}] shared: {shared.local_size()} states in this GPU's shard, {shared.size()} in the table")
    shared.close()

    # 4. exact synchronous steps on the shared table, owner computes (needs one shard per rank: world = 2^k > 1)
    if world > 1 and world & (world - 1) == 0:
        env = fresh()
        shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 24) // world)
        oc = gdist.OwnerComputesQLearning(env, shared, n_total, 0.1, 0.99, 0.1)
        for _ in range(16):
            oc.step()
        torch.cuda.synchronize()
        dist.barrier()
        _, rows = shared.export_local()
        t = torch.tensor([float(rows.astype("float64").sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        print(f"[rank {rank}] owner: table digest {float(t):.6f} (the same number as 'exact')")
        oc.close()
        shared.close()
        # the same exact step routed: no GPU touches another GPU's shard, only bulk lists cross NVLink
        env = fresh()
        shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 24) // world)
        rq = gdist.RoutedQLearning(env, shared, n_total, 0.1, 0.99, 0.1)
        for _ in range(16):
            rq.step()
        torch.cuda.synchronize()
        dist.barrier()
        _, rows = shared.export_local()
        t = torch.tensor([float(rows.astype("float64").sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        print(f"[rank {rank}] routed: table digest {float(t):.6f} (the same number again)")
        rq.close()
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
