"""Times the routed exact step (dist.RoutedQLearning) under torchrun, with G2048_ROUTED_PROFILE=1 the per-phase device
times: torchrun --nproc-per-node N tools/routed_timing.py [n_total_log2] [warm] [steps]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g2048  # noqa: E402
from g2048 import dist as gdist  # noqa: E402

lg, warm, steps = (int(x) for x in (sys.argv[1:] + ["23", "4", "12"])[:3])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if os.environ.get("G2048_ROUTED_PROFILE") and rank not in (0, world // 2):   # two ranks are enough
    del os.environ["G2048_ROUTED_PROFILE"]
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_total = 1 << lg
n = n_total // world
env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=8264, env_id_base=rank * n)
env.reset()
shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 30) // world)
rq = gdist.RoutedQLearning(env, shared, n_total, 0.1, 0.99, 0.1)
for _ in range(warm):
    rq.step()
t = torch.zeros(1, device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    rq.step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"routed step, {world} GPUs, 2^{lg} envs, steps {warm}..{warm + steps}: {ms.item():.3f} ms per step = "
          f"{n_total / ms.item() / 1e6:.2f} G env-steps/s")
rq.close()
shared.close()
dist.destroy_process_group()
