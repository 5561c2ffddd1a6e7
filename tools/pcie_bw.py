"""Host <-> device copy bandwidth of every GPU of the box at the same time (pinned memory, one process per GPU under
torchrun): what the end-to-end arm of bench.py can get from the host side.  Prints one line per rank."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
if "--no-bind" not in sys.argv:
    bench.bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mb = 64
h_in = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return mb * reps / 1024 / (time.perf_counter() - t0)


run(True, True, 3)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"rank {rank} cpus {sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))}) "
      f"H2D {a:.1f} GiB/s  D2H {b:.1f} GiB/s  both {c:.1f} + {c:.1f} GiB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
