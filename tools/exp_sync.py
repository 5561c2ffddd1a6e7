"""Kernel experiment driver: the exact synchronous step g2048_qlearn_step (deterministic) at bench size, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import g2048
n = int(os.environ.get("N", 1 << 20)); steps = int(os.environ.get("STEPS", 72)); mode = int(os.environ.get("MODE", 1))
g2048.init(0); L = g2048.lib(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
cap = 1 << 28
table = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
need = int(L.g2048_qlearn_scratch_bytes(n)); scratch = torch.empty(need, dtype=torch.uint8, device=dev)
b = torch.zeros(n, dtype=torch.int64, device=dev); a = torch.full((n,), 0xFF01, dtype=torch.int64, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev); cnt = torch.zeros(16, dtype=torch.int64, device=dev)
assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, 0x2048, 0, 0, st) == 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for t in range(steps):
    assert L.g2048_qlearn_step(b.data_ptr(), a.data_ptr(), s.data_ptr(), table.data_ptr(), cap, n, 0, 0.1, 0.99, 0.1, mode, 1, 0x2048, t, 0,
                               cnt.data_ptr(), None, None, None, scratch.data_ptr(), need, st) == 0
    ev[t + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("first steps", [round(x, 3) for x in ms[:6]], "last steps", [round(x, 3) for x in ms[-6:]])
