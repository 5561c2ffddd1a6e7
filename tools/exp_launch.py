"""Kernel experiment driver (not part of the product): per-launch times of the fused rollout at bench size, with the
per-warp time statistics an EXP_TIMING build leaves in counters[10..13]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import g2048
n = int(os.environ.get("N", 1 << 20)); k = int(os.environ.get("K", 16)); cap = 1 << int(os.environ.get("CAPLOG", 30))
launches = int(os.environ.get("LAUNCHES", 25)); sync = int(os.environ.get("SYNC", 1)); eps = float(os.environ.get("EPS", 0.1))
g2048.init(0); L = g2048.lib(); dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
b = torch.zeros(n, dtype=torch.int64, device=dev); a = torch.full((n,), 0xFF01, dtype=torch.int64, device=dev)
s = torch.zeros(n, dtype=torch.int32, device=dev)
assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, 0x2048, 0, 0, st) == 0
table = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
cs = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(launches)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
torch.cuda.synchronize(); ev[0].record()
for i in range(launches):
    rc = L.g2048_rollout_qlearn(b.data_ptr(), a.data_ptr(), s.data_ptr(), table.data_ptr(), cap, n, k, 0, 0.1, 0.99, eps, 0x2048, i * k, 0, cs[i].data_ptr(), st)
    assert rc == 0
    ev[i + 1].record()
    if sync: torch.cuda.synchronize()
torch.cuda.synchronize()
for i in range(launches):
    c = cs[i].tolist(); ms = ev[i].elapsed_time(ev[i + 1])
    w = max(c[13], 1)
    print(f"launch {i:2d}: {ms:6.3f} ms  {n*k/ms/1e6:6.2f} G/s  valid {c[1]/c[0]:.3f} new {c[6]/c[0]:.3f} retried {c[9]/c[0]:.3f} episodes {c[2]}"
          + (f"  rollout time max {c[10]/1e3:.0f} mean {c[11]/w/1e3:.0f} us" if c[13] else ""))
