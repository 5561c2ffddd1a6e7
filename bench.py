#!/usr/bin/env python
"""bench.py -- env-steps/s and tabular Q-updates/s of the fused 2048 Q-learning hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3"): 2^20 envs per GPU, penalty-flavour env, epsilon-greedy tabular
Q-learning (alpha 0.1, gamma 0.99, epsilon 0.1, Philox seed 0x2048) on an open-addressing hash Q-table in
HBM.  One bench "step" = one fused call g2048_rollout_qlearn advancing every env by ENV_STEPS_PER_LAUNCH
steps (each env step = choose_action + env.step + update_q_value, i.e. one Q-update; every update is applied,
`lost_update_fraction` is 0).  Multi-GPU `value` = weak scaling: every rank runs its own env shard (global env
ids) against its own table replica, no data-path collective; with N > 1 the same line carries `shared_learning`:
BASELINE configs[3] (2^23 envs in total learning ONE table across the GPUs: the exact routed step -- lookups and
records travel to the owner of a state as bulk lists --, the exact owner-computes exchange every step and every 16
steps, and the asynchronous table sharded over NVLink), the exact modes with their table digest compared with the
1-GPU run (DESIGN.md "Multi-GPU"), all timed after 160 warm-up env steps.

Prints ONE JSON line (rank 0).  `value`: device-resident throughput (CUDA events, max over ranks);
`e2e`: the same metric through the host-buffer C-ABI call g2048_ctx_rollout_qlearn with pinned HOST buffers,
host<->device copies inside the timed region; `roofline`: the fused call against the measured HBM peak
(32 algorithmic bytes per env step); `synchronous_step`: the exact batched step (atomic and deterministic apply) at
the same size; `cpu_baseline`: the C oracle port on the host cores (bounded sample of the same training run).
`--impl reference` times that same CPU run as the reference arm (one definition: cpu_reference()).

Optional side measurements (extras of the same line): --exchange [--exchange-envs M] the exact synchronous mode across
GPUs (replicated tables over NCCL / NVLink peer memory, owner-computes on one sharded table, every step and every 16
steps); --shared-table the fused rollout on ONE table sharded over the GPUs' HBM; --dqn BASELINE config 5 (197 M-parameter
network in the loop, data-parallel replay step); --eps the other exploration rate; --no-extras skips the explanatory
single-GPU measurements (random rollout, single-step API, stand-alone update, DQN feed, live HBM random-access peaks).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS_PER_GPU = 1 << 20
ENV_STEPS_PER_LAUNCH = 32
LR, GAMMA, EPS, SEED = 0.1, 0.99, 0.1, 0x2048
ALGO_BYTES_PER_STEP = 32        # key(s') 8 + row(s') 16 + Q(s,a) 4 R + 4 W, slot of s carried
ALGO_BYTES_PER_UPDATE = 40      # stand-alone update API: + key(s) 8
ALGO_BYTES_SINGLE_STEP = 38     # single-step env API, penalty flavour
METRIC = "env-steps/sec & tabular Q-updates/sec"
UNIT = "env-steps/s (1 Q-update per env step)"


def workload_name(n_envs, k):
    return (f"C3: {n_envs} envs/GPU epsilon-greedy tabular Q-learning, penalty env, fused {k}-step rollouts, "
            f"HBM hash Q-table")


def vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per env step (dram__bytes_read.sum + dram__bytes_write.sum of k_rollout_qlearn / its env steps) from
    the committed ncu --set full capture (profiles/r02_traffic.json), or None."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["dram_bytes_per_env_step"])
    except Exception:
        return None


def ncu_rollout_random_pipes():
    """Integer-pipe picture of k_rollout_random<penalty> from the committed ncu --set full capture: the env step is
    ALU-bound, not memory-bound (SURVEY.md 8d 'Roofline -- env step')."""
    try:
        ks = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_final_build.json")))["kernels"]
        k = next(x for x in ks if "k_rollout_random" in x["kernel"])
        num = lambda key: float(str(k[key]).split()[0])
        return {"alu_pipe_active_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                "issue_slots_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "thread_instructions_per_env_step": num("smsp__inst_executed.sum") *
                num("smsp__thread_inst_executed_per_inst_executed.ratio") / (float(1 << 20) * 64),
                "shared_wavefronts_per_warp_step": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") * 32 / (float(1 << 20) * 64),
                "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "source": "profiles/r01_ncu_full_final_build.json (ncu --set full)"}
    except Exception:
        return None


def rmw_peak():
    """Live: random 32-byte load + 4-byte store to the same sector over 8 GiB (tools/membench3) -- the access pattern
    of one Q-table lookup followed by the update of one of its values."""
    import re
    import subprocess
    try:
        out = subprocess.run([os.path.join(ROOT, "tools", "membench3")], capture_output=True, text=True, timeout=120).stdout
        m = re.search(r"mode 3 .*?([0-9.]+) Gops/s", out)
        return float(m.group(1)) * 1e9 if m else None
    except Exception:
        return None


def ncu_dram_ops_per_step():
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["dram_random_ops_per_env_step"])
    except Exception:
        return None


NVLINK_REQUESTS_PER_SEC = 6.69e9   # measured: tools/membench4.cu on 2 B200 (profiles/r01_membench.txt)


def random_access_peak():
    """Live measurement of what HBM gives random slot reads (tools/membench: dependent random 32-byte loads over
    8 GiB at 1024 threads/SM).  Every such miss moves a 128-byte line (profiles/r01_membench.txt)."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "tools", "membench")
    try:
        out = subprocess.run([exe, "8"], capture_output=True, text=True, timeout=120).stdout
        m = re.search(r"fill=1 mode=1 thr/SM=1024 ilp=1 :\s+([0-9.]+) Gops/s", out)
        return float(m.group(1)) * 1e9 if m else None
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons.update(k for k, bit in names.items() if r & bit)
            except Exception:
                pass
            time.sleep(0.0005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on (and allocate pinned host memory from) the CPU cores NVML reports as local to
    this GPU, so that the end-to-end arm's PCIe traffic does not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def fresh_host_envs(L, ctx, n, base, pinned):
    """Pinned host buffers with freshly reset boards (reset itself runs on the GPU through the C ABI)."""
    def alloc(dtype, count):
        nbytes = np.dtype(dtype).itemsize * count
        if pinned:
            p = L.g2048_host_alloc(nbytes)
            assert p, L.g2048_last_error()
            return np.frombuffer((C.c_char * nbytes).from_address(p), dtype=dtype, count=count), p
        return np.zeros(count, dtype), None
    b, pb = alloc(np.uint64, n)
    a, pa = alloc(np.uint64, n)
    s, ps = alloc(np.int32, n)
    a[:] = 0x000000000000FF01
    s[:] = 0
    rc = L.g2048_ctx_env_reset(ctx, vp(b), vp(s), None, None, n, SEED, 0, base)
    assert rc == 0, L.g2048_last_error()
    return (b, a, s), (pb, pa, ps)


CPU_ENVS_PER_THREAD = 2048


def cpu_reference(threads, steps, warmup, k):
    """THE CPU baseline, one definition for `cpu_baseline` (main line) and `--impl reference`: the C oracle's
    sequential-semantics Q-learning (oracle/g2048_oracle.c, main.py:91-101 env after env) as a training run like the
    GPU arm's -- `threads` host threads, each with its own env shard (2,048 envs) and its own PERSISTENT table (the
    most generous CPU figure: no sharing, no locks), k env steps per bench step, `warmup` untimed steps first."""
    import oracle
    oracle.load()
    n = CPU_ENVS_PER_THREAD * threads
    b = np.zeros(n, np.uint64)
    oracle.env_reset(b, None, None, None, seed=SEED)
    a, s = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
    tables = [oracle.QTable(1 << 22, f32=True) for _ in range(threads)]
    for w in range(warmup):
        oracle.rollout_qlearn_mt_tables(b, a, s, k, LR, GAMMA, EPS, tables, 0, SEED, w * k, 0)
    total, t0 = 0, time.perf_counter()
    for i in range(steps):
        total += int(oracle.rollout_qlearn_mt_tables(b, a, s, k, LR, GAMMA, EPS, tables, 0, SEED, (warmup + i) * k, 0)[0])
    dt = time.perf_counter() - t0
    sample = (f"C oracle port (oracle/g2048_oracle.c), {n} envs ({CPU_ENVS_PER_THREAD}/thread) x {k} env steps per bench "
              f"step, {threads} threads each with its own persistent 2^22-slot table, {warmup} warm-up + {steps} timed steps")
    return {"value": total / dt, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "host_cpus": os.cpu_count(), "ms_per_step": dt / max(steps, 1) * 1e3}


def cpu_baseline(n_threads):
    """Bounded sample for the main line (about 10 bench steps of the same run --impl reference times in full)."""
    return cpu_reference(n_threads, steps=10, warmup=2, k=ENV_STEPS_PER_LAUNCH)


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of the reference path (the reference itself is pure Python and does
    not exist on the GPU box), all host threads, the bench step of cpu_reference()."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    k = ENV_STEPS_PER_LAUNCH
    r = cpu_reference(threads, args.steps, args.warmup, k)
    v = r["value"]
    base = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"]}
    emit_result({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 boards / f32 Q rows", "data": "synthetic",
        "config": {"workload": workload_name(N_ENVS_PER_GPU, k), "cpu_sample": r["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_RESULT_FD = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: keep a private copy of it for the result and point fd 1 at
    stderr, so that native libraries that log to stdout (NCCL prints its version there) cannot get in front of it."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_result(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    global EPS
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS_PER_GPU)
    ap.add_argument("--eps", type=float, default=EPS, help="fixed exploration rate (BASELINE config 3 quotes 0.1 and 0.95)")
    ap.add_argument("--no-extras", action="store_true", help="skip the explanatory side measurements")
    ap.add_argument("--no-shared-learning", action="store_true",
                    help="N > 1: skip BASELINE config 4 (2^23 envs learning one table across the GPUs)")
    ap.add_argument("--shared-envs", type=int, default=1 << 23, help="total envs of the shared-learning measurement")
    ap.add_argument("--dqn", action="store_true",
                    help="also time BASELINE config 5: 65,536 nopenalty envs/GPU with the 197 M-parameter DQN forward in the loop")
    ap.add_argument("--exchange", action="store_true",
                    help="also time the synchronous mode with the cross-GPU record exchange (dist.ShardedQLearning)")
    ap.add_argument("--exchange-envs", type=int, default=0,
                    help="envs per GPU for --exchange (default min(--envs, 2^20)); BASELINE config 4 = 2^23 envs in total")
    ap.add_argument("--shared-table", action="store_true",
                    help="also time the fused rollout on ONE Q-table sharded over the GPUs' HBM (NVLink peer loads/atomics)")
    args = ap.parse_args()
    EPS = args.eps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist

    import g2048
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the g2048 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g2048.init(local_rank)
    L = g2048.lib()
    dev = torch.device("cuda", local_rank)
    n, k = args.envs, ENV_STEPS_PER_LAUNCH
    base = rank * n
    launches = args.steps + args.warmup
    # the table must stay under ~0.45 load for the whole run (<= 2^31 slots): long runs take fewer steps per launch
    k = int(max(1, min(k, 0.45 * (1 << 31) / (0.8 * n * launches))))

    # table: as many slots as keep the load factor under ~0.45 for one arm (<= 2^31 slots = 64 GiB)
    expected_inserts = 0.8 * n * k * launches
    free_bytes, _ = torch.cuda.mem_get_info()
    cap = 1 << 22
    while cap < (1 << 31) and cap * 0.45 < expected_inserts and cap * 2 * 32 < free_bytes * 0.6:
        cap <<= 1
    ctx = L.g2048_ctx_create(local_rank, n, cap)
    assert ctx, L.g2048_last_error()
    table = L.g2048_ctx_table(ctx)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_envs():
        b = torch.zeros(n, dtype=torch.int64, device=dev)
        a = torch.full((n,), 0x000000000000FF01, dtype=torch.int64, device=dev)
        s = torch.zeros(n, dtype=torch.int32, device=dev)
        assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, SEED, 0, base, stream) == 0
        return b, a, s

    counters = torch.zeros(16, dtype=torch.int64, device=dev)

    # ---- pre-warm: ramp clocks with random-policy rollouts (not counted) --------------------------------
    b, a, s = device_envs()
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        L.g2048_rollout_random(b.data_ptr(), a.data_ptr(), s.data_ptr(), n, 64, 0, SEED, 0, base, counters.data_ptr(), stream)
        torch.cuda.synchronize()

    # ---- arm 1: device-resident (value) -------------------------------------------------------------------
    b, a, s = device_envs()
    counters.zero_()

    def launch(i):
        rc = L.g2048_rollout_qlearn(b.data_ptr(), a.data_ptr(), s.data_ptr(), table, cap, n, k, 0, LR, GAMMA, EPS, SEED,
                                    i * k, base, counters.data_ptr(), stream)
        assert rc == 0, L.g2048_last_error()

    for i in range(args.warmup):
        launch(i)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    clocks = ClockSampler(local_rank)
    clocks.__enter__()          # sampled through both timed regions (device-resident arm and end-to-end arm)
    ev[0].record()
    for i in range(args.steps):
        launch(args.warmup + i)
        ev[i + 1].record()
    barrier()
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if os.environ.get("G2048_PRINT_LAUNCHES") and rank == 0:
        print("per-launch ms:", [round(x, 3) for x in per_launch_ms], "cap", cap, file=sys.stderr)
    elapsed_s = max_over_ranks(ev[0].elapsed_time(ev[-1]) / 1e3)
    c = counters.cpu().numpy()
    assert int(c[0]) == n * k * launches, (int(c[0]), n * k * launches)
    value = world * n * k * args.steps / elapsed_s
    kernel_ms = statistics.mean(per_launch_ms)
    peak, peak_src = measured_peak()
    achieved = n * k * ALGO_BYTES_PER_STEP / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    pst = torch.zeros(3, dtype=torch.int64, device=dev)
    assert L.g2048_qtable_probe_stats(table, cap, pst.data_ptr(), stream) == 0, L.g2048_last_error()
    pst = pst.tolist()
    table_stats = {"mean_probe_length": 1 + pst[1] / max(pst[0], 1), "max_probe_length": 1 + pst[2],
                   "inserts_per_sec": world * float(c[6]) / launches * args.steps / elapsed_s, "capacity_slots": cap, "table_GiB": cap * 32 / 2**30, "states": int(c[6]),
                   "load_factor_end": int(c[6]) / cap, "dropped": int(c[7]), "lost_updates": int(c[8]), "retried_updates": int(c[9]),
                   "retried_update_fraction": int(c[9]) / max(int(c[0]), 1),
                   "lost_update_fraction": int(c[8]) / max(int(c[0]), 1),
                   "new_state_fraction": int(c[6]) / max(int(c[0]), 1),
                   "valid_fraction": int(c[1]) / max(int(c[0]), 1)}

    # ---- arm 2: end to end through the host-buffer C-ABI call (pinned host memory) --------------------------
    assert L.g2048_ctx_qtable_clear(ctx) == 0
    (hb, ha, hs), pins = fresh_host_envs(L, ctx, n, base, pinned=True)
    hc = np.zeros(16, np.int64)

    def e2e_step(i):
        rc = L.g2048_ctx_rollout_qlearn(ctx, vp(hb), vp(ha), vp(hs), n, k, 0, LR, GAMMA, EPS, SEED, i * k, base, vp(hc))
        assert rc == 0, L.g2048_last_error()

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(args.warmup + i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks.__exit__(None, None, None)
    assert int(hc[0]) == n * k
    e2e = {"value": world * n * k * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 20,
           "d2h_bytes_per_step": n * 20 + 64, "ms_per_step": e2e_s / args.steps * 1e3,
           "api": "g2048_ctx_rollout_qlearn (host boards/aux/score in, boards/aux/score/counters out, pinned)"}
    for p in pins:
        L.g2048_host_free(p)

    sync_step = synchronous_step_measurement(L, torch, dev, n, base, stream) if rank == 0 else None
    shared_learning = None
    if world > 1 and not args.no_shared_learning:
        L.g2048_ctx_destroy(ctx)                   # the 64 GiB replica table makes room for the shared tables
        ctx = None
        torch.cuda.empty_cache()
        try:
            shared_learning = shared_learning_measurement(torch, dist, g2048, dev, rank, world, args.shared_envs,
                                                          max_over_ranks, barrier)
        except Exception as e:   # (a peer barrier that timed out raises on every rank) the main line is still reported
            shared_learning = [{"error": f"{type(e).__name__}: {e}"}]
    extras = {}
    if args.exchange:
        sync = sync_exchange_measurement(torch, dist, g2048, dev, rank, world, args.exchange_envs or min(n, 1 << 20),
                                         max_over_ranks, barrier)
        if rank == 0:
            extras["synchronous_exchange"] = sync
    if args.shared_table:
        sh = shared_table_measurement(torch, dist, g2048, dev, rank, world, min(n, 1 << 20), max_over_ranks, barrier)
        if rank == 0:
            extras["shared_table_over_nvlink"] = sh
    if args.dqn and world > 1:
        dp = dqn_data_parallel_measurement(torch, dist, g2048, dev, rank, world, max_over_ranks, barrier)
        if rank == 0:
            extras["dqn_data_parallel_replay"] = dp
    if args.dqn and rank == 0:
        extras["dqn_in_the_loop"] = dqn_measurement(torch, g2048, dev)
    if not args.no_extras and rank == 0 and ctx is not None:
        extras = dict(extras, **side_measurements(L, torch, dev, ctx, table, cap, n, base, stream, peak))
        rnd = random_access_peak()
        if rnd:
            # random DRAM operations the fused kernel needs per env step (committed ncu capture): line fills
            # (read sectors / 4) + sector write-backs
            per_step = ncu_dram_ops_per_step()
            extras["hbm_random_access"] = {
                "measured_line_fetches_per_sec": rnd, "bytes_moved_per_random_access": 128,
                "equivalent_GBps": rnd * 128 / 1e9, "frac_of_streaming_peak": rnd * 128 / 1e9 / peak,
                "fused_kernel_dram_ops_per_env_step_ncu": per_step,
                "fused_kernel_bound_env_steps_per_sec": (rnd / per_step) if per_step else None,
                "fused_kernel_frac_of_random_access_bound": (value / world * per_step / rnd) if per_step else None,
                "note": "a hash table in HBM is bounded by random line fetches, not by its algorithmic bytes"}
        rmw = rmw_peak()
        if rmw:
            c0 = counters.cpu().numpy()
            valid_frac = float(c0[1]) / max(float(c0[0]), 1.0)     # an invalid move needs no lookup (s' == s)
            extras["hbm_random_read_modify_write"] = {
                "measured_rmw_per_sec": rmw, "rmw_per_env_step": valid_frac,
                "fused_kernel_bound_env_steps_per_sec": rmw / valid_frac,
                "fused_kernel_frac_of_bound": value / world * valid_frac / rmw,
                "note": "one lookup of s' (32-byte sector) + one later 4-byte write into it per valid move"}
    if world > 1:
        dist.barrier()

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 boards / f64 reward / f32 Q rows", "data": "synthetic",
            "config": {"workload": workload_name(n, k), "envs_per_gpu": n, "env_steps_per_launch": k, "alpha": LR,
                       "gamma": GAMMA, "epsilon": EPS, "seed": SEED, "flavour": "penalty",
                       "mode": "fused asynchronous rollout, every update applied (lost races deferred to a grouped apply inside the call)",
                       "parallelism": f"env shards x{world}, table replica per GPU, no collective",
                       "l2": f"Q-table {cap * 32 / 2**30:.0f} GiB >> 126 MB L2 (random 32 B sectors); boards live in registers"},
            "q_updates_per_sec": value * (1.0 - table_stats["lost_update_fraction"]),
            "lost_update_fraction": table_stats["lost_update_fraction"],
            "retried_update_fraction": table_stats["retried_update_fraction"], "table": table_stats,
            "synchronous_step": sync_step, "shared_learning": shared_learning,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic * n * k) if traffic else None, "kernel": "k_rollout_qlearn<penalty>",
                         "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_STEP, "peak_source": peak_src,
                         "request_bound": {"table_visits_per_sec": 17.5e9, "env_steps_per_sec": 17.5e9 / table_stats["valid_fraction"],
                                           "frac_of_bound": value / world * table_stats["valid_fraction"] / 17.5e9,
                                           "source": "tools/membench5.cu: random 32 B load that misses L2 + ONE write-type request"},
                         "note": "achieved = 32 algorithmic B/env-step / duration of the whole call (k_rollout_qlearn = 89 % of "
                                 "it, the deferred-update kernels the rest; CUDA events around the call).  The memory system "
                                 "charges per request, not per byte (profiles/r02_membench.txt): a table visit of one load + "
                                 "one write-type request runs at 17.5 G/s at most = 8.6 % of the streaming peak at 32 B; "
                                 "traffic = ncu DRAM bytes (64 B fill + 32 B write-back per visit)"},
            "e2e": e2e, "gpu_launches": args.steps * 6, "clocks": clocks.summary(),
            "gpu_launches_note": "per fused call: k_rollout_qlearn + k_defer_count, k_defer_scan, k_defer_scan_sums, "
                                 "k_defer_scatter, k_defer_apply (device-resident arm; the e2e arm launches one rollout "
                                 "kernel per chunk on top)",
            "cpu_baseline": cpu_baseline(os.cpu_count() or 1) if world == 1 else None,
            "extras": extras,
        }
        emit_result(out)
    if ctx is not None:
        L.g2048_ctx_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()


def synchronous_step_measurement(L, torch, dev, n, base, stream, warm=64, steps=24):
    """The exact synchronous batched step g2048_qlearn_step (all envs choose and bootstrap on the
    table as it stands at step start; duplicates of one (state, action) applied one after another) at the headline
    size, both apply modes, mid-game (after `warm` steps: the envs have left the 480 start boards)."""
    out = {}
    cap = 1 << 28
    need = int(L.g2048_qlearn_scratch_bytes(n))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    table = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    for mode, name in ((0, "atomic"), (1, "deterministic")):
        table.zero_()
        b = torch.zeros(n, dtype=torch.int64, device=dev)
        a = torch.full((n,), 0x000000000000FF01, dtype=torch.int64, device=dev)
        s = torch.zeros(n, dtype=torch.int32, device=dev)
        assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, SEED, 0, base, stream) == 0

        def step(t):
            rc = L.g2048_qlearn_step(b.data_ptr(), a.data_ptr(), s.data_ptr(), table.data_ptr(), cap, n, 0, LR, GAMMA, EPS,
                                     mode, 1, SEED, t, base, cnt.data_ptr(), None, None, None, scratch.data_ptr(), need, stream)
            assert rc == 0, L.g2048_last_error()
        for t in range(warm):
            step(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(steps):
            step(warm + t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"env_steps_per_sec": n / ms * 1e3, "ms_per_step": ms, "envs": n,
                     "parity": "bit-exact vs the oracle (tests/test_gpu_agent.py)" if mode else
                               "float32 tolerance / hull of {old value, targets} vs the oracle (tests/test_gpu_parity.py)"}
    del table, scratch
    torch.cuda.empty_cache()
    return out


def device_table_digest(L, torch, dev, table_ptr, slots, stream):
    """Order-independent exact digest of a table (or shard) computed on the device: [rows with a non-zero Q value, sum
    of their float32 bit patterns, sum of the low words of their keys].  Replicas and shards differ only in untouched
    zero rows, which are left out."""
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    assert L.g2048_qtable_size(table_ptr, slots, cnt.data_ptr(), stream) == 0, L.g2048_last_error()
    m = int(cnt.item())
    keys = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    rows = torch.empty((max(m, 1), 4), dtype=torch.float32, device=dev)
    cnt.zero_()
    assert L.g2048_qtable_export(table_ptr, slots, keys.data_ptr(), rows.data_ptr(), m, cnt.data_ptr(), stream) == 0
    keys, rows = keys[:m], rows[:m]
    nz = (rows != 0).any(1)
    bits = rows.view(torch.int32).to(torch.int64)
    d = torch.stack([nz.sum(), (bits * nz[:, None]).sum(), ((keys & 0xFFFFFFFF) * nz).sum()])
    del keys, rows, bits
    return d


def shared_learning_measurement(torch, dist, g2048, dev, rank, world, n_total, max_over_ranks, barrier):
    """BASELINE configs[3]: n_total (2^23) envs sharded over the GPUs learn ONE Q-table.  Three modes, ms per env step
    of ALL envs: (1) exact, owner computes, exchange every step; (2) the same with the exchange every 16 steps (values
    frozen inside a window); (3) the asynchronous fused rollout on the table sharded over the GPUs' HBM (remote loads /
    atomics over NVLink inside the kernel).  For the exact modes the digest of the sharded table is compared with the
    digest of the SAME run on one GPU with one table (rank 0 plays all n_total envs with g2048_qlearn_step in
    deterministic mode): `digest_equal_to_1gpu`."""
    from g2048 import dist as gdist
    L = g2048.lib()
    if world & (world - 1) or n_total % world:
        return [{"skipped": "needs a power-of-two number of GPUs"}]
    n = n_total // world
    stream = torch.cuda.current_stream().cuda_stream
    slots_total = 1 << 30
    out = []

    def run_owner(window, warm, steps, routed=False):
        env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
        env.reset()
        shared = gdist.SharedQTable(L, dev, slots_total // world)
        if routed:
            oc = gdist.RoutedQLearning(env, shared, n_total, LR, GAMMA, EPS)
            oc.flush = lambda: 0
        else:
            oc = gdist.OwnerComputesQLearning(env, shared, n_total, LR, GAMMA, EPS, window=window)
        for _ in range(warm):
            oc.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            oc.step()
        e1.record()
        barrier()
        dt = max_over_ranks(e0.elapsed_time(e1) / 1e3)
        oc.flush()
        torch.cuda.synchronize()
        d = device_table_digest(L, torch, dev, shared.ptrs[rank], shared.slots_per_shard, stream)
        dist.all_reduce(d)
        cnt = env.counters.clone()
        dist.all_reduce(cnt)
        oc.close()
        shared.close()
        del env
        torch.cuda.empty_cache()
        return dt, d.tolist(), cnt.tolist()

    def one_gpu_reference(window, total_steps):
        """rank 0: the same run on one GPU, one table (the others wait)"""
        d = None
        if rank == 0:
            cap = slots_total
            table = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
            b = torch.zeros(n_total, dtype=torch.int64, device=dev)
            a = torch.full((n_total,), 0x000000000000FF01, dtype=torch.int64, device=dev)
            s = torch.zeros(n_total, dtype=torch.int32, device=dev)
            cnt = torch.zeros(16, dtype=torch.int64, device=dev)
            assert L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n_total, SEED, 0, 0, stream) == 0
            m = n_total * window
            need = int(L.g2048_qlearn_scratch_bytes(m))
            scratch = torch.empty(need, dtype=torch.uint8, device=dev)
            if window == 1:
                for t in range(total_steps):
                    assert L.g2048_qlearn_step(b.data_ptr(), a.data_ptr(), s.data_ptr(), table.data_ptr(), cap, n_total, 0, LR,
                                               GAMMA, EPS, 1, 1, SEED, t, 0, cnt.data_ptr(), None, None, None,
                                               scratch.data_ptr(), need, stream) == 0, L.g2048_last_error()
            else:   # values frozen for `window` steps: emit the records of every step, apply them all in (step, env) order
                rk = torch.empty(m, dtype=torch.int64, device=dev)
                ra = torch.empty(m, dtype=torch.uint8, device=dev)
                rt = torch.empty(m, dtype=torch.float32, device=dev)
                for t0 in range(0, total_steps, window):
                    for j in range(window):
                        o = j * n_total
                        assert L.g2048_qlearn_step(b.data_ptr(), a.data_ptr(), s.data_ptr(), table.data_ptr(), cap, n_total, 0,
                                                   LR, GAMMA, EPS, 1, 0, SEED, t0 + j, 0, cnt.data_ptr(),
                                                   rk.data_ptr() + 8 * o, ra.data_ptr() + o, rt.data_ptr() + 4 * o,
                                                   None, 0, stream) == 0, L.g2048_last_error()
                    assert L.g2048_qtable_apply_targets(table.data_ptr(), cap, rk.data_ptr(), ra.data_ptr(), rt.data_ptr(), m, LR,
                                                        1, scratch.data_ptr(), need, stream) == 0, L.g2048_last_error()
                del rk, ra, rt
            torch.cuda.synchronize()
            d = device_table_digest(L, torch, dev, table.data_ptr(), cap, stream).tolist()
            del table, scratch, b, a, s
            torch.cuda.empty_cache()
        barrier()
        return d

    refs = {}                 # both every-step modes are compared with the same single-GPU run
    # 160 warm-up env steps like the main arm and the asynchronous mode below: the games have left the 480 start boards
    for window, warm, steps, routed in ((1, 160, 12, True), (1, 160, 12, False), (16, 160, 32, False)):
        dt, digest, cnt = run_owner(window, warm, steps, routed)
        if (window, warm + steps) not in refs:
            refs[(window, warm + steps)] = one_gpu_reference(window, warm + steps)
        ref = refs[(window, warm + steps)]
        if rank == 0:
            name = ("routed, exact, exchange every step (lookups and records travel to the owner as bulk lists; no remote table access)"
                    if routed else f"owner computes, exact, exchange every {window} step" + ("s" if window > 1 else ""))
            out.append({"mode": name,
                        "env_steps_per_sec": n_total * steps / dt, "ms_per_step": dt / steps * 1e3, "envs_total": n_total,
                        "envs_per_gpu": n, "steps_timed": steps, "table_digest": digest, "digest_1gpu": ref,
                        "digest_equal_to_1gpu": digest == ref, "lost": int(cnt[8]), "dropped": int(cnt[7])})
    # (3) asynchronous, one table sharded over the GPUs
    k, warm, launches = 16, 10, 8       # 160 warm-up env steps like the main arm: the games have left the 480 start boards
    env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
    env.reset()
    shared = gdist.SharedQTable(L, dev, (1 << 31) // world)
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(warm):
        tot += shared.rollout(env, k, LR, GAMMA, EPS)
    tot.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        tot += shared.rollout(env, k, LR, GAMMA, EPS)
    e1.record()
    barrier()
    dt = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    dist.all_reduce(tot)
    c = tot.tolist()
    shared.close()
    del env
    torch.cuda.empty_cache()
    if rank == 0:
        out.append({"mode": "asynchronous fused rollout, one table sharded over the GPUs (NVLink loads / atomics in the kernel)",
                    "env_steps_per_sec": n_total * k * launches / dt, "ms_per_step": dt / (k * launches) * 1e3,
                    "envs_total": n_total, "envs_per_gpu": n, "steps_timed": k * launches, "digest_equal_to_1gpu": None,
                    "lost": int(c[8]), "dropped": int(c[7]), "retried_update_fraction": c[9] / max(c[0], 1),
                    "remote_access_fraction": (world - 1) / world})
    return out


def dqn_measurement(torch, g2048, dev, n=65536, steps=3):
    """BASELINE config 5 on one GPU: the env side in one fused launch per step, the reference's 197,204,996-parameter
    CNN (bf16, PyTorch/cuDNN/cuBLAS) evaluated on all 65,536 boards per step -- expected network-bound."""
    from g2048 import dqn
    env = g2048.BatchedGame2048Env(n, "nopenalty", device=dev.index, seed=SEED)
    agent = dqn.BatchedDQNAgent(device=dev.index, dtype=torch.bfloat16, memory_size=1 << 20, seed=SEED)
    env.reset()
    feed = dqn.FusedDQNFeed(env, agent)

    def forward():
        with torch.no_grad():
            agent.model.eval()
            return torch.cat([agent.model(c).float() for c in feed.onehot.split(8192)]).contiguous()

    def one():
        feed.step(forward())

    one()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    for _ in range(steps):
        feed.step(torch.zeros((n, 4), device=dev))
    e2.record()
    torch.cuda.synchronize()
    t_all, t_env = e0.elapsed_time(e1) / steps * 1e-3, e1.elapsed_time(e2) / steps * 1e-3
    flops = 2.0 * n * (64 * (4 * 512 * 30) + 2 * 64 * (2048 * 512 * 30) + 131072 * 1024 + 1024 * 4)
    return {"env_steps_per_sec_with_network": n / t_all, "ms_per_step": t_all * 1e3, "env_side_ms_per_step": t_env * 1e3,
            "network_share": 1 - t_env / t_all, "network_TFLOPs": flops / (t_all - t_env) / 1e12, "envs": n,
            "parameters": sum(p.numel() for p in agent.model.parameters()), "dtype": "bf16"}


def dqn_data_parallel_measurement(torch, dist, g2048, dev, rank, world, max_over_ranks, barrier, n=4096, steps=5):
    """Config 5, training side across GPUs: every rank feeds its own env shard into its replay memory and replay()
    averages the gradients of the 197 M-parameter network with ONE all-reduce of a flat fp32 buffer (789 MB) over
    NVLink (BatchedDQNAgent.data_parallel / dist.GradientAllReduce)."""
    from g2048 import dqn
    env = g2048.BatchedGame2048Env(n, "nopenalty", device=dev.index, seed=SEED, env_id_base=rank * n)
    agent = dqn.BatchedDQNAgent(device=dev.index, memory_size=1 << 16, batch_size=64, epsilon=0.5, seed=SEED + rank)
    env.reset()
    agent.data_parallel()
    for _ in range(4):
        dqn.dqn_step(env, agent)
    agent.replay()
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps):
        agent.replay()
    e1.record()
    for _ in range(steps):
        agent.grad_sync()
    e2.record()
    barrier()
    t_replay = max_over_ranks(e0.elapsed_time(e1) / steps)
    t_ar = max_over_ranks(e1.elapsed_time(e2) / steps)
    nbytes = agent.grad_sync.flat.numel() * agent.grad_sync.flat.element_size()
    p0 = torch.cat([p.detach().flatten()[:64] for p in agent.model.parameters()]).double().sum()
    allp = [torch.zeros_like(p0) for _ in range(world)]
    dist.all_gather(allp, p0)
    del agent, env
    torch.cuda.empty_cache()
    return {"ms_per_replay_step_batch64": t_replay, "ms_allreduce_alone": t_ar, "gradient_bytes": nbytes,
            "allreduce_bus_GBps": nbytes * 2 * (world - 1) / world / (t_ar * 1e-3) / 1e9,
            "replicas_identical_after_training": len({float(x) for x in allp}) == 1, "dtype": "fp32"}


def sync_exchange_measurement(torch, dist, g2048, dev, rank, world, n, max_over_ranks, barrier, steps=24):
    """Synchronous data-parallel Q-learning: every step all ranks exchange their packed (state, action, target)
    records and apply the whole list deterministically, so the replicas stay identical.
    Two transports: "peer" = records read in place from the owner's HBM over NVLink inside the apply kernel after a
    flag barrier in peer memory (no collective); "nccl" = one all_gather_into_tensor, then the same kernel."""
    from g2048 import dist as gdist
    out = {"envs_per_gpu": n, "records_per_step": world * n, "bytes_pulled_per_rank_per_step": (world - 1) * n * 16,
           "mode": "deterministic apply of all ranks' records on every replica"}
    transports = ["nccl"] + (["peer"] if world > 1 else [])
    for transport in transports:
        env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
        agent = g2048.BatchedQLearningAgent(1000, 4, LR, GAMMA, EPS, capacity=1 << 29, device=dev.index, seed=SEED)
        env.reset()
        sh = gdist.ShardedQLearning(gdist.TorchEngine(env, agent), n * world, transport=transport)
        for _ in range(3):
            sh.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            sh.step()
        e1.record()
        barrier()
        dt = max_over_ranks(e0.elapsed_time(e1) / 1e3)
        keys, rows = agent.export()
        nz = np.abs(rows).sum(1) > 0          # replicas differ only in untouched zero rows (their own shard's lookups)
        digest = float(rows[nz].astype(np.float64).sum())   # sorted by key, same rows on every replica: same order
        digests = [[digest, float(nz.sum())]]
        if world > 1:
            t = torch.tensor([digest, float(nz.sum())], dtype=torch.float64, device=dev)
            all_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            digests = [x.tolist() for x in all_t]
        sh.close()
        del sh, env, agent
        torch.cuda.empty_cache()
        out[transport] = {"env_steps_per_sec": world * n * steps / dt, "ms_per_step": dt / steps * 1e3,
                          "replica_digests_sum_and_count_of_nonzero_rows": digests,
                          "replicas_identical": len({tuple(d) for d in digests}) == 1}
    if world > 1 and world & (world - 1) == 0:
        # owner computes: ONE table sharded over the GPUs, every GPU sorts/applies only the records for its shard
        env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
        env.reset()
        shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 29) // world)
        oc = gdist.OwnerComputesQLearning(env, shared, n * world, LR, GAMMA, EPS)
        for _ in range(3):
            oc.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            oc.step()
        e1.record()
        barrier()
        dt = max_over_ranks(e0.elapsed_time(e1) / 1e3)
        keys, rows = shared.export_local()
        nz = np.abs(rows).sum(1) > 0
        t = torch.tensor([float(rows[nz].astype(np.float64).sum()), float(nz.sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        oc.close()
        shared.close()
        del env
        torch.cuda.empty_cache()
        out["owner_computes"] = {"env_steps_per_sec": world * n * steps / dt, "ms_per_step": dt / steps * 1e3,
                                 "table_digest_sum_and_count_of_nonzero_rows": t.tolist(),
                                 "note": "one sharded table; digest = sum over all shards, to compare with one replica's"}
        # the same with the exchange every 16 steps (values frozen inside a window)
        env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
        env.reset()
        shared = gdist.SharedQTable(g2048.lib(), dev, (1 << 29) // world)
        oc = gdist.OwnerComputesQLearning(env, shared, n * world, LR, GAMMA, EPS, window=16)
        for _ in range(16):
            oc.step()
        barrier()
        e0.record()
        for _ in range(32):
            oc.step()
        e1.record()
        barrier()
        dt = max_over_ranks(e0.elapsed_time(e1) / 1e3)
        oc.close()
        shared.close()
        del env
        torch.cuda.empty_cache()
        out["owner_computes_window_16"] = {"env_steps_per_sec": world * n * 32 / dt, "ms_per_step": dt / 32 * 1e3,
                                           "note": "exchange + apply once per 16 env steps"}
    if "peer" in out:
        out["transports_agree"] = (out["peer"]["replica_digests_sum_and_count_of_nonzero_rows"] ==
                                   out["nccl"]["replica_digests_sum_and_count_of_nonzero_rows"])
    return out


def shared_table_measurement(torch, dist, g2048, dev, rank, world, n, max_over_ranks, barrier, launches=8, warm=2, k=16):
    launches = int(os.environ.get('G2048_SHARED_LAUNCHES', launches))
    """The fused asynchronous rollout of every rank on ONE Q-table: shard j of the slot range in the HBM of GPU j,
    remote lookups / compare-and-swap updates over NVLink inside the kernel (dist.SharedQTable).  (world-1)/world of
    all table accesses are remote.  Table = 2^31 slots (64 GiB) in total."""
    from g2048 import dist as gdist
    if world & (world - 1):
        return {"skipped": "needs a power-of-two number of GPUs"}
    slots = (1 << 31) // world
    env = g2048.BatchedGame2048Env(n, "penalty", device=dev.index, seed=SEED, env_id_base=rank * n)
    env.reset()
    if world > 1:
        shared = gdist.SharedQTable(g2048.lib(), dev, slots)
    else:
        shared = gdist.SharedQTable(g2048.lib(), dev, slots, shards=[torch.zeros(slots * 4, dtype=torch.int64, device=dev)])
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(warm):
        tot += shared.rollout(env, k, LR, GAMMA, EPS)      # (also loads torch's add kernel outside the timed region)
    tot.zero_()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
    ev[0].record()
    for i in range(launches):
        tot += shared.rollout(env, k, LR, GAMMA, EPS)
        ev[i + 1].record()
    barrier()
    dt = max_over_ranks(ev[0].elapsed_time(ev[-1]) / 1e3)
    per_launch = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(launches)]
    c = tot.cpu().numpy()
    states = shared.size()
    shared.close()
    # table requests of one env step: one update CAS, a lookup (x probes) when the move changed the board or the game
    # was reset, an insert CAS per new state; (world-1)/world of them cross NVLink, which carries
    # NVLINK_REQUESTS_PER_SEC small requests per GPU whatever their kind (tools/membench4.cu, profiles/r01_membench.txt)
    req = (float(c[0]) + float(c[1]) + float(c[2]) + float(c[6])) / max(float(c[0]), 1.0)
    remote = (world - 1) / world
    bound = NVLINK_REQUESTS_PER_SEC / (req * remote) if remote else None
    out = {"env_steps_per_sec": world * n * k * launches / dt, "per_gpu": n * k * launches / dt, "ms_per_launch": dt / launches * 1e3,
           "table_requests_per_env_step": req, "nvlink_small_requests_per_sec_per_gpu": NVLINK_REQUESTS_PER_SEC,
           "nvlink_bound_env_steps_per_sec_per_gpu": bound,
           "frac_of_nvlink_request_bound": (n * k * launches / dt / bound) if bound else None,
           "rank0_ms_of_each_launch": per_launch,
           "envs_per_gpu": n, "env_steps_per_launch": k, "table_slots_total": 1 << 31, "slots_per_gpu": slots,
           "states_in_table": states, "load_factor_end": states / float(1 << 31), "remote_access_fraction": (world - 1) / world,
           "rank0_lost_update_fraction": float(c[8]) / max(float(c[1]), 1.0), "rank0_dropped": int(c[7]),
           "mode": "one table for all GPUs, async single-shot CAS updates at the owner's L2, no exchange step"}
    del env
    torch.cuda.empty_cache()
    return out


def side_measurements(L, torch, dev, ctx, table, cap, n, base, stream, peak):
    """Explanatory numbers beside the headline: random-policy rollout, single-step env API, stand-alone update."""
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    b = torch.zeros(n, dtype=torch.int64, device=dev)
    a = torch.full((n,), 0x000000000000FF01, dtype=torch.int64, device=dev)
    s = torch.zeros(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    L.g2048_env_reset(b.data_ptr(), s.data_ptr(), None, None, n, SEED, 0, base, stream)
    # (a) fused random-policy rollout, 64 steps per launch
    for flavour, name in ((0, "penalty"), (1, "nopenalty")):
        dt = timed(lambda: L.g2048_rollout_random(b.data_ptr(), a.data_ptr(), s.data_ptr(), n, 64, flavour, SEED, 0, base,
                                                  cnt.data_ptr(), stream), 10)
        out[f"random_policy_rollout_{name}"] = {"env_steps_per_sec": n * 64 / dt, "ms_per_launch": dt * 1e3}
        if flavour == 0:
            out[f"random_policy_rollout_{name}"]["bound"] = "integer ALU pipe"
            out[f"random_policy_rollout_{name}"]["ncu"] = ncu_rollout_random_pipes()
    # (b) single-step env API, state in HBM every call (HBM bound: 38 B/step)
    act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
    r64 = torch.zeros(n, dtype=torch.float64, device=dev)
    fl = torch.zeros(n, dtype=torch.uint8, device=dev)
    ml = torch.zeros(n, dtype=torch.uint8, device=dev)
    dt = timed(lambda: L.g2048_env_step(b.data_ptr(), a.data_ptr(), s.data_ptr(), act.data_ptr(), None, r64.data_ptr(),
                                        None, fl.data_ptr(), ml.data_ptr(), None, n, 0, SEED, 1, base, stream), 50)
    gbs = n * ALGO_BYTES_SINGLE_STEP / dt / 1e9
    out["single_step_env_api"] = {"env_steps_per_sec": n / dt, "achieved_GBps_at_38B": gbs, "frac_of_hbm_peak": gbs / peak,
                                  "note": f"{n} envs = {n * 38 / 1e6:.0f} MB per call, L2-resident"}
    # (b2) the same at 2^24 envs: 640 MB of state per call, far beyond L2 -- the HBM-resident case
    nb = 1 << 24
    bb = torch.zeros(nb, dtype=torch.int64, device=dev)
    ab = torch.full((nb,), 0x000000000000FF01, dtype=torch.int64, device=dev)
    sb = torch.zeros(nb, dtype=torch.int32, device=dev)
    L.g2048_env_reset(bb.data_ptr(), sb.data_ptr(), None, None, nb, SEED, 0, base, stream)
    actb = torch.randint(0, 4, (nb,), dtype=torch.uint8, device=dev)
    rb = torch.zeros(nb, dtype=torch.float64, device=dev)
    fb = torch.zeros(nb, dtype=torch.uint8, device=dev)
    mb = torch.zeros(nb, dtype=torch.uint8, device=dev)
    dt = timed(lambda: L.g2048_env_step(bb.data_ptr(), ab.data_ptr(), sb.data_ptr(), actb.data_ptr(), None, rb.data_ptr(),
                                        None, fb.data_ptr(), mb.data_ptr(), None, nb, 0, SEED, 1, base, stream), 10)
    gbs = nb * ALGO_BYTES_SINGLE_STEP / dt / 1e9
    out["single_step_env_api_16M_envs"] = {"env_steps_per_sec": nb / dt, "achieved_GBps_at_38B": gbs,
                                           "frac_of_hbm_peak": gbs / peak, "note": "640 MB of env state per call (HBM-resident)"}
    del bb, ab, sb, actb, rb, fb, mb
    # (c) stand-alone batched update on random transitions over a pre-filled table (HBM bound: 40 B/update)
    m = 1 << 22
    L.g2048_ctx_qtable_clear(ctx)
    pool = torch.randint(1, 1 << 62, (min(cap // 4, 1 << 27),), dtype=torch.int64, device=dev)
    idx = torch.randint(0, pool.numel(), (m,), device=dev)
    idx2 = torch.randint(0, pool.numel(), (m,), device=dev)
    s1, s2 = pool[idx].contiguous(), pool[idx2].contiguous()
    aa = torch.randint(0, 4, (m,), dtype=torch.uint8, device=dev)
    rr = torch.randn(m, dtype=torch.float32, device=dev)
    dd = torch.zeros(m, dtype=torch.uint8, device=dev)
    need = int(L.g2048_qlearn_scratch_bytes(m))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    rows = torch.empty((1 << 22, 4), dtype=torch.float32, device=dev)
    for lo in range(0, pool.numel(), 1 << 22):  # insert the whole pool
        chunk = pool[lo:lo + (1 << 22)]
        L.g2048_qtable_lookup(table, cap, chunk.data_ptr(), chunk.numel(), rows.data_ptr(), None, 1, stream)
    for mode, name in ((0, "atomic"), (1, "deterministic")):
        dt = timed(lambda: L.g2048_qtable_update(table, cap, s1.data_ptr(), aa.data_ptr(), rr.data_ptr(), s2.data_ptr(),
                                                 dd.data_ptr(), m, LR, GAMMA, mode, scratch.data_ptr(), need, stream), 10)
        gbs = m * ALGO_BYTES_PER_UPDATE / dt / 1e9
        out[f"standalone_q_update_{name}"] = {"updates_per_sec": m / dt, "achieved_GBps_at_40B": gbs,
                                              "frac_of_hbm_peak": gbs / peak, "batch": m, "table_states": pool.numel()}
    L.g2048_ctx_qtable_clear(ctx)
    # (c2) the headline workload at the other exploration rate BASELINE config 3 quotes (epsilon 0.95), fresh table
    be = torch.zeros(n, dtype=torch.int64, device=dev)
    ae = torch.full((n,), 0x000000000000FF01, dtype=torch.int64, device=dev)
    se = torch.zeros(n, dtype=torch.int32, device=dev)
    L.g2048_env_reset(be.data_ptr(), se.data_ptr(), None, None, n, SEED, 0, base, stream)
    ce = torch.zeros(16, dtype=torch.int64, device=dev)
    ke, launch_no = 16, [0]

    def rollout_095():
        L.g2048_rollout_qlearn(be.data_ptr(), ae.data_ptr(), se.data_ptr(), table, cap, n, ke, 0, LR, GAMMA, 0.95, SEED,
                               launch_no[0] * ke, base, ce.data_ptr(), stream)
        launch_no[0] += 1
    for _ in range(3):
        rollout_095()
    dt = timed(rollout_095, 12)
    cz = ce.cpu().numpy()
    out["fused_rollout_epsilon_0.95"] = {"env_steps_per_sec": n * ke / dt, "ms_per_launch": dt * 1e3,
                                         "new_state_fraction": float(cz[6]) / max(float(cz[0]), 1.0),
                                         "lost_update_fraction": float(cz[8]) / max(float(cz[0]), 1.0), "dropped": int(cz[7]),
                                         "load_factor_end": float(cz[6]) / cap}
    L.g2048_ctx_qtable_clear(ctx)
    # (d) BASELINE config 5, env side only: 65,536 nopenalty envs feeding a DQN -- per step: select_action on
    # (placeholder) network outputs with the legal mask, env step, one-hot encode of the new boards
    m5 = 65536
    b5 = torch.zeros(m5, dtype=torch.int64, device=dev)
    s5 = torch.zeros(m5, dtype=torch.int32, device=dev)
    L.g2048_env_reset(b5.data_ptr(), s5.data_ptr(), None, None, m5, SEED, 0, base, stream)
    qv = torch.randn((m5, 4), dtype=torch.float32, device=dev)
    a5 = torch.zeros(m5, dtype=torch.uint8, device=dev)
    f5 = torch.zeros(m5, dtype=torch.uint8, device=dev)
    lm5 = torch.full((m5,), 15, dtype=torch.uint8, device=dev)
    r5 = torch.zeros(m5, dtype=torch.float32, device=dev)
    dm5 = torch.zeros(m5, dtype=torch.uint8, device=dev)
    step_no = [0]
    for dt_name, code, width in (("f32", 0, 4), ("bf16", 1, 2)):
        enc = torch.empty((m5, 16, 4, 4), dtype=torch.float32 if code == 0 else torch.bfloat16, device=dev)

        def feed():
            t = step_no[0]
            L.g2048_select_action(qv.data_ptr(), lm5.data_ptr(), a5.data_ptr(), m5, 0.1, SEED, t, base, stream)
            L.g2048_env_step(b5.data_ptr(), None, s5.data_ptr(), a5.data_ptr(), None, None, r5.data_ptr(), f5.data_ptr(),
                             None, None, m5, 1, SEED, t, base, stream)
            torch.bitwise_right_shift(f5, 4, out=lm5)
            torch.bitwise_and(f5, 4, out=dm5)   # done envs start a new game
            L.g2048_env_reset(b5.data_ptr(), s5.data_ptr(), dm5.data_ptr(), None, m5, SEED, t + 1, base, stream)
            L.g2048_encode_onehot(b5.data_ptr(), enc.data_ptr(), m5, code, stream)
            step_no[0] += 1
        dt = timed(feed, 50)
        out[f"dqn_feed_65536_envs_{dt_name}"] = {"env_steps_per_sec": m5 / dt, "us_per_step": dt * 1e6,
                                                "onehot_GBps": m5 * 256 * width / dt / 1e9,
                                                "kernels_per_step": "select_action + env_step(nopenalty) + masked reset + encode_onehot"}

        def fused():
            t = step_no[0]
            L.g2048_dqn_env_step(b5.data_ptr(), s5.data_ptr(), qv.data_ptr(), lm5.data_ptr(), a5.data_ptr(), None, None,
                                 r5.data_ptr(), dm5.data_ptr(), lm5.data_ptr(), enc.data_ptr(), code, m5, 0.1, 3, SEED, t,
                                 t + 1, base, stream)
            step_no[0] += 1
        dt = timed(fused, 50)
        out[f"dqn_feed_65536_envs_{dt_name}_fused"] = {"env_steps_per_sec": m5 / dt, "us_per_step": dt * 1e6,
                                                      "onehot_GBps": m5 * 256 * width / dt / 1e9,
                                                      "kernels_per_step": "g2048_dqn_env_step (one launch)"}
    return out


if __name__ == "__main__":
    main()
