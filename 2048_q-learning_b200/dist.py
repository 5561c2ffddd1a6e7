"""Env shards over the GPUs of one box + the ways the GPUs learn together (one process per GPU, torch.distributed).

Envs are independent (the reference has exactly one, main.py:66), so rank r owns the contiguous global env
ids [lo, hi) and its Philox draws are keyed by the GLOBAL env id: results do not depend on the sharding.

  ShardedQLearning        replicated tables, exact: the (state key, action, target) records of one synchronous step
                          (16 B per transition) are exchanged -- transport "peer": every rank's apply kernel reads
                          each record straight from its owner over NVLink 5 / NVSwitch (CUDA-IPC buffers, flag
                          barrier, no collective call); "nccl": one all_gather_into_tensor; "auto": generic engines
                          (the oracle on gloo in the CPU tests) -- and every replica applies the whole list with the
                          same deterministic kernel.  Cannot scale (every replica applies every record); kept for comparison.
  SharedQTable            ONE table for the box, shard j in the HBM of rank j; the fused asynchronous rollout reaches
                          remote slots itself (NVLink loads and atomics inside the kernel).
  OwnerComputesQLearning  exact synchronous steps on that table: lookups are remote, each record goes to the owner of
                          its slot, every GPU sorts and applies only its share (optionally every K steps).
  RoutedQLearning         the same exact step with every table access LOCAL: keys travel to the owner as bulk lists,
                          {slot, max Q} and rows come back, records are pushed into the owner's sort input in env
                          order (g2048_routed_*) -- the fastest exact mode.
  GradientAllReduce       data-parallel DQN: one flat gradient buffer, one NCCL all-reduce per training step.

The fused asynchronous rollout on a LOCAL table (agent.rollout) has no exchange step: across GPUs it runs as independent
replicas.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous global env-id range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _share_device_memory(lib, device, nbytes: int, group=None):
    """Allocate `nbytes` of zeroed device memory on this rank, exchange the CUDA-IPC handles through the process group
    (any backend) and map every peer's allocation.  Returns (own pointer, [pointer of rank r's allocation for all r],
    [pointers that must be g2048_peer_close'd])."""
    import ctypes
    from ._lib import check
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
    with torch.cuda.device(device):
        check(lib.g2048_peer_alloc(nbytes, ctypes.byref(mine), handle), "g2048_peer_alloc")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        base, opened = [], []
        for r in range(world):
            if r == rank:
                base.append(mine.value)
                continue
            p = ctypes.c_void_p()
            check(lib.g2048_peer_open((ctypes.c_ubyte * 64)(*handles[r]), ctypes.byref(p)), "g2048_peer_open")
            base.append(p.value)
            opened.append(p.value)
    dist.barrier(group=group)          # every rank has mapped every buffer before anyone touches them
    return mine.value, base, opened


def _unshare_device_memory(lib, device, mine, opened, group=None):
    with torch.cuda.device(device):
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=group)   # nobody unmaps while a peer may still read
        for p in opened:
            lib.g2048_peer_close(p)
        if mine:
            lib.g2048_peer_free(mine)


class TorchEngine:
    """The GPU engine: BatchedGame2048Env + BatchedQLearningAgent of this rank."""

    def __init__(self, env, agent):
        self.env, self.agent = env, agent

    def emit(self):
        return self.agent.step_sync(self.env, mode="deterministic", apply=False, records=True)

    def apply(self, keys, actions, targets):
        self.agent.apply_targets(keys, actions, targets, mode="deterministic")

    # packed 16-byte records (g2048_qlearn_emit / g2048_qtable_apply_records)
    def emit_records(self, records):
        self.agent.emit_records(self.env, records)

    def apply_records(self, lists, counts):
        self.agent.apply_records(lists, counts, mode="deterministic")


class PeerRecordBuffers:
    """Per rank one CUDA-IPC allocation [flags: 256 B][records slot 0][records slot 1], mapped by every other rank
    of the box.  Slots alternate per step, so a rank may write step t+1's records while a slower peer still reads
    step t's; the flag barrier of step t+1 is what releases slot t for reuse at step t+2."""

    FLAG_BYTES = 256

    def __init__(self, lib, device: torch.device, n_max: int, group=None):
        import ctypes
        self.lib, self.device, self.n_max, self.group = lib, device, int(n_max), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("peer exchange supports up to 16 GPUs of one box")
        self.slot_bytes = ((self.n_max * 16 + 255) // 256) * 256
        nbytes = self.FLAG_BYTES + 2 * self.slot_bytes
        self._mine, self.base, self._opened = _share_device_memory(lib, device, nbytes, group)
        self._flags = (ctypes.c_void_p * self.world)(*self.base)
        # the barrier's time-out flag lives in pinned host memory: the kernel stores into it, the host reads it at every
        # step without a synchronisation, and the apply kernels look at it before they consume a peer's records
        self.timed_out = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.epoch = 0

    def _check(self, rc, what):
        from ._lib import check
        check(rc, what)

    def records(self, rank: int, slot: int) -> int:
        return self.base[rank] + self.FLAG_BYTES + (slot & 1) * self.slot_bytes

    def barrier(self, timeout_ns: int = 0):
        """Stream-ordered barrier over all ranks (k_peer_barrier); returns immediately on the host."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            self._check(self.lib.g2048_peer_barrier(self._flags, self.rank, self.world, self.epoch, timeout_ns,
                                                    self.timed_out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "g2048_peer_barrier")

    def check_timeout(self):
        """Raises once a barrier of this rank has timed out (a peer stalled for more than the time-out): the exchange
        is over -- the device side already ignores the peers' records (g2048_peer_barrier in include/g2048.h)."""
        v = int(self.timed_out[0])
        if v:
            raise RuntimeError(f"peer barrier timed out waiting for rank {v - 1}: the ranks are no longer in step")

    def close(self):
        _unshare_device_memory(self.lib, self.device, self._mine, self._opened, self.group)
        self._mine, self._opened = None, []


class ShardedQLearning:
    """Synchronous data-parallel tabular Q-learning: step() = local emit -> exchange -> apply everywhere.

    transport: "auto" (generic emit()/apply() engines, e.g. the oracle engine on gloo), "nccl" (packed records,
    one all_gather_into_tensor) or "peer" (NVLink peer memory, no collective on the data path)."""

    def __init__(self, engine, n_total: int, group=None, transport: str = "auto"):
        self.engine, self.group, self.transport = engine, group, transport
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.sizes = [hi - lo for lo, hi in (shard_range(n_total, r, self.world) for r in range(self.world))]
        self.pad = max(self.sizes)
        self.t = 0
        self.peers = None
        if transport == "peer":
            agent = engine.agent
            self.peers = PeerRecordBuffers(agent.lib, agent.device, self.pad, group)
        elif transport == "nccl":
            dev = getattr(engine, "device", None) or engine.agent.device
            self._mine = torch.zeros((self.pad, 2), dtype=torch.int64, device=dev)
            self._all = torch.zeros((self.world, self.pad, 2), dtype=torch.int64, device=dev)
        elif transport != "auto":
            raise ValueError("transport must be auto, nccl or peer")

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t
        if t.numel() < self.pad:  # all_gather needs equal sizes: pad, then cut each rank's tail
            t = torch.cat([t, t.new_zeros(self.pad - t.numel())])
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return torch.cat([o[:n] for o, n in zip(out, self.sizes)])

    def step(self):
        if self.transport == "peer":
            p, slot = self.peers, self.t & 1
            p.check_timeout()                       # a barrier of an earlier step gave up: stop before anything else is applied
            self.engine.emit_records(p.records(self.rank, slot))
            p.barrier()
            self.engine.apply_records([p.records(r, slot) for r in range(self.world)], self.sizes)
        elif self.transport == "nccl":
            self.engine.emit_records(self._mine)
            if self.world > 1:
                dist.all_gather_into_tensor(self._all.view(-1), self._mine.view(-1), group=self.group)
                self.engine.apply_records([self._all[r] for r in range(self.world)], self.sizes)
            else:
                self.engine.apply_records([self._mine], self.sizes)
        else:
            keys, actions, targets = self.engine.emit()
            self.engine.apply(self._gather(keys), self._gather(actions), self._gather(targets))
        self.t += 1

    def close(self):
        if self.peers is not None:
            self.peers.check_timeout()
            self.peers.close()
            self.peers = None


class SharedQTable:
    """ONE Q-table for all GPUs of the box (the reference's single `q_table`, main.py:16, at box scale): shard j of
    the slot range lives in the HBM of rank j, every rank maps every shard (CUDA IPC) and runs the fused rollout on
    its own env shard; lookups and compare-and-swap updates of remote slots travel over NVLink 5 / NVSwitch inside
    the kernel (g2048_rollout_qlearn_sharded).  No exchange step, no replicas, table memory adds up over the GPUs.

    `shards` (list of int64 tensors on this device) builds a table from local allocations instead -- the
    single-process form used by the tests and by 1-GPU runs."""

    def __init__(self, lib, device: torch.device, slots_per_shard: int, group=None, shards=None):
        import ctypes
        if slots_per_shard & (slots_per_shard - 1):
            raise ValueError("slots_per_shard must be a power of two")
        self.lib, self.device, self.group, self.slots_per_shard = lib, device, group, int(slots_per_shard)
        self._opened, self._mine, self._local = [], None, None
        from ._lib import check
        self._check = check
        if shards is not None:
            self.rank, self.world = 0, 1
            self._tensors = list(shards)
            self.ptrs = [t.data_ptr() for t in self._tensors]
            self._local = self.ptrs            # all shards are local
        else:
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self._mine, self.ptrs, self._opened = _share_device_memory(lib, device, self.slots_per_shard * 32, group)
            self._local = [self._mine]
        m = len(self.ptrs)
        if m & (m - 1) or m * self.slots_per_shard > (1 << 31):
            raise ValueError("number of shards must be a power of two and the table at most 2^31 slots")
        self.n_shards = m
        self._arr = (ctypes.c_void_p * m)(*self.ptrs)
        self._count = torch.zeros(1, dtype=torch.int64, device=device)

    @property
    def capacity(self) -> int:
        return self.n_shards * self.slots_per_shard

    def rollout(self, env, k_steps: int, lr: float, gamma: float, eps: float) -> torch.Tensor:
        """k_steps of main.py:91-101 for every env of this rank on the shared table (asynchronous updates)."""
        with torch.cuda.device(self.device):
            env.counters.zero_()
            self._check(self.lib.g2048_rollout_qlearn_sharded(
                env.boards.data_ptr(), env.aux.data_ptr(), env.score.data_ptr(), self._arr, self.n_shards,
                self.slots_per_shard, env.n, k_steps, env.flavour, lr, gamma, float(eps), env.seed, env.step_idx,
                env.env_id_base, env.counters.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "g2048_rollout_qlearn_sharded")
        env.step_idx += k_steps
        return env.counters

    def lookup(self, boards: torch.Tensor, insert: bool = False):
        n = boards.numel()
        rows = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        found = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.g2048_qtable_lookup_sharded(self._arr, self.n_shards, self.slots_per_shard,
                                                             boards.data_ptr(), n, rows.data_ptr(), found.data_ptr(),
                                                             int(insert), torch.cuda.current_stream().cuda_stream),
                        "g2048_qtable_lookup_sharded")
        return rows, found != 0

    def local_size(self) -> int:
        """States stored in the shard(s) this rank owns."""
        total = 0
        with torch.cuda.device(self.device):
            for p in self._local:
                self._check(self.lib.g2048_qtable_size(p, self.slots_per_shard, self._count.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "g2048_qtable_size")
                total += int(self._count.item())
        return total

    def size(self) -> int:
        n = self.local_size()
        if self.world > 1:
            t = torch.tensor([n], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, group=self.group)
            n = int(t.item())
        return n

    def export_local(self):
        """(keys uint64[n], rows float32[n,4]) of this rank's shard(s), sorted by key."""
        import numpy as np
        ks, rs = [], []
        with torch.cuda.device(self.device):
            for p in self._local:
                self._check(self.lib.g2048_qtable_size(p, self.slots_per_shard, self._count.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "g2048_qtable_size")
                n = int(self._count.item())
                keys = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
                rows = torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.device)
                self._count.zero_()
                self._check(self.lib.g2048_qtable_export(p, self.slots_per_shard, keys.data_ptr(), rows.data_ptr(), n,
                                                         self._count.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream), "g2048_qtable_export")
                ks.append(keys[:n].cpu().numpy().view(np.uint64))
                rs.append(rows[:n].cpu().numpy())
        k, r = np.concatenate(ks), np.concatenate(rs)
        order = np.argsort(k)
        return k[order], r[order]

    def close(self):
        if self._mine or self._opened:
            _unshare_device_memory(self.lib, self.device, self._mine, self._opened, self.group)
        self._mine, self._opened = None, []


class OwnerComputesQLearning:
    """Exact synchronous Q-learning on ONE table sharded over the GPUs (`SharedQTable`), owner computes: per step every
    rank advances its envs against the shared table and appends each transition's record to the list of the GPU that
    owns the slot of its state; after a flag barrier every GPU sorts and applies only the records for ITS shard, which
    it reads in place from the other GPUs' memory.  Same result as the single-GPU deterministic step (bit for bit),
    but the sort/apply work per GPU is 1/G of the global batch instead of all of it (`ShardedQLearning`)."""

    HEAD = 512                                   # flags (256 B) + 2 x 16 counters

    def __init__(self, env, shared: SharedQTable, n_total: int, lr: float, gamma: float, eps: float, group=None,
                 window: int = 1):
        """window = K > 1: the table's values stay frozen for K env steps (new states are still inserted, as zero
        rows), the records of all K steps are then exchanged and applied at once in (step, global env) order -- the
        "exchange every K steps" form: one barrier pair, one sort and one host read per K steps."""
        import ctypes
        from ._lib import check
        self._check, self._ct = check, ctypes
        self.env, self.shared, self.group = env, shared, group
        self.lr, self.gamma, self.eps = lr, gamma, eps
        self.lib, self.device = shared.lib, shared.device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if shared.n_shards != self.world:
            raise ValueError("one shard per rank")
        self.n_total, self.window = int(n_total), int(window)
        self.idx_bits = max(1, (self.n_total * self.window - 1).bit_length())
        self.lo, _ = shard_range(self.n_total, self.rank, self.world)
        self.cap = max(hi - lo for lo, hi in (shard_range(self.n_total, r, self.world) for r in range(self.world)))
        self.list_bytes = ((self.cap * self.window * 16 + 255) // 256) * 256
        nbytes = self.HEAD + 2 * self.world * self.list_bytes
        self._mine, self.base, self._opened = _share_device_memory(self.lib, self.device, nbytes, group)
        with torch.cuda.device(self.device):
            self._flags = (ctypes.c_void_p * self.world)(*self.base)
            self.timed_out = torch.zeros(1, dtype=torch.int32).pin_memory()   # see PeerRecordBuffers
            self._scratch = None
            self._carry_slot = torch.zeros(env.n, dtype=torch.int32, device=self.device) if self.window > 1 else None
            self._carry_row = torch.zeros((env.n, 4), dtype=torch.float32, device=self.device) if self.window > 1 else None
        self.epoch, self.t, self.k = 0, 0, 0      # barrier epoch, window number, step inside the window
        # everything a step passes by pointer, built once per buffer slot (the step is short: per-step Python work --
        # ctypes arrays, pointer arithmetic -- would be a third of it)
        ct = ctypes
        self._args = []
        for slot in (0, 1):
            self._args.append({
                "lists": (ct.c_void_p * self.world)(*[self._list(self.rank, slot, j) for j in range(self.world)]),
                "counts": self._counts(self.rank, slot),
                "src": (ct.c_void_p * self.world)(*[self._counts(r, slot) + 8 * self.rank for r in range(self.world)]),
                "mine": (ct.c_void_p * self.world)(*[self._list(r, slot, self.rank) for r in range(self.world)]),
            })
        self._host = (ct.c_uint64 * self.world)()
        self._counts_i64 = (ct.c_int64 * self.world)()
        self._ptrs = (env.boards.data_ptr(), env.aux.data_ptr(), env.score.data_ptr(), env.counters.data_ptr(),
                      None if self._carry_slot is None else self._carry_slot.data_ptr(),
                      None if self._carry_row is None else self._carry_row.data_ptr(), self.timed_out.data_ptr())
        with torch.cuda.device(self.device):      # scratch for twice the even share of a window's records (grown if ever needed)
            need = int(self.lib.g2048_qlearn_scratch_bytes(2 * self.cap * self.window))
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)

    # layout helpers: counts[slot][j] and list[slot][j] inside rank r's buffer
    def _counts(self, r, slot):
        return self.base[r] + 256 + slot * 128

    def _list(self, r, slot, j):
        return self.base[r] + self.HEAD + (slot * self.world + j) * self.list_bytes

    def _barrier(self, st=None):
        self.epoch += 1
        rc = self.lib.g2048_peer_barrier(self._flags, self.rank, self.world, self.epoch, 0, self._ptrs[6],
                                         st if st is not None else torch.cuda.current_stream().cuda_stream)
        if rc:
            self._check(rc, "g2048_peer_barrier")

    def step(self):
        """One env step of every env of this rank; returns the number of records this rank applied to its shard (0
        inside a window)."""
        env, sh, slot = self.env, self.shared, self.t & 1
        if env.boards.data_ptr() != self._ptrs[0]:           # the env's buffers were replaced: take the new addresses
            self._ptrs = (env.boards.data_ptr(), env.aux.data_ptr(), env.score.data_ptr(), env.counters.data_ptr()) + self._ptrs[4:]
        args, p = self._args[slot], self._ptrs
        st = torch.cuda.current_stream().cuda_stream
        total = 0
        self._check_timeout()
        with torch.cuda.device(self.device):
            if self.k == 0:
                rc = self.lib.g2048_peer_memset(args["counts"], 0, 128, st)
                if rc:
                    self._check(rc, "g2048_peer_memset")
            rc = self.lib.g2048_qlearn_emit_owned(
                p[0], p[1], p[2], sh._arr, sh.n_shards, sh.slots_per_shard, env.n, env.flavour, self.gamma, float(self.eps),
                env.seed, env.step_idx, env.env_id_base, self.k * self.n_total + self.lo, self.idx_bits, p[3],
                args["lists"], args["counts"], p[4], p[5], int(self.k > 0), st)
            if rc:
                self._check(rc, "g2048_qlearn_emit_owned")
            env.step_idx += 1
            self.k += 1
            if self.k == self.window:
                total = self._exchange_and_apply(slot, st)
                self.k = 0
                self.t += 1
        return total

    def flush(self):
        """Apply the records of a partly filled window now (every rank must call it at the same step)."""
        total = 0
        if self.k:
            with torch.cuda.device(self.device):
                total = self._exchange_and_apply(self.t & 1, torch.cuda.current_stream().cuda_stream)
            self.k = 0
            self.t += 1
        return total

    def _exchange_and_apply(self, slot, st):
        sh, args, host, counts = self.shared, self._args[slot], self._host, self._counts_i64
        self._barrier(st)                                   # every rank's records and counts are written
        rc = self.lib.g2048_peer_read_u64(args["src"], self.world, host, st)
        if rc:
            self._check(rc, "g2048_peer_read_u64")
        self._check_timeout()                               # (the read synchronised the stream: the barrier is over)
        total = 0
        for r in range(self.world):
            counts[r] = host[r]
            total += host[r]
        if total > 2 * self.cap * self.window:              # a very uneven split: more scratch
            need = int(self.lib.g2048_qlearn_scratch_bytes(total))
            if self._scratch.numel() < need:
                self._scratch = torch.empty(int(need * 1.25), dtype=torch.uint8, device=self.device)
        rc = self.lib.g2048_qtable_apply_owned(sh.ptrs[self.rank], sh.slots_per_shard, args["mine"], counts, self.world,
                                               self.idx_bits, self.lr, self._scratch.data_ptr(), self._scratch.numel(), st)
        if rc:
            self._check(rc, "g2048_qtable_apply_owned")
        self._barrier(st)                                   # all shards updated before anyone reads them again
        return total

    def _check_timeout(self):
        v = int(self.timed_out[0])
        if v:
            raise RuntimeError(f"peer barrier timed out waiting for rank {v - 1}: the ranks are no longer in step")

    def close(self):
        try:
            self.flush()
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
            self._check_timeout()
        finally:
            _unshare_device_memory(self.lib, self.device, self._mine, self._opened, self.group)
            self._mine, self._opened = None, []


class RoutedQLearning:
    """Exact synchronous Q-learning on ONE table sharded over the GPUs, ROUTED (`g2048_routed_*`, include/g2048.h): same
    result as `OwnerComputesQLearning` with window 1 and as the single-GPU deterministic step, but no GPU ever touches
    another GPU's shard -- the envs' lookups travel to the owner of the state as bulk lists (keys out, {slot, max Q} and
    rows back), the records likewise, all of it coalesced NVLink traffic (48 B per env step) instead of ~2.4 small
    remote requests per env.  Four flag barriers and one stream synchronisation per step, the whole step is one C call."""

    def __init__(self, env, shared: SharedQTable, n_total: int, lr: float, gamma: float, eps: float, group=None):
        import ctypes
        from ._lib import check
        self._check, self._ct = check, ctypes
        self.env, self.shared, self.group = env, shared, group
        self.lr, self.gamma, self.eps = lr, gamma, eps
        self.lib, self.device = shared.lib, shared.device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if shared.n_shards != self.world:
            raise ValueError("one shard per rank")
        self.n_total = int(n_total)
        self.cap = max(hi - lo for lo, hi in (shard_range(self.n_total, r, self.world) for r in range(self.world)))
        if env.n > self.cap:
            raise ValueError("this rank holds more envs than its share of n_total")
        nbytes = int(self.lib.g2048_routed_buffer_bytes(self.world, self.cap))
        self._mine, self.base, self._opened = _share_device_memory(self.lib, self.device, nbytes, group)
        with torch.cuda.device(self.device):
            bufs = (ctypes.c_void_p * self.world)(*self.base)
            self._h = self.lib.g2048_routed_create(self.rank, self.world, self.cap, self.n_total, bufs,
                                                   shared.ptrs[self.rank], shared.slots_per_shard)
            if not self._h:
                self._check(-1, "g2048_routed_create")
            self._applied = ctypes.c_int64(0)
            self.prime()

    def prime(self):
        """Look the envs' current boards up (after a reset, or when the boards were changed from outside); collective."""
        with torch.cuda.device(self.device):
            self._check(self.lib.g2048_routed_prime(self._h, self.env.boards.data_ptr(), self.env.n,
                                                    torch.cuda.current_stream().cuda_stream), "g2048_routed_prime")

    def step(self) -> int:
        """One env step of every env of this rank (collective); returns the records applied to this rank's shard."""
        env = self.env
        with torch.cuda.device(self.device):
            rc = self.lib.g2048_routed_step(self._h, env.boards.data_ptr(), env.aux.data_ptr(), env.score.data_ptr(), env.n,
                                            env.flavour, self.lr, self.gamma, float(self.eps), env.seed, env.step_idx,
                                            env.env_id_base, env.counters.data_ptr(), self._ct.byref(self._applied),
                                            torch.cuda.current_stream().cuda_stream)
            if rc:
                self._check(rc, "g2048_routed_step")
        env.step_idx += 1
        return self._applied.value

    def close(self):
        try:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
        finally:
            if self._h:
                self.lib.g2048_routed_destroy(self._h)
                self._h = None
            _unshare_device_memory(self.lib, self.device, self._mine, self._opened, self.group)
            self._mine, self._opened = None, []


class GradientAllReduce:
    """Data-parallel DQN (BASELINE config 5): every parameter's .grad is a view into ONE flat buffer,
    so a training step costs a single all-reduce over NVSwitch (no per-tensor launches, no bucket copies), followed
    by the division by the world size.  `sync_parameters()` broadcasts rank 0's weights once at start."""

    def __init__(self, model: torch.nn.Module, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        p0 = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def sync_parameters(self, src: int = 0):
        if self.world > 1:
            for p in self.params:
                dist.broadcast(p.data, src, group=self.group)

    def zero_grad(self):
        self.flat.zero_()

    def __call__(self):
        """Average the gradients of all ranks in place (call between backward() and optimizer.step())."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)
