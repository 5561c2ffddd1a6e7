"""The cross-GPU record exchange of the synchronous step: packed 16-byte records, the apply kernel that reads
record lists in place (local or NVLink peer memory), CUDA-IPC peer buffers and the flag barrier.

One GPU is enough for all of it: the two-process test maps each process's buffer into the other through CUDA IPC
(both on cuda:0), which is the same code path two GPUs take -- only the wire differs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, SEED, CAP = 3001, 10, 123, 1 << 18


def np_boards(t):
    return t.detach().cpu().numpy().view(np.uint64)


def single_process_result(g, n, steps, flavour="penalty", cap=CAP):
    env = g.BatchedGame2048Env(n, flavour, seed=SEED)
    agent = g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=cap, seed=SEED)
    env.reset()
    for _ in range(steps):
        agent.step_sync(env, mode="deterministic")
    keys, rows = agent.export()
    nz = np.abs(rows).sum(1) > 0
    return np_boards(env.boards).copy(), keys[nz], rows[nz]


def test_packed_records_equal_the_three_array_records():
    import torch
    import g2048
    n = 4096
    envs = [g2048.BatchedGame2048Env(n, "penalty", seed=SEED) for _ in range(2)]
    agents = [g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=CAP, seed=SEED) for _ in range(2)]
    for e in envs:
        e.reset()
    for t in range(6):
        k, a, tg = agents[0].step_sync(envs[0], mode="deterministic", apply=False, records=True)
        rec = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
        agents[1].emit_records(envs[1], rec)
        r = rec.cpu().numpy().view(np.uint64)
        assert np.array_equal(r[:, 0], np_boards(k))
        assert np.array_equal((r[:, 1] & np.uint64(3)).astype(np.uint8), a.cpu().numpy())
        assert np.array_equal((r[:, 1] >> np.uint64(32)).astype(np.uint32).view(np.float32), tg.cpu().numpy())
        assert torch.equal(envs[0].boards, envs[1].boards) and torch.equal(envs[0].aux, envs[1].aux)
        agents[0].apply_targets(k, a, tg)
        agents[1].apply_records([rec], [n])
    (k0, r0), (k1, r1) = agents[0].export(), agents[1].export()
    assert np.array_equal(k0, k1) and np.array_equal(r0, r1)


def test_record_lists_applied_in_place_equal_the_single_shard_step():
    """Ragged virtual shards (sizes 1001/0/1500/500): lists are read where they lie, in order, empty lists allowed."""
    import torch
    import g2048
    sizes = [1001, 0, 1500, 500]
    n = sum(sizes)
    boards1, keys1, rows1 = single_process_result(g2048, n, STEPS)
    lo = np.concatenate([[0], np.cumsum(sizes)])
    envs, recs = [], []
    for r, m in enumerate(sizes):
        e = g2048.BatchedGame2048Env(max(m, 1), "penalty", seed=SEED, env_id_base=int(lo[r]))
        e.reset()
        envs.append(e)
        recs.append(torch.zeros((max(m, 1), 2), dtype=torch.int64, device="cuda"))
    replica = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=CAP, seed=SEED)
    for t in range(STEPS):
        for r, m in enumerate(sizes):
            if m:
                replica.emit_records(envs[r], recs[r])
        replica.apply_records(recs, sizes)
    got = np.concatenate([np_boards(envs[r].boards)[:m] for r, m in enumerate(sizes)])
    assert np.array_equal(got, boards1)
    k, rows = replica.export()
    nz = np.abs(rows).sum(1) > 0
    assert np.array_equal(k[nz], keys1) and np.array_equal(rows[nz], rows1)


def test_apply_records_rejects_bad_arguments():
    import ctypes
    import torch
    import g2048
    L = g2048.lib()
    agent = g2048.BatchedQLearningAgent(10, capacity=1 << 10)
    rec = torch.zeros((8, 2), dtype=torch.int64, device="cuda")
    ptrs = (ctypes.c_void_p * 1)(rec.data_ptr() + 8)          # misaligned list
    cnt = (ctypes.c_int64 * 1)(4)
    assert L.g2048_qtable_apply_records(agent.table.data_ptr(), 1 << 10, ptrs, cnt, 1, 0.1, 1, None, 0, None) == -1
    ptrs = (ctypes.c_void_p * 1)(rec.data_ptr())
    assert L.g2048_qtable_apply_records(agent.table.data_ptr(), 1 << 10, ptrs, cnt, 17, 0.1, 1, None, 0, None) == -1
    assert L.g2048_qtable_apply_records(agent.table.data_ptr(), 1 << 10, ptrs, cnt, 1, 0.1, 1, None, 0, None) == -3
    assert b"scratch" in L.g2048_last_error()
    flags = (ctypes.c_void_p * 1)(0)
    assert L.g2048_peer_barrier(flags, 0, 1, 1, 0, None, None) == -1


def _worker(rank, world, port, out, transport):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import g2048
    from g2048 import dist as gdist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(N_TOTAL, rank, world)
    env = g2048.BatchedGame2048Env(hi - lo, "penalty", seed=SEED, env_id_base=lo)
    agent = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=CAP, seed=SEED)
    env.reset()
    sh = gdist.ShardedQLearning(gdist.TorchEngine(env, agent), N_TOTAL, transport=transport)
    for _ in range(STEPS):
        sh.step()
    torch.cuda.synchronize()
    sh.peers.check_timeout()
    keys, rows = agent.export()
    nz = np.abs(rows).sum(1) > 0
    np.savez(os.path.join(out, f"rank{rank}.npz"), boards=np_boards(env.boards), keys=keys[nz], rows=rows[nz], lo=lo, hi=hi,
             epoch=sh.peers.epoch)
    sh.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_processes_exchange_through_ipc_peer_memory(tmp_path):
    """Two processes (both on cuda:0) map each other's record buffers with CUDA IPC, synchronise with the flag
    barrier kernel and read each other's records in place: boards and table equal the single-process run."""
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), "peer"), nprocs=2, join=True)
    boards1, keys1, rows1 = single_process_result(g2048, N_TOTAL, STEPS)
    for rank in range(2):
        d = np.load(tmp_path / f"rank{rank}.npz")
        assert int(d["epoch"]) == STEPS
        assert np.array_equal(d["boards"], boards1[int(d["lo"]):int(d["hi"])])
        assert np.array_equal(d["keys"], keys1) and np.array_equal(d["rows"], rows1)


# ---------------------------------------------------------------- one Q-table spread over several GPUs' memory
def test_sharded_table_with_one_env_is_the_local_table_bit_for_bit():
    """N = 1 is sequential, hence deterministic: 4 shards of 2^12 slots hold exactly the bytes of one 2^14 table."""
    import torch
    import g2048
    from g2048 import dist as gdist
    L = g2048.lib()
    for flavour in ("penalty", "nopenalty"):
        env_a = g2048.BatchedGame2048Env(1, flavour, seed=SEED)
        env_b = g2048.BatchedGame2048Env(1, flavour, seed=SEED)
        agent = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.3, capacity=1 << 14, seed=SEED)
        shards = [torch.zeros((1 << 12) * 4, dtype=torch.int64, device="cuda") for _ in range(4)]
        shared = gdist.SharedQTable(L, torch.device("cuda", 0), 1 << 12, shards=shards)
        env_a.reset(); env_b.reset()
        for _ in range(5):
            ca = agent.rollout(env_a, 200).clone()
            cb = shared.rollout(env_b, 200, 0.1, 0.99, 0.3).clone()
            assert torch.equal(ca, cb)
        assert torch.equal(env_a.boards, env_b.boards) and torch.equal(env_a.aux, env_b.aux)
        assert torch.equal(torch.cat(shards), agent.table)
        assert shared.size() == len(agent) and shared.capacity == 1 << 14


def test_sharded_table_many_envs_random_policy():
    """epsilon = 1: trajectories do not depend on the table, so counters and the SET of stored states must equal the
    local-table run exactly (values may differ by update order); every stored state is found by the sharded lookup."""
    import torch
    import g2048
    from g2048 import dist as gdist
    L = g2048.lib()
    n = 50_000
    env_a = g2048.BatchedGame2048Env(n, "penalty", seed=SEED)
    env_b = g2048.BatchedGame2048Env(n, "penalty", seed=SEED)
    agent = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 1.0, capacity=1 << 23, seed=SEED)
    shards = [torch.zeros((1 << 22) * 4, dtype=torch.int64, device="cuda") for _ in range(2)]
    shared = gdist.SharedQTable(L, torch.device("cuda", 0), 1 << 22, shards=shards)
    env_a.reset(); env_b.reset()
    ca = agent.rollout(env_a, 48).clone()
    cb = shared.rollout(env_b, 48, 0.1, 0.99, 1.0).clone()
    assert torch.equal(ca[:8], cb[:8])                      # everything but the lost-update count
    assert torch.equal(env_a.boards, env_b.boards)
    k_local, r_local = agent.export()
    k_shared, r_shared = shared.export_local()
    assert np.array_equal(k_local, k_shared)
    assert np.isfinite(r_shared).all() and np.abs(r_shared).max() <= 10.0 / (1 - 0.99) + 1e-3
    assert abs(float(np.abs(r_shared).mean()) / float(np.abs(r_local).mean()) - 1) < 0.10
    keys = torch.from_numpy(k_shared.view(np.int64)).cuda()
    rows, found = shared.lookup(keys)
    assert bool(found.all()) and np.array_equal(rows.cpu().numpy(), r_shared)
    # both halves of the slot range are in use
    assert all(int((s.view(-1, 4)[:, 0] != 0).sum()) > 0.4 * len(k_shared) for s in shards)


def test_sharded_rollout_rejects_bad_shard_lists():
    import ctypes
    import torch
    import g2048
    L = g2048.lib()
    t = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
    b = torch.zeros(8, dtype=torch.int64, device="cuda")
    three = (ctypes.c_void_p * 3)(t.data_ptr(), t.data_ptr(), t.data_ptr())
    args = (8, 4, 0, 0.1, 0.9, 0.1, 1, 0, 0, None, None)
    assert L.g2048_rollout_qlearn_sharded(b.data_ptr(), None, None, three, 3, 1024, *args) == -1     # not 2^k shards
    two = (ctypes.c_void_p * 2)(t.data_ptr(), 0)
    assert L.g2048_rollout_qlearn_sharded(b.data_ptr(), None, None, two, 2, 1024, *args) == -1        # null shard
    one = (ctypes.c_void_p * 1)(t.data_ptr())
    assert L.g2048_rollout_qlearn_sharded(b.data_ptr(), None, None, one, 1, 1000, *args) == -1        # not 2^k slots
    assert L.g2048_rollout_qlearn_sharded(b.data_ptr(), None, None, one, 1, 1 << 32, *args) == -1     # > 2^31 slots


def _shared_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import g2048
    from g2048 import dist as gdist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(20_000, rank, world)
    env = g2048.BatchedGame2048Env(hi - lo, "penalty", seed=SEED, env_id_base=lo)
    env.reset()
    shared = gdist.SharedQTable(g2048.lib(), torch.device("cuda", 0), 1 << 21)
    for _ in range(3):
        shared.rollout(env, 16, 0.1, 0.99, 1.0)
    torch.cuda.synchronize()
    dist.barrier()
    total = shared.size()
    keys, rows = shared.export_local()
    np.savez(os.path.join(out, f"shared{rank}.npz"), keys=keys, rows=rows, total=total, boards=np_boards(env.boards))
    shared.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_processes_learn_one_table_through_peer_memory(tmp_path):
    """Two processes each own half of the slot range and run the fused rollout on their env shard at the same time;
    the union of the two shards holds exactly the states a single process visits with all the envs."""
    import torch
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_shared_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    env = g2048.BatchedGame2048Env(20_000, "penalty", seed=SEED)
    agent = g2048.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 1.0, capacity=1 << 22, seed=SEED)
    env.reset()
    for _ in range(3):
        agent.rollout(env, 16)
    k1, r1 = agent.export()
    d = [np.load(tmp_path / f"shared{r}.npz") for r in range(2)]
    union = np.sort(np.concatenate([d[0]["keys"], d[1]["keys"]]))
    assert np.array_equal(union, k1)
    assert int(d[0]["total"]) == int(d[1]["total"]) == len(k1)
    assert min(len(d[0]["keys"]), len(d[1]["keys"])) > 0.4 * len(k1)
    assert np.array_equal(np.concatenate([d[0]["boards"], d[1]["boards"]]), np_boards(env.boards))


# ---------------------------------------------------------------- exact synchronous step, owner computes
@pytest.mark.parametrize("flavour,code", [("penalty", 0), ("nopenalty", 1)])
def test_owner_computes_step_equals_the_single_table_deterministic_step(flavour, code):
    """Virtual ranks in one process through the raw C ABI: 2 env shards (ragged), a table of 2 shards; every owner sorts
    and applies only the records for its shard.  Boards and table content equal the single-GPU deterministic run."""
    import ctypes
    import torch
    import g2048
    from g2048 import dist as gdist
    L = g2048.lib()
    sizes, G, steps, slots = [1700, 1301], 2, 14, 1 << 17
    n = sum(sizes)
    boards1, keys1, rows1 = single_process_result(g2048, n, steps, flavour)       # capacity CAP = 2 * slots
    assert CAP == G * slots
    lo = [0, sizes[0]]
    idx_bits = (n - 1).bit_length()
    shards = [torch.zeros(slots * 4, dtype=torch.int64, device="cuda") for _ in range(G)]
    shared = gdist.SharedQTable(L, torch.device("cuda", 0), slots, shards=shards)
    envs = [g2048.BatchedGame2048Env(sizes[r], flavour, seed=SEED, env_id_base=lo[r]) for r in range(G)]
    for e in envs:
        e.reset()
    lists = [[torch.zeros((max(sizes), 2), dtype=torch.int64, device="cuda") for _ in range(G)] for _ in range(G)]
    counts = [torch.zeros(G, dtype=torch.int64, device="cuda") for _ in range(G)]
    scratch = torch.empty(L.g2048_qlearn_scratch_bytes(n), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    applied = 0
    for t in range(steps):
        for r in range(G):
            counts[r].zero_()
            arr = (ctypes.c_void_p * G)(*[x.data_ptr() for x in lists[r]])
            rc = L.g2048_qlearn_emit_owned(envs[r].boards.data_ptr(), envs[r].aux.data_ptr(), envs[r].score.data_ptr(),
                                           shared._arr, G, slots, sizes[r], code, 0.99, 0.4, SEED, t, lo[r], lo[r], idx_bits,
                                           envs[r].counters.data_ptr(), arr, counts[r].data_ptr(), None, None, 0, st)
            assert rc == 0, L.g2048_last_error()
        host = torch.stack(counts).cpu().numpy()                 # [rank][owner]
        assert host.sum() == n
        for j in range(G):
            arr = (ctypes.c_void_p * G)(*[lists[r][j].data_ptr() for r in range(G)])
            cnt = (ctypes.c_int64 * G)(*[int(host[r][j]) for r in range(G)])
            rc = L.g2048_qtable_apply_owned(shards[j].data_ptr(), slots, arr, cnt, G, idx_bits, 0.1, scratch.data_ptr(),
                                            scratch.numel(), st)
            assert rc == 0, L.g2048_last_error()
            applied += int(host[:, j].sum())
    assert applied == n * steps
    got = np.concatenate([np_boards(e.boards) for e in envs])
    assert np.array_equal(got, boards1)
    k, rows = shared.export_local()
    nz = np.abs(rows).sum(1) > 0
    assert np.array_equal(k[nz], keys1) and np.array_equal(rows[nz], rows1)
    # both owners had work
    assert min(int((s.view(-1, 4)[:, 0] != 0).sum()) for s in shards) > 1000
    # misuse: record index that does not fit idx_bits
    arr = (ctypes.c_void_p * G)(*[x.data_ptr() for x in lists[0]])
    assert L.g2048_qlearn_emit_owned(envs[0].boards.data_ptr(), None, None, shared._arr, G, slots, sizes[0], 0, 0.99, 0.4,
                                     SEED, 0, 0, 1 << 20, 8, None, arr, counts[0].data_ptr(), None, None, 0, st) == -1


def windowed_single_process_result(g, n, steps, window):
    """Reference for the K-step window: K steps on frozen values (records only), then all records applied in
    (step, env) order -- on ONE table in one process."""
    import torch
    env = g.BatchedGame2048Env(n, "penalty", seed=SEED)
    agent = g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=CAP, seed=SEED)
    env.reset()
    for w in range(0, steps, window):
        recs = [agent.step_sync(env, mode="deterministic", apply=False, records=True) for _ in range(min(window, steps - w))]
        agent.apply_targets(torch.cat([r[0] for r in recs]), torch.cat([r[1] for r in recs]), torch.cat([r[2] for r in recs]))
    keys, rows = agent.export()
    nz = np.abs(rows).sum(1) > 0
    return np_boards(env.boards).copy(), keys[nz], rows[nz]


def _owner_worker(rank, world, port, out, window=1):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import g2048
    from g2048 import dist as gdist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(N_TOTAL, rank, world)
    env = g2048.BatchedGame2048Env(hi - lo, "penalty", seed=SEED, env_id_base=lo)
    env.reset()
    shared = gdist.SharedQTable(g2048.lib(), torch.device("cuda", 0), CAP // world)
    oc = gdist.OwnerComputesQLearning(env, shared, N_TOTAL, 0.1, 0.99, 0.4, window=window)
    steps = STEPS if window == 1 else 14
    handled = [oc.step() for _ in range(steps)]
    handled.append(oc.flush())                       # the partly filled last window (a no-op for window = 1)
    torch.cuda.synchronize()
    dist.barrier()
    keys, rows = shared.export_local()
    nz = np.abs(rows).sum(1) > 0
    np.savez(os.path.join(out, f"owner{rank}.npz"), boards=np_boards(env.boards), keys=keys[nz], rows=rows[nz], lo=lo, hi=hi,
             handled=np.array(handled))
    oc.close()
    shared.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_processes_owner_computes_through_ipc_peer_memory(tmp_path):
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_owner_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    boards1, keys1, rows1 = single_process_result(g2048, N_TOTAL, STEPS)
    d = [np.load(tmp_path / f"owner{r}.npz") for r in range(2)]
    for x in d:
        assert np.array_equal(x["boards"], boards1[int(x["lo"]):int(x["hi"])])
    assert np.array_equal(d[0]["handled"] + d[1]["handled"], np.array([N_TOTAL] * STEPS + [0]))
    keys = np.concatenate([d[0]["keys"], d[1]["keys"]])
    rows = np.concatenate([d[0]["rows"], d[1]["rows"]])
    order = np.argsort(keys)
    assert np.array_equal(keys[order], keys1) and np.array_equal(rows[order], rows1)


@pytest.mark.timeout(600)
def test_two_processes_owner_computes_with_a_4_step_window(tmp_path):
    """Exchange every K = 4 steps: values frozen inside the window, all 4 x N records applied at once in (step, env)
    order by their owners -- equal to the single-table, single-process run of the same rule."""
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_owner_worker, args=(2, port, str(tmp_path), 4), nprocs=2, join=True)
    boards1, keys1, rows1 = windowed_single_process_result(g2048, N_TOTAL, 14, 4)
    d = [np.load(tmp_path / f"owner{r}.npz") for r in range(2)]
    for x in d:
        assert np.array_equal(x["boards"], boards1[int(x["lo"]):int(x["hi"])])
    handled = d[0]["handled"] + d[1]["handled"]
    assert handled.tolist() == [0, 0, 0, 4 * N_TOTAL] * 3 + [0, 0, 2 * N_TOTAL]
    keys = np.concatenate([d[0]["keys"], d[1]["keys"]])
    rows = np.concatenate([d[0]["rows"], d[1]["rows"]])
    order = np.argsort(keys)
    assert np.array_equal(keys[order], keys1) and np.array_equal(rows[order], rows1)


# ---------------------------------------------------------------- exact synchronous step, routed (every table access local)
def _routed_worker(rank, world, port, out, flavour, n_total=N_TOTAL, steps=None, cap=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import g2048
    from g2048 import dist as gdist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(n_total, rank, world)
    env = g2048.BatchedGame2048Env(hi - lo, flavour, seed=SEED, env_id_base=lo)
    env.reset()
    shared = gdist.SharedQTable(g2048.lib(), torch.device("cuda", 0), (cap or CAP_ROUTED) // world)
    rq = gdist.RoutedQLearning(env, shared, n_total, 0.1, 0.99, 0.4)
    handled = [rq.step() for _ in range(steps or STEPS_ROUTED)]
    torch.cuda.synchronize()
    dist.barrier()
    keys, rows = shared.export_local()
    nz = np.abs(rows).sum(1) > 0
    np.savez(os.path.join(out, f"routed{rank}.npz"), boards=np_boards(env.boards), keys=keys[nz], rows=rows[nz], lo=lo, hi=hi,
             handled=np.array(handled), counters=env.counters.cpu().numpy(), all_keys=keys)
    rq.close()
    shared.close()
    dist.destroy_process_group()


STEPS_ROUTED = 160         # long enough for games to end: the request for the fresh board after a game over is exercised
CAP_ROUTED = 1 << 21       # (about 450,000 states)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("flavour", ["penalty", "nopenalty"])
def test_two_processes_routed_step_equals_the_single_table_deterministic_step(tmp_path, flavour):
    """g2048_routed_*: keys travel to the owner of their home slot, {slot, max Q} and rows come back, records go to the
    owner of s; no process touches the other's shard.  Boards and every non-zero row equal the single-process
    deterministic step on one table, bit for bit; every state a shard holds is owned by it."""
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_routed_worker, args=(2, port, str(tmp_path), flavour), nprocs=2, join=True)
    boards1, keys1, rows1 = single_process_result(g2048, N_TOTAL, STEPS_ROUTED, flavour, CAP_ROUTED)
    d = [np.load(tmp_path / f"routed{r}.npz") for r in range(2)]
    for x in d:
        assert np.array_equal(x["boards"], boards1[int(x["lo"]):int(x["hi"])])
    assert np.array_equal(d[0]["handled"] + d[1]["handled"], np.array([N_TOTAL] * STEPS_ROUTED))
    assert min(d[0]["handled"].min(), d[1]["handled"].min()) > 0.3 * N_TOTAL        # both owners had work
    keys = np.concatenate([d[0]["keys"], d[1]["keys"]])
    rows = np.concatenate([d[0]["rows"], d[1]["rows"]])
    order = np.argsort(keys)
    assert np.array_equal(keys[order], keys1) and np.array_equal(rows[order], rows1)
    assert len(np.intersect1d(d[0]["all_keys"], d[1]["all_keys"])) == 0              # no state lives in both shards
    episodes = int(d[0]["counters"][2] + d[1]["counters"][2])
    assert episodes > 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n_total,steps,cap,flavour", [(80001, 5, CAP_ROUTED, "penalty"), ((1 << 20) + 77, 3, 1 << 23, "penalty"),
                                                       ((1 << 20) + 77, 3, 1 << 23, "nopenalty")])
def test_routed_step_with_many_envs_per_rank(tmp_path, n_total, steps, cap, flavour):
    """40,000 envs per rank: the places of the records come from more than one tile of k_routed_scan.  524,000 envs per
    rank: k_routed_request runs with the row LUT staged in shared memory (the path BASELINE-size runs take)."""
    import torch.multiprocessing as mp
    import g2048
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_routed_worker, args=(2, port, str(tmp_path), flavour, n_total, steps, cap), nprocs=2, join=True)
    boards1, keys1, rows1 = single_process_result(g2048, n_total, steps, flavour, cap)
    d = [np.load(tmp_path / f"routed{r}.npz") for r in range(2)]
    for x in d:
        assert np.array_equal(x["boards"], boards1[int(x["lo"]):int(x["hi"])])
    assert np.array_equal(d[0]["handled"] + d[1]["handled"], np.array([n_total] * steps))
    keys = np.concatenate([d[0]["keys"], d[1]["keys"]])
    rows = np.concatenate([d[0]["rows"], d[1]["rows"]])
    order = np.argsort(keys)
    assert np.array_equal(keys[order], keys1) and np.array_equal(rows[order], rows1)
