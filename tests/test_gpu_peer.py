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


def single_process_result(g, n, steps):
    env = g.BatchedGame2048Env(n, "penalty", seed=SEED)
    agent = g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.4, capacity=CAP, seed=SEED)
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
