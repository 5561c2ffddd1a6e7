"""The N > 1 path on CPU: two gloo processes run the sharded synchronous Q-learning exchange
(dist.ShardedQLearning) with an oracle-backed engine and must reproduce the single-process table and
boards exactly -- sharding invariance through global env ids and rank-ordered record gathering."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, SEED = 1001, 12, 77   # odd size: unequal shards exercise the padding


class OracleEngine:
    def __init__(self, lo, hi):
        import oracle
        self.o, self.lo = oracle, lo
        n = hi - lo
        self.boards = np.zeros(n, np.uint64)
        oracle.env_reset(self.boards, None, None, None, seed=SEED, episode_idx=0, env_id_base=lo)
        self.aux = np.full(n, oracle.AUX_INIT, np.uint64)
        self.score = np.zeros(n, np.int32)
        self.tab = oracle.QTable(1 << 16, f32=True)
        self.t = 0

    def emit(self):
        _, (k, a, d) = self.o.qlearn_step_sync(self.boards, self.aux, self.score, self.tab, 0.1, 0.99, 0.4, 0, SEED,
                                               self.t, self.lo, records=True, apply=False)
        self.t += 1
        return torch.from_numpy(k.view(np.int64)), torch.from_numpy(a), torch.from_numpy(d)

    def apply(self, keys, actions, deltas):
        self.tab.apply_targets_f32(keys.numpy().view(np.uint64).copy(), actions.numpy().copy(), deltas.numpy().copy(), 0.1)

    # the packed 16-byte record protocol of the GPU engine (g2048_record: key | action + float32 target bits << 32)
    device = torch.device("cpu")

    def emit_records(self, records):
        k, a, d = self.emit()
        n = k.numel()
        at = a.numpy().astype(np.uint64) | (d.numpy().view(np.uint32).astype(np.uint64) << np.uint64(32))
        records[:n, 0] = k
        records[:n, 1] = torch.from_numpy(at.view(np.int64))

    def apply_records(self, lists, counts):
        rec = np.concatenate([l[:c].numpy().view(np.uint64) for l, c in zip(lists, counts)])
        self.tab.apply_targets_f32(rec[:, 0].copy(), (rec[:, 1] & np.uint64(3)).astype(np.uint8),
                                   (rec[:, 1] >> np.uint64(32)).astype(np.uint32).view(np.float32).copy(), 0.1)


def _worker(rank, world, port, out, transport="auto"):
    sys.path.insert(0, ROOT)
    import g2048
    from g2048 import dist as gdist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(N_TOTAL, rank, world)
    eng = OracleEngine(lo, hi)
    sh = gdist.ShardedQLearning(eng, N_TOTAL, transport=transport)
    for _ in range(STEPS):
        sh.step()
    keys, rows = eng.tab.export()
    nz = np.abs(rows).sum(1) > 0
    np.savez(os.path.join(out, f"rank{rank}.npz"), boards=eng.boards, keys=keys[nz], rows=rows[nz], lo=lo, hi=hi)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from g2048 import dist as gdist
    for n, w in ((10, 3), (8, 8), (1 << 23, 8), (5, 8)):
        r = [gdist.shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


@pytest.mark.timeout(300)
@pytest.mark.parametrize("transport", ["auto", "nccl"])     # "nccl" = one gather of packed records (gloo here)
def test_two_rank_exchange_equals_single_process(tmp_path, transport):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), transport), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    single = OracleEngine(0, N_TOTAL)
    for t in range(STEPS):
        single.o.qlearn_step_sync(single.boards, single.aux, single.score, single.tab, 0.1, 0.99, 0.4, 0, SEED, t, 0)
    keys, rows = single.tab.export()
    nz = np.abs(rows).sum(1) > 0
    for rank in range(2):
        d = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(d["boards"], single.boards[int(d["lo"]):int(d["hi"])])
        assert np.array_equal(d["keys"], keys[nz])
        assert np.array_equal(d["rows"], rows[nz])     # float32 sums in the same (global env id) order: bit-identical


def _grad_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import oracle
    from g2048 import dist as gdist
    from g2048 import dqn
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different initial weights: sync_parameters must fix that
    model = dqn.DQNModel(width=8, hidden=8).eval()      # eval: no dropout noise in the comparison
    sync = gdist.GradientAllReduce(model)
    sync.sync_parameters()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    rng = np.random.RandomState(rank)
    boards = rng.randint(1, 1 << 62, size=32).astype(np.uint64)
    x = torch.from_numpy(oracle.encode_onehot(boards))
    y = torch.from_numpy(rng.standard_normal((32, 4)).astype(np.float32))
    local = []
    for _ in range(3):
        sync.zero_grad()
        loss = torch.mean((model(x) - y) ** 2)
        loss.backward()
        local.append(sync.flat.clone())
        sync()
        opt.step()
    torch.save({"local": local, "avg": sync.flat.clone(), "params": [p.detach().clone() for p in model.parameters()],
                "views": all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())},
               os.path.join(out, f"grad{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dqn_gradient_allreduce_keeps_two_replicas_identical(tmp_path):
    """Data-parallel DQN plumbing: one flat gradient buffer, one all-reduce, identical replicas."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"grad{r}.pt") for r in range(2))
    assert a["views"] and b["views"]
    assert torch.allclose(a["avg"], (a["local"][-1] + b["local"][-1]) / 2, atol=1e-7)
    assert torch.equal(a["avg"], b["avg"])
    assert not torch.equal(a["local"][0], b["local"][0])          # the ranks really saw different data
    for p, q in zip(a["params"], b["params"]):
        assert torch.equal(p, q)


# ---------------------------------------------------------------- the routed exact step (dist.RoutedQLearning) as a protocol
def _mix64(x):
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def _owner(keys, world, slot_bits=16):
    """owner(key) of g2048_routed_*: the top bits of the global home slot (csrc/g2048.cu, RoutedLocal.owner_shift)."""
    return ((_mix64(keys) >> np.uint64(slot_bits)) & np.uint64(world - 1)).astype(np.int64)


def _exchange(per_dest):
    """per_dest[d] = what this rank sends to rank d; returns [what rank r sent to this rank for all r]."""
    box = [None] * dist.get_world_size()
    dist.all_gather_object(box, per_dest)
    return [box[r][dist.get_rank()] for r in range(dist.get_world_size())]


def _routed_worker(rank, world, port, out, flavour):
    """The message flow of k_routed_request / _lookup / _records / apply / _rows with the oracle's pieces on gloo: a rank
    only ever reads and writes the states it owns; keys travel to the owner, rows and max Q come back, records go to the
    owner of s and are applied there in ascending global env order."""
    sys.path.insert(0, ROOT)
    import oracle
    from g2048 import dist as gdist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(N_TOTAL, rank, world)
    n, lr, gamma, eps = hi - lo, np.float32(0.1), np.float32(0.99), 0.4
    boards = np.zeros(n, np.uint64)
    oracle.env_reset(boards, None, None, None, seed=SEED, episode_idx=0, env_id_base=lo)
    aux, score = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
    shard = oracle.QTable(1 << 16, f32=True)
    thresh = oracle.eps_threshold(eps)

    def ask(keys):                       # keys -> rows, each answered by the owner of the key from ITS shard
        own = _owner(keys, world)
        got = _exchange([keys[own == d] for d in range(world)])
        answers = _exchange([np.array([shard.get(k)[0] for k in q], np.float64).reshape(-1, 4).astype(np.float32) for q in got])
        rows = np.zeros((len(keys), 4), np.float32)
        for d in range(world):
            rows[own == d] = answers[d]
        return rows

    for t in range(STEPS_ROUTED):
        rows = ask(boards)                                             # the rows as they are after the last apply
        draws = np.array([oracle.philox(SEED, lo + i, t, 0) for i in range(n)], np.uint32)
        actions = np.where(draws[:, 2] < thresh, draws[:, 3] >> 30, rows.argmax(1)).astype(np.uint8)
        s = boards.copy()
        reward, flags, _, _ = oracle.env_step(boards, aux, score, actions, None, flavour, SEED, t, lo)
        done = ((flags >> 2) & 1).astype(bool)
        assert not done.any()                                          # (no game ends this early: no reset in the emulation)
        best = ask(boards).max(1)                                      # max Q(s') BEFORE this step's apply
        target = reward.astype(np.float32) + np.where(done, np.float32(0), gamma * best).astype(np.float32)
        own = _owner(s, world)
        got = _exchange([(s[own == d], actions[own == d], target[own == d]) for d in range(world)])
        keys = np.concatenate([g[0] for g in got])                     # rank order = ascending global env order
        assert np.all(_owner(keys, world) == rank)
        shard.apply_targets_f32(keys, np.concatenate([g[1] for g in got]), np.concatenate([g[2] for g in got]), float(lr))
    keys, rows = shard.export()
    nz = np.abs(rows).sum(1) > 0
    np.savez(os.path.join(out, f"routed{rank}.npz"), boards=boards, keys=keys[nz], rows=rows[nz], lo=lo, hi=hi)
    dist.destroy_process_group()


STEPS_ROUTED = 10


@pytest.mark.timeout(300)
@pytest.mark.parametrize("flavour", [0, 1])
def test_routed_protocol_on_two_ranks_equals_the_single_table_step(tmp_path, flavour):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_routed_worker, args=(2, port, str(tmp_path), flavour), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import oracle
    single = OracleEngine(0, N_TOTAL)
    for t in range(STEPS_ROUTED):
        oracle.qlearn_step_sync(single.boards, single.aux, single.score, single.tab, 0.1, 0.99, 0.4, flavour, SEED, t, 0)
    keys, rows = single.tab.export()
    nz = np.abs(rows).sum(1) > 0
    d = [np.load(tmp_path / f"routed{r}.npz") for r in range(2)]
    for x in d:
        assert np.array_equal(x["boards"], single.boards[int(x["lo"]):int(x["hi"])])
    assert len(np.intersect1d(d[0]["keys"], d[1]["keys"])) == 0 and min(len(d[0]["keys"]), len(d[1]["keys"])) > 100
    k = np.concatenate([d[0]["keys"], d[1]["keys"]])
    r = np.concatenate([d[0]["rows"], d[1]["rows"]])
    order = np.argsort(k)
    assert np.array_equal(k[order], keys[nz]) and np.array_equal(r[order], rows[nz])
