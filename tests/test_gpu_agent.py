"""GPU tests of the host-side mirror of the reference interface: batched env/agent classes (torch tensors
as device buffers) and the N = 1 drop-in adapters, against the oracle and the reference goldens."""
import random

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import g2048
    g2048.init(0)
    return g2048


def np_boards(t):
    return t.detach().cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("flavour,code", [("penalty", 0), ("nopenalty", 1)])
def test_synchronous_deterministic_qlearning_matches_oracle(g, flavour, code):
    """Batched synchronous Q-learning, deterministic mode: boards AND the float32 table are
    bit-identical to the oracle after every step, for a batch with heavy start-state collisions."""
    import torch
    n, steps, seed = 4096, 48, 4321
    env = g.BatchedGame2048Env(n, flavour, seed=seed)
    agent = g.BatchedQLearningAgent(1000, learning_rate=0.1, discount_factor=0.99, exploration_rate=0.3,
                                    capacity=1 << 20, seed=seed)
    env.reset()
    cb = np_boards(env.boards).copy()
    ca, cs = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
    tab = oracle.QTable(1 << 20, f32=True)
    total = np.zeros(16, np.int64)
    env.counters.zero_()
    for t in range(steps):
        agent.step_sync(env, mode="deterministic")
        c, _ = oracle.qlearn_step_sync(cb, ca, cs, tab, 0.1, 0.99, 0.3, code, seed, t, 0)
        total += c
        assert np.array_equal(np_boards(env.boards), cb), t
        if t % 8 == 7 or t == steps - 1:
            keys, rows = agent.export()
            wk, wr = tab.export()
            assert np.array_equal(keys, wk), t
            assert np.array_equal(rows, wr.astype(np.float32)), t
    total[4] = 0
    got = env.counters.cpu().numpy().copy()
    got[4] = 0
    assert np.array_equal(got[:4], total[:4]) and got[5] == total[5] and got[6] == total[6]
    assert np.array_equal(env.aux.cpu().numpy().view(np.uint64), ca)
    assert np.array_equal(env.score.cpu().numpy(), cs)


def test_sharded_delta_exchange_equals_single_gpu(g):
    """Multi-GPU semantics on one device: G virtual shards each emit (key, action, delta) records, the records
    are concatenated in rank order (what all_gather returns) and applied by every replica in deterministic
    mode -- the result equals the 1-GPU table bit for bit."""
    import torch
    n, G, steps, seed = 2048, 4, 24, 99
    one_env = g.BatchedGame2048Env(n, "penalty", seed=seed)
    one = g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.5, capacity=1 << 18, seed=seed)
    one_env.reset()
    shard_envs, replicas = [], []
    h = n // G
    for r in range(G):
        e = g.BatchedGame2048Env(h, "penalty", seed=seed, env_id_base=r * h)
        e.boards.copy_(one_env.boards[r * h:(r + 1) * h])
        shard_envs.append(e)
        replicas.append(g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.5, capacity=1 << 18, seed=seed))
    for t in range(steps):
        one.step_sync(one_env, mode="deterministic")
        recs = [replicas[r].step_sync(shard_envs[r], mode="deterministic", apply=False, records=True) for r in range(G)]
        keys = torch.cat([x[0] for x in recs]); acts = torch.cat([x[1] for x in recs]); dl = torch.cat([x[2] for x in recs])
        for r in range(G):
            replicas[r].apply_targets(keys, acts, dl, mode="deterministic")
    k1, r1 = one.export()
    for r in range(G):
        assert np.array_equal(np_boards(shard_envs[r].boards), np_boards(one_env.boards)[r * h:(r + 1) * h])
        # a replica only holds the states its own shard touched plus every updated state; compare on the updated ones
        kr, rr = replicas[r].export()
        common = np.isin(k1, kr)
        nz = np.abs(r1).sum(1) > 0
        assert common[nz].all()
        idx = np.searchsorted(kr, k1[nz])
        assert np.array_equal(rr[idx], r1[nz])


def test_batched_env_step_api_and_encodings(g):
    import torch
    n, seed = 10_000, 17
    for flavour, code in (("penalty", 0), ("nopenalty", 1)):
        env = g.BatchedGame2048Env(n, flavour, seed=seed, env_id_base=5)
        env.reset()
        cb = np_boards(env.boards).copy()
        ca, cs = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
        rng = np.random.RandomState(code)
        for t in range(40):
            actions = rng.randint(0, 4, n).astype(np.uint8)
            boards, reward, done, max_number = env.step(torch.from_numpy(actions))
            wr, wf, wm, wms = oracle.env_step(cb, ca, cs, actions, None, code, seed, t, 5)
            assert np.array_equal(np_boards(boards), cb)
            assert np.array_equal(reward.cpu().numpy().view(np.uint64), wr.view(np.uint64))
            assert np.array_equal(done.cpu().numpy(), (wf & 4) != 0)
            assert np.array_equal(max_number.cpu().numpy(), 1 << wm.astype(np.int64))
            assert np.array_equal(env.flags.cpu().numpy(), wf)
        # unpack / pack / one-hot / legal mask on the resulting boards
        tiles = env.tiles()
        assert np.array_equal(tiles.cpu().numpy(), oracle.unpack_i64(cb))
        env2 = g.BatchedGame2048Env(n, flavour)
        env2.set_tiles(tiles)
        assert np.array_equal(np_boards(env2.boards), cb)
        assert np.array_equal(env.encode_onehot().cpu().numpy(), oracle.encode_onehot(cb))
        bf = env.encode_onehot(dtype=torch.bfloat16).float().cpu().numpy()
        assert np.array_equal(bf, oracle.encode_onehot(cb))
        assert np.array_equal(env.legal_mask().cpu().numpy(), oracle.legal_mask(cb))
        bad = tiles.clone(); bad[3, 1, 2] = 6
        with pytest.raises(ValueError):
            env2.set_tiles(bad)


def test_select_action_matches_dqn_agent_rules(g):
    """act / act_ripetitive (Dqn8TestNOPERCNN.py:312-336) restated in numpy on the same Philox draws."""
    import torch
    L = g.lib()
    n, seed, step, base = 50_000, 5, 9, 77
    rng = np.random.RandomState(0)
    q = rng.standard_normal((n, 4)).astype(np.float32)
    q[::7, 2] = q[::7, 0]   # ties: first maximum wins
    legal = rng.randint(0, 16, n).astype(np.uint8)
    qd, ld = torch.from_numpy(q).cuda(), torch.from_numpy(legal).cuda()
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for eps in (0.0, 0.4, 1.0):
        thresh = oracle.eps_threshold(eps)
        x = np.array([oracle.philox(seed, base + i, step, 0) for i in range(0, n, 50)])
        for use_legal in (False, True):
            rc = L.g2048_select_action(qd.data_ptr(), ld.data_ptr() if use_legal else None, out.data_ptr(), n, eps, seed,
                                       step, base, torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            got = out.cpu().numpy()[::50]
            for j, i in enumerate(range(0, n, 50)):
                explore = int(x[j][2]) < thresh
                lm = int(legal[i]) if use_legal else 0
                moves = [a for a in range(4) if (lm >> a) & 1]
                if not moves:
                    want = int(x[j][3]) >> 30 if explore else int(np.argmax(q[i]))
                elif explore:
                    want = moves[(int(x[j][3]) * len(moves)) >> 32]
                else:
                    want = moves[int(np.argmax([q[i][a] for a in moves]))]
                assert got[j] == want, (i, eps, use_legal)


# ------------------------------------------------------------------------------------------ drop-in adapters
@pytest.mark.parametrize("flavour", ["penalty", "nopenalty"])
def test_env_adapter_reproduces_reference_under_numpy_seed(g, golden, flavour):
    """`Game2048_env` adapter with rng="numpy": same np.random.seed, same actions -> the reference's boards,
    rewards (float64 bits), dones, max tiles and scores, step for step (recorded by oracle/make_golden.py)."""
    gd = golden("compat_seeded")
    np.random.seed(4242)
    env = g.Game2048_env(flavour=flavour)
    assert env.action_space.n == 4 and env.observation_space.shape == (4, 4)
    assert g.pack_tiles(env.game.board) == int(gd[f"{flavour}_first"])
    for t, a in enumerate(gd[f"{flavour}_actions"]):
        board, reward, done, max_number = env.step(int(a))
        assert float(reward) == gd[f"{flavour}_rewards"][t], t
        assert bool(done) == bool(gd[f"{flavour}_dones"][t]) and int(max_number) == gd[f"{flavour}_max"][t], t
        assert env.score == gd[f"{flavour}_scores"][t], t
        if flavour == "nopenalty":
            env.game.board = board
        if done:
            board = env.reset()
        assert g.pack_tiles(board) == int(gd[f"{flavour}_boards"][t]), t


def test_reference_training_loop_runs_unchanged_on_the_adapters(g, golden):
    """The loop of QLearningBase/Agent/main.py:80-109 on the adapters, seeded like the golden run: same actions,
    boards and rewards as the reference; final Q-table within float32 tolerance of the reference's float64 one."""
    gd = golden("compat_seeded")
    episodes, lr, gamma, eps0 = gd["loop_params"]
    np.random.seed(1)
    random.seed(1)
    env = g.Game2048_env()
    agent = g.QLearningAgent(int(episodes), action_space=env.action_space.n, learning_rate=lr, discount_factor=gamma,
                             exploration_rate=eps0)
    i = 0
    for episode in range(int(episodes)):
        state = tuple(map(tuple, env.reset()))
        done = False
        total_reward = 0
        while not done:
            action = agent.choose_action(state)
            next_state, reward, done, info = env.step(action)
            next_state = tuple(map(tuple, next_state))
            q_values = agent.q_table[state]
            agent.update_q_value(state, action, reward, next_state, done)
            state = next_state
            total_reward += reward
            assert action == gd["loop_actions"][i] and reward == gd["loop_rewards"][i], (episode, i)
            assert g.pack_tiles(next_state) == int(gd["loop_boards"][i]), (episode, i)
            i += 1
        assert total_reward == gd["loop_totals"][episode]
        agent.decay_exploration(episode)
    assert i == len(gd["loop_actions"]) and agent.epsilon == float(gd["loop_eps_final"])
    keys, rows = agent.export()
    assert np.array_equal(keys, gd["loop_q_keys"]) and len(agent.q_table) == len(keys)
    np.testing.assert_allclose(rows, gd["loop_q_rows"], rtol=1e-5, atol=1e-5)
    some = agent.to_dict()
    assert len(some) == len(keys) and all(len(k) == 4 for k in list(some)[:5])


def test_drivers_and_checkpoint(g, golden, tmp_path):
    """train_tabular == the reference loop (same seeded trajectory as the golden run, CSV in the reference's format);
    train_tabular_batched learns something; Q-table checkpoint round trip."""
    import csv
    gd = golden("compat_seeded")
    episodes = int(gd["loop_params"][0])
    np.random.seed(1)
    random.seed(1)
    env = g.Game2048_env()
    agent = g.QLearningAgent(episodes, action_space=4, learning_rate=0.1, discount_factor=0.99, exploration_rate=0.95)
    log = tmp_path / "debug_log.csv"
    hist = g.train_tabular(env, agent, episodes, log_file=str(log))
    assert [h[0] for h in hist] == gd["loop_totals"].tolist()
    rows = list(csv.reader(open(log)))
    assert rows[0] == ["Episode", "Action", "Q-Values", "Reward", "Total-Reward", "Max Value"] and len(rows) == episodes + 1

    benv = g.BatchedGame2048Env(1 << 12, "penalty", seed=5)
    bagent = g.BatchedQLearningAgent(12, 4, 0.1, 0.99, 0.9, capacity=1 << 22, seed=5)
    hist = g.train_tabular_batched(benv, bagent, 12, steps_per_epoch=16, log_file=str(tmp_path / "batched.csv"))
    assert len(hist) == 12 and hist[-1][1] < hist[0][1]            # epsilon decayed
    assert hist[-1][7] > hist[0][7] > 0                           # table grows
    assert bagent.epsilon == pytest.approx(0.01)
    keys, rows_ = bagent.export()
    path = tmp_path / "q.pt"
    bagent.save(str(path))
    other = g.BatchedQLearningAgent(12, 4, capacity=1 << 22)
    other.load(str(path))
    k2, r2 = other.export()
    assert np.array_equal(keys, k2) and np.array_equal(rows_, r2) and other.epsilon == bagent.epsilon
    d = bagent.to_dict()
    k0 = next(iter(d))
    assert len(d) == len(keys) and len(k0) == 4 and len(k0[0]) == 4


def test_adapter_game_core_api(g):
    """Game2048 adapter surface used by the reference's drivers: .board readable and assignable
    (main.py:85, mainDQL_CNN_step2.py:163,237), move(a, trial=True) -> (legal, score) without side effects
    (mainDQL_CNN_step2.py:169-174), is_game_over(), add_number(), showMatrix()."""
    np.random.seed(3)
    env = g.Game2048_env(flavour="nopenalty")
    board = np.array([[2, 2, 4, 0], [0, 0, 0, 0], [8, 8, 8, 8], [2, 4, 2, 4]], dtype=np.int64)
    env.game.board = board.copy()
    state = np.random.get_state()[1].copy()
    legal = []
    for a in range(4):
        moved, score = env.game.move(a, trial=True)
        nb, mv, sc = oracle.move(np.array([g.pack_tiles(board)], np.uint64), np.array([a], np.uint8))
        assert bool(moved) == bool(mv[0]) and int(score) == int(sc[0])
        legal.append(bool(moved))
    assert legal == [True, True, True, True]
    assert np.array_equal(env.game.board, board)                       # trial moves do not touch the board
    assert np.array_equal(np.random.get_state()[1], state)            # ... nor the global RNG
    assert not env.game.is_game_over()
    dead = np.array([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]], dtype=np.int64)
    env.game.board = dead
    assert env.game.is_game_over() and all(not env.game.move(a, trial=True)[0] for a in range(4))
    nxt, reward, done, mx = env.step(0)                                 # a step from a dead board ends the episode
    assert done and reward == 0 and mx == 4 and np.array_equal(nxt, dead)
    penv = g.Game2048_env()                                              # penalty flavour: move() mutates and spawns
    penv.game.board = board.copy()
    moved, score = penv.game.move(0)
    assert moved and score == 4 + 16 + 16 and (penv.game.board != 0).sum() == (board != 0).sum() - 3 + 1
    before = (penv.game.board != 0).sum()
    penv.game.add_number()
    assert (penv.game.board != 0).sum() == before + 1
    penv.showMatrix()


def test_evaluate_tabular_greedy_play(g):
    np.random.seed(11)
    random.seed(11)
    env = g.Game2048_env()
    agent = g.QLearningAgent(5, action_space=4, exploration_rate=0.9)
    g.train_tabular(env, agent, 5)
    eps = agent.epsilon
    res = g.evaluate_tabular(env, agent, episodes=3)
    assert len(res) == 3 and all(r[2] > 0 and r[1] >= 2 for r in res) and agent.epsilon == eps


def test_deterministic_update_with_very_long_runs_matches_oracle(g):
    """The first steps after a reset of 300,000 envs: a few hundred (state, action) pairs receive thousands of
    targets each (runs far beyond 32 records, handled by the warp-cooperative k_long_run_apply) -- the float32 table
    must still be the sequential result bit for bit, and the step must not take long."""
    import time
    import torch
    n, steps, seed = 300_000, 6, 99
    env = g.BatchedGame2048Env(n, "penalty", seed=seed)
    agent = g.BatchedQLearningAgent(1000, learning_rate=0.1, discount_factor=0.99, exploration_rate=0.05,
                                    capacity=1 << 22, seed=seed)
    env.reset()
    cb = np_boards(env.boards).copy()
    ca, cs = np.full(n, oracle.AUX_INIT, np.uint64), np.zeros(n, np.int32)
    tab = oracle.QTable(1 << 22, f32=True)
    agent.step_sync(env, mode="deterministic")
    oracle.qlearn_step_sync(cb, ca, cs, tab, 0.1, 0.99, 0.05, 0, seed, 0, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(1, steps):
        agent.step_sync(env, mode="deterministic")
    torch.cuda.synchronize()
    per_step = (time.perf_counter() - t0) / (steps - 1)
    for t in range(1, steps):
        oracle.qlearn_step_sync(cb, ca, cs, tab, 0.1, 0.99, 0.05, 0, seed, t, 0)
    assert np.array_equal(np_boards(env.boards), cb)
    keys, rows = agent.export()
    wk, wr = tab.export()
    assert np.array_equal(keys, wk) and np.array_equal(rows, wr.astype(np.float32))
    # the hottest (state, action) really is a long run
    k, a, _ = agent.step_sync(env, mode="deterministic", apply=False, records=True)
    pairs = (k.cpu().numpy().view(np.uint64) << np.uint64(2)) | a.cpu().numpy().astype(np.uint64)
    assert np.unique(pairs, return_counts=True)[1].max() > 32
    assert per_step < 0.02, per_step          # 16 ms per step at 8 M records before the cooperative kernel


def test_probe_stats_match_a_host_recount(g):
    """g2048_qtable_probe_stats: states, mean and max probe length recomputed on the host from the raw slots."""
    n = 20_000
    env = g.BatchedGame2048Env(n, "penalty", seed=3)
    agent = g.BatchedQLearningAgent(1000, 4, 0.1, 0.99, 0.5, capacity=1 << 18, seed=3)   # fills to a load of ~0.7
    env.reset()
    agent.rollout(env, 12)
    st = agent.probe_stats()
    keys = agent.table.view(-1, 4)[:, 0].cpu().numpy().view(np.uint64)
    pos = np.nonzero(keys)[0].astype(np.uint64)
    k = keys[pos.astype(np.int64)]
    with np.errstate(over="ignore"):
        x = k.copy()
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    mask = np.uint64((1 << 18) - 1)
    home = x & mask
    # the probe sequence goes pair by pair (the two slots of one 64-byte fill), the home slot's side first
    one = np.uint64(1)
    d = np.uint64(2) * (((pos >> one) - (home >> one)) & (mask >> one)) + ((pos ^ home) & one)
    assert st["states"] == len(k) == len(agent)
    assert st["max_probe_length"] == 1 + int(d.max())
    assert abs(st["mean_probe_length"] - (1 + d.astype(np.float64).mean())) < 1e-9
    assert 0.3 < st["load_factor"] < 0.95 and st["mean_probe_length"] > 1.1
