"""The PyTorch DQN restatement: architecture facts on CPU, the batched agent on GPU."""
import numpy as np
import pytest
import torch

import oracle


def test_model_has_the_reference_parameter_count():
    """Dqn8TestNOPERCNN.py:17 '198 milioni di parametri'; exact Keras count 197,204,996."""
    from g2048 import dqn
    with torch.device("meta"):
        m = dqn.DQNModel()
    assert sum(p.numel() for p in m.parameters()) == 197_204_996


def test_forward_shapes_and_same_padding_on_cpu():
    from g2048 import dqn
    torch.manual_seed(0)
    m = dqn.DQNModel(width=16, hidden=8).eval()
    boards = np.array([0x0000000000000021, 0x123456789ABCDEF1], np.uint64)
    x = torch.from_numpy(oracle.encode_onehot(boards))
    y = m(x)
    assert y.shape == (2, 4) and torch.isfinite(y).all()
    # 'same' for the even kernels pads (k-1)//2 before and k//2 after, like TensorFlow
    conv = m.blocks[0].convs[1]           # kernel 2
    xin = x.permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xin, (0, 1, 0, 1)), conv.weight, conv.bias)
    assert torch.allclose(conv(xin), ref, atol=1e-6)


def keras_semantics_forward(x, params):
    """The reference network (Dqn8TestNOPERCNN.py:209-246) written out in numpy the way Keras/TensorFlow evaluates it:
    Input(shape=(16, 4, 4)) is read channels-last (H = 16 levels, W = 4 rows, C = 4 columns); Conv2D kernels are
    [kh, kw, c_in, c_out] with TF 'SAME' padding -- (k - 1) // 2 zeros before, k // 2 after, per spatial dimension;
    conv_block = four parallel convolutions (k = 1..4) concatenated on the channel axis, ReLU; Flatten in (h, w, c)
    order; Dense kernels are [in, out]; Dropout is the identity at inference."""
    def conv_same(a, kernel, bias):                       # a [N, H, W, C]
        kh, kw, _, cout = kernel.shape
        ap = np.pad(a, ((0, 0), ((kh - 1) // 2, kh // 2), ((kw - 1) // 2, kw // 2), (0, 0)))
        n, h, w, _ = a.shape
        out = np.zeros((n, h, w, cout), np.float64)
        for i in range(kh):
            for j in range(kw):
                out += np.einsum("nhwc,co->nhwo", ap[:, i:i + h, j:j + w, :], kernel[i, j])
        return out + bias
    a = x.astype(np.float64)                              # [N, 16, 4, 4] read as NHWC
    for block in params["blocks"]:
        a = np.maximum(np.concatenate([conv_same(a, k, b) for k, b in block], axis=-1), 0.0)
    flat = a.reshape(a.shape[0], -1)
    hid = np.maximum(flat @ params["fc1"][0] + params["fc1"][1], 0.0)
    return hid @ params["fc2"][0] + params["fc2"][1]


def test_model_forward_equals_a_numpy_keras_semantics_reference():
    """DQNModel (PyTorch, channels-first, padding='same') against the numpy restatement above with the weights
    transposed into Keras layout -- pins the channels-last reading of Input(16,4,4), the SAME padding of the even
    kernels, the concatenation order and the Flatten order at reduced width (the arithmetic is width-independent)."""
    from g2048 import dqn
    torch.manual_seed(3)
    m = dqn.DQNModel(width=24, hidden=10).double().eval()
    rng = np.random.RandomState(0)
    boards = np.array([int(rng.randint(0, 1 << 62)) for _ in range(6)] + [0x0000000000000021, 0xFFFFFFFFFFFFFFFF], np.uint64)
    x = oracle.encode_onehot(boards)                                        # [N, 16, 4, 4] as encode_state returns it
    params = {"blocks": [[(c.weight.detach().numpy().transpose(2, 3, 1, 0), c.bias.detach().numpy()) for c in blk.convs]
                         for blk in m.blocks],
              "fc1": (m.fc1.weight.detach().numpy().T, m.fc1.bias.detach().numpy()),
              "fc2": (m.fc2.weight.detach().numpy().T, m.fc2.bias.detach().numpy())}
    want = keras_semantics_forward(x, params)
    got = m(torch.from_numpy(x).double()).detach().numpy()
    assert got.shape == want.shape == (8, 4)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_reference_driver_plays_invalid_moves_and_filters_repeats():
    """dqn_step(reference_driver=True) = mainDQL_CNN_step2.py:176-185, :220 + Dqn8TestNOPERCNN.py:279-297 per env:
    act() is unrestricted (invalid moves are played, cost -10, are stored once), a transition repeating the env's
    previously stored (state, next_state) is dropped and the next action is then act_ripetitive() (a legal move)."""
    import g2048
    from g2048 import dqn
    torch.manual_seed(0)
    n = 2048
    env = g2048.BatchedGame2048Env(n, "nopenalty", seed=5)
    agent = dqn.BatchedDQNAgent(width=16, hidden=16, memory_size=1 << 17, batch_size=64, epsilon=0.9, epsilon_decay=1.0)
    env.reset()
    stored, invalid_rewards, forced_legal = 0, 0, 0
    for t in range(60):
        before = env.boards.clone()
        legal = env.legal_mask()
        unsaved = None if agent.memory_saved is None else ~agent.memory_saved
        entries = agent.nb_entries
        reward, done = dqn.dqn_step(env, agent, reference_driver=True)
        kept = agent.nb_entries - entries
        stored += kept
        assert 0 < kept <= n
        invalid_rewards += int((reward == -10).sum())
        if unsaved is not None and bool(unsaved.any()):
            # envs whose previous transition was dropped were given a legal move: their board changed (or the game was over)
            moved = (env.boards != before) | done | (legal == 0)
            assert bool(moved[unsaved].all())
            forced_legal += int(unsaved.sum())
    assert invalid_rewards > 0 and forced_legal > 0 and stored < 60 * n      # repeats were dropped
    # no two consecutive stored transitions of an env are the same (state, next_state) pair unless a game ended
    assert agent.nb_entries == stored


@pytest.mark.gpu
def test_remember_with_more_transitions_than_slots_keeps_the_newest():
    from g2048 import dqn
    agent = dqn.BatchedDQNAgent(width=16, hidden=16, memory_size=1000, batch_size=8)
    n = 2500
    dev = agent.device
    st = torch.arange(1, n + 1, dtype=torch.int64, device=dev)
    agent.remember(st, torch.zeros(n, dtype=torch.uint8, device=dev), st.float(), torch.zeros(n, dtype=torch.bool, device=dev), st + 7)
    assert agent.nb_entries == 1000
    order = torch.argsort(agent.mem_state)
    assert torch.equal(agent.mem_state[order], st[-1000:])                  # the newest 1000, each exactly once
    assert torch.equal(agent.mem_next[order], st[-1000:] + 7) and torch.equal(agent.mem_reward[order], st[-1000:].float())


@pytest.mark.gpu
def test_evaluate_random_and_dqn_play_modes():
    """GameDemo.py:272-316 headless: random play on the N = 1 adapter, greedy-legal network play on a batched env."""
    import g2048
    from g2048 import dqn
    env1 = g2048.Game2048_env(flavour="nopenalty", seed=3)
    res = g2048.evaluate_random(env1, episodes=2, rng=np.random.RandomState(1))
    assert len(res) == 2 and all(steps > 20 and tile >= 8 for _, tile, steps in res)
    env = g2048.BatchedGame2048Env(256, "nopenalty", seed=9)
    agent = dqn.BatchedDQNAgent(width=16, hidden=16, memory_size=1024, batch_size=8)
    out = g2048.evaluate_dqn(env, agent, episodes=1, max_steps=3000)
    assert len(out["scores"]) == len(out["max_tiles"]) == 256 and min(out["max_tiles"]) >= 4
    assert agent.nb_entries == 0                                             # nothing was learned or stored


@pytest.mark.gpu
def test_batched_dqn_agent_trains_on_gpu_envs():
    import g2048
    from g2048 import dqn
    torch.manual_seed(0)
    n = 4096
    env = g2048.BatchedGame2048Env(n, "nopenalty", seed=11)
    agent = dqn.BatchedDQNAgent(width=32, hidden=64, memory_size=1 << 16, batch_size=256, epsilon=0.5, learning_rate=1e-3)
    env.reset()
    # encode_state == the oracle's restatement of Dqn8TestNOPERCNN.py:271-277
    enc = agent.encode_state(env.boards).cpu().numpy()
    assert np.array_equal(enc, oracle.encode_onehot(env.boards.cpu().numpy().view(np.uint64)))
    for t in range(40):
        state = env.boards.clone()
        legal = env.legal_mask()
        a = agent.act_ripetitive(state, legal)
        ok = ((legal.to(torch.int64) >> a.to(torch.int64)) & 1).bool() | (legal == 0)
        assert bool(ok.all())                      # never an illegal move while a legal one exists
        dqn.dqn_step(env, agent)
    assert agent.nb_entries == 40 * n and agent.nb_entries <= agent.memory_size or agent.nb_entries == agent.memory_size
    losses = [agent.replay() for _ in range(60)]
    assert all(np.isfinite(l) for l in losses)
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    agent.update_target_model()
    for p, q in zip(agent.model.parameters(), agent.target_model.parameters()):
        assert torch.equal(p, q)
    # terminal bonus rule of the driver (mainDQL_CNN_step2.py:202-213)
    b = torch.tensor([0xB, 0xAA, 0xA9, 0x9], dtype=torch.int64, device="cuda")
    d = torch.tensor([True, True, True, True], device="cuda")
    assert dqn.terminal_bonus(b, d).tolist() == [100.0, 50.0, 0.0, 0.0]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_dqn_env_step_equals_the_unfused_pieces(dtype):
    """g2048_dqn_env_step (one launch) == select_action -> env_step -> terminal bonus -> masked reset -> legal mask ->
    encode_onehot, each of which is checked against the oracle elsewhere; late-game boards so that episodes end."""
    import g2048
    from g2048 import dqn
    L = g2048.lib()
    n, seed, base = 10_000, 21, 7
    rng = np.random.RandomState(0)
    lv = rng.randint(1, 12, size=(n, 16)) * (rng.random_sample((n, 16)) < 0.9)
    packed = np.zeros(n, np.uint64)
    for j in range(16):
        packed |= lv[:, j].astype(np.uint64) << np.uint64(4 * j)
    packed[packed == 0] = 1
    start = torch.from_numpy(packed.view(np.int64)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    code = 0 if dtype == torch.float32 else 1
    fused = g2048.BatchedGame2048Env(n, "nopenalty", seed=seed, env_id_base=base)
    plain = g2048.BatchedGame2048Env(n, "nopenalty", seed=seed, env_id_base=base)
    fused.boards.copy_(start)
    plain.boards.copy_(start)
    legal_f = fused.legal_mask()
    out = {k: torch.empty(n, dtype=d, device="cuda") for k, d in (("a", torch.uint8), ("s", torch.int64), ("s2", torch.int64),
                                                                 ("r", torch.float32), ("d", torch.uint8))}
    onehot = torch.empty((n, 16, 4, 4), dtype=dtype, device="cuda")
    total_done = 0
    for t in range(30):
        q = torch.randn((n, 4), device="cuda")
        q[::5, 1] = q[::5, 0]
        eps = 0.3
        # unfused pieces
        legal_p = plain.legal_mask()
        a_p = torch.empty(n, dtype=torch.uint8, device="cuda")
        assert L.g2048_select_action(q.data_ptr(), legal_p.data_ptr(), a_p.data_ptr(), n, eps, seed, t, base, st) == 0
        s_p = plain.boards.clone()
        plain.step_idx = t
        nxt, r_p, d_p, _ = plain.step(a_p)
        nxt = nxt.clone()
        r_p = r_p.to(torch.float32) + dqn.terminal_bonus(nxt, d_p)
        plain.episode_idx = 100 + t
        plain.reset(mask=d_p)
        # fused
        rc = L.g2048_dqn_env_step(fused.boards.data_ptr(), fused.score.data_ptr(), q.data_ptr(), legal_f.data_ptr(),
                                  out["a"].data_ptr(), out["s"].data_ptr(), out["s2"].data_ptr(), out["r"].data_ptr(),
                                  out["d"].data_ptr(), legal_f.data_ptr(), onehot.data_ptr(), code, n, eps, 3, seed, t,
                                  100 + t, base, st)
        assert rc == 0, L.g2048_last_error()
        assert torch.equal(out["a"], a_p) and torch.equal(out["s"], s_p) and torch.equal(out["s2"], nxt), t
        assert torch.equal(out["r"], r_p) and torch.equal(out["d"] != 0, d_p), t
        assert torch.equal(fused.boards, plain.boards) and torch.equal(fused.score, plain.score), t
        assert torch.equal(legal_f, plain.legal_mask()), t
        assert torch.equal(onehot, plain.encode_onehot(dtype=dtype)), t
        total_done += int(d_p.sum())
    assert total_done > 100 and float(out["r"].max()) >= 50.0 or total_done > 100


@pytest.mark.gpu
def test_fused_dqn_feed_drives_training():
    import g2048
    from g2048 import dqn
    torch.manual_seed(1)
    env = g2048.BatchedGame2048Env(2048, "nopenalty", seed=3)
    agent = dqn.BatchedDQNAgent(width=16, hidden=32, memory_size=1 << 15, batch_size=128, epsilon=0.5, learning_rate=1e-3)
    env.reset()
    feed = dqn.FusedDQNFeed(env, agent)
    for _ in range(20):
        reward, done = feed.step()
    assert agent.nb_entries == min(20 * 2048, agent.memory_size) and agent.step_counter == 20
    assert np.isfinite(agent.replay())


@pytest.mark.gpu
def test_train_dqn_driver_cadence(tmp_path):
    """train_dqn = mainDQL_CNN_step2.py:151-333 batched: episode bookkeeping, replay burst at episode end, target
    sync, lr cut after a game that reached 1024, periodic save, CSV log."""
    import csv
    import g2048
    from g2048 import dqn
    torch.manual_seed(2)
    n = 1024
    env = g2048.BatchedGame2048Env(n, "nopenalty", seed=5)
    agent = dqn.BatchedDQNAgent(width=16, hidden=32, memory_size=1 << 14, batch_size=64, epsilon=0.5, learning_rate=1e-3)
    env.reset()
    rng = np.random.RandomState(1)
    lv = rng.randint(1, 11, size=(n, 16)) * (rng.random_sample((n, 16)) < 0.95)      # crowded late-game boards
    lv[:256, 0] = 10                                                                   # a few hold a 1024 tile
    packed = np.zeros(n, np.uint64)
    for j in range(16):
        packed |= lv[:, j].astype(np.uint64) << np.uint64(4 * j)
    env.boards.copy_(torch.from_numpy(packed.view(np.int64)).cuda())
    log = tmp_path / "dqn.csv"
    out = g2048.train_dqn(env, agent, 40, replays_per_episode=2, max_replays_per_step=6, target_sync_episodes=20,
                          save_every_episodes=50, save_dir=str(tmp_path / "saves"), log_file=str(log))
    assert out["steps"] == 40 and out["episodes"] > 50
    assert len(out["max_tile_list"]) == out["episodes"] == len(out["score_list"])
    assert out["best_tile"] == max(out["max_tile_list"]) and all(t & (t - 1) == 0 for t in out["max_tile_list"])
    assert all(np.isfinite(l) for l in out["loss_history"]) and out["loss_history"]
    rows = list(csv.reader(open(log)))
    assert rows[0] == g2048.train.DQN_CSV_HEADER and len(rows) == 41
    assert int(rows[-1][1]) == out["episodes"]
    assert agent.optimizer.param_groups[0]["lr"] < 1e-3            # at least one finished game held a 1024 tile
    saves = sorted(p.name for p in (tmp_path / "saves").iterdir())
    assert saves and all(s.startswith("agent_episode_") for s in saves)
    # resume from the save
    other = dqn.BatchedDQNAgent(width=16, hidden=32, memory_size=1 << 14, batch_size=64)
    other.load_agent_state(str(tmp_path / "saves" / saves[-1]))
    assert other.nb_entries > 0 and other.step_counter > 0


@pytest.mark.gpu
def test_encode_and_greedy_action_selection_match_the_reference_agent(golden):
    """g2048_encode_onehot and the greedy branch of g2048_select_action against goldens recorded from the reference's
    own DQNAgent.encode_state / act / act_ripetitive (Dqn8TestNOPERCNN.py:271-277, :312-336): argmax with ties, the
    restriction to the legal moves, and the fallback when no move is legal."""
    import g2048
    g2048.init(0)
    L = g2048.lib()
    g = golden("dqn_agent")
    st = torch.cuda.current_stream().cuda_stream
    boards, bad = oracle.pack_i64(g["tiles"])
    assert bad == 0
    b = torch.from_numpy(boards.view(np.int64)).cuda()
    for code, dtype in ((0, torch.float32), (1, torch.bfloat16)):
        out = torch.empty((len(boards), 16, 4, 4), dtype=dtype, device="cuda")
        assert L.g2048_encode_onehot(b.data_ptr(), out.data_ptr(), len(boards), code, st) == 0
        assert np.array_equal(out.float().cpu().numpy(), g["onehot"])
    n = len(g["q"])
    q = torch.from_numpy(g["q"]).cuda()
    legal = torch.from_numpy(g["legal"]).cuda()
    a = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert L.g2048_select_action(q.data_ptr(), None, a.data_ptr(), n, 0.0, 1, 0, 0, st) == 0
    assert np.array_equal(a.cpu().numpy(), g["act"])
    assert L.g2048_select_action(q.data_ptr(), legal.data_ptr(), a.data_ptr(), n, 0.0, 1, 0, 0, st) == 0
    assert np.array_equal(a.cpu().numpy(), g["act_ripetitive"])
    assert (g["act"] != g["act_ripetitive"]).sum() > 100          # the legal-move restriction matters in the sample
