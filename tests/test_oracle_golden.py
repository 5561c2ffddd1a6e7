"""The CPU oracle (oracle/g2048_oracle.c) against golden vectors recorded from the reference itself.

The goldens were produced by oracle/make_golden.py importing the unmodified reference
(QLearningBase/environment/Game2048_env.py, Deep_QLearning/environment/Game2048_nopenalty_env.py,
QLearningBase/Agent/main.py).  Everything here is bit-exact: boards, flags, scores, aux state and the
float64 shaped reward (same libm expressions in the same order).
"""
import numpy as np
import pytest

import oracle


def replay_env(g, flavour, step_fn):
    """Carried replay: the implementation keeps its own boards/aux/score across steps; boards are
    reloaded only where the golden says a new episode (reset/teleport) starts."""
    E, T = g["action"].shape
    boards = g["board_in"][:, 0].copy()
    aux = np.full(E, oracle.AUX_INIT, np.uint64)
    score = np.zeros(E, np.int32)
    for t in range(T):
        reload = g["reload"][:, t].astype(bool)
        boards[reload] = g["board_in"][reload, t]
        if t:
            prev_done = (g["flags"][:, t - 1] >> 2) & 1
            score[prev_done.astype(bool)] = 0  # env.reset() zeroes env.score (Game2048_env.py:190)
        assert np.array_equal(boards, g["board_in"][:, t]), f"carried board differs at step {t}"
        reward, flags, maxlvl, ms = step_fn(boards, aux, score, np.ascontiguousarray(g["action"][:, t]),
                                            np.ascontiguousarray(g["draws"][:, t]), flavour)
        assert np.array_equal(boards, g["board_out"][:, t]), f"board_out step {t}"
        assert np.array_equal(flags & 7, g["flags"][:, t]), f"flags step {t}"
        assert np.array_equal(maxlvl, g["maxlvl"][:, t]), f"maxlvl step {t}"
        assert np.array_equal(ms, g["move_score"][:, t]), f"move_score step {t}"
        assert np.array_equal(score, g["env_score"][:, t]), f"env.score step {t}"
        assert np.array_equal(reward.view(np.uint64), g["reward"][:, t].view(np.uint64)), f"reward bits step {t}"
        if flavour == oracle.FLAVOUR_PENALTY:
            assert np.array_equal(aux, g["aux_out"][:, t]), f"aux step {t}"


def oracle_step(boards, aux, score, actions, draws, flavour):
    return oracle.env_step(boards, aux, score, actions, draws, flavour)


FNV = np.uint64(0x100000001B3)


def replay_config2(g, flavour, step_fn, reset_fn):
    """BASELINE config 2 at its stated size (4,096 envs x 512 steps, SURVEY.md 8d "C2"): replay the recorded actions and
    spawn draws of the reference step by step (all envs at once), folding everything a step returns into a per-env
    64-bit digest the way oracle/make_golden.py:record_config2 did on the reference's own outputs; the digests, the
    final boards, scores and aux words must be the reference's."""
    E, T, _ = (int(x) for x in g["shape"])
    a4 = g["actions4"]
    actions = np.zeros((E, T), np.uint8)
    for j in range(4):
        actions[:, j::4] = (a4 >> (2 * j)) & 3
    boards, score = np.zeros(E, np.uint64), np.zeros(E, np.int32)
    aux = np.full(E, oracle.AUX_INIT, np.uint64)
    reset_fn(boards, score, None, np.ascontiguousarray(g["start_draws"]))
    resets = g["resets"]
    by_step = {}
    for row in resets:
        by_step.setdefault(int(row[1]), []).append(row)
    h = np.zeros(E, np.uint64)
    with np.errstate(over="ignore"):
        for t in range(T):
            sp, qk = g["spawn"][:, t], g["quirk"][:, t]
            draws = np.full((E, 4), 255, np.uint8)
            has, hq = sp != 255, qk != 255
            draws[has, 0], draws[has, 1] = sp[has] & 15, sp[has] >> 4
            draws[hq, 2], draws[hq, 3] = qk[hq] & 15, qk[hq] >> 4
            reward, flags, maxlvl, ms = step_fn(boards, aux, score, np.ascontiguousarray(actions[:, t]), draws, flavour)
            words = (flags & 7).astype(np.uint64) | (maxlvl.astype(np.uint64) << np.uint64(8)) | \
                    (ms.astype(np.int64).astype(np.uint64) << np.uint64(16))
            aux_word = aux if flavour == oracle.FLAVOUR_PENALTY else np.zeros(E, np.uint64)
            for w in (boards, reward.view(np.uint64), words, score.astype(np.int64).astype(np.uint64), aux_word):
                h = h * FNV + w
            done = ((flags >> 2) & 1).astype(bool)
            rows = by_step.get(t, [])
            assert sorted(int(r[0]) for r in rows) == np.nonzero(done)[0].tolist(), f"done envs differ at step {t}"
            if rows:
                mask = done.astype(np.uint8)
                rd = np.zeros((E, 4), np.uint8)
                for r in rows:
                    rd[int(r[0])] = r[2:6]
                reset_fn(boards, score, mask, rd)
    assert np.array_equal(h, g["digest"]), f"{(h != g['digest']).sum()} of {E} trajectories differ"
    assert np.array_equal(boards, g["final_board"]) and np.array_equal(score, g["final_score"])
    if flavour == oracle.FLAVOUR_PENALTY:
        assert np.array_equal(aux, g["final_aux"])
    return E * T, len(resets)


def oracle_reset(boards, score, mask, draws):
    oracle.env_reset(boards, score, mask, draws)


@pytest.mark.parametrize("name,flavour", [("config2_penalty", oracle.FLAVOUR_PENALTY),
                                          ("config2_nopenalty", oracle.FLAVOUR_NOPENALTY)])
def test_config2_full_size_replay_matches_reference(golden, name, flavour):
    steps, resets = replay_config2(golden(name), flavour, oracle_step, oracle_reset)
    assert steps == 4096 * 512 and resets > 4096          # every env finishes several games


@pytest.mark.parametrize("name,flavour", [("env_penalty", oracle.FLAVOUR_PENALTY),
                                          ("env_nopenalty", oracle.FLAVOUR_NOPENALTY)])
def test_env_step_matches_reference(golden, name, flavour):
    g = golden(name)
    replay_env(g, flavour, oracle_step)


def test_golden_covers_the_branches(golden):
    g = golden("env_penalty")
    fl, lv = g["flags"], g["maxlvl"]
    assert ((fl >> 2) & 1).sum() > 100 and ((fl >> 1) & 1).sum() > 100
    assert lv.max() == 15 and (lv >= 9).sum() > 1000                     # the >=512 reward terms
    assert ((g["aux_out"] >> np.uint64(32)) > 100).any()                  # stall termination
    assert (((g["aux_out"] >> np.uint64(16)) & np.uint64(0xFF)) == 25).any()  # saturated stall penalty
    assert (g["reward"] < -10).any() and (g["reward"] == 10).any()
    over_hi = (((fl >> 1) & 1) == 1) & ((fl & 1) == 0) & (lv >= 9) & (lv <= 11)
    assert over_hi.any()                                                  # game over on 512/1024/2048
    n = golden("env_nopenalty")
    quirk = n["draws"][:, :, 2] != 255
    assert quirk.sum() > 100                                              # full-board quirk (App. A.3)
    assert ((n["flags"] & 1) == 0)[quirk].any()                           # ... also on an invalid agent move


@pytest.mark.parametrize("name", ["env_penalty", "env_nopenalty"])
def test_reset_matches_reference(golden, name):
    g = golden(name)
    boards = np.zeros(len(g["reset_board"]), np.uint64)
    score = np.ones(len(boards), np.int32)
    oracle.env_reset(boards, score, None, np.ascontiguousarray(g["reset_draws"]))
    assert np.array_equal(boards, g["reset_board"])
    assert not score.any()


def test_stall_penalty_table():
    lib = oracle.load()
    p, ref = -1, []
    for _ in range(40):  # Game2048_env.py:124-125
        p = max(p * 1.1, -10)
        ref.append(float(p))
    got = [lib.orc_stall_penalty(k) for k in range(1, 41)]
    assert got == ref
    assert got[oracle.PEN_SAT - 2] > -10 and got[oracle.PEN_SAT - 1] == -10


def test_tabular_agent_matches_reference(golden):
    g = golden("qlearn_ref")
    episodes, lr, gamma, eps0, eps_min = g["params"]
    tab = oracle.QTable(1 << 17, f32=False)
    actions = tab.replay_agent_f64(g["s"], g["explore"], g["rand_action"], g["r"], g["s2"], g["done"], lr, gamma)
    assert np.array_equal(actions, g["a"])           # greedy choices (first-max argmax) follow the reference
    keys, rows = tab.export()
    assert np.array_equal(keys, g["q_keys"])         # same set of states (defaultdict inserts on read)
    assert np.array_equal(rows.view(np.uint64), g["q_rows"].view(np.uint64))  # float64 bit-exact

    tab2 = oracle.QTable(1 << 17, f32=False)
    tab2.update_seq_f64(g["s"], g["a"], g["r"], g["s2"], g["done"], lr, gamma)
    assert np.array_equal(tab2.export()[1].view(np.uint64), g["q_rows"].view(np.uint64))


def test_epsilon_schedule_matches_reference(golden):
    g = golden("qlearn_ref")
    episodes, lr, gamma, eps0, eps_min = g["params"]
    sched = oracle.decay_exploration_schedule(int(episodes), eps0, eps_min)
    assert np.array_equal(sched.view(np.uint64), g["eps"].view(np.uint64))


def test_batch_update_n1_is_sequential_update(golden):
    """Batched synchronous float32 update with N = 1 per call == update_q_value in float32."""
    g = golden("qlearn_ref")
    n = 4000
    lr, gamma = np.float32(0.1), np.float32(0.99)
    tab = oracle.QTable(1 << 15, f32=True)
    ref = {}
    r32 = g["r"].astype(np.float32)
    for i in range(n):
        s, a, s2, d = int(g["s"][i]), int(g["a"][i]), int(g["s2"][i]), int(g["done"][i])
        tab.update_batch_f32(g["s"][i:i + 1], g["a"][i:i + 1], r32[i:i + 1], g["s2"][i:i + 1], g["done"][i:i + 1],
                             float(lr), float(gamma))
        q2 = ref.setdefault(s2, np.zeros(4, np.float32))
        q = ref.setdefault(s, np.zeros(4, np.float32))
        target = r32[i] + (np.float32(0) if d else gamma * q2.max())
        q[a] = q[a] + lr * (target - q[a])
    keys, rows = tab.export()
    assert len(keys) == len(ref)
    for k, row in zip(keys, rows):
        assert np.array_equal(row.astype(np.float32), ref[int(k)])
    # and float32 stays within the stated tolerance of the reference's float64 table
    tab64 = oracle.QTable(1 << 15, f32=False)
    tab64.update_seq_f64(g["s"][:n], g["a"][:n], g["r"][:n], g["s2"][:n], g["done"][:n], 0.1, 0.99)
    k64, r64 = tab64.export()
    assert np.array_equal(k64, keys)
    np.testing.assert_allclose(rows, r64, rtol=1e-5, atol=1e-5)


def test_dqn_encode_state_and_epsilon_follow_the_reference_agent(golden):
    """Golden recorded by running the reference's own DQNAgent.encode_state / update_epsilon source
    (Dqn8TestNOPERCNN.py:271-277, :341-343; oracle/make_golden.py:record_dqn, TF ops stubbed with numpy)."""
    g = golden("dqn_agent")
    boards, bad = oracle.pack_i64(g["tiles"])
    assert bad == 0
    assert np.array_equal(oracle.encode_onehot(boards), g["onehot"])
    assert g["onehot"].sum() == 16 * len(boards) and g["onehot"][:, 15].sum() > 0        # 32768 tiles are in the sample

    from g2048 import dqn

    class Stand:
        epsilon_start, epsilon_min, epsilon_decay, epsilon, step_counter = 0.9, 0.001, 0.9999, 0.9, 0
    me = Stand()
    for step, want in zip(g["eps_steps"], g["eps"]):
        me.step_counter = int(step)
        dqn.BatchedDQNAgent.update_epsilon(me)
        assert me.epsilon == want                      # float64, same expression: bit-equal
