"""Where the reference checkout is present (the build container; never the GPU box): run the UNMODIFIED reference
envs live on trajectories that are NOT in the committed goldens (other seeds), record their draws with
oracle/make_golden.py's recorder and replay them through the oracle -- boards, flags, scores, aux state and the
float64 reward bits must be identical.  Skipped when /root/reference does not exist."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import oracle
from test_oracle_golden import oracle_step, replay_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


def recorder(offset, n_envs, n_steps):
    spec = importlib.util.spec_from_file_location("make_golden_live", os.path.join(ROOT, "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.SEED_OFFSET, mg.N_ENVS, mg.N_STEPS = offset, n_envs, n_steps
    return mg


@pytest.mark.timeout(280)
@pytest.mark.parametrize("flavour,code,offset", [("penalty", oracle.FLAVOUR_PENALTY, 50_000),
                                                 ("nopenalty", oracle.FLAVOUR_NOPENALTY, 60_000)])
def test_oracle_replays_fresh_reference_trajectories(flavour, code, offset):
    g = recorder(offset, 64, 320).record_env(flavour)
    replay_env(g, code, oracle_step)
    fl = g["flags"]
    assert fl.size == 64 * 320 and (fl & 1).mean() > 0.3 and ((fl >> 2) & 1).sum() > 0      # valid moves and finished games


@pytest.mark.timeout(280)
def test_oracle_replays_a_fresh_reference_training_run():
    """The reference's QLearningAgent + penalty env in the loop of main.py:80-109 under seeds other than the committed
    golden's: the oracle's float64 agent must take the same greedy actions and end with the same table, bit for bit."""
    g = recorder(777, 1, 1).record_qlearn(episodes=40)
    episodes, lr, gamma, eps0, eps_min = g["params"]
    tab = oracle.QTable(1 << 17, f32=False)
    actions = tab.replay_agent_f64(g["s"], g["explore"], g["rand_action"], g["r"], g["s2"], g["done"], lr, gamma)
    assert np.array_equal(actions, g["a"]) and len(g["s"]) > 3000
    keys, rows = tab.export()
    assert np.array_equal(keys, g["q_keys"])
    assert np.array_equal(rows.view(np.uint64), g["q_rows"].view(np.uint64))
    sched = oracle.decay_exploration_schedule(int(episodes), eps0, eps_min)
    assert np.array_equal(sched.view(np.uint64), g["eps"].view(np.uint64))


@pytest.mark.timeout(280)
@pytest.mark.parametrize("flavour", ["penalty", "nopenalty"])
def test_committed_goldens_are_what_the_recorder_produces(flavour, golden):
    """tests/golden/env_*.npz == a fresh run of oracle/make_golden.py:record_env against the reference checkout."""
    want = golden(f"env_{flavour}")
    mg = recorder(0, 96, 384)
    got = mg.record_env(flavour)
    assert set(got) == set(want)
    for k in want:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a, b.view(np.uint8) if b.dtype.kind == "f" else b), k
