"""Load the UNMODIFIED reference (Rocco9999/2048_Q-Learning) for golden-vector generation.

TEST INFRASTRUCTURE.  Works only where the reference checkout exists (the build
container: /root/reference, or $G2048_REF_ROOT); it is never imported by the
product, by `-m gpu` tests, by smoke() or by bench.py.

The reference imports `gymnasium` (for gym.Env / spaces.Discrete / spaces.Box,
QLearningBase/environment/Game2048_env.py:1-3,78,89-90) and `matplotlib`
(dead import, QLearningBase/Agent/main.py:9); neither is installed here, so two
minimal stub modules are injected into sys.modules before the import.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("G2048_REF_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "QLearningBase/environment/Game2048_env.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class Env:  # gym.Env: only used as a base class
            pass

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        gym.Env = Env
        spaces.Discrete = Discrete
        spaces.Box = Box
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def _load(name: str, relpath: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_penalty_env():
    """QLearningBase/environment/Game2048_env.py -> module (Game2048, Game2048_env)."""
    _install_stubs()
    return _load("ref_penalty_env", "QLearningBase/environment/Game2048_env.py")


def load_nopenalty_env():
    """Deep_QLearning/environment/Game2048_nopenalty_env.py -> module."""
    _install_stubs()
    return _load("ref_nopenalty_env", "Deep_QLearning/environment/Game2048_nopenalty_env.py")


def load_tabular_agent():
    """QLearningBase/Agent/main.py -> module (QLearningAgent); its training loop is __main__-guarded."""
    _install_stubs()
    # main.py does `from environment.Game2048_env import Game2048_env` after appending its parent to sys.path
    return _load("ref_tabular_main", "QLearningBase/Agent/main.py")
