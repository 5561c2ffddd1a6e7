"""Load the UNMODIFIED reference (Rocco9999/2048_Q-Learning) for golden-vector generation.

TEST INFRASTRUCTURE.  Works only where the reference checkout exists (the build
container: /root/reference, or $G2048_REF_ROOT); it is never imported by the
product, by `-m gpu` tests, by smoke() or by bench.py.

The DQN agent (Deep_QLearning/main_dir/Dqn8TestNOPERCNN.py) imports tensorflow, keras and keras-rl, none of which
is installable here.  `load_dqn_agent()` injects stubs for them: the three TensorFlow ops the action-selection path uses
(`tf.one_hot`, `tf.reshape`, `tf.transpose`, Dqn8TestNOPERCNN.py:271-277) are restated with numpy according to their
documented semantics; the Keras / keras-rl names are placeholders that are never called (the network and the replay
memory are not constructed).  Goldens recorded that way pin `encode_state`, `act`, `act_ripetitive` and `update_epsilon`
as the reference's own code executes them -- relative to those three numpy restatements.

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


def _install_dqn_stubs() -> None:
    import numpy as np

    if "tensorflow" not in sys.modules:
        tf = types.ModuleType("tensorflow")
        tf.one_hot = lambda indices, depth: np.eye(depth, dtype=np.float32)[np.asarray(indices, dtype=np.int64)]
        tf.reshape = lambda x, shape: np.reshape(np.asarray(x), shape)
        tf.transpose = lambda x, perm=None: np.transpose(np.asarray(x), perm)
        sys.modules["tensorflow"] = tf

    class _Placeholder:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            raise RuntimeError("Keras / keras-rl are stubbed: only the action-selection path of the DQN agent can run")

    def module(name, names):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        for n in names:
            setattr(m, n, type(n, (_Placeholder,), {}))
        sys.modules[name] = m

    module("keras", [])
    module("keras.models", ["Sequential", "Model"])
    module("keras.layers", ["Dense", "Flatten", "Input", "Dropout", "Conv2D", "ReLU", "Concatenate", "GlobalAveragePooling2D"])
    module("keras.optimizers", ["Adam", "SGD"])
    module("keras.callbacks", ["ReduceLROnPlateau", "LearningRateScheduler", "Callback"])
    module("rl", [])
    module("rl.memory", ["SequentialMemory", "RingBuffer"])


def load_dqn_agent():
    """Deep_QLearning/main_dir/Dqn8TestNOPERCNN.py -> module (DQNAgent class; do not construct it: use its methods on a
    stand-in object that carries epsilon, step_counter, action_space and a `model` with predict())."""
    _install_dqn_stubs()
    return _load("ref_dqn_agent", "Deep_QLearning/main_dir/Dqn8TestNOPERCNN.py")
