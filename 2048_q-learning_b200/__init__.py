"""2048_q-learning_b200 -- the B200-native hot path of Rocco9999/2048_Q-Learning.

Batched 2048 env reset()/step(), epsilon-greedy choose_action and the tabular Q-update as hand-written
sm_100a CUDA kernels behind a C ABI (include/g2048.h, libg2048.so), plus the host-side mirror of the
reference's Python interfaces.  `import g2048` (repo root) is an importable alias of this package.
"""
from ._lib import G2048Error, build, declared_symbols, init, lib  # noqa: F401
from .compat import Game2048, Game2048_env, QLearningAgent, pack_tiles, unpack_tiles  # noqa: F401
from .train import (evaluate_dqn, evaluate_random, evaluate_tabular, train_dqn, train_tabular,  # noqa: F401
                    train_tabular_batched)


def __getattr__(name):  # the batched classes need torch: import lazily
    if name in ("BatchedGame2048Env", "boards_to_numpy", "boards_from_numpy"):
        from . import env
        return getattr(env, name)
    if name == "BatchedQLearningAgent":
        from . import agent
        return agent.BatchedQLearningAgent
    if name in ("BatchedDQNAgent", "DQNModel", "dqn_step", "terminal_bonus", "FusedDQNFeed"):
        from . import dqn
        return getattr(dqn, name)
    if name in ("ShardedQLearning", "shard_range", "PeerRecordBuffers", "TorchEngine", "GradientAllReduce", "SharedQTable", "OwnerComputesQLearning"):
        from . import dist
        return getattr(dist, name)
    raise AttributeError(name)
