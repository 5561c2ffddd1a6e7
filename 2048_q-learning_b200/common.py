"""Constants and the epsilon schedule shared by the batched classes and the N = 1 adapters (torch-free)."""
from __future__ import annotations

FLAVOURS = {"penalty": 0, "nopenalty": 1}
AUX_INIT = 0x000000000000FF01
N_COUNTERS = 16
COUNTER_NAMES = ("steps", "valid", "episodes", "score", "maxlvl", "reward_fx", "inserts", "dropped", "lost", "retried")
MODES = {"atomic": 0, "deterministic": 1}



class ActionSpace:
    n = 4


class ObservationSpace:
    shape = (4, 4)


def epsilon_schedule_step(agent, current_epoch: int) -> float:
    """decay_exploration (main.py:45-57) on any object with the reference's schedule attributes."""
    if current_epoch < agent.first_decay_limit:
        agent.epsilon = max(agent.epsilon_min * 1.5, agent.epsilon - agent.slow_decay_1)
    elif current_epoch < agent.second_decay_limit:
        agent.epsilon = max(agent.epsilon_min * 1.1, agent.epsilon - agent.fast_decay)
    elif current_epoch < agent.third_decay_limit:
        agent.epsilon = max(agent.epsilon_min, agent.epsilon - agent.slow_decay_2)
    else:
        agent.epsilon = agent.epsilon_min
    return agent.epsilon


def init_schedule(agent, total_epochs, exploration_rate, exploration_min):
    """The schedule constants of QLearningAgent.__init__ (main.py:19-32)."""
    agent.epsilon, agent.epsilon_min, agent.total_epochs = exploration_rate, exploration_min, total_epochs
    agent.epsilon_decay_linear = (exploration_rate - exploration_min) / (total_epochs * 0.75)
    agent.first_decay_limit = total_epochs * 0.30
    agent.second_decay_limit = total_epochs * 0.60
    agent.third_decay_limit = total_epochs * 0.80
    agent.slow_decay_1 = (exploration_rate - (exploration_min * 1.5)) / agent.first_decay_limit
    agent.fast_decay = ((exploration_rate - exploration_min) - (exploration_min * 1.5)) / (
        agent.second_decay_limit - agent.first_decay_limit)
    agent.slow_decay_2 = (exploration_min * 1.1 - exploration_min) / (agent.third_decay_limit - agent.second_decay_limit)
