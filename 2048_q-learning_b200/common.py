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


def schedule_phases(total_epochs, eps0, eps_min):
    """The four phases of the reference's per-episode epsilon schedule (main.py:19-32, :45-57) as rows
    (first epoch NOT in the phase, floor, decrement): linear to 1.5 eps_min over the first 30 % of the epochs, on to
    1.1 eps_min until 60 %, to eps_min until 80 %, eps_min afterwards.  The decrements are built in float64 from the same
    operands in the same order as the reference's, so the schedule is bit-equal (tests/test_host_tables.py)."""
    t1, t2, t3 = total_epochs * 0.30, total_epochs * 0.60, total_epochs * 0.80
    return ((t1, eps_min * 1.5, (eps0 - (eps_min * 1.5)) / t1),
            (t2, eps_min * 1.1, ((eps0 - eps_min) - (eps_min * 1.5)) / (t2 - t1)),
            (t3, eps_min, (eps_min * 1.1 - eps_min) / (t3 - t2)),
            (float("inf"), eps_min, float("inf")))


def epsilon_schedule_step(agent, current_epoch: int) -> float:
    """decay_exploration(current_epoch): one row of the agent's phase table."""
    for limit, floor, dec in agent.schedule:
        if current_epoch < limit:
            agent.epsilon = max(floor, agent.epsilon - dec)
            break
    return agent.epsilon


def init_schedule(agent, total_epochs, exploration_rate, exploration_min):
    """Sets epsilon and the phase table; the reference's attribute names stay readable on the agent (duck-typed API)."""
    agent.epsilon, agent.epsilon_min, agent.total_epochs = exploration_rate, exploration_min, total_epochs
    agent.schedule = schedule_phases(total_epochs, exploration_rate, exploration_min)
    (agent.first_decay_limit, _, agent.slow_decay_1), (agent.second_decay_limit, _, agent.fast_decay), \
        (agent.third_decay_limit, _, agent.slow_decay_2) = agent.schedule[:3]
    agent.epsilon_decay_linear = (exploration_rate - exploration_min) / (total_epochs * 0.75)
