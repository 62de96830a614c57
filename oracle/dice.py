"""NumPy restatement of the hello-world Dice agent (``README.md:35-129`` =
``examples/basic/hello_world/hello_world.ipynb``), the workload of BASELINE
config 1 (and, with a 20-sided die, of the churn config 5).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``sides`` generalises the
reference's ``randint(key, (1,), 1, 7)`` to ``randint(key, (1,), 1, sides+1)``.
"""
import numpy as np

from . import agent_methods as am
from . import jax_random as jr


def create_agent(params, unique_id, active_state, agent_type, keys, sides=6):
    """Batched ``Dice.create_agent`` (README.md:37-53)."""
    sub = jr.split(keys, 2)[:, 1, :]
    draw_active = jr.randint(sub, 1, 1, sides + 1)
    draw = np.where(active_state[:, None] != 0, draw_active, np.int32(0)).astype(np.int32)
    n = len(unique_id)
    return {
        "state": {"draw": draw, "key": sub.astype(np.uint32)},
        "unique_id": unique_id.astype(np.int32),
        "agent_type": agent_type.astype(np.int32),
        "active_state": active_state.astype(np.float32),
        "age": np.zeros(n, np.float32),
    }


def step_agent(params, inp, agents, sides=6):
    """Batched ``Dice.step_agent`` (README.md:55-71): inactive slots pass through."""
    act = agents["active_state"] != 0
    sub = jr.split(agents["state"]["key"], 2)[:, 1, :]
    draw = jr.randint(sub, 1, 1, sides + 1)
    return {
        "state": {
            "draw": np.where(act[:, None], draw, agents["state"]["draw"]).astype(np.int32),
            "key": np.where(act[:, None], sub, agents["state"]["key"]).astype(np.uint32),
        },
        "unique_id": agents["unique_id"],
        "agent_type": agents["agent_type"],
        "active_state": agents["active_state"],
        "age": np.where(act, agents["age"] + np.float32(1.0), agents["age"]).astype(np.float32),
    }


def add_agent(params, agents, idx, key, sides=6):
    """``Dice.add_agent`` (README.md:73-81): one row."""
    row = am.tree_row(agents, idx)
    sub = jr.split(row["state"]["key"], 2)[1]
    row["state"] = {"draw": jr.randint(sub, 1, 1, sides + 1), "key": sub}
    row["active_state"] = True
    return row, key


def remove_agent(params, agents, idx):
    """``Dice.remove_agent`` (README.md:83-90): one row."""
    row = am.tree_row(agents, idx)
    row["state"] = {"draw": np.array([0], np.int32), "key": row["state"]["key"]}
    row["active_state"] = False
    return row


def is_value(v):
    def f(agents, select_params):
        return agents["state"]["draw"].reshape(-1) == v
    return f


def churn_tick(agents, num_total, sides, kill=1, birth=None):
    """One hello-world tick (README.md:98-129): step -> select(kill) -> remove ->
    sort(active desc) -> select(birth) -> clip -> add."""
    birth = sides if birth is None else birth
    import functools
    agents = step_agent(None, None, agents, sides)
    n_dead, rem = am.select_agents(is_value(kill), None, agents)
    agents = am.remove_agents(remove_agent, n_dead, {"remove_ids": rem}, agents)
    agents, _ = am.sort_agents(agents["active_state"], False, agents)
    n_add, _ = am.select_agents(is_value(birth), None, agents)
    n_act = np.sum(agents["active_state"], dtype=np.float32).astype(np.int32)
    n_add = min(int(n_add), num_total - int(n_act))
    agents, _ = am.add_agents(functools.partial(add_agent, sides=sides), n_add, None, agents, None)
    return agents
