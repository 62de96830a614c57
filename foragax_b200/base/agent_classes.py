"""Agent containers -- the B200 mirror of ``foragax/base/agent_classes.py:1-83``.

Same names and fields as the reference (``Signal, Params, State, Policy, Agent,
Agent_Set``); every leaf of a *set* of agents is a CUDA tensor whose leading axis is the
slot axis (the struct-of-arrays layout ``jax.vmap`` produces in the reference).

Difference forced by the platform (hard part 2): the reference traces
arbitrary per-agent Python callbacks with ``jax.vmap`` / ``lax.fori_loop``.  Here a
per-agent function is a *native batched op*: a static method that receives the whole
struct-of-arrays set and launches a CUDA kernel through the C ABI.  Such methods are
marked with :func:`native`; anything else is rejected with a clear error -- there is no
interpreter or CPU fallback.
"""
from .. import struct


def native(fn):
    """Mark a batched, kernel-backed per-agent function (create/step/add/remove/set)."""
    fn.fgx_native = True
    return fn


def is_native(fn):
    return bool(getattr(fn, "fgx_native", False))


def require_native(fn, role):
    if fn is None or not is_native(fn):
        name = getattr(fn, "__qualname__", repr(fn))
        raise TypeError(
            f"{role}={name} is not a native batched op: foragax_b200 runs per-agent functions as CUDA kernels "
            "(see foragax_b200.base.agent_classes.native); arbitrary per-agent Python callbacks are not "
            "traced and there is no CPU fallback")
    return fn


@struct.dataclass
class Signal:
    content: dict


@struct.dataclass
class Params:
    content: dict


@struct.dataclass
class State:
    content: dict


@struct.dataclass
class Policy:
    """agent_classes.py:17-31"""
    state: State
    params: Params
    deterministic: bool

    @staticmethod
    def create_policy(params, key):
        pass

    @staticmethod
    def step_policy(input, policy):
        pass

    @staticmethod
    def reset_policy(input, policy):
        pass


@struct.dataclass
class Agent:
    """agent_classes.py:33-65.  ``active_state`` is float32 1.0/0.0 (agent_methods.py:14)."""
    params: Params
    state: State
    unique_id: object
    agent_type: object
    active_state: object
    age: object
    policy: Policy

    @staticmethod
    def create_agent(create_params, unique_id, active_state, agent_type, key):
        pass

    @staticmethod
    def step_agent(step_params, input, agents):
        pass

    @staticmethod
    def add_agent(add_params, agents, idx, key):
        pass

    @staticmethod
    def set_agent(set_params, agents, idx, key):
        pass

    @staticmethod
    def reset_agent():
        pass

    @staticmethod
    def remove_agent(remove_params, agents, idx):
        pass


class Agent_Set:
    """agent_classes.py:67-83.  The reference builds ``vmap(create_agent)`` and
    ``jit(vmap(step_agent))`` here; the native per-agent ops are already batched, so they
    are bound as they are."""
    agents: Agent

    def __init__(self, agent, num_total_agents, num_active_agents, agent_type):
        self.create_agents = require_native(agent.create_agent, "create_agent")
        self.step_agents = require_native(agent.step_agent, "step_agent")
        self.reset_agents = agent.reset_agent
        self.num_total_agents = num_total_agents
        self.num_active_agents = num_active_agents
        self.num_inactive_agents = num_total_agents - num_active_agents
        self.agent_type = agent_type
        print("AgentSet initialized")
