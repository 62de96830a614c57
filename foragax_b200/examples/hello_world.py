"""The hello-world Dice agent (reference ``README.md:35-90`` =
``examples/basic/hello_world/hello_world.ipynb``) with native batched ops.

State leaves as in the reference: ``state.content['draw']`` int32[N,1],
``state.content['key']`` uint32[N,2]; ``params`` / ``policy`` are ``None``.
``sides`` (class attribute, default 6) generalises the die for the churn stress
configuration (a D20: ``randint(key, (1,), 1, 21)``).
"""
import torch

from .. import _lib as L
from .. import struct
from ..base.agent_classes import Agent, State, native
from ..base.predicates import P


@struct.dataclass
class Dice(Agent):
    sides = 6

    # README.md:37-53
    @staticmethod
    @native
    def create_agent(params, unique_id, active_state, agent_type, key, sides=None):
        sides = Dice.sides if sides is None else sides
        n = unique_id.shape[0]
        dev = unique_id.device
        draw = torch.empty((n, 1), dtype=torch.int32, device=dev)
        skey = torch.empty((n, 2), dtype=torch.uint32, device=dev)
        age = torch.empty(n, dtype=torch.float32, device=dev)
        L.check(L.lib.fgx_dice_create(n, sides, L.ptr(key), L.ptr(active_state), L.ptr(draw), L.ptr(skey),
                                      L.ptr(age), L.stream()))
        return Dice(params=params, unique_id=unique_id, agent_type=agent_type, active_state=active_state,
                    state=State(content={'draw': draw, 'key': skey}), policy=None, age=age)

    # README.md:55-71 (in place)
    @staticmethod
    @native
    def step_agent(params, input, dice_agent, sides=None):
        sides = Dice.sides if sides is None else sides
        c = dice_agent.state.content
        n = dice_agent.unique_id.shape[0]
        L.check(L.lib.fgx_dice_step(n, sides, L.ptr(dice_agent.active_state), L.ptr(c['draw']), L.ptr(c['key']),
                                    L.ptr(dice_agent.age), L.stream()))
        return dice_agent

    # README.md:73-81 (rows = RowRange; in place)
    @staticmethod
    @native
    def add_agent(params, dice_agents, rows, key, sides=None):
        sides = Dice.sides if sides is None else sides
        c = dice_agents.state.content
        n = dice_agents.unique_id.shape[0]
        L.check(L.lib.fgx_dice_add(n, sides, L.ptr(rows.start), L.ptr(rows.count), L.ptr(c['draw']), L.ptr(c['key']),
                                   L.ptr(dice_agents.active_state), L.stream()))
        return dice_agents, key

    # README.md:83-90 (rows = RowIds; in place)
    @staticmethod
    @native
    def remove_agent(params, dice_agents, rows):
        c = dice_agents.state.content
        n = dice_agents.unique_id.shape[0]
        L.check(L.lib.fgx_dice_remove(n, L.ptr(rows.ids), L.ptr(rows.count), L.ptr(c['draw']),
                                      L.ptr(dice_agents.active_state), L.stream()))
        return dice_agents


def make_dice(sides):
    """A Dice agent class with a different die (config 5 uses a D20)."""
    import functools

    @struct.dataclass
    class DiceN(Dice):
        pass

    DiceN.sides = sides
    for name in ("create_agent", "step_agent", "add_agent"):
        base = getattr(Dice, name)
        fn = functools.partial(base, sides=sides)
        fn.fgx_native = True
        setattr(DiceN, name, staticmethod(fn))
    DiceN.__name__ = f"Dice{sides}"
    return DiceN


def is_value(v):
    """README.md:100-102,117-119: select agents whose draw equals ``v`` (in-kernel predicate)."""
    def select(dice_agent, select_params):
        return P(dice_agent.state.content['draw']) == v
    return select


is_one = is_value(1)
is_six = is_value(6)
