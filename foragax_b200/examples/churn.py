"""BASELINE config 5: the add / remove / sort churn stress -- a large Dice set (hello-world agent,
``README.md:35-129`` of the reference) with a D20 die: ``draw == 1`` dies, ``draw == 20`` spawns, i.e. ~5 % + 5 %
turnover per tick.  One tick is exactly the hello-world chain (``hello_world.ipynb:104-144``):

    step_agents -> jit_select_agents(draw == 1) -> jit_remove_agents -> jit_sort_agents(active_state, descending)
    -> jit_select_agents(draw == sides) -> clip to the free slots -> jit_add_agents

all on-stream with device-resident counts.  Used by ``bench.py`` (leg ``c5_churn``), ``scripts/bench_churn.py``
and the parity tests.
"""
import torch

from ..base import agent_classes as fgx_classes
from ..base import agent_methods as fgx_methods
from .. import random as fgx_random
from .. import struct
from .hello_world import is_value, make_dice

OPS = ("step", "select_dead", "remove", "sort", "select_birth", "count+clip", "add")
ROW_BYTES = 28  # unique_id, agent_type, active_state, age, draw (4 B each) + key (8 B): SURVEY.md Appendix B


def algorithmic_bytes(n_slots, n_active, turnover):
    """Algorithmic HBM bytes of each op of one tick, ``turnover`` = rows removed ~ rows added."""
    return {
        "step": 32 * n_active + 4 * (n_slots - n_active),   # key, age, flag in; draw, key, age out | flag only
        "select_dead": 8 * n_slots,                          # 4 B leaf in, 4 B index out
        "remove": 12 * turnover,                             # id in, draw + flag out
        "sort": (4 + 4 + 2 * ROW_BYTES) * n_slots,           # key in, index out, every leaf in and out
        "select_birth": 8 * n_slots,
        "count+clip": 4 * n_slots,
        "add": 28 * turnover,                                # key in; draw, key, flag out (+ the count pass: see add)
    }


class ChurnWorld:
    def __init__(self, n_slots, sides=20, n_active=None, key=None, shard=None):
        self.n, self.sides = int(n_slots), int(sides)
        self.D = make_dice(self.sides)
        n_active = self.n // 2 if n_active is None else int(n_active)
        self.set = fgx_classes.Agent_Set(agent=self.D, num_total_agents=self.n, num_active_agents=n_active, agent_type=0)
        self.set.agents = fgx_methods.create_agents(None, self.set, fgx_random.PRNGKey(0) if key is None else key,
                                                    shard=shard)
        self._dead, self._birth = is_value(1), is_value(self.sides)
        self._spare = None  # second buffer set of the sort (two_buffers())

    @property
    def agents(self):
        return self.set.agents

    def tick(self, events=None):
        """One tick; ``events`` (optional): len(OPS) + 1 CUDA events recorded around the ops."""
        s, D = self.set, self.D
        rec = (lambda k: events[k].record()) if events is not None else (lambda k: None)
        rec(0)
        s.agents = fgx_methods.step_agents(None, None, s)
        rec(1)
        n_dead, rem = fgx_methods.jit_select_agents(self._dead, None, s.agents)
        rec(2)
        s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                                 s.agents)
        rec(3)
        if self._spare is None:
            s.agents, _ = fgx_methods.jit_sort_agents(s.agents.active_state, False, s.agents)
        else:  # ping-pong: the sort writes the other buffer set, the one it read becomes the spare
            old = s.agents
            s.agents, _ = fgx_methods.jit_sort_agents(old.active_state, False, old, out=self._spare)
            self._spare = old
        rec(4)
        n_add, _ = fgx_methods.jit_select_agents(self._birth, None, s.agents)
        rec(5)
        n_act, n_add = fgx_methods.count_active_clip(s.agents, n_add)  # min(n_add, N - n_active), README.md:124-125
        rec(6)
        s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, n_add, None, s.agents, None, num_active=n_act)
        rec(7)
        return n_act


    def two_buffers(self):
        """Let the sort alternate between two fixed buffer sets (instead of fresh tensors every tick), so that the
        state is back in the set it started from after every second tick."""
        if self._spare is None:
            self._spare = struct.tree_map(torch.empty_like, self.set.agents)

    def capture_two_ticks(self):
        """Two ticks as ONE CUDA graph (13 launches each, two of them cooperative); replaying it advances the world
        by two ticks.  Everything in a tick is on-stream with device-resident counts, so nothing else changes."""
        self.two_buffers()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream (real ticks: they advance the state)
            self.tick()
            self.tick()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.tick()
            self._n_act = self.tick()
        return graph


def time_graph_ticks(world, ticks, warmup=2):
    """Mean tick ms of `ticks` ticks replayed as ticks / 2 launches of the two-tick graph."""
    graph = world.capture_two_ticks()
    for _ in range(warmup):
        graph.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, ticks // 2)
    t0.record()
    for _ in range(reps):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / (2 * reps), graph


def time_ticks(world, ticks, warmup=3):
    """(mean tick ms over `ticks` back-to-back ticks, {op: mean ms} from a second, event-instrumented pass,
    mean active count)."""
    for _ in range(warmup):
        world.tick()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(ticks):
        world.tick()
    t1.record()
    torch.cuda.synchronize()
    tick_ms = t0.elapsed_time(t1) / ticks
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(OPS) + 1)]
    acc = [0.0] * len(OPS)
    act = 0.0
    for _ in range(ticks):
        n_act = world.tick(ev)
        torch.cuda.synchronize()
        for k in range(len(OPS)):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
        act += float(n_act)
    return tick_ms, {name: a / ticks for name, a in zip(OPS, acc)}, act / ticks
