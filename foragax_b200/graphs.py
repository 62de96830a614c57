"""CUDA-graph replay of a whole simulation tick.

The reference compiles each ``jit_*`` call with XLA and still pays one Python dispatch per call
(``EcoEvoJax/sim.py:651-666``: five dispatches and a host round-trip per tick).  The native ops are all
stream-asynchronous with device-resident counts, so an entire tick -- interaction, sensing, steps,
select / remove / sort / add / set -- can be captured ONCE into a CUDA graph and replayed with a single
launch; small worlds (config C2: 1,000 + 6,000 slots) are launch-bound, not bandwidth-bound.

A tick maps a state pytree to a new state pytree (``sort_agents`` gathers into fresh buffers).  For
replay the state must live at fixed addresses, so the captured body ends by copying every leaf of the new
state back into the static buffers it started from.
"""
import torch

from . import struct


def _copy_back(static, new):
    def cp(dst, src):
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src.reshape(dst.shape))
        return dst
    return struct.tree_map(cp, static, new)


class GraphedTick:
    """``tick_fn(state) -> (new_state, outputs)`` captured into a CUDA graph.

    ``state`` is a pytree (dataclasses / dicts / lists) of CUDA tensors; its leaves become the static
    buffers.  ``outputs`` (a pytree of tensors, e.g. per-tick counters) are static too: read them after
    ``replay()`` (they are overwritten by the next replay)."""

    def __init__(self, tick_fn, state, warmup=2):
        self.state = state
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up ticks (real ticks: they advance the state)
            for _ in range(warmup):
                new_state, _ = tick_fn(self.state)
                _copy_back(self.state, new_state)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            new_state, self.outputs = tick_fn(self.state)
            _copy_back(self.state, new_state)
        self.ticks_done = warmup  # capturing does not execute: the first replay is tick number warmup

    def replay(self):
        self.graph.replay()
        return self.outputs
