"""Config C2: one Foraging_world tick at the constants of EcoEvoJax/sim.py (1,000 forager /
6,000 resource slots, 9 rays, 4 walls, 40 neurons), PRNGKey(1): eager native calls vs the whole tick
replayed as ONE CUDA graph, next to the NumPy oracle on the host.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foragax_b200 import random as fgx_random  # noqa: E402
from foragax_b200.examples import foraging as fg  # noqa: E402
from oracle import foraging as ofo  # noqa: E402

TICKS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
fg.use_variant("sim")
cfg = ofo.Config("sim")


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


w = fg.Foraging_world(fg.default_params("sim"), fgx_random.PRNGKey(cfg.seed), TICKS)
for _ in range(5):
    w.step()
eager_ms = timed(w.step, TICKS)
w2 = fg.Foraging_world(fg.default_params("sim"), fgx_random.PRNGKey(cfg.seed), TICKS)
g = w2.capture(warmup=5)
graph_ms = timed(g.replay, TICKS)
nodes = None
ow = ofo.World(cfg)
t = time.perf_counter()
for _ in range(10):
    ow.step()
cpu_ms = (time.perf_counter() - t) / 10 * 1e3
slots = cfg.for_total + cfg.res_total
print(json.dumps({"config": "C2 foraging world tick (sim.py constants)", "ticks": TICKS, "eager_ms_per_tick": eager_ms,
                  "graph_ms_per_tick": graph_ms, "numpy_oracle_ms_per_tick": cpu_ms,
                  "agent_steps_per_sec_graph": slots / (graph_ms * 1e-3),
                  "active_foragers": int(g.outputs[0]), "active_resources": int(g.outputs[1])}))
