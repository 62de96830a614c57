"""Config C3: predator-prey churn at 100,000 slots per species (50 % active), a 1000 x 1000 grass
lattice (1,000,000 grass slots), reproduction 0.08 / 0.2 (wolf_sheep_grass/sim.ipynb:570-572), plus the synthetic
sensing stage: every animal casts 9 rays of length 10 at all animals (rad 0.5) through the uniform grid.
Tick = interaction -> 3 steps -> (select, remove, sort) x 2 -> (select, add, set, sort) x 2 [-> sensing].
Prints one JSON line.  python scripts/bench_wolf_sheep.py [ticks] [max_animals]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foragax_b200.base.space_methods as sm  # noqa: E402
from foragax_b200 import random as fgx_random  # noqa: E402
from foragax_b200.examples.wolf_sheep_grass import Ecosystem  # noqa: E402

TICKS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
MAXA = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
XY = 1000
eco = Ecosystem(grass_regrowth_time=30, wolf_reproduction_rate=0.08, wolf_energy=30, init_wolves=MAXA // 2,
                sheep_reproduction_rate=0.2, sheep_energy=5, init_sheeps=MAXA // 2, X_max=XY, Y_max=XY,
                sim_steps=TICKS, key=fgx_random.PRNGKey(0), max_animals=MAXA)


def sense_inputs():
    s, w = eco.sheeps.agents, eco.wolves.agents
    f = lambda a, k: a.state.content[k].reshape(-1).to(torch.float32)  # noqa: E731
    x = torch.cat([f(s, 'X_pos'), f(w, 'X_pos')])
    y = torch.cat([f(s, 'Y_pos'), f(w, 'Y_pos')])
    act = torch.cat([s.active_state, w.active_state])
    ang = (x * 0.618 + y * 0.382) % (2 * math.pi) - math.pi  # any heading: Random_Policy ignores observations
    return x, y, ang, act, torch.full_like(x, 0.5)


_inputs = None


def sense():
    """fgx_raycast_agents only (grid build + sensing); the float pose tensors are prepared once per state."""
    x, y, ang, act, rad = _inputs
    return sm.raycast_agents((x, y, ang), math.pi / 4, 10.0, (x, y, rad, act), (-1.0, float(XY), -1.0, float(XY)), 0.5,
                             act, 9)


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


for _ in range(3):
    eco.step()
tick_ms, _ = timed(eco.step, TICKS)
_inputs = sense_inputs()
per_call = [timed(sense, 1)[0] for _ in range(12)]
print("sensing per call (ms):", [round(v, 2) for v in per_call], file=sys.stderr)
sense_ms, (dist, hit) = timed(sense, 10)
g = eco.capture(warmup=2)
graph_ms, (ns, nw) = timed(g.replay, TICKS)
slots = 2 * MAXA + XY * XY
print(json.dumps({"config": f"C3 wolf-sheep-grass, {MAXA} slots per species, {XY}x{XY} grass", "ticks": TICKS,
                  "eager_ms_per_tick": tick_ms, "graph_ms_per_tick": graph_ms,
                  "agent_steps_per_sec_graph": slots / (graph_ms * 1e-3),
                  "animal_steps_per_sec_graph": 2 * MAXA / (graph_ms * 1e-3), "sensing_ms": sense_ms,
                  "rays_hitting": float((hit >= 0).float().mean()), "active_sheep": int(ns), "active_wolves": int(nw)}))
