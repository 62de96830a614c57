"""Config C5: add/remove/sort churn stress on a 10M-slot Dice set with a D20 die
(draw==1 dies, draw==20 spawns: ~5 % + 5 % turnover per tick).  Tick = step -> select -> remove ->
sort(active desc) -> select -> clip -> add, all on-stream with device-resident counts.

Prints per-op CUDA-event times and the achieved HBM bandwidth against each op's ALGORITHMIC bytes
(BASELINE.md section 3), then one JSON line.  `--check` compares the final state with the oracle
at a reduced size."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foragax_b200.base.agent_classes as fgx_classes  # noqa: E402
import foragax_b200.base.agent_methods as fgx_methods  # noqa: E402
from foragax_b200 import random as fgx_random  # noqa: E402
from foragax_b200.examples.hello_world import is_value, make_dice  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=10_000_000)
ap.add_argument("--ticks", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

N, SIDES = args.slots, 20
D = make_dice(SIDES)
s = fgx_classes.Agent_Set(agent=D, num_total_agents=N, num_active_agents=N // 2, agent_type=0)
s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(0))
ops = ["step", "select_dead", "remove", "sort(argsort)", "sort(gather)", "select_birth", "count+clip", "add"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
acc = np.zeros(len(ops))
active_hist = []


def tick(timed):
    e = iter(ev)
    if timed:
        next(e).record()
    s.agents = fgx_methods.step_agents(None, None, s)
    if timed:
        next(e).record()
    n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
    if timed:
        next(e).record()
    s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                             s.agents)
    if timed:
        next(e).record()
    ids = fgx_methods.argsort(s.agents.active_state, False)
    if timed:
        next(e).record()
    s.agents = fgx_methods.take_rows(s.agents, ids)
    if timed:
        next(e).record()
    n_add, _ = fgx_methods.jit_select_agents(is_value(SIDES), None, s.agents)
    if timed:
        next(e).record()
    n_act = fgx_methods.count_active(s.agents)
    n_add = torch.minimum(n_add, N - n_act)
    if timed:
        next(e).record()
    s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, n_add, None, s.agents, None)
    if timed:
        next(e).record()
        torch.cuda.synchronize()
        for k in range(len(ops)):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
        active_hist.append(int(n_act))


for _ in range(args.warmup):
    tick(False)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.ticks):
    tick(False)
t1.record()
torch.cuda.synchronize()
tick_ms = t0.elapsed_time(t1) / args.ticks
for _ in range(args.ticks):
    tick(True)
acc /= args.ticks
# the same tick replayed as ONE CUDA graph (no Python / launch overhead between the ~20 kernels)
from foragax_b200.graphs import GraphedTick  # noqa: E402


def graph_tick(state):
    s.agents = state
    tick(False)
    return s.agents, fgx_methods.count_active(s.agents)


g = GraphedTick(graph_tick, s.agents, warmup=1)
for _ in range(3):
    g.replay()
t0.record()
for _ in range(args.ticks):
    g.replay()
t1.record()
torch.cuda.synchronize()
graph_ms = t0.elapsed_time(t1) / args.ticks
n_act = float(np.mean(active_hist))
k = 0.05 * n_act
ROW = 28  # Dice row bytes
alg = {  # algorithmic bytes per op
    "step": 32 * n_act + 4 * (N - n_act),
    "select_dead": 8 * N, "select_birth": 8 * N,
    "remove": (4 + 4 + 4) * k, "add": (4 + 8 + 8 + 4 + 4) * k,
    "sort(argsort)": 8 * N, "sort(gather)": (2 * ROW + 4) * N,
    "count+clip": 4 * N,
}
peak = 6509.1
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
print(f"C5 churn: {N} slots, ~{n_act:.0f} active, D{SIDES}; graph-replayed tick {graph_ms:.3f} ms; eager tick {tick_ms:.3f} ms "
      f"({N / tick_ms / 1e3:.1f} M slot-steps/s, {n_act / tick_ms / 1e3:.1f} M active agent-steps/s)")
rows = {}
for name, ms in zip(ops, acc):
    gbs = alg[name] / (ms * 1e-3) / 1e9
    rows[name] = {"ms": float(ms), "alg_bytes": float(alg[name]), "gbs": gbs, "frac_of_hbm": gbs / peak}
    print(f"  {name:14s} {ms:8.3f} ms   {alg[name] / 1e6:9.1f} MB algorithmic   {gbs:8.1f} GB/s   {gbs / peak:5.2f} of HBM")
print(json.dumps({"config": "C5", "slots": N, "active": n_act, "tick_ms": tick_ms, "graph_tick_ms": graph_ms,
                  "slot_steps_per_sec": N / (tick_ms * 1e-3), "slot_steps_per_sec_graph": N / (graph_ms * 1e-3), "ops": rows, "hbm_peak_gbs": peak}))
