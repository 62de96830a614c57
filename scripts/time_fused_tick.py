"""C4 tick, one B200: culled ray cast + forager step as two kernels vs fused into one (fgx_forager_sense_step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import foragax_b200.base.space_methods as sm  # noqa: E402
from foragax_b200.base.agent_classes import Params, Signal  # noqa: E402
from foragax_b200.base.agent_methods import step_agents  # noqa: E402
from foragax_b200.examples.foraging import Forager  # noqa: E402

dev = torch.device("cuda", 0)
n = int(os.environ.get("N", bench.N_AGENTS))
work = bench.make_workload(n)
fset, space = bench.build_foragers(work, 0, 1, dev, n_agents=n)
st = fset.agents.state.content
out = (torch.empty((n, bench.N_RAYS), device=dev), torch.empty((n, bench.N_RAYS), dtype=torch.int32, device=dev))
energy_in = torch.zeros(n, device=dev)
params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0, 'energy_max': 15.0})
bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)


def two():
    sm.raycast_walls((st["X"], st["Y"], st["ang"]), bench.RAY_SPAN, bench.RAY_LEN, space.walls,
                     active_state=fset.agents.active_state, n_rays=bench.N_RAYS, out=out, bounds=bounds)
    fset.agents = step_agents(params, Signal(content={'obs': out[0], 'energy_in': energy_in}), fset)


def fused():
    Forager.sense_step(params, energy_in, fset.agents, space.walls, bounds, bench.RAY_SPAN, bench.RAY_LEN, out=out)


from foragax_b200 import struct  # noqa: E402
from foragax_b200.base.agent_classes import Agent_Set  # noqa: E402

CHUNKS = int(os.environ.get("CHUNKS", "8"))
side = torch.cuda.Stream()
bounds_idx = [(i * n // CHUNKS, (i + 1) * n // CHUNKS) for i in range(CHUNKS)]
views = []
for lo, hi in bounds_idx:   # per-chunk views of every leaf (in-place ops on them update the full set)
    sub = struct.tree_map(lambda x: x[lo:hi] if x.dim() and x.shape[0] == n else x, fset.agents)
    vs = Agent_Set(agent=Forager, num_total_agents=hi - lo, num_active_agents=0, agent_type=1)
    vs.agents = sub
    views.append((vs, (out[0][lo:hi], out[1][lo:hi]), energy_in[lo:hi]))
evs = [torch.cuda.Event() for _ in range(CHUNKS)]


def pipelined():
    """ray cast of chunk i+1 (issue-bound, no HBM traffic) runs on a side stream while the step of chunk i
    (HBM-bound, idle issue slots) runs on the main stream"""
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        for i, (vs, o, e) in enumerate(views):
            c = vs.agents.state.content
            sm.raycast_walls((c["X"], c["Y"], c["ang"]), bench.RAY_SPAN, bench.RAY_LEN, space.walls,
                             active_state=vs.agents.active_state, n_rays=bench.N_RAYS, out=o, bounds=bounds)
            evs[i].record(side)
    for i, (vs, o, e) in enumerate(views):
        main.wait_event(evs[i])
        vs.agents = step_agents(params, Signal(content={'obs': o[0], 'energy_in': e}), vs)


def t(fn, k=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


print(f"n = {n}: two kernels {t(two):.3f} ms per tick, fused {t(fused):.3f} ms per tick, "
      f"pipelined over {CHUNKS} chunks on two streams {t(pipelined):.3f} ms per tick")
