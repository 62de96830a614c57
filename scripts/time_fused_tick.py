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
work = bench.make_workload()
fset, space = bench.build_foragers(work, 0, 1, dev)
n = bench.N_AGENTS
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


print(f"two kernels {t(two):.3f} ms per tick, fused {t(fused):.3f} ms per tick")
