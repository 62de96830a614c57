"""Times fgx_forager_step (K2) alone on the C4 workload: python scripts/time_forager_step.py
(FGX_FS_PREFETCH=0 disables the L2 prefetch of the next agent's rows, FGX_FS_SLOTS=2 selects double buffering)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from foragax_b200.base.agent_classes import Params, Signal  # noqa: E402
from foragax_b200.base.agent_methods import step_agents  # noqa: E402

dev = torch.device("cuda", 0)
work = bench.make_workload()
fset, space = bench.build_foragers(work, 0, 1, dev)
n = bench.N_AGENTS
obs = torch.rand((n, bench.N_RAYS), device=dev) * 60
energy_in = torch.zeros(n, device=dev)
params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0, 'energy_max': 15.0})
inp = Signal(content={'obs': obs, 'energy_in': energy_in})
for _ in range(3):
    fset.agents = step_agents(params, inp, fset)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
K = 20
for _ in range(K):
    fset.agents = step_agents(params, inp, fset)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
act = int(work["active"].sum())
gbs = (bench.K2_BYTES_ACTIVE * act + 4 * (n - act)) / (ms * 1e-3) / 1e9
print(f"forager_step: {ms:.3f} ms, {gbs:.0f} GB/s algorithmic")
