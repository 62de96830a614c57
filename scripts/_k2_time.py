import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from foragax_b200.base.agent_classes import Params, Signal
from foragax_b200.base.agent_methods import step_agents
dev = torch.device("cuda", 0)
work = bench.make_workload()
fset, space = bench.build_foragers(work, 0, 1, dev)
n = bench.N_AGENTS
obs = torch.rand((n, bench.N_RAYS), device=dev) * 60
energy_in = torch.zeros(n, device=dev)
params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0, 'energy_max': 15.0})
def run(): fset.agents = step_agents(params, Signal(content={'obs': obs, 'energy_in': energy_in}), fset)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print("K2 alone", sys.argv[1:], e0.elapsed_time(e1) / 20, "ms")
