"""Times fgx_raycast_agents (+ merge with walls) on the C4 workload: 1M agents sensing each other."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import foragax_b200.base.space_methods as sm  # noqa: E402
from foragax_b200.base.space_classes import Point, Wall  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_AGENTS
work = bench.make_workload(n)
up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
x, y, ang, act, rad = (up(work[k]) for k in ("x", "y", "ang", "active", "rad"))
w = work["walls"]
walls = Wall(Point(up(w[0]), up(w[1])), Point(up(w[2]), up(w[3])))
out = (torch.empty((n, 32), dtype=torch.float32, device=dev), torch.empty((n, 32), dtype=torch.int32, device=dev))
targets = torch.stack([x, y, rad, act], dim=1).contiguous()
bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)


def run():
    sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out)
    sm.raycast_agents((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, targets, bounds, 10.0, act, 32, merge_walls=True,
                      out=out)


for _ in range(2):
    run()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.synchronize()
tw = ta = 0.0
K = 5
for _ in range(K):
    e[0].record()
    sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out)
    e[1].record()
    sm.raycast_agents((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, targets, bounds, 10.0, act, 32, merge_walls=True,
                      out=out)
    e[2].record()
    torch.cuda.synchronize()
    tw += e[0].elapsed_time(e[1]) / K
    ta += e[1].elapsed_time(e[2]) / K
hit = out[1]
print(f"n={n}: walls {tw:.2f} ms, agents(+grid build, merge) {ta:.2f} ms; rays hitting an agent "
      f"{float(((hit >= 0) & (hit < n)).float().mean()):.3f}, a wall {float((hit >= n).float().mean()):.3f}")
