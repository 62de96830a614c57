import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
import foragax_b200.base.space_methods as sm
from foragax_b200.base.space_classes import Point, Wall
dev = torch.device("cuda", 0)
n = bench.N_AGENTS
work = bench.make_workload(n)
up = lambda a: torch.from_numpy(a).to(dev)
x, y, ang, act = (up(work[k]) for k in ("x", "y", "ang", "active"))
w = work["walls"]
walls = Wall(Point(up(w[0]), up(w[1])), Point(up(w[2]), up(w[3])))
out0 = (torch.empty((n, 32), dtype=torch.float32, device=dev), torch.empty((n, 32), dtype=torch.int32, device=dev))
out1 = (torch.empty_like(out0[0]), torch.empty_like(out0[1]))
b = (0.0, bench.ARENA, 0.0, bench.ARENA)
def t(fn, k=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ta = t(lambda: sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out0))
tg = t(lambda: sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out1, bounds=b), 20)
print(f"all-pairs {ta:.3f} ms, grid {tg:.3f} ms, equal dist {torch.equal(out0[0], out1[0])} hit {torch.equal(out0[1], out1[1])}")
