"""One rank's share of the agent-agent sensing leg at G GPUs, emulated on ONE GPU: n / G sensing agents (strided like a
round-robin shard) against all n targets -- the uniform-grid build over the n targets (count, scan, scatter) is the part
every rank repeats.  Prints event times; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import foragax_b200.base.space_methods as sm  # noqa: E402
from foragax_b200.base.space_methods import create_space  # noqa: E402

dev = torch.device("cuda", 0)
work = bench.make_workload()
space = create_space(0.0, bench.ARENA, 0.0, bench.ARENA, False, bench.wall_array(work))
bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)
n = bench.N_AGENTS
tg = torch.from_numpy(np.stack([work["x"], work["y"], work["rad"], work["active"]], 1).astype(np.float32)).to(dev)
out = {}
for G in (1, 2, 4, 8):
    idx = slice(0, None, G)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a[idx])).to(dev)  # noqa: E731
    x, y, ang, act = (up(work[k]) for k in ("x", "y", "ang", "active"))
    d = torch.empty((x.shape[0], bench.N_RAYS), dtype=torch.float32, device=dev)
    h = torch.empty((x.shape[0], bench.N_RAYS), dtype=torch.int32, device=dev)

    def walls():
        sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, space.walls, active_state=act, n_rays=bench.N_RAYS,
                         out=(d, h), bounds=bounds)

    def agents():
        sm.raycast_agents((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, tg, bounds, 15.0, active_state=act,
                          n_rays=bench.N_RAYS, merge_walls=True, out=(d, h))

    def t(fn, k=10):
        for _ in range(2):
            walls()
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = 0.0
        for _ in range(k):
            walls()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            acc += e0.elapsed_time(e1)
        return acc / k
    out[str(G)] = {"sensors": int(x.shape[0]), "raycast_agents_ms": t(agents)}
print(json.dumps({"targets": n, "by_gpus": out}))
