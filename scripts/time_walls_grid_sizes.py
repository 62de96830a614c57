"""K1c (`fgx_raycast_walls_grid`) against the agent count on ONE GPU: the C4 arena and walls, the first n agents of
the C4 workload, n = 125 k ... 1 M (125 k is one rank's share at 8 GPUs).  Two timings per size: back to back
(wall grid L2-resident) and with an L2 flush (a 256 MB memset) before every launch, which is what the grid tables
see in the tick when the step kernel's 1.4 GB per rank has just streamed through.  Prints one JSON line."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for n in (125_000, 250_000, 500_000, 1_000_000):
    idx = slice(0, None, 1_000_000 // n)  # strided like a round-robin shard
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a[idx])).to(dev)  # noqa: E731
    x, y, ang, act = (up(work[k]) for k in ("x", "y", "ang", "active"))
    d = torch.empty((x.shape[0], bench.N_RAYS), dtype=torch.float32, device=dev)
    h = torch.empty((x.shape[0], bench.N_RAYS), dtype=torch.int32, device=dev)
    run = lambda: sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, space.walls, active_state=act,  # noqa: E731
                                   n_rays=bench.N_RAYS, out=(d, h), bounds=bounds)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 20
    cold = []
    for _ in range(10):
        flush.zero_()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1))
    import hashlib
    dig = hashlib.sha256(d.cpu().numpy().tobytes() + h.cpu().numpy().tobytes()).hexdigest()[:16]
    out[str(x.shape[0])] = {"digest": dig, "ms_back_to_back": warm, "ms_after_l2_flush": float(np.median(cold)),
                           "ns_per_agent_back_to_back": warm * 1e6 / x.shape[0]}
print(json.dumps({"kernel": "raycast_walls_grid_kernel", "env": {k: v for k, v in os.environ.items() if k.startswith("FGX_")},
                  "by_agents": out}))
