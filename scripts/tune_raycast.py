"""Times fgx_raycast_walls on the C4 workload for each tuning variant (FGX_RC_VARIANT):
rays per thread x minimum CTAs per SM.  CUDA events, 3 warm-ups, 10 timed launches each."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import foragax_b200.base.space_methods as sm  # noqa: E402
from foragax_b200.base.space_classes import Point, Wall  # noqa: E402

dev = torch.device("cuda", 0)
work = bench.make_workload()
up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
x, y, ang, act = (up(work[k]) for k in ("x", "y", "ang", "active"))
w = work["walls"]
walls = Wall(Point(up(w[0]), up(w[1])), Point(up(w[2]), up(w[3])))
out = (torch.empty((bench.N_AGENTS, 32), dtype=torch.float32, device=dev),
       torch.empty((bench.N_AGENTS, 32), dtype=torch.int32, device=dev))
names = {"": "default(RPT16/1)", "1": "RPT8/minb3", "2": "RPT4/minb3", "3": "RPT4/minb4", "4": "RPT4/minb5", "5": "RPT2/minb6",
         "6": "RPT4/minb2", "7": "RPT8/minb2"}
ref = None
for v in (sys.argv[1:] or list(names)):
    if v:
        os.environ["FGX_RC_VARIANT"] = v
    else:
        os.environ.pop("FGX_RC_VARIANT", None)
    for _ in range(3):
        sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        sm.raycast_walls((x, y, ang), bench.RAY_SPAN, bench.RAY_LEN, walls, act, 32, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    chk = (float(out[0].double().sum()), int(out[1].long().sum()))
    ref = ref or chk
    print(f"variant {v or '-':>2} {names.get(v, '?'):12s} {ms:8.3f} ms  {3.2e10 / ms / 1e9:8.1f} Gtests/s  same_output={chk == ref}")
