"""Times fgx_argsort (general multi-valued float keys and two-valued flags) at a few sizes.
python scripts/time_argsort.py   (FGX_SORT_MULTI=1 selects one launch per radix phase)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foragax_b200.base.agent_methods as am  # noqa: E402

dev = torch.device("cuda", 0)
for n in (20_000, 100_000, 1_000_000, 10_000_000):
    g = torch.Generator(device=dev).manual_seed(n)
    keys = {"float": torch.rand(n, device=dev, generator=g) * 100 - 50,
            "flag": (torch.rand(n, device=dev, generator=g) < 0.5).float()}
    for name, k in keys.items():
        for _ in range(3):
            perm = am.argsort(k, ascend=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            perm = am.argsort(k, ascend=False)
        e1.record()
        torch.cuda.synchronize()
        ref = torch.sort(-k, stable=True).indices.to(torch.int32)
        print(f"n={n:>9} {name:5s}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us  exact={bool(torch.equal(perm, ref))}")
