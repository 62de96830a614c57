"""A few eager C2 foraging ticks, for an ncu launch list (python scripts/c2_ticks.py [ticks])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foragax_b200 import random as fgx_random  # noqa: E402
from foragax_b200.examples import foraging as fg  # noqa: E402

fg.use_variant("sim")
w = fg.Foraging_world(fg.default_params("sim"), fgx_random.PRNGKey(1), 100)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    w.step()
torch.cuda.synchronize()
print("ok")
