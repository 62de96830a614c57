"""C5: fgx_sort_rows (two-valued key + 7 payload leaves, 10 M slots) against the key pattern -- the churn state, identity
permutations (all active / first half active) and random flags -- and a plain torch copy_ of the same seven leaves for
reference: the sort moves its 640 MB as fast as a device copy moves the 560 MB of rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foragax_b200.examples.churn import ChurnWorld
from foragax_b200.base import agent_methods as M
from foragax_b200 import struct
dev = torch.device("cuda", 0)
n = 10_000_000
w = ChurnWorld(n)
for _ in range(3): w.tick()
ag = w.agents
out = struct.tree_map(torch.empty_like, ag)
def t(flags, tag):
    a = ag.replace(active_state=flags)
    for _ in range(3): M.jit_sort_agents(a.active_state, False, a, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): M.jit_sort_agents(a.active_state, False, a, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{tag:40s} {ms:.4f} ms  {640e6 / ms / 1e6:.0f} GB/s algorithmic")
g = torch.Generator(device=dev); g.manual_seed(0)
t(ag.active_state.clone(), "churn state (packed front + 5% holes)")
t(torch.ones(n, device=dev), "all active (identity copy)")
half = torch.zeros(n, device=dev); half[: n // 2] = 1
t(half, "first half active (identity, two-valued)")
t((torch.rand(n, device=dev, generator=g) < 0.5).float(), "random 50 %")
t((torch.rand(n, device=dev, generator=g) < 0.05).float(), "random 5 %")
# plain device copy of the same bytes for reference
leaves = struct.tree_leaves(ag); outs = struct.tree_leaves(out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    for a, b in zip(leaves, outs): b.copy_(a)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    for a, b in zip(leaves, outs): b.copy_(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"{'torch copy_ of the 7 leaves (560 MB)':40s} {ms:.4f} ms  {560e6 / ms / 1e6:.0f} GB/s")
