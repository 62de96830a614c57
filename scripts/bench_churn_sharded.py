"""Config C5 sharded over N GPUs: 10 M Dice slots (D20) in contiguous blocks, one process per GPU.
Tick = local step / select / remove -> GLOBAL pack of the actives (all-gathered counts + all-to-all of rows over
NCCL) -> local select -> GLOBAL add range (all-gathered counts) -> local add.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_churn_sharded.py
Rank 0 prints one JSON line (max over ranks of the CUDA-event time)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foragax_b200.base.agent_classes as fgx_classes  # noqa: E402
import foragax_b200.base.agent_methods as fgx_methods  # noqa: E402
from foragax_b200 import distributed as fd  # noqa: E402
from foragax_b200 import random as fgx_random  # noqa: E402
from foragax_b200.examples.hello_world import is_value, make_dice  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=10_000_000)
ap.add_argument("--ticks", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--graph", action="store_true",
                help="single rank, peer exchange: capture TWO ticks (buffer set A -> B -> A) into one CUDA graph and "
                     "replay it (with 2 ranks the capture of the NCCL collectives hung, so it is not offered there)")
ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                help="peer: state in symmetric memory, ONE scatter kernel over NVLink per pack (fgx_pack_rows_peer); "
                     "nccl: local sort + all_to_all_single per leaf + local sort (split sizes via the host)")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1 or args.exchange == "peer":
    if "RANK" not in os.environ:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", RANK="0", WORLD_SIZE="1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
N, SIDES = args.slots, 20
D = make_dice(SIDES)
lo, hi = fd.shard_bounds(N, world, rank)
n_local = hi - lo
n_active = N // 2
s = fgx_classes.Agent_Set(agent=D, num_total_agents=n_local, num_active_agents=0, agent_type=0)
# this rank's rows of the single-GPU set (fgx_create_base_shard): first n_active GLOBAL slots active
s.agents = fgx_methods.create_agents(None, fgx_classes.Agent_Set(agent=D, num_total_agents=N, num_active_agents=n_active,
                                                                  agent_type=0),
                                     fgx_random.PRNGKey(0), shard=(lo, n_local))
sort_active = lambda a: fgx_methods.jit_sort_agents(a.active_state, False, a)[0]  # noqa: E731


st = None
if args.exchange == "peer":
    from foragax_b200.peer import PeerPackedSet  # noqa: E402
    st = PeerPackedSet(s.agents)
    s.agents = st.agents


def tick_peer():
    fgx_methods.step_agents(None, None, s)
    n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
    fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}), s.agents)
    s.agents = st.pack_active()
    n_sel, _ = fgx_methods.jit_select_agents(is_value(SIDES), None, s.agents)
    k, local_count, _ = fd.global_add_range(n_sel, fgx_methods.count_active(s.agents), n_local, lo, N)
    fgx_methods.jit_add_agents(D.add_agent, local_count, None, s.agents, None)


def tick_nccl():
    s.agents = fgx_methods.step_agents(None, None, s)
    n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
    s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                             s.agents)
    s.agents = fd.global_pack_active(s.agents, sort_active, fgx_methods.count_active(s.agents))
    n_sel, _ = fgx_methods.jit_select_agents(is_value(SIDES), None, s.agents)
    k, local_count, _ = fd.global_add_range(n_sel, fgx_methods.count_active(s.agents), n_local, lo, N)
    s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, local_count, None, s.agents, None)


tick = tick_peer if args.exchange == "peer" else tick_nccl
for _ in range(args.warmup):
    tick()
torch.cuda.synchronize()
if dist.is_initialized():
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.ticks):
    tick()
e1.record()
torch.cuda.synchronize()
eager_ms = e0.elapsed_time(e1) / args.ticks
graph_ms = None
if args.graph and args.exchange == "peer" and world == 1:  # capture with NCCL collectives inside hung on this stack
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        tick()
        tick()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    # thread_local: the NCCL watchdog thread keeps polling events while this thread captures
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        tick()
        tick()   # two packs: the state is back in the buffer set it started from
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(args.ticks // 2):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / (2 * (args.ticks // 2))
ms = torch.tensor([eager_ms, graph_ms if graph_ms is not None else 0.0], device="cuda")
act = fgx_methods.count_active(s.agents).to(torch.int64).reshape(1)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(act)
if rank == 0:
    best = float(ms[1]) if graph_ms is not None else float(ms[0])
    print(json.dumps({"config": "C5 sharded churn", "slots": N, "n_gpus": world, "tick_ms": float(ms[0]),
                      "graph_tick_ms": float(ms[1]) if graph_ms is not None else None,
                      "slot_steps_per_sec": N / (best * 1e-3), "active": int(act),
                      "exchange": args.exchange}))
if dist.is_initialized():
    dist.destroy_process_group()
