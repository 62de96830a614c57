"""Config C5 sharded over N GPUs: 10 M Dice slots (D20) in contiguous blocks, one process per GPU
(`foragax_b200/examples/churn_sharded.py`; the same leg rides in `bench.py`'s JSON line as `c5_churn` when N > 1).
Tick = local step / select / remove -> GLOBAL pack of the actives (one kernel: compaction + peer-memory stores over
NVLink; counts and ordering through peer-memory mailboxes, no NCCL) -> local select -> global add range -> local add;
two ticks are captured as one CUDA graph per rank and replayed.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_churn_sharded.py
Rank 0 prints one JSON line (max over ranks of the CUDA-event times)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foragax_b200.examples.churn import ROW_BYTES  # noqa: E402
from foragax_b200.examples.churn_sharded import ShardedChurnWorld  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=10_000_000)
ap.add_argument("--ticks", type=int, default=20)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if "RANK" not in os.environ:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", RANK="0", WORLD_SIZE="1")
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
w = ShardedChurnWorld(args.slots)
res = w.time_ticks(args.ticks)
t = torch.tensor([res["ms_per_tick"], res["eager_ms_per_tick"]], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
c = torch.tensor([res["final_active_local"], res["rows_sent_remote_per_tick"]], dtype=torch.float64, device="cuda")
dist.all_reduce(c)
if rank == 0:
    print(json.dumps({"config": "C5 sharded churn", "slots": args.slots, "n_gpus": world, "tick_ms": float(t[0]),
                      "eager_tick_ms": float(t[1]), "mode": res["mode"], "slot_steps_per_sec": args.slots / (float(t[0]) * 1e-3),
                      "active": int(c[0]), "nvlink_row_bytes_per_tick": float(c[1]) * ROW_BYTES}))
dist.destroy_process_group()
