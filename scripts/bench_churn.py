"""Config C5: add/remove/sort churn stress on a 10M-slot Dice set with a D20 die
(draw==1 dies, draw==20 spawns: ~5 % + 5 % turnover per tick).  Tick = step -> select -> remove ->
sort(active desc) -> select -> clip -> add, all on-stream with device-resident counts
(foragax_b200/examples/churn.py; the same leg rides in bench.py's JSON line as `c5_churn`).

Prints per-op CUDA-event times and the achieved HBM bandwidth against each op's ALGORITHMIC bytes
(BASELINE.md section 3), then one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foragax_b200.examples import churn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=10_000_000)
ap.add_argument("--ticks", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

w = churn.ChurnWorld(args.slots)
tick_ms, ops, n_act = churn.time_ticks(w, args.ticks, args.warmup)
alg = churn.algorithmic_bytes(args.slots, n_act, 0.05 * n_act)
peak = 6509.1
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
N = args.slots
print(f"C5 churn: {N} slots, ~{n_act:.0f} active, D20; eager tick {tick_ms:.3f} ms "
      f"({N / tick_ms / 1e3:.1f} M slot-steps/s, {n_act / tick_ms / 1e3:.1f} M active agent-steps/s)")
rows = {}
for name in churn.OPS:
    ms = ops[name]
    gbs = alg[name] / (ms * 1e-3) / 1e9
    rows[name] = {"ms": float(ms), "alg_bytes": float(alg[name]), "gbs": gbs, "frac_of_hbm": gbs / peak}
    print(f"  {name:14s} {ms:8.3f} ms   {alg[name] / 1e6:9.1f} MB algorithmic   {gbs:8.1f} GB/s   {gbs / peak:5.2f} of HBM")
print(json.dumps({"config": "C5", "slots": N, "active": n_act, "tick_ms": tick_ms,
                  "slot_steps_per_sec": N / (tick_ms * 1e-3), "ops": rows, "hbm_peak_gbs": peak}))
