"""Per-instruction executed counts of one kernel from an `ncu --set full --import-source on` capture: total warp
instructions, the histogram of execution counts (which groups the SASS into "per agent", "per list entry", ...
classes) and a few headline metrics.  Usage: python scripts/ncu_inst_hist.py report.ncu-rep [listing.txt]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iA, iS, iSm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[iA]) for r in data)
print("kernel:", rows[0][1][:100])
print("total warp instructions:", tot)
hist = Counter(int(r[iA]) for r in data)
for cnt, k in sorted(hist.items(), reverse=True)[:28]:
    print(f"  executed {cnt:>10} x {k:4d} instructions = {cnt * k / 1e6:8.1f} M")
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        for k, r in enumerate(data):
            f.write(f"{k} {r[iA]:>10} {r[iSm]:>6}  {r[iS][:90]}\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
want = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_alu.sum.pct",
        "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_lsu.sum.pct", "sm__inst_executed_pipe_xu.sum.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled")
for a, b in zip(rr[0], rr[2]):
    if any(a.startswith(w) for w in want) and "per_second" not in a:
        print(f"  {a} = {b}")
