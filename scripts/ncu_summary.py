"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small markdown table
for profiles/.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
name_i = hdr.index("Kernel Name")
print(f"# ncu summary of `{rep}`\n")
for r in data:
    print(f"## {r[name_i][:110]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {r[i]} | {units[i]} |")
    print("\nwarp stall reasons (warps stalled per issue-active cycle):\n")
    print("| reason | ratio |\n|---|---|")
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            v = float(r[i] or 0)
            if v >= 0.01:
                print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
    print()
