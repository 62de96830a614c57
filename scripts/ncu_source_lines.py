"""Per-source-line instruction counts and stall samples from an ncu report captured with --import-source on:
python scripts/ncu_source_lines.py rep.ncu-rep [top]  (reads `ncu --page source --csv --print-source cuda,sass`)."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 8 and r[0].isdigit():
        num = lambda v: float(v) if v not in ("", "-") else 0.0  # noqa: E731
        lines.append((cur, int(r[0]), r[1].strip()[:90], int(num(r[hdr["Instructions Executed"]])),
                      int(num(r[hdr["# Samples"]])), num(r[hdr["Avg. Threads Executed"]])))
tot_i = sum(x[3] for x in lines) or 1
tot_s = sum(x[4] for x in lines) or 1
print(f"total warp instructions {tot_i:,}; stall samples {tot_s:,}")
print("| file:line | inst % | samples % | avg threads | source |\n|---|---|---|---|---|")
for f, ln, src, ins, smp, thr in sorted(lines, key=lambda x: -x[4])[:top]:
    print(f"| {f}:{ln} | {100 * ins / tot_i:.1f} | {100 * smp / tot_s:.1f} | {thr:.0f} | `{src}` |")
