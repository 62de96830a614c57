"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list:
python scripts/ncu_launches.py gpurun_out/launches.csv [first_kernel_substring]  -> one period of launches."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():
        k = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", "")[:60])
        per.setdefault(k, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]])
seq = list(per.items())
mark = sys.argv[2] if len(sys.argv) > 2 else None
if mark:
    starts = [i for i, (k, v) in enumerate(seq) if mark in k[1]]
    if len(starts) >= 3:
        seq = seq[starts[1]:starts[2]]
tot = sum(v.get("gpu__time_duration.sum", 0) for _, v in seq) / 1e3
print(f"| kernel | us | share | dram rd MB | dram wr MB |\n|---|---|---|---|---|")
for k, v in seq:
    t = v.get("gpu__time_duration.sum", 0) / 1e3
    print(f"| {k[1]} | {t:.1f} | {t / tot:.3f} | {v.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
          f"{v.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
print(f"\nsum of kernel durations: {tot:.1f} us over {len(seq)} launches")
