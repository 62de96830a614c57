"""SASS evidence for profiles/: per kernel of libforagax_b200.so, counts of the mnemonics that show the sm_100a
features the kernels are written for -- UBLKCP (cp.async.bulk = 1-D TMA), SYNCS (mbarrier), FADD2 / FMUL2 /
FFMA2 (Blackwell packed FP32), REDUX (warp reductions), ATOM / RED, LDG.E.128.  Usage:
  python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "foragax_b200", "libforagax_b200.so")
PAT = {"UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS", "FADD2": r"\bFADD2", "FMUL2": r"\bFMUL2", "FFMA2": r"\bFFMA2",
       "REDUX": r"\bREDUX", "LDG.128": r"\bLDG\.E\.128|\bLDG\.E\.128\.CONSTANT|\bLDG\.E\.CONSTANT\.128", "STG.128": r"\bSTG\.E\.128",
       "MATCH": r"\bMATCH", "VOTE": r"\bVOTE", "ATOM/RED": r"\bATOM|\bRED\."}
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts, total, name = collections.OrderedDict(), collections.Counter(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        counts[name] = collections.Counter()
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[name]["instr"] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                counts[name][k] += 1
                total[k] += 1
print("# SASS summary of `foragax_b200/libforagax_b200.so` (round 2)\n")
print(f"`cuobjdump -sass` of the in-tree library; cubins: **{', '.join(arch)}** only ({len(counts)} kernels).  Counts of the "
      "mnemonics that show what the kernels are built from (PTX names never appear in SASS: `cp.async.bulk` = `UBLKCP`, "
      "`mbarrier.*` = `SYNCS`, packed FP32 `add/mul.f32x2` = `FADD2` / `FMUL2`).\n")
keys = list(PAT)
print("| kernel | instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for n, c in counts.items():
    if any(c[k] for k in keys):
        print(f"| `{n[:70]}` | {c['instr']} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
print("\n**Totals**: " + ", ".join(f"{k} {total[k]}" for k in keys))
