#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: split the SASS into
segments of similar execution count and print each segment's share of executed warp
instructions / stall samples with its dominant source lines and opcodes."""
import collections
import csv
import sys

path = sys.argv[1]
nctas = int(sys.argv[2]) if len(sys.argv) > 2 else 1
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.012
rows = list(csv.reader(open(path)))
cur_file = cur_line = hdr = None
sass = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or not r:
        continue
    if r[0] not in ("", "Function Name") and r[2] == "-":
        cur_line = int(r[0])
        continue
    if r[0] == "" and r[2].startswith("0x"):
        try:
            sass[int(r[2], 16)] = (cur_file, cur_line, r[3].strip(), int(r[hdr.index("# Samples")]),
                                   int(r[hdr.index("Instructions Executed")]))
        except ValueError:
            pass
addrs = sorted(sass)
ti = sum(sass[a][4] for a in addrs)
tsmp = sum(sass[a][3] for a in addrs)
segs, cur = [], None
for a in addrs:
    f, l, s, smp, ins = sass[a]
    if cur is None or ins == 0 or cur["last"] == 0 or not (0.67 < ins / max(cur["last"], 1) < 1.5):
        cur = {"start": a, "n": 0, "ins": 0, "smp": 0, "lines": collections.Counter(), "last": ins, "ops": collections.Counter()}
        segs.append(cur)
    cur["n"] += 1
    cur["ins"] += ins
    cur["smp"] += smp
    cur["last"] = ins if ins else cur["last"]
    cur["lines"][(f, l)] += ins
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    cur["ops"][op.split(".")[0]] += ins
print(f"static SASS {len(addrs)}, executed warp-instr {ti} ({ti // nctas}/CTA), samples {tsmp}")
cov = 0
for s in segs:
    if s["ins"] / ti <= thresh and s["smp"] / max(tsmp, 1) <= thresh:
        continue
    cov += s["ins"]
    top = ", ".join(f"{f.replace('wsmg_', '')}:{l}" for (f, l), _ in s["lines"].most_common(4))
    ops = ", ".join(f"{o}:{100 * c / max(s['ins'],1):.0f}%" for o, c in s["ops"].most_common(6))
    print(f"@{addrs.index(s['start']):5d} n={s['n']:4d} instr {100 * s['ins'] / ti:5.1f}% samples {100 * s['smp'] / max(tsmp,1):5.1f}% "
          f"exec/instr/CTA {s['ins'] // max(s['n'], 1) // nctas:6d}  [{top}] [{ops}]")
print(f"covered {100 * cov / ti:.1f}% of instructions")
