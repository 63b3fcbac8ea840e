#!/usr/bin/env python
"""Executed warp-instructions and stall samples of k_fused per PHASE of the item body, from an
`ncu --page source --csv --print-source cuda,sass` export: SASS instructions are attributed to the source line ncu
reports, and the line is mapped to a phase through the `// ---- phase` markers of wsmg_body.h."""
import csv
import re
import sys

path, body, nitems = sys.argv[1], sys.argv[2], int(sys.argv[3])
marks = []
for n, line in enumerate(open(body), 1):
    m = re.match(r"\s*// ---- (.*?)(-{3,}.*)?$", line)
    if m:
        marks.append((n, m.group(1).strip()[:60]))
rows = list(csv.reader(open(path)))
cur_file = cur_line = hdr = None
acc = {}
tot_i = tot_s = 0
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
            smp, ins = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        if cur_file == "wsmg_body.h":
            ph = "prologue"
            for n, name in marks:
                if cur_line >= n:
                    ph = name
        else:
            ph = f"({cur_file})"
        a = acc.setdefault(ph, [0, 0])
        a[0] += ins
        a[1] += smp
        tot_i += ins
        tot_s += smp
print(f"total {tot_i} warp-instr ({tot_i // nitems}/item), {tot_s} samples")
for ph, (i, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * i / tot_i:5.1f}% instr {i // nitems:7d}/item   {100 * s / max(tot_s, 1):5.1f}% samples   {ph}")
