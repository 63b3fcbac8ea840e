#!/usr/bin/env python
"""Per-phase totals of an `ncu --page source --csv --print-source cuda,sass` export of k_fused: executed warp
instructions, stall samples and shared-memory wavefronts per phase of wsmg_body.h.  A SASS instruction belongs to the
phase of the nearest wsmg_body.h line at or before it in address order (inlined helpers carry other files' lines).
Usage: ncu_phases2.py <csv> <n_ctas> <line:name> [<line:name> ...]   (phase start lines of the profiled source)"""
import csv, sys, collections
path, nctas = sys.argv[1], int(sys.argv[2])
marks = sorted((int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[3:])
rows = list(csv.reader(open(path)))
hdr = cur_file = cur_line = None
sass = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if not hdr or not r: continue
    if r[0] not in ("", "Function Name") and r[2] == "-":
        cur_line = int(r[0]); continue
    if r[0] == "" and r[2].startswith("0x"):
        def col(name):
            try: return int(r[hdr.index(name)])
            except ValueError: return 0
        a = int(r[2], 16)
        prev = sass.get(a)
        # the same SASS row appears under several source files; keep the wsmg_body.h attribution when there is one
        if prev is None or (cur_file == "wsmg_body.h" and prev[0] != "wsmg_body.h"):
            sass[a] = (cur_file, cur_line, r[3].strip(), col("# Samples"), col("Instructions Executed"), col("L1 Wavefronts Shared"),
                       col("L1 Tag Requests Global"), col("stall_barrier"), col("stall_short_sb"), col("stall_long_sb"), col("stall_mio"), col("stall_wait"))
def phase_of(line):
    name = "prologue"
    for l, n in marks:
        if line >= l: name = n
    return name
tot = collections.defaultdict(lambda: [0] * 9)
cur = "prologue"
for a in sorted(sass):
    f, l, s, smp, ins, wf, tag, sb, ss, sl, sm, sw = sass[a]
    if f == "wsmg_body.h": cur = phase_of(l)
    t = tot[cur]
    for i, v in enumerate((ins, smp, wf, tag, sb, ss, sl, sm, sw)): t[i] += v
ti = sum(t[0] for t in tot.values()); ts = sum(t[1] for t in tot.values())
print(f"{'phase':16s} {'instr/CTA':>10s} {'%instr':>7s} {'%samples':>8s} {'smem wf/CTA':>11s} {'gtag/CTA':>9s}   stall samples: barrier short_sb long_sb mio wait (% of all)")
for name in ["prologue"] + [n for _, n in marks]:
    t = tot.get(name)
    if not t: continue
    print(f"{name:16s} {t[0] / nctas:10.0f} {100 * t[0] / ti:7.1f} {100 * t[1] / ts:8.1f} {t[2] / nctas:11.0f} {t[3] / nctas:9.0f}   " +
          " ".join(f"{100 * v / ts:5.1f}" for v in t[4:]))
print(f"{'total':16s} {ti / nctas:10.0f} {100.0:7.1f} {100.0:8.1f} {sum(t[2] for t in tot.values()) / nctas:11.0f} {sum(t[3] for t in tot.values()) / nctas:9.0f}")
