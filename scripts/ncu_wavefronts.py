#!/usr/bin/env python
"""Per-instruction shared-memory wavefronts / global tag requests from an
`ncu --page source --csv --print-source cuda,sass` export (LSU pipe pressure)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nctas = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cur_file = cur_line = hdr = None
seen, out = set(), []
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
    if r[0] == "" and r[2].startswith("0x") and r[2] not in seen:
        seen.add(r[2])
        d = {h: r[i] for i, h in enumerate(hdr)}
        try:
            w, wi = int(d["L1 Wavefronts Shared"] or 0), int(d["L1 Wavefronts Shared Ideal"] or 0)
            tag = int(d["L1 Tag Requests Global"] or 0)
            ex = int(d["Instructions Executed"])
        except ValueError:
            continue
        if w or tag:
            out.append((cur_file, cur_line, r[3].strip(), w, wi, tag, ex))
tw, ti, tt = sum(o[3] for o in out), sum(o[4] for o in out), sum(o[5] for o in out)
print(f"shared wavefronts {tw // nctas}/CTA (ideal {ti // nctas}), global tag requests {tt // nctas}/CTA")
out.sort(key=lambda o: -(o[3] + o[5]))
for o in out[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{o[0].replace('wsmg_', '')[:12]:12s}:{o[1]:4d} wf/CTA {o[3] // nctas:6d} ideal {o[4] // nctas:6d} gtag/CTA {o[5] // nctas:6d} "
          f"exec/CTA {o[6] // nctas:5d}  {o[2][:70]}")
