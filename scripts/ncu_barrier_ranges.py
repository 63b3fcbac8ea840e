#!/usr/bin/env python
"""Split an `ncu --page source --csv --print-source cuda,sass` export of k_fused at its CTA-wide barriers (BAR.SYNC) and
print, per range of SASS between two barriers, the executed warp instructions, stall samples and shared-memory
wavefronts per CTA.  The ranges are the kernel's phases in program order: prologue | (first barrier inside the scatter)
| scatter | key decode (2 ranges) | first rotation | translation tables | band loop (its named barrier counts) |
output rotation.  Usage: ncu_barrier_ranges.py <csv> [n_ctas]"""
import csv, sys
path=sys.argv[1]; NC=int(sys.argv[2]) if len(sys.argv)>2 else 4096
rows=list(csv.reader(open(path)))
hdr=None; sass={}
for r in rows:
    if r and r[0]=="Line No": hdr=r; continue
    if not hdr or not r: continue
    if r[0]=="" and len(r)>3 and r[2].startswith("0x"):
        a=int(r[2],16)
        try: sass[a]=(r[3].strip(), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), int(r[hdr.index("L1 Wavefronts Shared")] or 0))
        except ValueError: pass
addrs=sorted(sass); base=addrs[0]
marks=[]
for a in addrs:
    op=sass[a][0]
    if op.startswith("BAR.SYNC") or "BAR.SYNC" in op: marks.append(a)
print("barriers at", [(m-base)//16 for m in marks])
tot=sum(v[1] for v in sass.values()); ts=sum(v[2] for v in sass.values())
prev=base
for m in marks+[addrs[-1]+16]:
    seg=[sass[a] for a in addrs if prev<=a<m]
    ins=sum(x[1] for x in seg); smp=sum(x[2] for x in seg); wf=sum(x[3] for x in seg)
    print(f"[{(prev-base)//16:5d},{(m-base)//16:5d}) n={len(seg):5d} instr/CTA {ins/NC:8.0f} ({100*ins/tot:5.1f}%) samples {100*smp/ts:5.1f}% wf/CTA {wf/NC:7.0f}")
    prev=m
