#!/bin/bash
# Opcode histogram of the shipped kernels (cuobjdump -sass of ws-mgmap_b200/lib/libwsmg.so) -> profiles/<tag>_sass_summary.txt
TAG=${1:?tag}
cd "$(dirname "$0")/.."
LIB=ws-mgmap_b200/lib/libwsmg.so
OUT=profiles/${TAG}_sass_summary.txt
cuobjdump -sass $LIB > /tmp/sass_all.txt
{
  echo "# cuobjdump -sass $LIB  ($(cuobjdump -lelf $LIB | head -3 | tr '\n' ' '))"
  echo "# per kernel: SASS instruction count and the 24 most frequent opcodes; then the Blackwell / TMA markers"
  python - <<'PY'
import re, collections
cur = None; ops = collections.OrderedDict()
for line in open('/tmp/sass_all.txt'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); ops[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        ops[cur][m.group(1)] += 1
import subprocess
for fn, c in ops.items():
    name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()[:150]
    n = sum(c.values())
    print(f"\n## {name}\n   {n} instructions: " + ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
    mark = {k: v for k, v in c.items() if re.match(r'(UTMA|UBLKCP|SYNCS|ATOMS|ATOMG|REDG|RED|FFMA2|FMUL2|FADD2|LDGSTS|UTCBAR|TCGEN|ACQBULK|LDGDEPBAR|BAR|WARPSYNC)', k)}
    print("   markers: " + ", ".join(f"{k} {v}" for k, v in sorted(mark.items())))
PY
} > $OUT
wc -l $OUT; grep -A3 "k_fused<100, 240, 50176, true, true, 0>" $OUT | cut -c1-400
