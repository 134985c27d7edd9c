#!/usr/bin/env python
"""FP64 instruction counts per CUDA source line from `ncu --page source --csv --print-source cuda,sass` output:
which source lines issue the DFMA/DMUL/DADD/DSETP/MUFU.*64 instructions (the FP64 pipe is the bottleneck of the BA kernels)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for i, r in enumerate(rows[:10]):
    if 'Source' in r:
        hdr, start = r, i + 1
        break
ex = hdr.index('Instructions Executed')
cur, per, txt, allinst = None, collections.Counter(), {}, 0
for r in rows[start:]:
    if len(r) <= ex:
        continue
    if r[0].strip():
        cur = r[0].strip(); txt[cur] = r[1].strip(); continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)\b', r[3].strip())
    if not m:
        continue
    try:
        n = int(r[ex])
    except ValueError:
        continue
    allinst += n
    op = m.group(2)
    if op in ('DFMA', 'DMUL', 'DADD') or op.startswith('DSETP') or op.startswith('DMNMX') or '64H' in op or op.startswith('DMMA'):
        per[cur] += n
tot = sum(per.values())
print(f"# FP64-pipe warp instructions {tot} of {allinst} executed ({100.0 * tot / max(allinst, 1):.1f} %)")
for k, v in per.most_common(top):
    print(f"{100.0 * v / tot:5.1f}%  {v:9d}  line {k:>5s}  {txt.get(k, '')[:120]}")
