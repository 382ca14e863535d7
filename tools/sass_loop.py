#!/usr/bin/env python3
"""Static view of the Euler loop of a step kernel in a SASS listing (tools/sass_listing.sh): the backward branch whose body
holds the most DFMA without containing another such loop; prints its instruction mix.  Usage: tools/sass_loop.py <listing> [-v]"""
import collections
import re
import sys

ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, s in ins:
    m = re.search(r'BRA(?:\.\w+)*\s+(?:.*?)0x([0-9a-f]+)', s)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
cand = []
for t, a in loops:
    body = [x for x in ins if t <= x[0] <= a]
    if sum('DFMA' in x[1] for x in body) >= 100:
        cand.append((t, a, body))
inner = [c for c in cand if not any(o is not c and c[0] <= o[0] and o[1] <= c[1] for o in cand)]
best = max(inner, key=lambda c: sum('DFMA' in x[1] for x in c[2]))
t, a, body = best
mix = collections.Counter()
for _, s in body:
    w = s.split()
    m = w[1] if w[0].startswith('@') else w[0]
    mix[m.split('.')[0] + ('.MOV' if '.MOV' in m else '')] += 1
fp = sum(mix[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"loop {t:#x} .. {a:#x}: {len(body)} instructions, {fp} on the FP64 pipe, model cycles 2*fp64 + other = {2 * fp + len(body) - fp}")
print(' '.join(f"{k}:{v}" for k, v in mix.most_common()))
if len(sys.argv) > 2:
    for ad, s in body:
        print(f"{ad:#06x} {s}")
