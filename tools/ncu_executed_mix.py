#!/usr/bin/env python3
"""Executed-instruction mix per warp-step of the fused step kernel from an `ncu --set full --import-source on` report:
the SASS instructions that execute at least once per two warp-steps (the Euler loop), grouped by mnemonic.
Usage: tools/ncu_executed_mix.py <report.ncu-rep> <aircraft> <euler steps> > profiles/<name>.md"""
import collections
import csv
import io
import subprocess
import sys

rep, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
data = [(r[ia].strip(), int(r[ie])) for r in rows if len(r) > ie and r[ie].isdigit()]
steps = (n + 31) // 32 * k
hot = [(s, e) for s, e in data if e >= 0.5 * steps]
mix = collections.Counter()
for s, e in hot:
    t = s.split()
    m = t[1] if t[0].startswith("@") else t[0]
    mix[m.split(".")[0]] += e / steps
fp64 = sum(v for m, v in mix.items() if m in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"# Executed SASS per warp-step of the Euler loop ({rep})\n")
print(f"{n} aircraft x {k} steps = {steps} warp-steps; {len(hot)} static instructions execute >= 0.5 times per warp-step.\n")
print("| mnemonic | executed per warp-step |\n|---|---|")
for m, v in mix.most_common():
    print(f"| {m} | {v:.1f} |")
tot = sum(mix.values())
print(f"| **total** | **{tot:.1f}** |\n")
print(f"FP64-pipe instructions (DFMA + DMUL + DADD + DSETP): {fp64:.0f} of {tot:.0f}; executed flop per aircraft-step "
      f"(DFMA = 2): {2 * mix['DFMA'] + mix['DMUL'] + mix['DADD']:.0f}.\n")
print("DSETP in the loop (FP64 pipe, 0 flop):\n")
for s, e in hot:
    if "DSETP" in s:
        print(f"    {s}")
