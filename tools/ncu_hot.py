#!/usr/bin/env python3
"""Per-opcode executed-instruction counts and stall-sample distribution of one kernel in an ncu report (source page).
Usage: tools/ncu_hot.py <report.ncu-rep> <warp_tasks> [window]"""
import collections, csv, io, subprocess, sys
rep, tasks = sys.argv[1], float(sys.argv[2])
win = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0    # which launch of the report
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:110])
hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
iS, iE, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot, stall, n, ns = collections.Counter(), collections.Counter(), 0, 0
def opc(src):
    t = src.split()
    m = t[1] if t[0].startswith('@') else t[0]
    return m.split('.')[0].rstrip(';')
for r in data:
    if not r[iE].isdigit(): continue
    e, s = int(r[iE]), int(r[iSt]); m = opc(r[iS])
    tot[m] += e; stall[m] += s; n += e; ns += s
print(f"executed warp instructions {n} = {n / tasks:.1f} per warp-task; stall samples {ns}")
for m, c in tot.most_common(22): print(f"  {m:10s} {c / tasks:8.1f}   samples {100 * stall[m] / max(ns, 1):5.1f}%")
print("top stall lines:")
for r in sorted([r for r in data if r[iSt].isdigit()], key=lambda r: -int(r[iSt]))[:14]:
    print(f"  {100 * int(r[iSt]) / max(ns, 1):5.1f}%  x{int(r[iE]) / tasks:5.2f}  {r[iS][:100]}")
if win:
    for i in range(0, len(data), win):
        w = [r for r in data[i:i + win] if r[iSt].isdigit()]
        s = sum(int(r[iSt]) for r in w); e = max([int(r[iE]) for r in w] or [0])
        if e: print(f"  [{i:5d}] {100 * s / max(ns, 1):5.1f}%  x{e / tasks:5.2f}  {sorted(set(opc(r[iS]) for r in w) & {'UTMALDG','UBLKCP','SYNCS','STG','LDG','LDL','STL','CALL','LDS','DFMA','BAR'})}")
