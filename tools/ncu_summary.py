#!/usr/bin/env python3
"""Summarise an ncu capture for profiles/: key counters of every launch in one `--set full` report and the launch list
of a bench run.  Usage: tools/ncu_summary.py <report.ncu-rep> <launches.csv | -> <out.md> [bench.json]"""
import csv
import io
import json
import subprocess
import sys

rep, launches, out = sys.argv[1:4]
bench = sys.argv[4] if len(sys.argv) > 4 else None

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = []
for li, vals in enumerate(rows[2:]):
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def g(name):
        v, u = m.get(name, ("n/a", ""))
        return f"{v} {u}".strip()

    def f(name):
        try:
            return float(m[name][0].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")

    kernel = m["Kernel Name"][0]
    grid, block = m["Grid Size"][0], m["Block Size"][0]
    lines += [f"# ncu summary: `{kernel[:120]}`" + (f" (launch {li + 1} of the report)" if len(rows) > 3 else ""), "",
              f"report `{rep}` (ncu --set full --clock-control none --import-source on, grid {grid}, block {block})", ""]
    lds = f("smsp__sass_inst_executed_op_shared_ld.sum")
    wf = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    keys = [
        ("duration", "gpu__time_duration.sum"),
        ("registers / thread", "launch__registers_per_thread"),
        ("dynamic shared memory / CTA", "launch__shared_mem_per_block_dynamic"),
        ("achieved warps / SM", "sm__warps_active.avg.per_cycle_active"),
        ("FP64 pipe active (% of peak)", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue slots busy (%)", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("shared-memory data pipe (% of peak wavefronts)", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("ALU pipe (%)", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("LSU pipe (%)", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("XU pipe (%)", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("DFMA thread-inst / cycle / SMSP", "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed"),
        ("DMUL thread-inst / cycle / SMSP", "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed"),
        ("DADD thread-inst / cycle / SMSP", "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed"),
        ("branch efficiency (%)", "smsp__sass_average_branch_targets_threads_uniform.pct"),
        ("DRAM read", "dram__bytes_read.sum"),
        ("DRAM write", "dram__bytes_write.sum"),
        ("DRAM throughput (% of peak)", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1/TEX hit rate (%)", "l1tex__t_sector_hit_rate.pct"),
        ("warp instructions executed", "smsp__inst_executed.sum"),
        ("LDS instructions executed", "smsp__sass_inst_executed_op_shared_ld.sum"),
        ("shared wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ]
    lines += ["| counter | value |", "|---|---|"]
    for label, k in keys:
        lines.append(f"| {label} | {g(k)} |")
    if lds == lds and lds > 0:
        lines.append(f"| shared wavefronts per LDS | {wf / lds:.2f} |")
    lines += ["", "Warp stall reasons (average warps stalled per issued instruction):", "", "| reason | ratio |", "|---|---|"]
    stalls = sorted(((h.split("issue_stalled_")[1].split("_per_issue")[0], float(v[0])) for h, v in m.items()
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h),
                    key=lambda kv: -kv[1])
    for k, v in stalls[:9]:
        lines.append(f"| {k} | {v:.3f} |")
    lines.append("")

# launch list
if launches != "-":
    lines += [f"## Launch list (`{launches}`: ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache)", ""]
    agg = {}
    with open(launches) as fh:
        rd = csv.reader(l for l in fh if l.startswith('"'))
        head = next(rd)
        ki, vi = head.index("Kernel Name"), head.index("Metric Value")
        for r in rd:
            name = r[ki].split("(")[0][-60:]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    lines += ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |")
if bench:
    d = json.loads(open(bench).readline())
    lines += ["", "## Bench line of the same build (no profiler attached)", "",
              f"value {d['value']:.4g} {d['unit']}, ms/step {d['ms_per_step']:.2f}, roofline frac {d['roofline']['frac']:.3f} of "
              f"{d['roofline']['peak']:.1f} TFLOP/s FP64 (measured DFMA rate), kernel_ms {d['roofline']['kernel_ms']:.2f}, "
              f"e2e {d['e2e']['value']:.4g}, clocks {d['clocks']}"]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
