#!/usr/bin/env python
"""Summarises an .ncu-rep (needs `ncu` on PATH): key raw metrics + time by source file + hottest lines.
Usage: python profiles/ncu_summary.py report.ncu-rep [n_lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"  {h} [{u}] = {v}")
    stalls = sorted(((int(float(v)), h) for h, v in zip(hdr, r) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v), reverse=True)
    print("  stalls:", ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {n}" for n, h in stalls[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
cur, fn = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
lines = {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1].split("(")[0].split("::")[-1].split("<")[0]
        continue
    if r[0] == "Line No":
        continue
    if len(r) > 8 and r[2] == "-" and r[0].isdigit():
        try:
            samp, inst, tinst = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
        except ValueError:
            continue
        a = agg[(fn, cur)]
        a[0] += samp; a[1] += inst; a[2] += tinst
        lines[(fn, cur, r[0])] = (samp, inst, tinst, r[1][:90])
for kernel in sorted({k[0] for k in agg}):
    sub = {k[1]: v for k, v in agg.items() if k[0] == kernel}
    tot = sum(v[0] for v in sub.values()) or 1
    toti = sum(v[1] for v in sub.values()) or 1
    print(f"== {kernel}: by file (samples %, instructions %, active lanes per instruction):")
    for k, v in sorted(sub.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:24s} {100 * v[0] / tot:5.1f}% {100 * v[1] / toti:5.1f}% {v[2] / max(v[1], 1):5.1f}")
    print("  hottest lines (samples %, instructions, lanes):")
    sel = {k: v for k, v in lines.items() if k[0] == kernel}
    for (_, f, ln), (samp, inst, tinst, text) in sorted(sel.items(), key=lambda kv: -kv[1][0])[:n_lines]:
        print(f"  {100 * samp / tot:5.1f}% {inst:10d} {tinst / max(inst, 1):5.1f}  {f}:{ln}  {text}")
