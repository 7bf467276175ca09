#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel, stall reasons, hottest source lines.
Usage: python tools/ncu_summary.py report.ncu-rep [kernel_index] [top_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ["Duration", "DRAM Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "Executed Ipc Active",
        "Issue Slots Busy", "No Eligible", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
        "Warp Cycles Per Issued Instruction", "Registers Per Thread", "Dynamic Shared Memory Per Block",
        "Theoretical Occupancy", "Achieved Occupancy", "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Busy", "Max Bandwidth",
        "Block Size", "Grid Size", "Local Load", "Local Store", "Shared Load", "Bank"]
for line in det.splitlines():
    if "kp_" in line and "(" in line and "Context" in line:
        print(line.strip()[:120])
    elif any(k in line for k in keys):
        print("   ", " ".join(line.split()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, kernel, hdr, agg = None, 0, None, {}
stall = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        kernel += 1
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or kernel == 0:
        continue
    # kernels repeat per file; kernel index = (kernel-1) // nfiles is unknown, so key on first-seen order of (file) groups
    if r and r[0].isdigit():
        key = (kernel, cur, int(r[0]))
        s = int(r[4]) if r[4].isdigit() else 0
        ins = int(r[7]) if r[7].isdigit() else 0
        agg[key] = (s, ins, r[1].strip())
    elif len(r) == len(hdr) and r[0] == "" and r[2].startswith("0x"):
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h and r[idx[h]].isdigit():
                stall[(kernel, h)] = stall.get((kernel, h), 0) + int(r[idx[h]])
kernels = sorted({k[0] for k in agg})
files = sorted({k[1] for k in agg})
nf = len(files)
sel = [k for k in kernels][(kidx - 1) * nf: kidx * nf]
tot_s = sum(v[0] for k, v in agg.items() if k[0] in sel)
tot_i = sum(v[1] for k, v in agg.items() if k[0] in sel)
print(f"kernel #{kidx}: samples {tot_s} warp-instructions {tot_i}")
st = {}
for (k, h), v in stall.items():
    if k in sel:
        st[h] = st.get(h, 0) + v
ssum = sum(st.values()) or 1
print("stalls:", ", ".join(f"{h[6:]} {100 * v / ssum:.1f}%" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
byfile = {}
for k, v in agg.items():
    if k[0] in sel:
        byfile.setdefault(k[1], [0, 0])
        byfile[k[1]][0] += v[0]
        byfile[k[1]][1] += v[1]
print("by file:", {f: (f"{100 * a / max(tot_s, 1):.1f}% smp", f"{100 * b / max(tot_i, 1):.1f}% ins") for f, (a, b) in byfile.items()})
for k, (s, i, srcline) in sorted(((k, v) for k, v in agg.items() if k[0] in sel), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[1][:14]:14s}:{k[2]:4d} {100 * s / max(tot_s, 1):5.1f}% smp {100 * i / max(tot_i, 1):5.1f}% ins  {srcline[:100]}")
