#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + the most-sampled SASS instructions.  usage: ncu_top.py rep [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
vals = rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} = {vals[i]} {units[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.2:
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {float(vals[i]):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# one block per profiled kernel: a "Kernel Name" row, a header row, then the SASS lines; argv[3] picks the block
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
# the source page repeats a block per view; take the kidx-th DISTINCT kernel, in launch order
names, firsts = [], []
for i in starts:
    if rows[i][1] not in names:
        names.append(rows[i][1])
        firsts.append(i)
b0 = firsts[min(kidx, len(firsts) - 1)]
b1 = min([i for i in starts if i > b0] + [len(rows)])
print("source block:", rows[b0][1][:100])
hdr = rows[b0 + 1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[b0 + 2:b1] if len(r) > max(isamp, iex)]
tot = sum(int(r[isamp]) for r in data)
texec = sum(int(r[iex]) for r in data)
print(f"total samples {tot}, SASS instructions {len(data)}, warp-instructions executed {texec}")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:n]:
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{idx:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9} {r[isrc].strip()[:60]:60s} {st}")
