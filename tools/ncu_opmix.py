#!/usr/bin/env python3
"""Executed-instruction mix of one kernel of an .ncu-rep (source page, no GPU): opcode -> executed warp-instructions, share, and how
many static instructions of the main loop carry it.  usage: ncu_opmix.py rep kernel_index"""
import csv, io, subprocess, sys, collections, re
rep, kidx = sys.argv[1], int(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
names, firsts = [], []
for i in starts:
    if rows[i][1] not in names:
        names.append(rows[i][1]); firsts.append(i)
b0 = firsts[kidx]; b1 = min([i for i in starts if i > b0] + [len(rows)])
hdr = rows[b0 + 1]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
print(rows[b0][1][:90])
h = collections.Counter(); tot = 0
loop = collections.Counter()
for r in rows[b0 + 2:b1]:
    if len(r) <= max(isrc, iex): continue
    ex = int(r[iex]); s = r[isrc].strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    op = s.split()[0] if s else "?"
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if "." in op else "")
    h[op] += ex; tot += ex
    if 4.0e6 < ex < 4.5e6: loop[op] += 1
print("total", tot)
for op, c in h.most_common(40): print(f"{op:22s} {c/1e6:10.1f}M {100*c/tot:5.1f}%   static-in-loop {loop[op]}")
print("static loop instrs", sum(loop.values()))
