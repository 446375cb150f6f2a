#!/usr/bin/env python3
"""Per-kernel SASS mnemonic summary of libpandora_b200.so (cuobjdump -sass; no GPU needed): which kernels use bulk TMA (UBLKCP),
cp.async (LDGSTS), DPX packed min/max (VIMNMX*.U16x2), warp reductions (REDUX / CREDUX), mbarrier (SYNCS), and how many
instructions they have.  usage: python tools/sass_summary.py > profiles/<round>_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pandora_b200", "_lib", "libpandora_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
KEYS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "VIMNMX3.U16x2", "VIMNMX.U16x2", "VIADDMNMX", "REDUX", "CREDUX", "POPC", "RED.E", "ATOMG", "SHFL", "DFMA", "DMUL",
        "BAR.SYNC", "STG.E.128", "LDG.E.128", "FFMA", "FADD"]
print("arch:", re.search(r"arch = (\S+)", sass).group(1), " library:", os.path.relpath(lib, ROOT))
print(f"{'kernel':78s} {'regs':>4s} {'instr':>6s}  " + " ".join(f"{k}" for k in KEYS))
cur, counts, n = None, collections.Counter(), 0


def flush():
    if cur is None:
        return
    name = demangle(cur)
    name = name.replace("(anonymous namespace)::", "").replace("pb200::", "")
    name = re.sub(r"\((int|bool|unsigned int)\)", "", name)
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    r = regs.get(cur, (0, 0, 0))
    marks = " ".join(f"{k}={counts[k]}" for k in KEYS if counts[k])
    print(f"{name[:78]:78s} {r[0]:4d} {n:6d}  {marks}")


for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush()
        cur, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.x]+)", line)
    if m and cur:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[k] += 1
flush()
