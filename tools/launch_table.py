#!/usr/bin/env python3
"""Print an ncu --csv launch log (metrics gpu__time_duration.sum, dram__bytes_*.sum) as one line per launch."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:70]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), v in d.items():
    t = v.get("gpu__time_duration.sum", 0) / 1e6
    rd, wr = v.get("dram__bytes_read.sum", 0) / 1e9, v.get("dram__bytes_write.sum", 0) / 1e9
    bw = (rd + wr) / t if t else 0
    print(f"{i:4d} {t:9.3f} ms  rd {rd:7.2f} GB  wr {wr:7.2f} GB  {bw:6.2f} TB/s  {k}")
