#!/usr/bin/env python3
"""Dynamic opcode histogram + top stall lines from `ncu -i X.ncu-rep --page source --csv`.
Usage: ncu_src_hist.py <csv> [kernel-index]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
# the file holds one table per profiled launch: split on the "Kernel Name" marker rows
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = tables[k]
hdr = t["rows"][0]
ix = {h: i for i, h in enumerate(hdr)}
hist, total, samples = collections.Counter(), 0, collections.Counter()
lines = []
for r in t["rows"][1:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m:
        continue
    n = int(r[ix["Instructions Executed"]] or 0)
    op = m.group(1).split(".")[0]
    hist[op] += n
    total += n
    lines.append((int(r[ix["# Samples"]] or 0), n, src))
print(t["name"])
print("warp-level instructions executed:", total)
for op, c in hist.most_common(30):
    print("%10d %5.1f%% %s" % (c, 100.0 * c / total, op))
print("top sampled lines:")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print("%6d %9d  %s" % (s, n, src))
