#!/usr/bin/env python3
"""Opcode histogram of one kernel from a .so / .cubin (cuobjdump -sass).  Usage: sass_hist.py <binary> <substring>"""
import collections
import re
import subprocess
import sys

binary, needle = sys.argv[1], sys.argv[2]
text = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or needle not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        hist[op.split(".")[0] if len(sys.argv) < 4 else op] += 1
        total += 1
print("total", total)
for op, c in hist.most_common(45):
    print("%6d %s" % (c, op))
