#!/usr/bin/env python3
"""Prints the SASS of the functions of a .so whose mangled name contains a substring, one instruction per line
(address, predicate+opcode, operands), plus an opcode histogram.  Usage: sass_fn.py <lib.so> <substring> [--hist-only]"""
import collections
import re
import subprocess
import sys

lib, sub = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, keep = None, False
hist = collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        keep = sub in cur
        if keep:
            print("#### " + cur)
        continue
    if not keep:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins = m.group(2).strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        hist[op] += 1
        if "--hist-only" not in sys.argv:
            print(m.group(1), ins)
print("#### static opcode histogram")
for op, c in hist.most_common(40):
    print("%6d %s" % (c, op))
