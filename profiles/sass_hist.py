#!/usr/bin/env python
"""Opcode histogram (warp instructions executed) of one kernel from `ncu --page source --csv` output."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_ex, i_samp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= i_ex:
        continue
    try:
        n = int(r[i_ex])
    except ValueError:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[i_src])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += n
    tot += n
    try:
        samp[op] += int(r[i_samp])
    except ValueError:
        pass
print("total warp instructions", tot, "static instructions", len(rows) - 2)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:10s} {n:10d} {100 * n / tot:5.1f}%  stall samples {samp[op]}")
