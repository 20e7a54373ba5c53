#!/usr/bin/env python
"""Opcode mix / hottest SASS lines of one kernel from `ncu --page source --csv` output.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv [kernel_index] [top_lines]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1][:120])
hdr = blk[1]
data = [r for r in blk[2:] if len(r) > 8]
isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in data)
nsamp = sum(int(r[iss]) for r in data)
print(f"total warp instructions {tot / 1e9:.2f} G, stall samples {nsamp}, SASS lines {len(data)}")
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc].strip())
    o = m.group(2).split(".")[0] if m else r[isrc]
    op[o] += int(r[ie])
    ops[o] += int(r[iss])
for o, c in op.most_common(24):
    print(f"  {o:12s} {c / 1e9:7.2f} G {100 * c / tot:5.1f}%   stall samples {100 * ops[o] / max(nsamp, 1):5.1f}%")
if top:
    print("hottest lines by stall samples:")
    for r in sorted(data, key=lambda r: -int(r[iss]))[:top]:
        print(f"  {int(r[iss]):7d} {int(r[ie]) / 1e6:9.1f}M  {r[isrc].strip()[:90]}")
