#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel of an .ncu-rep, grouped by SOURCE LINE.

    python tools/ncu_by_line.py prof.ncu-rep build/obj.o <kernel substring of the mangled name> [kernel index] [top N]

ncu's source page lists the SASS in program order; `nvdisasm --print-line-info` of the same cubin gives the
source line of every SASS instruction in the same order, so the two are joined by position."""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def line_table(obj, fn):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=d, capture_output=True)
        cub = next(Path(d).glob("*.cubin"))
        txt = subprocess.run(["nvdisasm", "--print-line-info", str(cub)], capture_output=True, text=True).stdout
    out, on, cur = [], False, None
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            on = fn in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((cur, m.group(2)))
    return out


def main():
    rep, obj, fn = sys.argv[1:4]
    kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    lines = line_table(obj, fn)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(txt.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    # (ncu prints every kernel twice when the report holds source and SASS views: keep one of each pair)
    if len(blocks) >= 2 and all(blocks[i]["name"] == blocks[i + 1]["name"] and len(blocks[i]["rows"]) == len(blocks[i + 1]["rows"])
                                for i in range(0, len(blocks) - 1, 2)):
        blocks = blocks[::2]
    blk = blocks[kidx]
    h = blk["hdr"]
    ie, ss = h.index("Instructions Executed"), h.index("# Samples")
    extra = [c for c in ("stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving",
                         "stall_math", "stall_lg", "stall_mio") if c in h]
    xi = [h.index(c) for c in extra]
    xs = {c: collections.Counter() for c in extra}
    assert len(blk["rows"]) == len(lines), (len(blk["rows"]), len(lines))
    inst, samp = collections.Counter(), collections.Counter()
    for (src, _), row in zip(lines, blk["rows"]):
        inst[src] += int(row[ie])
        samp[src] += int(row[ss])
        for c, i in zip(extra, xi):
            xs[c][src] += int(row[i] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(blk["name"][:120])
    print(f"total warp instructions {ti:.4g}, samples {ts}; of all samples: " +
          ", ".join(f"{c[6:]} {100 * sum(xs[c].values()) / max(ts, 1):.1f}%" for c in extra))
    for src, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        top2 = sorted(((xs[c][src], c[6:]) for c in extra), reverse=True)[:3]
        print(f"{str(src):45s} inst {v:12d} {100 * v / ti:5.1f}%   samples {100 * samp[src] / max(ts, 1):5.1f}%   " +
              " ".join(f"{n}={100 * x / max(ts, 1):.1f}" for x, n in top2))


if __name__ == "__main__":
    main()
