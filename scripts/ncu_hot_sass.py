#!/usr/bin/env python
"""Rank the SASS instructions of one captured launch by warp-stall samples.

    ncu -i <rep> --page source --csv --print-source sass --launch-skip N --launch-count 1 > /tmp/src.csv
    python scripts/ncu_hot_sass.py /tmp/src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[h + 1:]):
    if len(r) > isamp and r[isamp].isdigit():
        data.append((int(r[isamp]), i, r[isrc].strip(), int(r[iex] or 0)))
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}, {len(data)} instructions")
for s, i, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  @{i:5d} ex={ex:8d}  {src[:110]}")
