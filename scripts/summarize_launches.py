#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:

    python scripts/summarize_launches.py gpurun_out/launches.csv "title" > profiles/<name>_summary.csv

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"(?:\w+::)*gemm_kernel<(?:\w+::)*(?:\(anonymous namespace\)::)?([\w<>, ]+?)>\(", name)
    if m:
        return "gemm_kernel<" + m.group(1) + ">"
    name = re.sub(r"\(.*", "", name)
    return name.split("::")[-1] if "at::native" in name else name


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = OrderedDict()
    for r in rows:
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0, r[gi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3
    total = sum(a[1] for a in agg.values())
    print(f"# {title}")
    print("# per-launch times are cold-cache and serialised: compare SHARES")
    print("kernel,launches,total_us,avg_us,share,grid")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{a[1] / total:.3f},\"{a[2]}\"")


if __name__ == "__main__":
    main()
