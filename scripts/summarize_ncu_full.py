#!/usr/bin/env python
"""Summarise an `ncu --set full` report per captured launch (one CSV row each):

    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/summarize_ncu_full.py /tmp/raw.csv "title" > profiles/<name>_summary.csv

Columns: duration, DRAM read / write bytes (the `roofline.traffic` of bench.py = their sum, see profiles/ncu_traffic.json),
DRAM utilisation, tensor-pipe activity, resident warps, registers, grid, L2 -> SM read bytes, SM active / elapsed cycles
(their difference is the launch + drain overhead of a single-wave kernel)."""
import csv
import re
import sys

COLS = [("duration_us", "gpu__time_duration.sum", 1.0),
        ("dram_read_MB", "dram__bytes_read.sum", 1.0),
        ("dram_write_MB", "dram__bytes_write.sum", 1.0),
        ("dram_pct_of_peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("regs_per_thread", "launch__registers_per_thread", 1.0),
        ("grid", "launch__grid_size", 1.0),
        ("waves_per_sm", "launch__waves_per_multiprocessor", 1.0),
        ("l2_to_sm_read_MB", "lts__t_sectors_srcunit_tex_op_read.sum", 32e-6),
        ("sm_cycles_active_max", "sm__cycles_active.max", 1.0),
        ("sm_cycles_elapsed_max", "sm__cycles_elapsed.max", 1.0)]


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("void ", "")
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*", "", name)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
    print("kernel," + ",".join(c[0] for c in COLS))
    for r in rows:
        out = [short(r[ki])]
        for _, metric, scale in COLS:
            if metric not in hdr:
                out.append("")
                continue
            i = hdr.index(metric)
            try:
                v = float(r[i].replace(",", "")) * scale
                u = units[i]
                if metric.startswith("dram__bytes") and u == "byte":
                    v /= 1e6
                elif metric.startswith("dram__bytes") and u == "Kbyte":
                    v /= 1e3
                elif metric.startswith("dram__bytes") and u == "Gbyte":
                    v *= 1e3
                elif metric == "gpu__time_duration.sum" and u in ("ns", "nsecond"):
                    v /= 1e3
                out.append(f"{v:.3f}".rstrip("0").rstrip("."))
            except ValueError:
                out.append(r[i])
        print(",".join(f'"{o}"' if "," in o else o for o in out))


if __name__ == "__main__":
    main()
