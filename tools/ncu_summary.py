#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into a markdown table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof_r01c.ncu-rep [...] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp insts"),
]


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("launch__grid_size")],
             "block": r[hdr.index("launch__block_size")]}
        for m, _ in METRICS:
            if m in hdr:
                v = r[hdr.index(m)]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                d[m] = f"{v} {units[hdr.index(m)]}".strip()
        yield d


def main():
    for rep in sys.argv[1:]:
        print(f"## `{rep}`\n")
        print("| kernel | grid x block | " + " | ".join(n for _, n in METRICS) + " |")
        print("|---|---|" + "---:|" * len(METRICS))
        for d in rows_of(rep):
            name = d["kernel"].replace("b200vc::", "").replace("void ", "")
            name = name[:name.index("(")] if "(" in name else name
            print(f"| `{name}` | {d['grid']} x {d['block']} | " + " | ".join(d.get(m, "") for m, _ in METRICS) + " |")
        print()


if __name__ == "__main__":
    main()
