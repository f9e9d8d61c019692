#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel family.
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        elif r["Metric Unit"] == "ms":
            ns *= 1e6
        rows.append((r["Kernel Name"], ns))


def family(name):
    name = re.sub(r"^void ", "", name)
    base = re.split(r"[<(]", name)[0]
    # every kernel of the library lives in namespace b200vc (ncu drops the outer namespace of nested ones: tc::, wt::, w3::)
    if base.startswith("b200vc::") or re.match(r"(tc|tc192|wt|w2|w3|wf)::", base):
        return "b200vc::" + base.replace("b200vc::", "")
    m = re.match(r"((?:at::native::|at::|cudnn::|cutlass::)?[\w:]*?\w+)[<(]", name)
    base = m.group(1) if m else name.split("(")[0]
    if base.startswith("at::"):
        return "at:: (torch elementwise / copy / pooling / upsample)"
    return base


tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    a = agg[family(k)]
    a[0] += 1
    a[1] += ns
print(f"{len(rows)} launches, {tot / 1e6:.1f} ms of device time.\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.2f} % |")
mine = sum(ns for k, ns in rows if family(k).startswith("b200vc::"))
print(f"\nb200vc kernels: {mine / 1e6:.2f} ms = **{100 * mine / tot:.2f} %** of the step.")
