#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel family.
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

MINE = ("gdn_tc_kernel", "gdn_fp32_kernel", "gdn_prepare_kernel", "warp_tma_kernel", "warp2_tma_kernel", "warp_kernel",
        "warp2_lhbdc_kernel", "warp2_half_sse_kernel", "warp_sse_kernel", "blend_kernel", "sse_u8_kernel",
        "sum_partials_kernel", "gauss_cond_kernel", "eb_prepare_kernel", "entropy_bottleneck_kernel",
        "spynet_level_kernel", "spynet_pyramid_kernel", "rans_", "deform_conv2d", "dcn_to_group_last_kernel",
        "round_checker_kernel", "checker_mask_kernel")
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
    m = re.match(r"((?:b200vc::)?(?:tc::|wt::)?(?:%s))" % "|".join(MINE), name)
    if m:
        return "b200vc::" + m.group(1).replace("b200vc::", "")
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
