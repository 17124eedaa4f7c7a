#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per (kernel, grid, block) and print each
kernel's share of one device-resident step (the launches whose grid is that of the full batch).
Usage: launch_list.py launches.csv FRAMES > profiles/rN_launches.txt"""
import collections
import csv
import sys

path, frames = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
h = rows[0]
ik, ib, ig, iv = h.index("Kernel Name"), h.index("Block Size"), h.index("Grid Size"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ik].split("(")[0] if not r[ik].startswith("void") else r[ik][5:].split("(")[0], r[ig], r[ib])
    agg.setdefault(key, []).append(float(r[iv]) / 1e3)
print("ncu launch list (gpu__time_duration.sum, --clock-control none) of: python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("per-launch times are serialised and cold-cache; the kernel's SHARE of a step is what must agree with bench.py\n")
print("%-44s %-18s %-14s %5s %12s %12s" % ("kernel", "grid", "block", "n", "mean us", "min us"))
for (k, g, b), v in agg.items():
    print("%-44s %-18s %-14s %5d %12.1f %12.1f" % (k[:44], g, b, len(v), sum(v) / len(v), min(v)))
# the device-resident steps come first in bench.py: everything up to the first launch whose grid does not belong to the
# full batch (the end-to-end leg decodes in chunks) is "steps"; a kernel's share is its time per step there
step, nsteps, seen_full = collections.OrderedDict(), 0, False
for r in rows[1:]:
    k = r[ik][5:].split("(")[0] if r[ik].startswith("void") else r[ik].split("(")[0]
    dims = [int(x) for x in r[ig].strip("()").split(",")]
    if "build_lut" in k:
        continue
    if frames in dims:
        seen_full = True
    elif seen_full and ("scan" in k or "rtj_idct_kernel" in k):
        break                                   # a scan or K2 launch of another size: the chunked leg has begun
    step[k] = step.get(k, 0.0) + float(r[iv]) / 1e3
    if "rtj_idct_hard_kernel" in k:
        nsteps += 1
tot = sum(step.values())
print("\nshare of one device-resident step (%d frames, mean of %d steps):" % (frames, nsteps))
for k, v in step.items():
    print("  %-46s %8.1f us  %5.1f%%" % (k[:46], v / nsteps, 100 * v / tot))
