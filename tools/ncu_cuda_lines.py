#!/usr/bin/env python3
"""Per-CUDA-line table of one kernel from an ncu report taken with --import-source on: ncu's own attribution
(innermost source line) of executed warp instructions and stall samples.
Usage: ncu_cuda_lines.py report.ncu-rep kernel-regex [--min PCT] [--nth N]"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    mn = float(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0.5
    nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 0
    out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "sass,cuda",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path"] + [len(rows)]
    rows = rows[starts[nth]:starts[nth + 1]]
    print("#", rows[1][1][:120])
    h = rows[2]
    iex, ism = h.index("Instructions Executed"), h.index("# Samples")
    num = lambda v: int(v) if v.isdigit() else 0
    lines = [(int(r[0]), r[1], num(r[iex]), num(r[ism])) for r in rows[3:] if r and r[0].isdigit()]
    ti, ts = sum(l[2] for l in lines), sum(l[3] for l in lines)
    print("# %d warp instructions, %d samples" % (ti, ts))
    for ln, src, ex, sm in lines:
        if 100.0 * ex / ti >= mn or 100.0 * sm / ts >= mn:
            print("%5d  inst %6.2f%%  samples %6.2f%%  %s" % (ln, 100.0 * ex / ti, 100.0 * sm / ts, src.strip()[:110]))


if __name__ == "__main__":
    main()
