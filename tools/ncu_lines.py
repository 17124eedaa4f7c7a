#!/usr/bin/env python3
"""Per-source-line table of one kernel from an ncu report: executed warp instructions and stall samples of every
SASS instruction, attributed through the cubin's line table (nvdisasm -gi) to the line of the KERNEL BODY it was
inlined into (and, with --inner, to the innermost line as well).
Usage: ncu_lines.py report.ncu-rep cubin kernel-regex [--inner] [--min PCT] [--nth N]"""
import collections
import csv
import os
import re
import subprocess
import sys


def main():
    rep, cubin, pat = sys.argv[1:4]
    inner = "--inner" in sys.argv
    minpct = float(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    src = list(csv.reader(out.splitlines()))
    # one listing per matching launch: "--nth N" picks one (default the first)
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]
    nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 0
    print("#", src[starts[nth]][1][:110])
    src = src[starts[nth]:starts[nth + 1]]
    h = src[1]
    iex, ism, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    rows = src[2:]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    seqs, fn, grp, fresh = collections.defaultdict(list), None, [], True
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1) if re.search(pat, m.group(1)) else None
            grp = []
            continue
        if fn is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                grp = []
                fresh = False
            grp.append((os.path.basename(m.group(1)), int(m.group(2))))
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            seqs[fn].append(tuple(grp))
            fresh = True
    seq = next((v for v in seqs.values() if len(v) == len(rows)), None)
    if seq is None:
        sys.exit("no function of %d instructions matches %s (have %s)" % (len(rows), pat, {k: len(v) for k, v in seqs.items()}))
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = ts = 0
    for g, r in zip(seq, rows):
        key = (g[-1], g[0]) if inner and g else (g[-1] if g else None)
        ex, sm = int(r[iex]), int(r[ism])
        agg[key][0] += ex
        agg[key][1] += sm
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
        agg[key][2][op] += ex
        tot += ex
        ts += sm
    print("# %s: %d warp instructions, %d samples" % (pat, tot, ts))
    for k in sorted(agg, key=lambda k: (str(k))):
        ex, sm, ops = agg[k]
        if 100.0 * ex / tot < minpct and 100.0 * sm / max(ts, 1) < minpct:
            continue
        name = ("%s:%d <- %s:%d" % (k[0][0], k[0][1], k[1][0], k[1][1])) if inner and k else ("%s:%d" % k if k else "?")
        print("%-44s inst %6.2f%% (%10d)  samples %6.2f%%  %s" % (name, 100.0 * ex / tot, ex, 100.0 * sm / max(ts, 1),
              " ".join("%s:%d" % (o, c * 1000 // max(ex, 1)) for o, c in ops.most_common(4))))


if __name__ == "__main__":
    main()
