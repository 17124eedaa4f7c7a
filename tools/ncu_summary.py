#!/usr/bin/env python3
"""Text summary of an ncu report (.ncu-rep) for profiles/: key metrics per kernel, executed-opcode
mix (pipe shares) and the hottest source lines.  Usage: ncu_summary.py report.ncu-rep [cubin-dir]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]
ALU = {"LOP3", "SHF", "PRMT", "IADD3", "VIADD", "ISETP", "SEL", "VIMNMX", "IABS", "LEA", "MOV", "VIADDMNMX",
       "IMNMX", "SGXT", "BMSK", "VIMNMX3", "PLOP3", "UMOV"}
FMA = {"IMAD", "FMUL", "FFMA"}


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    cubins = glob.glob(os.path.join(sys.argv[2], "*.cubin")) if len(sys.argv) > 2 else []
    rows = list(csv.reader(ncu(rep, "--page", "raw").splitlines()))
    hdr, units = rows[0], rows[1]
    names = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        names.append(name)
        print("== kernel", name)
        for k in KEYS:
            if k in hdr:
                print("  %-86s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for name in dict.fromkeys(names):
        base = name.split("(")[0].split("<")[0].replace("void ", "").strip()
        src = list(csv.reader(ncu(rep, "--page", "source", "--kernel-name", "regex:" + base).splitlines()))
        if len(src) < 3:
            continue
        # one listing per matching launch: keep the first
        for j in range(2, len(src)):
            if src[j] == src[1]:
                src = src[:j - 1]
                break
        h = src[1]
        isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        mix = collections.Counter()
        for r in src[2:]:
            op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
            mix[op] += int(r[iex])
        tot = sum(mix.values()) or 1
        a = sum(v for k, v in mix.items() if k in ALU)
        f = sum(v for k, v in mix.items() if k in FMA)
        print("== opcode mix", base, "(warp instructions executed: %d; ALU pipe %.1f%%, FMA pipe %.1f%%)" % (tot, 100 * a / tot, 100 * f / tot))
        print("  " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in mix.most_common(16)))
        # source lines through the cubin's line table
        for cb in cubins:
            dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
            # one instruction sequence per function whose name matches (template instantiations each
            # have their own); the one as long as the report's listing is the kernel that ran
            seqs, cur, fn = collections.defaultdict(list), None, None
            for l in dis.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", l)
                if m:
                    fn = m.group(1) if base in m.group(1) else None
                    continue
                if fn is None:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
                    seqs[fn].append(cur)
            seq = next((v for v in seqs.values() if len(v) == len(src) - 2), None)
            if seq is None:
                continue
            agg = collections.defaultdict(lambda: [0, 0])
            for ln, r in zip(seq, src[2:]):
                agg[ln][0] += int(r[iex])
                agg[ln][1] += int(r[ism])
            ts = sum(v[1] for v in agg.values()) or 1
            print("== hottest source lines", base, "(share of executed instructions | of stall samples)")
            for ln, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / ts))[:24]:
                print("  %-28s inst %5.1f%%  samples %5.1f%%" % ("%s:%d" % ln if ln else "?", 100 * v[0] / tot, 100 * v[1] / ts))
            break


if __name__ == "__main__":
    main()
