#!/usr/bin/env python
"""Per-source-line instruction / stall breakdown of one kernel from an ncu report (dev tool).

usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{rx}", "--launch-count", "1"], capture_output=True, text=True).stdout
    agg = collections.Counter()
    stall = collections.Counter()
    srcs = {}
    f = hdr = None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            print(r[1][:120])
            continue
        if r[0] == "Line No":
            hdr = r
            i_n, i_s = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if r[0] == "" or hdr is None:
            continue
        try:
            ln, n, st = int(r[0]), int(r[i_n]), int(r[i_s])
        except ValueError:
            continue
        agg[(f, ln)] += n
        stall[(f, ln)] += st
        srcs[(f, ln)] = r[1].strip()[:100]
    tot, tst = sum(agg.values()) or 1, sum(stall.values()) or 1
    print(f"warp instructions {tot:,}   stall samples {tst:,}")
    for k, n in agg.most_common(top):
        print(f"{k[0]:16s}{k[1]:5d} {n / 1e6:9.2f}M {100 * n / tot:5.1f}%  stall {100 * stall[k] / tst:5.1f}%  {srcs[k]}")


if __name__ == "__main__":
    main()
