#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel out of an .ncu-rep (read here, no GPU):
python scripts/ncu_lines.py file.ncu-rep kernel_regex [top_n] [launch_skip]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    fname = ""
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            i_s, i_n, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr and len(r) > 8 and r[0] != "":
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(r[i_s]), int(r[i_n]), int(r[i_t])))
            except ValueError:
                pass
    ts = sum(x[3] for x in lines) or 1
    tn = sum(x[4] for x in lines) or 1
    print("total samples %d, warp instructions %d" % (ts, tn))
    for f, ln, src, s_, n_, t_ in sorted(lines, key=lambda x: -x[3])[:top]:
        print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s" % (100.0 * s_ / ts, 100.0 * n_ / tn, t_ / max(n_, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main()
