#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, mean duration and share per kernel, and the
shares of the four kernels of one timed step. usage: python scripts/launch_summary.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    v = float(r[14].replace(",", ""))
    if r[13] == "ns":
        v /= 1000.0
    tot[name] += v
    cnt[name] += 1
alltime = sum(tot.values())
print("# launch list of `python bench.py --steps 2 --warmup 1 --no-cpu --no-membrane` under ncu (gpu__time_duration.sum, --clock-control none): cold-cache, serialised")
print("# kernel | launches | mean us | share of all GPU time in the run")
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-44s %5d %10.1f %6.1f%%" % (k, cnt[k], tot[k] / cnt[k], 100.0 * tot[k] / alltime))
step = [k for k in tot if k.startswith(("k_gate_rows<1>", "k_cheap_flat", "k_patch_flat", "k_combine_flat"))]
st = sum(tot[k] / cnt[k] for k in step)
print("# one timed step = k_gate_rows<1> + k_cheap_flat + k_patch_flat + k_combine_flat; shares of the step (mean launch durations):")
for k in step:
    print("#   %-32s %6.1f us %5.1f%%" % (k, tot[k] / cnt[k], 100.0 * tot[k] / cnt[k] / st))
