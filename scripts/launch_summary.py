#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list, restricted to
the launches of whole device-timed training steps (k_gather_split<false> ... k_row_update): kernel, launches
per step, mean duration, share of the summed kernel time of a step.

    python scripts/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, ist = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Stream")
launches = [(re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("ocf::", ""), float(r[iv]) / 1e3, r[ist]) for r in rows[1:]]
# device-timed train steps: regather (k_gather_split<0>) ... k_metrics
steps, cur = [], None
for name, us, stream in launches:
    if name.startswith("k_gather_split<0>"):
        cur = []
    if cur is not None:
        cur.append((name, us, stream))
        if name.startswith("k_row_update"):
            steps.append(cur)
            cur = None
tot = collections.defaultdict(float)
cnt = collections.Counter()
for st in steps:
    for name, us, _ in st:
        tot[name] += us
        cnt[name] += 1
n = max(len(steps), 1)
total = sum(tot.values())
print("%d device-timed train steps in the list; summed kernel time per step %.1f us (cold-cache, serialised)" % (len(steps), total / n))
print("%-28s %8s %10s %7s" % ("kernel", "per step", "mean us", "share"))
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-28s %8.1f %10.1f %6.1f%%" % (name, cnt[name] / n, t / cnt[name], 100 * t / total))
other = collections.defaultdict(list)
for name, us, _ in launches:
    if name.startswith(("k_mt_words", "tc::k_score_tc", "k_gather_split<1>")):
        other[name].append(us)
for name, v in other.items():
    print("%-28s launches %3d  mean %8.1f us" % (name, len(v), sum(v) / len(v)))
