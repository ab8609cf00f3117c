"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_summary.py <launches.csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    k = r[4][:64]
    n, t = tot.get(k, (0, 0))
    tot[k] = (n + 1, t + int(float(r[14])))
total = sum(t for _, t in tot.values())
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-66s launches=%4d total_ns=%12d share=%5.1f%%" % (k, n, t, 100.0 * t / total))
