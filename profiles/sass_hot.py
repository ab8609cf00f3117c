"""Summarise `ncu --page source --csv` output: total executed warp-instructions, opcode mix and the
hottest SASS lines.  usage: python profiles/sass_hot.py <csv> [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in body)
smp = sum(int(r[ismp] or 0) for r in body)
print("SASS instructions: %d   executed warp-instr: %d   samples: %d" % (len(body), tot, smp))
ops = Counter()
for r in body:
    op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
    ops[op.split(".")[0]] += int(r[iex])
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(14)))
print("hottest lines (by samples):")
for r in sorted(body, key=lambda r: -int(r[ismp] or 0))[:top]:
    print("  %6s smp %9s exec  %s" % (r[ismp], r[iex], r[isrc][:90]))
