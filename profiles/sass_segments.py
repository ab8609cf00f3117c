"""Group `ncu --page source --csv` SASS lines into contiguous segments with similar execution
counts (= loop nests) and print each segment's share.  usage: python profiles/sass_segments.py <csv> [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
hdr = rows[1]
isrc, iex, ith, ismp = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
body = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in body)
smp = sum(int(r[ismp] or 0) for r in body)
segs, cur = [], None
for i, r in enumerate(body):
    e = int(r[iex])
    if cur and abs(e - cur["e"]) <= 0.15 * max(e, cur["e"]):
        cur["n"] += 1; cur["sum"] += e; cur["thr"] += int(r[ith] or 0); cur["end"] = i; cur["smp"] += int(r[ismp] or 0)
    else:
        if cur:
            segs.append(cur)
        cur = {"e": e, "n": 1, "sum": e, "thr": int(r[ith] or 0), "start": i, "end": i, "smp": int(r[ismp] or 0)}
segs.append(cur)
print("SASS lines %d, executed warp-instr %d, samples %d" % (len(body), tot, smp))
for s in sorted(segs, key=lambda s: -s["sum"])[:top]:
    print("lines %5d-%5d n=%4d exec/line=%8d share=%5.1f%% samples=%4.1f%% lanes=%4.1f  %s" % (
        s["start"], s["end"], s["n"], s["e"], 100 * s["sum"] / tot, 100 * s["smp"] / max(smp, 1),
        s["thr"] / max(1, s["sum"]), body[s["start"]][isrc][:40]))
