"""Per-CUDA-source-line totals of one ncu report (needs -lineinfo and --import-source on).
usage: python profiles/src_hot.py <rep> [top]   -- executed warp-instructions and stall samples per source line"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iex, ismp, ithr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > iex and r[0].isdigit() and r[iex].isdigit():
        num = lambda x: int(x) if x.isdigit() else 0
        lines.append((int(r[0]), r[1].strip(), int(r[iex]), num(r[ismp]), num(r[ithr])))
tot = sum(l[2] for l in lines)
smp = sum(l[3] for l in lines)
print("executed warp-instr %d, samples %d" % (tot, smp))
for ln, src, ex, sm, th in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%5d  %5.1f%% instr  %5.1f%% smp  lanes %4.1f  %s" % (ln, 100.0 * ex / tot, 100.0 * sm / max(smp, 1), th / max(ex, 1), src[:100]))
