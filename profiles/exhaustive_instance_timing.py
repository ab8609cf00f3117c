"""Scan kernel time on the 100 Mb bench contig with the shipped motif tables (consensus prefilter) and with one
extra heptamer that breaks the consensus property (exhaustive instance of the same kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bench
from igdetective_b200.engine import Engine
from igdetective_b200.rss import SPACER_LENGTH
motifs, canon = bench.load_data()
rng = np.random.default_rng(1)
custom = {t: {"7": set(motifs[t]["7"]) | {"AAAAAAA"}, "9": set(motifs[t]["9"])} for t in motifs}   # breaks the consensus property
contig, _ = bench.make_workload(2, int(100e6), motifs, canon)
for name, m in (("shipped", motifs), ("shipped + AAAAAAA (exhaustive instance)", custom)):
    eng = Engine(0, m)
    eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])
    for _ in range(3):
        n = eng.scan_rss_count(SPACER_LENGTH)
    print(name, "hits", n, eng.timing()["scan_ms"], "ms")
    eng.close()
