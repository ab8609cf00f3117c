import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, bench
from igdetective_b200.engine import Engine
from igdetective_b200.rss import SPACER_LENGTH
motifs, canon = bench.load_data()
contig, _ = bench.make_workload(2, int(100e6), motifs, canon)
eng = Engine(0, motifs)
eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])
for i in range(3):
    print("RUN", i, flush=True)
    eng.scan_rss_count(SPACER_LENGTH)
    eng.sync()
    print("ms", eng.timing()["scan_ms"], flush=True)
