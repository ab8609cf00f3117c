"""Per-warp timeline of a 100 Mb scan: run with a library built with -DSCAN_TIMELINE
(python profiles/build_variant.py tl -DSCAN_TIMELINE; IGD_B200_LIB=variants/libigd_tl.so python profiles/scan_timeline.py > log;
python profiles/scan_timeline_summary.py log).  Every warp prints: CTA, warp, start, first data, end (globaltimer ns), strips."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
