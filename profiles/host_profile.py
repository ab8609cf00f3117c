"""cProfile of one resident step of the bench workload (host glue hot spots)."""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from igdetective_b200 import igdetective  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402

motifs, canon = bench.load_data()
contig, _ = bench.make_workload(2, int(100e6), motifs, canon)
seqs = {"c": contig.tobytes().decode("latin-1")}
eng = Engine(0, motifs)
eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])
for _ in range(2):
    igdetective.detect(eng, seqs, canon, order="canonical", load=False)
pr = cProfile.Profile()
pr.enable()
igdetective.detect(eng, seqs, canon, order="canonical", load=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
