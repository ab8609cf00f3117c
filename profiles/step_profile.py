"""Where the bench step's wall time goes outside the DP kernel: cProfile of K resident steps + the library's own timers.
usage: python profiles/step_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from igdetective_b200 import igdetective
from igdetective_b200.engine import Engine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
motifs, canon = bench.load_data()
contig, _ = bench.make_workload(2, 100_000_000, motifs, canon)
seqs = {"c": contig.tobytes().decode("latin-1")}
eng = Engine(0, motifs)
eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])
for _ in range(3):
    igdetective.detect(eng, seqs, canon, order="canonical", load=False)
t0 = time.perf_counter()
for _ in range(K):
    igdetective.detect(eng, seqs, canon, order="canonical", load=False)
dt = (time.perf_counter() - t0) / K
t = eng.timing()
print("step %.3f ms; align kernels %.3f ms; scan kernel %.3f ms; outside the kernels %.3f ms" % (1e3 * dt, t["align_ms"], t["scan_ms"], 1e3 * dt - t["align_ms"] - t["scan_ms"]))
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    igdetective.detect(eng, seqs, canon, order="canonical", load=False)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
