"""Wall-clock timeline of one resident step of the bench workload (where the non-kernel time goes):
detect() with perf_counter wrappers around its stages and around the two C-ABI alignment calls."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from igdetective_b200 import igdetective as I  # noqa: E402
from igdetective_b200 import rss as R  # noqa: E402
from igdetective_b200 import scoring as S  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402

motifs, canon = bench.load_data()
contig, _ = bench.make_workload(2, int(100e6), motifs, canon)
seqs = {"c": contig.tobytes().decode("latin-1")}
eng = Engine(0, motifs)
eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])

acc = {}


def wrap(obj, name, label):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, g)


wrap(eng, "align_dense", "C igd_align_dense")
wrap(eng, "align", "C igd_align_batch (detail)")
wrap(eng, "scan_rss", "C igd_scan + fetch")
wrap(S, "align_fragment_to_genes", "align_fragment_to_genes")
wrap(I, "align_fragment_to_genes", "align_fragment_to_genes")
wrap(I, "compute_alignments", "compute_alignments")
wrap(I, "vj_gene_rows", "vj_gene_rows")
wrap(I, "d_gene_rows", "d_gene_rows")
wrap(I, "combine_D_RSS", "combine_D_RSS")
wrap(I, "get_s_fragment_from_RSS", "get_s_fragment_from_RSS")
wrap(I, "RssScan", "RssScan()")

for rep in range(4):
    acc.clear()
    k0 = eng.align_ms_total
    eng.sync()
    t0 = time.perf_counter()
    I.detect(eng, seqs, canon, order="canonical", load=False)
    eng.sync()
    total = time.perf_counter() - t0
    if rep == 3:
        for k, v in acc.items():
            print("%-34s %7.2f ms" % (k, 1e3 * v))
        print("gotoh kernels (CUDA events)        %7.2f ms" % (eng.align_ms_total - k0))
        print("detect() total                     %7.2f ms" % (1e3 * total))
