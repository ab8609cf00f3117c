"""Wall-clock timeline of one resident step of the bench workload (where the non-kernel time goes)."""
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
for rep in range(3):
    T = [("start", time.perf_counter())]
    mark = lambda n: T.append((n, time.perf_counter()))
    scan = R.RssScan(eng, seqs, order="canonical", load=False); mark("scan+fetch")
    info = {st: {sd: scan.contigwise(st, sd) for sd in "+-"} for st in R.SIGNAL_TYPES}; mark("contigwise x8")
    info["D"] = {sd: R.combine_D_RSS(info["D_left"][sd], info["D_right"][sd], seqs, sd) for sd in "+-"}; mark("combine_D")
    frags = {g: {sd: R.get_s_fragment_from_RSS(g, sd, info, seqs) for sd in "+-"} for g in "VJ"}; mark("fragments")
    drows = I.d_gene_rows(seqs, info["D"]); mark("D rows")
    for g in "VJ":
        ids, glist = list(canon[g].keys()), list(canon[g].values())
        flat = [f for sd in "+-" for c in frags[g][sd] for f in frags[g][sd][c]]
        k0 = eng.align_ms_total
        pi, ln = S.align_fragment_to_genes(flat, glist, engine=eng); mark("%s align_fragment_to_genes (kernel %.2f ms)" % (g, eng.align_ms_total - k0))
        rows = I.vj_gene_rows(eng, seqs, g, info[g], frags[g], canon[g]); mark("%s vj_gene_rows incl. a second align" % g)
    if rep == 2:
        for (a, ta), (b, tb) in zip(T, T[1:]):
            print("%-62s %7.2f ms" % (b, 1e3 * (tb - ta)))
        print("total %.2f ms" % (1e3 * (T[-1][1] - T[0][1])))
