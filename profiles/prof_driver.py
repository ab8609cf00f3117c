"""Small driver for ncu captures: runs only the scan kernel or only the Gotoh kernel on the bench
workload.  usage: python profiles/prof_driver.py scan|align|windows|bounds [mbases] [iters]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402
from igdetective_b200.rss import SPACER_LENGTH  # noqa: E402

what = sys.argv[1]
mbases = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
motifs, canon = bench.load_data()
eng = Engine(0, motifs)
if what == "scan":
    contig, _ = bench.make_workload(2, int(mbases * 1e6), motifs, canon)
    eng.load_contigs((contig, np.array([0, len(contig)], np.uint64)), ["c"])
    for _ in range(iters):
        n = eng.scan_rss_count(SPACER_LENGTH)
        print("hits", n, eng.timing()["scan_ms"], "ms")
elif what == "bounds":
    import time
    import synth
    rng = np.random.default_rng(5)
    genes = list(canon["V"].values())
    frags = [synth.background(rng, 350).tobytes().decode() for _ in range(int(mbases))]   # here: number of unrelated 350-nt fragments
    for _ in range(iters):
        eng.sync()
        t0 = time.perf_counter()
        lcs, y = eng.match_bounds(frags, genes)
        dt = time.perf_counter() - t0
        glen = np.array([len(g) for g in genes])
        print("bounds: %d x %d pairs in %.2f ms (both kernels + copies); (LCS - 1) / len mean %.3f, Y / (2 len) mean %.3f max %.3f"
              % (len(frags), len(genes), dt * 1e3, ((lcs - 1) / glen).mean(), (y / (2.0 * glen)).mean(), (y / (2.0 * glen)).max()))
elif what == "windows":
    import synth
    rng = np.random.default_rng(1)
    genes = list(canon["V"].values())
    wins = []
    for i in range(int(mbases)):          # here: number of 800-nt raw-case windows
        g = genes[rng.integers(len(genes))]
        w = (synth.background(rng, 300).tobytes().decode() + synth.mutate(rng, g, 0.1, 0.01) + synth.background(rng, 500).tobytes().decode())[:800]
        if i % 2:
            w = w[:200] + w[200:500].lower() + w[500:]
        wins.append(w)
    pt, pq = [x.ravel() for x in np.meshgrid(np.arange(len(wins)), np.arange(len(genes)), indexing="ij")]
    for detail in (False, True):
        for _ in range(iters):
            eng.align(wins, genes, pt, pq, detail=detail)
            t = eng.timing()
            print("windows detail=%s" % detail, t["align_ms"], "ms", t["align_cells"] / t["align_ms"] / 1e6, "GCUPS", t["align_launches"], "launches")
else:
    import synth
    rng = np.random.default_rng(0)
    genes = list(canon["V"].values())
    frags = []
    for i in range(int(mbases)):          # here: number of 350-nt fragments
        g = genes[rng.integers(len(genes))]
        frags.append((synth.background(rng, 350).tobytes().decode() + synth.mutate(rng, g, 0.1, 0.01))[-350:])
    pt, pq = [x.ravel() for x in np.meshgrid(np.arange(len(frags)), np.arange(len(genes)), indexing="ij")]
    for _ in range(iters):
        eng.align(frags, genes, pt, pq)
        t = eng.timing()
        print("align", t["align_ms"], "ms", t["align_cells"] / t["align_ms"] / 1e6, "GCUPS")
