"""cProfile of the host glue on a fragmented assembly (config 5 at 25 % scale)."""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import run_configs as rc  # noqa: E402
import synth  # noqa: E402
from igdetective_b200 import datafiles, igdetective  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402

motifs = datafiles.load_motifs(rc.DATA)
canon = {g: datafiles.load_canonical_genes("IGH", g, "combined_reference_genes", rc.DATA) for g in ("V", "J")}
rng = np.random.default_rng(5)
n = 50_000
lengths = np.minimum((1000 * (1 - rng.random(n)) ** (-1 / 1.1)).astype(np.int64), 5_000_000)
contigs = [synth.background(rng, int(L)) for L in lengths]
ids = ["ctg%06d" % i for i in range(n)]
eng = Engine(0, motifs)
buf = np.concatenate(contigs)
off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lengths, dtype=np.uint64)
eng.load_contigs((buf, off), ids)
seqs = {i: rc.LazySeq(c) for i, c in zip(ids, contigs)}
pr = cProfile.Profile()
pr.enable()
igdetective.detect(eng, seqs, canon, order="canonical", load=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
