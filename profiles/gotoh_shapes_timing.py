"""Packed Gotoh kernel throughput by target length (= kernel shape): 300 targets x the 490 IGHV genes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bench, synth
from igdetective_b200.engine import Engine
motifs, canon = bench.load_data()
eng = Engine(0, motifs)
rng = np.random.default_rng(0)
genes = list(canon["V"].values())
for L in (350, 400, 500, 800):
    frags = []
    for i in range(300):
        g = genes[rng.integers(len(genes))]
        frags.append((synth.background(rng, L).tobytes().decode() + synth.mutate(rng, g, 0.1, 0.01))[-L:])
    pt, pq = [x.ravel() for x in np.meshgrid(np.arange(len(frags)), np.arange(len(genes)), indexing="ij")]
    for _ in range(3):
        eng.align(frags, genes, pt, pq)
        t = eng.timing()
    print(L, "nt targets:", round(t["align_ms"], 2), "ms", round(t["align_cells"] / t["align_ms"] / 1e6, 1), "GCUPS")
