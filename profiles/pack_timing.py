"""pack_kernel on 1 Gb of pinned ASCII: device time and HBM rate (1 B read + 0.375 B written per base)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from igdetective_b200.engine import Engine

motifs, canon = bench.load_data()
eng = Engine(0, motifs)
contig, _ = bench.make_workload(2, 100_000_000, motifs, canon)
big = np.tile(contig, 10)
pin = torch.empty(len(big), dtype=torch.uint8).pin_memory()
pin.numpy()[:] = big
for _ in range(4):
    eng.load_contigs((pin.data_ptr(), np.array([0, len(big)], np.uint64)), ["c"])
    ms = eng.timing()["pack_ms"]
    print("pack_kernel %.3f ms per %d bases = %.0f GB/s of HBM traffic" % (ms, len(big), 1.375 * len(big) / ms / 1e6))
