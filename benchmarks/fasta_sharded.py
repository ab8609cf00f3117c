#!/usr/bin/env python
"""The CLI path of a sharded run from a FASTA FILE on disk: every rank parses its byte range of the file
(igd_fasta_open_range), the ranks agree on the contig table with one all-gather, load, scan, score, and rank 0 gathers
the rows (igdetective_b200.multi.detect_byte_range).  Compares with every rank parsing the whole file (--whole-file).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/fasta_sharded.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from igdetective_b200 import multi, shard  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402
from igdetective_b200.fasta import NativeFasta  # noqa: E402


def write_fasta(path, asm, width=60):
    with open(path, "wb") as f:
        for ci, L in enumerate(asm.lengths):
            f.write((">%s synthetic\n" % asm.ids[ci]).encode())
            for a in range(0, L, 6_000_000):                       # a multiple of the line width
                b = min(L, a + 6_000_000)
                buf = np.empty(b - a, np.uint8)
                asm.fill(ci, a, b, buf)
                full = (b - a) // width * width
                lines = np.empty((full // width, width + 1), np.uint8)
                lines[:, :width] = buf[:full].reshape(-1, width)
                lines[:, width] = 10
                f.write(lines.tobytes())
                if full < b - a:
                    f.write(buf[full:].tobytes() + b"\n")


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--path", default="/tmp/igd_config3.fa")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    motifs, canon = bench.load_data()
    asm = bench.Assembly3(motifs, canon, a.scale)
    if rank == 0 and not os.path.exists(a.path):
        t0 = time.perf_counter()
        write_fasta(a.path, asm)
        print("wrote %s (%.2f GB) in %.1f s" % (a.path, os.path.getsize(a.path) / 1e9, time.perf_counter() - t0), file=sys.stderr)
    eng = Engine(local, motifs)
    shard.gather_rows({"warmup": []}, asm.ids)
    out = {}
    for mode in ("byte_ranges", "whole_file"):
        samples = []
        for it in range(1 + a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "byte_ranges":
                rows, ids = multi.detect_byte_range(eng, a.path, rank, world, canon)
            else:
                seqs = NativeFasta(a.path)
                ids = list(seqs.keys())
                mine = shard.plan_pieces([len(seqs[c]) for c in ids], world)[rank]
                rows = multi.detect_pieces(eng, seqs, ids, mine, canon)
            eng.sync()
            t1 = time.perf_counter()
            merged = shard.gather_rows(rows, ids)
            t2 = time.perf_counter()
            v = [t2 - t0, t1 - t0, t2 - t1]
            if world > 1:
                t = torch.tensor(v, dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                v = t.tolist()
            if it:
                samples.append(v)
        med = np.median(np.array(samples), axis=0).tolist()
        if rank == 0:
            out[mode] = {"whole_job_s": med[0], "parse_load_scan_score_s": med[1], "gather_s": med[2],
                         "rows": {g: len(merged[g]) for g in ("V", "D", "J")}}
    if rank == 0:
        print(json.dumps({"config": "configs[2] from a FASTA file on disk (page cache warm), %d GPUs" % world, "n_gpus": world,
                          "file_bytes": os.path.getsize(a.path), "bases": int(sum(asm.lengths)), **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
