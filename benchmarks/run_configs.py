#!/usr/bin/env python
"""Measurements on BASELINE.json configs 3-5 (the bench.py line covers config 2; config 1 is the parity
suite).  Prints one JSON line per config; run on a B200 box:

    python benchmarks/run_configs.py [--configs 3,4,5] [--scale 1.0] [--device 0]

config 3: ~3 Gb mammalian-shaped assembly (40 SUPER_* scaffolds + 460 unplaced), scanned whole, V/J/D
          candidates scored vs combined IGHV/IGHJ -- "scan + score a 3 Gb assembly"
config 4: iterative-mode scoring: candidates (350-nt V flanks / 70-nt J flanks / 800-nt raw-case windows)
          against ALL V (1172) / J (235) genes of combined_reference_genes, both orientations
config 5: fragmented assembly: 200 k contigs, Pareto lengths 1 kb-5 Mb
--scale shrinks the sizes proportionally (1.0 = the sizes of SURVEY.md 8(d)).
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
import synth  # noqa: E402
from igdetective_b200 import datafiles, igdetective  # noqa: E402
from igdetective_b200.engine import Engine  # noqa: E402
from igdetective_b200.fasta import reverse_complement  # noqa: E402
from igdetective_b200.rss import SPACER_LENGTH  # noqa: E402
from igdetective_b200.scoring import align_fragment_to_genes, window_alignments  # noqa: E402

DATA = os.path.join(ROOT, "tests", "golden", "datafiles")


class LazySeq:
    """str-like view of a uint8 contig: slicing (Python semantics) returns a str, nothing else is copied."""

    def __init__(self, a):
        self.a = a

    def __len__(self):
        return len(self.a)

    def __getitem__(self, s):
        return self.a[s].tobytes().decode("latin-1")


def genes(kind):
    out = {}
    d = os.path.join(DATA, "combined_reference_genes")
    for fn in sorted(os.listdir(d)):
        if fn[3] == kind:
            from igdetective_b200.fasta import read_fasta
            for rid, s in read_fasta(os.path.join(d, fn)):
                if 0 < len(s) <= 511:
                    out[fn[:4] + "|" + rid] = s.upper()
    return out


def big_assembly(seed, lengths, motifs, canon, clusters, n_runs, only=None):
    """Contig i depends only on (seed, i), so a rank can build just the contigs it owns (`only`)."""
    contigs, truth = {}, []
    cl = {ci: (nv, nd, nj) for ci, nv, nd, nj in clusters}
    for i, L in enumerate(lengths):
        if only is not None and i not in only:
            continue
        rng = np.random.default_rng([seed, i])
        buf = np.empty(L, np.uint8)
        for p in range(0, L, 50_000_000):                     # blockwise: rng.choice makes int64 temporaries
            n = min(50_000_000, L - p)
            buf[p:p + n] = synth.background(rng, n)
        if i in cl:
            lo = int(rng.integers(0, max(1, L - 2_000_000)))
            truth += synth.plant_locus(rng, buf, motifs, canon["V"], canon["J"], *cl[i], lo, min(L, lo + 2_000_000))
        if i < 40:
            for _ in range(n_runs):
                n = int(rng.integers(100, 50_000))
                p = int(rng.integers(0, max(1, L - n)))
                buf[p:p + n] = ord("N")
        contigs[i] = buf
    return contigs, truth


def scan_and_score(eng, contigs, ids, canon, label, extra):
    n_bases = int(sum(len(c) for c in contigs))
    buf = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    off = np.zeros(len(contigs) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in contigs], dtype=np.uint64)
    seqs = {i: LazySeq(c) for i, c in zip(ids, contigs)}
    t0 = time.perf_counter()
    eng.load_contigs((buf, off), ids)
    t_load = time.perf_counter() - t0
    pack_ms = eng.timing()["pack_ms"]
    times = []
    for _ in range(3):
        eng.scan_rss_count(SPACER_LENGTH)
        times.append(eng.timing()["scan_ms"])
    scan_ms = float(np.median(times))
    t0 = time.perf_counter()
    info, rows = igdetective.detect(eng, seqs, canon, order="canonical", load=False)
    t_detect = time.perf_counter() - t0
    out = {"config": label, "bases": n_bases, "contigs": len(contigs),
           "load_h2d_pack_s": t_load, "pack_kernel_ms": pack_ms, "scan_kernel_ms": scan_ms,
           "scan_gb_per_s_kernel": n_bases / scan_ms / 1e6, "scan_hbm_gbs": n_bases / 4 / scan_ms / 1e6,
           "scan_plus_scoring_s": t_detect, "whole_job_s": t_load + t_detect,
           "rss": {st: sum(len(v) for sd in info[st] for v in info[st][sd].values()) for st in info},
           "rows": {g: len(r) for g, r in rows.items()}}
    out.update(extra)
    print(json.dumps(out), flush=True)


def config3(eng, motifs, canon, scale):
    rng = np.random.default_rng(3)
    sup = np.exp(rng.uniform(np.log(20e6), np.log(250e6), 40))
    sup = (sup / sup.sum() * 2.9e9 * scale).astype(np.int64)
    unp = np.exp(rng.uniform(np.log(10e3), np.log(2e6), 460)).astype(np.int64)
    lengths = sup.tolist() + (unp * min(1.0, scale * 4)).astype(np.int64).tolist()
    ids = ["SUPER_%d" % (i + 1) for i in range(40)] + ["scaffold_%d" % i for i in range(460)]
    clusters = [(i, 60, 6, 3) for i in (2, 7, 11, 23, 31)]
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        return config3_sharded(eng, motifs, canon, lengths, ids, clusters)
    t0 = time.perf_counter()
    contigs, truth = big_assembly(3, lengths, motifs, canon, clusters, 3)
    scan_and_score(eng, [contigs[i] for i in range(len(lengths))], ids, canon,
                   "3: synthetic mammalian-shaped assembly, whole-genome scan + scoring",
                   {"generate_s": time.perf_counter() - t0, "planted": len(truth)})


def config3_sharded(eng, motifs, canon, lengths, ids, clusters):
    """The same assembly sharded by contig over the ranks of a torchrun launch (LPT on length): every
    rank loads, scans and scores its own contigs; rank 0 gathers the row lists.  No data-path collective."""
    import torch
    import torch.distributed as dist
    from igdetective_b200 import shard
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard.partition_contigs(lengths, world)[rank]
    t0 = time.perf_counter()
    contigs, truth = big_assembly(3, lengths, motifs, canon, clusters, 3, only=set(mine))
    t_gen = time.perf_counter() - t0
    local = [contigs[i] for i in mine]
    buf = np.concatenate(local) if len(local) > 1 else local[0]
    off = np.zeros(len(local) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in local], dtype=np.uint64)
    seqs = {ids[i]: LazySeq(contigs[i]) for i in mine}
    shard.gather_rows({"warmup": []}, ids)                 # first object collective pays communicator set-up
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.load_contigs((buf, off), [ids[i] for i in mine])
    t_load = time.perf_counter() - t0
    info, rows = igdetective.detect(eng, seqs, canon, order="canonical", load=False)
    t_job = time.perf_counter() - t0
    scan_ms = eng.timing()["scan_ms"]
    merged = shard.gather_rows(rows, ids)
    t_all = time.perf_counter() - t0
    stats = torch.tensor([t_job, t_load, scan_ms, float(len(buf)), t_all], dtype=torch.float64, device="cuda")
    mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"config": "3 sharded by contig over %d GPUs (LPT), whole-genome scan + scoring" % world,
                          "bases": int(sm[3].item()), "contigs": len(lengths), "n_gpus": world,
                          "whole_job_s_max_over_ranks": mx[0].item(), "incl_gather_s": mx[4].item(),
                          "load_h2d_pack_s_max": mx[1].item(), "scan_kernel_ms_max": mx[2].item(),
                          "max_shard_bases": int(mx[3].item()), "generate_s": t_gen,
                          "rows": {g: len(r) for g, r in merged.items()}}), flush=True)


def config5(eng, motifs, canon, scale):
    rng = np.random.default_rng(5)
    n = int(200_000 * scale)
    lengths = np.minimum((1000 * (1 - rng.random(n)) ** (-1 / 1.1)).astype(np.int64), 5_000_000)
    t0 = time.perf_counter()
    contigs = [synth.background(rng, int(L)) for L in lengths]
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in synth.SIG}
    for c in contigs[::20]:                                     # RSS at contig edges in 5 % of contigs
        u = np.frombuffer(synth.rss_unit(rng, motifs, synth.SIG[rng.integers(4)], sets).encode(), np.uint8)
        if len(c) >= len(u):
            p = 0 if rng.integers(2) else len(c) - len(u)
            c[p:p + len(u)] = u
    ids = ["ctg%06d" % i for i in range(n)]
    scan_and_score(eng, contigs, ids, canon, "5: fragmented assembly, skewed contig lengths",
                   {"generate_s": time.perf_counter() - t0})


def config4(eng, canon, scale):
    rng = np.random.default_rng(4)
    gv, gj = genes("V"), genes("J")
    vlist, jlist = list(gv.values()), list(gj.values())
    n = int(100_000 * scale)
    nv, nj, nw = n // 4, n // 4, n // 2

    def bg(k):
        return synth.background(rng, k).tobytes().decode()
    res = {"config": "4: iterative-mode scoring vs all V (%d) / J (%d) genes of combined_reference_genes" % (len(vlist), len(jlist)),
           "candidates": {"V_flank_350": nv, "J_flank_70": nj, "window_800": nw}}
    vf = [(bg(350) + synth.mutate(rng, vlist[rng.integers(len(vlist))], 0.1, 0.01))[-350:] if rng.random() < 0.3 else bg(350)
          for _ in range(nv)]
    jf = [(synth.mutate(rng, jlist[rng.integers(len(jlist))], 0.1, 0.01) + bg(70))[:70] if rng.random() < 0.3 else bg(70)
          for _ in range(nj)]
    wf = []
    for i in range(nw):
        w = (bg(300) + synth.mutate(rng, vlist[rng.integers(len(vlist))], 0.1, 0.01) + bg(500))[:800] if rng.random() < 0.3 else bg(800)
        if i % 3 == 0:
            w = w[:250] + w[250:600].lower() + w[600:]          # soft-masked stretch
        wf.append(w)
    # under torchrun every rank builds the same candidate lists and scores one contiguous block of each,
    # cut at equal DP cells (shard.partition_by_cost); the small per-candidate results are all-gathered
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist
        from igdetective_b200 import shard
        shard.all_gather_lists([0])                             # first object collective pays communicator set-up
        res["config"] += ", candidates sharded over %d GPUs by DP cells" % world
        res["n_gpus"] = world
    for name, frs, gl, fn in (("V_flank_350", vf, vlist, "A"), ("J_flank_70", jf, jlist, "A"), ("window_800", wf, list(gv.items()), "B")):
        lo, hi = (0, len(frs)) if world == 1 else shard.partition_by_cost([len(f) for f in frs], world)[rank]
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        c0, m0, wall = eng.align_cells_total, eng.align_ms_total, time.perf_counter()
        small = []
        for b0 in range(lo, hi, 512):                           # batches bound the result matrices (~1.2 M pairs)
            chunk = frs[b0:min(hi, b0 + 512)]
            if fn == "A":
                pi, ln = align_fragment_to_genes(chunk, gl, engine=eng)
                best = np.argmax(pi, axis=1)
                small += [(int(b), float(pi[i, b]), float(ln[i, b])) for i, b in enumerate(best)]
            else:
                small += [(al.gene_seq, al.gene_id, al.pi, sd) for al, sd in window_alignments(eng, chunk, gl)]
        if world > 1:
            small = shard.all_gather_lists(small)
        assert len(small) == len(frs)
        wall = time.perf_counter() - wall
        cells, ms = eng.align_cells_total - c0, eng.align_ms_total - m0
        if world > 1:
            mx = torch.tensor([wall, ms], dtype=torch.float64, device="cuda"); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = torch.tensor([float(cells)], dtype=torch.float64, device="cuda"); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            wall, ms, cells = mx[0].item(), mx[1].item(), sm[0].item()
        res[name] = {"cells": int(cells), "kernel_s": ms / 1e3, "gcups_kernel": cells / ms / 1e6 if ms else 0.0,
                     "wall_s": wall, "gcups_wall": cells / wall / 1e9}
    if rank == 0:
        print(json.dumps(res), flush=True)


def ingest(eng, motifs, canon, scale):
    """FASTA ingest ("next" row 1): C++ reader + igd_load_fasta vs the Python reader the reference path implies"""
    import tempfile
    from igdetective_b200.fasta import NativeFasta, read_fasta
    rng = np.random.default_rng(9)
    n = int(1e9 * scale)
    lengths = [n // 2, n // 4, n // 8, n // 8]
    path = os.path.join(tempfile.gettempdir(), "igd_ingest_%d.fa" % os.getpid())
    with open(path, "wb") as f:
        for i, L in enumerate(lengths):
            f.write(b">contig_%d synthetic\n" % i)
            for p in range(0, L, 60_000_000):
                blk = synth.background(rng, min(60_000_000, L - p))
                pad = (-len(blk)) % 60
                rows = np.concatenate([blk, np.full(pad, 32, np.uint8)]).reshape(-1, 60)
                out = np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], axis=1).ravel()
                f.write(out.tobytes())
    size = os.path.getsize(path)
    t0 = time.perf_counter(); nf = NativeFasta(path); t_parse = time.perf_counter() - t0
    t0 = time.perf_counter(); eng.load_fasta(nf); t_load = time.perf_counter() - t0
    hits = eng.scan_rss_count(SPACER_LENGTH)
    t0 = time.perf_counter(); recs = read_fasta(path); t_py = time.perf_counter() - t0
    ok = [r[0] for r in recs] == nf.ids and all(len(r[1]) == L for r, L in zip(recs, nf.lengths)) \
        and recs[3][1][-1000:] == nf[nf.ids[3]][-1000:]
    os.remove(path)
    print(json.dumps({"config": "ingest: %d-byte FASTA, 60-column lines" % size, "bases": int(sum(nf.lengths)),
                      "native_parse_s": t_parse, "native_parse_gb_per_s": size / t_parse / 1e9,
                      "h2d_pack_s": t_load, "python_reader_s": t_py, "python_reader_gb_per_s": size / t_py / 1e9,
                      "same_records": bool(ok), "rss_hits": int(hits)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    motifs = datafiles.load_motifs(DATA)
    canon = {g: datafiles.load_canonical_genes("IGH", g, "combined_reference_genes", DATA) for g in ("V", "J")}
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:       # torchrun: one process per GPU
        import torch
        import torch.distributed as dist
        a.device = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(a.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", a.device))
    eng = Engine(a.device, motifs)
    for c in a.configs.split(","):
        if c == "3":
            config3(eng, motifs, canon, a.scale)
        elif c == "4":
            config4(eng, canon, a.scale)
        elif c == "5":
            config5(eng, motifs, canon, a.scale)
        elif c == "ingest":
            ingest(eng, motifs, canon, a.scale)
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
