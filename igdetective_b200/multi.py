"""The de-novo stage sharded over the ranks of a torchrun launch (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m igdetective_b200.multi -i genome.fasta -o out_dir -l IGH

Contigs -- cut into ~32 Mb pieces with 512 bases of context when longer than that, so that one long scaffold
does not bound a shard -- are assigned by longest-processing-time bin packing (shard.plan_pieces); every rank
scans and scores its own pieces with no data-path collective; rank 0 gathers the row lists, orders them as a
single-process run would and writes the same genes_{V,D,J}.tsv.  The merged files do not depend on the number
of ranks.
"""
import argparse
import os

import numpy as np

from . import datafiles as dfiles
from . import igdetective, shard
from .engine import Engine
from .fasta import NativeFasta, fasta_dict
from .rss import D, J, V


def detect_pieces(eng, seqs, ids, pieces, canonical, order="canonical", part=None):
    """igdetective.detect on this rank's pieces (shard.plan_pieces): each piece is loaded as a contig of its own,
    the rows it owns are kept and moved to contig coordinates.  seqs: {contig id: str-like}.  `part`: {piece
    number: str-like} already loaded on the engine by the caller (byte buffers in pinned memory)."""
    if not pieces:
        return {V: [], D: [], J: []}
    load = part is None
    if load:
        part = {}
        for k, (ci, a, b, p0, p1) in enumerate(pieces):
            s = seqs[ids[ci]]
            part[k] = s if (p0 == 0 and p1 == len(s)) else s[p0:p1]
    _, rows = igdetective.detect(eng, part, canonical, order, load=load)
    out = {V: [], D: [], J: []}
    by_piece = {g: {} for g in out}
    for g in out:
        for r in rows[g]:
            by_piece[g].setdefault(r[0], []).append(r)
    for k, piece in enumerate(pieces):
        own = shard.own_rows({g: by_piece[g].get(k, []) for g in out}, piece, ids[piece[0]])
        for g in out:
            out[g].extend(own[g])
    return out


def detect_pieces_pipelined(engines, ids, pieces, canonical, part, buffer, groups=4, timings=None):
    """load + detect_pieces of this rank's pieces in `groups` runs of consecutive pieces over len(engines) handles of the
    same GPU (one host thread each): the upload of group k + 1 (H2D + pack, one at a time, in order) runs while group k is
    scanned and scored on another handle's stream.  buffer = (address of the pieces' letters in pinned host memory,
    uint64 offsets of the pieces in it); part = {piece number: str-like view of its letters}.  timings (optional list)
    receives (group, upload seconds, scan + scoring seconds, scan kernel ms, pack kernel ms, bases, start of the upload
    in seconds since the call, scoring kernel ms)."""
    import threading
    import time
    ptr, off = buffer
    n = len(pieces)
    if n == 0:
        return {V: [], D: [], J: []}
    groups = max(1, min(groups, n))
    cuts = sorted(set([0, n] + [int(np.searchsorted(off, int(off[n]) * g // groups)) for g in range(1, groups)]))
    n_groups = len(cuts) - 1
    turn = [threading.Event() for _ in range(n_groups + 1)]
    turn[0].set()
    results, errors = [None] * n_groups, []
    t_begin = time.perf_counter()

    def worker(e):
        try:
            for k in range(e, n_groups, len(engines)):
                lo, hi = cuts[k], cuts[k + 1]
                turn[k].wait()
                t0 = time.perf_counter()
                try:
                    engines[e].load_contigs((ptr + int(off[lo]), (off[lo:hi + 1] - off[lo]).astype(np.uint64)), list(range(hi - lo)),
                                            wait=False)           # the next upload may start: only the bus is taken in turns
                finally:
                    turn[k + 1].set()
                t1 = time.perf_counter()
                results[k] = detect_pieces(engines[e], None, ids, pieces[lo:hi], canonical, "canonical",
                                           part={j - lo: part[j] for j in range(lo, hi)})
                engines[e].sync()
                if timings is not None:
                    tk = engines[e].timing()
                    timings.append((k, t1 - t0, time.perf_counter() - t1, tk["scan_ms"], tk["pack_ms"], int(off[hi] - off[lo]),
                                    t0 - t_begin, tk["align_ms"]))
        except BaseException as exc:                       # surfaced in the caller's thread
            errors.append(exc)
            for ev in turn:
                ev.set()

    threads = [threading.Thread(target=worker, args=(e,)) for e in range(min(len(engines), n_groups))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    out = {V: [], D: [], J: []}
    for r in results:
        for g in out:
            out[g].extend(r[g])
    return out


def detect_byte_range(eng, input_path, rank, world, canonical, group=None, margin=65536):
    """This rank's 1/world of a plain FASTA file: it parses only the lines of its byte range (plus a margin of context
    either side, igd_fasta_open_range), the ranks exchange what their ranges hold (shard.range_summary: a few KB,
    one all_gather_object) to agree on the contig table, and every record of the range is loaded as a contig of its
    own whose rows inside the range are kept.  No rank reads the whole file; the shards are balanced by bytes.
    Returns (rows in contig coordinates, contig ids of the whole file)."""
    import torch.distributed as dist
    size = os.path.getsize(input_path)
    fa = NativeFasta(input_path, byte_range=(size * rank // world, size * (rank + 1) // world), margin=margin)
    mine = shard.range_summary(fa)
    if world > 1:
        summaries = [None] * world
        dist.all_gather_object(summaries, mine, group=group)
    else:
        summaries = [mine]
    ids, lens, starts = shard.contig_table(summaries)
    pieces = shard.range_pieces(fa, ids, lens, starts[rank])
    part = {k: fa.view(k) for k in range(len(pieces))}
    eng.load_fasta(fa)
    return detect_pieces(eng, None, ids, pieces, canonical, "canonical", part=part), ids


def run_sharded(input_path, output_path, locus="IGH", datafiles=None, gene_dir="combined_reference_genes",
                order="canonical", device=None, rank=None, world=None, group=None, tile=32_000_000, byte_ranges=None):
    """byte_ranges: every rank parses only its 1/world of the file (detect_byte_range) instead of the whole file and
    its share of the pieces; default: on for plain FASTA files when there is more than one rank."""
    import torch.distributed as dist
    if rank is None:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    if order != "canonical":
        raise ValueError("sharded runs emit the canonical row order (the reference's order is a property of one "
                         "process's set iteration; use igdetective.run for byte-identical files)")
    if byte_ranges is None:
        byte_ranges = world > 1 and not str(input_path).endswith(".gz")
    canonical = {g: dfiles.load_canonical_genes(locus, g, gene_dir, datafiles) for g in (V, J)}
    with Engine(device, dfiles.load_motifs(datafiles)) as eng:
        if byte_ranges:
            rows, ids = detect_byte_range(eng, input_path, rank, world, canonical, group)
        else:
            seqs = fasta_dict(input_path) if str(input_path).endswith(".gz") else NativeFasta(input_path)
            ids = list(seqs.keys())
            mine = shard.plan_pieces([len(seqs[c]) for c in ids], world, tile)[rank]
            rows = detect_pieces(eng, seqs, ids, mine, canonical, order)
    merged = shard.gather_rows(rows, ids, group) if world > 1 else shard.merge_rows([rows], ids)
    if merged is not None:
        os.makedirs(output_path, exist_ok=True)
        for gene in (V, D, J):
            igdetective._write_tsv("{}/genes_{}.tsv".format(output_path, gene),
                                   [igdetective.D_HEADER if gene == D else igdetective.VJ_HEADER] + merged.get(gene, []))
    return merged


def main(argv=None):
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-i", "--input_file", required=True)
    ap.add_argument("-o", "--output_directory", required=True)
    ap.add_argument("-l", "--locus", default="IGH", choices=["IGH", "IGK", "IGL", "TRA", "TRB", "TRG"])
    ap.add_argument("--datafiles")
    ap.add_argument("--gene_dir", default="combined_reference_genes")
    ap.add_argument("--order", default="canonical", choices=["canonical"])
    ap.add_argument("--tile", type=int, default=32_000_000, help="contigs longer than 1.5 x this are cut into pieces")
    ap.add_argument("--whole-file", action="store_true", help="every rank parses the whole FASTA and takes its share of the "
                    "pieces (LPT) instead of parsing only its byte range of the file")
    a = ap.parse_args(argv)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        run_sharded(a.input_file, a.output_directory, a.locus, a.datafiles, a.gene_dir, a.order, local, tile=a.tile,
                    byte_ranges=False if a.whole_file else None)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
