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


def run_sharded(input_path, output_path, locus="IGH", datafiles=None, gene_dir="combined_reference_genes",
                order="canonical", device=None, rank=None, world=None, group=None, tile=32_000_000):
    import torch.distributed as dist
    if rank is None:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    if order != "canonical":
        raise ValueError("sharded runs emit the canonical row order (the reference's order is a property of one "
                         "process's set iteration; use igdetective.run for byte-identical files)")
    seqs = fasta_dict(input_path) if str(input_path).endswith(".gz") else NativeFasta(input_path)
    ids = list(seqs.keys())
    mine = shard.plan_pieces([len(seqs[c]) for c in ids], world, tile)[rank]
    canonical = {g: dfiles.load_canonical_genes(locus, g, gene_dir, datafiles) for g in (V, J)}
    with Engine(device, dfiles.load_motifs(datafiles)) as eng:
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
    a = ap.parse_args(argv)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        run_sharded(a.input_file, a.output_directory, a.locus, a.datafiles, a.gene_dir, a.order, local, tile=a.tile)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
