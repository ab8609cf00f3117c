"""The de-novo stage sharded by contig over the ranks of a torchrun launch (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m igdetective_b200.multi -i genome.fasta -o out_dir -l IGH

Contigs are assigned by longest-processing-time bin packing (shard.partition_contigs); every rank
scans and scores its own contigs with no data-path collective; rank 0 gathers the row lists, orders
them as a single-process run would (strand, then contig order of the file) and writes the same
genes_{V,D,J}.tsv.  The merged files do not depend on the number of ranks.
"""
import argparse
import os

from . import datafiles as dfiles
from . import igdetective, shard
from .engine import Engine
from .fasta import NativeFasta, fasta_dict
from .rss import D, J, V


class _Subset(dict):
    """the rank's share of input_seq_dict, in file order"""


def run_sharded(input_path, output_path, locus="IGH", datafiles=None, gene_dir="combined_reference_genes",
                order="canonical", device=None, rank=None, world=None, group=None):
    import torch.distributed as dist
    if rank is None:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    seqs = fasta_dict(input_path) if str(input_path).endswith(".gz") else NativeFasta(input_path)
    ids = list(seqs.keys())
    mine = shard.partition_contigs([len(seqs[c]) for c in ids], world)[rank]
    part = _Subset((ids[i], seqs[ids[i]]) for i in mine)
    canonical = {g: dfiles.load_canonical_genes(locus, g, gene_dir, datafiles) for g in (V, J)}
    with Engine(device, dfiles.load_motifs(datafiles)) as eng:
        if part:
            _, rows = igdetective.detect(eng, part, canonical, order)
        else:
            rows = {V: [], D: [], J: []}
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
    ap.add_argument("--order", default="canonical", choices=["canonical", "reference"])
    a = ap.parse_args(argv)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        run_sharded(a.input_file, a.output_directory, a.locus, a.datafiles, a.gene_dir, a.order, local)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
