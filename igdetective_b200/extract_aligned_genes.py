"""Iterative stage: same main(genome_fasta, gene_fasta, output_dir) and outputs (genes.tsv,
genes.fasta) as the reference's py/extract_aligned_genes.py:129-194, with the 2 x G alignments of
every window running on the GPU.  minimap2 is still what seeds the windows (north star keeps the
prefilter); `sam_file` lets a caller supply an existing SAM instead of shelling out."""
import os
import shutil
import sys

import pandas as pd

from .datafiles import load_gene_records
from .engine import Engine
from .fasta import NativeFasta, read_fasta
from .scoring import Alignment, window_alignments

_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_CODON = {a + b + c: _AAS[16 * i + 4 * j + k]
          for i, a in enumerate(_BASES) for j, b in enumerate(_BASES) for k, c in enumerate(_BASES)}
_AMBIG = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG",
          "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "ACGT", "N": "ACGT"}


def translate(seq):
    """str(Seq(seq).translate()) (:169): standard table, trailing partial codon dropped."""
    s = seq.upper().replace("U", "T")
    aa = []
    for i in range(0, len(s) - len(s) % 3, 3):
        c = s[i:i + 3]
        x = _CODON.get(c)
        if x is None:
            opts = [_AMBIG[ch] for ch in c]
            xs = {_CODON[a + b + d] for a in opts[0] for b in opts[1] for d in opts[2]}
            x = xs.pop() if len(xs) == 1 else "X"
        aa.append(x)
    return "".join(aa)


def PrepareOutputDir(output_dir):
    if os.path.exists(output_dir):
        shutil.rmtree(output_dir)
    os.mkdir(output_dir)


def ProcessSamFile(sam_file):
    """:112-127 -- {contig: [1-based POS, ...]} in first-seen order, duplicates dropped."""
    position_dict = dict()
    seen = dict()
    for l in open(sam_file).readlines():
        l = l.strip()
        if l[0] == '@':
            continue
        splits = l.split()
        contig_id, pos = splits[2], int(splits[3])
        if contig_id == '*':
            continue
        if contig_id not in position_dict:
            position_dict[contig_id] = []
            seen[contig_id] = set()
        if pos not in seen[contig_id]:
            seen[contig_id].add(pos)
            position_dict[contig_id].append(pos)
    return position_dict


def find_genes(engine, contig_dict, position_dict, genes, gene_len=400, batch=256, rank=0, world=1, group=None):
    """:156-184.  Whether a position is used depends on whether the previous used position produced
    a gene (prev_pos moves only then), so every position's window is scored first (independent units,
    one GPU batch per `batch` windows) and the reference's sequential rule is replayed afterwards.
    With world > 1 (one process per GPU) the windows are split into contiguous blocks of equal DP cells
    (shard.partition_by_cost), every rank scores its block, and one all-gather of the small result
    tuples gives every rank the full list: no data-path collective, same output for any world size."""
    from . import shard
    df = {k: [] for k in ('Contig', 'Pos', 'Seq', 'AASeq', 'PI', 'BestHit', 'Productive', 'Strand')}
    items = [(c_id, p) for c_id in position_dict for p in sorted(position_dict[c_id])]
    glen = sum(len(g[1]) for g in genes)

    def window(c_id, p):
        s = contig_dict[c_id]
        return s[max(0, p - gene_len):min(len(s), p + gene_len)]
    if world > 1:
        costs = [(min(len(contig_dict[c]), p + gene_len) - max(0, p - gene_len)) * glen for c, p in items]
        lo, hi = shard.partition_by_cost(costs, world)[rank]
    else:
        lo, hi = 0, len(items)
    mine = []
    for b0 in range(lo, hi, batch):
        chunk = items[b0:min(hi, b0 + batch)]
        for a, strand in window_alignments(engine, [window(c, p) for c, p in chunk], genes):
            mine.append((a.gene_seq, a.gene_id, a.pi, strand))
    everything = shard.all_gather_lists(mine, group) if world > 1 else mine
    scored = {}
    for (c_id, p), (seq, gid, pi, strand) in zip(items, everything):
        a = Alignment()
        if seq != "":
            a.Initiate(seq, gid, pi)
        scored[(c_id, p)] = (a, strand)
    for c_id in position_dict:
        positions = sorted(position_dict[c_id])
        results = {p: scored[(c_id, p)] for p in positions}
        # ... then replay the reference's sequential skip rule
        prev_pos = -1
        for pos in positions:
            if pos - prev_pos <= gene_len:
                continue
            alignment, strand = results[pos]
            if alignment.Empty():
                continue
            aa_seq = translate(alignment.gene_seq)
            df['Contig'].append(c_id); df['Pos'].append(pos); df['Seq'].append(alignment.gene_seq)
            df['AASeq'].append(aa_seq); df['PI'].append(alignment.pi); df['BestHit'].append(alignment.gene_id)
            df['Productive'].append(aa_seq.find('*') == -1); df['Strand'].append(strand)
            prev_pos = pos
    return df


def _dist_rank_world():
    """(rank, world) of a torchrun launch (one process per GPU), (0, 1) otherwise."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def main(genome_fasta, gene_fasta, output_dir, engine=None, sam_file=None, device=0):
    """py/extract_aligned_genes.py:129-194.  Inside an initialised torch.distributed group the windows
    are scored in shards (find_genes) and rank 0 alone runs minimap2 and writes the outputs."""
    rank, world = _dist_rank_world()
    if rank == 0:
        PrepareOutputDir(output_dir)
        if sam_file is None:
            print('Running minimap...')
            print('Alignment of IG genes ' + gene_fasta + ' to ' + genome_fasta)
            sam_file = os.path.join(output_dir, 'alignment.sam')
            os.system('minimap2 -a ' + genome_fasta + ' ' + gene_fasta + ' -o ' + sam_file + '> /dev/null 2>&1')
    if world > 1:
        import torch.distributed as dist
        box = [sam_file]
        dist.broadcast_object_list(box, src=0)      # also the barrier after minimap2
        sam_file = box[0]
    if rank == 0:
        print('Processing SAM file...')
    position_dict = ProcessSamFile(sam_file)
    if len(position_dict) == 0:
        if rank == 0:
            print('no matches were found')
        return
    if str(genome_fasta).endswith(".gz"):
        contig_dict = {rid: s for rid, s in read_fasta(genome_fasta) if rid in position_dict}
    else:                                    # native reader: only the 800-nt windows are ever copied out
        contig_dict = NativeFasta(genome_fasta)
    genes = load_gene_records(gene_fasta)
    own = engine is None
    if own:
        engine = Engine(device)
    try:
        df = find_genes(engine, contig_dict, position_dict, genes, rank=rank, world=world)
    finally:
        if own:
            engine.close()
    if rank != 0:
        return
    df = pd.DataFrame(df)
    df.to_csv(os.path.join(output_dir, 'genes.tsv'), index=False, sep='\t')
    print('# detected genes: ' + str(len(df)))
    with open(os.path.join(output_dir, 'genes.fasta'), 'w') as fh:
        for i in range(len(df)):
            fh.write('>Contig:' + df['Contig'][i] + '|Pos:' + str(df['Pos'][i]) + '\n' + df['Seq'][i] + '\n')


if __name__ == '__main__':
    if len(sys.argv) != 4:
        print('Invalid arguments')
        print('python -m igdetective_b200.extract_aligned_genes genome.fasta reference_IG_genes.fasta output_dir')
        sys.exit(1)
    main(sys.argv[1], sys.argv[2], sys.argv[3])
