"""Iterative stage behind the interface of the reference's py/extract_aligned_genes.py:

    main(genome_fasta, gene_fasta, output_dir)        :129-194   -> genes.tsv, genes.fasta
    ComputeAlignment(aligner, query_list, strand_list, gene_seqs) -> (Alignment, strand)   :85-105
    SetupAligner(aligner)   :72-83      BioAlign   :9-56      Alignment   :58-70

The 2 x G global alignments of every 800-nt window run on the GPU (igd_score_windows); minimap2 still
seeds the windows (the north star keeps the prefilter's interface).  The repository's `py/` directory
holds a same-name shim of this module, so the unmodified run_iterative_igdetective.py imports it as
`gene_finding_tools` (:17) exactly like the original."""
import os
import shutil
import subprocess
import sys

import numpy as np

from .datafiles import load_gene_records
from .engine import REFERENCE_SCORING, Engine
from .fasta import NativeFasta, read_fasta, reverse_complement
from .scoring import Alignment, BioAlignView, pi_from, window_alignments

BioAlign = BioAlignView
GENE_LEN = 400                 # half-width of a window and minimal distance between used positions (:156)
TSV_COLUMNS = ('Contig', 'Pos', 'Seq', 'AASeq', 'PI', 'BestHit', 'Productive', 'Strand')     # :157

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


# ---- the aligner object of the reference's seam ---------------------------------------------------

class GpuAligner:
    """Stands where Align.PairwiseAligner() stands in the reference (:153): holds the ten gap / match
    attributes SetupAligner assigns, and the Engine that runs the alignments."""

    _ATTRS = {"match_score": "match", "mismatch_score": "mismatch",
              "target_internal_open_gap_score": "target_internal_open",
              "target_internal_extend_gap_score": "target_internal_extend",
              "target_end_open_gap_score": "target_end_open", "target_end_extend_gap_score": "target_end_extend",
              "query_internal_open_gap_score": "query_internal_open",
              "query_internal_extend_gap_score": "query_internal_extend",
              "query_end_open_gap_score": "query_end_open", "query_end_extend_gap_score": "query_end_extend"}

    def __init__(self, engine=None, device=0):
        self.mode = "global"
        self.match_score, self.mismatch_score = 1, 0                    # PairwiseAligner defaults
        for a in list(self._ATTRS)[2:]:
            setattr(self, a, 0)
        self._engine, self._device, self._own = engine, device, False

    @property
    def engine(self):
        if self._engine is None:
            self._engine, self._own = Engine(self._device), True
        return self._engine

    def close(self):
        if self._own and self._engine is not None:
            self._engine.close()
        self._engine, self._own = None, False


def scoring_of(aligner):
    """The igd_scoring fields of any object carrying PairwiseAligner's attribute names."""
    if getattr(aligner, "mode", "global") != "global":
        raise ValueError("only global mode is implemented")
    sc = {}
    for attr, field in GpuAligner._ATTRS.items():
        v = getattr(aligner, attr)
        if int(v) != v:
            raise ValueError("%s must be an integer" % attr)
        sc[field] = int(v)
    return sc


def SetupAligner(aligner):
    """:72-83, attribute by attribute."""
    aligner.mode = 'global'
    for attr, field in GpuAligner._ATTRS.items():
        setattr(aligner, attr, REFERENCE_SCORING[field])


def ComputeAlignment(aligner, query_list, strand_list, gene_seqs):
    """:85-105 with the reference's signature: every query of `query_list` against every gene, scan order
    queries outer / genes inner, the LAST alignment with maximal PI (>=, from 0) wins; returns
    (Alignment(QuerySeq, gene id, PI), strand) or (Alignment(), '') when nothing has PI >= 0 or the
    winner's QuerySeq is empty.  Genes are records with .seq / .id (or (id, seq) pairs)."""
    engine = aligner.engine if hasattr(aligner, "engine") else GpuAligner().engine
    genes = [(g.id, str(g.seq)) if hasattr(g, "seq") else (g[0], str(g[1])) for g in gene_seqs]
    queries = [str(q) for q in query_list]
    strands = list(strand_list)[:len(queries)]
    queries = queries[:len(strands)]                                      # zip() semantics (:89)
    if not queries or not genes:
        return Alignment(), ''
    r = engine.align_dense(queries, [g[1] for g in genes], scoring_of(aligner))
    pi = pi_from(r["matches"], r["alen"])                                # [Q, G] float64, BioAlign.PI()
    best_pi, best = 0, None
    for k, v in enumerate(pi.ravel().tolist()):
        if v >= best_pi:
            best_pi, best = v, k
    if best is None:
        return Alignment(), ''
    qi, gi = divmod(best, len(genes))
    d = engine.align([queries[qi]], [genes[gi][1]], [0], [0], scoring_of(aligner), detail=True)
    view = BioAlignView(queries[qi], d["matches"][0], d["start"][0], d["end"][0], d["ient"][0])
    if len(view) == 0:                                                   # len(best_alignment) == 0, :101
        return Alignment(), ''
    out = Alignment()
    out.Initiate(view.QuerySeq(), genes[gi][0], best_pi)
    return out, strands[qi]


# ---- files ----------------------------------------------------------------------------------------

def PrepareOutputDir(output_dir):
    """A fresh, empty output directory (:107-110)."""
    shutil.rmtree(output_dir, ignore_errors=True)
    os.makedirs(output_dir)


def sam_positions(sam_file):
    """Columns 3 / 4 of the alignment records of a SAM file -> [(contig id, ascending unique 1-based POS
    as int64 array)], contigs in order of first appearance; unmapped records ('*') are ignored.  The
    content of ProcessSamFile's dict (:112-127), in the shape the batch scorer wants."""
    order, per = [], {}
    with open(sam_file) as f:
        for line in f:
            if line.startswith('@'):
                continue
            cols = line.split()
            if len(cols) < 4 or cols[2] == '*':
                continue
            if cols[2] not in per:
                per[cols[2]] = []
                order.append(cols[2])
            per[cols[2]].append(int(cols[3]))
    return [(c, np.unique(np.asarray(per[c], np.int64))) for c in order]


def ProcessSamFile(sam_file):
    """{contig: [POS, ...]} as the reference returns it (:112-127): first-seen order, duplicates dropped."""
    out = {}
    with open(sam_file) as f:
        for line in f:
            if line.startswith('@'):
                continue
            cols = line.split()
            if len(cols) < 4 or cols[2] == '*':
                continue
            lst = out.setdefault(cols[2], [])
            if int(cols[3]) not in lst:
                lst.append(int(cols[3]))
    return out


def _window(seq, pos, gene_len):
    return seq[max(0, pos - gene_len):min(len(seq), pos + gene_len)]


def find_genes(engine, contig_dict, positions, genes, gene_len=GENE_LEN, batch=256, rank=0, world=1, group=None):
    """:156-184.  `positions`: [(contig id, ascending positions)] (sam_positions) or the reference's dict.
    A position is aligned only if it lies more than `gene_len` behind the last position that produced a
    gene (:161-163), so which positions are needed depends on earlier results.  Waves: the positions the
    rule reaches when every not-yet-scored position is assumed to produce a gene are scored in one GPU
    batch; the rule is replayed with the real results; positions it now reaches that were never scored
    (only after a window came back empty) form the next wave.  Normally one wave: the positions within
    `gene_len` of an accepted one -- most of a real SAM file, where hits cluster on genes -- are never aligned
    (only a window chosen downstream of one that then comes back empty can be scored in vain).
    Returns the table's columns as a dict of lists.
    With world > 1 (one process per GPU) each wave is cut into contiguous blocks of equal DP cells
    (shard.partition_by_cost), every rank scores its block, one all-gather of the small result tuples
    gives every rank the full list: no data-path collective, same output for any world size."""
    from . import shard
    if isinstance(positions, dict):
        positions = [(c, np.unique(np.asarray(p, np.int64))) for c, p in positions.items()]
    glen = sum(len(g[1]) for g in genes)
    scored = {}                                           # (contig, pos) -> (gene_seq, gene_id, pi, strand)

    def replay(collect):
        """The reference's loop; unknown positions are collected and assumed to produce a gene."""
        rows = []
        for c_id, plist in positions:
            prev = -1
            for pos in plist.tolist():
                if pos - prev <= gene_len:
                    continue
                r = scored.get((c_id, pos))
                if r is None:
                    collect.append((c_id, pos))
                    prev = pos
                    continue
                if r[0] == '' or r[0][0] == ' ':          # Alignment.Empty(), :69-70
                    continue
                rows.append((c_id, pos, r))
                prev = pos
        return rows

    while True:
        wave = []
        rows = replay(wave)
        if not wave:
            break
        if world > 1:
            costs = [len(_window(contig_dict[c], p, gene_len)) * glen for c, p in wave]
            lo, hi = shard.partition_by_cost(costs, world)[rank]
        else:
            lo, hi = 0, len(wave)
        mine = []
        for b0 in range(lo, hi, batch):
            chunk = wave[b0:min(hi, b0 + batch)]
            for a, strand in window_alignments(engine, [_window(contig_dict[c], p, gene_len) for c, p in chunk], genes):
                mine.append((a.gene_seq, a.gene_id, a.pi, strand))
        everything = shard.all_gather_lists(mine, group) if world > 1 else mine
        scored.update(zip(wave, everything))
    table = {k: [] for k in TSV_COLUMNS}
    for c_id, pos, (seq, gid, pi, strand) in rows:
        aa = translate(seq)                               # :169
        for k, v in zip(TSV_COLUMNS, (c_id, pos, seq, aa, pi, gid, '*' not in aa, strand)):
            table[k].append(v)
    return table


def write_outputs(table, output_dir):
    """genes.tsv as DataFrame.to_csv(sep='\\t', index=False) writes it (:186-187: LF line ends, float64
    repr for PI, True / False) and genes.fasta (:191-194)."""
    import pandas as pd
    pd.DataFrame(table, columns=list(TSV_COLUMNS)).to_csv(os.path.join(output_dir, 'genes.tsv'), index=False, sep='\t')
    with open(os.path.join(output_dir, 'genes.fasta'), 'w') as fh:
        fh.writelines('>Contig:%s|Pos:%d\n%s\n' % (c, p, s) for c, p, s in zip(table['Contig'], table['Pos'], table['Seq']))


def _dist_rank_world():
    """(rank, world) of a torchrun launch (one process per GPU), (0, 1) otherwise."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def run_minimap2(genome_fasta, gene_fasta, sam_file):
    """`minimap2 -a genome genes -o sam` (:135), found on PATH like the reference finds it."""
    with open(os.devnull, 'w') as null:
        subprocess.call(['minimap2', '-a', str(genome_fasta), str(gene_fasta), '-o', sam_file], stdout=null, stderr=null)


def check_genes(genes, path):
    """The kernels hold a gene in 10-bit fields: say so before minimap2 and the GPU are started."""
    from ._lib import MAX_QUERY_LEN
    for gid, s in genes:
        if len(s) == 0 or len(s) > MAX_QUERY_LEN:
            raise ValueError("%s: gene record %r has %d letters; supported are 1..%d" % (path, gid, len(s), MAX_QUERY_LEN))


def main(genome_fasta, gene_fasta, output_dir, engine=None, sam_file=None, device=0):
    """py/extract_aligned_genes.py:129-194.  Inside an initialised torch.distributed group the windows
    are scored in shards (find_genes) and rank 0 alone runs minimap2 and writes the outputs."""
    rank, world = _dist_rank_world()
    genes = load_gene_records(gene_fasta)
    check_genes(genes, gene_fasta)
    if rank == 0:
        PrepareOutputDir(output_dir)
        if sam_file is None:
            print('Running minimap...')
            print('Alignment of IG genes ' + str(gene_fasta) + ' to ' + str(genome_fasta))
            sam_file = os.path.join(output_dir, 'alignment.sam')
            run_minimap2(genome_fasta, gene_fasta, sam_file)
    if world > 1:
        import torch.distributed as dist
        box = [sam_file]
        dist.broadcast_object_list(box, src=0)      # also the barrier after minimap2
        sam_file = box[0]
    if rank == 0:
        print('Processing SAM file...')
    positions = sam_positions(sam_file)
    if len(positions) == 0:
        if rank == 0:
            print('no matches were found')
        return
    if str(genome_fasta).endswith(".gz"):
        wanted = {c for c, _ in positions}
        contig_dict = {rid: s for rid, s in read_fasta(genome_fasta) if rid in wanted}
    else:                                    # native reader: only the 800-nt windows are ever copied out
        contig_dict = NativeFasta(genome_fasta)
    own = engine is None
    if own:
        engine = Engine(device)
    try:
        table = find_genes(engine, contig_dict, positions, genes, rank=rank, world=world)
    finally:
        if own:
            engine.close()
    if rank != 0:
        return
    write_outputs(table, output_dir)
    print('# detected genes: ' + str(len(table['Pos'])))


if __name__ == '__main__':
    if len(sys.argv) != 4:
        print('Invalid arguments')
        print('python extract_aligned_genes.py genome.fasta reference_IG_genes.fasta output_dir')
        sys.exit(1)
    main(sys.argv[1], sys.argv[2], sys.argv[3])
