"""De-novo stage: same command line and output files as the reference's py/IGDetective.py, with the
RSS scan and the candidate scoring running on the GPU.

    python -m igdetective_b200.igdetective -i genome.fasta -o out_dir -m 1 -l IGH [-r] [-g vdj]

Outputs (byte-identical to the reference run on this repository's BioPython model): genes_V.tsv, genes_D.tsv,
genes_J.tsv, or rss_<type>.csv with -r  (py/IGDetective.py:227-255,424-436,439-487).
Extra long options that the reference does not have: --datafiles DIR, --gene_dir NAME,
--device N, --order reference|canonical.
"""
import csv
import getopt
import os
import sys

import numpy as np

from . import datafiles as dfiles
from .engine import Engine
from .fasta import NativeFasta, fasta_dict, reverse_complement
from .rss import (D, DL, DR, FWD, GENE_LENGTH, J, REV, SIGNAL_TYPES, V, RssScan, combine_D_RSS,
                  get_s_fragment_from_RSS)
from .scoring import align_fragment_to_genes, compute_alignments

PI_CUTOFF = {"strict": {V: 70, J: 70}, "relax": {V: 60, J: 65}}     # py/IGDetective.py:31
MAXK_CUTOFF = {V: 15, J: 11}                                         # :32
ALIGNMENT_EXTENSION = {V: REV, J: FWD, D: None}                      # :30

D_HEADER = ['reference contig', 'strand', 'left heptamer index', 'left nonamer index',
            'left heptamer', 'left nonamer', 'right heptamer index', 'right nonamer index',
            'right heptamer', 'right nonamer', 'start of gene', 'end of gene', 'gene sequence']
VJ_HEADER = ['reference contig', 'strand', 'heptamer index', 'nonamer index', 'heptamer', 'nonamer',
             'start of gene', 'end of gene', 'best aligned human gene', 'alignment direction',
             'alignment PI', 'longest common k-mer', 'gene sequence']
RSS_HEADER = ['7-mer index', '9-mer index', '7-mer', '9-mer', 'reference contig', 'strand']


def _kmer(seq, i, k, strand):
    s = seq[i:i + k]
    return (s if strand == FWD else reverse_complement(s)).upper()


def rss_rows(rss_idx, parent_seq):
    """write_rss_to_file, py/IGDetective.py:227-251."""
    rows = [RSS_HEADER]
    for strand in (FWD, REV):
        for contig in parent_seq:
            for h, n in rss_idx[strand][contig]:
                rows.append([h, n, _kmer(parent_seq[contig], h, 7, strand), _kmer(parent_seq[contig], n, 9, strand),
                             contig, strand])
    return rows


def _write_tsv(path, rows):
    with open(path, "w", newline="") as f:             # csv 'excel' dialect => CRLF, like the reference
        csv.writer(f, delimiter="\t").writerows(rows)


def d_gene_rows(parent_seq, rss_idx):
    """extract_genes for D, py/IGDetective.py:366-378."""
    rows = []
    for strand in rss_idx:
        for contig, entries in rss_idx[strand].items():
            if not entries:
                continue
            s = parent_seq[contig]
            for e in entries:
                lh, ln = _kmer(s, e[0], 7, strand), _kmer(s, e[1], 9, strand)
                rh, rn = _kmer(s, e[2], 7, strand), _kmer(s, e[3], 9, strand)
                if strand == FWD:
                    gs, ge, gene = e[0] + 7, e[2] - 1, s[e[0] + 7:e[2]].upper()
                else:
                    gs, ge, gene = e[2] + 7, e[0] - 1, reverse_complement(s[e[2] + 7:e[0]]).upper()
                rows.append([contig, strand, e[0], e[1], lh, ln, e[2], e[3], rh, rn, gs, ge, gene])
    return rows


def vj_gene_rows(engine, parent_seq, gene, rss_idx, fragments, canonical):
    """Scoring + argmax (:463-477) + extract_genes for V / J (:380-420)."""
    ids = list(canonical.keys())
    glist = list(canonical.values())
    flat = []
    for strand in (FWD, REV):
        for frs in fragments[strand].values():
            if frs:
                flat.extend(frs)
    # the reference runs the same computation twice ('AFFINE' and 'MAXK', :467-468); once is enough
    pi_mat, len_mat = align_fragment_to_genes(flat, glist, "AFFINE", gene, engine=engine)
    winners, k = [], 0
    best_all = np.argmax(pi_mat, axis=1).tolist() if len(flat) else []        # first maximum, like :469
    rows_idx = np.arange(len(best_all))
    pi_best, len_best = pi_mat[rows_idx, best_all], len_mat[rows_idx, best_all]
    for strand in (FWD, REV):
        for contig, frs in fragments[strand].items():
            if not frs:
                continue
            for i, frag in enumerate(frs):
                best, pi, maxk = best_all[k], pi_best[k], len_best[k]
                k += 1
                if pi >= PI_CUTOFF["strict"][gene] or (pi >= PI_CUTOFF["relax"][gene] and maxk >= MAXK_CUTOFF[gene]):
                    winners.append((strand, contig, i, frag, best, pi, maxk))
    rows = []
    if not winners:
        return rows
    detail = compute_alignments(engine, [(w[3], w[4]) for w in winners], glist, want_ient=True)   # :388
    for (strand, contig, i, frag, best, pi, maxk), (aln, direction, _) in zip(winners, detail):
        r = rss_idx[strand][contig][i]
        s = parent_seq[contig]
        start, end = aln.AlignmentRange()
        if strand == FWD:
            if gene == V:
                ge = r[0] - 1
                gs = r[0] - GENE_LENGTH[gene] + start
            else:
                gs = r[0] + 7
                ge = gs + end
        else:
            if gene == V:
                gs = r[0] + 7
                ge = gs + GENE_LENGTH[gene] - start - 1
            else:
                ge = r[0] - 1
                gs = r[0] - end
        rows.append([contig, strand, r[0], r[1], _kmer(s, r[0], 7, strand), _kmer(s, r[1], 9, strand),
                     gs, ge, ids[best], direction, int(round(pi)), maxk, aln.QuerySeq()])
    return rows


class RssInfo(dict):
    """input_rss_info of the reference ({signal type or 'D': {strand: {contig: [tuples]}}}), built from the scan's hit
    list the first time an entry is asked for: a whole-genome run that only writes the gene tables never pays for
    Python lists of every RSS."""

    def __init__(self, scan, input_seq_dict):
        super().__init__()
        self._scan, self._seqs = scan, input_seq_dict

    def __missing__(self, st):
        if st == D:
            v = {sd: combine_D_RSS(self[DL][sd], self[DR][sd], self._seqs, sd, sparse=self._scan.sparse) for sd in (FWD, REV)}
        elif st in SIGNAL_TYPES:
            v = {sd: self._scan.contigwise(st, sd) for sd in (FWD, REV)}
        else:
            raise KeyError(st)
        self[st] = v
        return v

    def __iter__(self):
        return iter(SIGNAL_TYPES + [D])

    def keys(self):
        return SIGNAL_TYPES + [D]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


def _rows_from_device(res, ids, parent_seq, canonical_genes, scan, order):
    """Strings of genes_{V,D,J}.tsv for the integer rows of igd_detect (py/IGDetective.py:364-420); with
    order="reference" the rows of every (strand, contig) are put into the reference's RSS list order."""
    vj, dr = res["vj"], res["d"]
    gene_ids = {0: list(canonical_genes[V].keys()), 1: list(canonical_genes[J].keys())}
    gene_len = {0: [len(s) for s in canonical_genes[V].values()], 1: [len(s) for s in canonical_genes[J].values()]}
    rows = {V: [], J: [], D: []}
    keys = {V: [], J: [], D: []}
    ranks = {}

    def rank(sig_type, strand, ci):
        k = (sig_type, strand, ci)
        if k not in ranks:
            from .order import rank_map
            ranks[k] = rank_map(scan._first_set(sig_type, strand, ci))
        return ranks[k]

    def scan_index(sig_type, strand, ci, first_pos):
        k = 7 if sig_type in (V, DR) else 9
        return first_pos if strand == FWD else scan.lens[ci] - first_pos - k

    # group sizes, to ask for a rank map only where the order is not trivial
    want_order = order == "reference"
    if len(vj):
        # the numeric columns for all rows at once (:391-414); only the strings are made row by row
        hept, start, alen_a = vj["hept"].astype(np.int64), vj["start"].astype(np.int64), vj["alen"].astype(np.int64)
        is_v, fwd, has = vj["gene_type"] == 0, vj["strand"] == 0, vj["frag_len"] > 0
        glen = np.where(is_v, GENE_LENGTH[V], GENE_LENGTH[J])
        best_len = np.where(is_v, np.asarray(gene_len[0])[np.minimum(vj["best_gene"], len(gene_len[0]) - 1)],
                            np.asarray(gene_len[1] or [0])[np.minimum(vj["best_gene"], max(len(gene_len[1]) - 1, 0))])
        end = np.where(has, start + alen_a, best_len)                   # empty slice: ComputeAlignment('', gene) at :388
        lead = is_v == fwd                                              # the gene lies in front of the heptamer
        gs = np.where(lead, np.where(is_v, hept - glen + start, hept - end), hept + 7)
        ge = np.where(lead, hept - 1, np.where(is_v, hept + 7 + glen - start - 1, hept + 7 + end))
        pi = np.rint((vj["matches"].astype(np.int64) - 1) / alen_a * 100).astype(np.int64)     # int(round(pi)): half to even
        first = np.where(is_v, hept, vj["nona"])
        cols = zip(hept.tolist(), vj["nona"].tolist(), vj["frag_start"].tolist(), vj["contig"].tolist(), vj["frag_len"].tolist(),
                   vj["best_gene"].tolist(), start.tolist(), vj["ient"].tolist(), vj["gene_type"].tolist(), vj["strand"].tolist(),
                   vj["direction"].tolist(), gs.tolist(), ge.tolist(), pi.tolist(), alen_a.tolist(), first.tolist())
        text, to = res.get("text"), res.get("text_off")
        k = 0
        for h, n, fstart, ci, flen, best, st, ient, gt, sd, direction, g0, g1, p, al, fi in cols:
            gene, strand, contig = (V, J)[gt], (FWD, REV)[sd], ids[ci]
            if text is not None:                                        # the rows' letters came with the rows (igd_fetch_row_text)
                h7, n9, query = text[to[k]:to[k + 1]], text[to[k + 1]:to[k + 2]], text[to[k + 2]:to[k + 3]]
                k += 3
            else:
                s = parent_seq[contig]
                h7, n9 = _kmer(s, h, 7, strand), _kmer(s, n, 9, strand)
                if flen > 0:
                    # the aligned orientation of the fragment: reverse-complemented once for a '-' strand RSS and once more
                    # when the reverse complement aligned better (:309-317) -- two flips cancel
                    if (sd == 1) != (direction == 45):                  # 45 == ord('-')
                        query = reverse_complement(s[fstart:fstart + flen])[st:ient].upper()
                    else:
                        query = s[fstart + st:fstart + min(ient, flen)].upper()
                else:
                    query = ""
            rows[gene].append([contig, strand, h, n, h7, n9, g0, g1, gene_ids[gt][best], chr(direction), p, float(al), query])
            keys[gene].append((sd, ci, fi))
    text, to = res.get("text"), res.get("text_off")
    k = 3 * len(vj)
    for r in dr.tolist():
        dlh, dln, drh, drn, ci, sd, _ = r
        strand, contig = (FWD, REV)[sd], ids[ci]
        gs, ge = (dlh + 7, drh - 1) if strand == FWD else (drh + 7, dlh - 1)
        if text is not None:
            lh, ln, rh, rn, seq = (text[to[k + i]:to[k + i + 1]] for i in range(5))
            k += 5
        else:
            s = parent_seq[contig]
            lh, ln = _kmer(s, dlh, 7, strand), _kmer(s, dln, 9, strand)
            rh, rn = _kmer(s, drh, 7, strand), _kmer(s, drn, 9, strand)
            seq = s[dlh + 7:drh].upper() if strand == FWD else reverse_complement(s[drh + 7:dlh]).upper()
        rows[D].append([contig, strand, dlh, dln, lh, ln, drh, drn, rh, rn, gs, ge, seq])
        keys[D].append((sd, ci, drh, dln))
    if want_order:
        import collections
        for gene in (V, J):
            n = collections.Counter((k[0], k[1]) for k in keys[gene])
            def key(i, gene=gene):
                sd, ci, first = keys[gene][i]
                if n[(sd, ci)] == 1:
                    return (sd, ci, 0)
                strand = (FWD, REV)[sd]
                return (sd, ci, rank(gene, strand, ci)[scan_index(gene, strand, ci, first)])
            idx = sorted(range(len(rows[gene])), key=key)
            rows[gene] = [rows[gene][i] for i in idx]
        n = collections.Counter((k[0], k[1]) for k in keys[D])
        def dkey(i):
            sd, ci, drh, dln = keys[D][i]
            if n[(sd, ci)] == 1:
                return (sd, ci, 0, 0)
            strand = (FWD, REV)[sd]
            return (sd, ci, rank(DR, strand, ci)[scan_index(DR, strand, ci, drh)],
                    rank(DL, strand, ci)[scan_index(DL, strand, ci, dln)])
        idx = sorted(range(len(rows[D])), key=dkey)
        rows[D] = [rows[D][i] for i in idx]
    return rows


_GENE_TABLES = {}


def _gene_table(genes):
    """(uint8 buffer, int32 offsets) of a canonical-gene dict, built once per dict object (a pipeline passes the same
    two dicts for every contig set it scores)."""
    from .engine import concat
    hit = _GENE_TABLES.get(id(genes))
    if hit is None or hit[0] is not genes or hit[1] != len(genes):
        if len(_GENE_TABLES) > 16:
            _GENE_TABLES.clear()
        hit = (genes, len(genes), concat(list(genes.values())))
        _GENE_TABLES[id(genes)] = hit
    return hit[2]


def detect(engine, input_seq_dict, canonical_genes, order="reference", rss_only=False, load=True, by_seams=False):
    """The module body of py/IGDetective.py:439-487 as a function.
    Returns (input_rss_info, {gene type: rows}) -- rows is None with rss_only.
    One C call (igd_detect) goes from the resident contigs to the integer fields of the rows; Python only builds the
    strings of emitted rows.  by_seams=True runs the same stage function by function through the reference's seams
    (get_contigwise_rss, combine_D_RSS, get_s_fragment_from_RSS, align_fragment_to_genes, ComputeAlignment) -- the
    path the seam-level parity tests exercise, and the one taken for gene tables the packed kernels cannot hold
    (a gene longer than 511 letters)."""
    if not rss_only and not by_seams:
        from ._lib import IGD_ERR_UNSUPPORTED, IgdError
        from .rss import SPACER_LENGTH
        if order not in ("reference", "canonical"):
            raise ValueError("order must be 'reference' or 'canonical'")
        ids = list(input_seq_dict.keys())
        if load:
            if hasattr(input_seq_dict, "_f"):
                engine.load_fasta(input_seq_dict)
            else:
                engine.load_contigs([str(input_seq_dict[c]) for c in ids], ids)
        try:
            res = engine.detect(SPACER_LENGTH, _gene_table(canonical_genes[V]), _gene_table(canonical_genes[J]),
                                gene_length=(GENE_LENGTH[V], GENE_LENGTH[J]), d_gene_length=GENE_LENGTH[D],
                                pi_strict=(PI_CUTOFF["strict"][V], PI_CUTOFF["strict"][J]),
                                pi_relax=(PI_CUTOFF["relax"][V], PI_CUTOFF["relax"][J]),
                                maxk_cutoff=(MAXK_CUTOFF[V], MAXK_CUTOFF[J]))
        except IgdError as e:
            if e.code != IGD_ERR_UNSUPPORTED:
                raise
            return detect(engine, input_seq_dict, canonical_genes, order, rss_only, load=False, by_seams=True)
        scan = RssScan(engine, input_seq_dict, order=order, load=False, hits=res["hits"])
        return RssInfo(scan, input_seq_dict), _rows_from_device(res, ids, input_seq_dict, canonical_genes, scan, order)
    scan = RssScan(engine, input_seq_dict, order=order, load=load)
    info = {st: {sd: scan.contigwise(st, sd) for sd in (FWD, REV)} for st in SIGNAL_TYPES}
    sparse = scan.sparse
    info[D] = {sd: combine_D_RSS(info[DL][sd], info[DR][sd], input_seq_dict, sd, sparse=sparse) for sd in (FWD, REV)}
    if rss_only:
        return info, None
    frags = {g: {sd: get_s_fragment_from_RSS(g, sd, info, input_seq_dict, sparse=sparse) for sd in (FWD, REV)}
             for g in (V, J)}
    rows = {D: d_gene_rows(input_seq_dict, info[D])}
    for gene in (V, J):
        rows[gene] = vj_gene_rows(engine, input_seq_dict, gene, info[gene], frags[gene], canonical_genes[gene])
    return info, rows


def run(input_path, output_path, locus="IGH", rss_only=False, engine=None, datafiles=None,
        gene_dir="combined_reference_genes", order="reference", device=0):
    own = engine is None
    if own:
        engine = Engine(device)
    try:
        engine.set_motifs(dfiles.load_motifs(datafiles))
        print("Finding immunoglobulin genes for locus " + locus + "...")
        # plain FASTA goes through the native reader (igd_fasta_open); .gz through the Python one
        input_seq_dict = fasta_dict(input_path) if str(input_path).endswith(".gz") else NativeFasta(input_path)
        canonical = {g: dfiles.load_canonical_genes(locus, g, gene_dir, datafiles) for g in (V, J)}
        os.makedirs(output_path, exist_ok=True)
        print("Finding candidate RSS...", end=" ")
        info, rows = detect(engine, input_seq_dict, canonical, order, rss_only)
        print("Done")
        if rss_only:
            for st in SIGNAL_TYPES:
                _write_tsv("{}/rss_{}.csv".format(output_path, st), rss_rows(info[st], input_seq_dict))
            return
        print("Aligning candidate genes...", end=" ")
        print("Done")
        for gene in (V, D, J):
            _write_tsv("{}/genes_{}.tsv".format(output_path, gene), [D_HEADER if gene == D else VJ_HEADER] + rows[gene])
        print("Please see {}/ for gene predictions".format(output_path))
    finally:
        if own:
            engine.close()


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    options = "hi:o:m:rg:l:"
    long_options = ["help", "input_file=", "output_directory=", "multi_process=", "rss_only", "genes_type=",
                    "locus=", "datafiles=", "gene_dir=", "device=", "order="]
    try:
        arguments, _ = getopt.getopt(argv, options, long_options)
    except getopt.error as err:
        print(str(err))
        sys.exit(0)
    inp = out = None
    locus, rss_mode, datafiles, gene_dir, device, order = "IGH", False, None, "combined_reference_genes", 0, "reference"
    for a, v in arguments:
        if a in ("-h", "--help"):
            print(__doc__)
            return
        elif a in ("-i", "--input_file"):
            inp = str(v)
        elif a in ("-o", "--output_directory"):
            out = str(v)
        elif a in ("-l", "--locus"):
            if v not in ["IGH", "IGK", "IGL", "TRA", "TRB", "TRG"]:
                print("Incorrect locus argument: " + v)
                sys.exit(1)
            locus = v
        elif a in ("-m", "--multi_process"):
            int(v)                       # accepted for compatibility; the GPU path does not fork
        elif a in ("-r", "--rss_only"):
            rss_mode = True
        elif a == "--datafiles":
            datafiles = v
        elif a == "--gene_dir":
            gene_dir = v
        elif a == "--device":
            device = int(v)
        elif a == "--order":
            order = v
    if inp is None:
        raise NameError("no input file was given")
    if out is None:
        out = ".".join(inp.split(".")[:-1])
    run(inp, out, locus, rss_mode, None, datafiles, gene_dir, order, device)


if __name__ == "__main__":
    main()
