"""Scoring stage behind the reference's function seams.

  ComputeAlignment(seq_A, seq_B, ext)                      py/IGDetective.py:307-317
  align_fragment_to_genes(fragments, canon_genes, ...)     py/IGDetective.py:319-360
  BioAlign                                                 py/extract_aligned_genes.py:9-56
  ComputeAlignment(aligner, query_list, strand_list, genes) py/extract_aligned_genes.py:85-105

The GPU returns, per (target, query) pair, the integers BioAlign derives from BioPython's first
optimal alignment; this module turns them into the reference's return values (float64 PI etc.)."""
import numpy as np

from .fasta import reverse_complement


class BioAlignView:
    """The observable surface of BioAlign (extract_aligned_genes.py:9-56) over GPU results."""

    def __init__(self, target, matches, start, end, ient=None):
        self._target = target
        self._matches = int(matches)
        self.gene_range = (int(start), int(end))
        self._ient = None if ient is None else int(ient)

    def NumMatches(self):
        return self._matches - 1            # the reference starts counting at -1 (:16,33-38)

    def __len__(self):
        return self.gene_range[1] - self.gene_range[0]

    def PI(self):
        return self.NumMatches() / len(self) * 100

    def AlignmentRange(self):
        return self.gene_range

    def QuerySeq(self):
        if self._ient is None:
            raise RuntimeError("alignment was computed without ient")
        return self._target[self.gene_range[0]:self._ient].upper()


def pi_from(matches, alen):
    """float64 PI exactly as BioAlign.PI(): (matches-1) / len * 100, evaluated left to right."""
    return (np.asarray(matches, np.int64) - 1) / np.asarray(alen, np.int64) * 100


def compute_alignments(engine, pairs, genes, want_ient=False):
    """IGDetective.ComputeAlignment for many (fragment, gene index) pairs in one GPU batch.
    Returns a list of (BioAlignView, direction, matches-1): direction '+' iff the forward
    orientation has strictly more matches (:314-317)."""
    targets, pt, pq = [], [], []
    index = {}
    for frag, g in pairs:
        q = str(frag).upper()
        if q not in index:
            index[q] = len(targets)
            targets.append(q)
            targets.append(reverse_complement(q))
        pt += [index[q], index[q] + 1]
        pq += [g, g]
    r = engine.align(targets, genes, pt, pq, detail=True)
    out = []
    for k in range(len(pairs)):
        f, b = 2 * k, 2 * k + 1
        pick, d = (f, "+") if r["matches"][f] - 1 > r["matches"][b] - 1 else (b, "-")
        view = BioAlignView(targets[pt[pick]], r["matches"][pick], r["start"][pick], r["end"][pick], r["ient"][pick])
        out.append((view, d, int(r["matches"][pick]) - 1))
    return out


def ComputeAlignment(seq_A, seq_B, extend_alignment=None, engine=None):
    """Drop-in for py/IGDetective.py:307 (one pair)."""
    return compute_alignments(engine, [(seq_A, 0)], [str(seq_B)], want_ient=True)[0]


def align_fragment_to_genes(fragments, canon_genes, scoring_scheme=None, gene=None, engine=None):
    """Drop-in for py/IGDetective.py:319-360 -> (pi_mat, alig_len_mat), float64 [F, G].
    `scoring_scheme` is ignored, as in the reference (set_aligner, :303-304)."""
    genes = [str(g) for g in canon_genes]
    F, G = len(fragments), len(genes)
    if F == 0 or G == 0:
        return np.zeros((F, G)), np.zeros((F, G))
    # upper-casing, the reverse-complement orientation, the choice between the two (:314-317), the
    # 'A' * len(gene) stand-in of an empty fragment (:339-340) and PI happen behind the C ABI
    return engine.score_fragments([str(f) for f in fragments], genes)


class Alignment:
    """extract_aligned_genes.Alignment (:58-70)."""

    def __init__(self):
        self.gene_seq, self.gene_id, self.pi = "", "", 0

    def Initiate(self, gene_seq, gene_id, pi):
        self.gene_seq, self.gene_id, self.pi = gene_seq, gene_id, pi

    def Empty(self):
        return self.gene_seq == "" or self.gene_seq[0] == " "


def window_alignments(engine, windows, genes):
    """extract_aligned_genes.ComputeAlignment (:85-105) for many windows in one GPU batch.
    windows: raw-case strings; genes: [(id, SEQ)].  For each window the scan order is
    [window, RC(window)] x genes, keeping the LAST maximum of PI (>=, from 0).
    Returns [(Alignment, strand)]."""
    G = len(genes)
    if len(windows) == 0 or G == 0:
        return [(Alignment(), "") for _ in windows]
    windows = [str(w) for w in windows]
    r = engine.score_windows(windows, [g[1] for g in genes])
    out = []
    for w, b in enumerate(r["best"].tolist()):
        a = Alignment()
        s, e, ie = int(r["start"][w]), int(r["end"][w]), int(r["ient"][w])
        if b < 0 or e - s == 0:
            out.append((a, ""))
            continue
        query = windows[w] if b < G else reverse_complement(windows[w])
        a.Initiate(query[s:ie].upper(), genes[b % G][0], float(r["pi"][w]))
        out.append((a, "+-"[b // G]))
    return out
