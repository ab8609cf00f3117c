"""RSS stage behind the reference's function seams (py/IGDetective.py:138-300).

One GPU pass (Engine.scan_rss) evaluates both strands and the four signal types; the functions
below serve its result in the reference's own shapes:
  get_contigwise_rss(sig_type, strand, parent_seq) -> {contig: [(hept_idx, nona_idx), ...]}   (:180)
  combine_D_RSS(D_left_idx, D_right_idx, input_seq_dict, strand, Dgene_len)                   (:210)
  get_s_fragment_from_RSS(gene, strand, ...)                                                  (:268)
"""
import numpy as np

from .engine import SIGNAL_INDEX, SIGNAL_NAMES
from .fasta import reverse_complement
from .order import rank_map

V, D, J, DR, DL = "V", "D", "J", "D_right", "D_left"
FWD, REV = "+", "-"
SIGNAL_TYPES = [V, J, DL, DR]                        # :119-125 with the default GENE_TYPES_TOFIND
# InitializeVariables (:36-51) assigns locals only, so every locus runs with these
SPACER_LENGTH = {V: 23, DL: 12, DR: 12, J: 23}
GENE_LENGTH = {V: 350, J: 70, D: 150}
HEPT_FIRST = {V: True, DR: True, J: False, DL: False}
# Whole-genome runs (order="canonical": hundreds of thousands of contigs, most without any RSS of a given type)
# share ONE immutable empty value between the contigs that have nothing instead of building a list per contig and
# call; the reference-order mode keeps the reference's own shape, a fresh list per contig.
_NOTHING = ()


def _empty_map(keys, sparse):
    return dict.fromkeys(keys, _NOTHING) if sparse else {c: [] for c in keys}


class RssScan:
    """Result of one scan of `parent_seq` ({contig id: str}) on an Engine."""

    def __init__(self, engine, parent_seq, spacers=None, order="reference", load=True, hits=None):
        if order not in ("reference", "canonical"):
            raise ValueError("order must be 'reference' or 'canonical'")
        self.engine = engine
        self.seqs = parent_seq
        self.ids = list(parent_seq.keys())
        self.lens = [len(parent_seq[c]) for c in self.ids]
        self.spacers = dict(spacers or SPACER_LENGTH)
        self.order = order
        self.sparse = order == "canonical"
        if load:                       # load=False: the caller already put these contigs on the engine
            if hasattr(parent_seq, "_f"):                  # fasta.NativeFasta: no Python copy of the letters
                engine.load_fasta(parent_seq)
            else:
                engine.load_contigs([str(parent_seq[c]) for c in self.ids], self.ids)
        # hits: the hit list of a scan already run on these contigs (Engine.detect), not to scan twice
        self.hits = engine.scan_rss(self.spacers) if hits is None else hits
        self._kmers = {}
        self._cache = {}

    def _kmer_hits(self, k):
        """(hits of find_valid_motif_idx for this k, sorted by (contig, pos); first hit of every contig)"""
        if k not in self._kmers:
            kh = self.engine.scan_kmers(k)
            bounds = np.searchsorted(kh["contig"], np.arange(len(self.ids) + 1))       # one split per k, not a mask per list
            self._kmers[k] = (kh, bounds)
        return self._kmers[k]

    def _first_set(self, sig_type, strand, ci):
        """Ascending scanned-strand indices of the 5' k-mer set of find_valid_rss (:153-160)."""
        k = 7 if HEPT_FIRST[sig_type] else 9
        kh, bounds = self._kmer_hits(k)
        bit = 1 << (SIGNAL_INDEX[sig_type] + (4 if strand == REV else 0))
        mine = kh[bounds[ci]:bounds[ci + 1]]
        sel = mine[(mine["flags"] & bit) != 0]["pos"].astype(np.int64)
        if strand == REV:
            sel = (self.lens[ci] - sel - k)[::-1]
        return sel.tolist()

    def contigwise(self, sig_type, strand):
        key = (sig_type, strand)
        if key in self._cache:
            return self._cache[key]
        t, s = SIGNAL_INDEX[sig_type], (0 if strand == FWD else 1)
        h = self.hits[(self.hits["sig_type"] == t) & (self.hits["strand"] == s)]
        # hits arrive sorted by (contig, type, strand, scan-order index): one split per contig that has any
        res = _empty_map(self.ids, self.sparse)
        if len(h):
            cs, first = np.unique(h["contig"], return_index=True)
            bounds = np.append(first, len(h)).tolist()
            hept, nona = h["hept"].tolist(), h["nona"].tolist()
            for ci, lo, hi in zip(cs.tolist(), bounds[:-1], bounds[1:]):
                tuples = list(zip(hept[lo:hi], nona[lo:hi]))
                if self.order == "reference" and len(tuples) > 1:
                    k = 7 if HEPT_FIRST[sig_type] else 9
                    L = self.lens[ci]
                    rank = rank_map(self._first_set(sig_type, strand, ci))
                    first_el = 0 if HEPT_FIRST[sig_type] else 1
                    if strand == FWD:
                        tuples.sort(key=lambda tp: rank[tp[first_el]])
                    else:
                        tuples.sort(key=lambda tp: rank[L - tp[first_el] - k])
                res[self.ids[ci]] = tuples
        self._cache[key] = res
        return res


def get_contigwise_rss(sig_type, strand, parent_seq, scan=None, engine=None, order="reference"):
    """Drop-in for py/IGDetective.py:180.  Pass a shared RssScan to serve all eight
    (sig_type, strand) calls of :443 from one GPU pass."""
    if scan is None:
        if engine is None:
            raise ValueError("get_contigwise_rss needs an Engine or an RssScan")
        scan = RssScan(engine, parent_seq, order=order)
    return scan.contigwise(sig_type, strand)


def combine_D_RSS(D_left_idx, D_right_idx, input_seq_dict, strand, Dgene_len=GENE_LENGTH[D], sparse=False):
    """py/IGDetective.py:210-224: every (dr, dl) with left < right and right-(left+7) <= Dgene_len,
    dr-outer / dl-inner in the given list orders.  The reference's double loop is replaced by one
    sorted search per contig; the output (content and order) is the same."""
    res = _empty_map(input_seq_dict.keys(), sparse)
    for c, drs in D_right_idx.items():
        if not drs:
            continue
        dls = D_left_idx[c]
        if not dls:
            continue
        dl0 = np.fromiter((x[0] for x in dls), np.int64, len(dls))
        dr0 = np.fromiter((x[0] for x in drs), np.int64, len(drs))
        order = np.argsort(dl0, kind="stable")
        sdl = dl0[order]
        if strand == FWD:        # left = dl, right = dr:  dr-157 <= dl <= dr-1
            lo = np.searchsorted(sdl, dr0 - (Dgene_len + 7), "left")
            hi = np.searchsorted(sdl, dr0 - 1, "right")
        else:                    # left = dr, right = dl:  dr+1 <= dl <= dr+157
            lo = np.searchsorted(sdl, dr0 + 1, "left")
            hi = np.searchsorted(sdl, dr0 + (Dgene_len + 7), "right")
        out = []
        for k in np.nonzero(hi > lo)[0]:
            dr = drs[k]
            for i in np.sort(order[lo[k]:hi[k]]):      # dl-inner order = position in the dl list
                dl = dls[i]
                out.append((dl[0], dl[1], dr[0], dr[1]))
        if out or not sparse:
            res[c] = out
    return res


def extract_s_fragment(index, extract_dir, length, parent_seq):
    """py/IGDetective.py:260-266 -- plain Python slicing, including its negative-start behaviour."""
    if extract_dir == FWD:
        return parent_seq[index:index + length]
    return parent_seq[index - length:index]


def get_s_fragment_from_RSS(gene, strand, input_rss_info, input_seq_dict, sparse=False):
    """py/IGDetective.py:268-300 (the globals it reads are parameters here)."""
    out = _empty_map(input_rss_info[gene][strand], sparse)
    for c, rss in input_rss_info[gene][strand].items():
        if not rss:
            continue
        s = input_seq_dict[c]
        frs = []
        for r in rss:
            if (gene == V and strand == FWD) or (gene == J and strand == REV):
                frs.append(extract_s_fragment(r[0], REV, GENE_LENGTH[gene], s))
            elif gene in (V, J):
                frs.append(extract_s_fragment(r[0] + 7, FWD, GENE_LENGTH[gene], s))
            else:
                index = r[0] + 7 if strand == FWD else r[2] + 7
                glen = (r[2] if strand == FWD else r[0]) - index
                frs.append(extract_s_fragment(index, FWD, glen, s))
        if strand == REV:
            frs = [reverse_complement(x) for x in frs]
        out[c] = frs
    return out
