"""CPU oracle for the RSS scan + V/D/J scoring path.  TEST INFRASTRUCTURE ONLY.

A restatement, function by function, of what the reference computes on this path:
  /root/reference/py/IGDetective.py            (scan :138-207, D pairing :210-224,
                                                fragments :260-300, scoring :307-360,
                                                argmax/threshold :463-477, rows :364-422,
                                                TSV writers :227-255, :424-436)
  /root/reference/py/extract_aligned_genes.py  (BioAlign :9-56, SetupAligner :72-83,
                                                ComputeAlignment :85-105, main :129-194)
The aligner itself (BioPython's PairwiseAligner) is restated in oracle/gotoh_ref.c.

Pinning: this module is checked in tests/test_oracle_golden.py against TSVs that the
UNMODIFIED reference scripts produced when run on oracle/bio_shim (see
tests/golden/make_golden.py).  That pins every line of reference logic restated
here; the aligner tie-break underneath remains PARITY UNPINNED (no real BioPython
available) -- see oracle/gotoh_ref.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module.  Nothing under igdetective_b200/ does.
"""
import csv
import ctypes
import os
import pickle

import numpy as np

V, D, J, DR, DL = "V", "D", "J", "D_right", "D_left"
FWD, REV = "+", "-"
SIGNAL_TYPES = [V, J, DL, DR]                       # IGDetective.py:119-125 with GENE_TYPES_TOFIND=[V,D,J]
SPACER_LENGTH = {V: 23, DL: 12, DR: 12, J: 23}      # :28 -- InitializeVariables(:36-51) is a no-op
GENE_LENGTH = {V: 350, J: 70, D: 150}               # :29
PI_CUTOFF = {"strict": {V: 70, J: 70}, "relax": {V: 60, J: 65}}   # :31
MAXK_CUTOFF = {V: 15, J: 11}                        # :32

_HERE = os.path.dirname(os.path.abspath(__file__))
_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa")


def revcomp(s):
    return s.translate(_COMP)[::-1]


# ----------------------------------------------------------------------------- aligner
class _Scoring(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "match", "mismatch",
        "t_int_open", "t_int_ext", "t_left_open", "t_left_ext", "t_right_open", "t_right_ext",
        "q_int_open", "q_int_ext", "q_left_open", "q_left_ext", "q_right_open", "q_right_ext")]


class _Stats(ctypes.Structure):
    _fields_ = [("score", ctypes.c_double), ("ncols", ctypes.c_int), ("start", ctypes.c_int),
                ("end", ctypes.c_int), ("matches", ctypes.c_int), ("lead", ctypes.c_int),
                ("ient", ctypes.c_int)]


STATS_DTYPE = np.dtype([("score", "<f8"), ("ncols", "<i4"), ("start", "<i4"), ("end", "<i4"),
                        ("matches", "<i4"), ("lead", "<i4"), ("ient", "<i4")], align=True)

# SetupAligner, extract_aligned_genes.py:72-83
REFERENCE_SCORING = dict(match=2, mismatch=0,
                         t_int_open=-2, t_int_ext=-1, t_left_open=-2, t_left_ext=-1,
                         t_right_open=-2, t_right_ext=-1,
                         q_int_open=-2, q_int_ext=-1, q_left_open=0, q_left_ext=0,
                         q_right_open=0, q_right_ext=0)

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libigd_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle library missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(path)
        _lib.igo_gotoh_stats_batch.restype = ctypes.c_int
        _lib.igo_gotoh_stats_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(_Scoring),
            ctypes.c_void_p, ctypes.c_int]
    return _lib


def concat_bytes(seqs):
    """list[str] -> (uint8 array, int32 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int32)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return buf, off


def align_stats(targets, queries, pair_t, pair_q, scoring=None, threads=1):
    """First-optimal-alignment statistics for pairs (targets[pair_t[k]], queries[pair_q[k]]).

    target = fragment / window (BioPython's first argument), query = gene.
    Returns a structured array (STATS_DTYPE); `matches` is the TRUE match count.
    """
    sc = _Scoring(**{k: float(v) for k, v in (scoring or REFERENCE_SCORING).items()})
    tb, to = concat_bytes(targets)
    qb, qo = concat_bytes(queries)
    pt = np.ascontiguousarray(pair_t, dtype=np.int32)
    pq = np.ascontiguousarray(pair_q, dtype=np.int32)
    out = np.zeros(len(pt), dtype=STATS_DTYPE)
    assert STATS_DTYPE.itemsize == ctypes.sizeof(_Stats)
    if len(pt):
        rc = lib().igo_gotoh_stats_batch(tb.ctypes.data, to.ctypes.data, qb.ctypes.data, qo.ctypes.data,
                                         pt.ctypes.data, pq.ctypes.data, len(pt), ctypes.byref(sc),
                                         out.ctypes.data, int(threads))
        if rc:
            raise MemoryError("oracle aligner")
    return out


# ----------------------------------------------------------------------------- inputs
def read_fasta(path):
    """SeqIO.parse(path, 'fasta') -> [(id, seq)] in file order; id = first token of the header."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    recs, title, chunks = [], None, []
    with op(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    recs.append((title, "".join(chunks)))
                t = line[1:].rstrip()
                title = t.split(None, 1)[0] if t.strip() else ""
                chunks = []
            elif title is not None:
                chunks.append("".join(line.split()))
    if title is not None:
        recs.append((title, "".join(chunks)))
    return recs


def fasta_dict(path, upper=False):
    """{rec.id: seq} as the reference builds it (IGDetective.py:129,134): later duplicates overwrite."""
    d = {}
    for rid, s in read_fasta(path):
        d[rid] = s.upper() if upper else s
    return d


def load_motifs_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def load_motifs_text(path):
    """tests/golden/datafiles/motif_sets.txt: '<type>\t<k>\t<kmer>' per line."""
    m = {}
    with open(path) as f:
        for line in f:
            t, k, kmer = line.split()
            m.setdefault(t, {}).setdefault(k, set()).add(kmer)
    return m


# ----------------------------------------------------------------------------- scan
def find_valid_motif_idx_py(locus, motifs, k):
    """IGDetective.py:138-144, literally (pure Python; used for small cases and CPU timing)."""
    idx = []
    for i in range(0, len(locus) - int(k) + 1):
        if locus[i:i + int(k)].upper() in motifs:
            idx.append(i)
    return idx


_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def find_valid_motif_idx(locus, motifs, k):
    """Same result as find_valid_motif_idx_py, vectorised: ascending list of window starts whose
    upper-cased k-mer is in `motifs` (motif sets are pure upper-case ACGT)."""
    k = int(k)
    n = len(locus) - k + 1
    if n <= 0:
        return []
    codes = _CODE[np.frombuffer(locus.encode("latin-1"), dtype=np.uint8)]
    bad = (codes > 3).astype(np.int32)
    val = np.zeros(n, dtype=np.int64)
    nbad = np.zeros(n, dtype=np.int32)
    for d in range(k):
        val = val * 4 + (codes[d:d + n] & 3)
        nbad += bad[d:d + n]
    table = np.zeros(4 ** k, dtype=bool)
    for m in motifs:
        if len(m) == k and all(c in "ACGT" for c in m):
            v = 0
            for c in m:
                v = v * 4 + "ACGT".index(c)
            table[v] = True
    return np.nonzero(table[val] & (nbad == 0))[0].tolist()


def reference_set_order(ascending):
    """Iteration order of the index set as find_valid_rss sees it (IGDetective.py:144,197-204):
    set(list) built in a pool worker, pickled to the parent, pickled again to a worker."""
    s = set(ascending)
    s = pickle.loads(pickle.dumps(s))
    s = pickle.loads(pickle.dumps(s))
    return list(s)


def find_valid_rss(hept_order, hept_set, nona_order, nona_set, sig_type, strand, seq_length, spacers=None):
    """IGDetective.py:147-177.  *_order = iteration order of the first set, *_set = membership."""
    spacer = (spacers or SPACER_LENGTH)[sig_type]
    if sig_type in (V, DR):
        k_first, first, second = 7, hept_order, nona_set
    else:
        k_first, first, second = 9, nona_order, hept_set
    out = []
    for idx in first:
        for s in (spacer, spacer - 1, spacer + 1):
            if s + idx + k_first in second:
                out.append((idx, s + idx + k_first))
                break
    if sig_type in (J, DL):
        out = [(b, a) for a, b in out]
    if strand == REV:
        out = [(seq_length - h - 7, seq_length - n - 9) for h, n in out]
    return out


def get_contigwise_rss(sig_type, strand, seqs, motifs, order="reference", scan=find_valid_motif_idx, spacers=None):
    """IGDetective.py:180-207.  seqs: {contig: str}.  order='reference' reproduces the CPython
    set order of the original; 'canonical' iterates the first set ascending."""
    res = {}
    for contig, s in seqs.items():
        sequence = s if strand == FWD else revcomp(s)
        h = scan(sequence, motifs[sig_type]["7"], 7)
        n = scan(sequence, motifs[sig_type]["9"], 9)
        if order == "reference":
            ho, no = reference_set_order(h), reference_set_order(n)
        else:
            ho, no = h, n
        res[contig] = find_valid_rss(ho, set(h), no, set(n), sig_type, strand, len(s), spacers)
    return res


def combine_D_RSS(d_left, d_right, seqs, strand, dgene_len=GENE_LENGTH[D]):
    """IGDetective.py:210-224."""
    res = {c: [] for c in seqs}
    for c in seqs:
        for dr in d_right[c]:
            for dl in d_left[c]:
                if strand == FWD:
                    right_h, left_h = dr[0], dl[0]
                else:
                    right_h, left_h = dl[0], dr[0]
                if right_h - (left_h + 7) <= dgene_len and left_h < right_h:
                    res[c].append((dl[0], dl[1], dr[0], dr[1]))
    return res


def scan_all(seqs, motifs, order="reference", scan=find_valid_motif_idx):
    """input_rss_info of IGDetective.py:443-446."""
    info = {st: {sd: get_contigwise_rss(st, sd, seqs, motifs, order, scan) for sd in (FWD, REV)}
            for st in SIGNAL_TYPES}
    info[D] = {sd: combine_D_RSS(info[DL][sd], info[DR][sd], seqs, sd) for sd in (FWD, REV)}
    return info


# ----------------------------------------------------------------------------- fragments
def s_fragments(gene, strand, rss, seqs):
    """get_s_fragment_from_RSS, IGDetective.py:260-300 (Python slice semantics kept on purpose)."""
    out = {}
    for contig in rss[gene][strand]:
        s = seqs[contig]
        frs = []
        for r in rss[gene][strand][contig]:
            if (gene == V and strand == FWD) or (gene == J and strand == REV):
                frs.append(s[r[0] - GENE_LENGTH[gene]:r[0]])
            elif gene in (V, J):
                frs.append(s[r[0] + 7:r[0] + 7 + GENE_LENGTH[gene]])
            else:
                if strand == FWD:
                    index = r[0] + 7
                    glen = r[2] - index
                else:
                    index = r[2] + 7
                    glen = r[0] - index
                frs.append(s[index:index + glen])
        if strand == REV:
            frs = [revcomp(x) for x in frs]
        out[contig] = frs
    return out


# ----------------------------------------------------------------------------- scoring A
def compute_alignment_batch(frag_list, gene_list, pairs, threads=1):
    """IGDetective.ComputeAlignment (:307-317) for pairs [(fragment string, gene index)].
    Returns per pair: (direction, stats-of-chosen-orientation) where direction is '+' iff
    fwd NumMatches > rev NumMatches."""
    targets, pt, pq = [], [], []
    for frag, g in pairs:
        q = frag.upper()
        targets.append(q)
        targets.append(revcomp(q))
        pt += [len(targets) - 2, len(targets) - 1]
        pq += [g, g]
    st = align_stats(targets, gene_list, pt, pq, threads=threads)
    res = []
    for k in range(len(pairs)):
        f, r = st[2 * k], st[2 * k + 1]
        if int(f["matches"]) - 1 > int(r["matches"]) - 1:
            res.append(("+", f, targets[2 * k]))
        else:
            res.append(("-", r, targets[2 * k + 1]))
    return res


def _pi(st):
    return (int(st["matches"]) - 1) / (int(st["end"]) - int(st["start"])) * 100


def align_fragment_to_genes(fragments, gene_list, threads=1):
    """IGDetective.py:319-360 -> (pi_mat, alig_len_mat), float64 [F, G]."""
    pairs = []
    for f in fragments:
        for g, gs in enumerate(gene_list):
            pairs.append((f if len(f) else "A" * len(gs), g))
    res = compute_alignment_batch(fragments, gene_list, pairs, threads)
    F, G = len(fragments), len(gene_list)
    pi = np.zeros((F, G))
    ln = np.zeros((F, G))
    for k, (_, st, _) in enumerate(res):
        pi[k // G][k % G] = _pi(st)
        ln[k // G][k % G] = int(st["end"]) - int(st["start"])
    return pi, ln


def igdetective(seqs, motifs, canonical, order="reference", threads=1, scan=find_valid_motif_idx):
    """The module body of IGDetective.py:439-487 -> {'V': rows, 'D': rows, 'J': rows} (no header)
    plus the rss info.  canonical = {'V': {id: SEQ}, 'J': {id: SEQ}} (upper-cased, :134)."""
    rss = scan_all(seqs, motifs, order, scan)
    frags = {g: {sd: s_fragments(g, sd, rss, seqs) for sd in (FWD, REV)} for g in (V, D, J)}
    rows = {}
    # D rows, :366-378
    drows = []
    for sd in (FWD, REV):
        for contig in rss[D][sd]:
            s = seqs[contig]
            for e in rss[D][sd][contig]:
                if sd == FWD:
                    lh, ln_, rh, rn = (s[e[0]:e[0] + 7].upper(), s[e[1]:e[1] + 9].upper(),
                                       s[e[2]:e[2] + 7].upper(), s[e[3]:e[3] + 9].upper())
                    gs, ge, gene = e[0] + 7, e[2] - 1, s[e[0] + 7:e[2]].upper()
                else:
                    lh, ln_, rh, rn = (revcomp(s[e[0]:e[0] + 7]).upper(), revcomp(s[e[1]:e[1] + 9]).upper(),
                                       revcomp(s[e[2]:e[2] + 7]).upper(), revcomp(s[e[3]:e[3] + 9]).upper())
                    gs, ge, gene = e[2] + 7, e[0] - 1, revcomp(s[e[2] + 7:e[0]]).upper()
                drows.append([contig, sd, e[0], e[1], lh, ln_, e[2], e[3], rh, rn, gs, ge, gene])
    rows[D] = drows
    for gene in (V, J):
        ids = list(canonical[gene].keys())
        glist = list(canonical[gene].values())
        flat = []
        for sd in (FWD, REV):
            for contig in frags[gene][sd]:
                flat.extend(frags[gene][sd][contig])
        pi_mat, len_mat = align_fragment_to_genes(flat, glist, threads)
        out = []
        k = 0
        winners = []
        for sd in (FWD, REV):
            for contig in frags[gene][sd]:
                for i, f in enumerate(frags[gene][sd][contig]):
                    b = int(np.argmax(pi_mat[k]))
                    pi, mk = pi_mat[k][b], len_mat[k][b]
                    k += 1
                    if pi >= PI_CUTOFF["strict"][gene] or (pi >= PI_CUTOFF["relax"][gene] and mk >= MAXK_CUTOFF[gene]):
                        winners.append((sd, contig, i, f, b, pi, mk))
        # :388 -- the winner is re-aligned with the raw (possibly empty) fragment
        res = compute_alignment_batch(None, glist, [(w[3], w[4]) for w in winners], threads)
        for (sd, contig, i, f, b, pi, mk), (direction, st, tgt) in zip(winners, res):
            r = rss[gene][sd][contig][i]
            s = seqs[contig]
            start, end = int(st["start"]), int(st["end"])
            predicted = tgt[int(st["lead"]):int(st["ient"])]
            if sd == FWD:
                if gene == V:
                    ge = r[0] - 1
                    gs = r[0] - GENE_LENGTH[gene] + start
                else:
                    gs = r[0] + 7
                    ge = gs + end
                h, n = s[r[0]:r[0] + 7].upper(), s[r[1]:r[1] + 9].upper()
            else:
                if gene == V:
                    gs = r[0] + 7
                    ge = gs + GENE_LENGTH[gene] - start - 1
                else:
                    ge = r[0] - 1
                    gs = r[0] - end
                h, n = revcomp(s[r[0]:r[0] + 7]).upper(), revcomp(s[r[1]:r[1] + 9]).upper()
            out.append([contig, sd, r[0], r[1], h, n, gs, ge, ids[b], direction, int(round(pi)), mk, predicted])
        rows[gene] = out
    return rows, rss


D_HEADER = ['reference contig', 'strand', 'left heptamer index', 'left nonamer index',
            'left heptamer', 'left nonamer', 'right heptamer index', 'right nonamer index',
            'right heptamer', 'right nonamer', 'start of gene', 'end of gene', 'gene sequence']
VJ_HEADER = ['reference contig', 'strand', 'heptamer index', 'nonamer index', 'heptamer', 'nonamer',
             'start of gene', 'end of gene', 'best aligned human gene', 'alignment direction',
             'alignment PI', 'longest common k-mer', 'gene sequence']


def write_genes_tsv(path, gene, rows):
    """print_predicted_genes, IGDetective.py:424-436 (csv excel dialect => CRLF)."""
    with open(path, "w", newline="") as f:
        csv.writer(f, delimiter="\t").writerows([D_HEADER if gene == D else VJ_HEADER] + rows)


def rss_rows(rss_for_type, seqs):
    """write_rss_to_file, IGDetective.py:227-255 (min/max heptamer arguments unused by the caller)."""
    rows = [['7-mer index', '9-mer index', '7-mer', '9-mer', 'reference contig', 'strand']]
    for sd in (FWD, REV):
        for contig in seqs:
            s = seqs[contig]
            for h, n in rss_for_type[sd][contig]:
                if sd == FWD:
                    hs, ns = s[h:h + 7].upper(), s[n:n + 9].upper()
                else:
                    hs, ns = revcomp(s[h:h + 7]).upper(), revcomp(s[n:n + 9]).upper()
                rows.append([h, n, hs, ns, contig, sd])
    return rows


def write_rss_csv(path, rss_for_type, seqs):
    with open(path, "w", newline="") as f:
        csv.writer(f, delimiter="\t").writerows(rss_rows(rss_for_type, seqs))


# ----------------------------------------------------------------------------- scoring B
_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_CODON = {a + b + c: _AAS[16 * i + 4 * j + k]
          for i, a in enumerate(_BASES) for j, b in enumerate(_BASES) for k, c in enumerate(_BASES)}
_AMBIG = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG",
          "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "ACGT", "N": "ACGT"}


def translate(s):
    """Seq(s).translate(), extract_aligned_genes.py:169: standard table, trailing partial codon dropped,
    ambiguous codons -> the shared amino acid if unique else 'X'."""
    s = s.upper().replace("U", "T")
    out = []
    for i in range(0, len(s) - len(s) % 3, 3):
        c = s[i:i + 3]
        aa = _CODON.get(c)
        if aa is None:
            opts = [_AMBIG[x] for x in c]
            aas = {_CODON[a + b + d] for a in opts[0] for b in opts[1] for d in opts[2]}
            aa = aas.pop() if len(aas) == 1 else "X"
        out.append(aa)
    return "".join(out)


def window_best_alignment(window, genes, threads=1):
    """extract_aligned_genes.ComputeAlignment (:85-105) for one window: queries [window, RC(window)]
    in RAW case against upper-cased genes [(id, SEQ)], scan order orientation-major, keep the LAST
    maximum of PI (>=, starting from 0).  Returns (gene_seq, gene_id, pi, strand) or None."""
    targets = [window, revcomp(window)]
    G = len(genes)
    st = align_stats(targets, [g[1] for g in genes], [0] * G + [1] * G, list(range(G)) * 2, threads=threads)
    best_pi, best = 0, None
    for k in range(2 * G):
        pi = _pi(st[k])
        if pi >= best_pi:
            best_pi, best = pi, k
    if best is None:
        return None
    s = st[best]
    if int(s["end"]) - int(s["start"]) == 0:
        return None
    seq = targets[best // G][int(s["lead"]):int(s["ient"])].upper()
    if seq == "" or seq[0] == " ":
        return None
    return seq, genes[best % G][0], best_pi, "+-"[best // G]


def extract_aligned_genes(contigs, position_dict, genes, threads=1):
    """extract_aligned_genes.main :141-184 after the SAM has been read.
    contigs {id: raw str}; position_dict {id: [pos,...]} (ProcessSamFile, :112-127);
    genes [(id, SEQ upper)].  Returns the genes.tsv columns as a dict of lists."""
    gene_len = 400
    df = {k: [] for k in ("Contig", "Pos", "Seq", "AASeq", "PI", "BestHit", "Productive", "Strand")}
    for cid in position_dict:
        cs = contigs[cid]
        prev = -1
        for pos in sorted(position_dict[cid]):
            if pos - prev <= gene_len:
                continue
            frag = cs[max(0, pos - gene_len):min(len(cs), pos + gene_len)]
            r = window_best_alignment(frag, genes, threads)
            if r is None:
                continue
            seq, gid, pi, strand = r
            aa = translate(seq)
            df["Contig"].append(cid); df["Pos"].append(pos); df["Seq"].append(seq); df["AASeq"].append(aa)
            df["PI"].append(pi); df["BestHit"].append(gid); df["Productive"].append(aa.find("*") == -1)
            df["Strand"].append(strand)
            prev = pos
    return df
