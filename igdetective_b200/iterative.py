"""Gene-set glue of the iterative stage, behind the reference driver's own function:

    CombineIGGenes(genes_fasta, igdetective_tsv, output_fasta)      run_iterative_igdetective.py:130-151

The reference keeps the sequences that are not a substring of another sequence with a double loop of
str.find over all pairs -- O(n^2 L), minutes at 10^4 sequences and hopeless at the 10^5 of a whole-genome
candidate set ("next" row 3 of the scope table).  Here the same set is built in the same way (so that its
iteration order, which decides the seq_<i> names and through them every later tie, is the reference's own
under the same PYTHONHASHSEED) and the substring test runs against an index: all sequences joined into one
text; a sequence is contained in another one exactly when it occurs in the text somewhere else than at its
own place.  Small sets search the text directly (one C-level find per sequence); large ones look their
16-letter prefix up in a sorted array of rolling hashes of every text position and verify the few candidates.
"""
import os

import numpy as np

from .fasta import read_fasta

_K = 16
_SEP = "\x00"


def _tsv_gene_sequences(path):
    """df['gene sequence'] of pd.read_csv(path, sep='\\t') (:136-138): the last column of genes_V.tsv."""
    import csv
    with open(path, newline="") as f:
        rows = list(csv.reader(f, delimiter="\t"))
    if not rows:
        return []
    col = rows[0].index("gene sequence")
    return [r[col] for r in rows[1:] if len(r) > col]


def contained_in_another(seqs):
    """[True iff seqs[i] is a substring of some seqs[j] != seqs[i]] for a list of DISTINCT strings."""
    n = len(seqs)
    out = [False] * n
    if n < 2:
        return out
    starts = np.zeros(n + 1, np.int64)
    starts[1:] = np.cumsum([len(s) + 1 for s in seqs])
    text = _SEP.join(seqs) + _SEP
    if n <= 2000 or min(len(s) for s in seqs) < _K:
        small = True
    else:
        small = False
    if small:
        for i, s in enumerate(seqs):
            if s == "":
                out[i] = True                      # ''.find succeeds in every other string (:143)
                continue
            p = text.find(s)
            if p == starts[i]:
                p = text.find(s, p + 1)
            out[i] = p != -1
        return out
    # large sets: rolling hash of every K-letter window of the text, sorted once
    b = np.frombuffer(text.encode("latin-1"), np.uint8).astype(np.uint64)
    m = len(b) - _K + 1
    code = np.zeros(m, np.uint64)
    with np.errstate(over="ignore"):
        for j in range(_K):
            code = code * np.uint64(1099511628211) + b[j:j + m]
    order = np.argsort(code, kind="stable")
    sorted_code = code[order]
    own = code[starts[:-1]]                         # every sequence has at least K letters here
    lo = np.searchsorted(sorted_code, own, "left")
    hi = np.searchsorted(sorted_code, own, "right")
    tb = text
    for i, s in enumerate(seqs):
        for p in order[lo[i]:hi[i]].tolist():
            if p != starts[i] and tb.startswith(s, p):
                out[i] = True
                break
    return out


def CombineIGGenes(genes_fasta, igdetective_tsv, output_fasta):
    """run_iterative_igdetective.py:130-151 with the same signature and the same output file."""
    nucl_seqs = set()
    if os.path.exists(genes_fasta):
        for _, s in read_fasta(genes_fasta):
            nucl_seqs.add(s)
    if os.path.exists(igdetective_tsv):
        for s in _tsv_gene_sequences(igdetective_tsv):
            nucl_seqs.add(s)
    seqs = list(nucl_seqs)                          # the reference's iteration order (:140)
    drop = contained_in_another(seqs)
    with open(output_fasta, "w") as fh:
        k = 0
        for s, d in zip(seqs, drop):
            if not d:
                fh.write(">seq_" + str(k) + "\n" + s + "\n")
                k += 1
