"""Locating and loading the reference's data files (datafiles/motifs, datafiles/*_reference_genes).

Two layouts are understood: the reference checkout's own (`motifs` pickle + `<LOCUS><GENE>.fa`) and
the re-encoded copy under tests/golden/datafiles (`motif_sets.txt` + `.fa.gz`) that travels to
machines without the reference checkout."""
import os
import pickle

from .fasta import fasta_dict, read_fasta

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ["datafiles", "/root/reference/datafiles",
               os.path.join(_HERE, "..", "tests", "golden", "datafiles")]


def find_datafiles(explicit=None):
    for d in ([explicit] if explicit else []) + [os.environ.get("IGD_DATAFILES")] + _CANDIDATES:
        if d and (os.path.exists(os.path.join(d, "motifs")) or os.path.exists(os.path.join(d, "motif_sets.txt"))):
            return os.path.abspath(d)
    raise FileNotFoundError("could not find the datafiles directory (motifs / reference genes); "
                            "pass datafiles=... or set IGD_DATAFILES")


def load_motifs(datafiles=None):
    """VALID_MOTIFS, py/IGDetective.py:54-58: {sig_type: {'7': set, '9': set}}."""
    d = find_datafiles(datafiles)
    p = os.path.join(d, "motifs")
    if os.path.exists(p):
        with open(p, "rb") as f:
            return pickle.load(f)
    m = {}
    with open(os.path.join(d, "motif_sets.txt")) as f:
        for line in f:
            t, k, kmer = line.split()
            m.setdefault(t, {}).setdefault(k, set()).add(kmer)
    return m


def gene_fasta_path(locus_gene, gene_dir="combined_reference_genes", datafiles=None):
    d = find_datafiles(datafiles)
    for ext in (".fa", ".fa.gz"):
        p = os.path.join(d, gene_dir, locus_gene + ext)
        if os.path.exists(p):
            return p
    raise FileNotFoundError("%s/%s.fa[.gz] not found under %s" % (gene_dir, locus_gene, d))


def load_canonical_genes(locus, gene, gene_dir="combined_reference_genes", datafiles=None):
    """canonical_genes[gene], py/IGDetective.py:131-134: {rec.id: upper-cased sequence}."""
    return fasta_dict(gene_fasta_path(locus + gene, gene_dir, datafiles), upper=True)


def load_gene_records(path):
    """genes list of extract_aligned_genes.main (:148-151): [(id, upper-cased sequence)]."""
    return [(rid, s.upper()) for rid, s in read_fasta(path)]
