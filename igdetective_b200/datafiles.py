"""Locating and loading the reference's data files (datafiles/motifs, datafiles/*_reference_genes).

Where: the explicit argument (`--datafiles`), else $IGD_DATAFILES, else `datafiles/` under the working
directory -- the reference's own convention (py/IGDetective.py:54-58,133 open 'datafiles/...' relative
to the directory it is started from).  Nothing else is probed: the product does not depend on the test
tree or on a development checkout.

Two layouts are understood: the reference checkout's (`motifs` pickle + `<LOCUS><GENE>.fa`) and the
re-encoded text form (`motif_sets.txt`, one `type k kmer` per line, + `.fa.gz`).  The pickle is read
by an unpickler that resolves no globals at all: the shipped file holds only dict / set / str."""
import os
import pickle

from .fasta import fasta_dict, read_fasta


def find_datafiles(explicit=None):
    for d in (explicit, os.environ.get("IGD_DATAFILES"), "datafiles"):
        if d and (os.path.exists(os.path.join(d, "motifs")) or os.path.exists(os.path.join(d, "motif_sets.txt"))):
            return os.path.abspath(d)
    raise FileNotFoundError("could not find the datafiles directory (motifs / reference genes): "
                            "pass --datafiles / datafiles=..., set IGD_DATAFILES, or start from the directory "
                            "that holds datafiles/ (as py/IGDetective.py expects)")


class _PlainUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        raise pickle.UnpicklingError("datafiles/motifs must hold only dict / set / str (found %s.%s)" % (module, name))


def load_motifs(datafiles=None):
    """VALID_MOTIFS, py/IGDetective.py:54-58: {sig_type: {'7': set, '9': set}}."""
    d = find_datafiles(datafiles)
    p = os.path.join(d, "motif_sets.txt")
    if os.path.exists(p):
        m = {}
        with open(p) as f:
            for line in f:
                t, k, kmer = line.split()
                m.setdefault(t, {}).setdefault(k, set()).add(kmer)
        return m
    with open(os.path.join(d, "motifs"), "rb") as f:
        return _PlainUnpickler(f).load()


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
