"""FASTA reading with the semantics of SeqIO.parse(path, 'fasta') as the reference uses it
(py/IGDetective.py:129,134; py/extract_aligned_genes.py:143-151): record id = first
whitespace-delimited token of the header, sequence lines concatenated, case preserved."""
import gzip

_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa")


def reverse_complement(s):
    """str(Seq(s).reverse_complement()): IUPAC-aware, case preserving."""
    return s.translate(_COMP)[::-1]


def read_fasta(path):
    """-> [(id, sequence)] in file order."""
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
    """{rec.id: seq} built like the reference's dict comprehensions: a repeated id keeps its first
    position and takes the last sequence."""
    d = {}
    for rid, s in read_fasta(path):
        d[rid] = s.upper() if upper else s
    return d
