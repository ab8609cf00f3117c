"""Minimal stand-in for the parts of BioPython the reference imports.

TEST INFRASTRUCTURE ONLY.  BioPython (>= 1.82, README.md:8 of the reference) is
not installed in this environment and cannot be fetched, so the unmodified
reference scripts (/root/reference/py/IGDetective.py,
/root/reference/py/extract_aligned_genes.py) are run against this shim to
produce the golden fixtures under tests/golden/.  Only behaviour those two
scripts use is provided: Bio.Seq.Seq, Bio.SeqIO.parse(path, "fasta"),
Bio.Align.PairwiseAligner.  The aligner core is oracle/gotoh_ref.c.
PARITY UNPINNED: see the header of oracle/gotoh_ref.c.
"""
__version__ = "0.0-igd-oracle-shim"
