"""Pins the oracle: its restatement of the reference's logic must reproduce, byte for byte, the
files the UNMODIFIED reference scripts wrote (tests/golden/, see make_golden.py), plus the
known answers SURVEY.md §8(c) lists for the human IGH locus.  CPU only."""
import csv
import io
import os

import pandas as pd
import pytest

from conftest import DATAFILES, GOLDEN, canonical, golden_input
from oracle import igd_oracle as O

THREADS = len(os.sched_getaffinity(0))


def _tsv_bytes(rows):
    buf = io.StringIO(newline="")
    csv.writer(buf, delimiter="\t").writerows(rows)
    return buf.getvalue().encode()


@pytest.fixture(scope="module")
def cow(motifs):
    seqs = O.fasta_dict(golden_input("cow"))
    return seqs, O.igdetective(seqs, motifs, canonical("combined_reference_genes"), threads=THREADS)


def test_oracle_genes_tsv_cow(cow):
    seqs, (rows, rss) = cow
    gdir = os.path.join(GOLDEN, "igdetective", "cow__combined_reference_genes")
    for g in ("V", "D", "J"):
        want = open(os.path.join(gdir, "genes_%s.tsv" % g), "rb").read()
        assert _tsv_bytes([O.D_HEADER if g == "D" else O.VJ_HEADER] + rows[g]) == want, g


@pytest.mark.parametrize("locus", ["IGK", "TRB"])
def test_oracle_genes_tsv_other_loci(motifs, locus):
    seqs = O.fasta_dict(golden_input("cow"))
    rows, _ = O.igdetective(seqs, motifs, canonical("combined_reference_genes", locus), threads=THREADS)
    gdir = os.path.join(GOLDEN, "igdetective", "cow__combined_reference_genes__" + locus)
    for g in ("V", "D", "J"):
        want = open(os.path.join(gdir, "genes_%s.tsv" % g), "rb").read()
        assert _tsv_bytes([O.D_HEADER if g == "D" else O.VJ_HEADER] + rows[g]) == want, (locus, g)


def test_oracle_multi_contig_corner_cases(motifs):
    seqs = O.fasta_dict(golden_input("multi"))
    assert list(seqs) == ["ctgA", "ctgB", "ctgC", "short_v", "early_v", "rev_units", "empty_record", "short_j"]
    rows, rss = O.igdetective(seqs, motifs, canonical("combined_reference_genes"), threads=THREADS)
    gdir = os.path.join(GOLDEN, "igdetective", "multi__combined_reference_genes")
    for g in ("V", "D", "J"):
        want = open(os.path.join(gdir, "genes_%s.tsv" % g), "rb").read()
        assert _tsv_bytes([O.D_HEADER if g == "D" else O.VJ_HEADER] + rows[g]) == want, g
    for st in O.SIGNAL_TYPES:
        want = open(os.path.join(gdir, "rss_%s.csv" % st), "rb").read()
        assert _tsv_bytes(O.rss_rows(rss[st], seqs)) == want, st


def test_oracle_rss_csv_cow(cow):
    seqs, (rows, rss) = cow
    gdir = os.path.join(GOLDEN, "igdetective", "cow__combined_reference_genes")
    for st in O.SIGNAL_TYPES:
        want = open(os.path.join(gdir, "rss_%s.csv" % st), "rb").read()
        assert _tsv_bytes(O.rss_rows(rss[st], seqs)) == want, st


@pytest.mark.parametrize("name", ["human", "mouse"])
def test_oracle_rss_csv(motifs, name):
    seqs = O.fasta_dict(golden_input(name))
    rss = O.scan_all(seqs, motifs)
    gdir = os.path.join(GOLDEN, "igdetective", "%s__combined_reference_genes" % name)
    for st in O.SIGNAL_TYPES:
        want = open(os.path.join(gdir, "rss_%s.csv" % st), "rb").read()
        assert _tsv_bytes(O.rss_rows(rss[st], seqs)) == want, st


def test_oracle_human_j_known_answers(motifs):
    """IGHJ1,2,3,4,6 on '+' at these heptamer indices with PI = (len-1)/len."""
    seqs = O.fasta_dict(golden_input("human"))
    (cid, s), = seqs.items()
    rss = {"J": {sd: O.get_contigwise_rss("J", sd, seqs, motifs) for sd in "+-"}}
    frags = {sd: O.s_fragments("J", sd, rss, seqs) for sd in "+-"}
    genes = canonical("combined_reference_genes")["J"]
    ids, glist = list(genes), list(genes.values())
    flat = frags["+"][cid] + frags["-"][cid]
    pi, ln = O.align_fragment_to_genes(flat, glist, threads=THREADS)
    found = {}
    for k, r in enumerate(rss["J"]["+"][cid]):
        b = int(pi[k].argmax())
        if pi[k][b] >= 70:
            found[r[0]] = (ids[b], round(float(pi[k][b]), 2), int(ln[k][b]))
    assert found == {1018253: ("IGHJ1*01", 98.08, 52), 1018460: ("IGHJ2*01", 98.11, 53),
                     1019075: ("IGHJ3*02", 98.00, 50), 1019449: ("IGHJ4*02", 97.92, 48),
                     1020451: ("IGHJ6*03", 98.41, 63)}


def test_oracle_invariants(motifs, cow):
    """SURVEY.md §8(c)(3)."""
    seqs, (rows, rss) = cow
    for r in rows["V"]:
        assert r[4] == "CACAGTG" and r[5] in motifs["V"]["9"]
    for r in rows["J"]:
        assert r[4] in motifs["J"]["7"] and r[5] in motifs["J"]["9"]
    for r in rows["D"]:
        assert len(r[12]) <= 150 and r[4] in motifs["D_left"]["7"] and r[8] in motifs["D_right"]["7"]


@pytest.mark.parametrize("scenario,example", [("human_IGHV", "human"), ("cow_IGHV", "cow"), ("multi_IGHJ", "multi")])
def test_oracle_scoring_b(scenario, example):
    gdir = os.path.join(GOLDEN, "scoring_b", scenario)
    contigs = O.fasta_dict(golden_input(example))
    pos = {}
    for line in open(os.path.join(gdir, "alignment.sam")):
        if line[0] == "@":
            continue
        f = line.split()
        if f[2] != "*" and int(f[3]) not in pos.setdefault(f[2], []):
            pos[f[2]].append(int(f[3]))
    genes = [(i, s.upper()) for i, s in O.read_fasta(os.path.join(DATAFILES, "combined_reference_genes",
                                                                   scenario.split("_")[1] + ".fa.gz"))]
    df = pd.DataFrame(O.extract_aligned_genes(contigs, pos, genes, threads=THREADS))
    want = open(os.path.join(gdir, "genes.tsv")).read()
    assert df.to_csv(index=False, sep="\t") == want
