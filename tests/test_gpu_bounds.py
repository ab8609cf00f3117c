"""The two upper bounds igd_detect prunes its V scoring with (csrc/lcs.cu, through igd_match_bounds): the kernels against
the plain recurrences, and the recurrences' claim -- they bound what the oracle's BioPython restatement returns."""
import numpy as np
import pytest

import synth
from conftest import canonical
from kernel_model import lcs_bound, y_bound
from oracle import igd_oracle as O

pytestmark = pytest.mark.gpu


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))


def rc(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def check_kernels(engine, frags, genes):
    lcs, y = engine.match_bounds(frags, genes)
    for u, f in enumerate(frags):
        for o, row in enumerate((f.upper(), rc(f.upper()))):
            for g, q in enumerate(genes):
                assert lcs[u, o, g] == lcs_bound(row, q), ("lcs", u, o, g, len(f), len(q))
                assert y[u, o, g] == y_bound(row, q), ("y", u, o, g, len(f), len(q))
    return lcs, y


def test_bound_kernels_equal_the_recurrences_on_awkward_shapes(engine):
    rng = np.random.default_rng(11)
    frags = [rand_seq(rng, n) for n in (1, 2, 21, 22, 23, 44, 350, 351, 352, 353, 400, 511, 512)]
    frags += [rand_seq(rng, 350, "ACGTN"), "A" * 300, rand_seq(rng, 349, "acgt")]
    genes = [rand_seq(rng, n) for n in (1, 2, 31, 32, 33, 64, 65, 290, 352, 353, 511)]
    genes += [rand_seq(rng, 296, "ACGTNR"), rand_seq(rng, 300, "acgtACGT"), "T" * 280]
    check_kernels(engine, frags, genes)


def test_bounds_hold_for_the_alignments_the_oracle_returns(engine):
    """m equal columns and c = len(alignment) - len(gene) inserted letters: m <= LCS and 2 m - c <= Y, for true genes,
    mutated ones, unrelated fragments, in both orientations; and the bounds do their job: every unrelated pair is ruled
    out by Y at the 60 % cut-off, none by the LCS"""
    rng = np.random.default_rng(12)
    genes = list(canonical("combined_reference_genes")["V"].values())[::7]
    frags, related = [], []
    for i in range(18):
        g = genes[rng.integers(len(genes))]
        if i % 3 == 0:
            frags.append(synth.background(rng, 350).tobytes().decode()); related.append(False)
        else:
            body = synth.mutate(rng, g.upper(), 0.02 * i, 0.01 * (i % 4))
            f = (rand_seq(rng, 400) + body)[-350:]
            frags.append(f if i % 2 else rc(f)); related.append(True)
    lcs, y = engine.match_bounds(frags, genes)
    rows = []
    for f in frags:
        rows += [f.upper(), rc(f.upper())]
    pt, pq = np.meshgrid(np.arange(len(rows)), np.arange(len(genes)), indexing="ij")
    st = O.align_stats(rows, genes, pt.ravel(), pq.ravel(), threads=8)
    m = st["matches"].astype(np.int64).reshape(len(frags), 2, len(genes))
    alen = (st["end"].astype(np.int64) - st["start"]).reshape(len(frags), 2, len(genes))
    glen = np.array([len(g) for g in genes])
    assert (m <= lcs).all()
    assert (2 * m - (alen - glen) <= y).all()
    pi = (m - 1) / alen * 100
    by_y = 50.0 * y >= 60.0 * glen + 100.0                   # what igd_detect tests
    assert (by_y | (pi < 60.0)).all()
    unrelated = ~np.array(related)
    assert not by_y[unrelated].any() and ((lcs[unrelated] - 1) / glen * 100 >= 60.0).mean() > 0.9
    assert by_y[np.array(related)].any()
    # spot-check the kernels against the recurrences on these inputs too
    for u in (0, 1, 5):
        for g in (0, 3):
            assert lcs[u, 1, g] == lcs_bound(rows[2 * u + 1], genes[g]) and y[u, 0, g] == y_bound(rows[2 * u], genes[g])
