"""Alignment parity: CUDA Gotoh (through the C ABI) vs the oracle's BioPython restatement.  Bit-exact."""
import numpy as np
import pytest

import synth
from conftest import canonical
from oracle import igd_oracle as O

pytestmark = pytest.mark.gpu
KEYS = ("score", "matches", "start", "end", "ient")


def compare(engine, targets, queries, pt, pq):
    """both kernels: detail=True -> two-register kernel (all fields), detail=False -> packed kernel
    wherever the pair qualifies (score, matches, alignment length)"""
    want = O.align_stats(targets, queries, pt, pq, threads=8)
    w = {k: want[k].astype(np.int64) for k in KEYS}
    w["alen"] = w["end"] - w["start"]
    for detail, keys in ((True, KEYS + ("alen",)), (False, ("score", "matches", "alen"))):
        got = engine.align(targets, queries, pt, pq, detail=detail)
        for k in keys:
            bad = np.nonzero(got[k] != w[k])[0]
            assert bad.size == 0, (detail, k, bad[:5], got[k][bad[:5]], w[k][bad[:5]],
                                   [(targets[pt[i]], queries[pq[i]]) for i in bad[:1]])


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))


def test_align_small_exhaustive_lengths(engine):
    rng = np.random.default_rng(1)
    targets = [rand_seq(rng, n) for n in list(range(1, 40)) + [71, 72, 73, 191, 192, 193]]
    queries = [rand_seq(rng, n) for n in list(range(1, 30)) + [50, 63]]
    pt, pq = np.meshgrid(np.arange(len(targets)), np.arange(len(queries)), indexing="ij")
    compare(engine, targets, queries, pt.ravel(), pq.ravel())


def test_align_bucket_edges(engine):
    rng = np.random.default_rng(2)
    lens = [351, 352, 353, 511, 512, 513, 799, 800, 801, 1022, 1023]
    targets = [rand_seq(rng, n) for n in lens]
    queries = [rand_seq(rng, n) for n in (1, 38, 294, 325, 700, 1023)]
    # related pairs too: a query embedded in the target
    for n in lens:
        q = queries[2]
        t = rand_seq(rng, max(0, n - len(q)) // 2) + synth.mutate(rng, q, 0.1, 0.02)
        targets.append((t + rand_seq(rng, n))[:n])
    pt, pq = np.meshgrid(np.arange(len(targets)), np.arange(len(queries)), indexing="ij")
    compare(engine, targets, queries, pt.ravel(), pq.ravel())


def test_align_v_fragments_vs_gene_set(engine):
    """350-nt fragments (true, mutated, background, 'A'*len) x the IGHV set, both orientations"""
    rng = np.random.default_rng(3)
    genes = list(canonical("combined_reference_genes")["V"].values())
    frags = []
    for i in range(24):
        g = genes[rng.integers(len(genes))]
        if i % 3 == 0:
            f = rand_seq(rng, 350)
        else:
            m = synth.mutate(rng, g, rng.uniform(0.02, 0.25), rng.uniform(0, 0.02))
            f = (rand_seq(rng, 350) + m)[-350:]
        frags.append(f)
        frags.append(O.revcomp(f))
    frags.append("A" * 296)
    frags.append("ACGT" * 80 + "NNNNNNNNNN" + "Y" * 5)
    gsel = list(rng.choice(len(genes), 120, replace=False))
    queries = [genes[i] for i in gsel]
    pt, pq = np.meshgrid(np.arange(len(frags)), np.arange(len(queries)), indexing="ij")
    compare(engine, frags, queries, pt.ravel(), pq.ravel())


def test_align_j_fragments(engine):
    rng = np.random.default_rng(4)
    genes = list(canonical("combined_reference_genes")["J"].values())
    frags = []
    for i in range(64):
        g = genes[rng.integers(len(genes))]
        f = (synth.mutate(rng, g, 0.1, 0.02) + rand_seq(rng, 70))[:70] if i % 2 else rand_seq(rng, int(rng.integers(1, 71)))
        frags.append(f)
    pt, pq = np.meshgrid(np.arange(len(frags)), np.arange(len(genes)), indexing="ij")
    compare(engine, frags, genes, pt.ravel(), pq.ravel())


def test_align_windows_raw_case(engine):
    """800-nt raw-case windows (scoring B): lower-case letters mismatch in the DP but count as matches"""
    rng = np.random.default_rng(5)
    genes = list(canonical("combined_reference_genes")["V"].values())
    wins = []
    for i in range(10):
        g = genes[rng.integers(len(genes))]
        w = rand_seq(rng, 250) + synth.mutate(rng, g, 0.08, 0.01) + rand_seq(rng, 400)
        w = w[:800]
        if i % 2:
            a = int(rng.integers(0, 600))
            w = w[:a] + w[a:a + 150].lower() + w[a + 150:]
        if i == 9:
            w = w.lower()                       # fully soft-masked: every path ties, pure tie-break test
        wins.append(w)
        wins.append(O.revcomp(w))
    gsel = list(rng.choice(len(genes), 60, replace=False))
    queries = [genes[i] for i in gsel]
    pt, pq = np.meshgrid(np.arange(len(wins)), np.arange(len(queries)), indexing="ij")
    compare(engine, wins, queries, pt.ravel(), pq.ravel())


def test_align_low_complexity_ties(engine):
    """homopolymers / repeats maximise co-optimal paths: the M > Ix > Iy preference decides everything"""
    targets = ["A" * 50, "AC" * 40, "A" * 20 + "C" * 20, "ACGT" * 30, "T" * 7, "G"]
    queries = ["A" * 10, "AC" * 7, "CA" * 7, "AAAC", "C", "ACGT" * 5, "T" * 30]
    pt, pq = np.meshgrid(np.arange(len(targets)), np.arange(len(queries)), indexing="ij")
    compare(engine, targets, queries, pt.ravel(), pq.ravel())


def test_align_empty_target_and_errors(engine):
    from igdetective_b200._lib import IgdError
    r = engine.align(["", "ACGT"], ["ACGTAC"], [0, 1], [0, 0], detail=True)
    assert (r["score"][0], r["matches"][0], r["start"][0], r["end"][0], r["ient"][0]) == (-7, 0, 0, 6, 0)
    with pytest.raises(IgdError):
        engine.align(["A" * 1024], ["ACGT"], [0], [0])
    with pytest.raises(IgdError):
        engine.align(["ACGT"], [""], [0], [0])
    with pytest.raises(IgdError):
        engine.align(["ACGT"], ["ACGT"], [0], [0], scoring=dict(match=2, mismatch=0, target_internal_open=-2,
                     target_internal_extend=-1, target_end_open=0, target_end_extend=0, query_internal_open=-2,
                     query_internal_extend=-1, query_end_open=0, query_end_extend=0))


def test_align_other_schemes(engine):
    rng = np.random.default_rng(6)
    targets = [rand_seq(rng, int(rng.integers(1, 120))) for _ in range(40)]
    queries = [rand_seq(rng, int(rng.integers(1, 60))) for _ in range(20)]
    pt, pq = [a.ravel() for a in np.meshgrid(np.arange(40), np.arange(20), indexing="ij")]
    for sc in (dict(match=1, mismatch=-1, to=-3, te=-1, qo=-3, qe=-1, qeo=-3, qee=-1),
               dict(match=5, mismatch=-4, to=-10, te=-1, qo=-8, qe=-2, qeo=0, qee=0),
               dict(match=1, mismatch=0, to=0, te=0, qo=0, qe=0, qeo=0, qee=0)):
        gsc = dict(match=sc["match"], mismatch=sc["mismatch"], target_internal_open=sc["to"],
                   target_internal_extend=sc["te"], target_end_open=sc["to"], target_end_extend=sc["te"],
                   query_internal_open=sc["qo"], query_internal_extend=sc["qe"],
                   query_end_open=sc["qeo"], query_end_extend=sc["qee"])
        osc = dict(match=sc["match"], mismatch=sc["mismatch"], t_int_open=sc["to"], t_int_ext=sc["te"],
                   t_left_open=sc["to"], t_left_ext=sc["te"], t_right_open=sc["to"], t_right_ext=sc["te"],
                   q_int_open=sc["qo"], q_int_ext=sc["qe"], q_left_open=sc["qeo"], q_left_ext=sc["qee"],
                   q_right_open=sc["qeo"], q_right_ext=sc["qee"])
        want = O.align_stats(targets, queries, pt, pq, scoring=osc, threads=8)
        got = engine.align(targets, queries, pt, pq, scoring=gsc, detail=True)
        for k in KEYS:
            assert np.array_equal(got[k], want[k].astype(np.int64)), (sc, k)
        fast = engine.align(targets, queries, pt, pq, scoring=gsc)
        assert np.array_equal(fast["score"], want["score"].astype(np.int64)), sc
        assert np.array_equal(fast["matches"], want["matches"]), sc
        assert np.array_equal(fast["alen"], want["end"] - want["start"]), sc


def test_align_dense_equals_oracle(engine):
    """igd_align_dense: fast path (all pairs fit the packed word), with an empty target, and the fallback
    to pair tables (a query longer than 511)"""
    rng = np.random.default_rng(8)
    genes = list(canonical("combined_reference_genes")["V"].values())[:70]
    targets = [rand_seq(rng, n) for n in (350, 350, 70, 1, 200, 352, 353, 511)] + [""]
    targets += [(rand_seq(rng, 300) + synth.mutate(rng, genes[3], 0.1, 0.02) + rand_seq(rng, 300))[:800].lower()[:400] + rand_seq(rng, 400)]
    for queries in (genes, genes[:5] + [rand_seq(rng, 600)]):
        T, Q = len(targets), len(queries)
        pt, pq = [a.ravel() for a in np.meshgrid(np.arange(T), np.arange(Q), indexing="ij")]
        want = O.align_stats(targets, queries, pt, pq, threads=8)
        got = engine.align_dense(targets, queries)
        assert np.array_equal(got["score"].ravel(), want["score"].astype(np.int64))
        assert np.array_equal(got["matches"].ravel(), want["matches"])
        assert np.array_equal(got["alen"].ravel(), want["end"] - want["start"])


def test_score_fragments_equals_reference_loops(engine):
    """igd_score_fragments = align_fragment_to_genes (py/IGDetective.py:319-360): both orientations,
    the forward one only on strictly more matches, float64 PI; duplicates, lower case, IUPAC letters,
    an empty fragment ('A' * len(gene)) and the pair-table fallback (a gene longer than 511)"""
    rng = np.random.default_rng(21)
    genes = list(canonical("combined_reference_genes")["V"].values())[:40]
    frags = [rand_seq(rng, 350) for _ in range(3)]
    frags += [(rand_seq(rng, 100) + synth.mutate(rng, genes[i], 0.1, 0.02))[-350:] for i in (1, 5, 9)]
    frags += [O.revcomp((rand_seq(rng, 80) + synth.mutate(rng, genes[7], 0.05, 0.01))[-350:])]
    frags += [frags[0], frags[4].lower(), "", rand_seq(rng, 60, "ACGTNRY"), "acgtn" * 30, rand_seq(rng, 1), ""]
    for gl in (genes, genes[:6] + [rand_seq(rng, 600)]):
        want_pi, want_len = O.align_fragment_to_genes(frags, gl, threads=8)
        pi, ln, dirs = engine.score_fragments(frags, gl, direction=True)
        assert pi.dtype == np.float64 and pi.shape == (len(frags), len(gl))
        assert np.array_equal(pi, want_pi)            # bit-exact float64
        assert np.array_equal(ln, want_len)
        pairs = [(f if len(f) else "A" * len(g), k) for f in frags for k, g in enumerate(gl)]
        want_dir = [d for d, _, _ in O.compute_alignment_batch(frags, gl, pairs, threads=8)]
        assert [chr(c) for c in dirs.ravel()] == want_dir
    assert engine.score_fragments([], genes)[0].shape == (0, len(genes))
    assert engine.score_fragments(["", ""], genes[:3])[0].shape == (2, 3)


def test_score_windows_equals_reference_loop(engine):
    """igd_score_windows = extract_aligned_genes.ComputeAlignment (py/extract_aligned_genes.py:85-105):
    raw-case windows, both orientations, LAST maximum of PI among PI >= 0; soft-masked stretches, a
    window with no match at all, duplicated genes (ties resolved towards the later one), short
    windows at a contig end, and the pair-table fallback (a gene longer than 511)"""
    from igdetective_b200.scoring import window_alignments
    rng = np.random.default_rng(33)
    gl = list(canonical("combined_reference_genes")["V"].items())[:30]
    genes = [(k, v.upper()) for k, v in gl]
    genes += [("dup_of_3", genes[3][1]), ("dup_of_7", genes[7][1])]          # equal PI: the later gene must win
    wins = []
    for i in range(10):
        g = genes[int(rng.integers(len(genes)))][1]
        w = (rand_seq(rng, 300) + synth.mutate(rng, g, 0.1, 0.01) + rand_seq(rng, 500))[:800]
        if i % 2:
            w = w[:250] + w[250:600].lower() + w[600:]
        if i % 3 == 0:
            w = O.revcomp(w)
        wins.append(w)
    wins += [rand_seq(rng, 800), "N" * 50, rand_seq(rng, 37), genes[3][1], rand_seq(rng, 1)]
    for gs in (genes, genes[:4] + [("long", rand_seq(rng, 600))]):
        got = window_alignments(engine, wins, gs)
        for w, (a, strand) in zip(wins, got):
            want = O.window_best_alignment(w, gs, threads=4)
            if want is None:
                assert a.Empty() and strand == "", (w[:30], a.gene_seq[:30])
            else:
                assert (a.gene_seq, a.gene_id, a.pi, strand) == want, (w[:30], a.gene_id, want[1])
    assert window_alignments(engine, [], genes) == []


def test_align_lower_case_on_either_side(engine):
    """lower-case letters in the queries, in the targets, in both (same letter: same case is a match, different
    case only counts), with IUPAC letters in between -- every case class of the packed kernel"""
    rng = np.random.default_rng(44)
    base = [rand_seq(rng, n) for n in (350, 120, 33, 800)]

    def mask(s, p):
        m = rng.random(len(s)) < p
        return "".join(c.lower() if k else c for c, k in zip(s, m))
    rel = [synth.mutate(rng, base[0], 0.1, 0.02), synth.mutate(rng, base[1], 0.05, 0.0), base[2], base[0][50:300]]
    for tp, qp in ((0.0, 0.4), (0.4, 0.0), (0.5, 0.5), (1.0, 1.0)):
        targets = [mask(s, tp) for s in base] + ["acgtNNacgtRYACGT" * 5]
        queries = [mask(s, qp) for s in rel] + ["ACGTnnACGTryacgt" * 3]
        pt, pq = np.meshgrid(np.arange(len(targets)), np.arange(len(queries)), indexing="ij")
        compare(engine, targets, queries, pt.ravel(), pq.ravel())
        want = O.align_stats(targets, queries, pt.ravel(), pq.ravel(), threads=8)
        got = engine.align_dense(targets, queries)
        assert np.array_equal(got["matches"].ravel(), want["matches"])
        assert np.array_equal(got["alen"].ravel(), want["end"] - want["start"])


def test_compute_alignment_seam_has_the_reference_signature(engine):
    """extract_aligned_genes.ComputeAlignment(aligner, query_list, strand_list, gene_seqs) -> (Alignment, strand),
    py/extract_aligned_genes.py:85-105, called the way main() calls it (:153-154,165-166): an aligner object set
    up by SetupAligner, [window, RC(window)], ['+', '-'], gene records with .seq / .id"""
    import collections

    from igdetective_b200 import extract_aligned_genes as X
    Rec = collections.namedtuple("Rec", "id seq")
    rng = np.random.default_rng(66)
    gl = list(canonical("combined_reference_genes")["V"].items())[40:60]
    genes = [(k, v.upper()) for k, v in gl] + [("dup", gl[5][1].upper())]
    recs = [Rec(k, v) for k, v in genes]
    aligner = X.GpuAligner(engine)
    X.SetupAligner(aligner)
    assert X.scoring_of(aligner) == dict(match=2, mismatch=0, target_internal_open=-2, target_internal_extend=-1,
                                         target_end_open=-2, target_end_extend=-1, query_internal_open=-2,
                                         query_internal_extend=-1, query_end_open=0, query_end_extend=0)
    wins = []
    for i in range(6):
        g = genes[int(rng.integers(len(genes)))][1]
        w = (rand_seq(rng, 200) + synth.mutate(rng, g, 0.08, 0.01) + rand_seq(rng, 400))[:800]
        if i % 2:
            w = w[:100] + w[100:500].lower() + w[500:]
        wins.append(O.revcomp(w) if i % 3 == 0 else w)
    wins += [rand_seq(rng, 30), "N" * 12]
    for w in wins:
        got, strand = X.ComputeAlignment(aligner, [w, O.revcomp(w)], ['+', '-'], recs)
        want = O.window_best_alignment(w, genes, threads=4)
        if want is None:
            assert got.Empty() and strand == ''
        else:
            assert (got.gene_seq, got.gene_id, got.pi, strand) == want
    # another scheme through the same seam: plain attribute assignment like a PairwiseAligner
    aligner.mismatch_score = -1
    aligner.query_internal_open_gap_score = -3
    got, strand = X.ComputeAlignment(aligner, [wins[1]], ['+'], recs[:5])
    sc = dict(O.REFERENCE_SCORING, mismatch=-1, q_int_open=-3)
    st = O.align_stats([wins[1]], [g[1] for g in genes[:5]], [0] * 5, list(range(5)), scoring=sc)
    best_pi, best = 0, None
    for k in range(5):
        pi = (int(st[k]["matches"]) - 1) / (int(st[k]["end"]) - int(st[k]["start"])) * 100
        if pi >= best_pi:
            best_pi, best = pi, k
    assert best is not None and strand == '+'
    assert (got.gene_id, got.pi) == (genes[best][0], best_pi)
    assert got.gene_seq == wins[1][int(st[best]["lead"]):int(st[best]["ient"])].upper()
    assert X.ComputeAlignment(aligner, [], [], recs)[1] == ''
