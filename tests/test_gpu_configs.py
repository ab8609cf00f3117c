"""BASELINE.json configs[2] and configs[4] at sizes the CPU oracle finishes in seconds, rows against the ORACLE
(oracle/igd_oracle.py: the reference's logic restated, pinned to the unmodified reference's files), not only against
the library itself: the same generator bench.py times at full size, the same piece / shard plumbing
(shard.plan_pieces, multi.detect_pieces, shard.merge_rows), and a fragmented assembly of many short contigs whose
RSSs sit at contig edges (negative slice starts, empty fragments, py/IGDetective.py:265,339)."""
import numpy as np
import pytest

from conftest import canonical
from oracle import igd_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_rows(seqs, motifs, canon):
    rows, _ = O.igdetective(seqs, motifs, canon, order="canonical", threads=8)
    return {g: [list(r) for r in rows[g]] for g in ("V", "D", "J")}


def _norm(rows):
    return {g: [[(int(x) if isinstance(x, (np.integer,)) else (float(x) if isinstance(x, np.floating) else x)) for x in r]
                for r in rows[g]] for g in ("V", "D", "J")}


def test_config3_shape_pieces_and_shards_equal_the_oracle(engine, motifs):
    """the strong-scaling leg of bench.py at 1/500 size: 500 contigs, loci planted on 5 scaffolds; pieces of 64 kb
    (scaffolds are cut), three shards, rows merged -> equal to the oracle's rows on the whole assembly"""
    import bench
    from igdetective_b200 import multi, shard
    from igdetective_b200.fasta import ArraySeq
    canon = canonical("combined_reference_genes")
    asm = bench.Assembly3(motifs, canon, scale=0.002)
    # plant_locus windows are 2 Mb at full size; here a scaffold is ~100-500 kb: the patches still fit (w = min(L, 2e6))
    bufs = [np.empty(L, np.uint8) for L in asm.lengths]
    for ci, L in enumerate(asm.lengths):
        asm.fill(ci, 0, L, bufs[ci])
    seqs = {cid: b.tobytes().decode() for cid, b in zip(asm.ids, bufs)}
    want = _oracle_rows(seqs, motifs, canon)
    assert len(want["V"]) >= 100 and len(want["D"]) >= 5 and len(want["J"]) >= 3
    for world, tile in ((1, 32_000_000), (3, 64_000)):
        parts = shard.plan_pieces(asm.lengths, world, tile=tile)
        per_rank = []
        for pieces in parts:
            part = {k: ArraySeq(bufs[ci][p0:p1]) for k, (ci, _, _, p0, p1) in enumerate(pieces)}
            off = np.zeros(len(pieces) + 1, np.uint64)
            off[1:] = np.cumsum([p[4] - p[3] for p in pieces], dtype=np.uint64)
            buf = np.concatenate([bufs[ci][p0:p1] for ci, _, _, p0, p1 in pieces]) if pieces else np.zeros(1, np.uint8)
            engine.load_contigs((buf, off), list(part))
            per_rank.append(multi.detect_pieces(engine, None, asm.ids, pieces, canon, "canonical", part=part))
        got = _norm(shard.merge_rows(per_rank, asm.ids))
        for g in ("V", "D", "J"):
            assert got[g] == _norm(want)[g], (world, g)


def test_config5_shape_fragmented_assembly_equals_the_oracle(engine, motifs):
    """2000 contigs, truncated-Pareto lengths from 1 kb (scaled down from configs[4]), an RSS at a contig edge in
    every 20th contig, V units 100-300 nt into contigs shorter and longer than the 350-nt flank"""
    import synth
    from igdetective_b200 import igdetective
    canon = canonical("combined_reference_genes")
    rng = np.random.default_rng(5)
    n = 2000
    lengths = np.minimum((1000 * (1 - rng.random(n)) ** (-1 / 1.1)).astype(np.int64), 60_000)
    contigs = [synth.background(rng, int(L)) for L in lengths]
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in synth.SIG}
    vs = list(canon["V"].values())
    for k, c in enumerate(contigs):
        if k % 20 == 0:
            u = np.frombuffer(synth.rss_unit(rng, motifs, synth.SIG[rng.integers(4)], sets).encode(), np.uint8)
            p = 0 if rng.integers(2) else len(c) - len(u)
            c[p:p + len(u)] = u
        if k % 50 == 7 and len(c) > 700:
            g = synth.mutate(rng, vs[rng.integers(len(vs))], 0.05, 0.005)
            u = np.frombuffer((g + synth.rss_unit(rng, motifs, "V", sets)).encode(), np.uint8)
            if rng.integers(2):
                u = synth.revcomp_bytes(u)
            p = int(rng.integers(0, len(c) - len(u)))
            c[p:p + len(u)] = u
    seqs = {"ctg%06d" % i: c.tobytes().decode() for i, c in enumerate(contigs)}
    _, got = igdetective.detect(engine, seqs, canon, order="canonical")
    want = _oracle_rows(seqs, motifs, canon)
    assert len(want["V"]) >= 10
    for g in ("V", "D", "J"):
        assert _norm(got)[g] == _norm(want)[g], g


def test_config4_shape_iterative_candidates_vs_all_loci_genes(engine):
    """configs[3] scaled down: 350-nt V flanks, 70-nt J flanks and 800-nt raw-case windows against the V / J genes of
    ALL loci of combined_reference_genes (1172 / 235 genes, mixed-case files upper-cased as the stage scripts do):
    the float64 PI / length matrices of scoring A and the winners of scoring B against the oracle"""
    import os

    import synth
    from conftest import DATAFILES
    from igdetective_b200.scoring import align_fragment_to_genes, window_alignments
    rng = np.random.default_rng(4)
    gv, gj = [], []
    d = os.path.join(DATAFILES, "combined_reference_genes")
    for fn in sorted(os.listdir(d)):
        for rid, s in O.fasta_dict(os.path.join(d, fn), upper=True).items():
            if 0 < len(s) <= 511:
                (gv if fn[3] == "V" else gj).append((fn[:4] + "|" + rid, s))
    assert len(gv) > 1000 and len(gj) > 200

    def bg(k):
        return synth.background(rng, k).tobytes().decode()
    vf = [(bg(350) + synth.mutate(rng, gv[rng.integers(len(gv))][1], 0.1, 0.01))[-350:] if i % 3 else bg(350) for i in range(24)]
    jf = [(synth.mutate(rng, gj[rng.integers(len(gj))][1], 0.1, 0.01) + bg(70))[:70] if i % 3 else bg(70) for i in range(40)]
    for frags, genes in ((vf, gv), (jf, gj)):
        glist = [g[1] for g in genes]
        pi, alen = align_fragment_to_genes(frags, glist, engine=engine)
        wpi, wlen = O.align_fragment_to_genes(frags, glist, threads=8)
        assert np.array_equal(pi, wpi) and np.array_equal(alen, wlen)
    wins = []
    for i in range(12):
        g = gv[rng.integers(len(gv))][1]
        w = (bg(300) + synth.mutate(rng, g, 0.1, 0.01) + bg(500))[:800]
        if i % 2:
            w = w[:200] + w[200:500].lower() + w[500:]
        wins.append(O.revcomp(w) if i % 3 == 0 else w)
    got = window_alignments(engine, wins, gv)
    for w, (a, strand) in zip(wins, got):
        want = O.window_best_alignment(w, gv, threads=8)
        assert want is not None and (a.gene_seq, a.gene_id, a.pi, strand) == want
