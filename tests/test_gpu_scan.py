"""RSS scan parity: CUDA path (through the C ABI) vs the oracle.  Integer/index work: bit-exact."""
import numpy as np
import pytest

import synth
from conftest import golden_input
from oracle import igd_oracle as O

pytestmark = pytest.mark.gpu
SIG = synth.SIG


def oracle_hits(seqs, motifs):
    """{(contig index, type, strand): sorted list of (hept, nona)}"""
    out = {}
    ids = list(seqs.keys())
    for t in SIG:
        for sd in "+-":
            r = O.get_contigwise_rss(t, sd, seqs, motifs, order="canonical")
            for ci, c in enumerate(ids):
                out[(ci, t, sd)] = sorted(r[c])
    return out


def gpu_hits(engine, seqs):
    engine.load_contigs(seqs)
    h = engine.scan_rss(O.SPACER_LENGTH)
    out = {(ci, t, sd): [] for ci in range(len(seqs)) for t in SIG for sd in "+-"}
    for r in h:
        out[(int(r["contig"]), SIG[r["sig_type"]], "+-"[r["strand"]])].append((int(r["hept"]), int(r["nona"])))
    return {k: sorted(v) for k, v in out.items()}, h


def check(engine, seqs, motifs):
    want = oracle_hits(seqs, motifs)
    got, raw = gpu_hits(engine, seqs)
    assert want.keys() == got.keys()
    for k in want:
        assert got[k] == want[k], k
    return raw


@pytest.mark.parametrize("name", ["cow", "human", "mouse"])
def test_scan_examples(engine, motifs, name):
    seqs = O.fasta_dict(golden_input(name))
    raw = check(engine, seqs, motifs)
    assert len(raw) > 50


def test_kmer_hits_examples(engine, motifs):
    seqs = O.fasta_dict(golden_input("cow"))
    (cid, s), = seqs.items()
    engine.load_contigs(seqs)
    for k in (7, 9):
        kh = engine.scan_kmers(k)
        assert np.all(np.diff(kh["pos"]) > 0)
        for ti, t in enumerate(SIG):
            fw = O.find_valid_motif_idx(s, motifs[t][str(k)], k)
            assert kh["pos"][(kh["flags"] >> ti) & 1 == 1].tolist() == fw
            rv = O.find_valid_motif_idx(O.revcomp(s), motifs[t][str(k)], k)
            got = sorted((len(s) - kh["pos"][(kh["flags"] >> (4 + ti)) & 1 == 1] - k).tolist())
            assert got == rv


def _strs(contigs):
    return {"c%d" % i: c.tobytes().decode() for i, c in enumerate(contigs)}


def test_scan_synthetic_ragged(engine, motifs):
    """many contigs of awkward lengths: empty, shorter than a k-mer, around chunk / tile boundaries"""
    lengths = [0, 1, 6, 7, 8, 9, 39, 40, 41, 63, 64, 65, 127, 128, 129, 1000, 16383, 16384, 16385,
               16384 * 2 + 17, 50_000, 3, 70_001]
    contigs, _ = synth.assembly(11, lengths, motifs, plant_every=300, n_runs=1, soft_mask=0.3)
    check(engine, _strs(contigs), motifs)


def test_scan_edge_planted(engine, motifs):
    """RSS units flush with contig starts / ends and straddling chunk and tile boundaries"""
    rng = np.random.default_rng(5)
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in SIG}
    contigs = []
    for L in (64 * 256 * 2 + 100, 5000, 1000, 200):
        buf = synth.background(rng, L)
        spots = [0, 1, L, 64 * 256, 64 * 256 * 2, 64, 128, 4096] + list(rng.integers(0, L, 20))
        for sp in spots:
            t = SIG[rng.integers(4)]
            u = np.frombuffer(synth.rss_unit(rng, motifs, t, sets).encode(), np.uint8)
            if rng.integers(2):
                u = synth.revcomp_bytes(u)
            for p in (sp - len(u), sp - len(u) // 2, sp - 3, sp):
                p = int(min(max(p, 0), L - len(u)))
                if rng.random() < 0.5:
                    buf[p:p + len(u)] = u
        contigs.append(buf)
    check(engine, _strs(contigs), motifs)


def test_scan_invalid_letters(engine, motifs):
    """N / IUPAC letters inside heptamers, nonamers and spacers; lower-case everywhere"""
    rng = np.random.default_rng(9)
    contigs, _ = synth.assembly(9, [40_000], motifs, plant_every=120)
    buf = contigs[0]
    for p in rng.integers(0, len(buf), 600):
        buf[p] = ord("NRYKMnry"[rng.integers(8)])
    buf[1000:9000] |= 0x20
    check(engine, _strs([buf]), motifs)


def test_scan_reverse_complement_symmetry(engine, motifs):
    """size-independent property: scanning RC(contig) mirrors every hit onto the other strand"""
    contigs, _ = synth.assembly(21, [3_000_000], motifs, plant_every=5000, n_runs=2, soft_mask=0.2)
    a = contigs[0]
    L = len(a)
    engine.load_contigs([a.tobytes()])
    h1 = engine.scan_rss(O.SPACER_LENGTH)
    engine.load_contigs([synth.revcomp_bytes(a).tobytes()])
    h2 = engine.scan_rss(O.SPACER_LENGTH)
    s1 = {(int(r["sig_type"]), int(r["strand"]), int(r["hept"]), int(r["nona"])) for r in h1}
    s2 = {(int(r["sig_type"]), 1 - int(r["strand"]), L - int(r["hept"]) - 7, L - int(r["nona"]) - 9) for r in h2}
    assert len(s1) == len(h1) and len(s1) > 500
    assert s1 == s2


def test_scan_sharding_invariance(engine, motifs):
    """hits do not depend on which contigs share a load (what multi-GPU sharding relies on)"""
    lengths = [200_000, 1_000, 64_000, 333_333, 5_000, 90_001]
    contigs, _ = synth.assembly(31, lengths, motifs, plant_every=2000, n_runs=1)
    bs = [c.tobytes() for c in contigs]
    engine.load_contigs(bs)
    whole = engine.scan_rss(O.SPACER_LENGTH)
    parts = []
    for idx in ([0, 2, 4], [1, 3, 5]):
        engine.load_contigs([bs[i] for i in idx])
        h = engine.scan_rss(O.SPACER_LENGTH).copy()
        h["contig"] = np.asarray(idx, np.uint32)[h["contig"]]
        parts.append(h)
    merged = np.concatenate(parts)
    key = lambda h: sorted(map(tuple, h[["contig", "sig_type", "strand", "hept", "nona", "spacer"]].tolist()))
    assert key(merged) == key(whole)


def test_scan_fragmented_assembly(engine, motifs):
    """config 5 in miniature: thousands of short, skewed-length contigs with RSS units at their edges;
    every contig up to 20 kb is checked against the oracle, the rest through bounds and ordering"""
    rng = np.random.default_rng(55)
    n = 6000
    lengths = np.minimum((1000 * (1 - rng.random(n)) ** (-1 / 1.1)).astype(np.int64), 2_000_000)
    lengths[:50] = rng.integers(0, 80, 50)                      # a few degenerate ones
    contigs, _ = synth.assembly(55, lengths.tolist(), motifs, plant_every=4000)
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in SIG}
    for c in contigs[::20]:                                      # units flush with both ends
        for at_end in (False, True):
            u = np.frombuffer(synth.rss_unit(rng, motifs, SIG[rng.integers(4)], sets).encode(), np.uint8)
            if rng.integers(2):
                u = synth.revcomp_bytes(u)
            if len(c) >= len(u):
                p = len(c) - len(u) if at_end else 0
                c[p:p + len(u)] = u
    engine.load_contigs([c.tobytes() for c in contigs])
    h = engine.scan_rss(O.SPACER_LENGTH)
    L = lengths[h["contig"]]
    assert np.all(h["hept"] >= 0) and np.all(h["hept"] + 7 <= L)
    assert np.all(h["nona"] >= 0) and np.all(h["nona"] + 9 <= L)
    key = h["contig"].astype(np.int64) * 16 + h["sig_type"] * 2 + h["strand"]
    assert np.all(np.diff(key) >= 0)
    small = [i for i in range(n) if lengths[i] <= 20_000]
    seqs = {i: contigs[i].tobytes().decode() for i in small}
    got = {}
    for r in h:
        got.setdefault((int(r["contig"]), SIG[r["sig_type"]], "+-"[r["strand"]]), []).append((int(r["hept"]), int(r["nona"])))
    nchecked = 0
    for t in SIG:
        for sd in "+-":
            want = O.get_contigwise_rss(t, sd, seqs, motifs, order="canonical")
            for i in small:
                assert sorted(got.get((i, t, sd), [])) == sorted(want[i]), (i, t, sd)
                nchecked += len(want[i])
    assert nchecked > 300


def test_scan_large_properties(engine, motifs):
    """~0.4 Gb in 40 scaffolds: hit set invariant under reverse complement and under resharding
    (too big for the CPU oracle, so only size-independent properties are checked)"""
    rng = np.random.default_rng(77)
    lengths = (rng.uniform(2e6, 2e7, 40)).astype(np.int64).tolist()
    contigs, _ = synth.assembly(77, lengths, motifs, plant_every=50_000, n_runs=1, soft_mask=0.1)
    bs = [c.tobytes() for c in contigs]
    engine.load_contigs(bs)
    h = engine.scan_rss(O.SPACER_LENGTH)
    assert len(h) > 10_000
    full = set(map(tuple, h[["contig", "sig_type", "strand", "hept", "nona"]].tolist()))
    assert len(full) == len(h)
    # reverse complement of every contig
    engine.load_contigs([synth.revcomp_bytes(c).tobytes() for c in contigs])
    r = engine.scan_rss(O.SPACER_LENGTH)
    Lr = np.asarray(lengths)[r["contig"]]
    mirrored = set(zip(r["contig"].tolist(), r["sig_type"].tolist(), (1 - r["strand"]).tolist(),
                       (Lr - r["hept"] - 7).tolist(), (Lr - r["nona"] - 9).tolist()))
    assert mirrored == full
    # two shards
    parts = set()
    for idx in (list(range(0, 40, 2)), list(range(1, 40, 2))):
        engine.load_contigs([bs[i] for i in idx])
        p = engine.scan_rss(O.SPACER_LENGTH)
        parts |= set(zip(np.asarray(idx)[p["contig"]].tolist(), p["sig_type"].tolist(), p["strand"].tolist(),
                         p["hept"].tolist(), p["nona"].tolist()))
    assert parts == full


def test_scan_arbitrary_motif_tables_and_spacers(motifs):
    """tables that break the consensus property (-> exhaustive kernel instance) and four different
    spacers (-> eight distinct nonamer windows): still exact"""
    from igdetective_b200.engine import Engine
    rng = np.random.default_rng(123)

    def rand_set(k, n):
        return {"".join("ACGT"[i] for i in rng.integers(0, 4, k)) for _ in range(n)}
    custom = {t: {"7": rand_set(7, 60), "9": rand_set(9, 2500)} for t in SIG}
    custom["V"]["7"] |= {"AAAAAAA", "TTTTTTT"}          # also hit by packed N runs / padding ('A' code)
    custom["J"]["9"] |= {"AAAAAAAAA"}
    spacers = {"V": 20, "D_left": 5, "D_right": 9, "J": 14}
    contigs, _ = synth.assembly(8, [150_000, 777, 20_001], motifs, n_runs=2, soft_mask=0.2)
    # plant units of the custom sets
    for c in contigs:
        for _ in range(len(c) // 400):
            t = SIG[rng.integers(4)]
            h = sorted(custom[t]["7"])[rng.integers(len(custom[t]["7"]))]
            n = sorted(custom[t]["9"])[rng.integers(len(custom[t]["9"]))]
            sp = "".join("ACGT"[i] for i in rng.integers(0, 4, spacers[t] + int(rng.integers(-1, 2))))
            u = np.frombuffer((h + sp + n if synth.HEPT_FIRST[t] else n + sp + h).encode(), np.uint8)
            if rng.integers(2):
                u = synth.revcomp_bytes(u)
            if len(c) > len(u):
                p = int(rng.integers(0, len(c) - len(u)))
                c[p:p + len(u)] = u
    seqs = {"c%d" % i: c.tobytes().decode() for i, c in enumerate(contigs)}
    with Engine(0, custom) as eng:
        eng.load_contigs(seqs)
        h = eng.scan_rss(spacers)
        got = {}
        for r in h:
            got.setdefault((int(r["contig"]), SIG[r["sig_type"]], "+-"[r["strand"]]), []).append((int(r["hept"]), int(r["nona"])))
        total = 0
        for t in SIG:
            for sd in "+-":
                want = O.get_contigwise_rss(t, sd, seqs, custom, order="canonical", spacers=spacers)
                for ci, c in enumerate(seqs):
                    assert sorted(got.get((ci, t, sd), [])) == sorted(want[c]), (ci, t, sd)
                    total += len(want[c])
        assert total > 200
        for k in (7, 9):
            kh = eng.scan_kmers(k)
            s = seqs["c0"]
            sel = kh[kh["contig"] == 0]
            for ti, t in enumerate(SIG):
                assert sel["pos"][(sel["flags"] >> ti) & 1 == 1].tolist() == O.find_valid_motif_idx(s, custom[t][str(k)], k)


def test_scan_pathological_density_and_buffer_regrow(engine, motifs):
    """a contig made of back-to-back RSS units: ~1.5 M hits overflow the initial hit buffer (regrow +
    rerun) and every warp queue / item list runs at its densest; prefix checked against the oracle,
    the rest through periodicity"""
    h7 = sorted(motifs["D_right"]["7"])[3]
    n9 = sorted(motifs["D_right"]["9"])[5]
    unit = h7 + "ACGTTGCAAGCT" + n9                     # 7 + 12 + 9 = 28
    assert len(unit) == 28
    n_units = 1_500_000
    s = unit * n_units
    engine.load_contigs([s])
    h = engine.scan_rss(O.SPACER_LENGTH)
    assert len(h) > 1_400_000                            # > max(2^20, bases / 32): the buffer had to grow
    want = O.scan_all({"c": s[:28 * 2000]}, motifs, order="canonical")
    lim = 28 * 2000 - 100
    for ti, t in enumerate(SIG):
        for si, sd in enumerate("+-"):
            sel = h[(h["sig_type"] == ti) & (h["strand"] == si)]
            got = [(int(a), int(b)) for a, b in zip(sel["hept"], sel["nona"]) if 100 <= a < lim and 100 <= b < lim]
            exp = [(a, b) for a, b in want[t][sd]["c"] if 100 <= a < lim and 100 <= b < lim]
            assert sorted(got) == sorted(exp), (t, sd)
    # periodicity: away from the ends every unit contributes the same hits
    mid = h[(h["hept"] >= 28 * 1000) & (h["hept"] < 28 * (n_units - 1000))]
    per_unit = len(mid) / (n_units - 2000)
    assert per_unit == int(per_unit) and per_unit >= 1
    ph = np.unique((mid["hept"] % 28).astype(np.int64) * 64 + mid["sig_type"] * 2 + mid["strand"], return_counts=True)[1]
    assert np.all(ph == n_units - 2000)
    kh = engine.scan_kmers(7)
    assert len(kh) >= n_units


@pytest.mark.parametrize("spacers", [{"V": 22, "D_left": 11, "D_right": 12, "J": 23},
                                     {"V": 23, "D_left": 12, "D_right": 9, "J": 23},
                                     {"V": 12, "D_left": 12, "D_right": 12, "J": 12},
                                     {"V": 2, "D_left": 23, "D_right": 3, "J": 17}])
def test_scan_shipped_motifs_other_spacers(engine, motifs, spacers):
    """the shipped motif tables (consensus prefilter on) with spacers other than 23/12/12/23: five to eight
    nonamer windows, other fetch groups (the generic window loop instead of the two-group instance), one
    window only; bit-exact against the oracle's loops with the same spacers"""
    rng = np.random.default_rng(sum(spacers.values()))
    contigs, _ = synth.assembly(11, [120_000, 4_100, 63, 30_001], motifs, n_runs=2, soft_mask=0.3)
    for c in contigs:
        for _ in range(len(c) // 300):
            t = SIG[rng.integers(4)]
            h = sorted(motifs[t]["7"])[rng.integers(len(motifs[t]["7"]))]
            n = sorted(motifs[t]["9"])[rng.integers(len(motifs[t]["9"]))]
            sp = "".join("ACGT"[i] for i in rng.integers(0, 4, max(0, spacers[t] + int(rng.integers(-1, 2)))))
            u = np.frombuffer((h + sp + n if synth.HEPT_FIRST[t] else n + sp + h).encode(), np.uint8)
            if rng.integers(2):
                u = synth.revcomp_bytes(u)
            if len(c) > len(u):
                p = int(rng.integers(0, len(c) - len(u)))
                c[p:p + len(u)] = u
    seqs = {"c%d" % i: c.tobytes().decode() for i, c in enumerate(contigs)}
    engine.load_contigs(seqs)
    h = engine.scan_rss(spacers)
    got = {}
    for r in h:
        got.setdefault((int(r["contig"]), SIG[r["sig_type"]], "+-"[r["strand"]]), []).append((int(r["hept"]), int(r["nona"])))
    total = 0
    for t in SIG:
        for sd in "+-":
            want = O.get_contigwise_rss(t, sd, seqs, motifs, order="canonical", spacers=spacers)
            for ci, c in enumerate(seqs):
                assert sorted(got.get((ci, t, sd), [])) == sorted(want[c]), (ci, t, sd)
                total += len(want[c])
    assert total > 100
    # and back to the reference's spacers on the same handle (the window-need table follows the spacers)
    engine.load_contigs(seqs)
    check(engine, seqs, motifs)


def test_upload_pageable_ring_equals_pinned(engine, motifs):
    """igd_load_contigs moves a large pageable genome through a ring of pinned pieces (several memcpy threads, one
    DMA per piece); the same bytes from pinned memory go in one DMA: identical hits, on one contig that spans
    several pieces and on many contigs whose boundaries fall inside pieces"""
    import torch
    contigs, _ = synth.assembly(21, [1_000_003], motifs, plant_every=5_000, n_runs=3, soft_mask=0.3)
    big = np.tile(contigs[0], 150)                         # 150 Mb > 4 pieces of 32 MiB
    n = len(big)
    pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = big
    for off in (np.array([0, n], np.uint64),
                np.array(sorted({0, n} | set(np.random.default_rng(3).integers(1, n, 300).tolist())), np.uint64)):
        engine.load_contigs((big, off), list(range(len(off) - 1)))
        a = engine.scan_rss(O.SPACER_LENGTH)
        engine.load_contigs((pinned.data_ptr(), off), list(range(len(off) - 1)))
        b = engine.scan_rss(O.SPACER_LENGTH)
        assert len(a) == len(b) > 10_000
        assert a.tobytes() == b.tobytes()


def test_scan_is_deterministic_across_launches(engine, motifs):
    """warps pick strips from a global counter and queue candidates in arrival order: whichever way a launch
    interleaves, the sorted hit list is the same -- twelve launches on one input, byte for byte (a race in the
    queues or in the key buffers would show up here as a lost or doubled hit)"""
    contigs, _ = synth.assembly(31, [3_000_000, 40_000, 700_001], motifs, plant_every=900, n_runs=4, soft_mask=0.4)
    seqs = {"c%d" % i: c.tobytes().decode() for i, c in enumerate(contigs)}
    engine.load_contigs(seqs)
    first = engine.scan_rss(O.SPACER_LENGTH).tobytes()
    assert len(first) > 4000 * 24
    for _ in range(11):
        assert engine.scan_rss(O.SPACER_LENGTH).tobytes() == first
