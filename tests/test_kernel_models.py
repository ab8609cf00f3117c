"""The algorithms the CUDA kernels implement, modelled in Python, against the oracle (CPU only)."""
import numpy as np
import pytest

from oracle import igd_oracle as O
import kernel_model as KM

SC = dict(match=2, mismatch=0, target_internal_open=-2, target_internal_extend=-1,
          target_end_open=-2, target_end_extend=-1, query_internal_open=-2, query_internal_extend=-1,
          query_end_open=0, query_end_extend=0)


def _plant(rng, motifs, L, n_plant):
    s = list(rng.choice(list("ACGT"), L, p=[0.295, 0.205, 0.205, 0.295]))
    for _ in range(n_plant):
        t = KM.SIG[rng.integers(4)]
        h = sorted(motifs[t]["7"])[rng.integers(len(motifs[t]["7"]))]
        n = sorted(motifs[t]["9"])[rng.integers(len(motifs[t]["9"]))]
        sp = O.SPACER_LENGTH[t] + int(rng.integers(-1, 2))
        spacer = "".join(rng.choice(list("ACGT"), sp))
        unit = h + spacer + n if KM.HEPT_FIRST[t] else n + spacer + h
        if rng.integers(2):
            unit = O.revcomp(unit)
        p = int(rng.integers(0, L - len(unit)))
        s[p:p + len(unit)] = list(unit)
    s = "".join(s)
    # soft-masking, an N run and a stray IUPAC code
    a = int(rng.integers(0, L - 50))
    s = s[:a] + s[a:a + 40].lower() + s[a + 40:]
    b = int(rng.integers(0, L - 30))
    s = s[:b] + "N" * 17 + s[b + 17:]
    c = int(rng.integers(0, L))
    return s[:c] + "R" + s[c + 1:]


@pytest.mark.parametrize("seed", range(6))
def test_scan_model_equals_oracle(motifs, seed):
    rng = np.random.default_rng(seed)
    seq = _plant(rng, motifs, 3000, 40)
    model = KM.scan_model("".join(ch if ch.upper() in "ACGT" else "#" for ch in seq), motifs, O.SPACER_LENGTH)
    for t in KM.SIG:
        for sd in "+-":
            want = O.get_contigwise_rss(t, sd, {"c": seq}, motifs, order="canonical")["c"]
            assert len(set(want)) == len(want)
            assert set(want) == model[(t, sd)], (t, sd)


@pytest.mark.parametrize("spacers,n_windows", [(None, 4), ({"V": 22, "D_left": 11, "D_right": 12, "J": 23}, 8),
                                               ({"V": 12, "D_left": 12, "D_right": 12, "J": 12}, 2),
                                               ({"V": 2, "D_left": 23, "D_right": 3, "J": 17}, 8)])
def test_scan_cascade_model_equals_oracle(motifs, spacers, n_windows):
    """the kernel's cascade (window table with merged windows, need bits, one centre probe per window, per-window
    exact resolution) against the reference's loops, for the reference's spacers and for sets that merge or
    split the windows differently"""
    sp = dict(O.SPACER_LENGTH) if spacers is None else spacers
    wins = KM.scan_windows(sp)
    assert len(wins) == n_windows and [w for w, _ in wins] == sorted(w for w, _ in wins)
    total = 0
    for seed in range(3):
        rng = np.random.default_rng(100 + seed)
        s = list(rng.choice(list("ACGT"), 4000, p=[0.295, 0.205, 0.205, 0.295]))
        for _ in range(60):
            t = KM.SIG[rng.integers(4)]
            h = sorted(motifs[t]["7"])[rng.integers(len(motifs[t]["7"]))]
            n = sorted(motifs[t]["9"])[rng.integers(len(motifs[t]["9"]))]
            spacer = "".join(rng.choice(list("ACGT"), max(0, sp[t] + int(rng.integers(-1, 2)))))
            unit = h + spacer + n if KM.HEPT_FIRST[t] else n + spacer + h
            if rng.integers(2):
                unit = O.revcomp(unit)
            p = int(rng.integers(0, len(s) - len(unit)))
            s[p:p + len(unit)] = list(unit)
        seq = "".join(s)
        b = int(rng.integers(0, len(seq) - 30))
        seq = seq[:b] + "N" * 11 + seq[b + 11:]
        model = KM.scan_cascade_model("".join(ch if ch in "ACGT" else "#" for ch in seq), motifs, sp)
        for t in KM.SIG:
            for sd in "+-":
                want = O.get_contigwise_rss(t, sd, {"c": seq}, motifs, order="canonical", spacers=sp)["c"]
                assert set(want) == model[(t, sd)], (t, sd, seed)
                total += len(want)
    assert total > 60


def _pairs(rng, n):
    out = []
    for _ in range(n):
        nb = int(rng.integers(1, 40))
        g = "".join(rng.choice(list("ACGT"), nb))
        kind = rng.integers(3)
        if kind == 0:
            a = "".join(rng.choice(list("ACGT"), int(rng.integers(1, 60))))
        else:
            mut = list(g)
            for _ in range(int(rng.integers(0, 6))):
                p = int(rng.integers(0, len(mut) + 1))
                op = rng.integers(3)
                if op == 0 and mut:
                    mut[min(p, len(mut) - 1)] = "ACGT"[rng.integers(4)]
                elif op == 1:
                    mut.insert(p, "ACGT"[rng.integers(4)])
                elif mut:
                    del mut[min(p, len(mut) - 1)]
            a = "".join(rng.choice(list("ACGT"), int(rng.integers(0, 15)))) + "".join(mut) + \
                "".join(rng.choice(list("ACGT"), int(rng.integers(0, 15))))
        if kind == 2:
            a = "".join(ch.lower() if rng.random() < 0.3 else ch for ch in a)
        if not a:
            a = "A"
        out.append((a, g))
    return out


@pytest.mark.parametrize("seed", range(4))
def test_gotoh_model_equals_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    pairs = _pairs(rng, 120)
    st = O.align_stats([p[0] for p in pairs], [p[1] for p in pairs], range(len(pairs)), range(len(pairs)))
    for k, (a, g) in enumerate(pairs):
        score, m, lead, intx = KM.gotoh_model(a, g, SC, 0)
        _, m2, trail, intx2 = KM.gotoh_model(a, g, SC, 1)
        s = st[k]
        assert score == int(s["score"]), (a, g)
        assert (m, m2) == (int(s["matches"]),) * 2, (a, g)
        assert lead == int(s["start"]) == int(s["lead"]), (a, g)
        assert lead + len(g) + intx == int(s["end"]), (a, g)
        assert intx == intx2
        assert len(a) - trail == int(s["ient"]), (a, g)


def test_pruning_bounds_hold_for_the_oracle_alignments():
    """csrc/lcs.cu's two recurrences bound what the aligner returns: m equal columns <= LCS and, with c letters inserted
    inside the gene range, 2 m - c <= Y -- related, diverged and unrelated pairs, lower-case and N letters included."""
    import synth
    rng = np.random.default_rng(21)
    genes = ["".join(rng.choice(list("ACGT"), n)) for n in (40, 96, 150)]
    genes.append(genes[1][:50].lower() + genes[1][50:])
    genes.append(genes[2][:70] + "NNR" + genes[2][73:])
    frags = []
    for i in range(9):
        g = genes[i % len(genes)].upper()
        body = synth.mutate(rng, g, 0.05 * i, 0.02 * (i % 3)) if i % 4 else ""
        f = "".join(rng.choice(list("ACGTN"), 60, p=[0.24, 0.24, 0.24, 0.24, 0.04])) + body + "".join(rng.choice(list("ACGT"), 30))
        frags.append(f[-170:])
    pt, pq = np.meshgrid(np.arange(len(frags)), np.arange(len(genes)), indexing="ij")
    st = O.align_stats(frags, genes, pt.ravel(), pq.ravel())
    k = 0
    tight = 0
    for f in frags:
        for g in genes:
            m, c = int(st["matches"][k]), int(st["end"][k]) - int(st["start"][k]) - len(g)
            assert c >= 0 and m <= KM.lcs_bound(f, g), (f, g)
            y = KM.y_bound(f, g)
            assert 2 * m - c <= y, (f, g, m, c, y)
            tight += 2 * m - c == y
            k += 1
    assert tight > 0                     # the bound is attained (closely related pairs), not vacuous


def test_pruning_theorems_on_random_short_sequences():
    """What igd_detect relies on, on many small random cases (where co-optimal alignments and end effects are common):
    m <= LCS;  2 m - c <= Y;  and the two tests it applies -- PI <= (LCS - 1) / len(gene) x 100, and PI >= p >= 50 implies
    50 Y >= p len(gene) + 100 -- for the alignment the oracle returns, in both orientations' worth of inputs."""
    rng = np.random.default_rng(33)
    frags, genes = [], []
    for _ in range(400):
        n, q = int(rng.integers(1, 48)), int(rng.integers(1, 40))
        g = "".join(rng.choice(list("ACGTNacgt"), q, p=[0.22, 0.22, 0.22, 0.22, 0.04, 0.02, 0.02, 0.02, 0.02]))
        if rng.random() < 0.5:                        # a related pair: the gene with a few edits, inside a random flank
            body = "".join(ch for ch in g.upper() if rng.random() > 0.1)
            f = "".join(rng.choice(list("ACGT"), int(rng.integers(0, 10)))) + body + "".join(rng.choice(list("ACGT"), int(rng.integers(0, 10))))
            f = f[:max(1, n + q)] or "A"
        else:
            f = "".join(rng.choice(list("ACGTN"), n, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        frags.append(f); genes.append(g)
    idx = np.arange(len(frags))
    st = O.align_stats(frags, genes, idx, idx)
    strong = 0
    for k, (f, g) in enumerate(zip(frags, genes)):
        m, alen = int(st["matches"][k]), int(st["end"][k]) - int(st["start"][k])
        c = alen - len(g)
        lcs, y = KM.lcs_bound(f, g), KM.y_bound(f, g)
        assert c >= 0 and m <= lcs and 2 * m - c <= y, (f, g, m, c, lcs, y)
        pi = (m - 1) / alen * 100
        assert pi <= (lcs - 1) / len(g) * 100 + 1e-9
        if pi >= 50:
            assert 50 * y >= pi * len(g) + 100 - 1e-9, (f, g, pi, y)
            strong += 1
    assert strong >= 60
