"""Seeded synthetic inputs shared by tests and bench.py: i.i.d. background with
P(A,C,G,T) = (0.295, 0.205, 0.205, 0.295), planted RSS units / mutated genes, N runs, soft-masking."""
import numpy as np

_COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[a] = b
SIG = ["V", "D_left", "D_right", "J"]
HEPT_FIRST = {"V": True, "D_right": True, "J": False, "D_left": False}
SPACER = {"V": 23, "D_left": 12, "D_right": 12, "J": 23}


def background(rng, n):
    return np.frombuffer(b"ACGT", np.uint8)[rng.choice(4, size=n, p=[0.295, 0.205, 0.205, 0.295])].copy()


def revcomp_bytes(a):
    return _COMP[a[::-1]]


def mutate(rng, seq, sub, indel):
    out = []
    for ch in seq:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(ch); out.append("ACGT"[rng.integers(4)]); continue
        out.append("ACGT"[rng.integers(4)] if rng.random() < sub else ch)
    return "".join(out)


def rss_unit(rng, motifs, t, sets=None):
    sets = sets or {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in SIG}
    h = sets[t][0][rng.integers(len(sets[t][0]))]
    n = sets[t][1][rng.integers(len(sets[t][1]))]
    sp = "".join("ACGT"[i] for i in rng.integers(0, 4, SPACER[t] + int(rng.integers(-1, 2))))
    return h + sp + n if HEPT_FIRST[t] else n + sp + h


def plant_locus(rng, buf, motifs, genes_v, genes_j, n_v, n_d, n_j, lo, hi):
    """Writes V / D / J segments with their RSSs into buf[lo:hi] on random strands; returns truth rows."""
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in SIG}
    vs, js = list(genes_v.values()), list(genes_j.values())
    truth = []

    def put(unit, kind):
        u = np.frombuffer(unit.encode(), np.uint8)
        strand = "+"
        if rng.integers(2):
            u = revcomp_bytes(u); strand = "-"
        p = int(rng.integers(lo, max(lo + 1, hi - len(u))))
        buf[p:p + len(u)] = u
        truth.append((kind, p, len(u), strand))
    for _ in range(n_v):
        g = mutate(rng, vs[rng.integers(len(vs))], rng.uniform(0.02, 0.25), rng.uniform(0, 0.02))
        put(g + rss_unit(rng, motifs, "V", sets), "V")
    for _ in range(n_d):
        d = "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(8, 40))))
        put(rss_unit(rng, motifs, "D_left", sets) + d + rss_unit(rng, motifs, "D_right", sets), "D")
    for _ in range(n_j):
        g = mutate(rng, js[rng.integers(len(js))], rng.uniform(0.02, 0.15), rng.uniform(0, 0.02))
        put(rss_unit(rng, motifs, "J", sets) + g, "J")
    return truth


def assembly(seed, lengths, motifs, genes_v=None, genes_j=None, plant_every=0, n_runs=0, soft_mask=0.0,
             clusters=None):
    """Returns (list of uint8 arrays, truth).  plant_every: one bare RSS unit per that many bases;
    clusters: [(contig index, n_v, n_d, n_j)] loci planted with mutated genes."""
    rng = np.random.default_rng(seed)
    sets = {x: (sorted(motifs[x]["7"]), sorted(motifs[x]["9"])) for x in SIG}
    contigs, truth = [], []
    for ci, L in enumerate(lengths):
        buf = background(rng, L)
        if plant_every and L > 200:
            for _ in range(max(1, L // plant_every)):
                t = SIG[rng.integers(4)]
                u = np.frombuffer(rss_unit(rng, motifs, t, sets).encode(), np.uint8)
                if rng.integers(2):
                    u = revcomp_bytes(u)
                p = int(rng.integers(0, L - len(u)))
                buf[p:p + len(u)] = u
        contigs.append(buf)
    for (ci, nv, nd, nj) in (clusters or []):
        L = len(contigs[ci])
        lo = int(rng.integers(0, max(1, L - min(L, 2_000_000))))
        truth += [(ci,) + r for r in plant_locus(rng, contigs[ci], motifs, genes_v, genes_j, nv, nd, nj,
                                                 lo, min(L, lo + 2_000_000))]
    for buf in contigs:
        L = len(buf)
        for _ in range(n_runs):
            n = int(min(L // 4, rng.integers(100, 50_000)))
            p = int(rng.integers(0, max(1, L - n)))
            buf[p:p + n] = ord("N")
        if soft_mask > 0 and L > 1000:
            # lower-case blocks covering ~soft_mask of the contig
            nblk = max(1, int(L * soft_mask / 5000))
            for _ in range(nblk):
                n = int(min(L // 2, rng.integers(500, 9500)))
                p = int(rng.integers(0, max(1, L - n)))
                buf[p:p + n] |= 0x20
    return contigs, truth
