"""Pure-Python models of what the CUDA kernels compute, used by the CPU tests to check the
ALGORITHMS (heptamer-gated forward-only scan; key/payload Gotoh) against the oracle before any GPU
time is spent.  Test infrastructure only."""
import numpy as np

SIG = ["V", "D_left", "D_right", "J"]
HEPT_FIRST = {"V": True, "D_right": True, "J": False, "D_left": False}
_COMP = str.maketrans("ACGT", "TGCA")


def _rc(s):
    return s.translate(_COMP)[::-1]


def scan_model(seq, motifs, spacers):
    """rss_scan.cu's formulation: forward coordinates only, reverse strand through RC'd motif
    sets, pairing gated on heptamer hits.  Returns {(type, strand): set of (hept, nona)}."""
    s = seq.upper()
    L = len(s)
    H = {t: {"+": motifs[t]["7"], "-": {_rc(m) for m in motifs[t]["7"]}} for t in SIG}
    N = {t: {"+": motifs[t]["9"], "-": {_rc(m) for m in motifs[t]["9"]}} for t in SIG}

    def hept(x, t, sd):
        return 0 <= x <= L - 7 and s[x:x + 7] in H[t][sd]

    def nona(x, t, sd):
        return 0 <= x <= L - 9 and s[x:x + 9] in N[t][sd]
    out = {(t, sd): set() for t in SIG for sd in "+-"}
    for q in range(L - 6):
        for t in SIG:
            sp = spacers[t]
            cand = (sp, sp - 1, sp + 1)
            if hept(q, t, "+"):
                if HEPT_FIRST[t]:
                    for s1 in cand:
                        if nona(q + 7 + s1, t, "+"):
                            out[(t, "+")].add((q, q + 7 + s1)); break
                else:
                    if nona(q - 9 - sp, t, "+"):
                        out[(t, "+")].add((q, q - 9 - sp))
                    if not hept(q + 1, t, "+") and nona(q - 9 - (sp - 1), t, "+"):
                        out[(t, "+")].add((q, q - 9 - (sp - 1)))
                    if not hept(q - 1, t, "+") and not hept(q - 2, t, "+") and nona(q - 9 - (sp + 1), t, "+"):
                        out[(t, "+")].add((q, q - 9 - (sp + 1)))
            if hept(q, t, "-"):
                if HEPT_FIRST[t]:
                    for s1 in cand:
                        if nona(q - s1 - 9, t, "-"):
                            out[(t, "-")].add((q, q - s1 - 9)); break
                else:
                    if nona(q + 7 + sp, t, "-"):
                        out[(t, "-")].add((q, q + 7 + sp))
                    if not hept(q - 1, t, "-") and nona(q + 7 + (sp - 1), t, "-"):
                        out[(t, "-")].add((q, q + 7 + (sp - 1)))
                    if not hept(q + 1, t, "-") and not hept(q + 2, t, "-") and nona(q + 7 + (sp + 1), t, "-"):
                        out[(t, "-")].add((q, q + 7 + (sp + 1)))
    return out


def scan_windows(spacers):
    """igd_launch_rss_scan's window table: one window per distinct (side, spacer), ascending offset.
    Returns [(woff, {(type, strand), ...})]: the three candidate nonamer starts of a window are
    q + woff, q + woff + 1, q + woff + 2 for a heptamer at q."""
    wins = {}
    for t in SIG:
        sp = spacers[t]
        right, left = 7 + sp - 1, -9 - sp - 1
        if HEPT_FIRST[t]:
            wins.setdefault(right, set()).add((t, "+"))
            wins.setdefault(left, set()).add((t, "-"))
        else:
            wins.setdefault(left, set()).add((t, "+"))
            wins.setdefault(right, set()).add((t, "-"))
    return sorted(wins.items())


def centre_set(motifs):
    """igd_center_bitmap as a set of strings: a nonamer of any set on either strand, or a 9-mer that
    overlaps one shifted by one base to either side."""
    out = set()
    for t in SIG:
        for m in motifs[t]["9"]:
            for x in (m, _rc(m)):
                out.add(x)
                for a in "ACGT":
                    out.add(x[1:] + a)
                    out.add(a + x[:-1])
    return out


def scan_cascade_model(seq, motifs, spacers):
    """rss_scan.cu's cascade after the prefilter: heptamer test -> window-need bits -> ONE centre probe per
    needed window -> exact resolution of the windows that pass (three nonamer starts, the reference's ranking).
    Letters outside ACGT must already be '#'.  Returns {(type, strand): set of (hept, nona)}."""
    s = seq.upper()
    L = len(s)
    H = {(t, "+"): motifs[t]["7"] for t in SIG}
    H.update({(t, "-"): {_rc(m) for m in motifs[t]["7"]} for t in SIG})
    N = {(t, "+"): motifs[t]["9"] for t in SIG}
    N.update({(t, "-"): {_rc(m) for m in motifs[t]["9"]} for t in SIG})
    wins = scan_windows(spacers)
    centres = centre_set(motifs)
    # the packed planes hold code 0 ('A') for invalid letters: the probe sees them as 'A' (conservative)
    probe_view = s.replace("#", "A")

    def flags7(x):
        return {k for k in H if 0 <= x <= L - 7 and s[x:x + 7] in H[k]}

    def flags9(x):
        return {k for k in N if 0 <= x <= L - 9 and s[x:x + 9] in N[k]}
    out = {(t, sd): set() for t in SIG for sd in "+-"}
    for q in range(L - 6):
        f = flags7(q)
        if not f:
            continue
        for woff, served in wins:
            fm = f & served                                   # need bit of this window
            if not fm:
                continue
            c = q + woff + 1
            centre = probe_view[max(c, 0):c + 9] if c >= 0 else ""
            if len(centre) == 9 and centre not in centres:    # windows hanging over an end go to the exact path
                continue
            x = q + woff
            nf = [flags9(x), flags9(x + 1), flags9(x + 2)]
            right = woff > 0
            for key in fm:
                t, sd = key
                n = [key in nf[0], key in nf[1], key in nf[2]]
                if not any(n):
                    continue
                # window positions ascend: on the right spacers s-1, s, s+1, on the left s+1, s, s-1
                minus, plus = (n[0], n[2]) if right else (n[2], n[0])
                pos = {"mid": x + 1, "minus": x if right else x + 2, "plus": x + 2 if right else x}
                if HEPT_FIRST[t]:
                    pick = "mid" if n[1] else ("minus" if minus else "plus")
                    out[key].add((q, pos[pick]))
                else:
                    h = lambda d: key in flags7(q + d)
                    if n[1]:
                        out[key].add((q, pos["mid"]))
                    if sd == "+":
                        if minus and not h(1):
                            out[key].add((q, pos["minus"]))
                        if plus and not h(-1) and not h(-2):
                            out[key].add((q, pos["plus"]))
                    else:
                        if minus and not h(-1):
                            out[key].add((q, pos["minus"]))
                        if plus and not h(1) and not h(2):
                            out[key].add((q, pos["plus"]))
    return out


NEG = -(1 << 26)


def gotoh_model(A, B, sc, pay_mode=0):
    """gotoh.cu's recurrence: integer keys 4*score+priority, payload selected by the same compare.
    sc: dict with the ten igd_scoring fields.  Returns (score, matches, lead_or_trail, intIx)."""
    A = A.encode("latin-1") if isinstance(A, str) else A
    B = B.encode("latin-1") if isinstance(B, str) else B
    nA, nB = len(A), len(B)
    up = lambda c: c - 32 if 97 <= c <= 122 else c
    m4, x4 = 4 * sc["match"], 4 * sc["mismatch"]
    to4, te4 = 4 * sc["target_internal_open"], 4 * sc["target_internal_extend"]
    qio4, qie4 = 4 * sc["query_internal_open"], 4 * sc["query_internal_extend"]
    qeo4, qee4 = 4 * sc["query_end_open"], 4 * sc["query_end_extend"]
    u_int, u_trail = 1 << 10, ((1 << 20) if pay_mode == 1 else 0)
    lead = (lambda i: i << 20) if pay_mode == 0 else (lambda i: 0)
    # column state, indexed by row i (0..nA): D key/pay, Iy key/pay for column j-1
    Dk = [2] + [(qeo4 + qee4 * (i - 1)) | 1 for i in range(1, nA + 1)]
    Dp = [0] + [lead(i) for i in range(1, nA + 1)]
    Yk = [NEG] * (nA + 1)
    Yp = [0] * (nA + 1)
    for j in range(1, nB + 1):
        last = j == nB
        qo4, qe4 = (qeo4, qee4 + 1) if last else (qio4, qie4 + 1)
        ux = u_trail if last else u_int
        nDk, nDp, nYk, nYp = [0] * (nA + 1), [0] * (nA + 1), [0] * (nA + 1), [0] * (nA + 1)
        nDk[0], nDp[0], nYk[0], nYp[0] = to4 + te4 * (j - 1), 0, to4 + te4 * (j - 1), 0
        xk, xp = NEG, 0
        for i in range(1, nA + 1):
            a, b = A[i - 1], B[j - 1]
            mk = ((Dk[i - 1] & ~3) | 2) + (m4 if a == b else x4)
            mp = Dp[i - 1] + (1 if up(a) == up(b) else 0)
            c1, c2 = nDk[i - 1] + qo4, xk + qe4
            nxk = max(c1, c2) & ~3
            nxp = (nDp[i - 1] if c1 >= c2 else xp) + ux
            e1, e2 = Dk[i] + to4, Yk[i] + te4
            nyk = max(e1, e2) & ~3
            nyp = Dp[i] if e1 >= e2 else Yp[i]
            xk1 = nxk | 1
            bk, bp = (mk, mp) if mk >= xk1 else (xk1, nxp)
            if not bk >= nyk:
                bk, bp = nyk, nyp
            nDk[i], nDp[i], nYk[i], nYp[i] = bk, bp, nyk, nyp
            xk, xp = nxk, nxp
        Dk, Dp, Yk, Yp = nDk, nDp, nYk, nYp
    k, p = Dk[nA], Dp[nA]
    return k >> 2, p & 1023, (p >> 20) & 1023, (p >> 10) & 1023


# ----------------------------------------------------------------------------- pruning bounds (csrc/lcs.cu)
def _classes(s):
    """letter classes of the bound kernels: A, C, G, T (either case) and one class for everything else"""
    s = np.frombuffer(s.upper().encode(), np.uint8)
    c = np.full(len(s), 4, np.int64)
    for k, ch in enumerate(b"ACGT"):
        c[s == ch] = k
    return c


def lcs_bound(a, b):
    """K4: length of the longest common subsequence of the class strings (plain row-by-row recurrence)"""
    ca, cb = _classes(a), _classes(b)
    prev = np.zeros(len(cb) + 1, np.int64)
    for x in ca:
        t = np.maximum(prev[:-1] + (cb == x), prev[1:])
        prev = np.concatenate([[0], np.maximum.accumulate(t)])
    return int(prev[-1])


def y_bound(a, b):
    """K5: max over all alignments of 2 x equal columns - fragment letters inserted inside the gene range, as the kernel
    states it: H(i,j) = max(H(i-1,j-1) + 2 [a_i ~ b_j], H(i-1,j) - [j < len(b)], H(i,j-1)), H(i,0) = H(0,j) = 0, answer
    H(len(a), len(b)) + 3 x (gene letters outside ACGT).  a_i ~ b_j: equal classes, or a_i outside ACGT (taken to match
    every A, C, G, T); a gene letter outside ACGT matches nothing and its diagonal move costs 1 -- the + 3 covers its
    possible match and that 1, so the value is >= the plain recurrence in which such letters match each other."""
    ca, cb = _classes(a), _classes(b)
    pen = np.ones(len(cb), np.int64)
    pen[-1] = 0
    other = cb == 4
    prev = np.zeros(len(cb) + 1, np.int64)
    for x in ca:
        eq = ((cb == x) | (x == 4)) & ~other
        t = np.maximum(prev[:-1] + 2 * eq - other, prev[1:] - pen)
        prev = np.concatenate([[0], np.maximum.accumulate(np.maximum(t, 0))])
    return int(prev[-1]) + 3 * int(other.sum())
