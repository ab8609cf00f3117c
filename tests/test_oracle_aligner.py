"""Self-consistency of the oracle aligner: an independent pure-Python Gotoh
with predecessor SETS enumerates the co-optimal alignments of small pairs; the C oracle's path must
be optimal, and must be the first one in M > Ix > Iy preference order.  CPU only."""
import ctypes

import numpy as np
import pytest

from oracle import igd_oracle as O

NEG = float("-inf")


def py_first_alignment(A, B, sc):
    nA, nB = len(A), len(B)
    M = [[NEG] * (nB + 1) for _ in range(nA + 1)]
    X = [[NEG] * (nB + 1) for _ in range(nA + 1)]
    Y = [[NEG] * (nB + 1) for _ in range(nA + 1)]
    tM = [[()] * (nB + 1) for _ in range(nA + 1)]
    tX = [[()] * (nB + 1) for _ in range(nA + 1)]
    tY = [[()] * (nB + 1) for _ in range(nA + 1)]

    def best(cands):
        m = max(c[0] for c in cands)
        return m, tuple(s for v, s in cands if v == m and v > NEG)
    M[0][0] = 0
    for j in range(1, nB + 1):
        Y[0][j] = sc["t_left_open"] + sc["t_left_ext"] * (j - 1)
        tY[0][j] = ("M",) if j == 1 else ("Y",)
    for i in range(1, nA + 1):
        X[i][0] = sc["q_left_open"] + sc["q_left_ext"] * (i - 1)
        tX[i][0] = ("M",) if i == 1 else ("X",)
        for j in range(1, nB + 1):
            qo, qe = (sc["q_right_open"], sc["q_right_ext"]) if j == nB else (sc["q_int_open"], sc["q_int_ext"])
            to, te = (sc["t_right_open"], sc["t_right_ext"]) if i == nA else (sc["t_int_open"], sc["t_int_ext"])
            v, t = best([(M[i - 1][j - 1], "M"), (X[i - 1][j - 1], "X"), (Y[i - 1][j - 1], "Y")])
            M[i][j], tM[i][j] = v + (sc["match"] if A[i - 1] == B[j - 1] else sc["mismatch"]), t
            X[i][j], tX[i][j] = best([(M[i - 1][j] + qo, "M"), (X[i - 1][j] + qe, "X"), (Y[i - 1][j] + qo, "Y")])
            Y[i][j], tY[i][j] = best([(M[i][j - 1] + to, "M"), (X[i][j - 1] + to, "X"), (Y[i][j - 1] + te, "Y")])
    score, ends = best([(M[nA][nB], "M"), (X[nA][nB], "X"), (Y[nA][nB], "Y")])
    i, j, st = nA, nB, ends[0] if ends else None
    rt, rq = [], []
    while st and (i > 0 or j > 0):
        if st == "M":
            tr = tM[i][j]; rt.append(A[i - 1]); rq.append(B[j - 1]); i -= 1; j -= 1
        elif st == "X":
            tr = tX[i][j]; rt.append(A[i - 1]); rq.append("-"); i -= 1
        else:
            tr = tY[i][j]; rt.append("-"); rq.append(B[j - 1]); j -= 1
        st = next((s for s in "MXY" if s in tr), None)
    return score, "".join(reversed(rt)), "".join(reversed(rq))


def score_of(rt, rq, sc, nA, nB):
    """score of a gapped pair under the scheme (end gaps by position)"""
    s, k, i, j = 0, 0, 0, 0
    n = len(rt)
    while k < n:
        if rt[k] != "-" and rq[k] != "-":
            s += sc["match"] if rt[k] == rq[k] else sc["mismatch"]; i += 1; j += 1; k += 1
            continue
        row = rq if rq[k] == "-" else rt
        e = k
        while e < n and row[e] == "-" and (rt if row is rq else rq)[e] != "-":
            e += 1
        ln = e - k
        if row is rq:      # query gap: column position j decides end vs internal
            o, x = (sc["q_left_open"], sc["q_left_ext"]) if j == 0 else \
                   (sc["q_right_open"], sc["q_right_ext"]) if j == nB else (sc["q_int_open"], sc["q_int_ext"])
            i += ln
        else:
            o, x = (sc["t_left_open"], sc["t_left_ext"]) if i == 0 else \
                   (sc["t_right_open"], sc["t_right_ext"]) if i == nA else (sc["t_int_open"], sc["t_int_ext"])
            j += ln
        s += o + x * (ln - 1)
        k = e
    return s


def c_first(A, B, sc):
    lib = O.lib()
    lib.igo_gotoh_first.restype = ctypes.c_int
    s = O._Scoring(**{k: float(v) for k, v in sc.items()})
    n = len(A) + len(B) + 1
    rt, rq = ctypes.create_string_buffer(n), ctypes.create_string_buffer(n)
    score, ncols = ctypes.c_double(), ctypes.c_int()
    rc = lib.igo_gotoh_first(A.encode(), len(A), B.encode(), len(B), ctypes.byref(s), ctypes.byref(score),
                             rt, rq, ctypes.byref(ncols))
    assert rc == 0
    return score.value, rt.raw[:ncols.value].decode(), rq.raw[:ncols.value].decode()


SCHEMES = [O.REFERENCE_SCORING,
           dict(match=1, mismatch=-1, t_int_open=-3, t_int_ext=-1, t_left_open=-3, t_left_ext=-1, t_right_open=-3,
                t_right_ext=-1, q_int_open=-2, q_int_ext=-2, q_left_open=0, q_left_ext=0, q_right_open=-1, q_right_ext=0)]


@pytest.mark.parametrize("scheme", range(len(SCHEMES)))
def test_c_oracle_equals_python_enumeration(scheme):
    sc = SCHEMES[scheme]
    rng = np.random.default_rng(scheme)
    for _ in range(300):
        A = "".join("ACGT"[i] for i in rng.integers(0, 2 + rng.integers(3), rng.integers(1, 14)))
        B = "".join("ACGT"[i] for i in rng.integers(0, 2 + rng.integers(3), rng.integers(1, 10)))
        ps, prt, prq = py_first_alignment(A, B, sc)
        cs, crt, crq = c_first(A, B, sc)
        assert cs == ps and (crt, crq) == (prt, prq), (A, B)
        assert crt.replace("-", "") == A and crq.replace("-", "") == B
        assert score_of(crt, crq, sc, len(A), len(B)) == cs
