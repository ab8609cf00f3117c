"""Bio.Align stand-in (test infrastructure): PairwiseAligner in global mode over
oracle/gotoh_ref.c.  align(target, query) -> sequence of alignments; [0] is the
first optimal path; alignment[0] / alignment[1] are the gapped target / query rows."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = ctypes.CDLL(os.path.join(_HERE, "..", "..", "libigd_oracle.so"))


class _Scoring(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "match", "mismatch",
        "t_int_open", "t_int_ext", "t_left_open", "t_left_ext", "t_right_open", "t_right_ext",
        "q_int_open", "q_int_ext", "q_left_open", "q_left_ext", "q_right_open", "q_right_ext")]


_LIB.igo_gotoh_first.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                 ctypes.POINTER(_Scoring), ctypes.POINTER(ctypes.c_double),
                                 ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
_LIB.igo_gotoh_first.restype = ctypes.c_int


class Alignment:
    def __init__(self, target_row, query_row, score):
        self._rows = (target_row, query_row)
        self.score = score

    def __getitem__(self, i):
        return self._rows[i]


class PairwiseAlignments:
    def __init__(self, alignment):
        self._a = alignment
        self.score = alignment.score

    def __len__(self):
        return 1          # the reference only tests len(...) == 0

    def __getitem__(self, i):
        if i != 0:
            raise IndexError("shim yields the first optimal alignment only")
        return self._a

    def __iter__(self):
        yield self._a


class PairwiseAligner:
    # BioPython defaults: match 1, mismatch 0, all gap scores 0
    _GAPS = ("target_internal_open_gap_score", "target_internal_extend_gap_score",
             "target_left_open_gap_score", "target_left_extend_gap_score",
             "target_right_open_gap_score", "target_right_extend_gap_score",
             "query_internal_open_gap_score", "query_internal_extend_gap_score",
             "query_left_open_gap_score", "query_left_extend_gap_score",
             "query_right_open_gap_score", "query_right_extend_gap_score")

    def __init__(self):
        object.__setattr__(self, "mode", "global")
        object.__setattr__(self, "match_score", 1.0)
        object.__setattr__(self, "mismatch_score", 0.0)
        for g in self._GAPS:
            object.__setattr__(self, g, 0.0)

    def __setattr__(self, name, value):
        if name in ("target_end_open_gap_score", "target_end_extend_gap_score",
                    "query_end_open_gap_score", "query_end_extend_gap_score"):
            side, _, kind = name.partition("_end_")
            for end in ("left", "right"):
                object.__setattr__(self, "%s_%s_%s" % (side, end, kind), float(value))
            return
        if name == "mode":
            if value != "global":
                raise ValueError("shim supports global mode only")
            object.__setattr__(self, name, value)
            return
        if name in ("match_score", "mismatch_score") or name in self._GAPS:
            object.__setattr__(self, name, float(value))
            return
        raise AttributeError("shim PairwiseAligner has no attribute %r" % name)

    def _scoring(self):
        return _Scoring(self.match_score, self.mismatch_score,
                        self.target_internal_open_gap_score, self.target_internal_extend_gap_score,
                        self.target_left_open_gap_score, self.target_left_extend_gap_score,
                        self.target_right_open_gap_score, self.target_right_extend_gap_score,
                        self.query_internal_open_gap_score, self.query_internal_extend_gap_score,
                        self.query_left_open_gap_score, self.query_left_extend_gap_score,
                        self.query_right_open_gap_score, self.query_right_extend_gap_score)

    def align(self, target, query):
        a = str(target).encode("latin-1")
        b = str(query).encode("latin-1")
        n = len(a) + len(b) + 1
        rt = ctypes.create_string_buffer(n)
        rq = ctypes.create_string_buffer(n)
        score = ctypes.c_double()
        ncols = ctypes.c_int()
        sc = self._scoring()
        if _LIB.igo_gotoh_first(a, len(a), b, len(b), ctypes.byref(sc), ctypes.byref(score),
                                rt, rq, ctypes.byref(ncols)):
            raise MemoryError("oracle aligner")
        return PairwiseAlignments(Alignment(rt.raw[:ncols.value].decode("latin-1"),
                                            rq.raw[:ncols.value].decode("latin-1"), score.value))
