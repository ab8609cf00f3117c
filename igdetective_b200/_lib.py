"""ctypes binding of libigd_b200.so -- the stub a reference maintainer would add.

The product path has no fallback: if the CUDA library is missing or no B200 is visible,
loading / Engine creation raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IGD_B200_LIB") or os.path.join(_HERE, "libigd_b200.so")

IGD_OK, IGD_ERR_CUDA, IGD_ERR_ARG, IGD_ERR_STATE = 0, -1, -2, -3
IGD_ERR_TOO_LONG, IGD_ERR_UNSUPPORTED, IGD_ERR_CAPACITY, IGD_ERR_INTERNAL = -4, -5, -6, -7
MAX_TARGET_LEN, MAX_QUERY_LEN = 1023, 1023
SIG_V, SIG_D_LEFT, SIG_D_RIGHT, SIG_J = 0, 1, 2, 3

HIT_DTYPE = np.dtype([("hept", "<i8"), ("nona", "<i8"), ("contig", "<u4"), ("sig_type", "u1"),
                      ("strand", "u1"), ("spacer", "u1"), ("reserved", "u1")])
KMER_HIT_DTYPE = np.dtype([("pos", "<i8"), ("contig", "<u4"), ("flags", "u1"), ("reserved", "u1", (3,))])
# igd_vj_row / igd_d_row (include/igd_b200.h): rows of genes_V/J.tsv and genes_D.tsv in integer form
VJ_ROW_DTYPE = np.dtype([("hept", "<i8"), ("nona", "<i8"), ("frag_start", "<i8"), ("contig", "<u4"), ("frag_len", "<i4"),
                         ("best_gene", "<i4"), ("matches", "<i4"), ("alen", "<i4"), ("start", "<i4"), ("ient", "<i4"),
                         ("gene_type", "u1"), ("strand", "u1"), ("direction", "u1"), ("reserved", "u1")])
D_ROW_DTYPE = np.dtype([("dl_hept", "<i8"), ("dl_nona", "<i8"), ("dr_hept", "<i8"), ("dr_nona", "<i8"), ("contig", "<u4"),
                        ("strand", "u1"), ("reserved", "u1", (3,))])
assert VJ_ROW_DTYPE.itemsize == 56 and D_ROW_DTYPE.itemsize == 40


class Scoring(ctypes.Structure):
    """igd_scoring == the ten PairwiseAligner attributes of SetupAligner (extract_aligned_genes.py:72-83)."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "match", "mismatch", "target_internal_open", "target_internal_extend",
        "target_end_open", "target_end_extend", "query_internal_open", "query_internal_extend",
        "query_end_open", "query_end_extend")]


class DetectParams(ctypes.Structure):
    """igd_detect_params == GENE_LENGTH, PI_CUTOFF, MAXK_CUTOFF of py/IGDetective.py:29-32 (index 0 = V, 1 = J)."""
    _fields_ = [("gene_length", ctypes.c_int32 * 2), ("d_gene_length", ctypes.c_int32),
                ("pi_strict", ctypes.c_double * 2), ("pi_relax", ctypes.c_double * 2), ("maxk_cutoff", ctypes.c_int32 * 2)]


class Timing(ctypes.Structure):
    _fields_ = [("pack_ms", ctypes.c_float), ("scan_ms", ctypes.c_float), ("align_ms", ctypes.c_float),
                ("scan_launches", ctypes.c_uint32), ("align_launches", ctypes.c_uint32),
                ("scan_bases", ctypes.c_uint64), ("align_cells", ctypes.c_uint64),
                ("total_launches", ctypes.c_uint64),
                ("dp_ms", ctypes.c_float), ("lcs_ms", ctypes.c_float), ("k5_ms", ctypes.c_float), ("reserved_f", ctypes.c_float),
                ("lcs_cells", ctypes.c_uint64), ("k5_cells", ctypes.c_uint64)]


class IgdError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("igd_b200 error %d: %s" % (code, message))
        self.code = code


EXPORTS = ("igd_last_error", "igd_device_count", "igd_create", "igd_destroy", "igd_sync", "igd_get_timing",
           "igd_set_motifs", "igd_center_bitmap", "igd_load_contigs", "igd_scan", "igd_fetch_hits", "igd_scan_kmers",
           "igd_fetch_kmer_hits", "igd_align_batch", "igd_fasta_open", "igd_fasta_close", "igd_fasta_count",
           "igd_fasta_info", "igd_fasta_read", "igd_load_fasta", "igd_align_dense", "igd_score_fragments", "igd_score_windows",
           "igd_detect", "igd_fetch_vj_rows", "igd_fetch_d_rows", "igd_fasta_open_range", "igd_fasta_range",
           "igd_fasta_header_byte", "igd_match_bounds", "igd_load_contigs_async", "igd_row_text_size",
           "igd_fetch_row_text")

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `python -m igdetective_b200.build` "
                           "(igdetective_b200 has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64)
    lib.igd_last_error.restype = ctypes.c_char_p
    lib.igd_last_error.argtypes = []
    lib.igd_device_count.restype = ctypes.c_int
    lib.igd_create.restype = vp
    lib.igd_create.argtypes = [ctypes.c_int]
    lib.igd_destroy.restype = None
    lib.igd_destroy.argtypes = [vp]
    lib.igd_sync.argtypes = [vp]
    lib.igd_get_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    lib.igd_set_motifs.argtypes = [vp, vp, vp]
    lib.igd_center_bitmap.argtypes = [vp, vp]
    lib.igd_load_contigs.argtypes = [vp, vp, vp, i32]
    lib.igd_load_contigs_async.argtypes = [vp, vp, vp, i32]
    lib.igd_load_contigs_async.restype = ctypes.c_int
    lib.igd_scan.argtypes = [vp, vp, u64p]
    lib.igd_fetch_hits.argtypes = [vp, vp, ctypes.c_uint64]
    lib.igd_scan_kmers.argtypes = [vp, i32, u64p]
    lib.igd_fetch_kmer_hits.argtypes = [vp, vp, ctypes.c_uint64]
    lib.igd_align_batch.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, i64,
                                    ctypes.POINTER(Scoring), vp, vp, vp, vp, vp]
    lib.igd_align_dense.argtypes = [vp, vp, vp, i32, vp, vp, i32, ctypes.POINTER(Scoring), vp, vp, vp]
    lib.igd_align_dense.restype = ctypes.c_int
    lib.igd_score_fragments.argtypes = [vp, vp, vp, i32, vp, vp, i32, ctypes.POINTER(Scoring), vp, vp, vp]
    lib.igd_score_fragments.restype = ctypes.c_int
    lib.igd_score_windows.argtypes = [vp, vp, vp, i32, vp, vp, i32, ctypes.POINTER(Scoring), vp, vp, vp, vp, vp]
    lib.igd_score_windows.restype = ctypes.c_int
    lib.igd_detect.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, ctypes.POINTER(Scoring), ctypes.POINTER(DetectParams),
                               u64p, u64p, u64p]
    lib.igd_detect.restype = ctypes.c_int
    lib.igd_fetch_vj_rows.argtypes = [vp, vp, ctypes.c_uint64]
    lib.igd_fetch_vj_rows.restype = ctypes.c_int
    lib.igd_fetch_d_rows.argtypes = [vp, vp, ctypes.c_uint64]
    lib.igd_fetch_d_rows.restype = ctypes.c_int
    lib.igd_row_text_size.argtypes = [vp, u64p, u64p]
    lib.igd_row_text_size.restype = ctypes.c_int
    lib.igd_fetch_row_text.argtypes = [vp, vp, ctypes.c_uint64, vp, ctypes.c_uint64]
    lib.igd_fetch_row_text.restype = ctypes.c_int
    lib.igd_match_bounds.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp]
    lib.igd_match_bounds.restype = ctypes.c_int
    lib.igd_fasta_open.restype = vp
    lib.igd_fasta_open.argtypes = [ctypes.c_char_p]
    lib.igd_fasta_open_range.restype = vp
    lib.igd_fasta_open_range.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    lib.igd_fasta_range.argtypes = [vp, u64p, u64p, u64p, u64p]
    lib.igd_fasta_range.restype = ctypes.c_int
    lib.igd_fasta_header_byte.argtypes = [vp, i32, u64p]
    lib.igd_fasta_header_byte.restype = ctypes.c_int
    lib.igd_fasta_close.restype = None
    lib.igd_fasta_close.argtypes = [vp]
    lib.igd_fasta_count.restype = i32
    lib.igd_fasta_count.argtypes = [vp]
    lib.igd_fasta_info.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_char_p), u64p]
    lib.igd_fasta_read.argtypes = [vp, i32, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.igd_load_fasta.argtypes = [vp, vp]
    for name in ("igd_fasta_info", "igd_fasta_read", "igd_load_fasta", "igd_align_dense"):
        getattr(lib, name).restype = ctypes.c_int
    for name in ("igd_sync", "igd_get_timing", "igd_set_motifs", "igd_center_bitmap", "igd_load_contigs", "igd_scan",
                 "igd_fetch_hits", "igd_scan_kmers", "igd_fetch_kmer_hits", "igd_align_batch"):
        getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != IGD_OK:
        raise IgdError(rc, load().igd_last_error().decode("utf-8", "replace"))
