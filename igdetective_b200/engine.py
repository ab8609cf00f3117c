"""Engine: one igd_handle (one GPU, one stream).  numpy arrays in, numpy arrays out."""
import ctypes

import numpy as np

from . import _lib

SIGNAL_INDEX = {"V": _lib.SIG_V, "D_left": _lib.SIG_D_LEFT, "D_right": _lib.SIG_D_RIGHT, "J": _lib.SIG_J}
SIGNAL_NAMES = ["V", "D_left", "D_right", "J"]

# SetupAligner, py/extract_aligned_genes.py:72-83
REFERENCE_SCORING = dict(match=2, mismatch=0,
                         target_internal_open=-2, target_internal_extend=-1,
                         target_end_open=-2, target_end_extend=-1,
                         query_internal_open=-2, query_internal_extend=-1,
                         query_end_open=0, query_end_extend=0)

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def motif_flag_tables(motifs):
    """VALID_MOTIFS dict (py/IGDetective.py:54-58) -> (flags7[16384], flags9[262144]) for igd_set_motifs.
    Index = base-4 value of the k-mer, first letter most significant; bit t = signal type t."""
    tabs = {7: np.zeros(4 ** 7, np.uint8), 9: np.zeros(4 ** 9, np.uint8)}
    for name, t in SIGNAL_INDEX.items():
        for k in (7, 9):
            for kmer in motifs[name][str(k)]:
                if len(kmer) != k or any(c not in _CODE for c in kmer):
                    continue        # cannot match: windows are upper-cased ACGT or rejected (:141-142)
                v = 0
                for c in kmer:
                    v = v * 4 + _CODE[c]
                tabs[k][v] |= 1 << t
    return tabs[7], tabs[9]


def concat(seqs):
    """list[str | bytes] -> (uint8 buffer, int32 offsets[n+1])"""
    bs = [s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1") for s in seqs]
    off = np.zeros(len(bs) + 1, np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    if off[-1] >= 2 ** 31:
        raise ValueError("sequence table exceeds 2 GiB")
    buf = np.frombuffer(b"".join(bs), np.uint8) if bs else np.zeros(0, np.uint8)
    return buf, off.astype(np.int32)


class Engine:
    def __init__(self, device=0, motifs=None):
        self._lib = _lib.load()
        self._h = self._lib.igd_create(int(device))
        if not self._h:
            raise _lib.IgdError(_lib.IGD_ERR_CUDA, self._lib.igd_last_error().decode("utf-8", "replace"))
        self.device = int(device)
        self.contig_ids = []
        self.contig_lens = np.zeros(0, np.int64)
        self.h2d_bytes = 0          # bytes handed to the C ABI as inputs / read back as results
        self.d2h_bytes = 0
        self.align_ms_total = 0.0   # running sums over align() calls (device time of the DP kernels, cell updates)
        self.align_cells_total = 0
        if motifs is not None:
            self.set_motifs(motifs)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self, "_lib", None) is not None:
            self._lib.igd_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown: the library may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- scan -------------------------------------------------------------------------------
    def set_motifs(self, motifs):
        f7, f9 = motifs if isinstance(motifs, tuple) else motif_flag_tables(motifs)
        f7 = np.ascontiguousarray(f7, np.uint8)
        f9 = np.ascontiguousarray(f9, np.uint8)
        assert f7.size == 16384 and f9.size == 262144
        _lib.check(self._lib.igd_set_motifs(self._h, f7.ctypes.data, f9.ctypes.data))

    def load_contigs(self, seqs, ids=None, wait=True):
        """seqs: dict {id: str} (parent_seq of the reference), list of str/bytes, or a
        (uint8 buffer, uint64 offsets) pair already concatenated (e.g. pinned host memory).
        wait=False (igd_load_contigs_async): return once the letters have crossed the bus; the pack kernel stays queued
        on this engine's stream, in front of whatever this engine is asked next."""
        if isinstance(seqs, dict):
            ids, seqs = list(seqs.keys()), list(seqs.values())
        if isinstance(seqs, tuple):
            buf, off = seqs
            off = np.ascontiguousarray(off, np.uint64)
            ptr = buf if isinstance(buf, int) else np.ascontiguousarray(buf, np.uint8).ctypes.data
        else:
            bs = [s if isinstance(s, (bytes, bytearray)) else str(s).encode("latin-1") for s in seqs]
            off = np.zeros(len(bs) + 1, np.uint64)
            if bs:
                off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
            buf = np.frombuffer(b"".join(bs), np.uint8) if bs else np.zeros(1, np.uint8)
            ptr = buf.ctypes.data
        n = len(off) - 1
        self.contig_ids = list(ids) if ids is not None else list(range(n))
        self.contig_lens = np.diff(off.astype(np.int64))
        _lib.check((self._lib.igd_load_contigs if wait else self._lib.igd_load_contigs_async)(self._h, ptr, off.ctypes.data, n))
        self.h2d_bytes += int(off[-1] - off[0]) + off.nbytes
        return n

    def load_fasta(self, fasta):
        """fasta: a fasta.NativeFasta.  Sends its contigs to the GPU straight from the library's buffer."""
        _lib.check(self._lib.igd_load_fasta(self._h, fasta._f))
        self.contig_ids = list(fasta.ids)
        self.contig_lens = np.asarray(fasta.lengths, np.int64)
        self.h2d_bytes += int(self.contig_lens.sum()) + 8 * (len(fasta.ids) + 1)
        return len(fasta.ids)

    def scan_rss(self, spacers):
        """spacers: {signal name: length} or 4 ints in (V, D_left, D_right, J) order.
        Returns a HIT_DTYPE array sorted by (contig, sig_type, strand, scan-order index)."""
        if isinstance(spacers, dict):
            spacers = [spacers[n] for n in SIGNAL_NAMES]
        sp = np.asarray(spacers, np.int32)
        n = ctypes.c_uint64()
        _lib.check(self._lib.igd_scan(self._h, sp.ctypes.data, ctypes.byref(n)))
        out = np.zeros(n.value, _lib.HIT_DTYPE)
        _lib.check(self._lib.igd_fetch_hits(self._h, out.ctypes.data, n.value))
        self.d2h_bytes += 8 + 8 * n.value       # count + one packed 64-bit key per hit cross the bus
        return out

    def scan_rss_count(self, spacers):
        """Scan only (hits stay on the device); returns the hit count."""
        sp = np.asarray([spacers[n] for n in SIGNAL_NAMES] if isinstance(spacers, dict) else spacers, np.int32)
        n = ctypes.c_uint64()
        _lib.check(self._lib.igd_scan(self._h, sp.ctypes.data, ctypes.byref(n)))
        return n.value

    def scan_kmers(self, k):
        n = ctypes.c_uint64()
        _lib.check(self._lib.igd_scan_kmers(self._h, int(k), ctypes.byref(n)))
        out = np.zeros(n.value, _lib.KMER_HIT_DTYPE)
        _lib.check(self._lib.igd_fetch_kmer_hits(self._h, out.ctypes.data, n.value))
        self.d2h_bytes += 8 + 8 * n.value
        return out

    # ---- alignment --------------------------------------------------------------------------
    def align(self, targets, queries, pair_t, pair_q, scoring=None, detail=False):
        """targets / queries: list[str|bytes] or (uint8 buffer, int32 offsets).  Returns a dict of int32
        arrays: score, matches (TRUE count), alen (= end - start) and, with detail=True, also start, end
        and ient (QuerySeq = target[start:ient]); detail costs the slower two-pass kernel."""
        tb, to = targets if isinstance(targets, tuple) else concat(targets)
        qb, qo = queries if isinstance(queries, tuple) else concat(queries)
        tb = np.ascontiguousarray(tb, np.uint8)
        qb = np.ascontiguousarray(qb, np.uint8)
        to = np.ascontiguousarray(to, np.int32)
        qo = np.ascontiguousarray(qo, np.int32)
        pt = np.ascontiguousarray(pair_t, np.int32)
        pq = np.ascontiguousarray(pair_q, np.int32)
        n = len(pt)
        assert len(pq) == n
        sc = _lib.Scoring(**(scoring or REFERENCE_SCORING))
        res = {k: np.zeros(n, np.int32) for k in (("score", "matches", "alen", "start", "ient") if detail
                                                   else ("score", "matches", "alen"))}
        _lib.check(self._lib.igd_align_batch(
            self._h, tb.ctypes.data if tb.size else None, to.ctypes.data, len(to) - 1,
            qb.ctypes.data if qb.size else None, qo.ctypes.data, len(qo) - 1,
            pt.ctypes.data, pq.ctypes.data, n, ctypes.byref(sc),
            res["score"].ctypes.data, res["matches"].ctypes.data, res["alen"].ctypes.data,
            res["start"].ctypes.data if detail else None, res["ient"].ctypes.data if detail else None))
        if detail:
            res["end"] = res["start"] + res["alen"]
        t = self.timing()
        self.align_ms_total += t["align_ms"]
        self.align_cells_total += t["align_cells"]
        self.h2d_bytes += tb.nbytes + to.nbytes + qb.nbytes + qo.nbytes + 3 * pt.nbytes
        self.d2h_bytes += (4 if detail else 2) * 4 * n
        return res

    def align_dense(self, targets, queries, scoring=None):
        """Every target against every query: dict of int32 arrays [T, Q] (score, matches, alen)."""
        tb, to = targets if isinstance(targets, tuple) else concat(targets)
        qb, qo = queries if isinstance(queries, tuple) else concat(queries)
        tb = np.ascontiguousarray(tb, np.uint8)
        qb = np.ascontiguousarray(qb, np.uint8)
        to = np.ascontiguousarray(to, np.int32)
        qo = np.ascontiguousarray(qo, np.int32)
        T, Q = len(to) - 1, len(qo) - 1
        sc = _lib.Scoring(**(scoring or REFERENCE_SCORING))
        res = {k: np.zeros((T, Q), np.int32) for k in ("score", "matches", "alen")}
        if T and Q:
            _lib.check(self._lib.igd_align_dense(
                self._h, tb.ctypes.data if tb.size else None, to.ctypes.data, T,
                qb.ctypes.data if qb.size else None, qo.ctypes.data, Q, ctypes.byref(sc),
                res["score"].ctypes.data, res["matches"].ctypes.data, res["alen"].ctypes.data))
            t = self.timing()
            self.align_ms_total += t["align_ms"]
            self.align_cells_total += t["align_cells"]
        self.h2d_bytes += tb.nbytes + to.nbytes + qb.nbytes + qo.nbytes
        self.d2h_bytes += 2 * 4 * T * Q
        return res

    def score_fragments(self, fragments, genes, scoring=None, direction=False):
        """align_fragment_to_genes on the device (igd_score_fragments): float64 [F, G] matrices
        (pi, alen) of the better orientation of every fragment against every gene; with
        direction=True also a [F, G] uint8 matrix of '+' / '-'."""
        fb, fo = fragments if isinstance(fragments, tuple) else concat(fragments)
        gb, go = genes if isinstance(genes, tuple) else concat(genes)
        fb = np.ascontiguousarray(fb, np.uint8)
        gb = np.ascontiguousarray(gb, np.uint8)
        fo = np.ascontiguousarray(fo, np.int32)
        go = np.ascontiguousarray(go, np.int32)
        F, G = len(fo) - 1, len(go) - 1
        sc = _lib.Scoring(**(scoring or REFERENCE_SCORING))
        pi = np.zeros((F, G), np.float64)
        alen = np.zeros((F, G), np.float64)
        dirs = np.zeros((F, G), np.uint8) if direction else None
        if F and G:
            _lib.check(self._lib.igd_score_fragments(
                self._h, fb.ctypes.data if fb.size else None, fo.ctypes.data, F,
                gb.ctypes.data if gb.size else None, go.ctypes.data, G, ctypes.byref(sc),
                pi.ctypes.data, alen.ctypes.data, dirs.ctypes.data if direction else None))
            t = self.timing()
            self.align_ms_total += t["align_ms"]
            self.align_cells_total += t["align_cells"]
        self.h2d_bytes += 2 * fb.nbytes + 2 * fo.nbytes + gb.nbytes + go.nbytes
        self.d2h_bytes += 4 * F * G
        return (pi, alen, dirs) if direction else (pi, alen)

    def score_windows(self, windows, genes, scoring=None):
        """extract_aligned_genes.ComputeAlignment for many raw-case windows (igd_score_windows): dict of
        per-window arrays best (orientation * G + gene, -1 = none), pi (float64), start, end, ient."""
        wb, wo = windows if isinstance(windows, tuple) else concat(windows)
        gb, go = genes if isinstance(genes, tuple) else concat(genes)
        wb = np.ascontiguousarray(wb, np.uint8)
        gb = np.ascontiguousarray(gb, np.uint8)
        wo = np.ascontiguousarray(wo, np.int32)
        go = np.ascontiguousarray(go, np.int32)
        W, G = len(wo) - 1, len(go) - 1
        sc = _lib.Scoring(**(scoring or REFERENCE_SCORING))
        res = {k: np.zeros(W, np.int32) for k in ("best", "start", "end", "ient")}
        res["best"][:] = -1
        res["pi"] = np.zeros(W, np.float64)
        if W and G:
            _lib.check(self._lib.igd_score_windows(
                self._h, wb.ctypes.data if wb.size else None, wo.ctypes.data, W,
                gb.ctypes.data if gb.size else None, go.ctypes.data, G, ctypes.byref(sc),
                res["best"].ctypes.data, res["pi"].ctypes.data, res["start"].ctypes.data,
                res["end"].ctypes.data, res["ient"].ctypes.data))
            t = self.timing()
            self.align_ms_total += t["align_ms"]
            self.align_cells_total += t["align_cells"]
        self.h2d_bytes += 2 * wb.nbytes + 2 * wo.nbytes + gb.nbytes + go.nbytes
        self.d2h_bytes += 16 * W + 16 * W
        return res

    def detect(self, spacers, v_genes, j_genes, scoring=None, gene_length=(350, 70), d_gene_length=150,
               pi_strict=(70, 70), pi_relax=(60, 65), maxk_cutoff=(15, 11)):
        """The whole de-novo stage on the loaded contigs in one call (igd_detect): scan, D pairing, fragment gather on
        the device, scoring of every V / J candidate against every gene in both orientations, argmax, threshold and the
        winners' detail pass.  v_genes / j_genes: list[str] (or (buffer, offsets)).  Returns {"hits": HIT_DTYPE array,
        "vj": VJ_ROW_DTYPE array (V rows, then J rows), "d": D_ROW_DTYPE array}, rows in the reference's order for
        ascending RSS lists."""
        sp = np.asarray([spacers[n] for n in SIGNAL_NAMES] if isinstance(spacers, dict) else spacers, np.int32)
        vb, vo = v_genes if isinstance(v_genes, tuple) else concat(v_genes)
        jb, jo = j_genes if isinstance(j_genes, tuple) else concat(j_genes)
        vb, jb = np.ascontiguousarray(vb, np.uint8), np.ascontiguousarray(jb, np.uint8)
        vo, jo = np.ascontiguousarray(vo, np.int32), np.ascontiguousarray(jo, np.int32)
        sc = _lib.Scoring(**(scoring or REFERENCE_SCORING))
        prm = _lib.DetectParams((ctypes.c_int32 * 2)(*gene_length), int(d_gene_length), (ctypes.c_double * 2)(*pi_strict),
                                (ctypes.c_double * 2)(*pi_relax), (ctypes.c_int32 * 2)(*maxk_cutoff))
        nh, nvj, nd = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _lib.check(self._lib.igd_detect(self._h, sp.ctypes.data, vb.ctypes.data if vb.size else None, vo.ctypes.data, len(vo) - 1,
                                        jb.ctypes.data if jb.size else None, jo.ctypes.data, len(jo) - 1,
                                        ctypes.byref(sc), ctypes.byref(prm), ctypes.byref(nh), ctypes.byref(nvj), ctypes.byref(nd)))
        hits = np.zeros(nh.value, _lib.HIT_DTYPE)
        _lib.check(self._lib.igd_fetch_hits(self._h, hits.ctypes.data, nh.value))
        vj = np.zeros(nvj.value, _lib.VJ_ROW_DTYPE)
        _lib.check(self._lib.igd_fetch_vj_rows(self._h, vj.ctypes.data, nvj.value))
        d = np.zeros(nd.value, _lib.D_ROW_DTYPE)
        _lib.check(self._lib.igd_fetch_d_rows(self._h, d.ctypes.data, nd.value))
        nb, ns = ctypes.c_uint64(), ctypes.c_uint64()
        _lib.check(self._lib.igd_row_text_size(self._h, ctypes.byref(nb), ctypes.byref(ns)))
        text = np.empty(nb.value, np.uint8)
        text_off = np.empty(ns.value + 1, np.uint64)
        _lib.check(self._lib.igd_fetch_row_text(self._h, text.ctypes.data if nb.value else None, nb.value, text_off.ctypes.data, ns.value + 1))
        t = self.timing()
        self.align_ms_total += t["align_ms"]
        self.align_cells_total += t["align_cells"]
        # over the bus: gene tables in; one 64-bit key per hit, 16 bytes per scored fragment (<= hits) and the
        # winners' words out -- the device never sends letters back
        n_cand = int(np.count_nonzero((hits["sig_type"] == _lib.SIG_V) | (hits["sig_type"] == _lib.SIG_J)))
        self.h2d_bytes += vb.nbytes + vo.nbytes + jb.nbytes + jo.nbytes + 24 * n_cand + 12 * int(nvj.value)
        self.d2h_bytes += 8 + 8 * int(nh.value) + 16 * n_cand + 8 * int(nvj.value) + int(nb.value)
        return {"hits": hits, "vj": vj, "d": d, "text": text.tobytes().decode("latin-1"), "text_off": text_off.tolist()}

    def match_bounds(self, fragments, genes):
        """Upper bounds on the match count / PI of fragments[u] (and its reverse complement) against genes[g]
        (igd_match_bounds): (lcs [U, 2, G] uint16, y [U, 2, G] uint16), index 1 = reverse complement."""
        from .fasta import reverse_complement
        rows = []
        for f in fragments:
            f = str(f).upper()
            rows += [f, reverse_complement(f)]
        tb, to = concat(rows)
        qb, qo = concat(genes)
        tb, qb = np.ascontiguousarray(tb, np.uint8), np.ascontiguousarray(qb, np.uint8)
        to, qo = np.ascontiguousarray(to, np.int32), np.ascontiguousarray(qo, np.int32)
        U, G = len(fragments), len(genes)
        lcs = np.zeros((U, 2, G), np.uint16)
        y = np.zeros((U, G), np.uint32)
        _lib.check(self._lib.igd_match_bounds(self._h, tb.ctypes.data if tb.size else None, to.ctypes.data, U,
                                              qb.ctypes.data if qb.size else None, qo.ctypes.data, G,
                                              lcs.ctypes.data, y.ctypes.data))
        return lcs, np.stack([(y & 0xffff).astype(np.uint16), (y >> 16).astype(np.uint16)], axis=1)

    def sync(self):
        _lib.check(self._lib.igd_sync(self._h))

    def timing(self):
        t = _lib.Timing()
        _lib.check(self._lib.igd_get_timing(self._h, ctypes.byref(t)))
        return {f[0]: getattr(t, f[0]) for f in _lib.Timing._fields_}
