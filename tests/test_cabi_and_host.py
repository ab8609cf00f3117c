"""CPU-only checks: the C ABI library loads and exports what include/igd_b200.h declares, fails
loudly without a GPU, and the host-side glue (order replay, D pairing, tables, FASTA) is right."""
import os
import re

import numpy as np
import pytest

from conftest import DATAFILES, GOLDEN, ROOT, golden_input
from oracle import igd_oracle as O


def test_library_exports_every_declared_symbol():
    from igdetective_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "igd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(igd_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from igdetective_b200._lib import IgdError
    from igdetective_b200.engine import Engine
    with pytest.raises(IgdError):
        Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "igdetective_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.replace("igd_oracle_unused", ""), fn


def test_motif_flag_tables(motifs):
    from igdetective_b200.engine import SIGNAL_INDEX, motif_flag_tables
    f7, f9 = motif_flag_tables(motifs)
    for name, t in SIGNAL_INDEX.items():
        assert int(((f7 >> t) & 1).sum()) == len(motifs[name]["7"])
        assert int(((f9 >> t) & 1).sum()) == len(motifs[name]["9"])
    v = 0
    for c in "CACAGTG":
        v = v * 4 + "ACGT".index(c)
    assert f7[v] & 1


def test_datafiles_loader_matches_oracle(motifs):
    from igdetective_b200 import datafiles
    assert datafiles.load_motifs(DATAFILES) == motifs
    g = datafiles.load_canonical_genes("IGH", "J", "combined_reference_genes", DATAFILES)
    assert g == O.fasta_dict(os.path.join(DATAFILES, "combined_reference_genes", "IGHJ.fa.gz"), upper=True)
    assert all(s == s.upper() for s in g.values())


def test_reference_set_order_replay():
    from igdetective_b200.order import reference_set_order
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 263, 5000):
        asc = sorted(set(rng.integers(0, 3_000_000, n).tolist()))
        assert reference_set_order(asc) == O.reference_set_order(asc)
        assert sorted(reference_set_order(asc)) == asc


def test_combine_d_rss_equals_reference_loops():
    from igdetective_b200.rss import combine_D_RSS
    rng = np.random.default_rng(1)
    seqs = {"a": "x", "b": "y", "c": "z"}
    for strand in "+-":
        dl = {c: [(int(p), int(p) - 21) for p in rng.integers(0, 3000, 60)] for c in seqs}
        dr = {c: [(int(p), int(p) + 19) for p in rng.integers(0, 3000, 50)] for c in seqs}
        dl["c"], dr["b"] = [], []
        assert combine_D_RSS(dl, dr, seqs, strand) == O.combine_D_RSS(dl, dr, seqs, strand)


def test_fragment_slicing_quirks():
    """negative slice start near the contig start gives an empty fragment, not a truncated one"""
    from igdetective_b200.rss import get_s_fragment_from_RSS
    s = "".join("ACGT"[i % 4] for i in range(1000))
    seqs = {"c": s}
    info = {"V": {"+": {"c": [(100, 130), (400, 430)]}, "-": {"c": [(900, 870)]}},
            "J": {"+": {"c": [(990, 950)]}, "-": {"c": [(30, 60), (500, 530)]}}}
    for g in "VJ":
        for sd in "+-":
            assert get_s_fragment_from_RSS(g, sd, info, seqs) == O.s_fragments(g, sd, info, seqs)
    assert get_s_fragment_from_RSS("V", "+", info, seqs)["c"][0] == ""
    assert len(get_s_fragment_from_RSS("J", "+", info, seqs)["c"][0]) == 3


def test_fasta_reader_and_translate():
    from igdetective_b200.extract_aligned_genes import translate
    from igdetective_b200.fasta import fasta_dict, read_fasta, reverse_complement
    p = golden_input("cow")
    assert read_fasta(p) == O.read_fasta(p)
    assert list(fasta_dict(p)) == ["KT723008.1|Bos"]
    assert reverse_complement("ACGTNacgtRy") == O.revcomp("ACGTNacgtRy") == "rYacgtNACGT"
    for s in ("ATGGCCTAA", "atggcNtga", "ATGNNNAC", "GCNTAR", ""):
        assert translate(s) == O.translate(s)
    assert translate("ATGGCCTAA") == "MA*" and translate("GCN") == "A" and translate("TAR") == "*"


def test_native_fasta_reader_equals_oracle(tmp_path):
    """igd_fasta_open (C++) vs the oracle's restatement of SeqIO.parse on awkward files"""
    import gzip

    from igdetective_b200._lib import IgdError
    from igdetective_b200.fasta import NativeFasta
    cases = {
        "plain": ">a desc here\nACGT\nacgtn\n>b\nTTTT\n",
        "crlf": ">a x\r\nACGT\r\nAC GT\r\n\r\n>b\r\nNNNN\r\n",
        "no_final_newline": ">only\nACGTACGT",
        "blank_lines_and_tabs": "\n\n>a\tb c\n\nAC\tGT\n  \nGG  \n>empty\n>c\nA\n",
        "text_before_first_header": "junk line\nACGT\n>a\nCC\n",
        "duplicate_ids": ">a\nAAAA\n>b\nCC\n>a\nGGGGGG\n>c\nT\n>b\nTTT\n",
        "empty_id": ">\nACGT\n> spaced id\nGG\n",
        "empty_file": "",
    }
    for name, text in cases.items():
        p = tmp_path / (name + ".fa")
        p.write_text(text, newline="")
        want = O.fasta_dict(str(p))
        nf = NativeFasta(str(p))
        assert list(nf) == list(want), name
        for k in want:
            assert str(nf[k]) == want[k], (name, k)
            s = want[k]
            for a, b in ((0, 3), (-3, None), (2, 100), (-350, 1), (1, -1), (5, 2)):
                assert nf[k][a:b] == s[a:b], (name, k, a, b)
        nf.close()
    # a real locus, byte for byte
    raw = tmp_path / "cow.fa"
    raw.write_bytes(gzip.open(golden_input("cow")).read())
    nf = NativeFasta(str(raw))
    want = O.fasta_dict(golden_input("cow"))
    assert list(nf) == list(want) and str(nf[list(want)[0]]) == want[list(want)[0]]
    with pytest.raises(IgdError):
        NativeFasta(golden_input("cow"))          # gzip is refused loudly, not misparsed
    with pytest.raises(IgdError):
        NativeFasta(str(tmp_path / "missing.fa"))


def test_native_fasta_reader_multithreaded(tmp_path):
    """a file large enough to be cut into several pieces (one per parser thread): records of skewed
    lengths, CRLF, a duplicate id, text in front of the first header; same result as the oracle reader"""
    from igdetective_b200.fasta import NativeFasta
    rng = np.random.default_rng(3)
    p = tmp_path / "big.fa"
    with open(p, "wb") as f:
        f.write(b"preamble that is not part of any record\nACGTACGT\n")
        for i in range(60):
            L = int(rng.integers(1, 4_000_000))
            name = b"dup" if i in (7, 41) else b"rec%d" % i
            f.write(b">" + name + b" some description\r\n")
            seq = np.frombuffer(b"ACGTNacgt", np.uint8)[rng.integers(0, 9, L)]
            w = int(rng.integers(50, 90))
            pad = (-L) % w
            rows = np.concatenate([seq, np.full(pad, 32, np.uint8)]).reshape(-1, w)
            f.write(np.concatenate([rows, np.full((len(rows), 1), 13, np.uint8), np.full((len(rows), 1), 10, np.uint8)], axis=1).tobytes())
    assert os.path.getsize(p) > 3 * (32 << 20)
    want = O.fasta_dict(str(p))
    nf = NativeFasta(str(p))
    assert list(nf) == list(want)
    for k in want:
        assert len(nf[k]) == len(want[k]), k
        assert str(nf[k]) == want[k], k


def test_center_bitmap_is_conservative(motifs):
    """igd_center_bitmap (no GPU): a clear bit for the 9-mer at x+1 proves that none of the three candidate
    nonamer starts x, x+1, x+2 of a spacer window (py/IGDetective.py:162-168) holds a motif on either strand;
    and the filter stays selective (a few per cent of all 9-mers)."""
    from igdetective_b200 import _lib
    from igdetective_b200.engine import motif_flag_tables
    lib = _lib.load()
    _, f9 = motif_flag_tables(motifs)
    f9 = np.ascontiguousarray(f9)
    bm = np.zeros(8192, np.uint32)
    assert lib.igd_center_bitmap(f9.ctypes.data, bm.ctypes.data) == 0
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")

    def kidx(s):          # kernel index: (high bits of the codes << 9) | low bits, first base = bit 0
        lo = hi = 0
        for i, c in enumerate(s):
            v = "ACGT".index(c)
            lo |= (v & 1) << i
            hi |= (v >> 1) << i
        return (hi << 9) | lo
    comp = str.maketrans("ACGT", "TGCA")
    members = set()
    for t in motifs:
        for m in motifs[t]["9"]:
            members.add(m)
            members.add(m.translate(comp)[::-1])
    for m in members:
        assert bits[kidx(m)]
        for a in "ACGT":
            assert bits[kidx(m[1:] + a)] and bits[kidx(a + m[:-1])]
    # exhaustive the other way round on a random sample: bit set => one of the three windows can be a member
    rng = np.random.default_rng(5)
    pre = {m[:8] for m in members}
    suf = {m[1:] for m in members}
    for _ in range(20000):
        c = "".join("ACGT"[i] for i in rng.integers(0, 4, 9))
        want = c in members or c[:8] in suf or c[1:] in pre
        assert bool(bits[kidx(c)]) == want
    assert 0.01 < bits.mean() < 0.08


def test_sparse_host_glue_equals_reference_shapes():
    """whole-genome mode (order="canonical") shares one empty value between the contigs that have nothing; the
    contents are those of the reference-shaped dicts: same keys in the same order, same lists where there is
    anything, an empty sequence elsewhere"""
    from igdetective_b200.rss import combine_D_RSS, get_s_fragment_from_RSS
    rng = np.random.default_rng(7)
    ids = ["c%d" % i for i in range(40)]
    s = "".join("ACGT"[i] for i in rng.integers(0, 4, 3000))
    seqs = {c: s for c in ids}
    for strand in "+-":
        dl = {c: ([(int(p), int(p) - 21) for p in rng.integers(0, 3000, 30)] if i % 3 == 0 else []) for i, c in enumerate(ids)}
        dr = {c: ([(int(p), int(p) + 19) for p in rng.integers(0, 3000, 30)] if i % 2 == 0 else []) for i, c in enumerate(ids)}
        dense = combine_D_RSS(dl, dr, seqs, strand)
        sparse = combine_D_RSS(dl, dr, seqs, strand, sparse=True)
        assert dense == O.combine_D_RSS(dl, dr, seqs, strand)
        assert list(sparse) == list(dense)
        assert sum(len(v) for v in dense.values()) > 0
        for c in ids:
            assert list(sparse[c]) == dense[c]
            assert dense[c] or sparse[c] == ()
        info = {"V": {strand: {c: [(int(p), int(p) + 30) for p in rng.integers(400, 2500, 3)] if i % 4 == 0 else []
                               for i, c in enumerate(ids)}},
                "D": {strand: dense}}
        for g in ("V", "D"):
            fd = get_s_fragment_from_RSS(g, strand, info, seqs)
            fs = get_s_fragment_from_RSS(g, strand, info, seqs, sparse=True)
            assert fd == O.s_fragments(g, strand, info, seqs)
            assert list(fs) == list(fd) and all(list(fs[c]) == fd[c] for c in ids)


def test_combine_ig_genes_equals_the_reference_definition_and_golden(tmp_path):
    """iterative.CombineIGGenes (run_iterative_igdetective.py:130-151): the indexed substring filter against the
    reference's double loop on random sets with planted substrings (both code paths), and byte for byte against the
    <gene>_combined.fasta files the unmodified driver wrote (same PYTHONHASHSEED: the set order names the sequences)."""
    import subprocess
    import sys

    from igdetective_b200 import iterative
    rng = np.random.default_rng(21)

    def rand(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    for n, lo, hi in ((60, 1, 80), (2500, 16, 60)):
        base = [rand(int(rng.integers(lo, hi))) for _ in range(n)]
        extra = [s[int(rng.integers(0, len(s) // 2 + 1)):int(rng.integers(len(s) // 2 + 1, len(s) + 1))] for s in base[::7]]
        seqs = list(dict.fromkeys(base + extra + ([""] if lo == 1 else [])))
        want = [any(a != b and b.find(a) != -1 for b in seqs) for a in seqs]
        assert iterative.contained_in_another(seqs) == want
        assert sum(want) >= len(extra) // 2
    case = os.path.join(GOLDEN, "driver", "tri_locus")
    code = ("import sys; sys.path.insert(0, %r); from igdetective_b200.iterative import CombineIGGenes; "
            "CombineIGGenes(sys.argv[1], sys.argv[2], sys.argv[3])" % ROOT)
    for gene in ("IGHV", "IGKV", "IGLV", "TRBV"):
        out = str(tmp_path / (gene + ".fasta"))
        subprocess.run([sys.executable, "-c", code, os.path.join(case, "iterative_search", gene + "_iter0", "genes.fasta"),
                        os.path.join(case, "denovo_search", "predicted_genes_" + gene[:3], "genes_V.tsv"), out],
                       check=True, env=dict(os.environ, PYTHONHASHSEED="0"))
        assert open(out).read() == open(os.path.join(case, "iterative_search", gene + "_combined.fasta")).read(), gene


def test_byte_range_reader_partitions_the_file_and_rebuilds_every_contig(tmp_path):
    """igd_fasta_open_range + shard.range_summary / contig_table / range_pieces (the per-rank parse of a sharded run):
    for 1..60 ranks the parts the ranks OWN, put back in contig coordinates, are exactly the records of the whole-file
    reader -- with wrapped and unwrapped lines, CRLF, blank lines, an empty record, letters in front of the first
    header, a header in the file's last line, ranges without any header and ranges smaller than a line."""
    from igdetective_b200 import shard
    from igdetective_b200.fasta import NativeFasta
    rng = np.random.default_rng(8)

    def rand(n):
        return "".join("ACGTNacgt"[i] for i in rng.integers(0, 9, n))
    recs = [("c%d" % i, rand(int(n))) for i, n in enumerate([5000, 0, 1, 61, 12000, 3, 700, 2500, 60, 9000])]
    text = "ACGTAC\nGT\n"                                        # dropped: in front of the first header
    for i, (rid, s) in enumerate(recs):
        text += ">%s some description\n" % rid
        w = (60, 1000000, 7)[i % 3]
        eol = "\r\n" if i % 4 == 1 else "\n"
        for k in range(0, len(s), w):
            text += s[k:k + w] + eol
        if i % 5 == 2:
            text += "\n"
    text += ">last_without_letters"
    recs.append(("last_without_letters", ""))
    path = str(tmp_path / "awkward.fa")
    open(path, "w", newline="").write(text)
    whole = NativeFasta(path)
    assert [(r, str(whole[r])) for r in whole.ids] == recs
    size = os.path.getsize(path)
    for world in (1, 2, 3, 7, 16, 60):
        fas = [NativeFasta(path, byte_range=(size * r // world, size * (r + 1) // world), margin=2000) for r in range(world)]
        assert [f.byte_range[0] for f in fas[1:]] == [f.byte_range[1] for f in fas[:-1]]            # the ranges tile the file
        ids, lens, starts = shard.contig_table([shard.range_summary(f) for f in fas])
        assert ids == [r for r, _ in recs] and lens == [len(s) for _, s in recs], world
        built = {r: ["?"] * len(s) for r, s in recs}
        for r, f in enumerate(fas):
            for k, (ci, a, b, p0, p1) in enumerate(shard.range_pieces(f, ids, lens, starts[r], halo=16)):
                view = f.view(k)
                assert p1 - p0 == len(view)
                if b > a:
                    piece = view[a - p0:b - p0]
                    assert built[ids[ci]][a:b] == ["?"] * (b - a)                                      # owned once
                    built[ids[ci]][a:b] = list(piece)
                    assert str(view) == recs[ci][1][p0:p1]                                             # context = the contig's own letters
        assert {r: "".join(v) for r, v in built.items()} == dict(recs), world
    # a repeated id cannot be sharded by byte ranges
    dup = str(tmp_path / "dup.fa")
    open(dup, "w").write(">a\nACGT\n>b\nAC\n>a\nGGGG\n")
    f = NativeFasta(dup, byte_range=(0, os.path.getsize(dup)))
    with pytest.raises(ValueError):
        shard.contig_table([shard.range_summary(f)])


def test_rows_from_device_text_and_slice_paths_agree():
    """igdetective._rows_from_device builds a row either from the letters that came with it (igd_fetch_row_text: three
    segments per V / J row, five per D row, in row order) or by slicing parent_seq itself; both must give the same rows,
    for both strands, both alignment directions, an empty fragment, and in reference order."""
    from igdetective_b200 import _lib, igdetective
    from igdetective_b200.fasta import reverse_complement
    from igdetective_b200.rss import RssScan
    from conftest import canonical
    canon = canonical("combined_reference_genes")
    rng = np.random.default_rng(4)
    L = 60_000
    s = np.frombuffer(b"ACGTacgtNR", np.uint8)[rng.integers(0, 10, L)].tobytes().decode()
    seqs = {"c0": s, "c1": s[:20_000][::-1]}
    n = 80
    vj = np.zeros(n, _lib.VJ_ROW_DTYPE)
    vj["contig"] = rng.integers(0, 2, n)
    vj["hept"] = np.sort(rng.integers(1000, 19_000, n)); vj["nona"] = vj["hept"] + 30
    vj["frag_start"] = vj["hept"] - 350; vj["frag_len"] = 350
    vj["gene_type"] = rng.integers(0, 2, n); vj["strand"] = rng.integers(0, 2, n)
    vj["best_gene"] = rng.integers(0, 13, n); vj["matches"] = rng.integers(150, 296, n); vj["alen"] = rng.integers(280, 320, n)
    vj["start"] = rng.integers(0, 60, n); vj["ient"] = rng.integers(300, 351, n)
    vj["direction"] = np.where(rng.integers(0, 2, n) == 1, 43, 45)
    vj["frag_len"][::11] = 0
    d = np.zeros(12, _lib.D_ROW_DTYPE)
    d["contig"] = rng.integers(0, 2, 12); d["strand"] = rng.integers(0, 2, 12)
    d["dl_hept"] = np.sort(rng.integers(1000, 19_000, 12)); d["dl_nona"] = d["dl_hept"] - 21
    d["dr_hept"] = np.where(d["strand"] == 0, d["dl_hept"] + 40, d["dl_hept"] - 40); d["dr_nona"] = d["dr_hept"] + 19
    hits = np.zeros(0, _lib.HIT_DTYPE)
    ids = list(seqs)
    scan = RssScan(None, seqs, order="canonical", load=False, hits=hits)
    res = {"vj": vj, "d": d, "hits": hits}
    want = igdetective._rows_from_device(res, ids, seqs, canon, scan, "canonical")
    # the segments as the library documents them (include/igd_b200.h: igd_fetch_row_text), from the rows just built
    segs, by_key = [], {}
    for g in ("V", "J"):
        for r in want[g]:
            by_key[(r[0], r[1], r[2], r[3], g)] = r
    for r in vj.tolist():
        hept, nona, ci, gt, sd = r[0], r[1], r[3], r[10], r[11]
        row = by_key[(ids[ci], "+-"[sd], hept, nona, "VJ"[gt])]
        segs += [row[4], row[5], row[12]]
    for r in want["D"]:
        segs += [r[4], r[5], r[8], r[9], r[12]]
    off = np.zeros(len(segs) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in segs])
    res_text = dict(res, text="".join(segs), text_off=off.tolist())
    blank = {c: "x" * len(v) for c, v in seqs.items()}            # the letters must come from the text, not from parent_seq
    got = igdetective._rows_from_device(res_text, ids, blank, canon, RssScan(None, blank, order="canonical", load=False, hits=hits), "canonical")
    assert got == want
    assert sum(len(want[g]) for g in want) == n + 12 and any(r[12] == "" for r in want["V"] + want["J"])
    assert reverse_complement("ACGTn") == "nACGT"
