#!/usr/bin/env python
"""bench.py -- headline benchmark of the RSS scan + V/D/J scoring hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mbases M]

Workload (BASELINE.json configs[1]): one synthetic 100 Mb contig per GPU, seed 2 (+rank): i.i.d.
background, 300 V + 30 D + 12 J mutated genes with their RSSs in 3 clusters, 20 N runs, 45 %
soft-masked; scored against combined_reference_genes IGHV (490) / IGHJ (13).
One step = one pass of the hot path: RSS scan of both strands for the 4 signal types, D pairing,
fragment extraction, scoring of every V / J candidate against every gene in both orientations
(the V table by exact pruning: bounds prove which alignments cannot matter, csrc/lcs.cu + detect.cu;
rows identical to the all-pairs scoring, which `align` / `roofline` time separately),
argmax + threshold + the winners' detail pass -> the rows of genes_{V,D,J}.tsv.
  value : Gb scanned per second over the whole step, contigs resident in HBM (packed)
  e2e   : the same through the public API from HOST buffers (pinned ASCII contig H2D + pack
          inside the timed region, hits and alignment results read back)
The reference arm (--impl reference) times the CPU oracle port of the same step on a bounded,
proportional sample (the reference itself needs BioPython, which is not installable here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DATA = os.path.join(ROOT, "tests", "golden", "datafiles")
METRIC, UNIT = "rss_scan_and_vdj_scoring_throughput", "Gb_scanned_per_s"


def load_data():
    from igdetective_b200 import datafiles
    motifs = datafiles.load_motifs(DATA)
    canon = {g: datafiles.load_canonical_genes("IGH", g, "combined_reference_genes", DATA) for g in ("V", "J")}
    return motifs, canon


def make_workload(seed, n_bases, motifs, canon):
    import synth
    rng = np.random.default_rng(seed)
    contigs, truth = synth.assembly(seed, [n_bases], motifs, canon["V"], canon["J"], n_runs=20, soft_mask=0.45,
                                    clusters=[(0, 100, 10, 4)] * 3)
    return contigs[0], truth


class Assembly3:
    """BASELINE.json configs[2]: a ~3 Gb mammalian-shaped assembly (40 SUPER_* scaffolds of 20-250 Mb summing to 2.9 Gb
    + 460 unplaced scaffolds of 10 kb-2 Mb), i.i.d. background with P(A,C,G,T) = (0.295, 0.205, 0.205, 0.295),
    IGH-like loci (60 V + 6 D + 3 J mutated genes with their RSSs, both strands) planted on 5 scaffolds, 3 N runs per
    scaffold.  Every 4 Mb block of every contig depends only on (seed, contig, block), and the planted loci / N runs are
    patches with fixed coordinates, so a rank can write just the pieces it owns -- straight into pinned memory -- and any
    sharding sees the same genome."""
    BLOCK = 1 << 22

    def __init__(self, motifs, canon, scale=1.0, seed=3):
        import synth
        rng = np.random.default_rng(seed)
        sup = np.exp(rng.uniform(np.log(20e6), np.log(250e6), 40))
        sup = (sup / sup.sum() * 2.9e9 * scale).astype(np.int64)
        unp = np.exp(rng.uniform(np.log(10e3), np.log(2e6), 460)).astype(np.int64)
        self.lengths = sup.tolist() + (unp * min(1.0, scale * 4)).astype(np.int64).tolist()
        self.ids = ["SUPER_%d" % (i + 1) for i in range(40)] + ["scaffold_%d" % i for i in range(460)]
        self.seed = seed
        t = np.array([0.295, 0.5, 0.705, 1.0]) * 65536
        self.lut = np.frombuffer(b"ACGT", np.uint8)[np.searchsorted(t, np.arange(65536), side="right").clip(0, 3)]
        self.patches = {}                                  # contig -> [(start, uint8 array)]
        for ci in (2, 7, 11, 23, 31):
            r = np.random.default_rng([seed, 1000 + ci])
            L = self.lengths[ci]
            w = min(L, 2_000_000)
            lo = int(r.integers(0, max(1, L - w)))
            win = synth.background(r, w)
            synth.plant_locus(r, win, motifs, canon["V"], canon["J"], 60, 6, 3, 0, w)
            self.patches.setdefault(ci, []).append((lo, win))
        for ci in range(40):
            r = np.random.default_rng([seed, 2000 + ci])
            for _ in range(3):
                n = int(r.integers(100, 50_000))
                p0 = int(r.integers(0, max(1, self.lengths[ci] - n)))
                self.patches.setdefault(ci, []).append((p0, np.full(n, ord("N"), np.uint8)))

    def fill(self, ci, a, b, out):
        """letters [a, b) of contig ci -> out[0 : b - a]"""
        B = self.BLOCK
        for blk in range(a // B, (b + B - 1) // B if b > a else a // B):
            lo, hi = blk * B, min((blk + 1) * B, self.lengths[ci])
            u = np.random.default_rng([self.seed, ci, blk]).integers(0, 65536, hi - lo, dtype=np.uint16)
            x, y = max(a, lo), min(b, hi)
            out[x - a:y - a] = self.lut[u[x - lo:y - lo]]
        for p0, arr in self.patches.get(ci, ()):
            x, y = max(a, p0), min(b, p0 + len(arr))
            if y > x:
                out[x - a:y - a] = arr[x - p0:y - p0]


def strong_config3(a, eng, motifs, canon, rank, world, barrier):
    """configs[2] strong-scaled: the 3.06 Gb assembly cut into pieces (shard.plan_pieces: contigs, scaffolds longer than
    48 Mb in ~32 Mb tiles with 512 bases of context), pieces assigned by LPT, every rank loads (pinned ASCII -> H2D ->
    pack), scans and scores its own pieces, rank 0 gathers the rows.  No data-path collective.  Whole-job time =
    max over ranks of (load + detect + gather), CUDA-synchronised, between barriers."""
    import hashlib
    import torch
    import torch.distributed as dist
    from igdetective_b200 import multi, shard
    from igdetective_b200.fasta import ArraySeq
    asm = Assembly3(motifs, canon, a.strong_scale)
    t0 = time.perf_counter()
    pieces = shard.plan_pieces(asm.lengths, world)[rank]
    n_mine = sum(p[4] - p[3] for p in pieces)
    pinned = torch.empty(max(n_mine, 1), dtype=torch.uint8).pin_memory()
    host = pinned.numpy()
    off = np.zeros(len(pieces) + 1, np.uint64)
    part = {}
    for k, (ci, _, _, p0, p1) in enumerate(pieces):
        off[k + 1] = off[k] + (p1 - p0)
        asm.fill(ci, p0, p1, host[int(off[k]):int(off[k + 1])])
        part[k] = ArraySeq(host[int(off[k]):int(off[k + 1])])
    t_gen = time.perf_counter() - t0
    shard.gather_rows({"warmup": []}, asm.ids)                 # first object collective pays communicator set-up
    samples, merged, scan_ms, tm = [], None, 0.0, []
    # auto: one group per ~400 MB of this rank's letters (3 Gb on one GPU: 8 groups; an eighth of it: one load, one detect)
    groups = max(1, a.strong_groups) if a.strong_groups > 0 else int(max(1, min(8, n_mine // 400_000_000)))
    engines = [eng]
    if groups > 1:
        from igdetective_b200.engine import Engine
        for _ in range(max(1, a.strong_handles - 1)):          # more handles (streams) on the same GPU
            engines.append(Engine(eng.device, motifs))
    for it in range(1 + a.strong_reps):
        barrier()
        t0 = time.perf_counter()
        if groups > 1:
            # the pieces in `groups` runs over two handles: group k + 1 is uploaded while group k is scanned and scored
            tm = []
            rows = multi.detect_pieces_pipelined(engines, asm.ids, pieces, canon, part, (pinned.data_ptr(), off), groups, tm)
            t2 = time.perf_counter()
            t1 = t0 + sum(x[1] for x in tm)                    # time spent uploading (the rest overlaps it)
            t_detect = sum(x[2] for x in tm)
        else:
            eng.load_contigs((pinned.data_ptr(), off), list(part))
            t1 = time.perf_counter()
            rows = multi.detect_pieces(eng, None, asm.ids, pieces, canon, "canonical", part=part)
            eng.sync()
            t2 = time.perf_counter()
            t_detect = t2 - t1
        merged = shard.gather_rows(rows, asm.ids)
        t3 = time.perf_counter()
        if groups > 1:                                         # kernel times: sums over the groups (together they cover the shard)
            v = [t3 - t0, t1 - t0, t_detect, t3 - t2, sum(x[3] for x in tm), sum(x[4] for x in tm)]
        else:
            v = [t3 - t0, t1 - t0, t_detect, t3 - t2, eng.timing()["scan_ms"], eng.timing()["pack_ms"]]
        if world > 1:
            t = torch.tensor(v, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v = t.tolist()
        if it > 0:
            samples.append(v)
    if rank != 0:
        return None
    med = np.median(np.array(samples), axis=0).tolist()
    text = "".join("\t".join(map(str, r)) + "\n" for g in ("V", "D", "J") for r in merged[g])
    bases = int(sum(asm.lengths))
    return {"workload": "configs[2]: synthetic %.2f Gb assembly, %d contigs (40 SUPER_* scaffolds + 460 unplaced), whole-genome "
                        "scan + scoring vs combined IGHV/IGHJ, minimap2 prefilter bypassed" % (bases / 1e9, len(asm.lengths)),
            "scaling": "strong", "n_gpus": world, "bases": bases, "reps": a.strong_reps,
            "whole_job_s": med[0], "gb_scanned_per_s": bases / med[0] / 1e9,
            "load_h2d_pack_s": med[1], "scan_and_scoring_s": med[2], "gather_s": med[3],
            "scan_kernel_ms_max": med[4], "pack_kernel_ms_max": med[5],
            "timing": "median over reps of the max over ranks; host ASCII in pinned memory, H2D + pack inside",
            "pipeline": ("%d groups of pieces over %d handles: the upload of group k + 1 overlaps the scan + scoring of group k, so "
                         "load_h2d_pack_s + scan_and_scoring_s (sums over the groups, like scan_kernel_ms_max) exceed "
                         "whole_job_s; the pack kernels are queued without a wait and not timed (0)" % (groups, len(engines))) if groups > 1 else "none (one load, one detect)",
            "groups_timeline_rank0": [{"group": x[0], "upload_start_s": x[6], "upload_s": x[1], "detect_s": x[2],
                                       "scoring_kernel_ms": x[7]} for x in sorted(tm)] if groups > 1 else None,
            "rank0_pieces": len(pieces), "rank0_bases": int(n_mine), "generate_s_rank0": t_gen,
            "rows": {g: len(merged[g]) for g in ("V", "D", "J")},
            "rows_sha1": hashlib.sha1(text.encode()).hexdigest()}


class Clocks:
    """nvidia-smi sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.first = [], None, 0
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_first(self, timeout):
        """until the sampler has produced its first line (it is then in its steady polling loop)"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """samples from here on belong to the timed regions"""
        self.first = max(0, len(self.rows) - 1)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        self.t.join(timeout=2)
        self.rows = self.rows[self.first:] or self.rows
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return None
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "samples": len(sm),
                "reasons": reasons}


# --------------------------------------------------------------------------------------- GPU arm
def gpu_arm(a):
    import torch
    import torch.distributed as dist
    from igdetective_b200 import igdetective
    from igdetective_b200.engine import Engine
    from igdetective_b200.rss import SPACER_LENGTH

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    motifs, canon = load_data()
    n_bases = int(a.mbases * 1e6)
    contig, truth = make_workload(2 + rank, n_bases, motifs, canon)
    # host side of the public API: the contig as the reference holds it (a str) and as pinned bytes
    pinned = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = contig
    seqs = {"synthetic_%d" % rank: contig.tobytes().decode("latin-1")}
    off = np.array([0, n_bases], np.uint64)
    eng = Engine(local, motifs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        eng.sync()
        if world > 1:
            dist.barrier()

    stats = {}

    def step(e2e, record=False):
        """One pass of the hot path: contigs -> the rows of genes_{V,D,J}.tsv (strings included)."""
        if e2e:
            eng.load_contigs((pinned.data_ptr(), off), list(seqs))
        info, rows = igdetective.detect(eng, seqs, canon, order="canonical", load=False)
        if record:                          # workload description for the JSON line, outside the timed steps
            stats["rows"] = {g: len(r) for g, r in rows.items()}
            stats["rss"] = {st: sum(len(v) for sd in info[st] for v in info[st][sd].values()) for st in info}
        return rows

    def timed(e2e, k):
        """K steps, each bracketed by barrier + synchronize; L2 flushed between steps outside the clock."""
        total, scan_ms, align_ms, cells, launches = 0.0, [], 0.0, 0, 0
        h0, d0 = eng.h2d_bytes, eng.d2h_bytes
        for _ in range(k):
            flush.zero_()
            barrier()
            l0 = eng.timing()["total_launches"]
            t0 = time.perf_counter()
            step(e2e)
            eng.sync()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            total += dt
            launches += eng.timing()["total_launches"] - l0
        return total, (eng.h2d_bytes - h0) / k, (eng.d2h_bytes - d0) / k, launches / k

    # the clock sampler starts BEFORE the warm-up: nvidia-smi takes a few hundred ms to come up, and that must not fall
    # into the 5-20 steps of the timed region (it did: steps of 4.5 ms measured 5.2 ms with the sampler starting beside them)
    clocks = Clocks(local) if rank == 0 else None
    eng.load_contigs((pinned.data_ptr(), off), list(seqs))
    for _ in range(a.warmup):
        step(False)
    step(False, record=True)
    if clocks:
        clocks.wait_first(2.0)
        clocks.mark()
    t_dev, _, _, launches = timed(False, a.steps)
    # per-kernel device times of the last resident step: CUDA events on the library's stream
    eng.load_contigs((pinned.data_ptr(), off), list(seqs))
    flush.zero_(); barrier()
    eng.scan_rss_count(SPACER_LENGTH)
    tm = eng.timing()
    scan_ms = tm["scan_ms"]
    scan_samples = []
    for _ in range(max(3, a.steps)):
        flush.zero_(); barrier()
        eng.scan_rss_count(SPACER_LENGTH)
        scan_samples.append(eng.timing()["scan_ms"])
    scan_ms = float(np.mean(scan_samples))
    kstat = kernel_stats(eng, seqs, canon, igdetective)
    flush.zero_(); barrier()
    step(False)
    tdet = eng.timing()              # the library's CUDA-event times of one resident step (igd_detect)
    t_e2e, h2d, d2h, _ = timed(True, a.steps)
    clk = clocks.stop() if clocks else None
    # scan kernel alone at assembly scale (rank 0): the workload contig repeated, so that the fixed cost of a launch
    # (~15 us of the ~45 us a 100 Mb scan takes) does not hide the kernel's streaming rate
    scan_large = None
    if rank == 0 and a.scan_mbases > a.mbases:
        reps = int(round(a.scan_mbases / a.mbases))
        big = np.tile(contig, reps)
        eng.load_contigs((big, np.array([0, len(big)], np.uint64)), ["large"])
        samples = []
        for _ in range(max(3, a.steps) + 1):
            flush.zero_(); torch.cuda.synchronize(); eng.sync()
            n_large = eng.scan_rss_count(SPACER_LENGTH)
            samples.append(eng.timing()["scan_ms"])
        ms = float(np.mean(samples[1:]))
        scan_large = {"bases": int(len(big)), "kernel_ms": ms, "hits": int(n_large),
                      "achieved": (len(big) + 3) // 4 / (ms * 1e-3) / 1e9}
        del big

    strong = None
    if a.strong_reps > 0:
        del pinned
        strong = strong_config3(a, eng, motifs, canon, rank, world, barrier)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    total_bases = n_bases * world
    alg_bytes = (n_bases + 3) // 4
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    sm_hz = (clk or {}).get("sm_mhz", 1965.0) * 1e6
    int_peak = 148 * 64 * sm_hz / 1e12            # ALU-pipe lane-ops/s (T), 64 lanes/clk/SM
    sass = gotoh_instruction_counts()
    traffic = ncu_traffic()
    alu_per_cell = sass["alu_pipe_per_cell"]
    line = {
        "metric": METRIC, "value": total_bases / (t_dev / a.steps) / 1e9, "unit": UNIT,
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic %.0f Mb single contig per GPU, planted RSS + mutated V/D/J, "
                               "both strands, 4 signal types, scored vs combined_reference_genes IGHV(490)/IGHJ(13)"
                               % a.mbases,
                   "order": "canonical", "l2": "flushed (256 MiB write) between steps, outside the timed regions",
                   "rows": stats.get("rows"), "rss": stats.get("rss")},
        "e2e": {"value": total_bases / (t_e2e / a.steps) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / a.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "scan": {"kernel_ms": scan_ms, "gb_per_s_both_strands": n_bases / (scan_ms * 1e-3) / 1e9,
                 "hits": kstat["hits"]},
        # the packed Gotoh kernel on ALL (candidate, gene, orientation) pairs of the step, as the matrix-returning seam
        # align_fragment_to_genes computes them: the kernel measurement behind `roofline`
        "align": {"kernel_ms": kstat["align_ms"], "gcups": kstat["gcups"], "pairs": kstat["pairs"],
                  "cells": kstat["cells"]},
        # what the step itself runs: LCS bounds + the alignments that can matter (exact pruning, csrc/detect.cu) + the
        # winners' detail pass
        "scoring_in_step": {"kernel_ms": tdet["align_ms"], "dp_cells": int(tdet["align_cells"]),
                            "fraction_of_all_cells": tdet["align_cells"] / max(kstat["cells"], 1),
                            "launches": int(tdet["align_launches"]),
                            # how the kernel time splits: Gotoh kernels (small batches: three rounds + the winners' detail
                            # pass), K4 = LCS bound + top-K, K5 = insertion-charging bound of the fragments that resemble no gene
                            "gotoh_ms": tdet["dp_ms"], "gotoh_gcups": tdet["align_cells"] / max(tdet["dp_ms"], 1e-6) / 1e6,
                            "k4_lcs_ms": tdet["lcs_ms"], "k4_letter_pairs": int(tdet["lcs_cells"]),
                            "k4_tera_letter_pairs_per_s": tdet["lcs_cells"] / max(tdet["lcs_ms"], 1e-6) / 1e9,
                            "k5_ms": tdet["k5_ms"], "k5_letter_pairs": int(tdet["k5_cells"]),
                            "k5_tera_letter_pairs_per_s": tdet["k5_cells"] / max(tdet["k5_ms"], 1e-6) / 1e9},
        "kernel_share_of_step": {"rss_scan": scan_ms / (1e3 * t_dev / a.steps),
                                 "scoring_kernels": tdet["align_ms"] / (1e3 * t_dev / a.steps)},
        # largest kernel of the step (60 % of GPU time in the ncu launch list, K5 20 %, K4 7 %): integer DP, bound by the
        # ALU pipe (neither HBM nor tensor); achieved = cell updates/s x ALU-pipe instructions per cell, on all pairs
        "roofline": {"kernel": "gotoh_packed_kernel<%s,%s>" % V_KERNEL[:2], "bound": "alu", "achieved": kstat["gcups"] * alu_per_cell / 1e3,
                     "peak": int_peak, "unit": "T lane-ops/s", "frac": kstat["gcups"] * alu_per_cell / 1e3 / int_peak,
                     # dram bytes of one launch of this kernel (inputs are KB per pair): ncu --set full capture of this round
                     "traffic": traffic["gotoh_packed"]["dram_bytes_per_launch"], "traffic_source": traffic["gotoh_packed"]["source"],
                     "gcups": kstat["gcups"], "alu_ops_per_cell": alu_per_cell, "all_ops_per_cell": sass["all_per_cell"],
                     "fma_ops_per_cell": sass["fma_pipe_per_cell"], "ops_source": sass["source"],
                     # the same against the issue port: every instruction of the hot path, 128 lanes/clk/SM
                     "issue_frac": kstat["gcups"] * sass["all_per_cell"] / 1e3 / (2 * int_peak),
                     "peak_source": "148 SMs x 64 ALU lanes/clk x SM clock sampled during the run",
                     # what the number is measured on, and what the same kernel does inside the timed steps
                     "measured_on": "all (candidate, gene, orientation) pairs of the step in one dense pass of this kernel after "
                                    "the timed steps (`align`); the step itself aligns only the pairs its bounds leave over",
                     "in_step": {"kernel_ms": tdet["dp_ms"], "cells": int(tdet["align_cells"]),
                                 "gcups": tdet["align_cells"] / max(tdet["dp_ms"], 1e-6) / 1e6,
                                 "share_of_step": tdet["dp_ms"] / (1e3 * t_dev / a.steps)}},
        # the HBM-bound kernel of the path (north star: scan vs HBM roofline), 0.25 algorithmic bytes per base
        "roofline_scan": {"kernel": "rss_scan_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                          "unit": "GB/s", "frac": achieved / hbm_peak, "algorithmic_bytes": alg_bytes,
                          # dram__bytes_read+write of one launch, ncu --set full capture in profiles/: equals the algorithmic bytes
                          "traffic": traffic["rss_scan_100Mb"]["dram_bytes_per_launch"] if a.mbases == 100.0 else None,
                          "traffic_source": traffic["rss_scan_100Mb"]["source"],
                          "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650"},
    }
    if scan_large:
        line["roofline_scan_large"] = {"kernel": "rss_scan_kernel", "bound": "hbm", "achieved": scan_large["achieved"],
                                       "peak": hbm_peak, "unit": "GB/s", "frac": scan_large["achieved"] / hbm_peak,
                                       "workload": "the %.0f Mb contig repeated to %.0f Mb, scan kernel only, L2 flushed"
                                                   % (a.mbases, scan_large["bases"] / 1e6),
                                       "kernel_ms": scan_large["kernel_ms"], "hits": scan_large["hits"],
                                       "gb_per_s_both_strands": scan_large["bases"] / (scan_large["kernel_ms"] * 1e-3) / 1e9}
    if strong:
        # the multi-GPU split of the north star (contigs / pieces sharded over the GPUs), strong-scaled at this N; `value`
        # above stays the same workload at every N (one configs[1] contig per GPU) so that the per-N values are comparable
        line["strong_config3"] = strong
        alg3 = (strong["rank0_bases"] + 3) // 4
        line["roofline_scan_config3"] = {"kernel": "rss_scan_kernel", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
                                         "achieved": alg3 / (strong["scan_kernel_ms_max"] * 1e-3) / 1e9,
                                         "frac": alg3 / (strong["scan_kernel_ms_max"] * 1e-3) / 1e9 / hbm_peak,
                                         "workload": "rank 0's shard of configs[2] (%d bases)" % strong["rank0_bases"]}
    if not a.no_cpu:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        line["cpu_baseline"] = cpu_sample(contig, seqs, motifs, canon, kstat["fragments"], a)
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()


V_KERNEL = ("16", "22", "0", "0", "1")       # gotoh_packed_kernel<G, C, MIXED, LONG, FPINC>: the instance the 350-nt V flanks run in
SASS_JSON = os.path.join(ROOT, "profiles", "r02_sass_counts_gotoh_16x22.json")
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def gotoh_instruction_counts():
    """ALU-pipe / all instructions per DP cell of the dominant kernel, counted in the SASS of the library that is loaded
    (profiles/sass_derive.py); the committed count of the same derivation is used only where cuobjdump is missing, and a
    difference between the two is reported.  No typed-in constants: without either source the bench fails."""
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import sass_derive
    from igdetective_b200 import _lib
    committed = json.load(open(SASS_JSON)) if os.path.exists(SASS_JSON) else None
    try:
        res, _ = sass_derive.derive(_lib.LIB_PATH, V_KERNEL, int(V_KERNEL[1]))
        res["source"] = "cuobjdump -sass of the loaded library in this run (profiles/sass_derive.py)"
        if committed and committed["counts"] != res["counts"]:
            res["source"] += "; differs from the committed " + os.path.relpath(SASS_JSON, ROOT)
        return res
    except (OSError, subprocess.CalledProcessError, RuntimeError) as e:
        if committed is None:
            raise RuntimeError("no instruction counts: cuobjdump failed (%s) and %s is missing" % (e, SASS_JSON))
        committed["source"] = "%s (cuobjdump unavailable here: %s)" % (os.path.relpath(SASS_JSON, ROOT), e)
        return committed


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two roofline kernels, from the ncu --set full
    captures of this round (profiles/ncu_traffic.py writes the file from the .ncu-rep reports)."""
    return json.load(open(TRAFFIC_JSON))


def kernel_stats(eng, seqs, canon, igdetective):
    """One resident step with the library's own CUDA-event timings pulled after each call."""
    from igdetective_b200.rss import (D, DL, DR, FWD, J, REV, SIGNAL_TYPES, V, RssScan, combine_D_RSS,
                                      get_s_fragment_from_RSS)
    from igdetective_b200.scoring import align_fragment_to_genes
    scan = RssScan(eng, seqs, order="canonical", load=False)
    info = {st: {sd: scan.contigwise(st, sd) for sd in (FWD, REV)} for st in SIGNAL_TYPES}
    frags = {g: {sd: get_s_fragment_from_RSS(g, sd, info, seqs) for sd in (FWD, REV)} for g in (V, J)}
    ms, cells, pairs, flat_all = 0.0, 0, 0, {}
    for g in (V, J):
        flat = [f for sd in (FWD, REV) for c in frags[g][sd] for f in frags[g][sd][c]]
        flat_all[g] = flat
        align_fragment_to_genes(flat, list(canon[g].values()), engine=eng)
        t = eng.timing()
        ms += t["align_ms"]; cells += t["align_cells"]; pairs += 2 * len(flat) * len(canon[g])
    return {"hits": int(len(scan.hits)), "align_ms": ms, "cells": int(cells), "pairs": int(pairs),
            "gcups": cells / (ms * 1e-3) / 1e9 if ms > 0 else 0.0, "fragments": flat_all}


# --------------------------------------------------------------------------------------- CPU arms
def _scan_piece(args):
    """The reference's scan loops, literally (py/IGDetective.py:138-177), on one piece of the contig."""
    from oracle import igd_oracle as O
    piece, motifs = args
    n = 0
    for st in O.SIGNAL_TYPES:
        for sd in "+-":
            r = O.get_contigwise_rss(st, sd, {"p": piece}, motifs, order="reference", scan=O.find_valid_motif_idx_py)
            n += len(r["p"])
    return n


def cpu_step(sample_seq, motifs, frag_sample, canon, cores):
    """Oracle port of one step on a sample: pure-Python scan loops over `sample_seq` (split over
    `cores` processes) + scoring of `frag_sample` = {gene type: fragments} against all genes."""
    import multiprocessing as mp
    from oracle import igd_oracle as O
    t0 = time.perf_counter()
    n = max(1, cores)
    sz = (len(sample_seq) + n - 1) // n
    pieces = [sample_seq[max(0, i * sz - 40):(i + 1) * sz] for i in range(n)]
    with mp.get_context("fork").Pool(n) as pool:
        pool.map(_scan_piece, [(p, motifs) for p in pieces])
    t_scan = time.perf_counter() - t0
    t0 = time.perf_counter()
    cells = 0
    for g, frs in frag_sample.items():
        if frs:
            glist = list(canon[g].values())
            # the reference runs this computation twice per gene type ('AFFINE' and 'MAXK', py/IGDetective.py:467-468:
            # set_aligner ignores its argument), so the CPU arm does too; cells count one pass per orientation
            for _ in range(2):
                O.align_fragment_to_genes(frs, glist, threads=cores)
            cells += 2 * sum(max(len(f), 1) for f in frs) * sum(len(x) for x in glist)
    t_score = time.perf_counter() - t0
    return t_scan, t_score, cells


def cpu_sample(contig, seqs, motifs, canon, fragments, a, frac=None):
    """Bounded proportional sample of the workload: `frac` of the contig and of the candidates."""
    cores = len(os.sched_getaffinity(0))
    n = len(contig)
    frac = frac or min(1.0, 2.0e6 / n)
    s_len = int(n * frac)
    sample_seq = next(iter(seqs.values()))[:s_len]
    fs = {g: fr[:max(1, int(round(len(fr) * frac)))] for g, fr in fragments.items()}
    t_scan, t_score, cells = cpu_step(sample_seq, motifs, fs, canon, cores)
    return {"value": s_len / (t_scan + t_score) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%.2f%% of the step: pure-Python reference scan loops over the first %d bases split over "
                      "%d processes (%.2fs) + C oracle aligner, %d threads, on %s of the V/J candidates x all genes "
                      "x 2 orientations, run twice as the reference does (%.2fs, %.1f MCUPS per pass)" % (100 * frac, s_len, cores, t_scan, cores,
                      {g: len(f) for g, f in fs.items()}, t_score, 2 * cells / max(t_score, 1e-9) / 1e6),
            "scan_mb_per_s": s_len / t_scan / 1e6, "align_mcups": 2 * cells / max(t_score, 1e-9) / 1e6}


def reference_arm(a):
    """CPU-only: the oracle port of the step on a bounded proportional sample, all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import igd_oracle as O
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    motifs, canon = load_data()
    n_bases = int(a.mbases * 1e6)
    contig, truth = make_workload(2, n_bases, motifs, canon)
    seqs = {"synthetic_0": contig.tobytes().decode("latin-1")}
    frac = min(1.0, 2.0e6 / n_bases)
    # candidates of the whole workload (untimed set-up; vectorised oracle scan), then a proportional share
    rss = {g: {sd: O.get_contigwise_rss(g, sd, seqs, motifs, order="canonical") for sd in "+-"} for g in ("V", "J")}
    fragments = {g: [f for sd in "+-" for c in seqs for f in O.s_fragments(g, sd, rss, seqs)[c]] for g in ("V", "J")}
    cores = len(os.sched_getaffinity(0))
    res = None
    t_all = 0.0
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        res = cpu_sample(contig, seqs, motifs, canon, fragments, a, frac)
        if i >= a.warmup:
            t_all += time.perf_counter() - t0
    s_len = int(n_bases * frac)
    val = s_len * a.steps / t_all / 1e9
    res["value"] = val
    emit_line({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", a.gpus)),
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_all / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic %.0f Mb single contig, bounded proportional sample (%.2f%% "
                               "of the contig and of the V/J candidates per step)" % (a.mbases, 100 * frac)},
        "cpu_baseline": res,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else a library prints there during the run (NCCL's version
    banner under NCCL_DEBUG=VERSION, for one) is sent to stderr, and the line goes out on the real stdout at the end."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mbases", type=float, default=100.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--strong-reps", type=int, default=3, help="timed repetitions of the strong-scaled configs[2] leg (0 = skip)")
    ap.add_argument("--strong-scale", type=float, default=1.0, help="size of the configs[2] assembly relative to 3.06 Gb")
    ap.add_argument("--strong-handles", type=int, default=4, help="configs[2] leg: handles the groups rotate over")
    ap.add_argument("--strong-groups", type=int, default=0, help="configs[2] leg: groups of pieces pipelined over the handles (1 = one load, one detect; 0 = one per ~400 MB)")
    ap.add_argument("--scan-mbases", type=float, default=1000.0,
                    help="size of the extra scan-only measurement (the workload contig repeated; 0 = skip)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)
