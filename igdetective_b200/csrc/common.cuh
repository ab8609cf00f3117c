// Shared declarations of the igd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/igd_b200.h"

// ---- packed contig layout in HBM ---------------------------------------------------------------
// One "chunk" = 64 consecutive bases = one 16-byte record {lo0, hi0, lo1, hi1}:
//   bit b of lo0 / hi0 = low / high bit of the 2-bit code of base b       (bases  0..31)
//   bit b of lo1 / hi1 = the same for base 32 + b                          (bases 32..63)
// so the array is also a sequence of aligned 8-byte (lo, hi) pairs, one per 32 bases: any 32
// consecutive positions of both planes are two adjacent pairs and two funnel shifts.
// codes A=0 C=1 G=2 T=3 (complement = both bits flipped); anything else packs as 0 and sets the
// base's bit in the validity plane inv[chunk] (one uint64 per chunk) and the chunk's bit in the
// summary bitmap invsum[chunk / 32].  The scan reads the 16-byte records (0.25 B/base) for every
// base and the validity plane only for windows that survive the motif lookups.
// Chunk 0 is an all-invalid pad; each contig starts on a chunk boundary and is followed by at
// least one all-invalid pad chunk, so no k-mer window or RSS can straddle two contigs; the array
// is padded with invalid chunks to a whole number of tiles plus one trailing halo chunk.
constexpr int IGD_CHUNK_BASES = 64;
constexpr int IGD_TILE_CHUNKS = 512;                  // chunks per scan tile = threads per CTA
constexpr int IGD_STAGE_CHUNKS = IGD_TILE_CHUNKS + 2; // one halo chunk each side (RSS span <= 40 bases)

struct igd_fasta {
    uint8_t *seq = nullptr;            // letters of all contigs, concatenated in dict order (malloc)
    std::vector<uint64_t> off;         // n + 1 offsets
    std::vector<std::string> ids;
    std::vector<uint64_t> hdr_byte;    // file offset of each record's header line (UINT64_MAX: the lead of a range)
    uint64_t range_lo = 0, range_hi = 0;   // igd_fasta_open_range: the line-aligned byte range that was asked for ...
    uint64_t own_lo = 0, own_hi = 0;       // ... and the letters [own_lo, own_hi) of seq that come from it (the rest is margin)
    ~igd_fasta() { free(seq); }
};

struct igd_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    int sm_count = 148;
    bool fp_increment_ok = false;  // exact denormal arithmetic verified on this device (gotoh.cu, FPINC)

    // motif tables (kernel index order: (hi bits << k) | lo bits, first base = bit 0)
    uint8_t *d_lut7 = nullptr;    // 16384 flags: bit t = strand +, bit 4+t = strand -
    uint8_t *d_lut9 = nullptr;    // 262144 flags
    uint32_t *d_any9 = nullptr;   // 262144-bit bitmap: lut9 != 0
    uint32_t *d_any9c = nullptr;  // 262144-bit bitmap: the 9-mer or a neighbour shifted by one base is in lut9 (igd_center_bitmap)
    uint32_t *d_any7 = nullptr;   // 16384-bit bitmap: lut7 != 0
    uint8_t *d_need7 = nullptr;   // 16384: windows of the last igd_scan's spacers that serve a 7-mer's flags (rss_scan.cu)
    std::vector<uint8_t> h_lut7;  // host copy of d_lut7
    bool need7_valid = false;
    bool scan_attr_set[4] = {false, false, false, false};   // dynamic shared memory size set for the scan kernel instances
    int32_t need7_spacer[4] = {0, 0, 0, 0};
    bool motifs_set = false;
    bool consensus_prefilter = false;   // every heptamer flag entry matches CAC.GTG in >= 4 of 6 positions

    // contigs
    uint4 *d_planes = nullptr;
    unsigned long long *d_inv = nullptr;
    uint32_t *d_invsum = nullptr;
    uint64_t n_chunks = 0;        // allocated chunks (incl. pads)
    uint64_t n_tiles = 0;
    std::vector<uint64_t> contig_chunk0;   // first chunk of each contig
    std::vector<uint64_t> contig_len;
    std::vector<uint64_t> contig_seq_off;  // first letter of each contig in the resident letter buffer (S_SEQ)
    uint64_t total_bases = 0;
    bool contigs_loaded = false;
    size_t planes_cap = 0;        // chunks of capacity currently allocated

    // hit buffers
    unsigned long long *d_hits = nullptr;
    uint64_t hit_cap = 0;
    unsigned long long *d_count = nullptr;
    std::vector<unsigned long long> host_hits;   // last scan's raw keys
    std::vector<igd_hit> sorted_hits;            // the same decoded and sorted (igd_fetch_hits order)
    bool sorted_hits_valid = false;
    std::vector<igd_vj_row> vj_rows;             // last igd_detect
    std::vector<igd_d_row> d_rows;
    std::vector<uint8_t> row_text;               // the rows' letters (igd_fetch_row_text), segment k = [row_text_off[k], row_text_off[k + 1])
    std::vector<uint64_t> row_text_off;
    uint8_t *d_comp = nullptr;                   // 256-byte complement table (Bio.Seq's IUPAC table), detect.cu
    int last_scan_kind = 0;                      // 0 none, 1 rss, 2 kmers
    int last_k = 0;
    int32_t last_spacer[4] = {0, 0, 0, 0};

    // generic growable device scratch for the aligner
    void *d_scratch[40] = {nullptr};
    size_t scratch_cap[40] = {0};
    void *h_stage = nullptr;      // pinned staging buffer for small result read-backs
    size_t h_stage_cap = 0;
    void *h_pin[3] = {nullptr, nullptr, nullptr};   // more pinned buffers: [0] bounds read-back, [1] / [2] run tables (alignments / K5) on their way to the device
    size_t h_pin_cap[3] = {0, 0, 0};
    void *h_ring = nullptr;       // ring of pinned pieces for the upload of pageable genomes (capi.cu, upload_letters)
    cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};

    igd_timing timing = {};
};

void igd_set_error(const char *fmt, ...);
int igd_cuda_fail(cudaError_t e, const char *what);
#define IGD_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return igd_cuda_fail(_e, #call);            \
    } while (0)

// growable device scratch slots (capi.cu, detect.cu)
enum { S_SEQ = 0, S_SEQOFF, S_CHUNK0, S_TGT, S_TGTOFF, S_QRY, S_QRYOFF, S_PT, S_PQ, S_WORK, S_SCORE, S_PAY, S_RUNS, S_PICK,
       S_FRAG, S_BEST, S_TILEC, S_LCS, S_MASKS, S_UB, S_TOP, S_QCLS, S_YRUNS, S_Y, S_HITREC, S_HITOUT, S_HITKEY, S_HITIDX, S_CUBTMP, S_YPASS, S_TCLS, S_TEXT, S_TEXTDESC, S_COUNT };
static_assert(S_COUNT <= 40, "igd_handle::d_scratch is too small");
int igd_reserve(igd_handle *h, int slot, size_t bytes);
int igd_stage_reserve(igd_handle *h, size_t bytes);       // pinned read-back buffer h->h_stage
int igd_pin_reserve(igd_handle *h, int which, size_t bytes);   // h->h_pin[which]
// 0: no lower-case letter anywhere, 1: only in targets, 2: in queries (and maybe targets)
int igd_case_mix(const uint8_t *tgt, size_t tbytes, const uint8_t *qry, size_t qbytes);
// Every target against every query through the packed run-mode kernels, results left on the device
// (S_SCORE, S_PAY: [n_tgt][n_qry]); *launched = false (and nothing done) when some pair does not fit the
// packed kernels.  tgt_resident: the targets and their offsets already sit in S_TGT / S_TGTOFF.
int igd_dense_launch(igd_handle *h, const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                     const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                     const igd_scoring *sc, bool *launched, bool tgt_resident);
bool igd_packed_scheme_ok(const igd_scoring *sc);
int igd_packed_layout_for(const igd_scoring *sc, int nA, int nB);     // 0: two-register kernel, 1 / 2: packed short / long layout
// lcs.cu: out[t][g] = LCS(upper(target t), upper(gene g)), an upper bound on BioAlign's match count + 1
int igd_launch_lcs(igd_handle *h, const uint8_t *d_tgt, const int32_t *d_tgt_off, int n_tgt, size_t tbytes,
                   const uint8_t *d_qry, const int32_t *d_qry_off, int n_qry, int max_q, uint32_t *d_masks, uint16_t *d_out);
int igd_launch_ub_topk(igd_handle *h, const uint16_t *d_lcs, const int32_t *d_qry_off, int n_unique, int n_qry, int K,
                       uint16_t *d_ub, int32_t *d_top);
int igd_launch_gene_classes(igd_handle *h, const uint8_t *d_qry, size_t n, uint8_t *d_cls);
// Y bound (lcs.cu, K5): out[slot + k] = Y(rc) << 16 | Y(forward) per run {fragment, first gene, count, slot}
int igd_launch_ybound(igd_handle *h, int shape, const uint8_t *d_tgt, const int32_t *d_tgt_off, const uint8_t *d_qcls,
                      const int32_t *d_qry_off, const int4 *d_runs, int n_runs, uint32_t *d_out, unsigned int *d_counter,
                      double cut, uint2 *d_pass, unsigned int *d_pass_count, unsigned int pass_cap);
// scan + hit decoding shared by igd_scan / igd_fetch_hits / igd_detect
int igd_run_rss_scan(igd_handle *h, const int32_t spacer[4], uint64_t *n_hits);
int igd_sorted_hits(igd_handle *h, uint64_t *n);          // fills h->sorted_hits from the last RSS scan
int igd_sort_hits_device(igd_handle *h, uint64_t n, bool *done);   // hits.cu: the same on the device

// pack.cu
int igd_launch_pack(igd_handle *h, const uint8_t *d_seq, const unsigned long long *d_seq_off,
                    const unsigned long long *d_chunk0, const int *d_tile_contig, int n_contigs);
// rss_scan.cu
// prepare_only: upload the spacer-dependent table and load / configure the kernel instance, no launch (keeps the
// first scan's one-off costs out of the timed window)
int igd_launch_rss_scan(igd_handle *h, const int32_t spacer[4], bool prepare_only = false);
int igd_launch_kmer_scan(igd_handle *h, int k);
// gotoh.cu
struct igd_align_job {
    const uint8_t *d_tgt; const int32_t *d_tgt_off;
    const uint8_t *d_qry; const int32_t *d_qry_off;
    const int32_t *d_pair_t; const int32_t *d_pair_q;
    const int32_t *d_work;      // pair indices of this bucket (two-register kernel)
    const int4 *d_runs;         // runs of this bucket (packed kernel): {target, first query, count, first output slot}
    int32_t n_work;
    int bucket;                 // kernel shape selector
    int mixed_case;             // 0: no lower-case letters; 1: only in targets; 2: in queries (and maybe targets)
    int packed;                 // 0: two-register kernel; 1 / 2: single-register kernel, short / long field layout
    int pay_mode;               // 0: (matches, internal Ix, lead)   1: (matches, internal Ix, trailing Ix)
    igd_scoring sc;
    int32_t *d_score; uint32_t *d_pay;
    unsigned int *d_counter;
};
int igd_align_bucket_of(int target_len);
int igd_probe_fp_increment(igd_handle *h, bool *ok);
// out[u][q] = matches | alen << 10 | forward << 31 of the better orientation of fragment u (rows 2u, 2u+1 of pay)
// out[w] = {o * n_qry + g of the last maximal PI >= 0 over [window, RC] x genes, or -1; matches; len; 0}
int igd_launch_select_windows(igd_handle *h, const uint32_t *d_pay, const int32_t *d_tgt_off, const int32_t *d_qry_off,
                              int32_t n_win, int32_t n_qry, int4 *d_out);
int igd_launch_pick(igd_handle *h, const uint32_t *d_pay, const int32_t *d_qry_off, int32_t n_unique, int32_t n_qry,
                    uint32_t *d_out);
int igd_launch_gotoh(igd_handle *h, const igd_align_job &job);
