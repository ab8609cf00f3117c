// K1: RSS scan over 2-bit packed contigs, both strands, all four signal types in one pass.
//
// Replaces find_valid_motif_idx + find_valid_rss + the per-strand loop of get_contigwise_rss
// (py/IGDetective.py:138-207): exact membership of every 7-mer / 9-mer window in the motif sets,
// then heptamer/nonamer pairing at spacer s, s-1, s+1 with the reference's priority.
//
// Formulation: the reverse strand is scanned in forward coordinates with
// reverse-complemented motif tables, and pairing is gated on the heptamer: every RSS of every
// type contains a heptamer hit q, so 9-mer tables are consulted only next to heptamer hits.
//   strand +, heptamer first (V, D_right):  nonamer at q+7+s'                (first s' that hits)
//   strand -, heptamer first             :  nonamer at q-s'-9                (first s' that hits)
//   strand +, nonamer first (J, D_left)  :  nonamer p=q-9-s' selects q iff no better-ranked
//                                           spacer of p finds a heptamer (q+1 | q-1,q-2)
//   strand -, nonamer first              :  nonamer n=q+7+s' selects q iff (q-1 | q+1,q+2) miss
// with s' tried in the order s, s-1, s+1 (:162-168).
//
// Execution: one persistent CTA of 24 warps per SM.  Every warp streams its own strips (4 rows of 32
// chunks = 8192 bases; single rows for the last ~2 rows per warp so that the SMs finish together)
// through a private double buffer: its own TMA bulk copies (cp.async.bulk + mbarrier), one halo chunk
// per side (64 bases >= the 40-base RSS span).  Strips are handed out in batches: two static batches
// per warp, then a global atomic counter asked one batch ahead -- no CTA-wide coupling between the
// prologue and the final key flush, every base is read from HBM once (0.25 B/base).  The motif
// tables (two more bulk copies) are shared by the SM's 24 warps.
// A cascade of ever rarer, ever more expensive steps, each one run with full lanes on a queue:
//   1. bit-parallel prefilter, one 64-base chunk per lane: every heptamer of the shipped motif file
//      (both strands) agrees with the consensus CAC.GTG in at least 4 of its 6 fixed positions; a
//      bit-sliced population count thresholded at 4 passes every possible hit and ~3.8 % of random
//      positions (igd_set_motifs verifies the property for the tables it is given; otherwise the
//      exhaustive instance, where every position is a candidate, runs).  Survivors are queued.
//   2. 64 survivors at a time (two per lane): the window-need byte of the 7-mer (need7, 16 KB, zero = not a
//      heptamer of any set) -- real heptamer hits (~1 %) are queued with their need bits.
//   3. 32 hits at a time: for every nonamer window the hit's types need, ONE probe of the "centre"
//      bitmap any9c (32 KB): the three candidate nonamers (spacer s-1, s, s+1) start at x, x+1,
//      x+2, and bit c of any9c is set iff the 9-mer c or a 9-mer overlapping it shifted by one base
//      is in some nonamer set, so the 9-mer at x+1 rules the window out for ~95 % of the pairs; one
//      32-position fetch serves the centres of all windows on one side of the heptamer.
//   4. what passes (~8e-4 per base) goes to a queue of GLOBAL positions that survives across strips
//      and is resolved exactly -- nonamer flags, validity plane, the reference's ranking -- from L2.
#include "common.cuh"

namespace {

#ifndef SCAN_WARPS_N
#define SCAN_WARPS_N 24
#endif
constexpr int V5_WARPS = SCAN_WARPS_N;
constexpr int V5_THREADS = V5_WARPS * 32;
#ifndef SCAN_MAXROWS
#define SCAN_MAXROWS 4
#endif
constexpr int V5_MAXROWS = SCAN_MAXROWS;
constexpr int V5_BUF_STRIDE = ((32 * V5_MAXROWS + 2) * 16 + 127) / 128 * 128;   // 2080 -> 2176 for 4 rows
constexpr int V5_Q1CAP = 256;                                   // survivors (16-bit stage positions)
constexpr int V5_Q2CAP = 128;                                   // heptamer hits (position | need << 16)
constexpr int V5_EBCAP = 64;                                    // found RSS keys waiting for one global atomic
constexpr int V5_SQCAP = 64;                                    // (global position << 8 | windows) for the exact path

struct ScanParams {
    const uint4 *planes;           // records {lo0, hi0, lo1, hi1}
    const unsigned long long *inv;
    const uint32_t *invsum;
    const uint8_t *lut7;
    const uint8_t *lut9;
    const uint8_t *need7;          // 7-mer -> bitmask of the windows that serve any of its flags
    const uint32_t *any9c;
    unsigned long long n_chunks;
    unsigned n_strips;
    unsigned batch;                // consecutive strips handed out together: 1, 2 or 4
    unsigned n_long;               // strips [0, n_long) hold V5_MAXROWS rows (32 chunks each), the rest one row (fine-grained tail)
    int nwin;                      // distinct (side, spacer) nonamer windows, <= 8, ascending offset
    int woff[8];                   // first candidate nonamer start relative to the heptamer start
    unsigned wmask[8];             // heptamer flag bits whose nonamers live in that window
    int wnew[8];                   // window opens a new fetch group
    int wbase[8];                  // fetch start of the window's group relative to the heptamer start
    unsigned wgrp[8];              // windows of the group (bitmask over w)
    int wsh[8];                    // position of the window's centre 9-mer inside the group's fetch
    unsigned long long *hits;
    unsigned long long cap;
    unsigned long long *count;     // [0] hits, [1] strip counter
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Shared-memory reads of the hot loops take 32-bit shared-window addresses, so that a table or stage base
// is one register and the access one instruction.  The stage is written by the async proxy: its loads are
// volatile asm, ordered after the volatile mbarrier wait; the tables are constant after the prologue.
__device__ __forceinline__ uint2 lds_stage64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_tab8(uint32_t addr)
{
    unsigned v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_tab32(uint32_t addr)
{
    unsigned v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// 32 consecutive positions of both bit-planes starting at stage position p: the (lo, hi) words of 32
// bases are one aligned 8-byte pair, so a fetch is two LDS.64 off one address and two funnel shifts
__device__ __forceinline__ void swin32x2(uint32_t sp, int p, unsigned &lo, unsigned &hi)
{
    const uint32_t at = sp + ((unsigned)(p >> 5) << 3);
    const uint2 a = lds_stage64(at), b = lds_stage64(at + 8);
    lo = __funnelshift_r(a.x, b.x, p);
    hi = __funnelshift_r(a.y, b.y, p);
}

// the same from global memory at global base g (rare path, served by L2)
__device__ __forceinline__ void gwin32x2(const ScanParams &P, unsigned long long g, unsigned &lo, unsigned &hi)
{
    const uint2 *gp = (const uint2 *)P.planes;
    const unsigned long long i = g >> 5, last = 2 * P.n_chunks - 1;
    const uint2 a = __ldg(&gp[i]);
    const uint2 b = __ldg(&gp[i + 1 <= last ? i + 1 : last]);
    lo = __funnelshift_r(a.x, b.x, (unsigned)g);
    hi = __funnelshift_r(a.y, b.y, (unsigned)g);
}

// true iff no base of [g, g+k) is marked invalid (global coordinates; rare path)
__device__ __forceinline__ bool window_valid(const ScanParams &P, unsigned long long g, int k)
{
    const unsigned long long c0 = g >> 6, c1 = (g + k - 1) >> 6;
    const int b0 = (int)(g & 63);
    if ((__ldg(&P.invsum[c0 >> 5]) >> (c0 & 31)) & 1u) {
        const unsigned long long m = __ldg(&P.inv[c0]);
        const int n0 = (b0 + k > 64) ? 64 - b0 : k;
        if (m & (((1ull << n0) - 1ull) << b0)) return false;
    }
    if (c1 != c0 && ((__ldg(&P.invsum[c1 >> 5]) >> (c1 & 31)) & 1u)) {
        const unsigned long long m = __ldg(&P.inv[c1]);
        if (m & ((1ull << (b0 + k - 64)) - 1ull)) return false;
    }
    return true;
}

// heptamer flags at global position g (0 if the window holds an invalid base)
__device__ __forceinline__ unsigned hept_flags(const ScanParams &P, unsigned long long g)
{
    unsigned lo, hi;
    gwin32x2(P, g, lo, hi);
    const unsigned f = __ldg(&P.lut7[((hi & 127u) << 7) | (lo & 127u)]);
    if (!f) return 0;
    return window_valid(P, g, 7) ? f : 0;
}

// Shared memory: one private region per warp (two stage buffers, the three queues, found RSS keys, two mbarriers), so that
// everything a warp touches sits at a constant offset from one base; then the tables shared by the SM.
constexpr int WR_BUF = 0;                                       // 2 x V5_BUF_STRIDE
constexpr int WR_Q1 = 2 * V5_BUF_STRIDE;                        // V5_Q1CAP x u16
constexpr int WR_Q2 = WR_Q1 + V5_Q1CAP * 2;                     // V5_Q2CAP x u32
constexpr int WR_SQ = WR_Q2 + V5_Q2CAP * 4;                     // V5_SQCAP x u64
constexpr int WR_EB = WR_SQ + V5_SQCAP * 8;                     // V5_EBCAP x u64: RSS keys waiting for one global atomic
constexpr int WR_BAR = WR_EB + V5_EBCAP * 8;                    // 2 x u64 mbarriers, then the u32 fill of the key buffer
constexpr int WR_EBN = WR_BAR + 16;
constexpr int WR_SIZE = (WR_EBN + 4 + 127) / 128 * 128;
constexpr size_t kScanSmem = (size_t)V5_WARPS * WR_SIZE + 16384 + 32768 + 16;   // + the tables' mbarrier

// An RSS is found ~6e-5 times per base: keys collect in the warp's buffer (shared-memory atomic) and leave with
// one global atomic per flush instead of one round trip to L2 per key.
__device__ __forceinline__ void emit(const ScanParams &P, unsigned char *wb, unsigned long long gq, int t, int strand, int d)
{
    const unsigned long long key = (gq << 8) | ((unsigned long long)t << 4) | ((unsigned long long)strand << 3) | (unsigned long long)d;
    const unsigned i = atomicAdd((unsigned *)(wb + WR_EBN), 1u);
    if (i < (unsigned)V5_EBCAP) ((unsigned long long *)(wb + WR_EB))[i] = key;
    else {                                                       // buffer full inside one round: straight to global memory
        const unsigned long long j = atomicAdd(P.count, 1ull);
        if (j < P.cap) P.hits[j] = key;
    }
}

// all lanes of the warp; empties the key buffer
__device__ __forceinline__ void flush_keys(const ScanParams &P, unsigned char *wb, unsigned lane)
{
    __syncwarp();
    const unsigned n = min(*(volatile unsigned *)(wb + WR_EBN), (unsigned)V5_EBCAP);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.count, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned long long *eb = (const unsigned long long *)(wb + WR_EB);
    for (unsigned i = lane; i < n; i += 32)
        if (base + i < P.cap) P.hits[base + i] = eb[i];
    __syncwarp();
    if (lane == 0) *(volatile unsigned *)(wb + WR_EBN) = 0;
    __syncwarp();
}

// Step 4, exact: item = heptamer position << 8 | windows whose centre probe passed.  Nonamer flags of the
// three candidate starts of each window, validity, and the reference's ranking (py/IGDetective.py:162-175,
// inverted for the nonamer-first types as in the header of this file).
__device__ __forceinline__ void exact_round(const ScanParams &P, unsigned char *wb, unsigned long long item, bool active)
{
    if (!active) return;
    const unsigned long long gq = item >> 8;
    unsigned pm = (unsigned)item & 0xffu;
    unsigned lo, hi;
    gwin32x2(P, gq, lo, hi);
    const unsigned f = __ldg(&P.lut7[((hi & 127u) << 7) | (lo & 127u)]);
    bool hv_known = false;
    unsigned hm2 = 0, hm1 = 0, hp1 = 0, hp2 = 0; bool have_nb = false;
    while (pm) {
        const int w = __ffs(pm) - 1;
        pm &= pm - 1;
        unsigned fm = f & P.wmask[w];
        const int wo = P.woff[w];
        const unsigned long long x = gq + (long long)wo;
        gwin32x2(P, x, lo, hi);
        unsigned nf0 = __ldg(&P.lut9[((hi & 511u) << 9) | (lo & 511u)]);
        unsigned nf1 = __ldg(&P.lut9[(((hi >> 1) & 511u) << 9) | ((lo >> 1) & 511u)]);
        unsigned nf2 = __ldg(&P.lut9[(((hi >> 2) & 511u) << 9) | ((lo >> 2) & 511u)]);
        if (!((nf0 | nf1 | nf2) & fm)) continue;
        if (!hv_known) { if (!window_valid(P, gq, 7)) return; hv_known = true; }
        if (nf0 && !window_valid(P, x, 9)) nf0 = 0;
        if (nf1 && !window_valid(P, x + 1, 9)) nf1 = 0;
        if (nf2 && !window_valid(P, x + 2, 9)) nf2 = 0;
        const bool right = wo > 0;
        while (fm) {
            const int fb = __ffs(fm) - 1;
            fm &= fm - 1;
            const unsigned bit = 1u << fb;
            const unsigned n = ((nf0 & bit) ? 1u : 0u) | ((nf1 & bit) ? 2u : 0u) | ((nf2 & bit) ? 4u : 0u);
            if (!n) continue;
            const int t = fb & 3, strand = fb >> 2;
            const bool hept_first = (t == IGD_SIG_V || t == IGD_SIG_D_RIGHT);
            // window positions ascend; on the right they are spacers s-1, s, s+1, on the left s+1, s, s-1
            const unsigned bm = right ? 1u : 4u, bp = right ? 4u : 1u;
            if (hept_first) {                                      // the heptamer picks: s, then s-1, then s+1
                emit(P, wb, gq, t, strand, (n & 2u) ? 0 : ((n & bm) ? 1 : 2));
            } else {                                               // each nonamer picks its best-ranked heptamer
                if (!have_nb && (n & 5u)) {
                    hm2 = hept_flags(P, gq - 2); hm1 = hept_flags(P, gq - 1);
                    hp1 = hept_flags(P, gq + 1); hp2 = hept_flags(P, gq + 2);
                    have_nb = true;
                }
                if (n & 2u) emit(P, wb, gq, t, strand, 0);
                if (strand == 0) {      // nonamer p = q-9-s' ranks heptamers p+9+s, +s-1, +s+1
                    if ((n & bm) && !(hp1 & bit)) emit(P, wb, gq, t, 0, 1);
                    if ((n & bp) && !(hm1 & bit) && !(hm2 & bit)) emit(P, wb, gq, t, 0, 2);
                } else {                // reverse strand: nonamer n = q+7+s' ranks heptamers n-s-7, +1, -1
                    if ((n & bm) && !(hm1 & bit)) emit(P, wb, gq, t, 1, 1);
                    if ((n & bp) && !(hp1 & bit) && !(hp2 & bit)) emit(P, wb, gq, t, 1, 2);
                }
            }
        }
    }
}

struct Warp {                      // per-warp state of the cascade (queue fills are warp-uniform)
    unsigned char *wb;             // the warp's shared-memory region
    uint32_t sp;                   // current stage buffer ((lo, hi) word pairs), shared-window address
    uint32_t need7, any9c;         // the SM's tables, shared-window addresses
    unsigned row0;                 // first row of the strip: global base index of stage position 0 = row0 * 2048
    int n1, n2, nsq;
    unsigned lane, lt;             // lane id, mask of lower lanes
    __device__ __forceinline__ uint16_t *q1() const { return (uint16_t *)(wb + WR_Q1); }
    __device__ __forceinline__ unsigned *q2() const { return (unsigned *)(wb + WR_Q2); }
    __device__ __forceinline__ unsigned long long *sq() const { return (unsigned long long *)(wb + WR_SQ); }
};

// one probe of the centre bitmap: 9-mer at bit `sh` of the fetched planes
__device__ __forceinline__ unsigned centre_probe(const Warp &W, unsigned lo, unsigned hi, int sh)
{
    const unsigned tl = lo >> sh, th = (hi >> sh) & 511u;
    const unsigned word = lds_tab32(W.any9c + ((th << 6) | ((tl >> 3) & 60u)));
    return (word >> (tl & 31u)) & 1u;
}

// Step 3: up to 32 heptamer hits (entry = stage position | need << 16; idle lanes pass a valid position
// with need 0).  STD4: four windows in two fetch groups {0, 1} and {2, 3} (any set of spacers with one
// value for V/J and one for D_left/D_right, like the reference's 23/12/12/23).
template <bool STD4>
__device__ __forceinline__ void hits_round(const ScanParams &P, Warp &W, unsigned entry)
{
    const int q = (int)(entry & 0xffffu);
    const unsigned need = entry >> 16;
    unsigned lo = 0, hi = 0, pm = 0;
    if (STD4) {
        swin32x2(W.sp, q + P.wbase[0], lo, hi);
        pm = centre_probe(W, lo, hi, 0) | (centre_probe(W, lo, hi, P.wsh[1]) << 1);
        swin32x2(W.sp, q + P.wbase[2], lo, hi);
        pm |= (centre_probe(W, lo, hi, 0) << 2) | (centre_probe(W, lo, hi, P.wsh[3]) << 3);
        pm &= need;
    } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < P.nwin) {
                if (P.wnew[w]) swin32x2(W.sp, q + P.wbase[w], lo, hi);
                pm |= centre_probe(W, lo, hi, P.wsh[w]) << w;
            }
        }
        pm &= need;
    }
    const unsigned m = __ballot_sync(0xffffffffu, pm != 0);
    if (pm) W.sq()[W.nsq + __popc(m & W.lt)] = (((unsigned long long)W.row0 * (32u * IGD_CHUNK_BASES) + (unsigned)q) << 8) | pm;
    W.nsq += __popc(m);
}

// Step 2: up to 64 prefilter survivors, two per lane (independent chains); needs W.n2 < 32 on entry
__device__ __forceinline__ void survivors_round(Warp &W, int take)
{
    const bool a0 = (int)W.lane < take, a1 = (int)W.lane + 32 < take;
    unsigned q0 = 0, q1 = 0, need0 = 0, need1 = 0;
    if (a0) {
        unsigned lo, hi;
        q0 = W.q1()[W.n1 + W.lane];
        swin32x2(W.sp, (int)q0, lo, hi);
        need0 = lds_tab8(W.need7 + (((hi & 127u) << 7) | (lo & 127u)));
    }
    if (a1) {
        unsigned lo, hi;
        q1 = W.q1()[W.n1 + 32 + W.lane];
        swin32x2(W.sp, (int)q1, lo, hi);
        need1 = lds_tab8(W.need7 + (((hi & 127u) << 7) | (lo & 127u)));
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, need0 != 0);
    const unsigned m1 = __ballot_sync(0xffffffffu, need1 != 0);
    if (need0) W.q2()[W.n2 + __popc(m0 & W.lt)] = q0 | (need0 << 16);
    W.n2 += __popc(m0);
    if (need1) W.q2()[W.n2 + __popc(m1 & W.lt)] = q1 | (need1 << 16);
    W.n2 += __popc(m1);
    __syncwarp();
}

// Runs full rounds while a queue holds enough for one, later stages first (so the hit queue never holds more
// than 95 entries and the exact queue never more than 63); with `flush` the queues of stage positions are
// emptied (end of a strip), with `final` the exact queue too (the warp's last strip).
template <bool STD4>
__device__ __forceinline__ void drain(const ScanParams &P, Warp &W, bool flush, bool final)
{
    for (;;) {
        // full rounds first (three plain comparisons on the hot path); the partial rounds of a flush only
        // when no queue holds a full one
        int stage = W.nsq >= 32 ? 3 : (W.n2 >= 32 ? 2 : (W.n1 >= 64 ? 1 : 0));
        if (stage == 0) {
            if (!flush) break;
            stage = W.n1 > 0 ? 1 : (W.n2 > 0 ? 2 : (final && W.nsq > 0 ? 3 : 0));
            if (stage == 0) break;
        }
        if (stage == 3) {
            const int take = min(W.nsq, 32);
            W.nsq -= take;
            const bool active = (int)W.lane < take;
            exact_round(P, W.wb, W.sq()[W.nsq + (active ? W.lane : 0u)], active);
            __syncwarp();
            if (*(volatile unsigned *)(W.wb + WR_EBN) >= 32u) flush_keys(P, W.wb, W.lane);
        } else if (stage == 2) {
            const int take = min(W.n2, 32);
            W.n2 -= take;
            hits_round<STD4>(P, W, (int)W.lane < take ? W.q2()[W.n2 + W.lane] : 64u);      // 64: any position inside the stage
            __syncwarp();
        } else {
            const int take = min(W.n1, 64);
            W.n1 -= take;
            survivors_round(W, take);
        }
    }
}

// ---- step 1: positions of word (w, next) whose 7-mer matches CAC.GTG in at least 4 of the 6 fixed positions
struct Ind { uint32_t a, c, g, t; };
__device__ __forceinline__ Ind indicators(uint32_t lo, uint32_t hi)
{
    Ind r; r.a = ~(lo | hi); r.c = lo & ~hi; r.g = hi & ~lo; r.t = lo & hi; return r;
}
__device__ __forceinline__ uint32_t consensus_ge4(const Ind &w, const Ind &n)
{
    const uint32_t m0 = w.c;
    const uint32_t m1 = __funnelshift_r(w.a, n.a, 1);
    const uint32_t m2 = __funnelshift_r(w.c, n.c, 2);
    const uint32_t m4 = __funnelshift_r(w.g, n.g, 4);
    const uint32_t m5 = __funnelshift_r(w.t, n.t, 5);
    const uint32_t m6 = __funnelshift_r(w.g, n.g, 6);
    const uint32_t s1 = m0 ^ m1 ^ m2, c1 = (m0 & m1) | (m2 & (m0 | m1));
    const uint32_t s2 = m4 ^ m5 ^ m6, c2 = (m4 & m5) | (m6 & (m4 | m5));
    const uint32_t c3 = s1 & s2;
    return (c1 & c2) | (c3 & (c1 | c2));          // ones + 2*twos >= 4  <=>  at least two carries
}

__device__ __forceinline__ int top_bit(unsigned m)
{
    int b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));
    return b;
}

template <bool PREFILTER, bool STD4>
__global__ void __launch_bounds__(V5_THREADS, 1)
rss_scan_kernel(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint8_t *s_need7 = smem + V5_WARPS * WR_SIZE;                            // 16384
    uint32_t *s_any9c = (uint32_t *)(s_need7 + 16384);                       // 32768

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef SCAN_TIMELINE
    unsigned long long tl0, tl1 = 0, tl2 = 0; unsigned tl_strips = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif
    Warp W;
    W.wb = smem + warp * WR_SIZE;
    uint64_t *full = (uint64_t *)(W.wb + WR_BAR);
    if (lane == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        *(volatile unsigned *)(W.wb + WR_EBN) = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // strip s: V5_MAXROWS rows from row s * V5_MAXROWS while s < n_long, then single rows
    const auto strip_row0 = [&](unsigned s) { return s < P.n_long ? s * V5_MAXROWS : P.n_long * V5_MAXROWS + (s - P.n_long); };
    const auto strip_rows = [&](unsigned s) { return s < P.n_long ? (unsigned)V5_MAXROWS : 1u; };
    // Strips are handed out in batches of P.batch consecutive strips: every warp owns two batches from the start
    // (no rush on the counter when the kernel starts), later ones come from a global atomic counter.  With batches
    // of several strips (large inputs) the answer is asked for two batches ahead; with single strips (small inputs)
    // one strip ahead only, because every reserved strip is one a warp may still have to work through alone when
    // the others have run out (seen as a tail of 8 us on 100 Mb).  All of this is warp-uniform except `nn` (lane 0).
    unsigned *const batch_counter = (unsigned *)(P.count + 1);
    const unsigned nwarps = gridDim.x * V5_WARPS, gw = blockIdx.x * V5_WARPS + warp;
    unsigned s_cur = gw * P.batch;                                       // current strip
    unsigned s_end = min(s_cur + P.batch, P.n_strips);                   // end of its batch
    const bool deep = P.batch > 1;
    unsigned nb = gw + nwarps;                                           // next batch (deep hand-out)
    unsigned nn = gw + nwarps;                                           // lane 0: the batch after that / the next one, may be in flight
    if (lane == 0) {
        if (deep) nn = 2u * nwarps + atomicAdd(batch_counter, 1u);
        if (s_cur < P.n_strips) {
            const unsigned bytes = (strip_rows(s_cur) * 32u + 2u) * 16u;
            mbar_expect_tx(&full[0], bytes);
            bulk_g2s(W.wb + WR_BUF, P.planes + (size_t)strip_row0(s_cur) * 32u, bytes, &full[0]);
        }
    }
    // motif tables -> shared memory: two bulk copies on a CTA-wide mbarrier (L2-resident after the first CTA)
    uint64_t *tbar = (uint64_t *)(s_any9c + 8192);
    if (tid == 0) {
        mbar_init(tbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(tbar, 16384 + 32768);
        bulk_g2s(s_need7, P.need7, 16384, tbar);
        bulk_g2s(s_any9c, P.any9c, 32768, tbar);
    }
    __syncthreads();
    while (!mbar_try_wait(tbar, 0)) { }

    W.need7 = smem_u32(s_need7); W.any9c = smem_u32(s_any9c);
    W.n1 = 0; W.n2 = 0; W.nsq = 0;
    W.lane = (unsigned)lane; W.lt = (1u << lane) - 1u;
    for (unsigned it = 0; s_cur < P.n_strips; ++it) {
        const bool switch_batch = s_cur + 1 >= s_end;
        unsigned s_nxt = s_cur + 1;
        if (switch_batch) {
            const unsigned b = deep ? nb : __shfl_sync(0xffffffffu, nn, 0);
            s_nxt = b < 0x40000000u ? b * P.batch : 0xffffffffu;
            if (!deep && lane == 0) nn = 2u * nwarps + atomicAdd(batch_counter, 1u);
        }
        if (lane == 0 && s_nxt < P.n_strips) {
            // the other buffer was read during iteration it-1 (all lanes passed the __syncwarp at its end)
            const unsigned bytes = (strip_rows(s_nxt) * 32u + 2u) * 16u;
            mbar_expect_tx(&full[(it + 1) & 1u], bytes);
            bulk_g2s(W.wb + WR_BUF + ((it + 1) & 1u) * V5_BUF_STRIDE, P.planes + (size_t)strip_row0(s_nxt) * 32u, bytes, &full[(it + 1) & 1u]);
        }
        __syncwarp();
        while (!mbar_try_wait(&full[it & 1u], (it >> 1) & 1u)) { }
#ifdef SCAN_TIMELINE
        if (it == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
        ++tl_strips;
#endif

        const uint4 *rec = (const uint4 *)(W.wb + WR_BUF + (it & 1u) * V5_BUF_STRIDE);
        W.sp = smem_u32(rec);
        W.row0 = strip_row0(s_cur);                                            // record 0 = chunk row0 * 32
        const unsigned rows = strip_rows(s_cur);
        const bool last_strip = s_nxt >= P.n_strips;
        for (unsigned r0 = 0; r0 < rows; r0 += 2) {
            // ---- step 1: candidate masks of the 2 x 64 windows starting in this lane's chunk of two rows
            uint32_t cm[4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                cm[2 * k] = 0; cm[2 * k + 1] = 0;
                if (r0 + k < rows) {
                    cm[2 * k] = 0xffffffffu; cm[2 * k + 1] = 0xffffffffu;
                    if (PREFILTER) {
                        const int c = (int)(r0 + k) * 32 + lane + 1;         // stage record of the chunk
                        const uint4 r = rec[c];
                        const uint2 n = *(const uint2 *)&rec[c + 1];
                        const Ind i0 = indicators(r.x, r.y), i1 = indicators(r.z, r.w), i2 = indicators(n.x, n.y);
                        cm[2 * k] = consensus_ge4(i0, i1);
                        cm[2 * k + 1] = consensus_ge4(i1, i2);
                    }
                }
            }
            // ---- survivors into the warp queue, rounds as they form
            const unsigned pos00 = (r0 * 32 + lane + 1) * 64;                  // stage position of bit 0 of cm[0]
            if (!PREFILTER) {
                // exhaustive instance: every position is a candidate, so no enumeration -- two positions per lane
                // and step go straight into the survivor queue (64 entries: one round of the heptamer test)
                const unsigned nrow = min(rows - r0, 2u);
                for (unsigned cur = 0; cur < nrow * 64u; cur += 2) {
                    const unsigned pos = pos00 + (cur >> 6) * 2048u + (cur & 63u);
                    W.q1()[W.n1 + lane] = (uint16_t)pos;
                    W.q1()[W.n1 + 32 + lane] = (uint16_t)(pos + 1);
                    W.n1 += 64;
                    __syncwarp();
                    const bool flush = cur + 2 >= nrow * 64u && r0 + 2 >= rows;
                    drain<STD4>(P, W, flush, flush && last_strip);
                }
                continue;
            }
            for (bool more = true; more;) {
                const int cnt = __popc(cm[0]) + __popc(cm[1]) + __popc(cm[2]) + __popc(cm[3]);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                if (tot <= V5_Q1CAP - W.n1) {                                  // the usual case: everything fits
                    uint16_t *dst = W.q1() + W.n1 + (incl - cnt);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned pos0 = pos00 + (k >> 1) * 2048 + (k & 1) * 32;
                        unsigned m = cm[k];
                        while (m) {
                            const int b = top_bit(m);
                            m ^= 1u << b;
                            *dst++ = (uint16_t)(pos0 + b);
                        }
                    }
                    W.n1 += tot;
                    more = false;
                } else {
                    // pathological density (or the exhaustive instance): one candidate per lane and pass
                    unsigned pos = 0;
                    bool have = true;
                    if (cm[0]) { const int b = top_bit(cm[0]); cm[0] ^= 1u << b; pos = pos00 + b; }
                    else if (cm[1]) { const int b = top_bit(cm[1]); cm[1] ^= 1u << b; pos = pos00 + 32 + b; }
                    else if (cm[2]) { const int b = top_bit(cm[2]); cm[2] ^= 1u << b; pos = pos00 + 2048 + b; }
                    else if (cm[3]) { const int b = top_bit(cm[3]); cm[3] ^= 1u << b; pos = pos00 + 2080 + b; }
                    else have = false;
                    const unsigned mv = __ballot_sync(0xffffffffu, have);
                    if (have) W.q1()[W.n1 + __popc(mv & W.lt)] = (uint16_t)pos;
                    W.n1 += __popc(mv);
                }
                __syncwarp();
                // the queues hold stage positions: they are emptied before the buffer is handed back
                const bool flush = !more && r0 + 2 >= rows;
                drain<STD4>(P, W, flush, flush && last_strip);
            }
        }
        __syncwarp();
        if (switch_batch) {
            s_end = s_nxt < P.n_strips ? min(s_nxt + P.batch, P.n_strips) : 0u;
            if (deep) {
                nb = __shfl_sync(0xffffffffu, nn, 0);
                if (lane == 0) nn = 2u * nwarps + atomicAdd(batch_counter, 1u);
            }
        }
        s_cur = s_nxt;
    }
#ifdef SCAN_TIMELINE
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl2));
    if (lane == 0) printf("TL %u %d %llu %llu %llu %u\n", blockIdx.x, warp, tl0, tl1, tl2, tl_strips);
#endif
    // what is left of the found keys leaves with ONE global atomic per CTA
    __syncthreads();
    unsigned *s_base = (unsigned *)s_any9c;                                  // the tables are no longer needed
    unsigned long long *s_gbase = (unsigned long long *)(s_base + 32);
    if (warp == 0) {
        unsigned n = lane < V5_WARPS ? min(*(volatile unsigned *)(smem + lane * WR_SIZE + WR_EBN), (unsigned)V5_EBCAP) : 0u;
        unsigned incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        s_base[lane] = incl - n;
        const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) *s_gbase = tot ? atomicAdd(P.count, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    {
        const unsigned n = min(*(volatile unsigned *)(W.wb + WR_EBN), (unsigned)V5_EBCAP);
        const unsigned long long base = *s_gbase + s_base[warp];
        const unsigned long long *eb = (const unsigned long long *)(W.wb + WR_EB);
        for (unsigned i = lane; i < n; i += 32)
            if (base + i < P.cap) P.hits[base + i] = eb[i];
    }
}

// ---- K1a: plain k-mer membership (find_valid_motif_idx for all sets of one k at once) ----------
struct KmerParams {
    const uint4 *planes;
    const unsigned long long *inv;
    const uint8_t *lut;
    const uint32_t *any;         // bitmap prefilter (k = 9) or nullptr
    unsigned long long n_chunks;
    int k;
    unsigned long long *hits;
    unsigned long long cap;
    unsigned long long *count;
};

__global__ void __launch_bounds__(256)
kmer_scan_kernel(const KmerParams P)
{
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned kmask = (1u << P.k) - 1u;
    for (unsigned long long c = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; c + 1 < P.n_chunks; c += nthreads) {
        const unsigned long long m0 = __ldg(&P.inv[c]);
        if (m0 == ~0ull) continue;                                   // pad chunk
        const uint4 r = __ldg(&P.planes[c]);
        const uint4 n = __ldg(&P.planes[c + 1]);
        const unsigned long long m1 = __ldg(&P.inv[c + 1]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t l0 = half ? r.z : r.x, l1 = half ? n.x : r.z;
            const uint32_t h0 = half ? r.w : r.y, h1 = half ? n.y : r.w;
            const uint32_t v0 = half ? (uint32_t)(m0 >> 32) : (uint32_t)m0;
            const uint32_t v1 = half ? (uint32_t)m1 : (uint32_t)(m0 >> 32);
            for (int b = 0; b < 32; ++b) {
                if (__funnelshift_r(v0, v1, b) & kmask) continue;   // invalid base in the window
                const unsigned lo = __funnelshift_r(l0, l1, b) & kmask;
                const unsigned hi = __funnelshift_r(h0, h1, b) & kmask;
                const unsigned idx = (hi << P.k) | lo;
                if (P.any && !((__ldg(&P.any[idx >> 5]) >> (idx & 31)) & 1u)) continue;
                const unsigned f = __ldg(&P.lut[idx]);
                if (f) {
                    const unsigned long long i = atomicAdd(P.count, 1ull);
                    if (i < P.cap) P.hits[i] = ((c * 64ull + half * 32 + b) << 8) | f;
                }
            }
        }
    }
}

#ifndef SCAN_TAIL_ROWS_PER_WARP
#define SCAN_TAIL_ROWS_PER_WARP 2
#endif

} // namespace

int igd_launch_rss_scan(igd_handle *h, const int32_t spacer[4], bool prepare_only)
{
    ScanParams P;
    P.planes = h->d_planes; P.inv = h->d_inv; P.invsum = h->d_invsum;
    P.lut7 = h->d_lut7; P.lut9 = h->d_lut9; P.need7 = h->d_need7; P.any9c = h->d_any9c;
    P.n_chunks = h->n_chunks;
    // one window per distinct (side, spacer): right of the heptamer for (heptamer-first, +) and
    // (nonamer-first, -), left of it for the other two combinations; kept in ascending offset order
    P.nwin = 0;
    for (int side = 1; side >= 0; --side)
        for (int t = 0; t < 4; ++t) {
            const bool hept_first = (t == IGD_SIG_V || t == IGD_SIG_D_RIGHT);
            const unsigned bit = side == 0 ? (hept_first ? (1u << t) : (16u << t)) : (hept_first ? (16u << t) : (1u << t));
            const int off = side == 0 ? 7 + spacer[t] - 1 : -9 - spacer[t] - 1;
            int w = 0;
            while (w < P.nwin && P.woff[w] != off) ++w;
            if (w == P.nwin) {
                int at = P.nwin;                                   // insertion sort by offset
                while (at > 0 && P.woff[at - 1] > off) { P.woff[at] = P.woff[at - 1]; P.wmask[at] = P.wmask[at - 1]; --at; }
                P.woff[at] = off; P.wmask[at] = 0; P.nwin++;
                w = at;
            }
            P.wmask[w] |= bit;
        }
    for (int w = P.nwin; w < 8; ++w) { P.woff[w] = 0; P.wmask[w] = 0; }
    // fetch groups: one 32-position fetch serves every window whose centre 9-mer (start woff + 1) lies inside it
    int g0 = 0;
    for (int w = 0; w < 8; ++w) { P.wnew[w] = 0; P.wbase[w] = 0; P.wgrp[w] = 0; P.wsh[w] = 0; }
    for (int w = 0; w < P.nwin; ++w) {
        const int cs = P.woff[w] + 1;
        if (w == 0 || cs - (P.woff[g0] + 1) > 23) { g0 = w; P.wnew[w] = 1; }
        P.wbase[w] = P.woff[g0] + 1;
        P.wsh[w] = cs - P.wbase[w];
        P.wgrp[g0] |= 1u << w;
    }
    // need7: which windows serve a 7-mer's flags (zero = in no heptamer set); rebuilt when the windows change
    if (!h->need7_valid || memcmp(h->need7_spacer, spacer, sizeof h->need7_spacer) != 0) {
        std::vector<uint8_t> need7(16384);
        for (int i = 0; i < 16384; ++i) {
            unsigned need = 0;
            for (int w = 0; w < P.nwin; ++w) need |= (h->h_lut7[i] & P.wmask[w]) ? (1u << w) : 0u;
            need7[i] = (uint8_t)need;
        }
        IGD_CUDA(cudaMemcpyAsync(h->d_need7, need7.data(), 16384, cudaMemcpyHostToDevice, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(h->need7_spacer, spacer, sizeof h->need7_spacer);
        h->need7_valid = true;
    }
    // long strips first, then a tail of single rows (two per warp) so that the SMs finish together
    const unsigned long long rows_total = h->n_tiles * (unsigned long long)(IGD_TILE_CHUNKS / 32);
    const unsigned long long warps = (unsigned long long)h->sm_count * V5_WARPS;
    const unsigned long long tail_rows = std::min<unsigned long long>(rows_total, SCAN_TAIL_ROWS_PER_WARP * warps);
    const unsigned long long n_long = (rows_total - tail_rows) / V5_MAXROWS;
    const unsigned long long n_strips = n_long + (rows_total - n_long * V5_MAXROWS);
    if (n_strips > 0xffffff00ull || rows_total > 0xffffff00ull) { igd_set_error("too many chunks for one scan"); return IGD_ERR_CAPACITY; }
    P.n_long = (unsigned)n_long;
    P.n_strips = (unsigned)n_strips;
    P.batch = n_strips >= 32 * warps ? 4u : (n_strips >= 16 * warps ? 2u : 1u);
    P.hits = h->d_hits; P.cap = h->hit_cap; P.count = h->d_count;
    const bool std4 = P.nwin == 4 && P.wnew[0] && !P.wnew[1] && P.wnew[2] && !P.wnew[3];
    auto kernel = h->consensus_prefilter ? (std4 ? rss_scan_kernel<true, true> : rss_scan_kernel<true, false>)
                                         : (std4 ? rss_scan_kernel<false, true> : rss_scan_kernel<false, false>);
    const int inst = (h->consensus_prefilter ? 2 : 0) + (std4 ? 1 : 0);
    if (!h->scan_attr_set[inst]) {                     // once per handle and kernel instance (the attribute is per device)
        IGD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
        h->scan_attr_set[inst] = true;
    }
    if (prepare_only) return IGD_OK;                 // tables uploaded, kernel instance loaded and configured
    unsigned long long blocks = (unsigned long long)h->sm_count;
    const unsigned long long wanted = (P.n_strips + V5_WARPS - 1) / V5_WARPS;
    if (blocks > wanted) blocks = wanted;
    if (blocks < 1) blocks = 1;
    kernel<<<(unsigned)blocks, V5_THREADS, kScanSmem, h->stream>>>(P);
    h->timing.total_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

int igd_launch_kmer_scan(igd_handle *h, int k)
{
    KmerParams P;
    P.planes = h->d_planes; P.inv = h->d_inv;
    P.lut = (k == 7) ? h->d_lut7 : h->d_lut9;
    P.any = (k == 7) ? nullptr : h->d_any9;
    P.n_chunks = h->n_chunks; P.k = k;
    P.hits = h->d_hits; P.cap = h->hit_cap; P.count = h->d_count;
    unsigned long long blocks = (h->n_chunks + 255) / 256;
    const unsigned long long cap = (unsigned long long)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kmer_scan_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(P);
    h->timing.total_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}
