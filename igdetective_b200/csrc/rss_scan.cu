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
// Data movement: a persistent CTA (8 warps) walks tiles of 512 chunks (32 Kbases, 8 KB of
// planes); each tile plus one halo chunk per side arrives in shared memory by one TMA bulk copy
// (cp.async.bulk + mbarrier; 3 stages, full/empty barrier pairs), so every base is read from
// HBM once (0.25 B/base).  There is no CTA-wide barrier in the loop: each warp owns two rows of
// 32 chunks per tile and releases the stage as soon as it has read them.
// Phase A, one 64-base chunk per lane, is bit-parallel on the planes: every heptamer of the
// shipped motif file (both strands) agrees with the consensus CAC.GTG in at least 4 of its 6
// fixed positions, so a bit-sliced population count of the six letter-match planes, thresholded
// at 4, passes every possible heptamer hit and only ~3.8 % of random positions (igd_set_motifs
// verifies that property for the tables it is given; tables that break it use the exhaustive
// instance of the same kernel).  Survivors are looked up in the exact 7-mer flag table in shared
// memory; real heptamer hits (~1 % of positions) are compacted into a per-warp queue.
// Phase B runs on the staged tile too (halo = 64 bases >= the 40-base RSS span): hits are expanded
// into (heptamer, window) items, 32 items are probed at a time against the 32 KB bitmap of all 9-mers
// (L1-resident; SCAN_ANY9_SMEM=1 keeps it in shared memory at the price of one CTA per SM less), and
// only the ~3 % of items that find a 9-mer touch the exact tables and the validity plane.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int ROWS_PER_WARP = IGD_TILE_CHUNKS / 32 / SCAN_WARPS;      // 2
#ifndef SCAN_ANY9_SMEM
#define SCAN_ANY9_SMEM 0
#endif
#ifndef SCAN_MIN_CTAS
#define SCAN_MIN_CTAS 3
#endif
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = IGD_STAGE_CHUNKS * 16;          // 8224
constexpr int STAGE_STRIDE = 8320;                          // 128-byte aligned
constexpr int WQCAP = 128;                                  // per-warp hit queue (flushes in rounds of 32)
constexpr int IQCAP = 288;                                  // per-warp (hit, window) items: < 32 waiting + 32 hits x 8 windows

struct ScanParams {
    const uint4 *planes;
    const unsigned long long *inv;
    const uint32_t *invsum;
    const uint8_t *lut7;
    const uint8_t *lut9;
    const uint32_t *any9;
    unsigned long long n_tiles;
    int sp[4];
    int nwin;                      // distinct (side, spacer) nonamer windows, <= 8
    int woff[8];                   // first candidate nonamer start relative to the heptamer start
    unsigned wmask[8];             // heptamer flag bits whose nonamers live in that window
    unsigned long long *hits;
    unsigned long long cap;
    unsigned long long *count;
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// ---- phase B helpers ----------------------------------------------------------------------------
struct Stage {
    const uint32_t *sw;            // stage words: record r = {lo0, lo1, hi0, hi1} at sw[4r..4r+3]
    const uint8_t *lut7;           // shared-memory copies of the motif tables
    const uint8_t *need;           // 7-mer flags -> bitmask of the windows that serve any of them
    const uint32_t *any9;
    unsigned long long g0;         // global base index of stage position 0
};

// 32 consecutive positions of both bit-planes starting at stage position p
__device__ __forceinline__ void swin32x2(const uint32_t *sw, int p, unsigned &lo, unsigned &hi)
{
    const int b0 = p >> 5, b1 = b0 + 1;
    const int w0 = 4 * (b0 >> 1) + (b0 & 1), w1 = 4 * (b1 >> 1) + (b1 & 1);
    lo = __funnelshift_r(sw[w0], sw[w1], p & 31);
    hi = __funnelshift_r(sw[w0 + 2], sw[w1 + 2], p & 31);
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

// heptamer flags at stage position p (0 if the window holds an invalid base)
__device__ __forceinline__ unsigned hept_flags(const ScanParams &P, const Stage &T, int p)
{
    unsigned lo, hi;
    swin32x2(T.sw, p, lo, hi);
    const unsigned f = T.lut7[((hi & 127u) << 7) | (lo & 127u)];
    if (!f) return 0;
    return window_valid(P, T.g0 + p, 7) ? f : 0;
}

__device__ __forceinline__ void emit(const ScanParams &P, unsigned long long gq, int t, int strand, int d)
{
    const unsigned long long i = atomicAdd(P.count, 1ull);
    if (i < P.cap) P.hits[i] = (gq << 8) | ((unsigned long long)t << 4) | ((unsigned long long)strand << 3) | (unsigned long long)d;
}

// Rare path: window w of the heptamer at stage position q has at least one nonamer carrying a wanted
// flag.  nf[i] = 9-mer flags at window position i (ascending).  Applies the reference's ranking.
__device__ __noinline__ void resolve_window(const ScanParams &P, const Stage &T, int q,
                                            unsigned fm, int w, unsigned nf0, unsigned nf1, unsigned nf2)
{
    const unsigned long long gq = T.g0 + q;
    if (!window_valid(P, gq, 7)) return;
    const unsigned long long x = gq + (long long)P.woff[w];
    if (nf0 && !window_valid(P, x, 9)) nf0 = 0;
    if (nf1 && !window_valid(P, x + 1, 9)) nf1 = 0;
    if (nf2 && !window_valid(P, x + 2, 9)) nf2 = 0;
    const bool right = P.woff[w] > 0;
    unsigned hm2 = 0, hm1 = 0, hp1 = 0, hp2 = 0; bool have_nb = false;
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
            emit(P, gq, t, strand, (n & 2u) ? 0 : ((n & bm) ? 1 : 2));
        } else {                                               // each nonamer picks its best-ranked heptamer
            if (!have_nb && (n & 5u)) {
                hm2 = hept_flags(P, T, q - 2); hm1 = hept_flags(P, T, q - 1);
                hp1 = hept_flags(P, T, q + 1); hp2 = hept_flags(P, T, q + 2);
                have_nb = true;
            }
            if (n & 2u) emit(P, gq, t, strand, 0);
            if (strand == 0) {      // nonamer p = q-9-s' ranks heptamers p+9+s, +s-1, +s+1
                if ((n & bm) && !(hp1 & bit)) emit(P, gq, t, 0, 1);
                if ((n & bp) && !(hm1 & bit) && !(hm2 & bit)) emit(P, gq, t, 0, 2);
            } else {                // reverse strand: nonamer n = q+7+s' ranks heptamers n-s-7, +1, -1
                if ((n & bm) && !(hm1 & bit)) emit(P, gq, t, 1, 1);
                if ((n & bp) && !(hp1 & bit) && !(hp2 & bit)) emit(P, gq, t, 1, 2);
            }
        }
    }
}

// One (heptamer, window) work item = stage position | window << 16 | served flag bits << 24: probe the
// three adjacent 9-mers of that window in the shared-memory bitmap; the exact flags (L2) and the
// validity plane are touched only when the bitmap says a 9-mer of some set is there (~3 % of items).
__device__ __forceinline__ void probe_window(const ScanParams &P, const Stage &T, unsigned item)
{
    const int q = (int)(item & 0xffffu);
    const int w = (int)(item >> 16) & 7;
    const unsigned fm = (item >> 24) & P.wmask[w];
    unsigned lo, hi;
    swin32x2(T.sw, q + P.woff[w], lo, hi);
    unsigned nf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const unsigned idx = (((hi >> i) & 511u) << 9) | ((lo >> i) & 511u);
        nf[i] = 0;
        if ((SCAN_ANY9_SMEM ? T.any9[idx >> 5] : __ldg(&T.any9[idx >> 5])) >> (idx & 31) & 1u) nf[i] = __ldg(&P.lut9[idx]);
    }
    if ((nf[0] | nf[1] | nf[2]) & fm) resolve_window(P, T, q, fm, w, nf[0], nf[1], nf[2]);
}

// Phase B for one round of up to 32 queued heptamer hits (lane `active` holds stage position q): look
// up the heptamer's flags, expand into one item per window that serves any of them, and probe items
// 32 at a time so that every lane does useful work.  ni = items waiting in iq (warp-uniform).
__device__ __noinline__ int phase_b(const ScanParams &P, const Stage &T, unsigned *iq, int ni,
                                   int q, bool active, bool drain)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned f = 0, need = 0;
    if (active) {
        unsigned lo, hi;
        swin32x2(T.sw, q, lo, hi);
        f = T.lut7[((hi & 127u) << 7) | (lo & 127u)];
        need = T.need[f];
    }
    const int cnt = __popc(need);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    int off = ni + incl - cnt;
    while (need) {
        const int w = __ffs(need) - 1;
        need &= need - 1;
        iq[off++] = (unsigned)q | ((unsigned)w << 16) | (f << 24);
    }
    ni += tot;
    __syncwarp();
    while (ni >= 32 || (drain && ni > 0)) {
        const int take = min(ni, 32);
        ni -= take;
        if ((int)lane < take) probe_window(P, T, iq[ni + lane]);
        __syncwarp();
    }
    return ni;
}

// ---- phase A helpers ---------------------------------------------------------------------------
// positions of word (w, next) whose 7-mer matches CAC.GTG in at least 4 of the 6 fixed positions
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

// exact heptamer hits among the candidate positions `m` of one 32-position word
__device__ __forceinline__ uint32_t lut_filter(uint32_t m, uint32_t l0, uint32_t l1, uint32_t h0, uint32_t h1,
                                               const uint8_t *s_lut7)
{
    uint32_t hit = 0;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const unsigned lo = __funnelshift_r(l0, l1, b) & 127u;
        const unsigned hi = __funnelshift_r(h0, h1, b) & 127u;
        if (s_lut7[(hi << 7) | lo]) hit |= 1u << b;
    }
    return hit;
}

template <bool PREFILTER>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_MIN_CTAS)
rss_scan_kernel(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_stage = smem;                                           // NSTAGE * STAGE_STRIDE
    uint8_t *s_lut7 = smem + NSTAGE * STAGE_STRIDE;                          // 16384
#if SCAN_ANY9_SMEM
    uint32_t *s_any9 = (uint32_t *)(s_lut7 + 16384);                         // 32768
    unsigned *s_wq = s_any9 + 8192;                                          // SCAN_WARPS * WQCAP
#else
    const uint32_t *s_any9 = P.any9;                                         // served by L1
    unsigned *s_wq = (unsigned *)(s_lut7 + 16384);
#endif
    unsigned *s_iq = s_wq + SCAN_WARPS * WQCAP;                              // SCAN_WARPS * IQCAP
    uint8_t *s_need = (uint8_t *)(s_iq + SCAN_WARPS * IQCAP);                // 256
    uint64_t *s_full = (uint64_t *)(s_need + 256);                           // NSTAGE
    uint64_t *s_empty = s_full + NSTAGE;                                     // NSTAGE

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], SCAN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    {   // motif tables -> shared memory (L2-resident after the first CTA)
        const uint4 *src7 = (const uint4 *)P.lut7; uint4 *dst7 = (uint4 *)s_lut7;
        for (int i = tid; i < 16384 / 16; i += SCAN_THREADS) dst7[i] = __ldg(&src7[i]);
#if SCAN_ANY9_SMEM
        const uint4 *src9 = (const uint4 *)P.any9; uint4 *dst9 = (uint4 *)s_any9;
        for (int i = tid; i < 32768 / 16; i += SCAN_THREADS) dst9[i] = __ldg(&src9[i]);
#endif
    }
    for (int f = tid; f < 256; f += SCAN_THREADS) {
        unsigned need = 0;
        for (int w = 0; w < P.nwin; ++w) need |= ((unsigned)f & P.wmask[w]) ? (1u << w) : 0u;
        s_need[f] = (uint8_t)need;
    }
    __syncthreads();

    const unsigned long long tile0 = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE - 1; ++i) {
            const unsigned long long tile = tile0 + (unsigned long long)i * stride;
            if (tile < P.n_tiles) {
                mbar_expect_tx(&s_full[i], STAGE_BYTES);
                bulk_g2s(s_stage + i * STAGE_STRIDE, P.planes + tile * IGD_TILE_CHUNKS, STAGE_BYTES, &s_full[i]);
            }
        }
    }

    unsigned *wq = s_wq + warp * WQCAP;
    unsigned *iq = s_iq + warp * IQCAP;
    unsigned it = 0;
    for (unsigned long long tile = tile0; tile < P.n_tiles; tile += stride, ++it) {
        const int stage = it % NSTAGE;
        const unsigned parity = (it / NSTAGE) & 1u;
        if (tid == 0) {
            // producer: refill the stage that was read during iteration it-1, once all warps left it
            const unsigned long long nxt = tile + (unsigned long long)(NSTAGE - 1) * stride;
            if (nxt < P.n_tiles) {
                const int ns = (it + NSTAGE - 1) % NSTAGE;
                if (it >= 1) {
                    const unsigned ep = ((it - 1) / NSTAGE) & 1u;
                    while (!mbar_try_wait(&s_empty[ns], ep)) { }
                }
                mbar_expect_tx(&s_full[ns], STAGE_BYTES);
                bulk_g2s(s_stage + ns * STAGE_STRIDE, P.planes + nxt * IGD_TILE_CHUNKS, STAGE_BYTES, &s_full[ns]);
            }
        }
        __syncwarp();
        while (!mbar_try_wait(&s_full[stage], parity)) { }

        const uint4 *rec = (const uint4 *)(s_stage + stage * STAGE_STRIDE);
        Stage T;
        T.sw = (const uint32_t *)rec; T.lut7 = s_lut7; T.need = s_need; T.any9 = s_any9;
        T.g0 = tile * (unsigned long long)(IGD_TILE_CHUNKS * IGD_CHUNK_BASES);     // stage record 0 = chunk tile*512
        int nq = 0, ni = 0;                               // queued hits / items of this warp (warp-uniform)
        // ---- phase A: exact heptamer hits of the 64 windows starting in this lane's chunk of each row
        unsigned long long hm[ROWS_PER_WARP];
#pragma unroll
        for (int row = 0; row < ROWS_PER_WARP; ++row) {
            const int c = (row * SCAN_WARPS + warp) * 32 + lane + 1;        // stage record of the chunk
            const uint4 r = rec[c];
            const uint4 n = rec[c + 1];
            uint32_t cand0 = 0xffffffffu, cand1 = 0xffffffffu;
            if (PREFILTER) {
                const Ind i0 = indicators(r.x, r.z), i1 = indicators(r.y, r.w), i2 = indicators(n.x, n.z);
                cand0 = consensus_ge4(i0, i1);
                cand1 = consensus_ge4(i1, i2);
            }
            const uint32_t hit0 = lut_filter(cand0, r.x, r.y, r.z, r.w, s_lut7);
            const uint32_t hit1 = lut_filter(cand1, r.y, n.x, r.w, n.z, s_lut7);
            hm[row] = (unsigned long long)hit0 | ((unsigned long long)hit1 << 32);
        }
        // ---- compaction of the tile's hit positions into the warp queue: one prefix sum per pass; a pass
        //      queues at most what fits, so pathological densities just take more passes
        for (;;) {
            int cnt = 0;
#pragma unroll
            for (int row = 0; row < ROWS_PER_WARP; ++row) cnt += __popcll(hm[row]);
            if (!__any_sync(0xffffffffu, cnt != 0)) break;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            const int space = WQCAP - nq;
            int off = incl - cnt;
#pragma unroll
            for (int row = 0; row < ROWS_PER_WARP; ++row) {
                const int c = (row * SCAN_WARPS + warp) * 32 + lane + 1;
                while (hm[row] && off < space) {
                    const int b = __ffsll((long long)hm[row]) - 1;
                    hm[row] &= hm[row] - 1;
                    wq[nq + off] = (unsigned)(c * 64 + b);
                    ++off;
                }
            }
            nq += min(tot, space);
            __syncwarp();
            while (nq >= 64) {                            // keep room for the next pass; full rounds only
                nq -= 32;
                ni = phase_b(P, T, iq, ni, (int)wq[nq + lane], true, false);
            }
        }
        // ---- phase B for what is left of this tile, then release the stage
        do {
            const int take = min(nq, 32);
            nq -= take;
            ni = phase_b(P, T, iq, ni, lane < take ? (int)wq[nq + lane] : 0, lane < take, nq == 0);
        } while (nq > 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
    }
}

// ================================================================================================
// v5: one persistent CTA of 24 warps per SM; every warp streams its own strips of R rows through a
// private double buffer (own TMA bulk copies, own mbarriers, strips handed out by a global atomic
// counter) -- no CTA-wide coupling after the prologue.  The motif tables are shared by the whole SM:
// the exact 7-mer flags (16 KB) and the 32 KB "centre" bitmap any9c.
//   any9c[c] = 1 iff the 9-mer c, or a 9-mer overlapping it shifted by one base to either side, is
//   in some nonamer set (either strand).  The three candidate nonamers of a window (spacer s-1, s,
//   s+1) start at x, x+1, x+2, so ONE look-up of the 9-mer at x+1 rules the window out for ~95 % of
//   the (heptamer, window) pairs; a 32-position fetch serves the centres of all windows on one side
//   of the heptamer.  Pairs that pass go to a per-warp queue of GLOBAL positions that survives across
//   strips and is resolved exactly (nonamer flags, validity plane, the reference's ranking) in full
//   rounds of 32 from L2.
constexpr int V5_WARPS = 24;
constexpr int V5_THREADS = V5_WARPS * 32;
constexpr int V5_MAXROWS = 4;
constexpr int V5_BUF_BYTES_MAX = (32 * V5_MAXROWS + 2) * 16;    // 2080
constexpr int V5_BUF_STRIDE = 2176;                             // 128-byte aligned
constexpr int V5_WQCAP = 256;                                   // per-warp heptamer-hit queue (stage positions)
constexpr int V5_SQCAP = 64;                                    // per-warp queue of (global position, window) pairs

struct ScanParams5 {
    const uint4 *planes;
    const unsigned long long *inv;
    const uint32_t *invsum;
    const uint8_t *lut7;
    const uint8_t *lut9;
    const uint32_t *any9c;
    unsigned long long n_chunks;
    unsigned long long n_strips;
    int rows;                      // rows (32 chunks each) per strip: 1, 2 or 4
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

// 32 consecutive positions of both bit-planes starting at global base g (rare path, served by L2)
__device__ __forceinline__ void gwin32x2(const ScanParams5 &P, unsigned long long g, unsigned &lo, unsigned &hi)
{
    const unsigned long long r = g >> 6;
    const unsigned long long r1 = (r + 1 < P.n_chunks) ? r + 1 : r;
    const uint4 a = __ldg(&P.planes[r]);
    const uint4 b = __ldg(&P.planes[r1]);
    const unsigned s = (unsigned)g & 63u;
    const bool up = s >= 32u;
    lo = __funnelshift_r(up ? a.y : a.x, up ? b.x : a.y, s);
    hi = __funnelshift_r(up ? a.w : a.z, up ? b.z : a.w, s);
}

__device__ __forceinline__ bool window_valid5(const ScanParams5 &P, unsigned long long g, int k)
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

__device__ __forceinline__ unsigned hept_flags5(const ScanParams5 &P, const uint8_t *s_lut7, unsigned long long g)
{
    unsigned lo, hi;
    gwin32x2(P, g, lo, hi);
    const unsigned f = s_lut7[((hi & 127u) << 7) | (lo & 127u)];
    if (!f) return 0;
    return window_valid5(P, g, 7) ? f : 0;
}

__device__ __forceinline__ void emit5(const ScanParams5 &P, unsigned long long gq, int t, int strand, int d)
{
    const unsigned long long i = atomicAdd(P.count, 1ull);
    if (i < P.cap) P.hits[i] = (gq << 8) | ((unsigned long long)t << 4) | ((unsigned long long)strand << 3) | (unsigned long long)d;
}

// Exact resolution of up to 32 (heptamer position, window) pairs: nonamer flags of the three candidate
// starts, validity, and the reference's ranking (py/IGDetective.py:162-175, inverted for the nonamer-first
// types as in the header of this file).
__device__ __noinline__ void slow_round5(const ScanParams5 &P, const uint8_t *s_lut7, unsigned long long item, bool active)
{
    if (!active) return;
    const unsigned long long gq = item >> 3;
    const int w = (int)(item & 7ull);
    unsigned lo, hi;
    gwin32x2(P, gq, lo, hi);
    unsigned fm = s_lut7[((hi & 127u) << 7) | (lo & 127u)] & P.wmask[w];
    if (!fm) return;
    const int wo = P.woff[w];
    const unsigned long long x = gq + (long long)wo;
    gwin32x2(P, x, lo, hi);
    unsigned nf0 = __ldg(&P.lut9[((hi & 511u) << 9) | (lo & 511u)]);
    unsigned nf1 = __ldg(&P.lut9[(((hi >> 1) & 511u) << 9) | ((lo >> 1) & 511u)]);
    unsigned nf2 = __ldg(&P.lut9[(((hi >> 2) & 511u) << 9) | ((lo >> 2) & 511u)]);
    if (!((nf0 | nf1 | nf2) & fm)) return;
    if (!window_valid5(P, gq, 7)) return;
    if (nf0 && !window_valid5(P, x, 9)) nf0 = 0;
    if (nf1 && !window_valid5(P, x + 1, 9)) nf1 = 0;
    if (nf2 && !window_valid5(P, x + 2, 9)) nf2 = 0;
    const bool right = wo > 0;
    unsigned hm2 = 0, hm1 = 0, hp1 = 0, hp2 = 0; bool have_nb = false;
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
            emit5(P, gq, t, strand, (n & 2u) ? 0 : ((n & bm) ? 1 : 2));
        } else {                                               // each nonamer picks its best-ranked heptamer
            if (!have_nb && (n & 5u)) {
                hm2 = hept_flags5(P, s_lut7, gq - 2); hm1 = hept_flags5(P, s_lut7, gq - 1);
                hp1 = hept_flags5(P, s_lut7, gq + 1); hp2 = hept_flags5(P, s_lut7, gq + 2);
                have_nb = true;
            }
            if (n & 2u) emit5(P, gq, t, strand, 0);
            if (strand == 0) {      // nonamer p = q-9-s' ranks heptamers p+9+s, +s-1, +s+1
                if ((n & bm) && !(hp1 & bit)) emit5(P, gq, t, 0, 1);
                if ((n & bp) && !(hm1 & bit) && !(hm2 & bit)) emit5(P, gq, t, 0, 2);
            } else {                // reverse strand: nonamer n = q+7+s' ranks heptamers n-s-7, +1, -1
                if ((n & bm) && !(hm1 & bit)) emit5(P, gq, t, 1, 1);
                if ((n & bp) && !(hp1 & bit) && !(hp2 & bit)) emit5(P, gq, t, 1, 2);
            }
        }
    }
}

// One round of up to 32 queued heptamer hits (lane `active` holds stage position q): the heptamer's flags
// say which windows matter; one centre look-up per window decides whether the pair is looked at exactly.
// nsq = pairs waiting in sq (warp-uniform, < 32 on entry and on return).
__device__ __forceinline__ int hits_round5(const ScanParams5 &P, const uint32_t *sw, const uint8_t *s_lut7,
                                           const uint8_t *s_need, const uint32_t *s_any9c, unsigned long long g0,
                                           int q, bool active, unsigned long long *sq, int nsq)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned need = 0, lo = 0, hi = 0;
    if (active) {
        swin32x2(sw, q, lo, hi);
        need = s_need[s_lut7[((hi & 127u) << 7) | (lo & 127u)]];
    }
    for (int w = 0; w < P.nwin; ++w) {
        if (P.wnew[w] && (need & P.wgrp[w])) swin32x2(sw, q + P.wbase[w], lo, hi);
        bool pass = false;
        if ((need >> w) & 1u) {
            const int sh = P.wsh[w];
            const unsigned idx = (((hi >> sh) & 511u) << 9) | ((lo >> sh) & 511u);
            pass = (s_any9c[idx >> 5] >> (idx & 31u)) & 1u;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            if (pass) sq[nsq + __popc(m & ((1u << lane) - 1u))] = ((g0 + (unsigned long long)q) << 3) | (unsigned long long)w;
            nsq += __popc(m);
            __syncwarp();
            if (nsq >= 32) {
                nsq -= 32;
                slow_round5(P, s_lut7, sq[nsq + lane], true);
                __syncwarp();
            }
        }
    }
    return nsq;
}

static constexpr size_t kScan5Smem = (size_t)V5_WARPS * 2 * V5_BUF_STRIDE + 16384 + 32768 + (size_t)V5_WARPS * V5_WQCAP * 4
                                     + (size_t)V5_WARPS * V5_SQCAP * 8 + 256 + (size_t)V5_WARPS * 2 * 8 + 16;

template <bool PREFILTER>
__global__ void __launch_bounds__(V5_THREADS, 1)
rss_scan5_kernel(const __grid_constant__ ScanParams5 P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_buf = smem;                                             // V5_WARPS * 2 * V5_BUF_STRIDE
    uint8_t *s_lut7 = smem + V5_WARPS * 2 * V5_BUF_STRIDE;                   // 16384
    uint32_t *s_any9c = (uint32_t *)(s_lut7 + 16384);                        // 32768
    unsigned *s_wq = s_any9c + 8192;                                         // V5_WARPS * V5_WQCAP
    unsigned long long *s_sq = (unsigned long long *)(s_wq + V5_WARPS * V5_WQCAP);   // V5_WARPS * V5_SQCAP
    uint8_t *s_need = (uint8_t *)(s_sq + V5_WARPS * V5_SQCAP);               // 256
    uint64_t *s_full = (uint64_t *)(s_need + 256);                           // V5_WARPS * 2

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *full = s_full + 2 * warp;
    if (lane == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // strips are handed out by a global counter; the warp's first copy is in flight during the prologue
    const unsigned rows = (unsigned)P.rows;
    const unsigned strip_chunks = rows * 32u;
    const unsigned buf_bytes = (strip_chunks + 2u) * 16u;
    unsigned char *buf0 = s_buf + (size_t)warp * 2 * V5_BUF_STRIDE;
    unsigned long long s_cur = 0, s_nxt = 0;
    if (lane == 0) {
        s_cur = atomicAdd(&P.count[1], 1ull);
        s_nxt = atomicAdd(&P.count[1], 1ull);
        if (s_cur < P.n_strips) {
            mbar_expect_tx(&full[0], buf_bytes);
            bulk_g2s(buf0, P.planes + s_cur * strip_chunks, buf_bytes, &full[0]);
        }
    }
    s_cur = __shfl_sync(0xffffffffu, s_cur, 0);
    {   // motif tables -> shared memory (L2-resident after the first CTA)
        const uint4 *src7 = (const uint4 *)P.lut7; uint4 *dst7 = (uint4 *)s_lut7;
        for (int i = tid; i < 16384 / 16; i += V5_THREADS) dst7[i] = __ldg(&src7[i]);
        const uint4 *src9 = (const uint4 *)P.any9c; uint4 *dst9 = (uint4 *)s_any9c;
        for (int i = tid; i < 32768 / 16; i += V5_THREADS) dst9[i] = __ldg(&src9[i]);
    }
    for (int f = tid; f < 256; f += V5_THREADS) {
        unsigned need = 0;
        for (int w = 0; w < P.nwin; ++w) need |= ((unsigned)f & P.wmask[w]) ? (1u << w) : 0u;
        s_need[f] = (uint8_t)need;
    }
    __syncthreads();

    unsigned *wq = s_wq + warp * V5_WQCAP;
    unsigned long long *sq = s_sq + warp * V5_SQCAP;
    int nsq = 0;
    for (unsigned it = 0; s_cur < P.n_strips; ++it) {
        unsigned long long s_nn = 0;
        if (lane == 0) {
            // the other buffer was read during iteration it-1 (all lanes passed the __syncwarp at its end)
            if (s_nxt < P.n_strips) {
                mbar_expect_tx(&full[(it + 1) & 1u], buf_bytes);
                bulk_g2s(buf0 + ((it + 1) & 1u) * V5_BUF_STRIDE, P.planes + s_nxt * strip_chunks, buf_bytes, &full[(it + 1) & 1u]);
            }
            s_nn = atomicAdd(&P.count[1], 1ull);
        }
        __syncwarp();
        while (!mbar_try_wait(&full[it & 1u], (it >> 1) & 1u)) { }

        const uint4 *rec = (const uint4 *)(buf0 + (it & 1u) * V5_BUF_STRIDE);
        const uint32_t *sw = (const uint32_t *)rec;
        const unsigned long long g0 = s_cur * (unsigned long long)(strip_chunks * IGD_CHUNK_BASES);   // record 0 = chunk s_cur * strip_chunks
        int nq = 0;
        for (unsigned r0 = 0; r0 < rows; r0 += 2) {
            // ---- phase A: exact heptamer hits of the 64 windows starting in this lane's chunk of each row
            unsigned long long hm[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                hm[k] = 0;
                if (r0 + k < rows) {
                    const int c = (int)(r0 + k) * 32 + lane + 1;             // stage record of the chunk
                    const uint4 r = rec[c];
                    const uint4 n = rec[c + 1];
                    uint32_t cand0 = 0xffffffffu, cand1 = 0xffffffffu;
                    if (PREFILTER) {
                        const Ind i0 = indicators(r.x, r.z), i1 = indicators(r.y, r.w), i2 = indicators(n.x, n.z);
                        cand0 = consensus_ge4(i0, i1);
                        cand1 = consensus_ge4(i1, i2);
                    }
                    const uint32_t hit0 = lut_filter(cand0, r.x, r.y, r.z, r.w, s_lut7);
                    const uint32_t hit1 = lut_filter(cand1, r.y, n.x, r.w, n.z, s_lut7);
                    hm[k] = (unsigned long long)hit0 | ((unsigned long long)hit1 << 32);
                }
            }
            // ---- compaction into the warp queue (a pass queues at most what fits), full rounds as they form
            for (;;) {
                const int cnt = __popcll(hm[0]) + __popcll(hm[1]);
                if (!__any_sync(0xffffffffu, cnt != 0)) break;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                const int space = V5_WQCAP - nq;
                int off = incl - cnt;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int c = (int)(r0 + k) * 32 + lane + 1;
                    while (hm[k] && off < space) {
                        const int b = __ffsll((long long)hm[k]) - 1;
                        hm[k] &= hm[k] - 1;
                        wq[nq + off] = (unsigned)(c * 64 + b);
                        ++off;
                    }
                }
                nq += min(tot, space);
                __syncwarp();
                while (nq >= 32) {
                    nq -= 32;
                    nsq = hits_round5(P, sw, s_lut7, s_need, s_any9c, g0, (int)wq[nq + lane], true, sq, nsq);
                }
                __syncwarp();
                if (tot <= space) break;
            }
        }
        if (nq > 0) nsq = hits_round5(P, sw, s_lut7, s_need, s_any9c, g0, lane < nq ? (int)wq[lane] : 0, lane < nq, sq, nsq);
        __syncwarp();
        s_cur = __shfl_sync(0xffffffffu, s_nxt, 0);
        s_nxt = s_nn;
    }
    if (nsq > 0) slow_round5(P, s_lut7, sq[lane < nsq ? lane : 0], lane < nsq);
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
            const uint32_t l0 = half ? r.y : r.x, l1 = half ? n.x : r.y;
            const uint32_t h0 = half ? r.w : r.z, h1 = half ? n.z : r.w;
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

} // namespace

static const size_t kScanSmem = NSTAGE * STAGE_STRIDE + 16384 + (SCAN_ANY9_SMEM ? 32768 : 0) + SCAN_WARPS * (WQCAP + IQCAP) * 4 + 256 + 2 * NSTAGE * 8 + 16;

static int launch_rss_scan_v4(igd_handle *h, const int32_t spacer[4])
{
    ScanParams P;
    P.planes = h->d_planes; P.inv = h->d_inv; P.invsum = h->d_invsum;
    P.lut7 = h->d_lut7; P.lut9 = h->d_lut9; P.any9 = h->d_any9;
    P.n_tiles = h->n_tiles;
    for (int i = 0; i < 4; ++i) P.sp[i] = spacer[i];
    // one window per distinct (side, spacer): right of the heptamer for (heptamer-first, +) and
    // (nonamer-first, -), left of it for the other two combinations
    P.nwin = 0;
    for (int side = 0; side < 2; ++side)
        for (int t = 0; t < 4; ++t) {
            const bool hept_first = (t == IGD_SIG_V || t == IGD_SIG_D_RIGHT);
            const unsigned bit = side == 0 ? (hept_first ? (1u << t) : (16u << t)) : (hept_first ? (16u << t) : (1u << t));
            const int off = side == 0 ? 7 + spacer[t] - 1 : -9 - spacer[t] - 1;
            int w = 0;
            while (w < P.nwin && P.woff[w] != off) ++w;
            if (w == P.nwin) { P.woff[w] = off; P.wmask[w] = 0; P.nwin++; }
            P.wmask[w] |= bit;
        }
    for (int w = P.nwin; w < 8; ++w) { P.woff[w] = 0; P.wmask[w] = 0; }
    P.hits = h->d_hits; P.cap = h->hit_cap; P.count = h->d_count;
    auto kernel = h->consensus_prefilter ? rss_scan_kernel<true> : rss_scan_kernel<false>;
    IGD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
    int per_sm = 0;
    IGD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SCAN_THREADS, kScanSmem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (unsigned long long)h->sm_count * per_sm;
    if (blocks > h->n_tiles) blocks = h->n_tiles;
    if (blocks < 1) blocks = 1;
    kernel<<<(unsigned)blocks, SCAN_THREADS, kScanSmem, h->stream>>>(P);
    h->timing.total_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

int igd_launch_rss_scan(igd_handle *h, const int32_t spacer[4])
{
    if (getenv("IGD_SCAN_V4")) return launch_rss_scan_v4(h, spacer);
    ScanParams5 P;
    P.planes = h->d_planes; P.inv = h->d_inv; P.invsum = h->d_invsum;
    P.lut7 = h->d_lut7; P.lut9 = h->d_lut9; P.any9c = h->d_any9c;
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
    // rows per strip: as long as every warp still gets several strips (tail balance), longer strips
    const unsigned long long rows_total = h->n_tiles * (unsigned long long)(IGD_TILE_CHUNKS / 32);
    const unsigned long long warps = (unsigned long long)h->sm_count * V5_WARPS;
    P.rows = 4;
    while (P.rows > 1 && rows_total / P.rows < 6 * warps) P.rows >>= 1;
    P.n_strips = rows_total / P.rows;
    P.hits = h->d_hits; P.cap = h->hit_cap; P.count = h->d_count;
    auto kernel = h->consensus_prefilter ? rss_scan5_kernel<true> : rss_scan5_kernel<false>;
    IGD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScan5Smem));
    unsigned long long blocks = (unsigned long long)h->sm_count;
    const unsigned long long wanted = (P.n_strips + V5_WARPS - 1) / V5_WARPS;
    if (blocks > wanted) blocks = wanted;
    if (blocks < 1) blocks = 1;
    kernel<<<(unsigned)blocks, V5_THREADS, kScan5Smem, h->stream>>>(P);
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
