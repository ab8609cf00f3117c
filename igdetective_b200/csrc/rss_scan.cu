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
// Data movement: a persistent CTA walks tiles of 512 chunks (32 Kbases, 8 KB of planes); each
// tile plus one halo chunk per side arrives in shared memory by one TMA bulk copy
// (cp.async.bulk + mbarrier, 3 stages in flight), so every base is read from HBM once
// (0.25 B/base).
// Phase A, one 64-base chunk per thread, is bit-parallel on the planes: every heptamer of the
// shipped motif file (both strands) agrees with the consensus CAC.GTG in at least 4 of its 6
// fixed positions, so a bit-sliced population count of the six letter-match planes, thresholded
// at 4, passes every possible heptamer hit and only ~3.8 % of random positions (igd_set_motifs
// verifies that property for the tables it is given; tables that break it use the exhaustive
// variant of the same kernel).  Survivors are looked up in the exact 7-mer flag table in shared
// memory; real heptamer hits (~1 % of positions) go to a shared-memory queue.
// Phase B pairs the queued heptamers with all lanes busy.
#include "common.cuh"

namespace {

constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = IGD_STAGE_CHUNKS * 16;          // 8224
constexpr int STAGE_STRIDE = 8320;                          // 128-byte aligned
constexpr int QCAP = 2048;
constexpr int INVSUM_WORDS = IGD_TILE_CHUNKS / 32 + 1;

struct ScanParams {
    const uint4 *planes;
    const unsigned long long *inv;
    const uint32_t *invsum;
    const uint8_t *lut7;
    const uint8_t *lut9;
    const uint32_t *any9;
    unsigned long long n_tiles;
    int sp[4];
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

struct TileCtx {
    const uint32_t *sw;            // stage words: record r = {lo0, lo1, hi0, hi1} at sw[4r..4r+3]
    const uint8_t *s_lut7;
    const uint32_t *s_invsum;      // validity summary bits of the stage's chunks
    unsigned long long g0;         // global base index of stage position 0
};

// 32 consecutive positions of one bit-plane starting at stage position p (plane = 0 lo, 2 hi)
__device__ __forceinline__ uint32_t win32(const uint32_t *sw, int p, int plane)
{
    const int b0 = p >> 5, b1 = b0 + 1;
    const uint32_t a = sw[4 * (b0 >> 1) + (b0 & 1) + plane];
    const uint32_t b = sw[4 * (b1 >> 1) + (b1 & 1) + plane];
    return __funnelshift_r(a, b, p & 31);
}

// true iff no base of stage window [x, x+k) is marked invalid
__device__ __forceinline__ bool window_valid(const ScanParams &P, const TileCtx &T, int x, int k)
{
    const int c0 = x >> 6, c1 = (x + k - 1) >> 6, b0 = x & 63;
    if ((T.s_invsum[c0 >> 5] >> (c0 & 31)) & 1u) {
        const unsigned long long m = __ldg(&P.inv[(T.g0 >> 6) + c0]);
        const int n0 = (b0 + k > 64) ? 64 - b0 : k;
        if (m & (((1ull << n0) - 1ull) << b0)) return false;
    }
    if (c1 != c0 && ((T.s_invsum[c1 >> 5] >> (c1 & 31)) & 1u)) {
        const unsigned long long m = __ldg(&P.inv[(T.g0 >> 6) + c1]);
        if (m & ((1ull << (b0 + k - 64)) - 1ull)) return false;
    }
    return true;
}

// heptamer flags at stage position x (0 if the window holds an invalid base)
__device__ __forceinline__ unsigned hept_flags(const ScanParams &P, const TileCtx &T, int x)
{
    const unsigned lo = win32(T.sw, x, 0) & 127u, hi = win32(T.sw, x, 2) & 127u;
    const unsigned f = T.s_lut7[(hi << 7) | lo];
    if (!f) return 0;
    return window_valid(P, T, x, 7) ? f : 0;
}

// flags of the three nonamers starting at x, x+1, x+2 (the spacers s+1, s, s-1 or s-1, s, s+1 of
// one heptamer), masked by `bit`, validity included: bit i of the result = nonamer x+i qualifies
__device__ __forceinline__ unsigned nona3(const ScanParams &P, const TileCtx &T, int x, unsigned bit)
{
    const unsigned lo = win32(T.sw, x, 0), hi = win32(T.sw, x, 2);
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const unsigned idx = (((hi >> i) & 511u) << 9) | ((lo >> i) & 511u);
        if ((__ldg(&P.any9[idx >> 5]) >> (idx & 31)) & 1u)
            if ((__ldg(&P.lut9[idx]) & bit) && window_valid(P, T, x + i, 9)) r |= 1u << i;
    }
    return r;
}

__device__ __forceinline__ void emit(const ScanParams &P, unsigned long long gq, int t, int strand, int d)
{
    const unsigned long long i = atomicAdd(P.count, 1ull);
    if (i < P.cap) P.hits[i] = (gq << 8) | ((unsigned long long)t << 4) | ((unsigned long long)strand << 3) | (unsigned long long)d;
}

// all RSSs whose heptamer starts at stage position q; f = LUT flags of that heptamer
__device__ __noinline__ void eval_heptamer(const ScanParams &P, const TileCtx &T, int q, unsigned f)
{
    if (!window_valid(P, T, q, 7)) return;
    const unsigned long long gq = T.g0 + q;
    unsigned hm2 = 0, hm1 = 0, hp1 = 0, hp2 = 0; bool have_nb = false;
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        const unsigned bF = 1u << t, bR = 16u << t;
        if (!(f & (bF | bR))) continue;
        const int s = P.sp[t];
        const bool hept_first = (t == IGD_SIG_V || t == IGD_SIG_D_RIGHT);
        if (hept_first) {
            if (f & bF) {              // nonamers at q+7+{s-1, s, s+1} = bits 0,1,2; priority s, s-1, s+1
                const unsigned n = nona3(P, T, q + 7 + s - 1, bF);
                if (n) emit(P, gq, t, 0, (n & 2u) ? 0 : ((n & 1u) ? 1 : 2));
            }
            if (f & bR) {              // nonamers at q-9-{s+1, s, s-1} = bits 0,1,2
                const unsigned n = nona3(P, T, q - 9 - s - 1, bR);
                if (n) emit(P, gq, t, 1, (n & 2u) ? 0 : ((n & 4u) ? 1 : 2));
            }
        } else {
            if (!have_nb) {
                hm2 = hept_flags(P, T, q - 2); hm1 = hept_flags(P, T, q - 1);
                hp1 = hept_flags(P, T, q + 1); hp2 = hept_flags(P, T, q + 2);
                have_nb = true;
            }
            if (f & bF) {              // nonamer p = q-9-s' ranks its heptamers p+9+s, +s-1, +s+1
                const unsigned n = nona3(P, T, q - 9 - s - 1, bF);     // bits: s+1, s, s-1
                if (n & 2u) emit(P, gq, t, 0, 0);
                if ((n & 4u) && !(hp1 & bF)) emit(P, gq, t, 0, 1);
                if ((n & 1u) && !(hm1 & bF) && !(hm2 & bF)) emit(P, gq, t, 0, 2);
            }
            if (f & bR) {              // reverse strand: nonamer n = q+7+s' ranks heptamers n-s-7, +1, -1
                const unsigned n = nona3(P, T, q + 7 + s - 1, bR);     // bits: s-1, s, s+1
                if (n & 2u) emit(P, gq, t, 1, 0);
                if ((n & 1u) && !(hm1 & bR)) emit(P, gq, t, 1, 1);
                if ((n & 4u) && !(hp1 & bR) && !(hp2 & bR)) emit(P, gq, t, 1, 2);
            }
        }
    }
}

// positions of word (w, next) whose 7-mer matches CAC.GTG in at least 4 of the 6 fixed positions.
// A/C/G/T = letter indicator planes of the word, nA.. = of the following word.
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

template <bool PREFILTER>
__global__ void __launch_bounds__(IGD_TILE_CHUNKS, 2)
rss_scan_kernel(const ScanParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_stage = smem;                                           // NSTAGE * STAGE_STRIDE
    uint8_t *s_lut7 = smem + NSTAGE * STAGE_STRIDE;                          // 16384
    uint32_t *s_queue = (uint32_t *)(s_lut7 + 16384);                        // QCAP
    uint32_t *s_invsum = s_queue + QCAP;                                     // INVSUM_WORDS (+pad)
    uint64_t *s_bar = (uint64_t *)(s_invsum + 32);                           // NSTAGE
    unsigned *s_qn = (unsigned *)(s_bar + NSTAGE);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(&s_bar[i], 1);
        *s_qn = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const unsigned long long tile0 = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE - 1; ++i) {
            const unsigned long long tile = tile0 + (unsigned long long)i * stride;
            if (tile < P.n_tiles) {
                mbar_expect_tx(&s_bar[i], STAGE_BYTES);
                bulk_g2s(s_stage + i * STAGE_STRIDE, P.planes + tile * IGD_TILE_CHUNKS, STAGE_BYTES, &s_bar[i]);
            }
        }
    }
    {   // exact 7-mer flag table -> shared memory (L2-resident after the first CTA)
        const uint4 *src7 = (const uint4 *)P.lut7; uint4 *dst7 = (uint4 *)s_lut7;
        for (int i = tid; i < 16384 / 16; i += IGD_TILE_CHUNKS) dst7[i] = __ldg(&src7[i]);
    }
    __syncthreads();

    unsigned it = 0;
    for (unsigned long long tile = tile0; tile < P.n_tiles; tile += stride, ++it) {
        const int stage = it % NSTAGE;
        const unsigned parity = (it / NSTAGE) & 1u;
        if (tid == 0) {                                   // refill the stage consumed last iteration
            const unsigned long long nxt = tile + (unsigned long long)(NSTAGE - 1) * stride;
            if (nxt < P.n_tiles) {
                const int ns = (it + NSTAGE - 1) % NSTAGE;
                mbar_expect_tx(&s_bar[ns], STAGE_BYTES);
                bulk_g2s(s_stage + ns * STAGE_STRIDE, P.planes + nxt * IGD_TILE_CHUNKS, STAGE_BYTES, &s_bar[ns]);
            }
        }
        if (tid < INVSUM_WORDS) s_invsum[tid] = __ldg(&P.invsum[tile * (IGD_TILE_CHUNKS / 32) + tid]);
        while (!mbar_try_wait(&s_bar[stage], parity)) { }

        const uint32_t *sw = (const uint32_t *)(s_stage + stage * STAGE_STRIDE);
        TileCtx T;
        T.sw = sw; T.s_lut7 = s_lut7; T.s_invsum = s_invsum;
        T.g0 = tile * (unsigned long long)(IGD_TILE_CHUNKS * IGD_CHUNK_BASES);   // stage record 0 = chunk tile*512

        // ---- phase A: heptamer candidates of the 64 windows starting in this thread's chunk
        {
            const uint4 r = ((const uint4 *)sw)[tid + 1];
            const uint4 n = ((const uint4 *)sw)[tid + 2];
            uint32_t cand0 = 0xffffffffu, cand1 = 0xffffffffu;
            if (PREFILTER) {
                const Ind i0 = indicators(r.x, r.z), i1 = indicators(r.y, r.w), i2 = indicators(n.x, n.z);
                cand0 = consensus_ge4(i0, i1);
                cand1 = consensus_ge4(i1, i2);
            }
            const int p0 = (tid + 1) * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t l0 = half ? r.y : r.x, l1 = half ? n.x : r.y;
                const uint32_t h0 = half ? r.w : r.z, h1 = half ? n.z : r.w;
                uint32_t m = half ? cand1 : cand0;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned lo = __funnelshift_r(l0, l1, b) & 127u;
                    const unsigned hi = __funnelshift_r(h0, h1, b) & 127u;
                    const unsigned f = s_lut7[(hi << 7) | lo];
                    if (f) {
                        const unsigned slot = atomicAdd(s_qn, 1u);
                        const int q = p0 + half * 32 + b;
                        if (slot < QCAP) s_queue[slot] = (unsigned)q | (f << 16);
                        else eval_heptamer(P, T, q, f);
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase B: pair the queued heptamers
        {
            const unsigned n = min(*s_qn, (unsigned)QCAP);
            for (unsigned i = tid; i < n; i += IGD_TILE_CHUNKS) {
                const unsigned e = s_queue[i];
                eval_heptamer(P, T, (int)(e & 0xffffu), e >> 16);
            }
        }
        __syncthreads();          // all reads of this stage, the queue and s_invsum are done
        if (tid == 0) *s_qn = 0;
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

static const size_t kScanSmem = NSTAGE * STAGE_STRIDE + 16384 + QCAP * 4 + 32 * 4 + NSTAGE * 8 + 16;

int igd_launch_rss_scan(igd_handle *h, const int32_t spacer[4])
{
    ScanParams P;
    P.planes = h->d_planes; P.inv = h->d_inv; P.invsum = h->d_invsum;
    P.lut7 = h->d_lut7; P.lut9 = h->d_lut9; P.any9 = h->d_any9;
    P.n_tiles = h->n_tiles;
    for (int i = 0; i < 4; ++i) P.sp[i] = spacer[i];
    P.hits = h->d_hits; P.cap = h->hit_cap; P.count = h->d_count;
    auto kernel = h->consensus_prefilter ? rss_scan_kernel<true> : rss_scan_kernel<false>;
    IGD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
    int per_sm = 0;
    IGD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, IGD_TILE_CHUNKS, kScanSmem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (unsigned long long)h->sm_count * per_sm;
    if (blocks > h->n_tiles) blocks = h->n_tiles;
    if (blocks < 1) blocks = 1;
    kernel<<<(unsigned)blocks, IGD_TILE_CHUNKS, kScanSmem, h->stream>>>(P);
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
