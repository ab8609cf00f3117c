// K4 / K5: upper bounds on BioAlign's match count and PI for (target, gene) pairs, used by igd_detect to skip alignments
// that cannot matter.  K4 (LCS, bit-parallel) first; K5 (a one-state DP that charges insertions, below) where K4 is too weak.
//
// The columns that NumMatches counts (py/extract_aligned_genes.py:33-38: equal upper-cased letters inside the gene
// range) form a common subsequence of the upper-cased fragment and gene, whatever alignment BioPython returns, so
//     NumMatches + 1  <=  LCS(upper(fragment), upper(gene))      and      len(alignment) >= len(gene),
// hence PI = NumMatches / len * 100 <= (LCS - 1) / len(gene) * 100.  The LCS length is computed bit-parallel
// (Crochemore et al. / Hyyro): the gene's columns are the bits of a multi-word vector V, one row step per fragment
// letter is  U = V & M[letter];  V = (V + U) | (V & ~M[letter]),  and the LCS is the number of zero bits at the end --
// 3 instructions per 32 gene columns and fragment letter, ~1.5 % of the work of the Gotoh DP over the same pair.
// Letters outside A, C, G, T fall into one class that matches itself (an over-estimate: still an upper bound).
//
// Mapping: one thread per gene, holding that gene's five match masks in registers; the threads of a CTA walk the same
// targets letter by letter (the letter is warp-uniform, so the choice of mask set is a uniform branch).
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace {

constexpr int LCS_THREADS = 128;

__device__ __forceinline__ int letter_class(unsigned c)
{
    if (c - 'a' < 26u) c -= 32u;
    return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : 4)));
}

// masks[g][class][w]: bit i of word w = upper(gene g)[32 w + i] is in the class.  One thread per (gene, word).
__global__ void __launch_bounds__(LCS_THREADS)
gene_masks_kernel(const uint8_t *__restrict__ qry, const int32_t *__restrict__ qry_off, int n_qry, int W, uint32_t *__restrict__ masks)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = k / W, w = k - g * W;
    if (g >= n_qry) return;
    const int b0 = qry_off[g], n = qry_off[g + 1] - b0;
    uint32_t m[5] = {0u, 0u, 0u, 0u, 0u};
    for (int i = 32 * w; i < n && i < 32 * w + 32; ++i) {
        const int c = letter_class(qry[b0 + i]);
#pragma unroll
        for (int x = 0; x < 5; ++x) m[x] |= (c == x ? 1u : 0u) << (i & 31);
    }
#pragma unroll
    for (int x = 0; x < 5; ++x) masks[((size_t)g * 5 + x) * W + w] = m[x];
}

// V += U over W words with carry propagation; blocks of four words, the carry travels in a register between blocks
template <int W>
__device__ __forceinline__ void add_words(const uint32_t (&v)[W], const uint32_t (&u)[W], uint32_t (&s)[W])
{
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < W; w += 4) {
        if (w + 4 <= W) {
            asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %5, 0xffffffff;\n\t"
                "addc.cc.u32 %0, %6, %10;\n\taddc.cc.u32 %1, %7, %11;\n\taddc.cc.u32 %2, %8, %12;\n\taddc.cc.u32 %3, %9, %13;\n\t"
                "addc.u32 %4, 0, 0;\n\t}"
                : "=&r"(s[w]), "=&r"(s[w + 1]), "=&r"(s[w + 2]), "=&r"(s[w + 3]), "=&r"(carry)
                : "r"(carry), "r"(v[w]), "r"(v[w + 1]), "r"(v[w + 2]), "r"(v[w + 3]), "r"(u[w]), "r"(u[w + 1]), "r"(u[w + 2]), "r"(u[w + 3]));
        } else {
#pragma unroll
            for (int k = w; k < W; ++k) {
                asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, 0xffffffff;\n\taddc.cc.u32 %0, %3, %4;\n\taddc.u32 %1, 0, 0;\n\t}"
                    : "=&r"(s[k]), "=&r"(carry) : "r"(carry), "r"(v[k]), "r"(u[k]));
            }
        }
    }
}

// u += v over eleven words in one carry chain (u and v are disjoint register sets: 22 operands)
__device__ __forceinline__ void add_in_place_11(uint32_t (&u)[11], const uint32_t (&v)[11])
{
    asm("add.cc.u32 %0, %0, %11;\n\taddc.cc.u32 %1, %1, %12;\n\taddc.cc.u32 %2, %2, %13;\n\taddc.cc.u32 %3, %3, %14;\n\t"
        "addc.cc.u32 %4, %4, %15;\n\taddc.cc.u32 %5, %5, %16;\n\taddc.cc.u32 %6, %6, %17;\n\taddc.cc.u32 %7, %7, %18;\n\t"
        "addc.cc.u32 %8, %8, %19;\n\taddc.cc.u32 %9, %9, %20;\n\taddc.u32 %10, %10, %21;"
        : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]), "+r"(u[9]), "+r"(u[10])
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]));
}

template <int W>
__device__ __forceinline__ void lcs_row(uint32_t (&V)[W], const uint32_t (&M)[W])
{
    uint32_t U[W];
#pragma unroll
    for (int w = 0; w < W; ++w) U[w] = V[w] & M[w];
    if constexpr (W == 11) {
        add_in_place_11(U, V);                                                   // U = V + (V & M)
#pragma unroll
        for (int w = 0; w < W; ++w) V[w] = U[w] | (V[w] & ~M[w]);                // one LOP3
    } else {
        uint32_t S[W];
        add_words<W>(V, U, S);
#pragma unroll
        for (int w = 0; w < W; ++w) V[w] = S[w] | (V[w] ^ U[w]);
    }
}

template <int W>
__global__ void __launch_bounds__(LCS_THREADS)
lcs_bound_kernel(const uint8_t *__restrict__ tgt /* class bytes */, const int32_t *__restrict__ tgt_off, int n_tgt, int tgt_per_block,
                 const int32_t *__restrict__ qry_off, int n_qry, const uint32_t *__restrict__ masks, uint16_t *__restrict__ out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = g < n_qry;
    uint32_t M[5][W];
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int w = 0; w < W; ++w) M[c][w] = have ? masks[((size_t)g * 5 + c) * W + w] : 0u;
    const int n = have ? qry_off[g + 1] - qry_off[g] : 0;
    const int t0 = blockIdx.y * tgt_per_block, t1 = min(n_tgt, t0 + tgt_per_block);
    for (int t = t0; t < t1; ++t) {
        const int a0 = tgt_off[t], nA = tgt_off[t + 1] - a0;
        uint32_t V[W];
#pragma unroll
        for (int w = 0; w < W; ++w) V[w] = 0xffffffffu;
        for (int i = 0; i < nA; ++i) {
            const int c = __ldg(&tgt[a0 + i]);                                  // class byte: the same for every thread
            switch (c) {
            case 0: lcs_row<W>(V, M[0]); break;
            case 1: lcs_row<W>(V, M[1]); break;
            case 2: lcs_row<W>(V, M[2]); break;
            case 3: lcs_row<W>(V, M[3]); break;
            default: lcs_row<W>(V, M[4]); break;
            }
        }
        if (have) {
            int zeros = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int bits = min(32, max(0, n - 32 * w));                    // gene columns held by this word
                const uint32_t valid = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
                zeros += __popc(~V[w] & valid);
            }
            out[(size_t)t * n_qry + g] = (uint16_t)zeros;
        }
    }
}

// ub[u][g] = max(LCS(fragment u, g), LCS(rc(fragment u), g)) (rows 2u, 2u+1 of lcs) and top[u][0..K) = the K genes with
// the largest (ub - 1) / len(g), in no particular order.  One warp per fragment; a lane owns genes lane, lane + 32, ...
// (at most 64 of them: n_qry <= 2048).  Which genes come first is a heuristic (single precision); everything that
// decides what must be aligned is exact and happens on the integers.
__global__ void __launch_bounds__(128)
ub_topk_kernel(const uint16_t *__restrict__ lcs, const int32_t *__restrict__ qry_off, int n_unique, int n_qry, int K,
               uint16_t *__restrict__ ub, int32_t *__restrict__ top)
{
    const int u = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n_unique) return;
    const uint16_t *f = lcs + (size_t)(2 * u) * n_qry, *r = f + n_qry;
    unsigned long long taken = 0;
    for (int g = lane; g < n_qry; g += 32) ub[(size_t)u * n_qry + g] = max(f[g], r[g]);
    for (int k = 0; k < K; ++k) {
        float best = -1e30f; int bi = -1;
        int slot = 0;
        for (int g = lane; g < n_qry; g += 32, ++slot) {
            if ((taken >> slot) & 1ull) continue;
            const float v = ((float)max(f[g], r[g]) - 1.0f) / (float)(qry_off[g + 1] - qry_off[g]);
            if (v > best) { best = v; bi = g; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
        }
        if (bi >= 0 && (bi & 31) == lane) taken |= 1ull << (bi >> 5);
        if (lane == 0) top[(size_t)u * K + k] = bi;
    }
}

// cls[i] = letter_class(qry[i]): the shift amount the bound kernel below uses for a gene letter
__global__ void __launch_bounds__(256)
gene_class_kernel(const uint8_t *__restrict__ qry, size_t n, uint8_t *__restrict__ cls)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        cls[i] = (uint8_t)letter_class(qry[i]);
}

// K5: Y(fragment, gene) = max over ALL alignments of  2 x (equal upper-cased letters) - (fragment letters inserted between
// the first and the last gene letter).  For the alignment BioPython returns, with m equal columns and c inserted letters
// inside the gene range, PI = (m - 1) / (len(gene) + c) (py/extract_aligned_genes.py:33-41, 52-53), so PI >= p with
// p >= 1/2 gives  2 m - c >= 2 p len(gene) + 2,  hence  Y < 2 p len(gene) + 2  proves PI < p.  Unlike the LCS bound
// ((LCS - 1) / len(gene) ~ 0.69 for unrelated DNA: useless against the 0.60 cut-off) this one charges the insertions a
// long common subsequence needs: Y / (2 len) ~ 0.55 for unrelated pairs, and no unrelated pair of the bench workloads
// reaches 0.60.  One DP state per cell, H(i,j) = max(H(i-1,j-1) + 2 [a_i == b_j], H(i-1,j) - 1, H(i,j-1)) with
// H(i,0) = H(0,j) = 0 and free vertical moves in the last column (trailing fragment letters), so the answer arrives in
// the bottom row.  Two cells per instruction: the fragment (low half) and its reverse complement (high half) against
// the same gene share every control decision, and a cell pair costs PRMT + 2 IMAD + VIMNMX3.S16x2 (ybound_column below)
// against 13.8 instructions per cell of the Gotoh kernel.  Same systolic mapping and run-mode work items as gotoh_packed_kernel: G lanes x C rows, runs of
// consecutive genes, persistent grid.
struct YParams {
    const uint8_t *tgt; const int32_t *tgt_off;          // rows 2u / 2u + 1: fragment u / its reverse complement (same length)
    const uint8_t *qcls; const int32_t *qry_off;
    const int4 *runs; int n_runs;                        // {u, first gene, count, first output slot}
    uint32_t *out;                                       // Y(rc) << 16 | Y(forward)
    unsigned int *counter;
    // optional: the pairs whose bound reaches the cut-off, as {fragment, gene}, appended in no particular order
    double cut; uint2 *pass; unsigned int *pass_count; unsigned int pass_cap;
    int one;                                             // 1, opaque to the compiler: x * one + y is an IMAD
};

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// The K5 recurrence in the ROW FRAME: every value of row i is stored plus i, so the vertical move (score - 1) is the
// stored value of the row above unchanged, the diagonal move carries + 1 on top of its 2 [match], and a cell is
// max3(diag + inc, up + pen, left).  Per cell pair: PRMT (increment look-up: the row's four bytes per orientation, 3 where
// the row letter is that class and 1 elsewhere, selected by the column's class; the upper byte of each half comes from
// the selector's sign-replication mode and is zero), one IMAD (an add of packed halves, on the FMA pipe) and one
// VIMNMX3.S16x2 that updates the row's value in place (in the gene's last column, where the vertical move is free, the
// row above enters with + 1 per half: a second copy of the loop, taken once per gene).  The diagonal term of the NEXT column, M = H(i-1, j) + inc(j + 1),
// is formed while H(i-1, j) is at hand (the gene letters are read one column ahead), so no state needs a second live
// copy and the loop has no register moves.  A gene letter that is none of A, C, G, T selects zero bytes (a diagonal move worth one less than a mismatch,
// never more than one below the true recurrence per such column) and 3 per such letter is added to the result -- its
// possible match and that one -- which keeps the bound a bound.
__device__ __forceinline__ unsigned ybound_selector(unsigned c) { return c < 4u ? 0x8480u + 0x0101u * c : 0x8888u; }   // bytes Af[c], 0, Ar[c], 0

template <int G, int C>
__global__ void __launch_bounds__(128, C > 16 ? 4 : (C > 11 ? 5 : 8))
ybound_kernel(const YParams P)
{
    constexpr int GROUPS_PER_WARP = 32 / G;
    const unsigned lane = threadIdx.x & 31u;
    const int l = lane % G, grp = lane / G;
    const unsigned one = (unsigned)P.one;
    const unsigned row0 = (unsigned)(l * C) * 0x00010001u;                  // frame offset of the row above this lane's first
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(P.counter, (unsigned)GROUPS_PER_WARP);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned)P.n_runs) break;
        const bool have = base + grp < (unsigned)P.n_runs;
        int nA = 0, nB = 1, nq = 0, out = 0, q = 0, steps = 0, frag = 0;
        const uint8_t *B = P.qcls;
        unsigned Af[C], Ar[C], H[C], M[C];
#pragma unroll
        for (int r = 0; r < C; ++r) Af[r] = Ar[r] = 0x01010101u;
        if (have) {
            const int4 run = P.runs[base + grp];
            const int a0 = P.tgt_off[2 * run.x], a1 = P.tgt_off[2 * run.x + 1];
            nA = a1 - a0;
            q = run.y; nq = run.z; out = run.w; frag = run.x;
            const int b0 = P.qry_off[q];
            nB = P.qry_off[q + 1] - b0;
            B += b0;
            steps = P.qry_off[q + nq] - b0 + G - 1;
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const int i = l * C + r;
                if (i < nA) {
                    const int cf = letter_class(P.tgt[a0 + i]), cr = letter_class(P.tgt[a1 + i]);
                    Af[r] = cf < 4 ? 0x01010101u + (2u << (8 * cf)) : 0x03030303u;   // other letters: taken to match every column
                    Ar[r] = cr < 4 ? 0x01010101u + (2u << (8 * cr)) : 0x03030303u;
                }
            }
        }
        unsigned c_cur = have ? B[0] : 0u, c_pref = have ? B[1] : 0u;
        {
            const unsigned sel = ybound_selector(c_cur);
#pragma unroll
            for (int r = 0; r < C; ++r) {
                H[r] = row0 + (unsigned)(r + 1) * 0x00010001u;                       // column 0: H = 0
                M[r] = row0 + (unsigned)r * 0x00010001u + prmt(Af[r], Ar[r], sel);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        unsigned bot = 0u, other = 0u;
        int j = 0;
        bool active = have && nA >= 1;
        for (int t = 0; t < steps; ++t) {
            unsigned up = __shfl_up_sync(0xffffffffu, bot, 1, G);
            if (active && t >= l) {
                ++j;
                if (l == 0) up = 0u;                                     // row 0: H(0, j) = 0, frame offset 0
                const bool lastcol = j == nB;
                const unsigned c_next = c_pref;                          // read one step ago: the load is off the critical path
                c_pref = B[j + 1];                                       // (the table has spare bytes at its end)
                const unsigned sel = ybound_selector(c_next);
                other += c_cur < 4u ? 0u : 0x00030003u;
                c_cur = c_next;
                unsigned uprev = up;
                if (!lastcol) {
#pragma unroll
                    for (int r = 0; r < C; ++r) {
                        H[r] = __vimax3_s16x2(M[r], uprev, H[r]);
                        M[r] = uprev * one + prmt(Af[r], Ar[r], sel);
                        uprev = H[r];
                    }
                } else {                                                 // the vertical move is free: + 1 per half in the frame
#pragma unroll
                    for (int r = 0; r < C; ++r) {
                        H[r] = __vimax3_s16x2(M[r], uprev + 0x00010001u, H[r]);
                        M[r] = uprev * one + prmt(Af[r], Ar[r], sel);
                        uprev = H[r];
                    }
                }
                bot = uprev;
                if (lastcol) {
                    if (l == G - 1) {
                        const unsigned y = uprev - (unsigned)(G * C) * 0x00010001u + other;      // out of the row frame
                        P.out[out] = y;
                        if (P.pass) {
                            // PI >= cut (>= 50) in either orientation needs 50 Y >= cut len + 100; the factor keeps the
                            // comparison on the safe side of the product's rounding
                            const double ymax = (double)max(y & 0xffffu, y >> 16);
                            if (50.0 * ymax >= (P.cut * (double)nB + 100.0) * (1.0 - 1e-12)) {
                                const unsigned k = atomicAdd(P.pass_count, 1u);
                                if (k < P.pass_cap) P.pass[k] = make_uint2((unsigned)frag, (unsigned)q);
                            }
                        }
                    }
                    ++out; ++q;
                    other = 0u;
                    if (--nq > 0) {
                        B += nB;
                        nB = P.qry_off[q + 1] - P.qry_off[q];
                        j = 0;
#pragma unroll
                        for (int r = 0; r < C; ++r) {
                            H[r] = row0 + (unsigned)(r + 1) * 0x00010001u;
                            M[r] = row0 + (unsigned)r * 0x00010001u + prmt(Af[r], Ar[r], sel);   // sel: the next gene's first letter
                        }
                    } else {
                        active = false;
                    }
                }
            }
        }
    }
}

} // namespace

int igd_launch_ub_topk(igd_handle *h, const uint16_t *d_lcs, const int32_t *d_qry_off, int n_unique, int n_qry, int K,
                       uint16_t *d_ub, int32_t *d_top)
{
    if (n_qry > 2048) { igd_set_error("igd_launch_ub_topk: more than 2048 genes"); return IGD_ERR_ARG; }
    ub_topk_kernel<<<(unsigned)(((size_t)n_unique * 32 + 127) / 128), 128, 0, h->stream>>>(d_lcs, d_qry_off, n_unique, n_qry, K, d_ub, d_top);
    IGD_CUDA(cudaGetLastError());
    h->timing.total_launches++; h->timing.align_launches++;
    return IGD_OK;
}

// out[t][g] = LCS(upper(target t), upper(gene g)) for the targets / genes resident in S_TGT / S_QRY (uint16, n_tgt x n_qry);
// max_q = longest gene (<= 512).  d_masks: 5 * W words per gene of scratch.
int igd_launch_lcs(igd_handle *h, const uint8_t *d_tgt, const int32_t *d_tgt_off, int n_tgt, size_t tbytes,
                   const uint8_t *d_qry, const int32_t *d_qry_off, int n_qry, int max_q, uint32_t *d_masks, uint16_t *d_out)
{
    if (max_q > 512) { igd_set_error("igd_launch_lcs: gene longer than 512"); return IGD_ERR_TOO_LONG; }
    const int W = max_q <= 352 ? 11 : 16;
    // class bytes of the target letters, once, instead of a classification per thread and row
    int rc;
    if ((rc = igd_reserve(h, S_TCLS, tbytes + 16))) return rc;
    if ((rc = igd_launch_gene_classes(h, d_tgt, tbytes, (uint8_t *)h->d_scratch[S_TCLS]))) return rc;
    d_tgt = (const uint8_t *)h->d_scratch[S_TCLS];
    gene_masks_kernel<<<(n_qry * W + LCS_THREADS - 1) / LCS_THREADS, LCS_THREADS, 0, h->stream>>>(d_qry, d_qry_off, n_qry, W, d_masks);
    IGD_CUDA(cudaGetLastError());
    const int gene_blocks = (n_qry + LCS_THREADS - 1) / LCS_THREADS;
    // enough CTAs for every SM to hold several, few enough that the masks are loaded for more than one target
    int tgt_blocks = std::max(1, std::min(n_tgt, (h->sm_count * 8 + gene_blocks - 1) / gene_blocks));
    const int per = (n_tgt + tgt_blocks - 1) / tgt_blocks;
    tgt_blocks = (n_tgt + per - 1) / per;
    const dim3 grid((unsigned)gene_blocks, (unsigned)tgt_blocks);
    if (W == 11) lcs_bound_kernel<11><<<grid, LCS_THREADS, 0, h->stream>>>(d_tgt, d_tgt_off, n_tgt, per, d_qry_off, n_qry, d_masks, d_out);
    else         lcs_bound_kernel<16><<<grid, LCS_THREADS, 0, h->stream>>>(d_tgt, d_tgt_off, n_tgt, per, d_qry_off, n_qry, d_masks, d_out);
    IGD_CUDA(cudaGetLastError());
    h->timing.total_launches += 2;
    h->timing.align_launches += 2;
    return IGD_OK;
}

int igd_launch_gene_classes(igd_handle *h, const uint8_t *d_qry, size_t n, uint8_t *d_cls)
{
    if (n == 0) return IGD_OK;
    gene_class_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8), 256, 0, h->stream>>>(d_qry, n, d_cls);
    IGD_CUDA(cudaGetLastError());
    h->timing.total_launches++; h->timing.align_launches++;
    return IGD_OK;
}

// out[slot + k] = Y(rc) << 16 | Y(forward) for fragment u against gene q0 + k of every run {u, q0, count, slot};
// shape 0: fragments of at most 352 letters (16 lanes x 22 rows), 1: at most 512 (32 x 16).  *d_counter must be zero.
int igd_launch_ybound(igd_handle *h, int shape, const uint8_t *d_tgt, const int32_t *d_tgt_off, const uint8_t *d_qcls,
                      const int32_t *d_qry_off, const int4 *d_runs, int n_runs, uint32_t *d_out, unsigned int *d_counter,
                      double cut, uint2 *d_pass, unsigned int *d_pass_count, unsigned int pass_cap)
{
    if (n_runs <= 0) return IGD_OK;
    YParams P;
    P.tgt = d_tgt; P.tgt_off = d_tgt_off; P.qcls = d_qcls; P.qry_off = d_qry_off; P.runs = d_runs; P.n_runs = n_runs;
    P.out = d_out; P.counter = d_counter;
    P.cut = cut; P.pass = d_pass; P.pass_count = d_pass_count; P.pass_cap = pass_cap; P.one = 1;
    const int groups_per_cta = shape == 0 ? 128 / 16 : 128 / 32;
    const unsigned blocks = (unsigned)std::max(1, std::min((n_runs + groups_per_cta - 1) / groups_per_cta, h->sm_count * (shape == 0 ? 4 : 5)));
    // 32 lanes x 11 rows for the 350-letter fragments (64 registers, 8 CTAs per SM, half the dependent chain per column):
    // 1.18 against 1.04 ms on the bench step -- the per-column overhead is paid twice as often
    if (shape == 0) ybound_kernel<16, 22><<<blocks, 128, 0, h->stream>>>(P);
    else            ybound_kernel<32, 16><<<blocks, 128, 0, h->stream>>>(P);
    IGD_CUDA(cudaGetLastError());
    h->timing.total_launches++; h->timing.align_launches++;
    return IGD_OK;
}

// C ABI (include/igd_b200.h): both bounds for every (fragment, gene) pair -- what igd_detect's pruning uses, exposed so
// that the tests can hold the kernels against a plain restatement of the two recurrences.
extern "C" int igd_match_bounds(igd_handle *h, const uint8_t *tgt, const int32_t *tgt_off, int32_t n_frag,
                                const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                                uint16_t *lcs, uint32_t *ybound)
{
    if (!h || !tgt_off || !qry_off || n_frag < 0 || n_qry < 0 || (n_frag && n_qry && (!tgt || !qry || !lcs || !ybound))) {
        igd_set_error("bad arguments to igd_match_bounds"); return IGD_ERR_ARG;
    }
    if (n_frag == 0 || n_qry == 0) return IGD_OK;
    int max_q = 0;
    for (int q = 0; q < n_qry; ++q) {
        const int n = qry_off[q + 1] - qry_off[q];
        if (n < 1 || n > 511) { igd_set_error("igd_match_bounds: gene lengths must lie in 1..511"); return IGD_ERR_TOO_LONG; }
        max_q = std::max(max_q, n);
    }
    std::vector<int4> runs[2];
    for (int u = 0; u < n_frag; ++u) {
        const int nA = tgt_off[2 * u + 1] - tgt_off[2 * u];
        if (nA != tgt_off[2 * u + 2] - tgt_off[2 * u + 1]) { igd_set_error("igd_match_bounds: rows 2u and 2u+1 must have the same length"); return IGD_ERR_ARG; }
        if (nA < 1 || nA > 512) { igd_set_error("igd_match_bounds: fragment lengths must lie in 1..512"); return IGD_ERR_TOO_LONG; }
        for (int g = 0; g < n_qry; g += 64) runs[nA <= 352 ? 0 : 1].push_back(make_int4(u, g, std::min(64, n_qry - g), u * n_qry + g));
    }
    IGD_CUDA(cudaSetDevice(h->device));
    const size_t tbytes = (size_t)tgt_off[2 * n_frag], qbytes = (size_t)qry_off[n_qry], cells = (size_t)n_frag * n_qry;
    std::vector<int4> all = runs[0];
    all.insert(all.end(), runs[1].begin(), runs[1].end());
    int rc;
    if ((rc = igd_reserve(h, S_TGT, tbytes))) return rc;
    if ((rc = igd_reserve(h, S_TGTOFF, ((size_t)2 * n_frag + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_QRY, qbytes))) return rc;
    if ((rc = igd_reserve(h, S_QRYOFF, ((size_t)n_qry + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_QCLS, qbytes + 16))) return rc;
    if ((rc = igd_reserve(h, S_MASKS, (size_t)n_qry * 5 * 16 * 4))) return rc;
    if ((rc = igd_reserve(h, S_LCS, 2 * cells * 2))) return rc;
    if ((rc = igd_reserve(h, S_YRUNS, all.size() * sizeof(int4)))) return rc;
    if ((rc = igd_reserve(h, S_Y, cells * 4))) return rc;
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGT], tgt, tbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGTOFF], tgt_off, ((size_t)2 * n_frag + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRY], qry, qbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRYOFF], qry_off, ((size_t)n_qry + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_YRUNS], all.data(), all.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemsetAsync(h->d_count + 32, 0, 2 * sizeof(unsigned long long), h->stream));
    if ((rc = igd_launch_lcs(h, (const uint8_t *)h->d_scratch[S_TGT], (const int32_t *)h->d_scratch[S_TGTOFF], 2 * n_frag, tbytes,
                             (const uint8_t *)h->d_scratch[S_QRY], (const int32_t *)h->d_scratch[S_QRYOFF], n_qry, max_q,
                             (uint32_t *)h->d_scratch[S_MASKS], (uint16_t *)h->d_scratch[S_LCS]))) return rc;
    if ((rc = igd_launch_gene_classes(h, (const uint8_t *)h->d_scratch[S_QRY], qbytes, (uint8_t *)h->d_scratch[S_QCLS]))) return rc;
    for (int shape = 0; shape < 2; ++shape)
        if ((rc = igd_launch_ybound(h, shape, (const uint8_t *)h->d_scratch[S_TGT], (const int32_t *)h->d_scratch[S_TGTOFF],
                                    (const uint8_t *)h->d_scratch[S_QCLS], (const int32_t *)h->d_scratch[S_QRYOFF],
                                    (const int4 *)h->d_scratch[S_YRUNS] + (shape ? runs[0].size() : 0), (int)runs[shape].size(),
                                    (uint32_t *)h->d_scratch[S_Y], (unsigned int *)(h->d_count + 32 + shape), 0.0, nullptr, nullptr, 0))) return rc;
    IGD_CUDA(cudaMemcpyAsync(lcs, h->d_scratch[S_LCS], 2 * cells * 2, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaMemcpyAsync(ybound, h->d_scratch[S_Y], cells * 4, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    return IGD_OK;
}
