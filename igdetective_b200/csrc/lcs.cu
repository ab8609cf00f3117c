// K4: an upper bound on BioAlign.NumMatches() for every (target, gene) pair, used by igd_detect to skip alignments that
// cannot matter.
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

namespace {

constexpr int LCS_THREADS = 128;

__device__ __forceinline__ int letter_class(unsigned c)
{
    if (c - 'a' < 26u) c -= 32u;
    return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : 4)));
}

// masks[g][class][w]: bit i of word w = upper(gene g)[32 w + i] is in the class
__global__ void __launch_bounds__(LCS_THREADS)
gene_masks_kernel(const uint8_t *__restrict__ qry, const int32_t *__restrict__ qry_off, int n_qry, int W, uint32_t *__restrict__ masks)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_qry) return;
    const int b0 = qry_off[g], n = qry_off[g + 1] - b0;
    uint32_t *m = masks + (size_t)g * 5 * W;
    for (int k = 0; k < 5 * W; ++k) m[k] = 0u;
    for (int i = 0; i < n && i < 32 * W; ++i) m[letter_class(qry[b0 + i]) * W + (i >> 5)] |= 1u << (i & 31);
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

template <int W>
__global__ void __launch_bounds__(LCS_THREADS)
lcs_bound_kernel(const uint8_t *__restrict__ tgt, const int32_t *__restrict__ tgt_off, int n_tgt, int tgt_per_block,
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
            const int c = letter_class(__ldg(&tgt[a0 + i]));                    // the same letter for every thread
            uint32_t U[W], S[W];
#define IGD_LCS_ROW(C)                                                                                   \
            {                                                                                            \
                _Pragma("unroll") for (int w = 0; w < W; ++w) U[w] = V[w] & M[C][w];                      \
                add_words<W>(V, U, S);                                                                   \
                _Pragma("unroll") for (int w = 0; w < W; ++w) V[w] = S[w] | (V[w] ^ U[w]);               \
            }
            switch (c) {
            case 0: IGD_LCS_ROW(0) break;
            case 1: IGD_LCS_ROW(1) break;
            case 2: IGD_LCS_ROW(2) break;
            case 3: IGD_LCS_ROW(3) break;
            default: IGD_LCS_ROW(4) break;
            }
#undef IGD_LCS_ROW
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
int igd_launch_lcs(igd_handle *h, const uint8_t *d_tgt, const int32_t *d_tgt_off, int n_tgt,
                   const uint8_t *d_qry, const int32_t *d_qry_off, int n_qry, int max_q, uint32_t *d_masks, uint16_t *d_out)
{
    if (max_q > 512) { igd_set_error("igd_launch_lcs: gene longer than 512"); return IGD_ERR_TOO_LONG; }
    const int W = max_q <= 352 ? 11 : 16;
    gene_masks_kernel<<<(n_qry + LCS_THREADS - 1) / LCS_THREADS, LCS_THREADS, 0, h->stream>>>(d_qry, d_qry_off, n_qry, W, d_masks);
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
