// K0: ASCII contigs -> 2-bit planes + validity plane (layout in common.cuh).
//
// Replaces, for the scan, the per-window `locus[i:i+k].upper()` of find_valid_motif_idx
// (py/IGDetective.py:140-143) and the str(seq.reverse_complement()) copies of
// get_contigwise_rss (:187-190): case is folded and non-ACGT letters are marked once here;
// the reverse strand is never materialised.
//
// HBM-bound: 1 B read + 0.375 B written per base.  A warp packs 32 consecutive chunks (one word of the
// summary bitmap) in four steps of 8 chunks; in a step every lane owns 16 consecutive letters, fetched
// by two aligned 16-byte loads (the contigs sit back to back in the letter buffer, so a chunk's letters
// start at any byte: the 16 wanted bytes are cut out of the 32 loaded ones by funnel shifts), classified
// through a 256-entry table in shared memory whose entries hold the low / high code bit and the
// "not ACGT" bit in three byte-wide fields, and accumulated with one shift-add per letter; four lanes'
// results are joined into the chunk's record by shuffles.  All eight loads of a group are issued before the
// arithmetic starts, and the contig of a group comes from a host-made table (one entry per tile of 512 chunks)
// instead of a binary search, so no warp sits in a chain of dependent loads.
#include "common.cuh"

namespace {

constexpr int PACK_THREADS = 256;

__device__ __forceinline__ uint4 ldg_nc_128(const uint8_t *p) { return __ldg((const uint4 *)p); }

__global__ void __launch_bounds__(PACK_THREADS)
pack_kernel(const uint8_t *__restrict__ seq, const unsigned long long *__restrict__ seq_off,
            const unsigned long long *__restrict__ chunk0, const int *__restrict__ tile_contig, int n_contigs,
            unsigned long long n_chunks, uint4 *__restrict__ planes,
            unsigned long long *__restrict__ inv, uint32_t *__restrict__ invsum)
{
    // entry: bit 0 = low code bit, bit 8 = high code bit, bit 16 = letter is not A/C/G/T in either case
    // (codes A0 C1 G2 T3)
    __shared__ uint32_t lut[256];
    for (int i = threadIdx.x; i < 256; i += PACK_THREADS) {
        const int c = i & 0xDF;
        uint32_t e = 1u << 16;
        if (c == 'A') e = 0u;
        else if (c == 'C') e = 1u;
        else if (c == 'G') e = 1u << 8;
        else if (c == 'T') e = 1u | (1u << 8);
        lut[i] = e;
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, sub = lane & 3u;
    const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long ngroups = (n_chunks + 31) / 32;

    for (unsigned long long g = warp; g < ngroups; g += nwarps) {
        // contig holding the first chunk of the group's tile (host-made table, one entry per 512 chunks), then forward
        int c = __ldg(&tile_contig[g >> 4]);
        // ---- all loads of the group first (8 x 16 bytes per lane in flight), then the arithmetic -------------
        uint4 A[4], B[4];
        int nv[4];
        unsigned kk[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const unsigned long long chunk = g * 32 + (unsigned)it * 8u + (lane >> 2);
            while (c + 1 < n_contigs && __ldg(&chunk0[c + 1]) <= chunk) ++c;
            nv[it] = 0; kk[it] = 0;
            A[it] = make_uint4(0, 0, 0, 0); B[it] = make_uint4(0, 0, 0, 0);
            if (c >= 0 && chunk < n_chunks) {
                const unsigned long long base = (chunk - __ldg(&chunk0[c])) * 64ull + 16u * sub;
                const unsigned long long o0 = __ldg(&seq_off[c]), len = __ldg(&seq_off[c + 1]) - o0;
                if (base < len) {
                    nv[it] = (int)min(16ull, len - base);
                    const uint8_t *src = seq + o0 + base;
                    const uint8_t *a0 = (const uint8_t *)((unsigned long long)src & ~15ull);
                    kk[it] = (unsigned)((unsigned long long)src & 15ull);
                    A[it] = ldg_nc_128(a0);
                    // the second half is wanted only when the letters start inside the first one (and the last
                    // contig may end in the buffer's last 16 bytes)
                    if (kk[it]) B[it] = ldg_nc_128(a0 + 16);
                }
            }
        }
        unsigned summary = 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const unsigned long long chunk = g * 32 + (unsigned)it * 8u + (lane >> 2);
            const int nvalid = nv[it];
            unsigned acc0 = 0, acc1 = 0;                       // letters 0..7 and 8..15: three 8-bit fields each
            if (nvalid > 0) {
                const unsigned k = kk[it];
                unsigned w0 = A[it].x, w1 = A[it].y, w2 = A[it].z, w3 = A[it].w, w4 = B[it].x, w5 = B[it].y, w6 = B[it].z, w7 = B[it].w;
                if (k & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = w7; }
                if (k & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
                const unsigned sh = (k & 3u) * 8u;
                const unsigned v[4] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                       __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh)};
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    acc0 += lut[(v[b >> 2] >> (8 * (b & 3))) & 255u] << b;
                    acc1 += lut[(v[2 + (b >> 2)] >> (8 * (b & 3))) & 255u] << b;
                }
            }
            // positions past the contig's end: code 0, marked invalid
            const unsigned keep = (1u << nvalid) - 1u;              // nvalid <= 16
            const unsigned lo16 = ((acc0 & 0xffu) | ((acc1 & 0xffu) << 8)) & keep;
            const unsigned hi16 = (((acc0 >> 8) & 0xffu) | (acc1 & 0xff00u)) & keep;
            const unsigned bad16 = ((((acc0 >> 16) & 0xffu) | ((acc1 >> 8) & 0xff00u)) & keep) | (~keep & 0xffffu);
            // four lanes -> one record {lo0, hi0, lo1, hi1} and one validity word
            const unsigned lh = lo16 | (hi16 << 16);
            const unsigned lh1 = __shfl_down_sync(0xffffffffu, lh, 1);
            const unsigned bd1 = __shfl_down_sync(0xffffffffu, bad16, 1);
            const unsigned lo32 = (lh & 0xffffu) | (lh1 << 16);
            const unsigned hi32 = (lh >> 16) | (lh1 & 0xffff0000u);
            const unsigned bad32 = bad16 | (bd1 << 16);
            const unsigned lo32b = __shfl_down_sync(0xffffffffu, lo32, 2);
            const unsigned hi32b = __shfl_down_sync(0xffffffffu, hi32, 2);
            const unsigned bad32b = __shfl_down_sync(0xffffffffu, bad32, 2);
            const bool any_bad = (bad32 | bad32b) != 0u;
            if (sub == 0 && chunk < n_chunks) {
                planes[chunk] = make_uint4(lo32, hi32, lo32b, hi32b);
                inv[chunk] = (unsigned long long)bad32 | ((unsigned long long)bad32b << 32);
            }
            const unsigned m = __ballot_sync(0xffffffffu, sub == 0 && any_bad && chunk < n_chunks);   // bits 0, 4, .., 28
            // one bit per chunk: gather every fourth bit
            unsigned byte = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) byte |= ((m >> (4 * q)) & 1u) << q;
            summary |= byte << (8 * it);
        }
        if (lane == 0) invsum[g] = summary;
    }
}

} // namespace

int igd_launch_pack(igd_handle *h, const uint8_t *d_seq, const unsigned long long *d_seq_off,
                    const unsigned long long *d_chunk0, const int *d_tile_contig, int n_contigs)
{
    const unsigned long long groups = (h->n_chunks + 31) / 32;
    unsigned long long blocks = (groups * 32 + PACK_THREADS - 1) / PACK_THREADS;
    const unsigned long long cap = (unsigned long long)h->sm_count * 8;          // resident CTAs of 256 threads per SM (registers permitting)
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pack_kernel<<<(unsigned)blocks, PACK_THREADS, 0, h->stream>>>(d_seq, d_seq_off, d_chunk0, d_tile_contig, n_contigs,
                                                                  h->n_chunks, h->d_planes, h->d_inv, h->d_invsum);
    h->timing.total_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}
