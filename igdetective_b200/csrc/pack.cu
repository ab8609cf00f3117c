// K0: ASCII contigs -> 2-bit planes + validity plane (layout in common.cuh).
//
// Replaces, for the scan, the per-window `locus[i:i+k].upper()` of find_valid_motif_idx
// (py/IGDetective.py:140-143) and the str(seq.reverse_complement()) copies of
// get_contigwise_rss (:187-190): case is folded and non-ACGT letters are marked once here;
// the reverse strand is never materialised.
#include "common.cuh"

namespace {

__device__ __forceinline__ void classify(unsigned char ch, bool in_range, unsigned &code, bool &bad)
{
    const unsigned c = ch & 0xDFu;                       // fold case
    const bool ok = in_range && (c == 'A' || c == 'C' || c == 'G' || c == 'T');
    const unsigned x = (c >> 1) & 3u;                    // A0 C1 G3 T2
    code = ok ? (x ^ (x >> 1)) : 0u;                     // A0 C1 G2 T3
    bad = !ok;
}

// One warp packs 32 consecutive chunks (so it owns one word of the summary bitmap).
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ seq, const unsigned long long *__restrict__ seq_off,
            const unsigned long long *__restrict__ chunk0, int n_contigs,
            unsigned long long n_chunks, uint4 *__restrict__ planes,
            unsigned long long *__restrict__ inv, uint32_t *__restrict__ invsum)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long ngroups = (n_chunks + 31) / 32;

    for (unsigned long long g = warp; g < ngroups; g += nwarps) {
        unsigned summary = 0;
        // contig holding the first chunk of the group: last c with chunk0[c] <= chunk
        unsigned long long chunk = g * 32;
        int lo = 0, hi = n_contigs;                      // first c with chunk0[c] > chunk
        while (lo < hi) { int mid = (lo + hi) >> 1; if (chunk0[mid] <= chunk) lo = mid + 1; else hi = mid; }
        int c = lo - 1;
        for (int i = 0; i < 32; ++i, ++chunk) {
            if (chunk >= n_chunks) break;
            while (c + 1 < n_contigs && chunk0[c + 1] <= chunk) ++c;
            unsigned long long base_in_contig = 0, len = 0, src = 0;
            if (c >= 0) {
                base_in_contig = (chunk - chunk0[c]) * 64ull;
                len = seq_off[c + 1] - seq_off[c];
                src = seq_off[c] + base_in_contig;
            }
            unsigned w[2][3];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const unsigned long long p = base_in_contig + half * 32 + lane;
                const bool in_range = (c >= 0) && p < len;
                const unsigned char ch = in_range ? seq[src + half * 32 + lane] : 0;
                unsigned code; bool bad;
                classify(ch, in_range, code, bad);
                w[half][0] = __ballot_sync(0xffffffffu, code & 1u);
                w[half][1] = __ballot_sync(0xffffffffu, code & 2u);
                w[half][2] = __ballot_sync(0xffffffffu, bad);
            }
            if (lane == 0) {
                planes[chunk] = make_uint4(w[0][0], w[0][1], w[1][0], w[1][1]);
                inv[chunk] = (unsigned long long)w[0][2] | ((unsigned long long)w[1][2] << 32);
            }
            if (w[0][2] | w[1][2]) summary |= 1u << i;
        }
        if (lane == 0) invsum[g] = summary;
    }
}

} // namespace

int igd_launch_pack(igd_handle *h, const uint8_t *d_seq, const unsigned long long *d_seq_off,
                    const unsigned long long *d_chunk0, int n_contigs)
{
    const int threads = 256;
    const unsigned long long groups = (h->n_chunks + 31) / 32;
    unsigned long long blocks = (groups * 32 + threads - 1) / threads;
    const unsigned long long cap = (unsigned long long)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pack_kernel<<<(unsigned)blocks, threads, 0, h->stream>>>(d_seq, d_seq_off, d_chunk0, n_contigs,
                                                            h->n_chunks, h->d_planes, h->d_inv, h->d_invsum);
    h->timing.total_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}
