// igd_detect: the module body of py/IGDetective.py:439-487 in one call on the contigs resident in the handle.
//
//   RSS scan (rss_scan.cu)  ->  hit list decoded and radix-sorted on the device (hits.cu), 24 bytes per hit to the host
//   combine_D_RSS (:210-224) -> two pointers over the D_left / D_right hits of a contig and strand, host
//   get_s_fragment_from_RSS (:268-300) -> gather_fragments_kernel over the resident ASCII letters: Python slice
//       semantics incl. the negative-start quirk of parent_seq[index-length:index] (:265), upper-cased (:308),
//       reverse-complemented for strand '-' with Bio.Seq's IUPAC table; each fragment and its reverse complement
//       become targets 2u / 2u+1
//   align_fragment_to_genes (:319-360) + np.argmax's first maximum (:473): the reference aligns every fragment with every
//       gene in both orientations; here (best_by_pruning) two upper bounds on PI (lcs.cu: K4, the LCS; K5, a one-state DP
//       that charges insertions) prove which of those alignments cannot decide a row, and the packed Gotoh kernels run on
//       the rest in three rounds -- the same rows from ~2 % of the cells.  Tables of fewer than 64 genes (J) and
//       IGD_DETECT_PRUNE=0 take the all-pairs path: packed kernels over [2U targets] x [all genes], the orientation
//       choice of ComputeAlignment (:314-317) and the first-maximum argmax on the device.
//   threshold (:387) on the host in float64, the winners' ComputeAlignment (:388) through the two-register
//       Gotoh kernel for AlignmentRange()[0] and the end of QuerySeq()
//   the rows' letters (:391-419): gather_text_kernel, fetched with igd_fetch_row_text
// Environment switches, read per call (tests compare the paths): IGD_DETECT_PRUNE=0, IGD_DETECT_YBOUND=0 (K4 only),
// IGD_DETECT_YPASS_CAP (size of K5's verdict list), IGD_DETECT_TOP_MIN / IGD_DETECT_TOP_DELTA (round one's lists),
// IGD_DETECT_TRACE=1 (phase times on stderr).
#include "common.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <functional>
#include <unordered_map>

namespace {

// IGD_DETECT_TRACE=1: host-clock time of every phase of igd_detect on stderr (profiles/step_profile.py)
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("IGD_DETECT_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what)
    {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[igd_detect] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count());
        t0 = t;
    }
};

struct FragDesc { unsigned long long g0; int32_t len; int32_t rc; };   // letters [g0, g0+len) of the letter buffer

// target 2u = upper(fragment u) (reverse-complemented when rc), target 2u+1 = its reverse complement
__global__ void __launch_bounds__(128)
gather_fragments_kernel(const uint8_t *__restrict__ seq, const FragDesc *__restrict__ fr, const int32_t *__restrict__ toff,
                        const uint8_t *__restrict__ comp, int n, uint8_t *__restrict__ tgt)
{
    __shared__ uint8_t s_comp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_comp[i] = comp[i];
    __syncthreads();
    for (int u = blockIdx.x; u < n; u += gridDim.x) {
        const FragDesc f = fr[u];
        uint8_t *F = tgt + toff[2 * u], *R = tgt + toff[2 * u + 1];
        for (int k = threadIdx.x; k < f.len; k += blockDim.x) {
            unsigned c = seq[f.g0 + (unsigned long long)(f.rc ? f.len - 1 - k : k)];
            if (c - 'a' < 26u) c -= 32u;
            if (f.rc) c = s_comp[c];
            F[k] = (uint8_t)c;
            R[f.len - 1 - k] = s_comp[c];
        }
    }
}

// The letters of the emitted rows (py/IGDetective.py:391-419, 238-249): segment k = upper(letters [g0, g0 + len)), reverse-
// complemented when rc, written at out[off[k] ...).  Same letter handling as the fragment gather above.
struct TextDesc { unsigned long long g0; unsigned long long off; int32_t len; int32_t rc; };

__global__ void __launch_bounds__(128)
gather_text_kernel(const uint8_t *__restrict__ seq, const TextDesc *__restrict__ d, const uint8_t *__restrict__ comp, int n,
                   uint8_t *__restrict__ out)
{
    __shared__ uint8_t s_comp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_comp[i] = comp[i];
    __syncthreads();
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const TextDesc t = d[k];
        for (int i = threadIdx.x; i < t.len; i += blockDim.x) {
            unsigned c = seq[t.g0 + (unsigned long long)(t.rc ? t.len - 1 - i : i)];
            if (c - 'a' < 26u) c -= 32u;
            if (t.rc) c = s_comp[c];
            out[t.off + i] = (uint8_t)c;
        }
    }
}

// np.argmax(pi_mat[k]) (py/IGDetective.py:473): FIRST maximum of PI = (matches-1)/len*100 over the genes, compared
// exactly by cross-multiplication (equal rationals give equal float64 PI, different ones differ by far more than
// an ulp).  pick[u][g] = matches | len << 10 | forward << 31 (pick_orientation_kernel).  One warp per fragment.
__global__ void __launch_bounds__(256)
argmax_pick_kernel(const uint32_t *__restrict__ pick, int n_unique, int n_qry, int4 *__restrict__ out)
{
    const int u = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n_unique) return;
    long long bn = 0, bd = 1; int bi = -1; uint32_t bw = 0;
    for (int g = lane; g < n_qry; g += 32) {
        const uint32_t w = pick[(size_t)u * n_qry + g];
        const long long num = (long long)(int)(w & 1023u) - 1, den = (long long)((w >> 10) & 0x1fffffu);
        if (bi < 0 || num * bd > bn * den) { bn = num; bd = den; bi = g; bw = w; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long on = __shfl_xor_sync(0xffffffffu, bn, o), od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const uint32_t ow = __shfl_xor_sync(0xffffffffu, bw, o);
        if (oi >= 0) {
            const long long l = on * bd, r = bn * od;
            if (bi < 0 || l > r || (l == r && oi < bi)) { bn = on; bd = od; bi = oi; bw = ow; }
        }
    }
    if (lane == 0) out[u] = make_int4(bi, (int)(bw & 1023u), (int)((bw >> 10) & 0x1fffffu), (int)(bw >> 31));
}

// parent_seq[a:b] with Python's slice semantics on a sequence of length L -> (start, length)
inline void py_slice(int64_t a, int64_t b, int64_t L, int64_t &s, int64_t &n)
{
    if (a < 0) { a += L; if (a < 0) a = 0; } else if (a > L) a = L;
    if (b < 0) { b += L; if (b < 0) b = 0; } else if (b > L) b = L;
    s = a; n = b > a ? b - a : 0;
}

struct Best { int32_t gene, matches, alen, fwd; };

struct SliceKey {
    unsigned long long g0; int32_t len, rc;
    bool operator==(const SliceKey &o) const { return g0 == o.g0 && len == o.len && rc == o.rc; }
};
struct SliceHash {
    size_t operator()(const SliceKey &k) const { return (size_t)(k.g0 * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.len << 1) ^ (size_t)k.rc; }
};
struct Acc { float ms = 0.f; uint64_t cells = 0; unsigned launches = 0; float lcs_ms = 0.f, k5_ms = 0.f; uint64_t lcs_cells = 0, k5_cells = 0; };

// the 'A' * len(gene) stand-in of an empty fragment (py/IGDetective.py:339-340), per gene: pick word as in
// pick_orientation_kernel (host pair-table path: a few hundred short pairs, and only when an empty fragment exists)
int empty_fragment_best(igd_handle *h, const uint8_t *gene, const int32_t *gene_off, int32_t n_gene, const igd_scoring *sc, Best *best)
{
    std::unordered_map<int, int32_t> by_len;
    std::vector<uint8_t> et; std::vector<int32_t> eoff(1, 0), pt, pq;
    for (int32_t g = 0; g < n_gene; ++g) {
        const int n = gene_off[g + 1] - gene_off[g];
        auto it = by_len.find(n);
        if (it == by_len.end()) {
            it = by_len.emplace(n, (int32_t)by_len.size()).first;
            et.insert(et.end(), (size_t)n, (uint8_t)'A'); eoff.push_back((int32_t)et.size());
            et.insert(et.end(), (size_t)n, (uint8_t)'T'); eoff.push_back((int32_t)et.size());
        }
        pt.push_back(2 * it->second); pq.push_back(g);
        pt.push_back(2 * it->second + 1); pq.push_back(g);
    }
    std::vector<int32_t> s_(pt.size()), m_(pt.size()), l_(pt.size());
    int rc = igd_align_batch(h, et.data(), eoff.data(), (int32_t)eoff.size() - 1, gene, gene_off, n_gene,
                             pt.data(), pq.data(), (int64_t)pt.size(), sc, s_.data(), m_.data(), l_.data(), nullptr, nullptr);
    if (rc) return rc;
    long long bn = 0, bd = 1; best->gene = -1;
    for (int32_t g = 0; g < n_gene; ++g) {
        const bool fwd = m_[2 * (size_t)g] > m_[2 * (size_t)g + 1];
        const size_t w = 2 * (size_t)g + (fwd ? 0 : 1);
        const long long num = m_[w] - 1, den = l_[w];
        if (best->gene < 0 || num * bd > bn * den) { bn = num; bd = den; best->gene = g; best->matches = m_[w]; best->alen = l_[w]; best->fwd = fwd; }
    }
    return IGD_OK;
}


// ---- exact pruning of the scoring ---------------------------------------------------------------
// What the stage needs per fragment is np.argmax over the genes of PI (first maximum, py/IGDetective.py:473) and that
// gene's (PI, length) for the cut-off test (:387) -- not the whole PI matrix.  With the bound of lcs.cu,
//     PI(f, g) <= UB(f, g) = (max(LCS(f, g), LCS(rc(f), g)) - 1) / len(g) * 100     for both orientations,
// a gene whose UB is below the best PI found so far can neither be the first maximum nor tie with it, and a fragment
// all of whose UBs are below the lowest cut-off yields no row whatever its argmax is.  Two rounds of the packed Gotoh
// kernel are therefore enough: the sixteen genes with the largest UB of every fragment, then every gene with
// UB >= thr(f), thr = the best PI of round one, or the cut-off when that best lies below it.  The threshold can only
// rise afterwards, so nothing else is ever needed.  Rationals are compared exactly by cross-multiplication; the cut-off
// test uses the same float64 expression as the reference (monotone in the rational, so UB bounds it too).  Rows are
// identical to the unpruned path (tests/test_gpu_pipeline.py::test_pruned_scoring_equals_full_scoring); planted genes
// need 1-15 of 490 alignments per orientation, background candidates (best PI ~ 54 < 60 <= UB ~ 69) need them all.
constexpr size_t PRUNE_TOP = 16;
constexpr int PRUNE_MIN = 8;            // swept on the bench contig (min 2 / 4 / 8 x delta 0.015 / 0.03 / 0.06 / off): 8 and 0.03 cost least
constexpr double PRUNE_DELTA = 0.03;
struct PrunedPair { int32_t u, q0, count, slot; };           // fwd results at slot + k, reverse-complement at slot + count + k

// DP over the listed (fragment, gene) pairs, both orientations; per fragment `lists[u]` must be ascending
// after_launch (optional): more work for the stream, queued behind the round's kernels and finished by the round's
// synchronisation; extra_stage: pinned bytes it may fill behind the round's own read-back (its argument)
static int prune_round(igd_handle *h, const std::vector<std::vector<int32_t>> &lists, const std::vector<int32_t> &toff,
                       const int32_t *gene_off, int32_t n_gene, const igd_scoring *sc, int mixed, Acc &acc,
                       std::vector<Best> &best, std::vector<uint8_t> &done,
                       size_t extra_stage = 0, const std::function<int(uint8_t *)> &after_launch = nullptr)
{
    constexpr int NB = 6;
    const int32_t U = (int32_t)lists.size();
    size_t n_pairs = 0;
    for (const auto &l : lists) n_pairs += 2 * l.size();
    if (n_pairs == 0 && !after_launch) return IGD_OK;
    const size_t slots = (size_t)h->sm_count * 16;
    const int chunk = (int)std::min<size_t>(64, std::max<size_t>(1, n_pairs / (slots * 16)));
    std::vector<int4> runs[NB];
    std::vector<PrunedPair> pp;
    size_t slot = 0;
    uint64_t cells = 0;
    for (int32_t u = 0; u < U; ++u) {
        const std::vector<int32_t> &l = lists[(size_t)u];
        const int nA = toff[2 * u + 1] - toff[2 * u];
        const int b = igd_align_bucket_of(nA);
        for (size_t i = 0; i < l.size();) {
            int n = 1;
            while (i + n < l.size() && n < chunk && l[i + n] == l[i] + n) ++n;
            runs[b].push_back(make_int4(2 * u, l[i], n, (int)slot));
            runs[b].push_back(make_int4(2 * u + 1, l[i], n, (int)(slot + n)));
            pp.push_back({u, l[i], n, (int32_t)slot});
            for (int k = 0; k < n; ++k) cells += 2ull * (uint64_t)nA * (uint64_t)(gene_off[l[i] + k + 1] - gene_off[l[i] + k]);
            slot += 2 * (size_t)n;
            i += n;
        }
    }
    int rc;
    std::vector<int4> all;
    size_t roff[NB + 1];
    for (int b = 0; b < NB; ++b) { roff[b] = all.size(); all.insert(all.end(), runs[b].begin(), runs[b].end()); }
    roff[NB] = all.size();
    if ((rc = igd_reserve(h, S_RUNS, all.size() * sizeof(int4)))) return rc;
    if ((rc = igd_reserve(h, S_SCORE, n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_PAY, n_pairs * 4))) return rc;
    if ((rc = igd_stage_reserve(h, n_pairs * 4 + extra_stage))) return rc;
    if (!all.empty()) {
        if ((rc = igd_pin_reserve(h, 1, all.size() * sizeof(int4)))) return rc;
        memcpy(h->h_pin[1], all.data(), all.size() * sizeof(int4));       // consumed before the round's synchronisation
        IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_RUNS], h->h_pin[1], all.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    }
    IGD_CUDA(cudaMemsetAsync(h->d_count + 8, 0, 3 * NB * sizeof(unsigned long long), h->stream));
    h->timing.align_launches = 0;
    IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int b = 0; b < 4; ++b) {
        if (roff[b + 1] == roff[b]) continue;
        igd_align_job job;
        job.d_tgt = (const uint8_t *)h->d_scratch[S_TGT]; job.d_tgt_off = (const int32_t *)h->d_scratch[S_TGTOFF];
        job.d_qry = (const uint8_t *)h->d_scratch[S_QRY]; job.d_qry_off = (const int32_t *)h->d_scratch[S_QRYOFF];
        job.d_pair_t = nullptr; job.d_pair_q = nullptr; job.d_work = nullptr;
        job.d_runs = (const int4 *)h->d_scratch[S_RUNS] + roff[b];
        job.n_work = (int32_t)(roff[b + 1] - roff[b]);
        job.bucket = b; job.mixed_case = mixed; job.packed = 1; job.pay_mode = 0; job.sc = *sc;
        job.d_score = (int32_t *)h->d_scratch[S_SCORE]; job.d_pay = (uint32_t *)h->d_scratch[S_PAY];
        job.d_counter = (unsigned int *)(h->d_count + 8 + NB + b);
        if ((rc = igd_launch_gotoh(h, job))) return rc;
    }
    IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
    if (after_launch && (rc = after_launch((uint8_t *)h->h_stage + n_pairs * 4))) return rc;
    IGD_CUDA(cudaEventRecord(h->ev2, h->stream));
    if (n_pairs) IGD_CUDA(cudaMemcpyAsync(h->h_stage, h->d_scratch[S_PAY], n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f, ms_after = 0.f;
    IGD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    IGD_CUDA(cudaEventElapsedTime(&ms_after, h->ev1, h->ev2));
    acc.ms += ms + ms_after; acc.cells += cells; acc.launches += h->timing.align_launches;
    if (after_launch) acc.k5_ms += ms_after;
    if (getenv("IGD_DETECT_TRACE")) fprintf(stderr, "[igd_detect]   round: %zu pairs in %zu runs (chunk %d), %.3f ms of kernels, %.1f GCUPS; queued behind them: %.3f ms\n",
                                            n_pairs, all.size(), chunk, ms, cells / (ms * 1e6 + 1e-9), ms_after);
    const uint32_t *pay = (const uint32_t *)h->h_stage;
    for (const PrunedPair &p : pp)
        for (int k = 0; k < p.count; ++k) {
            const int32_t g = p.q0 + k;
            const uint32_t f = pay[p.slot + k], r = pay[p.slot + p.count + k];
            const bool fwd = (f & 1023u) > (r & 1023u);
            const uint32_t w = fwd ? f : r;
            const int m = (int)(w & 1023u), alen = (gene_off[g + 1] - gene_off[g]) + (int)((w >> 10) & 1023u);
            done[(size_t)p.u * n_gene + g] = 1;
            Best &b = best[(size_t)p.u];
            const long long l = (long long)(m - 1) * b.alen, rr = (long long)(b.matches - 1) * alen;
            if (b.gene < 0 || l > rr || (l == rr && g < b.gene)) b = {g, m, alen, fwd ? 1 : 0};
        }
    return IGD_OK;
}

// best[u] = what argmax_pick_kernel would give for the fragments that can yield a row, and SOME (gene, PI below every
// cut-off) for the others.  *applicable = false when the tables do not fit this path (the caller then scores densely).
static int best_by_pruning(igd_handle *h, int gt, int32_t U, const std::vector<int32_t> &toff,
                           const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                           const igd_scoring *sc, const igd_detect_params *prm, int mixed, Acc &acc,
                           std::vector<int4> &best_out, bool *applicable)
{
    *applicable = false;
    const bool enabled = !(getenv("IGD_DETECT_PRUNE") && atoi(getenv("IGD_DETECT_PRUNE")) == 0);     // read per call: tests compare both
    if (!enabled || n_gene < 64 || n_gene > 2048 || U < 1 || !igd_packed_scheme_ok(sc)) return IGD_OK;
    if (sc->target_end_open != sc->target_internal_open || sc->target_end_extend != sc->target_internal_extend ||
        sc->target_internal_extend < sc->target_internal_open || sc->query_internal_extend < sc->query_internal_open ||
        sc->query_end_extend < sc->query_end_open) return IGD_OK;
    int max_q = 0, min_q = 1 << 30;
    for (int q = 0; q < n_gene; ++q) { const int n = gene_off[q + 1] - gene_off[q]; max_q = std::max(max_q, n); min_q = std::min(min_q, n); }
    if (min_q < 1 || max_q > 511) return IGD_OK;
    for (int32_t u = 0; u < U; ++u) {
        const int nA = toff[2 * u + 1] - toff[2 * u];
        if (nA < 1 || nA > 511 || igd_align_bucket_of(nA) > 3 || igd_packed_layout_for(sc, nA, max_q) != 1 ||
            igd_packed_layout_for(sc, nA, min_q) != 1) return IGD_OK;
    }
    *applicable = true;
    Trace tr;
    int rc;
    const size_t qbytes = (size_t)gene_off[n_gene];
    if ((rc = igd_reserve(h, S_QRY, qbytes))) return rc;
    if ((rc = igd_reserve(h, S_QRYOFF, ((size_t)n_gene + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_LCS, (size_t)2 * U * n_gene * 2))) return rc;
    if ((rc = igd_reserve(h, S_MASKS, (size_t)n_gene * 5 * 16 * 4))) return rc;
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRY], gene, qbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRYOFF], gene_off, ((size_t)n_gene + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    h->timing.align_launches = 0;
    IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
    if ((rc = igd_launch_lcs(h, (const uint8_t *)h->d_scratch[S_TGT], (const int32_t *)h->d_scratch[S_TGTOFF], 2 * U, (size_t)toff[2 * U],
                             (const uint8_t *)h->d_scratch[S_QRY], (const int32_t *)h->d_scratch[S_QRYOFF], n_gene, max_q,
                             (uint32_t *)h->d_scratch[S_MASKS], (uint16_t *)h->d_scratch[S_LCS]))) return rc;
    const int K = (int)std::min<size_t>(PRUNE_TOP, (size_t)n_gene);
    if ((rc = igd_reserve(h, S_UB, (size_t)U * n_gene * 2))) return rc;
    if ((rc = igd_reserve(h, S_TOP, (size_t)U * K * 4))) return rc;
    if ((rc = igd_launch_ub_topk(h, (const uint16_t *)h->d_scratch[S_LCS], (const int32_t *)h->d_scratch[S_QRYOFF], U, n_gene, K,
                                 (uint16_t *)h->d_scratch[S_UB], (int32_t *)h->d_scratch[S_TOP]))) return rc;
    IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
    // bounds of the better orientation (LCS; the bound on the match count is LCS - 1) and, per fragment, the K genes
    // with the largest bounds: round one (the gene with the best PI is among the 16 largest bounds for every planted
    // gene of the bench workload, among the 2 largest for two thirds of them)
    // read back into pinned memory: a copy to pageable memory blocks inside the driver, which also stalls the other
    // handles of this GPU when several of them work side by side (multi.detect_pieces_pipelined)
    const size_t ub_bytes = ((size_t)U * n_gene * 2 + 15) & ~(size_t)15;
    if ((rc = igd_pin_reserve(h, 0, ub_bytes + (size_t)U * K * 4))) return rc;
    const uint16_t *ub = (const uint16_t *)h->h_pin[0];
    const int32_t *top = (const int32_t *)((const uint8_t *)h->h_pin[0] + ub_bytes);
    IGD_CUDA(cudaMemcpyAsync(h->h_pin[0], h->d_scratch[S_UB], (size_t)U * n_gene * 2, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaMemcpyAsync((uint8_t *)h->h_pin[0] + ub_bytes, h->d_scratch[S_TOP], (size_t)U * K * 4, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    IGD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    acc.ms += ms; acc.launches += h->timing.align_launches;
    acc.lcs_ms += ms;
    for (int32_t u = 0; u < U; ++u) acc.lcs_cells += 2ull * (uint64_t)(toff[2 * u + 1] - toff[2 * u]) * (uint64_t)gene_off[n_gene];
    if (getenv("IGD_DETECT_TRACE")) fprintf(stderr, "[igd_detect]   LCS kernels %.3f ms for %d x %d pairs\n", ms, 2 * U, n_gene);
    tr.mark("LCS bounds + D2H");
    // Round one: the largest bounds of every fragment (ub_topk_kernel lists them in descending order) -- the first
    // PRUNE_MIN of them, and the rest of the K while they stay within PRUNE_DELTA of the largest: the gene of maximal PI is
    // almost always among those, and whatever round one leaves out round two still aligns if its bound demands it.
    // A fragment whose largest bound is below the cut-off cannot yield a row and is aligned against nothing.
    const double cut = std::min(prm->pi_strict[gt], prm->pi_relax[gt]);
    const double delta = getenv("IGD_DETECT_TOP_DELTA") ? atof(getenv("IGD_DETECT_TOP_DELTA")) : PRUNE_DELTA;
    const int kmin = getenv("IGD_DETECT_TOP_MIN") ? atoi(getenv("IGD_DETECT_TOP_MIN")) : PRUNE_MIN;
    std::vector<std::vector<int32_t>> lists((size_t)U);
    for (int32_t u = 0; u < U; ++u) {
        std::vector<int32_t> &l = lists[(size_t)u];
        double top_ratio = 0.0;
        for (int k = 0; k < K; ++k) {
            const int32_t g = top[(size_t)u * K + k];
            if (g < 0) break;
            const double ratio = ((double)ub[(size_t)u * n_gene + g] - 1.0) / (double)(gene_off[g + 1] - gene_off[g]);
            if (k == 0) top_ratio = ratio;
            if (k >= kmin && ratio < top_ratio - delta) break;
            l.push_back(g);
        }
        std::sort(l.begin(), l.end());
    }
    std::vector<Best> best((size_t)U, Best{-1, 0, 1, 0});
    std::vector<uint8_t> done((size_t)U * n_gene, 0);
    if ((rc = prune_round(h, lists, toff, gene_off, n_gene, sc, mixed, acc, best, done))) return rc;
    tr.mark("round 1 (largest bounds)");
    // Round two.  A fragment whose best PI so far reaches the cut-off needs the genes whose LCS bound reaches that PI
    // (a few relatives).  One that stays below resembles no gene -- most candidates of a real genome -- and its LCS bounds
    // rule out nothing ((LCS - 1) / len ~ 0.69 for unrelated DNA against a cut-off of 0.60): it gets the K5 bound against
    // every remaining gene, queued behind this round's alignments, and round three aligns what K5 lets through.
    const bool ybound_ok = cut >= 50.0 && !(getenv("IGD_DETECT_YBOUND") && atoi(getenv("IGD_DETECT_YBOUND")) == 0);
    std::vector<int32_t> yslot((size_t)U, -1);
    std::vector<long long> len_values;                          // distinct gene lengths and each gene's index into them
    std::vector<uint16_t> len_class((size_t)n_gene), min_lcs;
    for (int32_t g = 0; g < n_gene; ++g) {
        const long long len = gene_off[g + 1] - gene_off[g];
        size_t c = 0;
        while (c < len_values.size() && len_values[c] != len) ++c;
        if (c == len_values.size()) len_values.push_back(len);
        len_class[(size_t)g] = (uint16_t)c;
    }
    min_lcs.resize(len_values.size());
    std::vector<int4> yruns[2];
    size_t ywords = 0, needed = 0, n_below = 0;
    std::vector<int32_t> below;
    for (int32_t u = 0; u < U; ++u) {
        std::vector<int32_t> &l = lists[(size_t)u];
        l.clear();
        const Best &b = best[(size_t)u];
        const bool by_best = b.gene >= 0 && (double)(b.matches - 1) / (double)b.alen * 100.0 >= cut;
        if (!by_best) { ++n_below; if (ybound_ok) { below.push_back(u); continue; } }
        if (by_best) {
            // (LCS - 1) x alen >= (m - 1) x len  <=>  LCS >= 1 + ceil((m - 1) x len / alen): one threshold per distinct gene
            // length (a few dozen), then one 16-bit comparison per gene
            const long long m1 = b.matches - 1, al = b.alen;
            for (size_t c = 0; c < len_values.size(); ++c) min_lcs[c] = (uint16_t)std::min<long long>(65535, 1 + (m1 * len_values[c] + al - 1) / al);
            const uint16_t *ubu = ub + (size_t)u * n_gene;
            const uint8_t *dn = done.data() + (size_t)u * n_gene;
            for (int32_t g = 0; g < n_gene; ++g)
                if (ubu[g] >= min_lcs[len_class[(size_t)g]] && !dn[g]) l.push_back(g);
        } else {
            for (int32_t g = 0; g < n_gene; ++g) {
                if (done[(size_t)u * n_gene + g]) continue;
                const int32_t num = (int32_t)ub[(size_t)u * n_gene + g] - 1;
                const long long len = gene_off[g + 1] - gene_off[g];
                if ((double)num / (double)len * 100.0 >= cut) l.push_back(g);
            }
        }
        needed += l.size();
    }
    if (!below.empty()) {
        const size_t groups = (size_t)h->sm_count * 4 * 8;
        // >= 16 runs per lane group, at least two genes per run (swept 1..10 on the bench step: 2 is the fastest, 1.04 ms)
        const int ychunk = (int)std::min<size_t>(64, std::max<size_t>(2, below.size() * (size_t)n_gene / (groups * 16)));
        for (int32_t u : below) {
            const int nA = toff[2 * u + 1] - toff[2 * u];
            yslot[(size_t)u] = (int32_t)ywords;
            for (int32_t g = 0; g < n_gene; g += ychunk)
                yruns[nA <= 352 ? 0 : 1].push_back(make_int4(u, g, std::min(ychunk, n_gene - g), (int)(ywords + g)));
            ywords += (size_t)n_gene;
            acc.k5_cells += 2ull * (uint64_t)nA * (uint64_t)gene_off[n_gene];
        }
    }
    // IGD_DETECT_YPASS_CAP: tests shrink the list to exercise the overflow path
    const unsigned pass_cap = getenv("IGD_DETECT_YPASS_CAP") ? (unsigned)std::max(1, atoi(getenv("IGD_DETECT_YPASS_CAP")))
                                                             : (unsigned)std::max<size_t>(4096, ywords / 64);
    auto launch_ybound = [&](uint8_t *stage) -> int {
        if (ywords == 0) return IGD_OK;
        int rc2;
        std::vector<int4> all = yruns[0];
        all.insert(all.end(), yruns[1].begin(), yruns[1].end());
        if ((rc2 = igd_reserve(h, S_QCLS, qbytes + 16))) return rc2;
        if ((rc2 = igd_reserve(h, S_YRUNS, all.size() * sizeof(int4)))) return rc2;
        if ((rc2 = igd_reserve(h, S_Y, ywords * 4))) return rc2;
        if ((rc2 = igd_reserve(h, S_YPASS, (size_t)pass_cap * sizeof(uint2)))) return rc2;
        if ((rc2 = igd_pin_reserve(h, 2, all.size() * sizeof(int4)))) return rc2;
        memcpy(h->h_pin[2], all.data(), all.size() * sizeof(int4));
        IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_YRUNS], h->h_pin[2], all.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
        IGD_CUDA(cudaMemsetAsync(h->d_count + 32, 0, 3 * sizeof(unsigned long long), h->stream));
        if ((rc2 = igd_launch_gene_classes(h, (const uint8_t *)h->d_scratch[S_QRY], qbytes, (uint8_t *)h->d_scratch[S_QCLS]))) return rc2;
        for (int shape = 0; shape < 2; ++shape)
            if ((rc2 = igd_launch_ybound(h, shape, (const uint8_t *)h->d_scratch[S_TGT], (const int32_t *)h->d_scratch[S_TGTOFF],
                                         (const uint8_t *)h->d_scratch[S_QCLS], (const int32_t *)h->d_scratch[S_QRYOFF],
                                         (const int4 *)h->d_scratch[S_YRUNS] + (shape ? yruns[0].size() : 0), (int)yruns[shape].size(),
                                         (uint32_t *)h->d_scratch[S_Y], (unsigned int *)(h->d_count + 32 + shape),
                                         cut, (uint2 *)h->d_scratch[S_YPASS], (unsigned int *)(h->d_count + 34), pass_cap))) return rc2;
        // the verdicts only: {count, pairs that may reach the cut-off}; the bounds themselves stay on the device
        IGD_CUDA(cudaMemcpyAsync(stage, h->d_count + 34, 8, cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaMemcpyAsync(stage + 8, h->d_scratch[S_YPASS], (size_t)pass_cap * sizeof(uint2), cudaMemcpyDeviceToHost, h->stream));
        return IGD_OK;
    };
    if (getenv("IGD_DETECT_TRACE")) fprintf(stderr, "[igd_detect]   after round 1: %zu of %d fragments below the cut-off, %zu pairs for the others\n", n_below, U, needed);
    if ((rc = prune_round(h, lists, toff, gene_off, n_gene, sc, mixed, acc, best, done, ywords ? 8 + (size_t)pass_cap * sizeof(uint2) : 0, launch_ybound))) return rc;
    tr.mark("round 2 (LCS bounds >= best PI; K5 bounds of the fragments below the cut-off)");
    if (ywords) {
        const uint8_t *stage = (const uint8_t *)h->h_stage + needed * 2 * 4;
        const unsigned n_pass = *(const unsigned *)stage;
        std::vector<uint2> pass((const uint2 *)(stage + 8), (const uint2 *)(stage + 8) + std::min(n_pass, pass_cap));
        if (n_pass > pass_cap) {                         // more than 1 in 64 (never seen): read the bounds, test them here
            std::vector<uint32_t> yb(ywords);
            IGD_CUDA(cudaMemcpyAsync(yb.data(), h->d_scratch[S_Y], ywords * 4, cudaMemcpyDeviceToHost, h->stream));
            IGD_CUDA(cudaStreamSynchronize(h->stream));
            pass.clear();
            for (int32_t u : below)
                for (int32_t g = 0; g < n_gene; ++g) {
                    const uint32_t y = yb[(size_t)yslot[(size_t)u] + g];
                    const double len = (double)(gene_off[g + 1] - gene_off[g]);
                    if (50.0 * (double)std::max(y & 0xffffu, y >> 16) >= (cut * len + 100.0) * (1.0 - 1e-12)) pass.push_back(make_uint2((unsigned)u, (unsigned)g));
                }
        }
        size_t n_ypass = 0;
        for (auto &l : lists) l.clear();
        for (const uint2 &p : pass) {
            const int32_t u = (int32_t)p.x, g = (int32_t)p.y;
            if (done[(size_t)u * n_gene + g]) continue;
            const double len = (double)(gene_off[g + 1] - gene_off[g]);
            if (((double)ub[(size_t)u * n_gene + g] - 1.0) / len * 100.0 < cut) continue;
            lists[(size_t)u].push_back(g);
            ++n_ypass;
        }
        for (int32_t u : below) std::sort(lists[(size_t)u].begin(), lists[(size_t)u].end());
        if (getenv("IGD_DETECT_TRACE")) fprintf(stderr, "[igd_detect]   K5: %zu of %zu bounded pairs may reach the cut-off\n", n_ypass, ywords);
        // round three reuses the pinned buffer the bounds sit in: they are consumed by now
        if (n_ypass && (rc = prune_round(h, lists, toff, gene_off, n_gene, sc, mixed, acc, best, done))) return rc;
        tr.mark("round 3 (K5 bounds >= cut-off)");
    }
    best_out.resize((size_t)U);
    for (int32_t u = 0; u < U; ++u) best_out[(size_t)u] = make_int4(best[(size_t)u].gene, best[(size_t)u].matches, best[(size_t)u].alen, best[(size_t)u].fwd);
    return IGD_OK;
}

struct Cand { uint32_t hit; int64_t start; int32_t len; int32_t rc; };

// One gene type (V or J): score the candidates cand[] against the gene table, append the winners' rows.
int score_gene_type(igd_handle *h, int gt, const std::vector<Cand> &cand,
                    const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                    const igd_scoring *sc, const igd_detect_params *prm, Acc &acc)
{
    if (cand.empty() || n_gene <= 0) return IGD_OK;
    Trace tr;
    int rc;
    bool any_empty = false;
    for (const Cand &c : cand) any_empty |= c.len == 0;
    Best empty_best = {-1, 0, 0, 0};
    if (any_empty) {
        if ((rc = empty_fragment_best(h, gene, gene_off, n_gene, sc, &empty_best))) return rc;
        acc.ms += h->timing.align_ms; acc.cells += h->timing.align_cells; acc.launches += h->timing.align_launches;
    }
    const int mixed = igd_case_mix(nullptr, 0, gene, (size_t)gene_off[n_gene]);      // targets are upper-cased
    constexpr int NB = 6;
    const size_t max_pairs = (size_t)1 << 28;                                       // per batch: 2 x 1 GiB of result words
    const size_t batch_frag = std::max<size_t>(1, max_pairs / (2 * (size_t)n_gene));
    for (size_t c0 = 0; c0 < cand.size();) {
        // ---- batch: candidates [c0, c1) with at most batch_frag distinct non-empty fragments ---------------
        std::unordered_map<SliceKey, int32_t, SliceHash> seen;                      // identical slices are scored once
        std::vector<FragDesc> frags;
        std::vector<int32_t> uid;
        size_t c1 = c0;
        for (; c1 < cand.size(); ++c1) {
            const Cand &c = cand[c1];
            if (c.len == 0) { uid.push_back(-1); continue; }
            const SliceKey key = {h->contig_seq_off[h->sorted_hits[c.hit].contig] + (unsigned long long)c.start, c.len, c.rc};
            auto it = seen.find(key);
            if (it == seen.end()) {
                if (frags.size() >= batch_frag) break;
                it = seen.emplace(key, (int32_t)frags.size()).first;
                frags.push_back({key.g0, c.len, c.rc});
            }
            uid.push_back(it->second);
        }
        const int32_t U = (int32_t)frags.size();
        std::vector<int4> best((size_t)std::max(U, 1));
        std::vector<int32_t> toff((size_t)2 * U + 1, 0);
        if (U > 0) {
            for (int32_t u = 0; u < U; ++u) { toff[2 * u + 1] = toff[2 * u] + frags[u].len; toff[2 * u + 2] = toff[2 * u + 1] + frags[u].len; }
            const size_t tbytes = (size_t)toff[2 * U];
            if (tbytes > 0x7fffff00u) { igd_set_error("fragment batch exceeds 2 GiB"); return IGD_ERR_CAPACITY; }
            if ((rc = igd_reserve(h, S_FRAG, (size_t)U * sizeof(FragDesc)))) return rc;
            if ((rc = igd_reserve(h, S_TGT, tbytes))) return rc;
            if ((rc = igd_reserve(h, S_TGTOFF, ((size_t)2 * U + 1) * 4))) return rc;
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_FRAG], frags.data(), (size_t)U * sizeof(FragDesc), cudaMemcpyHostToDevice, h->stream));
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGTOFF], toff.data(), ((size_t)2 * U + 1) * 4, cudaMemcpyHostToDevice, h->stream));
            gather_fragments_kernel<<<std::min(U, h->sm_count * 16), 128, 0, h->stream>>>(
                (const uint8_t *)h->d_scratch[S_SEQ], (const FragDesc *)h->d_scratch[S_FRAG], (const int32_t *)h->d_scratch[S_TGTOFF],
                h->d_comp, U, (uint8_t *)h->d_scratch[S_TGT]);
            h->timing.total_launches++;
            IGD_CUDA(cudaGetLastError());
            tr.mark("batch tables + gather launch");
            bool pruned = false;
            if ((rc = best_by_pruning(h, gt, U, toff, gene, gene_off, n_gene, sc, prm, mixed, acc, best, &pruned))) return rc;
            bool launched = false;
            if (!pruned) {                              // allocations first: nothing but launches between the timing events
                if ((rc = igd_reserve(h, S_PICK, (size_t)U * n_gene * 4))) return rc;
                if ((rc = igd_reserve(h, S_BEST, (size_t)U * sizeof(int4)))) return rc;
            }
            if (!pruned && (rc = igd_dense_launch(h, nullptr, toff.data(), 2 * U, gene, gene_off, n_gene, sc, &launched, true))) return rc;
            if (pruned) goto have_best;
            tr.mark("dense launch (host side)");
            if (!launched) {
                igd_set_error("a (fragment, gene) pair does not fit the packed alignment kernels (gene longer than 511 or scores out of range)");
                return IGD_ERR_UNSUPPORTED;
            }
            if ((rc = igd_launch_pick(h, (const uint32_t *)h->d_scratch[S_PAY], (const int32_t *)h->d_scratch[S_QRYOFF],
                                      U, n_gene, (uint32_t *)h->d_scratch[S_PICK]))) return rc;
            argmax_pick_kernel<<<(unsigned)(((size_t)U * 32 + 255) / 256), 256, 0, h->stream>>>(
                (const uint32_t *)h->d_scratch[S_PICK], U, n_gene, (int4 *)h->d_scratch[S_BEST]);
            h->timing.total_launches++; h->timing.align_launches++;
            IGD_CUDA(cudaGetLastError());
            IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
            IGD_CUDA(cudaMemcpyAsync(best.data(), h->d_scratch[S_BEST], (size_t)U * sizeof(int4), cudaMemcpyDeviceToHost, h->stream));
            IGD_CUDA(cudaStreamSynchronize(h->stream));
            float ms = 0.f;
            IGD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            acc.ms += ms; acc.cells += h->timing.align_cells; acc.launches += h->timing.align_launches;
            tr.mark("dense kernels + argmax D2H");
        }
have_best:
        // ---- threshold (:387) in float64, exactly as the reference evaluates it ------------------------------
        struct Win { size_t cand; Best b; };
        std::vector<Win> wins;
        for (size_t c = c0; c < c1; ++c) {
            const int32_t u = uid[c - c0];
            Best b;
            if (u >= 0) { const int4 v = best[(size_t)u]; b = {v.x, v.y, v.z, v.w}; } else b = empty_best;
            if (b.gene < 0) continue;
            const double pi = (double)(b.matches - 1) / (double)b.alen * 100.0;
            const double maxk = (double)b.alen;
            if (pi >= prm->pi_strict[gt] || (pi >= prm->pi_relax[gt] && maxk >= (double)prm->maxk_cutoff[gt])) wins.push_back({c, b});
        }
        // ---- the winners' alignment once more for its range (:388): two passes of the two-register kernel -----
        std::vector<int32_t> w_start(wins.size(), 0), w_ient(wins.size(), 0);
        std::vector<int32_t> pt, pq, who;
        for (size_t i = 0; i < wins.size(); ++i) {
            const int32_t u = uid[wins[i].cand - c0];
            if (u < 0) continue;                                    // empty fragment: the reference aligns '' here
            pt.push_back(2 * u + (wins[i].b.fwd ? 0 : 1)); pq.push_back(wins[i].b.gene); who.push_back((int32_t)i);
        }
        if (!pt.empty()) {
            const size_t n = pt.size();
            std::vector<int32_t> work[NB], flat;
            for (size_t i = 0; i < n; ++i) work[igd_align_bucket_of(toff[pt[i] + 1] - toff[pt[i]])].push_back((int32_t)i);
            size_t woff[NB];
            for (int b = 0; b < NB; ++b) { woff[b] = flat.size(); flat.insert(flat.end(), work[b].begin(), work[b].end()); }
            if ((rc = igd_reserve(h, S_SCORE, n * 4))) return rc;          // the dense results are consumed: reuse their slots
            if ((rc = igd_reserve(h, S_PAY, 2 * n * 4))) return rc;
            if ((rc = igd_reserve(h, S_PT, n * 4))) return rc;
            if ((rc = igd_reserve(h, S_PQ, n * 4))) return rc;
            if ((rc = igd_reserve(h, S_WORK, n * 4))) return rc;
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_PT], pt.data(), n * 4, cudaMemcpyHostToDevice, h->stream));
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_PQ], pq.data(), n * 4, cudaMemcpyHostToDevice, h->stream));
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_WORK], flat.data(), n * 4, cudaMemcpyHostToDevice, h->stream));
            std::vector<uint32_t> pay(2 * n);
            h->timing.align_launches = 0;
            IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
            for (int pass = 0; pass < 2; ++pass) {
                IGD_CUDA(cudaMemsetAsync(h->d_count + 8, 0, 3 * NB * sizeof(unsigned long long), h->stream));
                for (int b = 0; b < NB; ++b) {
                    if (work[b].empty()) continue;
                    igd_align_job job;
                    job.d_tgt = (const uint8_t *)h->d_scratch[S_TGT]; job.d_tgt_off = (const int32_t *)h->d_scratch[S_TGTOFF];
                    job.d_qry = (const uint8_t *)h->d_scratch[S_QRY]; job.d_qry_off = (const int32_t *)h->d_scratch[S_QRYOFF];
                    job.d_pair_t = (const int32_t *)h->d_scratch[S_PT]; job.d_pair_q = (const int32_t *)h->d_scratch[S_PQ];
                    job.d_work = (const int32_t *)h->d_scratch[S_WORK] + woff[b]; job.d_runs = nullptr;
                    job.n_work = (int32_t)work[b].size();
                    job.bucket = b; job.mixed_case = mixed; job.packed = 0; job.pay_mode = pass; job.sc = *sc;
                    job.d_score = (int32_t *)h->d_scratch[S_SCORE];
                    job.d_pay = (uint32_t *)h->d_scratch[S_PAY] + (size_t)pass * n;
                    job.d_counter = (unsigned int *)(h->d_count + 8 + b);
                    if ((rc = igd_launch_gotoh(h, job))) return rc;
                }
            }
            IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
            IGD_CUDA(cudaMemcpyAsync(pay.data(), h->d_scratch[S_PAY], 2 * n * 4, cudaMemcpyDeviceToHost, h->stream));
            IGD_CUDA(cudaStreamSynchronize(h->stream));
            float ms = 0.f;
            IGD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            acc.ms += ms; acc.launches += h->timing.align_launches;
            tr.mark("winners' detail pass");
            for (size_t i = 0; i < n; ++i) {
                const int nA = toff[pt[i] + 1] - toff[pt[i]], nB = gene_off[pq[i] + 1] - gene_off[pq[i]];
                const Best &b = wins[(size_t)who[i]].b;
                const int m = (int)(pay[i] & 1023u), alen = nB + (int)((pay[i] >> 10) & 1023u);
                if (m != b.matches || alen != b.alen) { igd_set_error("internal: packed and two-register kernels disagree"); return IGD_ERR_INTERNAL; }
                acc.cells += 2ull * (uint64_t)nA * (uint64_t)nB;
                w_start[(size_t)who[i]] = (int)((pay[i] >> 20) & 1023u);
                w_ient[(size_t)who[i]] = nA - (int)((pay[n + i] >> 20) & 1023u);
            }
        }
        for (size_t i = 0; i < wins.size(); ++i) {
            const Cand &c = cand[wins[i].cand];
            const igd_hit &hit = h->sorted_hits[c.hit];
            igd_vj_row r;
            r.hept = hit.hept; r.nona = hit.nona; r.frag_start = c.start; r.contig = hit.contig; r.frag_len = c.len;
            r.best_gene = wins[i].b.gene; r.matches = wins[i].b.matches; r.alen = wins[i].b.alen;
            r.start = w_start[i]; r.ient = w_ient[i];
            r.gene_type = (uint8_t)gt; r.strand = hit.strand;
            r.direction = (uint8_t)((c.len > 0 && wins[i].b.fwd) ? '+' : '-');
            r.reserved = 0;
            h->vj_rows.push_back(r);
        }
        c0 = c1;
    }
    return IGD_OK;
}

} // namespace

extern "C" {

int igd_detect(igd_handle *h, const int32_t spacer[4],
               const uint8_t *vgene, const int32_t *vgene_off, int32_t n_vgene,
               const uint8_t *jgene, const int32_t *jgene_off, int32_t n_jgene,
               const igd_scoring *sc, const igd_detect_params *prm,
               uint64_t *n_hits, uint64_t *n_vj_rows, uint64_t *n_d_rows)
{
    if (!h || !spacer || !sc || !prm || !n_hits || !n_vj_rows || !n_d_rows || n_vgene < 0 || n_jgene < 0 ||
        (n_vgene && (!vgene || !vgene_off)) || (n_jgene && (!jgene || !jgene_off))) {
        igd_set_error("bad arguments to igd_detect"); return IGD_ERR_ARG;
    }
    if (prm->gene_length[0] < 1 || prm->gene_length[0] > IGD_MAX_TARGET_LEN || prm->gene_length[1] < 1 || prm->gene_length[1] > IGD_MAX_TARGET_LEN) {
        igd_set_error("gene_length must lie in 1..%d", IGD_MAX_TARGET_LEN); return IGD_ERR_ARG;
    }
    h->vj_rows.clear(); h->d_rows.clear(); h->row_text.clear(); h->row_text_off.clear();
    int rc;
    Trace tr;
    if ((rc = igd_run_rss_scan(h, spacer, n_hits))) return rc;
    tr.mark("scan");
    uint64_t n = 0;
    if ((rc = igd_sorted_hits(h, &n))) return rc;
    tr.mark("hits D2H + decode + sort");
    if (!h->d_comp) {
        uint8_t comp[256];
        for (int i = 0; i < 256; ++i) comp[i] = (uint8_t)i;
        const char *from = "ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", *to = "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa";
        for (int i = 0; from[i]; ++i) comp[(uint8_t)from[i]] = (uint8_t)to[i];
        IGD_CUDA(cudaMalloc(&h->d_comp, 256));
        IGD_CUDA(cudaMemcpyAsync(h->d_comp, comp, 256, cudaMemcpyHostToDevice, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
    }
    const std::vector<igd_hit> &hits = h->sorted_hits;
    // hits are sorted by (contig, type, strand, list index): one block per contig, eight sub-lists inside
    std::vector<Cand> cand[2][2];                       // [V, J][strand]
    std::vector<igd_d_row> drows[2];
    std::vector<std::pair<int64_t, uint32_t>> dl;           // (heptamer, list position)
    std::vector<std::pair<uint32_t, uint32_t>> win;         // per D_right hit: its window [x, y) of dl
    std::vector<uint32_t> sel;
    for (size_t b0 = 0; b0 < hits.size();) {
        size_t b1 = b0;
        while (b1 < hits.size() && hits[b1].contig == hits[b0].contig) ++b1;
        const int64_t L = (int64_t)h->contig_len[hits[b0].contig];
        size_t lo[4][2], hi[4][2];
        for (int t = 0; t < 4; ++t) for (int s = 0; s < 2; ++s) lo[t][s] = hi[t][s] = b0;
        for (size_t i = b0; i < b1;) {
            size_t j = i;
            while (j < b1 && hits[j].sig_type == hits[i].sig_type && hits[j].strand == hits[i].strand) ++j;
            lo[hits[i].sig_type][hits[i].strand] = i; hi[hits[i].sig_type][hits[i].strand] = j;
            i = j;
        }
        // V / J candidates: get_s_fragment_from_RSS (:268-286)
        for (int gt = 0; gt < 2; ++gt) {
            const int t = gt == 0 ? IGD_SIG_V : IGD_SIG_J;
            const int64_t glen = prm->gene_length[gt];
            for (int s = 0; s < 2; ++s)
                for (size_t i = lo[t][s]; i < hi[t][s]; ++i) {
                    const int64_t q = hits[i].hept;
                    int64_t st, ln;
                    const bool before = (gt == 0 && s == 0) || (gt == 1 && s == 1);     // the letters in front of the heptamer
                    if (before) py_slice(q - glen, q, L, st, ln); else py_slice(q + 7, q + 7 + glen, L, st, ln);
                    cand[gt][s].push_back({(uint32_t)i, st, (int32_t)ln, s});
                }
        }
        // D pairs: combine_D_RSS (:210-224), D_right outer / D_left inner
        for (int s = 0; s < 2; ++s) {
            const size_t l0 = lo[IGD_SIG_D_LEFT][s], l1 = hi[IGD_SIG_D_LEFT][s], r0 = lo[IGD_SIG_D_RIGHT][s], r1 = hi[IGD_SIG_D_RIGHT][s];
            if (l0 == l1 || r0 == r1) continue;
            // D_left by heptamer, ascending: the list is ordered by its nonamers in scanned-strand coordinates, so it is
            // (nearly) ascending on '+' and descending on '-' already
            dl.clear();
            if (s == 0) for (size_t i = l0; i < l1; ++i) dl.push_back({hits[i].hept, (uint32_t)(i - l0)});
            else        for (size_t i = l1; i-- > l0;)   dl.push_back({hits[i].hept, (uint32_t)(i - l0)});
            if (!std::is_sorted(dl.begin(), dl.end())) std::sort(dl.begin(), dl.end());
            const int64_t span = prm->d_gene_length + 7;
            // '+': left = dl, right = dr: dr-span <= dl <= dr-1;  '-': left = dr, right = dl: dr+1 <= dl <= dr+span.
            // The D_right list ascends by heptamer on '+' and descends on '-': walked by ascending heptamer, both window
            // ends only move forward (two pointers instead of two binary searches per D_right hit).
            win.assign(r1 - r0, std::make_pair((uint32_t)0, (uint32_t)0));
            {
                size_t x = 0, y = 0;
                bool ascending = true;
                int64_t prev = INT64_MIN;
                for (size_t n = 0; n < r1 - r0 && ascending; ++n) {
                    const size_t k = s == 0 ? r0 + n : r1 - 1 - n;
                    ascending = hits[k].hept >= prev; prev = hits[k].hept;
                }
                for (size_t n = 0; n < r1 - r0; ++n) {
                    const size_t k = s == 0 ? r0 + n : r1 - 1 - n;
                    const int64_t dr = hits[k].hept;
                    const int64_t a = s == 0 ? dr - span : dr + 1, b = s == 0 ? dr - 1 : dr + span;
                    if (ascending) {
                        while (x < dl.size() && dl[x].first < a) ++x;
                        if (y < x) y = x;
                        while (y < dl.size() && dl[y].first <= b) ++y;
                    } else {
                        x = (size_t)(std::lower_bound(dl.begin(), dl.end(), std::make_pair(a, (uint32_t)0)) - dl.begin());
                        y = (size_t)(std::upper_bound(dl.begin(), dl.end(), std::make_pair(b, (uint32_t)0xffffffffu)) - dl.begin());
                    }
                    win[k - r0] = {(uint32_t)x, (uint32_t)y};
                }
            }
            for (size_t k = r0; k < r1; ++k) {
                const size_t x = win[k - r0].first, y = win[k - r0].second;
                if (x == y) continue;
                sel.clear();
                for (size_t it = x; it < y; ++it) sel.push_back(dl[it].second);
                std::sort(sel.begin(), sel.end());
                for (uint32_t p : sel) {
                    const igd_hit &l = hits[l0 + p];
                    igd_d_row r;
                    r.dl_hept = l.hept; r.dl_nona = l.nona; r.dr_hept = hits[k].hept; r.dr_nona = hits[k].nona;
                    r.contig = hits[k].contig; r.strand = (uint8_t)s; r.reserved[0] = r.reserved[1] = r.reserved[2] = 0;
                    drows[s].push_back(r);
                }
            }
        }
        b0 = b1;
    }
    h->d_rows = drows[0];
    h->d_rows.insert(h->d_rows.end(), drows[1].begin(), drows[1].end());
    tr.mark("candidates + D pairing");
    const float scan_ms = h->timing.scan_ms;
    Acc acc;
    for (int gt = 0; gt < 2; ++gt) {
        std::vector<Cand> all = cand[gt][0];
        all.insert(all.end(), cand[gt][1].begin(), cand[gt][1].end());
        if ((rc = score_gene_type(h, gt, all, gt == 0 ? vgene : jgene, gt == 0 ? vgene_off : jgene_off, gt == 0 ? n_vgene : n_jgene,
                                  sc, prm, acc))) return rc;
    }
    // ---- the rows' letters, read from the resident contigs: heptamer, nonamer and aligned gene sequence of a V / J row,
    //      the four k-mers and the sequence between the heptamers of a D row; Python slice semantics, upper-cased ----
    {
        std::vector<TextDesc> segs;
        segs.reserve(3 * h->vj_rows.size() + 5 * h->d_rows.size());
        unsigned long long at = 0;
        auto push = [&](uint32_t contig, int64_t a, int64_t b, bool rcomp) {
            int64_t st, n;
            py_slice(a, b, (int64_t)h->contig_len[contig], st, n);
            segs.push_back({h->contig_seq_off[contig] + (unsigned long long)st, at, (int32_t)n, rcomp ? 1 : 0});
            at += (unsigned long long)n;
        };
        for (const igd_vj_row &r : h->vj_rows) {
            const bool rev = r.strand == IGD_STRAND_REV;
            push(r.contig, r.hept, r.hept + 7, rev);
            push(r.contig, r.nona, r.nona + 9, rev);
            if (r.frag_len > 0) {
                // the aligned orientation: reverse-complemented once for a '-' strand RSS and once more when the reverse
                // complement aligned better (:309-317); [start, ient) of THAT string
                const bool flip = rev != (r.direction == '-');
                const int64_t lo = std::min<int64_t>(r.start, r.frag_len), hi = std::max<int64_t>(lo, std::min<int64_t>(r.ient, r.frag_len));
                const int64_t a = flip ? r.frag_start + r.frag_len - hi : r.frag_start + lo;
                segs.push_back({h->contig_seq_off[r.contig] + (unsigned long long)a, at, (int32_t)(hi - lo), flip ? 1 : 0});
                at += (unsigned long long)(hi - lo);
            } else {
                segs.push_back({0ull, at, 0, 0});
            }
        }
        for (const igd_d_row &r : h->d_rows) {
            const bool rev = r.strand == IGD_STRAND_REV;
            push(r.contig, r.dl_hept, r.dl_hept + 7, rev);
            push(r.contig, r.dl_nona, r.dl_nona + 9, rev);
            push(r.contig, r.dr_hept, r.dr_hept + 7, rev);
            push(r.contig, r.dr_nona, r.dr_nona + 9, rev);
            if (!rev) push(r.contig, r.dl_hept + 7, r.dr_hept, false); else push(r.contig, r.dr_hept + 7, r.dl_hept, true);
        }
        h->row_text_off.resize(segs.size() + 1);
        for (size_t k = 0; k < segs.size(); ++k) h->row_text_off[k] = segs[k].off;
        h->row_text_off[segs.size()] = at;
        h->row_text.resize((size_t)at);
        if (!segs.empty() && at > 0) {
            if ((rc = igd_reserve(h, S_TEXT, (size_t)at))) return rc;
            if ((rc = igd_reserve(h, S_TEXTDESC, segs.size() * sizeof(TextDesc)))) return rc;
            if ((rc = igd_pin_reserve(h, 1, segs.size() * sizeof(TextDesc)))) return rc;
            if ((rc = igd_stage_reserve(h, (size_t)at))) return rc;
            memcpy(h->h_pin[1], segs.data(), segs.size() * sizeof(TextDesc));
            IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TEXTDESC], h->h_pin[1], segs.size() * sizeof(TextDesc), cudaMemcpyHostToDevice, h->stream));
            gather_text_kernel<<<(unsigned)std::min<size_t>(segs.size(), (size_t)h->sm_count * 16), 128, 0, h->stream>>>(
                (const uint8_t *)h->d_scratch[S_SEQ], (const TextDesc *)h->d_scratch[S_TEXTDESC], h->d_comp, (int)segs.size(),
                (uint8_t *)h->d_scratch[S_TEXT]);
            IGD_CUDA(cudaGetLastError());
            h->timing.total_launches++;
            IGD_CUDA(cudaMemcpyAsync(h->h_stage, h->d_scratch[S_TEXT], (size_t)at, cudaMemcpyDeviceToHost, h->stream));
            IGD_CUDA(cudaStreamSynchronize(h->stream));
            memcpy(h->row_text.data(), h->h_stage, (size_t)at);
        }
        tr.mark("row letters");
    }
    h->timing.scan_ms = scan_ms;
    h->timing.align_ms = acc.ms; h->timing.align_cells = acc.cells; h->timing.align_launches = acc.launches;
    h->timing.lcs_ms = acc.lcs_ms; h->timing.k5_ms = acc.k5_ms; h->timing.dp_ms = acc.ms - acc.lcs_ms - acc.k5_ms;
    h->timing.lcs_cells = acc.lcs_cells; h->timing.k5_cells = acc.k5_cells;
    *n_vj_rows = h->vj_rows.size();
    *n_d_rows = h->d_rows.size();
    return IGD_OK;
}

int igd_row_text_size(igd_handle *h, uint64_t *n_bytes, uint64_t *n_segments)
{
    if (!h || !n_bytes || !n_segments) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    *n_bytes = h->row_text.size();
    *n_segments = h->row_text_off.empty() ? 0 : h->row_text_off.size() - 1;
    return IGD_OK;
}

int igd_fetch_row_text(igd_handle *h, uint8_t *text, uint64_t text_capacity, uint64_t *offsets, uint64_t offsets_capacity)
{
    if (!h || (!text && text_capacity) || (!offsets && offsets_capacity)) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (h->row_text.size() > text_capacity || h->row_text_off.size() > offsets_capacity) {
        igd_set_error("capacity too small for %llu bytes / %llu offsets", (unsigned long long)h->row_text.size(), (unsigned long long)h->row_text_off.size());
        return IGD_ERR_CAPACITY;
    }
    if (!h->row_text.empty()) memcpy(text, h->row_text.data(), h->row_text.size());
    if (!h->row_text_off.empty()) memcpy(offsets, h->row_text_off.data(), h->row_text_off.size() * sizeof(uint64_t));
    return IGD_OK;
}

int igd_fetch_vj_rows(igd_handle *h, igd_vj_row *out, uint64_t capacity)
{
    if (!h || (!out && capacity)) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (h->vj_rows.size() > capacity) { igd_set_error("capacity %llu < %llu rows", (unsigned long long)capacity, (unsigned long long)h->vj_rows.size()); return IGD_ERR_CAPACITY; }
    if (!h->vj_rows.empty()) memcpy(out, h->vj_rows.data(), h->vj_rows.size() * sizeof(igd_vj_row));
    return IGD_OK;
}

int igd_fetch_d_rows(igd_handle *h, igd_d_row *out, uint64_t capacity)
{
    if (!h || (!out && capacity)) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (h->d_rows.size() > capacity) { igd_set_error("capacity %llu < %llu rows", (unsigned long long)capacity, (unsigned long long)h->d_rows.size()); return IGD_ERR_CAPACITY; }
    if (!h->d_rows.empty()) memcpy(out, h->d_rows.data(), h->d_rows.size() * sizeof(igd_d_row));
    return IGD_OK;
}

} // extern "C"
