// K3: batched integer Gotoh global aligner that carries the reference's alignment statistics
// forward through the DP instead of tracing back.
//
// Replaces ALIGNER.align(target, query)[0] + BioAlign (py/IGDetective.py:310-311,
// py/extract_aligned_genes.py:9-56,92-95) with the scheme of SetupAligner (:72-83).
// BioPython's aligner keeps, per cell and state, the set of optimal predecessors and returns as
// alignment [0] the path that prefers M over Ix over Iy at the end cell and at every step back.
// Here every state carries the statistics of exactly that path:
//   key  = 4 * score + priority of the state the value came from (M 2, Ix 1, Iy 0), so a plain
//          integer max reproduces the reference's tie-break;
//   pay  = matches | internal-Ix-moves << 10 | (lead or trailing-Ix-moves) << 20, selected by the
//          same comparison.  matches counts M moves whose upper-cased letters are equal (the DP
//          itself compares raw bytes, as the reference does on raw-case windows).
// From the end cell: start = lead, end = start + nB + internal Ix, ient = nA - trailing Ix.
//
// Mapping: one pair per group of G lanes; lane l owns C consecutive target rows in registers and
// sweeps the gene columns as a systolic wavefront (lane l works on column t-l at step t); the
// bottom row of each lane travels to the next lane by shuffle.  No shared memory, no tensor
// cores: the work is integer add / max / select.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int GT_THREADS = 128;
constexpr int NEGKEY = -(1 << 26);

struct GotohParams {
    const uint8_t *tgt; const int32_t *tgt_off;
    const uint8_t *qry; const int32_t *qry_off;
    const int32_t *pair_t; const int32_t *pair_q;
    const int32_t *work; int32_t n_work;
    int match4, mismatch4;       // 4 * substitution scores
    int to4, te4;                // 4 * target gap open / extend (internal == end)
    int qio4, qie4;              // 4 * query internal gap open / extend
    int qeo4, qee4;              // 4 * query end gap open / extend
    unsigned u_int, u_trail;     // payload units of an Ix move in 0<j<nB / j==nB
    int lead_shift;              // 20: store lead in bits 20..29, -1: do not
    int32_t *out_score; uint32_t *out_pay;
    unsigned int *counter;
};

__device__ __forceinline__ unsigned up8(unsigned c) { return (c - 'a' < 26u) ? c - 32u : c; }

template <int G, int C, bool MIXED>
__global__ void __launch_bounds__(GT_THREADS)
gotoh_kernel(const GotohParams P)
{
    constexpr int GROUPS_PER_WARP = 32 / G;
    const unsigned lane = threadIdx.x & 31u;
    const int l = lane % G;                  // lane within the group
    const int grp = lane / G;

    for (;;) {
        // ---- fetch one pair per group (warp-uniform control flow) -----------------------------
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(P.counter, (unsigned)GROUPS_PER_WARP);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned)P.n_work) break;
        const bool have = base + grp < (unsigned)P.n_work;
        int pair = 0, nA = 0, nB = 0;
        const uint8_t *A = P.tgt, *B = P.qry;
        if (have) {
            pair = P.work[base + grp];
            const int t = P.pair_t[pair], q = P.pair_q[pair];
            const int a0 = P.tgt_off[t], b0 = P.qry_off[q];
            nA = P.tgt_off[t + 1] - a0; nB = P.qry_off[q + 1] - b0;
            A += a0; B += b0;
        }
        int steps = have ? nB + G - 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));

        // ---- per-lane rows ------------------------------------------------------------------
        unsigned a[C];
        int Dk[C], Yk[C];
        unsigned Dp[C], Yp[C];
        const int i0 = l * C;                                   // row above this lane's first row
#pragma unroll
        for (int r = 0; r < C; ++r) {
            const int i = i0 + r + 1;
            unsigned ch = (have && i <= nA) ? A[i - 1] : 0u;
            a[r] = MIXED ? (ch | (up8(ch) << 8)) : ch;
            Dk[r] = (P.qeo4 + P.qee4 * (i - 1)) | 1;            // D(i,0) = Ix(i,0): free/left query-end gap
            Dp[r] = P.lead_shift >= 0 ? ((unsigned)i << P.lead_shift) : 0u;
            Yk[r] = NEGKEY; Yp[r] = 0u;
        }
        // D(i0, j-1) as seen by this lane's first row: column-0 boundary before the first step
        int diagk = (l == 0) ? 2 : ((P.qeo4 + P.qee4 * (i0 - 1)) | 1);
        unsigned diagp = (l == 0 || P.lead_shift < 0) ? 0u : ((unsigned)i0 << P.lead_shift);
        int botDk = NEGKEY, botXk = NEGKEY; unsigned botDp = 0, botXp = 0;

        for (int t = 0; t < steps; ++t) {
            // bottom row of the lane above, computed one step ago for the same column
            int upDk = __shfl_up_sync(0xffffffffu, botDk, 1, G);
            unsigned upDp = __shfl_up_sync(0xffffffffu, botDp, 1, G);
            int upXk = __shfl_up_sync(0xffffffffu, botXk, 1, G);
            unsigned upXp = __shfl_up_sync(0xffffffffu, botXp, 1, G);
            const int j = t - l + 1;
            if (j >= 1 && j <= nB) {
                if (l == 0) {                                   // row 0: Iy(0,j), target end gap
                    upDk = P.to4 + P.te4 * (j - 1); upDp = 0u; upXk = NEGKEY; upXp = 0u;
                }
                const unsigned braw = B[j - 1];
                const unsigned b = MIXED ? (braw | (up8(braw) << 8)) : braw;
                const bool lastcol = (j == nB);
                const int qo4 = lastcol ? P.qeo4 : P.qio4;
                const int qe4 = lastcol ? P.qee4 + 1 : P.qie4 + 1;   // +1: priority of the Ix state on the extend candidate
                const unsigned ux = lastcol ? P.u_trail : P.u_int;
                const int nextdiagk = upDk; const unsigned nextdiagp = upDp;
                int dk = diagk; unsigned dp = diagp;              // D(i-1, j-1)
                int uk = upDk; unsigned upay = upDp;              // D(i-1, j)
                int xk = upXk; unsigned xp = upXp;                // Ix(i-1, j), key without priority bits
#pragma unroll
                for (int r = 0; r < C; ++r) {
                    // M(i,j)
                    bool eq_s, eq_m;
                    if (MIXED) {
                        const unsigned x = a[r] ^ b;
                        eq_s = (x & 0xffu) == 0u; eq_m = (x & 0xff00u) == 0u;
                    } else {
                        eq_s = eq_m = (a[r] == b);
                    }
                    const int mk = ((dk & ~3) | 2) + (eq_s ? P.match4 : P.mismatch4);
                    const unsigned mp = dp + (eq_m ? 1u : 0u);
                    // Ix(i,j) from (i-1,j): open from D (priority of D's origin) vs extend Ix (priority 1)
                    const int c1 = uk + qo4, c2 = xk + qe4;
                    const bool px = c1 >= c2;
                    const int nxk = max(c1, c2) & ~3;             // stored without priority
                    const unsigned nxp = (px ? upay : xp) + ux;
                    // Iy(i,j) from (i,j-1): open from D vs extend Iy (priority 0)
                    const int e1 = Dk[r] + P.to4, e2 = Yk[r] + P.te4;
                    const bool py = e1 >= e2;
                    const int nyk = max(e1, e2) & ~3;
                    const unsigned nyp = py ? Dp[r] : Yp[r];
                    // D(i,j) = best of M (2), Ix (1), Iy (0)
                    const int xk1 = nxk | 1;
                    const bool p1 = mk >= xk1;
                    int bk = max(mk, xk1); unsigned bp = p1 ? mp : nxp;
                    const bool p2 = bk >= nyk;
                    bp = p2 ? bp : nyp; bk = max(bk, nyk);
                    // shift the window
                    dk = Dk[r]; dp = Dp[r];
                    Dk[r] = bk; Dp[r] = bp; Yk[r] = nyk; Yp[r] = nyp;
                    uk = bk; upay = bp; xk = nxk; xp = nxp;
                }
                botDk = uk; botDp = upay; botXk = xk; botXp = xp;
                diagk = nextdiagk; diagp = nextdiagp;
                if (l == 0) {                                   // D(0, j) for the next column's diagonal
                    diagk = P.to4 + P.te4 * (j - 1); diagp = 0u;
                }
            }
        }
        // ---- result: D(nA, nB) sits in the lane / row that owns target row nA ----------------
        if (have && nA >= 1) {
            const int ls = (nA - 1) / C, rs = (nA - 1) % C;
            if (l == ls) {
                int k = 0; unsigned p = 0;
#pragma unroll
                for (int r = 0; r < C; ++r) if (r == rs) { k = Dk[r]; p = Dp[r]; }
                P.out_score[pair] = k >> 2;
                P.out_pay[pair] = p;
            }
        }
    }
}

// ---- packed variant: one 32-bit word per state ------------------------------------------------
// word = score << 20 | priority << 18 | internal-Ix-moves << 9 | matches.  Candidates that tie on
// (score, priority) always descend from the same state and so carry the same payload, hence a
// single integer max selects score, tie-break and payload at once: 12 integer instructions per
// cell instead of 27.  Only (score, matches, alignment length) come out; lead / ient need the
// two-register kernel above.  The host routes a pair here only if every field provably fits
// (target <= 511 rows, query <= 511 columns, score bounds inside +-1900, see capi.cu).
// Two field layouts (LONG = false / true):
//   short: score 12 bits | priority 2 | internal Ix 9 | matches 9   target <= 511, query <= 511, |score| <= 1900
//   long : score 11 bits | priority 2 | internal Ix 10 | matches 9  target <= 1023, query <= 511, |score| <= 900
template <bool LONG> struct PK {
    static constexpr int S = LONG ? (1 << 21) : (1 << 20);
    static constexpr int PR = S >> 2, PRMASK = 3 * PR, UI = 1 << 9;
    static constexpr int NEG = (LONG ? -900 : -2000) * S;
    static constexpr int SHIFT = LONG ? 21 : 20;
    static constexpr unsigned IXMASK = LONG ? 1023u : 511u;
};

struct PackedParams {
    const uint8_t *tgt; const int32_t *tgt_off;
    const uint8_t *qry; const int32_t *qry_off;
    const int4 *runs; int32_t n_runs;   // {target, first query, number of consecutive queries, first output slot}
    int inc_match, inc_mismatch;  // score * S + 2 * PR (+1 match for inc_match)
    int to, te;                   // target gap open / extend, * S
    int qio, qie, qeo, qee;       // query internal / end gap open / extend, * S
    int m_clear, m_pr;            // ~PK_PRMASK and PK_PR as run-time values: (x & m_clear) | m_pr is then ONE LOP3
    int one;                      // run-time 1: x * one + c is an IMAD on the FMA pipe, off the saturated ALU pipe
    int32_t *out_score; uint32_t *out_pay;
    unsigned int *counter;
};

// A work item is a RUN: one target against nq consecutive queries of the table.  The lanes sweep
// the concatenated columns of the whole run, so the systolic pipeline fills and drains once per
// run instead of once per pair; at a gene boundary a lane stores the result (if it owns row nA)
// and re-initialises its column state while the lanes below are still finishing the previous gene.
//
// FPINC: the substitution increment is computed on the FMA pipe instead of the saturated ALU pipe.
// Letters are held as floats; eq = sat(1 - d*d) is exactly 1.0 or 0.0, and the increment words are used as float BIT PATTERNS: a non-negative
// integer below 2^24 is the bit pattern of the float n * 2^-149, whose sums are exact, so
// fma(eq, match - mismatch, mismatch) yields the integer increment bit for bit (FADD + FFMA.SAT + FFMA
// replace ISETP + SEL).  The host selects FPINC only when both increments lie in [0, 2^24).
// With lower-case letters the DP compares raw bytes and the match count upper-cased ones: equal letters
// of the same case are a match, equal letters of different case only count (mismatch + 1), so
// inc = mismatch + [up(a) == up(b)] * (same case ? match - mismatch : 1).  The letters are upper-cased
// once, the factor is a per-row constant (MIXED = 1: queries are all upper case) that a lower-case query
// letter swaps by one more FFMA per cell (MIXED = 2).
// resident CTAs per SM the register allocation aims for: 6 for the shapes that fit 80 registers
#ifndef IGD_V_FRAME
#define IGD_V_FRAME 0
#endif
#ifndef IGD_V_MB
#define IGD_V_MB 4
#endif
constexpr int packed_min_blocks(int C) { return C <= 11 ? 6 : (C == 22 ? IGD_V_MB : 1); }
// which instances keep their columns in the shifted frame (see the kernel)
__host__ __device__ constexpr bool packed_frame(int C, bool LONG) { return !LONG && (C != 22 || IGD_V_FRAME); }
// MIXED: 0 = no lower-case letter anywhere; 1 = only targets may hold lower-case letters; 2 = queries may too.
template <int G, int C, int MIXED, bool LONG, bool FPINC>
__global__ void __launch_bounds__(GT_THREADS, packed_min_blocks(C))
gotoh_packed_kernel(const PackedParams P)
{
    constexpr int PK_PR = PK<LONG>::PR, PK_PRMASK = PK<LONG>::PRMASK, PK_UI = PK<LONG>::UI, PK_NEG = PK<LONG>::NEG;
    constexpr int GROUPS_PER_WARP = 32 / G;
    const unsigned lane = threadIdx.x & 31u;
    const int l = lane % G;
    const int grp = lane / G;
    const int m_clear = P.m_clear, m_pr = P.m_pr, one = P.one;
    // Column frame: every value stored for column j is the true value minus j * te (the Iy extension score), so that
    // extending a gap along the row -- Iy(i,j) = max(D(i,j-1) + to, Iy(i,j-1) + te) -- needs no add on the extension
    // side: max(D + (to - te), Iy) in the frame, ONE VIADDMNMX.  The diagonal step crosses one column (the increment
    // words carry -te), the vertical step stays in its column (unchanged), the result leaves the frame at the end.
    // Short layout only: the long layout's 11-bit score field has no room for the frame's + j.
    constexpr bool FRAME = packed_frame(C, LONG);
    const int to_te = P.to - P.te;

    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(P.counter, (unsigned)GROUPS_PER_WARP);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned)P.n_runs) break;
        const bool have = base + grp < (unsigned)P.n_runs;
        int nA = 0, nB = 1, nq = 0, out = 0, q = 0;
        const uint8_t *A = P.tgt, *B = P.qry;
        int steps = 0;
        if (have) {
            const int4 run = P.runs[base + grp];
            const int a0 = P.tgt_off[run.x];
            nA = P.tgt_off[run.x + 1] - a0;
            A += a0;
            q = run.y; nq = run.z; out = run.w;
            const int b0 = P.qry_off[q];
            nB = P.qry_off[q + 1] - b0;
            B += b0;
            steps = P.qry_off[q + nq] - b0 + G - 1;            // all columns of the run + pipeline depth
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));

        unsigned a[C];
        float af[C], wdr[C];
        const float wx = __int_as_float(P.inc_mismatch), wd = __int_as_float(P.inc_match) - __int_as_float(P.inc_mismatch);
        const float w1 = __int_as_float(1);
        int D[C], Y[C], E[C];                                 // D(i, j-1) with priority, Iy(i, j-1) clean, D(i-1, j-1) clean
        const int i0 = l * C;
#pragma unroll
        for (int r = 0; r < C; ++r) {
            const int i = i0 + r + 1;
            const unsigned ch = (have && i <= nA) ? A[i - 1] : 0u;
            a[r] = MIXED ? (ch | (up8(ch) << 8)) : ch;
            af[r] = (float)(MIXED ? up8(ch) : ch);
            wdr[r] = (ch - 'a' < 26u) ? w1 : wd;
            D[r] = P.qeo + P.qee * (i - 1) + PK_PR;           // D(i,0) = Ix(i,0), priority 1
            Y[r] = PK_NEG;
            E[r] = (i == 1) ? 0 : P.qeo + P.qee * (i - 2);    // clean D(i-1, 0)
        }
        int botD = PK_NEG, botX = PK_NEG;
        const int ls = (nA - 1) / C, rs = (nA - 1) % C;       // lane / row that owns target row nA
        int j = 0;                                            // columns of the current gene already done
        bool active = have && nA >= 1;

        for (int t = 0; t < steps; ++t) {
            int upD = __shfl_up_sync(0xffffffffu, botD, 1, G);
            int upX = __shfl_up_sync(0xffffffffu, botX, 1, G);
            if (active && t >= l) {
                ++j;
                if (l == 0) { upD = FRAME ? to_te : P.to + P.te * (j - 1); upX = PK_NEG; }   // row 0: Iy(0,j) = to + te (j-1), priority 0
                const unsigned braw = B[j - 1];
                const unsigned b = MIXED ? (braw | (up8(braw) << 8)) : braw;
                const float bf = (float)(MIXED ? up8(braw) : braw);
                const bool blow = MIXED == 2 && (braw - 'a' < 26u);
                const float bsg = blow ? -1.0f : 1.0f, bof = blow ? wd + w1 : 0.0f;   // swaps wd <-> 1 in wdr[r]
                const bool lastcol = (j == nB);
                // an Ix move in an internal column counts one internal-Ix unit: the unit rides on the
                // gap score constants, so both candidates carry it into the max
                const int qo = lastcol ? P.qeo : P.qio + PK_UI;
                const int qe = lastcol ? P.qee : P.qie + PK_UI;
                int uD = upD, uX = upX;
#pragma unroll
                for (int r = 0; r < C; ++r) {
                    int inc;
                    if (FPINC) {
                        const float dd = af[r] - bf;
                        const float eq = __saturatef(fmaf(-dd, dd, 1.0f));
                        if (MIXED == 2) {
                            inc = __float_as_int(fmaf(eq, fmaf(wdr[r], bsg, bof), wx));
                        } else if (MIXED == 1) {
                            inc = __float_as_int(fmaf(eq, wdr[r], wx));
                        } else {
                            inc = __float_as_int(fmaf(eq, wd, wx));
                        }
                    } else if (MIXED) {
                        const unsigned x = a[r] ^ b;
                        inc = (x & 0xffu) == 0u ? P.inc_match : ((x & 0xff00u) == 0u ? P.inc_mismatch + 1 : P.inc_mismatch);
                    } else {
                        inc = (a[r] == b) ? P.inc_match : P.inc_mismatch;
                    }
                    // M(i,j), priority 2.  Short layout: the add is an IMAD (x * one + c), which moves it from the
                    // saturated ALU pipe to the FMA pipe (6 + 6 instead of 7 + 5 per cell: 1.83 -> 1.90 TCUPS); the long
                    // layout at 255 registers is slower that way (1.75 -> 1.70) and keeps the plain add
                    const int m = LONG ? E[r] + inc : E[r] * one + inc;
                    const int x = __viaddmax_s32(uD, qo, uX * one + qe);               // Ix(i,j) raw
                    const int xs = (x & m_clear) | m_pr;                               // priority 1 (one LOP3)
                    // Iy(i,j), priority 0: the extension is free in the column frame
                    const int y = (FRAME ? __viaddmax_s32(D[r], to_te, Y[r]) : __viaddmax_s32(D[r], P.to, Y[r] * one + P.te)) & ~PK_PRMASK;
                    const int d = __vimax3_s32(m, xs, y);
                    // the diagonal of the next column is updated in place once it has been read, so no
                    // state needs two live copies (saves the register moves of "dc = old D[r]")
                    E[r] = (r == 0 && l == 0) ? (FRAME ? to_te : P.to + P.te * (j - 1)) : (uD & ~PK_PRMASK);
                    D[r] = d; Y[r] = y;
                    uD = d; uX = xs;
                }
                botD = uD; botX = uX;
                if (lastcol) {                                // this lane is through with the current gene
                    if (l == ls) {
                        int k = 0;
#pragma unroll
                        for (int r = 0; r < C; ++r) if (r == rs) k = D[r];
                        P.out_score[out] = (k >> PK<LONG>::SHIFT) + (FRAME ? nB * (P.te >> PK<LONG>::SHIFT) : 0);   // out of the column frame
                        P.out_pay[out] = ((unsigned)k & 511u) | ((((unsigned)k >> 9) & PK<LONG>::IXMASK) << 10);   // wide-kernel layout
                    }
                    ++out; ++q;
                    if (--nq > 0) {
                        B += nB;
                        nB = P.qry_off[q + 1] - P.qry_off[q];
                        j = 0;
#pragma unroll
                        for (int r = 0; r < C; ++r) {
                            D[r] = P.qeo + P.qee * (i0 + r) + PK_PR; Y[r] = PK_NEG;
                            E[r] = (i0 + r == 0) ? 0 : P.qeo + P.qee * (i0 + r - 1);
                        }
                    } else {
                        active = false;
                    }
                }
            }
        }
    }
}

template <int G, int C, bool LONG>
int launch_packed(igd_handle *h, const PackedParams &P, int mixed)
{
    const unsigned warps_needed = (unsigned)((P.n_runs + (32 / G) - 1) / (32 / G));
    unsigned blocks = (warps_needed + (GT_THREADS / 32) - 1) / (GT_THREADS / 32);
    static const unsigned per_sm = getenv("IGD_GOTOH_BLOCKS_PER_SM") ? (unsigned)atoi(getenv("IGD_GOTOH_BLOCKS_PER_SM")) : 6u;
    const unsigned cap = (unsigned)h->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // increments usable as float bit patterns (see FPINC above); IGD_GOTOH_FPINC=0 forces the integer form
    static const bool fp_allowed = !(getenv("IGD_GOTOH_FPINC") && atoi(getenv("IGD_GOTOH_FPINC")) == 0);
    const bool fp = fp_allowed && h->fp_increment_ok && P.inc_match >= 0 && P.inc_match < (1 << 24) && P.inc_mismatch >= 0 && P.inc_mismatch + 1 < (1 << 24);
    if (fp) {
        if (mixed == 2)      gotoh_packed_kernel<G, C, 2, LONG, true><<<blocks, GT_THREADS, 0, h->stream>>>(P);
        else if (mixed == 1) gotoh_packed_kernel<G, C, 1, LONG, true><<<blocks, GT_THREADS, 0, h->stream>>>(P);
        else                 gotoh_packed_kernel<G, C, 0, LONG, true><<<blocks, GT_THREADS, 0, h->stream>>>(P);
    } else {                 // the integer form compares raw and upper-cased bytes directly: one MIXED instance
        if (mixed) gotoh_packed_kernel<G, C, 2, LONG, false><<<blocks, GT_THREADS, 0, h->stream>>>(P);
        else       gotoh_packed_kernel<G, C, 0, LONG, false><<<blocks, GT_THREADS, 0, h->stream>>>(P);
    }
    h->timing.total_launches++;
    h->timing.align_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

template <int G, int C>
int launch(igd_handle *h, const GotohParams &P, int mixed)
{
    const unsigned warps_needed = (unsigned)((P.n_work + (32 / G) - 1) / (32 / G));
    unsigned blocks = (warps_needed + (GT_THREADS / 32) - 1) / (GT_THREADS / 32);
    const unsigned cap = (unsigned)h->sm_count * 4u;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (mixed) gotoh_kernel<G, C, true><<<blocks, GT_THREADS, 0, h->stream>>>(P);
    else       gotoh_kernel<G, C, false><<<blocks, GT_THREADS, 0, h->stream>>>(P);
    h->timing.total_launches++;
    h->timing.align_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

} // namespace

// kernel shapes: target rows = G * C
//   0: 8 x 9   (<= 72,  J flanks of 70)      1: 32 x 6  (<= 192)
//   2: 16 x 22 (<= 352, V flanks of 350; packed kernel), 32 x 11 (two-register kernel)     3: 32 x 16 (<= 512)
//   4: 32 x 25 (<= 800, scoring-B windows)   5: 32 x 32 (<= 1024)
int igd_align_bucket_of(int target_len)
{
    if (target_len <= 72) return 0;
    if (target_len <= 192) return 1;
    if (target_len <= 352) return 2;
    if (target_len <= 512) return 3;
    if (target_len <= 800) return 4;
    if (target_len <= 1024) return 5;
    return -1;
}

template <bool LONG>
static void fill_packed(PackedParams &P, const igd_align_job &job, int C)
{
    constexpr int S = PK<LONG>::S, PR = PK<LONG>::PR;
    P.tgt = job.d_tgt; P.tgt_off = job.d_tgt_off; P.qry = job.d_qry; P.qry_off = job.d_qry_off;
    P.runs = job.d_runs; P.n_runs = job.n_work;
    // column frame of the short-layout kernel: the diagonal step crosses one column, so both increments carry -te
    const int frame = packed_frame(C, LONG) ? job.sc.target_internal_extend : 0;
    P.inc_match = (job.sc.match - frame) * S + 2 * PR + 1;
    P.inc_mismatch = (job.sc.mismatch - frame) * S + 2 * PR;
    P.to = job.sc.target_internal_open * S; P.te = job.sc.target_internal_extend * S;
    P.qio = job.sc.query_internal_open * S; P.qie = job.sc.query_internal_extend * S;
    P.qeo = job.sc.query_end_open * S; P.qee = job.sc.query_end_extend * S;
    P.m_clear = ~PK<LONG>::PRMASK; P.m_pr = PR; P.one = 1;
    P.out_score = job.d_score; P.out_pay = job.d_pay; P.counter = job.d_counter;
}

static int launch_gotoh_packed(igd_handle *h, const igd_align_job &job)
{
    PackedParams P;
    static const bool narrow = getenv("IGD_GOTOH_V_SHAPE") && !strcmp(getenv("IGD_GOTOH_V_SHAPE"), "32x11");
    static const int rows_of[6] = {9, 6, 22, 16, 25, 32};
    if (job.packed == 1) {
        fill_packed<false>(P, job, job.bucket == 2 && narrow ? 11 : rows_of[job.bucket]);
        switch (job.bucket) {
        case 0: return launch_packed<8, 9, false>(h, P, job.mixed_case);
        case 1: return launch_packed<32, 6, false>(h, P, job.mixed_case);
        case 2: {
            // 16 lanes x 22 rows: the ~29 per-column instructions (query letter, shuffles, boundary selects) are shared by
            // 22 cells instead of 11 -- 13.3 instead of 14.6 instructions per cell, 2.04 instead of 1.90 TCUPS at 128
            // registers / 4 CTAs per SM (8 x 44 at 255 registers / 2 CTAs: 1.96).  IGD_GOTOH_V_SHAPE=32x11 selects the old shape.
            return narrow ? launch_packed<32, 11, false>(h, P, job.mixed_case) : launch_packed<16, 22, false>(h, P, job.mixed_case);
        }
        case 3: return launch_packed<32, 16, false>(h, P, job.mixed_case);
        }
    } else {
        fill_packed<true>(P, job, rows_of[job.bucket]);
        switch (job.bucket) {
        case 4: return launch_packed<32, 25, true>(h, P, job.mixed_case);
        case 5: return launch_packed<32, 32, true>(h, P, job.mixed_case);
        }
    }
    igd_set_error("bad packed align bucket %d (layout %d)", job.bucket, job.packed);
    return IGD_ERR_ARG;
}

// The FPINC instances rely on exact denormal arithmetic (no flush-to-zero): checked once per handle with the
// very operations the kernel uses; a failure only switches the library to the integer form.
__global__ void probe_denormal_kernel(int a, int b, int *out)
{
    const float wx = __int_as_float(b), wd = __int_as_float(a) - __int_as_float(b), w1 = __int_as_float(1);
    const float eq = __saturatef(fmaf(-0.0f, 0.0f, 1.0f)), ne = __saturatef(fmaf(-3.0f, 3.0f, 1.0f));
    out[0] = __float_as_int(fmaf(eq, wd, wx));                       // a
    out[1] = __float_as_int(fmaf(ne, wd, wx));                       // b
    out[2] = __float_as_int(fmaf(eq, fmaf(wd, -1.0f, wd + w1), wx)); // b + 1
}

int igd_probe_fp_increment(igd_handle *h, bool *ok)
{
    const int a = (2 << 21) + (1 << 20) + 1, b = 1 << 20;
    int *d = (int *)h->d_count + 96, host[3] = {0, 0, 0};          // scratch words past the counters in use
    probe_denormal_kernel<<<1, 1, 0, h->stream>>>(a, b, d);
    IGD_CUDA(cudaGetLastError());
    IGD_CUDA(cudaMemcpyAsync(host, d, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    *ok = host[0] == a && host[1] == b && host[2] == b + 1;
    return IGD_OK;
}

// ComputeAlignment's choice between the two orientations of a fragment (py/IGDetective.py:314-317):
// forward iff its match count is strictly larger.  pay rows 2u / 2u+1 = fragment u / its reverse complement.
__global__ void __launch_bounds__(256)
pick_orientation_kernel(const uint32_t *pay, const int32_t *qry_off, int n_unique, int n_qry, uint32_t *out)
{
    const size_t n = (size_t)n_unique * n_qry;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t u = i / n_qry, q = i - u * n_qry;
        const uint32_t f = pay[(2 * u) * n_qry + q], r = pay[(2 * u + 1) * n_qry + q];
        const bool fwd = (f & 1023u) > (r & 1023u);
        const uint32_t w = fwd ? f : r;
        const uint32_t alen = (uint32_t)(qry_off[q + 1] - qry_off[q]) + ((w >> 10) & 1023u);
        out[i] = (w & 1023u) | (alen << 10) | (fwd ? 0x80000000u : 0u);
    }
}

// extract_aligned_genes.ComputeAlignment's scan (py/extract_aligned_genes.py:89-100) over
// [window, RC(window)] x genes in that order, keeping the LAST maximum of PI = (matches-1)/len*100 among
// PI >= 0 (best_pi starts at 0, the test is >=).  Fractions are compared exactly by cross-multiplication
// (equal rationals give equal float64 PI, different ones differ by far more than an ulp).
// One warp per window; pay rows 2w / 2w+1 = window w / its reverse complement.  out[w] = {index o*G+g or -1, matches, len, 0}.
__global__ void __launch_bounds__(256)
select_window_best_kernel(const uint32_t *pay, const int32_t *tgt_off, const int32_t *qry_off, int n_win, int n_qry, int4 *out)
{
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_win) return;
    long long bn = 0, bd = 1; int bi = -1, bm = 0;
    for (int k = lane; k < 2 * n_qry; k += 32) {
        const int o = k >= n_qry ? 1 : 0, g = k - o * n_qry, t = 2 * w + o;
        if (tgt_off[t + 1] == tgt_off[t]) continue;                 // empty target: no matches, PI < 0
        const uint32_t p = pay[(size_t)t * n_qry + g];
        const int m = (int)(p & 1023u);
        if (m < 1) continue;
        const long long num = m - 1, den = (qry_off[g + 1] - qry_off[g]) + (int)((p >> 10) & 1023u);
        if (bi < 0 || num * bd >= bn * den) { bn = num; bd = den; bi = k; bm = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long on = __shfl_xor_sync(0xffffffffu, bn, o), od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o), om = __shfl_xor_sync(0xffffffffu, bm, o);
        if (oi >= 0) {
            const long long l = on * bd, r = bn * od;
            if (bi < 0 || l > r || (l == r && oi > bi)) { bn = on; bd = od; bi = oi; bm = om; }
        }
    }
    if (lane == 0) out[w] = make_int4(bi, bm, (int)bd, 0);
}

int igd_launch_select_windows(igd_handle *h, const uint32_t *d_pay, const int32_t *d_tgt_off, const int32_t *d_qry_off,
                              int32_t n_win, int32_t n_qry, int4 *d_out)
{
    const unsigned blocks = (unsigned)(((size_t)n_win * 32 + 255) / 256);
    select_window_best_kernel<<<blocks ? blocks : 1, 256, 0, h->stream>>>(d_pay, d_tgt_off, d_qry_off, n_win, n_qry, d_out);
    h->timing.total_launches++;
    h->timing.align_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

int igd_launch_pick(igd_handle *h, const uint32_t *d_pay, const int32_t *d_qry_off, int32_t n_unique, int32_t n_qry,
                    uint32_t *d_out)
{
    const size_t n = (size_t)n_unique * n_qry;
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8);
    if (blocks < 1) blocks = 1;
    pick_orientation_kernel<<<blocks, 256, 0, h->stream>>>(d_pay, d_qry_off, n_unique, n_qry, d_out);
    h->timing.total_launches++;
    h->timing.align_launches++;
    IGD_CUDA(cudaGetLastError());
    return IGD_OK;
}

int igd_launch_gotoh(igd_handle *h, const igd_align_job &job)
{
    if (job.packed) return launch_gotoh_packed(h, job);
    GotohParams P;
    P.tgt = job.d_tgt; P.tgt_off = job.d_tgt_off; P.qry = job.d_qry; P.qry_off = job.d_qry_off;
    P.pair_t = job.d_pair_t; P.pair_q = job.d_pair_q; P.work = job.d_work; P.n_work = job.n_work;
    P.match4 = 4 * job.sc.match; P.mismatch4 = 4 * job.sc.mismatch;
    P.to4 = 4 * job.sc.target_internal_open; P.te4 = 4 * job.sc.target_internal_extend;
    P.qio4 = 4 * job.sc.query_internal_open; P.qie4 = 4 * job.sc.query_internal_extend;
    P.qeo4 = 4 * job.sc.query_end_open; P.qee4 = 4 * job.sc.query_end_extend;
    P.u_int = 1u << 10;
    P.u_trail = job.pay_mode == 1 ? (1u << 20) : 0u;
    P.lead_shift = job.pay_mode == 0 ? 20 : -1;
    P.out_score = job.d_score; P.out_pay = job.d_pay; P.counter = job.d_counter;
    switch (job.bucket) {
    case 0: return launch<8, 9>(h, P, job.mixed_case);
    case 1: return launch<32, 6>(h, P, job.mixed_case);
    case 2: return launch<32, 11>(h, P, job.mixed_case);
    case 3: return launch<32, 16>(h, P, job.mixed_case);
    case 4: return launch<32, 25>(h, P, job.mixed_case);
    case 5: return launch<32, 32>(h, P, job.mixed_case);
    }
    igd_set_error("bad align bucket %d", job.bucket);
    return IGD_ERR_ARG;
}
