/*
 * oracle/gotoh_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the pairwise aligner the reference calls for every
 * candidate-vs-gene comparison:
 *     ALIGNER.align(query, seq_B)[0]             (/root/reference/py/IGDetective.py:310-311)
 *     aligner.align(query, gene.seq) ; [0]       (/root/reference/py/extract_aligned_genes.py:92-95)
 * with the scoring scheme of SetupAligner        (/root/reference/py/extract_aligned_genes.py:72-83).
 *
 * The arithmetic lives in a third-party dependency that is NOT vendored in the
 * reference and NOT installed here: BioPython (>= 1.82, unpinned, README.md:8),
 * Bio.Align.PairwiseAligner, whose C core (Bio/Align/_pairwisealigner.c,
 * "_aligners.c" before 1.84) runs Gotoh's three-state global alignment with
 * double scores, stores for every cell and state the SET of optimal
 * predecessors, and yields optimal paths in a fixed order.  This file restates
 * that published algorithm:
 *   - states M (diagonal), Ix (vertical: consumes a target letter, '-' in the
 *     query row), Iy (horizontal: consumes a query letter, '-' in the target row);
 *   - candidate order M, Ix, Iy; a later candidate replaces the running best only
 *     if it is larger by more than epsilon (1e-6) and is ADDED to the trace set if
 *     within epsilon;
 *   - left/right end-gap scores apply in column 0 / column nB for Ix (query gaps)
 *     and in row 0 / row nA for Iy (target gaps);
 *   - the first alignment ([0]) starts from the first of M, Ix, Iy that attains
 *     the optimum in the corner cell and at every step moves to the first
 *     flagged predecessor in the order M, Ix, Iy.
 * PARITY UNPINNED: the reference ships no tests or golden vectors, and real
 * BioPython could not be run in this environment, so the tie-break order above
 * is restated from the published source, not checked against its output.
 * Supporting evidence only: tests/test_oracle_biopython_recalled.py lists seven
 * known-answer cases of BioPython's own Tests/test_pairwise_aligner.py, recalled
 * from memory (scores and first alignments); this file reproduces all of them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product path (igdetective_b200/) never does.
 */
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <ctype.h>

#define T_M  1
#define T_IX 2
#define T_IY 4

typedef struct {
    double match, mismatch;
    double t_int_open, t_int_ext;     /* target gaps ('-' in target row), internal rows 0<i<nA */
    double t_left_open, t_left_ext;   /* row 0 */
    double t_right_open, t_right_ext; /* row nA */
    double q_int_open, q_int_ext;     /* query gaps ('-' in query row), internal columns 0<j<nB */
    double q_left_open, q_left_ext;   /* column 0 */
    double q_right_open, q_right_ext; /* column nB */
} igo_scoring;

typedef struct {
    double score;
    int ncols;        /* alignment columns */
    int start, end;   /* BioAlign gene_range: [first non-gap query column, last non-gap query column + 1) */
    int matches;      /* TRUE count of columns in range whose upper-cased letters are equal (reference reports matches-1) */
    int lead;         /* number of target letters before the range (QuerySeq = A[lead:ient]) */
    int ient;         /* target letters consumed up to and including the range */
} igo_stats;

#define EPS 1e-6

/* running max with trace-set semantics of the BioPython selection macros */
#define SEL3(s1, s2, s3, outscore, outtrace) do {                  \
        double _sc = (s1), _t; unsigned char _tr = T_M;             \
        _t = (s2);                                                   \
        if (_t > _sc + EPS) { _sc = _t; _tr = T_IX; }                \
        else if (_t > _sc - EPS) _tr |= T_IX;                        \
        _t = (s3);                                                   \
        if (_t > _sc + EPS) { _sc = _t; _tr = T_IY; }                \
        else if (_t > _sc - EPS) _tr |= T_IY;                        \
        (outscore) = _sc; (outtrace) = _tr;                          \
    } while (0)

/*
 * Fills the three trace matrices and returns the optimal score.  Matrices are
 * (nA+1) x (nB+1), row-major.  Score rows are rolled.
 */
static double fill(const unsigned char *A, int nA, const unsigned char *B, int nB,
                   const igo_scoring *sc, unsigned char *tM, unsigned char *tX, unsigned char *tY,
                   unsigned char *end_states)
{
    const int W = nB + 1;
    double *M = (double *)malloc(sizeof(double) * W * 3);
    double *X = M + W, *Y = X + W;
    int i, j;

    M[0] = 0.0; X[0] = -DBL_MAX; Y[0] = -DBL_MAX;
    tM[0] = 0; tX[0] = 0; tY[0] = 0;
    for (j = 1; j <= nB; j++) {
        M[j] = -DBL_MAX; X[j] = -DBL_MAX;
        Y[j] = sc->t_left_open + sc->t_left_ext * (j - 1);
        tM[j] = 0; tX[j] = 0; tY[j] = (j == 1) ? T_M : T_IY;
    }
    for (i = 1; i <= nA; i++) {
        const unsigned char a = A[i - 1];
        unsigned char *rM = tM + (size_t)i * W, *rX = tX + (size_t)i * W, *rY = tY + (size_t)i * W;
        /* values of row i-1 at column j-1 (diagonal), saved before overwrite */
        double dM = M[0], dX = X[0], dY = Y[0];
        const double to = (i == nA) ? sc->t_right_open : sc->t_int_open;
        const double te = (i == nA) ? sc->t_right_ext : sc->t_int_ext;
        M[0] = -DBL_MAX;
        X[0] = sc->q_left_open + sc->q_left_ext * (i - 1);
        Y[0] = -DBL_MAX;
        rM[0] = 0; rX[0] = (i == 1) ? T_M : T_IX; rY[0] = 0;
        for (j = 1; j <= nB; j++) {
            const double qo = (j == nB) ? sc->q_right_open : sc->q_int_open;
            const double qe = (j == nB) ? sc->q_right_ext : sc->q_int_ext;
            const double uM = M[j], uX = X[j], uY = Y[j];   /* row i-1, column j */
            double s; unsigned char t;
            /* M: from any state at (i-1, j-1) */
            SEL3(dM, dX, dY, s, t);
            if (s > -DBL_MAX / 2) s += (a == B[j - 1]) ? sc->match : sc->mismatch;
            rM[j] = t;
            dM = uM; dX = uX; dY = uY;
            M[j] = s;
            /* Ix: from (i-1, j) */
            SEL3(uM > -DBL_MAX / 2 ? uM + qo : -DBL_MAX,
                 uX > -DBL_MAX / 2 ? uX + qe : -DBL_MAX,
                 uY > -DBL_MAX / 2 ? uY + qo : -DBL_MAX, s, t);
            rX[j] = t; X[j] = s;
            /* Iy: from (i, j-1), which this row already holds */
            SEL3(M[j - 1] > -DBL_MAX / 2 ? M[j - 1] + to : -DBL_MAX,
                 X[j - 1] > -DBL_MAX / 2 ? X[j - 1] + to : -DBL_MAX,
                 Y[j - 1] > -DBL_MAX / 2 ? Y[j - 1] + te : -DBL_MAX, s, t);
            rY[j] = t; Y[j] = s;
        }
    }
    {
        double s; unsigned char t;
        SEL3(M[nB], X[nB], Y[nB], s, t);
        *end_states = t;
        free(M);
        return s;
    }
}

/*
 * First optimal alignment as gapped rows.  row_t / row_q need nA + nB + 1 bytes.
 * Returns 0, or -1 on allocation failure.
 */
int igo_gotoh_first(const unsigned char *A, int nA, const unsigned char *B, int nB,
                    const igo_scoring *sc, double *score, char *row_t, char *row_q, int *ncols)
{
    const size_t cells = (size_t)(nA + 1) * (size_t)(nB + 1);
    unsigned char *tM = (unsigned char *)malloc(cells * 3);
    unsigned char *tX, *tY, ends;
    int i = nA, j = nB, n = 0, k, state;
    const int W = nB + 1;
    if (!tM) return -1;
    tX = tM + cells; tY = tX + cells;
    *score = fill(A, nA, B, nB, sc, tM, tX, tY, &ends);
    state = (ends & T_M) ? T_M : (ends & T_IX) ? T_IX : T_IY;
    if (nA == 0 && nB == 0) state = 0;
    /* walk back; rows are produced reversed and flipped afterwards */
    while (state && (i > 0 || j > 0)) {
        unsigned char tr;
        if (state == T_M)       { tr = tM[(size_t)i * W + j]; row_t[n] = (char)A[i - 1]; row_q[n] = (char)B[j - 1]; i--; j--; }
        else if (state == T_IX) { tr = tX[(size_t)i * W + j]; row_t[n] = (char)A[i - 1]; row_q[n] = '-';             i--; }
        else                    { tr = tY[(size_t)i * W + j]; row_t[n] = '-';             row_q[n] = (char)B[j - 1]; j--; }
        n++;
        state = (tr & T_M) ? T_M : (tr & T_IX) ? T_IX : (tr & T_IY) ? T_IY : 0;
    }
    for (k = 0; k < n / 2; k++) {
        char c = row_t[k]; row_t[k] = row_t[n - 1 - k]; row_t[n - 1 - k] = c;
        c = row_q[k]; row_q[k] = row_q[n - 1 - k]; row_q[n - 1 - k] = c;
    }
    row_t[n] = 0; row_q[n] = 0;
    *ncols = n;
    free(tM);
    return 0;
}

/*
 * The numbers the reference's BioAlign class derives from the first alignment
 * (/root/reference/py/extract_aligned_genes.py:9-56), computed from the
 * traceback instead of from Python loops over the gapped strings:
 *   start,end = _ComputeGeneRange()  (:19-31)
 *   matches   = TRUE number of equal upper-cased columns inside the range
 *               (NumMatches() at :33-38 starts counting at -1 and so returns matches-1)
 *   lead,ient : QuerySeq() (:43-50) == upper(A[lead:ient])
 */
int igo_gotoh_stats(const unsigned char *A, int nA, const unsigned char *B, int nB,
                    const igo_scoring *sc, igo_stats *out)
{
    char *rt = (char *)malloc((size_t)(nA + nB + 1) * 2);
    char *rq;
    int n, k, start, end, m = 0, lead = 0, cons = 0, ient = 0;
    if (!rt) return -1;
    rq = rt + (nA + nB + 1);
    if (igo_gotoh_first(A, nA, B, nB, sc, &out->score, rt, rq, &n)) { free(rt); return -1; }
    start = 0;
    for (k = 0; k < n; k++) if (rq[k] != '-') { start = k; break; }
    end = n;
    for (k = n - 1; k >= 0; k--) if (rq[k] != '-') { end = k + 1; break; }
    for (k = 0; k < n; k++) {
        if (k == start) lead = cons;
        if (k >= start && k < end && toupper((unsigned char)rt[k]) == toupper((unsigned char)rq[k])) m++;
        if (rt[k] != '-') cons++;
        if (k == end - 1) ient = cons;
    }
    if (n == 0) { lead = 0; ient = 0; }
    out->ncols = n; out->start = start; out->end = end; out->matches = m;
    out->lead = lead; out->ient = ient;
    free(rt);
    return 0;
}

/*
 * Batch form for the CPU baseline: pairs (pair_t[k], pair_q[k]) index into
 * concatenated byte arrays with offset tables; n_threads POSIX threads pull
 * pairs from a shared counter.
 */
#include <pthread.h>

typedef struct {
    const unsigned char *tgt; const int *tgt_off;
    const unsigned char *qry; const int *qry_off;
    const int *pair_t, *pair_q; long n_pairs;
    const igo_scoring *sc; igo_stats *out;
    long next; int err; pthread_mutex_t mu;
} batch_ctx;

static void *batch_worker(void *p)
{
    batch_ctx *c = (batch_ctx *)p;
    for (;;) {
        long k0, k1, k;
        pthread_mutex_lock(&c->mu);
        k0 = c->next; c->next += 8;
        pthread_mutex_unlock(&c->mu);
        if (k0 >= c->n_pairs) break;
        k1 = k0 + 8 < c->n_pairs ? k0 + 8 : c->n_pairs;
        for (k = k0; k < k1; k++) {
            const int t = c->pair_t[k], q = c->pair_q[k];
            if (igo_gotoh_stats(c->tgt + c->tgt_off[t], c->tgt_off[t + 1] - c->tgt_off[t],
                                c->qry + c->qry_off[q], c->qry_off[q + 1] - c->qry_off[q],
                                c->sc, &c->out[k]))
                c->err = -1;
        }
    }
    return 0;
}

int igo_gotoh_stats_batch(const unsigned char *tgt, const int *tgt_off,
                          const unsigned char *qry, const int *qry_off,
                          const int *pair_t, const int *pair_q, long n_pairs,
                          const igo_scoring *sc, igo_stats *out, int n_threads)
{
    batch_ctx c;
    pthread_t th[256];
    int i;
    c.tgt = tgt; c.tgt_off = tgt_off; c.qry = qry; c.qry_off = qry_off;
    c.pair_t = pair_t; c.pair_q = pair_q; c.n_pairs = n_pairs; c.sc = sc; c.out = out;
    c.next = 0; c.err = 0;
    pthread_mutex_init(&c.mu, 0);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    for (i = 1; i < n_threads; i++) pthread_create(&th[i], 0, batch_worker, &c);
    batch_worker(&c);
    for (i = 1; i < n_threads; i++) pthread_join(th[i], 0);
    pthread_mutex_destroy(&c.mu);
    return c.err;
}
