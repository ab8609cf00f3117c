/*
 * igd_b200.h -- C ABI of the B200-native RSS scan + V/D/J scoring library
 * (libigd_b200.so, hand-written sm_100a CUDA behind plain C entry points).
 *
 * The reference (Immunotools/IgDetective) has no FFI: its hot path is Python
 * functions over BioPython.  Each entry point below names the reference
 * function it replaces (paths relative to the reference checkout); the ctypes
 * binding a maintainer would add is shown in INTEGRATION.md and implemented in
 * igdetective_b200/_lib.py.
 *
 * Conventions: all pointers are HOST pointers to caller-owned, contiguous
 * buffers; every function returns 0 (IGD_OK) or a negative IGD_ERR_* code and
 * leaves a message for igd_last_error(); no exceptions cross the ABI; a handle
 * owns one device, one CUDA stream and its device buffers, and is
 * thread-compatible (one thread at a time per handle).  There is NO CPU
 * fallback: without a usable CUDA device igd_create fails.
 */
#ifndef IGD_B200_H
#define IGD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGD_OK               0
#define IGD_ERR_CUDA        -1   /* CUDA runtime / launch failure */
#define IGD_ERR_ARG         -2   /* invalid argument */
#define IGD_ERR_STATE       -3   /* call out of order (e.g. scan before load) */
#define IGD_ERR_TOO_LONG    -4   /* sequence longer than the kernel limits below */
#define IGD_ERR_UNSUPPORTED -5   /* scoring scheme outside what the kernel implements */
#define IGD_ERR_CAPACITY    -6   /* output buffer too small */
#define IGD_ERR_INTERNAL    -7   /* two kernels of the library disagree: a bug, never a property of the input */

#define IGD_MAX_TARGET_LEN  1023 /* fragment / window rows held in registers (32 lanes x 32); 10-bit payload fields */
#define IGD_MAX_QUERY_LEN   1023 /* gene columns; bounded by the 10-bit payload fields */

/* signal types, in the bit order of the motif flag tables */
#define IGD_SIG_V       0
#define IGD_SIG_D_LEFT  1
#define IGD_SIG_D_RIGHT 2
#define IGD_SIG_J       3
#define IGD_STRAND_FWD  0
#define IGD_STRAND_REV  1

typedef struct igd_handle igd_handle;

/* One RSS = (heptamer, nonamer) pair, forward-strand 0-based contig coordinates:
 * the tuples find_valid_rss returns (py/IGDetective.py:147-177). */
typedef struct {
    int64_t  hept;      /* start of the heptamer (of its reverse complement for strand REV) */
    int64_t  nona;      /* start of the nonamer  (ditto) */
    uint32_t contig;    /* index into the contigs given to igd_load_contigs */
    uint8_t  sig_type;  /* IGD_SIG_* */
    uint8_t  strand;    /* IGD_STRAND_* */
    uint8_t  spacer;    /* spacer length that matched (s, s-1 or s+1) */
    uint8_t  reserved;
} igd_hit;

/* One k-mer membership hit: an element of the set find_valid_motif_idx returns
 * (py/IGDetective.py:138-144), for all signal types and both strands at once. */
typedef struct {
    int64_t  pos;       /* forward-strand 0-based start of the k-mer window */
    uint32_t contig;
    uint8_t  flags;     /* bit t: window in motifs[t][k] (strand +); bit 4+t: RC(window) in motifs[t][k] (strand -) */
    uint8_t  reserved[3];
} igd_kmer_hit;

/* SetupAligner (py/extract_aligned_genes.py:72-83): the ten PairwiseAligner
 * attributes the reference sets, as integers.  "target" is the first argument
 * of align() (fragment / window), "query" the second (gene).  The kernel
 * requires target_end_* == target_internal_* (true for the reference scheme);
 * anything else returns IGD_ERR_UNSUPPORTED. */
typedef struct {
    int32_t match, mismatch;
    int32_t target_internal_open, target_internal_extend;
    int32_t target_end_open, target_end_extend;
    int32_t query_internal_open, query_internal_extend;
    int32_t query_end_open, query_end_extend;
} igd_scoring;

/* device-side timings of the most recent calls, CUDA events on the handle's stream */
typedef struct {
    float    pack_ms;          /* igd_load_contigs: pack kernel */
    float    scan_ms;          /* igd_scan / igd_scan_kmers: scan kernel only */
    float    align_ms;         /* igd_align_batch: all DP kernel launches of the call */
    uint32_t scan_launches;    /* kernels launched by the last scan call */
    uint32_t align_launches;   /* kernels launched by the last align call */
    uint64_t scan_bases;       /* contig bases covered by the last scan */
    uint64_t align_cells;      /* sum len(target)*len(query) over pairs (x passes) of the last align call */
    uint64_t total_launches;   /* kernels launched by this handle since igd_create */
    /* igd_detect only: how align_ms splits.  dp_ms: the Gotoh kernels (align_cells are theirs); lcs_ms: the LCS bound,
     * its top-K selection (K4); k5_ms: the insertion-charging bound (K5) over k5_cells letter pairs (both orientations) */
    float    dp_ms, lcs_ms, k5_ms, reserved_f;
    uint64_t lcs_cells, k5_cells;
} igd_timing;

const char *igd_last_error(void);
int         igd_device_count(void);
igd_handle *igd_create(int device);            /* NULL on failure (see igd_last_error) */
void        igd_destroy(igd_handle *h);
int         igd_sync(igd_handle *h);
int         igd_get_timing(igd_handle *h, igd_timing *out);

/* ---- RSS scan ------------------------------------------------------------------------------- */

/* VALID_MOTIFS (py/IGDetective.py:54-58, datafiles/motifs) as flag tables:
 * flags7[16384], flags9[262144]; index = base-4 value of the k-mer, first
 * letter most significant, A=0 C=1 G=2 T=3; bit t set iff the k-mer is in
 * motifs[t][k] (t = IGD_SIG_*).  Reverse-strand tables are derived inside. */
int igd_set_motifs(igd_handle *h, const uint8_t *flags7, const uint8_t *flags9);

/* Host-only helper of the scan (no GPU needed), exposed for tests: the 262144-bit "centre" bitmap
 * igd_set_motifs derives from flags9.  Bit c is set iff the 9-mer c, or a 9-mer overlapping it
 * shifted by one base to either side, is in some nonamer set on either strand; the scan looks up
 * the 9-mer in the middle of the three candidate nonamer starts (spacer s-1, s, s+1 of
 * find_valid_rss, py/IGDetective.py:162-168) to discard a heptamer's window with one probe.
 * Bit order: index = (high bits of the 2-bit codes << 9) | low bits, first base = bit 0. */
int igd_center_bitmap(const uint8_t *flags9, uint32_t *bitmap /* 8192 words */);

/* parent_seq of get_contigwise_rss (py/IGDetective.py:180): contigs as raw ASCII
 * (any case, IUPAC allowed), concatenated; offsets[n_contigs+1].  Copies to the
 * device and packs to 2-bit planes + validity plane there. */
int igd_load_contigs(igd_handle *h, const uint8_t *seq, const uint64_t *offsets, int32_t n_contigs);
/* The same, returning as soon as the letters have crossed the bus (seq / offsets may be reused); the pack kernel stays
 * queued on the handle's stream and every later call of this handle is ordered behind it.  For pipelines that upload the
 * next set of contigs on one handle while another handle of the same GPU scans and scores (igdetective_b200/multi.py). */
int igd_load_contigs_async(igd_handle *h, const uint8_t *seq, const uint64_t *offsets, int32_t n_contigs);

/* Ingest ("next" row of the scope table): input_seq_dict = {rec.id: rec.seq for rec in
 * SeqIO.parse(INPUT_PATH, "fasta")} (py/IGDetective.py:129) done natively.  igd_fasta_open parses
 * an uncompressed FASTA file on the host: id = first whitespace-delimited token of the header,
 * sequence lines concatenated with blanks / CR / LF removed, case preserved; a repeated id keeps
 * its first position and takes the last sequence, like the dict comprehension.  The letters stay
 * in host memory for slicing (igd_fasta_read == parent_seq[contig][a:a+n]); igd_load_fasta copies
 * them to the device and packs them, like igd_load_contigs.  No GPU is needed to open a file. */
typedef struct igd_fasta igd_fasta;
igd_fasta *igd_fasta_open(const char *path);    /* NULL on failure (see igd_last_error) */
void       igd_fasta_close(igd_fasta *f);
int32_t    igd_fasta_count(const igd_fasta *f);
int        igd_fasta_info(const igd_fasta *f, int32_t i, const char **id, uint64_t *length);
int        igd_fasta_read(const igd_fasta *f, int32_t i, uint64_t start, uint64_t len, uint8_t *out);
int        igd_load_fasta(igd_handle *h, const igd_fasta *f);

/* Range form for sharded runs (one process per GPU, each parsing 1/N of the file): parses the LINES whose first byte
 * lies in [byte_lo, byte_hi) -- a line belongs to the range that holds its first byte, so consecutive ranges
 * partition the file -- plus margin_bytes of lines on either side (context for rows next to the range's ends).
 * Record 0 is the "lead": the letters in front of the first header, which continue a record that started earlier
 * (id ""; possibly empty); records 1.. are the headers met, in file order, nothing merged.  igd_fasta_range gives
 * the line-aligned byte range and the letters [own_lo, own_hi) of the concatenated records that come from it (the
 * rest is margin); igd_fasta_header_byte the file offset of a record's header line (UINT64_MAX for the lead), which
 * says whether the header lies in the range proper.  No GPU is needed. */
igd_fasta *igd_fasta_open_range(const char *path, uint64_t byte_lo, uint64_t byte_hi, uint64_t margin_bytes);
int        igd_fasta_range(const igd_fasta *f, uint64_t *byte_lo, uint64_t *byte_hi,
                           uint64_t *own_lo_letter, uint64_t *own_hi_letter);
int        igd_fasta_header_byte(const igd_fasta *f, int32_t i, uint64_t *byte);

/* find_valid_motif_idx + find_valid_rss over both strands and the four signal
 * types (py/IGDetective.py:138-177, call site :443) in one pass over the packed
 * contigs.  spacer[t] = SPACER_LENGTH of signal type t.  Hits stay on the
 * device until fetched; *n_hits is their number. */
int igd_scan(igd_handle *h, const int32_t spacer[4], uint64_t *n_hits);

/* Hits of the last igd_scan, sorted by (contig, sig_type, strand, index of the
 * 5' k-mer in scanned-strand coordinates) -- the order find_valid_rss would emit
 * if its first_set were iterated ascending. */
int igd_fetch_hits(igd_handle *h, igd_hit *out, uint64_t capacity);

/* find_valid_motif_idx (py/IGDetective.py:138-144) for k = 7 or 9: every window
 * whose upper-cased k-mer (or its reverse complement) is in any motif set of
 * that k.  Hits sorted by (contig, pos). */
int igd_scan_kmers(igd_handle *h, int32_t k, uint64_t *n_hits);
int igd_fetch_kmer_hits(igd_handle *h, igd_kmer_hit *out, uint64_t capacity);

/* ---- candidate-vs-gene scoring -------------------------------------------------------------- */

/* ALIGNER.align(target, query)[0] followed by BioAlign
 * (py/IGDetective.py:310-311, py/extract_aligned_genes.py:9-56,92-95) for a
 * list of (target, query) pairs.  Sequences are raw bytes (case preserved:
 * the DP compares bytes, the match count compares upper-cased bytes, as the
 * reference does).  Outputs, one per pair (any may be NULL except score):
 *   score   optimal global score
 *   matches number of equal (upper-cased) columns inside the gene range
 *           -- the TRUE count; BioAlign.NumMatches() returns matches-1
 *   alen    len(BioAlign) == AlignmentRange()[1] - AlignmentRange()[0]
 *   start   BioAlign.AlignmentRange()[0]  (columns before the first gene letter)
 *   ient    BioAlign.QuerySeq() == upper(target[start:ient])
 * With start == ient == NULL every pair whose fields provably fit runs in the
 * packed single-register kernel (about twice as fast); asking for start uses
 * the two-register kernel, asking for ient adds a second pass of it.
 */
int igd_align_batch(igd_handle *h,
                    const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                    const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                    const int32_t *pair_t, const int32_t *pair_q, int64_t n_pairs,
                    const igd_scoring *sc,
                    int32_t *score, int32_t *matches, int32_t *alen, int32_t *start, int32_t *ient);

/* The same for EVERY target against EVERY query (the shape of align_fragment_to_genes,
 * py/IGDetective.py:335-341, and of the iterative stage's loop, py/extract_aligned_genes.py:89-92):
 * outputs are [n_tgt][n_qry], target-major; no pair tables, no start / ient. */
int igd_align_dense(igd_handle *h,
                    const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                    const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                    const igd_scoring *sc, int32_t *score, int32_t *matches, int32_t *alen);

/* Replaces align_fragment_to_genes(fragments, canon_genes, scheme, gene), py/IGDetective.py:319-360,
 * i.e. ComputeAlignment (:307-317) for every (fragment, gene) pair followed by BioAlign.PI() / len()
 * (py/extract_aligned_genes.py:38-45): each fragment is upper-cased (ASCII), aligned as it is and
 * reverse-complemented (Bio.Seq's IUPAC table) against every gene; the forward orientation is kept
 * iff it has strictly more matches; an empty fragment stands for 'A' * len(gene) (:339-340).
 * Outputs are [n_frag][n_gene], fragment-major, float64 like the reference's matrices:
 *   pi        (matches - 1) / len * 100, evaluated left to right in float64
 *   alen      len(alignment) (the gene range)
 *   direction '+' or '-' per pair (may be NULL)
 * Both orientations run on the device, the choice is made there, and only one word per
 * (fragment, gene) travels back. */
int igd_score_fragments(igd_handle *h,
                        const uint8_t *frag, const int32_t *frag_off, int32_t n_frag,
                        const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                        const igd_scoring *sc, double *pi, double *alen, uint8_t *direction);

/* Replaces extract_aligned_genes.ComputeAlignment(aligner, [window, RC(window)], ['+', '-'], gene_seqs),
 * py/extract_aligned_genes.py:85-105, for many windows at once.  Windows are raw case (the DP compares
 * raw bytes, the match count upper-cased ones, :12-13); the reverse complement keeps the case.  Per
 * window, over the reference's scan order (orientation-major, genes inner) the LAST alignment with
 * maximal PI among PI >= 0 is kept (best_pi starts at 0 and the test is >=, :96-100):
 *   best   orientation * n_gene + gene index, or -1 when no alignment has PI >= 0 (Alignment stays empty)
 *   pi     its PI, float64 (matches-1)/len*100
 *   start, end   BioAlign.AlignmentRange() ; ient: QuerySeq() == upper(chosen query[start:ient])
 * The dense batch runs in the packed kernels, the choice is made on the device, and only the winners
 * go through the two-register kernel for their range. */
int igd_score_windows(igd_handle *h,
                      const uint8_t *win, const int32_t *win_off, int32_t n_win,
                      const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                      const igd_scoring *sc, int32_t *best, double *pi, int32_t *start, int32_t *end, int32_t *ient);

/* ---- the whole de-novo stage in one call ------------------------------------------------------ */

/* Module constants of py/IGDetective.py:29-32 that the stage reads (index 0 = V, 1 = J). */
typedef struct {
    int32_t gene_length[2];     /* GENE_LENGTH[V], [J]: 350, 70  -- letters cut next to the heptamer (:260-300) */
    int32_t d_gene_length;      /* GENE_LENGTH[D]: 150           -- combine_D_RSS window (:210-224) */
    double  pi_strict[2];       /* PI_CUTOFF['strict']: 70, 70   (:31, :387) */
    double  pi_relax[2];        /* PI_CUTOFF['relax']:  60, 65 */
    int32_t maxk_cutoff[2];     /* MAXK_CUTOFF: 15, 11           (:32) */
} igd_detect_params;

/* One row of genes_V.tsv / genes_J.tsv in integer form: an RSS whose best-scoring gene passed the cut-off
 * (py/IGDetective.py:463-477 argmax, :380-420 extract_genes).  The strings of the row are slices of the
 * caller's own contig letters. */
typedef struct {
    int64_t  hept, nona;        /* the RSS, forward-strand contig coordinates (columns 3, 4) */
    int64_t  frag_start;        /* first contig letter of the extracted fragment (forward coordinates) */
    uint32_t contig;
    int32_t  frag_len;          /* letters extracted; 0 = the reference's empty slice ('A' * len(gene) was scored) */
    int32_t  best_gene;         /* index of the first maximal PI in the gene table (np.argmax, :473) */
    int32_t  matches;           /* TRUE match count of the kept orientation; PI = (matches - 1) / alen * 100 */
    int32_t  alen;              /* len(BioAlign): 'longest common k-mer' column as float */
    int32_t  start;             /* AlignmentRange()[0] */
    int32_t  ient;              /* QuerySeq() == upper(kept orientation of the fragment)[start:ient] */
    uint8_t  gene_type;         /* 0 = V, 1 = J */
    uint8_t  strand;            /* IGD_STRAND_* of the RSS */
    uint8_t  direction;         /* '+' : the fragment as extracted aligned better, '-' : its reverse complement (:314-317) */
    uint8_t  reserved;
} igd_vj_row;

/* One row of genes_D.tsv: a (D_left, D_right) RSS pair of combine_D_RSS (py/IGDetective.py:210-224). */
typedef struct {
    int64_t  dl_hept, dl_nona, dr_hept, dr_nona;
    uint32_t contig;
    uint8_t  strand;
    uint8_t  reserved[3];
} igd_d_row;

/* The module body of py/IGDetective.py:439-487 on the contigs loaded in the handle, in ONE call:
 *   get_contigwise_rss x 8 (:443)            -> the RSS scan kernel (as igd_scan; the hits stay fetchable)
 *   combine_D_RSS (:444-446)                 -> sorted search over the hit list (host, O(n log n))
 *   get_s_fragment_from_RSS (:455)           -> a gather kernel over the resident contig letters: Python slice
 *                                               semantics incl. the negative-start quirk (:265), upper-cased,
 *                                               reverse-complemented for strand '-' (Bio.Seq's IUPAC table)
 *   align_fragment_to_genes (:467)           -> packed Gotoh kernels, both orientations, orientation choice and
 *                                               first-maximum argmax (:473) on the device
 *   threshold (:387), ComputeAlignment (:388) of the winners -> two-register Gotoh kernel for start / ient
 * Rows come back in the reference's order for ascending RSS lists: V rows then J rows, each strand '+' then
 * '-', contigs in load order, RSS list order; D rows strand '+' then '-', contigs in order, D_right outer /
 * D_left inner.  Needs igd_load_contigs / igd_load_fasta (the ASCII letters stay resident) and igd_set_motifs. */
int igd_detect(igd_handle *h, const int32_t spacer[4],
               const uint8_t *vgene, const int32_t *vgene_off, int32_t n_vgene,
               const uint8_t *jgene, const int32_t *jgene_off, int32_t n_jgene,
               const igd_scoring *sc, const igd_detect_params *prm,
               uint64_t *n_hits, uint64_t *n_vj_rows, uint64_t *n_d_rows);
int igd_fetch_vj_rows(igd_handle *h, igd_vj_row *out, uint64_t capacity);
int igd_fetch_d_rows(igd_handle *h, igd_d_row *out, uint64_t capacity);
/* The letters of the rows of the last igd_detect, read from the resident contigs with Python's slice semantics and
 * upper-cased (py/IGDetective.py:238-249, 391-419): three segments per V / J row, in row order -- heptamer, nonamer (reverse-
 * complemented for strand '-'), the aligned part of the fragment in its aligned orientation ('gene sequence') -- then five
 * per D row: left heptamer, left nonamer, right heptamer, right nonamer, the sequence between the heptamers.
 * Segment k = text[offsets[k] .. offsets[k + 1]); offsets has n_segments + 1 entries. */
int igd_row_text_size(igd_handle *h, uint64_t *n_bytes, uint64_t *n_segments);
int igd_fetch_row_text(igd_handle *h, uint8_t *text, uint64_t text_capacity, uint64_t *offsets, uint64_t offsets_capacity);

/* The two upper bounds igd_detect prunes its scoring with, for every (fragment, gene) pair -- not a reference
 * interface: exposed so that tests can hold the bound kernels against the plain recurrences.
 *   tgt rows 2u / 2u + 1: fragment u and its reverse complement (same length, 1..512); genes 1..511 letters.
 *   lcs[2u + o][g]  = LCS(upper(row 2u + o), upper(gene g)) >= the equal columns of ANY alignment, i.e. >= BioAlign's
 *                     NumMatches() + 1 (py/extract_aligned_genes.py:33-38), so PI <= (lcs - 1) / len(gene) * 100.
 *   ybound[u][g]    = Y(row 2u + 1, g) << 16 | Y(row 2u, g),  Y = max over all alignments of 2 x equal columns - fragment
 *                     letters inserted inside the gene range; PI >= p (p >= 50) implies Y >= p / 50 * len(gene) + 2.
 * Letters other than A, C, G, T (either case): one class that matches itself in the LCS; in Y such a fragment letter
 * matches every A, C, G, T and such a gene letter adds 3 (tests/kernel_model.py:y_bound) -- still upper bounds. */
int igd_match_bounds(igd_handle *h, const uint8_t *tgt, const int32_t *tgt_off, int32_t n_frag,
                     const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                     uint16_t *lcs, uint32_t *ybound);

#ifdef __cplusplus
}
#endif
#endif /* IGD_B200_H */
