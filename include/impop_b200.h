/* impop_b200 -- C ABI of the B200-native windowed population-statistics library.
 *
 * The reference (pangenome/impop) has NO library / plugin / FFI boundary on this path:
 * layers talk through process spawns and text (SURVEY.md section 8 b).  Each entry point
 * below therefore cites the reference *process or function* it replaces; INTEGRATION.md
 * shows the ctypes stub a maintainer of the reference scripts would add.
 *
 * Conventions
 *  - Every call returns 0 (IMPOP_OK) or a negative error code; impop_last_error(ctx) gives text.
 *  - The caller owns every buffer.  Pointers named *_dev are device pointers on the
 *    context's device; pointers named *_host are host pointers.  The library never
 *    frees caller memory and keeps no global state (one context per device / thread).
 *  - All work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL =
 *    default stream).  Calls are asynchronous unless stated; impop_check() synchronises
 *    the stream and reports device-side errors (bad input, barrier time-out).
 *  - Bit layout of a window's presence matrix: haplotype i, node k -> bit (k & 31) of the
 *    32-bit word x[x_off + i * pitch_words + (k >> 5)].  pitch_words and x_off must be
 *    multiples of 4 (16-byte rows).  Bits at k >= m are ignored (their weight is zero), so a window may be a
 *    column range of a wider matrix: x_off = its first 128-node group, pitch_words = the wide matrix's pitch,
 *    node_len = 0 for the nodes of that group that precede the window (impop_b200/chromosome.py).
 *  - Exactness: intersections, path lengths and unions are exact integers; a window must
 *    satisfy sum(node_len) < 2^31 (checked on device -> IMPOP_ERR_RANGE at impop_check).
 *  - There is no CPU fallback anywhere: without a CUDA device impop_create fails.
 */
#ifndef IMPOP_B200_H
#define IMPOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMPOP_OK 0
#define IMPOP_ERR_ARG (-1)      /* bad argument (null pointer, misaligned pitch, ...) */
#define IMPOP_ERR_CUDA (-2)     /* CUDA runtime error; see impop_last_error */
#define IMPOP_ERR_NOMEM (-3)
#define IMPOP_ERR_RANGE (-4)    /* window violates sum(node_len) < 2^31, or n/m limits */
#define IMPOP_ERR_DEVICE (-5)   /* device-side failure flag (barrier time-out) */

/* Haplotype label bits (one byte per haplotype).  In the window kernels a haplotype carrying both
 * IMPOP_LAB_A and IMPOP_LAB_B is dropped from both populations, as h-fst.py:181-185 does. */
#define IMPOP_LAB_SUBSET 1u /* counted in pi / n (pica2.py sample subset, run_tajd.sh -l list) */
#define IMPOP_LAB_A 2u      /* population A of h-fst.py -a */
#define IMPOP_LAB_B 4u      /* population B of h-fst.py -b */
#define IMPOP_LAB_SEG 8u    /* rows that define segregating nodes S (run_tajd.sh:126-148 uses all) */

/* Columns of one statistics row (fp64). */
#define IMPOP_NSTATS 20
#define IMPOP_ST_PI 0           /* pica2.py:154, all-singleton groups (threshold >= max identity) */
#define IMPOP_ST_PI_PER_SITE 1  /* pica2.py:163-164; NaN when L == 0 */
#define IMPOP_ST_PI_A 2         /* h-fst.py:197; divided by L when L > 0 (h-fst.py:225-240) */
#define IMPOP_ST_PI_B 3
#define IMPOP_ST_PI_XY 4
#define IMPOP_ST_DXY 5
#define IMPOP_ST_DA 6
#define IMPOP_ST_FST 7          /* h-fst.py:214-222 */
#define IMPOP_ST_S 8            /* segregating nodes (replaces run_tajd.sh:148) */
#define IMPOP_ST_TAJIMA_D 9     /* tj_d.py:47-69 on (n, S, per-site pi if L > 0 else pi) as run_tajd.sh:166-180 */
#define IMPOP_ST_A1 10
#define IMPOP_ST_E1 11
#define IMPOP_ST_E2 12
#define IMPOP_ST_N 13
#define IMPOP_ST_SUM_S 14       /* raw sum of pi_ij over subset pairs */
#define IMPOP_ST_SUM_AA 15
#define IMPOP_ST_SUM_BB 16
#define IMPOP_ST_SUM_AB 17
#define IMPOP_ST_TAJIMA_D_RAW 18 /* tj_d.py on (n, S, pi) with pi not divided by L */
#define IMPOP_ST_S_BUBBLES 19    /* variant sites as a bubble caller counts them (run_tajd.sh:126-148: records of `povu gfa2vcf`): maximal runs,
                                   in node order, of segregating nodes not interrupted by a node every SEG row carries; parity unpinned.
                                   After ingest-time compaction the node order is gone: pass the count taken before (site_runs_host). */

/* Columns of one counts row (int64). */
#define IMPOP_NCOUNTS 8 /* nS nA nB pairsS pairsAA pairsBB pairsAB S */

/* Pairwise-kernel implementations (both are CUDA kernels; there is no CPU path). */
#define IMPOP_ALGO_TCGEN05 0 /* tcgen05.mma kind::i8, accumulators in TMEM (default) */
#define IMPOP_ALGO_SIMT 1    /* dp4a shared-memory tiles; cross-check and bring-up path */

typedef struct impop_ctx impop_ctx_t;
typedef struct impop_batch impop_batch_t;

/* A batch of W windows.  Descriptor arrays are HOST arrays of length `windows`; the bulk
 * arrays are DEVICE pointers and must stay valid and unchanged (x, node_len) for the
 * lifetime of the batch.  labels may be rewritten between calls. */
typedef struct {
    int32_t windows;
    const int32_t *n_host;           /* haplotypes per window */
    const int32_t *m_host;           /* nodes per window */
    const int32_t *pitch_words_host; /* row pitch in 32-bit words (multiple of 4) */
    const int64_t *x_off_host;       /* offset of the window in x_dev, in 32-bit words (multiple of 4) */
    const int64_t *len_off_host;     /* offset of the window in node_len_dev (elements) */
    const int64_t *lab_off_host;     /* offset of the window in labels_dev (bytes) */
    const int64_t *length_host;      /* BED window length L per window (0 = no per-site normalisation) */
    const uint32_t *x_dev;           /* 16-byte aligned (rows are read 16 bytes at a time) */
    const uint32_t *node_len_dev;
    const uint8_t *labels_dev;
    const uint32_t *node_len_host;   /* optional host copy of node_len (same offsets): lets batch set-up run without
                                        any device->host read, i.e. fully asynchronously on `stream` */
    void *stream;                    /* cudaStream_t the set-up copies are enqueued on (NULL = default stream); use the
                                        stream the kernels will run on */
    const int64_t *site_runs_host;   /* optional [windows]: IMPOP_ST_S_BUBBLES of each window as counted at ingest on the
                                        original node order (impop_compact_scan); < 0 or NULL: counted on the device */
    /* Affine form of a window (optional; what impop_compact_fill writes with IMPOP_COMPACT_PAIRS): the intersection is
     *     I_ij = sum_k node_len_k x_ik x_jk + C - R_i - R_j ,   A_i = I_ii ,   U_ij = A_i + A_j - I_ij
     * with a window constant C and a row term R_i, all exact integers: nodes every haplotype visits live in C alone,
     * and the two branches of a bi-allelic bubble (complementary columns) share one column.  All three NULL: C = R = 0
     * and every column stands for one node.  Requirements: sum(node_len) < 2^31 and 0 <= I, A, U < 2^31 as before. */
    const int32_t *row_adj_dev;      /* R_i per haplotype row: the rows of all windows in batch order (window w starts at n_0 + .. + n_{w-1}) */
    const int64_t *win_const_host;   /* C per window */
    const uint8_t *col_mult_dev;     /* per column, offsets len_off_host: how many nodes of positive length the column stands
                                        for in the segregating-node count S (1 = a plain node, 2 = a merged bubble, 0 = a
                                        further copy of a column whose weight was split) */
    const int32_t *heavy_entries_host; /* optional [windows]: heavy-table entries of each window, sum over its nodes of
                                        ceil(floor(node_len / 255) / 255), when the caller knows it from ingest: batch set-up
                                        then never reads node_len on the host.  Too small a number is caught on the
                                        device (IMPOP_ERR_RANGE at impop_check), a larger one only costs scratch. */
} impop_batch_desc_t;

int impop_version(void);
int impop_create(int device, impop_ctx_t **ctx_out);
int impop_destroy(impop_ctx_t *ctx);
const char *impop_last_error(impop_ctx_t *ctx);
/* Synchronise `stream`, then report sticky device-side errors raised by earlier kernels. */
int impop_check(impop_ctx_t *ctx, void *stream);
/* Number of kernels this library has launched on ctx so far (bench.py's gpu_launches). */
int64_t impop_launch_count(impop_ctx_t *ctx);

/* Optional per-kernel device timing for bench.py's roofline line: when enabled, every launch of the
 * kernels below is bracketed by CUDA events on the caller's stream (no synchronisation added).
 * impop_timing_read sums the elapsed time of the launches of one kernel recorded since the last
 * impop_timing_enable call (which also clears the record). */
#define IMPOP_KERNEL_PREP 0      /* path lengths, byte weights, heavy-node table and bits, S, label counts */
#define IMPOP_KERNEL_PAIRS 1     /* fused pairwise + fp64 reduction (tcgen05 or SIMT) */
#define IMPOP_KERNEL_SUMS 2      /* per-window fixed-order sum of tile partials */
#define IMPOP_KERNEL_COLSTAT 3   /* (unused since the prep pass forms S and the label counts) */
#define IMPOP_KERNEL_FINALIZE 4  /* derived statistics */
#define IMPOP_KERNEL_SITES 5     /* per-site allele counts */
int impop_timing_enable(impop_ctx_t *ctx, int32_t enable);
int impop_timing_read(impop_ctx_t *ctx, int32_t kernel_id, double *total_ms, int64_t *launches);

/* K1 ingest.  Replaces the text hand-off `impg similarity ... > tmp.sim` (run_pica2_impg.sh:162-168):
 * a dense 0/1 coverage matrix (n x m bytes, row pitch dense_pitch bytes) becomes bit-packed rows. */
int impop_pack_bits(impop_ctx_t *ctx, const uint8_t *dense_dev, int32_t n, int32_t m, int64_t dense_pitch,
                    uint32_t *x_dev, int32_t pitch_words, void *stream);

/* Batch set-up: uploads the descriptor tables (asynchronously, from pinned staging), sizes the scratch (path
 * lengths, byte weights, heavy-node table and bits, per-item partial sums).  Device blocks come from a pool kept
 * by the context, so repeated create / destroy does not call cudaMalloc / cudaFree.  With node_len_host given the
 * call does not synchronise; without it there is one small device->host read.
 * impop_batch_destroy may be called while the batch's kernels are still in flight: its blocks go back to the pool
 * behind an event on the stream of the batch's last call, and are handed to work on ANOTHER stream only once that event
 * has completed (work on the same stream is ordered behind the kernels anyway). */
int impop_batch_create(impop_ctx_t *ctx, const impop_batch_desc_t *desc, impop_batch_t **batch_out);
int impop_batch_destroy(impop_ctx_t *ctx, impop_batch_t *batch);
int64_t impop_batch_items(const impop_batch_t *batch); /* number of 128 x (<= 256) tile work items */

/* K2+K3 fused.  Replaces, per window, `impg similarity` / `odgi similarity` (a-0) followed by
 * pica2.analyze_similarity_matrix (pica2.py:60-169, threshold >= max identity), h-fst.calculate_fst
 * (h-fst.py:173-249), the segregating-site count (run_tajd.sh:126-148) and tj_d.tajimas_d
 * (tj_d.py:47-69).  stats_dev: W x IMPOP_NSTATS fp64; counts_dev: W x IMPOP_NCOUNTS int64. */
int impop_window_stats(impop_ctx_t *ctx, impop_batch_t *batch, int32_t algo, double *stats_dev,
                       int64_t *counts_dev, void *stream);

/* The same in two steps, for splitting one batch's tile grid over several GPUs (SURVEY.md 8 e):
 * rank r of `world` processes work items t with t % world == r and writes raw sums
 * (W x 4: S, AA, BB, AB).  After an all-gather, impop_window_finalize adds `parts` such arrays
 * ([parts][W][4], fixed order => run-to-run reproducible) and derives the statistics; the label counts and
 * segregating-node counts it uses were formed by the preceding impop_window_sums call on this batch. */
int impop_window_sums(impop_ctx_t *ctx, impop_batch_t *batch, int32_t algo, int32_t rank, int32_t world,
                      double *sums_dev, void *stream);
int impop_window_finalize(impop_ctx_t *ctx, impop_batch_t *batch, const double *sums_dev, int32_t parts,
                          double *stats_dev, int64_t *counts_dev, void *stream);

/* Materialising variant for one window of the batch (`--dump-similarity`, bit-exact tests):
 * I_dev n x n int64 (diagonal = path length), A_dev n int64, pi_dev n x n fp64 (diagonal 0).
 * Any output pointer may be NULL.  The table `impg similarity` would print is (I, A) -> identity = 1 - pi. */
int impop_pairwise(impop_ctx_t *ctx, impop_batch_t *batch, int32_t window, int32_t algo, int64_t *I_dev,
                   int64_t *A_dev, double *pi_dev, void *stream);

/* K3 alone ("TSV mode"): reductions of pica2.py:118-164 / h-fst.py:130-249 over a dense identity
 * matrix (n x n fp64, row stride ld, NaN = pair absent from the table; only i < j is read).
 * weight_dev (nullable, n fp64): pica2 group frequencies |G|/N on representatives, 0 elsewhere;
 * when given, wsum_dev[0] = sum_{i<j} (1 - s_ij) * w_i * w_j over present pairs (pica2.py:137-139),
 * wsum_dev[1] = number of such pairs with w_i * w_j != 0, wsum_dev[2] = pica2's grouped
 * pi = n/(n-1) * 2 * wsum_dev[0] (pica2.py:154; 0 when no pair) and wsum_dev[3] = pi / length (NaN when length == 0).
 * With BOTH labels_dev and weight_dev the weighted sum runs over the pairs with one row in A and the other in B only
 * (hudson/hud.py:235-263: grouped Dxy, weights |G_a|/n_A and |G_b|/n_B on the group representatives).
 * wsum_dev holds 4 doubles.  stats_dev: IMPOP_NSTATS, counts_dev: IMPOP_NCOUNTS (S and Tajima columns use seg_sites as S). */
int impop_reduce_identity(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld,
                          const uint8_t *labels_dev, const double *weight_dev, int64_t length, double seg_sites,
                          double *stats_dev, int64_t *counts_dev, double *wsum_dev, void *stream);

/* tj_d.tajimas_d (tj_d.py:47-69) for `count` independent (n, S, pi) triples.
 * parts_dev (nullable): count x 10 = a1 a2 b1 b2 c1 c2 e1 e2 numerator denominator. */
int impop_tajima_d(impop_ctx_t *ctx, const int64_t *n_dev, const double *S_dev, const double *pi_dev,
                   int32_t count, double *D_dev, double *parts_dev, void *stream);

/* round(x, digits) of every element, in place, exactly as CPython rounds a float (pica2.py:81-83 and h-fst.py:149-150
 * call round(sim, r) on every table value): the double's decimal value rounded half-even at that decimal, then the
 * nearest double.  0 <= digits <= 22.  NaN (pair absent) and infinities stay. */
int impop_round_decimal(impop_ctx_t *ctx, double *values_dev, int64_t count, int32_t digits, void *stream);

/* Rows as they are stored and transferred -- tight: `src_pitch_words` >= ceil(m / 32) words per haplotype, any number --
 * into the rows the kernels read (16-byte multiples: `dst_pitch_words` a multiple of 4, >= src), zero padded.  A window of
 * 286 columns travels as 9 words per row instead of 12: a quarter of the presence bits less to upload.  All tables are
 * HOST arrays of length `windows` (offsets in 32-bit words into src_dev / dst_dev); one small table upload and one kernel
 * on `stream`.  No counterpart in the reference (the hand-off there is text). */
int impop_repitch_rows(impop_ctx_t *ctx, int32_t windows, const int32_t *rows_host, const int32_t *src_pitch_words_host,
                       const int32_t *dst_pitch_words_host, const int64_t *src_off_host, const int64_t *dst_off_host,
                       const uint32_t *src_dev, uint32_t *dst_dev, void *stream);

/* Plain device memory for callers without a tensor library (the TSV-mode command lines: scripts/pica2.py, h-fst.py,
 * af.py, tj_d.py, hud.py run without importing torch).  impop_dev_copy: kind 0 host -> device, 1 device -> host,
 * 2 device -> device; synchronous with respect to `stream`. */
int impop_dev_alloc(impop_ctx_t *ctx, int64_t bytes, void **ptr_out);
int impop_dev_free(impop_ctx_t *ctx, void *ptr);
int impop_dev_copy(impop_ctx_t *ctx, void *dst, const void *src, int64_t bytes, int32_t kind, void *stream);

/* K4, BASELINE config 4: per-site allele counts of a site-major bit matrix (sites x words u64,
 * bit h of a row = haplotype h carries the allele) under `pops` population masks (pops x words u64).
 * counts_dev sites x pops int32; freq_dev (nullable) sites x pops fp64 = count / |mask|.
 * The only per-site allele-count semantics in the reference is scripts/wip/op-afs.py:26-45. */
int impop_site_counts(impop_ctx_t *ctx, const uint64_t *sites_dev, int64_t sites, int32_t words,
                      const uint64_t *masks_dev, int32_t pops, int32_t *counts_dev, double *freq_dev, void *stream);

/* K5: af.cluster (af.py:35-44): connected components of {identity >= threshold} over a dense
 * identity matrix (NaN = absent).  comp_dev[i] = smallest member index of i's component. */
int impop_cluster(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld, double threshold,
                  int32_t *comp_dev, void *stream);

/* pica2.analyze_similarity_matrix step 1 (pica2.py:94-112): greedy star grouping on `identity > threshold`
 * (strict), seeds taken in index order (= sorted-name order; the reference's set.pop() order is
 * hash-seed dependent).  group_dev[i] = index of the seed of i's group (its smallest
 * member, the representative pica2.py:128 uses); weight_dev (nullable, n fp64) = |G|/n on seeds, 0 elsewhere. */
int impop_greedy_groups(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld, double threshold,
                        int32_t *group_dev, double *weight_dev, void *stream);

/* ---- Ingest (host side, no device work): GFA v1 text of one window -> presence matrix ------------------------
 * Replaces the text hand-off in front of the similarity tool: `odgi similarity -i tmp.gfa` walks the paths of the
 * extracted window graph (run_pica2_odgi.sh:60-96; `impg similarity -r REGION`, run_h-fst.sh:65-67, does the same
 * from alignments).  Node k = the k-th S line (length = its sequence, or LN:i: when the sequence is '*'); one
 * matrix row per P or W line in file order, as odgi similarity without grouping flags makes one group per path.
 * Row names: the P line's path name as it stands (e.g. `HG00097#1#CM094061.1:109468899-109469099`, h-fst.py:21-22),
 * or `sample#hap#seqid[:start-end]` for a W line.  Orientations are ignored (presence of the node, as the
 * similarity tools count it).  A malformed line or a step over an undefined segment -> IMPOP_ERR_ARG with its
 * 1-based line number. */
typedef struct {
    int64_t segments;    /* S lines = nodes m */
    int64_t paths;       /* P + W lines = matrix rows n */
    int64_t name_bytes;  /* bytes of all row names, each NUL-terminated */
    int64_t steps;       /* path steps in total */
    int64_t error_line;  /* 0, or the line impop_gfa_scan stopped at */
} impop_gfa_info_t;
int impop_gfa_scan(const char *text, int64_t bytes, impop_gfa_info_t *info);
/* Caller-owned host buffers sized from impop_gfa_scan: x_bits [paths x pitch_words] u32 (pitch_words a multiple of 4,
 * 32 * pitch_words >= segments; zeroed here), node_len [segments], names [name_bytes] with name_off [paths + 1],
 * counts (optional, may be NULL) [paths x segments] u16 = how often the path visits the node (saturating; the
 * multiset coverage a cyclic graph needs).  revisits (optional): set to 1 when some path visits a node more than once,
 * else 0 -- a reader that wants visit counts only for such windows parses without them first (a chromosome's worth of
 * count matrices is gigabytes). */
int impop_gfa_fill(const char *text, int64_t bytes, int32_t pitch_words, uint32_t *x_bits_host, uint32_t *node_len_host,
                   uint16_t *counts_host, char *names_host, int64_t *name_off_host, int64_t *error_line, int32_t *revisits);

/* ---- Ingest (host side, no device work): column compaction of a batch of windows ---------------------------------
 * The similarity tools walk paths (run_pica2_odgi.sh:96 `odgi similarity -i tmp.gfa`, run_h-fst.sh:65-67); a presence
 * MATRIX of the same window carries columns that cannot change any result, and this step removes them once, before the
 * matrices are uploaded: nodes visited by every haplotype are merged into one node of their summed length (they add
 * the same constant to every intersection and path length), nodes visited by none and nodes of length 0 are dropped,
 * the rest is ordered by length (original order within equal lengths).  Exact: I, A, U, segregating-node counts and all
 * statistics of the compacted window equal those of the original.
 * flags = IMPOP_COMPACT_PAIRS writes the affine form of impop_batch_desc_t instead (row_adj_out / win_const_out /
 * col_mult_out are then required): the constant nodes go into C, nodes with identical presence columns are merged, and
 * two columns that are complementary over the window's rows -- the two branches of a bi-allelic bubble, x_r = 1 - x_a --
 * become ONE column of weight len_r + len_a with len_r added to C and len_r x_ai to R_i
 * (len_r x_ri x_rj + len_a x_ai x_aj = len_r - len_r x_ai - len_r x_aj + (len_r + len_a) x_ai x_aj): an HPRC-shaped
 * window keeps about a quarter of its columns.  | IMPOP_COMPACT_REPLICATE additionally spreads a weight >= 255 over
 * copies of its column with byte weights (col_mult 0 for the further copies) when the copies fit into the padding of
 * the window's 128-column chunks, so that the device needs no separate heavy columns for it.  Still exact.
 * Descriptor arrays as in impop_batch_desc_t, but every pointer is a HOST pointer.  impop_compact_scan reports the node
 * count of every compacted window; the caller sizes the outputs (out_pitch_words[w] * 32 >= m_out[w], a multiple of 4
 * words; len_out zero-filled beyond m_out; col_mult_out shares out_len_off, row_adj_out starts at out_row_off[w]) and
 * impop_compact_fill writes them (it reuses the plans of a scan over the same arrays and flags).  `threads` host
 * threads share the windows.  IMPOP_ERR_RANGE: a window's visited nodes sum to 2^31 or more. */
#define IMPOP_COMPACT_PAIRS 1u
#define IMPOP_COMPACT_REPLICATE 2u
int impop_compact_scan(int32_t windows, const int32_t *n, const int32_t *m, const int32_t *pitch_words, const int64_t *x_off,
                       const int64_t *len_off, const uint32_t *x_bits, const uint32_t *node_len, int32_t threads, uint32_t flags,
                       int32_t *m_out, int64_t *site_runs_out /* nullable: IMPOP_ST_S_BUBBLES over all rows, original order */);
int impop_compact_fill(int32_t windows, const int32_t *n, const int32_t *m, const int32_t *pitch_words, const int64_t *x_off,
                       const int64_t *len_off, const uint32_t *x_bits, const uint32_t *node_len, int32_t threads, uint32_t flags,
                       const int32_t *out_pitch_words, const int64_t *out_x_off, const int64_t *out_len_off,
                       uint32_t *x_out, uint32_t *len_out, const int64_t *out_row_off /* nullable with flags == 0 */,
                       int32_t *row_adj_out, int64_t *win_const_out, uint8_t *col_mult_out);

/* All-pairs similarity table (the TSV the similarity tools print; pica2.py:6-58 and h-fst.py:84-119 parse it with
 * csv.DictReader, 47-92 % of those scripts' run time at 466 haplotypes) -> names in sorted order + dense n x n identity
 * matrix (NaN = pair absent; a repeated pair keeps its last row, pica2.py:44), the form impop_reduce_identity takes.
 * Columns `group.a`, `group.b`, `estimated.identity` are located by name, extras ignored.  Only machine-clean text is
 * accepted (ASCII, no quote characters, every row wide enough, plain decimal / exponent numbers or nan / inf):
 * otherwise status = 1 and the caller falls back to its general reader, which mirrors the reference's handling of
 * odd input line by line.  impop_tsv_fill returns IMPOP_ERR_ARG on text impop_tsv_scan did not report clean. */
typedef struct {
    int64_t rows;        /* data rows */
    int64_t names;       /* distinct names n */
    int64_t name_bytes;  /* bytes of all names, each NUL-terminated */
    int32_t status;      /* 0 clean, 1 use the general reader */
    int32_t reserved;
} impop_tsv_info_t;
int impop_tsv_scan(const char *text, int64_t bytes, impop_tsv_info_t *info);
int impop_tsv_fill(const char *text, int64_t bytes, double *matrix_host, char *names_host, int64_t *name_off_host);

/* Device self-test: the epilogue's range-restricted division against __ddiv_rn on `count` pseudo-random
 * in-range operand triples (I, A_i, A_j).  *mismatches_host must come back 0.  Synchronous. */
int impop_selftest_division(impop_ctx_t *ctx, uint64_t seed, int64_t count, int64_t *mismatches_host, void *stream);

/* Development aid: cycle counters (wait / work / other) of the pairs kernel's warp roles in the last launch, 16
 * int64 per CTA; all zero unless the library was built with -DIMPOP_PROFILE_ROLES.  Synchronous. */
int impop_debug_role_times(impop_ctx_t *ctx, int64_t *out_host, int32_t ctas);

#ifdef __cplusplus
}
#endif
#endif /* IMPOP_B200_H */
