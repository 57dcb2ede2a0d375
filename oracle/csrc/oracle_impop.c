/* CPU oracle (plain C) for the impop windowed population-statistics hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Linked / loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Row a-0 (the all-pairs weighted intersection / union / estimated.identity) restates
 * arithmetic of `odgi similarity` / `impg similarity`, tools that are not in
 * /root/reference and whose version the reference does not pin: PARITY UNPINNED for
 * that row (definition: SURVEY.md section 8 a-0; call sites run_pica2_impg.sh:162-168,
 * run_h-fst.sh:65-67, run_tajd.sh:160).  The reductions restate scripts/pica2.py:118-164,
 * scripts/h-fst.py:130-249 and scripts/tj_d.py:41-69 and are pinned through
 * tests/test_oracle_c.py against oracle/popstats.py, itself pinned to the reference's
 * golden vectors.
 *
 * Layout (same as include/impop_b200.h): haplotype i, node k  ->  bit (k & 31) of
 * 32-bit word x[i * pitch_words + (k >> 5)].
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, so every
 * fp64 operation below is a single correctly-rounded IEEE operation).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LAB_SUBSET 1u
#define LAB_A 2u
#define LAB_B 4u
#define LAB_SEG 8u
#define ORACLE_NSTATS 20
#define ORACLE_NCOUNTS 8

/* --------------------------------------------------------------------------------
 * a-0, plain definition: I_ij = sum_k len_k x_ik x_jk by walking the set bits.
 * ------------------------------------------------------------------------------ */
int64_t oracle_intersection_naive(const uint32_t *xi, const uint32_t *xj, const uint32_t *len, int m) {
    int64_t acc = 0;
    int words = (m + 31) / 32;
    for (int w = 0; w < words; ++w) {
        uint32_t c = xi[w] & xj[w];
        while (c) {
            int b = __builtin_ctz(c);
            int k = w * 32 + b;
            if (k < m) acc += (int64_t)len[k];
            c &= c - 1;
        }
    }
    return acc;
}

/* pi_ij from exact integers, in the contract's operation order:
 *   J = (double)I / (double)U ; id = 2.0*J / (1.0 + J) ; pi = 1.0 - id ; U == 0 -> J = 0 */
double oracle_pi_from_counts(int64_t inter, int64_t ai, int64_t aj) {
    int64_t uni = ai + aj - inter;
    double jac = (uni == 0) ? 0.0 : (double)inter / (double)uni;
    double ident = (2.0 * jac) / (1.0 + jac);
    return 1.0 - ident;
}

double oracle_identity_from_counts(int64_t inter, int64_t ai, int64_t aj) {
    int64_t uni = ai + aj - inter;
    double jac = (uni == 0) ? 0.0 : (double)inter / (double)uni;
    return (2.0 * jac) / (1.0 + jac);
}

/* Same intersections, faster: per byte column a 256-entry table of summed lengths. */
typedef struct {
    int nbytes;
    int64_t *tab; /* nbytes x 256 */
} byte_lut_t;

static int lut_build(byte_lut_t *lut, const uint32_t *len, int m, int pitch_words) {
    lut->nbytes = pitch_words * 4;
    lut->tab = (int64_t *)malloc((size_t)lut->nbytes * 256 * sizeof(int64_t));
    if (!lut->tab) return -1;
    for (int b = 0; b < lut->nbytes; ++b) {
        int64_t *t = lut->tab + (size_t)b * 256;
        t[0] = 0;
        for (int v = 1; v < 256; ++v) {
            int low = __builtin_ctz((unsigned)v);
            int k = b * 8 + low;
            t[v] = t[v & (v - 1)] + ((k < m) ? (int64_t)len[k] : 0);
        }
    }
    return 0;
}

static inline int64_t lut_intersection(const byte_lut_t *lut, const uint32_t *xi, const uint32_t *xj, int pitch_words) {
    int64_t acc = 0;
    for (int w = 0; w < pitch_words; ++w) {
        uint32_t c = xi[w] & xj[w];
        if (!c) continue;
        const int64_t *t = lut->tab + (size_t)w * 4 * 256;
        acc += t[c & 255u] + t[256 + ((c >> 8) & 255u)] + t[512 + ((c >> 16) & 255u)] + t[768 + (c >> 24)];
    }
    return acc;
}

/* Materialise one window: A[n], I[n*n] (may be NULL), pi[n*n] (may be NULL).  use_lut=0 -> plain loop. */
int oracle_window_pairwise(const uint32_t *x, int n, int m, int pitch_words, const uint32_t *len,
                           int64_t *A, int64_t *I, double *pi, int use_lut) {
    byte_lut_t lut = {0, NULL};
    if (use_lut && lut_build(&lut, len, m, pitch_words)) return -1;
    for (int i = 0; i < n; ++i) {
        const uint32_t *xi = x + (size_t)i * pitch_words;
        A[i] = use_lut ? lut_intersection(&lut, xi, xi, pitch_words) : oracle_intersection_naive(xi, xi, len, m);
    }
    for (int i = 0; i < n; ++i) {
        const uint32_t *xi = x + (size_t)i * pitch_words;
        for (int j = i; j < n; ++j) {
            const uint32_t *xj = x + (size_t)j * pitch_words;
            int64_t v = use_lut ? lut_intersection(&lut, xi, xj, pitch_words) : oracle_intersection_naive(xi, xj, len, m);
            if (I) { I[(size_t)i * n + j] = v; I[(size_t)j * n + i] = v; }
            if (pi) {
                double p = (i == j) ? 0.0 : oracle_pi_from_counts(v, A[i], A[j]);
                pi[(size_t)i * n + j] = p; pi[(size_t)j * n + i] = p;
            }
        }
    }
    free(lut.tab);
    return 0;
}

/* Python >= 3.12 builtin sum() over floats (Neumaier), which tj_d.py:41-45 calls. */
typedef struct { double total, comp; } nsum_t;
static inline void nsum_add(nsum_t *s, double x) {
    double t = s->total + x;
    if (fabs(s->total) >= fabs(x)) s->comp += (s->total - t) + x;
    else s->comp += (x - t) + s->total;
    s->total = t;
}
static inline double nsum_value(const nsum_t *s) {
    double r = s->total;
    if (s->comp != 0.0 && isfinite(s->comp)) r += s->comp;
    return r;
}

/* a-7: tj_d.py:47-69.  parts = a1 a2 b1 b2 c1 c2 e1 e2 numerator denominator (may be NULL). */
double oracle_tajimas_d(int64_t n, double S, double pi, double *parts) {
    nsum_t s1 = {0.0, 0.0}, s2 = {0.0, 0.0};
    for (int64_t i = 1; i < n; ++i) {
        nsum_add(&s1, 1.0 / (double)i);
        nsum_add(&s2, 1.0 / ((double)i * (double)i));
    }
    double dn = (double)n;
    double a1 = nsum_value(&s1), a2 = nsum_value(&s2);
    double b1 = (dn + 1.0) / (3.0 * (dn - 1.0));
    double b2 = 2.0 * (dn * dn + dn + 3.0) / (9.0 * dn * (dn - 1.0));
    double c1 = b1 - (1.0 / a1);
    double c2 = b2 - ((dn + 2.0) / (a1 * dn)) + (a2 / (a1 * a1));
    double e1 = c1 / a1;
    double e2 = c2 / (a1 * a1 + a2);
    double num = pi - (S / a1);
    double den = (S > 0) ? sqrt(e1 * S + e2 * S * (S - 1.0)) : NAN;
    double d = (den != 0.0 && den == den) ? num / den : NAN;
    if (parts) {
        parts[0] = a1; parts[1] = a2; parts[2] = b1; parts[3] = b2; parts[4] = c1;
        parts[5] = c2; parts[6] = e1; parts[7] = e2; parts[8] = num; parts[9] = den;
    }
    return d;
}

/* a-8 replacement: segregating nodes among rows whose label has LAB_SEG. */
int64_t oracle_segregating_nodes(const uint32_t *x, int n, int m, int pitch_words, const uint32_t *len,
                                 const uint8_t *labels) {
    int64_t S = 0;
    int words = (m + 31) / 32;
    for (int w = 0; w < words; ++w) {
        uint32_t any = 0, all = 0xffffffffu;
        int rows = 0;
        for (int i = 0; i < n; ++i) {
            if (!(labels[i] & LAB_SEG)) continue;
            uint32_t v = x[(size_t)i * pitch_words + w];
            any |= v; all &= v; ++rows;
        }
        if (!rows) continue;
        uint32_t seg = any & ~all;
        while (seg) {
            int k = w * 32 + __builtin_ctz(seg);
            if (k < m && len[k] > 0) ++S;
            seg &= seg - 1;
        }
    }
    return S;
}

/* Variant sites as a bubble caller would count them (run_tajd.sh:126-148 counts `povu gfa2vcf` records; povu is not in the
 * reference tree: parity unpinned): maximal runs, in node order, of segregating nodes (among the LAB_SEG rows, length > 0) not
 * interrupted by a node every LAB_SEG row carries; nodes none of them carries and zero-length nodes are transparent. */
int64_t oracle_site_runs(const uint32_t *x, int n, int m, int pitch_words, const uint32_t *len, const uint8_t *labels) {
    int64_t runs = 0;
    int in_run = 0, rows = 0;
    for (int i = 0; i < n; ++i) rows += (labels[i] & LAB_SEG) != 0;
    if (!rows) return 0;
    for (int k = 0; k < m; ++k) {
        if (len[k] == 0) continue;
        int c = 0;
        for (int i = 0; i < n; ++i)
            if ((labels[i] & LAB_SEG) && ((x[(size_t)i * pitch_words + (k >> 5)] >> (k & 31)) & 1u)) ++c;
        if (c == rows) in_run = 0;
        else if (c > 0) { if (!in_run) { ++runs; in_run = 1; } }
    }
    return runs;
}

/* Derived statistics from raw sums -- the definition the device finalize kernel mirrors.
 * stats[20]: 0 pi 1 pi_per_site 2 pi_a 3 pi_b 4 pi_xy 5 dxy 6 da 7 fst 8 S 9 tajima_d 10 a1 11 e1 12 e2 13 n
 *            14 sum_S 15 sum_AA 16 sum_BB 17 sum_AB 18 tajima_d_raw 19 S_bubbles (set by oracle_window_stats)
 * counts[8]: nS nA nB pairsS pairsAA pairsBB pairsAB S */
void oracle_finalize(const double sums[4], const int64_t cnt[8], int64_t L, double *stats) {
    int64_t nS = cnt[0];
    double pi = 0.0;
    if (nS >= 2 && cnt[3] > 0) {
        double dn = (double)nS;
        double f = 1.0 / dn;                                    /* pica2.py:137-138, every group a singleton */
        pi = (dn / (dn - 1.0)) * (2.0 * ((sums[0] * f) * f));   /* pica2.py:154 */
    }
    double pps = (L > 0) ? pi / (double)L : NAN;                /* pica2.py:161-164 */
    double pi_a = cnt[4] > 0 ? sums[1] / (double)cnt[4] : 0.0;  /* h-fst.py:168-171 */
    double pi_b = cnt[5] > 0 ? sums[2] / (double)cnt[5] : 0.0;
    double pi_xy = 0.5 * (pi_a + pi_b);                         /* h-fst.py:203 */
    double dxy = cnt[6] > 0 ? sums[3] / (double)cnt[6] : 0.0;
    double fst = dxy > 0 ? (dxy - pi_xy) / dxy : 0.0;           /* h-fst.py:214-222 */
    double da = dxy - pi_xy;
    if (L > 0) {                                                /* h-fst.py:225-240 */
        double dl = (double)L;
        pi_a = pi_a / dl; pi_b = pi_b / dl; pi_xy = pi_xy / dl; dxy = dxy / dl; da = da / dl;
    }
    double parts[10];
    double S = (double)cnt[7];
    double d = NAN, d_raw = NAN;
    for (int k = 0; k < 10; ++k) parts[k] = NAN;
    if (nS >= 2) {
        d_raw = oracle_tajimas_d(nS, S, pi, parts);
        d = (L > 0) ? oracle_tajimas_d(nS, S, pps, NULL) : d_raw;   /* run_tajd.sh:166-180 passes per-site pi */
    }
    stats[0] = pi; stats[1] = pps; stats[2] = pi_a; stats[3] = pi_b; stats[4] = pi_xy; stats[5] = dxy;
    stats[6] = da; stats[7] = fst; stats[8] = S; stats[9] = d; stats[10] = parts[0]; stats[11] = parts[6];
    stats[12] = parts[7]; stats[13] = (double)nS; stats[14] = sums[0]; stats[15] = sums[1]; stats[16] = sums[2];
    stats[17] = sums[3]; stats[18] = d_raw; stats[19] = 0.0;
}

/* h-fst.py:181-185: a sequence listed in both populations is removed from both. */
static inline unsigned clean_label(unsigned f) {
    return ((f & (LAB_A | LAB_B)) == (LAB_A | LAB_B)) ? (f & ~(LAB_A | LAB_B)) : f;
}

/* One window end to end (fused: nothing n x n is stored). */
int oracle_window_stats(const uint32_t *x, int n, int m, int pitch_words, const uint32_t *len,
                        const uint8_t *labels, int64_t L, double *stats, int64_t *counts) {
    byte_lut_t lut;
    if (lut_build(&lut, len, m, pitch_words)) return -1;
    int64_t *A = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    if (!A) { free(lut.tab); return -1; }
    int64_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const uint32_t *xi = x + (size_t)i * pitch_words;
        A[i] = lut_intersection(&lut, xi, xi, pitch_words);
        cnt[0] += (clean_label(labels[i]) & LAB_SUBSET) != 0;
        cnt[1] += (clean_label(labels[i]) & LAB_A) != 0;
        cnt[2] += (clean_label(labels[i]) & LAB_B) != 0;
    }
    nsum_t acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int i = 0; i < n; ++i) {
        const uint32_t *xi = x + (size_t)i * pitch_words;
        unsigned li = clean_label(labels[i]);
        if (!(li & (LAB_SUBSET | LAB_A | LAB_B))) continue;
        for (int j = i + 1; j < n; ++j) {
            unsigned lj = clean_label(labels[j]);
            int inS = (li & lj & LAB_SUBSET) != 0;
            int inAA = (li & lj & LAB_A) != 0;
            int inBB = (li & lj & LAB_B) != 0;
            int inAB = ((li & LAB_A) && (lj & LAB_B)) || ((li & LAB_B) && (lj & LAB_A));
            if (!(inS | inAA | inBB | inAB)) continue;
            int64_t v = lut_intersection(&lut, xi, x + (size_t)j * pitch_words, pitch_words);
            double p = oracle_pi_from_counts(v, A[i], A[j]);
            if (inS) { nsum_add(&acc[0], p); ++cnt[3]; }
            if (inAA) { nsum_add(&acc[1], p); ++cnt[4]; }
            if (inBB) { nsum_add(&acc[2], p); ++cnt[5]; }
            if (inAB) { nsum_add(&acc[3], p); ++cnt[6]; }
        }
    }
    cnt[7] = oracle_segregating_nodes(x, n, m, pitch_words, len, labels);
    double sums[4];
    for (int k = 0; k < 4; ++k) sums[k] = nsum_value(&acc[k]);
    oracle_finalize(sums, cnt, L, stats);
    stats[19] = (double)oracle_site_runs(x, n, m, pitch_words, len, labels);
    if (counts) memcpy(counts, cnt, sizeof(cnt));
    free(A);
    free(lut.tab);
    return 0;
}

/* Batch over windows with per-window descriptors, fanned out over `threads` pthreads. */
typedef struct {
    const uint32_t *x; const uint32_t *len; const uint8_t *labels;
    const int64_t *x_off, *len_off, *lab_off, *L;
    const int32_t *n, *m, *pitch;
    double *stats; int64_t *counts;
    int W, tid, threads, rc;
} job_t;

static void *batch_worker(void *arg) {
    job_t *jb = (job_t *)arg;
    for (int w = jb->tid; w < jb->W; w += jb->threads) {
        int rc = oracle_window_stats(jb->x + jb->x_off[w], jb->n[w], jb->m[w], jb->pitch[w], jb->len + jb->len_off[w],
                                     jb->labels + jb->lab_off[w], jb->L[w], jb->stats + (size_t)w * ORACLE_NSTATS,
                                     jb->counts ? jb->counts + (size_t)w * ORACLE_NCOUNTS : NULL);
        if (rc) jb->rc = rc;
    }
    return NULL;
}

int oracle_batch_stats(int W, const int32_t *n, const int32_t *m, const int32_t *pitch, const int64_t *x_off,
                       const int64_t *len_off, const int64_t *lab_off, const int64_t *L, const uint32_t *x,
                       const uint32_t *len, const uint8_t *labels, double *stats, int64_t *counts, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < threads; ++t) {
        job_t jb = {x, len, labels, x_off, len_off, lab_off, L, n, m, pitch, stats, counts, W, t, threads, 0};
        jobs[t] = jb;
        if (pthread_create(&th[t], NULL, batch_worker, &jobs[t])) return -2;
    }
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    return rc;
}

/* BASELINE config 4: counts[s*P+p] = popcount(site_bits[s] & mask[p]); freq = count / popsize. */
void oracle_site_counts(const uint64_t *sites, int64_t M, int words, const uint64_t *masks, int P,
                        int32_t *counts, double *freq) {
    int64_t size[64];
    for (int p = 0; p < P && p < 64; ++p) {
        size[p] = 0;
        for (int w = 0; w < words; ++w) size[p] += __builtin_popcountll(masks[(size_t)p * words + w]);
    }
    for (int64_t s = 0; s < M; ++s) {
        for (int p = 0; p < P; ++p) {
            int c = 0;
            for (int w = 0; w < words; ++w) c += __builtin_popcountll(sites[(size_t)s * words + w] & masks[(size_t)p * words + w]);
            counts[s * P + p] = c;
            if (freq) freq[s * P + p] = size[p] > 0 ? (double)c / (double)size[p] : 0.0;
        }
    }
}
