// Auxiliary kernels: bit packing (ingest), TSV-mode reductions over a dense identity matrix,
// stand-alone Tajima's D, per-site allele counts, and threshold clustering.
#include "common.cuh"
#include "stats_math.cuh"

namespace impop {

// ------------------------------------------------------------------------------------------
// K1: dense 0/1 bytes -> bit-packed rows.  One warp per 32 columns: coalesced byte loads, ballot.
// ------------------------------------------------------------------------------------------
__global__ void pack_bits_kernel(const uint8_t *dense, int32_t n, int32_t m, int64_t dpitch, uint32_t *x,
                                 int32_t pitch_words) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n * pitch_words;
    for (int64_t t = warp_global; t < total; t += nwarps) {
        const int i = (int)(t / pitch_words), wd = (int)(t % pitch_words);
        const int k = wd * 32 + lane;
        const bool bit = (k < m) && dense[(size_t)i * dpitch + k] != 0;
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) x[(size_t)i * pitch_words + wd] = word;
    }
}

// ------------------------------------------------------------------------------------------
// K3 alone (TSV mode).  Stage 1: one CTA per row stripe, fixed thread->pair mapping, partials to
// scratch; stage 2: one CTA adds the partials in order and finalizes.  Deterministic.
// ------------------------------------------------------------------------------------------
constexpr int RI_THREADS = 256;
constexpr int RI_SUMS = 5;    // S AA BB AB weighted: compensated (hi, lo)
constexpr int RI_CNTS = 5;    // matching pair counts (exact small integers in fp64)
constexpr int RI_VALS = 2 * RI_SUMS + RI_CNTS;   // per block: hi[5], lo[5], counts[5]

__global__ void __launch_bounds__(RI_THREADS) reduce_identity_stage1(const double *ident, int32_t n, int64_t ld,
                                                                     const uint8_t *labels, const double *weight,
                                                                     double *partials) {
    __shared__ dd s_sum[RI_THREADS / 32][RI_SUMS];
    __shared__ double s_cnt[RI_THREADS / 32][RI_CNTS];
    dd v[RI_SUMS];
    double c[RI_CNTS];
#pragma unroll
    for (int k = 0; k < RI_SUMS; ++k) { v[k].hi = 0.0; v[k].lo = 0.0; c[k] = 0.0; }
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t fi = labels ? labels[i] : 0u;
        const double wi = weight ? weight[i] : 0.0;
        for (int j = i + 1 + threadIdx.x; j < n; j += RI_THREADS) {
            const double s = ident[(size_t)i * ld + j];
            if (s != s) continue;  // pair absent from the table: skipped, not counted (h-fst.py:147-153)
            const double p = __dadd_rn(1.0, -s);
            const uint32_t fj = labels ? labels[j] : 0u;
            if (fi & fj & IMPOP_LAB_SUBSET) { dd_add(v[0], p); c[0] += 1.0; }
            if (fi & fj & IMPOP_LAB_A) { dd_add(v[1], p); c[1] += 1.0; }
            if (fi & fj & IMPOP_LAB_B) { dd_add(v[2], p); c[2] += 1.0; }
            if (((fi & IMPOP_LAB_A) && (fj & IMPOP_LAB_B)) || ((fi & IMPOP_LAB_B) && (fj & IMPOP_LAB_A))) {
                dd_add(v[3], p); c[3] += 1.0;
            }
            if (weight) {
                const double wj = weight[j];
                // labels given as well: only pairs with one row in A and the other in B (hud.py:235-263, grouped Dxy)
                const bool counted = !labels || ((fi & IMPOP_LAB_A) && (fj & IMPOP_LAB_B)) || ((fi & IMPOP_LAB_B) && (fj & IMPOP_LAB_A));
                if (counted && wi != 0.0 && wj != 0.0) {
                    dd_add(v[4], __dmul_rn(__dmul_rn(p, wi), wj));  // pica2.py:139
                    c[4] += 1.0;
                }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < RI_SUMS; ++k) {
        dd t = warp_sum_dd(v[k]);
        double tc = warp_sum(c[k]);
        if (lane == 0) { s_sum[warp][k] = t; s_cnt[warp][k] = tc; }
    }
    __syncthreads();
    if (threadIdx.x < RI_SUMS) {
        dd t = s_sum[0][threadIdx.x];
        double tc = s_cnt[0][threadIdx.x];
        for (int wgt = 1; wgt < RI_THREADS / 32; ++wgt) { dd_merge(t, s_sum[wgt][threadIdx.x]); tc += s_cnt[wgt][threadIdx.x]; }
        double *out = partials + (size_t)blockIdx.x * RI_VALS;
        out[threadIdx.x] = t.hi;
        out[RI_SUMS + threadIdx.x] = t.lo;
        out[2 * RI_SUMS + threadIdx.x] = tc;
    }
}

__global__ void reduce_identity_stage2(const double *partials, int32_t blocks, const uint8_t *labels, int32_t n,
                                       int64_t L, double seg, const double2 *harm, int32_t harm_n, double *stats,
                                       int64_t *counts, double *wsum) {
    __shared__ double s_tot[RI_SUMS], s_pairs[RI_CNTS];
    __shared__ int s_n[3];
    if (threadIdx.x < 3) s_n[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < RI_SUMS) {
        dd t = {0.0, 0.0};
        double tc = 0.0;
        for (int b = 0; b < blocks; ++b) {
            const double *in = partials + (size_t)b * RI_VALS;
            dd p = {in[threadIdx.x], in[RI_SUMS + threadIdx.x]};
            dd_merge(t, p);
            tc += in[2 * RI_SUMS + threadIdx.x];
        }
        s_tot[threadIdx.x] = dd_value(t);
        s_pairs[threadIdx.x] = tc;
    }
    int c0 = 0, c1 = 0, c2 = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t f = labels ? labels[i] : 0u;
        c0 += (f & IMPOP_LAB_SUBSET) != 0; c1 += (f & IMPOP_LAB_A) != 0; c2 += (f & IMPOP_LAB_B) != 0;
    }
    if (c0) atomicAdd(&s_n[0], c0);
    if (c1) atomicAdd(&s_n[1], c1);
    if (c2) atomicAdd(&s_n[2], c2);
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t cnt[IMPOP_NCOUNTS];
        cnt[0] = s_n[0]; cnt[1] = s_n[1]; cnt[2] = s_n[2];
        cnt[3] = (int64_t)s_pairs[0]; cnt[4] = (int64_t)s_pairs[1]; cnt[5] = (int64_t)s_pairs[2]; cnt[6] = (int64_t)s_pairs[3];
        cnt[7] = (int64_t)seg;
        double sums[4] = {s_tot[0], s_tot[1], s_tot[2], s_tot[3]};
        double st[IMPOP_NSTATS];
        finalize_row(sums, cnt, L, seg, harm, harm_n, st);
        if (stats) for (int k = 0; k < IMPOP_NSTATS; ++k) stats[k] = st[k];
        if (counts) for (int k = 0; k < IMPOP_NCOUNTS; ++k) counts[k] = cnt[k];
        if (wsum) {   // pica2.py:137-164 with group frequencies as weights; n = every element of the table
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            double pw = 0.0;
            if (s_pairs[4] > 0.0 && n >= 2) {
                double dn = (double)n;
                pw = __dmul_rn(__ddiv_rn(dn, __dadd_rn(dn, -1.0)), __dmul_rn(2.0, s_tot[4]));   // pica2.py:154
            }
            wsum[0] = s_tot[4]; wsum[1] = s_pairs[4]; wsum[2] = pw;
            wsum[3] = (L > 0) ? __ddiv_rn(pw, (double)L) : nan;                                 // pica2.py:163-164
        }
    }
}

// ------------------------------------------------------------------------------------------
// tj_d.py:47-69 for independent triples (the tj_d.py CLI path).
// ------------------------------------------------------------------------------------------
__global__ void tajima_kernel(const int64_t *n, const double *S, const double *pi, int32_t count, double *D,
                              double *parts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int64_t nn = n[t];
    double t1 = 0.0, c1 = 0.0, t2 = 0.0, c2 = 0.0;
    for (int64_t i = 1; i < nn; ++i) {
        double di = (double)i;
        neumaier_add(t1, c1, __ddiv_rn(1.0, di));
        neumaier_add(t2, c2, __ddiv_rn(1.0, __dmul_rn(di, di)));
    }
    double a1 = neumaier_value(t1, c1), a2 = neumaier_value(t2, c2);
    D[t] = tajima_from_harmonics((double)nn, S[t], pi[t], a1, a2, parts ? parts + (size_t)t * 10 : nullptr);
}

// ------------------------------------------------------------------------------------------
// K4: per-site allele counts.  Thread per site; masks staged in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int SITE_THREADS = 256;
constexpr int SITE_MAX_MASK_WORDS = 2048;  // pops * words (u64) held in shared memory

__global__ void __launch_bounds__(SITE_THREADS) site_counts_kernel(const uint64_t *sites, int64_t M, int32_t words,
                                                                   const uint64_t *masks, int32_t P, int32_t *counts,
                                                                   double *freq) {
    __shared__ uint64_t s_mask[SITE_MAX_MASK_WORDS];
    __shared__ double s_size[64];
    for (int k = threadIdx.x; k < P * words; k += SITE_THREADS) s_mask[k] = masks[k];
    __syncthreads();
    if (threadIdx.x < P) {
        int c = 0;
        for (int w = 0; w < words; ++w) c += __popcll(s_mask[threadIdx.x * words + w]);
        s_size[threadIdx.x] = (double)c;
    }
    __syncthreads();
    for (int64_t s = (int64_t)blockIdx.x * SITE_THREADS + threadIdx.x; s < M; s += (int64_t)gridDim.x * SITE_THREADS) {
        const uint64_t *row = sites + (size_t)s * words;
        for (int p = 0; p < P; ++p) {
            int c = 0;
            for (int w = 0; w < words; ++w) c += __popcll(__ldg(row + w) & s_mask[p * words + w]);
            counts[(size_t)s * P + p] = c;
            if (freq) freq[(size_t)s * P + p] = s_size[p] > 0.0 ? __ddiv_rn((double)c, s_size[p]) : 0.0;
        }
    }
}

// Fast path for words == 8 (up to 512 haplotypes, the HPRC panel): the 64-byte site row is read
// once as four 16-byte loads and every population is counted from registers.  Results are transposed
// through shared memory so that a warp's 32 x P counts (and frequencies) leave as fully coalesced stores.
// POPC is a quarter-rate instruction (16 lanes / clk / SM): counting every 32-bit word of the row against every panel
// (16 x P) makes the XU pipe, not HBM, the limit.  Panels are sparse in the haplotype index (a superpopulation is a
// few contiguous runs), so 32-bit mask words that are zero are skipped with a block-uniform test: 19 instead of 80
// POPC per site for the five HPRC superpopulations, and the kernel is back on the HBM roofline.
template <int P>
__global__ void __launch_bounds__(SITE_THREADS) site_counts_w8_kernel(const uint64_t *sites, int64_t M,
                                                                      const uint64_t *masks, int32_t *counts,
                                                                      double *freq) {
    __shared__ uint64_t s_mask[P * 8];
    __shared__ double s_size[P];
    __shared__ uint32_t s_nz[P];               // bit w: 32-bit word w of the panel's mask is not zero
    __shared__ int32_t s_cnt[SITE_THREADS / 32][32 * P];
    __shared__ double s_frq[SITE_THREADS / 32][32 * P];
    for (int k = threadIdx.x; k < P * 8; k += SITE_THREADS) s_mask[k] = masks[k];
    __syncthreads();
    const uint32_t *s_mask32 = reinterpret_cast<const uint32_t *>(s_mask);
    if (threadIdx.x < P) {
        int c = 0;
        uint32_t nz = 0u;
        for (int w = 0; w < 16; ++w) {
            const uint32_t mw = s_mask32[threadIdx.x * 16 + w];
            c += __popc(mw);
            nz |= (mw != 0u ? 1u : 0u) << w;
        }
        s_size[threadIdx.x] = (double)c;
        s_nz[threadIdx.x] = nz;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (SITE_THREADS / 32);
    for (int64_t s0 = ((int64_t)blockIdx.x * (SITE_THREADS / 32) + warp) * 32; s0 < M; s0 += warps_total * 32) {
        const int64_t s = s0 + lane;
        uint32_t v[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) v[w] = 0u;
        if (s < M) {
            const uint4 *row = reinterpret_cast<const uint4 *>(sites + (size_t)s * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 r = __ldg(row + q);
                v[4 * q] = r.x; v[4 * q + 1] = r.y; v[4 * q + 2] = r.z; v[4 * q + 3] = r.w;
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint32_t nz = s_nz[p];
            int c = 0;
#pragma unroll
            for (int w = 0; w < 16; ++w)
                if ((nz >> w) & 1u) c += __popc(v[w] & s_mask32[p * 16 + w]);
            s_cnt[warp][lane * P + p] = c;                       // stride P words: conflict-free for odd P
            // counts and panel sizes are integers below 2^31: the short correctly rounded division (common.cuh)
            if (freq) s_frq[warp][lane * P + p] = s_size[p] > 0.0 ? div_rn_int31((double)c, s_size[p]) : 0.0;
        }
        __syncwarp();
        const int64_t valid = (M - s0 < 32 ? M - s0 : 32) * P;  // entries of this warp's 32 sites
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int k = q * 32 + lane;
            if (k < valid) {
                counts[(size_t)s0 * P + k] = s_cnt[warp][k];
                if (freq) freq[(size_t)s0 * P + k] = s_frq[warp][k];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K5: connected components of {identity >= threshold} (af.py:35-44).  Lock-free union-find that
// always links the larger root under the smaller, so every root is its component's smallest index.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(volatile int32_t *parent, int v) {
    int p = parent[v];
    while (p != v) { v = p; p = parent[v]; }
    return v;
}

__global__ void cluster_init_kernel(int32_t *parent, int32_t n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

__global__ void cluster_link_kernel(const double *ident, int32_t n, int64_t ld, double threshold, int32_t *parent) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
            const double s = ident[(size_t)i * ld + j];
            if (!(s >= threshold)) continue;  // NaN (absent pair) never links
            int a = i, b = j;
            while (true) {
                a = uf_find(parent, a);
                b = uf_find(parent, b);
                if (a == b) break;
                const int hi = a > b ? a : b, lo = a > b ? b : a;
                if (atomicCAS(&parent[hi], hi, lo) == hi) break;
            }
        }
    }
}

__global__ void cluster_flatten_kernel(int32_t *parent, int32_t n, int32_t *comp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) comp[i] = uf_find(parent, i);
}

// ------------------------------------------------------------------------------------------
// pica2.py:94-112 star grouping, made deterministic: the seed is always the smallest remaining
// index (names are sorted, so index order is name order).  The seed loop is
// sequential by definition; each step scans the seed's row in parallel.  One CTA.
// group[i] = seed index of i's group (the seed is the group's smallest member = its
// representative, pica2.py:110/128); weight[i] = |G|/N on seeds, 0 elsewhere (pica2.py:137-138).
// ------------------------------------------------------------------------------------------
constexpr int GROUP_THREADS = 1024;

__global__ void __launch_bounds__(GROUP_THREADS) greedy_groups_kernel(const double *ident, int32_t n, int64_t ld,
                                                                      double threshold, int32_t *group,
                                                                      double *weight) {
    for (int i = threadIdx.x; i < n; i += GROUP_THREADS) {
        group[i] = -1;
        if (weight) weight[i] = 0.0;
    }
    __syncthreads();
    for (int s = 0; s < n; ++s) {
        if (group[s] != -1) continue;   // uniform: written before the last barrier
        for (int j = s + 1 + threadIdx.x; j < n; j += GROUP_THREADS) {
            if (group[j] != -1) continue;
            const double v = ident[(size_t)s * ld + j];
            if (v > threshold) group[j] = s;   // strict; NaN (absent pair) never joins (pica2.py:106)
        }
        if (threadIdx.x == 0) group[s] = s;
        __syncthreads();
    }
    if (!weight) return;
    for (int i = threadIdx.x; i < n; i += GROUP_THREADS) atomicAdd(&weight[group[i]], 1.0);   // exact small integers
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += GROUP_THREADS) weight[i] = __ddiv_rn(weight[i], (double)n);
}

// ------------------------------------------------------------------------------------------
// round(x, r) as CPython rounds a float (pica2.py:81-83, h-fst.py:149-150): the decimal value of the double, rounded
// half-even at the r-th decimal, converted back to the nearest double.  rint(x * 10^r) / 10^r is NOT that: x * 10^r is
// rounded once before the tie test (one value in ~5 10^5 lands on the wrong side).  Here the product is
// exact as a double-double (p, e) = two_prod(|x|, 10^r) with 10^r exact for r <= 22; t = floor(p); the fraction's
// distance from one half, d = (p - t) - 0.5, is exact, and the sign of d + e is the sign of the exact sum (a
// floating-point sum of two doubles is zero only when the exact sum is).  k = t or t + 1 (ties to even), and the result is
// the correctly rounded quotient k / 10^r -- the nearest double to the decimal string CPython's dtoa round trip forms.
// From 2^52 on p is an integer and |e| <= 1/2: the one extra case is e = -1/2 exactly (tie between p - 1 and p).  Values
// with |x| 10^r >= 2^53 cannot move (the rounded decimal is within ulp(x) / 4 of x) and come back unchanged, like NaN and
// infinities.
// ------------------------------------------------------------------------------------------
__global__ void round_decimal_kernel(double *v, int64_t count, double pow10) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[t];
        const double ax = fabs(x);
        const double p = __dmul_rn(ax, pow10);
        if (!(p < 9007199254740992.0)) continue;             // NaN, infinities, nothing to round
        const double e = __fma_rn(ax, pow10, -p);            // exact error of the product
        const double fl = floor(p);
        const bool odd = fmod(fl, 2.0) != 0.0;
        double k = fl;
        if (p == fl && e == -0.5) {                          // exact value p - 1/2: tie between p - 1 and p
            if (odd) k = __dadd_rn(fl, -1.0);
        } else {
            const double d = __dadd_rn(__dadd_rn(p, -fl), -0.5);
            const double s = __dadd_rn(d, e);
            if (s > 0.0 || (s == 0.0 && odd)) k = __dadd_rn(fl, 1.0);
        }
        const double r = __ddiv_rn(k, pow10);
        v[t] = copysign(r, x);
    }
}

// ------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_round_decimal(double *v, int64_t count, int32_t digits, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    double p = 1.0;
    for (int k = 0; k < digits; ++k) p *= 10.0;               // exact up to 10^22
    int64_t blocks = (count + 255) / 256;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    round_decimal_kernel<<<(int)blocks, 256, 0, st>>>(v, count, p);
    return cudaGetLastError();
}

// Tight rows (as stored / transferred) -> the 16-byte-multiple rows the window kernels read, zero padded.
__global__ void repitch_rows_kernel(const RepitchDesc *desc, int32_t windows, const uint32_t *src, uint32_t *dst) {
    for (int w = blockIdx.x; w < windows; w += gridDim.x) {
        const RepitchDesc d = desc[w];
        const int64_t total = (int64_t)d.rows * d.dst_pitch;
        for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
            const int r = (int)(t / d.dst_pitch), c = (int)(t - (int64_t)r * d.dst_pitch);
            dst[d.dst_off + t] = c < d.src_pitch ? __ldg(src + d.src_off + (int64_t)r * d.src_pitch + c) : 0u;
        }
    }
}

cudaError_t launch_repitch_rows(const RepitchDesc *desc, int32_t windows, const uint32_t *src, uint32_t *dst, int sm_count,
                                cudaStream_t st) {
    if (windows == 0) return cudaSuccess;
    repitch_rows_kernel<<<min(windows, sm_count * 8), 256, 0, st>>>(desc, windows, src, dst);
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(const uint8_t *dense, int32_t n, int32_t m, int64_t dpitch, uint32_t *x,
                             int32_t pitch_words, int sm_count, cudaStream_t st) {
    int64_t total = (int64_t)n * pitch_words;
    if (total == 0) return cudaSuccess;
    int64_t blocks = (total + 7) / 8;
    if (blocks > (int64_t)sm_count * 32) blocks = (int64_t)sm_count * 32;
    pack_bits_kernel<<<(int)blocks, 256, 0, st>>>(dense, n, m, dpitch, x, pitch_words);
    return cudaGetLastError();
}

int reduce_identity_blocks(int32_t n, int sm_count) { return n < 1 ? 1 : (n < sm_count * 4 ? n : sm_count * 4); }

cudaError_t launch_reduce_identity(const double *ident, int32_t n, int64_t ld, const uint8_t *labels,
                                   const double *weight, int64_t L, double seg, const double2 *harm, int32_t harm_n,
                                   double *scratch, int sm_count, double *stats, int64_t *counts, double *wsum, cudaStream_t st) {
    int blocks = reduce_identity_blocks(n, sm_count);
    reduce_identity_stage1<<<blocks, RI_THREADS, 0, st>>>(ident, n, ld, labels, weight, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reduce_identity_stage2<<<1, 128, 0, st>>>(scratch, blocks, labels, n, L, seg, harm, harm_n, stats, counts, wsum);
    return cudaGetLastError();
}

cudaError_t launch_tajima(const int64_t *n, const double *S, const double *pi, int32_t count, double *D, double *parts,
                          cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    tajima_kernel<<<(count + 63) / 64, 64, 0, st>>>(n, S, pi, count, D, parts);
    return cudaGetLastError();
}

cudaError_t launch_site_counts(const uint64_t *sites, int64_t M, int32_t words, const uint64_t *masks, int32_t P,
                               int32_t *counts, double *freq, int sm_count, cudaStream_t st) {
    if (M == 0) return cudaSuccess;
    int64_t blocks = (M + SITE_THREADS - 1) / SITE_THREADS;
    int64_t cap = (int64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (words == 8 && P == 5) site_counts_w8_kernel<5><<<(int)blocks, SITE_THREADS, 0, st>>>(sites, M, masks, counts, freq);
    else if (words == 8 && P == 1) site_counts_w8_kernel<1><<<(int)blocks, SITE_THREADS, 0, st>>>(sites, M, masks, counts, freq);
    else if (words == 8 && P == 2) site_counts_w8_kernel<2><<<(int)blocks, SITE_THREADS, 0, st>>>(sites, M, masks, counts, freq);
    else site_counts_kernel<<<(int)blocks, SITE_THREADS, 0, st>>>(sites, M, words, masks, P, counts, freq);
    return cudaGetLastError();
}

cudaError_t launch_cluster(const double *ident, int32_t n, int64_t ld, double threshold, int32_t *parent, int32_t *comp,
                           int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    cluster_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(parent, n);
    cluster_link_kernel<<<n < sm_count * 8 ? n : sm_count * 8, 128, 0, st>>>(ident, n, ld, threshold, parent);
    cluster_flatten_kernel<<<(n + 255) / 256, 256, 0, st>>>(parent, n, comp);
    return cudaGetLastError();
}

cudaError_t launch_greedy_groups(const double *ident, int32_t n, int64_t ld, double threshold, int32_t *group,
                                 double *weight, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    greedy_groups_kernel<<<1, GROUP_THREADS, 0, st>>>(ident, n, ld, threshold, group, weight);
    return cudaGetLastError();
}

}  // namespace impop
