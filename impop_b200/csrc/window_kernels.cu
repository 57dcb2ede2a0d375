// Windowed pairwise kernels: prep (path lengths, byte weights, heavy-node table and bits), the fused
// pairwise + fp64 reduction kernel in two implementations (warp-specialised tcgen05 / TMEM, and dp4a
// SIMT as cross-check), segregating-node counts, and the per-window finalize.
//
// Replaces, per window: `impg similarity` / `odgi similarity` (reference call sites
// run_pica2_impg.sh:162-168, run_h-fst.sh:65-67, run_tajd.sh:160) + pica2.py:118-164 +
// h-fst.py:130-249 + run_tajd.sh:126-148 (S) + tj_d.py:47-69.
//
// Arithmetic: I_ij = sum_k len_k x_ik x_jk is computed as an unsigned 8-bit GEMM with exact
// int32 accumulation.  Each node k contributes the "dense" column (a = x_ik * (len_k % 255), b = x_jk)
// and, when len_k >= 255, ceil(q / 255) "heavy" columns (a = x_ik * c, b = 255 * x_jk) with the c's
// summing to q = len_k / 255, so a * b summed over a node's columns is exactly len_k.  The weights sit on
// the A operand because the A tile has at most half the rows of the B tile (fewer weighted expansions).
#include <stdio.h>

#include <type_traits>

#include "common.cuh"
#include "stats_math.cuh"

namespace impop {

// ==========================================================================================
// Prep, three small kernels (the second is the only one that touches the presence matrix):
//   prep_cols   one CTA per window: byte weights of the dense columns and their eight bit planes (ballot over 32 nodes),
//               heavy-node table, range check sum(len) < 2^31, label counts, reset of the any / all words
//   prep_rows   one CTA per (window, row slice) -- a window with many haplotypes is cut into slices so that a
//               batch of few large windows still fills the GPU: path lengths A_i as the sum over the non-zero bit
//               planes of the dense byte weights of 2^p popc(row word & plane word) (rows staged through a shared
//               tile: coalesced loads with lane = word, arithmetic with lane = row and the planes of the current
//               word broadcast) plus 255 c per present heavy column, presence bits of the heavy columns, any / all
//               masks over the SEG rows
//   seg_count   S = #{k : 0 < sum_{i in SEG} x_ik < |SEG|, len_k > 0} (replaces `povu gfa2vcf | wc -l`,
//               run_tajd.sh:126-148) -> counts row nS nA nB pS pAA pBB pAB S
// ==========================================================================================
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS) prep_cols_kernel(const __grid_constant__ WindowTab tab, int64_t *counts) {
    __shared__ int s_heavy, s_cnt[4];
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31;
    for (int w = blockIdx.x; w < tab.W; w += gridDim.x) {
        const int n = tab.n[w], m = tab.m[w];
        const uint32_t *len = tab.len + tab.len_off[w];
        const uint8_t *lab = tab.labels + tab.lab_off[w];
        uint8_t *w8 = tab.w8 + tab.w8_off[w], *w8n = tab.w8n + tab.w8_off[w];
        uint32_t *heavy = tab.heavy + tab.heavy_off[w];
        const int m64 = ((m + KCHUNK - 1) / KCHUNK) * KCHUNK;
        const int hpad = (int)(tab.heavy_off[w + 1] - tab.heavy_off[w]);
        if (threadIdx.x == 0) { s_heavy = 0; s_total = 0ull; }
        if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        {
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            for (int i = threadIdx.x; i < n; i += PREP_THREADS) {
                const uint32_t f = clean_label(lab[i]);
                c0 += (f & IMPOP_LAB_SUBSET) != 0; c1 += (f & IMPOP_LAB_A) != 0; c2 += (f & IMPOP_LAB_B) != 0;
                c3 += (f & IMPOP_LAB_SEG) != 0;
            }
            if (c0) atomicAdd(&s_cnt[0], c0);
            if (c1) atomicAdd(&s_cnt[1], c1);
            if (c2) atomicAdd(&s_cnt[2], c2);
            if (c3) atomicAdd(&s_cnt[3], c3);
        }
        unsigned long long tot = 0;
        for (int k = threadIdx.x; k < m64; k += PREP_THREADS) {
            uint32_t l = (k < m) ? __ldg(len + k) : 0u;
            tot += l;
            const uint32_t bwt = l % HEAVY_Q;
            w8[kperm(k)] = w8n[k] = (uint8_t)bwt;
            {   // bit planes of the byte weights for prep_rows (a warp covers 32 consecutive nodes; m64 is a multiple of 128)
                uint32_t *pl = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(tab.planes) + tab.w8_off[w]) + (k >> 5) * 8;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const uint32_t msk = __ballot_sync(0xffffffffu, (bwt >> p) & 1u);
                    if (lane == p) pl[p] = msk;
                }
                const uint32_t lv = __ballot_sync(0xffffffffu, l > 0u);        // nodes of positive length (l = 0 beyond m)
                if (lane == 0 && (k >> 5) < ((m + 31) >> 5)) tab.live[tab.word_off[w] + (k >> 5)] = lv;
            }
            uint32_t q = l / HEAVY_Q;
            while (q > 0) {
                uint32_t c = q < 255u ? q : 255u;
                int slot = atomicAdd(&s_heavy, 1);
                if (slot < hpad) heavy[slot] = ((uint32_t)k << 8) | c;
                q -= c;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
        if (lane == 0 && tot) atomicAdd(&s_total, tot);
        const int64_t wo = tab.word_off[w], words = tab.word_off[w + 1] - wo;
        for (int64_t k = threadIdx.x; k < words; k += PREP_THREADS) { tab.seg_any[wo + k] = 0u; tab.seg_all[wo + k] = 0xffffffffu; }
        {   // heavy presence bits start out empty (prep_rows sets the words that hold entries)
            const int64_t h0 = tab.xh_off[w], h1 = tab.xh_off[w + 1];
            for (int64_t k = h0 + threadIdx.x; k < h1; k += PREP_THREADS) tab.xh[k] = 0u;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t cw = tab.win_const ? tab.win_const[w] : 0;                     // affine form: window constant C
            if (s_total >= (1ull << 31) || s_heavy > hpad || cw < 0 || cw >= (1ll << 31)) atomicExch(tab.err, (int32_t)DEV_ERR_RANGE);
        }
        const int nh = s_heavy < hpad ? s_heavy : hpad;
        for (int s = nh + threadIdx.x; s < hpad; s += PREP_THREADS) heavy[s] = 0u;
        if (threadIdx.x == 0) tab.heavy_n[w] = nh;
        __syncthreads();
        for (int s = threadIdx.x; s < hpad; s += PREP_THREADS) w8[m64 + kperm(s)] = w8n[m64 + s] = (uint8_t)(heavy[s] & 255u);
        if (threadIdx.x == 0 && counts) {
            const int64_t nS = s_cnt[0], nA = s_cnt[1], nB = s_cnt[2];
            int64_t *row = counts + (size_t)w * IMPOP_NCOUNTS;
            row[0] = nS; row[1] = nA; row[2] = nB;
            row[3] = nS * (nS - 1) / 2; row[4] = nA * (nA - 1) / 2; row[5] = nB * (nB - 1) / 2; row[6] = nA * nB;
            row[7] = s_cnt[3];                  // number of SEG rows for now; seg_count_kernel replaces it by S
        }
        __syncthreads();
    }
}

// prep_rows.  A warp takes 32 rows of a window at a time, 32 presence words (1024 nodes) per pass:
//   load     lane = word: the 32 rows are read coalesced (128 bytes per row) into a padded tile in shared memory; the any /
//            all words over the SEG rows are folded in registers on the way
//   compute  lane = row: the row's words come back from the tile (conflict-free), the bit planes of the byte weights of
//            the current word are the same for every lane (shared memory, broadcast), so planes that are zero for all 32
//            nodes of a word are skipped with warp-uniform branches: a word whose nodes all weigh 1 (SNP-dominated windows
//            once the constant columns are merged and the columns ordered by weight, impop_compact_scan / _fill) costs
//            ONE popc per 32 rows, a padding word none.  Heavy columns: one tile probe per entry and 32 rows.
// PR_WORDS presence words (2048 nodes) of planes are staged per group.
constexpr int PR_WORDS = 64;
constexpr int PR_TILE = 33;                   // tile row pitch in words (32 + 1: lane = row reads hit 32 banks)
constexpr int PR_FAST_HEAVY = 8;              // register path: at most this many heavy entries per window (probed one by one)

#ifndef IMPOP_PREP_OCC
#define IMPOP_PREP_OCC 4
#endif
__global__ void __launch_bounds__(PREP_THREADS, IMPOP_PREP_OCC) prep_rows_kernel(const __grid_constant__ WindowTab tab) {
    __shared__ __align__(16) uint32_t s_pl[PR_WORDS][8];
    __shared__ uint32_t s_mk[PR_WORDS], s_any[PR_WORDS], s_all[PR_WORDS];
    __shared__ uint32_t s_tile[PREP_THREADS / 32][32 * PR_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *tile = s_tile[warp];
    for (int sidx = blockIdx.x; sidx < tab.n_slices; sidx += gridDim.x) {
        const int4 sl = __ldg(tab.slices + sidx);
        const int w = sl.x, row_lo = sl.y, row_hi = sl.z;
        const int m = tab.m[w], pitch = tab.pitch[w];
        const int wlim = min(pitch, ((m + 127) >> 7) << 2);      // words of a row that can hold nodes < m (16-byte groups):
                                                                 // a window may be a column range of a wider matrix
        const uint32_t *x = tab.x + tab.x_off[w];
        const uint8_t *lab = tab.labels + tab.lab_off[w];
        const uint32_t *heavy = tab.heavy + tab.heavy_off[w];
        const int hwords = (int)((tab.heavy_off[w + 1] - tab.heavy_off[w]) >> 5);
        const int heavy_n = tab.heavy_n[w];
        const int hw_used = (heavy_n + 31) >> 5;             // words of the heavy table that hold real entries
        uint32_t *xh = tab.xh + tab.xh_off[w];
        int32_t *A = tab.A + tab.row_off[w];
        const int64_t wo = tab.word_off[w];
        const uint32_t *planes = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(tab.planes) + tab.w8_off[w]);
        const int plane_words = (((m + KCHUNK - 1) / KCHUNK) * KCHUNK) >> 5;   // prep_cols wrote planes for the padded node count
        for (int w0 = 0; w0 < wlim || w0 == 0; w0 += PR_WORDS) {
            const int gw = max(0, min(PR_WORDS, wlim - w0));       // words of this group (a multiple of 4)
            for (int t = threadIdx.x; t < gw * 8; t += PREP_THREADS)
                s_pl[t >> 3][t & 7] = (w0 + (t >> 3) < plane_words) ? __ldg(planes + (size_t)(w0 + (t >> 3)) * 8 + (t & 7)) : 0u;
            if (threadIdx.x < PR_WORDS) { s_any[threadIdx.x] = 0u; s_all[threadIdx.x] = 0xffffffffu; }
            __syncthreads();
            if (threadIdx.x < gw) {
                uint32_t mk = 0u;
#pragma unroll
                for (int p = 0; p < 8; ++p) mk |= (s_pl[threadIdx.x][p] != 0u ? 1u : 0u) << p;
                s_mk[threadIdx.x] = mk;
            }
            __syncthreads();
            uint32_t any[PR_WORDS / 32], all[PR_WORDS / 32];        // lane = word of pass ps
#pragma unroll
            for (int ps = 0; ps < PR_WORDS / 32; ++ps) { any[ps] = 0u; all[ps] = 0xffffffffu; }
            // Register path (every window compacted at ingest takes it: rows back to back, <= 1 024 columns, a handful of heavy
            // entries at most): lane = row, the row's words arrive 16 bytes at a time straight into registers (next group
            // requested before the current one is processed) -- no shared tile, no transposition.  The bit planes and the
            // plane mask of a word are the same for all lanes (shared memory, broadcast, warp-uniform branches); the any / all
            // words over the SEG rows take one warp reduction each per word and 32 rows and are kept by lane = word.
            const bool fast = (pitch == wlim) && (pitch <= 32) && (heavy_n <= PR_FAST_HEAVY);
            if (fast) {
                const int G = pitch >> 2;                           // 16-byte groups per row
                for (int i0 = row_lo + warp * 32; i0 < row_hi; i0 += PREP_THREADS) {
                    const int i = i0 + lane;
                    const bool valid = i < row_hi;
                    const bool seg = valid && (lab[valid ? i : row_lo] & IMPOP_LAB_SEG);
                    const uint4 *rowp = reinterpret_cast<const uint4 *>(x + (size_t)(valid ? i : row_lo) * pitch);
                    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
                    uint32_t acc = 0u;
                    uint4 nxt = valid ? __ldg(rowp) : zero4;
#pragma unroll 1
                    for (int g = 0; g < G; ++g) {
                        const uint4 cur = nxt;
                        if (g + 1 < G) nxt = valid ? __ldg(rowp + g + 1) : zero4;
                        const uint32_t wv[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int wd = 4 * g + t;
                            const uint32_t v = wv[t];
                            const uint32_t ra = __reduce_or_sync(0xffffffffu, seg ? v : 0u);
                            const uint32_t rl = __reduce_and_sync(0xffffffffu, seg ? v : 0xffffffffu);
                            if (lane == wd) { any[0] |= ra; all[0] &= rl; }
                            const uint32_t mk = s_mk[wd];
                            if (!mk) continue;
                            if ((mk & (mk - 1u)) == 0u) {                              // one plane: every node of the word weighs 2^p (or 0)
                                const int pl = __ffs(mk) - 1;
                                acc += (uint32_t)__popc(v & s_pl[wd][pl]) << pl;
                            } else {
                                const uint4 lo4 = *reinterpret_cast<const uint4 *>(&s_pl[wd][0]);
                                acc += (uint32_t)__popc(v & lo4.x) + ((uint32_t)__popc(v & lo4.y) << 1) +
                                       ((uint32_t)__popc(v & lo4.z) << 2) + ((uint32_t)__popc(v & lo4.w) << 3);
                                if (mk >> 4) {
                                    const uint4 hi4 = *reinterpret_cast<const uint4 *>(&s_pl[wd][4]);
                                    acc += ((uint32_t)__popc(v & hi4.x) << 4) + ((uint32_t)__popc(v & hi4.y) << 5) +
                                           ((uint32_t)__popc(v & hi4.z) << 6) + ((uint32_t)__popc(v & hi4.w) << 7);
                                }
                            }
                        }
                    }
                    uint32_t hbits = 0u;                            // the few heavy entries: their words come back from L1
                    for (int e = 0; e < heavy_n; ++e) {
                        const uint32_t ent = __ldg(heavy + e);
                        if (!(ent & 255u)) continue;
                        const uint32_t wvv = valid ? __ldg(x + (size_t)i * pitch + (ent >> 13)) : 0u;
                        const uint32_t on = (wvv >> ((ent >> 8) & 31u)) & 1u;
                        acc += on * (HEAVY_Q * (ent & 255u));
                        hbits |= on << e;
                    }
                    if (valid) {
                        if (hbits) xh[(size_t)i * hwords] |= hbits;                   // (xh was zeroed by prep_cols; heavy_n <= 32: one word)
                        const uint32_t cw = tab.win_const ? (uint32_t)tab.win_const[w] : 0u;
                        const uint32_t ri = tab.row_adj ? (uint32_t)tab.row_adj[tab.row_off[w] + i] : 0u;
                        const uint32_t a = acc + cw - 2u * ri;                        // affine form: A_i = sum_k len_k x_ik + C - 2 R_i
                        A[i] = (int32_t)a;
                        if ((int32_t)a < 0) atomicExch(tab.err, (int32_t)DEV_ERR_RANGE);
                    }
                }
            } else
            for (int i0 = row_lo + warp * 32; i0 < row_hi; i0 += PREP_THREADS) {
                const int i = i0 + lane;
                const bool valid = i < row_hi;
                const uint32_t segrows = __ballot_sync(0xffffffffu, valid && (lab[valid ? i : row_lo] & IMPOP_LAB_SEG));
                const int nrows = min(32, row_hi - i0);
                uint32_t acc = 0u;
#pragma unroll
                for (int ps = 0; ps < PR_WORDS / 32; ++ps) {
                    const int pw = min(32, gw - ps * 32);           // words of this pass
                    if (pw <= 0) break;
                    {   // load: lane = word
                        const bool lane_ok = lane < pw;
                        const uint32_t *src = x + (size_t)i0 * pitch + w0 + ps * 32 + lane;
                        __syncwarp();
#pragma unroll 1
                        for (int rb = 0; rb < 32; rb += 8) {       // eight row loads in flight, then their stores (kept explicit:
                            uint32_t v[8];                         // left to the scheduler, a load and its store end up back to back)
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = (lane_ok && rb + q < nrows) ? __ldg(src + (size_t)(rb + q) * pitch) : 0u;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                tile[(rb + q) * PR_TILE + lane] = v[q];
                                if ((segrows >> (rb + q)) & 1u) { any[ps] |= v[q]; all[ps] &= v[q]; }
                            }
                        }
                        __syncwarp();
                    }
                    const uint32_t *mine = tile + lane * PR_TILE;   // compute: lane = row
                    for (int wd = 0; wd < pw; ++wd) {
                        const uint32_t mk = s_mk[ps * 32 + wd];
                        if (!mk) continue;
                        const uint32_t v = mine[wd];
                        if (mk == 1u) {
                            acc += (uint32_t)__popc(v & s_pl[ps * 32 + wd][0]);
                        } else if ((mk & (mk - 1u)) == 0u) {                       // one plane: every node of the word weighs 2^p (or 0)
                            const int p = __ffs(mk) - 1;
                            acc += (uint32_t)__popc(v & s_pl[ps * 32 + wd][p]) << p;
                        } else {
                            const uint4 lo4 = *reinterpret_cast<const uint4 *>(&s_pl[ps * 32 + wd][0]);
                            acc += (uint32_t)__popc(v & lo4.x) + ((uint32_t)__popc(v & lo4.y) << 1) +
                                   ((uint32_t)__popc(v & lo4.z) << 2) + ((uint32_t)__popc(v & lo4.w) << 3);
                            if (mk >> 4) {
                                const uint4 hi4 = *reinterpret_cast<const uint4 *>(&s_pl[ps * 32 + wd][4]);
                                acc += ((uint32_t)__popc(v & hi4.x) << 4) + ((uint32_t)__popc(v & hi4.y) << 5) +
                                       ((uint32_t)__popc(v & hi4.z) << 6) + ((uint32_t)__popc(v & hi4.w) << 7);
                            }
                        }
                    }
                    // heavy columns (nodes of >= 255 bp) whose node lies in this pass: entry e = (node << 8 | c) adds 255 c to
                    // the rows that carry the node and becomes bit e of the row's heavy presence words (operand of the heavy
                    // chunks of the pairs kernel)
                    const int pass_w0 = w0 + ps * 32;
                    for (int hw = 0; hw < hw_used; ++hw) {
                        uint32_t hbits = 0u;
                        bool touched = false;
                        const int e_end = min(32, heavy_n - hw * 32);                    // entries in use (the rest is padding)
                        for (int e = 0; e < e_end; ++e) {
                            const uint32_t ent = __ldg(heavy + hw * 32 + e);
                            const int rel = (int)(ent >> 13) - pass_w0;                  // word of the entry's node within this pass
                            if (!(ent & 255u) || rel < 0 || rel >= pw) continue;        // other pass (uniform)
                            const uint32_t on = (mine[rel] >> ((ent >> 8) & 31u)) & 1u;
                            acc += on * (HEAVY_Q * (ent & 255u));
                            hbits |= on << e;
                            touched = true;
                        }
                        if (valid && touched && hbits) xh[(size_t)i * hwords + hw] |= hbits;   // (xh was zeroed by prep_cols)
                    }
                }
                if (valid) {
                    // affine form: A_i = I_ii = sum_k len_k x_ik + C - 2 R_i
                    const uint32_t cw = tab.win_const ? (uint32_t)tab.win_const[w] : 0u;
                    const uint32_t ri = tab.row_adj ? (uint32_t)tab.row_adj[tab.row_off[w] + i] : 0u;
                    const uint32_t a = acc + (w0 ? (uint32_t)A[i] : cw - 2u * ri);
                    A[i] = (int32_t)a;
                    if (w0 + PR_WORDS >= wlim && (int32_t)a < 0) atomicExch(tab.err, (int32_t)DEV_ERR_RANGE);   // path length out of range
                }
            }
#pragma unroll
            for (int ps = 0; ps < PR_WORDS / 32; ++ps) {
                if (any[ps]) atomicOr(&s_any[ps * 32 + lane], any[ps]);
                if (all[ps] != 0xffffffffu) atomicAnd(&s_all[ps * 32 + lane], all[ps]);
            }
            __syncthreads();
            if (threadIdx.x < gw && w0 + threadIdx.x < ((m + 31) >> 5)) {
                atomicOr(&tab.seg_any[wo + w0 + threadIdx.x], s_any[threadIdx.x]);
                atomicAnd(&tab.seg_all[wo + w0 + threadIdx.x], s_all[threadIdx.x]);
            }
            __syncthreads();
        }
    }
}

// S (segregating nodes) and the number of variant SITES in the sense of a bubble caller (run_tajd.sh:126-148 counts the
// records `povu gfa2vcf` prints: one per bubble of the window graph, not one per node): with the nodes in graph order, a
// site is a maximal run of segregating nodes that no node carried by EVERY SEG row interrupts; nodes no SEG row carries and
// zero-length nodes neither extend nor interrupt a run.  (A bi-allelic SNP bubble = two segregating nodes = one site.)
// Parity unpinned (povu is not in the reference tree); restated in oracle/similarity.py:site_runs.
__global__ void seg_count_kernel(const __grid_constant__ WindowTab tab, int64_t *counts) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < tab.W; w += gridDim.x * warps_per_block) {
        const int m = tab.m[w];
        const int64_t wo = tab.word_off[w];
        const int words = (m + 31) >> 5;
        int64_t *row = counts + (size_t)w * IMPOP_NCOUNTS;
        const bool have_rows = row[7] > 0;                       // prep_cols left the number of SEG rows here
        int seg = 0, runs = 0;
        bool carry = false;                                       // the last relevant node so far is segregating (lane-uniform)
        for (int w0 = 0; w0 < words && have_rows; w0 += 32) {
            const int wd = w0 + lane;
            uint32_t sg = 0u, sep = 0u;
            if (wd < words) {
                const uint32_t any = tab.seg_any[wo + wd], all = tab.seg_all[wo + wd];
                const uint32_t live = tab.live[wo + wd];          // nodes < m of positive length (prep_cols)
                sg = any & ~all & live;
                sep = all & live;
            }
            if (tab.col_mult) {                                    // a column may stand for several nodes (merged at ingest) or none (a copy)
                const uint8_t *cm = tab.col_mult + tab.len_off[w] + (size_t)wd * 32;
                for (uint32_t b = sg; b; b &= b - 1) seg += cm[__ffs(b) - 1];
            } else {
                seg += __popc(sg);
            }
            // word summary: runs that start inside the word, and the kind of its first / last relevant node
            uint32_t rel = sg | sep;
            int internal = 0;
            bool have_prev = false, prev_seg = false, first_seg = false;
            while (rel) {
                const uint32_t b = rel & (0u - rel);
                const bool is_seg = (sg & b) != 0u;
                if (!have_prev) { first_seg = is_seg; have_prev = true; }
                else if (is_seg && !prev_seg) ++internal;
                prev_seg = is_seg;
                rel ^= b;
            }
            // chain the 32 word summaries in order (lane 0 .. 31)
            for (int l = 0; l < 32; ++l) {
                const int hv = __shfl_sync(0xffffffffu, (int)have_prev, l);
                const int fs = __shfl_sync(0xffffffffu, (int)first_seg, l);
                const int ls = __shfl_sync(0xffffffffu, (int)prev_seg, l);
                const int in = __shfl_sync(0xffffffffu, internal, l);
                if (hv) {
                    runs += in + ((fs && !carry) ? 1 : 0);
                    carry = ls != 0;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) seg += __shfl_xor_sync(0xffffffffu, seg, off);
        __syncwarp();
        if (lane == 0) {
            row[7] = seg;
            const int64_t given = tab.site_runs_given ? tab.site_runs_given[w] : -1;
            tab.site_runs[w] = given >= 0 ? (int32_t)given : runs;
        }
    }
}

// Heavy-entry count per window (batch creation: sizes the heavy table).
__global__ void heavy_count_kernel(const uint32_t *len, const int64_t *len_off, const int32_t *m, int32_t W,
                                   int32_t *out) {
    for (int w = blockIdx.x; w < W; w += gridDim.x) {
        __shared__ int s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int c = 0;
        for (int k = threadIdx.x; k < m[w]; k += blockDim.x) {
            uint32_t q = len[len_off[w] + k] / HEAVY_Q;
            c += (int)((q + 254u) / 255u);
        }
        if (c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (threadIdx.x == 0) out[w] = s_cnt;
        __syncthreads();
    }
}

// ==========================================================================================
// Work-item bookkeeping
// ==========================================================================================
struct Item {
    int w, bi, col0, ncols;   // window, row block, first column, columns (multiple of 16, <= 256)
};

__device__ __forceinline__ Item decode_item(const WindowTab &tab, int64_t t) {
    const int4 v = __ldg(tab.items + t);      // table built on the host at batch creation
    Item it;
    it.w = v.x; it.bi = v.y & ITEM_BI_MASK; it.col0 = v.z; it.ncols = v.w;
    return it;
}

__device__ __forceinline__ void pair_dump(const ItemParams &p, int n, int i, int j, uint32_t inter, uint32_t ai,
                                          uint32_t aj) {
    if (i < n && j < n && j >= i) {
        if (p.dumpI) {
            p.dumpI[(size_t)i * n + j] = (int64_t)inter;
            p.dumpI[(size_t)j * n + i] = (int64_t)inter;
        }
        if (p.dumpPi) {
            double v = (i == j) ? 0.0 : pi_from_counts(inter, ai, aj);
            p.dumpPi[(size_t)i * n + j] = v;
            p.dumpPi[(size_t)j * n + i] = v;
        }
    }
}

// Row-side combination of one warp's per-row totals (s, a, b = sums over the columns carrying
// SUBSET / A / B) into the four pair sums S, AA, BB, AB, reduced over the warp's 32 rows and
// written as one partial record (hi[4], lo[4]).
__device__ __forceinline__ void warp_partial(const dd &ts, const dd &ta, const dd &tb, uint32_t fi, double *out8) {
    const dd zero = {0.0, 0.0};
    dd v[4];
    v[0] = (fi & IMPOP_LAB_SUBSET) ? ts : zero;
    v[1] = (fi & IMPOP_LAB_A) ? ta : zero;
    v[2] = (fi & IMPOP_LAB_B) ? tb : zero;
    v[3] = (fi & IMPOP_LAB_A) ? tb : zero;
    if (fi & IMPOP_LAB_B) dd_merge(v[3], ta);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = warp_sum_dd(v[k]);
        if (lane == 0) { out8[k] = v[k].hi; out8[4 + k] = v[k].lo; }
    }
}

// ==========================================================================================
// tcgen05 implementation, warp specialised.  One CTA per SM, P producer + 4 service + E epilogue warps
// (shipped: P = 4, E = 16, 768 threads):
//   warps 0..P-1      producers: expand the presence bits of a raw slot into the u8 operand tiles of a ring of stages
//                     (P = 4: lane = one A row and two B rows; P = 8 / 12: four A warps + four / eight B warps); shared memory only
//   warp  P           MMA issuer (one elected lane): tcgen05.mma kind::i8 into one of two TMEM buffers
//   warp  P+1         table warp: per item, the column / row values and geometry the epilogue needs, into one of
//                     three tables in shared memory, up to three items ahead
//   warps P+2, P+3    loaders: cp.async of each chunk's packed presence bits and byte weights into a ring of raw slots
//   warps P+4 ..      epilogue (E / 4 per TMEM lane quarter = per SM scheduler): tcgen05.ld, exact integer union, fp64 pi_ij,
//                     per-lane sums
// The fp64 epilogue is the critical role (18 fp64 + 4 XU instructions per pair): it gets four warps per scheduler --
// what the stand-alone micro-benchmark of the same math needs to keep the fp64 pipe 80 % busy -- and the registers the
// service roles give up with setmaxnreg (P = 4, E = 16: 8 x 32 x 48 + 16 x 32 x 96 = 61 440 = the 768 x 80 registers
// the CTA owns; a larger sum makes setmaxnreg.inc wait forever).
// The integer pipes (expansion) and the fp64 pipe (epilogue) so run concurrently on different
// warps, and items flow through without CTA-wide barriers.
// ==========================================================================================
#ifndef IMPOP_PROD_SLEEP
#define IMPOP_PROD_SLEEP 200     // ns between polls of a waiting producer warp
#endif
#ifndef IMPOP_EPI_SLEEP
#define IMPOP_EPI_SLEEP 20       // ns between polls of a waiting epilogue warp
#endif
#ifndef IMPOP_PROD_WARPS
#define IMPOP_PROD_WARPS 4       // 4: every producer warp expands one A row and two B rows per lane; 8: four A warps + four B warps
#endif                           // (two rows per lane); 12: four A warps + eight B warps (one row per lane)
#ifndef IMPOP_EPI_WARPS
#define IMPOP_EPI_WARPS 16       // a multiple of 4: IMPOP_EPI_WARPS / 4 warps share each TMEM lane quarter
#endif
constexpr int WS_PROD_WARPS = IMPOP_PROD_WARPS;
constexpr int WS_EPI_WARPS = IMPOP_EPI_WARPS;
constexpr bool WS_PROD_SHARED = WS_PROD_WARPS == 4;                                   // every producer warp serves both operand tiles
constexpr int WS_B_WARPS = WS_PROD_SHARED ? 4 : WS_PROD_WARPS - 4;
constexpr int WS_B_RPL = TILE_N / (32 * WS_B_WARPS);                                  // B rows per producer lane
static_assert(WS_PROD_WARPS % 4 == 0 && WS_EPI_WARPS % 4 == 0 && WS_B_RPL * 32 * WS_B_WARPS == TILE_N, "warp roles");
static_assert(WS_EPI_WARPS == PART_SLOTS, "one partial record per epilogue warp");
constexpr int WS_MMA_WARP = WS_PROD_WARPS;
constexpr int WS_EPI_WARP0 = WS_PROD_WARPS + 4;
constexpr int WS_THREADS = 32 * (WS_EPI_WARP0 + WS_EPI_WARPS);        // 768 (P = 4, E = 16)
constexpr int WS_STAGES = 3;                                    // operand stages (expanded u8 tiles)
constexpr int WS_RAW = 8;                                       // raw ring: packed presence bits + byte weights of a chunk
constexpr int WS_LOADER_LANES = 64;                             // two loader warps
constexpr int RAW_ROWS = TILE_M + TILE_N;                       // 384 operand rows per chunk
constexpr int WS_TABLES = 3;                                    // item tables in flight (see the table warps)
constexpr int WS_TBL_WARPS = 1;
#ifndef IMPOP_REGS_LOW           // registers of the service warps (producer / MMA / table / loader) and of the epilogue warps
#if IMPOP_PROD_WARPS == 4 && IMPOP_EPI_WARPS == 16
#define IMPOP_REGS_LOW 48
#define IMPOP_REGS_EPI 96
#elif IMPOP_PROD_WARPS == 8 && IMPOP_EPI_WARPS == 16
#define IMPOP_REGS_LOW 40
#define IMPOP_REGS_EPI 96
#elif IMPOP_PROD_WARPS == 8 && IMPOP_EPI_WARPS == 12
#define IMPOP_REGS_LOW 48
#define IMPOP_REGS_EPI 112
#elif IMPOP_PROD_WARPS == 4 && IMPOP_EPI_WARPS == 12
#define IMPOP_REGS_LOW 48
#define IMPOP_REGS_EPI 128
#else
#define IMPOP_REGS_LOW 48
#define IMPOP_REGS_EPI 104       // 12 + 12: 16 x 32 x 48 + 12 x 32 x 104 = 64 512 = 896 x 72
#endif
#endif
constexpr int WS_LAUNCH_REGS = (65536 / WS_THREADS) & ~7;       // what __launch_bounds__(WS_THREADS, 1) lets ptxas allocate
static_assert((WS_PROD_WARPS + 4) * 32 * IMPOP_REGS_LOW + WS_EPI_WARPS * 32 * IMPOP_REGS_EPI <= WS_THREADS * WS_LAUNCH_REGS,
              "setmaxnreg targets exceed the registers the CTA owns");
#ifndef IMPOP_EPI_NP
#define IMPOP_EPI_NP 4           // pairs whose division chains advance together (x 4 warps per scheduler)
#endif
static_assert(IMPOP_EPI_NP == 4, "the epilogue loads and processes its accumulators four columns at a time");
#define IMPOP_STR2(x) #x
#define IMPOP_STR(x) IMPOP_STR2(x)
constexpr int A_STAGE_BYTES = TILE_M * KCHUNK;                  // 16 KB
constexpr int B_STAGE_BYTES = TILE_N * KCHUNK;                  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;      // 48 KB
constexpr uint32_t LBO_A = TILE_M * 16;                         // next 16-byte K slab of the A tile
constexpr uint32_t LBO_B = TILE_N * 16;
constexpr uint32_t SBO_AB = 128;                                // next 8-row group
constexpr uint32_t TMEM_COLS = 512;                             // two accumulator buffers of 256 columns
constexpr int EPI_COLS = TILE_N;                                // columns of one item

struct __align__(16) EpiCols {          // everything the epilogue needs of one item, built by the table warps
    int32_t n, r0, col0, ncols;         // haplotypes of the window, first row, first column, columns of the item
    int32_t have_acc, last, rev, pad1;  // m > 0 (an accumulator exists); last item of this CTA's visit to the window;
                                        // rev: TMEM lane quarter q holds row quarter 3 - q (see ITEM_REV)
    uint32_t aj[EPI_COLS];              // column value of the union: A_j - C + R_j (+ 1 for an empty path), see pi_batch
    uint32_t rj[EPI_COLS];              // R_j (0 for a plain window)
    double fs[EPI_COLS], fa[EPI_COLS], fb[EPI_COLS];   // 1.0 / 0.0: column carries SUBSET / A / B
    uint32_t cmask[EPI_COLS / 16];      // per 16-column chunk: bit 0 all columns valid and in SUBSET, bit 1 any A, bit 2 any B,
                                        // bit 3 all in A, bit 4 all in B (a column carrying a label is a valid column)
    uint32_t ai[TILE_M];                // row value of the union: A_i + R_i (+ 1 for an empty path)
    uint32_t ci[TILE_M];                // C - R_i
    uint32_t fi[TILE_M];                // cleaned labels of the item's rows
};

struct WsShared {
    uint64_t full[WS_STAGES], empty[WS_STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t tbl_full[WS_TABLES], tbl_empty[WS_TABLES];
    uint32_t tmem_base;
    uint32_t pad;
    uint64_t raw_full[WS_RAW], raw_empty[WS_RAW];
    EpiCols col[WS_TABLES];             // 3 x 7.2 KB
};
struct __align__(16) RawSlot {          // one chunk as it lies in global memory: 16 bytes of presence bits per operand row
    uint4 bits[RAW_ROWS];               // rows 0-127: A tile, 128-383: B tile
    uint4 w[KCHUNK / 16];               // the chunk's 128 byte weights (kperm order)
};
constexpr int WS_RAW_OFF = WS_STAGES * STAGE_BYTES;
constexpr int WS_SH_OFF = WS_RAW_OFF + WS_RAW * (int)sizeof(RawSlot);
constexpr int WS_SMEM_BYTES = WS_SH_OFF + (int)sizeof(WsShared);
static_assert(WS_SMEM_BYTES <= 232448, "pairs kernel exceeds the 227 KB of shared memory a CTA can have");

// Items are dealt to CTAs as CONTIGUOUS ranges (consecutive items of a CTA mostly belong to the same window:
// its rows stay in L1/L2, and the per-row sums can be carried across items and reduced once per window).
// Local item u of this launch is global item item_begin + u * world + rank.
#ifdef IMPOP_PROFILE_ROLES
#define PROF_DECL long long pf_wait = 0, pf_work = 0, pf_t = clock64(), pf_aux = 0;
#define PROF_WAIT_END { long long n_ = clock64(); pf_wait += n_ - pf_t; pf_t = n_; }
#define PROF_WORK_END { long long n_ = clock64(); pf_work += n_ - pf_t; pf_t = n_; }
#define PROF_AUX_END { long long n_ = clock64(); pf_aux += n_ - pf_t; pf_t = n_; }
#define PROF_STORE(slot) if (prm.prof && lane == 0) { long long *o_ = prm.prof + (size_t)blockIdx.x * 16 + (slot) * 3; o_[0] = pf_wait; o_[1] = pf_work; o_[2] = pf_aux; }
#else
#define PROF_DECL
#define PROF_WAIT_END
#define PROF_WORK_END
#define PROF_AUX_END
#define PROF_STORE(slot) {}
#endif

// Byte s of the result = 0xFF if bit 8 s + 7 of v is set, else 0 (PRMT with the sign-replicating selector).
// (inline PTX: the __byte_perm intrinsic masks the replicate bit of the selector nibbles away)
__device__ __forceinline__ uint32_t msb_to_bytes(uint32_t v) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(r) : "r"(v));
    return r;
}

// One 16-byte K slab (half h of presence word `word`) of an operand row, in kperm order.
//   MODE 0: bytes 0 / 1       (B operand, dense chunk)
//   MODE 1: bytes 0 / 255     (B operand, heavy chunk)
//   MODE 2: bytes 0 / weight  (A operand)
template <int MODE>
__device__ __forceinline__ uint4 expand_slab(uint32_t word, int h, const uint4 &wv) {
    uint4 o;
    if (MODE == 0) {
        o.x = (word >> (4 * h)) & 0x01010101u;
        o.y = (word >> (4 * h + 1)) & 0x01010101u;
        o.z = (word >> (4 * h + 2)) & 0x01010101u;
        o.w = (word >> (4 * h + 3)) & 0x01010101u;
    } else {
        o.x = msb_to_bytes(word << (7 - 4 * h));
        o.y = msb_to_bytes(word << (6 - 4 * h));
        o.z = msb_to_bytes(word << (5 - 4 * h));
        o.w = msb_to_bytes(word << (4 - 4 * h));
        if (MODE == 2) { o.x &= wv.x; o.y &= wv.y; o.z &= wv.z; o.w &= wv.w; }
    }
    return o;
}

// Per-lane sums of the epilogue.  IMPOP_EPI_PLAIN = 1: plain fp64 (every term is a non-negative pi_ij and a lane adds at
// most a few dozen chunk sums per window, so the relative error is bounded by (terms + tree depth) x 2^-53 ~ 1e-14, far
// inside the 1e-12 contract; the cross-item / cross-warp sums of window_sums_kernel and finalize stay compensated).
// IMPOP_EPI_PLAIN = 0: two-sum compensated (hi, lo) as in the other kernels.
#ifndef IMPOP_EPI_PLAIN
#define IMPOP_EPI_PLAIN 1
#endif
#if IMPOP_EPI_PLAIN
struct esum { double hi; };
__device__ __forceinline__ esum esum_zero() { return esum{0.0}; }
__device__ __forceinline__ void esum_add(esum &a, double x) { a.hi = __dadd_rn(a.hi, x); }
__device__ __forceinline__ void esum_merge(esum &a, const esum &b) { a.hi = __dadd_rn(a.hi, b.hi); }
__device__ __forceinline__ double esum_lo(const esum &) { return 0.0; }
__device__ __forceinline__ esum esum_warp(esum v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v.hi = __dadd_rn(v.hi, __shfl_xor_sync(0xffffffffu, v.hi, off));
    return v;
}
#else
typedef dd esum;
__device__ __forceinline__ esum esum_zero() { return esum{0.0, 0.0}; }
__device__ __forceinline__ void esum_add(esum &a, double x) { dd_add(a, x); }
__device__ __forceinline__ void esum_merge(esum &a, const esum &b) { dd_merge(a, b); }
__device__ __forceinline__ double esum_lo(const esum &a) { return a.lo; }
__device__ __forceinline__ esum esum_warp(esum v) { return warp_sum_dd(v); }
#endif

template <bool DUMP>
__global__ void __launch_bounds__(WS_THREADS, 1)
window_pairs_tc_kernel(const __grid_constant__ WindowTab tab, const __grid_constant__ ItemParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    WsShared &sh = *reinterpret_cast<WsShared *>(smem + WS_SH_OFF);
    RawSlot *raw = reinterpret_cast<RawSlot *>(smem + WS_RAW_OFF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&sh.full[s], WS_PROD_WARPS); mbar_init(&sh.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], WS_EPI_WARPS); }
        for (int b = 0; b < WS_TABLES; ++b) { mbar_init(&sh.tbl_full[b], WS_TBL_WARPS); mbar_init(&sh.tbl_empty[b], WS_EPI_WARPS); }
        for (int b = 0; b < WS_RAW; ++b) { mbar_init(&sh.raw_full[b], WS_LOADER_LANES); mbar_init(&sh.raw_empty[b], WS_PROD_WARPS); }
        fence_mbar_init();
    }
    if (warp == WS_MMA_WARP) tmem_alloc(&sh.tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh.tmem_base;
    const int64_t u_total = (prm.item_end - prm.item_begin - prm.rank + prm.world - 1) / prm.world;
    const int64_t u_lo = u_total * blockIdx.x / gridDim.x, u_hi = u_total * (blockIdx.x + 1) / gridDim.x;
    auto item_of = [&](int64_t u) { return prm.item_begin + u * prm.world + prm.rank; };
    bool alive = true;

    // Per-window values the producer / MMA roles need; re-read only when a CTA's next item belongs to another
    // window (a CTA's items are contiguous, so that is once per ~6 items at n = 466); the item table entry of
    // the next item is prefetched one item ahead.
    struct Win {
        int w = -1, n = 0, m = 0, pitch = 0, dense_chunks = 0, hwords = 0, nch = 0;
        int64_t x_off = 0, w8_off = 0, xh_off = 0;
    };
    auto load_win = [&](Win &wi, int w) {
        if (w == wi.w) return;
        wi.w = w;
        wi.n = tab.n[w]; wi.m = tab.m[w]; wi.pitch = tab.pitch[w];
        const int64_t h0 = tab.heavy_off[w], h1 = tab.heavy_off[w + 1];
        wi.x_off = tab.x_off[w]; wi.w8_off = tab.w8_off[w]; wi.xh_off = tab.xh_off[w];
        wi.dense_chunks = (wi.m + KCHUNK - 1) / KCHUNK;
        wi.hwords = (int)((h1 - h0) >> 5);
        wi.nch = wi.dense_chunks + wi.hwords / (KCHUNK / 32);
    };
    auto raw_item = [&](int64_t u) { return (u < u_hi) ? __ldg(tab.items + item_of(u)) : make_int4(-1, 0, 0, 0); };
    auto raw_ext = [&](int64_t u) { return (u < u_hi) ? __ldg(tab.items_ext + item_of(u)) : make_int4(0, 0, 0, 0); };

    if (warp < WS_PROD_WARPS) {
        // ================================================================ producers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " IMPOP_STR(IMPOP_REGS_LOW) ";");
        // Row tasks of this lane: the A row of its warp quarter (doA) and WS_B_RPL B rows (doB).  With four producer
        // warps every warp does both; otherwise warps 0-3 own the A tile and the rest the B tile.
        const bool doA = WS_PROD_SHARED || warp < 4;
        const bool doB = WS_PROD_SHARED || warp >= 4;
        const int rlA = (warp & 3) * 32 + lane;                                       // row of the A tile
        const int rlB = (WS_PROD_SHARED ? warp : warp - 4) * 32 + lane;               // first row of the B tile
        uint32_t g = 0;                                               // chunks produced so far
        Win wi;
        int4 nxt = raw_item(u_lo);
        PROF_DECL
        for (int64_t u = u_lo; u < u_hi; ++u) {
            const int4 cur = nxt;
            nxt = raw_item(u + 1);
            load_win(wi, cur.x);
            PROF_AUX_END
            const int ncols = cur.w;
            const int nch = wi.nch, dense_chunks = wi.dense_chunks;
            const int a_src = (((cur.y & ITEM_REV) ? 3 - warp : warp) & 3) * 32 + lane;   // A rows: block row behind tile row rlA
            // Presence bits and byte weights arrive through the raw ring (loader warps, cp.async): the producers touch
            // shared memory only, so the proxy fence after the expansion has no global load to wait for.
            for (int c = 0; c < nch; ++c, ++g) {
                const uint32_t s = g % WS_STAGES, rs = g % WS_RAW;
                if (alive) alive = mbar_wait<IMPOP_PROD_SLEEP>(&sh.raw_full[rs], (g / WS_RAW) & 1u, tab.err);
#ifdef IMPOP_PROFILE_ROLES2
                PROF_AUX_END          // aux = wait for the raw slot (+ item setup)
#endif
                const RawSlot &slot = raw[rs];
                uint4 bitsA = make_uint4(0u, 0u, 0u, 0u), bits[WS_B_RPL];
                if (doA) bitsA = slot.bits[a_src];
#pragma unroll
                for (int q = 0; q < WS_B_RPL; ++q) {
                    const int r = rlB + q * (TILE_N / WS_B_RPL);
                    bits[q] = (doB && r < ncols) ? slot.bits[TILE_M + r] : make_uint4(0u, 0u, 0u, 0u);
                }
                if (alive) alive = mbar_wait<IMPOP_PROD_SLEEP>(&sh.empty[s], ((g / WS_STAGES) & 1u) ^ 1u, tab.err);
                PROF_WAIT_END
#ifndef IMPOP_DBG_NO_EXPAND   // (timing experiment: skip the operand expansion)
                const uint4 none = make_uint4(0u, 0u, 0u, 0u);
                if (doA) {
                    uint8_t *dst = smem + s * STAGE_BYTES + rlA * 16;
                    const uint32_t bw[4] = {bitsA.x, bitsA.y, bitsA.z, bitsA.w};
                    // the byte weights of four slabs are loaded before the first store of the group: a load behind a store
                    // to the same shared-memory array is ordered after it (possible alias), which serialised load -> AND ->
                    // store eight times per chunk
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint4 wv[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) wv[q] = slot.w[half * 4 + q];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int slab = half * 4 + q;
                            *reinterpret_cast<uint4 *>(dst + slab * LBO_A) = expand_slab<2>(bw[slab >> 1], slab & 1, wv[q]);
                        }
                    }
                }
                if (doB) {
#pragma unroll
                    for (int q = 0; q < WS_B_RPL; ++q) {
                        const int r = rlB + q * (TILE_N / WS_B_RPL);
                        if (r >= ncols) continue;
                        uint8_t *dst = smem + s * STAGE_BYTES + A_STAGE_BYTES + r * 16;
                        const uint32_t bw[4] = {bits[q].x, bits[q].y, bits[q].z, bits[q].w};
                        if (c < dense_chunks) {
#pragma unroll
                            for (int slab = 0; slab < KCHUNK / 16; ++slab)
                                *reinterpret_cast<uint4 *>(dst + slab * LBO_B) = expand_slab<0>(bw[slab >> 1], slab & 1, none);
                        } else {
#pragma unroll
                            for (int slab = 0; slab < KCHUNK / 16; ++slab)
                                *reinterpret_cast<uint4 *>(dst + slab * LBO_B) = expand_slab<1>(bw[slab >> 1], slab & 1, none);
                        }
                    }
                }
                fence_proxy_async_smem();
#endif
                __syncwarp();
                if (lane == 0) { mbar_arrive(&sh.full[s]); mbar_arrive(&sh.raw_empty[rs]); }
                PROF_WORK_END
            }
        }
        if (warp == 0) PROF_STORE(0)
        if (warp == (WS_PROD_SHARED ? 1 : 4)) PROF_STORE(1)
    } else if (warp < WS_EPI_WARP0) {
        // ================================================================ MMA issuer (warp 12, one lane issues; 13-15 idle)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " IMPOP_STR(IMPOP_REGS_LOW) ";");
        if (warp == WS_MMA_WARP) {
            const uint32_t smem_base_u32 = smem_u32(smem);
            uint32_t g = 0, uses0 = 0u, uses1 = 0u;
            Win wi;
            int4 nxt = raw_item(u_lo);
            PROF_DECL
            for (int64_t u = u_lo; u < u_hi; ++u) {
                const int4 cur = nxt;
                nxt = raw_item(u + 1);
                load_win(wi, cur.x);
                const int nch = wi.nch;
                if (nch == 0) continue;
                PROF_WORK_END
                const uint32_t buf = (uint32_t)(u - u_lo) & 1u;           // item parity = TMEM buffer
                const uint32_t used = buf ? uses1 : uses0;
                if (alive) alive = mbar_wait<100>(&sh.acc_empty[buf], (used & 1u) ^ 1u, tab.err);
                PROF_AUX_END
                tc_fence_after();
                const uint32_t idesc = make_idesc_u8(TILE_M, (uint32_t)cur.w);
                const uint32_t tmem_d = tmem_base + buf * TILE_N;
                for (int c = 0; c < nch; ++c, ++g) {
                    const uint32_t s = g % WS_STAGES;
                    if (alive) alive = mbar_wait<32>(&sh.full[s], (g / WS_STAGES) & 1u, tab.err);
                    PROF_WAIT_END
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_base_u32 + s * STAGE_BYTES;
                        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
                        uint64_t da = make_smem_desc(a_addr, LBO_A, SBO_AB);
                        uint64_t db = make_smem_desc(b_addr, LBO_B, SBO_AB);
#pragma unroll
                        for (int k32 = 0; k32 < KCHUNK / 32; ++k32) {
#ifndef IMPOP_DBG_NO_MMA      // timing experiment only
                            tc_mma_i8(tmem_d, da, db, idesc, (c > 0 || k32 > 0) ? 1u : 0u);
#endif
                            da += (uint64_t)((2 * LBO_A) >> 4);            // start-address field advances by 2 K slabs
                            db += (uint64_t)((2 * LBO_B) >> 4);
                        }
                        tc_commit(&sh.empty[s]);
                        if (c == nch - 1) tc_commit(&sh.acc_full[buf]);
                    }
                    __syncwarp();
                    PROF_WORK_END
                }
                if (buf) ++uses1; else ++uses0;
            }
            PROF_STORE(2)
        } else if (warp == WS_MMA_WARP + 1) {
            // ============================================================ table warp (13): builds, up to three items ahead of
            // the epilogue, everything it needs of an item -- A_j and class flags of the columns, A_i and labels of the
            // rows, the item's geometry -- so that no epilogue warp ever waits for global memory or for another
            // epilogue warp between two items.  tbl_full[slot]: 1 arrival (this warp); tbl_empty[slot]: one per epilogue warp.
            int4 nxt = raw_item(u_lo), nxt_x = raw_ext(u_lo);
            for (int64_t u = u_lo; u < u_hi; ++u) {
                const int k = (int)(u - u_lo), slot = k % WS_TABLES;
                const int4 cur = nxt, ext = nxt_x;
                nxt = raw_item(u + 1); nxt_x = raw_ext(u + 1);
                const int n = ext.x, col0 = cur.z, ncols = cur.w;
                const uint8_t *lab = tab.labels + (size_t)ext.w;
                const int32_t *Aw = tab.A + (size_t)ext.z;
                const int32_t *Rw = tab.row_adj ? tab.row_adj + (size_t)ext.z : nullptr;       // affine form (pi_batch)
                const uint32_t cw = tab.win_const ? (uint32_t)__ldg(tab.win_const + cur.x) : 0u;
                // issue the global loads before waiting for the slot
                uint32_t cf[EPI_COLS / 32], ca[EPI_COLS / 32], cr[EPI_COLS / 32];
#pragma unroll
                for (int ps = 0; ps < EPI_COLS / 32; ++ps) {
                    const int cc = ps * 32 + lane, j = col0 + cc;
                    const bool ok = cc < ncols && j < n;
                    cf[ps] = ok ? clean_label(__ldg(lab + j)) : 0u;
                    const uint32_t a = ok ? (uint32_t)__ldg(Aw + j) : 1u, r = (ok && Rw) ? (uint32_t)__ldg(Rw + j) : 0u;
                    ca[ps] = ok ? a - cw + r + (a == 0u ? 1u : 0u) : 1u;     // empty path: + 1 (see pi_batch)
                    cr[ps] = r;
                }
                if (alive) alive = mbar_wait<100>(&sh.tbl_empty[slot], ((uint32_t)(k / WS_TABLES) & 1u) ^ 1u, tab.err);
                EpiCols &col = sh.col[slot];
#pragma unroll
                for (int ps = 0; ps < EPI_COLS / 32; ++ps) {
                    const int cc = ps * 32 + lane;
                    const uint32_t f = cf[ps];
                    col.aj[cc] = ca[ps];
                    col.rj[cc] = cr[ps];
                    col.fs[cc] = (f & IMPOP_LAB_SUBSET) ? 1.0 : 0.0;
                    col.fa[cc] = (f & IMPOP_LAB_A) ? 1.0 : 0.0;
                    col.fb[cc] = (f & IMPOP_LAB_B) ? 1.0 : 0.0;
                    const uint32_t bs = __ballot_sync(0xffffffffu, (f & IMPOP_LAB_SUBSET) != 0u);
                    const uint32_t ba = __ballot_sync(0xffffffffu, (f & IMPOP_LAB_A) != 0u);
                    const uint32_t bb = __ballot_sync(0xffffffffu, (f & IMPOP_LAB_B) != 0u);
                    if (lane < 2) {
                        const uint32_t hs = lane ? (bs >> 16) : (bs & 0xFFFFu), ha = lane ? (ba >> 16) : (ba & 0xFFFFu),
                                       hb = lane ? (bb >> 16) : (bb & 0xFFFFu);
                        col.cmask[ps * 2 + lane] = (hs == 0xFFFFu ? 1u : 0u) | (ha ? 2u : 0u) | (hb ? 4u : 0u) |
                                                   (ha == 0xFFFFu ? 8u : 0u) | (hb == 0xFFFFu ? 16u : 0u);
                    }
                }
#pragma unroll
                for (int ps = 0; ps < TILE_M / 32; ++ps) {     // the row side (loaded here: this warp runs items ahead, and has 48 registers)
                    const int i = (cur.y & ITEM_BI_MASK) * TILE_M + ps * 32 + lane;
                    const uint32_t a = i < n ? (uint32_t)__ldg(Aw + i) : 1u, r = (i < n && Rw) ? (uint32_t)__ldg(Rw + i) : 0u;
                    col.ai[ps * 32 + lane] = i < n ? a + r + (a == 0u ? 1u : 0u) : 1u;
                    col.ci[ps * 32 + lane] = i < n ? cw - r : 0u;
                    col.fi[ps * 32 + lane] = i < n ? clean_label(__ldg(lab + i)) : 0u;
                }
                if (lane == 0) {
                    col.n = n; col.r0 = (cur.y & ITEM_BI_MASK) * TILE_M; col.col0 = col0; col.ncols = ncols;
                    col.have_acc = ext.y > 0 ? 1 : 0; col.last = (nxt.x != cur.x) ? 1 : 0; col.rev = (cur.y & ITEM_REV) ? 1 : 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.tbl_full[slot]);
            }
        } else {
            // ============================================================ loader warps: copy every chunk as it lies in
            // global memory -- 16 bytes of presence bits per operand row, the 128 byte weights -- into the raw ring with
            // cp.async (no registers, no waiting: up to WS_RAW chunks in flight), completion counted by the slot's mbarrier.
            const int ll = (warp - WS_MMA_WARP - 2) * 32 + lane;     // 0..63
            constexpr int ROWS_PER_LANE = RAW_ROWS / WS_LOADER_LANES;   // 6
            uint32_t g = 0;
            Win wi;
            int4 nxt = raw_item(u_lo);
            const uint32_t raw_u32 = smem_u32(raw);
            PROF_DECL
            for (int64_t u = u_lo; u < u_hi; ++u) {
                const int4 cur = nxt;
                nxt = raw_item(u + 1);
                load_win(wi, cur.x);
                PROF_AUX_END
                const int nch = wi.nch, dense_chunks = wi.dense_chunks, hwords = wi.hwords;
                // Once per item: the global row behind each of this lane's slot rows and its copy size (a slot row without a
                // haplotype behind it copies 0 bytes -- zero fill -- from row 0).  Per chunk and row: one wide multiply-add
                // for the address and the copy.
                const uint32_t *xw = tab.x + wi.x_off, *xhw = tab.xh + wi.xh_off;
                const uint8_t *w8 = tab.w8 + wi.w8_off;
                uint32_t grow[ROWS_PER_LANE], sz[ROWS_PER_LANE];
#pragma unroll
                for (int r = 0; r < ROWS_PER_LANE; ++r) {
                    const int tr = ll + r * WS_LOADER_LANES;            // 0..383: row of the raw slot
                    const bool isArow = tr < TILE_M;
                    const int gr = isArow ? (cur.y & ITEM_BI_MASK) * TILE_M + tr : cur.z + (tr - TILE_M);
                    const bool ok = (isArow || tr - TILE_M < cur.w) && gr < wi.n;
                    grow[r] = ok ? (uint32_t)gr : 0u;
                    sz[r] = ok ? 16u : 0u;
                }
                for (int c = 0; c < nch; ++c, ++g) {
                    const uint32_t rs = g % WS_RAW;
                    if (alive) alive = mbar_wait<100>(&sh.raw_empty[rs], ((g / WS_RAW) & 1u) ^ 1u, tab.err);
                    PROF_WAIT_END
                    const uint32_t slot_u32 = raw_u32 + rs * (uint32_t)sizeof(RawSlot) + (uint32_t)ll * 16u;
                    const bool heavy_chunk = c >= dense_chunks;
                    const uint32_t *base = heavy_chunk ? xhw + 4 * (c - dense_chunks) : xw + 4 * c;
                    const uint32_t stride = (uint32_t)(heavy_chunk ? hwords : wi.pitch);
#ifndef IMPOP_DBG_NO_LOAD     // timing experiment only: no copies into the raw ring
#pragma unroll
                    for (int r = 0; r < ROWS_PER_LANE; ++r)
                        cp_async16(slot_u32 + (uint32_t)(r * WS_LOADER_LANES) * 16u, base + (size_t)grow[r] * stride, sz[r]);
                    if (ll < KCHUNK / 16)
                        cp_async16(slot_u32 + (uint32_t)RAW_ROWS * 16u, w8 + (size_t)c * KCHUNK + ll * 16, 16u);
#endif
                    cp_async_arrive_noinc(&sh.raw_full[rs]);
                    PROF_WORK_END
                }
            }
#ifdef IMPOP_PROFILE_ROLES2
            if (ll == 0) PROF_STORE(4)     // replaces epilogue team 1 in the report
#endif
        }
    } else {
        // ================================================================ epilogue.  All epilogue warps work on the same item
        // (warp -> 32-row quarter q4 of the TMEM lanes x one share of that quarter's valid 16-column chunks) while the
        // MMA of the next item fills the other TMEM buffer.  Everything else an item needs comes from its table in
        // shared memory (table warps above): a warp that finishes its chunks early moves on to the next item as soon
        // as the MMA has filled that accumulator -- no barrier couples the epilogue warps.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " IMPOP_STR(IMPOP_REGS_EPI) ";");
        const int e = warp - WS_EPI_WARP0;              // 0 .. WS_EPI_WARPS - 1
        const int q4 = warp & 3;                          // TMEM lane quarter this warp may read
        constexpr int H = WS_EPI_WARPS / 4;               // warps per TMEM lane quarter
        const int hsel = e >> 2;                          // which share of the quarter's chunks
        uint32_t uses0 = 0u, uses1 = 0u;
        esum v[4] = {esum_zero(), esum_zero(), esum_zero(), esum_zero()};   // this lane's S, AA, BB, AB sums of the current window
        PROF_DECL
        for (int64_t u = u_lo; u < u_hi; ++u) {
            const int k = (int)(u - u_lo);
            const int slot = k % WS_TABLES;
            const int64_t t = item_of(u);
            const uint32_t buf = (uint32_t)k & 1u;                        // item parity = TMEM buffer
            if (alive) alive = mbar_wait<IMPOP_EPI_SLEEP>(&sh.tbl_full[slot], (uint32_t)(k / WS_TABLES) & 1u, tab.err);
            const EpiCols &col = sh.col[slot];
            const int4 geo = *reinterpret_cast<const int4 *>(&col.n);     // n, first row, first column, columns
            const int n = geo.x, col0 = geo.z;
            const bool have_acc = col.have_acc != 0, last = col.last != 0;
            const int rq = col.rev ? 3 - q4 : q4;                         // row quarter behind this warp's TMEM lanes
            const int r0 = geo.y + rq * 32;
            const int i = r0 + lane;
            const uint32_t ai = col.ai[rq * 32 + lane], ci = col.ci[rq * 32 + lane];
            const uint32_t fi = col.fi[rq * 32 + lane];
            // this quarter's valid chunks [c_lo, c_hi): columns below n, not entirely left of the diagonal; split in two
            int c_lo = 0, c_hi = 0;
            const int cwidth = min(geo.w, n - col0);                      // columns of the item inside the matrix
            if (r0 < n) {
                const int width = cwidth;
                c_hi = (width + 15) >> 4;
                c_lo = max(0, (r0 - col0) >> 4);
                if (c_lo > c_hi) c_lo = c_hi;
            }
            const int hh = (hsel + k) % H, cnt = c_hi - c_lo;          // shares rotate with the item: remainders even out
            const int cbeg = c_lo + cnt * hh / H, cend = c_lo + cnt * (hh + 1) / H;
            PROF_AUX_END
            if (have_acc) {
                if (alive) alive = mbar_wait<IMPOP_EPI_SLEEP>(&sh.acc_full[buf], (buf ? uses1 : uses0) & 1u, tab.err);
                tc_fence_after();
            }
            PROF_WAIT_END
            esum ts = esum_zero(), ta = esum_zero(), tb = esum_zero();
            const uint32_t tm0 = tmem_base + buf * TILE_N + ((uint32_t)(q4 * 32) << 16);
            // One 16-column chunk of this warp's 32 rows, IMPOP_EPI_NP = 4 pairs at a time: the contract's fp64 sequence, the
            // group's pi_ij straight into the chunk's class sums (only the accumulator registers and one group's division
            // chains are live: 16 epilogue warps at 96 registers).  Two bodies behind one warp-uniform branch per chunk:
            //   FAST     every class holds all 16 columns or none: one tcgen05.ld of 16 columns, the four groups unrolled, one
            //            shared pairwise tree -- 17 fp64 + 9 other instructions per pair;
            //   general  per-column class flags (chunks that straddle a population boundary or the matrix edge): the group
            //            loop is NOT unrolled and loads four columns at a time, so that this rarely used body stays small.
            // Why two bodies: the kernel is bound by issue slots (an fp64 instruction holds the port for two cycles,
            // tools/micro/issue_mix.cu), and with the class tests inside one body ptxas predicates the flag paths -- 3 fp64 +
            // 1.5 load instructions per pair issued and discarded on every chunk.  Why the general one is small: the hot code
            // of a scheduler's four epilogue warps has to stay in the instruction cache; two unrolled bodies (2 x 8 KB) ran
            // 30 % slower than either alone.  (Also measured: ONE not-unrolled group loop for both variants with the next
            // group's tcgen05.ld issued under the current group's divisions -- 5 KB of hot code, but ~4 more instructions
            // per pair for the loop: 1.59 against 1.51 ms.)
            auto group = [&](const uint32_t *r, int cc, int g0, bool on_diag, auto fast_tag, double &cs, double &ca, double &cb,
                             double &call) {
                constexpr bool FAST = decltype(fast_tag)::value;
                const int jbase = col0 + cc;
                uint32_t aj[IMPOP_EPI_NP], rj[IMPOP_EPI_NP];
                double p[IMPOP_EPI_NP];
                {
                    const uint4 q = *reinterpret_cast<const uint4 *>(&col.aj[cc + g0]);
                    aj[0] = q.x; aj[1] = q.y; aj[2] = q.z; aj[3] = q.w;
                    const uint4 t = *reinterpret_cast<const uint4 *>(&col.rj[cc + g0]);
                    rj[0] = t.x; rj[1] = t.y; rj[2] = t.z; rj[3] = t.w;
                }
#ifdef IMPOP_DBG_NO_EPI      // timing experiment only: skip the fp64 math
#pragma unroll
                for (int q = 0; q < IMPOP_EPI_NP; ++q) p[q] = __hiloint2double(r[q] + aj[q] - rj[q], ai + ci);
#else
                pi_batch<IMPOP_EPI_NP>(r, ai, aj, ci, rj, p);
#endif
                if (on_diag) {
#pragma unroll
                    for (int q = 0; q < IMPOP_EPI_NP; ++q) p[q] = (jbase + g0 + q > i) ? p[q] : 0.0;
                }
                if (DUMP) {
#pragma unroll
                    for (int q = 0; q < IMPOP_EPI_NP; ++q)      // true intersection; ai + ci and aj - rj leave the union as it is
                        pair_dump(prm, n, i, jbase + g0 + q, r[q] + ci - rj[q], ai + ci, aj[q] - rj[q]);
                }
                if (FAST) {                                     // plain pairwise tree over the group, shared by the classes
                    call = __dadd_rn(call, __dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3])));
                } else {                                        // p * 1.0 or p * 0.0: exact
#pragma unroll
                    for (int q = 0; q < IMPOP_EPI_NP; ++q) cs = __fma_rn(p[q], col.fs[cc + g0 + q], cs);
#pragma unroll
                    for (int q = 0; q < IMPOP_EPI_NP; ++q) ca = __fma_rn(p[q], col.fa[cc + g0 + q], ca);
#pragma unroll
                    for (int q = 0; q < IMPOP_EPI_NP; ++q) cb = __fma_rn(p[q], col.fb[cc + g0 + q], cb);
                }
            };
#pragma unroll 1
            for (int c = cbeg; c < cend; ++c) {
#ifdef IMPOP_DBG_NO_CLASS      // timing experiment only: every chunk treated as all-SUBSET, no A / B columns
                const uint32_t cm = 1u;
#else
                const uint32_t cm = col.cmask[c];
#endif
#ifdef IMPOP_DBG_FORCE_GENERAL  // timing experiment only
                const bool fast = false;
#else
                // every column in SUBSET, and A (B) holds all of them or none
                // (a window without columns has no accumulator: its chunks take the general body, which tests for that)
                const bool fast = have_acc && (cm & 1u) && ((cm & (2u | 8u)) != 2u) && ((cm & (4u | 16u)) != 4u);
#endif
                const int cc = c << 4;
                const bool on_diag = col0 + cc <= r0 + 31;         // the chunk touches the diagonal of this warp's rows
                double cs = 0.0, ca = 0.0, cb = 0.0, call = 0.0;   // sums over the chunk's SUBSET / A / B / all columns
                if (fast) {
                    uint32_t r[16];
                    tmem_ld16(tm0 + (uint32_t)cc, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int g0 = 0; g0 < 16; g0 += IMPOP_EPI_NP) group(r + g0, cc, g0, on_diag, std::true_type{}, cs, ca, cb, call);
                    esum_add(ts, call);                             // fast implies all 16 columns in SUBSET
                    if (cm & 8u) esum_add(ta, call);
                    if (cm & 16u) esum_add(tb, call);
                } else {
                    // the last chunk of a row block usually holds only a few columns of the matrix: groups beyond them are skipped
                    const int glim = min(16, (cwidth - cc + IMPOP_EPI_NP - 1) & ~(IMPOP_EPI_NP - 1));
#pragma unroll 1
                    for (int g0 = 0; g0 < glim; g0 += IMPOP_EPI_NP) {
                        uint32_t r[IMPOP_EPI_NP];
                        if (have_acc) {
                            tmem_ld4(tm0 + (uint32_t)(cc + g0), r);
                            tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int q = 0; q < IMPOP_EPI_NP; ++q) r[q] = 0u;
                        }
                        group(r, cc, g0, on_diag, std::false_type{}, cs, ca, cb, call);
                    }
                    esum_add(ts, cs);
                    if (cm & 2u) esum_add(ta, ca);
                    if (cm & 4u) esum_add(tb, cb);
                }
            }
            if (have_acc) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&sh.acc_empty[buf]); mbar_arrive(&sh.tbl_empty[slot]); }
                if (buf) ++uses1; else ++uses0;
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.tbl_empty[slot]);
            }
            PROF_WORK_END
            // row-side class combination, carried in this lane across the CTA's items of the same window
#ifdef IMPOP_DBG_NO_REDUCE      // timing experiment only: skip the per-item merges and the per-window warp reduction
            v[0].hi += ts.hi + ta.hi + tb.hi;   // (esum or dd: .hi exists in both)
            double *rec = prm.partials + t * PART_STRIDE + e * 8;
            if (last) { if (lane < 8) rec[lane] = v[0].hi; v[0].hi = 0.0; }
            if (true) {
            } else if (last) {
#else
            if (fi & IMPOP_LAB_SUBSET) esum_merge(v[0], ts);
            if (fi & IMPOP_LAB_A) { esum_merge(v[1], ta); esum_merge(v[3], tb); }
            if (fi & IMPOP_LAB_B) { esum_merge(v[2], tb); esum_merge(v[3], ta); }
            double *rec = prm.partials + t * PART_STRIDE + e * 8;
            if (last) {
#endif                                        // reduce over the warp's 32 lanes, once per window visit
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const esum tot = esum_warp(v[q]);
                    if (lane == 0) { rec[q] = tot.hi; rec[4 + q] = esum_lo(tot); }
                    v[q] = esum_zero();
                }
            } else if (lane < 8) {
                rec[lane] = 0.0;                               // the sums travel on to the next item's record
            }
        }
        if (e == 0) PROF_STORE(3)
#ifndef IMPOP_PROFILE_ROLES2
        if (e == 4) PROF_STORE(4)
#endif
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WS_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ==========================================================================================
// SIMT implementation (dp4a), the cross-check path.  128 threads = the 128 rows of the item's row
// block; columns are processed 16 at a time with u32 accumulators in registers.  Same items, same
// virtual columns, same contract arithmetic (__ddiv_rn).
// ==========================================================================================
constexpr int SIMT_THREADS = 128;
constexpr int SIMT_COLS = 16;
constexpr int SIMT_K = 64;       // virtual columns per staging step

__global__ void __launch_bounds__(SIMT_THREADS) window_pairs_simt_kernel(const __grid_constant__ WindowTab tab,
                                                                         const __grid_constant__ ItemParams prm) {
    __shared__ __align__(16) uint32_t s_b[SIMT_COLS][SIMT_K / 4 + 4];  // +4 words: rows stay 16-byte aligned
    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t stride = (int64_t)gridDim.x * prm.world;
    for (int64_t t = prm.item_begin + (int64_t)blockIdx.x * prm.world + prm.rank; t < prm.item_end; t += stride) {
        const Item it = decode_item(tab, t);
        const int n = tab.n[it.w], pitch = tab.pitch[it.w], m = tab.m[it.w];
        const uint32_t *x = tab.x + tab.x_off[it.w];
        const int dense_chunks = ((m + KCHUNK - 1) / KCHUNK) * (KCHUNK / SIMT_K);   // in 64-column sub-chunks
        const int hwords = (int)((tab.heavy_off[it.w + 1] - tab.heavy_off[it.w]) >> 5);
        const int nch = dense_chunks + (hwords >> 1);
        const uint32_t *xh = tab.xh + tab.xh_off[it.w];
        const uint8_t *w8 = tab.w8n + tab.w8_off[it.w];
        const int32_t *Aw = tab.A + tab.row_off[it.w];
        const uint8_t *lab = tab.labels + tab.lab_off[it.w];
        const int i = it.bi * TILE_M + tid;
        const bool rvalid = i < n;
        const uint32_t ai = rvalid ? (uint32_t)Aw[i] : 0u;
        const uint32_t fi = rvalid ? clean_label(lab[i]) : 0u;
        const int32_t *Rw = tab.row_adj ? tab.row_adj + tab.row_off[it.w] : nullptr;         // affine form: I = cnt + C - R_i - R_j
        const uint32_t ci = (tab.win_const ? (uint32_t)tab.win_const[it.w] : 0u) - ((rvalid && Rw) ? (uint32_t)Rw[i] : 0u);
        const bool dump = (prm.dumpI != nullptr) || (prm.dumpPi != nullptr);
        dd ts = {0.0, 0.0}, ta = {0.0, 0.0}, tb = {0.0, 0.0};

        for (int cc = 0; cc < it.ncols; cc += SIMT_COLS) {
            const int jbase = it.col0 + cc;
            if (jbase >= n) break;
            if (jbase + SIMT_COLS - 1 < it.bi * TILE_M) continue;
            uint32_t cnt[SIMT_COLS];
#pragma unroll
            for (int j = 0; j < SIMT_COLS; ++j) cnt[j] = 0u;
            for (int c = 0; c < nch; ++c) {
                const bool is_heavy = c >= dense_chunks;
                __syncthreads();
                {   // stage B': thread -> (column jr = tid / 8, 8 virtual columns = one byte of bits)
                    const int jr = tid >> 3, slab = tid & 7;
                    const int gj = jbase + jr;
                    uint32_t word = 0;
                    if (gj < n)
                        word = is_heavy ? xh[(size_t)gj * hwords + 2 * (c - dense_chunks) + (slab >> 2)]
                                        : __ldg(x + (size_t)gj * pitch + 2 * c + (slab >> 2));
                    const uint32_t b8 = (word >> ((slab & 3) * 8)) & 0xFFu;
                    const uint2 wv = *reinterpret_cast<const uint2 *>(w8 + (size_t)c * SIMT_K + slab * 8);
                    uint2 o;
                    o.x = (nibble_to_bytes01(b8 & 0xFu) * 255u) & wv.x;
                    o.y = (nibble_to_bytes01(b8 >> 4) * 255u) & wv.y;
                    *reinterpret_cast<uint2 *>(&s_b[jr][slab * 2]) = o;
                }
                uint32_t a[SIMT_K / 4];
                {
                    uint2 bits = make_uint2(0u, 0u);
                    if (rvalid)
                        bits = is_heavy ? *reinterpret_cast<const uint2 *>(xh + (size_t)i * hwords + 2 * (c - dense_chunks))
                                        : __ldg(reinterpret_cast<const uint2 *>(x + (size_t)i * pitch + 2 * c));
                    const uint32_t mul = is_heavy ? 255u : 1u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        a[q] = nibble_to_bytes01((bits.x >> (4 * q)) & 0xFu) * mul;
                        a[8 + q] = nibble_to_bytes01((bits.y >> (4 * q)) & 0xFu) * mul;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < SIMT_COLS; ++j) {
#pragma unroll
                    for (int q4 = 0; q4 < SIMT_K / 16; ++q4) {
                        uint4 b = *reinterpret_cast<const uint4 *>(&s_b[j][q4 * 4]);
                        cnt[j] = __dp4a(a[q4 * 4 + 0], b.x, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 1], b.y, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 2], b.z, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 3], b.w, cnt[j]);
                    }
                }
            }
            double cs = 0.0, ca = 0.0, cb = 0.0;
#pragma unroll
            for (int k = 0; k < SIMT_COLS; ++k) {
                const int j = jbase + k;
                const bool jv = j < n;
                const uint32_t aj = jv ? (uint32_t)__ldg(Aw + j) : 0u;
                const uint32_t fj = jv ? clean_label(__ldg(lab + j)) : 0u;
                const uint32_t inter = cnt[k] + ci - ((jv && Rw) ? (uint32_t)__ldg(Rw + j) : 0u);
                if (fj != 0u) {
                    const double p = (rvalid && j > i) ? pi_from_counts(inter, ai, aj) : 0.0;
                    if (fj & IMPOP_LAB_SUBSET) cs = __dadd_rn(cs, p);
                    if (fj & IMPOP_LAB_A) ca = __dadd_rn(ca, p);
                    if (fj & IMPOP_LAB_B) cb = __dadd_rn(cb, p);
                }
                if (dump) pair_dump(prm, n, i, j, inter, ai, aj);
            }
            dd_add(ts, cs); dd_add(ta, ca); dd_add(tb, cb);
        }
        __syncthreads();
        warp_partial(ts, ta, tb, fi, prm.partials + t * PART_STRIDE + warp * 8);
        if (tid < (PART_SLOTS - SIMT_THREADS / 32) * 8)             // unused partial slots of this item
            prm.partials[t * PART_STRIDE + (SIMT_THREADS / 32) * 8 + tid] = 0.0;
    }
}

// ==========================================================================================
// Window sums (fixed-order reduction of the items' partial records) and finalize.
// ==========================================================================================
__global__ void window_sums_kernel(const __grid_constant__ WindowTab tab, const double *partials, int32_t rank,
                                   int32_t world, double *sums) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < tab.W; w += gridDim.x * warps_per_block) {
        const int64_t t0 = tab.item_off[w], t1 = tab.item_off[w + 1];
        dd v[4] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        // items of this window handled by `rank`: t == rank (mod world); lane -> (item, slot)
        const int64_t first = t0 + ((rank - (t0 % world)) % world + world) % world;
        const int64_t mine = (t1 > first) ? (t1 - first + world - 1) / world : 0;
        for (int64_t r = lane; r < mine * PART_SLOTS; r += 32) {
            const int64_t t = first + (r / PART_SLOTS) * world;
            const double *rec = partials + t * PART_STRIDE + (r % PART_SLOTS) * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dd p = {rec[k], rec[4 + k]};
                dd_merge(v[k], p);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = warp_sum_dd(v[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sums[(size_t)w * 4 + k] = dd_value(v[k]);
        }
    }
}

// The same for batches of few, large windows (10 000 haplotypes: ~3 200 items x 16 records each): one CTA per window, its
// eight warps take the records warp-strided, the eight warp sums are added in warp order -- fixed order, reproducible.
__global__ void __launch_bounds__(256) window_sums_block_kernel(const __grid_constant__ WindowTab tab, const double *partials, int32_t rank,
                                                                int32_t world, double *sums) {
    __shared__ double s_hi[8][4], s_lo[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = blockIdx.x; w < tab.W; w += gridDim.x) {
        const int64_t t0 = tab.item_off[w], t1 = tab.item_off[w + 1];
        dd v[4] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        const int64_t first = t0 + ((rank - (t0 % world)) % world + world) % world;
        const int64_t mine = (t1 > first) ? (t1 - first + world - 1) / world : 0;
        for (int64_t r = threadIdx.x; r < mine * PART_SLOTS; r += 256) {
            const int64_t t = first + (r / PART_SLOTS) * world;
            const double *rec = partials + t * PART_STRIDE + (r % PART_SLOTS) * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dd p = {rec[k], rec[4 + k]};
                dd_merge(v[k], p);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = warp_sum_dd(v[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { s_hi[warp][k] = v[k].hi; s_lo[warp][k] = v[k].lo; }
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            dd acc = {0.0, 0.0};
            for (int q = 0; q < 8; ++q) {
                dd p = {s_hi[q][threadIdx.x], s_lo[q][threadIdx.x]};
                dd_merge(acc, p);
            }
            sums[(size_t)w * 4 + threadIdx.x] = dd_value(acc);
        }
        __syncthreads();
    }
}

__global__ void finalize_kernel(const __grid_constant__ WindowTab tab, const double *sums, int32_t parts,
                                const int64_t *counts, double *stats) {
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < tab.W; w += gridDim.x * blockDim.x) {
        double s[4];
        {
            dd acc[4] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
            for (int p = 0; p < parts; ++p)
#pragma unroll
                for (int k = 0; k < 4; ++k) dd_add(acc[k], sums[((size_t)p * tab.W + w) * 4 + k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] = dd_value(acc[k]);
        }
        int64_t cnt[IMPOP_NCOUNTS];
#pragma unroll
        for (int k = 0; k < IMPOP_NCOUNTS; ++k) cnt[k] = counts[(size_t)w * IMPOP_NCOUNTS + k];
        finalize_row(s, cnt, tab.L[w], (double)cnt[7], tab.harm, tab.harm_n, stats + (size_t)w * IMPOP_NSTATS);
        stats[(size_t)w * IMPOP_NSTATS + IMPOP_ST_S_BUBBLES] = (double)tab.site_runs[w];
    }
}

// A (int32 scratch) -> caller's int64 array for one window.
__global__ void export_a_kernel(const int32_t *A, int32_t n, int64_t *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)A[i];
}

// Self-test of div_rn_inrange / pi_from_counts_fast against __ddiv_rn / pi_from_counts on pseudo-random
// in-range operands; out[0] += number of mismatching results.
__global__ void division_selftest_kernel(uint64_t seed, int64_t count, unsigned long long *out) {
    unsigned long long bad = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);   // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const uint32_t sh1 = (uint32_t)(z & 31u), sh2 = (uint32_t)((z >> 5) & 31u);
        uint32_t ai = ((uint32_t)(z >> 10) & 0x3FFFFFFFu) >> (sh1 % 30u);
        uint32_t aj = ((uint32_t)(z >> 33) & 0x3FFFFFFFu) >> (sh2 % 30u);
        uint64_t z2 = z * 0xD6E8FEB86659FD93ull;
        uint32_t lim = ai < aj ? ai : aj;
        uint32_t inter = lim ? (uint32_t)((z2 >> 20) % ((uint64_t)lim + 1u)) : 0u;
        if ((z2 & 7u) == 0u) inter = lim;                    // one path contained in the other
        if ((z2 & 15u) == 1u) inter = 0u;
        const double p0 = pi_from_counts(inter, ai, aj), p1 = pi_from_counts_fast(inter, ai, aj);
        if (__double_as_longlong(p0) != __double_as_longlong(p1)) ++bad;
        const uint32_t uni = ai + aj - inter;
        const double a = u32_to_double(inter), b = u32_to_double(uni ? uni : 1u);
        if (__double_as_longlong(__ddiv_rn(a, b)) != __double_as_longlong(div_rn_inrange(a, b))) ++bad;
        if (__double_as_longlong(__ddiv_rn(a, b)) != __double_as_longlong(div_rn_int31(a, b))) ++bad;
        {   // ratios of arbitrary 31-bit integers (not only I <= U), small denominators, near-equal operands
            const uint32_t a2 = (uint32_t)(z2 >> 13) & 0x7FFFFFFFu;
            uint32_t b2 = (uint32_t)(z >> 7) & 0x7FFFFFFFu;
            if ((z2 & 3u) == 2u) b2 >>= (sh2 % 31u);
            if ((z2 & 31u) == 3u) b2 = a2 + 1u - (uint32_t)((z2 >> 5) & 3u);
            b2 = (b2 & 0x7FFFFFFFu) ? (b2 & 0x7FFFFFFFu) : 1u;
            const double da = u32_to_double(a2), db = u32_to_double(b2);
            if (__double_as_longlong(__ddiv_rn(da, db)) != __double_as_longlong(div_rn_int31(da, db))) ++bad;
        }
        {   // the epilogue's own code path (pi_batch: four chains, seeds whose low words are the integers, affine operands):
            // the pair above plus three derived ones, each split at random into accumulator, row and column values
            uint32_t acc4[4], aj4[4], rj4[4], want_i[4], want_ai[4], want_aj[4];
            const uint32_t ai_t = (uint32_t)(z2 >> 3), ci = (uint32_t)(z >> 17) * 2654435761u;
            for (int q = 0; q < 4; ++q) {
                const uint32_t a_i = q == 0 ? ai : (ai >> q) + (uint32_t)q, a_j = q == 0 ? aj : (aj >> (q ^ 1)) + 1u;
                const uint32_t lm = a_i < a_j ? a_i : a_j;
                const uint32_t it = q == 0 ? inter : (lm ? (uint32_t)((z2 >> (8 + q)) % ((uint64_t)lm + 1u)) : 0u);
                uint32_t un = a_i + a_j - it;
                un = un ? un : 1u;
                rj4[q] = (uint32_t)(z2 >> (11 * q)) ^ (uint32_t)z;
                acc4[q] = it - ci + rj4[q];                       // inter = acc + ci - rj
                aj4[q] = un + acc4[q] - ai_t;                     // union = ai + aj - acc
                want_i[q] = it; want_ai[q] = a_i; want_aj[q] = a_j;
            }
            double p4[4];
            pi_batch<4>(acc4, ai_t, aj4, ci, rj4, p4);
            for (int q = 0; q < 4; ++q)
                if (__double_as_longlong(pi_from_counts(want_i[q], want_ai[q], want_aj[q])) != __double_as_longlong(p4[q])) ++bad;
        }
    }
    if (bad) atomicAdd(out, bad);
}

// ------------------------------------------------------------------------------------------
// Launchers (called from api.cu)
// ------------------------------------------------------------------------------------------
cudaError_t launch_heavy_count(const uint32_t *len, const int64_t *len_off, const int32_t *m, int32_t W, int32_t *out,
                               cudaStream_t st) {
    if (W == 0) return cudaSuccess;
    heavy_count_kernel<<<min(W, 4096), 128, 0, st>>>(len, len_off, m, W, out);
    return cudaGetLastError();
}

int prep_rows_ctas_per_sm() {                      // resident prep_rows CTAs per SM (registers / shared memory) on the current device
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, prep_rows_kernel, PREP_THREADS, 0) != cudaSuccess || v < 1) v = 2;
    return v;
}

cudaError_t launch_prep(const WindowTab &tab, int64_t *counts, int sm_count, int per_sm, cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    prep_cols_kernel<<<min(tab.W, sm_count * 8), PREP_THREADS, 0, st>>>(tab, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    prep_rows_kernel<<<max(1, min(tab.n_slices, sm_count * per_sm)), PREP_THREADS, 0, st>>>(tab);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    seg_count_kernel<<<min((tab.W + 3) / 4, sm_count * 8), 128, 0, st>>>(tab, counts);
    return cudaGetLastError();
}

cudaError_t configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(window_pairs_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
    if (e != cudaSuccess) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, window_pairs_tc_kernel<false>) == cudaSuccess)
            fprintf(stderr, "pairs kernel: dynamic shared memory %d B requested, static %zu B, max dynamic %d B\n",
                    WS_SMEM_BYTES, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
        return e;
    }
    return cudaFuncSetAttribute(window_pairs_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
}

cudaError_t launch_pairs(const WindowTab &tab, const ItemParams &prm, int algo, int sm_count, cudaStream_t st) {
    if (prm.item_end <= prm.item_begin) return cudaSuccess;
    int64_t items = (prm.item_end - prm.item_begin - prm.rank + prm.world - 1) / prm.world;
    if (items <= 0) return cudaSuccess;
    if (algo == IMPOP_ALGO_TCGEN05) {
        int grid = (int)(items < sm_count ? items : sm_count);
        if (prm.dumpI || prm.dumpPi) window_pairs_tc_kernel<true><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(tab, prm);
        else window_pairs_tc_kernel<false><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(tab, prm);
    } else {
        int64_t cap = (int64_t)sm_count * 8;
        int grid = (int)(items < cap ? items : cap);
        window_pairs_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(tab, prm);
    }
    return cudaGetLastError();
}

cudaError_t launch_window_sums(const WindowTab &tab, const double *partials, int rank, int world, double *sums,
                               int sm_count, cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    if (tab.W < sm_count * 4) {                       // few windows: a CTA each (their item lists are long when n is large)
        window_sums_block_kernel<<<tab.W, 256, 0, st>>>(tab, partials, rank, world, sums);
        return cudaGetLastError();
    }
    int blocks = min((tab.W + 3) / 4, sm_count * 8);
    window_sums_kernel<<<blocks, 128, 0, st>>>(tab, partials, rank, world, sums);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const WindowTab &tab, const double *sums, int parts, const int64_t *counts, double *stats,
                            int sm_count, cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    finalize_kernel<<<min((tab.W + 127) / 128, sm_count * 4), 128, 0, st>>>(tab, sums, parts, counts, stats);
    return cudaGetLastError();
}

cudaError_t launch_export_a(const int32_t *A, int32_t n, int64_t *out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    export_a_kernel<<<(n + 127) / 128, 128, 0, st>>>(A, n, out);
    return cudaGetLastError();
}

cudaError_t launch_division_selftest(uint64_t seed, int64_t count, unsigned long long *out, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    division_selftest_kernel<<<sm_count * 4, 256, 0, st>>>(seed, count, out);
    return cudaGetLastError();
}

}  // namespace impop
