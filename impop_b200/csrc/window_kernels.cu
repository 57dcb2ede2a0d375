// Windowed pairwise kernels: prep (path lengths, byte weights, heavy-node table), the fused
// pairwise + fp64 reduction kernel in two implementations (tcgen05 / TMEM and dp4a SIMT),
// segregating-node counts, and the per-window finalize.
//
// Replaces, per window: `impg similarity` / `odgi similarity` (reference call sites
// run_pica2_impg.sh:162-168, run_h-fst.sh:65-67, run_tajd.sh:160) + pica2.py:118-164 +
// h-fst.py:130-249 + run_tajd.sh:126-148 (S) + tj_d.py:47-69.
//
// Arithmetic: I_ij = sum_k len_k x_ik x_jk is computed as an unsigned 8-bit GEMM with exact
// int32 accumulation.  Each node k contributes the "dense" column (a = x_ik, b = x_jk * (len_k % 255))
// and, when len_k >= 255, ceil(q / 255) "heavy" columns (a = 255 * x_ik, b = x_jk * c) with the c's
// summing to q = len_k / 255, so a * b summed over a node's columns is exactly len_k.
#include "common.cuh"
#include "stats_math.cuh"

namespace impop {

// ==========================================================================================
// Prep: one CTA per window.
// ==========================================================================================
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS) prep_kernel(WindowTab tab, int32_t *counter) {
    __shared__ int s_heavy;
    __shared__ unsigned long long s_total;
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0;
    for (int w = blockIdx.x; w < tab.W; w += gridDim.x) {
        const int n = tab.n[w], m = tab.m[w], pitch = tab.pitch[w];
        const uint32_t *len = tab.len + tab.len_off[w];
        const uint32_t *x = tab.x + tab.x_off[w];
        uint8_t *w8 = tab.w8 + tab.w8_off[w];
        uint32_t *heavy = tab.heavy + tab.heavy_off[w];
        const int m64 = (int)(tab.w8_off[w + 1] - tab.w8_off[w]);
        const int hpad = (int)(tab.heavy_off[w + 1] - tab.heavy_off[w]);
        if (threadIdx.x == 0) { s_heavy = 0; s_total = 0ull; }
        __syncthreads();
        unsigned long long tot = 0;
        for (int k = threadIdx.x; k < m64; k += PREP_THREADS) {
            uint32_t l = (k < m) ? len[k] : 0u;
            tot += l;
            w8[k] = (uint8_t)(l % HEAVY_Q);
            uint32_t q = l / HEAVY_Q;
            while (q > 0) {
                uint32_t c = q < 255u ? q : 255u;
                int slot = atomicAdd(&s_heavy, 1);
                if (slot < hpad) heavy[slot] = ((uint32_t)k << 8) | c;
                q -= c;
            }
        }
        if (tot) atomicAdd(&s_total, tot);
        __syncthreads();
        if (threadIdx.x == 0 && (s_total >= (1ull << 31) || s_heavy > hpad)) atomicExch(tab.err, (int32_t)DEV_ERR_RANGE);
        for (int s = s_heavy + threadIdx.x; s < hpad; s += PREP_THREADS) heavy[s] = 0u;
        // path lengths: one warp per haplotype
        int32_t *A = tab.A + tab.row_off[w];
        const int words = (m + 31) >> 5;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int i = warp; i < n; i += PREP_THREADS / 32) {
            const uint32_t *row = x + (size_t)i * pitch;
            uint32_t acc = 0;
            for (int wd = lane; wd < words; wd += 32) {
                uint32_t v = __ldg(row + wd);
                while (v) {
                    int k = wd * 32 + (__ffs(v) - 1);
                    if (k < m) acc += __ldg(len + k);
                    v &= v - 1;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) A[i] = (int32_t)acc;
        }
        __syncthreads();
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

// Harmonic tables a1(n), a2(n) (see stats_math.cuh).  Sequential by nature; run once per context.
__global__ void harmonic_table_kernel(double2 *harm, int32_t nmax) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double t1 = 0.0, c1 = 0.0, t2 = 0.0, c2 = 0.0;
    harm[0] = make_double2(0.0, 0.0);
    if (nmax >= 1) harm[1] = make_double2(0.0, 0.0);
    for (int n = 2; n <= nmax; ++n) {  // a(n) sums i = 1 .. n-1
        double di = (double)(n - 1);
        neumaier_add(t1, c1, __ddiv_rn(1.0, di));
        neumaier_add(t2, c2, __ddiv_rn(1.0, __dmul_rn(di, di)));
        harm[n] = make_double2(neumaier_value(t1, c1), neumaier_value(t2, c2));
    }
}

// ==========================================================================================
// Work-item bookkeeping
// ==========================================================================================
struct Item {
    int w, bi, cb0, ncb;
};

__device__ Item decode_item(const WindowTab &tab, int64_t t) {
    int lo = 0, hi = tab.W - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tab.item_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    Item it;
    it.w = lo;
    int64_t r = t - tab.item_off[lo];
    int nb = (tab.n[lo] + TILE_M - 1) / TILE_M;
    int bi = 0;
    while (bi < nb) {
        int u = (nb - bi + 1) >> 1;
        if (r < u) break;
        r -= u;
        ++bi;
    }
    it.bi = bi;
    it.cb0 = bi + 2 * (int)r;
    it.ncb = min(2, nb - it.cb0);
    return it;
}

// Gather 32 presence bits of one haplotype row for 32 heavy-table entries.
__device__ __forceinline__ uint32_t gather_heavy_bits(const uint32_t *row, const uint32_t *entries, bool row_valid) {
    uint32_t bits = 0;
    if (row_valid) {
#pragma unroll 4
        for (int e = 0; e < 32; ++e) {
            uint32_t ent = __ldg(entries + e);
            uint32_t col = ent >> 8;
            uint32_t word = __ldg(row + (col >> 5));
            bits |= ((word >> (col & 31u)) & 1u) << e;
        }
    }
    return bits;
}

// Per-pair epilogue shared by both implementations.  Column class flags are warp-uniform
// (every lane of a warp looks at the same column j), so the branches do not diverge.
// Accumulation is two-level: plain adds into a chunk-local sum (<= 32 columns, terms of similar
// size), then a compensated add of the chunk sum into the running total (see dd in common.cuh).
struct PairAcc {
    double s, a, b;  // sums over columns carrying SUBSET / A / B, for this thread's row
};
struct PairTot {
    dd s, a, b;
};

__device__ __forceinline__ void pair_step(PairAcc &acc, uint32_t inter, uint32_t ai, uint32_t aj, uint32_t fj,
                                          bool valid) {
    if (fj == 0u) return;
    double p = valid ? pi_from_counts(inter, ai, aj) : 0.0;
    if (fj & IMPOP_LAB_SUBSET) acc.s = __dadd_rn(acc.s, p);
    if (fj & IMPOP_LAB_A) acc.a = __dadd_rn(acc.a, p);
    if (fj & IMPOP_LAB_B) acc.b = __dadd_rn(acc.b, p);
}

__device__ __forceinline__ void pair_fold(PairTot &tot, PairAcc &acc) {
    dd_add(tot.s, acc.s); dd_add(tot.a, acc.a); dd_add(tot.b, acc.b);
    acc.s = 0.0; acc.a = 0.0; acc.b = 0.0;
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

// Row-side combination + block reduction into partials[item][4] (fixed order => deterministic).
template <int NWARPS>
__device__ __forceinline__ void item_reduce(const PairTot &tot, uint32_t fi, dd (*s_red)[4], double *out4) {
    const dd zero = {0.0, 0.0};
    dd v[4];
    v[0] = (fi & IMPOP_LAB_SUBSET) ? tot.s : zero;
    v[1] = (fi & IMPOP_LAB_A) ? tot.a : zero;
    v[2] = (fi & IMPOP_LAB_B) ? tot.b : zero;
    v[3] = (fi & IMPOP_LAB_A) ? tot.b : zero;
    if (fi & IMPOP_LAB_B) dd_merge(v[3], tot.a);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = warp_sum_dd(v[k]);
        if (lane == 0) s_red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        dd t = s_red[0][threadIdx.x];
#pragma unroll
        for (int wgt = 1; wgt < NWARPS; ++wgt) dd_merge(t, s_red[wgt][threadIdx.x]);
        out4[threadIdx.x] = t.hi;          // partials are stored as (hi[4], lo[4])
        out4[4 + threadIdx.x] = t.lo;
    }
}

// ==========================================================================================
// tcgen05 implementation.  256 threads, 2 CTAs per SM (each owns 256 TMEM columns), so one
// CTA's fp64 epilogue overlaps the other's operand expansion + MMA.
// ==========================================================================================
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 4;
constexpr int A_STAGE_BYTES = TILE_M * KCHUNK;                  // 8 KB
constexpr int B_STAGE_BYTES = TILE_N * KCHUNK;                  // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;      // 24 KB
constexpr uint32_t LBO_A = TILE_M * 16;                         // next 16-byte K slab of the A tile
constexpr uint32_t LBO_B = TILE_N * 16;
constexpr uint32_t SBO_AB = 128;                                // next 8-row group
constexpr uint32_t TMEM_COLS = 256;

struct TcShared {
    uint64_t stage_free[TC_STAGES];
    uint64_t acc_full;
    uint32_t tmem_base;
    int32_t item_w, item_bi, item_cb0, item_ncb;
    long long item_id;
    int32_t a_col[TILE_N];
    uint8_t f_col[TILE_N];
    dd red[TC_THREADS / 32][4];
};
constexpr int TC_SMEM_BYTES = TC_STAGES * STAGE_BYTES + (int)sizeof(TcShared) + 1024;

__global__ void __launch_bounds__(TC_THREADS, 2) window_pairs_tc_kernel(WindowTab tab, ItemParams prm) {
    extern __shared__ uint8_t smem_raw[];
    // operand stages need 128-byte alignment (16-byte core-matrix rows); align generously
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    TcShared &sh = *reinterpret_cast<TcShared *>(smem + TC_STAGES * STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&sh.stage_free[s], 1);
        mbar_init(&sh.acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&sh.tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh.tmem_base;
    const uint32_t smem_base_u32 = smem_u32(smem);

    uint32_t gchunk = 0;      // chunks issued so far by this CTA (stage = gchunk % STAGES)
    uint32_t acc_uses = 0;    // completed phases of acc_full (items with at least one MMA)
    bool alive = true;

    while (true) {
        if (tid == 0) {
            int64_t t = prm.item_begin + ((int64_t)atomicAdd(prm.counter, 1) * prm.world + prm.rank);
            if (t < prm.item_end) {
                Item it = decode_item(tab, t);
                sh.item_w = it.w; sh.item_bi = it.bi; sh.item_cb0 = it.cb0; sh.item_ncb = it.ncb;
            } else {
                sh.item_w = -1;
            }
            sh.item_id = (long long)t;
        }
        __syncthreads();
        const int w = sh.item_w;
        if (w < 0) break;
        const int64_t item_id = sh.item_id;
        const int bi = sh.item_bi, cb0 = sh.item_cb0, ncb = sh.item_ncb;
        const int n = tab.n[w], pitch = tab.pitch[w];
        const uint32_t *x = tab.x + tab.x_off[w];
        const uint8_t *w8 = tab.w8 + tab.w8_off[w];
        const uint32_t *heavy = tab.heavy + tab.heavy_off[w];
        const int32_t *Aw = tab.A + tab.row_off[w];
        const uint8_t *lab = tab.labels + tab.lab_off[w];
        const int dense_chunks = (int)((tab.w8_off[w + 1] - tab.w8_off[w]) / KCHUNK);
        const int heavy_chunks = (int)((tab.heavy_off[w + 1] - tab.heavy_off[w]) / KCHUNK);
        const int nch = dense_chunks + heavy_chunks;
        const int ncols = ncb * TILE_M;
        const uint32_t idesc = make_idesc_u8(TILE_M, (uint32_t)ncols);

        {   // column-side path lengths and labels for the epilogue
            int j = cb0 * TILE_M + tid;
            bool ok = tid < ncols && j < n;
            sh.a_col[tid] = ok ? Aw[j] : 0;
            sh.f_col[tid] = ok ? (uint8_t)clean_label(lab[j]) : (uint8_t)0;
        }
        __syncthreads();

        // ------------------------------------------------------------------ K loop
        for (int c = 0; c < nch; ++c) {
            const uint32_t s = gchunk % TC_STAGES, use = gchunk / TC_STAGES;
            if (use >= 1 && alive) alive = mbar_wait(&sh.stage_free[s], (use - 1) & 1u, tab.err);
            uint8_t *stA = smem + s * STAGE_BYTES;
            uint8_t *stB = stA + A_STAGE_BYTES;
            const bool is_heavy = c >= dense_chunks;
            const int hc = c - dense_chunks;
            // 24 warp tasks: (operand, 32-row group, 32-column half)
            const int b_groups = ncb * 4;
            for (int task = warp; task < 8 + 2 * b_groups; task += TC_THREADS / 32) {
                const bool isA = task < 8;
                const int tt = isA ? task : task - 8;
                const int rg = tt >> 1, half = tt & 1;
                const int rl = rg * 32 + lane;
                const int grow = (isA ? bi : cb0) * TILE_M + rl;
                const bool rvalid = grow < n;
                const uint32_t *row = x + (size_t)grow * pitch;
                uint32_t bits = 0;
                uint32_t wv[8];
                if (!is_heavy) {
                    const int wd = c * 2 + half;
                    if (rvalid && wd < pitch) bits = __ldg(row + wd);
                    if (!isA) {
                        const uint4 *wp = reinterpret_cast<const uint4 *>(w8 + c * KCHUNK + half * 32);
                        uint4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                        wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                        wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
                    }
                } else {
                    const uint32_t *ent = heavy + hc * KCHUNK + half * 32;
                    bits = gather_heavy_bits(row, ent, rvalid);
                    if (!isA) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            wv[q] = (__ldg(ent + 4 * q) & 255u) | ((__ldg(ent + 4 * q + 1) & 255u) << 8) |
                                    ((__ldg(ent + 4 * q + 2) & 255u) << 16) | ((__ldg(ent + 4 * q + 3) & 255u) << 24);
                        }
                    }
                }
                uint32_t out[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t b01 = nibble_to_bytes01((bits >> (4 * q)) & 0xFu);
                    if (isA) out[q] = is_heavy ? b01 * 255u : b01;
                    else out[q] = (b01 * 255u) & wv[q];
                }
                uint8_t *dst = (isA ? stA : stB) + (uint32_t)(half * 2) * (isA ? LBO_A : LBO_B) + rl * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<uint4 *>(dst + (isA ? LBO_A : LBO_B)) = make_uint4(out[4], out[5], out[6], out[7]);
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_addr = smem_base_u32 + s * STAGE_BYTES;
                const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                for (int k32 = 0; k32 < KCHUNK / 32; ++k32) {
                    uint64_t da = make_smem_desc(a_addr + k32 * 2 * LBO_A, LBO_A, SBO_AB);
                    uint64_t db = make_smem_desc(b_addr + k32 * 2 * LBO_B, LBO_B, SBO_AB);
                    tc_mma_i8(tmem_base, da, db, idesc, (c > 0 || k32 > 0) ? 1u : 0u);
                }
                tc_commit(&sh.stage_free[s]);
                if (c == nch - 1) tc_commit(&sh.acc_full);
            }
            ++gchunk;
        }

        // ------------------------------------------------------------------ epilogue
        if (nch > 0 && alive) alive = mbar_wait(&sh.acc_full, acc_uses & 1u, tab.err);
        if (nch > 0) ++acc_uses;
        tc_fence_after();
        const int q4 = warp & 3, hsel = warp >> 2;
        const int i = bi * TILE_M + q4 * 32 + lane;
        const bool rvalid = i < n;
        const uint32_t ai = rvalid ? (uint32_t)Aw[i] : 0u;
        const uint32_t fi = rvalid ? clean_label(lab[i]) : 0u;
        const int half_cols = ncols / 2;
        const int warp_row_min = bi * TILE_M + q4 * 32;
        const bool dump = (prm.dumpI != nullptr) || (prm.dumpPi != nullptr);
        PairAcc acc = {0.0, 0.0, 0.0};
        PairTot tot = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        for (int cc = hsel * half_cols; cc < (hsel + 1) * half_cols; cc += 16) {
            const int jbase = cb0 * TILE_M + cc;
            if (jbase >= n) break;
            if (!dump && jbase + 15 <= warp_row_min) continue;  // every j <= every i of this warp
            uint32_t r[16];
            if (nch > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)cc, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int t = 0; t < 16; ++t) r[t] = 0u;
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int j = jbase + t;
                const uint32_t aj = (uint32_t)sh.a_col[cc + t];
                const uint32_t fj = sh.f_col[cc + t];
                pair_step(acc, r[t], ai, aj, fj, rvalid && j < n && j > i);
                if (dump) pair_dump(prm, n, i, j, r[t], ai, aj);
            }
            pair_fold(tot, acc);
        }
        tc_fence_before();
        __syncthreads();   // all TMEM reads done before the next item's first MMA overwrites the accumulator
        item_reduce<TC_THREADS / 32>(tot, fi, sh.red, prm.partials + item_id * 8);
        __syncthreads();
    }

    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ==========================================================================================
// SIMT implementation (dp4a).  128 threads = the 128 rows of the item's row block; columns are
// processed 32 at a time with u32 accumulators in registers.  Same items, same epilogue.
// ==========================================================================================
constexpr int SIMT_THREADS = 128;
constexpr int SIMT_COLS = 32;

__global__ void __launch_bounds__(SIMT_THREADS) window_pairs_simt_kernel(WindowTab tab, ItemParams prm) {
    __shared__ __align__(16) uint32_t s_b[SIMT_COLS][KCHUNK / 4 + 4];  // +4 words: rows stay 16-byte aligned
    __shared__ int32_t s_item[4];
    __shared__ long long s_item_id;
    __shared__ dd s_red[SIMT_THREADS / 32][4];
    const int tid = threadIdx.x;

    while (true) {
        if (tid == 0) {
            int64_t t = prm.item_begin + ((int64_t)atomicAdd(prm.counter, 1) * prm.world + prm.rank);
            s_item_id = t;
            if (t < prm.item_end) {
                Item it = decode_item(tab, t);
                s_item[0] = it.w; s_item[1] = it.bi; s_item[2] = it.cb0; s_item[3] = it.ncb;
            } else {
                s_item[0] = -1;
            }
        }
        __syncthreads();
        const int w = s_item[0];
        if (w < 0) break;
        const int64_t item_id = s_item_id;
        const int bi = s_item[1], cb0 = s_item[2], ncb = s_item[3];
        const int n = tab.n[w], pitch = tab.pitch[w];
        const uint32_t *x = tab.x + tab.x_off[w];
        const uint8_t *w8 = tab.w8 + tab.w8_off[w];
        const uint32_t *heavy = tab.heavy + tab.heavy_off[w];
        const int32_t *Aw = tab.A + tab.row_off[w];
        const uint8_t *lab = tab.labels + tab.lab_off[w];
        const int dense_chunks = (int)((tab.w8_off[w + 1] - tab.w8_off[w]) / KCHUNK);
        const int heavy_chunks = (int)((tab.heavy_off[w + 1] - tab.heavy_off[w]) / KCHUNK);
        const int nch = dense_chunks + heavy_chunks;
        const int i = bi * TILE_M + tid;
        const bool rvalid = i < n;
        const uint32_t *myrow = x + (size_t)i * pitch;
        const uint32_t ai = rvalid ? (uint32_t)Aw[i] : 0u;
        const uint32_t fi = rvalid ? clean_label(lab[i]) : 0u;
        const bool dump = (prm.dumpI != nullptr) || (prm.dumpPi != nullptr);
        PairAcc acc = {0.0, 0.0, 0.0};
        PairTot tot = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};

        for (int sub = 0; sub < ncb * (TILE_M / SIMT_COLS); ++sub) {
            const int jbase = cb0 * TILE_M + sub * SIMT_COLS;
            if (jbase >= n) break;
            if (!dump && jbase + SIMT_COLS - 1 <= bi * TILE_M) continue;
            uint32_t cnt[SIMT_COLS];
#pragma unroll
            for (int j = 0; j < SIMT_COLS; ++j) cnt[j] = 0u;
            for (int c = 0; c < nch; ++c) {
                const bool is_heavy = c >= dense_chunks;
                const int hc = c - dense_chunks;
                __syncthreads();
                {   // stage B': thread -> (column row = tid / 4, 16-byte slab = tid % 4)
                    const int jr = tid >> 2, slab = tid & 3;
                    const int gj = jbase + jr;
                    const bool jvalid = gj < n;
                    const uint32_t *jrow = x + (size_t)gj * pitch;
                    uint32_t bits16 = 0, wv[4];
                    if (!is_heavy) {
                        const int wd = c * 2 + (slab >> 1);
                        if (jvalid && wd < pitch) bits16 = (__ldg(jrow + wd) >> ((slab & 1) * 16)) & 0xFFFFu;
                        uint4 wq = __ldg(reinterpret_cast<const uint4 *>(w8 + c * KCHUNK + slab * 16));
                        wv[0] = wq.x; wv[1] = wq.y; wv[2] = wq.z; wv[3] = wq.w;
                    } else {
                        const uint32_t *ent = heavy + hc * KCHUNK + slab * 16;
#pragma unroll
                        for (int q = 0; q < 4; ++q) wv[q] = 0u;
                        for (int e = 0; e < 16; ++e) {
                            uint32_t en = __ldg(ent + e);
                            uint32_t col = en >> 8;
                            if (jvalid) bits16 |= ((__ldg(jrow + (col >> 5)) >> (col & 31u)) & 1u) << e;
                            wv[e >> 2] |= (en & 255u) << ((e & 3) * 8);
                        }
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) o[q] = (nibble_to_bytes01((bits16 >> (4 * q)) & 0xFu) * 255u) & wv[q];
                    *reinterpret_cast<uint4 *>(&s_b[jr][slab * 4]) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                // own A' words for this chunk
                uint32_t a[KCHUNK / 4];
                {
                    uint32_t b0 = 0, b1 = 0;
                    if (!is_heavy) {
                        if (rvalid && c * 2 < pitch) b0 = __ldg(myrow + c * 2);
                        if (rvalid && c * 2 + 1 < pitch) b1 = __ldg(myrow + c * 2 + 1);
                    } else {
                        b0 = gather_heavy_bits(myrow, heavy + hc * KCHUNK, rvalid);
                        b1 = gather_heavy_bits(myrow, heavy + hc * KCHUNK + 32, rvalid);
                    }
                    const uint32_t mul = is_heavy ? 255u : 1u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        a[q] = nibble_to_bytes01((b0 >> (4 * q)) & 0xFu) * mul;
                        a[8 + q] = nibble_to_bytes01((b1 >> (4 * q)) & 0xFu) * mul;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < SIMT_COLS; ++j) {
#pragma unroll
                    for (int q4 = 0; q4 < KCHUNK / 16; ++q4) {
                        uint4 b = *reinterpret_cast<const uint4 *>(&s_b[j][q4 * 4]);
                        cnt[j] = __dp4a(a[q4 * 4 + 0], b.x, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 1], b.y, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 2], b.z, cnt[j]);
                        cnt[j] = __dp4a(a[q4 * 4 + 3], b.w, cnt[j]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < SIMT_COLS; ++t) {
                const int j = jbase + t;
                const bool jv = j < n;
                const uint32_t aj = jv ? (uint32_t)__ldg(Aw + j) : 0u;
                const uint32_t fj = jv ? clean_label(__ldg(lab + j)) : 0u;
                pair_step(acc, cnt[t], ai, aj, fj, rvalid && jv && j > i);
                if (dump) pair_dump(prm, n, i, j, cnt[t], ai, aj);
            }
            pair_fold(tot, acc);
        }
        __syncthreads();
        item_reduce<SIMT_THREADS / 32>(tot, fi, s_red, prm.partials + item_id * 8);
        __syncthreads();
    }
}

// ==========================================================================================
// Segregating nodes + label counts: one CTA per window.  counts row = nS nA nB pS pAA pBB pAB S.
// ==========================================================================================
constexpr int COL_THREADS = 128;

__global__ void __launch_bounds__(COL_THREADS) colstat_kernel(WindowTab tab, int64_t *counts) {
    __shared__ int s_cnt[4];
    for (int w = blockIdx.x; w < tab.W; w += gridDim.x) {
        const int n = tab.n[w], m = tab.m[w], pitch = tab.pitch[w];
        const uint32_t *x = tab.x + tab.x_off[w];
        const uint32_t *len = tab.len + tab.len_off[w];
        const uint8_t *lab = tab.labels + tab.lab_off[w];
        if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int c0 = 0, c1 = 0, c2 = 0;
        for (int i = threadIdx.x; i < n; i += COL_THREADS) {
            uint32_t f = clean_label(lab[i]);
            c0 += (f & IMPOP_LAB_SUBSET) != 0; c1 += (f & IMPOP_LAB_A) != 0; c2 += (f & IMPOP_LAB_B) != 0;
        }
        if (c0) atomicAdd(&s_cnt[0], c0);
        if (c1) atomicAdd(&s_cnt[1], c1);
        if (c2) atomicAdd(&s_cnt[2], c2);
        const int words = (m + 31) >> 5;
        int seg = 0;
        for (int wd = threadIdx.x; wd < words; wd += COL_THREADS) {
            uint32_t any = 0u, all = 0xffffffffu;
            int rows = 0;
            for (int i = 0; i < n; ++i) {
                if (!(lab[i] & IMPOP_LAB_SEG)) continue;   // uniform across the CTA
                uint32_t v = __ldg(x + (size_t)i * pitch + wd);
                any |= v; all &= v; ++rows;
            }
            uint32_t sg = rows ? (any & ~all) : 0u;
            while (sg) {
                int k = wd * 32 + (__ffs(sg) - 1);
                if (k < m && __ldg(len + k) > 0u) ++seg;
                sg &= sg - 1;
            }
        }
        if (seg) atomicAdd(&s_cnt[3], seg);
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t nS = s_cnt[0], nA = s_cnt[1], nB = s_cnt[2];
            int64_t *row = counts + (size_t)w * IMPOP_NCOUNTS;
            row[0] = nS; row[1] = nA; row[2] = nB;
            row[3] = nS * (nS - 1) / 2; row[4] = nA * (nA - 1) / 2; row[5] = nB * (nB - 1) / 2; row[6] = nA * nB;
            row[7] = s_cnt[3];
        }
        __syncthreads();
    }
}

// ==========================================================================================
// Window sums (fixed-order reduction of item partials) and finalize.
// ==========================================================================================
__global__ void window_sums_kernel(WindowTab tab, const double *partials, int32_t rank, int32_t world, double *sums) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < tab.W; w += gridDim.x * warps_per_block) {
        const int64_t t0 = tab.item_off[w], t1 = tab.item_off[w + 1];
        dd v[4] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        // first item of this window handled by `rank`
        int64_t first = t0 + ((rank - (t0 % world)) % world + world) % world;
        for (int64_t t = first + (int64_t)lane * world; t < t1; t += 32ll * world) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dd p = {partials[t * 8 + k], partials[t * 8 + 4 + k]};
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

__global__ void finalize_kernel(WindowTab tab, const double *sums, int32_t parts, const int64_t *counts, double *stats) {
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
    }
}

// A (int32 scratch) -> caller's int64 array for one window.
__global__ void export_a_kernel(const int32_t *A, int32_t n, int64_t *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)A[i];
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

cudaError_t launch_harmonic_table(double2 *harm, int32_t nmax, cudaStream_t st) {
    harmonic_table_kernel<<<1, 32, 0, st>>>(harm, nmax);
    return cudaGetLastError();
}

cudaError_t launch_prep(const WindowTab &tab, int32_t *counter, cudaStream_t st) {
    prep_kernel<<<max(1, min(tab.W, 148 * 8)), PREP_THREADS, 0, st>>>(tab, counter);
    return cudaGetLastError();
}

cudaError_t configure_kernels() {
    return cudaFuncSetAttribute(window_pairs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
}

cudaError_t launch_pairs(const WindowTab &tab, const ItemParams &prm, int algo, int sm_count, cudaStream_t st) {
    if (prm.item_end <= prm.item_begin) return cudaSuccess;
    int64_t items = (prm.item_end - prm.item_begin + prm.world - 1) / prm.world;
    if (algo == IMPOP_ALGO_TCGEN05) {
        int64_t cap = (int64_t)sm_count * 2;
        int grid = (int)(items < cap ? items : cap);
        window_pairs_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tab, prm);
    } else {
        int64_t cap = (int64_t)sm_count * 8;
        int grid = (int)(items < cap ? items : cap);
        window_pairs_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(tab, prm);
    }
    return cudaGetLastError();
}

cudaError_t launch_colstat(const WindowTab &tab, int64_t *counts, cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    colstat_kernel<<<min(tab.W, 148 * 16), COL_THREADS, 0, st>>>(tab, counts);
    return cudaGetLastError();
}

cudaError_t launch_window_sums(const WindowTab &tab, const double *partials, int rank, int world, double *sums,
                               cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    int blocks = min((tab.W + 3) / 4, 148 * 8);
    window_sums_kernel<<<blocks, 128, 0, st>>>(tab, partials, rank, world, sums);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const WindowTab &tab, const double *sums, int parts, const int64_t *counts, double *stats,
                            cudaStream_t st) {
    if (tab.W == 0) return cudaSuccess;
    finalize_kernel<<<min((tab.W + 127) / 128, 148 * 4), 128, 0, st>>>(tab, sums, parts, counts, stats);
    return cudaGetLastError();
}

cudaError_t launch_export_a(const int32_t *A, int32_t n, int64_t *out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    export_a_kernel<<<(n + 127) / 128, 128, 0, st>>>(A, n, out);
    return cudaGetLastError();
}

}  // namespace impop
