// C ABI of libimpop_b200.so (see include/impop_b200.h).  Host-side plumbing only: argument
// checks, device tables, scratch ownership and kernel launches on the caller's stream.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>
#include <time.h>

#include <algorithm>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace impop {
cudaError_t launch_heavy_count(const uint32_t *, const int64_t *, const int32_t *, int32_t, int32_t *, cudaStream_t);
int prep_rows_ctas_per_sm();
cudaError_t launch_prep(const WindowTab &, int64_t *, int, int, cudaStream_t);
cudaError_t launch_division_selftest(uint64_t, int64_t, unsigned long long *, int, cudaStream_t);
cudaError_t configure_kernels();
cudaError_t launch_pairs(const WindowTab &, const ItemParams &, int, int, cudaStream_t);
cudaError_t launch_window_sums(const WindowTab &, const double *, int, int, double *, int, cudaStream_t);
cudaError_t launch_finalize(const WindowTab &, const double *, int, const int64_t *, double *, int, cudaStream_t);
cudaError_t launch_export_a(const int32_t *, int32_t, int64_t *, cudaStream_t);
cudaError_t launch_pack_bits(const uint8_t *, int32_t, int32_t, int64_t, uint32_t *, int32_t, int, cudaStream_t);
cudaError_t launch_repitch_rows(const RepitchDesc *, int32_t, const uint32_t *, uint32_t *, int, cudaStream_t);
int reduce_identity_blocks(int32_t n, int sm_count);
cudaError_t launch_reduce_identity(const double *, int32_t, int64_t, const uint8_t *, const double *, int64_t, double,
                                   const double2 *, int32_t, double *, int, double *, int64_t *, double *, cudaStream_t);
cudaError_t launch_tajima(const int64_t *, const double *, const double *, int32_t, double *, double *, cudaStream_t);
cudaError_t launch_site_counts(const uint64_t *, int64_t, int32_t, const uint64_t *, int32_t, int32_t *, double *, int,
                               cudaStream_t);
cudaError_t launch_cluster(const double *, int32_t, int64_t, double, int32_t *, int32_t *, int, cudaStream_t);
cudaError_t launch_greedy_groups(const double *, int32_t, int64_t, double, int32_t *, double *, cudaStream_t);
cudaError_t launch_round_decimal(double *, int64_t, int32_t, int, cudaStream_t);
}  // namespace impop

using namespace impop;

constexpr int32_t HARM_N = 1 << 16;

struct impop_ctx {
    int device = 0;
    int sm_count = 0;
    int prep_per_sm = 0;              // resident prep_rows CTAs per SM on this device (queried at the first window batch)
    bool kernels_configured = false;  // shared-memory attribute of the pairs kernel set (first window batch)
    int32_t *err_dev = nullptr;
    double2 *harm_dev = nullptr;
    long long *prof_dev = nullptr;    // role-time counters of the last pairs launch (IMPOP_PROFILE_ROLES builds)
    int32_t *cluster_parent = nullptr;
    int32_t cluster_cap = 0;
    int64_t launches = 0;
    std::string last_error;
    // optional per-kernel timing (impop_timing_enable): event pairs recorded on the caller's stream
    bool timing = false;
    struct Slot { cudaEvent_t a, b; int kid; };
    std::vector<Slot> slots;       // recorded launches since the last enable/read
    std::vector<Slot> spare;       // recycled event pairs
    // device scratch pool and pinned staging buffers, recycled across batches (no cudaMalloc / cudaFree,
    // which synchronise the device, on the per-batch path)
    // A block handed back while work on `st` may still use it carries an event: it is reused at once by work on the same
    // stream (ordered behind it), by another stream only after the event has completed.
    struct Block { void *p; size_t cap; bool used; cudaEvent_t ev; cudaStream_t st; bool pending; };
    std::vector<Block> pool;
    struct Staging { void *host; size_t cap; cudaEvent_t done; bool busy; };
    std::vector<Staging> staging;
};

static void *pool_get(impop_ctx *ctx, size_t bytes, cudaStream_t st) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    int best = -1;
    for (int k = 0; k < (int)ctx->pool.size(); ++k) {
        auto &b = ctx->pool[k];
        if (b.used || b.cap < bytes || b.cap > 2 * bytes + (1u << 20)) continue;
        if (b.pending && b.st != st) {                       // still (possibly) in use by another stream
            if (cudaEventQuery(b.ev) != cudaSuccess) { cudaGetLastError(); continue; }
            b.pending = false;
        }
        if (best < 0 || b.cap < ctx->pool[best].cap) best = k;
    }
    if (best >= 0) { ctx->pool[best].used = true; return ctx->pool[best].p; }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaDeviceSynchronize();                             // every pending block is free after this
        for (auto it = ctx->pool.begin(); it != ctx->pool.end();) {     // give unused blocks back and retry
            if (!it->used) { cudaFree(it->p); if (it->ev) cudaEventDestroy(it->ev); it = ctx->pool.erase(it); } else ++it;
        }
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    ctx->pool.push_back({p, bytes, true, nullptr, nullptr, false});
    return p;
}

static void pool_put(impop_ctx *ctx, void *p, cudaStream_t st) {
    for (auto &b : ctx->pool)
        if (b.p == p) {
            b.used = false;
            if (!b.ev && cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); b.ev = nullptr; }
            if (b.ev && cudaEventRecord(b.ev, st) == cudaSuccess) { b.pending = true; b.st = st; }
            else { cudaGetLastError(); cudaStreamSynchronize(st); b.pending = false; }
            return;
        }
}

// A pinned buffer whose previous asynchronous copy has completed.
static impop_ctx::Staging *staging_get(impop_ctx *ctx, size_t bytes) {
    for (auto &s : ctx->staging) {
        if (s.cap < bytes) continue;
        if (s.busy && cudaEventQuery(s.done) != cudaSuccess) { cudaGetLastError(); continue; }
        s.busy = false;
        return &s;
    }
    impop_ctx::Staging s{nullptr, bytes + (bytes >> 2) + 4096, nullptr, false};
    if (cudaMallocHost(&s.host, s.cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) { cudaFreeHost(s.host); return nullptr; }
    ctx->staging.push_back(s);
    return &ctx->staging.back();
}

constexpr size_t MAX_TIMING_SLOTS = 1 << 14;

// Launch `fn` on `st`, bracketed by events when timing is on.  Events add no synchronisation.
template <typename F>
static cudaError_t timed(impop_ctx *ctx, int kid, cudaStream_t st, F fn) {
    if (!ctx->timing || ctx->slots.size() >= MAX_TIMING_SLOTS) return fn();
    impop_ctx::Slot s;
    if (!ctx->spare.empty()) { s = ctx->spare.back(); ctx->spare.pop_back(); }
    else {
        cudaError_t e = cudaEventCreate(&s.a);
        if (e != cudaSuccess) return e;
        e = cudaEventCreate(&s.b);
        if (e != cudaSuccess) return e;
    }
    s.kid = kid;
    cudaEventRecord(s.a, st);
    cudaError_t e = fn();
    cudaEventRecord(s.b, st);
    ctx->slots.push_back(s);
    return e;
}

struct impop_batch {
    WindowTab tab{};
    void *tables = nullptr;           // pooled device blocks: descriptor tables / scratch
    void *scratch = nullptr;
    std::vector<int64_t> item_off;    // host copy
    std::vector<int32_t> n;
    int64_t items = 0;
    cudaStream_t last_stream = nullptr;   // stream of the most recent work on this batch (its blocks go back to the pool behind it)
    double *partials = nullptr;
    double *sums_tmp = nullptr;
    int64_t *counts_tmp = nullptr;
};

static int fail(impop_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

static int cuda_fail(impop_ctx *ctx, cudaError_t e, const char *where) {
    return fail(ctx, IMPOP_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

static int64_t items_of(int32_t n) {
    int64_t nb = (n + TILE_M - 1) / TILE_M, t = 0;
    for (int64_t bi = 0; bi < nb; ++bi) t += items_of_rowblock(n, (int)bi);
    return t;
}

extern "C" {

int impop_version(void) { return 100; }

int impop_create(int device, impop_ctx_t **ctx_out) {
    if (!ctx_out) return IMPOP_ERR_ARG;
    *ctx_out = nullptr;
    int count = 0;
    struct timespec ts_enter; clock_gettime(CLOCK_MONOTONIC, &ts_enter);
    const double t_enter = ts_enter.tv_sec + 1e-9 * ts_enter.tv_nsec;
    cudaError_t e0 = cudaGetDeviceCount(&count);
    if (e0 != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        fprintf(stderr, "impop_create: no usable CUDA device (%s, %d devices)\n", cudaGetErrorString(e0), count);
        return IMPOP_ERR_CUDA;
    }
    impop_ctx *ctx = new (std::nothrow) impop_ctx();
    if (!ctx) return IMPOP_ERR_NOMEM;
    ctx->device = device;
    const bool trace = getenv("IMPOP_TRACE_CREATE") != nullptr;       // where context creation spends its time (stderr)
    auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t_start = now();
    if (trace) fprintf(stderr, "impop_create: cudaGetDeviceCount (driver initialisation) took %.3f s\n", t_start - t_enter);
    cudaDeviceProp prop;
    if ((e0 = cudaSetDevice(device)) != cudaSuccess || (e0 = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        fprintf(stderr, "impop_create: %s\n", cudaGetErrorString(e0));
        delete ctx;
        return IMPOP_ERR_CUDA;
    }
    if (prop.major != 10) {   // sm_100a cubin only: fail loudly rather than fall back
        fprintf(stderr, "impop_create: device %d is sm_%d%d, this library is built for sm_100a only\n", device, prop.major, prop.minor);
        delete ctx;
        return IMPOP_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    // Harmonic tables a1(n), a2(n) exactly as Python >= 3.12 builtin sum() (Neumaier) forms them in tj_d.py:41-45: computed
    // on the host (IEEE double, no contraction: the same bits as the device's _rn intrinsics) and uploaded -- a serial
    // 65 536-step loop is 6.6 ms as a one-thread kernel and 0.3 ms here, on every context creation of every CLI call.
    std::vector<double2> harm((size_t)HARM_N + 1);
    {
        volatile double t1 = 0.0, c1 = 0.0, t2 = 0.0, c2 = 0.0;
        auto nadd = [](volatile double &total, volatile double &comp, double x) {
            const double tot = total;
            const volatile double t = tot + x;
            if (fabs(tot) >= fabs(x)) { const volatile double u = tot - t; const volatile double w = u + x; comp = comp + w; }
            else { const volatile double u = x - t; const volatile double w = u + tot; comp = comp + w; }
            total = t;
        };
        auto nval = [](double total, double comp) { return (comp != 0.0 && std::isfinite(comp)) ? total + comp : total; };
        harm[0] = make_double2(0.0, 0.0);
        harm[1] = make_double2(0.0, 0.0);
        for (int n = 2; n <= HARM_N; ++n) {      // a(n) sums i = 1 .. n-1
            const double di = (double)(n - 1);
            const volatile double sq = di * di;
            nadd(t1, c1, 1.0 / di);
            nadd(t2, c2, 1.0 / sq);
            harm[n] = make_double2(nval(t1, c1), nval(t2, c2));
        }
    }
    const char *step = "";
    cudaError_t es = cudaSuccess;
    auto run = [&](const char *what, cudaError_t r) { if (es == cudaSuccess && r != cudaSuccess) { es = r; step = what; } return es == cudaSuccess; };
    bool ok = run("alloc err flag", cudaMalloc(&ctx->err_dev, sizeof(int32_t))) &&
              run("memset", cudaMemset(ctx->err_dev, 0, sizeof(int32_t))) &&
              run("alloc harmonic table", cudaMalloc(&ctx->harm_dev, sizeof(double2) * (HARM_N + 1))) &&
              run("upload harmonic table", cudaMemcpy(ctx->harm_dev, harm.data(), sizeof(double2) * (HARM_N + 1), cudaMemcpyHostToDevice)) &&
              run("alloc counters", cudaMalloc(&ctx->prof_dev, sizeof(long long) * 16 * 1024)) &&
              run("memset", cudaMemset(ctx->prof_dev, 0, sizeof(long long) * 16 * 1024));
    // (the pairs kernel's shared-memory attribute and the prep occupancy query load those kernels' module: done at the first
    // window batch, so that the TSV-mode command lines -- one process per window in the reference's wrappers -- do not pay it)
    if (!ok) {
        fprintf(stderr, "impop_create: CUDA set-up failed at '%s': %s\n", step, cudaGetErrorString(es));
        cudaFree(ctx->err_dev); cudaFree(ctx->harm_dev); cudaFree(ctx->prof_dev);
        delete ctx;
        return IMPOP_ERR_CUDA;
    }
    if (trace) fprintf(stderr, "impop_create: device context, allocations and table upload took %.3f s\n", now() - t_start);
    ctx->launches = 0;
    *ctx_out = ctx;
    return IMPOP_OK;
}

int impop_destroy(impop_ctx_t *ctx) {
    if (!ctx) return IMPOP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->err_dev); cudaFree(ctx->harm_dev); cudaFree(ctx->prof_dev); cudaFree(ctx->cluster_parent);
    for (auto *v : {&ctx->slots, &ctx->spare})
        for (auto &s : *v) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto &b : ctx->pool) { cudaFree(b.p); if (b.ev) cudaEventDestroy(b.ev); }
    for (auto &s : ctx->staging) { cudaFreeHost(s.host); cudaEventDestroy(s.done); }
    delete ctx;
    return IMPOP_OK;
}

int impop_timing_enable(impop_ctx_t *ctx, int32_t enable) {
    if (!ctx) return IMPOP_ERR_ARG;
    for (auto &s : ctx->slots) ctx->spare.push_back(s);
    ctx->slots.clear();
    ctx->timing = enable != 0;
    return IMPOP_OK;
}

int impop_timing_read(impop_ctx_t *ctx, int32_t kernel_id, double *total_ms, int64_t *launches) {
    if (!ctx || !total_ms || !launches) return IMPOP_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    double tot = 0.0;
    int64_t cnt = 0;
    for (auto &s : ctx->slots) {
        if (s.kid != kernel_id) continue;
        CU(cudaEventSynchronize(s.b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.a, s.b));
        tot += ms;
        ++cnt;
    }
    *total_ms = tot;
    *launches = cnt;
    return IMPOP_OK;
}

const char *impop_last_error(impop_ctx_t *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int64_t impop_launch_count(impop_ctx_t *ctx) { return ctx ? ctx->launches : 0; }

int impop_check(impop_ctx_t *ctx, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    int32_t flag = 0;
    CU(cudaMemcpy(&flag, ctx->err_dev, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag != DEV_OK) {
        cudaMemset(ctx->err_dev, 0, sizeof(int32_t));
        if (flag == DEV_ERR_RANGE) return fail(ctx, IMPOP_ERR_RANGE, "a window violates sum(node_len) < 2^31");
        return fail(ctx, IMPOP_ERR_DEVICE, "device-side barrier time-out in the pairwise kernel");
    }
    return IMPOP_OK;
}

int impop_pack_bits(impop_ctx_t *ctx, const uint8_t *dense_dev, int32_t n, int32_t m, int64_t dense_pitch,
                    uint32_t *x_dev, int32_t pitch_words, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (n < 0 || m < 0 || pitch_words < 0 || (n > 0 && m > 0 && (!dense_dev || !x_dev)) || dense_pitch < m ||
        (int64_t)pitch_words * 32 < m)
        return fail(ctx, IMPOP_ERR_ARG, "impop_pack_bits: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_pack_bits(dense_dev, n, m, dense_pitch, x_dev, pitch_words, ctx->sm_count, (cudaStream_t)stream));
    ctx->launches += ((int64_t)n * pitch_words > 0);
    return IMPOP_OK;
}

int impop_batch_destroy(impop_ctx_t *ctx, impop_batch_t *batch) {
    if (!ctx || !batch) return IMPOP_ERR_ARG;
    // kernels still in flight on the batch's stream stay valid: the blocks are reused at once only by work on that same
    // stream (ordered behind them), by other streams after the event recorded here has completed
    cudaSetDevice(ctx->device);
    if (batch->tables) pool_put(ctx, batch->tables, batch->last_stream);
    if (batch->scratch) pool_put(ctx, batch->scratch, batch->last_stream);
    delete batch;
    return IMPOP_OK;
}

namespace {
// Sequential sub-allocation inside one block, 256-byte aligned.
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; }
};
}  // namespace

int impop_batch_create(impop_ctx_t *ctx, const impop_batch_desc_t *d, impop_batch_t **batch_out) {
    if (!ctx || !d || !batch_out) return IMPOP_ERR_ARG;
    *batch_out = nullptr;
    const int32_t W = d->windows;
    if (W < 0) return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: negative window count");
    if (W > 0 && (!d->n_host || !d->m_host || !d->pitch_words_host || !d->x_off_host || !d->len_off_host ||
                  !d->lab_off_host || !d->length_host))
        return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: null descriptor array");
    if (!ctx->kernels_configured) {
        CU(cudaSetDevice(ctx->device));
        CU(configure_kernels());
        ctx->prep_per_sm = prep_rows_ctas_per_sm();
        ctx->kernels_configured = true;
    }
    const int32_t *n = d->n_host, *m = d->m_host, *pitch = d->pitch_words_host;
    std::vector<int64_t> row_off(W + 1, 0), item_off(W + 1, 0), word_off(W + 1, 0);
    bool any_rows = false, any_nodes = false;
    for (int32_t w = 0; w < W; ++w) {
        if (n[w] < 0 || m[w] < 0 || n[w] > (1 << 24) || m[w] > (1 << 24))
            return fail(ctx, IMPOP_ERR_RANGE, "impop_batch_create: n or m out of range (max 2^24)");
        if (pitch[w] % 4 != 0 || d->x_off_host[w] % 4 != 0 || (int64_t)pitch[w] * 32 < m[w])
            return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: pitch_words / x_off must be multiples of 4 and cover m");
        if (d->x_off_host[w] < 0 || d->len_off_host[w] < 0 || d->lab_off_host[w] < 0)
            return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: negative offset");
        if (d->lab_off_host[w] > 0x7FFFFFFFll || row_off[w] + n[w] > 0x7FFFFFFFll)
            return fail(ctx, IMPOP_ERR_RANGE, "impop_batch_create: more than 2^31 haplotype rows / label bytes in one batch");
        row_off[w + 1] = row_off[w] + n[w];
        item_off[w + 1] = item_off[w] + items_of(n[w]);
        word_off[w + 1] = word_off[w] + (m[w] + 31) / 32;
        any_rows |= n[w] > 0;
        any_nodes |= m[w] > 0;
    }
    if ((any_rows && any_nodes && !d->x_dev) || (any_nodes && !d->node_len_dev) || (any_rows && !d->labels_dev))
        return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: null device array");
    if (((uintptr_t)d->x_dev & 15u) != 0u)          // rows are read 16 bytes at a time (cp.async in the pairs kernel, uint4 in prep_rows)
        return fail(ctx, IMPOP_ERR_ARG, "impop_batch_create: x_dev must be 16-byte aligned");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)d->stream;
    // prep row slices: one per window when the batch has enough windows to fill the GPU, else windows are cut
    // into slices of >= 32 rows so that about 4 CTAs per SM have work
    std::vector<int4> slices;
    {
        const int64_t target = (int64_t)ctx->sm_count * 4;
        const int per_window = (W > 0 && W < target) ? (int)((target + W - 1) / W) : 1;
        for (int32_t w = 0; w < W; ++w) {
            int cnt = std::min(per_window, std::max(1, (n[w] + 31) / 32));
            int rows = (((n[w] + cnt - 1) / std::max(cnt, 1)) + 31) & ~31;
            if (rows < 32) rows = 32;
            for (int lo = 0; lo < n[w] || lo == 0; lo += rows) slices.push_back(make_int4(w, lo, std::min(lo + rows, n[w]), 0));
        }
    }
    impop_batch *b = new (std::nothrow) impop_batch();
    if (!b) return fail(ctx, IMPOP_ERR_NOMEM, "impop_batch_create: out of host memory");
    WindowTab &t = b->tab;
    auto bail = [&](int code, const std::string &why) {
        impop_batch_destroy(ctx, b);
        return fail(ctx, code, why);
    };

    // ---- descriptor tables: one pooled device block, filled by one (two) asynchronous copies from pinned staging
    const size_t W1 = (size_t)W + 1;
    Carver ca;
    const size_t o_n = ca.take(4 * W1), o_m = ca.take(4 * W1), o_pitch = ca.take(4 * W1);
    const size_t o_xoff = ca.take(8 * W1), o_lenoff = ca.take(8 * W1), o_laboff = ca.take(8 * W1), o_L = ca.take(8 * W1);
    const size_t o_row = ca.take(8 * W1), o_item = ca.take(8 * W1);
    const size_t o_items = ca.take(16 * (size_t)(item_off[W] + 1)), o_iext = ca.take(16 * (size_t)(item_off[W] + 1));
    const size_t o_slices = ca.take(16 * (slices.size() + 1)), o_word = ca.take(8 * W1);
    const size_t phase1 = ca.off;
    const size_t o_heavy = ca.take(8 * W1), o_w8 = ca.take(8 * W1), o_xh = ca.take(8 * W1);
    const size_t o_cnt = ca.take(4 * W1);
    const size_t o_runs = ca.take(8 * W1), o_wconst = ca.take(8 * W1);
    const size_t tables_bytes = ca.off;
    b->last_stream = st;
    b->tables = pool_get(ctx, tables_bytes, st);
    impop_ctx::Staging *sg = staging_get(ctx, tables_bytes);
    if (!b->tables || !sg) return bail(IMPOP_ERR_NOMEM, "impop_batch_create: out of memory (tables)");
    char *hb = (char *)sg->host, *db = (char *)b->tables;
    if (W > 0) {
        memcpy(hb + o_n, n, 4 * (size_t)W); memcpy(hb + o_m, m, 4 * (size_t)W); memcpy(hb + o_pitch, pitch, 4 * (size_t)W);
        memcpy(hb + o_xoff, d->x_off_host, 8 * (size_t)W); memcpy(hb + o_lenoff, d->len_off_host, 8 * (size_t)W);
        memcpy(hb + o_laboff, d->lab_off_host, 8 * (size_t)W); memcpy(hb + o_L, d->length_host, 8 * (size_t)W);
    }
    memcpy(hb + o_row, row_off.data(), 8 * W1);
    memcpy(hb + o_item, item_off.data(), 8 * W1);
    memcpy(hb + o_word, word_off.data(), 8 * W1);
    if (!slices.empty()) memcpy(hb + o_slices, slices.data(), 16 * slices.size());
    {   // work-item table: (window, row block, first column, columns)
        //      and what the epilogue needs of the item's window in one load: (n, m, row_off, lab_off)
        int4 *items = (int4 *)(hb + o_items), *iext = (int4 *)(hb + o_iext);
        int64_t k = 0;
        for (int32_t w = 0; w < W; ++w) {
            const int nb = (n[w] + TILE_M - 1) / TILE_M;
            const int4 ext = make_int4(n[w], m[w], (int)row_off[w], (int)d->lab_off_host[w]);
            for (int bi = 0; bi < nb; ++bi) {
                const int cnt = items_of_rowblock(n[w], bi), width = width_of_rowblock(n[w], bi);
                for (int r = 0; r < cnt; ++r) {
                    const int rev = (r == 0 && (bi & 1)) ? ITEM_REV : 0;      // every other diagonal block (common.cuh)
                    iext[k] = ext;
                    items[k++] = make_int4(w, bi | rev, bi * TILE_M + r * width, width);
                }
            }
        }
    }
    cudaError_t e = cudaMemcpyAsync(db, hb, phase1, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return bail(IMPOP_ERR_CUDA, std::string("upload tables: ") + cudaGetErrorString(e));
    t.n = (const int32_t *)(db + o_n); t.m = (const int32_t *)(db + o_m); t.pitch = (const int32_t *)(db + o_pitch);
    t.x_off = (const int64_t *)(db + o_xoff); t.len_off = (const int64_t *)(db + o_lenoff);
    t.lab_off = (const int64_t *)(db + o_laboff); t.L = (const int64_t *)(db + o_L);
    t.row_off = (const int64_t *)(db + o_row); t.item_off = (const int64_t *)(db + o_item);
    t.items = (const int4 *)(db + o_items); t.items_ext = (const int4 *)(db + o_iext);
    t.slices = (const int4 *)(db + o_slices); t.word_off = (const int64_t *)(db + o_word); t.n_slices = (int32_t)slices.size();
    t.heavy_off = (const int64_t *)(db + o_heavy); t.w8_off = (const int64_t *)(db + o_w8); t.xh_off = (const int64_t *)(db + o_xh);
    t.x = d->x_dev; t.len = d->node_len_dev; t.labels = d->labels_dev;
    t.W = W; t.err = ctx->err_dev; t.harm = ctx->harm_dev; t.harm_n = HARM_N;

    // ---- heavy-entry counts: from the caller's host copy of node_len when given (no synchronisation),
    //      else counted on the device and read back (one stream synchronisation)
    int32_t *cnt = (int32_t *)(hb + o_cnt);
    if (W > 0 && d->heavy_entries_host) {
        for (int32_t w = 0; w < W; ++w) {
            if (d->heavy_entries_host[w] < 0) return bail(IMPOP_ERR_ARG, "impop_batch_create: negative heavy_entries_host");
            cnt[w] = d->heavy_entries_host[w];
        }
    } else if (W > 0 && d->node_len_host) {
        // branch-free so that the compiler vectorises it; a few host threads for large batches
        auto count_range = [&](int32_t w0, int32_t w1) {
            for (int32_t w = w0; w < w1; ++w) {
                const uint32_t *len = d->node_len_host + d->len_off_host[w];
                uint32_t c = 0;
                const int32_t mm = m[w];
                for (int32_t k = 0; k < mm; ++k) c += (len[k] / HEAVY_Q + 254u) / 255u;
                cnt[w] = (int32_t)c;
            }
        };
        int64_t nodes = 0;
        for (int32_t w = 0; w < W; ++w) nodes += m[w];
        const int threads = nodes > (1 << 18) ? 4 : 1;
        if (threads == 1) count_range(0, W);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t)
                pool.emplace_back(count_range, (int32_t)((int64_t)W * t / threads), (int32_t)((int64_t)W * (t + 1) / threads));
            for (auto &th : pool) th.join();
        }
    } else if (W > 0) {
        if ((e = launch_heavy_count(t.len, t.len_off, t.m, W, (int32_t *)(db + o_cnt), st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(cnt, db + o_cnt, 4 * (size_t)W, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess)
            return bail(IMPOP_ERR_CUDA, std::string("heavy_count: ") + cudaGetErrorString(e));
        ctx->launches += 1;
    }
    int64_t *heavy_off = (int64_t *)(hb + o_heavy), *w8_off = (int64_t *)(hb + o_w8), *xh_off = (int64_t *)(hb + o_xh);
    heavy_off[0] = w8_off[0] = xh_off[0] = 0;
    for (int32_t w = 0; w < W; ++w) {
        const int64_t hpad = ((int64_t)(cnt[w] + KCHUNK - 1) / KCHUNK) * KCHUNK;
        const int64_t m64 = ((int64_t)(m[w] + KCHUNK - 1) / KCHUNK) * KCHUNK;
        heavy_off[w + 1] = heavy_off[w] + hpad;
        w8_off[w + 1] = w8_off[w] + m64 + hpad;                 // virtual columns: dense | heavy
        xh_off[w + 1] = xh_off[w] + (int64_t)n[w] * (hpad / 32);
    }
    if (d->site_runs_host && W > 0) memcpy(hb + o_runs, d->site_runs_host, 8 * (size_t)W);
    t.site_runs_given = (d->site_runs_host && W > 0) ? (const int64_t *)(db + o_runs) : nullptr;
    if (e == cudaSuccess && t.site_runs_given) e = cudaMemcpyAsync(db + o_runs, hb + o_runs, 8 * (size_t)W, cudaMemcpyHostToDevice, st);
    // affine form (optional): window constants travel with the tables, row terms and column multiplicities are the caller's device arrays
    if (d->win_const_host && W > 0) memcpy(hb + o_wconst, d->win_const_host, 8 * (size_t)W);
    t.win_const = (d->win_const_host && W > 0) ? (const int64_t *)(db + o_wconst) : nullptr;
    if (e == cudaSuccess && t.win_const) e = cudaMemcpyAsync(db + o_wconst, hb + o_wconst, 8 * (size_t)W, cudaMemcpyHostToDevice, st);
    t.row_adj = d->row_adj_dev;
    t.col_mult = d->col_mult_dev;
    if (e == cudaSuccess) e = cudaMemcpyAsync(db + o_heavy, hb + o_heavy, o_cnt - o_heavy, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) { e = cudaEventRecord(sg->done, st); sg->busy = true; }
    if (e != cudaSuccess) return bail(IMPOP_ERR_CUDA, std::string("upload tables: ") + cudaGetErrorString(e));

    // ---- scratch: one pooled block
    b->items = item_off[W];
    Carver cs;
    const size_t s_A = cs.take(4 * (size_t)(row_off[W] + 1)), s_w8 = cs.take((size_t)w8_off[W] + 64),
                 s_w8n = cs.take((size_t)w8_off[W] + 64), s_planes = cs.take((size_t)w8_off[W] + 64);
    const size_t s_heavy = cs.take(4 * (size_t)(heavy_off[W] + 64)), s_xh = cs.take(4 * (size_t)(xh_off[W] + 4));
    const size_t s_part = cs.take(8 * (size_t)(b->items * PART_STRIDE + 1));
    const size_t s_sums = cs.take(8 * 4 * W1), s_counts = cs.take(8 * IMPOP_NCOUNTS * W1);
    const size_t s_any = cs.take(4 * (size_t)(word_off[W] + 1)), s_all = cs.take(4 * (size_t)(word_off[W] + 1));
    const size_t s_live = cs.take(4 * (size_t)(word_off[W] + 1));
    const size_t s_hn = cs.take(4 * W1), s_runs = cs.take(4 * W1);
    b->scratch = pool_get(ctx, cs.off, st);
    if (!b->scratch) return bail(IMPOP_ERR_NOMEM, "impop_batch_create: out of device memory (scratch)");
    char *sb = (char *)b->scratch;
    t.A = (int32_t *)(sb + s_A); t.w8 = (uint8_t *)(sb + s_w8); t.w8n = (uint8_t *)(sb + s_w8n); t.planes = (uint32_t *)(sb + s_planes); t.heavy = (uint32_t *)(sb + s_heavy);
    t.xh = (uint32_t *)(sb + s_xh); t.seg_any = (uint32_t *)(sb + s_any); t.seg_all = (uint32_t *)(sb + s_all);
    t.live = (uint32_t *)(sb + s_live);
    t.heavy_n = (int32_t *)(sb + s_hn); t.site_runs = (int32_t *)(sb + s_runs);
    b->partials = (double *)(sb + s_part); b->sums_tmp = (double *)(sb + s_sums); b->counts_tmp = (int64_t *)(sb + s_counts);
    b->item_off = item_off;
    b->n.assign(n, n + W);
    *batch_out = b;
    return IMPOP_OK;
}

int64_t impop_batch_items(const impop_batch_t *batch) { return batch ? batch->items : 0; }

static int run_sums(impop_ctx_t *ctx, impop_batch_t *b, int32_t algo, int32_t rank, int32_t world, double *sums_dev,
                    cudaStream_t st) {
    if (algo != IMPOP_ALGO_TCGEN05 && algo != IMPOP_ALGO_SIMT) return fail(ctx, IMPOP_ERR_ARG, "unknown algo");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, IMPOP_ERR_ARG, "bad rank/world");
    if (b->tab.W == 0) return IMPOP_OK;
    b->last_stream = st;
    CU(timed(ctx, IMPOP_KERNEL_PREP, st, [&] { return launch_prep(b->tab, b->counts_tmp, ctx->sm_count, ctx->prep_per_sm, st); }));
    ctx->launches += 3;                                        // prep_cols, prep_rows, seg_count
    ItemParams prm{};
    prm.partials = b->partials;
    prm.item_begin = 0; prm.item_end = b->items; prm.rank = rank; prm.world = world;
    prm.dumpI = nullptr; prm.dumpPi = nullptr; prm.prof = ctx->prof_dev;
    const int64_t mine = (b->items - rank + world - 1) / world;
    CU(timed(ctx, IMPOP_KERNEL_PAIRS, st, [&] { return launch_pairs(b->tab, prm, algo, ctx->sm_count, st); }));
    ctx->launches += (mine > 0);
    CU(timed(ctx, IMPOP_KERNEL_SUMS, st, [&] { return launch_window_sums(b->tab, b->partials, rank, world, sums_dev, ctx->sm_count, st); }));
    ctx->launches += 1;
    return IMPOP_OK;
}

int impop_window_sums(impop_ctx_t *ctx, impop_batch_t *batch, int32_t algo, int32_t rank, int32_t world,
                      double *sums_dev, void *stream) {
    if (!ctx || !batch || (!sums_dev && batch->tab.W > 0)) return fail(ctx, IMPOP_ERR_ARG, "impop_window_sums: null argument");
    CU(cudaSetDevice(ctx->device));
    return run_sums(ctx, batch, algo, rank, world, sums_dev, (cudaStream_t)stream);
}

int impop_window_finalize(impop_ctx_t *ctx, impop_batch_t *batch, const double *sums_dev, int32_t parts,
                          double *stats_dev, int64_t *counts_dev, void *stream) {
    if (!ctx || !batch) return IMPOP_ERR_ARG;
    if (batch->tab.W == 0) return IMPOP_OK;
    if (!sums_dev || !stats_dev || parts < 1) return fail(ctx, IMPOP_ERR_ARG, "impop_window_finalize: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // label counts and segregating nodes were formed by the prep pass of the preceding impop_window_sums / _stats
    batch->last_stream = st;
    CU(timed(ctx, IMPOP_KERNEL_FINALIZE, st, [&] { return launch_finalize(batch->tab, sums_dev, parts, batch->counts_tmp, stats_dev, ctx->sm_count, st); }));
    if (counts_dev)
        CU(cudaMemcpyAsync(counts_dev, batch->counts_tmp, sizeof(int64_t) * IMPOP_NCOUNTS * (size_t)batch->tab.W,
                           cudaMemcpyDeviceToDevice, st));
    ctx->launches += 1;
    return IMPOP_OK;
}

int impop_window_stats(impop_ctx_t *ctx, impop_batch_t *batch, int32_t algo, double *stats_dev, int64_t *counts_dev,
                       void *stream) {
    if (!ctx || !batch) return IMPOP_ERR_ARG;
    if (batch->tab.W == 0) return IMPOP_OK;
    if (!stats_dev) return fail(ctx, IMPOP_ERR_ARG, "impop_window_stats: null stats_dev");
    CU(cudaSetDevice(ctx->device));
    int rc = run_sums(ctx, batch, algo, 0, 1, batch->sums_tmp, (cudaStream_t)stream);
    if (rc != IMPOP_OK) return rc;
    return impop_window_finalize(ctx, batch, batch->sums_tmp, 1, stats_dev, counts_dev, stream);
}

int impop_pairwise(impop_ctx_t *ctx, impop_batch_t *batch, int32_t window, int32_t algo, int64_t *I_dev,
                   int64_t *A_dev, double *pi_dev, void *stream) {
    if (!ctx || !batch) return IMPOP_ERR_ARG;
    if (window < 0 || window >= batch->tab.W) return fail(ctx, IMPOP_ERR_ARG, "impop_pairwise: window out of range");
    if (algo != IMPOP_ALGO_TCGEN05 && algo != IMPOP_ALGO_SIMT) return fail(ctx, IMPOP_ERR_ARG, "unknown algo");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    batch->last_stream = st;
    CU(launch_prep(batch->tab, batch->counts_tmp, ctx->sm_count, ctx->prep_per_sm, st));
    ctx->launches += 3;
    if (I_dev || pi_dev) {
        ItemParams prm{};
        prm.partials = batch->partials;
        prm.item_begin = batch->item_off[window]; prm.item_end = batch->item_off[window + 1];
        prm.rank = 0; prm.world = 1; prm.dumpI = I_dev; prm.dumpPi = pi_dev;
        CU(launch_pairs(batch->tab, prm, algo, ctx->sm_count, st));
        ctx->launches += 1;
    }
    if (A_dev) {
        int64_t off = 0;
        for (int32_t w = 0; w < window; ++w) off += batch->n[w];
        CU(launch_export_a(batch->tab.A + off, batch->n[window], A_dev, st));
        ctx->launches += 1;
    }
    return IMPOP_OK;
}

int impop_reduce_identity(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld, const uint8_t *labels_dev,
                          const double *weight_dev, int64_t length, double seg_sites, double *stats_dev,
                          int64_t *counts_dev, double *wsum_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (n < 0 || ld < n || (n > 0 && !ident_dev)) return fail(ctx, IMPOP_ERR_ARG, "impop_reduce_identity: bad argument");
    CU(cudaSetDevice(ctx->device));
    // per-call scratch from the pool (calls on different streams never share it)
    cudaStream_t st = (cudaStream_t)stream;
    double *scratch = (double *)pool_get(ctx, sizeof(double) * 16 * (size_t)reduce_identity_blocks(n, ctx->sm_count), st);
    if (!scratch) return fail(ctx, IMPOP_ERR_NOMEM, "impop_reduce_identity: out of device memory");
    cudaError_t e = launch_reduce_identity(ident_dev, n, ld, labels_dev, weight_dev, length, seg_sites, ctx->harm_dev, HARM_N,
                                           scratch, ctx->sm_count, stats_dev, counts_dev, wsum_dev, st);
    pool_put(ctx, scratch, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "impop_reduce_identity");
    ctx->launches += 2;
    return IMPOP_OK;
}

int impop_tajima_d(impop_ctx_t *ctx, const int64_t *n_dev, const double *S_dev, const double *pi_dev, int32_t count,
                   double *D_dev, double *parts_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (count < 0 || (count > 0 && (!n_dev || !S_dev || !pi_dev || !D_dev)))
        return fail(ctx, IMPOP_ERR_ARG, "impop_tajima_d: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_tajima(n_dev, S_dev, pi_dev, count, D_dev, parts_dev, (cudaStream_t)stream));
    ctx->launches += (count > 0);
    return IMPOP_OK;
}

int impop_site_counts(impop_ctx_t *ctx, const uint64_t *sites_dev, int64_t sites, int32_t words,
                      const uint64_t *masks_dev, int32_t pops, int32_t *counts_dev, double *freq_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (sites < 0 || words < 1 || pops < 1 || pops > 64 || (int64_t)pops * words > 2048 ||
        (sites > 0 && (!sites_dev || !masks_dev || !counts_dev)))
        return fail(ctx, IMPOP_ERR_ARG, "impop_site_counts: bad argument (pops <= 64, pops * words <= 2048)");
    CU(cudaSetDevice(ctx->device));
    CU(timed(ctx, IMPOP_KERNEL_SITES, (cudaStream_t)stream, [&] {
        return launch_site_counts(sites_dev, sites, words, masks_dev, pops, counts_dev, freq_dev, ctx->sm_count,
                                  (cudaStream_t)stream);
    }));
    ctx->launches += (sites > 0);
    return IMPOP_OK;
}

int impop_cluster(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld, double threshold,
                  int32_t *comp_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (n < 0 || ld < n || (n > 0 && (!ident_dev || !comp_dev))) return fail(ctx, IMPOP_ERR_ARG, "impop_cluster: bad argument");
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->cluster_cap) {
        CU(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(ctx->cluster_parent);
        ctx->cluster_parent = nullptr;
        ctx->cluster_cap = 0;
        CU(cudaMalloc(&ctx->cluster_parent, sizeof(int32_t) * (size_t)n));
        ctx->cluster_cap = n;
    }
    CU(launch_cluster(ident_dev, n, ld, threshold, ctx->cluster_parent, comp_dev, ctx->sm_count, (cudaStream_t)stream));
    ctx->launches += (n > 0) ? 3 : 0;
    return IMPOP_OK;
}

int impop_greedy_groups(impop_ctx_t *ctx, const double *ident_dev, int32_t n, int64_t ld, double threshold,
                        int32_t *group_dev, double *weight_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (n < 0 || ld < n || (n > 0 && (!ident_dev || !group_dev)))
        return fail(ctx, IMPOP_ERR_ARG, "impop_greedy_groups: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_greedy_groups(ident_dev, n, ld, threshold, group_dev, weight_dev, (cudaStream_t)stream));
    ctx->launches += (n > 0);
    return IMPOP_OK;
}

int impop_selftest_division(impop_ctx_t *ctx, uint64_t seed, int64_t count, int64_t *mismatches_host, void *stream) {
    if (!ctx || !mismatches_host) return IMPOP_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    unsigned long long *dev = nullptr;
    CU(cudaMalloc(&dev, sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(dev, 0, sizeof(unsigned long long), (cudaStream_t)stream);
    if (e == cudaSuccess) e = launch_division_selftest(seed, count, dev, ctx->sm_count, (cudaStream_t)stream);
    unsigned long long host = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host, dev, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dev);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "impop_selftest_division");
    ctx->launches += 1;
    *mismatches_host = (int64_t)host;
    return IMPOP_OK;
}

int impop_round_decimal(impop_ctx_t *ctx, double *values_dev, int64_t count, int32_t digits, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (count < 0 || digits < 0 || digits > 22 || (count > 0 && !values_dev))
        return fail(ctx, IMPOP_ERR_ARG, "impop_round_decimal: bad argument (0 <= digits <= 22)");
    CU(cudaSetDevice(ctx->device));
    CU(launch_round_decimal(values_dev, count, digits, ctx->sm_count, (cudaStream_t)stream));
    ctx->launches += (count > 0);
    return IMPOP_OK;
}

int impop_repitch_rows(impop_ctx_t *ctx, int32_t windows, const int32_t *rows_host, const int32_t *src_pitch_words_host,
                       const int32_t *dst_pitch_words_host, const int64_t *src_off_host, const int64_t *dst_off_host,
                       const uint32_t *src_dev, uint32_t *dst_dev, void *stream) {
    if (!ctx) return IMPOP_ERR_ARG;
    if (windows < 0 || (windows > 0 && (!rows_host || !src_pitch_words_host || !dst_pitch_words_host || !src_off_host ||
                                        !dst_off_host || !src_dev || !dst_dev)))
        return fail(ctx, IMPOP_ERR_ARG, "impop_repitch_rows: null argument");
    if (windows == 0) return IMPOP_OK;
    for (int32_t w = 0; w < windows; ++w)
        if (rows_host[w] < 0 || src_pitch_words_host[w] < 0 || dst_pitch_words_host[w] < src_pitch_words_host[w] ||
            dst_pitch_words_host[w] % 4 != 0 || src_off_host[w] < 0 || dst_off_host[w] < 0 || dst_off_host[w] % 4 != 0)
            return fail(ctx, IMPOP_ERR_ARG, "impop_repitch_rows: pitches (dst a multiple of 4 words, >= src) / offsets out of range");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(RepitchDesc) * (size_t)windows;
    impop_ctx::Staging *sg = staging_get(ctx, bytes);
    void *table = pool_get(ctx, bytes, st);
    if (!sg || !table) {
        if (table) pool_put(ctx, table, st);
        return fail(ctx, IMPOP_ERR_NOMEM, "impop_repitch_rows: out of memory (tables)");
    }
    RepitchDesc *hd = (RepitchDesc *)sg->host;
    for (int32_t w = 0; w < windows; ++w)
        hd[w] = RepitchDesc{src_off_host[w], dst_off_host[w], rows_host[w], src_pitch_words_host[w], dst_pitch_words_host[w], 0};
    cudaError_t e = cudaMemcpyAsync(table, hd, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) { e = cudaEventRecord(sg->done, st); sg->busy = true; }
    if (e == cudaSuccess) e = launch_repitch_rows((const RepitchDesc *)table, windows, src_dev, dst_dev, ctx->sm_count, st);
    pool_put(ctx, table, st);                      // reused at once only by work on `st`, by other streams after the event
    if (e != cudaSuccess) return cuda_fail(ctx, e, "impop_repitch_rows");
    ctx->launches += 1;
    return IMPOP_OK;
}

int impop_dev_alloc(impop_ctx_t *ctx, int64_t bytes, void **ptr_out) {
    if (!ctx || !ptr_out || bytes < 0) return IMPOP_ERR_ARG;
    *ptr_out = nullptr;
    CU(cudaSetDevice(ctx->device));
    if (cudaMalloc(ptr_out, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, IMPOP_ERR_NOMEM, "impop_dev_alloc: out of device memory");
    }
    return IMPOP_OK;
}

int impop_dev_free(impop_ctx_t *ctx, void *ptr) {
    if (!ctx) return IMPOP_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(ptr));
    return IMPOP_OK;
}

int impop_dev_copy(impop_ctx_t *ctx, void *dst, const void *src, int64_t bytes, int32_t kind, void *stream) {
    if (!ctx || bytes < 0 || kind < 0 || kind > 2 || (bytes > 0 && (!dst || !src))) return IMPOP_ERR_ARG;
    if (bytes == 0) return IMPOP_OK;
    CU(cudaSetDevice(ctx->device));
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return IMPOP_OK;
}

int impop_debug_role_times(impop_ctx_t *ctx, int64_t *out_host, int32_t ctas) {
    if (!ctx || !out_host || ctas < 0 || ctas > 1024) return IMPOP_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(out_host, ctx->prof_dev, sizeof(long long) * 16 * (size_t)ctas, cudaMemcpyDeviceToHost));
    return IMPOP_OK;
}

}  // extern "C"
