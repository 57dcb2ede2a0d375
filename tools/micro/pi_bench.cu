// Micro-benchmark: throughput of the epilogue's per-pair fp64 math (pi_batch) against warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pi_bench pi_bench.cu && ./pi_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../../impop_b200/csrc/common.cuh"
using namespace impop;

template <int NP, int VARIANT>
__global__ void k(uint32_t seed, int iters, double *out) {
    uint32_t r[16], aj[16];
    uint32_t ai = 50000u + (threadIdx.x & 31u) * 7u;
#pragma unroll
    for (int q = 0; q < 16; ++q) { r[q] = 49000u + q * 13u + threadIdx.x; aj[q] = 50100u + q * 5u; }
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double p[16];
        if (VARIANT == 0) {
#pragma unroll
            for (int q = 0; q < 16; q += NP) pi_batch<NP>(r + q, ai, aj + q, p + q);
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) p[q] = pi_from_counts_fast(r[q], ai, aj[q]);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) { acc += p[q]; r[q] += 1u + (seed & 1u); }
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int NP, int VARIANT>
void run(const char *name, int warps) {
    double *out; cudaMalloc(&out, 8);
    const int iters = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NP, VARIANT><<<148, warps * 32>>>(1, 10, out);
    cudaEventRecord(a);
    k<NP, VARIANT><<<148, warps * 32>>>(1, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double pairs = 148.0 * warps * 32 * 16.0 * iters;
    double cyc_per_pairwarp_per_sm = (ms * 1e-3 * 1.965e9) / (warps * 16.0 * iters);   // SM cycles per (warp x pair)
    printf("%-28s warps/SM %2d: %7.3f ms  %.3e pairs/s  %6.2f SM-cycles per 32-pair warp-op (fp64 floor ~10-11)\n", name, warps, ms, pairs / (ms * 1e-3), cyc_per_pairwarp_per_sm);
    cudaFree(out);
}

// Dependent chain of DFMAs in one warp: cycles per instruction = latency of the fp64 pipe.
__global__ void dfma_latency(double *out, long long *cyc, int iters) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = __fma_rn(x, y, 1e-12);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 12345.0) out[0] = x;
}
// ILP independent chains in one warp: cycles per instruction -> issue interval of the pipe.
template <int ILP>
__global__ void dfma_ilp(double *out, long long *cyc, int iters) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
    const double y = 0.999999;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = __fma_rn(x[k], y, 1e-12);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 12345.0) out[0] = s;
}
template <int ILP>
void run_ilp(int warps) {
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 148);
    const int iters = 1000;
    dfma_ilp<ILP><<<1, warps * 32>>>(out, cyc, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dfma ILP %2d, %2d warps on one SM: %.2f cycles per warp-DFMA per SMSP-slot (total %lld)\n", ILP, warps,
           (double)h / (iters * 4.0 * ILP) / ((warps + 3) / 4), h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    {
        double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
        dfma_latency<<<1, 32>>>(out, cyc, 1000);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA latency: %.2f cycles\n", (double)h / 16000.0);
        cudaFree(out); cudaFree(cyc);
    }
    run_ilp<1>(4); run_ilp<2>(4); run_ilp<4>(4); run_ilp<8>(4); run_ilp<16>(4);
    run_ilp<4>(8); run_ilp<8>(8); run_ilp<8>(16); run_ilp<4>(16);
    for (int w : {4, 8, 16, 32}) run<8, 0>("layered NP=8", w);
    for (int w : {4, 8, 16, 32}) run<4, 0>("layered NP=4", w);
    for (int w : {4, 8, 16, 32}) run<16, 0>("layered NP=16", w);
    for (int w : {4, 8, 16, 32}) run<1, 1>("per pair (compiler order)", w);
    return 0;
}
