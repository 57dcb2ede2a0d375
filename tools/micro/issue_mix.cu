// Micro-benchmark: does an fp64 instruction occupy the SM sub-partition's issue port for two cycles?
// Per loop iteration a warp issues F independent DFMAs and I independent integer (LOP3 / IADD) instructions.
//   cycles per iteration = 2 F          -> the integer instructions hide in the fp64 pipe's second cycle
//   cycles per iteration = 2 F + I      -> an fp64 instruction blocks the issue port while it occupies the half-width pipe
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int F, int I>
__global__ void mix(double *out, long long *cyc, int iters, unsigned seed) {
    double x[F > 0 ? F : 1];
    unsigned v[I > 0 ? I : 1];
#pragma unroll
    for (int k = 0; k < (F > 0 ? F : 1); ++k) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
#pragma unroll
    for (int k = 0; k < (I > 0 ? I : 1); ++k) v[k] = threadIdx.x * 2654435761u + k + seed;
    const double y = 0.999999;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < (F > I ? F : I); ++k) {
                if (k < F) x[k] = __fma_rn(x[k], y, 1e-12);
                if (k < I) v[k] = (v[k] ^ (v[k] >> 7)) + seed;       // LOP3-class + IADD: two integer instructions
            }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    double s = 0; unsigned u = 0;
    for (int k = 0; k < (F > 0 ? F : 1); ++k) s += x[k];
    for (int k = 0; k < (I > 0 ? I : 1); ++k) u += v[k];
    if (s == 12345.0 || u == 0x12345u) out[0] = s + u;
}

template <int F, int I>
void run(int warps) {
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 64);
    const int iters = 2000;
    mix<F, I><<<1, warps * 32>>>(out, cyc, iters, 3u);
    long long h[64]; cudaMemcpy(h, cyc, 8 * warps, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int k = 0; k < warps; ++k) mx = h[k] > mx ? h[k] : mx;
    const double per_iter = (double)mx / (iters * 4.0) / ((warps + 3) / 4);     // cycles per (warp iteration) per SMSP slot
    printf("F=%2d fp64 + I=%2d x2 int per iteration, %2d warps/SM: %6.2f cycles per warp-iteration per SMSP  (2F = %d, 2F + 2I = %d, 2I = %d)\n",
           F, I, warps, per_iter, 2 * F, 2 * F + 2 * I, 2 * I);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 16}) {
        run<8, 0>(w); run<0, 8>(w); run<8, 4>(w); run<8, 8>(w); run<8, 2>(w); run<4, 8>(w);
    }
    return 0;
}
