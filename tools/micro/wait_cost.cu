// Micro-benchmark: what one iteration of an mbarrier wait loop costs on sm_100a -- mbarrier.try_wait on a phase that does
// not complete, followed by nanosleep(t); and try_wait with a suspend-time hint.  (How many issue slots does a waiting
// warp burn?)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wait_cost wait_cost.cu && ./wait_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(long long *out, int iters, unsigned t) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    __syncthreads();
    const uint32_t addr = smem_u32(&bar);
    uint32_t done = 0, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(0u) : "memory");
            if (t) __nanosleep(t);
        } else if (MODE == 1) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(0u), "r"(t) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(0u) : "memory");
            if (t) __nanosleep(t);
        }
        acc += done;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}

template <int MODE>
void run(const char *name, unsigned t) {
    long long *out; cudaMalloc(&out, 16);
    const int iters = 2000;
    k<MODE><<<1, 32>>>(out, iters, t);
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-44s t = %6u ns: %9.1f cycles per iteration\n", name, t, (double)h[0] / iters);
    cudaFree(out);
}

int main() {
    for (unsigned t : {0u, 20u, 50u, 100u, 200u, 500u, 1000u, 2000u}) run<0>("try_wait + nanosleep(t)", t);
    for (unsigned t : {20u, 100u, 500u, 1000u, 2000u, 10000u, 100000u}) run<1>("try_wait with suspend-time hint t", t);
    for (unsigned t : {0u, 100u, 500u}) run<2>("test_wait + nanosleep(t)", t);
    return 0;
}
