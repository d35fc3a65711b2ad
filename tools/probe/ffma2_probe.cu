// micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput and issue behaviour on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_ffma2(float* out, int iters) {
    float2 a0 = make_float2(threadIdx.x * 1e-3f, 1.f), a1 = make_float2(2.f, 3.f), a2 = make_float2(4.f, 5.f), a3 = make_float2(6.f, 7.f);
    float2 a4 = make_float2(8.f, 9.f), a5 = make_float2(10.f, 11.f), a6 = make_float2(12.f, 13.f), a7 = make_float2(14.f, 15.f);
    const float2 m = make_float2(0.999f + blockIdx.x * 1e-9f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c);
            a4 = __ffma2_rn(a4, m, c); a5 = __ffma2_rn(a5, m, c); a6 = __ffma2_rn(a6, m, c); a7 = __ffma2_rn(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a1.x + a2.x + a3.x + a4.x + a5.x + a6.x + a7.x + a0.y + a1.y + a2.y + a3.y + a4.y + a5.y + a6.y + a7.y;
}
// mixed: FFMA2 + integer/other work to see whether freed issue slots are usable
__global__ void k_mixed(float* out, int iters, int use2) {
    float2 a0 = make_float2(threadIdx.x * 1e-3f, 1.f), a1 = make_float2(2.f, 3.f), a2 = make_float2(4.f, 5.f), a3 = make_float2(6.f, 7.f);
    const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    unsigned x0 = threadIdx.x, x1 = 7, x2 = 11, x3 = 13;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (use2) { a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c); }
            else { a0.x = __fmaf_rn(a0.x, m.x, c.x); a0.y = __fmaf_rn(a0.y, m.y, c.y); a1.x = __fmaf_rn(a1.x, m.x, c.x); a1.y = __fmaf_rn(a1.y, m.y, c.y);
                   a2.x = __fmaf_rn(a2.x, m.x, c.x); a2.y = __fmaf_rn(a2.y, m.y, c.y); a3.x = __fmaf_rn(a3.x, m.x, c.x); a3.y = __fmaf_rn(a3.y, m.y, c.y); }
            x0 = x0 * 1664525u + 1013904223u; x1 ^= x0 >> 3; x2 = x2 * 22695477u + x1; x3 ^= x2 << 1;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a1.x + a2.x + a3.x + a0.y + a1.y + a2.y + a3.y + (float)(x0 ^ x1 ^ x2 ^ x3);
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int blocks = 148 * 8, threads = 256, iters = 8192;
    float t1 = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters); });
    float t2 = timeit([&] { k_ffma2<<<blocks, threads>>>(out, iters); });
    double f1 = 2.0 * 32 * iters * blocks * threads, f2 = 2.0 * f1;
    printf("FFMA  : %.3f ms  %.1f TFLOP/s\n", t1, f1 / t1 / 1e9);
    printf("FFMA2 : %.3f ms  %.1f TFLOP/s\n", t2, f2 / t2 / 1e9);
    float t3 = timeit([&] { k_mixed<<<blocks, threads>>>(out, iters, 0); });
    float t4 = timeit([&] { k_mixed<<<blocks, threads>>>(out, iters, 1); });
    printf("mixed 8 FMA + 8 int ops per unroll: scalar %.3f ms, packed %.3f ms\n", t3, t4);
    return 0;
}
