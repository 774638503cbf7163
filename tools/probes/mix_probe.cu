// Micro-benchmark: what the instruction MIX of the edge kernel's Gauss-Newton pair loop can reach on the FP32 pipes when
// nothing but the mix limits it (independent chains, no memory, full occupancy).  Per sphere pair the SASS of the dense
// loop holds 21 FFMA2, 8 FMUL2, 5 FADD2, 14 FMNMX and 2 MUFU.RSQ; the probe issues that mix on 8
// independent accumulators per thread and reports it in the kernel's own flop units (FFMA2 = 4, FMUL2 / FADD2 = 2,
// FMNMX = 1, MUFU = 1), next to a pure-FFMA2 stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mix_probe tools/probes/mix_probe.cu && /tmp/mix_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: FFMA2 only.  1: FFMA2 / FMUL2 / FADD2 in the loop's ratio.  2: + FMNMX.  3: + MUFU.RSQ (the whole mix).
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    f2 x[8];
    float m[8], s[8];
    for (int i = 0; i < 8; ++i) { x[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); m[i] = (float)i; s[i] = 1.f + i; }
    const f2 A = pk(a, a), B = pk(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {          // one "sphere pair" per accumulator
#pragma unroll
            for (int k = 0; k < 21; ++k) x[i] = fma2(x[i], A, B);
            if (MODE >= 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[i] = mul2(x[i], A);
#pragma unroll
                for (int k = 0; k < 5; ++k) x[i] = add2(x[i], B);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {          // operands straight from the packed registers: no extra FP32 op
                    float lo, hi;
                    upk(x[(i + 1 + k) & 7], lo, hi);
                    m[i] = fmaxf(m[i], lo);
                    m[(i + 1) & 7] = fminf(m[(i + 1) & 7], hi);
                }
            }
            if (MODE >= 3) { s[i] = rsq(s[i] + 1.5f); s[(i + 5) & 7] = rsq(s[(i + 5) & 7] + 2.5f); }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; upk(x[i], lo, hi); r += lo + hi + m[i] + s[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
static void run(const char* name, float* d, int grid, int block, int iters, double flops_per_pair) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) probe<MODE><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e0);
    probe<MODE><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = 8.0 * iters * (double)grid * block;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s in the kernel's flop units\n", name, ms, pairs * flops_per_pair / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8, block = 256, iters = 512;
    float* d;
    cudaMalloc(&d, sizeof(float) * grid * block);
    // flops per "pair": FFMA2 4 each, FMUL2 / FADD2 2 each, FMNMX 1 each (two lanes of one instruction count once per
    // lane-op: 14 scalar FMNMX), MUFU 1 each
    run<0>("21 FFMA2", d, grid, block, iters, 21 * 4.0);
    run<1>("21 FFMA2 + 8 FMUL2 + 5 FADD2", d, grid, block, iters, 21 * 4.0 + 13 * 2.0);
    run<2>("... + 14 FMNMX", d, grid, block, iters, 21 * 4.0 + 13 * 2.0 + 14.0);
    run<3>("... + 2 MUFU.RSQ (the pair loop's mix)", d, grid, block, iters, 21 * 4.0 + 13 * 2.0 + 14.0 + 2.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
