// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) issue and pipe throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe tools/probes/ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b),
             rc = *reinterpret_cast<uint64_t*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>   // 0: scalar FFMA, 1: FFMA2, 2: scalar FFMA + FMNMX interleaved, 3: FFMA2 + FMNMX interleaved
__global__ void probe(float* out, int iters, float a, float b) {
    float2 x[8];
    float m[8];
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); m[i] = x[i].x; }
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
                else x[i] = ffma2(x[i], A, B);
                if (MODE >= 2) m[i] = fmaxf(m[i] , x[(i + 3) & 7].y);      // one ALU-pipe FMNMX per pair of FMAs
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, float* d, int grid, int block, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) probe<MODE><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e0);
    probe<MODE><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 2.0 * 8 * 16 * (double)iters * grid * block;      // scalar-equivalent FMAs
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (fma flops only)\n", name, ms, 2.0 * fma / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8, block = 256, iters = 2048;
    float* d;
    cudaMalloc(&d, sizeof(float) * grid * block);
    run<0>("scalar FFMA", d, grid, block, iters);
    run<1>("packed FFMA2", d, grid, block, iters);
    run<2>("scalar FFMA + 1 FMNMX/pair", d, grid, block, iters);
    run<3>("packed FFMA2 + 1 FMNMX/pair", d, grid, block, iters);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
