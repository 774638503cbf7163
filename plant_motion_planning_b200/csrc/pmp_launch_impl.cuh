// Implementation of the launchers declared in pmp_launch.h for one sphere-count padding A.
#pragma once
#include "pmp_launch.h"

template <int A, int NS, bool FIRST>
static cudaError_t pmp_launch_edge_t(const PmpEdgeArgs& args, bool wsmem, int grid, int block, size_t smem, cudaStream_t st,
                                     int* regs) {
    cudaError_t e;
    cudaFuncAttributes fa;
    if (wsmem) {
        e = cudaFuncSetAttribute(pmp_edge_kernel<A, NS, true, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pmp_edge_kernel<A, NS, true, FIRST><<<grid, block, smem, st>>>(args);
        cudaFuncGetAttributes(&fa, pmp_edge_kernel<A, NS, true, FIRST>);
    } else {
        e = cudaFuncSetAttribute(pmp_edge_kernel<A, NS, false, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pmp_edge_kernel<A, NS, false, FIRST><<<grid, block, smem, st>>>(args);
        cudaFuncGetAttributes(&fa, pmp_edge_kernel<A, NS, false, FIRST>);
    }
    if (regs) *regs = fa.numRegs;
    return cudaGetLastError();
}

template <int A, int NS>
static cudaError_t pmp_launch_config_t(const PmpConfigArgs& args, dim3 grid, int block, size_t smem, cudaStream_t st,
                                       int* regs) {
    pmp_config_kernel<A, NS><<<grid, block, smem, st>>>(args);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, pmp_config_kernel<A, NS>);
    if (regs) *regs = fa.numRegs;
    return cudaGetLastError();
}

#define PMP_DEFINE_EDGE_LAUNCHER(NAME, A, FIRST)                                                                       \
    cudaError_t NAME(int nspad, const PmpEdgeArgs& args, bool wsmem, int grid, int block, size_t smem, cudaStream_t st, \
                     int* regs) {                                                                                      \
        if (nspad == 1) return pmp_launch_edge_t<A, 1, FIRST>(args, wsmem, grid, block, smem, st, regs);               \
        if (nspad == 2) return pmp_launch_edge_t<A, 2, FIRST>(args, wsmem, grid, block, smem, st, regs);               \
        return pmp_launch_edge_t<A, 4, FIRST>(args, wsmem, grid, block, smem, st, regs);                               \
    }

#define PMP_DEFINE_LAUNCHERS(A)                                                                                        \
    PMP_DEFINE_EDGE_LAUNCHER(pmp_launch_edge_a##A, A, false)                                                           \
    cudaError_t pmp_launch_config_a##A(int nspad, const PmpConfigArgs& args, dim3 grid, int block, size_t smem,        \
                                       cudaStream_t st, int* regs) {                                                   \
        if (nspad == 1) return pmp_launch_config_t<A, 1>(args, grid, block, smem, st, regs);                           \
        if (nspad == 2) return pmp_launch_config_t<A, 2>(args, grid, block, smem, st, regs);                           \
        return pmp_launch_config_t<A, 4>(args, grid, block, smem, st, regs);                                           \
    }

// first-violation-only edge kernels live in their own translation units (pmp_inst_a*f.cu)
#define PMP_DEFINE_FIRST_LAUNCHERS(A) PMP_DEFINE_EDGE_LAUNCHER(pmp_launch_edge_first_a##A, A, true)
