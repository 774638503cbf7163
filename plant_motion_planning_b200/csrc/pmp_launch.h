// Non-template launchers, one translation unit per sphere-count padding (pmp_inst_a12/a16/a32.cu) so that the
// template instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include "pmp_kernels.cuh"

#define PMP_DECLARE_LAUNCHERS(A)                                                                                       \
    cudaError_t pmp_launch_edge_a##A(int nspad, const PmpEdgeArgs& args, bool wsmem, int grid, int block, size_t smem, \
                                     cudaStream_t st, int* regs);                                                      \
    cudaError_t pmp_launch_edge_first_a##A(int nspad, const PmpEdgeArgs& args, bool wsmem, int grid, int block,        \
                                           size_t smem, cudaStream_t st, int* regs);                                   \
    cudaError_t pmp_launch_config_a##A(int nspad, const PmpConfigArgs& args, dim3 grid, int block, size_t smem,        \
                                       cudaStream_t st, int* regs);

PMP_DECLARE_LAUNCHERS(12)
PMP_DECLARE_LAUNCHERS(16)
PMP_DECLARE_LAUNCHERS(18)
PMP_DECLARE_LAUNCHERS(20)
PMP_DECLARE_LAUNCHERS(32)
