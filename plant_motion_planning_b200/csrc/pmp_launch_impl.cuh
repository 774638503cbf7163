// Implementation of the launchers declared in pmp_launch.h for one sphere-count padding A.
#pragma once
#include "pmp_launch.h"

// cudaFuncSetAttribute / cudaFuncGetAttributes are driver round trips; on the 47 us scalar collision_fn path they are
// not free.  Each instantiation remembers the largest dynamic shared-memory size it has been opted into (per device,
// the attribute is per context) and its register count.
struct PmpFuncCache {
    int smem_optin[16];
    int regs[16];
};

template <class K>
static cudaError_t pmp_prepare(K kernel, PmpFuncCache& fc, size_t smem, int* regs) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if ((int)smem > fc.smem_optin[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        fc.smem_optin[dev] = (int)smem;
    }
    if (fc.regs[dev] == 0) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        fc.regs[dev] = fa.numRegs;
    }
    if (regs) *regs = fc.regs[dev];
    return cudaSuccess;
}

template <int A, int NS, bool FIRST>
static cudaError_t pmp_launch_edge_t(const PmpEdgeArgs& args, bool wsmem, int grid, int block, size_t smem, cudaStream_t st,
                                     int* regs) {
    static PmpFuncCache fc_s = {}, fc_g = {}, fc_d = {};
    cudaError_t e;
    // roofline launches (pmp_params.dense): their own instantiation, so that the production kernel carries none of
    // that code; only the full-output, tables-in-shared-memory shape is measured that way
    // (other shapes run the production kernel, whose results are identical)
    if constexpr (!FIRST) {
        if (args.prm.dense && wsmem) {
            e = pmp_prepare(pmp_edge_kernel<A, NS, true, false, true>, fc_d, smem, regs);
            if (e != cudaSuccess) return e;
            pmp_edge_kernel<A, NS, true, false, true><<<grid, block, smem, st>>>(args);
            return cudaGetLastError();
        }
    }
    if (wsmem) {
        e = pmp_prepare(pmp_edge_kernel<A, NS, true, FIRST>, fc_s, smem, regs);
        if (e != cudaSuccess) return e;
        pmp_edge_kernel<A, NS, true, FIRST><<<grid, block, smem, st>>>(args);
    } else {
        e = pmp_prepare(pmp_edge_kernel<A, NS, false, FIRST>, fc_g, smem, regs);
        if (e != cudaSuccess) return e;
        pmp_edge_kernel<A, NS, false, FIRST><<<grid, block, smem, st>>>(args);
    }
    return cudaGetLastError();
}

template <int A, int NS>
static cudaError_t pmp_launch_config_t(const PmpConfigArgs& args, dim3 grid, int block, size_t smem, cudaStream_t st,
                                       int* regs) {
    static PmpFuncCache fc = {};
    cudaError_t e = pmp_prepare(pmp_config_kernel<A, NS>, fc, smem, regs);
    if (e != cudaSuccess) return e;
    pmp_config_kernel<A, NS><<<grid, block, smem, st>>>(args);
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
