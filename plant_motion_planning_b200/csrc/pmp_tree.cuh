// Device-resident tree growth for batched RRT extension: the callers either side of the edge
// kernel.  Small kernels launched on the caller's stream by pmp_tree_extend (pmp_edgecheck.cu) around the edge kernel:
//
//   pmp_nearest_min_kernel / pmp_nearest_kernel / pmp_nearest_finish_kernel
//        nearest tree node of every target under the reference metric -- argmin over the tree of
//        distance_fn(node, target) with the first minimum winning (motion_planners/motion_planners/utils.py:47-53,
//        primitives.py:207; pybullet_tools/utils.py:3604-3627).  A float32 pass over every PMP_NN_SUB-th block of
//        the tree gives an upper bound of the minimum (any sampled node is one); the full float32 pass then sends
//        every node at or below that bound (plus the filter's rounding band) -- about PMP_NN_SUB per target -- to a
//        float64 evaluation on the float32 node/target data in a fixed summation order, so the chosen INDEX (and
//        distance) is bit-identical to the float64 oracle while the tree is scanned 1 + 1/PMP_NN_SUB times.
//   pmp_edge_points_kernel
//        configuration number step[e] of edge e, the same arithmetic as the edge kernel's interpolation.
//   pmp_tree_append_kernel
//        appends the last safe configuration of every extension to the tree (deterministic order: batch order).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmp_kernels.cuh"   // packed-pair helpers (pmp_f2, pmp_fma2, ...)
#include "pmp_tables.h"

#define PMP_NN_TGT_MAX 8      // targets per CTA: 8 for D <= 8, 4 above (targets and running minima live in registers)
#define PMP_NN_BLOCK 256
#define PMP_NN_SUB 8          // pass 1 looks at every 8th block of PMP_NN_BLOCK nodes of a CTA's slice

struct PmpNearestArgs {
    const float* nodes;        // [T, D]
    const int32_t* size_dev;   // live node count on the device (NULL: T)
    int32_t T;                 // host upper bound of the node count
    const float* targets;      // [B, D]
    int32_t B, D;
    uint32_t circular_mask;
    int32_t splits;            // CTAs along the tree axis
    uint32_t* min32;           // [B] float32 bits of the smallest float32 squared distance (pass 1)
    double* part_r;            // [B, splits] partial minimum distance
    int32_t* part_i;           // [B, splits] its node index
    int32_t* index;            // [B]
    double* dist;              // [B] or null
    float* q_from;             // [B, D] gathered nearest node or null
};

// lexicographic (distance, index) "a better than b"
__device__ __forceinline__ bool pmp_nn_better(double ra, int ia, double rb, int ib) {
    return ra < rb || (ra == rb && ia < ib);
}

// Exact float64 distance of one (node, target) pair: sqrt of the per-joint squares summed in joint order, the
// difference wrapped on continuous joints with Python's float modulo ((d + pi) % 2pi - pi,
// pybullet_tools/utils.py:1666-1686, 3604-3627; fmod is exact).  Only near-minimal pairs get here.
__device__ __noinline__ double pmp_nn_exact(const float* __restrict__ node, const float* __restrict__ target, int D,
                                            uint32_t circ) {
    double s = 0.0;
    for (int j = 0; j < D; ++j) {
        double d = __dsub_rn((double)target[j], (double)node[j]);
        if ((circ >> j) & 1u) {
            const double two_pi = 6.283185307179586, pi = 3.141592653589793;
            double m = fmod(__dadd_rn(d, pi), two_pi);
            if (m < 0.0) m = __dadd_rn(m, two_pi);
            d = __dsub_rn(m, pi);
        }
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    return __dsqrt_rn(s);
}

// float32 squared distances of one node to TWO targets, used as the filter (relative error <= (D + 2) * 2^-24, plus
// 1.3e-5 per continuous joint): packed pairs, low half = target 2p, high half = target 2p + 1.
template <int DP, bool CIRC>
__device__ __forceinline__ pmp_f2 pmp_nn_s32x2(const pmp_f2 (&t)[DP], const float (&q)[DP], uint32_t circ) {
    pmp_f2 s = pmp_bc(0.f);
#pragma unroll
    for (int j = 0; j < DP; ++j) {
        pmp_f2 d = pmp_sub2(t[j], pmp_bc(q[j]));
        if (CIRC && ((circ >> j) & 1u)) {
            float d0, d1;
            pmp_upk(d, d0, d1);
            d0 = fmaf(-6.2831853071795865f, rintf(d0 * 0.15915494309189535f), d0);
            d1 = fmaf(-6.2831853071795865f, rintf(d1 * 0.15915494309189535f), d1);
            d = pmp_pk(d0, d1);
        }
        s = pmp_fma2(d, d, s);
    }
    return s;
}

__device__ __forceinline__ void pmp_nn_range(const PmpNearestArgs& a, int& lo, int& hi) {
    int T = a.T;
    if (a.size_dev) T = min(T, *a.size_dev);
    // contiguous slice of the tree per blockIdx.y, a multiple of the block size
    int per = (T + a.splits - 1) / a.splits;
    per = (per + PMP_NN_BLOCK - 1) / PMP_NN_BLOCK * PMP_NN_BLOCK;
    lo = blockIdx.y * per;
    hi = min(T, lo + per);
}

// Pass 1: smallest float32 squared distance of each target over a sample of the tree -- the first block of every
// slice and every PMP_NN_SUB-th after it, so node 0 is always in -- (atomicMin on the float bits: the values are
// non-negative).  Pass 2 only needs a bound that the true nearest node passes, which the distance of ANY node is; the
// tighter the bound the fewer float64 evaluations.  DP = DOF or its padding (6, 7, 8 or 16), TGT targets per CTA, all
// in registers; CIRC = the robot has continuous joints (their wrap costs three instructions per term, so arms
// without one get a kernel without it).  Padding joints have target = node = 0 and contribute nothing.
template <int DP, int TGT, bool CIRC>
__global__ void __launch_bounds__(PMP_NN_BLOCK, 2) pmp_nearest_min_kernel(const PmpNearestArgs a) {
    const int D = a.D;
    const int t0 = blockIdx.x * TGT;
    static_assert(TGT % 2 == 0, "targets are processed in pairs");
    pmp_f2 tg[TGT / 2][DP];
#pragma unroll
    for (int t = 0; t < TGT; t += 2)
#pragma unroll
        for (int j = 0; j < DP; ++j)
            tg[t >> 1][j] = pmp_pk((t0 + t < a.B && j < D) ? a.targets[(size_t)(t0 + t) * D + j] : 0.f,
                                   (t0 + t + 1 < a.B && j < D) ? a.targets[(size_t)(t0 + t + 1) * D + j] : 0.f);
    int lo, hi;
    pmp_nn_range(a, lo, hi);
    float best[TGT];
#pragma unroll
    for (int t = 0; t < TGT; ++t) best[t] = 3.0e38f;
    for (int n = lo + threadIdx.x; n < hi; n += PMP_NN_BLOCK * PMP_NN_SUB) {
        float q[DP];
#pragma unroll
        for (int j = 0; j < DP; ++j) q[j] = j < D ? a.nodes[(size_t)n * D + j] : 0.f;
#pragma unroll
        for (int t = 0; t < TGT; t += 2) {
            float s0, s1;
            pmp_upk(pmp_nn_s32x2<DP, CIRC>(tg[t >> 1], q, a.circular_mask), s0, s1);
            best[t] = fminf(best[t], s0);
            best[t + 1] = fminf(best[t + 1], s1);
        }
    }
#pragma unroll
    for (int t = 0; t < TGT; ++t) {
        float v = best[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0 && t0 + t < a.B) atomicMin(&a.min32[t0 + t], __float_as_uint(v));
    }
}

// Pass 2 (the whole tree): every node whose float32 distance is within the filter's error band of the pass-1 bound is
// evaluated exactly in float64; the lexicographic minimum of (distance, index) over those is the reference's argmin.
// (The node n* of the exact minimum passes: its float32 value is at most its exact value plus the band, which is at
// most the exact value of the node behind the bound plus the band, which is at most the bound plus two bands.)
template <int DP, int TGT, bool CIRC>
__global__ void __launch_bounds__(PMP_NN_BLOCK, 2) pmp_nearest_kernel(const PmpNearestArgs a) {
    __shared__ double sm_r[PMP_NN_BLOCK / 32][TGT];
    __shared__ int sm_i[PMP_NN_BLOCK / 32][TGT];
    const int D = a.D;
    const int t0 = blockIdx.x * TGT;
    pmp_f2 tg[TGT / 2][DP];
    float thr[TGT];
    const float band = 3.0e-5f * (float)__popc(a.circular_mask);     // absolute error of the wrapped terms
#pragma unroll
    for (int t = 0; t < TGT; t += 2)
#pragma unroll
        for (int j = 0; j < DP; ++j)
            tg[t >> 1][j] = pmp_pk((t0 + t < a.B && j < D) ? a.targets[(size_t)(t0 + t) * D + j] : 0.f,
                                   (t0 + t + 1 < a.B && j < D) ? a.targets[(size_t)(t0 + t + 1) * D + j] : 0.f);
#pragma unroll
    for (int t = 0; t < TGT; ++t) {
        const float m = t0 + t < a.B ? __uint_as_float(a.min32[t0 + t]) : 0.f;
        thr[t] = (m + 2.f * band) * 1.00001f + 1e-30f;    // + underflow range of the float32 squares
    }
    int lo, hi;
    pmp_nn_range(a, lo, hi);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double best_r[TGT];
    int best_i[TGT];
#pragma unroll
    for (int t = 0; t < TGT; ++t) { best_r[t] = inf; best_i[t] = 0x7fffffff; }
    float q[DP];
    {
        const int n0 = min(lo + (int)threadIdx.x, a.T - 1);
#pragma unroll
        for (int j = 0; j < DP; ++j) q[j] = j < D ? a.nodes[(size_t)n0 * D + j] : 0.f;
    }
    for (int n = lo + threadIdx.x; n < hi; n += PMP_NN_BLOCK) {
        // the next node's loads are in flight while this one's distances are computed
        float qn[DP];
        {
            const int nn = min(n + PMP_NN_BLOCK, a.T - 1);
#pragma unroll
            for (int j = 0; j < DP; ++j) qn[j] = j < D ? a.nodes[(size_t)nn * D + j] : 0.f;
        }
        float sq[TGT];
        bool any = false;
#pragma unroll
        for (int t2 = 0; t2 < TGT; t2 += 2) {
            pmp_upk(pmp_nn_s32x2<DP, CIRC>(tg[t2 >> 1], q, a.circular_mask), sq[t2], sq[t2 + 1]);
            any |= (sq[t2] <= thr[t2]) | (sq[t2 + 1] <= thr[t2 + 1]);
        }
        if (any) {                                   // rare: about PMP_NN_SUB nodes per target pass the filter
#pragma unroll
            for (int t = 0; t < TGT; ++t) {
                if (sq[t] <= thr[t] && t0 + t < a.B) {
                    const double r = pmp_nn_exact(a.nodes + (size_t)n * D, a.targets + (size_t)(t0 + t) * D, D,
                                                  a.circular_mask);
                    if (pmp_nn_better(r, n, best_r[t], best_i[t])) { best_r[t] = r; best_i[t] = n; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < DP; ++j) q[j] = qn[j];
    }
    // block reduction of (distance, index)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < TGT; ++t) {
        double r = best_r[t];
        int i = best_i[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double r2 = __shfl_xor_sync(0xffffffffu, r, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (pmp_nn_better(r2, i2, r, i)) { r = r2; i = i2; }
        }
        if (lane == 0) { sm_r[wid][t] = r; sm_i[wid][t] = i; }
    }
    __syncthreads();
    if (threadIdx.x < TGT && t0 + threadIdx.x < a.B) {
        const int t = threadIdx.x;
        double r = sm_r[0][t];
        int i = sm_i[0][t];
        for (int w = 1; w < PMP_NN_BLOCK / 32; ++w)
            if (pmp_nn_better(sm_r[w][t], sm_i[w][t], r, i)) { r = sm_r[w][t]; i = sm_i[w][t]; }
        a.part_r[(size_t)(t0 + t) * a.splits + blockIdx.y] = r;
        a.part_i[(size_t)(t0 + t) * a.splits + blockIdx.y] = i;
    }
}

__global__ void pmp_nearest_finish_kernel(const PmpNearestArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.B) return;
    double r = a.part_r[(size_t)t * a.splits];
    int i = a.part_i[(size_t)t * a.splits];
    for (int k = 1; k < a.splits; ++k) {
        const double r2 = a.part_r[(size_t)t * a.splits + k];
        const int i2 = a.part_i[(size_t)t * a.splits + k];
        if (pmp_nn_better(r2, i2, r, i)) { r = r2; i = i2; }
    }
    if (i == 0x7fffffff) i = 0;          // no finite distance (NaN target or empty tree): node 0
    a.index[t] = i;
    if (a.dist) a.dist[t] = r;
    if (a.q_from)
        for (int j = 0; j < a.D; ++j) a.q_from[(size_t)t * a.D + j] = a.nodes[(size_t)i * a.D + j];
}

// ---- interpolation point ---------------------------------------------------------------------------
// Configuration number `step` of the edge q0 -> q1 and its step count: the arithmetic of the edge kernel's set-up
// (pmp_kernels.cuh, "interpolation set-up"; pybullet_tools/utils.py:3604-3676), one thread per edge.
__device__ __forceinline__ int pmp_edge_point(const float* __restrict__ q0, const float* __restrict__ q1, int D,
                                              uint32_t circ, double resolution, int step, int backward,
                                              float* __restrict__ out) {
    double ssum = 0.0;
    float dq[PMP_MAX_DOF];
#pragma unroll
    for (int j = 0; j < PMP_MAX_DOF; ++j) {
        dq[j] = 0.f;
        if (j < D) {
            double d = __dsub_rn((double)q1[j], (double)q0[j]);
            if ((circ >> j) & 1u) {
                const double two_pi = 6.283185307179586, pi = 3.141592653589793;
                double x = d + pi;
                x = x - two_pi * floor(x / two_pi);
                d = x - pi;
            }
            const double v = __ddiv_rn(d, resolution);
            ssum = __dadd_rn(ssum, __dmul_rn(v, v));
            dq[j] = (float)d;
        }
    }
    const double nd = floor(__dsqrt_rn(ssum));
    const int n = (nd >= 0.0 && nd <= 1.0e6) ? (int)nd + 1 : (nd > 1.0e6 ? 1000001 : 1);
    if (out) {
        const int num = step + (backward ? 0 : 1);
        const float t = (float)num * (1.0f / (float)n);
        const bool at_end = (num == n);
#pragma unroll
        for (int j = 0; j < PMP_MAX_DOF; ++j)
            if (j < D) out[j] = (at_end && !((circ >> j) & 1u)) ? q1[j] : fmaf(dq[j], t, q0[j]);
    }
    return n;
}

__global__ void pmp_edge_points_kernel(const float* q0, const float* q1, int E, int D, uint32_t circ, double resolution,
                                       const int32_t* step, int backward, float* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float a[PMP_MAX_DOF], b[PMP_MAX_DOF], o[PMP_MAX_DOF];
#pragma unroll
    for (int j = 0; j < PMP_MAX_DOF; ++j) {
        a[j] = j < D ? q0[(size_t)e * D + j] : 0.f;
        b[j] = j < D ? q1[(size_t)e * D + j] : 0.f;
    }
    pmp_edge_point(a, b, D, circ, resolution, step[e], backward, o);
#pragma unroll
    for (int j = 0; j < PMP_MAX_DOF; ++j)
        if (j < D) out[(size_t)e * D + j] = o[j];
}

// ---- tree append -----------------------------------------------------------------------------------
struct PmpTreeAppendArgs {
    float* nodes;              // [capacity, D]
    int32_t* parent;           // [capacity]
    float* via;                // [capacity, D] target of the edge the node lies on
    int32_t* via_steps;        // [capacity] number of safe steps of that edge (node = its step via_steps-1)
    int32_t* size;             // [1]
    int32_t capacity;
    const float* q_from;       // [B, D] nearest nodes
    const float* targets;      // [B, D]
    const int32_t* from_index; // [B]
    const int32_t* first_hit;  // [B]
    const int32_t* n_steps;    // [B]
    int32_t B, D, max_steps;
    uint32_t circular_mask;
    double resolution;
    int32_t* new_index;        // [B] slot of the appended node, -1 if the extension made no progress
    int32_t* reached;          // [B] 1 if the target itself was reached (and is the appended node)
    float* new_cfg;            // [B, D] appended configuration, NaN if none (null ok)
    int32_t* stats;            // [4] appended, reached, first reached batch index (B if none), overflow
};

// One CTA: the slot of every new node is size + (number of progressing extensions before it in the batch), so the
// tree layout does not depend on scheduling.
__global__ void __launch_bounds__(1024) pmp_tree_append_kernel(const PmpTreeAppendArgs a) {
    __shared__ int sm_warp[32];
    __shared__ int sm_running, sm_reached, sm_first, sm_overflow;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) { sm_running = 0; sm_reached = 0; sm_first = a.B; sm_overflow = 0; }
    __syncthreads();
    const int base = *a.size;
    const int D = a.D;
    for (int e0 = 0; e0 < a.B; e0 += blockDim.x) {
        const int e = e0 + threadIdx.x;
        int fh = 0, n = 1;
        if (e < a.B) { fh = a.first_hit[e]; n = a.n_steps[e]; }
        const bool grow = fh >= 1;
        const uint32_t m = __ballot_sync(0xffffffffu, grow);
        if (lane == 0) sm_warp[wid] = __popc(m);
        __syncthreads();
        int before = sm_running;
        for (int w = 0; w < wid; ++w) before += sm_warp[w];
        const int slot = base + before + __popc(m & ((1u << lane) - 1u));
        if (e < a.B) {
            const bool fits = slot < a.capacity;
            const bool ok = grow && fits;
            if (grow && !fits) sm_overflow = 1;
            const bool hit_target = ok && n <= a.max_steps && fh >= n;
            float o[PMP_MAX_DOF];
            if (ok) {
                float q0[PMP_MAX_DOF], q1[PMP_MAX_DOF];
#pragma unroll
                for (int j = 0; j < PMP_MAX_DOF; ++j) {
                    q0[j] = j < D ? a.q_from[(size_t)e * D + j] : 0.f;
                    q1[j] = j < D ? a.targets[(size_t)e * D + j] : 0.f;
                }
                pmp_edge_point(q0, q1, D, a.circular_mask, a.resolution, fh - 1, 0, o);
#pragma unroll
                for (int j = 0; j < PMP_MAX_DOF; ++j) {
                    if (j < D) {
                        a.nodes[(size_t)slot * D + j] = o[j];
                        a.via[(size_t)slot * D + j] = q1[j];
                    }
                }
                a.parent[slot] = a.from_index[e];
                a.via_steps[slot] = fh;
            }
            a.new_index[e] = ok ? slot : -1;
            a.reached[e] = hit_target ? 1 : 0;
            if (a.new_cfg) {
                const float nan = __int_as_float(0x7fc00000);
                for (int j = 0; j < D; ++j) a.new_cfg[(size_t)e * D + j] = ok ? o[j] : nan;
            }
            if (hit_target) { atomicAdd(&sm_reached, 1); atomicMin(&sm_first, e); }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < nw; ++w) tot += sm_warp[w];
            sm_running += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int appended = min(sm_running, max(a.capacity - base, 0));
        *a.size = base + appended;
        a.stats[0] = appended; a.stats[1] = sm_reached; a.stats[2] = sm_first; a.stats[3] = sm_overflow;
    }
}
