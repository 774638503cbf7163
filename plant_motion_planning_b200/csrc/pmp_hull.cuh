// Exact rigid test on the convex hulls of the arm's collision meshes (sm_100a, FP32 CUDA cores).
//
// The reference decides self and obstacle collisions with p.getClosestPoints(distance=0) per link pair
// (pybullet_tools/utils.py:3390-3441, 3965-4016); PyBullet loads every URDF collision mesh as the convex hull of its
// vertices and reports a pair when the GJK distance of the margin-free shapes minus both collision margins is <= 0.
// This pre-pass evaluates exactly that for every interpolation step of an edge batch (or every explicit
// configuration) and leaves one bit per step -- limits | self | fixed -- for the edge / configuration kernels, which
// then only run the arm FK for the plant contact model.
//
// Mapping: warp = 32 consecutive steps of one edge (lane = step), like the edge kernel.  Each lane runs its FK and a
// broad phase on the arm's fitted spheres with their COVERING radii (no overlap there => no contact, exactly).  Steps
// that survive are resolved one at a time by the WHOLE warp: the 32 lanes recompute that step's joint frames, cull
// the hull pairs / hull-obstacle pairs in parallel (lane = pair) and run GJK cooperatively -- the support scan over a
// hull's vertices is strided over the lanes (coalesced float4 loads) and finished with a warp arg-max; the simplex
// logic is executed redundantly (uniform data, no divergence).  Oracle: oracle/hull_gjk.c (float64) through
// oracle/edgecheck_oracle.c config_eval.
//
// Around that core: INSCRIBED balls of the same spheres decide certain hits in the broad phase (exact: they lie inside
// hull (+) margin); hull pairs two joints apart are read from clearance tables built at set-up; warp items are claimed
// from a global counter (an item that reaches the narrow phase costs tens of times one that does not); small batches
// use items of 8 or 4 steps, so that one item gliding along an obstacle is not the kernel's makespan; a caller that
// asked for no rigid step mask reads only the first set bit of every 32-step word, so steps behind the first certain
// hit skip the narrow phase.  Robots without hulls run the same kernel in sphere mode (pair masks, static primitives).
#pragma once
#include "pmp_kernels.cuh"

// Experiment builds (-DPMP_RIGID_STATS): counters of the narrow phase, read back with cudaMemcpyFromSymbol.
//  [0] warp items  [1] steps sent to the narrow phase  [2] GJK calls on hull pairs  [3] GJK calls hull x obstacle
//  [4] GJK iterations  [5] steps decided by a table hit / limits  [6] narrow-phase steps that end in a hit
//  [16 + p] GJK calls of self pair p   [48 + f * 8 + h] GJK calls of (obstacle f, hull h)
#ifdef PMP_RIGID_STATS
__device__ unsigned long long pmp_rigid_stats[128];
#define PMP_STAT(i, v) do { const unsigned long long v_ = (unsigned long long)(v); if ((threadIdx.x & 31) == 0) atomicAdd(&pmp_rigid_stats[i], v_); } while (0)
#else
#define PMP_STAT(i, v) do { } while (0)
#endif

struct PmpGjkShape {          // one posed convex shape, identical in every lane (kept in registers)
    const float4* verts;      // hull vertices (x, y, z, -) in the shape frame; nullptr: box / point given by `half`
    int n;
    float R[9], t[3];         // world pose
    float half[3];            // verts == nullptr: core half extents of a box (all zero: a point)
};

struct PmpV3 { float x, y, z; };
__device__ __forceinline__ PmpV3 pmp_v3(float x, float y, float z) { PmpV3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ PmpV3 operator-(PmpV3 a, PmpV3 b) { return pmp_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float pmp_dot(PmpV3 a, PmpV3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ PmpV3 pmp_cross(PmpV3 a, PmpV3 b) {
    return pmp_v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ PmpV3 pmp_axpy(float s, PmpV3 a, PmpV3 b) { return pmp_v3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ bool pmp_same(PmpV3 a, PmpV3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// element k of (p0, p1, p2, p3) without dynamic indexing (the simplex stays in registers)
__device__ __forceinline__ PmpV3 pmp_pick(int k, PmpV3 p0, PmpV3 p1, PmpV3 p2, PmpV3 p3) {
    PmpV3 r = p0;
    if (k == 1) r = p1;
    if (k == 2) r = p2;
    if (k == 3) r = p3;
    return r;
}

// world support point of S in world direction d.  The scan over a hull's vertices is strided over the 32 lanes
// (coalesced 16-byte loads), then a warp arg-max; ties go to an arbitrary maximal vertex (the distance does not care).
__device__ __noinline__ PmpV3 pmp_gjk_support(const PmpGjkShape& S, PmpV3 d) {
    float lx, ly, lz;
    pmp_matT_vec(S.R, d.x, d.y, d.z, lx, ly, lz);
    float vx, vy, vz;
    if (S.verts) {
        const int lane = threadIdx.x & 31;
        float best = -3.0e38f, bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll 4
        for (int i = lane; i < S.n; i += 32) {
            const float4 v = __ldg(S.verts + i);
            const float val = fmaf(v.x, lx, fmaf(v.y, ly, v.z * lz));
            if (val > best) { best = val; bx = v.x; by = v.y; bz = v.z; }
        }
        // order-preserving map float -> uint, one warp max, the first lane holding it publishes its vertex
        const uint32_t u = __float_as_uint(best);
        const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        const uint32_t mx = __reduce_max_sync(PMP_FULL, key);
        const int src = __ffs(__ballot_sync(PMP_FULL, key == mx)) - 1;
        vx = __shfl_sync(PMP_FULL, bx, src); vy = __shfl_sync(PMP_FULL, by, src); vz = __shfl_sync(PMP_FULL, bz, src);
    } else {
        vx = copysignf(S.half[0], lx); vy = copysignf(S.half[1], ly); vz = copysignf(S.half[2], lz);
    }
    float x, y, z;
    pmp_mat_vec(S.R, vx, vy, vz, x, y, z);
    return pmp_v3(x + S.t[0], y + S.t[1], z + S.t[2]);
}

// closest point to the origin on triangle (a, b, c) by Voronoi regions; (i0, i1, i2) / n = supporting feature as
// indices into {0, 1, 2}
__device__ __forceinline__ PmpV3 pmp_gjk_triangle(PmpV3 a, PmpV3 b, PmpV3 c, int& i0, int& i1, int& i2, int& n) {
    const PmpV3 ab = b - a, ac = c - a;
    const float d1 = -pmp_dot(ab, a), d2 = -pmp_dot(ac, a);
    i0 = 0; i1 = 1; i2 = 2;
    if (d1 <= 0.f && d2 <= 0.f) { n = 1; i0 = 0; return a; }
    const float d3 = -pmp_dot(ab, b), d4 = -pmp_dot(ac, b);
    if (d3 >= 0.f && d4 <= d3) { n = 1; i0 = 1; return b; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { n = 2; i0 = 0; i1 = 1; return pmp_axpy(d1 / (d1 - d3), ab, a); }
    const float d5 = -pmp_dot(ab, c), d6 = -pmp_dot(ac, c);
    if (d6 >= 0.f && d5 <= d6) { n = 1; i0 = 2; return c; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { n = 2; i0 = 0; i1 = 2; return pmp_axpy(d2 / (d2 - d6), ac, a); }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        n = 2; i0 = 1; i1 = 2;
        return pmp_axpy((d4 - d3) / ((d4 - d3) + (d5 - d6)), c - b, b);
    }
    const float den = 1.f / (va + vb + vc);
    n = 3;
    return pmp_axpy(vc * den, ac, pmp_axpy(vb * den, ab, a));
}

// "GJK distance of the margin-free shapes <= msum" (msum = sum of both collision margins).  Early exits: the current
// closest point of the simplex is an upper bound of the distance, v.w / |v| a lower bound.  Every lane runs the same
// scalar code on the same data (only the support scans are split over the lanes).  ONE copy of this code per kernel
// (noinline, compact loops): inlined at every call site it overflowed the instruction cache (ncu: 29 "no instruction"
// stall cycles per issue).
__device__ __noinline__ bool pmp_gjk_hit(const PmpGjkShape& A, const PmpGjkShape& B, float msum) {
    PmpV3 p0, p1 = pmp_v3(0.f, 0.f, 0.f), p2 = p1, p3 = p1;
    int ns;
    const float m2 = msum * msum;
    PmpV3 v;
    // first point of A - B: the one furthest towards the origin as seen from the difference of the shape origins
    {
        PmpV3 d = pmp_v3(B.t[0] - A.t[0], B.t[1] - A.t[1], B.t[2] - A.t[2]);
        if (pmp_dot(d, d) < 1e-12f) d = pmp_v3(1.f, 0.f, 0.f);
        v = pmp_gjk_support(A, d) - pmp_gjk_support(B, pmp_v3(-d.x, -d.y, -d.z));
    }
    p0 = v;
    ns = 1;
    float vv = pmp_dot(v, v);
    for (int it = 0; it < 48; ++it) {
        PMP_STAT(4, 1);
        if (vv <= m2) return true;
        const PmpV3 w = pmp_gjk_support(A, pmp_v3(-v.x, -v.y, -v.z)) - pmp_gjk_support(B, v);
        const float vw = pmp_dot(v, w);
        if (vw > 0.f && vw * vw > m2 * vv) return false;          // separated by more than the margins
        if (vv - vw <= 1e-6f * vv) return false;                  // converged with |v|^2 > msum^2
        if (pmp_same(w, p0) || (ns > 1 && pmp_same(w, p1)) || (ns > 2 && pmp_same(w, p2))) return false;
        PmpV3 nv;
        if (ns == 1) {
            p1 = w;
            const PmpV3 ab = p1 - p0;
            const float t = -pmp_dot(p0, ab), den = pmp_dot(ab, ab);
            if (t <= 0.f || den <= 0.f) { nv = p0; ns = 1; }
            else if (t >= den) { nv = p1; p0 = p1; ns = 1; }
            else { nv = pmp_axpy(t / den, ab, p0); ns = 2; }
        } else if (ns == 2) {
            p2 = w;
            int i0, i1, i2, n;
            nv = pmp_gjk_triangle(p0, p1, p2, i0, i1, i2, n);
            const PmpV3 q0 = pmp_pick(i0, p0, p1, p2, p3), q1 = pmp_pick(i1, p0, p1, p2, p3), q2 = pmp_pick(i2, p0, p1, p2, p3);
            p0 = q0; p1 = q1; p2 = q2;
            ns = n;
        } else {
            p3 = w;
            // tetrahedron: faces the origin lies outside of; a (nearly) flat one has no reliable inside
            const PmpV3 e1 = p1 - p0, e2 = p2 - p0, e3 = p3 - p0;
            const float vol = fabsf(pmp_dot(e1, pmp_cross(e2, e3)));
            const float scale = sqrtf(pmp_dot(e1, e1) * pmp_dot(e2, e2) * pmp_dot(e3, e3));
            const bool flat = !(vol > 1e-4f * scale);
            float best = 3.0e38f;
            PmpV3 bp = pmp_v3(0.f, 0.f, 0.f);
            int b0 = 0, b1 = 0, b2 = 0, bn = 0;
            bool any = false;
#pragma unroll 1
            for (int f = 0; f < 4; ++f) {
                // faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
                const int fa = f == 3 ? 1 : 0, fb = f == 0 ? 1 : (f == 1 ? 2 : 3), fc = f == 0 ? 2 : (f == 1 ? 3 : (f == 2 ? 1 : 2)),
                          fd = f == 0 ? 3 : (f == 1 ? 1 : (f == 2 ? 2 : 0));
                const PmpV3 pa = pmp_pick(fa, p0, p1, p2, p3), pb = pmp_pick(fb, p0, p1, p2, p3), pc = pmp_pick(fc, p0, p1, p2, p3),
                            pd = pmp_pick(fd, p0, p1, p2, p3);
                const PmpV3 nn = pmp_cross(pb - pa, pc - pa);
                const float sp = -pmp_dot(pa, nn), sd = pmp_dot(pd - pa, nn);
                if (!(flat || sd == 0.f || sp * sd < 0.f)) continue;
                int i0, i1, i2, n;
                const PmpV3 q = pmp_gjk_triangle(pa, pb, pc, i0, i1, i2, n);
                const float dd = pmp_dot(q, q);
                if (dd < best) {
                    best = dd; bp = q; bn = n; any = true;
                    // map the triangle-local indices back to simplex indices
                    b0 = i0 == 0 ? fa : (i0 == 1 ? fb : fc);
                    b1 = i1 == 0 ? fa : (i1 == 1 ? fb : fc);
                    b2 = i2 == 0 ? fa : (i2 == 1 ? fb : fc);
                }
            }
            if (!any) return true;                                 // the origin is inside: the shapes intersect
            const PmpV3 q0 = pmp_pick(b0, p0, p1, p2, p3), q1 = pmp_pick(b1, p0, p1, p2, p3), q2 = pmp_pick(b2, p0, p1, p2, p3);
            p0 = q0; p1 = q1; p2 = q2;
            ns = bn;
            nv = bp;
        }
        const float nvv = pmp_dot(nv, nv);
        if (nvv >= vv) return false;                               // no progress (rounding): |v|^2 > msum^2 stands
        v = nv;
        vv = nvv;
    }
    return vv <= m2;
}

struct PmpRigidArgs {
    const PmpRobotTab* robot;
    const PmpHullTab* hulls;
    const float4* hull_verts;
    const uint8_t* pair_tables;
    PmpKernelParams prm;
    // edge mode (q == nullptr): steps of edges q0 -> q1, one bit per step into rigid_pre [E, C]
    const float* q0;
    const float* q1;
    int32_t E;
    int32_t max_steps;
    int32_t backward;
    uint32_t* rigid_pre;
    // configuration mode: explicit configurations q [B, D], one byte per configuration into rigid_flags [B]
    const float* q;
    int32_t B;
    uint8_t* rigid_flags;
    int32_t* work_counter;     // next unclaimed warp item (zeroed before the launch); null: static stride over the grid
    int32_t first_bit_only;    // edge mode: 1 = the caller reads only the FIRST set bit of every 32-step word (no rigid
                               // step mask was requested: first_hit = first violating step, and every other output is
                               // independent of the rigid bits), so steps after the first certain hit of a word skip the
                               // narrow phase and the word may lack later bits
    int32_t steps_per_item;    // edge mode: interpolation steps per warp item (32, 8 or 4).  Small batches (a planner's few
                               // hundred proposals) use short items: the narrow phase resolves an item's surviving steps
                               // one after the other, and one 32-step item gliding along an obstacle would otherwise
                               // be the whole kernel's makespan.  Items shorter than 32 steps OR their bits into the
                               // word (which the host zeroes first).
    int32_t cfgs_per_item;     // configuration mode: configurations per warp item (1 .. 32).  Small batches (the
                               // scalar collision_fn adapter) use few, so that their narrow-phase steps -- resolved one
                               // at a time by a whole warp -- spread over many warps instead of queueing in one
};

// Clearance-table lookup of hull pair p at joint values (qa, qb): 1 = certainly free, 2 = certainly colliding, 0 = ask GJK.
__device__ __forceinline__ int pmp_pair_table(const PmpHullTab& ht, const uint8_t* __restrict__ tables, int p, float qa, float qb) {
    const int na = ht.tab_na[p], nb = ht.tab_nb[p];
    const int ia = min(max(__float2int_rn((qa - ht.tab_lo_a[p]) * ht.tab_inv_a[p]), 0), na - 1);
    const int ib = min(max(__float2int_rn((qb - ht.tab_lo_b[p]) * ht.tab_inv_b[p]), 0), nb - 1);
    return (int)__ldg(tables + ht.tab_off[p] + (size_t)ia * nb + ib);
}

// Set-up kernel: fills the clearance table of one hull pair (h1 on joint j, h2 on joint j + 2; a = j + 1, b = j + 2).
// warp = grid node; the pose of h2 in h1's frame is X_a(qa) X_b(qb) with X_k = origin_k Rot(axis_k, q_k).
struct PmpPairTableArgs {
    const PmpRobotTab* robot;
    const PmpHullTab* hulls;
    const float4* hull_verts;
    uint8_t* table;
    int32_t pair;
    float band;        // Lipschitz bound of the clearance over half a grid step (+ float32 slack)
};

__global__ void __launch_bounds__(128) pmp_pair_table_kernel(const PmpPairTableArgs args) {
    __shared__ __align__(16) PmpRobotTab rt;
    __shared__ __align__(16) PmpHullTab ht;
    pmp_stage_words(&rt, args.robot, sizeof(PmpRobotTab) / 4);
    pmp_stage_words(&ht, args.hulls, sizeof(PmpHullTab) / 4);
    __syncthreads();
    const int p = args.pair, na = ht.tab_na[p], nb = ht.tab_nb[p];
    const int h1 = ht.pairs[2 * p], h2 = ht.pairs[2 * p + 1], ja = ht.tab_ja[p], jb = ht.tab_jb[p];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; node < na * nb; node += warps) {
        const int ia = node / nb, ib = node - ia * nb;
        const float qa = ht.tab_lo_a[p] + (float)ia / ht.tab_inv_a[p], qb = ht.tab_lo_b[p] + (float)ib / ht.tab_inv_b[p];
        PmpGjkShape SA, SB;
        SA.verts = args.hull_verts + ht.hull_off[h1]; SA.n = ht.hull_off[h1 + 1] - ht.hull_off[h1];
#pragma unroll
        for (int k = 0; k < 9; ++k) SA.R[k] = (k % 4 == 0) ? 1.f : 0.f;
        SA.t[0] = SA.t[1] = SA.t[2] = 0.f; SA.half[0] = SA.half[1] = SA.half[2] = 0.f;
        SB.verts = args.hull_verts + ht.hull_off[h2]; SB.n = ht.hull_off[h2 + 1] - ht.hull_off[h2];
        SB.half[0] = SB.half[1] = SB.half[2] = 0.f;
        float Ma[9], Mb[9], sa, ca, sb, cb;
        sincosf(qa, &sa, &ca);
        sincosf(qb, &sb, &cb);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            Ma[k] = fmaf(sa, rt.joint_B[9 * ja + k], fmaf(ca, rt.joint_A[9 * ja + k], rt.joint_C[9 * ja + k]));
            Mb[k] = fmaf(sb, rt.joint_B[9 * jb + k], fmaf(cb, rt.joint_A[9 * jb + k], rt.joint_C[9 * jb + k]));
        }
        pmp_mat_mul(Ma, Mb, SB.R);
        float x, y, z;
        pmp_mat_vec(Ma, rt.joint_t[3 * jb], rt.joint_t[3 * jb + 1], rt.joint_t[3 * jb + 2], x, y, z);
        SB.t[0] = x + rt.joint_t[3 * ja]; SB.t[1] = y + rt.joint_t[3 * ja + 1]; SB.t[2] = z + rt.joint_t[3 * ja + 2];
        const float msum = ht.hull_margin[h1] + ht.hull_margin[h2];
        const bool free_ = !pmp_gjk_hit(SA, SB, msum + args.band);
        const bool hit_ = msum > args.band && pmp_gjk_hit(SA, SB, msum - args.band);
        if (lane == 0) args.table[node] = (uint8_t)((free_ ? 1 : 0) | (hit_ ? 2 : 0));
    }
}

#define PMP_RIGID_WARPS 4
#ifndef PMP_RIGID_MINB
#define PMP_RIGID_MINB 4
#endif
__global__ void __launch_bounds__(32 * PMP_RIGID_WARPS, PMP_RIGID_MINB) pmp_rigid_kernel(const PmpRigidArgs args) {
    __shared__ __align__(16) PmpRobotTab rt;
    __shared__ __align__(16) PmpHullTab ht;
    __shared__ double sm_sq_all[PMP_RIGID_WARPS][PMP_MAX_DOF];
    __shared__ float sm_q_all[PMP_RIGID_WARPS][3 * PMP_MAX_DOF];
    __shared__ float sm_fr_all[PMP_RIGID_WARPS][PMP_MAX_DOF][12];
    __shared__ float sm_sc_all[PMP_RIGID_WARPS][PMP_MAX_SPHERES][3];
    __shared__ float sm_qL_all[PMP_RIGID_WARPS][PMP_MAX_DOF];
    extern __shared__ __align__(16) float sm_lane_centres[];        // [warps][3 A][32] (dynamic: 3 * A * 32 floats per warp)
    pmp_stage_words(&rt, args.robot, sizeof(PmpRobotTab) / 4);
    if (args.hulls) pmp_stage_words(&ht, args.hulls, sizeof(PmpHullTab) / 4);
    __syncthreads();
    // Without hulls (args.hulls == nullptr) the arm's SPHERE model decides: sphere pairs named by self_mask, spheres
    // against the static primitives named by fixed_mask, exact radii (oracle rigid_hit) -- the same bit per step.
    const bool hull_mode = args.hulls != nullptr;
    const PmpKernelParams prm = args.prm;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sm_sq = sm_sq_all[wid];
    float* sm_q0 = sm_q_all[wid];
    float* sm_q1 = sm_q0 + PMP_MAX_DOF;
    float* sm_dq = sm_q1 + PMP_MAX_DOF;
    float (*fr)[12] = sm_fr_all[wid];
    float (*sc)[3] = sm_sc_all[wid];
    float* qL = sm_qL_all[wid];
    float* lc = sm_lane_centres + (size_t)wid * 3 * rt.n_spheres * 32;
    const int D = rt.dof, A = rt.n_spheres, H = hull_mode ? ht.n_hulls : 0;
    // Ball around all SPHERE obstacles (the block's eight corner contact spheres): steps whose covering spheres stay
    // clear of it skip those obstacles in the broad phase altogether (exact: the ball contains them).
    float gcx = 0.f, gcy = 0.f, gcz = 0.f, gR = -1.f;
    if (hull_mode) {
        int cnt = 0;
        for (int f = 0; f < ht.n_fixed_exact; ++f)
            if (rt.fixed_kind[f] == PMP_PRIM_SPHERE) { gcx += rt.fixed_t[3 * f]; gcy += rt.fixed_t[3 * f + 1]; gcz += rt.fixed_t[3 * f + 2]; ++cnt; }
        if (cnt) {
            const float inv = 1.f / (float)cnt;
            gcx *= inv; gcy *= inv; gcz *= inv;
            for (int f = 0; f < ht.n_fixed_exact; ++f)
                if (rt.fixed_kind[f] == PMP_PRIM_SPHERE) {
                    const float dx = rt.fixed_t[3 * f] - gcx, dy = rt.fixed_t[3 * f + 1] - gcy, dz = rt.fixed_t[3 * f + 2] - gcz;
                    gR = fmaxf(gR, sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))) + rt.fixed_half[3 * f]);
                }
            gR = gR * 1.00001f + 1.0e-5f;
        }
    }
    const bool edge_mode = args.q == nullptr;
    const int C = (args.max_steps + 31) >> 5;
    const int cpi = args.cfgs_per_item;
    const int spi = edge_mode ? args.steps_per_item : 32, subs = 32 / spi;
    const int n_items = edge_mode ? args.E * C * subs : (args.B + cpi - 1) / cpi;
    const uint32_t circ = rt.circular_mask;
    const bool lag = prm.lag_kappa > 0.f;
    // Items are claimed one at a time: an item whose steps reach the narrow phase costs tens of times one that does not,
    // and with a static stride the CTAs' warps finished far apart (8.5 of 16 resident warps active on average).
    for (int round = 0;; ++round) {
        int item = 0;
        if (args.work_counter) {
            if (lane == 0) item = atomicAdd(args.work_counter, 1);
            item = __shfl_sync(PMP_FULL, item, 0);
        } else {
            item = (blockIdx.x + round * gridDim.x) * PMP_RIGID_WARPS + wid;
        }
        if (item >= n_items) break;
        bool live, at_end = false;
        float t = 0.f, tg = 0.f;
        int cfg = 0;
        if (edge_mode) {
            const int word = item / subs, sub = item - word * subs;
            const int e = word / C, chunk = word - e * C;
            const int n = pmp_edge_setup(rt, prm, args.q0, args.q1, e, lane, sm_sq, sm_q0, sm_q1, sm_dq);
            const int nc = min(n, args.max_steps);
            if (chunk * 32 + sub * spi >= nc) {
                if (lane == 0 && subs == 1) args.rigid_pre[word] = 0u;
                continue;
            }
            const int i = chunk * 32 + sub * spi + lane;
            live = lane < spi && i < nc;
            const int num = i + (args.backward ? 0 : 1);
            const float inv_n = 1.0f / (float)n;
            t = (float)num * inv_n;
            at_end = (num == n);
            tg = pmp_lagged_t(prm, i, t, inv_n);
        } else {
            cfg = item * cpi + lane;
            live = lane < cpi && cfg < args.B;
            if (!live) cfg = 0;
        }
        auto q_cmd = [&](int j) {
            if (!edge_mode) return __ldg(args.q + (size_t)cfg * D + j);
            return (at_end && !((circ >> j) & 1u)) ? sm_q1[j] : fmaf(sm_dq[j], t, sm_q0[j]);
        };
        auto q_geo = [&](int j) {
            if (!edge_mode) return __ldg(args.q + (size_t)cfg * D + j);
            return lag ? fmaf(sm_dq[j], tg, sm_q0[j])
                       : ((at_end && !((circ >> j) & 1u)) ? sm_q1[j] : fmaf(sm_dq[j], t, sm_q0[j]));
        };
        // ---- per lane: limits on the commanded configuration, broad phase on the covering spheres ----
        bool lim_hit = false;
        for (int j = 0; j < D; ++j) {
            const float ql = q_cmd(j);
            lim_hit |= ((circ >> j) & 1u) ? !(fabsf(ql) <= 3.0e38f) : !(ql >= rt.lower[j] && ql <= rt.upper[j]);
        }
        // FK of this lane's step; the sphere centres go to a per-warp shared-memory tile [sphere][coordinate][lane]
        // (conflict-free: a lane only touches its own column).  Rolled loops throughout: the unrolled sphere-pair
        // code of the edge kernels, instantiated for 32 spheres, alone overflowed the instruction cache here.
        bool near = false, sure = false;
        {
            float R[9], pp[3];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = rt.base_R[k];
            pp[0] = rt.base_t[0]; pp[1] = rt.base_t[1]; pp[2] = rt.base_t[2];
            int a = 0;
#pragma unroll 1
            for (int j = 0; j < D; ++j) {
                float sj, cj;
                sincosf(q_geo(j), &sj, &cj);
                float M[9], Rn[9], tx, ty, tz;
#pragma unroll
                for (int k = 0; k < 9; ++k) M[k] = fmaf(sj, rt.joint_B[9 * j + k], fmaf(cj, rt.joint_A[9 * j + k], rt.joint_C[9 * j + k]));
                pmp_mat_vec(R, rt.joint_t[3 * j], rt.joint_t[3 * j + 1], rt.joint_t[3 * j + 2], tx, ty, tz);
                pp[0] += tx; pp[1] += ty; pp[2] += tz;
                pmp_mat_mul(R, M, Rn);
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = Rn[k];
                for (; a < A && rt.sph_joint[a] == j; ++a) {      // spheres are sorted by joint
                    float x, y, z;
                    pmp_mat_vec(R, rt.sph_center[3 * a], rt.sph_center[3 * a + 1], rt.sph_center[3 * a + 2], x, y, z);
                    lc[(3 * a) * 32 + lane] = x + pp[0]; lc[(3 * a + 1) * 32 + lane] = y + pp[1]; lc[(3 * a + 2) * 32 + lane] = z + pp[2];
                }
            }
            if (!hull_mode) {
#pragma unroll 1
                for (int a1 = 0; a1 < A; ++a1) {
                    const uint32_t m = rt.self_mask[a1];
                    for (int b1 = a1 + 1; b1 < A; ++b1) {
                        if (!((m >> b1) & 1u)) continue;
                        const float dx = lc[(3 * a1) * 32 + lane] - lc[(3 * b1) * 32 + lane],
                                    dy = lc[(3 * a1 + 1) * 32 + lane] - lc[(3 * b1 + 1) * 32 + lane],
                                    dz = lc[(3 * a1 + 2) * 32 + lane] - lc[(3 * b1 + 2) * 32 + lane];
                        const float rr = rt.sph_radius[a1] + rt.sph_radius[b1];
                        sure |= fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr;
                    }
                }
#pragma unroll 1
                for (int f = 0; f < rt.n_fixed; ++f) {
                    const uint32_t m = rt.fixed_mask[f];
                    for (int a1 = 0; a1 < A; ++a1)
                        if ((m >> a1) & 1u)
                            sure |= pmp_prim_hit(rt.fixed_kind[f], rt.fixed_R + 9 * f, rt.fixed_t + 3 * f, rt.fixed_half + 3 * f,
                                                 lc[(3 * a1) * 32 + lane], lc[(3 * a1 + 1) * 32 + lane], lc[(3 * a1 + 2) * 32 + lane],
                                                 rt.sph_radius[a1]);
                }
            }
            // covering spheres of the hull pairs without a clearance table; the INSCRIBED balls of the same spheres
            // (inside hull (+) margin) decide a certain hit right here when two of them overlap
#pragma unroll 1
            for (int p = 0; hull_mode && p < ht.n_pairs; ++p) {
                if (ht.tab_off[p] >= 0) continue;
                const int h1 = ht.pairs[2 * p], h2 = ht.pairs[2 * p + 1];
                for (int a1 = ht.sph_lo[h1]; a1 < ht.sph_hi[h1]; ++a1)
                    for (int b1 = ht.sph_lo[h2]; b1 < ht.sph_hi[h2]; ++b1) {
                        const float dx = lc[(3 * a1) * 32 + lane] - lc[(3 * b1) * 32 + lane],
                                    dy = lc[(3 * a1 + 1) * 32 + lane] - lc[(3 * b1 + 1) * 32 + lane],
                                    dz = lc[(3 * a1 + 2) * 32 + lane] - lc[(3 * b1 + 2) * 32 + lane];
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        const float rr = rt.sph_radius_outer[a1] + rt.sph_radius_outer[b1];
                        near |= d2 <= rr * rr;
                        const float ia = rt.sph_radius_inner[a1], ib = rt.sph_radius_inner[b1];
                        sure |= ia > 0.f && ib > 0.f && d2 <= (ia + ib) * (ia + ib);
                    }
            }
            // ... and against the obstacles (Bullet's box = core box (half - margin) (+) margin; sphere = point (+) radius)
            bool grp = false;
            if (hull_mode && gR >= 0.f) {
                for (int a1 = 0; a1 < A; ++a1) {
                    const float dx = lc[(3 * a1) * 32 + lane] - gcx, dy = lc[(3 * a1 + 1) * 32 + lane] - gcy, dz = lc[(3 * a1 + 2) * 32 + lane] - gcz;
                    const float rr = rt.sph_radius_outer[a1] + gR;
                    grp |= fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr;
                }
                grp = __any_sync(PMP_FULL, grp);
            }
#pragma unroll 1
            for (int f = 0; hull_mode && f < ht.n_fixed_exact; ++f) {
                const int kind = rt.fixed_kind[f];
                if (kind == PMP_PRIM_SPHERE && !grp) continue;
                const float mB = kind == PMP_PRIM_BOX ? ht.fixed_margin : rt.fixed_half[3 * f];
                float core[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) core[k] = kind == PMP_PRIM_BOX ? fmaxf(rt.fixed_half[3 * f + k] - mB, 0.f) : 0.f;
                for (int a1 = 0; a1 < A; ++a1) {
                    const float cx = lc[(3 * a1) * 32 + lane], cy = lc[(3 * a1 + 1) * 32 + lane], cz = lc[(3 * a1 + 2) * 32 + lane];
                    near |= pmp_prim_hit(kind, rt.fixed_R + 9 * f, rt.fixed_t + 3 * f, rt.fixed_half + 3 * f, cx, cy, cz,
                                         rt.sph_radius_outer[a1]);
                    const float ia = rt.sph_radius_inner[a1];
                    sure |= ia > 0.f && pmp_prim_hit(PMP_PRIM_BOX, rt.fixed_R + 9 * f, rt.fixed_t + 3 * f, core, cx, cy, cz, ia + mB);
                }
            }
        }
        // hull pairs with a clearance table are decided by one lookup; only its narrow band goes on to GJK
        bool tab_hit = false;
        for (int p = 0; hull_mode && p < ht.n_pairs; ++p) {
            if (ht.tab_off[p] < 0) continue;
            const int r = pmp_pair_table(ht, args.pair_tables, p, q_geo(ht.tab_ja[p]), q_geo(ht.tab_jb[p]));
            tab_hit |= r == 2;
            near |= r == 0;
        }
        bool rigid = live && (lim_hit || tab_hit || sure);
        uint32_t todo = __ballot_sync(PMP_FULL, live && !lim_hit && !tab_hit && !sure && near);
        const bool first_bit_only = edge_mode && args.first_bit_only != 0;
        if (first_bit_only) {
            const uint32_t known = __ballot_sync(PMP_FULL, rigid);
            if (known) todo &= (1u << (__ffs(known) - 1)) - 1u;          // only steps before the first certain hit matter
        }
        PMP_STAT(0, 1); PMP_STAT(1, __popc(todo)); PMP_STAT(5, __popc(__ballot_sync(PMP_FULL, rigid)));
        // ---- narrow phase: one surviving step at a time, by the whole warp ----
        while (todo) {
            const int L = __ffs(todo) - 1;
            todo &= todo - 1;
            const float tL = __shfl_sync(PMP_FULL, t, L), tgL = __shfl_sync(PMP_FULL, tg, L);
            const bool endL = __shfl_sync(PMP_FULL, (int)at_end, L) != 0;
            const int cfgL = __shfl_sync(PMP_FULL, cfg, L);
            // joint frames of step L (every lane computes the same chain; lane 0 publishes them)
            {
                float R[9], p[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = rt.base_R[k];
                p[0] = rt.base_t[0]; p[1] = rt.base_t[1]; p[2] = rt.base_t[2];
                __syncwarp();
                for (int j = 0; j < D; ++j) {
                    float qj;
                    if (!edge_mode) qj = __ldg(args.q + (size_t)cfgL * D + j);
                    else qj = lag ? fmaf(sm_dq[j], tgL, sm_q0[j])
                                  : ((endL && !((circ >> j) & 1u)) ? sm_q1[j] : fmaf(sm_dq[j], tL, sm_q0[j]));
                    if (lane == 0) qL[j] = qj;
                    float s, co;
                    sincosf(qj, &s, &co);
                    float M[9], Rn[9], tx, ty, tz;
#pragma unroll
                    for (int k = 0; k < 9; ++k) M[k] = fmaf(s, rt.joint_B[9 * j + k], fmaf(co, rt.joint_A[9 * j + k], rt.joint_C[9 * j + k]));
                    pmp_mat_vec(R, rt.joint_t[3 * j], rt.joint_t[3 * j + 1], rt.joint_t[3 * j + 2], tx, ty, tz);
                    p[0] += tx; p[1] += ty; p[2] += tz;
                    pmp_mat_mul(R, M, Rn);
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = Rn[k];
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 9; ++k) fr[j][k] = R[k];
                        fr[j][9] = p[0]; fr[j][10] = p[1]; fr[j][11] = p[2];
                    }
                }
                __syncwarp();
                if (lane < A) {            // lane a: world centre of sphere a
                    const int j = rt.sph_joint[lane];
                    float x, y, z;
                    pmp_mat_vec(fr[j], rt.sph_center[3 * lane], rt.sph_center[3 * lane + 1], rt.sph_center[3 * lane + 2], x, y, z);
                    sc[lane][0] = x + fr[j][9]; sc[lane][1] = y + fr[j][10]; sc[lane][2] = z + fr[j][11];
                }
                __syncwarp();
            }
            auto load_hull = [&](int h, PmpGjkShape& S) {
                const int j = ht.hull_joint[h];
                S.verts = args.hull_verts + ht.hull_off[h];
                S.n = ht.hull_off[h + 1] - ht.hull_off[h];
#pragma unroll
                for (int k = 0; k < 9; ++k) S.R[k] = fr[j][k];
                S.t[0] = fr[j][9]; S.t[1] = fr[j][10]; S.t[2] = fr[j][11];
                S.half[0] = S.half[1] = S.half[2] = 0.f;
            };
            bool hit = false;
            // self pairs: lane p culls pair p on the covering spheres, then the flagged pairs go through GJK
            for (int p0 = 0; p0 < ht.n_pairs && !hit; p0 += 32) {
                bool cand = false;
                const int p = p0 + lane;
                if (p < ht.n_pairs && ht.tab_off[p] >= 0) {
                    cand = pmp_pair_table(ht, args.pair_tables, p, qL[ht.tab_ja[p]], qL[ht.tab_jb[p]]) == 0;
                } else if (p < ht.n_pairs) {
                    const int h1 = ht.pairs[2 * p], h2 = ht.pairs[2 * p + 1];
                    for (int a = ht.sph_lo[h1]; a < ht.sph_hi[h1]; ++a)
                        for (int b = ht.sph_lo[h2]; b < ht.sph_hi[h2]; ++b) {
                            const float dx = sc[a][0] - sc[b][0], dy = sc[a][1] - sc[b][1], dz = sc[a][2] - sc[b][2];
                            const float rr = rt.sph_radius_outer[a] + rt.sph_radius_outer[b];
                            cand |= fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr;
                        }
                }
                uint32_t cm = __ballot_sync(PMP_FULL, cand);
                while (cm && !hit) {
                    const int pp = p0 + __ffs(cm) - 1;
                    cm &= cm - 1;
                    const int h1 = ht.pairs[2 * pp], h2 = ht.pairs[2 * pp + 1];
                    PmpGjkShape SA, SB;
                    load_hull(h1, SA);
                    load_hull(h2, SB);
                    PMP_STAT(2, 1); PMP_STAT(16 + (pp & 31), 1);
                    hit = pmp_gjk_hit(SA, SB, ht.hull_margin[h1] + ht.hull_margin[h2]);
                }
            }
            // obstacles x hulls: lane = (obstacle, hull) pair
            const int n_fh = ht.n_fixed_exact * H;
            for (int p0 = 0; p0 < n_fh && !hit; p0 += 32) {
                bool cand = false;
                const int p = p0 + lane;
                if (p < n_fh) {
                    const int f = p / H, h = p - f * H;
                    for (int a = ht.sph_lo[h]; a < ht.sph_hi[h]; ++a)
                        cand |= pmp_prim_hit(rt.fixed_kind[f], rt.fixed_R + 9 * f, rt.fixed_t + 3 * f, rt.fixed_half + 3 * f,
                                             sc[a][0], sc[a][1], sc[a][2], rt.sph_radius_outer[a]);
                }
                uint32_t cm = __ballot_sync(PMP_FULL, cand);
                while (cm && !hit) {
                    const int pp = p0 + __ffs(cm) - 1;
                    cm &= cm - 1;
                    const int f = pp / H, h = pp - f * H;
                    PmpGjkShape SA, SB;
                    load_hull(h, SA);
                    SB.verts = nullptr; SB.n = 0;
#pragma unroll
                    for (int k = 0; k < 9; ++k) SB.R[k] = rt.fixed_R[9 * f + k];
                    SB.t[0] = rt.fixed_t[3 * f]; SB.t[1] = rt.fixed_t[3 * f + 1]; SB.t[2] = rt.fixed_t[3 * f + 2];
                    float mB;
                    if (rt.fixed_kind[f] == PMP_PRIM_BOX) {
                        mB = ht.fixed_margin;
#pragma unroll
                        for (int k = 0; k < 3; ++k) SB.half[k] = fmaxf(rt.fixed_half[3 * f + k] - mB, 0.f);
                    } else {               // sphere: a point with its radius as margin
                        mB = rt.fixed_half[3 * f];
                        SB.half[0] = SB.half[1] = SB.half[2] = 0.f;
                    }
                    PMP_STAT(3, 1); PMP_STAT(48 + ((f * 8 + h) & 63), 1);
                    hit = pmp_gjk_hit(SA, SB, ht.hull_margin[h] + mB);
                }
            }
            if (lane == L) rigid = hit;
            PMP_STAT(6, hit ? 1 : 0);
            if (first_bit_only && hit) break;                              // later steps of the word cannot be first
        }
        if (edge_mode) {
            const uint32_t m = __ballot_sync(PMP_FULL, rigid);
            if (lane == 0) {
                if (subs == 1) args.rigid_pre[item] = m;
                else if (m) atomicOr(args.rigid_pre + item / subs, m << ((item % subs) * spi));
            }
        } else if (live) {
            args.rigid_flags[cfg] = rigid ? 1 : 0;
        }
    }
}
