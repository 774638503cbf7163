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
#pragma once
#include "pmp_kernels.cuh"

struct PmpGjkShape {          // one posed convex shape, identical in every lane
    const float4* verts;      // hull vertices (x, y, z, -) in the shape frame; nullptr: box / point given by `half`
    int n;
    float R[9], t[3];         // world pose
    float half[3];            // verts == nullptr: core half extents of a box (all zero: a point)
};

// world support point of S in world direction (dx, dy, dz); ties resolve to the lowest vertex index
__device__ __forceinline__ void pmp_gjk_support(const PmpGjkShape& S, float dx, float dy, float dz, float (&o)[3]) {
    float lx, ly, lz;
    pmp_matT_vec(S.R, dx, dy, dz, lx, ly, lz);
    float vx, vy, vz;
    if (S.verts) {
        const int lane = threadIdx.x & 31;
        float best = -3.0e38f;
        int bi = 0x7fffffff;
        for (int i = lane; i < S.n; i += 32) {
            const float4 v = __ldg(S.verts + i);
            const float val = fmaf(v.x, lx, fmaf(v.y, ly, v.z * lz));
            if (val > best) { best = val; bi = i; }
        }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const float ob = __shfl_xor_sync(PMP_FULL, best, ofs);
            const int oi = __shfl_xor_sync(PMP_FULL, bi, ofs);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const float4 v = __ldg(S.verts + bi);
        vx = v.x; vy = v.y; vz = v.z;
    } else {
        vx = copysignf(S.half[0], lx); vy = copysignf(S.half[1], ly); vz = copysignf(S.half[2], lz);
    }
    float x, y, z;
    pmp_mat_vec(S.R, vx, vy, vz, x, y, z);
    o[0] = x + S.t[0]; o[1] = y + S.t[1]; o[2] = z + S.t[2];
}

__device__ __forceinline__ float pmp_dot3(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

// closest point to the origin on triangle (a, b, c) by Voronoi regions; idx / n = supporting feature
__device__ __forceinline__ void pmp_gjk_triangle(const float* a, const float* b, const float* c, float* out, int* idx, int& n) {
    float ab[3], ac[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; }
    const float d1 = -pmp_dot3(ab, a), d2 = -pmp_dot3(ac, a);
    if (d1 <= 0.f && d2 <= 0.f) { out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; idx[0] = 0; n = 1; return; }
    const float d3 = -pmp_dot3(ab, b), d4 = -pmp_dot3(ac, b);
    if (d3 >= 0.f && d4 <= d3) { out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; idx[0] = 1; n = 1; return; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = fmaf(v, ab[k], a[k]);
        idx[0] = 0; idx[1] = 1; n = 2; return;
    }
    const float d5 = -pmp_dot3(ab, c), d6 = -pmp_dot3(ac, c);
    if (d6 >= 0.f && d5 <= d6) { out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; idx[0] = 2; n = 1; return; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float w = d2 / (d2 - d6);
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = fmaf(w, ac[k], a[k]);
        idx[0] = 0; idx[1] = 2; n = 2; return;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = fmaf(w, c[k] - b[k], b[k]);
        idx[0] = 1; idx[1] = 2; n = 2; return;
    }
    const float den = 1.f / (va + vb + vc);
    const float v = vb * den, w = vc * den;
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = fmaf(ac[k], w, fmaf(ab[k], v, a[k]));
    idx[0] = 0; idx[1] = 1; idx[2] = 2; n = 3;
}

// "GJK distance of the margin-free shapes <= msum" (msum = sum of both collision margins).  Early exits: the current
// closest point of the simplex is an upper bound of the distance, v.w / |v| a lower bound.
__device__ __noinline__ bool pmp_gjk_hit(const PmpGjkShape& A, const PmpGjkShape& B, float msum) {
    float P[4][3];
    int ns = 0;
    float v[3], a[3], b[3], w[3];
    const float m2 = msum * msum;
    // first point of A - B: the one furthest towards the origin as seen from the difference of the shape origins
    {
        float dx = B.t[0] - A.t[0], dy = B.t[1] - A.t[1], dz = B.t[2] - A.t[2];
        if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < 1e-12f) { dx = 1.f; dy = 0.f; dz = 0.f; }
        pmp_gjk_support(A, dx, dy, dz, a);
        pmp_gjk_support(B, -dx, -dy, -dz, b);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = a[k] - b[k]; P[0][k] = v[k]; }
    ns = 1;
    float vv = pmp_dot3(v, v);
    for (int it = 0; it < 48; ++it) {
        if (vv <= m2) return true;
        pmp_gjk_support(A, -v[0], -v[1], -v[2], a);
        pmp_gjk_support(B, v[0], v[1], v[2], b);
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = a[k] - b[k];
        const float vw = pmp_dot3(v, w);
        if (vw > 0.f && vw * vw > m2 * vv) return false;          // separated by more than the margins
        if (vv - vw <= 1e-6f * vv) return false;                  // converged with |v|^2 > msum^2
        bool dup = false;
        for (int i = 0; i < ns; ++i) dup |= (P[i][0] == w[0] && P[i][1] == w[1] && P[i][2] == w[2]);
        if (dup) return false;
#pragma unroll
        for (int k = 0; k < 3; ++k) P[ns][k] = w[k];
        ++ns;
        float nv[3];
        if (ns == 2) {
            float ab[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) ab[k] = P[1][k] - P[0][k];
            const float t = -pmp_dot3(P[0], ab), den = pmp_dot3(ab, ab);
            if (t <= 0.f || den <= 0.f) { nv[0] = P[0][0]; nv[1] = P[0][1]; nv[2] = P[0][2]; ns = 1; }
            else if (t >= den) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { nv[k] = P[1][k]; P[0][k] = P[1][k]; }
                ns = 1;
            } else {
                const float u = t / den;
#pragma unroll
                for (int k = 0; k < 3; ++k) nv[k] = fmaf(u, ab[k], P[0][k]);
            }
        } else if (ns == 3) {
            int idx[3], n;
            pmp_gjk_triangle(P[0], P[1], P[2], nv, idx, n);
            float Q[3][3];
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) Q[i][k] = P[idx[i]][k];
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) P[i][k] = Q[i][k];
            ns = n;
        } else {
            // tetrahedron: faces the origin lies outside of; a (nearly) flat one has no reliable inside
            const int faces[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
            float e1[3], e2[3], e3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { e1[k] = P[1][k] - P[0][k]; e2[k] = P[2][k] - P[0][k]; e3[k] = P[3][k] - P[0][k]; }
            const float cr[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
            const float vol = fabsf(pmp_dot3(e1, cr));
            const float scale = sqrtf(pmp_dot3(e1, e1) * pmp_dot3(e2, e2) * pmp_dot3(e3, e3));
            const bool flat = !(vol > 1e-4f * scale);
            float best = 3.0e38f, bp[3] = {0.f, 0.f, 0.f};
            int bidx[3] = {0, 0, 0}, bn = 0, bf = -1;
            for (int f = 0; f < 4; ++f) {
                const float *pa = P[faces[f][0]], *pb = P[faces[f][1]], *pc = P[faces[f][2]], *pd = P[faces[f][3]];
                float ab[3], ac[3], ad[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { ab[k] = pb[k] - pa[k]; ac[k] = pc[k] - pa[k]; ad[k] = pd[k] - pa[k]; }
                const float nn[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
                const float sp = -pmp_dot3(pa, nn), sd = pmp_dot3(ad, nn);
                const bool outside = flat || sd == 0.f || sp * sd < 0.f;
                if (!outside) continue;
                float q[3];
                int idx[3], n;
                pmp_gjk_triangle(pa, pb, pc, q, idx, n);
                const float dd = pmp_dot3(q, q);
                if (dd < best) {
                    best = dd; bp[0] = q[0]; bp[1] = q[1]; bp[2] = q[2];
                    bidx[0] = idx[0]; bidx[1] = idx[1]; bidx[2] = idx[2]; bn = n; bf = f;
                }
            }
            if (bf < 0) return true;                               // the origin is inside: the shapes intersect
            float Q[3][3];
            for (int i = 0; i < bn; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) Q[i][k] = P[faces[bf][bidx[i]]][k];
            for (int i = 0; i < bn; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) P[i][k] = Q[i][k];
            ns = bn;
            nv[0] = bp[0]; nv[1] = bp[1]; nv[2] = bp[2];
        }
        const float nvv = pmp_dot3(nv, nv);
        if (nvv >= vv) return false;                               // no progress (rounding): |v|^2 > msum^2 stands
        v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
        vv = nvv;
    }
    return vv <= m2;
}

struct PmpRigidArgs {
    const PmpRobotTab* robot;
    const PmpHullTab* hulls;
    const float4* hull_verts;
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
};

#define PMP_RIGID_WARPS 4
__global__ void __launch_bounds__(32 * PMP_RIGID_WARPS) pmp_rigid_kernel(const PmpRigidArgs args) {
    __shared__ __align__(16) PmpRobotTab rt;
    __shared__ __align__(16) PmpHullTab ht;
    __shared__ double sm_sq_all[PMP_RIGID_WARPS][PMP_MAX_DOF];
    __shared__ float sm_q_all[PMP_RIGID_WARPS][3 * PMP_MAX_DOF];
    __shared__ float sm_fr_all[PMP_RIGID_WARPS][PMP_MAX_DOF][12];
    __shared__ float sm_sc_all[PMP_RIGID_WARPS][PMP_MAX_SPHERES][3];
    pmp_stage_words(&rt, args.robot, sizeof(PmpRobotTab) / 4);
    pmp_stage_words(&ht, args.hulls, sizeof(PmpHullTab) / 4);
    __syncthreads();
    const PmpKernelParams prm = args.prm;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sm_sq = sm_sq_all[wid];
    float* sm_q0 = sm_q_all[wid];
    float* sm_q1 = sm_q0 + PMP_MAX_DOF;
    float* sm_dq = sm_q1 + PMP_MAX_DOF;
    float (*fr)[12] = sm_fr_all[wid];
    float (*sc)[3] = sm_sc_all[wid];
    const int D = rt.dof, A = rt.n_spheres, H = ht.n_hulls;
    const bool edge_mode = args.q == nullptr;
    const int C = (args.max_steps + 31) >> 5;
    const int n_items = edge_mode ? args.E * C : (args.B + 31) >> 5;
    const uint32_t circ = rt.circular_mask;
    const bool lag = prm.lag_kappa > 0.f;
    for (int item = blockIdx.x * PMP_RIGID_WARPS + wid; item < n_items; item += gridDim.x * PMP_RIGID_WARPS) {
        bool live, at_end = false;
        float t = 0.f, tg = 0.f;
        int cfg = 0;
        if (edge_mode) {
            const int e = item / C, chunk = item - e * C;
            const int n = pmp_edge_setup(rt, prm, args.q0, args.q1, e, lane, sm_sq, sm_q0, sm_q1, sm_dq);
            const int nc = min(n, args.max_steps);
            if (chunk * 32 >= nc) {
                if (lane == 0) args.rigid_pre[item] = 0u;
                continue;
            }
            const int i = chunk * 32 + lane;
            live = i < nc;
            const int num = i + (args.backward ? 0 : 1);
            const float inv_n = 1.0f / (float)n;
            t = (float)num * inv_n;
            at_end = (num == n);
            tg = pmp_lagged_t(prm, i, t, inv_n);
        } else {
            cfg = item * 32 + lane;
            live = cfg < args.B;
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
        bool near;
        {
            float c[PMP_MAX_SPHERES][3], ee[3];
            auto q_in = [&](int j) { return rt.lower[j]; };      // the limit test is done above
            near = pmp_config_eval<PMP_MAX_SPHERES>(rt, true, q_geo, q_in, c, ee, rt.sph_radius_outer);
        }
        bool rigid = live && lim_hit;
        uint32_t todo = __ballot_sync(PMP_FULL, live && !lim_hit && near);
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
                if (p < ht.n_pairs) {
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
                    hit = pmp_gjk_hit(SA, SB, ht.hull_margin[h] + mB);
                }
            }
            if (lane == L) rigid = hit;
        }
        if (edge_mode) {
            const uint32_t m = __ballot_sync(PMP_FULL, rigid);
            if (lane == 0) args.rigid_pre[item] = m;
        } else if (live) {
            args.rigid_flags[cfg] = rigid ? 1 : 0;
        }
    }
}
