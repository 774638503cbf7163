/*
 * C twin (float64, scalar, pthreads over edges) of oracle/edgecheck_oracle.py.   *** TEST INFRASTRUCTURE ***
 *
 * Used (a) by tests/ as a second, faster checker at sizes the numpy oracle cannot reach and (b) by bench.py as
 * the timed CPU baseline ("port": the reference's PyBullet path cannot run in this image).
 * The product path never links or loads this file.
 *
 * It consumes the same packed tables as the CUDA library (include/pmp_edgecheck.h) and evaluates the same
 * definitions in double precision with libm.  Reference lines restated (paths relative to
 * /root/reference/plant_motion_planning/):
 *   interpolation ........ pybullet_tools/utils.py:3604-3610, 3641-3651, 3667-3676; motion_planners/motion_planners/primitives.py:16-19
 *   limits ............... pybullet_tools/utils.py:3753-3762
 *   rigid collision ...... pybullet_tools/utils.py:3390-3441, 3965-4016: exact on the convex hulls of the arm's
 *                          collision meshes when a pmp_hull_desc is given (GJK of oracle/hull_gjk.c behind a
 *                          covering-sphere broad phase), else on the sphere model alone
 *   plant response ....... env.py:127-163, 212-219  (quasi-static model on the stems' boxes, starting from the
 *                          spring/gravity rest pose; see the .py docstring), optional arm lag (env.py:154,170)
 *   observers ............ representation.py:27-51, 84-123, 220-235, 268-294
 *   world / edge loops ... env.py:230-255; motion_planners/motion_planners/primitives.py:196-255
 *   cost ................. pybullet_tools/kuka_primitives.py:467-521
 * PARITY STATUS: as the .py oracle (sampler / interpolation / observers pinned; Bullet collision and dynamics unpinned).
 */
#include <tgmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/pmp_edgecheck.h"

/* The arithmetic type of the plant / sphere model: double for the parity twin, float (-DPMP_REAL=float) for the
 * single-precision CPU baseline of bench.py.  Step counts and the hull narrow phase are always double. */
#ifndef PMP_REAL
#define PMP_REAL double
#endif
typedef PMP_REAL real;

#define PMP_GJK_STATIC_ONLY
#include "hull_gjk.c"

#define S_RS0 0
#define S_B0 9
#define S_HW 12
#define S_HH 13
#define S_ZC 14
#define S_MARGIN 15
#define S_OBSV 16
#define S_OBSP 19
#define S_COMZ 22
#define S_RAW0 23
#define S_PSLOT 24
#define S_OSLOT 25
#define S_FLAGS 26
#define S_LIM 27
#define S_TH0 28
#define S_RO 32
#define S_TO 41
#define S_TV0 44
#define SF_PLANT_START 1
#define SF_IS_MAIN 2
#define SF_OBS_ANGLE 4
#define SF_OBS_ROOT 8
#define SF_SNAP 16

typedef struct {
    real c[PMP_MAX_SPHERES][3];
    real ee[3];
    double frames[PMP_MAX_DOF][12];   /* world frame of every joint (R row-major, t): posed hulls for the narrow phase */
    int rigid;
} config_t;

static void mat_vec(const real* R, const real* v, real* o) {
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static void matT_vec(const real* R, const real* v, real* o) {
    o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
static void mat_mul(const real* X, const real* Y, real* O) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) O[3 * r + c] = X[3 * r] * Y[c] + X[3 * r + 1] * Y[3 + c] + X[3 * r + 2] * Y[6 + c];
}
static real clampd(real v, real lo, real hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int prim_hit(int kind, const float* Rf, const float* tf, const float* hf, const real* c, real r) {
    real R[9], d[3], l[3];
    for (int i = 0; i < 9; ++i) R[i] = Rf[i];
    for (int i = 0; i < 3; ++i) d[i] = c[i] - (real)tf[i];
    matT_vec(R, d, l);
    if (kind == PMP_PRIM_BOX) {
        real e2 = 0;
        for (int i = 0; i < 3; ++i) {
            real e = fabs(l[i]) - (real)hf[i];
            if (e > 0) e2 += e * e;
        }
        return e2 <= r * r;
    } else if (kind == PMP_PRIM_CYLINDER) {
        real rho = sqrt(l[0] * l[0] + l[1] * l[1]);
        real dr = rho - (real)hf[0], dz = fabs(l[2]) - (real)hf[1];
        if (dr < 0) dr = 0;
        if (dz < 0) dz = 0;
        return dr * dr + dz * dz <= r * r;
    }
    real rr = r + (real)hf[0];
    return l[0] * l[0] + l[1] * l[1] + l[2] * l[2] <= rr * rr;
}

/* signed Bullet distance test of one posed hull against a static primitive: box = core (half - margin) + margin,
 * sphere = point + radius.  (pybullet_tools/utils.py:3390-3415 on btConvexHullShape vs btBoxShape / btSphereShape) */
static int hull_vs_fixed(const pmp_robot_desc* rb, const pmp_hull_desc* hd, int h, const double* F, int f) {
    const int kind = rb->fixed_kind[f];
    double V[8][3], Ff[12], m;
    int n;
    for (int i = 0; i < 9; ++i) Ff[i] = rb->fixed_R[9 * f + i];
    for (int i = 0; i < 3; ++i) Ff[9 + i] = rb->fixed_t[3 * f + i];
    if (kind == PMP_PRIM_BOX) {
        m = hd->fixed_margin;
        n = 8;
        for (int k = 0; k < 8; ++k)
            for (int i = 0; i < 3; ++i) {
                double hc = (double)rb->fixed_half[3 * f + i] - m;
                if (hc < 0) hc = 0;
                V[k][i] = ((k >> i) & 1) ? hc : -hc;
            }
    } else {
        m = rb->fixed_half[3 * f];
        n = 1;
        V[0][0] = V[0][1] = V[0][2] = 0.0;
    }
    const int off = hd->hull_off[h], nv = hd->hull_off[h + 1] - off;
    double* HV = (double*)malloc(sizeof(double) * 3 * (size_t)nv);
    for (int i = 0; i < 3 * nv; ++i) HV[i] = hd->hull_verts[3 * off + i];
    shape_t A = {HV, nv, F}, B = {&V[0][0], n, Ff};
    const double d = gjk_distance(&A, &B, 0);
    free(HV);
    return d - (double)hd->hull_margin[h] - m <= 0.0;
}

static int hull_vs_hull(const pmp_hull_desc* hd, int h1, const double* F1, int h2, const double* F2) {
    const int o1 = hd->hull_off[h1], n1 = hd->hull_off[h1 + 1] - o1, o2 = hd->hull_off[h2], n2 = hd->hull_off[h2 + 1] - o2;
    double* V1 = (double*)malloc(sizeof(double) * 3 * (size_t)(n1 + n2));
    double* V2 = V1 + 3 * n1;
    for (int i = 0; i < 3 * n1; ++i) V1[i] = hd->hull_verts[3 * o1 + i];
    for (int i = 0; i < 3 * n2; ++i) V2[i] = hd->hull_verts[3 * o2 + i];
    shape_t A = {V1, n1, F1}, B = {V2, n2, F2};
    const double d = gjk_distance(&A, &B, 0);
    free(V1);
    return d - (double)hd->hull_margin[h1] - (double)hd->hull_margin[h2] <= 0.0;
}

/* arm FK + spheres + world-independent rigid test.  q: the configuration the geometry is in; qlim: the commanded
 * configuration the joint-limit test sees (== q without arm lag).  hd != NULL: self / fixed collisions are decided
 * on the convex hulls; the spheres (covering radii) only cull. */
static void config_eval(const pmp_robot_desc* rb, const pmp_hull_desc* hd, const double* q, const double* qlim,
                        config_t* out) {
    real R[9], p[3];
    for (int i = 0; i < 9; ++i) R[i] = rb->base_R[i];
    for (int i = 0; i < 3; ++i) p[i] = rb->base_t[i];
    int hit = 0;
    const int D = rb->dof, A = rb->n_spheres;
    for (int j = 0; j < D; ++j) {
        const int circ = rb->circular && rb->circular[j];
        if (circ ? !(fabs(qlim[j]) <= 3.0e38) : !(qlim[j] >= (double)rb->lower[j] && qlim[j] <= (double)rb->upper[j])) hit = 1;
        const real s = sin((real)q[j]), c = cos((real)q[j]);
        real M[9], t[3], tj[3], Rn[9];
        for (int k = 0; k < 9; ++k)
            M[k] = (real)rb->joint_C[9 * j + k] + c * (real)rb->joint_A[9 * j + k] + s * (real)rb->joint_B[9 * j + k];
        for (int k = 0; k < 3; ++k) tj[k] = rb->joint_t[3 * j + k];
        mat_vec(R, tj, t);
        for (int k = 0; k < 3; ++k) p[k] += t[k];
        mat_mul(R, M, Rn);
        memcpy(R, Rn, sizeof(R));
        for (int k = 0; k < 9; ++k) out->frames[j][k] = R[k];
        for (int k = 0; k < 3; ++k) out->frames[j][9 + k] = p[k];
        for (int a = 0; a < A; ++a) {
            if (rb->sph_joint[a] == j) {
                real l[3] = {rb->sph_center[3 * a], rb->sph_center[3 * a + 1], rb->sph_center[3 * a + 2]}, w[3];
                mat_vec(R, l, w);
                for (int k = 0; k < 3; ++k) out->c[a][k] = w[k] + p[k];
            }
        }
        if (rb->ee_joint == j) {
            real l[3] = {rb->ee_local[0], rb->ee_local[1], rb->ee_local[2]}, w[3];
            mat_vec(R, l, w);
            for (int k = 0; k < 3; ++k) out->ee[k] = w[k] + p[k];
        }
    }
    if (!hd || hd->n_hulls == 0) {
        for (int a = 0; a < A && !hit; ++a) {
            const uint32_t m = rb->self_mask[a];
            for (int b = a + 1; b < A; ++b) {
                if ((m >> b) & 1u) {
                    real dx = out->c[a][0] - out->c[b][0], dy = out->c[a][1] - out->c[b][1], dz = out->c[a][2] - out->c[b][2];
                    real rr = (real)rb->sph_radius[a] + (real)rb->sph_radius[b];
                    if (dx * dx + dy * dy + dz * dz <= rr * rr) { hit = 1; break; }
                }
            }
        }
        for (int f = 0; f < rb->n_fixed && !hit; ++f) {
            for (int a = 0; a < A; ++a) {
                if ((rb->fixed_mask[f] >> a) & 1u) {
                    if (prim_hit(rb->fixed_kind[f], rb->fixed_R + 9 * f, rb->fixed_t + 3 * f, rb->fixed_half + 3 * f,
                                 out->c[a], (real)rb->sph_radius[a])) { hit = 1; break; }
                }
            }
        }
        out->rigid = hit;
        return;
    }
    /* exact: hull pairs behind the covering-sphere broad phase */
    for (int pi = 0; pi < hd->n_pairs && !hit; ++pi) {
        const int h1 = hd->pairs[2 * pi], h2 = hd->pairs[2 * pi + 1];
        int near = 0;
        for (int a = 0; a < A && !near; ++a) {
            if (hd->sph_hull[a] != h1) continue;
            for (int b = 0; b < A; ++b) {
                if (hd->sph_hull[b] != h2) continue;
                double dx = out->c[a][0] - out->c[b][0], dy = out->c[a][1] - out->c[b][1], dz = out->c[a][2] - out->c[b][2];
                double rr = (double)hd->sph_radius_outer[a] + (double)hd->sph_radius_outer[b];
                if (dx * dx + dy * dy + dz * dz <= rr * rr) { near = 1; break; }
            }
        }
        if (near && hull_vs_hull(hd, h1, out->frames[hd->hull_joint[h1]], h2, out->frames[hd->hull_joint[h2]])) hit = 1;
    }
    for (int f = 0; f < hd->n_fixed_exact && !hit; ++f) {
        for (int h = 0; h < hd->n_hulls && !hit; ++h) {
            int near = 0;
            for (int a = 0; a < A; ++a) {
                if (hd->sph_hull[a] != h) continue;
                if (prim_hit(rb->fixed_kind[f], rb->fixed_R + 9 * f, rb->fixed_t + 3 * f, rb->fixed_half + 3 * f,
                             out->c[a], (real)hd->sph_radius_outer[a])) { near = 1; break; }
            }
            if (near && hull_vs_fixed(rb, hd, h, out->frames[hd->hull_joint[h]], f)) hit = 1;
        }
    }
    out->rigid = hit;
}

typedef struct {
    int plant_hit, exceeded;
    real max_defl, over;
} unit_t;

static int as_int(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

/* one (configuration, world) unit */
static void world_eval(const pmp_robot_desc* rb, const pmp_world_desc* wd, const pmp_params* prm, int w,
                       const config_t* cf, unit_t* out, real* defl_out, real* theta_out) {
    const int A = rb->n_spheres;
    const float* wt = wd->stems + (size_t)w * wd->max_stems * PMP_STEM_STRIDE;
    const int ns = wd->n_stems[w];
    real FR[PMP_MAX_SLOTS][9], FT[PMP_MAX_SLOTS][3];
    int FM[PMP_MAX_SLOTS];
    memset(FM, 0, sizeof(FM));
    const int solve = prm->mode == PMP_MODE_OUR_METHOD;
    const int K = solve ? prm->iterations : 1;
    real last = 0.0;
    out->plant_hit = 0; out->exceeded = 0; out->max_defl = 0.0; out->over = 0.0;
    for (int s = 0; s < ns; ++s) {
        const float* rec = wt + s * PMP_STEM_STRIDE;
        const int pslot = as_int(rec[S_PSLOT]), oslot = as_int(rec[S_OSLOT]), flags = as_int(rec[S_FLAGS]);
        real Ro[9], to[3], Rv[9], tv[3], Rp[9];
        for (int i = 0; i < 9; ++i) Ro[i] = rec[S_RO + i];
        for (int i = 0; i < 3; ++i) to[i] = rec[S_TO + i];
        const real th0x = rec[S_TH0], th0y = rec[S_TH0 + 1];
        int moved = 0;
        if (pslot < 0) {
            memcpy(Rv, Ro, sizeof(Rv));
            memcpy(tv, to, sizeof(tv));
        } else {
            real x[3];
            memcpy(Rp, FR[pslot], sizeof(Rp));
            mat_mul(Rp, Ro, Rv);
            mat_vec(Rp, to, x);
            for (int i = 0; i < 3; ++i) tv[i] = x[i] + FT[pslot][i];
            moved = FM[pslot];
        }
        const real hw = rec[S_HW], hh = rec[S_HH], zc = rec[S_ZC], mg = rec[S_MARGIN], lim = rec[S_LIM];
        real thx = th0x, thy = th0y;
        int touch = 0;
        for (int it = 0; it < K; ++it) {
            const real cx = cos(thx), sx = sin(thx), cy = cos(thy), sy = sin(thy);
            const real M[9] = {cy, 0, sy, sx * sy, cx, -sx * cy, -cx * sy, sx, cx * cy};
            real Rl[9];
            mat_mul(Rv, M, Rl);
            real Hxx = 0, Hxy = 0, Hyy = 0, gx = 0, gy = 0, S = 0;
            int any = 0;
#ifdef PMP_SIMD_BASELINE
            /* The timed float32 baseline: the same sums, branch-free over the spheres so that the compiler vectorises
             * the loop (AVX2 / AVX-512 with -march=native -fopenmp-simd); a sphere that does not penetrate adds zeros. */
            const real tvx = tv[0], tvy = tv[1], tvz = tv[2];
            const real r0 = Rl[0], r1 = Rl[1], r2 = Rl[2], r3 = Rl[3], r4 = Rl[4], r5 = Rl[5], r6 = Rl[6], r7 = Rl[7], r8 = Rl[8];
            const real (*cc)[3] = cf->c;
            const float* rad = rb->sph_radius;
#pragma omp simd reduction(+ : Hxx, Hxy, Hyy, gx, gy, S) reduction(| : any)
            for (int a = 0; a < A; ++a) {
                const real wx = cc[a][0] - tvx, wy = cc[a][1] - tvy, wz = cc[a][2] - tvz;
                const real pfx = r0 * wx + r3 * wy + r6 * wz, pfy = r1 * wx + r4 * wy + r7 * wz, pfz = r2 * wx + r5 * wy + r8 * wz;
                const real pz = pfz - zc;
                const real qx = pfx < -hw ? -hw : (pfx > hw ? hw : pfx), qy = pfy < -hw ? -hw : (pfy > hw ? hw : pfy),
                           qz = pz < -hh ? -hh : (pz > hh ? hh : pz);
                const real dx = pfx - qx, dy = pfy - qy, dz = pz - qz;
                const real dist = sqrtf(dx * dx + dy * dy + dz * dz);
                const real pe = (real)rad[a] + mg - dist;
                const real pen = pe > 0 ? pe : 0;
                any |= pe > 0;
                const real inv = (real)1.0 / (dist > (real)1e-12 ? dist : (real)1e-12);
                const real mx = dy * pfz - dz * pfy, my = dz * pfx - dx * pfz, mz = dx * pfy - dy * pfx;
                const real Jx = inv * (cy * mx + sy * mz), Jy = inv * my;
                Hxx += pen * Jx * Jx; Hxy += pen * Jx * Jy; Hyy += pen * Jy * Jy;
                gx += pen * pen * Jx; gy += pen * pen * Jy; S += pen;
            }
#else
            for (int a = 0; a < A; ++a) {
                const real wv[3] = {cf->c[a][0] - tv[0], cf->c[a][1] - tv[1], cf->c[a][2] - tv[2]};
                real pf[3];
                matT_vec(Rl, wv, pf);
                const real p[3] = {pf[0], pf[1], pf[2] - zc};
                const real q[3] = {clampd(p[0], -hw, hw), clampd(p[1], -hw, hw), clampd(p[2], -hh, hh)};
                const real d[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
                const real dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const real pen = (real)rb->sph_radius[a] + mg - dist;
                if (pen <= 0) continue;
                any = 1;
                const real inv = (real)1.0 / (dist > (real)1e-12 ? dist : (real)1e-12);
                const real mx = d[1] * pf[2] - d[2] * pf[1], my = d[2] * pf[0] - d[0] * pf[2], mz = d[0] * pf[1] - d[1] * pf[0];
                const real Jx = inv * (cy * mx + sy * mz), Jy = inv * my;
                Hxx += pen * Jx * Jx; Hxy += pen * Jx * Jy; Hyy += pen * Jy * Jy;
                gx += pen * pen * Jx; gy += pen * pen * Jy; S += pen;
            }
#endif
            if (it == 0) touch = any;
            if (!solve) { if (any) out->plant_hit = 1; break; }
            if (!touch || !(S > 0)) { if (prm->dense) continue; else break; }
            const real lam = (real)prm->reg_eps * S, a11 = Hxx + lam, a22 = Hyy + lam;
            real det = a11 * a22 - Hxy * Hxy;
            const real det_lo = lam * (lam + Hxx + Hyy);   /* = det when H is rank one; lower bound otherwise */
            if (det < det_lo) det = det_lo;
            real dx = (a22 * gx - Hxy * gy) / det, dy = (a11 * gy - Hxy * gx) / det;
            dx = clampd(dx, -(real)prm->max_step, (real)prm->max_step);
            dy = clampd(dy, -(real)prm->max_step, (real)prm->max_step);
            thx = clampd(thx + dx, -lim, lim);
            thy = clampd(thy + dy, -lim, lim);
        }
        if (thx != th0x || thy != th0y) moved = 1;
        real Rs[9];
        {
            const real cx = cos(thx), sx = sin(thx), cy = cos(thy), sy = sin(thy);
            const real M[9] = {cy, 0, sy, sx * sy, cx, -sx * cy, -cx * sy, sx, cx * cy};
            mat_mul(Rv, M, Rs);
        }
        real raw = rec[S_RAW0];           /* observer value of the rest pose (undeflected chain) */
        if (moved || prm->dense) {
            const real o[3] = {rec[S_OBSV], rec[S_OBSV + 1], rec[S_OBSV + 2]};
            real val;
            if (flags & SF_OBS_ANGLE) {
                const real cz = rec[S_COMZ];
                real v[3] = {cz * Rs[2] + tv[0] - (real)rec[S_OBSP], cz * Rs[5] + tv[1] - (real)rec[S_OBSP + 1],
                               cz * Rs[8] + tv[2] - (real)rec[S_OBSP + 2]};
                if (flags & SF_SNAP) {
                    if (fabs(v[0]) < (real)1e-5) v[0] = 0;
                    if (fabs(v[1]) < (real)1e-5) v[1] = 0;
                }
                const real nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                real cs = (v[0] * o[0] + v[1] * o[1] + v[2] * o[2]) / nv;
                val = acos(clampd(cs, (real)-1.0, (real)1.0));
            } else {
                real v[3];
                if (flags & SF_OBS_ROOT) {
                    mat_vec(Rs, o, v);
                } else {
                    real l[3];
                    matT_vec(Rs, o, l);
                    mat_vec(Rp, l, v);
                }
                const real sarg = -v[0];
                real roll, pitch;
                if (sarg <= (real)-0.99999) { roll = 0; pitch = (real)-1.5707963267948966; }
                else if (sarg >= (real)0.99999) { roll = 0; pitch = (real)1.5707963267948966; }
                else { roll = atan2(v[1], v[2]); pitch = asin(sarg); }
                val = sqrt(roll * roll + pitch * pitch);
            }
            if (moved) raw = val;
        }
        if (flags & SF_PLANT_START) last = 0.0;
        real d = raw - last;
        if (d < 0) d = 0;
        if (flags & SF_IS_MAIN) last = d;
        if (d > (real)prm->deflection_limit) { out->exceeded = 1; out->over += d - (real)prm->deflection_limit; }
        if (d > out->max_defl) out->max_defl = d;
        if (defl_out) defl_out[s] = d;
        if (theta_out) { theta_out[2 * s] = thx; theta_out[2 * s + 1] = thy; }
        if (oslot >= 0) {
            memcpy(FR[oslot], Rs, sizeof(Rs));
            memcpy(FT[oslot], tv, sizeof(tv));
            FM[oslot] = moved;
        }
    }
    if (prm->mode == PMP_MODE_AVOID_ALL) {
        const float* bx = wd->boxes + (size_t)w * wd->max_boxes * PMP_BOX_STRIDE;
        for (int b = 0; b < wd->n_boxes[w] && !out->plant_hit; ++b)
            for (int a = 0; a < A; ++a)
                if (prim_hit(PMP_PRIM_BOX, bx + b * PMP_BOX_STRIDE, bx + b * PMP_BOX_STRIDE + 9,
                             bx + b * PMP_BOX_STRIDE + 12, cf->c[a], (real)rb->sph_radius[a])) { out->plant_hit = 1; break; }
    }
}

static int edge_steps(const pmp_robot_desc* rb, const pmp_params* prm, const float* q0, const float* q1, double* dq) {
    double ssum = 0.0;
    for (int j = 0; j < rb->dof; ++j) {
        double d = (double)q1[j] - (double)q0[j];
        if (rb->circular && rb->circular[j]) {
            const double two_pi = 6.283185307179586, pi = 3.141592653589793;
            double x = d + pi;
            x = x - two_pi * floor(x / two_pi);
            d = x - pi;
        }
        dq[j] = d;
        const double v = d / prm->resolution;
        ssum = ssum + v * v;
    }
    const double nd = floor(sqrt(ssum));
    return (nd >= 0.0 && nd <= 1.0e6) ? (int)nd + 1 : (nd > 1.0e6 ? 1000001 : 1);
}

typedef struct {
    const pmp_robot_desc* rb; const pmp_world_desc* wd; const pmp_params* prm; const pmp_hull_desc* hd;
    const float* q0; const float* q1; int32_t E, max_steps, backward;
    int32_t* first_hit; int32_t* n_steps; float* max_defl; float* over; float* ee_len; float* cost;
    uint32_t* rigid_mask; uint32_t* exceed_mask;
    /* explicit configurations */
    const float* q; int32_t B; uint8_t* flags; double* defl; double* theta; double* ee;
    atomic_int next;
    int chunk;
} job_t;

static void one_edge(const job_t* jb, int e) {
    const pmp_robot_desc* rb = jb->rb; const pmp_world_desc* wd = jb->wd; const pmp_params* prm = jb->prm;
    const pmp_hull_desc* hd = jb->hd;
    const float* q0 = jb->q0; const float* q1 = jb->q1;
    const int32_t max_steps = jb->max_steps, backward = jb->backward;
    int32_t* first_hit = jb->first_hit; int32_t* n_steps = jb->n_steps;
    float* max_defl = jb->max_defl; float* over = jb->over; float* ee_len = jb->ee_len; float* cost = jb->cost;
    uint32_t* rigid_mask = jb->rigid_mask; uint32_t* exceed_mask = jb->exceed_mask;
    const int D = rb->dof, W = wd->n_worlds, C = (max_steps + 31) / 32;
        double dq[PMP_MAX_DOF], q[PMP_MAX_DOF], qa[PMP_MAX_DOF];
        /* arm lag (env.py:154,170): the motor closes c = 1 - (1 - gain)^substeps of the joint error per step; the first
         * step of an edge is reached by teleport (env.py:222-227) */
        const double lag_c = prm->arm_lag_gain > 0.f ? 1.0 - pow(1.0 - (double)prm->arm_lag_gain, (double)prm->arm_lag_substeps) : 1.0;
        const float* a = q0 + (size_t)e * D;
        const float* b = q1 + (size_t)e * D;
        const int n = edge_steps(rb, prm, a, b, dq);
        const int nc = n < max_steps ? n : max_steps;
        const int off = backward ? 0 : 1;
        double* md = (double*)calloc((size_t)2 * W, sizeof(double));
        double* ov = md + W;
        config_t cf;
        memset(&cf, 0, sizeof(cf));
        for (int j = 0; j < D; ++j) q[j] = a[j];
        config_eval(rb, NULL, q, q, &cf);
        double prev[3] = {cf.ee[0], cf.ee[1], cf.ee[2]};
        double eel = 0.0;
        int first = nc;
        if (rigid_mask) memset(rigid_mask + (size_t)e * W * C, 0, sizeof(uint32_t) * W * C);
        if (exceed_mask) memset(exceed_mask + (size_t)e * W * C, 0, sizeof(uint32_t) * W * C);
        for (int i = 0; i < nc; ++i) {
            const int num = i + off;
            for (int j = 0; j < D; ++j) {
                const int circ = rb->circular && rb->circular[j];
                q[j] = (num == n && !circ) ? (double)b[j] : (double)a[j] + dq[j] * ((double)num / (double)n);
                qa[j] = (i == 0 || lag_c >= 1.0) ? q[j] : qa[j] + lag_c * (q[j] - qa[j]);
            }
            config_eval(rb, hd, qa, q, &cf);
            const double dx = cf.ee[0] - prev[0], dy = cf.ee[1] - prev[1], dz = cf.ee[2] - prev[2];
            eel += sqrt(dx * dx + dy * dy + dz * dz);
            prev[0] = cf.ee[0]; prev[1] = cf.ee[1]; prev[2] = cf.ee[2];
            int viol = cf.rigid;
            for (int w = 0; w < W; ++w) {
                unit_t u;
                u.plant_hit = 0; u.exceeded = 0; u.max_defl = 0; u.over = 0;
                if (prm->mode != PMP_MODE_IGNORE_ALL) world_eval(rb, wd, prm, w, &cf, &u, NULL, NULL);
                const int r = cf.rigid || u.plant_hit;
                viol |= r | u.exceeded;
                if (u.max_defl > md[w]) md[w] = u.max_defl;
                ov[w] += u.over;
                const size_t o = ((size_t)e * W + w) * C + (i >> 5);
                if (rigid_mask && r) rigid_mask[o] |= 1u << (i & 31);
                if (exceed_mask && u.exceeded) exceed_mask[o] |= 1u << (i & 31);
            }
            if (viol && first == nc) first = i;
        }
        double osum = 0.0;
        for (int w = 0; w < W; ++w) {
            if (max_defl) max_defl[(size_t)e * W + w] = (float)md[w];
            if (over) over[(size_t)e * W + w] = (float)ov[w];
            osum += ov[w];
        }
        first_hit[e] = first;
        n_steps[e] = n;
        if (ee_len) ee_len[e] = (float)eel;
        if (cost) cost[e] = (float)(eel + (double)prm->alpha * osum / (double)W);
        free(md);
}

static int hw_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

static void* edge_worker(void* arg) {
    job_t* jb = (job_t*)arg;
    for (;;) {
        const int lo = atomic_fetch_add(&jb->next, jb->chunk);
        if (lo >= jb->E) break;
        const int hi = lo + jb->chunk < jb->E ? lo + jb->chunk : jb->E;
        for (int e = lo; e < hi; ++e) one_edge(jb, e);
    }
    return NULL;
}

static void run_threads(job_t* jb, void* (*fn)(void*), int n_threads) {
    if (n_threads <= 0) n_threads = hw_threads();
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { fn(jb); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads; ++i)
        if (pthread_create(&th[started], NULL, fn, jb) == 0) ++started;
    if (started == 0) fn(jb);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* Same contract as pmp_validate_edges_host (include/pmp_edgecheck.h), double accumulators, float outputs.
 * n_threads <= 0: one thread per online core.  Returns 0. */
int pmp_oracle_validate_edges(const pmp_robot_desc* rb, const pmp_world_desc* wd, const pmp_params* prm, const float* q0,
                              const float* q1, int32_t E, int32_t max_steps, int32_t backward, int32_t* first_hit,
                              int32_t* n_steps, float* max_defl, float* over, float* ee_len, float* cost,
                              uint32_t* rigid_mask, uint32_t* exceed_mask, int32_t n_threads, const pmp_hull_desc* hd) {
    job_t jb;
    memset(&jb, 0, sizeof(jb));
    jb.rb = rb; jb.wd = wd; jb.prm = prm; jb.hd = hd; jb.q0 = q0; jb.q1 = q1; jb.E = E; jb.max_steps = max_steps; jb.backward = backward;
    jb.first_hit = first_hit; jb.n_steps = n_steps; jb.max_defl = max_defl; jb.over = over; jb.ee_len = ee_len; jb.cost = cost;
    jb.rigid_mask = rigid_mask; jb.exceed_mask = exceed_mask; jb.chunk = 4;
    atomic_init(&jb.next, 0);
    run_threads(&jb, edge_worker, n_threads);
    return 0;
}

static void one_config(const job_t* jb, int b) {
    const pmp_robot_desc* rb = jb->rb; const pmp_world_desc* wd = jb->wd; const pmp_params* prm = jb->prm;
    const float* q = jb->q; uint8_t* flags = jb->flags; double* defl = jb->defl; double* theta = jb->theta; double* ee = jb->ee;
    const int D = rb->dof, W = wd->n_worlds, P = wd->max_stems;
        double qd[PMP_MAX_DOF];
        config_t cf;
        memset(&cf, 0, sizeof(cf));
        for (int j = 0; j < D; ++j) qd[j] = q[(size_t)b * D + j];
        config_eval(rb, jb->hd, qd, qd, &cf);
        if (ee) for (int k = 0; k < 3; ++k) ee[3 * (size_t)b + k] = cf.ee[k];
        for (int w = 0; w < W; ++w) {
            unit_t u;
            u.plant_hit = 0; u.exceeded = 0; u.max_defl = 0; u.over = 0;
            const size_t bw = (size_t)b * W + w;
            double* dptr = defl ? defl + bw * P : NULL;
            double* tptr = theta ? theta + bw * P * 2 : NULL;
            if (dptr) memset(dptr, 0, sizeof(double) * P);
            if (tptr) memset(tptr, 0, sizeof(double) * P * 2);
            if (prm->mode != PMP_MODE_IGNORE_ALL) {
                real td[PMP_MAX_STEMS], tt[2 * PMP_MAX_STEMS];
                memset(td, 0, sizeof(td)); memset(tt, 0, sizeof(tt));
                world_eval(rb, wd, prm, w, &cf, &u, td, tt);
                const int ns = wd->n_stems[w];
                if (dptr) for (int s2 = 0; s2 < ns; ++s2) dptr[s2] = td[s2];
                if (tptr) for (int s2 = 0; s2 < 2 * ns; ++s2) tptr[s2] = tt[s2];
            }
            flags[bw] = (uint8_t)(((cf.rigid || u.plant_hit) ? PMP_FLAG_RIGID : 0) | (u.exceeded ? PMP_FLAG_EXCEEDED : 0));
        }
}

static void* config_worker(void* arg) {
    job_t* jb = (job_t*)arg;
    for (;;) {
        const int lo = atomic_fetch_add(&jb->next, jb->chunk);
        if (lo >= jb->B) break;
        const int hi = lo + jb->chunk < jb->B ? lo + jb->chunk : jb->B;
        for (int b = lo; b < hi; ++b) one_config(jb, b);
    }
    return NULL;
}

/* Same contract as pmp_check_configs_host; defl / theta / ee in double for tight comparisons. */
int pmp_oracle_check_configs(const pmp_robot_desc* rb, const pmp_world_desc* wd, const pmp_params* prm, const float* q,
                             int32_t B, uint8_t* flags, double* defl, double* theta, double* ee, int32_t n_threads,
                             const pmp_hull_desc* hd) {
    job_t jb;
    memset(&jb, 0, sizeof(jb));
    jb.rb = rb; jb.wd = wd; jb.prm = prm; jb.hd = hd; jb.q = q; jb.B = B; jb.flags = flags; jb.defl = defl; jb.theta = theta; jb.ee = ee;
    jb.chunk = 16;
    atomic_init(&jb.next, 0);
    run_threads(&jb, config_worker, n_threads);
    return 0;
}

int pmp_oracle_max_threads(void) { return hw_threads(); }
