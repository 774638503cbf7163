// Plant-world sampler on the device: one thread draws one world of a fixed topology and writes its
// stem records and base box straight into HBM, so that large-W uncertainty sweeps never round-trip worlds through
// the host.  Restates create_plant_params (rand_gen_plants_programmatically_main.py:71-277: distributions, branch
// placement rule, joint layout) like model/batch_sampler.py, the static balance of model/statics.py (gravity against
// the external-torque springs of env.py:127-144, Newton with a finite-difference Jacobian, fixed step count) and the
// record layout of model/tables.build_stem_records -- all in float64, cast to float32 on store, so that the tables
// equal the host packer's on the same uniform draws (up to the last float32 bit where double sin / cos / atan2 of the
// two platforms differ in their last bit).  Set-up code: clarity over speed (local-memory arrays, no tuning).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pmp_tables.h"

#define PMP_SK_FIRST 0
#define PMP_SK_STACK 1
#define PMP_SK_BRANCH 2
#define PMP_SK_EXT 3

struct PmpSamplerArgs {
    pmp_sampler_desc d;
    int32_t W;
    uint64_t seed;
    const double* uniforms;   // [W, n_uniforms] or null: counter-based generator
    int32_t n_uniforms;
    float* stems;             // [W, P, STRIDE]
    float* boxes;             // [W, 1, BOX_STRIDE]
};

// splitmix64 of (seed, counter): the host mirror is model/batch_sampler.splitmix_uniforms
__device__ __forceinline__ double pmp_sampler_uniform(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct PmpM3 { double m[9]; };
__device__ __forceinline__ PmpM3 pmp_d_mm(const PmpM3& A, const PmpM3& B) {
    PmpM3 O;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            O.m[3 * r + c] = (A.m[3 * r] * B.m[c] + A.m[3 * r + 1] * B.m[3 + c]) + A.m[3 * r + 2] * B.m[6 + c];
    return O;
}
__device__ __forceinline__ void pmp_d_mv(const PmpM3& A, const double* v, double* o) {
    for (int r = 0; r < 3; ++r) o[r] = (A.m[3 * r] * v[0] + A.m[3 * r + 1] * v[1]) + A.m[3 * r + 2] * v[2];
}
__device__ __forceinline__ void pmp_d_mtv(const PmpM3& A, const double* v, double* o) {
    for (int c = 0; c < 3; ++c) o[c] = (A.m[c] * v[0] + A.m[3 + c] * v[1]) + A.m[6 + c] * v[2];
}
__device__ __forceinline__ PmpM3 pmp_d_rxry(double tx, double ty) {
    const double cx = cos(tx), sx = sin(tx), cy = cos(ty), sy = sin(ty);
    PmpM3 M;
    M.m[0] = cy;       M.m[1] = 0.0; M.m[2] = sy;
    M.m[3] = sx * sy;  M.m[4] = cx;  M.m[5] = -sx * cy;
    M.m[6] = -cx * sy; M.m[7] = sx;  M.m[8] = cx * cy;
    return M;
}
__device__ __forceinline__ PmpM3 pmp_d_rx(double tx) {
    const double cx = cos(tx), sx = sin(tx);
    PmpM3 M;
    M.m[0] = 1.0; M.m[1] = 0.0; M.m[2] = 0.0;
    M.m[3] = 0.0; M.m[4] = cx;  M.m[5] = -sx;
    M.m[6] = 0.0; M.m[7] = sx;  M.m[8] = cx;
    return M;
}
// Rz(yaw) Ry(pitch) Rx(roll)  (URDF rpy, model/transforms.rpy_to_matrix)
__device__ __forceinline__ PmpM3 pmp_d_rpy(double r, double p, double y) {
    const double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
    PmpM3 M;
    M.m[0] = cy * cp; M.m[1] = cy * sp * sr - sy * cr; M.m[2] = cy * sp * cr + sy * sr;
    M.m[3] = sy * cp; M.m[4] = sy * sp * sr + cy * cr; M.m[5] = sy * sp * cr - cy * sr;
    M.m[6] = -sp;     M.m[7] = cp * sr;                M.m[8] = cp * cr;
    return M;
}

// generalized force of springs + gravity on every joint (model/statics.static_residual)
__device__ void pmp_static_residual(const pmp_sampler_desc& d, const PmpM3* oR, const double (*ot)[3], const double (*th)[2],
                                    double* F) {
    const int P = d.n_stems;
    PmpM3 R[PMP_SAMPLER_MAX_STEMS];
    double t[PMP_SAMPLER_MAX_STEMS][3], ax[PMP_SAMPLER_MAX_STEMS][3], ay[PMP_SAMPLER_MAX_STEMS][3];
    for (int s = 0; s < P; ++s) {
        PmpM3 Rv;
        const int p = d.parent[s];
        if (p < 0) {
            Rv = oR[s];
            t[s][0] = ot[s][0]; t[s][1] = ot[s][1]; t[s][2] = ot[s][2];
        } else {
            Rv = pmp_d_mm(R[p], oR[s]);
            double x[3];
            pmp_d_mv(R[p], ot[s], x);
            t[s][0] = x[0] + t[p][0]; t[s][1] = x[1] + t[p][1]; t[s][2] = x[2] + t[p][2];
        }
        const PmpM3 Rx = pmp_d_mm(Rv, pmp_d_rx(th[s][0]));
        ax[s][0] = Rv.m[0]; ax[s][1] = Rv.m[3]; ax[s][2] = Rv.m[6];
        ay[s][0] = Rx.m[1]; ay[s][1] = Rx.m[4]; ay[s][2] = Rx.m[7];
        R[s] = pmp_d_mm(Rv, pmp_d_rxry(th[s][0], th[s][1]));
    }
    for (int k = 0; k < 2 * P; ++k) F[k] = 0.0;
    for (int s = 0; s < P; ++s) {
        const double tl[3] = {-d.kp * th[s][0], -d.kp * th[s][1], 0.0};
        double tw[3];
        pmp_d_mv(R[s], tl, tw);
        const double com[3] = {R[s].m[2] * d.com_z + t[s][0], R[s].m[5] * d.com_z + t[s][1], R[s].m[8] * d.com_z + t[s][2]};
        const double fgz = -d.mass * d.g;
        for (int a = s; a >= 0; a = d.parent[a]) {
            // tot = tw + (com - t_a) x (0, 0, fgz)
            const double rx = com[0] - t[a][0], ry = com[1] - t[a][1];
            const double tot[3] = {tw[0] + ry * fgz, tw[1] - rx * fgz, tw[2]};
            F[2 * a] += (ax[a][0] * tot[0] + ax[a][1] * tot[1]) + ax[a][2] * tot[2];
            F[2 * a + 1] += (ay[a][0] * tot[0] + ay[a][1] * tot[1]) + ay[a][2] * tot[2];
        }
    }
}

__global__ void __launch_bounds__(64) pmp_sample_worlds_kernel(const PmpSamplerArgs args) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= args.W) return;
    const pmp_sampler_desc& d = args.d;
    const int P = d.n_stems;
    const int NU = PMP_SAMPLER_HEAD + P * PMP_SAMPLER_PER_STEM;
    auto U = [&](int i) -> double {
        return args.uniforms ? args.uniforms[(size_t)w * args.n_uniforms + i]
                             : pmp_sampler_uniform(args.seed, (uint64_t)w * (uint64_t)NU + (uint64_t)i);
    };
    auto lerp = [](double a, double b, double x) { return a + (b - a) * x; };
    const double sf = d.scaling;
    // ---- draws -> joint origins (model/batch_sampler.draws_from_uniforms) ----
    const double base_x = lerp(d.xy_lo[0], d.xy_hi[0], U(0)), base_y = lerp(d.xy_lo[1], d.xy_hi[1], U(1));
    double bh[3];
    for (int k = 0; k < 3; ++k) bh[k] = lerp(sf * 0.07, sf * 0.15, U(2 + k));
    double hh[PMP_SAMPLER_MAX_STEMS], xyz[PMP_SAMPLER_MAX_STEMS][3], rpy[PMP_SAMPLER_MAX_STEMS][3];
    for (int k = 0; k < P; ++k) {
        const int B = PMP_SAMPLER_HEAD + k * PMP_SAMPLER_PER_STEM;
        hh[k] = lerp(sf * 0.7, sf * 1.0, U(B));
        xyz[k][0] = xyz[k][1] = xyz[k][2] = 0.0;
        rpy[k][0] = rpy[k][1] = rpy[k][2] = 0.0;
        const int ref = d.ref[k];
        if (d.kind[k] == PMP_SK_FIRST) {
            xyz[k][2] = bh[2];
        } else if (d.kind[k] == PMP_SK_STACK) {                 // rand_gen...:237-244
            xyz[k][2] = d.spacing + 2.0 * hh[ref];
            const double ang = lerp(-0.4, 0.4, U(B + 1));
            const bool pitch_side = U(B + 2) < 0.5;
            rpy[k][1] = pitch_side ? ang : 0.0;
            rpy[k][0] = pitch_side ? 0.0 : ang;
            rpy[k][2] = lerp(-0.4, 0.4, U(B + 3));
        } else if (d.kind[k] == PMP_SK_EXT) {                   // rand_gen...:183-188
            xyz[k][2] = d.spacing + 2.0 * hh[ref];
            const double ang = lerp(sf * -0.5, sf * 0.5, U(B + 1));
            const bool roll_side = U(B + 2) < 0.5;
            rpy[k][0] = roll_side ? ang : 0.0;
            rpy[k][1] = roll_side ? 0.0 : ang;
        } else {                                                // branch, rand_gen...:158-162, 196-213
            const double tol = sf * 0.2;
            double x = 0.0, y = 0.0;
            for (int t = 0; t < PMP_SAMPLER_TRIES; ++t) {
                const bool side = U(B + 8 + 3 * t) < 0.5;
                const double fixed = U(B + 9 + 3 * t) < 0.5 ? sf * 0.2 : sf * -0.2;
                const double fr = lerp(0.0, sf * 0.2, U(B + 10 + 3 * t));
                x = side ? fixed : fr;
                y = side ? fr : fixed;
                bool clash = false;
                for (int j = 0; j < k; ++j)
                    if (d.kind[j] == PMP_SK_BRANCH && d.ref[j] == ref)
                        clash |= fabs(x - xyz[j][0]) < tol && fabs(y - xyz[j][1]) < tol;
                if (!clash) break;
            }
            const double yaw = lerp(-0.5, 0.5, U(B + 1)), mag = lerp(0.7, 1.5, U(B + 2));
            const bool along_x = fabs(x) > fabs(y);
            rpy[k][1] = along_x ? (x > 0 ? mag : -mag) : 0.0;
            rpy[k][0] = along_x ? 0.0 : (y > 0 ? -mag : mag);
            rpy[k][2] = yaw;
            xyz[k][0] = x; xyz[k][1] = y;
            xyz[k][2] = d.spacing + U(B + 3) * 2.0 * hh[ref];
        }
    }
    PmpM3 oR[PMP_SAMPLER_MAX_STEMS];
    double ot[PMP_SAMPLER_MAX_STEMS][3];
    for (int k = 0; k < P; ++k) {
        oR[k] = pmp_d_rpy(rpy[k][0], rpy[k][1], rpy[k][2]);
        ot[k][0] = xyz[k][0]; ot[k][1] = xyz[k][1]; ot[k][2] = xyz[k][2];
    }
    ot[0][0] += base_x; ot[0][1] += base_y; ot[0][2] += bh[2];     // the first stem's origin is a world frame

    // ---- rest pose: Newton with a central-difference Jacobian, fixed step count (model/statics.rest_pose_batch) ----
    double th[PMP_SAMPLER_MAX_STEMS][2];
    for (int k = 0; k < P; ++k) th[k][0] = th[k][1] = 0.0;
    const int N = 2 * P;
    if (d.gravity_sag && d.g != 0.0) {
        double F[2 * PMP_SAMPLER_MAX_STEMS], Fp[2 * PMP_SAMPLER_MAX_STEMS], Fm[2 * PMP_SAMPLER_MAX_STEMS];
        double J[2 * PMP_SAMPLER_MAX_STEMS][2 * PMP_SAMPLER_MAX_STEMS + 1];
        const double eps = 1e-6;
        for (int it = 0; it < 10; ++it) {
            pmp_static_residual(d, oR, ot, th, F);
            for (int k = 0; k < N; ++k) {
                double* tk = &th[k >> 1][k & 1];
                const double keep = *tk;
                *tk = keep + eps;
                pmp_static_residual(d, oR, ot, th, Fp);
                *tk = keep - eps;
                pmp_static_residual(d, oR, ot, th, Fm);
                *tk = keep;
                for (int r = 0; r < N; ++r) J[r][k] = (Fp[r] - Fm[r]) / (2 * eps);
            }
            for (int r = 0; r < N; ++r) J[r][N] = -F[r];
            // Gaussian elimination with partial pivoting
            for (int c = 0; c < N; ++c) {
                int piv = c;
                for (int r = c + 1; r < N; ++r) if (fabs(J[r][c]) > fabs(J[piv][c])) piv = r;
                if (piv != c) for (int k = c; k <= N; ++k) { const double tmp = J[c][k]; J[c][k] = J[piv][k]; J[piv][k] = tmp; }
                const double inv = 1.0 / J[c][c];
                for (int r = c + 1; r < N; ++r) {
                    const double f = J[r][c] * inv;
                    for (int k = c; k <= N; ++k) J[r][k] -= f * J[c][k];
                }
            }
            double step[2 * PMP_SAMPLER_MAX_STEMS];
            for (int r = N - 1; r >= 0; --r) {
                double acc = J[r][N];
                for (int k = r + 1; k < N; ++k) acc -= J[r][k] * step[k];
                step[r] = acc / J[r][r];
            }
            for (int k = 0; k < N; ++k) {
                const double st = fmin(fmax(step[k], -0.3), 0.3);
                th[k >> 1][k & 1] = fmin(fmax(th[k >> 1][k & 1] + st, -d.limit), d.limit);
            }
        }
        pmp_static_residual(d, oR, ot, th, F);
        double fmx = 0.0, tmx = 0.0;
        for (int k = 0; k < N; ++k) { fmx = fmax(fmx, fabs(F[k])); tmx = fmax(tmx, fabs(th[k >> 1][k & 1])); }
        // a plant heavier than its springs carry has no small-deflection balance: it keeps theta = 0
        if (!(fmx <= 1e-7 * d.kp) || !(tmx <= 0.75))
            for (int k = 0; k < P; ++k) th[k][0] = th[k][1] = 0.0;
    }

    // ---- stem records (model/tables.build_stem_records) ----
    PmpM3 R0[PMP_SAMPLER_MAX_STEMS], Rr[PMP_SAMPLER_MAX_STEMS];
    double T0[PMP_SAMPLER_MAX_STEMS][3], Tr[PMP_SAMPLER_MAX_STEMS][3];
    const double margin = d.stem_margin;
    for (int s = 0; s < P; ++s) {
        const int p = d.parent[s];
        PmpM3 Rv;
        if (p < 0) {
            R0[s] = oR[s];
            Rv = oR[s];
            for (int k = 0; k < 3; ++k) { T0[s][k] = ot[s][k]; Tr[s][k] = ot[s][k]; }
        } else {
            R0[s] = pmp_d_mm(R0[p], oR[s]);
            double x[3];
            pmp_d_mv(R0[p], ot[s], x);
            for (int k = 0; k < 3; ++k) T0[s][k] = x[k] + T0[p][k];
            Rv = pmp_d_mm(Rr[p], oR[s]);
            pmp_d_mv(Rr[p], ot[s], x);
            for (int k = 0; k < 3; ++k) Tr[s][k] = x[k] + Tr[p][k];
        }
        Rr[s] = pmp_d_mm(Rv, pmp_d_rxry(th[s][0], th[s][1]));
        float* rec = args.stems + ((size_t)w * P + s) * PMP_STEM_STRIDE;
        for (int k = 0; k < PMP_STEM_STRIDE; ++k) rec[k] = 0.f;
        const double zc = d.spacing + hh[s];
        for (int k = 0; k < 9; ++k) rec[PMP_S_RS0 + k] = (float)Rr[s].m[k];
        double b0[3];
        pmp_d_mtv(Rr[s], Tr[s], b0);
        rec[PMP_S_B0] = (float)b0[0]; rec[PMP_S_B0 + 1] = (float)b0[1]; rec[PMP_S_B0 + 2] = (float)(b0[2] + zc);
        const double hw = fmax(d.half_width - margin, 0.0), hhc = fmax(hh[s] - margin, 0.0);
        rec[PMP_S_HW] = (float)hw; rec[PMP_S_HH] = (float)hhc; rec[PMP_S_ZC] = (float)zc; rec[PMP_S_MARGIN] = (float)margin;
        rec[PMP_S_COMZ] = (float)d.com_z; rec[PMP_S_LIM] = (float)d.limit;
        rec[PMP_S_TH0] = (float)th[s][0]; rec[PMP_S_TH0 + 1] = (float)th[s][1];
        rec[PMP_S_BRAD] = (float)((sqrt(2.0) * hw + margin) * 1.000001);
        for (int k = 0; k < 9; ++k) rec[PMP_S_RO + k] = (float)oR[s].m[k];
        for (int k = 0; k < 3; ++k) { rec[PMP_S_TO + k] = (float)ot[s][k]; rec[PMP_S_TV0 + k] = (float)Tr[s][k]; }
        // observer (MOD3): c captured at zero deflection, raw0 = its value in the rest pose
        double c[3], v[3];
        if (p < 0) {
            c[0] = R0[s].m[6]; c[1] = R0[s].m[7]; c[2] = R0[s].m[8];          // R_l0^T ez
            pmp_d_mv(Rr[s], c, v);
        } else {
            const double zp[3] = {R0[p].m[6], R0[p].m[7], R0[p].m[8]};
            pmp_d_mv(R0[s], zp, c);                                           // R_l0 R_p0^T ez
            double l[3];
            pmp_d_mtv(Rr[s], c, l);
            pmp_d_mv(Rr[p], l, v);
        }
        rec[PMP_S_OBSV] = (float)c[0]; rec[PMP_S_OBSV + 1] = (float)c[1]; rec[PMP_S_OBSV + 2] = (float)c[2];
        const double sarg = -v[0];
        const bool gim = fabs(sarg) >= 0.99999;
        const double roll = gim ? 0.0 : atan2(v[1], v[2]);
        const double pitch = gim ? (sarg > 0 ? 1.0 : (sarg < 0 ? -1.0 : 0.0)) * 0.5 * 3.141592653589793
                                 : asin(fmin(fmax(sarg, -1.0), 1.0));
        rec[PMP_S_RAW0] = (float)sqrt(roll * roll + pitch * pitch);
        rec[PMP_S_PSLOT] = __int_as_float(d.pslot[s]);
        rec[PMP_S_OSLOT] = __int_as_float(d.oslot[s]);
        rec[PMP_S_FLAGS] = __int_as_float(d.flags[s]);
    }
    float* bx = args.boxes + (size_t)w * PMP_BOX_STRIDE;
    for (int k = 0; k < PMP_BOX_STRIDE; ++k) bx[k] = 0.f;
    bx[0] = 1.f; bx[4] = 1.f; bx[8] = 1.f;
    bx[9] = (float)base_x; bx[10] = (float)base_y; bx[11] = (float)bh[2];
    bx[12] = (float)bh[0]; bx[13] = (float)bh[1]; bx[14] = (float)bh[2];
}
