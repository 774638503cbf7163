/* Convex-hull distance queries (GJK, float64).                      *** TEST INFRASTRUCTURE ***
 *
 * Restates the narrow phase behind the reference's boolean collision test
 *   plant_motion_planning/pybullet_tools/utils.py:3390-3441  (get_closest_points / pairwise_link_collision /
 *   any_link_pair_collision / pairwise_collision: p.getClosestPoints(distance=0) non-empty)
 * for the shapes the scripts load: PyBullet 3.2.0 (not in /root/reference; Bullet's public source, restated) turns
 * every URDF collision mesh into a btConvexHullShape of the mesh vertices and tests convex pairs with
 * btGjkPairDetector on the shapes *without* margin, then subtracts both collision margins from the distance; a point
 * is reported when that signed distance is <= the query distance (0 here).  This file computes the margin-free
 * distance between two posed vertex sets (= their convex hulls; a box is its 8 corners, a point is one vertex);
 * oracle/hull_oracle.py applies the margins and the pair lists.
 *
 * Algorithm: Gilbert-Johnson-Keerthi on the Minkowski difference A - B with the closest-point-on-simplex
 * sub-algorithm written out by region tests (segment / triangle / tetrahedron), support mapping by exhaustive scan
 * of the vertex set in the shape's local frame.  Terminates when the support point no longer improves the lower
 * bound |v|^2 - v.w <= eps |v|^2, or the simplex encloses the origin (distance 0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    const double* V;   /* [n,3] local vertices */
    int n;
    const double* F;   /* 12 doubles: R row-major (9), t (3): world = R v + t */
} shape_t;

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void sub3(const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* world-frame support point of a posed vertex set in world direction d */
static void support(const shape_t* s, const double* d, double* out) {
    const double* R = s->F;
    const double dl[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
                          R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
    int best = 0;
    double bv = -1e300;
    for (int i = 0; i < s->n; ++i) {
        const double v = s->V[3 * i] * dl[0] + s->V[3 * i + 1] * dl[1] + s->V[3 * i + 2] * dl[2];
        if (v > bv) { bv = v; best = i; }
    }
    const double* p = s->V + 3 * best;
    out[0] = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s->F[9];
    out[1] = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s->F[10];
    out[2] = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s->F[11];
}

typedef struct { int n; double p[4][3]; } simplex_t;

/* closest point to the origin on segment (a, b); reduces the simplex to the supporting feature */
static void closest_segment(simplex_t* s, double* out) {
    const double* a = s->p[0]; const double* b = s->p[1];
    double ab[3]; sub3(b, a, ab);
    const double t = -dot3(a, ab), den = dot3(ab, ab);
    if (t <= 0.0 || den <= 0.0) { memcpy(out, a, 24); s->n = 1; return; }
    if (t >= den) { memcpy(out, b, 24); memcpy(s->p[0], b, 24); s->n = 1; return; }
    const double u = t / den;
    out[0] = a[0] + u * ab[0]; out[1] = a[1] + u * ab[1]; out[2] = a[2] + u * ab[2];
}

/* closest point to the origin on triangle (a, b, c): Voronoi-region tests; writes the reduced feature into
 * (idx, *n_out) as indices into {0,1,2} */
static void closest_triangle_pts(const double* a, const double* b, const double* c, double* out, int* idx, int* n_out) {
    double ab[3], ac[3];
    sub3(b, a, ab); sub3(c, a, ac);
    const double d1 = -dot3(ab, a), d2 = -dot3(ac, a);
    if (d1 <= 0.0 && d2 <= 0.0) { memcpy(out, a, 24); idx[0] = 0; *n_out = 1; return; }
    const double d3 = -dot3(ab, b), d4 = -dot3(ac, b);
    if (d3 >= 0.0 && d4 <= d3) { memcpy(out, b, 24); idx[0] = 1; *n_out = 1; return; }
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        const double v = d1 / (d1 - d3);
        for (int k = 0; k < 3; ++k) out[k] = a[k] + v * ab[k];
        idx[0] = 0; idx[1] = 1; *n_out = 2; return;
    }
    const double d5 = -dot3(ab, c), d6 = -dot3(ac, c);
    if (d6 >= 0.0 && d5 <= d6) { memcpy(out, c, 24); idx[0] = 2; *n_out = 1; return; }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        const double w = d2 / (d2 - d6);
        for (int k = 0; k < 3; ++k) out[k] = a[k] + w * ac[k];
        idx[0] = 0; idx[1] = 2; *n_out = 2; return;
    }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; ++k) out[k] = b[k] + w * (c[k] - b[k]);
        idx[0] = 1; idx[1] = 2; *n_out = 2; return;
    }
    const double den = 1.0 / (va + vb + vc);
    const double v = vb * den, w = vc * den;
    for (int k = 0; k < 3; ++k) out[k] = a[k] + ab[k] * v + ac[k] * w;
    idx[0] = 0; idx[1] = 1; idx[2] = 2; *n_out = 3;
}

static void closest_triangle(simplex_t* s, double* out) {
    int idx[3], n;
    closest_triangle_pts(s->p[0], s->p[1], s->p[2], out, idx, &n);
    double q[3][3];
    for (int i = 0; i < n; ++i) memcpy(q[i], s->p[idx[i]], 24);
    for (int i = 0; i < n; ++i) memcpy(s->p[i], q[i], 24);
    s->n = n;
}

/* origin strictly on the far side of plane (a,b,c) from d?  (degenerate tetrahedron counts as outside) */
static int outside_plane(const double* a, const double* b, const double* c, const double* d) {
    double ab[3], ac[3], n[3], ad[3];
    sub3(b, a, ab); sub3(c, a, ac); cross3(ab, ac, n); sub3(d, a, ad);
    const double sp = -dot3(a, n), sd = dot3(ad, n);
    if (sd == 0.0) return 1;
    return sp * sd < 0.0;
}

/* returns 1 if the tetrahedron encloses the origin.  A (nearly) flat tetrahedron -- four support points of a face
 * with many coplanar vertices -- has no reliable inside: its four faces are then all treated as candidates. */
static int closest_tetrahedron(simplex_t* s, double* out) {
    static const int faces[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    double e1[3], e2[3], e3[3], cr[3];
    sub3(s->p[1], s->p[0], e1); sub3(s->p[2], s->p[0], e2); sub3(s->p[3], s->p[0], e3);
    cross3(e2, e3, cr);
    const double vol = fabs(dot3(e1, cr));
    const double scale = sqrt(dot3(e1, e1) * dot3(e2, e2) * dot3(e3, e3));
    const int flat = !(vol > 1e-9 * scale);
    double best = 1e300;
    int found = 0, bidx[3] = {0, 0, 0}, bn = 0, bf = 0;
    double bp[3] = {0, 0, 0};
    for (int f = 0; f < 4; ++f) {
        const double *a = s->p[faces[f][0]], *b = s->p[faces[f][1]], *c = s->p[faces[f][2]], *d = s->p[faces[f][3]];
        if (!flat && !outside_plane(a, b, c, d)) continue;
        double q[3]; int idx[3], n;
        closest_triangle_pts(a, b, c, q, idx, &n);
        const double dd = dot3(q, q);
        if (dd < best) { best = dd; memcpy(bp, q, 24); memcpy(bidx, idx, sizeof idx); bn = n; bf = f; found = 1; }
    }
    if (!found) return 1;
    double q[3][3];
    for (int i = 0; i < bn; ++i) memcpy(q[i], s->p[faces[bf][bidx[i]]], 24);
    for (int i = 0; i < bn; ++i) memcpy(s->p[i], q[i], 24);
    s->n = bn;
    memcpy(out, bp, 24);
    return 0;
}

/* margin-free distance between the convex hulls of two posed vertex sets (0 when they intersect) */
static double gjk_distance(const shape_t* A, const shape_t* B, int* iters_out) {
    simplex_t s;
    double v[3], a[3], b[3], w[3], nv[3];
    const double d0[3] = {1.0, 0.0, 0.0}, nd0[3] = {-1.0, 0.0, 0.0};
    support(A, d0, a); support(B, nd0, b);
    sub3(a, b, v);
    memcpy(s.p[0], v, 24); s.n = 1;
    double vv = dot3(v, v);
    /* absolute floor: shapes are metres-sized; 1e-22 m^2 is a 1e-11 m gap */
    int it = 0;
    for (; it < 256; ++it) {
        if (vv <= 1e-22) { vv = 0.0; break; }
        nv[0] = -v[0]; nv[1] = -v[1]; nv[2] = -v[2];
        support(A, nv, a); support(B, v, b);
        sub3(a, b, w);
        const double vw = dot3(v, w);
        if (vv - vw <= 1e-13 * vv) break;          /* lower bound met: |v| is the distance */
        int dup = 0;
        for (int i = 0; i < s.n; ++i)
            if (s.p[i][0] == w[0] && s.p[i][1] == w[1] && s.p[i][2] == w[2]) dup = 1;
        if (dup) break;
        memcpy(s.p[s.n], w, 24); s.n++;
        double nvv[3];
        if (s.n == 2) closest_segment(&s, nvv);
        else if (s.n == 3) closest_triangle(&s, nvv);
        else if (closest_tetrahedron(&s, nvv)) { vv = 0.0; break; }
        const double nvvn = dot3(nvv, nvv);
        if (nvvn >= vv) break;                     /* no progress (rounding): keep the previous estimate */
        memcpy(v, nvv, 24); vv = nvvn;
    }
    if (iters_out) *iters_out = it;
    return sqrt(vv);
}

/* ---- exported (not when this file is #included for its static functions) ------------------------------------ */
#ifndef PMP_GJK_STATIC_ONLY

/* one pair */
double hull_gjk_pair(const double* VA, int nA, const double* FA, const double* VB, int nB, const double* FB) {
    shape_t A = {VA, nA, FA}, B = {VB, nB, FB};
    return gjk_distance(&A, &B, 0);
}

/* Batch: B configurations, S shapes (vertex ranges off[s]..off[s+1] of verts), frames [B, S, 12], NP pairs (i, j).
 * out [B, NP] = margin-free hull distance.  Pairs whose bounding spheres (centre c[s] local, radius r[s]) are further
 * apart than `far` are not refined and report that lower bound (>= far).  Threads: OpenMP over configurations. */
int hull_gjk_batch(int B, int S, const int32_t* off, const double* verts, const double* bs_center, const double* bs_radius,
                   const double* frames, int NP, const int32_t* pairs, double far, double* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
        const double* Fb = frames + (size_t)b * S * 12;
        for (int p = 0; p < NP; ++p) {
            const int i = pairs[2 * p], j = pairs[2 * p + 1];
            const double *Fi = Fb + 12 * i, *Fj = Fb + 12 * j;
            double ci[3], cj[3];
            for (int k = 0; k < 3; ++k) {
                ci[k] = Fi[3 * k] * bs_center[3 * i] + Fi[3 * k + 1] * bs_center[3 * i + 1] + Fi[3 * k + 2] * bs_center[3 * i + 2] + Fi[9 + k];
                cj[k] = Fj[3 * k] * bs_center[3 * j] + Fj[3 * k + 1] * bs_center[3 * j + 1] + Fj[3 * k + 2] * bs_center[3 * j + 2] + Fj[9 + k];
            }
            double d[3]; sub3(ci, cj, d);
            const double lb = sqrt(dot3(d, d)) - bs_radius[i] - bs_radius[j];
            if (lb > far) { out[(size_t)b * NP + p] = lb; continue; }
            shape_t A = {verts + 3 * off[i], off[i + 1] - off[i], Fi}, Bs = {verts + 3 * off[j], off[j + 1] - off[j], Fj};
            out[(size_t)b * NP + p] = gjk_distance(&A, &Bs, 0);
        }
    }
    return 0;
}
#endif
