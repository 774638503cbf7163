// C-ABI of libpmp_edgecheck.so (see include/pmp_edgecheck.h).  Host glue only: table upload,
// kernel selection and launch.  No CPU compute path exists here -- without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <type_traits>
#include <vector>

#include "pmp_launch.h"
#include "pmp_hull.cuh"
#include "pmp_tree.cuh"
#include "pmp_sampler.cuh"

struct pmp_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::string err;
    bool has_robot = false, has_worlds = false, has_params = false;
    PmpRobotTab h_robot;
    PmpLinkTab h_links;
    PmpRobotTab* d_robot = nullptr;
    PmpLinkTab* d_links = nullptr;
    // convex hulls of the arm's collision meshes (pmp_set_robot_hulls) and the pre-pass scratch (one bit per step)
    bool has_hulls = false;
    PmpHullTab h_hulls;
    PmpHullTab* d_hulls = nullptr;
    float4* d_hull_verts = nullptr;
    void* d_rigid_scratch = nullptr;
    size_t rigid_scratch_bytes = 0;
    uint8_t* d_pair_tables = nullptr;      // clearance tables of the two-joint hull pairs (PmpHullTab::tab_*)
    uint32_t self_mask_spheres[PMP_MAX_SPHERES];   // the sphere model's own pair masks (hull mode edits the device copy)
    float* d_stems = nullptr;
    int32_t* d_nstems = nullptr;
    float* d_boxes = nullptr;
    int32_t* d_nboxes = nullptr;
    int n_worlds = 0, max_stems = 0, max_boxes = 0, n_slots = 1;
    PmpKernelParams prm;
    pmp_launch_info info;
    // scratch for the host-buffer entry points
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;        // second stream of the chunked host entry point (copies overlap kernels)
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned staging for small host calls (the per-step collision_fn adapter): one H2D, one kernel, one D2H
    void* h_stage = nullptr;
    size_t stage_bytes = 0;
    // scratch of the tree entry points (nearest-node partials, gathered nodes, per-extension edge results)
    int32_t* d_work_counter = nullptr;     // edge kernel: ring of 64 'next unclaimed edge' counters, one per launch
    unsigned work_slot = 0;
    void* d_tree_scratch = nullptr;
    size_t tree_scratch_bytes = 0;
    // per-kernel timing (pmp_set_kernel_timing): events around the hull pre-pass [0, 1] and the edge kernel [2, 3]
    bool timing = false, timed_rigid = false, timed_edge = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static std::string g_create_err;

static int fail(pmp_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, PMP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

extern "C" const char* pmp_last_error(pmp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int pmp_create(pmp_ctx** out, int device) {
    pmp_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, PMP_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(nullptr, PMP_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, PMP_E_INVALID, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, PMP_E_CUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, PMP_E_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PMP_E_CUDA, "this library carries sm_100a code only (Blackwell B200 required)");
    ctx = new pmp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    memset(&ctx->info, 0, sizeof(ctx->info));
    memset(&ctx->prm, 0, sizeof(ctx->prm));
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PMP_E_CUDA, "cudaStreamCreate failed");
    }
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, PMP_E_CUDA, "cudaStreamCreate failed");
    }
    *out = ctx;
    return PMP_OK;
}

static void free_worlds(pmp_ctx* ctx) {
    cudaFree(ctx->d_stems); cudaFree(ctx->d_nstems); cudaFree(ctx->d_boxes); cudaFree(ctx->d_nboxes);
    ctx->d_stems = nullptr; ctx->d_nstems = nullptr; ctx->d_boxes = nullptr; ctx->d_nboxes = nullptr;
}

extern "C" int pmp_destroy(pmp_ctx* ctx) {
    if (!ctx) return PMP_OK;
    cudaSetDevice(ctx->device);
    free_worlds(ctx);
    cudaFree(ctx->d_hulls); cudaFree(ctx->d_hull_verts); cudaFree(ctx->d_rigid_scratch); cudaFree(ctx->d_pair_tables);
    cudaFree(ctx->d_robot); cudaFree(ctx->d_links); cudaFree(ctx->d_scratch); cudaFree(ctx->d_tree_scratch); cudaFree(ctx->d_work_counter);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    delete ctx;
    return PMP_OK;
}

extern "C" int pmp_set_robot(pmp_ctx* ctx, const pmp_robot_desc* r) {
    if (!ctx || !r) return fail(ctx, PMP_E_INVALID, "null argument");
    if (r->dof < 1 || r->dof > PMP_MAX_DOF) return fail(ctx, PMP_E_INVALID, "dof out of range");
    if (r->n_spheres < 1 || r->n_spheres > PMP_MAX_SPHERES) return fail(ctx, PMP_E_INVALID, "n_spheres out of range");
    if (r->n_links < 0 || r->n_links > PMP_MAX_LINKS) return fail(ctx, PMP_E_INVALID, "n_links out of range");
    if (r->n_fixed < 0 || r->n_fixed > PMP_MAX_FIXED) return fail(ctx, PMP_E_INVALID, "n_fixed out of range");
    if (r->ee_joint < 0 || r->ee_joint >= r->dof) return fail(ctx, PMP_E_INVALID, "ee_joint out of range");
    CK(cudaSetDevice(ctx->device));
    // the tables are built and validated in locals and committed only on success: a rejected description leaves the
    // context (host mirrors and device tables) exactly as it was
    static_assert(std::is_trivially_copyable<PmpRobotTab>::value && std::is_trivially_copyable<PmpLinkTab>::value, "POD tables");
    PmpRobotTab t;
    memset(&t, 0, sizeof(t));
    const int D = r->dof, A = r->n_spheres, F = r->n_fixed, L = r->n_links;
    memcpy(t.joint_C, r->joint_C, sizeof(float) * 9 * D);
    memcpy(t.joint_A, r->joint_A, sizeof(float) * 9 * D);
    memcpy(t.joint_B, r->joint_B, sizeof(float) * 9 * D);
    memcpy(t.joint_t, r->joint_t, sizeof(float) * 3 * D);
    memcpy(t.lower, r->lower, sizeof(float) * D);
    memcpy(t.upper, r->upper, sizeof(float) * D);
    memcpy(t.base_R, r->base_R, sizeof(float) * 9);
    memcpy(t.base_t, r->base_t, sizeof(float) * 3);
    memcpy(t.sph_center, r->sph_center, sizeof(float) * 3 * A);
    memcpy(t.sph_radius, r->sph_radius, sizeof(float) * A);
    memcpy(t.sph_radius_outer, r->sph_radius, sizeof(float) * A);
    memcpy(t.self_mask, r->self_mask, sizeof(uint32_t) * A);
    memcpy(t.sph_joint, r->sph_joint, sizeof(int32_t) * A);
    // padding spheres: parked at PMP_FAR by the kernels; radius + stem margin stays negative and small (the contact sums
    // saturate distances at 1 m, pmp_kernels.cuh pmp_outside)
    for (int a = A; a < PMP_MAX_SPHERES; ++a) { t.sph_joint[a] = -1; t.sph_radius[a] = -0.5f; t.sph_radius_outer[a] = -1.0e3f; }
    for (int a = 0; a < A; ++a)
        if (!(t.sph_radius[a] > 0.f && t.sph_radius[a] <= 0.5f)) return fail(ctx, PMP_E_INVALID, "sph_radius outside (0, 0.5] m");
    for (int a = 0; a < A; ++a) {
        if (t.sph_joint[a] < 0 || t.sph_joint[a] >= D) return fail(ctx, PMP_E_INVALID, "sph_joint out of range");
        if (a && t.sph_joint[a] < t.sph_joint[a - 1]) return fail(ctx, PMP_E_INVALID, "spheres must be sorted by joint");
    }
    if (F) {
        memcpy(t.fixed_R, r->fixed_R, sizeof(float) * 9 * F);
        memcpy(t.fixed_t, r->fixed_t, sizeof(float) * 3 * F);
        memcpy(t.fixed_half, r->fixed_half, sizeof(float) * 3 * F);
        memcpy(t.fixed_mask, r->fixed_mask, sizeof(uint32_t) * F);
        memcpy(t.fixed_kind, r->fixed_kind, sizeof(int32_t) * F);
    }
    memcpy(t.ee_local, r->ee_local, sizeof(float) * 3);
    t.ee_joint = r->ee_joint; t.dof = D; t.n_spheres = A; t.n_fixed = F;
    t.circular_mask = 0;
    for (int j = 0; j < D; ++j) if (r->circular && r->circular[j]) t.circular_mask |= (1u << j);
    PmpLinkTab lt;
    memset(&lt, 0, sizeof(lt));
    lt.n_links = L;
    if (L) {
        memcpy(lt.link_joint, r->link_joint, sizeof(int32_t) * L);
        memcpy(lt.link_R, r->link_R, sizeof(float) * 9 * L);
        memcpy(lt.link_t, r->link_t, sizeof(float) * 3 * L);
        memcpy(lt.link_com, r->link_com, sizeof(float) * 3 * L);
        for (int l = 0; l < L; ++l)
            if (lt.link_joint[l] < -1 || lt.link_joint[l] >= D) return fail(ctx, PMP_E_INVALID, "link_joint out of range");
    }
    for (int f = 0; f < F; ++f)
        if (t.fixed_kind[f] < PMP_PRIM_BOX || t.fixed_kind[f] > PMP_PRIM_SPHERE) return fail(ctx, PMP_E_INVALID, "unknown fixed_kind");
    if (!ctx->d_robot) CK(cudaMalloc(&ctx->d_robot, sizeof(PmpRobotTab)));
    if (!ctx->d_links) CK(cudaMalloc(&ctx->d_links, sizeof(PmpLinkTab)));
    CK(cudaMemcpy(ctx->d_robot, &t, sizeof(t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_links, &lt, sizeof(lt), cudaMemcpyHostToDevice));
    ctx->h_robot = t;
    ctx->h_links = lt;
    memcpy(ctx->self_mask_spheres, t.self_mask, sizeof(t.self_mask));
    ctx->has_robot = true;
    ctx->has_hulls = false;      // hulls describe one robot: a new robot starts on its sphere model
    return PMP_OK;
}

extern "C" int pmp_set_robot_hulls(pmp_ctx* ctx, const pmp_hull_desc* h) {
    if (!ctx) return PMP_E_INVALID;
    if (!ctx->has_robot) return fail(ctx, PMP_E_STATE, "pmp_set_robot must come first");
    CK(cudaSetDevice(ctx->device));
    PmpRobotTab t = ctx->h_robot;
    const int A = t.n_spheres, D = t.dof;
    memcpy(t.self_mask, ctx->self_mask_spheres, sizeof(t.self_mask));
    if (!h || h->n_hulls == 0) {
        t.has_hulls = 0;
        for (int a = 0; a < A; ++a) { t.sph_radius_outer[a] = t.sph_radius[a]; t.sph_radius_inner[a] = 0.f; }
        CK(cudaMemcpy(ctx->d_robot, &t, sizeof(t), cudaMemcpyHostToDevice));
        ctx->h_robot = t;
        ctx->has_hulls = false;
        return PMP_OK;
    }
    if (h->n_hulls < 0 || h->n_hulls > PMP_MAX_HULLS) return fail(ctx, PMP_E_INVALID, "n_hulls out of range");
    if (h->n_pairs < 0 || h->n_pairs > PMP_MAX_HULL_PAIRS) return fail(ctx, PMP_E_INVALID, "n_pairs out of range");
    if (h->n_fixed_exact < 0 || h->n_fixed_exact > t.n_fixed) return fail(ctx, PMP_E_INVALID, "n_fixed_exact out of range");
    if (!h->hull_joint || !h->hull_off || !h->hull_verts || !h->hull_margin || !h->sph_hull || !h->sph_radius_outer ||
        (h->n_pairs && !h->pairs))
        return fail(ctx, PMP_E_INVALID, "null array in pmp_hull_desc");
    PmpHullTab ht;
    memset(&ht, 0, sizeof(ht));
    const int H = h->n_hulls;
    ht.n_hulls = H; ht.n_pairs = h->n_pairs; ht.n_fixed_exact = h->n_fixed_exact; ht.fixed_margin = h->fixed_margin;
    if (!(h->fixed_margin >= 0.f)) return fail(ctx, PMP_E_INVALID, "fixed_margin must be >= 0");
    if (h->hull_off[0] != 0) return fail(ctx, PMP_E_INVALID, "hull_off[0] must be 0");
    for (int i = 0; i < H; ++i) {
        if (h->hull_joint[i] < 0 || h->hull_joint[i] >= D) return fail(ctx, PMP_E_INVALID, "hull_joint out of range");
        if (h->hull_off[i + 1] <= h->hull_off[i]) return fail(ctx, PMP_E_INVALID, "every hull needs at least one vertex");
        if (!(h->hull_margin[i] >= 0.f)) return fail(ctx, PMP_E_INVALID, "hull_margin must be >= 0");
        ht.hull_joint[i] = h->hull_joint[i]; ht.hull_off[i] = h->hull_off[i]; ht.hull_margin[i] = h->hull_margin[i];
        ht.sph_lo[i] = 0; ht.sph_hi[i] = 0;
    }
    ht.hull_off[H] = h->hull_off[H];
    const int NV = h->hull_off[H];
    if (NV > PMP_MAX_HULL_VERTS) return fail(ctx, PMP_E_INVALID, "too many hull vertices");
    for (int p = 0; p < 2 * h->n_pairs; ++p) {
        if (h->pairs[p] < 0 || h->pairs[p] >= H) return fail(ctx, PMP_E_INVALID, "hull pair index out of range");
        ht.pairs[p] = h->pairs[p];
    }
    for (int f = 0; f < h->n_fixed_exact; ++f)
        if (t.fixed_kind[f] != PMP_PRIM_BOX && t.fixed_kind[f] != PMP_PRIM_SPHERE)
            return fail(ctx, PMP_E_INVALID, "the exact hull test supports box and sphere obstacles");
    // the spheres of a hull must be one contiguous run (they are sorted by joint and a hull rides on one joint)
    for (int i = 0; i < H; ++i) {
        int lo = -1, hi = -1;
        for (int a = 0; a < A; ++a) {
            if (h->sph_hull[a] == i) {
                if (lo < 0) lo = a;
                else if (hi != a) return fail(ctx, PMP_E_INVALID, "spheres of a hull must be contiguous");
                hi = a + 1;
            }
        }
        if (lo < 0) return fail(ctx, PMP_E_INVALID, "every hull needs at least one covering sphere");
        ht.sph_lo[i] = lo; ht.sph_hi[i] = hi;
    }
    for (int a = 0; a < A; ++a) {
        if (h->sph_hull[a] < -1 || h->sph_hull[a] >= H) return fail(ctx, PMP_E_INVALID, "sph_hull out of range");
        if (!(h->sph_radius_outer[a] >= t.sph_radius[a])) return fail(ctx, PMP_E_INVALID, "sph_radius_outer must be >= sph_radius");
        t.sph_radius_outer[a] = h->sph_radius_outer[a];
        t.sph_radius_inner[a] = 0.f;
        if (h->sph_radius_inner && h->sph_hull[a] >= 0) {
            const float ri = h->sph_radius_inner[a];
            if (!(ri >= 0.f) || ri > h->sph_radius_outer[a]) return fail(ctx, PMP_E_INVALID, "sph_radius_inner must be in [0, sph_radius_outer]");
            // 2e-6 m inside the true ball: float32 rounding of the posed centres never turns a miss into a hit
            t.sph_radius_inner[a] = ri > 2.0e-6f ? ri - 2.0e-6f : 0.f;
        }
    }
    std::vector<float4> verts((size_t)NV);
    for (int i = 0; i < NV; ++i) verts[i] = make_float4(h->hull_verts[3 * i], h->hull_verts[3 * i + 1], h->hull_verts[3 * i + 2], 0.f);
    t.has_hulls = 1;

    // ---- clearance tables for hull pairs whose relative pose depends on exactly two joints (a = j + 1, b = j + 2) ----
    // Lipschitz bound of the hull distance: a point of hull 2 moves at most rho_b per radian of q_b (its distance from
    // axis b) and at most rho_a per radian of q_a (distance of joint b's origin from axis a + the hull's radius).
    struct TabPlan { int pair, na, nb; float band; size_t off; };
    std::vector<TabPlan> plans;
    size_t tab_bytes = 0;
    for (int p = 0; p < h->n_pairs; ++p) {
        ht.tab_off[p] = -1;
        const int h1 = ht.pairs[2 * p], h2 = ht.pairs[2 * p + 1];
        const int j1 = ht.hull_joint[h1], j2 = ht.hull_joint[h2];
        if (j2 - j1 != 2) continue;
        const int ja = j1 + 1, jb = j2;
        if (((t.circular_mask >> ja) & 1u) || ((t.circular_mask >> jb) & 1u)) continue;
        const double ra = (double)t.upper[ja] - t.lower[ja], rb = (double)t.upper[jb] - t.lower[jb];
        if (!(ra > 0 && rb > 0 && ra < 13 && rb < 13)) continue;
        auto axis_of = [&](int j, double* a) {      // R0^T C = a a^T: take its largest column
            double R0[9], M[9];
            for (int k = 0; k < 9; ++k) R0[k] = (double)t.joint_C[9 * j + k] + t.joint_A[9 * j + k];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
                M[3 * r + c] = 0;
                for (int k = 0; k < 3; ++k) M[3 * r + c] += R0[3 * k + r] * t.joint_C[9 * j + 3 * k + c];
            }
            int kk = M[0] >= M[4] ? (M[0] >= M[8] ? 0 : 2) : (M[4] >= M[8] ? 1 : 2);
            const double nrm = sqrt(M[4 * kk] > 0 ? M[4 * kk] : 1.0);
            for (int r = 0; r < 3; ++r) a[r] = M[3 * r + kk] / nrm;
            // axis in the joint's own (post-origin) frame; as a direction in the PARENT frame it is R0 a
        };
        double ab[3], aa[3];
        axis_of(jb, ab);
        axis_of(ja, aa);
        double rho_b = 0, rmax = 0;
        for (int i = ht.hull_off[h2]; i < ht.hull_off[h2 + 1]; ++i) {
            const double v[3] = {h->hull_verts[3 * i], h->hull_verts[3 * i + 1], h->hull_verts[3 * i + 2]};
            const double along = v[0] * ab[0] + v[1] * ab[1] + v[2] * ab[2];
            const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            rho_b = fmax(rho_b, sqrt(fmax(n2 - along * along, 0.0)));
            rmax = fmax(rmax, sqrt(n2));
        }
        // joint b's origin in the frame of joint a (after a's rotation): joint_t[b]; axis a in that frame is `aa`
        const double tb[3] = {t.joint_t[3 * jb], t.joint_t[3 * jb + 1], t.joint_t[3 * jb + 2]};
        const double alongt = tb[0] * aa[0] + tb[1] * aa[1] + tb[2] * aa[2];
        const double rho_a = sqrt(fmax(tb[0] * tb[0] + tb[1] * tb[1] + tb[2] * tb[2] - alongt * alongt, 0.0)) + rmax;
        const double msum = (double)ht.hull_margin[h1] + ht.hull_margin[h2];
        double target = 0.35 * msum;                    // Lipschitz half of the band; float32 slack is added below
        if (!(target > 1e-5)) continue;
        double da = target / fmax(rho_a, 1e-3), db = target / fmax(rho_b, 1e-3);
        long long na = (long long)ceil(ra / da) + 1, nb = (long long)ceil(rb / db) + 1;
        if (na * nb > (1ll << 21)) continue;              // too fine to be worth a table
        da = ra / (double)(na - 1); db = rb / (double)(nb - 1);
        const float band = (float)(0.5 * (rho_a * da + rho_b * db) + 3.0e-5);
        if (!(band < msum)) continue;
        ht.tab_off[p] = (int32_t)tab_bytes; ht.tab_ja[p] = ja; ht.tab_jb[p] = jb; ht.tab_na[p] = (int32_t)na; ht.tab_nb[p] = (int32_t)nb;
        ht.tab_lo_a[p] = t.lower[ja]; ht.tab_inv_a[p] = (float)(1.0 / da);
        ht.tab_lo_b[p] = t.lower[jb]; ht.tab_inv_b[p] = (float)(1.0 / db);
        plans.push_back({p, (int)na, (int)nb, band, tab_bytes});
        tab_bytes += (size_t)(na * nb);
        tab_bytes = (tab_bytes + 15) & ~(size_t)15;
        // the covering spheres no longer cull this pair: drop its sphere pairs from the broad-phase mask
        for (int a = ht.sph_lo[h1]; a < ht.sph_hi[h1]; ++a)
            for (int b = ht.sph_lo[h2]; b < ht.sph_hi[h2]; ++b) t.self_mask[a < b ? a : b] &= ~(1u << (a < b ? b : a));
    }

    cudaFree(ctx->d_hull_verts); ctx->d_hull_verts = nullptr;
    cudaFree(ctx->d_pair_tables); ctx->d_pair_tables = nullptr;
    CK(cudaMalloc(&ctx->d_hull_verts, sizeof(float4) * (size_t)NV));
    if (tab_bytes) CK(cudaMalloc(&ctx->d_pair_tables, tab_bytes));
    if (!ctx->d_hulls) CK(cudaMalloc(&ctx->d_hulls, sizeof(PmpHullTab)));
    CK(cudaMemcpy(ctx->d_hull_verts, verts.data(), sizeof(float4) * (size_t)NV, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_hulls, &ht, sizeof(ht), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_robot, &t, sizeof(t), cudaMemcpyHostToDevice));
    for (const TabPlan& pl : plans) {
        PmpPairTableArgs a;
        a.robot = ctx->d_robot; a.hulls = ctx->d_hulls; a.hull_verts = ctx->d_hull_verts;
        a.table = ctx->d_pair_tables + pl.off; a.pair = pl.pair; a.band = pl.band;
        pmp_pair_table_kernel<<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(a);
        CK(cudaGetLastError());
        ctx->info.kernel_launches++;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_robot = t;
    ctx->h_hulls = ht;
    ctx->has_hulls = true;
    return PMP_OK;
}

// hull pre-pass: limits | self | fixed of every step (edge mode) or configuration, on `stream`
static int launch_rigid(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps, int32_t backward,
                        uint32_t* rigid_pre, const float* q, int32_t B, uint8_t* rigid_flags, cudaStream_t st,
                        bool first_bit_only = false) {
    PmpRigidArgs a;
    a.first_bit_only = first_bit_only ? 1 : 0;
    a.robot = ctx->d_robot; a.hulls = ctx->has_hulls ? ctx->d_hulls : nullptr; a.hull_verts = ctx->d_hull_verts; a.prm = ctx->prm;
    a.pair_tables = ctx->d_pair_tables;
    a.q0 = q0; a.q1 = q1; a.E = E; a.max_steps = max_steps; a.backward = backward ? 1 : 0; a.rigid_pre = rigid_pre;
    a.q = q; a.B = B; a.rigid_flags = rigid_flags;
    // configuration mode: as few configurations per warp item as keeps every SM busy
    a.cfgs_per_item = 32;
    if (q) {
        const long long warps = (long long)ctx->sm_count * PMP_RIGID_MINB * PMP_RIGID_WARPS;
        a.cfgs_per_item = B >= 32 * warps ? 32 : (B >= 8 * warps ? 8 : (B >= 2 * warps ? 2 : 1));
    }
    // edge mode: short items for small batches (see PmpRigidArgs::steps_per_item)
    a.steps_per_item = 32;
    if (!q) {
        const long long warps = (long long)ctx->sm_count * PMP_RIGID_MINB * PMP_RIGID_WARPS;
        const long long words = (long long)E * ((max_steps + 31) / 32);
        a.steps_per_item = words >= 8 * warps ? 32 : (words >= 2 * warps ? 8 : 4);
        if (a.steps_per_item < 32) CK(cudaMemsetAsync(rigid_pre, 0, sizeof(uint32_t) * (size_t)words, st));
    }
    const long long items = q ? (B + a.cfgs_per_item - 1) / a.cfgs_per_item
                              : (long long)E * ((max_steps + 31) / 32) * (32 / a.steps_per_item);
    long long grid = (items + PMP_RIGID_WARPS - 1) / PMP_RIGID_WARPS;
    const long long cap = (long long)ctx->sm_count * PMP_RIGID_MINB;            // the resident CTAs; items are claimed dynamically
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    a.work_counter = nullptr;
    if (!q) {               // edge mode: items differ by an order of magnitude in cost, claim them dynamically
        if (!ctx->d_work_counter) CK(cudaMalloc(&ctx->d_work_counter, 64 * sizeof(int32_t)));
        a.work_counter = ctx->d_work_counter + (ctx->work_slot++ & 63u);
        CK(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    }                       // configuration mode strides statically over the grid: no counter to clear on the scalar path
    const size_t smem = sizeof(float) * (size_t)PMP_RIGID_WARPS * 3 * ctx->h_robot.n_spheres * 32;   // <= 48 KB
    static int optin[16] = {0};
    if (!optin[ctx->device & 15]) {
        CK(cudaFuncSetAttribute(pmp_rigid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        optin[ctx->device & 15] = 1;
    }
    pmp_rigid_kernel<<<(int)grid, 32 * PMP_RIGID_WARPS, smem, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return PMP_OK;
}

static int ensure_rigid_scratch(pmp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->rigid_scratch_bytes) return PMP_OK;
    cudaFree(ctx->d_rigid_scratch);
    ctx->d_rigid_scratch = nullptr; ctx->rigid_scratch_bytes = 0;
    const size_t want = bytes + bytes / 4;
    CK(cudaMalloc(&ctx->d_rigid_scratch, want));
    ctx->rigid_scratch_bytes = want;
    return PMP_OK;
}

extern "C" int pmp_set_worlds(pmp_ctx* ctx, const pmp_world_desc* w) {
    if (!ctx || !w) return fail(ctx, PMP_E_INVALID, "null argument");
    if (w->n_worlds < 1) return fail(ctx, PMP_E_INVALID, "n_worlds < 1");
    if (w->max_stems < 1 || w->max_stems > PMP_MAX_STEMS) return fail(ctx, PMP_E_INVALID, "max_stems out of range");
    if (w->max_boxes < 1 || w->max_boxes > PMP_MAX_WORLD_BOXES) return fail(ctx, PMP_E_INVALID, "max_boxes out of range");
    if (w->n_slots < 1 || w->n_slots > 4)
        return fail(ctx, PMP_E_INVALID, "n_slots must be 1..4 (plants deeper than 4 live parent frames are not compiled in)");
    const int W = w->n_worlds, P = w->max_stems;
    for (int i = 0; i < W; ++i) {
        if (w->n_stems[i] < 0 || w->n_stems[i] > P) return fail(ctx, PMP_E_INVALID, "n_stems[w] out of range");
        if (w->n_boxes[i] < 0 || w->n_boxes[i] > w->max_boxes) return fail(ctx, PMP_E_INVALID, "n_boxes[w] out of range");
        for (int s = 0; s < w->n_stems[i]; ++s) {
            const float* rec = w->stems + ((size_t)i * P + s) * PMP_STEM_STRIDE;
            int32_t ps, os;
            memcpy(&ps, rec + PMP_S_PSLOT, 4);
            memcpy(&os, rec + PMP_S_OSLOT, 4);
            if (ps < -1 || ps >= w->n_slots || os < -1 || os >= w->n_slots)
                return fail(ctx, PMP_E_INVALID, "stem frame slot out of range");
            // sphere radius (<= 0.5) + margin < 1 m: the contact sums saturate distances there (pmp_outside)
            if (!(rec[PMP_S_MARGIN] >= 0.f && rec[PMP_S_MARGIN] <= 0.4f))
                return fail(ctx, PMP_E_INVALID, "stem margin outside [0, 0.4] m");
        }
    }
    CK(cudaSetDevice(ctx->device));
    free_worlds(ctx);
    const size_t sb = sizeof(float) * (size_t)W * P * PMP_STEM_STRIDE;
    const size_t bb = sizeof(float) * (size_t)W * w->max_boxes * PMP_BOX_STRIDE;
    CK(cudaMalloc(&ctx->d_stems, sb));
    CK(cudaMalloc(&ctx->d_nstems, sizeof(int32_t) * W));
    CK(cudaMalloc(&ctx->d_boxes, bb));
    CK(cudaMalloc(&ctx->d_nboxes, sizeof(int32_t) * W));
    CK(cudaMemcpy(ctx->d_stems, w->stems, sb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_nstems, w->n_stems, sizeof(int32_t) * W, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_boxes, w->boxes, bb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_nboxes, w->n_boxes, sizeof(int32_t) * W, cudaMemcpyHostToDevice));
    ctx->n_worlds = W; ctx->max_stems = P; ctx->max_boxes = w->max_boxes; ctx->n_slots = w->n_slots;
    ctx->has_worlds = true;
    return PMP_OK;
}

extern "C" int pmp_sample_worlds(pmp_ctx* ctx, const pmp_sampler_desc* d, int32_t W, uint64_t seed, const double* uniforms_dev,
                                 int32_t n_uniforms, float* stems_dev, float* boxes_dev, void* stream) {
    if (!ctx || !d) return fail(ctx, PMP_E_INVALID, "null argument");
    if (W < 1 || !stems_dev || !boxes_dev) return fail(ctx, PMP_E_INVALID, "W >= 1, stems and boxes are required");
    const int P = d->n_stems;
    if (P < 1 || P > PMP_SAMPLER_MAX_STEMS) return fail(ctx, PMP_E_INVALID, "n_stems out of range for the device sampler");
    if (d->n_slots < 1 || d->n_slots > 4) return fail(ctx, PMP_E_INVALID, "n_slots must be 1..4");
    for (int s = 0; s < P; ++s) {
        if (d->parent[s] < -1 || d->parent[s] >= s) return fail(ctx, PMP_E_INVALID, "stems must be listed parents first");
        if ((s == 0) != (d->kind[s] == PMP_SK_FIRST) || d->kind[s] < 0 || d->kind[s] > PMP_SK_EXT)
            return fail(ctx, PMP_E_INVALID, "bad stem kind");
        if (s > 0 && (d->ref[s] < 0 || d->ref[s] >= s)) return fail(ctx, PMP_E_INVALID, "bad stem ref");
        if (d->pslot[s] < -1 || d->pslot[s] >= d->n_slots || d->oslot[s] < -1 || d->oslot[s] >= d->n_slots)
            return fail(ctx, PMP_E_INVALID, "stem frame slot out of range");
    }
    if (!(d->stem_margin >= 0.0 && d->stem_margin <= 0.4)) return fail(ctx, PMP_E_INVALID, "stem_margin outside [0, 0.4] m");
    if (uniforms_dev && n_uniforms < PMP_SAMPLER_HEAD + P * PMP_SAMPLER_PER_STEM)
        return fail(ctx, PMP_E_INVALID, "too few uniforms per world");
    CK(cudaSetDevice(ctx->device));
    PmpSamplerArgs a;
    a.d = *d; a.W = W; a.seed = seed; a.uniforms = uniforms_dev; a.n_uniforms = n_uniforms; a.stems = stems_dev; a.boxes = boxes_dev;
    pmp_sample_worlds_kernel<<<(W + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return PMP_OK;
}

static __global__ void pmp_fill_i32_kernel(int32_t* p, int32_t n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int pmp_set_worlds_dev(pmp_ctx* ctx, int32_t W, int32_t n_stems, int32_t n_boxes, int32_t n_slots,
                                  const float* stems_dev, const float* boxes_dev, void* stream) {
    if (!ctx || !stems_dev || !boxes_dev) return fail(ctx, PMP_E_INVALID, "null argument");
    if (W < 1) return fail(ctx, PMP_E_INVALID, "n_worlds < 1");
    if (n_stems < 1 || n_stems > PMP_MAX_STEMS) return fail(ctx, PMP_E_INVALID, "n_stems out of range");
    if (n_boxes < 1 || n_boxes > PMP_MAX_WORLD_BOXES) return fail(ctx, PMP_E_INVALID, "n_boxes out of range");
    if (n_slots < 1 || n_slots > 4) return fail(ctx, PMP_E_INVALID, "n_slots must be 1..4");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaStreamSynchronize(st));          // the old tables may still be in use by launches on this stream
    free_worlds(ctx);
    ctx->has_worlds = false;
    const size_t sb = sizeof(float) * (size_t)W * n_stems * PMP_STEM_STRIDE;
    const size_t bb = sizeof(float) * (size_t)W * n_boxes * PMP_BOX_STRIDE;
    CK(cudaMalloc(&ctx->d_stems, sb));
    CK(cudaMalloc(&ctx->d_nstems, sizeof(int32_t) * W));
    CK(cudaMalloc(&ctx->d_boxes, bb));
    CK(cudaMalloc(&ctx->d_nboxes, sizeof(int32_t) * W));
    CK(cudaMemcpyAsync(ctx->d_stems, stems_dev, sb, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_boxes, boxes_dev, bb, cudaMemcpyDeviceToDevice, st));
    pmp_fill_i32_kernel<<<(W + 255) / 256, 256, 0, st>>>(ctx->d_nstems, W, n_stems);
    pmp_fill_i32_kernel<<<(W + 255) / 256, 256, 0, st>>>(ctx->d_nboxes, W, n_boxes);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->n_worlds = W; ctx->max_stems = n_stems; ctx->max_boxes = n_boxes; ctx->n_slots = n_slots;
    ctx->has_worlds = true;
    return PMP_OK;
}

extern "C" int pmp_set_params(pmp_ctx* ctx, const pmp_params* p) {
    if (!ctx || !p) return fail(ctx, PMP_E_INVALID, "null argument");
    if (!(p->resolution > 0.0)) return fail(ctx, PMP_E_INVALID, "resolution must be > 0");
    if (p->iterations < 1 || p->iterations > 64) return fail(ctx, PMP_E_INVALID, "iterations out of range");
    if (p->mode < 0 || p->mode > 2) return fail(ctx, PMP_E_INVALID, "unknown mode");
    ctx->prm.resolution = p->resolution;
    ctx->prm.deflection_limit = p->deflection_limit;
    ctx->prm.alpha = p->alpha;
    ctx->prm.max_step = p->max_step;
    ctx->prm.reg_eps = p->reg_eps;
    ctx->prm.iterations = p->iterations;
    ctx->prm.mode = p->mode;
    ctx->prm.dense = p->dense;
    if (!(p->gate_probability >= 0.f && p->gate_probability <= 1.f)) return fail(ctx, PMP_E_INVALID, "gate_probability outside [0, 1]");
    if (!(p->arm_lag_gain >= 0.f && p->arm_lag_gain < 1.f)) return fail(ctx, PMP_E_INVALID, "arm_lag_gain outside [0, 1)");
    if (p->arm_lag_gain > 0.f && p->arm_lag_substeps < 1) return fail(ctx, PMP_E_INVALID, "arm_lag_substeps must be >= 1");
    ctx->prm.lag_kappa = 0.f; ctx->prm.lag_log2rho = 0.f;
    if (p->arm_lag_gain > 0.f) {
        // closure per interpolation step c = 1 - (1 - gain)^substeps; lag of a uniform ramp = kappa (1 - rho^i) steps
        const double rho = pow(1.0 - (double)p->arm_lag_gain, (double)p->arm_lag_substeps), c = 1.0 - rho;
        ctx->prm.lag_kappa = (float)(rho / c);
        ctx->prm.lag_log2rho = (float)log2(rho);
    }
    ctx->prm.gate_seed = p->gate_seed;
    ctx->prm.gate_threshold = p->gate_probability >= 1.f ? 0xFFFFFFFFu : (uint32_t)((double)p->gate_probability * 4294967296.0);
    ctx->has_params = true;
    return PMP_OK;
}

extern "C" int pmp_num_worlds(pmp_ctx* ctx) { return ctx ? ctx->n_worlds : 0; }
extern "C" int pmp_max_stems(pmp_ctx* ctx) { return ctx ? ctx->max_stems : 0; }

static int ready(pmp_ctx* ctx) {
    if (!ctx) return PMP_E_INVALID;
    if (!ctx->has_robot || !ctx->has_worlds || !ctx->has_params)
        return fail(ctx, PMP_E_STATE, "robot, worlds and params must be set first");
    ctx->prm.n_worlds = ctx->n_worlds;
    ctx->prm.max_stems = ctx->max_stems;
    ctx->prm.max_boxes = ctx->max_boxes;
    return PMP_OK;
}

extern "C" int pmp_fk(pmp_ctx* ctx, const float* q_dev, int32_t B, float* pos, float* quat, float* com, void* stream) {
    if (!ctx || !ctx->has_robot) return fail(ctx, PMP_E_STATE, "robot not set");
    if (B < 0 || (B && !q_dev)) return fail(ctx, PMP_E_INVALID, "bad batch");
    if (B == 0) return PMP_OK;
    CK(cudaSetDevice(ctx->device));
    const int threads = 128;
    const size_t smem = sizeof(PmpRobotTab) + sizeof(PmpLinkTab);
    pmp_fk_kernel<<<(B + threads - 1) / threads, threads, smem, (cudaStream_t)stream>>>(ctx->d_robot, ctx->d_links, q_dev,
                                                                                      B, pos, quat, com);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return PMP_OK;
}

// ---- kernel selection (instantiations live in pmp_inst_a*.cu) ---------------------------------------
static void pads(pmp_ctx* ctx, int& apad, int& nspad) {
    const int A = ctx->h_robot.n_spheres;
    apad = A <= 12 ? 12 : (A <= 16 ? 16 : (A <= 18 ? 18 : (A <= 20 ? 20 : 32)));
    nspad = ctx->n_slots <= 1 ? 1 : (ctx->n_slots <= 2 ? 2 : 4);
}

static int validate_edges_impl(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps,
                               int32_t backward, int32_t* first_hit, int32_t* n_steps, float* max_defl, float* over,
                               float* ee_len, float* cost, uint32_t* rigid_mask, uint32_t* exceed_mask, void* stream,
                               int32_t edge_base, uint32_t* rigid_pre_scratch = nullptr);

extern "C" int pmp_validate_edges(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps,
                                  int32_t backward, int32_t* first_hit, int32_t* n_steps, float* max_defl, float* over,
                                  float* ee_len, float* cost, uint32_t* rigid_mask, uint32_t* exceed_mask, void* stream,
                                  int32_t edge_base) {
    return validate_edges_impl(ctx, q0, q1, E, max_steps, backward, first_hit, n_steps, max_defl, over, ee_len, cost,
                               rigid_mask, exceed_mask, stream, edge_base);
}

static int validate_edges_impl(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps,
                               int32_t backward, int32_t* first_hit, int32_t* n_steps, float* max_defl, float* over,
                               float* ee_len, float* cost, uint32_t* rigid_mask, uint32_t* exceed_mask, void* stream,
                               int32_t edge_base, uint32_t* rigid_pre_scratch) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (E < 0 || max_steps < 1) return fail(ctx, PMP_E_INVALID, "bad E / max_steps");
    if (E == 0) return PMP_OK;
    if (!q0 || !q1 || !first_hit || !n_steps) return fail(ctx, PMP_E_INVALID, "q0, q1, first_hit, n_steps are required");
    CK(cudaSetDevice(ctx->device));
    PmpEdgeArgs a;
    a.rigid_pre = nullptr;
    {
        // limits | self | obstacles of every step: one bit per step from the pre-pass (exact GJK test on the hulls, or
        // the arm's sphere model when the robot has none), consumed by the edge kernel below.  The scratch belongs to
        // the context (or to the caller's chunk); launches of one context are ordered by their stream.
        uint32_t* pre = rigid_pre_scratch;
        if (!pre) {
            rc = ensure_rigid_scratch(ctx, sizeof(uint32_t) * (size_t)E * ((max_steps + 31) / 32));
            if (rc) return rc;
            pre = (uint32_t*)ctx->d_rigid_scratch;
        }
        if (ctx->timing) CK(cudaEventRecord(ctx->ev[0], (cudaStream_t)stream));
        // without a rigid step mask in the outputs only the first rigid bit of every 32-step word is ever read
        rc = launch_rigid(ctx, q0, q1, E, max_steps, backward, pre, nullptr, 0, nullptr, (cudaStream_t)stream,
                          /*first_bit_only=*/rigid_mask == nullptr);
        if (rc) return rc;
        if (ctx->timing) { CK(cudaEventRecord(ctx->ev[1], (cudaStream_t)stream)); ctx->timed_rigid = true; }
        a.rigid_pre = pre;
    }
    a.robot = ctx->d_robot; a.stems = ctx->d_stems; a.n_stems = ctx->d_nstems; a.boxes = ctx->d_boxes;
    a.n_boxes = ctx->d_nboxes; a.prm = ctx->prm; a.q0 = q0; a.q1 = q1; a.E = E; a.max_steps = max_steps;
    a.backward = backward ? 1 : 0; a.first_hit = first_hit; a.n_steps = n_steps; a.max_defl = max_defl; a.over = over;
    a.ee_len = ee_len; a.cost = cost; a.rigid_mask = rigid_mask; a.exceed_mask = exceed_mask; a.edge_base = edge_base;
    if (!ctx->d_work_counter) CK(cudaMalloc(&ctx->d_work_counter, 64 * sizeof(int32_t)));
    a.work_counter = ctx->d_work_counter + (ctx->work_slot++ & 63u);   // launches in flight on other streams keep theirs
    CK(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), (cudaStream_t)stream));

    int apad, nspad;
    pads(ctx, apad, nspad);
    const int W = ctx->n_worlds;
    const size_t world_bytes = sizeof(float) * (size_t)W * ctx->max_stems * PMP_STEM_STRIDE;
    // geometry: one persistent CTA per SM with up to 16 warps (128 registers/thread fill the register file; measured
    // best on B200, profiles/r01_tuning.md); small batches use smaller CTAs so that they still spread over the SMs.
    // World tables go to shared memory when they fit, else they are read through L1.
    auto smem_for = [&](int blk, bool ws) {
        const int warps = blk / 32;
        return sizeof(PmpRobotTab) + (ws ? world_bytes + sizeof(int32_t) * ((W + 3) & ~3) : 0) +
               (size_t)warps * PMP_MAX_DOF * (sizeof(double) + 3 * sizeof(float)) +
               (size_t)warps * 3 * apad * 33 * sizeof(float);          // swept-bound staging tile per warp
    };
    // a call that asks for nothing but first_hit / n_steps gets the early-stopping instantiation; when such a batch is
    // too small to fill the machine with one warp per edge, the worlds of an edge are split over several warps
    const bool first_only = !max_defl && !over && !ee_len && !cost && !rigid_mask && !exceed_mask;
    int split = 1;
    if (first_only && W > 1) {
        const int target = ctx->sm_count * 16;
        if (E < target) split = (target + E - 1) / E;
        if (split > 16) split = 16;
        if (split > W) split = W;
    }
    a.world_split = split;
    const int items = E * split;
    if (split > 1) CK(cudaMemsetAsync(first_hit, 0x7f, sizeof(int32_t) * (size_t)E, (cudaStream_t)stream));
    int wpc = (items + ctx->sm_count - 1) / ctx->sm_count;      // warps (= work items in flight) per CTA
    if (wpc > 16) wpc = 16;
    if (wpc < 1) wpc = 1;
    // shared memory: robot table + per-warp staging always; world tables when they fit next to them.  A CTA of fewer
    // than 4 warps would spend longer staging the tables than using them.  If even the per-warp staging of `wpc`
    // warps does not fit (32-sphere arms: 12.7 KB of swept-bound tile per warp), use fewer warps.
    const size_t lim = (size_t)ctx->max_smem_optin;
    while (wpc > 1 && smem_for(32 * wpc, false) > lim) wpc >>= 1;
    int block = 32 * wpc, per_sm = 1;
    bool wsmem = wpc >= 4 && smem_for(block, true) <= lim;
    // tuning knobs (profiling only): PMP_EDGE_BLOCK = threads per CTA, PMP_EDGE_PER_SM = resident CTAs per SM
    if (const char* eb = getenv("PMP_EDGE_BLOCK")) { int v = atoi(eb); if (v >= 32 && v <= PMP_EDGE_MAXT && v % 32 == 0) block = v; }
    if (const char* ep = getenv("PMP_EDGE_PER_SM")) { int v = atoi(ep); if (v >= 1 && v <= 16) per_sm = v; }
    if (smem_for(block, wsmem) > lim) wsmem = false;
    if (smem_for(block, false) > lim) return fail(ctx, PMP_E_INVALID, "shared memory budget exceeded for this block size");
    const size_t smem = smem_for(block, wsmem);
    const int warps = block / 32;
    int grid = (items + warps - 1) / warps;
    const int cap = ctx->sm_count * per_sm;
    if (grid > cap) grid = cap;
    cudaError_t err;
    int regs = 0;
    cudaStream_t st_ = (cudaStream_t)stream;
    if (ctx->timing) CK(cudaEventRecord(ctx->ev[2], st_));
    if (first_only) {
        if (apad == 12) err = pmp_launch_edge_first_a12(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 16) err = pmp_launch_edge_first_a16(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 18) err = pmp_launch_edge_first_a18(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 20) err = pmp_launch_edge_first_a20(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else err = pmp_launch_edge_first_a32(nspad, a, wsmem, grid, block, smem, st_, &regs);
    } else {
        if (apad == 12) err = pmp_launch_edge_a12(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 16) err = pmp_launch_edge_a16(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 18) err = pmp_launch_edge_a18(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else if (apad == 20) err = pmp_launch_edge_a20(nspad, a, wsmem, grid, block, smem, st_, &regs);
        else err = pmp_launch_edge_a32(nspad, a, wsmem, grid, block, smem, st_, &regs);
    }
    ctx->info.regs_per_thread = regs;
    if (err != cudaSuccess) return fail(ctx, PMP_E_CUDA, std::string("edge kernel launch: ") + cudaGetErrorString(err));
    if (ctx->timing) { CK(cudaEventRecord(ctx->ev[3], st_)); ctx->timed_edge = true; }
    ctx->info.grid = grid; ctx->info.block = block; ctx->info.smem_bytes = (int)smem;
    ctx->info.worlds_in_smem = wsmem ? 1 : 0; ctx->info.n_sphere_pad = apad; ctx->info.n_slot_pad = nspad;
    ctx->info.kernel_launches++;
    return PMP_OK;
}

extern "C" int pmp_check_configs(pmp_ctx* ctx, const float* q, int32_t B, uint8_t* flags, float* defl, float* theta,
                                 float* ee, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (B < 0) return fail(ctx, PMP_E_INVALID, "bad B");
    if (B == 0) return PMP_OK;
    if (!q || !flags) return fail(ctx, PMP_E_INVALID, "q and flags are required");
    CK(cudaSetDevice(ctx->device));
    PmpConfigArgs a;
    a.robot = ctx->d_robot; a.stems = ctx->d_stems; a.n_stems = ctx->d_nstems; a.boxes = ctx->d_boxes;
    a.n_boxes = ctx->d_nboxes; a.prm = ctx->prm; a.q = q; a.B = B; a.flags = flags; a.defl = defl; a.theta = theta; a.ee = ee;
    a.rigid_pre = nullptr;
    {
        rc = ensure_rigid_scratch(ctx, (size_t)B);
        if (rc) return rc;
        rc = launch_rigid(ctx, nullptr, nullptr, 0, 32, 0, nullptr, q, B, (uint8_t*)ctx->d_rigid_scratch, (cudaStream_t)stream);
        if (rc) return rc;
        a.rigid_pre = (const uint8_t*)ctx->d_rigid_scratch;
    }
    int apad, nspad;
    pads(ctx, apad, nspad);
    // small batches (the scalar collision_fn adapter) spread over SMs: 32 threads per CTA
    const int block = B <= 32 * ctx->sm_count ? 32 : 128;
    const int gx = (B + block - 1) / block;
    // spread the worlds over blockIdx.y until the grid holds about two CTAs per SM
    int gy = (2 * ctx->sm_count + gx - 1) / gx;
    if (gy > ctx->n_worlds) gy = ctx->n_worlds;
    if (gy < 1) gy = 1;
    a.worlds_per_cta = (ctx->n_worlds + gy - 1) / gy;
    gy = (ctx->n_worlds + a.worlds_per_cta - 1) / a.worlds_per_cta;
    const dim3 grid(gx, gy, 1);
    const size_t smem = sizeof(PmpRobotTab);
    cudaError_t err;
    int regs = 0;
    if (apad == 12) err = pmp_launch_config_a12(nspad, a, grid, block, smem, (cudaStream_t)stream, &regs);
    else if (apad == 16) err = pmp_launch_config_a16(nspad, a, grid, block, smem, (cudaStream_t)stream, &regs);
    else if (apad == 18) err = pmp_launch_config_a18(nspad, a, grid, block, smem, (cudaStream_t)stream, &regs);
    else if (apad == 20) err = pmp_launch_config_a20(nspad, a, grid, block, smem, (cudaStream_t)stream, &regs);
    else err = pmp_launch_config_a32(nspad, a, grid, block, smem, (cudaStream_t)stream, &regs);
    ctx->info.regs_per_thread = regs;
    if (err != cudaSuccess) return fail(ctx, PMP_E_CUDA, std::string("config kernel launch: ") + cudaGetErrorString(err));
    ctx->info.grid = gx * gy; ctx->info.block = block; ctx->info.smem_bytes = (int)smem; ctx->info.worlds_in_smem = 0;
    ctx->info.n_sphere_pad = apad; ctx->info.n_slot_pad = nspad;
    ctx->info.kernel_launches++;
    return PMP_OK;
}

// ---- host-buffer entry points --------------------------------------------------------------------
static int ensure_scratch(pmp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return PMP_OK;
    cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4;
    CK(cudaMalloc(&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
    return PMP_OK;
}
static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int pmp_validate_edges_host(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps,
                                       int32_t backward, int32_t* first_hit, int32_t* n_steps, float* max_defl,
                                       float* over, float* ee_len, float* cost, uint32_t* rigid_mask,
                                       uint32_t* exceed_mask) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (E < 0 || max_steps < 1) return fail(ctx, PMP_E_INVALID, "bad E / max_steps");
    if (E == 0) return PMP_OK;
    if (!q0 || !q1 || !first_hit || !n_steps) return fail(ctx, PMP_E_INVALID, "q0, q1, first_hit, n_steps are required");
    CK(cudaSetDevice(ctx->device));
    const int D = ctx->h_robot.dof, W = ctx->n_worlds, C = (max_steps + 31) / 32;
    // Large batches go through in chunks on two streams, so that the copies of one chunk (inputs in, per-world
    // outputs back: up to 8 W + 16 bytes per edge) overlap the kernel of the other.
    const int n_chunks = E >= (1 << 16) ? 8 : 1;
    const int CE = (E + n_chunks - 1) / n_chunks;               // edges per chunk
    const size_t bq = up256(sizeof(float) * (size_t)CE * D), bi = up256(sizeof(int32_t) * (size_t)CE),
                 bw = up256(sizeof(float) * (size_t)CE * W), bm = up256(sizeof(uint32_t) * (size_t)CE * W * C);
    const size_t bp = up256(sizeof(uint32_t) * (size_t)CE * C);     // pre-pass bits
    const size_t per = 2 * bq + 2 * bi + (max_defl ? bw : 0) + (over ? bw : 0) + (ee_len ? bi : 0) + (cost ? bi : 0) +
                       (rigid_mask ? bm : 0) + (exceed_mask ? bm : 0) + bp;
    rc = ensure_scratch(ctx, (n_chunks > 1 ? 2 : 1) * per);
    if (rc) return rc;
    // An error inside the loop must not return while copies of earlier chunks are still writing into caller memory:
    // every failure path drains both streams first.
    auto bail = [&](int code) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream2);
        return code;
    };
#define CKB(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return bail(fail(ctx, PMP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)));           \
    } while (0)
    for (int c = 0; c < n_chunks; ++c) {
        const int e0 = c * CE;
        const int ne = E - e0 < CE ? E - e0 : CE;
        if (ne <= 0) break;
        cudaStream_t st = (c & 1) ? ctx->stream2 : ctx->stream;
        char* p = (char*)ctx->d_scratch + (size_t)(c & 1) * per;
        auto take = [&](size_t b) { char* r = p; p += b; return r; };
        float* dq0 = (float*)take(bq); float* dq1 = (float*)take(bq);
        int32_t* dfh = (int32_t*)take(bi); int32_t* dns = (int32_t*)take(bi);
        float* dmd = max_defl ? (float*)take(bw) : nullptr;
        float* dov = over ? (float*)take(bw) : nullptr;
        float* del = ee_len ? (float*)take(bi) : nullptr;
        float* dco = cost ? (float*)take(bi) : nullptr;
        uint32_t* drm = rigid_mask ? (uint32_t*)take(bm) : nullptr;
        uint32_t* dxm = exceed_mask ? (uint32_t*)take(bm) : nullptr;
        uint32_t* dpre = bp ? (uint32_t*)take(bp) : nullptr;
        const size_t o1 = (size_t)e0, oD = (size_t)e0 * D, oW = (size_t)e0 * W, oM = (size_t)e0 * W * C;
        CKB(cudaMemcpyAsync(dq0, q0 + oD, sizeof(float) * (size_t)ne * D, cudaMemcpyHostToDevice, st));
        CKB(cudaMemcpyAsync(dq1, q1 + oD, sizeof(float) * (size_t)ne * D, cudaMemcpyHostToDevice, st));
        rc = validate_edges_impl(ctx, dq0, dq1, ne, max_steps, backward, dfh, dns, dmd, dov, del, dco, drm, dxm, st, e0, dpre);
        if (rc) return bail(rc);
        CKB(cudaMemcpyAsync(first_hit + o1, dfh, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(n_steps + o1, dns, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost, st));
        if (max_defl) CKB(cudaMemcpyAsync(max_defl + oW, dmd, sizeof(float) * (size_t)ne * W, cudaMemcpyDeviceToHost, st));
        if (over) CKB(cudaMemcpyAsync(over + oW, dov, sizeof(float) * (size_t)ne * W, cudaMemcpyDeviceToHost, st));
        if (ee_len) CKB(cudaMemcpyAsync(ee_len + o1, del, sizeof(float) * (size_t)ne, cudaMemcpyDeviceToHost, st));
        if (cost) CKB(cudaMemcpyAsync(cost + o1, dco, sizeof(float) * (size_t)ne, cudaMemcpyDeviceToHost, st));
        if (rigid_mask)
            CKB(cudaMemcpyAsync(rigid_mask + oM, drm, sizeof(uint32_t) * (size_t)ne * W * C, cudaMemcpyDeviceToHost, st));
        if (exceed_mask)
            CKB(cudaMemcpyAsync(exceed_mask + oM, dxm, sizeof(uint32_t) * (size_t)ne * W * C, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_chunks > 1) CK(cudaStreamSynchronize(ctx->stream2));
    return PMP_OK;
#undef CKB
}

static int ensure_stage(pmp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return PMP_OK;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr; ctx->stage_bytes = 0;
    CK(cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocDefault));
    ctx->stage_bytes = bytes;
    return PMP_OK;
}

extern "C" int pmp_check_configs_host(pmp_ctx* ctx, const float* q, int32_t B, uint8_t* flags, float* defl,
                                      float* theta, float* ee) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (B < 0) return fail(ctx, PMP_E_INVALID, "bad B");
    if (B == 0) return PMP_OK;
    if (!q || !flags) return fail(ctx, PMP_E_INVALID, "q and flags are required");
    CK(cudaSetDevice(ctx->device));
    const int D = ctx->h_robot.dof, W = ctx->n_worlds, P = ctx->max_stems;
    const size_t nq = sizeof(float) * (size_t)B * D, nf = (size_t)B * W, nd = sizeof(float) * (size_t)B * W * P,
                 ne = sizeof(float) * (size_t)B * 3;
    const size_t bq = up256(nq), bf = up256(nf), bd = up256(nd), be = up256(ne);
    // outputs are laid out back to back so that a small call needs a single device-to-host copy
    const size_t out_bytes = bf + (defl ? bd : 0) + (theta ? 2 * bd : 0) + (ee ? be : 0);
    rc = ensure_scratch(ctx, bq + out_bytes);
    if (rc) return rc;
    char* p = (char*)ctx->d_scratch;
    auto take = [&](size_t b) { char* r = p; p += b; return r; };
    float* dq = (float*)take(bq);
    char* dout = p;
    uint8_t* dfl = (uint8_t*)take(bf);
    float* dd = defl ? (float*)take(bd) : nullptr;
    float* dt = theta ? (float*)take(2 * bd) : nullptr;
    float* de = ee ? (float*)take(be) : nullptr;
    cudaStream_t st = ctx->stream;
    const bool small = (bq + out_bytes) <= ((size_t)1 << 20);
    if (small) {
        rc = ensure_stage(ctx, bq + out_bytes);
        if (rc) return rc;
        char* hs = (char*)ctx->h_stage;
        memcpy(hs, q, nq);
        CK(cudaMemcpyAsync(dq, hs, nq, cudaMemcpyHostToDevice, st));
        rc = pmp_check_configs(ctx, dq, B, dfl, dd, dt, de, st);
        if (rc) return rc;
        char* ho = hs + bq;
        CK(cudaMemcpyAsync(ho, dout, out_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(flags, ho + ((char*)dfl - dout), nf);
        if (defl) memcpy(defl, ho + ((char*)dd - dout), nd);
        if (theta) memcpy(theta, ho + ((char*)dt - dout), 2 * nd);
        if (ee) memcpy(ee, ho + ((char*)de - dout), ne);
        return PMP_OK;
    }
    CK(cudaMemcpyAsync(dq, q, nq, cudaMemcpyHostToDevice, st));
    rc = pmp_check_configs(ctx, dq, B, dfl, dd, dt, de, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(flags, dfl, nf, cudaMemcpyDeviceToHost, st));
    if (defl) CK(cudaMemcpyAsync(defl, dd, nd, cudaMemcpyDeviceToHost, st));
    if (theta) CK(cudaMemcpyAsync(theta, dt, 2 * nd, cudaMemcpyDeviceToHost, st));
    if (ee) CK(cudaMemcpyAsync(ee, de, ne, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PMP_OK;
}

extern "C" int pmp_get_launch_info(pmp_ctx* ctx, pmp_launch_info* out) {
    if (!ctx || !out) return PMP_E_INVALID;
    *out = ctx->info;
    return PMP_OK;
}

#ifdef PMP_RIGID_STATS
// experiment builds only (not declared in the public header): read and clear the narrow-phase counters
extern "C" int pmp_debug_rigid_stats(unsigned long long* out128) {
    unsigned long long z[128] = {0};
    if (cudaMemcpyFromSymbol(out128, pmp_rigid_stats, sizeof(z)) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(pmp_rigid_stats, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int pmp_set_kernel_timing(pmp_ctx* ctx, int32_t on) {
    if (!ctx) return PMP_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (on)
        for (int i = 0; i < 4; ++i)
            if (!ctx->ev[i]) CK(cudaEventCreate(&ctx->ev[i]));
    ctx->timing = on != 0;
    ctx->timed_rigid = ctx->timed_edge = false;
    return PMP_OK;
}

extern "C" int pmp_get_kernel_times(pmp_ctx* ctx, float* rigid_ms, float* edge_ms) {
    if (!ctx || !rigid_ms || !edge_ms) return fail(ctx, PMP_E_INVALID, "null argument");
    *rigid_ms = 0.f; *edge_ms = 0.f;
    if (!ctx->timing || !ctx->timed_edge) return fail(ctx, PMP_E_STATE, "no timed pmp_validate_edges call yet");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev[3]));
    CK(cudaEventElapsedTime(edge_ms, ctx->ev[2], ctx->ev[3]));
    if (ctx->timed_rigid) CK(cudaEventElapsedTime(rigid_ms, ctx->ev[0], ctx->ev[1]));
    return PMP_OK;
}

extern "C" int pmp_measure_fp32_peak(pmp_ctx* ctx, int32_t repeats, double* tflops, double* ms) {
    if (!ctx || !tflops || !ms) return fail(ctx, PMP_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    const int block = 256, grid = ctx->sm_count * 8, iters = 4096;
    float* d_out = nullptr;
    CK(cudaMalloc(&d_out, sizeof(float) * (size_t)grid * block));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < 3; ++i) pmp_fma_probe_kernel<<<grid, block, 0, st>>>(d_out, iters, 0.999f, 0.001f);
    double best = 1e30;
    if (repeats < 1) repeats = 1;
    for (int r = 0; r < repeats; ++r) {
        CK(cudaEventRecord(e0, st));
        pmp_fma_probe_kernel<<<grid, block, 0, st>>>(d_out, iters, 0.999f, 0.001f);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, e0, e1));
        if (t < best) best = t;
        ctx->info.kernel_launches++;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * block;
    *ms = best;
    *tflops = flops / (best * 1e-3) / 1e12;
    return PMP_OK;
}

// ---- batched tree extension -------------------------------------------------------
static int ensure_tree_scratch(pmp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->tree_scratch_bytes) return PMP_OK;
    cudaFree(ctx->d_tree_scratch);
    ctx->d_tree_scratch = nullptr; ctx->tree_scratch_bytes = 0;
    size_t want = bytes + bytes / 2;
    CK(cudaMalloc(&ctx->d_tree_scratch, want));
    ctx->tree_scratch_bytes = want;
    return PMP_OK;
}

static int nearest_tile(pmp_ctx* ctx) { return ctx->h_robot.dof <= 8 ? 8 : 4; }

static int nearest_splits(pmp_ctx* ctx, int T, int B) {
    const int tgt = nearest_tile(ctx);
    const int tiles = (B + tgt - 1) / tgt;
    int splits = (16 * ctx->sm_count + tiles - 1) / tiles;           // about eight waves of two CTAs per SM: short tail
    const int max_splits = (T + 4 * PMP_NN_BLOCK - 1) / (4 * PMP_NN_BLOCK);   // at least 4 nodes per thread
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return splits;
}

// launches the two nearest-node kernels; part_* must hold B * splits entries
static int launch_nearest(pmp_ctx* ctx, const float* nodes, int T, const int32_t* size_dev, const float* targets, int B,
                          int splits, uint32_t* min32, double* part_r, int32_t* part_i, int32_t* index, double* dist,
                          float* q_from, cudaStream_t st) {
    PmpNearestArgs a;
    a.min32 = min32;
    a.nodes = nodes; a.size_dev = size_dev; a.T = T; a.targets = targets; a.B = B; a.D = ctx->h_robot.dof;
    a.circular_mask = ctx->h_robot.circular_mask; a.splits = splits; a.part_r = part_r; a.part_i = part_i;
    a.index = index; a.dist = dist; a.q_from = q_from;
    const int tgt = nearest_tile(ctx);
    const dim3 grid((B + tgt - 1) / tgt, splits, 1);
    CK(cudaMemsetAsync(min32, 0x7f, sizeof(uint32_t) * (size_t)B, st));      // 0x7f7f7f7f = 3.39e38f
    const bool circ = a.circular_mask != 0;
#define PMP_NN_LAUNCH(DP, TG, CI)                                                \
    do {                                                                         \
        pmp_nearest_min_kernel<DP, TG, CI><<<grid, PMP_NN_BLOCK, 0, st>>>(a);    \
        pmp_nearest_kernel<DP, TG, CI><<<grid, PMP_NN_BLOCK, 0, st>>>(a);        \
    } while (0)
    if (a.D == 7) { if (circ) PMP_NN_LAUNCH(7, 8, true); else PMP_NN_LAUNCH(7, 8, false); }        // no padding term
    else if (a.D == 6) { if (circ) PMP_NN_LAUNCH(6, 8, true); else PMP_NN_LAUNCH(6, 8, false); }
    else if (a.D <= 8) { if (circ) PMP_NN_LAUNCH(8, 8, true); else PMP_NN_LAUNCH(8, 8, false); }
    else { if (circ) PMP_NN_LAUNCH(16, 4, true); else PMP_NN_LAUNCH(16, 4, false); }
#undef PMP_NN_LAUNCH
    pmp_nearest_finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches += 3;
    return PMP_OK;
}

extern "C" int pmp_nearest(pmp_ctx* ctx, const float* nodes_dev, int32_t T, const int32_t* size_dev,
                           const float* targets_dev, int32_t B, int32_t* index_dev, double* dist_dev,
                           float* q_from_dev, void* stream) {
    if (!ctx) return PMP_E_INVALID;
    if (!ctx->has_robot) return fail(ctx, PMP_E_STATE, "robot must be set first");
    if (T < 1 || B < 0) return fail(ctx, PMP_E_INVALID, "bad T / B");
    if (B == 0) return PMP_OK;
    if (!nodes_dev || !targets_dev || !index_dev) return fail(ctx, PMP_E_INVALID, "nodes, targets and index are required");
    CK(cudaSetDevice(ctx->device));
    const int splits = nearest_splits(ctx, T, B);
    const size_t br = up256(sizeof(double) * (size_t)B * splits), bi = up256(sizeof(int32_t) * (size_t)B * splits),
                 bm = up256(sizeof(uint32_t) * (size_t)B);
    int rc = ensure_tree_scratch(ctx, br + bi + bm);
    if (rc) return rc;
    char* p = (char*)ctx->d_tree_scratch;
    return launch_nearest(ctx, nodes_dev, T, size_dev, targets_dev, B, splits, (uint32_t*)(p + br + bi), (double*)p,
                          (int32_t*)(p + br), index_dev, dist_dev, q_from_dev, (cudaStream_t)stream);
}

extern "C" int pmp_edge_points(pmp_ctx* ctx, const float* q0_dev, const float* q1_dev, int32_t E,
                               const int32_t* step_dev, int32_t backward, float* out_dev, void* stream) {
    if (!ctx) return PMP_E_INVALID;
    if (!ctx->has_robot || !ctx->has_params) return fail(ctx, PMP_E_STATE, "robot and params must be set first");
    if (E < 0) return fail(ctx, PMP_E_INVALID, "bad E");
    if (E == 0) return PMP_OK;
    if (!q0_dev || !q1_dev || !step_dev || !out_dev) return fail(ctx, PMP_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    pmp_edge_points_kernel<<<(E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        q0_dev, q1_dev, E, ctx->h_robot.dof, ctx->h_robot.circular_mask, ctx->prm.resolution, step_dev, backward ? 1 : 0,
        out_dev);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return PMP_OK;
}

extern "C" int pmp_tree_extend(pmp_ctx* ctx, const pmp_tree* tree, int32_t size_bound, const float* targets_dev,
                               int32_t B, int32_t max_steps, int32_t* new_index_dev, int32_t* reached_dev,
                               float* new_cfg_dev, int32_t* stats_dev, void* stream) {
    int rc = ready(ctx);
    if (rc) return rc;
    if (!tree || !tree->nodes || !tree->parent || !tree->via || !tree->via_steps || !tree->size)
        return fail(ctx, PMP_E_INVALID, "incomplete pmp_tree");
    if (B < 0 || max_steps < 1 || size_bound < 1 || size_bound > tree->capacity)
        return fail(ctx, PMP_E_INVALID, "bad B / max_steps / size_bound");
    if (B == 0) return PMP_OK;
    if (!targets_dev || !new_index_dev || !reached_dev || !stats_dev)
        return fail(ctx, PMP_E_INVALID, "targets, new_index, reached and stats are required");
    CK(cudaSetDevice(ctx->device));
    const int D = ctx->h_robot.dof;
    const int splits = nearest_splits(ctx, size_bound, B);
    const size_t br = up256(sizeof(double) * (size_t)B * splits), bp = up256(sizeof(int32_t) * (size_t)B * splits),
                 bi = up256(sizeof(int32_t) * (size_t)B), bq = up256(sizeof(float) * (size_t)B * D);
    rc = ensure_tree_scratch(ctx, br + bp + 4 * bi + bq);
    if (rc) return rc;
    char* p = (char*)ctx->d_tree_scratch;
    auto take = [&](size_t b) { char* r = p; p += b; return r; };
    double* part_r = (double*)take(br);
    int32_t* part_i = (int32_t*)take(bp);
    int32_t* from_index = (int32_t*)take(bi);
    int32_t* first_hit = (int32_t*)take(bi);
    int32_t* n_steps = (int32_t*)take(bi);
    uint32_t* min32 = (uint32_t*)take(bi);
    float* q_from = (float*)take(bq);
    cudaStream_t st = (cudaStream_t)stream;
    // 1. nearest node of every target (+ gather), 2. the edge kernel on (nearest node -> target),
    // 3. append the last safe configuration of every extension that made progress
    rc = launch_nearest(ctx, tree->nodes, size_bound, tree->size, targets_dev, B, splits, min32, part_r, part_i,
                        from_index, nullptr, q_from, st);
    if (rc) return rc;
    rc = pmp_validate_edges(ctx, q_from, targets_dev, B, max_steps, 0, first_hit, n_steps, nullptr, nullptr, nullptr,
                            nullptr, nullptr, nullptr, stream, 0);
    if (rc) return rc;
    PmpTreeAppendArgs a;
    a.nodes = tree->nodes; a.parent = tree->parent; a.via = tree->via; a.via_steps = tree->via_steps; a.size = tree->size;
    a.capacity = tree->capacity; a.q_from = q_from; a.targets = targets_dev; a.from_index = from_index;
    a.first_hit = first_hit; a.n_steps = n_steps; a.B = B; a.D = D; a.max_steps = max_steps;
    a.circular_mask = ctx->h_robot.circular_mask; a.resolution = ctx->prm.resolution; a.new_index = new_index_dev;
    a.reached = reached_dev; a.new_cfg = new_cfg_dev; a.stats = stats_dev;
    pmp_tree_append_kernel<<<1, B >= 1024 ? 1024 : ((B + 31) / 32) * 32, 0, st>>>(a);
    CK(cudaGetLastError());
    ctx->info.kernel_launches++;
    return PMP_OK;
}
