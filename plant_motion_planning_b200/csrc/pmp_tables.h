// Device-side table layouts shared by the kernels and the C-ABI glue.
// Mirrors plant_motion_planning_b200/model/tables.py (host packer) -- keep both in sync.
#pragma once
#include <stdint.h>
#include "../../include/pmp_edgecheck.h"

// ---- stem record: PMP_STEM_STRIDE floats (written by model/tables.py build_stem_records) ---------
//  [0..8]   Rs0  world rotation of the stem LINK frame in its rest pose (springs and gravity in balance), row-major
//  [9..11]  b0   = Rs0^T tv0 + zc e_z : p = Rs0^T c - b0 is a world point c relative to the box centre, link coords
//  [12] hw  [13] hh   core half width / half height of the stem's box (Bullet's margin carved out)
//  [14] zc  box centre height on the stem's z axis       [15] margin (added to the arm spheres' radii)
//  [16..18] observer vector  (MOD3: c, see tables.py; ANGLE: unit reference direction)
//  [19..21] observer reference point (ANGLE only)
//  [22] COM offset along local z     [23] raw0: observer value of the rest pose (chain undeflected)
//  [24] int parent frame slot (-1 = root)   [25] int own frame slot (-1 = leaf)   [26] int flags   [27] joint limit
//  [28..29] th0  rest angles (tx, ty)       [30] radius of the capsule that bounds the box (swept-bound cull)
//  [32..40] Ro   joint-frame rotation in the parent stem's link frame (world frame for roots), [41..43] to  origin
//  [44..46] tv0  world position of the joint in the rest pose
// The first 16 floats are all the rest-pose tests need (four 128-bit loads).
#define PMP_S_RS0 0
#define PMP_S_B0 9
#define PMP_S_HW 12
#define PMP_S_HH 13
#define PMP_S_ZC 14
#define PMP_S_MARGIN 15
#define PMP_S_OBSV 16
#define PMP_S_OBSP 19
#define PMP_S_COMZ 22
#define PMP_S_RAW0 23
#define PMP_S_PSLOT 24
#define PMP_S_OSLOT 25
#define PMP_S_FLAGS 26
#define PMP_S_LIM 27
#define PMP_S_TH0 28
#define PMP_S_BRAD 30
#define PMP_S_RO 32
#define PMP_S_TO 41
#define PMP_S_TV0 44

#define PMP_SF_PLANT_START 1  // observer chain restarts here           (representation.py:270)
#define PMP_SF_IS_MAIN 2      // updates the running main-stem angle    (representation.py:289)
#define PMP_SF_OBS_ANGLE 4    // ANGLE observer (else MOD3)
#define PMP_SF_OBS_ROOT 8     // MOD3 without parent link               (representation.py:207-209)
#define PMP_SF_SNAP 16        // zero |x|,|y| < 1e-5                    (representation.py:41-44)
#define PMP_SF_HAS_PARENT 32

// ---- robot + static obstacle table, resident in global memory and staged to shared memory ------
struct PmpRobotTab {
    float joint_C[PMP_MAX_DOF * 9];
    float joint_A[PMP_MAX_DOF * 9];
    float joint_B[PMP_MAX_DOF * 9];
    float joint_t[PMP_MAX_DOF * 3];
    float lower[PMP_MAX_DOF];
    float upper[PMP_MAX_DOF];
    float base_R[9];
    float base_t[3];
    float sph_center[PMP_MAX_SPHERES * 3];
    float sph_radius[PMP_MAX_SPHERES];
    float sph_radius_outer[PMP_MAX_SPHERES];   // covering radii (broad phase of the hull test); == sph_radius without hulls
    float sph_radius_inner[PMP_MAX_SPHERES];   // inscribed radii (ball inside hull (+) margin; 0 = unknown): certain hits
    uint32_t self_mask[PMP_MAX_SPHERES];
    int32_t sph_joint[PMP_MAX_SPHERES];
    float fixed_R[PMP_MAX_FIXED * 9];
    float fixed_t[PMP_MAX_FIXED * 3];
    float fixed_half[PMP_MAX_FIXED * 3];
    uint32_t fixed_mask[PMP_MAX_FIXED];
    int32_t fixed_kind[PMP_MAX_FIXED];
    float ee_local[3];
    int32_t ee_joint;
    int32_t dof;
    int32_t n_spheres;
    int32_t n_fixed;
    uint32_t circular_mask;   // bit j = joint j continuous
    int32_t has_hulls;        // 1: self / fixed collisions are decided by the hull pre-pass (PmpHullTab), not here
    int32_t pad[3];           // sizeof % 16 == 0
};
static_assert(sizeof(PmpRobotTab) % 16 == 0, "PmpRobotTab must be a whole number of uint4");

// Convex hulls of the arm's collision meshes (pmp_set_robot_hulls): the exact narrow phase of the rigid test.
// Vertices live in a separate float4 array (x, y, z, 0) in global memory.
struct PmpHullTab {
    int32_t n_hulls;
    int32_t n_pairs;
    int32_t n_fixed_exact;
    float fixed_margin;
    int32_t hull_joint[PMP_MAX_HULLS];
    int32_t hull_off[PMP_MAX_HULLS + 1];
    float hull_margin[PMP_MAX_HULLS];
    int32_t sph_lo[PMP_MAX_HULLS];      // arm spheres [sph_lo, sph_hi) approximate hull h (spheres are sorted by joint)
    int32_t sph_hi[PMP_MAX_HULLS];
    int32_t pairs[PMP_MAX_HULL_PAIRS * 2];
    // Clearance tables of hull pairs whose relative pose depends on exactly two joints (a, b): node (ia, ib) holds
    // bit 0 = "free for every configuration within half a grid step" and bit 1 = "colliding for every ...", decided
    // at set-up by the device GJK with the margins moved by the Lipschitz bound of the clearance over half a step.
    // tab_off < 0: no table (the covering spheres cull that pair).
    int32_t tab_off[PMP_MAX_HULL_PAIRS];
    int32_t tab_ja[PMP_MAX_HULL_PAIRS], tab_jb[PMP_MAX_HULL_PAIRS];
    int32_t tab_na[PMP_MAX_HULL_PAIRS], tab_nb[PMP_MAX_HULL_PAIRS];
    float tab_lo_a[PMP_MAX_HULL_PAIRS], tab_inv_a[PMP_MAX_HULL_PAIRS];
    float tab_lo_b[PMP_MAX_HULL_PAIRS], tab_inv_b[PMP_MAX_HULL_PAIRS];
    int32_t pad[3];
};
static_assert(sizeof(PmpHullTab) % 16 == 0, "PmpHullTab must be a whole number of uint4");

struct PmpLinkTab {
    int32_t n_links;
    int32_t link_joint[PMP_MAX_LINKS];
    float link_R[PMP_MAX_LINKS * 9];
    float link_t[PMP_MAX_LINKS * 3];
    float link_com[PMP_MAX_LINKS * 3];
    int32_t pad[3];
};
static_assert(sizeof(PmpLinkTab) % 16 == 0, "PmpLinkTab must be a whole number of uint4");

struct PmpKernelParams {
    double resolution;
    float deflection_limit;
    float alpha;
    float max_step;
    float reg_eps;
    int32_t iterations;
    int32_t mode;
    int32_t dense;
    uint32_t gate_seed;
    uint32_t gate_threshold;   // gate_probability * 2^32 (0 = gate off)
    int32_t n_worlds;
    int32_t max_stems;
    int32_t max_boxes;
    // arm lag (env.py:154,170): step i of an edge sees t_eff = t_i - lag_kappa (1 - lag_rho^i) / n; kappa = 0: off
    float lag_kappa;          // (1 - c) / c,  c = 1 - (1 - gain)^substeps
    float lag_log2rho;        // log2(1 - c)
};
