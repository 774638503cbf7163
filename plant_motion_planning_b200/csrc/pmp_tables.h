// Device-side table layouts shared by the kernels and the C-ABI glue.
// Mirrors plant_motion_planning_b200/model/tables.py (host packer) -- keep both in sync.
#pragma once
#include <stdint.h>
#include "../../include/pmp_edgecheck.h"

// ---- stem record: PMP_STEM_STRIDE floats -------------------------------------------------------
//  [0..8]   Ro   joint-frame rotation in the parent stem's link frame (world frame for roots), row-major
//  [9..11]  to   joint-frame origin
//  [12] z0  [13] z1   capsule segment span on the stem's local +z      [14] capsule radius
//  [15] joint limit (rad)
//  [16..18] observer vector  (MOD3: c, see tables.py; ANGLE: unit reference direction)
//  [19..21] observer reference point (ANGLE only)
//  [22] COM offset along local z     [23] unused
//  [24] int parent frame slot (-1 = root)   [25] int own frame slot (-1 = leaf)   [26] int flags
//  [32..40] Rw0  [41..43] tw0   world pose of the joint frame with every stem at rest (used while the
//           whole ancestor chain is undeflected, so untouched plants never compose frames)
#define PMP_S_RO 0
#define PMP_S_TO 9
#define PMP_S_Z0 12
#define PMP_S_Z1 13
#define PMP_S_RAD 14
#define PMP_S_LIM 15
#define PMP_S_OBSV 16
#define PMP_S_OBSP 19
#define PMP_S_COMZ 22
#define PMP_S_PSLOT 24
#define PMP_S_OSLOT 25
#define PMP_S_FLAGS 26
#define PMP_S_RW0 32
#define PMP_S_TW0 41

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
    int32_t pad[4];           // sizeof % 16 == 0
};
static_assert(sizeof(PmpRobotTab) % 16 == 0, "PmpRobotTab must be a whole number of uint4");

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
};
