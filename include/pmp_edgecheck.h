/*
 * pmp_edgecheck.h -- C ABI of the B200 edge-validation library (libpmp_edgecheck.so).
 *
 * The reference (UM-ARM-Lab/plant_motion_planning) is pure Python on top of PyBullet and has no FFI
 * of its own; the entry points below are what a ctypes binding on the reference side would bind to
 * replace its per-step PyBullet calls (see INTEGRATION.md for the stub).  Each entry point names the
 * reference interface it stands in for (paths relative to plant_motion_planning/ in the reference).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * PMP_E_* code on failure, with a message available from pmp_last_error().  "_dev" pointers are device
 * memory on the context's GPU and are owned by the caller; nothing is allocated inside the *_dev calls.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream); device calls are asynchronous on it.
 * A context is not thread-safe (the reference is single-threaded: one global Bullet client).
 */
#ifndef PMP_EDGECHECK_H
#define PMP_EDGECHECK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMP_MAX_DOF 16
#define PMP_MAX_SPHERES 32
#define PMP_MAX_LINKS 64
#define PMP_MAX_FIXED 16
#define PMP_MAX_STEMS 32
#define PMP_MAX_WORLD_BOXES 32
#define PMP_MAX_SLOTS 4   /* parent-frame register slots compiled into the kernels */
#define PMP_STEM_STRIDE 48 /* floats per stem record, layout in csrc/pmp_tables.h */
#define PMP_BOX_STRIDE 16  /* floats per plant-base box record: R[9] t[3] half[3] pad */

#define PMP_OK 0
#define PMP_E_INVALID (-1)  /* bad argument / table out of capacity */
#define PMP_E_CUDA (-2)     /* CUDA runtime error */
#define PMP_E_STATE (-3)    /* robot / worlds / params not set */
#define PMP_E_NOMEM (-4)

#define PMP_MODE_OUR_METHOD 0 /* rigid obstacles + deflection limit   (scripts/multi_world_our_method.py) */
#define PMP_MODE_AVOID_ALL 1  /* undeflected plants are obstacles too (env.py:75-77, avoid_all=True)      */
#define PMP_MODE_IGNORE_ALL 2 /* rigid obstacles only                 (collision_fn_benchmark, env.py:112) */

#define PMP_FLAG_RIGID 1    /* limits | self | fixed obstacle hit  (pybullet_tools/utils.py:3979-4016) */
#define PMP_FLAG_EXCEEDED 2 /* some stem deflection > limit        (pybullet_tools/utils.py:3841-3843, gate NOT applied) */

#define PMP_PRIM_BOX 0
#define PMP_PRIM_CYLINDER 1
#define PMP_PRIM_SPHERE 2

typedef struct pmp_ctx pmp_ctx;

/* Robot + static obstacles, host SoA arrays (copied by pmp_set_robot).
 * Replaces the setup-time closure state of get_collision_fn_with_controls_v3
 * (pybullet_tools/utils.py:3965-3977: check_link_pairs, moving_links, limits) and the loaded URDF. */
typedef struct pmp_robot_desc {
    int32_t dof;              /* planned joints D (<= PMP_MAX_DOF), one serial chain */
    int32_t n_spheres;        /* arm collision spheres A (<= PMP_MAX_SPHERES), sorted by joint */
    int32_t n_links;          /* links reported by pmp_fk (<= PMP_MAX_LINKS), PyBullet order */
    int32_t n_fixed;          /* static primitives: obstacles then the robot's welded geometry */
    const float* joint_C;     /* [D*9] row-major; R0*Rot(axis,q) = C + cos(q) A + sin(q) B */
    const float* joint_A;     /* [D*9] */
    const float* joint_B;     /* [D*9] */
    const float* joint_t;     /* [D*3] joint origin in the previous joint frame */
    const float* lower;       /* [D] */
    const float* upper;       /* [D] */
    const uint8_t* circular;  /* [D] 1 = continuous joint (difference wraps, no limit test) */
    const float* base_R;      /* [9] world pose of the chain root */
    const float* base_t;      /* [3] */
    const int32_t* sph_joint; /* [A] joint the sphere rides on */
    const float* sph_center;  /* [A*3] in that joint's frame */
    const float* sph_radius;  /* [A], each in (0, 0.5] m */
    const uint32_t* self_mask;/* [A] bit b (b > a) set => test sphere pair (a, b) */
    const int32_t* link_joint;/* [L] joint the link rides on, -1 = welded to the base */
    const float* link_R;      /* [L*9] link frame in the joint frame (or root frame) */
    const float* link_t;      /* [L*3] */
    const float* link_com;    /* [L*3] inertial origin in the link frame */
    int32_t ee_joint;         /* end-effector point for the path-length cost */
    float ee_local[3];
    const int32_t* fixed_kind;/* [F] PMP_PRIM_* */
    const float* fixed_R;     /* [F*9] world */
    const float* fixed_t;     /* [F*3] world */
    const float* fixed_half;  /* [F*3] box half extents | (r, half_len, -) | (r, -, -) */
    const uint32_t* fixed_mask;/* [F] bit a set => arm sphere a is tested against it */
} pmp_robot_desc;

/* Convex hulls of the arm's collision meshes (copied by pmp_set_robot_hulls; optional).  With hulls set, self and
 * fixed-obstacle collisions are decided EXACTLY on them -- Bullet's test behind p.getClosestPoints(distance=0),
 * pybullet_tools/utils.py:3390-3441: GJK distance of the margin-free convex shapes minus both collision margins <= 0 --
 * and the arm spheres (with their covering radii) only cull.  Without hulls the sphere model decides.
 * The plant contact solve always works on the spheres. */
#define PMP_MAX_HULLS 16
#define PMP_MAX_HULL_PAIRS 64
#define PMP_MAX_HULL_VERTS 16384
typedef struct pmp_hull_desc {
    int32_t n_hulls;               /* <= PMP_MAX_HULLS; 0 removes the hulls */
    const int32_t* hull_joint;     /* [n_hulls] joint the hull rides on */
    const int32_t* hull_off;       /* [n_hulls+1] vertex ranges in hull_verts */
    const float* hull_verts;       /* [3*hull_off[n_hulls]] vertices in the joint frame (<= PMP_MAX_HULL_VERTS) */
    const float* hull_margin;      /* [n_hulls] collision margin outside the hull (btConvexHullShape: 0.001) */
    int32_t n_pairs;               /* self pairs (get_self_link_pairs, pybullet_tools/utils.py:3740-3751) */
    const int32_t* pairs;          /* [2*n_pairs] hull indices */
    const int32_t* sph_hull;       /* [n_spheres] hull that arm sphere a approximates (-1 = none) */
    const float* sph_radius_outer; /* [n_spheres] radii with which the spheres of a hull COVER hull (+) margin */
    float fixed_margin;            /* collision margin carved out of the static boxes (btBoxShape: 0.001) */
    int32_t n_fixed_exact;         /* the first n_fixed_exact static primitives (boxes / spheres) are obstacles
                                      tested against every hull (product(moving links, fixed), utils.py:4008-4014) */
    const float* sph_radius_inner; /* [n_spheres] or NULL: radius of the ball around sphere a's centre that lies INSIDE
                                      hull (+) margin (depth of the centre in the hull + margin; 0 = unknown).  Two
                                      such balls overlapping, or one reaching an obstacle, decide "colliding" in the
                                      broad phase without GJK -- exact, the balls are subsets of the shapes */
} pmp_hull_desc;

/* Sampled plant worlds (copied by pmp_set_worlds).  Replaces the per-world Bullet bodies created by
 * SinglePlantEnv.__init__ (env.py:63-73) and their CharacterizePlant observers. */
typedef struct pmp_world_desc {
    int32_t n_worlds;
    int32_t max_stems;       /* row stride of `stems` in records (<= PMP_MAX_STEMS) */
    int32_t max_boxes;       /* row stride of `boxes` in records (<= PMP_MAX_WORLD_BOXES) */
    int32_t n_slots;         /* parent-frame register slots the stem records use (1..PMP_MAX_SLOTS) */
    const int32_t* n_stems;  /* [W] */
    const float* stems;      /* [W*max_stems*PMP_STEM_STRIDE] */
    const int32_t* n_boxes;  /* [W] plant-base boxes (tested in PMP_MODE_AVOID_ALL only) */
    const float* boxes;      /* [W*max_boxes*PMP_BOX_STRIDE] */
} pmp_world_desc;

/* ---- plant-world sampler on the device ----------------------------------------------------
 * Replaces, for large world counts, the per-world create_plant_params call of SinglePlantEnv.__init__
 * (env.py:63-73; rand_gen_plants_programmatically_main.py:71-277): one launch draws W worlds of one topology and
 * writes their stem records and base boxes into caller-owned device buffers, pmp_set_worlds_dev adopts them -- the
 * worlds never exist on the host.  Distributions, branch-placement rule and joint layout are the reference's; the ORDER
 * of draws is not (a world consumes a fixed block of uniforms, so it does not depend on which other worlds are drawn):
 * the bit-for-bit seed replay of the scripts stays with the host sampler (model/plants.sample_world). */
#define PMP_SAMPLER_MAX_STEMS 8
#define PMP_SAMPLER_TRIES 16      /* branch-placement attempts (rand_gen...:196-213 loops until no overlap) */
#define PMP_SAMPLER_HEAD 5        /* uniforms per world: plant x, y, base half extents */
#define PMP_SAMPLER_PER_STEM (8 + 3 * PMP_SAMPLER_TRIES)   /* + per stem: height, angles, placement attempts */
typedef struct pmp_sampler_desc {
    int32_t n_stems;                           /* <= PMP_SAMPLER_MAX_STEMS, parents first */
    int32_t parent[PMP_SAMPLER_MAX_STEMS];     /* -1 for the first stem */
    int32_t kind[PMP_SAMPLER_MAX_STEMS];       /* 0 first, 1 stacked vertical stem, 2 branch, 3 extension */
    int32_t ref[PMP_SAMPLER_MAX_STEMS];        /* main stem of a branch / stacked stem, previous stem of an extension */
    int32_t flags[PMP_SAMPLER_MAX_STEMS];      /* stem-record flags (PMP_SF_*), a function of the topology */
    int32_t pslot[PMP_SAMPLER_MAX_STEMS];      /* parent / own frame slot (tables.assign_slots) */
    int32_t oslot[PMP_SAMPLER_MAX_STEMS];
    int32_t n_slots;
    int32_t gravity_sag;                       /* 1: records hold the spring / gravity rest pose (model/statics.py) */
    double xy_lo[2], xy_hi[2];                 /* plant_pos_xy_limits (env.py:63-64) */
    double spacing;                            /* stem_base_spacing (rand_gen...:19-20) */
    double scaling;                            /* 0.25 (rand_gen...:18) */
    double half_width, com_z, limit;           /* 0.025, 0.25, 1.5 (rand_gen...:74-75, 265, 307) */
    double stem_margin;                        /* Bullet's margin carved out of the stem boxes, in [0, 0.4] m */
    double kp, mass, g;                        /* 500, 10, 9.8 (env.py:128; rand_gen...:255; utils.py:1218-1221) */
} pmp_sampler_desc;
/* stems_dev [W * n_stems * PMP_STEM_STRIDE], boxes_dev [W * PMP_BOX_STRIDE] (one base box per world).
 * uniforms_dev: NULL = counter-based generator keyed by (seed, world, draw); else [W * n_uniforms] doubles in [0, 1)
 * with n_uniforms >= PMP_SAMPLER_HEAD + n_stems * PMP_SAMPLER_PER_STEM (parity tests pass the draws in). */
int pmp_sample_worlds(pmp_ctx* ctx, const pmp_sampler_desc* desc, int32_t W, uint64_t seed, const double* uniforms_dev,
                      int32_t n_uniforms, float* stems_dev, float* boxes_dev, void* stream);
/* Adopt W worlds whose tables are already in device memory (every world: n_stems stems, n_boxes boxes; copied device
 * to device on `stream`, the caller may free its buffers afterwards).  The records are trusted, not validated. */
int pmp_set_worlds_dev(pmp_ctx* ctx, int32_t W, int32_t n_stems, int32_t n_boxes, int32_t n_slots, const float* stems_dev,
                       const float* boxes_dev, void* stream);

/* Constants the reference hard-codes in closures / module scope. */
typedef struct pmp_params {
    double resolution;       /* DEFAULT_RESOLUTION = radians(3), pybullet_tools/utils.py:3660 */
    float deflection_limit;  /* env.deflection_limit */
    float alpha;             /* 0.1, kuka_primitives.py:476 */
    float max_step;          /* per-iteration joint step cap of the plant solve (rad) */
    float reg_eps;           /* Gauss-Newton regularisation (m^2) */
    int32_t iterations;      /* K Gauss-Newton steps per stem */
    int32_t mode;            /* PMP_MODE_* */
    int32_t dense;           /* 1 = disable data-dependent early exits (every unit runs all K iterations
                                on every stem: results identical, used to measure the FP32 roofline).  Applies to
                                full-output pmp_validate_edges launches whose world tables fit shared memory (its own
                                kernel instantiation); every other shape runs the production kernel */
    uint32_t gate_seed;      /* seed of the on-device gate (below) */
    float gate_probability;  /* 0 = off.  Otherwise first_hit treats "deflection exceeded" at (edge e, step s,
                                world w) as a violation only if hash32(gate_seed, e, s, w) < gate_probability * 2^32:
                                the reference's 5 % random pass-through (pybullet_tools/utils.py:3843, probability 0.95)
                                with a counter-based generator, so that it can run inside the kernel.  The raw step
                                masks are never gated; replay them on the host to reproduce the reference's
                                np.random stream exactly. */
    float arm_lag_gain;      /* 0 = the arm is at the commanded configuration.  Otherwise the per-sub-step position gain
                                of the reference's arm motor (0.01, env.py:154,170): collision and deflection tests of
                                step i see a_i = a_{i-1} + c (q_i - a_{i-1}), a_0 = q_0, c = 1 - (1 - gain)^substeps,
                                the joint-limit test sees q_i (pybullet_tools/utils.py:3980) */
    int32_t arm_lag_substeps;/* simulation sub-steps per interpolation step (50, env.py:157,216) */
} pmp_params;

/* The counter-based hash of the on-device gate (murmur3 finaliser over the mixed coordinates). */
#ifdef __CUDACC__
#define PMP_HOST_DEVICE __host__ __device__
#else
#define PMP_HOST_DEVICE
#endif
static inline PMP_HOST_DEVICE uint32_t pmp_gate_hash(uint32_t seed, uint32_t e, uint32_t s, uint32_t w) {
    uint32_t h = seed ^ (e * 0x9E3779B1u) ^ (s * 0x85EBCA77u) ^ (w * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

int pmp_create(pmp_ctx** out, int device);
/* after pmp_set_robot (which clears any hulls); NULL or n_hulls == 0 switches back to the sphere model */
int pmp_set_robot_hulls(pmp_ctx* ctx, const pmp_hull_desc* hulls);
int pmp_destroy(pmp_ctx* ctx);
const char* pmp_last_error(pmp_ctx* ctx); /* ctx may be NULL: last creation error */
int pmp_set_robot(pmp_ctx* ctx, const pmp_robot_desc* robot);
int pmp_set_worlds(pmp_ctx* ctx, const pmp_world_desc* worlds);
int pmp_set_params(pmp_ctx* ctx, const pmp_params* params);
int pmp_num_worlds(pmp_ctx* ctx);
int pmp_max_stems(pmp_ctx* ctx);

/* Batched forward kinematics: q_dev [B*D] -> link frame position [B*L*3], quaternion xyzw [B*L*4],
 * COM position [B*L*3] (any output may be NULL).
 * Replaces set_joint_positions + get_link_state per link (pybullet_tools/utils.py:1996, 2197-2203). */
int pmp_fk(pmp_ctx* ctx, const float* q_dev, int32_t B, float* link_pos_dev, float* link_quat_dev,
           float* com_pos_dev, void* stream);

/* Explicit configurations x all worlds.
 *   flags_dev [B*W] u8 (PMP_FLAG_*), defl_dev [B*W*max_stems] chained deflections (NULL ok),
 *   theta_dev [B*W*max_stems*2] stem joint angles (NULL ok), ee_dev [B*3] (NULL ok).
 * Replaces MultiPlantWorld.step(q) + per-world collision_fn(q) + plant_rep.observe_all()
 * (env.py:212-238; pybullet_tools/utils.py:3836-3855; representation.py:268-294). */
int pmp_check_configs(pmp_ctx* ctx, const float* q_dev, int32_t B, uint8_t* flags_dev, float* defl_dev,
                      float* theta_dev, float* ee_dev, void* stream);

/* Edge batches: E straight joint-space edges q0 -> q1, interpolated on the device at `resolution`
 * (n = int(|dq/res|_2)+1 configurations; backward=0: q1 included, q0 excluded; backward=1: the reversed
 * sequence of asymmetric_extend, q0 included, q1 excluded), every configuration checked in every world.
 *   n_steps_dev   [E]      n (unclamped)
 *   first_hit_dev [E]      first step with any flag in any world, min(n, max_steps) if none
 *   max_defl_dev  [E*W]    max chained deflection over steps and stems          (NULL ok)
 *   over_dev      [E*W]    sum over steps and stems of (deflection - limit)+    (NULL ok)
 *   ee_len_dev    [E]      end-effector path length along the edge               (NULL ok)
 *   cost_dev      [E]      ee_len + alpha * mean_w over[e,w]                      (NULL ok)
 *   rigid_mask_dev / exceed_mask_dev [E*W*C] u32, C = ceil(max_steps/32): bit (s%32) of word s/32 is the
 *                          raw flag of step s in world w -- what the host needs to replay the reference's
 *                          5 % random pass-through in its sequential order     (NULL ok)
 * Steps >= max_steps are not evaluated.
 * A call with every optional output NULL asks the planner's accept/reject question only; the kernel then stops the
 * way the reference's loops do (steps after the first violating one and the remaining worlds are not evaluated,
 * primitives.py:225-241, env.py:230-238).  first_hit / n_steps are identical to those of a full call.
 * Replaces the loop of extend_towards_with_angle_constraint_multi_world (motion_planners/motion_planners/
 * primitives.py:196-255), check_forward_path_multi_world (rrt_connect_with_angle_constraints_v2.py:143-198)
 * and the cost accumulation of Command.execute_multi_world (pybullet_tools/kuka_primitives.py:467-521).
 *
 * edge_base: index of edge 0 of this call in the caller's whole batch.  It enters only the on-device gate's hash
 * (seed, edge_base + e, step, world), so that a batch validated in pieces (chunks, shards over ranks) draws exactly
 * the pass-through decisions of the unsplit batch; pass 0 for a whole batch. */
int pmp_validate_edges(pmp_ctx* ctx, const float* q0_dev, const float* q1_dev, int32_t E, int32_t max_steps,
                       int32_t backward, int32_t* first_hit_dev, int32_t* n_steps_dev, float* max_defl_dev,
                       float* over_dev, float* ee_len_dev, float* cost_dev, uint32_t* rigid_mask_dev,
                       uint32_t* exceed_mask_dev, void* stream, int32_t edge_base);

/* Same with HOST buffers: copies q0/q1 to the device, runs the kernel, copies the requested outputs back
 * and synchronises.  This is the call a CPU-side planner makes. */
int pmp_validate_edges_host(pmp_ctx* ctx, const float* q0, const float* q1, int32_t E, int32_t max_steps,
                            int32_t backward, int32_t* first_hit, int32_t* n_steps, float* max_defl, float* over,
                            float* ee_len, float* cost, uint32_t* rigid_mask, uint32_t* exceed_mask);

/* Host-buffer version of pmp_check_configs (used by the scalar collision_fn adapter). */
int pmp_check_configs_host(pmp_ctx* ctx, const float* q, int32_t B, uint8_t* flags, float* defl, float* theta,
                           float* ee);

/* Launch statistics of the last pmp_validate_edges / pmp_check_configs call. */
typedef struct pmp_launch_info {
    int32_t grid, block, smem_bytes, regs_per_thread;
    int32_t worlds_in_smem;  /* 1 = world tables staged in shared memory, 0 = read through L1 */
    int32_t n_sphere_pad, n_slot_pad;
    int64_t kernel_launches; /* cumulative count of this library's kernel launches on the context */
} pmp_launch_info;
int pmp_get_launch_info(pmp_ctx* ctx, pmp_launch_info* out);

/* Per-kernel device timing of pmp_validate_edges (bench.py's roofline leg).  With `on` != 0 every
 * pmp_validate_edges call records CUDA events on ITS stream around the hull pre-pass and around the edge kernel;
 * pmp_get_kernel_times waits for the last timed call and returns both durations in milliseconds (0 for a kernel
 * that did not run).  Off by default: four event records per call are not free on the scalar adapter path. */
int pmp_set_kernel_timing(pmp_ctx* ctx, int32_t on);
int pmp_get_kernel_times(pmp_ctx* ctx, float* rigid_ms, float* edge_ms);

/* FP32 FMA-pipe micro-benchmark (dependent FMA chains, 8 independent chains per thread) used as the
 * measured roofline denominator: returns achieved TFLOP/s in *tflops and the kernel time in *ms. */
int pmp_measure_fp32_peak(pmp_ctx* ctx, int32_t repeats, double* tflops, double* ms);

/* ---- batched tree extension: the caller of the edge path (motion_planners/motion_planners/primitives.py:196-255,
 * rrt_connect_with_angle_constraints_v2.py:240-290), many proposals per launch ------------------------------- */

/* Nearest tree node of each target: argmin over nodes of distance_fn(node, target) (weights 1, 2-norm of
 * difference_fn, continuous joints wrapped), first minimum wins -- motion_planners/motion_planners/utils.py:47-53
 * as called from primitives.py:207; pybullet_tools/utils.py:3604-3627.
 *   nodes_dev [T*D], targets_dev [B*D]; size_dev (NULL ok): device int holding the live node count (<= T)
 *   index_dev [B]; dist_dev [B] float64 distance (NULL ok); q_from_dev [B*D] the chosen nodes (NULL ok)
 * float64 arithmetic on the float32 data in joint order: indices are bit-identical to the oracle. A target with a
 * NaN coordinate gets node 0. */
int pmp_nearest(pmp_ctx* ctx, const float* nodes_dev, int32_t T, const int32_t* size_dev, const float* targets_dev,
                int32_t B, int32_t* index_dev, double* dist_dev, float* q_from_dev, void* stream);

/* Configuration number step_dev[e] of edge e (0-based, the numbering of pmp_validate_edges), out_dev [E*D].
 * The arithmetic is the edge kernel's, so these are exactly the configurations that were validated
 * (get_refine_fn / get_extend_fn, pybullet_tools/utils.py:3641-3676). */
int pmp_edge_points(pmp_ctx* ctx, const float* q0_dev, const float* q1_dev, int32_t E, const int32_t* step_dev,
                    int32_t backward, float* out_dev, void* stream);

/* A search tree in caller-owned device memory. Node i lies on the edge parent[i] -> via[i] (its step via_steps[i]-1,
 * all earlier steps of that edge validated). */
typedef struct pmp_tree {
    float* nodes;        /* [capacity*D] */
    int32_t* parent;     /* [capacity], -1 for the root */
    float* via;          /* [capacity*D] */
    int32_t* via_steps;  /* [capacity] */
    int32_t* size;       /* [1] live node count */
    int32_t capacity;
    int32_t reserved;
} pmp_tree;

/* One batched extension: for each target, the nearest node is extended towards it with the full multi-world check
 * (forward sequence, q_near excluded, target included) and the last safe configuration -- if any step was safe --
 * is appended to the tree in batch order.  Five kernels on `stream`, no host synchronisation.
 *   size_bound     host upper bound of *tree->size (for grid sizing), 1 <= size_bound <= capacity
 *   new_index_dev  [B] slot of the appended node, -1 if none
 *   reached_dev    [B] 1 if the target itself was reached (first_hit >= n_steps)
 *   new_cfg_dev    [B*D] appended configuration, NaN where none (NULL ok) -- the targets of the connect step
 *   stats_dev      [4] nodes appended, targets reached, first reached batch index (B if none), capacity overflow
 * Replaces extend_towards_with_angle_constraint_multi_world (primitives.py:196-255) for B proposals at once. */
int pmp_tree_extend(pmp_ctx* ctx, const pmp_tree* tree, int32_t size_bound, const float* targets_dev, int32_t B,
                    int32_t max_steps, int32_t* new_index_dev, int32_t* reached_dev, float* new_cfg_dev,
                    int32_t* stats_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMP_EDGECHECK_H */
