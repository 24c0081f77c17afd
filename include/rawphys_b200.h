/* rawphys_b200.h -- C ABI of librawphys_b200.so: the XPBD frame step of felipeek/raw-physics on NVIDIA B200 (sm_100a).
 *
 * The seam this library replaces is the reference's pair of free functions (src/physics/pbd.h:93-94)
 *
 *     void pbd_simulate(r64 dt, Entity** entities, u32 num_substeps, u32 num_pos_iters, boolean enable_collisions);
 *     void pbd_simulate_with_constraints(r64 dt, Entity** entities, Constraint* external_constraints,
 *                                        u32 num_substeps, u32 num_pos_iters, boolean enable_collisions);
 *
 * and everything they call (broad.cpp, collider.cpp, gjk.cpp, epa.cpp, clipping.cpp, support.cpp,
 * pbd_base_constraints.cpp, physics_util.cpp). A *scene* is the template the reference builds with
 * collider_convex_hull_create / collider_sphere_create / entity_create[_fixed] / pbd_*_constraint_init; a *batch* is
 * n_worlds independent instances of that scene resident in the HBM of one GPU; rp_batch_step is one
 * pbd_simulate_with_constraints call applied to every world. Plain pointers and sizes only; no global state; every
 * call returns a status code (0 = RP_OK) instead of aborting. There is NO CPU fallback: without a CUDA device every
 * batch call fails with RP_ERR_CUDA.
 *
 * Body ids are the order of rp_scene_add_body calls and equal the reference's eids when its eid_counter starts at 0.
 *
 * The reference's compile-time switches (pbd.cpp:12-16, pbd_base_constraints.cpp:4): ENABLE_SIMULATION_ISLANDS and the three
 * sleeping constants are run-time fields of rp_batch_cfg; USE_QUATERNIONS_LINEARIZED_FORMULAS stays a compile-time switch -- the
 * same sources built with -DRP_EXACT_QUATERNIONS give librawphys_b200_exactq.so (same ABI), the reference WITHOUT that define:
 * orientation updates by axis-angle quaternions through sin / cos (raw-physics_b200/build.py build_exactq).
 */
#ifndef RAWPHYS_B200_H
#define RAWPHYS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rp_scene rp_scene;
typedef struct rp_batch rp_batch;

enum {
	RP_OK = 0,
	RP_ERR_ARG = 1,      /* bad argument */
	RP_ERR_CUDA = 2,     /* CUDA runtime error or no device; rp_last_error() has the text */
	RP_ERR_CAPACITY = 3  /* a fixed-capacity device buffer overflowed; see rp_batch_get_status */
};

/* per-world status bits (rp_batch_get_status); a non-zero word means the reference would have aborted on an assert,
 * printed a warning, or the device ran out of a fixed capacity in that world */
enum {
	RP_ST_GJK_SIMPLEX_OVERFLOW = 1 << 0, /* gjk.cpp:24-26 */
	RP_ST_EPA_DEGENERATE = 1 << 1,       /* epa.cpp:41,72 */
	RP_ST_EPA_NO_CONVERGENCE = 1 << 2,   /* epa.cpp:233 */
	RP_ST_EPA_CAPACITY = 1 << 3,
	RP_ST_CLIP_CAPACITY = 1 << 4,
	RP_ST_EDGE_PARALLEL = 1 << 5,        /* clipping.cpp:279 */
	RP_ST_CONTACT_CAPACITY = 1 << 6,
	RP_ST_PAIR_CAPACITY = 1 << 7,
	RP_ST_SOLVER_SINGULAR = 1 << 8,      /* pbd_base_constraints.cpp:40,154 */
	RP_ST_NAN = 1 << 9
};

/* PBD_Axis_Type (src/physics/pbd.h:5-12) */
enum { RP_POSITIVE_X_AXIS = 0, RP_NEGATIVE_X_AXIS, RP_POSITIVE_Y_AXIS, RP_NEGATIVE_Y_AXIS, RP_POSITIVE_Z_AXIS, RP_NEGATIVE_Z_AXIS };

const char* rp_last_error(void);
/* number of CUDA devices visible, or 0 */
int rp_device_count(void);

/* ------------------------------------------------------------------------------------------------ scene template */
rp_scene* rp_scene_create(void);
void rp_scene_destroy(rp_scene* s);

/* collider_convex_hull_create (collider.cpp:194): triangle soup, vertices already scaled (3 doubles each), indices in
 * triples. The collider is queued for the NEXT rp_scene_add_body call. Returns its index within that body, or -1. */
int rp_scene_collider_hull(rp_scene* s, const double* vertices_xyz, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices);
/* Where rp_scene_collider_hull builds the hull topology from now on: cuda_device >= 0 = on that GPU (csrc/rp_hull.cuh: the
 * quadratic passes of collider_convex_hull_create as one thread per output row, the order-defining flood fill as one thread;
 * the arrays are identical to the host build's), -1 = on the host (the default). rp_scene_hull_build_stats: hulls built so far
 * (identical soups share one) and the time that took, milliseconds. */
int rp_scene_set_hull_device(rp_scene* s, int cuda_device);
int rp_scene_hull_build_stats(const rp_scene* s, int* hulls_built, double* milliseconds);
/* collider_sphere_create (collider.cpp:12) */
int rp_scene_collider_sphere(rp_scene* s, float radius);
/* entity_create / entity_create_fixed (entity.cpp:67-77): consumes the queued colliders. Returns the body id or -1. */
int rp_scene_add_body(rp_scene* s, const double position[3], const double rotation_xyzw[4], double mass, int fixed,
	double static_friction, double dynamic_friction, double restitution);

/* Adoption of objects the reference has ALREADY built (what the pbd_simulate shim does, raw-physics_b200/shim/pbd_b200.cpp):
 * a Collider_Convex_Hull as it lies in memory (src/physics/collider.h:19-29: vertices, faces[].elements / .normal,
 * vertex_to_faces, vertex_to_neighbors, face_to_neighbors) as CSR arrays (ptr arrays have count + 1 entries), queued for
 * the next body like rp_scene_collider_hull; and an Entity's derived parameters (src/entity.h:27-40: inverse_mass,
 * inertia_tensor, inverse_inertia_tensor row-major, bounding_sphere_radius, fixed, coefficients) taken as they are
 * instead of being recomputed by entity_create_ex. Return the collider index / body id, or -1. */
int rp_scene_collider_hull_topology(rp_scene* s, const double* vertices_xyz, uint32_t n_vertices, const double* face_normals_xyz,
	uint32_t n_faces, const uint32_t* face_ptr, const uint32_t* face_idx, const uint32_t* v2f_ptr, const uint32_t* v2f_idx,
	const uint32_t* v2n_ptr, const uint32_t* v2n_idx, const uint32_t* f2n_ptr, const uint32_t* f2n_idx);
int rp_scene_add_body_params(rp_scene* s, const double position[3], const double rotation_xyzw[4], double inverse_mass,
	const double inertia[9], const double inverse_inertia[9], double bounding_sphere_radius, int fixed, double static_friction,
	double dynamic_friction, double restitution);

/* pbd_positional_constraint_init ... pbd_spherical_joint_constraint_init (pbd.cpp:18-79); return the constraint index */
int rp_scene_add_positional_constraint(rp_scene* s, int e1, int e2, const double r1_lc[3], const double r2_lc[3], double compliance,
	const double distance[3]);
int rp_scene_add_mutual_orientation_constraint(rp_scene* s, int e1, int e2, double compliance);
int rp_scene_add_hinge_joint_constraint(rp_scene* s, int e1, int e2, const double r1_lc[3], const double r2_lc[3], double compliance,
	int e1_aligned_axis, int e2_aligned_axis, int limited, int e1_limit_axis, int e2_limit_axis, double lower_limit, double upper_limit);
int rp_scene_add_spherical_joint_constraint(rp_scene* s, int e1, int e2, const double r1_lc[3], const double r2_lc[3], int e1_swing_axis,
	int e2_swing_axis, int e1_twist_axis, int e2_twist_axis, double swing_lower, double swing_upper, double twist_lower, double twist_upper);

int rp_scene_num_bodies(const rp_scene* s);
/* static per-body parameters as entity_create_ex computes them, 25 doubles per body:
 * inverse_mass, inertia[9], inverse_inertia[9], bounding_sphere_radius, mu_s, mu_d, restitution, fixed, n_colliders */
#define RP_PARAM_STRIDE 25
int rp_scene_get_params(const rp_scene* s, double* out);
/* hull topology in the reference's order (for parity checks): sizes = V, F, sum face elems, sum v2f, sum v2n, sum f2n
 * (V = -1 for a sphere); the dump fills CSR arrays whose ptr arrays have count + 1 entries */
int rp_scene_hull_sizes(const rp_scene* s, int body, int collider, int32_t out6[6]);
int rp_scene_hull_dump(const rp_scene* s, int body, int collider, double* verts, double* normals, uint32_t* face_ptr, uint32_t* face_idx,
	uint32_t* v2f_ptr, uint32_t* v2f_idx, uint32_t* v2n_ptr, uint32_t* v2n_idx, uint32_t* f2n_ptr, uint32_t* f2n_idx);

/* The scene as it was described -- what a caller needs to build the same scene somewhere else (the parity tests feed the
 * oracle with it). body_desc: position[3], rotation xyzw[4], mass, fixed, static friction, dynamic friction, restitution,
 * number of colliders, 3 reserved. A sphere collider reports nverts = nidx = 0 and its radius; a hull its input soup.
 * joint_desc: ints = type (Constraint_Type, pbd.h:14-20), e1, e2, limited, the four axis selectors; vals = r1_lc[3], r2_lc[3],
 * distance[3], compliance, lower, upper, lower2, upper2. initial_state: RP_STATE_STRIDE doubles per body. */
int rp_scene_initial_state(const rp_scene* s, double* out);
int rp_scene_body_desc(const rp_scene* s, int body, double out16[16]);
int rp_scene_collider_soup_size(const rp_scene* s, int body, int collider, uint32_t* nverts, uint32_t* nidx, float* radius);
int rp_scene_collider_soup(const rp_scene* s, int body, int collider, double* verts_xyz, uint32_t* indices);
int rp_scene_num_joints(const rp_scene* s);
int rp_scene_joint_desc(const rp_scene* s, int joint, int32_t ints8[8], double vals14[14]);

/* ---------------------------------------------------------------------------------------------------- built-in scenes */
/* The init() halves of the reference's 14 examples (src/examples/<name>.cpp) and the benchmark worlds built from the same
 * pieces ("w256", "pile", "tumble", "spheres"), as scene templates; `info` receives what the example's update() passes to
 * pbd_simulate (substeps, positional iterations, collisions) and its gravity. `params` (optional, 0 = default) per scene:
 * stack {cubes}; w256 {stacks_x, stacks_z, height}; brick_wall {rows, cols}; cube_storm {n}; spot_storm {n, seed};
 * pile {n_side, seed, spacing, layers}; tumble {n, seed}; spheres {n}. `perturb` != 0 gives the joint scenes (hinge_joints, arm,
 * triple_pendula, rott_pendulum), which are at rest until a user interacts, initial angular velocities. `mesh_dir` = directory of
 * the <mesh>.f32 triangle soups (NULL: assets/meshes next to the library). Returns NULL on failure (rp_example_error()). */
typedef struct {
	uint32_t substeps, pos_iters;
	int32_t collisions;
	int32_t reserved0;
	double gravity;
} rp_example_info;
int rp_example_count(void);
const char* rp_example_name(int index);
const char* rp_example_error(void);
rp_scene* rp_example_create(const char* name, const double* params, uint32_t n_params, int perturb, const char* mesh_dir_or_null,
	rp_example_info* info_or_null);
/* the same with the hulls built on a GPU (rp_scene_set_hull_device; -1 = host) */
rp_scene* rp_example_create_on(const char* name, const double* params, uint32_t n_params, int perturb, const char* mesh_dir_or_null,
	rp_example_info* info_or_null, int hull_cuda_device);

/* ------------------------------------------------------------------------------------------------------- batches */
/* Order of the Gauss-Seidel sweeps over a world's constraints (rp_batch_cfg.solve_order).
 * RP_ORDER_REFERENCE: the reference's array order (external constraints, then contacts in broadphase-pair order,
 *   pbd.cpp:580-620), run as dependency levels -- results identical to the reference's. The default.
 * RP_ORDER_COLOURED: a greedy colouring of the constraint graph rebuilt every frame; colours run in ascending order. Far
 *   fewer sequential stages for one large scene (a 32 x 32 brick wall: ~10 instead of ~150), trajectories equal to the
 *   reference's only within solver accuracy. */
enum { RP_ORDER_REFERENCE = 0, RP_ORDER_COLOURED = 1 };

typedef struct {
	uint32_t max_pairs_per_world;     /* broadphase (collider-)pair capacity; 0 = derive from the initial poses */
	uint32_t max_contacts_per_world;  /* contact capacity per substep; 0 = derive */
	uint32_t disable_cull;            /* 1 = run GJK on every broadphase pair (the exact-safe bounds cull is on by default) */
	uint32_t solve_order;             /* RP_ORDER_REFERENCE (default) or RP_ORDER_COLOURED */
	uint32_t sweep_block_worlds;      /* worlds per CTA of the world-block Gauss-Seidel sweeps; 0 = choose (level-major sweeps for small batches) */
	uint32_t large_scene;             /* per-frame prologue for ONE LARGE SCENE (uniform-grid broadphase, union-find islands, parallel graph
	                                     colouring): 0 = when a world has >= 4096 bodies (>= 1024 in the coloured order), 1 = never, 2 = always */
	uint32_t disable_islands;         /* 1 = the reference built without ENABLE_SIMULATION_ISLANDS (pbd.cpp:12): no islands, nothing sleeps */
	uint32_t sweep_form;              /* how the Gauss-Seidel sweeps keep the reference's order (same results either way): 0 = choose (dataflow --
	                                     a unit waits for the previous live unit of each of its two bodies, no grid barriers -- for contact-only scenes:
	                                     batches of at least 64 worlds and one large / coloured scene; grid barriers between levels otherwise), 1 = grid
	                                     barriers (a batch of worlds that differ from each other measured 3 % faster with them), 2 = dataflow wherever
	                                     it applies (scenes with joints included) */
	double linear_sleeping_threshold;   /* pbd.cpp:13, default 0.10 */
	double angular_sleeping_threshold;  /* pbd.cpp:14, default 0.10 */
	double deactivation_time;           /* pbd.cpp:15, default 1.0 */
} rp_batch_cfg;
void rp_batch_cfg_default(rp_batch_cfg* cfg);

/* 1 <= n_worlds <= 65535 per batch (several batches may live on one device, each with its own stream); RP_ERR_ARG otherwise.
 * Device memory is allocated here and nowhere else: about 1.1 MB per world of 257 bodies. */
int rp_batch_create(const rp_scene* scene, uint32_t n_worlds, int cuda_device, const rp_batch_cfg* cfg_or_null, rp_batch** out);
/* Entity counts that change mid-run (examples_util_throw_object, examples_util.cpp:52-95, creates bodies between frames;
 * entity_destroy, entity.cpp:67-77, removes them): a new batch for `scene` -- same world count, same device as `src` -- whose
 * body i takes its state in every world from body new_from_old[i] of `src` (device to device), or starts from the scene's initial
 * state where new_from_old[i] = -1. `src` stays valid and is destroyed separately. Between frames only. */
int rp_batch_create_from(const rp_scene* scene, rp_batch* src, const int32_t* new_from_old, const rp_batch_cfg* cfg_or_null, rp_batch** out);
void rp_batch_destroy(rp_batch* b);
uint32_t rp_batch_num_worlds(const rp_batch* b);
uint32_t rp_batch_num_bodies(const rp_batch* b);

/* External forces for the coming step(s), shared by every world: the sums of entity_add_force(e, position, force,
 * local_coords=false) calls (entity.cpp:176-193) in call order, as calculate_external_force/torque form them. */
int rp_batch_clear_forces(rp_batch* b);
int rp_batch_add_force(rp_batch* b, int body, const double position[3], const double force[3]);
/* the examples' gravity idiom: (0, -g * 1.0 / inverse_mass, 0) at the centre of every body (stack.cpp:93-96) */
int rp_batch_add_gravity(rp_batch* b, double g);

/* pbd_simulate_with_constraints for every world. The whole frame -- broadphase, islands + sleeping, schedule, then the
 * substeps -- is enqueued on the batch's CUDA stream as ONE graph launch (captured once per (dt, substeps, iters,
 * collisions)) and the call returns without waiting for it; nothing about the frame is read back by the host. */
int rp_batch_step(rp_batch* b, double dt, uint32_t num_substeps, uint32_t num_pos_iters, int enable_collisions);
/* Waits for the batch's stream. Returns RP_ERR_CAPACITY if a fixed-capacity device buffer (broadphase pairs, contacts per
 * substep: rp_batch_cfg, defaults max(128, 2 * pairs of the initial poses + 64) pairs and max(256, 8 * bodies) contacts per
 * world) ran out in some world since creation or the last rp_batch_clear_status: pairs / contacts were dropped there and the
 * world no longer follows the reference (rp_batch_get_status says which). rp_batch_step_host and rp_batch_download_state
 * report the same way, after doing their work. */
int rp_batch_sync(rp_batch* b);
/* kernel nodes of the CUDA graph one rp_batch_step launches (0 before the first step) */
int rp_batch_graph_kernels(const rp_batch* b);
/* `frames` consecutive steps timed on the device with CUDA events on the batch's stream (milliseconds). */
int rp_batch_run(rp_batch* b, uint32_t frames, double dt, uint32_t num_substeps, uint32_t num_pos_iters, int enable_collisions,
	float* gpu_ms_out);

/* Per-body dynamic state record, RP_STATE_STRIDE doubles: world_position[3], world_rotation xyzw[4],
 * linear_velocity[3], angular_velocity[3], active (0/1), deactivation_time, previous_linear_velocity[3],
 * previous_angular_velocity[3] (entity.h:21-46). Host arrays are [world][body][RP_STATE_STRIDE]. */
#define RP_STATE_STRIDE 21
int rp_batch_upload_state(rp_batch* b, uint32_t first_world, uint32_t n_worlds, const double* host_state);
int rp_batch_download_state(rp_batch* b, uint32_t first_world, uint32_t n_worlds, double* host_state);
/* same state for every world (the scene's initial poses are loaded at creation) */
int rp_batch_broadcast_state(rp_batch* b, const double* host_state_one_world);
/* Host-buffer form of the step (what a pbd_simulate shim calls): upload `host_state_in` (all worlds), step, download
 * into `host_state_out`, synchronise. Either pointer may be NULL to skip that copy. */
int rp_batch_step_host(rp_batch* b, const double* host_state_in, double* host_state_out, double dt, uint32_t num_substeps,
	uint32_t num_pos_iters, int enable_collisions);

/* status words accumulate (bitwise or) from batch creation or the last rp_batch_clear_status */
int rp_batch_get_status(rp_batch* b, int32_t* status_per_world);
int rp_batch_clear_status(rp_batch* b);
/* cumulative work counters since creation: [0] narrowphase collider-pair tests, [1] GJK hits (EPA runs),
 * [2] contacts, [3] broadphase pairs (per frame, summed), [4] solver levels (per frame, summed), [5] frames,
 * [6] GJK runs (pair tests that survived the bounds cull) */
int rp_batch_get_counters(rp_batch* b, uint64_t out8[8]);

/* Parity instrumentation: one frame of ONE world stepped substep by substep, logging what the reference's
 * colliders_get_contacts (collider.cpp:560) returns for every narrowphase pair in pair order: calls_out gets 4 u32 per
 * call (e1, e2, contact count, index of first contact), contacts_out 9 doubles per contact (point1, point2, normal).
 * All other worlds are stepped too. Returns counts through n_calls / n_contacts. */
int rp_batch_step_logged(rp_batch* b, double dt, uint32_t num_substeps, uint32_t num_pos_iters, int enable_collisions, uint32_t world,
	uint32_t* calls_out, uint32_t max_calls, double* contacts_out, uint32_t max_contacts, uint32_t* n_calls, uint32_t* n_contacts);
/* broadphase pairs of one world for its current poses (broad_get_collision_pairs, broad.cpp:6): (e1, e2) body ids */
int rp_batch_broad_pairs(rp_batch* b, uint32_t world, uint32_t* pairs_out, uint32_t max_pairs, uint32_t* n_pairs);

/* the schedule of one world for its current poses: every broadphase collider pair (bodies a < b) with the dependency level
 * (RP_ORDER_REFERENCE) or colour (RP_ORDER_COLOURED) the sweeps run it at; 0 = skipped (both sides fixed or asleep, pbd.cpp:594) */
int rp_batch_pair_levels(rp_batch* b, uint32_t world, uint32_t* pairs_out, int32_t* levels_out, uint32_t max_pairs, uint32_t* n_pairs);

/* Profiling aids (bench.py). Kernel families of one frame, in launch order. */
enum {
	RP_K_BROAD = 0, RP_K_ISLANDS, RP_K_SCHEDULE, RP_K_INTEGRATE, RP_K_CULL, RP_K_GJK, RP_K_MANIFOLD, RP_K_SOLVE_POS, RP_K_DERIVE,
	RP_K_SOLVE_VEL, RP_K_EPA, RP_NUM_KERNEL_FAMILIES
};
/* device milliseconds per kernel family over `frames` un-graphed frames (CUDA events at every kernel boundary) */
int rp_batch_profile(rp_batch* b, uint32_t frames, double dt, uint32_t num_substeps, uint32_t num_pos_iters, int enable_collisions,
	float ms_out[RP_NUM_KERNEL_FAMILIES]);
/* measured FP64 CUDA-core rate, TFLOP/s: out2[0] = DMUL+DADD chains (no contraction, as this library is built),
 * out2[1] = DFMA chains */
int rp_measure_fp64_peak(int cuda_device, double out2[2]);

#ifdef __cplusplus
}
#endif
#endif
