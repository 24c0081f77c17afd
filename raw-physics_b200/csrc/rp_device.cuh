// rp_device.cuh -- device data model of a batch (n_worlds instances of one scene template in HBM) and the kernels of
// the XPBD frame step. One launch works on ALL worlds; the per-substep sequence is
//   k_integrate -> [k_bounds] -> k_cull -> k_transform -> k_gjk [k_gjk_warp] -> k_epa [k_epa_warp] -> k_manifold
//   -> k_solve_pos -> k_solve_vel
// preceded once per frame by k_broad_cells / k_broad_scan / k_broad_write / k_islands / k_schedule and followed by k_derive
// (pbd_simulate_with_constraints, src/physics/pbd.cpp:468-747). See DESIGN.md for the layout and the roofline of each.
#ifndef RP_DEVICE_CUH
#define RP_DEVICE_CUH

#include <cuda_runtime.h>

#include "rp_solve.h"

namespace rp {

// per-world, per-body dynamic state (entity.h:21-46): x[3] q[4] v[3] w[3] prev_x[3] prev_q[4] prev_v[3] prev_w[3], stored
// world-minor: component f of body b in world w at dyn[(b * RP_DYN_DOUBLES + f) * WS + w] (see rp_kernels.cuh "layout")
#define RP_DYN_DOUBLES 26
// Physical parameters of a body, shared by all worlds (template) and de-duplicated: bodies with bit-identical mass,
// tensors and coefficients share one record (W256: 2 classes for 257 bodies), so the lanes of a warp that work on
// different bodies of the same class read the SAME addresses (one broadcast wavefront instead of 32 scattered ones).
struct BodyClass {
	real inv_mass;
	M3 inertia, inv_inertia;
	real mu_s, mu_d, rest;
	real ii_bound;   // tensor_bound(inv_inertia), see solve_contact
};
// per-body static parameters, shared by all worlds (template)
struct BodyStatic {
	real radius;
	real rfar;              // (radius + 0.1) * (1 + 1e-9): this body's share of the broadphase's axis-reject threshold
	int cls;                  // index into DevView::bclass
	int fixed, col0, ncol;
	int tv0, tvn, tn0, tnn;   // extent of the body's colliders in a world's transformed vertex / normal arrays
};
struct __align__(16) PairRec {  // one broadphase pair, expanded to collider granularity: bodies a < b, global collider indices ca, cb
	int a, b, ca, cb;
};
struct __align__(16) SolveItem {  // one unit of a sweep: a collider pair's manifold, with what the solver needs to start on it
	int w, a, b;      // world, bodies (a < b)
	int coff, cnt;    // the manifold's run in the world's contact buffer
	int pad;
	V3 normal;        // shared by the whole manifold
};
struct EpaOut {  // result of EPA (or the analytic sphere-sphere test) for one hit: work item of k_manifold
	V3 normal;
	real depth;
	int ok, pad;
	int sup_a, sup_b;  // support vertices of the two hulls along +normal / -normal (EPA's last support call), -1 when unknown
};

enum { CNT_PAIR_TESTS = 0, CNT_HITS = 1, CNT_CONTACTS = 2, CNT_BROAD_PAIRS = 3, CNT_LEVELS = 4, CNT_FRAMES = 5, CNT_CANDS = 6 };

struct DevView {
	int W, NB, NC, NJ, TV, TN;
	int WS;              // world stride of every world-minor array: W rounded up to a multiple of 32 (W itself when W < 32)
	int max_pairs, max_contacts, max_units;
	real lin_sleep, ang_sleep, sleep_time;
	// template
	const BodyStatic* bstat;
	const BodyClass* bclass;
	const ColliderDesc* cols;
	HullPool pool;
	const Joint* joints;
	const V3* force;
	const V3* torque;
	// per world, world-minor ([...][WS]) unless noted
	real* dyn;         // [NB][RP_DYN_DOUBLES][WS]
	int* active;         // [NB][WS]
	int* vstamp;         // [NB][WS] substep counter value at which the body's velocities were last derived (lazy k_derive)
	int* epoch;          // [1] substep counter, +1 per substep (k_substep_reset)
	real* deact;       // [NB][WS]
	real* tv;          // [TV][3][WS] transformed vertices (sphere: centre)
	real* tn;          // [TN][3][WS] transformed face normals
	PairRec* pairs;      // [max_pairs][WS]
	int* n_pairs;        // [W]
	int n_cells;         // broadphase cells: (row i, 32 consecutive j) pieces of the i < j triangle, in (i, j) order
	const int2* cells;   // [n_cells] (i, first j), template constant
	unsigned int* cell_mask;  // [n_cells][WS] bit k = body pair (i, j0 + k) is near
	int* cell_off;       // [n_cells][WS] collider pairs of the cell -> offset of its first pair in the world's list
	int* label;          // [W][NB] island labels (world-major scratch of k_islands)
	int* isl_flag;       // [W][NB] island "all members may sleep"
	int* last_level;     // [NB][WS] schedule scratch
	unsigned long long* colour_tab;  // [NB][WS] schedule scratch of the coloured order (when the table does not fit shared memory)
	int* pair_level;     // [max_pairs][WS] dependency level of each broadphase pair (0 = skipped this frame)
	int* lvl_hist;       // [max_levels + 2][WS] schedule scratch (per-world level histogram)
	int* geom_stamp;     // [NC][WS] substep counter value at which k_cull last asked for the collider's transformed geometry
	float* aabb;         // [NC][6][WS] world-space bounds of every collider (min xyz rounded down, max xyz rounded up)
	uint4* cands;        // [W * max_pairs] (world, pair, collider a, collider b) that survived the skip rule and the bounds cull
	unsigned int* cand_count;
	unsigned int* big_count;   // candidates with hulls too large to stage per thread: kept at the END of cands, taken by k_gjk_warp
	unsigned int cand_cap;     // W * max_pairs
	int split_bounds;          // bodies carry so much geometry that the bounds are computed per collider (k_bounds), not per body in k_integrate
	int split_big;             // the scene has such pairs: k_epa leaves them to k_epa_warp, k_manifold takes their supports from big_sup
	int2* big_sup;             // [W * max_pairs] per hit of a large pair: support vertices of the two hulls along +-normal (k_epa_warp)
	uint4* hits;         // [W * max_pairs] the colliding candidates' records, dense (one load tells EPA / clipping where their inputs are)
	real* simplex;     // [12][W * max_pairs] final GJK tetrahedron of each hit, as component planes (st_simplex)
	unsigned int* hit_count;
	EpaOut* epa_out;     // [W * max_pairs] per hit
	// level-major work lists shared by all worlds: the pairs of dependency level l that have contacts this substep
	int max_levels;
	int* lvl_cap;        // [max_levels + 2] pairs scheduled at level l over all worlds (per frame)
	int* lvl_off;        // [max_levels + 2] exclusive scan of lvl_cap
	int* lvl_fill;       // [max_levels + 2] pairs of level l with contacts (per substep)
	int* lvl_max;        // [1] deepest level of the frame over all worlds
	SolveItem* lvl_items;  // [W * max_pairs] level l occupies [lvl_off[l], lvl_off[l] + lvl_fill[l])
	// dataflow sweeps (pos_flow / vel_flow in rp_kernels.cuh): which levels of a body's chain of units are live this substep, and
	// how far each body's chain has got in the current sweep -- these stand in for the grid-wide barrier between levels
	int flow_mode;                   // 1: the sweeps run without grid barriers
	unsigned long long* body_live;   // [NB][WS] bit l: a unit of level l with contacts this substep touches this body (k_manifold; cleared by k_integrate)
	unsigned long long* body_done;   // [2][NB][WS] (pass << 6 | level) of the last unit finished on this body in the positional / velocity sweep
	const unsigned long long* joint_body_mask;  // [NB] bit l: a joint of level l touches this body (template constant)
	unsigned int* flow_cursor;       // [2] next unclaimed item of the positional / velocity sweep; [2] set when a lane gave up waiting
	// world-block sweeps (k_solve_block): the units of each WORLD that have contacts this substep, and each world block's
	// units sorted by level
	int block_mode;      // 1: the sweeps run per block of `worlds per block` consecutive worlds, CTA-scoped barriers between levels
	uint2* live;         // [max_pairs][WS] per world: (pair, level) of its units with contacts, in arrival order (k_manifold)
	int* n_live;         // [W]
	unsigned int* blk_items;  // [W * max_pairs] block b's region starts at (first world of b) * max_pairs: (pair << 6 | world in block)
	// template-constant schedule of the external constraints (they head the constraint array in every world)
	const int* joint_sched;   // [NJ] joints sorted by level
	const int* joint_lptr;    // [joint_levels + 1]
	const int* joint_last;    // [NB] level of the last joint touching each body (0 = none)
	const unsigned long long* joint_colours;  // [NB] coloured order: colours the joints of each body have taken (SchedEntry<true>)
	int joint_levels;
	V3* pair_normal;     // [max_pairs][WS]
	int* pair_coff;      // [max_pairs][WS]
	int* pair_ccnt;      // [max_pairs][WS]
	real* contacts;    // [max_contacts][8][WS] contact records (r1_lc, r2_lc, lambda_n, lambda_t), see contact_ptr
	int* n_contacts;     // [W]
	JointLambda* lambdas;  // [NJ][WS]
	int* status;         // [W]
	unsigned long long* counters;
	// parity instrumentation (rp_batch_step_logged)
	int dbg_world;
	V3* dbg_points;      // [max_contacts][2]
};

}  // namespace rp
#endif
