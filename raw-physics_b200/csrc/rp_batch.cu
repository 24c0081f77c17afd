// rp_batch.cu -- the C ABI of include/rawphys_b200.h: scene templates on the host, world batches in device memory,
// frame stepping through a CUDA graph. No CPU fallback: every batch entry point needs a CUDA device.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/rawphys_b200.h"
#include "rp_kernels.cuh"
#include "rp_large.cuh"
#include "rp_scene.h"
#include "rp_hull.cuh"

using namespace rp;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
	g_err = msg;
	return code;
}

#define RP_CUDA(call)                                                                                                  \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if (e_ != cudaSuccess) {                                                                                       \
			cudaGetLastError();                                                                                        \
			return fail(RP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                              \
		}                                                                                                              \
	} while (0)

#ifndef RP_CULL_CTAS_PER_SM
#define RP_CULL_CTAS_PER_SM 16  // k_cull grid: CTAs of 8 warps per SM's worth of pair indices (a warp walks the rest in trips)
#endif
#define RP_COLOUR_ROUNDS 40  // independent-set rounds of the parallel colouring enqueued per frame (a round that has nothing left to do returns at once)
#define RP_SCHED_SMEM_MAX (160 * 1024)  // dynamic shared memory k_schedule<true> may ask for (opted in at batch creation)

struct GraphKey {
	double dt;
	uint32_t substeps, iters;
	int collisions;
	bool operator==(const GraphKey& o) const {
		return dt == o.dt && substeps == o.substeps && iters == o.iters && collisions == o.collisions;
	}
};

struct rp_batch {
	Scene scene;  // private copy of the template (forces are accumulated here)
	DevView d;
	int device = 0;
	cudaStream_t stream = 0;
	cudaEvent_t ev0 = 0, ev1 = 0;
	std::vector<void*> allocs;
	V3* force_dev = 0;
	V3* torque_dev = 0;
	bool forces_dirty = true;
	double* rec_dev = 0;  // staging for state records, [W][NB][RP_STATE_STRIDE]
	int cull_chunks = 1;
	int sm_count = 148;
	unsigned int pos_grid = 148, vel_grid = 148;  // resident CTAs of the cooperative sweep kernels
	int cull = 1;            // exact-safe bounds cull before GJK (rp_batch_cfg.disable_cull turns it off)
	int coloured = 0;        // rp_batch_cfg.solve_order == RP_ORDER_COLOURED
	int live_lists = 0;        // the scene can have deep, mostly empty schedules (compound bodies): the sweeps list the levels with work
	size_t live_smem = 0;      // dynamic shared memory of the sweep kernels for that list
	unsigned int transform_slices = 1;  // gridDim.z of k_transform: threads that share one collider's vertices and normals
	bool has_big_pairs = false;   // some collider pair is too large for k_gjk's per-thread staging: k_gjk_warp is launched too
	bool no_islands = false;      // rp_batch_cfg.disable_islands: nothing ever falls asleep (the reference without ENABLE_SIMULATION_ISLANDS)
	bool no_restitution = false;  // every body's restitution coefficient is zero (k_integrate's store_velocities)
	// one large scene (rp_large.cuh): uniform-grid broadphase, union-find islands, parallel colouring
	bool large = false;
	GridView grid;
	ColourView col;
	int grid_tiles_table = 0, grid_tiles_rows = 0;
	// CTAs per SM the per-pair narrowphase kernels are launched with (grid-stride loops over the work lists, whose lengths only
	// the device knows): a multiple of what is resident at once, so that an empty or short list costs one wave of CTAs
	unsigned int grid_gjk = 8, grid_epa = 5, grid_manifold = 4;  // = resident CTAs per SM: measured on config 5 (1.82 -> 1.35 ms/frame vs 16 each) and W256 (no change)
	size_t expected_pairs = 0;    // worlds x collider pairs of the initial poses: what sizes the narrowphase grids
	int sweep_wpb = 0;            // worlds per CTA of the world-block sweeps (k_solve_block); 0 = level-major cooperative sweeps
	size_t sweep_smem = 0;        // dynamic shared memory of k_solve_block (its per-level cursors)
	std::vector<int> joint_level;  // template-constant levels of the external constraints
	int* overflow_dev = 0;        // [1] OR of the capacity bits of every world's status word (k_status_overflow)
	int* overflow_pin = 0;        // pinned host copy
	int graph_kernels = 0;        // kernel nodes of the captured frame graph
	bool have_graph = false;
	GraphKey graph_key;
	cudaGraph_t graph = 0;
	cudaGraphExec_t graph_exec = 0;
};

template <class T>
static int dev_alloc(rp_batch* b, T** out, size_t n, bool zero = true) {
	void* p = 0;
	size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
	RP_CUDA(cudaMalloc(&p, bytes));
	b->allocs.push_back(p);
	if (zero) RP_CUDA(cudaMemsetAsync(p, 0, bytes, b->stream));
	*out = (T*)p;
	return RP_OK;
}
template <class T>
static int dev_upload(rp_batch* b, const T** out, const std::vector<T>& v) {
	T* p = 0;
	int rc = dev_alloc(b, &p, v.size());
	if (rc) return rc;
	if (!v.empty()) RP_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, b->stream));
	*out = p;
	return RP_OK;
}

// CTAs for one thread per (item, world) (flat_item_world)
static dim3 flat_grid(const rp_batch* b, size_t n_items, unsigned int threads) {
	const size_t W = (size_t)b->d.W;
	if (W >= 2 * (size_t)threads && n_items > 0 && n_items < 65536) return dim3((unsigned int)n_items, (unsigned int)((W + threads - 1) / threads));
	return dim3((unsigned int)((n_items * W + threads - 1) / threads));
}

// launches `kernel` as a cooperative grid (cg::this_grid().sync() inside); capturable into the frame graph
template <class... Args>
static cudaError_t launch_cooperative(void (*kernel)(Args...), unsigned int grid, unsigned int block, size_t smem, cudaStream_t stream,
	Args... args) {
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr;
	attr.id = cudaLaunchAttributeCooperative;
	attr.val.cooperative = 1;
	cfg.attrs = &attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static std::vector<double> initial_records(const Scene& s) {
	std::vector<double> rec(s.bodies.size() * RP_STATE_STRIDE, 0.0);
	for (size_t i = 0; i < s.bodies.size(); ++i) {
		double* r = &rec[i * RP_STATE_STRIDE];
		const BodyInit& b = s.bodies[i];
		r[0] = b.x.x; r[1] = b.x.y; r[2] = b.x.z;
		r[3] = b.q.x; r[4] = b.q.y; r[5] = b.q.z; r[6] = b.q.w;
		r[7] = b.v0.x; r[8] = b.v0.y; r[9] = b.v0.z;
		r[10] = b.w0.x; r[11] = b.w0.y; r[12] = b.w0.z;
		r[13] = 1.0;  // entity->active = true (entity.cpp:50)
	}
	return rec;
}

extern "C" {

const char* rp_last_error(void) { return g_err.c_str(); }

int rp_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

// ---------------------------------------------------------------------------------------------------------- scenes
rp_scene* rp_scene_create(void) { return new rp_scene(); }
void rp_scene_destroy(rp_scene* s) { delete s; }

int rp_scene_collider_hull(rp_scene* s, const double* v, uint32_t nv, const uint32_t* idx, uint32_t nidx) {
	if (!s || !v || !idx || nv == 0 || nidx < 3) return -1;
	for (uint32_t i = 0; i < nidx; ++i) {
		if (idx[i] >= nv) return -1;
	}
	return s->s.add_hull_collider(v, nv, idx, nidx);
}
int rp_scene_set_hull_device(rp_scene* s, int cuda_device) {
	if (!s) return RP_ERR_ARG;
	if (cuda_device < 0) {
		s->s.hull_builder = nullptr;
		s->s.hull_device = -1;
		return RP_OK;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || cuda_device >= n) {
		cudaGetLastError();
		return fail(RP_ERR_CUDA, "rp_scene_set_hull_device: no such CUDA device");
	}
	s->s.hull_builder = rp::build_hull_device;
	s->s.hull_device = cuda_device;
	return RP_OK;
}
int rp_scene_hull_build_stats(const rp_scene* s, int* hulls_built, double* milliseconds) {
	if (!s) return RP_ERR_ARG;
	if (hulls_built) *hulls_built = s->s.hulls_built;
	if (milliseconds) *milliseconds = s->s.hull_build_ms;
	return RP_OK;
}
int rp_scene_collider_sphere(rp_scene* s, float radius) {
	if (!s) return -1;
	return s->s.add_sphere_collider(radius);
}
int rp_scene_add_body(rp_scene* s, const double pos[3], const double quat[4], double mass, int fixed, double mu_s, double mu_d, double rest) {
	if (!s || !pos || !quat) return -1;
	// the reference would divide by a zero mass / invert a singular tensor silently (entity.cpp:40-47): refused here
	if (!fixed && !(mass > 0.0)) return -1;
	return s->s.add_body(pos, quat, mass, fixed, mu_s, mu_d, rest);
}

static bool csr_ok(const uint32_t* ptr, const uint32_t* idx, uint32_t rows, uint32_t limit) {
	if (!ptr || ptr[0] != 0) return false;
	for (uint32_t r = 0; r < rows; ++r) {
		if (ptr[r + 1] < ptr[r]) return false;
	}
	if (ptr[rows] && !idx) return false;
	for (uint32_t k = 0; k < ptr[rows]; ++k) {
		if (idx[k] >= limit) return false;
	}
	return true;
}
static void csr_copy(const uint32_t* ptr, const uint32_t* idx, uint32_t rows, std::vector<int>& optr, std::vector<int>& oidx) {
	optr.assign(ptr, ptr + rows + 1);
	oidx.assign(idx, idx + ptr[rows]);
}
int rp_scene_collider_hull_topology(rp_scene* s, const double* verts, uint32_t nv, const double* normals, uint32_t nf, const uint32_t* face_ptr,
	const uint32_t* face_idx, const uint32_t* v2f_ptr, const uint32_t* v2f_idx, const uint32_t* v2n_ptr, const uint32_t* v2n_idx,
	const uint32_t* f2n_ptr, const uint32_t* f2n_idx) {
	if (!s || !verts || !normals || nv == 0 || nf == 0) return -1;
	if (!csr_ok(face_ptr, face_idx, nf, nv) || !csr_ok(v2f_ptr, v2f_idx, nv, nf) || !csr_ok(v2n_ptr, v2n_idx, nv, nv) ||
		!csr_ok(f2n_ptr, f2n_idx, nf, nf)) return -1;
	HullHost h;
	h.verts.resize(nv);
	h.normals.resize(nf);
	for (uint32_t i = 0; i < nv; ++i) h.verts[i] = v3((real)verts[3 * i], (real)verts[3 * i + 1], (real)verts[3 * i + 2]);
	for (uint32_t i = 0; i < nf; ++i) h.normals[i] = v3((real)normals[3 * i], (real)normals[3 * i + 1], (real)normals[3 * i + 2]);
	csr_copy(face_ptr, face_idx, nf, h.face_ptr, h.face_idx);
	csr_copy(v2f_ptr, v2f_idx, nv, h.v2f_ptr, h.v2f_idx);
	csr_copy(v2n_ptr, v2n_idx, nv, h.v2n_ptr, h.v2n_idx);
	csr_copy(f2n_ptr, f2n_idx, nf, h.f2n_ptr, h.f2n_idx);
	return s->s.add_hull_topology(h);
}
int rp_scene_add_body_params(rp_scene* s, const double pos[3], const double quat[4], double inverse_mass, const double inertia[9],
	const double inverse_inertia[9], double radius, int fixed, double mu_s, double mu_d, double rest) {
	if (!s || !pos || !quat || !inertia || !inverse_inertia) return -1;
	return s->s.add_body_params(pos, quat, inverse_mass, inertia, inverse_inertia, radius, fixed, mu_s, mu_d, rest);
}

static bool valid_pair(const rp_scene* s, int e1, int e2) {
	if (!s) return false;
	int n = (int)s->s.bodies.size();
	return e1 >= 0 && e2 >= 0 && e1 < n && e2 < n && e1 != e2;
}
static V3 vec(const double* p) { return v3(p[0], p[1], p[2]); }
static Joint blank_joint(int type, int e1, int e2) {
	Joint j;
	memset(&j, 0, sizeof(j));
	j.type = type; j.e1 = e1; j.e2 = e2;
	return j;
}

int rp_scene_add_positional_constraint(rp_scene* s, int e1, int e2, const double r1[3], const double r2[3], double compliance, const double dist[3]) {
	if (!s || !valid_pair(s, e1, e2) || !r1 || !r2 || !dist) return -1;
	Joint j = blank_joint(JOINT_POSITIONAL, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2); j.compliance = compliance; j.distance = vec(dist);
	s->s.joints.push_back(j);
	return (int)s->s.joints.size() - 1;
}
int rp_scene_add_mutual_orientation_constraint(rp_scene* s, int e1, int e2, double compliance) {
	if (!s || !valid_pair(s, e1, e2)) return -1;
	Joint j = blank_joint(JOINT_MUTUAL_ORIENTATION, e1, e2);
	j.compliance = compliance;
	s->s.joints.push_back(j);
	return (int)s->s.joints.size() - 1;
}
int rp_scene_add_hinge_joint_constraint(rp_scene* s, int e1, int e2, const double r1[3], const double r2[3], double compliance, int a1, int a2,
	int limited, int l1, int l2, double lower, double upper) {
	if (!s || !valid_pair(s, e1, e2) || !r1 || !r2 || a1 < 0 || a1 > 5 || a2 < 0 || a2 > 5 || l1 < 0 || l1 > 5 || l2 < 0 || l2 > 5) return -1;
	Joint j = blank_joint(JOINT_HINGE, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2); j.compliance = compliance;
	j.axis[0] = a1; j.axis[1] = a2; j.axis[2] = l1; j.axis[3] = l2;
	j.limited = limited ? 1 : 0; j.lower = lower; j.upper = upper;
	s->s.joints.push_back(j);
	return (int)s->s.joints.size() - 1;
}
int rp_scene_add_spherical_joint_constraint(rp_scene* s, int e1, int e2, const double r1[3], const double r2[3], int sw1, int sw2, int tw1,
	int tw2, double swing_lower, double swing_upper, double twist_lower, double twist_upper) {
	if (!s || !valid_pair(s, e1, e2) || !r1 || !r2 || sw1 < 0 || sw1 > 5 || sw2 < 0 || sw2 > 5 || tw1 < 0 || tw1 > 5 || tw2 < 0 || tw2 > 5) return -1;
	Joint j = blank_joint(JOINT_SPHERICAL, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2);
	j.axis[0] = sw1; j.axis[1] = sw2; j.axis[2] = tw1; j.axis[3] = tw2;
	j.lower = swing_lower; j.upper = swing_upper; j.lower2 = twist_lower; j.upper2 = twist_upper;
	s->s.joints.push_back(j);
	return (int)s->s.joints.size() - 1;
}

int rp_scene_num_bodies(const rp_scene* s) { return s ? (int)s->s.bodies.size() : -1; }

int rp_scene_get_params(const rp_scene* s, double* out) {
	if (!s || !out) return RP_ERR_ARG;
	for (size_t i = 0; i < s->s.bodies.size(); ++i) {
		const BodyInit& b = s->s.bodies[i];
		double* o = out + RP_PARAM_STRIDE * i;
		o[0] = b.inv_mass;
		for (int r = 0; r < 3; ++r) {
			for (int c = 0; c < 3; ++c) {
				o[1 + 3 * r + c] = b.inertia.m[r][c];
				o[10 + 3 * r + c] = b.inv_inertia.m[r][c];
			}
		}
		o[19] = b.radius; o[20] = b.mu_s; o[21] = b.mu_d; o[22] = b.rest; o[23] = b.fixed ? 1.0 : 0.0; o[24] = (double)b.ncol;
	}
	return RP_OK;
}

static const HullHost* find_hull(const rp_scene* s, int body, int collider, bool* sphere) {
	*sphere = false;
	if (!s || body < 0 || body >= (int)s->s.bodies.size()) return 0;
	const BodyInit& b = s->s.bodies[body];
	if (collider < 0 || collider >= b.ncol) return 0;
	const ColliderDesc& c = s->s.colliders[b.col0 + collider];
	if (c.type != SHAPE_HULL) {
		*sphere = true;
		return 0;
	}
	return &s->s.hulls[c.hull];
}

int rp_scene_hull_sizes(const rp_scene* s, int body, int collider, int32_t out6[6]) {
	if (!out6) return RP_ERR_ARG;
	bool sphere;
	const HullHost* h = find_hull(s, body, collider, &sphere);
	if (sphere) {
		out6[0] = -1;
		return RP_OK;
	}
	if (!h) return RP_ERR_ARG;
	out6[0] = (int32_t)h->verts.size(); out6[1] = (int32_t)h->normals.size(); out6[2] = (int32_t)h->face_idx.size();
	out6[3] = (int32_t)h->v2f_idx.size(); out6[4] = (int32_t)h->v2n_idx.size(); out6[5] = (int32_t)h->f2n_idx.size();
	return RP_OK;
}

static void copy_u32(const std::vector<int>& v, uint32_t* out) {
	for (size_t i = 0; i < v.size(); ++i) out[i] = (uint32_t)v[i];
}

int rp_scene_hull_dump(const rp_scene* s, int body, int collider, double* verts, double* normals, uint32_t* face_ptr, uint32_t* face_idx,
	uint32_t* v2f_ptr, uint32_t* v2f_idx, uint32_t* v2n_ptr, uint32_t* v2n_idx, uint32_t* f2n_ptr, uint32_t* f2n_idx) {
	bool sphere;
	const HullHost* h = find_hull(s, body, collider, &sphere);
	if (!h || !verts || !normals || !face_ptr || !face_idx || !v2f_ptr || !v2f_idx || !v2n_ptr || !v2n_idx || !f2n_ptr || !f2n_idx) return RP_ERR_ARG;
	for (size_t i = 0; i < h->verts.size(); ++i) {
		verts[3 * i] = h->verts[i].x; verts[3 * i + 1] = h->verts[i].y; verts[3 * i + 2] = h->verts[i].z;
	}
	for (size_t i = 0; i < h->normals.size(); ++i) {
		normals[3 * i] = h->normals[i].x; normals[3 * i + 1] = h->normals[i].y; normals[3 * i + 2] = h->normals[i].z;
	}
	copy_u32(h->face_ptr, face_ptr); copy_u32(h->face_idx, face_idx);
	copy_u32(h->v2f_ptr, v2f_ptr); copy_u32(h->v2f_idx, v2f_idx);
	copy_u32(h->v2n_ptr, v2n_ptr); copy_u32(h->v2n_idx, v2n_idx);
	copy_u32(h->f2n_ptr, f2n_ptr); copy_u32(h->f2n_idx, f2n_idx);
	return RP_OK;
}

// the scene as it was described: what a caller needs to build the same scene elsewhere (the tests feed the oracle with it)
int rp_scene_initial_state(const rp_scene* s, double* out) {
	if (!s || !out) return RP_ERR_ARG;
	const std::vector<double> rec = initial_records(s->s);
	memcpy(out, rec.data(), rec.size() * sizeof(double));
	return RP_OK;
}
int rp_scene_body_desc(const rp_scene* s, int body, double out[16]) {
	if (!s || !out || body < 0 || body >= (int)s->s.bodies.size()) return RP_ERR_ARG;
	const BodyInit& b = s->s.bodies[body];
	const double v[16] = {b.x.x, b.x.y, b.x.z, b.q.x, b.q.y, b.q.z, b.q.w, b.mass, b.fixed ? 1.0 : 0.0, b.mu_s, b.mu_d, b.rest, (double)b.ncol, 0, 0, 0};
	memcpy(out, v, sizeof(v));
	return RP_OK;
}
int rp_scene_collider_soup_size(const rp_scene* s, int body, int collider, uint32_t* nverts, uint32_t* nidx, float* radius) {
	if (!s || !nverts || !nidx || !radius || body < 0 || body >= (int)s->s.bodies.size()) return RP_ERR_ARG;
	const BodyInit& b = s->s.bodies[body];
	if (collider < 0 || collider >= b.ncol) return RP_ERR_ARG;
	const ColliderDesc& c = s->s.colliders[b.col0 + collider];
	*radius = c.radius;
	if (c.type != SHAPE_HULL) {
		*nverts = *nidx = 0;
		return RP_OK;
	}
	const HullHost& h = s->s.hulls[c.hull];
	*nverts = (uint32_t)(h.key.size() / 3);
	*nidx = (uint32_t)h.key_idx.size();
	return RP_OK;
}
int rp_scene_collider_soup(const rp_scene* s, int body, int collider, double* verts, uint32_t* idx) {
	bool sphere;
	const HullHost* h = 0;
	if (s && body >= 0 && body < (int)s->s.bodies.size()) {
		const BodyInit& b = s->s.bodies[body];
		if (collider >= 0 && collider < b.ncol && s->s.colliders[b.col0 + collider].type == SHAPE_HULL) h = &s->s.hulls[s->s.colliders[b.col0 + collider].hull];
	}
	(void)sphere;
	if (!h || !verts || !idx) return RP_ERR_ARG;
	memcpy(verts, h->key.data(), h->key.size() * sizeof(double));
	memcpy(idx, h->key_idx.data(), h->key_idx.size() * sizeof(uint32_t));
	return RP_OK;
}
int rp_scene_num_joints(const rp_scene* s) { return s ? (int)s->s.joints.size() : -1; }
int rp_scene_joint_desc(const rp_scene* s, int joint, int32_t ints[8], double vals[14]) {
	if (!s || !ints || !vals || joint < 0 || joint >= (int)s->s.joints.size()) return RP_ERR_ARG;
	const Joint& j = s->s.joints[joint];
	const int32_t iv[8] = {j.type, j.e1, j.e2, j.limited, j.axis[0], j.axis[1], j.axis[2], j.axis[3]};
	const double dv[14] = {j.r1_lc.x, j.r1_lc.y, j.r1_lc.z, j.r2_lc.x, j.r2_lc.y, j.r2_lc.z, j.distance.x, j.distance.y, j.distance.z, j.compliance,
		j.lower, j.upper, j.lower2, j.upper2};
	memcpy(ints, iv, sizeof(iv));
	memcpy(vals, dv, sizeof(dv));
	return RP_OK;
}

// --------------------------------------------------------------------------------------------------------- batches
void rp_batch_cfg_default(rp_batch_cfg* cfg) {
	if (!cfg) return;
	memset(cfg, 0, sizeof(*cfg));
	cfg->linear_sleeping_threshold = 0.10;
	cfg->angular_sleeping_threshold = 0.10;
	cfg->deactivation_time = 1.0;
}

void rp_batch_destroy(rp_batch* b) {
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	if (b->graph_exec) cudaGraphExecDestroy(b->graph_exec);
	if (b->graph) cudaGraphDestroy(b->graph);
	for (size_t i = 0; i < b->allocs.size(); ++i) cudaFree(b->allocs[i]);
	if (b->ev0) cudaEventDestroy(b->ev0);
	if (b->ev1) cudaEventDestroy(b->ev1);
	if (b->overflow_pin) cudaFreeHost(b->overflow_pin);
	if (b->stream) cudaStreamDestroy(b->stream);
	cudaGetLastError();
	delete b;
}

static int create_impl(const rp_scene* scene, uint32_t n_worlds, int device, const rp_batch_cfg* cfg_in, rp_batch* b) {
	const Scene& s = scene->s;
	rp_batch_cfg cfg;
	rp_batch_cfg_default(&cfg);
	if (cfg_in) cfg = *cfg_in;
	int ndev = 0;
	RP_CUDA(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return fail(RP_ERR_CUDA, "no such CUDA device");
	RP_CUDA(cudaSetDevice(device));
	b->device = device;
	b->scene = s;
	b->scene.pending.clear();
	RP_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	RP_CUDA(cudaEventCreate(&b->ev0));
	RP_CUDA(cudaEventCreate(&b->ev1));
	cudaDeviceProp prop;
	RP_CUDA(cudaGetDeviceProperties(&prop, device));
	b->sm_count = prop.multiProcessorCount;

	if (cfg.solve_order > RP_ORDER_COLOURED) return fail(RP_ERR_ARG, "rp_batch_create: unknown solve_order");
	b->coloured = cfg.solve_order == RP_ORDER_COLOURED ? 1 : 0;
	DevView& d = b->d;
	memset(&d, 0, sizeof(d));
	d.W = (int)n_worlds;
	// world stride: a multiple of 32 so that a warp's 32 worlds start on an aligned line -- except for batches smaller than
	// a warp (one large scene), where padding would only put 31 unused slots between consecutive bodies, pairs and
	// contacts of the one world and leave the item-per-lane kernels (GJK, EPA, clipping, the sweeps) with strided accesses
	d.WS = d.W >= 32 ? (d.W + 31) / 32 * 32 : d.W;
	d.NB = (int)s.bodies.size();
	d.NC = (int)s.colliders.size();
	d.NJ = (int)s.joints.size();
	d.TV = s.total_tv;
	d.TN = s.total_tn;
	d.lin_sleep = cfg.linear_sleeping_threshold;
	d.ang_sleep = cfg.angular_sleeping_threshold;
	d.sleep_time = cfg.deactivation_time;
	d.dbg_world = -1;

	// One large scene (rp_large.cuh) or many small worlds? Bodies far larger than the typical one (the floor) stay out of the
	// uniform grid: its cell edge is set by the largest of the others.
	if (const char* e = getenv("RP_LARGE_SCENE")) cfg.large_scene = (uint32_t)atoi(e);  // tuning aid
	if (cfg.large_scene > 2) return fail(RP_ERR_ARG, "rp_batch_create: unknown large_scene");
	// (the coloured order's schedule is one sequential thread per world otherwise: worth replacing from ~1000 bodies)
	b->large = cfg.large_scene == 2 || (cfg.large_scene == 0 && d.NB >= (b->coloured ? 1024 : 4096));
	std::vector<unsigned char> is_large(d.NB, 0);
	std::vector<int> large_ids;
	double r_small = 0.0;
	if (b->large) {
		std::vector<double> radii(d.NB);
		for (int i = 0; i < d.NB; ++i) radii[i] = s.bodies[i].radius;
		std::nth_element(radii.begin(), radii.begin() + d.NB / 2, radii.end());
		const double cut = 4.0 * radii[d.NB / 2];
		for (int i = 0; i < d.NB; ++i) {
			if (s.bodies[i].radius > cut) {
				is_large[i] = 1;
				large_ids.push_back(i);
			} else {
				r_small = std::max(r_small, s.bodies[i].radius);
			}
		}
	}
	const double cell_edge = (2.0 * r_small + 0.1) * (1.0 + 1e-6);

	// capacities: derived from the broadphase of the initial poses unless given
	size_t init_pairs = 0;
	if (!b->large) {
		for (int i = 0; i < d.NB; ++i) {
			for (int j = i + 1; j < d.NB; ++j) {
				double dist = length(sub(s.bodies[i].x, s.bodies[j].x));
				if (dist <= s.bodies[i].radius + s.bodies[j].radius + 0.1) init_pairs += (size_t)s.bodies[i].ncol * s.bodies[j].ncol;
			}
		}
	} else if (!cfg.max_pairs_per_world) {
		// the same count through a host-side grid of the same cells (the quadratic loop above takes minutes for 65 k bodies)
		std::vector<std::pair<long long, int>> keyed;
		auto key_of = [&](long long cx, long long cy, long long cz) { return ((cx + (1ll << 20)) << 42) ^ ((cy + (1ll << 20)) << 21) ^ (cz + (1ll << 20)); };
		for (int i = 0; i < d.NB; ++i) {
			if (is_large[i]) continue;
			const V3 x = s.bodies[i].x;
			keyed.push_back(std::make_pair(key_of((long long)floor(x.x / cell_edge), (long long)floor(x.y / cell_edge), (long long)floor(x.z / cell_edge)), i));
		}
		std::sort(keyed.begin(), keyed.end());
		for (size_t k = 0; k < keyed.size(); ++k) {
			const int i = keyed[k].second;
			const V3 x = s.bodies[i].x;
			const long long cx = (long long)floor(x.x / cell_edge), cy = (long long)floor(x.y / cell_edge), cz = (long long)floor(x.z / cell_edge);
			for (long long dz = -1; dz <= 1; ++dz) {
				for (long long dy = -1; dy <= 1; ++dy) {
					for (long long dx = -1; dx <= 1; ++dx) {
						const long long key = key_of(cx + dx, cy + dy, cz + dz);
						auto it = std::lower_bound(keyed.begin(), keyed.end(), std::make_pair(key, -1));
						for (; it != keyed.end() && it->first == key; ++it) {
							const int j = it->second;
							if (j > i && length(sub(x, s.bodies[j].x)) <= s.bodies[i].radius + s.bodies[j].radius + 0.1) init_pairs += (size_t)s.bodies[i].ncol * s.bodies[j].ncol;
						}
					}
				}
			}
		}
		for (size_t l = 0; l < large_ids.size(); ++l) {
			const int i = large_ids[l];
			for (int j = 0; j < d.NB; ++j) {
				if (j == i || (is_large[j] && j < i)) continue;
				if (length(sub(s.bodies[i].x, s.bodies[j].x)) <= s.bodies[i].radius + s.bodies[j].radius + 0.1) init_pairs += (size_t)s.bodies[i].ncol * s.bodies[j].ncol;
			}
		}
	}
	// (a large scene's bodies are expected to pile up: room for 8 pairs per body on top of what the initial poses need)
	size_t mp = cfg.max_pairs_per_world ? cfg.max_pairs_per_world : std::max<size_t>(128, 2 * init_pairs + 64 + (b->large ? 8 * (size_t)d.NB : 0));
	mp = (mp + 127) / 128 * 128;
	d.max_pairs = (int)mp;
	b->expected_pairs = (size_t)d.W * std::max<size_t>(init_pairs, b->large ? mp / 2 : 1);
	d.max_contacts = (int)(cfg.max_contacts_per_world ? cfg.max_contacts_per_world : std::max<size_t>(256, 8 * (size_t)d.NB));
	d.max_units = d.NJ + d.max_pairs;
	b->cull = cfg.disable_cull ? 0 : 1;
	b->no_islands = cfg.disable_islands != 0;
	{
		// k_cull: grid.y = groups of 32 worlds (lane = world), a CTA's 8 warps take 8 pair indices per trip
		const int groups = (d.W + 31) / 32;
		int want = (b->sm_count * RP_CULL_CTAS_PER_SM + groups - 1) / groups;
		int most = (d.max_pairs + 7) / 8;
		b->cull_chunks = std::max(1, std::min(want, most));
	}
	{
		int coop = 0, per_sm = 0;
		RP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
		if (!coop) return fail(RP_ERR_CUDA, "device does not support cooperative launches");
		// deep, mostly empty schedules come from compound bodies (a pair of bodies with m and n colliders chains m * n units)
		int most_colliders = 0;
		for (size_t i = 0; i < s.bodies.size(); ++i) most_colliders = std::max(most_colliders, s.bodies[i].ncol);
		b->live_lists = most_colliders > 2 ? 1 : 0;
		b->live_smem = b->live_lists ? sizeof(LiveLevels) : 0;
		if (d.NJ > 0) RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_pos<true>, RP_POS_THREADS, b->live_smem));
		else RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_pos<false>, RP_POS_THREADS, b->live_smem));
		if (per_sm < 1) return fail(RP_ERR_CUDA, "k_solve_pos does not fit an SM");
		b->pos_grid = (unsigned int)(b->sm_count * per_sm);
		RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_vel, RP_VEL_THREADS, b->live_smem));
		if (per_sm < 1) return fail(RP_ERR_CUDA, "k_solve_vel does not fit an SM");
		b->vel_grid = (unsigned int)(b->sm_count * per_sm);
	}
	if (const char* e = getenv("RP_NARROW_GRID")) {  // tuning aid: "gjk,epa,manifold" CTAs per SM
		unsigned int a = 0, c = 0, m = 0;
		if (sscanf(e, "%u,%u,%u", &a, &c, &m) == 3 && a && c && m) {
			b->grid_gjk = a; b->grid_epa = c; b->grid_manifold = m;
		}
	}
	if (const char* e = getenv("RP_CARVEOUT")) {  // tuning aid: one shared-memory carve-out (percent) for every kernel of the substep
		const int pct = atoi(e);
		cudaFuncSetAttribute(k_integrate, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_cull, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_gjk, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_epa, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_manifold, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_solve_pos<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_solve_pos<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_solve_vel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_substep_reset, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaGetLastError();
	}
	// the polytope / clip-polygon stores of the narrowphase live in (dynamic) shared memory
	RP_CUDA(cudaFuncSetAttribute(k_epa, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_EPA_SMEM_BYTES));
	RP_CUDA(cudaFuncSetAttribute(k_manifold, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_MANIFOLD_SMEM_BYTES));
	RP_CUDA(cudaFuncSetAttribute(k_schedule<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_SCHED_SMEM_MAX));
	RP_CUDA(cudaFuncSetAttribute(k_schedule<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_SCHED_SMEM_MAX));

	// Dependency levels of the external constraints: they head the constraint array (pbd.cpp:580) in every world, so
	// their part of the schedule is a constant of the template.
	std::vector<int> jlast(d.NB, 0), jlevel(d.NJ, 0);
	std::vector<unsigned long long> jcolours(d.NB, 0ull);
	int jl_max = 0;
	for (int u = 0; u < d.NJ; ++u) {
		const Joint& j = s.joints[u];
		const int fa = s.bodies[j.e1].fixed, fb = s.bodies[j.e2].fixed;
		int lvl;
		if (b->coloured) {
			unsigned long long na, nb;
			lvl = SchedEntry<true>::place(fa ? 0ull : jcolours[j.e1], fb ? 0ull : jcolours[j.e2], &na, &nb);
			if (!fa) jcolours[j.e1] = na;
			if (!fb) jcolours[j.e2] = nb;
		} else {
			int na, nb;
			lvl = SchedEntry<false>::place(fa ? 0 : jlast[j.e1], fb ? 0 : jlast[j.e2], &na, &nb);
			if (!fa) jlast[j.e1] = na;
			if (!fb) jlast[j.e2] = nb;
		}
		jlevel[u] = lvl;
		jl_max = std::max(jl_max, lvl);
	}
	std::vector<int> jsched, jlptr(jl_max + 1, 0);
	for (int l = 1; l <= jl_max; ++l) {
		for (int u = 0; u < d.NJ; ++u) {
			if (jlevel[u] == l) jsched.push_back(u);
		}
		jlptr[l] = (int)jsched.size();
	}
	d.joint_levels = jl_max;
	d.max_levels = jl_max + d.max_pairs;
	b->joint_level = jlevel;

	// template: per-body records + de-duplicated physical classes
	std::vector<BodyStatic> bs(d.NB);
	std::vector<BodyClass> classes;
	b->no_restitution = true;
	for (int i = 0; i < d.NB; ++i) b->no_restitution = b->no_restitution && s.bodies[i].rest == 0.0;
	for (int i = 0; i < d.NB; ++i) {
		const BodyInit& bi = s.bodies[i];
		BodyStatic& o = bs[i];
		memset(&o, 0, sizeof(o));
		BodyClass bc;
		memset(&bc, 0, sizeof(bc));
		bc.inv_mass = bi.inv_mass;
		bc.inertia = bi.inertia; bc.inv_inertia = bi.inv_inertia;
		bc.mu_s = bi.mu_s; bc.mu_d = bi.mu_d; bc.rest = bi.rest;
		bc.ii_bound = tensor_bound(bi.inv_inertia);
		int cls = -1;
		// bitwise comparison (signed zeros and all); scenes have few classes, and the most recent one usually matches
		for (int k = (int)classes.size() - 1; k >= 0 && k >= (int)classes.size() - 64; --k) {
			if (memcmp(&classes[k], &bc, sizeof(bc)) == 0) {
				cls = k;
				break;
			}
		}
		if (cls < 0) {
			cls = (int)classes.size();
			classes.push_back(bc);
		}
		o.cls = cls;
		o.radius = bi.radius;
		o.rfar = (bi.radius + 0.1) * (1.0 + 1e-9);
		o.fixed = bi.fixed; o.col0 = bi.col0; o.ncol = bi.ncol;
		o.tv0 = bi.ncol ? s.colliders[bi.col0].tv0 : 0;
		o.tn0 = bi.ncol ? s.colliders[bi.col0].tn0 : 0;
		for (int c = bi.col0; c < bi.col0 + bi.ncol; ++c) {
			const ColliderDesc& cd = s.colliders[c];
			o.tvn += cd.type == SHAPE_HULL ? (int)s.hulls[cd.hull].verts.size() : 1;
			o.tnn += cd.type == SHAPE_HULL ? (int)s.hulls[cd.hull].normals.size() : 0;
		}
	}
	HullPoolHost hp = pool_hulls(s);
	int rc;
	if ((rc = dev_upload(b, &d.bstat, bs))) return rc;
	if ((rc = dev_upload(b, &d.bclass, classes))) return rc;
	if ((rc = dev_upload(b, &d.cols, s.colliders))) return rc;
	if ((rc = dev_upload(b, &d.joints, s.joints))) return rc;
	if ((rc = dev_upload(b, &d.joint_sched, jsched))) return rc;
	if ((rc = dev_upload(b, &d.joint_lptr, jlptr))) return rc;
	if ((rc = dev_upload(b, &d.joint_last, jlast))) return rc;
	if ((rc = dev_upload(b, &d.joint_colours, jcolours))) return rc;
	if ((rc = dev_upload(b, &d.pool.hulls, hp.hulls))) return rc;
	if ((rc = dev_upload(b, &d.pool.verts, hp.verts))) return rc;
	if ((rc = dev_upload(b, &d.pool.normals, hp.normals))) return rc;
	if ((rc = dev_upload(b, &d.pool.face_ptr, hp.face_ptr))) return rc;
	if ((rc = dev_upload(b, &d.pool.face_idx, hp.face_idx))) return rc;
	if ((rc = dev_upload(b, &d.pool.v2f_ptr, hp.v2f_ptr))) return rc;
	if ((rc = dev_upload(b, &d.pool.v2f_idx, hp.v2f_idx))) return rc;
	if ((rc = dev_upload(b, &d.pool.v2n_ptr, hp.v2n_ptr))) return rc;
	if ((rc = dev_upload(b, &d.pool.v2n_idx, hp.v2n_idx))) return rc;
	if ((rc = dev_upload(b, &d.pool.f2n_ptr, hp.f2n_ptr))) return rc;
	if ((rc = dev_upload(b, &d.pool.f2n_idx, hp.f2n_idx))) return rc;
	if ((rc = dev_alloc(b, &b->force_dev, (size_t)d.NB))) return rc;
	if ((rc = dev_alloc(b, &b->torque_dev, (size_t)d.NB))) return rc;
	d.force = b->force_dev;
	d.torque = b->torque_dev;

	// per world (world-minor arrays are padded to WS worlds)
	const size_t W = (size_t)d.W, WS = (size_t)d.WS, WB = W * d.NB, SB = WS * d.NB, WP = W * d.max_pairs, SP = WS * d.max_pairs;
	if ((rc = dev_alloc(b, &d.dyn, SB * RP_DYN_DOUBLES))) return rc;
	if ((rc = dev_alloc(b, &d.active, SB))) return rc;
	if ((rc = dev_alloc(b, &d.vstamp, SB))) return rc;
	if ((rc = dev_alloc(b, &d.epoch, 1))) return rc;
	if ((rc = dev_alloc(b, &d.deact, SB))) return rc;
	// transformed geometry is not stored: the narrowphase evaluates vertices and face normals from the poses (PoseShape)
	if ((rc = dev_alloc(b, &d.tv, 1))) return rc;
#if defined(RP_STORED_NORMALS)
	if ((rc = dev_alloc(b, &d.tn, WS * std::max(d.TN, 1) * 3))) return rc;
#else
	if ((rc = dev_alloc(b, &d.tn, 1))) return rc;
#endif
	if ((rc = dev_alloc(b, &d.pairs, SP))) return rc;
	if ((rc = dev_alloc(b, &d.n_pairs, W))) return rc;
	if (b->large) {
		GridView& g = b->grid;
		memset(&g, 0, sizeof(g));
		g.table = 1024;
		while (g.table < 2 * d.NB) g.table *= 2;
		g.inv_cell = 1.0 / cell_edge;
		g.n_large = (int)large_ids.size();
		const int most = std::max(std::max(g.table + 1, d.NB + 1), 1);
		const size_t tiles = ((size_t)most + RP_SCAN_TILE - 1) / RP_SCAN_TILE;
		if ((rc = dev_upload(b, &g.large, large_ids))) return rc;
		if ((rc = dev_upload(b, &g.is_large, is_large))) return rc;
		if ((rc = dev_alloc(b, &g.bucket, WB))) return rc;
		if ((rc = dev_alloc(b, &g.start, W * ((size_t)g.table + 1)))) return rc;
		if ((rc = dev_alloc(b, &g.cursor, W * (size_t)g.table))) return rc;
		if ((rc = dev_alloc(b, &g.sorted, WB))) return rc;
		if ((rc = dev_alloc(b, &g.row, W * ((size_t)d.NB + 1)))) return rc;
		if ((rc = dev_alloc(b, &g.sums, W * tiles))) return rc;
		if ((rc = dev_alloc(b, &g.totals, W))) return rc;
		if (b->coloured) {
			ColourView& c = b->col;
			memset(&c, 0, sizeof(c));
			if ((rc = dev_alloc(b, &c.deg, W * ((size_t)d.NB + 1)))) return rc;
			if ((rc = dev_alloc(b, &c.fill, WB))) return rc;
			if ((rc = dev_alloc(b, &c.adj, 2 * WP))) return rc;
			if ((rc = dev_alloc(b, &c.colour, WP))) return rc;
			if ((rc = dev_alloc(b, &c.pending, WP))) return rc;
			if ((rc = dev_alloc(b, &c.remaining, RP_COLOUR_ROUNDS + 1))) return rc;
			c.sums = g.sums;
			c.totals = g.totals;
		}
		d.n_cells = 0;
		if ((rc = dev_alloc(b, &d.cells, 1))) return rc;
		if ((rc = dev_alloc(b, &d.cell_mask, 1))) return rc;
		if ((rc = dev_alloc(b, &d.cell_off, 1))) return rc;
	} else {
		std::vector<int2> cells;
		for (int i = 0; i + 1 < d.NB; ++i) {
			for (int j0 = i + 1; j0 < d.NB; j0 += 32) cells.push_back(make_int2(i, j0));
		}
		d.n_cells = (int)cells.size();
		if ((rc = dev_upload(b, &d.cells, cells))) return rc;
		if ((rc = dev_alloc(b, &d.cell_mask, WS * std::max(d.n_cells, 1)))) return rc;
		if ((rc = dev_alloc(b, &d.cell_off, WS * std::max(d.n_cells, 1)))) return rc;
	}
	if ((rc = dev_alloc(b, &d.label, WB))) return rc;
	if ((rc = dev_alloc(b, &d.isl_flag, WB))) return rc;
	if ((rc = dev_alloc(b, &d.last_level, SB))) return rc;
	if ((rc = dev_alloc(b, &d.colour_tab, b->coloured ? SB : 1))) return rc;
	if ((rc = dev_alloc(b, &d.pair_level, SP))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_hist, WS * (d.max_levels + 2)))) return rc;
	if ((rc = dev_alloc(b, &d.aabb, WS * std::max(d.NC, 1) * 6))) return rc;
	if ((rc = dev_alloc(b, &d.geom_stamp, WS * std::max(d.NC, 1)))) return rc;
	if ((rc = dev_alloc(b, &d.cands, WP, false))) return rc;
	if ((rc = dev_alloc(b, &d.cand_count, 1))) return rc;
	if ((rc = dev_alloc(b, &d.big_count, 1))) return rc;
	d.cand_cap = (unsigned int)WP;
	{
		// does any pair of colliders exceed the per-thread staging block of k_gjk? (two largest vertex counts)
		int n1 = 0, n2 = 0;
		for (size_t c = 0; c < s.colliders.size(); ++c) {
			const int nv = s.colliders[c].nv;
			if (nv > n1) { n2 = n1; n1 = nv; } else if (nv > n2) n2 = nv;
		}
		b->has_big_pairs = warp_pair_verts(n1 + n2);
		// heavy geometry: bounds per collider instead of per body, and several threads per collider in k_transform
		int body_most = 0, collider_most = 0;
		for (size_t i = 0; i < s.bodies.size(); ++i) {
			int nv = 0;
			for (int c = s.bodies[i].col0; c < s.bodies[i].col0 + s.bodies[i].ncol; ++c) {
				const ColliderDesc& cd = s.colliders[c];
				nv += cd.nv;
				if (cd.type == SHAPE_HULL) collider_most = std::max(collider_most, cd.nv + (int)s.hulls[cd.hull].normals.size());
			}
			body_most = std::max(body_most, nv);
		}
		d.split_bounds = body_most > 64 ? 1 : 0;
		b->transform_slices = (unsigned int)std::min(16, std::max(1, collider_most / 64));
		d.split_big = b->has_big_pairs ? 1 : 0;
		if ((rc = dev_alloc(b, &d.big_sup, b->has_big_pairs ? WP : 1, false))) return rc;
	}
	if ((rc = dev_alloc(b, &d.simplex, WP * 12, false))) return rc;
	if ((rc = dev_alloc(b, &d.hits, WP, false))) return rc;
	if ((rc = dev_alloc(b, &d.hit_count, 1))) return rc;
	if ((rc = dev_alloc(b, &d.epa_out, WP, false))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_cap, (size_t)d.max_levels + 2))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_off, (size_t)d.max_levels + 2))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_fill, ((size_t)d.max_levels + 2) * RP_LVL_STRIDE))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_max, 2))) return rc;
	if ((rc = dev_alloc(b, &d.lvl_items, WP, false))) return rc;
	if ((rc = dev_alloc(b, &d.pair_normal, SP, false))) return rc;
	if ((rc = dev_alloc(b, &d.pair_coff, SP))) return rc;
	if ((rc = dev_alloc(b, &d.pair_ccnt, SP))) return rc;
	{
		// World-block sweeps (k_solve_block) on request (rp_batch_cfg.sweep_block_worlds), if the per-level cursors fit shared
		// memory. Measured on the north-star batch (4096 x W256, frames 40..59, ms per 400 substeps): level-major cooperative
		// sweeps 180; world blocks of 2 / 4 / 8 / 16 worlds 355 / 227 / 219 / 193 -- a block only has its own worlds' units of
		// a level to fill its lanes with, where the level-major lists pack every world's, so the default stays level-major.
		int wpb = 0;
		if (cfg.sweep_block_worlds) wpb = (int)cfg.sweep_block_worlds;
		if (const char* e = getenv("RP_SWEEP_WPB")) wpb = atoi(e);  // tuning aid
		const size_t smem = ((size_t)d.max_levels + 2) * sizeof(int);
		if (wpb > RP_SB_MAX_WPB) wpb = RP_SB_MAX_WPB;
		if (wpb > 0 && smem <= 96 * 1024) {
			b->sweep_wpb = wpb;
			b->sweep_smem = smem;
			d.block_mode = 1;
			RP_CUDA(cudaFuncSetAttribute(k_solve_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			RP_CUDA(cudaFuncSetAttribute(k_solve_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		}
		if ((rc = dev_alloc(b, &d.live, d.block_mode ? SP : 1, false))) return rc;
		if ((rc = dev_alloc(b, &d.n_live, W))) return rc;
		if ((rc = dev_alloc(b, &d.blk_items, d.block_mode ? WP : 1, false))) return rc;
	}
	{
		// Dataflow sweeps (pos_flow / vel_flow: a unit waits for the previous live unit of each of its two bodies, no grid barriers)
		// for contact-only scenes: batches of at least two warps of worlds, and one large / coloured scene. Measured (B200, ms per
		// frame, barrier form -> dataflow): w256 x 4096 frames 40..59 24.1 -> 20.5, stack x 4096 2.56 -> 2.24, brick wall 32 x 32
		// coloured 5.39 -> 4.23, the 65,600-body pile 54.3 -> 39.2 (its sweeps 4.2x: a 64-contact manifold no longer holds a whole
		// level up). Scenes with external constraints can run it (their joints are links of the chains) but keep the barrier
		// form by default: the levers x 16384 have two or three levels per sweep and nothing to gain (1.25 -> 1.31).
		int flow = d.NJ == 0 && (d.W >= 64 || b->large || b->coloured) && !d.block_mode ? 1 : 0;
		if (cfg.sweep_form == 1) flow = 0;
		if (cfg.sweep_form == 2 && !d.block_mode) flow = 1;
		if (const char* e = getenv("RP_FLOW")) {  // tuning aid, overrides rp_batch_cfg.sweep_form: 0 = grid barriers between levels, 2 = dataflow for any batch
			const int v = atoi(e);
			if (v == 0) flow = 0;
			if (v == 2 && !d.block_mode) flow = 1;
		}
		d.flow_mode = flow;
		if ((rc = dev_alloc(b, &d.body_live, flow ? SB : 1))) return rc;
		if ((rc = dev_alloc(b, &d.body_done, flow ? 2 * SB : 1))) return rc;
		if ((rc = dev_alloc(b, &d.flow_cursor, 4))) return rc;
		// levels of the joints on each body (template constant): they are links of the body's chain in every substep
		std::vector<unsigned long long> jmask((size_t)d.NB, 0ull);
		for (int u = 0; u < d.NJ; ++u) {
			if (b->joint_level[u] < 1 || b->joint_level[u] > RP_FLOW_LEVELS) continue;  // (deeper schedules keep the barrier form)
			jmask[s.joints[u].e1] |= 1ull << b->joint_level[u];
			jmask[s.joints[u].e2] |= 1ull << b->joint_level[u];
		}
		if ((rc = dev_upload(b, &d.joint_body_mask, jmask))) return rc;
	}
	if ((rc = dev_alloc(b, &d.contacts, WS * d.max_contacts * 8, false))) return rc;
	if ((rc = dev_alloc(b, &d.n_contacts, W))) return rc;
	if ((rc = dev_alloc(b, &d.lambdas, WS * std::max(d.NJ, 1)))) return rc;
	if ((rc = dev_alloc(b, &d.status, W))) return rc;
	if ((rc = dev_alloc(b, &d.counters, 8))) return rc;
	if ((rc = dev_alloc(b, &d.dbg_points, 2 * (size_t)d.max_contacts, false))) return rc;
	if ((rc = dev_alloc(b, &b->rec_dev, WB * RP_STATE_STRIDE, false))) return rc;
	if ((rc = dev_alloc(b, &b->overflow_dev, 1))) return rc;
	RP_CUDA(cudaMallocHost((void**)&b->overflow_pin, sizeof(int)));
	*b->overflow_pin = 0;

	std::vector<double> rec = initial_records(s);
	RP_CUDA(cudaMemcpyAsync(b->rec_dev, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
	k_unpack_state<<<dim3(d.NB, (d.W + 127) / 128), 128, 0, b->stream>>>(d, b->rec_dev, 0, d.W, 1);
	RP_CUDA(cudaGetLastError());
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}

int rp_batch_create(const rp_scene* scene, uint32_t n_worlds, int device, const rp_batch_cfg* cfg, rp_batch** out) {
	if (!scene || !out || n_worlds == 0 || n_worlds > 65535 || scene->s.bodies.empty()) return fail(RP_ERR_ARG, "rp_batch_create: bad argument");
	if (!scene->s.pending.empty()) return fail(RP_ERR_ARG, "rp_batch_create: colliders queued without a body");
	rp_batch* b = new rp_batch();
	int rc = create_impl(scene, n_worlds, device, cfg, b);
	if (rc) {
		std::string keep = g_err;
		rp_batch_destroy(b);
		g_err = keep;
		return rc;
	}
	*out = b;
	return RP_OK;
}

int rp_batch_create_from(const rp_scene* scene, rp_batch* src, const int32_t* new_from_old, const rp_batch_cfg* cfg, rp_batch** out) {
	if (!scene || !src || !new_from_old || !out || scene->s.bodies.empty()) return fail(RP_ERR_ARG, "rp_batch_create_from: bad argument");
	if (!scene->s.pending.empty()) return fail(RP_ERR_ARG, "rp_batch_create_from: colliders queued without a body");
	const int nb = (int)scene->s.bodies.size();
	std::vector<int> map(new_from_old, new_from_old + nb);
	for (int i = 0; i < nb; ++i) {
		if (map[i] >= src->d.NB) return fail(RP_ERR_ARG, "rp_batch_create_from: mapping points past the source's bodies");
	}
	rp_batch* b = new rp_batch();
	int rc = create_impl(scene, (uint32_t)src->d.W, src->device, cfg, b);
	if (!rc) {
		// state moves device to device: the source's stream first finishes what it was given
		const int* map_dev = 0;
		if (cudaStreamSynchronize(src->stream) != cudaSuccess) rc = fail(RP_ERR_CUDA, "rp_batch_create_from: source stream");
		if (!rc) rc = dev_upload(b, &map_dev, map);
		if (!rc) {
			k_adopt_bodies<<<flat_grid(b, (size_t)nb, 128), 128, 0, b->stream>>>(b->d, src->d, map_dev);
			if (cudaStreamSynchronize(b->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = fail(RP_ERR_CUDA, "rp_batch_create_from: copy");
		}
	}
	if (rc) {
		std::string keep = g_err;
		rp_batch_destroy(b);
		g_err = keep;
		return rc;
	}
	*out = b;
	return RP_OK;
}

uint32_t rp_batch_num_worlds(const rp_batch* b) { return b ? (uint32_t)b->d.W : 0; }
uint32_t rp_batch_num_bodies(const rp_batch* b) { return b ? (uint32_t)b->d.NB : 0; }

int rp_batch_clear_forces(rp_batch* b) {
	if (!b) return RP_ERR_ARG;
	b->scene.clear_forces();
	b->forces_dirty = true;
	return RP_OK;
}
int rp_batch_add_force(rp_batch* b, int body, const double position[3], const double force[3]) {
	if (!b || body < 0 || body >= b->d.NB || !position || !force) return RP_ERR_ARG;
	b->scene.add_force(body, vec(position), vec(force));
	b->forces_dirty = true;
	return RP_OK;
}
int rp_batch_add_gravity(rp_batch* b, double g) {
	if (!b) return RP_ERR_ARG;
	b->scene.add_gravity(g);
	b->forces_dirty = true;
	return RP_OK;
}

static int flush_forces(rp_batch* b) {
	if (!b->forces_dirty) return RP_OK;
	// the copies are stream-ordered; the host vectors stay alive in b->scene and pageable copies are staged by the runtime
	RP_CUDA(cudaMemcpyAsync(b->force_dev, b->scene.force.data(), sizeof(V3) * b->d.NB, cudaMemcpyHostToDevice, b->stream));
	RP_CUDA(cudaMemcpyAsync(b->torque_dev, b->scene.torque.data(), sizeof(V3) * b->d.NB, cudaMemcpyHostToDevice, b->stream));
	b->forces_dirty = false;
	return RP_OK;
}

// exclusive scan of data[world][0..n) in place (k_scan_*), grand totals to totals[world] (or nowhere)
static void launch_scan(rp_batch* b, int* data, int n, size_t world_stride, int* sums, int* totals) {
	const int tiles = (n + RP_SCAN_TILE - 1) / RP_SCAN_TILE;
	k_scan_tiles<<<dim3(tiles, b->d.W), RP_SCAN_THREADS, 0, b->stream>>>(data, n, world_stride, sums, tiles);
	k_scan_sums<<<b->d.W, RP_SCAN_THREADS, 0, b->stream>>>(sums, tiles, totals);
	k_scan_add<<<dim3(tiles, b->d.W), RP_SCAN_THREADS, 0, b->stream>>>(data, n, world_stride, sums, tiles);
}
// coloured order of a large scene: per-body adjacency of the units, then rounds of independent sets (rp_large.cuh)
static void launch_colouring(rp_batch* b, int collisions) {
	const DevView& d = b->d;
	const ColourView& c = b->col;
	const size_t W = (size_t)d.W;
	cudaMemsetAsync(c.deg, 0, W * (d.NB + 1) * sizeof(int), b->stream);
	cudaMemsetAsync(c.fill, 0, W * d.NB * sizeof(int), b->stream);
	cudaMemsetAsync(c.pending, 0, W * d.max_pairs * sizeof(int), b->stream);
	cudaMemsetAsync(c.remaining, 0, (RP_COLOUR_ROUNDS + 1) * sizeof(int), b->stream);
	const dim3 per_pair = flat_grid(b, (size_t)d.max_pairs, 256);
	k_col_degree<<<per_pair, 256, 0, b->stream>>>(d, c, collisions);
	launch_scan(b, c.deg, d.NB + 1, (size_t)d.NB + 1, c.sums, 0);
	k_col_fill<<<per_pair, 256, 0, b->stream>>>(d, c);
	for (int r = 0; r < RP_COLOUR_ROUNDS; ++r) {
		k_col_round<<<per_pair, 256, 0, b->stream>>>(d, c, r);
		k_col_commit<<<per_pair, 256, 0, b->stream>>>(d, c, r);
	}
	k_col_leftover<<<per_pair, 256, 0, b->stream>>>(d, c, RP_COLOUR_ROUNDS);
	k_col_levels<<<1, 32, 0, b->stream>>>(d);
}
static void launch_islands(rp_batch* b, double dt) {
	const DevView& d = b->d;
	if (b->no_islands) return;  // pbd.cpp:476-533 compiled out: every body stays active, no deactivation timers
	if (!b->large) {
		k_islands<<<d.W, 256, 0, b->stream>>>(d, dt);
		return;
	}
	const dim3 per_body = flat_grid(b, (size_t)d.NB, 256);
	k_uf_init<<<per_body, 256, 0, b->stream>>>(d);
	k_uf_hook<<<flat_grid(b, (size_t)d.max_pairs + d.NJ, 256), 256, 0, b->stream>>>(d);
	k_uf_sleep<<<per_body, 256, 0, b->stream>>>(d, dt);
	k_uf_apply<<<per_body, 256, 0, b->stream>>>(d);
}

static void launch_schedule(rp_batch* b, int collisions) {
	const DevView& d = b->d;
	if (b->large && b->coloured) {
		launch_colouring(b, collisions);
		return;
	}
	const size_t entry = b->coloured ? sizeof(unsigned long long) : sizeof(int);
	const size_t lanes = d.W < 32 ? d.W : 32;  // columns of the shared tables (k_schedule's SL)
	const size_t smem = ((size_t)d.NB * lanes * entry + RP_SCHED_HIST * lanes * sizeof(int) + (size_t)d.NB * lanes + 15) / 16 * 16;
	const unsigned int grid = (unsigned int)((d.W + 31) / 32);
	if (b->coloured) {
		if (smem <= RP_SCHED_SMEM_MAX) k_schedule<true, true><<<grid, 32, smem, b->stream>>>(d, collisions);
		else k_schedule<false, true><<<grid, 32, 0, b->stream>>>(d, collisions);
	} else {
		if (smem <= RP_SCHED_SMEM_MAX) k_schedule<true, false><<<grid, 32, smem, b->stream>>>(d, collisions);
		else k_schedule<false, false><<<grid, 32, 0, b->stream>>>(d, collisions);
	}
}
static void launch_broad_grid(rp_batch* b) {
	const DevView& d = b->d;
	const GridView& g = b->grid;
	const size_t W = (size_t)d.W;
	cudaMemsetAsync(g.start, 0, W * (g.table + 1) * sizeof(int), b->stream);
	cudaMemsetAsync(g.cursor, 0, W * g.table * sizeof(int), b->stream);
	cudaMemsetAsync(g.row, 0, W * (d.NB + 1) * sizeof(int), b->stream);
	const dim3 per_body = flat_grid(b, (size_t)d.NB, 256);
	k_grid_count<<<per_body, 256, 0, b->stream>>>(d, g);
	launch_scan(b, g.start, g.table + 1, (size_t)g.table + 1, g.sums, 0);
	k_grid_fill<<<per_body, 256, 0, b->stream>>>(d, g);
	k_grid_rowcount<<<flat_grid(b, (size_t)d.NB, 128), 128, 0, b->stream>>>(d, g);
	if (g.n_large) k_grid_large_count<<<dim3(g.n_large, d.W), RP_SCAN_THREADS, 0, b->stream>>>(d, g);
	launch_scan(b, g.row, d.NB + 1, (size_t)d.NB + 1, g.sums, g.totals);
	k_grid_finish<<<(d.W + 63) / 64, 64, 0, b->stream>>>(d, g);
	k_grid_rowwrite<<<flat_grid(b, (size_t)d.NB, 128), 128, 0, b->stream>>>(d, g);
	if (g.n_large) k_grid_large_write<<<dim3(g.n_large, d.W), RP_SCAN_THREADS, 0, b->stream>>>(d, g);
}
static void launch_broad(rp_batch* b) {
	const DevView& d = b->d;
	if (b->large) {
		launch_broad_grid(b);
		return;
	}
	if (d.n_cells > 0) {
		const dim3 grid = flat_grid(b, (size_t)d.n_cells, 256);
		k_broad_cells<<<grid, 256, 0, b->stream>>>(d);
		k_broad_scan<<<(d.W + 31) / 32, dim3(32, RP_BROAD_SEGS), 0, b->stream>>>(d);
		k_broad_write<<<grid, 256, 0, b->stream>>>(d);
	}
}

// the per-frame prologue: broadphase, islands + sleeping, dependency-level schedule (pbd.cpp:474-533)
static void enqueue_prologue(rp_batch* b, double dt, int collisions) {
	const DevView& d = b->d;
	launch_broad(b);
	launch_islands(b, dt);
	k_level_reset<<<1, 256, 0, b->stream>>>(d);
	launch_schedule(b, collisions);
	k_level_offsets<<<1, 1, 0, b->stream>>>(d);
}

// The Gauss-Seidel sweeps are cooperative grids (grid-wide barrier between levels) of refill loops over warp-owned chunks
// (WarpQueue): launch exactly the CTAs that are resident at once (occupancy measured at batch creation).
static void launch_solve_pos(rp_batch* b, double h, uint32_t iters, int collisions) {
	if (iters == 0) return;
	if (b->d.NJ > 0) launch_cooperative(k_solve_pos<true>, b->pos_grid, (unsigned int)RP_POS_THREADS, b->live_smem, b->stream, b->d, (real)h, (int)iters, collisions, b->live_lists);
	else launch_cooperative(k_solve_pos<false>, b->pos_grid, (unsigned int)RP_POS_THREADS, b->live_smem, b->stream, b->d, (real)h, (int)iters, collisions, b->live_lists);
}
static void launch_solve_vel(rp_batch* b, double h, uint32_t iters) {
	launch_cooperative(k_solve_vel, b->vel_grid, (unsigned int)RP_VEL_THREADS, b->live_smem, b->stream, b->d, (real)h, b->live_lists, (int)iters);
}

static void enqueue_integrate(rp_batch* b, double h, bool last_substep) {
	const DevView& d = b->d;
	k_substep_reset<<<(unsigned int)((std::max(d.W, d.max_levels + 2) + 255) / 256), 256, 0, b->stream>>>(d);
	const int store_velocities = last_substep || !b->no_restitution ? 1 : 0;
	k_integrate<<<flat_grid(b, (size_t)d.NB, RP_INT_THREADS), RP_INT_THREADS, 0, b->stream>>>(d, h, store_velocities);
}
// grid of the per-hit kernel: the hit count lives on the device, so the launch covers the candidate capacity in
// grid-stride trips of at most this many CTAs
// at most the resident CTAs (one wave), at least one CTA per SM, in between what the pairs of the initial poses would fill:
// worlds of a few bodies (config 5) pay for every CTA of a mostly empty launch (592 CTAs of k_manifold finding no hit: 9 us)
static unsigned int narrow_grid(const rp_batch* b, unsigned int per_sm, unsigned int threads) {
	const size_t want = (b->expected_pairs + threads - 1) / threads;
	const size_t most = (size_t)b->sm_count * per_sm, least = (size_t)b->sm_count;
	return (unsigned int)std::min(most, std::max(least, want));
}
// the warp-per-pair kernels: 4 pairs per CTA at a time, 186 registers -> 2 CTAs resident per SM; two waves
static unsigned int warp_grid(const rp_batch* b) { return narrow_grid(b, 4, 4); }
static unsigned int manifold_grid(const rp_batch* b) { return narrow_grid(b, b->grid_manifold, RP_MANIFOLD_THREADS); }
static unsigned int epa_grid(const rp_batch* b) { return narrow_grid(b, b->grid_epa, RP_EPA_THREADS); }
static unsigned int gjk_grid(const rp_batch* b) { return narrow_grid(b, b->grid_gjk, RP_GJK_THREADS); }
static void launch_cull(rp_batch* b) {
	const DevView& d = b->d;
	if (d.NC == 0) return;  // bodies without colliders (joint-only scenes): no pairs, no candidates, nothing to transform
	if (d.split_bounds) k_bounds<<<flat_grid(b, (size_t)d.NC, RP_INT_THREADS), RP_INT_THREADS, 0, b->stream>>>(d);
	// lane = world (warp = one pair index of 32 worlds) for batches of at least a warp of worlds, thread = (pair, world) below
	if (d.W >= 32) k_cull<<<dim3(b->cull_chunks, (d.W + 31) / 32), 256, 0, b->stream>>>(d, b->cull);
	else k_cull_flat<<<flat_grid(b, (size_t)d.max_pairs, 256), 256, 0, b->stream>>>(d, b->cull);
#if defined(RP_STORED_NORMALS)
	k_transform<<<dim3(d.NC, (unsigned int)((d.W + RP_INT_THREADS - 1) / RP_INT_THREADS), b->transform_slices), RP_INT_THREADS, 0, b->stream>>>(d);
#endif
}
static void launch_gjk(rp_batch* b) {
	k_gjk<<<gjk_grid(b), RP_GJK_THREADS, 0, b->stream>>>(b->d);
	if (b->has_big_pairs) k_gjk_warp<<<warp_grid(b), RP_GJK_WARP_THREADS, 0, b->stream>>>(b->d);
}
static void launch_manifold(rp_batch* b) {
	k_epa<<<epa_grid(b), RP_EPA_THREADS, RP_EPA_SMEM_BYTES, b->stream>>>(b->d);
	if (b->has_big_pairs) k_epa_warp<<<warp_grid(b), RP_GJK_WARP_THREADS, 0, b->stream>>>(b->d);
	k_manifold<<<manifold_grid(b), RP_MANIFOLD_THREADS, RP_MANIFOLD_SMEM_BYTES, b->stream>>>(b->d);
	if (b->has_big_pairs) k_manifold_warp<<<warp_grid(b), RP_CLIPW_THREADS, 0, b->stream>>>(b->d);
}
static void enqueue_narrow(rp_batch* b) {
	launch_cull(b);
	launch_gjk(b);
	launch_manifold(b);
}
static void enqueue_solve(rp_batch* b, double h, uint32_t iters, int collisions) {
	if (b->sweep_wpb > 0) {
		// both sweeps of the substep in one launch, CTA = a block of worlds (k_solve_block)
		const unsigned int grid = (unsigned int)((b->d.W + b->sweep_wpb - 1) / b->sweep_wpb);
		if (b->d.NJ > 0) k_solve_block<true><<<grid, RP_SB_THREADS, b->sweep_smem, b->stream>>>(b->d, h, (int)iters, collisions, b->sweep_wpb);
		else k_solve_block<false><<<grid, RP_SB_THREADS, b->sweep_smem, b->stream>>>(b->d, h, (int)iters, collisions, b->sweep_wpb);
		return;
	}
	launch_solve_pos(b, h, iters, collisions);
	// velocity derivation (pbd.cpp:623-643) is lazy: a body's velocities are derived by the first velocity-level unit
	// that touches it, else by the next substep's k_integrate, else by k_derive at the end of the frame
	if (collisions) launch_solve_vel(b, h, iters);
}
static void enqueue_frame_end(rp_batch* b, double h) {
	const DevView& d = b->d;
	k_derive<<<flat_grid(b, (size_t)d.NB, 128), 128, 0, b->stream>>>(d, h);
	k_count_frame<<<1, 1, 0, b->stream>>>(d);
}

// one whole frame: prologue + substeps + end of frame. The sweep depth of the frame stays on the device (the sweep kernels
// read it), so nothing here depends on the state and the sequence is captured once per (dt, substeps, iters, collisions).
static void enqueue_frame(rp_batch* b, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	const double h = dt / substeps;  // pbd.cpp:472
	enqueue_prologue(b, dt, collisions);
	for (uint32_t s = 0; s < substeps; ++s) {
		enqueue_integrate(b, h, s + 1 == substeps);
		if (collisions) enqueue_narrow(b);
		enqueue_solve(b, h, iters, collisions);
	}
	enqueue_frame_end(b, h);
}

int rp_batch_step(rp_batch* b, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	if (!b || substeps == 0) return fail(RP_ERR_ARG, "rp_batch_step: bad argument");
	if (dt <= 0.0) return RP_OK;  // pbd.cpp:471
	RP_CUDA(cudaSetDevice(b->device));
	int rc = flush_forces(b);
	if (rc) return rc;
	GraphKey key;
	key.dt = dt; key.substeps = substeps; key.iters = iters; key.collisions = collisions ? 1 : 0;
	if (!b->have_graph || !(b->graph_key == key)) {
		if (b->graph_exec) cudaGraphExecDestroy(b->graph_exec);
		if (b->graph) cudaGraphDestroy(b->graph);
		b->graph_exec = 0;
		b->graph = 0;
		b->have_graph = false;
		RP_CUDA(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
		enqueue_frame(b, dt, substeps, iters, key.collisions);
		cudaError_t e = cudaStreamEndCapture(b->stream, &b->graph);
		if (e != cudaSuccess) return fail(RP_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
		RP_CUDA(cudaGraphInstantiate(&b->graph_exec, b->graph, 0));
		size_t nodes = 0;
		RP_CUDA(cudaGraphGetNodes(b->graph, 0, &nodes));
		b->graph_kernels = (int)nodes;
		b->graph_key = key;
		b->have_graph = true;
	}
	RP_CUDA(cudaGraphLaunch(b->graph_exec, b->stream));
	return RP_OK;
}

// A fixed-capacity device buffer that ran out (broadphase pairs, contacts, the full-size polytope / clip stores) drops work:
// the worlds concerned carry the bit in their status word, and every call that synchronises with the device reports it --
// rp_batch_sync, rp_batch_step_host, rp_batch_download_state return RP_ERR_CAPACITY (after doing their work: the state is
// downloaded, the stream is idle) until rp_batch_clear_status. Costs one tiny kernel and a 4-byte copy per synchronising call.
static int enqueue_overflow_check(rp_batch* b) {
	RP_CUDA(cudaMemsetAsync(b->overflow_dev, 0, sizeof(int), b->stream));
	k_status_overflow<<<(unsigned int)((b->d.W + 255) / 256), 256, 0, b->stream>>>(b->d, b->overflow_dev);
	RP_CUDA(cudaGetLastError());
	RP_CUDA(cudaMemcpyAsync(b->overflow_pin, b->overflow_dev, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
	return RP_OK;
}
static int overflow_result(rp_batch* b) {
	const int bits = *b->overflow_pin;
	if (!bits) return RP_OK;
	char msg[160];
	snprintf(msg, sizeof(msg), "capacity exhausted in some world (status bits 0x%x): raise rp_batch_cfg.max_pairs_per_world / max_contacts_per_world", bits);
	return fail(RP_ERR_CAPACITY, msg);
}

int rp_batch_sync(rp_batch* b) {
	if (!b) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	int rc = enqueue_overflow_check(b);
	if (rc) return rc;
	RP_CUDA(cudaStreamSynchronize(b->stream));
	RP_CUDA(cudaGetLastError());
	return overflow_result(b);
}
int rp_batch_graph_kernels(const rp_batch* b) { return b ? b->graph_kernels : -1; }

int rp_batch_run(rp_batch* b, uint32_t frames, double dt, uint32_t substeps, uint32_t iters, int collisions, float* ms_out) {
	if (!b) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	RP_CUDA(cudaEventRecord(b->ev0, b->stream));
	for (uint32_t f = 0; f < frames; ++f) {
		int rc = rp_batch_step(b, dt, substeps, iters, collisions);
		if (rc) return rc;
	}
	RP_CUDA(cudaEventRecord(b->ev1, b->stream));
	RP_CUDA(cudaEventSynchronize(b->ev1));
	float ms = 0.f;
	RP_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
	if (ms_out) *ms_out = ms;
	return RP_OK;
}

// ----------------------------------------------------------------------------------------------------------- state
static int upload_impl(rp_batch* b, uint32_t first, uint32_t n, const double* host, int broadcast) {
	if (!b || !host || n == 0 || first >= (uint32_t)b->d.W || n > (uint32_t)b->d.W - first) return fail(RP_ERR_ARG, "state transfer: bad range");
	RP_CUDA(cudaSetDevice(b->device));
	const size_t nrec = (size_t)(broadcast ? 1 : n) * b->d.NB;
	RP_CUDA(cudaMemcpyAsync(b->rec_dev, host, nrec * RP_STATE_STRIDE * sizeof(double), cudaMemcpyHostToDevice, b->stream));
	k_unpack_state<<<dim3(b->d.NB, (n + 127) / 128), 128, 0, b->stream>>>(b->d, b->rec_dev, (int)first, (int)n, broadcast);
	RP_CUDA(cudaGetLastError());
	return RP_OK;
}

int rp_batch_upload_state(rp_batch* b, uint32_t first, uint32_t n, const double* host) {
	int rc = upload_impl(b, first, n, host, 0);
	if (rc) return rc;
	RP_CUDA(cudaStreamSynchronize(b->stream));  // the caller may reuse `host` right away
	return RP_OK;
}
int rp_batch_broadcast_state(rp_batch* b, const double* host_one_world) {
	if (!b) return RP_ERR_ARG;
	int rc = upload_impl(b, 0, (uint32_t)b->d.W, host_one_world, 1);
	if (rc) return rc;
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}
static int download_async(rp_batch* b, uint32_t first, uint32_t n, double* host) {
	if (!b || !host || n == 0 || first >= (uint32_t)b->d.W || n > (uint32_t)b->d.W - first) return fail(RP_ERR_ARG, "state transfer: bad range");
	RP_CUDA(cudaSetDevice(b->device));
	const size_t total = (size_t)n * b->d.NB;
	k_pack_state<<<dim3(b->d.NB, (n + 127) / 128), 128, 0, b->stream>>>(b->d, b->rec_dev, (int)first, (int)n);
	RP_CUDA(cudaGetLastError());
	RP_CUDA(cudaMemcpyAsync(host, b->rec_dev, total * RP_STATE_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
	return RP_OK;
}
int rp_batch_download_state(rp_batch* b, uint32_t first, uint32_t n, double* host) {
	int rc = download_async(b, first, n, host);
	if (rc) return rc;
	if ((rc = enqueue_overflow_check(b))) return rc;
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return overflow_result(b);
}

int rp_batch_step_host(rp_batch* b, const double* in, double* out, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	if (!b) return RP_ERR_ARG;
	int rc;
	if (in && (rc = upload_impl(b, 0, (uint32_t)b->d.W, in, 0))) return rc;
	if ((rc = rp_batch_step(b, dt, substeps, iters, collisions))) return rc;
	if (out && (rc = download_async(b, 0, (uint32_t)b->d.W, out))) return rc;
	if ((rc = enqueue_overflow_check(b))) return rc;
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return overflow_result(b);
}

int rp_batch_get_status(rp_batch* b, int32_t* out) {
	if (!b || !out) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	RP_CUDA(cudaMemcpyAsync(out, b->d.status, sizeof(int) * b->d.W, cudaMemcpyDeviceToHost, b->stream));
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}
int rp_batch_clear_status(rp_batch* b) {
	if (!b) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	RP_CUDA(cudaMemsetAsync(b->d.status, 0, sizeof(int) * b->d.W, b->stream));
	*b->overflow_pin = 0;
	return RP_OK;
}
int rp_batch_get_counters(rp_batch* b, uint64_t out8[8]) {
	if (!b || !out8) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	RP_CUDA(cudaMemcpyAsync(out8, b->d.counters, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, b->stream));
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}

}  // extern "C"

// --------------------------------------------------------------------------------------------- parity instrumentation
template <class T>
static int fetch(rp_batch* b, std::vector<T>& host, const T* dev, size_t n) {
	host.resize(n);
	if (n) RP_CUDA(cudaMemcpyAsync(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, b->stream));
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}

// one world's column of a world-minor array: element e at dev[e * stride + w]
template <class T>
static int fetch_world(rp_batch* b, std::vector<T>& host, const T* dev, size_t n, int w) {
	host.resize(n);
	if (n) {
		RP_CUDA(cudaMemcpy2DAsync(host.data(), sizeof(T), dev + w, (size_t)b->d.WS * sizeof(T), sizeof(T), n, cudaMemcpyDeviceToHost,
			b->stream));
	}
	RP_CUDA(cudaStreamSynchronize(b->stream));
	return RP_OK;
}

extern "C" {

int rp_batch_broad_pairs(rp_batch* b, uint32_t world, uint32_t* pairs_out, uint32_t max_pairs, uint32_t* n_out) {
	if (!b || world >= (uint32_t)b->d.W || !n_out) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	const DevView& d = b->d;
	launch_broad(b);
	RP_CUDA(cudaGetLastError());
	std::vector<int> np;
	int rc = fetch(b, np, d.n_pairs + world, 1);
	if (rc) return rc;
	std::vector<PairRec> pr;
	if ((rc = fetch_world(b, pr, d.pairs, (size_t)np[0], (int)world))) return rc;
	uint32_t n = 0;
	for (int i = 0; i < np[0]; ++i) {
		if (i > 0 && pr[i].a == pr[i - 1].a && pr[i].b == pr[i - 1].b) continue;  // collider pairs of one body pair
		if (n < max_pairs && pairs_out) {
			pairs_out[2 * n] = (uint32_t)pr[i].a;
			pairs_out[2 * n + 1] = (uint32_t)pr[i].b;
		}
		++n;
	}
	*n_out = n;
	return RP_OK;
}

// the schedule of one world for its current poses: every broadphase collider pair with the level (reference order) or colour
// (coloured order) the sweeps would run it at (0 = skipped: both sides fixed or asleep)
int rp_batch_pair_levels(rp_batch* b, uint32_t world, uint32_t* pairs_out, int32_t* levels_out, uint32_t max_pairs, uint32_t* n_out) {
	if (!b || world >= (uint32_t)b->d.W || !n_out) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(b->device));
	const DevView& d = b->d;
	launch_broad(b);
	launch_islands(b, 0.0);  // (dt = 0: the deactivation timers do not advance)
	k_level_reset<<<1, 256, 0, b->stream>>>(d);
	launch_schedule(b, 1);
	RP_CUDA(cudaGetLastError());
	std::vector<int> np, lv;
	std::vector<PairRec> pr;
	int rc = fetch(b, np, d.n_pairs + world, 1);
	if (rc) return rc;
	if ((rc = fetch_world(b, pr, d.pairs, (size_t)np[0], (int)world))) return rc;
	if ((rc = fetch_world(b, lv, d.pair_level, (size_t)np[0], (int)world))) return rc;
	for (int i = 0; i < np[0] && (uint32_t)i < max_pairs; ++i) {
		if (pairs_out) {
			pairs_out[2 * i] = (uint32_t)pr[i].a;
			pairs_out[2 * i + 1] = (uint32_t)pr[i].b;
		}
		if (levels_out) levels_out[i] = lv[i];
	}
	*n_out = (uint32_t)np[0];
	return RP_OK;
}

int rp_batch_step_logged(rp_batch* b, double dt, uint32_t substeps, uint32_t iters, int collisions, uint32_t world, uint32_t* calls_out,
	uint32_t max_calls, double* contacts_out, uint32_t max_contacts, uint32_t* n_calls, uint32_t* n_contacts) {
	if (!b || substeps == 0 || world >= (uint32_t)b->d.W || !n_calls || !n_contacts) return fail(RP_ERR_ARG, "rp_batch_step_logged: bad argument");
	*n_calls = 0;
	*n_contacts = 0;
	if (dt <= 0.0) return RP_OK;
	RP_CUDA(cudaSetDevice(b->device));
	int rc = flush_forces(b);
	if (rc) return rc;
	DevView& d = b->d;
	struct DbgGuard {  // whatever way this function is left, the batch's view goes back to "no world is logged"
		DevView& d;
		~DbgGuard() { d.dbg_world = -1; }
	} guard{d};
	d.dbg_world = (int)world;
	const double h = dt / substeps;
	enqueue_prologue(b, dt, collisions ? 1 : 0);
	RP_CUDA(cudaGetLastError());
	std::vector<int> np, active, ccnt, coff;
	std::vector<PairRec> pr;
	std::vector<BodyStatic> bs;
	std::vector<V3> normals, pts;
	if ((rc = fetch(b, np, d.n_pairs + world, 1))) return rc;
	if ((rc = fetch_world(b, pr, d.pairs, (size_t)np[0], (int)world))) return rc;
	if ((rc = fetch_world(b, active, d.active, (size_t)d.NB, (int)world))) return rc;
	if ((rc = fetch(b, bs, d.bstat, (size_t)d.NB))) return rc;
	uint32_t nc = 0, nk = 0;
	for (uint32_t s = 0; s < substeps; ++s) {
		enqueue_integrate(b, h, s + 1 == substeps);
		if (collisions) {
			enqueue_narrow(b);
			RP_CUDA(cudaGetLastError());
			if ((rc = fetch_world(b, ccnt, d.pair_ccnt, (size_t)np[0], (int)world))) return rc;
			if ((rc = fetch_world(b, coff, d.pair_coff, (size_t)np[0], (int)world))) return rc;
			if ((rc = fetch_world(b, normals, d.pair_normal, (size_t)np[0], (int)world))) return rc;
			if ((rc = fetch(b, pts, d.dbg_points, 2 * (size_t)d.max_contacts))) return rc;
			for (int i = 0; i < np[0]; ++i) {
				const PairRec& p = pr[i];
				if ((bs[p.a].fixed || !active[p.a]) && (bs[p.b].fixed || !active[p.b])) continue;
				bool first = !(i > 0 && pr[i - 1].a == p.a && pr[i - 1].b == p.b);
				if (first) {
					if (nc < max_calls && calls_out) {
						calls_out[4 * nc] = (uint32_t)p.a; calls_out[4 * nc + 1] = (uint32_t)p.b;
						calls_out[4 * nc + 2] = 0; calls_out[4 * nc + 3] = nk;
					}
					++nc;
				}
				for (int k = 0; k < ccnt[i]; ++k) {
					if (nk < max_contacts && contacts_out) {
						double* o = contacts_out + 9 * (size_t)nk;
						const V3& p1 = pts[2 * (coff[i] + k)];
						const V3& p2 = pts[2 * (coff[i] + k) + 1];
						o[0] = p1.x; o[1] = p1.y; o[2] = p1.z; o[3] = p2.x; o[4] = p2.y; o[5] = p2.z;
						o[6] = normals[i].x; o[7] = normals[i].y; o[8] = normals[i].z;
					}
					++nk;
				}
				if (nc - 1 < max_calls && calls_out) calls_out[4 * (nc - 1) + 2] += (uint32_t)ccnt[i];
			}
		}
		enqueue_solve(b, h, iters, collisions ? 1 : 0);
	}
	enqueue_frame_end(b, h);
	RP_CUDA(cudaGetLastError());
	RP_CUDA(cudaStreamSynchronize(b->stream));
	d.dbg_world = -1;
	*n_calls = nc;
	*n_contacts = nk;
	return RP_OK;
}

// Per-kernel device time of `frames` frames, launched WITHOUT the graph with a CUDA event at every kernel boundary:
// ms_out[RP_K_BROAD .. RP_K_SOLVE] accumulate the time of each kernel family. Profiling aid for bench.py's roofline
// block; the headline numbers come from rp_batch_run.
int rp_batch_profile(rp_batch* b, uint32_t frames, double dt, uint32_t substeps, uint32_t iters, int collisions, float* ms_out) {
	if (!b || !ms_out || substeps == 0 || dt <= 0.0) return fail(RP_ERR_ARG, "rp_batch_profile: bad argument");
	RP_CUDA(cudaSetDevice(b->device));
	int rc = flush_forces(b);
	if (rc) return rc;
	for (int k = 0; k < RP_NUM_KERNEL_FAMILIES; ++k) ms_out[k] = 0.f;
	const DevView& d = b->d;
	const double h = dt / substeps;
	std::vector<cudaEvent_t> ev;
	std::vector<int> fam;
	struct EventGuard {
		std::vector<cudaEvent_t>& ev;
		~EventGuard() {
			for (size_t i = 0; i < ev.size(); ++i) cudaEventDestroy(ev[i]);
		}
	} guard{ev};
	auto mark = [&](int family) -> int {
		cudaEvent_t e;
		RP_CUDA(cudaEventCreate(&e));
		RP_CUDA(cudaEventRecord(e, b->stream));
		ev.push_back(e);
		fam.push_back(family);
		return RP_OK;
	};
	for (uint32_t f = 0; f < frames; ++f) {
		if ((rc = mark(-1))) return rc;
		launch_broad(b);
		if ((rc = mark(RP_K_BROAD))) return rc;
		launch_islands(b, dt);
		if ((rc = mark(RP_K_ISLANDS))) return rc;
		k_level_reset<<<1, 256, 0, b->stream>>>(d);
		launch_schedule(b, collisions ? 1 : 0);
		k_level_offsets<<<1, 1, 0, b->stream>>>(d);
		if ((rc = mark(RP_K_SCHEDULE))) return rc;
		RP_CUDA(cudaStreamSynchronize(b->stream));
		if ((rc = mark(-1))) return rc;
		for (uint32_t s = 0; s < substeps; ++s) {
			enqueue_integrate(b, h, s + 1 == substeps);
			if ((rc = mark(RP_K_INTEGRATE))) return rc;
			if (collisions) {
				launch_cull(b);
				if ((rc = mark(RP_K_CULL))) return rc;
				launch_gjk(b);
				if ((rc = mark(RP_K_GJK))) return rc;
				k_epa<<<epa_grid(b), RP_EPA_THREADS, RP_EPA_SMEM_BYTES, b->stream>>>(d);
				if (b->has_big_pairs) k_epa_warp<<<warp_grid(b), RP_GJK_WARP_THREADS, 0, b->stream>>>(d);
				if ((rc = mark(RP_K_EPA))) return rc;
				k_manifold<<<manifold_grid(b), RP_MANIFOLD_THREADS, RP_MANIFOLD_SMEM_BYTES, b->stream>>>(d);
				if (b->has_big_pairs) k_manifold_warp<<<warp_grid(b), RP_CLIPW_THREADS, 0, b->stream>>>(d);
				if ((rc = mark(RP_K_MANIFOLD))) return rc;
			}
			if (b->sweep_wpb > 0) {
				enqueue_solve(b, h, iters, collisions ? 1 : 0);  // both sweeps in one launch: reported as solve_pos
				if ((rc = mark(RP_K_SOLVE_POS))) return rc;
			} else {
				launch_solve_pos(b, h, iters, collisions ? 1 : 0);
				if ((rc = mark(RP_K_SOLVE_POS))) return rc;
				if (collisions) {
					launch_solve_vel(b, h, iters);
					if ((rc = mark(RP_K_SOLVE_VEL))) return rc;
				}
			}
		}
		enqueue_frame_end(b, h);
		if ((rc = mark(RP_K_DERIVE))) return rc;
	}
	RP_CUDA(cudaGetLastError());
	RP_CUDA(cudaStreamSynchronize(b->stream));
	for (size_t i = 1; i < ev.size(); ++i) {
		if (fam[i] < 0) continue;
		float ms = 0.f;
		RP_CUDA(cudaEventElapsedTime(&ms, ev[i - 1], ev[i]));
		ms_out[fam[i]] += ms;
	}
	return RP_OK;
}

// Measured FP64 rate of the CUDA-core pipe on `device`, TFLOP/s (mul and add counted as one flop each, an FMA as two):
// out2[0] = independent DMUL+DADD chains (the parity build's instruction mix), out2[1] = DFMA chains.
int rp_measure_fp64_peak(int device, double out2[2]) {
	if (!out2) return RP_ERR_ARG;
	RP_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	RP_CUDA(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
	double* buf = 0;
	cudaEvent_t e0 = 0, e1 = 0;
	struct ProbeGuard {
		double*& buf;
		cudaEvent_t &e0, &e1;
		~ProbeGuard() {
			if (e0) cudaEventDestroy(e0);
			if (e1) cudaEventDestroy(e1);
			if (buf) cudaFree(buf);
		}
	} guard{buf, e0, e1};
	RP_CUDA(cudaMalloc(&buf, sizeof(double) * blocks * threads));
	RP_CUDA(cudaEventCreate(&e0));
	RP_CUDA(cudaEventCreate(&e1));
	for (int mode = 0; mode < 2; ++mode) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; ++rep) {
			RP_CUDA(cudaEventRecord(e0, 0));
			if (mode == 0) k_fp64_probe<false><<<blocks, threads>>>(buf, iters, 1.0000001, 1e-9);
			else k_fp64_probe<true><<<blocks, threads>>>(buf, iters, 1.0000001, 1e-9);
			RP_CUDA(cudaEventRecord(e1, 0));
			RP_CUDA(cudaEventSynchronize(e1));
			float ms = 0.f;
			RP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
			if (rep > 0 && ms < best) best = ms;
		}
		const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
		out2[mode] = flops / (best * 1e-3) / 1e12;
	}
	return RP_OK;
}

}  // extern "C"
