// rp_scene.h -- host-side scene template: bodies, colliders, convex-hull topology pool, external constraints.
//
// One Scene is the template every world of a batch is instantiated from (static parameters and hull topology are
// shared by all worlds; only poses, velocities and sleep state are per world). Building it is setup-time work, but
// parity-defining: hull vertex/face/adjacency ORDER decides support-point ties and clipping order
// (SURVEY.md 8 a20).
//
// Reference: src/physics/collider.cpp:194-364 (collider_convex_hull_create), :448-521 (default inertia tensor,
// bounding radius), src/entity.cpp:24-65 (entity_create_ex).
#ifndef RP_SCENE_H
#define RP_SCENE_H

#include <string>
#include <vector>
#include "rp_solve.h"

namespace rp {

struct HullHost {  // one hull in the reference's index order, CSR
	std::vector<V3> verts, normals;
	std::vector<int> face_ptr, face_idx, v2f_ptr, v2f_idx, v2n_ptr, v2n_idx, f2n_ptr, f2n_idx;
	std::vector<double> key;  // the input soup it was built from (deduplication of identical colliders)
	std::vector<uint32_t> key_idx;
};

struct BodyInit {
	V3 x; Q4 q;
	double inv_mass;
	M3 inertia, inv_inertia;
	double radius;          // bounding_sphere_radius
	double mu_s, mu_d, rest;
	int fixed;
	int col0, ncol;         // collider range
	double mass;            // as given to add_body (0 when the body was adopted through add_body_params)
	V3 v0, w0;              // initial velocities (zero unless an example's `perturb` variant sets them)
};

struct Scene {
	std::vector<HullHost> hulls;
	std::vector<ColliderDesc> colliders;      // tv0/tn0 filled by finalize()
	std::vector<BodyInit> bodies;
	std::vector<Joint> joints;
	std::vector<ColliderDesc> pending;        // colliders of the body being assembled
	int total_tv = 0, total_tn = 0;           // transformed vertices / normals per world
	std::vector<V3> force, torque;            // per-body external force / torque sums for the coming frame(s)
	// hull topology builder: build_hull() on the host, or (set through rp_scene_set_hull_device) the device build of rp_hull.cuh,
	// reached through a pointer because this file is also compiled into the CPU checker, which has no CUDA
	bool (*hull_builder)(int device, const double* verts_xyz, uint32_t nverts, const uint32_t* indices, uint32_t nidx, HullHost* out,
		float* ms_out, std::string* err) = nullptr;
	int hull_device = -1;
	double hull_build_ms = 0.0;               // time spent building hull topology (host: wall clock; device: CUDA events)
	int hulls_built = 0;

	int add_hull_collider(const double* verts_xyz, uint32_t nverts, const uint32_t* indices, uint32_t nidx);
	int add_sphere_collider(float radius);
	// a hull whose topology already exists (Collider_Convex_Hull as the reference holds it, collider.h:19-29): adopted as is
	int add_hull_topology(const HullHost& h);
	// an entity whose derived parameters already exist (Entity fields, entity.h:27-40): adopted as they are
	int add_body_params(const double* pos, const double* quat_xyzw, double inv_mass, const double* inertia9, const double* inv_inertia9,
		double radius, int fixed, double mu_s, double mu_d, double rest);
	int add_body(const double* pos, const double* quat_xyzw, double mass, int fixed, double mu_s, double mu_d, double rest);
	void clear_forces();
	void add_force(int body, V3 position, V3 f);   // entity_add_force with local_coords = false (entity.cpp:176-193)
	void add_gravity(double g);                    // the examples' idiom (stack.cpp:93-96)
};

// pooled, pointer-free copy of the hull data ready for one cudaMemcpy per array (or direct host use)
struct HullPoolHost {
	std::vector<HullTopo> hulls;
	std::vector<V3> verts, normals;
	std::vector<int> face_ptr, face_idx, v2f_ptr, v2f_idx, v2n_ptr, v2n_idx, f2n_ptr, f2n_idx;
	HullPool view() const;
};
HullPoolHost pool_hulls(const Scene& s);

HullHost build_hull(const double* verts_xyz, uint32_t nverts, const uint32_t* indices, uint32_t nidx);

}  // namespace rp

struct rp_scene {  // the opaque handle of include/rawphys_b200.h
	rp::Scene s;
};
#endif
