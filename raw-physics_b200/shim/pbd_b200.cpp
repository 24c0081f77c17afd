// pbd_b200.cpp -- same-signature replacement for the reference's two entry points (src/physics/pbd.h:93-94)
//
//     void pbd_simulate(r64 dt, Entity** entities, u32 num_substeps, u32 num_pos_iters, boolean enable_collisions);
//     void pbd_simulate_with_constraints(r64 dt, Entity** entities, Constraint* external_constraints,
//                                        u32 num_substeps, u32 num_pos_iters, boolean enable_collisions);
//
// on top of the C ABI of librawphys_b200.so (include/rawphys_b200.h). It is compiled INSIDE a raw-physics tree: it
// includes the reference's own headers (entity.h, physics/pbd.h, light_array.h) and reads the reference's own structs,
// so `src/examples/*.cpp` link against it unchanged. One call = gather the Entity table into a scene template + state
// records, one rp_batch_step_host on a single world, scatter the state back into the entities.
//
// Argument meaning and side effects follow pbd.cpp:464-747: dt <= 0 is a no-op (:471); forces are read from
// entity->forces (the caller adds them before and clears them after the call); external constraints are not written
// (their lambdas restart at 0 every substep in a private copy, :580); entities' pose, velocities, previous velocities,
// active flag and deactivation time are updated in place. Error behaviour: the reference asserts; so does this (a
// message on stderr, then abort) -- for the reference's own assert sites, which the library reports as status bits, and
// for CUDA failures (there is no CPU fallback to continue on).
//
// Two ways to link it:
//   * as the definition of the two functions (drop pbd.cpp's own two definitions), or
//   * -DRP_SHIM_WRAP together with `-Wl,--wrap=_Z12pbd_simulatedPP6Entityjji
//     -Wl,--wrap=_Z29pbd_simulate_with_constraintsdPP6EntityP10Constraintjji`, leaving every reference source untouched
//     (what the test harness does for tests/test_gpu_shim.py, where the reference sources must stay as they are).
#include <light_array.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "physics/pbd.h"
#include "rawphys_b200.h"

namespace {

// Everything of the entity table and the constraint list that is constant between calls in an ordinary run. The library
// keeps one scene template + one single-world batch alive for as long as this image does not change.
struct Image {
	std::vector<unsigned char> bytes;
	template <class T>
	void put(const T& v) {
		const unsigned char* p = (const unsigned char*)&v;
		bytes.insert(bytes.end(), p, p + sizeof(T));
	}
	void put_array(const void* p, size_t n) {
		if (n) bytes.insert(bytes.end(), (const unsigned char*)p, (const unsigned char*)p + n);
	}
};

rp_scene* g_scene = 0;
rp_batch* g_batch = 0;
Image g_image;
uint32_t g_max_pairs = 0, g_max_contacts = 0;

void die(const char* what) {
	fprintf(stderr, "pbd_b200: %s: %s\n", what, rp_last_error());
	abort();
}

int device_index() {
	const char* e = getenv("RAWPHYS_B200_DEVICE");
	return e ? atoi(e) : 0;
}

int index_of(Entity** es, eid id) {
	for (u32 i = 0; i < array_length(es); ++i) {
		if (es[i]->id == id) return (int)i;
	}
	fprintf(stderr, "pbd_b200: a constraint names entity %llu, which is not in the entity list\n", (unsigned long long)id);
	abort();
}

void append_csr(u32** lists, u32 rows, std::vector<uint32_t>& ptr, std::vector<uint32_t>& idx) {
	ptr.assign(1, 0u);
	idx.clear();
	for (u32 r = 0; r < rows; ++r) {
		for (u32 k = 0; k < array_length(lists[r]); ++k) idx.push_back(lists[r][k]);
		ptr.push_back((uint32_t)idx.size());
	}
}

// flat copies of one Collider_Convex_Hull (collider.h:19-29), in the reference's own order
struct HullArrays {
	std::vector<double> verts, normals;
	std::vector<uint32_t> face_ptr, face_idx, v2f_ptr, v2f_idx, v2n_ptr, v2n_idx, f2n_ptr, f2n_idx;
	uint32_t nv, nf;
};

void flatten(const Collider_Convex_Hull* h, HullArrays& o) {
	o.nv = (uint32_t)array_length(h->vertices);
	o.nf = (uint32_t)array_length(h->faces);
	o.verts.clear();
	o.normals.clear();
	for (u32 i = 0; i < o.nv; ++i) {
		o.verts.push_back(h->vertices[i].x); o.verts.push_back(h->vertices[i].y); o.verts.push_back(h->vertices[i].z);
	}
	o.face_ptr.assign(1, 0u);
	o.face_idx.clear();
	for (u32 f = 0; f < o.nf; ++f) {
		o.normals.push_back(h->faces[f].normal.x); o.normals.push_back(h->faces[f].normal.y); o.normals.push_back(h->faces[f].normal.z);
		for (u32 k = 0; k < array_length(h->faces[f].elements); ++k) o.face_idx.push_back(h->faces[f].elements[k]);
		o.face_ptr.push_back((uint32_t)o.face_idx.size());
	}
	append_csr(h->vertex_to_faces, o.nv, o.v2f_ptr, o.v2f_idx);
	append_csr(h->vertex_to_neighbors, o.nv, o.v2n_ptr, o.v2n_idx);
	append_csr(h->face_to_neighbors, o.nf, o.f2n_ptr, o.f2n_idx);
}

template <class T>
void put_vec(Image& im, const std::vector<T>& v) {
	im.put((uint64_t)v.size());
	im.put_array(v.data(), v.size() * sizeof(T));
}

// Walks the entity table once: serialises its constant part into `im` and, when `scene` is given, builds the template.
void describe(Entity** es, Constraint* cs, Image& im, rp_scene* scene) {
	const u32 n = (u32)array_length(es);
	im.put(n);
	HullArrays ha;
	for (u32 i = 0; i < n; ++i) {
		Entity* e = es[i];
		const u32 nc = e->colliders ? (u32)array_length(e->colliders) : 0;
		im.put(nc);
		for (u32 c = 0; c < nc; ++c) {
			Collider* col = &e->colliders[c];
			im.put((int)col->type);
			if (col->type == COLLIDER_TYPE_SPHERE) {
				im.put(col->sphere.radius);
				if (scene && rp_scene_collider_sphere(scene, col->sphere.radius) < 0) die("rp_scene_collider_sphere");
			} else {
				flatten(&col->convex_hull, ha);
				put_vec(im, ha.verts); put_vec(im, ha.normals); put_vec(im, ha.face_ptr); put_vec(im, ha.face_idx);
				put_vec(im, ha.v2f_ptr); put_vec(im, ha.v2f_idx); put_vec(im, ha.v2n_ptr); put_vec(im, ha.v2n_idx);
				put_vec(im, ha.f2n_ptr); put_vec(im, ha.f2n_idx);
				if (scene && rp_scene_collider_hull_topology(scene, ha.verts.data(), ha.nv, ha.normals.data(), ha.nf, ha.face_ptr.data(),
					ha.face_idx.data(), ha.v2f_ptr.data(), ha.v2f_idx.data(), ha.v2n_ptr.data(), ha.v2n_idx.data(), ha.f2n_ptr.data(),
					ha.f2n_idx.data()) < 0) die("rp_scene_collider_hull_topology");
			}
		}
		double inertia[9], inv_inertia[9];
		for (int r = 0; r < 3; ++r) {
			for (int c = 0; c < 3; ++c) {
				inertia[3 * r + c] = e->inertia_tensor.data[r][c];
				inv_inertia[3 * r + c] = e->inverse_inertia_tensor.data[r][c];
			}
		}
		const int fixed = e->fixed ? 1 : 0;
		im.put(e->inverse_mass); im.put(inertia); im.put(inv_inertia); im.put(e->bounding_sphere_radius); im.put(fixed);
		im.put(e->static_friction_coefficient); im.put(e->dynamic_friction_coefficient); im.put(e->restitution_coefficient);
		if (scene) {
			const double p[3] = {e->world_position.x, e->world_position.y, e->world_position.z};
			const double q[4] = {e->world_rotation.x, e->world_rotation.y, e->world_rotation.z, e->world_rotation.w};
			if (rp_scene_add_body_params(scene, p, q, e->inverse_mass, inertia, inv_inertia, e->bounding_sphere_radius, fixed,
				e->static_friction_coefficient, e->dynamic_friction_coefficient, e->restitution_coefficient) < 0) die("rp_scene_add_body_params");
		}
	}
	const u32 nk = cs ? (u32)array_length(cs) : 0;
	im.put(nk);
	for (u32 k = 0; k < nk; ++k) {
		const Constraint* c = &cs[k];
		const int e1 = index_of(es, c->e1_id), e2 = index_of(es, c->e2_id);
		im.put((int)c->type); im.put(e1); im.put(e2);
		int rc = 0;
		switch (c->type) {
			case POSITIONAL_CONSTRAINT: {
				const Positional_Constraint* p = &c->positional_constraint;
				const double r1[3] = {p->r1_lc.x, p->r1_lc.y, p->r1_lc.z}, r2[3] = {p->r2_lc.x, p->r2_lc.y, p->r2_lc.z};
				const double dist[3] = {p->distance.x, p->distance.y, p->distance.z};
				im.put(r1); im.put(r2); im.put(p->compliance); im.put(dist);
				if (scene) rc = rp_scene_add_positional_constraint(scene, e1, e2, r1, r2, p->compliance, dist);
			} break;
			case MUTUAL_ORIENTATION_CONSTRAINT: {
				im.put(c->mutual_orientation_constraint.compliance);
				if (scene) rc = rp_scene_add_mutual_orientation_constraint(scene, e1, e2, c->mutual_orientation_constraint.compliance);
			} break;
			case HINGE_JOINT_CONSTRAINT: {
				const Hinge_Joint_Constraint* h = &c->hinge_joint_constraint;
				const double r1[3] = {h->r1_lc.x, h->r1_lc.y, h->r1_lc.z}, r2[3] = {h->r2_lc.x, h->r2_lc.y, h->r2_lc.z};
				const int limited = h->limited ? 1 : 0;
				// the limit fields are only meaningful (and only initialised, pbd.cpp:35-63) for a limited hinge
				const int l1 = limited ? (int)h->e1_limit_axis : 0, l2 = limited ? (int)h->e2_limit_axis : 0;
				const double lo = limited ? h->lower_limit : 0.0, hi = limited ? h->upper_limit : 0.0;
				im.put(r1); im.put(r2); im.put(h->compliance); im.put((int)h->e1_aligned_axis); im.put((int)h->e2_aligned_axis);
				im.put(limited); im.put(l1); im.put(l2); im.put(lo); im.put(hi);
				if (scene) {
					rc = rp_scene_add_hinge_joint_constraint(scene, e1, e2, r1, r2, h->compliance, (int)h->e1_aligned_axis, (int)h->e2_aligned_axis,
						limited, l1, l2, lo, hi);
				}
			} break;
			case SPHERICAL_JOINT_CONSTRAINT: {
				const Spherical_Joint_Constraint* s = &c->spherical_joint_constraint;
				const double r1[3] = {s->r1_lc.x, s->r1_lc.y, s->r1_lc.z}, r2[3] = {s->r2_lc.x, s->r2_lc.y, s->r2_lc.z};
				im.put(r1); im.put(r2); im.put((int)s->e1_swing_axis); im.put((int)s->e2_swing_axis); im.put((int)s->e1_twist_axis);
				im.put((int)s->e2_twist_axis); im.put(s->swing_lower_limit); im.put(s->swing_upper_limit); im.put(s->twist_lower_limit);
				im.put(s->twist_upper_limit);
				if (scene) {
					rc = rp_scene_add_spherical_joint_constraint(scene, e1, e2, r1, r2, (int)s->e1_swing_axis, (int)s->e2_swing_axis,
						(int)s->e1_twist_axis, (int)s->e2_twist_axis, s->swing_lower_limit, s->swing_upper_limit, s->twist_lower_limit,
						s->twist_upper_limit);
				}
			} break;
			default:
				// the reference's solve_constraint would walk a user-supplied COLLISION_CONSTRAINT too, but no caller builds
				// one: collision constraints are pbd's own per-substep output (pbd.cpp:584-611)
				fprintf(stderr, "pbd_b200: external constraint %u has unsupported type %d\n", k, (int)c->type);
				abort();
		}
		if (rc < 0) die("rp_scene_add_*_constraint");
	}
}

void drop_batch() {
	if (g_batch) rp_batch_destroy(g_batch);
	g_batch = 0;
}

void make_batch() {
	drop_batch();
	rp_batch_cfg cfg;
	rp_batch_cfg_default(&cfg);
	cfg.max_pairs_per_world = g_max_pairs;
	cfg.max_contacts_per_world = g_max_contacts;
	if (rp_batch_create(g_scene, 1, device_index(), &cfg, &g_batch) != RP_OK) die("rp_batch_create");
}

void rebuild(Entity** es, Constraint* cs) {
	drop_batch();
	if (g_scene) rp_scene_destroy(g_scene);
	g_scene = rp_scene_create();
	Image unused;
	describe(es, cs, unused, g_scene);
	g_max_pairs = g_max_contacts = 0;  // derived from the poses at hand
	make_batch();
}

void step(r64 dt, Entity** es, Constraint* cs, u32 substeps, u32 iters, boolean collisions) {
	if (dt <= 0.0) return;  // pbd.cpp:471
	Image now;
	describe(es, cs, now, 0);
	if (!g_batch || now.bytes != g_image.bytes) {
		rebuild(es, cs);
		g_image.bytes.swap(now.bytes);
	}
	const u32 n = (u32)array_length(es);
	std::vector<double> in((size_t)n * RP_STATE_STRIDE), out((size_t)n * RP_STATE_STRIDE);
	for (u32 i = 0; i < n; ++i) {
		const Entity* e = es[i];
		double* r = &in[(size_t)i * RP_STATE_STRIDE];
		r[0] = e->world_position.x; r[1] = e->world_position.y; r[2] = e->world_position.z;
		r[3] = e->world_rotation.x; r[4] = e->world_rotation.y; r[5] = e->world_rotation.z; r[6] = e->world_rotation.w;
		r[7] = e->linear_velocity.x; r[8] = e->linear_velocity.y; r[9] = e->linear_velocity.z;
		r[10] = e->angular_velocity.x; r[11] = e->angular_velocity.y; r[12] = e->angular_velocity.z;
		r[13] = e->active ? 1.0 : 0.0;
		r[14] = e->deactivation_time;
		r[15] = e->previous_linear_velocity.x; r[16] = e->previous_linear_velocity.y; r[17] = e->previous_linear_velocity.z;
		r[18] = e->previous_angular_velocity.x; r[19] = e->previous_angular_velocity.y; r[20] = e->previous_angular_velocity.z;
	}
	for (int attempt = 0;; ++attempt) {
		if (rp_batch_clear_forces(g_batch) != RP_OK) die("rp_batch_clear_forces");
		for (u32 i = 0; i < n; ++i) {  // calculate_external_force / _torque (physics_util.cpp:5-23) sum these in list order
			const Entity* e = es[i];
			for (u32 f = 0; e->forces && f < array_length(e->forces); ++f) {
				const double fp[3] = {e->forces[f].position.x, e->forces[f].position.y, e->forces[f].position.z};
				const double fv[3] = {e->forces[f].force.x, e->forces[f].force.y, e->forces[f].force.z};
				if (rp_batch_add_force(g_batch, (int)i, fp, fv) != RP_OK) die("rp_batch_add_force");
			}
		}
		// (RP_ERR_CAPACITY: the step ran, but a fixed device buffer overflowed -- handled through the status word just below)
		const int step_rc = rp_batch_step_host(g_batch, in.data(), out.data(), dt, substeps, iters, collisions ? 1 : 0);
		if (step_rc != RP_OK && step_rc != RP_ERR_CAPACITY) die("rp_batch_step_host");
		int32_t status = 0;
		if (rp_batch_get_status(g_batch, &status) != RP_OK) die("rp_batch_get_status");
		if (status && rp_batch_clear_status(g_batch) != RP_OK) die("rp_batch_clear_status");
		const int32_t capacity = RP_ST_PAIR_CAPACITY | RP_ST_CONTACT_CAPACITY;
		if ((status & capacity) && attempt < 8) {
			// a fixed device capacity ran out: the input records are still on the host, so grow and redo the frame
			if (status & RP_ST_PAIR_CAPACITY) g_max_pairs = g_max_pairs ? 2 * g_max_pairs : 4 * (n + 64) * 4;
			if (status & RP_ST_CONTACT_CAPACITY) g_max_contacts = g_max_contacts ? 2 * g_max_contacts : 64 * (n + 16);
			make_batch();
			continue;
		}
		if (status & RP_ST_EPA_NO_CONVERGENCE) printf("Warning: EPA did not converge.\n");  // epa.cpp:233 carries on the same way
		if (status & ~RP_ST_EPA_NO_CONVERGENCE) {
			fprintf(stderr, "pbd_b200: status 0x%x -- the reference asserts at this point (include/rawphys_b200.h RP_ST_*)\n", (unsigned)status);
			abort();
		}
		break;
	}
	for (u32 i = 0; i < n; ++i) {
		Entity* e = es[i];
		const double* r = &out[(size_t)i * RP_STATE_STRIDE];
		e->world_position = (vec3){r[0], r[1], r[2]};
		e->world_rotation = (Quaternion){r[3], r[4], r[5], r[6]};
		e->linear_velocity = (vec3){r[7], r[8], r[9]};
		e->angular_velocity = (vec3){r[10], r[11], r[12]};
		e->active = r[13] != 0.0;
		e->deactivation_time = r[14];
		e->previous_linear_velocity = (vec3){r[15], r[16], r[17]};
		e->previous_angular_velocity = (vec3){r[18], r[19], r[20]};
		// previous_world_position / previous_world_rotation are overwritten at the top of every substep before any use
		// (pbd.cpp:540-541), so their values after the call are not observable and stay as they were
	}
}

}  // namespace

#ifdef RP_SHIM_WRAP
extern "C" void __wrap__Z29pbd_simulate_with_constraintsdPP6EntityP10Constraintjji(r64 dt, Entity** entities, Constraint* external_constraints,
	u32 num_substeps, u32 num_pos_iters, boolean enable_collisions) {
	step(dt, entities, external_constraints, num_substeps, num_pos_iters, enable_collisions);
}
extern "C" void __wrap__Z12pbd_simulatedPP6Entityjji(r64 dt, Entity** entities, u32 num_substeps, u32 num_pos_iters, boolean enable_collisions) {
	step(dt, entities, NULL, num_substeps, num_pos_iters, enable_collisions);
}
#else
void pbd_simulate_with_constraints(r64 dt, Entity** entities, Constraint* external_constraints, u32 num_substeps, u32 num_pos_iters,
	boolean enable_collisions) {
	step(dt, entities, external_constraints, num_substeps, num_pos_iters, enable_collisions);
}
void pbd_simulate(r64 dt, Entity** entities, u32 num_substeps, u32 num_pos_iters, boolean enable_collisions) {
	step(dt, entities, NULL, num_substeps, num_pos_iters, enable_collisions);
}
#endif

// lets a host program release the device before exit (optional; the reference has no counterpart)
extern "C" void pbd_b200_shutdown(void) {
	drop_batch();
	if (g_scene) rp_scene_destroy(g_scene);
	g_scene = 0;
	g_image.bytes.clear();
}
