// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin extern "C" harness around the UNMODIFIED reference sources of felipeek/raw-physics, compiled where they
// lie under /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so (recipe: SURVEY.md 8c;
// flags -O2 -ffp-contract=off, asserts on, no -ffast-math). It replaces the GLFW loop (src/main.cpp:115-167) and
// the example scenes' init/update halves (e.g. src/examples/stack.cpp:35-104) with calls a test can drive:
// create entities from triangle soups, step pbd_simulate_with_constraints (src/physics/pbd.cpp:468), read the
// Entity state back, and log every colliders_get_contacts call (src/physics/collider.cpp:560) through a linker
// --wrap hook so per-substep contact sets can be compared.
//
// Nothing here restates reference arithmetic: every number comes out of the reference's own object code.
#define GRAPHICS_MATH_IMPLEMENT
#define C_FEK_HASH_MAP_IMPLEMENT
#include <gm.h>
#include <hash_map.h>
#include <light_array.h>
#include <vector>
#include <string.h>
#include <chrono>
#include "entity.h"
#include "physics/pbd.h"
#include "physics/gjk.h"
#include "physics/epa.h"
#include "physics/clipping.h"
#include "physics/support.h"
#include "physics/broad.h"

// entity.cpp globals (src/entity.cpp:7-9)
extern Entity** entities;
extern Hash_Map entities_map;
extern eid eid_counter;

// The two GL entry points entity.cpp references (src/entity.cpp:147-151); never called by the physics path.
extern "C" {
void* __glewDeleteBuffers = 0;
void* __glewDeleteVertexArrays = 0;
}

namespace {

struct Persistent_Force { u32 entity_index; vec3 position; vec3 force; };

bool g_inited = false;
std::vector<Persistent_Force> g_forces;
Constraint* g_constraints = NULL;  // light_array, NULL when no external constraints
Collider* g_pending_colliders = NULL;
bool g_gravity = false;
double g_gravity_value = 10.0;

// contact log
struct Contact_Log_Entry { u32 substep_call; u64 e1, e2; u32 count; u32 first_contact; };
std::vector<Contact_Log_Entry> g_log;
std::vector<Collider_Contact> g_log_contacts;
bool g_log_enabled = false;
u64 g_calls = 0;

Entity* find_owner(Collider* c) {
	for (u32 i = 0; i < array_length(entities); ++i) {
		if (entities[i]->colliders == c) return entities[i];
	}
	return NULL;
}

}

// --wrap hook: one call per narrowphase pair per substep, in pair order (src/physics/pbd.cpp:601)
extern "C" Collider_Contact* __real__Z22colliders_get_contactsP8ColliderS0_(Collider*, Collider*);
extern "C" Collider_Contact* __wrap__Z22colliders_get_contactsP8ColliderS0_(Collider* c1, Collider* c2) {
	Collider_Contact* r = __real__Z22colliders_get_contactsP8ColliderS0_(c1, c2);
	++g_calls;
	if (g_log_enabled) {
		Contact_Log_Entry e;
		Entity* e1 = find_owner(c1);
		Entity* e2 = find_owner(c2);
		e.substep_call = (u32)g_calls;
		e.e1 = e1 ? e1->id : (u64)-1;
		e.e2 = e2 ? e2->id : (u64)-1;
		e.count = r ? (u32)array_length(r) : 0;
		e.first_contact = (u32)g_log_contacts.size();
		for (u32 i = 0; i < e.count; ++i) g_log_contacts.push_back(r[i]);
		g_log.push_back(e);
	}
	return r;
}

extern "C" {

void ref_reset() {
	if (g_inited) {
		Entity** all = entity_get_all();
		for (u32 i = 0; i < array_length(all); ++i) {
			Entity* e = all[i];
			colliders_destroy(e->colliders);
			array_free(e->colliders);
			entity_destroy(e);
		}
		array_free(all);
		entity_module_destroy();
	}
	entity_module_init();
	eid_counter = 0;  // q13: ids == array indices for every scene built through this harness
	g_inited = true;
	g_forces.clear();
	if (g_constraints) { array_free(g_constraints); g_constraints = NULL; }
	g_pending_colliders = NULL;
	g_gravity = false;
	g_log.clear();
	g_log_contacts.clear();
	g_log_enabled = false;
	g_calls = 0;
}

// --- collider assembly for the next entity (mirrors examples_util.cpp:5-31 without the float->double scale step,
// which the caller performs so the same doubles reach both implementations)
void ref_collider_begin() { g_pending_colliders = array_new(Collider); }

void ref_collider_add_hull(const double* vertices_xyz, u32 num_vertices, const u32* indices, u32 num_indices) {
	vec3* v = array_new(vec3);
	for (u32 i = 0; i < num_vertices; ++i) {
		vec3 p = (vec3){vertices_xyz[3 * i + 0], vertices_xyz[3 * i + 1], vertices_xyz[3 * i + 2]};
		array_push(v, p);
	}
	u32* idx = array_new(u32);
	for (u32 i = 0; i < num_indices; ++i) array_push(idx, indices[i]);
	Collider c = collider_convex_hull_create(v, idx);
	array_free(v);
	array_free(idx);
	array_push(g_pending_colliders, c);
}

void ref_collider_add_sphere(float radius) {
	Collider c = collider_sphere_create(radius);
	array_push(g_pending_colliders, c);
}

// returns the entity id (== index)
u64 ref_entity_create(const double* pos, const double* quat_xyzw, double mass, int fixed, double mu_s, double mu_d, double restitution) {
	Mesh m;
	memset(&m, 0, sizeof(m));
	vec3 p = (vec3){pos[0], pos[1], pos[2]};
	Quaternion q = (Quaternion){quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3]};
	vec3 scale = (vec3){1.0, 1.0, 1.0};
	vec4 color = (vec4){1.0, 1.0, 1.0, 1.0};
	eid id;
	if (fixed) id = entity_create_fixed(m, p, q, scale, color, g_pending_colliders, mu_s, mu_d, restitution);
	else id = entity_create(m, p, q, scale, color, mass, g_pending_colliders, mu_s, mu_d, restitution);
	g_pending_colliders = NULL;
	return id;
}

u32 ref_num_entities() { return (u32)array_length(entities); }

// quaternion_new (degrees) as the examples call it, so tests can build the same initial rotations (src/quaternion.cpp:18)
void ref_quaternion_new(const double* axis, double angle_degrees, double* out_xyzw) {
	Quaternion q = quaternion_new((vec3){axis[0], axis[1], axis[2]}, angle_degrees);
	out_xyzw[0] = q.x; out_xyzw[1] = q.y; out_xyzw[2] = q.z; out_xyzw[3] = q.w;
}

// A force re-added before every step and cleared after it, as the examples' update() do (stack.cpp:93-102).
void ref_add_persistent_force(u32 entity_index, const double* position, const double* force) {
	Persistent_Force f;
	f.entity_index = entity_index;
	f.position = (vec3){position[0], position[1], position[2]};
	f.force = (vec3){force[0], force[1], force[2]};
	g_forces.push_back(f);
}

// The examples' gravity idiom: force (0, -G * 1.0 / inverse_mass, 0) at the COM of EVERY entity (stack.cpp:93-96).
void ref_set_gravity(int enabled, double g) { g_gravity = enabled != 0; g_gravity_value = g; }

static void ensure_constraints() { if (!g_constraints) g_constraints = array_new(Constraint); }

void ref_add_positional_constraint(u64 e1, u64 e2, const double* r1, const double* r2, double compliance, const double* distance) {
	ensure_constraints();
	Constraint c;
	memset(&c, 0, sizeof(c));
	pbd_positional_constraint_init(&c, e1, e2, (vec3){r1[0], r1[1], r1[2]}, (vec3){r2[0], r2[1], r2[2]}, compliance,
		(vec3){distance[0], distance[1], distance[2]});
	array_push(g_constraints, c);
}

void ref_add_mutual_orientation_constraint(u64 e1, u64 e2, double compliance) {
	ensure_constraints();
	Constraint c;
	memset(&c, 0, sizeof(c));
	pbd_mutual_orientation_constraint_init(&c, e1, e2, compliance);
	array_push(g_constraints, c);
}

void ref_add_hinge_constraint(u64 e1, u64 e2, const double* r1, const double* r2, double compliance, int e1_aligned, int e2_aligned,
	int limited, int e1_limit, int e2_limit, double lower, double upper) {
	ensure_constraints();
	Constraint c;
	memset(&c, 0, sizeof(c));
	if (limited) {
		pbd_hinge_joint_constraint_limited_init(&c, e1, e2, (vec3){r1[0], r1[1], r1[2]}, (vec3){r2[0], r2[1], r2[2]}, compliance,
			(PBD_Axis_Type)e1_aligned, (PBD_Axis_Type)e2_aligned, (PBD_Axis_Type)e1_limit, (PBD_Axis_Type)e2_limit, lower, upper);
	} else {
		pbd_hinge_joint_constraint_unlimited_init(&c, e1, e2, (vec3){r1[0], r1[1], r1[2]}, (vec3){r2[0], r2[1], r2[2]}, compliance,
			(PBD_Axis_Type)e1_aligned, (PBD_Axis_Type)e2_aligned);
	}
	array_push(g_constraints, c);
}

void ref_add_spherical_constraint(u64 e1, u64 e2, const double* r1, const double* r2, int e1_swing, int e2_swing, int e1_twist, int e2_twist,
	double swing_lower, double swing_upper, double twist_lower, double twist_upper) {
	ensure_constraints();
	Constraint c;
	memset(&c, 0, sizeof(c));
	pbd_spherical_joint_constraint_init(&c, e1, e2, (vec3){r1[0], r1[1], r1[2]}, (vec3){r2[0], r2[1], r2[2]},
		(PBD_Axis_Type)e1_swing, (PBD_Axis_Type)e2_swing, (PBD_Axis_Type)e1_twist, (PBD_Axis_Type)e2_twist,
		swing_lower, swing_upper, twist_lower, twist_upper);
	array_push(g_constraints, c);
}

// One frame exactly as an example's update() does it (stack.cpp:86-104 / hinge_joints.cpp:120-141):
// colliders_update for all, add forces, pbd_simulate_with_constraints, clear forces.
void ref_step(double dt, u32 num_substeps, u32 num_pos_iters, int enable_collisions) {
	Entity** all = entity_get_all();
	for (u32 i = 0; i < array_length(all); ++i) {
		Entity* e = all[i];
		colliders_update(e->colliders, e->world_position, &e->world_rotation);
	}
	if (g_gravity) {
		for (u32 i = 0; i < array_length(all); ++i) {
			entity_add_force(all[i], (vec3){0.0, 0.0, 0.0}, (vec3){0.0, -g_gravity_value * 1.0 / all[i]->inverse_mass, 0.0}, false);
		}
	}
	for (size_t i = 0; i < g_forces.size(); ++i) {
		entity_add_force(all[g_forces[i].entity_index], g_forces[i].position, g_forces[i].force, false);
	}
	pbd_simulate_with_constraints(dt, all, g_constraints, num_substeps, num_pos_iters, enable_collisions);
	for (u32 i = 0; i < array_length(all); ++i) entity_clear_forces(all[i]);
	array_free(all);
}

// Runs `frames` steps and returns wall seconds spent inside them (CPU baseline timing).
double ref_run_timed(u32 frames, double dt, u32 num_substeps, u32 num_pos_iters, int enable_collisions) {
	auto t0 = std::chrono::steady_clock::now();
	for (u32 f = 0; f < frames; ++f) ref_step(dt, num_substeps, num_pos_iters, enable_collisions);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// state record per entity: pos[3] quat[4] linvel[3] angvel[3] active deactivation_time  = 15 doubles
#define REF_STATE_STRIDE 15
void ref_get_state(double* out) {
	for (u32 i = 0; i < array_length(entities); ++i) {
		Entity* e = entities[i];
		double* o = out + REF_STATE_STRIDE * i;
		o[0] = e->world_position.x; o[1] = e->world_position.y; o[2] = e->world_position.z;
		o[3] = e->world_rotation.x; o[4] = e->world_rotation.y; o[5] = e->world_rotation.z; o[6] = e->world_rotation.w;
		o[7] = e->linear_velocity.x; o[8] = e->linear_velocity.y; o[9] = e->linear_velocity.z;
		o[10] = e->angular_velocity.x; o[11] = e->angular_velocity.y; o[12] = e->angular_velocity.z;
		o[13] = e->active ? 1.0 : 0.0;
		o[14] = e->deactivation_time;
	}
}

// previous_linear_velocity / previous_angular_velocity of every entity (entity.h:45-46), 6 doubles each
void ref_get_prev_velocities(double* out) {
	for (u32 i = 0; i < array_length(entities); ++i) {
		Entity* e = entities[i];
		double* o = out + 6 * i;
		o[0] = e->previous_linear_velocity.x; o[1] = e->previous_linear_velocity.y; o[2] = e->previous_linear_velocity.z;
		o[3] = e->previous_angular_velocity.x; o[4] = e->previous_angular_velocity.y; o[5] = e->previous_angular_velocity.z;
	}
}

void ref_set_state(const double* in) {
	for (u32 i = 0; i < array_length(entities); ++i) {
		Entity* e = entities[i];
		const double* o = in + REF_STATE_STRIDE * i;
		e->world_position = (vec3){o[0], o[1], o[2]};
		e->world_rotation = (Quaternion){o[3], o[4], o[5], o[6]};
		e->linear_velocity = (vec3){o[7], o[8], o[9]};
		e->angular_velocity = (vec3){o[10], o[11], o[12]};
		e->active = o[13] != 0.0;
		e->deactivation_time = o[14];
	}
}

// static per-entity parameters: inverse_mass, inertia[9], inverse_inertia[9], radius, mu_s, mu_d, e, fixed = 25 doubles
#define REF_PARAM_STRIDE 25
void ref_get_params(double* out) {
	for (u32 i = 0; i < array_length(entities); ++i) {
		Entity* e = entities[i];
		double* o = out + REF_PARAM_STRIDE * i;
		o[0] = e->inverse_mass;
		for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
			o[1 + 3 * r + c] = e->inertia_tensor.data[r][c];
			o[10 + 3 * r + c] = e->inverse_inertia_tensor.data[r][c];
		}
		o[19] = e->bounding_sphere_radius;
		o[20] = e->static_friction_coefficient;
		o[21] = e->dynamic_friction_coefficient;
		o[22] = e->restitution_coefficient;
		o[23] = e->fixed ? 1.0 : 0.0;
		o[24] = (double)array_length(e->colliders);
	}
}

// ---- hull topology dump (collider_convex_hull_create, src/physics/collider.cpp:194-364)
// sizes: [V, F, sum face elems, sum v2f, sum v2n, sum f2n]; -1 V for a sphere collider
void ref_hull_sizes(u32 entity_index, u32 collider_index, s32* out6) {
	Collider* c = &entities[entity_index]->colliders[collider_index];
	if (c->type != COLLIDER_TYPE_CONVEX_HULL) { out6[0] = -1; return; }
	Collider_Convex_Hull* h = &c->convex_hull;
	u32 V = (u32)array_length(h->vertices), F = (u32)array_length(h->faces);
	u32 fe = 0, v2f = 0, v2n = 0, f2n = 0;
	for (u32 i = 0; i < F; ++i) { fe += (u32)array_length(h->faces[i].elements); f2n += (u32)array_length(h->face_to_neighbors[i]); }
	for (u32 i = 0; i < V; ++i) { v2f += (u32)array_length(h->vertex_to_faces[i]); v2n += (u32)array_length(h->vertex_to_neighbors[i]); }
	out6[0] = V; out6[1] = F; out6[2] = fe; out6[3] = v2f; out6[4] = v2n; out6[5] = f2n;
}

// CSR dump in the reference's own order. ptr arrays have (count+1) entries.
void ref_hull_dump(u32 entity_index, u32 collider_index, double* verts, double* normals,
	u32* face_ptr, u32* face_idx, u32* v2f_ptr, u32* v2f_idx, u32* v2n_ptr, u32* v2n_idx, u32* f2n_ptr, u32* f2n_idx) {
	Collider_Convex_Hull* h = &entities[entity_index]->colliders[collider_index].convex_hull;
	u32 V = (u32)array_length(h->vertices), F = (u32)array_length(h->faces);
	for (u32 i = 0; i < V; ++i) { verts[3 * i] = h->vertices[i].x; verts[3 * i + 1] = h->vertices[i].y; verts[3 * i + 2] = h->vertices[i].z; }
	u32 a = 0, b = 0;
	for (u32 i = 0; i < F; ++i) {
		normals[3 * i] = h->faces[i].normal.x; normals[3 * i + 1] = h->faces[i].normal.y; normals[3 * i + 2] = h->faces[i].normal.z;
		face_ptr[i] = a; f2n_ptr[i] = b;
		for (u32 k = 0; k < array_length(h->faces[i].elements); ++k) face_idx[a++] = h->faces[i].elements[k];
		for (u32 k = 0; k < array_length(h->face_to_neighbors[i]); ++k) f2n_idx[b++] = h->face_to_neighbors[i][k];
	}
	face_ptr[F] = a; f2n_ptr[F] = b;
	a = 0; b = 0;
	for (u32 i = 0; i < V; ++i) {
		v2f_ptr[i] = a; v2n_ptr[i] = b;
		for (u32 k = 0; k < array_length(h->vertex_to_faces[i]); ++k) v2f_idx[a++] = h->vertex_to_faces[i][k];
		for (u32 k = 0; k < array_length(h->vertex_to_neighbors[i]); ++k) v2n_idx[b++] = h->vertex_to_neighbors[i][k];
	}
	v2f_ptr[V] = a; v2n_ptr[V] = b;
}

// ---- per-function known-answer probes on the CURRENT poses of two entities (first collider of each unless given)
// out: [0]=gjk verdict, [1..12]=simplex a,b,c,d, [13]=epa converged, [14..16]=normal, [17]=penetration, [18]=num contacts
// contacts_out: up to max_contacts * 9 doubles (p1, p2, n)
void ref_probe_pair(u32 ia, u32 ca, u32 ib, u32 cb, double* out, double* contacts_out, u32 max_contacts) {
	Entity* ea = entities[ia];
	Entity* eb = entities[ib];
	colliders_update(ea->colliders, ea->world_position, &ea->world_rotation);
	colliders_update(eb->colliders, eb->world_position, &eb->world_rotation);
	Collider* A = &ea->colliders[ca];
	Collider* B = &eb->colliders[cb];
	for (int i = 0; i < 19; ++i) out[i] = 0.0;
	GJK_Simplex s;
	memset(&s, 0, sizeof(s));
	boolean hit = gjk_collides(A, B, &s);
	out[0] = hit ? 1.0 : 0.0;
	if (!hit) return;
	const vec3* sv[4] = {&s.a, &s.b, &s.c, &s.d};
	for (int i = 0; i < 4; ++i) { out[1 + 3 * i] = sv[i]->x; out[2 + 3 * i] = sv[i]->y; out[3 + 3 * i] = sv[i]->z; }
	vec3 n; r64 pen;
	boolean ok = epa(A, B, &s, &n, &pen);
	out[13] = ok ? 1.0 : 0.0;
	if (!ok) return;
	out[14] = n.x; out[15] = n.y; out[16] = n.z; out[17] = pen;
	Collider_Contact* contacts = array_new_len(Collider_Contact, 16);
	clipping_get_contact_manifold(A, B, n, pen, &contacts);
	u32 nc = (u32)array_length(contacts);
	out[18] = (double)nc;
	for (u32 i = 0; i < nc && i < max_contacts; ++i) {
		double* o = contacts_out + 9 * i;
		o[0] = contacts[i].collision_point1.x; o[1] = contacts[i].collision_point1.y; o[2] = contacts[i].collision_point1.z;
		o[3] = contacts[i].collision_point2.x; o[4] = contacts[i].collision_point2.y; o[5] = contacts[i].collision_point2.z;
		o[6] = contacts[i].normal.x; o[7] = contacts[i].normal.y; o[8] = contacts[i].normal.z;
	}
	array_free(contacts);
}

// broadphase pairs on current poses (src/physics/broad.cpp:6-29); returns count, writes up to max pairs (e1,e2)
u32 ref_broad_pairs(u64* out_pairs, u32 max_pairs) {
	Entity** all = entity_get_all();
	Broad_Collision_Pair* p = broad_get_collision_pairs(all);
	u32 n = (u32)array_length(p);
	for (u32 i = 0; i < n && i < max_pairs; ++i) { out_pairs[2 * i] = p[i].e1_id; out_pairs[2 * i + 1] = p[i].e2_id; }
	array_free(p);
	array_free(all);
	return n;
}

// ---- contact log (filled by the --wrap hook)
void ref_log_enable(int on) { g_log_enabled = on != 0; }
void ref_log_clear() { g_log.clear(); g_log_contacts.clear(); }
u32 ref_log_num_calls() { return (u32)g_log.size(); }
u32 ref_log_num_contacts() { return (u32)g_log_contacts.size(); }
// calls_out: 4 u32 per call (e1, e2, count, first_contact); contacts_out: 9 doubles per contact
void ref_log_get(u32* calls_out, double* contacts_out) {
	for (size_t i = 0; i < g_log.size(); ++i) {
		calls_out[4 * i] = (u32)g_log[i].e1; calls_out[4 * i + 1] = (u32)g_log[i].e2;
		calls_out[4 * i + 2] = g_log[i].count; calls_out[4 * i + 3] = g_log[i].first_contact;
	}
	if (contacts_out) {
		for (size_t i = 0; i < g_log_contacts.size(); ++i) {
			double* o = contacts_out + 9 * i;
			const Collider_Contact& c = g_log_contacts[i];
			o[0] = c.collision_point1.x; o[1] = c.collision_point1.y; o[2] = c.collision_point1.z;
			o[3] = c.collision_point2.x; o[4] = c.collision_point2.y; o[5] = c.collision_point2.z;
			o[6] = c.normal.x; o[7] = c.normal.y; o[8] = c.normal.z;
		}
	}
}

u64 ref_total_narrowphase_calls() { return g_calls; }

}
