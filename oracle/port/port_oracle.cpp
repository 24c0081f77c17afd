// oracle/port/port_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, loaded by or called from the product).
//
// CPU restatement of the reference's frame step: a plain sequential walk over one world in the reference's own order
// (broadphase -> islands/sleep -> per substep {integrate, contacts, Gauss-Seidel positional solve, velocity derive,
// velocity solve}), pbd_simulate_with_constraints, src/physics/pbd.cpp:468-747. The per-body / per-pair / per-constraint
// arithmetic is the shared host+device core under raw-physics_b200/csrc (rp_math.h, rp_shape.h, rp_narrow.h, rp_solve.h,
// rp_scene.cpp), compiled here for the host with g++ -O2 -ffp-contract=off; each of those routines cites the reference
// file:line it follows. What this file adds is only the ORDER of operations, which is the reference's.
//
// Pinning: this restatement is checked bit-for-bit against the UNMODIFIED reference compiled into
// oracle/_ref/libref_oracle.so (tests/test_oracle.py): hull topology, per-pair GJK/EPA/manifold outputs,
// per-substep contact logs and whole trajectories for every scene in tests/scenes.py. The reference has no tests or
// golden vectors of its own; the committed fixtures under tests/golden/ were generated from that library.
//
// Its role next to the CUDA path: (1) it lets the arithmetic core be debugged in a container without a GPU, and
// (2) it is the `port` CPU baseline / checker on machines where oracle/_ref is absent.
//
// The exported functions mirror oracle/ref_driver.cpp one for one (ref_* -> port_*), so tests drive both through the
// same Python class.
#include <chrono>
#include <random>
#include <string.h>
#include <vector>

#include "rp_scene.h"

using namespace rp;

namespace {

struct SeqContact {
	int e1, e2;
	V3 normal;
	Contact c;
};
struct LogEntry {
	uint32_t e1, e2, count, first;
};
struct LoggedContact {
	V3 p1, p2, n;
};
struct PersistentForce {
	uint32_t body;
	V3 position, force;
};

struct World {
	Scene scene;
	HullPoolHost pool;
	bool pooled = false;
	std::vector<Body> bodies;
	std::vector<double> deact;
	std::vector<V3> tv, tn;
	std::vector<JointLambda> lambdas;
	std::vector<PersistentForce> forces;
	bool gravity = false;
	double gravity_value = 10.0;
	int status = 0;
	bool log_enabled = false;
	std::vector<LogEntry> log;
	std::vector<LoggedContact> log_contacts;
	uint64_t calls = 0;
};

World* g = 0;

void sync_bodies(World& w) {  // instantiate per-world state for bodies added since the last call
	while (w.bodies.size() < w.scene.bodies.size()) {
		const BodyInit& bi = w.scene.bodies[w.bodies.size()];
		Body b;
		memset(&b, 0, sizeof(b));
		b.x = bi.x; b.q = bi.q;
		b.inv_mass = bi.inv_mass;
		b.inertia = bi.inertia; b.inv_inertia = bi.inv_inertia;
		b.mu_s = bi.mu_s; b.mu_d = bi.mu_d; b.rest = bi.rest;
		b.ii_bound = tensor_bound(bi.inv_inertia);
		b.fixed = bi.fixed;
		b.active = 1;
		w.bodies.push_back(b);
		w.deact.push_back(0.0);
	}
	if (!w.pooled || w.tv.size() != (size_t)w.scene.total_tv) {
		w.pool = pool_hulls(w.scene);
		w.pooled = true;
		w.tv.resize(w.scene.total_tv);
		w.tn.resize(w.scene.total_tn);
	}
	w.lambdas.resize(w.scene.joints.size());
}

// colliders_update (collider.cpp:440-445) for every collider of one body
void update_colliders(World& w, int bi) {
	const BodyInit& init = w.scene.bodies[bi];
	const Body& b = w.bodies[bi];
	Pose34 M = model_matrix(b.q, b.x);
	for (int c = init.col0; c < init.col0 + init.ncol; ++c) {
		const ColliderDesc& cd = w.scene.colliders[c];
		if (cd.type == SHAPE_SPHERE) {
			w.tv[cd.tv0] = b.x;
		} else {
			const HullTopo& t = w.pool.hulls[cd.hull];
			for (int k = 0; k < t.nv; ++k) w.tv[cd.tv0 + k] = transform_point(M, w.pool.verts[t.vert0 + k]);
			for (int k = 0; k < t.nf; ++k) w.tn[cd.tn0 + k] = transform_normal(M, w.pool.normals[t.face0 + k]);
		}
	}
}

static int g_epa_reruns = 0, g_clip_reruns = 0;  // pairs that overflowed the small stores (port_tier_reruns)

struct VecSink {
	std::vector<LoggedContact>* out;
	V3 n;
	void operator()(V3 p1, V3 p2) {
		LoggedContact c;
		c.p1 = p1; c.p2 = p2; c.n = n;
		out->push_back(c);
	}
};

// collider_get_contacts (collider.cpp:523-558) for one collider pair; contacts are appended to `out`
void narrow_pair(World& w, int ca, int cb, std::vector<LoggedContact>& out) {
	HullPool pv = w.pool.view();
	Shape A = make_shape(pv, w.scene.colliders[ca], w.tv.data(), w.tn.data());
	Shape B = make_shape(pv, w.scene.colliders[cb], w.tv.data(), w.tn.data());
	V3 normal;
	double depth;
	static EpaScratch es;
	static ClipScratch cs;
	// the small stores the CUDA kernels keep in shared memory, with the full ones as second tier (rp_narrow.h)
	static EpaSmallArrays es_small;
	static ClipSmallArrays cs_small;
	int sup_a = -1, sup_b = -1;
	if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
		if (!sphere_sphere(A, B, &normal, &depth)) return;
	} else {
		Simplex s;
		if (!gjk(A, B, &s, &w.status, 0)) return;
		if (!epa_tiered(A, B, s, es_small, es, &normal, &depth, &w.status, &g_epa_reruns, &sup_a, &sup_b)) return;
	}
	VecSink sink;
	sink.out = &out;
	sink.n = normal;
	// (EPA's last support vertices stand in for the manifold's own support scans, as on the device)
	manifold_tiered(A, B, normal, depth, cs_small, cs, &w.status, sink, &g_clip_reruns, sup_a, sup_b);
}

int uf_find(std::vector<int>& p, int x) {
	while (p[x] != x) x = p[x];
	return x;
}

// pbd_simulate_with_constraints (pbd.cpp:468-747)
void simulate(World& w, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	if (dt <= 0.0) return;
	const int n = (int)w.bodies.size();
	double h = dt / substeps;

	// broad_get_collision_pairs (broad.cpp:6-29)
	std::vector<std::pair<int, int>> pairs;
	for (int i = 0; i < n; ++i) {
		for (int j = i + 1; j < n; ++j) {
			double dist = length(sub(w.bodies[i].x, w.bodies[j].x));
			double maxd = w.scene.bodies[i].radius + w.scene.bodies[j].radius + 0.1;
			if (dist <= maxd) pairs.push_back(std::make_pair(i, j));
		}
	}

	// broad_collect_simulation_islands (broad.cpp:70-116) + sleep bookkeeping (pbd.cpp:476-506)
	{
		std::vector<int> parent(n);
		for (int i = 0; i < n; ++i) parent[i] = i;
		for (size_t k = 0; k < pairs.size(); ++k) {
			int a = pairs[k].first, b = pairs[k].second;
			if (!w.bodies[a].fixed && !w.bodies[b].fixed) parent[uf_find(parent, b)] = uf_find(parent, a);
		}
		for (size_t k = 0; k < w.scene.joints.size(); ++k) {
			int a = w.scene.joints[k].e1, b = w.scene.joints[k].e2;
			if (!w.bodies[a].fixed && !w.bodies[b].fixed) parent[uf_find(parent, b)] = uf_find(parent, a);
		}
		std::vector<char> all_inactive(n, 1);
		for (int i = 0; i < n; ++i) {
			Body& b = w.bodies[i];
			if (b.fixed) continue;
			if (length(b.v) < 0.10 && length(b.w) < 0.10) w.deact[i] += dt;
			else w.deact[i] = 0.0;
			if (w.deact[i] < 1.0) all_inactive[uf_find(parent, i)] = 0;
		}
		for (int i = 0; i < n; ++i) {
			if (!w.bodies[i].fixed) w.bodies[i].active = !all_inactive[uf_find(parent, i)];
		}
	}

	std::vector<SeqContact> contacts;
	std::vector<LoggedContact> found;
	for (uint32_t s = 0; s < substeps; ++s) {
		for (int i = 0; i < n; ++i) integrate(w.bodies[i], h, w.scene.force[i], w.scene.torque[i]);
		for (size_t k = 0; k < w.lambdas.size(); ++k) w.lambdas[k].a = w.lambdas[k].b = w.lambdas[k].c = 0.0;  // copy_constraints

		contacts.clear();
		if (collisions) {
			for (size_t k = 0; k < pairs.size(); ++k) {
				int a = pairs[k].first, b = pairs[k].second;
				const Body& b1 = w.bodies[a];
				const Body& b2 = w.bodies[b];
				if ((b1.fixed || !b1.active) && (b2.fixed || !b2.active)) continue;
				update_colliders(w, a);
				update_colliders(w, b);
				found.clear();
				const BodyInit& i1 = w.scene.bodies[a];
				const BodyInit& i2 = w.scene.bodies[b];
				for (int ca = i1.col0; ca < i1.col0 + i1.ncol; ++ca) {  // colliders_get_contacts (collider.cpp:560-572)
					for (int cb = i2.col0; cb < i2.col0 + i2.ncol; ++cb) narrow_pair(w, ca, cb, found);
				}
				++w.calls;
				if (w.log_enabled) {
					LogEntry le;
					le.e1 = (uint32_t)a; le.e2 = (uint32_t)b; le.count = (uint32_t)found.size(); le.first = (uint32_t)w.log_contacts.size();
					w.log.push_back(le);
					w.log_contacts.insert(w.log_contacts.end(), found.begin(), found.end());
				}
				for (size_t l = 0; l < found.size(); ++l) {
					SeqContact sc;
					sc.e1 = a; sc.e2 = b;
					sc.normal = found[l].n;
					sc.c = make_contact(b1, b2, found[l].p1, found[l].p2);
					contacts.push_back(sc);
				}
			}
		}

		for (uint32_t it = 0; it < iters; ++it) {
			for (size_t k = 0; k < w.scene.joints.size(); ++k) {
				const Joint& j = w.scene.joints[k];
				solve_joint(j, w.lambdas[k], w.bodies[j.e1], w.bodies[j.e2], h, &w.status);
			}
			for (size_t k = 0; k < contacts.size(); ++k) {
				SeqContact& sc = contacts[k];
				solve_contact(sc.c, sc.normal, w.bodies[sc.e1], w.bodies[sc.e2], h, &w.status);
			}
		}
		for (int i = 0; i < n; ++i) derive_velocity(w.bodies[i], h);
		for (size_t k = 0; k < contacts.size(); ++k) {
			SeqContact& sc = contacts[k];
			solve_contact_velocity(sc.c, sc.normal, w.bodies[sc.e1], w.bodies[sc.e2], h);
		}
	}
}

void frame(World& w, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	sync_bodies(w);
	w.scene.clear_forces();
	if (w.gravity) w.scene.add_gravity(w.gravity_value);
	for (size_t i = 0; i < w.forces.size(); ++i) w.scene.add_force((int)w.forces[i].body, w.forces[i].position, w.forces[i].force);
	simulate(w, dt, substeps, iters, collisions);
	w.scene.clear_forces();
}

V3 vec(const double* p) { return v3(p[0], p[1], p[2]); }

Joint blank_joint(int type, uint64_t e1, uint64_t e2) {
	Joint j;
	memset(&j, 0, sizeof(j));
	j.type = type;
	j.e1 = (int)e1;
	j.e2 = (int)e2;
	return j;
}

}  // namespace

extern "C" {

void port_reset() {
	delete g;
	g = new World();
}

void port_collider_begin() { g->scene.pending.clear(); }
void port_collider_add_hull(const double* v, uint32_t nv, const uint32_t* idx, uint32_t nidx) { g->scene.add_hull_collider(v, nv, idx, nidx); }
void port_collider_add_sphere(float radius) { g->scene.add_sphere_collider(radius); }
uint64_t port_entity_create(const double* pos, const double* quat, double mass, int fixed, double mu_s, double mu_d, double rest) {
	uint64_t id = (uint64_t)g->scene.add_body(pos, quat, mass, fixed, mu_s, mu_d, rest);
	sync_bodies(*g);
	return id;
}
uint32_t port_num_entities() { return (uint32_t)g->scene.bodies.size(); }

void port_quaternion_new(const double* axis, double angle_degrees, double* out) {  // quaternion_new (quaternion.cpp:18-31)
	V3 a = vec(axis);
	if (length(a) != 0.0) a = normalize(a);
	double rad = RP_PI_F * angle_degrees / 180.0;  // gm_radians (gm.h:757)
	double s = sin(rad / 2.0);
	out[3] = cos(rad / 2.0);
	out[0] = a.x * s; out[1] = a.y * s; out[2] = a.z * s;
}

void port_add_persistent_force(uint32_t body, const double* position, const double* force) {
	PersistentForce f;
	f.body = body; f.position = vec(position); f.force = vec(force);
	g->forces.push_back(f);
}
void port_set_gravity(int enabled, double gval) { g->gravity = enabled != 0; g->gravity_value = gval; }

void port_add_positional_constraint(uint64_t e1, uint64_t e2, const double* r1, const double* r2, double compliance, const double* distance) {
	Joint j = blank_joint(JOINT_POSITIONAL, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2); j.compliance = compliance; j.distance = vec(distance);
	g->scene.joints.push_back(j);
}
void port_add_mutual_orientation_constraint(uint64_t e1, uint64_t e2, double compliance) {
	Joint j = blank_joint(JOINT_MUTUAL_ORIENTATION, e1, e2);
	j.compliance = compliance;
	g->scene.joints.push_back(j);
}
void port_add_hinge_constraint(uint64_t e1, uint64_t e2, const double* r1, const double* r2, double compliance, int e1_aligned, int e2_aligned,
	int limited, int e1_limit, int e2_limit, double lower, double upper) {
	Joint j = blank_joint(JOINT_HINGE, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2); j.compliance = compliance;
	j.axis[0] = e1_aligned; j.axis[1] = e2_aligned; j.axis[2] = e1_limit; j.axis[3] = e2_limit;
	j.limited = limited; j.lower = lower; j.upper = upper;
	g->scene.joints.push_back(j);
}
void port_add_spherical_constraint(uint64_t e1, uint64_t e2, const double* r1, const double* r2, int e1_swing, int e2_swing, int e1_twist,
	int e2_twist, double swing_lower, double swing_upper, double twist_lower, double twist_upper) {
	Joint j = blank_joint(JOINT_SPHERICAL, e1, e2);
	j.r1_lc = vec(r1); j.r2_lc = vec(r2);
	j.axis[0] = e1_swing; j.axis[1] = e2_swing; j.axis[2] = e1_twist; j.axis[3] = e2_twist;
	j.lower = swing_lower; j.upper = swing_upper; j.lower2 = twist_lower; j.upper2 = twist_upper;
	g->scene.joints.push_back(j);
}

void port_step(double dt, uint32_t substeps, uint32_t iters, int collisions) { frame(*g, dt, substeps, iters, collisions); }

double port_run_timed(uint32_t frames, double dt, uint32_t substeps, uint32_t iters, int collisions) {
	auto t0 = std::chrono::steady_clock::now();
	for (uint32_t f = 0; f < frames; ++f) frame(*g, dt, substeps, iters, collisions);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

#define STATE_STRIDE 15
void port_get_state(double* out) {
	for (size_t i = 0; i < g->bodies.size(); ++i) {
		const Body& b = g->bodies[i];
		double* o = out + STATE_STRIDE * i;
		o[0] = b.x.x; o[1] = b.x.y; o[2] = b.x.z;
		o[3] = b.q.x; o[4] = b.q.y; o[5] = b.q.z; o[6] = b.q.w;
		o[7] = b.v.x; o[8] = b.v.y; o[9] = b.v.z;
		o[10] = b.w.x; o[11] = b.w.y; o[12] = b.w.z;
		o[13] = b.active ? 1.0 : 0.0;
		o[14] = g->deact[i];
	}
}
// previous_linear_velocity / previous_angular_velocity of every body (entity.h:45-46), 6 doubles each
void port_get_prev_velocities(double* out) {
	for (size_t i = 0; i < g->bodies.size(); ++i) {
		const Body& b = g->bodies[i];
		double* o = out + 6 * i;
		o[0] = b.pv.x; o[1] = b.pv.y; o[2] = b.pv.z; o[3] = b.pw.x; o[4] = b.pw.y; o[5] = b.pw.z;
	}
}
void port_set_state(const double* in) {
	sync_bodies(*g);
	for (size_t i = 0; i < g->bodies.size(); ++i) {
		Body& b = g->bodies[i];
		const double* o = in + STATE_STRIDE * i;
		b.x = v3(o[0], o[1], o[2]);
		b.q = q4(o[3], o[4], o[5], o[6]);
		b.v = v3(o[7], o[8], o[9]);
		b.w = v3(o[10], o[11], o[12]);
		b.active = o[13] != 0.0;
		g->deact[i] = o[14];
	}
}

#define PARAM_STRIDE 25
void port_get_params(double* out) {
	for (size_t i = 0; i < g->scene.bodies.size(); ++i) {
		const BodyInit& b = g->scene.bodies[i];
		double* o = out + PARAM_STRIDE * i;
		o[0] = b.inv_mass;
		for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
			o[1 + 3 * r + c] = b.inertia.m[r][c];
			o[10 + 3 * r + c] = b.inv_inertia.m[r][c];
		}
		o[19] = b.radius; o[20] = b.mu_s; o[21] = b.mu_d; o[22] = b.rest; o[23] = b.fixed ? 1.0 : 0.0; o[24] = (double)b.ncol;
	}
}

void port_hull_sizes(uint32_t body, uint32_t collider, int32_t* out6) {
	const ColliderDesc& c = g->scene.colliders[g->scene.bodies[body].col0 + collider];
	if (c.type != SHAPE_HULL) { out6[0] = -1; return; }
	const HullHost& h = g->scene.hulls[c.hull];
	out6[0] = (int32_t)h.verts.size(); out6[1] = (int32_t)h.normals.size(); out6[2] = (int32_t)h.face_idx.size();
	out6[3] = (int32_t)h.v2f_idx.size(); out6[4] = (int32_t)h.v2n_idx.size(); out6[5] = (int32_t)h.f2n_idx.size();
}

static void copy_u32(const std::vector<int>& v, uint32_t* out) { for (size_t i = 0; i < v.size(); ++i) out[i] = (uint32_t)v[i]; }

void port_hull_dump(uint32_t body, uint32_t collider, double* verts, double* normals, uint32_t* face_ptr, uint32_t* face_idx, uint32_t* v2f_ptr,
	uint32_t* v2f_idx, uint32_t* v2n_ptr, uint32_t* v2n_idx, uint32_t* f2n_ptr, uint32_t* f2n_idx) {
	const ColliderDesc& c = g->scene.colliders[g->scene.bodies[body].col0 + collider];
	const HullHost& h = g->scene.hulls[c.hull];
	memcpy(verts, h.verts.data(), sizeof(V3) * h.verts.size());
	memcpy(normals, h.normals.data(), sizeof(V3) * h.normals.size());
	copy_u32(h.face_ptr, face_ptr); copy_u32(h.face_idx, face_idx);
	copy_u32(h.v2f_ptr, v2f_ptr); copy_u32(h.v2f_idx, v2f_idx);
	copy_u32(h.v2n_ptr, v2n_ptr); copy_u32(h.v2n_idx, v2n_idx);
	copy_u32(h.f2n_ptr, f2n_ptr); copy_u32(h.f2n_idx, f2n_idx);
}

// same output layout as ref_probe_pair (oracle/ref_driver.cpp)
void port_probe_pair(uint32_t ia, uint32_t ca, uint32_t ib, uint32_t cb, double* out, double* contacts_out, uint32_t max_contacts) {
	World& w = *g;
	sync_bodies(w);
	update_colliders(w, (int)ia);
	update_colliders(w, (int)ib);
	HullPool pv = w.pool.view();
	Shape A = make_shape(pv, w.scene.colliders[w.scene.bodies[ia].col0 + ca], w.tv.data(), w.tn.data());
	Shape B = make_shape(pv, w.scene.colliders[w.scene.bodies[ib].col0 + cb], w.tv.data(), w.tn.data());
	for (int i = 0; i < 19; ++i) out[i] = 0.0;
	Simplex s;
	int st = 0;
	bool hit = gjk(A, B, &s, &st, 0);
	out[0] = hit ? 1.0 : 0.0;
	if (!hit) return;
	const V3* sv[4] = {&s.a, &s.b, &s.c, &s.d};
	for (int i = 0; i < 4; ++i) { out[1 + 3 * i] = sv[i]->x; out[2 + 3 * i] = sv[i]->y; out[3 + 3 * i] = sv[i]->z; }
	static EpaScratch es;
	static ClipScratch cs;
	V3 n;
	double depth;
	bool ok = epa(A, B, s, es, &n, &depth, &st, 0);
	out[13] = ok ? 1.0 : 0.0;
	if (!ok) return;
	out[14] = n.x; out[15] = n.y; out[16] = n.z; out[17] = depth;
	std::vector<LoggedContact> found;
	VecSink sink;
	sink.out = &found;
	sink.n = n;
	manifold(A, B, n, depth, cs, &st, sink);
	out[18] = (double)found.size();
	for (size_t i = 0; i < found.size() && i < max_contacts; ++i) {
		double* o = contacts_out + 9 * i;
		o[0] = found[i].p1.x; o[1] = found[i].p1.y; o[2] = found[i].p1.z;
		o[3] = found[i].p2.x; o[4] = found[i].p2.y; o[5] = found[i].p2.z;
		o[6] = found[i].n.x; o[7] = found[i].n.y; o[8] = found[i].n.z;
	}
}

uint32_t port_broad_pairs(uint64_t* out_pairs, uint32_t max_pairs) {
	World& w = *g;
	uint32_t n = 0;
	for (size_t i = 0; i < w.bodies.size(); ++i) {
		for (size_t j = i + 1; j < w.bodies.size(); ++j) {
			double dist = length(sub(w.bodies[i].x, w.bodies[j].x));
			if (dist <= w.scene.bodies[i].radius + w.scene.bodies[j].radius + 0.1) {
				if (n < max_pairs) { out_pairs[2 * n] = i; out_pairs[2 * n + 1] = j; }
				++n;
			}
		}
	}
	return n;
}

void port_log_enable(int on) { g->log_enabled = on != 0; }
void port_log_clear() { g->log.clear(); g->log_contacts.clear(); }
uint32_t port_log_num_calls() { return (uint32_t)g->log.size(); }
uint32_t port_log_num_contacts() { return (uint32_t)g->log_contacts.size(); }
void port_log_get(uint32_t* calls_out, double* contacts_out) {
	for (size_t i = 0; i < g->log.size(); ++i) {
		calls_out[4 * i] = g->log[i].e1; calls_out[4 * i + 1] = g->log[i].e2;
		calls_out[4 * i + 2] = g->log[i].count; calls_out[4 * i + 3] = g->log[i].first;
	}
	if (contacts_out) {
		for (size_t i = 0; i < g->log_contacts.size(); ++i) {
			double* o = contacts_out + 9 * i;
			const LoggedContact& c = g->log_contacts[i];
			o[0] = c.p1.x; o[1] = c.p1.y; o[2] = c.p1.z; o[3] = c.p2.x; o[4] = c.p2.y; o[5] = c.p2.z; o[6] = c.n.x; o[7] = c.n.y; o[8] = c.n.z;
		}
	}
}
uint64_t port_total_narrowphase_calls() { return g->calls; }

// Soundness probe for the bounds cull of the CUDA path (k_cull, RP_CULL_MARGIN = 1e-7): random box / octahedron pairs at
// random, axis-aligned and nearly-touching poses; counts pairs whose world-space bounds are separated by more than the
// margin and which the reference-order GJK nevertheless reports as colliding. Must return 0.
uint64_t port_cull_soundness(uint64_t trials, uint64_t seed, uint64_t* separated_out, uint64_t* hits_out) {
	std::mt19937_64 rng(seed);
	std::uniform_real_distribution<double> U(-1.0, 1.0);
	std::vector<V3> la, lb, ta, tb;
	uint64_t viol = 0, sep = 0, hits = 0;
	for (uint64_t t = 0; t < trials; ++t) {
		la.clear(); lb.clear(); ta.clear(); tb.clear();
		for (int side = 0; side < 2; ++side) {
			std::vector<V3>& v = side ? lb : la;
			double big = side ? 25.0 : 2.0;
			if (t % (side ? 5 : 3) == 0) {
				double s = 0.5 + fabs(U(rng));
				v.push_back(v3(s, 0, 0)); v.push_back(v3(-s, 0, 0)); v.push_back(v3(0, s, 0));
				v.push_back(v3(0, -s, 0)); v.push_back(v3(0, 0, s)); v.push_back(v3(0, 0, -s));
			} else {
				double sx = 0.2 + fabs(U(rng)) * big, sy = 0.2 + fabs(U(rng)), sz = 0.2 + fabs(U(rng)) * big;
				for (int i = 0; i < 8; ++i) v.push_back(v3((i & 1) ? sx : -sx, (i & 2) ? sy : -sy, (i & 4) ? sz : -sz));
			}
		}
		Q4 qa = normalize(q4(U(rng), U(rng), U(rng), U(rng))), qb = normalize(q4(U(rng), U(rng), U(rng), U(rng)));
		if (t % 4 == 0) { qa = q4(0, 0, 0, 1); qb = q4(0, 0, 0, 1); }
		V3 xa = v3(3 * U(rng), 3 * U(rng), 3 * U(rng)), xb = v3(3 * U(rng), 3 * U(rng), 3 * U(rng));
		if (t % 7 == 0) xb = v3(xa.x, xa.y - 1.0 - 2e-6 * U(rng), xa.z);
		Pose34 Ma = model_matrix(qa, xa), Mb = model_matrix(qb, xb);
		double lo[2][3], hi[2][3];
		for (int side = 0; side < 2; ++side) {
			for (int k = 0; k < 3; ++k) { lo[side][k] = 1e300; hi[side][k] = -1e300; }
			const std::vector<V3>& src = side ? lb : la;
			std::vector<V3>& dst = side ? tb : ta;
			for (size_t i = 0; i < src.size(); ++i) {
				V3 p = transform_point(side ? Mb : Ma, src[i]);
				dst.push_back(p);
				double c[3] = {p.x, p.y, p.z};
				for (int k = 0; k < 3; ++k) { if (c[k] < lo[side][k]) lo[side][k] = c[k]; if (c[k] > hi[side][k]) hi[side][k] = c[k]; }
			}
		}
		Shape A, B;
		memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
		A.type = B.type = SHAPE_HULL;
		A.vp = (const double*)ta.data(); A.vs = 3; A.vcs = 1; A.nv = (int)ta.size();
		B.vp = (const double*)tb.data(); B.vs = 3; B.vcs = 1; B.nv = (int)tb.size();
		bool separated = false;
		for (int k = 0; k < 3; ++k) {
			if (lo[0][k] - hi[1][k] > 1e-7 || lo[1][k] - hi[0][k] > 1e-7) separated = true;
		}
		Simplex s;
		int st = 0;
		bool hit = gjk(A, B, &s, &st, 0);
		hits += hit;
		if (separated) { ++sep; if (hit) ++viol; }
	}
	if (separated_out) *separated_out = sep;
	if (hits_out) *hits_out = hits;
	return viol;
}
int port_status() { return g->status; }
// pairs whose polytope / clip polygon outgrew the small (shared-memory sized) stores and were rerun on the full ones
void port_tier_reruns(int out2[2]) { out2[0] = g_epa_reruns; out2[1] = g_clip_reruns; }

}

#if defined(RP_COUNT_FRICTION)
// diagnostics build only (make -C oracle port CXXFLAGS+=-DRP_COUNT_FRICTION): how often the static-friction bound of
// solve_contact settles the branch, how often the exact evaluation runs, and how often the branch is taken
extern "C" void port_friction_counts(long out3[3]) {
	out3[0] = rp::g_friction_skipped; out3[1] = rp::g_friction_evaluated; out3[2] = rp::g_friction_taken;
}
#endif
