// rp_scene.cpp -- host-side scene template construction (see rp_scene.h).
#include "rp_scene.h"

#include <algorithm>
#include <chrono>
#include <string.h>

namespace rp {

namespace {

struct Tri {
	int x, y, z;
};
struct Edge {
	int x, y;
};

bool tris_share_vertex(const Tri& a, const Tri& b) {  // collider.cpp:27-31
	return a.x == b.x || a.x == b.y || a.x == b.z || a.y == b.x || a.y == b.y || a.y == b.z || a.z == b.x || a.z == b.y || a.z == b.z;
}

V3 tri_normal(const std::vector<V3>& hull, const Tri& t) {
	V3 v1 = hull[t.x], v2 = hull[t.y], v3_ = hull[t.z];
	return normalize(cross(sub(v2, v1), sub(v3_, v1)));
}

// collect_faces_planar_to (collider.cpp:47-78): depth-first flood fill over triangles sharing any vertex whose normal
// matches the seed's within 1e-6; visiting order = neighbour-list order, which fixes the order of `out`.
void flood_coplanar(const std::vector<V3>& hull, const std::vector<Tri>& tris, const std::vector<std::vector<int>>& nbrs,
	std::vector<char>& done, int t, V3 target, std::vector<Tri>& out) {
	const double EPS = 0.000001;
	V3 n = tri_normal(hull, tris[t]);
	if (done[t]) return;
	double proj = dot(n, target);
	if ((proj - 1.0) > -EPS && (proj - 1.0) < EPS) {
		out.push_back(tris[t]);
		done[t] = 1;
		for (size_t i = 0; i < nbrs[t].size(); ++i) flood_coplanar(hull, tris, nbrs, done, nbrs[t][i], target, out);
	}
}

int find_edge(const std::vector<Edge>& edges, Edge e) {  // get_edge_index (collider.cpp:80-92)
	for (size_t i = 0; i < edges.size(); ++i) {
		if (edges[i].x == e.x && edges[i].y == e.y) return (int)i;
		if (edges[i].x == e.y && edges[i].y == e.x) return (int)i;
	}
	return -1;
}

void toggle_edge(std::vector<Edge>& edges, Edge e) {
	int k = find_edge(edges, e);
	if (k >= 0) {  // array_remove: swap with last (light_array.h:146)
		edges[k] = edges.back();
		edges.pop_back();
	} else {
		edges.push_back(e);
	}
}

// create_convex_hull_face (collider.cpp:94-166): boundary loop of a set of coplanar triangles
std::vector<int> face_loop(const std::vector<Tri>& tris) {
	std::vector<Edge> edges;
	for (size_t i = 0; i < tris.size(); ++i) {
		toggle_edge(edges, Edge{tris[i].x, tris[i].y});
		toggle_edge(edges, Edge{tris[i].y, tris[i].z});
		toggle_edge(edges, Edge{tris[i].z, tris[i].x});
	}
	// "nicely order the edges": note there is no early exit from the inner loop in the reference
	for (size_t i = 0; i < edges.size(); ++i) {
		Edge cur = edges[i];
		for (size_t j = i + 1; j < edges.size(); ++j) {
			Edge cand = edges[j];
			if (cur.y != cand.x && cur.y != cand.y) continue;
			if (cur.y == cand.y) std::swap(cand.x, cand.y);
			Edge tmp = edges[i + 1];
			edges[i + 1] = cand;
			edges[j] = tmp;
		}
	}
	std::vector<int> loop;
	for (size_t i = 0; i < edges.size(); ++i) loop.push_back(edges[i].x);
	return loop;
}

bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

void to_csr(const std::vector<std::vector<int>>& lists, std::vector<int>& ptr, std::vector<int>& idx) {
	ptr.clear();
	idx.clear();
	for (size_t i = 0; i < lists.size(); ++i) {
		ptr.push_back((int)idx.size());
		idx.insert(idx.end(), lists[i].begin(), lists[i].end());
	}
	ptr.push_back((int)idx.size());
}

}  // namespace

// collider_convex_hull_create (collider.cpp:194-364)
HullHost build_hull(const double* vx, uint32_t nverts, const uint32_t* indices, uint32_t nidx) {
	HullHost h;
	h.key.assign(vx, vx + 3 * (size_t)nverts);
	h.key_idx.assign(indices, indices + nidx);

	// unique vertices in first-seen order (the reference keys a hash map by exact coordinate equality)
	std::vector<int> remap(nverts);
	for (uint32_t i = 0; i < nverts; ++i) {
		V3 p = v3(vx[3 * i], vx[3 * i + 1], vx[3 * i + 2]);
		int found = -1;
		for (size_t k = 0; k < h.verts.size(); ++k) {
			if (equal(h.verts[k], p)) {
				found = (int)k;
				break;
			}
		}
		if (found < 0) {
			found = (int)h.verts.size();
			h.verts.push_back(p);
		}
		remap[i] = found;
	}
	std::vector<Tri> tris;
	for (uint32_t i = 0; i + 2 < nidx; i += 3) tris.push_back(Tri{remap[indices[i]], remap[indices[i + 1]], remap[indices[i + 2]]});

	const int V = (int)h.verts.size(), T = (int)tris.size();
	std::vector<std::vector<int>> v2f(V), v2n(V), tnbr(T);
	for (int i = 0; i < T; ++i) {
		for (int j = 0; j < T; ++j) {
			if (i != j && tris_share_vertex(tris[i], tris[j])) tnbr[i].push_back(j);
		}
		const Tri& t = tris[i];
		if (!contains(v2n[t.x], t.y)) v2n[t.x].push_back(t.y);
		if (!contains(v2n[t.x], t.z)) v2n[t.x].push_back(t.z);
		if (!contains(v2n[t.y], t.x)) v2n[t.y].push_back(t.x);
		if (!contains(v2n[t.y], t.z)) v2n[t.y].push_back(t.z);
		if (!contains(v2n[t.z], t.x)) v2n[t.z].push_back(t.x);
		if (!contains(v2n[t.z], t.y)) v2n[t.z].push_back(t.y);
	}

	std::vector<std::vector<int>> faces;
	std::vector<char> done(T, 0);
	for (int i = 0; i < T; ++i) {
		if (done[i]) continue;
		V3 n = tri_normal(h.verts, tris[i]);
		std::vector<Tri> planar;
		flood_coplanar(h.verts, tris, tnbr, done, i, n, planar);
		int fi = (int)faces.size();
		faces.push_back(face_loop(planar));
		h.normals.push_back(n);
		for (size_t k = 0; k < planar.size(); ++k) {  // duplicates are kept, as in the reference
			v2f[planar[k].x].push_back(fi);
			v2f[planar[k].y].push_back(fi);
			v2f[planar[k].z].push_back(fi);
		}
	}
	const int F = (int)faces.size();
	std::vector<std::vector<int>> f2n(F);
	for (int i = 0; i < F; ++i) {
		for (int j = 0; j < F; ++j) {
			if (i == j) continue;
			bool share = false;
			for (size_t a = 0; a < faces[i].size() && !share; ++a) share = contains(faces[j], faces[i][a]);
			if (share) f2n[i].push_back(j);
		}
	}
	to_csr(faces, h.face_ptr, h.face_idx);
	to_csr(v2f, h.v2f_ptr, h.v2f_idx);
	to_csr(v2n, h.v2n_ptr, h.v2n_idx);
	to_csr(f2n, h.f2n_ptr, h.f2n_idx);
	return h;
}

int Scene::add_hull_collider(const double* vx, uint32_t nverts, const uint32_t* indices, uint32_t nidx) {
	int id = -1;
	for (size_t k = 0; k < hulls.size(); ++k) {
		const HullHost& o = hulls[k];
		if (o.key.size() == 3 * (size_t)nverts && o.key_idx.size() == nidx && memcmp(o.key.data(), vx, sizeof(double) * 3 * nverts) == 0 &&
			memcmp(o.key_idx.data(), indices, sizeof(uint32_t) * nidx) == 0) {
			id = (int)k;
			break;
		}
	}
	if (id < 0) {
		HullHost h;
		if (hull_builder) {
			float ms = 0.f;
			std::string err;
			if (!hull_builder(hull_device, vx, nverts, indices, nidx, &h, &ms, &err)) return -1;
			hull_build_ms += ms;
		} else {
			const auto t0 = std::chrono::steady_clock::now();
			h = build_hull(vx, nverts, indices, nidx);
			hull_build_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		}
		++hulls_built;
		// A zero-area triangle has no normal (NaN after normalisation in the reference, collider.cpp:180-192) and leaves a face
		// that clipping would index out of its row; an empty face likewise. Refused instead of built.
		if (h.face_ptr.size() < 2) return -1;
		for (size_t f = 0; f + 1 < h.face_ptr.size(); ++f) {
			const V3 n = h.normals[f];
			if (h.face_ptr[f + 1] - h.face_ptr[f] < 3 || !(n.x == n.x && n.y == n.y && n.z == n.z) || (n.x == 0.0 && n.y == 0.0 && n.z == 0.0)) return -1;
		}
		id = (int)hulls.size();
		hulls.push_back(h);
	}
	ColliderDesc c;
	c.type = SHAPE_HULL;
	c.hull = id;
	c.radius = 0.0f;
	c.tv0 = c.tn0 = 0;
	c.nv = (int)hulls[id].verts.size();
	c.body = -1;
	pending.push_back(c);
	return (int)pending.size() - 1;
}

int Scene::add_hull_topology(const HullHost& h) {
	int id = -1;
	for (size_t k = 0; k < hulls.size(); ++k) {  // identical topologies share one pool entry
		const HullHost& o = hulls[k];
		if (o.verts.size() == h.verts.size() && o.normals.size() == h.normals.size() &&
			memcmp(o.verts.data(), h.verts.data(), sizeof(V3) * h.verts.size()) == 0 &&
			memcmp(o.normals.data(), h.normals.data(), sizeof(V3) * h.normals.size()) == 0 && o.face_ptr == h.face_ptr &&
			o.face_idx == h.face_idx && o.v2f_ptr == h.v2f_ptr && o.v2f_idx == h.v2f_idx && o.v2n_ptr == h.v2n_ptr && o.v2n_idx == h.v2n_idx &&
			o.f2n_ptr == h.f2n_ptr && o.f2n_idx == h.f2n_idx) {
			id = (int)k;
			break;
		}
	}
	if (id < 0) {
		id = (int)hulls.size();
		hulls.push_back(h);
	}
	ColliderDesc c;
	c.type = SHAPE_HULL;
	c.hull = id;
	c.radius = 0.0f;
	c.tv0 = c.tn0 = 0;
	c.nv = (int)hulls[id].verts.size();
	c.body = -1;
	pending.push_back(c);
	return (int)pending.size() - 1;
}

int Scene::add_sphere_collider(float radius) {  // collider_sphere_create (collider.cpp:12-18)
	ColliderDesc c;
	c.type = SHAPE_SPHERE;
	c.hull = -1;
	c.radius = radius;
	c.tv0 = c.tn0 = 0;
	c.nv = 0;
	c.body = -1;
	pending.push_back(c);
	return (int)pending.size() - 1;
}

static int commit_body(Scene& sc, BodyInit& b);

// entity_create_ex (entity.cpp:24-65)
int Scene::add_body(const double* pos, const double* quat, double mass, int fixed, double mu_s, double mu_d, double rest) {
	BodyInit b;
	b.x = v3(pos[0], pos[1], pos[2]);
	b.q = q4(quat[0], quat[1], quat[2], quat[3]);
	b.col0 = (int)colliders.size();
	b.ncol = (int)pending.size();
	b.fixed = fixed ? 1 : 0;
	b.mu_s = mu_s; b.mu_d = mu_d; b.rest = rest;
	b.mass = mass;
	b.v0 = b.w0 = v3(0.0, 0.0, 0.0);

	// colliders_get_bounding_sphere_radius (collider.cpp:510-521)
	double rmax = -1.7976931348623157e308;
	for (size_t i = 0; i < pending.size(); ++i) {
		double r;
		if (pending[i].type == SHAPE_SPHERE) {
			r = (double)pending[i].radius;
		} else {
			r = 0.0;
			const HullHost& h = hulls[pending[i].hull];
			for (size_t k = 0; k < h.verts.size(); ++k) {
				double d = length(h.verts[k]);
				if (d > r) r = d;
			}
		}
		if (r > rmax) rmax = r;
	}
	b.radius = rmax;

	memset(&b.inertia, 0, sizeof(M3));
	memset(&b.inv_inertia, 0, sizeof(M3));
	if (fixed) {
		b.inv_mass = 0.0;
	} else {
		b.inv_mass = 1.0 / mass;
		// colliders_get_default_inertia_tensor (collider.cpp:448-494), quirk q9: off-diagonals summed with + sign
		if (pending.size() == 1 && pending[0].type == SHAPE_SPHERE) {
			double I = (2.0 / 5.0) * mass * pending[0].radius * pending[0].radius;
			b.inertia.m[0][0] = I; b.inertia.m[1][1] = I; b.inertia.m[2][2] = I;
		} else {
			uint32_t total = 0;
			for (size_t i = 0; i < pending.size(); ++i) {
				if (pending[i].type == SHAPE_HULL) total += (uint32_t)hulls[pending[i].hull].verts.size();
			}
			double mpv = mass / total;
			M3& r = b.inertia;
			for (size_t i = 0; i < pending.size(); ++i) {
				if (pending[i].type != SHAPE_HULL) continue;  // the reference asserts hull here
				const HullHost& h = hulls[pending[i].hull];
				for (size_t k = 0; k < h.verts.size(); ++k) {
					V3 v = h.verts[k];
					r.m[0][0] += mpv * (v.y * v.y + v.z * v.z);
					r.m[0][1] += mpv * v.x * v.y;
					r.m[0][2] += mpv * v.x * v.z;
					r.m[1][0] += mpv * v.x * v.y;
					r.m[1][1] += mpv * (v.x * v.x + v.z * v.z);
					r.m[1][2] += mpv * v.y * v.z;
					r.m[2][0] += mpv * v.x * v.z;
					r.m[2][1] += mpv * v.y * v.z;
					r.m[2][2] += mpv * (v.x * v.x + v.y * v.y);
				}
			}
		}
		// a tensor that cannot be inverted (no hull vertices, all of them on one line ...) is refused; the reference would go on
		// with whatever gm_mat3_inverse left (entity.cpp:45-47 asserts on it)
		if (!inverse(b.inertia, &b.inv_inertia)) {
			pending.clear();
			return -1;
		}
	}
	return commit_body(*this, b);
}

int Scene::add_body_params(const double* pos, const double* quat, double inv_mass, const double* inertia9, const double* inv_inertia9,
	double radius, int fixed, double mu_s, double mu_d, double rest) {
	BodyInit b;
	b.x = v3(pos[0], pos[1], pos[2]);
	b.q = q4(quat[0], quat[1], quat[2], quat[3]);
	b.col0 = (int)colliders.size();
	b.ncol = (int)pending.size();
	b.fixed = fixed ? 1 : 0;
	b.mu_s = mu_s; b.mu_d = mu_d; b.rest = rest;
	b.radius = radius;
	b.inv_mass = inv_mass;
	b.mass = 0.0;
	b.v0 = b.w0 = v3(0.0, 0.0, 0.0);
	for (int r = 0; r < 3; ++r) {
		for (int c = 0; c < 3; ++c) {
			b.inertia.m[r][c] = inertia9[3 * r + c];
			b.inv_inertia.m[r][c] = inv_inertia9[3 * r + c];
		}
	}
	return commit_body(*this, b);
}

// moves the queued colliders under the new body and appends it
static int commit_body(Scene& sc, BodyInit& b) {
	std::vector<ColliderDesc>& pending = sc.pending;
	std::vector<ColliderDesc>& colliders = sc.colliders;
	std::vector<HullHost>& hulls = sc.hulls;
	int& total_tv = sc.total_tv;
	int& total_tn = sc.total_tn;
	std::vector<BodyInit>& bodies = sc.bodies;
	std::vector<V3>& force = sc.force;
	std::vector<V3>& torque = sc.torque;
	for (size_t i = 0; i < pending.size(); ++i) {
		ColliderDesc c = pending[i];
		c.body = (int)bodies.size();
		c.tv0 = total_tv;
		c.tn0 = total_tn;
		if (c.type == SHAPE_HULL) {
			total_tv += (int)hulls[c.hull].verts.size();
			total_tn += (int)hulls[c.hull].normals.size();
		} else {
			total_tv += 1;
		}
		colliders.push_back(c);
	}
	pending.clear();
	bodies.push_back(b);
	force.push_back(v3(0.0, 0.0, 0.0));
	torque.push_back(v3(0.0, 0.0, 0.0));
	return (int)bodies.size() - 1;
}

void Scene::clear_forces() {
	for (size_t i = 0; i < force.size(); ++i) {
		force[i] = v3(0.0, 0.0, 0.0);
		torque[i] = v3(0.0, 0.0, 0.0);
	}
}

// running sums equal to calculate_external_force / calculate_external_torque (physics_util.cpp:5-23) over the
// entity's force list in insertion order
void Scene::add_force(int body, V3 position, V3 f) {
	force[body] = add(force[body], f);
	V3 dist = sub(position, v3(0.0, 0.0, 0.0));
	torque[body] = add(torque[body], cross(dist, f));
}

void Scene::add_gravity(double g) {
	for (size_t i = 0; i < bodies.size(); ++i) add_force((int)i, v3(0.0, 0.0, 0.0), v3(0.0, -g * 1.0 / bodies[i].inv_mass, 0.0));
}

HullPoolHost pool_hulls(const Scene& s) {
	HullPoolHost p;
	for (size_t k = 0; k < s.hulls.size(); ++k) {
		const HullHost& h = s.hulls[k];
		HullTopo t;
		t.nv = (int)h.verts.size();
		t.nf = (int)h.normals.size();
		t.vert0 = (int)p.verts.size();
		t.face0 = (int)p.normals.size();
		t.v2f0 = (int)p.v2f_ptr.size();
		t.v2n0 = (int)p.v2n_ptr.size();
		t.f2n0 = (int)p.f2n_ptr.size();
		t.fptr0 = (int)p.face_ptr.size();
		// *_ptr entries are rebased so they index the pooled *_idx arrays directly
		int fb = (int)p.face_idx.size(), vfb = (int)p.v2f_idx.size(), vnb = (int)p.v2n_idx.size(), fnb = (int)p.f2n_idx.size();
		for (size_t i = 0; i < h.face_ptr.size(); ++i) p.face_ptr.push_back(h.face_ptr[i] + fb);
		for (size_t i = 0; i < h.v2f_ptr.size(); ++i) p.v2f_ptr.push_back(h.v2f_ptr[i] + vfb);
		for (size_t i = 0; i < h.v2n_ptr.size(); ++i) p.v2n_ptr.push_back(h.v2n_ptr[i] + vnb);
		for (size_t i = 0; i < h.f2n_ptr.size(); ++i) p.f2n_ptr.push_back(h.f2n_ptr[i] + fnb);
		p.face_idx.insert(p.face_idx.end(), h.face_idx.begin(), h.face_idx.end());
		p.v2f_idx.insert(p.v2f_idx.end(), h.v2f_idx.begin(), h.v2f_idx.end());
		p.v2n_idx.insert(p.v2n_idx.end(), h.v2n_idx.begin(), h.v2n_idx.end());
		p.f2n_idx.insert(p.f2n_idx.end(), h.f2n_idx.begin(), h.f2n_idx.end());
		p.verts.insert(p.verts.end(), h.verts.begin(), h.verts.end());
		p.normals.insert(p.normals.end(), h.normals.begin(), h.normals.end());
		p.hulls.push_back(t);
	}
	return p;
}

HullPool HullPoolHost::view() const {
	HullPool v;
	v.hulls = hulls.data();
	v.verts = verts.data();
	v.normals = normals.data();
	v.face_ptr = face_ptr.data(); v.face_idx = face_idx.data();
	v.v2f_ptr = v2f_ptr.data(); v.v2f_idx = v2f_idx.data();
	v.v2n_ptr = v2n_ptr.data(); v.v2n_idx = v2n_idx.data();
	v.f2n_ptr = f2n_ptr.data(); v.f2n_idx = f2n_idx.data();
	return v;
}

}  // namespace rp
