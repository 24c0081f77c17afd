// rp_narrow.h -- narrowphase for one collider pair: boolean GJK, EPA, contact-manifold clipping.
//
// Fixed-capacity scratch instead of the reference's growable arrays (SURVEY.md 7 "hard parts" 3); running out of
// capacity, or reaching a state where the reference would abort on an assert, raises a status bit and yields
// "no contact" instead of trapping (failure detection row).
//
// Reference: src/physics/gjk.cpp (gjk_collides :350, do_simplex_2/3/4 :36,:57,:122), src/physics/epa.cpp (epa :118,
// get_face_normal_and_distance_to_origin :32, add_edge :79), src/physics/clipping.cpp (whole file),
// src/physics/collider.cpp:523-558 (collider_get_contacts).
#ifndef RP_NARROW_H
#define RP_NARROW_H

#include "rp_shape.h"

namespace rp {

// status bits (per world on the device, per call on the host)
enum {
	ST_GJK_SIMPLEX_OVERFLOW = 1 << 0,   // gjk.cpp:24-26 assert(0): add_to_simplex on a 4-point simplex (quirk q3)
	ST_EPA_DEGENERATE = 1 << 1,         // epa.cpp:41 / :72 asserts
	ST_EPA_NO_CONVERGENCE = 1 << 2,     // epa.cpp:233 "EPA did not converge"
	ST_EPA_CAPACITY = 1 << 3,           // polytope scratch exhausted (reference arrays are unbounded)
	ST_CLIP_CAPACITY = 1 << 4,          // polygon scratch exhausted
	ST_EDGE_PARALLEL = 1 << 5,          // clipping.cpp:279 assert (skew-line system singular)
	ST_CONTACT_CAPACITY = 1 << 6,       // per-world contact buffer exhausted
	ST_PAIR_CAPACITY = 1 << 7,          // per-world broadphase pair buffer exhausted
	ST_SOLVER_SINGULAR = 1 << 8,        // pbd_base_constraints.cpp:40,:154 assert(w1 + w2 != 0)
	ST_NAN = 1 << 9
};

#define RP_EPA_MAX_VERTS 104   // 4 + 100 iterations (epa.cpp:149)
#define RP_EPA_MAX_FACES 256
#define RP_EPA_MAX_EDGES 192
#define RP_CLIP_MAX_POINTS 160
#define RP_GJK_MAX_ITERS 100
#define RP_EPA_MAX_ITERS 100

struct Simplex {
	V3 a, b, c, d;
	int num;
};

// triple_cross (gjk.cpp:32)
RP_HD V3 cross3(V3 a, V3 b, V3 c) { return cross(cross(a, b), c); }

// Region test shared by do_simplex_3 and the three single-plane cases of do_simplex_4 for the triangle (A, P, Q) with
// normal n = AP x AQ. `aq_alt` is the vector the "AP region reached through the failed AQ test" branch feeds to
// triple_cross: it is AP itself everywhere except gjk.cpp:213 (case 0x2), which passes AB (quirk q2).
// Returns true when the origin projects inside the triangle (caller finishes that region); on the "A region" branches
// only `a` is (re)written and num is left as it was (quirk q3).
RP_HD bool gjk_triangle_edges(Simplex* s, V3* dir, V3 a, V3 P, V3 Q, V3 ap, V3 aq, V3 n, V3 ao, V3 ap_alt) {
	if (dot(cross(n, aq), ao) >= RL(0.0)) {
		if (dot(aq, ao) >= RL(0.0)) {
			s->a = a; s->b = Q; s->num = 2;
			*dir = cross3(aq, ao, aq);
		} else if (dot(ap, ao) >= RL(0.0)) {
			s->a = a; s->b = P; s->num = 2;
			*dir = cross3(ap_alt, ao, ap_alt);
		} else {
			s->a = a;
			*dir = ao;
		}
		return false;
	}
	if (dot(cross(ap, n), ao) >= RL(0.0)) {
		if (dot(ap, ao) >= RL(0.0)) {
			s->a = a; s->b = P; s->num = 2;
			*dir = cross3(ap, ao, ap);
		} else {
			s->a = a;
			*dir = ao;
		}
		return false;
	}
	return true;
}

RP_HD void gjk_line(Simplex* s, V3* dir, V3 a, V3 P, V3 ap, V3 ao) {
	if (dot(ap, ao) >= RL(0.0)) {
		s->a = a; s->b = P; s->num = 2;
		*dir = cross3(ap, ao, ap);
	} else {
		s->a = a; s->num = 1;
		*dir = ao;
	}
}

// do_simplex (gjk.cpp:339-348). Returns true on intersection (tetrahedron encloses the origin).
RP_HD bool gjk_do_simplex(Simplex* s, V3* dir) {
	V3 a = s->a, b = s->b;
	V3 ao = zero_minus(a);
	V3 ab = sub(b, a);
	if (s->num == 2) {  // gjk.cpp:36-55
		gjk_line(s, dir, a, b, ab, ao);
		return false;
	}
	V3 c = s->c;
	V3 ac = sub(c, a);
	V3 abc = cross(ab, ac);
	if (s->num == 3) {  // gjk.cpp:57-120
		if (gjk_triangle_edges(s, dir, a, b, c, ab, ac, abc, ao, ab)) {
			if (dot(abc, ao) >= RL(0.0)) {
				s->a = a; s->b = b; s->c = c; s->num = 3;
				*dir = abc;
			} else {
				s->a = a; s->b = c; s->c = b; s->num = 3;
				*dir = zero_minus(abc);
			}
		}
		return false;
	}
	// gjk.cpp:122-337
	V3 d = s->d;
	V3 ad = sub(d, a);
	V3 acd = cross(ac, ad);
	V3 adb = cross(ad, ab);
	int planes = 0;
	if (dot(abc, ao) >= RL(0.0)) planes |= 1;
	if (dot(acd, ao) >= RL(0.0)) planes |= 2;
	if (dot(adb, ao) >= RL(0.0)) planes |= 4;
	switch (planes) {
		case 0: return true;
		case 1:
			if (gjk_triangle_edges(s, dir, a, b, c, ab, ac, abc, ao, ab)) {
				s->a = a; s->b = b; s->c = c; s->num = 3;
				*dir = abc;
			}
			break;
		case 2:
			if (gjk_triangle_edges(s, dir, a, c, d, ac, ad, acd, ao, ab)) {  // ap_alt = ab: quirk q2
				s->a = a; s->b = c; s->c = d; s->num = 3;
				*dir = acd;
			}
			break;
		case 3: gjk_line(s, dir, a, c, ac, ao); break;
		case 4:
			if (gjk_triangle_edges(s, dir, a, d, b, ad, ab, adb, ao, ad)) {
				s->a = a; s->b = d; s->c = b; s->num = 3;
				*dir = adb;
			}
			break;
		case 5: gjk_line(s, dir, a, b, ab, ao); break;
		case 6: gjk_line(s, dir, a, d, ad, ao); break;
		default:
			s->a = a; s->num = 1;
			*dir = ao;
			break;
	}
	return false;
}

// gjk_collides (gjk.cpp:350-379), cut into its loop-carried pieces so that a GPU lane can interleave the iterations of
// successive pairs (k_gjk refills a lane the moment its pair is decided): gjk_begin is the code before the loop,
// gjk_step one trip of it.
enum { GJK_CONTINUE = 0, GJK_HIT = 1, GJK_MISS = 2 };

template <class SA, class SB>
RP_HD void gjk_begin(const SA& A, const SB& B, Simplex* s, V3* dir) {
	s->a = support_minkowski(A, B, v3(RL(0.0), RL(0.0), RL(1.0)));
	s->b = s->c = s->d = v3(RL(0.0), RL(0.0), RL(0.0));
	s->num = 1;
	*dir = scale(-RL(1.0), s->a);
}

template <class SA, class SB>
RP_HD int gjk_step(const SA& A, const SB& B, Simplex* s, V3* dir, int* status) {
	V3 p = support_minkowski(A, B, *dir);
	if (dot(p, *dir) < RL(0.0)) return GJK_MISS;
	// add_to_simplex (gjk.cpp:7-30)
	if (s->num == 1) {
		s->b = s->a;
	} else if (s->num == 2) {
		s->c = s->b; s->b = s->a;
	} else if (s->num == 3) {
		s->d = s->c; s->c = s->b; s->b = s->a;
	} else {
		*status |= ST_GJK_SIMPLEX_OVERFLOW;  // the reference aborts here
		return GJK_MISS;
	}
	s->a = p;
	++s->num;
	return gjk_do_simplex(s, dir) ? GJK_HIT : GJK_CONTINUE;
}

// `iters` (optional) receives the number of support iterations, for statistics.
template <class SA, class SB>
RP_HD bool gjk(const SA& A, const SB& B, Simplex* out, int* status, int* iters) {
	Simplex s;
	V3 dir;
	gjk_begin(A, B, &s, &dir);
	for (int i = 0; i < RP_GJK_MAX_ITERS; ++i) {
		const int r = gjk_step(A, B, &s, &dir, status);
		if (r != GJK_CONTINUE) {
			if (iters) *iters = i + 1;
			if (r == GJK_HIT) *out = s;
			return r == GJK_HIT;
		}
	}
	if (iters) *iters = RP_GJK_MAX_ITERS;
	return false;
}

// ---------------------------------------------------------------------------------------------------------------- EPA
//
// The polytope lives in a STORE: vertices, face index triples with their planes, and the edge list of the current
// expansion, behind accessors, so the same routine runs on plain arrays (host checker; the device's full-capacity fallback in
// local memory) and on a small thread-interleaved block of shared memory (k_epa: round 1's ncu showed the 11.8 kB
// per-thread arrays missing L1 at every step). A store with SOFT = true reports exhaustion as EPA_OVERFLOW WITHOUT touching
// the status word: the caller reruns the pair on a larger store, and since the routine is deterministic the rerun is the run
// the reference would have made. A store with SOFT = false raises ST_EPA_CAPACITY (the reference's arrays are unbounded).
template <int MAXV_, int MAXF_, int MAXE_, bool SOFT_>
struct EpaArrays {
	enum { MAXV = MAXV_, MAXF = MAXF_, MAXE = MAXE_, SOFT = SOFT_ ? 1 : 0 };
	V3 verts[MAXV];
	V3 normals[MAXF];
	real dists[MAXF];
	uint8_t faces[MAXF][3];
	uint8_t edges[MAXE][2];
	int nverts, nfaces, nedges;
	V3 min_normal;   // closest face so far (first strictly smaller distance wins, epa.cpp:141-144,:222-228)
	real min_dist;
	RP_HD V3 vert(int i) const { return verts[i]; }
	RP_HD void set_vert(int i, V3 v) { verts[i] = v; }
	RP_HD V3 normal(int i) const { return normals[i]; }
	RP_HD real dist(int i) const { return dists[i]; }
	RP_HD void set_plane(int i, V3 n, real d) { normals[i] = n; dists[i] = d; }
	RP_HD void face(int i, int* x, int* y, int* z) const { *x = faces[i][0]; *y = faces[i][1]; *z = faces[i][2]; }
	RP_HD void set_face(int i, int x, int y, int z) { faces[i][0] = (uint8_t)x; faces[i][1] = (uint8_t)y; faces[i][2] = (uint8_t)z; }
	RP_HD void move_face(int dst, int src) {
		faces[dst][0] = faces[src][0]; faces[dst][1] = faces[src][1]; faces[dst][2] = faces[src][2];
		dists[dst] = dists[src];
		normals[dst] = normals[src];
	}
	RP_HD void edge(int i, int* x, int* y) const { *x = edges[i][0]; *y = edges[i][1]; }
	RP_HD void set_edge(int i, int x, int y) { edges[i][0] = (uint8_t)x; edges[i][1] = (uint8_t)y; }
};
typedef EpaArrays<RP_EPA_MAX_VERTS, RP_EPA_MAX_FACES, RP_EPA_MAX_EDGES, false> EpaScratch;
// capacities of the small store (the polytope of a box pair: EPA converges within 3 iterations for 97 % of the W256 pairs,
// SURVEY.md 6; 4 expansions = 8 vertices, 4 + 2 * 4 = 12 faces, horizons of up to 12 edges in flight)
#ifndef RP_EPA_SMALL_VERTS
#define RP_EPA_SMALL_VERTS 8
#define RP_EPA_SMALL_FACES 12
#define RP_EPA_SMALL_EDGES 12
#endif
typedef EpaArrays<RP_EPA_SMALL_VERTS, RP_EPA_SMALL_FACES, RP_EPA_SMALL_EDGES, true> EpaSmallArrays;

// get_face_normal_and_distance_to_origin (epa.cpp:32-77)
template <class E>
RP_HD bool epa_face_plane(const E& e, int ia, int ib, int ic, V3* normal_out, real* dist_out) {
	V3 a = e.vert(ia);
	V3 n = normalize(cross(sub(e.vert(ib), a), sub(e.vert(ic), a)));
	if (!(n.x != RL(0.0) || n.y != RL(0.0) || n.z != RL(0.0))) return false;  // epa.cpp:41
	real dist = dot(n, a);
	if (dist < -RL(0.0)) {
		n = zero_minus(n);
		dist = -dist;
	} else if (dist >= -RL(0.0) && dist <= RL(0.0)) {
		bool found = false;
		for (int i = 0; i < e.nverts; ++i) {
			real t = dot(n, e.vert(i));
			if (t < -RL(0.0) || t > RL(0.0)) {
				n = t < -RL(0.0) ? n : zero_minus(n);
				found = true;
				break;
			}
		}
		if (!found) return false;  // epa.cpp:72
	}
	*normal_out = n;
	*dist_out = dist;
	return true;
}

// add_edge (epa.cpp:79-110): an edge seen twice cancels -- by index in either direction or by coordinate-equal
// endpoints (quirk q8); removal is swap-with-last (light_array.h:146)
template <class E>
RP_HD bool epa_toggle_edge(E& e, int x, int y) {
	for (int i = 0; i < e.nedges; ++i) {
		int cx, cy;
		e.edge(i, &cx, &cy);
		bool hit = (x == cx && y == cy) || (x == cy && y == cx);
		if (!hit) {
			V3 c1 = e.vert(cx), c2 = e.vert(cy), e1 = e.vert(x), e2 = e.vert(y);
			hit = (equal(c1, e1) && equal(c2, e2)) || (equal(c1, e2) && equal(c2, e1));
		}
		if (hit) {
			--e.nedges;
			int lx, ly;
			e.edge(e.nedges, &lx, &ly);
			e.set_edge(i, lx, ly);
			return true;
		}
	}
	if (e.nedges >= E::MAXE) return false;
	e.set_edge(e.nedges, x, y);
	++e.nedges;
	return true;
}

// epa (epa.cpp:118-238), cut into its loop-carried pieces like gjk above: epa_begin is the code before the loop
// (polytope from the GJK tetrahedron, epa.cpp:8-30,:122-146), epa_step one trip of it. The running minimum lives in
// the store.
enum { EPA_CONTINUE = 0, EPA_DONE = 1, EPA_FAIL = 2, EPA_OVERFLOW = 3 };

template <class E>
RP_HD int epa_out_of_room(int* status) {
	if (E::SOFT) return EPA_OVERFLOW;
	*status |= ST_EPA_CAPACITY;
	return EPA_FAIL;
}

template <class E>
RP_HD int epa_begin(const Simplex& s, E& e, int* status) {
	e.set_vert(0, s.a); e.set_vert(1, s.b); e.set_vert(2, s.c); e.set_vert(3, s.d);
	e.nverts = 4;
	e.nfaces = 0;
	e.nedges = 0;
	e.min_normal = v3(RL(0.0), RL(0.0), RL(0.0));
	e.min_dist = RL(RP_REAL_MAX);
	for (int i = 0; i < 4; ++i) {
		// initial faces (0,1,2) (0,2,3) (0,3,1) (1,2,3) (epa.cpp:8-30)
		const int fa = i == 3 ? 1 : 0, fb = i == 0 ? 1 : (i == 3 ? 2 : i + 1), fc = i == 0 ? 2 : (i == 2 ? 1 : 3);
		V3 n; real d;
		if (!epa_face_plane(e, fa, fb, fc, &n, &d)) {
			*status |= ST_EPA_DEGENERATE;
			return EPA_FAIL;
		}
		e.set_face(i, fa, fb, fc);
		e.set_plane(i, n, d);
		e.nfaces = i + 1;
		if (d < e.min_dist) {
			e.min_dist = d;
			e.min_normal = n;
		}
	}
	return EPA_CONTINUE;
}

template <class SA, class SB, class E>
RP_HD int epa_step(const SA& A, const SB& B, E& e, int* status, int* sup_a = 0, int* sup_b = 0) {
	const V3 min_normal = e.min_normal;
	int ia = -1, ib = -1;
	V3 sp = support_minkowski_idx(A, B, min_normal, &ia, &ib);
	if (sup_a) {  // (on EPA_DONE these are the support vertices along the returned normal)
		*sup_a = ia;
		*sup_b = ib;
	}
	real d = dot(min_normal, sp);
	if (fabs(d - e.min_dist) < RL(0.0001)) return EPA_DONE;  // result: e.min_normal, e.min_dist
	if (e.nverts >= E::MAXV) return epa_out_of_room<E>(status);  // (the full store holds 4 + RP_EPA_MAX_ITERS: never reached there)
	int new_index = e.nverts;
	e.set_vert(e.nverts++, sp);

	// faces that see the new point are removed (swap-with-last while scanning, quirk q7), their edges toggled
	int i = 0;
	while (i < e.nfaces) {
		int fx, fy, fz;
		e.face(i, &fx, &fy, &fz);
		V3 centroid = scale(RL(1.0) / RL(3.0), add(add(e.vert(fy), e.vert(fz)), e.vert(fx)));  // triangle_centroid (epa.cpp:112)
		if (dot(e.normal(i), sub(sp, centroid)) > RL(0.0)) {
			if (!epa_toggle_edge(e, fx, fy) || !epa_toggle_edge(e, fy, fz) || !epa_toggle_edge(e, fz, fx)) return epa_out_of_room<E>(status);
			int last = --e.nfaces;
			e.move_face(i, last);
		} else {
			++i;
		}
	}
	for (int k = 0; k < e.nedges; ++k) {
		if (e.nfaces >= E::MAXF) return epa_out_of_room<E>(status);
		int ex, ey;
		e.edge(k, &ex, &ey);
		V3 n; real dd;
		if (!epa_face_plane(e, ex, ey, new_index, &n, &dd)) {
			*status |= ST_EPA_DEGENERATE;
			return EPA_FAIL;
		}
		int f = e.nfaces++;
		e.set_face(f, ex, ey, new_index);
		e.set_plane(f, n, dd);
	}
	e.min_dist = RL(RP_REAL_MAX);
	for (int k = 0; k < e.nfaces; ++k) {
		const real dk = e.dist(k);
		if (dk < e.min_dist) {
			e.min_dist = dk;
			e.min_normal = e.normal(k);
		}
	}
	e.nedges = 0;
	return EPA_CONTINUE;
}

// returns EPA_DONE (normal, depth written), EPA_FAIL (status says why) or, for a SOFT store, EPA_OVERFLOW (nothing written)
template <class SA, class SB, class E>
RP_HD int epa_run(const SA& A, const SB& B, const Simplex& s, E& e, V3* normal_out, real* depth_out, int* status, int* iters,
	int* sup_a = 0, int* sup_b = 0) {
	if (epa_begin(s, e, status) == EPA_FAIL) return EPA_FAIL;
	for (int it = 0; it < RP_EPA_MAX_ITERS; ++it) {
		const int r = epa_step(A, B, e, status, sup_a, sup_b);
		if (r == EPA_DONE) {
			*normal_out = e.min_normal;
			*depth_out = e.min_dist;
			if (iters) *iters = it + 1;
			return EPA_DONE;
		}
		if (r != EPA_CONTINUE) return r;
	}
	*status |= ST_EPA_NO_CONVERGENCE;
	if (iters) *iters = RP_EPA_MAX_ITERS;
	return EPA_FAIL;
}

#if defined(RP_REAL_F32)
// Single precision only. In float, boxes at rest are EXACTLY axis-aligned (rotation noise below 6e-8 rounds away), their
// Minkowski difference is an exact grid, and EPA keeps meeting exactly collinear / coplanar points where the reference asserts
// (epa.cpp:41, :72): the same failure every substep, the pair stays contact-free, the body sinks, and the deep overlap that is
// finally detected throws the stack apart. When EPA gives up, the penetration is taken from the FACE normals instead: the support
// function of A - B, h(n) = n . (support_A(n) - support_B(-n)), is minimised over the faces of A and the reversed faces of B
// (its minimum over ALL directions, which is what EPA looks for, lies on a face of A - B; the edge-edge faces are left out, a
// resting contact is a face contact). Not part of the FP64 product: there such pairs are flagged, as the reference aborts.
template <class SA, class SB>
RP_HD bool sat_face_fallback(const SA& A, const SB& B, V3* normal_out, real* depth_out, int* sup_a, int* sup_b) {
	if (A.type != SHAPE_HULL || B.type != SHAPE_HULL) return false;
	real best = RL(RP_REAL_MAX);
	for (int pass = 0; pass < 2; ++pass) {
		const int nf = pass == 0 ? A.nf : B.nf;
		for (int i = 0; i < nf; ++i) {
			const V3 n = pass == 0 ? fnormal(A, i) : zero_minus(fnormal(B, i));
			int ia, ib;
			const V3 sp = support_minkowski_idx(A, B, n, &ia, &ib);
			const real hgt = dot(n, sp);
			if (hgt < best) {
				best = hgt;
				*normal_out = n;
				if (sup_a) { *sup_a = ia; *sup_b = ib; }
			}
		}
	}
	if (!(best > RL(0.0)) || !(best < RL(RP_REAL_MAX))) return false;  // separated along some face normal (or no faces): no contact
	*depth_out = best;
	return true;
}
#endif

template <class SA, class SB>
RP_HD bool epa(const SA& A, const SB& B, const Simplex& s, EpaScratch& e, V3* normal_out, real* depth_out, int* status,
	int* iters) {
	return epa_run(A, B, s, e, normal_out, depth_out, status, iters) == EPA_DONE;
}

// ----------------------------------------------------------------------------------------------------------- clipping

struct ClipPlane {
	V3 normal, point;
};

#define RP_MAXF(a, b) (((a) > (b)) ? (a) : (b))
#define RP_MINF(a, b) (((a) < (b)) ? (a) : (b))

// is_point_in_plane (clipping.cpp:12-19): the plane offset is rounded to float (quirk q6). `offset` is
// (float)(-dot(normal, point)), hoisted out of the per-vertex calls (same inputs, same value).
RP_HD float clip_offset(const ClipPlane& pl) { return (float)(-dot(pl.normal, pl.point)); }
RP_HD bool clip_inside(const ClipPlane& pl, float offset, V3 p) { return !(dot(p, pl.normal) + (real)offset < RL(0.0)); }

// plane_edge_intersection (clipping.cpp:21-46): ab_p, the plane offset and the edge factor pass through float
RP_HD bool clip_edge(const ClipPlane& pl, float offset, V3 start, V3 end, V3* out) {
	V3 ab = sub(end, start);
	float ab_p = (float)dot(pl.normal, ab);
	if (fabs((real)ab_p) > RL(0.000001)) {
		V3 p_co = scale((real)(-offset), pl.normal);
		float fac = (float)fdiv(-dot(pl.normal, sub(start, p_co)), (real)ab_p);
		fac = (float)RP_MINF(RP_MAXF((real)fac, RL(0.0)), RL(1.0));
		*out = add(start, scale((real)fac, ab));
		return true;
	}
	return false;
}

// The two polygon buffers of Sutherland-Hodgman behind accessors, like the EPA store above: plain arrays of full capacity
// (host checker; the device's fallback in local memory) or a small thread-interleaved block of shared memory (k_manifold: a
// box face clipped against four planes never exceeds 8 points, while round 1's 160-point arrays were 7.7 kB of local memory
// per thread). SOFT = true: running out of room is reported to the caller (who reruns on the full store) and leaves the
// status word alone.
template <int CAP_, bool SOFT_>
struct ClipArrays {
	enum { CAP = CAP_, SOFT = SOFT_ ? 1 : 0 };
	V3 buf[2][CAP];
	RP_HD V3 get(int b, int i) const { return buf[b][i]; }
	RP_HD void set(int b, int i, V3 v) { buf[b][i] = v; }
};
typedef ClipArrays<RP_CLIP_MAX_POINTS, false> ClipScratch;
#define RP_CLIP_SMALL_POINTS 8
typedef ClipArrays<RP_CLIP_SMALL_POINTS, true> ClipSmallArrays;

// One Sutherland-Hodgman pass of buffer `src` against one plane into the other buffer (body of the loop at
// clipping.cpp:63-108). The reference tests every vertex twice (as the end of one edge and the start of the next); the
// verdict is a pure function of the vertex, so it is carried over instead. Returns the number of points written, or -1
// when the store is full (a point that does not fit is never dropped silently).
template <class C>
RP_HD int clip_pass(const ClipPlane& pl, C& cs, int src, int n_in, bool remove_only) {
	const int dst = src ^ 1;
	int n_out = 0;
	const float offset = clip_offset(pl);
	V3 start = cs.get(src, n_in - 1);
	bool s_in = clip_inside(pl, offset, start);
	for (int j = 0; j < n_in; ++j) {
		V3 end = cs.get(src, j);
		bool e_in = clip_inside(pl, offset, end);
		V3 tmp;
		if (remove_only) {
			if (e_in) {
				if (n_out >= C::CAP) return -1;
				cs.set(dst, n_out++, end);
			}
		} else if (s_in && e_in) {
			if (n_out >= C::CAP) return -1;
			cs.set(dst, n_out++, end);
		} else if (s_in && !e_in) {
			if (clip_edge(pl, offset, start, end, &tmp)) {
				if (n_out >= C::CAP) return -1;
				cs.set(dst, n_out++, tmp);
			}
		} else if (!s_in && e_in) {
			if (clip_edge(pl, offset, start, end, &tmp)) {
				if (n_out >= C::CAP) return -1;
				cs.set(dst, n_out++, tmp);
			}
			if (n_out >= C::CAP) return -1;
			cs.set(dst, n_out++, end);
		}
		start = end;
		s_in = e_in;
	}
	return n_out;
}

// get_face_with_most_fitting_normal (clipping.cpp:136-152). vertex_to_faces lists a face once per triangle of it that touches
// the vertex (cube corner: [0, 0, 2, 5, 5]); a repeated entry projects exactly as its first occurrence and can never win
// the strict comparison, so consecutive repeats are skipped (on the device a face normal costs a matrix product and a
// normalisation). `normal_out` receives the chosen face's normal.
template <class S>
RP_HD int clip_best_face(const S& s, int support_idx, V3 normal, V3* normal_out) {
	real best = -RL(RP_REAL_MAX);
	int sel = 0, prev = -1;
	V3 sel_n = v3(RL(0.0), RL(0.0), RL(0.0));
	bool any = false;
	for (int k = s.v2f_ptr[support_idx]; k < s.v2f_ptr[support_idx + 1]; ++k) {
		int f = s.v2f_idx[k];
		if (f == prev) continue;
		prev = f;
		V3 fn = fnormal(s, f);
		real proj = dot(fn, normal);
		if (proj > best) {
			best = proj;
			sel = f;
			sel_n = fn;
			any = true;
		}
	}
	*normal_out = any ? sel_n : fnormal(s, sel);  // (no face beat -DBL_MAX: an empty list or NaNs; the reference returns face 0)
	return sel;
}

// collision_distance_between_skew_lines (clipping.cpp:210-247), quirk q5: the names are swapped but self-consistent
RP_HD bool clip_skew_lines(V3 p1, V3 d1, V3 p2, V3 d2, V3* l1, V3* l2) {
	real n1 = d1.x * d2.x + d1.y * d2.y + d1.z * d2.z;
	real n2 = d2.x * d2.x + d2.y * d2.y + d2.z * d2.z;
	real m1 = -d1.x * d1.x - d1.y * d1.y - d1.z * d1.z;
	real m2 = -d2.x * d1.x - d2.y * d1.y - d2.z * d1.z;
	real r1 = -d1.x * p2.x + d1.x * p1.x - d1.y * p2.y + d1.y * p1.y - d1.z * p2.z + d1.z * p1.z;
	real r2 = -d2.x * p2.x + d2.x * p1.x - d2.y * p2.y + d2.y * p1.y - d2.z * p2.z + d2.z * p1.z;
	if ((n1 * m2) - (n2 * m1) == 0) return false;
	real n = fdiv((r1 * m2) - (r2 * m1), (n1 * m2) - (n2 * m1));
	real m = fdiv((n1 * r2) - (n2 * r1), (n1 * m2) - (n2 * m1));
	*l1 = add(p1, scale(m, d1));
	*l2 = add(p2, scale(n, d2));
	return true;
}

// convex_convex_contact_manifold (clipping.cpp:249-341), in two halves so that a caller can learn how many contacts there
// are before it decides where they go (k_manifold counts, takes a run of the world's contact buffer, then writes the
// records straight from the clip buffer -- no staging copy of the manifold):
//   manifold_clip  everything up to and including the clipping: either the single edge-edge contact, or the candidate
//                  points (incident face clipped against the reference face's side planes, then culled against the
//                  reference plane) left in buffer `cur` of the store;
//   manifold_emit  the final loop (clipping.cpp:322-338): penetration of each candidate along the normal, contact if < 0.
// `sup1_known` / `sup2_known` (>= 0): the support vertices along +-normal when a caller has already found them (the GPU
// finds them with a whole warp for large hulls, k_epa_warp); the scan here would return the same indices.
enum { CLIP_OK = 0, CLIP_OVERFLOW = 1 };
struct ClipResult {
	int kind;          // 0: nothing (failed / empty), 1: one edge-edge contact (l1, l2), 2: `n` candidate points in buffer `cur`
	int n, cur;
	bool ref1;         // the reference face is on hull 1
	V3 l1, l2;         // kind 1
	V3 rp_normal, rp_point;  // reference plane (inverted face normal, first face vertex), kind 2
};

// front half of convex_convex_contact_manifold (clipping.cpp:249-300): support vertices, best-aligned faces, the edge-edge
// test; the result is either the single edge contact or which face is clipped against which
struct FaceChoice {
	int kind;          // 0: nothing (the reference would have aborted), 1: edge-edge contact (l1, l2), 2: clip iface against rface's neighbours
	bool ref1;         // the reference face is on hull 1
	int rface, iface;
	V3 l1, l2;
	V3 ref_normal;     // normal of the reference face
};
template <class S>
RP_HD void manifold_select(const S& h1, const S& h2, V3 normal, int* status, FaceChoice* out, int sup1_known = -1, int sup2_known = -1) {
	out->kind = 0; out->ref1 = false; out->rface = out->iface = 0;
	V3 inv_normal = zero_minus(normal);
	int sup1 = sup1_known >= 0 ? sup1_known : support_index(h1, normal);
	int sup2 = sup2_known >= 0 ? sup2_known : support_index(h2, inv_normal);
	V3 f1n, f2n;
	int face1 = clip_best_face(h1, sup1, normal, &f1n);
	int face2 = clip_best_face(h2, sup2, inv_normal, &f2n);

	real dot1 = dot(f1n, normal);
	real dot2 = dot(f2n, inv_normal);
	const real EPS = RL(0.0001);

	// get_edge_with_most_fitting_normal (clipping.cpp:154-201) picks, over all pairs of edges leaving the two support
	// vertices, the unit cross product best aligned with the collision normal (first maximum wins, both signs tried).
	// Its result only matters if it beats BOTH face alignments by EPS (clipping.cpp:270). Two exact shortcuts:
	//  (1) every candidate is a dot of two vectors normalised by gm_vec3_normalize, so it is <= 1 + 12 ulp; if a face
	//      alignment + EPS already exceeds that bound the edge branch cannot be taken and the search is skipped;
	//  (2) otherwise a float pre-pass scores every candidate (error << 1e-5) and the FP64 evaluation, in the original
	//      order with the original strict comparison, runs only on candidates within 1e-3 of the best score -- the
	//      first exact maximum is always among them, and everything before it is exactly smaller, so the selected
	//      edge pair, the edge normal and the maximum are the ones the full loop would return.
	real best = -RL(RP_REAL_MAX);
	int e1n = 0, e2n = 0;
	V3 edge_normal = v3(RL(0.0), RL(0.0), RL(0.0));
	const bool edge_possible = !(dot1 + EPS > RL(1.000000000001) || dot2 + EPS > RL(1.000000000001));
	if (edge_possible) {
		V3 s1 = vert(h1, sup1), s2 = vert(h2, sup2);
		const float nx = (float)normal.x, ny = (float)normal.y, nz = (float)normal.z;
		float best_score = -1.0f;
		for (int pass = 0; pass < 2; ++pass) {
			for (int i = h1.v2n_ptr[sup1]; i < h1.v2n_ptr[sup1 + 1]; ++i) {
				V3 edge1 = sub(s1, vert(h1, h1.v2n_idx[i]));
				const float ax = (float)edge1.x, ay = (float)edge1.y, az = (float)edge1.z;
				const float a2 = ax * ax + ay * ay + az * az;
				for (int j = h2.v2n_ptr[sup2]; j < h2.v2n_ptr[sup2 + 1]; ++j) {
					V3 edge2 = sub(s2, vert(h2, h2.v2n_idx[j]));
					const float bx = (float)edge2.x, by = (float)edge2.y, bz = (float)edge2.z;
					const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
					const float len2 = cx * cx + cy * cy + cz * cz;
					const float b2 = bx * bx + by * by + bz * bz;
					// nearly parallel edges (sin < 1e-2): the float cross product is cancellation noise, so such a
					// candidate is never scored, always evaluated exactly, and does not set the bar for the others
					const bool degenerate = !(len2 > 1e-4f * a2 * b2);
					float score = 0.0f;
					if (!degenerate) score = fabsf(cx * nx + cy * ny + cz * nz) / sqrtf(len2);
					if (pass == 0) {
						if (!degenerate && score > best_score) best_score = score;
						continue;
					}
					if (!degenerate && score < best_score - 1e-3f) continue;
					V3 cn = normalize(cross(edge1, edge2));
					V3 cni = zero_minus(cn);
					real t = dot(cn, normal);
					if (t > best) {
						best = t; e1n = h1.v2n_idx[i]; e2n = h2.v2n_idx[j]; edge_normal = cn;
					}
					t = dot(cni, normal);
					if (t > best) {
						best = t; e1n = h1.v2n_idx[i]; e2n = h2.v2n_idx[j]; edge_normal = cni;
					}
				}
			}
		}
	}
	real dote = dot(edge_normal, normal);
	if (edge_possible && dote > dot1 + EPS && dote > dot2 + EPS) {
		V3 p1 = vert(h1, sup1);
		V3 d1 = sub(vert(h1, e1n), p1);
		V3 p2 = vert(h2, sup2);
		V3 d2 = sub(vert(h2, e2n), p2);
		if (!clip_skew_lines(p1, d1, p2, d2, &out->l1, &out->l2)) {
			*status |= ST_EDGE_PARALLEL;  // the reference aborts (clipping.cpp:279)
			return;
		}
		out->kind = 1;
		return;
	}
	out->kind = 2;
	out->ref1 = dot1 > dot2;
	out->rface = out->ref1 ? face1 : face2;
	out->iface = out->ref1 ? face2 : face1;
	out->ref_normal = out->ref1 ? f1n : f2n;
}

template <class S, class C>
RP_HD int manifold_clip(const S& h1, const S& h2, V3 normal, C& cs, int* status, ClipResult* out, int sup1_known = -1,
	int sup2_known = -1) {
	out->kind = 0; out->n = 0; out->cur = 0; out->ref1 = false;
	FaceChoice fc;
	manifold_select(h1, h2, normal, status, &fc, sup1_known, sup2_known);
	if (fc.kind == 0) return CLIP_OK;
	if (fc.kind == 1) {
		out->kind = 1; out->l1 = fc.l1; out->l2 = fc.l2;
		return CLIP_OK;
	}
	const bool ref1 = fc.ref1;
	// (copies, not references: a reference chosen at run time keeps BOTH shapes -- two model matrices on the device -- alive
	// through the whole clipping loop; as copies the incident hull is dead once its polygon has been read)
	const S R = ref1 ? h1 : h2;   // reference hull
	const S I = ref1 ? h2 : h1;   // incident hull
	const int rface = fc.rface, iface = fc.iface;

	// incident polygon (get_vertices_of_faces, clipping.cpp:241-247)
	int cur = 0;
	int n = 0;
	for (int k = I.face_ptr[iface]; k < I.face_ptr[iface + 1]; ++k) {
		if (n >= C::CAP) {
			if (C::SOFT) return CLIP_OVERFLOW;
			*status |= ST_CLIP_CAPACITY;
			return CLIP_OK;
		}
		cs.set(0, n++, vert(I, I.face_idx[k]));
	}
	// boundary planes of the reference face (build_boundary_planes, clipping.cpp:121-134), clipped one at a time
	// (sutherland_hodgman, clipping.cpp:52-113)
	for (int k = R.f2n_ptr[rface]; k < R.f2n_ptr[rface + 1]; ++k) {
		if (n == 0) break;
		int nf = R.f2n_idx[k];
		ClipPlane pl;
		pl.point = vert(R, R.face_idx[R.face_ptr[nf]]);
		pl.normal = zero_minus(fnormal(R, nf));
		n = clip_pass(pl, cs, cur, n, false);
		if (n < 0) {
			if (C::SOFT) return CLIP_OVERFLOW;
			*status |= ST_CLIP_CAPACITY;
			return CLIP_OK;
		}
		cur ^= 1;
	}
	out->rp_normal = zero_minus(fc.ref_normal);
	out->rp_point = vert(R, R.face_idx[R.face_ptr[rface]]);
	if (n != 0) {
		ClipPlane rp;
		rp.normal = out->rp_normal;
		rp.point = out->rp_point;
		n = clip_pass(rp, cs, cur, n, true);  // (never grows: cannot overflow)
		cur ^= 1;
	}
	out->kind = 2; out->n = n; out->cur = cur; out->ref1 = ref1;
	return CLIP_OK;
}

// one candidate point of the clipped polygon (clipping.cpp:322-338): its penetration along the normal; a contact if negative
RP_HD bool manifold_point(V3 p, V3 rp_normal, V3 rp_point, bool ref1, V3 normal, V3* p1, V3* p2) {
	// get_closest_point_polygon (clipping.cpp:115-119)
	real dd = dot(scale(-RL(1.0), rp_normal), rp_point);
	V3 closest = sub(p, scale(dot(rp_normal, p) + dd, rp_normal));
	V3 diff = sub(p, closest);
	if (ref1) {
		real pen = dot(diff, normal);
		if (!(pen < RL(0.0))) return false;
		*p1 = sub(p, scale(pen, normal));
		*p2 = p;
	} else {
		real pen = -dot(diff, normal);
		if (!(pen < RL(0.0))) return false;
		*p1 = p;
		*p2 = add(p, scale(pen, normal));
	}
	return true;
}

// Sink: void operator()(V3 p1, V3 p2), called once per contact in the reference's order
template <class C, class Sink>
RP_HD void manifold_emit(const C& cs, const ClipResult& r, V3 normal, Sink& sink) {
	if (r.kind == 1) {
		sink(r.l1, r.l2);
		return;
	}
	if (r.kind != 2) return;
	for (int k = 0; k < r.n; ++k) {
		V3 p1, p2;
		if (manifold_point(cs.get(r.cur, k), r.rp_normal, r.rp_point, r.ref1, normal, &p1, &p2)) sink(p1, p2);
	}
}

template <class Sink>
RP_HD void manifold_hull_hull(const Shape& h1, const Shape& h2, V3 normal, ClipScratch& cs, int* status, Sink& sink, int sup1_known = -1,
	int sup2_known = -1) {
	ClipResult r;
	manifold_clip(h1, h2, normal, cs, status, &r, sup1_known, sup2_known);
	manifold_emit(cs, r, normal, sink);
}

// collider_get_contacts (collider.cpp:523-558) + clipping_get_contact_manifold (clipping.cpp:343-371) for one collider
// pair whose GJK verdict is already known to be "colliding" (hull involved) -- see narrow_pair below for the front half.
template <class Sink>
RP_HD void manifold(const Shape& A, const Shape& B, V3 normal, real depth, ClipScratch& cs, int* status, Sink& sink, int sup1_known = -1,
	int sup2_known = -1) {
	if (A.type == SHAPE_SPHERE) {
		V3 p = support(A, normal);
		sink(p, sub(p, scale(depth, normal)));
	} else if (B.type == SHAPE_SPHERE) {
		V3 p = support(B, zero_minus(normal));
		sink(add(p, scale(depth, normal)), p);
	} else {
		manifold_hull_hull(A, B, normal, cs, status, sink, sup1_known, sup2_known);
	}
}

// Two-tier drivers: the small store first, the full store when it runs out (what the CUDA kernels do with their
// shared-memory stores; the CPU restatement calls these, so the rerun logic is checked against the compiled reference too).
// `reruns` (optional) counts the pairs that needed the second tier.
template <class SA, class SB, class Small, class Full>
RP_HD bool epa_tiered(const SA& A, const SB& B, const Simplex& s, Small& small, Full& full, V3* normal_out, real* depth_out, int* status,
	int* reruns, int* sup_a = 0, int* sup_b = 0) {
	int r = epa_run(A, B, s, small, normal_out, depth_out, status, 0, sup_a, sup_b);
	if (r == EPA_OVERFLOW) {
		if (reruns) ++*reruns;
		r = epa_run(A, B, s, full, normal_out, depth_out, status, 0, sup_a, sup_b);
	}
	return r == EPA_DONE;
}
template <class Small, class Full, class Sink>
RP_HD void manifold_tiered(const Shape& A, const Shape& B, V3 normal, real depth, Small& small, Full& full, int* status, Sink& sink,
	int* reruns, int sup1_known = -1, int sup2_known = -1) {
	if (A.type == SHAPE_SPHERE || B.type == SHAPE_SPHERE) {
		manifold(A, B, normal, depth, full, status, sink);
		return;
	}
	ClipResult r;
	if (manifold_clip(A, B, normal, small, status, &r, sup1_known, sup2_known) == CLIP_OVERFLOW) {
		if (reruns) ++*reruns;
		manifold_clip(A, B, normal, full, status, &r, sup1_known, sup2_known);
		manifold_emit(full, r, normal, sink);
	} else {
		manifold_emit(small, r, normal, sink);
	}
}

// sphere-sphere analytic test (collider.cpp:530-542): squared distance and radius sum are float (quirk q6)
RP_HD bool sphere_sphere(const Shape& A, const Shape& B, V3* normal, real* depth) {
	V3 dv = sub(A.center, B.center);
	float dist2 = (float)dot(dv, dv);
	float min_dist = A.radius + B.radius;
	if (dist2 < (min_dist * min_dist)) {
		*normal = normalize(sub(B.center, A.center));
		*depth = (real)(min_dist - sqrtf(dist2));
		return true;
	}
	return false;
}

}  // namespace rp
#endif
