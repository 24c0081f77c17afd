// rp_shape.h -- collider instances: CSR hull topology shared by every world, per-pose transform, support mapping.
//
// Reference: src/physics/collider.h:6-44 (types), collider.cpp:409-445 (collider_update), util.cpp:60-74
// (util_get_model_matrix_no_scale), support.cpp:5-39 (support mapping).
#ifndef RP_SHAPE_H
#define RP_SHAPE_H

#include "rp_math.h"

namespace rp {

enum { SHAPE_SPHERE = 0, SHAPE_HULL = 1 };  // Collider_Type (collider.h:31-34)

// One convex hull's topology in the reference's own index order (collider_convex_hull_create, collider.cpp:194-364):
// CSR arrays; v2f keeps the reference's duplicate entries, v2n includes triangulation diagonals, f2n = faces sharing
// any vertex. Offsets index the pooled arrays of HullPool.
struct HullTopo {
	int nv, nf;
	int vert0;                // first local vertex in HullPool::verts
	int face0;                // first face in HullPool::normals / face_ptr
	int fptr0;                // first entry of this hull in HullPool::face_ptr (nf + 1 entries, rebased to the pooled face_idx)
	int v2f0, v2n0, f2n0;     // first entries of this hull in the other *_ptr arrays (count + 1 entries each, rebased likewise)
};

struct HullPool {
	const HullTopo* hulls;
	const V3* verts;          // local vertices
	const V3* normals;        // local face normals
	const int* face_ptr; const int* face_idx;
	const int* v2f_ptr; const int* v2f_idx;
	const int* v2n_ptr; const int* v2n_idx;
	const int* f2n_ptr; const int* f2n_idx;
};

// A collider of the scene template (shared by all worlds): which hull / sphere radius, and where its transformed
// vertices and normals live inside one world's transformed-geometry block.
struct ColliderDesc {
	int type;
	int hull;                 // index into HullPool::hulls (SHAPE_HULL)
	float radius;             // r32, as in Collider_Sphere (collider.h:26-29)
	int tv0;                  // first transformed vertex (and, for a sphere, the slot holding its centre)
	int tn0;                  // first transformed normal
	int nv;                   // vertices of the hull (0 for a sphere): decides which narrowphase kernel takes a pair
	int body;                 // the body that carries this collider
};

// A collider at a pose, as the narrowphase sees it.
struct Shape {
	int type;
	float radius;
	V3 center;                // sphere centre (collider.cpp:433)
	// transformed vertices / (re-normalised) face normals, addressed as base[i * stride + c * cstride] so the same code
	// reads the AoS arrays in global memory (stride 3, cstride 1) and thread-interleaved copies staged in shared memory
	const real* vp; int vs, vcs;
	const real* np; int ns, ncs;
	int nv, nf;
	const int* face_ptr; const int* face_idx;
	const int* v2f_ptr; const int* v2f_idx;
	const int* v2n_ptr; const int* v2n_idx;
	const int* f2n_ptr; const int* f2n_idx;
};

RP_HD Shape make_shape(const HullPool& pool, const ColliderDesc& c, const V3* world_tv, const V3* world_tn) {
	Shape s;
	s.type = c.type;
	s.radius = c.radius;
	s.vp = (const real*)(world_tv + c.tv0); s.vs = 3; s.vcs = 1;
	s.np = (const real*)(world_tn + c.tn0); s.ns = 3; s.ncs = 1;
	if (c.type == SHAPE_SPHERE) {
		s.center = world_tv[c.tv0];
		s.nv = 0; s.nf = 0;
		s.face_ptr = s.face_idx = s.v2f_ptr = s.v2f_idx = s.v2n_ptr = s.v2n_idx = s.f2n_ptr = s.f2n_idx = 0;
	} else {
		const HullTopo h = pool.hulls[c.hull];
		s.center = v3(RL(0.0), RL(0.0), RL(0.0));
		s.nv = h.nv; s.nf = h.nf;
		s.face_ptr = pool.face_ptr + h.fptr0; s.face_idx = pool.face_idx;
		s.v2f_ptr = pool.v2f_ptr + h.v2f0; s.v2f_idx = pool.v2f_idx;
		s.v2n_ptr = pool.v2n_ptr + h.v2n0; s.v2n_idx = pool.v2n_idx;
		s.f2n_ptr = pool.f2n_ptr + h.f2n0; s.f2n_idx = pool.f2n_idx;
	}
	return s;
}

RP_HD V3 vert(const Shape& s, int i) {
	const real* p = s.vp + (size_t)i * s.vs;
	return v3(p[0], p[s.vcs], p[2 * s.vcs]);
}
// A shape whose vertices were staged into a thread-interleaved block of NT columns (k_gjk / k_epa): the strides are
// compile-time constants, so an unrolled support scan addresses the vertices with immediate offsets instead of
// recomputing 64-bit addresses per vertex (ncu, round 1: address arithmetic and loads were half of k_gjk's
// instructions). The routines that only need vertices are templates over the shape type for this.
template <int NT>
struct StagedShape : Shape {
	RP_HD StagedShape() {}
	RP_HD explicit StagedShape(const Shape& s) : Shape(s) {}
};
template <int NT>
RP_HD V3 vert(const StagedShape<NT>& s, int i) {
	const real* p = s.vp + i * (3 * NT);
	return v3(p[0], p[NT], p[2 * NT]);
}
RP_HD V3 fnormal_stored(const Shape& s, int i) {
	const real* p = s.np + (size_t)i * s.ns;
	return v3(p[0], p[s.ncs], p[2 * s.ncs]);
}
RP_HD V3 fnormal(const Shape& s, int i) {
	const real* p = s.np + (size_t)i * s.ns;
	return v3(p[0], p[s.ncs], p[2 * s.ncs]);
}

// Rows 0..2 of the model matrix T * R exactly as gm_mat4_multiply(translation, rotation) forms them
// (util.cpp:60-74 with quaternion_get_matrix, quaternion.cpp:107): every entry keeps its four-term sum, including the
// products with the matrices' structural zeros, so signed zeros come out as in the reference.
struct Pose34 {
	real m[3][4];
};

RP_HD Pose34 model_matrix(Q4 q, V3 t) {
	M3 R = to_mat3(q);
	real T[3][4] = {{RL(1.0), RL(0.0), RL(0.0), t.x}, {RL(0.0), RL(1.0), RL(0.0), t.y}, {RL(0.0), RL(0.0), RL(1.0), t.z}};
	Pose34 M;
#pragma unroll
	for (int i = 0; i < 3; ++i) {
#pragma unroll
		for (int j = 0; j < 3; ++j) M.m[i][j] = T[i][0] * R.m[0][j] + T[i][1] * R.m[1][j] + T[i][2] * R.m[2][j] + T[i][3] * RL(0.0);
		M.m[i][3] = T[i][0] * RL(0.0) + T[i][1] * RL(0.0) + T[i][2] * RL(0.0) + T[i][3] * RL(1.0);
	}
	return M;
}

// collider_update, vertex half (collider.cpp:414-422): gm_mat4_multiply_vec4(M, (v,1)); the following scale by
// 1/w is the identity because row 3 of M is (+0,+0,+0,1) and so w == 1 exactly.
RP_HD V3 transform_point(const Pose34& M, V3 v) {
	return v3(v.x * M.m[0][0] + v.y * M.m[0][1] + v.z * M.m[0][2] + RL(1.0) * M.m[0][3],
	          v.x * M.m[1][0] + v.y * M.m[1][1] + v.z * M.m[1][2] + RL(1.0) * M.m[1][3],
	          v.x * M.m[2][0] + v.y * M.m[2][1] + v.z * M.m[2][2] + RL(1.0) * M.m[2][3]);
}
// collider_update, normal half (collider.cpp:425-429): rotate then RE-NORMALISE (not a bitwise no-op)
RP_HD V3 transform_normal(const Pose34& M, V3 n) {
	V3 r = v3(M.m[0][0] * n.x + M.m[0][1] * n.y + M.m[0][2] * n.z,
	          M.m[1][0] * n.x + M.m[1][1] * n.y + M.m[1][2] * n.z,
	          M.m[2][0] * n.x + M.m[2][1] * n.y + M.m[2][2] * n.z);
	return normalize(r);
}

// A collider whose transformed geometry is NOT stored: every vertex (and face normal) is evaluated from the body's model
// matrix and the hull's local data at the moment it is asked for -- the same expressions collider_update evaluates
// (collider.cpp:414-429), so the same bits. The CUDA narrowphase uses it for everything: a world's poses (56 B per body)
// stay in L2 where the transformed hulls (336 B per cube, re-read by GJK, EPA and clipping: round 1's kernels stalled on
// exactly those loads, ncu: a third of all stall samples of k_epa and k_manifold) did not, and the FP64 pipe has the room.
struct PoseShape : Shape {
	Pose34 M;
	const V3* lv;   // local vertices of the hull (template, shared by all worlds)
	const V3* ln;   // local face normals
};
RP_HD V3 ld_v3(const V3* p) {
#ifdef __CUDA_ARCH__
	return v3(__ldg(&p->x), __ldg(&p->y), __ldg(&p->z));
#else
	return *p;
#endif
}
RP_HD V3 vert(const PoseShape& s, int i) { return transform_point(s.M, ld_v3(s.lv + i)); }
RP_HD V3 fnormal(const PoseShape& s, int i) {
#if defined(RP_STORED_NORMALS)
	return fnormal_stored(s, i);
#else
	return transform_normal(s.M, ld_v3(s.ln + i));
#endif
}

// support_point_get_index (support.cpp:5-17): first maximum wins (strict >), starting from -DBL_MAX
template <class S>
RP_HD int support_index(const S& s, V3 d) {
	int best = 0;
	real best_dot = -RL(RP_REAL_MAX);
#pragma unroll 4
	for (int i = 0; i < s.nv; ++i) {
		real t = dot(vert(s, i), d);
		if (t > best_dot) {
			best = i;
			best_dot = t;
		}
	}
	return best;
}
// support_point (support.cpp:19-32)
template <class S>
RP_HD V3 support(const S& s, V3 d) {
	if (s.type == SHAPE_HULL) return vert(s, support_index(s, d));
	return add(s.center, scale((real)s.radius, normalize(d)));
}
// support_point_of_minkowski_difference (support.cpp:34-39)
template <class SA, class SB>
RP_HD V3 support_minkowski(const SA& a, const SB& b, V3 d) {
	V3 s1 = support(a, d);
	V3 s2 = support(b, scale(-RL(1.0), d));
	return sub(s1, s2);
}
// the same, also telling which hull vertices were chosen (-1 for a sphere): EPA's last support call is made along the normal it
// returns, and convex_convex_contact_manifold starts from the support vertices along +-normal (clipping.cpp:255-256). Its
// second direction is (0,0,0) - n where this one is -1.0 * n: the two differ at most in the sign of a zero component, which
// changes no comparison of the scan (+0 == -0), so the indices are the ones the manifold's own scans would find.
template <class SA, class SB>
RP_HD V3 support_minkowski_idx(const SA& a, const SB& b, V3 d, int* ia, int* ib) {
	V3 s1, s2;
	if (a.type == SHAPE_HULL) {
		*ia = support_index(a, d);
		s1 = vert(a, *ia);
	} else {
		*ia = -1;
		s1 = support(a, d);
	}
	const V3 nd = scale(-RL(1.0), d);
	if (b.type == SHAPE_HULL) {
		*ib = support_index(b, nd);
		s2 = vert(b, *ib);
	} else {
		*ib = -1;
		s2 = support(b, nd);
	}
	return sub(s1, s2);
}

}  // namespace rp
#endif
