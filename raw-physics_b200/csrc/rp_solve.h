// rp_solve.h -- per-body integration, XPBD positional / angular primitives, contact, friction and joint solves,
// velocity derivation and the velocity (dynamic friction + restitution) pass. All routines work on register-resident
// Body records; the callers (CUDA kernels, or the sequential CPU checker) decide where bodies live and in which order
// constraints run.
//
// Reference: src/physics/pbd.cpp (integrate :537-577, solves :81-406, contact->constraint :408-424, velocity derive
// :623-643, velocity solve :646-711), src/physics/pbd_base_constraints.cpp (:6-227), src/physics/physics_util.cpp.
#ifndef RP_SOLVE_H
#define RP_SOLVE_H

#include "rp_narrow.h"

namespace rp {

#define RP_PI_F RL(3.14159265358979)  // include/gm.h:8

struct Body {
	V3 x; Q4 q;          // world_position, world_rotation
	V3 v; V3 w;          // linear_velocity, angular_velocity
	V3 px; Q4 pq;        // previous_world_position / rotation
	V3 pv; V3 pw;        // previous_linear_velocity / angular_velocity
	real inv_mass;
	M3 inertia, inv_inertia;  // body-frame tensors (entity.h:32-33)
	const M3* inv_inertia_p;  // device code: where inv_inertia lives in the template (read when needed, see inv_inertia_of)
	real mu_s, mu_d, rest;
	real ii_bound;     // an upper bound of the largest eigenvalue of inv_inertia (its infinity norm; tensor_bound)
	int fixed, active;
};

// The body-frame inverse inertia tensor as the constraint solves read it. On the device the nine doubles are fetched from
// the template (read-only path, L1-resident, the same address for every lane working on bodies of one class) each time a
// world-space tensor is formed, instead of occupying 18 registers per body for the whole life of a manifold: the
// positional sweep needs two bodies, their tensors in both frames and a contact at once and does not fit 255 registers
// otherwise. The host checker reads the copy inside the Body.
#ifdef __CUDA_ARCH__
RP_HD M3 inv_inertia_of(const Body& b) {
	M3 r;
#pragma unroll
	for (int i = 0; i < 3; ++i) {
#pragma unroll
		for (int j = 0; j < 3; ++j) r.m[i][j] = __ldg(&b.inv_inertia_p->m[i][j]);
	}
	return r;
}
#else
RP_HD const M3& inv_inertia_of(const Body& b) { return b.inv_inertia; }
#endif

// infinity norm of a symmetric (or any) 3x3 tensor: >= its spectral radius
RP_HD real tensor_bound(const M3& a) {
	real best = RL(0.0);
	for (int i = 0; i < 3; ++i) {
		real row = fabs(a.m[i][0]) + fabs(a.m[i][1]) + fabs(a.m[i][2]);
		if (row > best) best = row;
	}
	return best;
}

// ------------------------------------------------------------------------------------------------------- integration
// pbd.cpp:537-577 for one body. `force`/`torque` are the sums of calculate_external_force/torque (physics_util.cpp:5-23).
RP_HD void integrate(Body& b, real h, V3 force, V3 torque) {
	b.px = b.x;
	b.pq = b.q;
	if (b.fixed || !b.active) return;
	b.v = add(b.v, scale(h * b.inv_mass, force));
	b.x = add(b.x, scale(h, b.v));
	M3 iinv = world_tensor(b.q, b.inv_inertia);
	M3 iw = world_tensor(b.q, b.inertia);
	b.w = add(b.w, scale(h, mul(iinv, sub(torque, cross(b.w, mul(iw, b.w))))));
#if defined(RP_EXACT_QUATERNIONS)
	// the reference built without USE_QUATERNIONS_LINEARIZED_FORMULAS (pbd.cpp:16, :570-575): rotation by |w| h about w
	const real angle = length(b.w) * h;
	const Q4 change = quat_axis_angle(normalize(b.w), angle);
	b.q = normalize(mul(change, b.q));
#else
	Q4 aux = q4(b.w.x, b.w.y, b.w.z, RL(0.0));
	Q4 dq = mul(aux, b.q);
	b.q.x = b.q.x + h * RL(0.5) * dq.x;
	b.q.y = b.q.y + h * RL(0.5) * dq.y;
	b.q.z = b.q.z + h * RL(0.5) * dq.z;
	b.q.w = b.q.w + h * RL(0.5) * dq.w;
	b.q = normalize(b.q);
#endif
}

// pbd.cpp:623-643 for one body
RP_HD void derive_velocity(Body& b, real h) {
	if (b.fixed || !b.active) return;
	b.pv = b.v;
	b.pw = b.w;
	b.v = scale(RL(1.0) / h, sub(b.x, b.px));
	Q4 dq = mul(b.q, conj(b.pq));
	if (dq.w >= RL(0.0)) b.w = scale(RL(2.0) / h, v3(dq.x, dq.y, dq.z));
	else b.w = scale(-RL(2.0) / h, v3(dq.x, dq.y, dq.z));
}

// ---------------------------------------------------------------------------------------------- positional primitive
struct PosPre {  // Position_Constraint_Preprocessed_Data (pbd_base_constraints.h:5-12)
	V3 r1, r2;
	M3 ii1, ii2;
};

// A fixed body (entity_create_fixed, entity.cpp:24-65 with mass -> inverse_mass 0 and zero tensors; Scene::add_body)
// has inverse mass +0 and an all +0 inverse inertia tensor. For such a body R * 0 * R^T is a matrix of signed zeros,
// every product with it is a signed zero, and its generalised inverse mass 0 + (+-0) is +0 exactly -- so the tensor
// and the w term are not evaluated at all; a fixed body is never written (pbd_base_constraints.cpp:73-103), so its
// tensor has no other use. The CPU restatement runs this same code against the compiled reference.
RP_HD M3 zero_m3() {
	M3 z;
#pragma unroll
	for (int i = 0; i < 3; ++i) {
#pragma unroll
		for (int j = 0; j < 3; ++j) z.m[i][j] = RL(0.0);
	}
	return z;
}
// calculate_positional_constraint_preprocessed_data (pbd_base_constraints.cpp:6-15)
RP_HD PosPre pos_pre(const Body& b1, const Body& b2, V3 r1_lc, V3 r2_lc) {
	PosPre p;
	p.r1 = rotate(b1.q, r1_lc);
	p.r2 = rotate(b2.q, r2_lc);
	p.ii1 = b1.fixed ? zero_m3() : world_tensor(b1.q, inv_inertia_of(b1));
	p.ii2 = b2.fixed ? zero_m3() : world_tensor(b2.q, inv_inertia_of(b2));
	return p;
}
// generalised inverse mass of one body along n at arm r (pbd_base_constraints.cpp:38-39, pbd.cpp:697-698)
RP_HD real inv_mass_along(const Body& b, V3 r, const M3& ii, V3 n) {
	if (b.fixed) return RL(0.0);
	return b.inv_mass + dot(cross(r, n), mul(ii, cross(r, n)));
}

// positional_constraint_get_delta_lambda (pbd_base_constraints.cpp:17-47)
RP_HD real pos_delta_lambda(const PosPre& p, const Body& b1, const Body& b2, real h, real compliance, real lambda, V3 dx,
	int* status) {
	real c = length(dx);
	if (c <= RL(1e-50)) return RL(0.0);
	V3 n = divide(dx, c);
	real w1 = inv_mass_along(b1, p.r1, p.ii1, n);
	real w2 = inv_mass_along(b2, p.r2, p.ii2, n);
	if (!(w1 + w2 != RL(0.0))) *status |= ST_SOLVER_SINGULAR;  // the reference asserts
	real til = fdiv(compliance, h * h);
	return (-c - til * lambda) / (w1 + w2 + til);
}

// quaternion update shared by both apply routines (pbd_base_constraints.cpp:87-103, :189-207): q +/- 0.5 * ((aux,0) (x) q)
RP_HD void apply_rotation(Body& b, V3 aux, real sign_half) {
#if defined(RP_EXACT_QUATERNIONS)
	// the non-linearised branch (pbd_base_constraints.cpp:105-121, :209-225): rotation by +-|aux| about aux
	const real angle = sign_half > RL(0.0) ? length(aux) : -length(aux);
	const Q4 change = quat_axis_angle(normalize(aux), angle);
	b.q = normalize(mul(change, b.q));
	return;
#endif
	Q4 d = mul(q4(aux.x, aux.y, aux.z, RL(0.0)), b.q);
	if (sign_half > RL(0.0)) {
		b.q.x = b.q.x + RL(0.5) * d.x; b.q.y = b.q.y + RL(0.5) * d.y; b.q.z = b.q.z + RL(0.5) * d.z; b.q.w = b.q.w + RL(0.5) * d.w;
	} else {
		b.q.x = b.q.x - RL(0.5) * d.x; b.q.y = b.q.y - RL(0.5) * d.y; b.q.z = b.q.z - RL(0.5) * d.z; b.q.w = b.q.w - RL(0.5) * d.w;
	}
	b.q = normalize(b.q);
}

// positional_constraint_apply (pbd_base_constraints.cpp:50-123)
RP_HD void pos_apply(const PosPre& p, Body& b1, Body& b2, real dl, V3 dx) {
	real c = length(dx);
	if (c <= RL(1e-50)) return;
	V3 n = divide(dx, c);
	V3 imp = scale(dl, n);
	if (!b1.fixed) b1.x = add(b1.x, scale(b1.inv_mass, imp));
	if (!b2.fixed) b2.x = add(b2.x, scale(-b2.inv_mass, imp));
	if (!b1.fixed) apply_rotation(b1, mul(p.ii1, cross(p.r1, imp)), RL(1.0));
	if (!b2.fixed) apply_rotation(b2, mul(p.ii2, cross(p.r2, imp)), -RL(1.0));
}

// ------------------------------------------------------------------------------------------------- angular primitive
struct AngPre {
	M3 ii1, ii2;
};
// calculate_angular_constraint_preprocessed_data (pbd_base_constraints.cpp:125-131)
RP_HD AngPre ang_pre(const Body& b1, const Body& b2) {
	AngPre a;
	a.ii1 = b1.fixed ? zero_m3() : world_tensor(b1.q, inv_inertia_of(b1));
	a.ii2 = b2.fixed ? zero_m3() : world_tensor(b2.q, inv_inertia_of(b2));
	return a;
}
// angular_constraint_get_delta_lambda (pbd_base_constraints.cpp:133-161)
RP_HD real ang_delta_lambda(const AngPre& a, real h, real compliance, real lambda, V3 dq, int* status) {
	real theta = length(dq);
	if (theta <= RL(1e-50)) return RL(0.0);
	V3 n = divide(dq, theta);
	// a fixed body's term n . (0 n) is a signed zero; w1 + w2 is then the other body's term (or +-0, flagged below)
	real w1 = dot(n, mul(a.ii1, n));
	real w2 = dot(n, mul(a.ii2, n));
	if (!(w1 + w2 != RL(0.0))) *status |= ST_SOLVER_SINGULAR;
	real til = fdiv(compliance, h * h);
	return (-theta - til * lambda) / (w1 + w2 + til);
}
// angular_constraint_apply (pbd_base_constraints.cpp:164-227)
RP_HD void ang_apply(const AngPre& a, Body& b1, Body& b2, real dl, V3 dq) {
	real theta = length(dq);
	if (theta <= RL(1e-50)) return;
	V3 n = divide(dq, theta);
	V3 imp = scale(-dl, n);
	if (!b1.fixed) apply_rotation(b1, mul(a.ii1, imp), RL(1.0));
	if (!b2.fixed) apply_rotation(b2, mul(a.ii2, imp), -RL(1.0));
}

// ------------------------------------------------------------------------------------------------------ contact solve
struct Contact {  // Collision_Constraint minus the normal, which is shared by a collider pair's whole manifold
	V3 r1_lc, r2_lc;
	real lambda_n, lambda_t;
};

// clipping_contact_to_collision_constraint (pbd.cpp:408-424)
RP_HD Contact make_contact(const Body& b1, const Body& b2, V3 p1, V3 p2) {
	Contact c;
	c.r1_lc = rotate(conj(b1.q), sub(p1, b1.x));
	c.r2_lc = rotate(conj(b2.q), sub(p2, b2.x));
	c.lambda_n = RL(0.0);
	c.lambda_t = RL(0.0);
	return c;
}

#if defined(RP_COUNT_FRICTION) && !defined(__CUDA_ARCH__)
static long g_friction_skipped = 0, g_friction_taken = 0, g_friction_evaluated = 0;  // diagnostics build of the port only
#endif
// collision_constraint_solve (pbd.cpp:107-154), incl. quirk q1 (static friction reuses the normal correction vector).
// `prev(b1, b2)` is called right before the previous poses (px, pq) are read -- only inside the static-friction branch,
// which is rarely taken -- so a caller may leave them unloaded until then (PrevInBody: they are already in the bodies).
struct PrevInBody {
	RP_HD void operator()(Body&, Body&) const {}
};
template <class PrevPose>
RP_HD void solve_contact(Contact& c, V3 normal, Body& b1, Body& b2, real h, int* status, const PrevPose& prev) {
	PosPre p = pos_pre(b1, b2, c.r1_lc, c.r2_lc);
	V3 p1 = add(b1.x, p.r1);
	V3 p2 = add(b2.x, p.r2);
	real d = dot(sub(p1, p2), normal);
	if (d > RL(0.0)) {
		V3 dx = scale(d, normal);
		real dl = pos_delta_lambda(p, b1, b2, h, RL(0.0), c.lambda_n, dx, status);
		pos_apply(p, b1, b2, dl, dx);
		c.lambda_n += dl;

		// The rest of the routine only acts if lambda_t + dl_t > mu * lambda_n (pbd.cpp:138), with
		// dl_t = -|dx| / (w1' + w2') evaluated at the corrected poses (contacts have zero compliance). By quirk q1 dl_t
		// is of the size of the normal step, so for mu < 1 the test practically always fails -- after a second round of
		// world-space tensors. It is decided without them whenever a bound already settles it: w_i' = 1/m_i +
		// (r_i' x n)^T I_i'^-1 (r_i' x n) <= 1/m_i + |r_i|^2 lmax(I_i^-1) =: W_i (rotations keep |r_i| and the
		// eigenvalues, |n| = 1), hence lambda_t + dl_t <= lambda_t - |dx| / (W_1 + W_2); if that is below mu * lambda_n by
		// more than `margin` -- a millionth of the magnitudes involved, nine orders of magnitude above the rounding
		// error of either side -- the floating-point test cannot come out true, and nothing else in the skipped code
		// has an effect (lambda_t and the bodies are only written inside the branch).
		{
			const real wub = (b1.fixed ? RL(0.0) : b1.inv_mass + dot(c.r1_lc, c.r1_lc) * b1.ii_bound) +
			                   (b2.fixed ? RL(0.0) : b2.inv_mass + dot(c.r2_lc, c.r2_lc) * b2.ii_bound);
			const real mu_ln = ((b1.mu_s + b2.mu_s) / RL(2.0)) * c.lambda_n;
			if (wub > RL(0.0)) {
				const real step = length(dx) / wub;
				const real margin = RL(1e-6) * (fabs(c.lambda_t) + step + fabs(mu_ln));
				if (c.lambda_t - step + margin < mu_ln) {
#if defined(RP_COUNT_FRICTION) && !defined(__CUDA_ARCH__)
					++g_friction_skipped;
#endif
					return;
				}
			}
		}

#if defined(RP_COUNT_FRICTION) && !defined(__CUDA_ARCH__)
		++g_friction_evaluated;
#endif
		p = pos_pre(b1, b2, c.r1_lc, c.r2_lc);
		p1 = add(b1.x, p.r1);
		p2 = add(b2.x, p.r2);
		dl = pos_delta_lambda(p, b1, b2, h, RL(0.0), c.lambda_t, dx, status);
		real mu = (b1.mu_s + b2.mu_s) / RL(2.0);
		real lambda_n = c.lambda_n;
		real lambda_t = c.lambda_t + dl;
		if (lambda_t > mu * lambda_n) {
			prev(b1, b2);
			V3 p1t = add(b1.px, rotate(b1.pq, c.r1_lc));
			V3 p2t = add(b2.px, rotate(b2.pq, c.r2_lc));
			V3 dp = sub(sub(p1, p1t), sub(p2, p2t));
			V3 dpt = sub(dp, scale(dot(dp, normal), normal));
			pos_apply(p, b1, b2, dl, dpt);
			c.lambda_t += dl;
#if defined(RP_COUNT_FRICTION) && !defined(__CUDA_ARCH__)
			++g_friction_taken;
#endif
		}
	}
}

RP_HD void solve_contact(Contact& c, V3 normal, Body& b1, Body& b2, real h, int* status) {
	solve_contact(c, normal, b1, b2, h, status, PrevInBody());
}

// velocity solve for one contact (pbd.cpp:648-711). The world-space inverse inertia tensors of the preprocessed data
// depend on the orientations only, and the velocity pass never writes an orientation: a caller that walks several
// contacts of one body pair computes them once (vel_tensors) and passes them in; the reference recomputes the same
// values per contact (pbd.cpp:651).
RP_HD AngPre vel_tensors(const Body& b1, const Body& b2) { return ang_pre(b1, b2); }

RP_HD void solve_contact_velocity(const Contact& c, V3 n, Body& b1, Body& b2, real h, const AngPre& t) {
	PosPre p;
	p.r1 = rotate(b1.q, c.r1_lc);
	p.r2 = rotate(b2.q, c.r2_lc);
	p.ii1 = t.ii1;
	p.ii2 = t.ii2;
	V3 v = sub(add(b1.v, cross(b1.w, p.r1)), add(b2.v, cross(b2.w, p.r2)));
	real vn = dot(n, v);
	V3 vt = sub(v, scale(vn, n));
	V3 dv = v3(RL(0.0), RL(0.0), RL(0.0));
	real mu = (b1.mu_d + b2.mu_d) / RL(2.0);
	real fn = fdiv(c.lambda_n, h);
	real fact = RP_MINF(mu * fabs(fn), length(vt));
	dv = add(dv, scale(-fact, normalize(vt)));
	real e = b1.rest * b2.rest;
	if (e == RL(0.0)) {
		// -e * vn_til is a signed zero (or NaN), which the reference's MIN(x, 0) turns into +0.0 either way: the previous
		// velocities are not needed at all (callers may leave pv / pw unloaded when either restitution is zero)
		fact = -vn + RL(0.0);
	} else {
		V3 vtil = sub(add(b1.pv, cross(b1.pw, p.r1)), add(b2.pv, cross(b2.pw, p.r2)));
		real vn_til = dot(n, vtil);
		fact = -vn + RP_MINF(-e * vn_til, RL(0.0));
	}
	dv = add(dv, scale(fact, n));
	real w1 = inv_mass_along(b1, p.r1, p.ii1, n);
	real w2 = inv_mass_along(b2, p.r2, p.ii2, n);
	V3 imp = scale(RL(1.0) / (w1 + w2), dv);
	if (!b1.fixed) {
		b1.v = add(b1.v, scale(b1.inv_mass, imp));
		b1.w = add(b1.w, mul(p.ii1, cross(p.r1, imp)));
	}
	if (!b2.fixed) {
		b2.v = add(b2.v, zero_minus(scale(b2.inv_mass, imp)));
		b2.w = add(b2.w, zero_minus(mul(p.ii2, cross(p.r2, imp))));
	}
}
RP_HD void solve_contact_velocity(const Contact& c, V3 n, Body& b1, Body& b2, real h) {
	solve_contact_velocity(c, n, b1, b2, h, vel_tensors(b1, b2));
}

// -------------------------------------------------------------------------------------------------------------- joints
enum { JOINT_POSITIONAL = 0, JOINT_MUTUAL_ORIENTATION = 2, JOINT_HINGE = 3, JOINT_SPHERICAL = 4 };  // Constraint_Type (pbd.h:14-20)

struct Joint {  // the external Constraint union (pbd.h:22-91), flattened
	int type;
	int e1, e2;
	int limited;            // hinge
	int axis[4];            // hinge: e1_aligned, e2_aligned, e1_limit, e2_limit; spherical: e1_swing, e2_swing, e1_twist, e2_twist
	V3 r1_lc, r2_lc;
	V3 distance;            // positional
	real compliance;
	real lower, upper;    // hinge limits / spherical swing limits
	real lower2, upper2;  // spherical twist limits
};

struct JointLambda {  // reset to zero every substep (copy_constraints, pbd.cpp:426-462)
	real a, b, c;
};

// limit_angle (pbd.cpp:175-217)
RP_HD bool limit_angle(V3 n, V3 n1, V3 n2, real alpha, real beta, V3* dq) {
	real phi = asin(dot(cross(n1, n2), n));
	if (dot(n1, n2) < RL(0.0)) phi = RP_PI_F - phi;
	if (phi > RP_PI_F) phi = phi - RL(2.0) * RP_PI_F;
	if (phi < -RP_PI_F) phi = phi + RL(2.0) * RP_PI_F;
	if (phi < alpha || phi > beta) {
		phi = ((phi) > (beta)) ? (beta) : (((phi) < (alpha)) ? (alpha) : (phi));  // CLAMP (common.h:22)
		Q4 rot = quat_axis_angle(n, phi);
		n1 = rotate(rot, n1);
		*dq = cross(n1, n2);
		return true;
	}
	return false;
}

// solve_constraint for the four external types (pbd.cpp:81-97, :156-173, :245-299, :301-379)
RP_HD void solve_joint(const Joint& j, JointLambda& l, Body& b1, Body& b2, real h, int* status) {
	if (j.type == JOINT_POSITIONAL) {
		V3 dx = sub(sub(b1.x, b2.x), j.distance);
		PosPre p = pos_pre(b1, b2, j.r1_lc, j.r2_lc);
		real dl = pos_delta_lambda(p, b1, b2, h, j.compliance, l.a, dx, status);
		pos_apply(p, b1, b2, dl, dx);
		l.a += dl;
	} else if (j.type == JOINT_MUTUAL_ORIENTATION) {
		AngPre a = ang_pre(b1, b2);
		Q4 aux = mul(b1.q, conj(b2.q));
		V3 dq = v3(RL(2.0) * aux.x, RL(2.0) * aux.y, RL(2.0) * aux.z);
		real dl = ang_delta_lambda(a, h, j.compliance, l.a, dq, status);
		ang_apply(a, b1, b2, dl, dq);
		l.a += dl;
	} else if (j.type == JOINT_HINGE) {
		// l.a = lambda_aligned_axes, l.b = lambda_pos, l.c = lambda_limit_axes
		AngPre a = ang_pre(b1, b2);
		V3 a1 = axis_world(b1.q, j.axis[0]);
		V3 a2 = axis_world(b2.q, j.axis[1]);
		V3 dq = cross(a1, a2);
		real dl = ang_delta_lambda(a, h, j.compliance, l.a, dq, status);
		ang_apply(a, b1, b2, dl, dq);
		l.a += dl;

		PosPre p = pos_pre(b1, b2, j.r1_lc, j.r2_lc);
		V3 dx = sub(add(b1.x, p.r1), add(b2.x, p.r2));
		dl = pos_delta_lambda(p, b1, b2, h, RL(0.0), l.b, dx, status);
		pos_apply(p, b1, b2, dl, dx);
		l.b += dl;

		if (j.limited) {
			V3 n1 = axis_world(b1.q, j.axis[2]);
			V3 n2 = axis_world(b2.q, j.axis[3]);
			V3 n = axis_world(b1.q, j.axis[0]);
			if (limit_angle(n, n1, n2, j.lower, j.upper, &dq)) {
				AngPre a2p = ang_pre(b1, b2);
				real dl2 = ang_delta_lambda(a2p, h, RL(0.0), l.c, dq, status);
				ang_apply(a2p, b1, b2, dl2, dq);
				l.c += dl2;
			}
		}
	} else {
		// spherical: l.a = lambda_pos, l.b = lambda_swing, l.c = lambda_twist
		const real EPS = RL(1e-50);
		PosPre p = pos_pre(b1, b2, j.r1_lc, j.r2_lc);
		V3 dx = sub(add(b1.x, p.r1), add(b2.x, p.r2));
		real dl = pos_delta_lambda(p, b1, b2, h, RL(0.0), l.a, dx, status);
		pos_apply(p, b1, b2, dl, dx);
		l.a += dl;

		V3 n1 = axis_world(b1.q, j.axis[0]);
		V3 n2 = axis_world(b2.q, j.axis[1]);
		V3 n = cross(n1, n2);
		real nl = length(n);
		if (nl > EPS) {
			n = divide(n, nl);
			V3 dq;
			if (limit_angle(n, n1, n2, j.lower, j.upper, &dq)) {
				AngPre a = ang_pre(b1, b2);
				real d2 = ang_delta_lambda(a, h, RL(0.0), l.b, dq, status);
				ang_apply(a, b1, b2, d2, dq);
				l.b += d2;
			}
		}
		V3 a1 = axis_world(b1.q, j.axis[0]);
		V3 t1 = axis_world(b1.q, j.axis[2]);
		V3 a2 = axis_world(b2.q, j.axis[1]);
		V3 t2 = axis_world(b2.q, j.axis[3]);
		n = add(a1, a2);
		nl = length(n);
		if (nl > EPS) {
			n = divide(n, nl);
			n1 = sub(t1, scale(dot(n, t1), n));
			n2 = sub(t2, scale(dot(n, t2), n));
			real l1 = length(n1), l2 = length(n2);
			if (l1 > EPS && l2 > EPS) {
				n1 = divide(n1, l1);
				n2 = divide(n2, l2);
				V3 dq;
				if (limit_angle(n, n1, n2, j.lower2, j.upper2, &dq)) {
					AngPre a = ang_pre(b1, b2);
					real d3 = ang_delta_lambda(a, h, RL(0.0), l.c, dq, status);
					ang_apply(a, b1, b2, d3, dq);
					l.c += d3;
				}
			}
		}
	}
}

}  // namespace rp
#endif
