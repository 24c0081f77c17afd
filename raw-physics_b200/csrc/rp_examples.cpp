// rp_examples.cpp -- the scenes of the reference's examples as scene templates: the init() half of every file under
// src/examples (cited per builder) plus the benchmark worlds built from them, behind rp_example_create (rawphys_b200.h).
// Their update() halves are all the same sequence -- gravity force on every entity, pbd_simulate[_with_constraints] with the
// example's (substeps, iterations, collisions), clear forces (stack.cpp:86-104) -- which rp_example_info carries.
//
// One description for every consumer: rp_headless, bench.py and __graft_entry__.smoke() build their scenes here, and
// tests/test_examples.py checks each builder against the test-side description (tests/scenes.py, which also feeds the
// oracle) field by field, so the numbers are the ones the reference's init() computes: OBJ positions are float (obj.cpp:73-81)
// promoted to double, then scaled in double (examples_util.cpp:8-16); angles go through quaternion_new in degrees.
// Meshes are read from `<name>.f32` files (raw float triples of the triangle soup obj_parse returns).
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "../../include/rawphys_b200.h"
#include "rp_scene.h"

using namespace rp;

namespace {

const double PI_F = 3.14159265358979;  // include/gm.h:8

struct Build {
	rp_scene* sc;
	std::string meshes;
	bool perturb;
	std::string err;
};

bool load_soup(Build& b, const char* name, double sx, double sy, double sz, std::vector<double>& v) {
	const std::string path = b.meshes + "/" + name + ".f32";
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) {
		b.err = "cannot open mesh " + path;
		return false;
	}
	std::vector<float> raw;
	float buf[3];
	while (fread(buf, sizeof(float), 3, f) == 3) raw.insert(raw.end(), buf, buf + 3);
	fclose(f);
	if (raw.empty() || raw.size() % 9 != 0) {
		b.err = "mesh " + path + " is not a triangle soup";
		return false;
	}
	v.resize(raw.size());
	for (size_t i = 0; i < raw.size(); i += 3) {  // position.x *= scale.x in double (examples_util.cpp:8-16)
		v[i] = (double)raw[i] * sx;
		v[i + 1] = (double)raw[i + 1] * sy;
		v[i + 2] = (double)raw[i + 2] * sz;
	}
	return true;
}

bool add_hull(Build& b, const std::vector<double>& soup) {
	std::vector<uint32_t> idx(soup.size() / 3);
	for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
	if (b.sc->s.add_hull_collider(soup.data(), (uint32_t)idx.size(), idx.data(), (uint32_t)idx.size()) < 0) {
		b.err = "hull construction failed";
		return false;
	}
	return true;
}
bool add_mesh(Build& b, const char* name, double sx, double sy, double sz) {
	std::vector<double> soup;
	return load_soup(b, name, sx, sy, sz, soup) && add_hull(b, soup);
}

// quaternion_new (src/quaternion.cpp:18-31): axis normalised when non-zero, angle in degrees (gm_radians, PI_F)
Q4 quaternion_new(double ax, double ay, double az, double degrees) {
	const double len = sqrt(ax * ax + ay * ay + az * az);
	if (len != 0.0) {
		ax = ax / len; ay = ay / len; az = az / len;
	}
	const double rad = PI_F * degrees / 180.0;
	const double sn = sin(rad / 2.0);
	return q4(ax * sn, ay * sn, az * sn, cos(rad / 2.0));
}

int add_body(Build& b, V3 p, Q4 q, double mass, bool fixed, double mu_s, double mu_d, double e) {
	const double pos[3] = {p.x, p.y, p.z};
	const double rot[4] = {q.x, q.y, q.z, q.w};
	return b.sc->s.add_body(pos, rot, mass, fixed ? 1 : 0, mu_s, mu_d, e);
}
void spin(Build& b, int body, V3 w) {  // initial angular velocity of the `perturb` variants (scenes at rest without user input)
	if (b.perturb) b.sc->s.bodies[body].w0 = w;
}

// every example's floor: cube.obj scaled (50, 1, 50), fixed at (0, -2, 0), friction 0.5 (stack.cpp:48-51)
bool add_floor(Build& b) {
	return add_mesh(b, "cube", 50.0, 1.0, 50.0) && add_body(b, v3(0.0, -2.0, 0.0), quaternion_new(0, 1, 0, 0.0), 0.0, true, 0.5, 0.5, 0.0) >= 0;
}

// numpy.random.RandomState(seed): MT19937, rand() = 53-bit doubles from two draws, randn() = the legacy polar method
struct NumpyRng {
	std::mt19937 gen;
	bool has_gauss = false;
	double gauss = 0.0;
	explicit NumpyRng(uint32_t seed) : gen(seed) {}
	double rand() {
		const uint32_t a = (uint32_t)gen() >> 5, c = (uint32_t)gen() >> 6;
		return (a * 67108864.0 + c) / 9007199254740992.0;
	}
	double randn() {
		if (has_gauss) {
			has_gauss = false;
			return gauss;
		}
		double f, x1, x2, r2;
		do {
			x1 = 2.0 * rand() - 1.0;
			x2 = 2.0 * rand() - 1.0;
			r2 = x1 * x1 + x2 * x2;
		} while (r2 >= 1.0 || r2 == 0.0);
		f = sqrt(-2.0 * log(r2) / r2);
		gauss = f * x1;
		has_gauss = true;
		return f * x2;
	}
	// generate_random_quaternion (spot_storm.cpp:47-54): axis in U[0,1)^3, angle in U[-180,180), from a seeded generator
	Q4 quaternion() {
		const double x = rand(), y = rand(), z = rand();
		const double ang = -180.0 + 360.0 * rand();
		return quaternion_new(x, y, z, ang);
	}
};

double param(const double* p, uint32_t n, uint32_t i, double dflt) { return p && i < n && p[i] != 0.0 ? p[i] : dflt; }

// src/examples/stack.cpp:35-69: n cubes (scale 1.5, 1, 1; mass 1; friction 0.4) 2.5 apart above the floor
bool ex_stack(Build& b, const double* p, uint32_t n) {
	if (!add_floor(b)) return false;
	std::vector<double> cube;
	if (!load_soup(b, "cube", 1.5, 1.0, 1.0, cube)) return false;
	double y = 0.0;
	const int cubes = (int)param(p, n, 0, 8);
	for (int i = 0; i < cubes; ++i) {
		if (!add_hull(b, cube)) return false;
		add_body(b, v3(0.0, y, 0.0), quaternion_new(0, 1, 0, 0.0), 1.0, false, 0.4, 0.4, 0.0);
		y += 2.5;
	}
	return true;
}

// the north-star world: stacks_x * stacks_z of those stacks on one floor, stack k at x = (k mod sx) * 8 - 28,
// z = (k / sx) * 8 - 12 (32 stacks of 8 = 256 cubes + floor)
bool ex_w256(Build& b, const double* p, uint32_t n) {
	if (!add_floor(b)) return false;
	std::vector<double> cube;
	if (!load_soup(b, "cube", 1.5, 1.0, 1.0, cube)) return false;
	const int sx = (int)param(p, n, 0, 8), sz = (int)param(p, n, 1, 4), height = (int)param(p, n, 2, 8);
	for (int k = 0; k < sx * sz; ++k) {
		const double x = (k % sx) * 8.0 - 28.0, z = (k / sx) * 8.0 - 12.0;
		double y = 0.0;
		for (int i = 0; i < height; ++i) {
			if (!add_hull(b, cube)) return false;
			add_body(b, v3(x, y, z), quaternion_new(0, 1, 0, 0.0), 1.0, false, 0.4, 0.4, 0.0);
			y += 2.5;
		}
	}
	return true;
}

// src/examples/brick_wall.cpp:39-81 with the row / column counts as parameters (32 x 32 = the large-scene config)
bool ex_brick_wall(Build& b, const double* p, uint32_t n) {
	if (!add_floor(b)) return false;
	const double brick_height = 0.35, brick_width = 0.8;
	const double mu_s = (double)0.5f, mu_d = (double)0.4f;  // static r32 in the source (brick_wall.cpp:19-20)
	std::vector<double> brick;
	if (!load_soup(b, "cube", brick_width, brick_height, brick_height, brick)) return false;
	const int rows = (int)param(p, n, 0, 6), cols = (int)param(p, n, 1, 4);
	double y = -1.0;
	for (int i = 0; i < rows; ++i) {
		y += 2 * brick_height + 0.01;
		double x = i % 2 == 0 ? -2.0 : -2.0 + brick_width / 2;
		for (int j = 0; j < cols; ++j) {
			if (!add_hull(b, brick)) return false;
			add_body(b, v3(x, y, 0.0), quaternion_new(0, 1, 0, 0.0), 0.5, false, mu_s, mu_d, 0.0);
			x += 2 * brick_width + 0.01;
		}
	}
	return true;
}

// src/examples/cube_storm.cpp:35-80: n^3 unit cubes, gap 2.01, friction 0.8
bool ex_cube_storm(Build& b, const double* p, uint32_t np) {
	if (!add_floor(b)) return false;
	std::vector<double> cube;
	if (!load_soup(b, "cube", 1.0, 1.0, 1.0, cube)) return false;
	const int n = (int)param(p, np, 0, 3);
	const double gap = 2.01;
	double y = 2.0;
	for (int i = 0; i < n; ++i) {
		y += gap;
		double x = -2.0 * (n / 2.0);
		for (int j = 0; j < n; ++j) {
			x += gap;
			double z = -2.0 * (n / 2.0);
			for (int k = 0; k < n; ++k) {
				z += gap;
				if (!add_hull(b, cube)) return false;
				add_body(b, v3(x, y, z), quaternion_new(0, 1, 0, 0.0), 1.0, false, 0.8, 0.8, 0.0);
			}
		}
	}
	return true;
}

// src/examples/seesaw.cpp:34-76 (contacts only, despite the name)
bool ex_seesaw(Build& b, const double*, uint32_t) {
	if (!add_floor(b)) return false;
	if (!add_mesh(b, "seesaw_support", 2.0, 0.5, 0.25)) return false;
	add_body(b, v3(0.0, (double)-0.2f, 0.0), quaternion_new(0, 1, 0, 90.0), 1.0, false, 0.8, 0.8, 0.0);
	if (!add_mesh(b, "cube", 5.0, 0.03, 1.0)) return false;
	add_body(b, v3(0.0, (double)0.5f, 0.0), quaternion_new(1, 0, 0, 0.0), 1.0, false, 0.8, 0.8, 0.0);
	if (!add_mesh(b, "cube", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(4.0, (double)2.0f, 0.0), quaternion_new(1, 0, 0, 0.0), 0.5, false, 0.8, 0.8, 0.0);
	return true;
}

// src/examples/cube_and_ramp.cpp:36-70: fixed ramp + cube, static friction 1.0 / dynamic 0.7 (static r32)
bool ex_cube_and_ramp(Build& b, const double*, uint32_t) {
	const double mu_s = (double)1.0f, mu_d = (double)0.7f;
	if (!add_mesh(b, "ramp", 2.0, 4.0, 10.0)) return false;
	add_body(b, v3(0.0, -2.0, 0.0), quaternion_new(0, 1, 0, -90.0), 0.0, true, mu_s, mu_d, 0.0);
	if (!add_mesh(b, "cube", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(-5.0, 4.0, 0.0), quaternion_new(1, 0, 0, 0.0), 1.0, false, mu_s, mu_d, 0.0);
	return true;
}

// src/examples/coin.cpp:36-70: 128-vertex cylinder scaled (3, 0.1, 3), tilted 30 degrees, restitution 0.5; floor.obj
bool ex_coin(Build& b, const double*, uint32_t) {
	const double e = (double)0.5f;
	if (!add_mesh(b, "cylinder", 3.0, 0.1, 3.0)) return false;
	add_body(b, v3(0.0, 4.0, 0.0), quaternion_new(1, 0, 1, 30.0), 1.0, false, 0.5, 0.5, e);
	if (!add_mesh(b, "floor", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(0.0, -2.0, 0.0), quaternion_new(0, 1, 0, 0.0), 0.0, true, 0.5, 0.5, e);
	return true;
}

bool add_positional(Build& b, int e1, int e2, V3 r1, V3 r2, double compliance, V3 dist) {
	const double a[3] = {r1.x, r1.y, r1.z}, c[3] = {r2.x, r2.y, r2.z}, d[3] = {dist.x, dist.y, dist.z};
	return rp_scene_add_positional_constraint(b.sc, e1, e2, a, c, compliance, d) >= 0;
}
bool add_hinge(Build& b, int e1, int e2, V3 r1, V3 r2, int a1, int a2, bool limited, int l1, int l2, double lower, double upper) {
	const double a[3] = {r1.x, r1.y, r1.z}, c[3] = {r2.x, r2.y, r2.z};
	return rp_scene_add_hinge_joint_constraint(b.sc, e1, e2, a, c, 0.0, a1, a2, limited ? 1 : 0, l1, l2, lower, upper) >= 0;
}
bool add_spherical(Build& b, int e1, int e2, V3 r1, V3 r2, int s1, int s2, int t1, int t2, double sl, double su, double tl, double tu) {
	const double a[3] = {r1.x, r1.y, r1.z}, c[3] = {r2.x, r2.y, r2.z};
	return rp_scene_add_spherical_joint_constraint(b.sc, e1, e2, a, c, s1, s2, t1, t2, sl, su, tl, tu) >= 0;
}

// src/examples/spring.cpp:33-70: one positional constraint with compliance 0.001 between a fixed anchor and a cube
bool ex_spring(Build& b, const double*, uint32_t) {
	if (!add_floor(b)) return false;
	const Q4 q = quaternion_new(1, 1, 1, 33.0);
	if (!add_mesh(b, "cube", 0.1, 0.1, 0.1)) return false;
	add_body(b, v3(0.0, 6.0, 0.0), q, 0.0, true, 0.5, 0.5, 0.0);
	if (!add_mesh(b, "cube", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(0.0, 2.0, 0.0), q, 1.0, false, 0.8, 0.8, 0.0);
	return add_positional(b, 2, 1, v3(0, 0, 0), v3(0, 0, 0), 0.001, v3(0.0, -3.0, 0.0));
}

// reset_joint_distance (arm.cpp:32-42 and its copies): e2 is moved so that both attachment points coincide
V3 reset_joint(V3 p1, Q4 q1, V3 p2, Q4 q2, V3 r1, V3 r2) {
	const V3 a = add(p1, mul(to_mat3(q1), r1)), c = add(p2, mul(to_mat3(q2), r2));
	return add(p2, sub(a, c));
}

// src/examples/hinge_joints.cpp:32-103: three fixed supports, each carrying a lever on a limited hinge. `perturb`: 2 rad/s about
// every lever's hinge axis (the scene is at rest until the user throws something at it)
bool ex_hinge_joints(Build& b, const double*, uint32_t) {
	struct Spec { V3 pos; Q4 rot; double limit; };
	const Spec specs[3] = {{v3(0.0, 0.0, 0.0), quaternion_new(1.0, 0.0, 0.0, 0.0), 0.9},
	                       {v3(5.0, 0.0, 0.0), quaternion_new(0.0, 0.0, 1.0, 45.0), 0.5},
	                       {v3(-5.0, 0.0, 0.0), quaternion_new(0.0, 0.0, -1.0, 90.0), 0.5}};
	std::vector<double> support, lever;
	if (!load_soup(b, "lever_support", 1.0, 1.0, 1.0, support) || !load_soup(b, "lever", 1.0, 1.0, 1.0, lever)) return false;
	for (const Spec& sp : specs) {
		if (!add_hull(b, support)) return false;
		const int sid = add_body(b, sp.pos, sp.rot, 0.0, true, 0.5, 0.5, 0.0);
		const V3 r1 = v3(0.0, 0.0, 0.0), r2 = v3(0.0, 3.0, 0.0);
		const V3 lever_pos = reset_joint(sp.pos, sp.rot, sp.pos, sp.rot, r1, r2);  // create_lever (hinge_joints.cpp:62-77)
		if (!add_hull(b, lever)) return false;
		const int lid = add_body(b, lever_pos, sp.rot, 1.0, false, 0.6, 0.6, 0.0);
		if (!add_hinge(b, sid, lid, r1, r2, RP_POSITIVE_X_AXIS, RP_POSITIVE_X_AXIS, true, RP_POSITIVE_Y_AXIS, RP_POSITIVE_Y_AXIS, -PI_F * sp.limit,
			PI_F * sp.limit)) return false;
		spin(b, lid, mul(to_mat3(sp.rot), v3(2.0, 0.0, 0.0)));
	}
	return true;
}

// the chains of arm.cpp / triple_pendula.cpp: a fixed support and three links, poses by reset_joint_distance in creation order
struct Chain {
	V3 pos[4];
	V3 scale[4];
	V3 r1[3], r2[3];
};
bool chain_bodies(Build& b, Chain& c, Q4 q) {
	for (int k = 0; k < 3; ++k) c.pos[k + 1] = reset_joint(c.pos[k], q, c.pos[k + 1], q, c.r1[k], c.r2[k]);
	for (int k = 0; k < 4; ++k) {
		if (!add_mesh(b, "cube", c.scale[k].x, c.scale[k].y, c.scale[k].z)) return false;
		if (k == 0) add_body(b, c.pos[k], q, 0.0, true, 0.5, 0.5, 0.0);
		else add_body(b, c.pos[k], q, 1.0, false, 0.6, 0.6, 0.0);
	}
	return true;
}

// src/examples/arm.cpp:44-112: spherical shoulder, limited hinge elbow, spherical wrist with swing 0 and twist limits
bool ex_arm(Build& b, const double*, uint32_t) {
	Chain c = {{v3(0.0, 15.0, 0.0), v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0)},
	           {v3(0.2, 0.1, 0.1), v3(0.2, 1.0, 0.1), v3(0.15, 1.0, 0.1), v3(0.3, 0.3, 0.1)},
	           {v3(0.0, -10.0, 0.0), v3(0.0, -1.2, 0.0), v3(0.0, -1.2, 0.0)},
	           {v3(0.0, 1.0, 0.0), v3(0.0, 1.2, 0.0), v3(0.0, 0.5, 0.0)}};
	if (!chain_bodies(b, c, quaternion_new(1, 0, 0, 0.0))) return false;
	const int X = RP_POSITIVE_X_AXIS, Y = RP_POSITIVE_Y_AXIS;
	if (!add_spherical(b, 0, 1, c.r1[0], c.r2[0], X, X, Y, Y, -PI_F, PI_F, -PI_F, PI_F)) return false;
	if (!add_hinge(b, 1, 2, c.r1[1], c.r2[1], X, X, true, Y, Y, 0.0, 0.9 * PI_F)) return false;
	if (!add_spherical(b, 2, 3, c.r1[2], c.r2[2], X, X, Y, Y, -0.0 * PI_F, 0.0 * PI_F, -0.3 * PI_F, 0.05 * PI_F)) return false;
	spin(b, 1, v3(1.5, 0.3, 0.8));
	spin(b, 2, v3(-2.0, 0.5, 0.2));
	spin(b, 3, v3(0.7, 2.0, -1.0));
	return true;
}

// src/examples/triple_pendula.cpp:45-110: three unlimited hinges in a chain; 50 substeps x 50 iterations, collisions off (:153)
bool ex_triple_pendula(Build& b, const double*, uint32_t) {
	Chain c = {{v3(0.0, 0.0, -2.0), v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0)},
	           {v3(0.1, 0.1, 0.1), v3(0.1, 1.0, 0.1), v3(0.1, 1.0, 0.1), v3(0.1, 0.5, 0.1)},
	           {v3(0.0, 0.0, 2.0), v3(0.0, -1.1, 0.0), v3(0.0, -1.1, 0.0)},
	           {v3(0.0, 1.1, 0.0), v3(0.0, 1.1, 0.0), v3(0.0, 0.55, 0.0)}};
	if (!chain_bodies(b, c, quaternion_new(0, 0, 0, 0.0))) return false;
	const int Z = RP_POSITIVE_Z_AXIS;
	for (int k = 0; k < 3; ++k) {
		if (!add_hinge(b, k, k + 1, c.r1[k], c.r2[k], Z, Z, false, 0, 0, 0.0, 0.0)) return false;
	}
	spin(b, 1, v3(0.0, 0.0, 3.0));
	spin(b, 2, v3(0.0, 0.0, -2.0));
	spin(b, 3, v3(0.0, 0.0, 4.0));
	return true;
}

// src/examples/rott_pendulum.cpp:45-112: support, base, free piece, static piece; two unlimited hinges and one limited to
// [0, 0]; 50 substeps x 50 iterations, collisions off (:154)
bool ex_rott_pendulum(Build& b, const double*, uint32_t) {
	const Q4 q = quaternion_new(0, 0, 0, 0.0);
	const V3 support = v3(0.0, 0.0, -2.0);
	const V3 j0a = v3(0.0, 0.0, 2.0), j0b = v3(0.0, 0.0, 0.0);
	const V3 j1a = v3(0.9, 0.0, 0.0), j1b = v3(0.0, 1.1, -0.25);
	const V3 j2a = v3(-0.9, 0.0, 0.0), j2b = v3(0.0, 1.1, 0.0);
	const V3 base = reset_joint(support, q, v3(0, 0, 0), q, j0a, j0b);
	const V3 free_piece = reset_joint(base, q, v3(0, 0, 0), q, j1a, j1b);
	const V3 static_piece = reset_joint(base, q, v3(0, 0, 0), q, j2a, j2b);
	if (!add_mesh(b, "cube", 0.1, 0.1, 0.1)) return false;
	add_body(b, support, q, 0.0, true, 0.5, 0.5, 0.0);
	if (!add_mesh(b, "cube", 1.0, 0.1, 0.1)) return false;
	add_body(b, base, q, 1.0, false, 0.6, 0.6, 0.0);
	if (!add_mesh(b, "cube", 0.1, 1.0, 0.1)) return false;
	add_body(b, free_piece, q, 1.0, false, 0.6, 0.6, 0.0);
	if (!add_mesh(b, "cube", 0.1, 1.0, 0.1)) return false;
	add_body(b, static_piece, q, 1.0, false, 0.6, 0.6, 0.0);
	const int Z = RP_POSITIVE_Z_AXIS, Y = RP_POSITIVE_Y_AXIS;
	if (!add_hinge(b, 0, 1, j0a, j0b, Z, Z, false, 0, 0, 0.0, 0.0)) return false;
	if (!add_hinge(b, 1, 2, j1a, j1b, Z, Z, false, 0, 0, 0.0, 0.0)) return false;
	if (!add_hinge(b, 1, 3, j2a, j2b, Z, Z, true, Y, Y, 0.0, 0.0)) return false;
	spin(b, 1, v3(0.0, 0.0, 2.5));
	spin(b, 2, v3(0.0, 0.0, -3.0));
	spin(b, 3, v3(0.0, 0.0, 1.0));
	return true;
}

// src/examples/mirror_cube.cpp:45-68: one body carrying two hull colliders (collider.cpp:563-569)
bool ex_mirror_cube(Build& b, const double*, uint32_t) {
	if (!add_floor(b)) return false;
	if (!add_mesh(b, "mirror_cube_collider1", 1.0, 1.0, 1.0) || !add_mesh(b, "mirror_cube_collider2", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(0.0, 2.0, 0.0), quaternion_new(1, 1, 1, 33.0), 1.0, false, 0.8, 0.8, 0.0);
	return true;
}

// src/examples/spot_storm.cpp:35-139: n^3 compound bodies of 11 hulls each (res/spot/spot-hull-1..11.obj, scale 2), gap 3.5;
// orientations as generate_random_quaternion (:47-54) but from a seeded generator (the reference calls unseeded rand())
bool ex_spot_storm(Build& b, const double* p, uint32_t np) {
	if (!add_floor(b)) return false;
	const int n = (int)param(p, np, 0, 2);
	NumpyRng rng((uint32_t)param(p, np, 1, 99));
	std::vector<std::vector<double>> hulls(11);
	for (int i = 0; i < 11; ++i) {
		char name[32];
		snprintf(name, sizeof(name), "spot-hull-%d", i + 1);
		if (!load_soup(b, name, 2.0, 2.0, 2.0, hulls[i])) return false;
	}
	const double gap = 3.5;
	double y = 2.0;
	for (int i = 0; i < n; ++i) {
		y += gap;
		double x = -2.0 * (n / 2.0);
		for (int j = 0; j < n; ++j) {
			x += gap;
			double z = -2.0 * (n / 2.0);
			for (int k = 0; k < n; ++k) {
				z += gap;
				const Q4 q = rng.quaternion();
				for (int h = 0; h < 11; ++h) {
					if (!add_hull(b, hulls[h])) return false;
				}
				add_body(b, v3(x, y, z), q, 1.0, false, 0.8, 0.8, 0.0);
			}
		}
	}
	return true;
}

// src/examples/debug.cpp:47-66: floor.obj as the floor and nothing else (the user spawns bodies)
bool ex_debug(Build& b, const double*, uint32_t) {
	if (!add_mesh(b, "floor", 1.0, 1.0, 1.0)) return false;
	add_body(b, v3(0.0, -2.0, 0.0), quaternion_new(0, 1, 0, 0.0), 0.0, true, 0.5, 0.5, 0.0);
	return true;
}

// BASELINE config 4: a lattice of n_side^3 bodies alternating ico hull / cylinder hull scaled (1, 0.5, 1) / analytic sphere r = 1
// above the floor, seeded random orientations; params: n_side, seed, spacing, layers (default n_side)
bool ex_pile(Build& b, const double* p, uint32_t np) {
	if (!add_floor(b)) return false;
	const int n = (int)param(p, np, 0, 4);
	NumpyRng rng((uint32_t)param(p, np, 1, 12345));
	const double spacing = param(p, np, 2, 3.0);
	const int layers = (int)param(p, np, 3, n);
	std::vector<double> ico, cyl;
	if (!load_soup(b, "ico", 1.0, 1.0, 1.0, ico) || !load_soup(b, "cylinder", 1.0, 0.5, 1.0, cyl)) return false;
	int kind = 0;
	for (int i = 0; i < layers; ++i) {
		for (int j = 0; j < n; ++j) {
			for (int l = 0; l < n; ++l) {
				const V3 pos = v3((j - n / 2.0) * spacing, 1.5 + i * spacing, (l - n / 2.0) * spacing);
				const Q4 q = rng.quaternion();
				if (kind % 3 == 0) {
					if (!add_hull(b, ico)) return false;
				} else if (kind % 3 == 1) {
					if (!add_hull(b, cyl)) return false;
				} else {
					b.sc->s.add_sphere_collider(1.0f);
				}
				++kind;
				add_body(b, pos, q, 1.0, false, 0.8, 0.8, 0.0);
			}
		}
	}
	return true;
}

// randomly oriented unit cubes dropped close together (edge / vertex contacts, EPA expansion); params: n, seed
bool ex_tumble(Build& b, const double* p, uint32_t np) {
	if (!add_floor(b)) return false;
	const int n = (int)param(p, np, 0, 6);
	NumpyRng rng((uint32_t)param(p, np, 1, 7));
	std::vector<double> cube;
	if (!load_soup(b, "cube", 1.0, 1.0, 1.0, cube)) return false;
	for (int i = 0; i < n; ++i) {
		const Q4 q = rng.quaternion();
		const double px = 0.3 * rng.randn();
		const double pz = 0.3 * rng.randn();
		if (!add_hull(b, cube)) return false;
		add_body(b, v3(px, 0.6 + 1.9 * i, pz), q, 1.0, false, 0.6, 0.5, 0.2);
	}
	return true;
}

// analytic spheres dropped on the floor and on each other (collider.cpp:530-542, clipping.cpp:348-364)
bool ex_spheres(Build& b, const double* p, uint32_t np) {
	if (!add_floor(b)) return false;
	const int n = (int)param(p, np, 0, 5);
	for (int i = 0; i < n; ++i) {
		b.sc->s.add_sphere_collider(1.0f);
		add_body(b, v3(0.15 * i, 0.2 + 2.05 * i, -0.1 * i), q4(0.0, 0.0, 0.0, 1.0), 1.0, false, 0.7, 0.6, 0.3);
	}
	return true;
}

struct Entry {
	const char* name;
	bool (*build)(Build&, const double*, uint32_t);
	uint32_t substeps, iters;
	int collisions;
};
const Entry EXAMPLES[] = {
	{"stack", ex_stack, 20, 1, 1}, {"brick_wall", ex_brick_wall, 20, 1, 1}, {"hinge_joints", ex_hinge_joints, 20, 1, 1},
	{"seesaw", ex_seesaw, 20, 1, 1}, {"cube_storm", ex_cube_storm, 20, 1, 1}, {"spot_storm", ex_spot_storm, 1, 1, 1},
	{"arm", ex_arm, 20, 1, 1}, {"spring", ex_spring, 20, 1, 1}, {"coin", ex_coin, 20, 1, 1}, {"mirror_cube", ex_mirror_cube, 20, 1, 1},
	{"cube_and_ramp", ex_cube_and_ramp, 20, 1, 1}, {"rott_pendulum", ex_rott_pendulum, 50, 50, 0},
	{"triple_pendula", ex_triple_pendula, 50, 50, 0}, {"debug", ex_debug, 20, 1, 1},
	// not examples of the reference: the benchmark worlds and test scenes built from the same pieces
	{"w256", ex_w256, 20, 1, 1}, {"pile", ex_pile, 20, 1, 1}, {"tumble", ex_tumble, 20, 1, 1}, {"spheres", ex_spheres, 20, 1, 1},
};
const int N_EXAMPLES = (int)(sizeof(EXAMPLES) / sizeof(EXAMPLES[0]));

std::string default_mesh_dir() {
	Dl_info info;
	if (dladdr((const void*)&default_mesh_dir, &info) && info.dli_fname) {
		std::string lib = info.dli_fname;
		const size_t slash = lib.find_last_of('/');
		return (slash == std::string::npos ? std::string(".") : lib.substr(0, slash)) + "/assets/meshes";
	}
	return "assets/meshes";
}

thread_local std::string g_example_err;

}  // namespace

extern "C" {

int rp_example_count(void) { return N_EXAMPLES; }
const char* rp_example_name(int index) { return index >= 0 && index < N_EXAMPLES ? EXAMPLES[index].name : 0; }
const char* rp_example_error(void) { return g_example_err.c_str(); }

rp_scene* rp_example_create(const char* name, const double* params, uint32_t n_params, int perturb, const char* mesh_dir, rp_example_info* info) {
	return rp_example_create_on(name, params, n_params, perturb, mesh_dir, info, -1);
}

rp_scene* rp_example_create_on(const char* name, const double* params, uint32_t n_params, int perturb, const char* mesh_dir, rp_example_info* info,
	int hull_cuda_device) {
	g_example_err.clear();
	if (!name) {
		g_example_err = "rp_example_create: no name";
		return 0;
	}
	for (int i = 0; i < N_EXAMPLES; ++i) {
		if (strcmp(EXAMPLES[i].name, name) != 0) continue;
		Build b;
		b.sc = rp_scene_create();
		if (hull_cuda_device >= 0 && rp_scene_set_hull_device(b.sc, hull_cuda_device) != RP_OK) {
			g_example_err = std::string("rp_example_create_on: ") + rp_last_error();
			rp_scene_destroy(b.sc);
			return 0;
		}
		b.meshes = mesh_dir && *mesh_dir ? mesh_dir : default_mesh_dir();
		b.perturb = perturb != 0;
		if (!EXAMPLES[i].build(b, params, n_params)) {
			g_example_err = b.err.empty() ? std::string("building example ") + name + " failed" : b.err;
			rp_scene_destroy(b.sc);
			return 0;
		}
		if (info) {
			// spot_storm.cpp:184 steps with ONE substep; everything else as listed in the table
			info->substeps = EXAMPLES[i].substeps;
			info->pos_iters = EXAMPLES[i].iters;
			info->collisions = EXAMPLES[i].collisions;
			info->gravity = 10.0;
		}
		return b.sc;
	}
	g_example_err = std::string("unknown example ") + name;
	return 0;
}

}  // extern "C"
