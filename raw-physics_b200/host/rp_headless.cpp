// rp_headless.cpp -- headless driver: what is left of the reference's GLFW loop (src/main.cpp:115-167) when there is no
// window -- build a scene the way an example's init() does, then call the frame step the way its update() does, for any
// number of independent worlds on one GPU, through the C ABI only (include/rawphys_b200.h). Host code next to the
// library: plain C++, no CUDA, no Python.
//
//   rp_headless --scene stack|w256|brick_wall|levers [--rows R] [--cols C] [--worlds W] [--frames F] [--dt DT]
//               [--substeps S] [--iters I] [--device D] [--meshes DIR] [--dump FILE] [--dump-every K] [--no-collisions]
//               [--coloured]   (RP_ORDER_COLOURED: graph-coloured sweeps for one large scene, not bit-comparable)
//
// Scenes restate the init() halves of the reference's examples (cited per builder); their update() halves are all the
// same sequence -- gravity force on every entity, pbd_simulate, clear forces (stack.cpp:86-104) -- which is what run()
// does per frame. Meshes are the reference's OBJ files as triangle soups of float positions (what obj_parse returns,
// obj.cpp:73-81), one `<name>.f32` file each (raw little-endian float triples).
//
// --dump writes the state of world 0 every K frames (default: every frame): a 32-byte header {"RPHD", u32 version = 1,
// u32 bodies, u32 records, u32 stride = RP_STATE_STRIDE, u32 every, u64 reserved} followed by `records` blocks of
// {u32 frame, u32 0, double state[bodies][stride]}. Frame numbers count completed steps (1 = after the first step).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rawphys_b200.h"
#include "../csrc/rp_math.h"

namespace {

struct Options {
	std::string scene = "stack", meshes, dump;
	int rows = 32, cols = 32, worlds = 1, frames = 60, substeps = 20, iters = 1, device = 0, dump_every = 1;
	double dt = 1.0 / 60.0;
	bool collisions = true, coloured = false;
};

[[noreturn]] void die(const std::string& what) {
	fprintf(stderr, "rp_headless: %s", what.c_str());
	const char* e = rp_last_error();
	if (e && *e) fprintf(stderr, " (%s)", e);
	fprintf(stderr, "\n");
	exit(1);
}

std::vector<double> load_soup(const Options& o, const char* name, double sx, double sy, double sz) {
	const std::string path = o.meshes + "/" + name + ".f32";
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) die("cannot open mesh " + path);
	std::vector<float> raw;
	float buf[3];
	while (fread(buf, sizeof(float), 3, f) == 3) raw.insert(raw.end(), buf, buf + 3);
	fclose(f);
	if (raw.empty() || raw.size() % 9 != 0) die("mesh " + path + " is not a triangle soup");
	// float positions promoted to double, then scaled in double (examples_util.cpp:8-16)
	std::vector<double> v(raw.size());
	for (size_t i = 0; i < raw.size(); i += 3) {
		v[i] = (double)raw[i] * sx;
		v[i + 1] = (double)raw[i + 1] * sy;
		v[i + 2] = (double)raw[i + 2] * sz;
	}
	return v;
}

void add_hull(rp_scene* s, const std::vector<double>& soup) {
	std::vector<uint32_t> idx(soup.size() / 3);
	for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
	if (rp_scene_collider_hull(s, soup.data(), (uint32_t)idx.size(), idx.data(), (uint32_t)idx.size()) < 0) die("rp_scene_collider_hull");
}

// quaternion_new (src/quaternion.cpp:18-31): axis normalised when non-zero, angle in degrees (gm_radians, PI_F)
rp::Q4 quaternion_new(double ax, double ay, double az, double degrees) {
	const double len = sqrt(ax * ax + ay * ay + az * az);
	if (len != 0.0) {
		ax = ax / len; ay = ay / len; az = az / len;
	}
	const double rad = 3.14159265358979 * degrees / 180.0;
	const double sn = sin(rad / 2.0);
	return rp::q4(ax * sn, ay * sn, az * sn, cos(rad / 2.0));
}

int add_body(rp_scene* s, rp::V3 p, rp::Q4 q, double mass, bool fixed, double mu_s, double mu_d, double e) {
	const double pos[3] = {p.x, p.y, p.z};
	const double rot[4] = {q.x, q.y, q.z, q.w};
	const int id = rp_scene_add_body(s, pos, rot, mass, fixed ? 1 : 0, mu_s, mu_d, e);
	if (id < 0) die("rp_scene_add_body");
	return id;
}

// every example's floor: cube.obj scaled (50, 1, 50), fixed at (0, -2, 0), friction 0.5 (stack.cpp:48-51)
void add_floor(const Options& o, rp_scene* s) {
	add_hull(s, load_soup(o, "cube", 50.0, 1.0, 50.0));
	add_body(s, rp::v3(0.0, -2.0, 0.0), quaternion_new(0, 1, 0, 0.0), 0.0, true, 0.5, 0.5, 0.0);
}

// src/examples/stack.cpp:35-69: eight cubes (scale 1.5, 1, 1; mass 1; friction 0.4) 2.5 apart above the floor
void scene_stack(const Options& o, rp_scene* s) {
	add_floor(o, s);
	const std::vector<double> cube = load_soup(o, "cube", 1.5, 1.0, 1.0);
	double y = 0.0;
	for (int i = 0; i < 8; ++i) {
		add_hull(s, cube);
		add_body(s, rp::v3(0.0, y, 0.0), quaternion_new(0, 1, 0, 0.0), 1.0, false, 0.4, 0.4, 0.0);
		y += 2.5;
	}
}

// the north-star world: 32 of those stacks on one floor, stack k at x = (k mod 8) * 8 - 28, z = (k / 8) * 8 - 12
void scene_w256(const Options& o, rp_scene* s) {
	add_floor(o, s);
	const std::vector<double> cube = load_soup(o, "cube", 1.5, 1.0, 1.0);
	for (int k = 0; k < 32; ++k) {
		const double x = (k % 8) * 8.0 - 28.0, z = (k / 8) * 8.0 - 12.0;
		double y = 0.0;
		for (int i = 0; i < 8; ++i) {
			add_hull(s, cube);
			add_body(s, rp::v3(x, y, z), quaternion_new(0, 1, 0, 0.0), 1.0, false, 0.4, 0.4, 0.0);
			y += 2.5;
		}
	}
}

// src/examples/brick_wall.cpp:39-81 with the row / column counts as parameters (32 x 32 = the large-scene config)
void scene_brick_wall(const Options& o, rp_scene* s) {
	add_floor(o, s);
	const double brick_height = 0.35, brick_width = 0.8;
	const double mu_s = (double)0.5f, mu_d = (double)0.4f;  // static r32 in the source (brick_wall.cpp:19-20)
	const std::vector<double> brick = load_soup(o, "cube", brick_width, brick_height, brick_height);
	double y = -1.0;
	for (int i = 0; i < o.rows; ++i) {
		y += 2 * brick_height + 0.01;
		double x = i % 2 == 0 ? -2.0 : -2.0 + brick_width / 2;
		for (int j = 0; j < o.cols; ++j) {
			add_hull(s, brick);
			add_body(s, rp::v3(x, y, 0.0), quaternion_new(0, 1, 0, 0.0), 0.5, false, mu_s, mu_d, 0.0);
			x += 2 * brick_width + 0.01;
		}
	}
}

// src/examples/hinge_joints.cpp:32-103: three fixed supports, each carrying a lever on a limited hinge. The levers get an
// angular velocity of 2 rad/s about their hinge axis (`spin`), because the scene is at rest until the user throws
// something at it.
void scene_levers(const Options& o, rp_scene* s, std::vector<std::pair<int, rp::V3>>& spin) {
	struct Spec { rp::V3 pos; rp::Q4 rot; double limit; };
	const Spec specs[3] = {{rp::v3(0.0, 0.0, 0.0), quaternion_new(1.0, 0.0, 0.0, 0.0), 0.9},
	                       {rp::v3(5.0, 0.0, 0.0), quaternion_new(0.0, 0.0, 1.0, 45.0), 0.5},
	                       {rp::v3(-5.0, 0.0, 0.0), quaternion_new(0.0, 0.0, -1.0, 90.0), 0.5}};
	const std::vector<double> support = load_soup(o, "lever_support", 1.0, 1.0, 1.0), lever = load_soup(o, "lever", 1.0, 1.0, 1.0);
	for (const Spec& sp : specs) {
		add_hull(s, support);
		const int sid = add_body(s, sp.pos, sp.rot, 0.0, true, 0.5, 0.5, 0.0);
		// create_lever (hinge_joints.cpp:62-77): move the lever so that both anchor points coincide
		const rp::M3 R = rp::to_mat3(sp.rot);
		const rp::V3 r1 = rp::v3(0.0, 0.0, 0.0), r2 = rp::v3(0.0, 3.0, 0.0);
		const rp::V3 p1 = rp::add(sp.pos, rp::mul(R, r1)), p2 = rp::add(sp.pos, rp::mul(R, r2));
		const rp::V3 lever_pos = rp::add(sp.pos, rp::sub(p1, p2));
		add_hull(s, lever);
		const int lid = add_body(s, lever_pos, sp.rot, 1.0, false, 0.6, 0.6, 0.0);
		const double a1[3] = {r1.x, r1.y, r1.z}, a2[3] = {r2.x, r2.y, r2.z};
		if (rp_scene_add_hinge_joint_constraint(s, sid, lid, a1, a2, 0.0, RP_POSITIVE_X_AXIS, RP_POSITIVE_X_AXIS, 1, RP_POSITIVE_Y_AXIS,
			RP_POSITIVE_Y_AXIS, -3.14159265358979 * sp.limit, 3.14159265358979 * sp.limit) < 0) die("rp_scene_add_hinge_joint_constraint");
		spin.push_back(std::make_pair(lid, rp::mul(R, rp::v3(2.0, 0.0, 0.0))));
	}
}

void parse(int argc, char** argv, Options& o) {
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto next = [&]() -> const char* {
			if (i + 1 >= argc) die("missing value after " + a);
			return argv[++i];
		};
		if (a == "--scene") o.scene = next();
		else if (a == "--rows") o.rows = atoi(next());
		else if (a == "--cols") o.cols = atoi(next());
		else if (a == "--worlds") o.worlds = atoi(next());
		else if (a == "--frames") o.frames = atoi(next());
		else if (a == "--substeps") o.substeps = atoi(next());
		else if (a == "--iters") o.iters = atoi(next());
		else if (a == "--device") o.device = atoi(next());
		else if (a == "--dt") o.dt = atof(next());
		else if (a == "--meshes") o.meshes = next();
		else if (a == "--dump") o.dump = next();
		else if (a == "--dump-every") o.dump_every = atoi(next());
		else if (a == "--no-collisions") o.collisions = false;
		else if (a == "--coloured") o.coloured = true;
		else die("unknown argument " + a);
	}
	if (o.worlds < 1 || o.frames < 0 || o.substeps < 1 || o.iters < 0 || o.dump_every < 1 || o.rows < 1 || o.cols < 1) die("bad argument value");
}

std::string default_mesh_dir(const char* argv0) {
	std::string exe = argv0;
	const size_t slash = exe.find_last_of('/');
	const std::string dir = slash == std::string::npos ? "." : exe.substr(0, slash);
	return dir + "/assets/meshes";
}

}  // namespace

int main(int argc, char** argv) {
	Options o;
	o.meshes = default_mesh_dir(argv[0]);
	parse(argc, argv, o);
	if (rp_device_count() < 1) die("no CUDA device: this library has no CPU path");

	rp_scene* scene = rp_scene_create();
	std::vector<std::pair<int, rp::V3>> spin;
	if (o.scene == "stack") scene_stack(o, scene);
	else if (o.scene == "w256") scene_w256(o, scene);
	else if (o.scene == "brick_wall") scene_brick_wall(o, scene);
	else if (o.scene == "levers") scene_levers(o, scene, spin);
	else die("unknown scene " + o.scene);

	rp_batch* batch = 0;
	rp_batch_cfg cfg;
	rp_batch_cfg_default(&cfg);
	cfg.solve_order = o.coloured ? RP_ORDER_COLOURED : RP_ORDER_REFERENCE;
	if (rp_batch_create(scene, (uint32_t)o.worlds, o.device, &cfg, &batch) != RP_OK) die("rp_batch_create");
	const uint32_t nb = rp_batch_num_bodies(batch);
	std::vector<double> state((size_t)nb * RP_STATE_STRIDE);
	if (!spin.empty()) {  // initial angular velocities: edit world 0's records and hand them to every world
		if (rp_batch_download_state(batch, 0, 1, state.data()) != RP_OK) die("rp_batch_download_state");
		for (const auto& sp : spin) {
			double* r = &state[(size_t)sp.first * RP_STATE_STRIDE];
			r[10] = sp.second.x; r[11] = sp.second.y; r[12] = sp.second.z;
		}
		if (rp_batch_broadcast_state(batch, state.data()) != RP_OK) die("rp_batch_broadcast_state");
	}

	FILE* dump = 0;
	uint32_t records = 0;
	if (!o.dump.empty()) {
		dump = fopen(o.dump.c_str(), "wb");
		if (!dump) die("cannot write " + o.dump);
		const uint32_t head[6] = {0x44485052u /* "RPHD" */, 1u, nb, 0u, (uint32_t)RP_STATE_STRIDE, (uint32_t)o.dump_every};
		const uint64_t reserved = 0;
		fwrite(head, sizeof(head), 1, dump);
		fwrite(&reserved, sizeof(reserved), 1, dump);
	}

	const auto t0 = std::chrono::steady_clock::now();
	for (int f = 1; f <= o.frames; ++f) {
		// an example's update(): gravity on every entity, simulate, clear forces (stack.cpp:93-102)
		if (rp_batch_clear_forces(batch) != RP_OK || rp_batch_add_gravity(batch, 10.0) != RP_OK) die("forces");
		if (rp_batch_step(batch, o.dt, (uint32_t)o.substeps, (uint32_t)o.iters, o.collisions ? 1 : 0) != RP_OK) die("rp_batch_step");
		if (dump && (f % o.dump_every == 0 || f == o.frames)) {
			if (rp_batch_download_state(batch, 0, 1, state.data()) != RP_OK) die("rp_batch_download_state");
			const uint32_t tag[2] = {(uint32_t)f, 0u};
			fwrite(tag, sizeof(tag), 1, dump);
			fwrite(state.data(), sizeof(double), state.size(), dump);
			++records;
		}
	}
	if (rp_batch_sync(batch) != RP_OK) die("rp_batch_sync");
	const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (dump) {
		fseek(dump, 12, SEEK_SET);
		fwrite(&records, sizeof(records), 1, dump);
		fclose(dump);
	}

	// every world started from the same poses: they must still agree bit for bit, and no world may have raised a flag
	std::vector<int32_t> status((size_t)o.worlds);
	if (rp_batch_get_status(batch, status.data()) != RP_OK) die("rp_batch_get_status");
	int32_t bits = 0;
	for (int32_t s : status) bits |= s;
	int diverged = 0;
	if (o.worlds > 1) {
		std::vector<double> first((size_t)nb * RP_STATE_STRIDE), other(first.size());
		if (rp_batch_download_state(batch, 0, 1, first.data()) != RP_OK) die("rp_batch_download_state");
		const int probes[3] = {1, o.worlds / 2, o.worlds - 1};
		for (int w : probes) {
			if (rp_batch_download_state(batch, (uint32_t)w, 1, other.data()) != RP_OK) die("rp_batch_download_state");
			if (memcmp(first.data(), other.data(), first.size() * sizeof(double)) != 0) ++diverged;
		}
	}
	uint64_t counters[8];
	if (rp_batch_get_counters(batch, counters) != RP_OK) die("rp_batch_get_counters");
	const double units = (double)nb * o.worlds * o.substeps * o.frames;
	printf("{\"scene\": \"%s\", \"order\": \"%s\", \"sweep_depth\": %.1f, \"worlds\": %d, \"bodies\": %u, \"frames\": %d, \"substeps\": %d, \"seconds\": %.6f, \"ms_per_frame\": %.4f, "
	       "\"body_substeps_per_s\": %.6g, \"status_bits\": %d, \"diverged_worlds\": %d, \"pair_tests\": %llu, \"epa_runs\": %llu, "
	       "\"contacts\": %llu}\n",
		o.scene.c_str(), o.coloured ? "coloured" : "reference", counters[5] ? (double)counters[4] / (double)counters[5] / o.worlds : 0.0, o.worlds, nb,
		o.frames, o.substeps, seconds, o.frames ? 1e3 * seconds / o.frames : 0.0,
		seconds > 0.0 ? units / seconds : 0.0, (int)bits, diverged, (unsigned long long)counters[0], (unsigned long long)counters[1],
		(unsigned long long)counters[2]);
	rp_batch_destroy(batch);
	rp_scene_destroy(scene);
	return bits || diverged ? 2 : 0;
}
