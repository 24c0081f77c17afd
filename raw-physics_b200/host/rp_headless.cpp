// rp_headless.cpp -- headless driver: what is left of the reference's GLFW loop (src/main.cpp:115-167) when there is no
// window -- build a scene the way an example's init() does, then call the frame step the way its update() does, for any
// number of independent worlds on one GPU, through the C ABI only (include/rawphys_b200.h). Host code next to the
// library: plain C++, no CUDA, no Python.
//
//   rp_headless --list
//   rp_headless --scene NAME [--params a,b,c] [--perturb] [--rows R] [--cols C] [--worlds W] [--frames F] [--dt DT]
//               [--substeps S] [--iters I] [--device D] [--meshes DIR] [--dump FILE] [--dump-every K] [--dump-worlds N] [--no-collisions]
//               [--device-hulls]   (hull topology built on the GPU: rp_example_create_on, csrc/rp_hull.cuh)
//               [--coloured]   (RP_ORDER_COLOURED: graph-coloured sweeps for one large scene, not bit-comparable)
//               [--gpus G]     (the worlds in G contiguous shards, one batch per CUDA device d, d + 1, ...: worlds never interact,
//                               so there is no exchange between them; every frame is enqueued on all devices before any is waited for)
//
// NAME is any built-in scene of the library (rp_example_create: the init() halves of all 14 examples under src/examples plus
// the benchmark worlds; `levers` = hinge_joints --perturb); substeps, iterations and collisions default to what that
// example's update() passes to pbd_simulate. The update() halves are all the same sequence -- gravity force on every
// entity, pbd_simulate, clear forces (stack.cpp:86-104) -- which is what the frame loop below does. Meshes are the
// reference's OBJ files as triangle soups of float positions (what obj_parse returns, obj.cpp:73-81), one `<name>.f32`
// file each (raw little-endian float triples), by default in assets/meshes next to the library.
//
// --dump writes the state of the first N worlds (--dump-worlds, default 1) every K frames (default: every frame): a 32-byte
// header {"RPHD", u32 version (1: one world, 2: several), u32 bodies, u32 records, u32 stride = RP_STATE_STRIDE, u32 every,
// u32 worlds (version 2), u32 reserved} followed by `records` blocks of {u32 frame, u32 0, double state[worlds][bodies][stride]}.
// Frame numbers count completed steps (1 = after the first step). raw-physics_b200/viewer.py turns a dump into an animated GIF
// or an .obj (the counterpart of looking at the reference's window).
#include <algorithm>
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
	std::vector<double> params;
	int rows = 0, cols = 0, worlds = 1, frames = 60, substeps = 0, iters = -1, device = 0, dump_every = 1, dump_worlds = 1, gpus = 1;
	double dt = 1.0 / 60.0;
	bool collisions = true, coloured = false, perturb = false, list = false, device_hulls = false;
};

[[noreturn]] void die(const std::string& what) {
	fprintf(stderr, "rp_headless: %s", what.c_str());
	const char* e = rp_last_error();
	if (e && *e) fprintf(stderr, " (%s)", e);
	fprintf(stderr, "\n");
	exit(1);
}

void parse(int argc, char** argv, Options& o) {
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto next = [&]() -> const char* {
			if (i + 1 >= argc) die("missing value after " + a);
			return argv[++i];
		};
		if (a == "--scene") o.scene = next();
		else if (a == "--list") o.list = true;
		else if (a == "--perturb") o.perturb = true;
		else if (a == "--params") {
			std::string v = next();
			size_t at = 0;
			while (at <= v.size()) {
				const size_t comma = v.find(',', at);
				o.params.push_back(atof(v.substr(at, comma == std::string::npos ? std::string::npos : comma - at).c_str()));
				if (comma == std::string::npos) break;
				at = comma + 1;
			}
		}
		else if (a == "--rows") o.rows = atoi(next());
		else if (a == "--cols") o.cols = atoi(next());
		else if (a == "--worlds") o.worlds = atoi(next());
		else if (a == "--frames") o.frames = atoi(next());
		else if (a == "--substeps") o.substeps = atoi(next());
		else if (a == "--iters") o.iters = atoi(next());
		else if (a == "--device") o.device = atoi(next());
		else if (a == "--gpus") o.gpus = atoi(next());
		else if (a == "--dt") o.dt = atof(next());
		else if (a == "--meshes") o.meshes = next();
		else if (a == "--dump") o.dump = next();
		else if (a == "--dump-every") o.dump_every = atoi(next());
		else if (a == "--dump-worlds") o.dump_worlds = atoi(next());
		else if (a == "--device-hulls") o.device_hulls = true;
		else if (a == "--no-collisions") o.collisions = false;
		else if (a == "--coloured") o.coloured = true;
		else die("unknown argument " + a);
	}
	if (o.worlds < 1 || o.frames < 0 || o.substeps < 0 || o.dump_every < 1 || o.rows < 0 || o.cols < 0 || o.gpus < 1 || o.gpus > o.worlds || o.dump_worlds < 1) die("bad argument value");
}

}  // namespace

int main(int argc, char** argv) {
	Options o;
	parse(argc, argv, o);
	if (o.list) {
		for (int i = 0; i < rp_example_count(); ++i) printf("%s\n", rp_example_name(i));
		return 0;
	}
	if (rp_device_count() < 1) die("no CUDA device: this library has no CPU path");

	if (o.scene == "levers") {  // (round-1 name)
		o.scene = "hinge_joints";
		o.perturb = true;
	}
	if (o.scene == "brick_wall" && o.params.empty()) {
		o.params.push_back(o.rows ? o.rows : 32);
		o.params.push_back(o.cols ? o.cols : 32);
	}
	rp_example_info info;
	rp_scene* scene = rp_example_create_on(o.scene.c_str(), o.params.empty() ? 0 : o.params.data(), (uint32_t)o.params.size(), o.perturb ? 1 : 0,
		o.meshes.empty() ? 0 : o.meshes.c_str(), &info, o.device_hulls ? o.device : -1);
	if (!scene) die(std::string("rp_example_create: ") + rp_example_error());
	int hulls_built = 0;
	double hull_ms = 0.0;
	rp_scene_hull_build_stats(scene, &hulls_built, &hull_ms);
	if (o.substeps == 0) o.substeps = (int)info.substeps;
	if (o.iters < 0) o.iters = (int)info.pos_iters;
	if (!info.collisions) o.collisions = false;

	if (o.device + o.gpus > rp_device_count()) die("not that many CUDA devices");
	rp_batch_cfg cfg;
	rp_batch_cfg_default(&cfg);
	cfg.solve_order = o.coloured ? RP_ORDER_COLOURED : RP_ORDER_REFERENCE;
	// shards: contiguous blocks of worlds, sizes differing by at most one (the partition of raw-physics_b200/multi.py)
	std::vector<rp_batch*> shards((size_t)o.gpus, (rp_batch*)0);
	std::vector<int> shard_worlds((size_t)o.gpus);
	for (int g = 0; g < o.gpus; ++g) {
		shard_worlds[g] = o.worlds / o.gpus + (g < o.worlds % o.gpus ? 1 : 0);
		if (rp_batch_create(scene, (uint32_t)shard_worlds[g], o.device + g, &cfg, &shards[g]) != RP_OK) die("rp_batch_create");
	}
	rp_batch* batch = shards[0];
	const uint32_t nb = rp_batch_num_bodies(batch);
	const uint32_t dump_worlds = (uint32_t)std::min(o.dump_worlds, shard_worlds[0]);
	std::vector<double> state((size_t)dump_worlds * nb * RP_STATE_STRIDE);

	FILE* dump = 0;
	uint32_t records = 0;
	if (!o.dump.empty()) {
		dump = fopen(o.dump.c_str(), "wb");
		if (!dump) die("cannot write " + o.dump);
		const uint32_t head[8] = {0x44485052u /* "RPHD" */, dump_worlds > 1 ? 2u : 1u, nb, 0u, (uint32_t)RP_STATE_STRIDE, (uint32_t)o.dump_every,
			dump_worlds > 1 ? dump_worlds : 0u, 0u};
		fwrite(head, sizeof(head), 1, dump);
	}

	const auto t0 = std::chrono::steady_clock::now();
	for (int f = 1; f <= o.frames; ++f) {
		// an example's update(): gravity on every entity, simulate, clear forces (stack.cpp:93-102) -- on every shard; the step
		// only enqueues the frame's graph on the shard's stream, so the devices run side by side
		for (rp_batch* b : shards) {
			if (rp_batch_clear_forces(b) != RP_OK || rp_batch_add_gravity(b, info.gravity) != RP_OK) die("forces");
			if (rp_batch_step(b, o.dt, (uint32_t)o.substeps, (uint32_t)o.iters, o.collisions ? 1 : 0) != RP_OK) die("rp_batch_step");
		}
		if (dump && (f % o.dump_every == 0 || f == o.frames)) {
			const int rc = rp_batch_download_state(batch, 0, dump_worlds, state.data());
			if (rc != RP_OK && rc != RP_ERR_CAPACITY) die("rp_batch_download_state");
			const uint32_t tag[2] = {(uint32_t)f, 0u};
			fwrite(tag, sizeof(tag), 1, dump);
			fwrite(state.data(), sizeof(double), state.size(), dump);
			++records;
		}
	}
	for (rp_batch* b : shards) {
		const int rc = rp_batch_sync(b);
		if (rc != RP_OK && rc != RP_ERR_CAPACITY) die("rp_batch_sync");  // (capacity: reported through the status bits below)
	}
	const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (dump) {
		fseek(dump, 12, SEEK_SET);
		fwrite(&records, sizeof(records), 1, dump);
		fclose(dump);
	}

	// every world started from the same poses: they must still agree bit for bit -- within a shard and across the devices --
	// and no world may have raised a flag
	int32_t bits = 0;
	int diverged = 0;
	uint64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	std::vector<double> first((size_t)nb * RP_STATE_STRIDE), other(first.size());
	for (int g = 0; g < o.gpus; ++g) {
		std::vector<int32_t> status((size_t)shard_worlds[g]);
		if (rp_batch_get_status(shards[g], status.data()) != RP_OK) die("rp_batch_get_status");
		for (int32_t s : status) bits |= s;
		if (bits) rp_batch_clear_status(shards[g]);
		const int probes[3] = {0, shard_worlds[g] / 2, shard_worlds[g] - 1};
		for (int w : probes) {
			if (rp_batch_download_state(shards[g], (uint32_t)w, 1, g == 0 && w == 0 ? first.data() : other.data()) != RP_OK) die("rp_batch_download_state");
			if (!(g == 0 && w == 0) && memcmp(first.data(), other.data(), first.size() * sizeof(double)) != 0) ++diverged;
		}
		uint64_t c[8];
		if (rp_batch_get_counters(shards[g], c) != RP_OK) die("rp_batch_get_counters");
		for (int k = 0; k < 8; ++k) counters[k] += (k == 5 && g > 0) ? 0 : c[k];  // (frames: counted once)
	}
	const double units = (double)nb * o.worlds * o.substeps * o.frames;
	printf("{\"scene\": \"%s\", \"order\": \"%s\", \"sweep_depth\": %.1f, \"worlds\": %d, \"gpus\": %d, \"bodies\": %u, \"frames\": %d, \"substeps\": %d, \"seconds\": %.6f, "
	       "\"ms_per_frame\": %.4f, \"body_substeps_per_s\": %.6g, \"status_bits\": %d, \"diverged_worlds\": %d, \"pair_tests\": %llu, \"epa_runs\": %llu, "
	       "\"contacts\": %llu, \"hulls_built\": %d, \"hull_build_ms\": %.3f, \"hull_builder\": \"%s\"}\n",
		o.scene.c_str(), o.coloured ? "coloured" : "reference", counters[5] ? (double)counters[4] / (double)counters[5] / o.worlds : 0.0, o.worlds, o.gpus, nb,
		o.frames, o.substeps, seconds, o.frames ? 1e3 * seconds / o.frames : 0.0,
		seconds > 0.0 ? units / seconds : 0.0, (int)bits, diverged, (unsigned long long)counters[0], (unsigned long long)counters[1],
		(unsigned long long)counters[2], hulls_built, hull_ms, o.device_hulls ? "device" : "host");
	for (rp_batch* b : shards) rp_batch_destroy(b);
	rp_scene_destroy(scene);
	return bits || diverged ? 2 : 0;
}
