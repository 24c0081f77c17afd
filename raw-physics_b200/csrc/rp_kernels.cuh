// rp_kernels.cuh -- CUDA kernels of the frame step (sm_100a). Included once, by rp_batch.cu.
//
// Arithmetic is FP64 and is the shared core (rp_math.h ... rp_solve.h) compiled with --fmad=false; the kernels only
// decide WHO computes WHAT and WHEN. Order sensitivity of the reference's sequential Gauss-Seidel is preserved by the
// dependency-level schedule built in k_schedule: two constraints commute exactly when they share no non-fixed body
// (fixed bodies are never written, pbd_base_constraints.cpp:73-103), so running each level's units in parallel and the
// levels in sequence reproduces the sequential result bit for bit.
#ifndef RP_KERNELS_CUH
#define RP_KERNELS_CUH

#include "rp_device.cuh"

// occupancy knobs (min resident CTAs per SM handed to __launch_bounds__); tuned on B200, see profiles/
#ifndef RP_MINB_INTEGRATE
#define RP_MINB_INTEGRATE 4
#endif
#ifndef RP_MINB_GJK
#define RP_MINB_GJK 8
#endif
#ifndef RP_MINB_MANIFOLD
#define RP_MINB_MANIFOLD 4
#endif
#ifndef RP_MINB_POS
#define RP_MINB_POS 2
#endif
#ifndef RP_MINB_VEL
#define RP_MINB_VEL 2
#endif

#define RP_GJK_THREADS 64
#define RP_GJK_STAGE 48          // doubles of shared memory per thread: two hulls of up to 16 vertices in total
#define RP_MANIFOLD_THREADS 128
#define RP_MANIFOLD_STAGE 0      // doubles of shared memory per thread for staged hulls in k_manifold (0 = off: measured slower)

#define RP_LVL_SMEM 64     // levels ranked through shared memory in k_manifold
#define RP_LVL_STRIDE 32   // ints between consecutive level fill counters (one 128-byte line each)

namespace rp {

// ------------------------------------------------------------------------------------------------------ body access
__device__ __forceinline__ V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ Q4 ld4(const double* p) { return q4(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ void st4(double* p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

__device__ __forceinline__ void load_static(Body& b, const BodyStatic& s) {
	b.inv_mass = s.inv_mass;
	b.inertia = s.inertia;
	b.inv_inertia = s.inv_inertia;
	b.mu_s = s.mu_s; b.mu_d = s.mu_d; b.rest = s.rest;
	b.fixed = s.fixed;
}
__device__ __forceinline__ void load_dyn(Body& b, const BodyDyn& d) {
	b.x = ld3(d.x); b.q = ld4(d.q); b.v = ld3(d.v); b.w = ld3(d.w);
	b.px = ld3(d.px); b.pq = ld4(d.pq); b.pv = ld3(d.pv); b.pw = ld3(d.pw);
}

// ------------------------------------------------------------------------------------------------------- broadphase
// broad_get_collision_pairs (broad.cpp:6-29): all i < j with |x_i - x_j| <= r_i + r_j + 0.1, emitted in (i, j) order.
// Row i is one thread; all threads of a CTA walk j together so the position loads broadcast.
template <bool WRITE>
__global__ void __launch_bounds__(128) k_broad_rows(DevView d) {
	const int w = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int row0 = blockIdx.x * blockDim.x;
	const BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
	V3 xi = v3(0.0, 0.0, 0.0);
	double ri = 0.0;
	int ci0 = 0, nci = 0;
	if (i < d.NB) {
		xi = ld3(dyn[i].x);
		ri = d.bstat[i].radius;
		ci0 = d.bstat[i].col0;
		nci = d.bstat[i].ncol;
	}
	int count = 0;
	int out = 0;
	PairRec* pairs = d.pairs + (size_t)w * d.max_pairs;
	if (WRITE && i < d.NB) out = d.row_off[(size_t)w * d.NB + i];
	for (int j = row0 + 1; j < d.NB; ++j) {
		if (i < d.NB && j > i) {
			V3 xj = ld3(dyn[j].x);
			double dist = length(sub(xi, xj));
			double maxd = ri + d.bstat[j].radius + 0.1;
			if (dist <= maxd) {
				int ncj = d.bstat[j].ncol;
				if (WRITE) {
					int cj0 = d.bstat[j].col0;
					for (int a = 0; a < nci; ++a) {
						for (int b = 0; b < ncj; ++b) {
							if (out < d.max_pairs) {
								PairRec pr;
								pr.a = i; pr.b = j; pr.ca = ci0 + a; pr.cb = cj0 + b;
								pairs[out] = pr;
							}
							++out;
						}
					}
				} else {
					count += nci * ncj;
				}
			}
		}
	}
	if (!WRITE && i < d.NB) d.row_off[(size_t)w * d.NB + i] = count;
}

// exclusive scan of the row counts of one world (one CTA per world)
__global__ void __launch_bounds__(256) k_broad_scan(DevView d) {
	const int w = blockIdx.x;
	int* row = d.row_off + (size_t)w * d.NB;
	__shared__ int warp_sums[8];
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (int base = 0; base < d.NB; base += blockDim.x) {
		int i = base + threadIdx.x;
		int v = i < d.NB ? row[i] : 0;
		int x = v;
		for (int o = 1; o < 32; o <<= 1) {
			int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) warp_sums[wid] = x;
		__syncthreads();
		int prefix = carry;
		for (int k = 0; k < wid; ++k) prefix += warp_sums[k];
		if (i < d.NB) row[i] = prefix + x - v;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) carry = prefix + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		int total = carry;
		if (total > d.max_pairs) {
			atomicOr(&d.status[w], ST_PAIR_CAPACITY);
			total = d.max_pairs;
		}
		d.n_pairs[w] = total;
		atomicAdd(&d.counters[CNT_BROAD_PAIRS], (unsigned long long)total);
	}
}

// ------------------------------------------------------------------------------------------------- islands + sleeping
// broad_collect_simulation_islands (broad.cpp:70-116) + the sleep bookkeeping of pbd.cpp:476-506. Islands are the
// connected components of {pairs, external constraints} restricted to non-fixed bodies; the result does not depend on
// the order in which unions happen, so min-label propagation replaces the reference's union-find. One CTA per world.
__global__ void __launch_bounds__(256) k_islands(DevView d, double dt) {
	const int w = blockIdx.x;
	int* label = d.label + (size_t)w * d.NB;
	int* flag = d.isl_flag + (size_t)w * d.NB;
	BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
	int* active = d.active + (size_t)w * d.NB;
	double* deact = d.deact + (size_t)w * d.NB;
	const PairRec* pairs = d.pairs + (size_t)w * d.max_pairs;
	const int np = d.n_pairs[w];
	__shared__ int changed;
	for (int b = threadIdx.x; b < d.NB; b += blockDim.x) {
		label[b] = b;
		flag[b] = 1;
	}
	__syncthreads();
	for (;;) {
		if (threadIdx.x == 0) changed = 0;
		__syncthreads();
		for (int e = threadIdx.x; e < np + d.NJ; e += blockDim.x) {
			int a, b;
			if (e < np) {
				a = pairs[e].a; b = pairs[e].b;
			} else {
				a = d.joints[e - np].e1; b = d.joints[e - np].e2;
			}
			if (d.bstat[a].fixed || d.bstat[b].fixed) continue;
			int la = label[a], lb = label[b];
			if (la != lb) {
				int m = la < lb ? la : lb;
				atomicMin(&label[a], m);
				atomicMin(&label[b], m);
				changed = 1;
			}
		}
		__syncthreads();
		int c = changed;
		__syncthreads();
		if (!c) break;
	}
	for (int b = threadIdx.x; b < d.NB; b += blockDim.x) {
		if (d.bstat[b].fixed) continue;
		double lv = length(ld3(dyn[b].v));
		double av = length(ld3(dyn[b].w));
		double t = deact[b];
		if (lv < d.lin_sleep && av < d.ang_sleep) t += dt;
		else t = 0.0;
		deact[b] = t;
		if (t < d.sleep_time) flag[label[b]] = 0;
	}
	__syncthreads();
	for (int b = threadIdx.x; b < d.NB; b += blockDim.x) {
		if (d.bstat[b].fixed) continue;
		active[b] = flag[label[b]] ? 0 : 1;
	}
}

// ------------------------------------------------------------------------------------------------------ level schedule
// Units of the Gauss-Seidel sweep in the reference's array order: the external constraints first (copy_constraints
// output is the head of the array, pbd.cpp:580), then the broadphase (collider-)pairs in pair order, each pair standing
// for its whole manifold (pbd.cpp:584-611). level(u) = 1 + max(level of the previous unit touching either of u's
// NON-FIXED bodies). The joints' levels are the same in every world (host, at batch creation); this kernel continues
// the recurrence over one world's pairs (one thread per world, once per frame) and adds the world's per-level pair
// counts to the global capacities of the level-major work lists.
__global__ void __launch_bounds__(64) k_schedule(DevView d, int collisions) {
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= d.W) return;
	int* last = d.last_level + (size_t)w * d.NB;
	int* plevel = d.pair_level + (size_t)w * d.max_pairs;
	int* hist = d.lvl_hist + (size_t)w * (d.max_levels + 2);
	const int* active = d.active + (size_t)w * d.NB;
	const PairRec* pairs = d.pairs + (size_t)w * d.max_pairs;
	const int np = collisions ? d.n_pairs[w] : 0;
	for (int b = 0; b < d.NB; ++b) last[b] = d.joint_last[b];
	int nl = d.joint_levels;
	for (int p = 0; p < np; ++p) {
		const int a = pairs[p].a, b = pairs[p].b;
		const int fa = d.bstat[a].fixed, fb = d.bstat[b].fixed;
		// pbd.cpp:594: nothing to do when both sides are fixed or asleep
		if ((fa || !active[a]) && (fb || !active[b])) {
			plevel[p] = 0;
			continue;
		}
		const int la = fa ? 0 : last[a], lb = fb ? 0 : last[b];
		const int lvl = 1 + (la > lb ? la : lb);
		if (!fa) last[a] = lvl;
		if (!fb) last[b] = lvl;
		plevel[p] = lvl;
		if (lvl > nl) nl = lvl;
	}
	for (int l = 0; l <= nl + 1; ++l) hist[l] = 0;
	for (int p = 0; p < np; ++p) hist[plevel[p]] += 1;
	for (int l = 1; l <= nl; ++l) {
		if (hist[l]) atomicAdd(&d.lvl_cap[l], hist[l]);
	}
	atomicMax(d.lvl_max, nl);
	atomicAdd(&d.counters[CNT_LEVELS], (unsigned long long)nl);
}

// zeroes the per-frame level capacities (before k_schedule) / turns them into list offsets (after it)
__global__ void __launch_bounds__(256) k_level_reset(DevView d) {
	for (int l = threadIdx.x; l < d.max_levels + 2; l += blockDim.x) d.lvl_cap[l] = 0;
	if (threadIdx.x == 0) *d.lvl_max = 0;
}
__global__ void k_level_offsets(DevView d) {
	int run = 0;
	const int nl = *d.lvl_max;
	for (int l = 0; l <= nl + 1; ++l) {
		d.lvl_off[l] = run;
		run += d.lvl_cap[l];
	}
}

// ---------------------------------------------------------------------------------------- integrate + collider update
// pbd.cpp:537-577 (integration) and collider.cpp:409-445 (collider_update) for one body per thread. The reference
// re-transforms both colliders of every pair every substep (39 % of its time); the same pose gives the same result,
// so once per body per substep is exactly equivalent. Also leaves each collider's world-space bounds
// for k_cull and resets the per-substep counters.
// per-substep counters
__global__ void __launch_bounds__(256) k_substep_reset(DevView d) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		*d.hit_count = 0u;
		*d.cand_count = 0u;
	}
	if (i < d.max_levels + 2) d.lvl_fill[(size_t)i * RP_LVL_STRIDE] = 0;
	if (i < d.W) d.n_contacts[i] = 0;
}

#define RP_INT_MAXV 8   // staged write path of k_integrate: bodies of up to 8 transformed vertices and 6 normals (boxes)
#define RP_INT_MAXF 6
__global__ void __launch_bounds__(128, RP_MINB_INTEGRATE) k_integrate(DevView d, double h) {
	// Transformed geometry of the CTA's 128 bodies, staged so that the global writes are coalesced. Rows are padded to an
	// odd number of doubles: per-thread rows are then free of bank conflicts.
	__shared__ double s_tv[128 * (RP_INT_MAXV * 3 + 1)];
	__shared__ double s_tn[128 * (RP_INT_MAXF * 3 + 1)];
	const int w = blockIdx.y;
	const int b0 = blockIdx.x * blockDim.x;
	const int b = b0 + threadIdx.x;
	const int nb = min(128, d.NB - b0);
	// staged path only when every body of the CTA has the same small footprint, laid out back to back
	const BodyStatic& s0 = d.bstat[b0];
	const int tvn = s0.tvn, tnn = s0.tnn;
	bool uniform = tvn <= RP_INT_MAXV && tnn <= RP_INT_MAXF;
	if (b < d.NB) {
		const BodyStatic& sb = d.bstat[b];
		uniform = uniform && sb.tvn == tvn && sb.tnn == tnn && sb.tv0 == s0.tv0 + (b - b0) * tvn && sb.tn0 == s0.tn0 + (b - b0) * tnn;
	}
	const bool staged = __syncthreads_and(uniform) != 0;
	const int rv = tvn * 3 + 1, rn = tnn * 3 + 1;
	if (b < d.NB) {
		const size_t gid = (size_t)w * d.NB + b;
		if (b < d.NJ) {  // copy_constraints resets every lambda each substep (pbd.cpp:426-462)
			for (int j = b; j < d.NJ; j += d.NB) {
				JointLambda z;
				z.a = z.b = z.c = 0.0;
				d.lambdas[(size_t)w * d.NJ + j] = z;
			}
		}
		const BodyStatic& s = d.bstat[b];
		BodyDyn& dd = d.dyn[gid];
		Body body;
		load_static(body, s);
		body.x = ld3(dd.x); body.q = ld4(dd.q); body.v = ld3(dd.v); body.w = ld3(dd.w);
		body.active = d.active[gid];
		integrate(body, h, d.force[b], d.torque[b]);
		st3(dd.px, body.px); st4(dd.pq, body.pq);
		if (!(body.fixed || !body.active)) {
			st3(dd.x, body.x); st4(dd.q, body.q); st3(dd.v, body.v); st3(dd.w, body.w);
		}
		Pose34 M = model_matrix(body.q, body.x);
		V3* tv = d.tv + (size_t)w * d.TV;
		V3* tn = d.tn + (size_t)w * d.TN;
		double* row_v = s_tv + (size_t)threadIdx.x * rv;
		double* row_n = s_tn + (size_t)threadIdx.x * rn;
		int ov = 0, on = 0;
		for (int c = s.col0; c < s.col0 + s.ncol; ++c) {
			const ColliderDesc cd = d.cols[c];
			double* bb = d.aabb + ((size_t)w * d.NC + c) * 6;
			if (cd.type == SHAPE_SPHERE) {
				if (staged) { row_v[ov] = body.x.x; row_v[ov + 1] = body.x.y; row_v[ov + 2] = body.x.z; ov += 3; }
				else tv[cd.tv0] = body.x;
				const double r = (double)cd.radius;
				bb[0] = body.x.x - r; bb[1] = body.x.y - r; bb[2] = body.x.z - r;
				bb[3] = body.x.x + r; bb[4] = body.x.y + r; bb[5] = body.x.z + r;
			} else {
				const HullTopo t = d.pool.hulls[cd.hull];
				double lo0 = 1.7976931348623157e308, lo1 = lo0, lo2 = lo0, hi0 = -lo0, hi1 = -lo0, hi2 = -lo0;
				for (int k = 0; k < t.nv; ++k) {
					const V3 p = transform_point(M, d.pool.verts[t.vert0 + k]);
					if (staged) { row_v[ov] = p.x; row_v[ov + 1] = p.y; row_v[ov + 2] = p.z; ov += 3; }
					else tv[cd.tv0 + k] = p;
					lo0 = fmin(lo0, p.x); lo1 = fmin(lo1, p.y); lo2 = fmin(lo2, p.z);
					hi0 = fmax(hi0, p.x); hi1 = fmax(hi1, p.y); hi2 = fmax(hi2, p.z);
				}
				for (int k = 0; k < t.nf; ++k) {
					const V3 n = transform_normal(M, d.pool.normals[t.face0 + k]);
					if (staged) { row_n[on] = n.x; row_n[on + 1] = n.y; row_n[on + 2] = n.z; on += 3; }
					else tn[cd.tn0 + k] = n;
				}
				bb[0] = lo0; bb[1] = lo1; bb[2] = lo2; bb[3] = hi0; bb[4] = hi1; bb[5] = hi2;
			}
		}
	}
	if (staged) {
		__syncthreads();
		double* gv = (double*)(d.tv + (size_t)w * d.TV + s0.tv0);
		double* gn = (double*)(d.tn + (size_t)w * d.TN + s0.tn0);
		const int nv3 = tvn * 3, nn3 = tnn * 3;
		for (int g = threadIdx.x; g < nb * nv3; g += blockDim.x) gv[g] = s_tv[(g / nv3) * rv + g % nv3];
		for (int g = threadIdx.x; g < nb * nn3; g += blockDim.x) gn[g] = s_tn[(g / nn3) * rn + g % nn3];
	}
}

// Copies a small hull's transformed vertices (and, optionally, face normals) from the world's AoS arrays into the calling
// thread's column of a thread-interleaved shared-memory block: element e of the thread lives at base[e * nthreads], so
// the 32 lanes of a warp touch 32 consecutive doubles per access (2 wavefronts) instead of 32 scattered sectors. The
// narrowphase scans the same vertices many times (support mapping), so this turns an L1-wavefront-bound kernel back
// into an FP64-bound one. Returns the number of doubles used.
__device__ __forceinline__ int stage_shape(Shape& s, double* col, int nthreads, bool with_normals) {
	int e = 0;
	const double* src = s.vp;
	for (int k = 0; k < s.nv * 3; ++k) col[(size_t)(e + k) * nthreads] = src[k];
	s.vp = col + (size_t)e * nthreads; s.vs = 3 * nthreads; s.vcs = nthreads;
	e += s.nv * 3;
	if (with_normals) {
		src = s.np;
		for (int k = 0; k < s.nf * 3; ++k) col[(size_t)(e + k) * nthreads] = src[k];
		s.np = col + (size_t)e * nthreads; s.ns = 3 * nthreads; s.ncs = nthreads;
		e += s.nf * 3;
	}
	return e;
}

// warp-aggregated append: every lane of the warp calls this; lanes with want == true get consecutive slots
__device__ __forceinline__ unsigned int warp_append(unsigned int* counter, bool want) {
	const unsigned int mask = __ballot_sync(0xffffffffu, want);
	if (!mask) return 0u;
	const int lane = threadIdx.x & 31;
	const int leader = __ffs(mask) - 1;
	unsigned int base = 0;
	if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + __popc(mask & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------------------- cull
// One thread per (world, collider pair): the pbd.cpp:594 skip rule, then an exact-safe bounds test. If the world-space
// boxes of the two colliders (bounds of the very vertex sets GJK would scan; sphere: centre +/- radius, its support set)
// are separated by more than RP_CULL_MARGIN along an axis, every Minkowski-difference support point has that coordinate
// strictly positive (or strictly negative), the origin is outside the difference, and gjk_collides returns false: the
// pair yields no contacts, exactly as if GJK had run. Survivors go to the dense candidate list of k_gjk.
#define RP_CULL_MARGIN 1e-7
__global__ void __launch_bounds__(256) k_cull(DevView d, int cull) {
	const int w = blockIdx.y;
	const int np = d.n_pairs[w];
	const int* active = d.active + (size_t)w * d.NB;
	int tested = 0;
	for (int p0 = blockIdx.x * blockDim.x; p0 < np; p0 += gridDim.x * blockDim.x) {
		const int p = p0 + threadIdx.x;
		bool keep = false;
		if (p < np) {
			const size_t pg = (size_t)w * d.max_pairs + p;
			d.pair_ccnt[pg] = 0;
			const PairRec pr = d.pairs[pg];
			if (!((d.bstat[pr.a].fixed || !active[pr.a]) && (d.bstat[pr.b].fixed || !active[pr.b]))) {
				++tested;
				keep = true;
				if (cull) {
					const double* A = d.aabb + ((size_t)w * d.NC + pr.ca) * 6;
					const double* B = d.aabb + ((size_t)w * d.NC + pr.cb) * 6;
					const bool both_spheres = d.cols[pr.ca].type == SHAPE_SPHERE && d.cols[pr.cb].type == SHAPE_SPHERE;
					if (!both_spheres) {  // sphere-sphere pairs never reach GJK (collider.cpp:530)
						for (int k = 0; k < 3; ++k) {
							if (A[k] - B[3 + k] > RP_CULL_MARGIN || B[k] - A[3 + k] > RP_CULL_MARGIN) keep = false;
						}
					}
				}
			}
		}
		const unsigned int slot = warp_append(d.cand_count, keep);
		if (keep) d.cands[slot] = make_uint2((unsigned int)w, (unsigned int)p);
	}
	for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(0xffffffffu, tested, o);
	if ((threadIdx.x & 31) == 0 && tested) atomicAdd(&d.counters[CNT_PAIR_TESTS], (unsigned long long)tested);
}

// ----------------------------------------------------------------------------------------------------- narrowphase 1
// One thread per candidate pair: sphere-sphere test or boolean GJK (collider.cpp:523-547). Colliding pairs are appended
// to the global hit list for k_manifold.
__global__ void __launch_bounds__(RP_GJK_THREADS, RP_MINB_GJK) k_gjk(DevView d) {
	const unsigned int nc = *d.cand_count;
	__shared__ double s_stage[RP_GJK_STAGE * RP_GJK_THREADS];
	for (unsigned int c0 = blockIdx.x * blockDim.x; c0 < nc; c0 += gridDim.x * blockDim.x) {
		const unsigned int ci = c0 + threadIdx.x;
		bool hit = false;
		Simplex s;
		s.a = s.b = s.c = s.d = v3(0.0, 0.0, 0.0);
		int w = 0, p = 0;
		if (ci < nc) {
			const uint2 cd = d.cands[ci];
			w = (int)cd.x; p = (int)cd.y;
			const PairRec pr = d.pairs[(size_t)w * d.max_pairs + p];
			const V3* tv = d.tv + (size_t)w * d.TV;
			const V3* tn = d.tn + (size_t)w * d.TN;
			Shape A = make_shape(d.pool, d.cols[pr.ca], tv, tn);
			Shape B = make_shape(d.pool, d.cols[pr.cb], tv, tn);
			int st = 0;
			if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
				V3 n;
				double depth;
				hit = sphere_sphere(A, B, &n, &depth);
			} else {
				if ((A.nv + B.nv) * 3 <= RP_GJK_STAGE) {
					double* col = s_stage + threadIdx.x;
					const int used = stage_shape(A, col, RP_GJK_THREADS, false);
					stage_shape(B, col + (size_t)used * RP_GJK_THREADS, RP_GJK_THREADS, false);
				}
				hit = gjk(A, B, &s, &st, 0);
			}
			if (st) atomicOr(&d.status[w], st);
		}
		const unsigned int slot = warp_append(d.hit_count, hit);
		if (hit) {
			HitRec hr;
			hr.world = w; hr.pair = p;
			hr.sa = s.a; hr.sb = s.b; hr.sc = s.c; hr.sd = s.d;
			d.hits[slot] = hr;
		}
	}
}

// ----------------------------------------------------------------------------------------------------- narrowphase 2
struct StageSink {
	V3* stage;
	int n, cap;
	__device__ __forceinline__ void operator()(V3 p1, V3 p2) {
		if (n < cap) {
			stage[2 * n] = p1;
			stage[2 * n + 1] = p2;
		}
		++n;
	}
};

struct ManifoldScratch {
	union {
		EpaScratch epa;
		struct {
			ClipScratch clip;
			V3 stage[2 * RP_CLIP_MAX_POINTS];
		} m;
	};
};

// One thread per colliding collider pair: EPA (epa.cpp:118), manifold (clipping.cpp:343), contact -> constraint
// (pbd.cpp:408-424). The pair's contacts get a contiguous run in the world's contact buffer (allocation order between
// pairs is irrelevant: the solver walks pairs, not the buffer), and the pair is appended to the work list of its
// dependency level.
__global__ void __launch_bounds__(RP_MANIFOLD_THREADS, RP_MINB_MANIFOLD) k_manifold(DevView d) {
	__shared__ double s_stage[RP_MANIFOLD_STAGE * RP_MANIFOLD_THREADS + 1];
	const unsigned int nh = *d.hit_count;
	ManifoldScratch sc;
	__shared__ int s_cnt[RP_LVL_SMEM], s_base[RP_LVL_SMEM];
	int made = 0;
	const int lane = threadIdx.x & 31;
	for (unsigned int h0 = blockIdx.x * blockDim.x; h0 < nh; h0 += gridDim.x * blockDim.x) {
		const unsigned int hi = h0 + threadIdx.x;
		int n = 0, lvl = -1, w = 0, pair = 0;
		if (hi < nh) {
			const HitRec hr = d.hits[hi];
			w = hr.world; pair = hr.pair;
			const size_t pg = (size_t)w * d.max_pairs + hr.pair;
			const PairRec pr = d.pairs[pg];
			const V3* tv = d.tv + (size_t)w * d.TV;
			const V3* tn = d.tn + (size_t)w * d.TN;
			Shape A = make_shape(d.pool, d.cols[pr.ca], tv, tn);
			Shape B = make_shape(d.pool, d.cols[pr.cb], tv, tn);
			if (A.type == SHAPE_HULL && B.type == SHAPE_HULL && (A.nv + B.nv + A.nf + B.nf) * 3 <= RP_MANIFOLD_STAGE) {
				double* col = s_stage + threadIdx.x;
				const int used = stage_shape(A, col, RP_MANIFOLD_THREADS, true);
				stage_shape(B, col + (size_t)used * RP_MANIFOLD_THREADS, RP_MANIFOLD_THREADS, true);
			}
			V3 normal;
			double depth;
			int st = 0;
			bool ok;
			if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
				ok = sphere_sphere(A, B, &normal, &depth);
			} else {
				Simplex s;
				s.a = hr.sa; s.b = hr.sb; s.c = hr.sc; s.d = hr.sd;
				s.num = 4;
				ok = epa(A, B, s, sc.epa, &normal, &depth, &st, 0);
			}
			if (ok) {
				StageSink sink;
				sink.stage = sc.m.stage;
				sink.n = 0;
				sink.cap = RP_CLIP_MAX_POINTS;
				manifold(A, B, normal, depth, sc.m.clip, &st, sink);
				n = sink.n;
				if (n > sink.cap) {
					st |= ST_CLIP_CAPACITY;
					n = sink.cap;
				}
			}
			if (n > 0) {
				int off = atomicAdd(&d.n_contacts[w], n);
				if (off + n > d.max_contacts) {
					st |= ST_CONTACT_CAPACITY;
					n = d.max_contacts - off;
					if (n < 0) n = 0;
				}
				const BodyDyn& da = d.dyn[(size_t)w * d.NB + pr.a];
				const BodyDyn& db = d.dyn[(size_t)w * d.NB + pr.b];
				Body b1, b2;
				b1.x = ld3(da.x); b1.q = ld4(da.q);
				b2.x = ld3(db.x); b2.q = ld4(db.q);
				Contact* out = d.contacts + (size_t)w * d.max_contacts + off;
				for (int k = 0; k < n; ++k) {
					out[k] = make_contact(b1, b2, sc.m.stage[2 * k], sc.m.stage[2 * k + 1]);
					if (w == d.dbg_world) {
						d.dbg_points[2 * (off + k)] = sc.m.stage[2 * k];
						d.dbg_points[2 * (off + k) + 1] = sc.m.stage[2 * k + 1];
					}
				}
				d.pair_normal[pg] = normal;
				d.pair_coff[pg] = off;
				d.pair_ccnt[pg] = n;
				made += n;
				if (n > 0) lvl = d.pair_level[pg];
			}
			if (st) atomicOr(&d.status[w], st);
		}
		// append (world, pair) to the list of its level: ranks within the CTA through shared-memory counters, then ONE
		// global atomic per (CTA, level) on a counter that owns its 128-byte line (RP_LVL_STRIDE)
		__syncthreads();
		if (threadIdx.x < RP_LVL_SMEM) s_cnt[threadIdx.x] = 0;  // RP_MANIFOLD_THREADS >= RP_LVL_SMEM
		__syncthreads();
		int rank = 0;
		if (lvl > 0) {
			if (lvl < RP_LVL_SMEM) rank = atomicAdd(&s_cnt[lvl], 1);
			else rank = atomicAdd(&d.lvl_fill[(size_t)lvl * RP_LVL_STRIDE], 1);  // very deep schedules: direct
		}
		__syncthreads();
		if (threadIdx.x < RP_LVL_SMEM && s_cnt[threadIdx.x] > 0) {
			s_base[threadIdx.x] = atomicAdd(&d.lvl_fill[(size_t)threadIdx.x * RP_LVL_STRIDE], s_cnt[threadIdx.x]);
		}
		__syncthreads();
		if (lvl > 0) {
			const int slot = (lvl < RP_LVL_SMEM ? s_base[lvl] : 0) + rank;
			d.lvl_items[d.lvl_off[lvl] + slot] = make_uint2((unsigned int)w, (unsigned int)pair);
		}
	}
	for (int o = 16; o > 0; o >>= 1) made += __shfl_down_sync(0xffffffffu, made, o);
	if (lane == 0 && made) atomicAdd(&d.counters[CNT_CONTACTS], (unsigned long long)made);
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&d.counters[CNT_HITS], (unsigned long long)nh);
}

// -------------------------------------------------------------------------------------------------------------- solve
// Level-major Gauss-Seidel across ALL worlds: launch l runs every constraint of dependency level l, one thread per
// unit -- the joints of level l of every world, then the (world, pair) items of level l that have contacts this substep
// (a pair's manifold is a sequential chain on its two bodies and stays in one thread, bodies in registers). Kernel
// boundaries are the barriers between levels, so the result equals the reference's sequential sweep (pbd.cpp:615-620).
__global__ void __launch_bounds__(128, RP_MINB_POS) k_pos_level(DevView d, double h, int level, int collisions) {
	const int nj = level <= d.joint_levels ? d.joint_lptr[level] - d.joint_lptr[level - 1] : 0;
	const int njw = nj * d.W;
	const int np = collisions ? d.lvl_fill[(size_t)level * RP_LVL_STRIDE] : 0;
	int st = 0, stw = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < njw + np; i += gridDim.x * blockDim.x) {
		if (i < njw) {
			const int w = i / nj;
			const int u = d.joint_sched[d.joint_lptr[level - 1] + i % nj];
			const Joint j = d.joints[u];
			BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
			Body b1, b2;
			load_static(b1, d.bstat[j.e1]);
			load_static(b2, d.bstat[j.e2]);
			BodyDyn& d1 = dyn[j.e1];
			BodyDyn& d2 = dyn[j.e2];
			b1.x = ld3(d1.x); b1.q = ld4(d1.q);
			b2.x = ld3(d2.x); b2.q = ld4(d2.q);
			JointLambda lam = d.lambdas[(size_t)w * d.NJ + u];
			solve_joint(j, lam, b1, b2, h, &st);
			d.lambdas[(size_t)w * d.NJ + u] = lam;
			if (!b1.fixed) { st3(d1.x, b1.x); st4(d1.q, b1.q); }
			if (!b2.fixed) { st3(d2.x, b2.x); st4(d2.q, b2.q); }
			stw = w;
		} else {
			const uint2 item = d.lvl_items[d.lvl_off[level] + (i - njw)];
			const int w = (int)item.x;
			const size_t pg = (size_t)w * d.max_pairs + item.y;
			const int cnt = d.pair_ccnt[pg];
			const PairRec pr = d.pairs[pg];
			const V3 normal = d.pair_normal[pg];
			Contact* cs = d.contacts + (size_t)w * d.max_contacts + d.pair_coff[pg];
			BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
			Body b1, b2;
			load_static(b1, d.bstat[pr.a]);
			load_static(b2, d.bstat[pr.b]);
			BodyDyn& d1 = dyn[pr.a];
			BodyDyn& d2 = dyn[pr.b];
			b1.x = ld3(d1.x); b1.q = ld4(d1.q); b1.px = ld3(d1.px); b1.pq = ld4(d1.pq);
			b2.x = ld3(d2.x); b2.q = ld4(d2.q); b2.px = ld3(d2.px); b2.pq = ld4(d2.pq);
			for (int c = 0; c < cnt; ++c) {
				Contact ct = cs[c];
				solve_contact(ct, normal, b1, b2, h, &st);
				cs[c].lambda_n = ct.lambda_n;
				cs[c].lambda_t = ct.lambda_t;
			}
			if (!b1.fixed) { st3(d1.x, b1.x); st4(d1.q, b1.q); }
			if (!b2.fixed) { st3(d2.x, b2.x); st4(d2.q, b2.q); }
			stw = w;
		}
		if (st) {
			atomicOr(&d.status[stw], st);
			st = 0;
		}
	}
}

// velocity derivation (pbd.cpp:623-643), one thread per body
__global__ void __launch_bounds__(128) k_derive(DevView d, double h) {
	const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (size_t)d.W * d.NB) return;
	const int b = (int)(gid % d.NB);
	Body body;
	body.fixed = d.bstat[b].fixed;
	body.active = d.active[gid];
	if (body.fixed || !body.active) return;
	BodyDyn& dd = d.dyn[gid];
	body.x = ld3(dd.x); body.q = ld4(dd.q); body.px = ld3(dd.px); body.pq = ld4(dd.pq);
	body.v = ld3(dd.v); body.w = ld3(dd.w);
	derive_velocity(body, h);
	st3(dd.v, body.v); st3(dd.w, body.w); st3(dd.pv, body.pv); st3(dd.pw, body.pw);
}

// velocity pass over the contacts of one level (pbd.cpp:646-711); the hinge branch of the reference's velocity pass is
// an empty TODO (pbd.cpp:712-739), so joints take no part
__global__ void __launch_bounds__(128, RP_MINB_VEL) k_vel_level(DevView d, double h, int level) {
	const int np = d.lvl_fill[(size_t)level * RP_LVL_STRIDE];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
		const uint2 item = d.lvl_items[d.lvl_off[level] + i];
		const int w = (int)item.x;
		const size_t pg = (size_t)w * d.max_pairs + item.y;
		const int cnt = d.pair_ccnt[pg];
		const PairRec pr = d.pairs[pg];
		const V3 normal = d.pair_normal[pg];
		const Contact* cs = d.contacts + (size_t)w * d.max_contacts + d.pair_coff[pg];
		BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
		Body b1, b2;
		load_static(b1, d.bstat[pr.a]);
		load_static(b2, d.bstat[pr.b]);
		BodyDyn& d1 = dyn[pr.a];
		BodyDyn& d2 = dyn[pr.b];
		b1.q = ld4(d1.q); b1.v = ld3(d1.v); b1.w = ld3(d1.w); b1.pv = ld3(d1.pv); b1.pw = ld3(d1.pw);
		b2.q = ld4(d2.q); b2.v = ld3(d2.v); b2.w = ld3(d2.w); b2.pv = ld3(d2.pv); b2.pw = ld3(d2.pw);
		for (int c = 0; c < cnt; ++c) {
			const Contact ct = cs[c];
			solve_contact_velocity(ct, normal, b1, b2, h);
		}
		if (!b1.fixed) { st3(d1.v, b1.v); st3(d1.w, b1.w); }
		if (!b2.fixed) { st3(d2.v, b2.v); st3(d2.w, b2.w); }
	}
}

// ------------------------------------------------------------------------------------------------ FP64 pipe probe
// Roofline denominator for this path: sustained FP64 rate of the CUDA-core pipe with independent DADD/DMUL chains (the
// form the parity build issues: --fmad=false) or DFMA chains (what the pipe could do if contraction were allowed).
template <bool FMA>
__global__ void __launch_bounds__(256) k_fp64_probe(double* out, int iters, double a, double b) {
	double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
	for (int i = 0; i < iters; ++i) {
		if (FMA) {
			x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
			x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
		} else {
			x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b); x2 = __dadd_rn(__dmul_rn(x2, a), b);
			x3 = __dadd_rn(__dmul_rn(x3, a), b); x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
			x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
		}
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_count_frame(DevView d) { atomicAdd(&d.counters[CNT_FRAMES], 1ull); }

// -------------------------------------------------------------------------------------------------- state pack/unpack
// host record (rawphys_b200.h RP_STATE_STRIDE = 21 doubles) <-> BodyDyn + active + deactivation time
__global__ void __launch_bounds__(128) k_unpack_state(DevView d, const double* rec, int first_world, int n_worlds, int broadcast) {
	const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (size_t)n_worlds * d.NB) return;
	const int wl = (int)(gid / d.NB), b = (int)(gid % d.NB);
	const double* r = rec + (broadcast ? (size_t)b : gid) * 21;
	const size_t o = (size_t)(first_world + wl) * d.NB + b;
	BodyDyn& dd = d.dyn[o];
	for (int k = 0; k < 3; ++k) { dd.x[k] = r[k]; dd.v[k] = r[7 + k]; dd.w[k] = r[10 + k]; dd.pv[k] = r[15 + k]; dd.pw[k] = r[18 + k]; dd.px[k] = r[k]; }
	for (int k = 0; k < 4; ++k) { dd.q[k] = r[3 + k]; dd.pq[k] = r[3 + k]; }
	d.active[o] = r[13] != 0.0 ? 1 : 0;
	d.deact[o] = r[14];
}
__global__ void __launch_bounds__(128) k_pack_state(DevView d, double* rec, int first_world, int n_worlds) {
	const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (size_t)n_worlds * d.NB) return;
	const int wl = (int)(gid / d.NB), b = (int)(gid % d.NB);
	double* r = rec + gid * 21;
	const size_t o = (size_t)(first_world + wl) * d.NB + b;
	const BodyDyn& dd = d.dyn[o];
	for (int k = 0; k < 3; ++k) { r[k] = dd.x[k]; r[7 + k] = dd.v[k]; r[10 + k] = dd.w[k]; r[15 + k] = dd.pv[k]; r[18 + k] = dd.pw[k]; }
	for (int k = 0; k < 4; ++k) r[3 + k] = dd.q[k];
	r[13] = d.active[o] ? 1.0 : 0.0;
	r[14] = d.deact[o];
}

}  // namespace rp
#endif
