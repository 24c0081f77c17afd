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
#ifndef RP_MINB_EPA
#define RP_MINB_EPA 8
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
#define RP_EPA_THREADS 64

#define RP_LVL_SMEM 64     // levels ranked through shared memory in k_manifold
#define RP_LVL_STRIDE 32   // ints between consecutive level fill counters (one 128-byte line each: [0] front, [1] back)
#ifndef RP_SMALL_MANIFOLD
#define RP_SMALL_MANIFOLD 1000000  // (off: the refill loops of the solver kernels make manifold length irrelevant)
// manifolds of up to this many contacts fill a level list from the front, larger ones from the back
#endif

namespace rp {

// ------------------------------------------------------------------------------------------------------ body access
__device__ __forceinline__ V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ Q4 ld4(const double* p) { return q4(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ void st4(double* p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

__device__ __forceinline__ void load_static(Body& b, const DevView& d, int body) {
	const BodyStatic& s = d.bstat[body];
	const BodyClass& c = d.bclass[s.cls];
	b.inv_mass = c.inv_mass;
	b.inertia = c.inertia;
	b.inv_inertia = c.inv_inertia;
	b.mu_s = c.mu_s; b.mu_d = c.mu_d; b.rest = c.rest;
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
	if (i < d.max_levels + 2) {
		d.lvl_fill[(size_t)i * RP_LVL_STRIDE] = 0;
		d.lvl_fill[(size_t)i * RP_LVL_STRIDE + 1] = 0;
	}
	if (i < d.W) d.n_contacts[i] = 0;
}

#define RP_INT_MAXV 8   // staged write path of k_integrate: bodies of up to 8 transformed vertices and 6 normals (boxes)
#define RP_INT_MAXF 6
#define RP_INT_THREADS 128
#define RP_DYN_DOUBLES 26   // sizeof(BodyDyn) / 8
#define RP_DYN_LIVE 20      // x q v w px pq: the part of a record k_integrate reads or writes (pv, pw stay untouched)
#define RP_DYN_ROW 27       // odd row pitch in shared memory: per-thread rows are free of bank conflicts
// One thread per (world, body), flat over the whole batch: a CTA owns 128 consecutive BodyDyn records, which are
// contiguous in HBM, so they are moved with coalesced loads/stores through shared memory (per-lane gathers of 208-byte
// records were the kernel's long-scoreboard stall). The same shared memory is then reused to stage the transformed
// geometry of the CTA's bodies so that those writes are coalesced too.
__global__ void __launch_bounds__(RP_INT_THREADS, RP_MINB_INTEGRATE) k_integrate(DevView d, double h) {
	__shared__ double s_buf[RP_INT_THREADS * ((RP_INT_MAXV * 3 + 1) + (RP_INT_MAXF * 3 + 1))];
	static_assert(RP_INT_THREADS * RP_DYN_ROW <= RP_INT_THREADS * ((RP_INT_MAXV * 3 + 1) + (RP_INT_MAXF * 3 + 1)), "staging buffer");
	const size_t total = (size_t)d.W * d.NB;
	const size_t g0 = (size_t)blockIdx.x * RP_INT_THREADS;
	const int nb = (int)(total - g0 < (size_t)RP_INT_THREADS ? total - g0 : (size_t)RP_INT_THREADS);
	const size_t gid = g0 + threadIdx.x;
	const bool live = threadIdx.x < nb;
	const int w = live ? (int)(gid / d.NB) : 0;
	const int b = live ? (int)(gid % d.NB) : 0;
	// ---- coalesced load of the CTA's records
	{
		const double* g = (const double*)(d.dyn + g0);
		int r = threadIdx.x / RP_DYN_DOUBLES, c = threadIdx.x - r * RP_DYN_DOUBLES;
		for (int e = threadIdx.x; e < nb * RP_DYN_DOUBLES; e += RP_INT_THREADS) {
			if (c < 13) s_buf[r * RP_DYN_ROW + c] = g[e];
			r += RP_INT_THREADS / RP_DYN_DOUBLES; c += RP_INT_THREADS % RP_DYN_DOUBLES;
			if (c >= RP_DYN_DOUBLES) { c -= RP_DYN_DOUBLES; ++r; }
		}
	}
	__syncthreads();
	Body body;
	BodyStatic s;
	s.fixed = 1; s.col0 = 0; s.ncol = 0; s.tv0 = 0; s.tvn = 0; s.tn0 = 0; s.tnn = 0; s.cls = 0; s.radius = 0.0;
	double* row = s_buf + threadIdx.x * RP_DYN_ROW;
	if (live) {
		if (b < d.NJ) {  // copy_constraints resets every lambda each substep (pbd.cpp:426-462)
			for (int j = b; j < d.NJ; j += d.NB) {
				JointLambda z;
				z.a = z.b = z.c = 0.0;
				d.lambdas[(size_t)w * d.NJ + j] = z;
			}
		}
		s = d.bstat[b];
		load_static(body, d, b);
		body.x = ld3(row); body.q = ld4(row + 3); body.v = ld3(row + 7); body.w = ld3(row + 10);
		body.active = d.active[gid];
		integrate(body, h, d.force[b], d.torque[b]);
		st3(row, body.x); st4(row + 3, body.q); st3(row + 7, body.v); st3(row + 10, body.w);  // unchanged when fixed or asleep
		st3(row + 13, body.px); st4(row + 16, body.pq);
	}
	__syncthreads();
	{
		double* g = (double*)(d.dyn + g0);
		int r = threadIdx.x / RP_DYN_DOUBLES, c = threadIdx.x - r * RP_DYN_DOUBLES;
		for (int e = threadIdx.x; e < nb * RP_DYN_DOUBLES; e += RP_INT_THREADS) {
			if (c < RP_DYN_LIVE) g[e] = s_buf[r * RP_DYN_ROW + c];
			r += RP_INT_THREADS / RP_DYN_DOUBLES; c += RP_INT_THREADS % RP_DYN_DOUBLES;
			if (c >= RP_DYN_DOUBLES) { c -= RP_DYN_DOUBLES; ++r; }
		}
	}
	// ---- collider update. Staged path only when every body of the CTA has the same small footprint and the CTA's
	// blocks of transformed vertices / normals are laid out back to back (they are, across world boundaries too, when
	// all bodies of the scene have that footprint: TV = NB * tvn).
	size_t voff = 0, noff = 0;
	if (live) {
		voff = (size_t)w * d.TV + s.tv0;
		noff = (size_t)w * d.TN + s.tn0;
	}
	__shared__ size_t s_off[2];
	__shared__ int s_foot[2];
	if (threadIdx.x == 0) { s_off[0] = voff; s_off[1] = noff; s_foot[0] = s.tvn; s_foot[1] = s.tnn; }
	__syncthreads();  // also fences the write-back reads of s_buf against the staging writes below
	const int tvn = s_foot[0], tnn = s_foot[1];
	bool uniform = tvn <= RP_INT_MAXV && tnn <= RP_INT_MAXF;
	if (live) uniform = uniform && s.tvn == tvn && s.tnn == tnn && voff == s_off[0] + (size_t)threadIdx.x * tvn && noff == s_off[1] + (size_t)threadIdx.x * tnn;
	const bool staged = __syncthreads_and(uniform) != 0;
	const int rv = tvn * 3 + 1, rn = tnn * 3 + 1;
	double* s_tv = s_buf;
	double* s_tn = s_buf + RP_INT_THREADS * (RP_INT_MAXV * 3 + 1);
	if (live) {
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
		double* gv = (double*)(d.tv + s_off[0]);
		double* gn = (double*)(d.tn + s_off[1]);
		const int nv3 = tvn * 3, nn3 = tnn * 3;
		// g = r * n3 + c walks in steps of the CTA size; (r, c) are advanced without dividing
		if (nv3 > 0) {
			int r = threadIdx.x / nv3, c = threadIdx.x - r * nv3;
			const int dr = RP_INT_THREADS / nv3, dc = RP_INT_THREADS - dr * nv3;
			for (int g = threadIdx.x; g < nb * nv3; g += RP_INT_THREADS) {
				gv[g] = s_tv[r * rv + c];
				r += dr; c += dc;
				if (c >= nv3) { c -= nv3; ++r; }
			}
		}
		if (nn3 > 0) {
			int r = threadIdx.x / nn3, c = threadIdx.x - r * nn3;
			const int dr = RP_INT_THREADS / nn3, dc = RP_INT_THREADS - dr * nn3;
			for (int g = threadIdx.x; g < nb * nn3; g += RP_INT_THREADS) {
				gn[g] = s_tn[r * rn + c];
				r += dr; c += dc;
				if (c >= nn3) { c -= nn3; ++r; }
			}
		}
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

// ----------------------------------------------------------------------------------------------------- narrowphase
// GJK, EPA and the contact solves are loops whose trip counts differ from pair to pair (GJK 1..15 support iterations,
// EPA 1..8, manifolds of 1..8 contacts). With one work item per lane a warp runs as long as its slowest lane and the
// other lanes idle (ncu, round 1: 9..15 of 32 lanes active). The kernels below are therefore written as REFILL loops:
// a warp owns a contiguous chunk of the work list, every trip of the loop is one iteration of the algorithm for
// whatever item a lane currently holds, and a lane whose item is finished takes the warp's next item before the next
// trip. The warp's cursor is warp-uniform (ballot + popc), so taking work needs no atomics.
struct WarpQueue {
	unsigned int next, end;
	__device__ __forceinline__ void init(unsigned int n_items) {
		const unsigned int warps = gridDim.x * (blockDim.x >> 5);
		const unsigned int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
		unsigned int chunk = (n_items + warps - 1) / warps;
		if (chunk < 32u) chunk = 32u;
		const unsigned long long lo = (unsigned long long)wid * chunk;
		next = lo < n_items ? (unsigned int)lo : n_items;
		end = lo + chunk < n_items ? (unsigned int)(lo + chunk) : n_items;
	}
	// every lane calls this; lanes with want == true get the next items of the chunk (or 0xffffffff when it is used up)
	__device__ __forceinline__ unsigned int take(bool want) {
		const unsigned int mask = __ballot_sync(0xffffffffu, want);
		const unsigned int mine = next + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u));
		next += __popc(mask);
		if (next > end) next = end;
		return want && mine < end ? mine : 0xffffffffu;
	}
	__device__ __forceinline__ bool empty() const { return next >= end; }
};

// One lane per candidate pair at a time: sphere-sphere test or boolean GJK (collider.cpp:523-547). Writes the verdict of
// every candidate and, for colliding pairs, the final simplex, both at the candidate's own index (k_hits compacts).
__global__ void __launch_bounds__(RP_GJK_THREADS, RP_MINB_GJK) k_gjk(DevView d) {
	const unsigned int nc = *d.cand_count;
	__shared__ double s_stage[RP_GJK_STAGE * RP_GJK_THREADS];
	WarpQueue q;
	q.init(nc);
	bool have = false;
	unsigned int ci = 0;
	int w = 0, iter = 0, st = 0;
	Simplex s;
	s.a = s.b = s.c = s.d = v3(0.0, 0.0, 0.0);
	s.num = 0;
	V3 dir = v3(0.0, 0.0, 0.0);
	Shape A, B;
	A.type = B.type = SHAPE_SPHERE; A.nv = B.nv = 0;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			ci = got;
			const uint2 cd = d.cands[ci];
			w = (int)cd.x;
			const PairRec pr = d.pairs[(size_t)w * d.max_pairs + cd.y];
			const V3* tv = d.tv + (size_t)w * d.TV;
			const V3* tn = d.tn + (size_t)w * d.TN;
			A = make_shape(d.pool, d.cols[pr.ca], tv, tn);
			B = make_shape(d.pool, d.cols[pr.cb], tv, tn);
			if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
				V3 n;
				double depth;
				d.verdict[ci] = sphere_sphere(A, B, &n, &depth) ? 1 : 0;  // decided on the spot; the lane refills next trip
			} else {
				if ((A.nv + B.nv) * 3 <= RP_GJK_STAGE) {
					double* col = s_stage + threadIdx.x;
					const int used = stage_shape(A, col, RP_GJK_THREADS, false);
					stage_shape(B, col + (size_t)used * RP_GJK_THREADS, RP_GJK_THREADS, false);
				}
				gjk_begin(A, B, &s, &dir);
				iter = 0;
				have = true;
			}
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			int r = gjk_step(A, B, &s, &dir, &st);
			if (r == GJK_CONTINUE && ++iter >= RP_GJK_MAX_ITERS) r = GJK_MISS;  // gjk.cpp:358
			if (r != GJK_CONTINUE) {
				d.verdict[ci] = r == GJK_HIT ? 1 : 0;
				if (r == GJK_HIT) {
					V3* o = d.simplex + (size_t)ci * 4;
					o[0] = s.a; o[1] = s.b; o[2] = s.c; o[3] = s.d;
				}
				if (st) {
					atomicOr(&d.status[w], st);
					st = 0;
				}
				have = false;
			}
		}
	}
}

// candidate verdicts -> dense hit list (order is irrelevant downstream)
__global__ void __launch_bounds__(256) k_hits(DevView d) {
	const unsigned int nc = *d.cand_count;
	for (unsigned int c0 = blockIdx.x * blockDim.x; c0 < nc; c0 += gridDim.x * blockDim.x) {
		const unsigned int ci = c0 + threadIdx.x;
		const bool hit = ci < nc && d.verdict[ci] != 0;
		const unsigned int slot = warp_append(d.hit_count, hit);
		if (hit) d.hits[slot] = ci;
	}
}

// EPA (epa.cpp:118) for every hit, refill loop over its iterations. The polytope lives in the lane's local memory; the
// two hulls' vertices are staged in shared memory as in k_gjk (EPA only ever asks for support points).
__global__ void __launch_bounds__(RP_EPA_THREADS, RP_MINB_EPA) k_epa(DevView d) {
	const unsigned int nh = *d.hit_count;
	__shared__ double s_stage[RP_GJK_STAGE * RP_EPA_THREADS];
	WarpQueue q;
	q.init(nh);
	bool have = false;
	unsigned int hi = 0;
	int w = 0, iter = 0, st = 0;
	EpaScratch e;
	Shape A, B;
	A.type = B.type = SHAPE_SPHERE; A.nv = B.nv = 0;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			hi = got;
			const unsigned int ci = d.hits[hi];
			const uint2 cd = d.cands[ci];
			w = (int)cd.x;
			const PairRec pr = d.pairs[(size_t)w * d.max_pairs + cd.y];
			const V3* tv = d.tv + (size_t)w * d.TV;
			const V3* tn = d.tn + (size_t)w * d.TN;
			A = make_shape(d.pool, d.cols[pr.ca], tv, tn);
			B = make_shape(d.pool, d.cols[pr.cb], tv, tn);
			EpaOut out;
			out.ok = 0; out.pad = 0; out.depth = 0.0; out.normal = v3(0.0, 0.0, 0.0);
			if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
				out.ok = sphere_sphere(A, B, &out.normal, &out.depth) ? 1 : 0;
				d.epa_out[hi] = out;
			} else {
				if ((A.nv + B.nv) * 3 <= RP_GJK_STAGE) {
					double* col = s_stage + threadIdx.x;
					const int used = stage_shape(A, col, RP_EPA_THREADS, false);
					stage_shape(B, col + (size_t)used * RP_EPA_THREADS, RP_EPA_THREADS, false);
				}
				const V3* sp = d.simplex + (size_t)ci * 4;
				Simplex s;
				s.a = sp[0]; s.b = sp[1]; s.c = sp[2]; s.d = sp[3];
				s.num = 4;
				if (epa_begin(s, e, &st) == EPA_FAIL) {
					d.epa_out[hi] = out;
					atomicOr(&d.status[w], st);
					st = 0;
				} else {
					iter = 0;
					have = true;
				}
			}
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			int r = epa_step(A, B, e, &st);
			if (r == EPA_CONTINUE && ++iter >= RP_EPA_MAX_ITERS) {
				st |= ST_EPA_NO_CONVERGENCE;  // epa.cpp:233
				r = EPA_FAIL;
			}
			if (r != EPA_CONTINUE) {
				EpaOut out;
				out.ok = r == EPA_DONE ? 1 : 0;
				out.pad = 0;
				out.normal = e.min_normal;
				out.depth = e.min_dist;
				d.epa_out[hi] = out;
				if (st) {
					atomicOr(&d.status[w], st);
					st = 0;
				}
				have = false;
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&d.counters[CNT_HITS], (unsigned long long)nh);
}

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
	ClipScratch clip;
	V3 stage[2 * RP_CLIP_MAX_POINTS];
};

// One thread per colliding collider pair whose EPA converged: manifold (clipping.cpp:343), contact -> constraint
// (pbd.cpp:408-424). The pair's contacts get a contiguous run in the world's contact buffer (allocation order between
// pairs is irrelevant: the solver walks pairs, not the buffer), and the pair is appended to the work list of its
// dependency level.
__global__ void __launch_bounds__(RP_MANIFOLD_THREADS, RP_MINB_MANIFOLD) k_manifold(DevView d) {
	const unsigned int nh = *d.hit_count;
	ManifoldScratch sc;
	__shared__ int s_cnt[2 * RP_LVL_SMEM], s_base[2 * RP_LVL_SMEM];
	int made = 0;
	const int lane = threadIdx.x & 31;
	for (unsigned int h0 = blockIdx.x * blockDim.x; h0 < nh; h0 += gridDim.x * blockDim.x) {
		const unsigned int hi = h0 + threadIdx.x;
		int n = 0, lvl = -1, w = 0, pair = 0;
		if (hi < nh) {
			const EpaOut eo = d.epa_out[hi];
			const uint2 cd = d.cands[d.hits[hi]];
			w = (int)cd.x; pair = (int)cd.y;
			const size_t pg = (size_t)w * d.max_pairs + pair;
			const PairRec pr = d.pairs[pg];
			int st = 0;
			if (eo.ok) {
				const V3* tv = d.tv + (size_t)w * d.TV;
				const V3* tn = d.tn + (size_t)w * d.TN;
				Shape A = make_shape(d.pool, d.cols[pr.ca], tv, tn);
				Shape B = make_shape(d.pool, d.cols[pr.cb], tv, tn);
				StageSink sink;
				sink.stage = sc.stage;
				sink.n = 0;
				sink.cap = RP_CLIP_MAX_POINTS;
				manifold(A, B, eo.normal, eo.depth, sc.clip, &st, sink);
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
					out[k] = make_contact(b1, b2, sc.stage[2 * k], sc.stage[2 * k + 1]);
					if (w == d.dbg_world) {
						d.dbg_points[2 * (off + k)] = sc.stage[2 * k];
						d.dbg_points[2 * (off + k) + 1] = sc.stage[2 * k + 1];
					}
				}
				d.pair_normal[pg] = eo.normal;
				d.pair_coff[pg] = off;
				d.pair_ccnt[pg] = n;
				made += n;
				if (n > 0) lvl = d.pair_level[pg];
			}
			if (st) atomicOr(&d.status[w], st);
		}
		// append (world, pair) to the list of its level: ranks within the CTA through shared-memory counters, then ONE
		// global atomic per (CTA, level, end) on a counter line of its own (RP_LVL_STRIDE). A level's list can be filled
		// from both ends -- small manifolds (<= RP_SMALL_MANIFOLD contacts) from the front, large ones from the back --
		// which groups manifolds of similar length (order within a level is free: its units commute).
		__syncthreads();
		if (threadIdx.x < 2 * RP_LVL_SMEM) s_cnt[threadIdx.x] = 0;  // RP_MANIFOLD_THREADS >= 2 * RP_LVL_SMEM
		__syncthreads();
		int rank = 0;
		const int big = n > RP_SMALL_MANIFOLD ? 1 : 0;
		if (lvl > 0) {
			if (lvl < RP_LVL_SMEM) rank = atomicAdd(&s_cnt[2 * lvl + big], 1);
			else rank = atomicAdd(&d.lvl_fill[(size_t)lvl * RP_LVL_STRIDE + big], 1);  // very deep schedules: direct
		}
		__syncthreads();
		if (threadIdx.x < 2 * RP_LVL_SMEM && s_cnt[threadIdx.x] > 0) {
			s_base[threadIdx.x] = atomicAdd(&d.lvl_fill[(size_t)(threadIdx.x >> 1) * RP_LVL_STRIDE + (threadIdx.x & 1)], s_cnt[threadIdx.x]);
		}
		__syncthreads();
		if (lvl > 0) {
			const int slot = (lvl < RP_LVL_SMEM ? s_base[2 * lvl + big] : 0) + rank;
			const int at = big ? d.lvl_off[lvl + 1] - 1 - slot : d.lvl_off[lvl] + slot;
			d.lvl_items[at] = make_uint2((unsigned int)w, (unsigned int)pair);
		}
	}
	for (int o = 16; o > 0; o >>= 1) made += __shfl_down_sync(0xffffffffu, made, o);
	if (lane == 0 && made) atomicAdd(&d.counters[CNT_CONTACTS], (unsigned long long)made);
}

// -------------------------------------------------------------------------------------------------------------- solve
// Level-major Gauss-Seidel across ALL worlds: launch l runs every constraint of dependency level l, one thread per
// unit -- the joints of level l of every world, then the (world, pair) items of level l that have contacts this substep
// (a pair's manifold is a sequential chain on its two bodies and stays in one thread, bodies in registers). Kernel
// boundaries are the barriers between levels, so the result equals the reference's sequential sweep (pbd.cpp:615-620).
__global__ void __launch_bounds__(128, RP_MINB_POS) k_pos_level(DevView d, double h, int level, int collisions) {
	const int nj = level <= d.joint_levels ? d.joint_lptr[level] - d.joint_lptr[level - 1] : 0;
	const int njw = nj * d.W;
	int st = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < njw; i += gridDim.x * blockDim.x) {
		const int w = i / nj;
		const int u = d.joint_sched[d.joint_lptr[level - 1] + i % nj];
		const Joint j = d.joints[u];
		BodyDyn* dyn = d.dyn + (size_t)w * d.NB;
		Body b1, b2;
		load_static(b1, d, j.e1);
		load_static(b2, d, j.e2);
		BodyDyn& d1 = dyn[j.e1];
		BodyDyn& d2 = dyn[j.e2];
		b1.x = ld3(d1.x); b1.q = ld4(d1.q);
		b2.x = ld3(d2.x); b2.q = ld4(d2.q);
		JointLambda lam = d.lambdas[(size_t)w * d.NJ + u];
		solve_joint(j, lam, b1, b2, h, &st);
		d.lambdas[(size_t)w * d.NJ + u] = lam;
		if (!b1.fixed) { st3(d1.x, b1.x); st4(d1.q, b1.q); }
		if (!b2.fixed) { st3(d2.x, b2.x); st4(d2.q, b2.q); }
		if (st) {
			atomicOr(&d.status[w], st);
			st = 0;
		}
	}
	if (!collisions) return;
	// contacts: refill loop, one trip = one contact of whatever manifold a lane holds (see WarpQueue)
	const int npf = d.lvl_fill[(size_t)level * RP_LVL_STRIDE];
	const int np = npf + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1];
	const int off0 = d.lvl_off[level], off1 = d.lvl_off[level + 1];
	WarpQueue q;
	q.init((unsigned int)np);
	bool have = false;
	int w = 0, cnt = 0, c = 0;
	Contact* cs = 0;
	BodyDyn* d1 = 0;
	BodyDyn* d2 = 0;
	V3 normal = v3(0.0, 0.0, 0.0);
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			const int k = (int)got;
			const uint2 item = d.lvl_items[k < npf ? off0 + k : off1 - 1 - (k - npf)];
			w = (int)item.x;
			const size_t pg = (size_t)w * d.max_pairs + item.y;
			cnt = d.pair_ccnt[pg];
			const PairRec pr = d.pairs[pg];
			normal = d.pair_normal[pg];
			cs = d.contacts + (size_t)w * d.max_contacts + d.pair_coff[pg];
			load_static(b1, d, pr.a);
			load_static(b2, d, pr.b);
			d1 = d.dyn + (size_t)w * d.NB + pr.a;
			d2 = d.dyn + (size_t)w * d.NB + pr.b;
			b1.x = ld3(d1->x); b1.q = ld4(d1->q); b1.px = ld3(d1->px); b1.pq = ld4(d1->pq);
			b2.x = ld3(d2->x); b2.q = ld4(d2->q); b2.px = ld3(d2->px); b2.pq = ld4(d2->pq);
			c = 0;
			have = cnt > 0;
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			Contact ct = cs[c];
			solve_contact(ct, normal, b1, b2, h, &st);
			cs[c].lambda_n = ct.lambda_n;
			cs[c].lambda_t = ct.lambda_t;
			if (++c == cnt) {
				if (!b1.fixed) { st3(d1->x, b1.x); st4(d1->q, b1.q); }
				if (!b2.fixed) { st3(d2->x, b2.x); st4(d2->q, b2.q); }
				if (st) {
					atomicOr(&d.status[w], st);
					st = 0;
				}
				have = false;
			}
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
	const int npf = d.lvl_fill[(size_t)level * RP_LVL_STRIDE];
	const int np = npf + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1];
	const int off0 = d.lvl_off[level], off1 = d.lvl_off[level + 1];
	WarpQueue q;
	q.init((unsigned int)np);
	bool have = false;
	int cnt = 0, c = 0;
	const Contact* cs = 0;
	BodyDyn* d1 = 0;
	BodyDyn* d2 = 0;
	V3 normal = v3(0.0, 0.0, 0.0);
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			const int k = (int)got;
			const uint2 item = d.lvl_items[k < npf ? off0 + k : off1 - 1 - (k - npf)];
			const int w = (int)item.x;
			const size_t pg = (size_t)w * d.max_pairs + item.y;
			cnt = d.pair_ccnt[pg];
			const PairRec pr = d.pairs[pg];
			normal = d.pair_normal[pg];
			cs = d.contacts + (size_t)w * d.max_contacts + d.pair_coff[pg];
			load_static(b1, d, pr.a);
			load_static(b2, d, pr.b);
			d1 = d.dyn + (size_t)w * d.NB + pr.a;
			d2 = d.dyn + (size_t)w * d.NB + pr.b;
			b1.q = ld4(d1->q); b1.v = ld3(d1->v); b1.w = ld3(d1->w); b1.pv = ld3(d1->pv); b1.pw = ld3(d1->pw);
			b2.q = ld4(d2->q); b2.v = ld3(d2->v); b2.w = ld3(d2->w); b2.pv = ld3(d2->pv); b2.pw = ld3(d2->pw);
			c = 0;
			have = cnt > 0;
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			const Contact ct = cs[c];
			solve_contact_velocity(ct, normal, b1, b2, h);
			if (++c == cnt) {
				if (!b1.fixed) { st3(d1->v, b1.v); st3(d1->w, b1.w); }
				if (!b2.fixed) { st3(d2->v, b2.v); st3(d2->w, b2.w); }
				have = false;
			}
		}
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
