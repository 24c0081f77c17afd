// rp_kernels.cuh -- CUDA kernels of the frame step (sm_100a). Included once, by rp_batch.cu.
//
// Arithmetic is FP64 and is the shared core (rp_math.h ... rp_solve.h) compiled with --fmad=false; the kernels only
// decide WHO computes WHAT and WHEN. Order sensitivity of the reference's sequential Gauss-Seidel is preserved by the
// dependency-level schedule built in k_schedule: two constraints commute exactly when they share no non-fixed body
// (fixed bodies are never written, pbd_base_constraints.cpp:73-103), so running each level's units in parallel and the
// levels in sequence reproduces the sequential result bit for bit.
#ifndef RP_KERNELS_CUH
#define RP_KERNELS_CUH

#include <cooperative_groups.h>

#include "rp_device.cuh"

// occupancy knobs (min resident CTAs per SM handed to __launch_bounds__); tuned on B200, see profiles/
#ifndef RP_MINB_INTEGRATE
#define RP_MINB_INTEGRATE 6
#endif
#ifndef RP_MINB_GJK
#define RP_MINB_GJK 8
#endif
#ifndef RP_MINB_MANIFOLD
#define RP_MINB_MANIFOLD 4
#endif
#ifndef RP_MINB_EPA
#define RP_MINB_EPA 5
#endif
#ifndef RP_MINB_POS
#define RP_MINB_POS 3
#endif
#ifndef RP_MINB_VEL
#define RP_MINB_VEL 4
#endif
#ifndef RP_POS_THREADS
#define RP_POS_THREADS 128  // CTA sizes of the cooperative sweeps (with RP_MINB_*: resident warps, hence trips per level)
#endif
#ifndef RP_VEL_THREADS
#define RP_VEL_THREADS 128
#endif
#ifdef RP_POS_MAXNREG  // tuning: an explicit register cap instead of a minimum CTA count
#define RP_POS_BOUNDS __maxnreg__(RP_POS_MAXNREG)
#else
#define RP_POS_BOUNDS __launch_bounds__(RP_POS_THREADS, RP_MINB_POS)
#endif

#define RP_GJK_THREADS 64
#define RP_GJK_STAGE 48          // doubles of shared memory per thread: two hulls of up to 16 vertices in total
// Three sizes of collider pair. Up to RP_GJK_STAGE / 3 vertices in total: one thread per pair, both hulls staged in shared memory
// (k_gjk). Up to RP_WARP_PAIR_VERTS: still one thread per pair, vertices evaluated from the pose at every use (icosahedra, the
// 58-vertex spot hulls against a box: a warp per pair would idle most of its lanes in every scan). Above: one WARP per pair
// (k_gjk_warp, k_epa_warp, k_manifold_warp: cylinders with 128 vertices, 64-gon faces).
#ifndef RP_WARP_PAIR_VERTS
#define RP_WARP_PAIR_VERTS 72
#endif
__host__ __device__ __forceinline__ bool warp_pair_verts(int total_vertices) { return total_vertices > RP_WARP_PAIR_VERTS; }
#define RP_MANIFOLD_THREADS 128
#define RP_EPA_THREADS 64

#define RP_LVL_STRIDE 32   // ints between consecutive level fill counters (one 128-byte line each: [0] front, [1] back)
#define RP_FLOW_LEVELS 62  // deepest schedule the dataflow sweeps take (their per-level tables sit in static shared memory)
#ifndef RP_SMALL_MANIFOLD
#define RP_SMALL_MANIFOLD 1000000  // (off: the refill loops of the solver kernels make manifold length irrelevant)
// manifolds of up to this many contacts fill a level list from the front, larger ones from the back
#endif

namespace cg = cooperative_groups;

namespace rp {

// ---------------------------------------------------------------------------------------------------------- layout
// Everything that is per world is stored WORLD-MINOR: element e of world w lives at base[e * WS + w] (WS = number of
// worlds rounded up to a multiple of 32). The kernels map lane -> world, so the 32 lanes of a warp -- the same body, the
// same pair index or the same contact slot in 32 neighbouring worlds -- read 32 consecutive doubles: every per-lane
// gather of the world-major layout (32 wavefronts per load instruction; ncu, round 1: k_manifold at 68 % of the LSU
// wavefront peak with 24 stall cycles on the long scoreboard per issued instruction) becomes one coalesced request.
enum { DF_X = 0, DF_Q = 3, DF_V = 7, DF_W = 10, DF_PX = 13, DF_PQ = 16, DF_PV = 20, DF_PW = 23 };  // fields of a body's dynamic record

struct DynRef {  // one body's dynamic record: field component f at p[f * s]
	real* p;
	size_t s;
};
__device__ __forceinline__ DynRef dyn_ref(const DevView& d, int w, int b) {
	DynRef r;
	r.p = d.dyn + (size_t)b * RP_DYN_DOUBLES * d.WS + w;
	r.s = d.WS;
	return r;
}
__device__ __forceinline__ V3 ld3(const DynRef& r, int f) { return v3(r.p[f * r.s], r.p[(f + 1) * r.s], r.p[(f + 2) * r.s]); }
__device__ __forceinline__ Q4 ld4(const DynRef& r, int f) { return q4(r.p[f * r.s], r.p[(f + 1) * r.s], r.p[(f + 2) * r.s], r.p[(f + 3) * r.s]); }
// the same through L2 only (ld.global.cg): for state that units running on OTHER SMs write during the same launch (dataflow sweeps),
// where a line cached in this SM's L1 by an earlier unit would be stale
__device__ __forceinline__ V3 ld3cg(const DynRef& r, int f) { return v3(__ldcg(r.p + f * r.s), __ldcg(r.p + (f + 1) * r.s), __ldcg(r.p + (f + 2) * r.s)); }
__device__ __forceinline__ Q4 ld4cg(const DynRef& r, int f) {
	return q4(__ldcg(r.p + f * r.s), __ldcg(r.p + (f + 1) * r.s), __ldcg(r.p + (f + 2) * r.s), __ldcg(r.p + (f + 3) * r.s));
}
__device__ __forceinline__ void st3(const DynRef& r, int f, V3 v) { r.p[f * r.s] = v.x; r.p[(f + 1) * r.s] = v.y; r.p[(f + 2) * r.s] = v.z; }
__device__ __forceinline__ void st4(const DynRef& r, int f, Q4 q) {
	r.p[f * r.s] = q.x; r.p[(f + 1) * r.s] = q.y; r.p[(f + 2) * r.s] = q.z; r.p[(f + 3) * r.s] = q.w;
}
// contact record (r1_lc, r2_lc, lambda_n, lambda_t) of slot k of a world: component f at p[(k * 8 + f) * WS]
__device__ __forceinline__ real* contact_ptr(const DevView& d, int w, int slot) { return d.contacts + (size_t)slot * 8 * d.WS + w; }
__device__ __forceinline__ Contact ld_contact(const real* p, size_t S) {
	Contact c;
	c.r1_lc = v3(p[0], p[S], p[2 * S]);
	c.r2_lc = v3(p[3 * S], p[4 * S], p[5 * S]);
	c.lambda_n = p[6 * S];
	c.lambda_t = p[7 * S];
	return c;
}
__device__ __forceinline__ void st_contact(real* p, size_t S, const Contact& c) {
	p[0] = c.r1_lc.x; p[S] = c.r1_lc.y; p[2 * S] = c.r1_lc.z;
	p[3 * S] = c.r2_lc.x; p[4 * S] = c.r2_lc.y; p[5 * S] = c.r2_lc.z;
	p[6 * S] = c.lambda_n; p[7 * S] = c.lambda_t;
}
__device__ __forceinline__ size_t bidx(const DevView& d, int b, int w) { return (size_t)b * d.WS + w; }  // per-body arrays
__device__ __forceinline__ size_t pidx(const DevView& d, int p, int w) { return (size_t)p * d.WS + w; }  // per-pair arrays

// dataflow sweeps: a unit of level `lvl` with contacts this substep touches bodies a and b (fixed bodies are never written and carry no chain)
__device__ __forceinline__ void flow_mark_live(const DevView& d, int a, int b, int w, int lvl) {
	if (!d.bstat[a].fixed) atomicOr(&d.body_live[bidx(d, a, w)], 1ull << lvl);
	if (!d.bstat[b].fixed) atomicOr(&d.body_live[bidx(d, b, w)], 1ull << lvl);
}

// Thread -> (item, world) for the kernels that do one thing per item per world, world fastest: consecutive threads are
// consecutive worlds of one item, so world-minor accesses coalesce; a batch of fewer worlds than a warp (one large scene: WS = W
// then, see rp_batch_create) gets consecutive ITEMS in a warp instead of 31 idle lanes per item, and the accesses still
// coalesce because its world stride is W. Launch flat_grid() CTAs: ceil(n_items * W / blockDim.x), or (items, world blocks) for
// batches of at least two CTAs' worth of worlds (consecutive CTAs then walk the items of one world block, as round 1 did).
__device__ __forceinline__ bool flat_item_world(const DevView& d, int n_items, int* item, int* w) {
	if (gridDim.y > 1) {  // batches of whole CTAs of worlds: grid = (items, world blocks), a CTA = one item in blockDim.x worlds
		*item = (int)blockIdx.x;
		*w = (int)(blockIdx.y * blockDim.x + threadIdx.x);
		return *w < d.W;
	}
	const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long total = (unsigned long long)n_items * (unsigned long long)d.W;
	if (t >= total) return false;
	if (total <= 0xffffffffull) {  // (a 64-bit division is a ~100-instruction subroutine; every batch but the very largest fits 32 bits)
		const unsigned int it = (unsigned int)t / (unsigned int)d.W;
		*item = (int)it;
		*w = (int)((unsigned int)t - it * (unsigned int)d.W);
	} else {
		const unsigned long long it = t / (unsigned long long)d.W;
		*item = (int)it;
		*w = (int)(t - it * (unsigned long long)d.W);
	}
	return true;
}

__device__ __forceinline__ void load_static(Body& b, const DevView& d, int body) {
	const BodyStatic& s = d.bstat[body];
	const BodyClass& c = d.bclass[s.cls];
	b.inv_mass = c.inv_mass;
	b.inertia = c.inertia;
	b.inv_inertia = c.inv_inertia;
	b.inv_inertia_p = &c.inv_inertia;
	b.mu_s = c.mu_s; b.mu_d = c.mu_d; b.rest = c.rest;
	b.ii_bound = c.ii_bound;
	b.fixed = s.fixed;
}

// A collider of world w at its current pose, as a PoseShape (rp_shape.h): nothing of its transformed geometry is read from memory -- the body's pose
// (7 doubles, world-minor, coalesced) gives the model matrix, and vertices / face normals are evaluated from the hull's local
// data (template arrays: the same addresses for every lane that works on the same hull, L1-resident) when they are needed.
__device__ __forceinline__ PoseShape dev_pose_shape(const DevView& d, const ColliderDesc& c, int w) {
	PoseShape s;
	s.type = c.type;
	s.radius = c.radius;
	const DynRef r = dyn_ref(d, w, c.body);
	const V3 x = ld3(r, DF_X);
	s.vp = 0; s.vs = 0; s.vcs = 0;
	s.np = d.tn + (size_t)c.tn0 * 3 * d.WS + w; s.ns = 3 * d.WS; s.ncs = d.WS;  // (read only in the RP_STORED_NORMALS build)
	if (c.type == SHAPE_SPHERE) {
		s.center = x;  // collider.cpp:433
		s.nv = 0; s.nf = 0;
		s.face_ptr = s.face_idx = s.v2f_ptr = s.v2f_idx = s.v2n_ptr = s.v2n_idx = s.f2n_ptr = s.f2n_idx = 0;
		s.lv = s.ln = 0;
		s.M = model_matrix(q4(RL(0.0), RL(0.0), RL(0.0), RL(1.0)), x);
	} else {
		const HullTopo h = d.pool.hulls[c.hull];
		s.M = model_matrix(ld4(r, DF_Q), x);
		s.center = v3(RL(0.0), RL(0.0), RL(0.0));
		s.nv = h.nv; s.nf = h.nf;
		s.lv = d.pool.verts + h.vert0;
		s.ln = d.pool.normals + h.face0;
		s.face_ptr = d.pool.face_ptr + h.fptr0; s.face_idx = d.pool.face_idx;
		s.v2f_ptr = d.pool.v2f_ptr + h.v2f0; s.v2f_idx = d.pool.v2f_idx;
		s.v2n_ptr = d.pool.v2n_ptr + h.v2n0; s.v2n_idx = d.pool.v2n_idx;
		s.f2n_ptr = d.pool.f2n_ptr + h.f2n0; s.f2n_idx = d.pool.f2n_idx;
	}
	return s;
}

// ------------------------------------------------------------------------------------------------------- broadphase
// broad_get_collision_pairs (broad.cpp:6-29): all i < j with |x_i - x_j| <= r_i + r_j + 0.1, emitted in (i, j) order.
#ifndef RP_BROAD_BATCH
#define RP_BROAD_BATCH 8
#endif
#define RP_BROAD_SEGS 8
// The (i, j) triangle is cut into CELLS of one row i and up to 32 consecutive j (DevView::cells, in (i, j) order; a
// constant of the template): one thread per (world, cell), lane = world. A row per thread leaves the rows of big bodies
// -- the floor pairs with everything -- as 256-step dependent chains that the whole launch waits for; cells are at most 32
// steps each. Pass 1 leaves a bit mask of the near j of every cell and the number of collider pairs they expand to;
// k_broad_scan turns the counts into offsets (cell order = pair order); pass 2 expands the masks. Nothing is tested twice.
__global__ void __launch_bounds__(256) k_broad_cells(DevView d) {
	int w, c;
	if (!flat_item_world(d, d.n_cells, &c, &w)) return;
	const int2 cell = d.cells[c];
	const int i = cell.x, j0 = cell.y;
	const int j1 = j0 + 32 < d.NB ? j0 + 32 : d.NB;
	const real* X = d.dyn + w;  // x of body j: X[(j * RP_DYN_DOUBLES + c) * WS]
	const size_t S = d.WS;
	const size_t body_stride = (size_t)RP_DYN_DOUBLES * S;
	const V3 xi = v3(X[(size_t)i * body_stride], X[(size_t)i * body_stride + S], X[(size_t)i * body_stride + 2 * S]);
	const real ri = d.bstat[i].radius;
	const int nci = d.bstat[i].ncol;
	// Axis rejects first: |x_i - x_j| along one axis already above (r_i + r_j + 0.1)(1 + 1e-9) puts the distance the
	// reference computes (sqrt of a sum of rounded squares, each term within 3 ulp) above r_i + r_j + 0.1 as well, so the
	// pair is not emitted -- decided after one load and three FP64 instructions instead of fifteen. The threshold is the
	// sum of two per-body terms prepared once (BodyStatic::rfar), both rounded up by 1e-9 >> the 1e-16 roundings involved.
	// Survivors get the reference's own expression below.
	const real ri_far = ri * (RL(1.0) + RL(1e-9));
	unsigned int mask = 0u;
	int count = 0;
	// the x coordinates of RP_BROAD_BATCH bodies are fetched together (independent loads in flight), then looked at in order
	for (int jb = j0; jb < j1; jb += RP_BROAD_BATCH) {
		real xs[RP_BROAD_BATCH], thrs[RP_BROAD_BATCH];
#pragma unroll
		for (int k = 0; k < RP_BROAD_BATCH; ++k) {
			const int jj = jb + k < j1 ? jb + k : j1 - 1;
			xs[k] = X[(size_t)jj * body_stride];
			thrs[k] = ri_far + d.bstat[jj].rfar;
		}
#pragma unroll
		for (int k = 0; k < RP_BROAD_BATCH; ++k) {
			const int j = jb + k;
			if (j >= j1) break;
			const real thr = thrs[k];
			const real* xj = X + (size_t)j * body_stride;
			V3 dv;
			dv.x = xi.x - xs[k];
			if (fabs(dv.x) > thr) continue;
			dv.z = xi.z - xj[2 * S];
			if (fabs(dv.z) > thr) continue;
			dv.y = xi.y - xj[S];
			if (fabs(dv.y) > thr) continue;
			// broad.cpp:19-20 compares sqrt(|xi - xj|^2) with ri + rj + 0.1. Squared distances outside a 4e-12 relative
			// band around maxd^2 decide the comparison without the square root (sqrt is monotonic and both roundings are
			// 1e-16 effects); inside the band the reference's expression is evaluated as written.
			const real d2 = dv.x * dv.x + dv.y * dv.y + dv.z * dv.z;
			const real maxd = ri + d.bstat[j].radius + RL(0.1);
			const real m2 = maxd * maxd;
			bool near = d2 < m2 * (RL(1.0) - RL(4e-12));
			if (!near && !(d2 > m2 * (RL(1.0) + RL(4e-12)))) near = sqrt(d2) <= maxd;
			if (near) {
				mask |= 1u << (j - j0);
				count += nci * d.bstat[j].ncol;
			}
		}
	}
	d.cell_mask[(size_t)c * S + w] = mask;
	d.cell_off[(size_t)c * S + w] = count;
}

// exclusive scan of the cell counts of every world: thread = (world, segment of the cell list); blockDim = (32, RP_BROAD_SEGS)
__global__ void __launch_bounds__(32 * RP_BROAD_SEGS) k_broad_scan(DevView d) {
	const int w = blockIdx.x * 32 + threadIdx.x;
	const int seg = threadIdx.y;
	const bool live = w < d.W;
	const size_t S = d.WS;
	int* off = d.cell_off + (live ? w : 0);
	const int per = (d.n_cells + RP_BROAD_SEGS - 1) / RP_BROAD_SEGS;
	const int c0 = seg * per < d.n_cells ? seg * per : d.n_cells;
	const int c1 = c0 + per < d.n_cells ? c0 + per : d.n_cells;
	__shared__ int s_sum[RP_BROAD_SEGS][32];
	int sum = 0;
	if (live) {
#pragma unroll 8
		for (int c = c0; c < c1; ++c) sum += off[(size_t)c * S];
	}
	s_sum[seg][threadIdx.x] = sum;
	__syncthreads();
	int run = 0, total = 0;
	for (int k = 0; k < RP_BROAD_SEGS; ++k) {
		if (k < seg) run += s_sum[k][threadIdx.x];
		total += s_sum[k][threadIdx.x];
	}
	if (live) {
		for (int cb = c0; cb < c1; cb += 8) {
			int v[8];
#pragma unroll
			for (int k = 0; k < 8; ++k) v[k] = cb + k < c1 ? off[(size_t)(cb + k) * S] : 0;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				if (cb + k < c1) off[(size_t)(cb + k) * S] = run;
				run += v[k];
			}
		}
		if (seg == 0) {
			if (total > d.max_pairs) {
				atomicOr(&d.status[w], ST_PAIR_CAPACITY);
				total = d.max_pairs;
			}
			d.n_pairs[w] = total;
		}
	}
	if (seg == 0) {
		unsigned long long t = live ? (unsigned long long)total : 0ull;
		for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
		if (threadIdx.x == 0 && t) atomicAdd(&d.counters[CNT_BROAD_PAIRS], t);
	}
}

// pass 2: every cell writes the pairs of its mask at its offset (collider pairs of one body pair stay adjacent, sub-collider
// i outer, j inner: collider.cpp:563-571)
__global__ void __launch_bounds__(256) k_broad_write(DevView d) {
	int w, c;
	if (!flat_item_world(d, d.n_cells, &c, &w)) return;
	unsigned int mask = d.cell_mask[(size_t)c * d.WS + w];
	if (!mask) return;
	int out = d.cell_off[(size_t)c * d.WS + w];
	const int2 cell = d.cells[c];
	const int i = cell.x;
	const int ci0 = d.bstat[i].col0, nci = d.bstat[i].ncol;
	while (mask) {
		const int j = cell.y + __ffs(mask) - 1;
		mask &= mask - 1u;
		const int cj0 = d.bstat[j].col0, ncj = d.bstat[j].ncol;
		for (int a = 0; a < nci; ++a) {
			for (int b = 0; b < ncj; ++b) {
				if (out < d.max_pairs) {
					PairRec pr;
					pr.a = i; pr.b = j; pr.ca = ci0 + a; pr.cb = cj0 + b;
					d.pairs[pidx(d, out, w)] = pr;
				}
				++out;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------- islands + sleeping
// broad_collect_simulation_islands (broad.cpp:70-116) + the sleep bookkeeping of pbd.cpp:476-506. Islands are the
// connected components of {pairs, external constraints} restricted to non-fixed bodies; the result does not depend on
// the order in which unions happen, so min-label propagation replaces the reference's union-find. One CTA per world
// (labels and flags are [W][NB] scratch; the strided reads of the world-minor arrays are once per frame).
__global__ void __launch_bounds__(256) k_islands(DevView d, real dt) {
	const int w = blockIdx.x;
	int* label = d.label + (size_t)w * d.NB;
	int* flag = d.isl_flag + (size_t)w * d.NB;
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
				const PairRec pr = d.pairs[pidx(d, e, w)];
				a = pr.a; b = pr.b;
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
		const DynRef r = dyn_ref(d, w, b);
		real lv = length(ld3(r, DF_V));
		real av = length(ld3(r, DF_W));
		real t = d.deact[bidx(d, b, w)];
		if (lv < d.lin_sleep && av < d.ang_sleep) t += dt;
		else t = RL(0.0);
		d.deact[bidx(d, b, w)] = t;
		if (t < d.sleep_time) flag[label[b]] = 0;
	}
	__syncthreads();
	for (int b = threadIdx.x; b < d.NB; b += blockDim.x) {
		if (d.bstat[b].fixed) continue;
		d.active[bidx(d, b, w)] = flag[label[b]] ? 0 : 1;
	}
}

// How k_schedule places a unit given the table entries of its two bodies (0 for a fixed body), and what it leaves in them.
template <bool COLOUR>
struct SchedEntry;
template <>
struct SchedEntry<false> {  // dependency levels: entry = level of the last unit that touched the body
	typedef int type;
	__device__ static __forceinline__ int* scratch(const DevView& d) { return d.last_level; }
	__device__ static __forceinline__ int initial(const DevView& d, int body) { return d.joint_last[body]; }
	__host__ __device__ static __forceinline__ int place(int la, int lb, int* na, int* nb) {
		const int lvl = 1 + (la > lb ? la : lb);
		*na = lvl; *nb = lvl;
		return lvl;
	}
};
template <>
struct SchedEntry<true> {  // colours: entry = mask of colours 1..32 seen (low word) | overflow depth (high word)
	typedef unsigned long long type;
	__device__ static __forceinline__ unsigned long long* scratch(const DevView& d) { return d.colour_tab; }
	__device__ static __forceinline__ unsigned long long initial(const DevView& d, int body) { return d.joint_colours[body]; }
	__host__ __device__ static __forceinline__ int place(unsigned long long la, unsigned long long lb, unsigned long long* na,
		unsigned long long* nb) {
		const unsigned int seen = (unsigned int)(la | lb);
		if (seen != 0xffffffffu) {
			int c = 0;
			while ((seen >> c) & 1u) ++c;  // lowest free colour (0-based)
			*na = la | (1ull << c);
			*nb = lb | (1ull << c);
			return c + 1;
		}
		const unsigned long long ea = la >> 32, eb = lb >> 32;
		const unsigned long long e = (ea > eb ? ea : eb) + 1ull;
		*na = (la & 0xffffffffull) | (e << 32);
		*nb = (lb & 0xffffffffull) | (e << 32);
		return 32 + (int)e;
	}
};

// ------------------------------------------------------------------------------------------------------ level schedule
// Units of the Gauss-Seidel sweep in the reference's array order: the external constraints first (copy_constraints
// output is the head of the array, pbd.cpp:580), then the broadphase (collider-)pairs in pair order, each pair standing
// for its whole manifold (pbd.cpp:584-611). level(u) = 1 + max(level of the previous unit touching either of u's
// NON-FIXED bodies). The joints' levels are the same in every world (host, at batch creation); this kernel continues
// the recurrence over one world's pairs (one thread per world, once per frame; all its arrays are world-minor, so the
// 32 worlds of a warp read consecutive words) and adds the world's per-level pair counts to the global capacities of
// the level-major work lists.
// SMEM: the tables the recurrence reads in a dependent chain live in shared memory for the CTA's 32 worlds -- per body
// the level of the last unit that touched it ([NB][32] ints) and its "takes part" flags (fixed / active, [NB][32]
// bytes), and the per-world level histogram ([RP_SCHED_HIST][32] ints) -- so an iteration's only global access is the
// (coalesced, prefetchable) pair record. Scenes too large for that use the world-minor global scratch.
// COLOUR = true is the large-scene order (rp_batch_cfg.solve_order = RP_ORDER_COLOURED): instead of the dependency level
// a unit gets the lowest COLOUR no earlier unit on either of its non-fixed bodies has taken -- a greedy colouring of the
// constraint graph, rebuilt every frame from that frame's pairs. Units of one colour share no non-fixed body and run in
// parallel exactly like the units of one level; the sweep walks the colours in ascending order. That is a Gauss-Seidel
// order too, but not the reference's: a brick wall needs ~10 colours where the reference's array order chains ~150
// levels deep, so a single large scene takes an order of magnitude fewer grid-wide barriers -- and its trajectories
// agree with the reference's only within solver accuracy, not bit for bit. A body's table entry holds a 32-bit mask of
// the colours it has seen (low word) and, for bodies with more than 32 units, a counter that continues with the
// dependency recurrence above colour 32 (high word).
#define RP_SCHED_HIST 64
template <bool SMEM, bool COLOUR>
__global__ void __launch_bounds__(32) k_schedule(DevView d, int collisions) {
	typedef typename SchedEntry<COLOUR>::type Entry;
	extern __shared__ __align__(8) unsigned char s_sched_raw[];
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = w < d.W;
	const size_t S = d.WS;
	// shared tables are [..][SL] with SL = the lanes that hold a world: 32, or the whole batch when it is smaller than a warp
	// (one large scene must not pay 32 columns for its one world -- its tables would not fit)
	const int SL = d.W < 32 ? d.W : 32;
	Entry* s_tab = reinterpret_cast<Entry*>(s_sched_raw);
	Entry* last = SMEM ? s_tab + (live ? threadIdx.x : 0) : SchedEntry<COLOUR>::scratch(d) + (live ? w : 0);  // [NB][SL] or [NB][WS]
	const size_t LS = SMEM ? SL : S;
	int* s_sched = reinterpret_cast<int*>(s_tab + (size_t)d.NB * SL);
	int* s_hist = s_sched + (live ? threadIdx.x : 0);                                          // [RP_SCHED_HIST][SL] (SMEM only)
	unsigned char* s_flag = (unsigned char*)(s_sched + RP_SCHED_HIST * SL) + (live ? threadIdx.x : 0);  // [NB][SL] (SMEM only)
	int* plevel = d.pair_level + (live ? w : 0);    // [max_pairs][WS]
	int* hist = d.lvl_hist + (live ? w : 0);        // [max_levels + 2][WS]
	const int* active = d.active + (live ? w : 0);  // [NB][WS]
	const int np = (collisions && live) ? d.n_pairs[w] : 0;
	// flag bit 0: fixed, bit 1: fixed or asleep (pbd.cpp:594)
	if (live) {
		for (int b = 0; b < d.NB; ++b) {
			last[b * LS] = SchedEntry<COLOUR>::initial(d, b);
			if (SMEM) {
				const int f = d.bstat[b].fixed;
				s_flag[b * SL] = (unsigned char)((f ? 1 : 0) | ((f || !active[b * S]) ? 2 : 0));
			}
		}
		if (SMEM) {
			for (int l = 0; l < RP_SCHED_HIST; ++l) s_hist[l * SL] = 0;
		}
	}
	int nl = d.joint_levels;
	int deep = 0;  // pairs scheduled at levels >= RP_SCHED_HIST (counted through the global histogram)
	// the recurrence is a dependent chain through shared memory; its only global reads, the pair records, are fetched
	// eight at a time ahead of it (one thread per world has no other way to keep several loads in flight)
	for (int p0 = 0; p0 < np; p0 += 8) {
		int2 ab[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if (p0 + k < np) ab[k] = *reinterpret_cast<const int2*>(&d.pairs[pidx(d, p0 + k, w)]);
		}
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const int p = p0 + k;
			if (p >= np) break;
			const int a = ab[k].x, b = ab[k].y;
			int fa, fb, sa, sb;
			if (SMEM) {
				const int ga = s_flag[a * SL], gb = s_flag[b * SL];
				fa = ga & 1; fb = gb & 1; sa = ga & 2; sb = gb & 2;
			} else {
				fa = d.bstat[a].fixed; fb = d.bstat[b].fixed;
				sa = fa || !active[a * S]; sb = fb || !active[b * S];
			}
			// pbd.cpp:594: nothing to do when both sides are fixed or asleep
			if (sa && sb) {
				plevel[p * S] = 0;
				continue;
			}
			const Entry la = fa ? 0 : last[a * LS], lb = fb ? 0 : last[b * LS];
			Entry na, nb;
			const int lvl = SchedEntry<COLOUR>::place(la, lb, &na, &nb);
			if (!fa) last[a * LS] = na;
			if (!fb) last[b * LS] = nb;
			plevel[p * S] = lvl;
			if (lvl > nl) nl = lvl;
			if (SMEM) {
				if (lvl < RP_SCHED_HIST) s_hist[lvl * SL] += 1;
				else ++deep;
			}
		}
	}
	// per-level pair counts of the warp's 32 worlds -> one atomic per (warp, level): per-thread atomics on the handful of
	// lvl_cap words serialise (4096 worlds x 14 levels on 14 addresses cost 0.6 ms per frame)
	int nl_warp = live ? nl : 0;
	for (int o = 16; o > 0; o >>= 1) nl_warp = max(nl_warp, __shfl_xor_sync(0xffffffffu, nl_warp, o));
	const int deep_any = __any_sync(0xffffffffu, deep != 0);
	const bool use_smem_hist = SMEM && !deep_any;
	if (!use_smem_hist && live) {
		for (int l = 0; l <= nl + 1; ++l) hist[l * S] = 0;
		for (int p = 0; p < np; ++p) hist[plevel[p * S] * S] += 1;
	}
	for (int l = 1; l <= nl_warp; ++l) {
		int c = 0;
		if (live && l <= nl) c = use_smem_hist ? s_hist[l * SL] : hist[l * S];
		c = __reduce_add_sync(0xffffffffu, c);
		if (threadIdx.x == 0 && c) atomicAdd(&d.lvl_cap[l], c);
	}
	const int nl_sum = __reduce_add_sync(0xffffffffu, live ? nl : 0);
	if (threadIdx.x == 0) {
		atomicMax(d.lvl_max, nl_warp);
		atomicAdd(&d.counters[CNT_LEVELS], (unsigned long long)nl_sum);
	}
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
// per-substep counters
__global__ void __launch_bounds__(256) k_substep_reset(DevView d) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		*d.hit_count = 0u;
		*d.cand_count = 0u;
		*d.big_count = 0u;
		*d.epoch += 1;
	}
	if (i < d.max_levels + 2) {
		d.lvl_fill[(size_t)i * RP_LVL_STRIDE] = 0;
		d.lvl_fill[(size_t)i * RP_LVL_STRIDE + 1] = 0;
	}
	if (i < d.W) {
		d.n_contacts[i] = 0;
		d.n_live[i] = 0;
	}
	if (i < 2 && d.flow_mode) d.flow_cursor[i] = 0u;
}

#define RP_INT_THREADS 128
// float bounds rounded outwards from the arithmetic type (in single precision the value is a float already)
#if defined(RP_REAL_F32)
__device__ __forceinline__ float bound_down(real x) { return x; }
__device__ __forceinline__ float bound_up(real x) { return x; }
#else
__device__ __forceinline__ float bound_down(real x) { return __double2float_rd(x); }
__device__ __forceinline__ float bound_up(real x) { return __double2float_ru(x); }
#endif
// world-space bounds of one collider from its body's pose (collider_update's bounds, see k_integrate / k_bounds)
__device__ __forceinline__ void collider_bounds(const DevView& d, const ColliderDesc& cd, const Pose34& M, V3 x, float* bb, size_t S) {
	if (cd.type == SHAPE_SPHERE) {
		const real rad = (real)cd.radius;
		bb[0] = bound_down(x.x - rad); bb[S] = bound_down(x.y - rad); bb[2 * S] = bound_down(x.z - rad);
		bb[3 * S] = bound_up(x.x + rad); bb[4 * S] = bound_up(x.y + rad); bb[5 * S] = bound_up(x.z + rad);
	} else {
		const HullTopo t = d.pool.hulls[cd.hull];
		real lo0 = RL(RP_REAL_MAX), lo1 = lo0, lo2 = lo0, hi0 = -lo0, hi1 = -lo0, hi2 = -lo0;
		for (int k = 0; k < t.nv; ++k) {
			const V3 p = transform_point(M, d.pool.verts[t.vert0 + k]);
			lo0 = fmin(lo0, p.x); lo1 = fmin(lo1, p.y); lo2 = fmin(lo2, p.z);
			hi0 = fmax(hi0, p.x); hi1 = fmax(hi1, p.y); hi2 = fmax(hi2, p.z);
		}
		bb[0] = bound_down(lo0); bb[S] = bound_down(lo1); bb[2 * S] = bound_down(lo2);
		bb[3 * S] = bound_up(hi0); bb[4 * S] = bound_up(hi1); bb[5 * S] = bound_up(hi2);
	}
}
// pbd.cpp:537-577 (integration) and collider.cpp:409-445 (collider_update) for one body per thread. The reference
// re-transforms both colliders of every pair every substep (39 % of its time); the same pose gives the same result,
// so once per body per substep is exactly equivalent. Also leaves each collider's world-space bounds
// for k_cull. Grid = (bodies, world blocks): a CTA is ONE body in 128 consecutive worlds, so the statics and the hull
// are the same for every lane and every load/store of the world-minor arrays is coalesced.
__global__ void __launch_bounds__(RP_INT_THREADS, RP_MINB_INTEGRATE) k_integrate(DevView d, real h, int store_velocities) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	const int epoch = *d.epoch;
	const size_t S = d.WS;
	if (b < d.NJ) {  // copy_constraints resets every lambda each substep (pbd.cpp:426-462)
		for (int j = b; j < d.NJ; j += d.NB) {
			JointLambda z;
			z.a = z.b = z.c = RL(0.0);
			d.lambdas[(size_t)j * S + w] = z;
		}
	}
	const BodyStatic s = d.bstat[b];
	if (d.flow_mode) d.body_live[bidx(d, b, w)] = 0ull;  // this substep's k_manifold marks the levels at which the body has live units
	Body body;
	load_static(body, d, b);
	const DynRef r = dyn_ref(d, w, b);
	body.x = ld3(r, DF_X); body.q = ld4(r, DF_Q);
	body.active = d.active[bidx(d, b, w)];
	const bool moving = !(body.fixed || !body.active);
	// lazy velocity derivation of the PREVIOUS substep (pbd.cpp:623-643) for a body no velocity-level unit touched
	// there: (x, q, prev x, prev q) are still exactly what the reference's derivation pass would have seen, and the
	// derived velocities depend on nothing else -- the stored velocities of such a body are not even read. (What the
	// derivation would have left as previous velocities is not stored either: nothing reads it before this substep's
	// derivation rewrites it.)
	if (moving && d.vstamp[bidx(d, b, w)] != epoch - 1) {
		body.v = body.w = v3(RL(0.0), RL(0.0), RL(0.0));
		body.px = ld3(r, DF_PX); body.pq = ld4(r, DF_PQ);
		derive_velocity(body, h);
	} else {
		body.v = ld3(r, DF_V); body.w = ld3(r, DF_W);
	}
	integrate(body, h, d.force[b], d.torque[b]);
	st3(r, DF_PX, body.px); st4(r, DF_PQ, body.pq);
	if (moving) {
		st3(r, DF_X, body.x); st4(r, DF_Q, body.q);
		// The velocities as integrated are read by one thing only: this substep's derivation turns them into the body's
		// PREVIOUS velocities (pbd.cpp:629-630), which the velocity pass reads for restitution and the caller sees after
		// the frame. In a scene without restitution (every coefficient zero: solve_contact_velocity never looks at them)
		// they are dead until the last substep, and the host asks for them to be stored only there.
		if (store_velocities) { st3(r, DF_V, body.v); st3(r, DF_W, body.w); }
	}
	// collider_update (collider.cpp:409-445) is split in two. Here: the world-space bounds of every collider of the body
	// (from the transformed vertices, which stay in registers) for k_cull. The transformed vertices and face normals
	// themselves are written by k_transform, after the cull, and only for colliders that are part of a surviving
	// candidate pair: geometry nothing will look at is not written (it is the larger half of what this kernel, which runs
	// at the HBM roof, used to store).
	if (d.split_bounds) return;  // k_bounds does it per collider
	// bounds are kept in float, lower ends rounded down and upper ends up: still boxes AROUND the vertex sets, so the cull
	// stays exact-safe (it can only keep a pair it might have dropped), at half the bytes written here and read there
	const Pose34 M = model_matrix(body.q, body.x);
	for (int c = s.col0; c < s.col0 + s.ncol; ++c) collider_bounds(d, d.cols[c], M, body.x, d.aabb + (size_t)c * 6 * S + w, S);
}

// the bounds of every collider, one thread per (collider, world): for scenes whose bodies carry so many vertices (compound
// bodies of large hulls) that k_integrate's one thread per body would walk thousands of them
__global__ void __launch_bounds__(RP_INT_THREADS) k_bounds(DevView d) {
	int w, c;
	if (!flat_item_world(d, d.NC, &c, &w)) return;
	const ColliderDesc cd = d.cols[c];
	const DynRef r = dyn_ref(d, w, cd.body);
	const V3 x = ld3(r, DF_X);
	const Pose34 M = model_matrix(ld4(r, DF_Q), x);
	collider_bounds(d, cd, M, x, d.aabb + (size_t)c * 6 * d.WS + w, d.WS);
}

// Only in the RP_STORED_NORMALS build (a tuning variant; by default the narrowphase evaluates face normals from the pose as it
// does vertices, see PoseShape): re-normalised face normals (collider.cpp:425-429) of the colliders k_cull marked
// (geom_stamp == this substep), after k_cull. Thread = (collider, world, slice): lane = world; gridDim.z slices share out the
// normals of large hulls.
__global__ void __launch_bounds__(RP_INT_THREADS) k_transform(DevView d) {
	const int w = blockIdx.y * RP_INT_THREADS + threadIdx.x;
	const int c = blockIdx.x;
	if (w >= d.W) return;
	const size_t S = d.WS;
	if (d.geom_stamp[(size_t)c * S + w] != *d.epoch) return;
	const ColliderDesc cd = d.cols[c];
	if (cd.type == SHAPE_SPHERE) return;
	const DynRef r = dyn_ref(d, w, cd.body);
	const Pose34 M = model_matrix(ld4(r, DF_Q), ld3(r, DF_X));
	const int first = blockIdx.z, step = gridDim.z;
	const HullTopo t = d.pool.hulls[cd.hull];
	real* tn = d.tn + (size_t)cd.tn0 * 3 * S + w;
	for (int k = first; k < t.nf; k += step) {
		const V3 n = transform_normal(M, d.pool.normals[t.face0 + k]);
		tn[(size_t)(3 * k) * S] = n.x; tn[(size_t)(3 * k + 1) * S] = n.y; tn[(size_t)(3 * k + 2) * S] = n.z;
	}
}

// Copies a small hull's transformed vertices (and, optionally, face normals) from the world-minor arrays into the
// calling thread's column of a thread-interleaved shared-memory block: element e of the thread lives at
// base[e * nthreads], so the 32 lanes of a warp touch 32 consecutive doubles per access whatever (world, body) each lane
// holds. The narrowphase scans the same vertices many times (support mapping), and 16 vertices x 3 x 32 lanes of a warp
// are 12 kB -- more than a warp's share of L1. Returns the number of doubles used.
__device__ __forceinline__ int stage_shape(PoseShape& s, real* col, int nthreads) {
	// the vertices are evaluated from the pose (collider_update's expression, collider.cpp:414-422), not loaded
	for (int k = 0; k < s.nv; ++k) {
		const V3 p = vert(s, k);
		col[(size_t)(3 * k) * nthreads] = p.x; col[(size_t)(3 * k + 1) * nthreads] = p.y; col[(size_t)(3 * k + 2) * nthreads] = p.z;
	}
	s.vp = col; s.vs = 3 * nthreads; s.vcs = nthreads;
	return s.nv * 3;
}

// final GJK tetrahedron of hit `slot`, stored as 12 component planes of `cand_cap` doubles: the lanes of a warp hold
// consecutive hits, so every access is one coalesced request (as 4 x V3 records per hit the twelve loads of k_epa each touched
// 32 different sectors: 9 % of its stall samples, ncu round 2)
__device__ __forceinline__ void st_simplex(const DevView& d, unsigned int slot, const Simplex& s) {
	real* o = d.simplex + slot;
	const size_t S = d.cand_cap;
	o[0] = s.a.x; o[S] = s.a.y; o[2 * S] = s.a.z; o[3 * S] = s.b.x; o[4 * S] = s.b.y; o[5 * S] = s.b.z;
	o[6 * S] = s.c.x; o[7 * S] = s.c.y; o[8 * S] = s.c.z; o[9 * S] = s.d.x; o[10 * S] = s.d.y; o[11 * S] = s.d.z;
}
__device__ __forceinline__ Simplex ld_simplex(const DevView& d, unsigned int slot) {
	const real* o = d.simplex + slot;
	const size_t S = d.cand_cap;
	Simplex s;
	s.a = v3(o[0], o[S], o[2 * S]); s.b = v3(o[3 * S], o[4 * S], o[5 * S]);
	s.c = v3(o[6 * S], o[7 * S], o[8 * S]); s.d = v3(o[9 * S], o[10 * S], o[11 * S]);
	s.num = 4;
	return s;
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
#if defined(RP_REAL_F32)
#define RP_CULL_MARGIN 1e-5f  // (single precision: the bounds and the vertices GJK sees carry float rounding)
#else
#define RP_CULL_MARGIN 1e-7
#endif
#ifndef RP_CULL_ILP
#define RP_CULL_ILP 1
#endif
// Work order. The lanes of a warp take the SAME pair index of 32 CONSECUTIVE worlds (lane = world, warp = pair index),
// and every list built downstream (candidates -> hits -> level lists) keeps that order. The worlds of a batch are
// instances of one template, so their pair lists line up and the same pair in neighbouring worlds is in a similar
// configuration: GJK/EPA iteration counts, clipping cases and manifold sizes are correlated across the lanes of a warp
// (identical for identical worlds), which keeps the warps of the branchy narrowphase and of the solver converged and
// their world-minor loads coalesced. With unrelated worlds the order is merely as good as any other.
__global__ void __launch_bounds__(256) k_cull(DevView d, int cull) {
	const int lane = threadIdx.x & 31;
	const int w = blockIdx.y * 32 + lane;
	const bool wlive = w < d.W;
	const int np = wlive ? d.n_pairs[w] : 0;
	int np_max = np;
	for (int o = 16; o > 0; o >>= 1) np_max = max(np_max, __shfl_xor_sync(0xffffffffu, np_max, o));
	const int w0 = wlive ? w : 0;
	const int* active = d.active + w0;
	const size_t S = d.WS;
#if defined(RP_STORED_NORMALS)
	const int epoch = *d.epoch;
#endif
	int tested = 0;
	const int warps_per_cta = blockDim.x >> 5;
	const int stride = gridDim.x * warps_per_cta;
	// The test is three dependent round trips to memory and nothing else (pair record -> sleep flags + bounds), and the
	// bounds are what it moves: 12 doubles per pair, 260 MB per substep through L2 for the north-star batch. So: (1) the
	// bounds are fetched one axis at a time, vertical axis first -- resting and stacked bodies, the floor pairs above all,
	// separate along y -- and the other two axes only if some lane of the warp still needs them; (2) RP_CULL_ILP pair
	// indices per trip, every load of a stage issued for all of them before its first use.
	for (int p0 = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); p0 < np_max; p0 += RP_CULL_ILP * stride) {
		PairRec pr[RP_CULL_ILP];
		bool in[RP_CULL_ILP], keep[RP_CULL_ILP], bounds[RP_CULL_ILP];
#pragma unroll
		for (int u = 0; u < RP_CULL_ILP; ++u) {
			const int p = p0 + u * stride;
			in[u] = p < np;
			pr[u] = d.pairs[pidx(d, in[u] ? p : 0, w0)];
		}
		int fa[RP_CULL_ILP], fb[RP_CULL_ILP], aa[RP_CULL_ILP], ab[RP_CULL_ILP], ta[RP_CULL_ILP], tb[RP_CULL_ILP];
		real lo_a[RP_CULL_ILP], hi_a[RP_CULL_ILP], lo_b[RP_CULL_ILP], hi_b[RP_CULL_ILP];
#pragma unroll
		for (int u = 0; u < RP_CULL_ILP; ++u) {
			fa[u] = d.bstat[pr[u].a].fixed; fb[u] = d.bstat[pr[u].b].fixed;
			aa[u] = active[pr[u].a * S]; ab[u] = active[pr[u].b * S];
			ta[u] = d.cols[pr[u].ca].type; tb[u] = d.cols[pr[u].cb].type;
			const float* pa = d.aabb + (size_t)pr[u].ca * 6 * S + w0;
			const float* pb = d.aabb + (size_t)pr[u].cb * 6 * S + w0;
			lo_a[u] = pa[S]; hi_a[u] = pa[4 * S]; lo_b[u] = pb[S]; hi_b[u] = pb[4 * S];
		}
#pragma unroll
		for (int u = 0; u < RP_CULL_ILP; ++u) {
			// pbd.cpp:594 skip rule; sphere-sphere pairs never reach GJK (collider.cpp:530) and are never culled
			keep[u] = in[u] && !((fa[u] || !aa[u]) && (fb[u] || !ab[u]));
			if (keep[u]) ++tested;
			bounds[u] = keep[u] && cull && !(ta[u] == SHAPE_SPHERE && tb[u] == SHAPE_SPHERE);
			if (bounds[u] && (lo_a[u] - hi_b[u] > RP_CULL_MARGIN || lo_b[u] - hi_a[u] > RP_CULL_MARGIN)) keep[u] = bounds[u] = false;
		}
#pragma unroll
		for (int u = 0; u < RP_CULL_ILP; ++u) {
			if (__any_sync(0xffffffffu, bounds[u])) {
				const float* pa = d.aabb + (size_t)pr[u].ca * 6 * S + w0;
				const float* pb = d.aabb + (size_t)pr[u].cb * 6 * S + w0;
				const real ax0 = pa[0], ax1 = pa[3 * S], az0 = pa[2 * S], az1 = pa[5 * S];
				const real bx0 = pb[0], bx1 = pb[3 * S], bz0 = pb[2 * S], bz1 = pb[5 * S];
				if (bounds[u] && (ax0 - bx1 > RP_CULL_MARGIN || bx0 - ax1 > RP_CULL_MARGIN || az0 - bz1 > RP_CULL_MARGIN ||
					bz0 - az1 > RP_CULL_MARGIN)) keep[u] = false;
			}
		}
#pragma unroll
		for (int u = 0; u < RP_CULL_ILP; ++u) {
			const int p = p0 + u * stride;
			if (in[u]) d.pair_ccnt[pidx(d, p, w)] = 0;
			// pairs whose hulls do not fit k_gjk's per-thread staging block go to the back of the list, for k_gjk_warp
			const bool big = keep[u] && warp_pair_verts(d.cols[pr[u].ca].nv + d.cols[pr[u].cb].nv);
			const unsigned int front = warp_append(d.cand_count, keep[u] && !big);
			const unsigned int back = warp_append(d.big_count, big);
			const unsigned int slot = big ? d.cand_cap - 1u - back : front;
			if (keep[u]) {
				d.cands[slot] = make_uint4((unsigned int)w, (unsigned int)p, (unsigned int)pr[u].ca, (unsigned int)pr[u].cb);
#if defined(RP_STORED_NORMALS)
				// k_transform writes the face normals of these two colliders (every writer of a stamp writes the same value)
				d.geom_stamp[(size_t)pr[u].ca * S + w] = epoch;
				d.geom_stamp[(size_t)pr[u].cb * S + w] = epoch;
#endif
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(0xffffffffu, tested, o);
	if (lane == 0 && tested) atomicAdd(&d.counters[CNT_PAIR_TESTS], (unsigned long long)tested);
}

// ----------------------------------------------------------------------------------------------------- narrowphase
// A warp's cursor over a contiguous chunk of a work list (used by the solver's refill loops): every trip of the loop is
// one step for whatever item a lane currently holds, and a lane whose item is finished takes the warp's next item
// before the next trip. The cursor is warp-uniform (ballot + popc), so taking work needs no atomics.
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

// One thread per candidate pair: sphere-sphere test or boolean GJK (collider.cpp:523-547). Colliding pairs are appended
// (warp-aggregated, order-preserving within the warp) to the hit list together with their final simplex.
__global__ void __launch_bounds__(RP_GJK_THREADS, RP_MINB_GJK) k_gjk(DevView d) {
	const unsigned int nc = *d.cand_count;
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&d.counters[CNT_CANDS], (unsigned long long)nc + *d.big_count);
	__shared__ real s_stage[RP_GJK_STAGE * RP_GJK_THREADS];
	for (unsigned int c0 = blockIdx.x * blockDim.x; c0 < nc; c0 += gridDim.x * blockDim.x) {
		const unsigned int ci = c0 + threadIdx.x;
		bool hit = false;
		Simplex s;
		s.a = s.b = s.c = s.d = v3(RL(0.0), RL(0.0), RL(0.0));
		uint4 cd = make_uint4(0u, 0u, 0u, 0u);
		if (ci < nc) {
			cd = d.cands[ci];  // (world, pair, collider a, collider b): everything the narrowphase needs to find its inputs
			const int w = (int)cd.x;
			PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
			PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
			int st = 0;
			if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
				V3 n;
				real depth;
				hit = sphere_sphere(A, B, &n, &depth);
			} else {
				if ((A.nv + B.nv) * 3 <= RP_GJK_STAGE) {
					real* col = s_stage + threadIdx.x;
					const int used = stage_shape(A, col, RP_GJK_THREADS);
					stage_shape(B, col + (size_t)used * RP_GJK_THREADS, RP_GJK_THREADS);
					hit = gjk(StagedShape<RP_GJK_THREADS>(A), StagedShape<RP_GJK_THREADS>(B), &s, &st, 0);
				} else {
					hit = gjk(A, B, &s, &st, 0);
				}
			}
			if (st) atomicOr(&d.status[w], st);
		}
		const unsigned int slot = warp_append(d.hit_count, hit);
		if (hit) {
			d.hits[slot] = cd;
			st_simplex(d, slot, s);
		}
	}
}

// ---- one WARP per candidate pair, for hulls too large to stage per thread (a cylinder has 128 vertices, a tessellated
// sphere over a thousand): the 32 lanes run the same GJK instance in lockstep, and the support scans -- everything such a
// pair spends its time in -- are split across them: every lane keeps the first maximum of its share of the vertices
// (ascending, strict >, as support_point_get_index does, support.cpp:5-17), then a butterfly of shuffles picks the largest
// value and, among equal values, the lowest index: exactly the vertex the sequential scan returns. Hulls of up to
// RP_WARP_HULL_MAX vertices are staged in shared memory first (component planes, so the lanes' reads are consecutive).
#define RP_GJK_WARP_THREADS 128
#define RP_WARP_HULL_MAX 128
struct WarpShape : PoseShape {
	__device__ WarpShape() {}
	__device__ explicit WarpShape(const PoseShape& s) : PoseShape(s) {}
};
// staged in the warp's block (vp set by warp_stage) or, for hulls too large for it, evaluated from the pose in place
__device__ __forceinline__ V3 vert(const WarpShape& s, int i) {
	return s.vp ? vert(static_cast<const Shape&>(s), i) : vert(static_cast<const PoseShape&>(s), i);
}
__device__ __forceinline__ int support_index(const WarpShape& s, V3 d) {
	const int lane = threadIdx.x & 31;
	int best = 0x7fffffff;
	real best_dot = -RL(RP_REAL_MAX);
	for (int i = lane; i < s.nv; i += 32) {
		const real t = dot(vert(s, i), d);
		if (t > best_dot) {
			best = i;
			best_dot = t;
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const real od = __shfl_xor_sync(0xffffffffu, best_dot, o);
		const int oi = __shfl_xor_sync(0xffffffffu, best, o);
		if (od > best_dot || (od == best_dot && oi < best)) {
			best_dot = od;
			best = oi;
		}
	}
	return best == 0x7fffffff ? 0 : best;
}
// evaluates a hull's transformed vertices into the warp's block of shared memory as component planes
__device__ __forceinline__ void warp_stage(PoseShape& s, real* block) {
	if (s.type != SHAPE_HULL || s.nv > RP_WARP_HULL_MAX) return;  // spheres have no vertices; larger hulls are evaluated in place
	const int lane = threadIdx.x & 31;
	for (int k = lane; k < s.nv; k += 32) {
		const V3 p = vert(s, k);
		block[k] = p.x; block[RP_WARP_HULL_MAX + k] = p.y; block[2 * RP_WARP_HULL_MAX + k] = p.z;
	}
	s.vp = block; s.vs = 1; s.vcs = RP_WARP_HULL_MAX;
}
__global__ void __launch_bounds__(RP_GJK_WARP_THREADS) k_gjk_warp(DevView d) {
	const unsigned int nb = *d.big_count;
	__shared__ real s_hull[RP_GJK_WARP_THREADS / 32][2][3 * RP_WARP_HULL_MAX];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const unsigned int warps = gridDim.x * (RP_GJK_WARP_THREADS / 32);
	for (unsigned int k = blockIdx.x * (RP_GJK_WARP_THREADS / 32) + wib; k < nb; k += warps) {
		const uint4 cd = d.cands[d.cand_cap - 1u - k];
		const int w = (int)cd.x;
		PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
		PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
		__syncwarp();  // the previous pair's scans are done with the block
		warp_stage(A, s_hull[wib][0]);
		warp_stage(B, s_hull[wib][1]);
		__syncwarp();
		Simplex s;
		int st = 0;
		const bool hit = gjk(WarpShape(A), WarpShape(B), &s, &st, 0);  // the same instance on all 32 lanes
		if (lane == 0) {
			if (st) atomicOr(&d.status[w], st);
			if (hit) {
				const unsigned int slot = atomicAdd(d.hit_count, 1u);
				d.hits[slot] = cd;
				st_simplex(d, slot, s);
			}
		}
	}
}

// The EPA polytope of one thread in a thread-interleaved block of shared memory (element e of thread t at base[e * NT + t]:
// the 32 lanes of a warp touch 32 consecutive words whatever each lane indexes). Store concept of rp_narrow.h, small
// capacities, SOFT: a pair whose polytope outgrows it is rerun on the full-capacity store in local memory (epa_full).
// Round 1 kept the full store (11.8 kB) in every thread's local memory; at 12 warps per SM it missed L1 on every step
// (ncu: L1 hit 53 %, 6 long-scoreboard stall cycles per issued instruction, FP64 pipe 9.7 %).
#define RP_EPA_POLY_DOUBLES (3 * RP_EPA_SMALL_VERTS + 4 * RP_EPA_SMALL_FACES)  // vertices, face normals, face distances
#define RP_EPA_POLY_INTS (RP_EPA_SMALL_FACES + RP_EPA_SMALL_EDGES)             // packed face triples, packed edges
template <int NT>
struct EpaShared {
	enum { MAXV = RP_EPA_SMALL_VERTS, MAXF = RP_EPA_SMALL_FACES, MAXE = RP_EPA_SMALL_EDGES, SOFT = 1 };
	real* col;  // this thread's column of doubles: vertices [0, 3V), normals [3V, 3V + 3F), distances [3V + 3F, 3V + 4F)
	int* icol;    // this thread's column of ints: faces [0, F), edges [F, F + E)
	int nverts, nfaces, nedges;
	V3 min_normal;
	real min_dist;
	__device__ __forceinline__ V3 vert(int i) const { const real* p = col + (3 * i) * NT; return v3(p[0], p[NT], p[2 * NT]); }
	__device__ __forceinline__ void set_vert(int i, V3 v) { real* p = col + (3 * i) * NT; p[0] = v.x; p[NT] = v.y; p[2 * NT] = v.z; }
	__device__ __forceinline__ V3 normal(int i) const { const real* p = col + (3 * MAXV + 3 * i) * NT; return v3(p[0], p[NT], p[2 * NT]); }
	__device__ __forceinline__ real dist(int i) const { return col[(3 * MAXV + 3 * MAXF + i) * NT]; }
	__device__ __forceinline__ void set_plane(int i, V3 n, real d) {
		real* p = col + (3 * MAXV + 3 * i) * NT;
		p[0] = n.x; p[NT] = n.y; p[2 * NT] = n.z;
		col[(3 * MAXV + 3 * MAXF + i) * NT] = d;
	}
	__device__ __forceinline__ void face(int i, int* x, int* y, int* z) const {
		const int f = icol[i * NT];
		*x = f & 255; *y = (f >> 8) & 255; *z = f >> 16;
	}
	__device__ __forceinline__ void set_face(int i, int x, int y, int z) { icol[i * NT] = x | (y << 8) | (z << 16); }
	__device__ __forceinline__ void move_face(int dst, int src) {
		icol[dst * NT] = icol[src * NT];
		set_plane(dst, normal(src), dist(src));
	}
	__device__ __forceinline__ void edge(int i, int* x, int* y) const {
		const int e = icol[(MAXF + i) * NT];
		*x = e & 255; *y = e >> 8;
	}
	__device__ __forceinline__ void set_edge(int i, int x, int y) { icol[(MAXF + i) * NT] = x | (y << 8); }
};

// second tier of k_epa: the full-capacity polytope in local memory, for the rare pair that outgrows the shared store.
// Out of line so that its 11.8 kB frame and its registers stay out of the common path.
__device__ __noinline__ int epa_full(const DevView& d, uint4 cd, const Simplex& s, V3* normal, real* depth, int* status, int* sup_a, int* sup_b) {
	EpaScratch e;
	const PoseShape A = dev_pose_shape(d, d.cols[cd.z], (int)cd.x);
	const PoseShape B = dev_pose_shape(d, d.cols[cd.w], (int)cd.x);
	return epa_run(A, B, s, e, normal, depth, status, 0, sup_a, sup_b);
}

// EPA (epa.cpp:118) for every hit, one thread per hit, the polytope in shared memory (EpaShared). RP_EPA_STAGE_HULLS = 1
// also stages the two hulls' vertices in shared memory as k_gjk does (EPA only ever asks for support points); with the
// polytope there as well that halves the resident warps, and EPA scans each hull once or twice where GJK scans it four
// times, so by default the vertices are read in place (world-minor arrays: coalesced across the lanes of a warp). Kept
// apart from k_manifold: lanes leave EPA after different numbers of iterations, and the kernel boundary is what brings a
// warp back together before the clipping code.
#ifndef RP_EPA_STAGE_HULLS
#define RP_EPA_STAGE_HULLS 0
#endif
#define RP_EPA_SMEM_BYTES ((RP_EPA_POLY_DOUBLES * (int)sizeof(real) + RP_EPA_POLY_INTS * 4 + (RP_EPA_STAGE_HULLS ? RP_GJK_STAGE * (int)sizeof(real) : 0)) * RP_EPA_THREADS)
__global__ void __launch_bounds__(RP_EPA_THREADS, RP_MINB_EPA) k_epa(DevView d) {
	const unsigned int nh = *d.hit_count;
	extern __shared__ __align__(16) unsigned char s_epa_raw[];
	real* s_poly = reinterpret_cast<real*>(s_epa_raw);
	real* s_stage = s_poly + RP_EPA_POLY_DOUBLES * RP_EPA_THREADS;
	int* s_idx = reinterpret_cast<int*>(s_stage + (RP_EPA_STAGE_HULLS ? RP_GJK_STAGE * RP_EPA_THREADS : 0));
	EpaShared<RP_EPA_THREADS> e;
	e.col = s_poly + threadIdx.x;
	e.icol = s_idx + threadIdx.x;
	for (unsigned int hi = blockIdx.x * blockDim.x + threadIdx.x; hi < nh; hi += gridDim.x * blockDim.x) {
		const uint4 cd = d.hits[hi];
		const int w = (int)cd.x;
		if (d.split_big && warp_pair_verts(d.cols[cd.z].nv + d.cols[cd.w].nv)) continue;  // k_epa_warp's
		PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
		PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
		EpaOut out;
		out.ok = 0; out.pad = 0; out.depth = RL(0.0); out.normal = v3(RL(0.0), RL(0.0), RL(0.0));
		out.sup_a = out.sup_b = -1;
		int st = 0;
		if (A.type == SHAPE_SPHERE && B.type == SHAPE_SPHERE) {
			out.ok = sphere_sphere(A, B, &out.normal, &out.depth) ? 1 : 0;
		} else {
			const Simplex s = ld_simplex(d, hi);
			int r;
#if RP_EPA_STAGE_HULLS
			if ((A.nv + B.nv) * 3 <= RP_GJK_STAGE) {
				PoseShape SA = A, SB = B;
				real* col = s_stage + threadIdx.x;
				const int used = stage_shape(SA, col, RP_EPA_THREADS);
				stage_shape(SB, col + (size_t)used * RP_EPA_THREADS, RP_EPA_THREADS);
				r = epa_run(StagedShape<RP_EPA_THREADS>(SA), StagedShape<RP_EPA_THREADS>(SB), s, e, &out.normal, &out.depth, &st, 0, &out.sup_a, &out.sup_b);
			} else
#endif
			r = epa_run(A, B, s, e, &out.normal, &out.depth, &st, 0, &out.sup_a, &out.sup_b);
			if (r == EPA_OVERFLOW) r = epa_full(d, cd, s, &out.normal, &out.depth, &st, &out.sup_a, &out.sup_b);
#if defined(RP_REAL_F32)
			if (r == EPA_FAIL && sat_face_fallback(A, B, &out.normal, &out.depth, &out.sup_a, &out.sup_b)) {  // (rp_narrow.h: single precision only)
				r = EPA_DONE;
				st &= ~(ST_EPA_DEGENERATE | ST_EPA_NO_CONVERGENCE);
			}
#endif
			out.ok = r == EPA_DONE ? 1 : 0;
		}
		d.epa_out[hi] = out;
		if (st) atomicOr(&d.status[w], st);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&d.counters[CNT_HITS], (unsigned long long)nh);
}

// EPA for the hits of large pairs, one warp per hit like k_gjk_warp (the polytope is small; the support scans over the
// big hulls are what is shared out). While the hulls are still staged, the warp also finds the two support vertices along
// +-normal that clipping starts from (convex_convex_contact_manifold, clipping.cpp:255-256) and leaves them for k_manifold.
__global__ void __launch_bounds__(RP_GJK_WARP_THREADS) k_epa_warp(DevView d) {
	const unsigned int nh = *d.hit_count;
	__shared__ real s_hull[RP_GJK_WARP_THREADS / 32][2][3 * RP_WARP_HULL_MAX];
	EpaScratch e;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const unsigned int warps = gridDim.x * (RP_GJK_WARP_THREADS / 32);
	for (unsigned int hi = blockIdx.x * (RP_GJK_WARP_THREADS / 32) + wib; hi < nh; hi += warps) {
		const uint4 cd = d.hits[hi];
		if (!(warp_pair_verts(d.cols[cd.z].nv + d.cols[cd.w].nv))) continue;  // k_epa's
		const int w = (int)cd.x;
		PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
		PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
		__syncwarp();
		warp_stage(A, s_hull[wib][0]);
		warp_stage(B, s_hull[wib][1]);
		__syncwarp();
		const Simplex s = ld_simplex(d, hi);
		EpaOut out;
		out.ok = 0; out.pad = 0; out.depth = RL(0.0); out.normal = v3(RL(0.0), RL(0.0), RL(0.0));
		out.sup_a = out.sup_b = -1;
		int st = 0;
		out.ok = epa(WarpShape(A), WarpShape(B), s, e, &out.normal, &out.depth, &st, 0) ? 1 : 0;  // the same instance on all lanes
		int2 sup = make_int2(-1, -1);
		if (out.ok && A.type == SHAPE_HULL && B.type == SHAPE_HULL) {
			sup.x = support_index(WarpShape(A), out.normal);
			sup.y = support_index(WarpShape(B), zero_minus(out.normal));
		}
		if (lane == 0) {
			d.epa_out[hi] = out;
			d.big_sup[hi] = sup;
			if (st) atomicOr(&d.status[w], st);
		}
	}
}

// The two Sutherland-Hodgman polygon buffers of one thread in a thread-interleaved block of shared memory (store concept of
// rp_narrow.h): RP_CLIP_SMALL_POINTS points each -- a box face clipped against the four side planes of another box face has
// at most 8 vertices. A polygon that outgrows it (a cylinder's 64-gon cap) is rerun on the full store in local memory
// (manifold_full). Round 1 kept 2 x 160 points plus a 320-point staging array (15 kB) in every thread's local memory.
template <int NT>
struct ClipShared {
	enum { CAP = RP_CLIP_SMALL_POINTS, SOFT = 1 };
	real* col;  // point i of buffer b: components at col[((b * CAP + i) * 3 + c) * NT]
	__device__ __forceinline__ V3 get(int b, int i) const { const real* p = col + ((b * CAP + i) * 3) * NT; return v3(p[0], p[NT], p[2 * NT]); }
	__device__ __forceinline__ void set(int b, int i, V3 v) { real* p = col + ((b * CAP + i) * 3) * NT; p[0] = v.x; p[NT] = v.y; p[2 * NT] = v.z; }
};
#define RP_MANIFOLD_SMEM_BYTES (2 * RP_CLIP_SMALL_POINTS * 3 * (int)sizeof(real) * RP_MANIFOLD_THREADS)

struct CountSink {
	int n;
	__device__ __forceinline__ void operator()(V3, V3) { ++n; }
};
// writes the contact records (pbd.cpp:408-424) of one manifold straight into the world's contact buffer
struct ContactSink {
	const DevView* d;
	Body b1, b2;   // poses only
	int w, off, n, cap;
	__device__ __forceinline__ void operator()(V3 p1, V3 p2) {
		if (n < cap) {
			st_contact(contact_ptr(*d, w, off + n), d->WS, make_contact(b1, b2, p1, p2));
			if (w == d->dbg_world) {
				d->dbg_points[2 * (off + n)] = p1;
				d->dbg_points[2 * (off + n) + 1] = p2;
			}
		}
		++n;
	}
};
// Second half of a manifold, shared by both tiers: count the contacts of the clipped polygon (manifold_emit with a counting
// sink), take a contiguous run of the world's contact buffer (allocation order between pairs is irrelevant: the solver walks
// pairs, not the buffer), emit again into it. Returns the number of contacts stored; *off_out = first slot.
template <class C>
__device__ __forceinline__ int emit_contacts(const DevView& d, const C& cs, const ClipResult& r, V3 normal, int w, const PairRec& pr, int* off_out,
	int* status) {
	CountSink cnt;
	cnt.n = 0;
	manifold_emit(cs, r, normal, cnt);
	if (cnt.n == 0) return 0;
	ContactSink sink;
	sink.d = &d; sink.w = w; sink.n = 0;
	sink.off = atomicAdd(&d.n_contacts[w], cnt.n);
	sink.cap = cnt.n;
	if (sink.off + cnt.n > d.max_contacts) {
		*status |= ST_CONTACT_CAPACITY;
		sink.cap = d.max_contacts - sink.off;
		if (sink.cap < 0) sink.cap = 0;
	}
	const DynRef ra = dyn_ref(d, w, pr.a);
	const DynRef rb = dyn_ref(d, w, pr.b);
	sink.b1.x = ld3(ra, DF_X); sink.b1.q = ld4(ra, DF_Q);
	sink.b2.x = ld3(rb, DF_X); sink.b2.q = ld4(rb, DF_Q);
	manifold_emit(cs, r, normal, sink);
	*off_out = sink.off;
	return sink.cap;
}
// second tier of k_manifold: full-capacity polygon buffers in local memory. Out of line (7.7 kB frame).
__device__ __noinline__ int manifold_full(const DevView& d, uint4 cd, V3 normal, PairRec pr, int sup1, int sup2, int* off_out, int* status) {
	ClipScratch cs;
	ClipResult r;
	const int w = (int)cd.x;
	const PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
	const PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
	manifold_clip(A, B, normal, cs, status, &r, sup1, sup2);
	return emit_contacts(d, cs, r, normal, w, pr, off_out, status);
}

// One thread per colliding collider pair whose EPA converged: manifold (clipping.cpp:343), contact -> constraint
// (pbd.cpp:408-424), and the pair is appended to the work list of its dependency level.
__global__ void __launch_bounds__(RP_MANIFOLD_THREADS, RP_MINB_MANIFOLD) k_manifold(DevView d) {
	const unsigned int nh = *d.hit_count;
	extern __shared__ __align__(16) unsigned char s_clip_raw[];
	ClipShared<RP_MANIFOLD_THREADS> cs;
	cs.col = reinterpret_cast<real*>(s_clip_raw) + threadIdx.x;
	int made = 0;
	const int lane = threadIdx.x & 31;
	for (unsigned int h0 = blockIdx.x * blockDim.x; h0 < nh; h0 += gridDim.x * blockDim.x) {
		const unsigned int hi = h0 + threadIdx.x;
		int n = 0, lvl = -1, w = 0, pair = 0;
		SolveItem item;
		item.w = item.a = item.b = item.coff = item.cnt = item.pad = 0;
		item.normal = v3(RL(0.0), RL(0.0), RL(0.0));
		if (hi < nh) {
			const EpaOut eo = d.epa_out[hi];
			const uint4 cd = d.hits[hi];
			w = (int)cd.x; pair = (int)cd.y;
			const size_t pg = pidx(d, pair, w);
			int st = 0;
			// pairs of large hulls are k_manifold_warp's (the whole warp clips one polygon)
			const bool warp_pair = d.split_big && warp_pair_verts(d.cols[cd.z].nv + d.cols[cd.w].nv);
			if (eo.ok && !warp_pair) {
				const PairRec pr = d.pairs[pg];
				const int ta = d.cols[cd.z].type, tb = d.cols[cd.w].type;
				int off = 0;
				if (ta == SHAPE_SPHERE || tb == SHAPE_SPHERE) {
					const PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
					const PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
					// clipping_get_contact_manifold's sphere cases (clipping.cpp:348-364): one contact
					ClipResult r;
					r.kind = 1; r.n = 0; r.cur = 0; r.ref1 = false;
					r.rp_normal = r.rp_point = v3(RL(0.0), RL(0.0), RL(0.0));
					if (A.type == SHAPE_SPHERE) {
						r.l1 = support(A, eo.normal);
						r.l2 = sub(r.l1, scale(eo.depth, eo.normal));
					} else {
						r.l2 = support(B, zero_minus(eo.normal));
						r.l1 = add(r.l2, scale(eo.depth, eo.normal));
					}
					n = emit_contacts(d, cs, r, eo.normal, w, pr, &off, &st);
				} else {
					const int sup1 = eo.sup_a, sup2 = eo.sup_b;  // EPA's last support vertices = the ones clipping starts from
					ClipResult r;
					int ov;
					{
						const PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
						const PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
						ov = manifold_clip(A, B, eo.normal, cs, &st, &r, sup1, sup2);
					}
					if (ov == CLIP_OVERFLOW) {
						n = manifold_full(d, cd, eo.normal, pr, sup1, sup2, &off, &st);
					} else {
						n = emit_contacts(d, cs, r, eo.normal, w, pr, &off, &st);
					}
				}
				if (n > 0) {
					d.pair_normal[pg] = eo.normal;
					d.pair_coff[pg] = off;
					d.pair_ccnt[pg] = n;
					made += n;
					lvl = d.pair_level[pg];
					item.w = w; item.a = pr.a; item.b = pr.b; item.coff = off; item.cnt = n; item.normal = eo.normal;
				}
			}
			if (st) atomicOr(&d.status[w], st);
		}
		// append the unit (a self-contained SolveItem: the sweeps start on it after one load) to the list of its level.
		// Ranks come from the WARP: lanes with the same (level, end) elect a leader that takes their slots with one
		// atomic on that level's fill counter (each counter on a line of its own, RP_LVL_STRIDE). With lane = world the
		// lanes of a warp hold the same pair of neighbouring worlds and so, as a rule, the same level: one atomic per warp
		// and no CTA-wide barrier (four __syncthreads per trip made every warp wait for the CTA's slowest clipping case).
		// A level's list can be filled from both ends -- small manifolds (<= RP_SMALL_MANIFOLD contacts) from the front,
		// large ones from the back -- which groups manifolds of similar length (order within a level is free).
		// World-block sweeps (k_solve_block) only need to know which of a world's units are live: (pair, level) appended to
		// the world's own list; the block's CTA sorts them by level itself.
		if (d.block_mode) {
			if (lvl > 0) {
				const int slot = atomicAdd(&d.n_live[w], 1);
				d.live[(size_t)slot * d.WS + w] = make_uint2((unsigned int)pair, (unsigned int)lvl);
			}
		} else {
			const int big = n > RP_SMALL_MANIFOLD ? 1 : 0;
			const int key = lvl > 0 ? 2 * lvl + big : -1;
			const unsigned int peers = __match_any_sync(0xffffffffu, key);
			int slot = 0;
			if (lvl > 0) {
				const int leader = __ffs(peers) - 1;
				if (lane == leader) slot = atomicAdd(&d.lvl_fill[(size_t)lvl * RP_LVL_STRIDE + big], __popc(peers));
				slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1u));
				const int at = big ? d.lvl_off[lvl + 1] - 1 - slot : d.lvl_off[lvl] + slot;
				d.lvl_items[at] = item;
				if (d.flow_mode && lvl <= RP_FLOW_LEVELS) flow_mark_live(d, item.a, item.b, w, lvl);
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) made += __shfl_down_sync(0xffffffffu, made, o);
	if (lane == 0 && made) atomicAdd(&d.counters[CNT_CONTACTS], (unsigned long long)made);
}

// ---- one WARP per colliding pair of LARGE hulls (the pairs k_gjk_warp / k_epa_warp took): Sutherland-Hodgman with the lanes
// sharing out the polygon. A cylinder cap is a 64-gon clipped against the 64 side planes of the other cap (clipping.cpp:52-113);
// one thread walking that is 4096 dependent inside tests per pair. Here, per plane, lane l takes vertices l, l + 32, ...: the
// inside test of the vertex and of its predecessor, the edge intersection if they differ (the same expressions, float round
// trips included), and a warp prefix sum of the 0 / 1 / 2 points each lane emits puts them where the sequential loop would
// (intersection before end point, edges in order). Face selection and the edge-edge test run redundantly on all lanes (the
// same instance, as in k_epa_warp). The final cull against the reference plane and the contact emission are compactions of
// the same kind.
#define RP_CLIPW_THREADS 128
__device__ __forceinline__ int warp_exclusive_scan(int v, int* total) {
	const int lane = threadIdx.x & 31;
	int inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	*total = __shfl_sync(0xffffffffu, inc, 31);
	return inc - v;
}
// one pass of `in` (n_in points) against a plane into `out`; returns the number of points written, -1 if they do not fit
__device__ __forceinline__ int warp_clip_pass(const ClipPlane& pl, const V3* in, int n_in, V3* out, bool remove_only) {
	const int lane = threadIdx.x & 31;
	const float offset = clip_offset(pl);
	int base = 0;
	bool overflow = false;
	for (int j0 = 0; j0 < n_in; j0 += 32) {
		const int j = j0 + lane;
		int cnt = 0;
		bool has_x = false;
		V3 x = v3(RL(0.0), RL(0.0), RL(0.0)), end = v3(RL(0.0), RL(0.0), RL(0.0));
		bool e_in = false;
		if (j < n_in) {
			end = in[j];
			const V3 start = in[j == 0 ? n_in - 1 : j - 1];
			e_in = clip_inside(pl, offset, end);
			if (remove_only) {
				cnt = e_in ? 1 : 0;
			} else {
				const bool s_in = clip_inside(pl, offset, start);
				if (s_in != e_in) has_x = clip_edge(pl, offset, start, end, &x);
				cnt = (has_x ? 1 : 0) + (e_in ? 1 : 0);
			}
		}
		int total;
		int at = base + warp_exclusive_scan(cnt, &total);
		if (base + total > RP_CLIP_MAX_POINTS) {
			overflow = true;
			break;
		}
		if (has_x) out[at++] = x;
		if (e_in && j < n_in) out[at] = end;
		base += total;
	}
	__syncwarp();
	return overflow ? -1 : base;
}

__global__ void __launch_bounds__(RP_CLIPW_THREADS) k_manifold_warp(DevView d) {
	const unsigned int nh = *d.hit_count;
	__shared__ V3 s_poly[RP_CLIPW_THREADS / 32][2][RP_CLIP_MAX_POINTS];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const unsigned int warps = gridDim.x * (RP_CLIPW_THREADS / 32);
	int made = 0;
	for (unsigned int hi = blockIdx.x * (RP_CLIPW_THREADS / 32) + wib; hi < nh; hi += warps) {
		const uint4 cd = d.hits[hi];
		if (!(warp_pair_verts(d.cols[cd.z].nv + d.cols[cd.w].nv))) continue;  // k_manifold's
		const EpaOut eo = d.epa_out[hi];
		if (!eo.ok) continue;
		const int w = (int)cd.x, pair = (int)cd.y;
		const size_t pg = pidx(d, pair, w);
		const PairRec pr = d.pairs[pg];
		const PoseShape A = dev_pose_shape(d, d.cols[cd.z], w);
		const PoseShape B = dev_pose_shape(d, d.cols[cd.w], w);
		const V3 normal = eo.normal;
		int st = 0;
		V3* buf0 = s_poly[wib][0];
		V3* buf1 = s_poly[wib][1];
		__syncwarp();
		// the n contacts end up as (pts1[k], pts2[k])
		int n = 0;
		V3* pts1 = buf0;
		V3* pts2 = buf1;
		if (A.type == SHAPE_SPHERE || B.type == SHAPE_SPHERE) {  // clipping.cpp:348-364
			if (lane == 0) {
				if (A.type == SHAPE_SPHERE) {
					buf0[0] = support(A, normal);
					buf1[0] = sub(buf0[0], scale(eo.depth, normal));
				} else {
					buf1[0] = support(B, zero_minus(normal));
					buf0[0] = add(buf1[0], scale(eo.depth, normal));
				}
			}
			n = 1;
		} else {
			const int2 sup = d.big_sup[hi];
			FaceChoice fc;
			manifold_select(A, B, normal, &st, &fc, sup.x, sup.y);
			if (fc.kind == 1) {
				if (lane == 0) {
					buf0[0] = fc.l1;
					buf1[0] = fc.l2;
				}
				n = 1;
			} else if (fc.kind == 2) {
				const PoseShape& R = fc.ref1 ? A : B;
				const PoseShape& I = fc.ref1 ? B : A;
				// incident polygon (clipping.cpp:241-247)
				int m = I.face_ptr[fc.iface + 1] - I.face_ptr[fc.iface];
				if (m > RP_CLIP_MAX_POINTS) {
					st |= ST_CLIP_CAPACITY;
					m = 0;
				}
				for (int k = lane; k < m; k += 32) buf0[k] = vert(I, I.face_idx[I.face_ptr[fc.iface] + k]);
				__syncwarp();
				V3* cur = buf0;
				V3* oth = buf1;
				for (int k = R.f2n_ptr[fc.rface]; k < R.f2n_ptr[fc.rface + 1] && m > 0; ++k) {
					const int nf = R.f2n_idx[k];
					ClipPlane pl;
					pl.point = vert(R, R.face_idx[R.face_ptr[nf]]);
					pl.normal = zero_minus(fnormal(R, nf));
					m = warp_clip_pass(pl, cur, m, oth, false);
					if (m < 0) {
						st |= ST_CLIP_CAPACITY;
						m = 0;
					}
					V3* t = cur; cur = oth; oth = t;
				}
				ClipPlane rp;
				rp.normal = zero_minus(fc.ref_normal);
				rp.point = vert(R, R.face_idx[R.face_ptr[fc.rface]]);
				if (m > 0) {
					m = warp_clip_pass(rp, cur, m, oth, true);
					V3* t = cur; cur = oth; oth = t;
				}
				// penetration of every candidate (clipping.cpp:322-338) and ordered compaction of the contacts: p1 into `oth`, p2 in
				// place over `cur` (slot at <= j, and every candidate of the trip has been read before any slot is written)
				int base = 0;
				for (int j0 = 0; j0 < m; j0 += 32) {
					const int j = j0 + lane;
					V3 p1 = v3(RL(0.0), RL(0.0), RL(0.0)), p2 = p1;
					const bool hit = j < m && manifold_point(cur[j], rp.normal, rp.point, fc.ref1, normal, &p1, &p2);
					int total;
					const int at = base + warp_exclusive_scan(hit ? 1 : 0, &total);
					__syncwarp();
					if (hit) {
						oth[at] = p1;
						cur[at] = p2;
					}
					base += total;
				}
				n = base;
				pts1 = oth;
				pts2 = cur;
			}
		}
		__syncwarp();
		int off = 0;
		if (n > 0) {
			if (lane == 0) {
				off = atomicAdd(&d.n_contacts[w], n);
				if (off + n > d.max_contacts) {
					st |= ST_CONTACT_CAPACITY;
					n = d.max_contacts - off;
					if (n < 0) n = 0;
				}
			}
			off = __shfl_sync(0xffffffffu, off, 0);
			n = __shfl_sync(0xffffffffu, n, 0);
			Body b1, b2;
			const DynRef ra = dyn_ref(d, w, pr.a);
			const DynRef rb = dyn_ref(d, w, pr.b);
			b1.x = ld3(ra, DF_X); b1.q = ld4(ra, DF_Q);
			b2.x = ld3(rb, DF_X); b2.q = ld4(rb, DF_Q);
			for (int k = lane; k < n; k += 32) {
				st_contact(contact_ptr(d, w, off + k), d.WS, make_contact(b1, b2, pts1[k], pts2[k]));
				if (w == d.dbg_world) {
					d.dbg_points[2 * (off + k)] = pts1[k];
					d.dbg_points[2 * (off + k) + 1] = pts2[k];
				}
			}
		}
		if (lane == 0) {
			if (n > 0) {
				d.pair_normal[pg] = normal;
				d.pair_coff[pg] = off;
				d.pair_ccnt[pg] = n;
				made += n;
				const int lvl = d.pair_level[pg];
				if (lvl > 0) {
					if (d.block_mode) {
						const int slot = atomicAdd(&d.n_live[w], 1);
						d.live[(size_t)slot * d.WS + w] = make_uint2((unsigned int)pair, (unsigned int)lvl);
					} else {
						SolveItem item;
						item.w = w; item.a = pr.a; item.b = pr.b; item.coff = off; item.cnt = n; item.pad = 0; item.normal = normal;
						const int big = n > RP_SMALL_MANIFOLD ? 1 : 0;
						const int slot = atomicAdd(&d.lvl_fill[(size_t)lvl * RP_LVL_STRIDE + big], 1);
						d.lvl_items[big ? d.lvl_off[lvl + 1] - 1 - slot : d.lvl_off[lvl] + slot] = item;
						if (d.flow_mode && lvl <= RP_FLOW_LEVELS) flow_mark_live(d, item.a, item.b, w, lvl);
					}
				}
			}
			if (st) atomicOr(&d.status[w], st);
		}
	}
	if (lane == 0 && made) atomicAdd(&d.counters[CNT_CONTACTS], (unsigned long long)made);
}

// -------------------------------------------------------------------------------------------------------------- solve
// Level-major Gauss-Seidel across ALL worlds: launch l runs every constraint of dependency level l -- the joints of
// level l of every world, then the (world, pair) items of level l that have contacts this substep (a pair's manifold is
// a sequential chain on its two bodies and stays in one thread, bodies in registers). Kernel boundaries are the barriers
// between levels, so the result equals the reference's sequential sweep (pbd.cpp:615-620). The contact part is a refill
// loop (WarpQueue): one trip = one contact of whatever manifold a lane holds, so lanes with short manifolds do not wait
// for lanes with long ones.
//
// ONE launch runs the whole sweep (k_solve_pos: every level of every positional iteration; k_solve_vel: every level of
// the velocity pass): a cooperative grid of exactly the resident CTAs walks the levels with a grid-wide barrier between
// two levels that both have work. Whether a level has work (joints of that level, or lvl_fill > 0 after this substep's
// k_manifold) is the same for every CTA, so EMPTY levels cost nothing -- no launch, no barrier. In the first second of
// the north-star window more than half of the scheduled levels hold no contact yet (ncu, frame 40: 6 of 11 level
// launches ran empty at ~5 us each).
// previous poses for the static-friction branch of solve_contact, fetched only if it is taken
struct PrevFromDyn {
	DynRef r1, r2;
	__device__ __forceinline__ void operator()(Body& b1, Body& b2) const {
		b1.px = ld3(r1, DF_PX); b1.pq = ld4(r1, DF_PQ);
		b2.px = ld3(r2, DF_PX); b2.pq = ld4(r2, DF_PQ);
	}
};
template <bool JOINTS>
__device__ __forceinline__ void pos_level(const DevView& d, real h, int level, int nj, int collisions) {
	const int njw = JOINTS ? nj * d.W : 0;
	int st = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < njw; i += gridDim.x * blockDim.x) {
		const int w = i % d.W;  // lane = world
		const int u = d.joint_sched[d.joint_lptr[level - 1] + i / d.W];
		const Joint j = d.joints[u];
		Body b1, b2;
		load_static(b1, d, j.e1);
		load_static(b2, d, j.e2);
		const DynRef r1 = dyn_ref(d, w, j.e1);
		const DynRef r2 = dyn_ref(d, w, j.e2);
		b1.x = ld3(r1, DF_X); b1.q = ld4(r1, DF_Q);
		b2.x = ld3(r2, DF_X); b2.q = ld4(r2, DF_Q);
		JointLambda lam = d.lambdas[(size_t)u * d.WS + w];
		solve_joint(j, lam, b1, b2, h, &st);
		d.lambdas[(size_t)u * d.WS + w] = lam;
		if (!b1.fixed) { st3(r1, DF_X, b1.x); st4(r1, DF_Q, b1.q); }
		if (!b2.fixed) { st3(r2, DF_X, b2.x); st4(r2, DF_Q, b2.q); }
		if (st) {
			atomicOr(&d.status[w], st);
			st = 0;
		}
	}
	if (!collisions) return;
	const int npf = d.lvl_fill[(size_t)level * RP_LVL_STRIDE];
	const int np = npf + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1];
	const int off0 = d.lvl_off[level], off1 = d.lvl_off[level + 1];
	WarpQueue q;
	q.init((unsigned int)np);
	bool have = false;
	int w = 0, cnt = 0, c = 0;
	real* cs = 0;
	DynRef r1, r2;
	r1.p = r2.p = 0; r1.s = r2.s = d.WS;
	V3 normal = v3(RL(0.0), RL(0.0), RL(0.0));
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			const int k = (int)got;
			const SolveItem item = d.lvl_items[k < npf ? off0 + k : off1 - 1 - (k - npf)];
			w = item.w;
			cnt = item.cnt;
			normal = item.normal;
			cs = contact_ptr(d, w, item.coff);
			load_static(b1, d, item.a);
			load_static(b2, d, item.b);
			r1 = dyn_ref(d, w, item.a);
			r2 = dyn_ref(d, w, item.b);
			b1.x = ld3(r1, DF_X); b1.q = ld4(r1, DF_Q);
			b2.x = ld3(r2, DF_X); b2.q = ld4(r2, DF_Q);
			c = 0;
			have = cnt > 0;
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			real* cp = cs + (size_t)c * 8 * d.WS;
			Contact ct = ld_contact(cp, d.WS);
			solve_contact(ct, normal, b1, b2, h, &st, PrevFromDyn{r1, r2});
			cp[6 * (size_t)d.WS] = ct.lambda_n;
			cp[7 * (size_t)d.WS] = ct.lambda_t;
			if (++c == cnt) {
				if (!b1.fixed) { st3(r1, DF_X, b1.x); st4(r1, DF_Q, b1.q); }
				if (!b2.fixed) { st3(r2, DF_X, b2.x); st4(r2, DF_Q, b2.q); }
				if (st) {
					atomicOr(&d.status[w], st);
					st = 0;
				}
				have = false;
			}
		}
	}
}

// JOINTS = false is the build for scenes without external constraints: the joint solves (hinge, spherical, their libm
// calls) are three quarters of this kernel's 14 k instructions, and the contact loop already runs short of instruction
// cache at two warps per scheduler (ncu, round 1: no_instruction is its second largest stall).
// Deep schedules are mostly empty: a pair of compound bodies expands to (colliders x colliders) units that all share the
// two bodies and therefore chain (spot_storm: 11 x 11 = 121 units per body pair, sweeps 1300+ levels deep), of which a
// handful have contacts in any one substep. Walking every level just to find it empty costs two dependent loads each, so
// for schedules deeper than RP_LIVE_MIN every CTA first builds the ascending list of levels that have work (same inputs,
// same list in every CTA, so the grid barriers stay matched) and the sweep walks that.
#define RP_LIVE_MIN 64
#define RP_LIVE_MAX 4096
struct LiveLevels {
	int list[RP_LIVE_MAX];
	unsigned int mask[RP_LIVE_MAX / 32];
	int off[RP_LIVE_MAX / 32];
	int n;
};
// returns the number of listed levels, or -1 when the plain walk over 1..levels is to be used. The list lives in DYNAMIC
// shared memory that the host only asks for when the scene can have deep schedules (`enabled`): the sweeps of the headline
// scenes keep those 17 kB as L1.
template <bool JOINTS>
__device__ __forceinline__ int list_live_levels(const DevView& d, LiveLevels& s, int levels, int collisions, int enabled) {
	if (!enabled || levels <= RP_LIVE_MIN || levels > RP_LIVE_MAX) return -1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nchunks = (levels + 31) >> 5;
	for (int c = warp; c < nchunks; c += nwarps) {
		const int l = 1 + (c << 5) + lane;
		bool live = false;
		if (l <= levels) {
			if (collisions) live = d.lvl_fill[(size_t)l * RP_LVL_STRIDE] + d.lvl_fill[(size_t)l * RP_LVL_STRIDE + 1] > 0;
			if (JOINTS && l <= d.joint_levels) live = live || d.joint_lptr[l] > d.joint_lptr[l - 1];
		}
		const unsigned int m = __ballot_sync(0xffffffffu, live);
		if (lane == 0) s.mask[c] = m;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		int run = 0;
		for (int c = 0; c < nchunks; ++c) {
			s.off[c] = run;
			run += __popc(s.mask[c]);
		}
		s.n = run;
	}
	__syncthreads();
	for (int c = warp; c < nchunks; c += nwarps) {
		const unsigned int m = s.mask[c];
		if ((m >> lane) & 1u) s.list[s.off[c] + __popc(m & ((1u << lane) - 1u))] = 1 + (c << 5) + lane;
	}
	__syncthreads();
	return s.n;
}


// ------------------------------------------------------------------------------------------------------ dataflow sweeps
// A unit only has to wait for the units that touch one of its two bodies and come before it in the reference's order -- on each
// body those form a chain of strictly increasing levels. The level-major sweeps above hold the WHOLE grid at a barrier between two
// levels anyway (ncu, frame 40: 4.2 barrier-stall cycles per issued instruction, and 22.6 of 32 lanes active because every warp's
// share of a level ends in a partly filled trip). The dataflow form keeps the level-major item lists but replaces the barriers
// by the chains themselves:
//  * k_manifold marks, per body, the levels at which the body has a unit with contacts this substep (body_live: one 64-bit mask
//    per body and world, cleared by k_integrate); joints add the template-constant mask of their levels;
//  * all levels (and positional iterations) form ONE item sequence in level-major order; a warp claims the next 32 items with
//    one atomic on a global cursor and hands them to its lanes as they fall free (lanes stay full across level boundaries);
//  * a unit of level l waits, for each of its non-fixed bodies, until body_done says that the body's previous live unit (the
//    highest marked level below l; in a later positional iteration and with none below, the body's LAST unit of the iteration
//    before) has finished; a finished unit stores its bodies and then, with release semantics, (pass << 6 | level) into body_done
//    of both. Pass numbers grow from substep to substep (substep counter * 256 + iteration), so nothing is ever reset.
// A first version kept one counter per WORLD (units finished / units at lower levels): right, and as fast on copies of one scene,
// but a world's units of one level are spread over the whole level list, so its level was only complete when the list's tail
// was, and worlds with uneven work had 10 - 13 of 32 lanes running.
// No deadlock: items are claimed in sequence order and a warp hands its claimed items to lanes in order, so the lowest
// unfinished item of the sequence is always held by a lane, everything it waits for has a lower level or an earlier pass and has
// finished, and it runs (the grid is the cooperative launch's resident CTAs, so every claiming warp is running). Every body's
// units still run in level order, which is the reference's order: results are bit-identical to the barrier form. A lane that
// polls 2^16 times without success gives up, flags the world (RP_ST_SOLVER_SINGULAR), tells every other lane to stop waiting and
// proceeds: a logic error cannot hang the device.
struct FlowTables {
	int cum[RP_FLOW_LEVELS + 2];   // items of levels < l (cum[levels + 1] = all)
	int jn[RP_FLOW_LEVELS + 2];    // (joint, world) items that head level l: its joints x W
	int npf[RP_FLOW_LEVELS + 2], off0[RP_FLOW_LEVELS + 2], off1[RP_FLOW_LEVELS + 2];
};
// every CTA builds the same tables from the level fill counters; returns the number of items of one pass over the levels.
// A level's items are its joints in every world (world fastest: lane = world), then its contact units.
template <bool JOINTS>
__device__ __forceinline__ unsigned int flow_tables(const DevView& d, FlowTables& t, int levels, int collisions) {
	if (threadIdx.x == 0) {
		long long run = 0;
		for (int l = 1; l <= levels; ++l) {
			const int f = collisions ? d.lvl_fill[(size_t)l * RP_LVL_STRIDE] : 0, bk = collisions ? d.lvl_fill[(size_t)l * RP_LVL_STRIDE + 1] : 0;
			const int nj = JOINTS && l <= d.joint_levels ? d.joint_lptr[l] - d.joint_lptr[l - 1] : 0;
			t.cum[l] = (int)run;
			t.jn[l] = nj * d.W;
			t.npf[l] = f;
			t.off0[l] = d.lvl_off[l];
			t.off1[l] = d.lvl_off[l + 1];
			run += (long long)nj * d.W + f + bk;
			if (run > 0x7fffff00ll) run = 0x7fffff00ll;  // (the caller falls back to the barrier form)
		}
		t.cum[levels + 1] = (int)run;
	}
	__syncthreads();
	return (unsigned int)t.cum[levels + 1];
}
// The chain state is read with RELAXED loads and the state a unit then loads goes through L2 (ld3cg / ld4cg / __ldcg): an acquire
// load (or fence) is LD + CCTL.IVALL on sm_100a -- it drops the SM's whole L1 on every poll, for every warp on the SM, although
// the only lines that can be stale are the ones read through L2 anyway. Ordering: the producer's stores are performed at L2
// (release) before its body_done update; the consumer issues its loads only after the poll's value has come back and been
// tested, and they read L2.
__device__ __forceinline__ unsigned long long flow_poll(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void flow_signal(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// what body_done of `body` must have reached before a unit of `level` in pass `pass` (iteration `it` of its sweep) may run; 0 = nothing
template <bool JOINTS>
__device__ __forceinline__ unsigned long long flow_need(const DevView& d, int body, int w, int level, unsigned long long pass, int it) {
	if (d.bstat[body].fixed) return 0ull;
	unsigned long long m = d.body_live[bidx(d, body, w)];
	if (JOINTS) m |= d.joint_body_mask[body];
	const unsigned long long below = m & ((1ull << level) - 1ull);
	if (below) return (pass << 6) | (unsigned long long)(63 - __clzll((long long)below));
	if (it > 0 && m) return ((pass - 1ull) << 6) | (unsigned long long)(63 - __clzll((long long)m));
	return 0ull;
}
#define RP_FLOW_SPIN_LIMIT (1 << 16)
#define RP_FLOW_MAX_ITERS 250  // positional iterations per pass number block (pass = substep counter * 256 + iteration + 1)
// one failed poll: counts it; true when the lane should stop waiting -- it ran out of patience (and then raises the batch-wide
// "broken" word behind the cursors, which every other waiting lane looks at now and then), or someone else did
__device__ __forceinline__ bool flow_give_up(const DevView& d, int* spins) {
	unsigned int* broken = d.flow_cursor + 2;
	if (++*spins > RP_FLOW_SPIN_LIMIT) {
		atomicExch(broken, 1u);
		return true;
	}
	return (*spins & 255) == 0 && *reinterpret_cast<volatile unsigned int*>(broken) != 0u;
}
// a warp's claim on the item sequence: [next, end) is claimed and not yet handed to a lane
struct FlowQueue {
	unsigned int next, end, total;
	bool more;  // the global cursor may still have items
	__device__ __forceinline__ void init(unsigned int total_items) {
		next = end = 0u;
		total = total_items;
		more = total_items > 0u;
	}
	// every lane calls this; lanes with want == true get the next items in sequence order (0xffffffff: none left)
	__device__ __forceinline__ unsigned int take(bool want, unsigned int* cursor) {
		const unsigned int mask = __ballot_sync(0xffffffffu, want);
		if (mask == 0u) return 0xffffffffu;
		const unsigned int n = __popc(mask), rank = __popc(mask & ((1u << (threadIdx.x & 31)) - 1u));
		const unsigned int avail = end - next;
		unsigned int got = 0xffffffffu;
		if (avail >= n || !more) {
			if (want && rank < avail) got = next + rank;
			next += n < avail ? n : avail;
			return got;
		}
		unsigned int base = 0u;
		if ((threadIdx.x & 31) == 0) base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		unsigned int end2 = base + 32u;
		if (base >= total) {
			more = false;
			base = end2 = total;
		} else if (end2 > total) {
			end2 = total;
		}
		if (want) {
			if (rank < avail) got = next + rank;
			else if (base + (rank - avail) < end2) got = base + (rank - avail);
		}
		const unsigned int used = n - avail;
		next = base + used < end2 ? base + used : end2;
		end = end2;
		return got;
	}
	__device__ __forceinline__ bool drained() const { return next >= end && !more; }
};
// item g of the sequence -> iteration, level, and either the slot of the level-major contact list (return value >= 0) or,
// for one of the level's joint items, -1 and its index among them (*joint_item: joint-in-level * W + world)
__device__ __forceinline__ int flow_locate(const FlowTables& t, int levels, unsigned int per_pass, unsigned int g, int* it, int* level, int* joint_item) {
	const unsigned int pass = g / per_pass;
	const int k = (int)(g - pass * per_pass);
	int l = 1;
	while (l < levels && k >= t.cum[l + 1]) ++l;
	*it = (int)pass;
	*level = l;
	int j = k - t.cum[l];
	if (j < t.jn[l]) {
		*joint_item = j;
		return -1;
	}
	j -= t.jn[l];
	return j < t.npf[l] ? t.off0[l] + j : t.off1[l] - 1 - (j - t.npf[l]);
}

// positional sweep, dataflow form: every level of every iteration in one pass (see above). A joint item is solved in one trip
// of the loop, a contact unit in one trip per contact.
template <bool JOINTS>
__device__ __forceinline__ void pos_flow(const DevView& d, const FlowTables& t, real h, int levels, unsigned int per_pass, int iters) {
	FlowQueue q;
	q.init(per_pass * (unsigned int)iters);
	const unsigned long long pass0 = (unsigned long long)(unsigned int)*d.epoch * 256ull + 1ull;  // pass number of iteration 0
	unsigned long long* const done = d.body_done;
	int st = 0;
	bool have = false, ready = false;
	int w = 0, cnt = 0, c = 0, ia = 0, ib = 0, spins = 0;
	int ju = -1;  // the joint a lane holds (JOINTS), -1: a contact unit
	unsigned long long need1 = 0ull, need2 = 0ull, mine = 0ull;
	real* cs = 0;
	DynRef r1, r2;
	r1.p = r2.p = 0; r1.s = r2.s = d.WS;
	V3 normal = v3(RL(0.0), RL(0.0), RL(0.0));
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have, d.flow_cursor);
		if (got != 0xffffffffu) {
			int it, level, jitem = 0;
			const int slot = flow_locate(t, levels, per_pass, got, &it, &level, &jitem);
			if (JOINTS && slot < 0) {
				const unsigned int jl = (unsigned int)jitem / (unsigned int)d.W;
				w = (int)((unsigned int)jitem - jl * (unsigned int)d.W);
				ju = d.joint_sched[d.joint_lptr[level - 1] + (int)jl];
				const Joint j = d.joints[ju];
				ia = j.e1; ib = j.e2;
				cnt = 1;
			} else {
				const SolveItem item = d.lvl_items[slot];
				w = item.w;
				cnt = item.cnt;
				normal = item.normal;
				ia = item.a; ib = item.b;
				cs = contact_ptr(d, w, item.coff);
				ju = -1;
			}
			// how far the chains of its two bodies must have got, and what it will write there itself
			need1 = flow_need<JOINTS>(d, ia, w, level, pass0 + (unsigned long long)it, it);
			need2 = flow_need<JOINTS>(d, ib, w, level, pass0 + (unsigned long long)it, it);
			mine = ((pass0 + (unsigned long long)it) << 6) | (unsigned long long)level;
			have = cnt > 0;
			ready = false;
			spins = 0;
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.drained()) break;
			continue;
		}
		if (have && !ready) {
			bool go = (need1 == 0ull || flow_poll(done + bidx(d, ia, w)) >= need1) && (need2 == 0ull || flow_poll(done + bidx(d, ib, w)) >= need2);
			if (!go && flow_give_up(d, &spins)) {
				st |= ST_SOLVER_SINGULAR;
				go = true;
			}
			if (go) {
				load_static(b1, d, ia);
				load_static(b2, d, ib);
				r1 = dyn_ref(d, w, ia);
				r2 = dyn_ref(d, w, ib);
				b1.x = ld3cg(r1, DF_X); b1.q = ld4cg(r1, DF_Q);
				b2.x = ld3cg(r2, DF_X); b2.q = ld4cg(r2, DF_Q);
				c = 0;
				ready = true;
			}
		}
		if (have && ready) {
			if (JOINTS && ju >= 0) {
				const Joint j = d.joints[ju];
				const real* lp = &d.lambdas[(size_t)ju * d.WS + w].a;  // (an earlier iteration's value may come from another SM)
				JointLambda lam;
				lam.a = __ldcg(lp); lam.b = __ldcg(lp + 1); lam.c = __ldcg(lp + 2);
				solve_joint(j, lam, b1, b2, h, &st);
				d.lambdas[(size_t)ju * d.WS + w] = lam;
				c = cnt;
			} else {
				real* cp = cs + (size_t)c * 8 * d.WS;
				Contact ct = ld_contact(cp, d.WS);
				ct.lambda_n = __ldcg(cp + 6 * (size_t)d.WS);  // (written by the previous iteration's unit, possibly on another SM)
				ct.lambda_t = __ldcg(cp + 7 * (size_t)d.WS);
				solve_contact(ct, normal, b1, b2, h, &st, PrevFromDyn{r1, r2});
				cp[6 * (size_t)d.WS] = ct.lambda_n;
				cp[7 * (size_t)d.WS] = ct.lambda_t;
				++c;
			}
			if (c == cnt) {
				if (!b1.fixed) { st3(r1, DF_X, b1.x); st4(r1, DF_Q, b1.q); }
				if (!b2.fixed) { st3(r2, DF_X, b2.x); st4(r2, DF_Q, b2.q); }
				if (st) {
					atomicOr(&d.status[w], st);
					st = 0;
				}
				if (!b1.fixed) flow_signal(done + bidx(d, ia, w), mine);
				if (!b2.fixed) flow_signal(done + bidx(d, ib, w), mine);
				have = false;
			}
		}
	}
}

template <bool JOINTS>
__global__ void RP_POS_BOUNDS k_solve_pos(DevView d, real h, int iters, int collisions, int live_lists) {
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) unsigned char s_live_raw[];
	LiveLevels& s_live = *reinterpret_cast<LiveLevels*>(s_live_raw);
	const int levels = *d.lvl_max;  // this frame's sweep depth over all worlds (k_schedule): read here, the host never needs it
	if (d.flow_mode && levels <= RP_FLOW_LEVELS) {
		__shared__ FlowTables s_flow;
		const unsigned int per_pass = flow_tables<JOINTS>(d, s_flow, levels, collisions);
		if (per_pass == 0u) return;
		if ((unsigned long long)per_pass * (unsigned long long)iters < 0x7fffff00ull && iters <= RP_FLOW_MAX_ITERS) {  // (same decision in every CTA)
			pos_flow<JOINTS>(d, s_flow, h, levels, per_pass, iters);
			return;
		}
	}
	const int n_live = list_live_levels<JOINTS>(d, s_live, levels, collisions, live_lists);
	const int trips = n_live >= 0 ? n_live : levels;
	bool dirty = false;
	for (int it = 0; it < iters; ++it) {
		for (int i = 0; i < trips; ++i) {
			const int level = n_live >= 0 ? s_live.list[i] : i + 1;
			const int nj = JOINTS && level <= d.joint_levels ? d.joint_lptr[level] - d.joint_lptr[level - 1] : 0;
			const int np = collisions ? d.lvl_fill[(size_t)level * RP_LVL_STRIDE] + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1] : 0;
			if (nj == 0 && np == 0) continue;
			if (dirty) grid.sync();
			pos_level<JOINTS>(d, h, level, nj, collisions);
			dirty = true;
		}
	}
}

// velocity derivation (pbd.cpp:623-643), one thread per body, for the bodies whose velocities are still pending at the
// end of the frame (not touched by a velocity-level unit of the last substep). Derivation is a pure function of the
// body's own (x, q, prev x, prev q, v, w), none of which the velocity pass writes before deriving, so WHEN it happens
// between the positional sweep and the first read of v/w does not change a bit.
__global__ void __launch_bounds__(128) k_derive(DevView d, real h) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	const int epoch = *d.epoch;
	// every body leaves the frame stamped "current" (a body that wakes up next frame must not look pending)
	if (d.vstamp[bidx(d, b, w)] == epoch) return;
	d.vstamp[bidx(d, b, w)] = epoch;
	Body body;
	body.fixed = d.bstat[b].fixed;
	body.active = d.active[bidx(d, b, w)];
	if (body.fixed || !body.active) return;
	const DynRef r = dyn_ref(d, w, b);
	body.x = ld3(r, DF_X); body.q = ld4(r, DF_Q); body.px = ld3(r, DF_PX); body.pq = ld4(r, DF_PQ);
	body.v = ld3(r, DF_V); body.w = ld3(r, DF_W);
	derive_velocity(body, h);
	st3(r, DF_V, body.v); st3(r, DF_W, body.w); st3(r, DF_PV, body.pv); st3(r, DF_PW, body.pw);
}

// Loads what the velocity pass needs of one body. If the body's velocities have not been derived in this substep yet
// (first velocity-level unit that touches it), derives them here (pbd.cpp:623-643), stores the previous velocities the
// derivation leaves (they are part of the body's state) and stamps the body; returns true if it did. The previous
// velocities are kept in registers only if the restitution term will read them (`need_prev`).
// CG: velocities, previous velocities and the stamp are read through L2 (dataflow form: other SMs write them during the launch).
template <bool CG = false>
__device__ __forceinline__ bool load_for_velocity(Body& b, const DynRef& r, int active, int* stamp, int epoch, real h, bool need_prev) {
	b.q = ld4(r, DF_Q);
	if (CG) { b.v = ld3cg(r, DF_V); b.w = ld3cg(r, DF_W); } else { b.v = ld3(r, DF_V); b.w = ld3(r, DF_W); }
	b.active = active;
	if (!(b.fixed || !b.active) && (CG ? __ldcg(stamp) : *stamp) != epoch) {
		b.x = ld3(r, DF_X); b.px = ld3(r, DF_PX); b.pq = ld4(r, DF_PQ);
		derive_velocity(b, h);
		st3(r, DF_PV, b.pv); st3(r, DF_PW, b.pw);
		*stamp = epoch;
		return true;
	}
	if (need_prev) {
		if (CG) { b.pv = ld3cg(r, DF_PV); b.pw = ld3cg(r, DF_PW); } else { b.pv = ld3(r, DF_PV); b.pw = ld3(r, DF_PW); }
	}
	return false;
}

// velocity pass over the contacts of one level (pbd.cpp:646-711); the hinge branch of the reference's velocity pass is
// an empty TODO (pbd.cpp:712-739), so joints take no part
__device__ __forceinline__ void vel_level(const DevView& d, real h, int level) {
	const int npf = d.lvl_fill[(size_t)level * RP_LVL_STRIDE];
	const int np = npf + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1];
	const int off0 = d.lvl_off[level], off1 = d.lvl_off[level + 1];
	WarpQueue q;
	q.init((unsigned int)np);
	bool have = false;
	int cnt = 0, c = 0;
	const real* cs = 0;
	DynRef r1, r2;
	r1.p = r2.p = 0; r1.s = r2.s = d.WS;
	AngPre tens;
	tens.ii1 = tens.ii2 = zero_m3();
	const int epoch = *d.epoch;
	V3 normal = v3(RL(0.0), RL(0.0), RL(0.0));
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have);
		if (got != 0xffffffffu) {
			const int k = (int)got;
			const SolveItem item = d.lvl_items[k < npf ? off0 + k : off1 - 1 - (k - npf)];
			const int w = item.w;
			cnt = item.cnt;
			if (cnt > 0) {  // (k_manifold only lists pairs with contacts)
				const PairRec pr = {item.a, item.b, 0, 0};
				normal = item.normal;
				cs = contact_ptr(d, w, item.coff);
				load_static(b1, d, pr.a);
				load_static(b2, d, pr.b);
				r1 = dyn_ref(d, w, pr.a);
				r2 = dyn_ref(d, w, pr.b);
				// restitution 0 on either side: the velocity solve never reads the previous velocities (solve_contact_velocity)
				const bool need_prev = b1.rest * b2.rest != RL(0.0);
				load_for_velocity(b1, r1, d.active[bidx(d, pr.a, w)], d.vstamp + bidx(d, pr.a, w), epoch, h, need_prev);
				load_for_velocity(b2, r2, d.active[bidx(d, pr.b, w)], d.vstamp + bidx(d, pr.b, w), epoch, h, need_prev);
				tens = vel_tensors(b1, b2);
				c = 0;
				have = true;
			}
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.empty()) break;
			continue;
		}
		if (have) {
			const Contact ct = ld_contact(cs + (size_t)c * 8 * d.WS, d.WS);
			solve_contact_velocity(ct, normal, b1, b2, h, tens);
			if (++c == cnt) {
				if (!b1.fixed) { st3(r1, DF_V, b1.v); st3(r1, DF_W, b1.w); }
				if (!b2.fixed) { st3(r2, DF_V, b2.v); st3(r2, DF_W, b2.w); }
				have = false;
			}
		}
	}
}


// velocity pass, dataflow form (the prefix table is the one the positional kernel of this substep left)
__device__ __forceinline__ void vel_flow(const DevView& d, const FlowTables& t, real h, int levels, unsigned int per_pass) {
	FlowQueue q;
	q.init(per_pass);
	unsigned long long* const done = d.body_done + (size_t)d.NB * d.WS;
	const unsigned long long pass = (unsigned long long)(unsigned int)*d.epoch * 256ull + 1ull;
	bool have = false, ready = false;
	int w = 0, cnt = 0, c = 0, ia = 0, ib = 0, spins = 0;
	unsigned long long need1 = 0ull, need2 = 0ull, mine = 0ull;
	const real* cs = 0;
	DynRef r1, r2;
	r1.p = r2.p = 0; r1.s = r2.s = d.WS;
	AngPre tens;
	tens.ii1 = tens.ii2 = zero_m3();
	const int epoch = *d.epoch;
	V3 normal = v3(RL(0.0), RL(0.0), RL(0.0));
	Body b1, b2;
	b1.fixed = b2.fixed = 1;
	for (;;) {
		const unsigned int got = q.take(!have, d.flow_cursor + 1);
		if (got != 0xffffffffu) {
			int it, level, jitem;
			const SolveItem item = d.lvl_items[flow_locate(t, levels, per_pass, got, &it, &level, &jitem)];
			w = item.w;
			cnt = item.cnt;
			normal = item.normal;
			ia = item.a; ib = item.b;
			cs = contact_ptr(d, w, item.coff);
			need1 = flow_need<false>(d, ia, w, level, pass, 0);  // (joints take no part in the velocity pass)
			need2 = flow_need<false>(d, ib, w, level, pass, 0);
			mine = (pass << 6) | (unsigned long long)level;
			have = cnt > 0;
			ready = false;
			spins = 0;
		}
		if (!__any_sync(0xffffffffu, have)) {
			if (q.drained()) break;
			continue;
		}
		if (have && !ready) {
			bool go = (need1 == 0ull || flow_poll(done + bidx(d, ia, w)) >= need1) && (need2 == 0ull || flow_poll(done + bidx(d, ib, w)) >= need2);
			if (!go && flow_give_up(d, &spins)) {
				atomicOr(&d.status[w], (int)ST_SOLVER_SINGULAR);
				go = true;
			}
			if (go) {
				load_static(b1, d, ia);
				load_static(b2, d, ib);
				r1 = dyn_ref(d, w, ia);
				r2 = dyn_ref(d, w, ib);
				const bool need_prev = b1.rest * b2.rest != RL(0.0);
				load_for_velocity<true>(b1, r1, d.active[bidx(d, ia, w)], d.vstamp + bidx(d, ia, w), epoch, h, need_prev);
				load_for_velocity<true>(b2, r2, d.active[bidx(d, ib, w)], d.vstamp + bidx(d, ib, w), epoch, h, need_prev);
				tens = vel_tensors(b1, b2);
				c = 0;
				ready = true;
			}
		}
		if (have && ready) {
			const Contact ct = ld_contact(cs + (size_t)c * 8 * d.WS, d.WS);
			solve_contact_velocity(ct, normal, b1, b2, h, tens);
			if (++c == cnt) {
				if (!b1.fixed) { st3(r1, DF_V, b1.v); st3(r1, DF_W, b1.w); }
				if (!b2.fixed) { st3(r2, DF_V, b2.v); st3(r2, DF_W, b2.w); }
				if (!b1.fixed) flow_signal(done + bidx(d, ia, w), mine);
				if (!b2.fixed) flow_signal(done + bidx(d, ib, w), mine);
				have = false;
			}
		}
	}
}

__global__ void __launch_bounds__(RP_VEL_THREADS, RP_MINB_VEL) k_solve_vel(DevView d, real h, int live_lists, int flow_iters) {
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) unsigned char s_live_raw[];
	LiveLevels& s_live = *reinterpret_cast<LiveLevels*>(s_live_raw);
	const int levels = *d.lvl_max;
	// (flow_iters = the positional iterations of this substep's k_solve_pos: the same decision as there, or 0 for the barrier form)
	if (flow_iters > 0 && d.flow_mode && levels <= RP_FLOW_LEVELS) {
		__shared__ FlowTables s_flow;
		// (the same bound as the positional kernel's, whose item count includes the joints)
		const unsigned int per_pass = flow_tables<false>(d, s_flow, levels, 1);
		if (per_pass == 0u) return;
		if (((unsigned long long)per_pass + (unsigned long long)d.NJ * d.W) * (unsigned long long)flow_iters < 0x7fffff00ull && flow_iters <= RP_FLOW_MAX_ITERS) {
			vel_flow(d, s_flow, h, levels, per_pass);
			return;
		}
	}
	const int n_live = list_live_levels<false>(d, s_live, levels, 1, live_lists);
	const int trips = n_live >= 0 ? n_live : levels;
	bool dirty = false;
	for (int i = 0; i < trips; ++i) {
		const int level = n_live >= 0 ? s_live.list[i] : i + 1;
		if (d.lvl_fill[(size_t)level * RP_LVL_STRIDE] + d.lvl_fill[(size_t)level * RP_LVL_STRIDE + 1] == 0) continue;
		if (dirty) grid.sync();
		vel_level(d, h, level);
		dirty = true;
	}
}

// --------------------------------------------------------------------------------------------------- world-block sweeps
// Dependencies between constraints never cross worlds, so a barrier between two levels only has to hold back the threads
// that work on the SAME worlds. k_solve_block gives every CTA a block of `wpb` consecutive worlds and walks that block's
// levels -- every positional iteration, then the velocity pass -- with __syncthreads() between them: no grid-wide barrier
// (ncu, round 2 start: 45 % of k_solve_pos's and 30 % of k_solve_vel's stall samples sat in grid.sync()), no cooperative
// launch holding whole SMs while most of the grid waits, one launch per substep for both sweeps, and blocks that finish
// early make room for the next ones. The CTA first sorts its worlds' live units (k_manifold's per-world lists) by level
// into its region of blk_items: histogram and cursors in shared memory; lanes of a warp that hold the same level take
// consecutive slots, so the `wpb` worlds of one pair stay adjacent and their loads of the world-minor arrays share sectors.
// Results are bit-identical to the level-major sweeps: the same units run in the same per-world level order.
#define RP_SB_THREADS 128
#ifndef RP_MINB_SB
#define RP_MINB_SB 3
#endif
#define RP_SB_MAX_WPB 64

__device__ __forceinline__ void pos_unit(const DevView& d, real h, int w, int pair, int* st) {
	const size_t pg = pidx(d, pair, w);
	const int2 ab = *reinterpret_cast<const int2*>(&d.pairs[pg]);
	const int cnt = d.pair_ccnt[pg];
	const V3 normal = d.pair_normal[pg];
	real* cs = contact_ptr(d, w, d.pair_coff[pg]);
	Body b1, b2;
	load_static(b1, d, ab.x);
	load_static(b2, d, ab.y);
	const DynRef r1 = dyn_ref(d, w, ab.x);
	const DynRef r2 = dyn_ref(d, w, ab.y);
	b1.x = ld3(r1, DF_X); b1.q = ld4(r1, DF_Q);
	b2.x = ld3(r2, DF_X); b2.q = ld4(r2, DF_Q);
	for (int c = 0; c < cnt; ++c) {
		real* cp = cs + (size_t)c * 8 * d.WS;
		Contact ct = ld_contact(cp, d.WS);
		solve_contact(ct, normal, b1, b2, h, st, PrevFromDyn{r1, r2});
		cp[6 * (size_t)d.WS] = ct.lambda_n;
		cp[7 * (size_t)d.WS] = ct.lambda_t;
	}
	if (!b1.fixed) { st3(r1, DF_X, b1.x); st4(r1, DF_Q, b1.q); }
	if (!b2.fixed) { st3(r2, DF_X, b2.x); st4(r2, DF_Q, b2.q); }
}

__device__ __forceinline__ void vel_unit(const DevView& d, real h, int w, int pair, int epoch) {
	const size_t pg = pidx(d, pair, w);
	const int2 ab = *reinterpret_cast<const int2*>(&d.pairs[pg]);
	const int cnt = d.pair_ccnt[pg];
	const V3 normal = d.pair_normal[pg];
	const real* cs = contact_ptr(d, w, d.pair_coff[pg]);
	Body b1, b2;
	load_static(b1, d, ab.x);
	load_static(b2, d, ab.y);
	const DynRef r1 = dyn_ref(d, w, ab.x);
	const DynRef r2 = dyn_ref(d, w, ab.y);
	const bool need_prev = b1.rest * b2.rest != RL(0.0);  // restitution 0 on either side: the previous velocities are never read
	load_for_velocity(b1, r1, d.active[bidx(d, ab.x, w)], d.vstamp + bidx(d, ab.x, w), epoch, h, need_prev);
	load_for_velocity(b2, r2, d.active[bidx(d, ab.y, w)], d.vstamp + bidx(d, ab.y, w), epoch, h, need_prev);
	const AngPre tens = vel_tensors(b1, b2);
	for (int c = 0; c < cnt; ++c) {
		const Contact ct = ld_contact(cs + (size_t)c * 8 * d.WS, d.WS);
		solve_contact_velocity(ct, normal, b1, b2, h, tens);
	}
	if (!b1.fixed) { st3(r1, DF_V, b1.v); st3(r1, DF_W, b1.w); }
	if (!b2.fixed) { st3(r2, DF_V, b2.v); st3(r2, DF_W, b2.w); }
}

template <bool JOINTS>
__global__ void __launch_bounds__(RP_SB_THREADS, RP_MINB_SB) k_solve_block(DevView d, real h, int iters, int collisions, int wpb) {
	extern __shared__ __align__(16) int s_lv[];  // [max_levels + 2]: counts -> starts -> ends of the block's levels
	__shared__ int s_nl[RP_SB_MAX_WPB];
	__shared__ int s_max, s_nlmax;
	const int tid = threadIdx.x, lane = tid & 31;
	const int w0 = blockIdx.x * wpb;
	const int nw = d.W - w0 < wpb ? d.W - w0 : wpb;
	unsigned int* items = d.blk_items + (size_t)w0 * d.max_pairs;
	int lmax = 0;
	if (collisions) {
		for (int l = tid; l < d.max_levels + 2; l += RP_SB_THREADS) s_lv[l] = 0;
		if (tid == 0) { s_max = 0; s_nlmax = 0; }
		__syncthreads();
		if (tid < nw) {
			const int n = d.n_live[w0 + tid];
			s_nl[tid] = n;
			atomicMax(&s_nlmax, n);
		}
		__syncthreads();
		const int total = s_nlmax * nw;  // entry e = (slot e / nw of world e % nw): a pair's worlds are neighbours
		for (int e = tid; e < total; e += RP_SB_THREADS) {
			const int slot = e / nw, wl = e - slot * nw;
			if (slot < s_nl[wl]) {
				const int lvl = (int)d.live[(size_t)slot * d.WS + w0 + wl].y;
				atomicAdd(&s_lv[lvl], 1);
				atomicMax(&s_max, lvl);
			}
		}
		__syncthreads();
		lmax = s_max;
		if (tid < 32) {  // exclusive scan of the counts of levels 1..lmax
			int run = 0;
			for (int base = 1; base <= lmax; base += 32) {
				const int l = base + lane;
				const int v = l <= lmax ? s_lv[l] : 0;
				int inc = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int t = __shfl_up_sync(0xffffffffu, inc, o);
					if (lane >= o) inc += t;
				}
				if (l <= lmax) s_lv[l] = run + inc - v;
				run += __shfl_sync(0xffffffffu, inc, 31);
			}
		}
		__syncthreads();
		const int rounds = (total + RP_SB_THREADS - 1) / RP_SB_THREADS;
		for (int r = 0; r < rounds; ++r) {
			const int e = r * RP_SB_THREADS + tid;
			int lvl = -1;
			unsigned int enc = 0u;
			if (e < total) {
				const int slot = e / nw, wl = e - slot * nw;
				if (slot < s_nl[wl]) {
					const uint2 rec = d.live[(size_t)slot * d.WS + w0 + wl];
					lvl = (int)rec.y;
					enc = (rec.x << 6) | (unsigned int)wl;
				}
			}
			const unsigned int peers = __match_any_sync(0xffffffffu, lvl);
			if (lvl > 0) {
				const int leader = __ffs(peers) - 1;
				int base = 0;
				if (lane == leader) base = atomicAdd(&s_lv[lvl], __popc(peers));
				base = __shfl_sync(peers, base, leader);
				items[base + __popc(peers & ((1u << lane) - 1u))] = enc;
			}
		}
		__syncthreads();  // s_lv[l] is now the END of level l (the start of level l + 1); s_lv[0] = 0
	}
	const int jl = JOINTS ? d.joint_levels : 0;
	const int levels = lmax > jl ? lmax : jl;
	int st = 0;
	for (int it = 0; it < iters; ++it) {
		for (int level = 1; level <= levels; ++level) {
			bool any = false;
			if (JOINTS && level <= jl) {
				const int j0 = d.joint_lptr[level - 1], nj = d.joint_lptr[level] - j0;
				any = nj > 0;
				for (int i = tid; i < nj * nw; i += RP_SB_THREADS) {
					const int ju = i / nw, w = w0 + (i - ju * nw);
					const int u = d.joint_sched[j0 + ju];
					const Joint j = d.joints[u];
					Body b1, b2;
					load_static(b1, d, j.e1);
					load_static(b2, d, j.e2);
					const DynRef r1 = dyn_ref(d, w, j.e1);
					const DynRef r2 = dyn_ref(d, w, j.e2);
					b1.x = ld3(r1, DF_X); b1.q = ld4(r1, DF_Q);
					b2.x = ld3(r2, DF_X); b2.q = ld4(r2, DF_Q);
					JointLambda lam = d.lambdas[(size_t)u * d.WS + w];
					solve_joint(j, lam, b1, b2, h, &st);
					d.lambdas[(size_t)u * d.WS + w] = lam;
					if (!b1.fixed) { st3(r1, DF_X, b1.x); st4(r1, DF_Q, b1.q); }
					if (!b2.fixed) { st3(r2, DF_X, b2.x); st4(r2, DF_Q, b2.q); }
					if (st) {
						atomicOr(&d.status[w], st);
						st = 0;
					}
				}
			}
			if (collisions && level <= lmax) {
				const int beg = s_lv[level - 1], end = s_lv[level];
				any = any || end > beg;
				for (int i = beg + tid; i < end; i += RP_SB_THREADS) {
					const unsigned int enc = items[i];
					const int w = w0 + (int)(enc & 63u);
					pos_unit(d, h, w, (int)(enc >> 6), &st);
					if (st) {
						atomicOr(&d.status[w], st);
						st = 0;
					}
				}
			}
			if (any) __syncthreads();
		}
	}
	if (!collisions) return;
	// velocity pass over the contacts (pbd.cpp:646-711), levels in the same order; joints take no part (pbd.cpp:712-739)
	const int epoch = *d.epoch;
	for (int level = 1; level <= lmax; ++level) {
		const int beg = s_lv[level - 1], end = s_lv[level];
		if (end == beg) continue;
		for (int i = beg + tid; i < end; i += RP_SB_THREADS) {
			const unsigned int enc = items[i];
			vel_unit(d, h, w0 + (int)(enc & 63u), (int)(enc >> 6), epoch);
		}
		__syncthreads();
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

// rp_batch_create_from: the per-world state of body `map[i]` of another batch becomes the state of body i of this one
__global__ void __launch_bounds__(128) k_adopt_bodies(DevView dst, DevView src, const int* map) {
	int w, b;
	if (!flat_item_world(dst, dst.NB, &b, &w)) return;
	const int o = map[b];
	if (o < 0) return;
	const DynRef to = dyn_ref(dst, w, b), from = dyn_ref(src, w, o);
#pragma unroll
	for (int f = 0; f < RP_DYN_DOUBLES; ++f) to.p[f * to.s] = from.p[f * from.s];
	dst.active[bidx(dst, b, w)] = src.active[bidx(src, o, w)];
	dst.deact[bidx(dst, b, w)] = src.deact[bidx(src, o, w)];
	dst.vstamp[bidx(dst, b, w)] = *dst.epoch;  // between frames every body's velocities are current (k_derive)
}

// OR of the capacity bits of every world's status word (rp_batch_sync and the other synchronising calls report it)
__global__ void __launch_bounds__(256) k_status_overflow(DevView d, int* out) {
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	int bits = w < d.W ? d.status[w] & (ST_EPA_CAPACITY | ST_CLIP_CAPACITY | ST_CONTACT_CAPACITY | ST_PAIR_CAPACITY) : 0;
	for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
	if ((threadIdx.x & 31) == 0 && bits) atomicOr(out, bits);
}

// -------------------------------------------------------------------------------------------------- state pack/unpack
// host record (rawphys_b200.h RP_STATE_STRIDE = 21 doubles, [world][body]) <-> world-minor dynamic state + active +
// deactivation time. Thread = (world, body) with lane = world: the device side is coalesced, the record side strided.
__global__ void __launch_bounds__(128) k_unpack_state(DevView d, const double* rec, int first_world, int n_worlds, int broadcast) {
	const int wl = blockIdx.y * blockDim.x + threadIdx.x;
	const int b = blockIdx.x;
	if (wl >= n_worlds) return;
	const double* r = rec + (broadcast ? (size_t)b : (size_t)wl * d.NB + b) * 21;  // (the host record is double whatever `real` is)
	const int w = first_world + wl;
	const DynRef o = dyn_ref(d, w, b);
	st3(o, DF_X, v3((real)r[0], (real)r[1], (real)r[2])); st4(o, DF_Q, q4((real)r[3], (real)r[4], (real)r[5], (real)r[6]));
	st3(o, DF_V, v3((real)r[7], (real)r[8], (real)r[9])); st3(o, DF_W, v3((real)r[10], (real)r[11], (real)r[12]));
	st3(o, DF_PX, v3((real)r[0], (real)r[1], (real)r[2])); st4(o, DF_PQ, q4((real)r[3], (real)r[4], (real)r[5], (real)r[6]));
	st3(o, DF_PV, v3((real)r[15], (real)r[16], (real)r[17])); st3(o, DF_PW, v3((real)r[18], (real)r[19], (real)r[20]));
	d.active[bidx(d, b, w)] = r[13] != 0.0 ? 1 : 0;
	d.deact[bidx(d, b, w)] = (real)r[14];
	d.vstamp[bidx(d, b, w)] = *d.epoch;  // uploaded velocities are current
}
__global__ void __launch_bounds__(128) k_pack_state(DevView d, double* rec, int first_world, int n_worlds) {
	const int wl = blockIdx.y * blockDim.x + threadIdx.x;
	const int b = blockIdx.x;
	if (wl >= n_worlds) return;
	double* r = rec + ((size_t)wl * d.NB + b) * 21;
	const int w = first_world + wl;
	const DynRef o = dyn_ref(d, w, b);
	const V3 x = ld3(o, DF_X), v = ld3(o, DF_V), om = ld3(o, DF_W), pv = ld3(o, DF_PV), pw = ld3(o, DF_PW);
	const Q4 q = ld4(o, DF_Q);
	r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = q.x; r[4] = q.y; r[5] = q.z; r[6] = q.w;
	r[7] = v.x; r[8] = v.y; r[9] = v.z; r[10] = om.x; r[11] = om.y; r[12] = om.z;
	r[13] = d.active[bidx(d, b, w)] ? 1.0 : 0.0;
	r[14] = d.deact[bidx(d, b, w)];
	r[15] = pv.x; r[16] = pv.y; r[17] = pv.z; r[18] = pw.x; r[19] = pw.y; r[20] = pw.z;
}

}  // namespace rp
#endif
