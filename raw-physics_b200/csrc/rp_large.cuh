// rp_large.cuh -- the per-frame prologue for ONE LARGE SCENE (thousands to ~10^5 bodies per world, few worlds): the
// pieces of the frame step whose batched-worlds forms are per-world sequential or quadratic in bodies per world.
//
//   * uniform-grid broadphase: bodies hashed into cells of edge >= 2 r + 0.1 (r = largest radius among the "small" bodies),
//     a one-digit radix (counting) sort of the bodies on their cell key, a sweep of the 27 neighbouring cells per body, and an
//     ORDERED compaction -- per-row counts, an exclusive scan over the rows, rows sorted by j -- so that the pair list is
//     exactly broad_get_collision_pairs' (broad.cpp:6-29): every i < j with |x_i - x_j| <= r_i + r_j + 0.1 in (i asc, j asc)
//     order, expanded to collider pairs. Bodies far larger than the rest (the floor: radius 70.7) would blow the cell size up;
//     they are kept out of the grid and tested against everything, one CTA per such body (block-wide ordered compaction).
//   * islands by union-find (compare-and-swap hooking of the larger root under the smaller, path halving), grid-wide: the partition
//     into islands does not depend on union order (broad.cpp:70-116), so this gives the reference's islands.
//   * graph colouring by independent sets (Jones-Plassmann with hashed priorities): a unit takes the lowest colour none of
//     its coloured neighbours holds once every neighbour of higher priority is coloured; deterministic, O(log n) rounds.
//     Used by the coloured solve order only (north star: "large-scene mode, graph-coloured Gauss-Seidel over a
//     contact-constraint colouring rebuilt each step"); the reference order keeps k_schedule's sequential recurrence.
//   * the (pair, world) form of k_cull for batches of fewer worlds than a warp.
// Included by rp_batch.cu after rp_kernels.cuh.
#ifndef RP_LARGE_CUH
#define RP_LARGE_CUH

namespace rp {

// ------------------------------------------------------------------------------------------------- multi-block scan
// Exclusive scan of data[world][0..n) (ints) in three launches: per-block scan + block totals, scan of the totals by one
// block, add-back. blockDim = RP_SCAN_THREADS, RP_SCAN_ITEMS consecutive elements per thread.
#define RP_SCAN_THREADS 256
#define RP_SCAN_ITEMS 8
#define RP_SCAN_TILE (RP_SCAN_THREADS * RP_SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {  // blockDim.x == RP_SCAN_THREADS
	__shared__ int s_warp[RP_SCAN_THREADS / 32];
	__shared__ int s_total;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) s_warp[wid] = inc;
	__syncthreads();
	if (wid == 0) {
		int w = lane < RP_SCAN_THREADS / 32 ? s_warp[lane] : 0;
		int winc = w;
#pragma unroll
		for (int o = 1; o < RP_SCAN_THREADS / 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, winc, o);
			if (lane >= o) winc += t;
		}
		if (lane < RP_SCAN_THREADS / 32) s_warp[lane] = winc - w;
		if (lane == RP_SCAN_THREADS / 32 - 1) s_total = winc;
	}
	__syncthreads();
	const int r = s_warp[wid] + inc - v;
	*total = s_total;
	__syncthreads();
	return r;
}

// phase 1: data -> per-tile exclusive scan, tile totals to sums[world][tile]
__global__ void __launch_bounds__(RP_SCAN_THREADS) k_scan_tiles(int* data, int n, size_t world_stride, int* sums, int tiles) {
	int* x = data + (size_t)blockIdx.y * world_stride;
	const int base = blockIdx.x * RP_SCAN_TILE + threadIdx.x * RP_SCAN_ITEMS;
	int v[RP_SCAN_ITEMS], run = 0;
#pragma unroll
	for (int k = 0; k < RP_SCAN_ITEMS; ++k) {
		v[k] = base + k < n ? x[base + k] : 0;
		run += v[k];
	}
	int total;
	int off = block_exclusive_scan(run, &total);
#pragma unroll
	for (int k = 0; k < RP_SCAN_ITEMS; ++k) {
		if (base + k < n) x[base + k] = off;
		off += v[k];
	}
	if (threadIdx.x == 0) sums[(size_t)blockIdx.y * tiles + blockIdx.x] = total;
}
// phase 2: exclusive scan of the tile totals of one world by one block (any number of tiles), grand total to totals[world]
__global__ void __launch_bounds__(RP_SCAN_THREADS) k_scan_sums(int* sums, int tiles, int* totals) {
	int* s = sums + (size_t)blockIdx.x * tiles;
	int carry = 0;
	for (int base = 0; base < tiles; base += RP_SCAN_THREADS) {
		const int i = base + threadIdx.x;
		const int v = i < tiles ? s[i] : 0;
		int total;
		const int off = block_exclusive_scan(v, &total);
		if (i < tiles) s[i] = carry + off;
		carry += total;
	}
	if (threadIdx.x == 0 && totals) totals[blockIdx.x] = carry;
}
// phase 3
__global__ void __launch_bounds__(RP_SCAN_THREADS) k_scan_add(int* data, int n, size_t world_stride, const int* sums, int tiles) {
	int* x = data + (size_t)blockIdx.y * world_stride;
	const int add = sums[(size_t)blockIdx.y * tiles + blockIdx.x];
	const int base = blockIdx.x * RP_SCAN_TILE + threadIdx.x * RP_SCAN_ITEMS;
#pragma unroll
	for (int k = 0; k < RP_SCAN_ITEMS; ++k) {
		if (base + k < n) x[base + k] += add;
	}
}

// ------------------------------------------------------------------------------------------------ grid broadphase
struct GridView {
	int table;            // buckets (power of two)
	real inv_cell;      // 1 / cell edge; cell edge = (2 * r_small_max + 0.1) * (1 + 1e-6)
	int n_large;          // bodies kept out of the grid
	const int* large;     // [n_large] their indices, ascending
	const unsigned char* is_large;  // [NB]
	int* bucket;          // [W][NB] bucket of every small body
	int* start;           // [W][table + 1] counts -> starts of the buckets in `sorted`
	int* cursor;          // [W][table]
	int* sorted;          // [W][NB] small bodies ordered by bucket
	int* row;             // [W][NB + 1] collider pairs per row i -> row offsets
	int* sums;            // scan scratch
	int* totals;          // [W]
};

__device__ __forceinline__ int3 grid_cell(const GridView& g, V3 x) {
	return make_int3((int)floor(x.x * g.inv_cell), (int)floor(x.y * g.inv_cell), (int)floor(x.z * g.inv_cell));
}
__device__ __forceinline__ int grid_bucket(const GridView& g, int cx, int cy, int cz) {
	unsigned int h = (unsigned int)cx * 73856093u ^ (unsigned int)cy * 19349663u ^ (unsigned int)cz * 83492791u;
	h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
	return (int)(h & (unsigned int)(g.table - 1));
}
__device__ __forceinline__ V3 body_x(const DevView& d, int w, int b) { return ld3(dyn_ref(d, w, b), DF_X); }
// broad.cpp:19-20 as written
__device__ __forceinline__ bool broad_near(V3 xi, real ri, V3 xj, real rj) {
	const V3 dv = sub(xi, xj);
	return sqrt(dv.x * dv.x + dv.y * dv.y + dv.z * dv.z) <= ri + rj + RL(0.1);
}

__global__ void __launch_bounds__(256) k_grid_count(DevView d, GridView g) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	if (g.is_large[b]) return;
	const int3 c = grid_cell(g, body_x(d, w, b));
	const int k = grid_bucket(g, c.x, c.y, c.z);
	g.bucket[(size_t)w * d.NB + b] = k;
	atomicAdd(&g.start[(size_t)w * (g.table + 1) + k], 1);
}
__global__ void __launch_bounds__(256) k_grid_fill(DevView d, GridView g) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	if (g.is_large[b]) return;
	const int k = g.bucket[(size_t)w * d.NB + b];
	const int slot = g.start[(size_t)w * (g.table + 1) + k] + atomicAdd(&g.cursor[(size_t)w * g.table + k], 1);
	g.sorted[(size_t)w * d.NB + slot] = b;
}

// Visits every j > i that is near body i (small i): the small bodies of the 27 neighbouring cells (distinct buckets only: two
// cells may share a bucket, and a body must be seen once) and the large bodies. f(j) is called in NO particular order.
template <class F>
__device__ __forceinline__ void grid_row_visit(const DevView& d, const GridView& g, int w, int i, F f) {
	const V3 xi = body_x(d, w, i);
	const real ri = d.bstat[i].radius;
	const int3 c = grid_cell(g, xi);
	int seen[27];
	int ns = 0;
	const int* start = g.start + (size_t)w * (g.table + 1);
	const int* sorted = g.sorted + (size_t)w * d.NB;
	for (int dz = -1; dz <= 1; ++dz) {
		for (int dy = -1; dy <= 1; ++dy) {
			for (int dx = -1; dx <= 1; ++dx) {
				const int k = grid_bucket(g, c.x + dx, c.y + dy, c.z + dz);
				bool dup = false;
				for (int s = 0; s < ns; ++s) dup = dup || seen[s] == k;
				if (dup) continue;
				seen[ns++] = k;
				for (int p = start[k]; p < start[k + 1]; ++p) {
					const int j = sorted[p];
					if (j > i && broad_near(xi, ri, body_x(d, w, j), d.bstat[j].radius)) f(j);
				}
			}
		}
	}
	for (int l = 0; l < g.n_large; ++l) {
		const int j = g.large[l];
		if (j > i && broad_near(xi, ri, body_x(d, w, j), d.bstat[j].radius)) f(j);
	}
}

// rows of the small bodies: collider pairs per row
__global__ void __launch_bounds__(128) k_grid_rowcount(DevView d, GridView g) {
	int w, i;
	if (!flat_item_world(d, d.NB, &i, &w)) return;
	if (g.is_large[i]) return;  // k_grid_large_count
	const int nci = d.bstat[i].ncol;
	int count = 0;
	grid_row_visit(d, g, w, i, [&](int j) { count += nci * d.bstat[j].ncol; });
	g.row[(size_t)w * (d.NB + 1) + i] = count;
}
// row of a large body: one CTA walks every j > i (blockIdx.x = index into the large list, blockIdx.y = world)
__global__ void __launch_bounds__(RP_SCAN_THREADS) k_grid_large_count(DevView d, GridView g) {
	const int i = g.large[blockIdx.x], w = blockIdx.y;
	const V3 xi = body_x(d, w, i);
	const real ri = d.bstat[i].radius;
	const int nci = d.bstat[i].ncol;
	int count = 0;
	for (int j = i + 1 + threadIdx.x; j < d.NB; j += RP_SCAN_THREADS) {
		if (broad_near(xi, ri, body_x(d, w, j), d.bstat[j].radius)) count += nci * d.bstat[j].ncol;
	}
	int total;
	block_exclusive_scan(count, &total);
	if (threadIdx.x == 0) g.row[(size_t)w * (d.NB + 1) + i] = total;
}

__device__ __forceinline__ void write_pairs(const DevView& d, int w, int i, int j, int* out) {
	const int ci0 = d.bstat[i].col0, nci = d.bstat[i].ncol, cj0 = d.bstat[j].col0, ncj = d.bstat[j].ncol;
	for (int a = 0; a < nci; ++a) {  // sub-collider i outer, j inner (collider.cpp:563-571)
		for (int b = 0; b < ncj; ++b) {
			if (*out < d.max_pairs) {
				PairRec pr;
				pr.a = i; pr.b = j; pr.ca = ci0 + a; pr.cb = cj0 + b;
				d.pairs[pidx(d, *out, w)] = pr;
			}
			++*out;
		}
	}
}

// rows of the small bodies, written in ascending j. A row is gathered into a small local list and sorted; a row longer than
// the list is emitted by rank (for every near j: how many near j' are smaller), which needs no storage.
#define RP_GRID_ROW_LOCAL 48
__global__ void __launch_bounds__(128) k_grid_rowwrite(DevView d, GridView g) {
	int w, i;
	if (!flat_item_world(d, d.NB, &i, &w)) return;
	if (g.is_large[i]) return;
	const int* row = g.row + (size_t)w * (d.NB + 1);
	int out = row[i];
	if (row[i + 1] == out) return;
	int list[RP_GRID_ROW_LOCAL];
	int n = 0;
	grid_row_visit(d, g, w, i, [&](int j) {
		if (n < RP_GRID_ROW_LOCAL) list[n] = j;
		++n;
	});
	if (n <= RP_GRID_ROW_LOCAL) {
		for (int a = 1; a < n; ++a) {  // insertion sort
			const int v = list[a];
			int b = a - 1;
			while (b >= 0 && list[b] > v) {
				list[b + 1] = list[b];
				--b;
			}
			list[b + 1] = v;
		}
		for (int a = 0; a < n; ++a) write_pairs(d, w, i, list[a], &out);
	} else {
		const int nci = d.bstat[i].ncol;
		grid_row_visit(d, g, w, i, [&](int j) {
			int before = 0;
			grid_row_visit(d, g, w, i, [&](int j2) {
				if (j2 < j) before += nci * d.bstat[j2].ncol;
			});
			int at = out + before;
			write_pairs(d, w, i, j, &at);
		});
	}
}
// row of a large body in ascending j: block-wide ordered compaction, RP_SCAN_THREADS candidates per trip
__global__ void __launch_bounds__(RP_SCAN_THREADS) k_grid_large_write(DevView d, GridView g) {
	const int i = g.large[blockIdx.x], w = blockIdx.y;
	const V3 xi = body_x(d, w, i);
	const real ri = d.bstat[i].radius;
	const int nci = d.bstat[i].ncol;
	int base = g.row[(size_t)w * (d.NB + 1) + i];
	for (int j0 = i + 1; j0 < d.NB; j0 += RP_SCAN_THREADS) {
		const int j = j0 + threadIdx.x;
		const bool near = j < d.NB && broad_near(xi, ri, body_x(d, w, j), d.bstat[j].radius);
		int total;
		int at = base + block_exclusive_scan(near ? nci * d.bstat[j].ncol : 0, &total);
		if (near) write_pairs(d, w, i, j, &at);
		base += total;
	}
}
// after the scan of the rows: the world's pair count (clamped to capacity) and the counters k_broad_scan maintains
__global__ void k_grid_finish(DevView d, GridView g) {
	const int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= d.W) return;
	int total = g.totals[w];
	g.row[(size_t)w * (d.NB + 1) + d.NB] = total;
	atomicAdd(&d.counters[CNT_BROAD_PAIRS], (unsigned long long)total);
	if (total > d.max_pairs) {
		atomicOr(&d.status[w], ST_PAIR_CAPACITY);
		total = d.max_pairs;
	}
	d.n_pairs[w] = total;
}

// -------------------------------------------------------------------------------------------------- islands, grid-wide
// broad_collect_simulation_islands (broad.cpp:70-116) + pbd.cpp:476-506 for worlds too large for one CTA's label
// propagation: union-find over {pairs, joints} among non-fixed bodies. label[] holds parent pointers; a union hooks the
// larger root under the smaller, so the final root of a component is its smallest body -- whatever the order of the unions.
__device__ __forceinline__ int uf_root(int* parent, int x) {
	// parent pointers are read past L1 (other SMs re-parent nodes while this runs; a stale pointer is still an ancestor, but
	// a stale "I am a root" would make the caller's compare-and-swap fail forever)
	int p = __ldcg(&parent[x]);
	while (p != x) {  // path halving
		const int gp = __ldcg(&parent[p]);
		if (gp != p) parent[x] = gp;
		x = p;
		p = gp;
	}
	return x;
}
__global__ void __launch_bounds__(256) k_uf_init(DevView d) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	d.label[(size_t)w * d.NB + b] = b;
	d.isl_flag[(size_t)w * d.NB + b] = 1;
}
// one pass over the edges: every edge loops until its two bodies share a root. A union is a compare-and-swap on a ROOT
// (a plain atomicMin could re-parent a node somebody else had just hooked and lose that union); path halving only ever
// writes non-roots, and a non-root never becomes a root again, so the two do not interfere.
__global__ void __launch_bounds__(256) k_uf_hook(DevView d) {
	int w, e;
	if (!flat_item_world(d, d.max_pairs + d.NJ, &e, &w)) return;
	int a, b;
	if (e < d.max_pairs) {
		if (e >= d.n_pairs[w]) return;
		const PairRec pr = d.pairs[pidx(d, e, w)];
		a = pr.a; b = pr.b;
	} else {
		a = d.joints[e - d.max_pairs].e1; b = d.joints[e - d.max_pairs].e2;
	}
	if (d.bstat[a].fixed || d.bstat[b].fixed) return;
	int* parent = d.label + (size_t)w * d.NB;
	for (;;) {
		const int ra = uf_root(parent, a), rb = uf_root(parent, b);
		if (ra == rb) break;
		const int hi = ra > rb ? ra : rb, lo = ra > rb ? rb : ra;
		if (atomicCAS(&parent[hi], hi, lo) == hi) break;
	}
}
__global__ void __launch_bounds__(256) k_uf_sleep(DevView d, real dt) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	if (d.bstat[b].fixed) return;
	int* parent = d.label + (size_t)w * d.NB;
	const int root = uf_root(parent, b);
	parent[b] = root;
	const DynRef r = dyn_ref(d, w, b);
	const real lv = length(ld3(r, DF_V)), av = length(ld3(r, DF_W));
	real t = d.deact[bidx(d, b, w)];
	if (lv < d.lin_sleep && av < d.ang_sleep) t += dt;
	else t = RL(0.0);
	d.deact[bidx(d, b, w)] = t;
	if (t < d.sleep_time) d.isl_flag[(size_t)w * d.NB + root] = 0;
}
__global__ void __launch_bounds__(256) k_uf_apply(DevView d) {
	int w, b;
	if (!flat_item_world(d, d.NB, &b, &w)) return;
	if (d.bstat[b].fixed) return;
	const int root = d.label[(size_t)w * d.NB + b];  // flattened by k_uf_sleep (roots point to themselves)
	d.active[bidx(d, b, w)] = d.isl_flag[(size_t)w * d.NB + uf_root(d.label + (size_t)w * d.NB, root)] ? 0 : 1;
}

// -------------------------------------------------------------------------------------------- colouring, grid-wide
// Coloured solve order for a large scene: units = broadphase (collider) pairs that pass the skip rule (pbd.cpp:594) + the
// joints (coloured on the host, constant); two units conflict when they share a non-fixed body. Per-body adjacency (CSR over
// the units touching each non-fixed body) is rebuilt every frame; then rounds of independent sets.
struct ColourView {
	int* deg;         // [W][NB + 1] units per body -> CSR offsets
	int* fill;        // [W][NB]
	int* adj;         // [W][2 * max_pairs] unit ids per body
	int* colour;      // [W][max_pairs] 0 = not yet coloured (committed colours), -1 = unit skipped
	int* pending;     // [W][max_pairs] colours decided in the current round
	int* remaining;   // [rounds + 1] units still uncoloured after round r (all worlds)
	int* sums;
	int* totals;
};
__device__ __forceinline__ unsigned int unit_priority(int u) {
	unsigned int h = (unsigned int)u * 0x9e3779b1u;
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
__device__ __forceinline__ bool unit_before(int u, int v) {  // u has priority over v
	const unsigned int pu = unit_priority(u), pv = unit_priority(v);
	return pu > pv || (pu == pv && u < v);
}
__global__ void __launch_bounds__(256) k_col_degree(DevView d, ColourView c, int collisions) {
	int w, p;
	if (!flat_item_world(d, d.max_pairs, &p, &w)) return;
	const int np = collisions ? d.n_pairs[w] : 0;
	if (p >= np) {
		c.colour[(size_t)w * d.max_pairs + p] = -1;
		return;
	}
	const PairRec pr = d.pairs[pidx(d, p, w)];
	const int fa = d.bstat[pr.a].fixed, fb = d.bstat[pr.b].fixed;
	const bool sa = fa || !d.active[bidx(d, pr.a, w)], sb = fb || !d.active[bidx(d, pr.b, w)];
	if (sa && sb) {  // pbd.cpp:594
		c.colour[(size_t)w * d.max_pairs + p] = -1;
		d.pair_level[pidx(d, p, w)] = 0;
		return;
	}
	c.colour[(size_t)w * d.max_pairs + p] = 0;
	if (!fa) atomicAdd(&c.deg[(size_t)w * (d.NB + 1) + pr.a], 1);
	if (!fb) atomicAdd(&c.deg[(size_t)w * (d.NB + 1) + pr.b], 1);
}
__global__ void __launch_bounds__(256) k_col_fill(DevView d, ColourView c) {
	int w, p;
	if (!flat_item_world(d, d.max_pairs, &p, &w)) return;
	if (c.colour[(size_t)w * d.max_pairs + p] != 0) return;
	const PairRec pr = d.pairs[pidx(d, p, w)];
	const int* off = c.deg + (size_t)w * (d.NB + 1);
	int* adj = c.adj + (size_t)w * 2 * d.max_pairs;
	if (!d.bstat[pr.a].fixed) adj[off[pr.a] + atomicAdd(&c.fill[(size_t)w * d.NB + pr.a], 1)] = p;
	if (!d.bstat[pr.b].fixed) adj[off[pr.b] + atomicAdd(&c.fill[(size_t)w * d.NB + pr.b], 1)] = p;
}
// one round: every uncoloured unit whose higher-priority neighbours are all coloured takes the lowest colour above the
// joints' that none of its coloured neighbours holds. Colours go to `pending` and are committed by k_col_commit, so a round
// only ever reads the previous rounds' colours (two neighbours are never both ready in one round: one of them has priority).
__global__ void __launch_bounds__(256) k_col_round(DevView d, ColourView c, int round) {
	if (round > 0 && c.remaining[round - 1] == 0) return;
	int w, p;
	if (!flat_item_world(d, d.max_pairs, &p, &w)) return;
	const int* colour = c.colour + (size_t)w * d.max_pairs;
	if (colour[p] != 0) return;
	const PairRec pr = d.pairs[pidx(d, p, w)];
	const int* off = c.deg + (size_t)w * (d.NB + 1);
	const int* adj = c.adj + (size_t)w * 2 * d.max_pairs;
	// Colours are numbered from 1 in one space shared with the joints (coloured on the host, SchedEntry<true>): `used` bit k =
	// colour k + 1 is held by a joint on one of this unit's bodies or by a neighbouring unit; colours past 64 are tracked by
	// their maximum only.
	unsigned long long used = 0ull;
	int beyond = 64;
	bool ready = true;
	const int bodies[2] = {pr.a, pr.b};
	for (int s = 0; s < 2 && ready; ++s) {
		const int b = bodies[s];
		if (d.bstat[b].fixed) continue;
		const unsigned long long jm = d.joint_colours[b];  // low word: colours 1..32 seen; high word: e = colours 33..32+e taken
		const unsigned long long e = jm >> 32;
		used |= (jm & 0xffffffffull) | ((e >= 32ull ? 0xffffffffull : ((1ull << e) - 1ull)) << 32);
		for (int k = off[b]; k < off[b + 1]; ++k) {
			const int v = adj[k];
			if (v == p) continue;
			const int cv = colour[v];
			if (cv == 0) {
				if (unit_before(v, p)) {
					ready = false;
					break;
				}
			} else if (cv > 0) {
				if (cv <= 64) used |= 1ull << (cv - 1);
				else if (cv > beyond) beyond = cv;
			}
		}
	}
	if (!ready) {
		atomicAdd(&c.remaining[round], 1);
		return;
	}
	const int col = used != ~0ull ? __ffsll((long long)~used) : beyond + 1;
	c.pending[(size_t)w * d.max_pairs + p] = col;
}
__global__ void __launch_bounds__(256) k_col_commit(DevView d, ColourView c, int round) {
	if (round > 0 && c.remaining[round - 1] == 0) return;
	int w, p;
	if (!flat_item_world(d, d.max_pairs, &p, &w)) return;
	const int col = c.pending[(size_t)w * d.max_pairs + p];
	if (col > 0 && c.colour[(size_t)w * d.max_pairs + p] == 0) {
		c.colour[(size_t)w * d.max_pairs + p] = col;
		d.pair_level[pidx(d, p, w)] = col;
		atomicAdd(&d.lvl_cap[col], 1);
		atomicMax(d.lvl_max, col);
	}
}
// units the fixed number of rounds left uncoloured (never seen; the expected number of rounds is O(log n)) get colours of
// their own past everything else, in index order: still a valid colouring
__global__ void __launch_bounds__(256) k_col_leftover(DevView d, ColourView c, int rounds) {
	if (c.remaining[rounds - 1] == 0) return;
	int w, p;
	if (!flat_item_world(d, d.max_pairs, &p, &w)) return;
	if (c.colour[(size_t)w * d.max_pairs + p] != 0) return;
	const int col = d.max_levels - atomicAdd(&c.remaining[rounds], 1);  // from the top of the level space, downwards
	if (col > 4096) {
		c.colour[(size_t)w * d.max_pairs + p] = col;
		d.pair_level[pidx(d, p, w)] = col;
		atomicAdd(&d.lvl_cap[col], 1);
		atomicMax(d.lvl_max, col);
	} else {
		d.pair_level[pidx(d, p, w)] = 0;
		atomicOr(&d.status[w], ST_PAIR_CAPACITY);
	}
}
__global__ void k_col_levels(DevView d) {  // the joints' colours count too; CNT_LEVELS as k_schedule maintains it
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		atomicMax(d.lvl_max, d.joint_levels);
		atomicAdd(&d.counters[CNT_LEVELS], (unsigned long long)*d.lvl_max * d.W);
	}
}

// ------------------------------------------------------------------------------------------------- cull, (pair, world)
// k_cull for batches of fewer worlds than a warp: one thread per (pair, world) instead of lane = world. Same skip rule, same
// exact-safe bounds test, same outputs.
__global__ void __launch_bounds__(256) k_cull_flat(DevView d, int cull) {
	int w = 0, p = 0;
	const bool in_range = flat_item_world(d, d.max_pairs, &p, &w);
	const bool in = in_range && p < d.n_pairs[w];
	bool keep = false, big = false;
	PairRec pr;
	pr.a = pr.b = pr.ca = pr.cb = 0;
	const size_t S = d.WS;
	if (in) {
		pr = d.pairs[pidx(d, p, w)];
		const int fa = d.bstat[pr.a].fixed, fb = d.bstat[pr.b].fixed;
		const int aa = d.active[bidx(d, pr.a, w)], ab = d.active[bidx(d, pr.b, w)];
		keep = !((fa || !aa) && (fb || !ab));
		d.pair_ccnt[pidx(d, p, w)] = 0;
	}
	int tested = keep ? 1 : 0;
	if (keep && cull && !(d.cols[pr.ca].type == SHAPE_SPHERE && d.cols[pr.cb].type == SHAPE_SPHERE)) {
		const float* pa = d.aabb + (size_t)pr.ca * 6 * S + w;
		const float* pb = d.aabb + (size_t)pr.cb * 6 * S + w;
#pragma unroll
		for (int ax = 0; ax < 3; ++ax) {
			const real lo_a = pa[ax * S], hi_a = pa[(3 + ax) * S], lo_b = pb[ax * S], hi_b = pb[(3 + ax) * S];
			if (lo_a - hi_b > RP_CULL_MARGIN || lo_b - hi_a > RP_CULL_MARGIN) keep = false;
		}
	}
	if (keep) big = warp_pair_verts(d.cols[pr.ca].nv + d.cols[pr.cb].nv);
	const unsigned int front = warp_append(d.cand_count, keep && !big);
	const unsigned int back = warp_append(d.big_count, big);
	if (keep) {
		const unsigned int slot = big ? d.cand_cap - 1u - back : front;
		d.cands[slot] = make_uint4((unsigned int)w, (unsigned int)p, (unsigned int)pr.ca, (unsigned int)pr.cb);
#if defined(RP_STORED_NORMALS)
		d.geom_stamp[(size_t)pr.ca * S + w] = *d.epoch;
		d.geom_stamp[(size_t)pr.cb * S + w] = *d.epoch;
#endif
	}
	for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(0xffffffffu, tested, o);
	if ((threadIdx.x & 31) == 0 && tested) atomicAdd(&d.counters[CNT_PAIR_TESTS], (unsigned long long)tested);
}

}  // namespace rp
#endif
