// rp_hull.cuh -- convex-hull topology built on the device. Included once, by rp_batch.cu.
//
// collider_convex_hull_create (src/physics/collider.cpp:194-364) is quadratic in the triangle count on one host thread:
// the vertex merge, the triangle-shares-a-vertex lists (:249-262), the vertex neighbour lists and the face neighbour lists
// (:330-352) each compare everything with everything. Its OUTPUT ORDER is parity-defining (vertex order decides support
// ties, face / neighbour order decides clipping), so the device build must produce the very arrays build_hull()
// (rp_scene.cpp) does. What is order-DEFINING and sequential -- the depth-first flood over coplanar triangles
// (collect_faces_planar_to, :47-78) -- is linear in the neighbour lists and stays one thread; everything quadratic is one
// thread per output row:
//   k_hull_first      vertex i -> first vertex with the same coordinates            (O(n^2) compares, thread per vertex)
//   k_hull_tris       remapped triangles, their normals, per-vertex incidence counts
//   k_hull_tnbr       triangle i -> triangles j != i sharing a vertex, ascending j   (count pass, scan, fill pass)
//   k_hull_v2n        vertex -> neighbour vertices in the reference's push order      (thread per vertex over all triangles)
//   k_hull_flood      ONE thread: seeds in index order, explicit-stack DFS in neighbour-list order -> triangle sequence, faces
//   k_hull_loops      thread per face: boundary loop by edge toggling + the reference's (quirky) edge ordering (:94-166)
//   k_hull_v2f        vertex -> face of every incident triangle, in flood sequence order (duplicates kept, :300-304)
//   k_hull_f2n        face i -> faces j != i sharing a vertex, ascending j            (count pass, scan, fill pass)
// Arithmetic (triangle normals) is the shared FP64 core under --fmad=false: the normals are the host's bit for bit.
#ifndef RP_HULL_CUH
#define RP_HULL_CUH

#include <cuda_runtime.h>

#include <vector>

#include "rp_scene.h"

namespace rp {

struct HullBuildView {
	int n_in, T;             // input vertices, triangles
	const double* vin;       // [n_in][3]
	const unsigned int* idx; // [3 T]
	int* first;              // [n_in] first input vertex with equal coordinates
	int* is_first;           // [n_in]
	int* vid;                // [n_in + 1] exclusive scan of is_first -> hull vertex id of a first occurrence; [n_in] = V
	V3* verts;               // [V]
	int3* tris;              // [T] hull vertex ids
	V3* tnorm;               // [T]
	int* deg;                // [n_in + 1] triangles incident to hull vertex v (then its exclusive scan: v2f_ptr)
	int* tn_cnt;             // [T + 1] -> scan = tn_ptr
	int* tn_ptr;
	int* tn_idx;             // triangle neighbour lists
	int* v2n_cap;            // [V + 1] scan of 2 * deg: scratch offsets
	int* v2n_tmp;            // [6 T] scratch lists
	int* v2n_cnt;            // [V + 1] -> scan = v2n_ptr
	int* v2n_ptr;
	int* v2n_idx;
	int* seq;                // [T] triangles in flood order
	int* order;              // [T] position of triangle t in seq
	int* tri_face;           // [T]
	int* face_start;         // [T + 1] face f = seq[face_start[f] .. face_start[f + 1])
	int* n_faces;            // [1]
	int* stack;              // [2 T] DFS frames
	V3* fnorm;               // [F]
	int2* edges;             // [3 T] scratch of the boundary loops (face f's region starts at 3 * face_start[f])
	int* loop_cnt;           // [T + 1] -> scan = face_ptr
	int* face_ptr;
	int* face_idx;
	int* v2f_ptr;            // = scan of deg
	int* v2f_idx;            // [3 T]
	int* f2n_cnt;            // [T + 1] -> scan
	int* f2n_ptr;
	int* f2n_idx;
};

__global__ void __launch_bounds__(128) k_hull_first(HullBuildView h) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= h.n_in) return;
	const V3 p = v3(h.vin[3 * i], h.vin[3 * i + 1], h.vin[3 * i + 2]);
	int f = i;
	for (int j = 0; j < i; ++j) {
		if (equal(v3(h.vin[3 * j], h.vin[3 * j + 1], h.vin[3 * j + 2]), p)) {
			f = j;
			break;
		}
	}
	h.first[i] = f;
	h.is_first[i] = f == i ? 1 : 0;
}

// exclusive scan of in[0..n) into out[0..n], out[n] = total; one CTA (the arrays here are a few thousand entries)
__global__ void __launch_bounds__(1024) k_hull_scan(const int* in, int* out, int n) {
	__shared__ int s_warp[32];
	__shared__ int s_run;
	if (threadIdx.x == 0) s_run = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = i < n ? in[i] : 0;
		int inc = v;
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += t;
		}
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			int ws = s_warp[lane];
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, ws, o);
				if (lane >= o) ws += t;
			}
			s_warp[lane] = ws;
		}
		__syncthreads();
		const int before = s_run + (warp > 0 ? s_warp[warp - 1] : 0);
		if (i < n) out[i] = before + inc - v;
		__syncthreads();
		if (threadIdx.x == 1023) s_run = before + inc;
		__syncthreads();
	}
	if (threadIdx.x == 0) out[n] = s_run;
}

__global__ void __launch_bounds__(128) k_hull_verts(HullBuildView h) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= h.n_in) return;
	if (h.is_first[i]) h.verts[h.vid[i]] = v3(h.vin[3 * i], h.vin[3 * i + 1], h.vin[3 * i + 2]);
	h.deg[i] = 0;
}

__global__ void __launch_bounds__(128) k_hull_tris(HullBuildView h) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= h.T) return;
	int3 tr;
	tr.x = h.vid[h.first[h.idx[3 * t]]];
	tr.y = h.vid[h.first[h.idx[3 * t + 1]]];
	tr.z = h.vid[h.first[h.idx[3 * t + 2]]];
	h.tris[t] = tr;
	const V3 v1 = h.verts[tr.x], v2 = h.verts[tr.y], v3_ = h.verts[tr.z];
	h.tnorm[t] = normalize(cross(sub(v2, v1), sub(v3_, v1)));  // find_triangle_normal (collider.cpp:180-192)
	atomicAdd(&h.deg[tr.x], 1);
	atomicAdd(&h.deg[tr.y], 1);
	atomicAdd(&h.deg[tr.z], 1);
}

__device__ __forceinline__ bool hull_tris_share(int3 a, int3 b) {  // do_triangles_share_same_vertex (collider.cpp:27-31)
	return a.x == b.x || a.x == b.y || a.x == b.z || a.y == b.x || a.y == b.y || a.y == b.z || a.z == b.x || a.z == b.y || a.z == b.z;
}
template <bool FILL>
__global__ void __launch_bounds__(128) k_hull_tnbr(HullBuildView h) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= h.T) return;
	const int3 a = h.tris[i];
	int n = 0;
	int* out = FILL ? h.tn_idx + h.tn_ptr[i] : 0;
	for (int j = 0; j < h.T; ++j) {
		if (j != i && hull_tris_share(a, h.tris[j])) {
			if (FILL) out[n] = j;
			++n;
		}
	}
	if (!FILL) h.tn_cnt[i] = n;
}

// vertex_to_neighbors (collider.cpp:264-296): triangles in order, each pushes its other two vertices unless already listed
__global__ void __launch_bounds__(128) k_hull_v2n(HullBuildView h, int V) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= V) return;
	int* list = h.v2n_tmp + h.v2n_cap[v];
	int n = 0;
	for (int t = 0; t < h.T; ++t) {
		const int3 tr = h.tris[t];
		int a, b;
		if (tr.x == v) { a = tr.y; b = tr.z; }
		else if (tr.y == v) { a = tr.x; b = tr.z; }
		else if (tr.z == v) { a = tr.x; b = tr.y; }
		else continue;
		bool seen = false;
		for (int k = 0; k < n && !seen; ++k) seen = list[k] == a;
		if (!seen) list[n++] = a;
		seen = false;
		for (int k = 0; k < n && !seen; ++k) seen = list[k] == b;
		if (!seen) list[n++] = b;
		// (a triangle that names v twice lists it under both positions in the reference; degenerate input is refused upstream)
	}
	h.v2n_cnt[v] = n;
}
__global__ void __launch_bounds__(128) k_hull_v2n_pack(HullBuildView h, int V) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= V) return;
	const int* list = h.v2n_tmp + h.v2n_cap[v];
	int* out = h.v2n_idx + h.v2n_ptr[v];
	const int n = h.v2n_cnt[v];
	for (int k = 0; k < n; ++k) out[k] = list[k];
}
__global__ void __launch_bounds__(128) k_hull_double(const int* in, int* out, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = 2 * in[i];
}

// collect_faces_planar_to over every seed (collider.cpp:47-78, :298-316): the recursion visits a triangle's neighbours in list
// order and descends at once, so the explicit stack holds (triangle, next neighbour) frames
__global__ void k_hull_flood(HullBuildView h) {
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	const double EPS = 0.000001;
	int placed = 0, faces = 0;
	for (int t = 0; t < h.T; ++t) h.tri_face[t] = -1;
	for (int seed = 0; seed < h.T; ++seed) {
		if (h.tri_face[seed] >= 0) continue;
		const V3 target = h.tnorm[seed];
		h.face_start[faces] = placed;
		h.fnorm[faces] = target;
		int sp = 0;
		int cur = seed;
		// visit(cur): not done, and planar to the target -> take it and walk its neighbours
		for (;;) {
			bool descend = false;
			if (cur >= 0 && h.tri_face[cur] < 0) {
				const double proj = dot(h.tnorm[cur], target);
				if ((proj - 1.0) > -EPS && (proj - 1.0) < EPS) {
					h.seq[placed] = cur;
					h.order[cur] = placed;
					++placed;
					h.tri_face[cur] = faces;
					h.stack[2 * sp] = cur;
					h.stack[2 * sp + 1] = h.tn_ptr[cur];
					++sp;
					descend = true;
				}
			}
			(void)descend;
			// next neighbour of the innermost open frame
			cur = -1;
			while (sp > 0) {
				const int top = h.stack[2 * (sp - 1)];
				const int at = h.stack[2 * (sp - 1) + 1];
				if (at < h.tn_ptr[top + 1]) {
					h.stack[2 * (sp - 1) + 1] = at + 1;
					cur = h.tn_idx[at];
					break;
				}
				--sp;
			}
			if (cur < 0) break;
		}
		++faces;
	}
	h.face_start[faces] = placed;
	*h.n_faces = faces;
}

// create_convex_hull_face (collider.cpp:94-166) for one face per thread
__global__ void __launch_bounds__(64) k_hull_loops(HullBuildView h) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= *h.n_faces) return;
	const int s0 = h.face_start[f], s1 = h.face_start[f + 1];
	int2* e = h.edges + 3 * (size_t)s0;
	int n = 0;
	for (int s = s0; s < s1; ++s) {
		const int3 tr = h.tris[h.seq[s]];
		const int ex[3] = {tr.x, tr.y, tr.z}, ey[3] = {tr.y, tr.z, tr.x};
		for (int k = 0; k < 3; ++k) {
			int found = -1;
			for (int i = 0; i < n && found < 0; ++i) {
				if ((e[i].x == ex[k] && e[i].y == ey[k]) || (e[i].x == ey[k] && e[i].y == ex[k])) found = i;
			}
			if (found >= 0) {  // array_remove: swap with the last
				e[found] = e[n - 1];
				--n;
			} else {
				e[n++] = make_int2(ex[k], ey[k]);
			}
		}
	}
	for (int i = 0; i < n; ++i) {  // no early exit from the inner loop, as in the reference
		const int2 cur = e[i];
		for (int j = i + 1; j < n; ++j) {
			int2 cand = e[j];
			if (cur.y != cand.x && cur.y != cand.y) continue;
			if (cur.y == cand.y) {
				const int t = cand.x;
				cand.x = cand.y;
				cand.y = t;
			}
			const int2 tmp = e[i + 1];
			e[i + 1] = cand;
			e[j] = tmp;
		}
	}
	h.loop_cnt[f] = n;
}
__global__ void __launch_bounds__(64) k_hull_loops_pack(HullBuildView h) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= *h.n_faces) return;
	const int2* e = h.edges + 3 * (size_t)h.face_start[f];
	int* out = h.face_idx + h.face_ptr[f];
	const int n = h.loop_cnt[f];
	for (int i = 0; i < n; ++i) out[i] = e[i].x;
}

// vertex_to_faces (collider.cpp:300-304): every triangle of every face, in flood order, pushes its face to its three vertices
__global__ void __launch_bounds__(128) k_hull_v2f(HullBuildView h, int V) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= V) return;
	int* out = h.v2f_idx + h.v2f_ptr[v];
	int n = 0;
	// incident triangles sorted by their position in the flood sequence (insertion sort on short lists); a triangle naming v
	// k times pushes k entries
	for (int t = 0; t < h.T; ++t) {
		const int3 tr = h.tris[t];
		const int times = (tr.x == v ? 1 : 0) + (tr.y == v ? 1 : 0) + (tr.z == v ? 1 : 0);
		for (int r = 0; r < times; ++r) {
			int k = n++;
			while (k > 0 && h.order[out[k - 1]] > h.order[t]) {
				out[k] = out[k - 1];
				--k;
			}
			out[k] = t;
		}
	}
	for (int k = 0; k < n; ++k) out[k] = h.tri_face[out[k]];
}

// face_to_neighbors (collider.cpp:330-352)
template <bool FILL>
__global__ void __launch_bounds__(128) k_hull_f2n(HullBuildView h) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int F = *h.n_faces;
	if (i >= F) return;
	const int a0 = h.face_ptr[i], a1 = h.face_ptr[i + 1];
	int n = 0;
	int* out = FILL ? h.f2n_idx + h.f2n_ptr[i] : 0;
	for (int j = 0; j < F; ++j) {
		if (j == i) continue;
		const int b0 = h.face_ptr[j], b1 = h.face_ptr[j + 1];
		bool share = false;
		for (int a = a0; a < a1 && !share; ++a) {
			const int va = h.face_idx[a];
			for (int b = b0; b < b1 && !share; ++b) share = h.face_idx[b] == va;
		}
		if (share) {
			if (FILL) out[n] = j;
			++n;
		}
	}
	if (!FILL) h.f2n_cnt[i] = n;
}

// host driver: returns false (and the CUDA error text) on failure
struct HullScratch {
	std::vector<void*> ptrs;
	~HullScratch() {
		for (size_t i = 0; i < ptrs.size(); ++i) cudaFree(ptrs[i]);
	}
	template <class T>
	bool get(T** p, size_t n) {
		void* q = 0;
		if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return false;
		ptrs.push_back(q);
		*p = (T*)q;
		return true;
	}
};

static bool build_hull_device(int device, const double* vx, uint32_t nverts, const uint32_t* indices, uint32_t nidx, HullHost* out, float* ms_out,
	std::string* err) {
#define HB_CHECK(call)                                                          \
	do {                                                                        \
		cudaError_t e_ = (call);                                                \
		if (e_ != cudaSuccess) {                                                \
			*err = std::string(#call) + ": " + cudaGetErrorString(e_);          \
			return false;                                                       \
		}                                                                       \
	} while (0)
	HB_CHECK(cudaSetDevice(device));
	const int n = (int)nverts, T = (int)(nidx / 3);
	HullScratch sc;
	HullBuildView h;
	memset(&h, 0, sizeof(h));
	h.n_in = n;
	h.T = T;
	double* vin;
	unsigned int* idx;
	bool ok = sc.get(&vin, 3 * (size_t)n) && sc.get(&idx, 3 * (size_t)T) && sc.get(&h.first, n) && sc.get(&h.is_first, n) && sc.get(&h.vid, n + 1) &&
	          sc.get(&h.verts, n) && sc.get(&h.tris, T) && sc.get(&h.tnorm, T) && sc.get(&h.deg, n + 1) && sc.get(&h.tn_cnt, T + 1) &&
	          sc.get(&h.tn_ptr, T + 1) && sc.get(&h.v2n_cap, n + 1) && sc.get(&h.v2n_tmp, 6 * (size_t)T) && sc.get(&h.v2n_cnt, n + 1) &&
	          sc.get(&h.v2n_ptr, n + 1) && sc.get(&h.v2n_idx, 6 * (size_t)T) && sc.get(&h.seq, T) && sc.get(&h.order, T) && sc.get(&h.tri_face, T) &&
	          sc.get(&h.face_start, T + 1) && sc.get(&h.n_faces, 1) && sc.get(&h.stack, 2 * (size_t)T + 2) && sc.get(&h.fnorm, T) &&
	          sc.get(&h.edges, 3 * (size_t)T) && sc.get(&h.loop_cnt, T + 1) && sc.get(&h.face_ptr, T + 1) && sc.get(&h.face_idx, 3 * (size_t)T) &&
	          sc.get(&h.v2f_ptr, n + 1) && sc.get(&h.v2f_idx, 3 * (size_t)T) && sc.get(&h.f2n_cnt, T + 1) && sc.get(&h.f2n_ptr, T + 1);
	if (!ok) {
		*err = "build_hull_device: cudaMalloc failed";
		return false;
	}
	h.vin = vin;
	h.idx = idx;
	cudaStream_t st = 0;
	cudaEvent_t e0, e1;
	HB_CHECK(cudaEventCreate(&e0));
	HB_CHECK(cudaEventCreate(&e1));
	HB_CHECK(cudaMemcpyAsync(vin, vx, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
	HB_CHECK(cudaMemcpyAsync(idx, indices, sizeof(unsigned int) * 3 * (size_t)T, cudaMemcpyHostToDevice, st));
	HB_CHECK(cudaEventRecord(e0, st));
	const unsigned int gn = (unsigned int)((n + 127) / 128), gt = (unsigned int)((T + 127) / 128);
	k_hull_first<<<gn, 128, 0, st>>>(h);
	k_hull_scan<<<1, 1024, 0, st>>>(h.is_first, h.vid, n);
	k_hull_verts<<<gn, 128, 0, st>>>(h);
	k_hull_tris<<<gt, 128, 0, st>>>(h);
	int V = 0;
	HB_CHECK(cudaMemcpyAsync(&V, h.vid + n, sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	const unsigned int gv = (unsigned int)((V + 127) / 128);
	k_hull_scan<<<1, 1024, 0, st>>>(h.deg, h.v2f_ptr, V);
	k_hull_tnbr<false><<<gt, 128, 0, st>>>(h);
	k_hull_scan<<<1, 1024, 0, st>>>(h.tn_cnt, h.tn_ptr, T);
	int tn_total = 0;
	HB_CHECK(cudaMemcpyAsync(&tn_total, h.tn_ptr + T, sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	if (!sc.get(&h.tn_idx, (size_t)tn_total)) {
		*err = "build_hull_device: cudaMalloc failed";
		return false;
	}
	k_hull_tnbr<true><<<gt, 128, 0, st>>>(h);
	k_hull_double<<<gv, 128, 0, st>>>(h.deg, h.v2n_cnt, V);  // (v2n_cnt borrowed as the input of the capacity scan)
	k_hull_scan<<<1, 1024, 0, st>>>(h.v2n_cnt, h.v2n_cap, V);
	k_hull_v2n<<<gv, 128, 0, st>>>(h, V);
	k_hull_scan<<<1, 1024, 0, st>>>(h.v2n_cnt, h.v2n_ptr, V);
	k_hull_v2n_pack<<<gv, 128, 0, st>>>(h, V);
	k_hull_flood<<<1, 1, 0, st>>>(h);
	int F = 0;
	HB_CHECK(cudaMemcpyAsync(&F, h.n_faces, sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	const unsigned int gf = (unsigned int)((F + 63) / 64), gf128 = (unsigned int)((F + 127) / 128);
	k_hull_loops<<<gf, 64, 0, st>>>(h);
	k_hull_scan<<<1, 1024, 0, st>>>(h.loop_cnt, h.face_ptr, F);
	k_hull_loops_pack<<<gf, 64, 0, st>>>(h);
	k_hull_v2f<<<gv, 128, 0, st>>>(h, V);
	k_hull_f2n<false><<<gf128, 128, 0, st>>>(h);
	k_hull_scan<<<1, 1024, 0, st>>>(h.f2n_cnt, h.f2n_ptr, F);
	int f2n_total = 0;
	HB_CHECK(cudaMemcpyAsync(&f2n_total, h.f2n_ptr + F, sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	if (!sc.get(&h.f2n_idx, (size_t)f2n_total)) {
		*err = "build_hull_device: cudaMalloc failed";
		return false;
	}
	k_hull_f2n<true><<<gf128, 128, 0, st>>>(h);
	HB_CHECK(cudaEventRecord(e1, st));
	HB_CHECK(cudaGetLastError());

	HullHost r;
	r.key.assign(vx, vx + 3 * (size_t)nverts);
	r.key_idx.assign(indices, indices + nidx);
	r.verts.resize(V);
	r.normals.resize(F);
	r.face_ptr.resize(F + 1);
	r.v2f_ptr.resize(V + 1);
	r.v2n_ptr.resize(V + 1);
	r.f2n_ptr.resize(F + 1);
	HB_CHECK(cudaMemcpyAsync(r.verts.data(), h.verts, sizeof(V3) * V, cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.normals.data(), h.fnorm, sizeof(V3) * F, cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.face_ptr.data(), h.face_ptr, sizeof(int) * (F + 1), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.v2f_ptr.data(), h.v2f_ptr, sizeof(int) * (V + 1), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.v2n_ptr.data(), h.v2n_ptr, sizeof(int) * (V + 1), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.f2n_ptr.data(), h.f2n_ptr, sizeof(int) * (F + 1), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	r.face_idx.resize(r.face_ptr[F]);
	r.v2f_idx.resize(r.v2f_ptr[V]);
	r.v2n_idx.resize(r.v2n_ptr[V]);
	r.f2n_idx.resize(r.f2n_ptr[F]);
	HB_CHECK(cudaMemcpyAsync(r.face_idx.data(), h.face_idx, sizeof(int) * r.face_idx.size(), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.v2f_idx.data(), h.v2f_idx, sizeof(int) * r.v2f_idx.size(), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.v2n_idx.data(), h.v2n_idx, sizeof(int) * r.v2n_idx.size(), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaMemcpyAsync(r.f2n_idx.data(), h.f2n_idx, sizeof(int) * r.f2n_idx.size(), cudaMemcpyDeviceToHost, st));
	HB_CHECK(cudaStreamSynchronize(st));
	if (ms_out) HB_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*out = r;
	return true;
#undef HB_CHECK
}

}  // namespace rp
#endif
