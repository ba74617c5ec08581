// gcm_collide.cuh -- collision discovery on the device: BruteforceCollisionDetector::find_collisions, the
// local/local part (gcm/system/BruteforceCollisionDetector.cpp:7-92), with the helpers it calls
// (CollisionDetector.cpp:34-103, TetrMesh_1stOrder::vector_intersects_triangle / interpolate_triangle,
// gcm/meshtypes/TetrMesh_1stOrder.cpp:2194-2311).
//
// What is static is laid out once on the host (gcm_host.cpp: build_collision_plan): mesh outlines, hence the
// intersection box of every ordered pair of local meshes, the border nodes of the first mesh and the border faces
// of the second inside it (both in the reference's list order), and a 2-D grid over those faces.  What changes
// every step -- the direction ksi[0] of each node's fresh random basis and the values on the faces -- is read on
// the device, so a non-static detector needs no host round trip:
//   collide_clear_kernel   clear_contact_state (TetrMeshSet.cpp:116-117)
//   collide_find_kernel    per candidate node: the FIRST face of the list whose plane the node's line meets inside
//                          the triangle (:44-84).  Only the faces of the grid bins the projected line crosses are
//                          tried, each bin in ascending list order, keeping the smallest passing position: the face
//                          the full scan would return.
//   collide_number_kernel  virt_nodes.size() at the moment each hit is pushed: an exclusive scan over the
//                          candidates in (pair, node) order -- the order of the reference's loops
//   collide_build_kernel   the virtual node: point on the face, 13 interpolated values, borrowed base node;
//                          contact flag and contact_data->axis_plus[0] of the real node (the later pair wins, as
//                          the reference's overwrite does)
// The routines are __host__ __device__: the host detector (find_collisions, used by host-only sets and the CPU
// emulation tests) runs the same arithmetic.
#pragma once
#include "gcm_step.cuh"

namespace gcm {

// TetrMesh_1stOrder::find_border_elem_normal's cross product (:333-370)
GCM_HD void tri_normal(const float* p1, const float* p2, const float* p3, float n[3])
{
	float v0x = p2[0] - p1[0], v0y = p2[1] - p1[1], v0z = p2[2] - p1[2];
	float v1x = p3[0] - p1[0], v1y = p3[1] - p1[1], v1z = p3[2] - p1[2];
	n[0] = v0y * v1z - v0z * v1y;
	n[1] = v0z * v1x - v0x * v1z;
	n[2] = v0x * v1y - v0y * v1x;
	float l = sqrtf( n[0] * n[0] + n[1] * n[1] + n[2] * n[2] );
	n[0] /= l;
	n[1] /= l;
	n[2] /= l;
}

GCM_HD bool bad_float(float x) { return !(fabsf(x) <= 3.4028235e38f); }    // NaN or Inf

// vector_intersects_triangle (:2194-2290).  As written there, the distance / direction test (t < 0 || t > l) and
// the three same_orientation tests are overwritten before use: the answer is the area test alone (plus vn != 0).
// Returns 1 (p = meeting point), 0, or -(1..6) for the reference's "NAN or INF in p1 / p2 / p3 / p0 / v / n" throws.
GCM_HD int vector_intersects_triangle(const float* p1, const float* p2, const float* p3, const float* p0, const float* v, float* p)
{
	const float* chk[5] = {p1, p2, p3, p0, v};
	for(int a = 0; a < 5; a++)
		for(int i = 0; i < 3; i++)
			if(bad_float(chk[a][i])) return -(a + 1);
	float n[3];
	tri_normal(p1, p2, p3, n);
	for(int i = 0; i < 3; i++)
		if(bad_float(n[i])) return -6;
	float vn = n[0]*v[0] + n[1]*v[1] + n[2]*v[2];
	if(vn == 0) return 0;
	float d = - (n[0]*p1[0] + n[1]*p1[1] + n[2]*p1[2]);
	float t = - ((n[0]*p0[0] + n[1]*p0[1] + n[2]*p0[2]) + d) / vn;
	for(int i = 0; i < 3; i++) p[i] = p0[i] + t * v[i];
	float area = fabsf( tri_area(p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2], p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]) );
	float areas[3];
	areas[0] = fabsf( tri_area(p3[0] - p[0], p3[1] - p[1], p3[2] - p[2], p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]) );
	areas[1] = fabsf( tri_area(p1[0] - p[0], p1[1] - p[1], p1[2] - p[2], p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]) );
	areas[2] = fabsf( tri_area(p3[0] - p[0], p3[1] - p[1], p3[2] - p[2], p1[0] - p[0], p1[1] - p[1], p1[2] - p[2]) );
	if( (double)fabsf(areas[0] + areas[1] + areas[2] - area) > (double)area * 0.0001 )
		return 0;
	return 1;
}

// interpolate_triangle (:2292-2311); false on "NAN or INF in virt node coords"
GCM_HD bool interpolate_triangle(const float* p1, const float* p2, const float* p3, const float* p,
                                 const float* v1, const float* v2, const float* v3, float* v, int n)
{
	for(int i = 0; i < 3; i++)
		if(bad_float(p[i])) return false;
	float areas[3];
	areas[0] = fabsf( tri_area(p3[0] - p[0], p3[1] - p[1], p3[2] - p[2], p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]) );
	areas[1] = fabsf( tri_area(p1[0] - p[0], p1[1] - p[1], p1[2] - p[2], p2[0] - p[0], p2[1] - p[1], p2[2] - p[2]) );
	areas[2] = fabsf( tri_area(p3[0] - p[0], p3[1] - p[1], p3[2] - p[2], p1[0] - p[0], p1[1] - p[1], p1[2] - p[2]) );
	float l = areas[0] + areas[1] + areas[2];
	float _l1 = areas[0] / l;
	float _l2 = areas[2] / l;
	float _l3 = areas[1] / l;
	for(int i = 0; i < n; i++)
		v[i] = _l1*v1[i] + _l2*v2[i] + _l3*v3[i];
	return true;
}

// One ordered pair (mesh a, mesh b) of local meshes whose outlines intersect: candidate nodes of a (g indices
// [cand_off, cand_off + n_cand) of the candidate arrays), faces of b in the box ([face_off, +n_faces) of the face
// array, ascending face id) and the 2-D bins over them (gcm_host.cpp: FaceGrid).
struct CollPair {
	int cand_off, n_cand;
	int face_off, n_faces;
	int grid_ok, ax0, ax1, n0, n1;
	int start_off, items_off;      // CSR: bin -> positions in the pair's face list, ascending
	int phase;                     // VirtNode::phase of the virtual nodes of this pair
	double org0, org1, cell, pad;
	int ax2;                       // the axis the grid drops; [lo2, hi2] = extent of the pair's faces along it
	double lo2, hi2;
};

struct CollFace { int v0, v1, v2, face; };     // merged vertex ids, face id inside its zone

#if defined(__CUDACC__)

__device__ __forceinline__ void report_error(int* err, int code, int zone, int node, int stage);   // gcm_kernels.cuh

__device__ __forceinline__ void coll_xyz(const float* __restrict__ u, size_t rs, int node, float c[3])
{
	const float4 q = rec_quad(u, rs, node, 2);
	c[0] = q.y; c[1] = q.z; c[2] = q.w;
}

__global__ void __launch_bounds__(256)
collide_clear_kernel(int* __restrict__ flags, const int* __restrict__ border_by_slot, int* __restrict__ axis_plus0, int n_border)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= n_border) return;
	axis_plus0[s] = -1;
	flags[border_by_slot[s]] &= ~1;        // setContactType(Free); only border nodes ever carry the bit
}

// positions [e0, e1) of a bin's list: try the faces in order, stop at the first pass or at the best so far
__device__ __forceinline__ int coll_try_bin(const int* __restrict__ items, int e0, int e1, const CollFace* __restrict__ faces,
                                            const float* __restrict__ u, size_t rs, const float p0[3], const float dir[3],
                                            int& best, float best_p[3])
{
	for(int e = e0; e < e1; e++) {
		const int l = items ? items[e] : e;
		if(l >= best) break;
		const CollFace f = faces[l];
		float a[3], b[3], c[3], p[3];
		coll_xyz(u, rs, f.v0, a); coll_xyz(u, rs, f.v1, b); coll_xyz(u, rs, f.v2, c);
		const int r = vector_intersects_triangle(a, b, c, p0, dir, p);
		if(r < 0) return r;
		if(r) { best = l; best_p[0] = p[0]; best_p[1] = p[1]; best_p[2] = p[2]; break; }
	}
	return 0;
}

__global__ void __launch_bounds__(128)
collide_find_kernel(const float* __restrict__ u, size_t rs, const float* __restrict__ basis,
                    const CollPair* __restrict__ pairs, const int* __restrict__ cand_pair, const int* __restrict__ cand_node,
                    const int* __restrict__ cand_slot, int n_cand, const CollFace* __restrict__ all_faces,
                    const int* __restrict__ grid_start, const int* __restrict__ grid_items,
                    int* __restrict__ hit_l, float* __restrict__ hit_p, int* err)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= n_cand) return;
	const CollPair pr = pairs[cand_pair[g]];
	const int node = cand_node[g];
	float p0[3], dir[3];
	coll_xyz(u, rs, node, p0);
	const float* bs = basis + 9 * (size_t)cand_slot[g];        // local_basis->ksi[0]
	dir[0] = bs[0]; dir[1] = bs[1]; dir[2] = bs[2];
	const CollFace* faces = all_faces + pr.face_off;
	int best = 0x7fffffff;
	float bp[3] = {0.f, 0.f, 0.f};
	int rc = 0;
	bool finite = true;
	for(int a = 0; a < 3; a++) finite = finite && !bad_float(p0[a]) && !bad_float(dir[a]);
	if(!pr.grid_ok || !finite) {
		rc = coll_try_bin(nullptr, 0, pr.n_faces, faces, u, rs, p0, dir, best, bp);
	} else {
		// FaceGrid::candidates (gcm_host.cpp): the bins the projected line crosses, in double as on the host
		const int* start = grid_start + pr.start_off;
		const int* items = grid_items + pr.items_off;
		const int n[2] = {pr.n0, pr.n1};
		const double org[2] = {pr.org0, pr.org1}, cell = pr.cell, pad = pr.pad;
		const double p[2] = {(double)p0[pr.ax0], (double)p0[pr.ax1]}, d[2] = {(double)dir[pr.ax0], (double)dir[pr.ax1]};
		const int M = fabs(d[0]) >= fabs(d[1]) ? 0 : 1;
		const int m = 1 - M;
		#define GCM_CLAMPI(x, dim) max(0, min(n[dim] - 1, (int)floor(((x) - org[dim]) / cell)))
		if(d[M] == 0) {
			if(!(p[M] + pad < org[M] || p[M] - pad > org[M] + n[M] * cell || p[m] + pad < org[m] || p[m] - pad > org[m] + n[m] * cell)) {
				const int i0 = GCM_CLAMPI(p[M] - pad, M), i1 = GCM_CLAMPI(p[M] + pad, M);
				const int j0 = GCM_CLAMPI(p[m] - pad, m), j1 = GCM_CLAMPI(p[m] + pad, m);
				for(int i = i0; i <= i1 && rc == 0; i++)
					for(int j = j0; j <= j1 && rc == 0; j++) {
						const size_t b = M == 1 ? (size_t)j * n[1] + i : (size_t)i * n[1] + j;
						rc = coll_try_bin(items, start[b], start[b + 1], faces, u, rs, p0, dir, best, bp);
					}
			}
		} else {
			// A face can only pass with its meeting point inside the faces' extent along the dropped axis: that bounds
			// the parameter of the line, hence the columns it can reach (a near-normal direction projects to a
			// direction of pure rounding noise, whose unbounded line would cross the whole grid)
			int ia = 0, ib = n[M] - 1;
			{
				const double pz = (double)p0[pr.ax2], dz = (double)dir[pr.ax2];
				bool none = false;
				if(dz != 0) {
					const double ta = (pr.lo2 - pad - pz) / dz, tb = (pr.hi2 + pad - pz) / dz;
					const double xa = p[M] + fmin(ta, tb) * d[M], xb = p[M] + fmax(ta, tb) * d[M];
					const double xlo = fmin(xa, xb) - pad, xhi = fmax(xa, xb) + pad;
					if(xlo == xlo && xhi == xhi) {       // finite enough to use
						if(xhi < org[M] || xlo > org[M] + n[M] * cell) none = true;
						else { ia = GCM_CLAMPI(xlo, M); ib = GCM_CLAMPI(xhi, M); }
					}
				} else if(pz < pr.lo2 - pad || pz > pr.hi2 + pad) none = true;
				if(none) { ia = 1; ib = 0; }
			}
			const double slope = d[m] / d[M];
			for(int i = ia; i <= ib && rc == 0; i++) {
				const double x0 = org[M] + i * cell - pad, x1 = org[M] + (i + 1) * cell + pad;
				const double y0 = p[m] + slope * (x0 - p[M]), y1 = p[m] + slope * (x1 - p[M]);
				const double ylo = fmin(y0, y1) - pad, yhi = fmax(y0, y1) + pad;
				if(yhi < org[m] || ylo > org[m] + n[m] * cell) continue;
				const int j0 = GCM_CLAMPI(ylo, m), j1 = GCM_CLAMPI(yhi, m);
				for(int j = j0; j <= j1 && rc == 0; j++) {
					const size_t b = M == 1 ? (size_t)j * n[1] + i : (size_t)i * n[1] + j;
					rc = coll_try_bin(items, start[b], start[b + 1], faces, u, rs, p0, dir, best, bp);
				}
			}
		}
		#undef GCM_CLAMPI
	}
	if(rc < 0) { report_error(err, ERR_COLLISION_NAN, -1, node, -rc); hit_l[g] = -1; return; }
	hit_l[g] = best == 0x7fffffff ? -1 : best;
	hit_p[3 * (size_t)g] = bp[0]; hit_p[3 * (size_t)g + 1] = bp[1]; hit_p[3 * (size_t)g + 2] = bp[2];
}

// virt_nodes.size() when each hit is pushed: exclusive scan of the hit flags in candidate order.  One block.
__global__ void __launch_bounds__(1024)
collide_number_kernel(const int* __restrict__ hit_l, int n_cand, int* __restrict__ hit_virt, int* __restrict__ n_virt)
{
	__shared__ int warp_sum[32];
	__shared__ int carry;
	if(threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for(int base = 0; base < n_cand; base += 1024) {
		const int g = base + threadIdx.x;
		const int flag = (g < n_cand && hit_l[g] >= 0) ? 1 : 0;
		int x = flag;
		#pragma unroll
		for(int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if(lane >= o) x += y; }
		if(lane == 31) warp_sum[w] = x;
		__syncthreads();
		if(w == 0) {
			int s = warp_sum[lane];
			#pragma unroll
			for(int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if(lane >= o) s += y; }
			warp_sum[lane] = s;
		}
		__syncthreads();
		const int before = carry + (w ? warp_sum[w - 1] : 0) + x - flag;
		if(g < n_cand) hit_virt[g] = flag ? before : -1;
		__syncthreads();
		if(threadIdx.x == 1023) carry = before + flag;
		__syncthreads();
	}
	if(threadIdx.x == 0) *n_virt = carry;
}

__global__ void __launch_bounds__(128)
collide_build_kernel(const float* __restrict__ u, size_t rs, const float* __restrict__ mat, size_t stride,
                     const CollPair* __restrict__ pairs, const int* __restrict__ cand_pair, const int* __restrict__ cand_node,
                     const int* __restrict__ cand_slot, int n_cand, const CollFace* __restrict__ all_faces,
                     const int* __restrict__ hit_l, const float* __restrict__ hit_p, const int* __restrict__ hit_virt,
                     VirtNode* __restrict__ virt, int* __restrict__ axis_plus0, int* __restrict__ flags, int* err)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= n_cand) return;
	const int idx = hit_virt[g];
	if(idx < 0) return;
	const CollPair pr = pairs[cand_pair[g]];
	const CollFace f = all_faces[pr.face_off + hit_l[g]];
	const int vs[3] = {f.v0, f.v1, f.v2};
	float c[3][3], vals[3][13];
	for(int j = 0; j < 3; j++) {
		float r[12];
		rec_load12(u, rs, vs[j], r);
		for(int k = 0; k < 9; k++) vals[j][k] = r[k];
		c[j][0] = r[9]; c[j][1] = r[10]; c[j][2] = r[11];
		for(int k = 0; k < 4; k++) vals[j][9 + k] = mat[(size_t)k * stride + vs[j]];
	}
	VirtNode v;
	v.pos[0] = hit_p[3 * (size_t)g]; v.pos[1] = hit_p[3 * (size_t)g + 1]; v.pos[2] = hit_p[3 * (size_t)g + 2];
	float out[13];
	if(!interpolate_triangle(c[0], c[1], c[2], v.pos, vals[0], vals[1], vals[2], out, 13)) {
		report_error(err, ERR_COLLISION_NAN, -1, cand_node[g], 7);
		return;
	}
	for(int k = 0; k < 9; k++) v.val[k] = out[k];
	v.la = out[9]; v.mu = out[10]; v.rho = out[11]; v.yield = out[12];
	v.base = f.v0;
	v.real = cand_node[g];
	v.phase = pr.phase;
	virt[idx] = v;
	atomicMax(&axis_plus0[cand_slot[g]], idx);        // a later pair overwrites an earlier one (:58-60)
	atomicOr(&flags[cand_node[g]], 1);                // setContactType(InContact)
}

#endif // __CUDACC__

} // namespace gcm
