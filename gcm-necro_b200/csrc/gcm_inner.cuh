// gcm_inner.cuh -- the hot kernel: one axis stage for INNER nodes (93-97 % of a mesh).
//
// Same arithmetic as the generic path of gcm_step.cuh (which restates
// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:190-258 for any node), specialised for what is
// invariant at inner nodes in the reference:
//   * the local basis is the identity (create_random_axis overwrites the rotation with E,
//     TetrMethod_Plastic_1stOrder.cpp:712-722), so the displacement of every characteristic foot
//     has one non-zero component (the stage axis) and Omega / Omega^-1 have the fixed sparsity
//     patterns below -- their values depend on (la, mu, rho) only and come from a per-material
//     table built with the general routine elastic_matrix() (bit-identical by construction);
//   * never more than one alpha try (all nine characteristics are inner), never a border solve;
//   * the zero-speed characteristics interpolate at the node itself, inside elements[0]: that
//     weight is static geometry (w0, computed once with the general routines).
// Owner search = the reference's first-passing-candidate rule, but each candidate first goes
// through a conservative barycentric pre-filter (plain float, FMA allowed) that can only answer
// "certainly fails" / "certainly passes" / "don't know"; "don't know" (e.g. feet lying exactly
// on mesh edges, the norm on 6-tet-per-hex meshes) falls to the reference's own volume-sum
// predicate, evaluated division-free when the answer is clear of the 1.001 threshold by > 1e-5
// and literally (point_in_tetr) otherwise.  Owners and values stay bit-identical to the
// reference; tests/test_gpu_parity.py checks that on every case.
#pragma once
#include "gcm_step.cuh"

namespace gcm {

#define GCM_MAX_MATERIALS 8
#if defined(__CUDACC__)
#define GCM_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define GCM_CONSTEXPR_HD constexpr
#endif

// Identity-basis eigensystem of one material for one stage (from elastic_matrix()).
struct MatTab { float L[9]; float U[81]; float U1[81]; };

// Structural non-zeros of Omega (U) and Omega^-1 (U1) rows for q = e_stage (bit j = column j).
// Derived by evaluating elastic_matrix() for several materials; re-verified against the actual
// table entries at upload time (gcm_capi.cu: build_material_table) -- a material whose table
// has a non-zero outside the pattern is routed to the generic kernel.
GCM_CONSTEXPR_HD unsigned short u_pattern(int stage, int row)
{
	constexpr unsigned short P[3][9] = {
		{0x009, 0x009, 0x024, 0x024, 0x012, 0x012, 0x080, 0x140, 0x148},
		{0x042, 0x042, 0x084, 0x084, 0x011, 0x011, 0x020, 0x108, 0x148},
		{0x104, 0x104, 0x082, 0x082, 0x021, 0x021, 0x010, 0x048, 0x148}};
	return P[stage][row];
}
GCM_CONSTEXPR_HD unsigned short u1_pattern(int stage, int row)
{
	constexpr unsigned short P[3][9] = {
		{0x003, 0x030, 0x00c, 0x003, 0x030, 0x00c, 0x183, 0x040, 0x183},
		{0x030, 0x003, 0x00c, 0x183, 0x030, 0x040, 0x003, 0x00c, 0x183},
		{0x030, 0x00c, 0x003, 0x183, 0x040, 0x030, 0x183, 0x00c, 0x003}};
	return P[stage][row];
}

#if defined(__CUDACC__)

__device__ __forceinline__ float comp(const Vec3& v, int s) { return s == 0 ? v.x : (s == 1 ? v.y : v.z); }

// cross(a, b)[s]
__device__ __forceinline__ float cross_comp(const Vec3& a, const Vec3& b, int s)
{
	if(s == 0) return a.y * b.z - a.z * b.y;
	if(s == 1) return a.z * b.x - a.x * b.z;
	return a.x * b.y - a.y * b.x;
}

// Owner search for a foot displaced along the stage axis by `dks` from inner node `node`.
// Returns the owner tet (filling g) or -1; *expand as in find_owner_first_ring.
template<int S>
__device__ __forceinline__ int find_owner_axis(const MeshView& m, int node, const Vec3& X, float dks,
                                               int e0, int e1, TetGeom& g, bool* expand)
{
	// displacement (dks * ksi[S][j]): +-0 off-axis, so R2 == dks*dks and delta == sqrtf(R2)
	const float R2 = dks * dks;
	const float delta = sqrtf(R2);
	const float sgn = dks < 0 ? -1.0f : 1.0f;
	bool inside_R = false;
	for(int e = e0; e < e1; e++) {
		const int t = __ldg(m.n2t + e);
		const int4 q = __ldg(reinterpret_cast<const int4*>(m.tets) + t);
		const float2 hv = __ldg(reinterpret_cast<const float2*>(m.tet_hv) + t);
		g.v[0] = q.x; g.v[1] = q.y; g.v[2] = q.z; g.v[3] = q.w;
		g.h = hv.x; g.vol = hv.y;
		#pragma unroll
		for(int k = 0; k < 4; k++) g.p[k] = (g.v[k] == node) ? X : load_xyz(m, g.v[k]);

		const bool rescale = (delta > 0) && ((double)delta < (double)g.h * 0.99);
		int verdict = 0;   // 0 unknown, 1 certainly in, -1 certainly out
		if(rescale && hv.y > 0) {   // hv.y > 0: well-shaped tet (sign bit of the stored |Vol| is the flag)
			// barycentric coordinates of P = X + 0.99 h e_S sgn w.r.t. the three vertices other than X
			Vec3 ea, eb, ec;
			const int k = (g.v[0] == node) ? 0 : (g.v[1] == node) ? 1 : (g.v[2] == node) ? 2 : 3;
			const Vec3& A = g.p[k == 0 ? 1 : 0];
			const Vec3& B = g.p[k <= 1 ? 2 : 1];
			const Vec3& C = g.p[k == 3 ? 2 : 3];
			ea.x = A.x - X.x; ea.y = A.y - X.y; ea.z = A.z - X.z;
			eb.x = B.x - X.x; eb.y = B.y - X.y; eb.z = B.z - X.z;
			ec.x = C.x - X.x; ec.y = C.y - X.y; ec.z = C.z - X.z;
			const float cbc_x = eb.y * ec.z - eb.z * ec.y;
			const float cbc_y = eb.z * ec.x - eb.x * ec.z;
			const float cbc_z = eb.x * ec.y - eb.y * ec.x;
			const float D = ea.x * cbc_x + ea.y * cbc_y + ea.z * cbc_z;
			const float mm = sgn * 0.99f * g.h;
			const float na = mm * (S == 0 ? cbc_x : (S == 1 ? cbc_y : cbc_z));
			const float nb = mm * cross_comp(ec, ea, S);
			const float nc = mm * cross_comp(ea, eb, S);
			// lambda_j = n_j / D; compare n_j * D against +-1e-3 * D^2
			const float D2 = D * D;
			const float la_ = na * D, lb_ = nb * D, lc_ = nc * D;
			const float lo = -1e-3f * D2, hi = 1e-3f * D2;
			if(la_ < lo || lb_ < lo || lc_ < lo) verdict = -1;
			else if(la_ > hi && lb_ > hi && lc_ > hi && (la_ + lb_ + lc_) < 0.999f * D2) verdict = 1;
		}
		if(verdict == 0) {
			// the reference's predicate (TetrMesh_1stOrder.cpp:1295-1402); only coordinate S of P moves
			float d_s = dks;
			if(rescale) d_s = (float)( (double)(dks * g.h) * 0.99 / (double)delta );
			Vec3 P = X;
			if(S == 0) P.x = X.x + d_s; else if(S == 1) P.y = X.y + d_s; else P.z = X.z + d_s;
			const float a0x = g.p[0].x - P.x, a0y = g.p[0].y - P.y, a0z = g.p[0].z - P.z;
			const float a1x = g.p[1].x - P.x, a1y = g.p[1].y - P.y, a1z = g.p[1].z - P.z;
			const float a2x = g.p[2].x - P.x, a2y = g.p[2].y - P.y, a2z = g.p[2].z - P.z;
			const float a3x = g.p[3].x - P.x, a3y = g.p[3].y - P.y, a3z = g.p[3].z - P.z;
			const float d0 = determinant(a1x, a1y, a1z, a2x, a2y, a2z, a3x, a3y, a3z);
			const float d1 = determinant(a0x, a0y, a0z, a2x, a2y, a2z, a3x, a3y, a3z);
			const float d2 = determinant(a1x, a1y, a1z, a0x, a0y, a0z, a3x, a3y, a3z);
			const float d3 = determinant(a1x, a1y, a1z, a2x, a2y, a2z, a0x, a0y, a0z);
			const float S6 = fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
			const float av = fabsf(g.vol);
			if(S6 < av * 6.0059f) verdict = 1;
			else if(S6 > av * 6.0061f) verdict = -1;
			else {
				const float dxyz[3] = {S == 0 ? dks : dks * 0.0f, S == 1 ? dks : dks * 0.0f, S == 2 ? dks : dks * 0.0f};
				verdict = point_in_tetr(X, dxyz[0], dxyz[1], dxyz[2], delta, g.p[0], g.p[1], g.p[2], g.p[3], g.h, g.vol) ? 1 : -1;
			}
		}
		if(verdict > 0) return t;
		if(!inside_R) {
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				if(g.v[j] == node) break;
				const float lx = X.x - g.p[j].x, ly = X.y - g.p[j].y, lz = X.z - g.p[j].z;
				if(lx * lx + ly * ly + lz * lz < R2) inside_R = true;
			}
		}
	}
	*expand = inside_R;
	return -1;
}

// One thread per listed inner node, axis stage S.  mat_id == nullptr: single material (id 0).
template<int S>
__global__ void __launch_bounds__(128)
stage_inner_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list,
                   float* __restrict__ un, int* err, int zone, StageTap* tap,
                   const MatTab* __restrict__ tab, int n_mat, const unsigned char* __restrict__ mat_id,
                   const float* __restrict__ w0)
{
	__shared__ float sU[GCM_MAX_MATERIALS][81];
	__shared__ float sU1[GCM_MAX_MATERIALS][81];
	__shared__ float sL[GCM_MAX_MATERIALS][4];
	for(int i = threadIdx.x; i < n_mat * 81; i += blockDim.x) {
		const MatTab& mt = tab[(i / 81) * 3 + S];
		sU[i / 81][i % 81] = mt.U[i % 81];
		sU1[i / 81][i % 81] = mt.U1[i % 81];
	}
	for(int i = threadIdx.x; i < n_mat * 4; i += blockDim.x) sL[i / 4][i % 4] = tab[(i / 4) * 3 + S].L[i % 4];
	__syncthreads();

	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const int mid = mat_id ? mat_id[node] : 0;
	const Vec3 X = load_xyz(m, node);
	float own[9];
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) {
		own[k] = m.u[(size_t)k * m.stride + node];
		bad |= (isnan(own[k]) || isinf(own[k]));
	}
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }

	const int e0 = m.n2t_off[node], e1 = m.n2t_off[node + 1];
	float val[4][9];
	StageTap t;
	#pragma unroll
	for(int p = 0; p < 4; p++) {
		const float dks = - sL[mid][p] * m.tau;      // dksi[i] = - L(i,i) * time_step
		TetGeom g;
		bool expand = false;
		const int owner = find_owner_axis<S>(m, node, X, dks, e0, e1, g, &expand);
		t.owner[p] = owner;
		if(owner < 0) {
			report_error(err, expand ? ERR_RING_OVERFLOW : ERR_INNER_NO_OWNER, zone, node, S);
			return;
		}
		// foot coordinates = stored coordinates + dx (dx = +-0 off-axis)
		Vec3 F = X;
		if(S == 0) F.x = X.x + dks; else if(S == 1) F.y = X.y + dks; else F.z = X.z + dks;
		float f[4];
		if(!interp_factors(g.p[0], g.p[1], g.p[2], g.p[3], F.x, F.y, F.z, g.vol, f)) {
			report_error(err, ERR_INTERP_SUM, zone, node, S);
			return;
		}
		#pragma unroll
		for(int k = 0; k < 9; k++) {
			const float* pl = m.u + (size_t)k * m.stride;
			val[p][k] = interp4(pl[g.v[0]], pl[g.v[1]], pl[g.v[2]], pl[g.v[3]], f);
		}
	}
	// zero-speed characteristics: interpolation at the node itself inside elements[0]; every
	// factor but the node's own is exactly 0, the own one is the static weight w0
	float vz[9];
	const float wz = w0[node];
	#pragma unroll
	for(int k = 0; k < 9; k++) vz[k] = own[k] * wz;
	if(tap) {
		t.owner[4] = m.n2t[e0];
		t.outer_count = 0; t.tries = 1;
		tap[(size_t)S * m.n_nodes + node] = t;
	}

	// omega = Omega * u(foot), u' = Omega^-1 * omega (SimpleVolumeCalculator.cpp:7-32), skipping
	// the structural zeros; ppoint_num = (0,1,2,3,2,3,4,4,4)
	const float* U = sU[mid];
	const float* U1 = sU1[mid];
	float omega[9];
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		const float* v = (i < 4) ? val[i] : (i < 6 ? val[i - 2] : vz);
		float om = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u_pattern(S, i) >> j) & 1) om += U[i * 9 + j] * v[j];
		omega[i] = om;
	}
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		float o = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u1_pattern(S, i) >> j) & 1) o += U1[i * 9 + j] * omega[j];
		un[(size_t)i * m.stride + node] = o;
	}
}

#endif // __CUDACC__

} // namespace gcm
