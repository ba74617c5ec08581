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

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 sel3(bool c, const F3& a, const F3& b) { F3 r; r.x = c ? a.x : b.x; r.y = c ? a.y : b.y; r.z = c ? a.z : b.z; return r; }
template<int S> __device__ __forceinline__ float comp(const F3& v) { return S == 0 ? v.x : (S == 1 ? v.y : v.z); }
// cross(a, b)[S]
template<int S> __device__ __forceinline__ float cross_comp(const F3& a, const F3& b)
{
	if(S == 0) return a.y * b.z - a.z * b.y;
	if(S == 1) return a.z * b.x - a.x * b.z;
	return a.x * b.y - a.y * b.x;
}

// One candidate tetrahedron as seen from inner node X: the three other vertices (A, B, C in
// tet order), their third record quad (szz, x, y, z) and the position k of X in the tet.
struct Cand {
	int t; int4 tv; float h, vol; int k; int ia, ib, ic; float4 qa, qb, qc;
};

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
	const float4* __restrict__ rec = reinterpret_cast<const float4*>(m.u);
	const float4 o0 = rec[3 * (size_t)node], o1 = rec[3 * (size_t)node + 1], o2 = rec[3 * (size_t)node + 2];
	const float own[9] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x};
	F3 X; X.x = o2.y; X.y = o2.z; X.z = o2.w;
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) bad |= (isnan(own[k]) || isinf(own[k]));
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }

	const int e0 = m.n2t_off[node], e1 = m.n2t_off[node + 1];
	const int4* __restrict__ tets4 = reinterpret_cast<const int4*>(m.tets);
	const float2* __restrict__ hv2 = reinterpret_cast<const float2*>(m.tet_hv);
	const float* U = sU[mid];
	const float* U1 = sU1[mid];
	float omega[9];
	StageTap tp;
	#pragma unroll 1
	for(int p = 0; p < 4; p++) {
		const float dks = - sL[mid][p] * m.tau;      // dksi[i] = - L(i,i) * time_step
		// displacement dks * ksi[S][j]: +-0 off-axis, so R2 == dks*dks and delta == sqrtf(R2)
		const float R2 = dks * dks;
		const float delta = sqrtf(R2);
		const float sgn = dks < 0 ? -1.0f : 1.0f;
		bool inside_R = false;
		int owner = -1;
		Cand c;
		F3 A, B, C;
		for(int e = e0; e < e1; e++) {
			c.t = __ldg(m.n2t + e);
			c.tv = __ldg(tets4 + c.t);
			const float2 hv = __ldg(hv2 + c.t);
			c.h = hv.x; c.vol = hv.y;
			c.k = (c.tv.x == node) ? 0 : (c.tv.y == node) ? 1 : (c.tv.z == node) ? 2 : 3;
			c.ia = (c.k == 0) ? c.tv.y : c.tv.x;
			c.ib = (c.k <= 1) ? c.tv.z : c.tv.y;
			c.ic = (c.k == 3) ? c.tv.z : c.tv.w;
			c.qa = __ldg(rec + 3 * (size_t)c.ia + 2);
			c.qb = __ldg(rec + 3 * (size_t)c.ib + 2);
			c.qc = __ldg(rec + 3 * (size_t)c.ic + 2);
			A.x = c.qa.y; A.y = c.qa.z; A.z = c.qa.w;
			B.x = c.qb.y; B.y = c.qb.z; B.z = c.qb.w;
			C.x = c.qc.y; C.y = c.qc.z; C.z = c.qc.w;

			const bool rescale = (delta > 0) && ((double)delta < (double)c.h * 0.99);
			int verdict = 0;   // 0 unknown, 1 certainly in, -1 certainly out
			if(rescale && c.vol > 0) {   // vol > 0: well-shaped tet (sign bit of the stored |Vol| is the flag)
				// barycentric coordinates of P = X + 0.99 h sgn e_S w.r.t. A, B, C (Cramer)
				F3 ea, eb, ec;
				ea.x = A.x - X.x; ea.y = A.y - X.y; ea.z = A.z - X.z;
				eb.x = B.x - X.x; eb.y = B.y - X.y; eb.z = B.z - X.z;
				ec.x = C.x - X.x; ec.y = C.y - X.y; ec.z = C.z - X.z;
				F3 cbc;
				cbc.x = eb.y * ec.z - eb.z * ec.y;
				cbc.y = eb.z * ec.x - eb.x * ec.z;
				cbc.z = eb.x * ec.y - eb.y * ec.x;
				const float D = ea.x * cbc.x + ea.y * cbc.y + ea.z * cbc.z;
				const float mm = sgn * 0.99f * c.h;
				// lambda_j = n_j / D  ->  compare n_j * D with multiples of D^2
				const float la_ = mm * comp<S>(cbc) * D;
				const float lb_ = mm * cross_comp<S>(ec, ea) * D;
				const float lc_ = mm * cross_comp<S>(ea, eb) * D;
				const float D2 = D * D;
				const float lx_ = D2 - (la_ + lb_ + lc_);
				const float lo = -1e-3f * D2, ok = -1e-4f * D2;
				// any coordinate below -1e-3: sum|s| > 1.0019 Vol, fails; all above -1e-4: sum|s| < 1.0007 Vol, passes
				if(la_ < lo || lb_ < lo || lc_ < lo || lx_ < lo) verdict = -1;
				else if(la_ > ok && lb_ > ok && lc_ > ok && lx_ > ok) verdict = 1;
			}
			if(verdict == 0) {
				// the reference's predicate (TetrMesh_1stOrder.cpp:1295-1402); only coordinate S of P moves
				float d_s = dks;
				if(rescale) d_s = (float)( (double)(dks * c.h) * 0.99 / (double)delta );
				F3 P = X;
				if(S == 0) P.x = X.x + d_s; else if(S == 1) P.y = X.y + d_s; else P.z = X.z + d_s;
				const F3 p0 = sel3(c.k == 0, X, A);
				const F3 p1 = sel3(c.k == 0, A, sel3(c.k == 1, X, B));
				const F3 p2 = sel3(c.k <= 1, B, sel3(c.k == 2, X, C));
				const F3 p3 = sel3(c.k == 3, X, C);
				const float a0x = p0.x - P.x, a0y = p0.y - P.y, a0z = p0.z - P.z;
				const float a1x = p1.x - P.x, a1y = p1.y - P.y, a1z = p1.z - P.z;
				const float a2x = p2.x - P.x, a2y = p2.y - P.y, a2z = p2.z - P.z;
				const float a3x = p3.x - P.x, a3y = p3.y - P.y, a3z = p3.z - P.z;
				const float d0 = determinant(a1x, a1y, a1z, a2x, a2y, a2z, a3x, a3y, a3z);
				const float d1 = determinant(a0x, a0y, a0z, a2x, a2y, a2z, a3x, a3y, a3z);
				const float d2 = determinant(a1x, a1y, a1z, a0x, a0y, a0z, a3x, a3y, a3z);
				const float d3 = determinant(a1x, a1y, a1z, a2x, a2y, a2z, a0x, a0y, a0z);
				const float S6 = fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
				const float av = fabsf(c.vol);
				if(S6 < av * 6.0059f) verdict = 1;
				else if(S6 > av * 6.0061f) verdict = -1;
				else {
					Vec3 q0, q1, q2, q3, XX;
					q0.x = p0.x; q0.y = p0.y; q0.z = p0.z; q1.x = p1.x; q1.y = p1.y; q1.z = p1.z;
					q2.x = p2.x; q2.y = p2.y; q2.z = p2.z; q3.x = p3.x; q3.y = p3.y; q3.z = p3.z;
					XX.x = X.x; XX.y = X.y; XX.z = X.z;
					const float z0 = dks * 0.0f;
					verdict = point_in_tetr(XX, S == 0 ? dks : z0, S == 1 ? dks : z0, S == 2 ? dks : z0, delta,
					                        q0, q1, q2, q3, c.h, c.vol) ? 1 : -1;
				}
			}
			if(verdict > 0) { owner = c.t; break; }
			if(!inside_R) {
				// vertices in tet order up to (not including) the node itself, :1541-1555
				const float da = (X.x - A.x) * (X.x - A.x) + (X.y - A.y) * (X.y - A.y) + (X.z - A.z) * (X.z - A.z);
				const float db = (X.x - B.x) * (X.x - B.x) + (X.y - B.y) * (X.y - B.y) + (X.z - B.z) * (X.z - B.z);
				const float dc = (X.x - C.x) * (X.x - C.x) + (X.y - C.y) * (X.y - C.y) + (X.z - C.z) * (X.z - C.z);
				if(c.k >= 1 && da < R2) inside_R = true;
				if(c.k >= 2 && db < R2) inside_R = true;
				if(c.k >= 3 && dc < R2) inside_R = true;
			}
		}
		if(p == 0) tp.owner[0] = owner;
		if(p == 1) tp.owner[1] = owner;
		if(p == 2) tp.owner[2] = owner;
		if(p == 3) tp.owner[3] = owner;
		if(owner < 0) {
			report_error(err, inside_R ? ERR_RING_OVERFLOW : ERR_INNER_NO_OWNER, zone, node, S);
			return;
		}
		// interpolation at the foot = stored coordinates + dx (TetrMesh_1stOrder.cpp:1764-1880)
		F3 Fp = X;
		if(S == 0) Fp.x = X.x + dks; else if(S == 1) Fp.y = X.y + dks; else Fp.z = X.z + dks;
		const F3 p0 = sel3(c.k == 0, X, A);
		const F3 p1 = sel3(c.k == 0, A, sel3(c.k == 1, X, B));
		const F3 p2 = sel3(c.k <= 1, B, sel3(c.k == 2, X, C));
		const F3 p3 = sel3(c.k == 3, X, C);
		Vec3 q0, q1, q2, q3;
		q0.x = p0.x; q0.y = p0.y; q0.z = p0.z; q1.x = p1.x; q1.y = p1.y; q1.z = p1.z;
		q2.x = p2.x; q2.y = p2.y; q2.z = p2.z; q3.x = p3.x; q3.y = p3.y; q3.z = p3.z;
		float f[4];
		if(!interp_factors(q0, q1, q2, q3, Fp.x, Fp.y, Fp.z, c.vol, f)) {
			report_error(err, ERR_INTERP_SUM, zone, node, S);
			return;
		}
		// factors of A, B, C and of the node itself
		const float fX = c.k == 0 ? f[0] : (c.k == 1 ? f[1] : (c.k == 2 ? f[2] : f[3]));
		const float fA = c.k == 0 ? f[1] : f[0];
		const float fB = c.k <= 1 ? f[2] : f[1];
		const float fC = c.k == 3 ? f[2] : f[3];
		const float4 a0 = __ldg(rec + 3 * (size_t)c.ia), a1 = __ldg(rec + 3 * (size_t)c.ia + 1);
		const float4 b0 = __ldg(rec + 3 * (size_t)c.ib), b1 = __ldg(rec + 3 * (size_t)c.ib + 1);
		const float4 c0 = __ldg(rec + 3 * (size_t)c.ic), c1 = __ldg(rec + 3 * (size_t)c.ic + 1);
		const float va[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, c.qa.x};
		const float vb[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, c.qb.x};
		const float vc[9] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c.qc.x};
		// ((v0 f0 + v1 f1) + v2 f2) + v3 f3 in tet order; a + b == b + a, so k = 0 and k = 1 coincide
		float v[9];
		#pragma unroll
		for(int q = 0; q < 9; q++) {
			const float tA = va[q] * fA, tB = vb[q] * fB, tC = vc[q] * fC, tX = own[q] * fX;
			const float s1 = tA + (c.k <= 1 ? tX : tB);
			const float s2 = s1 + (c.k <= 1 ? tB : (c.k == 2 ? tX : tC));
			v[q] = s2 + (c.k == 3 ? tX : tC);
		}
		// omega_i = Omega(i,:) . u(foot of i)  (SimpleVolumeCalculator.cpp:14-23): foot p feeds row p
		// and, for p = 2, 3, row p + 2 (ppoint_num = 0,1,2,3,2,3,4,4,4).  Rows 0-3 are summed over the
		// union of the row-0/1 and row-2/3 patterns: the extra entries are exact zeros of Omega, as in
		// the reference's dense loop.
		float oa = 0, ob = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++) {
			if(((u_pattern(S, 0) | u_pattern(S, 2)) >> j) & 1) oa += U[p * 9 + j] * v[j];
			if((u_pattern(S, 4) >> j) & 1) ob += U[(p + 2) * 9 + j] * v[j];
		}
		if(p == 0) omega[0] = oa;
		if(p == 1) omega[1] = oa;
		if(p == 2) { omega[2] = oa; omega[4] = ob; }
		if(p == 3) { omega[3] = oa; omega[5] = ob; }
	}
	// zero-speed characteristics: interpolation at the node itself inside elements[0]; every
	// factor but the node's own is exactly 0, the own one is the static weight w0
	float vz[9];
	const float wz = w0[node];
	#pragma unroll
	for(int k = 0; k < 9; k++) vz[k] = own[k] * wz;
	if(tap) {
		tp.owner[4] = m.n2t[e0];
		tp.outer_count = 0; tp.tries = 1;
		tap[(size_t)S * m.n_nodes + node] = tp;
	}

	// rows 6-8 read the zero-speed foot; then u' = Omega^-1 * omega (SimpleVolumeCalculator.cpp:24-31),
	// skipping the structural zeros
	#pragma unroll
	for(int i = 6; i < 9; i++) {
		float om = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u_pattern(S, i) >> j) & 1) om += U[i * 9 + j] * vz[j];
		omega[i] = om;
	}
	float res[9];
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		float o = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u1_pattern(S, i) >> j) & 1) o += U1[i * 9 + j] * omega[j];
		res[i] = o;
	}
	float4* __restrict__ out = reinterpret_cast<float4*>(un) + 3 * (size_t)node;
	out[0] = make_float4(res[0], res[1], res[2], res[3]);
	out[1] = make_float4(res[4], res[5], res[6], res[7]);
	out[2] = make_float4(res[8], X.x, X.y, X.z);
}

#endif // __CUDACC__

} // namespace gcm
