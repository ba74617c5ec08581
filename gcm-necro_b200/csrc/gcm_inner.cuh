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
//     weight is static geometry (w0, computed once with the general routines);
//   * point_in_tetr rescales every short displacement to 0.99 h(tet) (TetrMesh_1stOrder.cpp:
//     1304-1307), so the tested point depends on the DIRECTION of the foot only (up to rounding):
//     the -c1 and -c2 feet share one candidate scan, and so do the +c1 and +c2 feet.
// Owner search = the reference's first-passing-candidate rule, but each candidate first goes
// through a conservative barycentric pre-filter (plain float) that can only answer "certainly
// fails" / "certainly passes" / "don't know"; "don't know" falls to the reference's own
// volume-sum predicate, evaluated division-free when the answer is clear of the 1.001 threshold
// by > 1e-5 and literally (point_in_tetr) otherwise.  Owners and values stay bit-identical to the
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
// table entries at preprocessing time (gcm_host.cpp: build_fast_path_tables) -- a material whose
// table has a non-zero outside the pattern is routed to the generic kernel.
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
__device__ __forceinline__ Vec3 to_vec3(const F3& a) { Vec3 v; v.x = a.x; v.y = a.y; v.z = a.z; return v; }
template<int S> __device__ __forceinline__ float comp(const F3& v) { return S == 0 ? v.x : (S == 1 ? v.y : v.z); }
// cross(a, b)[S]
template<int S> __device__ __forceinline__ float cross_comp(const F3& a, const F3& b)
{
	if(S == 0) return a.y * b.z - a.z * b.y;
	if(S == 1) return a.z * b.x - a.x * b.z;
	return a.x * b.y - a.y * b.x;
}

// One candidate tetrahedron as seen from inner node X.  The node->tet CSR is expanded for this
// kernel into n2x[e] = (ia, ib, ic, t | k << 30): the three OTHER vertices in tet order and the
// position k of the node in the tet, so that the vertex gathers do not wait for a tets[t] lookup.
struct Cand {
	int t, k, ia, ib, ic;
	float h, vol;
	float4 qa, qb, qc;      // third record quad of A, B, C: (szz, x, y, z)
};

__device__ __forceinline__ void tet_order(const Cand& c, const F3& X, F3& p0, F3& p1, F3& p2, F3& p3)
{
	F3 A, B, C;
	A.x = c.qa.y; A.y = c.qa.z; A.z = c.qa.w;
	B.x = c.qb.y; B.y = c.qb.z; B.z = c.qb.w;
	C.x = c.qc.y; C.y = c.qc.z; C.z = c.qc.w;
	p0 = sel3(c.k == 0, X, A);
	p1 = sel3(c.k == 0, A, sel3(c.k == 1, X, B));
	p2 = sel3(c.k <= 1, B, sel3(c.k == 2, X, C));
	p3 = sel3(c.k == 3, X, C);
}

// The reference's predicate (TetrMesh_1stOrder.cpp:1295-1402) for a displacement dks along axis
// S; only coordinate S of the tested point moves.  Division-free unless within 1e-5 of 1.001.
// Arguments by value (registers): a by-reference Cand would pin the candidate to the stack of the
// hot loop for the sake of this rarely taken call.
template<int S>
__device__ __noinline__ bool exact_in_tetr(int k, float h, float vol, float ax, float ay, float az, float bx, float by, float bz,
                                           float cx, float cy, float cz, float Xx, float Xy, float Xz, float dks, float delta)
{
	Cand c;
	c.k = k; c.h = h; c.vol = vol;
	c.qa = make_float4(0.f, ax, ay, az); c.qb = make_float4(0.f, bx, by, bz); c.qc = make_float4(0.f, cx, cy, cz);
	F3 X; X.x = Xx; X.y = Xy; X.z = Xz;
	const bool rescale = (delta > 0) && ((double)delta < (double)c.h * 0.99);
	float d_s = dks;
	if(rescale) d_s = (float)( (double)(dks * c.h) * 0.99 / (double)delta );
	F3 P = X;
	if(S == 0) P.x = X.x + d_s; else if(S == 1) P.y = X.y + d_s; else P.z = X.z + d_s;
	F3 p0, p1, p2, p3;
	tet_order(c, X, p0, p1, p2, p3);
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
	if(S6 < av * 6.0059f) return true;
	if(S6 > av * 6.0061f) return false;
	const float z0 = dks * 0.0f;
	return point_in_tetr(to_vec3(X), S == 0 ? dks : z0, S == 1 ? dks : z0, S == 2 ? dks : z0, delta,
	                     to_vec3(p0), to_vec3(p1), to_vec3(p2), to_vec3(p3), c.h, c.vol);
}

// Pre-filter: barycentric coordinates (Cramer) of P = X + 0.99 h sgn e_S w.r.t. A, B, C and X.
// Returns -1 when some coordinate is below -1e-3 (then sum|s| > 1.0019 Vol: the predicate
// fails), +1 when all are above -1e-4 (sum|s| < 1.0007 Vol: it passes), 0 otherwise.  Only for
// well-shaped tets (stored |Vol| > 0) whose displacement is rescaled (delta < 0.99 h).
template<int S>
__device__ __forceinline__ int prefilter(const Cand& c, const F3& X, float sgn)
{
	F3 ea, eb, ec;
	ea.x = c.qa.y - X.x; ea.y = c.qa.z - X.y; ea.z = c.qa.w - X.z;
	eb.x = c.qb.y - X.x; eb.y = c.qb.z - X.y; eb.z = c.qb.w - X.z;
	ec.x = c.qc.y - X.x; ec.y = c.qc.z - X.y; ec.z = c.qc.w - X.z;
	F3 cbc;
	cbc.x = eb.y * ec.z - eb.z * ec.y;
	cbc.y = eb.z * ec.x - eb.x * ec.z;
	cbc.z = eb.x * ec.y - eb.y * ec.x;
	const float D = ea.x * cbc.x + ea.y * cbc.y + ea.z * cbc.z;
	const float mD = sgn * 0.99f * c.h * D;
	// lambda_j = n_j / D  ->  compare n_j * D with multiples of D^2
	const float la_ = mD * comp<S>(cbc);
	const float lb_ = mD * cross_comp<S>(ec, ea);
	const float lc_ = mD * cross_comp<S>(ea, eb);
	const float D2 = D * D;
	const float lx_ = D2 - (la_ + lb_ + lc_);
	const float lo = -1e-3f * D2, ok = -1e-4f * D2;
	const float mn = fminf(fminf(la_, lb_), fminf(lc_, lx_));
	return mn < lo ? -1 : (mn > ok ? 1 : 0);
}

template<int S>
__device__ __forceinline__ void load_cand(Cand& c, const int4 x, const float* __restrict__ rec, size_t rs, const float2* __restrict__ hv2)
{
	c.ia = x.x; c.ib = x.y; c.ic = x.z;
	c.t = x.w & 0x3fffffff; c.k = (unsigned)x.w >> 30;
	c.qa = rec_quad_ldg(rec, rs, c.ia, 2);
	c.qb = rec_quad_ldg(rec, rs, c.ib, 2);
	c.qc = rec_quad_ldg(rec, rs, c.ic, 2);
	const float2 hv = __ldg(hv2 + c.t);
	c.h = hv.x; c.vol = hv.y;
}

// Cold path: a node the specialised scan cannot settle (the two feet of a direction disagree on a
// candidate, or no owner in the first ring) goes through the literal per-node routine, which also
// classifies the failure exactly as the reference does.
__device__ __noinline__ void inner_node_generic(const MeshView& m, int node, int stage, float* __restrict__ un,
                                                int* err, int zone, StageTap* tap, const float* __restrict__ yield_plane)
{
	float out[9];
	StageTap t;
	const int e = node_stage_nocontact(m, node, stage, out, tap ? &t : nullptr);
	if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
	if(e) { report_error(err, e, zone, node, stage); return; }
	if(yield_plane && !plastic_epilogue(out, yield_plane[node])) { report_error(err, ERR_NAN, zone, node, 3); return; }
	const float4 c = rec_quad(m.u, m.rs, node, 2);
	rec_quad_store(un, m.rs, node, 0, make_float4(out[0], out[1], out[2], out[3]));
	rec_quad_store(un, m.rs, node, 1, make_float4(out[4], out[5], out[6], out[7]));
	rec_quad_store(un, m.rs, node, 2, make_float4(out[8], c.y, c.z, c.w));
}

// Interpolated state at the foot X + dks e_S inside candidate c (TetrMesh_1stOrder.cpp:1764-1880).
// va/vb/vc = the 9 values of A, B, C; own = the node's.  Returns false when the reference throws.
template<int S>
__device__ __forceinline__ bool interp_axis(const Cand& c, const F3& X, float dks, const float own[9],
                                            const float va[9], const float vb[9], const float vc[9], float v[9])
{
	F3 Fp = X;
	if(S == 0) Fp.x = X.x + dks; else if(S == 1) Fp.y = X.y + dks; else Fp.z = X.z + dks;
	F3 p0, p1, p2, p3;
	tet_order(c, X, p0, p1, p2, p3);
	float f[4];
	if(!interp_factors(to_vec3(p0), to_vec3(p1), to_vec3(p2), to_vec3(p3), Fp.x, Fp.y, Fp.z, c.vol, f))
		return false;
	const float fX = c.k == 0 ? f[0] : (c.k == 1 ? f[1] : (c.k == 2 ? f[2] : f[3]));
	const float fA = c.k == 0 ? f[1] : f[0];
	const float fB = c.k <= 1 ? f[2] : f[1];
	const float fC = c.k == 3 ? f[2] : f[3];
	// ((v0 f0 + v1 f1) + v2 f2) + v3 f3 in tet order; a + b == b + a, so k = 0 and k = 1 coincide
	#pragma unroll
	for(int q = 0; q < 9; q++) {
		const float tA = va[q] * fA, tB = vb[q] * fB, tC = vc[q] * fC, tX = own[q] * fX;
		const float s1 = tA + (c.k <= 1 ? tX : tB);
		const float s2 = s1 + (c.k <= 1 ? tB : (c.k == 2 ? tX : tC));
		v[q] = s2 + (c.k == 3 ? tX : tC);
	}
	return true;
}

__device__ __forceinline__ void load_values(const Cand& c, const float* __restrict__ rec, size_t rs, float va[9], float vb[9], float vc[9])
{
	const float4 a0 = rec_quad_ldg(rec, rs, c.ia, 0), a1 = rec_quad_ldg(rec, rs, c.ia, 1);
	const float4 b0 = rec_quad_ldg(rec, rs, c.ib, 0), b1 = rec_quad_ldg(rec, rs, c.ib, 1);
	const float4 c0 = rec_quad_ldg(rec, rs, c.ic, 0), c1 = rec_quad_ldg(rec, rs, c.ic, 1);
	va[0] = a0.x; va[1] = a0.y; va[2] = a0.z; va[3] = a0.w; va[4] = a1.x; va[5] = a1.y; va[6] = a1.z; va[7] = a1.w; va[8] = c.qa.x;
	vb[0] = b0.x; vb[1] = b0.y; vb[2] = b0.z; vb[3] = b0.w; vb[4] = b1.x; vb[5] = b1.y; vb[6] = b1.z; vb[7] = b1.w; vb[8] = c.qb.x;
	vc[0] = c0.x; vc[1] = c0.y; vc[2] = c0.z; vc[3] = c0.w; vc[4] = c1.x; vc[5] = c1.y; vc[6] = c1.z; vc[7] = c1.w; vc[8] = c.qc.x;
}

// One thread per listed inner node, axis stage S, scanning the node's whole candidate list.
// This is the index-free form of the inner kernel (kept as the A/B reference of the indexed one
// below: GCM_B200_DIR_INDEX=0).  mat_id == nullptr: single material (id 0).
// MB = minimum resident blocks per SM the register allocation is tuned for (occupancy vs spills).
template<int S, int MB>
__global__ void __launch_bounds__(128, MB)
stage_inner_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list,
                   float* __restrict__ un, int* err, int zone, StageTap* tap,
                   const MatTab* __restrict__ tab, int n_mat, const unsigned char* __restrict__ mat_id,
                   const float* __restrict__ w0, const int4* __restrict__ n2x, const float* __restrict__ yield_plane)
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
	const float* __restrict__ rec = m.u;
	const size_t rs = m.rs;
	const float2* __restrict__ hv2 = reinterpret_cast<const float2*>(m.tet_hv);
	const float4 o0 = rec_quad(rec, rs, node, 0), o1 = rec_quad(rec, rs, node, 1), o2 = rec_quad(rec, rs, node, 2);
	const float own[9] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x};
	F3 X; X.x = o2.y; X.y = o2.z; X.z = o2.w;
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) bad |= (isnan(own[k]) || isinf(own[k]));
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }

	const int e0 = m.n2t_off[node], e1 = m.n2t_off[node + 1];
	const float* U = sU[mid];
	const float* U1 = sU1[mid];
	float omega[9];
	StageTap tp;

	// dir 0: dksi < 0 (characteristics 0 and 2: -c1 tau, -c2 tau); dir 1: dksi > 0 (1 and 3)
	#pragma unroll
	for(int dir = 0; dir < 2; dir++) {
		const float dk1 = - sL[mid][dir] * m.tau;        // dksi[i] = - L(i,i) * time_step
		const float dk2 = - sL[mid][2 + dir] * m.tau;
		const float sgn = dir == 0 ? -1.0f : 1.0f;
		const float de1 = sqrtf(dk1 * dk1), de2 = sqrtf(dk2 * dk2);   // delta; off-axis components are +-0
		Cand c;
		int owner = -1;
		bool cold = false;
		int4 xn = __ldg(n2x + e0);
		for(int e = e0; e < e1; e++) {
			const int4 x = xn;
			if(e + 1 < e1) xn = __ldg(n2x + e + 1);
			load_cand<S>(c, x, rec, rs, hv2);
			const double h99 = (double)c.h * 0.99;
			int v1 = 0;
			if(c.vol > 0 && de1 > 0 && de2 > 0 && (double)de1 < h99 && (double)de2 < h99)
				v1 = prefilter<S>(c, X, sgn);
			if(v1 == 0) {
				const bool in1 = exact_in_tetr<S>(c.k, c.h, c.vol, c.qa.y, c.qa.z, c.qa.w, c.qb.y, c.qb.z, c.qb.w,
				                                  c.qc.y, c.qc.z, c.qc.w, X.x, X.y, X.z, dk1, de1);
				const bool in2 = exact_in_tetr<S>(c.k, c.h, c.vol, c.qa.y, c.qa.z, c.qa.w, c.qb.y, c.qb.z, c.qb.w,
				                                  c.qc.y, c.qc.z, c.qc.w, X.x, X.y, X.z, dk2, de2);
				if(in1 != in2) { cold = true; break; }
				v1 = in1 ? 1 : -1;
			}
			if(v1 > 0) { owner = c.t; break; }
		}
		if(cold || owner < 0) {
			inner_node_generic(m, node, S, un, err, zone, tap, yield_plane);
			return;
		}
		tp.owner[dir] = owner;
		tp.owner[2 + dir] = owner;

		float va[9], vb[9], vc[9], v[9];
		load_values(c, rec, rs, va, vb, vc);
		if(!interp_axis<S>(c, X, dk1, own, va, vb, vc, v)) { report_error(err, ERR_INTERP_SUM, zone, node, S); return; }
		// omega_i = Omega(i,:) . u(foot of i)  (SimpleVolumeCalculator.cpp:14-23); ppoint_num = (0,1,2,3,2,3,4,4,4)
		{
			float om = 0;
			#pragma unroll
			for(int j = 0; j < 9; j++)
				if((u_pattern(S, dir) >> j) & 1) om += U[dir * 9 + j] * v[j];
			omega[dir] = om;
		}
		if(!interp_axis<S>(c, X, dk2, own, va, vb, vc, v)) { report_error(err, ERR_INTERP_SUM, zone, node, S); return; }
		{
			float om = 0, om4 = 0;
			#pragma unroll
			for(int j = 0; j < 9; j++) {
				if((u_pattern(S, 2 + dir) >> j) & 1) om += U[(2 + dir) * 9 + j] * v[j];
				if((u_pattern(S, 4 + dir) >> j) & 1) om4 += U[(4 + dir) * 9 + j] * v[j];
			}
			omega[2 + dir] = om;
			omega[4 + dir] = om4;
		}
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
	#pragma unroll
	for(int i = 6; i < 9; i++) {
		float om = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u_pattern(S, i) >> j) & 1) om += U[i * 9 + j] * vz[j];
		omega[i] = om;
	}
	// u' = Omega^-1 * omega (SimpleVolumeCalculator.cpp:24-31), skipping the structural zeros
	float res[9];
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		float o = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u1_pattern(S, i) >> j) & 1) o += U1[i * 9 + j] * omega[j];
		res[i] = o;
	}
	if(yield_plane && !plastic_epilogue(res, yield_plane[node])) { report_error(err, ERR_NAN, zone, node, 3); return; }
	rec_quad_store(un, rs, node, 0, make_float4(res[0], res[1], res[2], res[3]));
	rec_quad_store(un, rs, node, 1, make_float4(res[4], res[5], res[6], res[7]));
	rec_quad_store(un, rs, node, 2, make_float4(res[8], X.x, X.y, X.z));
}

#endif // __CUDACC__ (index-free kernel)

// ---- step-invariant cache of the inner-node stage ------------------------------------------------
// For an inner node nothing the stage computes before it touches the VALUES changes from step to step:
// the basis is the identity, tau is fixed (TetrMeshSet.cpp:105-106), the mesh does not move
// (move_coords is disabled, TetrMeshSet.cpp:170-171) and the material is static.  So the owner
// tetrahedron of every characteristic foot and its four interpolation factors are functions of static
// data -- the legal cache of SURVEY section 7.  It is built once per (tau, mesh) by the LITERAL per-node
// routine of gcm_step.cuh (find_nodes_on_previous_time_layer: the reference's own search, predicate and
// interpolate arithmetic, no pre-filter), one 96-byte entry per (axis, node):
//   a0 b0 c0 | flags      dksi < 0: the owner's other three vertices in tet order; flags = k0 | k1 << 2 | valid << 31
//   w10[4]                factors (own, A, B, C) of the foot  -c1 tau
//   w20[4]                factors of the foot  -c2 tau
//   a1 b1 c1 | w0         dksi > 0: the same; w0 = own factor of the zero-speed foot (inside elements[0])
//   w11[4], w21[4]
// k = position of the node in the owner's vertex list (decides the association of the 4-term sum).
// An entry is valid only if the literal routine took its plainest path: one alpha try, all nine
// characteristics inner, both feet of a direction in the same first-ring tetrahedron, the zero-speed
// foot in elements[0] with all weight on the node.  Anything else (or any error) leaves the entry
// invalid and the node on the `cold` list, advanced every stage by the literal routine itself.
// The step-time kernel then is: entry + own record + six gathered records -> 12 interpolated values ->
// sparse Omega rows -> sparse Omega^-1: ~250 instructions per node-stage, bounded by HBM.
struct InnerCacheEntry {
	int a0, b0, c0; unsigned flags;
	float w10[4], w20[4];
	int a1, b1, c1; float w0;
	float w11[4], w21[4];
};
#define GCM_IC_VALID 0x80000000u

// components of a gathered record the stage reads at all: rows 0..5 of Omega
GCM_CONSTEXPR_HD unsigned short gather_mask(int stage) { return u_pattern(stage, 0) | u_pattern(stage, 2) | u_pattern(stage, 4); }

// Builds the entry of (node, stage); returns true when it is valid.  owner_tets (optional, 2 ints):
// the owner per direction, for the parity tap.
GCM_HD bool build_inner_cache_entry(const MeshView& m, int node, int stage, InnerCacheEntry& e, int* owner_tets)
{
	e.a0 = e.b0 = e.c0 = e.a1 = e.b1 = e.c1 = 0; e.flags = 0; e.w0 = 0;
	for(int k = 0; k < 4; k++) { e.w10[k] = e.w20[k] = e.w11[k] = e.w21[k] = 0; }
	if(owner_tets) { owner_tets[0] = -1; owner_tets[1] = -1; }
	if(m.flags[node] & NF_BORDER) return false;
	CurNode cur;
	cur.base = node;
	cur.pos = load_xyz(m, node);
	for(int k = 0; k < 9; k++) cur.val[k] = 0.f;      // values do not enter the search or the factors
	cur.la = m.mat[node]; cur.mu = m.mat[m.stride + node]; cur.rho = m.mat[2*m.stride + node];
	cur.is_border = false; cur.has_contact_data = false;
	for(int k = 0; k < 3; k++) { cur.axis_plus[k] = -1; cur.axis_minus[k] = -1; }
	float ksi[3] = {0.f, 0.f, 0.f};
	ksi[stage] = 1.f;                                   // create_random_axis: basis := E (:712-722)
	float L[9];
	if(elastic_matrix(cur.la, cur.mu, cur.rho, ksi[0], ksi[1], ksi[2], L, nullptr, nullptr)) return false;
	float dksi[9];
	for(int i = 0; i < 9; i++) dksi[i] = - L[i] * m.tau;
	FeetOut feet;
	FootGeom fg[5];
	for(int p = 0; p < 5; p++) for(int k = 0; k < 4; k++) { fg[p].v[k] = -1; fg[p].f[k] = 0; }
	const int outer = find_nodes_on_previous_time_layer(m, cur, stage, 0.0f, dksi, ksi, feet, false, fg);
	if(outer != 0 || feet.n_points != 5) return false;
	const int want[9] = {0, 1, 2, 3, 2, 3, 4, 4, 4};
	for(int i = 0; i < 9; i++) if(feet.ppoint[i] != want[i]) return false;
	for(int p = 0; p < 5; p++) if(feet.owner[p] < 0) return false;
	if(feet.owner[0] != feet.owner[2] || feet.owner[1] != feet.owner[3]) return false;
	if(feet.owner[4] != m.n2t[m.n2t_off[node]]) return false;
	// zero-speed foot: all weight on the node itself
	int k4 = -1;
	for(int k = 0; k < 4; k++) { if(fg[4].v[k] == node) k4 = k; else if(fg[4].f[k] != 0.0f) return false; }
	if(k4 < 0) return false;
	e.w0 = fg[4].f[k4];
	int kk[2];
	for(int d = 0; d < 2; d++) {
		const FootGeom& f1 = fg[d];        // foot -+ c1 tau
		const FootGeom& f2 = fg[2 + d];    // foot -+ c2 tau (same tetrahedron)
		int k = -1;
		for(int j = 0; j < 4; j++) {
			if(f1.v[j] != f2.v[j]) return false;
			if(f1.v[j] == node && k < 0) k = j;
		}
		if(k < 0) return false;            // an owner outside the first ring: the literal routine keeps this node
		kk[d] = k;
		const int ia = f1.v[k == 0 ? 1 : 0], ib = f1.v[k <= 1 ? 2 : 1], ic = f1.v[k == 3 ? 2 : 3];
		float* w1 = d == 0 ? e.w10 : e.w11;
		float* w2 = d == 0 ? e.w20 : e.w21;
		w1[0] = f1.f[k]; w1[1] = f1.f[k == 0 ? 1 : 0]; w1[2] = f1.f[k <= 1 ? 2 : 1]; w1[3] = f1.f[k == 3 ? 2 : 3];
		w2[0] = f2.f[k]; w2[1] = f2.f[k == 0 ? 1 : 0]; w2[2] = f2.f[k <= 1 ? 2 : 1]; w2[3] = f2.f[k == 3 ? 2 : 3];
		if(d == 0) { e.a0 = ia; e.b0 = ib; e.c0 = ic; } else { e.a1 = ia; e.b1 = ib; e.c1 = ic; }
		if(owner_tets) owner_tets[d] = feet.owner[d];
	}
	e.flags = (unsigned)kk[0] | ((unsigned)kk[1] << 2) | GCM_IC_VALID;
	return true;
}

// ((v0 f0 + v1 f1) + v2 f2) + v3 f3 in tet order (TetrMesh_1stOrder.cpp:1864-1877) with the node at
// position k; w = (own, A, B, C).  a + b == b + a, so k = 0 and k = 1 coincide.
GCM_HD float cached_value(int k, const float w[4], float own, float a, float b, float c)
{
	const float tA = a * w[1], tB = b * w[2], tC = c * w[3], tX = own * w[0];
	const float s1 = tA + (k <= 1 ? tX : tB);
	const float s2 = s1 + (k <= 1 ? tB : (k == 2 ? tX : tC));
	return s2 + (k == 3 ? tX : tC);
}

// the components of record i the stage reads (the others are left untouched)
template<int S>
GCM_HD void gather_record(const float* __restrict__ u, size_t rs, int i, float v[9])
{
#if defined(__CUDA_ARCH__)
	const float4 a = rec_quad_ldg(u, rs, i, 0), b = rec_quad_ldg(u, rs, i, 1);
	v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
	if((gather_mask(S) >> 8) & 1) v[8] = rec_quad_ldg(u, rs, i, 2).x; else v[8] = 0.f;
#else
	for(int k = 0; k < 9; k++) v[k] = rec_get(u, rs, i, k);
#endif
}

// The two feet of one direction and their three Riemann invariants (rows dir, 2 + dir, 4 + dir of Omega).
template<int S>
GCM_HD void cached_dir(const float* __restrict__ u, size_t rs, int dir, int k, int ia, int ib, int ic, const float w1[4], const float w2[4],
                       const float own[9], const float* __restrict__ U, float omega[9])
{
	float va[9], vb[9], vc[9];
	gather_record<S>(u, rs, ia, va);
	gather_record<S>(u, rs, ib, vb);
	gather_record<S>(u, rs, ic, vc);
	// omega_i = Omega(i,:) . u(foot of i)  (SimpleVolumeCalculator.cpp:14-23); ppoint_num = (0,1,2,3,2,3,4,4,4)
	float om = 0, om2 = 0, om4 = 0;
	#pragma unroll
	for(int j = 0; j < 9; j++) {
		if((u_pattern(S, 0) >> j) & 1) om += U[dir * 9 + j] * cached_value(k, w1, own[j], va[j], vb[j], vc[j]);
		if(((u_pattern(S, 2) | u_pattern(S, 4)) >> j) & 1) {
			const float v2 = cached_value(k, w2, own[j], va[j], vb[j], vc[j]);
			if((u_pattern(S, 2) >> j) & 1) om2 += U[(2 + dir) * 9 + j] * v2;
			if((u_pattern(S, 4) >> j) & 1) om4 += U[(4 + dir) * 9 + j] * v2;
		}
	}
	omega[dir] = om;
	omega[2 + dir] = om2;
	omega[4 + dir] = om4;
}

// One inner node, axis stage S, from its cache entry: the volume update of
// SimpleVolumeCalculator.cpp:7-32 with the structural zeros of Omega / Omega^-1 skipped.
template<int S>
GCM_HD void inner_cached_stage(const float* __restrict__ u, size_t rs, const InnerCacheEntry& e, const float own[9],
                               const float* __restrict__ U, const float* __restrict__ U1, float res[9])
{
	static_assert(u_pattern(S, 0) == u_pattern(S, 1) && u_pattern(S, 2) == u_pattern(S, 3) && u_pattern(S, 4) == u_pattern(S, 5),
	              "rows 2q and 2q+1 of Omega must share their sparsity pattern");
	float omega[9];
	// dir 0: dksi < 0 (characteristics 0, 2, 4); dir 1: dksi > 0 (1, 3, 5)
	cached_dir<S>(u, rs, 0, (int)(e.flags & 3u), e.a0, e.b0, e.c0, e.w10, e.w20, own, U, omega);
	cached_dir<S>(u, rs, 1, (int)((e.flags >> 2) & 3u), e.a1, e.b1, e.c1, e.w11, e.w21, own, U, omega);
	// zero-speed characteristics: interpolation at the node itself inside elements[0]; every factor
	// but the node's own is exactly 0 (their products are signed zeros, absorbed by the sums below,
	// which start from +0), the own one is w0
	float vz[9];
	#pragma unroll
	for(int k = 0; k < 9; k++) vz[k] = own[k] * e.w0;
	#pragma unroll
	for(int i = 6; i < 9; i++) {
		float om = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u_pattern(S, i) >> j) & 1) om += U[i * 9 + j] * vz[j];
		omega[i] = om;
	}
	// u' = Omega^-1 * omega (SimpleVolumeCalculator.cpp:24-31), skipping the structural zeros
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		float o = 0;
		#pragma unroll
		for(int j = 0; j < 9; j++)
			if((u1_pattern(S, i) >> j) & 1) o += U1[i * 9 + j] * omega[j];
		res[i] = o;
	}
}

// ---- pattern table: the cache without its redundancy ---------------------------------------------
// On a mesh with translational regularity (create_cube's 6 tets per hex, extruded layers, ...) most
// entries are equal once the six vertex ids are written relative to the node: same stencil, same
// factors bit for bit.  After the build the entries of an axis are de-duplicated on the device (hash
// table on the 24 relative words, then an exact word-by-word comparison with the pattern's
// representative, so equal means EQUAL); when an axis has few patterns the step kernel reads a 4-byte
// pattern id per node and the 96-byte pattern from a table that lives in L1/L2 -- 100 instead of 192
// bytes of DRAM traffic per node-stage, same arithmetic on the same numbers.  An irregular mesh (more
// than GCM_IC_MAX_PATTERNS distinct entries on an axis) keeps the per-node entries.
#define GCM_IC_WORDS 24
#define GCM_IC_MAX_PATTERNS 32768
#define GCM_IC_NO_PATTERN 0xffffffffu
#define GCM_IC_SHELL 0x80000000u          // flag on a pattern id: the node's stencil touches another process's halo node

// entry -> the 24 words of its pattern (ids relative to the node)
GCM_HD void ic_relative(const InnerCacheEntry& e, int node, unsigned w[GCM_IC_WORDS])
{
	const unsigned* src = reinterpret_cast<const unsigned*>(&e);
	for(int k = 0; k < GCM_IC_WORDS; k++) w[k] = src[k];
	w[0] = (unsigned)(e.a0 - node); w[1] = (unsigned)(e.b0 - node); w[2] = (unsigned)(e.c0 - node);
	w[12] = (unsigned)(e.a1 - node); w[13] = (unsigned)(e.b1 - node); w[14] = (unsigned)(e.c1 - node);
}

// pattern + node -> entry
GCM_HD void ic_absolute(const InnerCacheEntry& pat, int node, InnerCacheEntry& e)
{
	e = pat;
	e.a0 = pat.a0 + node; e.b0 = pat.b0 + node; e.c0 = pat.c0 + node;
	e.a1 = pat.a1 + node; e.b1 = pat.b1 + node; e.c1 = pat.c1 + node;
}

#if defined(__CUDACC__)

// One thread per listed inner node: the literal routine once, its outcome into the cache
// (and into `cold`, cold[0] = count, when the node has to stay with the literal routine).
__global__ void __launch_bounds__(128)
build_inner_cache_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list, int stage,
                         InnerCacheEntry* __restrict__ cache, int* __restrict__ cold, int* __restrict__ owner_tap)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	InnerCacheEntry e;
	int own2[2];
	const bool ok = build_inner_cache_entry(m, node, stage, e, own2);
	uint4* o = reinterpret_cast<uint4*>(cache + node);
	const uint4* src = reinterpret_cast<const uint4*>(&e);
	#pragma unroll
	for(int q = 0; q < 6; q++) o[q] = src[q];
	if(owner_tap) { owner_tap[2 * (size_t)node] = own2[0]; owner_tap[2 * (size_t)node + 1] = own2[1]; }
	if(!ok && stage == 0) cold[1 + atomicAdd(&cold[0], 1)] = node;
}

// stage > 0: a node invalid on this axis only joins the list too (once)
__global__ void __launch_bounds__(256)
merge_cold_kernel(const InnerCacheEntry* __restrict__ c0, const InnerCacheEntry* __restrict__ c1, const InnerCacheEntry* __restrict__ c2,
                  const int* __restrict__ list, int n_list, int* __restrict__ cold)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const bool v0 = (c0[node].flags & GCM_IC_VALID) != 0, v1 = (c1[node].flags & GCM_IC_VALID) != 0, v2 = (c2[node].flags & GCM_IC_VALID) != 0;
	if(v0 && !(v1 && v2)) cold[1 + atomicAdd(&cold[0], 1)] = node;   // !v0: already listed by the stage-0 build
}

// The hot kernel.  One thread per node of a contiguous range; the entry says whether the node is
// an inner node of the cached path at all (all-zero entries: border, halo, generic, cold).
// PLASTIC: stage 3 folded in (only instantiated for S == 2).
template<int S, bool PLASTIC>
__global__ void __launch_bounds__(256)
stage_inner_cached_kernel(const float* __restrict__ u, size_t rs, int node_begin, int n_range, float* __restrict__ un,
                          const InnerCacheEntry* __restrict__ cache, const MatTab* __restrict__ tab, int n_mat,
                          const unsigned char* __restrict__ mat_id, const float* __restrict__ yield_plane,
                          int* err, int zone, StageTap* tap, const int* __restrict__ owner_tap,
                          const int* __restrict__ n2t_off, const int* __restrict__ n2t, int n_nodes)
{
	__shared__ float sU[GCM_MAX_MATERIALS][81];
	__shared__ float sU1[GCM_MAX_MATERIALS][81];
	for(int i = threadIdx.x; i < n_mat * 81; i += blockDim.x) {
		const MatTab& mt = tab[(i / 81) * 3 + S];
		sU[i / 81][i % 81] = mt.U[i % 81];
		sU1[i / 81][i % 81] = mt.U1[i % 81];
	}
	__syncthreads();
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_range) return;
	const int node = node_begin + idx;
	InnerCacheEntry e;
	{
		const uint4* src = reinterpret_cast<const uint4*>(cache + node);
		uint4* dst = reinterpret_cast<uint4*>(&e);
		const uint4 q0 = __ldg(src);
		if(!(q0.w & GCM_IC_VALID)) return;
		dst[0] = q0;
		#pragma unroll
		for(int q = 1; q < 6; q++) dst[q] = __ldg(src + q);
	}
	const float4 o0 = rec_quad(u, rs, node, 0), o1 = rec_quad(u, rs, node, 1), o2 = rec_quad(u, rs, node, 2);
	const float own[9] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x};
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) bad |= !(fabsf(own[k]) <= 3.4028235e38f);   // NaN or Inf (TetrMethod_Plastic_1stOrder.cpp:192-194)
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }
	const int mid = mat_id ? mat_id[node] : 0;
	float res[9];
	inner_cached_stage<S>(u, rs, e, own, sU[mid], sU1[mid], res);
	if(tap) {
		StageTap tp;
		tp.owner[0] = tp.owner[2] = owner_tap ? owner_tap[2 * (size_t)node] : -2;
		tp.owner[1] = tp.owner[3] = owner_tap ? owner_tap[2 * (size_t)node + 1] : -2;
		tp.owner[4] = n2t[n2t_off[node]];
		tp.outer_count = 0; tp.tries = 1;
		tap[(size_t)S * n_nodes + node] = tp;
	}
	if(PLASTIC) {
		if(!plastic_epilogue(res, yield_plane[node])) { report_error(err, ERR_NAN, zone, node, 3); return; }
	}
	rec_quad_store(un, rs, node, 0, make_float4(res[0], res[1], res[2], res[3]));
	rec_quad_store(un, rs, node, 1, make_float4(res[4], res[5], res[6], res[7]));
	rec_quad_store(un, rs, node, 2, make_float4(res[8], o2.y, o2.z, o2.w));
}

// --- de-duplication of one axis of the cache (see "pattern table" above) ---
// table: `cap` (power of two) slots of (64-bit key, representative node); stats[0] = distinct keys,
// stats[1] = failure flag (table full / hash collision between different entries), stats[2] = patterns numbered.
__device__ __forceinline__ unsigned long long ic_hash(const unsigned w[GCM_IC_WORDS])
{
	unsigned long long h = 0x9E3779B97F4A7C15ull;
	#pragma unroll
	for(int k = 0; k < GCM_IC_WORDS; k++) {
		h ^= w[k];
		h *= 0xBF58476D1CE4E5B9ull;
		h ^= h >> 29;
	}
	return h | 1ull;        // 0 marks an empty slot
}

__global__ void __launch_bounds__(256)
ic_dedup_insert_kernel(const InnerCacheEntry* __restrict__ cache, const int* __restrict__ list, int n_list,
                       unsigned long long* __restrict__ keys, int* __restrict__ rep, unsigned cap, int* __restrict__ slot_of, int* __restrict__ stats)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const InnerCacheEntry e = cache[node];
	if(!(e.flags & GCM_IC_VALID)) { slot_of[node] = -1; return; }
	unsigned w[GCM_IC_WORDS];
	ic_relative(e, node, w);
	const unsigned long long key = ic_hash(w);
	unsigned h = (unsigned)(key >> 20) & (cap - 1);
	for(unsigned probe = 0; probe < cap; probe++) {
		const unsigned long long old = atomicCAS(&keys[h], 0ull, key);
		if(old == 0ull) { atomicAdd(&stats[0], 1); }
		if(old == 0ull || old == key) { atomicMin(&rep[h], node); slot_of[node] = (int)h; return; }
		h = (h + 1) & (cap - 1);
		if(probe > 4096) break;
	}
	slot_of[node] = -1;
	atomicExch(&stats[1], 1);
}

__global__ void __launch_bounds__(256)
ic_dedup_number_kernel(const InnerCacheEntry* __restrict__ cache, const int* __restrict__ list, int n_list,
                       const int* __restrict__ rep, const int* __restrict__ slot_of, int* __restrict__ slot_pid,
                       InnerCacheEntry* __restrict__ patterns, int max_patterns, int* __restrict__ stats)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const int slot = slot_of[node];
	if(slot < 0) return;
	const int r = rep[slot];
	unsigned w[GCM_IC_WORDS], wr[GCM_IC_WORDS];
	ic_relative(cache[node], node, w);
	if(r != node) {
		// equal hash is not enough: the entry must equal its pattern word for word
		ic_relative(cache[r], r, wr);
		bool same = true;
		#pragma unroll
		for(int k = 0; k < GCM_IC_WORDS; k++) same = same && (w[k] == wr[k]);
		if(!same) atomicExch(&stats[1], 1);
		return;
	}
	const int pid = atomicAdd(&stats[2], 1);
	if(pid >= max_patterns) { atomicExch(&stats[1], 1); return; }
	slot_pid[slot] = pid;
	unsigned* dst = reinterpret_cast<unsigned*>(patterns + pid);
	#pragma unroll
	for(int k = 0; k < GCM_IC_WORDS; k++) dst[k] = w[k];
}

__global__ void __launch_bounds__(256)
ic_dedup_assign_kernel(const int* __restrict__ list, int n_list, const int* __restrict__ slot_of, const int* __restrict__ slot_pid,
                       unsigned* __restrict__ pat_idx)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const int slot = slot_of[node];
	pat_idx[node] = slot < 0 ? GCM_IC_NO_PATTERN : (unsigned)slot_pid[slot];
}

// Flags the nodes whose pattern (any axis) gathers a node of `peer_halo` (1 = halo node owned by another process)
// and lists them: shell[0] = count, shell[1..] = nodes.
__global__ void __launch_bounds__(256)
ic_mark_shell_kernel(const int* __restrict__ list, int n_list, unsigned* __restrict__ pat_idx, const InnerCacheEntry* __restrict__ patterns,
                     int n_nodes, int max_patterns, const unsigned char* __restrict__ peer_halo, int* __restrict__ shell)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	bool touches = false;
	for(int st = 0; st < 3; st++) {
		const unsigned pid = pat_idx[(size_t)st * n_nodes + node];
		if(pid == GCM_IC_NO_PATTERN) return;                 // not a node of the pattern path
		const InnerCacheEntry& p = patterns[(size_t)st * max_patterns + pid];
		touches = touches || peer_halo[p.a0 + node] || peer_halo[p.b0 + node] || peer_halo[p.c0 + node]
		                  || peer_halo[p.a1 + node] || peer_halo[p.b1 + node] || peer_halo[p.c1 + node];
	}
	if(!touches) return;
	for(int st = 0; st < 3; st++) pat_idx[(size_t)st * n_nodes + node] |= GCM_IC_SHELL;
	shell[1 + atomicAdd(&shell[0], 1)] = node;
}

// The hot kernel over the pattern table: identical to stage_inner_cached_kernel but for where the
// entry comes from (pattern id -> pattern -> absolute vertex ids).
// MB = resident blocks per SM the register allocation aims at (1: unconstrained; occupancy against a few spilled words)
template<int S, bool PLASTIC, int MB>
__global__ void __launch_bounds__(256, MB)
stage_inner_pattern_kernel(const float* __restrict__ u, size_t rs, int node_begin, int n_range, float* __restrict__ un,
                           const unsigned* __restrict__ pat_idx, const InnerCacheEntry* __restrict__ patterns,
                           const MatTab* __restrict__ tab, int n_mat,
                           const unsigned char* __restrict__ mat_id, const float* __restrict__ yield_plane,
                           int* err, int zone, StageTap* tap, const int* __restrict__ owner_tap,
                           const int* __restrict__ n2t_off, const int* __restrict__ n2t, int n_nodes,
                           const int* __restrict__ shell_list, int n_shell)
{
	__shared__ float sU[GCM_MAX_MATERIALS][81];
	__shared__ float sU1[GCM_MAX_MATERIALS][81];
	for(int i = threadIdx.x; i < n_mat * 81; i += blockDim.x) {
		const MatTab& mt = tab[(i / 81) * 3 + S];
		sU[i / 81][i % 81] = mt.U[i % 81];
		sU1[i / 81][i % 81] = mt.U1[i % 81];
	}
	__syncthreads();
	// Two ways to be launched: over the contiguous node range [node_begin, node_begin + n_range) (shell_list == NULL:
	// nodes flagged GCM_IC_SHELL are left to the other launch), or over shell_list (its first n_shell entries, those
	// inside the range): the inner nodes whose stencil touches a halo node owned by another process, which have to
	// wait for the exchange while everything else runs beside it (gcm_capi.cu: run_stage).
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	int node;
	unsigned pid;
	if(shell_list) {
		if(idx >= n_shell) return;
		node = shell_list[idx];
		if(node < node_begin || node >= node_begin + n_range) return;
		pid = __ldg(pat_idx + node) & ~GCM_IC_SHELL;
	} else {
		if(idx >= n_range) return;
		node = node_begin + idx;
		pid = __ldg(pat_idx + node);
		if(pid & GCM_IC_SHELL) return;          // GCM_IC_NO_PATTERN has the bit too
	}
	const float4 o0 = rec_quad(u, rs, node, 0), o1 = rec_quad(u, rs, node, 1), o2 = rec_quad(u, rs, node, 2);
	InnerCacheEntry e;
	{
		const uint4* src = reinterpret_cast<const uint4*>(patterns + pid);
		uint4* dst = reinterpret_cast<uint4*>(&e);
		#pragma unroll
		for(int q = 0; q < 6; q++) dst[q] = __ldg(src + q);
		e.a0 += node; e.b0 += node; e.c0 += node; e.a1 += node; e.b1 += node; e.c1 += node;
	}
	const float own[9] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x};
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) bad |= !(fabsf(own[k]) <= 3.4028235e38f);   // NaN or Inf (TetrMethod_Plastic_1stOrder.cpp:192-194)
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }
	const int mid = mat_id ? mat_id[node] : 0;
	float res[9];
	inner_cached_stage<S>(u, rs, e, own, sU[mid], sU1[mid], res);
	if(tap) {
		StageTap tp;
		tp.owner[0] = tp.owner[2] = owner_tap ? owner_tap[2 * (size_t)node] : -2;
		tp.owner[1] = tp.owner[3] = owner_tap ? owner_tap[2 * (size_t)node + 1] : -2;
		tp.owner[4] = n2t[n2t_off[node]];
		tp.outer_count = 0; tp.tries = 1;
		tap[(size_t)S * n_nodes + node] = tp;
	}
	if(PLASTIC) {
		if(!plastic_epilogue(res, yield_plane[node])) { report_error(err, ERR_NAN, zone, node, 3); return; }
	}
	rec_quad_store(un, rs, node, 0, make_float4(res[0], res[1], res[2], res[3]));
	rec_quad_store(un, rs, node, 1, make_float4(res[4], res[5], res[6], res[7]));
	rec_quad_store(un, rs, node, 2, make_float4(res[8], o2.y, o2.z, o2.w));
}

#endif // __CUDACC__

} // namespace gcm
