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
__device__ __forceinline__ void load_cand(Cand& c, const int4 x, const float4* __restrict__ rec, const float2* __restrict__ hv2)
{
	c.ia = x.x; c.ib = x.y; c.ic = x.z;
	c.t = x.w & 0x3fffffff; c.k = (unsigned)x.w >> 30;
	c.qa = __ldg(rec + 3 * (size_t)c.ia + 2);
	c.qb = __ldg(rec + 3 * (size_t)c.ib + 2);
	c.qc = __ldg(rec + 3 * (size_t)c.ic + 2);
	const float2 hv = __ldg(hv2 + c.t);
	c.h = hv.x; c.vol = hv.y;
}

// Cold path: a node the specialised scan cannot settle (the two feet of a direction disagree on a
// candidate, or no owner in the first ring) goes through the literal per-node routine, which also
// classifies the failure exactly as the reference does.
__device__ __noinline__ void inner_node_generic(const MeshView& m, int node, int stage, float* __restrict__ un,
                                                int* err, int zone, StageTap* tap)
{
	float out[9];
	StageTap t;
	const int e = node_stage_nocontact(m, node, stage, out, tap ? &t : nullptr);
	if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
	if(e) { report_error(err, e, zone, node, stage); return; }
	const float* r = m.u + (size_t)node * GCM_REC;
	float4* __restrict__ o = reinterpret_cast<float4*>(un) + 3 * (size_t)node;
	o[0] = make_float4(out[0], out[1], out[2], out[3]);
	o[1] = make_float4(out[4], out[5], out[6], out[7]);
	o[2] = make_float4(out[8], r[9], r[10], r[11]);
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

__device__ __forceinline__ void load_values(const Cand& c, const float4* __restrict__ rec, float va[9], float vb[9], float vc[9])
{
	const float4 a0 = __ldg(rec + 3 * (size_t)c.ia), a1 = __ldg(rec + 3 * (size_t)c.ia + 1);
	const float4 b0 = __ldg(rec + 3 * (size_t)c.ib), b1 = __ldg(rec + 3 * (size_t)c.ib + 1);
	const float4 c0 = __ldg(rec + 3 * (size_t)c.ic), c1 = __ldg(rec + 3 * (size_t)c.ic + 1);
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
                   const float* __restrict__ w0, const int4* __restrict__ n2x)
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
	const float2* __restrict__ hv2 = reinterpret_cast<const float2*>(m.tet_hv);
	const float4 o0 = rec[3 * (size_t)node], o1 = rec[3 * (size_t)node + 1], o2 = rec[3 * (size_t)node + 2];
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
			load_cand<S>(c, x, rec, hv2);
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
			inner_node_generic(m, node, S, un, err, zone, tap);
			return;
		}
		tp.owner[dir] = owner;
		tp.owner[2 + dir] = owner;

		float va[9], vb[9], vc[9], v[9];
		load_values(c, rec, va, vb, vc);
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
	float4* __restrict__ out = reinterpret_cast<float4*>(un) + 3 * (size_t)node;
	out[0] = make_float4(res[0], res[1], res[2], res[3]);
	out[1] = make_float4(res[4], res[5], res[6], res[7]);
	out[2] = make_float4(res[8], X.x, X.y, X.z);
}

// ---- angular candidate index ------------------------------------------------------------------
// Built once on the device after the mesh upload.  The pre-filter's verdict for a candidate
// depends on static data only (the vertex coordinates, h(tet) and the SIGN of the displacement
// along axis S -- the length is rescaled away, see the file header), so the candidates it rejects
// ("certainly fails") are the same on every step.  Per (node, axis, sign) the index keeps the
// positions of the surviving candidates as a 32-bit mask over the node's list and, inline, the
// expanded-CSR entry and (h, Vol) of the FIRST survivor -- with the rejects gone that one is the
// owner almost always, and having it in the index saves the two scattered 64-byte DRAM accesses
// (CSR entry, tet_hv) per direction that dominated the kernel's traffic.  The per-step search is
// unchanged: survivors in list order through the same pre-filter / exact predicate, first pass
// wins.  Same function, same inputs, same compiler flags at build and at step time -> the skipped
// verdicts are exactly the -1 ones.  Entry of one (axis, node), 4 x uint4 = 64 bytes:
//   q0 = first survivor, dksi < 0: (ia, ib, ic, t | k << 30)
//   q1 = (its h, its Vol, survivors after it, min h over the whole list)
//   q2, q3 = the same for dksi > 0, with w0 in the last word (sign bit: "shell", see the kernel)
// so that each direction reads its own 32 bytes when it needs them.
// min h == 0 marks "not a node of the fast inner path" (all-zero entry: border, halo, generic);
// ia == -1 "no survivor".  The step-time guard  0 < delta < 0.99 * min h  implies the pre-filter's
// own precondition delta < 0.99 h(tet) for every candidate (rounding of h * 0.99 is monotone in
// h); a node that fails it, or has more than 32 candidates (min h stored as -1), takes the
// literal per-node routine.
#define GCM_DIDX_QUADS 4
template<int S>
__global__ void __launch_bounds__(128)
build_dir_index_kernel(const float* __restrict__ u, const int* __restrict__ n2t_off, const float* __restrict__ tet_hv,
                       const int* __restrict__ list, int n_list, const float* __restrict__ w0,
                       const int4* __restrict__ n2x, uint4* __restrict__ didx)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	const float4* __restrict__ rec = reinterpret_cast<const float4*>(u);
	const float2* __restrict__ hv2 = reinterpret_cast<const float2*>(tet_hv);
	const float4 o2 = rec[3 * (size_t)node + 2];
	F3 X; X.x = o2.y; X.y = o2.z; X.z = o2.w;
	const int e0 = n2t_off[node], e1 = n2t_off[node + 1];
	unsigned mm = 0, mp = 0;
	float hmin = -1.0f;
	if(e1 - e0 <= 32) {
		for(int e = e0; e < e1; e++) {
			Cand c;
			load_cand<S>(c, __ldg(n2x + e), rec, hv2);
			hmin = (e == e0 || c.h < hmin) ? c.h : hmin;
			const bool shaped = c.vol > 0;
			if(!(shaped && prefilter<S>(c, X, -1.0f) < 0)) mm |= 1u << (e - e0);
			if(!(shaped && prefilter<S>(c, X, 1.0f) < 0)) mp |= 1u << (e - e0);
		}
	}
	int4 xm = make_int4(-1, -1, -1, 0), xp = make_int4(-1, -1, -1, 0);
	float2 hm = make_float2(0.f, 0.f), hp = make_float2(0.f, 0.f);
	if(mm) { xm = __ldg(n2x + e0 + __ffs(mm) - 1); hm = __ldg(hv2 + (xm.w & 0x3fffffff)); mm &= mm - 1; }
	if(mp) { xp = __ldg(n2x + e0 + __ffs(mp) - 1); hp = __ldg(hv2 + (xp.w & 0x3fffffff)); mp &= mp - 1; }
	uint4* __restrict__ o = didx + GCM_DIDX_QUADS * (size_t)node;
	o[0] = make_uint4((unsigned)xm.x, (unsigned)xm.y, (unsigned)xm.z, (unsigned)xm.w);
	o[1] = make_uint4(__float_as_uint(hm.x), __float_as_uint(hm.y), mm, __float_as_uint(hmin));
	o[2] = make_uint4((unsigned)xp.x, (unsigned)xp.y, (unsigned)xp.z, (unsigned)xp.w);
	o[3] = make_uint4(__float_as_uint(hp.x), __float_as_uint(hp.y), mp, __float_as_uint(w0[node]));
}

// Sets the "shell" bit (sign of w0) of the listed nodes in the three planes of the index.
__global__ void mark_shell_kernel(uint4* __restrict__ didx, size_t n_nodes, const int* __restrict__ list, int n_list)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	for(int sgl = 0; sgl < 3; sgl++) {
		uint4* e = didx + GCM_DIDX_QUADS * ((size_t)sgl * n_nodes + node);
		if(e[1].w != 0u) e[3].w |= 0x80000000u;
	}
}

// ---- inner kernel, indexed -------------------------------------------------------------------
// One thread per NODE of a contiguous node range (no list indirection): the node's index entry
// says whether it is a fast inner node at all.  Two rounds of dependent loads per node-stage:
// (index entry, own record) -> (records of the two first survivors' vertices: coordinates for the
// predicate and values for the interpolation in one go).  A direction whose first survivor is not
// the owner continues down its mask through the expanded CSR.  Arithmetic identical to
// stage_inner_kernel.
struct CandFull {
	Cand c;
	float4 a0, a1, b0, b1, c0, c1;     // first two record quads of A, B, C (the third is c.qa/qb/qc)
};

__device__ __forceinline__ void load_value_quads(CandFull& f, const float4* __restrict__ rec)
{
	f.a0 = __ldg(rec + 3 * (size_t)f.c.ia); f.a1 = __ldg(rec + 3 * (size_t)f.c.ia + 1);
	f.b0 = __ldg(rec + 3 * (size_t)f.c.ib); f.b1 = __ldg(rec + 3 * (size_t)f.c.ib + 1);
	f.c0 = __ldg(rec + 3 * (size_t)f.c.ic); f.c1 = __ldg(rec + 3 * (size_t)f.c.ic + 1);
}

// values = false: only the coordinate quads (what the search and the interpolation factors need);
// the value quads then follow in feet_of_dir<S, true>, when the factors are known
__device__ __forceinline__ void load_records(CandFull& f, const float4* __restrict__ rec, bool values = true)
{
	f.c.qa = __ldg(rec + 3 * (size_t)f.c.ia + 2);
	f.c.qb = __ldg(rec + 3 * (size_t)f.c.ib + 2);
	f.c.qc = __ldg(rec + 3 * (size_t)f.c.ic + 2);
	if(values) load_value_quads(f, rec);
}

__device__ __forceinline__ void set_entry(CandFull& f, const uint4 x, float h, float vol)
{
	f.c.ia = (int)x.x; f.c.ib = (int)x.y; f.c.ic = (int)x.z;
	f.c.t = (int)(x.w & 0x3fffffffu); f.c.k = (int)(x.w >> 30);
	f.c.h = h; f.c.vol = vol;
}

// Owner search of one direction starting from the (already loaded) first survivor; `rest` = the
// later survivors.  Returns false when the node must take the literal routine.
template<int S, bool LATE>
__device__ __forceinline__ bool settle_dir(CandFull& f, unsigned rest, int node, const int* __restrict__ n2t_off, const F3& X, float sgn,
                                           float dk1, float de1, float dk2, float de2,
                                           const int4* __restrict__ n2x, const float4* __restrict__ rec, const float2* __restrict__ hv2)
{
	for(;;) {
		const Cand& c = f.c;
		int v1 = 0;
		if(c.vol > 0) v1 = prefilter<S>(c, X, sgn);
		if(v1 == 0) return false;   // within 1e-3 of a face: the literal routine decides (deferred)
		if(v1 > 0) return true;
		if(!rest) return false;
		const int4 x = __ldg(n2x + n2t_off[node] + __ffs(rest) - 1);
		rest &= rest - 1;
		const float2 hv = __ldg(hv2 + (x.w & 0x3fffffff));
		set_entry(f, make_uint4((unsigned)x.x, (unsigned)x.y, (unsigned)x.z, (unsigned)x.w), hv.x, hv.y);
		load_records(f, rec, !LATE);
	}
}

// Interpolation factors of the foot X + dks e_S inside candidate c, as (own, A, B, C)
// (TetrMesh_1stOrder.cpp:1764-1857).  Returns false when the reference throws.
template<int S>
__device__ __forceinline__ bool axis_factors(const Cand& c, const F3& X, float dks, float w[4])
{
	F3 Fp = X;
	if(S == 0) Fp.x = X.x + dks; else if(S == 1) Fp.y = X.y + dks; else Fp.z = X.z + dks;
	F3 p0, p1, p2, p3;
	tet_order(c, X, p0, p1, p2, p3);
	float f[4];
	if(!interp_factors(to_vec3(p0), to_vec3(p1), to_vec3(p2), to_vec3(p3), Fp.x, Fp.y, Fp.z, c.vol, f))
		return false;
	w[0] = c.k == 0 ? f[0] : (c.k == 1 ? f[1] : (c.k == 2 ? f[2] : f[3]));
	w[1] = c.k == 0 ? f[1] : f[0];
	w[2] = c.k <= 1 ? f[2] : f[1];
	w[3] = c.k == 3 ? f[2] : f[3];
	return true;
}

// ((v0 f0 + v1 f1) + v2 f2) + v3 f3 in tet order (:1858-1871); a + b == b + a, so k = 0 and k = 1 coincide
__device__ __forceinline__ float axis_value(int k, const float w[4], float own, float a, float b, float c)
{
	const float tA = a * w[1], tB = b * w[2], tC = c * w[3], tX = own * w[0];
	const float s1 = tA + (k <= 1 ? tX : tB);
	const float s2 = s1 + (k <= 1 ? tB : (k == 2 ? tX : tC));
	return s2 + (k == 3 ? tX : tC);
}

// The two feet of one direction inside its owner and their three Riemann invariants
// (rows dir, 2 + dir, 4 + dir of Omega).  Returns false when the reference throws.
// LATE: the value quads of the owner's vertices are requested here, after the factors of both feet
// (they are 24 registers that need not be live during the search and the determinants).
template<int S, bool LATE>
__device__ __forceinline__ bool feet_of_dir(CandFull& f, int dir, const F3& X, float dk1, float dk2,
                                            const float own[9], const float* __restrict__ U, float omega[9],
                                            const float4* __restrict__ rec)
{
	float w1[4], w2[4];
	if(!axis_factors<S>(f.c, X, dk1, w1)) return false;
	if(!axis_factors<S>(f.c, X, dk2, w2)) return false;
	if(LATE) load_value_quads(f, rec);
	const float va[9] = {f.a0.x, f.a0.y, f.a0.z, f.a0.w, f.a1.x, f.a1.y, f.a1.z, f.a1.w, f.c.qa.x};
	const float vb[9] = {f.b0.x, f.b0.y, f.b0.z, f.b0.w, f.b1.x, f.b1.y, f.b1.z, f.b1.w, f.c.qb.x};
	const float vc[9] = {f.c0.x, f.c0.y, f.c0.z, f.c0.w, f.c1.x, f.c1.y, f.c1.z, f.c1.w, f.c.qc.x};
	// omega_i = Omega(i,:) . u(foot of i)  (SimpleVolumeCalculator.cpp:14-23); ppoint_num = (0,1,2,3,2,3,4,4,4)
	float om = 0, om2 = 0, om4 = 0;
	#pragma unroll
	for(int j = 0; j < 9; j++) {
		if((u_pattern(S, 0) >> j) & 1) om += U[dir * 9 + j] * axis_value(f.c.k, w1, own[j], va[j], vb[j], vc[j]);
		if(((u_pattern(S, 2) | u_pattern(S, 4)) >> j) & 1) {
			const float v2 = axis_value(f.c.k, w2, own[j], va[j], vb[j], vc[j]);
			if((u_pattern(S, 2) >> j) & 1) om2 += U[(2 + dir) * 9 + j] * v2;
			if((u_pattern(S, 4) >> j) & 1) om4 += U[(4 + dir) * 9 + j] * v2;
		}
	}
	omega[dir] = om;
	omega[2 + dir] = om2;
	omega[4 + dir] = om4;
	return true;
}

// PF: 1 = the records of both directions' first survivors are requested together, 0 = the second
// direction's after the first direction's search, 2 = after the first direction's interpolation
// (fewest live registers).
// list / skip_shell: see the body.
// A node the index cannot settle (guard, disagreeing feet, no owner in the first ring) is appended to
// `cold` (cold[0] = count) and advanced by stage_deferred_kernel right after this launch: keeping the
// literal routine out of this kernel leaves its hot path spill-free at 128 registers (-7 % time).
template<int S, int MB, int PF>
__global__ void __launch_bounds__(128, MB)
stage_inner_direct_kernel(const __grid_constant__ MeshView m, int node_begin, int n_range,
                          float* __restrict__ un, int* err, int zone, StageTap* tap,
                          const MatTab* __restrict__ tab, int n_mat, const unsigned char* __restrict__ mat_id,
                          const int4* __restrict__ n2x, const uint4* __restrict__ didx, int* __restrict__ cold,
                          const int* __restrict__ list, int skip_shell)
{
	static_assert(u_pattern(S, 0) == u_pattern(S, 1) && u_pattern(S, 2) == u_pattern(S, 3) && u_pattern(S, 4) == u_pattern(S, 5),
	              "rows 2q and 2q+1 of Omega must share their sparsity pattern");
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
	if(idx >= n_range) return;
	// list == NULL: the contiguous range; else the listed nodes (the shell pass of the overlapped stage)
	const int node = list ? list[idx] : node_begin + idx;
	const uint4* __restrict__ ent = didx + GCM_DIDX_QUADS * (size_t)node;
	const uint4 e0q = __ldg(ent), e1q = __ldg(ent + 1);
	if(e1q.w == 0u) return;                       // min h == 0: not a node of the fast inner path
	uint4 e2q = make_uint4(0u, 0u, 0u, 0u), e3q = e2q;
	if(PF != 2 || skip_shell) { e2q = __ldg(ent + 2); e3q = __ldg(ent + 3); }
	// sign bit of w0 (a weight, >= 0) = "shell": within two rings of a halo node owned by another process;
	// the deep pass leaves those to the pass that runs after the exchange
	if(skip_shell && (e3q.w & 0x80000000u)) return;
	const float4* __restrict__ rec = reinterpret_cast<const float4*>(m.u);
	const float2* __restrict__ hv2 = reinterpret_cast<const float2*>(m.tet_hv);
	const float4 o0 = rec[3 * (size_t)node], o1 = rec[3 * (size_t)node + 1], o2 = rec[3 * (size_t)node + 2];
	const bool have = (int)e0q.x >= 0 && (PF == 2 || (int)e2q.x >= 0);
	CandFull f0, f1;
	set_entry(f0, e0q, __uint_as_float(e1q.x), __uint_as_float(e1q.y));
	if(PF != 2) set_entry(f1, e2q, __uint_as_float(e3q.x), __uint_as_float(e3q.y));
	if(have) load_records(f0, rec, PF != 2);
	if(PF == 1 && have) load_records(f1, rec);

	const int mid = mat_id ? mat_id[node] : 0;
	const float own[9] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x};
	F3 X; X.x = o2.y; X.y = o2.z; X.z = o2.w;
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) bad |= !(fabsf(own[k]) <= 3.4028235e38f);   // NaN or Inf (TetrMethod_Plastic_1stOrder.cpp:192-194): one compare
	if(bad) { report_error(err, ERR_NAN, zone, node, S); return; }
	const float* U = sU[mid];
	const float* U1 = sU1[mid];
	float dk[4], de[4];
	{
		// all four displacements must be rescaled by every candidate (then the index is valid)
		const double h99 = (double)__uint_as_float(e1q.w) * 0.99;
		bool ok = have;
		#pragma unroll
		for(int q = 0; q < 4; q++) {
			dk[q] = - sL[mid][q] * m.tau;             // dksi[i] = - L(i,i) * time_step
			de[q] = sqrtf(dk[q] * dk[q]);             // delta; off-axis components are +-0
			ok = ok && de[q] > 0 && (double)de[q] < h99;
		}
		if(!ok) { cold[1 + atomicAdd(&cold[0], 1)] = node; return; }
	}
	float omega[9];
	// dir 0: dksi < 0 (characteristics 0 and 2: -c1 tau, -c2 tau); dir 1: dksi > 0 (1 and 3)
	if(!settle_dir<S, PF == 2>(f0, e1q.z, node, m.n2t_off, X, -1.0f, dk[0], de[0], dk[2], de[2], n2x, rec, hv2)) {
		cold[1 + atomicAdd(&cold[0], 1)] = node; return;
	}
	if(PF == 0) load_records(f1, rec);
	if(!feet_of_dir<S, PF == 2>(f0, 0, X, dk[0], dk[2], own, U, omega, rec)) { report_error(err, ERR_INTERP_SUM, zone, node, S); return; }
	if(PF == 2) {
		// nothing of the second direction is live while the first one is interpolated: its entry is
		// read again (the line is in L1) and its records are requested only now
		e2q = __ldg(ent + 2); e3q = __ldg(ent + 3);
		if((int)e2q.x < 0) { cold[1 + atomicAdd(&cold[0], 1)] = node; return; }   // no survivor on this side
		set_entry(f1, e2q, __uint_as_float(e3q.x), __uint_as_float(e3q.y));
		load_records(f1, rec, false);
	}
	if(!settle_dir<S, PF == 2>(f1, e3q.z, node, m.n2t_off, X, 1.0f, dk[1], de[1], dk[3], de[3], n2x, rec, hv2)) {
		cold[1 + atomicAdd(&cold[0], 1)] = node; return;
	}
	if(!feet_of_dir<S, PF == 2>(f1, 1, X, dk[1], dk[3], own, U, omega, rec)) { report_error(err, ERR_INTERP_SUM, zone, node, S); return; }

	// zero-speed characteristics: interpolation at the node itself inside elements[0]; every
	// factor but the node's own is exactly 0, the own one is the static weight w0
	float vz[9];
	const float wz = __uint_as_float(e3q.w & 0x7fffffffu);
	#pragma unroll
	for(int k = 0; k < 9; k++) vz[k] = own[k] * wz;
	if(tap) {
		StageTap tp;
		tp.owner[0] = f0.c.t; tp.owner[2] = f0.c.t; tp.owner[1] = f1.c.t; tp.owner[3] = f1.c.t;
		tp.owner[4] = m.n2t[m.n2t_off[node]];
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
	float4* __restrict__ out = reinterpret_cast<float4*>(un) + 3 * (size_t)node;
	out[0] = make_float4(res[0], res[1], res[2], res[3]);
	out[1] = make_float4(res[4], res[5], res[6], res[7]);
	out[2] = make_float4(res[8], X.x, X.y, X.z);
}

#endif // __CUDACC__

} // namespace gcm
