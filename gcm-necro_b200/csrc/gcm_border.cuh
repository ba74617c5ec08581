// gcm_border.cuh -- one axis stage of a border node without contact, the B200-first replacement of
// the literal per-node routine (which stays as the fallback for whatever this one declines).
//
// What the reference computes for such a node (gcm/methods/TetrMethod_Plastic_1stOrder.cpp:190-297,
// first alpha try) is kept operation for operation wherever a value is produced; what is redesigned is
// everything around it:
//   * owner search: every first-ring candidate's barycentric gradients, scaled by 0.99 h(tet), are
//     static geometry and are precomputed once in double (gcm_host.cpp: build_border_flat_tables),
//     stored warp-transposed so that 32 nodes read 32 consecutive quads.  point_in_tetr rescales every
//     short displacement to 0.99 h (TetrMesh_1stOrder.cpp:1304-1307), so a candidate's verdict for the
//     direction q is three dot products q^ . G -- one walk over the list serves both senses and all
//     four moving feet.  The verdict is conservative (certainly fails / certainly passes / don't know,
//     thresholds MeshView::pf_fail / pf_pass); "don't know", a foot the rescale does not cover or a ring
//     expansion hand the node to the literal routine (returns -1).
//   * the alpha-retry loop of prepare_node (:99-108) collapses: tries 1..9 only shorten the bent feet
//     along the same direction, which the rescale undoes, so their verdicts are those of try 0 and none
//     of them can decide the node; a node try 0 does not decide (edges and corners on their tangential
//     axes) is decided by the 11th try, whose bent feet all sit at the centre of elements[0] (:569-584).
//     Two evaluations instead of eleven, tries reported as the reference counts them.
//   * no 81-float Omega / Omega^-1 per thread: the rows are generated from the three direction vectors
//     where they are consumed (same expressions as ElasticMatrix3D.cpp:371-517, rows 2q / 2q+1 differ
//     by sign only), the Riemann invariants are accumulated straight from the interpolated values.
//   * the 9x9 double system of the border calculators (gcm/methods/border/*.cpp + GSL LU) lives in a
//     caller-provided matrix -- shared memory, element-major across the block, on the device -- and the
//     elimination keeps the pivot row in registers; same pivoting, same operation order as lu_solve<9>.
// The routine is __host__ __device__: tests/emul runs it on the CPU against the reference's fixtures.
#pragma once
#include "gcm_step.cuh"

namespace gcm {

#if defined(__CUDACC__)
typedef float4 gcm_f4;
#else
struct gcm_f4 { float x, y, z, w; };
#endif

// Static per border slot (merged index).  deg == 0: not a node of this path (sharp, or its zero-speed
// foot does not sit in elements[0] with all weight on the node).
struct BndSlotStatic {
	int deg;        // candidates = len(elements)
	int e0;         // first CSR entry of the node (n2x[e0 + c] = candidate c)
	float hmin;     // min h(tet) over the candidates
	float dmin2;    // min squared distance node <-> candidate vertex under the reference's scan rule (:1541-1545)
	float w0;       // own factor of the zero-speed foot
	int cand_off;   // index (in quads) of candidate 0, quad 0 of this node in the transposed table
	float h0;       // h(elements[0]): the last-chance foot (centre of elements[0]) is rescaled to 0.99 h0 when nearer
	int pad1;
};
#define GCM_BND_LANES 32      // candidates of one warp of flat nodes are interleaved: quad (c, q) of lane l at off + (3 c + q) * 32 + l
#define GCM_BND_DEFER (-1)
#if defined(GCM_BND_STATS) && !defined(__CUDA_ARCH__)
static long g_bnd_defer_reason[16];
#define GCM_BND_DEFER_R(r) (__sync_fetch_and_add(&g_bnd_defer_reason[r], 1), GCM_BND_DEFER)
#else
#define GCM_BND_DEFER_R(r) GCM_BND_DEFER
#endif

// 9x9 double system in caller-provided storage
struct LocalMat9 { double a[81]; GCM_HD double& at(int i, int j) { return a[i * 9 + j]; } };
template<int STRIDE> struct StridedMat9 { double* base; GCM_HD double& at(int i, int j) { return base[(i * 9 + j) * STRIDE]; } };

// lu_solve<9> (gcm_math.cuh: GSL LU_decomp + LU_solve semantics) on a Mat: the row swaps are applied
// to b as they happen (equal to gathering b[perm[i]] afterwards), the pivot row sits in registers while
// the rows below it are eliminated.  Every product, difference and quotient is the one lu_solve makes.
template<class Mat>
GCM_HD void lu_solve9(Mat& A, double b[9])
{
	#pragma unroll
	for(int j = 0; j + 1 < 9; j++) {
		double mx = fabs(A.at(j, j));
		int ip = j;
		#pragma unroll
		for(int i = j + 1; i < 9; i++) {
			const double a = fabs(A.at(i, j));
			if(a > mx) { mx = a; ip = i; }
		}
		double prow[9];
		if(ip != j) {
			#pragma unroll
			for(int k = 0; k < 9; k++) {
				const double t = A.at(j, k), u = A.at(ip, k);
				A.at(j, k) = u; A.at(ip, k) = t;
				prow[k] = u;
			}
			#pragma unroll
			for(int i = j + 1; i < 9; i++)
				if(i == ip) { const double t = b[j]; b[j] = b[i]; b[i] = t; }
		} else {
			#pragma unroll
			for(int k = 0; k < 9; k++) prow[k] = (k >= j) ? A.at(j, k) : 0.0;
		}
		const double ajj = prow[j];
		if(ajj != 0.0) {
			#pragma unroll
			for(int i = j + 1; i < 9; i++) {
				const double aij = div_z(A.at(i, j), ajj);
				A.at(i, j) = aij;
				#pragma unroll
				for(int k = j + 1; k < 9; k++)
					A.at(i, k) = A.at(i, k) - aij * prow[k];
			}
		}
	}
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		double t = b[i];
		#pragma unroll
		for(int j = 0; j < i; j++) t -= A.at(i, j) * b[j];
		b[i] = t;
	}
	#pragma unroll
	for(int i = 8; i >= 0; i--) {
		double t = b[i];
		#pragma unroll
		for(int j = i + 1; j < 9; j++) t -= A.at(i, j) * b[j];
		b[i] = div_z(t, A.at(i, i));
	}
}

// The eigen-directions of one (node, axis) and what the rows of Omega / Omega^-1 are made of.
struct BndDirs {
	float n[3][3];            // n0 = q/|q|, n1, n2 (ElasticMatrix3D.cpp:284-339)
	float l, c1, c2, c3;
	float la, mu, ro;
};

// N_ab[i] of ElasticMatrix3D.cpp:522-530
GCM_HD float bnd_N(const BndDirs& d, int a, int b, int i)
{
	const int r = i == 0 ? 0 : (i <= 2 ? 0 : (i <= 4 ? 1 : 2));     // (xx,xy,xz,yy,yz,zz) -> first index
	const int c = i == 0 ? 0 : (i == 1 ? 1 : (i == 2 ? 2 : (i == 3 ? 1 : (i == 4 ? 2 : 2))));
	return (d.n[a][r] * d.n[b][c] + d.n[b][r] * d.n[a][c]) / 2;
}

// Row i of Omega (U), ElasticMatrix3D.cpp:429-517 -- the expressions of elastic_matrix() / elastic_U_row()
GCM_HD void bnd_U_row(const BndDirs& d, int i, float row[9])
{
	const float w[6] = {1, 2, 2, 1, 2, 1};
	const int a = i < 2 ? 0 : (i < 4 ? 1 : (i < 6 ? 2 : -1));
	#pragma unroll
	for(int k = 0; k < 3; k++) row[k] = a >= 0 ? d.n[a][k] : 0;
	#pragma unroll
	for(int q = 0; q < 6; q++) {
		float v;
		switch(i) {
		case 0: v = div_z(-w[q] * bnd_N(d, 0, 0, q), d.c1 * d.ro); break;
		case 1: v = div_z(w[q] * bnd_N(d, 0, 0, q), d.c1 * d.ro); break;
		case 2: v = div_z(-w[q] * bnd_N(d, 0, 1, q), d.c2 * d.ro); break;
		case 3: v = div_z(w[q] * bnd_N(d, 0, 1, q), d.c2 * d.ro); break;
		case 4: v = div_z(-w[q] * bnd_N(d, 0, 2, q), d.c2 * d.ro); break;
		case 5: v = div_z(w[q] * bnd_N(d, 0, 2, q), d.c2 * d.ro); break;
		case 6: v = w[q] * bnd_N(d, 1, 2, q); break;
		case 7: v = w[q] * (bnd_N(d, 1, 1, q) - bnd_N(d, 2, 2, q)); break;
		default: v = w[q] * (bnd_N(d, 1, 1, q) + bnd_N(d, 2, 2, q) - 2 * d.la * bnd_N(d, 0, 0, q) / (d.la + 2 * d.mu)); break;
		}
		row[3 + q] = v;
	}
}

// Rows 0..5 of Omega come in pairs that differ by the sign of their stress part (-x / y == -(x / y) in
// IEEE arithmetic): the 18 quotients are made once per (node, axis).
struct BndRows { float s[3][6]; };
GCM_HD void bnd_rows_prepare(const BndDirs& d, BndRows& R)
{
	const float w[6] = {1, 2, 2, 1, 2, 1};
	#pragma unroll
	for(int q = 0; q < 6; q++) {
		R.s[0][q] = div_z(w[q] * bnd_N(d, 0, 0, q), d.c1 * d.ro);
		R.s[1][q] = div_z(w[q] * bnd_N(d, 0, 1, q), d.c2 * d.ro);
		R.s[2][q] = div_z(w[q] * bnd_N(d, 0, 2, q), d.c2 * d.ro);
	}
}
// Row i of Omega from the prepared quotients (bitwise the row bnd_U_row gives)
GCM_HD void bnd_U_row_p(const BndDirs& d, const BndRows& R, int i, float row[9])
{
	if(i >= 6) { bnd_U_row(d, i, row); return; }
	const int a = i >> 1;
	#pragma unroll
	for(int k = 0; k < 3; k++) row[k] = d.n[a][k];
	#pragma unroll
	for(int q = 0; q < 6; q++) row[3 + q] = (i & 1) ? R.s[a][q] : -R.s[a][q];
}

// Row r of Omega^-1 (U1), ElasticMatrix3D.cpp:371-427
GCM_HD void bnd_U1_row(const BndDirs& d, int r, float row[9])
{
	const float c1 = d.c1, c2 = d.c2, c3 = d.c3, ro = d.ro;
	if(r < 3) {
		row[0] = d.n[0][r]; row[1] = d.n[0][r];
		row[2] = d.n[1][r]; row[3] = d.n[1][r];
		row[4] = d.n[2][r]; row[5] = d.n[2][r];
		row[6] = 0; row[7] = 0; row[8] = 0;
	} else {
		const int i = r - 3;
		const float I = (i == 0 || i == 3 || i == 5) ? 1.f : 0.f;
		const float N00 = bnd_N(d, 0, 0, i), N01 = bnd_N(d, 0, 1, i), N02 = bnd_N(d, 0, 2, i);
		row[0] = -ro * ((c1 - c3) * N00 + c3 * I);
		row[1] = ro * ((c1 - c3) * N00 + c3 * I);
		row[2] = -2 * ro * c2 * N01;
		row[3] = 2 * ro * c2 * N01;
		row[4] = -2 * ro * c2 * N02;
		row[5] = 2 * ro * c2 * N02;
		row[6] = 4 * bnd_N(d, 1, 2, i);
		row[7] = bnd_N(d, 1, 1, i) - bnd_N(d, 2, 2, i);
		row[8] = I - N00;
	}
	#pragma unroll
	for(int c = 0; c < 9; c++) row[c] /= 2;
}

// Condition row k of the five border calculators written into row i of the system (zeroed by the caller).
template<class Mat>
GCM_HD void bnd_cond_row(const BorderCond& bc, int k, const float outer_normal[3], const float local_n[3][3],
                         Mat& A, int i, double* rhs)
{
	const float scale = bc.scale;
	switch(bc.type) {
	case BC_FREE: {
		// FreeBorderCalculator.cpp:59-74
		const int c0 = k == 0 ? 3 : (k == 1 ? 4 : 5), c1 = k == 0 ? 4 : (k == 1 ? 6 : 7), c2 = k == 0 ? 5 : (k == 1 ? 7 : 8);
		A.at(i, c0) = (double)outer_normal[0];
		A.at(i, c1) = (double)outer_normal[1];
		A.at(i, c2) = (double)outer_normal[2];
	} break;
	case BC_FIXED:
		A.at(i, k) = 1;      // FixedBorderCalculator.cpp:52-58
		break;
	case BC_EXT_FORCE: {
		// ExternalForceCalculator.cpp:76-110
		const float* n0 = local_n[0];
		if(k == 0) {
			A.at(i, 3) = (double)(n0[0] * n0[0]);
			A.at(i, 4) = (double)(2 * n0[1] * n0[0]);
			A.at(i, 5) = (double)(2 * n0[2] * n0[0]);
			A.at(i, 6) = (double)(n0[1] * n0[1]);
			A.at(i, 7) = (double)(2 * n0[2] * n0[1]);
			A.at(i, 8) = (double)(n0[2] * n0[2]);
			*rhs = (double)(scale * bc.p_normal);
		} else {
			const float* t = local_n[k];
			A.at(i, 3) = (double)(n0[0] * t[0]);
			A.at(i, 4) = (double)(n0[1] * t[0] + n0[0] * t[1]);
			A.at(i, 5) = (double)(n0[2] * t[0] + n0[0] * t[2]);
			A.at(i, 6) = (double)(n0[1] * t[1]);
			A.at(i, 7) = (double)(n0[2] * t[1] + n0[1] * t[2]);
			A.at(i, 8) = (double)(n0[2] * t[2]);
			*rhs = (double)(scale * bc.p_tangential * (t[0]*bc.dir[0] + t[1]*bc.dir[1] + t[2]*bc.dir[2]));
		}
	} break;
	case BC_EXT_VELOCITY: {
		// ExternalVelocityCalculator.cpp:70-98
		const float* t = local_n[k];
		A.at(i, 0) = (double)t[0];
		A.at(i, 1) = (double)t[1];
		A.at(i, 2) = (double)t[2];
		if(k == 0) *rhs = (double)(scale * bc.p_normal);
		else *rhs = (double)(scale * bc.p_tangential * (t[0]*bc.dir[0] + t[1]*bc.dir[1] + t[2]*bc.dir[2]));
	} break;
	case BC_EXT_VALUES:
		// ExternalValuesCalculator.cpp:63-78
		A.at(i, bc.vars_index[k]) = 1;
		*rhs = (double)bc.vars_values[k];
		break;
	}
}

// The two feet of one sense inside their owner: interpolated values (TetrMesh_1stOrder.cpp:1764-1880).
// x = (ia, ib, ic, t | k << 30): expanded CSR entry of the owner.  Returns a StepError.
GCM_HD int bnd_feet_values(const MeshView& m, int node, const Vec3& X, const float own[9], const int x[4],
                           const float fc1[3], const float fc2[3], float v1[9], float v2[9])
{
	const int ia = x[0], ib = x[1], ic = x[2];
	const int k = (int)((unsigned)x[3] >> 30);
	const int t = x[3] & 0x3fffffff;
	float ra[12], rb[12], rc[12];
	rec_load12(m.u, m.rs, ia, ra);
	rec_load12(m.u, m.rs, ib, rb);
	rec_load12(m.u, m.rs, ic, rc);
	Vec3 A, B, C;
	A.x = ra[9]; A.y = ra[10]; A.z = ra[11];
	B.x = rb[9]; B.y = rb[10]; B.z = rb[11];
	C.x = rc[9]; C.y = rc[10]; C.z = rc[11];
	const Vec3 p0 = k == 0 ? X : A;
	const Vec3 p1 = k == 0 ? A : (k == 1 ? X : B);
	const Vec3 p2 = k <= 1 ? B : (k == 2 ? X : C);
	const Vec3 p3 = k == 3 ? X : C;
	const float vol = m.tet_hv[2*(size_t)t + 1];
	float f1[4], f2[4];
	if(!interp_factors(p0, p1, p2, p3, fc1[0], fc1[1], fc1[2], vol, f1)) return ERR_INTERP_SUM;
	if(!interp_factors(p0, p1, p2, p3, fc2[0], fc2[1], fc2[2], vol, f2)) return ERR_INTERP_SUM;
	#pragma unroll
	for(int j = 0; j < 9; j++) {
		const float a0 = k == 0 ? own[j] : ra[j];
		const float a1 = k == 0 ? ra[j] : (k == 1 ? own[j] : rb[j]);
		const float a2 = k <= 1 ? rb[j] : (k == 2 ? own[j] : rc[j]);
		const float a3 = k == 3 ? own[j] : rc[j];
		v1[j] = interp4(a0, a1, a2, a3, f1);
		v2[j] = interp4(a0, a1, a2, a3, f2);
	}
	return ERR_NONE;
}

// Conservative verdict of candidate quads (g0, g1, g2) for the unit direction (qx, qy, qz):
// v1 for the sense +q, v0 for -q (-1 certainly fails, +1 certainly passes, 0 don't know).
GCM_HD void bnd_verdicts(const gcm_f4& g0, const gcm_f4& g1, const gcm_f4& g2, float qx, float qy, float qz,
                         float pf_fail, float pf_pass, int& v0, int& v1)
{
	// barycentric coordinates of X + 0.99 h q^ w.r.t. A, B, C; those of the opposite sense are the negatives
	const float la = qx * g0.x + qy * g0.y + qz * g0.z;
	const float lb = qx * g0.w + qy * g1.x + qz * g1.y;
	const float lc = qx * g1.z + qy * g1.w + qz * g2.x;
	const float sum = la + lb + lc;
	const float mn1 = fminf(fminf(la, lb), fminf(lc, 1.0f - sum));
	const float mn0 = fminf(fminf(-la, -lb), fminf(-lc, 1.0f + sum));
	v1 = mn1 < pf_fail ? -1 : (mn1 > pf_pass ? 1 : 0);
	v0 = mn0 < pf_fail ? -1 : (mn0 > pf_pass ? 1 : 0);
}

// One border node without contact, axis `stage`.  cand = the node's candidate quads (already offset to
// its lane; quad (c, q) at cand[(3 c + q) * GCM_BND_LANES]).  A = storage of the 9x9 system or NULL
// (then a node that needs a border solve is deferred).  Returns a StepError, or GCM_BND_DEFER.
template<class Mat>
GCM_HD int node_stage_border_fast(const MeshView& m, const BndSlotStatic& bs, const gcm_f4* __restrict__ cand,
                                  int node, int stage, float out[9], StageTap* tap, Mat* A)
{
	if(bs.deg <= 0 || m.exact_only) return GCM_BND_DEFER_R(1);
	float own[9];
	Vec3 X;
	{
		float r[12];
		rec_load12(m.u, m.rs, node, r);
		#pragma unroll
		for(int k = 0; k < 9; k++) own[k] = r[k];
		X.x = r[9]; X.y = r[10]; X.z = r[11];
	}
	{
		bool bad = false;
		#pragma unroll
		for(int k = 0; k < 9; k++) bad |= !(fabsf(own[k]) <= 3.4028235e38f);   // NaN or Inf (:192-194)
		if(bad) return ERR_NAN;
	}
	BndDirs d;
	d.la = m.mat[node]; d.mu = m.mat[m.stride + node]; d.ro = m.mat[2*m.stride + node];
	const int slot = m.border_slot[node];
	const float* basis = m.basis + 9*(size_t)slot;
	const float ksi[3] = {basis[3*stage], basis[3*stage + 1], basis[3*stage + 2]};
	if((d.la <= 0) || (d.mu <= 0) || (d.ro <= 0)) return ERR_BAD_RHEOLOGY;
	if(ksi[0]*ksi[0] + ksi[1]*ksi[1] + ksi[2]*ksi[2] <= 0) return ERR_BAD_Q;
	elastic_axes(ksi[0], ksi[1], ksi[2], d.n, &d.l);
	d.c1 = sqrtf((d.la + 2*d.mu) / d.ro);
	d.c2 = sqrtf(d.mu / d.ro);
	d.c3 = sqrtf(d.la*d.la / (d.ro*(d.la + 2*d.mu)));
	// dksi[i] = - L(i,i) * tau of the five distinct points (characteristics 0, 1, 2, 3, 6)
	float dk[5];
	dk[0] = - (d.c1 * d.l) * m.tau;
	dk[1] = - (-d.c1 * d.l) * m.tau;
	dk[2] = - (d.c2 * d.l) * m.tau;
	dk[3] = - (-d.c2 * d.l) * m.tau;
	dk[4] = - 0.0f * m.tau;
	{
		// the dedupe of :479-490 must give the five-point pattern (0,1,2,3,2,3,4,4,4)
		bool usual = true;
		#pragma unroll
		for(int p = 1; p < 5; p++)
			#pragma unroll
			for(int q = 0; q < p; q++)
				if( (double)fabsf(dk[p] - dk[q]) <= 0.01 * (double)fabsf(dk[p] + dk[q]) ) usual = false;
		if(!usual || !(dk[0] < 0) || !(dk[2] < 0) || !(dk[1] > 0) || !(dk[3] > 0)) return GCM_BND_DEFER_R(2);
	}
	// displacements and feet of the four moving points on the first try (real node, no contact, alpha == 0: :553-558)
	float fc[4][3], R2[4];
	#pragma unroll
	for(int p = 0; p < 4; p++) {
		float dx[3];
		#pragma unroll
		for(int j = 0; j < 3; j++) {
			dx[j] = dk[p] * ksi[j] + 0.0f;
			fc[p][j] = (j == 0 ? X.x : (j == 1 ? X.y : X.z)) + dx[j];
		}
		R2[p] = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
		const float delta = sqrtf(R2[p]);
		// the shared scan answers only for feet every candidate rescales (the bent feet of tries 1..9 are
		// shorter still, and longer than 0 down to alpha = 0.9)
		if(!(delta > 0) || !((double)delta < (double)bs.hmin * 0.99) || !(0.05f * delta > 0)) return GCM_BND_DEFER_R(3);
	}

	// ---- one walk over the first ring, both senses (0: -q, 1: +q); four candidates' quads in flight ----
	int own_c[2] = {-1, -1};
	{
		const float qn = sqrtf(ksi[0]*ksi[0] + ksi[1]*ksi[1] + ksi[2]*ksi[2]);
		const float iq = 1.0f / qn;                 // the filter is conservative: its rounding is free
		const float qx = ksi[0] * iq, qy = ksi[1] * iq, qz = ksi[2] * iq;
		bool open0 = true, open1 = true;
		for(int c0 = 0; c0 < bs.deg && (open0 || open1); c0 += 4) {
			gcm_f4 g[4][3];
			#pragma unroll
			for(int j = 0; j < 4; j++)        // the table is padded: reading up to 3 candidates past the list is in bounds
				#pragma unroll
				for(int q = 0; q < 3; q++) g[j][q] = cand[(3 * (c0 + j) + q) * GCM_BND_LANES];
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				if(c0 + j >= bs.deg || !(open0 || open1)) break;
				if(!(g[j][2].y > 0)) return GCM_BND_DEFER_R(4);    // ill-shaped tetrahedron: only the literal predicate may judge it
				int v0, v1;
				bnd_verdicts(g[j][0], g[j][1], g[j][2], qx, qy, qz, m.pf_fail, m.pf_pass, v0, v1);
				if(open0) { if(v0 > 0) { own_c[0] = c0 + j; open0 = false; } else if(v0 == 0) return GCM_BND_DEFER_R(5); }
				if(open1) { if(v1 > 0) { own_c[1] = c0 + j; open1 = false; } else if(v1 == 0) return GCM_BND_DEFER_R(6); }
			}
		}
	}
	// a sense without owner: NULL only if no candidate vertex lies within |dx| (else the reference expands the ring)
	bool inn[5];
	#pragma unroll
	for(int p = 0; p < 4; p++) {
		inn[p] = own_c[p & 1] >= 0;
		if(!inn[p] && bs.dmin2 < R2[p]) return GCM_BND_DEFER_R(7);
	}
	inn[4] = true;     // dksi == 0: the node itself (interpolated inside elements[0], weight w0)
	int outer_count = (inn[0] ? 0 : 1) + (inn[1] ? 0 : 1) + (inn[2] ? 0 : 2) + (inn[3] ? 0 : 2);
	int tries = 1;
	// ---- the retry loop (:99-108).  Tries 1..9 move the bent feet (dksi < 0) to dksi (1 - alpha) ksi: same
	// direction, rescaled by point_in_tetr to the same tested point, so the same verdicts and the same
	// outer count as the first try.  The 11th (alpha = 1) puts every foot but the outward ones of the
	// normal axis at the centre of elements[0] (:569-584).
	bool centre[2] = {false, false};       // per sense: its feet sit at the centre of elements[0]
	float fcc[3] = {0.f, 0.f, 0.f};
	int x0[4] = {0, 0, 0, 0};
	if(outer_count != 0 && outer_count != 3) {
		tries = 11;
		{
			const int* x = m.n2x_scan + 4*(size_t)bs.e0;
#if defined(__CUDA_ARCH__)
			const int4 q4 = *reinterpret_cast<const int4*>(x);
			x0[0] = q4.x; x0[1] = q4.y; x0[2] = q4.z; x0[3] = q4.w;
#else
			x0[0] = x[0]; x0[1] = x[1]; x0[2] = x[2]; x0[3] = x[3];
#endif
		}
		const int k = (int)((unsigned)x0[3] >> 30);
		const Vec3 Pa = load_xyz(m, x0[0]), Pb = load_xyz(m, x0[1]), Pc = load_xyz(m, x0[2]);
		const Vec3 p0 = k == 0 ? X : Pa, p1 = k == 0 ? Pa : (k == 1 ? X : Pb), p2 = k <= 1 ? Pb : (k == 2 ? X : Pc), p3 = k == 3 ? X : Pc;
		float dxc[3];
		dxc[0] = ( p0.x + p1.x + p2.x + p3.x ) / 4 - X.x;
		dxc[1] = ( p0.y + p1.y + p2.y + p3.y ) / 4 - X.y;
		dxc[2] = ( p0.z + p1.z + p2.z + p3.z ) / 4 - X.z;
		fcc[0] = X.x + dxc[0]; fcc[1] = X.y + dxc[1]; fcc[2] = X.z + dxc[2];
		// the search from X along dxc must stop at elements[0] (candidate 0) with certainty
		const float dl = sqrtf(dxc[0]*dxc[0] + dxc[1]*dxc[1] + dxc[2]*dxc[2]);
		if(!(dl > 0)) return GCM_BND_DEFER_R(8);
		const gcm_f4 g0 = cand[0], g1 = cand[GCM_BND_LANES], g2 = cand[2 * GCM_BND_LANES];
		if(!(g2.y > 0)) return GCM_BND_DEFER_R(8);
		// tested point: X + 0.99 h0 dxc/|dxc| when |dxc| < 0.99 h0, else X + dxc itself; G = 0.99 h0 grad
		const bool resc = (double)dl < (double)bs.h0 * 0.99;
		const float sc = resc ? 1.0f / dl : 1.0f / (0.99f * bs.h0);
		int v0, v1;
		bnd_verdicts(g0, g1, g2, dxc[0] * sc, dxc[1] * sc, dxc[2] * sc, m.pf_fail, m.pf_pass, v0, v1);
		if(v1 <= 0) return GCM_BND_DEFER_R(8);
		centre[0] = true;                   // dksi < 0: always
		centre[1] = stage != 0;             // dksi > 0: unless border node on its normal axis (:569)
		#pragma unroll
		for(int p = 0; p < 4; p++) if(centre[p & 1]) inn[p] = true;
		outer_count = (inn[0] ? 0 : 1) + (inn[1] ? 0 : 1) + (inn[2] ? 0 : 2) + (inn[3] ? 0 : 2);
	}
	if(tap) {
		// owners of the last try
		#pragma unroll
		for(int sgn = 0; sgn < 2; sgn++) {
			int t = -1;
			if(centre[sgn]) t = x0[3] & 0x3fffffff;
			else if(own_c[sgn] >= 0) t = m.n2x_scan[4*(size_t)(bs.e0 + own_c[sgn]) + 3] & 0x3fffffff;
			tap->owner[sgn] = t; tap->owner[2 + sgn] = t;
		}
		tap->owner[4] = m.n2t[m.n2t_off[node]];
		tap->outer_count = outer_count; tap->tries = tries;
	}
	if(outer_count != 0 && outer_count != 3) return ERR_OUTER_COUNT;     // :234-239, after the last try
	if(outer_count == 3 && !A) return GCM_BND_DEFER_R(9);

	// ---- Riemann invariants omega_i = Omega(i,:) . u(foot of i), ppoint_num = (0,1,2,3,2,3,4,4,4) ----
	BndRows R;
	bnd_rows_prepare(d, R);
	float omega[9];
	float row[9];
	// the 11th try: the zero-speed foot (dksi = -0, not > 0) moves to the centre of elements[0] as well (:569-584)
	float vz[9];
	if(tries == 11) {
		float vdup[9];
		const int e = bnd_feet_values(m, node, X, own, x0, fcc, fcc, vz, vdup);
		if(e) return e;
	} else {
		#pragma unroll
		for(int k = 0; k < 9; k++) vz[k] = own[k] * bs.w0;
	}
	#pragma unroll
	for(int sgn = 0; sgn < 2; sgn++) {
		omega[sgn] = 0; omega[2 + sgn] = 0; omega[4 + sgn] = 0;
		if(!inn[sgn]) continue;
		if(centre[sgn]) {
			bnd_U_row_p(d, R, sgn, row);
			omega[sgn] = omega_row(row, 0, vz);
			bnd_U_row_p(d, R, 2 + sgn, row);
			omega[2 + sgn] = omega_row(row, 0, vz);
			bnd_U_row_p(d, R, 4 + sgn, row);
			omega[4 + sgn] = omega_row(row, 0, vz);
			continue;
		}
		float v1[9], v2[9];
		int xs[4];
		const int* x = m.n2x_scan + 4*(size_t)(bs.e0 + own_c[sgn]);
#if defined(__CUDA_ARCH__)
		const int4 q4 = *reinterpret_cast<const int4*>(x);
		xs[0] = q4.x; xs[1] = q4.y; xs[2] = q4.z; xs[3] = q4.w;
#else
		xs[0] = x[0]; xs[1] = x[1]; xs[2] = x[2]; xs[3] = x[3];
#endif
		const int e = bnd_feet_values(m, node, X, own, xs, fc[sgn], fc[2 + sgn], v1, v2);
		if(e) return e;
		bnd_U_row_p(d, R, sgn, row);
		omega[sgn] = omega_row(row, 0, v1);
		bnd_U_row_p(d, R, 2 + sgn, row);
		omega[2 + sgn] = omega_row(row, 0, v2);
		bnd_U_row_p(d, R, 4 + sgn, row);
		omega[4 + sgn] = omega_row(row, 0, v2);
	}
	#pragma unroll
	for(int i = 6; i < 9; i++) {
		bnd_U_row(d, i, row);
		omega[i] = omega_row(row, 0, vz);
	}

	if(outer_count == 0) {
		// SimpleVolumeCalculator.cpp:24-31
		#pragma unroll
		for(int r = 0; r < 9; r++) {
			bnd_U1_row(d, r, row);
			float v = 0;
			#pragma unroll
			for(int j = 0; j < 9; j++) v += row[j] * omega[j];
			out[r] = v;
		}
		return ERR_NONE;
	}

	// ---- outer_count == 3: the border calculators (gcm/methods/border/*.cpp) ----
	BorderCond def;
	def.type = BC_FREE; def.scale = 1; def.active = 1;
	const BorderCond* bc = &def;
	const int cid = m.cond_id ? m.cond_id[slot] : -1;
	if(cid >= 0 && m.conds[cid].active) bc = &m.conds[cid];
	float local_n[3][3];
	local_n[0][0] = ksi[0]; local_n[0][1] = ksi[1]; local_n[0][2] = ksi[2];
	local_n[1][0] = local_n[1][1] = local_n[1][2] = 0; local_n[2][0] = local_n[2][1] = local_n[2][2] = 0;
	if(bc->type == BC_EXT_FORCE || bc->type == BC_EXT_VELOCITY)
		create_local_basis(local_n[0], local_n[1], local_n[2]);
	double b[9];
	const int pp[9] = {0, 1, 2, 3, 2, 3, 4, 4, 4};
	int k = 0;
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		if(inn[pp[i]]) {
			bnd_U_row_p(d, R, i, row);
			b[i] = (double)omega[i];
			#pragma unroll
			for(int j = 0; j < 9; j++) A->at(i, j) = (double)row[j];
		} else {
			b[i] = 0;
			#pragma unroll
			for(int j = 0; j < 9; j++) A->at(i, j) = 0;
			if(k < 3) bnd_cond_row(*bc, k, ksi, local_n, *A, i, &b[i]);
			k++;
		}
	}
	lu_solve9(*A, b);
	#pragma unroll
	for(int j = 0; j < 9; j++)
		out[j] = (float)b[j];
	return ERR_NONE;
}

} // namespace gcm
