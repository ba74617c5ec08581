// gcm_math.cuh -- scalar building blocks of the grid-characteristic step.
//
// Every function restates, operation by operation and in the reference's float/double mix,
// one routine of avasyukov/gcm-necro (cited per function as gcm/<file>:<lines>).  The
// reference is built for x86-64 without FMA contraction, so this translation unit MUST be
// compiled with  nvcc -fmad=false -prec-div=true -prec-sqrt=true -ftz=false  (device) or
// g++ -ffp-contract=off (the CPU emulation used by tests/ only).  Where a fused multiply-add
// is wanted on purpose (conservative pre-filters) it is spelled __fmaf_rn explicitly.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GCM_HD __host__ __device__ __forceinline__
#define GCM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GCM_HD inline
#define GCM_HD_NOINLINE
#endif

namespace gcm {

// ---- error codes written by kernels (first error wins) ------------------------------------
// Mapped by the host wrapper to the reference's GCMException codes (gcm/system/GCMException.h:19-29).
enum StepError {
	ERR_NONE = 0,
	ERR_NAN = 1,              // "NAN or INF detected. Instability?"  TetrMethod_Plastic_1stOrder.cpp:192-194
	ERR_OUTER_COUNT = 2,      // "Illegal number of outer characteristics"  :234-239
	ERR_NOT_BORDER = 3,       // "Border node is not marked as border"  :268-272
	ERR_BOTH_CONTACT = 4,     // "Both direction of contact are marked as contact"  :280-284
	ERR_INTERP_SUM = 5,       // "Sum of factors is greater than 1.0"  TetrMesh_1stOrder.cpp:1855
	ERR_INNER_NO_OWNER = 6,   // inner node without owner: find_border_cross path (:619-643), not built
	ERR_RING_OVERFLOW = 7,    // owner search ring expansion exceeded the device scratch
	ERR_BAD_RHEOLOGY = 8,     // "Bad rheology" ElasticMatrix3D.cpp:28-29
	ERR_VIRT_OUTER_COUNT = 9, // virt node outer count != 3  TetrMethod_Plastic_1stOrder.cpp:348-356
	ERR_BAD_Q = 10,           // "Bad vector q" ElasticMatrix3D.cpp:31-32
	ERR_ZERO_VOLUME = 11,     // "Tetr volume is zero" / "Tri area is zero" TetrMesh_1stOrder.cpp:1010-1015
	ERR_FOOT_PATTERN = 12,    // characteristic dedupe did not give the 5-point pattern (0,1,2,3,2,3,4,4,4)
	ERR_PEER_TIMEOUT = 13,    // a neighbour process never delivered its halo (direct peer-store exchange)
	ERR_COLLISION_NAN = 14    // "NAN or INF in p1 / p2 / p3 / p0 / v / n / virt node coords"  TetrMesh_1stOrder.cpp:2196-2230,2297-2299
};

// x / y with the IEEE result, arranged so that a zero numerator never reaches the divider.
// nvcc expands a division into an inline fast path plus a subroutine for "special" operands, and
// a ZERO numerator is special: one lane with x == 0 (a foot on a tet face, a structural zero of
// the border system -- both the common case here) drags its whole warp through the subroutine;
// ncu showed a quarter of the border kernel's instructions there.  0 / y = +-0 for any non-zero,
// non-NaN y, so those lanes divide 1 / y instead and select the signed zero afterwards.
// (The device build spells the division as the __fdiv_rn / __ddiv_rn intrinsic -- the same
// round-to-nearest IEEE operation -- because the optimiser folds  x == 0 ? +-0 : x / c  back into
// the plain x / c for a constant c, which is exactly what this helper is here to avoid.)
GCM_HD float div_z(float x, float y)
{
	const bool z = (x == 0.0f) && (fabsf(y) > 0.0f);
#if defined(__CUDA_ARCH__)
	const float q = __fdiv_rn(z ? 1.0f : x, y);
#else
	const float q = (z ? 1.0f : x) / y;
#endif
	const float zr = copysignf(0.0f, x);
	return z ? (signbit(y) ? -zr : zr) : q;
}
GCM_HD double div_z(double x, double y)
{
	const bool z = (x == 0.0) && (fabs(y) > 0.0);
#if defined(__CUDA_ARCH__)
	const double q = __ddiv_rn(z ? 1.0 : x, y);
#else
	const double q = (z ? 1.0 : x) / y;
#endif
	const double zr = copysign(0.0, x);
	return z ? (signbit(y) ? -zr : zr) : q;
}

// x / 6 (the `/ 6` of quick_math.cpp:9), correctly rounded, without the divider: with
// c = RN(1/6),  q0 = RN(x c),  r = x - 6 q0 (exact in an FMA),  q = RN(q0 + r c)  equals RN(x / 6)
// for every float with biased exponent 2..254 -- checked exhaustively over all 2^32 bit patterns
// by tests/emul/div6_check.c (the only misses are zeros/denormals/results that go denormal, and
// infinities); those, never seen in practice, take the real division.  A third of the inner
// kernel's instructions were inline division expansions before this.
GCM_HD float div6(float x)
{
#if defined(__CUDA_ARCH__)
	const float ax = fabsf(x);
	if(ax >= 2.3509887e-38f && ax <= 3.4028235e38f) {     // 2^-125 <= |x| <= FLT_MAX
		const float c = 0.16666667163372039794921875f;
		const float q0 = __fmul_rn(x, c);
		const float r = __fmaf_rn(-6.0f, q0, x);
		return __fmaf_rn(r, c, q0);
	}
	if(x == 0.0f) return x;
	return __fdiv_rn(x, 6.0f);
#else
	return x / 6;
#endif
}

// gcm/system/quick_math.cpp:26-29
GCM_HD float determinant(float x1, float y1, float z1, float x2, float y2, float z2, float x3, float y3, float z3)
{
	return (x1*(y2*z3-y3*z2) - x2*(y1*z3-y3*z1) + x3*(y1*z2-y2*z1));
}

// gcm/system/quick_math.cpp:7-10
GCM_HD float tetr_volume(float x1, float y1, float z1, float x2, float y2, float z2, float x3, float y3, float z3)
{
	return div6(determinant(x1, y1, z1, x2, y2, z2, x3, y3, z3));
}

// gcm/system/quick_math.cpp:13-23
GCM_HD float tri_area(float x1, float y1, float z1, float x2, float y2, float z2)
{
	float a2 = x1*x1 + y1*y1 + z1*z1;
	float b2 = x2*x2 + y2*y2 + z2*z2;
	float ab = x1*x2 + y1*y2 + z1*z2;
	float diff = a2*b2 - ab*ab;
	if(diff < 0)
		return 0;
	return sqrtf(diff) / 2;
}

struct Vec3 { float x, y, z; };

// Signed volume of the tetrahedron on (v1-v0, v2-v0, v3-v0); the `Vol` of
// gcm/meshtypes/TetrMesh_1stOrder.cpp:1330-1340 (after fabs) and :1766-1776 (signed).
GCM_HD float tet_signed_volume(const Vec3& v0, const Vec3& v1, const Vec3& v2, const Vec3& v3)
{
	return tetr_volume(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z,
	                   v2.x - v0.x, v2.y - v0.y, v2.z - v0.z,
	                   v3.x - v0.x, v3.y - v0.y, v3.z - v0.z);
}

// gcm/meshtypes/TetrMesh_1stOrder.cpp:951-1029.  Returns < 0 when the reference would throw
// (zero volume or a zero-area face).
GCM_HD float tetr_h(const Vec3& v0, const Vec3& v1, const Vec3& v2, const Vec3& v3)
{
	float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;
	float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
	float cx = v3.x - v0.x, cy = v3.y - v0.y, cz = v3.z - v0.z;
	float vol = tetr_volume(ax, ay, az, bx, by, bz, cx, cy, cz);
	float area[4];
	area[0] = tri_area(ax, ay, az, bx, by, bz);
	area[1] = tri_area(ax, ay, az, cx, cy, cz);
	area[2] = tri_area(bx, by, bz, cx, cy, cz);
	area[3] = tri_area(v2.x - v1.x, v2.y - v1.y, v2.z - v1.z, v3.x - v1.x, v3.y - v1.y, v3.z - v1.z);
	if(vol == 0) return -1.f;
	for(int j = 0; j < 4; j++)
		if(area[j] == 0) return -1.f;
	float min_h = fabsf(3*vol/area[0]);
	for(int j = 1; j < 4; j++) {
		float h = fabsf(3*vol/area[j]);
		if(h < min_h) min_h = h;
	}
	return min_h;
}

// The four sub-volumes of point P against the faces of (v0,v1,v2,v3), in the vertex orderings
// of gcm/meshtypes/TetrMesh_1stOrder.cpp:1348-1394 (point_in_tetr) == :1780-1826 (interpolate).
GCM_HD void sub_volumes(const Vec3& v0, const Vec3& v1, const Vec3& v2, const Vec3& v3,
                        float x, float y, float z, float s[4])
{
	float a0x = v0.x - x, a0y = v0.y - y, a0z = v0.z - z;
	float a1x = v1.x - x, a1y = v1.y - y, a1z = v1.z - z;
	float a2x = v2.x - x, a2y = v2.y - y, a2z = v2.z - z;
	float a3x = v3.x - x, a3y = v3.y - y, a3z = v3.z - z;
	s[0] = tetr_volume(a1x, a1y, a1z, a2x, a2y, a2z, a3x, a3y, a3z);
	s[1] = tetr_volume(a0x, a0y, a0z, a2x, a2y, a2z, a3x, a3y, a3z);
	s[2] = tetr_volume(a1x, a1y, a1z, a0x, a0y, a0z, a3x, a3y, a3z);
	s[3] = tetr_volume(a1x, a1y, a1z, a2x, a2y, a2z, a0x, a0y, a0z);
}

// gcm/meshtypes/TetrMesh_1stOrder.cpp:1295-1402 with the per-tet statics hoisted:
// min_h = tetr_h(tet) and vol = signed Vol(tet) are functions of the (static) coordinates only.
// base = coordinates of the base node, (dx,dy,dz) the raw displacement, delta = sqrtf(dx^2+dy^2+dz^2)
// (the same for every candidate, so the caller computes it once).
GCM_HD bool point_in_tetr(const Vec3& base, float _dx, float _dy, float _dz, float delta,
                          const Vec3& v0, const Vec3& v1, const Vec3& v2, const Vec3& v3,
                          float min_h, float vol)
{
	float dx, dy, dz;
	if( (delta > 0) && ((double)delta < (double)min_h * 0.99) ) {
		dx = (float)div_z( (double)(_dx * min_h) * 0.99, (double)delta );
		dy = (float)div_z( (double)(_dy * min_h) * 0.99, (double)delta );
		dz = (float)div_z( (double)(_dz * min_h) * 0.99, (double)delta );
	} else {
		dx = _dx; dy = _dy; dz = _dz;
	}
	float Vol = fabsf(vol);
	float x = base.x + dx;
	float y = base.y + dy;
	float z = base.z + dz;
	float s[4];
	sub_volumes(v0, v1, v2, v3, x, y, z, s);
	return (double)(fabsf(s[0]) + fabsf(s[1]) + fabsf(s[2]) + fabsf(s[3])) < (double)Vol * 1.001;
}

// Barycentric factors of gcm/meshtypes/TetrMesh_1stOrder.cpp:1764-1857.  (x,y,z) = foot
// coordinates, vol = signed Vol(tet).  Returns false when the reference throws (sum >= 1.25).
GCM_HD bool interp_factors(const Vec3& v0, const Vec3& v1, const Vec3& v2, const Vec3& v3,
                           float x, float y, float z, float vol, float f[4])
{
	float s[4];
	sub_volumes(v0, v1, v2, v3, x, y, z, s);
	// fabs(s / Vol); an exactly zero sub-volume (the foot lies in a face plane of the owner: half
	// of them on axis-aligned meshes) needs no division: |+-0 / Vol| = +0 for any Vol != 0
	#pragma unroll
	for(int i = 0; i < 4; i++) {
		if(s[i] == 0.0f && fabsf(vol) > 0.0f) f[i] = 0.0f;
		else {
#if defined(__CUDA_ARCH__)
			f[i] = fabsf(__fdiv_rn(s[i], vol));   // the numerator is not zero on this branch (or Vol is degenerate)
#else
			f[i] = fabsf(s[i] / vol);
#endif
		}
	}
	float sum = f[0] + f[1] + f[2] + f[3];
	if((double)sum > 1.0) {
		if((double)sum < 1.25) {
			for(int i = 0; i < 4; i++)
				f[i] = (float)( (double)f[i] * 0.995 / (double)sum );
		} else {
			return false;
		}
	}
	return true;
}

// One interpolated quantity, gcm/meshtypes/TetrMesh_1stOrder.cpp:1864-1877.
GCM_HD float interp4(float a0, float a1, float a2, float a3, const float f[4])
{
	return (a0*f[0] + a1*f[1] + a2*f[2] + a3*f[3]);
}

// ---- elastic eigensystem -----------------------------------------------------------------
// gcm/datatypes/ElasticMatrix3D.cpp:522-530
GCM_HD void create_matrix_N(const float n[3][3], int i, int j, float res[6])
{
	res[0] = (n[i][0]*n[j][0] + n[j][0]*n[i][0])/2;
	res[1] = (n[i][0]*n[j][1] + n[j][0]*n[i][1])/2;
	res[2] = (n[i][0]*n[j][2] + n[j][0]*n[i][2])/2;
	res[3] = (n[i][1]*n[j][1] + n[j][1]*n[i][1])/2;
	res[4] = (n[i][1]*n[j][2] + n[j][1]*n[i][2])/2;
	res[5] = (n[i][2]*n[j][2] + n[j][2]*n[i][2])/2;
}

// Direction triple (n0 = q/|q|, n1, n2) of gcm/datatypes/ElasticMatrix3D.cpp:284-339.
GCM_HD void elastic_axes(float qjx, float qjy, float qjz, float n[3][3], float* l_out)
{
	float l = sqrtf(qjx*qjx + qjy*qjy + qjz*qjz);
	n[0][0] = qjx/l;
	n[0][1] = qjy/l;
	n[0][2] = qjz/l;
	if(fabsf(n[0][0]) <= fabsf(n[0][1])) {
		if(fabsf(n[0][0]) <= fabsf(n[0][2])) {
			n[1][0] = 0;         n[1][1] = -n[0][2]; n[1][2] = n[0][1];
		} else {
			n[1][0] = -n[0][1];  n[1][1] = n[0][0];  n[1][2] = 0;
		}
	} else {
		if(fabsf(n[0][1]) <= fabsf(n[0][2])) {
			n[1][0] = -n[0][2];  n[1][1] = 0;        n[1][2] = n[0][0];
		} else {
			n[1][0] = -n[0][1];  n[1][1] = n[0][0];  n[1][2] = 0;
		}
	}
	float dtmp = sqrtf(n[1][0]*n[1][0] + n[1][1]*n[1][1] + n[1][2]*n[1][2]);
	n[1][0] /= dtmp;
	n[1][1] /= dtmp;
	n[1][2] /= dtmp;
	n[2][0] = n[0][1]*n[1][2] - n[0][2]*n[1][1];
	n[2][1] = n[0][2]*n[1][0] - n[0][0]*n[1][2];
	n[2][2] = n[0][0]*n[1][1] - n[0][1]*n[1][0];
	*l_out = l;
}

// Lambda (diagonal), Omega (U) and Omega^-1 (U1) of
// gcm/datatypes/ElasticMatrix3D.cpp:243-519 (CreateGeneralizedMatrix); A is never read by
// the step and is not built.  U/U1 are row-major 9x9.  Returns a StepError.
GCM_HD int elastic_matrix(float la, float mu, float ro, float qjx, float qjy, float qjz,
                          float L[9], float* U, float* U1)
{
	if((la <= 0) || (mu <= 0) || (ro <= 0))
		return ERR_BAD_RHEOLOGY;
	if( qjx*qjx + qjy*qjy + qjz*qjz <= 0 )
		return ERR_BAD_Q;
	float n[3][3];
	float l;
	elastic_axes(qjx, qjy, qjz, n, &l);

	float c1 = sqrtf((la+2*mu)/ro);
	float c2 = sqrtf(mu/ro);
	float c3 = sqrtf(la*la/(ro*(la+2*mu)));

	L[0] = c1 * l;  L[1] = -c1 * l;
	L[2] = c2 * l;  L[3] = -c2 * l;
	L[4] = c2 * l;  L[5] = -c2 * l;
	L[6] = 0; L[7] = 0; L[8] = 0;

	float I[6] = {1, 0, 0, 1, 0, 1};
	float N00[6], N01[6], N02[6], N11[6], N12[6], N22[6];
	create_matrix_N(n, 0, 0, N00);
	create_matrix_N(n, 0, 1, N01);
	create_matrix_N(n, 0, 2, N02);
	create_matrix_N(n, 1, 1, N11);
	create_matrix_N(n, 1, 2, N12);
	create_matrix_N(n, 2, 2, N22);

	if(U1) {
#define GCM_U1(i,j) U1[(i)*9+(j)]
		for(int k = 0; k < 3; k++) {
			GCM_U1(k,0) = n[0][k]; GCM_U1(k,1) = n[0][k];
			GCM_U1(k,2) = n[1][k]; GCM_U1(k,3) = n[1][k];
			GCM_U1(k,4) = n[2][k]; GCM_U1(k,5) = n[2][k];
			GCM_U1(k,6) = 0; GCM_U1(k,7) = 0; GCM_U1(k,8) = 0;
		}
		for(int i = 0; i < 6; i++) {
			GCM_U1(3+i,0) = -ro*((c1-c3)*N00[i] + c3*I[i]);
			GCM_U1(3+i,1) = ro*((c1-c3)*N00[i] + c3*I[i]);
			GCM_U1(3+i,2) = -2*ro*c2*N01[i];
			GCM_U1(3+i,3) = 2*ro*c2*N01[i];
			GCM_U1(3+i,4) = -2*ro*c2*N02[i];
			GCM_U1(3+i,5) = 2*ro*c2*N02[i];
			GCM_U1(3+i,6) = 4*N12[i];
			GCM_U1(3+i,7) = N11[i] - N22[i];
			GCM_U1(3+i,8) = I[i] - N00[i];
		}
		for(int i = 0; i < 81; i++)
			U1[i] /= 2;
#undef GCM_U1
	}
	if(U) {
#define GCM_U(i,j) U[(i)*9+(j)]
		const float w[6] = {1, 2, 2, 1, 2, 1};
		for(int k = 0; k < 3; k++) {
			GCM_U(0,k) = n[0][k]; GCM_U(1,k) = n[0][k];
			GCM_U(2,k) = n[1][k]; GCM_U(3,k) = n[1][k];
			GCM_U(4,k) = n[2][k]; GCM_U(5,k) = n[2][k];
			GCM_U(6,k) = 0; GCM_U(7,k) = 0; GCM_U(8,k) = 0;
		}
		// the reference writes  -N/(c*ro)  for w==1 and  -2*N/(c*ro)  for w==2; 1*N == N exactly
		for(int i = 0; i < 6; i++) {
			GCM_U(0,3+i) = div_z(-w[i]*N00[i], c1*ro);
			GCM_U(1,3+i) = div_z(w[i]*N00[i], c1*ro);
			GCM_U(2,3+i) = div_z(-w[i]*N01[i], c2*ro);
			GCM_U(3,3+i) = div_z(w[i]*N01[i], c2*ro);
			GCM_U(4,3+i) = div_z(-w[i]*N02[i], c2*ro);
			GCM_U(5,3+i) = div_z(w[i]*N02[i], c2*ro);
			GCM_U(6,3+i) = w[i]*N12[i];
			GCM_U(7,3+i) = w[i]*(N11[i]-N22[i]);
			GCM_U(8,3+i) = w[i]*(N11[i]+N22[i]-2*la*N00[i]/(la+2*mu));
		}
#undef GCM_U
	}
	return ERR_NONE;
}

// Row-wise access to the same eigensystem, for kernels that spread the rows of one node over the
// lanes of a warp (stage_border_coop_kernel).  elastic_dirs() = everything elastic_matrix() derives
// from (la, mu, ro, q) before filling the matrices; elastic_U_row / elastic_U1_row evaluate one row
// with the very expressions elastic_matrix() uses (tests/test_host_logic.py compares them bitwise).
struct ElasticDirs {
	float n[3][3];
	float l, c1, c2, c3;
	float N00[6], N01[6], N02[6], N11[6], N12[6], N22[6];
};

GCM_HD int elastic_dirs(float la, float mu, float ro, float qjx, float qjy, float qjz, ElasticDirs& d)
{
	if((la <= 0) || (mu <= 0) || (ro <= 0))
		return ERR_BAD_RHEOLOGY;
	if( qjx*qjx + qjy*qjy + qjz*qjz <= 0 )
		return ERR_BAD_Q;
	elastic_axes(qjx, qjy, qjz, d.n, &d.l);
	d.c1 = sqrtf((la+2*mu)/ro);
	d.c2 = sqrtf(mu/ro);
	d.c3 = sqrtf(la*la/(ro*(la+2*mu)));
	create_matrix_N(d.n, 0, 0, d.N00);
	create_matrix_N(d.n, 0, 1, d.N01);
	create_matrix_N(d.n, 0, 2, d.N02);
	create_matrix_N(d.n, 1, 1, d.N11);
	create_matrix_N(d.n, 1, 2, d.N12);
	create_matrix_N(d.n, 2, 2, d.N22);
	return ERR_NONE;
}

// Lambda(i,i)
GCM_HD float elastic_L(const ElasticDirs& d, int i)
{
	switch(i) {
	case 0: return d.c1 * d.l;
	case 1: return -d.c1 * d.l;
	case 2: case 4: return d.c2 * d.l;
	case 3: case 5: return -d.c2 * d.l;
	default: return 0;
	}
}

// Row i of Omega (U)
GCM_HD void elastic_U_row(const ElasticDirs& d, float la, float mu, float ro, int i, float row[9])
{
	const float w[6] = {1, 2, 2, 1, 2, 1};
	const float c1 = d.c1, c2 = d.c2;
	const int a = i < 2 ? 0 : (i < 4 ? 1 : (i < 6 ? 2 : -1));
	for(int k = 0; k < 3; k++) row[k] = a >= 0 ? d.n[a][k] : 0;
	for(int q = 0; q < 6; q++) {
		float v;
		switch(i) {
		case 0: v = div_z(-w[q]*d.N00[q], c1*ro); break;
		case 1: v = div_z(w[q]*d.N00[q], c1*ro); break;
		case 2: v = div_z(-w[q]*d.N01[q], c2*ro); break;
		case 3: v = div_z(w[q]*d.N01[q], c2*ro); break;
		case 4: v = div_z(-w[q]*d.N02[q], c2*ro); break;
		case 5: v = div_z(w[q]*d.N02[q], c2*ro); break;
		case 6: v = w[q]*d.N12[q]; break;
		case 7: v = w[q]*(d.N11[q]-d.N22[q]); break;
		default: v = w[q]*(d.N11[q]+d.N22[q]-2*la*d.N00[q]/(la+2*mu)); break;
		}
		row[3+q] = v;
	}
}

// Row r of Omega^-1 (U1)
GCM_HD void elastic_U1_row(const ElasticDirs& d, float ro, int r, float row[9])
{
	const float c1 = d.c1, c2 = d.c2, c3 = d.c3;
	if(r < 3) {
		row[0] = d.n[0][r]; row[1] = d.n[0][r];
		row[2] = d.n[1][r]; row[3] = d.n[1][r];
		row[4] = d.n[2][r]; row[5] = d.n[2][r];
		row[6] = 0; row[7] = 0; row[8] = 0;
	} else {
		const int i = r - 3;
		const float I[6] = {1, 0, 0, 1, 0, 1};
		row[0] = -ro*((c1-c3)*d.N00[i] + c3*I[i]);
		row[1] = ro*((c1-c3)*d.N00[i] + c3*I[i]);
		row[2] = -2*ro*c2*d.N01[i];
		row[3] = 2*ro*c2*d.N01[i];
		row[4] = -2*ro*c2*d.N02[i];
		row[5] = 2*ro*c2*d.N02[i];
		row[6] = 4*d.N12[i];
		row[7] = d.N11[i] - d.N22[i];
		row[8] = I[i] - d.N00[i];
	}
	for(int c = 0; c < 9; c++) row[c] /= 2;
}

// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:51-71 (drop_deviator).  In place: s = values[3..8]
// (sxx,sxy,sxz,syy,syz,szz); writes only when yielding, exactly as the reference.
GCM_HD bool drop_deviator(float v[9], float yield_limit)
{
	float J2 = sqrtf( ( (v[3] - v[6]) * (v[3] - v[6])
			+ (v[6] - v[8]) * (v[6] - v[8])
			+ (v[3] - v[8]) * (v[3] - v[8])
			+ 6 * ( (v[4]) * (v[4]) + (v[5]) * (v[5])
				+ (v[7]) * (v[7])) ) / 6 );
	float p = (v[3] + v[6] + v[8]) / 3;
	if (yield_limit < J2)
	{
		float n3 = ((v[3] - p) * yield_limit / J2) + p;
		float n4 = v[4] * yield_limit / J2;
		float n5 = v[5] * yield_limit / J2;
		float n6 = ((v[6] - p) * yield_limit / J2) + p;
		float n7 = v[7] * yield_limit / J2;
		float n8 = ((v[8] - p) * yield_limit / J2) + p;
		v[3] = n3; v[4] = n4; v[5] = n5; v[6] = n6; v[7] = n7; v[8] = n8;
		return true;
	}
	return false;
}

// Stage 3 folded into the writers of stage 2 (TetrMethod_Plastic_1stOrder.cpp:192-199 on the values
// stage 2 just produced: the NaN guard of the stage-3 call, then drop_deviator).  Returns false on NaN/Inf.
GCM_HD bool plastic_epilogue(float v[9], float yield_limit)
{
	bool bad = false;
	for(int k = 0; k < 9; k++) bad |= !(fabsf(v[k]) <= 3.4028235e38f);
	if(bad) return false;
	drop_deviator(v, yield_limit);
	return true;
}

// TetrMesh_1stOrder::calc_destruction_criterias (gcm/meshtypes/TetrMesh_1stOrder.cpp:1957-1995), the part that
// depends on the current values only: c = (max_compression, max_tension, max_shear, max_deviator).
GCM_HD void destruction_now(const float v[9], float c[4])
{
	float ten = 0, com = 0, sh = 0;
	if(v[3] > ten) ten = v[3];
	if(v[3] < com) com = v[3];
	if(v[6] > ten) ten = v[6];
	if(v[6] < com) com = v[6];
	if(v[8] > ten) ten = v[8];
	if(v[8] < com) com = v[8];
	com = fabsf(com);
	if(fabsf(v[4]) > sh) sh = fabsf(v[4]);
	if(fabsf(v[5]) > sh) sh = fabsf(v[5]);
	if(fabsf(v[7]) > sh) sh = fabsf(v[7]);
	const float dev = sqrtf( ( (v[3] - v[6]) * (v[3] - v[6])
		+ (v[6] - v[8]) * (v[6] - v[8])
		+ (v[3] - v[8]) * (v[3] - v[8])
		+ 6 * ( (v[4]) * (v[4]) + (v[5]) * (v[5])
		+ (v[7]) * (v[7])) ) / 6 );
	c[0] = com; c[1] = ten; c[2] = sh; c[3] = dev;
}

// ---- dense LU in double (GSL gsl_linalg_LU_decomp / gsl_linalg_LU_solve semantics) --------
// GSL is a third-party dependency of the reference, absent from /root/reference and unpinned
// (gcm/CMakeLists.txt:88).  Published GSL 1.x algorithm: Gaussian elimination with partial
// pivoting, pivot = first row of maximal |a_ij| (strict >), multipliers stored in place,
// rank-1 update; solve = permute, unit-lower forward substitution, upper back substitution.
// A is row-major NxN and is destroyed; b is overwritten by x.
template<int N>
GCM_HD void lu_solve(double* A, double* b)
{
	int perm[N];
	for(int i = 0; i < N; i++) perm[i] = i;
	for(int j = 0; j + 1 < N; j++) {
		double mx = fabs(A[j*N+j]);
		int ip = j;
		for(int i = j + 1; i < N; i++) {
			double a = fabs(A[i*N+j]);
			if(a > mx) { mx = a; ip = i; }
		}
		if(ip != j) {
			for(int k = 0; k < N; k++) { double t = A[j*N+k]; A[j*N+k] = A[ip*N+k]; A[ip*N+k] = t; }
			int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t;
		}
		double ajj = A[j*N+j];
		if(ajj != 0.0) {
			// The pivot row and the row being eliminated pass through registers, all N entries with compile-time
			// indices (those at k <= j are carried along and not stored): the N loads of a row are in flight
			// together instead of one dependent load-update-store per entry.  Each stored entry is the same
			// product and difference as before.
			double prow[N];
			#pragma unroll
			for(int k = 0; k < N; k++) prow[k] = A[j*N+k];
			for(int i = j + 1; i < N; i++) {
				double row[N];
				#pragma unroll
				for(int k = 0; k < N; k++) row[k] = A[i*N+k];
				double aij = div_z(A[i*N+j], ajj);
				A[i*N+j] = aij;
				#pragma unroll
				for(int k = 0; k < N; k++)
					if(k > j) A[i*N+k] = row[k] - aij * prow[k];
			}
		}
	}
	double x[N];
	for(int i = 0; i < N; i++) x[i] = b[perm[i]];
	for(int i = 0; i < N; i++) {
		double t = x[i];
		for(int j = 0; j < i; j++) t -= A[i*N+j] * x[j];
		x[i] = t;
	}
	for(int i = N - 1; i >= 0; i--) {
		double t = x[i];
		for(int j = i + 1; j < N; j++) t -= A[i*N+j] * x[j];
		x[i] = div_z(t, A[i*N+i]);
	}
	for(int i = 0; i < N; i++) b[i] = x[i];
}

// gcm/system/quick_math.cpp:85-115 (create_local_basis)
GCM_HD void create_local_basis(const float n[3], float n1[3], float n2[3])
{
	if(fabsf(n[0]) <= fabsf(n[1])) {
		if(fabsf(n[0]) <= fabsf(n[2])) { n1[0] = 0; n1[1] = -n[2]; n1[2] = n[1]; }
		else { n1[0] = -n[1]; n1[1] = n[0]; n1[2] = 0; }
	} else {
		if(fabsf(n[1]) <= fabsf(n[2])) { n1[0] = -n[2]; n1[1] = 0; n1[2] = n[0]; }
		else { n1[0] = -n[1]; n1[1] = n[0]; n1[2] = 0; }
	}
	float dtmp = sqrtf(n1[0]*n1[0] + n1[1]*n1[1] + n1[2]*n1[2]);
	n1[0] /= dtmp;
	n1[1] /= dtmp;
	n1[2] /= dtmp;
	// vector_product, quick_math.cpp:38-43
	n2[0] = n[1] * n1[2] - n1[1] * n[2];
	n2[1] = n1[0] * n[2] - n[0] * n1[2];
	n2[2] = n[0] * n1[1] - n1[0] * n[1];
}

// create_rotation_matrix_with_normal, gcm/methods/TetrMethod_Plastic_1stOrder.cpp:843-901:
// local basis of a border node = (normal, a perpendicular, their cross product) rotated about
// the normal by teta.  ct/st = cos(teta)/sin(teta) as the host libm gives them for the float
// angle (the reference calls cos/sin on a float; 360 distinct angles, tabulated by the host).
GCM_HD void rotation_basis_with_normal(float x, float y, float z, float ct, float st, float out[9])
{
	float a[3][3];
	a[0][0] = x; a[0][1] = y; a[0][2] = z;
	if( (fabsf(x) >= fabsf(y)) && (fabsf(x) >= fabsf(z)) ) { a[1][0] = y; a[1][1] = -x; a[1][2] = 0; }
	else if( (fabsf(y) >= fabsf(x)) && (fabsf(y) >= fabsf(z)) ) { a[1][0] = y; a[1][1] = -x; a[1][2] = 0; }
	else { a[1][0] = 0; a[1][1] = -z; a[1][2] = y; }
	float modul = sqrtf( a[1][0] * a[1][0] + a[1][1] * a[1][1] + a[1][2] * a[1][2] );
	a[1][0] /= modul; a[1][1] /= modul; a[1][2] /= modul;
	a[2][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
	a[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
	a[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
	float rot[3][3];
	rot[0][0] = ct + x * x * (1 - ct);
	rot[0][1] = y * x * (1 - ct) + z * st;
	rot[0][2] = z * x * (1 - ct) - y * st;
	rot[1][0] = y * x * (1 - ct) - z * st;
	rot[1][1] = ct + y * y * (1 - ct);
	rot[1][2] = y * z * (1 - ct) + x * st;
	rot[2][0] = x * z * (1 - ct) + y * st;
	rot[2][1] = y * z * (1 - ct) - x * st;
	rot[2][2] = ct + z * z * (1 - ct);
	for(int r = 0; r < 3; r++)
		for(int c = 0; c < 3; c++) {
			float v = 0;
			for(int k = 0; k < 3; k++)
				v += rot[k][c] * a[r][k];
			out[3*r + c] = v;
		}
}

} // namespace gcm
