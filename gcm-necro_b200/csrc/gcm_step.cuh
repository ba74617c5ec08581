// gcm_step.cuh -- one node, one stage of the grid-characteristic method, as straight-line
// __host__ __device__ code over an SoA mesh view.
//
// Follows gcm/methods/TetrMethod_Plastic_1stOrder.cpp:190-423 (do_next_part_step),
// :78-111 (prepare_node), :430-681 (find_nodes_on_previous_time_layer) and
// gcm/meshtypes/TetrMesh_1stOrder.cpp:1468-1602 (find_owner_tetr), :1764-1880 (interpolate).
// The kernels in gcm_kernels.cu call these per thread; tests/ compile the same header with
// g++ to single-step the logic on the CPU (debug aid only -- not part of the product library).
#pragma once
#include "gcm_math.cuh"

namespace gcm {

// node_flags bits, gcm/datatypes/Node.h:4-27
enum NodeFlags { NF_IN_CONTACT = 1, NF_LOCAL = 2, NF_USED = 4, NF_BORDER = 8, NF_GCM = 16, NF_SPH = 32 };

// Border calculators (gcm/methods/border/*.cpp) as device-side condition ids.
enum BorderCalcType { BC_FREE = 0, BC_FIXED = 1, BC_EXT_FORCE = 2, BC_EXT_VELOCITY = 3, BC_EXT_VALUES = 4 };
// Contact calculators (gcm/methods/contact/*.cpp).
enum ContactCalcType { CC_ADHESION = 0, CC_FRICTION = 1 };

// One BorderCondition = calculator + its parameters + the pulse form evaluated for this step
// (gcm/system/BorderCondition.cpp:3-6, gcm/system/forms/StepPulseForm.cpp:5-14).
struct BorderCond {
	int type;            // BorderCalcType
	float p_normal;      // normal_stress / normal_v
	float p_tangential;  // tangential_stress / tangential_v
	float dir[3];        // tangential_direction, already normalised (set_parameters)
	int vars_index[3];   // ExternalValuesCalculator
	float vars_values[3];
	float scale;         // form->calcMagnitudeNorm(t) for the current step
	int active;          // form->isActive(t); inactive -> default condition (check_stresses)
};

#define GCM_MAX_BORDER_CONDS 16

// Read-only view of one zone mesh in device (or, in the test emulation, host) memory.
struct MeshView {
	int n_nodes, n_tets;
	size_t stride;           // plane stride of the material planes (floats)
	const float* u;          // current layer: node records as three PLANES of 16-byte quads,
	                         //   plane 0 = vx,vy,vz,sxx | plane 1 = sxy,sxz,syy,syz | plane 2 = szz,x,y,z;
	                         //   quad q of node i at float4 index q * rs + i (rec_get / rec_quad below)
	size_t rs;               // quads per plane
	const float* mat;        // [4][stride]  la, mu, rho, yield_limit
	const int* flags;        // [n_nodes]
	const int* n2t_off;      // [n_nodes+1]  node -> incident tets (ascending ids), CSR
	const int* n2t;
	const int* n2x;          // [len(n2t)][4] or NULL: CSR entry -> other three vertices (tet order), tet | k << 30
	const int* n2x_scan;     // the same table (or NULL) for scan_ring_both_senses, whatever n2x says
	const int* tets;         // [n_tets][4]
	const float* tet_hv;     // [n_tets][2]  (tetr_h, signed volume) -- static geometry
	const int* border_slot;  // [n_nodes]    index into the border arrays or -1
	const float* basis;      // [n_border][9] local basis ksi[3][3] of this step (border nodes)
	const int* cond_id;      // [n_border]   index into conds[] or -1 = default
	const int* contact;      // [n_border][6] axis_plus[3], axis_minus[3] (virt node ids) or NULL
	int* ring_ws;            // NULL, or GCM_RING_WS ints per thread of the launch: workspace of find_owner_general
	// Thresholds of the conservative owner pre-filter (barycentric units, both negative): a candidate
	// certainly fails below pf_fail and certainly passes above pf_pass.  They depend on how much the
	// reference's own float evaluation of the tested point is perturbed, i.e. on max|coordinate| / min h
	// (gcm_capi.cu: prefilter_thresholds); exact_only != 0 switches the pre-filter off altogether
	// (huge coordinates, and the one-time build of the inner cache, where speed does not matter).
	float pf_fail, pf_pass;
	int exact_only;
	float tau;
	float time;
	float min_h;
	BorderCond conds[GCM_MAX_BORDER_CONDS];
	int n_conds;
};

// Node record: the 9 evolving values and the (static) coordinates of a node are three float4 quads.
// The quads are stored plane by plane (all first quads, all second quads, all third quads): a warp
// working on 32 consecutive nodes -- or gathering 32 consecutive neighbours, the common case in a
// mesh numbered with any locality -- then moves 512 contiguous bytes per LDG.128 / STG.128 instead of
// 16-byte pieces at a 48-byte stride (ncu: 12 L1 tag requests per instruction against 4).
#define GCM_REC 12
GCM_HD size_t rec_index(size_t rs, int i, int k) { return (((size_t)(k >> 2)) * rs + (size_t)i) * 4 + (size_t)(k & 3); }
GCM_HD float rec_get(const float* u, size_t rs, int i, int k) { return u[rec_index(rs, i, k)]; }
GCM_HD void rec_set(float* u, size_t rs, int i, int k, float v) { u[rec_index(rs, i, k)] = v; }
#if defined(__CUDACC__)
__device__ __forceinline__ float4 rec_quad(const float* u, size_t rs, int i, int q) { return reinterpret_cast<const float4*>(u)[(size_t)q * rs + (size_t)i]; }
__device__ __forceinline__ float4 rec_quad_ldg(const float* u, size_t rs, int i, int q) { return __ldg(reinterpret_cast<const float4*>(u) + (size_t)q * rs + (size_t)i); }
__device__ __forceinline__ void rec_quad_store(float* u, size_t rs, int i, int q, float4 v) { reinterpret_cast<float4*>(u)[(size_t)q * rs + (size_t)i] = v; }
#endif
// all 12 floats of record i
GCM_HD void rec_load12(const float* u, size_t rs, int i, float r[12])
{
#if defined(__CUDA_ARCH__)
	#pragma unroll
	for(int w = 0; w < 3; w++) {
		const float4 a = rec_quad(u, rs, i, w);
		r[4*w] = a.x; r[4*w+1] = a.y; r[4*w+2] = a.z; r[4*w+3] = a.w;
	}
#else
	for(int k = 0; k < 12; k++) r[k] = rec_get(u, rs, i, k);
#endif
}

GCM_HD Vec3 load_xyz(const MeshView& m, int i)
{
	Vec3 v;
#if defined(__CUDA_ARCH__)
	const float4 q = rec_quad(m.u, m.rs, i, 2);   // (szz, x, y, z)
	v.x = q.y; v.y = q.z; v.z = q.w;
#else
	v.x = rec_get(m.u, m.rs, i, 9); v.y = rec_get(m.u, m.rs, i, 10); v.z = rec_get(m.u, m.rs, i, 11);
#endif
	return v;
}

struct TetGeom { int v[4]; Vec3 p[4]; float h, vol; };

GCM_HD void load_tet_row(const MeshView& m, int t, int v[4])
{
#if defined(__CUDA_ARCH__)
	const int4 q = reinterpret_cast<const int4*>(m.tets)[t];
	v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#else
	for(int k = 0; k < 4; k++) v[k] = m.tets[4*(size_t)t + k];
#endif
}

GCM_HD void load_tet_hv(const MeshView& m, int t, float& h, float& vol)
{
#if defined(__CUDA_ARCH__)
	const float2 hv = reinterpret_cast<const float2*>(m.tet_hv)[t];
	h = hv.x; vol = hv.y;
#else
	h = m.tet_hv[2*(size_t)t]; vol = m.tet_hv[2*(size_t)t + 1];
#endif
}

GCM_HD void load_tet(const MeshView& m, int t, TetGeom& g)
{
#if defined(__CUDA_ARCH__)
	int4 q = reinterpret_cast<const int4*>(m.tets)[t];
	g.v[0] = q.x; g.v[1] = q.y; g.v[2] = q.z; g.v[3] = q.w;
	float2 hv = reinterpret_cast<const float2*>(m.tet_hv)[t];
	g.h = hv.x; g.vol = hv.y;
#else
	for(int k = 0; k < 4; k++) g.v[k] = m.tets[4*(size_t)t + k];
	g.h = m.tet_hv[2*(size_t)t]; g.vol = m.tet_hv[2*(size_t)t + 1];
#endif
	for(int k = 0; k < 4; k++) g.p[k] = load_xyz(m, g.v[k]);
}

// Conservative pre-filter of point_in_tetr for a candidate tet that has the base node X as a
// vertex (every first-ring candidate has).  The predicate tests P = X + d * 0.99 h / |d| (the
// rescale of TetrMesh_1stOrder.cpp:1304-1307) with  sum|sub-volumes| < 1.001 Vol,  i.e. the sum
// of the negative barycentric coordinates of P must stay above -5e-4.  The coordinates follow
// from Cramer's rule on the edges X->A, X->B, X->C in plain float (error ~1e-5 for tets with
// 6|Vol| > 0.02 * longest edge^3, flagged by a positive stored |Vol|):
//   -1  some coordinate < -1e-3        -> the predicate certainly fails
//   +1  all four coordinates > -1e-4   -> it certainly passes
//    0  anything else (or tet / displacement outside the filter's assumptions) -> evaluate it
GCM_HD int owner_prefilter(const TetGeom& g, int base, const Vec3& X, float dx, float dy, float dz, float delta,
                           float pf_fail, float pf_pass)
{
	if(!(g.vol > 0) || !(delta > 0) || !((double)delta < (double)g.h * 0.99)) return 0;
	const int k = (g.v[0] == base) ? 0 : (g.v[1] == base) ? 1 : (g.v[2] == base) ? 2 : (g.v[3] == base) ? 3 : -1;
	if(k < 0) return 0;
	const Vec3 A = g.p[k == 0 ? 1 : 0], B = g.p[k <= 1 ? 2 : 1], C = g.p[k == 3 ? 2 : 3];
	const float eax = A.x - X.x, eay = A.y - X.y, eaz = A.z - X.z;
	const float ebx = B.x - X.x, eby = B.y - X.y, ebz = B.z - X.z;
	const float ecx = C.x - X.x, ecy = C.y - X.y, ecz = C.z - X.z;
	const float bcx = eby * ecz - ebz * ecy, bcy = ebz * ecx - ebx * ecz, bcz = ebx * ecy - eby * ecx;
	const float cax = ecy * eaz - ecz * eay, cay = ecz * eax - ecx * eaz, caz = ecx * eay - ecy * eax;
	const float abx = eay * ebz - eaz * eby, aby = eaz * ebx - eax * ebz, abz = eax * eby - eay * ebx;
	const float D = eax * bcx + eay * bcy + eaz * bcz;
	const float sD = 0.99f * g.h / delta * D;
	// lambda_j = n_j / D  ->  compare n_j * D with multiples of D^2
	const float la = sD * (dx * bcx + dy * bcy + dz * bcz);
	const float lb = sD * (dx * cax + dy * cay + dz * caz);
	const float lc = sD * (dx * abx + dy * aby + dz * abz);
	const float D2 = D * D;
	const float lx = D2 - (la + lb + lc);
	const float mn = fminf(fminf(la, lb), fminf(lc, lx));
	return mn < pf_fail * D2 ? -1 : (mn > pf_pass * D2 ? 1 : 0);
}

// First ring of gcm/meshtypes/TetrMesh_1stOrder.cpp:1468-1602: the candidates are the base
// node's incident tets in stored order, the first one passing point_in_tetr wins.
//   base      = index whose `elements` list and stored coordinates are used (nodes[base_node])
//   node_pos  = node->coords (differs from the stored coordinates only for virtual nodes)
// Returns the tet id, -1 for "not found and the search would stop" (NULL), and sets *expand
// when the reference would go on to the next ring (some candidate vertex closer than |dx|).
GCM_HD int find_owner_first_ring(const MeshView& m, int base, const Vec3& node_pos,
                                 float dx, float dy, float dz, TetGeom& g, bool* expand)
{
	Vec3 X = load_xyz(m, base);
	float R2 = dx * dx + dy * dy + dz * dz;
	float delta = sqrtf(R2);
	bool inside_R = false;
	int e0 = m.n2t_off[base], e1 = m.n2t_off[base + 1];
	for(int e = e0; e < e1; e++) {
		int t;
		if(m.n2x) {
			// expanded CSR entry: the other three vertices (tet order) and the position k of `base`;
			// saves the tets[t] hop and the load of the base node's own coordinates
			const int* x = m.n2x + 4*(size_t)e;
#if defined(__CUDA_ARCH__)
			const int4 q = *reinterpret_cast<const int4*>(x);
			const int xa = q.x, xb = q.y, xc = q.z, w = q.w;
#else
			const int xa = x[0], xb = x[1], xc = x[2], w = x[3];
#endif
			const int k = (int)((unsigned)w >> 30);
			t = w & 0x3fffffff;
			const Vec3 A = load_xyz(m, xa), B = load_xyz(m, xb), C = load_xyz(m, xc);
			g.v[0] = k == 0 ? base : xa;  g.p[0] = k == 0 ? X : A;
			g.v[1] = k == 0 ? xa : (k == 1 ? base : xb);  g.p[1] = k == 0 ? A : (k == 1 ? X : B);
			g.v[2] = k <= 1 ? xb : (k == 2 ? base : xc);  g.p[2] = k <= 1 ? B : (k == 2 ? X : C);
			g.v[3] = k == 3 ? base : xc;  g.p[3] = k == 3 ? X : C;
			g.h = m.tet_hv[2*(size_t)t]; g.vol = m.tet_hv[2*(size_t)t + 1];
		} else {
			t = m.n2t[e];
			load_tet(m, t, g);
		}
		int verdict = m.exact_only ? 0 : owner_prefilter(g, base, X, dx, dy, dz, delta, m.pf_fail, m.pf_pass);
		if(verdict == 0)
			verdict = point_in_tetr(X, dx, dy, dz, delta, g.p[0], g.p[1], g.p[2], g.p[3], g.h, g.vol) ? 1 : -1;
		if(verdict > 0)
			return t;
		if(!inside_R) {
			for(int j = 0; j < 4; j++) {
				if(g.v[j] == base) break;   // the reference `break`s (not `continue`s) here, :1544-1545
				float lx = g.p[j].x, ly = g.p[j].y, lz = g.p[j].z;
				if( ( (node_pos.x - lx) * (node_pos.x - lx)
					+ (node_pos.y - ly) * (node_pos.y - ly)
					+ (node_pos.z - lz) * (node_pos.z - lz) ) < R2 )
					inside_R = true;
			}
		}
	}
	*expand = inside_R;
	return -1;
}

GCM_HD bool tet_has_vertex(const MeshView& m, int t, int v)
{
	const int* tv = m.tets + 4*(size_t)t;
	return tv[0] == v || tv[1] == v || tv[2] == v || tv[3] == v;
}

// Second ring of find_owner_tetr (:1564-1597): all tets incident to any vertex of the first
// ring, ascending, de-duplicated, minus the first ring; the first one passing wins.  Evaluated
// without materialising the list: the winner is the smallest passing tet id, so every unchecked
// candidate with an id below the best so far is tested (a tet reachable through several vertices
// is simply tested again).  Each distinct ring-1 vertex is expanded once.  *expand = some vertex
// of the second ring lies within |dx| of the node (the reference would build a third ring).
// Needed for virtual contact nodes, whose borrowed base node can be a full face edge away.
GCM_HD int find_owner_second_ring(const MeshView& m, int base, const Vec3& node_pos,
                                  float dx, float dy, float dz, TetGeom& g, bool* expand)
{
	const Vec3 X = load_xyz(m, base);
	const float R2 = dx * dx + dy * dy + dz * dz;
	const float delta = sqrtf(R2);
	const int e0 = m.n2t_off[base], e1 = m.n2t_off[base + 1];
	int best = 0x7fffffff;
	bool inside_R = false;
	// The search is a chain of dependent loads (tet id -> vertex ids -> coordinates -> predicate), ~400 candidates
	// long for a virtual contact node: the ring-1 rows are read once, and the candidates of a vertex go through in
	// groups of four whose loads are all issued before the first predicate -- same candidates, same order, same
	// arithmetic, a quarter of the round trips.
	const int R1_MAX = 32;
	int r1[R1_MAX][4];
	const int n1 = e1 - e0;
	const bool have_r1 = n1 <= R1_MAX;
	if(have_r1)
		for(int a = 0; a < n1; a++) {
			load_tet_row(m, m.n2t[e0 + a], r1[a]);
		}
	for(int a = e0; a < e1; a++) {
		const int* tv1 = have_r1 ? r1[a - e0] : m.tets + 4*(size_t)m.n2t[a];
		for(int j = 0; j < 4; j++) {
			const int v = tv1[j];
			if(v == base) continue;
			bool seen = false;
			if(have_r1) { for(int b = 0; b < a - e0 && !seen; b++) seen = r1[b][0] == v || r1[b][1] == v || r1[b][2] == v || r1[b][3] == v; }
			else { for(int b = e0; b < a && !seen; b++) seen = tet_has_vertex(m, m.n2t[b], v); }
			if(seen) continue;
			const int c1 = m.n2t_off[v + 1];
			for(int c = m.n2t_off[v]; c < c1; c += 4) {
				int tt[4], tv4[4][4];
				bool take[4];
				TetGeom q[4];
				#pragma unroll
				for(int u = 0; u < 4; u++) tt[u] = c + u < c1 ? m.n2t[c + u] : 0x7fffffff;
				#pragma unroll
				for(int u = 0; u < 4; u++) {
					take[u] = tt[u] < best;
					if(take[u]) {
						load_tet_row(m, tt[u], tv4[u]);
						take[u] = !(tv4[u][0] == base || tv4[u][1] == base || tv4[u][2] == base || tv4[u][3] == base);
					}
				}
				#pragma unroll
				for(int u = 0; u < 4; u++)
					if(take[u]) {
						for(int k = 0; k < 4; k++) { q[u].v[k] = tv4[u][k]; q[u].p[k] = load_xyz(m, tv4[u][k]); }
						load_tet_hv(m, tt[u], q[u].h, q[u].vol);
					}
				#pragma unroll
				for(int u = 0; u < 4; u++) {
					if(!take[u] || tt[u] >= best) continue;      // an earlier one of the group may have lowered `best`
					if(point_in_tetr(X, dx, dy, dz, delta, q[u].p[0], q[u].p[1], q[u].p[2], q[u].p[3], q[u].h, q[u].vol)) { best = tt[u]; continue; }
					for(int k = 0; k < 4; k++) {
						const float lx = q[u].p[k].x, ly = q[u].p[k].y, lz = q[u].p[k].z;
						if( ( (node_pos.x - lx) * (node_pos.x - lx)
							+ (node_pos.y - ly) * (node_pos.y - ly)
							+ (node_pos.z - lz) * (node_pos.z - lz) ) < R2 )
							inside_R = true;
					}
				}
			}
		}
	}
	if(best != 0x7fffffff) { load_tet(m, best, g); *expand = false; return best; }
	*expand = inside_R;
	return -1;
}

// ---- find_owner_tetr with any number of rings (:1468-1602), lists materialised -------------------
// Used when the first two rings (above, list-free) do not settle the search: virtual contact
// nodes on non-matching surface meshes, whose displacement from the borrowed base node is a face
// edge long.  `ws` = this thread's workspace of GCM_RING_WS ints:
//   checked  [0, 3C)   every tet of the rings tested so far, ascending
//   checking [3C, 4C)  the ring being tested (ring 1: the base node's list in stored order; later
//                      rings ascending, as the reference's sort + unique leaves them)
//   next     [4C, 5C)  the ring under construction, ascending, no duplicates
// Returns the owner tet, -1 = NULL, -2 = workspace exhausted.
#define GCM_RING_CAP 2048
#define GCM_RING_WS (5 * GCM_RING_CAP)

GCM_HD int ring_lower_bound(const int* a, int n, int v)
{
	int lo = 0, hi = n;
	while(lo < hi) { const int mid = (lo + hi) >> 1; if(a[mid] < v) lo = mid + 1; else hi = mid; }
	return lo;
}

// insert v into the ascending array a[0..n) unless present; returns the new length or -1 (full)
GCM_HD int ring_insert(int* a, int n, int cap, int v)
{
	const int p = ring_lower_bound(a, n, v);
	if(p < n && a[p] == v) return n;
	if(n >= cap) return -1;
	for(int i = n; i > p; i--) a[i] = a[i - 1];
	a[p] = v;
	return n + 1;
}

#if defined(__CUDACC__)
static __host__ __device__ __noinline__     // cold: keep it out of the hot kernels' register allocation
#else
static inline
#endif
int find_owner_general(const MeshView& m, int base, const Vec3& node_pos,
                       float dx, float dy, float dz, TetGeom& g, int* ws)
{
	const int C = GCM_RING_CAP;
	int* checked = ws; int* checking = ws + 3 * C; int* next = ws + 4 * C;
	int n_checked = 0, n_checking = 0;
	const Vec3 X = load_xyz(m, base);
	const float R2 = dx * dx + dy * dy + dz * dz;
	const float delta = sqrtf(R2);
	for(int e = m.n2t_off[base]; e < m.n2t_off[base + 1]; e++) {
		if(n_checking >= C) return -2;
		checking[n_checking++] = m.n2t[e];
	}
	bool inside_R = true;
	while(inside_R) {
		inside_R = false;
		for(int i = 0; i < n_checking; i++) {
			const int t = checking[i];
			load_tet(m, t, g);
			if(point_in_tetr(X, dx, dy, dz, delta, g.p[0], g.p[1], g.p[2], g.p[3], g.h, g.vol))
				return t;
			if(!inside_R) {
				for(int j = 0; j < 4; j++) {
					if(g.v[j] == base) break;   // `break`, not `continue` (:1544-1545)
					const float lx = g.p[j].x, ly = g.p[j].y, lz = g.p[j].z;
					if( ( (node_pos.x - lx) * (node_pos.x - lx)
						+ (node_pos.y - ly) * (node_pos.y - ly)
						+ (node_pos.z - lz) * (node_pos.z - lz) ) < R2 )
						inside_R = true;
				}
			}
		}
		if(!inside_R) return -1;
		// checked += checking; next = (all tets incident to a vertex of `checking`) minus checked
		for(int i = 0; i < n_checking; i++) {
			n_checked = ring_insert(checked, n_checked, 3 * C, checking[i]);
			if(n_checked < 0) return -2;
		}
		int n_next = 0;
		for(int i = 0; i < n_checking; i++) {
			const int* tv = m.tets + 4*(size_t)checking[i];
			for(int j = 0; j < 4; j++) {
				const int v = tv[j];
				for(int c = m.n2t_off[v]; c < m.n2t_off[v + 1]; c++) {
					const int t2 = m.n2t[c];
					const int p = ring_lower_bound(checked, n_checked, t2);
					if(p < n_checked && checked[p] == t2) continue;
					n_next = ring_insert(next, n_next, C, t2);
					if(n_next < 0) return -2;
				}
			}
		}
		for(int i = 0; i < n_next; i++) checking[i] = next[i];
		n_checking = n_next;
	}
	return -1;
}

// Interpolated state at a foot (gcm/meshtypes/TetrMesh_1stOrder.cpp:1764-1880).
// out[0..8] = values, out[9..11] = la, mu, rho (only virtual nodes read those).
// What a foot's interpolation is made of: the owner's vertices (tet order) and their factors.  Filled
// on request (the step-invariant cache of the inner kernel is built from it, gcm_inner.cuh).
struct FootGeom { int v[4]; float f[4]; };

GCM_HD int interpolate_foot(const MeshView& m, const TetGeom& g, float x, float y, float z,
                            float out[12], bool with_material, FootGeom* fg = nullptr)
{
	float f[4];
	if(!interp_factors(g.p[0], g.p[1], g.p[2], g.p[3], x, y, z, g.vol, f))
		return ERR_INTERP_SUM;
	if(fg) for(int k = 0; k < 4; k++) { fg->v[k] = g.v[k]; fg->f[k] = f[k]; }
	float r0[12], r1[12], r2[12], r3[12];
	rec_load12(m.u, m.rs, g.v[0], r0);
	rec_load12(m.u, m.rs, g.v[1], r1);
	rec_load12(m.u, m.rs, g.v[2], r2);
	rec_load12(m.u, m.rs, g.v[3], r3);
	#pragma unroll
	for(int k = 0; k < 9; k++)
		out[k] = interp4(r0[k], r1[k], r2[k], r3[k], f);
	if(with_material) {
		for(int k = 0; k < 3; k++) {
			const float* pl = m.mat + (size_t)k * m.stride;
			out[9 + k] = interp4(pl[g.v[0]], pl[g.v[1]], pl[g.v[2]], pl[g.v[3]], f);
		}
	}
	return ERR_NONE;
}

// The node being advanced: a mesh node, or a virtual (contact) node borrowing the element
// list and local_num of a mesh node (gcm/system/TetrMeshSet.cpp:236-271).
struct CurNode {
	int base;          // local_num: index into the mesh arrays
	Vec3 pos;          // cur_node->coords
	float val[9];      // cur_node->values[0..8]
	float la, mu, rho;
	bool is_border;
	int axis_plus[3], axis_minus[3];  // contact_data (all -1 when none)
	bool has_contact_data;
};

struct FeetOut {
	float val[5][12];  // previous_nodes[count].values (+ la, mu, rho)
	bool inner[9];
	int ppoint[9];
	int owner[5];      // owner tet per distinct point (-1 = none); debug/parity tap
	int n_points;
};

// One pass over the first ring for BOTH senses of one direction q (alpha == 0: every moving foot of
// a stage is X -+ c tau q, and point_in_tetr rescales the displacement to 0.99 h(tet), so the tested
// point depends on the sense only).  The four moving feet of a border node otherwise walk the same
// candidate list four times -- ncu: gathers + pre-filter were ~40 % of the border kernel.
// Per sense s (0: -q, 1: +q):
//   own[s]   first candidate the conservative pre-filter certainly accepts, all earlier ones being
//            certainly rejected; -1 = every candidate certainly rejected
//   amb[s]   some candidate before a decision was not clear-cut -> the foot takes the literal search
//   hmin[s]  min h(tet) over the candidates looked at for that sense: a foot may use the result only
//            if 0 < |dx| < 0.99 hmin (the pre-filter's own precondition, for every such candidate)
//   dmin2    (valid when a sense ends without owner: the whole list was visited) min squared distance
//            node <-> candidate vertex under the reference's scan rule (:1541-1545) = the inside_R test
struct RingScan {
	bool valid;
	int own[2];
	bool amb[2];
	float hmin[2];
	float dmin2;
};

GCM_HD void scan_ring_both_senses(const MeshView& m, int base, const Vec3& X, const float q[3], RingScan& rs)
{
	rs.valid = true;
	rs.own[0] = rs.own[1] = -1;
	rs.amb[0] = rs.amb[1] = false;
	rs.hmin[0] = rs.hmin[1] = 3.4e38f;
	rs.dmin2 = 3.4e38f;
	bool open0 = true, open1 = true;
	const float qn = sqrtf(q[0]*q[0] + q[1]*q[1] + q[2]*q[2]);
	const float inv_qn = qn > 0 ? 1.0f / qn : 0.0f;    // the filter is conservative: its rounding is free
	const int e0 = m.n2t_off[base], e1 = m.n2t_off[base + 1];
	for(int e = e0; e < e1 && (open0 || open1); e++) {
		int t, k;
		float gh, gvol;
		Vec3 A, B, C;
		if(m.n2x_scan) {
			// expanded CSR entry: the other three vertices (tet order) and the position k of the node:
			// no tets[t] hop, and the node's own coordinates are not loaded again
			const int* x = m.n2x_scan + 4*(size_t)e;
#if defined(__CUDA_ARCH__)
			const int4 q4 = *reinterpret_cast<const int4*>(x);
			const int xa = q4.x, xb = q4.y, xc = q4.z, w = q4.w;
#else
			const int xa = x[0], xb = x[1], xc = x[2], w = x[3];
#endif
			k = (int)((unsigned)w >> 30);
			t = w & 0x3fffffff;
			A = load_xyz(m, xa); B = load_xyz(m, xb); C = load_xyz(m, xc);
			gh = m.tet_hv[2*(size_t)t]; gvol = m.tet_hv[2*(size_t)t + 1];
		} else {
			t = m.n2t[e];
			TetGeom g;
			load_tet(m, t, g);
			k = (g.v[0] == base) ? 0 : (g.v[1] == base) ? 1 : (g.v[2] == base) ? 2 : (g.v[3] == base) ? 3 : -1;
			A = g.p[k == 0 ? 1 : 0]; B = g.p[k <= 1 ? 2 : 1]; C = g.p[k == 3 ? 2 : 3];
			gh = g.h; gvol = g.vol;
		}
		if(open0) rs.hmin[0] = fminf(rs.hmin[0], gh);
		if(open1) rs.hmin[1] = fminf(rs.hmin[1], gh);
		int v0 = 0, v1 = 0;    // verdicts for -q and +q
		if(gvol > 0 && k >= 0 && qn > 0) {
			const float eax = A.x - X.x, eay = A.y - X.y, eaz = A.z - X.z;
			const float ebx = B.x - X.x, eby = B.y - X.y, ebz = B.z - X.z;
			const float ecx = C.x - X.x, ecy = C.y - X.y, ecz = C.z - X.z;
			const float bcx = eby * ecz - ebz * ecy, bcy = ebz * ecx - ebx * ecz, bcz = ebx * ecy - eby * ecx;
			const float cax = ecy * eaz - ecz * eay, cay = ecz * eax - ecx * eaz, caz = ecx * eay - ecy * eax;
			const float abx = eay * ebz - eaz * eby, aby = eaz * ebx - eax * ebz, abz = eax * eby - eay * ebx;
			const float D = eax * bcx + eay * bcy + eaz * bcz;
			const float sD = 0.99f * gh * inv_qn * D;
			// barycentric coordinates (times D^2) of X + 0.99 h q/|q|; those of the opposite sense are the negatives
			const float la = sD * (q[0] * bcx + q[1] * bcy + q[2] * bcz);
			const float lb = sD * (q[0] * cax + q[1] * cay + q[2] * caz);
			const float lc = sD * (q[0] * abx + q[1] * aby + q[2] * abz);
			const float D2 = D * D;
			const float sum = la + lb + lc;
			const float mn1 = fminf(fminf(la, lb), fminf(lc, D2 - sum));
			const float mn0 = fminf(fminf(-la, -lb), fminf(-lc, D2 + sum));
			v1 = mn1 < m.pf_fail * D2 ? -1 : (mn1 > m.pf_pass * D2 ? 1 : 0);
			v0 = mn0 < m.pf_fail * D2 ? -1 : (mn0 > m.pf_pass * D2 ? 1 : 0);
		}
		if(open0) { if(v0 > 0) { rs.own[0] = t; open0 = false; } else if(v0 == 0) { rs.amb[0] = true; open0 = false; } }
		if(open1) { if(v1 > 0) { rs.own[1] = t; open1 = false; } else if(v1 == 0) { rs.amb[1] = true; open1 = false; } }
		// the reference scans a tet's vertices in order and `break`s (not `continue`s) at the node itself
		// (:1544-1545): only the vertices BEFORE position k are measured = the first k of (A, B, C)
		for(int j = 0; j < 3 && j < (k < 0 ? 3 : k); j++) {
			const Vec3& P = j == 0 ? A : (j == 1 ? B : C);
			const float d2 = (X.x - P.x) * (X.x - P.x) + (X.y - P.y) * (X.y - P.y) + (X.z - P.z) * (X.z - P.z);
			rs.dmin2 = fminf(rs.dmin2, d2);
		}
	}
}

// One distinct characteristic foot of find_nodes_on_previous_time_layer (:491-671): displacement
// (bent by alpha / replaced by the "last chance" centre of elements[0]), owner search, interpolation
// and the inner / outer classification.  val[12] must hold the node's own values on entry
// (previous_nodes[count] = *cur_node).  alpha is already clamped to <= 1.  Returns a StepError.
// rs (optional): scan_ring_both_senses(ksi) of this node, usable when alpha == 0 and the node sits
// on its base node (not a virtual node).
GCM_HD int foot_point(const MeshView& m, const CurNode& cur, int stage, float alpha, float dks, const float ksi[3],
                      const float Xc[3], const float curc[3], const float tetr_center[3], bool with_material,
                      bool* inner, int* owner_out, float val[12], const RingScan* rs = nullptr, FootGeom* fg = nullptr)
{
	float dx[3], dx_ksi[3];
	float safe_direction[3];
	for(int j = 0; j < 3; j++)
		safe_direction[j] = - ( curc[j] - Xc[j] );
	const bool contact_axis = cur.has_contact_data
		&& ( (cur.axis_plus[stage] != -1) || (cur.axis_minus[stage] != -1) );
	for(int j = 0; j < 3; j++)
		dx_ksi[j] = dks * ksi[j];
	float safe_direction_projection_modul = 1;
	float fc[3];
	for(int j = 0; j < 3; j++) {
		bool bend;
		if(contact_axis) {
			if(cur.axis_plus[stage] != -1) bend = !(dks >= 0);
			else bend = !(dks <= 0);
		} else {
			bend = !(dks >= 0);
		}
		if(!bend) {
			dx[j] = dx_ksi[j];
		} else {
			float sdp = safe_direction_projection_modul * safe_direction[j];
			dx[j] = dx_ksi[j] * (1 - alpha) + sdp * alpha;
		}
		dx[j] += ( curc[j] - Xc[j] );
		fc[j] = Xc[j] + dx[j];
	}
	if((double)alpha > 0.95) {
		if( ! ( cur.is_border && (stage == 0) && (dks > 0) ) )
			for(int z = 0; z < 3; z++) {
				dx[z] = tetr_center[z] - Xc[z];
				fc[z] = Xc[z] + dx[z];
			}
	}

	// A foot on the contact side is classified `outer` whatever the search returns (:594-603), so
	// the search (which the reference still runs, ring after ring into the gap) is skipped;
	// its tap entry is -3 ("not searched").
	const bool forced_outer = cur.has_contact_data
		&& ( ((cur.axis_plus[stage] != -1) && (dks > 0)) || ((cur.axis_minus[stage] != -1) && (dks < 0)) );
	TetGeom g;
	bool expand = false;
	int owner = -3;
	bool settled = false;
	if(!forced_outer && rs && rs->valid && alpha == 0.0f && dks != 0) {
		// the shared scan answers for this foot if it was clear-cut and the foot is one the pre-filter covers
		const int s = dks < 0 ? 0 : 1;
		const float R2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
		const float delta = sqrtf(R2);
		if(!rs->amb[s] && delta > 0 && (double)delta < (double)rs->hmin[s] * 0.99) {
			if(rs->own[s] >= 0) { owner = rs->own[s]; load_tet(m, owner, g); settled = true; }
			else if(!(rs->dmin2 < R2)) { owner = -1; settled = true; }   // not found, no vertex within |dx|: NULL
		}
	}
	if(!forced_outer && !settled) {
		owner = find_owner_first_ring(m, cur.base, cur.pos, dx[0], dx[1], dx[2], g, &expand);
		if(owner < 0 && expand) {
			owner = find_owner_second_ring(m, cur.base, cur.pos, dx[0], dx[1], dx[2], g, &expand);
			if(owner < 0 && expand) {
				// a third ring is needed: the list-building search, where the launch provides a workspace
				// (the deep pass of the contact kernel); elsewhere the node is reported / deferred
				int* ws = m.ring_ws;
#if defined(__CUDA_ARCH__)
				if(ws) ws += (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * GCM_RING_WS;
#endif
				if(!ws) return ERR_RING_OVERFLOW;
				owner = find_owner_general(m, cur.base, cur.pos, dx[0], dx[1], dx[2], g, ws);
				if(owner == -2) return ERR_RING_OVERFLOW;
			}
		}
	}
	*owner_out = owner;

	if( cur.has_contact_data && (cur.axis_plus[stage] != -1) && (dks > 0) ) {
		*inner = false;
	} else if( cur.has_contact_data && (cur.axis_minus[stage] != -1) && (dks < 0) ) {
		*inner = false;
	} else if(owner >= 0) {
		int e = interpolate_foot(m, g, fc[0], fc[1], fc[2], val, with_material, fg);
		if(e) return e;
		*inner = true;
	} else if(dks == 0) {
		*inner = true;   // previous_nodes[count] = *cur_node (already copied)
	} else if(!cur.is_border) {
		return ERR_INNER_NO_OWNER;
	} else {
		*inner = false;
	}
	return ERR_NONE;
}

// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:430-681.  ksi = random_axis[basis_num].ksi[stage].
// Returns the outer count, or -(StepError) on a condition the reference throws on.
GCM_HD int find_nodes_on_previous_time_layer(const MeshView& m, const CurNode& cur, int stage,
	float alpha, const float dksi[9], const float ksi[3], FeetOut& out, bool with_material, FootGeom* fgs = nullptr)
{
	Vec3 X = load_xyz(m, cur.base);
	float Xc[3] = {X.x, X.y, X.z};
	float curc[3] = {cur.pos.x, cur.pos.y, cur.pos.z};
	float safe_direction[3];
	for(int j = 0; j < 3; j++)
		safe_direction[j] = - ( curc[j] - Xc[j] );

	if(alpha > 1.0)
		alpha = 1.0;
	// centre of elements[0]: the "last chance" foot, read by foot_point only when alpha > 0.95
	float tetr_center[3] = {0.f, 0.f, 0.f};
	if((double)alpha > 0.95) {
		int tetr_num = m.n2t[m.n2t_off[cur.base]];
		const int* tv = m.tets + 4*(size_t)tetr_num;
		Vec3 a = load_xyz(m, tv[0]), b = load_xyz(m, tv[1]), c = load_xyz(m, tv[2]), d = load_xyz(m, tv[3]);
		tetr_center[0] = ( a.x + b.x + c.x + d.x ) / 4;
		tetr_center[1] = ( a.y + b.y + c.y + d.y ) / 4;
		tetr_center[2] = ( a.z + b.z + c.z + d.z ) / 4;
	}

	// first try of a real node: one shared walk over the first ring serves all moving feet
	RingScan rs;
	rs.valid = false;
	if(alpha == 0.0f && !m.exact_only && curc[0] == Xc[0] && curc[1] == Xc[1] && curc[2] == Xc[2])
		scan_ring_both_senses(m, cur.base, X, ksi, rs);

	int count = 0;

	for(int i = 0; i < 9; i++)
	{
		bool already_found = false;
		for(int j = 0; j < i; j++) {
			if( (double)fabsf(dksi[i] - dksi[j]) <= 0.01 * (double)fabsf(dksi[i] + dksi[j]) ) {
				already_found = true;
				out.ppoint[i] = out.ppoint[j];
				out.inner[i] = out.inner[j];
			}
		}
		if(already_found) continue;
		if(count >= 5) return -ERR_FOOT_PATTERN;
		out.ppoint[i] = count;
		// previous_nodes[count] = *cur_node
		for(int k = 0; k < 9; k++) out.val[count][k] = cur.val[k];
		out.val[count][9] = cur.la; out.val[count][10] = cur.mu; out.val[count][11] = cur.rho;

		bool inn = false;
		int own = -1;
		int fe = foot_point(m, cur, stage, alpha, dksi[i], ksi, Xc, curc, tetr_center, with_material, &inn, &own, out.val[count], &rs, fgs ? fgs + count : nullptr);
		if(fe) return -fe;
		out.inner[i] = inn;
		out.owner[count] = own;
		count++;
	}
	out.n_points = count;
	int outer_count = 0;
	for(int i = 0; i < 9; i++)
		if(!out.inner[i]) outer_count++;
	return outer_count;
}

// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:78-111 (prepare_node): eigensystem for axis
// `stage` of basis `ksi3x3`, dksi, and the alpha-retry loop.  *tries = number of searches run.
GCM_HD int prepare_node(const MeshView& m, const CurNode& cur, int stage, const float* ksi3x3,
                        float L[9], float* U, float* U1, FeetOut& feet, bool with_material, int* tries)
{
	// prepare_part_step: q = column `stage` of basis^-1 = row `stage` of the basis (:40-49, :903-919)
	const float* ksi = ksi3x3 + 3*stage;
	int e = elastic_matrix(cur.la, cur.mu, cur.rho, ksi[0], ksi[1], ksi[2], L, U, U1);
	if(e) return -e;
	float dksi[9];
	for(int i = 0; i < 9; i++)
		dksi[i] = - L[i] * m.tau;
	float alpha = 0;
	int outer_count;
	int n = 0;
	do {
		outer_count = find_nodes_on_previous_time_layer(m, cur, stage, alpha, dksi, ksi, feet, with_material);
		n++;
		if(outer_count < 0) break;
		alpha = (float)((double)alpha + 0.1);
	} while( (outer_count != 0) && (outer_count != 3) && ((double)alpha < 1.01) );
	*tries = n;
	return outer_count;
}

// omega_i = sum_j U(i,j) * u^{p(i)}_j  (the inner rows of every calculator)
GCM_HD float omega_row(const float* U, int i, const float* vals)
{
	float om = 0;
	for(int j = 0; j < 9; j++)
		om += U[i*9+j] * vals[j];
	return om;
}

// gcm/methods/volume/SimpleVolumeCalculator.cpp:7-32
GCM_HD void volume_calc(const float* U, const float* U1, const FeetOut& feet, float out[9])
{
	float omega[9];
	for(int i = 0; i < 9; i++)
		omega[i] = omega_row(U, i, feet.val[feet.ppoint[i]]);
	for(int i = 0; i < 9; i++) {
		float v = 0;
		for(int j = 0; j < 9; j++)
			v += U1[i*9+j] * omega[j];
		out[i] = v;
	}
}

// Condition row k (0, 1, 2 = first, second, third outer characteristic) of the five border
// calculators: the non-zero entries of the row (the caller zeroed it) and its right-hand side.
// local_n = (n, t1, t2) from create_local_basis(outer_normal), needed by ExternalForce / Velocity.
GCM_HD void border_cond_row(const BorderCond& bc, int k, const float outer_normal[3], const float local_n[3][3],
                            double* row, double* rhs)
{
	const float scale = bc.scale;
	switch(bc.type) {
	case BC_FREE: {
		// FreeBorderCalculator.cpp:59-74
		const int c[3][3] = {{3,4,5},{4,6,7},{5,7,8}};
		row[c[k][0]] = (double)outer_normal[0];
		row[c[k][1]] = (double)outer_normal[1];
		row[c[k][2]] = (double)outer_normal[2];
	} break;
	case BC_FIXED:
		// FixedBorderCalculator.cpp:52-58
		row[k] = 1;
		break;
	case BC_EXT_FORCE: {
		// ExternalForceCalculator.cpp:76-110
		const float* n0 = local_n[0];
		if(k == 0) {
			row[3] = (double)(n0[0] * n0[0]);
			row[4] = (double)(2 * n0[1] * n0[0]);
			row[5] = (double)(2 * n0[2] * n0[0]);
			row[6] = (double)(n0[1] * n0[1]);
			row[7] = (double)(2 * n0[2] * n0[1]);
			row[8] = (double)(n0[2] * n0[2]);
			*rhs = (double)(scale * bc.p_normal);
		} else {
			const float* t = local_n[k];
			row[3] = (double)(n0[0] * t[0]);
			row[4] = (double)(n0[1] * t[0] + n0[0] * t[1]);
			row[5] = (double)(n0[2] * t[0] + n0[0] * t[2]);
			row[6] = (double)(n0[1] * t[1]);
			row[7] = (double)(n0[2] * t[1] + n0[1] * t[2]);
			row[8] = (double)(n0[2] * t[2]);
			*rhs = (double)(scale * bc.p_tangential * (t[0]*bc.dir[0] + t[1]*bc.dir[1] + t[2]*bc.dir[2]));
		}
	} break;
	case BC_EXT_VELOCITY: {
		// ExternalVelocityCalculator.cpp:70-98
		const float* t = local_n[k];
		row[0] = (double)t[0];
		row[1] = (double)t[1];
		row[2] = (double)t[2];
		if(k == 0) *rhs = (double)(scale * bc.p_normal);
		else *rhs = (double)(scale * bc.p_tangential * (t[0]*bc.dir[0] + t[1]*bc.dir[1] + t[2]*bc.dir[2]));
	} break;
	case BC_EXT_VALUES:
		// ExternalValuesCalculator.cpp:63-78
		row[bc.vars_index[k]] = 1;
		*rhs = (double)bc.vars_values[k];
		break;
	}
}

// The five border calculators of gcm/methods/border/*.cpp behind one switch:
// inner rows = Omega rows, the three outer rows (in index order) = condition rows.
// outer_normal = basis.ksi[stage] (TetrMethod_Plastic_1stOrder.cpp:289-297).
GCM_HD void border_calc(const BorderCond& bc, const float* U, const FeetOut& feet,
                        const float outer_normal[3], float out[9])
{
	double A[81];
	double b[9];
	float local_n[3][3];
	local_n[0][0] = outer_normal[0]; local_n[0][1] = outer_normal[1]; local_n[0][2] = outer_normal[2];
	if(bc.type == BC_EXT_FORCE || bc.type == BC_EXT_VELOCITY)
		create_local_basis(local_n[0], local_n[1], local_n[2]);
	const float scale = bc.scale;
	int outer_count = 3;
	for(int i = 0; i < 9; i++) {
		if(feet.inner[i]) {
			b[i] = (double)omega_row(U, i, feet.val[feet.ppoint[i]]);
			for(int j = 0; j < 9; j++) A[i*9+j] = (double)U[i*9+j];
			continue;
		}
		b[i] = 0;
		for(int j = 0; j < 9; j++) A[i*9+j] = 0;
		const int k = 3 - outer_count;   // 0, 1, 2: which condition row
		if(outer_count <= 0) continue;
		outer_count--;
		border_cond_row(bc, k, outer_normal, local_n, A + i*9, &b[i]);
	}
	lu_solve<9>(A, b);
	for(int j = 0; j < 9; j++)
		out[j] = (float)b[j];
}

// Debug/parity tap of one node-stage: owner per distinct point of the LAST alpha try.
struct StageTap { int owner[5]; int outer_count; int tries; };

// One mesh node, one axis stage (stage 0..2), without contact:
// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:190-297.  Writes out[9]; returns a StepError.
// `Decide` (only_try >= 0): callable bool(outer_count, try, ctx) that returns true on the one lane
// whose try decides the node; every lane of the group must call it (it votes).  The default
// (only_try = -1) runs the sequential retry loop.  Returns -1000 on the lanes that lost the vote.
struct NoDecide { GCM_HD bool operator()(int, int, void*) const { return true; } };

template<class Decide>
GCM_HD int node_stage_nocontact_t(const MeshView& m, int node, int stage, float out[9], StageTap* tap,
                                  int only_try, Decide decide, void* decide_ctx);

GCM_HD int node_stage_nocontact(const MeshView& m, int node, int stage, float out[9], StageTap* tap)
{
	return node_stage_nocontact_t(m, node, stage, out, tap, -1, NoDecide(), nullptr);
}

template<class Decide>
GCM_HD int node_stage_nocontact_t(const MeshView& m, int node, int stage, float out[9], StageTap* tap,
                                  int only_try, Decide decide, void* decide_ctx)
{
	CurNode cur;
	cur.base = node;
	bool bad_value = false;
	{
		float r[12];
		rec_load12(m.u, m.rs, node, r);
		for(int k = 0; k < 9; k++) cur.val[k] = r[k];
		cur.pos.x = r[9]; cur.pos.y = r[10]; cur.pos.z = r[11];
	}
	for(int k = 0; k < 9; k++) {
		const float v = cur.val[k];
		if(isnan(v) || isinf(v)) { if(only_try < 0) return ERR_NAN; bad_value = true; }
	}
	cur.la = m.mat[node]; cur.mu = m.mat[m.stride + node]; cur.rho = m.mat[2*m.stride + node];
	cur.is_border = (m.flags[node] & NF_BORDER) != 0;
	cur.has_contact_data = cur.is_border;
	for(int k = 0; k < 3; k++) { cur.axis_plus[k] = -1; cur.axis_minus[k] = -1; }

	const float E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	const float* basis = E;
	int slot = -1;
	if(cur.is_border) {
		slot = m.border_slot[node];
		basis = m.basis + 9*(size_t)slot;
	}
	// The feet need the wave speeds only; Omega (and Omega^-1 when every characteristic is inner) is
	// built after the search, in ONE pass of the same routine on the same inputs -- the shared part of
	// the eigensystem is then computed once per stage instead of twice and is not live across the search.
	float L[9], U[81];
	FeetOut feet;
	int tries = 0;
	int outer_count;
	if(only_try < 0) {
		outer_count = prepare_node(m, cur, stage, basis, L, nullptr, nullptr, feet, false, &tries);
	} else {
		// one alpha try of the retry loop (speculative evaluation of all 11 tries by 11 lanes,
		// stage_sharp_kernel): same eigensystem, same dksi, alpha = value the loop would have then
		const float* ksi = basis + 3*stage;
		int e = (bad_value || only_try > 10) ? 0 : elastic_matrix(cur.la, cur.mu, cur.rho, ksi[0], ksi[1], ksi[2], L, nullptr, nullptr);
		if(only_try > 10) outer_count = 1;          // padding lane of the group: votes "no"
		else if(bad_value) outer_count = -ERR_NAN;
		else if(e) outer_count = -e;
		else {
			float dksi[9];
			for(int i = 0; i < 9; i++) dksi[i] = - L[i] * m.tau;
			float alpha = 0;
			for(int t = 0; t < only_try; t++) alpha = (float)((double)alpha + 0.1);
			outer_count = find_nodes_on_previous_time_layer(m, cur, stage, alpha, dksi, ksi, feet, false);
		}
		tries = only_try + 1;
		// the lanes of a node now agree on which try decides: the first with a legal outer count or an
		// error, else the last one (prepare_node's loop, :105-108)
		if(!decide(outer_count, only_try, decide_ctx)) return -1000;
	}
	if(tap) {
		for(int k = 0; k < 5; k++) tap->owner[k] = (outer_count >= 0 && k < feet.n_points) ? feet.owner[k] : -2;
		tap->outer_count = outer_count; tap->tries = tries;
	}
	if(outer_count < 0) return -outer_count;
	const float* q = basis + 3*stage;
	if(outer_count == 0) {
		float U1[81];
		elastic_matrix(cur.la, cur.mu, cur.rho, q[0], q[1], q[2], L, U, U1);
		volume_calc(U, U1, feet, out);
		return ERR_NONE;
	}
	if(outer_count != 3) return ERR_OUTER_COUNT;
	if(!cur.is_border) return ERR_NOT_BORDER;
	elastic_matrix(cur.la, cur.mu, cur.rho, q[0], q[1], q[2], L, U, nullptr);
	// default condition: FreeBorderCalculator with an always-on form (TetrMeshSet.cpp:11-21)
	BorderCond def;
	def.type = BC_FREE; def.scale = 1; def.active = 1;
	const BorderCond* bc = &def;
	int cid = m.cond_id ? m.cond_id[slot] : -1;
	if(cid >= 0 && m.conds[cid].active) bc = &m.conds[cid];
	border_calc(*bc, U, feet, basis + 3*stage, out);
	return ERR_NONE;
}

// Out-of-line copy of the general routine for rarely taken fallbacks (keeps its 1.7 KB of arrays
// off the caller's stack frame).
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
int node_stage_nocontact_cold(const MeshView& m, int node, int stage, float out[9], StageTap* tap)
{
	return node_stage_nocontact(m, node, stage, out, tap);
}

// Border node without contact, small-footprint variant of node_stage_nocontact: the same calls
// (foot_point per distinct foot, the retry loop of prepare_node, the calculators) but Omega and
// Omega^-1 are produced row by row (elastic_U_row / elastic_U1_row, bitwise equal to the entries of
// elastic_matrix) instead of as two 81-float arrays, and the five-point pattern of the
// characteristic speeds is assumed (checked; anything else goes to the general routine).  The
// thread's stack shrinks from ~1.7 KB to ~0.4 KB on the two tangential stages (outer == 0), which
// is what keeps the border launch out of L2-resident local memory.
GCM_HD int node_stage_border_lean(const MeshView& m, int node, int stage, float out[9], StageTap* tap)
{
	CurNode cur;
	cur.base = node;
	cur.pos = load_xyz(m, node);
	for(int k = 0; k < 9; k++) {
		float v = rec_get(m.u, m.rs, node, k);
		if(isnan(v) || isinf(v)) return ERR_NAN;
		cur.val[k] = v;
	}
	cur.la = m.mat[node]; cur.mu = m.mat[m.stride + node]; cur.rho = m.mat[2*m.stride + node];
	cur.is_border = true;
	cur.has_contact_data = true;
	for(int k = 0; k < 3; k++) { cur.axis_plus[k] = -1; cur.axis_minus[k] = -1; }
	const int slot = m.border_slot[node];
	const float* basis = m.basis + 9*(size_t)slot;
	const float ksi[3] = {basis[3*stage], basis[3*stage + 1], basis[3*stage + 2]};

	ElasticDirs d;
	{
		int e = elastic_dirs(cur.la, cur.mu, cur.rho, ksi[0], ksi[1], ksi[2], d);
		if(e) return e;
	}
	// dksi[i] = - L(i,i) * tau of the five distinct points (characteristics 0, 1, 2, 3, 6)
	float dk[5];
	for(int p = 0; p < 5; p++) dk[p] = - elastic_L(d, p < 4 ? p : 6) * m.tau;
	bool usual = true;
	for(int p = 1; p < 5; p++)
		for(int q = 0; q < p; q++)
			if( (double)fabsf(dk[p] - dk[q]) <= 0.01 * (double)fabsf(dk[p] + dk[q]) ) usual = false;
	if(!usual) return node_stage_nocontact_cold(m, node, stage, out, tap);

	const float Xc[3] = {cur.pos.x, cur.pos.y, cur.pos.z};
	float val[5][9];
	bool inn[5];
	int own[5];
	int outer_count = 0, tries = 0;
	float alpha = 0;
	RingScan rs;
	scan_ring_both_senses(m, node, cur.pos, ksi, rs);   // serves the first try (alpha == 0)
	do {
		const float a = ((double)alpha > 1.0) ? 1.0f : alpha;
		// centre of elements[0]: the "last chance" foot, read by foot_point only when alpha > 0.95
		float tetr_center[3] = {0.f, 0.f, 0.f};
		if((double)a > 0.95) {
			int tetr_num = m.n2t[m.n2t_off[node]];
			const int* tv = m.tets + 4*(size_t)tetr_num;
			Vec3 ta = load_xyz(m, tv[0]), tb = load_xyz(m, tv[1]), tc = load_xyz(m, tv[2]), te = load_xyz(m, tv[3]);
			tetr_center[0] = ( ta.x + tb.x + tc.x + te.x ) / 4;
			tetr_center[1] = ( ta.y + tb.y + tc.y + te.y ) / 4;
			tetr_center[2] = ( ta.z + tb.z + tc.z + te.z ) / 4;
		}
		int ferr = 0;
		for(int p = 0; p < 5 && !ferr; p++) {
			float v12[12];
			for(int k = 0; k < 9; k++) v12[k] = cur.val[k];
			v12[9] = cur.la; v12[10] = cur.mu; v12[11] = cur.rho;
			inn[p] = true; own[p] = -2;
			ferr = foot_point(m, cur, stage, a, dk[p], ksi, Xc, Xc, tetr_center, false, &inn[p], &own[p], v12, &rs);
			for(int k = 0; k < 9; k++) val[p][k] = v12[k];
		}
		tries++;
		if(ferr) {
			if(tap) { for(int k = 0; k < 5; k++) tap->owner[k] = -2; tap->outer_count = -ferr; tap->tries = tries; }
			return ferr;
		}
		outer_count = (inn[0] ? 0 : 1) + (inn[1] ? 0 : 1) + (inn[2] ? 0 : 2) + (inn[3] ? 0 : 2) + (inn[4] ? 0 : 3);
		alpha = (float)((double)alpha + 0.1);
	} while( (outer_count != 0) && (outer_count != 3) && ((double)alpha < 1.01) );
	if(tap) {
		for(int k = 0; k < 5; k++) tap->owner[k] = own[k];
		tap->outer_count = outer_count; tap->tries = tries;
	}
	if(outer_count != 0 && outer_count != 3) return ERR_OUTER_COUNT;

	const int pp[9] = {0, 1, 2, 3, 2, 3, 4, 4, 4};
	float omega[9];
	float row[9];
	if(outer_count == 0) {
		for(int i = 0; i < 9; i++) {
			elastic_U_row(d, cur.la, cur.mu, cur.rho, i, row);
			omega[i] = omega_row(row, 0, val[pp[i]]);
		}
		for(int r = 0; r < 9; r++) {
			elastic_U1_row(d, cur.rho, r, row);
			float v = 0;
			for(int j = 0; j < 9; j++) v += row[j] * omega[j];
			out[r] = v;
		}
		return ERR_NONE;
	}
	// outer_count == 3: border calculators
	BorderCond def;
	def.type = BC_FREE; def.scale = 1; def.active = 1;
	const BorderCond* bc = &def;
	int cid = m.cond_id ? m.cond_id[slot] : -1;
	if(cid >= 0 && m.conds[cid].active) bc = &m.conds[cid];
	float local_n[3][3];
	local_n[0][0] = ksi[0]; local_n[0][1] = ksi[1]; local_n[0][2] = ksi[2];
	if(bc->type == BC_EXT_FORCE || bc->type == BC_EXT_VELOCITY)
		create_local_basis(local_n[0], local_n[1], local_n[2]);
	double A[81];
	double b[9];
	int k = 0;
	for(int i = 0; i < 9; i++) {
		if(inn[pp[i]]) {
			elastic_U_row(d, cur.la, cur.mu, cur.rho, i, row);
			b[i] = (double)omega_row(row, 0, val[pp[i]]);
			for(int j = 0; j < 9; j++) A[i*9+j] = (double)row[j];
		} else {
			b[i] = 0;
			for(int j = 0; j < 9; j++) A[i*9+j] = 0;
			if(k < 3) border_cond_row(*bc, k, ksi, local_n, A + i*9, &b[i]);
			k++;
		}
	}
	lu_solve<9>(A, b);
	for(int j = 0; j < 9; j++)
		out[j] = (float)b[j];
	return ERR_NONE;
}

// ---- two-body contact ------------------------------------------------------------------------
// A virtual node as the device sees it (built from gcmh::VirtNodeHost): position on the partner
// body's border face, the 13 values interpolated there at discovery time, and the partner-mesh
// node whose `elements` / local_num it borrows (TetrMeshSet.cpp:236-271).
struct VirtNode {
	float pos[3];
	float val[9];
	float la, mu, rho, yield;
	int base;      // merged index of the borrowed node (vert[0] of the face)
	int real;      // merged index of the border node it belongs to
	int phase;     // 0: partner zone comes later in local_meshes (reads the current layer);
	               // 1: partner zone was already advanced this stage (reads the next layer)
};

// AdhesionContactCalculator::do_calc, gcm/methods/contact/AdhesionContactCalculator.cpp:19-142:
// 18 unknowns (real 9 + virtual 9); rows 0-5 real inner Omega rows, 6-11 virtual inner Omega rows,
// 12-14 equal velocities, 15-17 equal traction (sigma . n).  x[0..8] = real node, x[9..17] = virtual.
GCM_HD void adhesion_calc(const float* U, const FeetOut& feet, const float* Uv, const FeetOut& vfeet,
                          const float outer_normal[3], double x[18])
{
	double A[324];
	for(int i = 0; i < 324; i++) A[i] = 0;
	for(int i = 0; i < 18; i++) x[i] = 0;
	int pos = 0;
	for(int i = 0; i < 9; i++)
		if(feet.inner[i] && pos < 6) {
			x[pos] = (double)omega_row(U, i, feet.val[feet.ppoint[i]]);
			for(int j = 0; j < 9; j++) A[pos*18 + j] = (double)U[i*9 + j];
			pos++;
		}
	pos = 0;
	for(int i = 0; i < 9; i++)
		if(vfeet.inner[i] && pos < 6) {
			x[6 + pos] = (double)omega_row(Uv, i, vfeet.val[vfeet.ppoint[i]]);
			for(int j = 0; j < 9; j++) A[(6 + pos)*18 + 9 + j] = (double)Uv[i*9 + j];
			pos++;
		}
	A[12*18 + 0] = 1; A[12*18 + 9] = -1;
	A[13*18 + 1] = 1; A[13*18 + 10] = -1;
	A[14*18 + 2] = 1; A[14*18 + 11] = -1;
	const int c[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};
	for(int r = 0; r < 3; r++)
		for(int k = 0; k < 3; k++) {
			A[(15 + r)*18 + c[r][k]] = (double)outer_normal[k];
			A[(15 + r)*18 + 9 + c[r][k]] = (double)(-outer_normal[k]);
		}
	lu_solve<18>(A, x);
}

// One border node in contact on axis `stage` (axis_plus[stage] = its virtual node), i.e. the
// contact branch of do_next_part_step, gcm/methods/TetrMethod_Plastic_1stOrder.cpp:299-421.
//   m   view whose `u` is the layer the REAL node reads (its own mesh, current layer)
//   mv  view whose `u` is the layer the VIRTUAL node's feet read in the partner mesh: the current
//       layer, or the next one when the partner was already advanced in this stage (the
//       reference runs its meshes one after the other, TetrMeshSet.cpp:158-167)
//   normal = find_border_node_normal of the real node (prepare_node :83-84)
// first_step: this is the first step after loading (only matters to Friction, see below).
// out[9] = new values of the real node.  virt_out (optional) receives the virtual node's values
// when the calculator rewrites them (Friction's swapped Adhesion call); *virt_written says so.
GCM_HD int node_stage_contact(const MeshView& m, const MeshView& mv, int node, const VirtNode& vn, int vidx,
                              int stage, const float normal[3], int calc_type, float out[9], StageTap* tap,
                              float virt_out[9], bool* virt_written, bool first_step)
{
	if(virt_written) *virt_written = false;
	CurNode cur;
	cur.base = node;
	cur.pos = load_xyz(m, node);
	for(int k = 0; k < 9; k++) {
		float v = rec_get(m.u, m.rs, node, k);
		if(isnan(v) || isinf(v)) return ERR_NAN;
		cur.val[k] = v;
	}
	cur.la = m.mat[node]; cur.mu = m.mat[m.stride + node]; cur.rho = m.mat[2*m.stride + node];
	cur.is_border = true;
	cur.has_contact_data = true;
	for(int k = 0; k < 3; k++) { cur.axis_plus[k] = -1; cur.axis_minus[k] = -1; }
	cur.axis_plus[stage] = vidx;
	const int slot = m.border_slot[node];
	const float* basis = m.basis + 9*(size_t)slot;

	float L[9], U[81];
	FeetOut feet;
	int tries = 0;
	int outer_count = prepare_node(m, cur, stage, basis, L, U, nullptr, feet, false, &tries);
	if(tap) {
		for(int k = 0; k < 5; k++) tap->owner[k] = (outer_count >= 0 && k < feet.n_points) ? feet.owner[k] : -2;
		tap->outer_count = outer_count; tap->tries = tries;
	}
	if(outer_count < 0) return -outer_count;
	if(outer_count == 0) {
		float U1[81];
		const float* q = basis + 3*stage;
		elastic_matrix(cur.la, cur.mu, cur.rho, q[0], q[1], q[2], L, nullptr, U1);
		volume_calc(U, U1, feet, out);
		return ERR_NONE;
	}
	if(outer_count != 3) return ERR_OUTER_COUNT;

	// the virtual node, marked as in contact from the other side (:310-315)
	CurNode vc;
	vc.base = vn.base;
	vc.pos.x = vn.pos[0]; vc.pos.y = vn.pos[1]; vc.pos.z = vn.pos[2];
	for(int k = 0; k < 9; k++) vc.val[k] = vn.val[k];
	vc.la = vn.la; vc.mu = vn.mu; vc.rho = vn.rho;
	vc.is_border = true;
	vc.has_contact_data = true;
	for(int k = 0; k < 3; k++) { vc.axis_plus[k] = -1; vc.axis_minus[k] = -1; }
	vc.axis_minus[stage] = vidx;
	float Lv[9], Uv[81];
	FeetOut vfeet;
	int vtries = 0;
	int virt_outer = prepare_node(mv, vc, stage, basis, Lv, Uv, nullptr, vfeet, false, &vtries);   // basis_num = the real node's
	if(virt_outer < 0) return -virt_outer;
	if(virt_outer != 3) return ERR_VIRT_OUTER_COUNT;

	double x[18];
	if(calc_type == CC_ADHESION) {
		adhesion_calc(U, feet, Uv, vfeet, normal, x);
		for(int j = 0; j < 9; j++) out[j] = (float)x[j];
		return ERR_NONE;
	}

	// FrictionContactCalculator::do_calc, gcm/methods/contact/FrictionContactCalculator.cpp:23-132
	// (k = 0.1; the free-border shortcut is disabled there by `if (false)`)
	const float* on = normal;
	float vel_rel[3] = {cur.val[0] - vn.val[0], cur.val[1] - vn.val[1], cur.val[2] - vn.val[2]};
	float dsxx = -vn.val[3] + cur.val[3], dsxy = -vn.val[4] + cur.val[4], dsxz = -vn.val[5] + cur.val[5],
	      dsyy = -vn.val[6] + cur.val[6], dsyz = -vn.val[7] + cur.val[7], dszz = -vn.val[8] + cur.val[8];
	float rho = cur.rho, c1 = sqrtf( (cur.la + 2 * cur.mu) / rho );
	float vel_norm_abs = vel_rel[0]*on[0] + vel_rel[1]*on[1] + vel_rel[2]*on[2];
	float vel_tang[3] = {vel_rel[0] - vel_norm_abs*on[0], vel_rel[1] - vel_norm_abs*on[1], vel_rel[2] - vel_norm_abs*on[2]};
	float vel_tang_abs = sqrtf(vel_tang[0]*vel_tang[0] + vel_tang[1]*vel_tang[1] + vel_tang[2]*vel_tang[2]);
	const float kf = (float)0.1;
	adhesion_calc(U, feet, Uv, vfeet, normal, x);
	float nd[9];
	for(int j = 0; j < 9; j++) nd[j] = (float)x[j];
	float frc_nd[3] = { nd[3]*on[0] + nd[4]*on[1] + nd[5]*on[2],
	                    nd[4]*on[0] + nd[6]*on[1] + nd[7]*on[2],
	                    nd[5]*on[0] + nd[7]*on[1] + nd[8]*on[2] };
	float frc_tan = (vel_tang[0]*frc_nd[0] + vel_tang[1]*frc_nd[1] + vel_tang[2]*frc_nd[2]) / vel_tang_abs;
	float frc_nrm = on[0]*frc_nd[0] + on[1]*frc_nd[1] + on[2]*frc_nd[2];
	if(kf * frc_nrm > frc_tan) {
		// sliding: ExternalForceCalculator with sigma_n, k sigma_n along the tangential velocity
		float tp[3] = { dsxx*on[0] + dsxy*on[1] + dsxz*on[2],
		                dsxy*on[0] + dsyy*on[1] + dsyz*on[2],
		                dsxz*on[0] + dsyz*on[1] + dszz*on[2] };
		// - 0.5*rho*c1*vel_norm_abs - 0.5*scalar_product(tp, n): the 0.5 literals promote to double
		float sn = (float)( ((-0.5 * (double)rho) * (double)c1) * (double)vel_norm_abs
		                    - 0.5 * (double)(tp[0]*on[0] + tp[1]*on[1] + tp[2]*on[2]) );
		if((double)sn > 0.0) sn = -sn;
		BorderCond bc;
		bc.type = BC_EXT_FORCE; bc.scale = 1; bc.active = 1;
		// set_parameters(sn, k*sn, vel_tang / |vel_tang|): the direction is normalised once more there
		bc.p_normal = sn; bc.p_tangential = kf * sn;
		float d0 = vel_tang[0] / vel_tang_abs, d1 = vel_tang[1] / vel_tang_abs, d2 = vel_tang[2] / vel_tang_abs;
		float dtmp = sqrtf(d0*d0 + d1*d1 + d2*d2);
		bc.dir[0] = d0 / dtmp; bc.dir[1] = d1 / dtmp; bc.dir[2] = d2 / dtmp;
		border_calc(bc, U, feet, normal, out);
		return ERR_NONE;
	}
	// "adhesion" branch with new_node / virt_node swapped (:123): the solution lands in the VIRTUAL
	// node's values and the real node keeps what new_node held (the previous step's result)
	// new_nodes == nodes at the start of every step but the first, where new_nodes still holds the
	// values the node was added with (add_node runs before any initial state is written: zeros)
	for(int j = 0; j < 9; j++) out[j] = first_step ? 0.0f : cur.val[j];
	if(virt_out) { for(int j = 0; j < 9; j++) virt_out[j] = nd[j]; if(virt_written) *virt_written = true; }
	return ERR_NONE;
}

} // namespace gcm
