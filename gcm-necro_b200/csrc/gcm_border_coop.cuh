// gcm_border_coop.cuh -- border nodes, one axis stage, 16 lanes per node.
//
// The per-thread routine (gcm_step.cuh: node_stage_nocontact) walks 5 feet, builds two 9x9
// matrices and solves a 9x9 system in double, all in one thread and mostly in local memory: a
// long, latency-bound thread, and the border launch is too small to hide it (it decides the
// strong-scaling floor).  Here the same arithmetic is spread over a half-warp:
//   lanes 0..4  one characteristic foot each (foot_point(): displacement, owner search,
//               interpolation) -- all alpha tries are taken together, the outer count is a vote;
//   lanes 0..8  one row each of Omega / Omega^-1 (elastic_U_row / elastic_U1_row) and of the 9x9
//               border system; LU with partial pivoting runs across the lanes with shuffles, each
//               matrix element seeing exactly the operations, in the order, of lu_solve<9>.
// Results are bit-identical to the per-thread routine (tests: every border case of the golden set).
// Nodes whose characteristic speeds do not form the usual 5-point pattern fall back to the
// per-thread routine on lane 0.
#pragma once
#include "gcm_step.cuh"

namespace gcm {

#if defined(__CUDACC__)

__device__ __forceinline__ bool close_dksi(float a, float b)
{
	return (double)fabsf(a - b) <= 0.01 * (double)fabsf(a + b);   // TetrMethod_Plastic_1stOrder.cpp:483
}

__global__ void __launch_bounds__(128)
stage_border_coop_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list,
                         int stage, float* __restrict__ un, int* err, int zone, StageTap* tap)
{
	const int lane = threadIdx.x & 31;
	const int gbase = lane & 16;
	const int l = lane & 15;
	const unsigned gmask = 0xFFFFu << gbase;
	const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
	if(idx >= n_list) return;                 // the whole half-warp leaves together
	const int node = list[idx];

	// ---- data every lane needs -------------------------------------------------------------
	CurNode cur;
	cur.base = node;
	cur.pos = load_xyz(m, node);
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) {
		const float v = m.u[(size_t)node * GCM_REC + k];
		bad |= (isnan(v) || isinf(v));
		cur.val[k] = v;
	}
	if(bad) { if(l == 0) report_error(err, ERR_NAN, zone, node, stage); return; }
	cur.la = m.mat[node]; cur.mu = m.mat[m.stride + node]; cur.rho = m.mat[2 * m.stride + node];
	cur.is_border = true;
	cur.has_contact_data = true;
	#pragma unroll
	for(int k = 0; k < 3; k++) { cur.axis_plus[k] = -1; cur.axis_minus[k] = -1; }
	const int slot = m.border_slot[node];
	const float* basis = m.basis + 9 * (size_t)slot;
	const float ksi[3] = {basis[3 * stage], basis[3 * stage + 1], basis[3 * stage + 2]};

	ElasticDirs d;
	{
		const int e = elastic_dirs(cur.la, cur.mu, cur.rho, ksi[0], ksi[1], ksi[2], d);
		if(e) { if(l == 0) report_error(err, e, zone, node, stage); return; }
	}
	// dksi[i] = - L(i,i) * tau of the five distinct points (characteristics 0, 1, 2, 3, 6)
	const float dk0 = - elastic_L(d, 0) * m.tau, dk1 = - elastic_L(d, 1) * m.tau;
	const float dk2 = - elastic_L(d, 2) * m.tau, dk3 = - elastic_L(d, 3) * m.tau, dk4 = - elastic_L(d, 6) * m.tau;
	const bool usual = !close_dksi(dk1, dk0) && !close_dksi(dk2, dk0) && !close_dksi(dk2, dk1)
		&& !close_dksi(dk3, dk0) && !close_dksi(dk3, dk1) && !close_dksi(dk3, dk2)
		&& !close_dksi(dk4, dk0) && !close_dksi(dk4, dk1) && !close_dksi(dk4, dk2) && !close_dksi(dk4, dk3);
	if(!usual) {
		if(l == 0) {
			float out[9];
			StageTap t;
			const int e = node_stage_nocontact(m, node, stage, out, tap ? &t : nullptr);
			if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
			if(e) report_error(err, e, zone, node, stage);
			else for(int k = 0; k < 9; k++) un[(size_t)node * GCM_REC + k] = out[k];
		}
		return;
	}
	const float dks = l == 0 ? dk0 : (l == 1 ? dk1 : (l == 2 ? dk2 : (l == 3 ? dk3 : dk4)));

	const float Xc[3] = {cur.pos.x, cur.pos.y, cur.pos.z};
	float tetr_center[3];
	{
		const int tetr_num = m.n2t[m.n2t_off[node]];
		const int* tv = m.tets + 4 * (size_t)tetr_num;
		const Vec3 a = load_xyz(m, tv[0]), b = load_xyz(m, tv[1]), c = load_xyz(m, tv[2]), e = load_xyz(m, tv[3]);
		tetr_center[0] = ( a.x + b.x + c.x + e.x ) / 4;
		tetr_center[1] = ( a.y + b.y + c.y + e.y ) / 4;
		tetr_center[2] = ( a.z + b.z + c.z + e.z ) / 4;
	}

	// ---- prepare_node's retry loop, the five feet in parallel ---------------------------------
	float val[12];
	bool inn = true;
	int own = -2;
	int outer = 0, tries = 0;
	unsigned om = 0;
	float alpha = 0;
	do {
		const float a = ((double)alpha > 1.0) ? 1.0f : alpha;
		int ferr = 0;
		if(l < 5) {
			#pragma unroll
			for(int k = 0; k < 9; k++) val[k] = cur.val[k];
			val[9] = cur.la; val[10] = cur.mu; val[11] = cur.rho;
			ferr = foot_point(m, cur, stage, a, dks, ksi, Xc, Xc, tetr_center, false, &inn, &own, val);
		}
		const unsigned em = (__ballot_sync(gmask, l < 5 && ferr != 0) >> gbase) & 0x1Fu;
		tries++;
		if(em) {   // the per-thread routine stops at the first failing point
			const int code = __shfl_sync(gmask, ferr, gbase + __ffs(em) - 1);
			if(l == 0) report_error(err, code, zone, node, stage);
			return;
		}
		om = (__ballot_sync(gmask, l < 5 && !inn) >> gbase) & 0x1Fu;   // bit p: point p is outer
		outer = (int)(om & 1u) + (int)((om >> 1) & 1u) + 2 * (int)((om >> 2) & 1u) + 2 * (int)((om >> 3) & 1u) + 3 * (int)((om >> 4) & 1u);
		alpha = (float)((double)alpha + 0.1);
	} while( (outer != 0) && (outer != 3) && ((double)alpha < 1.01) );

	if(tap) {
		StageTap t;
		#pragma unroll
		for(int p = 0; p < 5; p++) t.owner[p] = __shfl_sync(gmask, own, gbase + p);
		t.outer_count = outer; t.tries = tries;
		if(l == 0) tap[(size_t)stage * m.n_nodes + node] = t;
	}
	if(outer != 0 && outer != 3) { if(l == 0) report_error(err, ERR_OUTER_COUNT, zone, node, stage); return; }

	// ---- omega_i = Omega(i,:) . u(foot of i) on the lane that owns the foot ---------------------
	// characteristic i -> point (0,1,2,3,2,3,4,4,4): lane 2 also serves i = 4, lane 3 i = 5, lane 4 i = 6,7,8
	float oa = 0, ob = 0, oc = 0;
	if(l < 5 && inn) {
		float row[9];
		const int ia = l < 4 ? l : 6;
		elastic_U_row(d, cur.la, cur.mu, cur.rho, ia, row);
		oa = omega_row(row, 0, val);
		if(l >= 2) {
			elastic_U_row(d, cur.la, cur.mu, cur.rho, l == 4 ? 7 : l + 2, row);
			ob = omega_row(row, 0, val);
		}
		if(l == 4) {
			elastic_U_row(d, cur.la, cur.mu, cur.rho, 8, row);
			oc = omega_row(row, 0, val);
		}
	}
	float omega[9];
	omega[0] = __shfl_sync(gmask, oa, gbase + 0);
	omega[1] = __shfl_sync(gmask, oa, gbase + 1);
	omega[2] = __shfl_sync(gmask, oa, gbase + 2);
	omega[3] = __shfl_sync(gmask, oa, gbase + 3);
	omega[4] = __shfl_sync(gmask, ob, gbase + 2);
	omega[5] = __shfl_sync(gmask, ob, gbase + 3);
	omega[6] = __shfl_sync(gmask, oa, gbase + 4);
	omega[7] = __shfl_sync(gmask, ob, gbase + 4);
	omega[8] = __shfl_sync(gmask, oc, gbase + 4);

	if(outer == 0) {
		// SimpleVolumeCalculator.cpp:24-31, row l of Omega^-1 on lane l
		if(l < 9) {
			float row[9];
			elastic_U1_row(d, cur.rho, l, row);
			float v = 0;
			#pragma unroll
			for(int j = 0; j < 9; j++) v += row[j] * omega[j];
			un[(size_t)node * GCM_REC + l] = v;
		}
		return;
	}

	// ---- outer == 3: the border calculators (gcm/methods/border/*.cpp), row l on lane l ---------
	BorderCond def;
	def.type = BC_FREE; def.scale = 1; def.active = 1;
	const BorderCond* bc = &def;
	const int cid = m.cond_id ? m.cond_id[slot] : -1;
	if(cid >= 0 && m.conds[cid].active) bc = &m.conds[cid];
	float local_n[3][3];
	local_n[0][0] = ksi[0]; local_n[0][1] = ksi[1]; local_n[0][2] = ksi[2];
	if(bc->type == BC_EXT_FORCE || bc->type == BC_EXT_VELOCITY)
		create_local_basis(local_n[0], local_n[1], local_n[2]);
	// inner[i] of the nine characteristics from the five point bits
	const unsigned outer_chars = (om & 1u) | (om & 2u) | (om & 4u) | (om & 8u) | (((om >> 2) & 1u) << 4) | (((om >> 3) & 1u) << 5)
		| (((om >> 4) & 1u) * 0x1C0u);
	double a[9];
	double b = 0;
	#pragma unroll
	for(int j = 0; j < 9; j++) a[j] = 0;
	if(l < 9) {
		if(!((outer_chars >> l) & 1u)) {
			float row[9];
			elastic_U_row(d, cur.la, cur.mu, cur.rho, l, row);
			#pragma unroll
			for(int j = 0; j < 9; j++) a[j] = (double)row[j];
			float om_l = omega[0];
			#pragma unroll
			for(int j = 1; j < 9; j++) if(l == j) om_l = omega[j];
			b = (double)om_l;
		} else {
			const int k = __popc(outer_chars & ((1u << l) - 1u));   // 0, 1, 2: which condition row
			if(k < 3) {
				// border_cond_row indexes the row dynamically: build it in a small local array
				double row[9];
				for(int j = 0; j < 9; j++) row[j] = 0;
				border_cond_row(*bc, k, ksi, local_n, row, &b);
				#pragma unroll
				for(int j = 0; j < 9; j++) a[j] = row[j];
			}
		}
	}

	// ---- lu_solve<9> across lanes 0..8: gsl_linalg_LU_decomp / LU_solve semantics ---------------
	int pos = l < 9 ? l : 100 + l;      // row position this lane's row currently occupies
	#pragma unroll
	for(int j = 0; j < 8; j++) {
		// pivot = first row (in position order) of maximal |a_ij| among positions >= j
		double mag = (l < 9 && pos >= j) ? fabs(a[j]) : -1.0;
		int ppos = pos;
		#pragma unroll
		for(int off = 8; off >= 1; off >>= 1) {
			const double omag = __shfl_xor_sync(gmask, mag, off);
			const int opos = __shfl_xor_sync(gmask, ppos, off);
			if(omag > mag || (omag == mag && opos < ppos)) { mag = omag; ppos = opos; }
		}
		if(ppos != j) { if(pos == ppos) pos = j; else if(pos == j) pos = ppos; }
		const unsigned pm = (__ballot_sync(gmask, pos == j) >> gbase) & 0xFFFFu;
		const int pl = gbase + __ffs(pm) - 1;
		const double ajj = __shfl_sync(gmask, a[j], pl);
		const bool below = (l < 9) && (pos > j) && (ajj != 0.0);
		double aij = 0;
		if(below) { aij = a[j] / ajj; a[j] = aij; }
		#pragma unroll
		for(int k = j + 1; k < 9; k++) {
			const double ajk = __shfl_sync(gmask, a[k], pl);
			if(below) a[k] = a[k] - aij * ajk;
		}
	}
	// forward substitution, unit lower: rows in position order, subtractions in ascending j
	double x[9];
	#pragma unroll
	for(int i = 0; i < 9; i++) {
		double t = b;
		#pragma unroll
		for(int j = 0; j < i; j++) t -= a[j] * x[j];
		const unsigned pm = (__ballot_sync(gmask, pos == i) >> gbase) & 0xFFFFu;
		x[i] = __shfl_sync(gmask, t, gbase + __ffs(pm) - 1);
	}
	// back substitution, upper: position 8 down to 0, subtractions in ascending j
	float outv = 0;
	#pragma unroll
	for(int i = 8; i >= 0; i--) {
		double t = x[i];
		#pragma unroll
		for(int j = i + 1; j < 9; j++) t -= a[j] * x[j];
		t = t / a[i];
		const unsigned pm = (__ballot_sync(gmask, pos == i) >> gbase) & 0xFFFFu;
		x[i] = __shfl_sync(gmask, t, gbase + __ffs(pm) - 1);
		if(l == i) outv = (float)x[i];
	}
	if(l < 9) un[(size_t)node * GCM_REC + l] = outv;
}

#endif // __CUDACC__

} // namespace gcm
