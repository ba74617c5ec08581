// gcm_host.cpp -- see gcm_host.h.  Compiled with  g++ -O2 -ffp-contract=off -fopenmp.
#include "gcm_host.h"
#include "gcm_collide.cuh"
#include "gcm_math.cuh"
#include "gcm_inner.cuh"
#include "gcm_border.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace gcmh {

using gcm::Vec3;

void GlibcRand::seed(unsigned s)
{
	if(s == 0) s = 1;
	r[0] = s;
	long word = (long)(int32_t)s;
	for(int i = 1; i < 31; i++) {
		long hi = word / 127773;
		long lo = word % 127773;
		word = 16807 * lo - 2836 * hi;
		if(word < 0) word += 2147483647;
		r[i] = (uint32_t)word;
	}
	f = 3; b = 0;
	for(int i = 0; i < 310; i++) (void)next();
}

static inline Vec3 P(const std::vector<float>& c, int i)
{
	Vec3 v; v.x = c[3*(size_t)i]; v.y = c[3*(size_t)i+1]; v.z = c[3*(size_t)i+2];
	return v;
}

void HostMesh::translate(float dx, float dy, float dz)
{
	// TetrMesh::translate, gcm/meshtypes/TetrMesh.cpp:29-45
	for(int i = 0; i < N; i++) {
		coords[3*(size_t)i] += dx; coords[3*(size_t)i+1] += dy; coords[3*(size_t)i+2] += dz;
	}
	for(int k = 0; k < 3; k++) { outline_min[k] += (k == 0 ? dx : k == 1 ? dy : dz); outline_max[k] += (k == 0 ? dx : k == 1 ? dy : dz); }
}

// quick_math::solid_angle, gcm/system/quick_math.cpp:70-82 (atanf of the host libm, as the reference)
static float solid_angle(float x1, float y1, float z1, float x2, float y2, float z2, float x3, float y3, float z3)
{
	static const float PI = atanf(1) * 4;
	float det = fabsf( gcm::determinant(x1, y1, z1, x2, y2, z2, x3, y3, z3) );
	float norm_v1 = sqrtf( x1*x1 + y1*y1 + z1*z1 );
	float norm_v2 = sqrtf( x2*x2 + y2*y2 + z2*z2 );
	float norm_v3 = sqrtf( x3*x3 + y3*y3 + z3*z3 );
	float div = norm_v1 * norm_v2 * norm_v3 + (x1*x2 + y1*y2 + z1*z2) * norm_v3
					+ (x1*x3 + y1*y3 + z1*z3) * norm_v2 + (x2*x3 + y2*y3 + z2*z3) * norm_v1;
	float at = atanf( det / div );
	if (at < 0) { at += PI; }
	return 2 * at;
}

// find_border_elem_normal, gcm/meshtypes/TetrMesh_1stOrder.cpp:333-370
static void elem_normal(const Vec3& p1, const Vec3& p2, const Vec3& p3, float n[3])
{
	float v0x = p2.x - p1.x, v0y = p2.y - p1.y, v0z = p2.z - p1.z;
	float v1x = p3.x - p1.x, v1y = p3.y - p1.y, v1z = p3.z - p1.z;
	n[0] = v0y * v1z - v0z * v1y;
	n[1] = v0z * v1x - v0x * v1z;
	n[2] = v0x * v1y - v0y * v1x;
	float l = sqrtf( n[0] * n[0] + n[1] * n[1] + n[2] * n[2] );
	n[0] /= l;
	n[1] /= l;
	n[2] /= l;
}

int HostMesh::find_owner_tetr(int base, const float node_pos[3], float dx, float dy, float dz) const
{
	const Vec3 X = P(coords, base);
	const float R2 = dx * dx + dy * dy + dz * dz;
	const float delta = sqrtf(R2);
	std::vector<int> checked, checking, tmp;
	for(int e = n2t_off[base]; e < n2t_off[base+1]; e++) checking.push_back(n2t[e]);
	bool inside_R = true;
	while(inside_R) {
		inside_R = false;
		for(size_t i = 0; i < checking.size(); i++) {
			const int t = checking[i];
			const int* tv = &tets[4*(size_t)t];
			Vec3 p[4] = { P(coords, tv[0]), P(coords, tv[1]), P(coords, tv[2]), P(coords, tv[3]) };
			float h, vol;
			if(!tet_hv.empty()) { h = tet_hv[2*(size_t)t]; vol = tet_hv[2*(size_t)t+1]; }
			else { h = gcm::tetr_h(p[0], p[1], p[2], p[3]); vol = gcm::tet_signed_volume(p[0], p[1], p[2], p[3]); }
			if(h < 0) throw HostError{MESH_EXCEPTION, "Tetr volume is zero"};
			if(gcm::point_in_tetr(X, dx, dy, dz, delta, p[0], p[1], p[2], p[3], h, vol))
				return t;
			if(!inside_R) {
				for(int j = 0; j < 4; j++) {
					if(tv[j] == base) break;
					float lx = p[j].x, ly = p[j].y, lz = p[j].z;
					if( ( (node_pos[0] - lx) * (node_pos[0] - lx)
						+ (node_pos[1] - ly) * (node_pos[1] - ly)
						+ (node_pos[2] - lz) * (node_pos[2] - lz) ) < R2 )
						inside_R = true;
				}
			}
		}
		if(!inside_R) return -1;
		checked.insert(checked.end(), checking.begin(), checking.end());
		tmp.clear();
		for(size_t i = 0; i < checking.size(); i++)
			for(int j = 0; j < 4; j++) {
				int v = tets[4*(size_t)checking[i] + j];
				for(int e = n2t_off[v]; e < n2t_off[v+1]; e++) tmp.push_back(n2t[e]);
			}
		std::sort(tmp.begin(), tmp.end());
		tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
		std::vector<int> sorted_checked(checked);
		std::sort(sorted_checked.begin(), sorted_checked.end());
		checking.clear();
		for(size_t i = 0; i < tmp.size(); i++)
			if(!std::binary_search(sorted_checked.begin(), sorted_checked.end(), tmp[i]))
				checking.push_back(tmp[i]);
	}
	return -1;
}

void HostMesh::find_border_node_normal(int node, float h, float out[3]) const
{
	float fn[3] = {0, 0, 0};
	float cur[3];
	const int f0 = n2f_off[node], f1 = n2f_off[node+1];
	const int count = f1 - f0;
	for(int e = f0; e < f1; e++) {
		const int* fv = &faces[3*(size_t)n2f[e]];
		elem_normal(P(coords, fv[0]), P(coords, fv[1]), P(coords, fv[2]), cur);
		fn[0] += cur[0]; fn[1] += cur[1]; fn[2] += cur[2];
	}
	fn[0] /= count; fn[1] /= count; fn[2] /= count;
	float l = sqrtf( fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2] );
	fn[0] /= l; fn[1] /= l; fn[2] /= l;
	out[0] = fn[0]; out[1] = fn[1]; out[2] = fn[2];
	const float* X = &coords[3*(size_t)node];
	if(find_owner_tetr(node, X, - h * fn[0], - h * fn[1], - h * fn[2]) != -1)
		return;

	// last chance option (:414-439): direction from the centre of elements[0]
	{
		const int* tv = &tets[4*(size_t)n2t[n2t_off[node]]];
		for(int i = 0; i < 3; i++) {
			float ve = 0;
			for(int j = 0; j < 4; j++) ve += coords[3*(size_t)tv[j] + i];
			ve /= 4;
			fn[i] = X[i] - ve;
		}
	}
	l = sqrtf( fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2] );
	fn[0] /= l; fn[1] /= l; fn[2] /= l;
	out[0] = fn[0]; out[1] = fn[1]; out[2] = fn[2];
	if(find_owner_tetr(node, X, - h * fn[0], - h * fn[1], - h * fn[2]) != -1)
		return;

	// failback option (:441-478): min + max of the incident face normals
	float mn[3], mx[3];
	{
		const int* fv = &faces[3*(size_t)n2f[f0]];
		elem_normal(P(coords, fv[0]), P(coords, fv[1]), P(coords, fv[2]), cur);
		for(int i = 0; i < 3; i++) { mn[i] = cur[i]; mx[i] = cur[i]; }
	}
	for(int e = f0 + 1; e < f1; e++) {
		const int* fv = &faces[3*(size_t)n2f[e]];
		elem_normal(P(coords, fv[0]), P(coords, fv[1]), P(coords, fv[2]), cur);
		for(int i = 0; i < 3; i++) {
			if(mx[i] < cur[i]) mx[i] = cur[i];
			if(mn[i] > cur[i]) mn[i] = cur[i];
		}
	}
	for(int i = 0; i < 3; i++) fn[i] = mn[i] + mx[i];
	l = sqrtf( fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2] );
	fn[0] /= l; fn[1] /= l; fn[2] /= l;
	out[0] = fn[0]; out[1] = fn[1]; out[2] = fn[2];
	if(find_owner_tetr(node, X, - h * fn[0], - h * fn[1], - h * fn[2]) == -1)
		throw HostError{MESH_EXCEPTION, "Can not create reasonable normal for node - is the border too sharp?"};
}

void HostMesh::pre_process_mesh()
{
	faces.clear();
	tet_hv.assign(2*(size_t)T, 0.f);

	// node numbering checks of :86-102 are implicit (index == local_num by construction)
	for(size_t k = 0; k < tets.size(); k++)
		if(tets[k] < 0 || tets[k] >= N) throw HostError{MESH_EXCEPTION, "Invalid node numbering"};

	// ---- per-tet statics + get_max_h() over ALL tets (every node is still `used`, :81) -----
	int bad = 0;
	float mx_h = -1;
	#pragma omp parallel for schedule(static) reduction(max:mx_h) reduction(+:bad)
	for(int t = 0; t < T; t++) {
		const int* tv = &tets[4*(size_t)t];
		Vec3 p0 = P(coords, tv[0]), p1 = P(coords, tv[1]), p2 = P(coords, tv[2]), p3 = P(coords, tv[3]);
		float h = gcm::tetr_h(p0, p1, p2, p3);
		tet_hv[2*(size_t)t] = h;
		// Both uses of Vol are sign-blind (fabs(Vol) in point_in_tetr, fabs(s/Vol) in interpolate), so
		// |Vol| is stored and its sign bit carries a shape flag for the inner kernel's pre-filter:
		// positive = well shaped (6|Vol| > 0.02 * longest edge^3), negative = use the exact test only.
		{
			float vol = fabsf(gcm::tet_signed_volume(p0, p1, p2, p3));
			const Vec3 pp[4] = {p0, p1, p2, p3};
			float l2 = 0;
			for(int a = 0; a < 4; a++) for(int b = a + 1; b < 4; b++) {
				float dx = pp[a].x - pp[b].x, dy = pp[a].y - pp[b].y, dz = pp[a].z - pp[b].z;
				float d2 = dx*dx + dy*dy + dz*dz;
				if(d2 > l2) l2 = d2;
			}
			float l3 = l2 * sqrtf(l2);
			tet_hv[2*(size_t)t+1] = (6 * vol > 0.02f * l3) ? vol : -vol;
		}
		if(h < 0) bad++;
		if(h > mx_h) mx_h = h;
	}
	if(bad) throw HostError{MESH_EXCEPTION, "Tetr volume is zero"};
	max_h = mx_h;
	const float step_h = max_h / 4;

	// ---- volume reverse lookups (:104-118) ------------------------------------------------
	n2t_off.assign((size_t)N + 1, 0);
	for(size_t k = 0; k < tets.size(); k++) n2t_off[tets[k] + 1]++;
	for(int i = 0; i < N; i++) n2t_off[i+1] += n2t_off[i];
	n2t.resize(tets.size());
	{
		std::vector<int> pos(n2t_off.begin(), n2t_off.end() - 1);
		for(int t = 0; t < T; t++)
			for(int j = 0; j < 4; j++)
				n2t[pos[tets[4*(size_t)t + j]]++] = t;
	}

	// ---- unused nodes (:120-152) ----------------------------------------------------------
	for(int i = 0; i < N; i++)
		if(n2t_off[i+1] == n2t_off[i]) flags[i] &= ~4;
	{
		std::vector<char> drop(N, 0);
		#pragma omp parallel for schedule(static)
		for(int i = 0; i < N; i++) {
			if(!is_remote(i)) continue;
			int count = 0;
			for(int e = n2t_off[i]; e < n2t_off[i+1] && !count; e++)
				for(int k = 0; k < 4; k++)
					if(is_local(tets[4*(size_t)n2t[e] + k])) { count++; break; }
			if(count == 0) drop[i] = 1;
		}
		for(int i = 0; i < N; i++) if(drop[i]) flags[i] &= ~4;
	}

	// ---- border nodes by solid angle (:156-180) -------------------------------------------
	static const float PI = atanf(1) * 4;
	#pragma omp parallel for schedule(static)
	for(int i = 0; i < N; i++) {
		if(!is_local(i)) continue;
		float sa = 0;
		const Vec3 X = P(coords, i);
		for(int e = n2t_off[i]; e < n2t_off[i+1]; e++) {
			const int* tv = &tets[4*(size_t)n2t[e]];
			int vert[3], c = 0;
			for(int k = 0; k < 4; k++) if(tv[k] != i) { if(c < 3) vert[c] = tv[k]; c++; }
			if(c != 3) continue;   // degenerate tet: the reference aborts preprocessing here
			Vec3 a = P(coords, vert[0]), b = P(coords, vert[1]), d = P(coords, vert[2]);
			sa += solid_angle(a.x - X.x, a.y - X.y, a.z - X.z, b.x - X.x, b.y - X.y, b.z - X.z,
			                  d.x - X.x, d.y - X.y, d.z - X.z);
		}
		if( (4 * PI - sa) > PI / 100 ) flags[i] |= 8;
	}

	// ---- border triangles (:182-191, check_triangle_to_be_border :261-325) ------------------
	// The test of each (tet, face) is independent; the list order is tet-major, face-minor.
	static const int FV[4][4] = {{0,1,2,3},{0,1,3,2},{0,2,3,1},{1,2,3,0}};
	std::vector<int> cand;   // tet*4+face for faces whose three vertices are local border nodes
	for(int t = 0; t < T; t++) {
		const int* tv = &tets[4*(size_t)t];
		int nb = 0;
		for(int k = 0; k < 4; k++) nb += (is_border(tv[k]) && is_local(tv[k])) ? 1 : 0;
		if(nb < 3) continue;
		for(int f = 0; f < 4; f++) {
			int a = tv[FV[f][0]], b = tv[FV[f][1]], c = tv[FV[f][2]];
			if(is_border(a) && is_border(b) && is_border(c) && is_local(a) && is_local(b) && is_local(c))
				cand.push_back(t*4 + f);
		}
	}
	std::vector<int> cand_face(3*cand.size(), -1);
	std::vector<char> cand_ok(cand.size(), 0);
	std::string err;
	#pragma omp parallel for schedule(dynamic, 256)
	for(long ci = 0; ci < (long)cand.size(); ci++) {
		const int t = cand[ci] / 4, f = cand[ci] % 4;
		const int* tv = &tets[4*(size_t)t];
		int v1 = tv[FV[f][0]], v2 = tv[FV[f][1]], v3 = tv[FV[f][2]], tetr_vert = tv[FV[f][3]];
		float normal[3], dx[3], tri_center[3], dir[3], plane_move[3];
		elem_normal(P(coords, v1), P(coords, v2), P(coords, v3), normal);
		for(int i = 0; i < 3; i++)
			tri_center[i] = (coords[3*(size_t)v1+i] + coords[3*(size_t)v2+i] + coords[3*(size_t)v3+i]) / 3;
		for(int i = 0; i < 3; i++)
			dir[i] = coords[3*(size_t)tetr_vert+i] - tri_center[i];
		if(normal[0] * dir[0] + normal[1] * dir[1] + normal[2] * dir[2] > 0) {
			for(int i = 0; i < 3; i++) normal[i] = -normal[i];
			int s = v2; v2 = v3; v3 = s;
		}
		for(int i = 0; i < 3; i++) dx[i] = step_h * normal[i];
		for(int i = 0; i < 3; i++) plane_move[i] = tri_center[i] - coords[3*(size_t)v1+i];
		try {
			if(find_owner_tetr(v1, &coords[3*(size_t)v1], dx[0] + plane_move[0], dx[1] + plane_move[1], dx[2] + plane_move[2]) == -1) {
				cand_ok[ci] = 1;
				cand_face[3*ci] = v1; cand_face[3*ci+1] = v2; cand_face[3*ci+2] = v3;
			}
		} catch(HostError& e) {
			#pragma omp critical
			err = e.msg;
		}
	}
	if(!err.empty()) throw HostError{MESH_EXCEPTION, err};
	for(size_t ci = 0; ci < cand.size(); ci++)
		if(cand_ok[ci]) { faces.push_back(cand_face[3*ci]); faces.push_back(cand_face[3*ci+1]); faces.push_back(cand_face[3*ci+2]); }

	// ---- surface reverse lookups (:192-200) ------------------------------------------------
	const int F = (int)(faces.size() / 3);
	n2f_off.assign((size_t)N + 1, 0);
	for(size_t k = 0; k < faces.size(); k++) n2f_off[faces[k] + 1]++;
	for(int i = 0; i < N; i++) n2f_off[i+1] += n2f_off[i];
	n2f.resize(faces.size());
	{
		std::vector<int> pos(n2f_off.begin(), n2f_off.end() - 1);
		for(int f = 0; f < F; f++)
			for(int j = 0; j < 3; j++)
				n2f[pos[faces[3*(size_t)f + j]]++] = f;
	}

	// ---- node lists -------------------------------------------------------------------------
	border_nodes.clear(); inner_nodes.clear();
	border_slot.assign(N, -1);
	for(int i = 0; i < N; i++) {
		if(!is_local(i)) continue;
		if(is_border(i)) { border_slot[i] = (int)border_nodes.size(); border_nodes.push_back(i); }
		else inner_nodes.push_back(i);
	}
	const int NB = (int)border_nodes.size();

	// ---- outer-normal check of :202-234 (only its exceptions are observable) ---------------
	#pragma omp parallel for schedule(dynamic, 256)
	for(int s = 0; s < NB; s++) {
		float n[3];
		try { find_border_node_normal(border_nodes[s], step_h, n); }
		catch(HostError& e) {
			#pragma omp critical
			err = e.msg;
		}
	}
	if(!err.empty()) throw HostError{MESH_EXCEPTION, err};

	// ---- outline (:238-254) -------------------------------------------------------------------
	if(N) {
		for(int j = 0; j < 3; j++) outline_min[j] = outline_max[j] = coords[j];
		for(int i = 0; i < N; i++)
			for(int j = 0; j < 3; j++) {
				float c = coords[3*(size_t)i + j];
				if(c > outline_max[j]) outline_max[j] = c;
				if(c < outline_min[j]) outline_min[j] = c;
			}
	}

	// ---- get_min_h() over tets whose vertices are all used (:1040-1065) ----------------------
	float mn_h = -1;
	for(int t = 0; t < T; t++) {
		const int* tv = &tets[4*(size_t)t];
		if(!is_used(tv[0]) || !is_used(tv[1]) || !is_used(tv[2]) || !is_used(tv[3])) continue;
		float h = tet_hv[2*(size_t)t];
		if(mn_h < 0) mn_h = h;
		if(h < mn_h) mn_h = h;
	}
	min_h = mn_h;

	// ---- static node normals: find_border_node_normal(i) == (i, get_min_h()) (:377-380) -------
	normals.assign(3*(size_t)NB, 0.f);
	#pragma omp parallel for schedule(dynamic, 256)
	for(int s = 0; s < NB; s++) {
		try { find_border_node_normal(border_nodes[s], min_h, &normals[3*(size_t)s]); }
		catch(HostError& e) {
			#pragma omp critical
			err = e.msg;
		}
	}
	if(!err.empty()) throw HostError{MESH_EXCEPTION, err};
	basis.assign(9*(size_t)NB, 0.f);
	cond_id.assign(NB, -1);
	// launch order (scheduling only): a node whose incident border faces are not coplanar sits on
	// an edge or corner and typically burns all 11 alpha tries on the tangential axes
	{
		std::vector<char> sharp(NB, 0);
		#pragma omp parallel for schedule(static)
		for(int s = 0; s < NB; s++) {
			const int i = border_nodes[s];
			const float* nn = &normals[3*(size_t)s];
			for(int e = n2f_off[i]; e < n2f_off[i+1]; e++) {
				const int* fv = &faces[3*(size_t)n2f[e]];
				float fnr[3];
				elem_normal(P(coords, fv[0]), P(coords, fv[1]), P(coords, fv[2]), fnr);
				if(!(fnr[0]*nn[0] + fnr[1]*nn[1] + fnr[2]*nn[2] > 0.999f)) sharp[s] = 1;
			}
		}
		// flat nodes are grouped by the direction of their normal (26 sectors), node order inside a group:
		// the 32 nodes of a warp then walk candidate lists of the same shape and take the same branches
		// (on a box every face is one group; nothing but the launch order depends on this)
		auto sector = [&](int s) {
			const float* nn = &normals[3*(size_t)s];
			int code = 0;
			for(int k = 0; k < 3; k++) code = code * 3 + (nn[k] > 0.38f ? 2 : (nn[k] < -0.38f ? 0 : 1));
			return code;
		};
		border_launch.clear();
		for(int s = 0; s < NB; s++) if(sharp[s]) border_launch.push_back(border_nodes[s]);
		n_sharp = (int)border_launch.size();
		std::vector<std::pair<int, int> > keyed;
		for(int s = 0; s < NB; s++) if(!sharp[s]) keyed.push_back(std::make_pair(sector(s), border_nodes[s]));
		std::sort(keyed.begin(), keyed.end());
		for(size_t q = 0; q < keyed.size(); q++) border_launch.push_back(keyed[q].second);
	}
	build_fast_path_tables();
	preprocessed = true;
}

// Static inputs of gcm_inner.cuh.  Everything here is computed with the general routines, so
// the specialised kernel reproduces them by construction; whatever does not fit its
// assumptions is routed to the generic kernel (inner_generic).
void HostMesh::build_fast_path_tables()
{
	// ---- distinct materials of local nodes and their identity-basis eigensystems -----------
	MaterialRegistry& reg = registry ? *registry : own_registry;
	std::vector<float>& materials = reg.materials; std::vector<float>& mat_table = reg.table; std::vector<char>& mat_ok = reg.ok;
	mat_id.assign(N, 255);
	for(int i = 0; i < N; i++) {
		if(!is_local(i)) continue;
		const float* mt = &mat[4*(size_t)i];
		int id = -1;
		for(size_t k = 0; k < materials.size() / 3; k++)
			if(memcmp(&materials[3*k], mt, 3 * sizeof(float)) == 0) { id = (int)k; break; }
		if(id < 0) {
			if(materials.size() / 3 >= GCM_MAX_MATERIALS) continue;   // not tabulated -> generic kernel
			id = (int)(materials.size() / 3);
			materials.insert(materials.end(), mt, mt + 3);
			bool ok = true;
			for(int st = 0; st < 3; st++) {
				float L[9], U[81], U1[81];
				float q[3] = {0, 0, 0}; q[st] = 1;
				if(gcm::elastic_matrix(mt[0], mt[1], mt[2], q[0], q[1], q[2], L, U, U1) != 0) ok = false;
				for(int r = 0; r < 9; r++)
					for(int c = 0; c < 9; c++) {
						if(!((gcm::u_pattern(st, r) >> c) & 1) && U[r*9+c] != 0) ok = false;
						if(!((gcm::u1_pattern(st, r) >> c) & 1) && U1[r*9+c] != 0) ok = false;
					}
				// the five-point pattern (0,1,2,3,2,3,4,4,4) of the characteristic dedupe (:479-490)
				float dksi[9]; int pp[9], count = 0;
				for(int r = 0; r < 9; r++) dksi[r] = - L[r] * 0.002f;
				for(int r = 0; r < 9; r++) {
					bool found = false;
					for(int c = 0; c < r; c++)
						if( (double)fabsf(dksi[r] - dksi[c]) <= 0.01 * (double)fabsf(dksi[r] + dksi[c]) ) { found = true; pp[r] = pp[c]; }
					if(!found) pp[r] = count++;
				}
				static const int want[9] = {0, 1, 2, 3, 2, 3, 4, 4, 4};
				if(count != 5 || memcmp(pp, want, sizeof want) != 0 || L[2] != L[4] || L[3] != L[5] || L[6] != 0) ok = false;
				mat_table.insert(mat_table.end(), L, L + 9);
				mat_table.insert(mat_table.end(), U, U + 81);
				mat_table.insert(mat_table.end(), U1, U1 + 81);
			}
			mat_ok.push_back(ok ? 1 : 0);
		}
		mat_id[i] = mat_ok[id] ? (unsigned char)id : 255;
	}

	// ---- zero-speed foot: owner must be elements[0], all weight on the node itself ------------
	w0.assign(N, 0.f);
	std::vector<char> fast(N, 0);
	const int NI = (int)inner_nodes.size();
	#pragma omp parallel for schedule(static)
	for(int q = 0; q < NI; q++) {
		const int i = inner_nodes[q];
		if(mat_id[i] == 255 || n2t_off[i+1] == n2t_off[i]) continue;
		const Vec3 X = P(coords, i);
		const int t = n2t[n2t_off[i]];
		const int* tv = &tets[4*(size_t)t];
		const Vec3 p[4] = { P(coords, tv[0]), P(coords, tv[1]), P(coords, tv[2]), P(coords, tv[3]) };
		const float h = tet_hv[2*(size_t)t], vol = tet_hv[2*(size_t)t+1];
		if(!gcm::point_in_tetr(X, 0.f, 0.f, 0.f, 0.f, p[0], p[1], p[2], p[3], h, vol)) continue;
		float f[4];
		if(!gcm::interp_factors(p[0], p[1], p[2], p[3], X.x, X.y, X.z, vol, f)) continue;
		int k = -1; bool ok = true;
		for(int j = 0; j < 4; j++) { if(tv[j] == i) k = j; else if(f[j] != 0) ok = false; }
		if(k < 0 || !ok) continue;
		w0[i] = f[k];
		fast[i] = 1;
	}
	inner_fast.clear(); inner_generic.clear();
	for(int q = 0; q < NI; q++) (fast[inner_nodes[q]] ? inner_fast : inner_generic).push_back(inner_nodes[q]);

	// ---- node -> tet CSR expanded with the other three vertices (removes the tets[t] hop) ------
	if(T >= (1 << 30)) throw HostError{CONFIG_EXCEPTION, "too many tetrahedrons for the packed candidate table"};
	n2x.assign(4 * n2t.size(), 0);
	#pragma omp parallel for schedule(static)
	for(int i = 0; i < N; i++)
		for(int e = n2t_off[i]; e < n2t_off[i+1]; e++) {
			const int t = n2t[e];
			const int* tv = &tets[4*(size_t)t];
			int k = -1, o[3], c = 0;
			for(int j = 0; j < 4; j++) { if(tv[j] == i && k < 0) k = j; else if(c < 3) o[c++] = tv[j]; }
			if(k < 0) k = 3;
			while(c < 3) o[c++] = i;
			n2x[4*(size_t)e] = o[0]; n2x[4*(size_t)e+1] = o[1]; n2x[4*(size_t)e+2] = o[2];
			n2x[4*(size_t)e+3] = (int)((unsigned)t | ((unsigned)k << 30));
		}
}

// The reference evaluates point_in_tetr at P = X + d rounded to float: each coordinate of P is off by up
// to ulp(|coordinate|)/2, i.e. each barycentric coordinate by up to  n = sqrt(3) ulp(Cmax) / (2 min h).
// The predicate passes iff the negative barycentric coordinates sum to less than 5e-4 (sum|s| < 1.001
// Vol); with at most three of them negative, "all > -t_p" is certain to pass while 3 t_p + 3 n + 1e-5 <
// 5e-4, and "one < -t_f" certain to fail once t_f > 5e-4 + 3 n + 1e-5.  2e-5 of each is left to the
// filter's own float rounding.  At the benchmark scale (|coordinates| < 256, h = 0.577) this gives the
// 1e-4 / 1e-3 the filter was validated with; a body far from the origin gets tighter values and, when
// no positive t_p is left (|coordinate| / h beyond ~2000), no pre-filter at all: every candidate through
// the literal predicate.
void prefilter_thresholds(const std::vector<const HostMesh*>& meshes, float* pf_fail, float* pf_pass, int* exact_only)
{
	float cmax = 0, hmin = -1;
	for(const HostMesh* hm : meshes) {
		for(size_t k = 0; k < hm->coords.size(); k++) cmax = fmaxf(cmax, fabsf(hm->coords[k]));
		if(hm->min_h > 0 && (hmin < 0 || hm->min_h < hmin)) hmin = hm->min_h;
	}
	*pf_fail = -1e-3f; *pf_pass = -1e-4f; *exact_only = 0;
	if(!(hmin > 0) || !std::isfinite(cmax)) { *exact_only = 1; return; }
	cmax += 2 * hmin;                                         // the tested point is up to 0.99 h away from a node
	const float ulp = nextafterf(cmax, INFINITY) - cmax;
	const float n = 0.8660254f * ulp / hmin;
	const float tp = fminf(1e-4f, (4.9e-4f - 3 * n) / 3 * 0.9f - 2e-5f);
	const float tf = fmaxf(1e-3f, 1.1f * (5.1e-4f + 3 * n) + 2e-5f);
	if(!(tp >= 2e-5f)) { *exact_only = 1; return; }
	*pf_pass = -tp; *pf_fail = -tf;
}

// Static inputs of gcm_border.cuh: per flat border node its candidate list with the barycentric
// gradients of the other three vertices scaled by 0.99 h(tet) (the point point_in_tetr tests for a
// direction q^ is X + 0.99 h q^, so a candidate's barycentric coordinates there are q^ . G), computed in
// double from the float coordinates, plus what the search needs to know when it finds nothing.
void build_border_flat_tables(const std::vector<BorderFlatZone>& zones, std::vector<int>& slot_static8, std::vector<float>& cand_quads)
{
	size_t NB = 0, NF = 0;
	for(const auto& z : zones) { NB += z.hm->border_nodes.size(); NF += z.hm->border_launch.size(); }
	slot_static8.assign(8 * NB, 0);
	gcm::BndSlotStatic* st = reinterpret_cast<gcm::BndSlotStatic*>(slot_static8.data());
	static_assert(sizeof(gcm::BndSlotStatic) == 32, "BndSlotStatic is 8 words");
	// border nodes in launch order
	struct Ref { const BorderFlatZone* z; int node; };
	std::vector<Ref> flat; flat.reserve(NF);
	for(const auto& z : zones)
		for(size_t q = 0; q < z.hm->border_launch.size(); q++) flat.push_back(Ref{&z, z.hm->border_launch[q]});
	const size_t NW = (NF + GCM_BND_LANES - 1) / GCM_BND_LANES;
	std::vector<size_t> warp_off(NW + 1, 0);
	for(size_t w = 0; w < NW; w++) {
		int dmax = 0;
		for(size_t p = w * GCM_BND_LANES; p < std::min(NF, (w + 1) * GCM_BND_LANES); p++) {
			const HostMesh& hm = *flat[p].z->hm; const int i = flat[p].node;
			dmax = std::max(dmax, hm.n2t_off[i+1] - hm.n2t_off[i]);
		}
		warp_off[w+1] = warp_off[w] + (size_t)3 * dmax * GCM_BND_LANES;
	}
	if(warp_off[NW] >= 0x7fffffffull) throw HostError{CONFIG_EXCEPTION, "border candidate table too large for 32-bit offsets"};
	cand_quads.assign(4 * (warp_off[NW] + (size_t)3 * 4 * GCM_BND_LANES), 0.f);   // + 4 candidates of padding: the kernel loads them four at a time
	#pragma omp parallel for schedule(dynamic, 256)
	for(long p = 0; p < (long)NF; p++) {
		const BorderFlatZone& z = *flat[p].z;
		const HostMesh& hm = *z.hm;
		const int i = flat[p].node;
		gcm::BndSlotStatic& s = st[z.slot_off + hm.border_slot[i]];
		const int e0 = hm.n2t_off[i], e1 = hm.n2t_off[i+1];
		const Vec3 X = P(hm.coords, i);
		// zero-speed foot: elements[0] must pass the literal predicate at the node itself with all weight on it
		{
			const int t = hm.n2t[e0];
			const int* tv = &hm.tets[4*(size_t)t];
			const Vec3 pp[4] = { P(hm.coords, tv[0]), P(hm.coords, tv[1]), P(hm.coords, tv[2]), P(hm.coords, tv[3]) };
			const float h = hm.tet_hv[2*(size_t)t], vol = hm.tet_hv[2*(size_t)t+1];
			if(!gcm::point_in_tetr(X, 0.f, 0.f, 0.f, 0.f, pp[0], pp[1], pp[2], pp[3], h, vol)) continue;
			float f[4];
			if(!gcm::interp_factors(pp[0], pp[1], pp[2], pp[3], X.x, X.y, X.z, vol, f)) continue;
			int k = -1; bool ok = true;
			for(int j = 0; j < 4; j++) { if(tv[j] == i && k < 0) k = j; else if(f[j] != 0) ok = false; }
			if(k < 0 || !ok) continue;
			s.w0 = f[k];
		}
		const size_t base = warp_off[p / GCM_BND_LANES] + (size_t)(p % GCM_BND_LANES);
		float hmin = 3.4e38f, dmin2 = 3.4e38f;
		for(int e = e0; e < e1; e++) {
			const int c = e - e0;
			const int* x = &hm.n2x[4*(size_t)e];
			const int t = (int)((unsigned)x[3] & 0x3fffffffu), k = (int)((unsigned)x[3] >> 30);
			const Vec3 A = P(hm.coords, x[0]), B = P(hm.coords, x[1]), C = P(hm.coords, x[2]);
			const float h = hm.tet_hv[2*(size_t)t], vflag = hm.tet_hv[2*(size_t)t+1];
			hmin = std::min(hmin, h);
			// the reference measures only the vertices BEFORE the node in tet order (`break`, :1544-1545)
			for(int j = 0; j < 3 && j < k; j++) {
				const Vec3& Q = j == 0 ? A : (j == 1 ? B : C);
				const float d2 = (X.x - Q.x) * (X.x - Q.x) + (X.y - Q.y) * (X.y - Q.y) + (X.z - Q.z) * (X.z - Q.z);
				dmin2 = std::min(dmin2, d2);
			}
			const double ea[3] = {(double)A.x - X.x, (double)A.y - X.y, (double)A.z - X.z};
			const double eb[3] = {(double)B.x - X.x, (double)B.y - X.y, (double)B.z - X.z};
			const double ec[3] = {(double)C.x - X.x, (double)C.y - X.y, (double)C.z - X.z};
			const double bc[3] = {eb[1]*ec[2] - eb[2]*ec[1], eb[2]*ec[0] - eb[0]*ec[2], eb[0]*ec[1] - eb[1]*ec[0]};
			const double ca[3] = {ec[1]*ea[2] - ec[2]*ea[1], ec[2]*ea[0] - ec[0]*ea[2], ec[0]*ea[1] - ec[1]*ea[0]};
			const double ab[3] = {ea[1]*eb[2] - ea[2]*eb[1], ea[2]*eb[0] - ea[0]*eb[2], ea[0]*eb[1] - ea[1]*eb[0]};
			const double D = ea[0]*bc[0] + ea[1]*bc[1] + ea[2]*bc[2];
			// the node must be a vertex of the candidate (it is: the list is the node's own) and the tet well shaped
			const bool usable = vflag > 0 && D != 0 && x[0] != i && x[1] != i && x[2] != i;
			const double sc = usable ? 0.99 * (double)h / D : 0.0;
			float* q = &cand_quads[4 * (base + (size_t)(3 * c) * GCM_BND_LANES)];
			float* q1 = q + 4 * GCM_BND_LANES; float* q2 = q1 + 4 * GCM_BND_LANES;
			q[0] = (float)(sc * bc[0]); q[1] = (float)(sc * bc[1]); q[2] = (float)(sc * bc[2]); q[3] = (float)(sc * ca[0]);
			q1[0] = (float)(sc * ca[1]); q1[1] = (float)(sc * ca[2]); q1[2] = (float)(sc * ab[0]); q1[3] = (float)(sc * ab[1]);
			q2[0] = (float)(sc * ab[2]); q2[1] = usable ? 1.f : 0.f; q2[2] = 0.f; q2[3] = 0.f;
		}
		s.deg = e1 - e0;
		s.e0 = (int)(z.entry_off + e0);
		s.hmin = hmin; s.dmin2 = dmin2;
		s.h0 = hm.tet_hv[2*(size_t)hm.n2t[e0]];
		s.cand_off = (int)base;
	}
}

// create_rotation_matrix_with_normal + the rand() bookkeeping of create_random_axis,
// gcm/methods/TetrMethod_Plastic_1stOrder.cpp:697-748, :843-901.
void HostMesh::create_random_axes(GlibcRand& rng)
{
	const double PI = atan(1) * 4;
	for(int i = 0; i < N; i++) {
		if(!is_local(i)) continue;
		if(!is_border(i)) {
			(void)rng.next(); (void)rng.next(); (void)rng.next();   // phi, psi, teta; basis := E
			continue;
		}
		const int s = border_slot[i];
		const float x = normals[3*(size_t)s], y = normals[3*(size_t)s+1], z = normals[3*(size_t)s+2];
		float teta = (rng.next() % 360) * 2 * PI / 360;
		gcm::rotation_basis_with_normal(x, y, z, cosf(teta), sinf(teta), &basis[9*(size_t)s]);
	}
}

// ---- the rand() stream as a linear recurrence ----------------------------------------------
// o[n] = o[n-31] + o[n-3] (mod 2^32), rand() = o[n] >> 1.  State in stream order
// W = (o[n-31] .. o[n-1]); one call maps W -> (W[1] .. W[30], W[0] + W[28]).
void GlibcRand::get_window(uint32_t w[31]) const { for(int i = 0; i < 31; i++) w[i] = r[(f + i) % 31]; }
void GlibcRand::set_window(const uint32_t w[31]) { for(int i = 0; i < 31; i++) r[i] = w[i]; f = 0; b = 28; }

static void mat_mul(const uint32_t* a, const uint32_t* b, uint32_t* c)
{
	for(int i = 0; i < 31; i++)
		for(int j = 0; j < 31; j++) {
			uint32_t acc = 0;
			for(int k = 0; k < 31; k++) acc += a[i*31+k] * b[k*31+j];
			c[i*31+j] = acc;
		}
}

void rng_jump_matrix(unsigned long long k, uint32_t out[961])
{
	std::vector<uint32_t> base(961, 0), res(961, 0), tmp(961);
	for(int i = 0; i < 30; i++) base[i*31 + i + 1] = 1;
	base[30*31 + 0] = 1; base[30*31 + 28] = 1;
	for(int i = 0; i < 31; i++) res[i*31+i] = 1;
	while(k) {
		if(k & 1) { mat_mul(base.data(), res.data(), tmp.data()); res = tmp; }
		mat_mul(base.data(), base.data(), tmp.data()); base = tmp;
		k >>= 1;
	}
	memcpy(out, res.data(), 961 * sizeof(uint32_t));
}

void rng_apply_jump(const uint32_t m[961], uint32_t w[31])
{
	uint32_t n[31];
	for(int i = 0; i < 31; i++) {
		uint32_t acc = 0;
		for(int j = 0; j < 31; j++) acc += m[i*31+j] * w[j];
		n[i] = acc;
	}
	memcpy(w, n, sizeof n);
}

void rng_teta_tables(float cos_t[360], float sin_t[360])
{
	const double PI = atan(1) * 4;
	for(int k = 0; k < 360; k++) {
		float teta = k * 2 * PI / 360;   // (rand() % 360) * 2 * PI / 360, TetrMethod_Plastic_1stOrder.cpp:729
		cos_t[k] = cosf(teta);
		sin_t[k] = sinf(teta);
	}
}

// ---- collision discovery (host; the per-step contact kernel consumes its output) ------------
// TetrMesh_1stOrder::vector_intersects_triangle / interpolate_triangle (gcm/meshtypes/TetrMesh_1stOrder.cpp:2194-2311):
// the arithmetic is gcm_collide.cuh's (shared with the device detector); here its error codes become the
// reference's exceptions.
static bool vector_intersects_triangle(const float* p1, const float* p2, const float* p3, const float* p0,
                                       const float* v, float /*l*/, float* p)
{
	static const char* what[6] = {"NAN or INF in p1", "NAN or INF in p2", "NAN or INF in p3", "NAN or INF in p0", "NAN or INF in v", "NAN or INF in n"};
	const int r = gcm::vector_intersects_triangle(p1, p2, p3, p0, v, p);
	if(r < 0) throw HostError{MESH_EXCEPTION, what[-r - 1]};
	return r != 0;
}

static void interpolate_triangle(const float* p1, const float* p2, const float* p3, const float* p,
                                 const float* v1, const float* v2, const float* v3, float* v, int n)
{
	if(!gcm::interpolate_triangle(p1, p2, p3, p, v1, v2, v3, v, n))
		throw HostError{MESH_EXCEPTION, "NAN or INF in virt node coords"};
}

// 2-D bins over the faces that take part in one collision pass (see find_collisions).  Only a
// conservative filter: whatever it returns is still put through vector_intersects_triangle in list
// order, so the result is that of the full scan.
struct FaceGrid {
	bool ok = false;               // false: few faces / degenerate / non-finite input -> scan every face
	int ax0 = 0, ax1 = 1;          // the two binned coordinate axes (longest sides of the faces' box)
	double org[2] = {0, 0}, cell = 1, pad = 0;
	int n[2] = {0, 0};
	std::vector<int> start, items; // CSR: bin -> positions in fcs, ascending

	void build(const HostMesh& m2, const std::vector<int>& fcs)
	{
		ok = false;
		const size_t F = fcs.size();
		if(F < 64) return;
		if(const char* e = getenv("GCM_B200_FACE_GRID")) if(atoi(e) == 0) return;   // A/B switch for the tests
		double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, mean = 0;
		for(size_t l = 0; l < F; l++) {
			double flo[3] = {1e300, 1e300, 1e300}, fhi[3] = {-1e300, -1e300, -1e300};
			for(int j = 0; j < 3; j++) {
				const float* c = &m2.coords[3*(size_t)m2.faces[3*(size_t)fcs[l] + j]];
				for(int a = 0; a < 3; a++) {
					if(!std::isfinite(c[a])) return;
					flo[a] = std::min(flo[a], (double)c[a]); fhi[a] = std::max(fhi[a], (double)c[a]);
				}
			}
			for(int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], flo[a]); hi[a] = std::max(hi[a], fhi[a]); }
			mean += std::max(fhi[0] - flo[0], std::max(fhi[1] - flo[1], fhi[2] - flo[2]));
		}
		mean /= (double)F;
		// drop the shortest side of the box
		int drop = 0;
		for(int a = 1; a < 3; a++) if(hi[a] - lo[a] < hi[drop] - lo[drop]) drop = a;
		ax0 = drop == 0 ? 1 : 0; ax1 = drop == 2 ? 1 : 2;
		const double e0 = hi[ax0] - lo[ax0], e1 = hi[ax1] - lo[ax1];
		if(!(mean > 0) || !(e0 > 0) || !(e1 > 0)) return;
		cell = std::max(mean, std::max(e0, e1) / 2048.0);
		// a meeting point passes the area test up to ~1e-4 of the face size outside the triangle (plus
		// float rounding of the point itself): bins overlap by 100x that
		double big = 0;
		for(int a = 0; a < 3; a++) big = std::max(big, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
		pad = 1e-2 * cell + 1e-5 * big;
		org[0] = lo[ax0] - 2 * pad; org[1] = lo[ax1] - 2 * pad;
		n[0] = (int)((e0 + 4 * pad) / cell) + 1; n[1] = (int)((e1 + 4 * pad) / cell) + 1;
		std::vector<int> count((size_t)n[0] * n[1] + 1, 0);
		auto range = [&](size_t l, int r[4]) {
			double flo[2] = {1e300, 1e300}, fhi[2] = {-1e300, -1e300};
			for(int j = 0; j < 3; j++) {
				const float* c = &m2.coords[3*(size_t)m2.faces[3*(size_t)fcs[l] + j]];
				flo[0] = std::min(flo[0], (double)c[ax0]); fhi[0] = std::max(fhi[0], (double)c[ax0]);
				flo[1] = std::min(flo[1], (double)c[ax1]); fhi[1] = std::max(fhi[1], (double)c[ax1]);
			}
			for(int d = 0; d < 2; d++) {
				r[2*d] = std::max(0, std::min(n[d] - 1, (int)std::floor((flo[d] - pad - org[d]) / cell)));
				r[2*d+1] = std::max(0, std::min(n[d] - 1, (int)std::floor((fhi[d] + pad - org[d]) / cell)));
			}
		};
		for(size_t l = 0; l < F; l++) {
			int r[4]; range(l, r);
			for(int i = r[0]; i <= r[1]; i++) for(int j = r[2]; j <= r[3]; j++) count[(size_t)i * n[1] + j + 1]++;
		}
		start.assign(count.size(), 0);
		for(size_t b = 1; b < count.size(); b++) start[b] = start[b-1] + count[b];
		items.assign((size_t)start.back(), 0);
		std::vector<int> fill(start.begin(), start.end() - 1);
		for(size_t l = 0; l < F; l++) {      // ascending l keeps every bin's list ascending
			int r[4]; range(l, r);
			for(int i = r[0]; i <= r[1]; i++) for(int j = r[2]; j <= r[3]; j++) items[(size_t)fill[(size_t)i * n[1] + j]++] = (int)l;
		}
		ok = true;
	}

	// Positions in fcs of every face whose bins the line p0 + t v (all t) crosses, ascending.
	// Returns false when the caller has to scan every face (no grid, or non-finite node data: the
	// full scan then raises the reference's own exception).
	bool candidates(const float* p0, const float* v, std::vector<int>& out) const
	{
		if(!ok) return false;
		for(int a = 0; a < 3; a++) if(!std::isfinite(p0[a]) || !std::isfinite(v[a])) return false;
		const double p[2] = {p0[ax0], p0[ax1]}, d[2] = {v[ax0], v[ax1]};
		out.clear();
		auto add = [&](int i, int j0, int j1, bool swapped) {
			for(int j = j0; j <= j1; j++) {
				const size_t b = swapped ? (size_t)j * n[1] + i : (size_t)i * n[1] + j;
				out.insert(out.end(), items.begin() + start[b], items.begin() + start[b+1]);
			}
		};
		const int M = std::fabs(d[0]) >= std::fabs(d[1]) ? 0 : 1;   // major axis of the projected line
		const int m = 1 - M;
		auto clampi = [&](double x, int dim) { return std::max(0, std::min(n[dim] - 1, (int)std::floor((x - org[dim]) / cell))); };
		if(d[M] == 0) {
			// the line is normal to the binned plane: one point
			const int i0 = clampi(p[M] - pad, M), i1 = clampi(p[M] + pad, M);
			if(p[M] + pad < org[M] || p[M] - pad > org[M] + n[M] * cell || p[m] + pad < org[m] || p[m] - pad > org[m] + n[m] * cell) return true;
			for(int i = i0; i <= i1; i++) add(M == 0 ? i : i, clampi(p[m] - pad, m), clampi(p[m] + pad, m), M == 1);
		} else {
			const double slope = d[m] / d[M];                        // |slope| <= 1
			for(int i = 0; i < n[M]; i++) {
				const double x0 = org[M] + i * cell - pad, x1 = org[M] + (i + 1) * cell + pad;
				const double y0 = p[m] + slope * (x0 - p[M]), y1 = p[m] + slope * (x1 - p[M]);
				const double ylo = std::min(y0, y1) - pad, yhi = std::max(y0, y1) + pad;
				if(yhi < org[m] || ylo > org[m] + n[m] * cell) continue;
				add(i, clampi(ylo, m), clampi(yhi, m), M == 1);
			}
		}
		std::sort(out.begin(), out.end());
		out.erase(std::unique(out.begin(), out.end()), out.end());
		return true;
	}
};

void find_collisions(std::vector<HostMesh*>& meshes, float threshold, std::vector<VirtNodeHost>& virt,
                     std::vector<std::vector<int> >& axis_plus0)
{
	const int M = (int)meshes.size();
	virt.clear();
	axis_plus0.assign(M, std::vector<int>());
	for(int a = 0; a < M; a++) {
		HostMesh& hm = *meshes[a];
		axis_plus0[a].assign(hm.border_nodes.size(), -1);
		for(int i = 0; i < hm.N; i++) hm.flags[i] &= ~1;   // clear_contact_state: setContactType(Free)
	}
	for(int a = 0; a < M; a++)
		for(int b = 0; b < M; b++) {
			if(a == b) continue;
			HostMesh& m1 = *meshes[a];
			HostMesh& m2 = *meshes[b];
			// CollisionDetector::find_intersection, gcm/system/CollisionDetector.cpp:34-47
			float lo[3], hi[3];
			bool hit = true;
			for(int j = 0; j < 3; j++) {
				lo[j] = fmaxf(m1.outline_min[j] - threshold, m2.outline_min[j] - threshold);
				hi[j] = fminf(m1.outline_max[j] + threshold, m2.outline_max[j] + threshold);
				if(lo[j] > hi[j]) { hit = false; break; }
			}
			if(!hit) continue;
			// find_nodes_in_intersection (:49-66): local border nodes inside the box, index order
			std::vector<int> nodes;
			for(size_t s = 0; s < m1.border_nodes.size(); s++) {
				const float* c = &m1.coords[3*(size_t)m1.border_nodes[s]];
				if(c[0] < lo[0] || c[0] > hi[0] || c[1] < lo[1] || c[1] > hi[1] || c[2] < lo[2] || c[2] > hi[2]) continue;
				nodes.push_back((int)s);
			}
			// find_faces_in_intersection (:87-103): faces with at least one vertex inside the box
			std::vector<int> fcs;
			const int F = (int)(m2.faces.size() / 3);
			for(int f = 0; f < F; f++) {
				int out = 0;
				for(int j = 0; j < 3; j++) {
					const float* c = &m2.coords[3*(size_t)m2.faces[3*(size_t)f + j]];
					if(c[0] < lo[0] || c[0] > hi[0] || c[1] < lo[1] || c[1] > hi[1] || c[2] < lo[2] || c[2] > hi[2]) out++;
				}
				if(out != 3) fcs.push_back(f);
			}
			// each node against the faces in order, first hit wins (:44-84); the node loop is parallel,
			// the virt_nodes numbering is then assigned in node order as the serial loop would
			const int nn = (int)nodes.size();
			std::vector<int> hit_face(nn, -1);
			std::vector<float> hit_p(3*(size_t)nn, 0.f);
			std::string err;
			// The test a face has to pass is "the LINE through the node along its first basis vector meets the
			// face's plane inside the triangle" (no distance limit survives in the reference, see above), and
			// the first face of the list that passes wins.  A face can only pass if the meeting point lies in
			// its bounding box, so the faces are binned on a 2-D grid over the two longest sides of the
			// intersection box and a node looks only at the bins its (projected) line crosses -- in list order,
			// which gives the same face as the full scan, at O(bins crossed) instead of O(faces) per node.
			FaceGrid grid;
			grid.build(m2, fcs);
			#pragma omp parallel for schedule(dynamic, 64)
			for(int k = 0; k < nn; k++) {
				const int s = nodes[k];
				const float* p0 = &m1.coords[3*(size_t)m1.border_nodes[s]];
				const float* dir = &m1.basis[9*(size_t)s];     // local_basis->ksi[0]
				try {
					std::vector<int> cand;       // positions in fcs, ascending
					const bool all = !grid.candidates(p0, dir, cand);
					const size_t nc = all ? fcs.size() : cand.size();
					for(size_t c = 0; c < nc; c++) {
						const size_t l = all ? c : (size_t)cand[c];
						const int* fv = &m2.faces[3*(size_t)fcs[l]];
						float p[3];
						if(vector_intersects_triangle(&m2.coords[3*(size_t)fv[0]], &m2.coords[3*(size_t)fv[1]], &m2.coords[3*(size_t)fv[2]],
						                              p0, dir, threshold, p)) {
							hit_face[k] = fcs[l];
							hit_p[3*(size_t)k] = p[0]; hit_p[3*(size_t)k+1] = p[1]; hit_p[3*(size_t)k+2] = p[2];
							break;
						}
					}
				} catch(HostError& e) {
					#pragma omp critical
					err = e.msg;
				}
			}
			if(!err.empty()) throw HostError{MESH_EXCEPTION, err};
			for(int k = 0; k < nn; k++) {
				if(hit_face[k] < 0) continue;
				const int s = nodes[k];
				const int* fv = &m2.faces[3*(size_t)hit_face[k]];
				VirtNodeHost vn;
				memcpy(vn.coords, &hit_p[3*(size_t)k], 3 * sizeof(float));
				float vals[3][13];
				for(int j = 0; j < 3; j++) {
					memcpy(vals[j], &m2.values[9*(size_t)fv[j]], 9 * sizeof(float));
					memcpy(vals[j] + 9, &m2.mat[4*(size_t)fv[j]], 4 * sizeof(float));
				}
				interpolate_triangle(&m2.coords[3*(size_t)fv[0]], &m2.coords[3*(size_t)fv[1]], &m2.coords[3*(size_t)fv[2]],
				                     vn.coords, vals[0], vals[1], vals[2], vn.values, 13);
				vn.remote_zone = m2.zone; vn.remote_face = hit_face[k]; vn.base_node = fv[0];
				vn.real_zone = m1.zone; vn.real_node = m1.border_nodes[s];
				m1.flags[vn.real_node] |= 1;                 // setContactType(InContact)
				axis_plus0[a][s] = (int)virt.size();         // contact_data->axis_plus[0]
				virt.push_back(vn);
			}
		}
}

// The static half of the device detector (gcm_collide.cuh): every ordered pair of local meshes whose outlines
// intersect, with its candidate nodes, faces and face grid -- exactly the lists find_collisions builds per pass.
void build_collision_plan(std::vector<HostMesh*>& meshes, float threshold, CollisionPlan& plan)
{
	const int M = (int)meshes.size();
	plan = CollisionPlan();
	plan.mesh_cand_off.assign(M, 0); plan.mesh_cand_n.assign(M, 0);
	for(int a = 0; a < M; a++) {
		plan.mesh_cand_off[a] = (int)plan.cand_node.size();
		for(int b = 0; b < M; b++) {
			if(a == b) continue;
			HostMesh& m1 = *meshes[a];
			HostMesh& m2 = *meshes[b];
			float lo[3], hi[3];
			bool hit = true;
			for(int j = 0; j < 3; j++) {
				lo[j] = fmaxf(m1.outline_min[j] - threshold, m2.outline_min[j] - threshold);
				hi[j] = fminf(m1.outline_max[j] + threshold, m2.outline_max[j] + threshold);
				if(lo[j] > hi[j]) { hit = false; break; }
			}
			if(!hit) continue;
			gcm::CollPair pr;
			memset(&pr, 0, sizeof pr);
			pr.cand_off = (int)plan.cand_node.size();
			for(size_t s = 0; s < m1.border_nodes.size(); s++) {
				const float* c = &m1.coords[3*(size_t)m1.border_nodes[s]];
				if(c[0] < lo[0] || c[0] > hi[0] || c[1] < lo[1] || c[1] > hi[1] || c[2] < lo[2] || c[2] > hi[2]) continue;
				plan.cand_pair.push_back((int)plan.pairs.size());
				plan.cand_node.push_back(m1.border_nodes[s]);
				plan.cand_slot.push_back((int)s);
				plan.cand_mesh.push_back(a);
			}
			pr.n_cand = (int)plan.cand_node.size() - pr.cand_off;
			std::vector<int> fcs;
			const int F = (int)(m2.faces.size() / 3);
			for(int f = 0; f < F; f++) {
				int out = 0;
				for(int j = 0; j < 3; j++) {
					const float* c = &m2.coords[3*(size_t)m2.faces[3*(size_t)f + j]];
					if(c[0] < lo[0] || c[0] > hi[0] || c[1] < lo[1] || c[1] > hi[1] || c[2] < lo[2] || c[2] > hi[2]) out++;
				}
				if(out != 3) fcs.push_back(f);
			}
			pr.face_off = (int)plan.faces.size();
			pr.n_faces = (int)fcs.size();
			for(int f : fcs) {
				gcm::CollFace cf;
				cf.v0 = m2.faces[3*(size_t)f]; cf.v1 = m2.faces[3*(size_t)f + 1]; cf.v2 = m2.faces[3*(size_t)f + 2]; cf.face = f;
				plan.faces.push_back(cf);
				plan.face_mesh.push_back(b);
			}
			FaceGrid grid;
			grid.build(m2, fcs);
			pr.grid_ok = grid.ok ? 1 : 0;
			pr.ax0 = grid.ax0; pr.ax1 = grid.ax1; pr.n0 = grid.n[0]; pr.n1 = grid.n[1];
			pr.org0 = grid.org[0]; pr.org1 = grid.org[1]; pr.cell = grid.cell; pr.pad = grid.pad;
			pr.ax2 = 3 - grid.ax0 - grid.ax1;
			pr.lo2 = 1e300; pr.hi2 = -1e300;
			for(int f : fcs)
				for(int j = 0; j < 3; j++) {
					const double c = m2.coords[3*(size_t)m2.faces[3*(size_t)f + j] + pr.ax2];
					pr.lo2 = std::min(pr.lo2, c); pr.hi2 = std::max(pr.hi2, c);
				}
			pr.start_off = (int)plan.grid_start.size();
			pr.items_off = (int)plan.grid_items.size();
			if(grid.ok) {
				plan.grid_start.insert(plan.grid_start.end(), grid.start.begin(), grid.start.end());
				plan.grid_items.insert(plan.grid_items.end(), grid.items.begin(), grid.items.end());
			}
			pr.phase = b < a ? 1 : 0;      // partner earlier in local_meshes: already advanced when this mesh runs its stage
			plan.pair_a.push_back(a); plan.pair_b.push_back(b);
			plan.pairs.push_back(pr);
		}
		plan.mesh_cand_n[a] = (int)plan.cand_node.size() - plan.mesh_cand_off[a];
	}
}

} // namespace gcmh
