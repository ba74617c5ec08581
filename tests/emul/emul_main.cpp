// emul_main.cpp -- TEST-ONLY debugging aid (never part of libgcm_b200.so, never shipped).
//
// Single-steps the `__host__ __device__` per-node logic of gcm-necro_b200/csrc/gcm_step.cuh on
// the CPU so that the restatement can be compared with oracle/_ref in a container that has no
// GPU.  Takes the same files and flags as oracle/ref_build/driver.cpp and writes the same
// outputs (<out>.z<k>.values.f32, .flags.i32, .faces.i32, .bases.f32, .tap.i32).
// The product has no CPU path: kernels are validated on the B200 by tests/ -m gpu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <memory>

#define GCM_BND_STATS 1
#include "../../gcm-necro_b200/csrc/gcm_host.h"
#include "../../gcm-necro_b200/csrc/gcm_step.cuh"
#include "../../gcm-necro_b200/csrc/gcm_inner.cuh"
#include "../../gcm-necro_b200/csrc/gcm_border.cuh"

using gcmh::HostMesh;
using std::string;
using std::vector;

static void die(const string& m) { fprintf(stderr, "gcm_emul: %s\n", m.c_str()); exit(2); }
template<class T> static void read_vec(FILE* f, vector<T>& v, size_t n) { v.resize(n); if(n && fread(v.data(), sizeof(T), n, f) != n) die("short read"); }
template<class T> static void write_file(const string& name, const vector<T>& v)
{
	FILE* f = fopen(name.c_str(), "wb"); if(!f) die("cannot open " + name);
	if(v.size()) fwrite(v.data(), sizeof(T), v.size(), f);
	fclose(f);
}

struct Z {
	HostMesh hm;
	size_t stride, rs = 0;                    // rs: quads per plane of the node records (gcm_step.cuh)
	vector<float> xyz, u[2], mat;
	int cur = 0;
	vector<gcm::InnerCacheEntry> icache[3];   // the step-invariant cache of the inner kernel (gcm_inner.cuh)
	vector<int> icache_owner[3];
	vector<int> bnd_static;                   // flat-border tables (gcm_border.cuh), as build_border_flat_tables makes them
	vector<float> bnd_cand;
	vector<float> destr_hist;                 // running maxima of calc_destruction_criterias, 4 per node
};
struct Cond { gcm::BorderCond bc; float start, dur, box[6]; };

int main(int argc, char** argv)
{
	vector<string> mesh_files, state_files, mat_files;
	vector<Cond> conds;
	float la = 7e9f, mu = 1e9f, rho = 7.8e6f, yield = 1e16f;
	int steps = 1, seed = 12345, tap_step = -1, ring_check = 0;
	string out, detector = "void", contact_calc = "adhesion";
	bool detector_static = false;
	vector<vector<float> > translates;
	for(int a = 1; a < argc; a++) {
		string s = argv[a];
		if(s == "--mesh") { while(a + 1 < argc && argv[a+1][0] != '-') mesh_files.push_back(argv[++a]); }
		else if(s == "--state") { while(a + 1 < argc && argv[a+1][0] != '-') state_files.push_back(argv[++a]); }
		else if(s == "--mat") { while(a + 1 < argc && argv[a+1][0] != '-') mat_files.push_back(argv[++a]); }
		else if(s == "--la") la = atof(argv[++a]);
		else if(s == "--mu") mu = atof(argv[++a]);
		else if(s == "--rho") rho = atof(argv[++a]);
		else if(s == "--yield") yield = atof(argv[++a]);
		else if(s == "--steps") steps = atoi(argv[++a]);
		else if(s == "--seed") seed = atoi(argv[++a]);
		else if(s == "--out") out = argv[++a];
		else if(s == "--tap-step") tap_step = atoi(argv[++a]);
		else if(s == "--ring-check") ring_check = atoi(argv[++a]);
		else if(s == "--detector") detector = argv[++a];
		else if(s == "--static") detector_static = true;
		else if(s == "--contact-calc") contact_calc = argv[++a];
		else if(s == "--translate") { vector<float> t; for(int k = 0; k < 4; k++) t.push_back(atof(argv[++a])); translates.push_back(t); }
		else if(s == "--border") {
			Cond c; memset(&c, 0, sizeof c);
			string t = argv[++a];
			float p[5]; for(int k = 0; k < 5; k++) p[k] = atof(argv[++a]);
			c.start = atof(argv[++a]); c.dur = atof(argv[++a]);
			for(int k = 0; k < 6; k++) c.box[k] = atof(argv[++a]);
			c.bc.type = t == "free" ? 0 : t == "fixed" ? 1 : t == "extforce" ? 2 : t == "extvel" ? 3 : t == "extvalues" ? 4 : -1;
			if(c.bc.type < 0) die("unknown border calculator");
			if(c.bc.type == 4) {
				// p0 = the three variable indices as decimal digits, p1..p3 = their values (as the oracle driver)
				const int code = (int)p[0];
				c.bc.vars_index[0] = code / 100; c.bc.vars_index[1] = (code / 10) % 10; c.bc.vars_index[2] = code % 10;
				for(int k = 0; k < 3; k++) c.bc.vars_values[k] = p[1 + k];
			} else {
				c.bc.p_normal = p[0]; c.bc.p_tangential = p[1];
				float d = sqrtf(p[2]*p[2] + p[3]*p[3] + p[4]*p[4]);
				c.bc.dir[0] = p[2] / d; c.bc.dir[1] = p[3] / d; c.bc.dir[2] = p[4] / d;
			}
			conds.push_back(c);
		}
		else die("unknown argument " + s);
	}
	const int NZ = (int)mesh_files.size();
	const bool lean_border = !(getenv("GCM_EMUL_LEAN_BORDER") && atoi(getenv("GCM_EMUL_LEAN_BORDER")) == 0);
	const bool inner_cache = !(getenv("GCM_EMUL_INNER_CACHE") && atoi(getenv("GCM_EMUL_INNER_CACHE")) == 0);
	vector<std::unique_ptr<Z>> zs;
	for(int z = 0; z < NZ; z++) {
		FILE* f = fopen(mesh_files[z].c_str(), "rb"); if(!f) die("cannot open mesh");
		int hdr[5]; if(fread(hdr, 4, 5, f) != 5 || hdr[0] != 0x424D4347) die("bad mesh header");
		int N = hdr[3], T = hdr[4];
		vector<int> placement;
		zs.emplace_back(new Z()); HostMesh& hm = zs.back()->hm;
		hm.zone = z; hm.N = N; hm.T = T;
		read_vec(f, placement, N); read_vec(f, hm.coords, 3*(size_t)N);
		read_vec(f, hm.remote_zone, N); read_vec(f, hm.remote_num, N); read_vec(f, hm.tets, 4*(size_t)T);
		fclose(f);
		hm.flags.resize(N); hm.mat.resize(4*(size_t)N); hm.values.assign(9*(size_t)N, 0.f);
		for(int i = 0; i < N; i++) {
			hm.flags[i] = 16 | 4 | (placement[i] ? 2 : 0);
			if(!placement[i]) for(int k = 0; k < 3; k++) hm.coords[3*(size_t)i+k] = 0;
			hm.mat[4*(size_t)i] = la; hm.mat[4*(size_t)i+1] = mu; hm.mat[4*(size_t)i+2] = rho; hm.mat[4*(size_t)i+3] = yield;
		}
		if((int)mat_files.size() > z) {   // per-node (la, mu, rho, yield), as the oracle driver's --mat
			FILE* g = fopen(mat_files[z].c_str(), "rb"); if(!g) die("cannot open material file");
			read_vec(g, hm.mat, 4*(size_t)N); fclose(g);
		}
	}
	auto host_sync = [&](bool with_coords) {
		for(auto& zp : zs) {
			HostMesh& d = zp->hm;
			for(int i = 0; i < d.N; i++) {
				if(d.flags[i] & 2) continue;
				HostMesh& o = zs[d.remote_zone[i]]->hm;
				int si = d.remote_num[i];
				if(with_coords) {
					memcpy(&d.coords[3*(size_t)i], &o.coords[3*(size_t)si], 12);
					memcpy(&d.mat[4*(size_t)i], &o.mat[4*(size_t)si], 16);
				}
				memcpy(&d.values[9*(size_t)i], &o.values[9*(size_t)si], 36);
			}
		}
	};
	for(auto& t : translates) zs[(int)t[0]]->hm.translate(t[1], t[2], t[3]);
	try {
		host_sync(true);
		for(auto& zp : zs) zp->hm.pre_process_mesh();
	} catch(gcmh::HostError& e) { die("preprocess: " + e.msg); }
	for(size_t c = 0; c < conds.size(); c++)
		for(auto& zp : zs) {
			HostMesh& hm = zp->hm;
			for(size_t b = 0; b < hm.border_nodes.size(); b++) {
				const float* x = &hm.coords[3*(size_t)hm.border_nodes[b]]; const float* box = conds[c].box;
				if(x[0] < box[1] && x[0] > box[0] && x[1] < box[3] && x[1] > box[2] && x[2] < box[5] && x[2] > box[4]) hm.cond_id[b] = (int)c;
			}
		}
	for(int z = 0; z < NZ && z < (int)state_files.size(); z++) {
		HostMesh& hm = zs[z]->hm;
		FILE* f = fopen(state_files[z].c_str(), "rb"); if(!f) die("cannot open state");
		vector<float> st; read_vec(f, st, 9*(size_t)hm.N); fclose(f);
		for(int i = 0; i < hm.N; i++) if(hm.is_local(i)) memcpy(&hm.values[9*(size_t)i], &st[9*(size_t)i], 36);
	}
	// SoA
	for(auto& zp : zs) {
		Z& z = *zp; HostMesh& hm = z.hm;
		z.stride = (hm.N + 31) / 32 * 32;
		z.rs = hm.N;
		z.mat.assign(4*z.stride, 0); z.u[0].assign(GCM_REC*(size_t)hm.N, 0);
		for(int i = 0; i < hm.N; i++) {
			for(int k = 0; k < 3; k++) z.u[0][gcm::rec_index(z.rs, i, 9+k)] = hm.coords[3*(size_t)i+k];
			for(int k = 0; k < 4; k++) z.mat[k*z.stride+i] = hm.mat[4*(size_t)i+k];
			for(int k = 0; k < 9; k++) z.u[0][gcm::rec_index(z.rs, i, k)] = hm.values[9*(size_t)i+k];
		}
		z.u[1] = z.u[0];
	}
	float pf_fail = -1e-3f, pf_pass = -1e-4f; int exact_only = 0;     // as the product computes them (gcm_capi.cu)
	{
		std::vector<const HostMesh*> hms;
		for(auto& zp : zs) hms.push_back(&zp->hm);
		gcmh::prefilter_thresholds(hms, &pf_fail, &pf_pass, &exact_only);
	}
	const int border_mode = getenv("GCM_EMUL_BORDER") ? atoi(getenv("GCM_EMUL_BORDER")) : 2;   // 0 literal, 1 lean, 2 flat (+ literal for what it defers)
	long n_flat_done = 0, n_flat_deferred = 0;
	for(auto& zp : zs) {
		std::vector<gcmh::BorderFlatZone> one(1, gcmh::BorderFlatZone{&zp->hm, 0, 0, 0});
		gcmh::build_border_flat_tables(one, zp->bnd_static, zp->bnd_cand);
	}
	long n_cached = 0, n_cold = 0;
	if(inner_cache) {
		// exactly what build_inner_cache_kernel does on the device
		for(auto& zp : zs) {
			Z& z = *zp; HostMesh& hm = z.hm;
			gcm::MeshView m; memset(&m, 0, sizeof m);
			m.n_nodes = hm.N; m.n_tets = hm.T; m.stride = z.stride;
			m.u = z.u[0].data(); m.rs = z.rs; m.mat = z.mat.data(); m.flags = hm.flags.data();
			m.n2t_off = hm.n2t_off.data(); m.n2t = hm.n2t.data(); m.n2x = nullptr; m.n2x_scan = hm.n2x.data(); m.tets = hm.tets.data(); m.tet_hv = hm.tet_hv.data();
			m.border_slot = hm.border_slot.data();
			m.tau = 0.002f; m.time = 0; m.min_h = hm.min_h;
			m.pf_fail = pf_fail; m.pf_pass = pf_pass; m.exact_only = 1;   // as ensure_inner_cache (gcm_capi.cu)
			for(int st = 0; st < 3; st++) {
				z.icache[st].assign(hm.N, gcm::InnerCacheEntry());
				memset(z.icache[st].data(), 0, hm.N * sizeof(gcm::InnerCacheEntry));
				z.icache_owner[st].assign(2*(size_t)hm.N, -1);
				#pragma omp parallel for schedule(dynamic, 1024) reduction(+:n_cached,n_cold)
				for(size_t q = 0; q < hm.inner_fast.size(); q++) {
					const int i = hm.inner_fast[q];
					if(gcm::build_inner_cache_entry(m, i, st, z.icache[st][i], &z.icache_owner[st][2*(size_t)i])) n_cached++; else n_cold++;
				}
			}
		}
	}
	if(ring_check > 0) {
		// --ring-check N: the list-building device search (gcm::find_owner_general, what the deep contact
		// pass runs) against the host's literal find_owner_tetr on N random (node, displacement) pairs per
		// zone, displacements up to 2.5 mesh steps so that rings 1..4 and "not found" all occur.
		// Prints "ring-check: <tested> <found> <mismatches>".
		long tested = 0, found = 0, bad = 0;
		std::vector<int> ws(GCM_RING_WS);
		gcmh::GlibcRand rr; rr.seed(seed);
		auto uni = [&]() { return (rr.next() % 2000001) / 1000000.0f - 1.0f; };
		for(auto& zp : zs) {
			Z& z = *zp; HostMesh& hm = z.hm;
			gcm::MeshView m; memset(&m, 0, sizeof m);
			m.n_nodes = hm.N; m.n_tets = hm.T; m.stride = z.stride;
			m.u = z.u[0].data(); m.rs = z.rs; m.mat = z.mat.data(); m.flags = hm.flags.data();
			m.n2t_off = hm.n2t_off.data(); m.n2t = hm.n2t.data(); m.tets = hm.tets.data(); m.tet_hv = hm.tet_hv.data();
			for(int k = 0; k < ring_check; k++) {
				const int node = (int)(rr.next() % (unsigned)hm.N);
				if(!hm.is_local(node)) continue;
				const float len = 2.5f * hm.min_h * 2.0f * fabsf(uni());
				float d[3] = {uni(), uni(), uni()};
				const float dn = sqrtf(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]);
				if(!(dn > 0)) continue;
				for(int j = 0; j < 3; j++) d[j] = d[j] / dn * len;
				const float* X = &hm.coords[3*(size_t)node];
				// half of the probes start from a position off the base node, as virtual nodes do
				float pos[3] = {X[0], X[1], X[2]};
				if(k & 1) for(int j = 0; j < 3; j++) pos[j] += 0.3f * hm.min_h * uni();
				const int want = hm.find_owner_tetr(node, pos, d[0], d[1], d[2]);
				gcm::Vec3 np; np.x = pos[0]; np.y = pos[1]; np.z = pos[2];
				gcm::TetGeom g;
				const int got = gcm::find_owner_general(m, node, np, d[0], d[1], d[2], g, ws.data());
				tested++;
				if(want >= 0) found++;
				if(got != want) { if(bad < 5) fprintf(stderr, "ring-check mismatch: zone %d node %d want %d got %d\n", hm.zone, node, want, got); bad++; }
			}
		}
		printf("ring-check: %ld %ld %ld\n", tested, found, bad);
		return bad ? 3 : 0;
	}
	gcmh::GlibcRand rng; rng.seed(seed);
	vector<int> tap;   // records: zone node stage owner[5] outer tries
	int rc = 0;
	const float tau = 0.002f;
	vector<gcmh::VirtNodeHost> virt;
	vector<vector<int> > axis_plus0(NZ);
	for(int z = 0; z < NZ; z++) axis_plus0[z].assign(zs[z]->hm.border_nodes.size(), -1);
	auto make_view = [&](Z& z) {
		HostMesh& hm = z.hm;
		gcm::MeshView m; memset(&m, 0, sizeof m);
		m.n_nodes = hm.N; m.n_tets = hm.T; m.stride = z.stride;
		m.u = z.u[z.cur].data(); m.rs = z.rs; m.mat = z.mat.data(); m.flags = hm.flags.data();
		m.n2t_off = hm.n2t_off.data(); m.n2t = hm.n2t.data(); m.n2x = hm.n2x.data(); m.n2x_scan = hm.n2x.data(); m.tets = hm.tets.data(); m.tet_hv = hm.tet_hv.data();
		m.border_slot = hm.border_slot.data(); m.basis = hm.basis.data(); m.cond_id = hm.cond_id.data();
		m.tau = tau; m.time = hm.current_time; m.min_h = hm.min_h;
		m.pf_fail = pf_fail; m.pf_pass = pf_pass; m.exact_only = exact_only;
		return m;
	};
	for(int step = 0; step < steps && !rc; step++) {
		for(auto& zp : zs) zp->hm.create_random_axes(rng);
		// TetrMeshSet.cpp:111-120: collisions unless the detector is static and t != 0
		if(detector == "brute" && (!detector_static || step == 0)) {
			vector<HostMesh*> ms;
			for(auto& zp : zs) {
				Z& z = *zp;
				for(int i = 0; i < z.hm.N; i++) for(int k = 0; k < 9; k++) z.hm.values[9*(size_t)i+k] = z.u[z.cur][gcm::rec_index(z.rs, i, k)];
				ms.push_back(&z.hm);
			}
			try { gcmh::find_collisions(ms, 1.0f, virt, axis_plus0); }
			catch(gcmh::HostError& e) { die("collisions: " + e.msg); }
		}
		for(int stage = 0; stage < 4 && !rc; stage++) {
			// sync_nodes on the SoA layers
			for(auto& zp : zs) {
				Z& d = *zp;
				for(int i = 0; i < d.hm.N; i++) {
					if(d.hm.flags[i] & 2) continue;
					Z& o = *zs[d.hm.remote_zone[i]]; int si = d.hm.remote_num[i];
					for(int k = 0; k < 9; k++) d.u[d.cur][gcm::rec_index(d.rs, i, k)] = o.u[o.cur][gcm::rec_index(o.rs, si, k)];
				}
			}
			for(int zi = 0; zi < NZ && !rc; zi++) {
				Z& z = *zs[zi]; HostMesh& hm = z.hm;
				float* ucur = z.u[z.cur].data(); float* un = z.u[z.cur^1].data();
				if(stage == 3) {
					for(int i = 0; i < hm.N; i++) {
						if(!hm.is_local(i)) continue;
						float v[9]; for(int k = 0; k < 9; k++) v[k] = ucur[gcm::rec_index(z.rs, i, k)];
						if(gcm::drop_deviator(v, z.mat[3*z.stride+i])) for(int k = 3; k < 9; k++) ucur[gcm::rec_index(z.rs, i, k)] = v[k];
					}
					continue;
				}
				gcm::MeshView m; memset(&m, 0, sizeof m);
				m.n_nodes = hm.N; m.n_tets = hm.T; m.stride = z.stride;
				m.u = ucur; m.rs = z.rs; m.mat = z.mat.data(); m.flags = hm.flags.data();
				m.n2t_off = hm.n2t_off.data(); m.n2t = hm.n2t.data(); m.n2x = hm.n2x.data(); m.n2x_scan = hm.n2x.data(); m.tets = hm.tets.data(); m.tet_hv = hm.tet_hv.data();
				m.border_slot = hm.border_slot.data(); m.basis = hm.basis.data(); m.cond_id = hm.cond_id.data();
				m.tau = tau; m.time = hm.current_time; m.min_h = hm.min_h;
				m.pf_fail = pf_fail; m.pf_pass = pf_pass; m.exact_only = exact_only;
				m.n_conds = (int)conds.size();
				for(size_t c = 0; c < conds.size(); c++) {
					m.conds[c] = conds[c].bc;
					float t = hm.current_time;
					m.conds[c].active = (conds[c].dur < 0) ? 1 : !(t > conds[c].start + conds[c].dur);
					m.conds[c].scale = t < conds[c].start ? 0.f : conds[c].dur < 0 ? 1.f : (t <= conds[c].start + conds[c].dur ? 1.f : 0.f);
				}
				int err = 0;
				#pragma omp parallel for schedule(dynamic, 1024)
				for(int i = 0; i < hm.N; i++) {
					if(!hm.is_local(i)) continue;
					float o[9]; gcm::StageTap t;
					int e;
					const int slot = hm.border_slot[i];
					const int vidx = (stage == 0 && slot >= 0) ? axis_plus0[zi][slot] : -1;
					if(vidx >= 0) {
						// contact branch; the partner mesh is read as it is NOW (zones advance one after the other)
						gcmh::VirtNodeHost& vh = virt[vidx];
						gcm::MeshView mv = make_view(*zs[vh.remote_zone]);
						gcm::VirtNode vn;
						for(int k = 0; k < 3; k++) vn.pos[k] = vh.coords[k];
						for(int k = 0; k < 9; k++) vn.val[k] = vh.values[k];
						vn.la = vh.values[9]; vn.mu = vh.values[10]; vn.rho = vh.values[11];
						vn.base = vh.base_node; vn.real = i; vn.phase = 0;
						float vo[9]; bool vw = false;
						// workspace of the list-building owner search (the product's deep contact pass)
						static thread_local std::vector<int> ring_ws(GCM_RING_WS);
						gcm::MeshView mc = m;
						mc.ring_ws = ring_ws.data(); mv.ring_ws = ring_ws.data();
						e = gcm::node_stage_contact(mc, mv, i, vn, vidx, stage, &hm.normals[3*(size_t)slot], contact_calc == "friction" ? 1 : 0, o, &t, vo, &vw, step == 0);
						if(vw) for(int k = 0; k < 9; k++) vh.values[k] = vo[k];
					} else if(slot >= 0 && border_mode == 2) {
						// stage_border_flat_kernel, then the literal routine for what it defers
						const gcm::BndSlotStatic& bs = reinterpret_cast<const gcm::BndSlotStatic*>(z.bnd_static.data())[slot];
						gcm::LocalMat9 A;
						e = gcm::node_stage_border_fast(m, bs, reinterpret_cast<const gcm::gcm_f4*>(z.bnd_cand.data()) + bs.cand_off, i, stage, o, &t, &A);
						if(e == GCM_BND_DEFER) {
							e = gcm::node_stage_nocontact(m, i, stage, o, &t);
							#pragma omp atomic
							n_flat_deferred++;
						} else {
							#pragma omp atomic
							n_flat_done++;
						}
					} else if(slot >= 0 && border_mode == 1 && lean_border)
						e = gcm::node_stage_border_lean(m, i, stage, o, &t);
					else if(inner_cache && slot < 0 && (z.icache[stage][i].flags & GCM_IC_VALID)
					        && (z.icache[0][i].flags & z.icache[1][i].flags & z.icache[2][i].flags & GCM_IC_VALID)) {
						// stage_inner_cached_kernel
						const gcm::InnerCacheEntry& ce = z.icache[stage][i];
						const gcmh::MaterialRegistry& reg = hm.registry ? *hm.registry : hm.own_registry;
						const gcm::MatTab* tab = reinterpret_cast<const gcm::MatTab*>(reg.table.data()) + 3 * (size_t)hm.mat_id[i] + stage;
						float own[9]; bool bad = false;
						for(int k = 0; k < 9; k++) { own[k] = ucur[gcm::rec_index(z.rs, i, k)]; bad |= !(fabsf(own[k]) <= 3.4028235e38f); }
						e = bad ? gcm::ERR_NAN : 0;
						if(!e) {
							if(stage == 0) gcm::inner_cached_stage<0>(ucur, z.rs, ce, own, tab->U, tab->U1, o);
							else if(stage == 1) gcm::inner_cached_stage<1>(ucur, z.rs, ce, own, tab->U, tab->U1, o);
							else gcm::inner_cached_stage<2>(ucur, z.rs, ce, own, tab->U, tab->U1, o);
						}
						t.owner[0] = t.owner[2] = z.icache_owner[stage][2*(size_t)i];
						t.owner[1] = t.owner[3] = z.icache_owner[stage][2*(size_t)i+1];
						t.owner[4] = hm.n2t[hm.n2t_off[i]];
						t.outer_count = 0; t.tries = 1;
					} else
						e = gcm::node_stage_nocontact(m, i, stage, o, &t);
					if(e) {
						#pragma omp critical
						{ if(!err) { err = e; fprintf(stderr, "gcm_emul: error %d zone %d node %d stage %d\n", e, zi, i, stage); } }
						continue;
					}
					for(int k = 0; k < 9; k++) un[gcm::rec_index(z.rs, i, k)] = o[k];
					if(step == tap_step) {
						#pragma omp critical
						{ tap.push_back(zi); tap.push_back(i); tap.push_back(stage); for(int k = 0; k < 5; k++) tap.push_back(t.owner[k]); tap.push_back(t.outer_count); tap.push_back(t.tries); }
					}
				}
				if(err) rc = 10 + err;
				// the copy-back loop (TetrMesh_1stOrder.cpp:1925-1932) touches local nodes only: a Remote entry keeps the
				// value synced before the stage, and that is what a later mesh's virtual nodes read in this mesh
				for(int i = 0; i < hm.N; i++)
					if(!(hm.flags[i] & 2)) for(int k = 0; k < 9; k++) un[gcm::rec_index(z.rs, i, k)] = ucur[gcm::rec_index(z.rs, i, k)];
				z.cur ^= 1;
			}
		}
		// calc_destruction_criterias (TetrMeshSet.cpp:177-178): running maxima of every local node
		for(auto& zp : zs) {
			Z& z = *zp;
			if(z.destr_hist.empty()) z.destr_hist.assign(4*(size_t)z.hm.N, 0.f);
			for(int i = 0; i < z.hm.N; i++) {
				if(!z.hm.is_local(i)) continue;
				float v[9], c[4];
				for(int k = 0; k < 9; k++) v[k] = z.u[z.cur][gcm::rec_index(z.rs, i, k)];
				gcm::destruction_now(v, c);
				for(int k = 0; k < 4; k++) if(c[k] > z.destr_hist[4*(size_t)i+k]) z.destr_hist[4*(size_t)i+k] = c[k];
			}
		}
		for(auto& zp : zs) zp->hm.current_time += tau;
	}
	if(out.size()) {
		for(int zi = 0; zi < NZ; zi++) {
			Z& z = *zs[zi]; HostMesh& hm = z.hm;
			vector<float> vals(9*(size_t)hm.N), bases(9*(size_t)hm.N, 0.f);
			for(int i = 0; i < hm.N; i++) {
				for(int k = 0; k < 9; k++) vals[9*(size_t)i+k] = z.u[z.cur][gcm::rec_index(z.rs, i, k)];
				if(hm.is_local(i)) {
					int b = hm.border_slot[i];
					const float E[9] = {1,0,0,0,1,0,0,0,1};
					memcpy(&bases[9*(size_t)i], b >= 0 ? &hm.basis[9*(size_t)b] : E, 36);
				}
			}
			vector<int> fl(hm.N);
			for(int i = 0; i < hm.N; i++) fl[i] = hm.flags[i] & 15;
			char zs_[32]; snprintf(zs_, sizeof zs_, ".z%d", zi);
			write_file(out + zs_ + ".values.f32", vals);
			write_file(out + zs_ + ".flags.i32", fl);
			write_file(out + zs_ + ".faces.i32", hm.faces);
			write_file(out + zs_ + ".bases.f32", bases);
			vector<float> destr(8*(size_t)hm.N, 0.f);
			for(int i = 0; i < hm.N && steps > 0; i++) {
				if(!hm.is_local(i)) continue;
				gcm::destruction_now(&vals[9*(size_t)i], &destr[8*(size_t)i]);
				for(int k = 0; k < 4; k++) destr[8*(size_t)i+4+k] = z.destr_hist[4*(size_t)i+k];
			}
			write_file(out + zs_ + ".destr.f32", destr);
		}
		if(tap_step >= 0) write_file(out + ".tap.i32", tap);
	}
	if(getenv("GCM_EMUL_BND_STATS")) { fprintf(stderr, "flat defer reasons:"); for(int r = 1; r < 10; r++) fprintf(stderr, " %d:%ld", r, gcm::g_bnd_defer_reason[r]); fprintf(stderr, "\n"); }
	printf("{\"impl\": \"emul\", \"steps\": %d, \"rc\": %d, \"inner_cached\": %ld, \"inner_cold\": %ld, \"flat_done\": %ld, \"flat_deferred\": %ld, \"exact_only\": %d, \"pf_pass\": %g}\n", steps, rc, n_cached, n_cold, n_flat_done, n_flat_deferred, exact_only, pf_pass);
	return rc;
}
