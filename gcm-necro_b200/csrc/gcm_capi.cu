// gcm_capi.cu -- the C ABI of include/gcm_b200.h: TetrMeshSet-level orchestration of the
// device-resident zones (gcm/system/TetrMeshSet.cpp:101-184) over the kernels of
// gcm_kernels.cuh and the host-side mesh logic of gcm_host.cpp.
#include "../../include/gcm_b200.h"
#include "gcm_host.h"
#include "gcm_kernels.cuh"

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using gcmh::HostError;
using gcmh::HostMesh;

namespace {

thread_local std::string g_last_error;

int fail(int exc_code, const std::string& msg)
{
	g_last_error = msg;
	return -(exc_code + 1);
}

struct CudaError { std::string msg; };
#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
	throw CudaError{std::string(#call) + ": " + cudaGetErrorString(e_)}; } while(0)

template<class T> struct DevBuf {
	T* p = nullptr; size_t n = 0;
	void alloc(size_t count) { free(); n = count; if(count) CK(cudaMalloc(&p, count * sizeof(T))); }
	void upload(const std::vector<T>& h, cudaStream_t st) {
		alloc(h.size());
		if(h.size()) CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
	}
	void free() { if(p) cudaFree(p); p = nullptr; n = 0; }
	~DevBuf() { free(); }
};

struct CondSpec {   // one BorderCondition: calculator + StepPulseForm + BoxArea
	gcm::BorderCond bc;
	float start, duration;
	float box[6];
};

struct HaloGroup {  // nodes of dst zone copied from src zone (same process)
	int dst_zone, src_zone;
	std::vector<int> h_dst, h_src;
	DevBuf<int> dst, src;
};

struct PeerPlan {   // exchange with one other process
	std::vector<int> recv_zone, recv_node;       // my halo slots, in request order
	std::vector<int> req_pairs;                  // (owner zone, owner node) per slot
	std::vector<int> send_zone, send_node;       // my nodes the peer asked for
	// per zone segments for the device kernels
	struct Seg { int zone, offset, n; DevBuf<int> idx; };
	std::vector<std::unique_ptr<Seg>> recv_segs, send_segs;
	bool dev_ready = false;
};

struct Zone {
	HostMesh hm;
	bool loaded = false;
	size_t stride = 0;
	DevBuf<float> xyz, u0, u1, mat, tet_hv, basis, aos, normals;
	int slot_offset = 0;                       // first global border slot of this zone (device RNG)
	DevBuf<int> flags, n2t_off, n2t, tets, border_slot, border_nodes, inner_nodes, cond_id, inner_fast, inner_generic;
	DevBuf<float> w0, mat_tab;
	DevBuf<int> n2x;
	DevBuf<unsigned char> mat_id;
	DevBuf<gcm::StageTap> tap;
	int cur = 0;
	float* h_basis[2] = {nullptr, nullptr};   // pinned staging, double buffered
	cudaEvent_t basis_ev[2] = {nullptr, nullptr};
	int basis_flip = 0;
	bool on_device = false;
	float* ucur() { return cur ? u1.p : u0.p; }
	float* unext() { return cur ? u0.p : u1.p; }
	~Zone() {
		for(int k = 0; k < 2; k++) { if(h_basis[k]) cudaFreeHost(h_basis[k]); if(basis_ev[k]) cudaEventDestroy(basis_ev[k]); }
	}
};

} // namespace

struct gcm_b200_set {
	int device = 0, n_zones = 0, my_rank = 0;
	std::vector<int> zone_rank;
	std::vector<std::unique_ptr<Zone>> zones;   // indexed by zone number; null when remote
	std::vector<int> local_zones;               // TetrMeshSet::local_meshes order
	gcmh::GlibcRand rng;
	cudaStream_t stream = nullptr;
	int* d_err = nullptr;
	int* h_err = nullptr;
	std::vector<CondSpec> conds;
	std::vector<std::unique_ptr<HaloGroup>> halo_groups;
	std::map<int, PeerPlan> peers;
	bool plan_built = false;
	bool preprocessed = false;
	bool tap_on = false;
	int detector_type = 0, detector_static = 0; float detector_threshold = 1;
	int contact_calc = 0;
	long long launches = 0;
	float time_step = 0.002f;
	bool in_step = false;
	// device-side rand() stream + bases (rng_chunk_kernel / basis_kernel)
	bool fast_inner = true;   // specialised inner-node kernel (gcm_inner.cuh); false = generic kernel for every node
	bool basis_on_device = true;
	bool rng_ready = false;
	long long rng_draws_per_step = 0;
	int rng_chunks = 0;
	uint32_t rng_jump[961];
	DevBuf<uint32_t> d_chunk_state, d_jump, d_slot_pos;
	DevBuf<int> d_chunk_first_slot;
	DevBuf<unsigned short> d_teta_idx;
	DevBuf<float> d_cos, d_sin;
	// per-kernel CUDA-event profile (gcm_b200_set_profile): 0 inner stage, 1 border stage, 2 plastic, 3 halo
	bool profile = false;
	struct ProfRec { int which; long long units; cudaEvent_t a, b; };
	std::vector<ProfRec> prof;
	double prof_ms[6] = {0, 0, 0, 0, 0, 0};
	long long prof_launches[6] = {0, 0, 0, 0, 0, 0}, prof_units[6] = {0, 0, 0, 0, 0, 0};
};

namespace {

Zone* get_zone(gcm_b200_set* s, int zone)
{
	if(!s || zone < 0 || zone >= s->n_zones || !s->zones[zone]) return nullptr;
	return s->zones[zone].get();
}

void build_halo_plan(gcm_b200_set* s)
{
	if(s->plan_built) return;
	s->halo_groups.clear();
	s->peers.clear();
	std::map<std::pair<int,int>, HaloGroup*> groups;
	for(int zi : s->local_zones) {
		HostMesh& hm = s->zones[zi]->hm;
		for(int i = 0; i < hm.N; i++) {
			if(hm.flags[i] & gcm::NF_LOCAL) continue;
			int oz = hm.remote_zone[i], on = hm.remote_num[i];
			if(oz < 0 || oz >= s->n_zones) throw HostError{gcmh::INPUT_EXCEPTION, "remote node refers to an unknown zone"};
			int owner_rank = s->zone_rank[oz];
			if(owner_rank == s->my_rank) {
				auto key = std::make_pair(zi, oz);
				HaloGroup*& g = groups[key];
				if(!g) { s->halo_groups.emplace_back(new HaloGroup()); g = s->halo_groups.back().get(); g->dst_zone = zi; g->src_zone = oz; }
				g->h_dst.push_back(i); g->h_src.push_back(on);
			} else {
				PeerPlan& p = s->peers[owner_rank];
				p.recv_zone.push_back(zi); p.recv_node.push_back(i);
				p.req_pairs.push_back(oz); p.req_pairs.push_back(on);
			}
		}
	}
	s->plan_built = true;
}

// DataBus::sync_nodes, same-process branch, on the host mirrors (start-up: coords + 13 values).
void host_sync_nodes(gcm_b200_set* s)
{
	for(auto& gp : s->halo_groups) {
		HostMesh& d = s->zones[gp->dst_zone]->hm;
		HostMesh& o = s->zones[gp->src_zone]->hm;
		for(size_t k = 0; k < gp->h_dst.size(); k++) {
			int di = gp->h_dst[k], si = gp->h_src[k];
			if(si < 0 || si >= o.N) throw HostError{gcmh::INPUT_EXCEPTION, "remote node index out of range"};
			memcpy(&d.coords[3*(size_t)di], &o.coords[3*(size_t)si], 3 * sizeof(float));
			memcpy(&d.values[9*(size_t)di], &o.values[9*(size_t)si], 9 * sizeof(float));
			memcpy(&d.mat[4*(size_t)di], &o.mat[4*(size_t)si], 4 * sizeof(float));
		}
	}
}

// Host node-major values -> both device layers (start-up; includes halo slots).
void upload_values(gcm_b200_set* s, Zone& z)
{
	const HostMesh& hm = z.hm;
	const int n = hm.N;
	if(!n) return;
	if(!z.aos.p) z.aos.alloc(9 * (size_t)n);
	CK(cudaMemcpyAsync(z.aos.p, hm.values.data(), 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
	gcm::values_to_records_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(z.u0.p, z.u1.p, z.aos.p, z.flags.p, n, 0);
	s->launches++;
	CK(cudaStreamSynchronize(s->stream));
	z.cur = 0;
}

void upload_zone(gcm_b200_set* s, Zone& z)
{
	HostMesh& hm = z.hm;
	const size_t N = hm.N;
	z.stride = (N + 31) / 32 * 32;
	z.xyz.upload(hm.coords, s->stream);   // node-major staging; lives on in the records
	std::vector<float> soa(4 * z.stride, 0.f);
	for(size_t i = 0; i < N; i++)
		for(int k = 0; k < 4; k++) soa[k * z.stride + i] = hm.mat[4*i + k];
	z.mat.upload(soa, s->stream);
	CK(cudaStreamSynchronize(s->stream));
	z.u0.alloc(GCM_REC * (N + 1)); z.u1.alloc(GCM_REC * (N + 1));
	CK(cudaMemsetAsync(z.u0.p, 0, GCM_REC * (N + 1) * sizeof(float), s->stream));
	CK(cudaMemsetAsync(z.u1.p, 0, GCM_REC * (N + 1) * sizeof(float), s->stream));
	if(N) {
		gcm::coords_to_records_kernel<<<(int)((N + 255) / 256), 256, 0, s->stream>>>(z.u0.p, z.u1.p, z.xyz.p, (int)N);
		s->launches++;
	}
	z.flags.upload(hm.flags, s->stream);
	z.n2t_off.upload(hm.n2t_off, s->stream);
	z.n2t.upload(hm.n2t, s->stream);
	z.tets.upload(hm.tets, s->stream);
	z.tet_hv.upload(hm.tet_hv, s->stream);
	z.border_slot.upload(hm.border_slot, s->stream);
	z.border_nodes.upload(hm.border_nodes, s->stream);
	z.inner_nodes.upload(hm.inner_nodes, s->stream);
	z.inner_fast.upload(hm.inner_fast, s->stream);
	z.inner_generic.upload(hm.inner_generic, s->stream);
	z.w0.upload(hm.w0, s->stream);
	z.mat_tab.upload(hm.mat_table, s->stream);
	z.mat_id.upload(hm.mat_id, s->stream);
	z.n2x.upload(hm.n2x, s->stream);
	z.cond_id.upload(hm.cond_id, s->stream);
	z.basis.alloc(hm.basis.size());
	CK(cudaStreamSynchronize(s->stream));
	for(int k = 0; k < 2; k++) {
		if(hm.basis.size()) CK(cudaMallocHost(&z.h_basis[k], hm.basis.size() * sizeof(float)));
		CK(cudaEventCreateWithFlags(&z.basis_ev[k], cudaEventDisableTiming));
	}
	upload_values(s, z);
	z.on_device = true;
}

gcm::MeshView make_view(gcm_b200_set* s, Zone& z)
{
	gcm::MeshView m;
	memset(&m, 0, sizeof(m));
	m.n_nodes = z.hm.N; m.n_tets = z.hm.T; m.stride = z.stride;
	m.u = z.ucur(); m.mat = z.mat.p; m.flags = z.flags.p;
	m.n2t_off = z.n2t_off.p; m.n2t = z.n2t.p; m.tets = z.tets.p; m.tet_hv = z.tet_hv.p;
	m.border_slot = z.border_slot.p; m.basis = z.basis.p; m.cond_id = z.cond_id.p; m.contact = nullptr;
	m.tau = s->time_step; m.time = z.hm.current_time; m.min_h = z.hm.min_h;
	m.n_conds = (int)s->conds.size();
	for(int c = 0; c < m.n_conds; c++) {
		const CondSpec& cs = s->conds[c];
		m.conds[c] = cs.bc;
		// check_stresses (TetrMesh_1stOrder.cpp:2024-2025) + BorderCondition::do_calc's scale
		const float t = z.hm.current_time;
		bool active = (cs.duration < 0) ? true : !(t > cs.start + cs.duration);    // PulseForm::isActive
		float scale;                                                            // StepPulseForm::calcMagnitudeNorm
		if(t < cs.start) scale = 0.0f;
		else if(cs.duration < 0) scale = 1.0f;
		else if(t <= cs.start + cs.duration) scale = 1.0f;
		else scale = 0.0f;
		m.conds[c].active = active ? 1 : 0;
		m.conds[c].scale = scale;
	}
	return m;
}

struct ProfScope {
	gcm_b200_set* s; int which; long long units; cudaEvent_t a = nullptr, b = nullptr;
	ProfScope(gcm_b200_set* s_, int w, long long u) : s(s_), which(w), units(u) {
		if(!s->profile) return;
		CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
		CK(cudaEventRecord(a, s->stream));
	}
	~ProfScope() {
		if(!a) return;
		cudaEventRecord(b, s->stream);
		s->prof.push_back({which, units, a, b});
	}
};

void prof_drain(gcm_b200_set* s)
{
	for(auto& r : s->prof) {
		float ms = 0;
		cudaEventSynchronize(r.b);
		cudaEventElapsedTime(&ms, r.a, r.b);
		s->prof_ms[r.which] += ms; s->prof_launches[r.which]++; s->prof_units[r.which] += r.units;
		cudaEventDestroy(r.a); cudaEventDestroy(r.b);
	}
	s->prof.clear();
}

// Lay the step's rand() stream out for the device: draw positions of the border nodes, chunk
// start windows from the current host generator state, the one-step jump matrix.
void rng_device_init(gcm_b200_set* s)
{
	std::vector<uint32_t> slot_pos;
	unsigned long long pos = 0;
	int slot_total = 0;
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		const HostMesh& hm = z.hm;
		z.slot_offset = slot_total;
		for(int i = 0; i < hm.N; i++) {
			if(!hm.is_local(i)) continue;
			if(hm.is_border(i)) { slot_pos.push_back((uint32_t)pos); pos += 1; }
			else pos += 3;
		}
		slot_total += (int)hm.border_nodes.size();
	}
	if(pos >= 0xffffffffull) throw HostError{gcmh::CONFIG_EXCEPTION, "too many rand() draws per step for the device generator"};
	s->rng_draws_per_step = (long long)pos;
	s->rng_chunks = (int)((pos + GCM_RNG_CHUNK - 1) / GCM_RNG_CHUNK);
	const int nc = s->rng_chunks;
	std::vector<uint32_t> st((size_t)31 * nc);
	std::vector<int> first(nc + 1, 0);
	{
		gcmh::GlibcRand g = s->rng;
		size_t sl = 0;
		for(int c = 0; c < nc; c++) {
			uint32_t w[31];
			g.get_window(w);
			for(int i = 0; i < 31; i++) st[(size_t)i * nc + c] = w[i];
			while(sl < slot_pos.size() && slot_pos[sl] < (uint32_t)c * GCM_RNG_CHUNK) sl++;
			first[c] = (int)sl;
			if(c + 1 < nc) for(int k = 0; k < GCM_RNG_CHUNK; k++) (void)g.next();
		}
		first[nc] = (int)slot_pos.size();
	}
	gcmh::rng_jump_matrix((unsigned long long)pos, s->rng_jump);
	std::vector<uint32_t> jm(s->rng_jump, s->rng_jump + 961);
	float ct[360], sn[360];
	gcmh::rng_teta_tables(ct, sn);
	s->d_chunk_state.upload(st, s->stream);
	s->d_jump.upload(jm, s->stream);
	s->d_slot_pos.upload(slot_pos, s->stream);
	s->d_chunk_first_slot.upload(first, s->stream);
	s->d_teta_idx.alloc(slot_pos.size() + 1);
	s->d_cos.upload(std::vector<float>(ct, ct + 360), s->stream);
	s->d_sin.upload(std::vector<float>(sn, sn + 360), s->stream);
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		if(!z.normals.p) z.normals.upload(z.hm.normals, s->stream);
	}
	CK(cudaStreamSynchronize(s->stream));
	s->rng_ready = true;
}

void begin_step(gcm_b200_set* s)
{
	// get_max_possible_tau(): only its side effect survives: new local bases,
	// meshes in local_meshes order, nodes in index order, one libc rand() stream.
	if(s->basis_on_device) {
		if(!s->rng_ready) rng_device_init(s);
		if(s->rng_chunks) {
			ProfScope ps(s, 4, s->rng_draws_per_step);
			gcm::rng_chunk_kernel<<<(s->rng_chunks + 127) / 128, 128, 0, s->stream>>>(s->d_chunk_state.p, s->rng_chunks,
				s->d_jump.p, s->d_slot_pos.p, s->d_chunk_first_slot.p, s->d_teta_idx.p);
			s->launches++;
			for(int zi : s->local_zones) {
				Zone& z = *s->zones[zi];
				int nb = (int)z.hm.border_nodes.size();
				if(!nb) continue;
				gcm::basis_kernel<<<(nb + 127) / 128, 128, 0, s->stream>>>(z.basis.p, z.normals.p, s->d_teta_idx.p, z.slot_offset, nb, s->d_cos.p, s->d_sin.p);
				s->launches++;
			}
		}
		// keep the host generator in step with the device one
		uint32_t w[31];
		s->rng.get_window(w);
		gcmh::rng_apply_jump(s->rng_jump, w);
		s->rng.set_window(w);
		s->in_step = true;
		return;
	}
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		z.hm.create_random_axes(s->rng);
		const size_t nb = z.hm.basis.size();
		if(nb) {
			int b = z.basis_flip; z.basis_flip ^= 1;
			CK(cudaEventSynchronize(z.basis_ev[b]));
			memcpy(z.h_basis[b], z.hm.basis.data(), nb * sizeof(float));
			CK(cudaMemcpyAsync(z.basis.p, z.h_basis[b], nb * sizeof(float), cudaMemcpyHostToDevice, s->stream));
			CK(cudaEventRecord(z.basis_ev[b], s->stream));
		}
	}
	s->in_step = true;
}

void sync_nodes_device(gcm_b200_set* s)
{
	for(auto& gp : s->halo_groups) {
		Zone& d = *s->zones[gp->dst_zone];
		Zone& o = *s->zones[gp->src_zone];
		int n = (int)gp->h_dst.size();
		if(!n) continue;
		if(!gp->dst.p) { gp->dst.upload(gp->h_dst, s->stream); gp->src.upload(gp->h_src, s->stream); }
		{
			ProfScope ps(s, 3, n);
			gcm::halo_copy_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(d.ucur(), gp->dst.p, o.ucur(), gp->src.p, n);
		}
		s->launches++;
	}
}

void zone_stage(gcm_b200_set* s, int zi, float tau, int stage)
{
	Zone& z = *s->zones[zi];
	if(stage == 3) {
		int n = z.hm.N;
		if(n) {
			ProfScope ps(s, 2, (long long)(z.hm.border_nodes.size() + z.hm.inner_nodes.size()));
			gcm::plastic_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(z.ucur(), z.mat.p, z.flags.p, n, z.stride, s->d_err, zi);
			s->launches++;
		}
		return;
	}
	gcm::MeshView m = make_view(s, z);
	m.tau = tau;
	gcm::StageTap* tap = s->tap_on ? z.tap.p : nullptr;
	int nb = (int)z.hm.border_nodes.size(), ni = (int)z.hm.inner_nodes.size();
	if(nb) {
		ProfScope ps(s, 1, nb);
		gcm::stage_generic_kernel<<<(nb + 127) / 128, 128, 0, s->stream>>>(m, z.border_nodes.p, nb, stage, z.unext(), s->d_err, zi, tap);
		s->launches++;
	}
	if(!s->fast_inner) {
		if(ni) {
			ProfScope ps(s, 0, ni);
			gcm::stage_generic_kernel<<<(ni + 127) / 128, 128, 0, s->stream>>>(m, z.inner_nodes.p, ni, stage, z.unext(), s->d_err, zi, tap);
			s->launches++;
		}
	} else {
		const int ng = (int)z.hm.inner_generic.size(), nf = (int)z.hm.inner_fast.size();
		if(ng) {
			ProfScope ps(s, 5, ng);
			gcm::stage_generic_kernel<<<(ng + 127) / 128, 128, 0, s->stream>>>(m, z.inner_generic.p, ng, stage, z.unext(), s->d_err, zi, tap);
			s->launches++;
		}
		if(nf) {
			ProfScope ps(s, 0, nf);
			const int n_mat = (int)(z.hm.materials.size() / 3);
			const unsigned char* mid = n_mat > 1 ? z.mat_id.p : nullptr;
			const gcm::MatTab* tab = reinterpret_cast<const gcm::MatTab*>(z.mat_tab.p);
			const dim3 grid((nf + 127) / 128), block(128);
			if(stage == 0) gcm::stage_inner_kernel<0><<<grid, block, 0, s->stream>>>(m, z.inner_fast.p, nf, z.unext(), s->d_err, zi, tap, tab, n_mat, mid, z.w0.p, reinterpret_cast<const int4*>(z.n2x.p));
			else if(stage == 1) gcm::stage_inner_kernel<1><<<grid, block, 0, s->stream>>>(m, z.inner_fast.p, nf, z.unext(), s->d_err, zi, tap, tab, n_mat, mid, z.w0.p, reinterpret_cast<const int4*>(z.n2x.p));
			else gcm::stage_inner_kernel<2><<<grid, block, 0, s->stream>>>(m, z.inner_fast.p, nf, z.unext(), s->d_err, zi, tap, tab, n_mat, mid, z.w0.p, reinterpret_cast<const int4*>(z.n2x.p));
			s->launches++;
		}
	}
	// copy-back loop of TetrMesh_1stOrder.cpp:1925-1932 == buffer swap: the halo slots of the new
	// current layer are refreshed by the sync_nodes that precedes every stage
	z.cur ^= 1;
}

const char* step_error_text(int code)
{
	switch(code) {
	case gcm::ERR_NAN: return "NAN or INF detected. Instability?";
	case gcm::ERR_OUTER_COUNT: return "Illegal number of outer characteristics";
	case gcm::ERR_NOT_BORDER: return "Border node is not marked as border";
	case gcm::ERR_BOTH_CONTACT: return "Both direction of contact are marked as contact";
	case gcm::ERR_INTERP_SUM: return "Sum of factors is greater than 1.0";
	case gcm::ERR_INNER_NO_OWNER: return "Inner node without owner tetrahedron (find_border_cross path is not built)";
	case gcm::ERR_RING_OVERFLOW: return "Owner search needs a second ring of tetrahedrons (not built on device)";
	case gcm::ERR_BAD_RHEOLOGY: return "Bad rheology";
	case gcm::ERR_VIRT_OUTER_COUNT: return "Illegal number of outer characteristics";
	case gcm::ERR_BAD_Q: return "Bad vector q";
	case gcm::ERR_ZERO_VOLUME: return "Tetr volume is zero";
	case gcm::ERR_FOOT_PATTERN: return "More than 5 distinct characteristic feet";
	default: return "Unknown device error";
	}
}

int step_error_exc(int code)
{
	switch(code) {
	case gcm::ERR_INTERP_SUM: case gcm::ERR_ZERO_VOLUME: return gcmh::MESH_EXCEPTION;
	case gcm::ERR_BAD_RHEOLOGY: case gcm::ERR_BAD_Q: return gcmh::CONFIG_EXCEPTION;
	case gcm::ERR_INNER_NO_OWNER: case gcm::ERR_RING_OVERFLOW: return gcmh::UNIMPLEMENTED_EXCEPTION;
	default: return gcmh::METHOD_EXCEPTION;
	}
}

} // namespace

#define GUARD_BEGIN try {
#define GUARD_END } catch(HostError& e) { return fail(e.code, e.msg); } \
	catch(CudaError& e) { return fail(gcmh::CONFIG_EXCEPTION, "CUDA: " + e.msg); } \
	catch(std::exception& e) { return fail(gcmh::CONFIG_EXCEPTION, e.what()); }

extern "C" {

const char* gcm_b200_version(void) { return "gcm_b200 0.1 (sm_100a)"; }
const char* gcm_b200_last_error(void) { return g_last_error.c_str(); }

int gcm_b200_set_create(gcm_b200_set** out, int device, int n_zones, const int* zone_rank, int my_rank)
{
	GUARD_BEGIN
	if(!out || n_zones <= 0) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	std::unique_ptr<gcm_b200_set> s(new gcm_b200_set());
	if(device >= 0) {
		int ndev = 0;
		if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
			return fail(gcmh::CONFIG_EXCEPTION, "no CUDA device: libgcm_b200 has no CPU execution path");
		if(device >= ndev) return fail(gcmh::CONFIG_EXCEPTION, "bad device ordinal");
		CK(cudaSetDevice(device));
	}
	s->device = device; s->n_zones = n_zones; s->my_rank = my_rank;
	s->zone_rank.assign(n_zones, my_rank);
	if(zone_rank) s->zone_rank.assign(zone_rank, zone_rank + n_zones);
	s->zones.resize(n_zones);
	for(int z = 0; z < n_zones; z++)
		if(s->zone_rank[z] == my_rank) { s->zones[z].reset(new Zone()); s->local_zones.push_back(z); }
	if(device >= 0) {
		CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
		CK(cudaMalloc(&s->d_err, 4 * sizeof(int)));
		CK(cudaMemset(s->d_err, 0, 4 * sizeof(int)));
		CK(cudaMallocHost(&s->h_err, 4 * sizeof(int)));
	}
	s->rng.seed(1);   // libc default when srand() was never called
	*out = s.release();
	return 0;
	GUARD_END
}

int gcm_b200_set_destroy(gcm_b200_set* s)
{
	if(!s) return 0;
	if(s->device >= 0) cudaSetDevice(s->device);
	if(s->stream) { cudaStreamSynchronize(s->stream); }
	s->zones.clear();
	s->halo_groups.clear();
	s->peers.clear();
	if(s->d_err) cudaFree(s->d_err);
	if(s->h_err) cudaFreeHost(s->h_err);
	if(s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return 0;
}

int gcm_b200_set_seed(gcm_b200_set* s, unsigned seed)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	s->rng.seed(seed);
	s->rng_ready = false;
	return 0;
}

int gcm_b200_set_kernel_mode(gcm_b200_set* s, int fast_inner)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	s->fast_inner = fast_inner != 0;
	return 0;
}

int gcm_b200_set_basis_mode(gcm_b200_set* s, int on_device)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	s->basis_on_device = on_device != 0;
	s->rng_ready = false;
	return 0;
}

int gcm_b200_set_collision_detector(gcm_b200_set* s, int type, int is_static, float threshold)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	if(type != 0) return fail(gcmh::UNIMPLEMENTED_EXCEPTION, "only VoidCollisionDetector is built so far");
	s->detector_type = type; s->detector_static = is_static; s->detector_threshold = threshold;
	return 0;
}

int gcm_b200_set_contact_calculator(gcm_b200_set* s, int type)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	s->contact_calc = type;
	return 0;
}

int gcm_b200_zone_load(gcm_b200_set* s, int zone, int n_nodes, int n_tets, const float* coords,
                       const int* placement, const int* remote_zone, const int* remote_num,
                       const int* tets, const float* material)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z) return fail(gcmh::CONFIG_EXCEPTION, "zone is not local to this process");
	if(n_nodes < 0 || n_tets < 0 || !coords || !tets) return fail(gcmh::INPUT_EXCEPTION, "bad mesh arrays");
	HostMesh& hm = z->hm;
	hm.zone = zone; hm.N = n_nodes; hm.T = n_tets;
	hm.coords.assign(coords, coords + 3*(size_t)n_nodes);
	hm.tets.assign(tets, tets + 4*(size_t)n_tets);
	hm.flags.resize(n_nodes); hm.remote_zone.assign(n_nodes, -1); hm.remote_num.assign(n_nodes, -1);
	for(int i = 0; i < n_nodes; i++) {
		bool local = placement ? placement[i] != 0 : true;
		// load_msh_file: setIsBorder(false); addOwner(GCM); setContactType(Free); setPlacement(...) -> Used
		hm.flags[i] = gcm::NF_GCM | gcm::NF_USED | (local ? gcm::NF_LOCAL : 0);
		if(!local) {
			hm.remote_zone[i] = remote_zone ? remote_zone[i] : -1;
			hm.remote_num[i] = remote_num ? remote_num[i] : -1;
			for(int k = 0; k < 3; k++) hm.coords[3*(size_t)i + k] = 0;   // arrive via sync_nodes
		}
	}
	hm.mat.resize(4*(size_t)n_nodes);
	if(material) memcpy(hm.mat.data(), material, 4*(size_t)n_nodes * sizeof(float));
	else for(int i = 0; i < n_nodes; i++) { hm.mat[4*(size_t)i] = 0; hm.mat[4*(size_t)i+1] = 0; hm.mat[4*(size_t)i+2] = 0; hm.mat[4*(size_t)i+3] = 0; }
	hm.values.assign(9*(size_t)n_nodes, 0.f);
	hm.current_time = 0;
	z->loaded = true;
	s->plan_built = false;
	return 0;
	GUARD_END
}

int gcm_b200_zone_translate(gcm_b200_set* s, int zone, float dx, float dy, float dz)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	if(s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "translate after preprocess is not supported");
	z->hm.translate(dx, dy, dz);
	return 0;
}

int gcm_b200_zone_set_values(gcm_b200_set* s, int zone, const float* values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !values) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	HostMesh& hm = z->hm;
	// the reference harness writes local nodes only; halo slots are refreshed by sync_nodes
	for(int i = 0; i < hm.N; i++)
		if(hm.flags[i] & gcm::NF_LOCAL)
			memcpy(&hm.values[9*(size_t)i], &values[9*(size_t)i], 9 * sizeof(float));
	if(z->on_device) { CK(cudaSetDevice(s->device)); upload_values(s, *z); }
	return 0;
	GUARD_END
}

int gcm_b200_zone_get_values(gcm_b200_set* s, int zone, float* values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !values) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	HostMesh& hm = z->hm;
	if(z->on_device) {
		CK(cudaSetDevice(s->device));
		const int n = hm.N;
		if(n) {
			if(!z->aos.p) z->aos.alloc(9 * (size_t)n);
			gcm::records_to_values_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(z->aos.p, z->ucur(), n);
			s->launches++;
			CK(cudaMemcpyAsync(hm.values.data(), z->aos.p, 9 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
			CK(cudaStreamSynchronize(s->stream));
		}
	}
	memcpy(values, hm.values.data(), 9*(size_t)hm.N * sizeof(float));
	return 0;
	GUARD_END
}

int gcm_b200_zone_upload_values_async(gcm_b200_set* s, int zone, const float* host_values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->on_device || !host_values) return fail(gcmh::CONFIG_EXCEPTION, "zone is not on the device");
	CK(cudaSetDevice(s->device));
	const int n = z->hm.N;
	if(!n) return 0;
	if(!z->aos.p) z->aos.alloc(9 * (size_t)n);
	CK(cudaMemcpyAsync(z->aos.p, host_values, 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
	gcm::values_to_records_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(z->ucur(), nullptr, z->aos.p, z->flags.p, n, 1);
	s->launches++;
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_zone_download_values_async(gcm_b200_set* s, int zone, float* host_values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->on_device || !host_values) return fail(gcmh::CONFIG_EXCEPTION, "zone is not on the device");
	CK(cudaSetDevice(s->device));
	const int n = z->hm.N;
	if(!n) return 0;
	if(!z->aos.p) z->aos.alloc(9 * (size_t)n);
	gcm::records_to_values_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(z->aos.p, z->ucur(), n);
	s->launches++;
	CK(cudaMemcpyAsync(host_values, z->aos.p, 9 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
	return 0;
	GUARD_END
}

int gcm_b200_set_add_border_condition(gcm_b200_set* s, int calc_type, const float* params,
                                      const int* vars, const float* vals,
                                      float start, float duration, const float* box)
{
	GUARD_BEGIN
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "call gcm_b200_set_preprocess first");
	if((int)s->conds.size() >= GCM_MAX_BORDER_CONDS) return fail(gcmh::CONFIG_EXCEPTION, "too many border conditions");
	if(calc_type < 0 || calc_type > 4 || !box) return fail(gcmh::CONFIG_EXCEPTION, "bad border condition");
	CondSpec cs;
	memset(&cs, 0, sizeof(cs));
	cs.bc.type = calc_type;
	if(calc_type == gcm::BC_EXT_FORCE || calc_type == gcm::BC_EXT_VELOCITY) {
		if(!params) return fail(gcmh::CONFIG_EXCEPTION, "missing calculator parameters");
		// set_parameters, ExternalForceCalculator.cpp:19-27
		cs.bc.p_normal = params[0]; cs.bc.p_tangential = params[1];
		float dtmp = sqrtf(params[2]*params[2] + params[3]*params[3] + params[4]*params[4]);
		cs.bc.dir[0] = params[2] / dtmp; cs.bc.dir[1] = params[3] / dtmp; cs.bc.dir[2] = params[4] / dtmp;
	}
	if(calc_type == gcm::BC_EXT_VALUES) {
		if(!vars || !vals) return fail(gcmh::CONFIG_EXCEPTION, "missing calculator parameters");
		for(int k = 0; k < 3; k++) {
			if(vars[k] < 0 || vars[k] > 8) return fail(gcmh::CONFIG_EXCEPTION, "bad variable index");
			cs.bc.vars_index[k] = vars[k]; cs.bc.vars_values[k] = vals[k];
		}
	}
	cs.start = start; cs.duration = duration;
	memcpy(cs.box, box, 6 * sizeof(float));
	int id = (int)s->conds.size();
	s->conds.push_back(cs);
	if(s->device >= 0) CK(cudaSetDevice(s->device));
	// BoxArea::isInArea (areas/BoxArea.cpp:15-26) over every node of every local mesh; a later
	// condition overrides an earlier one (TaskPreparator.cpp:432-438)
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		HostMesh& hm = z.hm;
		for(size_t b = 0; b < hm.border_nodes.size(); b++) {
			const float* c = &hm.coords[3*(size_t)hm.border_nodes[b]];
			if(c[0] < box[1] && c[0] > box[0] && c[1] < box[3] && c[1] > box[2] && c[2] < box[5] && c[2] > box[4])
				hm.cond_id[b] = id;
		}
		if(z.on_device) {
			z.cond_id.upload(hm.cond_id, s->stream);
			CK(cudaStreamSynchronize(s->stream));
		}
	}
	return 0;
	GUARD_END
}

int gcm_b200_set_preprocess(gcm_b200_set* s)
{
	GUARD_BEGIN
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	for(int zi : s->local_zones)
		if(!s->zones[zi]->loaded) return fail(gcmh::CONFIG_EXCEPTION, "a local zone was not loaded");
	build_halo_plan(s);
	host_sync_nodes(s);
	for(int zi : s->local_zones) s->zones[zi]->hm.pre_process_mesh();
	if(s->device >= 0) {
		CK(cudaSetDevice(s->device));
		for(int zi : s->local_zones) upload_zone(s, *s->zones[zi]);
	}
	s->preprocessed = true;
	return 0;
	GUARD_END
}

int gcm_b200_set_begin_step(gcm_b200_set* s)
{
	GUARD_BEGIN
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	if(s->device < 0) return fail(gcmh::CONFIG_EXCEPTION, "host-only set (device -1): there is no CPU execution path for the step");
	CK(cudaSetDevice(s->device));
	s->time_step = 0.002f;   // TetrMeshSet.cpp:105-106
	begin_step(s);
	return 0;
	GUARD_END
}

int gcm_b200_set_sync_nodes(gcm_b200_set* s)
{
	GUARD_BEGIN
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	CK(cudaSetDevice(s->device));
	sync_nodes_device(s);
	return 0;
	GUARD_END
}

int gcm_b200_zone_do_next_part_step(gcm_b200_set* s, int zone, float tau, int stage)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->on_device) return fail(gcmh::CONFIG_EXCEPTION, "zone is not on the device");
	if(stage < 0 || stage > 3) return fail(gcmh::METHOD_EXCEPTION, "Bad stage number");
	CK(cudaSetDevice(s->device));
	if(s->tap_on && !z->tap.p) { z->tap.alloc(3 * (size_t)z->hm.N); CK(cudaMemsetAsync(z->tap.p, 0xff, 3 * (size_t)z->hm.N * sizeof(gcm::StageTap), s->stream)); }
	zone_stage(s, zone, tau, stage);
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_set_end_step(gcm_b200_set* s)
{
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	// proceed_rheology: VoidRheologyCalculator (no-op); calc_destruction_criterias: evaluated on
	// download; update_current_time (TetrMesh_1stOrder.cpp:2403-2406)
	for(int zi : s->local_zones) s->zones[zi]->hm.current_time += s->time_step;
	s->in_step = false;
	return 0;
}

int gcm_b200_set_do_next_step(gcm_b200_set* s)
{
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	if(!s->peers.empty()) return fail(gcmh::CONFIG_EXCEPTION, "multi-process set: drive the stages with gcm_b200_halo_* between them");
	int rc = gcm_b200_set_begin_step(s);
	if(rc) return rc;
	for(int stage = 0; stage < 4; stage++) {
		if((rc = gcm_b200_set_sync_nodes(s))) return rc;
		for(int zi : s->local_zones)
			if((rc = gcm_b200_zone_do_next_part_step(s, zi, s->time_step, stage))) return rc;
	}
	return gcm_b200_set_end_step(s);
}

int gcm_b200_set_do_steps(gcm_b200_set* s, int n_steps)
{
	for(int i = 0; i < n_steps; i++) { int rc = gcm_b200_set_do_next_step(s); if(rc) return rc; }
	return 0;
}

int gcm_b200_set_synchronize(gcm_b200_set* s)
{
	GUARD_BEGIN
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	CK(cudaSetDevice(s->device));
	CK(cudaMemcpyAsync(s->h_err, s->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[0]) {
		char buf[256];
		snprintf(buf, sizeof buf, "%s (zone %d, node %d, stage %d)", step_error_text(s->h_err[0]), s->h_err[1], s->h_err[2], s->h_err[3]);
		int exc = step_error_exc(s->h_err[0]);
		CK(cudaMemsetAsync(s->d_err, 0, 4 * sizeof(int), s->stream));
		return fail(exc, buf);
	}
	return 0;
	GUARD_END
}

float gcm_b200_set_current_time(gcm_b200_set* s)
{
	if(!s || s->local_zones.empty()) return -1;
	return s->zones[s->local_zones[0]]->hm.current_time;
}

void* gcm_b200_set_stream(gcm_b200_set* s) { return s ? (void*)s->stream : nullptr; }

// ---- halo exchange between processes ----------------------------------------------------
int gcm_b200_halo_counts(gcm_b200_set* s, int peer_rank, int* n_recv, int* n_send)
{
	GUARD_BEGIN
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	build_halo_plan(s);
	auto it = s->peers.find(peer_rank);
	if(n_recv) *n_recv = it == s->peers.end() ? 0 : (int)it->second.recv_node.size();
	if(n_send) *n_send = it == s->peers.end() ? 0 : (int)it->second.send_node.size();
	return 0;
	GUARD_END
}

int gcm_b200_halo_get_requests(gcm_b200_set* s, int peer_rank, int* pairs)
{
	GUARD_BEGIN
	if(!s || !pairs) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	build_halo_plan(s);
	auto it = s->peers.find(peer_rank);
	if(it != s->peers.end() && !it->second.req_pairs.empty())
		memcpy(pairs, it->second.req_pairs.data(), it->second.req_pairs.size() * sizeof(int));
	return 0;
	GUARD_END
}

int gcm_b200_halo_set_sends(gcm_b200_set* s, int peer_rank, int n_send, const int* pairs)
{
	GUARD_BEGIN
	if(!s || (n_send && !pairs)) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	build_halo_plan(s);
	PeerPlan& p = s->peers[peer_rank];
	p.send_zone.clear(); p.send_node.clear(); p.dev_ready = false;
	for(int k = 0; k < n_send; k++) {
		int z = pairs[2*k], n = pairs[2*k+1];
		Zone* zz = get_zone(s, z);
		if(!zz || n < 0 || n >= zz->hm.N) return fail(gcmh::SYNC_EXCEPTION, "peer requested a node this process does not own");
		p.send_zone.push_back(z); p.send_node.push_back(n);
	}
	return 0;
	GUARD_END
}

int gcm_b200_halo_pack_host(gcm_b200_set* s, int peer_rank, float* buf16)
{
	GUARD_BEGIN
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	PeerPlan& p = it->second;
	for(size_t k = 0; k < p.send_node.size(); k++) {
		HostMesh& hm = s->zones[p.send_zone[k]]->hm;
		int n = p.send_node[k];
		float* o = buf16 + 16 * k;
		memcpy(o, &hm.coords[3*(size_t)n], 3 * sizeof(float));
		memcpy(o + 3, &hm.values[9*(size_t)n], 9 * sizeof(float));
		memcpy(o + 12, &hm.mat[4*(size_t)n], 4 * sizeof(float));
	}
	return 0;
	GUARD_END
}

int gcm_b200_halo_unpack_host(gcm_b200_set* s, int peer_rank, const float* buf16)
{
	GUARD_BEGIN
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	PeerPlan& p = it->second;
	for(size_t k = 0; k < p.recv_node.size(); k++) {
		HostMesh& hm = s->zones[p.recv_zone[k]]->hm;
		int n = p.recv_node[k];
		const float* o = buf16 + 16 * k;
		memcpy(&hm.coords[3*(size_t)n], o, 3 * sizeof(float));
		memcpy(&hm.values[9*(size_t)n], o + 3, 9 * sizeof(float));
		memcpy(&hm.mat[4*(size_t)n], o + 12, 4 * sizeof(float));
	}
	return 0;
	GUARD_END
}

static void peer_dev_prepare(gcm_b200_set* s, PeerPlan& p)
{
	if(p.dev_ready) return;
	auto build = [&](const std::vector<int>& zs, const std::vector<int>& ns, std::vector<std::unique_ptr<PeerPlan::Seg>>& segs) {
		segs.clear();
		size_t k = 0;
		while(k < zs.size()) {
			size_t e = k;
			while(e < zs.size() && zs[e] == zs[k]) e++;
			std::unique_ptr<PeerPlan::Seg> sg(new PeerPlan::Seg());
			sg->zone = zs[k]; sg->offset = (int)k; sg->n = (int)(e - k);
			std::vector<int> idx(ns.begin() + k, ns.begin() + e);
			sg->idx.upload(idx, s->stream);
			CK(cudaStreamSynchronize(s->stream));
			segs.push_back(std::move(sg));
			k = e;
		}
	};
	build(p.recv_zone, p.recv_node, p.recv_segs);
	build(p.send_zone, p.send_node, p.send_segs);
	p.dev_ready = true;
}

int gcm_b200_halo_pack(gcm_b200_set* s, int peer_rank, float* dev_buf)
{
	GUARD_BEGIN
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	CK(cudaSetDevice(s->device));
	PeerPlan& p = it->second;
	peer_dev_prepare(s, p);
	int total = (int)p.send_node.size();
	for(auto& sg : p.send_segs) {
		Zone& z = *s->zones[sg->zone];
		gcm::halo_pack_kernel<<<(sg->n + 255) / 256, 256, 0, s->stream>>>(dev_buf, z.ucur(), sg->idx.p, sg->n, total, sg->offset);
		s->launches++;
	}
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_halo_unpack(gcm_b200_set* s, int peer_rank, const float* dev_buf)
{
	GUARD_BEGIN
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	CK(cudaSetDevice(s->device));
	PeerPlan& p = it->second;
	peer_dev_prepare(s, p);
	int total = (int)p.recv_node.size();
	for(auto& sg : p.recv_segs) {
		Zone& z = *s->zones[sg->zone];
		gcm::halo_unpack_kernel<<<(sg->n + 255) / 256, 256, 0, s->stream>>>(z.ucur(), dev_buf, sg->idx.p, sg->n, total, sg->offset);
		s->launches++;
	}
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

// ---- inspection -----------------------------------------------------------------------------
int gcm_b200_zone_counts(gcm_b200_set* s, int zone, int* n_nodes, int* n_tets, int* n_faces,
                         int* n_border_nodes, int* n_local_nodes)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	HostMesh& hm = z->hm;
	if(n_nodes) *n_nodes = hm.N;
	if(n_tets) *n_tets = hm.T;
	if(n_faces) *n_faces = (int)(hm.faces.size() / 3);
	if(n_border_nodes) *n_border_nodes = (int)hm.border_nodes.size();
	if(n_local_nodes) { int c = 0; for(int i = 0; i < hm.N; i++) c += hm.is_local(i) ? 1 : 0; *n_local_nodes = c; }
	return 0;
}

int gcm_b200_zone_get_flags(gcm_b200_set* s, int zone, int* flags)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !flags) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	memcpy(flags, z->hm.flags.data(), z->hm.N * sizeof(int));
	return 0;
}

int gcm_b200_zone_get_faces(gcm_b200_set* s, int zone, int* faces)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !faces) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	if(!z->hm.faces.empty()) memcpy(faces, z->hm.faces.data(), z->hm.faces.size() * sizeof(int));
	return 0;
}

int gcm_b200_zone_get_bases(gcm_b200_set* s, int zone, float* bases)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->hm.preprocessed || !bases) return fail(gcmh::CONFIG_EXCEPTION, "zone not preprocessed");
	HostMesh& hm = z->hm;
	if(z->on_device && s->basis_on_device && !hm.basis.empty()) {
		cudaSetDevice(s->device);
		if(cudaMemcpyAsync(hm.basis.data(), z->basis.p, hm.basis.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess
		   || cudaStreamSynchronize(s->stream) != cudaSuccess)
			return fail(gcmh::CONFIG_EXCEPTION, "CUDA: reading the bases failed");
	}
	static const float E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	for(int i = 0; i < hm.N; i++) {
		float* o = bases + 9*(size_t)i;
		if(!hm.is_local(i)) { memset(o, 0, 9 * sizeof(float)); continue; }
		int b = hm.border_slot[i];
		memcpy(o, b >= 0 ? &hm.basis[9*(size_t)b] : E, 9 * sizeof(float));
	}
	return 0;
}

int gcm_b200_zone_get_min_max_h(gcm_b200_set* s, int zone, float* min_h, float* max_h)
{
	Zone* z = get_zone(s, zone);
	if(!z || !z->hm.preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "zone not preprocessed");
	if(min_h) *min_h = z->hm.min_h;
	if(max_h) *max_h = z->hm.max_h;
	return 0;
}

int gcm_b200_set_enable_tap(gcm_b200_set* s, int on) { if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set"); s->tap_on = on != 0; return 0; }

int gcm_b200_zone_get_tap(gcm_b200_set* s, int zone, int* owners, int* outer, int* tries)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->tap.p) return fail(gcmh::CONFIG_EXCEPTION, "tap was not enabled");
	CK(cudaSetDevice(s->device));
	size_t n = 3 * (size_t)z->hm.N;
	std::vector<gcm::StageTap> h(n);
	CK(cudaMemcpyAsync(h.data(), z->tap.p, n * sizeof(gcm::StageTap), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	for(size_t i = 0; i < n; i++) {
		if(owners) for(int k = 0; k < 5; k++) owners[5*i + k] = h[i].owner[k];
		if(outer) outer[i] = h[i].outer_count;
		if(tries) tries[i] = h[i].tries;
	}
	return 0;
	GUARD_END
}

long long gcm_b200_set_launch_count(gcm_b200_set* s) { return s ? s->launches : 0; }

int gcm_b200_set_profile(gcm_b200_set* s, int on)
{
	if(!s) return fail(gcmh::CONFIG_EXCEPTION, "null set");
	cudaSetDevice(s->device);
	prof_drain(s);
	s->profile = on != 0;
	if(on) for(int k = 0; k < 6; k++) { s->prof_ms[k] = 0; s->prof_launches[k] = 0; s->prof_units[k] = 0; }
	return 0;
}

int gcm_b200_set_profile_read(gcm_b200_set* s, int which, double* total_ms, long long* launches, long long* units)
{
	if(!s || which < 0 || which > 5) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	cudaSetDevice(s->device);
	prof_drain(s);
	if(total_ms) *total_ms = s->prof_ms[which];
	if(launches) *launches = s->prof_launches[which];
	if(units) *units = s->prof_units[which];
	return 0;
}
int gcm_b200_set_virt_count(gcm_b200_set* s) { (void)s; return 0; }

} // extern "C"
