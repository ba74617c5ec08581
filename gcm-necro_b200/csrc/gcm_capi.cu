// gcm_capi.cu -- the C ABI of include/gcm_b200.h: TetrMeshSet-level orchestration
// (gcm/system/TetrMeshSet.cpp:101-184) over the kernels of gcm_kernels.cuh / gcm_inner.cuh and
// the host-side mesh logic of gcm_host.cpp.
//
// Device layout: the zones a process owns are concatenated into ONE device mesh (node, tet,
// CSR-entry and border-slot indices shifted by per-zone offsets).  Within a stage the zones are
// independent (Jacobi; TetrMeshSet.cpp:148-167 syncs the halo before the per-mesh loop), so one
// launch per kernel class covers all of them; the halo between two zones of the same process is
// an index-to-index copy inside the merged arrays (DataBus.cpp:64-79).
#include "../../include/gcm_b200.h"
#include "gcm_host.h"
#include "gcm_kernels.cuh"

#include <dlfcn.h>
#include <algorithm>
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
	void upload(const std::vector<T>& h, cudaStream_t st) { alloc(h.size()); put(h, 0, st); }
	// copy h into [offset, offset + h.size()) and wait (h may be a temporary)
	void put(const std::vector<T>& h, size_t offset, cudaStream_t st) {
		if(h.empty()) return;
		CK(cudaMemcpyAsync(p + offset, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
	}
	void free() { if(p) cudaFree(p); p = nullptr; n = 0; }
	~DevBuf() { free(); }
};

struct CondSpec {   // one BorderCondition: calculator + StepPulseForm + BoxArea
	gcm::BorderCond bc;
	float start, duration;
	float box[6];
};

struct PeerPlan {   // exchange with one other process
	std::vector<int> recv_zone, recv_node;       // my halo slots, in request order
	std::vector<int> req_pairs;                  // (owner zone, owner node) per slot
	std::vector<int> send_zone, send_node;       // my nodes the peer asked for
	DevBuf<int> d_recv, d_send;                  // merged-mesh indices
	DevBuf<float> send_buf, recv_buf;            // [9][n] staging of the library-driven (NCCL) exchange
	bool dev_ready = false;
	// direct peer stores: the neighbour's mapped buffers and where my send nodes live in its arrays
	void* map_u0 = nullptr; void* map_u1 = nullptr; void* map_flags = nullptr;    // what cudaIpcOpenMemHandle returned (to close)
	float* peer_u[2] = {nullptr, nullptr}; int* peer_flags = nullptr;             // the neighbour's arrays inside those mappings
	unsigned long long peer_rs = 0;
};

// NCCL entry points, resolved at run time (the library has no link-time dependency on NCCL and
// loads without it; a process that already carries NCCL -- torch does -- shares that copy).
struct NcclId { char internal[128]; };           // ncclUniqueId
struct NcclApi {
	void* lib = nullptr;
	int (*GetUniqueId)(NcclId*) = nullptr;
	int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};

struct Zone {
	HostMesh hm;
	bool loaded = false;
	size_t node_off = 0, tet_off = 0, entry_off = 0, slot_off = 0;
	size_t fast_off = 0, generic_off = 0, inner_off = 0;   // ranges of this zone in the merged node lists
	size_t st0_off = 0, st0_n = 0, c0_off = 0, c0_n = 0, c1_off = 0, c1_n = 0;   // contact launch lists
	size_t border_off = 0;                       // range in the merged border list
	bool stage_done = false;
	// pipelined host <-> device traffic of the *_async calls (see gcm_b200_zone_upload_values_async)
	cudaEvent_t ev_h2d_done = nullptr, ev_in_consumed = nullptr, ev_out_ready = nullptr, ev_d2h_done = nullptr;
};

// The merged device mesh of every local zone.
struct DeviceMesh {
	size_t N = 0, T = 0, E = 0, NB = 0, stride = 0, rs = 0;   // rs: quads per plane of the node records (gcm_step.cuh)
	size_t n_fast = 0, n_generic = 0, n_inner = 0;
	DevBuf<float> u0, u1, mat, tet_hv, basis, basis_next, normals, w0, aos, aos_out, mat_tab;
	DevBuf<int> flags, n2t_off, n2t, tets, border_slot, cond_id, n2x, cold;
	// step-invariant cache of the inner kernel (gcm_inner.cuh): 3 planes of N entries, built lazily for a tau
	DevBuf<gcm::InnerCacheEntry> icache;
	DevBuf<int> icache_owner;        // [3][N][2] owner tets per direction, only kept when the parity tap is on
	// the cache de-duplicated into patterns (per axis; pattern_ok[axis] says whether the axis compressed)
	DevBuf<unsigned> ic_pat_idx;     // [3][N] pattern id per node or GCM_IC_NO_PATTERN
	DevBuf<gcm::InnerCacheEntry> ic_patterns;   // [3][GCM_IC_MAX_PATTERNS]
	bool pattern_ok[3] = {false, false, false};
	DevBuf<int> shell;               // [1 + n]: inner nodes whose stencil touches another process's halo node (gcm_inner.cuh: GCM_IC_SHELL)
	int n_shell = 0;
	int n_patterns[3] = {0, 0, 0};
	float icache_tau = -1.f;
	bool icache_ready = false;
	int n_cold = 0;                  // inner nodes that stay with the literal routine (cold[0] on the device)
	// static tables of the flat-border kernel (gcm_border.cuh) and the list of nodes it defers per launch
	DevBuf<int> bnd_static, bnd_defer;
	DevBuf<float> bnd_cand;
	float pf_fail = -1e-3f, pf_pass = -1e-4f;   // thresholds of the conservative owner pre-filter (prefilter_thresholds)
	int exact_only = 0;
	DevBuf<int> border_list, fast_list, generic_list, inner_list;
	size_t n_sharp = 0, n_flat = 0;
	DevBuf<unsigned char> mat_id;
	DevBuf<float> destr_hist;        // [N] float4 running maxima of calc_destruction_criterias, when tracked
	DevBuf<gcm::StageTap> tap;
	DevBuf<int> halo_dst, halo_src;
	size_t n_halo = 0;
	int n_mat = 0;
	int cur = 0;
	bool ready = false;
	float* ucur() { return cur ? u1.p : u0.p; }
	float* unext() { return cur ? u0.p : u1.p; }
};

enum ProfClass { PROF_INNER = 0, PROF_BORDER = 1, PROF_PLASTIC = 2, PROF_HALO = 3, PROF_RNG = 4, PROF_INNER_GENERIC = 5, PROF_N = 6 };

} // namespace

struct gcm_b200_set {
	int device = 0, n_zones = 0, my_rank = 0;
	std::vector<int> zone_rank;
	std::vector<std::unique_ptr<Zone>> zones;   // indexed by zone number; null when remote
	std::vector<int> local_zones;               // TetrMeshSet::local_meshes order
	gcmh::GlibcRand rng;
	gcmh::MaterialRegistry materials;
	DeviceMesh dm;
	cudaStream_t stream = nullptr, aux = nullptr, aux2 = nullptr, rng_stream = nullptr, io_in = nullptr, io_out = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_rng_done = nullptr, ev_step_mark = nullptr, ev_halo = nullptr;
	bool overlap = true;         // multi-process whole-step calls: the halo exchange beside the inner nodes that do not need it
	bool bases_ahead = true;     // device generator: the NEXT step's bases are produced on rng_stream during this step
	bool ahead_valid = false;    // dm.basis_next holds the bases of the coming step
	int* d_err = nullptr;
	int* h_err = nullptr;
	std::vector<CondSpec> conds;
	std::vector<int> halo_dst, halo_src;        // same-process halo, merged indices
	std::map<int, PeerPlan> peers;
	void* nccl_comm = nullptr;   // library-driven halo exchange (gcm_b200_set_comm_init); NULL = the host moves the halos
	int comm_rank = -1, comm_size = 0;
	// direct peer-store exchange (gcm_kernels.cuh: halo_push_kernel / halo_wait_kernel), set up by comm_init
	bool p2p = true, p2p_ready = false;
	bool p2p_dma_wanted = false, p2p_dma = false;   // contiguous halo layers: records by cudaMemcpyAsync (copy engines) instead of the push kernel
	std::vector<int> p2p_dma_ranges;              // per neighbour: first send node, first slot over there
	int p2p_seq = 0;
	DevBuf<int> p2p_flags, p2p_send_node, p2p_remote_slot, p2p_peer_ranks;
	DevBuf<unsigned> p2p_counter;
	DevBuf<gcm::PeerTarget> p2p_more;
	std::vector<gcm::PeerTarget> p2p_targets;
	bool plan_built = false;
	bool preprocessed = false;
	bool tap_on = false;
	int detector_type = 0, detector_static = 0; float detector_threshold = 1;
	int contact_calc = 0;
	long long launches = 0;
	float time_step = 0.002f;
	bool fast_inner = true;     // specialised inner-node kernel (gcm_inner.cuh)
	bool concurrent = true;     // border kernel on a second stream, next to the inner kernel
	bool flat_border = true;     // border nodes through gcm_border.cuh (0 = the literal per-node routine for every border node)
	bool inner_cache = true;     // inner nodes through the step-invariant owner/factor cache (gcm_inner.cuh); 0 = index-free scan kernel
	bool inner_patterns = true;  // de-duplicate the cache into a pattern table where an axis has few distinct entries
	bool track_destruction = false;   // calc_destruction_criterias at the end of every step (its running maxima need every step)
	int border_lu_block = 64;    // block size of the stage-0 border kernel: 64 (five blocks per SM), or 352 (one block on all of an SM's
	                             // shared memory: 10 % more systems per wave, but nothing else fits beside it -- slower at 8 GPUs, 0.333 vs 0.301 ms)
	int inner_mb = 5;            // register budget of the pattern kernel: resident 256-thread blocks per SM aimed at (4, 5, 6; anything else: 1 block, up to 255 registers); 5 = 48 registers measured fastest at 216^3 (1.47 vs 1.49 / 1.53 ms per step)
	bool fuse_plastic = true;    // whole-step calls: stage 3 as the epilogue of the stage-2 writers (and no exchange before it)
	// contact: virtual nodes of the last collision pass and the launch lists derived from them
	std::vector<gcmh::VirtNodeHost> virt;
	std::vector<std::vector<int> > axis_plus0;   // per local zone (local_meshes order), per border slot
	DevBuf<gcm::VirtNode> d_virt;
	DevBuf<int> d_axis_plus0, d_border_stage0, d_contact_nodes, d_border_by_slot;
	DevBuf<int> d_deep, d_ring_ws;   // contact nodes deferred to the list-building owner search, and its workspaces
	DevBuf<float> d_gather;
	size_t st0_total = 0, c0_total = 0;
	bool contact_active = false;
	// collision discovery on the device (gcm_collide.cuh); the host detector stays as the A/B (GCM_B200_DEVICE_COLLISIONS=0)
	bool coll_device = true, coll_ready = false, coll_host_stale = false;
	int coll_n_cand = 0;
	std::vector<int> coll_zone_off, coll_zone_n;     // per local zone (local_meshes order): its candidate range
	DevBuf<gcm::CollPair> coll_pairs;
	DevBuf<gcm::CollFace> coll_faces;
	DevBuf<int> coll_cand_pair, coll_cand_node, coll_cand_slot, coll_cand_phase, coll_grid_start, coll_grid_items, coll_hit_l, coll_hit_virt, coll_n_virt;
	DevBuf<float> coll_hit_p;
	long long step_count = 0;
	// device-side rand() stream + bases (rng_chunk_kernel / basis_kernel)
	bool basis_on_device = true;
	bool rng_ready = false;
	long long rng_draws_per_step = 0;
	int rng_chunks = 0;
	uint32_t rng_jump[961];
	DevBuf<uint32_t> d_chunk_state, d_jump, d_slot_pos;
	DevBuf<int> d_chunk_first_slot;
	DevBuf<unsigned short> d_teta_idx;
	DevBuf<float> d_cos, d_sin;
	float* h_basis[2] = {nullptr, nullptr};     // host-mode bases: pinned staging, double buffered
	cudaEvent_t basis_ev[2] = {nullptr, nullptr};
	int basis_flip = 0;
	// per-kernel CUDA-event profile (gcm_b200_set_profile)
	bool profile = false;
	struct ProfRec { int which; long long units; cudaEvent_t a, b; };
	std::vector<ProfRec> prof;
	double prof_ms[PROF_N] = {0, 0, 0, 0, 0, 0};
	long long prof_launches[PROF_N] = {0, 0, 0, 0, 0, 0}, prof_units[PROF_N] = {0, 0, 0, 0, 0, 0};
};

namespace {

Zone* get_zone(gcm_b200_set* s, int zone)
{
	if(!s || zone < 0 || zone >= s->n_zones || !s->zones[zone]) return nullptr;
	return s->zones[zone].get();
}

struct ProfScope {
	gcm_b200_set* s; int which; long long units; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
	ProfScope(gcm_b200_set* s_, int w, long long u, cudaStream_t st_ = nullptr) : s(s_), which(w), units(u), st(st_ ? st_ : s_->stream) {
		if(!s->profile) return;
		CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
		CK(cudaEventRecord(a, st));
	}
	~ProfScope() {
		if(!a) return;
		cudaEventRecord(b, st);
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

void assign_offsets(gcm_b200_set* s)
{
	size_t n = 0, t = 0, e = 0, b = 0;
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		z.node_off = n; z.tet_off = t; z.entry_off = e; z.slot_off = b;
		n += z.hm.N; t += z.hm.T; e += z.hm.n2t.size(); b += z.hm.border_nodes.size();
	}
}

// Halo plan from the placement flags: every Remote node is refreshed from its owner before each
// stage (the reference builds its index lists before preprocessing too, main.cpp:124-130).
void build_halo_plan(gcm_b200_set* s)
{
	if(s->plan_built) return;
	s->halo_dst.clear(); s->halo_src.clear();
	s->peers.clear();
	size_t n = 0;
	for(int zi : s->local_zones) { s->zones[zi]->node_off = n; n += s->zones[zi]->hm.N; }
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		HostMesh& hm = z.hm;
		for(int i = 0; i < hm.N; i++) {
			if(hm.flags[i] & gcm::NF_LOCAL) continue;
			int oz = hm.remote_zone[i], on = hm.remote_num[i];
			if(oz < 0 || oz >= s->n_zones) throw HostError{gcmh::INPUT_EXCEPTION, "remote node refers to an unknown zone"};
			int owner_rank = s->zone_rank[oz];
			if(owner_rank == s->my_rank) {
				Zone* o = s->zones[oz].get();
				if(!o || !o->loaded || on < 0 || on >= o->hm.N) throw HostError{gcmh::INPUT_EXCEPTION, "remote node index out of range"};
				s->halo_dst.push_back((int)(z.node_off + i));
				s->halo_src.push_back((int)(o->node_off + on));
			} else {
				PeerPlan& p = s->peers[owner_rank];
				p.recv_zone.push_back(zi); p.recv_node.push_back(i);
				p.req_pairs.push_back(oz); p.req_pairs.push_back(on);
			}
		}
	}
	s->plan_built = true;
}

// merged index -> (host mesh, local index)
HostMesh& locate(gcm_b200_set* s, int merged, int* local)
{
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		if((size_t)merged >= z.node_off && (size_t)merged < z.node_off + z.hm.N) { *local = merged - (int)z.node_off; return z.hm; }
	}
	throw HostError{gcmh::INPUT_EXCEPTION, "bad merged node index"};
}

// DataBus::sync_nodes, same-process branch, on the host mirrors (start-up: coords + 13 values).
void host_sync_nodes(gcm_b200_set* s)
{
	for(size_t k = 0; k < s->halo_dst.size(); k++) {
		int di, si;
		HostMesh& d = locate(s, s->halo_dst[k], &di);
		HostMesh& o = locate(s, s->halo_src[k], &si);
		memcpy(&d.coords[3*(size_t)di], &o.coords[3*(size_t)si], 3 * sizeof(float));
		memcpy(&d.values[9*(size_t)di], &o.values[9*(size_t)si], 9 * sizeof(float));
		memcpy(&d.mat[4*(size_t)di], &o.mat[4*(size_t)si], 4 * sizeof(float));
	}
}

// Host node-major values of one zone -> both device layers (start-up; includes halo slots).
void upload_values(gcm_b200_set* s, Zone& z)
{
	DeviceMesh& dm = s->dm;
	const int n = z.hm.N;
	if(!n) return;
	float* stage = dm.aos.p + 9 * z.node_off;
	if(s->io_in) CK(cudaStreamSynchronize(s->io_in));
	CK(cudaMemcpyAsync(stage, z.hm.values.data(), 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
	gcm::values_to_records_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(dm.u0.p + 4 * z.node_off, dm.u1.p + 4 * z.node_off, dm.rs,
		stage, dm.flags.p + z.node_off, n, 0);
	s->launches++;
	CK(cudaStreamSynchronize(s->stream));
}

template<class T> std::vector<T> shifted(const std::vector<T>& v, long long off)
{
	std::vector<T> r(v.size());
	for(size_t i = 0; i < v.size(); i++) r[i] = (T)(v[i] + off);
	return r;
}

void prefilter_thresholds(gcm_b200_set* s)
{
	std::vector<const HostMesh*> ms;
	for(int zi : s->local_zones) ms.push_back(&s->zones[zi]->hm);
	gcmh::prefilter_thresholds(ms, &s->dm.pf_fail, &s->dm.pf_pass, &s->dm.exact_only);
}

void upload_mesh(gcm_b200_set* s)
{
	DeviceMesh& dm = s->dm;
	cudaStream_t st = s->stream;
	assign_offsets(s);
	size_t N = 0, T = 0, E = 0, NB = 0, nf = 0, ng = 0, ni = 0, nsh = 0, nfl = 0;
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		z.fast_off = nf; z.generic_off = ng; z.inner_off = ni; z.border_off = nsh + nfl;
		nsh += z.hm.n_sharp; nfl += z.hm.border_launch.size() - z.hm.n_sharp;
		N += z.hm.N; T += z.hm.T; E += z.hm.n2t.size(); NB += z.hm.border_nodes.size();
		nf += z.hm.inner_fast.size(); ng += z.hm.inner_generic.size(); ni += z.hm.inner_nodes.size();
	}
	if(T >= (1u << 30) || E >= 0x7fffffffull || N >= 0x7fffffffull / GCM_REC * 4) throw HostError{gcmh::CONFIG_EXCEPTION, "mesh too large for 32-bit device indices"};
	dm.N = N; dm.T = T; dm.E = E; dm.NB = NB; dm.n_fast = nf; dm.n_generic = ng; dm.n_inner = ni; dm.n_sharp = nsh; dm.n_flat = nfl;
	dm.stride = (N + 31) / 32 * 32;
	dm.rs = N + 1;
	dm.u0.alloc(GCM_REC * (N + 1)); dm.u1.alloc(GCM_REC * (N + 1));
	CK(cudaMemsetAsync(dm.u0.p, 0, GCM_REC * (N + 1) * sizeof(float), st));
	CK(cudaMemsetAsync(dm.u1.p, 0, GCM_REC * (N + 1) * sizeof(float), st));
	dm.mat.alloc(4 * dm.stride); dm.tet_hv.alloc(2 * T); dm.basis.alloc(9 * NB); dm.normals.alloc(3 * NB);
	dm.w0.alloc(N); dm.aos.alloc(9 * N); dm.flags.alloc(N); dm.n2t_off.alloc(N + 1); dm.n2t.alloc(E); dm.tets.alloc(4 * T);
	dm.border_slot.alloc(N); dm.cond_id.alloc(NB); dm.n2x.alloc(4 * E); dm.mat_id.alloc(N);
	dm.border_list.alloc(nsh + nfl); dm.fast_list.alloc(nf); dm.generic_list.alloc(ng); dm.inner_list.alloc(ni);
	if(dm.basis.p) CK(cudaMemsetAsync(dm.basis.p, 0, 9 * NB * sizeof(float), st));
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		HostMesh& hm = z.hm;
		const size_t n = hm.N;
		// node records: coordinates (static) + values
		{
			DevBuf<float> xyz; xyz.upload(hm.coords, st);
			if(n) {
				gcm::coords_to_records_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(dm.u0.p + 4 * z.node_off, dm.u1.p + 4 * z.node_off, dm.rs, xyz.p, (int)n);
				s->launches++;
			}
			CK(cudaStreamSynchronize(st));
		}
		for(int k = 0; k < 4; k++) {
			std::vector<float> pl(n);
			for(size_t i = 0; i < n; i++) pl[i] = hm.mat[4*i + k];
			dm.mat.put(pl, k * dm.stride + z.node_off, st);
		}
		dm.flags.put(hm.flags, z.node_off, st);
		{
			std::vector<int> off(hm.n2t_off.begin(), hm.n2t_off.end() - 1);
			dm.n2t_off.put(shifted(off, (long long)z.entry_off), z.node_off, st);
		}
		dm.n2t.put(shifted(hm.n2t, (long long)z.tet_off), z.entry_off, st);
		dm.tets.put(shifted(hm.tets, (long long)z.node_off), 4 * z.tet_off, st);
		dm.tet_hv.put(hm.tet_hv, 2 * z.tet_off, st);
		{
			std::vector<int> bs(hm.border_slot);
			for(auto& v : bs) if(v >= 0) v += (int)z.slot_off;
			dm.border_slot.put(bs, z.node_off, st);
		}
		dm.cond_id.put(hm.cond_id, z.slot_off, st);
		{
			std::vector<int> x(hm.n2x);
			for(size_t e = 0; e < x.size() / 4; e++) {
				x[4*e] += (int)z.node_off; x[4*e+1] += (int)z.node_off; x[4*e+2] += (int)z.node_off;
				unsigned w = (unsigned)x[4*e+3];
				x[4*e+3] = (int)((((w & 0x3fffffffu) + (unsigned)z.tet_off) & 0x3fffffffu) | (w & 0xc0000000u));
			}
			dm.n2x.put(x, 4 * z.entry_off, st);
		}
		dm.w0.put(hm.w0, z.node_off, st);
		dm.mat_id.put(hm.mat_id, z.node_off, st);
		dm.normals.put(hm.normals, 3 * z.slot_off, st);
		dm.border_list.put(shifted(hm.border_launch, (long long)z.node_off), z.border_off, st);
		dm.fast_list.put(shifted(hm.inner_fast, (long long)z.node_off), z.fast_off, st);
		dm.generic_list.put(shifted(hm.inner_generic, (long long)z.node_off), z.generic_off, st);
		dm.inner_list.put(shifted(hm.inner_nodes, (long long)z.node_off), z.inner_off, st);
	}
	{
		std::vector<int> last(1, (int)E);
		dm.n2t_off.put(last, N, st);
	}
	dm.n_mat = (int)(s->materials.materials.size() / 3);
	dm.mat_tab.upload(s->materials.table, st);
	dm.halo_dst.upload(s->halo_dst, st); dm.halo_src.upload(s->halo_src, st);
	dm.n_halo = s->halo_dst.size();
	CK(cudaStreamSynchronize(st));
	dm.cur = 0;
	dm.ready = true;
	dm.cold.alloc(N + 1);   // inner nodes the cache could not take: count, then node ids
	if(NB) {
		std::vector<gcmh::BorderFlatZone> bz;
		for(int zi : s->local_zones) { Zone& z = *s->zones[zi]; bz.push_back(gcmh::BorderFlatZone{&z.hm, z.node_off, z.entry_off, z.slot_off}); }
		std::vector<int> st8; std::vector<float> cq;
		gcmh::build_border_flat_tables(bz, st8, cq);
		dm.bnd_static.upload(st8, st);
		if(cq.empty()) cq.assign(4, 0.f);
		dm.bnd_cand.upload(cq, st);
		dm.bnd_defer.alloc(NB + 1);
	}
	dm.icache_ready = false;
	prefilter_thresholds(s);
	for(int zi : s->local_zones) upload_values(s, *s->zones[zi]);
}

gcm::MeshView make_view(gcm_b200_set* s, float tau)
{
	DeviceMesh& dm = s->dm;
	gcm::MeshView m;
	memset(&m, 0, sizeof(m));
	m.n_nodes = (int)dm.N; m.n_tets = (int)dm.T; m.stride = dm.stride;
	m.u = dm.ucur(); m.rs = dm.rs; m.mat = dm.mat.p; m.flags = dm.flags.p;
	m.n2t_off = dm.n2t_off.p; m.n2t = dm.n2t.p; m.tets = dm.tets.p; m.tet_hv = dm.tet_hv.p;
	m.n2x = nullptr;
	m.pf_fail = dm.pf_fail; m.pf_pass = dm.pf_pass; m.exact_only = dm.exact_only;
	m.n2x_scan = dm.n2x.p;
	m.border_slot = dm.border_slot.p; m.basis = dm.basis.p; m.cond_id = dm.cond_id.p; m.contact = nullptr;
	const float t = s->zones[s->local_zones[0]]->hm.current_time;   // TetrMeshSet::get_current_time()
	m.tau = tau; m.time = t; m.min_h = 0;
	m.n_conds = (int)s->conds.size();
	for(int c = 0; c < m.n_conds; c++) {
		const CondSpec& cs = s->conds[c];
		m.conds[c] = cs.bc;
		// check_stresses (TetrMesh_1stOrder.cpp:2024-2025) + BorderCondition::do_calc's scale
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

// Lay the step's rand() stream out for the device: draw positions of the border nodes, chunk
// start windows from the current host generator state, the one-step jump matrix.
void rng_device_init(gcm_b200_set* s)
{
	std::vector<uint32_t> slot_pos;
	unsigned long long pos = 0;
	for(int zi : s->local_zones) {
		const HostMesh& hm = s->zones[zi]->hm;
		for(int i = 0; i < hm.N; i++) {
			if(!hm.is_local(i)) continue;
			if(hm.is_border(i)) { slot_pos.push_back((uint32_t)pos); pos += 1; }
			else pos += 3;
		}
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
	CK(cudaStreamSynchronize(s->stream));
	s->rng_ready = true;
}

void run_collision_detection(gcm_b200_set* s);

void begin_step_bases(gcm_b200_set* s)
{
	// get_max_possible_tau(): only its side effect survives: new local bases,
	// meshes in local_meshes order, nodes in index order, one libc rand() stream.
	DeviceMesh& dm = s->dm;
	for(int zi : s->local_zones) s->zones[zi]->stage_done = false;
	if(s->basis_on_device) {
		if(!s->rng_ready) { CK(cudaStreamSynchronize(s->rng_stream)); rng_device_init(s); s->ahead_valid = false; }
		if(s->rng_chunks && dm.NB) {
			// one call = the rand() draws of one step (the chunk states advance in place) + the bases from them
			auto produce = [&](float* dst, cudaStream_t st) {
				ProfScope ps(s, PROF_RNG, s->rng_draws_per_step, st);
				gcm::rng_chunk_kernel<<<(s->rng_chunks + 127) / 128, 128, 0, st>>>(s->d_chunk_state.p, s->rng_chunks,
					s->d_jump.p, s->d_slot_pos.p, s->d_chunk_first_slot.p, s->d_teta_idx.p);
				gcm::basis_kernel<<<(int)((dm.NB + 127) / 128), 128, 0, st>>>(dst, dm.normals.p, s->d_teta_idx.p, 0, (int)dm.NB, s->d_cos.p, s->d_sin.p);
				s->launches += 2;
			};
			if(s->ahead_valid) {
				// produced during the previous step (the stream depends on nothing but its own position)
				CK(cudaStreamWaitEvent(s->stream, s->ev_rng_done, 0));
				std::swap(dm.basis.p, dm.basis_next.p);
				s->ahead_valid = false;
			} else {
				produce(dm.basis.p, s->stream);
			}
			if(s->bases_ahead) {
				// the other buffer was read by the previous step's kernels: start after everything queued so far
				if(!dm.basis_next.p) dm.basis_next.alloc(9 * dm.NB);
				CK(cudaEventRecord(s->ev_step_mark, s->stream));
				CK(cudaStreamWaitEvent(s->rng_stream, s->ev_step_mark, 0));
				produce(dm.basis_next.p, s->rng_stream);
				CK(cudaEventRecord(s->ev_rng_done, s->rng_stream));
				s->ahead_valid = true;
			}
		}
		// keep the host generator in step with the device one
		uint32_t w[31];
		s->rng.get_window(w);
		gcmh::rng_apply_jump(s->rng_jump, w);
		s->rng.set_window(w);
		return;
	}
	if(!dm.NB) { for(int zi : s->local_zones) s->zones[zi]->hm.create_random_axes(s->rng); return; }
	if(!s->h_basis[0])
		for(int k = 0; k < 2; k++) {
			CK(cudaMallocHost(&s->h_basis[k], 9 * dm.NB * sizeof(float)));
			CK(cudaEventCreateWithFlags(&s->basis_ev[k], cudaEventDisableTiming));
		}
	const int b = s->basis_flip; s->basis_flip ^= 1;
	CK(cudaEventSynchronize(s->basis_ev[b]));
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		z.hm.create_random_axes(s->rng);
		if(!z.hm.basis.empty()) memcpy(s->h_basis[b] + 9 * z.slot_off, z.hm.basis.data(), z.hm.basis.size() * sizeof(float));
	}
	CK(cudaMemcpyAsync(dm.basis.p, s->h_basis[b], 9 * dm.NB * sizeof(float), cudaMemcpyHostToDevice, s->stream));
	CK(cudaEventRecord(s->basis_ev[b], s->stream));
}

void begin_step(gcm_b200_set* s)
{
	begin_step_bases(s);
	// TetrMeshSet.cpp:111-120: a static detector runs at t == 0 only
	if(s->detector_type != 0 && (!s->detector_static || s->zones[s->local_zones[0]]->hm.current_time == 0.0f))
		run_collision_detection(s);
}

void sync_nodes_device(gcm_b200_set* s)
{
	DeviceMesh& dm = s->dm;
	const int n = (int)dm.n_halo;
	if(!n) return;
	ProfScope ps(s, PROF_HALO, n);
	gcm::halo_copy_kernel<<<(3 * n + 255) / 256, 256, 0, s->stream>>>(dm.ucur(), dm.halo_dst.p, dm.ucur(), dm.halo_src.p, n, dm.rs);
	s->launches++;
}

// Builds the step-invariant cache of the inner kernel for time step tau (once; again if tau changes
// or the parity tap is switched on afterwards): the literal per-node routine over every inner node
// of the fast list, three axes.
void ensure_inner_cache(gcm_b200_set* s, float tau)
{
	DeviceMesh& dm = s->dm;
	const bool want_tap = s->tap_on;
	if(dm.icache_ready && dm.icache_tau == tau && (!want_tap || dm.icache_owner.p)) return;
	const size_t N = dm.N;
	const int nf = (int)dm.n_fast;
	if(!dm.icache.p) dm.icache.alloc(3 * N);
	if(want_tap && !dm.icache_owner.p) dm.icache_owner.alloc(3 * 2 * N);
	CK(cudaMemsetAsync(dm.icache.p, 0, 3 * N * sizeof(gcm::InnerCacheEntry), s->stream));
	CK(cudaMemsetAsync(dm.cold.p, 0, sizeof(int), s->stream));
	dm.n_cold = 0;
	if(nf) {
		gcm::MeshView m = make_view(s, tau);
		m.exact_only = 1;      // the reference's own predicate for every candidate: nothing to be conservative about
		const int g = (nf + 127) / 128;
		for(int st = 0; st < 3; st++)
			gcm::build_inner_cache_kernel<<<g, 128, 0, s->stream>>>(m, dm.fast_list.p, nf, st, dm.icache.p + st * N, dm.cold.p,
				dm.icache_owner.p ? dm.icache_owner.p + (size_t)st * 2 * N : nullptr);
		gcm::merge_cold_kernel<<<(nf + 255) / 256, 256, 0, s->stream>>>(dm.icache.p, dm.icache.p + N, dm.icache.p + 2 * N, dm.fast_list.p, nf, dm.cold.p);
		s->launches += 4;
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(s->h_err + 4, dm.cold.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		dm.n_cold = s->h_err[4];
		for(int st = 0; st < 3; st++) { dm.pattern_ok[st] = false; dm.n_patterns[st] = 0; }
		if(s->inner_patterns) {
			// de-duplicate each axis: hash insert -> exact comparison with the representative + numbering -> ids
			const unsigned cap = 1u << 18;
			DevBuf<unsigned long long> keys; DevBuf<int> rep, slot_of, slot_pid, stats;
			keys.alloc(cap); rep.alloc(cap); slot_pid.alloc(cap); slot_of.alloc(N); stats.alloc(4);
			if(!dm.ic_pat_idx.p) { dm.ic_pat_idx.alloc(3 * N); dm.ic_patterns.alloc(3 * (size_t)GCM_IC_MAX_PATTERNS); }
			CK(cudaMemsetAsync(dm.ic_pat_idx.p, 0xff, 3 * N * sizeof(unsigned), s->stream));
			const int g256 = (nf + 255) / 256;
			for(int st = 0; st < 3; st++) {
				CK(cudaMemsetAsync(keys.p, 0, cap * sizeof(unsigned long long), s->stream));
				CK(cudaMemsetAsync(rep.p, 0x7f, cap * sizeof(int), s->stream));
				CK(cudaMemsetAsync(stats.p, 0, 4 * sizeof(int), s->stream));
				const gcm::InnerCacheEntry* ce = dm.icache.p + (size_t)st * N;
				gcm::ic_dedup_insert_kernel<<<g256, 256, 0, s->stream>>>(ce, dm.fast_list.p, nf, keys.p, rep.p, cap, slot_of.p, stats.p);
				gcm::ic_dedup_number_kernel<<<g256, 256, 0, s->stream>>>(ce, dm.fast_list.p, nf, rep.p, slot_of.p, slot_pid.p,
					dm.ic_patterns.p + (size_t)st * GCM_IC_MAX_PATTERNS, GCM_IC_MAX_PATTERNS, stats.p);
				gcm::ic_dedup_assign_kernel<<<g256, 256, 0, s->stream>>>(dm.fast_list.p, nf, slot_of.p, slot_pid.p, dm.ic_pat_idx.p + (size_t)st * N);
				s->launches += 3;
				CK(cudaGetLastError());
				CK(cudaMemcpyAsync(s->h_err + 4, stats.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
				CK(cudaStreamSynchronize(s->stream));
				dm.pattern_ok[st] = s->h_err[5] == 0 && s->h_err[6] <= GCM_IC_MAX_PATTERNS;
				dm.n_patterns[st] = s->h_err[6];
			}
		}
	}
	// shell nodes: flagged and listed when every axis runs from its pattern table and there are other processes
	dm.n_shell = 0;
	if(nf && !s->peers.empty() && dm.pattern_ok[0] && dm.pattern_ok[1] && dm.pattern_ok[2]) {
		std::vector<unsigned char> mask(N, 0);
		for(auto& kv : s->peers)
			for(size_t k = 0; k < kv.second.recv_node.size(); k++)
				mask[s->zones[kv.second.recv_zone[k]]->node_off + kv.second.recv_node[k]] = 1;
		DevBuf<unsigned char> d_mask; d_mask.upload(mask, s->stream);
		dm.shell.alloc((size_t)nf + 1);
		CK(cudaMemsetAsync(dm.shell.p, 0, sizeof(int), s->stream));
		gcm::ic_mark_shell_kernel<<<(nf + 255) / 256, 256, 0, s->stream>>>(dm.fast_list.p, nf, dm.ic_pat_idx.p, dm.ic_patterns.p, (int)N,
			GCM_IC_MAX_PATTERNS, d_mask.p, dm.shell.p);
		s->launches++;
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(s->h_err + 4, dm.shell.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		dm.n_shell = s->h_err[4];
	}
	dm.icache_tau = tau;
	dm.icache_ready = true;
}

// One stage for the nodes of one zone (zone >= 0) or of every local zone (zone < 0).
// fuse: stage 2 only -- fold the plastic stage into the writers (whole-step calls).
void comm_exchange(gcm_b200_set* s, cudaStream_t st);

// exchange_here: the cross-process halo exchange of this stage has NOT been issued yet and is to run inside the
// stage, on the border stream, beside the inner nodes that do not depend on it (whole-step calls of multi-process sets).
void run_stage(gcm_b200_set* s, int zone, float tau, int stage, bool fuse = false, bool exchange_here = false)
{
	DeviceMesh& dm = s->dm;
	Zone* z = zone >= 0 ? s->zones[zone].get() : nullptr;
	if(stage == 3) {
		const size_t off = z ? z->node_off : 0;
		const int n = (int)(z ? (size_t)z->hm.N : dm.N);
		if(n) {
			ProfScope ps(s, PROF_PLASTIC, n);
			gcm::plastic_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(dm.ucur(), dm.rs, dm.mat.p, dm.flags.p, (int)off, n, dm.stride, s->d_err, zone);
			s->launches++;
		}
		return;
	}
	const bool cached = s->fast_inner && s->inner_cache;
	if(cached) ensure_inner_cache(s, tau);
	gcm::MeshView m = make_view(s, tau);
	const float* yp = (fuse && stage == 2) ? dm.mat.p + 3 * dm.stride : nullptr;
	if(s->tap_on && !dm.tap.p) { dm.tap.alloc(3 * dm.N); CK(cudaMemsetAsync(dm.tap.p, 0xff, 3 * dm.N * sizeof(gcm::StageTap), s->stream)); }
	gcm::StageTap* tap = s->tap_on ? dm.tap.p : nullptr;
	// border nodes: one thread per node
	int nb = (int)(z ? z->hm.border_launch.size() : dm.NB);
	const int* bl = dm.border_list.p + (z ? z->border_off : 0);
	const bool contact = stage == 0 && s->contact_active;
	const bool dev_lists = contact && s->coll_ready;      // contact state built on the device: full lists + filters
	const int* skip_contact = dev_lists ? s->d_axis_plus0.p : nullptr;
	if(contact && !dev_lists) {   // border nodes in contact on axis 0 take the contact kernel instead
		nb = (int)(z ? z->st0_n : s->st0_total);
		bl = s->d_border_stage0.p + (z ? z->st0_off : 0);
	}
	cudaStream_t bst = s->stream;
	const bool border_work = nb > 0;
	// Overlapped stage: push + wait run on the border stream while the main stream already advances every inner node
	// whose stencil stays clear of other processes' halo nodes; border nodes and the shell nodes follow the exchange.
	const bool overlap = exchange_here && s->overlap && s->concurrent && cached && !contact && !tap && zone < 0
		&& dm.pattern_ok[0] && dm.pattern_ok[1] && dm.pattern_ok[2] && s->fast_inner && dm.n_generic == 0;
	if(exchange_here && !overlap) comm_exchange(s, s->stream);
	if(overlap) {
		bst = s->aux;
		CK(cudaEventRecord(s->ev_fork, s->stream));
		CK(cudaStreamWaitEvent(s->aux, s->ev_fork, 0));
		comm_exchange(s, s->aux);
		CK(cudaEventRecord(s->ev_halo, s->aux));
	}
	if(border_work) {
		if(s->concurrent && !overlap) {
			bst = s->aux;
			CK(cudaEventRecord(s->ev_fork, s->stream));
			CK(cudaStreamWaitEvent(s->aux, s->ev_fork, 0));
		}
		ProfScope ps(s, PROF_BORDER, nb, bst);
		if(s->flat_border) {
			// the redesigned kernel (gcm_border.cuh), then the literal routine for what it deferred
			CK(cudaMemsetAsync(dm.bnd_defer.p, 0, sizeof(int), bst));
			const gcm::BndSlotStatic* bstat = reinterpret_cast<const gcm::BndSlotStatic*>(dm.bnd_static.p);
			const float4* cq = reinterpret_cast<const float4*>(dm.bnd_cand.p);
			// stage 0 = the normal axis: every flat node needs the 9x9 solve, 648 bytes of shared memory per thread.
			// One 352-thread block per SM on all of its shared memory keeps 352 systems resident against 5 x 64 = 320:
			// 52 096 nodes per wave instead of 47 360 -- one wave for a rank that owns a whole 216^2 face
			if(stage == 0 && s->border_lu_block == 352) gcm::stage_border_fast_kernel<true, 352><<<(nb + 351) / 352, 352, 81 * 352 * sizeof(double), bst>>>(m, bl, nb, stage, dm.unext(), s->d_err, zone, tap, yp, bstat, cq, dm.bnd_defer.p, skip_contact);
			else if(stage == 0) gcm::stage_border_fast_kernel<true, 64><<<(nb + 63) / 64, 64, 0, bst>>>(m, bl, nb, stage, dm.unext(), s->d_err, zone, tap, yp, bstat, cq, dm.bnd_defer.p, skip_contact);
			else gcm::stage_border_fast_kernel<false, 128><<<(nb + 127) / 128, 128, 0, bst>>>(m, bl, nb, stage, dm.unext(), s->d_err, zone, tap, yp, bstat, cq, dm.bnd_defer.p);
			gcm::stage_deferred_kernel<<<64, 128, 0, bst>>>(m, dm.bnd_defer.p, stage, dm.unext(), s->d_err, zone, tap, yp, 0, 0x7fffffff);
			s->launches += 2;
		} else {
			gcm::stage_generic_kernel<<<(nb + 127) / 128, 128, 0, bst>>>(m, bl, nb, stage, dm.unext(), s->d_err, zone, tap, yp, skip_contact);
			s->launches++;
		}
	}
	if(!s->fast_inner) {
		const int ni = (int)(z ? z->hm.inner_nodes.size() : dm.n_inner);
		if(ni) {
			ProfScope ps(s, PROF_INNER, ni);
			gcm::stage_generic_kernel<<<(ni + 127) / 128, 128, 0, s->stream>>>(m, dm.inner_list.p + (z ? z->inner_off : 0), ni, stage, dm.unext(), s->d_err, zone, tap, yp);
			s->launches++;
		}
	} else {
		const int ng = (int)(z ? z->hm.inner_generic.size() : dm.n_generic);
		const int nf = (int)(z ? z->hm.inner_fast.size() : dm.n_fast);
		if(ng) {
			ProfScope ps(s, PROF_INNER_GENERIC, ng);
			gcm::stage_generic_kernel<<<(ng + 127) / 128, 128, 0, s->stream>>>(m, dm.generic_list.p + (z ? z->generic_off : 0), ng, stage, dm.unext(), s->d_err, zone, tap, yp);
			s->launches++;
		}
		if(nf) {
			const unsigned char* mid = dm.n_mat > 1 ? dm.mat_id.p : nullptr;
			const gcm::MatTab* tab = reinterpret_cast<const gcm::MatTab*>(dm.mat_tab.p);
			if(cached) {
				// the contiguous node range of the zone (or of the whole device mesh); entries of other nodes are invalid
				const int nbeg = (int)(z ? z->node_off : 0), nrange = (int)(z ? (size_t)z->hm.N : dm.N);
				const dim3 grid((unsigned)((nrange + 255) / 256)), block(256);
				const gcm::InnerCacheEntry* ce = dm.icache.p + (size_t)stage * dm.N;
				const int* otap = (tap && dm.icache_owner.p) ? dm.icache_owner.p + (size_t)stage * 2 * dm.N : nullptr;
				{
					ProfScope ps(s, PROF_INNER, nf - dm.n_cold);
					const unsigned* pidx = dm.ic_pat_idx.p + (size_t)stage * dm.N;
					const gcm::InnerCacheEntry* pats = dm.ic_patterns.p + (size_t)stage * GCM_IC_MAX_PATTERNS;
					const bool pat = dm.pattern_ok[stage];
					const int* shell_list = nullptr; int n_shell = 0;      // first launch: the node range, shell nodes left out
#define GCM_LAUNCH_PATTERN(S_, P_, MB_) gcm::stage_inner_pattern_kernel<S_, P_, MB_><<<grid, block, 0, s->stream>>>(dm.ucur(), dm.rs, nbeg, nrange, dm.unext(), pidx, pats, tab, dm.n_mat, mid, yp, \
						s->d_err, zone, tap, otap, dm.n2t_off.p, dm.n2t.p, (int)dm.N, shell_list, n_shell)
#define GCM_LAUNCH_CACHED(S_, P_) do { if(pat) { if(s->inner_mb == 4) GCM_LAUNCH_PATTERN(S_, P_, 4); else if(s->inner_mb == 5) GCM_LAUNCH_PATTERN(S_, P_, 5); \
						else if(s->inner_mb == 6) GCM_LAUNCH_PATTERN(S_, P_, 6); else GCM_LAUNCH_PATTERN(S_, P_, 1); } \
					else gcm::stage_inner_cached_kernel<S_, P_><<<grid, block, 0, s->stream>>>(dm.ucur(), dm.rs, nbeg, nrange, dm.unext(), ce, tab, dm.n_mat, mid, yp, \
						s->d_err, zone, tap, otap, dm.n2t_off.p, dm.n2t.p, (int)dm.N); } while(0)
					if(stage == 0) GCM_LAUNCH_CACHED(0, false);
					else if(stage == 1) GCM_LAUNCH_CACHED(1, false);
					else if(yp) GCM_LAUNCH_CACHED(2, true);
					else GCM_LAUNCH_CACHED(2, false);
					s->launches++;
					if(dm.n_shell && pat) {
						// the shell nodes, once the halo exchange running beside the launch above has delivered (overlapped
						// stage), or right behind it (everything else)
						if(overlap) CK(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
						shell_list = dm.shell.p + 1; n_shell = dm.n_shell;
						const dim3 grid((unsigned)((n_shell + 255) / 256));
						if(stage == 0) GCM_LAUNCH_CACHED(0, false);
						else if(stage == 1) GCM_LAUNCH_CACHED(1, false);
						else if(yp) GCM_LAUNCH_CACHED(2, true);
						else GCM_LAUNCH_CACHED(2, false);
						s->launches++;
					} else if(overlap) CK(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));   // the cold list below may touch the halo
#undef GCM_LAUNCH_CACHED
#undef GCM_LAUNCH_PATTERN
				}
				if(dm.n_cold) {
					// whatever the cache could not take: the literal routine, right behind (a per-zone call strides
					// over the whole list; nodes of other zones are skipped by the range test below)
					ProfScope ps(s, PROF_INNER_GENERIC, dm.n_cold);
					gcm::stage_deferred_kernel<<<(dm.n_cold + 127) / 128 < 296 ? (dm.n_cold + 127) / 128 : 296, 128, 0, s->stream>>>(m, dm.cold.p, stage, dm.unext(), s->d_err, zone, tap, yp,
						nbeg, nbeg + nrange);
					s->launches++;
				}
			} else {
				ProfScope ps(s, PROF_INNER, nf);
				const int4* n2x = reinterpret_cast<const int4*>(dm.n2x.p);
				const int* fl = dm.fast_list.p + (z ? z->fast_off : 0);
				const dim3 grid((nf + 127) / 128), block(128);
#define GCM_LAUNCH_INNER(S_, MB_) gcm::stage_inner_kernel<S_, MB_><<<grid, block, 0, s->stream>>>(m, fl, nf, dm.unext(), s->d_err, zone, tap, tab, dm.n_mat, mid, dm.w0.p, n2x, yp)
				if(stage == 0) GCM_LAUNCH_INNER(0, 5);
				else if(stage == 1) GCM_LAUNCH_INNER(1, 5);
				else GCM_LAUNCH_INNER(2, 5);
#undef GCM_LAUNCH_INNER
				s->launches++;
			}
		}
	}
	if((border_work && s->concurrent) || overlap) {
		CK(cudaEventRecord(s->ev_join, s->aux));
		CK(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
	}
	if(contact) {
		// phase 0: partner body comes later in local_meshes -> its current layer is what the reference
		// reads; phase 1: partner already advanced -> read the next layer, zone after zone
		// (TetrMeshSet.cpp:158-167 advances the meshes one after the other)
		const int first = s->step_count == 0 ? 1 : 0;
		auto launch = [&](size_t off, size_t n, int partner_next, int zn) {
			if(!n) return;
			ProfScope ps(s, PROF_BORDER, (long long)n);
			// pass 1: two list-free rings; nodes that need more are collected in d_deep and advanced
			// by pass 2, whose few threads own a find_owner_general workspace each
			const int deep_blocks = 32;
			if(s->d_deep.n < n + 1) s->d_deep.alloc(n + 1);
			if(!s->d_ring_ws.p) s->d_ring_ws.alloc((size_t)deep_blocks * 64 * GCM_RING_WS);
			CK(cudaMemsetAsync(s->d_deep.p, 0, sizeof(int), s->stream));
			const int* list = (dev_lists ? s->coll_cand_node.p : s->d_contact_nodes.p) + off;
			const int* sel = dev_lists ? s->coll_hit_virt.p + off : nullptr;
			const int* sel_phase = dev_lists ? s->coll_cand_phase.p + off : nullptr;
			gcm::stage_contact_kernel<<<(int)((n + 63) / 64), 64, 0, s->stream>>>(m, list, (int)n, stage, dm.unext(),
				s->d_virt.p, s->d_axis_plus0.p, dm.normals.p, s->contact_calc, partner_next, first, s->d_err, zn, tap, s->d_deep.p, 0, sel, sel_phase, partner_next);
			gcm::MeshView md = m;
			md.ring_ws = s->d_ring_ws.p;
			gcm::stage_contact_kernel<<<deep_blocks, 64, 0, s->stream>>>(md, list, (int)n, stage, dm.unext(),
				s->d_virt.p, s->d_axis_plus0.p, dm.normals.p, s->contact_calc, partner_next, first, s->d_err, zn, tap, s->d_deep.p, 1);
			s->launches += 2;
		};
		// a partner that was already advanced is read in its NEXT layer; its Remote (halo) entries are not written by
		// a stage and keep the values synced before it (the copy-back loop, TetrMesh_1stOrder.cpp:1925-1932, touches
		// local nodes only), so those slots are carried over from the current layer first
		if(dm.n_halo && (dev_lists ? (size_t)s->coll_n_cand : (z ? z->c1_n : s->virt.size()))) {
			const int n = (int)dm.n_halo;
			gcm::halo_copy_kernel<<<(3 * n + 255) / 256, 256, 0, s->stream>>>(dm.unext(), dm.halo_dst.p, dm.ucur(), dm.halo_dst.p, n, dm.rs);
			s->launches++;
		}
		if(dev_lists) {
			// every candidate of the pass is listed (zone after zone); the kernel keeps those that hit, by phase
			auto zpos = [&](int zn) { for(size_t a = 0; a < s->local_zones.size(); a++) if(s->local_zones[a] == zn) return a; return (size_t)0; };
			if(z) { const size_t a = zpos(zone); launch(s->coll_zone_off[a], s->coll_zone_n[a], 0, zone); launch(s->coll_zone_off[a], s->coll_zone_n[a], 1, zone); }
			else {
				launch(0, s->coll_n_cand, 0, -1);
				for(size_t a = 0; a < s->local_zones.size(); a++) launch(s->coll_zone_off[a], s->coll_zone_n[a], 1, s->local_zones[a]);
			}
		} else if(z) { launch(z->c0_off, z->c0_n, 0, zone); launch(z->c1_off, z->c1_n, 1, zone); }
		else {
			launch(0, s->c0_total, 0, -1);
			for(int zi : s->local_zones) launch(s->zones[zi]->c1_off, s->zones[zi]->c1_n, 1, zi);
		}
	}
	CK(cudaGetLastError());
}

// CollisionDetector::find_collisions + the bookkeeping around it in TetrMeshSet::do_next_step
// (TetrMeshSet.cpp:111-120): runs on the host (gcmh::find_collisions) on the border nodes' current
// values and bases, then lays the result out for stage_contact_kernel.
// The device detector: static plan once, then four small launches per pass on the set's stream -- no host round trip.
void run_collision_detection_device(gcm_b200_set* s)
{
	DeviceMesh& dm = s->dm;
	cudaStream_t st = s->stream;
	if(!s->coll_ready) {
		std::vector<HostMesh*> meshes;
		for(int zi : s->local_zones) meshes.push_back(&s->zones[zi]->hm);
		gcmh::CollisionPlan plan;
		gcmh::build_collision_plan(meshes, s->detector_threshold, plan);
		// zone-local numbers -> the merged device mesh
		std::vector<int> phase(plan.cand_node.size());
		for(size_t g = 0; g < plan.cand_node.size(); g++) {
			Zone& za = *s->zones[s->local_zones[plan.cand_mesh[g]]];
			plan.cand_node[g] += (int)za.node_off;
			plan.cand_slot[g] += (int)za.slot_off;
			phase[g] = plan.pairs[plan.cand_pair[g]].phase;
		}
		for(size_t f = 0; f < plan.faces.size(); f++) {
			const int off = (int)s->zones[s->local_zones[plan.face_mesh[f]]]->node_off;
			plan.faces[f].v0 += off; plan.faces[f].v1 += off; plan.faces[f].v2 += off;
		}
		s->coll_n_cand = (int)plan.cand_node.size();
		s->coll_zone_off.assign(plan.mesh_cand_off.begin(), plan.mesh_cand_off.end());
		s->coll_zone_n.assign(plan.mesh_cand_n.begin(), plan.mesh_cand_n.end());
		s->coll_pairs.upload(plan.pairs, st); s->coll_faces.upload(plan.faces, st);
		s->coll_cand_pair.upload(plan.cand_pair, st); s->coll_cand_node.upload(plan.cand_node, st); s->coll_cand_slot.upload(plan.cand_slot, st);
		s->coll_cand_phase.upload(phase, st);
		if(plan.grid_start.empty()) plan.grid_start.push_back(0);
		if(plan.grid_items.empty()) plan.grid_items.push_back(0);
		s->coll_grid_start.upload(plan.grid_start, st); s->coll_grid_items.upload(plan.grid_items, st);
		const size_t nc = (size_t)std::max(1, s->coll_n_cand);
		s->coll_hit_l.alloc(nc); s->coll_hit_virt.alloc(nc); s->coll_hit_p.alloc(3 * nc); s->coll_n_virt.alloc(1);
		s->d_virt.alloc(nc);
		CK(cudaMemsetAsync(s->coll_n_virt.p, 0, sizeof(int), st));
		if(!s->d_border_by_slot.p && dm.NB) {
			std::vector<int> bys;
			for(int zi : s->local_zones) {
				Zone& z = *s->zones[zi];
				for(int b : z.hm.border_nodes) bys.push_back((int)(z.node_off + b));
			}
			s->d_border_by_slot.upload(bys, st);
			s->d_gather.alloc(9 * dm.NB);
		}
		s->d_axis_plus0.alloc(std::max<size_t>(1, dm.NB));
		s->coll_ready = true;
	}
	const int nc = s->coll_n_cand;
	if(dm.NB) {
		gcm::collide_clear_kernel<<<(int)((dm.NB + 255) / 256), 256, 0, st>>>(dm.flags.p, s->d_border_by_slot.p, s->d_axis_plus0.p, (int)dm.NB);
		s->launches++;
	}
	if(nc) {
		ProfScope ps(s, PROF_BORDER, nc);
		gcm::collide_find_kernel<<<(nc + 127) / 128, 128, 0, st>>>(dm.ucur(), dm.rs, dm.basis.p, s->coll_pairs.p, s->coll_cand_pair.p, s->coll_cand_node.p,
			s->coll_cand_slot.p, nc, s->coll_faces.p, s->coll_grid_start.p, s->coll_grid_items.p, s->coll_hit_l.p, s->coll_hit_p.p, s->d_err);
		gcm::collide_number_kernel<<<1, 1024, 0, st>>>(s->coll_hit_l.p, nc, s->coll_hit_virt.p, s->coll_n_virt.p);
		gcm::collide_build_kernel<<<(nc + 127) / 128, 128, 0, st>>>(dm.ucur(), dm.rs, dm.mat.p, dm.stride, s->coll_pairs.p, s->coll_cand_pair.p, s->coll_cand_node.p,
			s->coll_cand_slot.p, nc, s->coll_faces.p, s->coll_hit_l.p, s->coll_hit_p.p, s->coll_hit_virt.p, s->d_virt.p, s->d_axis_plus0.p, dm.flags.p, s->d_err);
		s->launches += 3;
	}
	CK(cudaGetLastError());
	s->contact_active = nc > 0;
	s->coll_host_stale = true;
}

// Host mirrors of a device pass (virtual nodes, contact flags), fetched when somebody asks for them.
void collisions_to_host(gcm_b200_set* s)
{
	if(!s->coll_ready || !s->coll_host_stale) return;
	DeviceMesh& dm = s->dm;
	CK(cudaSetDevice(s->device));
	int n = 0;
	CK(cudaMemcpyAsync(&n, s->coll_n_virt.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	std::vector<gcm::VirtNode> dv((size_t)n);
	if(n) CK(cudaMemcpyAsync(dv.data(), s->d_virt.p, (size_t)n * sizeof(gcm::VirtNode), cudaMemcpyDeviceToHost, s->stream));
	std::vector<int> fl(dm.N);
	if(dm.N) CK(cudaMemcpyAsync(fl.data(), dm.flags.p, dm.N * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	s->virt.assign((size_t)n, gcmh::VirtNodeHost());
	for(int k = 0; k < n; k++) {
		gcmh::VirtNodeHost& h = s->virt[k];
		memset(&h, 0, sizeof h);
		for(int q = 0; q < 3; q++) h.coords[q] = dv[k].pos[q];
		for(int q = 0; q < 9; q++) h.values[q] = dv[k].val[q];
		h.values[9] = dv[k].la; h.values[10] = dv[k].mu; h.values[11] = dv[k].rho; h.values[12] = dv[k].yield;
	}
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		for(int i = 0; i < z.hm.N; i++) z.hm.flags[i] = (z.hm.flags[i] & ~1) | (fl[z.node_off + i] & 1);
	}
	s->coll_host_stale = false;
}

void run_collision_detection(gcm_b200_set* s)
{
	if(s->coll_device) { run_collision_detection_device(s); return; }
	DeviceMesh& dm = s->dm;
	std::vector<HostMesh*> meshes;
	for(int zi : s->local_zones) meshes.push_back(&s->zones[zi]->hm);
	if(dm.NB) {
		// current values of the border nodes and (device mode) this step's bases -> host mirrors
		if(!s->d_border_by_slot.p) {
			std::vector<int> bys;
			for(int zi : s->local_zones) {
				Zone& z = *s->zones[zi];
				for(int b : z.hm.border_nodes) bys.push_back((int)(z.node_off + b));
			}
			s->d_border_by_slot.upload(bys, s->stream);
			s->d_gather.alloc(9 * dm.NB);
		}
		gcm::gather_records_kernel<<<(int)((dm.NB + 255) / 256), 256, 0, s->stream>>>(s->d_gather.p, dm.ucur(), dm.rs, s->d_border_by_slot.p, (int)dm.NB);
		s->launches++;
		std::vector<float> g(9 * dm.NB), bs(9 * dm.NB);
		CK(cudaMemcpyAsync(g.data(), s->d_gather.p, g.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
		if(s->basis_on_device) CK(cudaMemcpyAsync(bs.data(), dm.basis.p, bs.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		for(int zi : s->local_zones) {
			Zone& z = *s->zones[zi];
			HostMesh& hm = z.hm;
			for(size_t b = 0; b < hm.border_nodes.size(); b++)
				memcpy(&hm.values[9*(size_t)hm.border_nodes[b]], &g[9 * (z.slot_off + b)], 9 * sizeof(float));
			if(s->basis_on_device && !hm.basis.empty()) memcpy(hm.basis.data(), &bs[9 * z.slot_off], hm.basis.size() * sizeof(float));
		}
	}
	gcmh::find_collisions(meshes, s->detector_threshold, s->virt, s->axis_plus0);

	std::map<int, int> order;   // zone number -> position in local_meshes
	for(size_t a = 0; a < s->local_zones.size(); a++) order[s->local_zones[a]] = (int)a;
	std::vector<gcm::VirtNode> dv(s->virt.size());
	for(size_t k = 0; k < s->virt.size(); k++) {
		const gcmh::VirtNodeHost& h = s->virt[k];
		gcm::VirtNode& v = dv[k];
		for(int q = 0; q < 3; q++) v.pos[q] = h.coords[q];
		for(int q = 0; q < 9; q++) v.val[q] = h.values[q];
		v.la = h.values[9]; v.mu = h.values[10]; v.rho = h.values[11]; v.yield = h.values[12];
		v.base = (int)(s->zones[h.remote_zone]->node_off + h.base_node);
		v.real = (int)(s->zones[h.real_zone]->node_off + h.real_node);
		v.phase = order[h.remote_zone] < order[h.real_zone] ? 1 : 0;
	}
	std::vector<int> axis(dm.NB, -1), st0, c0, c1;
	for(size_t a = 0; a < s->local_zones.size(); a++) {
		Zone& z = *s->zones[s->local_zones[a]];
		for(size_t b = 0; b < z.hm.border_nodes.size(); b++) axis[z.slot_off + b] = s->axis_plus0[a][b];
	}
	// lists: stage-0 border nodes not in contact (sharp / flat, launch order), contact nodes by phase
	for(size_t a = 0; a < s->local_zones.size(); a++) {
		Zone& z = *s->zones[s->local_zones[a]];
		z.st0_off = st0.size(); z.c0_off = c0.size();
		for(size_t q = 0; q < z.hm.border_launch.size(); q++) {
			const int node = z.hm.border_launch[q];
			const int v = s->axis_plus0[a][z.hm.border_slot[node]];
			if(v < 0) st0.push_back((int)(z.node_off + node));
			else if(dv[v].phase == 0) c0.push_back((int)(z.node_off + node));
		}
		z.st0_n = st0.size() - z.st0_off; z.c0_n = c0.size() - z.c0_off;
	}
	for(size_t a = 0; a < s->local_zones.size(); a++) {
		Zone& z = *s->zones[s->local_zones[a]];
		z.c1_off = c0.size() + c1.size();
		for(int node : z.hm.border_launch) {
			const int v = s->axis_plus0[a][z.hm.border_slot[node]];
			if(v >= 0 && dv[v].phase == 1) c1.push_back((int)(z.node_off + node));
		}
		z.c1_n = c0.size() + c1.size() - z.c1_off;
	}
	s->st0_total = st0.size(); s->c0_total = c0.size();
	std::vector<int> cn(c0); cn.insert(cn.end(), c1.begin(), c1.end());
	s->d_virt.upload(dv, s->stream);
	s->d_axis_plus0.upload(axis, s->stream);
	s->d_border_stage0.upload(st0, s->stream);
	s->d_contact_nodes.upload(cn, s->stream);
	s->contact_active = !s->virt.empty();
}

// copy-back loop of TetrMesh_1stOrder.cpp:1925-1932 == buffer swap, once every local zone has
// run the stage; the halo slots of the new current layer are refreshed by the next sync_nodes
void finish_axis_stage(gcm_b200_set* s)
{
	s->dm.cur ^= 1;
	for(int zi : s->local_zones) s->zones[zi]->stage_done = false;
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
	case gcm::ERR_RING_OVERFLOW: return "Owner search needs more than the first and second ring of tetrahedrons (only contact nodes search further on the device; a ring is capped at 2048 tetrahedrons)";
	case gcm::ERR_BAD_RHEOLOGY: return "Bad rheology";
	case gcm::ERR_VIRT_OUTER_COUNT: return "Illegal number of outer characteristics";
	case gcm::ERR_BAD_Q: return "Bad vector q";
	case gcm::ERR_ZERO_VOLUME: return "Tetr volume is zero";
	case gcm::ERR_FOOT_PATTERN: return "More than 5 distinct characteristic feet";
	case gcm::ERR_PEER_TIMEOUT: return "A neighbour process did not deliver its halo nodes";
	case gcm::ERR_COLLISION_NAN: return "NAN or INF in collision detection (stage field: 1 p1, 2 p2, 3 p3, 4 p0, 5 v, 6 n, 7 virt node coords)";
	default: return "Unknown device error";
	}
}

int step_error_exc(int code)
{
	switch(code) {
	case gcm::ERR_INTERP_SUM: case gcm::ERR_ZERO_VOLUME: case gcm::ERR_COLLISION_NAN: return gcmh::MESH_EXCEPTION;
	case gcm::ERR_BAD_RHEOLOGY: case gcm::ERR_BAD_Q: return gcmh::CONFIG_EXCEPTION;
	case gcm::ERR_INNER_NO_OWNER: case gcm::ERR_RING_OVERFLOW: return gcmh::UNIMPLEMENTED_EXCEPTION;
	case gcm::ERR_PEER_TIMEOUT: return gcmh::SYNC_EXCEPTION;
	default: return gcmh::METHOD_EXCEPTION;
	}
}

void peer_dev_prepare(gcm_b200_set* s, PeerPlan& p)
{
	if(p.dev_ready) return;
	std::vector<int> r(p.recv_node.size()), q(p.send_node.size());
	for(size_t k = 0; k < r.size(); k++) r[k] = (int)(s->zones[p.recv_zone[k]]->node_off + p.recv_node[k]);
	for(size_t k = 0; k < q.size(); k++) q[k] = (int)(s->zones[p.send_zone[k]]->node_off + p.send_node[k]);
	p.d_recv.upload(r, s->stream);
	p.d_send.upload(q, s->stream);
	p.dev_ready = true;
}

NcclApi& nccl_api()
{
	static NcclApi api;
	if(api.lib) return api;
	// the copy already in the process (torch's) if there is one, else the system library
	void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
	if(!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if(!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if(!h) throw HostError{gcmh::CONFIG_EXCEPTION, std::string("NCCL is not loadable: ") + (dlerror() ? dlerror() : "?")};
#define GCM_NCCL_SYM(field, name) do { *(void**)(&api.field) = dlsym(h, name); \
	if(!api.field) throw HostError{gcmh::CONFIG_EXCEPTION, std::string("NCCL symbol missing: ") + name}; } while(0)
	GCM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
	GCM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
	GCM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
	GCM_NCCL_SYM(Send, "ncclSend");
	GCM_NCCL_SYM(Recv, "ncclRecv");
	GCM_NCCL_SYM(GroupStart, "ncclGroupStart");
	GCM_NCCL_SYM(GroupEnd, "ncclGroupEnd");
	GCM_NCCL_SYM(AllGather, "ncclAllGather");
	GCM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef GCM_NCCL_SYM
	api.lib = h;
	return api;
}

void nccl_check(int rc, const char* what)
{
	if(rc != 0) throw HostError{gcmh::CONFIG_EXCEPTION, std::string("NCCL ") + what + ": " + nccl_api().GetErrorString(rc)};
}

// Base address of the allocation p lies in (cuMemGetAddressRange through the runtime's driver entry points: the
// library does not link libcuda).  An IPC handle always names a whole allocation, and cudaIpcOpenMemHandle returns
// its base: a pointer the driver carved out of a larger block (small cudaMalloc's share 2 MiB blocks) must travel
// as handle-of-base + offset.
bool allocation_base(const void* p, char** base)
{
	typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
	static range_fn fn = nullptr;
	if(!fn) {
		void* f = nullptr;
		cudaDriverEntryPointQueryResult q;
		if(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { cudaGetLastError(); return false; }
		fn = (range_fn)f;
	}
	unsigned long long b = 0; size_t sz = 0;
	if(fn(&b, &sz, (unsigned long long)(uintptr_t)p) != 0) return false;
	*base = (char*)(uintptr_t)b;
	return true;
}

// Maps every neighbour's record layers and flag array into this process and learns where its send
// nodes live in the neighbour's arrays.  Collective over the communicator.  Leaves p2p_ready false
// (the NCCL exchange stays in charge) when anything about it is not available.
void p2p_setup(gcm_b200_set* s)
{
	if(!s->p2p || s->peers.empty() || s->peers.size() > GCM_MAX_PEERS) return;
	NcclApi& nc = nccl_api();
	DeviceMesh& dm = s->dm;
	cudaStream_t st = s->stream;
	const int R = s->comm_size, me = s->comm_rank;
	const bool dbg = getenv("GCM_B200_P2P_DEBUG") != nullptr;
	// flags[r] = sequence number raised by rank r; flags[R + r] = scratch of the mapping check below.
	// Its own 2 MiB allocation: nothing else of this process becomes visible to the neighbours through the handle.
	s->p2p_flags.alloc((size_t)(2u << 20) / sizeof(int));
	CK(cudaMemsetAsync(s->p2p_flags.p, 0, 2 * (size_t)R * sizeof(int), st));
	// 1. everybody's handles: u0, u1, flags (handle of the allocation + offset inside it) + rs + ok marker
	struct Blob { cudaIpcMemHandle_t h[3]; unsigned long long off[3]; unsigned long long rs; unsigned long long ok; };
	Blob mine; memset(&mine, 0, sizeof mine);
	bool ok = true;
	{
		const void* ptr[3] = {dm.u0.p, dm.u1.p, s->p2p_flags.p};
		for(int k = 0; k < 3 && ok; k++) {
			char* base = nullptr;
			ok = allocation_base(ptr[k], &base) && cudaIpcGetMemHandle(&mine.h[k], base) == cudaSuccess;
			if(ok) mine.off[k] = (unsigned long long)((const char*)ptr[k] - base);
		}
		cudaGetLastError();
	}
	mine.rs = dm.rs; mine.ok = ok ? 1 : 0;
	DevBuf<char> d_mine, d_all;
	d_mine.alloc(sizeof(Blob)); d_all.alloc(sizeof(Blob) * (size_t)R);
	CK(cudaMemcpyAsync(d_mine.p, &mine, sizeof(Blob), cudaMemcpyHostToDevice, st));
	nccl_check(nc.AllGather(d_mine.p, d_all.p, sizeof(Blob), 0 /* ncclInt8 */, s->nccl_comm, st), "all gather");
	std::vector<Blob> all((size_t)R);
	CK(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Blob) * (size_t)R, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	bool all_ok = true;
	for(int r = 0; r < R; r++) all_ok = all_ok && all[r].ok;
	// 2. my send nodes' slots in each neighbour's arrays = the neighbour's recv list, in the agreed order
	std::map<int, DevBuf<int> > remote;
	for(auto& kv : s->peers) { peer_dev_prepare(s, kv.second); if(!kv.second.send_node.empty()) remote[kv.first].alloc(kv.second.send_node.size()); }
	nccl_check(nc.GroupStart(), "group start");
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		if(!p.recv_node.empty()) nccl_check(nc.Send(p.d_recv.p, p.recv_node.size(), 2 /* ncclInt32 */, kv.first, s->nccl_comm, st), "send");
		if(!p.send_node.empty()) nccl_check(nc.Recv(remote[kv.first].p, p.send_node.size(), 2, kv.first, s->nccl_comm, st), "recv");
	}
	nccl_check(nc.GroupEnd(), "group end");
	CK(cudaStreamSynchronize(st));
	// every process must take the same decision at every point it could give up: agree on it
	auto agree = [&](bool mine_ok) {
		int flag = mine_ok ? 1 : 0;
		DevBuf<int> d1, dR; d1.alloc(1); dR.alloc((size_t)R);
		CK(cudaMemcpyAsync(d1.p, &flag, sizeof(int), cudaMemcpyHostToDevice, st));
		nccl_check(nc.AllGather(d1.p, dR.p, 1, 2, s->nccl_comm, st), "all gather");
		std::vector<int> f((size_t)R);
		CK(cudaMemcpyAsync(f.data(), dR.p, R * sizeof(int), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		bool r = true;
		for(int k = 0; k < R; k++) r = r && f[k];
		return r;
	};
	if(!all_ok) { if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] no IPC handles\n", me); return; }
	// 3. map the neighbours' buffers
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		const Blob& b = all[kv.first];
		if(cudaIpcOpenMemHandle(&p.map_u0, b.h[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
		   cudaIpcOpenMemHandle(&p.map_u1, b.h[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
		   cudaIpcOpenMemHandle(&p.map_flags, b.h[2], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = false; }
		p.peer_rs = b.rs;
		p.peer_u[0] = p.map_u0 ? (float*)((char*)p.map_u0 + b.off[0]) : nullptr;
		p.peer_u[1] = p.map_u1 ? (float*)((char*)p.map_u1 + b.off[1]) : nullptr;
		p.peer_flags = p.map_flags ? (int*)((char*)p.map_flags + b.off[2]) : nullptr;
		if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] peer %d: u0 %p+%llu u1 %p+%llu flags %p+%llu rs %llu send %zu recv %zu\n", me, kv.first,
			p.map_u0, b.off[0], p.map_u1, b.off[1], p.map_flags, b.off[2], b.rs, p.send_node.size(), p.recv_node.size());
	}
	if(!agree(all_ok)) { if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] mapping failed somewhere\n", me); return; }
	// 4. check the mapping end to end before trusting it with the state: a token into every neighbour's scratch
	// flag and into the padding record (index N = rs - 1, never a node) of both its layers, then read back what arrived here
	{
		const int token = 0x5eed0000 + me;
		for(auto& kv : s->peers) {
			PeerPlan& p = kv.second;
			CK(cudaMemcpyAsync(p.peer_flags + R + me, &token, sizeof(int), cudaMemcpyHostToDevice, st));
			const float ft = (float)(1000 + me);
			for(int l = 0; l < 2; l++)
				CK(cudaMemcpyAsync(p.peer_u[l] + 4 * ((size_t)((me % 12) / 4) * p.peer_rs + (p.peer_rs - 1)) + me % 4, &ft, sizeof(float), cudaMemcpyHostToDevice, st));
		}
		CK(cudaStreamSynchronize(st));
		agree(true);                       // barrier: every token is on its way before anybody looks
		bool good = true;
		std::vector<int> fl(2 * (size_t)R);
		CK(cudaMemcpyAsync(fl.data(), s->p2p_flags.p, 2 * (size_t)R * sizeof(int), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		for(auto& kv : s->peers) {
			good = good && fl[R + kv.first] == 0x5eed0000 + kv.first;
			// rank r writes float r % 12 of the padding record (two neighbours congruent mod 12: the check fails, NCCL stays)
			for(int l = 0; l < 2; l++) {
				float got = 0;
				const float* src = (l ? dm.u1.p : dm.u0.p) + 4 * ((size_t)((kv.first % 12) / 4) * dm.rs + (dm.rs - 1)) + kv.first % 4;
				CK(cudaMemcpy(&got, src, sizeof(float), cudaMemcpyDeviceToHost));
				good = good && got == (float)(1000 + kv.first);
				const float zero = 0;
				CK(cudaMemcpy((void*)src, &zero, sizeof(float), cudaMemcpyHostToDevice));
			}
		}
		if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] mapping check %s\n", me, good ? "passed" : "FAILED");
		if(!agree(good)) return;
	}
	// 5. concatenated send list + per-neighbour targets
	std::vector<int> send_all, peer_ranks;
	size_t n_total = 0;
	for(auto& kv : s->peers) n_total += kv.second.send_node.size();
	s->p2p_send_node.alloc(n_total); s->p2p_remote_slot.alloc(n_total);
	s->p2p_targets.clear();
	size_t off = 0;
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		gcm::PeerTarget t;
		t.u[0] = p.peer_u[0]; t.u[1] = p.peer_u[1]; t.rs = p.peer_rs; t.flag = p.peer_flags;
		t.begin = (int)off; t.end = (int)(off + p.send_node.size());
		if(!p.send_node.empty()) {
			CK(cudaMemcpyAsync(s->p2p_send_node.p + off, p.d_send.p, p.send_node.size() * sizeof(int), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(s->p2p_remote_slot.p + off, remote[kv.first].p, p.send_node.size() * sizeof(int), cudaMemcpyDeviceToDevice, st));
		}
		off += p.send_node.size();
		s->p2p_targets.push_back(t);
		peer_ranks.push_back(kv.first);
	}
	{
		// every slot must lie inside the neighbour's node range (a mismatch of the two lists would write anywhere)
		std::vector<int> rs_host(n_total), sn_host(n_total);
		if(n_total) {
			CK(cudaMemcpyAsync(rs_host.data(), s->p2p_remote_slot.p, n_total * sizeof(int), cudaMemcpyDeviceToHost, st));
			CK(cudaMemcpyAsync(sn_host.data(), s->p2p_send_node.p, n_total * sizeof(int), cudaMemcpyDeviceToHost, st));
		}
		CK(cudaStreamSynchronize(st));
		bool good = true;
		for(size_t k = 0; k < s->p2p_targets.size(); k++) {
			const gcm::PeerTarget& t = s->p2p_targets[k];
			int lo = 0x7fffffff, hi = -1;
			for(int i = t.begin; i < t.end; i++) { lo = std::min(lo, rs_host[i]); hi = std::max(hi, rs_host[i]); }
			if(t.end > t.begin && (lo < 0 || (unsigned long long)hi + 1 >= t.rs)) good = false;
			if(dbg && t.end > t.begin) fprintf(stderr, "[gcm_b200 p2p %d] target %zu (rank %d): entries [%d, %d) my nodes %d..%d -> its slots %d..%d (range %d..%d) of rs %llu\n",
				me, k, peer_ranks[k], t.begin, t.end, sn_host[t.begin], sn_host[t.end - 1], rs_host[t.begin], rs_host[t.end - 1], lo, hi, t.rs);
		}
		if(!agree(good)) { if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] slot lists out of range somewhere\n", me); return; }
		// halo layers that are contiguous node ranges on both sides (z-slab zones) can go through the copy engines
		s->p2p_dma_ranges.clear();
		bool contiguous = s->p2p_dma_wanted;
		for(size_t k = 0; k < s->p2p_targets.size() && contiguous; k++) {
			const gcm::PeerTarget& t = s->p2p_targets[k];
			for(int i = t.begin; i < t.end && contiguous; i++)
				contiguous = sn_host[i] == sn_host[t.begin] + (i - t.begin) && rs_host[i] == rs_host[t.begin] + (i - t.begin);
			if(t.end > t.begin) { s->p2p_dma_ranges.push_back(sn_host[t.begin]); s->p2p_dma_ranges.push_back(rs_host[t.begin]); }
			else { s->p2p_dma_ranges.push_back(0); s->p2p_dma_ranges.push_back(0); }
		}
		s->p2p_dma = contiguous;
		if(dbg) fprintf(stderr, "[gcm_b200 p2p %d] copy-engine exchange: %s\n", me, s->p2p_dma ? "yes" : "no");
	}
	s->p2p_peer_ranks.upload(peer_ranks, st);
	s->p2p_more.upload(s->p2p_targets, st);
	s->p2p_counter.alloc(1);
	CK(cudaMemsetAsync(s->p2p_counter.p, 0, sizeof(unsigned), st));
	CK(cudaStreamSynchronize(st));
	s->p2p_seq = 0;
	s->p2p_ready = true;
}

// The exchange before a stage by peer stores: push my nodes into the neighbours' current layer, raise my
// flag there, wait for theirs.  A neighbour's current layer is safe to write: nothing of it but local
// nodes is written by the neighbour's own kernels, and its halo slots were last read two stages ago
// (the neighbour cannot be further behind than that: I needed its previous push to get here).
void p2p_exchange(gcm_b200_set* s, cudaStream_t st)
{

	DeviceMesh& dm = s->dm;
	const int np = (int)s->p2p_targets.size();
	const int n_total = (int)s->p2p_send_node.n;
	const int seq = ++s->p2p_seq;
	if(s->p2p_dma) {
		// three plane-wise peer copies per neighbour on the copy engines, then one small kernel: flags up, wait
		ProfScope ps(s, PROF_HALO, n_total, st);
		for(int k = 0; k < np; k++) {
			const gcm::PeerTarget& t = s->p2p_targets[k];
			const size_t cnt = (size_t)(t.end - t.begin);
			if(!cnt) continue;
			const size_t node0 = (size_t)s->p2p_dma_ranges[2 * k], slot0 = (size_t)s->p2p_dma_ranges[2 * k + 1];
			for(int q = 0; q < 3; q++)
				CK(cudaMemcpyAsync(t.u[dm.cur] + 4 * ((size_t)q * t.rs + slot0), dm.ucur() + 4 * ((size_t)q * dm.rs + node0),
				                   cnt * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
		}
		gcm::halo_flag_wait_kernel<<<1, 32, 0, st>>>(s->p2p_more.p, s->p2p_flags.p, s->p2p_peer_ranks.p, np, s->comm_rank, seq, s->d_err);
		s->launches++;
		CK(cudaGetLastError());
		return;
	}
	if(n_total) {
		ProfScope ps(s, PROF_HALO, n_total, st);
		const gcm::PeerTarget& t0 = s->p2p_targets[0];
		const gcm::PeerTarget& t1 = s->p2p_targets[np > 1 ? 1 : 0];
		gcm::halo_push_kernel<<<(3 * n_total + 255) / 256, 256, 0, st>>>(dm.ucur(), dm.rs, s->p2p_send_node.p, s->p2p_remote_slot.p, n_total,
			t0, t1, s->p2p_more.p, np, dm.cur, s->comm_rank, seq, s->p2p_counter.p);
		s->launches++;
	}
	gcm::halo_wait_kernel<<<1, 32, 0, st>>>(s->p2p_flags.p, s->p2p_peer_ranks.p, np, seq, s->d_err);
	s->launches++;
	CK(cudaGetLastError());

}

// DataBus::sync_nodes across processes (gcm/system/DataBus.cpp:81-115), all on the set's stream:
// gather the nodes every neighbour asked for, one grouped ncclSend/ncclRecv per neighbour over
// NVLink, scatter what arrived into the halo slots.
void comm_exchange(gcm_b200_set* s, cudaStream_t st)
{
	if(s->peers.empty()) return;
	if(s->p2p_ready) { p2p_exchange(s, st); return; }
	NcclApi& nc = nccl_api();
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		peer_dev_prepare(s, p);
		if(!p.recv_node.empty() && !p.recv_buf.p) p.recv_buf.alloc(9 * p.recv_node.size());
		const int n = (int)p.send_node.size();
		if(!n) continue;
		if(!p.send_buf.p) p.send_buf.alloc(9 * (size_t)n);
		ProfScope ps(s, PROF_HALO, n, st);
		gcm::halo_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(p.send_buf.p, s->dm.ucur(), s->dm.rs, p.d_send.p, n, n, 0);
		s->launches++;
	}
	CK(cudaGetLastError());
	nccl_check(nc.GroupStart(), "group start");
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		const size_t ns = p.send_node.size(), nr = p.recv_node.size();
		if(ns) nccl_check(nc.Send(p.send_buf.p, 9 * ns, 7 /* ncclFloat32 */, kv.first, s->nccl_comm, st), "send");
		if(nr) nccl_check(nc.Recv(p.recv_buf.p, 9 * nr, 7, kv.first, s->nccl_comm, st), "recv");
	}
	nccl_check(nc.GroupEnd(), "group end");
	for(auto& kv : s->peers) {
		PeerPlan& p = kv.second;
		const int n = (int)p.recv_node.size();
		if(!n) continue;
		ProfScope ps(s, PROF_HALO, n, st);
		gcm::halo_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(s->dm.ucur(), s->dm.rs, p.recv_buf.p, p.d_recv.p, n, n, 0);
		s->launches++;
	}
	CK(cudaGetLastError());
}


} // namespace

#define GUARD_BEGIN try {
#define GUARD_END } catch(HostError& e) { return fail(e.code, e.msg); } \
	catch(CudaError& e) { return fail(gcmh::CONFIG_EXCEPTION, "CUDA: " + e.msg); } \
	catch(std::exception& e) { return fail(gcmh::CONFIG_EXCEPTION, e.what()); }

#define NEED_SET(s) do { if(!(s)) return fail(gcmh::CONFIG_EXCEPTION, "null set"); } while(0)
#define NEED_DEVICE(s) do { if(!(s) || !(s)->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed"); \
	if((s)->local_zones.empty()) return fail(gcmh::CONFIG_EXCEPTION, "this process owns no zone (more processes than zones?): nothing to step"); \
	if((s)->device < 0) return fail(gcmh::CONFIG_EXCEPTION, "host-only set (device -1): there is no CPU execution path for the step"); \
	CK(cudaSetDevice((s)->device)); } while(0)

extern "C" {

const char* gcm_b200_version(void) { return "gcm_b200 0.2 (sm_100a)"; }
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
		if(s->zone_rank[z] == my_rank) {
			s->zones[z].reset(new Zone());
			s->zones[z]->hm.registry = &s->materials;
			s->local_zones.push_back(z);
		}
	if(device >= 0) {
		// the border kernels are short and latency-bound: their streams outrank the inner kernel's, so their
		// blocks are placed as soon as they are launched instead of queueing behind a 40 000-block grid
		int prio_lo = 0, prio_hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		if(const char* e = getenv("GCM_B200_AUX_PRIO")) if(atoi(e) == 0) prio_hi = prio_lo;
		CK(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_lo));
		CK(cudaStreamCreateWithPriority(&s->aux, cudaStreamNonBlocking, prio_hi));
		CK(cudaStreamCreateWithPriority(&s->aux2, cudaStreamNonBlocking, prio_hi));
		CK(cudaEventCreateWithFlags(&s->ev_join2, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming));
		CK(cudaStreamCreateWithFlags(&s->rng_stream, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&s->io_in, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&s->io_out, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&s->ev_rng_done, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_step_mark, cudaEventDisableTiming));
		CK(cudaMalloc(&s->d_err, 4 * sizeof(int)));
		CK(cudaMemset(s->d_err, 0, 4 * sizeof(int)));
		CK(cudaMallocHost(&s->h_err, 8 * sizeof(int)));
		CK(cudaFuncSetAttribute(gcm::stage_border_fast_kernel<true, 352>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(81 * 352 * sizeof(double))));
	}
	s->rng.seed(1);   // libc default when srand() was never called
	if(const char* e = getenv("GCM_B200_CONCURRENT")) s->concurrent = atoi(e) != 0;   // profiling aid
	if(const char* e = getenv("GCM_B200_FAST_INNER")) s->fast_inner = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_FLAT_BORDER")) s->flat_border = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_BASES_AHEAD")) s->bases_ahead = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_INNER_CACHE")) s->inner_cache = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_FUSE_PLASTIC")) s->fuse_plastic = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_INNER_PATTERNS")) s->inner_patterns = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_P2P")) s->p2p = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_INNER_MB")) s->inner_mb = atoi(e);
	if(const char* e = getenv("GCM_B200_OVERLAP")) s->overlap = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_DEVICE_COLLISIONS")) s->coll_device = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_P2P_DMA")) s->p2p_dma_wanted = atoi(e) != 0;
	if(const char* e = getenv("GCM_B200_BORDER_LU_BLOCK")) s->border_lu_block = atoi(e);
	*out = s.release();
	return 0;
	GUARD_END
}

int gcm_b200_set_destroy(gcm_b200_set* s)
{
	if(!s) return 0;
	if(s->device >= 0) {
		cudaSetDevice(s->device);
		if(s->stream) cudaStreamSynchronize(s->stream);
		if(s->aux) cudaStreamSynchronize(s->aux);
		if(s->aux2) { cudaStreamSynchronize(s->aux2); cudaStreamDestroy(s->aux2); }
		if(s->ev_join2) cudaEventDestroy(s->ev_join2);
		if(s->rng_stream) cudaStreamSynchronize(s->rng_stream);
		if(s->io_in) { cudaStreamSynchronize(s->io_in); cudaStreamDestroy(s->io_in); }
		if(s->io_out) { cudaStreamSynchronize(s->io_out); cudaStreamDestroy(s->io_out); }
		for(auto& zp : s->zones) if(zp) {
			if(zp->ev_h2d_done) cudaEventDestroy(zp->ev_h2d_done);
			if(zp->ev_in_consumed) cudaEventDestroy(zp->ev_in_consumed);
			if(zp->ev_out_ready) cudaEventDestroy(zp->ev_out_ready);
			if(zp->ev_d2h_done) cudaEventDestroy(zp->ev_d2h_done);
		}
		for(auto& kv : s->peers) {
			PeerPlan& p = kv.second;
			if(p.map_u0) cudaIpcCloseMemHandle(p.map_u0);
			if(p.map_u1) cudaIpcCloseMemHandle(p.map_u1);
			if(p.map_flags) cudaIpcCloseMemHandle(p.map_flags);
			p.map_u0 = p.map_u1 = p.map_flags = nullptr;
		}
		if(s->nccl_comm) { nccl_api().CommDestroy(s->nccl_comm); s->nccl_comm = nullptr; }
		prof_drain(s);
		for(int k = 0; k < 2; k++) { if(s->h_basis[k]) cudaFreeHost(s->h_basis[k]); if(s->basis_ev[k]) cudaEventDestroy(s->basis_ev[k]); }
		if(s->ev_fork) cudaEventDestroy(s->ev_fork);
		if(s->ev_join) cudaEventDestroy(s->ev_join);
		if(s->ev_halo) cudaEventDestroy(s->ev_halo);
		if(s->ev_rng_done) cudaEventDestroy(s->ev_rng_done);
		if(s->ev_step_mark) cudaEventDestroy(s->ev_step_mark);
		if(s->rng_stream) cudaStreamDestroy(s->rng_stream);
		if(s->d_err) cudaFree(s->d_err);
		if(s->h_err) cudaFreeHost(s->h_err);
	}
	cudaStream_t a = s->stream, b = s->aux;
	delete s;   // frees the device buffers
	if(a) cudaStreamDestroy(a);
	if(b) cudaStreamDestroy(b);
	return 0;
}

int gcm_b200_set_seed(gcm_b200_set* s, unsigned seed)
{
	NEED_SET(s);
	s->rng.seed(seed);
	s->rng_ready = false;
	return 0;
}

int gcm_b200_set_kernel_mode(gcm_b200_set* s, int fast_inner)
{
	NEED_SET(s);
	s->fast_inner = fast_inner != 0;
	return 0;
}

int gcm_b200_set_concurrency(gcm_b200_set* s, int on)
{
	NEED_SET(s);
	s->concurrent = on != 0;
	return 0;
}

int gcm_b200_set_basis_mode(gcm_b200_set* s, int on_device)
{
	NEED_SET(s);
	s->basis_on_device = on_device != 0;
	s->rng_ready = false;
	return 0;
}

int gcm_b200_set_collision_detector(gcm_b200_set* s, int type, int is_static, float threshold)
{
	NEED_SET(s);
	if(type != 0 && type != 1) return fail(gcmh::UNIMPLEMENTED_EXCEPTION, "unknown collision detector (0 = Void, 1 = Bruteforce)");
	if(type != 0 && s->local_zones.size() != (size_t)s->n_zones)
		return fail(gcmh::UNIMPLEMENTED_EXCEPTION, "contact between bodies of different processes (sync_faces_in_intersection / sync_tetrs) is not built");
	s->detector_type = type; s->detector_static = is_static; s->detector_threshold = threshold;
	return 0;
}

int gcm_b200_set_contact_calculator(gcm_b200_set* s, int type)
{
	NEED_SET(s);
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
	if(s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "zones cannot be reloaded after preprocess");
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
	hm.mat.assign(4*(size_t)n_nodes, 0.f);
	if(material) memcpy(hm.mat.data(), material, 4*(size_t)n_nodes * sizeof(float));
	hm.values.assign(9*(size_t)n_nodes, 0.f);
	hm.current_time = 0;
	z->loaded = true;
	s->plan_built = false;
	return 0;
	GUARD_END
}

int gcm_b200_zone_load_msh_file(gcm_b200_set* s, int zone, const char* path, float la, float mu, float rho, float yield_limit)
{
	GUARD_BEGIN
	if(!path) return fail(gcmh::MESH_EXCEPTION, "Can not open msh file");
	gcmh::MshData m;
	gcmh::read_msh_file(path, m);
	const int n = (int)m.placement.size();
	std::vector<float> mat(4 * (size_t)n);
	for(int i = 0; i < n; i++) { mat[4*(size_t)i] = la; mat[4*(size_t)i+1] = mu; mat[4*(size_t)i+2] = rho; mat[4*(size_t)i+3] = yield_limit; }
	return gcm_b200_zone_load(s, zone, n, (int)(m.tets.size() / 4), m.coords.data(), m.placement.data(), m.remote_zone.data(),
	                          m.remote_num.data(), m.tets.data(), mat.data());
	GUARD_END
}

int gcm_b200_read_zones_map(const char* path, int* zone_rank, int capacity, int* n_zones)
{
	GUARD_BEGIN
	if(!path || !n_zones) return fail(gcmh::CONFIG_EXCEPTION, "Can not open zone map file");
	std::vector<int> z;
	gcmh::read_zones_map(path, z);
	*n_zones = (int)z.size();
	if(zone_rank) for(int k = 0; k < (int)z.size() && k < capacity; k++) zone_rank[k] = z[k];
	return 0;
	GUARD_END
}

int gcm_b200_zone_dump_vtk(gcm_b200_set* s, int zone, const char* path)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !path) return fail(gcmh::SNAP_EXCEPTION, "zone not loaded");
	HostMesh& hm = z->hm;
	gcmh::VtuData d;
	d.coords = hm.coords; d.material = hm.mat; d.tets = hm.tets;
	d.values.resize(9 * (size_t)hm.N);
	if(int rc = gcm_b200_zone_get_values(s, zone, d.values.data())) return rc;
	d.flags.resize(hm.N);
	// Node::getFlags(): the reference's bit layout (Node.h:4-27) is the one the host mesh keeps
	for(int i = 0; i < hm.N; i++) d.flags[i] = hm.flags[i];
	if(s->dm.ready && s->device >= 0) {
		d.destruction.resize(8 * (size_t)hm.N);
		if(int rc = gcm_b200_zone_get_destruction_criterias(s, zone, d.destruction.data())) return rc;
	}
	gcmh::write_vtu_file(path, d);
	return 0;
	GUARD_END
}

int gcm_b200_zone_load_vtu_file(gcm_b200_set* s, int zone, const char* path)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !path) return fail(gcmh::CONFIG_EXCEPTION, "zone is not local to this process");
	gcmh::VtuData d;
	gcmh::read_vtu_file(path, d);
	const int n = (int)d.flags.size();
	// load_vtu_file keeps the stored flags: placement comes from them; a snapshot holds coordinates for every node
	std::vector<int> placement(n), rz(n, -1), rn(n, -1);
	for(int i = 0; i < n; i++) placement[i] = (d.flags[i] & gcm::NF_LOCAL) ? 1 : 0;
	for(int i = 0; i < n; i++) if(!placement[i]) return fail(gcmh::UNIMPLEMENTED_EXCEPTION, "vtu restart of a zone with halo nodes (the snapshot does not store their owners)");
	if(int rc = gcm_b200_zone_load(s, zone, n, (int)(d.tets.size() / 4), d.coords.data(), placement.data(), rz.data(), rn.data(), d.tets.data(), d.material.data())) return rc;
	memcpy(z->hm.values.data(), d.values.data(), 9 * (size_t)n * sizeof(float));
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
	if(s->dm.ready) { CK(cudaSetDevice(s->device)); upload_values(s, *z); }
	return 0;
	GUARD_END
}

int gcm_b200_zone_get_values(gcm_b200_set* s, int zone, float* values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !z->loaded || !values) return fail(gcmh::CONFIG_EXCEPTION, "zone not loaded");
	HostMesh& hm = z->hm;
	if(s->dm.ready && hm.N) {
		CK(cudaSetDevice(s->device));
		DeviceMesh& dm = s->dm;
		const int n = hm.N;
		float* stage = dm.aos.p + 9 * z->node_off;
		CK(cudaStreamSynchronize(s->io_in));      // an *_async upload may still be filling this staging slice
		gcm::records_to_values_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(stage, dm.ucur() + 4 * z->node_off, dm.rs, n);
		s->launches++;
		CK(cudaMemcpyAsync(hm.values.data(), stage, 9 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	memcpy(values, hm.values.data(), 9*(size_t)hm.N * sizeof(float));
	return 0;
	GUARD_END
}

// The *_async pair is a three-stage pipeline over two extra streams, so that a caller who keeps the state in pinned
// host memory and streams it through every step pays max(H2D, step + D2H) instead of their sum (PCIe is full duplex):
//   io_in  : H2D copy of the zone's values into the input staging slice   (waits until the slice was consumed)
//   stream : transpose into the records -- the step -- transpose into the output staging slice
//   io_out : D2H copy of the output slice                                  (the next transpose-out waits for it)
// Separate staging buffers for the two directions; one event per hand-over and zone.
static void zone_io_events(Zone& z)
{
	if(z.ev_h2d_done) return;
	CK(cudaEventCreateWithFlags(&z.ev_h2d_done, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&z.ev_in_consumed, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&z.ev_out_ready, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&z.ev_d2h_done, cudaEventDisableTiming));
}

int gcm_b200_zone_upload_values_async(gcm_b200_set* s, int zone, const float* host_values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !s->dm.ready || !host_values) return fail(gcmh::CONFIG_EXCEPTION, "zone is not on the device");
	CK(cudaSetDevice(s->device));
	DeviceMesh& dm = s->dm;
	const int n = z->hm.N;
	if(!n) return 0;
	zone_io_events(*z);
	float* stage = dm.aos.p + 9 * z->node_off;
	CK(cudaStreamWaitEvent(s->io_in, z->ev_in_consumed, 0));      // the previous upload of this slice was transposed
	CK(cudaMemcpyAsync(stage, host_values, 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->io_in));
	CK(cudaEventRecord(z->ev_h2d_done, s->io_in));
	CK(cudaStreamWaitEvent(s->stream, z->ev_h2d_done, 0));
	gcm::values_to_records_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(dm.ucur() + 4 * z->node_off, nullptr, dm.rs, stage, dm.flags.p + z->node_off, n, 1);
	s->launches++;
	CK(cudaEventRecord(z->ev_in_consumed, s->stream));
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_zone_download_values_async(gcm_b200_set* s, int zone, float* host_values)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !s->dm.ready || !host_values) return fail(gcmh::CONFIG_EXCEPTION, "zone is not on the device");
	CK(cudaSetDevice(s->device));
	DeviceMesh& dm = s->dm;
	const int n = z->hm.N;
	if(!n) return 0;
	zone_io_events(*z);
	if(!dm.aos_out.p) dm.aos_out.alloc(9 * dm.N);
	float* stage = dm.aos_out.p + 9 * z->node_off;
	CK(cudaStreamWaitEvent(s->stream, z->ev_d2h_done, 0));        // the previous download of this slice left the device
	gcm::records_to_values_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(stage, dm.ucur() + 4 * z->node_off, dm.rs, n);
	s->launches++;
	CK(cudaEventRecord(z->ev_out_ready, s->stream));
	CK(cudaStreamWaitEvent(s->io_out, z->ev_out_ready, 0));
	CK(cudaMemcpyAsync(host_values, stage, 9 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->io_out));
	CK(cudaEventRecord(z->ev_d2h_done, s->io_out));
	return 0;
	GUARD_END
}

int gcm_b200_set_join_io(gcm_b200_set* s)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	for(int zi : s->local_zones) {
		Zone& z = *s->zones[zi];
		if(z.ev_d2h_done) CK(cudaStreamWaitEvent(s->stream, z.ev_d2h_done, 0));
	}
	return 0;
	GUARD_END
}

// One BorderCondition without its area: calculator + StepPulseForm.  Returns its id through *id.
static int define_border_condition(gcm_b200_set* s, int calc_type, const float* params, const int* vars, const float* vals,
                                   float start, float duration, const float* box, int* id)
{
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "call gcm_b200_set_preprocess first");
	if((int)s->conds.size() >= GCM_MAX_BORDER_CONDS) return fail(gcmh::CONFIG_EXCEPTION, "too many border conditions");
	const bool dir_normalised = (calc_type & GCM_B200_DIR_NORMALISED) != 0;
	calc_type &= ~GCM_B200_DIR_NORMALISED;
	if(calc_type < 0 || calc_type > 4) return fail(gcmh::CONFIG_EXCEPTION, "bad border condition");
	CondSpec cs;
	memset(&cs, 0, sizeof(cs));
	cs.bc.type = calc_type;
	if(calc_type == gcm::BC_EXT_FORCE || calc_type == gcm::BC_EXT_VELOCITY) {
		if(!params) return fail(gcmh::CONFIG_EXCEPTION, "missing calculator parameters");
		// set_parameters, ExternalForceCalculator.cpp:19-27
		cs.bc.p_normal = params[0]; cs.bc.p_tangential = params[1];
		float dtmp = dir_normalised ? 1.0f : sqrtf(params[2]*params[2] + params[3]*params[3] + params[4]*params[4]);
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
	if(box) memcpy(cs.box, box, 6 * sizeof(float));
	*id = (int)s->conds.size();
	s->conds.push_back(cs);
	return 0;
}

int gcm_b200_set_add_border_condition(gcm_b200_set* s, int calc_type, const float* params,
                                      const int* vars, const float* vals,
                                      float start, float duration, const float* box)
{
	GUARD_BEGIN
	if(!box) return fail(gcmh::CONFIG_EXCEPTION, "bad border condition");
	int id = -1;
	if(int rc = define_border_condition(s, calc_type, params, vars, vals, start, duration, box, &id)) return rc;
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
		if(s->dm.ready) s->dm.cond_id.put(hm.cond_id, z.slot_off, s->stream);
	}
	return 0;
	GUARD_END
}

int gcm_b200_set_define_border_condition(gcm_b200_set* s, int calc_type, const float* params,
                                         const int* vars, const float* vals, float start, float duration, int* id)
{
	GUARD_BEGIN
	if(!id) return fail(gcmh::CONFIG_EXCEPTION, "null id");
	return define_border_condition(s, calc_type, params, vars, vals, start, duration, nullptr, id);
	GUARD_END
}

int gcm_b200_zone_assign_border_conditions(gcm_b200_set* s, int zone, const int* cond_of_node)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !s->preprocessed || !cond_of_node) return fail(gcmh::CONFIG_EXCEPTION, "zone is not preprocessed");
	HostMesh& hm = z->hm;
	for(size_t b = 0; b < hm.border_nodes.size(); b++) {
		const int id = cond_of_node[hm.border_nodes[b]];
		if(id >= (int)s->conds.size()) return fail(gcmh::CONFIG_EXCEPTION, "unknown border condition id");
		hm.cond_id[b] = id < 0 ? -1 : id;
	}
	if(s->dm.ready) { CK(cudaSetDevice(s->device)); s->dm.cond_id.put(hm.cond_id, z->slot_off, s->stream); }
	return 0;
	GUARD_END
}

int gcm_b200_set_preprocess(gcm_b200_set* s)
{
	GUARD_BEGIN
	NEED_SET(s);
	if(s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is already preprocessed");
	for(int zi : s->local_zones)
		if(!s->zones[zi]->loaded) return fail(gcmh::CONFIG_EXCEPTION, "a local zone was not loaded");
	build_halo_plan(s);
	host_sync_nodes(s);
	for(int zi : s->local_zones) s->zones[zi]->hm.pre_process_mesh();
	assign_offsets(s);
	if(s->device >= 0) {
		CK(cudaSetDevice(s->device));
		upload_mesh(s);
	}
	s->preprocessed = true;
	return 0;
	GUARD_END
}

int gcm_b200_set_begin_step(gcm_b200_set* s)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	s->time_step = 0.002f;   // TetrMeshSet.cpp:105-106
	begin_step(s);
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_set_sync_nodes(gcm_b200_set* s)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	sync_nodes_device(s);
	return 0;
	GUARD_END
}

int gcm_b200_zone_do_next_part_step(gcm_b200_set* s, int zone, float tau, int stage)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	Zone* z = get_zone(s, zone);
	if(!z) return fail(gcmh::CONFIG_EXCEPTION, "zone is not local to this process");
	if(stage < 0 || stage > 3) return fail(gcmh::METHOD_EXCEPTION, "Bad stage number");
	if(z->stage_done) return fail(gcmh::CONFIG_EXCEPTION, "zone already ran this stage; every local zone must run a stage before the next one");
	run_stage(s, zone, tau, stage);
	if(stage < 3) {
		z->stage_done = true;
		bool all = true;
		for(int zi : s->local_zones) all = all && s->zones[zi]->stage_done;
		if(all) finish_axis_stage(s);
	}
	return 0;
	GUARD_END
}

int gcm_b200_set_do_next_part_step(gcm_b200_set* s, float tau, int stage)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	if(stage < 0 || stage > 3) return fail(gcmh::METHOD_EXCEPTION, "Bad stage number");
	run_stage(s, -1, tau, stage);
	if(stage < 3) finish_axis_stage(s);
	return 0;
	GUARD_END
}

int gcm_b200_set_end_step(gcm_b200_set* s)
{
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	// proceed_rheology: VoidRheologyCalculator (no-op); calc_destruction_criterias: evaluated on
	// download; update_current_time (TetrMesh_1stOrder.cpp:2403-2406)
	if(s->track_destruction && s->dm.ready && s->dm.N) {
		// calc_destruction_criterias (TetrMeshSet.cpp:177-178): the criteria of the current values are evaluated when
		// they are read (gcm_b200_zone_get_destruction_criterias); the running maxima need every step
		GUARD_BEGIN
		CK(cudaSetDevice(s->device));
		DeviceMesh& dm = s->dm;
		gcm::destruction_history_kernel<<<(int)((dm.N + 255) / 256), 256, 0, s->stream>>>(dm.ucur(), dm.rs, dm.flags.p, (int)dm.N,
			reinterpret_cast<float4*>(dm.destr_hist.p));
		s->launches++;
		CK(cudaGetLastError());
		GUARD_END
	}
	for(int zi : s->local_zones) s->zones[zi]->hm.current_time += s->time_step;
	s->step_count++;
	return 0;
}

int gcm_b200_set_track_destruction(gcm_b200_set* s, int on)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	DeviceMesh& dm = s->dm;
	if(on && !dm.destr_hist.p && dm.N) {
		dm.destr_hist.alloc(4 * dm.N);
		CK(cudaMemsetAsync(dm.destr_hist.p, 0, 4 * dm.N * sizeof(float), s->stream));
	}
	s->track_destruction = on != 0;
	return 0;
	GUARD_END
}

int gcm_b200_zone_get_destruction_criterias(gcm_b200_set* s, int zone, float* out8)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	Zone* z = get_zone(s, zone);
	if(!z || !out8) return fail(gcmh::CONFIG_EXCEPTION, "zone is not local to this process");
	DeviceMesh& dm = s->dm;
	const int n = z->hm.N;
	if(!n) return 0;
	DevBuf<float> tmp; tmp.alloc(8 * (size_t)n);
	// before the first step every criterion is 0 (load_msh_file zeroes them, TetrMesh_1stOrder.cpp:793)
	gcm::destruction_gather_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(tmp.p, dm.ucur(), dm.rs, dm.flags.p,
		reinterpret_cast<const float4*>(dm.destr_hist.p), (int)z->node_off, n, s->step_count > 0 ? 1 : 0);
	s->launches++;
	CK(cudaMemcpyAsync(out8, tmp.p, 8 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	return 0;
	GUARD_END
}

int gcm_b200_comm_get_unique_id(void* id128)
{
	GUARD_BEGIN
	if(!id128) return fail(gcmh::CONFIG_EXCEPTION, "null id buffer");
	NcclId id;
	nccl_check(nccl_api().GetUniqueId(&id), "get unique id");
	memcpy(id128, &id, sizeof id);
	return 0;
	GUARD_END
}

int gcm_b200_set_comm_init(gcm_b200_set* s, const void* id128, int rank, int n_ranks)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	if(!id128 || rank < 0 || rank >= n_ranks) return fail(gcmh::CONFIG_EXCEPTION, "bad communicator arguments");
	if(s->nccl_comm) return fail(gcmh::CONFIG_EXCEPTION, "communicator already initialised");
	NcclId id;
	memcpy(&id, id128, sizeof id);
	nccl_check(nccl_api().CommInitRank(&s->nccl_comm, n_ranks, id, rank), "comm init");
	s->comm_rank = rank; s->comm_size = n_ranks;
	{
		// Every pair of processes must agree on what travels between them, or the grouped send/recv of the first
		// exchange would hang without a word: all-gather (nodes I send to r, nodes I expect from r) and compare.
		NcclApi& nc = nccl_api();
		const int R = n_ranks;
		std::vector<int> mine(2 * (size_t)R, 0), all(2 * (size_t)R * R, 0);
		for(auto& kv : s->peers) {
			if(kv.first < 0 || kv.first >= R) return fail(gcmh::SYNC_EXCEPTION, "halo plan names a rank outside the communicator");
			mine[kv.first] = (int)kv.second.send_node.size();
			mine[R + kv.first] = (int)kv.second.recv_node.size();
		}
		DevBuf<int> d1, dall; d1.alloc(2 * (size_t)R); dall.alloc(2 * (size_t)R * R);
		CK(cudaMemcpyAsync(d1.p, mine.data(), mine.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
		nccl_check(nc.AllGather(d1.p, dall.p, 2 * (size_t)R, 2 /* ncclInt32 */, s->nccl_comm, s->stream), "all gather");
		CK(cudaMemcpyAsync(all.data(), dall.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		for(int a = 0; a < R; a++)
			for(int b = 0; b < R; b++)
				if(all[(size_t)a * 2 * R + b] != all[(size_t)b * 2 * R + R + a]) {
					char buf[200];
					snprintf(buf, sizeof buf, "halo plans disagree: rank %d sends %d nodes to rank %d, which expects %d (gcm_b200_halo_set_sends not called with the peer's requests?)",
					         a, all[(size_t)a * 2 * R + b], b, all[(size_t)b * 2 * R + R + a]);
					return fail(gcmh::SYNC_EXCEPTION, buf);
				}
	}
	p2p_setup(s);
	return 0;
	GUARD_END
}

int gcm_b200_set_comm_sync_nodes(gcm_b200_set* s)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	if(!s->nccl_comm) return fail(gcmh::CONFIG_EXCEPTION, "no communicator: call gcm_b200_set_comm_init first");
	comm_exchange(s, s->stream);
	return 0;
	GUARD_END
}

int gcm_b200_set_do_next_step(gcm_b200_set* s)
{
	if(!s || !s->preprocessed) return fail(gcmh::CONFIG_EXCEPTION, "set is not preprocessed");
	if(!s->peers.empty() && !s->nccl_comm)
		return fail(gcmh::CONFIG_EXCEPTION, "multi-process set: call gcm_b200_set_comm_init, or drive the stages with gcm_b200_halo_* between them");
	int rc = gcm_b200_set_begin_step(s);
	if(rc) return rc;
	// Whole-step call: the plastic stage (per node, no neighbours) runs as the epilogue of the stage-2
	// writers, and the halo refresh in front of it -- never read: stage 3 touches local nodes only and the
	// next step's first refresh overwrites the same slots -- is not issued.
	const bool fuse = s->fuse_plastic;
	for(int stage = 0; stage < (fuse ? 3 : 4); stage++) {
		if((rc = gcm_b200_set_sync_nodes(s))) return rc;
		GUARD_BEGIN
		NEED_DEVICE(s);
		if(s->nccl_comm && stage == 3) comm_exchange(s, s->stream);     // only when stage 3 is its own pass (GCM_B200_FUSE_PLASTIC=0)
		run_stage(s, -1, s->time_step, stage, fuse, s->nccl_comm != nullptr && stage < 3);
		if(stage < 3) finish_axis_stage(s);
		GUARD_END
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
	NEED_DEVICE(s);
	CK(cudaMemcpyAsync(s->h_err, s->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	CK(cudaStreamSynchronize(s->io_out));      // downloads of the *_async calls have landed
	CK(cudaStreamSynchronize(s->io_in));
	if(s->h_err[0]) {
		char buf[256];
		int zone = -1, node = s->h_err[2];   // kernels report merged node indices
		for(int zi : s->local_zones) {
			Zone& z = *s->zones[zi];
			if((size_t)node >= z.node_off && (size_t)node < z.node_off + z.hm.N) { zone = zi; node -= (int)z.node_off; break; }
		}
		snprintf(buf, sizeof buf, "%s (zone %d, node %d, stage %d)", step_error_text(s->h_err[0]), zone, node, s->h_err[3]);
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
	NEED_SET(s);
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
	NEED_SET(s);
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
	NEED_SET(s);
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

int gcm_b200_halo_pack(gcm_b200_set* s, int peer_rank, float* dev_buf)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	PeerPlan& p = it->second;
	peer_dev_prepare(s, p);
	const int n = (int)p.send_node.size();
	if(n) {
		ProfScope ps(s, PROF_HALO, n);
		gcm::halo_pack_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(dev_buf, s->dm.ucur(), s->dm.rs, p.d_send.p, n, n, 0);
		s->launches++;
	}
	CK(cudaGetLastError());
	return 0;
	GUARD_END
}

int gcm_b200_halo_unpack(gcm_b200_set* s, int peer_rank, const float* dev_buf)
{
	GUARD_BEGIN
	NEED_DEVICE(s);
	auto it = s->peers.find(peer_rank);
	if(it == s->peers.end()) return 0;
	PeerPlan& p = it->second;
	peer_dev_prepare(s, p);
	const int n = (int)p.recv_node.size();
	if(n) {
		ProfScope ps(s, PROF_HALO, n);
		gcm::halo_unpack_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->dm.ucur(), s->dm.rs, dev_buf, p.d_recv.p, n, n, 0);
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
	GUARD_BEGIN
	collisions_to_host(s);      // the contact bit of a device collision pass
	GUARD_END
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
	if(s->dm.ready && s->basis_on_device && !hm.basis.empty()) {
		cudaSetDevice(s->device);
		if(cudaMemcpyAsync(hm.basis.data(), s->dm.basis.p + 9 * z->slot_off, hm.basis.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess
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

int gcm_b200_set_enable_tap(gcm_b200_set* s, int on) { NEED_SET(s); s->tap_on = on != 0; return 0; }

int gcm_b200_zone_get_tap(gcm_b200_set* s, int zone, int* owners, int* outer, int* tries)
{
	GUARD_BEGIN
	Zone* z = get_zone(s, zone);
	if(!z || !s->dm.tap.p) return fail(gcmh::CONFIG_EXCEPTION, "tap was not enabled");
	CK(cudaSetDevice(s->device));
	const size_t n = z->hm.N;
	std::vector<gcm::StageTap> h(n);
	for(int st = 0; st < 3; st++) {
		CK(cudaMemcpyAsync(h.data(), s->dm.tap.p + (size_t)st * s->dm.N + z->node_off, n * sizeof(gcm::StageTap), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		for(size_t i = 0; i < n; i++) {
			const size_t o = (size_t)st * n + i;
			if(owners) for(int k = 0; k < 5; k++) owners[5*o + k] = h[i].owner[k] >= 0 ? h[i].owner[k] - (int)z->tet_off : h[i].owner[k];
			if(outer) outer[o] = h[i].outer_count;
			if(tries) tries[o] = h[i].tries;
		}
	}
	return 0;
	GUARD_END
}

long long gcm_b200_set_launch_count(gcm_b200_set* s) { return s ? s->launches : 0; }

int gcm_b200_set_profile(gcm_b200_set* s, int on)
{
	NEED_SET(s);
	if(s->device >= 0) cudaSetDevice(s->device);
	prof_drain(s);
	s->profile = on != 0;
	if(on) for(int k = 0; k < PROF_N; k++) { s->prof_ms[k] = 0; s->prof_launches[k] = 0; s->prof_units[k] = 0; }
	return 0;
}

int gcm_b200_set_profile_read(gcm_b200_set* s, int which, double* total_ms, long long* launches, long long* units)
{
	if(!s || which < 0 || which >= PROF_N) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	if(s->device >= 0) cudaSetDevice(s->device);
	prof_drain(s);
	if(total_ms) *total_ms = s->prof_ms[which];
	if(launches) *launches = s->prof_launches[which];
	if(units) *units = s->prof_units[which];
	return 0;
}

int gcm_b200_set_kernel_stats(gcm_b200_set* s, long long* out16)
{
	GUARD_BEGIN
	if(!s || !out16) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	for(int k = 0; k < 16; k++) out16[k] = 0;
	DeviceMesh& dm = s->dm;
	if(!dm.ready) return 0;
	CK(cudaSetDevice(s->device));
	out16[0] = dm.icache_ready ? (long long)dm.n_fast - dm.n_cold : 0;
	out16[1] = dm.icache_ready ? dm.n_cold : (long long)dm.n_fast;
	for(int k = 0; k < 3; k++) { out16[2 + k] = dm.n_patterns[k]; out16[5 + k] = dm.pattern_ok[k] ? 1 : 0; }
	out16[8] = (long long)dm.n_flat; out16[9] = (long long)dm.n_sharp;
	if(dm.bnd_defer.p) {
		CK(cudaMemcpyAsync(s->h_err + 4, dm.bnd_defer.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		out16[10] = s->h_err[4];
	}
	out16[11] = dm.exact_only;
	out16[12] = s->p2p_ready ? 1 : 0;
	out16[13] = dm.n_shell;
	return 0;
	GUARD_END
}

int gcm_b200_set_virt_count(gcm_b200_set* s)
{
	if(!s) return 0;
	try { collisions_to_host(s); } catch(...) { return 0; }
	return (int)s->virt.size();
}

int gcm_b200_set_get_virt(gcm_b200_set* s, float* out16)
{
	GUARD_BEGIN
	if(!s || !out16) return fail(gcmh::CONFIG_EXCEPTION, "bad arguments");
	collisions_to_host(s);
	// the friction calculator writes into its virtual node (the swapped-argument Adhesion call of
	// FrictionContactCalculator.cpp:123): the device copy is the current one
	std::vector<gcm::VirtNode> dv;
	if(s->d_virt.p && s->d_virt.n >= s->virt.size() && !s->virt.empty() && s->device >= 0) {
		CK(cudaSetDevice(s->device));
		dv.resize(s->virt.size());
		CK(cudaMemcpyAsync(dv.data(), s->d_virt.p, dv.size() * sizeof(gcm::VirtNode), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	for(size_t k = 0; k < s->virt.size(); k++) {
		memcpy(out16 + 16 * k, s->virt[k].coords, 3 * sizeof(float));
		memcpy(out16 + 16 * k + 3, s->virt[k].values, 13 * sizeof(float));
		if(!dv.empty()) memcpy(out16 + 16 * k + 3, dv[k].val, 9 * sizeof(float));
	}
	return 0;
	GUARD_END
}

} // extern "C"
