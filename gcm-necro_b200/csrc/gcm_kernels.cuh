// gcm_kernels.cuh -- sm_100a kernels of the GCM stage.  Included by gcm_capi.cu only.
#pragma once
#include <cuda_runtime.h>
#include "gcm_step.cuh"

namespace gcm {

// err[0] = StepError (0 = none), err[1] = zone, err[2] = node, err[3] = stage; first error wins.
__device__ __forceinline__ void report_error(int* err, int code, int zone, int node, int stage)
{
	if(atomicCAS(&err[0], 0, code) == 0) { err[1] = zone; err[2] = node; err[3] = stage; }
}

} // namespace gcm
#include "gcm_inner.cuh"
#include "gcm_border.cuh"
#include "gcm_collide.cuh"
namespace gcm {

// Writes one node's stage result: the 9 values into the next layer; with yield_plane != NULL the
// plastic stage is folded in (whole-step calls run stage 3 as the epilogue of stage 2).
__device__ __forceinline__ void store_node(float* __restrict__ un, size_t rs, int node, float out[9], const float* __restrict__ yield_plane,
                                           int* err, int zone)
{
	if(yield_plane && !plastic_epilogue(out, yield_plane[node])) { report_error(err, ERR_NAN, zone, node, 3); return; }
	rec_quad_store(un, rs, node, 0, make_float4(out[0], out[1], out[2], out[3]));
	rec_quad_store(un, rs, node, 1, make_float4(out[4], out[5], out[6], out[7]));
	un[rec_index(rs, node, 8)] = out[8];      // the coordinates in the rest of the third quad are static
}

// One thread per listed node, one axis stage: TetrMesh_1stOrder::do_next_part_step's two
// node loops (gcm/meshtypes/TetrMesh_1stOrder.cpp:1906-1923); the copy-back loop (:1925-1932)
// is the caller's buffer swap.
__global__ void __launch_bounds__(128)
stage_generic_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list,
                     int stage, float* __restrict__ un, int* err, int zone, StageTap* tap, const float* __restrict__ yield_plane,
                     const int* __restrict__ skip_contact = nullptr)
{
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= n_list) return;
	int node = list[idx];
	// device-built contact state: a border node in contact on this axis belongs to the contact kernel
	if(skip_contact && skip_contact[m.border_slot[node]] >= 0) return;
	float out[9];
	StageTap t;
	int e = node_stage_nocontact(m, node, stage, out, tap ? &t : nullptr);
	if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
	if(e) { report_error(err, e, zone, node, stage); return; }
	store_node(un, m.rs, node, out, yield_plane, err, zone);
}

// The nodes a specialised kernel could not settle (cold[0] = how many, cold[1..] = which), through
// the literal per-node routine; a small fixed grid strides over the list (it is empty on regular
// meshes, a handful of nodes on irregular ones).
__global__ void __launch_bounds__(128)
stage_deferred_kernel(const __grid_constant__ MeshView m, const int* __restrict__ cold, int stage,
                      float* __restrict__ un, int* err, int zone, StageTap* tap, const float* __restrict__ yield_plane,
                      int node_begin, int node_end)
{
	const int n = cold[0];
	for(int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
		const int node = cold[1 + idx];
		if(node < node_begin || node >= node_end) continue;   // a per-zone call: the list covers every local zone
		float out[9];
		StageTap t;
		const int e = node_stage_nocontact_cold(m, node, stage, out, tap ? &t : nullptr);
		if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
		if(e) { report_error(err, e, zone, node, stage); continue; }
		store_node(un, m.rs, node, out, yield_plane, err, zone);
	}
}

// Border nodes without contact (gcm_border.cuh), one thread per node.  LU: the launch provides the shared-memory
// 9x9 double systems (element-major across the block: thread t owns column t of an [81][BLOCK] array),
// needed on the normal axis (stage 0) where three characteristics leave the body; the tangential
// stages run without it at full occupancy.  Whatever the routine defers (ambiguous candidate, alpha
// retries, ...) is appended to `defer` (defer[0] = count) for the literal routine right behind.
template<bool LU, int BLOCK>
__global__ void __launch_bounds__(BLOCK, LU ? (BLOCK > 64 ? 1 : 5) : 4)
stage_border_fast_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list, int stage,
                         float* __restrict__ un, int* err, int zone, StageTap* tap, const float* __restrict__ yield_plane,
                         const BndSlotStatic* __restrict__ bstat, const float4* __restrict__ cand, int* __restrict__ defer,
                         const int* __restrict__ skip_contact = nullptr)
{
	// BLOCK <= 64: static array, 5 blocks per SM; BLOCK = 352: one block per SM on 228 096 bytes of dynamic shared
	// memory (the whole SM: 352 instead of 320 systems resident)
	__shared__ double sA_static[(LU && BLOCK <= 64) ? 81 * BLOCK : 1];
	extern __shared__ double sA_dynamic[];
	double* sA = BLOCK > 64 ? sA_dynamic : sA_static;
	const int idx = blockIdx.x * BLOCK + threadIdx.x;
	if(idx >= n_list) return;
	const int node = list[idx];
	if(skip_contact && skip_contact[m.border_slot[node]] >= 0) return;     // in contact on this axis: the contact kernel's node
	BndSlotStatic bs;
	{
		const uint4* q = reinterpret_cast<const uint4*>(bstat + m.border_slot[node]);
		const uint4 a = __ldg(q), b = __ldg(q + 1);
		bs.deg = (int)a.x; bs.e0 = (int)a.y; bs.hmin = __uint_as_float(a.z); bs.dmin2 = __uint_as_float(a.w);
		bs.w0 = __uint_as_float(b.x); bs.cand_off = (int)b.y; bs.h0 = __uint_as_float(b.z); bs.pad1 = 0;
	}
	StridedMat9<BLOCK> A;
	A.base = sA + threadIdx.x;
	float out[9];
	StageTap tp;
	const int e = node_stage_border_fast(m, bs, cand + bs.cand_off, node, stage, out, tap ? &tp : nullptr, LU ? &A : (StridedMat9<BLOCK>*)nullptr);
	if(e == GCM_BND_DEFER) { defer[1 + atomicAdd(&defer[0], 1)] = node; return; }
	if(tap) tap[(size_t)stage * m.n_nodes + node] = tp;
	if(e) { report_error(err, e, zone, node, stage); return; }
	store_node(un, m.rs, node, out, yield_plane, err, zone);
}

// Border nodes in contact on axis `stage` (only axis 0 can be: BruteforceCollisionDetector.cpp:58-60):
// the contact branch of do_next_part_step (TetrMethod_Plastic_1stOrder.cpp:299-421), one thread per
// node.  list[i] = real node, its virtual node = axis_plus0[border_slot].  partner_next = 1 when
// the partner body was already advanced in this stage (the reference advances its meshes one
// after the other): the virtual node's feet then read the NEXT layer `un`.
//
// Two passes.  The first (m.ring_ws == NULL, one thread per listed node) searches the first two
// rings without any list; a node whose search would need a third ring (virtual nodes on
// non-matching surface meshes) is appended to `deep` (deep[0] = count, deep[1..] = nodes) instead
// of being advanced.  The second pass (deep_pass != 0, a fixed small grid whose threads each own a
// find_owner_general workspace in m.ring_ws) strides over that list.  No node reads another
// contact node's result of the same launch, so the split does not change any value.
__global__ void __launch_bounds__(64)
stage_contact_kernel(const __grid_constant__ MeshView m, const int* __restrict__ list, int n_list, int stage,
                     float* __restrict__ un, VirtNode* __restrict__ virt, const int* __restrict__ axis_plus0,
                     const float* __restrict__ normals, int calc_type, int partner_next, int first_step,
                     int* err, int zone, StageTap* tap, int* deep, int deep_pass,
                     const int* __restrict__ sel = nullptr, const int* __restrict__ sel_phase = nullptr, int want_phase = 0)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int n_work = deep_pass ? deep[0] : n_list;
	const int step = deep_pass ? gridDim.x * blockDim.x : n_work;
	for(int idx = tid; idx < n_work; idx += step) {
		const int node = deep_pass ? deep[1 + idx] : list[idx];
		const int slot = m.border_slot[node];
		const int vidx = axis_plus0[slot];
		// device-built lists (gcm_collide.cuh): `list` holds every candidate of the collision pass; an entry runs if
		// it hit a face, its virtual node is the one the real node ended up with, and its pair has the wanted phase
		if(sel && !deep_pass && (sel[idx] < 0 || sel[idx] != vidx || sel_phase[idx] != want_phase)) continue;
		MeshView mv = m;
		if(partner_next) mv.u = un;
		VirtNode vn = virt[vidx];
		float out[9], vout[9];
		bool vw = false;
		StageTap t;
		const float nrm[3] = {normals[3 * (size_t)slot], normals[3 * (size_t)slot + 1], normals[3 * (size_t)slot + 2]};
		int e = node_stage_contact(m, mv, node, vn, vidx, stage, nrm, calc_type, out, tap ? &t : nullptr, vout, &vw, first_step != 0);
		if(e == ERR_RING_OVERFLOW && !deep_pass && deep) { deep[1 + atomicAdd(&deep[0], 1)] = node; continue; }
		if(tap) tap[(size_t)stage * m.n_nodes + node] = t;
		if(e) { report_error(err, e, zone, node, stage); continue; }
		store_node(un, m.rs, node, out, nullptr, err, zone);
		if(vw) for(int k = 0; k < 9; k++) virt[vidx].val[k] = vout[k];
	}
}

// Gather the records of listed nodes (collision discovery reads border-node values on the host).
__global__ void gather_records_kernel(float* __restrict__ out9, const float* __restrict__ u, size_t rs, const int* __restrict__ list, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const int node = list[i];
	#pragma unroll
	for(int k = 0; k < 9; k++) out9[9 * (size_t)i + k] = rec_get(u, rs, node, k);
}

// Stage 3: the plastic corrector drop_deviator (TetrMethod_Plastic_1stOrder.cpp:51-71,192-199),
// in place on the current layer for every local node.
__global__ void __launch_bounds__(256)
plastic_kernel(float* __restrict__ u, size_t rs, const float* __restrict__ mat, const int* __restrict__ flags,
               int first, int n_nodes, size_t stride, int* err, int zone)   // u: node records, mat: planes
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_nodes) return;
	i += first;
	if((flags[i] & (NF_LOCAL | NF_USED)) != (NF_LOCAL | NF_USED)) return;
	float v[9];
	bool bad = false;
	#pragma unroll
	for(int k = 0; k < 9; k++) {
		v[k] = rec_get(u, rs, i, k);
		bad |= (isnan(v[k]) || isinf(v[k]));
	}
	if(bad) { report_error(err, ERR_NAN, zone, i, 3); return; }
	float yield_limit = mat[3 * stride + i];
	if(drop_deviator(v, yield_limit)) {
		#pragma unroll
		for(int k = 3; k < 9; k++)
			rec_set(u, rs, i, k, v[k]);
	}
}

// calc_destruction_criterias (TetrMesh_1stOrder.cpp:1957-1995), end of every step when tracking is on: the
// running maxima hist[node] = (compression, tension, shear, deviator)_history of local nodes.
__global__ void __launch_bounds__(256)
destruction_history_kernel(const float* __restrict__ u, size_t rs, const int* __restrict__ flags, int n, float4* __restrict__ hist)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	if((flags[i] & (NF_LOCAL | NF_USED)) != (NF_LOCAL | NF_USED)) return;
	float v[9], c[4];
	#pragma unroll
	for(int k = 0; k < 9; k++) v[k] = rec_get(u, rs, i, k);
	destruction_now(v, c);
	float4 h = hist[i];
	if(c[0] > h.x) h.x = c[0];
	if(c[1] > h.y) h.y = c[1];
	if(c[2] > h.z) h.z = c[2];
	if(c[3] > h.w) h.w = c[3];
	hist[i] = h;
}

// ElasticNode::destruction_criterias[8] of a node range, node-major: the four criteria of the current values, then
// their running maxima (zeros when tracking was never switched on; remote nodes: zeros, as load_msh_file leaves them).
__global__ void __launch_bounds__(256)
destruction_gather_kernel(float* __restrict__ out8, const float* __restrict__ u, size_t rs, const int* __restrict__ flags,
                          const float4* __restrict__ hist, int first, int n, int have_now)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if(k >= n) return;
	const int i = first + k;
	float c[4] = {0.f, 0.f, 0.f, 0.f};
	float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
	if(have_now && (flags[i] & (NF_LOCAL | NF_USED)) == (NF_LOCAL | NF_USED)) {
		float v[9];
		#pragma unroll
		for(int q = 0; q < 9; q++) v[q] = rec_get(u, rs, i, q);
		destruction_now(v, c);
		if(hist) h = hist[i];
	}
	float4* o = reinterpret_cast<float4*>(out8 + 8 * (size_t)k);
	o[0] = make_float4(c[0], c[1], c[2], c[3]);
	o[1] = h;
}

// DataBus::sync_nodes, same-process branch (gcm/system/DataBus.cpp:64-79): halo node records of one
// zone copied from the owner zone's nodes, three threads per node, one 16-byte quad each (the third
// quad carries the coordinates, which are equal on both sides since the start-up sync).
__global__ void __launch_bounds__(256)
halo_copy_kernel(float* __restrict__ dst_u, const int* __restrict__ dst_idx,
                 const float* __restrict__ src_u, const int* __restrict__ src_idx, int n, size_t rs)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	const int q = g / n, i = g - q * n;        // plane-major: consecutive threads, consecutive list entries
	if(q >= 3) return;
	const int d = dst_idx[i], s = src_idx[i];
	rec_quad_store(dst_u, rs, d, q, rec_quad(src_u, rs, s, q));
}

// Cross-process halo: gather into / scatter from a plane-major [9][n] exchange buffer.
__global__ void halo_pack_kernel(float* __restrict__ buf, const float* __restrict__ u, size_t rs,
                                 const int* __restrict__ idx, int n, int n_total, int offset)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	int s = idx[i];
	#pragma unroll
	for(int k = 0; k < 9; k++)
		buf[(size_t)k * n_total + offset + i] = rec_get(u, rs, s, k);
}

__global__ void halo_unpack_kernel(float* __restrict__ u, size_t rs, const float* __restrict__ buf,
                                   const int* __restrict__ idx, int n, int n_total, int offset)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	int d = idx[i];
	#pragma unroll
	for(int k = 0; k < 9; k++)
		rec_set(u, rs, d, k, buf[(size_t)k * n_total + offset + i]);
}

// ---- cross-process halo by direct peer stores (DataBus::sync_nodes, gcm/system/DataBus.cpp:81-115) -------
// The record buffers and a small flag array of every neighbour process are mapped into this process
// (cudaIpcOpenMemHandle, once); before a stage each process WRITES the records of the nodes its
// neighbours asked for straight into their halo slots over NVLink and then raises its flag in the
// neighbour's flag array; the neighbour waits for the flag before its stage kernels start.  One small
// kernel on either side instead of pack -> ncclSend/ncclRecv -> unpack.
struct PeerTarget {
	float* u[2];                 // the neighbour's two record layers (mapped)
	unsigned long long rs;       // its quads per plane
	int* flag;                   // its flag array (mapped); entry [my rank] is mine to raise
	int begin, end;              // my range in the concatenated send list
};
#define GCM_MAX_PEERS 8

// entries [begin, end) of peer k: quad q of my node send_node[i] -> quad q of the peer's node remote_slot[i].
// The last block to finish (all stores fenced system-wide) raises the flags.
__global__ void __launch_bounds__(256)
halo_push_kernel(const float* __restrict__ u, size_t rs, const int* __restrict__ send_node, const int* __restrict__ remote_slot,
                 int n_total, const __grid_constant__ PeerTarget t0, const __grid_constant__ PeerTarget t1,
                 const PeerTarget* __restrict__ more, int n_peers, int layer, int my_rank, int seq, unsigned* __restrict__ done_counter)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	const int q = g / n_total, i = g - q * n_total;
	if(q < 3) {
		// field by field, no local PeerTarget copy indexed by `layer`: nvcc 12.9 lost the selection of such a copy
		// on the n_peers == 2 path (the copy kept t1's pointers for t0's entries; seen in the SASS and on 4 GPUs)
		const bool second = n_peers > 1 && i >= t1.begin && i < t1.end;
		float* base = second ? (layer ? t1.u[1] : t1.u[0]) : (layer ? t0.u[1] : t0.u[0]);
		unsigned long long prs = second ? t1.rs : t0.rs;
		for(int k = 2; k < n_peers; k++)
			if(i >= more[k].begin && i < more[k].end) { base = layer ? more[k].u[1] : more[k].u[0]; prs = more[k].rs; }
		const float4 v = rec_quad(u, rs, send_node[i], q);
		reinterpret_cast<float4*>(base)[(size_t)q * prs + (size_t)remote_slot[i]] = v;
	}
	// one system-wide fence per block, by the thread that then counts the block in: the barrier orders the block's
	// stores before it, the fence is cumulative (every thread fencing on its own cost 20 us per exchange at 8 GPUs)
	__shared__ bool last;
	__syncthreads();
	if(threadIdx.x == 0) {
		__threadfence_system();
		last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if(last && threadIdx.x < n_peers) {
		int* flag = threadIdx.x == 0 ? t0.flag : (threadIdx.x == 1 ? t1.flag : more[threadIdx.x].flag);
		__threadfence_system();
		*reinterpret_cast<volatile int*>(flag + my_rank) = seq;
		if(threadIdx.x == 0) *done_counter = 0;
	}
}

// The exchange when the copy engines moved the records (contiguous halo layers, cudaMemcpyAsync on the same stream
// in front of this kernel): thread k raises this process's flag at neighbour k, then waits for that neighbour's.
__global__ void halo_flag_wait_kernel(const PeerTarget* __restrict__ targets, const int* __restrict__ flags, const int* __restrict__ peer_ranks,
                                      int n_peers, int my_rank, int seq, int* err)
{
	const int k = threadIdx.x;
	if(k >= n_peers) return;
	__threadfence_system();
	*reinterpret_cast<volatile int*>(targets[k].flag + my_rank) = seq;
	const volatile int* f = flags + peer_ranks[k];
	const long long t0 = clock64();
	while(*f < seq) {
		__nanosleep(100);
		if(clock64() - t0 > 4000000000ll) { if(atomicCAS(&err[0], 0, ERR_PEER_TIMEOUT) == 0) { err[1] = -1; err[2] = peer_ranks[k]; err[3] = seq; } break; }
	}
}

// One thread per neighbour: wait until its flag reaches seq.  A neighbour that never arrives is reported
// (SYNC error) after ~2 s instead of hanging the device.
__global__ void halo_wait_kernel(const int* __restrict__ flags, const int* __restrict__ peer_ranks, int n_peers, int seq, int* err)
{
	const int k = threadIdx.x;
	if(k >= n_peers) return;
	const volatile int* f = flags + peer_ranks[k];
	const long long t0 = clock64();
	while(*f < seq) {
		__nanosleep(200);
		if(clock64() - t0 > 4000000000ll) { if(atomicCAS(&err[0], 0, ERR_PEER_TIMEOUT) == 0) { err[1] = -1; err[2] = peer_ranks[k]; err[3] = seq; } break; }
	}
}

// ---- the per-step random local bases, generated on the device ------------------------------
// create_random_axis (gcm/methods/TetrMethod_Plastic_1stOrder.cpp:697-748) consumes libc rand():
// 3 draws per inner node (discarded: basis := E) and 1 per border node, nodes in index order.
// rand() is the linear recurrence o[n] = o[n-31] + o[n-3] (mod 2^32), value o[n] >> 1, so a
// step's stream is cut into chunks of GCM_RNG_CHUNK draws; one thread owns one chunk, keeps the
// 31-word window at the chunk start, (a) jumps it by one whole step (31x31 matrix A^C over
// Z/2^32, built by the host) for the next call and (b) walks its chunk, storing
// (draw % 360) for the draws that belong to border nodes.
#define GCM_RNG_CHUNK (31 * 32)

__global__ void __launch_bounds__(128)
rng_chunk_kernel(uint32_t* __restrict__ chunk_state, int n_chunks, const uint32_t* __restrict__ jump,
                 const uint32_t* __restrict__ slot_pos, const int* __restrict__ chunk_first_slot,
                 unsigned short* __restrict__ teta_idx)
{
	__shared__ uint32_t J[961];
	for(int i = threadIdx.x; i < 961; i += blockDim.x) J[i] = jump[i];
	__syncthreads();
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c >= n_chunks) return;
	uint32_t w[31];
	#pragma unroll
	for(int i = 0; i < 31; i++) w[i] = chunk_state[(size_t)i * n_chunks + c];
	#pragma unroll
	for(int i = 0; i < 31; i++) {
		uint32_t acc = 0;
		#pragma unroll
		for(int j = 0; j < 31; j++) acc += J[i * 31 + j] * w[j];
		chunk_state[(size_t)i * n_chunks + c] = acc;
	}
	int slot = chunk_first_slot[c];
	const int slot_end = chunk_first_slot[c + 1];
	if(slot == slot_end) return;
	const uint32_t base = (uint32_t)c * GCM_RNG_CHUNK;
	uint32_t next = slot_pos[slot] - base;
	for(int blk = 0; blk < GCM_RNG_CHUNK / 31; blk++) {
		#pragma unroll
		for(int i = 0; i < 31; i++) {
			w[i] += w[(i + 28) % 31];
			if((uint32_t)(blk * 31 + i) == next) {
				teta_idx[slot] = (unsigned short)((w[i] >> 1) % 360u);
				slot++;
				next = slot < slot_end ? slot_pos[slot] - base : 0xffffffffu;
			}
		}
		if(slot >= slot_end) break;
	}
}

// create_rotation_matrix_with_normal for every border node of a zone (:843-901), from the static
// node normal and the drawn angle index; cos/sin of the 360 possible angles come from the host.
__global__ void basis_kernel(float* __restrict__ basis, const float* __restrict__ normals,
                             const unsigned short* __restrict__ teta_idx, int slot_offset, int nb,
                             const float* __restrict__ cos_t, const float* __restrict__ sin_t)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= nb) return;
	int k = teta_idx[slot_offset + s];
	float out[9];
	rotation_basis_with_normal(normals[3 * (size_t)s], normals[3 * (size_t)s + 1], normals[3 * (size_t)s + 2],
	                           cos_t[k], sin_t[k], out);
	#pragma unroll
	for(int q = 0; q < 9; q++) basis[9 * (size_t)s + q] = out[q];
}

// Host-facing state layout is the reference's node-major values[9] (ElasticNode.h:28-44); the
// device keeps 12-float node records.  These kernels sit on either side of a pinned-memory copy;
// values_to_records with local_only writes local nodes only (halo slots belong to sync_nodes).
__global__ void values_to_records_kernel(float* __restrict__ u0, float* __restrict__ u1, size_t rs,
                                         const float* __restrict__ aos, const int* __restrict__ flags, int n, int local_only)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	if(local_only && !(flags[i] & NF_LOCAL)) return;
	#pragma unroll
	for(int k = 0; k < 9; k++) {
		float v = aos[9 * (size_t)i + k];
		rec_set(u0, rs, i, k, v);
		if(u1) rec_set(u1, rs, i, k, v);
	}
}

__global__ void records_to_values_kernel(float* __restrict__ aos, const float* __restrict__ u, size_t rs, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	#pragma unroll
	for(int k = 0; k < 9; k++)
		aos[9 * (size_t)i + k] = rec_get(u, rs, i, k);
}

// Static coordinates into the xyz slots of both layers' records (once, at upload).
__global__ void coords_to_records_kernel(float* __restrict__ u0, float* __restrict__ u1, size_t rs, const float* __restrict__ xyz, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	#pragma unroll
	for(int k = 0; k < 3; k++) {
		float v = xyz[3 * (size_t)i + k];
		rec_set(u0, rs, i, 9 + k, v);
		rec_set(u1, rs, i, 9 + k, v);
	}
}

} // namespace gcm
