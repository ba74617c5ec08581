// gcm_host.h -- host-side (C++) half of the GCM step: the mesh container and everything the
// reference keeps on the CPU around the hot loop -- reverse lookups, border detection,
// border faces, node normals (gcm/meshtypes/TetrMesh_1stOrder.cpp:66-479), the per-step
// random local bases (gcm/methods/TetrMethod_Plastic_1stOrder.cpp:697-919) and the libc
// rand() stream they consume.  Nothing here does the per-node update: that is CUDA only.
#pragma once
#include <cstdint>
#include "gcm_collide.cuh"
#include <string>
#include <vector>

namespace gcmh {

// GCMException codes, gcm/system/GCMException.h:19-29
enum ExcCode { MPI_EXCEPTION = 0, SYNC_EXCEPTION = 1, CONFIG_EXCEPTION = 2, METHOD_EXCEPTION = 3,
               MESH_EXCEPTION = 4, SNAP_EXCEPTION = 5, COLLISION_EXCEPTION = 6, MATH_EXCEPTION = 7,
               UNIMPLEMENTED_EXCEPTION = 8, INPUT_EXCEPTION = 9 };

struct HostError { int code; std::string msg; };

// glibc rand()/srand() (random_r TYPE_3: additive feedback, degree 31, separation 3).
// glibc is a platform dependency of the reference, not under /root/reference; the algorithm
// is restated from its published description so that `srand(seed)` streams match bit for bit
// (checked against the C library in tests/test_host_logic.py).
struct GlibcRand {
	uint32_t r[34];
	int f, b;   // front / rear positions
	void seed(unsigned s);
	inline int next()
	{
		uint32_t v = (r[f] += r[b]);
		if(++f >= 31) f = 0;
		if(++b >= 31) b = 0;
		return (int)(v >> 1);
	}
	void get_window(uint32_t w[31]) const;   // state in stream order (o[n-31] .. o[n-1])
	void set_window(const uint32_t w[31]);
};

// Jump-ahead for the rand() recurrence: out = A^k (31x31 over Z/2^32, row-major), and its
// application to a stream-order window.  Lets the GPU produce a whole step's draws in parallel
// chunks (gcm_kernels.cuh: rng_chunk_kernel) instead of 3 host rand() calls per inner node.
void rng_jump_matrix(unsigned long long k, uint32_t out[961]);
void rng_apply_jump(const uint32_t m[961], uint32_t w[31]);
// cosf/sinf of the 360 angles create_random_axis can draw, from the host libm.
void rng_teta_tables(float cos_t[360], float sin_t[360]);

// Distinct (la, mu, rho) triples of the local nodes of a mesh set and their identity-basis
// eigensystems: [n_mat][3 stages][9 + 81 + 81] = L, U, U1 for q = e_stage (gcm_inner.cuh: MatTab).
struct MaterialRegistry { std::vector<float> materials, table; std::vector<char> ok; };

struct HostMesh {
	int zone = -1;
	int N = 0, T = 0;
	std::vector<float> coords;        // [N][3]
	std::vector<int> flags;           // node_flags bits (Node.h:4-27)
	std::vector<int> remote_zone, remote_num;
	std::vector<float> mat;           // [N][4] la, mu, rho, yield_limit
	std::vector<float> values;        // [N][9] host mirror (load / snapshot time only)
	std::vector<int> tets;            // [T][4]

	// built by pre_process_mesh
	std::vector<int> n2t_off, n2t;    // node -> tets (`elements`), ascending
	std::vector<int> faces;           // [F][3] border triangles, outward oriented
	std::vector<int> n2f_off, n2f;    // node -> border faces (`border_elements`)
	std::vector<float> tet_hv;        // [T][2] tetr_h, signed volume
	std::vector<int> border_nodes;    // local border nodes, ascending
	std::vector<int> border_slot;     // [N] index into border_nodes or -1
	int n_sharp = 0;                  // the first n_sharp entries of border_launch are the sharp ones
	std::vector<int> border_launch;   // border_nodes permuted for the kernel launch: nodes on sharp
	                                  // edges (many alpha retries) first, so warps stay homogeneous
	std::vector<int> inner_nodes;     // local inner nodes, ascending
	std::vector<float> normals;       // [n_border][3] find_border_node_normal(min_h)
	std::vector<float> basis;         // [n_border][9] local basis of the current step
	std::vector<int> cond_id;         // [n_border] border condition id or -1 (default)
	// static inputs of the specialised inner-node kernel (gcm_inner.cuh), built by
	// build_fast_path_tables(): nodes it may take, their zero-foot weight, the material table
	std::vector<int> inner_fast, inner_generic;   // partition of inner_nodes
	std::vector<float> w0;                        // [N] own-vertex factor of the zero-speed foot
	std::vector<unsigned char> mat_id;            // [N] index into materials (255 = not tabulated)
	MaterialRegistry* registry = nullptr;         // shared by the zones of one set (else own_registry)
	MaterialRegistry own_registry;
	std::vector<int> n2x;                         // [len(n2t)][4] CSR entry -> (other vertices a, b, c in tet order, tet | k << 30)
	void build_fast_path_tables();
	float min_h = -1, max_h = -1;
	float outline_min[3] = {0, 0, 0}, outline_max[3] = {0, 0, 0};
	float current_time = 0;
	bool preprocessed = false;

	bool is_used(int i) const { return (flags[i] & 4) != 0; }
	bool is_local(int i) const { return (flags[i] & 6) == 6; }
	bool is_remote(int i) const { return (flags[i] & 6) == 4; }
	bool is_border(int i) const { return (flags[i] & 8) != 0; }

	// TetrMesh_1stOrder::pre_process_mesh (:66-259). Throws HostError.
	void pre_process_mesh();
	// TetrMesh_1stOrder::find_owner_tetr (:1468-1602), all rings. Returns tet id or -1.
	int find_owner_tetr(int base, const float node_pos[3], float dx, float dy, float dz) const;
	// TetrMesh_1stOrder::find_border_node_normal (:382-479). Throws HostError.
	void find_border_node_normal(int node, float h, float n[3]) const;
	// GCM_..._Rotate_Axis::create_random_axis over all local nodes in index order, as driven
	// by TetrMesh_1stOrder::get_max_possible_tau (:1882-1900). Fills `basis`.
	void create_random_axes(GlibcRand& rng);
	void translate(float dx, float dy, float dz);
};

// Thresholds of the conservative owner pre-filter (MeshView::pf_fail / pf_pass / exact_only) for the
// meshes of one process; see the definition for the error budget behind them.
void prefilter_thresholds(const std::vector<const HostMesh*>& meshes, float* pf_fail, float* pf_pass, int* exact_only);

// Static tables of the border-node kernel (gcm_border.cuh), over the zones of one process in
// local_meshes order: one BndSlotStatic per border slot (merged numbering, laid out as 8 ints), and the
// warp-transposed candidate quads.  The nodes are taken in launch order (border_launch, zone after
// zone) = the order of the device's border list.
struct BorderFlatZone { const HostMesh* hm; size_t node_off, entry_off, slot_off; };
void build_border_flat_tables(const std::vector<BorderFlatZone>& zones, std::vector<int>& slot_static8, std::vector<float>& cand_quads);

// A virtual (contact) node: the point where the ray from a border node along its first basis axis
// meets a border face of another body, with the state interpolated on that face
// (gcm/system/BruteforceCollisionDetector.cpp:44-84).
struct VirtNodeHost {
	float coords[3];
	float values[13];          // 9 state values + la, mu, rho, yield_limit
	int remote_zone;           // zone of the face it lies on (remote_zone_num)
	int remote_face;           // that face's index (remote_num)
	int base_node;             // vert[0] of the face: the node whose `elements` the virt node borrows (absolute_num)
	int real_zone, real_node;  // the border node it was created for
};

// BruteforceCollisionDetector::find_collisions, local/local part (:7-92), over the meshes of one
// process in local_meshes order.  Needs the current bases (hm.basis) and the current values of
// the border nodes (hm.values).  Clears and refills `virt`; axis_plus0[m][slot] = virt index or -1
// for every border slot of mesh m (clear_contact_state + the detector's marks); sets / clears the
// InContact flag.  Throws HostError (the NaN checks of vector_intersects_triangle).
void find_collisions(std::vector<HostMesh*>& meshes, float threshold, std::vector<VirtNodeHost>& virt,
                     std::vector<std::vector<int> >& axis_plus0);

// The static half of collision discovery for the device detector (gcm_collide.cuh): zone-local node / slot / vertex
// numbers, candidates in the reference's (mesh a, mesh b, node) loop order.
struct CollisionPlan {
	std::vector<gcm::CollPair> pairs;
	std::vector<int> pair_a, pair_b;                       // positions in the meshes vector
	std::vector<int> cand_pair, cand_node, cand_slot, cand_mesh;
	std::vector<gcm::CollFace> faces;
	std::vector<int> face_mesh;
	std::vector<int> grid_start, grid_items;
	std::vector<int> mesh_cand_off, mesh_cand_n;           // per mesh a: its contiguous candidate range
};
void build_collision_plan(std::vector<HostMesh*>& meshes, float threshold, CollisionPlan& plan);

// ---- on-disk formats (gcm_io.cpp) ----------------------------------------------------------------
struct MshData {           // what TetrMesh_1stOrder::load_msh_file leaves in nodes / tetrs
	std::vector<float> coords;                       // [3N]; zeros for halo nodes
	std::vector<int> placement, remote_zone, remote_num;   // [N]; 1 = Local; 0-based remote_num
	std::vector<int> tets;                           // [4T], 0-based
};
void read_msh_file(const std::string& path, MshData& out);
void read_zones_map(const std::string& path, std::vector<int>& zone_cpu);
struct VtuData {           // one zone's snapshot: VTKSnapshotWriter::dump_vtk / TetrMesh_1stOrder::load_vtu_file
	std::vector<float> coords, values, material, destruction;   // [3N] [9N] [4N] [8N or empty]
	std::vector<int> flags, tets;                                 // [N] Node::getFlags(), [4T]
};
void write_vtu_file(const std::string& path, const VtuData& d);
void read_vtu_file(const std::string& path, VtuData& d);

} // namespace gcmh
