/* gcm_b200.h -- C ABI of libgcm_b200.so: the B200-native grid-characteristic time step of
 * gcm3d (avasyukov/gcm-necro) on unstructured tetrahedral meshes.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own; the seam is the set of
 * C++ entry points its orchestrator calls around the hot loop.  Each function below names the
 * reference interface it replaces (paths relative to the reference root):
 *
 *   gcm_b200_set_*        TetrMeshSet            gcm/system/TetrMeshSet.cpp (singleton :379-383)
 *   gcm_b200_zone_*       TetrMesh_1stOrder      gcm/meshtypes/TetrMesh_1stOrder.cpp
 *   gcm_b200_halo_*       DataBus::sync_nodes    gcm/system/DataBus.cpp:59-122
 *
 * Conventions: every function returns 0 on success or  -(GCMException code + 1)  on error
 * (codes: gcm/system/GCMException.h:19-29; e.g. METHOD_EXCEPTION=3 -> -4, MESH_EXCEPTION=4 -> -5);
 * gcm_b200_last_error() returns the message the reference would have thrown.  No C++
 * exceptions cross this boundary.  Host arrays are borrowed for the duration of a call only;
 * device memory is owned by the library behind the opaque handle.  One host thread drives one
 * handle; handles are independent.  There is no CPU execution path: every compute entry
 * point fails with CONFIG_EXCEPTION when no CUDA device is available.
 */
#ifndef GCM_B200_H
#define GCM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcm_b200_set gcm_b200_set;

/* ---- library / error ------------------------------------------------------------------- */
const char* gcm_b200_version(void);
/* Message of the last failing call on this thread (GCMException::getMessage()). */
const char* gcm_b200_last_error(void);

/* ---- TetrMeshSet ------------------------------------------------------------------------ */
/* TetrMeshSet::init_mesh_container (TetrMeshSet.cpp:350-377) + DataBus::load_zones_info
 * (DataBus.cpp:124-129): n_zones zones, zone_rank[z] = owning process ("cpu" of the zones XML,
 * TaskPreparator.cpp:65-99), my_rank = this process, device = CUDA ordinal used by it. */
int gcm_b200_set_create(gcm_b200_set** out, int device, int n_zones, const int* zone_rank, int my_rank);
int gcm_b200_set_destroy(gcm_b200_set* s);
/* srand(seed) of gcm/main.cpp:26 -- the stream consumed by create_random_axis. */
int gcm_b200_set_seed(gcm_b200_set* s, unsigned seed);
/* Where create_random_axis runs: 1 (default) = on the device (the rand() recurrence jumped and
 * walked in parallel chunks, bit-identical to libc for the given seed); 0 = on the host with the
 * restated libc generator, bases uploaded every step. */
int gcm_b200_set_basis_mode(gcm_b200_set* s, int on_device);
/* 1 (default): inner nodes go through the specialised kernel (identity basis, tabulated
 * eigensystems, pre-filtered owner search); 0: every node through the generic per-node routine
 * (the literal restatement of the reference).  Results are bit-identical; 0 exists for A/B tests. */
int gcm_b200_set_kernel_mode(gcm_b200_set* s, int fast_inner);
/* CollisionDetector selection, TaskPreparator.cpp:225-254: type 0 = Void, 1 = Bruteforce;
 * is_static as <collision_detector static="true">; threshold as set_treshold(). */
int gcm_b200_set_collision_detector(gcm_b200_set* s, int type, int is_static, float threshold);
/* Default ContactCondition calculator (TetrMeshSet.cpp:13,19-21): 0 = Adhesion, 1 = Friction. */
int gcm_b200_set_contact_calculator(gcm_b200_set* s, int type);

/* TetrMesh_1stOrder::add_node / add_tetr for a whole zone (TetrMesh_1stOrder.cpp:51-64, and the
 * per-node initialisation of load_msh_file :773-823).  coords[3N] (ignored for remote nodes
 * unless have_remote_coords), placement[N] (1 = Local, 0 = Remote), remote_zone/remote_num[N]
 * (0-based; -1 for local nodes), tets[4T] (0-based), material[4N] = la, mu, rho, yield_limit. */
int gcm_b200_zone_load(gcm_b200_set* s, int zone, int n_nodes, int n_tets, const float* coords,
                       const int* placement, const int* remote_zone, const int* remote_num,
                       const int* tets, const float* material);
/* TetrMesh_1stOrder::load_msh_file (TetrMesh_1stOrder.cpp:735-866): the gmsh-2 ASCII subset it accepts, with the
 * zone splitter's halo lines (-id remote_zone remote_id).  la/mu/rho/yield_limit: the rheology TaskPreparator
 * assigns to every node of the mesh (TaskPreparator.cpp:335-386 with an unbounded area). */
int gcm_b200_zone_load_msh_file(gcm_b200_set* s, int zone, const char* path, float la, float mu, float rho, float yield_limit);
/* TaskPreparator::load_zones_info (TaskPreparator.cpp:65-99): <zones_map><zone num="i" cpu="r"/>...; *n_zones
 * receives the number of zones, zone_rank[0..min(n, capacity)) their cpu.  Feeds gcm_b200_set_create. */
int gcm_b200_read_zones_map(const char* path, int* zone_rank, int capacity, int* n_zones);
/* VTKSnapshotWriter::dump_vtk (VTKSnapshotWriter.cpp:49-206) for one zone: .vtu UnstructuredGrid (ASCII arrays) with
 * the same point arrays and names (velocity, sxx..szz, lambda, mu, rho, yieldLimit, contact, flags, max*),
 * current device state; and the restart reader TetrMesh_1stOrder::load_vtu_file (:868-949) for such a file
 * (instead of gcm_b200_zone_load; the stored values become the initial state). */
int gcm_b200_zone_dump_vtk(gcm_b200_set* s, int zone, const char* path);
int gcm_b200_zone_load_vtu_file(gcm_b200_set* s, int zone, const char* path);
/* TetrMesh::translate (TetrMesh.cpp:29-45); before preprocess only. */
int gcm_b200_zone_translate(gcm_b200_set* s, int zone, float dx, float dy, float dz);
/* Initial state, written straight into nodes[i].values[0..8]; values[9N] node-major. */
int gcm_b200_zone_set_values(gcm_b200_set* s, int zone, const float* values);
int gcm_b200_zone_get_values(gcm_b200_set* s, int zone, float* values);
/* Stream-ordered variants for callers that keep the state in (pinned) host memory between
 * steps: H2D copy + transpose into the current layer (local nodes only) / transpose + D2H copy.
 * They return before the copy is done; gcm_b200_set_synchronize() waits. */
int gcm_b200_zone_upload_values_async(gcm_b200_set* s, int zone, const float* host_values);
int gcm_b200_zone_download_values_async(gcm_b200_set* s, int zone, float* host_values);
/* The two calls above run their copies on streams of their own (H2D of the next input and D2H of the last result
 * overlap each other and the step: PCIe is full duplex), hand-overs by events.  gcm_b200_set_join_io makes the set's
 * stream (gcm_b200_set_stream) wait for every download issued so far, for callers that time or order work on it. */
int gcm_b200_set_join_io(gcm_b200_set* s);
/* A BorderCondition (calculator + StepPulseForm + BoxArea), TaskPreparator.cpp:388-441 and
 * gcm/methods/border/ (every .cpp there).  calc_type: 0 Free, 1 Fixed, 2 ExternalForce, 3 ExternalVelocity,
 * 4 ExternalValues.  params[5] = set_parameters(normal, tangential, dir x,y,z) for 2/3;
 * for 4: vars[3] / vals[3].  start/duration = StepPulseForm; box[6] = BoxArea(minX,maxX,...).
 * Applies to every local zone; call after gcm_b200_set_preprocess. */
int gcm_b200_set_add_border_condition(gcm_b200_set* s, int calc_type, const float* params,
                                      const int* vars, const float* vals,
                                      float start, float duration, const float* box);
/* The same for a host that keeps the reference's plugin objects and assigns them per node
 * (nodes[i].border_condition = bc, as TaskPreparator.cpp:432-438 does after the area test): define a
 * condition (calculator + StepPulseForm, no area; *id receives its number), then hand every node of a
 * zone its condition id (cond_of_node[N]; -1 = the default Free border of TetrMeshSet.cpp:12,16-18;
 * entries of non-border nodes are ignored).  After gcm_b200_set_preprocess. */
#define GCM_B200_DIR_NORMALISED 0x100  /* or-ed into calc_type 2/3: params[2..4] is already the unit vector
                                        * set_parameters stored (ExternalForceCalculator.cpp:23-26), do not divide again */
int gcm_b200_set_define_border_condition(gcm_b200_set* s, int calc_type, const float* params,
                                         const int* vars, const float* vals, float start, float duration, int* id);
int gcm_b200_zone_assign_border_conditions(gcm_b200_set* s, int zone, const int* cond_of_node);

/* main.cpp:124-130: DataBus::create_custom_types + sync_nodes (halo coordinates/material from
 * the owner zones of this process) + TetrMeshSet::pre_process_meshes, then the SoA upload.
 * Multi-process sets must have received remote-rank halo data first (gcm_b200_halo_* below). */
int gcm_b200_set_preprocess(gcm_b200_set* s);

/* TetrMeshSet::do_next_step (TetrMeshSet.cpp:101-184): bases, collisions, check_stresses,
 * 4 x [sync_nodes, per-zone stage], rheology, time update.  Asynchronous on the set's stream;
 * errors raised by kernels are reported by the next gcm_b200_set_synchronize(). */
int gcm_b200_set_do_next_step(gcm_b200_set* s);
int gcm_b200_set_do_steps(gcm_b200_set* s, int n_steps);
/* Pieces of do_next_step, for callers that interleave their own exchange between stages: */
int gcm_b200_set_begin_step(gcm_b200_set* s);            /* :105-145 */
int gcm_b200_set_sync_nodes(gcm_b200_set* s);            /* DataBus::sync_nodes, same-process part */
/* TetrMesh_1stOrder::do_next_part_step(tau, stage), TetrMesh_1stOrder.cpp:1902-1934 */
int gcm_b200_zone_do_next_part_step(gcm_b200_set* s, int zone, float tau, int stage);
/* The per-mesh loop of one stage for every local zone at once, TetrMeshSet.cpp:158-167: the zones
 * of a process live in one merged device mesh, so this is one launch per kernel class.  The
 * per-zone call above may be used instead, but then every local zone must run a stage before
 * the next gcm_b200_set_sync_nodes (the layer swap happens when the last one has). */
int gcm_b200_set_do_next_part_step(gcm_b200_set* s, float tau, int stage);
int gcm_b200_set_end_step(gcm_b200_set* s);              /* :173-181 */
/* TetrMesh_1stOrder::calc_destruction_criterias (TetrMesh_1stOrder.cpp:1957-1995, called by TetrMeshSet.cpp:177-178
 * at the end of every step; consumed by the snapshot writers only).  The four criteria of the current values are
 * evaluated when they are read; their running maxima need every step, so they are kept only while tracking is on
 * (one extra pass of 56 B per node at the end of a step).  out8[8N] = ElasticNode::destruction_criterias per node:
 * max_compression, max_tension, max_shear, max_deviator, then the four *_history (0 while never tracked). */
int gcm_b200_set_track_destruction(gcm_b200_set* s, int on);
int gcm_b200_zone_get_destruction_criterias(gcm_b200_set* s, int zone, float* out8);
/* 1 (default): the border-node kernel runs on a second CUDA stream next to the inner-node kernel
 * of the same stage (they write disjoint nodes); 0: everything on the one stream. */
int gcm_b200_set_concurrency(gcm_b200_set* s, int on);
/* Wait for the stream and surface device-side errors (NaN guard, outer-count, ...). */
int gcm_b200_set_synchronize(gcm_b200_set* s);
float gcm_b200_set_current_time(gcm_b200_set* s);
/* cudaStream_t the set launches on (for CUDA-event timing by the caller). */
void* gcm_b200_set_stream(gcm_b200_set* s);

/* ---- halo exchange between processes (DataBus::sync_nodes cross-rank part, :81-115) ------ */
/* Number of halo nodes this process needs from / must send to `peer_rank` per stage. */
int gcm_b200_halo_counts(gcm_b200_set* s, int peer_rank, int* n_recv, int* n_send);
/* The (zone, node) pairs requested from peer_rank, 2*n_recv ints: what create_custom_types
 * exchanges (DataBus.cpp:313-326). The peer passes them to gcm_b200_halo_set_sends. */
int gcm_b200_halo_get_requests(gcm_b200_set* s, int peer_rank, int* pairs);
int gcm_b200_halo_set_sends(gcm_b200_set* s, int peer_rank, int n_send, const int* pairs);
/* Start-up exchange on host memory (coords 3 + values 13 per node = 16 floats, MPI_ELNODE). */
int gcm_b200_halo_pack_host(gcm_b200_set* s, int peer_rank, float* buf16);
int gcm_b200_halo_unpack_host(gcm_b200_set* s, int peer_rank, const float* buf16);
/* Per-stage exchange on device memory: 9 floats per node, plane-major [9][n]; dev_buf is a
 * device pointer owned by the caller (e.g. an NCCL send/recv buffer). Runs on the set's stream. */
int gcm_b200_halo_pack(gcm_b200_set* s, int peer_rank, float* dev_buf);
int gcm_b200_halo_unpack(gcm_b200_set* s, int peer_rank, const float* dev_buf);
/* Library-driven exchange: the set owns an NCCL communicator (NCCL is resolved with dlopen, the
 * process's own copy if it has one) and moves the halos itself, so that a whole step is one call.
 * Replaces MPI::COMM_WORLD / MPI_Isend / MPI_Irecv of gcm/system/DataBus.cpp:81-115.
 *   rank 0: gcm_b200_comm_get_unique_id(id) (128 bytes, ncclUniqueId); the host broadcasts it;
 *   every rank, after gcm_b200_halo_set_sends: gcm_b200_set_comm_init(s, id, rank, n_ranks) (collective).
 * gcm_b200_set_comm_sync_nodes = pack + grouped ncclSend/ncclRecv with every neighbour + unpack on the
 * set's stream; gcm_b200_set_do_next_step calls it before each stage once a communicator exists. */
int gcm_b200_comm_get_unique_id(void* id128);
int gcm_b200_set_comm_init(gcm_b200_set* s, const void* id128, int rank, int n_ranks);
int gcm_b200_set_comm_sync_nodes(gcm_b200_set* s);

/* ---- inspection (what VTKSnapshotWriter / a debugger would read) -------------------------- */
int gcm_b200_zone_counts(gcm_b200_set* s, int zone, int* n_nodes, int* n_tets, int* n_faces,
                         int* n_border_nodes, int* n_local_nodes);
int gcm_b200_zone_get_flags(gcm_b200_set* s, int zone, int* flags);           /* node_flags[N] */
int gcm_b200_zone_get_faces(gcm_b200_set* s, int zone, int* faces);           /* [3F] */
int gcm_b200_zone_get_bases(gcm_b200_set* s, int zone, float* bases);         /* [9N], E for inner */
int gcm_b200_zone_get_min_max_h(gcm_b200_set* s, int zone, float* min_h, float* max_h);
/* Owner-tet tap: when enabled, every stage records per node the owner tetrahedron of the 5
 * distinct characteristic feet of the last alpha try, the outer count and the number of tries.
 * owners[stage][N][5], outer[stage][N], tries[stage][N] for the 3 axis stages of the last step. */
int gcm_b200_set_enable_tap(gcm_b200_set* s, int on);
int gcm_b200_zone_get_tap(gcm_b200_set* s, int zone, int* owners, int* outer, int* tries);
/* Kernel launches issued so far by this set (bench.py's gpu_launches). */
long long gcm_b200_set_launch_count(gcm_b200_set* s);
/* Per-kernel CUDA-event timing for roofline accounting (bench.py): while on, every launch of
 * kernel class `which` (0 inner-node stage, 1 border-node stage, 2 plastic corrector, 3 halo copy)
 * is bracketed by events on the set's stream; read returns summed ms, launches and nodes. */
int gcm_b200_set_profile(gcm_b200_set* s, int on);
int gcm_b200_set_profile_read(gcm_b200_set* s, int which, double* total_ms, long long* launches, long long* units);
/* What the specialised kernels cover (tests, bench.py): out[0] inner nodes on the cached path,
 * [1] inner nodes left to the literal routine, [2..4] patterns per axis of the de-duplicated cache,
 * [5..7] 1 where that axis runs from the pattern table, [8] flat border nodes, [9] sharp border nodes,
 * [10] flat border nodes the last border launch deferred to the literal routine, [11] 1 when the
 * conservative owner pre-filter is switched off for this mesh (huge coordinates), [12] 1 when the cross-process
 * halo exchange runs on direct peer stores (else NCCL send/recv, or none), [13] inner nodes whose stencil touches another
 * process's halo node (they wait for the exchange; the others run beside it), [14..15] 0. */
int gcm_b200_set_kernel_stats(gcm_b200_set* s, long long* out16);
/* Virtual (contact) nodes of the last collision pass: n and [16] floats each (coords, values). */
int gcm_b200_set_virt_count(gcm_b200_set* s);
int gcm_b200_set_get_virt(gcm_b200_set* s, float* out16);

#ifdef __cplusplus
}
#endif
#endif /* GCM_B200_H */
