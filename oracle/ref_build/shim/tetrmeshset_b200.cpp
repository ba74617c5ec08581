// The INTEGRATION.md shim as real C++: the two bodies a maintainer replaces in the reference so that its
// hot path runs in libgcm_b200.so, linked with the reference's OWN, otherwise unmodified classes and driver
// (oracle/build_ref.sh, target gcm3d_b200).  Test infrastructure: proves the C ABI of include/gcm_b200.h
// is a drop-in for
//     TetrMeshSet::do_next_step()                      gcm/system/TetrMeshSet.cpp:101-184
//     TetrMesh_1stOrder::do_next_part_step(tau, stage) gcm/meshtypes/TetrMesh_1stOrder.cpp:1902-1934
// The reference translation units are compiled with -Ddo_next_step=do_next_step_host
// -Ddo_next_part_step=do_next_part_step_host, so their own bodies stay in the binary under another name
// and the two definitions below are the ones every caller reaches.
//
// What crosses the boundary, and where it comes from in the reference's objects:
//   meshes            nodes[i].coords / placement / remote_zone_num / remote_num / la mu rho yield_limit,
//                     tetrs[i].vert                                          -> gcm_b200_zone_load
//   srand(seed)       gcm/main.cpp:26 (intercepted below)                    -> gcm_b200_set_seed
//   plugin objects    CollisionDetector subclass, is_static(), get_treshold() -> gcm_b200_set_collision_detector
//                     nodes[i].contact_condition->calc (Adhesion / Friction)  -> gcm_b200_set_contact_calculator
//                     nodes[i].border_condition (calc subclass + its parameters, StepPulseForm)
//                                                     -> gcm_b200_set_define_border_condition + per-node ids
//   state             nodes[i].values[0..8]                                  -> gcm_b200_zone_set_values
// and back after every step (the host objects stay authoritative for snapshots): values, contact flags,
// local bases, virtual nodes, destruction criteria, current time.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <dlfcn.h>

// The calculators keep their parameters private and have no getters; a maintainer would add accessors
// (or a friend declaration).  The proof build reads them in place.
#define private public
#define protected public
#include "system/TetrMeshSet.h"
#include "system/DataBus.h"
#include "system/VoidCollisionDetector.h"
#include "system/BruteforceCollisionDetector.h"
#include "system/BorderCondition.h"
#include "system/ContactCondition.h"
#include "system/forms/StepPulseForm.h"
#include "methods/border/FreeBorderCalculator.h"
#include "methods/border/FixedBorderCalculator.h"
#include "methods/border/ExternalForceCalculator.h"
#include "methods/border/ExternalVelocityCalculator.h"
#include "methods/border/ExternalValuesCalculator.h"
#include "methods/contact/AdhesionContactCalculator.h"
#include "methods/contact/FrictionContactCalculator.h"
#undef private
#undef protected

#include "gcm_b200.h"

// ---- srand(): the library regenerates libc's rand() stream on the device and needs the seed ----------
static unsigned g_seed = 1;        // libc default when srand() is never called
extern "C" void srand(unsigned seed)
{
	typedef void (*srand_fn)(unsigned);
	static srand_fn real = (srand_fn)dlsym(RTLD_NEXT, "srand");
	g_seed = seed;
	if(real) real(seed);
}

namespace {

void gcm_check(int rc)                                 // C ABI -> GCMException, as main.cpp:171-180 expects
{
	if(rc) throw GCMException(-rc - 1, gcm_b200_last_error());
}

struct B200 {
	gcm_b200_set* set;
	std::vector< std::vector<basis> > bases;           // storage behind nodes[i].local_basis, per local mesh
	B200() : set(NULL) { }
};
B200& b200() { static B200 g; return g; }

// BorderCalculator subclass -> calc_type of gcm_b200_set_define_border_condition, parameters in the ABI's layout
int describe(BorderCalculator* c, float params[5], int vars[3], float vals[3])
{
	for(int k = 0; k < 5; k++) params[k] = 0;
	for(int k = 0; k < 3; k++) { vars[k] = 0; vals[k] = 0; }
	if(dynamic_cast<FreeBorderCalculator*>(c)) return 0;
	if(dynamic_cast<FixedBorderCalculator*>(c)) return 1;
	if(ExternalForceCalculator* f = dynamic_cast<ExternalForceCalculator*>(c)) {
		params[0] = f->normal_stress; params[1] = f->tangential_stress;
		for(int k = 0; k < 3; k++) params[2 + k] = f->tangential_direction[k];
		return 2 | GCM_B200_DIR_NORMALISED;             // set_parameters already divided by the norm
	}
	if(ExternalVelocityCalculator* v = dynamic_cast<ExternalVelocityCalculator*>(c)) {
		params[0] = v->normal_v; params[1] = v->tangential_v;
		for(int k = 0; k < 3; k++) params[2 + k] = v->tangential_direction[k];
		return 3 | GCM_B200_DIR_NORMALISED;
	}
	if(ExternalValuesCalculator* e = dynamic_cast<ExternalValuesCalculator*>(c)) {
		for(int k = 0; k < 3; k++) { vars[k] = e->vars_index[k]; vals[k] = e->vars_values[k]; }
		return 4;
	}
	throw GCMException(GCMException::CONFIG_EXCEPTION, "border calculator without a device counterpart");
}

} // namespace

// First call: everything the host objects hold goes to the device (TetrMeshSet::init_mesh_container,
// load_msh_file, pre_process_meshes and the TaskPreparator assignments have run on the host by now).
static void b200_attach(TetrMeshSet* ms, vector<TetrMesh_1stOrder*>& local_meshes, CollisionDetector* cd)
{
	B200& g = b200();
	const int device = getenv("GCM_B200_DEVICE") ? atoi(getenv("GCM_B200_DEVICE")) : 0;
	const int nz = ms->get_number_of_meshes();
	gcm_check(gcm_b200_set_create(&g.set, device, nz, NULL, 0));
	gcm_check(gcm_b200_set_seed(g.set, g_seed));
	int contact_calc = 0;
	for(size_t l = 0; l < local_meshes.size(); l++) {
		TetrMesh_1stOrder* m = local_meshes[l];
		const size_t N = m->nodes.size(), T = m->tetrs.size();
		std::vector<float> coords(3 * N), mat(4 * N);
		std::vector<int> placement(N), rz(N), rn(N), tets(4 * T);
		for(size_t i = 0; i < N; i++) {
			ElasticNode& n = m->nodes[i];
			for(int k = 0; k < 3; k++) coords[3 * i + k] = n.coords[k];
			placement[i] = n.isLocal() ? 1 : 0;
			rz[i] = n.remote_zone_num; rn[i] = n.remote_num;
			mat[4 * i] = n.la; mat[4 * i + 1] = n.mu; mat[4 * i + 2] = n.rho; mat[4 * i + 3] = n.yield_limit;
			if(n.contact_condition && dynamic_cast<FrictionContactCalculator*>(n.contact_condition->calc)) contact_calc = 1;
		}
		for(size_t i = 0; i < T; i++)
			for(int k = 0; k < 4; k++) tets[4 * i + k] = m->tetrs[i].vert[k];
		gcm_check(gcm_b200_zone_load(g.set, m->zone_num, (int)N, (int)T, &coords[0], &placement[0], &rz[0], &rn[0], &tets[0], &mat[0]));
	}
	int det = 0;
	if(dynamic_cast<BruteforceCollisionDetector*>(cd)) det = 1;
	else if(!dynamic_cast<VoidCollisionDetector*>(cd)) throw GCMException(GCMException::CONFIG_EXCEPTION, "collision detector without a device counterpart");
	gcm_check(gcm_b200_set_collision_detector(g.set, det, cd->is_static() ? 1 : 0, cd->get_treshold()));
	gcm_check(gcm_b200_set_contact_calculator(g.set, contact_calc));
	gcm_check(gcm_b200_set_preprocess(g.set));
	// BorderCondition objects, numbered in the order the nodes meet them
	std::map<BorderCondition*, int> ids;
	for(size_t l = 0; l < local_meshes.size(); l++) {
		TetrMesh_1stOrder* m = local_meshes[l];
		std::vector<int> cond(m->nodes.size(), -1);
		bool any = false;
		for(size_t i = 0; i < m->nodes.size(); i++) {
			BorderCondition* bc = m->nodes[i].border_condition;
			if(!bc || bc == ms->getDefaultBorderCondition()) continue;
			std::map<BorderCondition*, int>::iterator it = ids.find(bc);
			if(it == ids.end()) {
				float params[5], vals[3]; int vars[3], id = -1;
				const int type = describe(bc->calc, params, vars, vals);
				StepPulseForm* f = dynamic_cast<StepPulseForm*>(bc->form);
				if(!f) throw GCMException(GCMException::CONFIG_EXCEPTION, "pulse form without a device counterpart");
				gcm_check(gcm_b200_set_define_border_condition(g.set, type, params, vars, vals, f->startTime, f->duration, &id));
				it = ids.insert(std::make_pair(bc, id)).first;
			}
			cond[i] = it->second;
			any = true;
		}
		if(any) gcm_check(gcm_b200_zone_assign_border_conditions(g.set, m->zone_num, &cond[0]));
	}
	g.bases.resize(local_meshes.size());
	for(size_t l = 0; l < local_meshes.size(); l++) {
		TetrMesh_1stOrder* m = local_meshes[l];
		std::vector<float> v(9 * m->nodes.size());
		for(size_t i = 0; i < m->nodes.size(); i++)
			for(int k = 0; k < 9; k++) v[9 * i + k] = m->nodes[i].values[k];
		gcm_check(gcm_b200_zone_set_values(g.set, m->zone_num, &v[0]));
		g.bases[l].resize(m->nodes.size());
	}
	gcm_check(gcm_b200_set_track_destruction(g.set, 1));
}

// After a step: what the host side of the reference reads from its objects (snapshot writers, the driver's dumps).
static void b200_download(TetrMeshSet* ms, vector<TetrMesh_1stOrder*>& local_meshes)
{
	B200& g = b200();
	for(size_t l = 0; l < local_meshes.size(); l++) {
		TetrMesh_1stOrder* m = local_meshes[l];
		const size_t N = m->nodes.size();
		std::vector<float> v(9 * N), b(9 * N), d(8 * N);
		std::vector<int> fl(N);
		gcm_check(gcm_b200_zone_get_values(g.set, m->zone_num, &v[0]));
		gcm_check(gcm_b200_zone_get_bases(g.set, m->zone_num, &b[0]));
		gcm_check(gcm_b200_zone_get_flags(g.set, m->zone_num, &fl[0]));
		gcm_check(gcm_b200_zone_get_destruction_criterias(g.set, m->zone_num, &d[0]));
		for(size_t i = 0; i < N; i++) {
			ElasticNode& n = m->nodes[i];
			if(!n.isLocal()) continue;
			for(int k = 0; k < 9; k++) n.values[k] = v[9 * i + k];
			for(int k = 0; k < 8; k++) n.destruction_criterias[k] = d[8 * i + k];
			n.setContactType((fl[i] & 1) ? InContact : Free);
			for(int k = 0; k < 9; k++) g.bases[l][i].ksi[k / 3][k % 3] = b[9 * i + k];
			n.local_basis = &g.bases[l][i];
		}
	}
	const int nv = gcm_b200_set_virt_count(g.set);
	std::vector<float> virt(16 * (size_t)(nv > 0 ? nv : 1));
	if(nv > 0) gcm_check(gcm_b200_set_get_virt(g.set, &virt[0]));
	ms->virt_nodes.resize(nv > 0 ? nv : 0);
	for(int i = 0; i < nv; i++) {
		for(int k = 0; k < 3; k++) ms->virt_nodes[i].coords[k] = virt[16 * i + k];
		for(int k = 0; k < 13; k++) ms->virt_nodes[i].values[k] = virt[16 * i + 3 + k];
	}
}

// gcm/meshtypes/TetrMesh_1stOrder.cpp:1902 -- the method keeps its signature
int TetrMesh_1stOrder::do_next_part_step(float tau, int stage)
{
	gcm_check(gcm_b200_zone_do_next_part_step(b200().set, zone_num, tau, stage));
	return 0;
}

// gcm/system/TetrMeshSet.cpp:101 -- INTEGRATION.md section 2
int TetrMeshSet::do_next_step()
{
	B200& g = b200();
	if(!g.set) b200_attach(this, local_meshes, collision_detector);
	float time_step = 0.002;                              // TetrMeshSet.cpp:105-106
	gcm_check(gcm_b200_set_begin_step(g.set));            // get_max_possible_tau(): new local bases; then
	                                                      // find_collisions unless static && t != 0 (:111-120)
	for(int s = 0; s < 4; s++) {
		data_bus->sync_nodes();                           // cross-rank halo (one rank in this build: host copies only)
		gcm_check(gcm_b200_set_sync_nodes(g.set));        // same-rank halo (DataBus.cpp:64-79)
		for(int l = 0; l < local_meshes.size(); l++)      // TetrMeshSet.cpp:158-167
			if( local_meshes[l]->do_next_part_step(time_step, s) < 0 )
				throw GCMException( GCMException::MESH_EXCEPTION, "Do next part step failed");
	}
	gcm_check(gcm_b200_set_end_step(g.set));              // rheology (void), destruction criteria, time update
	gcm_check(gcm_b200_set_synchronize(g.set));           // device-side GCMExceptions surface here
	b200_download(this, local_meshes);
	for(int i = 0; i < local_meshes.size(); i++)
		local_meshes[i]->update_current_time(time_step);
	return 0;
}
