// Driver for the UNMODIFIED reference hot path (oracle/_ref build only -- test infrastructure).
//
// Reproduces gcm/main.cpp's call order without TaskPreparator (TinyXML is absent):
//   srand(seed) -> TetrMeshSet/DataBus singletons -> per-zone mesh from a binary .gcmb file
//   (same per-node initialisation as load_msh_file, TetrMesh_1stOrder.cpp:773-823) or, for a *.msh
//   argument, through the reference's own load_msh_file ->
//   sync_nodes -> pre_process_meshes -> initial state -> N x TetrMeshSet::do_next_step().
// It dumps what the parity tests compare against: final values[9] per node, node flags,
// border faces, border-node bases of the last step, and (tap build) one record per owner
// search of a chosen step.  It also prints one JSON line with the step-loop wall time so
// bench.py can use it as the `reference` CPU arm.
//
// File formats (little endian):
//   .gcmb mesh: int32 magic 'GCMB', int32 version(1), int32 zone, int32 N, int32 T,
//               int32 placement[N] (1 local, 0 remote), float32 coords[3N],
//               int32 remote_zone[N], int32 remote_num[N] (0-based), int32 tets[4T] (0-based)
//   state:      float32 values[9N]        material: float32 (la,mu,rho,yield)[4N]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "system/Logger.h"
#include "system/TetrMeshSet.h"
#include "system/DataBus.h"
#include "system/VoidCollisionDetector.h"
#include "system/BruteforceCollisionDetector.h"
#include "system/BorderCondition.h"
#include "system/ContactCondition.h"
#include "system/forms/StepPulseForm.h"
#include "system/areas/BoxArea.h"
#include "rheotypes/VoidRheologyCalculator.h"
#include "methods/TetrMethod_Plastic_1stOrder.h"
#include "methods/border/FreeBorderCalculator.h"
#include "methods/border/FixedBorderCalculator.h"
#include "methods/border/ExternalForceCalculator.h"
#include "methods/border/ExternalVelocityCalculator.h"
#include "methods/border/ExternalValuesCalculator.h"
#include "methods/contact/AdhesionContactCalculator.h"
#include "methods/contact/FrictionContactCalculator.h"

using std::string;
using std::vector;

static void die(const char* msg) { fprintf(stderr, "gcm3d_ref: %s\n", msg); exit(2); }

template<class T> static void read_vec(FILE* f, vector<T>& v, size_t n)
{
	v.resize(n);
	if(n && fread(&v[0], sizeof(T), n, f) != n) die("short read");
}

template<class T> static void write_file(const string& name, const vector<T>& v)
{
	FILE* f = fopen(name.c_str(), "wb");
	if(!f) die("cannot open output file");
	if(v.size()) fwrite(&v[0], sizeof(T), v.size(), f);
	fclose(f);
}

// ---- tap (only linked into the tap build; see ref.mk) ---------------------------------
struct TapRec { int step, zone, node, is_virt, stage, ch, count, owner; float alpha, dx[3]; };
static vector<TapRec> g_tap;
static int g_tap_step = -1, g_cur_step = 0;
#ifdef GCM_ORACLE_TAP
void gcm_oracle_tap(ElasticNode* cur_node, TetrMesh* mesh, int stage, int ch, int count, float alpha, float* dx, Tetrahedron* t)
{
	if(g_cur_step != g_tap_step) return;
	TapRec r;
	r.step = g_cur_step; r.zone = mesh->zone_num; r.node = cur_node->local_num;
	r.is_virt = (cur_node != &mesh->nodes[cur_node->local_num]);
	r.stage = stage; r.ch = ch; r.count = count; r.owner = t ? t->local_num : -1;
	r.alpha = alpha; r.dx[0] = dx[0]; r.dx[1] = dx[1]; r.dx[2] = dx[2];
	g_tap.push_back(r);
}
#endif

struct CondSpec { string type; float p[5]; int vars[3]; float start, dur; float box[6]; };

int main(int argc, char** argv)
{
	vector<string> mesh_files, state_files, mat_files;
	vector<CondSpec> conds;
	vector< vector<float> > translates; // zone dx dy dz
	float la = 7e9f, mu = 1e9f, rho = 7.8e6f, yield = 1e16f;
	int steps = 1, seed = 12345, warmup = 0;
	string detector = "void", out, contact_calc = "adhesion";
	bool detector_static = false;

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
		else if(s == "--warmup") warmup = atoi(argv[++a]);   // untimed leading steps (bench.py)
		else if(s == "--detector") detector = argv[++a];
		else if(s == "--static") detector_static = true;
		else if(s == "--contact-calc") contact_calc = argv[++a];
		else if(s == "--out") out = argv[++a];
		else if(s == "--tap-step") g_tap_step = atoi(argv[++a]);
		else if(s == "--translate") { vector<float> t; for(int k = 0; k < 4; k++) t.push_back(atof(argv[++a])); translates.push_back(t); }
		else if(s == "--border") {
			// --border <type> p0 p1 p2 p3 p4 start dur minX maxX minY maxY minZ maxZ
			CondSpec c; c.type = argv[++a];
			for(int k = 0; k < 5; k++) c.p[k] = atof(argv[++a]);
			c.start = atof(argv[++a]); c.dur = atof(argv[++a]);
			for(int k = 0; k < 6; k++) c.box[k] = atof(argv[++a]);
			conds.push_back(c);
		}
		else die(("unknown argument " + s).c_str());
	}
	if(mesh_files.empty()) die("no --mesh given");
	const int Z = mesh_files.size();

	srand(seed); // gcm/main.cpp:26 seeds with time(0); the harness pins it

	Logger* logger = Logger::getInstace();
	logger->setFileOutput("/dev/null");
	TetrMeshSet* ms = TetrMeshSet::getInstance();
	DataBus* db = DataBus::getInstance();
	ms->init();
	logger->init();

	vector<int> zones(Z, 0); // every zone on cpu 0
	db->load_zones_info(&zones);
	ms->init_mesh_container(zones);

	ms->attach(new VoidRheologyCalculator());
	CollisionDetector* cd;
	if(detector == "void") cd = new VoidCollisionDetector();
	else if(detector == "brute") cd = new BruteforceCollisionDetector();
	else die("unknown detector");
	cd->set_treshold(1); // TaskPreparator.cpp:241-254
	cd->set_static(detector_static);
	ms->attach(cd);
	db->attach(cd);

	for(int z = 0; z < Z; z++) {
		if(mesh_files[z].size() > 4 && mesh_files[z].substr(mesh_files[z].size() - 4) == ".msh") {
			// the reference's own reader (TetrMesh_1stOrder::load_msh_file, :735-866), then what TaskPreparator
			// does with a loaded mesh: rheology for every node (TaskPreparator.cpp:335-386), method, attach
			TetrMesh_1stOrder* m = ms->get_mesh_by_zone_num(z);
			m->load_msh_file(const_cast<char*>(mesh_files[z].c_str()));
			for(size_t i = 0; i < m->nodes.size(); i++) {
				m->nodes[i].la = la; m->nodes[i].mu = mu; m->nodes[i].rho = rho; m->nodes[i].yield_limit = yield;
			}
			m->attach(new GCM_Tetr_Plastic_Interpolation_1stOrder_Rotate_Axis());
			ms->attach(m);
			continue;
		}
		FILE* f = fopen(mesh_files[z].c_str(), "rb");
		if(!f) die("cannot open mesh file");
		int hdr[5];
		if(fread(hdr, 4, 5, f) != 5 || hdr[0] != 0x424D4347 || hdr[1] != 1) die("bad mesh header");
		int zone = hdr[2], N = hdr[3], T = hdr[4];
		if(zone != z) die("zone files must be given in zone order");
		vector<int> placement, rz, rn, tets; vector<float> coords;
		read_vec(f, placement, N); read_vec(f, coords, 3 * (size_t)N);
		read_vec(f, rz, N); read_vec(f, rn, N); read_vec(f, tets, 4 * (size_t)T);
		fclose(f);
		vector<float> mat;
		if((int)mat_files.size() > z) {
			FILE* g = fopen(mat_files[z].c_str(), "rb"); if(!g) die("cannot open material file");
			read_vec(g, mat, 4 * (size_t)N); fclose(g);
		}

		TetrMesh_1stOrder* m = ms->get_mesh_by_zone_num(z);
		for(int i = 0; i < N; i++) {
			ElasticNode nn; // same initialisation as load_msh_file
			nn.remote_num = -1; nn.remote_zone_num = -1;
			nn.local_num = nn.absolute_num = i;
			nn.local_zone_num = z;
			for(int k = 0; k < 3; k++) nn.coords[k] = nn.fixed_coords[k] = 0;
			for(int k = 0; k < 13; k++) nn.values[k] = 0;
			nn.elements = NULL; nn.border_elements = NULL; nn.contact_data = NULL; nn.local_basis = NULL;
			nn.setIsBorder(false); nn.addOwner(GCM); nn.setContactType(Free); nn.setUsed(true);
			nn.volume_calculator = NULL; nn.border_condition = NULL; nn.contact_condition = NULL;
			for(int k = 0; k < 8; k++) nn.destruction_criterias[k] = 0;
			if(placement[i]) {
				for(int k = 0; k < 3; k++) nn.coords[k] = coords[3 * (size_t)i + k];
				nn.setPlacement(Local);
			} else {
				nn.remote_zone_num = rz[i]; nn.remote_num = rn[i];
				nn.setPlacement(Remote);
			}
			if(mat.size()) { nn.la = mat[4*(size_t)i]; nn.mu = mat[4*(size_t)i+1]; nn.rho = mat[4*(size_t)i+2]; nn.yield_limit = mat[4*(size_t)i+3]; }
			else { nn.la = la; nn.mu = mu; nn.rho = rho; nn.yield_limit = yield; }
			m->add_node(&nn);
		}
		for(int i = 0; i < T; i++) {
			Tetrahedron_1st_order t;
			t.local_num = i; t.absolute_num = i;
			for(int k = 0; k < 4; k++) t.vert[k] = tets[4 * (size_t)i + k];
			m->add_tetr(&t);
		}
		m->attach(new GCM_Tetr_Plastic_Interpolation_1stOrder_Rotate_Axis());
		ms->attach(m);
	}
	for(size_t t = 0; t < translates.size(); t++)
		ms->get_mesh_by_zone_num((int)translates[t][0])->translate(translates[t][1], translates[t][2], translates[t][3]);

	int rc = 0;
	double seconds = 0;
	long local_nodes = 0;
	try {
		db->create_custom_types();
		db->sync_nodes();
		ms->pre_process_meshes();

		// border conditions, as TaskPreparator::load_task does (TaskPreparator.cpp:388-441)
		for(size_t c = 0; c < conds.size(); c++) {
			BorderCondition* bc = new BorderCondition();
			const CondSpec& s = conds[c];
			if(s.type == "extforce") { ExternalForceCalculator* k = new ExternalForceCalculator(); k->set_parameters(s.p[0], s.p[1], s.p[2], s.p[3], s.p[4]); bc->calc = k; }
			else if(s.type == "extvel") { ExternalVelocityCalculator* k = new ExternalVelocityCalculator(); k->set_parameters(s.p[0], s.p[1], s.p[2], s.p[3], s.p[4]); bc->calc = k; }
			else if(s.type == "fixed") bc->calc = new FixedBorderCalculator();
			else if(s.type == "free") bc->calc = new FreeBorderCalculator();
			else if(s.type == "extvalues") {
				// p0 = the three variable indices as decimal digits (i0 i1 i2), p1..p3 = their values
				ExternalValuesCalculator* k = new ExternalValuesCalculator();
				const int code = (int)s.p[0];
				int vars[3] = {code / 100, (code / 10) % 10, code % 10};
				float vals[3] = {s.p[1], s.p[2], s.p[3]};
				k->set_parameters(vars, vals);
				bc->calc = k;
			}
			else die("unknown border calculator");
			bc->form = new StepPulseForm(s.start, s.dur);
			bc->area = new BoxArea(s.box[0], s.box[1], s.box[2], s.box[3], s.box[4], s.box[5]);
			for(int l = 0; l < ms->get_number_of_local_meshes(); l++) {
				vector<ElasticNode>& nodes = ms->get_local_mesh(l)->nodes;
				for(size_t i = 0; i < nodes.size(); i++)
					if(bc->area->isInArea(&nodes[i])) nodes[i].border_condition = bc;
			}
		}
		if(contact_calc == "friction") {
			ContactCondition* cc = new ContactCondition();
			cc->calc = new FrictionContactCalculator();
			cc->form = new StepPulseForm(-1, -1);
			cc->area = NULL;
			for(int l = 0; l < ms->get_number_of_local_meshes(); l++) {
				vector<ElasticNode>& nodes = ms->get_local_mesh(l)->nodes;
				for(size_t i = 0; i < nodes.size(); i++) nodes[i].contact_condition = cc;
			}
		}

		// initial state straight into nodes[i].values
		for(int z = 0; z < Z && z < (int)state_files.size(); z++) {
			TetrMesh_1stOrder* m = ms->get_mesh_by_zone_num(z);
			FILE* f = fopen(state_files[z].c_str(), "rb"); if(!f) die("cannot open state file");
			vector<float> st; read_vec(f, st, 9 * m->nodes.size()); fclose(f);
			for(size_t i = 0; i < m->nodes.size(); i++)
				if(m->nodes[i].isLocal())
					for(int k = 0; k < 9; k++) m->nodes[i].values[k] = st[9 * i + k];
		}
		for(int z = 0; z < Z; z++) {
			TetrMesh_1stOrder* m = ms->get_mesh_by_zone_num(z);
			for(size_t i = 0; i < m->nodes.size(); i++) if(m->nodes[i].isLocal()) local_nodes++;
		}

		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for(g_cur_step = 0; g_cur_step < steps; g_cur_step++) {
			if(g_cur_step == warmup) clock_gettime(CLOCK_MONOTONIC, &t0);
			if(ms->do_next_step() < 0) { rc = 3; break; }
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	} catch(GCMException& e) {
		fprintf(stderr, "gcm3d_ref: GCMException %d: %s\n", e.getCode(), e.getMessage().c_str());
		rc = 10 + e.getCode();
	}

	if(out.size()) {
		for(int z = 0; z < Z; z++) {
			TetrMesh_1stOrder* m = ms->get_mesh_by_zone_num(z);
			const size_t N = m->nodes.size();
			vector<float> vals(9 * N), crd(3 * N), bases(9 * N, 0.f), destr(8 * N);
			vector<int> flags(N), faces(3 * m->border.size()), contact(6 * N, -1);
			for(size_t i = 0; i < N; i++) {
				ElasticNode& n = m->nodes[i];
				for(int k = 0; k < 9; k++) vals[9 * i + k] = n.values[k];
				for(int k = 0; k < 3; k++) crd[3 * i + k] = n.coords[k];
				for(int k = 0; k < 8; k++) destr[8 * i + k] = n.destruction_criterias[k];
				flags[i] = (n.isInContact() ? 1 : 0) | (n.isLocal() ? 2 : 0) | (n.isUsed() ? 4 : 0) | (n.isBorder() ? 8 : 0);
				if(n.local_basis) for(int k = 0; k < 9; k++) bases[9 * i + k] = n.local_basis->ksi[k / 3][k % 3];
				if(n.contact_data) for(int k = 0; k < 3; k++) { contact[6*i+k] = n.contact_data->axis_plus[k]; contact[6*i+3+k] = n.contact_data->axis_minus[k]; }
			}
			for(size_t i = 0; i < m->border.size(); i++)
				for(int k = 0; k < 3; k++) faces[3 * i + k] = m->border[i].vert[k];
			char zs[32]; snprintf(zs, sizeof zs, ".z%d", z);
			write_file(out + zs + ".values.f32", vals);
			write_file(out + zs + ".coords.f32", crd);
			write_file(out + zs + ".destr.f32", destr);
			write_file(out + zs + ".flags.i32", flags);
			write_file(out + zs + ".faces.i32", faces);
			write_file(out + zs + ".bases.f32", bases);
			write_file(out + zs + ".contact.i32", contact);
		}
		vector<float> virt(16 * ms->virt_nodes.size());
		for(size_t i = 0; i < ms->virt_nodes.size(); i++) {
			for(int k = 0; k < 3; k++) virt[16*i+k] = ms->virt_nodes[i].coords[k];
			for(int k = 0; k < 13; k++) virt[16*i+3+k] = ms->virt_nodes[i].values[k];
		}
		write_file(out + ".virt.f32", virt);
		if(g_tap_step >= 0) write_file(out + ".tap.bin", g_tap);
	}
	printf("{\"impl\": \"reference\", \"steps\": %d, \"timed_steps\": %d, \"seconds\": %.6f, \"local_nodes\": %ld, \"virt_nodes\": %ld, \"rc\": %d}\n",
		steps, steps - warmup, seconds, local_nodes, (long)ms->virt_nodes.size(), rc);
	return rc;
}
