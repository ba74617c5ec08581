// gcm_io.cpp -- on-disk formats either side of the hot path (host C++, no third-party libraries):
//   * gmsh-2 ASCII .msh as TetrMesh_1stOrder::load_msh_file reads it (gcm/meshtypes/TetrMesh_1stOrder.cpp:735-866),
//     including the negative-id halo lines the zone splitter (gcm/tools/create_cube.c:38-81) writes;
//   * the zones map XML of TaskPreparator::load_zones_info (gcm/system/TaskPreparator.cpp:65-99; TinyXML there);
//   * .vtu snapshots with the point arrays of VTKSnapshotWriter::dump_vtk (gcm/system/VTKSnapshotWriter.cpp:49-206;
//     VTK's XML writer there) and the restart reader TetrMesh_1stOrder::load_vtu_file (:868-949) for the ASCII
//     flavour of that file.
// The reference leans on iostream extraction, TinyXML and VTK; the subset each call site accepts is restated here.
#include "gcm_host.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace gcmh {

// ---- gmsh-2 .msh ---------------------------------------------------------------------------------------
// Token stream exactly as the reference consumes it with operator>>: "$MeshFormat" v t s "$EndMeshFormat"
// "$Nodes" n, then per node  id x y z  (id > 0: local)  or  -id remote_zone remote_id  (halo; ids 1-based;
// coordinates arrive with the first sync_nodes), "$EndNodes" "$Elements" m, then per element  num type ...;
// only type 4 is read:  num 4 ntags t1 t2 v0 v1 v2 v3  (exactly THREE integers between the type and the vertices,
// i.e. two tags, :843-845), anything else is skipped to the end of its line.
void read_msh_file(const std::string& path, MshData& out)
{
	std::ifstream in(path.c_str());
	if(!in.is_open()) throw HostError{MESH_EXCEPTION, "Can not open msh file"};
	auto expect = [&](const char* tag) {
		std::string s;
		in >> s;
		if(s != tag) throw HostError{MESH_EXCEPTION, "Wrong file format"};
	};
	expect("$MeshFormat");
	float f;
	in >> f >> f >> f;
	expect("$EndMeshFormat");
	expect("$Nodes");
	int n_nodes = 0;
	in >> n_nodes;
	if(!in || n_nodes < 0) throw HostError{MESH_EXCEPTION, "Wrong file format"};
	out.coords.assign(3 * (size_t)n_nodes, 0.f);
	out.placement.assign(n_nodes, 1);
	out.remote_zone.assign(n_nodes, -1);
	out.remote_num.assign(n_nodes, -1);
	for(int i = 0; i < n_nodes; i++) {
		int id = 0;
		in >> id;
		if(!in || id == 0) throw HostError{MESH_EXCEPTION, "Wrong file format"};
		// the reference stores the node at position i whatever its id says (nodes.push_back) and uses
		// local_num = |id| - 1 as its index everywhere: the two must agree
		const int num = (id > 0 ? id : -id) - 1;
		if(num != i) throw HostError{MESH_EXCEPTION, "Wrong file format (node ids must equal their position)"};
		if(id > 0) {
			in >> out.coords[3 * (size_t)i] >> out.coords[3 * (size_t)i + 1] >> out.coords[3 * (size_t)i + 2];
		} else {
			int rz = -1, rn = 0;
			in >> rz >> rn;
			out.placement[i] = 0;
			out.remote_zone[i] = rz;
			out.remote_num[i] = rn - 1;
		}
		if(!in) throw HostError{MESH_EXCEPTION, "Wrong file format"};
	}
	expect("$EndNodes");
	expect("$Elements");
	int n_elem = 0;
	in >> n_elem;
	if(!in || n_elem < 0) throw HostError{MESH_EXCEPTION, "Wrong file format"};
	out.tets.clear();
	for(int i = 0; i < n_elem; i++) {
		int num = 0, type = 0;
		in >> num >> type;
		if(!in) throw HostError{MESH_EXCEPTION, "Wrong file format"};
		if(type != 4) {
			std::string rest;
			std::getline(in, rest);
			continue;
		}
		int t1, t2, t3, v[4];
		in >> t1 >> t2 >> t3 >> v[0] >> v[1] >> v[2] >> v[3];
		if(!in || v[0] <= 0 || v[1] <= 0 || v[2] <= 0 || v[3] <= 0) throw HostError{MESH_EXCEPTION, "Wrong file format"};
		for(int k = 0; k < 4; k++) {
			if(v[k] > n_nodes) throw HostError{MESH_EXCEPTION, "Wrong file format (vertex beyond the node list)"};
			out.tets.push_back(v[k] - 1);
		}
	}
	expect("$EndElements");
}

// ---- zones map -----------------------------------------------------------------------------------------
// <zones_map> <zone num="0" cpu="0"/> <zone num="1" cpu="0"/> ... </zones_map>; `num` must count up from 0
// (TaskPreparator.cpp:90-91), values go through atoi like there.  Comments and the XML declaration are skipped.
namespace {

struct XmlTag { std::string name; std::vector<std::pair<std::string, std::string> > attr; bool closing = false; };

// next tag at or after pos (comments, declarations and text are skipped); false at the end of the document
bool next_tag(const std::string& s, size_t& pos, XmlTag& tag)
{
	for(;;) {
		size_t lt = s.find('<', pos);
		if(lt == std::string::npos) return false;
		if(s.compare(lt, 4, "<!--") == 0) { size_t e = s.find("-->", lt + 4); if(e == std::string::npos) return false; pos = e + 3; continue; }
		if(lt + 1 < s.size() && (s[lt + 1] == '?' || s[lt + 1] == '!')) { size_t e = s.find('>', lt); if(e == std::string::npos) return false; pos = e + 1; continue; }
		size_t gt = s.find('>', lt);
		if(gt == std::string::npos) return false;
		std::string body = s.substr(lt + 1, gt - lt - 1);
		pos = gt + 1;
		tag = XmlTag();
		size_t i = 0;
		if(!body.empty() && body[0] == '/') { tag.closing = true; i = 1; }
		while(i < body.size() && !isspace((unsigned char)body[i]) && body[i] != '/') tag.name += body[i++];
		while(i < body.size()) {
			while(i < body.size() && (isspace((unsigned char)body[i]) || body[i] == '/')) i++;
			std::string k, v;
			while(i < body.size() && body[i] != '=' && !isspace((unsigned char)body[i])) k += body[i++];
			while(i < body.size() && isspace((unsigned char)body[i])) i++;
			if(i >= body.size() || body[i] != '=') break;
			i++;
			while(i < body.size() && isspace((unsigned char)body[i])) i++;
			if(i >= body.size() || (body[i] != '"' && body[i] != '\'')) break;
			const char q = body[i++];
			while(i < body.size() && body[i] != q) v += body[i++];
			i++;
			tag.attr.push_back(std::make_pair(k, v));
		}
		return true;
	}
}

const std::string* attribute(const XmlTag& t, const char* name)
{
	for(size_t k = 0; k < t.attr.size(); k++) if(t.attr[k].first == name) return &t.attr[k].second;
	return nullptr;
}

std::string slurp(const std::string& path, int exc, const char* msg)
{
	std::ifstream in(path.c_str(), std::ios::binary);
	if(!in.is_open()) throw HostError{exc, msg};
	std::stringstream ss;
	ss << in.rdbuf();
	return ss.str();
}

} // namespace

void read_zones_map(const std::string& path, std::vector<int>& zone_cpu)
{
	const std::string doc = slurp(path, CONFIG_EXCEPTION, "Can not open zone map file");
	size_t pos = 0;
	XmlTag t;
	bool in_map = false;
	zone_cpu.clear();
	while(next_tag(doc, pos, t)) {
		if(!in_map) {
			if(!t.closing && t.name == "zones_map") in_map = true;
			else if(!t.closing) throw HostError{CONFIG_EXCEPTION, "Malformed zone map file"};   // FirstChildElement("zones_map") of the document
			continue;
		}
		if(t.closing && t.name == "zones_map") break;
		if(t.closing || t.name != "zone") continue;     // FirstChildElement / NextSiblingElement("zone") skip other elements
		const std::string* num = attribute(t, "num");
		const std::string* cpu = attribute(t, "cpu");
		if(!num || !cpu) throw HostError{CONFIG_EXCEPTION, "Malformed zone map file"};
		if(atoi(num->c_str()) != (int)zone_cpu.size()) throw HostError{CONFIG_EXCEPTION, "Malformed zone map file"};
		zone_cpu.push_back(atoi(cpu->c_str()));
	}
	if(!in_map || zone_cpu.empty()) throw HostError{CONFIG_EXCEPTION, "Malformed zone map file"};
}

// ---- .vtu ------------------------------------------------------------------------------------------------
// UnstructuredGrid, ASCII data arrays.  Point arrays and their names as VTKSnapshotWriter::dump_vtk sets them:
// velocity (3 components, the active Vectors), sxx sxy sxz syy syz szz lambda mu rho yieldLimit contact flags
// maxCompression maxTension maxShear maxDeviator max*History; all Float64 but `flags` (Int32); cells = the
// tetrahedra (VTK_TETRA = 10).  %.9g / %.17g print a float / double so that reading it back gives the same bits.
void write_vtu_file(const std::string& path, const VtuData& d)
{
	const size_t N = d.coords.size() / 3, T = d.tets.size() / 4;
	if(d.values.size() != 9 * N || d.material.size() != 4 * N || d.flags.size() != N || (!d.destruction.empty() && d.destruction.size() != 8 * N))
		throw HostError{SNAP_EXCEPTION, "inconsistent snapshot arrays"};
	FILE* f = fopen(path.c_str(), "w");
	if(!f) throw HostError{SNAP_EXCEPTION, "Can not open snapshot file for writing"};
	fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <UnstructuredGrid>\n");
	fprintf(f, "    <Piece NumberOfPoints=\"%zu\" NumberOfCells=\"%zu\">\n", N, T);
	fprintf(f, "      <PointData Vectors=\"velocity\">\n");
	fprintf(f, "        <DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n");
	for(size_t i = 0; i < N; i++) fprintf(f, "%.9g %.9g %.9g\n", d.values[9*i], d.values[9*i+1], d.values[9*i+2]);
	fprintf(f, "        </DataArray>\n");
	auto scalar = [&](const char* name, const float* base, size_t stride, size_t off) {
		fprintf(f, "        <DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n", name);
		for(size_t i = 0; i < N; i++) fprintf(f, "%.9g\n", base ? base[stride * i + off] : 0.f);
		fprintf(f, "        </DataArray>\n");
	};
	static const char* sn[6] = {"sxx", "sxy", "sxz", "syy", "syz", "szz"};
	for(int k = 0; k < 6; k++) scalar(sn[k], d.values.data(), 9, 3 + k);
	static const char* mn[4] = {"lambda", "mu", "rho", "yieldLimit"};
	for(int k = 0; k < 4; k++) scalar(mn[k], d.material.data(), 4, k);
	fprintf(f, "        <DataArray type=\"Float64\" Name=\"contact\" format=\"ascii\">\n");
	for(size_t i = 0; i < N; i++) fprintf(f, "%d\n", (d.flags[i] & 1) ? 1 : 0);
	fprintf(f, "        </DataArray>\n");
	fprintf(f, "        <DataArray type=\"Int32\" Name=\"flags\" format=\"ascii\">\n");
	for(size_t i = 0; i < N; i++) fprintf(f, "%d\n", d.flags[i]);
	fprintf(f, "        </DataArray>\n");
	static const char* dn[8] = {"maxCompression", "maxTension", "maxShear", "maxDeviator",
	                            "maxCompressionHistory", "maxTensionHistory", "maxShearHistory", "maxDeviatorHistory"};
	for(int k = 0; k < 8; k++) scalar(dn[k], d.destruction.empty() ? nullptr : d.destruction.data(), 8, k);
	fprintf(f, "      </PointData>\n      <CellData>\n      </CellData>\n      <Points>\n");
	fprintf(f, "        <DataArray type=\"Float32\" Name=\"Points\" NumberOfComponents=\"3\" format=\"ascii\">\n");
	for(size_t i = 0; i < N; i++) fprintf(f, "%.9g %.9g %.9g\n", d.coords[3*i], d.coords[3*i+1], d.coords[3*i+2]);
	fprintf(f, "        </DataArray>\n      </Points>\n      <Cells>\n");
	fprintf(f, "        <DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n");
	for(size_t t = 0; t < T; t++) fprintf(f, "%d %d %d %d\n", d.tets[4*t], d.tets[4*t+1], d.tets[4*t+2], d.tets[4*t+3]);
	fprintf(f, "        </DataArray>\n        <DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n");
	for(size_t t = 0; t < T; t++) fprintf(f, "%zu\n", 4 * (t + 1));
	fprintf(f, "        </DataArray>\n        <DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
	for(size_t t = 0; t < T; t++) fprintf(f, "10\n");
	fprintf(f, "        </DataArray>\n      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n");
	if(fclose(f) != 0) throw HostError{SNAP_EXCEPTION, "writing the snapshot file failed"};
}

// load_vtu_file: points, velocity, the six stresses, lambda mu rho yieldLimit, flags, tetrahedra.  ASCII arrays
// only (what write_vtu_file and VTK's SetDataModeToAscii produce); binary / appended data is refused.
void read_vtu_file(const std::string& path, VtuData& d)
{
	const std::string doc = slurp(path, MESH_EXCEPTION, "Can not open vtu file");
	size_t pos = 0;
	XmlTag t;
	size_t N = 0, T = 0;
	bool have_piece = false;
	std::vector<double> arr;
	auto parse_numbers = [&](size_t from, size_t to, std::vector<double>& out) {
		out.clear();
		const char* p = doc.c_str() + from;
		const char* end = doc.c_str() + to;
		while(p < end) {
			while(p < end && isspace((unsigned char)*p)) p++;
			if(p >= end) break;
			char* q = nullptr;
			const double v = strtod(p, &q);
			if(q == p) throw HostError{MESH_EXCEPTION, "vtu: malformed number"};
			out.push_back(v);
			p = q;
		}
	};
	std::vector<double> pts, vel, s[6], m[4], fl, conn;
	while(next_tag(doc, pos, t)) {
		if(t.closing) continue;
		if(t.name == "Piece") {
			const std::string* np = attribute(t, "NumberOfPoints");
			const std::string* nc = attribute(t, "NumberOfCells");
			if(!np || !nc) throw HostError{MESH_EXCEPTION, "vtu: Piece without sizes"};
			N = (size_t)atoll(np->c_str()); T = (size_t)atoll(nc->c_str());
			have_piece = true;
		} else if(t.name == "DataArray") {
			const std::string* fmt = attribute(t, "format");
			if(fmt && *fmt != "ascii") throw HostError{UNIMPLEMENTED_EXCEPTION, "vtu: only ascii data arrays are read"};
			const std::string* name = attribute(t, "Name");
			const size_t begin = pos;
			const size_t end = doc.find("</DataArray>", begin);
			if(end == std::string::npos) throw HostError{MESH_EXCEPTION, "vtu: unterminated DataArray"};
			static const char* sn[6] = {"sxx", "sxy", "sxz", "syy", "syz", "szz"};
			static const char* mn[4] = {"lambda", "mu", "rho", "yieldLimit"};
			std::vector<double>* dst = nullptr;
			if(name) {
				if(*name == "Points") dst = &pts;
				else if(*name == "velocity") dst = &vel;
				else if(*name == "flags") dst = &fl;
				else if(*name == "connectivity") dst = &conn;
				for(int k = 0; k < 6; k++) if(*name == sn[k]) dst = &s[k];
				for(int k = 0; k < 4; k++) if(*name == mn[k]) dst = &m[k];
			}
			if(dst) parse_numbers(begin, end, *dst);
			pos = end;
		}
	}
	if(!have_piece || pts.size() != 3 * N || vel.size() != 3 * N || fl.size() != N || conn.size() != 4 * T)
		throw HostError{MESH_EXCEPTION, "vtu: missing or short arrays"};
	for(int k = 0; k < 6; k++) if(s[k].size() != N) throw HostError{MESH_EXCEPTION, "vtu: missing or short arrays"};
	for(int k = 0; k < 4; k++) if(m[k].size() != N) throw HostError{MESH_EXCEPTION, "vtu: missing or short arrays"};
	d.coords.resize(3 * N); d.values.resize(9 * N); d.material.resize(4 * N); d.flags.resize(N); d.tets.resize(4 * T);
	d.destruction.clear();
	for(size_t i = 0; i < N; i++) {
		for(int k = 0; k < 3; k++) { d.coords[3*i+k] = (float)pts[3*i+k]; d.values[9*i+k] = (float)vel[3*i+k]; }
		for(int k = 0; k < 6; k++) d.values[9*i+3+k] = (float)s[k][i];
		for(int k = 0; k < 4; k++) d.material[4*i+k] = (float)m[k][i];
		d.flags[i] = (int)fl[i];
	}
	for(size_t k = 0; k < 4 * T; k++) {
		if(conn[k] < 0 || conn[k] >= (double)N) throw HostError{MESH_EXCEPTION, "vtu: vertex beyond the point list"};
		d.tets[k] = (int)conn[k];
	}
}

} // namespace gcmh
