// Known-answer driver over the UNMODIFIED reference classes (oracle/_ref build only -- test
// infrastructure).  Prints, one item per line, what SURVEY section 4 lists as the reference's usable
// known-answer material, so that tests/ can pin the product's restatement (gcm_math.cuh) against it:
//   elastic <la> <mu> <ro> <qx> <qy> <qz> selfcheck=<0|1> [simple_eq=<0|1>] L=<9 hex> U=<81 hex> U1=<81 hex>
//       ElasticMatrix3D::prepare_matrix(la, mu, ro, q) (gcm/datatypes/ElasticMatrix3D.cpp:243-519), its
//       self_check() identities (:37-43) and, for an axis-aligned q, the simple-vs-generalised A comparison
//       of gcm/unittests/checkElasticMatrix3D.cpp:11-57
//   pit <dx> <dy> <dz> in=<0|1> h=<hex>
//       TetrMesh_1stOrder::point_in_tetr / tetr_h on the tetrahedron of
//       gcm/unittests/checkFindOwnerTetrAdvancedImpl.cpp:227-264, base node 3
// Floats are printed as their bit patterns.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>

#include "system/Logger.h"
#include "system/GCMException.h"
#include "datatypes/ElasticMatrix3D.h"
#include "meshtypes/TetrMesh_1stOrder.h"

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static void elastic(float la, float mu, float ro, float qx, float qy, float qz)
{
	ElasticMatrix3D m;
	m.prepare_matrix(la, mu, ro, qx, qy, qz);
	int ok = 1;
	try { m.self_check(); } catch(GCMException& e) { ok = 0; }
	printf("elastic %08x %08x %08x %08x %08x %08x selfcheck=%d", bits(la), bits(mu), bits(ro), bits(qx), bits(qy), bits(qz), ok);
	const int axis = (qx == 1 && qy == 0 && qz == 0) ? 0 : (qx == 0 && qy == 1 && qz == 0) ? 1 : (qx == 0 && qy == 0 && qz == 1) ? 2 : -1;
	if(axis >= 0) {
		ElasticMatrix3D s;
		s.prepare_matrix(la, mu, ro, axis);
		printf(" simple_eq=%d", (s.A != m.A) ? 0 : 1);
	}
	printf(" L=");
	for(int i = 0; i < 9; i++) printf("%08x", bits(m.L(i, i)));
	printf(" U=");
	for(int i = 0; i < 9; i++) for(int j = 0; j < 9; j++) printf("%08x", bits(m.U(i, j)));
	printf(" U1=");
	for(int i = 0; i < 9; i++) for(int j = 0; j < 9; j++) printf("%08x", bits(m.U1(i, j)));
	printf("\n");
}

int main()
{
	Logger* logger = Logger::getInstace();
	logger->setFileOutput("/dev/null");
	logger->init();
	// the unit-test material of checkElasticMatrix3D.cpp (self_check passes in float for it) and the
	// benchmark material (for which it does not), axis-aligned and general directions
	const float mats[3][3] = {{1000, 1000, 10}, {7e9f, 1e9f, 7.8e6f}, {3e9f, 2e9f, 2.7e6f}};
	const float qs[8][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.6f, 0.8f, 0}, {-0.4218390584f, 0.7677641511f, 0.4822759330f},
	                        {0.1f, -0.2f, 0.97f}, {2, 0, 0}, {-0.57735026f, -0.57735026f, 0.57735026f}};
	for(int a = 0; a < 3; a++) for(int b = 0; b < 8; b++) elastic(mats[a][0], mats[a][1], mats[a][2], qs[b][0], qs[b][1], qs[b][2]);

	// the tetrahedron of checkFindOwnerTetrAdvancedImpl.cpp:227-264
	TetrMesh_1stOrder mesh;
	const float xyz[4][3] = {{5.3074893951f, 1.6371172667f, -2.3064794540f}, {5.2448787689f, 1.4480520487f, -2.4296388626f},
	                         {5.4918704033f, 1.6011626720f, -2.3315827847f}, {5.4362969398f, 1.4015516043f, -2.4567565918f}};
	for(int i = 0; i < 4; i++) {
		ElasticNode nn;
		nn.local_num = nn.absolute_num = i; nn.local_zone_num = 0; nn.remote_num = -1; nn.remote_zone_num = -1;
		for(int k = 0; k < 3; k++) nn.coords[k] = nn.fixed_coords[k] = xyz[i][k];
		for(int k = 0; k < 13; k++) nn.values[k] = 0;
		nn.elements = NULL; nn.border_elements = NULL; nn.contact_data = NULL; nn.local_basis = NULL;
		nn.setIsBorder(false); nn.addOwner(GCM); nn.setContactType(Free); nn.setUsed(true); nn.setPlacement(Local);
		nn.volume_calculator = NULL; nn.border_condition = NULL; nn.contact_condition = NULL;
		mesh.add_node(&nn);
	}
	Tetrahedron_1st_order t;
	t.local_num = 0; t.absolute_num = 0;
	for(int k = 0; k < 4; k++) t.vert[k] = k;
	mesh.add_tetr(&t);
	const float h = mesh.get_min_h();   // = tetr_h(0): one tetrahedron (TetrMesh_1stOrder.cpp:1040-1065)
	// the unit test's displacement, then a deterministic sweep of directions and lengths around node 3
	const float ksi = -0.4218390584f, eta = 0.7677641511f, dzeta = 0.4822759330f, l = -0.0000101252f;
	float probes[96][3];
	int np = 0;
	probes[np][0] = - l*ksi*3; probes[np][1] = - l*eta*3; probes[np][2] = - l*dzeta*3; np++;
	unsigned s = 12345u;
	while(np < 64) {
		float d[3];
		for(int k = 0; k < 3; k++) { s = s * 1664525u + 1013904223u; d[k] = ((s >> 8) & 0xffff) / 32768.0f - 1.0f; }
		const float len = (np % 4 == 0) ? 0.001f : (np % 4 == 1) ? 0.05f : (np % 4 == 2) ? 0.2f : 1e-5f;
		for(int k = 0; k < 3; k++) probes[np][k] = d[k] * len;
		np++;
	}
	// probes aimed at points of the tetrahedron itself (and slightly beyond its faces) from node 3
	while(np < 96) {
		float w[4], sum = 0;
		for(int k = 0; k < 4; k++) { s = s * 1664525u + 1013904223u; w[k] = ((s >> 8) & 0xffff) / 65536.0f + ((np % 3 == 0 && k == 0) ? -0.05f : 0.0f); sum += w[k]; }
		const float scale = (np % 2) ? 1.0f : 0.002f;      // a full-length displacement, and one short enough to be rescaled
		for(int k = 0; k < 3; k++)
			probes[np][k] = ((w[0] * xyz[0][k] + w[1] * xyz[1][k] + w[2] * xyz[2][k] + w[3] * xyz[3][k]) / sum - xyz[3][k]) * scale;
		np++;
	}
	for(int i = 0; i < np; i++) {
		const bool in = mesh.point_in_tetr(3, probes[i][0], probes[i][1], probes[i][2], (Tetrahedron*)&mesh.tetrs[0]);
		printf("pit %08x %08x %08x in=%d h=%08x\n", bits(probes[i][0]), bits(probes[i][1]), bits(probes[i][2]), in ? 1 : 0, bits(h));
	}
	return 0;
}
