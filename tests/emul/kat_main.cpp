// kat_main.cpp -- TEST-ONLY: the product's restatement (gcm-necro_b200/csrc/gcm_math.cuh, the header the
// CUDA kernels compile) on the known-answer inputs of oracle/ref_build/kat_driver.cpp, printed in the same
// format so that tests/test_known_answers.py can compare the two outputs bit for bit.
#include <cstdio>
#include <cstring>
#include "../../gcm-necro_b200/csrc/gcm_math.cuh"

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static void elastic(float la, float mu, float ro, float qx, float qy, float qz)
{
	float L[9], U[81], U1[81];
	const int e = gcm::elastic_matrix(la, mu, ro, qx, qy, qz, L, U, U1);
	printf("elastic %08x %08x %08x %08x %08x %08x rc=%d L=", bits(la), bits(mu), bits(ro), bits(qx), bits(qy), bits(qz), e);
	for(int i = 0; i < 9; i++) printf("%08x", bits(L[i]));
	printf(" U=");
	for(int i = 0; i < 81; i++) printf("%08x", bits(U[i]));
	printf(" U1=");
	for(int i = 0; i < 81; i++) printf("%08x", bits(U1[i]));
	printf("\n");
}

int main()
{
	const float mats[3][3] = {{1000, 1000, 10}, {7e9f, 1e9f, 7.8e6f}, {3e9f, 2e9f, 2.7e6f}};
	const float qs[8][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.6f, 0.8f, 0}, {-0.4218390584f, 0.7677641511f, 0.4822759330f},
	                        {0.1f, -0.2f, 0.97f}, {2, 0, 0}, {-0.57735026f, -0.57735026f, 0.57735026f}};
	for(int a = 0; a < 3; a++) for(int b = 0; b < 8; b++) elastic(mats[a][0], mats[a][1], mats[a][2], qs[b][0], qs[b][1], qs[b][2]);

	const float xyz[4][3] = {{5.3074893951f, 1.6371172667f, -2.3064794540f}, {5.2448787689f, 1.4480520487f, -2.4296388626f},
	                         {5.4918704033f, 1.6011626720f, -2.3315827847f}, {5.4362969398f, 1.4015516043f, -2.4567565918f}};
	gcm::Vec3 p[4];
	for(int i = 0; i < 4; i++) { p[i].x = xyz[i][0]; p[i].y = xyz[i][1]; p[i].z = xyz[i][2]; }
	const float h = gcm::tetr_h(p[0], p[1], p[2], p[3]);
	const float vol = gcm::tet_signed_volume(p[0], p[1], p[2], p[3]);
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
		const float dx = probes[i][0], dy = probes[i][1], dz = probes[i][2];
		const float delta = sqrtf(dx * dx + dy * dy + dz * dz);
		const bool in = gcm::point_in_tetr(p[3], dx, dy, dz, delta, p[0], p[1], p[2], p[3], h, vol);
		printf("pit %08x %08x %08x in=%d h=%08x\n", bits(dx), bits(dy), bits(dz), in ? 1 : 0, bits(h));
	}
	return 0;
}
