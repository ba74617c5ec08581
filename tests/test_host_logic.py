"""Host-side logic that needs no GPU: the libc rand() stream restatement, mesh generator."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import util_parity as up
from gcm_b200 import meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RAND_SRC = r'''
#include <cstdio>
#include <cstdlib>
#include "gcm_host.h"
int main() {
    unsigned seeds[] = {1u, 12345u, 0u, 2147483647u, 4000000000u};
    for(unsigned s : seeds) {
        srand(s); gcmh::GlibcRand r; r.seed(s);
        for(int i = 0; i < 100000; i++) { int a = rand(), b = r.next(); if(a != b) { printf("MISMATCH seed %u i %d %d %d\n", s, i, a, b); return 1; } }
    }
    printf("OK\n"); return 0;
}
'''


def test_glibc_rand_restatement(built):
    """GlibcRand (gcm_host.h) reproduces srand()/rand() of the C library bit for bit."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "r.cpp")
        open(src, "w").write(RAND_SRC)
        exe = os.path.join(d, "r")
        csrc = os.path.join(ROOT, "gcm-necro_b200", "csrc")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-I", csrc, src, os.path.join(csrc, "gcm_host.cpp"), "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout


JUMP_SRC = r'''
#include <cstdio>
#include <cstring>
#include "gcm_host.h"
int main() {
    unsigned long long ks[] = {1ull, 3ull, 31ull, 992ull, 100003ull, 30233088ull};
    for(unsigned long long k : ks) {
        gcmh::GlibcRand a; a.seed(12345); gcmh::GlibcRand b = a;
        for(unsigned long long i = 0; i < k; i++) (void)a.next();
        uint32_t m[961], w[31], wa[31];
        gcmh::rng_jump_matrix(k, m);
        b.get_window(w); gcmh::rng_apply_jump(m, w); b.set_window(w);
        a.get_window(wa);
        if(memcmp(w, wa, sizeof w)) { printf("MISMATCH window k=%llu\n", k); return 1; }
        for(int i = 0; i < 1000; i++) if(a.next() != b.next()) { printf("MISMATCH stream k=%llu\n", k); return 1; }
    }
    printf("OK\n"); return 0;
}
'''


def test_rand_jump_ahead(built):
    """A^k applied to the 31-word window == k calls of rand() (the device generator's host half)."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "j.cpp")
        open(src, "w").write(JUMP_SRC)
        exe = os.path.join(d, "j")
        csrc = os.path.join(ROOT, "gcm-necro_b200", "csrc")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-I", csrc, src, os.path.join(csrc, "gcm_host.cpp"), "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout


def test_create_cube_counts():
    """create_cube.c semantics: 100^3 would give 5 821 794 tets; check the formula on 10^3."""
    z = meshgen.make_cube(10)
    assert z.n_nodes == 1000 and z.n_tets == 6 * 9 ** 3
    zs = meshgen.make_box_zones(6, 6, 8, 2)
    assert [x.n_local for x in zs] == [144, 144]
    assert [x.n_nodes for x in zs] == [180, 180]
    # halo of zone 0 = first local layer of zone 1
    h = np.nonzero(zs[0].placement == 0)[0]
    assert np.array_equal(zs[0].coords[h], zs[1].coords[zs[0].remote_num[h]])
    assert np.all(zs[1].placement[zs[0].remote_num[h]] == 1)


ROWS_SRC = r'''
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "gcm_math.cuh"
int main() {
    srand(7);
    for(int it = 0; it < 2000; it++) {
        float la = 1e3f + (rand() % 100000) * 7e4f, mu = 1e3f + (rand() % 100000) * 1e4f, ro = 1.f + (rand() % 10000) * 780.f;
        float q[3] = {(rand() % 2001 - 1000) / 1000.f, (rand() % 2001 - 1000) / 1000.f, (rand() % 2001 - 1000) / 1000.f};
        if(it < 3) { q[0] = it == 0; q[1] = it == 1; q[2] = it == 2; }
        if(q[0] == 0 && q[1] == 0 && q[2] == 0) continue;
        float L[9], U[81], U1[81], row[9];
        if(gcm::elastic_matrix(la, mu, ro, q[0], q[1], q[2], L, U, U1)) continue;
        gcm::ElasticDirs d;
        if(gcm::elastic_dirs(la, mu, ro, q[0], q[1], q[2], d)) { printf("dirs failed\n"); return 1; }
        for(int i = 0; i < 9; i++) {
            float li = gcm::elastic_L(d, i);
            if(memcmp(&li, &L[i], 4) && !(li == 0 && L[i] == 0)) { printf("L mismatch %d\n", i); return 1; }
            gcm::elastic_U_row(d, la, mu, ro, i, row);
            for(int j = 0; j < 9; j++) if(row[j] != U[i*9+j]) { printf("U mismatch %d %d\n", i, j); return 1; }
            gcm::elastic_U1_row(d, ro, i, row);
            for(int j = 0; j < 9; j++) if(row[j] != U1[i*9+j]) { printf("U1 mismatch %d %d\n", i, j); return 1; }
        }
    }
    printf("OK\n"); return 0;
}
'''


def test_elastic_row_functions_match_full_matrix(built):
    """elastic_U_row / elastic_U1_row / elastic_L (used by the half-warp border kernel) == elastic_matrix."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "r.cpp")
        open(src, "w").write(ROWS_SRC)
        exe = os.path.join(d, "r")
        csrc = os.path.join(ROOT, "gcm-necro_b200", "csrc")
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", csrc, src, "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout


def test_div6_fma_form_matches_ieee_division(tmp_path):
    """gcm_math.cuh div6: the FMA form equals x / 6.0f wherever the device code uses it (sampled here,
    every 64th mantissa of every exponent; run tests/emul/div6_check with stride 1 for all 2^32)."""
    import subprocess
    src = os.path.join(os.path.dirname(__file__), "emul", "div6_check.c")
    exe = str(tmp_path / "div6_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-mfma", src, "-o", exe, "-lm"])
    out = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "in exponents 2..254: 0" in out.stdout
    assert "0x3e2aaaab" in out.stdout


def test_host_only_sets_refuse_device_work_and_accept_host_calls(built):
    """A set created with device -1 does host work (load, preprocess, conditions, file formats) and refuses
    everything that would need the GPU with CONFIG_EXCEPTION: there is no CPU execution path to fall back to."""
    import gcm_b200
    from gcm_b200 import meshgen
    z = meshgen.make_cube(5)
    ms = gcm_b200.TetrMeshSet(n_zones=1, device=-1)
    m = ms.get_mesh_by_zone_num(0)
    m.load(z, la=7e9, mu=1e9, rho=7.8e6)
    ms.pre_process_meshes()
    # per-node border conditions: ids from define, unknown ids refused
    cid = ms.define_border_condition(gcm_b200.BC_EXT_FORCE, params=(-1e8, 3e7, 1.0, 0.5, 0.0), start=0.0, duration=0.003)
    assert cid == 0
    cid2 = ms.define_border_condition(gcm_b200.BC_FIXED)
    assert cid2 == 1
    ids = np.full(z.n_nodes, -1, np.int32)
    ids[z.coords[:, 2] > 3.5] = cid
    ids[z.coords[:, 2] < 0.5] = cid2
    m.assign_border_conditions(ids)
    ids[0] = 7
    with pytest.raises(gcm_b200.GCMException) as e:
        m.assign_border_conditions(ids)
    assert e.value.code == gcm_b200.GCMException.CONFIG_EXCEPTION and "unknown border condition id" in e.value.message
    for call in (ms.do_next_step, ms.begin_step, ms.join_io, lambda: ms.track_destruction(True), m.get_destruction_criterias,
                 lambda: ms.do_next_part_step(0.002, 0), ms.synchronize):
        with pytest.raises(gcm_b200.GCMException) as e:
            call()
        assert e.value.code == gcm_b200.GCMException.CONFIG_EXCEPTION
        assert "no CPU execution path" in e.value.message or "host-only" in e.value.message
    ms.close()
