"""CPU checks of the hot-path restatement (gcm_math.cuh / gcm_step.cuh / gcm_host.cpp), which
the CUDA kernels compile verbatim: the test-only emulation harness (tests/emul) against the
golden fixtures produced by the reference's own sources, and against oracle/_ref itself when it
is built here.  Bit-exact: flags, border faces, random bases, owner tetrahedra, alpha tries,
node values."""
import os

import pytest

import util_parity as up


@pytest.mark.parametrize("name", up.CASE_NAMES)
def test_emulation_matches_golden(built, name):
    case = up.make_case(name)
    got = up.run_emul(case)
    assert up.compare(up.load_golden(name), got) == []


@pytest.mark.parametrize("name", ["cube8", "zones3_jitter", "extforce8"])
def test_golden_is_what_the_reference_produces(built, name):
    """Pins the fixtures: regenerate with the reference binary and compare (skipped on the GPU box
    when oracle/_ref did not travel)."""
    if not os.path.exists(up.ORACLE_TAP):
        pytest.skip("oracle/_ref not built")
    ref = up.run_oracle(up.make_case(name))
    assert up.compare(ref, up.load_golden(name)) == []


def test_bigger_cube_against_reference(built):
    """16^3 cube, 5 steps, straight against the reference binary."""
    if not os.path.exists(up.ORACLE_TAP):
        pytest.skip("oracle/_ref not built")
    from gcm_b200 import meshgen
    z = meshgen.make_cube(16)
    case = dict(name="cube16", zones=[z], states=[meshgen.pwave_state(z.coords, up.LA, up.MU, up.RHO)],
                material=(up.LA, up.MU, up.RHO, 1e16), steps=5, seed=777, borders=[], tap_step=4)
    assert up.compare(up.run_oracle(case), up.run_emul(case)) == []


def test_face_grid_gives_the_full_scan_result(built):
    """Collision discovery bins the partner's faces on a 2-D grid (gcm_host.cpp: FaceGrid); the hit face
    and virtual node of every border node must be those of the reference's scan over all faces.
    Two jittered, shifted bodies (non-matching surfaces, every kind of node normal), detector every step:
    emulation with the grid == emulation without == the reference."""
    import os
    from gcm_b200 import meshgen
    za = meshgen.make_box_zones(14, 14, 6, 1, jitter=0.15, jitter_border=True, seed=11)[0]
    zb = meshgen.make_box_zones(10, 10, 5, 1, jitter=0.15, jitter_border=True, seed=12)[0]
    zb.zone = 1
    case = dict(name="grid_ab", zones=[za, zb],
                states=[meshgen.generic_state(za.coords, up.LA, up.MU, up.RHO, seed=3),
                        meshgen.generic_state(zb.coords, up.LA, up.MU, up.RHO, seed=4)],
                material=(up.LA, up.MU, up.RHO, 1e16), steps=2, seed=5, borders=[], tap_step=1,
                translates=[(1, 1.3, 2.1, 5.2)], detector="brute", static=False, contact_calc="adhesion")
    ref = up.run_oracle(case)
    old = os.environ.get("GCM_B200_FACE_GRID")
    try:
        os.environ["GCM_B200_FACE_GRID"] = "1"
        with_grid = up.run_emul(case)
        os.environ["GCM_B200_FACE_GRID"] = "0"
        without = up.run_emul(case)
    finally:
        if old is None:
            os.environ.pop("GCM_B200_FACE_GRID", None)
        else:
            os.environ["GCM_B200_FACE_GRID"] = old
    assert up.compare(ref, with_grid) == []
    assert up.compare(ref, without) == []
    n_contact = sum(int((r["flags"] & 1).sum()) for r in with_grid)
    assert n_contact > 50


def test_any_ring_search_matches_host_restatement(built):
    """gcm::find_owner_general (the list-building ring expansion the deep contact pass runs on the device)
    against HostMesh::find_owner_tetr (the literal restatement of TetrMesh_1stOrder.cpp:1468-1602 used by
    the preprocessing) on random nodes and displacements of up to 5 shortest altitudes -- first to fourth
    ring, found and not found, probes started off the base node as virtual nodes are."""
    import subprocess, tempfile
    from gcm_b200 import meshgen
    z = meshgen.make_box_zones(9, 8, 7, 1, jitter=0.2, jitter_border=True, seed=31)[0]
    with tempfile.TemporaryDirectory() as d:
        mp = os.path.join(d, "z.gcmb")
        meshgen.write_gcmb(mp, z)
        r = subprocess.run([up.EMUL, "--mesh", mp, "--seed", "77", "--ring-check", "4000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tested, found, bad = [int(x) for x in r.stdout.split(":")[1].split()]
    assert bad == 0
    assert tested > 3000 and 0.2 * tested < found < 0.95 * tested


def test_reference_aborts_on_meshes_finer_than_c1_tau(built):
    """Pins what `find_border_cross` means for parity: on a cube of spacing 0.05 (< c1 tau = 0.068) the reference
    reaches TetrMethod_Plastic_1stOrder.cpp:619-643 and aborts (assert in Node::isOwnedBy, Node.h:132, or a smashed
    stack in TetrMesh_1stOrder.cpp:1668-1675), with and without jitter.  There is no reference result for
    such meshes; the product reports them (tests/test_gpu_parity.py)."""
    if not os.path.exists(up.ORACLE_TAP):
        pytest.skip("oracle/_ref not built")
    from gcm_b200 import meshgen
    for jit in (0.0, 0.2):
        z = meshgen.make_cube(8, spacing=(0.05, 0.05, 0.05), jitter=jit, seed=41)
        case = dict(name="fine", zones=[z], states=[meshgen.generic_state(z.coords, up.LA, up.MU, up.RHO, seed=42)],
                    material=(up.LA, up.MU, up.RHO, 1e16), steps=1, seed=78, borders=[], tap_step=-1)
        with pytest.raises(RuntimeError):
            up.run_oracle(case)
