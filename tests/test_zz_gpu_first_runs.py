"""GPU parity cases that have only been verified on the CPU emulation so far: they run
last and are non-strict xfail, so that neither a mismatch nor a crash of a first run can hide the result of
the verified suite.  A pass shows as XPASS; move the case to util_parity.CASE_NAMES then."""
import pytest

import util_parity as up

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="first GPU run of a two-material mesh: the mat_id path of the inner kernel "
                                        "is verified on the CPU emulation only (DESIGN.md section 7)")
def test_gpu_two_materials(built):
    """Two (lambda, mu, rho) in one body: per-material eigensystem tables, feet interpolating across the
    material interface.  Not yet run on a GPU when it was written, hence non-strict xfail: a pass shows
    as XPASS, a mismatch does not stop the suite."""
    case = up.make_case("twomat8")
    ref = up.load_golden("twomat8")
    got = up.run_product(case)
    assert up.compare(ref, got) == []


@pytest.mark.xfail(strict=False, reason="first GPU run of an unstructured (Delaunay) mesh: verified on the CPU "
                                        "emulation only when written")
def test_gpu_unstructured_delaunay_mesh(built):
    """scipy-Delaunay mesh of a jittered lattice (variant ii): node degrees 1..36, so
    some inner nodes exceed the 32 candidates of the angular index and take the deferred literal routine."""
    pytest.importorskip("scipy")
    case = up.make_case("delaunay7")
    ref = up.load_golden("delaunay7")
    got = up.run_product(case)
    assert up.compare(ref, got) == []


@pytest.mark.xfail(strict=False, reason="first GPU run of contact between unstructured meshes")
@pytest.mark.parametrize("name", ["delaunay_contact", "delaunay_friction"])
def test_gpu_unstructured_contact(built, name):
    """Two Delaunay bodies 0.01 apart (non-matching surface triangles): Bruteforce detector every step,
    virtual nodes through the any-ring search, Adhesion / Friction."""
    pytest.importorskip("scipy")
    case = up.make_case(name)
    ref = up.load_golden(name)
    got = up.run_product(case)
    assert up.compare(ref, got) == []
