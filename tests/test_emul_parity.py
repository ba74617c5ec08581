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
