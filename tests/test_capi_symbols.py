"""The C-ABI library loads here (no GPU) and exports every symbol include/gcm_b200.h declares."""
import ctypes
import os
import re

import pytest

import gcm_b200
from gcm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gcm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gcm_b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 30
    assert sorted(capi.SIGNATURES) == names


def test_library_exports_every_symbol(built):
    lib = ctypes.CDLL(gcm_b200.library_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib2 = gcm_b200.load_library()
    assert b"sm_100a" in lib2.gcm_b200_version()


def test_no_cpu_fallback(built):
    """Without a CUDA device the compute entry points fail loudly (CONFIG_EXCEPTION)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gcm_b200.GCMException) as e:
        gcm_b200.TetrMeshSet(n_zones=1)
    assert e.value.code == gcm_b200.GCMException.CONFIG_EXCEPTION
    assert "no CPU execution path" in e.value.message


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(gcm_b200.GCMException):
        capi.load_library(str(tmp_path / "libgcm_b200.so"))


def test_comm_entry_points_refuse_host_only_sets(built):
    """The library-driven exchange is a device feature: a host-only set (device -1, what the gloo tests
    use for planning) cannot get a communicator, cannot exchange, and a multi-process step without one
    says what to do instead of running something else."""
    ms = gcm_b200.TetrMeshSet(n_zones=1, device=-1)
    z = __import__("gcm_b200.meshgen", fromlist=["make_cube"]).make_cube(4)
    ms.get_mesh_by_zone_num(0).load(z, la=7e9, mu=1e9, rho=7.8e6)
    ms.pre_process_meshes()
    for call in (lambda: ms.comm_init(b"\0" * 128, 0, 1), ms.comm_sync_nodes, ms.do_next_step):
        with pytest.raises(gcm_b200.GCMException) as e:
            call()
        assert e.value.code == gcm_b200.GCMException.CONFIG_EXCEPTION
    ms.close()


@pytest.mark.parametrize("compiler,flags", [("gcc", ["-std=c99", "-pedantic", "-x", "c"]), ("g++", ["-std=gnu++98", "-x", "c++"])])
def test_header_is_plain_c_and_cxx98(tmp_path, compiler, flags):
    """The boundary is a C ABI: include/gcm_b200.h must compile, warning-free, as C99 and as the C++98 the
    reference is written in (gnu++98, the dialect oracle/build_ref.sh compiles its sources with; `long long`
    in two inspection calls is the only thing beyond ISO C++98)."""
    import subprocess
    src = tmp_path / "hdr_check.c"
    src.write_text('#include "gcm_b200.h"\nint main(void) { return gcm_b200_version() != 0 ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run([compiler] + flags + ["-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
