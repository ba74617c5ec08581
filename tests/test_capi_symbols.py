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
