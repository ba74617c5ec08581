import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gcm-necro_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (library, emulation harness, oracle when possible)."""
    import util_parity as up
    import gcm_b200
    if not (os.path.exists(gcm_b200.library_path()) and os.path.exists(up.EMUL)):
        import __graft_entry__ as g
        g.build()
    return True


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
