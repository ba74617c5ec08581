"""gcm_b200 -- Python face of libgcm_b200.so (harness side: tests, bench, examples).

The product is the CUDA library behind the C ABI of ``include/gcm_b200.h``; this package only
binds it with ctypes, mirroring the reference's class names (``TetrMeshSet``,
``TetrMesh_1stOrder``) so that parity tests read like the reference's driver code.
There is no CPU execution path: importing works anywhere, but creating a ``TetrMeshSet``
raises ``GCMException`` unless the library is built and a CUDA device is present.
"""
from .capi import (GCMException, TetrMeshSet, TetrMesh_1stOrder, library_path, load_library,
                   BC_FREE, BC_FIXED, BC_EXT_FORCE, BC_EXT_VELOCITY, BC_EXT_VALUES)
from . import meshgen

__all__ = ["GCMException", "TetrMeshSet", "TetrMesh_1stOrder", "library_path", "load_library", "meshgen",
           "BC_FREE", "BC_FIXED", "BC_EXT_FORCE", "BC_EXT_VELOCITY", "BC_EXT_VALUES"]
