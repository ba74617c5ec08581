"""ctypes binding of include/gcm_b200.h, shaped after the reference's host classes.

``TetrMeshSet``        <- gcm/system/TetrMeshSet.{h,cpp}   (orchestrator: do_next_step)
``TetrMesh_1stOrder``  <- gcm/meshtypes/TetrMesh_1stOrder.{h,cpp} (one zone: do_next_part_step)
``GCMException``       <- gcm/system/GCMException.h:19-29
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

BC_FREE, BC_FIXED, BC_EXT_FORCE, BC_EXT_VELOCITY, BC_EXT_VALUES = range(5)

_EXC_NAMES = ["MPI_EXCEPTION", "SYNC_EXCEPTION", "CONFIG_EXCEPTION", "METHOD_EXCEPTION", "MESH_EXCEPTION",
              "SNAP_EXCEPTION", "COLLISION_EXCEPTION", "MATH_EXCEPTION", "UNIMPLEMENTED_EXCEPTION",
              "INPUT_EXCEPTION", "CONFIG_VERSION_MISSMATCH_EXCEPTION"]


class GCMException(Exception):
    """GCMException(code, message); codes as in gcm/system/GCMException.h:19-29."""
    MPI_EXCEPTION, SYNC_EXCEPTION, CONFIG_EXCEPTION, METHOD_EXCEPTION, MESH_EXCEPTION, SNAP_EXCEPTION, \
        COLLISION_EXCEPTION, MATH_EXCEPTION, UNIMPLEMENTED_EXCEPTION, INPUT_EXCEPTION, \
        CONFIG_VERSION_MISSMATCH_EXCEPTION = range(11)

    def __init__(self, code: int, message: str):
        super().__init__("%s: %s" % (_EXC_NAMES[code] if 0 <= code < len(_EXC_NAMES) else code, message))
        self.code = code
        self.message = message

    def getCode(self):
        return self.code

    def getMessage(self):
        return self.message


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgcm_b200.so")


_lib = None

# name -> (restype, argtypes); every symbol include/gcm_b200.h declares
_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
SIGNATURES = {
    "gcm_b200_version": (C.c_char_p, []),
    "gcm_b200_last_error": (C.c_char_p, []),
    "gcm_b200_set_create": (_i, [C.POINTER(_p), _i, _i, _ip, _i]),
    "gcm_b200_set_destroy": (_i, [_p]),
    "gcm_b200_set_seed": (_i, [_p, C.c_uint]),
    "gcm_b200_set_basis_mode": (_i, [_p, _i]),
    "gcm_b200_set_kernel_mode": (_i, [_p, _i]),
    "gcm_b200_set_collision_detector": (_i, [_p, _i, _i, _f]),
    "gcm_b200_set_contact_calculator": (_i, [_p, _i]),
    "gcm_b200_zone_load": (_i, [_p, _i, _i, _i, _fp, _ip, _ip, _ip, _ip, _fp]),
    "gcm_b200_zone_translate": (_i, [_p, _i, _f, _f, _f]),
    "gcm_b200_zone_load_msh_file": (_i, [_p, _i, C.c_char_p, _f, _f, _f, _f]),
    "gcm_b200_read_zones_map": (_i, [C.c_char_p, _ip, _i, _ip]),
    "gcm_b200_zone_dump_vtk": (_i, [_p, _i, C.c_char_p]),
    "gcm_b200_zone_load_vtu_file": (_i, [_p, _i, C.c_char_p]),
    "gcm_b200_zone_set_values": (_i, [_p, _i, _fp]),
    "gcm_b200_zone_get_values": (_i, [_p, _i, _fp]),
    "gcm_b200_zone_upload_values_async": (_i, [_p, _i, _p]),
    "gcm_b200_zone_download_values_async": (_i, [_p, _i, _p]),
    "gcm_b200_set_join_io": (_i, [_p]),
    "gcm_b200_set_add_border_condition": (_i, [_p, _i, _fp, _ip, _fp, _f, _f, _fp]),
    "gcm_b200_set_define_border_condition": (_i, [_p, _i, _fp, _ip, _fp, _f, _f, _ip]),
    "gcm_b200_zone_assign_border_conditions": (_i, [_p, _i, _ip]),
    "gcm_b200_set_preprocess": (_i, [_p]),
    "gcm_b200_set_do_next_step": (_i, [_p]),
    "gcm_b200_set_do_steps": (_i, [_p, _i]),
    "gcm_b200_set_begin_step": (_i, [_p]),
    "gcm_b200_set_sync_nodes": (_i, [_p]),
    "gcm_b200_zone_do_next_part_step": (_i, [_p, _i, _f, _i]),
    "gcm_b200_set_do_next_part_step": (_i, [_p, _f, _i]),
    "gcm_b200_set_concurrency": (_i, [_p, _i]),
    "gcm_b200_set_end_step": (_i, [_p]),
    "gcm_b200_set_track_destruction": (_i, [_p, _i]),
    "gcm_b200_zone_get_destruction_criterias": (_i, [_p, _i, _fp]),
    "gcm_b200_set_synchronize": (_i, [_p]),
    "gcm_b200_set_current_time": (_f, [_p]),
    "gcm_b200_set_stream": (_p, [_p]),
    "gcm_b200_halo_counts": (_i, [_p, _i, _ip, _ip]),
    "gcm_b200_halo_get_requests": (_i, [_p, _i, _ip]),
    "gcm_b200_halo_set_sends": (_i, [_p, _i, _i, _ip]),
    "gcm_b200_halo_pack_host": (_i, [_p, _i, _fp]),
    "gcm_b200_halo_unpack_host": (_i, [_p, _i, _fp]),
    "gcm_b200_halo_pack": (_i, [_p, _i, _p]),
    "gcm_b200_halo_unpack": (_i, [_p, _i, _p]),
    "gcm_b200_comm_get_unique_id": (_i, [_p]),
    "gcm_b200_set_comm_init": (_i, [_p, _p, _i, _i]),
    "gcm_b200_set_comm_sync_nodes": (_i, [_p]),
    "gcm_b200_zone_counts": (_i, [_p, _i, _ip, _ip, _ip, _ip, _ip]),
    "gcm_b200_zone_get_flags": (_i, [_p, _i, _ip]),
    "gcm_b200_zone_get_faces": (_i, [_p, _i, _ip]),
    "gcm_b200_zone_get_bases": (_i, [_p, _i, _fp]),
    "gcm_b200_zone_get_min_max_h": (_i, [_p, _i, _fp, _fp]),
    "gcm_b200_set_enable_tap": (_i, [_p, _i]),
    "gcm_b200_zone_get_tap": (_i, [_p, _i, _ip, _ip, _ip]),
    "gcm_b200_set_launch_count": (C.c_longlong, [_p]),
    "gcm_b200_set_kernel_stats": (_i, [_p, C.POINTER(C.c_longlong)]),
    "gcm_b200_set_virt_count": (_i, [_p]),
    "gcm_b200_set_get_virt": (_i, [_p, _fp]),
    "gcm_b200_set_profile": (_i, [_p, _i]),
    "gcm_b200_set_profile_read": (_i, [_p, _i, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
}


def load_library(path: str | None = None):
    """dlopen libgcm_b200.so and declare every prototype. Fails loudly when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or library_path()
    if not os.path.exists(p):
        raise GCMException(GCMException.CONFIG_EXCEPTION,
                           "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class TetrMesh_1stOrder:
    """One zone of a TetrMeshSet (gcm/meshtypes/TetrMesh_1stOrder.h)."""

    def __init__(self, mesh_set: "TetrMeshSet", zone_num: int):
        self.mesh_set = mesh_set
        self.zone_num = zone_num
        self._n_nodes = 0

    def _ck(self, rc):
        self.mesh_set._ck(rc)

    def load(self, zone_mesh, material=None, la=None, mu=None, rho=None, yield_limit=1e16):
        """add_node/add_tetr for a whole ``meshgen.ZoneMesh``; material [N,4] or constants."""
        n = zone_mesh.n_nodes
        if material is None:
            material = np.empty((n, 4), np.float32)
            material[:] = (la, mu, rho, yield_limit)
        material = np.ascontiguousarray(material, np.float32)
        coords = np.ascontiguousarray(zone_mesh.coords, np.float32)
        tets = np.ascontiguousarray(zone_mesh.tets, np.int32)
        pl = np.ascontiguousarray(zone_mesh.placement, np.int32)
        rz = np.ascontiguousarray(zone_mesh.remote_zone, np.int32)
        rn = np.ascontiguousarray(zone_mesh.remote_num, np.int32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_load(self.mesh_set.h, self.zone_num, n, zone_mesh.n_tets,
                                                     _fptr(coords), _iptr(pl), _iptr(rz), _iptr(rn), _iptr(tets),
                                                     _fptr(material)))
        self._n_nodes = n

    def load_msh_file(self, path, la, mu, rho, yield_limit=1e16):
        """TetrMesh_1stOrder::load_msh_file + the rheology TaskPreparator gives every node."""
        self._ck(self.mesh_set.lib.gcm_b200_zone_load_msh_file(self.mesh_set.h, self.zone_num, str(path).encode(), la, mu, rho, yield_limit))
        self._n_nodes = self.counts()["nodes"]

    def load_vtu_file(self, path):
        """TetrMesh_1stOrder::load_vtu_file: mesh, material and state of a snapshot (restart)."""
        self._ck(self.mesh_set.lib.gcm_b200_zone_load_vtu_file(self.mesh_set.h, self.zone_num, str(path).encode()))
        self._n_nodes = self.counts()["nodes"]

    def dump_vtk(self, path):
        """VTKSnapshotWriter::dump_vtk for this zone."""
        self._ck(self.mesh_set.lib.gcm_b200_zone_dump_vtk(self.mesh_set.h, self.zone_num, str(path).encode()))

    def assign_border_conditions(self, cond_of_node):
        """nodes[i].border_condition = condition cond_of_node[i] (-1: the default Free border)."""
        a = np.ascontiguousarray(cond_of_node, np.int32)
        assert a.shape == (self._n_nodes,)
        self._ck(self.mesh_set.lib.gcm_b200_zone_assign_border_conditions(self.mesh_set.h, self.zone_num, _iptr(a)))

    def translate(self, dx, dy, dz):
        self._ck(self.mesh_set.lib.gcm_b200_zone_translate(self.mesh_set.h, self.zone_num, dx, dy, dz))

    def set_values(self, values):
        v = np.ascontiguousarray(values, np.float32).reshape(self._n_nodes, 9)
        self._ck(self.mesh_set.lib.gcm_b200_zone_set_values(self.mesh_set.h, self.zone_num, _fptr(v)))

    def get_values(self):
        v = np.empty((self._n_nodes, 9), np.float32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_values(self.mesh_set.h, self.zone_num, _fptr(v)))
        return v

    def get_destruction_criterias(self):
        """[N, 8]: ElasticNode::destruction_criterias (4 of the current values + their running maxima)."""
        v = np.empty((self._n_nodes, 8), np.float32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_destruction_criterias(self.mesh_set.h, self.zone_num, _fptr(v)))
        return v

    def upload_values_async(self, host_ptr):
        self._ck(self.mesh_set.lib.gcm_b200_zone_upload_values_async(self.mesh_set.h, self.zone_num, C.c_void_p(host_ptr)))

    def download_values_async(self, host_ptr):
        self._ck(self.mesh_set.lib.gcm_b200_zone_download_values_async(self.mesh_set.h, self.zone_num, C.c_void_p(host_ptr)))

    def do_next_part_step(self, tau, stage):
        self._ck(self.mesh_set.lib.gcm_b200_zone_do_next_part_step(self.mesh_set.h, self.zone_num, tau, stage))

    def counts(self):
        c = [C.c_int() for _ in range(5)]
        self._ck(self.mesh_set.lib.gcm_b200_zone_counts(self.mesh_set.h, self.zone_num, *[C.byref(x) for x in c]))
        return dict(zip(["nodes", "tets", "faces", "border_nodes", "local_nodes"], [x.value for x in c]))

    def get_flags(self):
        f = np.empty(self._n_nodes, np.int32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_flags(self.mesh_set.h, self.zone_num, _iptr(f)))
        return f

    def get_faces(self):
        f = np.empty((self.counts()["faces"], 3), np.int32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_faces(self.mesh_set.h, self.zone_num, _iptr(f)))
        return f

    def get_bases(self):
        b = np.empty((self._n_nodes, 9), np.float32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_bases(self.mesh_set.h, self.zone_num, _fptr(b)))
        return b

    def get_min_max_h(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_min_max_h(self.mesh_set.h, self.zone_num, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_tap(self):
        n = self._n_nodes
        owners = np.empty((3, n, 5), np.int32)
        outer = np.empty((3, n), np.int32)
        tries = np.empty((3, n), np.int32)
        self._ck(self.mesh_set.lib.gcm_b200_zone_get_tap(self.mesh_set.h, self.zone_num, _iptr(owners), _iptr(outer), _iptr(tries)))
        return owners, outer, tries


class TetrMeshSet:
    """The orchestrator (gcm/system/TetrMeshSet.h) for the zones this process owns."""

    def __init__(self, n_zones=1, zone_rank=None, my_rank=0, device=0, lib=None):
        self.lib = lib or load_library()
        self.h = C.c_void_p()
        zr = None if zone_rank is None else np.ascontiguousarray(zone_rank, np.int32)
        rc = self.lib.gcm_b200_set_create(C.byref(self.h), device, n_zones, _iptr(zr), my_rank)
        self._ck(rc)
        self.n_zones = n_zones
        self.zone_rank = list(range(0)) if zr is None else list(zr)
        self.my_rank = my_rank
        self.meshes = {}
        for z in range(n_zones):
            if zr is None or zr[z] == my_rank:
                self.meshes[z] = TetrMesh_1stOrder(self, z)

    def _ck(self, rc):
        if rc != 0:
            raise GCMException(-rc - 1, self.lib.gcm_b200_last_error().decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.gcm_b200_set_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_mesh_by_zone_num(self, z) -> TetrMesh_1stOrder:
        return self.meshes[z]

    def srand(self, seed):
        self._ck(self.lib.gcm_b200_set_seed(self.h, seed))

    def set_kernel_mode(self, fast_inner=True):
        self._ck(self.lib.gcm_b200_set_kernel_mode(self.h, int(fast_inner)))

    def set_basis_mode(self, on_device=True):
        self._ck(self.lib.gcm_b200_set_basis_mode(self.h, int(on_device)))

    def set_collision_detector(self, kind="void", static=False, threshold=1.0):
        t = {"void": 0, "brute": 1}[kind]
        self._ck(self.lib.gcm_b200_set_collision_detector(self.h, t, int(static), threshold))

    def set_contact_calculator(self, kind="adhesion"):
        self._ck(self.lib.gcm_b200_set_contact_calculator(self.h, {"adhesion": 0, "friction": 1}[kind]))

    def pre_process_meshes(self):
        self._ck(self.lib.gcm_b200_set_preprocess(self.h))

    def add_border_condition(self, calc_type, box, params=None, start=-1.0, duration=-1.0, vars=None, vals=None):
        p = None if params is None else np.ascontiguousarray(params, np.float32)
        b = np.ascontiguousarray(box, np.float32)
        vi = None if vars is None else np.ascontiguousarray(vars, np.int32)
        vv = None if vals is None else np.ascontiguousarray(vals, np.float32)
        self._ck(self.lib.gcm_b200_set_add_border_condition(self.h, calc_type, _fptr(p), _iptr(vi), _fptr(vv),
                                                           start, duration, _fptr(b)))

    def define_border_condition(self, calc_type, params=None, start=-1.0, duration=-1.0, vars=None, vals=None):
        """A BorderCondition without its area (calculator + StepPulseForm); returns its id for
        TetrMesh_1stOrder.assign_border_conditions."""
        p = None if params is None else np.ascontiguousarray(params, np.float32)
        vi = None if vars is None else np.ascontiguousarray(vars, np.int32)
        vv = None if vals is None else np.ascontiguousarray(vals, np.float32)
        cid = C.c_int(-1)
        self._ck(self.lib.gcm_b200_set_define_border_condition(self.h, calc_type, _fptr(p), _iptr(vi), _fptr(vv),
                                                              start, duration, C.byref(cid)))
        return cid.value

    def do_next_step(self):
        self._ck(self.lib.gcm_b200_set_do_next_step(self.h))

    def do_steps(self, n):
        self._ck(self.lib.gcm_b200_set_do_steps(self.h, n))

    def begin_step(self):
        self._ck(self.lib.gcm_b200_set_begin_step(self.h))

    def sync_nodes(self):
        self._ck(self.lib.gcm_b200_set_sync_nodes(self.h))

    def do_next_part_step(self, tau, stage):
        """One stage for every local zone (the per-mesh loop of TetrMeshSet.cpp:158-167)."""
        self._ck(self.lib.gcm_b200_set_do_next_part_step(self.h, tau, stage))

    def set_concurrency(self, on=True):
        self._ck(self.lib.gcm_b200_set_concurrency(self.h, int(on)))

    def end_step(self):
        self._ck(self.lib.gcm_b200_set_end_step(self.h))

    def join_io(self):
        """The set's stream waits for every download_values_async issued so far."""
        self._ck(self.lib.gcm_b200_set_join_io(self.h))

    def track_destruction(self, on=True):
        """Keep the running maxima of calc_destruction_criterias (after pre_process_meshes)."""
        self._ck(self.lib.gcm_b200_set_track_destruction(self.h, int(on)))

    def synchronize(self):
        self._ck(self.lib.gcm_b200_set_synchronize(self.h))

    def get_current_time(self):
        return self.lib.gcm_b200_set_current_time(self.h)

    def stream(self):
        return self.lib.gcm_b200_set_stream(self.h)

    def enable_tap(self, on=True):
        self._ck(self.lib.gcm_b200_set_enable_tap(self.h, int(on)))

    def virt_count(self):
        return self.lib.gcm_b200_set_virt_count(self.h)

    def get_virt_nodes(self):
        """Virtual (contact) nodes of the last collision pass: [n, 16] = coords(3) + values(13)."""
        n = self.lib.gcm_b200_set_virt_count(self.h)
        v = np.empty((n, 16), np.float32)
        if n:
            self._ck(self.lib.gcm_b200_set_get_virt(self.h, _fptr(v)))
        return v

    def launch_count(self):
        return self.lib.gcm_b200_set_launch_count(self.h)

    def kernel_stats(self):
        """Coverage of the specialised kernels (include/gcm_b200.h: gcm_b200_set_kernel_stats)."""
        a = (C.c_longlong * 16)()
        self._ck(self.lib.gcm_b200_set_kernel_stats(self.h, a))
        return dict(inner_cached=a[0], inner_cold=a[1], patterns=[a[2], a[3], a[4]], pattern_axes=[a[5], a[6], a[7]],
                    border_flat=a[8], border_sharp=a[9], flat_deferred_last=a[10], exact_only=a[11], peer_stores=a[12], inner_shell=a[13])

    def profile(self, on=True):
        self._ck(self.lib.gcm_b200_set_profile(self.h, int(on)))

    def profile_read(self, which):
        ms, n, u = C.c_double(), C.c_longlong(), C.c_longlong()
        self._ck(self.lib.gcm_b200_set_profile_read(self.h, which, C.byref(ms), C.byref(n), C.byref(u)))
        return ms.value, n.value, u.value

    # ---- cross-process halo (DataBus role) ------------------------------------------------
    def halo_counts(self, peer):
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.gcm_b200_halo_counts(self.h, peer, C.byref(a), C.byref(b)))
        return a.value, b.value

    def halo_get_requests(self, peer):
        n, _ = self.halo_counts(peer)
        pairs = np.empty((n, 2), np.int32)
        self._ck(self.lib.gcm_b200_halo_get_requests(self.h, peer, _iptr(pairs)))
        return pairs

    def halo_set_sends(self, peer, pairs):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._ck(self.lib.gcm_b200_halo_set_sends(self.h, peer, len(pairs), _iptr(pairs)))

    def halo_pack_host(self, peer):
        _, n = self.halo_counts(peer)
        buf = np.empty((n, 16), np.float32)
        self._ck(self.lib.gcm_b200_halo_pack_host(self.h, peer, _fptr(buf)))
        return buf

    def halo_unpack_host(self, peer, buf):
        buf = np.ascontiguousarray(buf, np.float32)
        self._ck(self.lib.gcm_b200_halo_unpack_host(self.h, peer, _fptr(buf)))

    def halo_pack(self, peer, dev_ptr):
        self._ck(self.lib.gcm_b200_halo_pack(self.h, peer, C.c_void_p(dev_ptr)))

    def halo_unpack(self, peer, dev_ptr):
        self._ck(self.lib.gcm_b200_halo_unpack(self.h, peer, C.c_void_p(dev_ptr)))

    def comm_get_unique_id(self) -> bytes:
        """ncclUniqueId (128 bytes) for gcm_b200_set_comm_init; rank 0 makes it, the host broadcasts it."""
        buf = C.create_string_buffer(128)
        self._ck(self.lib.gcm_b200_comm_get_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """Collective: the set gets its own NCCL communicator and moves the halos itself from now on."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.gcm_b200_set_comm_init(self.h, buf, rank, n_ranks))

    def comm_sync_nodes(self):
        self._ck(self.lib.gcm_b200_set_comm_sync_nodes(self.h))
