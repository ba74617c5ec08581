"""On-disk formats either side of the hot path (gcm-necro_b200/csrc/gcm_io.cpp): the gmsh-2 .msh subset of
TetrMesh_1stOrder::load_msh_file with the zone splitter's halo lines, the zones-map XML of
TaskPreparator::load_zones_info, .vtu snapshots / restart.  CPU tests use host-only sets (device -1: loading
and preprocessing are host C++); the reference's own load_msh_file reads the same files in the oracle build."""
import ctypes as C
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import util_parity as up
import gcm_b200
from gcm_b200 import meshgen


def _host_set(n_zones):
    return gcm_b200.TetrMeshSet(n_zones=n_zones, device=-1)


def test_msh_files_give_the_golden_preprocessing(built, tmp_path):
    """zones3_jitter written as three .msh files (halo nodes as -id zone id), read back by the library: flags and
    border faces after pre_process_meshes equal the fixture the reference produced from the same meshes."""
    case = up.make_case("zones3_jitter")
    gold = up.load_golden("zones3_jitter")
    ms = _host_set(3)
    for k, z in enumerate(case["zones"]):
        p = tmp_path / ("zone%d.msh" % k)
        meshgen.write_msh(str(p), z)
        ms.get_mesh_by_zone_num(k).load_msh_file(p, up.LA, up.MU, up.RHO)
    ms.pre_process_meshes()
    for k, z in enumerate(case["zones"]):
        m = ms.get_mesh_by_zone_num(k)
        c = m.counts()
        assert c["nodes"] == z.n_nodes and c["tets"] == z.n_tets
        assert np.array_equal(m.get_flags() & 15, gold[k]["flags"])
        assert np.array_equal(m.get_faces(), gold[k]["faces"])
    ms.close()


def test_reference_reads_the_same_msh_files(built, tmp_path):
    """The reference's own TetrMesh_1stOrder::load_msh_file on the files the library reads: same flags / faces /
    values after the steps as from the binary meshes (pins write_msh's format to what the reference accepts)."""
    if not os.path.exists(up.ORACLE):
        pytest.skip("oracle/_ref not built")
    case = up.make_case("zones2")
    gold = up.load_golden("zones2")
    meshes, states = [], []
    for k, (z, st) in enumerate(zip(case["zones"], case["states"])):
        p = str(tmp_path / ("z%d.msh" % k))
        meshgen.write_msh(p, z)
        sp = str(tmp_path / ("z%d.state" % k))
        np.ascontiguousarray(st, np.float32).tofile(sp)
        meshes.append(p); states.append(sp)
    out = str(tmp_path / "o")
    r = subprocess.run([up.ORACLE] + up._cli(case, meshes, states, out), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    for k, z in enumerate(case["zones"]):
        v = np.fromfile("%s.z%d.values.f32" % (out, k), np.float32).reshape(-1, 9)
        fl = np.fromfile("%s.z%d.flags.i32" % (out, k), np.int32)
        assert np.array_equal(fl, gold[k]["flags"])
        loc = up.local_mask(fl)
        assert np.array_equal(v[loc], gold[k]["values"][loc])


@pytest.mark.parametrize("text, msg", [
    ("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n0 0 0 0\n$EndNodes\n", "Wrong file format"),
    ("$Mesh\n", "Wrong file format"),
    ("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 4 2 0 1 1 1 1 0\n$EndElements\n", "Wrong file format"),
])
def test_msh_errors_are_the_reference_exceptions(built, tmp_path, text, msg):
    p = tmp_path / "bad.msh"
    p.write_text(text)
    ms = _host_set(1)
    with pytest.raises(gcm_b200.GCMException) as e:
        ms.get_mesh_by_zone_num(0).load_msh_file(p, 1.0, 1.0, 1.0)
    assert e.value.code == gcm_b200.GCMException.MESH_EXCEPTION and msg in e.value.message
    with pytest.raises(gcm_b200.GCMException) as e:
        ms.get_mesh_by_zone_num(0).load_msh_file(tmp_path / "missing.msh", 1.0, 1.0, 1.0)
    assert "Can not open msh file" in e.value.message
    ms.close()


def _zones_map(path):
    lib = gcm_b200.load_library()
    arr = (C.c_int * 64)()
    n = C.c_int()
    rc = lib.gcm_b200_read_zones_map(str(path).encode(), arr, 64, C.byref(n))
    if rc:
        raise gcm_b200.GCMException(-rc - 1, lib.gcm_b200_last_error().decode())
    return [arr[k] for k in range(n.value)]


def test_zones_map_xml(built, tmp_path):
    """tasks/*-zones.xml shape (TaskPreparator.cpp:65-99): cpu per zone, num counting from 0."""
    p = tmp_path / "zones.xml"
    p.write_text('<?xml version="1.0"?>\n<!-- 8 zones on 4 cpus -->\n<zones_map>\n' +
                 "".join('\t<zone num="%d" cpu="%d" />\n' % (k, k // 2) for k in range(8)) + "</zones_map>\n")
    assert _zones_map(p) == [0, 0, 1, 1, 2, 2, 3, 3]
    bad = tmp_path / "bad.xml"
    bad.write_text('<zones_map><zone num="1" cpu="0"/></zones_map>')
    with pytest.raises(gcm_b200.GCMException) as e:
        _zones_map(bad)
    assert e.value.code == gcm_b200.GCMException.CONFIG_EXCEPTION and "Malformed zone map file" in e.value.message
    with pytest.raises(gcm_b200.GCMException) as e:
        _zones_map(tmp_path / "none.xml")
    assert "Can not open zone map file" in e.value.message
    ref_xml = "/root/reference/tasks/basic-contact-task-1cpu-zones.xml"
    if os.path.exists(ref_xml):          # the README run's zones file (BASELINE configs[0])
        assert _zones_map(ref_xml) == [0, 0]


def test_vtu_snapshot_round_trip(built, tmp_path):
    """dump_vtk on a host-only set (state as loaded): a well-formed UnstructuredGrid with the reference's array
    names, and load_vtu_file gives the same mesh, material and state back, bit for bit."""
    z = meshgen.make_cube(5, jitter=0.2, seed=3)
    st = meshgen.generic_state(z.coords, up.LA, up.MU, up.RHO, seed=9)
    ms = _host_set(1)
    m = ms.get_mesh_by_zone_num(0)
    m.load(z, la=up.LA, mu=up.MU, rho=up.RHO, yield_limit=1e7)
    m.set_values(st)
    ms.pre_process_meshes()
    p = tmp_path / "snap.vtu"
    m.dump_vtk(p)
    root = ET.parse(str(p)).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == z.n_nodes and int(piece.get("NumberOfCells")) == z.n_tets
    names = [a.get("Name") for a in piece.find("PointData")]
    assert names == ["velocity", "sxx", "sxy", "sxz", "syy", "syz", "szz", "lambda", "mu", "rho", "yieldLimit", "contact", "flags",
                     "maxCompression", "maxTension", "maxShear", "maxDeviator",
                     "maxCompressionHistory", "maxTensionHistory", "maxShearHistory", "maxDeviatorHistory"]
    assert piece.find("PointData").get("Vectors") == "velocity"
    sxx = np.array(piece.find("PointData")[1].text.split(), np.float64).astype(np.float32)
    assert np.array_equal(sxx, st[:, 3])
    flags = np.array(piece.find("PointData")[12].text.split(), np.int64)
    assert np.array_equal(flags, m.get_flags())
    ms2 = _host_set(1)
    m2 = ms2.get_mesh_by_zone_num(0)
    m2.load_vtu_file(p)
    assert m2.counts()["nodes"] == z.n_nodes and m2.counts()["tets"] == z.n_tets
    assert np.array_equal(m2.get_values(), st)
    ms2.pre_process_meshes()
    assert np.array_equal(m2.get_flags(), m.get_flags()) and np.array_equal(m2.get_faces(), m.get_faces())
    ms.close(); ms2.close()
