"""Synthetic tetrahedral meshes for the GCM step (harness side: tests, bench, golden vectors).

Restates the reference's generator ``gcm/tools/create_cube.c:21-152`` (N zones x M layers,
6 tets per hex in the vertex order of ``write_tetrs`` :120-147, halo nodes encoded as
"remote" entries, :49-81) without its 3-tag element bug (the loader
``TetrMesh_1stOrder.cpp:838-846`` expects 2 tags).  Meshes are produced in memory as numpy
arrays; ``write_gcmb`` / ``write_msh`` serialise them for the oracle driver / the
reference's own ``load_msh_file``.
"""
from __future__ import annotations

import dataclasses
import numpy as np

# 6 tets per hex, corner offsets (di, dj, dk) in the order of create_cube.c:120-147
_TET_PATTERN = np.array([
    [(0, 0, 0), (0, 1, 1), (0, 1, 0), (1, 1, 0)],
    [(0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 0)],
    [(0, 1, 1), (0, 0, 0), (1, 1, 1), (1, 1, 0)],
    [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)],
    [(0, 0, 1), (1, 0, 0), (1, 1, 1), (0, 0, 0)],
    [(1, 0, 1), (1, 1, 1), (0, 0, 1), (1, 0, 0)],
], dtype=np.int64)


@dataclasses.dataclass
class ZoneMesh:
    """One zone as the reference's loader would see it (all indices 0-based)."""
    zone: int
    coords: np.ndarray        # float32 [N,3]; halo rows hold the owner's coordinates
    placement: np.ndarray     # int32 [N]; 1 = local, 0 = remote (halo)
    remote_zone: np.ndarray   # int32 [N]; -1 for local nodes
    remote_num: np.ndarray    # int32 [N]; index in the owner zone, -1 for local nodes
    tets: np.ndarray          # int32 [T,4]

    @property
    def n_nodes(self) -> int:
        return int(self.coords.shape[0])

    @property
    def n_tets(self) -> int:
        return int(self.tets.shape[0])

    @property
    def n_local(self) -> int:
        return int(self.placement.sum())


def _slab_tets(nx: int, ny: int, nlayers: int) -> np.ndarray:
    """Tets of a slab of nlayers node layers, local numbering k*nx*ny + j*nx + i."""
    if nlayers < 2:
        return np.zeros((0, 4), np.int32)
    k, j, i = np.meshgrid(np.arange(nlayers - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    i = i.reshape(-1, 1, 1); j = j.reshape(-1, 1, 1); k = k.reshape(-1, 1, 1)
    p = _TET_PATTERN[None]  # [1,6,4,3]
    idx = (k + p[..., 2]) * (nx * ny) + (j + p[..., 1]) * nx + (i + p[..., 0])
    return idx.reshape(-1, 4).astype(np.int32)


def make_box_zones(nx: int, ny: int, nz: int, n_zones: int = 1, spacing=(1.0, 1.0, 1.0),
                   origin=(0.0, 0.0, 0.0), jitter: float = 0.0, jitter_border: bool = False,
                   seed: int = 0, only=None, layers=None) -> list[ZoneMesh]:
    """nx x ny x nz node box split into n_zones z-slabs, create_cube.c semantics.

    With nx == ny == nz == N*M and n_zones == N this is exactly ``create_cube`` with
    Nnum=N, Mnum=M.  ``jitter`` displaces nodes by a uniform random offset of up to
    +-jitter (in units of the spacing) so that owner searches are not all axis-aligned ties;
    the offset is a function of the GLOBAL node id, so every zone sees the same coordinates.
    ``only`` = iterable of zone numbers to build (others are returned as None): a rank of a
    multi-process run builds only the zones it owns.
    ``layers`` = per-zone number of owned z-layers (sum = nz) for slabs of unequal thickness (the end
    slabs carry a whole face of border nodes: thinner end slabs balance the work per zone).
    """
    if layers is None:
        if nz % n_zones:
            raise ValueError("nz must be divisible by n_zones")
        layers = [nz // n_zones] * n_zones
    layers = [int(x) for x in layers]
    if len(layers) != n_zones or sum(layers) != nz or min(layers) < 1:
        raise ValueError("layers must give every zone >= 1 layer and sum to nz")
    start = [0]
    for x in layers:
        start.append(start[-1] + x)
    sp = np.asarray(spacing, np.float64); org = np.asarray(origin, np.float64)
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ijk = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1)
    xyz = ijk.astype(np.float64)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, size=xyz.shape)
        if not jitter_border:
            on_border = ((ijk == 0) | (ijk == np.array([nx - 1, ny - 1, nz - 1]))).any(1)
            d[on_border] = 0.0
        xyz = xyz + d
    xyz = (xyz * sp + org).astype(np.float32)
    layer = nx * ny
    zones = []
    for n in range(n_zones):
        if only is not None and n not in only:
            zones.append(None)
            continue
        k_lo = start[n] - (1 if n > 0 else 0)
        k_hi = start[n + 1] + (1 if n < n_zones - 1 else 0)   # exclusive
        nl = k_hi - k_lo
        gid = np.arange(k_lo * layer, k_hi * layer)
        coords = xyz[gid].copy()
        placement = np.ones(nl * layer, np.int32)
        rz = np.full(nl * layer, -1, np.int32)
        rn = np.full(nl * layer, -1, np.int32)
        if n > 0:   # low halo layer k = start[n]-1, owned by zone n-1
            lo_prev = start[n - 1] - (1 if n - 1 > 0 else 0)
            sl = slice(0, layer)
            placement[sl] = 0; rz[sl] = n - 1
            rn[sl] = (start[n] - 1 - lo_prev) * layer + np.arange(layer)
        if n < n_zones - 1:  # top halo layer k = start[n+1], owned by zone n+1
            lo_next = start[n + 1] - 1
            sl = slice((nl - 1) * layer, nl * layer)
            placement[sl] = 0; rz[sl] = n + 1
            rn[sl] = (start[n + 1] - lo_next) * layer + np.arange(layer)
        zones.append(ZoneMesh(n, coords, placement, rz, rn, _slab_tets(nx, ny, nl)))
    return zones


def make_cube(n_side: int, **kw) -> ZoneMesh:
    """Single-zone n_side^3 cube (BASELINE config 2 with n_side = 100)."""
    return make_box_zones(n_side, n_side, n_side, 1, **kw)[0]


def pwave_state(coords: np.ndarray, la: float, mu: float, rho: float, z_c: float | None = None,
                width: float = 3.0, axis: int = 2, amp: float = 1.0) -> np.ndarray:
    """Gaussian plane P-wave along `axis`: v_a = g, s_aa = -rho*c1*g, lateral
    normal stresses = la/(la+2mu)*s_aa.  Returns float32 [N,9]."""
    x = coords[:, axis].astype(np.float64)
    if z_c is None:
        z_c = 0.5 * (x.min() + x.max())
    g = amp * np.exp(-((x - z_c) / width) ** 2)
    c1 = np.sqrt((la + 2 * mu) / rho)
    u = np.zeros((coords.shape[0], 9), np.float64)
    diag = [3, 6, 8]
    u[:, axis] = g
    s = -rho * c1 * g
    for a in range(3):
        u[:, diag[a]] = s if a == axis else la / (la + 2 * mu) * s
    return u.astype(np.float32)


def generic_state(coords: np.ndarray, la: float, mu: float, rho: float, seed: int = 1) -> np.ndarray:
    """Smooth non-symmetric field exercising all 9 variables (parity tests)."""
    c = coords.astype(np.float64)
    rng = np.random.default_rng(seed)
    k = rng.uniform(0.15, 0.45, size=(9, 3)); ph = rng.uniform(0, 6.28, size=9)
    u = np.sin(c @ k.T + ph)
    u[:, 3:] *= rho * np.sqrt((la + 2 * mu) / rho)
    return u.astype(np.float32)


def write_gcmb(path: str, z: ZoneMesh) -> None:
    """Binary mesh for oracle/ref_build/driver.cpp (format documented there)."""
    with open(path, "wb") as f:
        np.array([0x424D4347, 1, z.zone, z.n_nodes, z.n_tets], np.int32).tofile(f)
        z.placement.astype(np.int32).tofile(f)
        z.coords.astype(np.float32).tofile(f)
        z.remote_zone.astype(np.int32).tofile(f)
        z.remote_num.astype(np.int32).tofile(f)
        z.tets.astype(np.int32).tofile(f)


def write_msh(path: str, z: ZoneMesh) -> None:
    """gmsh-2 ASCII subset accepted by load_msh_file (TetrMesh_1stOrder.cpp:735-866)."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % z.n_nodes)
        for i in range(z.n_nodes):
            if z.placement[i]:
                f.write("%d %.9g %.9g %.9g\n" % (i + 1, *z.coords[i]))
            else:
                f.write("%d %d %d\n" % (-(i + 1), z.remote_zone[i], z.remote_num[i] + 1))
        f.write("$EndNodes\n$Elements\n%d\n" % z.n_tets)
        for t in range(z.n_tets):
            f.write("%d 4 2 0 1 %d %d %d %d\n" % (t + 1, *(z.tets[t] + 1)))
        f.write("$EndElements\n")


def make_delaunay_box(n_side: int, jitter: float = 0.3, seed: int = 0, min_quality: float = 0.02) -> ZoneMesh:
    """Unstructured stand-in for a gmsh mesh (variant ii): the Delaunay
    tetrahedralisation (scipy) of an n_side^3 lattice whose inner points are displaced by up to
    +-jitter; border points stay on the box faces so that the surface is flat.  Node degrees, the
    position of a node inside its tets and the tet orientations are all irregular.  Slivers
    (6 |Vol| < min_quality * longest edge^3) on the flat faces are dropped, as a mesher would not
    produce them.  Needs scipy (test-side utility; not used by the library)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    k, j, i = np.meshgrid(np.arange(n_side), np.arange(n_side), np.arange(n_side), indexing="ij")
    ijk = np.stack([i.ravel(), j.ravel(), k.ravel()], 1)
    xyz = ijk.astype(np.float64)
    d = rng.uniform(-jitter, jitter, size=xyz.shape)
    on_face = (ijk == 0) | (ijk == n_side - 1)
    d[on_face] = 0.0                       # a coordinate that lies on a face stays on it
    xyz = (xyz + d).astype(np.float32)
    tets = Delaunay(xyz.astype(np.float64)).simplices.astype(np.int32)
    p = xyz[tets].astype(np.float64)
    vol6 = np.abs(np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0]))
    edges = np.stack([np.linalg.norm(p[:, a] - p[:, b], axis=1) for a in range(4) for b in range(a + 1, 4)], 1)
    keep = vol6 > min_quality * edges.max(1) ** 3
    tets = np.ascontiguousarray(tets[keep])
    n = xyz.shape[0]
    return ZoneMesh(0, xyz, np.ones(n, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32), tets)


def write_msh(path: str, z: ZoneMesh, extra_elements: bool = True) -> None:
    """gmsh-2 ASCII file as TetrMesh_1stOrder::load_msh_file reads it: local nodes ``id x y z``, halo nodes
    ``-id remote_zone remote_id`` (1-based ids, the zone splitter's encoding, gcm/tools/create_cube.c:38-81),
    tetrahedra ``num 4 2 0 1 v0 v1 v2 v3`` -- exactly three integers between the type and the vertices (two tags;
    create_cube.c itself writes three tags and needs the sed of SURVEY fact 9).  ``extra_elements`` adds a few
    non-tetrahedral elements (points, triangles), which the reader must skip."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % z.n_nodes)
        for i in range(z.n_nodes):
            if z.placement[i]:
                f.write("%d %s %s %s\n" % (i + 1, repr(float(z.coords[i, 0])), repr(float(z.coords[i, 1])), repr(float(z.coords[i, 2]))))
            else:
                f.write("%d %d %d\n" % (-(i + 1), z.remote_zone[i], z.remote_num[i] + 1))
        extra = []
        if extra_elements:
            extra = ["%d 15 2 0 1 1" % 1, "%d 2 2 0 1 1 2 3" % 2]
        f.write("$EndNodes\n$Elements\n%d\n" % (z.n_tets + len(extra)))
        for e in extra:
            f.write(e + "\n")
        for t in range(z.n_tets):
            v = z.tets[t]
            f.write("%d 4 2 0 1 %d %d %d %d\n" % (t + 1 + len(extra), v[0] + 1, v[1] + 1, v[2] + 1, v[3] + 1))
        f.write("$EndElements\n")
