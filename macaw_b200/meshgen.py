"""Synthetic meshes for the BASELINE.json configs, in libMesh conventions.

Element / node / side numbering of the structured grids follows MOOSE's GeneratedMeshGenerator as
stated in-tree by ``moose/framework/src/meshgenerators/DistributedRectilinearMeshGenerator.C:652-776``
(node id = i + (nx+1)(j + k(ny+1)); HEX8 boundary ids back=0 bottom=1 right=2 top=3 left=4 front=5;
QUAD4 bottom=0 right=1 top=2 left=3).  Everything here is host-side numpy set-up code; the hot path
never sees it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

QUAD4, TET4, HEX8 = 0, 1, 2
BC_KILL, BC_REFLECT = 0, 1

SIDE_NODES = {
    QUAD4: np.array([[0, 1], [1, 2], [2, 3], [3, 0]]),
    TET4: np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]]),
    HEX8: np.array([[0, 3, 2, 1], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]]),
}
EDGE_NODES = {
    TET4: np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]]),
    HEX8: np.array([[0, 1], [1, 2], [2, 3], [0, 3], [0, 4], [1, 5], [2, 6], [3, 7], [4, 5], [5, 6], [6, 7], [4, 7]]),
}


@dataclass
class Mesh:
    elem_type: int
    xyz: np.ndarray            # (n_nodes, 3) float64
    conn: np.ndarray           # (n_elem, nv) int64
    block_id: np.ndarray       # (n_elem,) int32 raw subdomain ids
    neighbors: np.ndarray = None   # (n_elem, ns) int64, -1 on the boundary
    side_bnd_id: np.ndarray = None  # (n_elem, ns) int32, -1 interior
    bnd_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    name: str = ""

    @property
    def n_elem(self):
        return self.conn.shape[0]

    @property
    def n_sides(self):
        return SIDE_NODES[self.elem_type].shape[0]

    @property
    def dim(self):
        return 2 if self.elem_type == QUAD4 else 3

    def block_ids_sorted(self):
        return np.unique(self.block_id).astype(np.int32)

    def block_index(self, block_ids=None):
        """Position of every element's block in `block_ids` (default: the sorted unique ids)."""
        ids = self.block_ids_sorted() if block_ids is None else np.asarray(block_ids, np.int64)
        order = np.argsort(ids, kind="stable")
        pos = np.searchsorted(ids[order], self.block_id)
        return order[pos].astype(np.int32)

    def centroids(self):
        return self.xyz[self.conn].mean(axis=1)


def compute_neighbors(elem_type: int, conn: np.ndarray) -> np.ndarray:
    """Face-matching neighbour table (what libMesh's find_neighbors gives: elem->neighbor_ptr(s))."""
    sn = SIDE_NODES[elem_type]
    n_elem, ns = conn.shape[0], sn.shape[0]
    faces = np.sort(conn[:, sn].reshape(n_elem * ns, -1), axis=1)
    order = np.lexsort(faces.T[::-1])
    fs = faces[order]
    same = np.all(fs[1:] == fs[:-1], axis=1)
    nbr = np.full(n_elem * ns, -1, np.int64)
    a, b = order[:-1][same], order[1:][same]
    nbr[a] = b // ns
    nbr[b] = a // ns
    return nbr.reshape(n_elem, ns)


def _finish(mesh: Mesh, bnd_of_side, n_bnd: int, bc=BC_REFLECT) -> Mesh:
    mesh.neighbors = compute_neighbors(mesh.elem_type, mesh.conn)
    sb = np.full(mesh.neighbors.shape, -1, np.int32)
    mask = mesh.neighbors < 0
    sb[mask] = bnd_of_side(mask)
    mesh.side_bnd_id = sb
    mesh.bnd_kind = np.full(n_bnd, bc, np.int32) if np.isscalar(bc) else np.asarray(bc, np.int32)
    return mesh


def _box_bnd_ids(mesh: Mesh, lo, hi):
    """Boundary id of every boundary side of a box-shaped mesh from its side centroid."""
    sn = SIDE_NODES[mesh.elem_type]
    cent = mesh.xyz[mesh.conn[:, sn]].mean(axis=2)  # (n_elem, ns, 3)
    ext = np.maximum(np.asarray(hi, float) - np.asarray(lo, float), 1e-300)
    tol = 1e-9 * ext.max()

    def fn(mask):
        c = cent[mask]
        out = np.full(c.shape[0], -1, np.int32)
        if mesh.dim == 3:
            tests = [(2, lo[2], 0), (1, lo[1], 1), (0, hi[0], 2), (1, hi[1], 3), (0, lo[0], 4), (2, hi[2], 5)]
        else:
            tests = [(1, lo[1], 0), (0, hi[0], 1), (1, hi[1], 2), (0, lo[0], 3)]
        for axis, val, bid in tests:
            out[(np.abs(c[:, axis] - val) <= tol) & (out < 0)] = bid
        assert (out >= 0).all(), "boundary side not on the box"
        return out

    return fn


def hex_grid(nx, ny, nz, lo=(0, 0, 0), hi=(1, 1, 1), bc=BC_REFLECT, block_fn=None, name="hex_grid") -> Mesh:
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    xyz = np.stack([xs[I].ravel(), ys[J].ravel(), zs[K].ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()

    def nid(a, b, c):
        return a + (nx + 1) * (b + c * (ny + 1))

    conn = np.stack([nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)],
                    axis=1).astype(np.int64)
    m = Mesh(HEX8, xyz, conn, np.zeros(conn.shape[0], np.int32), name=name)
    if block_fn is not None:
        m.block_id = np.asarray(block_fn(m.centroids()), np.int32)
    # structured grid: neighbours and boundary ids follow from (i, j, k) -- side s faces (k-1, j-1, i+1, j+1, i-1, k+1)
    # with boundary ids back=0 bottom=1 right=2 top=3 left=4 front=5; face matching
    # (compute_neighbors) gives the same tables and is what the unstructured meshes use
    eid = i + nx * (j + ny * k)
    nbr = np.stack([np.where(k > 0, eid - nx * ny, -1), np.where(j > 0, eid - nx, -1), np.where(i < nx - 1, eid + 1, -1),
                    np.where(j < ny - 1, eid + nx, -1), np.where(i > 0, eid - 1, -1), np.where(k < nz - 1, eid + nx * ny, -1)],
                   axis=1).astype(np.int64)
    m.neighbors = nbr
    m.side_bnd_id = np.where(nbr < 0, np.arange(6, dtype=np.int32)[None, :], -1).astype(np.int32)
    m.bnd_kind = np.full(6, bc, np.int32) if np.isscalar(bc) else np.asarray(bc, np.int32)
    return m


def quad_grid(nx, ny, lo=(0, 0), hi=(1, 1), bc=BC_REFLECT, block_fn=None, name="quad_grid") -> Mesh:
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    J, I = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    xyz = np.stack([xs[I].ravel(), ys[J].ravel(), np.zeros(I.size)], axis=1)
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    i, j = i.ravel(), j.ravel()

    def nid(a, b):
        return a + (nx + 1) * b

    conn = np.stack([nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)], axis=1).astype(np.int64)
    m = Mesh(QUAD4, xyz, conn, np.zeros(conn.shape[0], np.int32), name=name)
    if block_fn is not None:
        m.block_id = np.asarray(block_fn(m.centroids()), np.int32)
    return _finish(m, _box_bnd_ids(m, (lo[0], lo[1], 0), (hi[0], hi[1], 0)), 4, bc)


# Kuhn (Freudenthal) split of a hex cell into 6 tets around the (0,0,0)-(1,1,1) diagonal; conforming
# across translated cells.  Each tet walks 000 -> +e_a -> +e_b -> 111; odd permutations are flipped so
# every tet is positively oriented (node 3 on the outward side of (0,1,2), libMesh Tet4 convention).
_KUHN = []
for perm in [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]:
    pts = [np.zeros(3, int)]
    for ax in perm:
        p = pts[-1].copy()
        p[ax] = 1
        pts.append(p)
    _KUHN.append(np.array(pts))


def tet_grid(nx, ny, nz, lo=(0, 0, 0), hi=(1, 1, 1), bc=BC_REFLECT, block_fn=None, name="tet_grid") -> Mesh:
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    xyz = np.stack([xs[I].ravel(), ys[J].ravel(), zs[K].ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()

    def nid(a, b, c):
        return a + (nx + 1) * (b + c * (ny + 1))

    tets = []
    for pts in _KUHN:
        cols = [nid(i + p[0], j + p[1], k + p[2]) for p in pts]
        tets.append(np.stack(cols, axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int64)  # 6 consecutive tets per cell
    # fix orientation
    p = xyz[conn]
    vol = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    flip = vol < 0
    conn[flip, 1], conn[flip, 2] = conn[flip, 2].copy(), conn[flip, 1].copy()
    m = Mesh(TET4, xyz, conn, np.zeros(conn.shape[0], np.int32), name=name)
    if block_fn is not None:
        m.block_id = np.asarray(block_fn(m.centroids()), np.int32)
    return _finish(m, _box_bnd_ids(m, lo, hi), 6, bc)


def tet_grid_24(nx, ny, nz, lo=(0, 0, 0), hi=(1, 1, 1), bc=BC_REFLECT, name="tet_grid_24") -> Mesh:
    """GeneratedMeshGenerator with elem_type = TET4 as libMesh builds it (MeshTools::Generation::build_cube):
    a HEX27 grid whose every hex is split into 24 tetrahedra -- for each of the six hex sides, four tets
    (side corner k, side centre, side corner k + 1, hex centre); the mid-edge nodes are dropped.  This is the mesh
    of the reference's TET4 trace tests (moose/modules/ray_tracing/test/tests/traceray/tet4/tests: 3 x 3 x 3
    on [0, 5]^3), whose gold total distances tests/test_oracle_pinning.py reproduces on it."""
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    pts, index = [], {}

    def node(p):
        key = (round(p[0], 12), round(p[1], 12), round(p[2], 12))
        if key not in index:
            index[key] = len(pts)
            pts.append(p)
        return index[key]

    conn = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = [(xs[i], ys[j], zs[k]), (xs[i + 1], ys[j], zs[k]), (xs[i + 1], ys[j + 1], zs[k]), (xs[i], ys[j + 1], zs[k]),
                     (xs[i], ys[j], zs[k + 1]), (xs[i + 1], ys[j], zs[k + 1]), (xs[i + 1], ys[j + 1], zs[k + 1]),
                     (xs[i], ys[j + 1], zs[k + 1])]
                c = [np.array(q, float) for q in c]
                apex = node(sum(c) / 8.0)
                for side in SIDE_NODES[HEX8]:
                    sc = node(sum(c[v] for v in side) / 4.0)
                    for t in range(4):
                        conn.append([node(c[side[t]]), sc, node(c[side[(t + 1) % 4]]), apex])
    xyz = np.array(pts, float)
    conn = np.array(conn, np.int64)
    p = xyz[conn]
    vol = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    flip = vol < 0
    conn[flip, 1], conn[flip, 2] = conn[flip, 2].copy(), conn[flip, 1].copy()
    m = Mesh(TET4, xyz, conn, np.zeros(conn.shape[0], np.int32), name=name)
    return _finish(m, _box_bnd_ids(m, lo, hi), 6, bc)


# ---- BASELINE.json configs -------------------------------------------------------------------------
PIN_PITCH = 1.25984  # cm, reference test/tests/study/geometry.xml:8-16
PIN_RADII = (0.39218, 0.40005, 0.45720)  # fuel / gap / clad outer radii


def pin_blocks(cent):
    r = np.hypot(cent[:, 0], cent[:, 1])
    return np.where(r < PIN_RADII[0], 0, np.where(r < PIN_RADII[1], 1, np.where(r < PIN_RADII[2], 2, 3)))


def infinite_medium(n=5) -> Mesh:
    """C1: examples/infinite_medium -- 5x5x5 HEX8 on [-5,5]^3, reflective, centre element = block 1."""
    def blocks(c):
        return (np.abs(c).max(axis=1) < 1.0 + 1e-9).astype(np.int32) if n % 2 == 1 else np.zeros(len(c), np.int32)
    return hex_grid(n, n, n, (-5, -5, -5), (5, 5, 5), BC_REFLECT, blocks, "infinite_medium")


def pin_cell_tet(n_xy=26, n_z=25, height=None) -> Mesh:
    """C2: UO2 pin cell, hex grid split into 6 tets/cell (26*26*25*6 = 101 400 TET4), blocks by radius."""
    h = PIN_PITCH / 2
    height = PIN_PITCH if height is None else height
    return tet_grid(n_xy, n_xy, n_z, (-h, -h, -height / 2), (h, h, height / 2), BC_REFLECT, pin_blocks, "pin_cell_tet")


def pin_cell_hex(n_xy=20, n_z=4) -> Mesh:
    h = PIN_PITCH / 2
    return hex_grid(n_xy, n_xy, n_z, (-h, -h, -h), (h, h, h), BC_REFLECT, pin_blocks, "pin_cell_hex")


def assembly_blocks(n_pins):
    def fn(c):
        x = c[:, 0] / PIN_PITCH
        y = c[:, 1] / PIN_PITCH
        lx = (x - np.floor(x) - 0.5) * PIN_PITCH
        ly = (y - np.floor(y) - 0.5) * PIN_PITCH
        return pin_blocks(np.stack([lx, ly], axis=1))
    return fn


def assembly_quad(n_pins=17, cells_per_pin=9) -> Mesh:
    """C3 (2-D): 17x17 assembly of pin cells, QUAD4 (17*9)^2 = 23 409 elements (reference mesh: 23 876)."""
    L = n_pins * PIN_PITCH
    n = n_pins * cells_per_pin
    return quad_grid(n, n, (0, 0), (L, L), BC_REFLECT, assembly_blocks(n_pins), "assembly_quad")


def assembly_hex(n_pins=17, cells_per_pin=9, n_z=4, height=4.0) -> Mesh:
    """C3 (3-D): the 2-D assembly extruded by 4 layers (reference run_assembly_3d.i:6-31)."""
    L = n_pins * PIN_PITCH
    n = n_pins * cells_per_pin
    return hex_grid(n, n, n_z, (0, 0, 0), (L, L, height), BC_REFLECT, assembly_blocks(n_pins), "assembly_hex")


def core_hex(n_asm=15, cells_per_pin=3, n_layers=17, layer_height=2.0) -> Mesh:
    """C4: a BEAVRS-core-sized synthetic (reference examples/beavrs_core/run_3d_core_cycle1.i: 3-D full core whose
    mesh files are not in the tree, vacuum boundary): n_asm x n_asm assemblies of 17 x 17 pin cells, `cells_per_pin`
    HEX8 cells per pin pitch, `n_layers` axial layers, every outer face killing the ray.
    15 x 15 assemblies x 3 cells x 17 layers = 9 948 825 HEX8."""
    n = n_asm * 17 * cells_per_pin
    L = n_asm * 17 * PIN_PITCH
    return hex_grid(n, n, n_layers, (0, 0, 0), (L, L, n_layers * layer_height), BC_KILL, assembly_blocks(n_asm * 17),
                    f"core_hex_{n_asm}_{cells_per_pin}_{n_layers}")


def brick_tet(n, bc=BC_KILL, checker=4) -> Mesh:
    """C5: n^3 hex brick split into tets (6 n^3 TET4), checkerboard of fuel / water blocks."""
    L = n * 0.05
    def blocks(c):
        q = np.floor(c / (L / checker)).astype(int)
        return np.where((q.sum(axis=1) % 2) == 0, 0, 3).astype(np.int32)
    return tet_grid(n, n, n, (0, 0, 0), (L, L, L), bc, blocks, f"brick_tet_{n}")


def _spread3(v):
    """bits of v (< 2^21) spread to every third position"""
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def renumber_elements(m: Mesh, order="morton") -> Mesh:
    """The same mesh with its elements renumbered along a space-filling curve (Morton order of the centroids): what a
    host hands over when the element records do not fit in L2, so that a history's successive crossings stay within a
    few DRAM pages.  Element ids are labels (libMesh renumbers too): tallies, banks and traces are in the new numbering,
    `m.renumbering[new] = old` maps back.  Physics is unchanged -- every history visits the same elements in the same
    order (tests/test_host_cpu.py)."""
    if order != "morton":
        raise ValueError("order must be 'morton'")
    cent = m.xyz[m.conn].mean(axis=1)
    lo, hi = cent.min(0), cent.max(0)
    q = ((cent - lo) / np.maximum(hi - lo, 1e-300) * (2 ** 21 - 1)).astype(np.uint64)
    code = _spread3(q[:, 0]) | (_spread3(q[:, 1]) << np.uint64(1)) | (_spread3(q[:, 2]) << np.uint64(2))
    perm = np.argsort(code, kind="stable")          # new -> old
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)                # old -> new
    out = Mesh(m.elem_type, m.xyz, m.conn[perm], m.block_id[perm], name=m.name + "_morton")
    if m.neighbors is not None:
        nb = m.neighbors[perm]
        out.neighbors = np.where(nb >= 0, inv[np.maximum(nb, 0)], -1).astype(np.int64)
    if m.side_bnd_id is not None:
        out.side_bnd_id = m.side_bnd_id[perm]
    out.bnd_kind = m.bnd_kind
    out.renumbering = perm
    return out


# BEAVRS assembly blocks -> role (examples/beavrs_assembly/run_assembly.i:39-44 with materials.xml):
# 10004 helium gap, 10006 and 11006 Zircaloy-4 (cladding, guide tubes), 10010 fuel 3.1 %, 10014 borated water
BEAVRS_BLOCK_ROLE = {10004: "gap", 10006: "clad", 10010: "fuel", 10014: "water", 11006: "clad"}


def beavrs_assembly_file(path) -> Mesh:
    """C3 on the reference's own mesh: the .npz written by tests/golden/make_beavrs_mesh.py from
    31enr_16_in.e (23 876 QUAD4, reflecting side sets 1-4)."""
    z = np.load(path)
    xyz = np.zeros((z["xy"].shape[0], 3))
    xyz[:, :2] = z["xy"]
    m = Mesh(QUAD4, xyz, z["conn"].astype(np.int64), z["block_id"].astype(np.int32), name="beavrs_assembly")
    m.neighbors = compute_neighbors(QUAD4, m.conn)
    m.side_bnd_id = z["side_bnd_id"].astype(np.int32)
    m.bnd_kind = z["bnd_kind"].astype(np.int32)
    return m


def extrude_quads(m: Mesh, num_layers, heights, direction=(0, 0, 1), bottom_boundary=None, top_boundary=None) -> Mesh:
    """QUAD4 -> HEX8 extrusion with the numbering of MOOSE's FancyExtruderGenerator
    (moose/framework/src/meshgenerators/FancyExtruderGenerator.C:255-285 nodes, :444-462 Hex8 nodes, :551 element
    ids, :604-646 sides): node id = node + plane * n_nodes, element id = elem + layer * n_elem, quad side s becomes
    hex side s + 1 with its boundary id, hex side 0 of the first layer / side 5 of the last layer get the
    bottom / top boundary ids.  Plane positions accumulate height / num_layers like the generator does."""
    assert m.elem_type == QUAD4
    d = np.asarray(direction, float)
    planes = [np.zeros(3)]
    for nl, h in zip(num_layers, heights):
        for _ in range(int(nl)):
            planes.append(planes[-1] + d * (float(h) / float(nl)))
    L = len(planes) - 1
    nn, ne = m.xyz.shape[0], m.n_elem
    xyz = np.concatenate([m.xyz + p for p in planes])
    conn = np.concatenate([np.concatenate([m.conn + k * nn, m.conn + (k + 1) * nn], axis=1) for k in range(L)])
    out = Mesh(HEX8, xyz, conn.astype(np.int64), np.tile(m.block_id, L), name=m.name + "_extruded")
    out.neighbors = compute_neighbors(HEX8, out.conn)
    sb = np.full((L * ne, 6), -1, np.int32)
    for k in range(L):
        sb[k * ne:(k + 1) * ne, 1:5] = m.side_bnd_id
    n_bnd = m.bnd_kind.size
    if bottom_boundary is not None:
        sb[:ne, 0] = bottom_boundary
        n_bnd = max(n_bnd, bottom_boundary + 1)
    if top_boundary is not None:
        sb[(L - 1) * ne:, 5] = top_boundary
        n_bnd = max(n_bnd, top_boundary + 1)
    out.side_bnd_id = sb
    out.bnd_kind = np.zeros(n_bnd, np.int32)
    out.bnd_kind[:m.bnd_kind.size] = m.bnd_kind
    return out
