"""Reader for Exodus-II meshes stored as classic NetCDF (what MaCaw's FileMeshGenerator inputs are, e.g.
examples/beavrs_assembly/31enr_16_in.e): coordinates, element blocks (QUAD4 / TET4 / HEX8), block ids
and side sets -> meshgen.Mesh.  scipy's netcdf_file reads the container; no libMesh / netCDF4 needed.
Host-side set-up code; the hot path never sees it.

Exodus numbers nodes, elements and sides from 1; its QUAD4 / HEX8 / TET4 node and side orders are the
libMesh ones (libMesh's ExodusII_IO conversion tables are the identity for these three types)."""
from __future__ import annotations

import numpy as np

from . import meshgen

_TYPES = {"QUAD4": meshgen.QUAD4, "QUAD": meshgen.QUAD4, "TETRA4": meshgen.TET4, "TETRA": meshgen.TET4,
          "TET4": meshgen.TET4, "HEX8": meshgen.HEX8, "HEX": meshgen.HEX8}


def read_exodus(path, reflect=(), kill=()) -> meshgen.Mesh:
    """`reflect` / `kill`: side-set ids carrying ReflectRayBC / KillRayBC (boundary ids of the input's
    [RayBCs] block); boundary sides in no listed set get no id (the ray is killed there)."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    dim = f.dimensions["num_dim"]
    xyz = np.zeros((f.dimensions["num_nodes"], 3))
    if "coord" in v:
        xyz[:, :dim] = np.asarray(v["coord"][:]).T
    else:
        for d, name in enumerate(("coordx", "coordy", "coordz")[:dim]):
            xyz[:, d] = v[name][:]
    conns, blocks, etype = [], [], None
    ids = np.asarray(v["eb_prop1"][:])
    for b in range(f.dimensions["num_el_blk"]):
        c = v[f"connect{b + 1}"]
        t = _TYPES.get(c.elem_type.decode().strip().upper())
        if t is None:
            raise ValueError(f"element type {c.elem_type!r} is not traced by this path (QUAD4, TET4, HEX8)")
        if etype is not None and t != etype:
            raise ValueError("mixed element types in one mesh")
        etype = t
        conns.append(np.asarray(c[:], np.int64) - 1)
        blocks.append(np.full(c.shape[0], ids[b], np.int32))
    conn, block_id = np.concatenate(conns), np.concatenate(blocks)
    m = meshgen.Mesh(etype, xyz, conn, block_id, name=str(path))
    m.neighbors = meshgen.compute_neighbors(etype, conn)
    sb = np.full(m.neighbors.shape, -1, np.int32)
    ss_ids = np.asarray(v["ss_prop1"][:]) if "ss_prop1" in v else np.zeros(0, int)
    wanted = {int(i): meshgen.BC_REFLECT for i in reflect}
    wanted.update({int(i): meshgen.BC_KILL for i in kill})
    for k, sid in enumerate(ss_ids):
        if int(sid) not in wanted:
            continue
        e = np.asarray(v[f"elem_ss{k + 1}"][:], np.int64) - 1
        s = np.asarray(v[f"side_ss{k + 1}"][:], np.int64) - 1
        on_boundary = m.neighbors[e, s] < 0
        sb[e[on_boundary], s[on_boundary]] = int(sid)
    m.side_bnd_id = sb
    n_bnd = (max(wanted) + 1) if wanted else 0
    m.bnd_kind = np.zeros(n_bnd, np.int32)
    for sid, kind in wanted.items():
        m.bnd_kind[sid] = kind
    f.close()
    return m


def write_exodus(path, m: meshgen.Mesh, version=2):
    """Minimal Exodus-II writer (classic NetCDF, 64-bit offsets): coordinates, one element block per
    subdomain id, one side set per boundary id present on the mesh.  Enough for read_exodus, the C++
    reader (host/exodus_reader.hpp) and libMesh's ExodusII_IO; used to hand generated meshes to the
    FileMeshGenerator path."""
    from scipy.io import netcdf_file
    names = {meshgen.QUAD4: "QUAD4", meshgen.TET4: "TETRA4", meshgen.HEX8: "HEX8"}
    ids = m.block_ids_sorted()
    bnd = np.unique(m.side_bnd_id[m.side_bnd_id >= 0])
    f = netcdf_file(path, "w", version=version)  # 1 = classic (32-bit offsets), 2 = 64-bit offsets
    f.title = "macaw_b200 mesh export"
    f.createDimension("len_string", 33)
    f.createDimension("num_dim", m.dim)
    f.createDimension("num_nodes", m.xyz.shape[0])
    f.createDimension("num_elem", m.n_elem)
    f.createDimension("num_el_blk", ids.size)
    if bnd.size:
        f.createDimension("num_side_sets", bnd.size)
    for d, name in enumerate(("coordx", "coordy", "coordz")[:m.dim]):
        v = f.createVariable(name, ">f8", ("num_nodes",))
        v[:] = m.xyz[:, d]
    v = f.createVariable("eb_prop1", ">i4", ("num_el_blk",))
    v[:] = ids
    v.name = "ID"
    order = []
    for b, bid in enumerate(ids):
        elems = np.nonzero(m.block_id == bid)[0]
        order.append(elems)
        f.createDimension(f"num_el_in_blk{b + 1}", elems.size)
        f.createDimension(f"num_nod_per_el{b + 1}", m.conn.shape[1])
        v = f.createVariable(f"connect{b + 1}", ">i4", (f"num_el_in_blk{b + 1}", f"num_nod_per_el{b + 1}"))
        v[:] = m.conn[elems] + 1
        v.elem_type = names[m.elem_type]
    # Exodus numbers elements block by block
    new_id = np.empty(m.n_elem, np.int64)
    new_id[np.concatenate(order)] = np.arange(m.n_elem)
    if bnd.size:
        v = f.createVariable("ss_prop1", ">i4", ("num_side_sets",))
        v[:] = bnd
        v.name = "ID"
        for k, sid in enumerate(bnd):
            e, sd = np.nonzero(m.side_bnd_id == sid)
            f.createDimension(f"num_side_ss{k + 1}", e.size)
            ve = f.createVariable(f"elem_ss{k + 1}", ">i4", (f"num_side_ss{k + 1}",))
            ve[:] = new_id[e] + 1
            vs = f.createVariable(f"side_ss{k + 1}", ">i4", (f"num_side_ss{k + 1}",))
            vs[:] = sd + 1
    f.close()
    return np.concatenate(order)  # old element index of every element in file order
