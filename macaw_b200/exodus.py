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


RAY_VAR_ORDER = ("aux_energy", "aux_id", "aux_n_progeny", "aux_particle_type", "aux_seed_URR", "aux_seed_source",
                 "aux_seed_tracking", "aux_weight", "intersections", "ray_id")


def ray_segments(traces):
    """Cached trace points (Context.traces(): ordered by ray and sequence) -> the segment mesh RayTracingExodus
    builds (RT src/outputs/RayTracingMeshOutput.C:252-445): one node per cached point, one EDGE2 element between
    consecutive points of a ray, and per element the ray id, the running intersection count and the ray's aux data
    as after that segment (:447-580).  Returns (coord, conn, {var name: values per element})."""
    from .cstructs import TRACE_AUX_NAMES
    t = traces
    coord = np.ascontiguousarray(t["point"], np.float64)
    same_ray = t["ray_id"][1:] == t["ray_id"][:-1]
    a = np.nonzero(same_ray)[0]
    conn = np.stack([a, a + 1], axis=1).astype(np.int64)
    end = t[a + 1]
    first_of_ray = np.concatenate([[0], np.nonzero(~same_ray)[0] + 1])
    ray_start = np.repeat(first_of_ray, np.diff(np.concatenate([first_of_ray, [len(t)]])))
    data = {"ray_id": end["ray_id"].astype(np.float64), "intersections": (a - ray_start[a]).astype(np.float64)}
    for k, name in enumerate(TRACE_AUX_NAMES):
        data["aux_" + name] = np.ascontiguousarray(end["aux"][:, k])
    return coord, conn, data


def write_ray_exodus(path, traces, version=2):
    """The `rays` output of the reference's study test (test/tests/study/openmc_study.i:77-83: RayTracingExodus with
    output_aux_data): EDGE2 segments with element variables, one time step.  Readable by read_ray_exodus and laid
    out like test/tests/study/gold/openmc_study_rays.e (variable names in the gold file's order)."""
    from scipy.io import netcdf_file
    coord, conn, data = ray_segments(traces)
    f = netcdf_file(path, "w", version=version)
    f.title = "macaw_b200 ray trace export"
    f.createDimension("len_string", 33)
    f.createDimension("len_name", 33)
    f.createDimension("time_step", 1)
    f.createDimension("num_dim", 3)
    f.createDimension("num_nodes", coord.shape[0])
    f.createDimension("num_elem", conn.shape[0])
    f.createDimension("num_el_blk", 1)
    f.createDimension("num_el_in_blk1", conn.shape[0])
    f.createDimension("num_nod_per_el1", 2)
    f.createDimension("num_elem_var", len(RAY_VAR_ORDER))
    for d, name in enumerate(("coordx", "coordy", "coordz")):
        v = f.createVariable(name, ">f8", ("num_nodes",))
        v[:] = coord[:, d]
    v = f.createVariable("eb_prop1", ">i4", ("num_el_blk",))
    v[:] = [0]
    v.name = "ID"
    v = f.createVariable("connect1", ">i4", ("num_el_in_blk1", "num_nod_per_el1"))
    v[:] = conn + 1
    v.elem_type = "EDGE2"
    v = f.createVariable("time_whole", ">f8", ("time_step",))
    v[:] = [1.0]
    v = f.createVariable("name_elem_var", "c", ("num_elem_var", "len_name"))
    for i, nm in enumerate(RAY_VAR_ORDER):
        v[i] = list(nm.ljust(33, "\0"))
    for i, nm in enumerate(RAY_VAR_ORDER):
        v = f.createVariable(f"vals_elem_var{i + 1}eb1", ">f8", ("time_step", "num_el_in_blk1"))
        v[0, :] = data[nm]
    f.close()


def read_ray_exodus(path):
    """(coord, conn, {name: element values of the last time step}) of a RayTracingExodus file -- the reference's gold
    file or one written by write_ray_exodus."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    names = ["".join(c.decode() for c in row).strip("\x00 ").strip() for row in v["name_elem_var"][:]]
    coord = np.stack([np.array(v["coordx"][:]), np.array(v["coordy"][:]), np.array(v["coordz"][:])], axis=1)
    conn = np.array(v["connect1"][:], np.int64) - 1
    data = {nm: np.array(v[f"vals_elem_var{i + 1}eb1"][:][-1], np.float64) for i, nm in enumerate(names)}
    f.close()
    return coord, conn, data
