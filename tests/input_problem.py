"""Python restatement of what host/input_file.hpp understands (MOOSE input + OpenMC settings.xml /
materials.xml), used to check the C++ front end: same files in, same mesh / materials / settings out,
and the same problem for the oracle."""
import re
import xml.etree.ElementTree as ET

import numpy as np

from macaw_b200 import cstructs, meshgen, nucdata

N_AVOGADRO, MASS_NEUTRON = 0.602214076, 1.00866491595  # OMC constants.h:81,89


def parse_hit(text):
    """{'A/B': {key: value}} for every block path of a MOOSE input file."""
    out, path = {}, []
    for raw in text.splitlines():
        line = re.sub(r"#.*$", "", raw).strip()
        if not line:
            continue
        m = re.match(r"^\[(.*)\]$", line)
        if m:
            n = m.group(1).strip()
            if n in ("", "../", ".."):
                depth = path.pop()
                continue
            parts = [p for p in n.lstrip("./").split("/") if p] if not n.startswith("./") else [n[2:]]
            path.append(parts)
            out.setdefault("/".join(sum(path, [])), {})
            continue
        k, v = line.split("=", 1)
        out["/".join(sum(path, []))][k.strip()] = v.strip().strip("'\"").strip()
    return out


def read_input(input_path, xml_dir, lib):
    """-> mesh, library (nuclides in order of first use), blocks, material indices, settings, batches, tallies, aux."""
    import os
    from macaw_b200 import exodus
    hit = parse_hit(open(input_path).read())
    bcs = [(v["type"], v["boundary"].split()) for k, v in hit.items() if k.startswith("RayBCs/")]
    fm = [v for v in hit.values() if v.get("type") == "FileMeshGenerator"]
    if fm:
        path = os.path.join(os.path.dirname(input_path), fm[0]["file"])
        mesh = exodus.read_exodus(path, reflect=[int(b) for t, bl in bcs if t == "ReflectRayBC" for b in bl],
                                  kill=[int(b) for t, bl in bcs if t == "KillRayBC" for b in bl])
        for v in hit.values():  # generators in file order
            if v.get("type") == "TransformGenerator":
                mesh.xyz += np.array(v["vector_value"].split(), float)
            elif v.get("type") == "FancyExtruderGenerator":
                kinds = mesh.bnd_kind
                mesh = meshgen.extrude_quads(mesh, [int(x) for x in v["num_layers"].split()],
                                             [float(x) for x in v["heights"].split()],
                                             [float(x) for x in v["direction"].split()],
                                             int(v["bottom_boundary"]) if "bottom_boundary" in v else None,
                                             int(v["top_boundary"]) if "top_boundary" in v else None)
                mesh.bnd_kind[:kinds.size] = kinds
                for t, bl in bcs:  # ids created by the extruder
                    for b in bl:
                        if int(b) < mesh.bnd_kind.size:
                            mesh.bnd_kind[int(b)] = meshgen.BC_REFLECT if t == "ReflectRayBC" else meshgen.BC_KILL
        return _finish_input(hit, mesh, mesh.dim, xml_dir, lib)
    g = next(v for k, v in hit.items() if v.get("type") == "GeneratedMeshGenerator")
    dim = int(g["dim"])
    lo = [float(g.get(a, 0)) for a in ("xmin", "ymin", "zmin")]
    hi = [float(g.get(a, 1)) for a in ("xmax", "ymax", "zmax")]
    boxes = [v for v in hit.values() if v.get("type") == "SubdomainBoundingBoxGenerator"]

    def blocks_of(c):
        b = np.zeros(len(c), np.int32)
        for bx in boxes:
            a, t = np.array(bx["bottom_left"].split(), float), np.array(bx["top_right"].split(), float)
            inside = np.all((c[:, :dim] >= a[:dim]) & (c[:, :dim] <= t[:dim]), axis=1)
            b[inside] = int(bx["block_id"])
        return b
    if dim == 3:
        mesh = meshgen.hex_grid(int(g["nx"]), int(g["ny"]), int(g["nz"]), lo, hi, meshgen.BC_KILL, blocks_of)
        names = ["back", "bottom", "right", "top", "left", "front"]
    else:
        mesh = meshgen.quad_grid(int(g["nx"]), int(g["ny"]), lo[:2], hi[:2], meshgen.BC_KILL, blocks_of)
        names = ["bottom", "right", "top", "left"]
    kinds = np.zeros(len(names), np.int32)
    for k, v in hit.items():
        if k.startswith("RayBCs/"):
            for b in v["boundary"].split():
                kinds[names.index(b)] = meshgen.BC_REFLECT if v["type"] == "ReflectRayBC" else meshgen.BC_KILL
    mesh.bnd_kind = kinds
    return _finish_input(hit, mesh, dim, xml_dir, lib)


def _finish_input(hit, mesh, dim, xml_dir, lib):
    ck = next(v for v in hit.values() if v.get("type") == "CollisionKernel")
    T0 = float(hit.get("Variables/temperature", hit.get("AuxVariables/temperature", {})).get("initial_condition", 0))
    try:
        T0 = float(ck.get("temperature"))  # a constant instead of a coupled variable (run_assembly_3d.i:62)
    except (TypeError, ValueError):
        pass

    st = ET.parse(f"{xml_dir}/settings.xml").getroot()
    box = [float(x) for x in st.find("source/space/parameters").text.split()]
    s = cstructs.make_settings(int(st.findtext("particles")), n_inactive=int(st.findtext("inactive", "0")),
                               is_2d=dim == 2, z_coord=float(ck.get("z_coord", 0)), source_box=box,
                               only_fissionable=st.find("source/space").get("type") == "fission",
                               seed=int(st.findtext("seed", "1")), record_traces=0)
    batches = int(hit.get("Executioner", {}).get("num_steps", st.findtext("batches")))

    by_name = {n.name: n for n in lib.nuclides}
    order, mats, index_of_id = [], [], {}
    for m in ET.parse(f"{xml_dir}/materials.xml").getroot().findall("material"):
        nucs = m.findall("nuclide")
        for n in nucs:
            if n.get("name") not in order:
                order.append(n.get("name"))
        awr = np.array([by_name[n.get("name")].awr for n in nucs])
        frac = np.array([float(n.get("ao")) if n.get("ao") is not None else -float(n.get("wo")) for n in nucs])
        if frac[0] < 0:
            frac = -frac / awr
        total = 0.0
        for f in frac:
            total += f
        d = m.find("density")
        units = d.get("units")
        density = total if units == "sum" else float(d.get("value")) * (1 if units == "atom/b-cm" else -1)  # g/cm3 | g/cc
        frac = frac / total
        if density < 0:
            sp = 0.0
            for f, a in zip(frac, awr):
                sp += f * a
            density = -density * N_AVOGADRO / MASS_NEUTRON * (1.0 / sp)
        iso = m.findtext("isotropic")
        mats.append(nucdata.Material(m.get("name"), [order.index(n.get("name")) for n in nucs], list(frac * density),
                                     [int(n.get("name") in iso.split()) for n in nucs] if iso else None))
        index_of_id[int(m.get("id"))] = len(mats) - 1
    library = nucdata.Library([by_name[n] for n in order], mats, lib.n_log_bins)
    blocks = [int(b) for b in ck["blocks"].split()]
    mat_index = [index_of_id[int(i)] for i in ck["materials"].split()]
    def active(path):
        parent, _, child = path.rpartition("/")
        p = hit.get(parent, {})
        if "active" in p and child not in p["active"].split():
            return False
        return child not in p.get("inactive", "").split()
    tallies = []
    for k, v in hit.items():
        if v.get("type") == "OpenMCTally" and active(k):
            filters = v["filters"].split()
            gran = "cell" if "cell" in filters else "universe" if "universe" in filters else "global"
            tallies.append((int(v["id"]), cstructs.ESTIMATOR[v.get("estimator", "COLLISION").lower()],
                            cstructs.GRANULARITY[gran], [cstructs.SCORE[x] for x in v["scores"].split()],
                            [float(x) for x in v["energy_bins"].split()] if "energy" in filters else None))
    aux = [v for k, v in hit.items() if v.get("type") == "OpenMCTallyAux" and active(k)]
    return mesh, library, blocks, mat_index, s, T0, batches, tallies, aux
