"""The reference's input-file surface for the Python host: a MOOSE input (hit syntax) next to OpenMC's
settings.xml / materials.xml -> mesh, library, CollisionKernel map, settings, tallies.  The same decks
are read by the C++ front end (host/input_file.hpp); tests/test_cpp_host.py checks the two against each
other field by field (exact coordinates, sides and atom densities), so each is the other's restatement.

  python -m macaw_b200.inputs --input examples/study_5x5/study.i [--grid-points G]    # runs the deck on cuda:0

Mesh generators: GeneratedMeshGenerator + SubdomainBoundingBoxGenerator, FileMeshGenerator (Exodus-II via
macaw_b200/exodus.py) + TransformGenerator TRANSLATE + FancyExtruderGenerator.  Set-up code only."""
import re
import xml.etree.ElementTree as ET

import numpy as np

from . import cstructs, meshgen, nucdata

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
    from . import exodus
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


def nuclide_names(materials_xml):
    names = []
    for n in ET.parse(materials_xml).getroot().iter("nuclide"):
        if n.get("name") not in names:
            names.append(n.get("name"))
    return names


def run_deck(input_path, xml_dir=None, grid_points=2000, device=0, batches=None, verbose=True, lib=None, rays=None,
             particles=None):
    """Run a deck on the GPU through the Python host objects (study.py); nuclear data = one synthetic table
    per nuclide name of materials.xml (nucdata.library_for_names).  Returns the per-batch results."""
    import os
    from . import study as st
    xml_dir = xml_dir or os.path.dirname(os.path.abspath(input_path))
    if lib is None:
        lib = nucdata.library_for_names(nuclide_names(os.path.join(xml_dir, "materials.xml")), G=grid_points)
    mesh, library, blocks, mat_index, s, T0, n_batches, tallies, aux = read_input(input_path, xml_dir, lib)
    if particles:
        s.n_particles = int(particles)
    if rays:
        s.record_traces = 2  # [Outputs] RayTracingExodus with output_aux_data (test/tests/study/openmc_study.i:77-83)
    kernel = st.CollisionKernel(blocks=blocks, materials=[m + 1 for m in mat_index], temperature=T0,
                                z_coord=s.z_coord if mesh.dim == 2 else None)
    inv_score = {v: k for k, v in cstructs.SCORE.items()}
    tl = []
    for tid, est, gran, scores, ebins in tallies:
        filters = [{0: "cell", 1: "universe"}[gran]] if gran in (0, 1) else []
        tl.append(st.OpenMCTally(id=tid, scores=[inv_score[x] for x in scores], filters=filters + (["energy"] if ebins else []),
                                 estimator="COLLISION" if est == 0 else "TRACKLENGTH", energy_bins=ebins))
    study = st.OpenMCStudy(mesh, library, kernel, s, tl, device=device)
    out = []
    for _ in range(batches or n_batches):
        r = study.execute()
        out.append(r)
        if verbose:
            print(f"Keff {r.keff} ({r.keff_std}) Generation: {r.k_generation}")  # MC OpenMCStudy.C:399-401
    if rays:  # execute_on = final: the cached traces of the last batch
        from . import exodus
        exodus.write_ray_exodus(rays, study.ctx.traces())
    return study, out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="run a MOOSE / OpenMC deck on cuda:0 with synthetic nuclear data")
    ap.add_argument("--input", required=True)
    ap.add_argument("--xml-dir")
    ap.add_argument("--grid-points", type=int, default=2000)
    ap.add_argument("--batches", type=int)
    ap.add_argument("--particles", type=int)
    ap.add_argument("--rays", help="write the last batch's ray traces as an Exodus EDGE2 file (RayTracingExodus layout)")
    a = ap.parse_args()
    run_deck(a.input, a.xml_dir, a.grid_points, batches=a.batches, rays=a.rays, particles=a.particles)
