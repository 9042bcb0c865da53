"""The C++ host layer (host/: OpenMCStudy / CollisionKernel / OpenMCTally / OpenMCTallyAux over the
C ABI + the mcw_run driver).  CPU: it builds, reads the problem file the Python side wrote, rejects
bad parameters with the reference's messages and fails loudly without a GPU.  GPU: its k per batch
and its tally read-back equal the oracle's on the same problem file."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from macaw_b200 import cstructs, problem_io
from tests import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
TALLIES = [(1, 0, 0, problems.ALL_SCORES, problems.EBINS), (2, 0, 1, [0, 3], None)]


def _build():
    from macaw_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", HOST, "NCCL=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(HOST, "mcw_run")


def _write(tmp_path, name, **kw):
    mesh, lib, bm, s = problems.PROBLEMS[name](**kw)
    path = str(tmp_path / f"{name}.bin")
    problem_io.write_problem(path, mesh, lib, list(range(len(bm))), [m + 1 for m in bm], s, TALLIES)
    return path, (mesh, lib, bm, s)


def test_problem_file_round_trip(tmp_path):
    path, (mesh, lib, bm, s) = _write(tmp_path, "pin_cell_tet", n_particles=100)
    r = problem_io.read_records(path)
    assert np.array_equal(r["mesh.conn"].reshape(mesh.conn.shape), mesh.conn)
    assert r["mesh.xyz"].tobytes() == np.ascontiguousarray(mesh.xyz, np.float64).tobytes()
    assert r["lib.counts"].tolist() == [len(lib.nuclides), len(lib.materials), lib.n_log_bins]
    assert r["nuc3.xs"].tobytes() == np.ascontiguousarray(lib.nuclides[3].xs, np.float64).tobytes()
    assert r["settings.ints"][0] == 100 and r["kernel.materials"].tolist() == [1, 2, 3, 4]


@pytest.mark.parametrize("name", ["pin_cell_tet_tabular", "pin_cell_tet_reactions", "pin_cell_tet_evaporation", "pin_cell_tet_fission",
                                  "pin_cell_tet_temperatures", "pin_cell_tet_sab", "pin_cell_tet_sab_temperatures", "pin_cell_tet_urr"])
def test_problem_files_carry_the_full_physics(tmp_path, name):
    """Tabulated distributions, thermal-scattering tables with their material assignments and probability tables travel
    through the problem file: the records are there, and the C++ reader gets as far as creating the device handle
    (which fails loudly here: no GPU) -- i.e. it parsed them."""
    import torch
    exe = _build()
    path, (mesh, lib, bm, s) = _write(tmp_path, name, n_particles=100)
    r = problem_io.read_records(path)
    if name.endswith("tabular"):
        assert any(k.endswith("el_angle.xpc") for k in r) and any(k.endswith("inel_energy.e") for k in r)
        assert any(k.endswith("law.i") for k in r) and any(k.endswith("law.f") for k in r)  # correlated, Kalbach-Mann, N-body
    if name.endswith("reactions"):
        assert r["nuc1.rx.head"][0] == 4 and r["nuc1.rx3.i"][0] == 875 and any(k.endswith("corr.f") for k in r)
    if name.endswith("evaporation"):
        assert "nuc1.rx.a" in r and "nuc1.rx2.y" in r and r["nuc1.law.i"][0] == 6 and r["nuc1.rx1.i"][2] == 6
    if name.endswith("fission") or name == "pin_cell_tet_temperatures":
        assert r["nuc2.fis.head"][0] == 7 and "nuc2.fis0.y" in r and any(k.endswith("en.xpc") and ".fis" in k for k in r)
    if name == "pin_cell_tet_temperatures":
        n = len(lib.nuclides)
        assert r["lib.set_of"].size == len(lib.temperature_sets()) > 10 and f"nuc{n}.energy" in r
        assert r["kernel.temperature"].size == mesh.n_elem
    if name.endswith("sab_temperatures"):
        assert r["lib.n_thermal"][0] == 9 and r["lib.sab_same"].tolist() == [0, 1, 2, 0, 0, 1, 1, 2, 2] and r["lib.sab_kT"].size == 9
    if name.endswith("sab"):
        assert r["lib.n_thermal"][0] == 3 and "mat3.sab_i" in r and r["sab0.i_offset"].size > 2
    if name.endswith("urr"):
        assert r["lib.n_urr"][0] == len(lib.urr_tables) and r["urr0.table"].size == lib.urr_tables[0].table.size
    out = subprocess.run([exe, path, "--batches", "1"], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert out.returncode == 1 and "no CPU fallback" in out.stderr


def test_cpp_host_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build()
    path, _ = _write(tmp_path, "infinite_medium", n_particles=100)
    r = subprocess.run([exe, path, "--batches", "1"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and r.stdout == ""
    r = subprocess.run([exe, str(tmp_path / "missing.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open problem file" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["infinite_medium", "pin_cell_tet", "assembly_quad", "pin_cell_tet_iso",
                                  # tabulated distributions, S(a,b) and probability tables through the problem file and
                                  # host/macaw_host.hpp's Library (mcw_set_thermal_tables / _material_thermal / _urr_tables)
                                  "pin_cell_tet_tabular", "pin_cell_tet_reactions", "pin_cell_tet_evaporation", "pin_cell_tet_fission",
                                  "pin_cell_tet_temperatures", "pin_cell_tet_sab", "pin_cell_tet_sab_temperatures", "pin_cell_tet_urr"])
def test_cpp_host_matches_oracle(oracle_bridge, tmp_path, name):
    exe = _build()
    path, (mesh, lib, bm, s) = _write(tmp_path, name, n_particles=3000)
    r = subprocess.run([exe, path, "--batches", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.splitlines()]
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, True, False)
    o.sample_initial_source()
    for b in (1, 2, 3):
        ro, rc = o.run_generation(b), lines[b - 1]
        assert rc["n_segments"] == ro.n_segments and rc["n_collisions"] == ro.n_collisions
        assert rc["n_bank_sites"] == ro.n_bank_sites
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength"):
            assert abs(rc[f] - getattr(ro, f)) <= 1e-11 * abs(getattr(ro, f)), f
    _, so, _, _ = o.get_tally(1)
    ns = mesh.n_elem
    flux = so.reshape(ns, len(problems.EBINS) - 1, len(problems.ALL_SCORES))[:, :, 0].sum(axis=1)
    t1 = [t for t in lines[-1]["tallies"] if t["id"] == 1][0]
    assert t1["n_spatial"] == ns
    assert abs(t1["sum"] - flux.sum()) <= 1e-10 * abs(flux.sum())
    assert abs(t1["first"] - flux[0]) <= 1e-10 * max(abs(flux[0]), 1e-300)


# ---- the input-file surface: MOOSE .i + OpenMC XML through host/input_file.hpp
EXAMPLE = os.path.join(ROOT, "examples", "study_5x5")


def _tables(tmp_path, G=400):
    from macaw_b200 import nucdata
    lib = nucdata.make_library(9, G=G)
    path = str(tmp_path / "tables.bin")
    problem_io.write_library(path, lib)
    return path, lib


def test_input_file_front_end_understands_the_example(tmp_path):
    """mcw_run --input ... --check (no GPU): mesh, boundary kinds, block map, OpenMC density
    normalisation (ao / wo / g/cm3 / atom/b-cm), <isotropic> lists and settings equal the Python
    restatement that reads the same three files."""
    from macaw_b200 import inputs as input_problem
    exe = _build()
    tables, lib = _tables(tmp_path)
    r = subprocess.run([exe, "--input", os.path.join(EXAMPLE, "study.i"), "--tables", tables, "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(
        os.path.join(EXAMPLE, "study.i"), EXAMPLE, lib)
    assert d["n_elem"] == mesh.n_elem == 25 and d["elem_type"] == mesh.elem_type
    assert np.array_equal(np.array(d["xyz"]).reshape(-1, 3), mesh.xyz)
    assert np.array_equal(np.array(d["conn"]).reshape(mesh.conn.shape), mesh.conn)
    assert d["block_id"] == mesh.block_id.tolist() and sum(d["block_id"]) == 1
    assert np.array_equal(np.array(d["side_bnd_id"]).reshape(mesh.side_bnd_id.shape), mesh.side_bnd_id)
    assert d["bnd_kind"] == mesh.bnd_kind.tolist() == [1] * 6
    assert d["blocks"] == blocks and [m - 1 for m in d["materials"]] == mat_index
    assert d["batches"] == batches == 4 and d["n_particles"] == s.n_particles and d["n_inactive"] == s.n_inactive
    assert d["temperature"] == T0 and d["source_box"] == [s.source_box[i] for i in range(6)]
    assert len(d["library_materials"]) == len(library.materials) == 4 and d["n_library_nuclides"] == len(library.nuclides)
    for a, b in zip(d["library_materials"], library.materials):
        assert a["nuclide"] == b.nuclides
        assert a["atom_density"] == b.densities  # same IEEE operations in the same order: exact
        assert a["isotropic"] == (b.isotropic or [])
    assert d["tallies"] == len(tallies) == 1 and d["aux"] == len(aux) == 1


def test_input_file_front_end_rejects_what_the_reference_rejects(tmp_path):
    exe = _build()
    tables, _ = _tables(tmp_path)
    text = open(os.path.join(EXAMPLE, "study.i")).read()
    cases = [(text.replace('materials = "4 1"', 'materials = "4 9"'), "is not in materials.xml"),
             (text.replace("type = CollisionKernel", "type = SomethingElse"), "no CollisionKernel"),
             (text.replace("dim = 3", "dim = 2"), "z_coord parameter must be specified for 2D"),
             (text.replace("GeneratedMeshGenerator", "FileMeshGenerator"), "a file mesh's side sets are addressed by id"),
             (text.replace("GeneratedMeshGenerator", "FileMeshGenerator\n    file = nowhere.e").replace(
                 "'back front top right left bottom'", "'1 2'"), "cannot open mesh file")]
    for k, (bad, msg) in enumerate(cases):
        d = tmp_path / f"case{k}"
        d.mkdir()
        (d / "study.i").write_text(bad)
        for f in ("settings.xml", "materials.xml"):
            (d / f).write_text(open(os.path.join(EXAMPLE, f)).read())
        r = subprocess.run([exe, "--input", str(d / "study.i"), "--tables", tables, "--check"], capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (msg, r.stderr)


@pytest.mark.gpu
def test_input_file_run_matches_oracle(oracle_bridge, tmp_path):
    """The example input through the C++ host on the GPU == the oracle on the Python restatement of
    the same files: k per batch, counters, and the OpenMCTallyAux values of every element."""
    from macaw_b200 import inputs as input_problem
    exe = _build()
    tables, lib = _tables(tmp_path)
    r = subprocess.run([exe, "--input", os.path.join(EXAMPLE, "study.i"), "--tables", tables], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.splitlines()]
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(
        os.path.join(EXAMPLE, "study.i"), EXAMPLE, lib)
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh, np.asarray(blocks, np.int32))
    o.set_library(library, mat_index)
    o.set_settings(s)
    for t in tallies:
        o.add_tally(*t)
    o.sample_initial_source()
    for b in range(1, batches + 1):
        ro, rc = o.run_generation(b), lines[b - 1]
        assert rc["n_segments"] == ro.n_segments and rc["n_collisions"] == ro.n_collisions
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength"):
            assert abs(rc[f] - getattr(ro, f)) <= 1e-11 * abs(getattr(ro, f)), f
    _, so, _, nr = o.get_tally(7)
    want = so.reshape(mesh.n_elem, 2, 2)[:, 0, 0]  # energy_bin = 0 (thermal), score flux
    got = np.array([a for a in lines if "aux" in a][0]["aux"][0]["values"])
    assert nr == batches - s.n_inactive and got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-10, atol=1e-300) and want.sum() > 0


# ---- FileMeshGenerator decks: Exodus mesh through host/exodus_reader.hpp
def _file_mesh_deck(tmp_path, version=2, name="assembly_file"):
    """examples/assembly_file with its mesh written from the committed fixture of the reference's mesh."""
    import shutil
    from macaw_b200 import exodus, meshgen
    d = tmp_path / name
    shutil.copytree(os.path.join(ROOT, "examples", "assembly_file"), d, ignore=shutil.ignore_patterns("*.e"))
    exodus.write_exodus(str(d / "assembly.e"),
                        meshgen.beavrs_assembly_file(os.path.join(ROOT, "tests", "golden", "beavrs_assembly_mesh.npz")),
                        version=version)
    return str(d)


def _check_against_restatement(d, mesh, library, blocks, mat_index, s, T0, batches):
    assert d["n_elem"] == mesh.n_elem and d["elem_type"] == mesh.elem_type and d["n_nodes"] == mesh.xyz.shape[0]
    assert np.array_equal(np.array(d["xyz"]).reshape(-1, 3), mesh.xyz)
    assert np.array_equal(np.array(d["conn"]).reshape(mesh.conn.shape), mesh.conn)
    assert d["block_id"] == mesh.block_id.tolist()
    assert np.array_equal(np.array(d["side_bnd_id"]).reshape(mesh.side_bnd_id.shape), mesh.side_bnd_id)
    assert d["bnd_kind"] == mesh.bnd_kind.tolist()
    assert d["blocks"] == blocks and [m - 1 for m in d["materials"]] == mat_index
    assert d["batches"] == batches and d["n_particles"] == s.n_particles and d["n_inactive"] == s.n_inactive
    assert d["temperature"] == T0 and d["source_box"] == [s.source_box[i] for i in range(6)]
    assert d["only_fissionable"] == s.source_only_fissionable
    assert len(d["library_materials"]) == len(library.materials) and d["n_library_nuclides"] == len(library.nuclides)
    for a, b in zip(d["library_materials"], library.materials):
        assert a["nuclide"] == b.nuclides and a["atom_density"] == b.densities and a["isotropic"] == (b.isotropic or [])


def test_file_mesh_deck_front_end(tmp_path):
    """FileMeshGenerator + TransformGenerator + numeric side-set boundaries + [AuxVariables] temperature +
    g/cc densities + <sab> note: the C++ front end (own NetCDF reader) equals the Python restatement
    (scipy's reader) on the same deck -- every coordinate, side, block and atom density."""
    from macaw_b200 import inputs as input_problem
    exe = _build()
    tables, lib = _tables(tmp_path)
    deck = _file_mesh_deck(tmp_path)
    r = subprocess.run([exe, "--input", os.path.join(deck, "run.i"), "--tables", tables, "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "S(alpha,beta) table c_H_in_H2O ignored" in r.stderr
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(os.path.join(deck, "run.i"), deck, lib)
    _check_against_restatement(json.loads(r.stdout), mesh, library, blocks, mat_index, s, T0, batches)
    assert mesh.n_elem == 23876 and batches == 3 and T0 == 300.0
    # the same mesh as a CDF-1 file (32-bit offsets): the C++ reader handles both header variants
    deck1 = _file_mesh_deck(tmp_path, version=1, name="assembly_file_cdf1")
    assert open(os.path.join(deck1, "assembly.e"), "rb").read(4) == b"CDF\x01"
    r1 = subprocess.run([exe, "--input", os.path.join(deck1, "run.i"), "--tables", tables, "--check"], capture_output=True, text=True)
    assert r1.returncode == 0, r1.stderr
    assert json.loads(r1.stdout) == json.loads(r.stdout)


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples/beavrs_assembly"), reason="reference tree not present")
def test_reference_beavrs_deck_is_understood_unmodified(tmp_path):
    """MaCaw's own examples/beavrs_assembly (run_assembly.i, settings.xml, materials.xml with 91 nuclides in 17
    materials, 31enr_16_in.e) read in place, unmodified; nuclear data = one synthetic table per nuclide name."""
    import xml.etree.ElementTree as ET
    from macaw_b200 import nucdata
    from macaw_b200 import inputs as input_problem
    exe = _build()
    deck = "/root/reference/examples/beavrs_assembly"
    names = []
    for n in ET.parse(os.path.join(deck, "materials.xml")).getroot().iter("nuclide"):
        if n.get("name") not in names:
            names.append(n.get("name"))
    lib = nucdata.library_for_names(names, G=200)
    tables = str(tmp_path / "beavrs_tables.bin")
    problem_io.write_library(tables, lib)
    r = subprocess.run([exe, "--input", os.path.join(deck, "run_assembly.i"), "--tables", tables, "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(
        os.path.join(deck, "run_assembly.i"), deck, lib)
    d = json.loads(r.stdout)
    _check_against_restatement(d, mesh, library, blocks, mat_index, s, T0, batches)
    assert d["n_elem"] == 23876 and d["n_particles"] == 200000 and d["batches"] == 5 and len(names) == 91
    assert len(library.materials) == 17 and blocks == [10000, 10001, 10004, 10006, 10008, 10009, 10010, 10013, 10014, 11006]


@pytest.mark.gpu
def test_file_mesh_deck_run_matches_oracle(oracle_bridge, tmp_path):
    from macaw_b200 import inputs as input_problem
    exe = _build()
    tables, lib = _tables(tmp_path)
    deck = _file_mesh_deck(tmp_path)
    r = subprocess.run([exe, "--input", os.path.join(deck, "run.i"), "--tables", tables], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.splitlines()]
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(os.path.join(deck, "run.i"), deck, lib)
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh, np.asarray(blocks, np.int32))
    o.set_library(library, mat_index)
    o.set_settings(s)
    for t in tallies:
        o.add_tally(*t)
    o.sample_initial_source()
    for b in range(1, batches + 1):
        ro, rc = o.run_generation(b), lines[b - 1]
        assert rc["n_segments"] == ro.n_segments and rc["n_collisions"] == ro.n_collisions
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength"):
            assert abs(rc[f] - getattr(ro, f)) <= 1e-11 * abs(getattr(ro, f)), f
    _, so, _, nr = o.get_tally(1)
    got = np.array([a for a in lines if "aux" in a][0]["aux"][0]["values"])
    assert nr == batches and got.shape == so.shape
    assert np.allclose(got, so, rtol=1e-10, atol=1e-300) and so.sum() > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/tests/study"), reason="reference tree not present")
def test_reference_study_test_deck_is_understood_unmodified(tmp_path):
    """MaCaw's own regression input test/tests/study/openmc_study.i with its settings.xml / materials.xml
    (30 particles, 3 batches, every nuclide <isotropic>), read in place."""
    import xml.etree.ElementTree as ET
    from macaw_b200 import nucdata
    from macaw_b200 import inputs as input_problem
    exe = _build()
    deck = "/root/reference/test/tests/study"
    names = []
    for n in ET.parse(os.path.join(deck, "materials.xml")).getroot().iter("nuclide"):
        if n.get("name") not in names:
            names.append(n.get("name"))
    lib = nucdata.library_for_names(names, G=200)
    tables = str(tmp_path / "study_tables.bin")
    problem_io.write_library(tables, lib)
    r = subprocess.run([exe, "--input", os.path.join(deck, "openmc_study.i"), "--tables", tables, "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(
        os.path.join(deck, "openmc_study.i"), deck, lib)
    d = json.loads(r.stdout)
    _check_against_restatement(d, mesh, library, blocks, mat_index, s, T0, batches)
    assert d["n_elem"] == 25 and d["n_particles"] == 30 and d["batches"] == 3 and d["n_inactive"] == 1
    assert sum(d["block_id"]) == 1 and d["block_id"][12] == 1  # the centre element of the 5 x 5 x 1 grid
    assert all(all(m["isotropic"]) for m in d["library_materials"])  # every nuclide isotropic-in-lab


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples/beavrs_assembly"), reason="reference tree not present")
def test_reference_beavrs_3d_deck_is_understood_unmodified(oracle_bridge, tmp_path):
    """run_assembly_3d.i: the same Exodus mesh translated and extruded by FancyExtruderGenerator into the
    95 504 HEX8 of SURVEY section 8 (C3 3-D), side sets 1-4 plus the extruder's 8 / 9 reflecting, a constant
    temperature, `inactive` tallies.  The C++ front end equals the Python restatement, and the oracle walks
    the extruded mesh without a failed trace."""
    import xml.etree.ElementTree as ET
    from macaw_b200 import nucdata
    from macaw_b200 import inputs as input_problem
    exe = _build()
    deck = "/root/reference/examples/beavrs_assembly"
    names = []
    for n in ET.parse(os.path.join(deck, "materials.xml")).getroot().iter("nuclide"):
        if n.get("name") not in names:
            names.append(n.get("name"))
    lib = nucdata.library_for_names(names, G=200)
    tables = str(tmp_path / "beavrs_tables.bin")
    problem_io.write_library(tables, lib)
    r = subprocess.run([exe, "--input", os.path.join(deck, "run_assembly_3d.i"), "--tables", tables, "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mesh, library, blocks, mat_index, s, T0, batches, tallies, aux = input_problem.read_input(
        os.path.join(deck, "run_assembly_3d.i"), deck, lib)
    d = json.loads(r.stdout)
    _check_against_restatement(d, mesh, library, blocks, mat_index, s, T0, batches)
    assert d["n_elem"] == 95504 and d["elem_type"] == 2 and d["bnd_kind"] == [0, 1, 1, 1, 1, 0, 0, 0, 1, 1]
    assert d["tallies"] == len(tallies) == 1 and d["aux"] == len(aux) == 1 and T0 == 300.0
    z = np.unique(mesh.xyz[:, 2])
    assert z.tolist() == [-100.0, 0.0, 100.0, 200.0, 300.0]
    s.n_particles = 400
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh, np.asarray(blocks, np.int32))
    o.set_library(library, mat_index)
    o.set_settings(s)
    o.sample_initial_source()
    res = o.run_generation(1, 4)
    assert res.n_failed_traces == 0 and res.n_segments > 100 * 400 and 0.1 < res.k_generation < 2.0


def test_hit_syntax_variants(tmp_path):
    """The older MOOSE block syntax ([./name] ... [../]), a value quoted over several lines, comments after
    values and a block given twice (MOOSE merges them) parse to the same problem as the plain deck."""
    exe = _build()
    tables, _ = _tables(tmp_path)
    text = open(os.path.join(EXAMPLE, "study.i")).read()
    alt = text.replace("  [add_subdomain]", "  [./add_subdomain]").replace("    block_name = 'center'\n  []", "    block_name = 'center'\n  [../]")
    alt = alt.replace("boundary = 'back front top right left bottom'", "boundary = 'back front top\n                right left bottom'  # split over two lines")
    alt = alt.replace("[Executioner]\n  type = Transient\n  num_steps = 4\n[]", "[Executioner]\n  type = Transient\n[]\n[Executioner]\n  num_steps = 4 # merged\n[]")
    assert alt != text and "[../]" in alt
    outs = []
    for k, t in enumerate((text, alt)):
        d = tmp_path / f"variant{k}"
        d.mkdir()
        (d / "study.i").write_text(t)
        for f in ("settings.xml", "materials.xml"):
            (d / f).write_text(open(os.path.join(EXAMPLE, f)).read())
        r = subprocess.run([exe, "--input", str(d / "study.i"), "--tables", tables, "--check"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(json.loads(r.stdout))
    assert outs[0] == outs[1] and outs[0]["batches"] == 4 and outs[0]["bnd_kind"] == [1] * 6


@pytest.mark.gpu
def test_python_and_cpp_hosts_agree_on_a_deck(tmp_path):
    """examples/study_5x5 through macaw_b200.inputs.run_deck (ctypes host) and through host/mcw_run (C++ host):
    same counters, k within the FP64 addition-order tolerance, same per-element tally read-back."""
    from macaw_b200 import inputs
    exe = _build()
    tables, lib = _tables(tmp_path)
    deck = os.path.join(EXAMPLE, "study.i")
    r = subprocess.run([exe, "--input", deck, "--tables", tables], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.splitlines()]
    study, res = inputs.run_deck(deck, lib=lib, verbose=False)
    assert len(res) == 4
    for b, rp in enumerate(res):
        rc = lines[b]
        assert rc["n_segments"] == rp.n_segments and rc["n_collisions"] == rp.n_collisions
        assert rc["n_bank_sites"] == rp.n_bank_sites
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength"):
            assert abs(rc[f] - getattr(rp, f)) <= 1e-11 * abs(getattr(rp, f)), f
    t, _, nr = study.ctx.get_tally(7)
    got = np.array([a for a in lines if "aux" in a][0]["aux"][0]["values"])
    assert nr == 3 and np.allclose(got, t[:, 0, 0], rtol=1e-10, atol=1e-300)
    study.ctx.close()


@pytest.mark.parametrize("name", ["pin_cell_tet", "pin_cell_tet_tabular", "pin_cell_tet_reactions", "pin_cell_tet_evaporation",
                                  "pin_cell_tet_fission", "pin_cell_tet_temperatures", "pin_cell_tet_sab_temperatures",
                                  "pin_cell_tet_urr", "beavrs_assembly_full"])
def test_cpp_host_hands_over_the_same_library_as_the_python_host(tmp_path, name):
    """No GPU needed: `mcw_run problem.bin --digest` reads the problem file with host/problem_io.hpp, builds the descriptors
    with host/macaw_host.hpp, has the library check them (mcw_validate_library) and prints the library's own digest of
    them (mcw_digest: every scalar and every array the descriptors point to).  The Python host's descriptors of the same
    problem (macaw_b200/nucdata.py pack_*) give the same four digests: nuclides, further temperature sets, S(a,b) tables,
    probability tables -- so both hosts hand the library the same data, bit for bit, for every kind of physics."""
    from macaw_b200 import lib as mlib, nucdata
    exe = _build()
    path, (mesh, lib, bm, s) = _write(tmp_path, name, n_particles=100)
    out = subprocess.run([exe, path, "--digest"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout)
    L = mlib.load()

    def digest(kind, addr, n):
        h = C.c_uint64()
        assert L.mcw_digest(kind, addr, n, C.byref(h)) == 0
        return f"{h.value:016x}"
    nd, _, keep = nucdata.pack_library(lib)
    assert got["nuclides"] == digest(0, C.addressof(nd), len(lib.nuclides))
    sets, keep_s = nucdata.pack_temperature_sets(lib)
    if sets:
        arr = (nucdata.NuclideDesc * len(sets))()
        for k, (_, d) in enumerate(sets):
            C.memmove(C.addressof(arr[k]), C.addressof(d), C.sizeof(nucdata.NuclideDesc))
        assert got["temperature_sets"] == digest(0, C.addressof(arr), len(sets))
    else:
        assert got["temperature_sets"] == digest(0, C.addressof(nd), 0)
    if getattr(lib, "thermal_tables", None):
        tables, _, _ = lib.thermal_entries()
        td, keep_t = nucdata.pack_thermal(tables)
        assert got["thermal"] == digest(1, C.addressof(td), len(tables))
    if getattr(lib, "urr_tables", None):
        ud, keep_u = nucdata.pack_urr(lib.urr_tables)
        assert got["urr"] == digest(2, C.addressof(ud), len(lib.urr_tables))
