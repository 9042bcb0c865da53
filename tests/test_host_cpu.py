"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), the work split / exchange plan arithmetic, and the world_size-2 bank
synchronisation (gloo) driven through the same macaw_b200.dist code the GPU study runs over NCCL,
with the CPU oracle as the per-rank engine."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "macaw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcw_[a-z_0-9]+)\s*\(", src)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from macaw_b200 import lib
    L = lib.load()
    names = _header_functions()
    assert len(names) >= 39
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == names  # the ctypes binding covers the whole header
    assert L.mcw_version() >= 100


def test_c_abi_struct_sizes_match_ctypes_mirror():
    from macaw_b200 import cstructs, lib
    L = lib.load()
    out = (C.c_int32 * 6)()
    assert L.mcw_struct_sizes(out, 6) == 6
    assert out[0] == 96 == cstructs.SITE_DTYPE.itemsize  # openmc::SourceSite, particle_data.h:43-53
    assert out[3] == C.sizeof(cstructs.Settings)
    assert out[4] == C.sizeof(cstructs.GenerationResult)
    assert out[5] == cstructs.RECORD_DTYPE.itemsize


def test_library_validation_without_a_device():
    """mcw_validate_library runs mcw_set_library's own checks and copies on the host (no handle, no GPU): every library
    the tests and the bench use is accepted -- including the nucleus descriptor size the ctypes mirror assumes --, and
    malformed data are refused with the reference's kind of message instead of reaching a kernel."""
    import copy
    from macaw_b200 import lib, nucdata
    from tests import problems
    L = lib.load()
    out = (C.c_int32 * 6)()
    L.mcw_struct_sizes(out, 6)
    assert out[1] == C.sizeof(nucdata.NuclideDesc)
    msg = C.create_string_buffer(512)

    def check(library):
        nd, _, keep = nucdata.pack_library(library)
        rc = L.mcw_validate_library(C.addressof(nd), len(library.nuclides), msg, 512)
        sets, keep_s = nucdata.pack_temperature_sets(library)  # further temperature sets: one descriptor each
        for _, d in sets:
            rc = rc or L.mcw_validate_library(C.addressof(d), 1, msg, 512)
        return rc, msg.value.decode()
    for name, make in problems.PROBLEMS.items():
        _, library, _, _ = make(n_particles=100)
        rc, text = check(library)
        assert rc == 0, (name, text)
    base = nucdata.add_reactions(nucdata.make_library(15, G=300, tabular=True))
    assert check(base)[0] == 0
    # a reaction whose threshold lies outside the grid
    bad = copy.deepcopy(base)
    bad.nuclides[2].more_reactions[1].threshold = 299
    bad.nuclides[2].more_reactions[1].xs = np.zeros(1)
    rc, text = check(bad)
    assert rc != 0 and "threshold" in text
    # a law the library does not implement
    bad = copy.deepcopy(base)
    bad.nuclides[2].more_reactions[0].law = 9
    rc, text = check(bad)
    assert rc != 0 and "not implemented" in text
    # N-body with six bodies (OMC secondary_nbody.cpp:57)
    bad = copy.deepcopy(base)
    r = [x for n in bad.nuclides for x in (n.more_reactions or []) if x.law == 4][0]
    r.nbody = (6, r.nbody[1], r.nbody[2])
    rc, text = check(bad)
    assert rc != 0 and "N-body" in text
    # fission products: a tabulated function with logarithmic interpolation, a product missing, an unknown spectrum
    full = nucdata.add_fission_data(nucdata.add_evaporation_laws(copy.deepcopy(base)))
    assert check(full)[0] == 0
    bad = copy.deepcopy(full)
    bad.nuclides[2].total_nu.interpolation = 5
    rc, text = check(bad)
    assert rc != 0 and "histogram and linear-linear" in text
    bad = copy.deepcopy(full)
    bad.nuclides[2].fission_products = bad.nuclides[2].fission_products[:-1]
    rc, text = check(bad)
    assert rc != 0 and "1 + n_precursor" in text
    bad = copy.deepcopy(full)
    bad.nuclides[2].fission_products[1].law = 4
    rc, text = check(bad)
    assert rc != 0 and "energy law not implemented" in text
    # an evaporation spectrum on a scattering reaction without its theta(E)
    bad = copy.deepcopy(full)
    r = [x for n in bad.nuclides for x in (n.more_reactions or []) if x.law == 6][0]
    r.a = None
    rc, text = check(bad)
    assert rc != 0 and "function it needs is missing" in text
    # an unsorted energy grid
    bad = copy.deepcopy(base)
    bad.nuclides[0].energy = bad.nuclides[0].energy[::-1].copy()
    rc, text = check(bad)
    assert rc != 0 and "not sorted" in text


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device mcw_create must fail (MCW_E_CUDA); nothing routes to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from macaw_b200.context import Context
    from macaw_b200.lib import McwError
    with pytest.raises(McwError):
        Context(0)


@pytest.mark.parametrize("n,p", [(10, 1), (10, 3), (7, 8), (200000, 8), (1000003, 4)])
def test_calculate_work_matches_openmc_rule(oracle_bridge, n, p):
    from macaw_b200 import dist, lib
    wi = dist.calculate_work(n, p)
    assert wi[0] == 0 and wi[-1] == n
    sizes = np.diff(wi)
    assert sizes.max() - sizes.min() <= 1 and np.all(sizes[:-1] >= sizes[1:])  # first n%p ranks get +1
    assert np.array_equal(wi, lib.calculate_work(n, p))
    o = np.zeros(p + 1, np.int64)
    oracle_bridge.lib().orc_calculate_work(n, p, o.ctypes.data)
    assert np.array_equal(wi, o)


def test_exchange_plan_conserves_and_orders_sites():
    from macaw_b200 import dist
    rng = np.random.default_rng(3)
    for _ in range(50):
        p = int(rng.integers(1, 9))
        n = int(rng.integers(p, 5000))
        wi = dist.calculate_work(n, p)
        cuts = np.sort(rng.integers(0, n + 1, p - 1))
        counts = np.diff(np.concatenate([[0], cuts, [n]]))
        plan = dist.exchange_plan(counts, wi)
        off = dist.exclusive_offsets(counts)
        got = np.full(n, -1, np.int64)
        for s, d, so, do, c in plan:
            assert c > 0 and so + c <= counts[s] and do + c <= wi[d + 1] - wi[d]
            got[wi[d] + do: wi[d] + do + c] = np.arange(off[s] + so, off[s] + so + c)
        assert np.array_equal(got, np.arange(n))  # global order of the sampled bank is preserved


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,name", [(2, "pin_cell_tet"), (3, "pin_cell_tet"),
                                        # temperature field + data sets at several temperatures, fission products, and
                                        # multiplicities that move the weights: the shards still add up to the one-rank run
                                        (2, "pin_cell_tet_temperatures"), (3, "pin_cell_tet_evaporation")])
def test_bank_synchronisation_world_size_n_equals_single_rank(oracle_bridge, tmp_path, world, name):
    """OpenMC's synchronize_bank is rank-count invariant: the next source bank and k_generation of an
    N-rank run equal the 1-rank run's (OMC eigenvalue.cpp:71-311).  Runs N gloo processes."""
    from tests import problems
    n_particles, n_batches = 600, 3
    port = _free_port()
    out = str(tmp_path / "w")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), str(r), str(world),
                               str(port), out, str(n_particles), str(n_batches), name], cwd=ROOT,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)

    mesh, lib, bm, s = problems.PROBLEMS[name](n_particles=n_particles)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, True, False)
    o.sample_initial_source()
    ks = []
    for b in range(1, n_batches + 1):
        r = o.run_generation(b)
        ks.append([r.k_generation, r.keff, r.k_collision, r.k_absorption, r.k_tracklength])
    ks = np.asarray(ks)
    parts = [np.load(f"{out}.{r}.npz") for r in range(world)]
    src = np.concatenate([p["source"] for p in parts])
    assert src.size == n_particles
    assert src.tobytes() == o.source_bank().tobytes()  # bit-identical next-generation source
    for p in parts:
        assert np.allclose(p["ks"], ks, rtol=1e-12, atol=0)  # summation order differs across ranks only
    _, t_sum, _, _ = o.get_tally(1)
    assert np.allclose(sum(p["tally_sum"] for p in parts), t_sum, rtol=1e-11, atol=1e-300)


def test_exodus_reader_and_beavrs_mesh_fixture(golden_dir):
    """macaw_b200/exodus.py on the reference's BEAVRS assembly mesh (when the reference tree is here) must
    reproduce the committed fixture; the fixture itself must be a closed, positively oriented QUAD4 mesh
    of the published size with every boundary side in one of the reflecting side sets 1-4."""
    import numpy as np
    from macaw_b200 import exodus, meshgen
    path = os.path.join(golden_dir, "beavrs_assembly_mesh.npz")
    m = meshgen.beavrs_assembly_file(path)
    assert m.n_elem == 23876 and m.xyz.shape[0] == 24081  # C3
    ids, counts = np.unique(m.block_id, return_counts=True)
    assert ids.tolist() == [10004, 10006, 10010, 10014, 11006] and counts.tolist() == [4224, 4224, 8448, 6580, 400]
    p = m.xyz[m.conn]
    area = 0.5 * ((p[:, 1, 0] - p[:, 3, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 3, 1]))
    assert area.min() > 0 and abs(area.sum() - (17 * 1.25984) ** 2) < 1e-6
    boundary = m.neighbors < 0
    assert boundary.sum() == 408 and set(np.unique(m.side_bnd_id[boundary]).tolist()) == {1, 2, 3, 4}
    assert (m.side_bnd_id[~boundary] == -1).all() and m.bnd_kind.tolist() == [0, 1, 1, 1, 1]
    # neighbour symmetry: the neighbour of my side s has me as a neighbour
    e, s = np.nonzero(~boundary)
    nb = m.neighbors[e, s]
    assert (m.neighbors[nb] == e[:, None]).any(axis=1).all()
    ref = "/root/reference/examples/beavrs_assembly/31enr_16_in.e"
    if os.path.exists(ref):
        r = exodus.read_exodus(ref, reflect=(1, 2, 3, 4))
        assert np.array_equal(r.xyz, m.xyz) and np.array_equal(r.conn, m.conn) and np.array_equal(r.block_id, m.block_id)
        assert np.array_equal(r.side_bnd_id, m.side_bnd_id) and np.array_equal(r.bnd_kind, m.bnd_kind)


def test_extrusion_reproduces_generated_hex_mesh():
    """meshgen.extrude_quads (MOOSE FancyExtruderGenerator numbering, FancyExtruderGenerator.C:255-285,444-462,551,
    604-646) of a generated QUAD4 grid is the generated HEX8 grid: same connectivity, neighbours and lateral
    boundary ids shifted by one side, bottom / top ids on hex sides 0 / 5."""
    import numpy as np
    from macaw_b200 import meshgen
    q = meshgen.quad_grid(3, 2, (0, 0), (3, 2))
    h = meshgen.extrude_quads(q, [2, 1], [1.0, 0.5], (0, 0, 1), 8, 9)
    g = meshgen.hex_grid(3, 2, 3, (0, 0, 0), (3, 2, 1.5))
    assert np.array_equal(h.conn, g.conn) and np.array_equal(h.neighbors, g.neighbors)
    assert np.allclose(h.xyz, g.xyz) and np.array_equal(h.block_id, g.block_id)
    assert np.array_equal(h.side_bnd_id[:, 1:5], np.tile(q.side_bnd_id, (3, 1)))
    assert (h.side_bnd_id[:6, 0] == 8).all() and (h.side_bnd_id[12:, 5] == 9).all()
    assert (h.side_bnd_id[6:, 0] == -1).all() and (h.side_bnd_id[:12, 5] == -1).all()
    # lateral ids: quad bottom / right / top / left = 0 / 1 / 2 / 3 sit where the hex grid has bottom / right / top / left = 1 / 2 / 3 / 4
    lat = g.side_bnd_id[:, 1:5]
    assert np.array_equal(np.where(lat >= 0, lat - 1, -1), h.side_bnd_id[:, 1:5])


def test_tally_nuclide_bins_follow_the_reference():
    """OpenMCTally `nuclides` (MC OpenMCTally.C:188-204) and OpenMCTallyAux's score index (OpenMCTallyAux.C:93-119),
    host logic only: a list gets the total appended; `all` keeps the reference's n + 2 bins (its nuclides_ starts with the
    default -1, which it never clears), of which the last is never scored; unknown names fail with OpenMC's message."""
    from macaw_b200 import nucdata, study
    lib = nucdata.make_library(9, G=50)
    names = [n.name for n in lib.nuclides]
    t = study.OpenMCTally(1, ["flux", "total"], ["cell"], nuclides=[names[3], names[0]])
    assert t.nuclide_indices(lib) == [3, 0, -1] and t.ref_nuclides == [3, 0, -1]
    t = study.OpenMCTally(2, ["flux"], ["cell"], nuclides=["all"])
    assert t.nuclide_indices(lib) == list(range(9)) + [-1, -2] and t.ref_nuclides == [-1] + list(range(9)) + [-1]
    t = study.OpenMCTally(3, ["flux"], ["cell"])
    assert t.nuclide_indices(lib) == [-1]
    with pytest.raises(ValueError, match="Could not find the nuclide Unobtainium"):
        study.OpenMCTally(4, ["flux"], ["cell"], nuclides=["Unobtainium"]).nuclide_indices(lib)

    class FakeStudy:
        library = lib
    t = study.OpenMCTally(1, ["flux", "total"], ["cell"], nuclides=[names[3], names[0]])
    t.nuclide_bins = t.nuclide_indices(lib)
    aux = study.OpenMCTallyAux(FakeStudy(), 1, "total", tally=t, nuclide=names[0])
    assert aux.score_index() == 1 * 3 + 1      # score_bin * n_nuclide_bins + nuclide_bin, as the reference computes it
    assert study.OpenMCTallyAux(FakeStudy(), 1, "flux", tally=t).score_index() == 2   # "all" -> the total bin


def test_morton_renumbering_leaves_the_physics_unchanged():
    """meshgen.renumber_elements (Morton order of the centroids, for meshes larger than L2): element ids are labels --
    the oracle walks the renumbered mesh through the same elements in the same order, so seeds, energies, counts and
    fates of every history are bit-identical and the per-element tally is the same array permuted."""
    from macaw_b200 import cstructs, meshgen, nucdata
    from oracle import bridge
    from tests import problems
    bridge.build(ref=False)
    m = meshgen.brick_tet(5, bc=meshgen.BC_REFLECT)
    m2 = meshgen.renumber_elements(m)
    assert sorted(m2.renumbering.tolist()) == list(range(m.n_elem)) and not np.array_equal(m2.renumbering, np.arange(m.n_elem))
    lib = nucdata.make_library(9, G=200)
    s = cstructs.make_settings(300, source_box=(0, 0, 0, 0.25, 0.25, 0.25))
    out = []
    for mesh in (m, m2):
        o = problems.configure(bridge.Oracle(), mesh, lib, [0, 1, 2, 3], s, tallies=True)
        o.sample_initial_source()
        for b in (1, 2, 3):
            o.run_generation(b)
        out.append((o.history_records().copy(), o.get_tally(1)[1].reshape(m.n_elem, -1).copy()))
    (a, ta), (b, tb) = out
    for f in ("seed_tracking", "seed_source", "n_segments", "n_collisions", "n_progeny", "fate"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["E"].view(np.uint64), b["E"].view(np.uint64))
    assert np.array_equal(m2.renumbering[b["last_elem"]], a["last_elem"])
    assert np.array_equal(tb, ta[m2.renumbering])
