"""GPU <-> oracle parity through the C ABI (run on the B200 box: pytest -m gpu).

Bar: bit-exact for every integer / index / byte result (seeds, element sequences, bank contents are
compared as raw bytes, which also makes the FP64 fields of histories and sites bit-exact); FP64
sums that depend on atomic ordering (tallies, k totals) within rtol 1e-11."""
import ctypes as C

import numpy as np
import pytest

from macaw_b200 import cstructs
from tests import lots_rays, problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_factory():
    from macaw_b200.context import Context
    made = []

    def make():
        c = Context(0)
        made.append(c)
        return c
    yield make
    for c in made:
        c.close()


def both(oracle_bridge, ctx_factory, name, tallies=True, track_tally=False, nuclide_tallies=False, **kw):
    mesh, lib, bm, s = problems.PROBLEMS[name](**kw)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies, track_tally, nuclide_tallies)
    g = problems.configure(ctx_factory(), mesh, lib, bm, s, tallies, track_tally, nuclide_tallies)
    return o, g, mesh, lib, s


@pytest.mark.parametrize("name", ["infinite_medium", "pin_cell_tet", "assembly_quad", "pin_cell_tet_sab", "beavrs_assembly_sab"])
def test_xs_lookup_bit_exact(oracle_bridge, ctx_factory, name):
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=False)
    E, mat = problems.lookup_points(lib, 20000)
    a, b = o.xs_lookup(E, mat), g.xs_lookup(E, mat)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    for i in range(len(lib.nuclides)):
        assert np.array_equal(o.grid_index(i, lib.n_log_bins), g.grid_index(i, lib.n_log_bins))


@pytest.mark.parametrize("name,n", [("infinite_medium", 70001), ("pin_cell_tet", 300), ("assembly_quad", 200000)])
def test_xs_lookup_sorted_staged_bit_exact(oracle_bridge, ctx_factory, name, n):
    """The sorted, shared-memory-staged lookup (mcw_xs.cuh) against the oracle: library-sorted
    (mode 2), caller-sorted (mode 1), and mode 1 fed UNSORTED arrays -- staging is an optimisation,
    never a correctness assumption.  Ragged sizes: partial tiles, tiles that straddle materials."""
    import torch
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=False)
    E, mat = problems.lookup_points(lib, n)
    want = o.xs_lookup(E, mat).view(np.uint64)
    dE, dm = torch.from_numpy(E).cuda(), torch.from_numpy(mat).cuda()
    out = torch.zeros(n * 4, dtype=torch.float64, device="cuda")
    for plan in (1, 0):  # k_xs_lookup_plan (all windows of a pass staged at once) / k_xs_lookup_staged (per nuclide)
        g.set_option("xs_plan", plan)
        for mode in (0, 2, 1):
            out.zero_()
            g.xs_lookup_device(n, dE.data_ptr(), dm.data_ptr(), out.data_ptr(), 1, mode)
            assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(n, 4), want), (plan, mode)
    order = np.lexsort((E, mat))
    dE, dm = torch.from_numpy(E[order].copy()).cuda(), torch.from_numpy(mat[order].copy()).cuda()
    out.zero_()
    g.xs_lookup_device(n, dE.data_ptr(), dm.data_ptr(), out.data_ptr(), 1, 1)
    assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(n, 4), want[order])


def test_inlined_division_equals_the_operator(ctx_factory):
    """mcw_div_fast (the compiler's expansion of a / b, written out for instruction-level parallelism in the tile
    lookup) against the plain operator and against IEEE division on the host: wherever its range checks hold the
    quotient is used as is and must be the correctly rounded one; elsewhere the kernels fall back to the operator."""
    g = ctx_factory()
    rng = np.random.default_rng(11)
    n = 400000
    a = np.concatenate([rng.uniform(0, 1, n // 2) * 10.0 ** rng.uniform(-12, 8, n // 2), rng.standard_normal(n // 4),
                        np.ldexp(rng.uniform(0.5, 1, n // 4), rng.integers(-1070, 1020, n // 4))])
    b = np.concatenate([rng.uniform(0, 1, n // 2) * 10.0 ** rng.uniform(-12, 8, n // 2) + 1e-300, rng.standard_normal(n // 4),
                        np.ldexp(rng.uniform(0.5, 1, n // 4), rng.integers(-1070, 1020, n // 4))])
    a[:8] = [0.0, 1.0, -1.0, 5e-324, 1e308, np.inf, np.nan, 3.0]
    b[:8] = [1.0, 0.0, 3.0, 2.0, 1e-308, 2.0, 1.0, 3.0]
    qf, qp, ok = g.test_division(a, b)
    with np.errstate(all="ignore"):
        want = a / b
    same = (qp.view(np.uint64) == want.view(np.uint64)) | (np.isnan(qp) & np.isnan(want))
    assert same.all()  # the device operator is IEEE division
    used = ok != 0
    assert used.mean() > 0.7  # the energy-grid operands of the lookup are all in this class
    assert np.array_equal(qf[used].view(np.uint64), qp[used].view(np.uint64))


@pytest.mark.parametrize("name", ["infinite_medium", "pin_cell_tet", "pin_cell_hex", "assembly_quad"])
def test_locate_and_source_bit_exact(oracle_bridge, ctx_factory, name):
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=False)
    rng = np.random.default_rng(2)
    lo, hi = mesh.xyz.min(0), mesh.xyz.max(0)
    pts = lo + rng.uniform(0, 1, (5000, 3)) * (hi - lo)
    pts[:200] = mesh.xyz[rng.integers(0, mesh.xyz.shape[0], 200)]  # exactly on vertices: tie-break rule
    rid = np.arange(5000, dtype=np.int64)
    assert np.array_equal(o.locate_points(pts, rid), g.locate_points(pts, rid))
    o.sample_initial_source()
    g.sample_initial_source()
    assert o.source_bank().tobytes() == g.source_bank().tobytes()


@pytest.mark.parametrize("et,kind", [("hex8", "vtv"), ("hex8", "ctv"), ("hex8", "ete"), ("hex8", "ctc"),
                                     ("quad4", "vtv"), ("quad4", "ctv"), ("quad4", "ctc"), ("tet4", "vtv"),
                                     ("tet4", "ctv"), ("tet4", "ete"), ("tet4", "ctc"), ("tet4k", "vtv"),
                                     ("tet4k", "ctv"), ("tet4k", "ete")])
def test_general_walk_matches_oracle_and_moose_gold(oracle_bridge, ctx_factory, golden_dir, et, kind):
    import json, os
    from macaw_b200 import meshgen, nucdata
    if et == "hex8":
        mesh = meshgen.hex_grid(5, 5, 5, (0, 0, 0), (5, 5, 5), meshgen.BC_KILL)
    elif et == "quad4":
        mesh = meshgen.quad_grid(5, 5, (0, 0), (5, 5), meshgen.BC_KILL)
    elif et == "tet4":
        # libMesh's own TET4 GeneratedMesh (24 tets per hex): the mesh of the reference's tet4 trace tests
        mesh = meshgen.tet_grid_24(3, 3, 3, (0, 0, 0), (5, 5, 5), meshgen.BC_KILL)
    else:
        mesh = meshgen.tet_grid(4, 4, 4, (0, 0, 0), (4, 4, 4), meshgen.BC_KILL)  # Kuhn split (the bench meshes)
    lib = nucdata.make_library(2, G=50)
    s = cstructs.make_settings(100)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, [0], s, tallies=False)
    g = problems.configure(ctx_factory(), mesh, lib, [0], s, tallies=False)
    start, dirs, elems, sides, maxd = lots_rays.lots(mesh, kind)
    md = maxd if kind == "ctc" else None
    d0, s0, h0, st = o.trace_rays(start, dirs, elems, sides, md)
    d1, s1, h1 = g.trace_rays(start, dirs, elems, sides, md)
    assert st["failed"] == 0
    assert np.array_equal(s0, s1) and np.array_equal(h0, h1)
    assert np.array_equal(d0.view(np.uint64), d1.view(np.uint64))
    gold_names = {"vtv": "ntn" if et == "hex8" else "vtv", "ctv": "ctn" if et == "hex8" else "ctv", "ete": "ete",
                  "ctc": "ctc"}
    if et != "tet4k":  # every element type has the reference's own gold total distance (TET4 on libMesh's 24-split)
        name = gold_names[kind] if et == "hex8" else kind
        gold = json.load(open(os.path.join(golden_dir, "rt_lots_gold.json")))[f"{et}/{name}"]
        assert abs(d1.sum() - gold["total_distance"]) <= 1e-8 * gold["total_distance"]
    if kind != "ctc":  # the reference's analytic check (LotsOfRaysRayStudy.C:117-256): distance to the bounding box
        lo, hi = mesh.xyz.min(0), mesh.xyz.max(0)
        exp = lots_rays.box_exit_distance(start, dirs, lo, hi, mesh.dim)
        assert np.allclose(d1, exp, rtol=1e-9, atol=1e-9)


# Event-loop regimes (mcw_set_option): EVENT = one round per collision throughout (tail_threshold 0),
# FUSED = every history keeps its lane through its collisions (the default for these small banks),
# MIXED = event rounds until fewer than 700 histories fly, then fused; SORTED = event rounds whose
# cross-section refresh goes through the energy binning + tile kernel (mcw_xs.cuh).
EVENT = {"tail_threshold": 0}
FUSED = {"tail_threshold": 1 << 30}
MIXED = {"tail_threshold": 700, "collide_min": 3}
SORTED = {"tail_threshold": 300, "xs_sort_min": 1}
# PUMP = the whole generation in the persistent kernel (mcw_pump.cuh: tracking / collision / general-walk warps
# connected by device-side rings); PUMP_TAIL = event rounds until fewer than 700 histories fly, then the pump
PUMP = {"pump_threshold": 1 << 30}
PUMP_TAIL = {"pump_threshold": 700, "tail_threshold": 0}


@pytest.mark.parametrize("name,kw,opts", [
    ("infinite_medium", {}, EVENT), ("infinite_medium", {}, FUSED), ("infinite_medium", {"sort_xs_queue": 1}, SORTED),
    ("pin_cell_tet", {}, EVENT), ("pin_cell_tet", {}, MIXED), ("pin_cell_tet", {"sort_xs_queue": 1}, SORTED),
    ("pin_cell_hex", {}, EVENT), ("pin_cell_hex", {}, FUSED),
    ("assembly_quad", {}, MIXED), ("assembly_quad", {"sort_xs_queue": 1}, SORTED),
    ("pin_cell_tet", {"sort_xs_queue": 1}, {"tail_threshold": 0, "xs_sort_min": 1, "ordered_queues": 0}),
    ("pin_cell_tet", {"seed_through_double": False}, MIXED),
    ("infinite_medium_iso", {}, EVENT), ("pin_cell_tet_iso", {}, MIXED),
    # the reference's own BEAVRS assembly mesh (31enr_16_in.e, raw block ids 10004 ... 11006), 35 nuclides
    ("beavrs_assembly", {}, MIXED), ("beavrs_assembly", {"sort_xs_queue": 1}, SORTED),
    # 20 000 histories: host syncs every 4th round, so the general-walk queue is served only after syncs
    ("pin_cell_tet", {"n_particles": 20000}, {"tail_threshold": 3000}),
    ("pin_cell_tet", {"n_particles": 20000}, {"tail_threshold": 3000, "fallback_every_round": 1}),
    ("pin_cell_hex", {"n_particles": 20000}, {"tail_threshold": 0, "sync_every": 7}),
    ("infinite_medium", {}, PUMP), ("pin_cell_tet", {}, PUMP), ("pin_cell_hex", {}, PUMP), ("assembly_quad", {}, PUMP_TAIL),
    ("beavrs_assembly", {}, PUMP), ("pin_cell_tet_iso", {}, PUMP_TAIL),
    ("pin_cell_tet", {"n_particles": 20000}, PUMP), ("pin_cell_tet", {"n_particles": 20000}, {"pump_threshold": 5000}),
    ("pin_cell_tet", {"n_particles": 20000}, {"pump_threshold": 1 << 30, "pump_track_warps": 1, "pump_patience": 1}),
    ("pin_cell_tet", {"n_particles": 20000}, {"pump_threshold": 1 << 30, "pump_track_warps": 3, "pump_patience": 64}),
    # the kernel build with the rarely used branches compiled in (csrc/mcw_kernels.h) on problems that do not need it:
    # same results as the common build
    ("pin_cell_tet", {}, {**MIXED, "force_rare": 1}), ("beavrs_assembly", {"sort_xs_queue": 1}, {**SORTED, "force_rare": 1}),
    ("pin_cell_hex", {}, {**EVENT, "force_rare": 1}), ("assembly_quad", {}, {**PUMP_TAIL, "force_rare": 1}),
    # tabulated secondary distributions (AngleDistribution, ContinuousTabular in CM and LAB)
    ("pin_cell_tet_tabular", {}, EVENT), ("pin_cell_tet_tabular", {}, FUSED), ("pin_cell_tet_tabular", {}, PUMP),
    ("pin_cell_tet_tabular", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # lists of scattering reactions with yields (further levels, continuum with a non-integral multiplicity, (n,2n))
    ("pin_cell_tet_reactions", {}, EVENT), ("pin_cell_tet_reactions", {}, FUSED), ("pin_cell_tet_reactions", {}, PUMP),
    ("pin_cell_tet_reactions", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # Maxwell / evaporation / Watt spectra and energy-dependent multiplicities on scattering reactions
    ("pin_cell_tet_evaporation", {}, EVENT), ("pin_cell_tet_evaporation", {}, FUSED), ("pin_cell_tet_evaporation", {}, PUMP),
    ("pin_cell_tet_evaporation", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # fission neutrons as nuclide files have them (function yields, total nu, Watt / Maxwell / evaporation / tabular spectra)
    ("pin_cell_tet_fission", {}, EVENT), ("pin_cell_tet_fission", {}, FUSED), ("pin_cell_tet_fission", {}, PUMP),
    ("pin_cell_tet_fission", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # several temperatures per nuclide (own grids, reaction lists, probability tables) under a temperature field: every
    # element's temperature picks each nuclide's nearest data set
    ("pin_cell_tet_temperatures", {}, EVENT), ("pin_cell_tet_temperatures", {}, FUSED),
    ("pin_cell_tet_temperatures", {"sort_xs_queue": 1}, SORTED), ("pin_cell_tet_temperatures", {"urr": False}, PUMP),
    ("pin_cell_tet_temperatures", {"urr": False, "sort_xs_queue": 1}, SORTED),
    ("pin_cell_tet_temperatures", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # thermal scattering S(a,b): bound cross sections below energy_max, sab_scatter (continuous / discrete inelastic,
    # coherent / incoherent-discrete elastic, partial binding)
    ("infinite_medium_sab", {}, EVENT), ("pin_cell_tet_sab", {}, EVENT), ("pin_cell_tet_sab", {}, FUSED),
    ("pin_cell_tet_sab", {"sort_xs_queue": 1}, SORTED), ("pin_cell_tet_sab", {}, PUMP),
    ("beavrs_assembly_sab", {"sort_xs_queue": 1}, MIXED), ("pin_cell_tet_sab", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # ... with tables at three temperatures under a temperature field
    ("pin_cell_tet_sab_temperatures", {}, EVENT), ("pin_cell_tet_sab_temperatures", {}, FUSED),
    ("pin_cell_tet_sab_temperatures", {"n_particles": 20000}, {"tail_threshold": 3000}),
    # unresolved-resonance probability tables: cross sections that depend on the history's URR stream, targets at rest
    ("pin_cell_tet_urr", {}, EVENT), ("pin_cell_tet_urr", {}, FUSED), ("pin_cell_tet_urr", {"sort_xs_queue": 1}, SORTED),
    ("pin_cell_tet_urr", {}, PUMP), ("pin_cell_tet_urr", {"n_particles": 20000}, {"tail_threshold": 3000}),
    ("pin_cell_tet_urr", {"seed_through_double": False}, MIXED),
    # all of it at once on the reference's BEAVRS mesh
    ("beavrs_assembly_full", {"sort_xs_queue": 1}, MIXED), ("beavrs_assembly_full", {}, EVENT)])
def test_generations_bit_exact(oracle_bridge, ctx_factory, name, kw, opts):
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=True, track_tally=True, **kw)
    for k, v in opts.items():
        g.set_option(k, v)
    o.sample_initial_source()
    g.sample_initial_source()
    laws = np.zeros(8, np.int64)
    o.L.orc_law_samples(laws.ctypes.data_as(C.POINTER(C.c_longlong)))  # reset
    o.L.orc_reaction_samples(np.zeros(8, np.int64).ctypes.data_as(C.POINTER(C.c_longlong)))
    for batch in range(1, 4):
        ro = o.run_generation(batch)
        rg = g.run_generation(batch)
        if name == "pin_cell_tet_reactions" and batch == 3:
            # every position of the reaction lists was sampled (MT 51, 52, 53, 91, 875)
            rxs = np.zeros(8, np.int64)
            o.L.orc_reaction_samples(rxs.ctypes.data_as(C.POINTER(C.c_longlong)))
            assert (rxs[:5] >= 30).all(), rxs
        if name == "pin_cell_tet_tabular" and batch == 3:
            # the problem reaches every inelastic law: continuous tabular, correlated, Kalbach-Mann, N-body
            o.L.orc_law_samples(laws.ctypes.data_as(C.POINTER(C.c_longlong)))
            assert laws[0] == 0 and (laws[1:5] >= 50).all(), laws
        if name == "pin_cell_tet_evaporation" and batch == 3:
            # ... and the Maxwell, evaporation and Watt spectra
            o.L.orc_law_samples(laws.ctypes.data_as(C.POINTER(C.c_longlong)))
            assert (laws[5:] >= 30).all(), laws
        a, b = o.history_records(), g.history_records()
        for f in ("seed_tracking", "seed_source", "seed_urr", "n_segments", "n_collisions", "n_progeny",
                  "last_elem", "fate", "trace_hash"):
            bad = np.flatnonzero(a[f] != b[f])
            assert bad.size == 0, (batch, f, bad.size, bad[:4].tolist(), [str(a[k]) for k in bad[:2]], [str(b[k]) for k in bad[:2]])
        for f in ("E", "k_collision", "k_absorption", "k_tracklength"):
            assert np.array_equal(a[f].view(np.uint64), b[f].view(np.uint64)), (batch, f)
        assert o.fission_bank().tobytes() == g.fission_bank().tobytes()
        assert o.source_bank().tobytes() == g.source_bank().tobytes()  # sampled next-generation source
        for f in ("n_segments", "n_collisions", "n_bank_sites", "n_failed_traces", "n_vertex_edge_hits"):
            assert getattr(ro, f) == getattr(rg, f), f
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength", "total_weight"):
            assert getattr(rg, f) == pytest.approx(getattr(ro, f), rel=1e-12), f
        assert rg.n_failed_traces == 0
    for tid in (1, 2, 3):
        _, so, qo, nro = o.get_tally(tid)
        sg, qg, nrg = g.get_tally(tid)
        assert nro == nrg == 2
        assert np.allclose(sg.ravel(), so, rtol=1e-11, atol=1e-300)
        assert np.allclose(qg.ravel(), qo, rtol=1e-11, atol=1e-300)
        assert so.sum() > 0


@pytest.mark.parametrize("name,opts", [("pin_cell_tet", MIXED), ("beavrs_assembly", EVENT), ("infinite_medium", FUSED),
                                       ("assembly_quad", PUMP), ("pin_cell_tet_temperatures", MIXED)])
def test_nuclide_bin_tallies(oracle_bridge, ctx_factory, name, opts):
    """Tallies with nuclide bins (mcw_set_tally_nuclides; MC OpenMCTally.C:188-204 `nuclides = all | list`) on both
    estimators against the oracle, whose per-bin scoring equals the reference's own score_all_nuclides / nuclide-bin loop
    bit for bit (tests/test_oracle_pinning.py::test_nuclide_bins_against_reference_tally_objects).  Sums over histories are
    atomic-order dependent: 1e-11 like the other tallies.  Bins of nuclides a material does not hold and the never-scored
    bin of `all` stay exactly zero; the per-nuclide totals add up to the total bin."""
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=True, track_tally=True, nuclide_tallies=True)
    for k, v in opts.items():
        g.set_option(k, v)
    o.sample_initial_source()
    g.sample_initial_source()
    for batch in range(1, 4):
        o.run_generation(batch)
        g.run_generation(batch)
    assert o.source_bank().tobytes() == g.source_bank().tobytes()
    n = len(lib.nuclides)
    for tid, n_bins, n_sc in ((4, 4, 5), (5, n + 2, 4), (1, 1, 5)):
        assert g.tally_nuclide_bins(tid) == n_bins
        _, so, qo, nro = o.get_tally(tid)
        sg, qg, nrg = g.get_tally(tid)
        assert nro == nrg == 2 and sg.shape[2] == n_bins * n_sc and sg.size == so.size
        assert np.allclose(sg.ravel(), so, rtol=1e-11, atol=1e-300)
        assert np.allclose(qg.ravel(), qo, rtol=1e-11, atol=1e-300)
        assert np.array_equal(sg.ravel() == 0, so == 0)
        assert so.sum() > 0
    s5 = g.get_tally(5)[0].reshape(n + 2, 4)
    assert not s5[n + 1].any()                                     # allocated, never scored
    assert np.allclose(s5[:n, 0].sum(), s5[n, 0], rtol=1e-10)      # total: sum over nuclides == material total



@pytest.mark.parametrize("name,opts", [("infinite_medium_iso", EVENT), ("pin_cell_tet", MIXED), ("pin_cell_hex", FUSED),
                                       ("assembly_quad", PUMP), ("beavrs_assembly", EVENT)])
def test_cached_trace_points_bit_exact(oracle_bridge, ctx_factory, tmp_path, name, opts):
    """mcw_get_traces (RT TraceData.h:15-29: the point and the ray's aux data cached per completed segment, what
    RayTracingExodus writes) against the oracle: every record of every ray, byte for byte -- positions on element
    sides, energies, the three stream seeds as doubles, progeny counts -- in event rounds, the fused tail and the
    persistent kernel; then through the Exodus EDGE2 writer and back."""
    from macaw_b200 import exodus
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, name, tallies=True, n_particles=300, record_traces=2)
    for k, v in opts.items():
        g.set_option(k, v)
    o.sample_initial_source()
    g.sample_initial_source()
    for batch in (1, 2):
        o.run_generation(batch)
        g.run_generation(batch)
        a, b = o.traces(), g.traces()
        assert a.size == b.size and a.size > 300
        assert a.tobytes() == b.tobytes()
    path = str(tmp_path / "rays.e")
    exodus.write_ray_exodus(path, b)
    coord, conn, d = exodus.read_ray_exodus(path)
    assert coord.shape[0] == b.size and conn.shape[0] == b.size - 300
    assert np.array_equal(d["aux_energy"], b["aux"][:, 0][b["kind"] != 0])


def test_e2e_host_buffers_round_trip(oracle_bridge, ctx_factory):
    """The reference-facing call sequence with HOST buffers: set_source_bank -> run -> get_source_bank."""
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, "pin_cell_tet")
    o.sample_initial_source()
    src = o.source_bank()
    g.set_source_bank(src)
    ro, rg = o.run_generation(1), g.run_generation(1)
    assert o.source_bank().tobytes() == g.source_bank().tobytes()
    assert rg.k_generation == pytest.approx(ro.k_generation, rel=1e-12)


def test_error_paths(ctx_factory):
    from macaw_b200.lib import McwError
    c = ctx_factory()
    with pytest.raises(McwError):
        c.transport_generation(1)  # nothing configured: MCW_E_NOT_READY
    mesh, lib, bm, s = problems.infinite_medium()
    c.set_mesh(mesh)
    with pytest.raises(McwError):
        c.set_library(lib, [0, 7])  # block mapped to an unknown material (MC CollisionKernel.C:84-90)
    # data the library does not carry is refused, never dropped: Debye-Waller incoherent elastic, log-log probability
    # tables, thermal grids that end below energy_max (the samplers would read past them), unknown tally nuclides
    from macaw_b200 import nucdata
    c = ctx_factory()
    mesh, lib, bm, s = problems.pin_cell_tet()
    nucdata.add_thermal_scattering(lib)
    lib.thermal_tables[1].elastic_type = 3
    c.set_mesh(mesh)
    with pytest.raises(McwError, match="unsupported elastic type"):
        c.set_library(lib, bm, np.arange(4, dtype=np.int32))
    lib.thermal_tables[1].elastic_type = 1
    lib.thermal_tables[1].elastic_x = lib.thermal_tables[1].elastic_x * 0.1
    with pytest.raises(McwError, match="ends below energy_max"):
        c.set_library(lib, bm, np.arange(4, dtype=np.int32))
    mesh, lib, bm, s = problems.pin_cell_tet()
    nucdata.add_urr_tables(lib)
    lib.urr_tables[0].interpolation = 5
    with pytest.raises(McwError, match="linear-linear"):
        c.set_library(lib, bm, np.arange(4, dtype=np.int32))
    c = problems.configure(ctx_factory(), *problems.pin_cell_tet()[:3], problems.pin_cell_tet()[3], tallies=True)
    with pytest.raises(McwError, match="Could not find the nuclide"):
        c.set_tally_nuclides(1, [0, 99, -1])
    # a source box that misses the mesh is the reference's fatal error, not an empty first generation
    mesh, lib, bm, s = problems.pin_cell_tet(n_particles=64)
    s.source_box[0], s.source_box[3] = 50.0, 51.0
    c = problems.configure(ctx_factory(), mesh, lib, bm, s, tallies=False)
    with pytest.raises(McwError, match="external source sites sampled were rejected"):
        c.sample_initial_source()


@pytest.mark.parametrize("n", [1, 7, 33])
def test_ragged_banks_bit_exact(oracle_bridge, ctx_factory, n):
    """Banks smaller than a warp / not a multiple of 32: partial warps in every kernel, a fission bank
    that can be empty (MCW_E_PHYSICS like OpenMC's "No fission sites banked", OMC eigenvalue.cpp:107-109)."""
    from macaw_b200.lib import McwError
    o, g, mesh, lib, s = both(oracle_bridge, ctx_factory, "pin_cell_tet", tallies=True, n_particles=n)
    o.sample_initial_source()
    g.sample_initial_source()
    assert o.source_bank().tobytes() == g.source_bank().tobytes()
    for batch in (1, 2):
        try:
            ro = o.run_generation(batch)
        except RuntimeError:  # the oracle banked no fission site
            with pytest.raises(McwError):
                g.run_generation(batch)
            return
        rg = g.run_generation(batch)
        a, b = o.history_records(), g.history_records()
        assert a.size == b.size == n
        for f in ("seed_tracking", "seed_urr", "n_segments", "n_collisions", "n_progeny", "last_elem", "fate", "trace_hash"):
            assert np.array_equal(a[f], b[f]), (batch, f)
        assert o.fission_bank().tobytes() == g.fission_bank().tobytes()
        assert o.source_bank().tobytes() == g.source_bank().tobytes() and g.source_bank().size == n
        assert rg.k_generation == pytest.approx(ro.k_generation, rel=1e-12)


def test_full_fission_bank(oracle_bridge, ctx_factory):
    """k about 4 with the keff guess 1 of the first generation: more sites than the bank's 3 n.  Which sites are dropped
    depends on the order histories reach the bank (thread / atomic order, in the reference too), so this is held by
    properties, on the GPU and on the oracle alike: tests/problems.py check_full_bank.  The next generation (keff now
    about 4) fits again and is an ordinary one."""
    n = 3000
    mesh, lib, bm, s = problems.infinite_medium_supercritical(n_particles=n, n_inactive=1)
    for obj in (problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s), problems.configure(ctx_factory(), mesh, lib, bm, s)):
        obj.sample_initial_source()
        r = obj.run_generation(1)
        problems.check_full_bank(r, obj.history_records(), obj.fission_bank(), n)
        assert 3.0 < r.k_generation < 6.0
        assert len(obj.source_bank()) == n
        r2 = obj.run_generation(2)
        assert r2.n_bank_dropped == 0 and 0.5 * n < r2.n_bank_sites < 2 * n and r2.n_failed_traces == 0
    assert "fission bank is full" in obj.L.mcw_last_error(obj.h).decode()


def test_empty_inputs(ctx_factory):
    """Zero-length calls are no-ops, not errors or crashes (the reference's containers are simply empty)."""
    import torch
    c = ctx_factory()
    mesh, lib, bm, s = problems.infinite_medium()
    problems.configure(c, mesh, lib, bm, s, tallies=False)
    assert c.xs_lookup(np.zeros(0), np.zeros(0, np.int32)).shape == (0, 4)
    d, seg, hsh = c.trace_rays(np.zeros((0, 3)), np.zeros((0, 3)))
    assert d.size == seg.size == hsh.size == 0
    assert c.locate_points(np.zeros((0, 3))).size == 0
    buf = torch.zeros(8, dtype=torch.float64, device="cuda")
    for mode in (0, 1, 2):
        c.xs_lookup_device(0, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 1, mode)
    one = c.xs_lookup(np.array([1.0e3]), np.array([0], np.int32))
    assert one.shape == (1, 4) and one[0, 0] > 0


def test_full_size_properties_pin_cell_1m():
    """BASELINE.json configs[1] at bench size (101 400 TET4, 10^6 histories, 9 nuclides x 10 000
    points): too big for the scalar oracle in a test, so size-independent properties of the path:
    counter checksums, a checksum of checksums for the collision tally (sum of the `total` score =
    number of collisions, because every history has weight 1), bank order, progeny bookkeeping,
    exact bank resampling, agreement of the three k estimators, and run-to-run determinism of every
    per-history quantity across event-loop regimes."""
    from macaw_b200 import meshgen, nucdata
    from macaw_b200.context import Context
    mesh = meshgen.pin_cell_tet(26, 25)
    lib = nucdata.make_library(9, G=10000, seed=1)
    h = meshgen.PIN_PITCH / 2
    n = 1_000_000

    def run(opts, sort_xs=0):
        s = cstructs.make_settings(n, n_inactive=1, source_box=(-h, -h, -h, h, h, h), record_traces=1, sort_xs_queue=sort_xs)
        c = Context(0)
        c.set_mesh(mesh)
        c.set_library(lib, [0, 1, 2, 3], [0, 1, 2, 3])
        c.set_settings(s)
        c.add_tally(1, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["cell"], [0, 1], None)
        for k, v in opts.items():
            c.set_option(k, v)
        c.sample_initial_source()
        r1 = c.run_generation(1)
        r2 = c.run_generation(2)
        rec, bank, src = c.history_records(), c.fission_bank(), c.source_bank()
        t, _, nr = c.get_tally(1)
        c.close()
        return r1, r2, rec, bank, src, t, nr

    r1, r2, rec, bank, src, t, nr = run({})
    assert r2.n_histories == n and rec.size == n and src.size == n
    assert int(rec["n_segments"].sum()) == r2.n_segments and int(rec["n_collisions"].sum()) == r2.n_collisions
    assert r2.n_segments > 1000 * n and r2.n_collisions > 10 * n  # fine mesh: segment dominated
    assert set(np.unique(rec["fate"]).tolist()) <= {0, 2}  # reflective box: absorbed, or a counted failed trace
    assert int((rec["fate"] == 2).sum()) == r2.n_failed_traces <= 2
    # fission bank: sorted by (parent, progeny), progeny ids 0..n_progeny-1 of each parent (OMC bank.cpp:60-104)
    key = bank["parent_id"].astype(np.int64) * 64 + bank["progeny_id"]
    assert np.all(np.diff(key) > 0) and bank["progeny_id"].max() < 64
    per_parent = np.bincount(bank["parent_id"] - 1, minlength=n)
    assert np.array_equal(per_parent, rec["n_progeny"]) and bank.size == r2.n_bank_sites
    # sites are born where fission happened: fuel elements only, inside the pin
    assert np.all(mesh.block_id[src["elem"]] == 0) and np.all(np.hypot(src["r"][:, 0], src["r"][:, 1]) < meshgen.PIN_RADII[0] + 0.05)
    # the three estimators of one generation agree statistically (sigma ~ 1e-3 at 1e6 histories)
    assert abs(r2.k_collision - r2.k_tracklength) < 5e-3 and abs(r2.k_absorption - r2.k_tracklength) < 5e-3
    assert abs(r2.k_generation - bank.size / n * r1.k_generation) < 5e-3 * r2.k_generation  # sites banked with 1/k weighting
    # checksum of checksums: batch 2 is the only active batch; total score = sig_t * wgt / sig_t per collision
    assert nr == 1 and t.shape == (mesh.n_elem, 1, 2)
    # (Tally::accumulate normalises by the number of source particles, OMC tally.cpp:650-660)
    assert abs(t[:, 0, 1].sum() * n - r2.n_collisions) <= 1e-9 * r2.n_collisions
    assert np.all(t >= 0) and t[:, 0, 0].sum() > 0
    # the same run through event rounds only, unordered queues, sorted refresh: every per-history result identical
    r1b, r2b, recb, bankb, srcb, tb, _ = run({"tail_threshold": 0, "ordered_queues": 0, "xs_sort_min": 1}, sort_xs=1)
    assert recb.tobytes() == rec.tobytes() and bankb.tobytes() == bank.tobytes() and srcb.tobytes() == src.tobytes()
    assert np.allclose(tb, t, rtol=1e-11, atol=1e-300) and r2b.k_generation == pytest.approx(r2.k_generation, rel=1e-12)


def test_full_size_oracle_parity_pin_cell_1m(oracle_bridge):
    """The bench workload at bench size against the oracle on all host threads (~20 s of CPU): one
    generation of 10^6 histories, ~1.5e9 element crossings.  Every history's crossing hash, seeds,
    counters, fate and final energy, and the whole fission bank, must be bit-identical."""
    import os
    from macaw_b200 import meshgen, nucdata
    from macaw_b200.context import Context
    mesh = meshgen.pin_cell_tet(26, 25)
    lib = nucdata.make_library(9, G=10000, seed=1)
    h = meshgen.PIN_PITCH / 2
    n = 1_000_000
    s = cstructs.make_settings(n, n_inactive=1, source_box=(-h, -h, -h, h, h, h), record_traces=1)
    g = Context(0)
    g.set_mesh(mesh)
    g.set_library(lib, [0, 1, 2, 3], [0, 1, 2, 3])
    g.set_settings(s)
    g.sample_initial_source()
    src = g.source_bank().copy()
    rg = g.run_generation(1)
    rec, bank = g.history_records(), g.fission_bank()
    g.close()
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh, np.arange(4, dtype=np.int32))
    o.set_library(lib, [0, 1, 2, 3])
    o.set_settings(s)
    o.set_source_bank(src)
    o.transport_generation(1, os.cpu_count() or 1)
    ro = o.history_records()
    for f in ("trace_hash", "seed_tracking", "seed_source", "seed_urr", "n_segments", "n_collisions", "n_progeny",
              "last_elem", "fate"):
        assert np.array_equal(rec[f], ro[f]), (f, int((rec[f] != ro[f]).sum()))
    for f in ("E", "k_collision", "k_absorption", "k_tracklength"):
        assert np.array_equal(rec[f].view(np.uint64), ro[f].view(np.uint64)), f
    assert int(ro["n_segments"].sum()) == rg.n_segments > 1_000_000_000
    ob = o.fission_bank()
    assert ob.tobytes() == bank.tobytes()


def test_xs_lookup_against_reference_objects_fixture(ctx_factory, golden_dir):
    """GPU against the REFERENCE's own outputs, no oracle in between: tests/golden/ref_collisions.npz holds what the
    reference's compiled Material::calculate_xs (OMC material.cpp:773-829 -> nuclide.cpp:550-750) and Nuclide::init_grid
    (:446-476) gave on Nuclide / Material objects built by its own constructors (tests/golden/make_ref_collisions.py).
    mcw_xs_lookup returns the same macroscopic total / absorption / fission / nu-fission bit for bit for all 4 x 1000
    energies (grid points included), and the library's logarithmic grid index equals the reference's for every nuclide."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_ref_collisions", os.path.join(golden_dir, "make_ref_collisions.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ref = np.load(os.path.join(golden_dir, "ref_collisions.npz"))
    lib = gen.library()
    for i, n in enumerate(lib.nuclides):
        if n.fissionable:
            n.xs = ref[f"xs_{i}"].copy()
    mesh, _, bm, s = problems.infinite_medium(n_particles=100, n_log_bins=lib.n_log_bins)
    g = problems.configure(ctx_factory(), mesh, lib, bm, s, False)
    for i_mat in range(len(lib.materials)):
        E, _, _, _ = gen.events(i_mat, lib)
        got = g.xs_lookup(E, np.full(E.size, i_mat, np.int32))
        assert np.array_equal(got.view(np.uint64), ref[f"m{i_mat}_macro"].view(np.uint64)), i_mat
    for i in range(len(lib.nuclides)):
        assert np.array_equal(g.grid_index(i, lib.n_log_bins), ref[f"grid_index_{i}"]), i


def test_xs_lookup_at_temperatures_against_reference_fixture(ctx_factory, golden_dir):
    """GPU against the REFERENCE's own outputs on a library with data sets at 293.6 / 600 / 900 K (own grids) and reaction
    lists: tests/golden/ref_collisions_full.npz holds what the reference's Material::calculate_xs gave at particle
    temperatures of 0, 446, 700, 1200 and 447.5 K, i.e. after Nuclide::calculate_xs picked each nuclide's nearest data set
    (OMC nuclide.cpp:614-626).  mcw_xs_lookup under mcw_set_temperature(T) returns the same macroscopic cross sections bit
    for bit, and every data set's logarithmic grid index equals Nuclide::init_grid's on its own grid."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_ref_collisions_full", os.path.join(golden_dir, "make_ref_collisions_full.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ref = np.load(os.path.join(golden_dir, "ref_collisions_full.npz"))
    lib = gen.library()
    for key, x in gen.data_sets(lib):
        if f"xs_{key}" in ref:
            x.xs = ref[f"xs_{key}"].copy()
    mesh, _, bm, s = problems.infinite_medium(n_particles=100, n_log_bins=lib.n_log_bins)
    g = problems.configure(ctx_factory(), mesh, lib, bm, s, False)
    differ = 0
    for i_mat in range(len(lib.materials)):
        E, _, _, _ = gen.events(i_mat, lib)
        g.set_temperature(gen.TEMPERATURES[i_mat])
        got = g.xs_lookup(E, np.full(E.size, i_mat, np.int32))
        assert np.array_equal(got.view(np.uint64), ref[f"m{i_mat}_macro"].view(np.uint64)), i_mat
        g.set_temperature(293.6)
        differ += int((g.xs_lookup(E, np.full(E.size, i_mat, np.int32))[:, 0] != got[:, 0]).sum())
    assert differ > 2000  # the 600 K and 900 K sets really are other numbers
    entries = [str(i) for i in range(len(lib.nuclides))] + [f"{i}_{k}" for i, n in enumerate(lib.nuclides)
                                                           for k in range(len(n.more_temperatures or []))]
    for e, key in enumerate(entries):
        assert np.array_equal(g.grid_index(e, lib.n_log_bins), ref[f"grid_index_{key}"]), key


@pytest.mark.parametrize("name,world", [("pin_cell_tet", 2), ("beavrs_assembly", 2)])
def test_two_gpus_equal_one_gpu(ctx_factory, tmp_path, name, world):
    """Rank-count invariance ON THE GPUS over NCCL (OMC eigenvalue.cpp:71-311 synchronize_bank): the next source
    bank of an N-GPU run, concatenated in rank order, equals the 1-GPU run's byte for byte; k per generation to
    1e-12 (summation order across ranks), tallies to 1e-11.  Skipped on a box with fewer GPUs."""
    import os, socket, subprocess, sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n_particles, n_batches = 40000, 3
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "g")
    procs = [subprocess.Popen([sys.executable, os.path.join(root, "tests", "dist_gpu_worker.py"), str(r), str(world),
                               str(port), out, name, str(n_particles), str(n_batches)], cwd=root,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    from macaw_b200 import study as st
    mesh, lib, bm, s = problems.PROBLEMS[name](n_particles=n_particles)
    blocks = list(range(len(bm))) if int(mesh.block_id.max()) < len(bm) else [int(b) for b in mesh.block_ids_sorted()]
    kernel = st.CollisionKernel(blocks=blocks, materials=[m + 1 for m in bm], z_coord=100.0 if mesh.dim == 2 else None)
    tally = st.OpenMCTally(id=1, scores=["flux", "nu-fission"], filters=["cell"], estimator="COLLISION")
    one = st.OpenMCStudy(mesh, lib, kernel, s, [tally], device=0)
    ks = []
    for _ in range(n_batches):
        r = one.execute()
        ks.append([r.k_generation, r.keff, r.k_collision, r.k_absorption, r.k_tracklength])
    parts = [np.load(f"{out}.{r}.npz") for r in range(world)]
    src = np.concatenate([p["source"] for p in parts])
    assert src.tobytes() == one.ctx.source_bank().tobytes()
    rec = np.concatenate([p["records"] for p in parts])
    assert rec.tobytes() == one.ctx.history_records().tobytes()  # every history of the last generation, bit for bit
    for p in parts:
        assert np.allclose(p["ks"], np.asarray(ks), rtol=1e-12, atol=0)
    t_sum, _, _ = one.ctx.get_tally(1)
    for p in parts:  # the VALUE arrays are all-reduced in place: every rank accumulates the global tally
        assert np.allclose(p["tally_sum"], t_sum, rtol=1e-11, atol=1e-300)
    one.ctx.close()
