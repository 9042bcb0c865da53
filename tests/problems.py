"""Identical problem set-ups for the CPU oracle and the GPU library (same numpy arrays on both sides)."""
import numpy as np

from macaw_b200 import cstructs, meshgen, nucdata

ALL_SCORES = [0, 1, 2, 3, 4]
EBINS = [1e-5, 0.625, 1.0e5, 2.0e7]


def configure(obj, mesh, lib, block_material, settings, tallies=True, track_tally=False, nuclide_tallies=False):
    # block ids are 0..n_materials-1 in every synthetic mesh; pass them explicitly because a coarse mesh
    # may not contain every block (e.g. no element centroid falls inside the thin gap)
    block_ids = np.arange(len(block_material), dtype=np.int32)
    if int(mesh.block_id.max()) >= len(block_material):  # raw Exodus subdomain ids (BEAVRS: 10004 ... 11006)
        block_ids = mesh.block_ids_sorted()
        assert block_ids.size == len(block_material)
    if hasattr(obj, "L") and hasattr(obj.L, "orc_create"):
        obj.set_mesh(mesh, block_ids)
        obj.set_library(lib, block_material)
    else:
        obj.set_mesh(mesh)
        obj.set_library(lib, block_material, block_ids)
    if getattr(lib, "element_temperature", None) is not None:  # the coupled temperature of CollisionKernel, per element
        obj.set_temperature(lib.element_temperature)
    obj.set_settings(settings)
    if tallies:
        obj.add_tally(1, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["cell"], ALL_SCORES, EBINS)
        obj.add_tally(2, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["universe"], [0, 3])
    if track_tally:
        obj.add_tally(3, cstructs.ESTIMATOR["tracklength"], cstructs.GRANULARITY["cell"], [0, 3, 4], EBINS)
    if nuclide_tallies:
        # OpenMCTally `nuclides` (OpenMCTally.C:188-204): a list (+ total) on a collision tally, `all` on a track-length one
        n = len(lib.nuclides)
        obj.add_tally(4, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["universe"], ALL_SCORES, EBINS)
        obj.set_tally_nuclides(4, [min(2, n - 1), 0, n - 1, -1])
        obj.add_tally(5, cstructs.ESTIMATOR["tracklength"], cstructs.GRANULARITY["global"], [1, 2, 4, 0])
        obj.set_tally_nuclides(5, list(range(n)) + [-1, -2])
    return obj


def infinite_medium(n_particles=2000, G=400, n_inactive=1, **kw):
    mesh = meshgen.infinite_medium(5)
    lib = nucdata.infinite_medium_library(G=G)
    s = cstructs.make_settings(n_particles, n_inactive=n_inactive, source_box=(-5, -5, -5, 5, 5, 5), **kw)
    return mesh, lib, [0, 1], s


def infinite_medium_supercritical(**kw):
    """nu ten times as large: k is about 4, so the first generation (keff guess 1) makes more than the 3 n sites the
    fission bank holds (OMC physics.cpp:203-209 "The shared fission bank is full")."""
    mesh, lib, bm, s = infinite_medium(**kw)
    for n in lib.nuclides:
        if n.fissionable:
            n.nu_poly = (10.0 * n.nu_poly[0], 10.0 * n.nu_poly[1])
            n.xs = n.xs.copy()
            n.xs[:, 3] = (n.nu_poly[1] * n.energy + n.nu_poly[0]) * n.xs[:, 2]
    return mesh, lib, bm, s


def pin_cell_tet(n_particles=2000, G=400, n_xy=8, n_z=3, n_inactive=1, **kw):
    mesh = meshgen.pin_cell_tet(n_xy, n_z)
    lib = nucdata.make_library(9, G=G)
    h = meshgen.PIN_PITCH / 2
    s = cstructs.make_settings(n_particles, n_inactive=n_inactive, source_box=(-h, -h, -h, h, h, h), **kw)
    return mesh, lib, [0, 1, 2, 3], s


def pin_cell_hex(n_particles=2000, G=400, n_xy=10, n_z=2, n_inactive=1, **kw):
    mesh = meshgen.pin_cell_hex(n_xy, n_z)
    lib = nucdata.make_library(9, G=G)
    h = meshgen.PIN_PITCH / 2
    s = cstructs.make_settings(n_particles, n_inactive=n_inactive, source_box=(-h, -h, -h, h, h, h), **kw)
    return mesh, lib, [0, 1, 2, 3], s


def assembly_quad(n_particles=2000, G=400, n_pins=3, cells_per_pin=7, n_inactive=1, **kw):
    mesh = meshgen.assembly_quad(n_pins, cells_per_pin)
    lib = nucdata.make_library(9, G=G)
    L = n_pins * meshgen.PIN_PITCH
    s = cstructs.make_settings(n_particles, n_inactive=n_inactive, is_2d=True, z_coord=100.0,
                               source_box=(0, 0, -1, L, L, 1), **kw)
    return mesh, lib, [0, 1, 2, 3], s


def infinite_medium_iso(**kw):
    """Every nuclide isotropic-in-lab, as in the reference's own study test (test/tests/study/materials.xml)."""
    mesh, lib, bm, s = infinite_medium(**kw)
    for m in lib.materials:
        m.isotropic = [1] * len(m.nuclides)
    return mesh, lib, bm, s


def pin_cell_tet_iso(**kw):
    """Fuel and water nuclides isotropic-in-lab, gap and clad not (mixed flags)."""
    mesh, lib, bm, s = pin_cell_tet(**kw)
    for k, m in enumerate(lib.materials):
        m.isotropic = [1 if (k in (0, 3) and j % 2 == 0) else 0 for j in range(len(m.nuclides))]
    return mesh, lib, bm, s


def beavrs_assembly(n_particles=3000, G=400, n_inactive=1, **kw):
    """BASELINE.json configs[2] on the reference's own mesh (tests/golden/beavrs_assembly_mesh.npz, 23 876 QUAD4
    from examples/beavrs_assembly/31enr_16_in.e) with the settings.xml source box; 35 synthetic nuclides."""
    import os
    mesh = meshgen.beavrs_assembly_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                                     "beavrs_assembly_mesh.npz"))
    lib = nucdata.make_library(35, G=G)
    s = cstructs.make_settings(n_particles, n_inactive=n_inactive, is_2d=True, z_coord=100.0,
                               source_box=(-0.6, -20.7, -1, 20.7, 0.6, 1), **kw)
    role = {"fuel": 0, "gap": 1, "clad": 2, "water": 3}
    return mesh, lib, [role[meshgen.BEAVRS_BLOCK_ROLE[b]] for b in sorted(meshgen.BEAVRS_BLOCK_ROLE)], s


def pin_cell_tet_tabular(**kw):
    """Tabulated secondary distributions in the data: angular distributions of elastic / inelastic reactions,
    ContinuousTabular laws in the CM and the LAB system (OMC distribution_angle.cpp, distribution_energy.cpp)."""
    mesh, lib, bm, s = pin_cell_tet(**kw)
    # 15 nuclides: enough inelastic reactions for every law (level-free here: continuous tabular in CM / LAB, correlated,
    # Kalbach-Mann, N-body)
    lib2 = nucdata.make_library(15, G=lib.nuclides[0].energy.size, tabular=True)
    # the nuclides whose inelastic reaction is a correlated / Kalbach-Mann / N-body product get a density at which
    # a few thousand histories sample each law; Pu239 (Kalbach-Mann in CM) joins the fuel
    fuel = lib2.materials[0]
    if 4 not in fuel.nuclides:
        fuel.nuclides.append(4)
        fuel.densities.append(1e-5)
    for m in lib2.materials:
        for j, i in enumerate(m.nuclides):
            if lib2.nuclides[i].inelastic.size and lib2.nuclides[i].inelastic_law >= 2:
                m.densities[j] = max(m.densities[j], 1e-2)
    return mesh, lib2, bm, s


def pin_cell_tet_reactions(**kw):
    """Every nuclide with an inelastic reaction carries a LIST of scattering reactions (OMC index_inelastic_scatter_):
    levels MT 51 / 52 / 53, the continuum MT 91 (for every third nuclide with multiplicity 1.25: the weight changes) and
    (n,2n) MT 875 with multiplicity 2 -- on top of the tabulated distributions of pin_cell_tet_tabular."""
    mesh, lib, bm, s = pin_cell_tet_tabular(**kw)
    return mesh, nucdata.add_reactions(lib), bm, s


def pin_cell_tet_evaporation(**kw):
    """Reaction lists with Maxwell / evaporation / Watt spectra and energy-dependent multiplicities on scattering reactions
    (nucdata.add_evaporation_laws), incl. the first (inline) reaction of some nuclides."""
    mesh, lib, bm, s = pin_cell_tet_reactions(**kw)
    return mesh, nucdata.add_evaporation_laws(lib), bm, s


def pin_cell_tet_fission(**kw):
    """Fission neutrons as evaluated files have them (tabulated total nu and yields, six delayed groups with Maxwell /
    evaporation / continuous tabular spectra, prompt Watt with tabulated parameters or continuous tabular), on top of the
    reaction lists and tabulated distributions."""
    mesh, lib, bm, s = pin_cell_tet_reactions(**kw)
    return mesh, nucdata.add_fission_data(lib), bm, s


def pin_cell_tet_temperatures(**kw):
    """Several temperatures per nuclide (data sets at 293.6 / 600 / 900 K with their own grids, reaction lists and
    probability tables; every fourth nuclide keeps one set) and a temperature FIELD on the mesh, hot in the middle of the
    fuel (1100 K) and cold at the cell corner (100 K): fuel and water each span two temperature classes, each nuclide's
    data set is the one nearest to the element's temperature (OMC nuclide.cpp:614-626, MC CollisionKernel.C:109)."""
    urr = kw.pop("urr", True)  # False: no probability tables, so that the staged (tile) refresh kernels run as well
    mesh, lib, bm, s = pin_cell_tet_reactions(**kw)
    nucdata.add_fission_data(lib)  # before the temperature sets, which share the nuclide's fission data
    nucdata.add_temperatures(nucdata.add_urr_tables(lib) if urr else lib)
    c = mesh.xyz[mesh.conn].mean(axis=1)
    r = np.hypot(c[:, 0], c[:, 1])
    lib.element_temperature = 1100.0 - 1000.0 * r / r.max()
    return mesh, lib, bm, s


def pin_cell_tet_sab(**kw):
    """S(alpha, beta) tables on hydrogen (continuous inelastic), on a second nuclide per material (coherent elastic +
    discrete inelastic, fraction 0.75) and a third (incoherent-discrete elastic + skewed discrete inelastic)."""
    mesh, lib, bm, s = pin_cell_tet(**kw)
    return mesh, nucdata.add_thermal_scattering(lib), bm, s


def pin_cell_tet_sab_temperatures(**kw):
    """S(a,b) tables at 293.6 / 400 / 600 K under a temperature field (1100 K in the middle of the fuel, 100 K at the cell
    corner): the water spans all three table temperatures (OMC thermal.cpp:149-171: the closer of the two that bound the
    element's temperature)."""
    mesh, lib, bm, s = pin_cell_tet_sab(**kw)
    nucdata.add_thermal_temperatures(lib)
    c = mesh.xyz[mesh.conn].mean(axis=1)
    r = np.hypot(c[:, 0], c[:, 1])
    lib.element_temperature = 1100.0 - 1000.0 * r / r.max()
    return mesh, lib, bm, s


def beavrs_assembly_sab(**kw):
    """The BEAVRS assembly with bound hydrogen in the water, as its materials.xml has it (<sab name="c_H_in_H2O"/>)."""
    mesh, lib, bm, s = beavrs_assembly(**kw)
    return mesh, nucdata.add_thermal_scattering(lib), bm, s


def pin_cell_tet_urr(**kw):
    """Unresolved-resonance probability tables on every nuclide heavier than A = 80 (absolute and multiply_smooth
    tables, inelastic competition); the band of a history follows its URR stream."""
    mesh, lib, bm, s = pin_cell_tet(**kw)
    return mesh, nucdata.add_urr_tables(lib), bm, s


def beavrs_assembly_full(**kw):
    """BEAVRS assembly with everything a real library brings: S(a,b) on hydrogen, probability tables on the heavy
    nuclides, tabulated secondary distributions."""
    import os
    mesh, lib, bm, s = beavrs_assembly(**kw)
    lib2 = nucdata.make_library(35, G=lib.nuclides[0].energy.size, tabular=True)
    return mesh, nucdata.add_urr_tables(nucdata.add_thermal_scattering(lib2)), bm, s


def infinite_medium_sab(**kw):
    mesh, lib, bm, s = infinite_medium(**kw)
    return mesh, nucdata.add_thermal_scattering(lib), bm, s


def check_full_bank(r, rec, bank, n):
    """What has to hold when the fission bank overflowed (OMC physics.cpp:203-209): the generation goes on with the sites
    that fit, dropped sites are counted, and n_progeny counts BANKED sites only, so that sort_fission_bank's offsets
    (exclusive scan of n_progeny, bank.cpp:60-104) stay inside the bank: the sorted bank has no holes."""
    assert r.n_bank_dropped > 0 and r.n_bank_sites == 3 * n == len(bank)
    assert int(rec["n_progeny"].sum()) == r.n_bank_sites
    assert bank["parent_id"].min() >= 1 and bank["parent_id"].max() <= n
    order = np.lexsort((bank["progeny_id"], bank["parent_id"]))
    assert np.array_equal(order, np.arange(len(bank)))  # sorted by (parent, progeny)
    counts = np.bincount(bank["parent_id"] - 1, minlength=n)
    assert np.array_equal(counts, rec["n_progeny"])      # every parent's sites are there ...
    first = np.concatenate([[0], np.cumsum(counts)[:-1]])
    assert np.array_equal(bank["progeny_id"], np.arange(len(bank)) - np.repeat(first, counts))  # ... numbered 0, 1, 2, ...
    assert np.isfinite(bank["E"]).all() and (bank["E"] > 0).all()


PROBLEMS = {"pin_cell_tet_sab_temperatures": pin_cell_tet_sab_temperatures, "pin_cell_tet_evaporation": pin_cell_tet_evaporation, "pin_cell_tet_fission": pin_cell_tet_fission, "pin_cell_tet_temperatures": pin_cell_tet_temperatures, "pin_cell_tet_reactions": pin_cell_tet_reactions, "pin_cell_tet_urr": pin_cell_tet_urr, "beavrs_assembly_full": beavrs_assembly_full,
            "pin_cell_tet_tabular": pin_cell_tet_tabular, "pin_cell_tet_sab": pin_cell_tet_sab,
            "beavrs_assembly_sab": beavrs_assembly_sab, "infinite_medium_sab": infinite_medium_sab,
            "beavrs_assembly": beavrs_assembly, "infinite_medium_iso": infinite_medium_iso, "pin_cell_tet_iso": pin_cell_tet_iso,
            "infinite_medium": infinite_medium, "pin_cell_tet": pin_cell_tet, "pin_cell_hex": pin_cell_hex,
            "assembly_quad": assembly_quad}


def lookup_points(lib, n, seed=5):
    rng = np.random.default_rng(seed)
    E = np.exp(rng.uniform(np.log(1.1e-5), np.log(1.9e7), n))
    # exact grid points, interval ends and duplicates exercise the search edge cases
    g = lib.nuclides[min(3, len(lib.nuclides) - 1)].energy
    E[: min(64, n)] = g[rng.integers(0, g.size, min(64, n))]
    E[0], E[1] = 1e-5, 2e7 * (1 - 1e-16)
    mat = rng.integers(0, len(lib.materials), n).astype(np.int32)
    return E, mat
