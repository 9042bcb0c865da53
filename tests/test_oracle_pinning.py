"""Pins the CPU oracle to the reference: its own compiled RNG/sampler sources (oracle/_ref), OpenMC's
unit-test known answers, the MOOSE ray-tracing lots.i gold distances, and MaCaw's gold ray file."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from macaw_b200 import cstructs, meshgen, nucdata
from tests import lots_rays

MULT, ADD, STRIDE = 6364136223846793005, 1442695040888963407, 152917
M64 = (1 << 64) - 1


def lcg(s):
    return (MULT * s + ADD) & M64


def test_rng_vs_reference_vectors(oracle_bridge, golden_dir):
    """oracle prn/future_seed/init_particle_seeds == the reference's random_lcg.cpp, bit for bit
    (vectors generated from oracle/_ref/libomc_ref.so by tests/golden/make_fixtures.py)."""
    g = np.load(os.path.join(golden_dir, "ref_rng_vectors.npz"))
    for libm in (False, True):
        L = oracle_bridge.lib(libm)
        s = C.c_uint64(1)
        mine = np.array([L.orc_prn(C.byref(s)) for _ in range(1000)])
        assert np.array_equal(mine, g["prn"])
        fut = np.array([L.orc_future_seed(int(n), 12345) for n in g["future_n"]], np.uint64)
        assert np.array_equal(fut, g["future_seed"])
        for pid, ref in zip(g["ids"], g["particle_seeds"]):
            for stream in range(4):
                assert L.orc_init_seed(int(pid), stream, 1) == int(ref[stream])
    # init_particle_seeds(1) and the first prn
    assert [int(x) for x in g["particle_seeds"][0][:3]] == [3009189780765489792, 5497575982607593661,
                                                             7985962184449697530]


def test_samplers_libm_build_bit_exact_vs_reference(oracle_bridge, golden_dir):
    """With glibc math the oracle's samplers reproduce the reference's compiled random_dist.cpp /
    math_functions.cpp exactly; with the deterministic math header they agree to a few ulp."""
    g = np.load(os.path.join(golden_dir, "ref_rng_vectors.npz"))
    for libm, exact in ((True, True), (False, False)):
        L = oracle_bridge.lib(libm)
        s = C.c_uint64(7)
        maxw = np.array([L.orc_maxwell_spectrum(1.3e6, C.byref(s)) for _ in range(200)])
        s = C.c_uint64(7)
        watt = np.array([L.orc_watt_spectrum(0.988e6, 2.249e-6, C.byref(s)) for _ in range(200)])
        s = C.c_uint64(11)
        rot = np.zeros((200, 3))
        u = (C.c_double * 3)(0.6, 0.0, 0.8)
        for i in range(200):
            s.value = lcg(s.value)  # uniform_distribution(-1,1) consumed one RN in the reference run
            word = (((s.value >> ((s.value >> 59) + 5)) ^ s.value) * 12605985483714917081) & M64
            mu = -1.0 + 2.0 * (((word >> 43) ^ word) * 2.0 ** -64)
            L.orc_rotate_angle(u, mu, None, C.byref(s))
            rot[i] = list(u)
        if exact:
            assert np.array_equal(maxw, g["maxwell"]) and np.array_equal(watt, g["watt"])
            assert np.array_equal(rot, g["rotate"])
        else:
            assert np.allclose(maxw, g["maxwell"], rtol=1e-13, atol=0)
            assert np.allclose(watt, g["watt"], rtol=1e-13, atol=0)
            assert np.allclose(rot, g["rotate"], rtol=0, atol=1e-9)  # 200 chained rotations accumulate ulps


def test_live_reference_library_if_present(oracle_bridge):
    r = oracle_bridge.ref_lib()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    L = oracle_bridge.lib(True)
    rng = np.random.default_rng(3)
    for seed0 in rng.integers(1, 2 ** 63, 50):
        a, b = C.c_uint64(int(seed0)), C.c_uint64(int(seed0))
        assert L.orc_prn(C.byref(a)) == r.ref_prn(C.byref(b)) and a.value == b.value
        n = int(rng.integers(0, 2 ** 62))
        assert L.orc_future_seed(n, int(seed0)) == r.ref_future_seed(n, int(seed0))
        a, b = C.c_uint64(int(seed0)), C.c_uint64(int(seed0))
        assert L.orc_watt_spectrum(0.9e6, 2.2e-6, C.byref(a)) == r.ref_watt_spectrum(0.9e6, 2.2e-6, C.byref(b))
        # isotropic_direction: the reference's own distribution_multi.cpp:73-79 (isotropic-in-lab scattering, source angle)
        if hasattr(r, "ref_isotropic_direction"):
            a, b = C.c_uint64(int(seed0)), C.c_uint64(int(seed0))
            ua, ub = np.zeros(3), np.zeros(3)
            L.orc_isotropic_direction(ua.ctypes.data_as(C.POINTER(C.c_double)), C.byref(a))
            r.ref_isotropic_direction(ub.ctypes.data_as(C.POINTER(C.c_double)), C.byref(b))
            assert a.value == b.value and np.array_equal(ua.view(np.uint64), ub.view(np.uint64))


def test_openmc_math_known_answers(oracle_bridge, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "omc_math_kat.json")))
    L = oracle_bridge.lib(True)  # glibc build: exact equality like the reference's own test
    s = C.c_uint64(kat["maxwell_spectrum"]["seed"])
    assert L.orc_maxwell_spectrum(kat["maxwell_spectrum"]["T"], C.byref(s)) == kat["maxwell_spectrum"]["ref"]
    s = C.c_uint64(kat["watt_spectrum"]["seed"])
    w = kat["watt_spectrum"]
    assert L.orc_watt_spectrum(w["a"], w["b"], C.byref(s)) == w["ref"]
    r = kat["rotate_angle"]
    for libm in (True, False):
        L = oracle_bridge.lib(libm)
        u = (C.c_double * 3)(*r["uvw0"])
        phi = C.c_double(r["phi"])
        L.orc_rotate_angle(u, r["mu"], C.byref(phi), None)
        assert list(u) == r["ref_phi0"]
        u = (C.c_double * 3)(*r["uvw0"])
        L.orc_rotate_angle(u, 1.0, C.byref(phi), None)
        assert list(u) == [1.0, 0.0, 0.0]
        u = (C.c_double * 3)(*r["uvw0_random"])
        s = C.c_uint64(r["seed"])
        L.orc_rotate_angle(u, r["mu_random"], None, C.byref(s))
        assert np.allclose(list(u), r["ref_random"], rtol=1e-12)
    # deterministic-math build within 2 ulp of the known answers
    L = oracle_bridge.lib(False)
    s = C.c_uint64(1)
    assert abs(L.orc_maxwell_spectrum(0.5, C.byref(s)) - 0.27767406743161277) < 1e-15


def test_detmath_within_one_ulp_of_glibc(oracle_bridge):
    L = oracle_bridge.lib(False)
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-40, 40, 20000))
    u = rng.uniform(0, 1, 20000)
    ph = 2 * np.pi * u
    def ulps(a, b):
        return np.abs(np.asarray(a).view(np.int64) - np.asarray(b).view(np.int64))
    assert ulps(np.array([L.orc_log(v) for v in x]), np.log(x)).max() <= 1
    assert ulps(np.array([L.orc_log(v) for v in u]), np.log(u)).max() <= 1
    s, c = np.array([L.orc_sin(v) for v in ph]), np.array([L.orc_cos(v) for v in ph])
    ok = np.abs(np.sin(ph)) > 1e-3
    assert ulps(s[ok], np.sin(ph)[ok]).max() <= 1
    ok = np.abs(np.cos(ph)) > 1e-3
    assert ulps(c[ok], np.cos(ph)[ok]).max() <= 1
    assert np.abs(s - np.sin(ph)).max() < 2e-16 and np.abs(c - np.cos(ph)).max() < 2e-16
    # exp / sinh (Kalbach-Mann slope parameters: O(1) arguments)
    a = rng.uniform(-30, 30, 20000)
    assert ulps(np.array([L.orc_exp(v) for v in a]), np.exp(a)).max() <= 1
    b = np.concatenate([rng.uniform(0.05, 8, 10000), -rng.uniform(0.05, 8, 10000)])
    sh = np.array([L.orc_sinh(v) for v in b])
    assert np.abs(sh / np.sinh(b) - 1).max() < 1e-14


def test_is_within_segment_known_answers(oracle_bridge):
    """RT unit/src/TraceRayToolsTest.C:269-274."""
    L = oracle_bridge.lib(False)
    def f(a, b, p):
        A, B, Pt = [(C.c_double * 3)(*v) for v in (a, b, p)]
        return bool(L.orc_is_within_segment(A, B, Pt, 1e-8))
    assert f((0, 0, 0), (1, 1, 1), (0.5, 0.5, 0.5))
    assert f((0, 0, 0), (1, 1, 1), (0, 0, 0)) and f((0, 0, 0), (1, 1, 1), (1, 1, 1))
    assert not f((0, 0, 0), (1, 1, 1), (1.1, 1.1, 1.1))
    assert not f((0, 0, 0), (1, 1, 1), (0.5, 0.5, 0.6))


LOTS = [("hex8", "vtv", "ntn"), ("hex8", "ctv", "ctn"), ("hex8", "ete", "ete"), ("hex8", "ctc", "ctc"),
        ("quad4", "vtv", "vtv"), ("quad4", "ctv", "ctv"), ("quad4", "ctc", "ctc"),
        # TET4 (the headline bench's element type): libMesh's 24-tets-per-hex GeneratedMesh, 3 x 3 x 3 on [0, 5]^3
        # (moose/modules/ray_tracing/test/tests/traceray/tet4/tests)
        ("tet4", "vtv", "vtv"), ("tet4", "ctv", "ctv"), ("tet4", "ete", "ete"), ("tet4", "ctc", "ctc")]


@pytest.mark.parametrize("et,kind,gold_name", LOTS)
def test_element_walk_reproduces_moose_lots_gold(oracle_bridge, golden_dir, et, kind, gold_name):
    """The oracle's element walk (incl. vertex/edge fallbacks: these rays run through vertices and
    along edges) reproduces the total traced distance of MOOSE's lots.i regression gold CSVs."""
    gold = json.load(open(os.path.join(golden_dir, "rt_lots_gold.json")))[f"{et}/{gold_name}"]
    if et == "hex8":
        mesh = meshgen.hex_grid(5, 5, 5, (0, 0, 0), (5, 5, 5), meshgen.BC_KILL)
    elif et == "tet4":
        mesh = meshgen.tet_grid_24(3, 3, 3, (0, 0, 0), (5, 5, 5), meshgen.BC_KILL)
    else:
        mesh = meshgen.quad_grid(5, 5, (0, 0), (5, 5), meshgen.BC_KILL)
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh)
    start, dirs, elems, sides, maxd = lots_rays.lots(mesh, kind)
    dist, seg, h, st = o.trace_rays(start, dirs, elems, sides, maxd if kind == "ctc" else None)
    assert st["failed"] == 0
    total = dist.sum()
    assert abs(total - gold["total_distance"]) <= 1e-8 * gold["total_distance"], (total, gold)
    if kind != "ctc":
        exp = lots_rays.box_exit_distance(start, dirs, (0, 0, 0), (5, 5, 5), mesh.dim)
        assert np.allclose(dist, exp, rtol=1e-9, atol=1e-9)
        assert st["vertex_edge_hits"] > 0  # the fallbacks really ran


def test_gold_ray_file_rng_facts(oracle_bridge, golden_dir):
    """MaCaw's own gold file (test/tests/study/gold/openmc_study_rays.e): seeds live in doubles
    (SURVEY F3), one tracking RN per segment, URR stream advanced by data::nuclides.size() = 9 per
    energy-changing collision, initial seeds = future_seed(((batch-1)*n + id)*stride, master+stream)."""
    g = np.load(os.path.join(golden_dir, "study_rays_gold.npz"))
    L = oracle_bridge.lib(False)
    E, pid = g["var_aux_energy"], g["var_aux_id"]
    st, ss, su = g["var_aux_seed_tracking"], g["var_aux_seed_source"], g["var_aux_seed_URR"]
    rid, npg = g["var_ray_id"], g["var_aux_n_progeny"]
    assert E.size == 1796 and np.unique(rid).size == 30
    assert np.all(g["var_aux_weight"] == 1.0) and np.all(g["var_aux_particle_type"] == 0)
    # every stored seed is a double that holds an integer < 2^64 exactly representable in 53 bits
    for arr in (st, ss, su):
        assert np.all(arr == np.floor(arr)) and np.all(arr < 2.0 ** 64)
        assert all(L.orc_seed_via_double(int(v)) == int(v) for v in arr[:200])
    n_one_step = n_same = n_urr = n_urr_ok = 0
    for r in np.unique(rid):
        idx = np.nonzero(rid == r)[0]
        # seed_source: constant per ray and equal to the rounded initial seed of stream 1
        p = int(pid[idx[0]])
        seed_id = ((3 - 1) * 30 + p) * STRIDE
        assert np.all(ss[idx] == ss[idx[0]])
        assert float(L.orc_seed_via_double(L.orc_future_seed(seed_id, 1 + 1))) == ss[idx[0]]
        for a, b in zip(idx[:-1], idx[1:]):
            if E[a] == E[b]:
                n_same += 1
                nxt = L.orc_seed_via_double(lcg(int(st[a])))
                n_one_step += float(nxt) == st[b]
                assert su[a] == su[b]  # URR stream untouched when the energy did not change
            else:
                # several collisions can happen between two recorded face crossings (collision points
                # are not recorded): k rounds of (skip 9, round through double)
                n_urr += 1
                adv, ok = int(su[a]), False
                for _ in range(40):
                    adv = L.orc_seed_via_double(L.orc_future_seed(9, adv))
                    if float(adv) == su[b]:
                        ok = True
                        break
                n_urr_ok += ok
    # 919 of the 939 same-energy transitions are exactly one LCG step from the double-rounded seed; the
    # other 20 are final records of absorbed rays (several RNs consumed in the collision)
    assert n_same == 939 and n_one_step == 919
    assert n_urr > 800 and n_urr_ok == n_urr
    # all recorded segment ends lie on faces of the 5x5x1 grid on [-5,5]^3
    ends = g["coord"][g["conn"][:, 1]]
    on_face = (np.isclose(np.abs(ends[:, 0]) % 2, 1, atol=1e-9) | np.isclose(np.abs(ends[:, 1]) % 2, 1, atol=1e-9)
               | np.isclose(np.abs(ends[:, 2]), 5, atol=1e-9))
    assert on_face.all()


def test_ray_trace_output_has_the_structure_of_the_reference_gold(oracle_bridge, golden_dir, tmp_path):
    """The `rays` output (RayTracingExodus, EDGE2 segments + aux data; RT RayTracingMeshOutput.C:252-580) of
    examples/study_5x5 -- the shape of the reference's own study test -- written by macaw_b200/exodus.py from the
    oracle's cached trace points, read back, and held to the same nuclear-data-free facts as the reference's gold
    file: one polyline per particle with the reference's ray ids, every cached point on an element side, sources
    interior, per-ray constant rounded source seed, exactly one tracking-stream step per same-energy record pair
    except a ray's last (absorbed) record."""
    from macaw_b200 import exodus, inputs, nucdata
    from tests import ray_facts
    L = oracle_bridge.lib(False)
    g = np.load(os.path.join(golden_dir, "study_rays_gold.npz"))
    gold = ray_facts.check(g["coord"], g["conn"], {k[4:]: g[k] for k in g.files if k.startswith("var_")}, 30, 3, L)
    assert gold == {"segments": 1796, "rays": 30, "same_energy": 939, "one_step": 919, "final_other": 20}

    deck = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "study_5x5")
    lib = nucdata.library_for_names(inputs.nuclide_names(os.path.join(deck, "materials.xml")), G=400)
    mesh, library, blocks, mat_index, s, T0, n_batches, tallies, aux = inputs.read_input(os.path.join(deck, "study.i"), deck, lib)
    s.n_particles, s.record_traces = 30, 2
    o = oracle_bridge.Oracle()
    o.set_mesh(mesh, np.asarray(blocks, np.int32))
    o.set_library(library, mat_index)
    o.set_settings(s)
    o.sample_initial_source()
    for b in (1, 2, 3):
        o.run_generation(b)
    path = str(tmp_path / "rays.e")
    exodus.write_ray_exodus(path, o.traces())
    coord, conn, d = exodus.read_ray_exodus(path)
    assert list(d) == list(exodus.RAY_VAR_ORDER) == [str(n) for n in g["names"]]  # same variables, same order
    mine = ray_facts.check(coord, conn, d, 30, 3, L)
    assert mine["rays"] == 30 and mine["segments"] > 300 and mine["same_energy"] > 100
    assert mine["one_step"] + mine["final_other"] == mine["same_energy"]


def test_isotropic_in_lab_direction_and_budget(oracle_bridge):
    """OMC physics.cpp:706-718 + distribution_multi.cpp:73-79: a scatter off an <isotropic> nuclide draws
    two numbers from the tracking stream -- phi = 2 pi xi_1 first, then mu = 2 xi_2 - 1 -- and the new
    direction is (mu, sqrt(1 - mu^2) cos phi, sqrt(1 - mu^2) sin phi).  The oracle's sampler is checked
    against that statement with the reference-pinned prn (libm build: bit for bit), and a run with every
    nuclide flagged (as in test/tests/study/materials.xml) must differ from the unflagged one only
    through the tracking stream."""
    import math
    from tests import problems
    L = oracle_bridge.lib(True)
    for seed0 in (1, 12345, 2 ** 61 + 7):
        s, t = C.c_uint64(seed0), C.c_uint64(seed0)
        uvw = np.zeros(3)
        L.orc_isotropic_direction(uvw.ctypes.data_as(C.POINTER(C.c_double)), C.byref(s))
        phi = 0.0 + (2.0 * math.pi - 0.0) * L.orc_prn(C.byref(t))
        mu = -1.0 + (1.0 - -1.0) * L.orc_prn(C.byref(t))
        assert s.value == t.value  # exactly two draws
        want = np.array([mu, math.sqrt(1.0 - mu * mu) * math.cos(phi), math.sqrt(1.0 - mu * mu) * math.sin(phi)])
        assert np.array_equal(uvw.view(np.uint64), want.view(np.uint64))
    recs = {}
    for name in ("infinite_medium", "infinite_medium_iso"):
        mesh, lib, bm, st = problems.PROBLEMS[name](n_particles=300)
        o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, st, tallies=False)
        o.sample_initial_source()
        o.run_generation(1)
        recs[name] = o.history_records().copy()
    a, b = recs["infinite_medium"], recs["infinite_medium_iso"]
    assert np.array_equal(a["seed_source"], b["seed_source"])
    assert not np.array_equal(a["seed_tracking"], b["seed_tracking"])
    assert not np.array_equal(a["n_segments"], b["n_segments"])


def test_initial_source_sites_from_reference_components(oracle_bridge):
    """Batch-1 source (OMC source.cpp:325-338,143-221): site i uses the stream init_seed(i + 1, STREAM_SOURCE) and
    draws SpatialBox::sample (3 numbers, x then y then z), isotropic_direction (2) and the Watt spectrum
    (0.988 MeV, 2.249e-6 /eV; resampled outside the data's energy range).  Composed here from the REFERENCE's own
    compiled SpatialBox, isotropic_direction and watt_spectrum, every site of the oracle's bank must match bit
    for bit (libm build; no rejection: the box lies inside the mesh and non-fissionable regions are accepted)."""
    from macaw_b200 import cstructs, meshgen, nucdata
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_box_sample"):
        pytest.skip("oracle/_ref not built with the distribution sources (no /root/reference on this box)")
    mesh = meshgen.infinite_medium(5)
    lib = nucdata.infinite_medium_library(G=300)
    box = (-4.5, -3.0, -2.0, 4.0, 3.5, 1.0)
    n = 500
    s = cstructs.make_settings(n, source_box=box, only_fissionable=False)
    o = oracle_bridge.Oracle(libm=True)
    o.set_mesh(mesh, np.array([0, 1], np.int32))
    o.set_library(lib, [0, 1])
    o.set_settings(s)
    o.sample_initial_source()
    bank = o.source_bank()
    e_min = max(nuc.energy[0] for nuc in lib.nuclides)
    e_max = min(nuc.energy[-1] for nuc in lib.nuclides)
    b6 = np.array(box)
    f64p = C.POINTER(C.c_double)
    r.ref_set_master_seed(1)
    for i in range(n):
        seed = C.c_uint64(r.ref_init_seed(i + 1, 1))  # STREAM_SOURCE = 1 (OMC random_lcg.h)
        xyz, u = np.zeros(3), np.zeros(3)
        r.ref_box_sample(b6.ctypes.data_as(f64p), C.byref(seed), xyz.ctypes.data_as(f64p))
        r.ref_isotropic_direction(u.ctypes.data_as(f64p), C.byref(seed))
        while True:
            E = r.ref_watt_spectrum(0.988e6, 2.249e-6, C.byref(seed))
            if e_min < E < e_max:
                break
        assert np.array_equal(bank["r"][i].view(np.uint64), xyz.view(np.uint64)), i
        assert np.array_equal(bank["u"][i].view(np.uint64), u.view(np.uint64)), i
        assert bank["E"][i] == E and bank["wgt"][i] == 1.0


def test_xs_lookup_against_reference_search_primitive(oracle_bridge):
    """Macroscopic lookup of the oracle (libm build) against a direct transcription of OMC nuclide.cpp:446-476
    (init_grid) and :646-691 (calculate_xs) whose energy-grid search is the REFERENCE's own compiled
    lower_bound_index (search.h:14-20), summed like material.cpp:790-829: bit for bit, including exact grid
    points, duplicated grid points and energies outside a nuclide's grid."""
    import math
    from macaw_b200 import cstructs, meshgen, nucdata
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_lower_bound_index"):
        pytest.skip("oracle/_ref not built with search.h (no /root/reference on this box)")
    lib = nucdata.make_library(9, G=300)
    dup = lib.nuclides[3]
    dup.energy[101] = dup.energy[100]  # the "rare case where two energy points are the same" (nuclide.cpp:662)
    lib.nuclides[5].energy[-1] *= 0.5  # a grid that ends early: E > energy.back() for the others' top range
    lib.nuclides[5].energy[:] = np.sort(lib.nuclides[5].energy)
    M = lib.n_log_bins
    o = oracle_bridge.Oracle(libm=True)
    o.set_mesh(meshgen.pin_cell_tet(4, 2), np.arange(4, dtype=np.int32))
    o.set_library(lib, [0, 1, 2, 3])
    o.set_settings(cstructs.make_settings(10))
    E_min = max(n.energy[0] for n in lib.nuclides)
    E_max = min(n.energy[-1] for n in lib.nuclides)
    spacing = math.log(E_max / E_min) / M
    step = (M * spacing - 0.0) / M  # xt::linspace(0.0, M*spacing, M+1)
    f64p = C.POINTER(C.c_double)
    grid_index = []
    for k, n in enumerate(lib.nuclides):
        gi, j = np.zeros(M + 1, np.int32), 0
        for b in range(M + 1):
            while math.log(n.energy[j + 1] / E_min) <= 0.0 + b * step:
                if j + 2 == n.energy.size:
                    break
                j += 1
            gi[b] = j
        assert np.array_equal(gi, o.grid_index(k, M)), k
        grid_index.append(gi)
    E, mat = problems.lookup_points(lib, 1500)
    E[2:6] = [dup.energy[100], dup.energy[99], dup.energy[102], lib.nuclides[5].energy[-1] * 1.5]
    got = o.xs_lookup(E, mat)
    for p in range(E.size):
        e = float(E[p])
        i_log = min(max(int(math.log(e / E_min) / spacing), 0), M - 1)
        tot = [0.0, 0.0, 0.0, 0.0]
        m = lib.materials[mat[p]]
        for i_nuc, dens in zip(m.nuclides, m.densities):
            n = lib.nuclides[i_nuc]
            en = np.ascontiguousarray(n.energy)
            if e < en[0]:
                ig = 0
            elif e > en[-1]:
                ig = en.size - 2
            else:
                lo, hi = int(grid_index[i_nuc][i_log]), int(grid_index[i_nuc][i_log + 1]) + 1
                ig = lo + r.ref_lower_bound_index(en[lo:].ctypes.data_as(f64p), hi - lo, e)
            if en[ig] == en[ig + 1]:
                ig += 1
            f = (e - float(en[ig])) / (float(en[ig + 1]) - float(en[ig]))
            cols = 4 if n.fissionable else 2
            for c in range(cols):
                micro = (1.0 - f) * float(n.xs[ig, c]) + f * float(n.xs[ig + 1, c])
                tot[c] += dens * micro
        assert np.array_equal(np.array(tot).view(np.uint64), got[p].view(np.uint64)), (p, e, mat[p])


def test_free_gas_target_velocity_against_reference_physics_cpp(oracle_bridge):
    """sample_cxs_target_velocity (OMC physics.cpp:955-1004: the constant-cross-section free-gas target sampler
    with its rejection loop, 5+ random numbers per try, then rotate_angle) -- the REFERENCE's own function, linked
    out of its physics.cpp (oracle/Makefile: stub hdf5.h + --gc-sections), against the oracle (libm build): same
    velocity bits and the same stream position for light and heavy targets, thermal to 10 eV."""
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_sample_cxs_target_velocity"):
        pytest.skip("oracle/_ref not built with physics.cpp (no /root/reference on this box)")
    L = oracle_bridge.lib(True)
    rng = np.random.default_rng(11)
    f64p = C.POINTER(C.c_double)
    kT = 8.617333262e-5 * 293.6
    n_draws = []
    for k in range(400):
        awr = float(rng.choice([0.9992, 15.858, 89.132, 236.006]))
        E = float(np.exp(rng.uniform(np.log(1e-4), np.log(10.0))))
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        seed0 = int(rng.integers(1, 2 ** 63))
        a, b = C.c_uint64(seed0), C.c_uint64(seed0)
        va, vb = np.zeros(3), np.zeros(3)
        L.orc_sample_cxs_target_velocity(awr, E, u.ctypes.data_as(f64p), kT, C.byref(a), va.ctypes.data_as(f64p))
        r.ref_sample_cxs_target_velocity(awr, E, u.ctypes.data_as(f64p), kT, C.byref(b), vb.ctypes.data_as(f64p))
        assert a.value == b.value, k
        assert np.array_equal(va.view(np.uint64), vb.view(np.uint64)), (k, va, vb)
        s, n = seed0, 0
        while s != a.value and n < 400:
            s, n = lcg(s), n + 1
        n_draws.append(n)
    assert min(n_draws) >= 6 and max(n_draws) > 6  # at least one try (5 or 6 draws) + the azimuth; rejections occur


def test_bank_sort_and_synchronize_against_reference_eigenvalue_cpp(oracle_bridge):
    """sort_fission_bank (OMC bank.cpp:60-104) and synchronize_bank (eigenvalue.cpp:71-311, single process) -- the
    REFERENCE's own functions working on the reference's own globals, linked out of bank.cpp / eigenvalue.cpp /
    simulation.cpp / settings.cpp (oracle/Makefile) -- fed the oracle's fission bank in shuffled order: the sorted
    bank and the sampled next-generation source bank must equal the oracle's byte for byte, in a generation that
    banks fewer sites than n_particles (padding / duplication branch) and in one that banks more (sampling)."""
    from macaw_b200.cstructs import SITE_DTYPE
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_sort_and_synchronize_bank"):
        pytest.skip("oracle/_ref not built with eigenvalue.cpp (no /root/reference on this box)")
    n = 3000
    mesh, lib, bm, s = problems.pin_cell_tet(n_particles=n)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    o.sample_initial_source()
    r.ref_set_master_seed(1)
    rng = np.random.default_rng(5)
    ratios = []
    for batch in (1, 2, 3):
        o.transport_generation(batch, 4)
        bank = o.fission_bank().copy()  # sorted by (parent, progeny)
        progeny = o.history_records()["n_progeny"].astype(np.int64)
        assert progeny.sum() == bank.size
        shuffled = bank[rng.permutation(bank.size)].copy()
        sorted_out = np.zeros(bank.size, SITE_DTYPE)
        source_out = np.zeros(n, SITE_DTYPE)
        got = r.ref_sort_and_synchronize_bank(shuffled.ctypes.data, bank.size, progeny.ctypes.data, n, 0, batch, 1, 1,
                                              sorted_out.ctypes.data, source_out.ctypes.data)
        assert got == n and sorted_out.tobytes() == bank.tobytes()
        o.L.orc_sample_bank(o.h, batch)
        o.L.orc_finish_sampling(o.h, 0, 1)
        mine = o.sampled()
        assert mine.size == n and mine.tobytes() == source_out.tobytes(), batch
        ratios.append(bank.size / n)
        o.set_source_bank(mine)
        o.L.orc_set_keff(o.h, 0.3 if batch == 1 else 1.0)  # sites ~ 1 / keff: batch 2 banks more than n, batch 3 fewer
    assert min(ratios) < 1.0 < max(ratios), ratios  # both branches of synchronize_bank were exercised


def test_keff_statistics_against_reference_eigenvalue_cpp(oracle_bridge):
    """calculate_average_keff (OMC eigenvalue.cpp:313-351), the reference's own function on its own globals, fed the
    oracle's k_generation sequence: keff and keff_std of every batch (inactive: last generation's k; active: running
    mean and standard deviation of the mean) equal the oracle's bit for bit."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_keff_statistics"):
        pytest.skip("oracle/_ref not built with eigenvalue.cpp (no /root/reference on this box)")
    mesh, lib, bm, s = problems.infinite_medium(n_particles=1500, n_inactive=2)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    o.sample_initial_source()
    res = [o.run_generation(b, 4) for b in range(1, 8)]
    kg = np.array([x.k_generation for x in res])
    keff, std = np.zeros(kg.size), np.zeros(kg.size)
    f64p = C.POINTER(C.c_double)
    r.ref_keff_statistics(kg.ctypes.data_as(f64p), kg.size, 2, keff.ctypes.data_as(f64p), std.ctypes.data_as(f64p))
    assert np.array_equal(keff.view(np.uint64), np.array([x.keff for x in res]).view(np.uint64))
    assert np.array_equal(std.view(np.uint64), np.array([x.keff_std for x in res]).view(np.uint64))
    assert keff[0] == kg[0] and keff[1] == kg[1] and std[2] == 0.0 and std[3] > 0.0


def test_generation_bookkeeping_against_reference_particle_and_eigenvalue_cpp(oracle_bridge):
    """End of generation on the reference's own Particle and globals: Particle::event_death (OMC particle.cpp:319-352) fed
    every history's k scores from the oracle's records in history order, the global-tally update MaCaw does (MC
    OpenMCStudy.C:411-428), initialize_generation, calculate_generation_keff, calculate_average_keff and accumulate_tallies
    (eigenvalue.cpp:49-69,313-351; tallies/tally.cpp) -- against the oracle's own generation results over 2 inactive + 4
    active batches: the three k sums, k_generation, keff and keff_std bit for bit."""
    from macaw_b200.cstructs import GenerationResult
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_generation_bookkeeping"):
        pytest.skip("oracle/_ref not built with particle.cpp (no /root/reference on this box)")
    n, n_inactive, n_gen = 1200, 2, 6
    mesh, lib, bm, s = problems.infinite_medium(n_particles=n, n_inactive=n_inactive)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    o.sample_initial_source()
    k3, results = np.zeros((n_gen, n, 3)), []
    for b in range(1, n_gen + 1):
        o.transport_generation(b, 4)
        rec = o.history_records()
        k3[b - 1] = np.stack([rec["k_collision"], rec["k_absorption"], rec["k_tracklength"]], axis=1)
        o.L.orc_sample_bank(o.h, b)
        o.L.orc_finish_sampling(o.h, 0, 1)
        res = GenerationResult()
        o.L.orc_finalize_generation(o.h, b, None, C.byref(res))
        results.append((res.k_collision, res.k_absorption, res.k_tracklength, res.k_generation, res.keff, res.keff_std,
                        res.total_weight))
        o.set_source_bank(o.sampled())
    f64p = C.POINTER(C.c_double)
    gt3, kg, keff, std = np.zeros((n_gen, 3)), np.zeros(n_gen), np.zeros(n_gen), np.zeros(n_gen)
    r.ref_generation_bookkeeping(n, n_inactive, n_gen, n, k3.ctypes.data_as(f64p), gt3.ctypes.data_as(f64p),
                                 kg.ctypes.data_as(f64p), keff.ctypes.data_as(f64p), std.ctypes.data_as(f64p))
    mine = np.array(results)
    assert np.all(mine[:, 6] == n)
    assert np.array_equal((gt3 / float(n)).view(np.uint64), np.ascontiguousarray(mine[:, :3]).view(np.uint64))
    for col, ref in ((3, kg), (4, keff), (5, std)):
        assert np.array_equal(np.ascontiguousarray(mine[:, col]).view(np.uint64), ref.view(np.uint64)), col
    assert std[n_inactive] == 0.0 and std[-1] > 0.0 and 0.1 < keff[-1] < 2.0


def test_tally_accumulate_against_reference_tally_cpp(oracle_bridge):
    """Tally::accumulate (OMC tallies/tally.cpp:639-670) on a reference Tally object, fed the VALUE arrays the oracle
    scored in three active batches: SUM, SUM_SQ (normalised by n_particles) and the realisation count equal the
    oracle's bit for bit."""
    from macaw_b200.cstructs import GenerationResult
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_tally_accumulate"):
        pytest.skip("oracle/_ref not built with tally.cpp (no /root/reference on this box)")
    n = 1500
    mesh, lib, bm, s = problems.infinite_medium(n_particles=n, n_inactive=1)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=True)
    o.sample_initial_source()
    values = []
    for b in range(1, 5):
        o.transport_generation(b, 4)
        v, _, _, _ = o.get_tally(1)
        if b > 1:  # batch 1 is inactive: nothing is scored
            values.append(v.copy())
            assert v.sum() > 0
        o.L.orc_sample_bank(o.h, b)
        o.L.orc_finish_sampling(o.h, 0, 1)
        res = GenerationResult()
        o.L.orc_finalize_generation(o.h, b, None, C.byref(res))
        o.set_source_bank(o.sampled())
    v_after, sm, sq, nr = o.get_tally(1)
    assert nr == 3 and not v_after.any()
    vals = np.ascontiguousarray(np.stack(values))
    f64p = C.POINTER(C.c_double)
    ref_sum, ref_sq = np.zeros(sm.size), np.zeros(sm.size)
    got = r.ref_tally_accumulate(vals.ctypes.data_as(f64p), 3, sm.size, n, ref_sum.ctypes.data_as(f64p), ref_sq.ctypes.data_as(f64p))
    assert got == 3
    assert np.array_equal(ref_sum.view(np.uint64), sm.view(np.uint64)) and np.array_equal(ref_sq.view(np.uint64), sq.view(np.uint64))


@pytest.mark.parametrize("use_universe", [0, 1])
@pytest.mark.parametrize("with_energy", [1, 0])
@pytest.mark.parametrize("tracklength", [0, 1])
def test_scoring_against_reference_tally_scoring_cpp(oracle_bridge, use_universe, with_energy, tracklength):
    """score_collision_tally -> score_general_ce (OMC tallies/tally_scoring.cpp:2354-2420, :516) with the reference's own
    CellFilter / UniverseFilter / EnergyFilter doing the binning and its Tally holding the results -- what MaCaw's
    CollisionKernel.C:185 event_collide() runs (particle.cpp:247) -- replayed over the same collision events as the oracle's scoring + binning: every
    (filter bin, score) VALUE equal bit for bit.  Events sit on energy-bin edges, outside the filter's range, and on the
    zero-absorption branch of the fission scores.  tracklength=1: score_tracklength_tally (:2289-2352) over (E, wgt,
    segment length) events instead (OpenMC's estimator; MaCaw's kernel leaves it as a TODO, CollisionKernel.C:147)."""
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_score_collision_events"):
        pytest.skip("oracle/_ref not built with tally_scoring.cpp (no /root/reference on this box)")
    L = oracle_bridge.lib(False)
    rng = np.random.default_rng(20 + 4 * tracklength + 2 * use_universe + with_energy)
    n_cells, n_uni, n = 37, 5, 6000
    edges = np.array([1e-5, 0.0253, 0.625, 1e3, 1e5, 2e7])
    cell = rng.integers(0, n_cells, n).astype(np.int32)
    uni = rng.integers(0, n_uni, n).astype(np.int32)
    E = np.exp(rng.uniform(np.log(2e-6), np.log(6e7), n))
    E[:6] = edges                          # on every edge, first and last included
    E[6], E[7] = np.nextafter(edges[0], 0), np.nextafter(edges[-1], 1e9)   # just outside
    w = rng.uniform(0.2, 1.7, n)
    macro = rng.uniform(0.05, 3.0, (n, 4))
    macro[::17, 1] = 0.0                   # no absorption: fission / nu-fission are skipped
    macro[::23, 2:] = 0.0                  # non-fissile material
    names = [b"flux", b"total", b"fission", b"nu-fission"]
    score_ids = np.array([0, 1, 2, 3], np.int32)   # MCW_SCORE_FLUX .. MCW_SCORE_NU_FISSION
    n_spatial = n_uni if use_universe else n_cells
    n_e = edges.size - 1 if with_energy else 1
    f64p = C.POINTER(C.c_double)
    ref = np.zeros(n_spatial * n_e * 4)
    dist = rng.uniform(1e-4, 3.0, n)
    head = (n_cells, n_uni, edges.ctypes.data_as(f64p), edges.size if with_energy else 0, use_universe,
            (C.c_char_p * 4)(*names), 4, n, cell.ctypes.data, uni.ctypes.data, E.ctypes.data_as(f64p), w.ctypes.data_as(f64p))
    tail = (macro.ctypes.data_as(f64p), ref.ctypes.data_as(f64p))
    if tracklength:
        got_n = r.ref_score_tracklength_events(*head, dist.ctypes.data_as(f64p), *tail)
    else:
        got_n = r.ref_score_collision_events(*head, *tail)
    assert got_n == ref.size
    mine = np.zeros_like(ref)
    spatial = uni if use_universe else cell
    assert L.orc_score_events(n_spatial, edges.ctypes.data_as(f64p), edges.size if with_energy else 0, score_ids.ctypes.data, 4,
                              n, spatial.ctypes.data, E.ctypes.data_as(f64p), w.ctypes.data_as(f64p),
                              dist.ctypes.data_as(f64p) if tracklength else None, macro.ctypes.data_as(f64p),
                              mine.ctypes.data_as(f64p)) == ref.size
    assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
    v = ref.reshape(n_spatial, n_e, 4)
    inside = (E >= edges[0]) & (E <= edges[-1]) if with_energy else np.ones(n, bool)
    flux = w * dist if tracklength else w / macro[:, 0]
    assert abs(v[..., 0].sum() - flux[inside].sum()) < 1e-9 * v[..., 0].sum()
    assert v[..., 2].sum() > 0 and v[..., 3].sum() > 0


def _reference_library(oracle_bridge, r, n_log_bins=500, G=600, tabular=False, thermal=False, urr=False, reactions=False,
                       temperatures=False, fission=False):
    """A 9-nuclide synthetic library loaded into the reference's own data objects; the oracle gets the reference's derived
    tables, so both sides hold the same numbers.  One material is given <isotropic> flags.  tabular: the nuclides carry
    tabulated angular distributions and ContinuousTabular laws, read by the reference's own AngleDistribution /
    ContinuousTabular constructors from the in-memory nuclide file."""
    from macaw_b200 import nucdata
    # tabular: 15 nuclides, so that every secondary law occurs (continuous tabular CM / LAB, correlated, Kalbach-Mann, N-body)
    lib = nucdata.make_library(n_nuclides=15 if tabular else 9, G=G, seed=3, n_log_bins=n_log_bins, tabular=tabular)
    lib.materials[3].isotropic = [1] + [0] * (len(lib.materials[3].nuclides) - 1)
    if reactions:
        # a list of scattering reactions after the first on every nuclide that has one: further levels, the continuum with
        # a non-integral multiplicity, (n,2n) -- read by the reference's own Reaction / ReactionProduct constructors
        nucdata.add_reactions(lib)
    if reactions == "evaporation":
        # ... some of them with Maxwell / evaporation / Watt spectra and energy-dependent multiplicities (Tabulated1D /
        # Polynomial yield_), read by the reference's MaxwellEnergy / Evaporation / WattEnergy constructors
        nucdata.add_evaporation_laws(lib)
    if tabular:
        # a fifth material holding every nuclide with an inelastic reaction at the same density, so that each secondary
        # law (continuous tabular CM / LAB, correlated CM / LAB, Kalbach-Mann CM / LAB, N-body) is sampled thousands of times
        inel = [i for i, n in enumerate(lib.nuclides) if n.inelastic.size]
        lib.materials.append(nucdata.Material("laws", inel, [0.01] * len(inel)))
    if thermal:  # S(a,b) tables, read by the reference's own ThermalScattering / ThermalData constructors
        nucdata.add_thermal_scattering(lib)
        lib.materials[3].isotropic = None
    if thermal == "temperatures":  # ... at 293.6 / 400 / 600 K each (ThermalScattering::kTs_, data_[t])
        nucdata.add_thermal_temperatures(lib)
    if urr:      # probability tables, read by the reference's own UrrData constructor
        nucdata.add_urr_tables(lib)
    if fission:
        # fission neutrons as evaluated files have them: tabulated total nu and yields, six delayed groups, Watt / Maxwell /
        # evaporation spectra with tabulated parameters and continuous tabular spectra -- read by the reference's own
        # ReactionProduct / Tabulated1D / WattEnergy / MaxwellEnergy / Evaporation / ContinuousTabular constructors
        nucdata.add_fission_data(lib)
    if temperatures:  # data sets at 600 K and 900 K with their own grids (and probability tables), read by the reference's
        nucdata.add_temperatures(lib)  # own Nuclide / Reaction / UrrData constructors for every temperature of `kTs`
    generator = [n.xs.copy() for n in lib.nuclides]
    tables = oracle_bridge.ref_load_library(r, lib)
    for n, tb in zip(lib.nuclides, tables):
        xs5, fissionable, n_precursor = tb[0], tb[1], tb[2]
        assert fissionable == n.fissionable and n_precursor == n.n_precursor
        n.xs = xs5.copy()
        if len(tb) > 4:  # the reference's derived tables of every temperature
            for x in [n] + list(n.more_temperatures):
                assert np.allclose(tb[4][id(x)][0][:, :4], x.xs[:, :4], rtol=1e-12, atol=0)
                x.xs = tb[4][id(x)][0].copy()
    return lib, generator, tables


def test_library_tables_and_grid_index_against_reference_nuclide_cpp(oracle_bridge):
    """The reference's own Nuclide constructor + create_derived (OMC nuclide.cpp:51-315, :317-390) and init_grid
    (:446-476), run on an in-memory image of a nuclide file (oracle/ref_hdf5_mem.*): the total / absorption / fission /
    nu-fission tables it derives from the reactions equal the ones macaw_b200.nucdata generates bit for bit (non-fissionable
    nuclides; for fissionable ones capture is handed over as absorption - fission, so only to rounding), and its
    logarithmic grid index equals the oracle's for every nuclide."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_collide_events"):
        pytest.skip("oracle/_ref not built with nuclide.cpp (no /root/reference on this box)")
    lib, generator, tables = _reference_library(oracle_bridge, r)
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    n_exact = 0
    for i, (n, gen, (xs5, _, _, gi)) in enumerate(zip(lib.nuclides, generator, tables)):
        if not n.fissionable:
            assert np.array_equal(xs5[:, :4].view(np.uint64), gen[:, :4].view(np.uint64)), n.name
            n_exact += 1
        else:
            assert np.allclose(xs5[:, :4], gen[:, :4], rtol=1e-12, atol=0) and xs5[:, 3].max() > 0
        assert np.array_equal(o.grid_index(i, lib.n_log_bins), gi), n.name
    assert n_exact >= 5


@pytest.mark.parametrize("tabular", [False, True, "thermal", "urr", "reactions", "temperatures", "fission", "evaporation",
                                     "thermal temperatures"])
@pytest.mark.parametrize("libm", [True, False])
def test_lookup_and_collision_against_reference_physics_cpp(oracle_bridge, libm, tabular):
    """The chain MaCaw's CollisionKernel.C:143,185 runs, on the reference's own compiled code and data objects:
    Material::calculate_xs -> Nuclide::calculate_xs (OMC material.cpp:773-829, nuclide.cpp:550-750) and
    Particle::event_collide -> collision -> sample_neutron_reaction with sample_nuclide, sample_fission,
    create_fission_sites, absorption, scatter, elastic_scatter (free-gas target), inelastic_scatter (LevelInelastic),
    sample_fission_neutron (Watt prompt / Maxwell delayed), isotropic-in-lab override and the URR stream advance
    (particle.cpp:219-288, physics.cpp:38-1120), replayed over 20 000 independent collisions in each of the four
    materials and held against the oracle's macro_xs + collide on the same states.  libm=True (the oracle built on
    glibc's log / sin / cos like the reference): macroscopic cross sections, outgoing energy and direction, alive flag,
    keff collision / absorption scores, sampled nuclide, all stream seeds, fission-site counts and every site equal BIT
    FOR BIT.  libm=False (the deterministic-math build the GPU is compared with): seeds, site counts, sampled nuclides,
    fates and the cross sections still equal exactly (same random-number consumption, same branches); energies and
    directions agree to 1e-12 (last-bit differences of the transcendental functions).
    tabular=True: the same with tabulated secondary distributions in the data -- AngleDistribution::sample (OMC
    distribution_angle.cpp:65-95) -> Tabular::sample (distribution.cpp:223-266) for the elastic and the inelastic
    reaction, ContinuousTabular::sample (distribution_energy.cpp:150-275; histogram and lin-lin, with and without
    discrete lines) in the centre-of-mass and in the laboratory system -- the reference's own objects built by its own
    constructors from the in-memory nuclide file, against the oracle's restatement."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_collide_events") or (tabular and not hasattr(r, "ref_stage_tables")) or \
            (tabular == "thermal" and not hasattr(r, "ref_add_thermal_table")):
        pytest.skip("oracle/_ref not built with physics.cpp + particle.cpp (no /root/reference on this box)")
    thermal, urr = tabular in ("thermal", "thermal temperatures"), tabular == "urr"
    if urr and not hasattr(r, "ref_stage_urr"):
        pytest.skip("oracle/_ref built without the probability-table doorway")
    thermal_T = tabular == "thermal temperatures"
    if thermal_T and not hasattr(r, "ref_stage_thermal_temperature"):
        pytest.skip("oracle/_ref built without the thermal-temperature doorway")
    temperatures = tabular == "temperatures"
    if temperatures and not hasattr(r, "ref_stage_temperature_set"):
        pytest.skip("oracle/_ref built without the temperature-set doorway")
    fission = tabular == "fission"
    evaporation = tabular == "evaporation"
    reactions = tabular == "reactions" or temperatures or fission or evaporation
    urr = urr or temperatures or fission
    lib, _, tables = _reference_library(oracle_bridge, r, tabular=tabular is True or reactions,
                                        thermal="temperatures" if thermal_T else thermal, urr=urr,
                                        reactions="evaporation" if evaporation else reactions, temperatures=temperatures,
                                        fission=fission)
    law_seen = np.zeros(8, np.int64)
    rx_seen = np.zeros(8, np.int64)
    o_reset = np.zeros(8, np.int64)
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(libm=libm), mesh, lib, bm, s, tallies=False)
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    n, keff = 20000, 1.02
    seen = {"absorbed": 0, "sites": 0, "delayed": 0, "inelastic": 0, "upscatter": 0}
    for i_mat, m in enumerate(lib.materials):
        rng = np.random.default_rng(5 + i_mat)
        E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), n))
        E[:8] = lib.nuclides[m.nuclides[0]].energy[[3, 4, 50, 51, 200, 300, 400, 500]]   # exactly on grid points
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1)[:, None]
        seeds = rng.integers(1, 2 ** 63, (n, 4), dtype=np.uint64)
        ids = np.arange(1, n + 1, dtype=np.int64)

        def run(f, *head, extra=()):
            macro, st, so = np.zeros((n, 4)), np.zeros((n, 8)), np.zeros((n, 4), np.uint64)
            ns, cap = np.zeros(n, np.int32), 3 * n
            sites, kappa = np.zeros((cap, 7)), np.zeros(n)
            nb = f(*head, n, E.ctypes.data_as(f64p), u.ctypes.data_as(f64p), seeds.ctypes.data_as(u64p), ids.ctypes.data, *extra,
                   macro.ctypes.data_as(f64p), st.ctypes.data_as(f64p), so.ctypes.data_as(u64p), ns.ctypes.data, cap,
                   sites.ctypes.data_as(f64p), kappa.ctypes.data_as(f64p))
            return macro, st, so, ns, sites[:nb], kappa
        sqrtkT = float(np.sqrt(lib.nuclides[0].kT))
        if thermal_T:
            # ThermalScattering::calculate_xs (thermal.cpp:149-171): the closer of the two temperatures that bound the
            # particle's: 0 K and 340 K -> 293.6 K, 347 K -> 400 K (midpoint 346.8), 501 K -> 600 K, 2000 K -> 600 K
            T = [0.0, 340.0, 347.0, 501.0, 2000.0][i_mat % 5]
            sqrtkT = float(np.sqrt(8.617333262e-5 * T))
            o.set_temperature(T)
        if temperatures:
            # the particle's temperature (MaCaw: the element's, MC CollisionKernel.C:109) picks every nuclide's data set
            # (nuclide.cpp:614-626 NEAREST): 0 K and 293.6 K -> the 293.6 K sets, 446 K -> 293.6 K (152.4 < 154), 447.5 K and
            # 700 K -> 600 K, 1200 K -> 900 K
            T = [0.0, 446.0, 700.0, 1200.0, 447.5][i_mat % 5]
            sqrtkT = float(np.sqrt(8.617333262e-5 * T))
            o.set_temperature(T)
        ref = run(r.ref_collide_events, i_mat, keff, extra=(sqrtkT,))
        o.L.orc_reaction_samples(o_reset.ctypes.data_as(C.POINTER(C.c_longlong)))
        mine = run(o.L.orc_collide_events, o.h, i_mat, keff)
        o.L.orc_reaction_samples(o_reset.ctypes.data_as(C.POINTER(C.c_longlong)))
        rx_seen += o_reset
        o.L.orc_law_samples(o_reset.ctypes.data_as(C.POINTER(C.c_longlong)))
        law_seen += o_reset
        assert np.array_equal(mine[0].view(np.uint64), ref[0].view(np.uint64))          # macroscopic cross sections
        if thermal_T and T > 346.8 and getattr(m, "sab", None):
            # ... the bound cross sections below energy_max really come from the hotter tables
            o.set_temperature(293.6)
            cold = run(o.L.orc_collide_events, o.h, i_mat, keff)
            assert (cold[0][:, 0] != mine[0][:, 0]).mean() > 0.2
            seen["hot tables"] = seen.get("hot tables", 0) + 1
        if temperatures and T > 447:
            # ... which really come from other data sets than at room temperature
            o.set_temperature(293.6)
            cold = run(o.L.orc_collide_events, o.h, i_mat, keff)
            assert (cold[0][:, 0] != mine[0][:, 0]).mean() > 0.9
            seen["hot sets"] = seen.get("hot sets", 0) + 1
        assert np.array_equal(mine[2], ref[2]) and np.array_equal(mine[3], ref[3])       # seeds, site counts
        assert mine[4].shape == ref[4].shape
        if libm:
            assert np.array_equal(mine[1].view(np.uint64), ref[1].view(np.uint64))
            assert np.array_equal(mine[4].view(np.uint64), ref[4].view(np.uint64))
        else:
            assert np.array_equal(mine[1][:, [4, 7]], ref[1][:, [4, 7]]) and np.array_equal(mine[4][:, 4:], ref[4][:, 4:])
            assert np.allclose(mine[1], ref[1], rtol=1e-12, atol=1e-14) and np.allclose(mine[4], ref[4], rtol=1e-12, atol=1e-14)
        # kappa-fission, collision estimator: q * sigma_f * N * flux summed nuclide by nuclide like the reference
        # (tally_scoring.cpp:1246-1255) -- bit for bit (the cross sections are bit-equal in both math builds)
        assert np.array_equal(mine[5].view(np.uint64), ref[5].view(np.uint64))
        seen["kappa"] = seen.get("kappa", 0) + int((ref[5] > 0).sum())
        st, sites = ref[1], ref[4]
        assert np.allclose(np.linalg.norm(st[:, 1:4], axis=1), 1.0)
        seen["absorbed"] += int((st[:, 4] == 0).sum())
        seen["sites"] += len(sites)
        seen["delayed"] += int((sites[:, 4] > 0).sum())
        if fission:
            for g in range(7):  # prompt and every delayed group: Watt / continuous, Maxwell, evaporation, continuous, ...
                seen[f"group {g}"] = seen.get(f"group {g}", 0) + int((sites[:, 4] == g).sum())
        seen["upscatter"] += int(((st[:, 0] > E) & (st[:, 4] > 0)).sum())
        seen["inelastic"] += int(((st[:, 0] < 0.5 * E) & (E > 1e6) & (st[:, 4] > 0)).sum())
        if evaporation:  # tabulated multiplicities between 1 and 1.6, a polynomial one just above 1
            seen["weight"] = seen.get("weight", 0) + int(((st[:, 4] > 1.0) & (st[:, 4] <= 1.6)).sum())
        elif reactions:  # st[4] = alive * weight: the continuum reaction of every third nuclide multiplies the weight by 1.25
            seen["weight"] = seen.get("weight", 0) + int((st[:, 4] == 1.25).sum())
            assert set(np.unique(st[:, 4])) <= {0.0, 1.0, 1.25}
    assert all(v > 0 for v in seen.values()), seen
    assert not thermal_T or seen.get("hot tables", 0) >= 2
    if evaporation:  # Maxwell (5), evaporation (6) and Watt (7) spectra were sampled, next to the other laws
        assert (law_seen[5:] >= 20).all(), law_seen
    if reactions:  # every position of the reaction list was sampled: MT 51, 52, 53, 91, 875
        assert (rx_seen[:5] >= 20).all() and rx_seen[5:].sum() == 0, rx_seen


@pytest.mark.parametrize("libm", [True, False])
def test_collision_fixture_from_reference_objects(oracle_bridge, libm):
    """tests/golden/ref_collisions.npz (made by tests/golden/make_ref_collisions.py where the reference tree exists):
    4 x 1000 collisions through the reference's own Material::calculate_xs + Particle::event_collide.  Needs neither
    /root/reference nor oracle/_ref, so it also holds on the GPU box and on a fresh clone: the oracle reproduces the
    reference's outputs bit for bit on glibc math, and to 1e-12 with identical seeds / counts / fates on its
    deterministic math."""
    import importlib.util
    import os
    from tests import problems
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_ref_collisions", os.path.join(here, "make_ref_collisions.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "ref_collisions.npz"))
    lib = gen.library()
    for i, n in enumerate(lib.nuclides):
        if n.fissionable:
            n.xs = g[f"xs_{i}"].copy()
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(libm=libm), mesh, lib, bm, s, tallies=False)
    for i in range(len(lib.nuclides)):
        assert np.array_equal(o.grid_index(i, lib.n_log_bins), g[f"grid_index_{i}"])
    for i_mat in range(len(lib.materials)):
        E, u, seeds, ids = gen.events(i_mat, lib)
        mine = gen.run_events(o.L.orc_collide_events, (o.h, i_mat, gen.KEFF), E, u, seeds, ids)
        ref = {k: g[f"m{i_mat}_{k}"] for k in mine}
        assert np.array_equal(mine["macro"].view(np.uint64), ref["macro"].view(np.uint64))
        assert np.array_equal(mine["seeds_out"], ref["seeds_out"]) and np.array_equal(mine["n_sites"], ref["n_sites"])
        assert mine["sites"].shape == ref["sites"].shape
        if libm:
            assert np.array_equal(mine["state"].view(np.uint64), ref["state"].view(np.uint64))
            assert np.array_equal(mine["sites"].view(np.uint64), ref["sites"].view(np.uint64))
        else:
            assert np.array_equal(mine["state"][:, [4, 7]], ref["state"][:, [4, 7]])
            assert np.allclose(mine["state"], ref["state"], rtol=1e-12, atol=1e-14)
            assert np.allclose(mine["sites"], ref["sites"], rtol=1e-12, atol=1e-14)
        assert np.array_equal(mine["kappa"].view(np.uint64), ref["kappa"].view(np.uint64))  # per-nuclide summation order


@pytest.mark.parametrize("physics", ["base", "full"])
@pytest.mark.parametrize("renormalise,libm", [(0, True), (1, True), (1, False)])
@pytest.mark.parametrize("i_mat", [0, 3])
def test_whole_histories_against_reference_history_loop(oracle_bridge, i_mat, renormalise, libm, physics):
    """Whole histories, all carried state: the reference's own initialize_history + transport_history_based_single_particle
    (OMC simulation.cpp:451,662-677: event_calculate_xs, event_advance, event_collide, event_revive_from_secondary,
    event_death) in a one-cell, no-surface geometry of one material, against the oracle's PRODUCTION history loop
    (run_history: locate, exit search, distance sample, collide) on a single 2e9-cm HEX8 element that no flight leaves, so
    that MaCaw's one-random-number-per-segment rule coincides with OpenMC's one-per-flight.  Oracle on glibc
    math, seed-through-double off (MaCaw's aux-data rounding, F3).  Per history: final energy, collision and progeny
    counts, all stream seeds and the three k scores bit for bit; the fission bank site by site (energy, direction,
    delayed group, parent / progeny ids; positions are the oracle's backward-from-exit points in a huge element and are
    not compared).  Material 0 is the fuel (fission, delayed neutrons); material 3 carries an <isotropic> nuclide.
    renormalise=0 switches off the one thing MaCaw does to the physics state that plain OpenMC does not -- Point::unit() on
    every new direction (RT Ray.C:216-226) -- and then everything is BIT-IDENTICAL.  renormalise=1 (the
    oracle's and the GPU's default, MaCaw's behaviour): the last-bit direction differences reach the energies (<= 1e-12
    relative; k scores 1e-10), while seeds, collision and progeny counts, bank size and ids stay identical.  libm=False:
    the deterministic-math build, i.e. exactly the configuration the GPU is bit-compared with -- LCG streams, counts and
    bank ids of whole histories still identical to the reference's (the north star's stream-parity bar, by transitivity
    GPU == oracle == OpenMC).
    physics="full": the same on a library with everything a real evaluation brings -- reaction lists with every secondary
    law and energy-dependent multiplicities (the weight changes along the history and with it every k score, the
    fission-site numbers and the tallies), fission products with function yields and tabulated spectra, probability
    tables, data sets at three temperatures, S(a,b) tables at three temperatures -- at a particle temperature of 700 K
    (the 600 K sets and tables).  Only (n,2n)'s
    integral multiplicity is set to 1 there: the reference's loop would go on to transport the second neutron, which
    MaCaw drops (MC CollisionKernel.C:232-236)."""
    from macaw_b200 import cstructs, meshgen
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_transport_histories"):
        pytest.skip("oracle/_ref not built with the history loop (no /root/reference on this box)")
    T = 0.0
    if physics == "full":
        if not hasattr(r, "ref_stage_temperature_set"):
            pytest.skip("oracle/_ref built without the temperature-set doorway")
        from macaw_b200 import nucdata
        _orig = nucdata.add_reactions

        def one_neutron(lib_, *a, **k):  # (n,2n) with multiplicity 1: no secondaries for the reference's loop to transport
            _orig(lib_, *a, **k)
            for nn in lib_.nuclides:
                for rx in (nn.more_reactions or []):
                    if rx.yield_ == 2.0:
                        rx.yield_ = 1.0
            return lib_
        nucdata.add_reactions = one_neutron
        try:
            lib, _, _ = _reference_library(oracle_bridge, r, tabular=True, reactions="evaporation", fission=True, urr=True,
                                           temperatures=True, thermal="temperatures")
        finally:
            nucdata.add_reactions = _orig
        T = 700.0
    else:
        lib, _, _ = _reference_library(oracle_bridge, r)
    n, batch, keff, master_seed, L = 1500, 3, 0.97, 11, 1.0e9
    mesh = meshgen.hex_grid(1, 1, 1, (-L, -L, -L), (L, L, L), meshgen.BC_REFLECT, None, "one_element")
    s = cstructs.make_settings(n, n_inactive=0, source_box=(-1, -1, -1, 1, 1, 1), seed=master_seed,
                               n_log_bins=lib.n_log_bins, seed_through_double=False)
    o = problems.configure(oracle_bridge.Oracle(libm=libm), mesh, lib, [i_mat], s, tallies=False)
    edges = np.array([1e-5, 0.0253, 0.625, 1e3, 1e5, 2e7])
    o.add_tally(1, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], edges)
    o.add_tally(2, cstructs.ESTIMATOR["tracklength"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], edges)
    rng = np.random.default_rng(77 + i_mat)
    sites = np.zeros(n, cstructs.SITE_DTYPE)
    sites["r"] = rng.uniform(-1, 1, (n, 3))
    u = rng.normal(size=(n, 3))
    sites["u"] = u / np.linalg.norm(u, axis=1)[:, None]
    sites["E"] = np.exp(rng.uniform(np.log(1e-2), np.log(1.0e7), n))
    sites["wgt"], sites["elem"] = 1.0, -1
    if physics == "full":
        o.set_temperature(T)
    o.set_source_bank(sites)
    o.L.orc_set_keff(o.h, keff)
    o.L.orc_set_renormalise_direction(o.h, renormalise)
    o.transport_generation(batch, 1)
    rec, bank = o.history_records(), o.fission_bank()
    assert int((rec["fate"] != 0).sum()) == 0 and int(rec["n_segments"].max()) > 20   # every history ends by absorption
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    site8 = np.ascontiguousarray(np.concatenate([sites["r"], sites["u"], sites["E"][:, None], sites["wgt"][:, None]], axis=1))
    state, seeds, cap = np.zeros((n, 8)), np.zeros((n, 4), np.uint64), 4 * n
    site7, ref_tallies = np.zeros((cap, 7)), np.zeros((2, edges.size - 1, 4))
    nb = r.ref_transport_histories(i_mat, keff, master_seed, batch, n, site8.ctypes.data_as(f64p),
                                   float(np.sqrt(8.617333262e-5 * T)) if physics == "full" else float(np.sqrt(lib.nuclides[0].kT)),
                                   state.ctypes.data_as(f64p), seeds.ctypes.data_as(u64p),
                                   cap, site7.ctypes.data_as(f64p), edges.ctypes.data_as(f64p), edges.size,
                                   ref_tallies.ctypes.data_as(f64p))
    r.ref_set_master_seed(1)
    assert np.all(state[:, 6] == 0)

    def same(a, b):
        a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
        return np.array_equal(a.view(np.uint64), b.view(np.uint64)) if not renormalise else np.allclose(a, b, rtol=1e-10, atol=1e-13)
    assert same(state[:, 0], rec["E"])
    assert np.array_equal(state[:, 1].astype(np.int64), rec["n_collisions"]) and np.array_equal(state[:, 1].astype(np.int64), rec["n_segments"])
    assert np.array_equal(state[:, 2].astype(np.int64), rec["n_progeny"])
    for col, f in ((0, "seed_tracking"), (1, "seed_source"), (2, "seed_urr")):
        assert np.array_equal(seeds[:, col], rec[f]), f
    for col, f in ((3, "k_collision"), (4, "k_absorption"), (5, "k_tracklength")):
        assert same(state[:, col], rec[f]), f
    # the oracle returns its bank sorted by (parent, progeny); the reference's is in creation order = the same order here
    assert nb == bank.size
    ref_sites = site7[:nb]
    assert np.array_equal(ref_sites[:, 5].astype(np.int64), bank["parent_id"]) and np.array_equal(ref_sites[:, 6].astype(np.int64), bank["progeny_id"])
    assert same(ref_sites[:, 0], bank["E"]) and same(ref_sites[:, 1:4], bank["u"])
    assert np.array_equal(ref_sites[:, 4].astype(np.int32), bank["delayed_group"])
    if i_mat == 0:
        assert nb > n // 2 and int((ref_sites[:, 4] > 0).sum()) > 0
    # tallies scored by the loops themselves, accumulated in history order on both sides (one oracle thread): collision
    # and track-length estimators x energy bins x (flux, total, fission, nu-fission)
    for t, tid in enumerate((1, 2)):
        v, _, _, _ = o.get_tally(tid)
        assert v.size == ref_tallies[t].size and ref_tallies[t][:, 0].sum() > 0
        assert same(v, ref_tallies[t].ravel()), tid


@pytest.mark.parametrize("i_mat", [0, 3])
def test_whole_history_fixture_from_reference_loop(oracle_bridge, i_mat):
    """tests/golden/ref_histories.npz (tests/golden/make_ref_histories.py): 600 whole histories per material through the
    reference's own history loop.  Needs neither /root/reference nor oracle/_ref: the oracle's production loop (glibc
    math, direction re-normalisation off) on one huge element reproduces final energies, stream seeds, collision and
    progeny counts, k scores, the fission bank and both tallies bit for bit."""
    import importlib.util
    import os
    from macaw_b200 import cstructs, meshgen
    from tests import problems
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_ref_histories", os.path.join(here, "make_ref_histories.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "ref_histories.npz"))
    lib, L = gen.library(), gen.HALF
    mesh = meshgen.hex_grid(1, 1, 1, (-L, -L, -L), (L, L, L), meshgen.BC_REFLECT, None, "one_element")
    s = cstructs.make_settings(gen.N, n_inactive=0, source_box=(-1, -1, -1, 1, 1, 1), seed=gen.MASTER_SEED,
                               n_log_bins=lib.n_log_bins, seed_through_double=False)
    o = problems.configure(oracle_bridge.Oracle(libm=True), mesh, lib, [i_mat], s, tallies=False)
    o.add_tally(1, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], gen.EDGES)
    o.add_tally(2, cstructs.ESTIMATOR["tracklength"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], gen.EDGES)
    o.set_source_bank(gen.sites(i_mat))
    o.L.orc_set_keff(o.h, gen.KEFF)
    o.L.orc_set_renormalise_direction(o.h, 0)
    o.transport_generation(gen.BATCH, 1)
    rec, bank = o.history_records(), o.fission_bank()
    state, seeds, sites, tallies = (g[f"m{i_mat}_{k}"] for k in ("state", "seeds", "sites", "tallies"))

    def bits(a):
        return np.ascontiguousarray(a, np.float64).view(np.uint64)
    assert np.array_equal(bits(state[:, 0]), bits(rec["E"]))
    assert np.array_equal(state[:, 1].astype(np.int64), rec["n_collisions"]) and np.array_equal(state[:, 2].astype(np.int64), rec["n_progeny"])
    for col, f in ((0, "seed_tracking"), (1, "seed_source"), (2, "seed_urr")):
        assert np.array_equal(seeds[:, col], rec[f]), f
    for col, f in ((3, "k_collision"), (4, "k_absorption"), (5, "k_tracklength")):
        assert np.array_equal(bits(state[:, col]), bits(rec[f])), f
    assert sites.shape[0] == bank.size
    assert np.array_equal(bits(sites[:, 0]), bits(bank["E"])) and np.array_equal(bits(sites[:, 1:4]), bits(bank["u"]))
    assert np.array_equal(sites[:, 4].astype(np.int32), bank["delayed_group"])
    assert np.array_equal(sites[:, 5].astype(np.int64), bank["parent_id"]) and np.array_equal(sites[:, 6].astype(np.int64), bank["progeny_id"])
    for t, tid in enumerate((1, 2)):
        v, _, _, _ = o.get_tally(tid)
        assert np.array_equal(bits(v), bits(tallies[t].ravel())), tid


def test_tracklength_kappa_fission_against_reference_objects(oracle_bridge):
    """kappa-fission on the track-length estimator (OMC tally_scoring.cpp:2289-2352 -> :1216-1258) through the reference's
    own Material::calculate_xs and score_tracklength_tally on its own Nuclide / Material / Tally objects: per segment equal
    to the oracle's macro_xs + tracklength_kappa BIT FOR BIT (same per-nuclide q * sigma_f * N * flux summation)."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_tracklength_kappa"):
        pytest.skip("oracle/_ref not built with tally_scoring.cpp + nuclide.cpp (no /root/reference on this box)")
    lib, _, _ = _reference_library(oracle_bridge, r)
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    f64p = C.POINTER(C.c_double)
    n, n_pos = 20000, 0
    for i_mat in range(len(lib.materials)):
        rng = np.random.default_rng(40 + i_mat)
        E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), n))
        w, d = rng.uniform(0.3, 1.5, n), rng.uniform(1e-4, 4.0, n)
        ref, mine = np.zeros(n), np.zeros(n)
        r.ref_tracklength_kappa(i_mat, n, E.ctypes.data_as(f64p), w.ctypes.data_as(f64p), d.ctypes.data_as(f64p),
                                float(np.sqrt(lib.nuclides[0].kT)), ref.ctypes.data_as(f64p))
        o.L.orc_tracklength_kappa(o.h, i_mat, n, E.ctypes.data_as(f64p), w.ctypes.data_as(f64p), d.ctypes.data_as(f64p),
                                  mine.ctypes.data_as(f64p))
        assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
        n_pos += int((ref > 0).sum())
    assert n_pos >= n   # the fuel material scores on every segment, the others never


@pytest.mark.parametrize("temperatures", [False, True])
@pytest.mark.parametrize("estimator", ["collision", "tracklength"])
@pytest.mark.parametrize("mode", ["all", "list"])
def test_nuclide_bins_against_reference_tally_objects(oracle_bridge, estimator, mode, temperatures):
    """Tallies with nuclide bins (MC OpenMCTally.C:188-204 `nuclides = all | list`): the reference's own Tally built the way
    OpenMCTally::initialize builds it (nuclides_ + all_nuclides_, or the list followed by total), scored by its own
    score_collision_tally / score_tracklength_tally -> score_all_nuclides / the nuclide-bin loop -> score_general_ce
    (OMC tally_scoring.cpp:2142-2175, :2381-2407, :516-1260) on its own Material / Nuclide objects, against the oracle's
    score_tally_nuclides: every (nuclide bin, score) VALUE of every event BIT FOR BIT -- flux, total, fission, nu-fission
    and kappa-fission; bins whose nuclide is not in the material stay zero on both sides.  temperatures=True: nuclides
    with data sets at three temperatures (and fission products with a tabulated nu), particles at 700 K: the bins are the
    NUCLIDES' whatever data set the temperature picked, on both sides."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_score_nuclide_bins"):
        pytest.skip("oracle/_ref not built with the nuclide-bin doorway (no /root/reference on this box)")
    if temperatures and not hasattr(r, "ref_stage_temperature_set"):
        pytest.skip("oracle/_ref built without the temperature-set doorway")
    lib, _, _ = _reference_library(oracle_bridge, r, tabular=temperatures, reactions=temperatures, fission=temperatures,
                                   temperatures=temperatures)
    T = 700.0 if temperatures else None
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, tallies=False)
    if temperatures:
        o.set_temperature(T)
    sqrtkT = float(np.sqrt(8.617333262e-5 * T)) if temperatures else float(np.sqrt(lib.nuclides[0].kT))
    f64p = C.POINTER(C.c_double)
    names = [b"flux", b"total", b"fission", b"nu-fission", b"kappa-fission"]
    scores = np.arange(5, dtype=np.int32)
    n_nuc = len(lib.nuclides)
    # "all": MaCaw appends to the default {-1} without clearing (OpenMCTally.C:190-197): n + 2 bins, the last never scored
    bins = list(range(n_nuc)) + [-1, -2] if mode == "all" else [4, 0, 7, 2, -1]  # list: no particular order, then total
    nb = np.asarray(bins, np.int32)
    n, n_pos = 4000, 0
    for i_mat in range(len(lib.materials)):
        rng = np.random.default_rng(70 + i_mat)
        E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), n))
        w, d = rng.uniform(0.3, 1.5, n), rng.uniform(1e-4, 4.0, n)
        dp = d.ctypes.data_as(f64p) if estimator == "tracklength" else None
        ref, mine = np.zeros((n, nb.size, 5)), np.zeros((n, nb.size, 5))
        r.ref_score_nuclide_bins(i_mat, int(mode == "all"), nb.ctypes.data, nb.size, (C.c_char_p * 5)(*names), 5, n,
                                 E.ctypes.data_as(f64p), w.ctypes.data_as(f64p), dp, sqrtkT, ref.ctypes.data_as(f64p))
        o.L.orc_score_nuclide_bins(o.h, i_mat, nb.ctypes.data, nb.size, scores.ctypes.data, 5, n, E.ctypes.data_as(f64p),
                                   w.ctypes.data_as(f64p), dp, mine.ctypes.data_as(f64p))
        assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
        held = set(lib.materials[i_mat].nuclides)
        for j, b in enumerate(bins):
            if b == -2 or (b >= 0 and b not in held):
                assert not ref[:, j].any()
            else:
                assert (ref[:, j, 0] > 0).all()  # flux scores in every bin the material feeds
        # the per-nuclide totals add up to the material total (to rounding)
        if mode == "all":
            assert np.allclose(ref[:, :-2, 1].sum(1), ref[:, -2, 1], rtol=1e-13)
        n_pos += int((ref[..., 4] > 0).sum())
    assert n_pos > 0


def test_history_identity_against_reference_initialize_history(oracle_bridge):
    """initialize_history (OMC simulation.cpp:451-481; MC OpenMCStudy.C:326), the reference's own function on its own
    Particle: id = work_index[rank] + index_source, seeds = init_particle_seeds((batch - 1) * n_particles + id) with the
    master seed of settings.xml, and from_source copying the site -- equal to the oracle's history_identity for 1 / 3 / 8
    ranks, several batches and sites (seed_through_double off: the rounding through aux data is MaCaw's)."""
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_initialize_history"):
        pytest.skip("oracle/_ref not built with particle.cpp (no /root/reference on this box)")
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    n = 1000
    for master_seed in (1, 7919):
        mesh, lib, bm, s = problems.infinite_medium(n_particles=n, seed=master_seed, seed_through_double=False)
        for nranks, rank in ((1, 0), (3, 1), (8, 7)):
            o = problems.configure(oracle_bridge.Oracle(rank=rank, nranks=nranks), mesh, lib, bm, s, tallies=False)
            for batch in (1, 2, 9):
                for i_local in (0, 1, 57, n // nranks - 1):
                    site = np.array([0.1, -0.2, 0.3, 0.0, 0.6, 0.8, 2.0e6 + i_local, 1.0])
                    rid, rseeds, state = C.c_int64(), np.zeros(4, np.uint64), np.zeros(8)
                    r.ref_initialize_history(n, nranks, rank, master_seed, batch, i_local + 1, site.ctypes.data_as(f64p),
                                             C.byref(rid), rseeds.ctypes.data_as(u64p), state.ctypes.data_as(f64p))
                    oid, oseeds = C.c_int64(), np.zeros(3, np.uint64)
                    o.L.orc_history_identity(o.h, i_local, batch, C.byref(oid), oseeds.ctypes.data_as(u64p))
                    assert oid.value == rid.value and np.array_equal(oseeds, rseeds[:3])
                    assert np.array_equal(state, site)
    r.ref_set_master_seed(1)


def test_work_split_against_reference_simulation_cpp(oracle_bridge):
    """calculate_work (OMC simulation.cpp:529-551), the reference's own function: the particle split used by the
    C ABI (mcw_calculate_work), by macaw_b200.dist and by the oracle for every rank count the study supports."""
    from macaw_b200 import dist, lib
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_calculate_work"):
        pytest.skip("oracle/_ref not built with simulation.cpp (no /root/reference on this box)")
    L = lib.load()
    for n in (1, 7, 30, 1000, 200000, 10 ** 8 + 3):
        for P in (1, 2, 3, 4, 8, 13):
            want = np.zeros(P + 1, np.int64)
            per_rank = [r.ref_calculate_work(n, P, k, want.ctypes.data) for k in range(P)]
            mine = np.zeros(P + 1, np.int64)
            assert L.mcw_calculate_work(n, P, mine.ctypes.data) == 0
            assert np.array_equal(mine, want) and np.array_equal(dist.calculate_work(n, P), want)
            assert per_rank == np.diff(want).tolist() and want[-1] == n


def test_full_fission_bank_keeps_the_bank_consistent(oracle_bridge):
    """More fission sites than the bank's 3 n (k about 4 under the first generation's keff guess 1): like
    create_fission_sites (OMC physics.cpp:203-209) the site that does not fit is not counted in n_progeny and no further
    ones are sampled for that collision, so sort_fission_bank's offsets stay inside the bank -- no holes, every parent's
    sites numbered 0, 1, 2, ... -- and the next generation is an ordinary one (tests/problems.py check_full_bank; the same
    properties are asserted of the GPU path in tests/test_gpu_parity.py::test_full_fission_bank)."""
    from tests import problems
    n = 3000
    mesh, lib, bm, s = problems.infinite_medium_supercritical(n_particles=n, n_inactive=1)
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s)
    o.sample_initial_source()
    r = o.run_generation(1)
    problems.check_full_bank(r, o.history_records(), o.fission_bank(), n)
    assert len(o.source_bank()) == n
    r2 = o.run_generation(2)
    assert r2.n_bank_dropped == 0 and 0.5 * n < r2.n_bank_sites < 2 * n


def _full_fixture():
    """tests/golden/ref_collisions_full.npz + its generator module + the library with the reference's derived tables"""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(here, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    gen, base = load("make_ref_collisions_full"), load("make_ref_collisions")
    g = np.load(os.path.join(here, "ref_collisions_full.npz"))
    lib = gen.library()
    for key, x in gen.data_sets(lib):
        if f"xs_{key}" in g:
            x.xs = g[f"xs_{key}"].copy()
    return gen, base, g, lib


@pytest.mark.parametrize("libm", [True, False])
def test_full_physics_fixture_from_reference_objects(oracle_bridge, libm):
    """tests/golden/ref_collisions_full.npz (tests/golden/make_ref_collisions_full.py, made where the reference tree
    exists): 5 x 1000 collisions through the reference's own Material::calculate_xs + Particle::event_collide on a
    library with reaction lists, every secondary law, energy-dependent multiplicities, fission products with function
    yields and tabulated spectra, and data sets at three temperatures, at particle temperatures that pick each set.  Needs
    neither /root/reference nor oracle/_ref: the oracle reproduces the reference's outputs bit for bit on glibc math --
    cross sections, outgoing energy / direction / weight, all seeds, fission sites -- and to 1e-12 with identical seeds /
    counts / fates on its deterministic math; every data set's logarithmic grid index equals Nuclide::init_grid's."""
    from tests import problems
    gen, base, g, lib = _full_fixture()
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    o = problems.configure(oracle_bridge.Oracle(libm=libm), mesh, lib, bm, s, tallies=False)
    entries = [str(i) for i in range(len(lib.nuclides))] + [f"{i}_{k}" for i, n in enumerate(lib.nuclides)
                                                           for k in range(len(n.more_temperatures or []))]
    for e, key in enumerate(entries):  # library entries: the nuclides, then the further sets in nuclide order
        assert np.array_equal(o.grid_index(e, lib.n_log_bins), g[f"grid_index_{key}"]), key
    weights = set()
    for i_mat in range(len(lib.materials)):
        E, u, seeds, ids = gen.events(i_mat, lib)
        o.set_temperature(gen.TEMPERATURES[i_mat])
        mine = base.run_events(o.L.orc_collide_events, (o.h, i_mat, gen.KEFF), E, u, seeds, ids)
        ref = {k: g[f"m{i_mat}_{k}"] for k in mine}
        assert np.array_equal(mine["macro"].view(np.uint64), ref["macro"].view(np.uint64))
        assert np.array_equal(mine["seeds_out"], ref["seeds_out"]) and np.array_equal(mine["n_sites"], ref["n_sites"])
        assert mine["sites"].shape == ref["sites"].shape
        if libm:
            assert np.array_equal(mine["state"].view(np.uint64), ref["state"].view(np.uint64))
            assert np.array_equal(mine["sites"].view(np.uint64), ref["sites"].view(np.uint64))
        else:
            assert np.array_equal(mine["state"][:, 7], ref["state"][:, 7]) and np.array_equal(mine["sites"][:, 4:], ref["sites"][:, 4:])
            assert np.allclose(mine["state"], ref["state"], rtol=1e-12, atol=1e-14)
            assert np.allclose(mine["sites"], ref["sites"], rtol=1e-12, atol=1e-14)
        assert np.array_equal(mine["kappa"].view(np.uint64), ref["kappa"].view(np.uint64))
        weights |= set(np.unique(ref["state"][:, 4]).tolist())
    assert len(weights) > 10 and max(weights) <= 1.6  # energy-dependent multiplicities moved the weight
    assert set(np.unique(np.concatenate([g[f"m{m}_sites"][:, 4] for m in range(5)])).tolist()) == set(range(7))  # all groups


def _same(a, b, path=""):
    """deep equality of what nucdata objects hold: arrays bit for bit, None == None, dataclasses field by field"""
    import dataclasses
    if path.endswith("n_discrete") and (a is None or b is None):
        assert (a is None or not np.any(a)) and (b is None or not np.any(b)), path  # no discrete lines, said either way
    elif a is None or b is None:
        assert a is None and b is None, path
    elif dataclasses.is_dataclass(a):
        for f in dataclasses.fields(a):
            _same(getattr(a, f.name), getattr(b, f.name), f"{path}.{f.name}")
    elif isinstance(a, (tuple, list)):
        assert len(a) == len(b), path
        for k, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{k}]")
    elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, (path, a.shape, b.shape)
        if a.dtype.kind == "f" or b.dtype.kind == "f":
            assert np.array_equal(a.astype(np.float64).view(np.uint64), b.astype(np.float64).view(np.uint64)), path
        else:
            assert np.array_equal(a, b), path
    else:
        assert a == b, (path, a, b)


def _everything_library(nucdata, G=600):
    lib = nucdata.make_library(n_nuclides=15, G=G, seed=3, n_log_bins=500, tabular=True)
    nucdata.add_evaporation_laws(nucdata.add_reactions(lib))
    nucdata.add_fission_data(lib)
    nucdata.add_urr_tables(lib)
    nucdata.add_temperatures(lib)
    return lib


def _check_converted(lib, conv):
    """the converted library equals the one the image was written from, field by field"""
    from macaw_b200 import nucdata
    assert len(conv.nuclides) == len(lib.nuclides)
    for n, c in zip(lib.nuclides, conv.nuclides):
        for x, y in zip([n] + list(n.more_temperatures or []), [c] + list(c.more_temperatures or [])):
            p = f"{n.name}@{x.kT:.4f}"
            assert y.kT == x.kT and y.awr == x.awr and y.fissionable == x.fissionable and y.n_precursor == x.n_precursor, p
            for f in ("energy", "elastic", "inelastic"):
                _same(getattr(x, f), getattr(y, f), f"{p}.{f}")
            assert np.array_equal(y.xs[:, :4].view(np.uint64), x.xs[:, :4].view(np.uint64)), p  # create_derived's tables
            assert y.inel_threshold == x.inel_threshold
            if x.inelastic.size:
                law = lambda z: z.inelastic_law if z.inelastic_law >= 0 else (1 if z.inelastic_energy is not None else 0)
                assert law(x) == law(y), p
                if law(x) == 0:
                    assert (y.inel_threshold_energy, y.inel_mass_ratio) == (x.inel_threshold_energy, x.inel_mass_ratio), p
                if law(x) != 0:
                    assert bool(y.inelastic_in_cm) == bool(x.inelastic_in_cm), p
                for f in ("inelastic_energy", "corr_angle", "km_r", "km_a", "nbody", "inelastic_yield_fn", "inelastic_a",
                          "inelastic_b"):
                    _same(getattr(x, f), getattr(y, f), f"{p}.{f}")
                if law(x) < 2 or law(x) >= 5:
                    _same(x.inelastic_angle, y.inelastic_angle, p + ".inelastic_angle")
                assert (x.inelastic_yield_fn is not None or y.inelastic_yield == x.inelastic_yield) and \
                    (law(x) < 5 or y.inelastic_u == x.inelastic_u), p
            _same(x.elastic_angle, y.elastic_angle, p + ".elastic_angle")
            assert len(x.more_reactions or []) == len(y.more_reactions or []), p
            for rx, ry in zip(x.more_reactions or [], y.more_reactions or []):
                q = f"{p}.MT{rx.mt}"
                assert (ry.mt, ry.threshold, ry.law) == (rx.mt, rx.threshold, rx.law), q
                assert rx.yield_fn is not None or ry.yield_ == rx.yield_, q  # a yield function stands in for the constant
                assert ry.law == 0 or bool(ry.in_cm) == bool(rx.in_cm), q
                if rx.law == 0:
                    assert (ry.threshold_energy, ry.mass_ratio) == (rx.threshold_energy, rx.mass_ratio), q
                if rx.law >= 5:
                    assert ry.u == rx.u, q
                for f in ("xs", "energy", "corr_angle", "km_r", "km_a", "nbody", "yield_fn", "a", "b"):
                    _same(getattr(rx, f), getattr(ry, f), f"{q}.{f}")
                if rx.law < 2 or rx.law >= 5:
                    _same(rx.angle, ry.angle, q + ".angle")
            if x.fissionable:
                assert y.q_fission == x.q_fission, p
                _same(x.total_nu, y.total_nu, p + ".total_nu")
                assert len(y.fission_products) == len(x.fission_products) == 7, p
                for k, (fx, fy) in enumerate(zip(x.fission_products, y.fission_products)):
                    assert fy.law == fx.law and (fx.law == 1 or fy.u == fx.u), (p, k)
                    for f in ("yield_", "a", "b", "energy"):
                        _same(getattr(fx, f), getattr(fy, f), f"{p}.fis{k}.{f}")
    ux = sorted(lib.urr_tables, key=lambda t: t.nuclide)
    uy = sorted(conv.urr_tables, key=lambda t: t.nuclide)
    assert [t.nuclide for t in ux] == [t.nuclide for t in uy]
    for a, b in zip(ux, uy):
        assert (a.inelastic_flag, a.absorption_flag, bool(a.multiply_smooth)) == (b.inelastic_flag, b.absorption_flag, bool(b.multiply_smooth))
        _same(a.energy, b.energy, "urr.energy")
        _same(a.table, b.table, "urr.table")


def test_nuclide_file_converter_against_reference_nuclide_cpp(oracle_bridge):
    """macaw_b200/openmc_hdf5.py, the converter from OpenMC's nuclide-file layout to the library's descriptors, on exactly the
    image the REFERENCE's own constructors read: a library with everything (reaction lists, every law, function yields,
    fission products, probability tables, three temperatures) is written as a nuclide-file image by oracle/_ref, read by
    the reference's Nuclide / Reaction / ReactionProduct / UrrData constructors -- and, through ref_dump_tree, by the
    converter.  The converter's derived total / absorption / fission / nu-fission tables equal the reference's own
    Nuclide::create_derived bit for bit at every temperature, its reaction lists, laws, tables, yields, fission products
    and probability tables equal what was written, and the converted library collides exactly like the original."""
    from macaw_b200 import nucdata, openmc_hdf5
    from tests import problems
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_dump_tree"):
        pytest.skip("oracle/_ref not built with the image dump (no /root/reference on this box)")
    lib = _everything_library(nucdata)
    tables = oracle_bridge.ref_load_library(r, lib)
    for n, tb in zip(lib.nuclides, tables):  # the reference's derived tables, every temperature
        n.xs = tb[0].copy()
        if len(tb) > 4:
            for x in [n] + list(n.more_temperatures):
                x.xs = tb[4][id(x)][0].copy()
    root = openmc_hdf5.read_image(oracle_bridge.ref_nuclide_file_image(r))
    conv = openmc_hdf5.library_from_groups([(n.name, root[n.name]) for n in lib.nuclides], lib.materials, lib.n_log_bins)
    _check_converted(lib, conv)
    # the converted library through the oracle: same collisions as the original, bit for bit
    mesh, _, bm, s = problems.infinite_medium(n_particles=100)
    rng = np.random.default_rng(9)
    n_ev = 3000
    E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), n_ev))
    u = rng.normal(size=(n_ev, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    seeds = rng.integers(1, 2 ** 63, (n_ev, 4), dtype=np.uint64)
    ids = np.arange(1, n_ev + 1, dtype=np.int64)
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_ref_collisions", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                       "golden", "make_ref_collisions.py"))
    base = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(base)
    outs = []
    for L in (lib, conv):
        o = problems.configure(oracle_bridge.Oracle(), mesh, L, bm, s, tallies=False)
        o.set_temperature(700.0)
        outs.append(base.run_events(o.L.orc_collide_events, (o.h, 0, 1.02), E, u, seeds, ids))
    for k in outs[0]:
        assert outs[0][k].tobytes() == outs[1][k].tobytes(), k
    # what the library does not carry is refused, not dropped
    bad = root["U235"]["reactions"]["reaction_018"]
    bad.attrs["mt"] = np.int64(19)
    with pytest.raises(openmc_hdf5.Refused):
        openmc_hdf5.nuclide_from_group(root["U235"], "U235")


def test_nuclide_file_converter_fixture():
    """tests/golden/nuclide_file_image.npz (tests/golden/make_nuclide_file_image.py, made where the reference tree exists):
    the bytes of a nuclide-file image the reference's own constructors read, and the tables its Nuclide::create_derived
    made of them.  Needs neither /root/reference nor oracle/_ref: macaw_b200/openmc_hdf5.py turns the image into a library
    whose derived tables equal the reference's bit for bit at every temperature and whose reactions, laws, yields, fission
    products and probability tables equal the library the image was written from."""
    import importlib.util
    import os
    from macaw_b200 import openmc_hdf5
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_nuclide_file_image", os.path.join(here, "make_nuclide_file_image.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "nuclide_file_image.npz"))
    lib = gen.library()
    for key, x in gen.data_sets(lib):
        x.xs = g[f"xs_{key}"].copy()
    root = openmc_hdf5.read_image(g["image"].tobytes())
    conv = openmc_hdf5.library_from_groups([(n.name, root[n.name]) for n in lib.nuclides], lib.materials, lib.n_log_bins)
    _check_converted(lib, conv)
    for t in lib.thermal_tables:  # the thermal-scattering tables of the image, every temperature
        c = openmc_hdf5.thermal_from_group(root[t.name], t.name)
        for x, y in zip([t] + list(t.more_temperatures), [c] + list(c.more_temperatures)):
            assert (y.kT, y.energy_max, y.inelastic_type, y.elastic_type) == (x.kT, x.energy_max, x.inelastic_type, x.elastic_type)
            for f in ("inelastic_xs", "inelastic_e_out", "inelastic_mu", "elastic_x", "elastic_y", "elastic_mu"):
                _same(getattr(x, f), getattr(y, f), f"{t.name}.{f}")


def test_nuclide_file_converter_refuses_what_the_library_lacks():
    """The converter never drops what the library cannot express: partial fission reactions, products with several
    distributions, Tabulated1D functions with logarithmic interpolation or several regions, log-log probability tables and
    unknown energy laws raise openmc_hdf5.Refused with the reason (image: tests/golden/nuclide_file_image.npz)."""
    import os
    from macaw_b200 import openmc_hdf5
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    image = np.load(os.path.join(here, "nuclide_file_image.npz"))["image"].tobytes()

    def refused(change, nuclide="U235"):
        root = openmc_hdf5.read_image(image)
        change(root[nuclide])
        with pytest.raises(openmc_hdf5.Refused) as e:
            openmc_hdf5.nuclide_from_group(root[nuclide], nuclide)
        return str(e.value)
    assert openmc_hdf5.nuclide_from_group(openmc_hdf5.read_image(image)["U235"], "U235")[0].fissionable  # unchanged: accepted

    def partial(g):
        g["reactions/reaction_018"].attrs["mt"] = np.int64(19)
    assert "partial fission" in refused(partial)

    def two_distributions(g):
        g["reactions/reaction_052/product_0"].attrs["n_distribution"] = np.int64(2)
    assert "several distributions" in refused(two_distributions)

    def log_interpolation(g):
        g["total_nu/yield"].attrs["interpolation"] = np.array([5])
    assert "logarithmic" in refused(log_interpolation)

    def two_regions(g):
        g["total_nu/yield"].attrs["interpolation"] = np.array([2, 1])
        g["total_nu/yield"].attrs["breakpoints"] = np.array([3, 8])
    assert "several regions" in refused(two_regions)
    # ... but regions that all interpolate the same way are one region
    root = openmc_hdf5.read_image(image)
    root["U235"]["total_nu/yield"].attrs["interpolation"] = np.array([2, 2])
    root["U235"]["total_nu/yield"].attrs["breakpoints"] = np.array([3, 8])
    same, _ = openmc_hdf5.nuclide_from_group(root["U235"], "U235")
    plain, _ = openmc_hdf5.nuclide_from_group(openmc_hdf5.read_image(image)["U235"], "U235")
    assert same.total_nu.interpolation == 2 and np.array_equal(same.xs, plain.xs)

    def loglog_urr(g):
        g["urr/294K"].attrs["interpolation"] = np.int64(5)
    assert "log-log" in refused(loglog_urr)

    def unknown_law(g):
        g["reactions/reaction_052/product_0/distribution_0/energy"].attrs["type"] = "madland-nix"
    assert "madland-nix" in refused(unknown_law)


def test_thermal_file_converter_against_reference_thermal_cpp(oracle_bridge):
    """openmc_hdf5.thermal_from_group on the thermal-scattering file image the REFERENCE's own ThermalScattering /
    ThermalData constructors read in oracle/_ref: the three kinds of tables (continuous inelastic; discrete inelastic +
    coherent elastic; skewed discrete inelastic + incoherent-discrete elastic) at three temperatures each come back as the
    tables the image was written from, array for array."""
    from macaw_b200 import nucdata, openmc_hdf5
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_dump_tree") or not hasattr(r, "ref_stage_thermal_temperature"):
        pytest.skip("oracle/_ref not built with the image dump (no /root/reference on this box)")
    lib = nucdata.make_library(n_nuclides=9, G=300, seed=3, n_log_bins=500)
    nucdata.add_thermal_temperatures(nucdata.add_thermal_scattering(lib))
    oracle_bridge.ref_load_library(r, lib)
    root = openmc_hdf5.read_image(oracle_bridge.ref_nuclide_file_image(r))
    for t in lib.thermal_tables:
        c = openmc_hdf5.thermal_from_group(root[t.name], t.name)
        assert c.applies_to == ["H1"]
        for x, y in zip([t] + list(t.more_temperatures), [c] + list(c.more_temperatures)):
            assert (y.kT, y.energy_max, y.inelastic_type, y.elastic_type, y.n_mu, y.n_e_out, bool(y.skewed)) == \
                (x.kT, x.energy_max, x.inelastic_type, x.elastic_type, x.n_mu, x.n_e_out, bool(x.skewed)), t.name
            for f in ("inelastic_xs_energy", "inelastic_xs", "inelastic_energy", "inelastic_offset", "inelastic_e_out",
                      "inelastic_pdf", "inelastic_cdf", "inelastic_mu", "elastic_x", "elastic_y", "elastic_mu"):
                _same(getattr(x, f), getattr(y, f), f"{t.name}.{f}")
    bad = root[lib.thermal_tables[1].name]
    bad[sorted(bad["kTs"].keys())[0]]["elastic"]["distribution"].attrs["type"] = "incoherent_elastic"
    with pytest.raises(openmc_hdf5.Refused):
        openmc_hdf5.thermal_from_group(bad)


def test_cpp_nuclide_reader_against_reference_nuclide_cpp(oracle_bridge):
    """host/openmc_nuclide_reader.hpp -- the MaCaw-side gatherLibrary() of INTEGRATION.md, written against OpenMC's own
    hdf5_interface.h -- on the nuclide-file image the REFERENCE's own constructors just read (oracle/_ref links the header's
    functions to an in-memory tree instead of libhdf5): the library's digest (mcw_digest: every scalar and every array of
    the descriptors) of what the C++ reader gathered equals the digest of the library the image was written from with the
    reference's own Nuclide::create_derived tables in it -- nuclides, further temperature sets and probability tables.
    So the adapter code a maintainer would add derives the same tables as OpenMC, bit for bit, and hands over every
    reaction list, law, yield and fission product unchanged."""
    from macaw_b200 import lib as mlib, nucdata
    r = oracle_bridge.ref_lib()
    if r is None or not hasattr(r, "ref_gather_digest"):
        pytest.skip("oracle/_ref not built with the C++ reader doorway (no /root/reference on this box)")
    lib = _everything_library(nucdata)
    tables = oracle_bridge.ref_load_library(r, lib)
    for n, tb in zip(lib.nuclides, tables):  # the reference's derived tables, every temperature
        n.xs = tb[0].copy()
        if len(tb) > 4:
            for x in [n] + list(n.more_temperatures):
                x.xs = tb[4][id(x)][0].copy()
    L = mlib.load()
    got, msg = (C.c_uint64 * 3)(), C.create_string_buffer(512)
    rc = r.ref_gather_digest(C.cast(L.mcw_digest, C.c_void_p), got, msg, 512, None, None, None, 0)
    assert rc == 0, msg.value.decode()

    def digest(kind, addr, n):
        h = C.c_uint64()
        assert L.mcw_digest(kind, addr, n, C.byref(h)) == 0
        return h.value
    nd, _, keep = nucdata.pack_library(lib)
    assert got[0] == digest(0, C.addressof(nd), len(lib.nuclides))
    sets, keep_s = nucdata.pack_temperature_sets(lib)
    arr = (nucdata.NuclideDesc * len(sets))()
    for k, (_, d) in enumerate(sets):
        C.memmove(C.addressof(arr[k]), C.addressof(d), C.sizeof(nucdata.NuclideDesc))
    assert len(sets) > 20 and got[1] == digest(0, C.addressof(arr), len(sets))
    ud, keep_u = nucdata.pack_urr(lib.urr_tables)
    assert got[2] == digest(2, C.addressof(ud), len(lib.urr_tables))
    # thermal-scattering files: three kinds of tables at three temperatures each, read by the reference's ThermalScattering
    lib2 = nucdata.make_library(n_nuclides=9, G=300, seed=3, n_log_bins=500)
    nucdata.add_thermal_temperatures(nucdata.add_thermal_scattering(lib2))
    oracle_bridge.ref_load_library(r, lib2)
    th, same, kT = (C.c_uint64 * 1)(), np.zeros(16, np.int32), np.zeros(16)
    rc = r.ref_gather_digest(C.cast(L.mcw_digest, C.c_void_p), got, msg, 512, th, same.ctypes.data, kT.ctypes.data, 16)
    assert rc == 0, msg.value.decode()
    tables, same_py, kT_py = lib2.thermal_entries()
    td, keep_t = nucdata.pack_thermal(tables)
    assert th[0] == digest(1, C.addressof(td), len(tables))
    assert np.array_equal(same[:len(tables)], same_py) and np.array_equal(kT[:len(tables)], kT_py)


@pytest.mark.parametrize("i_mat", [0, 3])
def test_whole_history_full_fixture_from_reference_loop(oracle_bridge, i_mat):
    """tests/golden/ref_histories_full.npz (tests/golden/make_ref_histories_full.py): 400 whole histories per material through
    the reference's own history loop on the library with everything at once (reaction lists and every law, multiplicities
    that move the weight, fission products, probability tables, nuclide and S(a,b) data at three temperatures), particles
    at 700 K.  Needs neither /root/reference nor oracle/_ref: the oracle's production loop (glibc math, direction
    re-normalisation off) on one huge element reproduces final energies, stream seeds, collision and progeny counts, k
    scores, the fission bank and both tallies bit for bit."""
    import importlib.util
    import os
    from macaw_b200 import cstructs, meshgen
    from tests import problems
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_ref_histories_full", os.path.join(here, "make_ref_histories_full.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "ref_histories_full.npz"))
    lib, L = gen.library(), gen.HALF
    for key, x in gen.data_sets(lib):
        if f"xs_{key}" in g:
            x.xs = g[f"xs_{key}"].copy()
    mesh = meshgen.hex_grid(1, 1, 1, (-L, -L, -L), (L, L, L), meshgen.BC_REFLECT, None, "one_element")
    s = cstructs.make_settings(gen.N, n_inactive=0, source_box=(-1, -1, -1, 1, 1, 1), seed=gen.MASTER_SEED,
                               n_log_bins=lib.n_log_bins, seed_through_double=False)
    o = problems.configure(oracle_bridge.Oracle(libm=True), mesh, lib, [i_mat], s, tallies=False)
    o.add_tally(1, cstructs.ESTIMATOR["collision"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], gen.EDGES)
    o.add_tally(2, cstructs.ESTIMATOR["tracklength"], cstructs.GRANULARITY["cell"], [0, 1, 2, 3], gen.EDGES)
    o.set_temperature(gen.TEMPERATURE)
    o.set_source_bank(gen.sites(i_mat))
    o.L.orc_set_keff(o.h, gen.KEFF)
    o.L.orc_set_renormalise_direction(o.h, 0)
    o.transport_generation(gen.BATCH, 1)
    rec, bank = o.history_records(), o.fission_bank()
    state, seeds, sites, tallies = (g[f"m{i_mat}_{k}"] for k in ("state", "seeds", "sites", "tallies"))

    def bits(a):
        return np.ascontiguousarray(a, np.float64).view(np.uint64)
    assert np.array_equal(bits(state[:, 0]), bits(rec["E"]))
    assert np.array_equal(state[:, 1].astype(np.int64), rec["n_collisions"]) and np.array_equal(state[:, 2].astype(np.int64), rec["n_progeny"])
    for col, f in ((0, "seed_tracking"), (1, "seed_source"), (2, "seed_urr")):
        assert np.array_equal(seeds[:, col], rec[f]), f
    for col, f in ((3, "k_collision"), (4, "k_absorption"), (5, "k_tracklength")):
        assert np.array_equal(bits(state[:, col]), bits(rec[f])), f
    assert len(sites) == bank.size
    assert np.array_equal(sites[:, 5].astype(np.int64), bank["parent_id"]) and np.array_equal(sites[:, 6].astype(np.int64), bank["progeny_id"])
    assert np.array_equal(bits(sites[:, 0]), bits(bank["E"])) and np.array_equal(bits(sites[:, 1:4]), bits(bank["u"]))
    assert np.array_equal(sites[:, 4].astype(np.int32), bank["delayed_group"])
    for t, tid in enumerate((1, 2)):
        v, _, _, _ = o.get_tally(tid)
        assert np.array_equal(bits(v), bits(tallies[t].ravel())), tid
