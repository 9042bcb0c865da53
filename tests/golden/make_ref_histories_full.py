"""Regenerates tests/golden/ref_histories_full.npz: whole histories through the REFERENCE's own history loop
(initialize_history + transport_history_based_single_particle, OMC simulation.cpp:451,662-677) like make_ref_histories.py,
on a library with everything at once -- tabulated angles, reaction lists with every secondary law and energy-dependent
multiplicities (the weight changes along a history), fission products with function yields and tabulated spectra,
probability tables, data sets at 293.6 / 600 / 900 K, S(a,b) tables at 293.6 / 400 / 600 K -- at a particle temperature
of 700 K.  (n,2n) carries multiplicity 1 here: the reference's loop would transport the second neutron, which MaCaw drops
(MC CollisionKernel.C:232-236).  Run in the build container, where /root/reference exists; everywhere else the committed
fixture is what tests/test_oracle_pinning.py::test_whole_history_full_fixture_from_reference_loop holds the oracle's
production loop against."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
N, BATCH, KEFF, MASTER_SEED, HALF, TEMPERATURE = 400, 3, 0.97, 11, 1.0e9, 700.0
EDGES = np.array([1e-5, 0.0253, 0.625, 1e3, 1e5, 2e7])
MATERIALS = (0, 3)


def library():
    from macaw_b200 import nucdata
    lib = nucdata.make_library(n_nuclides=15, G=600, seed=3, n_log_bins=500, tabular=True)
    nucdata.add_evaporation_laws(nucdata.add_reactions(lib))
    for n in lib.nuclides:
        for rx in (n.more_reactions or []):
            if rx.yield_ == 2.0:
                rx.yield_ = 1.0
    nucdata.add_thermal_temperatures(nucdata.add_thermal_scattering(lib))
    nucdata.add_fission_data(lib)
    nucdata.add_urr_tables(lib)
    nucdata.add_temperatures(lib)
    return lib


def data_sets(lib):
    out = []
    for i, n in enumerate(lib.nuclides):
        out.append((str(i), n))
        for k, m in enumerate(n.more_temperatures or []):
            out.append((f"{i}_{k}", m))
    return out


def sites(i_mat):
    from macaw_b200 import cstructs
    rng = np.random.default_rng(277 + i_mat)
    s = np.zeros(N, cstructs.SITE_DTYPE)
    s["r"] = rng.uniform(-1, 1, (N, 3))
    u = rng.normal(size=(N, 3))
    s["u"] = u / np.linalg.norm(u, axis=1)[:, None]
    s["E"] = np.exp(rng.uniform(np.log(1e-2), np.log(1.0e7), N))
    s["wgt"], s["elem"] = 1.0, -1
    return s


def main():
    from oracle import bridge
    r = bridge.ref_lib()
    assert r is not None and hasattr(r, "ref_stage_thermal_temperature"), "needs oracle/_ref built from /root/reference"
    lib = library()
    tables = bridge.ref_load_library(r, lib)
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    out = {}
    for i, (n, tb) in enumerate(zip(lib.nuclides, tables)):  # the reference's derived tables of the fissionable nuclides
        per = tb[4] if len(tb) > 4 else {id(n): (tb[0], tb[3])}
        for k, x in enumerate([n] + list(n.more_temperatures or [])):
            if n.fissionable:
                out["xs_" + (str(i) if k == 0 else f"{i}_{k - 1}")] = per[id(x)][0]
    sqrtkT = float(np.sqrt(8.617333262e-5 * TEMPERATURE))
    for i_mat in MATERIALS:
        s = sites(i_mat)
        site8 = np.ascontiguousarray(np.concatenate([s["r"], s["u"], s["E"][:, None], s["wgt"][:, None]], axis=1))
        state, seeds, cap = np.zeros((N, 8)), np.zeros((N, 4), np.uint64), 4 * N
        site7, tallies = np.zeros((cap, 7)), np.zeros((2, EDGES.size - 1, 4))
        nb = r.ref_transport_histories(i_mat, KEFF, MASTER_SEED, BATCH, N, site8.ctypes.data_as(f64p), sqrtkT,
                                       state.ctypes.data_as(f64p), seeds.ctypes.data_as(u64p), cap, site7.ctypes.data_as(f64p),
                                       EDGES.ctypes.data_as(f64p), EDGES.size, tallies.ctypes.data_as(f64p))
        out[f"m{i_mat}_state"], out[f"m{i_mat}_seeds"] = state, seeds
        out[f"m{i_mat}_sites"], out[f"m{i_mat}_tallies"] = site7[:nb].copy(), tallies
    r.ref_set_master_seed(1)
    np.savez_compressed(os.path.join(HERE, "ref_histories_full.npz"), **out)
    print("wrote ref_histories_full.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
