"""Regenerates tests/golden/ref_histories.npz: whole histories run through the REFERENCE's own history loop
(initialize_history + transport_history_based_single_particle, OMC simulation.cpp:451,662-677) on Nuclide / Material
objects built by its own constructors, in a one-cell geometry of one material (oracle/ref_shim_physics.cpp
ref_transport_histories).  Run in the build container, where /root/reference exists; everywhere else the committed fixture
is what tests/test_oracle_pinning.py::test_whole_history_fixture_from_reference_loop holds the oracle's production loop
against.  The library is the one of make_ref_collisions.py (its reference-derived tables are read from
ref_collisions.npz)."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
N, BATCH, KEFF, MASTER_SEED, HALF = 600, 3, 0.97, 11, 1.0e9
EDGES = np.array([1e-5, 0.0253, 0.625, 1e3, 1e5, 2e7])
MATERIALS = (0, 3)


def collisions_module():
    spec = importlib.util.spec_from_file_location("make_ref_collisions", os.path.join(HERE, "make_ref_collisions.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def library():
    """The make_ref_collisions library with the reference-derived tables of the fissionable nuclides."""
    lib = collisions_module().library()
    g = np.load(os.path.join(HERE, "ref_collisions.npz"))
    for i, n in enumerate(lib.nuclides):
        if n.fissionable:
            n.xs = g[f"xs_{i}"].copy()
    return lib


def sites(i_mat):
    from macaw_b200 import cstructs
    rng = np.random.default_rng(177 + i_mat)
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
    assert r is not None and hasattr(r, "ref_transport_histories"), "needs oracle/_ref built from /root/reference"
    lib = collisions_module().library()
    bridge.ref_load_library(r, lib)
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    out = {}
    for i_mat in MATERIALS:
        s = sites(i_mat)
        site8 = np.ascontiguousarray(np.concatenate([s["r"], s["u"], s["E"][:, None], s["wgt"][:, None]], axis=1))
        state, seeds, cap = np.zeros((N, 8)), np.zeros((N, 4), np.uint64), 4 * N
        site7, tallies = np.zeros((cap, 7)), np.zeros((2, EDGES.size - 1, 4))
        nb = r.ref_transport_histories(i_mat, KEFF, MASTER_SEED, BATCH, N, site8.ctypes.data_as(f64p),
                                       float(np.sqrt(lib.nuclides[0].kT)), state.ctypes.data_as(f64p),
                                       seeds.ctypes.data_as(u64p), cap, site7.ctypes.data_as(f64p), EDGES.ctypes.data_as(f64p),
                                       EDGES.size, tallies.ctypes.data_as(f64p))
        out[f"m{i_mat}_state"], out[f"m{i_mat}_seeds"] = state, seeds
        out[f"m{i_mat}_sites"], out[f"m{i_mat}_tallies"] = site7[:nb].copy(), tallies
    r.ref_set_master_seed(1)
    np.savez_compressed(os.path.join(HERE, "ref_histories.npz"), **out)
    print("wrote ref_histories.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
