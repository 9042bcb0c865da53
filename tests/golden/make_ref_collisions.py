"""Regenerates tests/golden/ref_collisions.npz: collision events run through the REFERENCE's own compiled
Material::calculate_xs + Particle::event_collide on Nuclide / Material objects built by its own constructors from an
in-memory nuclide file of the synthetic library (oracle/_ref, oracle/ref_shim_physics.cpp).  Run in the build container,
where /root/reference exists; everywhere else (the GPU box, a fresh clone without the reference tree) the committed
fixture is what tests/test_oracle_pinning.py::test_collision_fixture_from_reference_objects holds the oracle against.

Contents: the reference-derived cross-section tables of the fissionable nuclides (the others equal the generator's bit
for bit), every nuclide's logarithmic grid index from Nuclide::init_grid, and per material the event inputs (E, u, seeds, id) and the reference's outputs (macroscopic cross sections,
outgoing state, seeds, fission-site counts, sites, kappa-fission scores)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
N_EVENTS, KEFF, G, N_LOG_BINS, LIB_SEED = 1000, 1.02, 600, 500, 3


def library():
    from macaw_b200 import nucdata
    lib = nucdata.make_library(n_nuclides=9, G=G, seed=LIB_SEED, n_log_bins=N_LOG_BINS)
    lib.materials[3].isotropic = [1] + [0] * (len(lib.materials[3].nuclides) - 1)
    return lib


def events(i_mat, lib):
    rng = np.random.default_rng(100 + i_mat)
    E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), N_EVENTS))
    E[:8] = lib.nuclides[lib.materials[i_mat].nuclides[0]].energy[[3, 4, 50, 51, 200, 300, 400, 500]]
    u = rng.normal(size=(N_EVENTS, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    seeds = rng.integers(1, 2 ** 63, (N_EVENTS, 4), dtype=np.uint64)
    return E, u, seeds, np.arange(1, N_EVENTS + 1, dtype=np.int64)


def run_events(f, head, E, u, seeds, ids, extra=()):
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    n = E.size
    macro, st, so = np.zeros((n, 4)), np.zeros((n, 8)), np.zeros((n, 4), np.uint64)
    ns, cap = np.zeros(n, np.int32), 3 * n
    sites, kappa = np.zeros((cap, 7)), np.zeros(n)
    nb = f(*head, n, E.ctypes.data_as(f64p), u.ctypes.data_as(f64p), seeds.ctypes.data_as(u64p), ids.ctypes.data, *extra,
           macro.ctypes.data_as(f64p), st.ctypes.data_as(f64p), so.ctypes.data_as(u64p), ns.ctypes.data, cap,
           sites.ctypes.data_as(f64p), kappa.ctypes.data_as(f64p))
    return dict(macro=macro, state=st, seeds_out=so, n_sites=ns, sites=sites[:nb].copy(), kappa=kappa)


def main():
    from oracle import bridge
    r = bridge.ref_lib()
    assert r is not None and hasattr(r, "ref_collide_events"), "needs oracle/_ref built from /root/reference"
    lib = library()
    tables = bridge.ref_load_library(r, lib)
    out = {}
    for i, (n, (xs5, _, _, gi)) in enumerate(zip(lib.nuclides, tables)):
        out[f"grid_index_{i}"] = gi   # Nuclide::init_grid
        if n.fissionable:
            out[f"xs_{i}"] = xs5
        else:
            assert np.array_equal(xs5[:, :4].view(np.uint64), n.xs[:, :4].view(np.uint64))
    for i_mat in range(len(lib.materials)):
        E, u, seeds, ids = events(i_mat, lib)
        res = run_events(r.ref_collide_events, (i_mat, KEFF), E, u, seeds, ids, extra=(float(np.sqrt(lib.nuclides[0].kT)),))
        for k, v in res.items():
            out[f"m{i_mat}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "ref_collisions.npz"), **out)
    print("wrote ref_collisions.npz:", {k: v.shape for k, v in out.items() if k.startswith("m0_") or k.startswith("xs_")})


if __name__ == "__main__":
    main()
