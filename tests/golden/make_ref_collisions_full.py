"""Regenerates tests/golden/ref_collisions_full.npz: like make_ref_collisions.py, on a library with everything a real
evaluation brings except S(a,b) and probability tables (those are pinned live in tests/test_oracle_pinning.py) --
tabulated angular distributions, reaction lists with level / continuous tabular / correlated / Kalbach-Mann / N-body /
Maxwell / evaporation / Watt products and constant or energy-dependent multiplicities, fission neutrons with function
yields, six delayed groups and tabulated spectra, and data sets at 293.6 / 600 / 900 K with their own grids -- run
through the REFERENCE's own Material::calculate_xs + Particle::event_collide at particle temperatures that pick each of
the three sets.  Run in the build container, where /root/reference exists; everywhere else the committed fixture is what
tests/test_oracle_pinning.py::test_full_physics_fixture_from_reference_objects holds the oracle against, and
tests/test_gpu_parity.py::test_xs_lookup_at_temperatures_against_reference_fixture the GPU lookup.

Contents: the reference-derived tables of every data set of the fissionable nuclides, every data set's logarithmic grid
index, and per material the event inputs' outputs (macroscopic cross sections, outgoing state with the weight, seeds,
fission-site counts, sites, kappa-fission scores)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
N_EVENTS, KEFF, G, N_LOG_BINS, LIB_SEED = 1000, 1.02, 600, 500, 3
TEMPERATURES = [0.0, 446.0, 700.0, 1200.0, 447.5]  # per material: -> 293.6 K, 293.6 K, 600 K, 900 K, 600 K sets


def library():
    from macaw_b200 import nucdata
    lib = nucdata.make_library(n_nuclides=15, G=G, seed=LIB_SEED, n_log_bins=N_LOG_BINS, tabular=True)
    lib.materials[3].isotropic = [1] + [0] * (len(lib.materials[3].nuclides) - 1)
    nucdata.add_evaporation_laws(nucdata.add_reactions(lib))
    inel = [i for i, n in enumerate(lib.nuclides) if n.inelastic.size]
    lib.materials.append(nucdata.Material("laws", inel, [0.01] * len(inel)))
    nucdata.add_fission_data(lib)
    nucdata.add_temperatures(lib)
    return lib


def data_sets(lib):
    """[(key, data set)] of every nuclide's data sets: key = "<nuclide>" or "<nuclide>_<k>" for its k-th further set"""
    out = []
    for i, n in enumerate(lib.nuclides):
        out.append((str(i), n))
        for k, m in enumerate(n.more_temperatures or []):
            out.append((f"{i}_{k}", m))
    return out


def events(i_mat, lib):
    rng = np.random.default_rng(300 + i_mat)
    E = np.exp(rng.uniform(np.log(1.2e-5), np.log(1.9e7), N_EVENTS))
    u = rng.normal(size=(N_EVENTS, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    seeds = rng.integers(1, 2 ** 63, (N_EVENTS, 4), dtype=np.uint64)
    return E, u, seeds, np.arange(1, N_EVENTS + 1, dtype=np.int64)


def main():
    from oracle import bridge
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_collisions", os.path.join(HERE, "make_ref_collisions.py"))
    base = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(base)
    r = bridge.ref_lib()
    assert r is not None and hasattr(r, "ref_stage_temperature_set"), "needs oracle/_ref built from /root/reference"
    lib = library()
    tables = bridge.ref_load_library(r, lib)
    out = {}
    for i, (n, tb) in enumerate(zip(lib.nuclides, tables)):
        sets = [n] + list(n.more_temperatures or [])
        per = tb[4] if len(tb) > 4 else {id(n): (tb[0], tb[3])}
        for k, x in enumerate(sets):
            key = str(i) if k == 0 else f"{i}_{k - 1}"
            xs5, gi = per[id(x)]
            out[f"grid_index_{key}"] = gi
            if n.fissionable:
                out[f"xs_{key}"] = xs5
            else:
                assert np.array_equal(xs5[:, :4].view(np.uint64), x.xs[:, :4].view(np.uint64)), key
    for i_mat in range(len(lib.materials)):
        E, u, seeds, ids = events(i_mat, lib)
        sqrtkT = float(np.sqrt(8.617333262e-5 * TEMPERATURES[i_mat]))
        res = base.run_events(r.ref_collide_events, (i_mat, KEFF), E, u, seeds, ids, extra=(sqrtkT,))
        for k, v in res.items():
            out[f"m{i_mat}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "ref_collisions_full.npz"), **out)
    print("wrote ref_collisions_full.npz:", len(out), "arrays,", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
