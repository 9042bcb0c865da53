"""Regenerates tests/golden/nuclide_file_image.npz: the nuclide-file image oracle/_ref writes for a synthetic library
with everything (reaction lists, every secondary law, function yields, fission products, probability tables, three
temperatures, S(a,b) tables at three temperatures; 15 nuclides x 200 grid points) -- the bytes the REFERENCE's own Nuclide constructors read -- together
with the tables Nuclide::create_derived made of it at every temperature.  Run in the build container, where
/root/reference exists; tests/test_oracle_pinning.py::test_nuclide_file_converter_fixture then holds
macaw_b200/openmc_hdf5.py against it anywhere."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
G = 200


def library():
    from macaw_b200 import nucdata
    lib = nucdata.make_library(n_nuclides=15, G=G, seed=3, n_log_bins=500, tabular=True)
    nucdata.add_evaporation_laws(nucdata.add_reactions(lib))
    nucdata.add_fission_data(lib)
    nucdata.add_urr_tables(lib)
    nucdata.add_temperatures(lib)
    nucdata.add_thermal_temperatures(nucdata.add_thermal_scattering(lib))  # three S(a,b) tables x three temperatures
    return lib


def data_sets(lib):
    out = []
    for i, n in enumerate(lib.nuclides):
        out.append((str(i), n))
        for k, m in enumerate(n.more_temperatures or []):
            out.append((f"{i}_{k}", m))
    return out


def main():
    from oracle import bridge
    r = bridge.ref_lib()
    assert r is not None and hasattr(r, "ref_dump_tree"), "needs oracle/_ref built from /root/reference"
    lib = library()
    tables = bridge.ref_load_library(r, lib)
    out = {"image": np.frombuffer(bridge.ref_nuclide_file_image(r), np.uint8)}
    for i, (n, tb) in enumerate(zip(lib.nuclides, tables)):
        per = tb[4] if len(tb) > 4 else {id(n): (tb[0], tb[3])}
        for k, x in enumerate([n] + list(n.more_temperatures or [])):
            out["xs_" + (str(i) if k == 0 else f"{i}_{k - 1}")] = per[id(x)][0]
    np.savez_compressed(os.path.join(HERE, "nuclide_file_image.npz"), **out)
    print("wrote nuclide_file_image.npz:", out["image"].size, "image bytes,", len(out) - 1, "derived tables")


if __name__ == "__main__":
    main()
