"""Regenerates the golden fixtures under tests/golden/ from the reference tree (run in the build
container, where /root/reference exists; the GPU box only reads the committed outputs).

  rt_lots_gold.json     total-distance known answers of the MOOSE ray-tracing `lots.i` regression
                        (moose/modules/ray_tracing/test/tests/traceray/{hex8,quad4}/gold/*.csv)
  study_rays_gold.npz   per-record ray aux data of MaCaw's own regression gold file
                        (test/tests/study/gold/openmc_study_rays.e, classic NetCDF Exodus)
  omc_math_kat.json     OpenMC unit-test known answers (openmc/tests/unit_tests/test_math.py:156-205)
  ref_rng_vectors.npz   outputs of the reference's own random_lcg.cpp / random_dist.cpp /
                        math_functions.cpp compiled into oracle/_ref/libomc_ref.so
"""
import csv
import ctypes as C
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def lots_gold():
    out = {}
    base = f"{REF}/moose/modules/ray_tracing/test/tests/traceray"
    for et in ("hex8", "quad4", "tet4"):
        d = f"{base}/{et}/gold"
        for f in sorted(os.listdir(d)):
            rows = list(csv.DictReader(open(f"{d}/{f}")))
            out[f"{et}/{f[:-8]}"] = {"expected_distance": float(rows[-1]["expected_distance"]),
                                     "total_distance": float(rows[-1]["total_distance"])}
    json.dump(out, open(f"{HERE}/rt_lots_gold.json", "w"), indent=1)
    print("rt_lots_gold.json", len(out))


def study_gold():
    from scipy.io import netcdf_file
    f = netcdf_file(f"{REF}/test/tests/study/gold/openmc_study_rays.e", "r", mmap=False)
    v = f.variables
    names = ["".join(c.decode() for c in row).strip("\x00 ").strip() for row in v["name_elem_var"][:]]
    coord = np.stack([v["coordx"][:], v["coordy"][:], v["coordz"][:]], axis=1)
    conn = v["connect1"][:].astype(np.int64) - 1
    data = {}
    for i, nm in enumerate(names):
        data[nm] = np.array(v[f"vals_elem_var{i + 1}eb1"][:][-1], np.float64)
    np.savez_compressed(f"{HERE}/study_rays_gold.npz", coord=coord, conn=conn, names=np.array(names),
                        **{f"var_{k}": a for k, a in data.items()})
    print("study_rays_gold.npz", names, coord.shape, conn.shape)


def math_kat():
    src = open(f"{REF}/openmc/tests/unit_tests/test_math.py").read()
    kat = {
        "source": "openmc/tests/unit_tests/test_math.py:156-205",
        "rotate_angle": {"uvw0": [1.0, 0.0, 0.0], "phi": 0.0, "mu": 0.0, "ref_phi0": [0.0, 0.0, -1.0],
                         "seed": 1, "ref_random": [0.9, -0.422746750548505, 0.10623175090659095],
                         "uvw0_random": [1.0, 0.0, 0.0], "mu_random": 0.9},
        "maxwell_spectrum": {"T": 0.5, "seed": 1, "ref": 0.27767406743161277},
        "watt_spectrum": {"a": 0.5, "b": 0.75, "seed": 1, "ref": 0.30957476387766697},
    }
    # make sure the numbers above are the ones in the reference file
    for num in ("0.27767406743161277", "0.30957476387766697", "-0.422746750548505", "0.10623175090659095"):
        assert num in src, num
    json.dump(kat, open(f"{HERE}/omc_math_kat.json", "w"), indent=1)
    print("omc_math_kat.json")


def ref_vectors():
    from oracle import bridge
    r = bridge.ref_lib()
    seed = C.c_uint64(1)
    prn = np.array([r.ref_prn(C.byref(seed)) for _ in range(1000)])
    ns = np.array([0, 1, 2, 3, 10, 152917, 2 ** 31 + 7, 2 ** 40 + 12345, 2 ** 63 + 5], np.uint64)
    fut = np.array([r.ref_future_seed(int(n), 12345) for n in ns], np.uint64)
    ids = np.array([1, 2, 61, 90, 1000003, 10 ** 12], np.int64)
    pseeds = np.zeros((ids.size, 4), np.uint64)
    for i, pid in enumerate(ids):
        buf = (C.c_uint64 * 4)()
        r.ref_init_particle_seeds(int(pid), buf)
        pseeds[i] = list(buf)
    s = C.c_uint64(7)
    maxw = np.array([r.ref_maxwell_spectrum(1.3e6, C.byref(s)) for _ in range(200)])
    s = C.c_uint64(7)
    watt = np.array([r.ref_watt_spectrum(0.988e6, 2.249e-6, C.byref(s)) for _ in range(200)])
    s = C.c_uint64(11)
    rot = np.zeros((200, 3))
    u = (C.c_double * 3)(0.6, 0.0, 0.8)
    for i in range(200):
        mu = r.ref_uniform(-1.0, 1.0, C.byref(s))
        r.ref_rotate_angle(u, mu, None, C.byref(s))
        rot[i] = list(u)
    np.savez(f"{HERE}/ref_rng_vectors.npz", prn=prn, future_n=ns, future_seed=fut, ids=ids, particle_seeds=pseeds,
             maxwell=maxw, watt=watt, rotate=rot)
    print("ref_rng_vectors.npz")


if __name__ == "__main__":
    lots_gold()
    study_gold()
    math_kat()
    ref_vectors()
