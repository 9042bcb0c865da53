"""Structural facts of a RayTracingExodus file of the reference's study test (test/tests/study/openmc_study.i on the
5 x 5 x 1 HEX8 grid on [-5, 5]^3) that need no nuclear data -- they hold for the reference's gold file
test/tests/study/gold/openmc_study_rays.e and must hold for a file this repository writes."""
import ctypes as C

import numpy as np

STRIDE = 152917
MULT, ADD = 6364136223846793005, 1442695040888963407


def lcg(seed):
    return (MULT * seed + ADD) % (1 << 64)


def check(coord, conn, d, n_particles, last_batch, L):
    """L = the oracle library (seed helpers).  Returns counters for the caller to compare / report."""
    n_elem = conn.shape[0]
    rid = d["ray_id"]
    rays = np.unique(rid)
    # one polyline per ray: nodes = segments + rays, consecutive node numbering inside a ray
    assert coord.shape[0] == n_elem + rays.size and rays.size == n_particles
    assert rays.min() == (last_batch - 1) * n_particles and rays.max() == last_batch * n_particles - 1
    assert np.all(conn[:, 1] == conn[:, 0] + 1)
    assert np.all(d["aux_weight"] == 1.0) and np.all(d["aux_particle_type"] == 0)
    ends = coord[conn[:, 1]]
    on_face = (np.isclose(np.abs(ends[:, 0]) % 2, 1, atol=1e-9) | np.isclose(np.abs(ends[:, 1]) % 2, 1, atol=1e-9)
               | np.isclose(np.abs(ends[:, 2]), 5, atol=1e-9))
    assert on_face.all()  # collision points are not cached
    n_same = n_one = n_final_other = 0
    for r in rays:
        idx = np.nonzero(rid == r)[0]
        assert np.array_equal(d["intersections"][idx], np.arange(idx.size))
        start = coord[conn[idx[0], 0]]
        assert np.all(np.abs(start) < 5)  # sources are interior points
        pid = int(d["aux_id"][idx[0]])
        assert np.all(d["aux_id"][idx] == pid) and pid == int(r) - (last_batch - 1) * n_particles + 1
        ss = d["aux_seed_source"][idx]
        seed_id = ((last_batch - 1) * n_particles + pid) * STRIDE
        assert np.all(ss == ss[0]) and float(L.orc_seed_via_double(L.orc_future_seed(seed_id, 1 + 1))) == ss[0]
        assert np.all(np.diff(d["aux_n_progeny"][idx]) >= 0)
        E, st = d["aux_energy"][idx], d["aux_seed_tracking"][idx]
        for k in range(idx.size - 1):
            if E[k] == E[k + 1]:
                n_same += 1
                one = float(L.orc_seed_via_double(lcg(int(st[k])))) == st[k + 1]
                n_one += one
                if not one:
                    assert k + 1 == idx.size - 1  # only a ray's last record (absorbed) may consume more numbers
                    n_final_other += 1
    return {"segments": n_elem, "rays": int(rays.size), "same_energy": n_same, "one_step": n_one, "final_other": n_final_other}
