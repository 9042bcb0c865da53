"""ctypes mirrors of the plain-C structs in include/macaw_b200.h (kept field-for-field in sync;
tests/test_abi.py checks the sizes against the compiled library)."""
import ctypes as C

import numpy as np


class SourceSite(C.Structure):
    _fields_ = [("r", C.c_double * 3), ("u", C.c_double * 3), ("E", C.c_double), ("wgt", C.c_double),
                ("delayed_group", C.c_int32), ("surf_id", C.c_int32), ("particle", C.c_int32), ("elem", C.c_int32),
                ("parent_id", C.c_int64), ("progeny_id", C.c_int64)]


SITE_DTYPE = np.dtype([("r", "f8", 3), ("u", "f8", 3), ("E", "f8"), ("wgt", "f8"), ("delayed_group", "i4"),
                       ("surf_id", "i4"), ("particle", "i4"), ("elem", "i4"), ("parent_id", "i8"),
                       ("progeny_id", "i8")])
assert SITE_DTYPE.itemsize == 96 == C.sizeof(SourceSite)


class Settings(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("master_seed", C.c_int64), ("n_inactive", C.c_int32),
                ("n_log_bins", C.c_int32), ("is_2d", C.c_int32), ("seed_through_double", C.c_int32),
                ("z_coord", C.c_double), ("source_box", C.c_double * 6), ("source_only_fissionable", C.c_int32),
                ("max_events", C.c_int32), ("sort_xs_queue", C.c_int32), ("record_traces", C.c_int32)]


class GenerationResult(C.Structure):
    _fields_ = [("k_generation", C.c_double), ("keff", C.c_double), ("keff_std", C.c_double),
                ("k_collision", C.c_double), ("k_absorption", C.c_double), ("k_tracklength", C.c_double),
                ("leakage", C.c_double), ("total_weight", C.c_double),
                ("n_histories", C.c_int64), ("n_segments", C.c_int64), ("n_collisions", C.c_int64),
                ("n_xs_lookups", C.c_int64), ("n_nuclide_lookups", C.c_int64), ("n_bank_sites", C.c_int64),
                ("n_vertex_edge_hits", C.c_int64), ("n_moved_through_neighbors", C.c_int64),
                ("n_failed_traces", C.c_int64), ("n_event_rounds", C.c_int64), ("n_kernel_launches", C.c_int64),
                ("ms_total", C.c_double), ("ms_xs", C.c_double), ("ms_track", C.c_double),
                ("ms_collide", C.c_double), ("ms_bank", C.c_double), ("n_bank_dropped", C.c_int64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


RECORD_DTYPE = np.dtype([("seed_tracking", "u8"), ("seed_source", "u8"), ("seed_urr", "u8"), ("E", "f8"),
                         ("k_collision", "f8"), ("k_absorption", "f8"), ("k_tracklength", "f8"),
                         ("n_segments", "i4"), ("n_collisions", "i4"), ("n_progeny", "i4"), ("last_elem", "i4"),
                         ("fate", "i4"), ("n_xs", "i4"), ("trace_hash", "u8")])
assert RECORD_DTYPE.itemsize == 88

TRACE_DTYPE = np.dtype([("ray_id", "i8"), ("seq", "i4"), ("kind", "i4"), ("point", "f8", 3), ("aux", "f8", 8)])
assert TRACE_DTYPE.itemsize == 104
TRACE_AUX_NAMES = ("energy", "weight", "n_progeny", "id", "seed_tracking", "seed_source", "seed_URR", "particle_type")

SCORE = {"flux": 0, "total": 1, "fission": 2, "nu-fission": 3, "kappa-fission": 4}
ESTIMATOR = {"collision": 0, "tracklength": 1}
GRANULARITY = {"cell": 0, "universe": 1, "global": 2}


def make_settings(n_particles, n_inactive=0, is_2d=False, z_coord=0.0, source_box=(0, 0, 0, 1, 1, 1),
                  only_fissionable=True, seed=1, n_log_bins=8000, seed_through_double=True, max_events=1000000,
                  sort_xs_queue=0, record_traces=1):
    s = Settings()
    s.n_particles, s.master_seed, s.n_inactive, s.n_log_bins = n_particles, seed, n_inactive, n_log_bins
    s.is_2d, s.seed_through_double, s.z_coord = int(is_2d), int(seed_through_double), z_coord
    for i in range(6):
        s.source_box[i] = float(source_box[i])
    s.source_only_fissionable, s.max_events = int(only_fissionable), max_events
    s.sort_xs_queue, s.record_traces = sort_xs_queue, record_traces
    return s
