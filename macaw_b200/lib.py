"""ctypes binding of libmacaw_b200.so (include/macaw_b200.h).  There is no fallback: if the CUDA
library is missing it is built with nvcc; if that fails, or no sm_100 GPU is visible when a context
is created, the call raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .cstructs import GenerationResult, Settings

_LIB = None
_PATH = None


class McwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmacaw_b200 error {code}: {msg}")
        self.code = code


def load():
    global _LIB, _PATH
    if _LIB is not None:
        return _LIB
    path = os.environ.get("MCW_LIB_PATH") or _build.LIB  # MCW_LIB_PATH: an A/B build of the same CUDA library
    if not os.path.exists(path):
        _build.build()  # raises if nvcc is missing: no silent fallback
    L = C.CDLL(path)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    P = C.POINTER
    sig = {
        "mcw_create": [P(vp), C.c_int, C.c_int, C.c_int],
        "mcw_set_option": [vp, C.c_char_p, C.c_int],
        "mcw_prepare": [vp],
        "mcw_stream": [vp, P(vp)],
        "mcw_struct_sizes": [vp, C.c_int],
        "mcw_set_mesh": [vp, C.c_int, i64, vp, i64, vp, vp, vp, vp, C.c_int, vp],
        "mcw_set_library": [vp, vp, C.c_int],
        "mcw_validate_library": [vp, C.c_int, C.c_char_p, C.c_int],
        "mcw_digest": [C.c_int, vp, C.c_int, P(C.c_uint64)],
        "mcw_add_temperature_set": [vp, C.c_int, vp, P(C.c_int)],
        "mcw_set_materials": [vp, vp, C.c_int, vp, vp, C.c_int],
        "mcw_set_thermal_tables": [vp, vp, C.c_int],
        "mcw_set_urr_tables": [vp, vp, C.c_int],
        "mcw_set_thermal_temperatures": [vp, vp, vp, C.c_int],
        "mcw_set_material_thermal": [vp, C.c_int, C.c_int, vp, vp, vp],
        "mcw_set_temperature": [vp, vp, i64],
        "mcw_set_settings": [vp, P(Settings)],
        "mcw_add_tally": [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int],
        "mcw_set_tally_nuclides": [vp, C.c_int, vp, C.c_int],
        "mcw_tally_nuclide_bins": [vp, C.c_int, P(i32)],
        "mcw_sample_initial_source": [vp],
        "mcw_set_source_bank": [vp, vp, i64],
        "mcw_get_source_bank": [vp, vp, i64, P(i64)],
        "mcw_transport_generation": [vp, C.c_int],
        "mcw_finalize_generation": [vp, C.c_int, P(GenerationResult)],
        "mcw_run_generation": [vp, C.c_int, P(GenerationResult)],
        "mcw_bank_count": [vp, P(i64)],
        "mcw_bank_device_ptr": [vp, P(vp), P(i64)],
        "mcw_set_bank_window": [vp, i64, i64],
        "mcw_sample_bank": [vp, C.c_int, P(i64)],
        "mcw_finish_sampling": [vp, i64, C.c_int, P(i64)],
        "mcw_sampled_device_ptr": [vp, P(vp), P(i64)],
        "mcw_source_device_ptr": [vp, P(vp), P(i64)],
        "mcw_set_source_count": [vp, i64],
        "mcw_reduction_device_ptr": [vp, P(vp), P(i64)],
        "mcw_get_tally": [vp, C.c_int, vp, vp, i64, P(i32)],
        "mcw_tally_shape": [vp, C.c_int, P(i64), P(i32), P(i32)],
        "mcw_get_history_records": [vp, vp, i64, P(i64)],
        "mcw_get_fission_bank": [vp, vp, i64, P(i64)],
        "mcw_get_traces": [vp, vp, i64, P(i64)],
        "mcw_get_grid_index": [vp, C.c_int, vp, C.c_int],
        "mcw_xs_lookup": [vp, i64, vp, vp, vp],
        "mcw_xs_lookup_device": [vp, i64, vp, vp, vp, C.c_int, C.c_int, P(C.c_float)],
        "mcw_trace_rays": [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp],
        "mcw_locate_points": [vp, i64, vp, vp, vp],
        "mcw_calculate_work": [i64, C.c_int, vp],
        "mcw_test_division": [vp, i64, vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.mcw_destroy.argtypes = [vp]
    L.mcw_destroy.restype = None
    L.mcw_last_error.argtypes = [vp]
    L.mcw_last_error.restype = C.c_char_p
    L.mcw_version.restype = C.c_int
    _LIB, _PATH = L, os.path.relpath(os.path.abspath(path), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    return L


def loaded_path():
    """Path (relative to the repo root) of the CUDA library this process bound: bench.py prints it so an
    MCW_LIB_PATH override can never hide behind a bench line."""
    load()
    return _PATH


EXPORTS = ["mcw_create", "mcw_destroy", "mcw_last_error", "mcw_version", "mcw_set_option", "mcw_prepare", "mcw_stream",
           "mcw_struct_sizes", "mcw_set_mesh", "mcw_set_library", "mcw_validate_library", "mcw_digest", "mcw_add_temperature_set", "mcw_set_materials", "mcw_set_thermal_tables", "mcw_set_material_thermal", "mcw_set_urr_tables", "mcw_set_thermal_temperatures", "mcw_set_temperature",
           "mcw_set_settings", "mcw_add_tally", "mcw_set_tally_nuclides", "mcw_tally_nuclide_bins",
           "mcw_sample_initial_source", "mcw_set_source_bank",
           "mcw_get_source_bank", "mcw_transport_generation", "mcw_finalize_generation", "mcw_run_generation",
           "mcw_bank_count", "mcw_bank_device_ptr", "mcw_set_bank_window", "mcw_sample_bank", "mcw_finish_sampling",
           "mcw_sampled_device_ptr", "mcw_source_device_ptr", "mcw_set_source_count", "mcw_reduction_device_ptr",
           "mcw_get_tally", "mcw_tally_shape", "mcw_get_history_records", "mcw_get_fission_bank", "mcw_get_traces",
           "mcw_get_grid_index", "mcw_xs_lookup", "mcw_xs_lookup_device", "mcw_trace_rays", "mcw_locate_points",
           "mcw_calculate_work", "mcw_test_division"]


def calculate_work(n_particles, nranks):
    """OpenMC calculate_work (openmc/src/simulation.cpp:529-551); needs no GPU."""
    out = np.zeros(nranks + 1, np.int64)
    rc = load().mcw_calculate_work(n_particles, nranks, out.ctypes.data)
    if rc:
        raise McwError(rc, "mcw_calculate_work")
    return out
