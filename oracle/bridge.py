"""ctypes bridge to the CPU oracle (oracle/liboracle*.so).  TEST INFRASTRUCTURE: importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from macaw_b200 import meshgen, nucdata
from macaw_b200.cstructs import (RECORD_DTYPE, SITE_DTYPE, GenerationResult, Settings)

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref=True):
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build(ref=False)
    lib = C.CDLL(path)
    u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    lib.orc_prn.restype = C.c_double
    lib.orc_prn.argtypes = [u64p]
    lib.orc_future_seed.restype = C.c_uint64
    lib.orc_future_seed.argtypes = [C.c_uint64, C.c_uint64]
    lib.orc_init_seed.restype = C.c_uint64
    lib.orc_init_seed.argtypes = [C.c_int64, C.c_int, C.c_int64]
    lib.orc_seed_via_double.restype = C.c_uint64
    lib.orc_seed_via_double.argtypes = [C.c_uint64]
    lib.orc_maxwell_spectrum.restype = C.c_double
    lib.orc_maxwell_spectrum.argtypes = [C.c_double, u64p]
    lib.orc_watt_spectrum.restype = C.c_double
    lib.orc_watt_spectrum.argtypes = [C.c_double, C.c_double, u64p]
    lib.orc_rotate_angle.argtypes = [f64p, C.c_double, f64p, u64p]
    lib.orc_isotropic_direction.argtypes = [f64p, u64p]
    lib.orc_sample_cxs_target_velocity.argtypes = [C.c_double, C.c_double, f64p, C.c_double, u64p, f64p]
    lib.orc_tracklength_kappa.argtypes = [C.c_void_p, C.c_int, C.c_int64, f64p, f64p, f64p, f64p]
    lib.orc_set_tally_nuclides.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.orc_set_thermal_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_set_urr_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_set_material_thermal.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_score_nuclide_bins.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, f64p, f64p,
                                           f64p, f64p]
    lib.orc_history_identity.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), u64p]
    lib.orc_collide_events.restype = C.c_int64
    lib.orc_collide_events.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int64, f64p, f64p, u64p, C.c_void_p, f64p, f64p,
                                       u64p, C.c_void_p, C.c_int64, f64p, f64p]
    lib.orc_score_events.restype = C.c_int64
    lib.orc_score_events.argtypes = [C.c_int64, f64p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, f64p, f64p, f64p,
                                     f64p, f64p]
    lib.orc_set_thermal_temperatures.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_add_temperature_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.orc_set_temperature.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.orc_law_samples.argtypes = [C.POINTER(C.c_longlong)]
    lib.orc_law_samples.restype = None
    lib.orc_reaction_samples.argtypes = [C.POINTER(C.c_longlong)]
    lib.orc_reaction_samples.restype = None
    for f in ("orc_log", "orc_sin", "orc_cos", "orc_exp", "orc_sinh"):
        getattr(lib, f).restype = C.c_double
        getattr(lib, f).argtypes = [C.c_double]
    lib.orc_is_within_segment.argtypes = [f64p, f64p, f64p, C.c_double]
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_int, C.c_int]
    lib.orc_destroy.argtypes = [C.c_void_p]
    vp = C.c_void_p
    lib.orc_set_mesh.argtypes = [vp, C.c_int, C.c_int64, vp, C.c_int64, vp, vp, C.c_int, vp, vp, C.c_int, vp]
    lib.orc_set_library.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.orc_set_materials.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    lib.orc_set_settings.argtypes = [vp, C.POINTER(Settings)]
    lib.orc_add_tally.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.orc_get_grid.argtypes = [vp, f64p, f64p, f64p]
    lib.orc_get_grid_index.argtypes = [vp, C.c_int, vp]
    lib.orc_xs_lookup.argtypes = [vp, C.c_int64, vp, vp, vp]
    lib.orc_trace_rays.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.orc_locate_points.argtypes = [vp, C.c_int64, vp, vp, vp]
    lib.orc_sample_initial_source.argtypes = [vp]
    lib.orc_set_source_bank.argtypes = [vp, vp, C.c_int64]
    for f in ("orc_get_source_bank", "orc_get_fission_bank", "orc_get_sampled", "orc_get_history_records", "orc_get_traces"):
        getattr(lib, f).restype = C.c_int64
        getattr(lib, f).argtypes = [vp, vp, C.c_int64]
    lib.orc_transport_generation.argtypes = [vp, C.c_int, C.c_int]
    lib.orc_set_bank_window.argtypes = [vp, C.c_int64, C.c_int64]
    lib.orc_sample_bank.restype = C.c_int64
    lib.orc_sample_bank.argtypes = [vp, C.c_int]
    lib.orc_finish_sampling.restype = C.c_int64
    lib.orc_finish_sampling.argtypes = [vp, C.c_int64, C.c_int]
    lib.orc_finalize_generation.argtypes = [vp, C.c_int, vp, C.POINTER(GenerationResult)]
    lib.orc_run_generation.argtypes = [vp, C.c_int, C.c_int, C.POINTER(GenerationResult)]
    lib.orc_get_tally.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int64]
    lib.orc_tally_size.restype = C.c_int64
    lib.orc_tally_size.argtypes = [vp, C.c_int]
    lib.orc_set_keff.argtypes = [vp, C.c_double]
    lib.orc_set_renormalise_direction.argtypes = [vp, C.c_int]
    lib.orc_get_keff.restype = C.c_double
    lib.orc_get_keff.argtypes = [vp]
    lib.orc_calculate_work.argtypes = [C.c_int64, C.c_int, vp]
    return lib


_libs = {}


def lib(libm=False):
    key = "liboracle_libm.so" if libm else "liboracle.so"
    if key not in _libs:
        _libs[key] = _load(key)
    return _libs[key]


def ref_lib():
    """The reference's own RNG / sampler sources compiled by oracle/Makefile (None if never built)."""
    path = os.path.join(_HERE, "_ref", "libomc_ref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/openmc/src"):
            build(ref=True)
        else:
            return None
    r = C.CDLL(path)
    u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    r.ref_prn.restype = C.c_double
    r.ref_prn.argtypes = [u64p]
    r.ref_future_prn.restype = C.c_double
    r.ref_future_prn.argtypes = [C.c_int64, C.c_uint64]
    r.ref_future_seed.restype = C.c_uint64
    r.ref_future_seed.argtypes = [C.c_uint64, C.c_uint64]
    r.ref_init_seed.restype = C.c_uint64
    r.ref_init_seed.argtypes = [C.c_int64, C.c_int]
    r.ref_init_particle_seeds.argtypes = [C.c_int64, u64p]
    if hasattr(r, "ref_isotropic_direction"):
        r.ref_isotropic_direction.argtypes = [f64p, u64p]
    if hasattr(r, "ref_sort_and_synchronize_bank"):
        r.ref_sort_and_synchronize_bank.restype = C.c_int64
        r.ref_sort_and_synchronize_bank.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                                    C.c_int, C.c_void_p, C.c_void_p]
    if hasattr(r, "ref_generation_bookkeeping"):
        r.ref_generation_bookkeeping.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, f64p, f64p, f64p, f64p, f64p]
    if hasattr(r, "ref_initialize_history"):
        r.ref_initialize_history.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, f64p,
                                             C.POINTER(C.c_int64), u64p, f64p]
    if hasattr(r, "ref_calculate_work"):
        r.ref_calculate_work.restype = C.c_int64
        r.ref_calculate_work.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p]
    if hasattr(r, "ref_collide_events"):
        r.ref_library_reset.argtypes = []
        r.ref_add_nuclide.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_int, f64p, f64p, f64p, f64p, C.c_int, C.c_int,
                                      f64p, C.c_double, C.c_double, f64p, C.c_int, f64p, f64p, C.c_double, C.c_double, C.c_double]
        if hasattr(r, "ref_stage_tables"):
            r.ref_stage_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        r.ref_finish_library.argtypes = [C.c_int]
        r.ref_nuclide_tables.argtypes = [C.c_int, f64p, C.c_void_p, C.c_void_p, C.c_void_p]
        r.ref_add_material.argtypes = [C.c_int, C.c_void_p, f64p, C.c_void_p]
        if hasattr(r, "ref_transport_histories"):
            r.ref_transport_histories.restype = C.c_int64
            r.ref_transport_histories.argtypes = [C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int64, f64p, C.c_double, f64p,
                                                  u64p, C.c_int64, f64p, f64p, C.c_int, f64p]
        r.ref_tracklength_kappa.argtypes = [C.c_int, C.c_int64, f64p, f64p, f64p, C.c_double, f64p]
        if hasattr(r, "ref_stage_law"):
            r.ref_stage_law.argtypes = [C.c_void_p]
        if hasattr(r, "ref_stage_urr"):
            r.ref_stage_urr.argtypes = [C.c_void_p]
        if hasattr(r, "ref_gather_digest"):
            r.ref_gather_digest.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_char_p, C.c_int, C.POINTER(C.c_uint64),
                                            C.c_void_p, C.c_void_p, C.c_int]
        if hasattr(r, "ref_dump_tree"):
            r.ref_dump_tree.restype = C.c_int64
            r.ref_dump_tree.argtypes = [C.c_void_p, C.c_int64]
        if hasattr(r, "ref_stage_temperature_set"):
            r.ref_stage_temperature_set.argtypes = [C.c_void_p, C.c_void_p]
            r.ref_nuclide_temperatures.argtypes = [C.c_int, f64p]
            r.ref_nuclide_tables_t.argtypes = [C.c_int, C.c_int, f64p, C.c_void_p]
        if hasattr(r, "ref_add_thermal_table"):
            r.ref_add_thermal_table.argtypes = [C.c_char_p, C.c_void_p, C.c_double]
            if hasattr(r, "ref_stage_thermal_temperature"):
                r.ref_stage_thermal_temperature.argtypes = [C.c_void_p, C.c_double]
            r.ref_set_material_thermal.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, f64p]
        if hasattr(r, "ref_score_nuclide_bins"):
            r.ref_score_nuclide_bins.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int64,
                                                 f64p, f64p, f64p, C.c_double, f64p]
        r.ref_collide_events.restype = C.c_int64
        r.ref_collide_events.argtypes = [C.c_int, C.c_double, C.c_int64, f64p, f64p, u64p, C.c_void_p, C.c_double, f64p, f64p,
                                         u64p, C.c_void_p, C.c_int64, f64p, f64p]
    if hasattr(r, "ref_score_collision_events"):
        r.ref_score_collision_events.argtypes = [C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                                 C.c_int64, C.c_void_p, C.c_void_p, f64p, f64p, f64p, f64p]
        r.ref_score_tracklength_events.argtypes = [C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                                   C.c_int64, C.c_void_p, C.c_void_p, f64p, f64p, f64p, f64p, f64p]
    if hasattr(r, "ref_tally_accumulate"):
        r.ref_tally_accumulate.argtypes = [f64p, C.c_int, C.c_int, C.c_int64, f64p, f64p]
    if hasattr(r, "ref_keff_statistics"):
        r.ref_keff_statistics.argtypes = [f64p, C.c_int, C.c_int, f64p, f64p]
    if hasattr(r, "ref_sample_cxs_target_velocity"):
        r.ref_sample_cxs_target_velocity.argtypes = [C.c_double, C.c_double, f64p, C.c_double, u64p, f64p]
    if hasattr(r, "ref_lower_bound_index"):
        r.ref_lower_bound_index.restype = C.c_int64
        r.ref_lower_bound_index.argtypes = [f64p, C.c_int64, C.c_double]
    if hasattr(r, "ref_box_sample"):
        r.ref_box_sample.argtypes = [f64p, u64p, f64p]
    r.ref_set_master_seed.argtypes = [C.c_int64]
    r.ref_maxwell_spectrum.restype = C.c_double
    r.ref_maxwell_spectrum.argtypes = [C.c_double, u64p]
    r.ref_watt_spectrum.restype = C.c_double
    r.ref_watt_spectrum.argtypes = [C.c_double, C.c_double, u64p]
    r.ref_uniform.restype = C.c_double
    r.ref_uniform.argtypes = [C.c_double, C.c_double, u64p]
    r.ref_rotate_angle.argtypes = [f64p, C.c_double, f64p, u64p]
    return r


def _p(a):
    return a.ctypes.data if a is not None else None


def ref_nuclide_file_image(r) -> bytes:
    """The in-memory nuclide-file image the reference's constructors read in the last ref_load_library, as bytes
    (macaw_b200.openmc_hdf5.read_image parses them)."""
    n = r.ref_dump_tree(None, 0)
    buf = (C.c_ubyte * n)()
    assert r.ref_dump_tree(C.addressof(buf), n) == n
    return bytes(buf)


def ref_load_library(r, libr: nucdata.Library):
    """Hands a synthetic library to the REFERENCE's own constructors (oracle/ref_shim_physics.cpp: Nuclide, Reaction,
    ReactionProduct, distributions, Material, init_grid) and returns, per nuclide, the reference's derived tables
    (xs_[0] as [G][5], fissionable_, n_precursor_, grid_index).  Capture is handed over as absorption - fission."""
    f64p = C.POINTER(C.c_double)

    def dp(a):
        return a.ctypes.data_as(f64p)
    r.ref_library_reset()
    stash = []
    for n in libr.nuclides:
        e, el = np.ascontiguousarray(n.energy, np.float64), np.ascontiguousarray(n.elastic, np.float64)
        fis = np.ascontiguousarray(n.xs[:, 2])
        cap = np.ascontiguousarray(n.xs[:, 1] - n.xs[:, 2])
        inel = np.ascontiguousarray(n.inelastic if n.inelastic.size else np.zeros(1), np.float64)
        nu, dy, dth = np.array(n.nu_poly, np.float64), np.ascontiguousarray(n.delayed_yield), np.ascontiguousarray(n.delayed_theta)
        if any(getattr(n, a, None) is not None for a in ("elastic_angle", "inelastic_angle", "inelastic_energy")):
            from macaw_b200 import nucdata as _nd
            descs, keep = [], []
            for a in ("elastic_angle", "inelastic_angle", "inelastic_energy"):
                d = _nd.Tabular2DDesc()
                _nd._pack_table(d, getattr(n, a, None), keep)
                descs.append(d)
            r.ref_stage_tables(C.addressof(descs[0]), C.addressof(descs[1]), C.addressof(descs[2]), int(n.inelastic_in_cm))
        if getattr(n, "inelastic_law", -1) >= 2 or getattr(n, "more_reactions", None) or getattr(n, "inelastic_yield", 1.0) != 1.0 \
                or getattr(n, "fission_products", None):
            from macaw_b200 import nucdata as _nd
            one = _nd.Library([n], [])
            ndesc, _, keep_l = _nd.pack_library(one)
            r.ref_stage_law(C.addressof(ndesc))
        for m in (getattr(n, "more_temperatures", None) or []):  # further temperatures of this nuclide, with their URR tables
            from macaw_b200 import nucdata as _nd
            sd, _, keep_s = _nd.pack_library(_nd.Library([m], []))
            entry = len(libr.nuclides) + [id(x) for _, x in libr.temperature_sets()].index(id(m))
            su = [t for t in (getattr(libr, "urr_tables", None) or []) if t.nuclide == entry]
            ud_s, keep_us = _nd.pack_urr(su) if su else (None, None)
            stash.append((sd, keep_s, ud_s, keep_us))
            r.ref_stage_temperature_set(C.addressof(sd), C.addressof(ud_s) if su else None)
        urr = [t for t in (getattr(libr, "urr_tables", None) or []) if t.nuclide < len(libr.nuclides) and libr.nuclides[t.nuclide] is n]
        if urr:
            from macaw_b200 import nucdata as _nd
            ud, keep_u = _nd.pack_urr(urr)
            r.ref_stage_urr(C.addressof(ud))
        r.ref_add_nuclide(n.name.encode(), n.awr, n.kT, e.size, dp(e), dp(el), dp(cap), dp(fis) if n.fissionable else None,
                          int(n.inelastic.size), n.inel_threshold, dp(inel), n.inel_threshold_energy, n.inel_mass_ratio,
                          dp(nu), n.n_precursor, dp(dy), dp(dth), n.watt_a, n.watt_b, n.q_fission)
    r.ref_finish_library(libr.n_log_bins)
    out = []
    for i, n in enumerate(libr.nuclides):
        xs5, gi = np.zeros((n.energy.size, 5)), np.zeros(libr.n_log_bins + 1, np.int32)
        fz, npre = C.c_int32(), C.c_int32()
        r.ref_nuclide_tables(i, dp(xs5), C.byref(fz), C.byref(npre), gi.ctypes.data)
        out.append((xs5, bool(fz.value), npre.value, gi))
        # further temperatures: the reference's derived tables xs_[t] and grid_[t].grid_index, by increasing kT; the set
        # at position t of the sorted temperatures is the library's own entry or one of more_temperatures
        more = getattr(n, "more_temperatures", None) or []
        if more:
            kts = np.zeros(len(more) + 1)
            assert r.ref_nuclide_temperatures(i, dp(kts)) == len(more) + 1
            by_kt = sorted([n] + list(more), key=lambda x: x.kT)
            assert np.array_equal(kts, [x.kT for x in by_kt])
            per_set = {}
            for t, x in enumerate(by_kt):
                xs_t, gi_t = np.zeros((x.energy.size, 5)), np.zeros(libr.n_log_bins + 1, np.int32)
                r.ref_nuclide_tables_t(i, t, dp(xs_t), gi_t.ctypes.data)
                per_set[id(x)] = (xs_t, gi_t)
            out[-1] = out[-1] + (per_set,)
    for m in libr.materials:
        nn, dd = np.ascontiguousarray(m.nuclides, np.int32), np.ascontiguousarray(m.densities, np.float64)
        iso = np.ascontiguousarray(m.isotropic, np.int32) if m.isotropic is not None else None
        r.ref_add_material(nn.size, nn.ctypes.data, dp(dd), iso.ctypes.data if iso is not None else None)
    if getattr(libr, "thermal_tables", None):
        from macaw_b200 import nucdata as _nd
        td, keep = _nd.pack_thermal(libr.thermal_tables)
        for k, t in enumerate(libr.thermal_tables):
            more = getattr(t, "more_temperatures", None) or []
            if more:  # further temperatures of the table: one `<T>K` group each
                md, keep_m = _nd.pack_thermal(more)
                stash.append((md, keep_m))
                for j, m in enumerate(more):
                    r.ref_stage_thermal_temperature(C.addressof(md[j]), m.kT)
            r.ref_add_thermal_table(t.name.encode(), C.addressof(td[k]), getattr(t, "kT", libr.nuclides[0].kT))
        for im, m in enumerate(libr.materials):
            if getattr(m, "sab", None):
                pos = np.ascontiguousarray([a[0] for a in m.sab], np.int32)
                tab = np.ascontiguousarray([a[1] for a in m.sab], np.int32)
                fr = np.ascontiguousarray([a[2] for a in m.sab], np.float64)
                r.ref_set_material_thermal(im, pos.size, pos.ctypes.data, tab.ctypes.data, dp(fr))
    return out


class Oracle:
    """One oracle problem instance (mirrors the product's macaw_b200.Context call for call)."""

    def __init__(self, libm=False, rank=0, nranks=1):
        self.L = lib(libm)
        self.h = self.L.orc_create(rank, nranks)
        self._keep = []
        self.mesh = None

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_mesh(self, m: meshgen.Mesh, block_ids=None):
        self.mesh = m
        ids = m.block_ids_sorted() if block_ids is None else np.asarray(block_ids, np.int32)
        self.block_ids = ids
        bi = np.ascontiguousarray(m.block_index(ids), np.int32)
        xyz = np.ascontiguousarray(m.xyz, np.float64)
        conn = np.ascontiguousarray(m.conn, np.int64)
        nbr = np.ascontiguousarray(m.neighbors, np.int64)
        sb = np.ascontiguousarray(m.side_bnd_id, np.int32)
        bk = np.ascontiguousarray(m.bnd_kind, np.int32)
        self.L.orc_set_mesh(self.h, m.elem_type, xyz.shape[0], _p(xyz), conn.shape[0], _p(conn), _p(bi), len(ids),
                            _p(nbr), _p(sb), bk.size, _p(bk))

    def set_library(self, libr: nucdata.Library, block_material):
        nd, md, keep = nucdata.pack_library(libr)
        self._keep.append((nd, md, keep))
        self.L.orc_set_library(self.h, C.addressof(nd), len(libr.nuclides), libr.n_log_bins)
        sets, keep_s = nucdata.pack_temperature_sets(libr)  # further temperatures: entries len(nuclides), ...
        self._keep.append((sets, keep_s))
        for k, (i, d) in enumerate(sets):
            entry = C.c_int(-1)
            if self.L.orc_add_temperature_set(self.h, i, C.addressof(d), C.byref(entry)) or entry.value != len(libr.nuclides) + k:
                raise RuntimeError("orc_add_temperature_set failed")
        bm = np.ascontiguousarray(block_material, np.int32)
        self.L.orc_set_materials(self.h, C.addressof(md), len(libr.materials), _p(bm), bm.size)
        if getattr(libr, "urr_tables", None):
            ud, keep3 = nucdata.pack_urr(libr.urr_tables)
            self._keep.append((ud, keep3))
            if self.L.orc_set_urr_tables(self.h, C.addressof(ud), len(libr.urr_tables)):
                raise RuntimeError("orc_set_urr_tables failed")
        if getattr(libr, "thermal_tables", None):
            tables, same, kT = libr.thermal_entries()  # every temperature of every table as one entry
            td, keep2 = nucdata.pack_thermal(tables)
            self._keep.append((td, keep2))
            self.L.orc_set_thermal_tables(self.h, C.addressof(td), len(tables))
            if len(tables) > len(libr.thermal_tables) and self.L.orc_set_thermal_temperatures(self.h, _p(same), _p(kT), len(tables)):
                raise RuntimeError("orc_set_thermal_temperatures failed")
            for im, m in enumerate(libr.materials):
                if getattr(m, "sab", None):
                    pos = np.ascontiguousarray([a[0] for a in m.sab], np.int32)
                    tab = np.ascontiguousarray([a[1] for a in m.sab], np.int32)
                    fr = np.ascontiguousarray([a[2] for a in m.sab], np.float64)
                    if self.L.orc_set_material_thermal(self.h, im, pos.size, _p(pos), _p(tab), _p(fr)):
                        raise RuntimeError("orc_set_material_thermal failed")

    def set_temperature(self, T):
        T = np.ascontiguousarray(np.atleast_1d(T), np.float64)
        self.L.orc_set_temperature(self.h, _p(T), T.size)

    def set_settings(self, s: Settings):
        self.settings = s
        self.L.orc_set_settings(self.h, C.byref(s))

    def add_tally(self, tid, estimator, granularity, scores, energy_bins=None):
        sc = np.ascontiguousarray(scores, np.int32)
        eb = np.ascontiguousarray(energy_bins if energy_bins is not None else [], np.float64)
        self.L.orc_add_tally(self.h, tid, estimator, granularity, _p(sc), sc.size, _p(eb), eb.size)

    def set_tally_nuclides(self, tid, nuclides):
        nn = np.ascontiguousarray(nuclides, np.int32)
        if self.L.orc_set_tally_nuclides(self.h, tid, _p(nn), nn.size):
            raise RuntimeError("unknown tally id")

    def grid(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_get_grid(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def grid_index(self, i_nuc, n_log_bins):
        out = np.zeros(n_log_bins + 1, np.int32)
        self.L.orc_get_grid_index(self.h, i_nuc, _p(out))
        return out

    def xs_lookup(self, E, mat):
        E = np.ascontiguousarray(E, np.float64)
        mat = np.ascontiguousarray(mat, np.int32)
        out = np.zeros((E.size, 4))
        self.L.orc_xs_lookup(self.h, E.size, _p(E), _p(mat), _p(out))
        return out

    def trace_rays(self, start, dirs, start_elem=None, start_side=None, max_distance=None):
        start = np.ascontiguousarray(start, np.float64)
        dirs = np.ascontiguousarray(dirs, np.float64)
        n = start.shape[0]
        se = np.ascontiguousarray(start_elem if start_elem is not None else np.full(n, -1), np.int32)
        ss = np.ascontiguousarray(start_side, np.int32) if start_side is not None else None
        dist, seg, h, st = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.uint64), np.zeros(6, np.int64)
        md = np.ascontiguousarray(max_distance, np.float64) if max_distance is not None else None
        self.L.orc_trace_rays(self.h, n, _p(start), _p(dirs), _p(se), _p(ss), _p(md), _p(dist), _p(seg), _p(h),
                              _p(st))
        return dist, seg, h, dict(zip(["segments", "vertex_edge_hits", "moved_through_neighbors", "failed",
                                       "intersection_calls", "reflections"], st.tolist()))

    def locate_points(self, xyz, ray_id=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        rid = np.ascontiguousarray(ray_id, np.int64) if ray_id is not None else None
        out = np.zeros(xyz.shape[0], np.int32)
        self.L.orc_locate_points(self.h, xyz.shape[0], _p(xyz), _p(rid), _p(out))
        return out

    def sample_initial_source(self):
        self.L.orc_sample_initial_source(self.h)

    def set_source_bank(self, sites):
        sites = np.ascontiguousarray(sites, SITE_DTYPE)
        self.L.orc_set_source_bank(self.h, _p(sites), sites.size)

    def _get(self, fn, dtype):
        n = fn(self.h, None, 0)
        out = np.zeros(n, dtype)
        fn(self.h, _p(out), n)
        return out

    def source_bank(self):
        return self._get(self.L.orc_get_source_bank, SITE_DTYPE)

    def fission_bank(self):
        return self._get(self.L.orc_get_fission_bank, SITE_DTYPE)

    def sampled(self):
        return self._get(self.L.orc_get_sampled, SITE_DTYPE)

    def history_records(self):
        return self._get(self.L.orc_get_history_records, RECORD_DTYPE)

    def traces(self):
        from macaw_b200.cstructs import TRACE_DTYPE
        return self._get(self.L.orc_get_traces, TRACE_DTYPE)

    def run_generation(self, batch, n_threads=1):
        res = GenerationResult()
        rc = self.L.orc_run_generation(self.h, batch, n_threads, C.byref(res))
        if rc:
            raise RuntimeError(f"oracle run_generation failed rc={rc}")
        return res

    def transport_generation(self, batch, n_threads=1):
        self.L.orc_transport_generation(self.h, batch, n_threads)

    def get_tally(self, tid):
        n = self.L.orc_tally_size(self.h, tid)
        v, s, q = np.zeros(n), np.zeros(n), np.zeros(n)
        nr = self.L.orc_get_tally(self.h, tid, _p(v), _p(s), _p(q), n)
        return v, s, q, nr
