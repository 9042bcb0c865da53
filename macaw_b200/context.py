"""Thin object wrapper over the C ABI: one Context = one mcw_handle = one GPU / rank."""
import ctypes as C

import numpy as np

from . import meshgen, nucdata
from .cstructs import RECORD_DTYPE, SITE_DTYPE, TRACE_DTYPE, GenerationResult, Settings
from .lib import McwError, load


def _p(a):
    return a.ctypes.data if a is not None else None


class Context:
    def __init__(self, device=0, rank=0, nranks=1):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.mcw_create(C.byref(self.h), device, rank, nranks)
        if rc:
            raise McwError(rc, "mcw_create failed: a CUDA device of compute capability 10.x (B200) is required; "
                               "there is no CPU fallback")
        self.rank, self.nranks = rank, nranks
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.mcw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise McwError(rc, self.L.mcw_last_error(self.h).decode())

    def set_option(self, name, value):
        self._ck(self.L.mcw_set_option(self.h, name.encode(), int(value)))

    def prepare(self):
        self._ck(self.L.mcw_prepare(self.h))

    def stream(self):
        p = C.c_void_p()
        self._ck(self.L.mcw_stream(self.h, C.byref(p)))
        return p.value

    def set_mesh(self, m: meshgen.Mesh, pass_neighbors=True):
        self.mesh = m
        xyz = np.ascontiguousarray(m.xyz, np.float64)
        conn = np.ascontiguousarray(m.conn, np.int64)
        blk = np.ascontiguousarray(m.block_id, np.int32)
        nbr = np.ascontiguousarray(m.neighbors, np.int64) if (pass_neighbors and m.neighbors is not None) else None
        sb = np.ascontiguousarray(m.side_bnd_id, np.int32)
        bk = np.ascontiguousarray(m.bnd_kind, np.int32)
        self._ck(self.L.mcw_set_mesh(self.h, m.elem_type, xyz.shape[0], _p(xyz), conn.shape[0], _p(conn), _p(blk),
                                     _p(nbr), _p(sb), bk.size, _p(bk)))

    def set_library(self, libr: nucdata.Library, block_material, block_ids=None):
        nd, md, keep = nucdata.pack_library(libr)
        self._keep.append((nd, md, keep))
        self._ck(self.L.mcw_set_library(self.h, C.addressof(nd), len(libr.nuclides)))
        sets, keep_s = nucdata.pack_temperature_sets(libr)  # further temperatures of the nuclides: entries len(nuclides), ...
        self._keep.append((sets, keep_s))
        for k, (i, d) in enumerate(sets):
            entry = C.c_int(-1)
            self._ck(self.L.mcw_add_temperature_set(self.h, i, C.addressof(d), C.byref(entry)))
            assert entry.value == len(libr.nuclides) + k
        ids = np.ascontiguousarray(self.mesh.block_ids_sorted() if block_ids is None else block_ids, np.int32)
        bm = np.ascontiguousarray(block_material, np.int32)
        assert ids.size == bm.size
        self._ck(self.L.mcw_set_materials(self.h, C.addressof(md), len(libr.materials), _p(ids), _p(bm), bm.size))
        if getattr(libr, "urr_tables", None):  # unresolved-resonance probability tables
            ud, keep3 = nucdata.pack_urr(libr.urr_tables)
            self._keep.append((ud, keep3))
            self._ck(self.L.mcw_set_urr_tables(self.h, C.addressof(ud), len(libr.urr_tables)))
        if getattr(libr, "thermal_tables", None):  # S(a,b): data::thermal_scatt + the materials' <sab> entries
            tables, same, kT = libr.thermal_entries()  # every temperature of every table as one entry
            td, keep2 = nucdata.pack_thermal(tables)
            self._keep.append((td, keep2))
            self._ck(self.L.mcw_set_thermal_tables(self.h, C.addressof(td), len(tables)))
            if len(tables) > len(libr.thermal_tables):
                self._ck(self.L.mcw_set_thermal_temperatures(self.h, _p(same), _p(kT), len(tables)))
            for im, m in enumerate(libr.materials):
                if getattr(m, "sab", None):
                    pos = np.ascontiguousarray([a[0] for a in m.sab], np.int32)
                    tab = np.ascontiguousarray([a[1] for a in m.sab], np.int32)
                    fr = np.ascontiguousarray([a[2] for a in m.sab], np.float64)
                    self._ck(self.L.mcw_set_material_thermal(self.h, im, pos.size, _p(pos), _p(tab), _p(fr)))

    def set_temperature(self, T):
        T = np.ascontiguousarray(np.atleast_1d(T), np.float64)
        self._ck(self.L.mcw_set_temperature(self.h, _p(T), T.size))

    def set_settings(self, s: Settings):
        self.settings = s
        self._ck(self.L.mcw_set_settings(self.h, C.byref(s)))

    def add_tally(self, tid, estimator, granularity, scores, energy_bins=None):
        sc = np.ascontiguousarray(scores, np.int32)
        eb = np.ascontiguousarray(energy_bins if energy_bins is not None else [], np.float64)
        self._ck(self.L.mcw_add_tally(self.h, tid, estimator, granularity, _p(sc), sc.size, _p(eb), eb.size))

    def set_tally_nuclides(self, tid, nuclides):
        """nuclide bins of a tally: library indices, -1 = total (OpenMCTally.C:188-204)"""
        nn = np.ascontiguousarray(nuclides, np.int32)
        self._ck(self.L.mcw_set_tally_nuclides(self.h, tid, _p(nn), nn.size))

    # ---- source / banks
    def sample_initial_source(self):
        self._ck(self.L.mcw_sample_initial_source(self.h))

    def set_source_bank(self, sites):
        sites = np.ascontiguousarray(sites, SITE_DTYPE)
        self._ck(self.L.mcw_set_source_bank(self.h, _p(sites), sites.size))

    def source_bank(self, out=None):
        n = C.c_int64()
        self._ck(self.L.mcw_get_source_bank(self.h, None, 0, C.byref(n)))
        if out is None:
            out = np.zeros(n.value, SITE_DTYPE)
        self._ck(self.L.mcw_get_source_bank(self.h, _p(out), out.size, C.byref(n)))
        return out[:n.value]

    def fission_bank(self):
        n = C.c_int64()
        self._ck(self.L.mcw_get_fission_bank(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, SITE_DTYPE)
        self._ck(self.L.mcw_get_fission_bank(self.h, _p(out), out.size, C.byref(n)))
        return out

    def history_records(self):
        n = C.c_int64()
        self._ck(self.L.mcw_get_history_records(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, RECORD_DTYPE)
        self._ck(self.L.mcw_get_history_records(self.h, _p(out), out.size, C.byref(n)))
        return out

    def traces(self):
        """Cached trace points of the last generation (settings.record_traces = 2), ordered by (ray_id, seq)."""
        n = C.c_int64()
        self._ck(self.L.mcw_get_traces(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, TRACE_DTYPE)
        self._ck(self.L.mcw_get_traces(self.h, _p(out), out.size, C.byref(n)))
        return out

    # ---- generations
    def run_generation(self, batch):
        res = GenerationResult()
        self._ck(self.L.mcw_run_generation(self.h, batch, C.byref(res)))
        return res

    def transport_generation(self, batch):
        self._ck(self.L.mcw_transport_generation(self.h, batch))

    def finalize_generation(self, batch):
        res = GenerationResult()
        self._ck(self.L.mcw_finalize_generation(self.h, batch, C.byref(res)))
        return res

    def bank_count(self):
        n = C.c_int64()
        self._ck(self.L.mcw_bank_count(self.h, C.byref(n)))
        return n.value

    def set_bank_window(self, start, total):
        self._ck(self.L.mcw_set_bank_window(self.h, start, total))

    def sample_bank(self, batch):
        n = C.c_int64()
        self._ck(self.L.mcw_sample_bank(self.h, batch, C.byref(n)))
        return n.value

    def finish_sampling(self, start, is_last):
        n = C.c_int64()
        self._ck(self.L.mcw_finish_sampling(self.h, start, int(is_last), C.byref(n)))
        return n.value

    def _dev(self, fn):
        p, n = C.c_void_p(), C.c_int64()
        self._ck(fn(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def sampled_device(self):
        return self._dev(self.L.mcw_sampled_device_ptr)

    def source_device(self):
        return self._dev(self.L.mcw_source_device_ptr)

    def bank_device(self):
        return self._dev(self.L.mcw_bank_device_ptr)

    def reduction_device(self):
        return self._dev(self.L.mcw_reduction_device_ptr)

    def set_source_count(self, n):
        self._ck(self.L.mcw_set_source_count(self.h, n))

    # ---- tallies
    def tally_shape(self, tid):
        a, b, c = C.c_int64(), C.c_int32(), C.c_int32()
        self._ck(self.L.mcw_tally_shape(self.h, tid, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def tally_nuclide_bins(self, tid):
        a = C.c_int32()
        self._ck(self.L.mcw_tally_nuclide_bins(self.h, tid, C.byref(a)))
        return a.value

    def get_tally(self, tid):
        """SUM, SUM_SQ as [spatial][energy][score_index], score_index = nuclide_bin * n_scores + score (OpenMC's results_)"""
        ns, ne, nc = self.tally_shape(tid)
        nc *= self.tally_nuclide_bins(tid)
        n = ns * ne * nc
        s, q, nr = np.zeros(n), np.zeros(n), C.c_int32()
        self._ck(self.L.mcw_get_tally(self.h, tid, _p(s), _p(q), n, C.byref(nr)))
        return s.reshape(ns, ne, nc), q.reshape(ns, ne, nc), nr.value

    # ---- stand-alone kernels
    def xs_lookup(self, E, mat):
        E = np.ascontiguousarray(E, np.float64)
        mat = np.ascontiguousarray(mat, np.int32)
        out = np.zeros((E.size, 4))
        self._ck(self.L.mcw_xs_lookup(self.h, E.size, _p(E), _p(mat), _p(out)))
        return out

    def xs_lookup_device(self, n, d_E, d_mat, d_out, repeats=1, sorted=0):
        ms = C.c_float()
        self._ck(self.L.mcw_xs_lookup_device(self.h, n, d_E, d_mat, d_out, sorted, repeats, C.byref(ms)))
        return ms.value

    def test_division(self, a, b):
        a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
        qf, qp, ok = np.zeros(a.size), np.zeros(a.size), np.zeros(a.size, np.int32)
        self._ck(self.L.mcw_test_division(self.h, a.size, _p(a), _p(b), _p(qf), _p(qp), _p(ok)))
        return qf, qp, ok

    def grid_index(self, i_nuc, n_log_bins):
        out = np.zeros(n_log_bins + 1, np.int32)
        self._ck(self.L.mcw_get_grid_index(self.h, i_nuc, _p(out), out.size))
        return out

    def trace_rays(self, start, dirs, start_elem=None, start_side=None, max_distance=None):
        start = np.ascontiguousarray(start, np.float64)
        dirs = np.ascontiguousarray(dirs, np.float64)
        n = start.shape[0]
        se = np.ascontiguousarray(start_elem if start_elem is not None else np.full(n, -1), np.int32)
        ss = np.ascontiguousarray(start_side, np.int32) if start_side is not None else None
        md = np.ascontiguousarray(max_distance, np.float64) if max_distance is not None else None
        dist, seg, h = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.uint64)
        self._ck(self.L.mcw_trace_rays(self.h, n, _p(start), _p(dirs), _p(se), _p(ss), _p(md), _p(dist), _p(seg),
                                       _p(h)))
        return dist, seg, h

    def locate_points(self, xyz, ray_id=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        rid = np.ascontiguousarray(ray_id, np.int64) if ray_id is not None else None
        out = np.zeros(xyz.shape[0], np.int32)
        self._ck(self.L.mcw_locate_points(self.h, xyz.shape[0], _p(xyz), _p(rid), _p(out)))
        return out
