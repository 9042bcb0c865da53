"""Problem files: one self-contained binary with mesh, nuclear data, block->material map, settings and
tallies, so the C++ host (host/mcw_run) runs byte-identical inputs to the Python host and the oracle.

Layout (little endian): 8-byte magic ``MCWPROB1``, uint32 record count, then records of
``char name[32]; uint32 dtype (0 f64, 1 i32, 2 i64); uint32 reserved; uint64 count; data`` with the
data padded to a multiple of 8 bytes.  host/problem_io.hpp is the reader."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"MCWPROB1"
_DT = {0: np.float64, 1: np.int32, 2: np.int64}
_CODE = {np.dtype(np.float64): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2}


def _records(mesh, lib, blocks, materials, settings, tallies, temperature):
    r = {}
    r["mesh.elem_type"] = np.array([mesh.elem_type], np.int32)
    r["mesh.xyz"] = np.ascontiguousarray(mesh.xyz, np.float64).ravel()
    r["mesh.conn"] = np.ascontiguousarray(mesh.conn, np.int64).ravel()
    r["mesh.block_id"] = np.ascontiguousarray(mesh.block_id, np.int32)
    if mesh.neighbors is not None:
        r["mesh.neighbors"] = np.ascontiguousarray(mesh.neighbors, np.int64).ravel()
    r["mesh.side_bnd_id"] = np.ascontiguousarray(mesh.side_bnd_id, np.int32).ravel()
    r["mesh.bnd_kind"] = np.ascontiguousarray(mesh.bnd_kind, np.int32)
    r["lib.counts"] = np.array([len(lib.nuclides), len(lib.materials), lib.n_log_bins], np.int32)
    # further temperature sets of the nuclides follow them as records nuc<n>, nuc<n + 1>, ... (their library entries);
    # lib.set_of names the nuclide of each
    sets = lib.temperature_sets() if hasattr(lib, "temperature_sets") else []
    if sets:
        r["lib.set_of"] = np.array([i for i, _ in sets], np.int32)
    for k, n in enumerate(list(lib.nuclides) + [m for _, m in sets]):
        r[f"nuc{k}.ints"] = np.array([n.energy.size, int(n.fissionable), int(n.inelastic.size > 0), n.inel_threshold,
                                      n.n_precursor], np.int32)
        r[f"nuc{k}.scalars"] = np.array([n.awr, n.kT, n.inel_threshold_energy, n.inel_mass_ratio, n.nu_poly[0],
                                         n.nu_poly[1], n.watt_a, n.watt_b, n.q_fission], np.float64)
        r[f"nuc{k}.energy"] = np.ascontiguousarray(n.energy, np.float64)
        r[f"nuc{k}.xs"] = np.ascontiguousarray(n.xs, np.float64).ravel()
        r[f"nuc{k}.elastic"] = np.ascontiguousarray(n.elastic, np.float64)
        r[f"nuc{k}.inelastic"] = np.ascontiguousarray(n.inelastic, np.float64)
        r[f"nuc{k}.delayed_yield"] = np.ascontiguousarray(n.delayed_yield, np.float64)
        r[f"nuc{k}.delayed_theta"] = np.ascontiguousarray(n.delayed_theta, np.float64)
        # tabulated secondary distributions (mcw_tabular2d), when the nuclide has them
        for key, t in (("el_angle", getattr(n, "elastic_angle", None)), ("inel_angle", getattr(n, "inelastic_angle", None)),
                       ("inel_energy", getattr(n, "inelastic_energy", None))):
            if t is None:
                continue
            nd = t.n_discrete if t.n_discrete is not None else np.zeros(len(t.energy), np.int32)
            r[f"nuc{k}.{key}.i"] = np.concatenate([[int(t.histogram_in_energy), int(getattr(n, "inelastic_in_cm", True))],
                                                   t.offset, t.interpolation, nd]).astype(np.int32)
            r[f"nuc{k}.{key}.e"] = np.ascontiguousarray(t.energy, np.float64)
            r[f"nuc{k}.{key}.xpc"] = np.concatenate([t.x, t.pdf, t.cdf]).astype(np.float64)
        # product distributions that give energy and cosine together (MCW_LAW_CORRELATED / KALBACH_MANN / NBODY)
        law = getattr(n, "inelastic_law", -1)
        if law >= 2:
            li, lf = [law, int(n.inelastic_in_cm), 0], [0.0, 0.0]
            if law == 4:
                li[2], lf = int(n.nbody[0]), [float(n.nbody[1]), float(n.nbody[2])]
            elif law == 2:
                off, interp, mu, pdf, cdf = n.corr_angle
                li, lf = np.concatenate([li, off, interp]), np.concatenate([lf, mu, pdf, cdf])
            elif law == 3:
                lf = np.concatenate([lf, n.km_r, n.km_a])
            r[f"nuc{k}.law.i"] = np.asarray(li, np.int32)  # laws 5 / 6 / 7: law and in_cm only (parameters: rx.a / rx.b)
            r[f"nuc{k}.law.f"] = np.asarray(lf, np.float64)
        # fission neutrons as a nuclide file has them: products (prompt, delayed groups) and the total nu
        prods = getattr(n, "fission_products", None) or []

        def fn(key, f):  # a mcw_function1d: [kind, interpolation] + x + y
            if f is not None:
                r[key] = np.concatenate([[f.kind, f.interpolation], [] if f.x is None else f.x, f.y]).astype(np.float64)
        if prods:
            r[f"nuc{k}.fis.head"] = np.array([len(prods)] + [p.law for p in prods], np.int32)
            r[f"nuc{k}.fis.u"] = np.array([p.u for p in prods], np.float64)
            fn(f"nuc{k}.fis.nu", getattr(n, "total_nu", None))
            for j, pr in enumerate(prods):
                q = f"nuc{k}.fis{j}."
                fn(q + "y", pr.yield_)
                fn(q + "a", pr.a)
                fn(q + "b", pr.b)
                t = pr.energy
                if t is not None:
                    nd = t.n_discrete if t.n_discrete is not None else np.zeros(len(t.energy), np.int32)
                    r[q + "en.i"] = np.concatenate([[int(t.histogram_in_energy), 0], t.offset, t.interpolation, nd]).astype(np.int32)
                    r[q + "en.e"] = np.ascontiguousarray(t.energy, np.float64)
                    r[q + "en.xpc"] = np.concatenate([t.x, t.pdf, t.cdf]).astype(np.float64)
        # the scattering reactions after the first (mcw_reaction_desc) and the first one's yield
        more = getattr(n, "more_reactions", None) or []
        law0 = getattr(n, "inelastic_law", -1)
        if more or getattr(n, "inelastic_yield", 1.0) != 1.0 or law0 >= 5 or getattr(n, "inelastic_yield_fn", None) is not None:
            r[f"nuc{k}.rx.head"] = np.array([len(more), float(getattr(n, "inelastic_yield", 1.0)),
                                             float(getattr(n, "inelastic_u", 0.0))], np.float64)
            fn(f"nuc{k}.rx.y", getattr(n, "inelastic_yield_fn", None))
            fn(f"nuc{k}.rx.a", getattr(n, "inelastic_a", None))
            fn(f"nuc{k}.rx.b", getattr(n, "inelastic_b", None))
        for j, x in enumerate(more):
            q = f"nuc{k}.rx{j}."
            r[q + "i"] = np.array([x.mt, x.threshold, x.law, int(x.in_cm), 0 if x.nbody is None else int(x.nbody[0])], np.int32)
            r[q + "f"] = np.array([x.threshold_energy, x.mass_ratio, x.yield_, 0.0 if x.nbody is None else x.nbody[1],
                                   0.0 if x.nbody is None else x.nbody[2]], np.float64)
            r[q + "xs"] = np.ascontiguousarray(x.xs, np.float64)
            r[q + "u"] = np.array([float(x.u)], np.float64)
            fn(q + "y", x.yield_fn)
            fn(q + "a", x.a)
            fn(q + "b", x.b)
            for key, t in (("ang", x.angle), ("en", x.energy)):
                if t is None:
                    continue
                nd = t.n_discrete if t.n_discrete is not None else np.zeros(len(t.energy), np.int32)
                r[q + key + ".i"] = np.concatenate([[int(t.histogram_in_energy), int(x.in_cm)], t.offset, t.interpolation,
                                                    nd]).astype(np.int32)
                r[q + key + ".e"] = np.ascontiguousarray(t.energy, np.float64)
                r[q + key + ".xpc"] = np.concatenate([t.x, t.pdf, t.cdf]).astype(np.float64)
            if x.law == 2:
                off, interp, mu, pdf, cdf = x.corr_angle
                r[q + "corr.i"] = np.concatenate([off, interp]).astype(np.int32)
                r[q + "corr.f"] = np.concatenate([mu, pdf, cdf]).astype(np.float64)
            elif x.law == 3:
                r[q + "km"] = np.concatenate([x.km_r, x.km_a]).astype(np.float64)
    for k, m in enumerate(lib.materials):
        r[f"mat{k}.nuclide"] = np.ascontiguousarray(m.nuclides, np.int32)
        r[f"mat{k}.density"] = np.ascontiguousarray(m.densities, np.float64)
        iso = getattr(m, "isotropic", None)
        r[f"mat{k}.isotropic"] = np.ascontiguousarray(iso if iso is not None else np.zeros(len(m.nuclides)), np.int32)
        if getattr(m, "sab", None):  # (position, table, fraction)
            r[f"mat{k}.sab_i"] = np.array([[a[0], a[1]] for a in m.sab], np.int32).ravel()
            r[f"mat{k}.sab_f"] = np.array([a[2] for a in m.sab], np.float64)
    # thermal scattering tables (mcw_thermal_desc) and probability tables (mcw_urr_desc)
    th = getattr(lib, "thermal_tables", None) or []
    if th and any(getattr(t, "more_temperatures", None) for t in th):
        # every temperature of every table as one entry; lib.sab_same names the first entry of each entry's table
        th, same, kT = lib.thermal_entries()
        r["lib.sab_same"], r["lib.sab_kT"] = same, kT
    r["lib.n_thermal"] = np.array([len(th)], np.int32)
    for k, t in enumerate(th):
        zero = np.zeros(0)
        r[f"sab{k}.ints"] = np.array([t.inelastic_type, t.elastic_type, t.n_mu, t.n_e_out, int(t.skewed),
                                      0 if t.elastic_mu is None else t.elastic_mu.shape[1]], np.int32)
        r[f"sab{k}.emax"] = np.array([t.energy_max], np.float64)
        for key in ("inelastic_xs_energy", "inelastic_xs", "inelastic_energy", "inelastic_e_out", "inelastic_pdf",
                    "inelastic_cdf", "inelastic_mu", "elastic_x", "elastic_y", "elastic_mu"):
            a = getattr(t, key)
            r[f"sab{k}.{key.replace('inelastic_', 'i_').replace('elastic_', 'e_')}"] = \
                np.ascontiguousarray(a if a is not None else zero, np.float64).ravel()
        r[f"sab{k}.i_offset"] = np.ascontiguousarray(t.inelastic_offset if t.inelastic_offset is not None else [], np.int32)
    ur = getattr(lib, "urr_tables", None) or []
    r["lib.n_urr"] = np.array([len(ur)], np.int32)
    for k, t in enumerate(ur):
        r[f"urr{k}.ints"] = np.array([t.nuclide, t.interpolation, t.inelastic_flag, t.absorption_flag, int(t.multiply_smooth),
                                      t.energy.size, t.table.shape[2]], np.int32)
        r[f"urr{k}.energy"] = np.ascontiguousarray(t.energy, np.float64)
        r[f"urr{k}.table"] = np.ascontiguousarray(t.table, np.float64).ravel()
    r["kernel.blocks"] = np.ascontiguousarray(blocks, np.int32)
    r["kernel.materials"] = np.ascontiguousarray(materials, np.int32)  # OpenMC material ids, 1-based
    if getattr(lib, "element_temperature", None) is not None:
        temperature = lib.element_temperature
    r["kernel.temperature"] = np.atleast_1d(np.asarray(temperature, np.float64))
    s = settings
    r["settings.ints"] = np.array([s.n_particles, s.master_seed, s.n_inactive, s.n_log_bins, s.is_2d,
                                   s.seed_through_double, s.source_only_fissionable, s.max_events, s.sort_xs_queue,
                                   s.record_traces], np.int64)
    r["settings.reals"] = np.array([s.z_coord] + [s.source_box[i] for i in range(6)], np.float64)
    r["tallies.count"] = np.array([len(tallies)], np.int32)
    for k, tl in enumerate(tallies):
        tid, estimator, granularity, scores, ebins = tl[:5]
        r[f"tally{k}.ints"] = np.array([tid, estimator, granularity], np.int32)
        r[f"tally{k}.scores"] = np.ascontiguousarray(scores, np.int32)
        r[f"tally{k}.energy_bins"] = np.ascontiguousarray(ebins if ebins is not None else [], np.float64)
        if len(tl) > 5 and tl[5] is not None:  # nuclide bins (library indices, -1 total, -2 never scored)
            r[f"tally{k}.nuclides"] = np.ascontiguousarray(tl[5], np.int32)
    return r


def write_problem(path, mesh, lib, blocks, materials, settings, tallies=(), temperature=293.6):
    """tallies: iterable of (id, estimator, granularity, scores, energy_bins | None[, nuclide bins]) in C-ABI codes."""
    recs = _records(mesh, lib, blocks, materials, settings, list(tallies), temperature)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(recs)))
        for name, a in recs.items():
            nb = name.encode()
            assert len(nb) < 32
            f.write(nb.ljust(32, b"\0"))
            f.write(struct.pack("<IIQ", _CODE[a.dtype], 0, a.size))
            b = a.tobytes()
            f.write(b + b"\0" * (-len(b) % 8))


def read_records(path):
    """Inverse of write_problem at the record level (used by the CPU tests)."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            name = f.read(32).rstrip(b"\0").decode()
            code, _, cnt = struct.unpack("<IIQ", f.read(16))
            dt = np.dtype(_DT[code])
            nbytes = cnt * dt.itemsize
            out[name] = np.frombuffer(f.read(nbytes), dt).copy()
            f.read(-nbytes % 8)
    return out


def write_library(path, lib):
    """Nuclide table file for the C++ input-file front end (host/input_file.hpp): the continuous-energy
    tables by nuclide name, in place of cross_sections.xml + HDF5."""
    from . import cstructs, meshgen
    recs = {k: v for k, v in _records(meshgen.infinite_medium(1), lib, [0], [1], cstructs.make_settings(1), [], 0.0).items()
            if k.startswith("nuc")}
    recs["lib.names"] = np.frombuffer(";".join(n.name for n in lib.nuclides).encode(), np.uint8).astype(np.int32)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(recs)))
        for name, a in recs.items():
            f.write(name.encode().ljust(32, b"\0"))
            f.write(struct.pack("<IIQ", _CODE[a.dtype], 0, a.size))
            b = a.tobytes()
            f.write(b + b"\0" * (-len(b) % 8))


if __name__ == "__main__":
    import argparse
    from . import cstructs, exodus, nucdata
    ap = argparse.ArgumentParser(description="nuclide table files (--library) and problem files (--problem) for host/mcw_run")
    ap.add_argument("--library", help="write the synthetic nuclide table file read by mcw_run --tables")
    ap.add_argument("--names-from", help="materials.xml: one synthetic nuclide per nuclide name found there")
    ap.add_argument("--problem", help="write a problem file from an Exodus mesh (FileMeshGenerator inputs)")
    ap.add_argument("--exodus", help="Exodus-II mesh (classic NetCDF), e.g. examples/beavrs_assembly/31enr_16_in.e")
    ap.add_argument("--reflect", type=int, nargs="*", default=[], help="side-set ids with ReflectRayBC")
    ap.add_argument("--kill", type=int, nargs="*", default=[], help="side-set ids with KillRayBC")
    ap.add_argument("--blocks", type=int, nargs="*", help="CollisionKernel blocks")
    ap.add_argument("--materials", type=int, nargs="*", help="CollisionKernel materials (1-based ids into the library's materials)")
    ap.add_argument("--particles", type=int, default=100000)
    ap.add_argument("--inactive", type=int, default=0)
    ap.add_argument("--box", type=float, nargs=6, help="source box: x0 y0 z0 x1 y1 z1")
    ap.add_argument("--z-coord", type=float, default=0.0)
    ap.add_argument("--nuclides", type=int, default=9)
    ap.add_argument("--grid-points", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    lib = nucdata.make_library(a.nuclides, G=a.grid_points, seed=a.seed)
    if a.names_from:
        import xml.etree.ElementTree as ET
        names = []
        for n in ET.parse(a.names_from).getroot().iter("nuclide"):
            if n.get("name") not in names:
                names.append(n.get("name"))
        lib = nucdata.library_for_names(names, G=a.grid_points, seed=a.seed)
    if a.library:
        write_library(a.library, lib)
        print(a.library)
    if a.problem:
        if not (a.exodus and a.blocks and a.materials and a.box):
            ap.error("--problem needs --exodus, --blocks, --materials and --box")
        mesh = exodus.read_exodus(a.exodus, reflect=a.reflect, kill=a.kill)
        s = cstructs.make_settings(a.particles, n_inactive=a.inactive, is_2d=mesh.dim == 2, z_coord=a.z_coord,
                                   source_box=a.box, record_traces=0)
        write_problem(a.problem, mesh, lib, a.blocks, a.materials, s, [(1, 0, 0, [0], None)])
        print(a.problem, mesh.n_elem, "elements")
