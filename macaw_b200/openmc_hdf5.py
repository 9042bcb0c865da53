"""OpenMC nuclide files -> macaw_b200.nucdata.Library: the data hand-over of INTEGRATION.md section 3.2 for a host that
reads the HDF5 library itself instead of walking openmc::data::nuclides.

The functions take an h5py-like object -- ``g[name]``, ``name in g``, ``g.keys()``, ``g.attrs[name]``, datasets as
``d[()]`` with ``d.attrs`` -- so ``h5py.File("U235.h5")["U235"]`` works where h5py exists.  This image has neither h5py
nor nuclide files; what is here is held against the reference instead: oracle/_ref builds a nuclide-file
image in memory, the REFERENCE's own Nuclide / Reaction / ReactionProduct / ... constructors read it, and
tests/test_oracle_pinning.py::test_nuclide_file_converter_against_reference_nuclide_cpp reads the same image through
`read_image` + `nuclide_from_group` and finds the reference's derived tables (Nuclide::create_derived) bit for bit, its
reaction list, temperatures and probability tables, and that the converted library collides like the original.

Layout followed (the reference's readers): Nuclide::Nuclide OMC nuclide.cpp:50-302, create_derived :317-444, Reaction
reaction.cpp:20-65, ReactionProduct reaction_product.cpp:18-68, read_function endf.cpp:80-100, Tabulated1D :127-153,
AngleDistribution distribution_angle.cpp:20-63, UncorrelatedAngleEnergy secondary_uncorrelated.cpp:17-52, the energy laws
distribution_energy.cpp:20-148,265-340, CorrelatedAngleEnergy secondary_correlated.cpp:22-150, KalbachMann
secondary_kalbach.cpp:24-110, NBodyPhaseSpace secondary_nbody.cpp:16-23, UrrData urr.cpp:7-33.

What the library does not carry is refused with the reason (NotImplementedError), never dropped: see `Refused`."""
from __future__ import annotations

import struct

import numpy as np

from . import nucdata

K_BOLTZMANN = 8.617333262e-5  # OMC constants.h:90


class Refused(NotImplementedError):
    """A nuclide (or one of its reactions) uses something the library does not implement."""


# ---------------------------------------------------------------------------------------------- h5py-like image reader
class _Attrs(dict):
    pass


class MemNode:
    """One group or dataset of the image written by oracle/_ref ref_dump_tree, with the slice of h5py's API used below."""

    def __init__(self, name, is_dataset):
        self.name, self.is_dataset, self.attrs, self.children, self.data = name, is_dataset, _Attrs(), {}, None

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        if path == ():  # dataset[()]
            return self.data
        node = self
        for part in str(path).strip("/").split("/"):
            node = node.children[part]
        return node

    def keys(self):
        return list(self.children.keys())

    @property
    def shape(self):
        return self.data.shape


def read_image(buf: bytes) -> MemNode:
    """Parses the byte stream of ref_dump_tree (oracle/ref_shim_physics.cpp) into a tree of MemNode."""
    pos = 0

    def u32():
        nonlocal pos
        (v,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        return v

    def text():
        nonlocal pos
        n = u32()
        s = buf[pos:pos + n].decode()
        pos += n
        return s

    def array():
        nonlocal pos
        kind = buf[pos]
        pos += 1
        nd = u32()
        dims = struct.unpack_from(f"<{nd}Q", buf, pos)
        pos += 8 * nd
        if kind == 2:
            return text()
        n = u32()
        a = np.frombuffer(buf, np.float64 if kind == 0 else np.int64, n, pos).copy()
        pos += 8 * n
        return a.reshape(dims) if nd else a[0]

    def node():
        nonlocal pos
        tag = chr(buf[pos])
        pos += 1
        nd = MemNode(text(), tag == "D")
        for _ in range(u32()):
            k = text()
            nd.attrs[k] = array()
        if nd.is_dataset:
            nd.data = array()
        for _ in range(u32()):
            c = node()
            nd.children[c.name] = c
        return nd
    return node()


# ------------------------------------------------------------------------------------------------------------ helpers
def _s(x):
    """attribute strings come as str, bytes or numpy scalars depending on the reader"""
    if isinstance(x, bytes):
        return x.decode()
    if isinstance(x, np.ndarray) and x.shape == ():
        x = x[()]
    return x.decode() if isinstance(x, bytes) else str(x)


def _a(x):
    return np.atleast_1d(np.asarray(x))


def _data(d):
    return np.asarray(d[()])


def is_fission(mt):  # OMC endf.cpp:37-40
    return mt in (18, 19, 20, 21, 38)


def is_disappearance(mt):  # OMC endf.cpp:42-51
    return 101 <= mt <= 117 or mt in (155, 182, 191, 192, 193, 197)


def is_inelastic_scatter(mt):  # OMC endf.cpp:63-78
    if mt < 100:
        return (not is_fission(mt)) and mt >= 5 and mt != 27
    if mt <= 200:
        return not is_disappearance(mt)
    return 875 <= mt <= 891


def read_function(d) -> nucdata.Function1D:
    """read_function (OMC endf.cpp:80-100): Polynomial or Tabulated1D; one interpolation region, histogram or lin-lin"""
    kind = _s(d.attrs["type"])
    v = _data(d)
    if kind == "Polynomial":
        return nucdata.Function1D(1, np.asarray(v, np.float64).ravel())
    if kind != "Tabulated1D":
        raise Refused(f"function type '{kind}'")
    interp, nbt = _a(d.attrs["interpolation"]), _a(d.attrs["breakpoints"])
    # several regions that all interpolate the same way are one region (Tabulated1D::operator() only looks the code up,
    # endf.cpp:172-184)
    if np.any(interp != interp[0]) or int(nbt[-1]) != v.shape[1] or int(interp[0]) not in (1, 2):
        raise Refused("Tabulated1D with several regions of different interpolation, or logarithmic interpolation")
    return nucdata.Function1D(2, np.asarray(v[1], np.float64), np.asarray(v[0], np.float64), int(interp[0]))


def _tabular2d(energy, table, offsets, interpolation, n_discrete=None, histogram=False) -> nucdata.Tabular2D:
    """(3, N) x / pdf / cdf with `offsets` (start of every distribution) -> nucdata.Tabular2D"""
    off = np.concatenate([_a(offsets).astype(np.int64), [table.shape[1]]]).astype(np.int32)
    return nucdata.Tabular2D(np.asarray(energy, np.float64), off, _a(interpolation).astype(np.int32), np.asarray(table[0], np.float64),
                             np.asarray(table[1], np.float64), np.asarray(table[2], np.float64),
                             None if n_discrete is None else _a(n_discrete).astype(np.int32), bool(histogram))


def _one_region(d, what):
    """the `interpolation` attribute of an incoming-energy grid: (2, n_regions) breakpoints / codes; one region only"""
    it = np.asarray(d.attrs["interpolation"]).reshape(2, -1)
    if it.shape[1] != 1 or int(it[1, 0]) not in (1, 2):
        raise Refused(f"{what}: several incoming-energy interpolation regions")
    return int(it[1, 0]) == 1


class _Product:
    """What one neutron product gives: law, tables, parameters (the fields Reaction / FissionProduct / Nuclide share)"""
    law = 0
    in_cm = True
    angle = energy = corr_angle = km_r = km_a = nbody = a = b = None
    u = 0.0
    threshold_energy = mass_ratio = 0.0
    yield_const = 1.0
    yield_fn = None


def read_product(p, awr) -> _Product:
    """ReactionProduct (OMC reaction_product.cpp:18-68) with its one distribution"""
    out = _Product()
    if int(np.asarray(p.attrs["n_distribution"])) != 1:
        raise Refused("a product with several distributions (applicability)")
    y = read_function(p["yield"])
    if y.kind == 1 and y.y.size == 1:
        out.yield_const = float(y.y[0])
    else:
        out.yield_fn = y
    d = p["distribution_0"]
    kind = _s(d.attrs["type"])
    if kind == "uncorrelated":
        if "angle" in d:
            a = d["angle"]
            mu = a["mu"]
            out.angle = _tabular2d(_data(a["energy"]), _data(mu), mu.attrs["offsets"], mu.attrs["interpolation"])
        e = d["energy"] if "energy" in d else None
        law = _s(e.attrs["type"]) if e is not None else None
        if law is None:
            out.law = -1  # no energy law: elastic scattering (its energy follows from the kinematics)
        elif law == "level":
            out.law, out.threshold_energy, out.mass_ratio = 0, float(np.asarray(e.attrs["threshold"])), float(np.asarray(e.attrs["mass_ratio"]))
        elif law == "continuous":
            dist = e["distribution"]
            out.law = 1
            out.energy = _tabular2d(_data(e["energy"]), _data(dist), dist.attrs["offsets"], dist.attrs["interpolation"],
                                    dist.attrs["n_discrete_lines"], _one_region(e["energy"], "continuous tabular law"))
        elif law in ("maxwell", "evaporation", "watt"):
            out.law = {"maxwell": 5, "evaporation": 6, "watt": 7}[law]
            out.u = float(np.asarray(e.attrs["u"]))
            out.a = read_function(e["a"] if law == "watt" else e["theta"])
            out.b = read_function(e["b"]) if law == "watt" else None
        else:
            raise Refused(f"energy distribution type '{law}'")  # OMC secondary_uncorrelated.cpp:47
    elif kind in ("correlated", "kalbach-mann"):
        _one_region(d["energy"], kind)
        eo = d["energy_out" if kind == "correlated" else "distribution"]
        t = _data(eo)
        if np.any(_a(eo.attrs["n_discrete_lines"]) != 0):
            raise Refused(f"{kind} distribution with discrete lines")
        out.energy = _tabular2d(_data(d["energy"]), t[:3], eo.attrs["offsets"], eo.attrs["interpolation"])
        if kind == "correlated":
            mu = _data(d["mu"])
            off = np.concatenate([t[4].astype(np.int64), [mu.shape[1]]]).astype(np.int32)
            out.law, out.corr_angle = 2, (off, t[3].astype(np.int32), mu[0].copy(), mu[1].copy(), mu[2].copy())
        else:
            out.law, out.km_r, out.km_a = 3, t[3].copy(), t[4].copy()
    elif kind == "nbody":
        out.law = 4
        out.nbody = (int(np.asarray(d.attrs["n_particles"])), float(np.asarray(d.attrs["total_mass"])), float(np.asarray(d.attrs["q_value"])))
    else:
        raise Refused(f"angle-energy distribution type '{kind}'")
    return out


def _neutron_products(rx):
    out = []
    for k in sorted((n for n in rx.keys() if n.startswith("product_")), key=lambda n: int(n.split("_")[1])):
        if _s(rx[k].attrs["particle"]) == "neutron":
            out.append(rx[k])
    return out


def nuclide_from_group(g, name=None):
    """One nuclide group of a nuclide file -> (nucdata.Nuclide with more_temperatures, [(temperature index, UrrTable)]).
    Temperature index 0 is the nuclide's own entry (the lowest temperature), k > 0 its (k - 1)-th further set."""
    name = name or getattr(g, "name", "nuclide").strip("/")
    awr = float(np.asarray(g.attrs["atomic_weight_ratio"]))
    temps = sorted(g["kTs"].keys(), key=lambda t: float(np.asarray(_data(g["kTs"][t]))))
    rxs = g["reactions"]
    rx_names = sorted(n for n in rxs.keys() if n.startswith("reaction_"))  # name order = MT order (HDF5's link index)
    sets = []
    for T in temps:
        energy = np.asarray(_data(g["energy"][T]), np.float64)
        G = energy.size
        total, absorption, fission = np.zeros(G), np.zeros(G), np.zeros(G)
        n = nucdata.Nuclide(name, awr, energy, None, None, np.zeros(0))
        n.kT = float(np.asarray(_data(g["kTs"][T])))
        scatter = []
        for rn in rx_names:
            rx = rxs[rn]
            mt = int(np.asarray(rx.attrs["mt"]))
            xs_d = rx[T]["xs"]
            thr = int(np.asarray(xs_d.attrs["threshold_idx"]))
            xs = np.asarray(_data(xs_d), np.float64)
            if int(np.asarray(rx.attrs["redundant"])):
                continue
            # Nuclide::create_derived (OMC nuclide.cpp:317-390), same order of additions
            total[thr:thr + xs.size] += xs
            if is_disappearance(mt):
                absorption[thr:thr + xs.size] += xs
            if is_fission(mt):
                if mt != 18:
                    raise Refused("partial fission reactions (MT 19 / 20 / 21 / 38)")
                fission[thr:thr + xs.size] += xs
                absorption[thr:thr + xs.size] += xs
                n.fissionable = True
                n.q_fission = float(np.asarray(rx.attrs["Q_value"]))
                prods = [read_product(p, awr) for p in _neutron_products(rx)]
                n.fission_products = []
                for k, (p, node) in enumerate(zip(prods, _neutron_products(rx))):
                    if p.law not in (1, 5, 6, 7) or p.angle is not None:
                        raise Refused("fission product: law other than Watt / Maxwell / evaporation / continuous tabular, or an angle table")
                    y = p.yield_fn or nucdata.Function1D(1, np.array([p.yield_const]))
                    n.fission_products.append(nucdata.FissionProduct(y, p.law, p.u, p.a, p.b, p.energy))
                    if k > 0 and _s(node.attrs["emission_mode"]) != "delayed":
                        raise Refused("fission reaction with several prompt neutron products")
                n.n_precursor = len(n.fission_products) - 1
            elif mt == 2:
                if thr != 0:
                    raise Refused("elastic scattering with a threshold")
                n.elastic = xs
                p = read_product(_neutron_products(rx)[0], awr)
                n.elastic_angle = p.angle
            elif is_inelastic_scatter(mt):
                p = read_product(_neutron_products(rx)[0], awr)
                if p.law < 0:
                    raise Refused(f"MT {mt}: neutron product without an energy law")
                p.in_cm = bool(int(np.asarray(rx.attrs["center_of_mass"])))
                scatter.append((mt, thr, xs, p))
        if n.elastic is None:
            raise Refused("no elastic scattering reaction")
        if "total_nu" in g and n.fissionable:
            n.total_nu = read_function(g["total_nu"]["yield"])
        nu = (n.total_nu or n.fission_products[0].yield_)(energy) if n.fissionable else 0.0
        n.xs = np.stack([total, absorption, fission, nu * fission, np.zeros(G)], axis=1)
        # the scattering reactions in index_inelastic_scatter_ order: the first inline, the others as a list
        for j, (mt, thr, xs, p) in enumerate(scatter):
            if j == 0:
                n.inelastic, n.inel_threshold = xs, thr
                n.inel_threshold_energy, n.inel_mass_ratio = p.threshold_energy, p.mass_ratio
                n.inelastic_angle, n.inelastic_energy, n.inelastic_in_cm = p.angle, p.energy, p.in_cm
                n.inelastic_law = p.law if p.law >= 2 else -1
                n.corr_angle, n.km_r, n.km_a, n.nbody = p.corr_angle, p.km_r, p.km_a, p.nbody
                n.inelastic_yield, n.inelastic_yield_fn = p.yield_const, p.yield_fn
                n.inelastic_u, n.inelastic_a, n.inelastic_b = p.u, p.a, p.b
            else:
                n.more_reactions = n.more_reactions or []
                n.more_reactions.append(nucdata.Reaction(mt, thr, xs, p.threshold_energy, p.mass_ratio, p.yield_const, p.law, p.in_cm,
                                                         p.angle, p.energy, p.corr_angle, p.km_r, p.km_a, p.nbody, p.yield_fn, p.u, p.a, p.b))
        sets.append(n)
    base = sets[0]
    base.more_temperatures = sets[1:] or None
    urr = []
    if "urr" in g:
        for k, T in enumerate(temps):
            u = g["urr"][T]
            if int(np.asarray(u.attrs["interpolation"])) != 2:
                raise Refused("log-log probability tables")
            urr.append((k, nucdata.UrrTable(-1, np.asarray(_data(u["energy"]), np.float64), np.asarray(_data(u["table"]), np.float64),
                                            int(np.asarray(u.attrs["inelastic"])), int(np.asarray(u.attrs["absorption"])),
                                            bool(int(np.asarray(u.attrs["multiply_smooth"]))))))
    return base, urr


def library_from_groups(groups, materials=(), n_log_bins=8000) -> nucdata.Library:
    """groups: [(name, h5py-like nuclide group)] in the order of data::nuclides (the order materials refer to)"""
    nuclides, pending = [], []
    for name, g in groups:
        n, urr = nuclide_from_group(g, name)
        nuclides.append(n)
        pending.append(urr)
    lib = nucdata.Library(nuclides, list(materials), n_log_bins)
    # library entries: the nuclides, then their further temperature sets in nuclide order
    entry = {}
    e = len(nuclides)
    for i, n in enumerate(nuclides):
        entry[(i, 0)] = i
        for k in range(len(n.more_temperatures or [])):
            entry[(i, k + 1)] = e
            e += 1
    tables = []
    for i, urr in enumerate(pending):
        for k, t in urr:
            t.nuclide = entry[(i, k)]
            tables.append(t)
    lib.urr_tables = tables or None
    return lib


def library_from_files(paths, materials=(), n_log_bins=8000) -> nucdata.Library:
    """OpenMC nuclide files (one nuclide group each, e.g. the NNDC ENDF/B-VII.1 library MaCaw downloads,
    scripts/download-openmc-cross-sections.sh) in the order of data::nuclides.  Needs h5py, which this image does not have:
    untested here beyond the layout pins described at the top."""
    import h5py  # noqa: PLC0415 -- absent in the build image; the ImportError is the message
    groups = []
    for p in paths:
        f = h5py.File(p, "r")
        name = [k for k in f.keys()][0]
        groups.append((name, f[name]))
    return library_from_groups(groups, materials, n_log_bins)


def thermal_from_group(g, name=None) -> nucdata.ThermalTable:
    """One table group of a thermal-scattering file (OMC ThermalScattering thermal.cpp:38-147, ThermalData :177-227) ->
    nucdata.ThermalTable with its further temperatures in more_temperatures.  Inelastic: incoherent_inelastic (continuous
    outgoing energies, equiprobable discrete cosines per outgoing energy) or incoherent_inelastic_discrete; elastic: none,
    coherent_elastic (Bragg edges) or incoherent_elastic_discrete.  The Debye-Waller form is refused."""
    name = name or getattr(g, "name", "table").strip("/")
    energy_max = float(np.asarray(g.attrs["energy_max"]))
    temps = sorted(g["kTs"].keys(), key=lambda t: float(np.asarray(_data(g["kTs"][t]))))
    sets = []
    for T in temps:
        tg = g[T]
        inel = tg["inelastic"]
        xs = read_function(inel["xs"])
        if xs.kind != 2 or xs.interpolation != 2:
            raise Refused("thermal inelastic cross section that is not a lin-lin table")
        d = inel["distribution"]
        kind = _s(d.attrs["type"])
        if kind == "incoherent_inelastic":
            _one_region(d["energy"], "thermal inelastic distribution")
            eo, mu = _data(d["energy_out"]), _data(d["mu"])
            off = _a(d["energy_out"].attrs["offsets"]).astype(np.int64)
            n_pts = eo.shape[1]
            if np.any(eo[3] != 0) or n_pts == 0 or mu.shape[1] % n_pts:
                raise Refused("thermal inelastic distribution whose cosines are not equiprobable discrete ones")
            n_mu = mu.shape[1] // n_pts
            if np.any(eo[4] != np.arange(n_pts) * n_mu):
                raise Refused("thermal inelastic distribution with a varying number of cosines")
            t = nucdata.ThermalTable(name, energy_max, 0, xs.x, xs.y, n_mu)
            t.inelastic_energy = np.asarray(_data(d["energy"]), np.float64)
            t.inelastic_offset = np.concatenate([off, [n_pts]]).astype(np.int32)
            t.inelastic_e_out, t.inelastic_pdf, t.inelastic_cdf = eo[0].copy(), eo[1].copy(), eo[2].copy()
            t.inelastic_mu = mu[0].reshape(n_pts, n_mu).copy()
        elif kind == "incoherent_inelastic_discrete":
            e_out, mu_out = _data(d["energy_out"]), _data(d["mu_out"])
            t = nucdata.ThermalTable(name, energy_max, 1, xs.x, xs.y, int(mu_out.shape[2]))
            t.n_e_out, t.skewed = int(e_out.shape[1]), bool(np.asarray(_data(d["skewed"])))
            t.inelastic_e_out, t.inelastic_mu = e_out.copy(), mu_out.copy()
        else:
            raise Refused(f"thermal inelastic distribution type '{kind}'")
        if "elastic" in tg:
            el = tg["elastic"]
            ekind = _s(el["distribution"].attrs["type"])
            if ekind == "coherent_elastic":
                v = _data(el["xs"])
                t.elastic_type, t.elastic_x, t.elastic_y = 1, v[0].copy(), v[1].copy()
            elif ekind == "incoherent_elastic_discrete":
                f = read_function(el["xs"])
                t.elastic_type, t.elastic_x, t.elastic_y = 2, f.x, f.y
                t.elastic_mu = _data(el["distribution"]["mu_out"]).copy()
            else:
                raise Refused(f"thermal elastic distribution type '{ekind}' (the Debye-Waller form is not carried)")
        t.kT = float(np.asarray(_data(g["kTs"][T])))
        sets.append(t)
    sets[0].more_temperatures = sets[1:] or None
    # the nuclides the table binds (ThermalScattering::nuclides_, thermal.cpp:49; Material assigns tables by them,
    # material.cpp:317-360): kept on the returned object for the host that builds Material.sab
    who = g.attrs["nuclides"]
    sets[0].applies_to = [_s(w) for w in (who if isinstance(who, (list, tuple, np.ndarray)) and np.ndim(who) else [who])]
    return sets[0]
