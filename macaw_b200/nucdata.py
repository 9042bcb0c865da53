"""Synthetic continuous-energy nuclide library of a named shape (n_nuclides x energy points).

ENDF/B-VII.1 HDF5 files are not available offline, so runs use in-memory tables whose
*structure* is what OpenMC builds in ``Nuclide::create_derived`` (openmc/src/nuclide.cpp:309-444):
a per-nuclide strictly increasing energy grid on [1e-5, 2e7] eV clustered around resonances, and the
derived 5-column table (total, absorption, fission, nu-fission, photon production) plus the partial
elastic and level-inelastic cross sections the collision physics samples from.  All numbers come from
``numpy.random.Generator(PCG64(seed))`` so every run, CPU or GPU, reads identical bytes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

E_MIN, E_MAX = 1.0e-5, 2.0e7
KT_ROOM = 8.617333262e-5 * 293.6  # eV (OpenMC constants.h:90, temperature_default 293.6 K)
MAX_DELAYED_GROUPS = 8


@dataclass
class Nuclide:
    name: str
    awr: float
    energy: np.ndarray
    xs: np.ndarray         # (G, 5)
    elastic: np.ndarray    # (G,)
    inelastic: np.ndarray  # (G - threshold,) or empty
    inel_threshold: int = 0
    inel_threshold_energy: float = 0.0
    inel_mass_ratio: float = 0.0
    fissionable: bool = False
    nu_poly: tuple = (0.0, 0.0)
    delayed_yield: np.ndarray = field(default_factory=lambda: np.zeros(MAX_DELAYED_GROUPS))
    delayed_theta: np.ndarray = field(default_factory=lambda: np.zeros(MAX_DELAYED_GROUPS))
    n_precursor: int = 0
    watt_a: float = 0.988e6
    watt_b: float = 2.249e-6
    q_fission: float = 0.0
    kT: float = KT_ROOM
    # optional tabulated secondary distributions (Tabular2D below); None = isotropic in CM / level law
    elastic_angle: object = None
    inelastic_angle: object = None
    inelastic_energy: object = None  # outgoing-energy table of the inelastic reaction (then inelastic_in_cm as given)
    inelastic_in_cm: bool = True
    inelastic_law: int = -1          # -1: level, or continuous tabular when inelastic_energy is set; 2 correlated, 3 Kalbach-Mann, 4 N-body
    corr_angle: object = None        # law 2: (offset, interpolation, mu, pdf, cdf): one angular Tabular per point of inelastic_energy
    km_r: np.ndarray = None          # law 3: precompound fraction / slope per point of inelastic_energy
    km_a: np.ndarray = None
    nbody: tuple = None              # law 4: (n_particles, total_mass, q_value)
    inelastic_yield: float = 1.0     # neutron multiplicity of the inline inelastic reaction
    inelastic_yield_fn: object = None  # ... as a Function1D of the incoming energy
    inelastic_u: float = 0.0         # inline reaction with law 5 / 6 / 7: restriction energy and parameters
    inelastic_a: object = None
    inelastic_b: object = None
    more_reactions: list = None      # further scattering reactions (Reaction), index_inelastic_scatter_ order
    more_temperatures: list = None   # further data sets of this nuclide (Nuclide objects at other kT; OMC kTs_, grid_[t], xs_[t])
    fission_products: list = None    # FissionProduct x (1 + n_precursor): prompt, delayed groups (replaces nu_poly / watt / ...)
    total_nu: object = None          # Function1D or None (then the prompt yield stands in)


@dataclass
class Function1D:
    """mcw_function1d: kind 1 polynomial (y = coefficients), kind 2 tabulated (x, y; interpolation 1 histogram, 2 lin-lin)."""
    kind: int
    y: np.ndarray
    x: np.ndarray = None
    interpolation: int = 2

    def __call__(self, E):
        E = np.asarray(E, np.float64)
        if self.kind == 1:
            out = np.zeros_like(E)
            for c in self.y[::-1]:
                out = out * E + c
            return out
        # Tabulated1D::operator() (OMC endf.cpp:156-210) operation for operation, so that tables derived with it (nu-fission)
        # equal the reference's bit for bit: constant outside the range, i = lower_bound_index, r = (x - x0) / (x1 - x0)
        x, y = np.asarray(self.x, np.float64), np.asarray(self.y, np.float64)
        i = np.clip(np.searchsorted(x, E, side="left") - 1, 0, x.size - 2)
        i = np.where(E == x[0], 0, i)
        if self.interpolation == 1:
            out = y[i]
        else:
            r = (E - x[i]) / (x[i + 1] - x[i])
            out = y[i] + r * (y[i + 1] - y[i])
        return np.where(E < x[0], y[0], np.where(E > x[-1], y[-1], out))


@dataclass
class FissionProduct:
    """mcw_fission_product"""
    yield_: Function1D
    law: int                 # 7 Watt, 5 Maxwell, 6 evaporation, 1 continuous tabular
    u: float = -2.0e7
    a: Function1D = None     # Watt a(E) or theta(E)
    b: Function1D = None
    energy: object = None    # Tabular2D


@dataclass
class Reaction:
    """A further scattering reaction of a nuclide (mcw_reaction_desc; one entry of OpenMC's index_inelastic_scatter_)."""
    mt: int
    threshold: int                   # grid index
    xs: np.ndarray                   # (G - threshold,)
    threshold_energy: float = 0.0    # level law
    mass_ratio: float = 0.0
    yield_: float = 1.0
    law: int = 0                     # MCW_LAW_*
    in_cm: bool = True
    angle: object = None             # Tabular2D
    energy: object = None            # Tabular2D
    corr_angle: object = None
    km_r: np.ndarray = None
    km_a: np.ndarray = None
    nbody: tuple = None
    yield_fn: object = None          # Function1D: energy-dependent multiplicity (else the constant yield_)
    u: float = 0.0                   # laws 5 / 6 / 7 (Maxwell / evaporation / Watt): restriction energy and parameters
    a: object = None
    b: object = None


@dataclass
class Tabular2D:
    """Distributions on a grid of incoming energies (mcw_tabular2d): OpenMC's "angle" / "continuous" layouts."""
    energy: np.ndarray
    offset: np.ndarray         # (n_energy + 1,)
    interpolation: np.ndarray  # (n_energy,) 1 histogram, 2 lin-lin
    x: np.ndarray
    pdf: np.ndarray
    cdf: np.ndarray
    n_discrete: np.ndarray = None
    histogram_in_energy: bool = False


def _tabulated_pdf(rng, x, interp):
    """A positive pdf on the points x with the cdf the reference's Tabular / CTTable would rebuild from it,
    normalised so that cdf[-1] == 1.0 exactly (its own normalisation is then the identity)."""
    p = rng.uniform(0.2, 1.0, x.size)
    c = np.zeros(x.size)
    for i in range(1, x.size):
        c[i] = c[i - 1] + (p[i - 1] if interp == 1 else 0.5 * (p[i - 1] + p[i])) * (x[i] - x[i - 1])
    p, c = p / c[-1], c / c[-1]
    c[-1] = 1.0
    return p, c


def make_angle_table(rng, e_lo, e_hi, n_energy=5) -> Tabular2D:
    """Anisotropy growing with energy, alternating histogram / lin-lin distributions, 4-17 cosines each."""
    energy = np.exp(np.linspace(np.log(e_lo), np.log(e_hi), n_energy))
    xs, ps, cs, off, interp = [], [], [], [0], []
    for k in range(n_energy):
        n = int(rng.integers(4, 18))
        mu = np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)]))
        it = 1 if k % 2 else 2
        p, c = _tabulated_pdf(rng, mu, it)
        xs.append(mu); ps.append(p); cs.append(c); interp.append(it); off.append(off[-1] + n)
    return Tabular2D(energy, np.array(off, np.int32), np.array(interp, np.int32), np.concatenate(xs), np.concatenate(ps),
                     np.concatenate(cs))


def make_continuous_table(rng, e_thr, e_hi, n_energy=4, discrete=True) -> Tabular2D:
    """Outgoing energies up to ~0.8 (E_in - threshold) + a little; the first distributions carry discrete lines."""
    energy = np.exp(np.linspace(np.log(e_thr), np.log(e_hi), n_energy))
    xs, ps, cs, off, interp, nd = [], [], [], [0], [], []
    for k in range(n_energy):
        n_d = int(rng.integers(0, 3)) if (discrete and k < n_energy - 1) else 0
        n_c = int(rng.integers(3, 12))
        top = 0.8 * (energy[k] - 0.5 * e_thr) + 1.0e3
        e_c = np.sort(np.concatenate([[1.0e-2 * top, top], rng.uniform(1.0e-2 * top, top, n_c - 2)]))
        it = 1 if k % 2 == 0 else 2
        p_c, c_c = _tabulated_pdf(rng, e_c, it)
        w_d = rng.uniform(0.02, 0.1, n_d)  # probability of each discrete line
        e_d = np.sort(rng.uniform(0.05 * top, 0.5 * top, n_d))
        scale = 1.0 - w_d.sum()
        p = np.concatenate([w_d, p_c * scale])
        # CTTable cdf (OMC distribution_energy.cpp:115-141): running sum over the lines, then the continuous part
        c = np.concatenate([np.cumsum(w_d), w_d.sum() + c_c * scale])
        c[-1] = 1.0
        xs.append(np.concatenate([e_d, e_c])); ps.append(p); cs.append(c); interp.append(it); nd.append(n_d)
        off.append(off[-1] + n_d + n_c)
    return Tabular2D(energy, np.array(off, np.int32), np.array(interp, np.int32), np.concatenate(xs), np.concatenate(ps),
                     np.concatenate(cs), np.array(nd, np.int32), False)


def make_point_angles(rng, n_points):
    """One angular Tabular per point of an outgoing-energy table (the "mu" dataset of a correlated distribution)."""
    off, interp, xs, ps, cs = [0], [], [], [], []
    for j in range(n_points):
        n = int(rng.integers(3, 9))
        mu = np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)]))
        it = 1 if j % 3 == 0 else 2
        p, c = _tabulated_pdf(rng, mu, it)
        xs.append(mu); ps.append(p); cs.append(c); interp.append(it); off.append(off[-1] + n)
    return (np.array(off, np.int32), np.array(interp, np.int32), np.concatenate(xs), np.concatenate(ps), np.concatenate(cs))


@dataclass
class Material:
    name: str
    nuclides: list   # indices into the library
    densities: list  # atom / b-cm
    isotropic: list = None  # per-nuclide isotropic-in-lab flags (<isotropic> of materials.xml), None = none
    sab: list = None        # S(a,b) tables: (position in `nuclides`, table index, fraction), sorted by position


@dataclass
class ThermalTable:
    """One temperature of an S(alpha, beta) table (mcw_thermal_desc; OMC thermal.cpp ThermalData)."""
    name: str
    energy_max: float
    inelastic_type: int            # 0 continuous (IncoherentInelasticAE), 1 discrete (IncoherentInelasticAEDiscrete)
    inelastic_xs_energy: np.ndarray
    inelastic_xs: np.ndarray
    n_mu: int
    # continuous
    inelastic_energy: np.ndarray = None
    inelastic_offset: np.ndarray = None
    inelastic_e_out: np.ndarray = None   # continuous: per point; discrete: (n_xs, n_e_out)
    inelastic_pdf: np.ndarray = None
    inelastic_cdf: np.ndarray = None
    inelastic_mu: np.ndarray = None      # continuous: (points, n_mu); discrete: (n_xs, n_e_out, n_mu)
    n_e_out: int = 0
    skewed: bool = False
    elastic_type: int = 0          # 0 none, 1 coherent (Bragg edges), 2 incoherent discrete
    elastic_x: np.ndarray = None
    elastic_y: np.ndarray = None
    elastic_mu: np.ndarray = None  # (n_elastic, n_elastic_mu)
    kT: float = KT_ROOM            # the temperature of this set
    more_temperatures: list = None  # further temperatures of the table (ThermalTable objects; OMC ThermalScattering::data_[t])


def make_thermal_table(rng, name, kind="continuous", elastic="none", energy_max=4.0, n_mu=8) -> ThermalTable:
    """A synthetic S(a,b) table of the shape the ENDF/B thermal files have: an inelastic cross section that falls from
    the bound to the free value over [1e-5, energy_max] eV, and either continuous outgoing-energy distributions with
    n_mu equiprobable cosines per outgoing energy or equiprobable discrete (energy, cosine) bins; optionally Bragg edges
    (coherent elastic) or a tabulated elastic cross section with discrete cosines."""
    n_xs = 24
    ex = np.exp(np.linspace(np.log(1.0e-5), np.log(energy_max), n_xs))
    ixs = 20.0 + 60.0 / (1.0 + (ex / 0.05) ** 0.8) * rng.uniform(0.9, 1.1, n_xs)
    t = ThermalTable(name, float(energy_max), 0 if kind == "continuous" else 1, ex, ixs, n_mu)

    def cosines(n):
        m = np.sort(rng.uniform(-1.0, 1.0, (n, n_mu)), axis=1)
        return m
    if kind == "continuous":
        n_in = 9
        ein = np.exp(np.linspace(np.log(1.0e-5), np.log(energy_max), n_in))
        off, eo, pd, cd = [0], [], [], []
        for l in range(n_in):
            n = int(rng.integers(5, 14))
            top = ein[l] + 0.25 * rng.uniform(0.5, 1.5)   # up-scatter by about kT-ish amounts
            e = np.sort(np.concatenate([[1.0e-6 * top, top], rng.uniform(1.0e-6 * top, top, n - 2)]))
            p, c = _tabulated_pdf(rng, e, 2)
            eo.append(e); pd.append(p); cd.append(c); off.append(off[-1] + n)
        t.inelastic_energy = ein
        t.inelastic_offset = np.array(off, np.int32)
        t.inelastic_e_out, t.inelastic_pdf, t.inelastic_cdf = np.concatenate(eo), np.concatenate(pd), np.concatenate(cd)
        t.inelastic_mu = cosines(off[-1])
    else:
        t.n_e_out = 12
        t.skewed = kind == "discrete_skewed"
        t.inelastic_e_out = np.sort(ex[:, None] * rng.uniform(0.05, 1.6, (n_xs, t.n_e_out)) + 1.0e-4, axis=1)
        t.inelastic_mu = cosines(n_xs * t.n_e_out).reshape(n_xs, t.n_e_out, n_mu)
    if elastic == "coherent":
        t.elastic_type = 1
        # Bragg edges reach beyond energy_max, as in the evaluated files: CoherentElasticAE::sample reads the structure
        # factor of the edge ABOVE the incident energy (secondary_thermal.cpp:46-48)
        t.elastic_x = np.concatenate([np.sort(np.exp(rng.uniform(np.log(1.8e-3), np.log(0.9 * energy_max), 13))),
                                      [1.5 * energy_max]])
        t.elastic_y = np.cumsum(rng.uniform(0.002, 0.02, 14))
    elif elastic == "incoherent_discrete":
        t.elastic_type = 2
        t.elastic_x = np.exp(np.linspace(np.log(1.0e-5), np.log(energy_max), 11))
        t.elastic_y = 5.0 / (1.0 + t.elastic_x / 0.02) + 0.01
        t.elastic_mu = np.sort(rng.uniform(-1.0, 1.0, (11, 6)), axis=1)
    return t


@dataclass
class UrrTable:
    """Unresolved-resonance probability tables of one nuclide (mcw_urr_desc; OMC urr.cpp UrrData)."""
    nuclide: int
    energy: np.ndarray          # (n_energy,)
    table: np.ndarray           # (n_energy, 6, n_band): cum. probability, total, elastic, fission, (n,gamma), heating
    inelastic_flag: int = -1
    absorption_flag: int = -1
    multiply_smooth: bool = False
    interpolation: int = 2


def add_urr_tables(lib: "Library", seed=9, e_lo=2.0e3, e_hi=1.5e5):
    """Probability tables for every nuclide heavier than A = 80 (the resonance absorbers and structural nuclides):
    16 bands on 8-13 energies in [e_lo, e_hi], alternately as absolute cross sections and as factors on the smooth
    ones (multiply_smooth), with inelastic competition where the nuclide has an inelastic reaction."""
    rng = np.random.Generator(np.random.PCG64([seed, 909]))
    lib.urr_tables = []
    for i, n in enumerate(lib.nuclides):
        if n.awr < 80:
            continue
        n_e, n_b = int(rng.integers(8, 14)), 16
        lo = e_lo * rng.uniform(0.8, 1.5)
        energy = np.exp(np.linspace(np.log(lo), np.log(e_hi * rng.uniform(0.7, 1.3)), n_e))
        t = np.zeros((n_e, 6, n_b))
        smooth = len(lib.urr_tables) % 2 == 1
        for k in range(n_e):
            w = rng.uniform(0.2, 1.0, n_b)
            c = np.cumsum(w / w.sum())
            c[-1] = 1.0
            t[k, 0] = c
            if smooth:
                t[k, 2], t[k, 3], t[k, 4] = rng.lognormal(0, 0.3, n_b), rng.lognormal(0, 0.5, n_b), rng.lognormal(0, 0.8, n_b)
            else:
                t[k, 2] = 9.0 * rng.lognormal(0, 0.3, n_b)
                t[k, 3] = (2.0 * rng.lognormal(0, 0.5, n_b)) if n.fissionable else 0.0
                t[k, 4] = 0.8 * rng.lognormal(0, 0.8, n_b)
            t[k, 1] = t[k, 2] + t[k, 3] + t[k, 4]
            t[k, 5] = 1.0
        lib.urr_tables.append(UrrTable(i, energy, t, 51 if n.inelastic.size else -1, -1, smooth))
    return lib


def add_thermal_scattering(lib: "Library", seed=5):
    """Gives a library S(a,b) tables the way the BEAVRS deck has them (materials.xml <sab name="c_H_in_H2O"/>): hydrogen
    in every material that holds H1 is bound (continuous inelastic, no elastic); a second nuclide, where a material has
    one with A > 10, gets a table with coherent elastic + discrete inelastic at fraction 0.75 (the partial-binding
    branch of calculate_sab_xs), a third one incoherent-discrete elastic + skewed discrete inelastic."""
    rng = np.random.Generator(np.random.PCG64([seed, 4242]))
    lib.thermal_tables = [make_thermal_table(rng, "c_H_in_H2O", "continuous", "none", 4.0),
                          make_thermal_table(rng, "c_Graphite_like", "discrete", "coherent", 2.5),
                          make_thermal_table(rng, "c_ZrH_like", "discrete_skewed", "incoherent_discrete", 1.9)]
    for m in lib.materials:
        sab, used = [], 1
        for pos, i in enumerate(m.nuclides):
            n = lib.nuclides[i]
            if n.name == "H1":
                sab.append((pos, 0, 1.0))
            elif n.awr > 10 and not n.fissionable and used < 3:
                sab.append((pos, used, 0.75 if used == 1 else 1.0))
                used += 1
        m.sab = sorted(sab) if sab else None
    return lib


def add_thermal_temperatures(lib: "Library", temperatures=(400.0, 600.0), seed=17) -> "Library":
    """Gives every S(a,b) table further temperatures (tables of the same kinds drawn anew, same energy_max): the history's
    temperature picks the closer of the two that bound it (OMC thermal.cpp:149-171)."""
    rng = np.random.Generator(np.random.PCG64([seed, 77]))
    kinds = {(0, 0): ("continuous", "none"), (1, 1): ("discrete", "coherent"), (1, 2): ("discrete_skewed", "incoherent_discrete")}
    for t in lib.thermal_tables:
        t.more_temperatures = []
        kind, el = kinds[(t.inelastic_type, t.elastic_type)]
        if t.inelastic_type == 1 and not t.skewed:
            kind = "discrete"
        for T in temperatures:
            m = make_thermal_table(rng, t.name, kind, el, t.energy_max, t.n_mu)
            m.kT = 8.617333262e-5 * float(T)
            t.more_temperatures.append(m)
    return lib


@dataclass
class Library:
    nuclides: list
    materials: list
    n_log_bins: int = 8000
    thermal_tables: list = None
    urr_tables: list = None

    def total_points(self):
        return int(sum(n.energy.size for n in self.nuclides))

    def thermal_entries(self):
        """(tables, same_table_as, kT): every temperature of every S(a,b) table as one entry -- the tables themselves
        first (what materials name), then their further temperatures in table order."""
        tables = list(self.thermal_tables or [])
        same, kT = list(range(len(tables))), [t.kT for t in tables]
        for k, t in enumerate(self.thermal_tables or []):
            for m in (getattr(t, "more_temperatures", None) or []):
                tables.append(m)
                same.append(k)
                kT.append(m.kT)
        return tables, np.array(same, np.int32), np.array(kT, np.float64)

    def temperature_sets(self):
        """[(nuclide index, data set)] of the further temperatures, in library-entry order: set k is entry
        len(nuclides) + k (what mcw_add_temperature_set returns, and what an UrrTable of that set names)."""
        return [(i, m) for i, n in enumerate(self.nuclides) for m in (getattr(n, "more_temperatures", None) or [])]

    def bytes(self):
        return int(sum(n.energy.nbytes + n.xs.nbytes + n.elastic.nbytes + n.inelastic.nbytes for n in self.nuclides))


def _grid(rng, G, res_E, res_G):
    """Strictly increasing grid with `G` points: log-uniform background + clusters at resonances."""
    n_res_pts = min(G // 2, 12 * len(res_E))
    pts = [np.exp(rng.uniform(np.log(E_MIN), np.log(E_MAX), max(G - 2 - n_res_pts, 0)))]
    if len(res_E) and n_res_pts:
        k = rng.integers(0, len(res_E), n_res_pts)
        pts.append(res_E[k] + res_G[k] * rng.standard_cauchy(n_res_pts) * 0.5)
    e = np.concatenate(pts)
    e = e[(e > E_MIN * 1.0001) & (e < E_MAX * 0.9999)]
    e = np.unique(np.concatenate([[E_MIN, E_MAX], e]))
    while e.size < G:  # top up (collisions / clipped points)
        extra = np.exp(rng.uniform(np.log(E_MIN), np.log(E_MAX), G - e.size))
        e = np.unique(np.concatenate([e, extra]))
    if e.size > G:
        keep = np.sort(rng.choice(np.arange(1, e.size - 1), G - 2, replace=False))
        e = np.concatenate([[E_MIN], e[keep], [E_MAX]])
    assert e.size == G and np.all(np.diff(e) > 0)
    return e


def make_nuclide(name, awr, G, rng, sig_el=5.0, cap0=0.1, fis0=0.0, fast_fis=0.0, n_res=0,
                 res_scale=100.0, inel=0.0, inel_E=1.0e6, nu0=2.43, nu1=1.3e-7, q=193.0e6) -> Nuclide:
    res_E = np.exp(rng.uniform(np.log(0.5), np.log(2.0e4), n_res)) if n_res else np.zeros(0)
    res_G = res_E * rng.uniform(0.002, 0.03, n_res) if n_res else np.zeros(0)
    res_H = res_scale * rng.uniform(0.2, 3.0, n_res) / np.sqrt(np.maximum(res_E, 1e-3)) if n_res else np.zeros(0)
    e = _grid(rng, G, res_E, res_G)
    inv_v = np.sqrt(0.0253 / e)
    lor = np.zeros_like(e)
    for E0, Gm, H in zip(res_E, res_G, res_H):
        lor += H / (1.0 + ((e - E0) / (0.5 * Gm)) ** 2)
    fast = 1.0 / (1.0 + (e / 3.0e6) ** 2)
    elastic = sig_el * (0.25 + 0.75 * fast) + 0.3 * lor
    capture = cap0 * inv_v + (0.6 if fis0 > 0 else 1.0) * lor * (cap0 > 0)
    fission = np.zeros_like(e)
    fissionable = fis0 > 0 or fast_fis > 0
    if fissionable:
        fission = fis0 * inv_v + 0.8 * lor * (fis0 > 0) + fast_fis / (1.0 + np.exp(-(e - 1.2e6) / 1.5e5))
    thr = 0
    inelastic = np.zeros(0)
    mass_ratio = (awr / (awr + 1.0)) ** 2
    thr_E = 0.0
    if inel > 0:
        thr = int(np.searchsorted(e, inel_E))
        thr = min(max(thr, 1), G - 2)
        thr_E = e[thr] * 0.999  # LevelInelastic threshold_: just below the first tabulated point
        inelastic = inel * (1.0 - np.exp(-(e[thr:] - e[thr]) / (2.0 * inel_E)))
    inel_full = np.zeros_like(e)
    if inel > 0:
        inel_full[thr:] = inelastic
    # create_derived: total += xs per reaction in MT order (2, 18, 51, 102); absorption = fission + capture
    total = elastic + fission + inel_full + capture
    absorption = fission + capture
    nu = nu1 * e + nu0
    xs = np.zeros((G, 5))
    xs[:, 0] = total
    xs[:, 1] = absorption
    xs[:, 2] = fission
    xs[:, 3] = nu * fission if fissionable else 0.0
    dy = np.zeros(MAX_DELAYED_GROUPS)
    dth = np.zeros(MAX_DELAYED_GROUPS)
    npre = 0
    if fissionable:
        npre = 6
        dy[:6] = 0.0065 * nu0 * np.array([0.033, 0.219, 0.196, 0.395, 0.115, 0.042])
        dth[:6] = np.array([0.25e6, 0.46e6, 0.41e6, 0.45e6, 0.41e6, 0.35e6])
    return Nuclide(name, float(awr), e, xs, elastic, inelastic, thr, float(thr_E), float(mass_ratio),
                   bool(fissionable), (float(nu0), float(nu1)), dy, dth, npre, 0.988e6, 2.249e-6,
                   float(q if fissionable else 0.0))


_BASE = [
    # name, awr, kwargs
    ("H1", 0.9992, dict(sig_el=20.4, cap0=0.332)),
    ("O16", 15.858, dict(sig_el=3.9, cap0=1.9e-4, inel=0.4, inel_E=6.0e6)),
    ("U235", 233.025, dict(sig_el=14.0, cap0=98.7, fis0=585.0, fast_fis=1.2, n_res=40, inel=1.5, inel_E=4.0e4)),
    ("U238", 236.006, dict(sig_el=9.3, cap0=2.68, fast_fis=0.55, n_res=60, res_scale=600.0, inel=2.5, inel_E=5.0e4,
                           nu0=2.49, nu1=1.4e-7)),
    ("Pu239", 236.999, dict(sig_el=8.0, cap0=270.0, fis0=748.0, fast_fis=1.8, n_res=40, inel=1.2, inel_E=3.0e4,
                            nu0=2.88, nu1=1.4e-7, q=198.0e6)),
    ("Zr90", 89.132, dict(sig_el=6.2, cap0=0.011, n_res=10, res_scale=8.0, inel=1.0, inel_E=1.8e6)),
    ("He4", 3.968, dict(sig_el=0.76, cap0=0.0)),
    ("B10", 9.927, dict(sig_el=2.1, cap0=3840.0)),
    ("B11", 10.915, dict(sig_el=4.8, cap0=0.0055)),
]


def make_library(n_nuclides=9, G=2000, seed=1, n_log_bins=8000, tabular=False) -> Library:
    """First 9 nuclides are the named ones above; the rest are generic structural / fission-product
    like nuclides so the BEAVRS shape (91 nuclides) can be reproduced.  tabular = True gives the nuclides
    tabulated secondary distributions (drawn from a separate stream, so the cross sections are the same):
    angular distributions of the elastic reaction for every second nuclide and of the inelastic one for every
    third, and a ContinuousTabular law (alternately in the CM and the LAB system) instead of the level law for
    every second nuclide that has an inelastic reaction."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nucs = []
    for i in range(n_nuclides):
        if i < len(_BASE):
            name, awr, kw = _BASE[i]
        else:
            awr = float(rng.uniform(20, 240))
            name = f"X{i}"
            kw = dict(sig_el=float(rng.uniform(3, 12)), cap0=float(np.exp(rng.uniform(np.log(0.01), np.log(50)))),
                      n_res=int(rng.integers(0, 30)), res_scale=float(rng.uniform(5, 200)),
                      inel=float(rng.uniform(0, 2)), inel_E=float(np.exp(rng.uniform(np.log(5e4), np.log(3e6)))))
        g = G if np.isscalar(G) else int(G[i])
        nucs.append(make_nuclide(name, awr, g, rng, **kw))
    if tabular:
        trng = np.random.Generator(np.random.PCG64([seed, 77]))
        k_inel = 0
        for i, n in enumerate(nucs):
            if i % 2 == 0:
                n.elastic_angle = make_angle_table(trng, 1.0e2 * (1 + i), 1.5e7)
            if n.inelastic.size:
                if i % 3 == 0:
                    n.inelastic_angle = make_angle_table(trng, n.inel_threshold_energy, 1.8e7, 4)
                if k_inel % 2 == 0:
                    n.inelastic_energy = make_continuous_table(trng, n.inel_threshold_energy, 1.9e7, discrete=k_inel % 4 == 0)
                    n.inelastic_in_cm = k_inel % 4 != 0
                k_inel += 1
        # product distributions that give energy and cosine together (drawn after everything above, so the library
        # without them is unchanged): the inelastic reactions WITHOUT a table so far cycle through correlated,
        # Kalbach-Mann and N-body
        k_law = 0
        for n in nucs:
            if not n.inelastic.size or n.inelastic_energy is not None:
                continue
            law = (2, 3, 4)[k_law % 3]
            k_law += 1
            n.inelastic_law = law
            n.inelastic_angle = None
            if law == 4:
                n.nbody = (3 + (k_law // 3) % 3, float(trng.uniform(2.5, 4.0)), -0.5 * n.inel_threshold_energy)
                n.inelastic_in_cm = True
                continue
            n.inelastic_energy = make_continuous_table(trng, n.inel_threshold_energy, 1.9e7, discrete=False)
            n.inelastic_in_cm = k_law % 2 == 0
            npts = int(n.inelastic_energy.offset[-1])
            if law == 2:
                n.corr_angle = make_point_angles(trng, npts)
            else:
                n.km_r = trng.uniform(0.0, 1.0, npts)
                n.km_a = trng.uniform(0.3, 3.0, npts)
    mats = standard_materials(n_nuclides)
    return Library(nucs, mats, n_log_bins)


def library_for_names(names, G=2000, seed=1, n_log_bins=8000) -> Library:
    """One synthetic nuclide per name (e.g. every nuclide of a materials.xml), so an unmodified OpenMC
    deck can run without ENDF/B data: the named base nuclides keep their parameters, any other name gets
    parameters drawn from a stream seeded by the name (mass number from its digits, fissile / fertile
    actinides by element).  The shapes are plausible, the values are NOT evaluated nuclear data."""
    import re
    import zlib
    base = {b[0]: b for b in _BASE}
    nucs = []
    for name in names:
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        if name in base:
            _, awr, kw = base[name]
        else:
            m = re.match(r"([A-Za-z]+)(\d+)", name)
            a = int(m.group(2)) if m else 100
            el = m.group(1) if m else name
            awr = 0.9914 * a if a > 1 else 0.9992
            kw = dict(sig_el=float(rng.uniform(3, 12)), cap0=float(np.exp(rng.uniform(np.log(0.01), np.log(50)))),
                      n_res=int(rng.integers(0, 30)) if a > 20 else 0, res_scale=float(rng.uniform(5, 200)),
                      inel=float(rng.uniform(0, 2)) if a > 10 else 0.0,
                      inel_E=float(np.exp(rng.uniform(np.log(5e4), np.log(3e6)))))
            if el in ("U", "Pu", "Th", "Np", "Am", "Cm"):
                fissile = a % 2 == 1
                kw.update(fis0=float(rng.uniform(300, 800)) if fissile else 0.0, fast_fis=float(rng.uniform(0.4, 2.0)),
                          n_res=40, nu0=float(rng.uniform(2.3, 2.9)))
        nucs.append(make_nuclide(name, awr, G, rng, **kw))
    return Library(nucs, [], n_log_bins)


def add_reactions(lib: Library, seed=11) -> Library:
    """Gives every nuclide that has an inelastic reaction a realistic LIST of scattering reactions after it (MT order, as
    OpenMC's index_inelastic_scatter_ has them): two more levels MT 52 / 53 (level law; the second with an angular
    distribution), the continuum MT 91 (continuous tabular in the LAB system, or Kalbach-Mann for every second nuclide,
    with a NON-integral multiplicity 1.25 for every third nuclide: the weight changes) and (n,2n) as MT 875 (N-body for
    every second nuclide, else correlated; multiplicity 2: the second neutron is what MaCaw drops).  The total cross
    section grows by the new reactions, summed in the order Nuclide::create_derived adds them (MT 2, 18, 51, 52, 53, 91,
    102, 875)."""
    rng = np.random.Generator(np.random.PCG64([seed, 5]))
    k = 0
    for n in lib.nuclides:
        if not n.inelastic.size:
            continue
        e, G, thr0 = n.energy, n.energy.size, n.inel_threshold
        mr = (n.awr / (n.awr + 1.0)) ** 2

        def xs_from(thr, height, width):
            return height * (1.0 - np.exp(-(e[thr:] - e[thr]) / width))

        def later(f):  # a threshold above the first level's
            return int(min(max(np.searchsorted(e, e[thr0] * f), thr0 + 1), G - 3))
        rxs = []
        for mt, f, h in ((52, 1.6, 0.5), (53, 2.7, 0.35)):
            t = later(f)
            r = Reaction(mt, t, xs_from(t, h * float(rng.uniform(0.5, 1.5)), 2.0 * e[t]), float(e[t] * 0.999), mr)
            if mt == 53:
                r.angle = make_angle_table(rng, e[t], max(1.9e7, 1.2 * e[t]))
            rxs.append(r)
        t = later(4.0)
        r91 = Reaction(91, t, xs_from(t, 0.8 * float(rng.uniform(0.5, 1.5)), 3.0 * e[t]), law=1, in_cm=False,
                       yield_=1.25 if k % 3 == 0 else 1.0)
        top = max(1.9e7, 1.2 * e[t])  # the tables reach from the threshold to above it, wherever the threshold lies
        r91.energy = make_continuous_table(rng, e[t], top, discrete=k % 2 == 0)
        if k % 2:
            r91.law, r91.in_cm = 3, True
            r91.energy = make_continuous_table(rng, e[t], top, discrete=False)
            npts = int(r91.energy.offset[-1])
            r91.km_r, r91.km_a = rng.uniform(0.0, 1.0, npts), rng.uniform(0.3, 3.0, npts)
        rxs.append(r91)
        t = later(9.0)
        r2n = Reaction(875, t, xs_from(t, 0.6 * float(rng.uniform(0.5, 1.5)), 2.0 * e[t]), yield_=2.0)
        if k % 2:
            r2n.law, r2n.nbody = 4, (3 + (k // 2) % 3, float(rng.uniform(2.5, 4.0)), -0.5 * float(e[t]))
        else:
            r2n.law = 2
            r2n.energy = make_continuous_table(rng, e[t], max(1.9e7, 1.2 * e[t]), discrete=False)
            r2n.corr_angle = make_point_angles(rng, int(r2n.energy.offset[-1]))
        rxs.append(r2n)
        n.more_reactions = rxs
        # total in create_derived order; capture = absorption - fission (exact for the non-fissionable nuclides)
        fission, capture = n.xs[:, 2].copy(), n.xs[:, 1] - n.xs[:, 2]
        full = lambda thr, v: np.concatenate([np.zeros(thr), v])
        total = n.elastic + fission + full(thr0, n.inelastic)
        for r in rxs[:3]:
            total = total + full(r.threshold, r.xs)
        total = total + capture
        total = total + full(r2n.threshold, r2n.xs)
        n.xs = n.xs.copy()
        n.xs[:, 0] = total
        k += 1
    return lib


def add_temperatures(lib: Library, temperatures=(600.0, 900.0), seed=21, skip_every=4) -> Library:
    """Gives the nuclides further data sets at other temperatures (all but every `skip_every`-th, which stay
    single-temperature): each set has its OWN energy grid (a thinned and jittered copy, as Doppler-broadened evaluations
    have) and every reaction's cross section re-tabulated on it with a smooth temperature-dependent factor; thresholds
    are re-indexed on the new grid, the secondary distributions are the nuclide's (they do not depend on temperature in
    the data format).  Nuclides with probability tables get tables for every set (OpenMC reads `urr/<T>K` for each
    temperature it loads); those are addressed by the library ENTRY of the set: entries [0, n) are the nuclides, further
    sets follow in nuclide order (Library.entries())."""
    import copy
    rng = np.random.Generator(np.random.PCG64([seed, 31]))
    n_nuc = len(lib.nuclides)
    entry = n_nuc
    urr_of = {t.nuclide: t for t in (getattr(lib, "urr_tables", None) or [])}
    for i, n in enumerate(lib.nuclides):
        if skip_every and i % skip_every == skip_every - 1:
            continue
        n.more_temperatures = []
        e0, G = n.energy, n.energy.size
        thr_e = e0[n.inel_threshold] if n.inelastic.size else 0.0
        for T in temperatures:
            kT = 8.617333262e-5 * float(T)
            keep = np.sort(rng.choice(np.arange(1, G - 1), int(0.9 * (G - 2)), replace=False))
            mid = e0[keep] * (1.0 + 1e-4 * rng.uniform(-1, 1, keep.size))
            e = np.unique(np.concatenate([[e0[0] * 1.5], mid[(mid > e0[0] * 1.5) & (mid < e0[-1] * 0.97)], [e0[-1] * 0.97]]))
            fac = 1.0 + 0.08 * (T / 293.6 - 1.0) * np.sin(np.log(e))

            def on(x, thr):  # a reaction's values (from its threshold on) re-tabulated on the new grid
                full = np.concatenate([np.zeros(thr), x])
                return np.interp(e, e0, full) * fac
            m = copy.copy(n)
            m.more_temperatures, m.kT, m.energy = None, kT, e
            m.elastic = on(n.elastic, 0)
            fission = on(n.xs[:, 2], 0)
            capture = on(n.xs[:, 1] - n.xs[:, 2], 0)
            total = m.elastic + fission
            if n.inelastic.size:
                t = int(min(np.searchsorted(e, thr_e), e.size - 3))
                m.inel_threshold = t
                m.inelastic = on(n.inelastic, n.inel_threshold)[t:]
                total = total + np.concatenate([np.zeros(t), m.inelastic])
            if n.more_reactions:
                m.more_reactions = []
                for r in n.more_reactions:
                    r2 = copy.copy(r)
                    t = int(min(np.searchsorted(e, e0[r.threshold]), e.size - 3))
                    r2.threshold, r2.xs = t, on(r.xs, r.threshold)[t:]
                    m.more_reactions.append(r2)
            # total in create_derived order: MT 2, 18, 51 ... 91, 102, then the reactions filed after capture (875)
            for r2 in (m.more_reactions or []):
                if r2.mt < 102:
                    total = total + np.concatenate([np.zeros(r2.threshold), r2.xs])
            total = total + capture
            for r2 in (m.more_reactions or []):
                if r2.mt > 102:
                    total = total + np.concatenate([np.zeros(r2.threshold), r2.xs])
            xs = np.zeros((e.size, 5))
            xs[:, 0], xs[:, 1], xs[:, 2] = total, fission + capture, fission
            if n.fissionable and getattr(n, "fission_products", None):  # create_derived: nu(E, total) * fission
                xs[:, 3] = (n.total_nu or n.fission_products[0].yield_)(e) * fission
            else:
                xs[:, 3] = (n.nu_poly[1] * e + n.nu_poly[0]) * fission if n.fissionable else 0.0
            m.xs = xs
            n.more_temperatures.append(m)
            if i in urr_of:
                u = copy.copy(urr_of[i])
                u.nuclide = entry
                u.table = u.table.copy()
                u.table[:, 2:5] *= 1.0 + 0.05 * (T / 293.6 - 1.0)
                u.table[:, 1] = u.table[:, 2] + u.table[:, 3] + u.table[:, 4]
                lib.urr_tables.append(u)
            entry += 1
    return lib


def add_evaporation_laws(lib: Library) -> Library:
    """After add_reactions: evaporation-type energy laws and energy-dependent yields on scattering reactions.  Of every
    nuclide with a reaction list, MT 53 becomes an evaporation spectrum (keeping its angular distribution), the continuum
    MT 91 of every second one a Maxwell spectrum with a tabulated multiplicity rising from 1 to 1.6 (not integral: the
    weight follows it), (n,2n) MT 875 of the others a Watt spectrum; the first reaction (inline, MT 51) of every third
    nuclide whose law is level or continuous tabular becomes an evaporation spectrum in the LAB system with a polynomial
    multiplicity.  Spectrum parameters grow with the incoming energy (tabulated, lin-lin), restriction energies lie below
    the thresholds."""
    k = 0
    for n in lib.nuclides:
        if not n.more_reactions:
            continue
        e = n.energy
        def theta(t, f):
            lo = float(e[t])
            return Function1D(2, np.array([f * lo, f * 2.0e7]), np.array([lo, 2.0e7]), 2)
        r52, r53, r91, r2n = n.more_reactions
        r53.law, r53.in_cm, r53.u, r53.a = 6, True, 0.5 * float(e[r53.threshold]), theta(r53.threshold, 0.2)
        if k % 2 == 0:
            t = r91.threshold
            r91.law, r91.in_cm, r91.energy, r91.u, r91.a = 5, False, None, 0.4 * float(e[t]), theta(t, 0.15)
            r91.yield_fn = Function1D(2, np.array([1.0, 1.3, 1.6]), np.array([float(e[t]), 0.5 * (float(e[t]) + 2.0e7), 2.0e7]), 2)
        else:
            t = r2n.threshold
            r2n.law, r2n.in_cm, r2n.nbody, r2n.u = 7, True, None, -2.0e7
            r2n.a, r2n.b = theta(t, 0.05), Function1D(2, np.array([2.0e-6, 3.0e-6]), np.array([float(e[t]), 2.0e7]), 2)
        if k % 3 == 0 and n.inelastic_law == -1:
            t = n.inel_threshold
            n.inelastic_law, n.inelastic_in_cm, n.inelastic_angle, n.inelastic_energy = 6, False, None, None
            n.inelastic_u, n.inelastic_a = 0.5 * float(e[t]), theta(t, 0.25)
            n.inelastic_yield_fn = Function1D(1, np.array([1.0, 2.0e-9]))
        k += 1
    return lib


def add_fission_data(lib: Library, seed=41) -> Library:
    """Gives the fissionable nuclides fission-neutron data as evaluated files have them instead of the scalar description:
    a tabulated total nu (lin-lin), a prompt yield (tabulated, or a polynomial for every second nuclide), six delayed
    groups with energy-dependent yields (one of them a histogram table), and product spectra of every kind -- prompt: Watt
    with tabulated a(E), b(E) or a continuous tabular law; delayed: Maxwell, evaporation and continuous tabular.  The
    nu-fission column of the derived table becomes total_nu(E) * fission as Nuclide::create_derived forms it."""
    rng = np.random.Generator(np.random.PCG64([seed, 8]))
    k = 0
    for n in lib.nuclides:
        if not n.fissionable:
            continue
        eg = np.array([1.0e-5, 1.0, 1.0e3, 1.0e5, 1.0e6, 4.0e6, 1.0e7, 2.0e7])
        nu_t = n.nu_poly[0] + n.nu_poly[1] * eg + 0.02 * np.sin(np.log(eg))
        total = Function1D(2, nu_t, eg, 2)
        frac = 0.06 * np.array([0.1, 0.2, 0.2, 0.25, 0.15, 0.1])  # 6 % delayed: every group and law is sampled often
        prods = []
        dsum = np.zeros_like(eg)
        for g in range(6):
            yg = frac[g] * nu_t[0] * (1.0 - 0.3 * eg / 2.0e7) * (1.0 + 0.1 * g * (eg > 1.0e5))
            dsum += yg
            y = Function1D(2, yg, eg, 1 if g == 3 else 2)
            if g % 3 == 0:
                p = FissionProduct(y, 5, -2.0e7, Function1D(2, (0.25e6 + 2.0e4 * g) * (1 + 0.1 * np.log10(eg / 1e-5) / 12), eg, 2))
            elif g % 3 == 1:
                p = FissionProduct(y, 6, -2.0e7, Function1D(2, (0.4e6 + 1.0e4 * g) * (1 + 0.05 * np.log10(eg / 1e-5) / 12), eg, 2))
            else:
                p = FissionProduct(y, 1, energy=make_continuous_table(rng, 1.0e-5, 1.9e7, n_energy=3, discrete=False))
                p.energy.x = np.abs(p.energy.x) * 0.05 + 1.0e3  # delayed neutrons: a few hundred keV
                p.energy = _renormalise(p.energy)
            prods.append(p)
        if k % 2 == 0:
            prompt_y = Function1D(2, nu_t - dsum, eg, 2)
            prompt = FissionProduct(prompt_y, 7, -2.0e7, Function1D(2, n.watt_a * (1 + 0.02 * np.log10(eg / 1e-5)), eg, 2),
                                    Function1D(2, n.watt_b * (1 - 0.01 * np.log10(eg / 1e-5)), eg, 2))
        else:
            prompt_y = Function1D(1, np.array([n.nu_poly[0] - dsum[0], n.nu_poly[1], -1.0e-16]))
            tab = make_continuous_table(rng, 1.0e-5, 1.9e7, n_energy=4, discrete=False)
            tab.x = np.abs(tab.x) * 0.3 + 1.0e4
            prompt = FissionProduct(prompt_y, 1, energy=_renormalise(tab))
        n.fission_products = [prompt] + prods
        n.total_nu = total if k % 3 != 2 else None   # every third nuclide: no total nu, the prompt yield stands in
        n.n_precursor = 6
        nu_at = (n.total_nu or prompt_y)(n.energy)
        n.xs = n.xs.copy()
        n.xs[:, 3] = nu_at * n.xs[:, 2]
        for m in (getattr(n, "more_temperatures", None) or []):
            m.fission_products, m.total_nu, m.n_precursor = n.fission_products, n.total_nu, 6
            m.xs = m.xs.copy()
            m.xs[:, 3] = (n.total_nu or prompt_y)(m.energy) * m.xs[:, 2]
        k += 1
    return lib


def _renormalise(t: Tabular2D) -> Tabular2D:
    """pdf / cdf of every distribution recomputed for (possibly rescaled, re-sorted) abscissae, cdf ending at exactly 1"""
    x, p, c = t.x.copy(), t.pdf.copy(), t.cdf.copy()
    for k in range(len(t.energy)):
        a, b = int(t.offset[k]), int(t.offset[k + 1])
        xs = np.sort(x[a:b])
        for j in range(1, xs.size):  # strictly increasing
            if xs[j] <= xs[j - 1]:
                xs[j] = xs[j - 1] * (1 + 1e-9) + 1e-9
        pp = np.abs(p[a:b]) + 1e-30
        if t.interpolation[k] == 1:
            cc = np.concatenate([[0.0], np.cumsum(pp[:-1] * np.diff(xs))])
        else:
            cc = np.concatenate([[0.0], np.cumsum(0.5 * (pp[:-1] + pp[1:]) * np.diff(xs))])
        pp, cc = pp / cc[-1], cc / cc[-1]
        cc[-1] = 1.0
        x[a:b], p[a:b], c[a:b] = xs, pp, cc
    nd = None if t.n_discrete is None else np.zeros_like(t.n_discrete)
    return Tabular2D(t.energy, t.offset, t.interpolation, x, p, c, nd, t.histogram_in_energy)


def standard_materials(n_nuclides) -> list:
    """Material 0 fuel (UO2 3.1 %), 1 gap (He), 2 clad (Zr + generic alloying nuclides), 3 borated water.
    Indices follow `_BASE`; nuclides beyond the named ones are spread over clad and fuel like the
    25-nuclide Zircaloy / fission-product inventories of the BEAVRS materials.xml."""
    H1, O16, U235, U238, Pu239, Zr90, He4, B10, B11 = range(9)
    have = lambda i: i < n_nuclides
    fuel = [(U235, 7.18e-4), (U238, 2.21e-2), (O16, 4.57e-2)]
    gap = [(He4, 2.4e-5)] if have(He4) else [(O16, 1e-6)]
    clad = [(Zr90, 4.3e-2)] if have(Zr90) else [(O16, 4.3e-2)]
    water = [(H1, 6.6e-2), (O16, 3.3e-2)]
    if have(B10):
        water.append((B10, 1.0e-5))
    if have(B11):
        water.append((B11, 4.0e-5))
    extra = list(range(9, n_nuclides))
    for k, i in enumerate(extra):
        (clad if k % 3 else fuel).append((i, 1.0e-4 / (1 + k % 7)))
    mk = lambda name, lst: Material(name, [i for i, _ in lst if have(i)], [d for i, d in lst if have(i)])
    return [mk("fuel", fuel), mk("gap", gap), mk("clad", clad), mk("water", water)]


def infinite_medium_library(G=2000, seed=1) -> Library:
    """C1: one homogeneous material with the six nuclides of examples/infinite_medium/run_openmc_csg.py:32-38."""
    lib = make_library(6, G, seed)
    H1, O16, U235, U238, Pu239, Zr90 = range(6)
    mix = Material("mix", [H1, O16, U235, U238, Pu239, Zr90], [4.0e-2, 3.5e-2, 2.0e-4, 6.0e-3, 1.0e-5, 8.0e-3])
    he = Material("moderator", [H1, O16], [6.6e-2, 3.3e-2])
    lib.materials = [mix, he]
    return lib


# ---- ctypes mirror of include/macaw_b200.h ---------------------------------------------------------
class Tabular2DDesc(C.Structure):
    _fields_ = [("n_energy", C.c_int32), ("histogram_in_energy", C.c_int32), ("energy", C.c_void_p), ("offset", C.c_void_p),
                ("interpolation", C.c_void_p), ("n_discrete", C.c_void_p), ("x", C.c_void_p), ("pdf", C.c_void_p),
                ("cdf", C.c_void_p)]


def _pack_table(d: Tabular2DDesc, t, keep):
    if t is None:
        return
    # what mcw_set_library checks (csrc/mcw_api.cu take_table), so that bad synthetic data fail where there is no GPU too
    assert len(t.energy) >= 2 and np.all(np.diff(t.energy) > 0), "tabulated distribution: incoming energies must increase"
    assert np.all(np.diff(t.offset) >= 2) and set(np.unique(t.interpolation)) <= {1, 2}
    assert np.all(np.asarray(t.cdf)[np.asarray(t.offset[1:]) - 1] == 1.0), "tabulated distribution: the cdf must end at exactly 1"
    arrs = [np.ascontiguousarray(t.energy, np.float64), np.ascontiguousarray(t.offset, np.int32),
            np.ascontiguousarray(t.interpolation, np.int32), np.ascontiguousarray(t.x, np.float64),
            np.ascontiguousarray(t.pdf, np.float64), np.ascontiguousarray(t.cdf, np.float64)]
    keep += arrs
    d.n_energy, d.histogram_in_energy = arrs[0].size, int(t.histogram_in_energy)
    d.energy, d.offset, d.interpolation, d.x, d.pdf, d.cdf = [a.ctypes.data for a in arrs]
    if t.n_discrete is not None:
        nd = np.ascontiguousarray(t.n_discrete, np.int32)
        keep.append(nd)
        d.n_discrete = nd.ctypes.data


class Function1DDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("interpolation", C.c_int32), ("reserved0", C.c_int32),
                ("x", C.c_void_p), ("y", C.c_void_p)]


class FissionProductDesc(C.Structure):
    _fields_ = [("yield_", Function1DDesc), ("law", C.c_int32), ("reserved0", C.c_int32), ("u", C.c_double),
                ("a", Function1DDesc), ("b", Function1DDesc), ("energy", Tabular2DDesc)]


def _pack_function(d, f, keep):
    if f is None:
        return
    y = np.ascontiguousarray(f.y, np.float64)
    keep.append(y)
    d.kind, d.n, d.interpolation, d.y = f.kind, y.size, f.interpolation, y.ctypes.data
    if f.kind == 2:
        x = np.ascontiguousarray(f.x, np.float64)
        assert x.size == y.size and np.all(np.diff(x) > 0)
        keep.append(x)
        d.x = x.ctypes.data


class NuclideDesc(C.Structure):
    _fields_ = [("n_grid", C.c_int32), ("fissionable", C.c_int32), ("awr", C.c_double), ("kT", C.c_double),
                ("energy", C.c_void_p), ("xs", C.c_void_p), ("elastic", C.c_void_p),
                ("n_inelastic", C.c_int32), ("inel_threshold", C.c_int32), ("inelastic", C.c_void_p),
                ("inel_threshold_energy", C.c_double), ("inel_mass_ratio", C.c_double),
                ("nu_poly", C.c_double * 2), ("n_precursor", C.c_int32), ("reserved0", C.c_int32),
                ("delayed_yield", C.c_double * MAX_DELAYED_GROUPS), ("delayed_theta", C.c_double * MAX_DELAYED_GROUPS),
                ("watt_a", C.c_double), ("watt_b", C.c_double), ("q_fission", C.c_double),
                ("elastic_angle", Tabular2DDesc), ("inelastic_angle", Tabular2DDesc), ("inelastic_law", C.c_int32),
                ("inelastic_in_cm", C.c_int32), ("inelastic_energy", Tabular2DDesc),
                ("corr_offset", C.c_void_p), ("corr_interp", C.c_void_p), ("corr_mu", C.c_void_p), ("corr_pdf", C.c_void_p),
                ("corr_cdf", C.c_void_p), ("km_r", C.c_void_p), ("km_a", C.c_void_p), ("nbody_n", C.c_int32),
                ("reserved1", C.c_int32), ("nbody_mass_ratio", C.c_double), ("nbody_q", C.c_double),
                ("inelastic_yield", C.c_double), ("n_more_reactions", C.c_int32), ("reserved2", C.c_int32),
                ("more_reactions", C.c_void_p),
                ("n_fission_products", C.c_int32), ("reserved3", C.c_int32), ("fission_products", C.c_void_p),
                ("total_nu", Function1DDesc),
                ("inelastic_yield_fn", Function1DDesc), ("inelastic_u", C.c_double), ("inelastic_a", Function1DDesc),
                ("inelastic_b", Function1DDesc)]


class ReactionDesc(C.Structure):
    _fields_ = [("mt", C.c_int32), ("threshold", C.c_int32), ("xs", C.c_void_p), ("threshold_energy", C.c_double),
                ("mass_ratio", C.c_double), ("yield_", C.c_double), ("law", C.c_int32), ("in_cm", C.c_int32),
                ("angle", Tabular2DDesc), ("energy", Tabular2DDesc),
                ("corr_offset", C.c_void_p), ("corr_interp", C.c_void_p), ("corr_mu", C.c_void_p), ("corr_pdf", C.c_void_p),
                ("corr_cdf", C.c_void_p), ("km_r", C.c_void_p), ("km_a", C.c_void_p), ("nbody_n", C.c_int32),
                ("reserved0", C.c_int32), ("nbody_mass_ratio", C.c_double), ("nbody_q", C.c_double),
                ("yield_fn", Function1DDesc), ("u", C.c_double), ("a", Function1DDesc), ("b", Function1DDesc)]


class ThermalDesc(C.Structure):
    _fields_ = [("energy_max", C.c_double), ("inelastic_type", C.c_int32), ("elastic_type", C.c_int32),
                ("n_inelastic_xs", C.c_int32), ("n_inelastic_energy", C.c_int32), ("inelastic_xs_energy", C.c_void_p),
                ("inelastic_xs", C.c_void_p), ("inelastic_energy", C.c_void_p), ("inelastic_offset", C.c_void_p),
                ("inelastic_e_out", C.c_void_p), ("inelastic_pdf", C.c_void_p), ("inelastic_cdf", C.c_void_p),
                ("inelastic_mu", C.c_void_p), ("n_mu", C.c_int32), ("n_e_out", C.c_int32), ("skewed", C.c_int32),
                ("n_elastic", C.c_int32), ("elastic_x", C.c_void_p), ("elastic_y", C.c_void_p), ("elastic_mu", C.c_void_p),
                ("n_elastic_mu", C.c_int32), ("reserved0", C.c_int32)]


class UrrDesc(C.Structure):
    _fields_ = [("nuclide", C.c_int32), ("interpolation", C.c_int32), ("inelastic_flag", C.c_int32),
                ("absorption_flag", C.c_int32), ("multiply_smooth", C.c_int32), ("n_energy", C.c_int32),
                ("n_band", C.c_int32), ("reserved0", C.c_int32), ("energy", C.c_void_p), ("table", C.c_void_p)]


def pack_temperature_sets(lib: Library):
    """mcw_nuclide_desc of every further temperature set; returns ([(nuclide index, desc)], keepalive)."""
    out, keep = [], []
    for i, m in lib.temperature_sets():
        nd, _, k = pack_library(Library([m], []))
        keep += [nd, k]
        out.append((i, nd))
    return out, keep


def pack_urr(tables):
    keep = []
    ud = (UrrDesc * max(len(tables), 1))()
    for d, t in zip(ud, tables):
        e, tb = np.ascontiguousarray(t.energy, np.float64), np.ascontiguousarray(t.table, np.float64)
        keep += [e, tb]
        d.nuclide, d.interpolation, d.inelastic_flag, d.absorption_flag = t.nuclide, t.interpolation, t.inelastic_flag, t.absorption_flag
        d.multiply_smooth, d.n_energy, d.n_band = int(t.multiply_smooth), e.size, tb.shape[2]
        d.energy, d.table = e.ctypes.data, tb.ctypes.data
    return ud, keep


def pack_thermal(tables):
    """mcw_thermal_desc array of a list of ThermalTable; returns (descs, keepalive)."""
    keep = []
    td = (ThermalDesc * max(len(tables), 1))()

    def ptr(a, dt=np.float64):
        if a is None:
            return None
        b = np.ascontiguousarray(a, dt)
        keep.append(b)
        return b.ctypes.data
    for d, t in zip(td, tables):
        d.energy_max, d.inelastic_type, d.elastic_type = t.energy_max, t.inelastic_type, t.elastic_type
        d.n_inelastic_xs = int(np.size(t.inelastic_xs_energy))
        d.inelastic_xs_energy, d.inelastic_xs = ptr(t.inelastic_xs_energy), ptr(t.inelastic_xs)
        d.n_inelastic_energy = int(np.size(t.inelastic_energy)) if t.inelastic_energy is not None else 0
        d.inelastic_energy, d.inelastic_offset = ptr(t.inelastic_energy), ptr(t.inelastic_offset, np.int32)
        d.inelastic_e_out, d.inelastic_pdf, d.inelastic_cdf = ptr(t.inelastic_e_out), ptr(t.inelastic_pdf), ptr(t.inelastic_cdf)
        d.inelastic_mu = ptr(t.inelastic_mu)
        d.n_mu, d.n_e_out, d.skewed = t.n_mu, t.n_e_out, int(t.skewed)
        d.n_elastic = int(np.size(t.elastic_x)) if t.elastic_x is not None else 0
        d.elastic_x, d.elastic_y, d.elastic_mu = ptr(t.elastic_x), ptr(t.elastic_y), ptr(t.elastic_mu)
        d.n_elastic_mu = int(t.elastic_mu.shape[1]) if t.elastic_mu is not None else 0
    return td, keep


def _pack_product(d, law, n, keep):
    """data of the product distributions that give energy and cosine together (same field names in both descriptors)"""
    def ptr(a, dt):
        b = np.ascontiguousarray(a, dt)
        keep.append(b)
        return b.ctypes.data
    if law == 2:
        off, interp, mu, pdf, cdf = n.corr_angle
        d.corr_offset, d.corr_interp = ptr(off, np.int32), ptr(interp, np.int32)
        d.corr_mu, d.corr_pdf, d.corr_cdf = ptr(mu, np.float64), ptr(pdf, np.float64), ptr(cdf, np.float64)
    elif law == 3:
        d.km_r, d.km_a = ptr(n.km_r, np.float64), ptr(n.km_a, np.float64)
    elif law == 4:
        d.nbody_n, d.nbody_mass_ratio, d.nbody_q = int(n.nbody[0]), float(n.nbody[1]), float(n.nbody[2])


def pack_inelastic_law(d, n, keep):
    """inelastic_law of the inline reaction, its product distribution, its yield, and the further reactions"""
    law = getattr(n, "inelastic_law", -1)
    d.inelastic_law = law if law >= 0 else (1 if getattr(n, "inelastic_energy", None) is not None else 0)
    d.inelastic_in_cm = int(getattr(n, "inelastic_in_cm", True))
    _pack_product(d, d.inelastic_law, n, keep)
    d.inelastic_yield = float(getattr(n, "inelastic_yield", 1.0))
    _pack_function(d.inelastic_yield_fn, getattr(n, "inelastic_yield_fn", None), keep)
    d.inelastic_u = float(getattr(n, "inelastic_u", 0.0))
    _pack_function(d.inelastic_a, getattr(n, "inelastic_a", None), keep)
    _pack_function(d.inelastic_b, getattr(n, "inelastic_b", None), keep)
    more = getattr(n, "more_reactions", None) or []
    if more:
        rd = (ReactionDesc * len(more))()
        keep.append(rd)
        for r, x in zip(rd, more):
            xs = np.ascontiguousarray(x.xs, np.float64)
            keep.append(xs)
            r.mt, r.threshold, r.xs = x.mt, x.threshold, xs.ctypes.data
            r.threshold_energy, r.mass_ratio, r.yield_ = x.threshold_energy, x.mass_ratio, x.yield_
            r.law, r.in_cm = x.law, int(x.in_cm)
            _pack_table(r.angle, x.angle, keep)
            _pack_table(r.energy, x.energy, keep)
            _pack_product(r, x.law, x, keep)
            _pack_function(r.yield_fn, x.yield_fn, keep)
            r.u = float(x.u)
            _pack_function(r.a, x.a, keep)
            _pack_function(r.b, x.b, keep)
        d.n_more_reactions, d.more_reactions = len(more), C.addressof(rd)
    prods = getattr(n, "fission_products", None) or []
    if prods:  # fission neutrons as a nuclide file has them
        fd = (FissionProductDesc * len(prods))()
        keep.append(fd)
        for f, x in zip(fd, prods):
            _pack_function(f.yield_, x.yield_, keep)
            f.law, f.u = x.law, x.u
            _pack_function(f.a, x.a, keep)
            _pack_function(f.b, x.b, keep)
            _pack_table(f.energy, x.energy, keep)
        d.n_fission_products, d.fission_products = len(prods), C.addressof(fd)
        _pack_function(d.total_nu, getattr(n, "total_nu", None), keep)


class MaterialDesc(C.Structure):
    _fields_ = [("n_nuclides", C.c_int32), ("reserved0", C.c_int32), ("nuclide", C.c_void_p),
                ("atom_density", C.c_void_p), ("isotropic_in_lab", C.c_void_p)]


def pack_library(lib: Library):
    """Build the C descriptor arrays; returns (nuclide_descs, material_descs, keepalive)."""
    keep = []
    nd = (NuclideDesc * len(lib.nuclides))()
    for d, n in zip(nd, lib.nuclides):
        e = np.ascontiguousarray(n.energy, np.float64)
        xs = np.ascontiguousarray(n.xs, np.float64)
        el = np.ascontiguousarray(n.elastic, np.float64)
        inel = np.ascontiguousarray(n.inelastic if n.inelastic.size else np.zeros(1), np.float64)
        keep += [e, xs, el, inel]
        d.n_grid, d.fissionable, d.awr, d.kT = e.size, int(n.fissionable), n.awr, n.kT
        d.energy, d.xs, d.elastic = e.ctypes.data, xs.ctypes.data, el.ctypes.data
        d.n_inelastic = int(n.inelastic.size > 0)
        d.inel_threshold = n.inel_threshold
        d.inelastic = inel.ctypes.data
        d.inel_threshold_energy, d.inel_mass_ratio = n.inel_threshold_energy, n.inel_mass_ratio
        d.nu_poly[0], d.nu_poly[1] = n.nu_poly
        d.n_precursor = n.n_precursor
        for g in range(MAX_DELAYED_GROUPS):
            d.delayed_yield[g] = float(n.delayed_yield[g])
            d.delayed_theta[g] = float(n.delayed_theta[g])
        d.watt_a, d.watt_b, d.q_fission = n.watt_a, n.watt_b, n.q_fission
        _pack_table(d.elastic_angle, getattr(n, "elastic_angle", None), keep)
        _pack_table(d.inelastic_angle, getattr(n, "inelastic_angle", None), keep)
        _pack_table(d.inelastic_energy, getattr(n, "inelastic_energy", None), keep)
        pack_inelastic_law(d, n, keep)
    md = (MaterialDesc * len(lib.materials))()
    for d, m in zip(md, lib.materials):
        idx = np.ascontiguousarray(m.nuclides, np.int32)
        den = np.ascontiguousarray(m.densities, np.float64)
        keep += [idx, den]
        d.n_nuclides, d.nuclide, d.atom_density = idx.size, idx.ctypes.data, den.ctypes.data
        if getattr(m, "isotropic", None) is not None:
            iso = np.ascontiguousarray(m.isotropic, np.int32)
            assert iso.size == idx.size
            keep.append(iso)
            d.isotropic_in_lab = iso.ctypes.data
    return nd, md, keep
