// orc_core.h -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain scalar C++ restatement of the reference algorithms on MaCaw's particle-history hot path
// (SURVEY.md 8a).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, link or call anything under oracle/.  libmacaw_b200.so never does.
//
// Structure follows the reference on purpose: history-based, array-of-structs doubles, one scratch
// particle per history, per-nuclide micro cross-section cache, std::lower_bound searches.
// Every function cites the reference lines it restates (MC = MaCaw src/, RT =
// moose/modules/ray_tracing, OMC = openmc/).
//
// Parity pinning (see tests/test_oracle_*.py and DESIGN.md):
//   - RNG + samplers (prn, future_seed, init_seed, maxwell/watt spectrum, rotate_angle) are checked
//     bit-exactly against the reference's own sources compiled into oracle/_ref/libomc_ref.so and
//     against OpenMC's unit-test known answers (OMC tests/unit_tests/test_math.py:156-205);
//   - seed bookkeeping (seed-through-double rounding, RN count per event, URR stream advance) is
//     checked against the reference gold file test/tests/study/gold/openmc_study_rays.e through
//     tests/golden/study_rays_*.npz;
//   - the element walk is checked against RT's HEX8/QUAD4 `lots.i` total-distance gold CSVs;
//   - cross-section lookup, collision physics and tally scoring are checked bit for bit against the reference's own
//     compiled Material::calculate_xs / Particle::event_collide / score_collision_tally running on Nuclide / Material /
//     Tally objects built by its own constructors from an in-memory nuclide file of the synthetic library
//     (oracle/ref_shim_physics.cpp, ref_shim_tally.cpp, ref_hdf5_mem.*);
//   - "parity unpinned" remains only for physics the synthetic library does not carry (S(a,b), URR tables, tabulated
//     secondaries: needs ENDF/B HDF5) and for libMesh itself, stated in DESIGN.md.
//
// Math library: by default log/sin/cos come from macaw_b200/csrc/mcw_detmath.h (the deterministic
// libm replacement both sides use, so CPU and GPU histories are bit-identical); build with
// -DMCW_ORACLE_LIBM to use glibc like the reference does.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/macaw_b200.h"
#include "../macaw_b200/csrc/mcw_detmath.h"

namespace orc
{
#ifdef MCW_ORACLE_LIBM
inline double xlog(double x) { return std::log(x); }
inline double xsin(double x) { return std::sin(x); }
inline double xcos(double x) { return std::cos(x); }
inline double xexp(double x) { return std::exp(x); }
inline double xsinh(double x) { return std::sinh(x); }
#else
inline double xlog(double x) { return mcw_log(x); }
inline double xsin(double x) { return mcw_sin(x); }
inline double xcos(double x) { return mcw_cos(x); }
inline double xexp(double x) { return mcw_exp(x); }
inline double xsinh(double x) { return mcw_sinh(x); }
#endif

constexpr double PI = 3.1415926535897932384626433832795028841971693993751058209749445923078164062;
constexpr double K_BOLTZMANN = 8.617333262e-5; // OMC constants.h:90
constexpr double FREE_GAS_THRESHOLD = 400.0;    // OMC constants.h:42
constexpr double TRACE_TOLERANCE = 1e-8;        // RT TraceRayTools.h:45
constexpr double LOOSE_TRACE_TOLERANCE = 1e-5;  // RT TraceRayTools.h:47

// ------------------------------------------------------------------ RNG (OMC random_lcg.cpp:12-114)
constexpr uint64_t PRN_MULT = 6364136223846793005ULL;
constexpr uint64_t PRN_ADD = 1442695040888963407ULL;
constexpr uint64_t PRN_STRIDE = 152917ULL;
enum
{
  STREAM_TRACKING = 0,
  STREAM_SOURCE = 1,
  STREAM_URR = 2
};

inline double prn(uint64_t* seed)
{
  *seed = PRN_MULT * (*seed) + PRN_ADD;
  uint64_t word = ((*seed >> ((*seed >> 59u) + 5u)) ^ *seed) * 12605985483714917081ull;
  uint64_t result = (word >> 43u) ^ word;
  return (double)result * 5.42101086242752217003726400434970855712890625e-20; // ldexp(result,-64)
}

inline uint64_t future_seed(uint64_t n, uint64_t seed)
{
  uint64_t g = PRN_MULT, c = PRN_ADD, g_new = 1, c_new = 0;
  while (n > 0)
  {
    if (n & 1)
    {
      g_new *= g;
      c_new = c_new * g + c;
    }
    c *= (g + 1);
    g *= g;
    n >>= 1;
  }
  return g_new * seed + c_new;
}

inline uint64_t init_seed(int64_t id, int offset, int64_t master_seed)
{
  return future_seed((uint64_t)id * PRN_STRIDE, (uint64_t)(master_seed + offset));
}

// MaCaw keeps seeds in double-typed ray aux data (MC CollisionKernel.C:132-134,157-159,217-219,
// OpenMCStudy.C:345-347): uint64 -> double (round to nearest even) -> uint64.
inline uint64_t seed_via_double(uint64_t s)
{
  double d = (double)s;
  if (d >= 18446744073709551616.0)
    return 0; // x86 cvttsd2si wrap for 2^64; probability 2^-54 per conversion
  return (uint64_t)d;
}

// ------------------------------------------------------------------ samplers
inline double uniform_distribution(double a, double b, uint64_t* s) { return a + (b - a) * prn(s); }

// OMC random_dist.cpp:15-27
inline double maxwell_spectrum(double T, uint64_t* s)
{
  double r1 = prn(s), r2 = prn(s), r3 = prn(s);
  double c = xcos(PI / 2. * r3);
  return -T * (xlog(r1) + xlog(r2) * c * c);
}

// OMC random_dist.cpp:29-34
inline double watt_spectrum(double a, double b, uint64_t* s)
{
  double w = maxwell_spectrum(a, s);
  return w + 0.25 * a * a * b + uniform_distribution(-1., 1., s) * std::sqrt(a * a * b * w);
}

struct V3
{
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 cross(V3 a, V3 b)
{
  return {a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x};
}
inline V3 unit(V3 a)
{
  double l = norm(a);
  return {a.x / l, a.y / l, a.z / l};
}

// OMC math_functions.cpp:644-672 (phi sampled: 1 RN)
inline V3 rotate_angle(V3 u, double mu, const double* phi, uint64_t* s)
{
  double phi_ = phi ? *phi : 2.0 * PI * prn(s);
  double sinphi = xsin(phi_), cosphi = xcos(phi_);
  double a = std::sqrt(std::fmax(0., 1. - mu * mu));
  double b = std::sqrt(std::fmax(0., 1. - u.z * u.z));
  if (b > 1e-10)
  {
    return {mu * u.x + a * (u.x * u.z * cosphi - u.y * sinphi) / b,
            mu * u.y + a * (u.y * u.z * cosphi + u.x * sinphi) / b, mu * u.z - a * b * cosphi};
  }
  else
  {
    b = std::sqrt(1. - u.y * u.y);
    return {mu * u.x + a * (-u.x * u.y * sinphi + u.z * cosphi) / b, mu * u.y + a * b * sinphi,
            mu * u.z - a * (u.y * u.z * sinphi + u.x * cosphi) / b};
  }
}

// OMC distribution_multi.cpp:73-79
inline V3 isotropic_direction(uint64_t* s)
{
  double phi = uniform_distribution(0., 2.0 * PI, s);
  double mu = uniform_distribution(-1., 1., s);
  return {mu, std::sqrt(1.0 - mu * mu) * xcos(phi), std::sqrt(1.0 - mu * mu) * xsin(phi)};
}

// ------------------------------------------------------------------ nuclear data
// Distributions tabulated on a grid of incoming energies (mcw_tabular2d; the layouts OMC distribution_angle.cpp:20-63
// and distribution_energy.cpp:56-148 read)
struct Tabular2D
{
  std::vector<double> energy, x, p, c;
  std::vector<int> off, interp, ndisc;
  bool hist_in_energy = false;
  bool empty() const { return energy.empty(); }
  void take(const mcw_tabular2d& t, bool continuous)
  {
    *this = Tabular2D();
    if (t.n_energy == 0)
      return;
    const int n = t.n_energy;
    energy.assign(t.energy, t.energy + n);
    off.assign(t.offset, t.offset + n + 1);
    interp.assign(t.interpolation, t.interpolation + n);
    ndisc.assign(n, 0);
    if (continuous && t.n_discrete)
      ndisc.assign(t.n_discrete, t.n_discrete + n);
    hist_in_energy = continuous && t.histogram_in_energy != 0;
    x.assign(t.x, t.x + off[n]);
    p.assign(t.pdf, t.pdf + off[n]);
    c.assign(t.cdf, t.cdf + off[n]);
  }
};

// mcw_function1d: OMC Polynomial (endf.cpp:112-121) or Tabulated1D with one region, histogram or lin-lin (:156-210)
struct Fn
{
  int kind = MCW_FN_NONE, interp = 2;
  std::vector<double> x, y;
  void take(const mcw_function1d& f)
  {
    kind = f.kind;
    interp = f.interpolation;
    x.clear(), y.clear();
    if (kind == MCW_FN_NONE)
      return;
    y.assign(f.y, f.y + f.n);
    if (kind == MCW_FN_TABULATED)
      x.assign(f.x, f.x + f.n);
  }
  double operator()(double v) const
  {
    if (kind == MCW_FN_POLYNOMIAL)
    {
      double r = 0.0;
      for (size_t k = y.size(); k-- > 0;)
        r = r * v + y[k];
      return r;
    }
    const int n = (int)x.size();
    if (v < x[0])
      return y[0];
    if (v > x[n - 1])
      return y[n - 1];
    int i;
    if (x[0] == v)
      i = 0;
    else
      i = (int)(std::lower_bound(x.begin(), x.end(), v) - x.begin()) - 1;
    if (interp == 1)
      return y[i];
    const double r = (v - x[i]) / (x[i + 1] - x[i]);
    return y[i] + r * (y[i + 1] - y[i]);
  }
};

// mcw_fission_product: one neutron product of the fission reaction
struct FisProduct
{
  Fn yield, a, b;
  int law = MCW_LAW_WATT;
  double u = 0;
  Tabular2D energy;
};

// One scattering reaction other than elastic (an entry of OMC Nuclide::index_inelastic_scatter_): threshold, the neutron
// product's distribution and its multiplicity.  The nuclide's first one is the Nuclide itself (its base class: the
// descriptor carries it inline), further ones are Nuclide::more.
struct RxLaw
{
  int mt = 51;
  int inel_thr = 0;
  double inel_E = 0, inel_ratio = 0;
  Tabular2D inel_angle, inel_energy; // optional tabulated secondary distributions
  int inel_law = MCW_LAW_LEVEL;
  bool inel_in_cm = true;
  // product distributions that give energy and cosine together: correlated (one angular Tabular per point of inel_energy:
  // Tabular2D with one "incoming energy" per point), Kalbach-Mann (r, a per point), N-body
  Tabular2D corr_angle;
  std::vector<double> km_r, km_a;
  int nbody_n = 0;
  double nbody_mass_ratio = 0, nbody_q = 0;
  double yield = 1.0;         // constant multiplicity of the neutron product ...
  Fn yield_fn;                // ... or a function of the incoming energy (kind 0: the constant)
  double law_u = 0;           // MCW_LAW_MAXWELL / _EVAPORATION / _WATT: restriction energy, theta(E) or a(E), b(E)
  Fn law_a, law_b;
  std::vector<double> rx_xs;  // further reactions only: cross section from the threshold on
};

// What every cross-section lookup touches, kept together at the start of Nuclide (two cache lines): the lookup walks up
// to 35 nuclides per material and the CPU baseline is timed on it
struct NuclideHot
{
  int G = 0;
  bool fissionable = false;
  double awr = 1, kT = 0;
  std::vector<double> energy, xs;
  std::vector<int> grid_index;
  std::vector<double> urr_energy; // unresolved-resonance probability tables: energies (empty = none)
};

struct Nuclide : NuclideHot, RxLaw
{
  std::vector<double> elastic, inelastic;
  int n_inel = 0;
  std::vector<RxLaw> more; // the scattering reactions after the first
  // Several temperatures (OMC Nuclide::kTs_ / grid_[t] / xs_[t] / reactions_[r]->xs_[t] / urr_data_[t]): every data set is
  // an entry of Library::nuclides; entries [0, n_identities) are the nuclides themselves (materials, tallies and the URR
  // stream count those), further entries are additional data sets with `identity` naming their nuclide.  `sets` (on the
  // nuclide's own entry) lists all of its data sets by increasing kT.
  int identity = -1;
  std::vector<int> sets;
  double nu_poly[2] = {0, 0};
  int n_precursor = 0;
  double dyield[MCW_MAX_DELAYED_GROUPS], dtheta[MCW_MAX_DELAYED_GROUPS];
  double watt_a = 0, watt_b = 0, q = 0;
  Tabular2D el_angle;
  // unresolved-resonance probability tables (mcw_urr_desc; OMC urr.cpp UrrData); urr_energy empty = none
  std::vector<double> urr_table; // table[n_energy][6][n_band]
  int urr_bands = 0, urr_inel_flag = -1, urr_multiply_smooth = 0;
  double urr_prob(int i_energy, int param, int band) const { return urr_table[((size_t)i_energy * 6 + param) * urr_bands + band]; }

  // fission neutrons as a nuclide file has them (mcw_fission_product x (1 + n_precursor), total nu); empty: the scalars
  std::vector<FisProduct> fis;
  Fn total_nu_fn;

  // OMC endf.cpp:112-121 Polynomial (Horner, increasing powers) on nu_poly; nuclide.cpp:478-517
  double nu_total(double E) const
  {
    if (!fis.empty())
      return total_nu_fn.kind != MCW_FN_NONE ? total_nu_fn(E) : fis[0].yield(E); // :510-516
    double y = 0.0;
    y = y * E + nu_poly[1];
    y = y * E + nu_poly[0];
    return y;
  }
  double nu_delayed(double E) const
  {
    double nu = 0.0;
    if (!fis.empty())
    {
      for (size_t i = 1; i < fis.size(); ++i) // :494-503
        nu += fis[i].yield(E);
      return nu;
    }
    for (int g = 0; g < n_precursor; ++g)
      nu += dyield[g];
    return nu;
  }
};

struct Material
{
  std::vector<int> nuc;
  std::vector<double> dens;
  std::vector<int> iso; // isotropic-in-lab flags (OMC material.cpp:255-270 p0_)
  std::vector<int> sab_table;   // per nuclide position: S(a,b) table index or -1 (OMC material.cpp thermal_tables_); empty = none
  std::vector<double> sab_frac; // per nuclide position: fraction bound
  bool fissionable = false;
};

struct Micro // subset of OMC particle_data.h:82-112 NuclideMicroXS
{
  double total, absorption, fission, nu_fission;
  int i_grid;
  double f;
  int entry = -1; // the data set of the nuclide these values come from (micro.index_temp: Library::pick_set)
  // S(a,b) treatment (OMC nuclide.cpp:740-791): index_sab, thermal = sab_frac * (elastic + inelastic), thermal_elastic,
  // and the corrected micro.elastic; i_sab < 0 = free-atom scattering (elastic is then evaluated on demand)
  int i_sab = -1;
  double thermal = 0.0, thermal_elastic = 0.0, elastic = 0.0;
  bool use_ptable = false; // URR probability table used: `elastic` holds the table's value (OMC nuclide.cpp:795-924)
};

// One temperature of a thermal scattering table (mcw_thermal_desc; OMC thermal.cpp ThermalData + secondary_thermal.cpp)
struct ThermalTable
{
  double energy_max = 0;
  int inel_type = 0, el_type = 0, n_mu = 0, n_e_out = 0, skewed = 0, n_el_mu = 0;
  std::vector<double> ixs_e, ixs, ien, eout, pdf, cdf, mu, el_x, el_y, el_mu;
  std::vector<int> ioff;
};

struct Macro // OMC particle_data.h:134-146
{
  double total = 0, absorption = 0, fission = 0, nu_fission = 0;
};

struct Library
{
  std::vector<Nuclide> nuclides;
  std::vector<Material> materials;
  std::vector<ThermalTable> thermal;
  // Several temperatures of an S(a,b) table (OMC ThermalScattering::kTs_ / data_[t]): every temperature is an entry of
  // `thermal`; thermal_sets[t] (for the entry t that materials name) lists all entries of that table by increasing kT,
  // thermal_kT[k] is entry k's kT.  Empty: every table has one temperature.
  std::vector<std::vector<int>> thermal_sets;
  std::vector<double> thermal_kT;
  // ThermalScattering::calculate_xs (OMC thermal.cpp:149-171), nearest temperature: the closer of the two bounding ones
  int pick_thermal(int t, double kT) const
  {
    if (thermal_sets.empty() || thermal_sets[t].size() <= 1)
      return t;
    const std::vector<int>& s = thermal_sets[t];
    const size_t n = s.size();
    size_t i = 0;
    while (thermal_kT[s[i + 1]] < kT && i + 1 < n - 1)
      ++i;
    if (kT - thermal_kT[s[i]] > thermal_kT[s[i + 1]] - kT)
      ++i;
    return s[i];
  }
  double E_min = 0, E_max = 0, log_spacing = 0;
  int n_log_bins = 8000;
  int n_identities = 0; // nuclides (OMC data::nuclides.size()); entries beyond are further temperature sets.  0: all
  bool multi_temperature() const { return (n_identities > 0 && (size_t)n_identities < nuclides.size()) || !thermal_sets.empty(); }
  int identities() const { return n_identities > 0 ? n_identities : (int)nuclides.size(); }

  // OMC simulation.cpp:553-619 initialize_data + nuclide.cpp:446-476 init_grid
  void finalize()
  {
    E_min = 0.0;
    E_max = INFINITY;
    // every nuclide's FIRST (lowest-temperature) grid bounds the energy range (simulation.cpp:558-566 grid_[0])
    for (int i = 0; i < identities(); ++i)
    {
      const Nuclide& n = nuclides[nuclides[i].sets.empty() ? i : nuclides[i].sets[0]];
      E_min = std::max(E_min, n.energy.front());
      E_max = std::min(E_max, n.energy.back());
    }
    const int M = n_log_bins;
    double spacing = xlog(E_max / E_min) / M;
    for (auto& n : nuclides)
    {
      n.grid_index.assign(M + 1, 0);
      int j = 0;
      double step = (M * spacing - 0.0) / M; // xt::linspace(0, M*spacing, M+1)
      for (int k = 0; k <= M; ++k)
      {
        double umesh = 0.0 + k * step;
        while (xlog(n.energy[j + 1] / E_min) <= umesh)
        {
          if (j + 2 == n.G)
            break;
          ++j;
        }
        n.grid_index[k] = j;
      }
    }
    log_spacing = xlog(E_max / E_min) / M;
    for (auto& m : materials)
    {
      m.fissionable = false;
      for (int i : m.nuc)
        if (nuclides[i].fissionable)
          m.fissionable = true;
    }
  }

  // OMC nuclide.cpp:614-626 TemperatureMethod::NEAREST: the data set whose kT is closest to the particle's (the first of
  // equally close ones; kTs_ is sorted).  kT = sqrtkT * sqrtkT with sqrtkT = sqrt(K_BOLTZMANN * T) as MaCaw sets it
  // (MC CollisionKernel.C:109).
  int pick_set(int i_nuc, double kT) const
  {
    const Nuclide& n = nuclides[i_nuc];
    if (n.sets.size() <= 1)
      return i_nuc;
    int i_temp = -1;
    double max_diff = INFINITY;
    for (size_t t = 0; t < n.sets.size(); ++t)
    {
      const double diff = std::fabs(nuclides[n.sets[t]].kT - kT);
      if (diff < max_diff)
      {
        i_temp = (int)t;
        max_diff = diff;
      }
    }
    return n.sets[i_temp];
  }
  static double kT_of(double T)
  {
    const double sqrtkT = std::sqrt(K_BOLTZMANN * T);
    return sqrtkT * sqrtkT;
  }

  // OMC search.h:14-20
  static int lower_bound_index(const double* first, const double* last, double value)
  {
    if (*first == value)
      return 0;
    const double* it = std::lower_bound(first, last, value) - 1;
    return (int)(it - first);
  }

  // OMC nuclide.cpp:557-763 (NEAREST temperature, single data set, no WMP / S(a,b) / URR)
  Micro micro_xs(int i_nuc, double E, int i_log) const
  {
    const Nuclide& n = nuclides[i_nuc];
    Micro m;
    int i_grid;
    if (E < n.energy.front())
      i_grid = 0;
    else if (E > n.energy.back())
      i_grid = n.G - 2;
    else
    {
      int i_low = n.grid_index[i_log];
      int i_high = n.grid_index[i_log + 1] + 1;
      i_grid = i_low + lower_bound_index(&n.energy[i_low], &n.energy[i_high], E);
    }
    if (n.energy[i_grid] == n.energy[i_grid + 1])
      ++i_grid;
    double f = (E - n.energy[i_grid]) / (n.energy[i_grid + 1] - n.energy[i_grid]);
    m.i_grid = i_grid;
    m.f = f;
    const double* r0 = &n.xs[5 * (size_t)i_grid];
    const double* r1 = r0 + 5;
    m.total = (1.0 - f) * r0[0] + f * r1[0];
    m.absorption = (1.0 - f) * r0[1] + f * r1[1];
    if (n.fissionable)
    {
      m.fission = (1.0 - f) * r0[2] + f * r1[2];
      m.nu_fission = (1.0 - f) * r0[3] + f * r1[3];
    }
    else
    {
      m.fission = 0.0;
      m.nu_fission = 0.0;
    }
    return m;
  }

  int log_bin(double E) const
  {
    int i = (int)(xlog(E / E_min) / log_spacing); // OMC material.cpp:767-768
    if (i < 0)
      i = 0; // undefined in the reference (out-of-range read); clamp
    if (i > n_log_bins - 1)
      i = n_log_bins - 1;
    return i;
  }

  // Tabulated1D::operator() with one lin-lin region (OMC endf.cpp:153-212)
  static double tabulated1d(const std::vector<double>& x, const std::vector<double>& y, double v)
  {
    const int n = (int)x.size();
    if (v < x[0])
      return y[0];
    if (v > x[n - 1])
      return y[n - 1];
    const int i = lower_bound_index(x.data(), x.data() + n, v);
    const double r = (v - x[i]) / (x[i + 1] - x[i]);
    return y[i] + r * (y[i + 1] - y[i]);
  }

  // ThermalData::calculate_xs (OMC thermal.cpp:229-243) with CoherentElasticXS::operator() (endf.cpp:229-239)
  void thermal_xs(const ThermalTable& t, double E, double& elastic, double& inelastic) const
  {
    if (t.el_type == MCW_SAB_ELASTIC_COHERENT)
    {
      if (E < t.el_x[0])
        elastic = 0.0;
      else
        elastic = t.el_y[lower_bound_index(t.el_x.data(), t.el_x.data() + t.el_x.size(), E)] / E;
    }
    else if (t.el_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
      elastic = tabulated1d(t.el_x, t.el_y, E);
    else
      elastic = 0.0;
    inelastic = tabulated1d(t.ixs_e, t.ixs, E);
  }

  // Nuclide::calculate_sab_xs (OMC nuclide.cpp:763-791) on a freshly interpolated micro; the free-atom elastic cross
  // section is Nuclide::calculate_elastic_xs (:517-529)
  void apply_sab(const Nuclide& n, int i_sab, double sab_frac, double E, Micro& m) const
  {
    m.i_sab = i_sab;
    double elastic, inelastic;
    thermal_xs(thermal[i_sab], E, elastic, inelastic);
    m.thermal = sab_frac * (elastic + inelastic);
    m.thermal_elastic = sab_frac * elastic;
    const double free_elastic = (1.0 - m.f) * n.elastic[m.i_grid] + m.f * n.elastic[m.i_grid + 1];
    m.total = m.total + m.thermal - sab_frac * free_elastic;
    m.elastic = m.thermal + (1.0 - sab_frac) * free_elastic;
  }

  // Nuclide::calculate_urr_xs (OMC nuclide.cpp:795-924), lin-lin tables; seed = the history's URR stream
  void apply_urr(const Nuclide& n, int i_nuc, double E, uint64_t seed, Micro& m) const
  {
    m.use_ptable = true;
    int i_energy = 0;
    while (E >= n.urr_energy[i_energy + 1])
      ++i_energy;
    uint64_t fseed = future_seed((uint64_t)i_nuc, seed); // future_prn (random_lcg.cpp:49-53)
    const double r = prn(&fseed);
    int i_low = 0;
    while (n.urr_prob(i_energy, 0, i_low) <= r)
      ++i_low;
    int i_up = 0;
    while (n.urr_prob(i_energy + 1, 0, i_up) <= r)
      ++i_up;
    double f = (E - n.urr_energy[i_energy]) / (n.urr_energy[i_energy + 1] - n.urr_energy[i_energy]);
    double elastic = (1. - f) * n.urr_prob(i_energy, 2, i_low) + f * n.urr_prob(i_energy + 1, 2, i_up);
    double fission = (1. - f) * n.urr_prob(i_energy, 3, i_low) + f * n.urr_prob(i_energy + 1, 3, i_up);
    double capture = (1. - f) * n.urr_prob(i_energy, 4, i_low) + f * n.urr_prob(i_energy + 1, 4, i_up);
    double inelastic = 0.;
    if (n.urr_inel_flag != -1)
    {
      f = m.f;
      const int xs_index = m.i_grid - n.inel_thr;
      if (xs_index >= 0)
        inelastic = (1. - f) * n.inelastic[m.i_grid] + f * n.inelastic[m.i_grid + 1]; // stored on the full grid here
    }
    if (n.urr_multiply_smooth)
    {
      const double free_elastic = (1.0 - m.f) * n.elastic[m.i_grid] + m.f * n.elastic[m.i_grid + 1];
      elastic *= free_elastic;
      capture *= (m.absorption - m.fission);
      fission *= m.fission;
    }
    if (elastic < 0.)
      elastic = 0.;
    if (fission < 0.)
      fission = 0.;
    if (capture < 0.)
      capture = 0.;
    m.elastic = elastic;
    m.absorption = capture + fission;
    m.fission = fission;
    m.total = elastic + inelastic + capture + fission;
    if (n.fissionable)
      m.nu_fission = n.nu_total(E) * m.fission;
  }

  // OMC material.cpp:748-831.  seed_urr: the history's URR stream seed (null: no history, smooth cross sections).  kT: the
  // particle's temperature, which picks the data sets of libraries with several temperatures; the loop without the picks
  // is a separate instantiation (timed on a quiet host the picks cost the baseline 3 percent even when never taken:
  // profiles/oracle_ab_r02y.txt).
  Macro macro_xs(int i_mat, double E, Micro* cache /* [materials[i_mat].nuc.size()] or null */,
                 int64_t* n_nuclide_lookups = nullptr, const uint64_t* seed_urr = nullptr, double kT = 0.0) const
  {
    return multi_temperature() ? macro_xs_impl<true>(i_mat, E, cache, n_nuclide_lookups, seed_urr, kT)
                               : macro_xs_impl<false>(i_mat, E, cache, n_nuclide_lookups, seed_urr, kT);
  }
  template <bool PICK>
  Macro macro_xs_impl(int i_mat, double E, Micro* cache, int64_t* n_nuclide_lookups, const uint64_t* seed_urr, double kT) const
  {
    const Material& mat = materials[i_mat];
    Macro M;
    int i_log = log_bin(E);
    for (size_t i = 0; i < mat.nuc.size(); ++i)
    {
      const int entry = PICK ? pick_set(mat.nuc[i], kT) : mat.nuc[i];
      Micro m = micro_xs(entry, E, i_log);
      m.entry = entry;
      // S(a,b) table of this nuclide in this material, unless the energy is above the table (material.cpp:778-806)
      if (!mat.sab_table.empty() && mat.sab_table[i] >= 0 && !(E > thermal[mat.sab_table[i]].energy_max))
        apply_sab(nuclides[entry], PICK ? pick_thermal(mat.sab_table[i], kT) : mat.sab_table[i], mat.sab_frac[i], E, m);
      {
        // probability tables (those of the chosen data set, urr_data_[i_temp]), strictly inside their energy range
        // (nuclide.cpp:753-759); the band is drawn with the NUCLIDE's index (index_, :809)
        const Nuclide& n = nuclides[entry];
        if (seed_urr && !n.urr_energy.empty() && E > n.urr_energy.front() && E < n.urr_energy.back())
          apply_urr(n, mat.nuc[i], E, *seed_urr, m);
      }
      if (cache)
        cache[i] = m;
      double d = mat.dens[i];
      M.total += d * m.total;
      M.absorption += d * m.absorption;
      M.fission += d * m.fission;
      M.nu_fission += d * m.nu_fission;
    }
    if (n_nuclide_lookups)
      *n_nuclide_lookups += (int64_t)mat.nuc.size();
    return M;
  }
};

} // namespace orc
