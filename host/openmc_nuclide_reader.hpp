// openmc_nuclide_reader.hpp -- the MaCaw-side `gatherLibrary()` of INTEGRATION.md section 3.2: OpenMC nuclide files ->
// macaw::Library, written against OpenMC's own include/openmc/hdf5_interface.h (open_group, read_attribute, read_dataset,
// group_names, ...).  In a MaCaw build those functions sit on libhdf5 and `group` comes from
// openmc::file_open(path) / open_group(file, name); in this repository the same header is implemented over an in-memory
// tree (oracle/ref_hdf5_mem.cpp), which is how the reader is tested: oracle/_ref lets the reference's OWN Nuclide /
// Reaction / ReactionProduct / UrrData constructors read a nuclide-file image and this reader read the same image, and
// tests/test_oracle_pinning.py::test_cpp_nuclide_reader_against_reference_nuclide_cpp holds the library's digest
// (mcw_digest) of what it gathered against the digest of the library the image was written from, with the reference's
// Nuclide::create_derived tables in it -- i.e. the derived total / absorption / fission / nu-fission tables are summed in
// create_derived's order and come out bit-identical.  macaw_b200/openmc_hdf5.py is the same reader in Python.
//
// Why read the files when openmc::data::nuclides is in memory: what the secondary distributions hold is private without
// accessors in OpenMC (distribution_angle.h:33, distribution_energy.h:41-149, secondary_kalbach.h:34, ...).
//
// What the library does not carry is refused (macaw::Error with the reason), never dropped.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "openmc/hdf5_interface.h"
#include "xtensor/xarray.hpp"

#include "macaw_host.hpp"

namespace macaw
{
namespace omcfile
{
inline bool isFission(int mt) { return mt == 18 || mt == 19 || mt == 20 || mt == 21 || mt == 38; } // OMC endf.cpp:37-40
inline bool isDisappearance(int mt)                                                                  // :42-51
{
  return (mt >= 101 && mt <= 117) || mt == 155 || mt == 182 || mt == 191 || mt == 192 || mt == 193 || mt == 197;
}
inline bool isInelasticScatter(int mt) // :63-78
{
  if (mt < 100)
    return !isFission(mt) && mt >= 5 && mt != 27;
  if (mt <= 200)
    return !isDisappearance(mt);
  return mt >= 875 && mt <= 891;
}
[[noreturn]] inline void refuse(const std::string& what) { throw Error(MCW_E_UNSUPPORTED, "nuclide file: " + what); }

inline xt::xarray<double> array2d(hid_t group, const char* name)
{
  hid_t d = openmc::open_dataset(group, name);
  xt::xarray<double> a;
  openmc::read_dataset(d, a);
  openmc::close_dataset(d);
  return a;
}

// read_function (OMC endf.cpp:80-100): Polynomial, or Tabulated1D whose regions all interpolate the same way (1 or 2)
inline Function1D readFunction(hid_t group, const char* name)
{
  hid_t d = openmc::open_dataset(group, name);
  std::string type;
  openmc::read_attribute(d, "type", type);
  Function1D f;
  if (type == "Polynomial")
  {
    f.kind = MCW_FN_POLYNOMIAL;
    openmc::read_dataset(d, f.y);
  }
  else if (type == "Tabulated1D")
  {
    std::vector<int> nbt, interp;
    openmc::read_attribute(d, "breakpoints", nbt);
    openmc::read_attribute(d, "interpolation", interp);
    xt::xarray<double> a;
    openmc::read_dataset(d, a);
    const size_t n = a.shape()[1];
    for (int c : interp)
      if (c != interp[0])
        refuse("Tabulated1D with several regions of different interpolation");
    if ((interp[0] != 1 && interp[0] != 2) || (size_t)nbt.back() != n)
      refuse("Tabulated1D with logarithmic interpolation");
    f.kind = MCW_FN_TABULATED;
    f.interpolation = interp[0];
    f.x.assign(a.data(), a.data() + n);
    f.y.assign(a.data() + n, a.data() + 2 * n);
  }
  else
    refuse("function type '" + type + "'");
  openmc::close_dataset(d);
  return f;
}
// Tabulated1D::operator() / Polynomial::operator() (OMC endf.cpp:112-121,156-210), operation for operation
inline double evaluate(const Function1D& f, double v)
{
  if (f.kind == MCW_FN_POLYNOMIAL)
  {
    double r = 0.0;
    for (size_t k = f.y.size(); k-- > 0;)
      r = r * v + f.y[k];
    return r;
  }
  const size_t n = f.x.size();
  if (v < f.x[0])
    return f.y[0];
  if (v > f.x[n - 1])
    return f.y[n - 1];
  size_t i = f.x[0] == v ? 0 : (size_t)(std::lower_bound(f.x.begin(), f.x.end(), v) - f.x.begin()) - 1;
  if (f.interpolation == 1)
    return f.y[i];
  const double r = (v - f.x[i]) / (f.x[i + 1] - f.x[i]);
  return f.y[i] + r * (f.y[i + 1] - f.y[i]);
}

// (3, N) x / pdf / cdf rows with `offsets` / `interpolation` [/ `n_discrete_lines`] attributes on the dataset
inline Tabular2D readTable(hid_t group, const char* dataset, const std::vector<double>& energy, bool continuous, int rows = 3)
{
  Tabular2D t;
  t.energy = energy;
  hid_t d = openmc::open_dataset(group, dataset);
  std::vector<int> off, interp, nd;
  openmc::read_attribute(d, "offsets", off);
  openmc::read_attribute(d, "interpolation", interp);
  if (continuous)
    openmc::read_attribute(d, "n_discrete_lines", nd);
  xt::xarray<double> a;
  openmc::read_dataset(d, a);
  openmc::close_dataset(d);
  (void)rows;
  const size_t n = a.shape()[1];
  t.offset.assign(off.begin(), off.end());
  t.offset.push_back((int32_t)n);
  t.interpolation.assign(interp.begin(), interp.end());
  t.n_discrete.assign(nd.begin(), nd.end());
  t.x.assign(a.data(), a.data() + n);
  t.pdf.assign(a.data() + n, a.data() + 2 * n);
  t.cdf.assign(a.data() + 2 * n, a.data() + 3 * n);
  return t;
}
// the one incoming-energy interpolation region of an `energy` dataset: true = histogram
inline bool oneRegionIsHistogram(hid_t group, const char* what)
{
  hid_t d = openmc::open_dataset(group, "energy");
  xt::xarray<int> it;
  openmc::read_attribute(d, "interpolation", it);
  openmc::close_dataset(d);
  if (it.shape()[1] != 1 || (it(1, 0) != 1 && it(1, 0) != 2))
    refuse(std::string(what) + ": several incoming-energy interpolation regions");
  return it(1, 0) == 1;
}

// what one neutron product gives (ReactionProduct, OMC reaction_product.cpp:18-68, with its one distribution)
struct Product
{
  Reaction r; // law, in_cm, angle, energy, corr_*, km_*, nbody_*, yield, yield_fn, u, a, b, threshold_energy, mass_ratio
  bool has_law = true;
};
inline Product readProduct(hid_t p)
{
  Product out;
  Reaction& r = out.r;
  int n_dist = 0;
  openmc::read_attribute(p, "n_distribution", n_dist);
  if (n_dist != 1)
    refuse("a product with several distributions (applicability)");
  Function1D y = readFunction(p, "yield");
  if (y.kind == MCW_FN_POLYNOMIAL && y.y.size() == 1)
    r.yield = y.y[0];
  else
    r.yield_fn = y;
  hid_t d = openmc::open_group(p, "distribution_0");
  std::string type;
  openmc::read_attribute(d, "type", type);
  if (type == "uncorrelated")
  {
    if (openmc::object_exists(d, "angle"))
    {
      hid_t a = openmc::open_group(d, "angle");
      std::vector<double> e;
      openmc::read_dataset(a, "energy", e);
      r.angle = readTable(a, "mu", e, false);
      r.angle.n_discrete.clear();
      openmc::close_group(a);
    }
    if (!openmc::object_exists(d, "energy"))
      out.has_law = false; // elastic scattering: its energy follows from the kinematics
    else
    {
      hid_t e = openmc::open_group(d, "energy");
      std::string law;
      openmc::read_attribute(e, "type", law);
      if (law == "level")
      {
        r.law = MCW_LAW_LEVEL;
        openmc::read_attribute(e, "threshold", r.threshold_energy);
        openmc::read_attribute(e, "mass_ratio", r.mass_ratio);
      }
      else if (law == "continuous")
      {
        r.law = MCW_LAW_CONTINUOUS_TABULAR;
        std::vector<double> en;
        openmc::read_dataset(e, "energy", en);
        r.energy = readTable(e, "distribution", en, true);
        r.energy.histogram_in_energy = oneRegionIsHistogram(e, "continuous tabular law");
      }
      else if (law == "maxwell" || law == "evaporation" || law == "watt")
      {
        r.law = law == "maxwell" ? MCW_LAW_MAXWELL : (law == "evaporation" ? MCW_LAW_EVAPORATION : MCW_LAW_WATT);
        openmc::read_attribute(e, "u", r.u);
        r.a = readFunction(e, law == "watt" ? "a" : "theta");
        if (law == "watt")
          r.b = readFunction(e, "b");
      }
      else
        refuse("energy distribution type '" + law + "'"); // OMC secondary_uncorrelated.cpp:47
      openmc::close_group(e);
    }
  }
  else if (type == "correlated" || type == "kalbach-mann")
  {
    oneRegionIsHistogram(d, type.c_str());
    std::vector<double> en;
    openmc::read_dataset(d, "energy", en);
    const char* ds = type == "correlated" ? "energy_out" : "distribution";
    r.energy = readTable(d, ds, en, true);
    for (int32_t k : r.energy.n_discrete)
      if (k != 0)
        refuse(type + " distribution with discrete lines");
    r.energy.n_discrete.clear();
    const xt::xarray<double> t = array2d(d, ds);
    const size_t np = t.shape()[1];
    if (type == "correlated")
    {
      r.law = MCW_LAW_CORRELATED;
      const xt::xarray<double> mu = array2d(d, "mu");
      const size_t nm = mu.shape()[1];
      for (size_t k = 0; k < np; ++k)
      {
        r.corr_interp.push_back((int32_t)t(3, k));
        r.corr_offset.push_back((int32_t)t(4, k));
      }
      r.corr_offset.push_back((int32_t)nm);
      r.corr_mu.assign(mu.data(), mu.data() + nm);
      r.corr_pdf.assign(mu.data() + nm, mu.data() + 2 * nm);
      r.corr_cdf.assign(mu.data() + 2 * nm, mu.data() + 3 * nm);
    }
    else
    {
      r.law = MCW_LAW_KALBACH_MANN;
      for (size_t k = 0; k < np; ++k)
      {
        r.km_r.push_back(t(3, k));
        r.km_a.push_back(t(4, k));
      }
    }
  }
  else if (type == "nbody")
  {
    r.law = MCW_LAW_NBODY;
    openmc::read_attribute(d, "n_particles", r.nbody_n);
    openmc::read_attribute(d, "total_mass", r.nbody_mass_ratio);
    openmc::read_attribute(d, "q_value", r.nbody_q);
  }
  else
    refuse("angle-energy distribution type '" + type + "'");
  openmc::close_group(d);
  return out;
}
// the neutron products of a reaction group, in product order
inline std::vector<std::string> neutronProducts(hid_t rx)
{
  std::vector<std::pair<int, std::string>> found;
  for (const std::string& name : openmc::group_names(rx))
    if (name.rfind("product_", 0) == 0)
    {
      hid_t p = openmc::open_group(rx, name.c_str());
      std::string particle;
      openmc::read_attribute(p, "particle", particle);
      openmc::close_group(p);
      if (particle == "neutron")
        found.push_back({std::stoi(name.substr(8)), name});
    }
  std::sort(found.begin(), found.end());
  std::vector<std::string> out;
  for (auto& f : found)
    out.push_back(f.second);
  return out;
}

// One nuclide group: the nuclide (its lowest temperature), its further temperature sets, and its probability tables as
// (temperature index, table) -- index 0 is the nuclide's own entry.
struct GatheredNuclide
{
  Nuclide base;
  std::vector<Nuclide> more;
  std::vector<std::pair<int, UrrTable>> urr;
};
inline GatheredNuclide readNuclide(hid_t g, const std::string& name)
{
  GatheredNuclide out;
  double awr = 0;
  openmc::read_attribute(g, "atomic_weight_ratio", awr);
  hid_t kts = openmc::open_group(g, "kTs");
  std::vector<std::pair<double, std::string>> temps;
  for (const std::string& t : openmc::dataset_names(kts))
  {
    double kT = 0;
    openmc::read_dataset(kts, t.c_str(), kT);
    temps.push_back({kT, t});
  }
  openmc::close_group(kts);
  std::sort(temps.begin(), temps.end());
  hid_t rxs = openmc::open_group(g, "reactions");
  std::vector<std::string> rx_names;
  for (const std::string& n : openmc::group_names(rxs)) // name order = MT order, like Nuclide::Nuclide (nuclide.cpp:199-221)
    if (n.rfind("reaction_", 0) == 0)
      rx_names.push_back(n);
  hid_t eg = openmc::open_group(g, "energy");
  for (size_t it = 0; it < temps.size(); ++it)
  {
    const std::string& T = temps[it].second;
    Nuclide n;
    n.name = name;
    n.d.awr = awr;
    n.d.kT = temps[it].first;
    openmc::read_dataset(eg, T.c_str(), n.energy);
    const size_t G = n.energy.size();
    std::vector<double> total(G, 0.0), absorption(G, 0.0), fission(G, 0.0);
    std::vector<Reaction> scatter;
    bool have_elastic = false;
    for (const std::string& rn : rx_names)
    {
      hid_t rx = openmc::open_group(rxs, rn.c_str());
      int mt = 0, redundant = 0, cm = 0;
      openmc::read_attribute(rx, "mt", mt);
      openmc::read_attribute(rx, "redundant", redundant);
      openmc::read_attribute(rx, "center_of_mass", cm);
      hid_t tg = openmc::open_group(rx, T.c_str());
      hid_t xd = openmc::open_dataset(tg, "xs");
      int thr = 0;
      openmc::read_attribute(xd, "threshold_idx", thr);
      std::vector<double> xs;
      openmc::read_dataset(xd, xs);
      openmc::close_dataset(xd);
      openmc::close_group(tg);
      if (!redundant)
      {
        // Nuclide::create_derived (OMC nuclide.cpp:317-390), same order of additions
        for (size_t k = 0; k < xs.size(); ++k)
          total[thr + k] += xs[k];
        if (isDisappearance(mt))
          for (size_t k = 0; k < xs.size(); ++k)
            absorption[thr + k] += xs[k];
        if (isFission(mt))
        {
          if (mt != 18)
            refuse("partial fission reactions (MT 19 / 20 / 21 / 38)");
          for (size_t k = 0; k < xs.size(); ++k)
            fission[thr + k] += xs[k], absorption[thr + k] += xs[k];
          n.d.fissionable = 1;
          openmc::read_attribute(rx, "Q_value", n.d.q_fission);
          const auto prods = neutronProducts(rx);
          for (size_t k = 0; k < prods.size(); ++k)
          {
            hid_t p = openmc::open_group(rx, prods[k].c_str());
            const Product pr = readProduct(p);
            std::string mode;
            openmc::read_attribute(p, "emission_mode", mode);
            openmc::close_group(p);
            const Reaction& r = pr.r;
            if ((r.law != MCW_LAW_CONTINUOUS_TABULAR && r.law < MCW_LAW_MAXWELL) || !r.angle.energy.empty())
              refuse("fission product: law other than Watt / Maxwell / evaporation / continuous tabular, or an angle table");
            if (k > 0 && mode != "delayed")
              refuse("fission reaction with several prompt neutron products");
            FissionProduct f;
            if (r.yield_fn.kind != MCW_FN_NONE)
              f.yield = r.yield_fn;
            else
              f.yield.kind = MCW_FN_POLYNOMIAL, f.yield.y = {r.yield};
            f.law = r.law, f.u = r.u, f.a = r.a, f.b = r.b, f.energy = r.energy;
            n.fission_products.push_back(std::move(f));
          }
          n.d.n_precursor = (int32_t)n.fission_products.size() - 1;
        }
        else if (mt == 2)
        {
          if (thr != 0)
            refuse("elastic scattering with a threshold");
          n.elastic = xs;
          have_elastic = true;
          hid_t p = openmc::open_group(rx, neutronProducts(rx).at(0).c_str());
          n.elastic_angle = readProduct(p).r.angle;
          openmc::close_group(p);
        }
        else if (isInelasticScatter(mt))
        {
          hid_t p = openmc::open_group(rx, neutronProducts(rx).at(0).c_str());
          Product pr = readProduct(p);
          openmc::close_group(p);
          if (!pr.has_law)
            refuse("MT " + std::to_string(mt) + ": neutron product without an energy law");
          pr.r.mt = mt, pr.r.threshold = thr, pr.r.xs = xs, pr.r.in_cm = cm != 0;
          scatter.push_back(std::move(pr.r));
        }
      }
      openmc::close_group(rx);
    }
    if (!have_elastic)
      refuse("no elastic scattering reaction");
    if (n.d.fissionable && openmc::object_exists(g, "total_nu"))
    {
      hid_t nu = openmc::open_group(g, "total_nu");
      n.total_nu = readFunction(nu, "yield");
      openmc::close_group(nu);
    }
    n.xs.assign(5 * G, 0.0);
    for (size_t k = 0; k < G; ++k)
    {
      n.xs[5 * k] = total[k], n.xs[5 * k + 1] = absorption[k], n.xs[5 * k + 2] = fission[k];
      if (n.d.fissionable) // nu(E, total) * fission: the total-nu function, else the prompt yield (nuclide.cpp:510-516)
        n.xs[5 * k + 3] = evaluate(n.total_nu.kind != MCW_FN_NONE ? n.total_nu : n.fission_products.at(0).yield, n.energy[k]) * fission[k];
    }
    // the scattering reactions in index_inelastic_scatter_ order: the first inline, the others as a list
    for (size_t j = 0; j < scatter.size(); ++j)
    {
      Reaction& r = scatter[j];
      if (j > 0)
      {
        n.more_reactions.push_back(std::move(r));
        continue;
      }
      n.inelastic = r.xs;
      n.d.inel_threshold = r.threshold;
      n.d.inel_threshold_energy = r.threshold_energy, n.d.inel_mass_ratio = r.mass_ratio;
      n.inelastic_angle = r.angle, n.inelastic_energy = r.energy, n.inelastic_in_cm = r.in_cm;
      n.inelastic_law = r.law;
      n.corr_offset = r.corr_offset, n.corr_interp = r.corr_interp;
      n.corr_mu = r.corr_mu, n.corr_pdf = r.corr_pdf, n.corr_cdf = r.corr_cdf, n.km_r = r.km_r, n.km_a = r.km_a;
      n.d.nbody_n = r.nbody_n, n.d.nbody_mass_ratio = r.nbody_mass_ratio, n.d.nbody_q = r.nbody_q;
      n.inelastic_yield = r.yield, n.inelastic_yield_fn = r.yield_fn;
      n.inelastic_u = r.u, n.inelastic_a = r.a, n.inelastic_b = r.b;
    }
    if (it == 0)
      out.base = std::move(n);
    else
      out.more.push_back(std::move(n));
  }
  openmc::close_group(eg);
  openmc::close_group(rxs);
  if (openmc::object_exists(g, "urr"))
  {
    hid_t ug = openmc::open_group(g, "urr");
    for (size_t it = 0; it < temps.size(); ++it)
    {
      hid_t u = openmc::open_group(ug, temps[it].second.c_str());
      UrrTable t;
      openmc::read_attribute(u, "interpolation", t.interpolation);
      if (t.interpolation != 2)
        refuse("log-log probability tables");
      openmc::read_attribute(u, "inelastic", t.inelastic_flag);
      openmc::read_attribute(u, "absorption", t.absorption_flag);
      int ms = 0;
      openmc::read_attribute(u, "multiply_smooth", ms);
      t.multiply_smooth = ms;
      openmc::read_dataset(u, "energy", t.energy);
      hid_t td = openmc::open_dataset(u, "table");
      xt::xarray<double> tab;
      openmc::read_dataset(td, tab);
      openmc::close_dataset(td);
      t.n_band = (int)tab.shape()[2];
      t.table.assign(tab.data(), tab.data() + tab.size());
      openmc::close_group(u);
      out.urr.push_back({(int)it, std::move(t)});
    }
    openmc::close_group(ug);
  }
  return out;
}

// One table group of a thermal-scattering file (OMC ThermalScattering thermal.cpp:38-147, ThermalData :177-227): every
// temperature as one ThermalTable, by increasing kT, with the kTs.  Inelastic: incoherent_inelastic (continuous outgoing
// energies with equiprobable discrete cosines) or incoherent_inelastic_discrete; elastic: none, coherent_elastic (Bragg
// edges) or incoherent_elastic_discrete; the Debye-Waller form is refused.
inline std::vector<std::pair<double, ThermalTable>> readThermalTable(hid_t g)
{
  std::vector<std::pair<double, ThermalTable>> out;
  double energy_max = 0;
  openmc::read_attribute(g, "energy_max", energy_max);
  hid_t kts = openmc::open_group(g, "kTs");
  std::vector<std::pair<double, std::string>> temps;
  for (const std::string& t : openmc::dataset_names(kts))
  {
    double kT = 0;
    openmc::read_dataset(kts, t.c_str(), kT);
    temps.push_back({kT, t});
  }
  openmc::close_group(kts);
  std::sort(temps.begin(), temps.end());
  for (const auto& kt : temps)
  {
    ThermalTable t;
    t.energy_max = energy_max;
    hid_t tg = openmc::open_group(g, kt.second.c_str());
    hid_t in = openmc::open_group(tg, "inelastic");
    const Function1D xs = readFunction(in, "xs");
    if (xs.kind != MCW_FN_TABULATED || xs.interpolation != 2)
      refuse("thermal inelastic cross section that is not a lin-lin table");
    t.i_xs_energy = xs.x, t.i_xs = xs.y;
    hid_t d = openmc::open_group(in, "distribution");
    std::string type;
    openmc::read_attribute(d, "type", type);
    if (type == "incoherent_inelastic")
    {
      t.inelastic_type = MCW_SAB_INELASTIC_CONTINUOUS;
      oneRegionIsHistogram(d, "thermal inelastic distribution");
      openmc::read_dataset(d, "energy", t.i_energy);
      const Tabular2D eo = readTable(d, "energy_out", t.i_energy, true);
      const xt::xarray<double> raw = array2d(d, "energy_out"), mu = array2d(d, "mu");
      const size_t np = raw.shape()[1];
      if (np == 0 || mu.shape()[1] % np)
        refuse("thermal inelastic distribution whose cosines are not equiprobable discrete ones");
      t.n_mu = (int)(mu.shape()[1] / np);
      for (size_t k = 0; k < np; ++k)
        if (raw(3, k) != 0.0 || raw(4, k) != (double)(k * (size_t)t.n_mu))
          refuse("thermal inelastic distribution with tabulated cosines or a varying number of them");
      t.i_offset = eo.offset, t.i_e_out = eo.x, t.i_pdf = eo.pdf, t.i_cdf = eo.cdf;
      t.i_mu.assign(mu.data(), mu.data() + mu.shape()[1]);
    }
    else if (type == "incoherent_inelastic_discrete")
    {
      t.inelastic_type = MCW_SAB_INELASTIC_DISCRETE;
      const xt::xarray<double> eo = array2d(d, "energy_out"), mu = array2d(d, "mu_out");
      t.n_e_out = (int)eo.shape()[1], t.n_mu = (int)mu.shape()[2];
      double skewed = 0;
      openmc::read_dataset(d, "skewed", skewed);
      t.skewed = skewed != 0;
      t.i_e_out.assign(eo.data(), eo.data() + eo.size());
      t.i_mu.assign(mu.data(), mu.data() + mu.size());
    }
    else
      refuse("thermal inelastic distribution type '" + type + "'");
    openmc::close_group(d);
    openmc::close_group(in);
    if (openmc::object_exists(tg, "elastic"))
    {
      hid_t el = openmc::open_group(tg, "elastic");
      hid_t ed = openmc::open_group(el, "distribution");
      std::string etype;
      openmc::read_attribute(ed, "type", etype);
      if (etype == "coherent_elastic")
      {
        t.elastic_type = MCW_SAB_ELASTIC_COHERENT;
        const xt::xarray<double> v = array2d(el, "xs");
        const size_t n = v.shape()[1];
        t.e_x.assign(v.data(), v.data() + n), t.e_y.assign(v.data() + n, v.data() + 2 * n);
      }
      else if (etype == "incoherent_elastic_discrete")
      {
        t.elastic_type = MCW_SAB_ELASTIC_INCOHERENT_DISCRETE;
        const Function1D f = readFunction(el, "xs");
        t.e_x = f.x, t.e_y = f.y;
        const xt::xarray<double> mu = array2d(ed, "mu_out");
        t.n_elastic_mu = (int)mu.shape()[1];
        t.e_mu.assign(mu.data(), mu.data() + mu.size());
      }
      else
        refuse("thermal elastic distribution type '" + etype + "' (the Debye-Waller form is not carried)");
      openmc::close_group(ed);
      openmc::close_group(el);
    }
    openmc::close_group(tg);
    out.push_back({kt.first, std::move(t)});
  }
  return out;
}

// Thermal tables in the order of openmc::data::thermal_scatt: fills Library::thermal (every table's lowest temperature
// first -- what materials name --, then the further temperatures in table order), thermal_same and thermal_kT
inline void readThermalTables(const std::vector<hid_t>& groups, Library& lib)
{
  std::vector<std::vector<std::pair<double, ThermalTable>>> all;
  for (hid_t g : groups)
    all.push_back(readThermalTable(g));
  bool several = false;
  for (const auto& t : all)
    several |= t.size() > 1;
  for (size_t k = 0; k < all.size(); ++k)
  {
    lib.thermal.push_back(all[k][0].second);
    lib.thermal_same.push_back((int32_t)k), lib.thermal_kT.push_back(all[k][0].first);
  }
  for (size_t k = 0; k < all.size(); ++k)
    for (size_t j = 1; j < all[k].size(); ++j)
    {
      lib.thermal.push_back(all[k][j].second);
      lib.thermal_same.push_back((int32_t)k), lib.thermal_kT.push_back(all[k][j].first);
    }
  if (!several)
    lib.thermal_same.clear(), lib.thermal_kT.clear();
}

// groups: the nuclide groups in the order of openmc::data::nuclides (the order materials refer to)
inline Library readLibrary(const std::vector<std::pair<std::string, hid_t>>& groups)
{
  Library lib;
  std::vector<GatheredNuclide> all;
  for (const auto& ng : groups)
    all.push_back(readNuclide(ng.second, ng.first));
  // library entries: the nuclides, then their further temperature sets in nuclide order
  int entry = (int)all.size();
  for (size_t i = 0; i < all.size(); ++i)
  {
    std::vector<int> entry_of(1, (int)i);
    for (size_t k = 0; k < all[i].more.size(); ++k)
      entry_of.push_back(entry++);
    for (auto& u : all[i].urr)
    {
      u.second.nuclide = entry_of.at(u.first);
      lib.urr.push_back(std::move(u.second));
    }
  }
  for (size_t i = 0; i < all.size(); ++i)
  {
    lib.nuclides.push_back(std::move(all[i].base));
    for (auto& m : all[i].more)
      lib.temperature_sets.push_back({(int)i, std::move(m)});
  }
  std::sort(lib.urr.begin(), lib.urr.end(), [](const UrrTable& a, const UrrTable& b) { return a.nuclide < b.nuclide; });
  return lib;
}
} // namespace omcfile
} // namespace macaw
