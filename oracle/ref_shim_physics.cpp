// ref_shim_physics.cpp -- extern "C" doorways into the REFERENCE's own cross-section lookup and collision physics:
// Nuclide / Reaction / ReactionProduct / distribution objects built by their own constructors (openmc/src/nuclide.cpp:51,
// reaction.cpp:22, reaction_product.cpp:24, secondary_uncorrelated.cpp:17, distribution_energy.cpp:42,271,328,
// endf.cpp:80-151) from an in-memory tree (ref_hdf5_mem.*) laid out like an OpenMC HDF5 nuclide file; Nuclide::init_grid
// (nuclide.cpp:446); Material::calculate_xs (material.cpp:773-829) -> Nuclide::calculate_xs (nuclide.cpp:550-750); and
// Particle::event_collide (particle.cpp:219-288) -> collision (physics.cpp:38) -> sample_neutron_reaction (:86) with
// everything below it (sample_nuclide, sample_fission, create_fission_sites, absorption, scatter, elastic / inelastic
// scatter, sample_fission_neutron).  That is the chain MaCaw's CollisionKernel.C:143,185 runs.  Linked into
// oracle/_ref/libomc_ref.so by oracle/Makefile; test infrastructure only.  Nothing here restates an algorithm: the shim
// only lays arrays out as the data file would and sets / reads particle state.
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "openmc/bank.h"
#include "openmc/cell.h"
#include "openmc/message_passing.h"
#include "openmc/random_lcg.h"
#include "openmc/geometry.h"
#include "openmc/tallies/filter.h"
#include "openmc/tallies/filter_energy.h"
#include "openmc/tallies/tally.h"
#include "openmc/tallies/tally_scoring.h"
#include "openmc/constants.h"
#include "openmc/material.h"
#include "openmc/nuclide.h"
#include "openmc/particle.h"
#include "openmc/settings.h"
#include "openmc/simulation.h"
#include "openmc/thermal.h"
#include "ref_hdf5_mem.h"
#include "../include/macaw_b200.h" // mcw_tabular2d: the tables arrive in the product's own descriptor layout

namespace
{
std::unique_ptr<memh5::Node> g_root;
// tables staged by ref_stage_tables for the next ref_add_nuclide (null = none)
const mcw_tabular2d *g_el_angle = nullptr, *g_inel_angle = nullptr, *g_inel_energy = nullptr;
int g_inel_in_cm = 1;
const mcw_urr_desc* g_urr = nullptr; // probability tables staged for the next ref_add_nuclide
std::vector<std::pair<const mcw_nuclide_desc*, const mcw_urr_desc*>> g_sets; // further temperatures of the next nuclide
std::vector<std::pair<const mcw_thermal_desc*, double>> g_thermal_sets; // further temperatures of the next thermal table
const mcw_nuclide_desc* g_law = nullptr; // product distribution (correlated / Kalbach-Mann / N-body) of the next nuclide's inelastic reaction

// (3, N) dataset x / pdf / cdf with `offsets` and `interpolation` attributes: the layout of an angular distribution's `mu`
// (OMC distribution_angle.cpp:26-33) and of a continuous law's `distribution` (distribution_energy.cpp:80-92)
memh5::Node* table_dataset(memh5::Node* parent, const char* name, const mcw_tabular2d& t)
{
  const int n = t.n_energy, total = t.offset[n];
  std::vector<double> flat(3 * (size_t)total);
  for (int k = 0; k < total; ++k)
  {
    flat[k] = t.x[k];
    flat[total + k] = t.pdf[k];
    flat[2 * (size_t)total + k] = t.cdf[k];
  }
  memh5::Node* d = parent->dataset(name, flat, {3, (unsigned long long)total});
  d->attr("offsets", std::vector<int>(t.offset, t.offset + n));
  d->attr("interpolation", std::vector<int>(t.interpolation, t.interpolation + n));
  return d;
}
void angle_group(memh5::Node* distribution, const mcw_tabular2d& t)
{
  memh5::Node* a = distribution->group("angle");
  a->dataset("energy", std::vector<double>(t.energy, t.energy + t.n_energy));
  table_dataset(a, "mu", t);
}

void constant_function(memh5::Node* parent, const char* name, double v, double e0, double e1)
{
  // Tabulated1D, one lin-lin region, two points of equal value
  memh5::Node* d = parent->dataset(name, {e0, e1, v, v}, {2, 2});
  d->attr("type", "Tabulated1D");
  d->attr("breakpoints", std::vector<int>{2});
  d->attr("interpolation", std::vector<int>{2});
}
// a mcw_function1d as the dataset endf.cpp read_function reads: Polynomial (coefficients) or Tabulated1D ((2, n) x / y with
// one interpolation region)
void function_dataset(memh5::Node* parent, const char* name, const mcw_function1d& f)
{
  if (f.kind == MCW_FN_POLYNOMIAL)
  {
    memh5::Node* d = parent->dataset(name, std::vector<double>(f.y, f.y + f.n));
    d->attr("type", "Polynomial");
    return;
  }
  std::vector<double> xy(f.x, f.x + f.n);
  xy.insert(xy.end(), f.y, f.y + f.n);
  memh5::Node* d = parent->dataset(name, xy, {2, (unsigned long long)f.n});
  d->attr("type", "Tabulated1D");
  d->attr("breakpoints", std::vector<int>{f.n});
  d->attr("interpolation", std::vector<int>{f.interpolation});
}
void polynomial(memh5::Node* parent, const char* name, const std::vector<double>& coef)
{
  memh5::Node* d = parent->dataset(name, coef);
  d->attr("type", "Polynomial");
}
memh5::Node* reaction(memh5::Node* rxs, int mt, double q, bool cm, const std::string& temp, const double* xs, int n, int threshold)
{
  char nm[32];
  std::snprintf(nm, sizeof(nm), "reaction_%03d", mt);
  memh5::Node* r = rxs->group(nm);
  r->attr("Q_value", q);
  r->attr("mt", mt);
  r->attr("center_of_mass", cm ? 1 : 0);
  r->attr("redundant", 0);
  memh5::Node* d = r->group(temp)->dataset("xs", std::vector<double>(xs, xs + n));
  d->attr("threshold_idx", threshold);
  return r;
}
// the cross section of an existing reaction group at one more temperature (Reaction::Reaction reads `<T>K/xs` for every
// temperature of the nuclide, reaction.cpp:25-50)
void reaction_temperature(memh5::Node* r, const std::string& temp, const double* xs, int n, int threshold)
{
  memh5::Node* d = r->group(temp)->dataset("xs", std::vector<double>(xs, xs + n));
  d->attr("threshold_idx", threshold);
}
std::string temperature_name(double kT) { return std::to_string((int)std::round(kT / openmc::K_BOLTZMANN)) + "K"; }
memh5::Node* neutron_product(memh5::Node* rx, int k, const char* mode, const std::vector<double>& yield_coef)
{
  memh5::Node* p = rx->group("product_" + std::to_string(k));
  p->attr("particle", "neutron");
  p->attr("emission_mode", mode);
  p->attr("n_distribution", 1);
  polynomial(p, "yield", yield_coef);
  memh5::Node* d = p->group("distribution_0");
  d->attr("type", "uncorrelated"); // no "angle" group: isotropic (secondary_uncorrelated.cpp:60-63)
  return d;
}
// What the distribution of a reaction's neutron product is made from (either descriptor)
struct LawView
{
  int law;
  const mcw_tabular2d *angle, *energy;
  const int32_t *corr_offset, *corr_interp;
  const double *corr_mu, *corr_pdf, *corr_cdf, *km_r, *km_a;
  int nbody_n;
  double nbody_mass_ratio, nbody_q, level_E, level_ratio;
  double u;
  const mcw_function1d *a, *b;
};
// distribution_0 of a neutron product as the reference's constructors read it: "uncorrelated" (angle table or none +
// level / continuous energy law: secondary_uncorrelated.cpp:17-52), "correlated" (secondary_correlated.cpp:22-150),
// "kalbach-mann" (secondary_kalbach.cpp:24-110), "nbody" (secondary_nbody.cpp:16-23)
void fill_distribution(memh5::Node* dist, double awr, const LawView& L)
{
  const bool together = L.law >= MCW_LAW_CORRELATED && L.law <= MCW_LAW_NBODY; // energy and cosine from one distribution
  if (!together && L.angle)
    angle_group(dist, *L.angle);
  if (L.law == MCW_LAW_NBODY)
  {
    dist->attr("type", "nbody");
    dist->attr("n_particles", L.nbody_n);
    dist->attr("total_mass", L.nbody_mass_ratio);
    dist->attr("atomic_weight_ratio", awr);
    dist->attr("q_value", L.nbody_q);
  }
  else if (together)
  {
    const mcw_tabular2d& t = *L.energy;
    const int n = t.n_energy, np = t.offset[n];
    memh5::Node* en = dist->dataset("energy", std::vector<double>(t.energy, t.energy + n));
    en->attr("interpolation", std::vector<int>{n, 2}, {2, 1});
    std::vector<double> eo(5 * (size_t)np);
    for (int k = 0; k < np; ++k)
    {
      eo[k] = t.x[k], eo[np + k] = t.pdf[k], eo[2 * (size_t)np + k] = t.cdf[k];
      if (L.law == MCW_LAW_CORRELATED)
        eo[3 * (size_t)np + k] = L.corr_interp[k], eo[4 * (size_t)np + k] = L.corr_offset[k];
      else
        eo[3 * (size_t)np + k] = L.km_r[k], eo[4 * (size_t)np + k] = L.km_a[k];
    }
    memh5::Node* d = dist->dataset(L.law == MCW_LAW_CORRELATED ? "energy_out" : "distribution", eo, {5, (unsigned long long)np});
    d->attr("offsets", std::vector<int>(t.offset, t.offset + n));
    d->attr("interpolation", std::vector<int>(t.interpolation, t.interpolation + n));
    d->attr("n_discrete_lines", std::vector<int>(n, 0));
    if (L.law == MCW_LAW_CORRELATED)
    {
      dist->attr("type", "correlated");
      const int nm = L.corr_offset[np];
      std::vector<double> mu(3 * (size_t)nm);
      for (int k = 0; k < nm; ++k)
        mu[k] = L.corr_mu[k], mu[nm + k] = L.corr_pdf[k], mu[2 * (size_t)nm + k] = L.corr_cdf[k];
      dist->dataset("mu", mu, {3, (unsigned long long)nm});
    }
    else
      dist->attr("type", "kalbach-mann");
  }
  if (together)
    return; // the product distribution above replaces the uncorrelated angle + energy pair
  memh5::Node* e = dist->group("energy");
  if (L.law == MCW_LAW_CONTINUOUS_TABULAR)
  {
    // ContinuousTabular (OMC distribution_energy.cpp:56-148): `energy` with one incoming-energy interpolation region
    const mcw_tabular2d& t = *L.energy;
    e->attr("type", "continuous");
    memh5::Node* en = e->dataset("energy", std::vector<double>(t.energy, t.energy + t.n_energy));
    en->attr("interpolation", std::vector<int>{t.n_energy, t.histogram_in_energy ? 1 : 2}, {2, 1});
    memh5::Node* d = table_dataset(e, "distribution", t);
    std::vector<int> nd(t.n_energy, 0);
    if (t.n_discrete)
      nd.assign(t.n_discrete, t.n_discrete + t.n_energy);
    d->attr("n_discrete_lines", nd);
  }
  else if (L.law >= MCW_LAW_MAXWELL)
  {
    // MaxwellEnergy / Evaporation / WattEnergy of a scattering product (distribution_energy.cpp:265-356)
    e->attr("type", L.law == MCW_LAW_WATT ? "watt" : (L.law == MCW_LAW_MAXWELL ? "maxwell" : "evaporation"));
    e->attr("u", L.u);
    if (L.law == MCW_LAW_WATT)
    {
      function_dataset(e, "a", *L.a);
      function_dataset(e, "b", *L.b);
    }
    else
      function_dataset(e, "theta", *L.a);
  }
  else
  {
    e->attr("type", "level");
    e->attr("threshold", L.level_E);
    e->attr("mass_ratio", L.level_ratio);
  }
}
} // namespace

// the image's root for oracle/ref_gather.cpp (host/openmc_nuclide_reader.hpp reads the same image)
memh5::Node* ref_image_root() { return g_root.get(); }

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API void ref_library_reset()
{
  using namespace openmc;
  model::materials.clear();
  data::nuclides.clear();
  data::nuclide_map.clear();
  data::thermal_scatt.clear();
  data::thermal_scatt_map.clear();
  g_root.reset(new memh5::Node);
  g_sets.clear();
  g_thermal_sets.clear();
  settings::run_CE = true;
  settings::run_mode = RunMode::EIGENVALUE;
  settings::survival_biasing = false;
  settings::ufs_on = false;
  settings::photon_transport = false;
  settings::res_scat_on = false;
  settings::urr_ptables_on = true;
  settings::temperature_method = TemperatureMethod::NEAREST;
  model::n_coord_levels = 1;
  model::active_collision_tallies.clear();
  model::active_analog_tallies.clear();
  model::active_tallies.clear();
  model::active_surface_tallies.clear();
  model::active_meshsurf_tallies.clear();
}

// Tabulated secondary distributions of the NEXT nuclide added (null = none): angular distributions of the elastic and
// the inelastic reaction (read by AngleDistribution's own constructor), a continuous tabular energy law instead of the
// level law (read by ContinuousTabular's own constructor) and that reaction's centre-of-mass flag.
REF_API void ref_stage_tables(const mcw_tabular2d* elastic_angle, const mcw_tabular2d* inelastic_angle,
                              const mcw_tabular2d* inelastic_energy, int inelastic_in_cm)
{
  g_el_angle = elastic_angle && elastic_angle->n_energy ? elastic_angle : nullptr;
  g_inel_angle = inelastic_angle && inelastic_angle->n_energy ? inelastic_angle : nullptr;
  g_inel_energy = inelastic_energy && inelastic_energy->n_energy ? inelastic_energy : nullptr;
  g_inel_in_cm = inelastic_in_cm;
}

// The NEXT nuclide's own descriptor, for what ref_add_nuclide's argument list does not carry: the product distribution of
// its inelastic reaction when that gives energy and cosine together -- written as the "correlated" / "kalbach-mann" /
// "nbody" group the reference's own CorrelatedAngleEnergy / KalbachMann / NBodyPhaseSpace constructors read
// (secondary_correlated.cpp:22-150, secondary_kalbach.cpp:24-110, secondary_nbody.cpp:16-23) --, that reaction's yield,
// and the scattering reactions after it (more_reactions: one reaction_<MT> group each).
REF_API void ref_stage_law(const mcw_nuclide_desc* d) { g_law = d; }

// Probability tables of the NEXT nuclide added: written as the `urr/<T>K` group UrrData's own constructor reads
// (urr.cpp:7-33: attributes interpolation / inelastic / absorption / multiply_smooth, datasets energy and table)
REF_API void ref_stage_urr(const mcw_urr_desc* urr) { g_urr = urr; }

// One more temperature of the NEXT nuclide added (call once per further temperature): its own energy grid and every
// reaction's cross section on it, written as the `<T>K` members the reference's Nuclide / Reaction / UrrData constructors
// read for each temperature of `kTs` (nuclide.cpp:170-200, 225-237).  urr: that temperature's probability tables or null.
REF_API void ref_stage_temperature_set(const mcw_nuclide_desc* set, const mcw_urr_desc* urr) { g_sets.push_back({set, urr}); }

// One synthetic nuclide of macaw_b200/nucdata.py as a nuclide file: MT 2 elastic (isotropic in the centre of mass),
// optionally MT 51 (one level, isotropic CM), MT 102 capture, optionally MT 18 fission with a polynomial total nu, a Watt
// prompt spectrum and n_precursor Maxwell delayed groups of constant yield.  Returns the nuclide's index.
REF_API int ref_add_nuclide(const char* name, double awr, double kT, int G, const double* energy, const double* elastic,
                            const double* capture, const double* fission, int n_inel, int inel_thr, const double* inel,
                            double inel_E, double inel_ratio, const double* nu_poly, int n_precursor, const double* dyield,
                            const double* dtheta, double watt_a, double watt_b, double q_fission)
{
  using namespace openmc;
  memh5::Node* g = g_root->group(name);
  g->attr("Z", 1);
  g->attr("A", 1);
  g->attr("metastable", 0);
  g->attr("atomic_weight_ratio", awr);
  const int T = (int)std::round(kT / K_BOLTZMANN);
  const std::string temp = std::to_string(T) + "K";
  g->group("kTs")->dataset(temp, {kT}, {})->data.shape.clear();
  g->group("energy")->dataset(temp, std::vector<double>(energy, energy + G));
  memh5::Node* rxs = g->group("reactions");
  memh5::Node* r_el = reaction(rxs, 2, 0.0, true, temp, elastic, G, 0);
  memh5::Node *r_inel = nullptr, *r_cap = nullptr, *r_fis = nullptr;
  std::vector<memh5::Node*> r_more;
  memh5::Node* del = neutron_product(r_el, 0, "prompt", {1.0});
  if (g_el_angle)
    angle_group(del, *g_el_angle);
  if (n_inel > 0)
  {
    const double y0 = g_law && g_law->inelastic_yield != 0.0 ? g_law->inelastic_yield : 1.0;
    const bool cm0 = g_law ? (g_law->inelastic_law == MCW_LAW_LEVEL ? true : g_law->inelastic_in_cm != 0)
                           : (g_inel_energy ? g_inel_in_cm != 0 : true);
    r_inel = reaction(rxs, 51, -inel_E, cm0, temp, inel, n_inel, inel_thr);
    memh5::Node* dist = neutron_product(r_inel, 0, "prompt", {y0});
    LawView v{};
    v.law = g_law ? g_law->inelastic_law : (g_inel_energy ? MCW_LAW_CONTINUOUS_TABULAR : MCW_LAW_LEVEL);
    v.angle = g_inel_angle, v.energy = g_inel_energy;
    if (g_law)
    {
      const mcw_nuclide_desc& L = *g_law;
      if (L.inelastic_energy.n_energy)
        v.energy = &L.inelastic_energy;
      if (L.inelastic_angle.n_energy)
        v.angle = &L.inelastic_angle;
      v.corr_offset = L.corr_offset, v.corr_interp = L.corr_interp, v.corr_mu = L.corr_mu, v.corr_pdf = L.corr_pdf;
      v.corr_cdf = L.corr_cdf, v.km_r = L.km_r, v.km_a = L.km_a, v.nbody_n = L.nbody_n;
      v.nbody_mass_ratio = L.nbody_mass_ratio, v.nbody_q = L.nbody_q;
    }
    v.level_E = inel_E, v.level_ratio = inel_ratio;
    if (g_law)
      v.u = g_law->inelastic_u, v.a = &g_law->inelastic_a, v.b = &g_law->inelastic_b;
    fill_distribution(dist, awr, v);
    if (g_law && g_law->inelastic_yield_fn.kind != MCW_FN_NONE) // an energy-dependent multiplicity replaces the constant
    {
      r_inel->children.at("product_0")->children.erase("yield");
      function_dataset(r_inel->children.at("product_0").get(), "yield", g_law->inelastic_yield_fn);
    }
    // the scattering reactions after the first (index_inelastic_scatter_ follows the MT order of the group names)
    for (int j = 0; g_law && j < g_law->n_more_reactions; ++j)
    {
      const mcw_reaction_desc& r = g_law->more_reactions[j];
      r_more.push_back(reaction(rxs, r.mt, -r.threshold_energy, r.law == MCW_LAW_LEVEL ? true : r.in_cm != 0, temp, r.xs, G - r.threshold,
                                r.threshold));
      memh5::Node* dj = neutron_product(r_more.back(), 0, "prompt", {r.yield == 0.0 ? 1.0 : r.yield});
      LawView w{};
      w.law = r.law;
      w.angle = r.angle.n_energy ? &r.angle : nullptr, w.energy = r.energy.n_energy ? &r.energy : nullptr;
      w.corr_offset = r.corr_offset, w.corr_interp = r.corr_interp, w.corr_mu = r.corr_mu, w.corr_pdf = r.corr_pdf;
      w.corr_cdf = r.corr_cdf, w.km_r = r.km_r, w.km_a = r.km_a, w.nbody_n = r.nbody_n;
      w.nbody_mass_ratio = r.nbody_mass_ratio, w.nbody_q = r.nbody_q;
      w.level_E = r.threshold_energy, w.level_ratio = r.mass_ratio;
      w.u = r.u, w.a = &r.a, w.b = &r.b;
      fill_distribution(dj, awr, w);
      if (r.yield_fn.kind != MCW_FN_NONE)
      {
        r_more.back()->children.at("product_0")->children.erase("yield");
        function_dataset(r_more.back()->children.at("product_0").get(), "yield", r.yield_fn);
      }
    }
  }
  g_el_angle = g_inel_angle = g_inel_energy = nullptr;
  const mcw_nuclide_desc* g_fis = g_law; // the staged descriptor also carries the fission products
  g_law = nullptr;
  if (g_urr)
  {
    memh5::Node* u = g->group("urr")->group(temp);
    u->attr("interpolation", g_urr->interpolation);
    u->attr("inelastic", g_urr->inelastic_flag);
    u->attr("absorption", g_urr->absorption_flag);
    u->attr("multiply_smooth", g_urr->multiply_smooth);
    const unsigned long long ne = g_urr->n_energy, nb = g_urr->n_band;
    u->dataset("energy", std::vector<double>(g_urr->energy, g_urr->energy + ne));
    u->dataset("table", std::vector<double>(g_urr->table, g_urr->table + ne * 6 * nb), {ne, 6, nb});
    g_urr = nullptr;
  }
  r_cap = reaction(rxs, 102, 0.0, false, temp, capture, G, 0);
  if (fission)
  {
    memh5::Node* f = r_fis = reaction(rxs, 18, q_fission, false, temp, fission, G, 0);
    if (g_fis && g_fis->n_fission_products > 0)
    {
      // fission neutrons as a nuclide file has them: every product with its yield function and its own energy law, read by
      // the reference's ReactionProduct / WattEnergy / MaxwellEnergy / Evaporation / ContinuousTabular constructors
      for (int k = 0; k < g_fis->n_fission_products; ++k)
      {
        const mcw_fission_product& P = g_fis->fission_products[k];
        memh5::Node* p = f->group("product_" + std::to_string(k));
        p->attr("particle", "neutron");
        p->attr("emission_mode", k == 0 ? "prompt" : "delayed");
        if (k > 0)
          p->attr("decay_rate", 1.0);
        p->attr("n_distribution", 1);
        function_dataset(p, "yield", P.yield);
        memh5::Node* d = p->group("distribution_0");
        d->attr("type", "uncorrelated");
        memh5::Node* en = d->group("energy");
        if (P.law == MCW_LAW_CONTINUOUS_TABULAR)
        {
          const mcw_tabular2d& t = P.energy;
          en->attr("type", "continuous");
          memh5::Node* ee = en->dataset("energy", std::vector<double>(t.energy, t.energy + t.n_energy));
          ee->attr("interpolation", std::vector<int>{t.n_energy, t.histogram_in_energy ? 1 : 2}, {2, 1});
          memh5::Node* dd = table_dataset(en, "distribution", t);
          std::vector<int> nd(t.n_energy, 0);
          if (t.n_discrete)
            nd.assign(t.n_discrete, t.n_discrete + t.n_energy);
          dd->attr("n_discrete_lines", nd);
        }
        else
        {
          en->attr("type", P.law == MCW_LAW_WATT ? "watt" : (P.law == MCW_LAW_MAXWELL ? "maxwell" : "evaporation"));
          en->attr("u", P.u);
          if (P.law == MCW_LAW_WATT)
          {
            function_dataset(en, "a", P.a);
            function_dataset(en, "b", P.b);
          }
          else
            function_dataset(en, "theta", P.a);
        }
      }
      if (g_fis->total_nu.kind != MCW_FN_NONE)
        function_dataset(g->group("total_nu"), "yield", g_fis->total_nu);
    }
    else
    {
    const double e0 = energy[0], e1 = energy[G - 1];
    double nu_d = 0.0;
    for (int k = 0; k < n_precursor; ++k)
      nu_d += dyield[k];
    memh5::Node* pe = neutron_product(f, 0, "prompt", {nu_poly[0] - nu_d, nu_poly[1]})->group("energy");
    pe->attr("type", "watt");
    pe->attr("u", -2.0e7);
    constant_function(pe, "a", watt_a, e0, e1);
    constant_function(pe, "b", watt_b, e0, e1);
    for (int k = 0; k < n_precursor; ++k)
    {
      memh5::Node* dp = neutron_product(f, k + 1, "delayed", {dyield[k]});
      memh5::Node* de = dp->group("energy");
      de->attr("type", "maxwell");
      de->attr("u", -2.0e7);
      constant_function(de, "theta", dtheta[k], e0, e1);
    }
    for (int k = 0; k < n_precursor; ++k)
      f->children.at("product_" + std::to_string(k + 1))->attr("decay_rate", 1.0);
    polynomial(g->group("total_nu"), "yield", {nu_poly[0], nu_poly[1]});
    }
  }
  // further temperatures: kTs / energy / every reaction's xs (/ urr) under their own `<T>K`
  std::vector<double> temps{(double)T};
  for (const auto& su : g_sets)
  {
    const mcw_nuclide_desc& S = *su.first;
    const std::string tn = temperature_name(S.kT);
    temps.push_back(std::round(S.kT / K_BOLTZMANN));
    const int Gs = S.n_grid;
    g->children.at("kTs")->dataset(tn, {S.kT}, {})->data.shape.clear();
    g->children.at("energy")->dataset(tn, std::vector<double>(S.energy, S.energy + Gs));
    reaction_temperature(r_el, tn, S.elastic, Gs, 0);
    if (r_inel)
      reaction_temperature(r_inel, tn, S.inelastic, Gs - S.inel_threshold, S.inel_threshold);
    for (size_t j = 0; j < r_more.size(); ++j)
      reaction_temperature(r_more[j], tn, S.more_reactions[j].xs, Gs - S.more_reactions[j].threshold, S.more_reactions[j].threshold);
    std::vector<double> cap(Gs), fis(Gs);
    for (int k = 0; k < Gs; ++k)
      fis[k] = S.xs[5 * (size_t)k + 2], cap[k] = S.xs[5 * (size_t)k + 1] - S.xs[5 * (size_t)k + 2];
    reaction_temperature(r_cap, tn, cap.data(), Gs, 0);
    if (r_fis)
      reaction_temperature(r_fis, tn, fis.data(), Gs, 0);
    if (su.second)
    {
      const mcw_urr_desc& U = *su.second;
      memh5::Node* u = g->group("urr")->group(tn);
      u->attr("interpolation", U.interpolation);
      u->attr("inelastic", U.inelastic_flag);
      u->attr("absorption", U.absorption_flag);
      u->attr("multiply_smooth", U.multiply_smooth);
      const unsigned long long ne = U.n_energy, nb = U.n_band;
      u->dataset("energy", std::vector<double>(U.energy, U.energy + ne));
      u->dataset("table", std::vector<double>(U.table, U.table + ne * 6 * nb), {ne, 6, nb});
    }
  }
  g_sets.clear();
  data::nuclides.push_back(std::make_unique<Nuclide>(memh5::id_of(g), temps));
  return (int)data::nuclides.size() - 1;
}

// One S(a,b) table at the temperature of kT as a thermal-scattering file image, read by the reference's own
// ThermalScattering / ThermalData constructors (thermal.cpp:38-227) -- and through them Tabulated1D, CoherentElasticXS,
// CorrelatedAngleEnergy -> IncoherentInelasticAE, IncoherentInelasticAEDiscrete, CoherentElasticAE,
// IncoherentElasticAEDiscrete.  Returns its index in data::thermal_scatt.
// One more temperature of the NEXT thermal table added (ThermalScattering reads every `<T>K` group of `kTs`)
REF_API void ref_stage_thermal_temperature(const mcw_thermal_desc* t, double kT) { g_thermal_sets.push_back({t, kT}); }

REF_API int ref_add_thermal_table(const char* name, const mcw_thermal_desc* t0, double kT0)
{
  using namespace openmc;
  auto tabulated = [](memh5::Node* parent, const double* x, const double* y, int n) {
    std::vector<double> flat(2 * (size_t)n);
    for (int k = 0; k < n; ++k)
      flat[k] = x[k], flat[n + k] = y[k];
    memh5::Node* d = parent->dataset("xs", flat, {2, (unsigned long long)n});
    d->attr("type", "Tabulated1D");
    d->attr("breakpoints", std::vector<int>{n});
    d->attr("interpolation", std::vector<int>{2});
    return d;
  };
  memh5::Node* g = g_root->group(name);
  g->attr("atomic_weight_ratio", 0.999167);
  g->attr("energy_max", t0->energy_max);
  g->attrs["nuclides"] = memh5::Array{memh5::Array::TEXT, {1}, {}, {}, "H1"};
  std::vector<std::pair<const mcw_thermal_desc*, double>> all{{t0, kT0}};
  all.insert(all.end(), g_thermal_sets.begin(), g_thermal_sets.end());
  g_thermal_sets.clear();
  std::vector<double> temps;
  for (const auto& tk : all)
  {
  const mcw_thermal_desc* t = tk.first;
  const double kT = tk.second;
  const int T = (int)std::round(kT / K_BOLTZMANN);
  temps.push_back((double)T);
  const std::string temp = std::to_string(T) + "K";
  g->group("kTs")->dataset(temp, {kT}, {})->data.shape.clear();
  memh5::Node* tg = g->group(temp);
  // ---- inelastic
  memh5::Node* in = tg->group("inelastic");
  tabulated(in, t->inelastic_xs_energy, t->inelastic_xs, t->n_inelastic_xs);
  memh5::Node* d = in->group("distribution");
  if (t->inelastic_type == MCW_SAB_INELASTIC_CONTINUOUS)
  {
    // the correlated angle-energy layout (secondary_correlated.cpp:22-150): energy_out rows = E_out, pdf, cdf, cosine
    // interpolation (0 = discrete cosines), offset into mu; mu rows = cosine, pdf, cdf
    d->attr("type", "incoherent_inelastic");
    const int n_in = t->n_inelastic_energy, np = t->inelastic_offset[n_in], n_mu = t->n_mu;
    memh5::Node* en = d->dataset("energy", std::vector<double>(t->inelastic_energy, t->inelastic_energy + n_in));
    en->attr("interpolation", std::vector<int>{n_in, 2}, {2, 1});
    std::vector<double> eo(5 * (size_t)np), mu(3 * (size_t)np * n_mu);
    for (int k = 0; k < np; ++k)
    {
      eo[k] = t->inelastic_e_out[k];
      eo[np + k] = t->inelastic_pdf[k];
      eo[2 * (size_t)np + k] = t->inelastic_cdf[k];
      eo[3 * (size_t)np + k] = 0.0;
      eo[4 * (size_t)np + k] = (double)k * n_mu;
      for (int m = 0; m < n_mu; ++m)
      {
        const size_t c = (size_t)k * n_mu + m, M = (size_t)np * n_mu;
        mu[c] = t->inelastic_mu[c];
        mu[M + c] = 1.0 / n_mu;
        mu[2 * M + c] = (m + 1.0) / n_mu;
      }
    }
    memh5::Node* eod = d->dataset("energy_out", eo, {5, (unsigned long long)np});
    eod->attr("offsets", std::vector<int>(t->inelastic_offset, t->inelastic_offset + n_in));
    eod->attr("interpolation", std::vector<int>(n_in, 2));
    eod->attr("n_discrete_lines", std::vector<int>(n_in, 0));
    d->dataset("mu", mu, {3, (unsigned long long)np * n_mu});
  }
  else
  {
    d->attr("type", "incoherent_inelastic_discrete");
    const unsigned long long n_in = t->n_inelastic_xs, n_out = t->n_e_out, n_mu = t->n_mu;
    d->dataset("energy_out", std::vector<double>(t->inelastic_e_out, t->inelastic_e_out + n_in * n_out), {n_in, n_out});
    d->dataset("mu_out", std::vector<double>(t->inelastic_mu, t->inelastic_mu + n_in * n_out * n_mu), {n_in, n_out, n_mu});
    d->dataset("skewed", {(double)(t->skewed != 0)}, {})->data.shape.clear();
  }
  // ---- elastic
  if (t->elastic_type != MCW_SAB_ELASTIC_NONE)
  {
    memh5::Node* el = tg->group("elastic");
    memh5::Node* ed = el->group("distribution");
    const int n = t->n_elastic;
    if (t->elastic_type == MCW_SAB_ELASTIC_COHERENT)
    {
      std::vector<double> flat(2 * (size_t)n);
      for (int k = 0; k < n; ++k)
        flat[k] = t->elastic_x[k], flat[n + k] = t->elastic_y[k];
      el->dataset("xs", flat, {2, (unsigned long long)n})->attr("type", "CoherentElastic");
      ed->attr("type", "coherent_elastic");
    }
    else
    {
      tabulated(el, t->elastic_x, t->elastic_y, n);
      ed->attr("type", "incoherent_elastic_discrete");
      ed->dataset("mu_out", std::vector<double>(t->elastic_mu, t->elastic_mu + (size_t)n * t->n_elastic_mu),
                  {(unsigned long long)n, (unsigned long long)t->n_elastic_mu});
    }
  }
  } // every temperature
  data::thermal_scatt.push_back(std::make_unique<ThermalScattering>(memh5::id_of(g), temps));
  return (int)data::thermal_scatt.size() - 1;
}

// <sab> entries of a material (Material::thermal_tables_, material.cpp:317-360): sorted by nuclide position
REF_API void ref_set_material_thermal(int i_mat, int n, const int32_t* position, const int32_t* table, const double* fraction)
{
  using namespace openmc;
  std::vector<Material::ThermalTable> tables;
  for (int k = 0; k < n; ++k)
    tables.push_back({table[k], position[k], fraction[k]});
  model::materials[i_mat]->thermal_tables_ = tables;
}

// energy limits of the library, then every nuclide's logarithmic grid index (what read_ce_cross_sections does after loading)
REF_API void ref_finish_library(int n_log_bins)
{
  using namespace openmc;
  const int neutron = static_cast<int>(ParticleType::neutron);
  data::energy_min[neutron] = 0.0;
  data::energy_max[neutron] = INFTY;
  for (const auto& nuc : data::nuclides)
  {
    data::energy_min[neutron] = std::max(data::energy_min[neutron], nuc->grid_[0].energy.front());
    data::energy_max[neutron] = std::min(data::energy_max[neutron], nuc->grid_[0].energy.back());
  }
  settings::n_log_bins = n_log_bins;
  simulation::log_spacing = std::log(data::energy_max[neutron] / data::energy_min[neutron]) / settings::n_log_bins;
  for (auto& nuc : data::nuclides)
    nuc->init_grid();
}

// derived tables of nuclide i (create_derived, nuclide.cpp:317-390): xs_[0] as [G][5], and its flags
// The in-memory nuclide file image as a byte stream (so that macaw_b200/openmc_hdf5.py, the converter from OpenMC's
// nuclide-file layout to the library's descriptors, can be run on exactly what the reference's constructors just read):
//   node  = 'G' | 'D', name, n_attrs, {name, array}, [array if 'D'], n_children, {node}
//   array = kind (0 real, 1 int, 2 text), n_dims, dims (u64), payload (f64 / i64 / bytes)
// strings are u32 length + bytes; everything little endian.
namespace
{
void put_u32(std::vector<unsigned char>& o, uint32_t v) { o.insert(o.end(), (unsigned char*)&v, (unsigned char*)&v + 4); }
void put_str(std::vector<unsigned char>& o, const std::string& s)
{
  put_u32(o, (uint32_t)s.size());
  o.insert(o.end(), s.begin(), s.end());
}
void put_array(std::vector<unsigned char>& o, const memh5::Array& a)
{
  o.push_back((unsigned char)a.kind);
  put_u32(o, (uint32_t)a.shape.size());
  for (unsigned long long d : a.shape)
    o.insert(o.end(), (unsigned char*)&d, (unsigned char*)&d + 8);
  if (a.kind == memh5::Array::REAL)
  {
    put_u32(o, (uint32_t)a.real.size());
    o.insert(o.end(), (const unsigned char*)a.real.data(), (const unsigned char*)(a.real.data() + a.real.size()));
  }
  else if (a.kind == memh5::Array::INT)
  {
    put_u32(o, (uint32_t)a.integer.size());
    o.insert(o.end(), (const unsigned char*)a.integer.data(), (const unsigned char*)(a.integer.data() + a.integer.size()));
  }
  else
    put_str(o, a.text);
}
void put_node(std::vector<unsigned char>& o, const memh5::Node& n)
{
  o.push_back(n.is_dataset ? 'D' : 'G');
  put_str(o, n.name);
  put_u32(o, (uint32_t)n.attrs.size());
  for (const auto& kv : n.attrs)
  {
    put_str(o, kv.first);
    put_array(o, kv.second);
  }
  if (n.is_dataset)
    put_array(o, n.data);
  put_u32(o, (uint32_t)n.children.size());
  for (const auto& kv : n.children)
    put_node(o, *kv.second);
}
std::vector<unsigned char> g_dump;
} // namespace

// serialises the tree; returns the byte count, copies min(count, capacity) bytes into out (call with capacity 0 first)
REF_API int64_t ref_dump_tree(unsigned char* out, int64_t capacity)
{
  if (capacity <= 0)
  {
    g_dump.clear();
    if (g_root)
      put_node(g_dump, *g_root);
    return (int64_t)g_dump.size();
  }
  const int64_t n = std::min<int64_t>(capacity, (int64_t)g_dump.size());
  memcpy(out, g_dump.data(), (size_t)n);
  return n;
}

// number of temperatures nuclide i was read with and kTs_[t]
REF_API int ref_nuclide_temperatures(int i, double* kTs)
{
  const auto& nuc = *openmc::data::nuclides[i];
  for (size_t t = 0; kTs && t < nuc.kTs_.size(); ++t)
    kTs[t] = nuc.kTs_[t];
  return (int)nuc.kTs_.size();
}

// ref_nuclide_tables for temperature index t (xs_[t], grid_[t].grid_index)
REF_API void ref_nuclide_tables_t(int i, int t, double* xs5, int32_t* grid_index)
{
  const auto& nuc = *openmc::data::nuclides[i];
  const auto& xs = nuc.xs_[t];
  for (size_t k = 0; k < xs.shape()[0]; ++k)
    for (int c = 0; c < 5; ++c)
      xs5[5 * k + c] = xs(k, c);
  if (grid_index)
    for (size_t k = 0; k < nuc.grid_[t].grid_index.size(); ++k)
      grid_index[k] = nuc.grid_[t].grid_index[k];
}

REF_API void ref_nuclide_tables(int i, double* xs5, int32_t* fissionable, int32_t* n_precursor, int32_t* grid_index)
{
  const auto& nuc = *openmc::data::nuclides[i];
  const auto& xs = nuc.xs_[0];
  for (size_t k = 0; k < xs.shape()[0]; ++k)
    for (int c = 0; c < 5; ++c)
      xs5[5 * k + c] = xs(k, c);
  *fissionable = nuc.fissionable_;
  *n_precursor = nuc.n_precursor_;
  if (grid_index)
    for (size_t k = 0; k < nuc.grid_[0].grid_index.size(); ++k)
      grid_index[k] = nuc.grid_[0].grid_index[k];
}

// isotropic: per-nuclide <isotropic> flags (material.cpp:255-270 p0_) or null
REF_API int ref_add_material(int n, const int32_t* nuc, const double* dens, const int32_t* isotropic)
{
  using namespace openmc;
  model::materials.push_back(std::make_unique<Material>());
  Material& m = *model::materials.back();
  m.nuclide_.assign(nuc, nuc + n);
  m.atom_density_ = xt::zeros<double>({(size_t)n});
  for (int k = 0; k < n; ++k)
    m.atom_density_(k) = dens[k];
  if (isotropic)
  {
    m.p0_.resize(n);
    for (int k = 0; k < n; ++k)
      m.p0_[k] = isotropic[k] != 0;
  }
  m.init_nuclide_index();
  return (int)model::materials.size() - 1;
}

// Replays n independent collision events in material i_mat: each starts from (E, u, seeds[4], id, wgt 1), gets the macroscopic
// cross sections from Material::calculate_xs and collides through Particle::event_collide.  Outputs per event: macro4 =
// (total, absorption, fission, nu_fission); state8 = (E, ux, uy, uz, alive, keff_tally_collision, keff_tally_absorption,
// sampled nuclide); seeds_out[4]; n_sites.  Fission sites of all events in order: site7 = (E, ux, uy, uz, delayed_group,
// parent_id, progeny_id).  kappa (or null): per event, what score_collision_tally (called by event_collide, particle.cpp:247)
// adds to a filterless collision-estimator tally with the kappa-fission score (tally_scoring.cpp:1216-1258).  Returns the
// number of banked sites.
REF_API int64_t ref_collide_events(int i_mat, double keff, int64_t n, const double* E, const double* u3, const uint64_t* seeds4,
                                   const int64_t* id, double sqrtkT, double* macro4, double* state8, uint64_t* seeds_out,
                                   int32_t* n_sites, int64_t bank_capacity, double* site7, double* kappa)
{
  using namespace openmc;
  Tally* tally = nullptr;
  model::tallies.clear();
  model::tally_map.clear();
  model::tally_filters.clear();
  model::filter_map.clear();
  if (kappa)
  {
    tally = Tally::create(1);
    tally->set_scores(std::vector<std::string>{"kappa-fission"});
    tally->estimator_ = TallyEstimator::COLLISION;
    tally->set_active(true);
    tally->init_results();
  }
  setup_active_tallies();
  simulation::keff = keff;
  simulation::fission_bank.clear();
  simulation::fission_bank.reserve(bank_capacity);
  for (int64_t k = 0; k < n; ++k)
  {
    Particle p;
    p.type() = ParticleType::neutron;
    p.n_coord() = 1;
    p.material() = i_mat;
    p.sqrtkT() = sqrtkT;
    p.wgt() = 1.0;
    p.id() = id[k];
    p.E() = E[k];
    p.u() = {u3[3 * k], u3[3 * k + 1], u3[3 * k + 2]};
    p.r() = {0.25 * k, -0.5, 1.0};
    for (int s = 0; s < N_STREAMS; ++s)
      p.seeds(s) = seeds4[4 * k + s];
    p.stream() = STREAM_TRACKING;
    // event_calculate_xs (particle.cpp:86-135) minus the cell search: remember the pre-collision state, look up
    p.E_last() = p.E();
    p.u_last() = p.u();
    p.r_last() = p.r();
    p.wgt_last() = p.wgt();
    model::materials[i_mat]->calculate_xs(p);
    macro4[4 * k] = p.macro_xs().total;
    macro4[4 * k + 1] = p.macro_xs().absorption;
    macro4[4 * k + 2] = p.macro_xs().fission;
    macro4[4 * k + 3] = p.macro_xs().nu_fission;
    const int64_t before = simulation::fission_bank.size();
    if (tally)
      tally->reset();
    p.event_collide();
    if (tally)
      kappa[k] = tally->results_(0, 0, static_cast<int>(TallyResult::VALUE));
    double* st = state8 + 8 * k;
    st[0] = p.E(), st[1] = p.u().x, st[2] = p.u().y, st[3] = p.u().z, st[4] = p.alive() ? p.wgt() : 0.0;
    st[5] = p.keff_tally_collision(), st[6] = p.keff_tally_absorption(), st[7] = p.event_nuclide();
    for (int s = 0; s < N_STREAMS; ++s)
      seeds_out[4 * k + s] = p.seeds(s);
    n_sites[k] = (int32_t)(simulation::fission_bank.size() - before);
  }
  const int64_t nb = simulation::fission_bank.size();
  for (int64_t j = 0; j < nb; ++j)
  {
    const auto& s = simulation::fission_bank[j];
    double* o = site7 + 7 * j;
    o[0] = s.E, o[1] = s.u.x, o[2] = s.u.y, o[3] = s.u.z, o[4] = s.delayed_group, o[5] = (double)s.parent_id, o[6] = (double)s.progeny_id;
  }
  model::tallies.clear();
  model::tally_map.clear();
  setup_active_tallies();
  return nb;
}

// kappa-fission track-length score of n segments: Material::calculate_xs, then score_tracklength_tally (tally_scoring.cpp:
// 2289-2352 -> :1216-1258) on a filterless track-length tally, per segment
REF_API void ref_tracklength_kappa(int i_mat, int64_t n, const double* E, const double* wgt, const double* distance,
                                   double sqrtkT, double* out)
{
  using namespace openmc;
  model::tallies.clear();
  model::tally_map.clear();
  model::tally_filters.clear();
  model::filter_map.clear();
  Tally* tally = Tally::create(1);
  tally->set_scores(std::vector<std::string>{"kappa-fission"});
  tally->estimator_ = TallyEstimator::TRACKLENGTH;
  tally->set_active(true);
  tally->init_results();
  setup_active_tallies();
  Particle p;
  p.type() = ParticleType::neutron;
  p.n_coord() = 1;
  p.material() = i_mat;
  p.sqrtkT() = sqrtkT;
  for (int64_t k = 0; k < n; ++k)
  {
    p.E() = p.E_last() = E[k];
    p.wgt() = wgt[k];
    model::materials[i_mat]->calculate_xs(p);
    tally->reset();
    score_tracklength_tally(p, distance[k]);
    out[k] = tally->results_(0, 0, static_cast<int>(TallyResult::VALUE));
  }
  model::tallies.clear();
  model::tally_map.clear();
  setup_active_tallies();
}

// n independent scoring events in material i_mat on a filterless tally with NUCLIDE BINS, built the way MaCaw's
// OpenMCTally::initialize builds it (OpenMCTally.C:188-204): all != 0 -> every nuclide and -1 APPENDED to the default
// nuclides_ = {-1} (no clear: n + 2 bins) and all_nuclides_ (scored by score_all_nuclides, tally_scoring.cpp:2142-2175); else nuclides_ = the given list (library indices, -1 =
// total; scored by the bin loop :2381-2407 / :2310-2336).  Material::calculate_xs, then score_collision_tally
// (distance == nullptr, wgt = wgt_last) or score_tracklength_tally.  out = [event][nuclide bin][score] VALUEs.
REF_API void ref_score_nuclide_bins(int i_mat, int all, const int32_t* nuclides, int n_bins, const char* const* score_names,
                                    int n_scores, int64_t n, const double* E, const double* wgt, const double* distance,
                                    double sqrtkT, double* out)
{
  using namespace openmc;
  model::tallies.clear();
  model::tally_map.clear();
  model::tally_filters.clear();
  model::filter_map.clear();
  Tally* tally = Tally::create(1);
  tally->set_scores(std::vector<std::string>(score_names, score_names + n_scores));
  if (all)
  {
    for (size_t i = 0; i < data::nuclides.size(); ++i)
      tally->nuclides_.push_back((int)i);
    tally->nuclides_.push_back(-1);
    tally->all_nuclides_ = true;
  }
  else
  {
    tally->nuclides_.clear();
    for (int j = 0; j < n_bins; ++j)
      tally->nuclides_.push_back(nuclides[j]);
  }
  tally->estimator_ = distance ? TallyEstimator::TRACKLENGTH : TallyEstimator::COLLISION;
  tally->set_active(true);
  tally->init_results();
  setup_active_tallies();
  const int nb = (int)tally->nuclides_.size();
  Particle p;
  p.type() = ParticleType::neutron;
  p.n_coord() = 1;
  p.material() = i_mat;
  p.sqrtkT() = sqrtkT;
  for (int64_t k = 0; k < n; ++k)
  {
    p.E() = p.E_last() = E[k];
    p.wgt() = p.wgt_last() = wgt[k];
    model::materials[i_mat]->calculate_xs(p);
    tally->reset();
    if (distance)
      score_tracklength_tally(p, distance[k]);
    else
      score_collision_tally(p);
    for (int j = 0; j < nb * n_scores; ++j)
      out[(size_t)k * nb * n_scores + j] = tally->results_(0, j, static_cast<int>(TallyResult::VALUE));
  }
  model::tallies.clear();
  model::tally_map.clear();
  setup_active_tallies();
}

// Whole histories in an infinite medium of material i_mat through the reference's own history loop:
// initialize_history (simulation.cpp:451) + transport_history_based_single_particle (:662-677: event_calculate_xs,
// event_advance, event_collide, event_revive_from_secondary, event_death) in a one-cell, no-surface CSG geometry built here
// (cell 0 fills universe 0 with the material; every flight ends in a collision).  site8 = (r, u, E, wgt) per history.
// Outputs per history: state8 = (final E, n_collision, n_progeny, k_collision, k_absorption, k_tracklength, alive, n_event),
// seeds_out[4]; fission sites of all histories: site7 as in ref_collide_events.  Returns the number of banked sites.
// tallies4 (or null): VALUEs of two tallies scored by the loop itself (event_advance -> score_tracklength_tally,
// event_collide -> score_collision_tally), each EnergyFilter(ebins) x (flux, total, fission, nu-fission):
// [collision | tracklength][n_edges - 1][4].
REF_API int64_t ref_transport_histories(int i_mat, double keff, int64_t master_seed, int batch, int64_t n, const double* site8,
                                        double sqrtkT, double* state8, uint64_t* seeds_out, int64_t bank_capacity,
                                        double* site7, const double* ebins, int n_edges, double* tallies4)
{
  using namespace openmc;
  model::cells.clear();
  model::cell_map.clear();
  model::universes.clear();
  model::universe_map.clear();
  auto cell = std::make_unique<CSGCell>();
  cell->id_ = 1;
  cell->type_ = Fill::MATERIAL;
  cell->universe_ = 0;
  cell->material_ = {i_mat};
  cell->sqrtkT_ = {sqrtkT};
  cell->simple_ = true; // empty region: contains() is true everywhere, no surface to hit
  model::cells.push_back(std::move(cell));
  model::cell_map[1] = 0;
  model::universes.push_back(std::make_unique<Universe>());
  model::universes[0]->id_ = 0;
  model::universes[0]->cells_ = {0};
  model::universe_map[0] = 0;
  model::root_universe = 0;
  model::n_coord_levels = 1;
  model::tallies.clear();
  model::tally_map.clear();
  model::tally_filters.clear();
  model::filter_map.clear();
  Tally* tally[2] = {nullptr, nullptr};
  if (tallies4)
  {
    auto* ef = Filter::create<EnergyFilter>();
    ef->set_bins({ebins, (size_t)n_edges});
    std::vector<Filter*> filters{ef};
    for (int t = 0; t < 2; ++t)
    {
      tally[t] = Tally::create(t + 1);
      tally[t]->set_filters(filters);
      tally[t]->set_scores(std::vector<std::string>{"flux", "total", "fission", "nu-fission"});
      tally[t]->estimator_ = t == 0 ? TallyEstimator::COLLISION : TallyEstimator::TRACKLENGTH;
      tally[t]->set_active(true);
      tally[t]->init_results();
      tally[t]->reset();
    }
  }
  setup_active_tallies();
  settings::check_overlaps = false;
  settings::n_particles = n;
  settings::gen_per_batch = 1;
  settings::write_all_tracks = false;
  settings::track_identifiers.clear();
  mpi::n_procs = 1;
  mpi::rank = 0;
  simulation::work_index = {0, n};
  simulation::work_per_rank = n;
  simulation::progeny_per_particle.assign(n, 0);
  openmc_set_seed(master_seed);
  simulation::total_gen = 0;
  simulation::current_batch = batch;
  simulation::current_gen = 1;
  simulation::keff = keff;
  simulation::fission_bank.clear();
  simulation::fission_bank.reserve(bank_capacity);
  simulation::source_bank.assign(n, SourceSite{});
  for (int64_t k = 0; k < n; ++k)
  {
    SourceSite& s = simulation::source_bank[k];
    const double* v = site8 + 8 * k;
    s.r = {v[0], v[1], v[2]};
    s.u = {v[3], v[4], v[5]};
    s.E = v[6];
    s.wgt = v[7];
    s.particle = ParticleType::neutron;
  }
  for (int64_t k = 0; k < n; ++k)
  {
    Particle p;
    initialize_history(p, k + 1);
    global_tally_collision = global_tally_absorption = global_tally_tracklength = global_tally_leakage = 0.0;
    transport_history_based_single_particle(p);
    double* st = state8 + 8 * k;
    st[0] = p.E(), st[1] = p.n_collision(), st[2] = (double)p.n_progeny();
    st[3] = global_tally_collision, st[4] = global_tally_absorption, st[5] = global_tally_tracklength;
    st[6] = p.alive(), st[7] = p.n_event();
    for (int q = 0; q < N_STREAMS; ++q)
      seeds_out[4 * k + q] = p.seeds(q);
  }
  global_tally_collision = global_tally_absorption = global_tally_tracklength = global_tally_leakage = 0.0;
  if (tallies4)
  {
    size_t o = 0;
    for (int t = 0; t < 2; ++t)
      for (size_t i = 0; i < tally[t]->results_.shape()[0]; ++i)
        for (size_t j = 0; j < tally[t]->results_.shape()[1]; ++j)
          tallies4[o++] = tally[t]->results_(i, j, static_cast<int>(TallyResult::VALUE));
    model::tallies.clear();
    model::tally_map.clear();
    model::tally_filters.clear();
    model::filter_map.clear();
    setup_active_tallies();
  }
  const int64_t nb = simulation::fission_bank.size();
  for (int64_t j = 0; j < nb; ++j)
  {
    const auto& s = simulation::fission_bank[j];
    double* o = site7 + 7 * j;
    o[0] = s.E, o[1] = s.u.x, o[2] = s.u.y, o[3] = s.u.z, o[4] = s.delayed_group, o[5] = (double)s.parent_id, o[6] = (double)s.progeny_id;
  }
  return nb;
}
}
