// orc_transport.cpp -- CPU ORACLE (test infrastructure, NOT product code): history-based transport,
// eigenvalue bookkeeping and the ctypes-facing C entry points.  See orc_core.h for the rules on who
// may use this and how it is pinned to the reference.
#include <cstdio>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_mesh.h"

namespace orc
{
struct Tally
{
  int id, estimator, granularity;
  std::vector<int> scores;
  std::vector<double> ebins; // edges; empty = one bin
  std::vector<int> nuclides; // nuclide bins (library index, -1 = total); empty = the single bin "total"
  int64_t n_spatial = 1;
  int n_energy = 1;
  std::vector<double> value, sum, sum_sq;
  int n_realizations = 0;
  int n_nuc_bins() const { return nuclides.empty() ? 1 : (int)nuclides.size(); }
  int64_t bins() const { return n_spatial * n_energy * n_nuc_bins(); }
};

struct Problem
{
  Mesh mesh;
  Library lib;
  mcw_settings set{};
  std::vector<int32_t> block_material; // block index -> material index
  int n_blocks = 0;
  std::vector<double> temperature;
  // kT of the first (or only) temperature value: what calls without an element use (stand-alone lookups, event replays)
  double kT0() const { return temperature.empty() ? 0.0 : Library::kT_of(temperature[0]); }
  std::vector<Tally> tallies;
  int rank = 0, nranks = 1;
  std::vector<int64_t> work_index;

  std::vector<mcw_source_site> source_bank, fission_bank, sampled;
  std::vector<mcw_history_record> records;
  std::vector<mcw_trace_point> traces; // cached trace points of the last generation (settings.record_traces == 2)
  int64_t bank_global_start = 0, bank_global_total = -1;

  // eigenvalue state (OMC simulation.cpp:253-276)
  double keff = 1.0, keff_std = 0.0;
  bool renormalise_direction = true; // MaCaw / MOOSE re-normalise every new direction; off only to compare with plain OpenMC
  std::vector<double> k_generation;
  double k_sum[2] = {0, 0};
  int n_realizations = 0;
  double gt_value[4] = {0, 0, 0, 0}; // K_COLLISION, K_ABSORPTION, K_TRACKLENGTH, LEAKAGE
  double gt_sum[4] = {0, 0, 0, 0}, gt_sum_sq[4] = {0, 0, 0, 0};
  double total_weight = 0;
  TraceStats stats;
  int64_t n_collisions = 0, n_xs = 0, n_nuc_lookups = 0;
  int64_t n_bank_dropped = 0; // fission sites of the generation that did not fit into the bank (OMC physics.cpp:203-209)

  // point location grid
  int gx = 1, gy = 1, gz = 1;
  std::vector<int64_t> cell_off, cell_elems;
  std::string err;

  void buildLocator()
  {
    const ElemTopo& t = mesh.T();
    double n = std::max<double>(1.0, (double)mesh.n_elem);
    V3 ext = mesh.bb_max - mesh.bb_min;
    if (mesh.dim == 2)
    {
      int g = std::max(1, (int)std::sqrt(n));
      gx = gy = g;
      gz = 1;
    }
    else
    {
      int g = std::max(1, (int)std::cbrt(n));
      gx = gy = gz = g;
    }
    auto cellRange = [&](double lo, double hi, double mn, double extent, int g, int& a, int& b) {
      if (extent <= 0)
      {
        a = b = 0;
        return;
      }
      a = (int)std::floor((lo - mn) / extent * g - 1e-9);
      b = (int)std::floor((hi - mn) / extent * g + 1e-9);
      a = std::min(std::max(a, 0), g - 1);
      b = std::min(std::max(b, 0), g - 1);
    };
    std::vector<int64_t> cnt((size_t)gx * gy * gz + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
    {
      for (int64_t e = 0; e < mesh.n_elem; ++e)
      {
        V3 lo = mesh.P(e, 0), hi = lo;
        for (int v = 1; v < t.nv; ++v)
        {
          V3 p = mesh.P(e, v);
          lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
          hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
        }
        int ax, bx, ay, by, az, bz;
        cellRange(lo.x, hi.x, mesh.bb_min.x, ext.x, gx, ax, bx);
        cellRange(lo.y, hi.y, mesh.bb_min.y, ext.y, gy, ay, by);
        cellRange(lo.z, hi.z, mesh.bb_min.z, ext.z, gz, az, bz);
        for (int k = az; k <= bz; ++k)
          for (int j = ay; j <= by; ++j)
            for (int i = ax; i <= bx; ++i)
            {
              size_t c = ((size_t)k * gy + j) * gx + i;
              if (pass == 0)
                cnt[c + 1]++;
              else
                cell_elems[cnt[c]++] = e;
            }
      }
      if (pass == 0)
      {
        for (size_t c = 0; c + 1 < cnt.size(); ++c)
          cnt[c + 1] += cnt[c];
        cell_off = cnt;
        cell_elems.assign(cnt.back(), 0);
        cnt.assign(cell_off.begin(), cell_off.end());
      }
    }
  }

  // Point locator + ClaimRays tie-break (RT ClaimRays.C:104-139): smallest containing element id
  // for even ray ids, largest for odd.
  int64_t locate(V3 p, int64_t ray_id) const
  {
    V3 ext = mesh.bb_max - mesh.bb_min;
    auto cell = [&](double x, double mn, double extent, int g) {
      if (extent <= 0)
        return 0;
      int c = (int)std::floor((x - mn) / extent * g);
      return std::min(std::max(c, 0), g - 1);
    };
    int i = cell(p.x, mesh.bb_min.x, ext.x, gx), j = cell(p.y, mesh.bb_min.y, ext.y, gy),
        k = mesh.dim == 3 ? cell(p.z, mesh.bb_min.z, ext.z, gz) : 0;
    size_t c = ((size_t)k * gy + j) * gx + i;
    const bool smallest = (ray_id % 2) == 0;
    int64_t best = -1;
    for (int64_t q = cell_off[c]; q < cell_off[c + 1]; ++q)
    {
      int64_t e = cell_elems[q];
      if (!mesh.containsPoint(e, p))
        continue;
      if (best < 0 || (smallest && e < best) || (!smallest && e > best))
        best = e;
    }
    return best;
  }
};

// ------------------------------------------------------------------ collision physics
struct Hist
{
  V3 r, u;
  double E, wgt;
  int64_t id;
  int n_progeny = 0;
  uint64_t seeds[3];
  double k_col = 0, k_abs = 0, k_trk = 0;
  int n_coll = 0, n_xs = 0;
};

// OMC physics.cpp:955-1004
static V3 sample_cxs_target_velocity(double awr, double E, V3 u, double kT, uint64_t* seed)
{
  double beta_vn = std::sqrt(awr * E / kT);
  double alpha = 1.0 / (1.0 + std::sqrt(PI) * beta_vn / 2.0);
  double beta_vt_sq, mu;
  while (true)
  {
    double r1 = prn(seed), r2 = prn(seed);
    if (prn(seed) < alpha)
      beta_vt_sq = -xlog(r1 * r2);
    else
    {
      double c = xcos(PI / 2.0 * prn(seed));
      beta_vt_sq = -xlog(r1) - xlog(r2) * c * c;
    }
    double beta_vt = std::sqrt(beta_vt_sq);
    mu = uniform_distribution(-1., 1., seed);
    double accept_prob =
        std::sqrt(beta_vn * beta_vn + beta_vt_sq - 2 * beta_vn * beta_vt * mu) / (beta_vn + beta_vt);
    if (prn(seed) < accept_prob)
      break;
  }
  double vt = std::sqrt(beta_vt_sq * kT / awr);
  return vt * rotate_angle(u, mu, nullptr, seed);
}

// lower_bound_index (OMC search.h:14-20)
static int lower_bound_index(const std::vector<double>& v, double x)
{
  if (v.front() == x)
    return 0;
  return (int)(std::lower_bound(v.begin(), v.end(), x) - v.begin()) - 1;
}

// Tabular::sample (OMC distribution.cpp:223-266) of distribution k of a table
static double tabular_sample(const Tabular2D& t, int k, uint64_t* seed)
{
  const int b = t.off[k], n = t.off[k + 1] - b;
  const double* x = &t.x[b];
  const double* pd = &t.p[b];
  const double* cd = &t.c[b];
  double c = prn(seed);
  double c_i = cd[0];
  int i;
  for (i = 0; i < n - 1; ++i)
  {
    if (c <= cd[i + 1])
      break;
    c_i = cd[i + 1];
  }
  double x_i = x[i], p_i = pd[i];
  if (t.interp[k] == 1)
  {
    if (p_i > 0.0)
      return x_i + (c - c_i) / p_i;
    return x_i;
  }
  double x_i1 = x[i + 1], p_i1 = pd[i + 1];
  double m = (p_i1 - p_i) / (x_i1 - x_i);
  if (m == 0.0)
    return x_i + (c - c_i) / p_i;
  return x_i + (std::sqrt(std::max(0.0, p_i * p_i + 2 * m * (c - c_i))) - p_i) / m;
}

// AngleDistribution::sample (OMC distribution_angle.cpp:65-95)
static double angle_sample(const Tabular2D& t, double E, uint64_t* seed)
{
  const int n = (int)t.energy.size();
  int i;
  double r;
  if (E < t.energy[0])
  {
    i = 0;
    r = 0.0;
  }
  else if (E > t.energy[n - 1])
  {
    i = n - 2;
    r = 1.0;
  }
  else
  {
    i = lower_bound_index(t.energy, E);
    r = (E - t.energy[i]) / (t.energy[i + 1] - t.energy[i]);
  }
  if (r > prn(seed))
    ++i;
  double mu = tabular_sample(t, i, seed);
  if (std::fabs(mu) > 1.0)
    mu = std::copysign(1.0, mu);
  return mu;
}

// ContinuousTabular::sample (OMC distribution_energy.cpp:150-275)
static double continuous_tabular_sample(const Tabular2D& t, double E, uint64_t* seed)
{
  const bool histogram_interp = t.hist_in_energy;
  const int n_energy_in = (int)t.energy.size();
  int i;
  double r;
  if (E < t.energy[0])
  {
    i = 0;
    r = 0.0;
  }
  else if (E > t.energy[n_energy_in - 1])
  {
    i = n_energy_in - 2;
    r = 1.0;
  }
  else
  {
    i = lower_bound_index(t.energy, E);
    r = (E - t.energy[i]) / (t.energy[i + 1] - t.energy[i]);
  }
  int l;
  if (histogram_interp)
    l = i;
  else
    l = r > prn(seed) ? i + 1 : i;
  double E_i_1 = t.x[t.off[i] + t.ndisc[i]], E_i_K = t.x[t.off[i + 1] - 1];
  double E_i1_1 = t.x[t.off[i + 1] + t.ndisc[i + 1]], E_i1_K = t.x[t.off[i + 2] - 1];
  double E_1 = E_i_1 + r * (E_i1_1 - E_i_1);
  double E_K = E_i_K + r * (E_i1_K - E_i_K);
  const int b = t.off[l], n_energy_out = t.off[l + 1] - b, n_discrete = t.ndisc[l];
  const double* eo = &t.x[b];
  const double* pd = &t.p[b];
  const double* cd = &t.c[b];
  double r1 = prn(seed);
  double c_k = cd[0];
  int k = 0, end = n_energy_out - 2;
  for (int j = 0; j < n_discrete; ++j)
  {
    k = j;
    c_k = cd[k];
    if (r1 < c_k)
    {
      end = j;
      break;
    }
  }
  double c_k1;
  for (int j = n_discrete; j < end; ++j)
  {
    k = j;
    c_k1 = cd[k + 1];
    if (r1 < c_k1)
      break;
    k = j + 1;
    c_k = c_k1;
  }
  double E_l_k = eo[k], p_l_k = pd[k];
  double E_out = E_l_k;
  if (t.interp[l] == 1)
  {
    if (p_l_k > 0.0 && k >= n_discrete)
      E_out = E_l_k + (r1 - c_k) / p_l_k;
  }
  else
  {
    double E_l_k1 = eo[k + 1], p_l_k1 = pd[k + 1];
    if (E_l_k != E_l_k1)
    {
      double frac = (p_l_k1 - p_l_k) / (E_l_k1 - E_l_k);
      if (frac == 0.0)
        E_out = E_l_k + (r1 - c_k) / p_l_k;
      else
        E_out = E_l_k + (std::sqrt(std::max(0.0, p_l_k * p_l_k + 2.0 * frac * (r1 - c_k))) - p_l_k) / frac;
    }
  }
  if (!histogram_interp && n_energy_out > 1 && k >= n_discrete)
  {
    if (l == i)
      return E_1 + (E_out - E_i_1) * (E_K - E_1) / (E_i_K - E_i_1);
    return E_1 + (E_out - E_i1_1) * (E_K - E_1) / (E_i1_K - E_i1_1);
  }
  return E_out;
}

// ---- thermal scattering S(a,b): OMC secondary_thermal.cpp
// get_energy_index (:14-25)
static void get_energy_index(const double* e, int n, double E, int& i, double& f)
{
  i = 0;
  f = 0.0;
  if (E >= e[0])
  {
    i = Library::lower_bound_index(e, e + n, E);
    f = (E - e[i]) / (e[i + 1] - e[i]);
  }
}

// CoherentElasticAE::sample (:36-58)
static void sab_coherent_elastic(const ThermalTable& t, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  int i;
  double f;
  const int n = (int)t.el_x.size();
  get_energy_index(t.el_x.data(), n, E_in, i, f);
  const double prob = prn(seed) * t.el_y[i + 1];
  int k = 0;
  if (prob >= t.el_y[0])
    k = Library::lower_bound_index(t.el_y.data(), t.el_y.data() + (i + 1), prob);
  mu = 1.0 - 2.0 * t.el_x[k] / E_in;
  E_out = E_in;
}

// IncoherentElasticAEDiscrete::sample (:96-126)
static void sab_incoherent_elastic_discrete(const ThermalTable& t, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  int i;
  double f;
  get_energy_index(t.el_x.data(), (int)t.el_x.size(), E_in, i, f);
  const int n_mu = t.n_el_mu;
  const int k = (int)(prn(seed) * n_mu);
  const double* m0 = &t.el_mu[(size_t)i * n_mu];
  const double* m1 = m0 + n_mu;
  mu = m0[k] + f * (m1[k] - m0[k]);
  const double mu_left = (k == 0) ? -1.0 - (mu + 1.0) : m0[k - 1] + f * (m1[k - 1] - m0[k - 1]);
  const double mu_right = (k == n_mu - 1) ? 1.0 + (1.0 - mu) : m0[k + 1] + f * (m1[k + 1] - m0[k + 1]);
  mu += std::min(mu - mu_left, mu_right - mu) * (prn(seed) - 0.5);
  E_out = E_in;
}

// IncoherentInelasticAEDiscrete::sample (:138-188)
static void sab_inelastic_discrete(const ThermalTable& t, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  int i;
  double f;
  get_energy_index(t.ixs_e.data(), (int)t.ixs_e.size(), E_in, i, f);
  int j;
  const int n = t.n_e_out;
  if (!t.skewed)
    j = (int)(prn(seed) * n);
  else
  {
    const double r = prn(seed) * (n - 3);
    if (r > 1.0)
      j = (int)(r + 1);
    else if (r > 0.6)
      j = n - 2;
    else if (r > 0.5)
      j = n - 1;
    else if (r > 0.1)
      j = 1;
    else
      j = 0;
  }
  const double E_ij = t.eout[(size_t)i * n + j];
  const double E_i1j = t.eout[(size_t)(i + 1) * n + j];
  E_out = (1 - f) * E_ij + f * E_i1j;
  const int m = t.n_mu;
  const int k = (int)(prn(seed) * m);
  const double mu_ijk = t.mu[((size_t)i * n + j) * m + k];
  const double mu_i1jk = t.mu[((size_t)(i + 1) * n + j) * m + k];
  mu = (1 - f) * mu_ijk + f * mu_i1jk;
}

// IncoherentInelasticAE::sample (:236-332)
static void sab_inelastic_continuous(const ThermalTable& t, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  int i;
  double f;
  get_energy_index(t.ien.data(), (int)t.ien.size(), E_in, i, f);
  const int l = f > 0.5 ? i + 1 : i;
  const int b = t.ioff[l];
  const size_t n = (size_t)(t.ioff[l + 1] - b);
  const double* e_out = &t.eout[b];
  const double* pdf = &t.pdf[b];
  const double* cdf = &t.cdf[b];
  const double r1 = prn(seed);
  double c_j = cdf[0];
  double c_j1 = 0.0; // the reference leaves it uninitialised; it is always assigned when n >= 2
  size_t j;
  for (j = 0; j < n - 1; ++j)
  {
    c_j1 = cdf[j + 1];
    if (r1 < c_j1)
      break;
    c_j = c_j1;
  }
  j = std::min(j, n - 2);
  const double E_l_j = e_out[j], p_l_j = pdf[j];
  const double E_l_j1 = e_out[j + 1], p_l_j1 = pdf[j + 1];
  const double frac = (p_l_j1 - p_l_j) / (E_l_j1 - E_l_j);
  if (frac == 0.0)
    E_out = E_l_j + (r1 - c_j) / p_l_j;
  else
    E_out = E_l_j + (std::sqrt(std::max(0.0, p_l_j * p_l_j + 2.0 * frac * (r1 - c_j))) - p_l_j) / frac;
  const double E_l = t.ien[l];
  if (E_out < 0.5 * E_l)
    E_out *= 2.0 * E_in / E_l - 1.0;
  else
    E_out += E_in - E_l;
  const int n_mu = t.n_mu;
  const size_t k = (size_t)(prn(seed) * n_mu);
  const double* mu_j = &t.mu[((size_t)b + j) * n_mu];
  const double* mu_j1 = mu_j + n_mu;
  f = (r1 - c_j) / (c_j1 - c_j);
  mu = mu_j[k] + f * (mu_j1[k] - mu_j[k]);
  const double mu_left = (k == 0) ? -1.0 - (mu + 1.0) : mu_j[k - 1] + f * (mu_j1[k - 1] - mu_j[k - 1]);
  const double mu_right = (k == (size_t)n_mu - 1) ? 1.0 + (1.0 - mu) : mu_j[k + 1] + f * (mu_j1[k + 1] - mu_j[k + 1]);
  mu += std::min(mu - mu_left, mu_right - mu) * (prn(seed) - 0.5);
}

// sab_scatter (OMC physics.cpp:788-802) over ThermalData::sample (thermal.cpp:245-262)
static void sab_scatter(const ThermalTable& t, const Micro& micro, Hist& p)
{
  double E_out, mu;
  if (prn(&p.seeds[0]) < micro.thermal_elastic / micro.thermal)
  {
    if (t.el_type == MCW_SAB_ELASTIC_COHERENT)
      sab_coherent_elastic(t, p.E, E_out, mu, &p.seeds[0]);
    else
      sab_incoherent_elastic_discrete(t, p.E, E_out, mu, &p.seeds[0]);
  }
  else if (t.inel_type == MCW_SAB_INELASTIC_CONTINUOUS)
    sab_inelastic_continuous(t, p.E, E_out, mu, &p.seeds[0]);
  else
    sab_inelastic_discrete(t, p.E, E_out, mu, &p.seeds[0]);
  if (std::fabs(mu) > 1.0)
    mu = std::copysign(1.0, mu);
  p.E = E_out;
  p.u = rotate_angle(p.u, mu, nullptr, &p.seeds[0]);
}

// OMC physics.cpp:721-786 (tabular angular distribution :749-757, else "assume isotropic")
static void elastic_scatter(const Nuclide& nuc, double kT, Hist& p, bool use_ptable = false)
{
  double vel = std::sqrt(p.E);
  double awr = nuc.awr;
  V3 v_n = vel * p.u;
  V3 v_t = {0, 0, 0};
  // OMC physics.cpp:804-835 sample_target_velocity, non-resonant branch; a nuclide in its unresolved range is taken
  // at rest (:735-738)
  if (!use_ptable && !(p.E >= FREE_GAS_THRESHOLD * kT && awr > 1.0))
    v_t = sample_cxs_target_velocity(awr, p.E, p.u, kT, &p.seeds[0]);
  V3 v_cm = (v_n + awr * v_t) / (awr + 1.0);
  v_n = v_n - v_cm;
  vel = norm(v_n);
  double mu_cm;
  if (!nuc.el_angle.empty())
    mu_cm = angle_sample(nuc.el_angle, p.E, &p.seeds[0]);
  else
    mu_cm = uniform_distribution(-1., 1., &p.seeds[0]);
  V3 u_cm = v_n / vel;
  v_n = vel * rotate_angle(u_cm, mu_cm, nullptr, &p.seeds[0]);
  v_n = v_n + v_cm;
  p.E = dot(v_n, v_n);
  vel = std::sqrt(p.E);
  p.u = v_n / vel;
}

// The outgoing-energy part CorrelatedAngleEnergy::sample (OMC secondary_correlated.cpp:156-245) and KalbachMann::sample
// (secondary_kalbach.cpp:117-222) share: like ContinuousTabular but without the histogram-in-incoming-energy case and
// without its E_l_k != E_l_k1 guard.  Returns what the angular parts need.
struct EnergyPick
{
  double E_out, E_unscaled, r1, c_k, c_k1, E_l_k, E_l_k1;
  int l, k;      // distribution, point within it
  bool lin_lin;
};
static EnergyPick correlated_energy_sample(const Tabular2D& t, double E_in, uint64_t* seed)
{
  EnergyPick o;
  const int n_energy_in = (int)t.energy.size();
  int i;
  double r;
  if (E_in < t.energy[0])
  {
    i = 0;
    r = 0.0;
  }
  else if (E_in > t.energy[n_energy_in - 1])
  {
    i = n_energy_in - 2;
    r = 1.0;
  }
  else
  {
    i = lower_bound_index(t.energy, E_in);
    r = (E_in - t.energy[i]) / (t.energy[i + 1] - t.energy[i]);
  }
  const int l = r > prn(seed) ? i + 1 : i;
  const double E_i_1 = t.x[t.off[i] + t.ndisc[i]], E_i_K = t.x[t.off[i + 1] - 1];
  const double E_i1_1 = t.x[t.off[i + 1] + t.ndisc[i + 1]], E_i1_K = t.x[t.off[i + 2] - 1];
  const double E_1 = E_i_1 + r * (E_i1_1 - E_i_1);
  const double E_K = E_i_K + r * (E_i1_K - E_i_K);
  const int b = t.off[l], n_energy_out = t.off[l + 1] - b, n_discrete = t.ndisc[l];
  const double* eo = &t.x[b];
  const double* pd = &t.p[b];
  const double* cd = &t.c[b];
  const double r1 = prn(seed);
  double c_k = cd[0];
  int k = 0, end = n_energy_out - 2;
  for (int j = 0; j < n_discrete; ++j)
  {
    k = j;
    c_k = cd[k];
    if (r1 < c_k)
    {
      end = j;
      break;
    }
  }
  double c_k1 = 0.0; // uninitialised in the reference when the continuous loop does not run; the data here always run it
  for (int j = n_discrete; j < end; ++j)
  {
    k = j;
    c_k1 = cd[k + 1];
    if (r1 < c_k1)
      break;
    k = j + 1;
    c_k = c_k1;
  }
  const double E_l_k = eo[k], p_l_k = pd[k];
  double E_out;
  o.lin_lin = t.interp[l] != 1;
  o.E_l_k1 = 0.0;
  if (!o.lin_lin)
  {
    if (p_l_k > 0.0 && k >= n_discrete)
      E_out = E_l_k + (r1 - c_k) / p_l_k;
    else
      E_out = E_l_k;
  }
  else
  {
    const double E_l_k1 = eo[k + 1], p_l_k1 = pd[k + 1];
    o.E_l_k1 = E_l_k1;
    const double frac = (p_l_k1 - p_l_k) / (E_l_k1 - E_l_k);
    if (frac == 0.0)
      E_out = E_l_k + (r1 - c_k) / p_l_k;
    else
      E_out = E_l_k + (std::sqrt(std::max(0.0, p_l_k * p_l_k + 2.0 * frac * (r1 - c_k))) - p_l_k) / frac;
  }
  const double E_unscaled = E_out; // the Kalbach-Mann parameters interpolate with this value (:206-209)
  if (k >= n_discrete)
  {
    if (l == i)
      E_out = E_1 + (E_out - E_i_1) * (E_K - E_1) / (E_i_K - E_i_1);
    else
      E_out = E_1 + (E_out - E_i1_1) * (E_K - E_1) / (E_i1_K - E_i1_1);
  }
  o.E_out = E_out, o.E_unscaled = E_unscaled, o.r1 = r1, o.c_k = c_k, o.c_k1 = c_k1, o.E_l_k = E_l_k, o.l = l, o.k = k;
  return o;
}

// CorrelatedAngleEnergy::sample, the angular part (:247-261): the Tabular of the closer outgoing-energy point
static void correlated_sample(const RxLaw& nuc, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  const EnergyPick o = correlated_energy_sample(nuc.inel_energy, E_in, seed);
  E_out = o.E_out;
  const int point = nuc.inel_energy.off[o.l] + o.k; // absolute point index = "incoming energy" index of corr_angle
  if (o.r1 - o.c_k < o.c_k1 - o.r1 || !o.lin_lin)
    mu = tabular_sample(nuc.corr_angle, point, seed);
  else
    mu = tabular_sample(nuc.corr_angle, point + 1, seed);
}

// KalbachMann::sample, the angular part (:189-234)
static void kalbach_sample(const RxLaw& nuc, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  const EnergyPick o = correlated_energy_sample(nuc.inel_energy, E_in, seed);
  E_out = o.E_out;
  const int point = nuc.inel_energy.off[o.l] + o.k;
  double km_r, km_a;
  if (!o.lin_lin)
  {
    km_r = nuc.km_r[point];
    km_a = nuc.km_a[point];
  }
  else
  {
    km_r = nuc.km_r[point] + (o.E_unscaled - o.E_l_k) / (o.E_l_k1 - o.E_l_k) * (nuc.km_r[point + 1] - nuc.km_r[point]);
    km_a = nuc.km_a[point] + (o.E_unscaled - o.E_l_k) / (o.E_l_k1 - o.E_l_k) * (nuc.km_a[point + 1] - nuc.km_a[point]);
  }
  if (prn(seed) > km_r)
  {
    const double T = uniform_distribution(-1., 1., seed) * xsinh(km_a);
    mu = xlog(T + std::sqrt(T * T + 1.0)) / km_a;
  }
  else
  {
    const double r1 = prn(seed);
    mu = xlog(r1 * xexp(km_a) + (1.0 - r1) * xexp(-km_a)) / km_a;
  }
}

// NBodyPhaseSpace::sample (OMC secondary_nbody.cpp:25-68)
static void nbody_sample(const RxLaw& nuc, double awr, double E_in, double& E_out, double& mu, uint64_t* seed)
{
  mu = uniform_distribution(-1., 1., seed);
  const double Ap = nuc.nbody_mass_ratio;
  const double E_max = (Ap - 1.0) / Ap * (awr / (awr + 1.0) * E_in + nuc.nbody_q);
  const double x = maxwell_spectrum(1.0, seed);
  double y;
  if (nuc.nbody_n == 3)
    y = maxwell_spectrum(1.0, seed);
  else if (nuc.nbody_n == 4)
  {
    const double r1 = prn(seed), r2 = prn(seed), r3 = prn(seed);
    y = -xlog(r1 * r2 * r3);
  }
  else
  {
    const double r1 = prn(seed), r2 = prn(seed), r3 = prn(seed), r4 = prn(seed), r5 = prn(seed), r6 = prn(seed);
    const double c = xcos(PI / 2.0 * r6);
    y = -xlog(r1 * r2 * r3 * r4) - xlog(r5) * (c * c); // std::pow(cos, 2) is the correctly rounded square
  }
  const double v = x / (x + y);
  E_out = E_max * v;
}

// WattEnergy / MaxwellEnergy / Evaporation::sample with tabulated parameters and the restriction energy u (OMC
// distribution_energy.cpp:342-356, 279-291, 305-322): fission-neutron products and uncorrelated scattering products
static double evaporation_type_energy(int law, const Fn& fa, const Fn& fb, double u, double E, uint64_t* seed)
{
  if (law == MCW_LAW_WATT)
  {
    const double a = fa(E), b = fb(E);
    while (true)
    {
      const double E_out = watt_spectrum(a, b, seed);
      if (E_out <= E - u)
        return E_out;
    }
  }
  const double theta = fa(E);
  if (law == MCW_LAW_MAXWELL)
    while (true)
    {
      const double E_out = maxwell_spectrum(theta, seed);
      if (E_out <= E - u)
        return E_out;
    }
  // Evaporation (:305-322); the two random numbers of the product commute
  const double y = (E - u) / theta;
  const double v = 1.0 - xexp(-y);
  double x;
  while (true)
  {
    const double r1 = prn(seed), r2 = prn(seed);
    x = -xlog((1.0 - v * r1) * (1.0 - v * r2));
    if (x <= y)
      break;
  }
  return x * theta;
}

// the energy law of a fission-neutron product: one of the above, or ContinuousTabular (:150-275)
static double fission_product_energy(const FisProduct& f, double E, uint64_t* seed)
{
  if (f.law == MCW_LAW_CONTINUOUS_TABULAR)
    return continuous_tabular_sample(f.energy, E, seed);
  return evaporation_type_energy(f.law, f.a, f.b, f.u, E, seed);
}

// OMC physics.cpp:1073-1120 over UncorrelatedAngleEnergy::sample (secondary_uncorrelated.cpp:54-68): cosine from the
// reaction's angle table or uniform, energy from LevelInelastic (distribution_energy.cpp:48-51) or ContinuousTabular
// (:150-275), CM -> LAB when scatter_in_cm_, yield 1
static bool g_count_samples = false; // the counters below are shared by all threads: off unless a test reads them
static long long g_law_samples[8]; // how often each MCW_LAW_* was sampled (test coverage: orc_law_samples)
static long long g_rx_samples[8];  // ... and each position of the reaction list (orc_reaction_samples)
static void inelastic_scatter(const RxLaw& nuc, double awr, Hist& p)
{
  double E_in = p.E;
  double mu, E;
  if (g_count_samples && nuc.inel_law >= 0 && nuc.inel_law < 8)
  {
#pragma omp atomic
    g_law_samples[nuc.inel_law]++;
  }
  // ReactionProduct::sample -> the product's distribution: correlated, Kalbach-Mann and N-body give energy and cosine
  // together (reaction_product.cpp:70-79); uncorrelated = angle table (or isotropic) + energy law
  if (nuc.inel_law == MCW_LAW_CORRELATED)
    correlated_sample(nuc, E_in, E, mu, &p.seeds[0]);
  else if (nuc.inel_law == MCW_LAW_KALBACH_MANN)
    kalbach_sample(nuc, E_in, E, mu, &p.seeds[0]);
  else if (nuc.inel_law == MCW_LAW_NBODY)
    nbody_sample(nuc, awr, E_in, E, mu, &p.seeds[0]);
  else
  {
    if (!nuc.inel_angle.empty())
      mu = angle_sample(nuc.inel_angle, E_in, &p.seeds[0]);
    else
      mu = uniform_distribution(-1., 1., &p.seeds[0]);
    if (nuc.inel_law == MCW_LAW_CONTINUOUS_TABULAR)
      E = continuous_tabular_sample(nuc.inel_energy, E_in, &p.seeds[0]);
    else if (nuc.inel_law >= MCW_LAW_MAXWELL)
      E = evaporation_type_energy(nuc.inel_law, nuc.law_a, nuc.law_b, nuc.law_u, E_in, &p.seeds[0]);
    else
      E = nuc.inel_ratio * (E_in - nuc.inel_E);
  }
  if (nuc.inel_in_cm)
  {
    double E_cm = E;
    double A = awr;
    E = E_cm + (E_in + 2.0 * mu * (A + 1.0) * std::sqrt(E_in * E_cm)) / ((A + 1.0) * (A + 1.0));
    mu = mu * std::sqrt(E_cm / E) + 1.0 / (A + 1.0) * std::sqrt(E_in / E);
  }
  if (std::fabs(mu) > 1.0)
    mu = std::copysign(1.0, mu);
  p.E = E;
  p.u = rotate_angle(p.u, mu, nullptr, &p.seeds[0]);
  // the product's yield (:1108-1118): an integral one makes yield - 1 secondary neutrons -- which MaCaw pops from the
  // scratch particle's bank and drops (MC CollisionKernel.C:232-236) -- and leaves the history as it is; any other
  // multiplies the weight
  const double yield = nuc.yield_fn.kind != MCW_FN_NONE ? nuc.yield_fn(E_in) : nuc.yield;
  if (std::floor(yield) != yield)
    p.wgt *= yield;
}

// OMC physics.cpp:1006-1071
static void sample_fission_neutron(const Nuclide& nuc, double E_in, double E_max,
                                   mcw_source_site* site, uint64_t* seed)
{
  double nu_t = nuc.nu_total(E_in);
  double nu_d = nuc.nu_delayed(E_in);
  double beta = nu_d / nu_t;
  if (prn(seed) < beta)
  {
    double xi = prn(seed) * nu_d;
    double prob = 0.0;
    int group;
    for (group = 1; group < nuc.n_precursor; ++group)
    {
      prob += nuc.fis.empty() ? nuc.dyield[group - 1] : nuc.fis[group].yield(E_in);
      if (xi < prob)
        break;
    }
    group = std::min(group, nuc.n_precursor);
    site->delayed_group = group;
  }
  else
    site->delayed_group = 0;
  double mu;
  while (true)
  {
    // OMC reaction_product.cpp:85-105 (single law) -> secondary_uncorrelated.cpp:54-68
    mu = uniform_distribution(-1., 1., seed);
    if (!nuc.fis.empty())
      site->E = fission_product_energy(nuc.fis[site->delayed_group], E_in, seed);
    else if (site->delayed_group == 0)
      site->E = watt_spectrum(nuc.watt_a, nuc.watt_b, seed); // distribution_energy.cpp:342-356
    else
      site->E = maxwell_spectrum(nuc.dtheta[site->delayed_group - 1], seed); // :279-291
    if (site->E < E_max)
      break;
  }
  V3 u = rotate_angle({1., 0., 0.}, mu, nullptr, seed);
  site->u[0] = u.x;
  site->u[1] = u.y;
  site->u[2] = u.z;
}

enum
{
  EV_SCATTER = 0,
  EV_ABSORB = 1
};

struct CollisionOut
{
  int event;
  int nuclide = -1; // library index of the sampled nuclide
  double kappa_score = 0.0; // sum_i q_i sigma_f,i N_i flux, nuclide by nuclide (kappa-fission, OMC tally_scoring.cpp:1246-1255)
};

// OMC particle.cpp:219-288 event_collide + physics.cpp:38-152, :154-247, :442-465, :592-625, :627-719
static CollisionOut collide(Problem& P, Hist& p, int i_mat, const Macro& macro, const Micro* micro,
                            int elem, std::vector<mcw_source_site>& bank, int64_t bank_capacity)
{
  const Material& mat = P.lib.materials[i_mat];
  CollisionOut out;
  out.event = EV_SCATTER;
  p.k_col += p.wgt * macro.nu_fission / macro.total;
  ++p.n_coll;
  {
    // collision-estimator flux of this event (OMC tally_scoring.cpp:2354-2365): wgt_last / Sigma_t
    const double flux = p.wgt / macro.total;
    for (size_t i = 0; i < mat.nuc.size(); ++i)
    {
      const Nuclide& n = P.lib.nuclides[micro[i].entry >= 0 ? micro[i].entry : mat.nuc[i]];
      if (n.fissionable)
        out.kappa_score += n.q * micro[i].fission * mat.dens[i] * flux;
    }
  }
  // sample_nuclide
  double cutoff = prn(&p.seeds[0]) * macro.total;
  double prob = 0.0;
  size_t k = mat.nuc.size() - 1;
  for (size_t i = 0; i < mat.nuc.size(); ++i)
  {
    prob += mat.dens[i] * micro[i].total;
    if (prob >= cutoff)
    {
      k = i;
      break;
    }
  }
  const Micro& mx = micro[k];
  // the data set the cross sections came from (micro.index_temp): its reaction tables, its kT for the free-gas target
  const Nuclide& nuc = P.lib.nuclides[mx.entry >= 0 ? mx.entry : mat.nuc[k]];
  out.nuclide = mat.nuc[k];
  const double E_last = p.E;
  if (nuc.fissionable)
  {
    // create_fission_sites (eigenvalue mode, no UFS)
    double nu_t = p.wgt / P.keff * 1.0 * mx.nu_fission / mx.total;
    int nu = (int)nu_t;
    if (prn(&p.seeds[0]) <= (nu_t - nu))
      ++nu;
    for (int i = 0; i < nu; ++i)
    {
      mcw_source_site site;
      memset(&site, 0, sizeof(site));
      site.r[0] = p.r.x;
      site.r[1] = p.r.y;
      site.r[2] = p.r.z;
      site.particle = 0;
      site.wgt = 1. / 1.0;
      site.parent_id = p.id;
      site.progeny_id = p.n_progeny++;
      site.surf_id = 0;
      site.elem = elem;
      sample_fission_neutron(nuc, p.E, P.lib.E_max, &site, &p.seeds[0]);
      bool full = false;
#pragma omp critical(orc_bank)
      {
        if ((int64_t)bank.size() >= bank_capacity)
          full = true; // OMC shared_array.h:71-89 returns -1
        else
          bank.push_back(site);
      }
      if (full)
      {
        // OMC physics.cpp:203-209: the site is not banked and no further ones are sampled.  The id it took is handed
        // back so that progeny_per_particle counts banked sites only and sort_fission_bank's offsets stay inside the bank.
        --p.n_progeny;
#pragma omp atomic
        P.n_bank_dropped += nu - i;
        break;
      }
    }
  }
  // absorption (analog)
  if (mx.absorption > 0.0)
  {
    if (mx.absorption > prn(&p.seeds[0]) * mx.total)
    {
      p.k_abs += p.wgt * mx.nu_fission / mx.absorption;
      out.event = EV_ABSORB;
      return out;
    }
  }
  // scatter
  {
    double cut = prn(&p.seeds[0]) * (mx.total - mx.absorption);
    double f = mx.f;
    int ig = mx.i_grid;
    // micro.elastic: corrected by calculate_sab_xs when a table applies, else evaluated now (physics.cpp:644-647)
    double elastic = (mx.i_sab >= 0 || mx.use_ptable) ? mx.elastic : (1.0 - f) * nuc.elastic[ig] + f * nuc.elastic[ig + 1];
    // the scattering reactions in index_inelastic_scatter_ order (:676-697): each one above its threshold adds its cross
    // section until the cutoff is passed.  One reaction: it is the one sampled whatever its value (the loop ends there).
    // Should rounding leave the cutoff above the sum, the last reaction above its threshold is taken (the reference would
    // run off the list there: "Did not sample any reaction").
    const RxLaw* rx = nullptr;
    if (nuc.n_inel > 0)
    {
      if (nuc.more.empty())
      {
        if (ig >= nuc.inel_thr)
          rx = &nuc;
      }
      else
      {
        double prob = elastic;
        for (size_t j = 0; j <= nuc.more.size(); ++j)
        {
          const RxLaw& r = j == 0 ? (const RxLaw&)nuc : nuc.more[j - 1];
          if (ig < r.inel_thr)
            continue;
          rx = &r;
          const double* v = j == 0 ? &nuc.inelastic[ig] : &r.rx_xs[ig - r.inel_thr];
          prob += (1.0 - f) * v[0] + f * v[1];
          if (!(prob < cut))
            break;
        }
      }
    }
    if (elastic - mx.thermal > cut)
      elastic_scatter(nuc, nuc.kT, p, mx.use_ptable); // non-S(a,b) elastic (:649-661)
    else if (elastic > cut && mx.i_sab >= 0)
      sab_scatter(P.lib.thermal[mx.i_sab], mx, p); // :663-672
    else if (rx)
    {
      const int j = rx == (const RxLaw*)&nuc ? 0 : 1 + (int)(rx - nuc.more.data());
      if (g_count_samples && j < 8)
      {
#pragma omp atomic
        g_rx_samples[j]++;
      }
      inelastic_scatter(*rx, nuc.awr, p);
    }
    else
      elastic_scatter(nuc, nuc.kT, p, mx.use_ptable);
    // isotropic-in-lab override for nuclides of the material's <isotropic> list (OMC physics.cpp:706-718)
    if (!mat.iso.empty() && mat.iso[k])
      p.u = isotropic_direction(&p.seeds[0]);
  }
  // URR stream advance when the energy changed (OMC physics.cpp:140-144)
  if (p.E != E_last)
    p.seeds[STREAM_URR] = future_seed((uint64_t)P.lib.identities(), p.seeds[STREAM_URR]);
  return out;
}

static void score_tally(Tally& t, int64_t spatial, double E, const double* score_by_kind)
{
  // OMC filter_energy.cpp:61-75 + tally_scoring.cpp:1453
  int eb = 0;
  if (!t.ebins.empty())
  {
    if (!(E >= t.ebins.front() && E <= t.ebins.back()))
      return;
    eb = Library::lower_bound_index(t.ebins.data(), t.ebins.data() + t.ebins.size(), E);
    if (eb >= t.n_energy)
      eb = t.n_energy - 1;
  }
  int64_t bin = spatial * t.n_energy + eb;
  for (size_t s = 0; s < t.scores.size(); ++s)
  {
    double v = score_by_kind[t.scores[s]];
#pragma omp atomic
    t.value[bin * t.scores.size() + s] += v;
  }
}

// Tally with nuclide bins (MC OpenMCTally.C:188-204): the loop over nuclide bins of score_collision_tally /
// score_tracklength_tally (OMC tally_scoring.cpp:2381-2407, :2310-2336; "all" = score_all_nuclides :2142-2175) around
// score_general_ce with i_nuclide >= 0 (:600-606 total, :746-750 fission, :805-809 nu-fission, :1238-1244
// kappa-fission): micro * atom_density * flux; a nuclide the material does not hold scores nothing (:2391-2394).
// `micro` holds the material's nuclides in material order (what p.neutron_xs(i) holds after calculate_xs).
static void score_tally_nuclides(const Problem& P, Tally& t, int64_t spatial, double E, const Material& mat, const Micro* micro,
                                 double flux, double macro_absorption, const double* score_total)
{
  int eb = 0;
  if (!t.ebins.empty())
  {
    if (!(E >= t.ebins.front() && E <= t.ebins.back()))
      return;
    eb = Library::lower_bound_index(t.ebins.data(), t.ebins.data() + t.ebins.size(), E);
    if (eb >= t.n_energy)
      eb = t.n_energy - 1;
  }
  const size_t ns = t.scores.size();
  const int64_t base = (spatial * t.n_energy + eb) * t.n_nuc_bins();
  for (int j = 0; j < t.n_nuc_bins(); ++j)
  {
    double* out = &t.value[(size_t)(base + j) * ns];
    const int in = t.nuclides[j];
    if (in == -2)
      continue; // allocated, never scored (MaCaw's `nuclides = all`, see mcw_set_tally_nuclides)
    if (in < 0)
    {
      for (size_t s = 0; s < ns; ++s)
      {
#pragma omp atomic
        out[s] += score_total[t.scores[s]];
      }
      continue;
    }
    size_t q = 0;
    while (q < mat.nuc.size() && mat.nuc[q] != in)
      ++q;
    if (q == mat.nuc.size())
      continue;
    const Micro& m = micro[q];
    const double dens = mat.dens[q];
    const Nuclide& nuc = P.lib.nuclides[m.entry >= 0 ? m.entry : in];
    for (size_t s = 0; s < ns; ++s)
    {
      double v;
      switch (t.scores[s])
      {
        case MCW_SCORE_FLUX:
          v = flux;
          break;
        case MCW_SCORE_TOTAL:
          v = m.total * dens * flux;
          break;
        case MCW_SCORE_FISSION:
          if (macro_absorption == 0)
            continue;
          v = m.fission * dens * flux;
          break;
        case MCW_SCORE_NU_FISSION:
          if (macro_absorption == 0)
            continue;
          v = m.nu_fission * dens * flux;
          break;
        default:
          if (macro_absorption == 0.)
            continue;
          v = nuc.fissionable ? nuc.q * m.fission * dens * flux : 0.0;
          break;
      }
#pragma omp atomic
      out[s] += v;
    }
  }
}

// kappa-fission of one segment: q * sigma_f * N * flux nuclide by nuclide (OMC tally_scoring.cpp:1246-1255)
static double tracklength_kappa(const Problem& P, const Material& mat, const Micro* micro, double flux)
{
  double kappa = 0;
  for (size_t i = 0; i < mat.nuc.size(); ++i)
    if (P.lib.nuclides[mat.nuc[i]].fissionable) // q and the flag are the nuclide's, whatever data set was picked
      kappa += P.lib.nuclides[mat.nuc[i]].q * micro[i].fission * mat.dens[i] * flux;
  return kappa;
}

// collision-estimator scores of one collision (OMC tally_scoring.cpp:2354-2365 flux, :557, :600, :752, :799, :1240-1255)
static void collision_scores(double wgt_last, double total, double absorption, double fission, double nu_fission,
                             double kappa_score, double* sc)
{
  const double flux = wgt_last / total;
  sc[MCW_SCORE_FLUX] = flux;
  sc[MCW_SCORE_TOTAL] = total * flux;
  sc[MCW_SCORE_FISSION] = absorption == 0 ? 0.0 : fission * flux;
  sc[MCW_SCORE_NU_FISSION] = absorption == 0 ? 0.0 : nu_fission * flux;
  sc[MCW_SCORE_KAPPA_FISSION] = absorption == 0. ? 0.0 : kappa_score;
}

// track-length-estimator scores of one segment (OMC tally_scoring.cpp:2289-2300 flux = wgt * distance; same score cases);
// returns the flux, kappa-fission is filled in by the caller
static double tracklength_scores(double wgt, double distance, double total, double absorption, double fission,
                                 double nu_fission, double* sc)
{
  const double flux = wgt * distance;
  sc[MCW_SCORE_FLUX] = flux;
  sc[MCW_SCORE_TOTAL] = total * flux;
  sc[MCW_SCORE_FISSION] = absorption == 0 ? 0.0 : fission * flux;
  sc[MCW_SCORE_NU_FISSION] = absorption == 0 ? 0.0 : nu_fission * flux;
  sc[MCW_SCORE_KAPPA_FISSION] = 0.0;
  return flux;
}

static int64_t spatialBin(const Problem& P, const Tally& t, int64_t elem)
{
  return t.granularity == MCW_FILTER_CELL ? elem
                                          : (t.granularity == MCW_FILTER_UNIVERSE ? P.mesh.block[elem] : 0);
}

// id and stream seeds of the i_local-th history of this rank in `batch` (OMC simulation.cpp:466-481 initialize_history)
static void history_identity(const Problem& P, int64_t i_local, int batch, Hist& p)
{
  const mcw_settings& S = P.set;
  p.id = P.work_index[P.rank] + i_local + 1; // OMC simulation.cpp:469
  // overall_generation() = batch (gen_per_batch 1, current_gen 1: MC OpenMCStudy.C:309-310)
  int64_t particle_seed = (int64_t)(0 + batch - 1) * S.n_particles + p.id; // OMC simulation.cpp:477-480
  for (int s = 0; s < 3; ++s)
  {
    p.seeds[s] = future_seed((uint64_t)particle_seed * PRN_STRIDE, (uint64_t)(S.master_seed + s));
    if (S.seed_through_double)
      p.seeds[s] = seed_via_double(p.seeds[s]);
  }
}

// One history = MC OpenMCStudy.C:317-363 (defineRays) + RT TraceRay::trace + MC
// CollisionKernel::onSegment (CollisionKernel.C:98-275).
static void run_history(Problem& P, int64_t i_local, int batch, bool tallies_active,
                        std::vector<mcw_source_site>& bank, int64_t bank_capacity, TraceStats& st,
                        int64_t& n_coll, int64_t& n_xs, int64_t& n_nuc, mcw_history_record& rec)
{
  const mcw_source_site& src = P.source_bank[i_local];
  const mcw_settings& S = P.set;
  Hist p;
  p.r = {src.r[0], src.r[1], src.r[2]};
  V3 dir = {src.u[0], src.u[1], src.u[2]};
  if (S.is_2d)
    dir.z = 0;
  if (P.renormalise_direction)
    dir = unit(dir); // Ray::setStartingDirection (RT Ray.C:291)
  p.u = dir;
  p.E = src.E;
  p.wgt = src.wgt;
  history_identity(P, i_local, batch, p);
  int64_t ray_id = (int64_t)(batch - 1) * S.n_particles + (p.id - 1);
  int64_t elem = src.elem;
  if (elem < 0)
    elem = P.locate(p.r, ray_id);
  memset(&rec, 0, sizeof(rec));
  if (elem < 0)
  {
    rec.fate = 2;
    ++st.failed;
    return;
  }
  int inc_side = startingIncomingSide(P.mesh, elem, p.r, dir);
  std::vector<Micro> micro(32);
  uint64_t hash = 0xcbf29ce484222325ull;
  int64_t last_elem = elem;
  TraceStats my;
  // RT TraceData.h:15-29: the point and the ray's aux data (MC OpenMCStudy.C:190-197 order) cached per completed segment
  auto cache_point = [&](int kind, V3 pt) {
    if (S.record_traces != 2)
      return;
    mcw_trace_point r;
    r.ray_id = ray_id;
    r.seq = (int32_t)my.segments;
    r.kind = kind;
    r.point[0] = pt.x, r.point[1] = pt.y, r.point[2] = pt.z;
    r.aux[0] = p.E, r.aux[1] = p.wgt, r.aux[2] = (double)p.n_progeny, r.aux[3] = (double)p.id;
    r.aux[4] = (double)p.seeds[0], r.aux[5] = (double)p.seeds[1], r.aux[6] = (double)p.seeds[2], r.aux[7] = 0.0;
#pragma omp critical(orc_traces)
    P.traces.push_back(r);
  };
  cache_point(0, p.r);

  auto onSegment = [&](const SegmentCtx& c, V3& nstart, V3& ndir) -> int {
    int blk = P.mesh.block[c.elem];
    int i_mat = P.block_material[blk];
    const Material& mat = P.lib.materials[i_mat];
    if (micro.size() < mat.nuc.size())
      micro.resize(mat.nuc.size());
    // event_calculate_xs: macro recomputed every segment
    // p.sqrtkT from the coupled temperature of the element (MC CollisionKernel.C:109): picks each nuclide's data set
    const double kT = (P.temperature.empty() || !P.lib.multi_temperature())
                          ? 0.0
                          : Library::kT_of(P.temperature[P.temperature.size() == 1 ? 0 : (size_t)c.elem]);
    Macro macro = P.lib.macro_xs(i_mat, p.E, micro.data(), &n_nuc, &p.seeds[STREAM_URR], kT);
    ++p.n_xs;
    // distance to collision: 1 RN per segment (MC CollisionKernel.C:150)
    const double d = -xlog(prn(&p.seeds[0])) / macro.total;
    p.k_trk += p.wgt * std::min(d, c.length) * macro.nu_fission; // :153
    if (tallies_active)
      for (auto& t : P.tallies)
        if (t.estimator == MCW_ESTIMATOR_TRACKLENGTH)
        {
          // OMC tally_scoring.cpp:2289-2352: flux = wgt * distance
          double sc[MCW_N_SCORES];
          const double flux = tracklength_scores(p.wgt, std::min(d, c.length), macro.total, macro.absorption, macro.fission,
                                                 macro.nu_fission, sc);
          sc[MCW_SCORE_KAPPA_FISSION] = macro.absorption == 0. ? 0.0 : tracklength_kappa(P, mat, micro.data(), flux);
          if (!t.nuclides.empty())
            score_tally_nuclides(P, t, spatialBin(P, t, c.elem), p.E, mat, micro.data(), flux, macro.absorption, sc);
          else
            score_tally(t, spatialBin(P, t, c.elem), p.E, sc);
        }
    if (S.seed_through_double)
      for (int s = 0; s < 3; ++s)
        p.seeds[s] = seed_via_double(p.seeds[s]); // :157-159
    if (d > c.length)
    {
      cache_point(1, c.end);
      return 0;
    }
    // collision point computed backwards from the exit point (:173-175)
    V3 pos = c.end - (c.length - d) * c.dir;
    if (S.is_2d)
      pos.z = S.z_coord;
    p.r = pos;
    p.u = c.dir;
    const double E_last = p.E, wgt_last = p.wgt;
    CollisionOut co = collide(P, p, i_mat, macro, micro.data(), (int)c.elem, bank, bank_capacity);
    if (tallies_active)
    {
      // OMC tally_scoring.cpp:2354-2420 score_collision_tally, scores :557,:600,:752,:799,:1240-1255
      double sc[MCW_N_SCORES];
      collision_scores(wgt_last, macro.total, macro.absorption, macro.fission, macro.nu_fission, co.kappa_score, sc);
      for (auto& t : P.tallies)
        if (t.estimator == MCW_ESTIMATOR_COLLISION)
        {
          if (!t.nuclides.empty())
            // `micro` still holds the pre-collision cross sections of the material's nuclides (p.neutron_xs at E_last)
            score_tally_nuclides(P, t, spatialBin(P, t, c.elem), E_last, mat, micro.data(), wgt_last / macro.total,
                                 macro.absorption, sc);
          else
            score_tally(t, spatialBin(P, t, c.elem), E_last, sc);
        }
    }
    if (S.seed_through_double)
      for (int s = 0; s < 3; ++s)
        p.seeds[s] = seed_via_double(p.seeds[s]); // :217-219
    if (co.event == EV_ABSORB)
    {
      cache_point(3, c.end); // the killed ray still caches the end of this segment (RT TraceRay.C:1318-1364)
      return 1;
    }
    // changeRayStartDirection (RT RayKernelBase.C:62-97 -> Ray.C:216-226)
    V3 nd = p.u;
    if (S.is_2d)
      nd.z = 0;
    ndir = P.renormalise_direction ? unit(nd) : nd;
    nstart = pos;
    return 2;
  };

  int max_seg = S.max_events > 0 ? S.max_events : 1000000;
  int end = trace(P.mesh, elem, inc_side, p.r, dir, onSegment, my, &hash, nullptr, max_seg, &last_elem);
  rec.seed_tracking = p.seeds[0];
  rec.seed_source = p.seeds[1];
  rec.seed_urr = p.seeds[2];
  rec.E = p.E;
  rec.k_collision = p.k_col;
  rec.k_absorption = p.k_abs;
  rec.k_tracklength = p.k_trk;
  rec.n_segments = (int)my.segments;
  rec.n_collisions = p.n_coll;
  rec.n_progeny = p.n_progeny;
  rec.last_elem = (int)last_elem;
  rec.fate = end == END_KERNEL_KILL ? 0 : (end == END_BC_KILL ? 1 : (my.segments >= max_seg ? 3 : 2));
  rec.n_xs = p.n_xs;
  rec.trace_hash = hash;
  st.segments += my.segments;
  st.vertex_edge_hits += my.vertex_edge_hits;
  st.moved_through_neighbors += my.moved_through_neighbors;
  st.failed += my.failed;
  st.intersection_calls += my.intersection_calls;
  st.reflections += my.reflections;
  n_coll += p.n_coll;
  n_xs += p.n_xs;
}

static void calc_work(int64_t n_particles, int nranks, std::vector<int64_t>& wi)
{
  // OMC simulation.cpp:529-551
  int64_t min_work = n_particles / nranks, rem = n_particles % nranks;
  wi.assign(nranks + 1, 0);
  for (int i = 0; i < nranks; ++i)
    wi[i + 1] = wi[i] + (i < rem ? min_work + 1 : min_work);
}

} // namespace orc

using namespace orc;

extern "C" {

// ------------------------------------------------------------------ unit-level entry points
double orc_prn(uint64_t* seed) { return prn(seed); }
uint64_t orc_future_seed(uint64_t n, uint64_t seed) { return future_seed(n, seed); }
uint64_t orc_init_seed(int64_t id, int offset, int64_t master) { return init_seed(id, offset, master); }
uint64_t orc_seed_via_double(uint64_t s) { return seed_via_double(s); }
double orc_maxwell_spectrum(double T, uint64_t* seed) { return maxwell_spectrum(T, seed); }
double orc_watt_spectrum(double a, double b, uint64_t* seed) { return watt_spectrum(a, b, seed); }
void orc_rotate_angle(double* uvw, double mu, const double* phi, uint64_t* seed)
{
  V3 u = rotate_angle({uvw[0], uvw[1], uvw[2]}, mu, phi, seed);
  uvw[0] = u.x;
  uvw[1] = u.y;
  uvw[2] = u.z;
}
void orc_sample_cxs_target_velocity(double awr, double E, const double* u, double kT, uint64_t* seed, double* v)
{
  V3 r = sample_cxs_target_velocity(awr, E, V3{u[0], u[1], u[2]}, kT, seed);
  v[0] = r.x, v[1] = r.y, v[2] = r.z;
}
void orc_isotropic_direction(double* uvw, uint64_t* seed)
{
  V3 u = isotropic_direction(seed);
  uvw[0] = u.x;
  uvw[1] = u.y;
  uvw[2] = u.z;
}
// samples of each inelastic law since the last call (so a test can assert that its problem really reaches every law)
void orc_law_samples(long long* out8)
{
  g_count_samples = true; // counting starts with the first read (tests read once to reset)
  for (int i = 0; i < 8; ++i)
  {
    out8[i] = g_law_samples[i];
    g_law_samples[i] = 0;
  }
}
void orc_reaction_samples(long long* out8)
{
  g_count_samples = true;
  for (int i = 0; i < 8; ++i)
  {
    out8[i] = g_rx_samples[i];
    g_rx_samples[i] = 0;
  }
}
double orc_log(double x) { return xlog(x); }
double orc_sin(double x) { return xsin(x); }
double orc_exp(double x) { return xexp(x); }
double orc_sinh(double x) { return xsinh(x); }
double orc_cos(double x) { return xcos(x); }
int orc_is_within_segment(const double* s1, const double* s2, const double* p, double tol)
{
  return isWithinSegment({s1[0], s1[1], s1[2]}, {s2[0], s2[1], s2[2]}, {p[0], p[1], p[2]}, tol);
}
int orc_uses_libm(void)
{
#ifdef MCW_ORACLE_LIBM
  return 1;
#else
  return 0;
#endif
}

// ------------------------------------------------------------------ problem set-up
void* orc_create(int rank, int nranks)
{
  Problem* P = new Problem();
  P->rank = rank;
  P->nranks = nranks;
  return P;
}
void orc_destroy(void* h) { delete (Problem*)h; }

int orc_set_mesh(void* h, int elem_type, int64_t n_nodes, const double* xyz, int64_t n_elem,
                 const int64_t* conn, const int32_t* block_index, int n_blocks,
                 const int64_t* neighbors, const int32_t* side_bnd_id, int n_bnd,
                 const int32_t* bnd_kind)
{
  Problem& P = *(Problem*)h;
  Mesh& m = P.mesh;
  m.type = elem_type;
  const ElemTopo& t = m.T();
  m.n_nodes = n_nodes;
  m.n_elem = n_elem;
  m.xyz.resize(n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i)
    m.xyz[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  m.conn.assign(conn, conn + n_elem * t.nv);
  m.block.assign(block_index, block_index + n_elem);
  m.nbr.assign(neighbors, neighbors + n_elem * t.ns);
  m.bnd.assign(side_bnd_id, side_bnd_id + n_elem * t.ns);
  m.bnd_kind.assign(bnd_kind, bnd_kind + n_bnd);
  P.n_blocks = n_blocks;
  m.finalize(n_blocks);
  P.buildLocator();
  return 0;
}

// the law part of a reaction, from either descriptor
static void take_law(RxLaw& N, int law, int in_cm, const mcw_tabular2d& angle, const mcw_tabular2d& energy,
                     const int32_t* corr_offset, const int32_t* corr_interp, const double* corr_mu, const double* corr_pdf,
                     const double* corr_cdf, const double* km_r, const double* km_a, int nbody_n, double nbody_mass_ratio,
                     double nbody_q, double yield, const mcw_function1d& yield_fn, double u, const mcw_function1d& a,
                     const mcw_function1d& b)
{
  N.yield_fn.take(yield_fn);
  N.law_u = u;
  N.law_a.take(a);
  N.law_b.take(b);
  N.inel_angle.take(angle, false);
  N.inel_energy.take(energy, true);
  N.inel_law = law;
  N.inel_in_cm = law == MCW_LAW_LEVEL ? true : in_cm != 0;
  if (law == MCW_LAW_CORRELATED || law == MCW_LAW_KALBACH_MANN)
  {
    const int np = N.inel_energy.off.back();
    if (law == MCW_LAW_CORRELATED)
    {
      // one "incoming energy" per point: reuse Tabular2D (energy unused)
      Tabular2D& c = N.corr_angle;
      c = Tabular2D();
      c.energy.assign(np, 0.0);
      c.off.assign(corr_offset, corr_offset + np + 1);
      c.interp.assign(corr_interp, corr_interp + np);
      c.ndisc.assign(np, 0);
      c.x.assign(corr_mu, corr_mu + c.off[np]);
      c.p.assign(corr_pdf, corr_pdf + c.off[np]);
      c.c.assign(corr_cdf, corr_cdf + c.off[np]);
    }
    else
    {
      N.km_r.assign(km_r, km_r + np);
      N.km_a.assign(km_a, km_a + np);
    }
  }
  if (law >= MCW_LAW_CORRELATED && law <= MCW_LAW_NBODY)
    N.inel_angle = Tabular2D(); // the product distribution gives the cosine itself
  N.nbody_n = nbody_n;
  N.nbody_mass_ratio = nbody_mass_ratio;
  N.nbody_q = nbody_q;
  N.yield = yield == 0.0 ? 1.0 : yield;
}

// one data set of the descriptor array as the oracle's Nuclide
static Nuclide nuclide_from_desc(const mcw_nuclide_desc& d)
{
  Nuclide N;
  N.G = d.n_grid;
  N.fissionable = d.fissionable != 0;
  N.awr = d.awr;
  N.kT = d.kT;
  N.energy.assign(d.energy, d.energy + N.G);
  N.xs.assign(d.xs, d.xs + 5 * (size_t)N.G);
  N.elastic.assign(d.elastic, d.elastic + N.G);
  N.n_inel = d.n_inelastic;
  N.inel_thr = d.inel_threshold;
  if (N.n_inel)
  {
    // stored over the full grid with zeros below threshold so indexing matches i_grid
    N.inelastic.assign(N.G, 0.0);
    for (int k = N.inel_thr; k < N.G; ++k)
      N.inelastic[k] = d.inelastic[k - N.inel_thr];
  }
  N.inel_E = d.inel_threshold_energy;
  N.inel_ratio = d.inel_mass_ratio;
  N.nu_poly[0] = d.nu_poly[0];
  N.nu_poly[1] = d.nu_poly[1];
  N.n_precursor = d.n_precursor;
  for (int g = 0; g < MCW_MAX_DELAYED_GROUPS; ++g)
  {
    N.dyield[g] = d.delayed_yield[g];
    N.dtheta[g] = d.delayed_theta[g];
  }
  for (int k = 0; k < d.n_fission_products; ++k)
  {
    const mcw_fission_product& f = d.fission_products[k];
    FisProduct F;
    F.yield.take(f.yield);
    F.a.take(f.a);
    F.b.take(f.b);
    F.law = f.law;
    F.u = f.u;
    F.energy.take(f.energy, true);
    N.fis.push_back(std::move(F));
  }
  if (d.n_fission_products)
    N.total_nu_fn.take(d.total_nu);
  N.watt_a = d.watt_a;
  N.watt_b = d.watt_b;
  N.q = d.q_fission;
  N.el_angle.take(d.elastic_angle, false);
  take_law(N, d.inelastic_law, d.inelastic_in_cm, d.inelastic_angle, d.inelastic_energy, d.corr_offset,
           d.corr_interp, d.corr_mu, d.corr_pdf, d.corr_cdf, d.km_r, d.km_a, d.nbody_n,
           d.nbody_mass_ratio, d.nbody_q, d.inelastic_yield, d.inelastic_yield_fn, d.inelastic_u, d.inelastic_a, d.inelastic_b);
  for (int j = 0; j < d.n_more_reactions && N.n_inel; ++j)
  {
    const mcw_reaction_desc& r = d.more_reactions[j];
    RxLaw R;
    R.mt = r.mt;
    R.inel_thr = r.threshold;
    R.inel_E = r.threshold_energy;
    R.inel_ratio = r.mass_ratio;
    R.rx_xs.assign(r.xs, r.xs + (N.G - r.threshold));
    take_law(R, r.law, r.in_cm, r.angle, r.energy, r.corr_offset, r.corr_interp, r.corr_mu, r.corr_pdf, r.corr_cdf, r.km_r,
             r.km_a, r.nbody_n, r.nbody_mass_ratio, r.nbody_q, r.yield, r.yield_fn, r.u, r.a, r.b);
    N.more.push_back(std::move(R));
  }
  return N;
}

int orc_set_library(void* h, const mcw_nuclide_desc* d, int n, int n_log_bins)
{
  Problem& P = *(Problem*)h;
  P.lib.nuclides.clear();
  for (int i = 0; i < n; ++i)
  {
    Nuclide N = nuclide_from_desc(d[i]);
    N.identity = i;
    N.sets.assign(1, i);
    P.lib.nuclides.push_back(std::move(N));
  }
  P.lib.n_identities = n;
  P.lib.n_log_bins = n_log_bins;
  return 0;
}

// A further data set of nuclide `nuclide` at the temperature d->kT (OMC Nuclide::kTs_[t], grid_[t], xs_[t], every
// reaction's xs_[t]; nuclide.cpp:170-200).  It becomes library entry *entry (>= the number of nuclides), which
// orc_set_urr_tables can address for that temperature's probability tables.
int orc_add_temperature_set(void* h, int nuclide, const mcw_nuclide_desc* d, int* entry)
{
  Problem& P = *(Problem*)h;
  if (nuclide < 0 || nuclide >= P.lib.identities())
    return -1;
  Nuclide N = nuclide_from_desc(*d);
  N.identity = nuclide;
  const int e = (int)P.lib.nuclides.size();
  P.lib.nuclides.push_back(std::move(N));
  std::vector<int>& sets = P.lib.nuclides[nuclide].sets;
  sets.push_back(e);
  std::stable_sort(sets.begin(), sets.end(), [&](int a, int b) { return P.lib.nuclides[a].kT < P.lib.nuclides[b].kT; });
  if (entry)
    *entry = e;
  return 0;
}

// the coupled temperature variable of CollisionKernel: one value, or one per element (MC CollisionKernel.C:33,109)
int orc_set_temperature(void* h, const double* T, int64_t n)
{
  Problem& P = *(Problem*)h;
  P.temperature.assign(T, T + n);
  return 0;
}

int orc_set_materials(void* h, const mcw_material_desc* m, int n, const int32_t* block_material,
                      int n_blocks)
{
  Problem& P = *(Problem*)h;
  P.lib.materials.clear();
  for (int i = 0; i < n; ++i)
  {
    Material M;
    M.nuc.assign(m[i].nuclide, m[i].nuclide + m[i].n_nuclides);
    M.dens.assign(m[i].atom_density, m[i].atom_density + m[i].n_nuclides);
    if (m[i].isotropic_in_lab)
      M.iso.assign(m[i].isotropic_in_lab, m[i].isotropic_in_lab + m[i].n_nuclides);
    P.lib.materials.push_back(std::move(M));
  }
  P.block_material.assign(block_material, block_material + n_blocks);
  P.lib.finalize();
  return 0;
}

// data::thermal_scatt (OMC thermal.cpp:30-32), one temperature each
int orc_set_thermal_tables(void* h, const mcw_thermal_desc* d, int n)
{
  Problem& P = *(Problem*)h;
  P.lib.thermal.clear();
  P.lib.thermal_sets.clear();
  P.lib.thermal_kT.clear();
  for (int k = 0; k < n; ++k)
  {
    const mcw_thermal_desc& s = d[k];
    ThermalTable t;
    t.energy_max = s.energy_max;
    t.inel_type = s.inelastic_type;
    t.el_type = s.elastic_type;
    t.n_mu = s.n_mu;
    t.n_e_out = s.n_e_out;
    t.skewed = s.skewed;
    t.n_el_mu = s.n_elastic_mu;
    t.ixs_e.assign(s.inelastic_xs_energy, s.inelastic_xs_energy + s.n_inelastic_xs);
    t.ixs.assign(s.inelastic_xs, s.inelastic_xs + s.n_inelastic_xs);
    if (s.inelastic_type == MCW_SAB_INELASTIC_CONTINUOUS)
    {
      t.ien.assign(s.inelastic_energy, s.inelastic_energy + s.n_inelastic_energy);
      t.ioff.assign(s.inelastic_offset, s.inelastic_offset + s.n_inelastic_energy + 1);
      const int np = t.ioff.back();
      t.eout.assign(s.inelastic_e_out, s.inelastic_e_out + np);
      t.pdf.assign(s.inelastic_pdf, s.inelastic_pdf + np);
      t.cdf.assign(s.inelastic_cdf, s.inelastic_cdf + np);
      t.mu.assign(s.inelastic_mu, s.inelastic_mu + (size_t)np * s.n_mu);
    }
    else
    {
      const size_t np = (size_t)s.n_inelastic_xs * s.n_e_out;
      t.eout.assign(s.inelastic_e_out, s.inelastic_e_out + np);
      t.mu.assign(s.inelastic_mu, s.inelastic_mu + np * s.n_mu);
    }
    if (s.elastic_type != MCW_SAB_ELASTIC_NONE)
    {
      t.el_x.assign(s.elastic_x, s.elastic_x + s.n_elastic);
      t.el_y.assign(s.elastic_y, s.elastic_y + s.n_elastic);
      if (s.elastic_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
        t.el_mu.assign(s.elastic_mu, s.elastic_mu + (size_t)s.n_elastic * s.n_elastic_mu);
    }
    P.lib.thermal.push_back(std::move(t));
  }
  return 0;
}

// Several temperatures of S(a,b) tables: entry k of orc_set_thermal_tables is a temperature (kT[k]) of the table whose
// first entry is same_table_as[k] (k itself for a table's first or only temperature); materials name first entries.
int orc_set_thermal_temperatures(void* h, const int32_t* same_table_as, const double* kT, int n)
{
  Problem& P = *(Problem*)h;
  if (n != (int)P.lib.thermal.size())
    return -1;
  P.lib.thermal_sets.assign(n, {});
  P.lib.thermal_kT.assign(kT, kT + n);
  for (int k = 0; k < n; ++k)
  {
    if (same_table_as[k] < 0 || same_table_as[k] > k)
      return -1;
    P.lib.thermal_sets[same_table_as[k]].push_back(k);
  }
  for (auto& s : P.lib.thermal_sets)
    std::stable_sort(s.begin(), s.end(), [&](int a, int b) { return P.lib.thermal_kT[a] < P.lib.thermal_kT[b]; });
  return 0;
}

// urr_data_ of the nuclides that have probability tables (OMC nuclide.cpp:225-270, urr.cpp:7-33)
int orc_set_urr_tables(void* h, const mcw_urr_desc* d, int n)
{
  Problem& P = *(Problem*)h;
  for (auto& nuc : P.lib.nuclides)
    nuc.urr_energy.clear(), nuc.urr_table.clear();
  for (int k = 0; k < n; ++k)
  {
    if (d[k].nuclide < 0 || d[k].nuclide >= (int)P.lib.nuclides.size() || d[k].interpolation != 2)
      return -1;
    Nuclide& nuc = P.lib.nuclides[d[k].nuclide];
    nuc.urr_energy.assign(d[k].energy, d[k].energy + d[k].n_energy);
    nuc.urr_table.assign(d[k].table, d[k].table + (size_t)d[k].n_energy * 6 * d[k].n_band);
    nuc.urr_bands = d[k].n_band;
    nuc.urr_inel_flag = d[k].inelastic_flag;
    nuc.urr_multiply_smooth = d[k].multiply_smooth;
  }
  return 0;
}

// Material::thermal_tables_ (OMC material.cpp:140-160,317-360)
int orc_set_material_thermal(void* h, int material, int n, const int32_t* position, const int32_t* table, const double* fraction)
{
  Problem& P = *(Problem*)h;
  if (material < 0 || material >= (int)P.lib.materials.size())
    return -1;
  Material& M = P.lib.materials[material];
  M.sab_table.assign(M.nuc.size(), -1);
  M.sab_frac.assign(M.nuc.size(), 0.0);
  for (int k = 0; k < n; ++k)
  {
    if (position[k] < 0 || position[k] >= (int)M.nuc.size() || table[k] < 0 || table[k] >= (int)P.lib.thermal.size())
      return -1;
    M.sab_table[position[k]] = table[k];
    M.sab_frac[position[k]] = fraction[k];
  }
  if (n == 0)
    M.sab_table.clear(), M.sab_frac.clear();
  return 0;
}

int orc_set_settings(void* h, const mcw_settings* s)
{
  Problem& P = *(Problem*)h;
  P.set = *s;
  calc_work(s->n_particles, P.nranks, P.work_index);
  return 0;
}

int orc_add_tally(void* h, int id, int estimator, int granularity, const int32_t* scores,
                  int n_scores, const double* ebins, int n_edges)
{
  Problem& P = *(Problem*)h;
  Tally t;
  t.id = id;
  t.estimator = estimator;
  t.granularity = granularity;
  t.scores.assign(scores, scores + n_scores);
  if (n_edges >= 2)
  {
    t.ebins.assign(ebins, ebins + n_edges);
    t.n_energy = n_edges - 1;
  }
  t.n_spatial = granularity == MCW_FILTER_CELL
                    ? P.mesh.n_elem
                    : (granularity == MCW_FILTER_UNIVERSE ? P.n_blocks : 1);
  size_t n = (size_t)t.bins() * n_scores;
  t.value.assign(n, 0.0);
  t.sum.assign(n, 0.0);
  t.sum_sq.assign(n, 0.0);
  P.tallies.push_back(std::move(t));
  return 0;
}

// MC OpenMCTally.C:188-204 / OMC tally.cpp:543-558: nuclide bins of tally `id` (-1 = total)
int orc_set_tally_nuclides(void* h, int id, const int32_t* nuclides, int n)
{
  Problem& P = *(Problem*)h;
  for (auto& t : P.tallies)
    if (t.id == id)
    {
      t.nuclides.assign(nuclides, nuclides + n);
      size_t m = (size_t)t.bins() * t.scores.size();
      t.value.assign(m, 0.0);
      t.sum.assign(m, 0.0);
      t.sum_sq.assign(m, 0.0);
      return 0;
    }
  return -1;
}

// n independent scoring events in material i_mat on a filterless tally with nuclide bins: macro_xs, then the collision
// (distance == nullptr: flux = wgt / total) or track-length (flux = wgt * distance) scores of every nuclide bin; out is
// [event][nuclide bin][score].  Held against the reference's score_collision_tally / score_tracklength_tally on its own
// Tally (nuclides_ / all_nuclides_), Material and Nuclide objects (oracle/_ref ref_score_nuclide_bins).
void orc_score_nuclide_bins(void* h, int i_mat, const int32_t* nuclides, int n_bins, const int32_t* scores, int n_scores, int64_t n,
                            const double* E, const double* wgt, const double* distance, double* out)
{
  Problem& P = *(Problem*)h;
  const Material& mat = P.lib.materials[i_mat];
  std::vector<Micro> micro(mat.nuc.size());
  Tally t;
  t.id = 1, t.estimator = distance ? MCW_ESTIMATOR_TRACKLENGTH : MCW_ESTIMATOR_COLLISION, t.granularity = MCW_FILTER_GLOBAL;
  t.scores.assign(scores, scores + n_scores);
  t.nuclides.assign(nuclides, nuclides + n_bins);
  for (int64_t k = 0; k < n; ++k)
  {
    int64_t n_nuc = 0;
    const Macro macro = P.lib.macro_xs(i_mat, E[k], micro.data(), &n_nuc, nullptr, P.kT0());
    double sc[MCW_N_SCORES];
    double flux;
    if (distance)
    {
      flux = tracklength_scores(wgt[k], distance[k], macro.total, macro.absorption, macro.fission, macro.nu_fission, sc);
      sc[MCW_SCORE_KAPPA_FISSION] = macro.absorption == 0. ? 0.0 : tracklength_kappa(P, mat, micro.data(), flux);
    }
    else
    {
      flux = wgt[k] / macro.total;
      collision_scores(wgt[k], macro.total, macro.absorption, macro.fission, macro.nu_fission,
                       tracklength_kappa(P, mat, micro.data(), flux), sc);
    }
    t.value.assign((size_t)n_bins * n_scores, 0.0);
    score_tally_nuclides(P, t, 0, E[k], mat, micro.data(), flux, macro.absorption, sc);
    memcpy(out + (size_t)k * n_bins * n_scores, t.value.data(), sizeof(double) * n_bins * n_scores);
  }
}

// kappa-fission track-length score of n segments (E, wgt, length) in material i_mat, through macro_xs + tracklength_kappa;
// held against the reference's score_tracklength_tally on its own objects (oracle/_ref ref_tracklength_kappa)
void orc_tracklength_kappa(void* h, int i_mat, int64_t n, const double* E, const double* wgt, const double* distance, double* out)
{
  Problem& P = *(Problem*)h;
  const Material& mat = P.lib.materials[i_mat];
  std::vector<Micro> micro(mat.nuc.size());
  for (int64_t k = 0; k < n; ++k)
  {
    int64_t n_nuc = 0;
    const Macro macro = P.lib.macro_xs(i_mat, E[k], micro.data(), &n_nuc, nullptr, P.kT0());
    out[k] = macro.absorption == 0. ? 0.0 : tracklength_kappa(P, mat, micro.data(), wgt[k] * distance[k]);
  }
}

// history_identity for tests (held against the reference's own initialize_history, oracle/_ref ref_initialize_history)
void orc_history_identity(void* h, int64_t i_local, int batch, int64_t* id, uint64_t* seeds3)
{
  Hist p;
  history_identity(*(Problem*)h, i_local, batch, p);
  *id = p.id;
  for (int s = 0; s < 3; ++s)
    seeds3[s] = p.seeds[s];
}

// Replays n independent collision events in material i_mat through the lookup + collision of the transport loop above
// (Library::macro_xs, collide); exists so tests can hold them against the reference's own Material::calculate_xs and
// Particle::event_collide (oracle/_ref ref_collide_events, same argument layout: macro4 = total, absorption, fission,
// nu_fission; state8 = E, u, alive * weight, k_collision, k_absorption, sampled nuclide; site7 = E, u, delayed_group, parent_id,
// progeny_id).  seeds are 4 per event like the reference's streams; the oracle keeps the first 3.
int64_t orc_collide_events(void* h, int i_mat, double keff, int64_t n, const double* E, const double* u3, const uint64_t* seeds4,
                           const int64_t* id, double* macro4, double* state8, uint64_t* seeds_out, int32_t* n_sites,
                           int64_t bank_capacity, double* site7, double* kappa)
{
  Problem& P = *(Problem*)h;
  const double keff_saved = P.keff;
  P.keff = keff;
  std::vector<mcw_source_site> bank;
  std::vector<Micro> micro(P.lib.materials[i_mat].nuc.size());
  for (int64_t k = 0; k < n; ++k)
  {
    Hist p;
    p.r = {0.25 * k, -0.5, 1.0};
    p.u = {u3[3 * k], u3[3 * k + 1], u3[3 * k + 2]};
    p.E = E[k], p.wgt = 1.0, p.id = id[k];
    for (int s = 0; s < 3; ++s)
      p.seeds[s] = seeds4[4 * k + s];
    int64_t n_nuc = 0;
    const Macro macro = P.lib.macro_xs(i_mat, p.E, micro.data(), &n_nuc, &p.seeds[STREAM_URR], P.kT0());
    macro4[4 * k] = macro.total, macro4[4 * k + 1] = macro.absorption;
    macro4[4 * k + 2] = macro.fission, macro4[4 * k + 3] = macro.nu_fission;
    const size_t before = bank.size();
    const CollisionOut co = collide(P, p, i_mat, macro, micro.data(), 0, bank, bank_capacity);
    double* st = state8 + 8 * k;
    st[0] = p.E, st[1] = p.u.x, st[2] = p.u.y, st[3] = p.u.z, st[4] = co.event == EV_ABSORB ? 0.0 : p.wgt;
    st[5] = p.k_col, st[6] = p.k_abs, st[7] = co.nuclide;
    if (kappa)
    {
      double sc[MCW_N_SCORES];
      collision_scores(1.0, macro.total, macro.absorption, macro.fission, macro.nu_fission, co.kappa_score, sc);
      kappa[k] = sc[MCW_SCORE_KAPPA_FISSION];
    }
    for (int s = 0; s < 3; ++s)
      seeds_out[4 * k + s] = p.seeds[s];
    seeds_out[4 * k + 3] = seeds4[4 * k + 3];
    n_sites[k] = (int32_t)(bank.size() - before);
  }
  for (size_t j = 0; j < bank.size(); ++j)
  {
    double* o = site7 + 7 * j;
    o[0] = bank[j].E, o[1] = bank[j].u[0], o[2] = bank[j].u[1], o[3] = bank[j].u[2], o[4] = bank[j].delayed_group;
    o[5] = (double)bank[j].parent_id, o[6] = (double)bank[j].progeny_id;
  }
  P.keff = keff_saved;
  return (int64_t)bank.size();
}

// Replays collision events through the scoring + binning used by the transport loop above (collision_scores +
// score_tally); exists so tests can hold it against the reference's own score_collision_tally /
// score_tracklength_tally (oracle/_ref ref_score_collision_events, ref_score_tracklength_events).  macro4 = (total, absorption, fission, nu_fission) per event; value_out is
// [n_spatial][n_energy][n_scores].
// distance == nullptr: collision estimator; else track-length estimator with (E, wgt) the particle's current values.
int64_t orc_score_events(int64_t n_spatial, const double* ebins, int n_edges, const int32_t* scores, int n_scores,
                         int64_t n_events, const int32_t* spatial, const double* E_last, const double* wgt_last,
                         const double* distance, const double* macro4, double* value_out)
{
  Tally t;
  t.id = 1, t.estimator = distance ? MCW_ESTIMATOR_TRACKLENGTH : MCW_ESTIMATOR_COLLISION, t.granularity = MCW_FILTER_CELL;
  t.scores.assign(scores, scores + n_scores);
  if (n_edges >= 2)
  {
    t.ebins.assign(ebins, ebins + n_edges);
    t.n_energy = n_edges - 1;
  }
  t.n_spatial = n_spatial;
  t.value.assign((size_t)t.bins() * n_scores, 0.0);
  for (int64_t k = 0; k < n_events; ++k)
  {
    double sc[MCW_N_SCORES];
    if (distance)
      tracklength_scores(wgt_last[k], distance[k], macro4[4 * k], macro4[4 * k + 1], macro4[4 * k + 2], macro4[4 * k + 3], sc);
    else
      collision_scores(wgt_last[k], macro4[4 * k], macro4[4 * k + 1], macro4[4 * k + 2], macro4[4 * k + 3], 0.0, sc);
    score_tally(t, spatial[k], E_last[k], sc);
  }
  std::copy(t.value.begin(), t.value.end(), value_out);
  return (int64_t)t.value.size();
}

void orc_get_grid(void* h, double* e_min, double* e_max, double* log_spacing)
{
  Problem& P = *(Problem*)h;
  *e_min = P.lib.E_min;
  *e_max = P.lib.E_max;
  *log_spacing = P.lib.log_spacing;
}
int orc_get_grid_index(void* h, int i_nuc, int32_t* out)
{
  Problem& P = *(Problem*)h;
  auto& g = P.lib.nuclides[i_nuc].grid_index;
  for (size_t i = 0; i < g.size(); ++i)
    out[i] = g[i];
  return (int)g.size();
}

// macroscopic XS for n (E, material) pairs: out = n x 4
int orc_xs_lookup(void* h, int64_t n, const double* E, const int32_t* mat, double* out)
{
  Problem& P = *(Problem*)h;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
  {
    Macro M = P.lib.macro_xs(mat[i], E[i], nullptr, nullptr, nullptr, P.kT0());
    out[4 * i + 0] = M.total;
    out[4 * i + 1] = M.absorption;
    out[4 * i + 2] = M.fission;
    out[4 * i + 3] = M.nu_fission;
  }
  return 0;
}

// geometry-only walk (the RT lots.i workload): no kernels, stop at the first kill boundary
int orc_trace_rays(void* h, int64_t n, const double* start, const double* dirs,
                   const int32_t* start_elem, const int32_t* start_side, const double* max_distance,
                   double* out_distance, int32_t* out_segments, uint64_t* out_hash, int64_t* stats6)
{
  Problem& P = *(Problem*)h;
  TraceStats total;
  for (int64_t i = 0; i < n; ++i)
  {
    V3 p = {start[3 * i], start[3 * i + 1], start[3 * i + 2]};
    V3 d = unit(V3{dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]});
    int64_t e = start_elem[i];
    if (e < 0)
      e = P.locate(p, i);
    int side = start_side ? start_side[i] : INVALID;
    if (side == -2)
      side = startingIncomingSide(P.mesh, e, p, d);
    TraceStats st;
    uint64_t hash = 0xcbf29ce484222325ull;
    double dist = 0;
    auto noop = [](const SegmentCtx&, V3&, V3&) { return 0; };
    trace(P.mesh, e, side, p, d, noop, st, &hash, &dist, 10000000, nullptr,
          max_distance ? max_distance[i] : INFINITY);
    out_distance[i] = dist;
    out_segments[i] = (int32_t)st.segments;
    if (out_hash)
      out_hash[i] = hash;
    total.segments += st.segments;
    total.vertex_edge_hits += st.vertex_edge_hits;
    total.moved_through_neighbors += st.moved_through_neighbors;
    total.failed += st.failed;
    total.intersection_calls += st.intersection_calls;
    total.reflections += st.reflections;
  }
  if (stats6)
  {
    stats6[0] = total.segments;
    stats6[1] = total.vertex_edge_hits;
    stats6[2] = total.moved_through_neighbors;
    stats6[3] = total.failed;
    stats6[4] = total.intersection_calls;
    stats6[5] = total.reflections;
  }
  return 0;
}

int orc_locate_points(void* h, int64_t n, const double* xyz, const int64_t* ray_id, int32_t* elem)
{
  Problem& P = *(Problem*)h;
  for (int64_t i = 0; i < n; ++i)
    elem[i] = (int32_t)P.locate({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, ray_id ? ray_id[i] : i);
  return 0;
}

// ------------------------------------------------------------------ source bank
// OMC source.cpp:325-338 initialize_source + :143-231 IndependentSource::sample (box, isotropic,
// Watt(0.988 MeV, 2.249 /MeV)); geometry rejection = "no mesh element contains the point".
int orc_sample_initial_source(void* h)
{
  Problem& P = *(Problem*)h;
  const mcw_settings& S = P.set;
  int64_t n = P.work_index[P.rank + 1] - P.work_index[P.rank];
  P.source_bank.assign(n, mcw_source_site{});
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i)
  {
    int64_t id = 0 * S.n_particles + P.work_index[P.rank] + i + 1;
    uint64_t seed = init_seed(id, STREAM_SOURCE, S.master_seed);
    mcw_source_site site;
    memset(&site, 0, sizeof(site));
    site.wgt = 1.0;
    int64_t ray_id = P.work_index[P.rank] + i;
    int64_t e = -1;
    for (int tries = 0; tries < 1000000; ++tries)
    {
      double x1 = prn(&seed), x2 = prn(&seed), x3 = prn(&seed);
      V3 r = {S.source_box[0] + x1 * (S.source_box[3] - S.source_box[0]),
              S.source_box[1] + x2 * (S.source_box[4] - S.source_box[1]),
              S.source_box[2] + x3 * (S.source_box[5] - S.source_box[2])};
      e = P.locate(r, ray_id);
      bool found = e >= 0;
      if (found && S.source_only_fissionable)
        found = P.lib.materials[P.block_material[P.mesh.block[e]]].fissionable;
      if (found)
      {
        site.r[0] = r.x;
        site.r[1] = r.y;
        site.r[2] = r.z;
        break;
      }
      e = -1;
    }
    V3 u = isotropic_direction(&seed);
    site.u[0] = u.x;
    site.u[1] = u.y;
    site.u[2] = u.z;
    while (true)
    {
      site.E = watt_spectrum(0.988e6, 2.249e-6, &seed);
      if (site.E < P.lib.E_max && site.E > P.lib.E_min)
        break;
    }
    site.elem = (int32_t)e;
    site.parent_id = 0;
    site.progeny_id = 0;
    P.source_bank[i] = site;
  }
  return 0;
}

int orc_set_source_bank(void* h, const mcw_source_site* s, int64_t n)
{
  ((Problem*)h)->source_bank.assign(s, s + n);
  return 0;
}
int64_t orc_get_source_bank(void* h, mcw_source_site* s, int64_t cap)
{
  Problem& P = *(Problem*)h;
  int64_t n = std::min<int64_t>(cap, (int64_t)P.source_bank.size());
  if (s)
    memcpy(s, P.source_bank.data(), n * sizeof(mcw_source_site));
  return (int64_t)P.source_bank.size();
}
int64_t orc_get_fission_bank(void* h, mcw_source_site* s, int64_t cap)
{
  Problem& P = *(Problem*)h;
  int64_t n = std::min<int64_t>(cap, (int64_t)P.fission_bank.size());
  if (s)
    memcpy(s, P.fission_bank.data(), n * sizeof(mcw_source_site));
  return (int64_t)P.fission_bank.size();
}
int64_t orc_get_sampled(void* h, mcw_source_site* s, int64_t cap)
{
  Problem& P = *(Problem*)h;
  int64_t n = std::min<int64_t>(cap, (int64_t)P.sampled.size());
  if (s)
    memcpy(s, P.sampled.data(), n * sizeof(mcw_source_site));
  return (int64_t)P.sampled.size();
}
int64_t orc_get_traces(void* h, mcw_trace_point* r, int64_t cap)
{
  Problem& P = *(Problem*)h;
  std::sort(P.traces.begin(), P.traces.end(), [](const mcw_trace_point& a, const mcw_trace_point& b) {
    return a.ray_id != b.ray_id ? a.ray_id < b.ray_id : (a.seq != b.seq ? a.seq < b.seq : a.kind < b.kind);
  });
  int64_t n = std::min<int64_t>(cap, (int64_t)P.traces.size());
  if (r)
    memcpy(r, P.traces.data(), n * sizeof(mcw_trace_point));
  return (int64_t)P.traces.size();
}
int64_t orc_get_history_records(void* h, mcw_history_record* r, int64_t cap)
{
  Problem& P = *(Problem*)h;
  int64_t n = std::min<int64_t>(cap, (int64_t)P.records.size());
  if (r)
    memcpy(r, P.records.data(), n * sizeof(mcw_history_record));
  return (int64_t)P.records.size();
}

// ------------------------------------------------------------------ one generation
// Transport half: OMC simulation.cpp:296-332,392-405 (initialize_batch / generation), all
// histories, then bank.cpp:60-104 sort_fission_bank.
int orc_transport_generation(void* h, int batch, int n_threads)
{
  Problem& P = *(Problem*)h;
  const int64_t n = (int64_t)P.source_bank.size();
  P.fission_bank.clear();
  P.traces.clear();
  P.fission_bank.reserve(3 * n);
  P.n_bank_dropped = 0;
  P.records.assign(n, mcw_history_record{});
  P.total_weight = 0.0;
  const bool tallies_active = batch > P.set.n_inactive;
  TraceStats st;
  int64_t n_coll = 0, n_xs = 0, n_nuc = 0;
  (void)n_threads;
#ifdef _OPENMP
  if (n_threads > 0)
    omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    TraceStats my;
    int64_t c = 0, x = 0, nn = 0;
#pragma omp for schedule(dynamic, 20) // RT ParallelStudy.h:507-523
    for (int64_t i = 0; i < n; ++i)
      run_history(P, i, batch, tallies_active, P.fission_bank, 3 * n, my, c, x, nn, P.records[i]);
#pragma omp critical(orc_stats)
    {
      st.segments += my.segments;
      st.vertex_edge_hits += my.vertex_edge_hits;
      st.moved_through_neighbors += my.moved_through_neighbors;
      st.failed += my.failed;
      st.intersection_calls += my.intersection_calls;
      st.reflections += my.reflections;
      n_coll += c;
      n_xs += x;
      n_nuc += nn;
    }
  }
  P.stats = st;
  P.n_collisions = n_coll;
  P.n_xs = n_xs;
  P.n_nuc_lookups = n_nuc;
  // global tally accumulators in history order (OMC particle.cpp:331-345; simulation.cpp:409-425)
  double kc = 0, ka = 0, kt = 0, w = 0;
  for (int64_t i = 0; i < n; ++i)
  {
    kc += P.records[i].k_collision;
    ka += P.records[i].k_absorption;
    kt += P.records[i].k_tracklength;
    w += P.source_bank[i].wgt;
  }
  P.gt_value[0] += kc;
  P.gt_value[1] += ka;
  P.gt_value[2] += kt;
  P.total_weight = w;
  // sort by (parent_id, progeny_id): equivalent to OMC bank.cpp:60-104
  std::sort(P.fission_bank.begin(), P.fission_bank.end(),
            [](const mcw_source_site& a, const mcw_source_site& b) {
              return a.parent_id != b.parent_id ? a.parent_id < b.parent_id
                                                : a.progeny_id < b.progeny_id;
            });
  P.bank_global_start = 0;
  P.bank_global_total = (int64_t)P.fission_bank.size();
  return 0;
}

int orc_set_bank_window(void* h, int64_t start, int64_t total)
{
  Problem& P = *(Problem*)h;
  P.bank_global_start = start;
  P.bank_global_total = total;
  return 0;
}

// Sampling half of OMC eigenvalue.cpp:71-311 synchronize_bank on this rank's slice.  Returns the
// number of sites sampled locally (before the last-rank trim/pad, which the caller applies through
// orc_finish_sampling once it knows the global offsets).
int64_t orc_sample_bank(void* h, int batch)
{
  Problem& P = *(Problem*)h;
  const mcw_settings& S = P.set;
  const int64_t total = P.bank_global_total, start = P.bank_global_start;
  int64_t id = 0 + batch; // total_gen + overall_generation()
  uint64_t seed = init_seed(id, STREAM_TRACKING, S.master_seed);
  seed = future_seed((uint64_t)start, seed);
  int64_t sites_needed = total < S.n_particles ? S.n_particles % total : S.n_particles;
  double p_sample = (double)sites_needed / (double)total;
  P.sampled.clear();
  for (size_t i = 0; i < P.fission_bank.size(); ++i)
  {
    const auto& site = P.fission_bank[i];
    if (total < S.n_particles)
      for (int64_t j = 1; j <= S.n_particles / total; ++j)
        P.sampled.push_back(site);
    if (prn(&seed) < p_sample)
      P.sampled.push_back(site);
  }
  return (int64_t)P.sampled.size();
}

// Last-rank adjustment (OMC eigenvalue.cpp:184-206): `start` = global offset of this rank's sampled
// sites, is_last = this is the last rank.
int64_t orc_finish_sampling(void* h, int64_t start, int is_last)
{
  Problem& P = *(Problem*)h;
  const int64_t n = P.set.n_particles;
  int64_t finish = start + (int64_t)P.sampled.size();
  if (is_last)
  {
    if (finish > n)
      P.sampled.resize(n - start);
    else if (finish < n)
    {
      int64_t needed = n - finish;
      for (int64_t i = 0; i < needed; ++i)
      {
        int64_t ib = (int64_t)P.fission_bank.size() - needed + i;
        if (ib < 0)
          ib = 0; // MC OpenMCStudy.C:584
        P.sampled.push_back(P.fission_bank[ib]);
      }
    }
  }
  return (int64_t)P.sampled.size();
}

// k estimators + tally accumulation: OMC eigenvalue.cpp:49-69,313-351; tallies/tally.cpp:639-670,
// 805-846; simulation.cpp:334-347.  `gt` = this generation's global [k_col,k_abs,k_trk,leak] and
// total weight summed over ranks by the caller (or null = local values).
int orc_finalize_generation(void* h, int batch, const double* reduced5, mcw_generation_result* out)
{
  Problem& P = *(Problem*)h;
  const mcw_settings& S = P.set;
  double g[4] = {P.gt_value[0], P.gt_value[1], P.gt_value[2], P.gt_value[3]};
  double w = P.total_weight;
  if (reduced5)
  {
    for (int i = 0; i < 4; ++i)
      g[i] = reduced5[i];
    w = reduced5[4];
  }
  double k_gen = g[2] / (double)S.n_particles;
  P.k_generation.push_back(k_gen);
  int n = batch > S.n_inactive ? P.n_realizations + 1 : 0; // gen_per_batch*n_realizations + current_gen
  if (n <= 0)
    P.keff = k_gen;
  else
  {
    P.k_sum[0] += k_gen;
    P.k_sum[1] += k_gen * k_gen;
    P.keff = P.k_sum[0] / n;
    if (n > 1)
      P.keff_std = std::sqrt((P.k_sum[1] / n - P.keff * P.keff) / (n - 1));
  }
  // accumulate_tallies
  P.n_realizations += 1;
  for (int i = 0; i < 4; ++i)
  {
    double val = g[i] / w;
    P.gt_sum[i] += val;
    P.gt_sum_sq[i] += val * val;
  }
  if (out)
  {
    memset(out, 0, sizeof(*out));
    out->k_generation = k_gen;
    out->keff = P.keff;
    out->keff_std = P.keff_std;
    out->k_collision = g[0] / w;
    out->k_absorption = g[1] / w;
    out->k_tracklength = g[2] / w;
    out->leakage = g[3] / w;
    out->total_weight = w;
    out->n_histories = (int64_t)P.source_bank.size();
    out->n_segments = P.stats.segments;
    out->n_collisions = P.n_collisions;
    out->n_xs_lookups = P.n_xs;
    out->n_nuclide_lookups = P.n_nuc_lookups;
    out->n_bank_sites = (int64_t)P.fission_bank.size();
    out->n_bank_dropped = P.n_bank_dropped;
    out->n_vertex_edge_hits = P.stats.vertex_edge_hits;
    out->n_moved_through_neighbors = P.stats.moved_through_neighbors;
    out->n_failed_traces = P.stats.failed;
  }
  for (int i = 0; i < 4; ++i)
    P.gt_value[i] = 0.0;
  if (batch > S.n_inactive)
  {
    double norm_ = 1.0 / ((double)S.n_particles * 1.0);
    for (auto& t : P.tallies)
    {
      t.n_realizations += 1;
      for (size_t i = 0; i < t.value.size(); ++i)
      {
        double val = t.value[i] * norm_;
        t.value[i] = 0.0;
        t.sum[i] += val;
        t.sum_sq[i] += val * val;
      }
    }
  }
  else
  {
    for (int i = 0; i < 4; ++i)
      P.gt_sum[i] = P.gt_sum_sq[i] = 0.0;
    P.n_realizations = 0;
  }
  return 0;
}

// single-rank convenience: transport + sample + finalize; the sampled sites become the next source
int orc_run_generation(void* h, int batch, int n_threads, mcw_generation_result* out)
{
  Problem& P = *(Problem*)h;
  orc_transport_generation(h, batch, n_threads);
  if (P.fission_bank.empty())
    return MCW_E_PHYSICS;
  orc_sample_bank(h, batch);
  orc_finish_sampling(h, 0, 1);
  orc_finalize_generation(h, batch, nullptr, out);
  P.source_bank = P.sampled;
  return 0;
}

int orc_get_tally(void* h, int id, double* value, double* sum, double* sum_sq, int64_t n)
{
  Problem& P = *(Problem*)h;
  for (auto& t : P.tallies)
    if (t.id == id)
    {
      int64_t m = std::min<int64_t>(n, (int64_t)t.sum.size());
      if (value)
        memcpy(value, t.value.data(), m * 8);
      if (sum)
        memcpy(sum, t.sum.data(), m * 8);
      if (sum_sq)
        memcpy(sum_sq, t.sum_sq.data(), m * 8);
      return t.n_realizations;
    }
  return -1;
}
int64_t orc_tally_size(void* h, int id)
{
  Problem& P = *(Problem*)h;
  for (auto& t : P.tallies)
    if (t.id == id)
      return (int64_t)t.sum.size();
  return -1;
}
void orc_set_keff(void* h, double k) { ((Problem*)h)->keff = k; }
// test hook: 0 = keep directions as the collision physics returns them (plain OpenMC), 1 = MaCaw's Point::unit() (default)
void orc_set_renormalise_direction(void* h, int on) { ((Problem*)h)->renormalise_direction = on != 0; }
double orc_get_keff(void* h) { return ((Problem*)h)->keff; }
int orc_calculate_work(int64_t n, int nranks, int64_t* wi)
{
  std::vector<int64_t> v;
  calc_work(n, nranks, v);
  for (size_t i = 0; i < v.size(); ++i)
    wi[i] = v[i];
  return 0;
}
int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} // extern "C"
