// mcw_physics.cuh -- per-segment and per-collision physics (device).  Mirrors
// MC CollisionKernel.C:98-275 split at the collision point: the part before `if (d > L)` runs in
// the tracking kernel, event_collide() runs in the collision kernel.
#pragma once
#include "mcw_device.cuh"

// ------------------------------------------------------------------ tallies
// OMC filter_energy.cpp:61-75 energy bin + tally_scoring.cpp:1453 atomic update.
// energy bin of a tally (OMC filter_energy.cpp:61-75); -1 = outside the filter's range
MCW_D int mcw_tally_energy_bin(const McwTallyDev& t, double E)
{
  int eb = 0;
  if (t.n_edges >= 2)
  {
    const double* b = t.ebins;
    if (!(E >= b[0] && E <= b[t.n_edges - 1]))
      return -1;
    if (b[0] == E)
      eb = 0;
    else
    {
      int lo = 0, hi = t.n_edges;
      while (lo < hi)
      {
        int mid = (lo + hi) >> 1;
        if (b[mid] < E)
          lo = mid + 1;
        else
          hi = mid;
      }
      eb = lo - 1;
    }
    if (eb >= t.n_energy)
      eb = t.n_energy - 1;
  }
  return eb;
}

MCW_D void mcw_score_tally(const McwTallyDev& t, int64_t spatial, double E, const double* sc)
{
  const int eb = mcw_tally_energy_bin(t, E);
  long long bin = eb >= 0 ? (long long)(spatial * t.n_energy + eb) : -1ll;
  if (t.granularity == MCW_FILTER_CELL)
  {
    // one bin per element: the lanes of a warp sit in different elements (measured: 26.1 distinct bins per 26.3 active
    // lanes, profiles/ncu_r02m_track_tl.txt), so looking for equal bins first costs more than it saves -- and the
    // L2's RED units run at 4 % of their peak.  One FP64 RED per lane and score.
    if (bin >= 0)
      for (int s = 0; s < t.n_scores; ++s)
        atomicAdd(&t.value[bin * t.n_scores + s], sc[t.scores[s]]);
    return;
  }
  // few bins (one per block, or one): lanes that hit the same bin are pre-reduced with __match_any_sync so that each
  // distinct bin gets one FP64 atomic per warp
  unsigned active = __activemask();
  unsigned peers = __match_any_sync(active, bin);
  int lane = threadIdx.x & 31;
  int leader = __ffs(peers) - 1;
  for (int s = 0; s < t.n_scores; ++s)
  {
    double v = sc[t.scores[s]];
    // segmented reduction over the peer group (peers share `bin`)
    double sum = v;
    // every lane of the group iterates over the same peer set
    for (unsigned m = peers; m; m &= m - 1)
    {
      int src = __ffs(m) - 1;
      double o = __shfl_sync(peers, v, src);
      if (src != lane)
        sum += o;
    }
    if (lane == leader && bin >= 0)
      atomicAdd(&t.value[bin * t.n_scores + s], sum);
  }
}

// A tally with nuclide bins (MC OpenMCTally.C:188-204; OMC tally_scoring.cpp:2381-2407 and score_all_nuclides
// :2142-2175): bin j scores micro * atom_density * flux of its nuclide when the event's material holds it
// (score_general_ce, i_nuclide >= 0: :600-606 total, :746-750 fission, :805-809 nu-fission, :1238-1244 kappa-fission)
// and nothing otherwise (:2391-2394); the bin -1 scores the material totals `sc`.  score_index = j * n_scores + s.
// Rarely asked for, so out of line (the collision kernel's code stays compact) and with plain atomics.
__device__ __noinline__ void mcw_score_tally_nuclides(const McwLibraryDev* Lg, const McwTallyDev* tg, int64_t spatial, double E,
                                                      int i_mat, double flux, double sig_a, const double* sc, uint64_t seed_urr)
{
  const McwLibraryDev& L = *Lg; // global-memory copies: see McwLibraryDev::self_dev
  const McwTallyDev& t = *tg;
  const int eb = mcw_tally_energy_bin(t, E);
  if (eb < 0)
    return;
  const long long base = (long long)(spatial * t.n_energy + eb) * t.n_nuc_bins;
  const int i_log = mcw_log_bin(L, E);
  for (int j = 0; j < t.n_nuc_bins; ++j)
  {
    double* out = t.value + (base + j) * t.n_scores;
    const int in = t.nuclides[j];
    if (in == -2)
      continue; // a bin the reference allocates but never scores (mcw_set_tally_nuclides)
    if (in < 0)
    {
      for (int s = 0; s < t.n_scores; ++s)
        atomicAdd(out + s, sc[t.scores[s]]);
      continue;
    }
    int q = L.mat_off[i_mat];
    const int qe = L.mat_off[i_mat + 1];
    while (q < qe && L.nuc[L.mat_nuc[q]].identity != in) // the material's entry may be any of the nuclide's data sets
      ++q;
    if (q == qe)
      continue; // mat_nuclide_index_ == C_NONE
    const double dens = L.mat_dens[q];
    const Micro m = mcw_micro_xs_entry(L, q, E, i_log, L.n_urr ? &seed_urr : nullptr);
    for (int s = 0; s < t.n_scores; ++s)
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
          if (sig_a == 0)
            continue;
          v = m.fission * dens * flux;
          break;
        case MCW_SCORE_NU_FISSION:
          if (sig_a == 0)
            continue;
          v = m.nu_fission * dens * flux;
          break;
        default: // MCW_SCORE_KAPPA_FISSION
          if (sig_a == 0.)
            continue;
          v = L.nuc[in].fissionable ? L.nuc[in].q * m.fission * dens * flux : 0.0;
          break;
      }
      atomicAdd(out + s, v);
    }
  }
}

MCW_D int64_t mcw_spatial_bin(const McwTallyDev& t, int elem, int block)
{
  return t.granularity == MCW_FILTER_CELL ? (int64_t)elem
                                          : (t.granularity == MCW_FILTER_UNIVERSE ? (int64_t)block : 0);
}

// kappa-fission of a material: q_i * sigma_f,i * N_i * flux summed nuclide by nuclide (OMC tally_scoring.cpp:1246-1255)
MCW_D double mcw_kappa_macro(const McwLibraryDev& L, int i_mat, double E, double flux, uint64_t seed_urr)
{
  int i_log = mcw_log_bin(L, E);
  double kappa = 0.0;
  for (int i = L.mat_off[i_mat]; i < L.mat_off[i_mat + 1]; ++i)
  {
    int in = L.mat_nuc[i];
    if (L.nuc[in].fissionable)
    {
      Micro m = mcw_micro_xs_entry(L, i, E, i_log, (MCW_RARE && L.n_urr) ? &seed_urr : nullptr);
      kappa += L.nuc[in].q * m.fission * L.mat_dens[i] * flux;
    }
  }
  return kappa;
}

// ------------------------------------------------------------------ cached trace points
// RT TraceData.h:15-29 / RayTracingStudy's cached traces with aux data (what RayTracingExodus writes).  Off the hot
// path: only when settings.record_traces == 2.
MCW_D void mcw_trace_emit(const McwParams& P, int i, int seq, int kind, V3 pt, double E, double wgt, int n_progeny,
                          uint64_t seed_trk, uint64_t seed_src, uint64_t seed_urr)
{
  const unsigned long long k = atomicAdd(&P.counters->n_trace, 1ull);
  if ((int64_t)k >= P.trace_capacity)
    return;
  mcw_trace_point r;
  const int64_t id = P.work_offset + i + 1;
  r.ray_id = (int64_t)(P.batch - 1) * P.n_particles_global + (id - 1);
  r.seq = seq;
  r.kind = kind;
  r.point[0] = pt.x, r.point[1] = pt.y, r.point[2] = pt.z;
  r.aux[0] = E, r.aux[1] = wgt, r.aux[2] = (double)n_progeny, r.aux[3] = (double)id;
  r.aux[4] = (double)seed_trk, r.aux[5] = (double)seed_src, r.aux[6] = (double)seed_urr, r.aux[7] = 0.0;
  P.trace[k] = r;
}

// ------------------------------------------------------------------ segment part of onSegment
// MC CollisionKernel.C:143-175.  Returns true if a collision happens inside the segment, and then
// `pos` is the collision point (computed backwards from the exit point, :173-175).
struct SegState
{
  double E, wgt;
  uint64_t seed_trk, seed_src, seed_urr;
  double sig_t, sig_a, sig_f, sig_nf;
  double k_trk;
};

// `neg_log` = -log(prn(seed_trk)), with s.seed_trk already advanced by that draw (the caller may
// compute it ahead of the geometry to overlap it with the element-record load).
MCW_D bool mcw_segment_nl(const McwParams& P, SegState& s, double neg_log, int elem, int block, int i_mat, V3 end, V3 dir,
                          double L, V3& pos)
{
  const double d = neg_log / s.sig_t; // 1 RN per segment (:150): -log(prn) / Sigma_t
  const double dmin = (L < d) ? L : d;                       // std::min(d, L)
  s.k_trk += s.wgt * dmin * s.sig_nf;                        // :153
  if (P.tallies_active && P.any_tracklength)
  {
    // OMC tally_scoring.cpp:2289-2352 score_tracklength_tally (extension)
    double flux = s.wgt * dmin;
    double sc[MCW_N_SCORES];
    sc[MCW_SCORE_FLUX] = flux;
    sc[MCW_SCORE_TOTAL] = s.sig_t * flux;
    sc[MCW_SCORE_FISSION] = s.sig_a == 0 ? 0.0 : s.sig_f * flux;
    sc[MCW_SCORE_NU_FISSION] = s.sig_a == 0 ? 0.0 : s.sig_nf * flux;
    sc[MCW_SCORE_KAPPA_FISSION] = 0.0;
    for (int t = 0; t < P.n_tallies; ++t)
      if (P.tally[t].estimator == MCW_ESTIMATOR_TRACKLENGTH)
      {
        bool need_kappa = false;
        for (int q = 0; q < P.tally[t].n_scores; ++q)
          need_kappa |= P.tally[t].scores[q] == MCW_SCORE_KAPPA_FISSION;
        if (need_kappa)
          sc[MCW_SCORE_KAPPA_FISSION] = s.sig_a == 0. ? 0.0 : mcw_kappa_macro(P.lib, i_mat, s.E, flux, s.seed_urr);
        if (MCW_RARE && P.tally[t].nuclides)
          mcw_score_tally_nuclides(P.lib.self_dev, P.tally_dev + t, mcw_spatial_bin(P.tally[t], elem, block), s.E, i_mat, flux,
                                   s.sig_a, sc, s.seed_urr);
        else
          mcw_score_tally(P.tally[t], mcw_spatial_bin(P.tally[t], elem, block), s.E, sc);
      }
  }
  // :157-159 store all three seeds as doubles.  The source and URR seeds did not change since they
  // were last rounded (at initialisation or at the end of the last collision) and the rounding is
  // idempotent, so only the tracking seed needs it here.
  if (P.seed_through_double)
    s.seed_trk = mcw_seed_via_double(s.seed_trk);
  if (d > L)
    return false;
  pos = end - (L - d) * dir;
  if (P.is_2d)
    pos.z = P.z_coord;
  return true;
}

MCW_D bool mcw_segment(const McwParams& P, SegState& s, int elem, int block, int i_mat, V3 end, V3 dir, double L, V3& pos)
{
  const double neg_log = -mcw_log(mcw_prn(s.seed_trk));
  return mcw_segment_nl(P, s, neg_log, elem, block, i_mat, end, dir, L, pos);
}

// ------------------------------------------------------------------ collision physics
struct ColState
{
  V3 r, u;
  double E, wgt;
  uint64_t seed; // tracking stream
  int n_progeny;
};

// ---- thermal scattering S(a,b): OMC secondary_thermal.cpp; all tables live in McwLibraryDev::sab_pool
// get_energy_index (:14-25)
MCW_D void mcw_sab_energy_index(const double* __restrict__ e, int n, double E, int& i, double& f)
{
  i = 0;
  f = 0.0;
  if (E >= e[0])
  {
    i = mcw_lower_bound_index(e, n, E);
    f = (E - e[i]) / (e[i + 1] - e[i]);
  }
}

// sab_scatter (OMC physics.cpp:788-802) over ThermalData::sample (thermal.cpp:245-262): CoherentElasticAE (:36-58),
// IncoherentElasticAEDiscrete (:96-126), IncoherentInelasticAEDiscrete (:138-188), IncoherentInelasticAE (:236-332).
// Out of line: only libraries with S(a,b) tables come here.
__device__ __noinline__ void mcw_sab_scatter(const McwLibraryDev* Lg, int i_sab, double thermal, double thermal_elastic,
                                             ColState& p)
{
  const McwLibraryDev& L = *Lg; // the global-memory copy (McwLibraryDev::self_dev)
  const McwSabDev& t = L.sab[i_sab];
  const double* P = L.sab_pool;
  const double E_in = p.E;
  double E_out, mu, f;
  int i;
  if (mcw_prn(p.seed) < thermal_elastic / thermal)
  {
    if (t.el_type == MCW_SAB_ELASTIC_COHERENT)
    {
      const double* edges = P + t.el_x;
      const double* factors = P + t.el_y;
      mcw_sab_energy_index(edges, t.n_el, E_in, i, f);
      const double prob = mcw_prn(p.seed) * factors[i + 1];
      int k = 0;
      if (prob >= factors[0])
        k = mcw_lower_bound_index(factors, i + 1, prob);
      mu = 1.0 - 2.0 * edges[k] / E_in;
    }
    else
    {
      mcw_sab_energy_index(P + t.el_x, t.n_el, E_in, i, f);
      const int n_mu = t.n_el_mu;
      const int k = (int)(mcw_prn(p.seed) * n_mu);
      const double* m0 = P + t.el_mu + (size_t)i * n_mu;
      const double* m1 = m0 + n_mu;
      mu = m0[k] + f * (m1[k] - m0[k]);
      const double mu_left = (k == 0) ? -1.0 - (mu + 1.0) : m0[k - 1] + f * (m1[k - 1] - m0[k - 1]);
      const double mu_right = (k == n_mu - 1) ? 1.0 + (1.0 - mu) : m0[k + 1] + f * (m1[k + 1] - m0[k + 1]);
      mu += fmin(mu - mu_left, mu_right - mu) * (mcw_prn(p.seed) - 0.5);
    }
    E_out = E_in;
  }
  else if (t.inel_type == MCW_SAB_INELASTIC_DISCRETE)
  {
    mcw_sab_energy_index(P + t.ixs_e, t.n_ixs, E_in, i, f);
    int j;
    const int n = t.n_e_out;
    if (!t.skewed)
      j = (int)(mcw_prn(p.seed) * n);
    else
    {
      const double r = mcw_prn(p.seed) * (n - 3);
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
    const double E_ij = P[t.eout + (size_t)i * n + j];
    const double E_i1j = P[t.eout + (size_t)(i + 1) * n + j];
    E_out = (1 - f) * E_ij + f * E_i1j;
    const int m = t.n_mu;
    const int k = (int)(mcw_prn(p.seed) * m);
    const double mu_ijk = P[t.mu + ((size_t)i * n + j) * m + k];
    const double mu_i1jk = P[t.mu + ((size_t)(i + 1) * n + j) * m + k];
    mu = (1 - f) * mu_ijk + f * mu_i1jk;
  }
  else
  {
    const double* ien = P + t.ien;
    mcw_sab_energy_index(ien, t.n_ien, E_in, i, f);
    const int l = f > 0.5 ? i + 1 : i;
    const int b = L.sab_ipool[t.ioff + l];
    const int n = L.sab_ipool[t.ioff + l + 1] - b;
    const double* e_out = P + t.eout + b;
    const double* pdf = P + t.pdf + b;
    const double* cdf = P + t.cdf + b;
    const double r1 = mcw_prn(p.seed);
    double c_j = cdf[0], c_j1 = 0.0;
    int j;
    for (j = 0; j < n - 1; ++j)
    {
      c_j1 = cdf[j + 1];
      if (r1 < c_j1)
        break;
      c_j = c_j1;
    }
    j = min(j, n - 2);
    const double E_l_j = e_out[j], p_l_j = pdf[j];
    const double E_l_j1 = e_out[j + 1], p_l_j1 = pdf[j + 1];
    const double frac = (p_l_j1 - p_l_j) / (E_l_j1 - E_l_j);
    if (frac == 0.0)
      E_out = E_l_j + (r1 - c_j) / p_l_j;
    else
      E_out = E_l_j + (sqrt(fmax(0.0, p_l_j * p_l_j + 2.0 * frac * (r1 - c_j))) - p_l_j) / frac;
    const double E_l = ien[l];
    if (E_out < 0.5 * E_l)
      E_out *= 2.0 * E_in / E_l - 1.0;
    else
      E_out += E_in - E_l;
    const int n_mu = t.n_mu;
    const int k = (int)(mcw_prn(p.seed) * n_mu);
    const double* mu_j = P + t.mu + ((size_t)b + j) * n_mu;
    const double* mu_j1 = mu_j + n_mu;
    f = (r1 - c_j) / (c_j1 - c_j);
    mu = mu_j[k] + f * (mu_j1[k] - mu_j[k]);
    const double mu_left = (k == 0) ? -1.0 - (mu + 1.0) : mu_j[k - 1] + f * (mu_j1[k - 1] - mu_j[k - 1]);
    const double mu_right = (k == n_mu - 1) ? 1.0 + (1.0 - mu) : mu_j[k + 1] + f * (mu_j1[k + 1] - mu_j[k + 1]);
    mu += fmin(mu - mu_left, mu_right - mu) * (mcw_prn(p.seed) - 0.5);
  }
  if (fabs(mu) > 1.0)
    mu = copysign(1.0, mu);
  p.E = E_out;
  p.u = mcw_rotate_angle(p.u, mu, p.seed);
}

// OMC physics.cpp:955-1004
MCW_D V3 mcw_sample_cxs_target_velocity(double awr, double E, V3 u, double kT, uint64_t& seed)
{
  double beta_vn = sqrt(awr * E / kT);
  double alpha = 1.0 / (1.0 + sqrt(MCW_PI) * beta_vn / 2.0);
  double beta_vt_sq, mu;
  while (true)
  {
    double r1 = mcw_prn(seed), r2 = mcw_prn(seed);
    if (mcw_prn(seed) < alpha)
      beta_vt_sq = -mcw_log_c(r1 * r2);
    else
    {
      double c = mcw_cos_c(MCW_PI / 2.0 * mcw_prn(seed));
      beta_vt_sq = -mcw_log_c(r1) - mcw_log_c(r2) * c * c;
    }
    double beta_vt = sqrt(beta_vt_sq);
    mu = mcw_uniform(-1., 1., seed);
    double accept_prob =
        sqrt(beta_vn * beta_vn + beta_vt_sq - 2 * beta_vn * beta_vt * mu) / (beta_vn + beta_vt);
    if (mcw_prn(seed) < accept_prob)
      break;
  }
  double vt = sqrt(beta_vt_sq * kT / awr);
  return vt * mcw_rotate_angle(u, mu, seed);
}

// ------------------------------------------------------------------ tabulated secondary distributions
// lower_bound_index (OMC search.h:14-20) on the incoming-energy grid of one table
MCW_D int mcw_tab_energy_bin(const double* __restrict__ en, int n, double E)
{
  if (__ldg(en) == E)
    return 0;
  int lo = 0, hi = n;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (__ldg(en + mid) < E)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

// Tabular::sample (OMC distribution.cpp:223-266) of distribution `e` of the pools
MCW_D double mcw_tabular_sample(const McwLibraryDev& L, int e, uint64_t& seed)
{
  const int b = __ldg(L.tab_off + e), n = __ldg(L.tab_end + e) - b;
  const double* __restrict__ x = L.tab_x + b;
  const double* __restrict__ pd = L.tab_p + b;
  const double* __restrict__ cd = L.tab_c + b;
  const double c = mcw_prn(seed);
  double c_i = __ldg(cd);
  int i;
  for (i = 0; i < n - 1; ++i)
  {
    const double c1 = __ldg(cd + i + 1);
    if (c <= c1)
      break;
    c_i = c1;
  }
  const double x_i = __ldg(x + i), p_i = __ldg(pd + i);
  if (__ldg(L.tab_interp + e) == 1)
    return p_i > 0.0 ? x_i + (c - c_i) / p_i : x_i;
  const double x_i1 = __ldg(x + i + 1), p_i1 = __ldg(pd + i + 1);
  const double m = (p_i1 - p_i) / (x_i1 - x_i);
  if (m == 0.0)
    return x_i + (c - c_i) / p_i;
  return x_i + (sqrt(fmax(0.0, p_i * p_i + 2 * m * (c - c_i))) - p_i) / m;
}

// AngleDistribution::sample (OMC distribution_angle.cpp:65-95)
MCW_D double mcw_angle_sample(const McwLibraryDev& L, int first, int n, double E, uint64_t& seed)
{
  const double* __restrict__ en = L.tab_energy + first;
  int i;
  double r;
  if (E < __ldg(en))
  {
    i = 0;
    r = 0.0;
  }
  else if (E > __ldg(en + n - 1))
  {
    i = n - 2;
    r = 1.0;
  }
  else
  {
    i = mcw_tab_energy_bin(en, n, E);
    r = (E - __ldg(en + i)) / (__ldg(en + i + 1) - __ldg(en + i));
  }
  if (r > mcw_prn(seed))
    ++i;
  double mu = mcw_tabular_sample(L, first + i, seed);
  if (fabs(mu) > 1.0)
    mu = copysign(1.0, mu);
  return mu;
}

// ContinuousTabular::sample (OMC distribution_energy.cpp:150-275)
MCW_D double mcw_continuous_tabular_sample(const McwLibraryDev& L, int first, int n_in, bool histogram_interp, double E,
                                           uint64_t& seed)
{
  const double* __restrict__ en = L.tab_energy + first;
  int i;
  double r;
  if (E < __ldg(en))
  {
    i = 0;
    r = 0.0;
  }
  else if (E > __ldg(en + n_in - 1))
  {
    i = n_in - 2;
    r = 1.0;
  }
  else
  {
    i = mcw_tab_energy_bin(en, n_in, E);
    r = (E - __ldg(en + i)) / (__ldg(en + i + 1) - __ldg(en + i));
  }
  int l;
  if (histogram_interp)
    l = i;
  else
    l = r > mcw_prn(seed) ? i + 1 : i;
  // interpolation for energy E1 and EK
  const int b_i = __ldg(L.tab_off + first + i), e_i = __ldg(L.tab_end + first + i);
  const int b_i1 = __ldg(L.tab_off + first + i + 1), e_i1 = __ldg(L.tab_end + first + i + 1);
  const double E_i_1 = __ldg(L.tab_x + b_i + __ldg(L.tab_ndisc + first + i)), E_i_K = __ldg(L.tab_x + e_i - 1);
  const double E_i1_1 = __ldg(L.tab_x + b_i1 + __ldg(L.tab_ndisc + first + i + 1)), E_i1_K = __ldg(L.tab_x + e_i1 - 1);
  const double E_1 = E_i_1 + r * (E_i1_1 - E_i_1);
  const double E_K = E_i_K + r * (E_i1_K - E_i_K);
  // outgoing energy bin of distribution l
  const int b = __ldg(L.tab_off + first + l), n_energy_out = __ldg(L.tab_end + first + l) - b;
  const int n_discrete = __ldg(L.tab_ndisc + first + l);
  const double* __restrict__ eo = L.tab_x + b;
  const double* __restrict__ pd = L.tab_p + b;
  const double* __restrict__ cd = L.tab_c + b;
  const double r1 = mcw_prn(seed);
  double c_k = __ldg(cd);
  int k = 0, end = n_energy_out - 2;
  for (int j = 0; j < n_discrete; ++j) // discrete portion
  {
    k = j;
    c_k = __ldg(cd + k);
    if (r1 < c_k)
    {
      end = j;
      break;
    }
  }
  for (int j = n_discrete; j < end; ++j) // continuous portion
  {
    k = j;
    const double c_k1 = __ldg(cd + k + 1);
    if (r1 < c_k1)
      break;
    k = j + 1;
    c_k = c_k1;
  }
  const double E_l_k = __ldg(eo + k), p_l_k = __ldg(pd + k);
  double E_out = E_l_k;
  if (__ldg(L.tab_interp + first + l) == 1)
  {
    if (p_l_k > 0.0 && k >= n_discrete)
      E_out = E_l_k + (r1 - c_k) / p_l_k;
  }
  else
  {
    const double E_l_k1 = __ldg(eo + k + 1), p_l_k1 = __ldg(pd + k + 1);
    if (E_l_k != E_l_k1)
    {
      const double frac = (p_l_k1 - p_l_k) / (E_l_k1 - E_l_k);
      if (frac == 0.0)
        E_out = E_l_k + (r1 - c_k) / p_l_k;
      else
        E_out = E_l_k + (sqrt(fmax(0.0, p_l_k * p_l_k + 2.0 * frac * (r1 - c_k))) - p_l_k) / frac;
    }
  }
  // interpolate between incident energy bins i and i + 1
  if (!histogram_interp && n_energy_out > 1 && k >= n_discrete)
  {
    if (l == i)
      return E_1 + (E_out - E_i_1) * (E_K - E_1) / (E_i_K - E_i_1);
    return E_1 + (E_out - E_i1_1) * (E_K - E_1) / (E_i1_K - E_i1_1);
  }
  return E_out;
}

// OMC physics.cpp:721-786: elastic scattering, angle from the nuclide's tabular distribution when it has one
// (:749-757) else isotropic in CM; free gas below 400 kT (:804-835)
MCW_D void mcw_elastic_scatter(const McwLibraryDev& L, const McwNucParams& nuc, ColState& p, bool use_ptable = false)
{
  double vel = sqrt(p.E);
  double awr = nuc.awr;
  V3 v_n = vel * p.u;
  V3 v_t = v3(0, 0, 0);
  if (!use_ptable && !(p.E >= MCW_FREE_GAS_THRESHOLD * nuc.kT && awr > 1.0))
    v_t = mcw_sample_cxs_target_velocity(awr, p.E, p.u, nuc.kT, p.seed);
  V3 v_cm = (v_n + awr * v_t) / (awr + 1.0);
  v_n = v_n - v_cm;
  vel = norm(v_n);
  double mu_cm;
  if (MCW_RARE && nuc.el_ang_n > 0)
    mu_cm = mcw_angle_sample(L, nuc.el_ang_first, nuc.el_ang_n, p.E, p.seed);
  else
    mu_cm = mcw_uniform(-1., 1., p.seed);
  V3 u_cm = v_n / vel;
  v_n = vel * mcw_rotate_angle(u_cm, mu_cm, p.seed);
  v_n = v_n + v_cm;
  p.E = dot(v_n, v_n);
  vel = sqrt(p.E);
  p.u = v_n / vel;
}

// WattEnergy / MaxwellEnergy / Evaporation::sample (OMC distribution_energy.cpp:342-356, 279-291, 305-322) with the
// parameters pa (theta, or Watt's a) and pb (Watt's b) already evaluated at E_in and the restriction energy u: fission
// products and uncorrelated scattering products.  Out of line.
__device__ __noinline__ double mcw_evaporation_type_sample(int law, double pa, double pb, double u, double E_in, uint64_t* seed_io)
{
  uint64_t seed = *seed_io;
  double Eo;
  if (law == MCW_LAW_WATT)
  {
    do
      Eo = mcw_watt_spectrum(pa, pb, seed);
    while (!(Eo <= E_in - u));
  }
  else if (law == MCW_LAW_MAXWELL)
  {
    do
      Eo = mcw_maxwell_spectrum(pa, seed);
    while (!(Eo <= E_in - u));
  }
  else
  {
    const double y = (E_in - u) / pa;
    const double v = 1.0 - mcw_exp(-y);
    double x;
    do
    {
      const double r1 = mcw_prn(seed), r2 = mcw_prn(seed);
      x = -mcw_log_c((1.0 - v * r1) * (1.0 - v * r2));
    } while (!(x <= y));
    Eo = x * pa;
  }
  *seed_io = seed;
  return Eo;
}

// The outgoing-energy part CorrelatedAngleEnergy::sample (OMC secondary_correlated.cpp:156-245) and KalbachMann::sample
// (secondary_kalbach.cpp:117-222) share: ContinuousTabular without the histogram-in-incoming-energy case and without its
// E_l_k != E_l_k1 guard.  Returns what the angular parts need.
struct McwEnergyPick
{
  double E_out, E_unscaled, r1, c_k, c_k1, E_l_k, E_l_k1;
  int point; // index of the chosen point within the table (distribution l, point k)
  bool lin_lin;
};
MCW_D McwEnergyPick mcw_correlated_energy_sample(const McwLibraryDev& L, int first, int n_in, double E, uint64_t& seed)
{
  McwEnergyPick o;
  const double* __restrict__ en = L.tab_energy + first;
  int i;
  double r;
  if (E < __ldg(en))
  {
    i = 0;
    r = 0.0;
  }
  else if (E > __ldg(en + n_in - 1))
  {
    i = n_in - 2;
    r = 1.0;
  }
  else
  {
    i = mcw_tab_energy_bin(en, n_in, E);
    r = (E - __ldg(en + i)) / (__ldg(en + i + 1) - __ldg(en + i));
  }
  const int l = r > mcw_prn(seed) ? i + 1 : i;
  const int b_i = __ldg(L.tab_off + first + i), e_i = __ldg(L.tab_end + first + i);
  const int b_i1 = __ldg(L.tab_off + first + i + 1), e_i1 = __ldg(L.tab_end + first + i + 1);
  const double E_i_1 = __ldg(L.tab_x + b_i + __ldg(L.tab_ndisc + first + i)), E_i_K = __ldg(L.tab_x + e_i - 1);
  const double E_i1_1 = __ldg(L.tab_x + b_i1 + __ldg(L.tab_ndisc + first + i + 1)), E_i1_K = __ldg(L.tab_x + e_i1 - 1);
  const double E_1 = E_i_1 + r * (E_i1_1 - E_i_1);
  const double E_K = E_i_K + r * (E_i1_K - E_i_K);
  const int b = __ldg(L.tab_off + first + l), n_energy_out = __ldg(L.tab_end + first + l) - b;
  const int n_discrete = __ldg(L.tab_ndisc + first + l);
  const double* __restrict__ eo = L.tab_x + b;
  const double* __restrict__ pd = L.tab_p + b;
  const double* __restrict__ cd = L.tab_c + b;
  const double r1 = mcw_prn(seed);
  double c_k = __ldg(cd);
  int k = 0, end = n_energy_out - 2;
  for (int j = 0; j < n_discrete; ++j)
  {
    k = j;
    c_k = __ldg(cd + k);
    if (r1 < c_k)
    {
      end = j;
      break;
    }
  }
  double c_k1 = 0.0;
  for (int j = n_discrete; j < end; ++j)
  {
    k = j;
    c_k1 = __ldg(cd + k + 1);
    if (r1 < c_k1)
      break;
    k = j + 1;
    c_k = c_k1;
  }
  const double E_l_k = __ldg(eo + k), p_l_k = __ldg(pd + k);
  double E_out;
  o.lin_lin = __ldg(L.tab_interp + first + l) != 1;
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
    const double E_l_k1 = __ldg(eo + k + 1), p_l_k1 = __ldg(pd + k + 1);
    o.E_l_k1 = E_l_k1;
    const double frac = (p_l_k1 - p_l_k) / (E_l_k1 - E_l_k);
    if (frac == 0.0)
      E_out = E_l_k + (r1 - c_k) / p_l_k;
    else
      E_out = E_l_k + (sqrt(fmax(0.0, p_l_k * p_l_k + 2.0 * frac * (r1 - c_k))) - p_l_k) / frac;
  }
  o.E_unscaled = E_out; // the Kalbach-Mann parameters interpolate with this value (secondary_kalbach.cpp:206-209)
  if (k >= n_discrete)
  {
    if (l == i)
      E_out = E_1 + (E_out - E_i_1) * (E_K - E_1) / (E_i_K - E_i_1);
    else
      E_out = E_1 + (E_out - E_i1_1) * (E_K - E_1) / (E_i1_K - E_i1_1);
  }
  o.E_out = E_out;
  o.r1 = r1, o.c_k = c_k, o.c_k1 = c_k1, o.E_l_k = E_l_k;
  o.point = b - __ldg(L.tab_off + first) + k;
  return o;
}

// The product distributions that give energy and cosine together (OMC reaction_product.cpp:70-79):
// CorrelatedAngleEnergy::sample (secondary_correlated.cpp:156-262), KalbachMann::sample (secondary_kalbach.cpp:117-235),
// NBodyPhaseSpace::sample (secondary_nbody.cpp:25-68).  Out of line: three rarely sampled laws.
__device__ __noinline__ double2 mcw_product_sample(const McwLibraryDev* Lp, int i_rx, double awr, double E_in, uint64_t* seed_io)
{
  const McwLibraryDev& L = *Lp;
  const McwRxParams& rx = L.rx[i_rx];
  uint64_t seed = *seed_io;
  double E_out, mu;
  if (rx.law == MCW_LAW_NBODY)
  {
    mu = mcw_uniform(-1., 1., seed);
    const double Ap = rx.nbody_mass_ratio;
    const double E_max = (Ap - 1.0) / Ap * (awr / (awr + 1.0) * E_in + rx.nbody_q);
    const double x = mcw_maxwell_spectrum(1.0, seed);
    double y;
    if (rx.nbody_n == 3)
      y = mcw_maxwell_spectrum(1.0, seed);
    else if (rx.nbody_n == 4)
    {
      const double r1 = mcw_prn(seed), r2 = mcw_prn(seed), r3 = mcw_prn(seed);
      y = -mcw_log_c(r1 * r2 * r3);
    }
    else
    {
      const double r1 = mcw_prn(seed), r2 = mcw_prn(seed), r3 = mcw_prn(seed), r4 = mcw_prn(seed), r5 = mcw_prn(seed),
                   r6 = mcw_prn(seed);
      const double c = mcw_cos_c(MCW_PI / 2.0 * r6);
      y = -mcw_log_c(r1 * r2 * r3 * r4) - mcw_log_c(r5) * (c * c);
    }
    const double v = x / (x + y);
    E_out = E_max * v;
  }
  else
  {
    const McwEnergyPick o = mcw_correlated_energy_sample(L, rx.en_first, rx.en_n, E_in, seed);
    E_out = o.E_out;
    if (rx.law == MCW_LAW_CORRELATED)
    {
      // the angular distribution of the closer outgoing-energy point (:247-261)
      if (o.r1 - o.c_k < o.c_k1 - o.r1 || !o.lin_lin)
        mu = mcw_tabular_sample(L, rx.law_first + o.point, seed);
      else
        mu = mcw_tabular_sample(L, rx.law_first + o.point + 1, seed);
    }
    else
    {
      // Kalbach-Mann: precompound fraction r and slope a of the sampled point (:189-234)
      const int b = __ldg(L.tab_off + rx.law_first) + o.point;
      const double r_k = __ldg(L.tab_x + b), a_k = __ldg(L.tab_p + b);
      double km_r = r_k, km_a = a_k;
      if (o.lin_lin)
      {
        const double t = (o.E_unscaled - o.E_l_k) / (o.E_l_k1 - o.E_l_k);
        km_r = r_k + t * (__ldg(L.tab_x + b + 1) - r_k);
        km_a = a_k + t * (__ldg(L.tab_p + b + 1) - a_k);
      }
      if (mcw_prn(seed) > km_r)
      {
        const double T = mcw_uniform(-1., 1., seed) * mcw_sinh(km_a);
        mu = mcw_log_c(T + sqrt(T * T + 1.0)) / km_a;
      }
      else
      {
        const double r1 = mcw_prn(seed);
        mu = mcw_log_c(r1 * mcw_exp(km_a) + (1.0 - r1) * mcw_exp(-km_a)) / km_a;
      }
    }
  }
  *seed_io = seed;
  return make_double2(E_out, mu);
}

// OMC physics.cpp:1073-1120 inelastic_scatter over UncorrelatedAngleEnergy::sample (secondary_uncorrelated.cpp:54-68):
// cosine from the reaction's tabular angle distribution or uniform, outgoing energy from LevelInelastic
// (distribution_energy.cpp:48-51) or ContinuousTabular (:150-275), then CM -> LAB if the reaction scatters in CM
MCW_D void mcw_inelastic_scatter(const McwLibraryDev& L, const McwNucParams& nuc, ColState& p, int j_rx = 0)
{
  double E_in = p.E;
  double mu, E;
  bool in_cm = true;
  if (!MCW_RARE)
  {
    // the common build: one level per nuclide, isotropic in the centre of mass, yield 1
    mu = mcw_uniform(-1., 1., p.seed);
    E = nuc.inel_ratio * (E_in - nuc.inel_E);
  }
  else
  {
    const int i_rx = nuc.rx_first + j_rx;
    const McwRxParams& rx = L.rx[i_rx];
    if (rx.law >= MCW_LAW_CORRELATED && rx.law <= MCW_LAW_NBODY)
    {
      const double2 em = mcw_product_sample(L.self_dev, i_rx, nuc.awr, E_in, &p.seed);
      E = em.x;
      mu = em.y;
    }
    else
    {
      if (rx.ang_n > 0)
        mu = mcw_angle_sample(L, rx.ang_first, rx.ang_n, E_in, p.seed);
      else
        mu = mcw_uniform(-1., 1., p.seed);
      if (rx.law == MCW_LAW_CONTINUOUS_TABULAR)
        E = mcw_continuous_tabular_sample(L, rx.en_first, rx.en_n, rx.en_hist != 0, E_in, p.seed);
      else if (rx.law >= MCW_LAW_MAXWELL)
      {
        const double pa = mcw_fn_eval(L, rx.a, E_in);
        const double pb = rx.law == MCW_LAW_WATT ? mcw_fn_eval(L, rx.b, E_in) : 0.0;
        E = mcw_evaporation_type_sample(rx.law, pa, pb, rx.u, E_in, &p.seed);
      }
      else
        E = rx.level_ratio * (E_in - rx.level_E);
    }
    in_cm = rx.in_cm != 0; // the level law always scatters in the centre of mass
    // the product's yield (OMC physics.cpp:1108-1118): integral = secondaries, which MaCaw drops (MC
    // CollisionKernel.C:232-236); otherwise the weight is multiplied
    const double yield = rx.yield_fn.kind != MCW_FN_NONE ? mcw_fn_eval(L, rx.yield_fn, E_in) : rx.yield;
    if (floor(yield) != yield)
      p.wgt *= yield;
  }
  if (in_cm)
  {
    double E_cm = E;
    double A = nuc.awr;
    E = E_cm + (E_in + 2.0 * mu * (A + 1.0) * sqrt(E_in * E_cm)) / ((A + 1.0) * (A + 1.0));
    mu = mu * sqrt(E_cm / E) + 1.0 / (A + 1.0) * sqrt(E_in / E);
  }
  if (fabs(mu) > 1.0)
    mu = copysign(1.0, mu);
  p.E = E;
  p.u = mcw_rotate_angle(p.u, mu, p.seed);
}

// Which of the nuclide's scattering reactions a non-elastic scatter is (OMC physics.cpp:676-697): in list order, every
// reaction above its threshold adds its cross section at (i_grid, f) until the sum passes the cutoff.  inel0 / inel1 =
// the first reaction's values at i_grid and i_grid + 1 (el_inel).  -1: none is above its threshold.  Should rounding
// leave the cutoff above the sum the last eligible reaction is taken (the reference would run off its list there).
__device__ __noinline__ int mcw_pick_reaction(const McwLibraryDev* Lp, int i_nuc, int i_grid, double f, double elastic, double inel0,
                                              double inel1, double cut)
{
  const McwLibraryDev& L = *Lp;
  const McwNucParams& nuc = L.nuc[i_nuc];
  double prob = elastic;
  int pick = -1;
  for (int j = 0; j <= nuc.n_rx_more; ++j)
  {
    const McwRxParams& rx = L.rx[nuc.rx_first + j];
    if (i_grid < rx.thr)
      continue;
    pick = j;
    double v0 = inel0, v1 = inel1;
    if (j > 0)
    {
      const double* v = L.rx_xs + rx.xs_off + (i_grid - rx.thr);
      v0 = __ldg(v);
      v1 = __ldg(v + 1);
    }
    prob += (1.0 - f) * v0 + f * v1;
    if (!(prob < cut))
      break;
  }
  return pick;
}

// sample_fission_neutron (OMC physics.cpp:1006-1071) on fission data as a nuclide file has them (mcw_fission_product):
// beta from the yield functions, the group by the delayed yields, the product's own energy law -- WattEnergy / MaxwellEnergy
// / Evaporation::sample with tabulated parameters and restriction energy (distribution_energy.cpp:342-356, 279-291,
// 305-322) or ContinuousTabular -- resampled while E >= E_max.  Returns (E, mu), the group in *group_out.  Out of line.
__device__ __noinline__ double2 mcw_fission_neutron_general(const McwLibraryDev* Lp, int i_nuc, double E_in, int* group_out,
                                                            uint64_t* seed_io)
{
  const McwLibraryDev& L = *Lp;
  const McwNucParams& nuc = L.nuc[i_nuc];
  const McwFisProduct* fp = L.fis + nuc.fis_first;
  uint64_t seed = *seed_io;
  const double nu_t = mcw_nu_total(L, nuc, E_in);
  double nu_d = 0.0;
  for (int g = 1; g <= nuc.n_precursor; ++g)
    nu_d += mcw_fn_eval(L, fp[g].yield, E_in);
  const double beta = nu_d / nu_t;
  int group = 0;
  if (mcw_prn(seed) < beta)
  {
    const double xi = mcw_prn(seed) * nu_d;
    double prob = 0.0;
    for (group = 1; group < nuc.n_precursor; ++group)
    {
      prob += mcw_fn_eval(L, fp[group].yield, E_in);
      if (xi < prob)
        break;
    }
    group = min(group, nuc.n_precursor);
  }
  const McwFisProduct& f = fp[group];
  double pa = 0.0, pb = 0.0; // the law's parameters at E_in
  if (f.law != MCW_LAW_CONTINUOUS_TABULAR)
  {
    pa = mcw_fn_eval(L, f.a, E_in);
    if (f.law == MCW_LAW_WATT)
      pb = mcw_fn_eval(L, f.b, E_in);
  }
  double mu, Eo;
  while (true)
  {
    mu = mcw_uniform(-1., 1., seed); // secondary_uncorrelated.cpp:54-68
    if (f.law == MCW_LAW_CONTINUOUS_TABULAR)
      Eo = mcw_continuous_tabular_sample(L, f.en_first, f.en_n, f.en_hist != 0, E_in, seed);
    else
      Eo = mcw_evaporation_type_sample(f.law, pa, pb, f.u, E_in, &seed);
    if (Eo < L.E_max)
      break;
  }
  *seed_io = seed;
  *group_out = group;
  return make_double2(Eo, mu);
}

// OMC physics.cpp:1006-1071
MCW_D void mcw_sample_fission_neutron(const McwLibraryDev& L, const McwNucParams& nuc, double E_in, double E_max,
                                      mcw_source_site& site, uint64_t& seed)
{
  if (MCW_RARE && nuc.fis_first >= 0)
  {
    int group;
    const double2 em = mcw_fission_neutron_general(L.self_dev, (int)(&nuc - L.nuc), E_in, &group, &seed);
    site.delayed_group = group;
    site.E = em.x;
    V3 u = mcw_rotate_angle(v3(1., 0., 0.), em.y, seed);
    site.u[0] = u.x;
    site.u[1] = u.y;
    site.u[2] = u.z;
    return;
  }
  double nu_t = 0.0;
  nu_t = nu_t * E_in + nuc.nu_poly[1]; // OMC endf.cpp:112-121 Horner
  nu_t = nu_t * E_in + nuc.nu_poly[0];
  double nu_d = 0.0;
  for (int g = 0; g < nuc.n_precursor; ++g)
    nu_d += nuc.dyield[g];
  double beta = nu_d / nu_t;
  int group = 0;
  if (mcw_prn(seed) < beta)
  {
    double xi = mcw_prn(seed) * nu_d;
    double prob = 0.0;
    for (group = 1; group < nuc.n_precursor; ++group)
    {
      prob += nuc.dyield[group - 1];
      if (xi < prob)
        break;
    }
    group = min(group, nuc.n_precursor);
  }
  site.delayed_group = group;
  double mu, Eo;
  while (true)
  {
    mu = mcw_uniform(-1., 1., seed); // secondary_uncorrelated.cpp:54-68
    if (group == 0)
      Eo = mcw_watt_spectrum(nuc.watt_a, nuc.watt_b, seed);
    else
      Eo = mcw_maxwell_spectrum(nuc.dtheta[group - 1], seed);
    if (Eo < E_max)
      break;
  }
  site.E = Eo;
  V3 u = mcw_rotate_angle(v3(1., 0., 0.), mu, seed);
  site.u[0] = u.x;
  site.u[1] = u.y;
  site.u[2] = u.z;
}
