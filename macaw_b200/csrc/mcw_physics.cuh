// mcw_physics.cuh -- per-segment and per-collision physics (device).  Mirrors
// MC CollisionKernel.C:98-275 split at the collision point: the part before `if (d > L)` runs in
// the tracking kernel, event_collide() runs in the collision kernel.
#pragma once
#include "mcw_device.cuh"

// ------------------------------------------------------------------ tallies
// OMC filter_energy.cpp:61-75 energy bin + tally_scoring.cpp:1453 atomic update.  Lanes of a warp
// that hit the same bin are pre-reduced with __match_any_sync so each distinct bin gets one FP64
// atomic per warp.
MCW_D void mcw_score_tally(const McwTallyDev& t, int64_t spatial, double E, const double* sc)
{
  int eb = 0;
  bool ok = true;
  if (t.n_edges >= 2)
  {
    const double* b = t.ebins;
    if (!(E >= b[0] && E <= b[t.n_edges - 1]))
      ok = false;
    else if (b[0] == E)
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
  long long bin = ok ? (long long)(spatial * t.n_energy + eb) : -1ll;
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

MCW_D int64_t mcw_spatial_bin(const McwTallyDev& t, int elem, int block)
{
  return t.granularity == MCW_FILTER_CELL ? (int64_t)elem
                                          : (t.granularity == MCW_FILTER_UNIVERSE ? (int64_t)block : 0);
}

// kappa-fission macro: sum_i N_i q_i sigma_f,i (OMC tally_scoring.cpp:1240-1255)
MCW_D double mcw_kappa_macro(const McwLibraryDev& L, int i_mat, double E)
{
  int i_log = mcw_log_bin(L, E);
  double kappa = 0.0;
  for (int i = L.mat_off[i_mat]; i < L.mat_off[i_mat + 1]; ++i)
  {
    int in = L.mat_nuc[i];
    if (L.nuc[in].fissionable)
    {
      Micro m = mcw_micro_xs(L, in, E, i_log);
      kappa += L.nuc[in].q * m.fission * L.mat_dens[i];
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
          sc[MCW_SCORE_KAPPA_FISSION] = s.sig_a == 0. ? 0.0 : mcw_kappa_macro(P.lib, i_mat, s.E) * flux;
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
      beta_vt_sq = -mcw_log(r1 * r2);
    else
    {
      double c = mcw_cos(MCW_PI / 2.0 * mcw_prn(seed));
      beta_vt_sq = -mcw_log(r1) - mcw_log(r2) * c * c;
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

// OMC physics.cpp:721-786, isotropic-in-CM elastic, free gas below 400 kT (:804-835)
MCW_D void mcw_elastic_scatter(const McwNucParams& nuc, ColState& p)
{
  double vel = sqrt(p.E);
  double awr = nuc.awr;
  V3 v_n = vel * p.u;
  V3 v_t = v3(0, 0, 0);
  if (!(p.E >= MCW_FREE_GAS_THRESHOLD * nuc.kT && awr > 1.0))
    v_t = mcw_sample_cxs_target_velocity(awr, p.E, p.u, nuc.kT, p.seed);
  V3 v_cm = (v_n + awr * v_t) / (awr + 1.0);
  v_n = v_n - v_cm;
  vel = norm(v_n);
  double mu_cm = mcw_uniform(-1., 1., p.seed);
  V3 u_cm = v_n / vel;
  v_n = vel * mcw_rotate_angle(u_cm, mu_cm, p.seed);
  v_n = v_n + v_cm;
  p.E = dot(v_n, v_n);
  vel = sqrt(p.E);
  p.u = v_n / vel;
}

// OMC physics.cpp:1073-1120 with LevelInelastic (distribution_energy.cpp:48-51), CM system
MCW_D void mcw_inelastic_scatter(const McwNucParams& nuc, ColState& p)
{
  double E_in = p.E;
  double mu = mcw_uniform(-1., 1., p.seed);
  double E = nuc.inel_ratio * (E_in - nuc.inel_E);
  double E_cm = E;
  double A = nuc.awr;
  E = E_cm + (E_in + 2.0 * mu * (A + 1.0) * sqrt(E_in * E_cm)) / ((A + 1.0) * (A + 1.0));
  mu = mu * sqrt(E_cm / E) + 1.0 / (A + 1.0) * sqrt(E_in / E);
  if (fabs(mu) > 1.0)
    mu = copysign(1.0, mu);
  p.E = E;
  p.u = mcw_rotate_angle(p.u, mu, p.seed);
}

// OMC physics.cpp:1006-1071
MCW_D void mcw_sample_fission_neutron(const McwNucParams& nuc, double E_in, double E_max,
                                      mcw_source_site& site, uint64_t& seed)
{
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
