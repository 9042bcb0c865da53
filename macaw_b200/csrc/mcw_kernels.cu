// mcw_kernels.cu -- event-based sm_100a kernels of the particle-history path.
//
//   k_init_histories   source sites -> particle SoA            (MC OpenMCStudy.C:297-364, RT ClaimRays.C)
//   k_sample_source    batch-1 box source with rejection        (OMC source.cpp:143-231,325-338)
//   k_xs_lookup_array  macroscopic cross sections for (E, material) arrays
//                                                               (OMC material.cpp:763-831, nuclide.cpp:557-763)
//   k_track            element walk until collision / death
//                                                               (RT TraceRay.C:292-493,999-1592; MC CollisionKernel.C:143-175)
//   k_collide          event_collide + tallies + bank append    (OMC particle.cpp:219-288, physics.cpp:38-247,442-1120)
//   k_fallback         general vertex/edge/neighbour step       (RT TraceRay.C:564-997,1134-1246)
//   bank / tally / reduction helpers                            (OMC bank.cpp:60-104, eigenvalue.cpp:71-311, tally.cpp:639-670)
//
// One thread = one history slot; queues are index lists with warp-aggregated appends.
#include <algorithm>
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <math.h>

#include "mcw_detmath.h"
#include "mcw_kernels.h"

// MCW_RARE: 1 = with the rarely used branches (S(a,b), nuclide tally bins, cached trace points, tabulated secondary
// distributions), 0 = without them; one namespace each (mcw_kernels.h)
#ifndef MCW_RARE
#define MCW_RARE 1
#endif
#if MCW_RARE
#define MCW_NS mcw_r1
#else
#define MCW_NS mcw_r0
#endif

namespace MCW_NS
{
#include "mcw_physics.cuh"

#define MCW_BLOCK 128

// ------------------------------------------------------------------ init
template <int TYPE>
__global__ void __launch_bounds__(MCW_BLOCK)
k_init_histories(McwParams P, const mcw_source_site* __restrict__ src, int64_t n, int32_t* __restrict__ q_fly)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const mcw_source_site s = src[i];
  V3 r = v3(s.r[0], s.r[1], s.r[2]);
  V3 dir = v3(s.u[0], s.u[1], s.u[2]);
  if (P.is_2d)
    dir.z = 0;
  dir = unit(dir); // Ray::setStartingDirection (RT Ray.C:291)
  const int64_t id = P.work_offset + i + 1;                                      // OMC simulation.cpp:469
  const int64_t particle_seed = (int64_t)(P.batch - 1) * P.n_particles_global + id; // :477-480
  uint64_t seeds[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    seeds[k] = mcw_future_seed((uint64_t)particle_seed * MCW_PRN_STRIDE, (uint64_t)(P.master_seed + k));
    if (P.seed_through_double)
      seeds[k] = mcw_seed_via_double(seeds[k]); // MC OpenMCStudy.C:345-347
  }
  int elem = s.elem;
  if (elem < 0)
    elem = mcw_locate<TYPE>(P.mesh, r, (int64_t)(P.batch - 1) * P.n_particles_global + (id - 1));
  McwParticlesDev& p = P.p;
  p.x[i] = r.x;
  p.y[i] = r.y;
  p.z[i] = r.z;
  p.ux[i] = dir.x;
  p.uy[i] = dir.y;
  p.uz[i] = dir.z;
  p.E[i] = s.E;
  p.wgt[i] = s.wgt;
  p.seed_trk[i] = seeds[0];
  p.seed_src[i] = seeds[1];
  p.seed_urr[i] = seeds[2];
  p.k_col[i] = 0.0;
  p.k_abs[i] = 0.0;
  p.k_trk[i] = 0.0;
  p.n_progeny[i] = 0;
  p.n_seg[i] = 0;
  p.n_coll[i] = 0;
  p.n_xs[i] = 0;
  p.last_elem[i] = -1;
  p.ext_first[i] = -1;
  p.ext_second[i] = -1;
  p.hash[i] = 0xcbf29ce484222325ull;
  q_fly[i] = (int32_t)i;
  if (elem < 0)
  {
    p.elem[i] = -1;
    p.inc_side[i] = -1;
    p.status[i] = MCW_ST_DEAD;
    p.fate[i] = 2;
    atomicAdd(&P.counters->n_failed, 1ull);
    return;
  }
  ElemRec<TYPE> R(P.mesh, elem);
  p.elem[i] = elem;
  p.inc_side[i] = (int8_t)mcw_starting_incoming_side<TYPE>(R, r, dir);
  p.status[i] = MCW_ST_FLY;
  p.fate[i] = -1;
  if (MCW_RARE && P.trace)
    mcw_trace_emit(P, (int)i, 0, 0, r, s.E, s.wgt, 0, seeds[0], seeds[1], seeds[2]);
  // Particle::event_calculate_xs for the birth energy, every material slot at once
  mcw_lookup_all_slots(P, i, s.E);
  p.n_xs[i] = P.n_slots;
}

// OMC source.cpp:325-338 + :143-231 (box, isotropic, Watt 0.988 MeV / 2.249 per MeV)
template <int TYPE>
__global__ void __launch_bounds__(MCW_BLOCK)
k_sample_source(McwParams P, mcw_source_site* __restrict__ out, int64_t n, const double* box, int only_fissionable)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int64_t id = 0 * P.n_particles_global + P.work_offset + i + 1;
  uint64_t seed = mcw_future_seed((uint64_t)id * MCW_PRN_STRIDE, (uint64_t)(P.master_seed + 1));
  mcw_source_site site;
  site.wgt = 1.0;
  site.delayed_group = 0;
  site.surf_id = 0;
  site.particle = 0;
  site.parent_id = 0;
  site.progeny_id = 0;
  site.r[0] = site.r[1] = site.r[2] = 0.0;
  int64_t ray_id = P.work_offset + i;
  int e = -1;
  for (int tries = 0; tries < 1000000; ++tries)
  {
    double x1 = mcw_prn(seed), x2 = mcw_prn(seed), x3 = mcw_prn(seed);
    V3 r = v3(box[0] + x1 * (box[3] - box[0]), box[1] + x2 * (box[4] - box[1]), box[2] + x3 * (box[5] - box[2]));
    e = mcw_locate<TYPE>(P.mesh, r, ray_id);
    bool found = e >= 0;
    if (found && only_fissionable)
    {
      ElemRec<TYPE> R(P.mesh, e);
      found = P.lib.mat_fissionable[P.slot_mat[R.slot()]] != 0;
    }
    if (found)
    {
      site.r[0] = r.x;
      site.r[1] = r.y;
      site.r[2] = r.z;
      break;
    }
    e = -1;
  }
  if (e < 0)
    atomicAdd(&P.counters->n_failed, 1ull); // no position accepted: the host turns this into the reference's fatal error
  V3 u = mcw_isotropic_direction(seed);
  site.u[0] = u.x;
  site.u[1] = u.y;
  site.u[2] = u.z;
  while (true)
  {
    site.E = mcw_watt_spectrum(0.988e6, 2.249e-6, seed);
    if (site.E < P.lib.E_max && site.E > P.lib.E_min)
      break;
  }
  site.elem = e;
  out[i] = site;
}

// ------------------------------------------------------------------ cross-section lookup
// stand-alone lookup for (E, material) arrays: the XS-lookup bandwidth metric and unit parity
__global__ void __launch_bounds__(MCW_BLOCK)
k_xs_lookup_array(McwLibraryDev L, int64_t n, const double* __restrict__ E, const int32_t* __restrict__ mat,
                  double* __restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  Macro M = mcw_macro_xs(L, mat[i], E[i]);
  double2* o = reinterpret_cast<double2*>(out + 4 * i);
  o[0] = make_double2(M.total, M.absorption);
  o[1] = make_double2(M.fission, M.nu_fission);
}

#include "mcw_xs.cuh"

// resets the per-round device counters (one launch instead of three memsets)
__global__ void k_round_reset(McwCounters* c, int consumed, bool fallback_served)
{
  c->n_fly[consumed] = 0;
  c->n_collide = 0;
  if (fallback_served)
    c->n_fallback = 0;
  c->head = 0;
}

// ------------------------------------------------------------------ collision
// event_collide for history i (OMC particle.cpp:219-288).  Called by ALL 32 lanes of a warp (the
// fission-bank append is a full-warp aggregate); lanes with valid == false only take part in the
// shuffles.  Returns true when the history survives (scattered, status MCW_ST_FLY, cross sections
// of every material refreshed at the new energy).
// LOOKUP = false leaves the all-material cross-section refresh to k_xs_all_slots (one thread per
// (history, material) instead of one dependent chain of ~12 nuclide lookups per history).
template <int TYPE, bool LOOKUP>
MCW_D bool mcw_collide_history(const McwParams& P, int i, bool valid)
{
  const McwParticlesDev& p = P.p;
  const McwLibraryDev& L = P.lib;

  ColState c;
  double sig_t = 1, sig_a = 0, sig_f = 0, sig_nf = 0, E_last = 0, wgt_last = 0, kappa = 0.0;
  int i_mat = 0, elem = 0, block = 0, k_nuc = 0, q_nuc = 0, nu = 0;
  Micro mx;
  mx.total = mx.absorption = mx.fission = mx.nu_fission = mx.f = 0;
  mx.i_grid = 0;
  double k_col = 0, k_abs = 0;
  uint64_t seed_src = 0, seed_urr = 0, seed_urr_pre = 0; // _pre: the URR seed the pre-collision cross sections were made with
  if (valid)
  {
    c.r = v3(p.x[i], p.y[i], p.z[i]);
    c.u = v3(p.ux[i], p.uy[i], p.uz[i]);
    c.E = p.E[i];
    c.wgt = p.wgt[i];
    c.seed = p.seed_trk[i];
    c.n_progeny = p.n_progeny[i];
    seed_src = p.seed_src[i];
    seed_urr = p.seed_urr[i];
    seed_urr_pre = seed_urr;
    elem = p.elem[i];
    ElemRec<TYPE> R(P.mesh, elem);
    block = R.block();
    const int slot = R.slot();
    {
      const Macro M = mcw_load_xs(p.xs, P.n_slots, i, slot);
      sig_t = M.total;
      sig_a = M.absorption;
      sig_f = M.fission;
      sig_nf = M.nu_fission;
      i_mat = P.slot_mat[slot];
    }
    E_last = c.E;
    wgt_last = c.wgt;
    k_col = p.k_col[i] + c.wgt * sig_nf / sig_t; // OMC particle.cpp:223-226
    k_abs = p.k_abs[i];
    // sample_nuclide (OMC physics.cpp:442-465).  Its partial sums `prob` are the running totals the last
    // cross-section refresh stored for this energy (McwParticlesDev::xs_prefix), so the material's nuclides
    // are not looked up a second time: scan the stored sums, four independent loads at a time.
    const double cutoff = mcw_prn(c.seed) * sig_t;
    const int pb = P.slot_plan_off[slot], pe = P.slot_plan_off[slot + 1];
    {
      const double* pf = p.xs_prefix + (size_t)i * P.n_plan_pad; // plain loads: the fused kernel writes these rows itself
      int k = pe - 1; // `q == e - 1`: the last nuclide takes what rounding leaves over
      for (int q0 = pb; q0 < pe; q0 += 4)
      {
        const double v0 = pf[q0], v1 = pf[min(q0 + 1, pe - 1)], v2 = pf[min(q0 + 2, pe - 1)], v3 = pf[min(q0 + 3, pe - 1)];
        const int hit = v0 >= cutoff ? 0 : (v1 >= cutoff ? 1 : (v2 >= cutoff ? 2 : (v3 >= cutoff ? 3 : 4)));
        if (hit < 4)
        {
          k = min(q0 + hit, pe - 1);
          break;
        }
      }
      k_nuc = P.plan_nuc[k];
      q_nuc = L.mat_off[i_mat] + (k - pb);
    }
    const int i_log = mcw_log_bin(L, c.E);
    // with the S(a,b) correction of the total and the probability tables where they apply
    mx = mcw_micro_xs_entry(L, q_nuc, c.E, i_log, (MCW_RARE && L.n_urr) ? &seed_urr : nullptr);
    // kappa-fission, only when a collision tally asks for it: q * sigma_f * N * flux summed nuclide by
    // nuclide over the fissionable nuclides of the material (OMC tally_scoring.cpp:1246-1255)
    if (P.tallies_active && P.collision_kappa && sig_a != 0.)
    {
      const double flux = c.wgt / sig_t;
      for (int q = pb; q < pe; ++q)
      {
        const int in = P.plan_nuc[q];
        if (L.nuc[in].fissionable)
        {
          const Micro m = (in == k_nuc) ? mx
                                        : mcw_micro_xs_entry(L, L.mat_off[i_mat] + (q - pb), c.E, i_log,
                                                             (MCW_RARE && L.n_urr) ? &seed_urr : nullptr);
          kappa += L.nuc[in].q * m.fission * P.plan_dens[q] * flux;
        }
      }
    }
    // create_fission_sites: number of sites (OMC physics.cpp:154-170)
    if (L.nuc[k_nuc].fissionable)
    {
      double nu_t = c.wgt / P.keff * 1.0 * mx.nu_fission / mx.total;
      nu = (int)nu_t;
      if (mcw_prn(c.seed) <= (nu_t - nu))
        ++nu;
    }
  }
  // warp-aggregated fission-bank append: one atomicAdd per warp for all sites of the warp
  unsigned long long base = 0;
  {
    const unsigned full = 0xffffffffu;
    int lane = threadIdx.x & 31;
    int incl = nu;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      int t = __shfl_up_sync(full, incl, o);
      if (lane >= o)
        incl += t;
    }
    int total = __shfl_sync(full, incl, 31);
    unsigned long long wbase = 0;
    if (lane == 0 && total > 0)
      wbase = atomicAdd(&P.counters->n_bank, (unsigned long long)total);
    wbase = __shfl_sync(full, wbase, 0);
    base = wbase + (unsigned long long)(incl - nu);
  }
  if (!valid)
    return false;
  const McwNucParams& nuc = L.nuc[k_nuc];
  for (int k = 0; k < nu; ++k)
  {
    mcw_source_site site;
    site.r[0] = c.r.x;
    site.r[1] = c.r.y;
    site.r[2] = c.r.z;
    site.particle = 0;
    site.wgt = 1. / 1.0;
    site.parent_id = P.work_offset + i + 1;
    site.progeny_id = c.n_progeny++;
    site.surf_id = 0;
    site.elem = elem;
    mcw_sample_fission_neutron(L, nuc, c.E, L.E_max, site, c.seed);
    const unsigned long long bank_slot = base + k;
    if ((int64_t)bank_slot < P.bank_capacity)
      P.bank[bank_slot] = site;
    else
    {
      // "The shared fission bank is full" (OMC physics.cpp:203-209): this site is not banked and no further
      // ones are sampled.  The id it took is handed back, so n_progeny counts banked sites only and the bank
      // sort's offsets (exclusive scan of n_progeny) stay inside the bank; the host reports the count.
      --c.n_progeny;
      atomicAdd(&P.counters->n_bank_dropped, (unsigned long long)(nu - k));
      break;
    }
  }
  // absorption (analog, OMC physics.cpp:592-625)
  bool absorbed = false;
  if (mx.absorption > 0.0)
  {
    if (mx.absorption > mcw_prn(c.seed) * mx.total)
    {
      k_abs += c.wgt * mx.nu_fission / mx.absorption;
      absorbed = true;
    }
  }
  if (!absorbed)
  {
    // scatter (OMC physics.cpp:627-719)
    double cut = mcw_prn(c.seed) * (mx.total - mx.absorption);
    const double2* ei = reinterpret_cast<const double2*>(L.el_inel + 2 * (nuc.off + mx.i_grid));
    double2 r0 = __ldg(ei), r1 = __ldg(ei + 1);
    double elastic = (1.0 - mx.f) * r0.x + mx.f * r1.x;
    bool can_inelastic = nuc.n_inel > 0 && mx.i_grid >= nuc.inel_thr;
    // S(a,b): micro.elastic as corrected by calculate_sab_xs, micro.thermal and micro.thermal_elastic, re-derived here
    // from the same grid index and interpolation factor instead of being carried through the absorption / fission code
    double thermal = 0.0, thermal_el = 0.0;
    const int i_sab = mcw_sab_of_entry(L, q_nuc, E_last);
    if (i_sab >= 0)
    {
      Micro scratch = mx;
      mcw_sab_apply(L, q_nuc, i_sab, E_last, scratch, thermal, thermal_el, elastic);
    }
    // probability tables: micro.elastic is the table's value and the target is taken at rest (physics.cpp:735-738);
    // re-derived from the smooth values at the pre-collision energy and URR seed (seed_urr advances only below)
    const bool use_ptable = mcw_urr_in_range(L, k_nuc, E_last);
    if (use_ptable)
    {
      Micro smooth = mcw_micro_xs(L, k_nuc, E_last, mcw_log_bin(L, E_last));
      mcw_urr_apply(L, k_nuc, E_last, seed_urr, smooth, elastic);
    }
    if (elastic - thermal > cut)
      mcw_elastic_scatter(L, nuc, c, use_ptable); // non-S(a,b) elastic (OMC physics.cpp:649-661)
    else if (i_sab >= 0 && elastic > cut)
      mcw_sab_scatter(L.self_dev, i_sab, thermal, thermal_el, c); // :663-672
    else if (MCW_RARE && nuc.n_rx_more > 0)
    {
      // a list of scattering reactions (:676-697)
      const int j_rx = mcw_pick_reaction(L.self_dev, k_nuc, mx.i_grid, mx.f, elastic, r0.y, r1.y, cut);
      if (j_rx >= 0)
        mcw_inelastic_scatter(L, nuc, c, j_rx);
      else
        mcw_elastic_scatter(L, nuc, c, use_ptable);
    }
    else if (can_inelastic)
      mcw_inelastic_scatter(L, nuc, c);
    else
      mcw_elastic_scatter(L, nuc, c, use_ptable);
    if (L.mat_iso[q_nuc]) // isotropic-in-lab override (OMC physics.cpp:706-718): phi, then mu
      c.u = mcw_isotropic_direction(c.seed);
    if (c.E != E_last) // OMC physics.cpp:140-144
      seed_urr = mcw_future_seed((uint64_t)L.n_identities, seed_urr);
  }
  // collision-estimator tallies (OMC tally_scoring.cpp:2354-2420)
  if (P.tallies_active && P.any_collision_tally)
  {
    double flux = wgt_last / sig_t;
    double sc[MCW_N_SCORES];
    sc[MCW_SCORE_FLUX] = flux;
    sc[MCW_SCORE_TOTAL] = sig_t * flux;
    sc[MCW_SCORE_FISSION] = sig_a == 0 ? 0.0 : sig_f * flux;
    sc[MCW_SCORE_NU_FISSION] = sig_a == 0 ? 0.0 : sig_nf * flux;
    sc[MCW_SCORE_KAPPA_FISSION] = kappa; // 0 when the material absorbs nothing or no collision tally scores it
    for (int t = 0; t < P.n_tallies; ++t)
      if (P.tally[t].estimator == MCW_ESTIMATOR_COLLISION)
      {
        if (MCW_RARE && P.tally[t].nuclides)
          mcw_score_tally_nuclides(L.self_dev, P.tally_dev + t, mcw_spatial_bin(P.tally[t], elem, block), E_last, i_mat, flux,
                                   sig_a, sc, seed_urr_pre);
        else
          mcw_score_tally(P.tally[t], mcw_spatial_bin(P.tally[t], elem, block), E_last, sc);
      }
  }
  uint64_t seed_trk = c.seed;
  if (P.seed_through_double)
  {
    seed_trk = mcw_seed_via_double(seed_trk); // MC CollisionKernel.C:217-219
    seed_src = mcw_seed_via_double(seed_src);
    seed_urr = mcw_seed_via_double(seed_urr);
  }
  p.seed_trk[i] = seed_trk;
  p.seed_src[i] = seed_src;
  p.seed_urr[i] = seed_urr;
  p.k_col[i] = k_col;
  p.k_abs[i] = k_abs;
  p.n_progeny[i] = c.n_progeny;
  p.n_coll[i] += 1;
  p.E[i] = c.E;
  if (MCW_RARE && c.wgt != wgt_last) // a reaction with a non-integral yield (OMC physics.cpp:1116-1118)
    p.wgt[i] = c.wgt;
  atomicAdd(&P.counters->n_collisions, 1ull);
  if (absorbed)
  {
    p.status[i] = MCW_ST_DEAD;
    p.fate[i] = 0;
    if (MCW_RARE && P.trace) // the last cached point of the ray: the end of the segment it was absorbed in
      mcw_trace_emit(P, i, p.n_seg[i], 3, v3(p.tx[i], p.ty[i], p.tz[i]), c.E, c.wgt, c.n_progeny, seed_trk, seed_src, seed_urr);
  }
  else
  {
    // changeRayStartDirection (RT RayKernelBase.C:62-97 -> Ray.C:216-226); MC CollisionKernel.C:198-205
    V3 nd = c.u;
    if (P.is_2d)
      nd.z = 0;
    nd = unit(nd);
    p.ux[i] = nd.x;
    p.uy[i] = nd.y;
    p.uz[i] = nd.z;
    p.inc_side[i] = -1;
    p.ext_first[i] = -1;
    p.ext_second[i] = -1;
    p.status[i] = MCW_ST_FLY;
    // Particle::event_calculate_xs at the new energy, every material slot at once (if the
    // collision kept the energy the sectors are rewritten with the same values)
    if (LOOKUP)
      mcw_lookup_all_slots(P, i, c.E);
    p.n_xs[i] += P.n_slots;
  }
  return !absorbed;
}

// Particle::event_calculate_xs for the survivors of k_collide (OMC particle.cpp:85-147,
// material.cpp:763-831): thread (h, s) computes material slot s at the new energy of the h-th
// history of the collision queue; the four slots of a history are neighbouring threads, so the
// 32-byte results form one 128-byte store per history.
__global__ void __launch_bounds__(MCW_BLOCK)
k_xs_all_slots(McwParams P, const int32_t* __restrict__ queue, const unsigned long long* __restrict__ count)
{
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t h = idx / P.n_slots;
  const int s = (int)(idx - h * P.n_slots);
  if (h >= (int64_t)*count)
    return;
  const int i = queue[h];
  if (P.p.status[i] != MCW_ST_FLY)
    return; // absorbed
  uint64_t seed_urr = 0;
  const bool urr = MCW_RARE && P.lib.n_urr != 0; // the history's URR stream picks the probability-table band
  if (urr)
    seed_urr = P.p.seed_urr[i];
  const Macro M = mcw_macro_xs(P.lib, P.slot_mat[s], P.p.E[i], P.p.xs_prefix + (size_t)i * P.n_plan_pad + P.slot_plan_off[s],
                               urr ? &seed_urr : nullptr);
  double2* o = reinterpret_cast<double2*>(P.p.xs + ((size_t)i * P.n_slots + s) * 4);
  o[0] = make_double2(M.total, M.absorption);
  o[1] = make_double2(M.fission, M.nu_fission);
}

// Particle::event_calculate_xs for the histories of `m_need` (lanes of this warp; lane l holds history
// i and its new energy E), done by the whole warp: the (slot, nuclide) entries of the refresh plan are
// dealt to the 32 lanes, each lane looks its nuclides up, and the products N_i * sigma_i are then added
// in plan order -- the order of Material::calculate_xs (OMC material.cpp:790-829) -- so the sums are
// bit-identical to one thread walking the material.  Used by the fused tracking kernel, where one lane
// doing ~40 dependent nuclide lookups would hold its whole warp for the length of that chain.
// Called by all 32 lanes.
MCW_D void mcw_warp_lookup_all_slots(const McwParams& P, unsigned m_need, int i, double E)
{
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  while (m_need)
  {
    const int src = __ffs(m_need) - 1;
    m_need &= m_need - 1;
    const int hi = __shfl_sync(full, i, src);
    const double he = __shfl_sync(full, E, src);
    const int i_log = mcw_log_bin(P.lib, he);
    double* xs_rows = P.p.xs + (size_t)hi * P.n_slots * 4;
    double* prefix = P.p.xs_prefix + (size_t)hi * P.n_plan_pad;
    double Mt = 0.0, Ma = 0.0, Mf = 0.0, Mn = 0.0;
    int cur_slot = 0;
    for (int base = 0; base < P.n_plan; base += 32)
    {
      const int q = base + lane;
      double t = 0.0, a = 0.0, f = 0.0, nf = 0.0, mine = 0.0;
      if (q < P.n_plan)
      {
        Micro m;
        if (MCW_RARE && (P.lib.n_sab | P.lib.n_urr))
        {
          // the plan entry's position in its material: S(a,b) correction of the total, probability tables
          const int sl = P.plan_slot[q];
          const int ent = P.lib.mat_off[P.slot_mat[sl]] + q - P.slot_plan_off[sl];
          const uint64_t hs = P.p.seed_urr[hi];
          m = mcw_micro_xs_entry(P.lib, ent, he, i_log, P.lib.n_urr ? &hs : nullptr);
        }
        else
          m = mcw_micro_xs(P.lib, P.plan_nuc[q], he, i_log);
        const double d = P.plan_dens[q];
        t = d * m.total;
        a = d * m.absorption;
        f = d * m.fission;
        nf = d * m.nu_fission;
      }
      const int n_here = min(32, P.n_plan - base);
      for (int k = 0; k < n_here; ++k)
      {
        const int sk = P.plan_slot[base + k];
        if (sk != cur_slot)
        {
          if (lane == 0)
          {
            double2* o = reinterpret_cast<double2*>(xs_rows + 4 * cur_slot);
            o[0] = make_double2(Mt, Ma);
            o[1] = make_double2(Mf, Mn);
          }
          Mt = Ma = Mf = Mn = 0.0;
          cur_slot = sk;
        }
        Mt += __shfl_sync(full, t, k);
        Ma += __shfl_sync(full, a, k);
        Mf += __shfl_sync(full, f, k);
        Mn += __shfl_sync(full, nf, k);
        if (lane == k)
          mine = Mt;
      }
      if (q < P.n_plan)
        prefix[q] = mine; // sample_nuclide's partial sum after this entry
    }
    if (lane == 0)
    {
      double2* o = reinterpret_cast<double2*>(xs_rows + 4 * cur_slot);
      o[0] = make_double2(Mt, Ma);
      o[1] = make_double2(Mf, Mn);
    }
  }
  __syncwarp(); // rows written by other lanes are read by the history's own lane next
}

#ifndef MCW_COLLIDE_MIN_BLOCKS
#define MCW_COLLIDE_MIN_BLOCKS 8 // 64 registers: the kernel waits on scattered loads, twice the warps beat the spills (A/B r01f)
#endif
template <int TYPE>
__global__ void __launch_bounds__(MCW_BLOCK, MCW_COLLIDE_MIN_BLOCKS)
k_collide(McwParams P, const int32_t* __restrict__ queue, const unsigned long long* __restrict__ count,
          int32_t* __restrict__ q_fly_next, unsigned long long* __restrict__ cnt_next, uint32_t* __restrict__ keys,
          int32_t* __restrict__ hist, uint32_t* __restrict__ bins)
{
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < (int64_t)*count;
  if (!__any_sync(0xffffffffu, active))
    return; // grids are sized from the host's last known queue length
  // no early return inside a warp: the bank append and the queue append are full-warp aggregates
  const int i = active ? queue[idx] : -1;
  const bool valid = active && P.p.status[i] == MCW_ST_COLLIDE;
  const bool survived = mcw_collide_history<TYPE, false>(P, i, valid);
  if (keys && active)
  {
    // energy bucket of the survivor for the binned cross-section refresh (mcw_xs.cuh), counted here
    const uint32_t k = survived ? mcw_refresh_key(P.p.E[i]) : MCW_KEY_DEAD;
    keys[idx] = k;
    hist[idx] = i;
    if (survived)
      atomicAdd(bins + k, 1u);
  }
  if (P.ordered_queues)
    return;
  unsigned long long k = mcw_warp_append(cnt_next, survived);
  if (survived)
    q_fly_next[k] = i;
}

// ------------------------------------------------------------------ tracking (may collide inline: after mcw_collide_history)
#include "mcw_track.cuh"

// ------------------------------------------------------------------ general (rare-path) step
#include "mcw_fallback.cuh"

// ------------------------------------------------------------------ persistent generation kernel
#include "mcw_pump.cuh"

// ------------------------------------------------------------------ bank helpers
// OMC bank.cpp:60-104: position of a site in the sorted bank = exclusive_scan(progeny_per_particle)
// [parent] + progeny_id
__global__ void k_bank_scatter(const mcw_source_site* __restrict__ bank, int64_t n_bank,
                               const int64_t* __restrict__ progeny_scan, int64_t work_offset,
                               mcw_source_site* __restrict__ sorted)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bank)
    return;
  mcw_source_site s = bank[i];
  int64_t off = s.parent_id - 1 - work_offset;
  const int64_t dst = progeny_scan[off] + s.progeny_id;
  if (dst >= 0 && dst < n_bank) // n_progeny counts banked sites only, so this always holds; never write outside the bank
    sorted[dst] = s;
}

__global__ void k_progeny_to_i64(const int32_t* __restrict__ n_progeny, int64_t n, int64_t* __restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = n_progeny[i];
}

// OMC eigenvalue.cpp:112-160: how many copies of bank site i go to the next source
__global__ void k_sample_counts(int64_t n_bank, int64_t global_start, int64_t total, int64_t n_particles,
                                uint64_t seed0, int64_t* __restrict__ counts)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bank)
    return;
  int64_t sites_needed = total < n_particles ? n_particles % total : n_particles;
  double p_sample = (double)sites_needed / (double)total;
  uint64_t seed = mcw_future_seed((uint64_t)(global_start + i), seed0);
  int64_t c = total < n_particles ? n_particles / total : 0;
  if (mcw_prn(seed) < p_sample)
    ++c;
  counts[i] = c;
}

__global__ void k_sample_scatter(const mcw_source_site* __restrict__ sorted, int64_t n_bank,
                                 const int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                 mcw_source_site* __restrict__ out, int64_t cap)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bank)
    return;
  int64_t c = counts[i], o = offsets[i];
  if (c == 0)
    return;
  mcw_source_site s = sorted[i];
  for (int64_t k = 0; k < c; ++k)
    if (o + k < cap)
      out[o + k] = s;
}

// OMC eigenvalue.cpp:195-203: repeat sites from the very end of the fission bank
__global__ void k_pad_from_bank_end(const mcw_source_site* __restrict__ sorted, int64_t n_bank,
                                    mcw_source_site* __restrict__ out, int64_t have, int64_t needed)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= needed)
    return;
  int64_t ib = n_bank - needed + i;
  if (ib < 0)
    ib = 0; // MC OpenMCStudy.C:584
  out[have + i] = sorted[ib];
}

// OMC tallies/tally.cpp:639-670 Tally::accumulate
__global__ void k_tally_accumulate(double* __restrict__ value, double* __restrict__ sum, double* __restrict__ sum_sq,
                                   int64_t n, double norm)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double val = value[i] * norm;
  value[i] = 0.0;
  sum[i] += val;
  sum_sq[i] += val * val;
}

__global__ void k_fill_records(McwParticlesDev p, int64_t n, mcw_history_record* __restrict__ rec)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  mcw_history_record r;
  r.seed_tracking = p.seed_trk[i];
  r.seed_source = p.seed_src[i];
  r.seed_urr = p.seed_urr[i];
  r.E = p.E[i];
  r.k_collision = p.k_col[i];
  r.k_absorption = p.k_abs[i];
  r.k_tracklength = p.k_trk[i];
  r.n_segments = p.n_seg[i];
  r.n_collisions = p.n_coll[i];
  r.n_progeny = p.n_progeny[i];
  r.last_elem = p.elem[i];
  r.fate = p.fate[i];
  r.n_xs = p.n_xs[i];
  r.trace_hash = p.hash[i];
  rec[i] = r;
}

// geometry-only walk (RT lots.i workload)
template <int TYPE>
__global__ void __launch_bounds__(MCW_BLOCK)
k_trace_rays(McwParams P, int64_t n, const double* __restrict__ start, const double* __restrict__ dirs,
             const int32_t* __restrict__ start_elem, const int32_t* __restrict__ start_side,
             const double* __restrict__ max_distance, double* __restrict__ out_distance,
             int32_t* __restrict__ out_segments, uint64_t* __restrict__ out_hash, int32_t* __restrict__ out_status)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  V3 p = v3(start[3 * i], start[3 * i + 1], start[3 * i + 2]);
  V3 d = unit(v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
  int e = start_elem[i];
  if (e < 0)
    e = mcw_locate<TYPE>(P.mesh, p, i);
  int side = start_side ? start_side[i] : MCW_INVALID;
  if (e >= 0 && side == -2)
  {
    ElemRec<TYPE> R(P.mesh, e);
    side = mcw_starting_incoming_side<TYPE>(R, p, d);
  }
  GeneralWalk<TYPE> W;
  W.cur = e;
  W.last = MCW_INVALID;
  W.inc_side = side;
  W.point = p;
  W.dir = d;
  W.ext.invalidate();
  double maxd = max_distance ? max_distance[i] : INFINITY;
  double travelled = 0;
  int segs = 0;
  uint64_t hash = 0xcbf29ce484222325ull;
  int end = -1;
  unsigned long long n_ve = 0, n_moved = 0;
  while (end < 0 && e >= 0 && segs < 10000000)
  {
    StepResult st = W.step(P.mesh, n_ve, n_moved);
    if (st.failed)
    {
      end = 2;
      break;
    }
    ++segs;
    // RT TraceRay.C:1268-1315
    bool at_end = false;
    if (maxd != INFINITY)
    {
      if (fabs(travelled + st.idist - maxd) <= 1e-10 * fmax(1.0, maxd))
        at_end = true;
      else if (travelled + st.idist > maxd)
      {
        st.idist -= travelled + st.idist - maxd;
        st.iside = MCW_INVALID;
        at_end = true;
      }
    }
    travelled += st.idist;
    hash = mcw_hash_step(hash, st.elem, st.iside);
    if (at_end)
    {
      end = 0;
      break;
    }
    int c = W.cross(P.mesh, st);
    if (c == 1)
      end = 1;
    else if (c == 2)
      end = 2;
  }
  out_distance[i] = travelled;
  out_segments[i] = segs;
  if (out_hash)
    out_hash[i] = hash;
  if (out_status)
    out_status[i] = end;
}

__global__ void k_locate_points(McwParams P, int64_t n, const double* __restrict__ xyz, const int64_t* __restrict__ ray_id,
                                int32_t* __restrict__ elem)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  V3 p = v3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  int64_t rid = ray_id ? ray_id[i] : i;
  int e;
  if (P.mesh.type == MCW_TET4)
    e = mcw_locate<MCW_TET4>(P.mesh, p, rid);
  else if (P.mesh.type == MCW_HEX8)
    e = mcw_locate<MCW_HEX8>(P.mesh, p, rid);
  else
    e = mcw_locate<MCW_QUAD4>(P.mesh, p, rid);
  elem[i] = e;
}

// diagnostics: the division the tile lookup inlines (mcw_div_fast) next to the compiler's own a / b
__global__ void k_test_division(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ q_fast,
                                double* __restrict__ q_plain, int32_t* __restrict__ ok)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  bool good;
  q_fast[i] = mcw_div_fast(a[i], b[i], good);
  q_plain[i] = a[i] / b[i];
  ok[i] = good;
}

void mcwk_test_division(int64_t n, const double* a, const double* b, double* q_fast, double* q_plain, int32_t* ok, cudaStream_t st)
{
  if (n > 0)
    k_test_division<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, b, q_fast, q_plain, ok);
}

// ------------------------------------------------------------------ host launchers
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + MCW_BLOCK - 1) / MCW_BLOCK); }

#define MCW_DISPATCH(type, CALL)                                                                               \
  do                                                                                                           \
  {                                                                                                            \
    if ((type) == MCW_TET4)                                                                                    \
    {                                                                                                          \
      constexpr int TY = MCW_TET4;                                                                             \
      CALL;                                                                                                    \
    }                                                                                                          \
    else if ((type) == MCW_HEX8)                                                                               \
    {                                                                                                          \
      constexpr int TY = MCW_HEX8;                                                                             \
      CALL;                                                                                                    \
    }                                                                                                          \
    else                                                                                                       \
    {                                                                                                          \
      constexpr int TY = MCW_QUAD4;                                                                            \
      CALL;                                                                                                    \
    }                                                                                                          \
  } while (0)

void mcwk_init_histories(const McwParams& P, const mcw_source_site* src, int64_t n, int32_t* q_fly, cudaStream_t st)
{
  if (n <= 0)
    return;
  MCW_DISPATCH(P.mesh.type, (k_init_histories<TY><<<grid_for(n), MCW_BLOCK, 0, st>>>(P, src, n, q_fly)));
}

void mcwk_sample_source(const McwParams& P, mcw_source_site* out, int64_t n, const double* d_box, int only_fissionable,
                        cudaStream_t st)
{
  if (n <= 0)
    return;
  MCW_DISPATCH(P.mesh.type, (k_sample_source<TY><<<grid_for(n), MCW_BLOCK, 0, st>>>(P, out, n, d_box, only_fissionable)));
}

void mcwk_xs_lookup_array(const McwLibraryDev& L, int64_t n, const double* E, const int32_t* mat, double* out,
                          cudaStream_t st)
{
  if (n <= 0)
    return;
  k_xs_lookup_array<<<grid_for(n), MCW_BLOCK, 0, st>>>(L, n, E, mat, out);
}

int mcwk_track_blocks_per_sm(int type, bool fused)
{
  int nb = 0;
  if (fused)
  {
    MCW_DISPATCH(type, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_track<TY, true>, MCW_TRACK_BLOCK, 0)));
  }
  else
  {
    MCW_DISPATCH(type, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_track<TY, false>, MCW_TRACK_BLOCK, 0)));
  }
  return nb;
}

void mcwk_track(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper, int max_blocks,
                int32_t* q_fly_next, unsigned long long* cnt_next, int32_t* q_collide, int32_t* q_fallback,
                int drain_lanes, bool fused, int collide_min, cudaStream_t st)
{
  if (upper <= 0)
    return;
  // persistent grid: never more blocks than fit on the device at once, never more than the queue needs
  unsigned g = (unsigned)std::min<int64_t>((upper + MCW_TRACK_BLOCK - 1) / MCW_TRACK_BLOCK, (int64_t)max_blocks);
  if (fused)
  {
    MCW_DISPATCH(P.mesh.type, (k_track<TY, true><<<g, MCW_TRACK_BLOCK, 0, st>>>(P, queue, count, &P.counters->head, q_fly_next,
                                                                               cnt_next, q_collide, q_fallback, 0, collide_min)));
  }
  else
  {
    MCW_DISPATCH(P.mesh.type, (k_track<TY, false><<<g, MCW_TRACK_BLOCK, 0, st>>>(P, queue, count, &P.counters->head, q_fly_next,
                                                                                cnt_next, q_collide, q_fallback, drain_lanes, 32)));
  }
}

int mcwk_pump_blocks_per_sm(int type)
{
  int nb = 0;
  MCW_DISPATCH(type, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pump<TY>, MCW_TRACK_BLOCK, 0)));
  return nb;
}

// Everything that is still alive, from here to the end of the generation, in one launch (mcw_pump.cuh).
void mcwk_pump(const McwParams& P, const McwPump& Q, void* ctl, size_t ctl_bytes, const int32_t* q_fly,
               const unsigned long long* n_fly, const int32_t* q_fallback, const unsigned long long* n_fallback, int blocks,
               int n_track_warps, int patience, cudaStream_t st)
{
  cudaMemsetAsync(ctl, 0, ctl_bytes, st);
  k_pump_fill<<<592, 256, 0, st>>>(Q, P.p.status, q_fly, n_fly, q_fallback, n_fallback);
  MCW_DISPATCH(P.mesh.type, (k_pump<TY><<<blocks, MCW_TRACK_BLOCK, 0, st>>>(P, Q, n_track_warps, patience)));
}

void mcwk_round_reset(McwCounters* c, int consumed, bool fallback_served, cudaStream_t st)
{
  k_round_reset<<<1, 1, 0, st>>>(c, consumed, fallback_served);
}

void mcwk_collide(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper,
                  int32_t* q_fly_next, unsigned long long* cnt_next, uint32_t* keys, int32_t* hist, uint32_t* bins,
                  cudaStream_t st)
{
  if (upper <= 0)
    return;
  MCW_DISPATCH(P.mesh.type,
               (k_collide<TY><<<grid_for(upper), MCW_BLOCK, 0, st>>>(P, queue, count, q_fly_next, cnt_next, keys, hist, bins)));
}

// Cross sections of every material slot (and sample_nuclide's partial sums) at the new energy of the
// survivors of the collision queue.
// binned = false: one thread per (history, slot), direct table reads.
// binned = true : k_collide counted the survivors per energy bucket; scan + scatter place them (the
//                 event loop's particle-queue sort, mcw_xs.cuh), then the tile kernel serves each
//                 256 neighbouring energies from shared-memory windows.
int mcwk_xs_refresh(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper, bool binned,
                    const uint32_t* keys, const int32_t* hist, uint32_t* bins, uint32_t* bin_offsets, int32_t* hist_binned,
                    cudaStream_t st)
{
  if (upper <= 0)
    return 0;
  if (!binned)
  {
    k_xs_all_slots<<<grid_for(upper * P.n_slots), MCW_BLOCK, 0, st>>>(P, queue, count);
    return 1;
  }
  k_bin_scan<<<1, 1024, 0, st>>>(bins, bin_offsets, &P.counters->n_refresh);
  k_bin_scatter<<<(unsigned)((upper + 255) / 256), 256, 0, st>>>(keys, hist, count, bin_offsets, hist_binned);
  k_xs_refresh_tile<<<(unsigned)((upper + MCW_XS_TILE - 1) / MCW_XS_TILE), MCW_XS_TILE, 0, st>>>(P, hist_binned,
                                                                                                 &P.counters->n_refresh);
  return 3;
}

// Stand-alone lookup of (E, material) arrays through the sort + staged kernel; out in caller order.
void mcwk_xs_sort(int64_t n, const double* E, const int32_t* mat, void* temp, size_t temp_bytes, uint32_t* keys,
                  uint32_t* keys_out, int32_t* idx, int32_t* idx_sorted, cudaStream_t st)
{
  if (n <= 0)
    return;
  k_xs_sort_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, E, mat, keys, idx);
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, idx, idx_sorted, n, 0, 24, st);
}

void mcwk_xs_lookup_staged(const McwLibraryDev& L, int64_t n, const double* E, const int32_t* mat, const int32_t* order,
                           double* out, bool plan, cudaStream_t st)
{
  if (n <= 0)
    return;
  const unsigned tiles = (unsigned)((n + MCW_XS_TILE - 1) / MCW_XS_TILE);
  if (plan)
    k_xs_lookup_plan<<<tiles, MCW_XS_TILE, 0, st>>>(L, n, E, mat, order, out);
  else
    k_xs_lookup_staged<<<tiles, MCW_XS_TILE, 0, st>>>(L, n, E, mat, order, out);
}

// queue = the histories whose status equals `want`, in history-index order; *count = how many
struct McwStatusIs
{
  int8_t want;
  __host__ __device__ bool operator()(int8_t s) const { return s == want; }
};
void mcwk_compact_status(void* temp, size_t temp_bytes, const int8_t* status, int64_t n, int want, int32_t* queue,
                         unsigned long long* count, cudaStream_t st)
{
  cub::CountingInputIterator<int32_t> ids(0);
  cub::TransformInputIterator<bool, McwStatusIs, const int8_t*> flags(status, McwStatusIs{(int8_t)want});
  cub::DeviceSelect::Flagged(temp, temp_bytes, ids, flags, queue, count, (int)n, st);
}

size_t mcwk_compact_temp_bytes(int64_t n)
{
  size_t c = 0;
  cub::CountingInputIterator<int32_t> ids(0);
  cub::TransformInputIterator<bool, McwStatusIs, const int8_t*> flags(nullptr, McwStatusIs{0});
  cub::DeviceSelect::Flagged(nullptr, c, ids, flags, (int32_t*)nullptr, (unsigned long long*)nullptr, (int)n);
  return c + 256;
}

int mcwk_refresh_bins(void) { return MCW_REFRESH_BINS; }

size_t mcwk_sort_temp_bytes(int64_t n)
{
  size_t c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, c, (uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, n);
  return c + 256;
}

void mcwk_fallback(const McwParams& P, const int32_t* queue, const unsigned long long* count, int n_blocks,
                   int32_t* q_fly_next, unsigned long long* cnt_next, int32_t* q_collide, cudaStream_t st)
{
  if (n_blocks <= 0)
    return;
  MCW_DISPATCH(P.mesh.type, (k_fallback<TY><<<n_blocks, 64, 0, st>>>(P, queue, count, q_fly_next, cnt_next, q_collide)));
}

void mcwk_trace_rays(const McwParams& P, int64_t n, const double* start, const double* dirs, const int32_t* start_elem,
                     const int32_t* start_side, const double* max_distance, double* out_distance, int32_t* out_segments,
                     uint64_t* out_hash, int32_t* out_status, cudaStream_t st)
{
  if (n <= 0)
    return;
  MCW_DISPATCH(P.mesh.type, (k_trace_rays<TY><<<grid_for(n), MCW_BLOCK, 0, st>>>(P, n, start, dirs, start_elem, start_side,
                                                                                max_distance, out_distance, out_segments,
                                                                                out_hash, out_status)));
}

void mcwk_locate_points(const McwParams& P, int64_t n, const double* xyz, const int64_t* ray_id, int32_t* elem,
                        cudaStream_t st)
{
  if (n <= 0)
    return;
  k_locate_points<<<grid_for(n), MCW_BLOCK, 0, st>>>(P, n, xyz, ray_id, elem);
}

void mcwk_fill_records(const McwParticlesDev& p, int64_t n, mcw_history_record* rec, cudaStream_t st)
{
  if (n <= 0)
    return;
  k_fill_records<<<grid_for(n), MCW_BLOCK, 0, st>>>(p, n, rec);
}

void mcwk_tally_accumulate(double* value, double* sum, double* sum_sq, int64_t n, double norm, cudaStream_t st)
{
  if (n <= 0)
    return;
  k_tally_accumulate<<<grid_for(n), MCW_BLOCK, 0, st>>>(value, sum, sum_sq, n, norm);
}

// ---- CUB-backed utilities (scan / reduce / sort are library calls; they are not on the hot path) ----
size_t mcwk_scan_temp_bytes(int64_t n)
{
  size_t a = 0, b = 0, c = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (int64_t*)nullptr, (int64_t*)nullptr, n);
  cub::DeviceReduce::Sum(nullptr, b, (double*)nullptr, (double*)nullptr, n);
  cub::DeviceRadixSort::SortPairs(nullptr, c, (uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                                  n);
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

void mcwk_exclusive_scan(void* temp, size_t temp_bytes, const int64_t* in, int64_t* out, int64_t n, cudaStream_t st)
{
  if (n > 0)
    cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, n, st);
}

void mcwk_reduce_sum(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st)
{
  cub::DeviceReduce::Sum(temp, temp_bytes, in, out, n, st);
}

void mcwk_bank_sort(void* temp, size_t temp_bytes, const McwParticlesDev& p, int64_t n_particles, int64_t work_offset,
                    const mcw_source_site* bank, int64_t n_bank, int64_t* scratch_counts, int64_t* scratch_scan,
                    mcw_source_site* sorted, cudaStream_t st)
{
  if (n_particles <= 0)
    return;
  k_progeny_to_i64<<<grid_for(n_particles), MCW_BLOCK, 0, st>>>(p.n_progeny, n_particles, scratch_counts);
  cub::DeviceScan::ExclusiveSum(temp, temp_bytes, scratch_counts, scratch_scan, n_particles, st);
  if (n_bank > 0)
    k_bank_scatter<<<grid_for(n_bank), MCW_BLOCK, 0, st>>>(bank, n_bank, scratch_scan, work_offset, sorted);
}

void mcwk_sample_bank(void* temp, size_t temp_bytes, const mcw_source_site* sorted, int64_t n_bank, int64_t global_start,
                      int64_t total, int64_t n_particles, uint64_t seed0, int64_t* counts, int64_t* offsets,
                      mcw_source_site* out, int64_t cap, cudaStream_t st)
{
  if (n_bank <= 0)
    return;
  k_sample_counts<<<grid_for(n_bank), MCW_BLOCK, 0, st>>>(n_bank, global_start, total, n_particles, seed0, counts);
  cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, offsets, n_bank, st);
  k_sample_scatter<<<grid_for(n_bank), MCW_BLOCK, 0, st>>>(sorted, n_bank, counts, offsets, out, cap);
}

void mcwk_pad_from_bank_end(const mcw_source_site* sorted, int64_t n_bank, mcw_source_site* out, int64_t have,
                            int64_t needed, cudaStream_t st)
{
  if (needed > 0)
    k_pad_from_bank_end<<<grid_for(needed), MCW_BLOCK, 0, st>>>(sorted, n_bank, out, have, needed);
}

uint64_t mcwk_host_init_seed(int64_t id, int offset, int64_t master_seed)
{
  // host copy of OMC random_lcg.cpp:58-61 (init_seed) for the bank-sampling stream
  uint64_t n = (uint64_t)id * MCW_PRN_STRIDE, seed = (uint64_t)(master_seed + offset);
  uint64_t g = MCW_PRN_MULT, c = MCW_PRN_ADD, g_new = 1, c_new = 0;
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
} // namespace MCW_NS
