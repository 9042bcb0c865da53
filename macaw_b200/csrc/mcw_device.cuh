// mcw_device.cuh -- device functions of the particle-history path (sm_100a).
//
// Each function states the reference lines it implements (MC = MaCaw src/, RT =
// moose/modules/ray_tracing, OMC = openmc/).  FP64 throughout; compiled with -fmad=false so every
// +,-,*,/ rounds exactly as on the host, and log/sin/cos come from mcw_detmath.h, so a history is
// bit-reproducible against the CPU oracle (tests/test_gpu_parity.py).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "mcw_detmath.h"
#include "mcw_types.h"

#define MCW_D __device__ __forceinline__
#define MCW_PI 3.14159265358979323846264338327950288
#define MCW_TRACE_TOL 1e-8        // RT TraceRayTools.h:45
#define MCW_LOOSE_TRACE_TOL 1e-5  // RT TraceRayTools.h:47
#define MCW_FREE_GAS_THRESHOLD 400.0 // OMC constants.h:42
#define MCW_INVALID (-1)

// ------------------------------------------------------------------ RNG: OMC random_lcg.cpp:12-114
#define MCW_PRN_MULT 6364136223846793005ULL
#define MCW_PRN_ADD 1442695040888963407ULL
#define MCW_PRN_STRIDE 152917ULL

MCW_D double mcw_prn(uint64_t& seed)
{
  seed = MCW_PRN_MULT * seed + MCW_PRN_ADD;
  uint64_t word = ((seed >> ((seed >> 59u) + 5u)) ^ seed) * 12605985483714917081ull;
  uint64_t result = (word >> 43u) ^ word;
  return (double)result * 5.42101086242752217003726400434970855712890625e-20; // ldexp(result, -64)
}

MCW_D uint64_t mcw_future_seed(uint64_t n, uint64_t seed)
{
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

// uint64 -> double -> uint64, the rounding MaCaw's double-typed ray aux data applies to every seed
// at every segment boundary (MC CollisionKernel.C:132-134,157-159,217-219)
MCW_D uint64_t mcw_seed_via_double(uint64_t s)
{
  double d = (double)s; // cvt.rn.f64.u64
  if (d >= 18446744073709551616.0)
    return 0;
  return (uint64_t)d;
}

// ------------------------------------------------------------------ small vector type
struct V3
{
  double x, y, z;
};
MCW_D V3 v3(double x, double y, double z)
{
  V3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
MCW_D V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
MCW_D V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
MCW_D V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
MCW_D V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
MCW_D V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
MCW_D double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MCW_D double norm(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
MCW_D V3 cross(V3 a, V3 b)
{
  return v3(a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x);
}
MCW_D V3 unit(V3 a)
{
  double l = norm(a);
  return v3(a.x / l, a.y / l, a.z / l);
}

// The collision physics calls the software log / sin / cos (mcw_detmath.h, a few hundred instructions each) from many
// places; inlined everywhere they make the collision code several times larger than its own arithmetic, and the tail of
// a generation (one or two warps per scheduler) stalls on instruction fetch.  MCW_COLLISION_MATH_OOL = 1 calls them out
// of line from the collision samplers; the tracking loop and the cross-section lookups keep the inline versions.
#ifndef MCW_COLLISION_MATH_OOL
#define MCW_COLLISION_MATH_OOL 1
#endif
#if MCW_COLLISION_MATH_OOL
__device__ __noinline__ double mcw_log_c(double x) { return mcw_log(x); }
__device__ __noinline__ double mcw_cos_c(double x) { return mcw_cos(x); }
__device__ __noinline__ double2 mcw_sincos_c2(double x)
{
  double s, c;
  mcw_sincos(x, &s, &c);
  return make_double2(s, c);
}
MCW_D void mcw_sincos_c(double x, double* s, double* c)
{
  const double2 r = mcw_sincos_c2(x);
  *s = r.x;
  *c = r.y;
}
#else
#define mcw_log_c mcw_log
#define mcw_cos_c mcw_cos
#define mcw_sincos_c mcw_sincos
#endif

// ------------------------------------------------------------------ samplers
MCW_D double mcw_uniform(double a, double b, uint64_t& s) { return a + (b - a) * mcw_prn(s); }

MCW_D double mcw_maxwell_spectrum(double T, uint64_t& s) // OMC random_dist.cpp:15-27
{
  double r1 = mcw_prn(s), r2 = mcw_prn(s), r3 = mcw_prn(s);
  double c = mcw_cos_c(MCW_PI / 2. * r3);
  return -T * (mcw_log_c(r1) + mcw_log_c(r2) * c * c);
}

MCW_D double mcw_watt_spectrum(double a, double b, uint64_t& s) // OMC random_dist.cpp:29-34
{
  double w = mcw_maxwell_spectrum(a, s);
  return w + 0.25 * a * a * b + mcw_uniform(-1., 1., s) * sqrt(a * a * b * w);
}

MCW_D V3 mcw_rotate_angle(V3 u, double mu, uint64_t& s) // OMC math_functions.cpp:644-672
{
  double phi = 2.0 * MCW_PI * mcw_prn(s);
  double sinphi, cosphi;
  mcw_sincos_c(phi, &sinphi, &cosphi);
  double a = sqrt(fmax(0., 1. - mu * mu));
  double b = sqrt(fmax(0., 1. - u.z * u.z));
  if (b > 1e-10)
  {
    return v3(mu * u.x + a * (u.x * u.z * cosphi - u.y * sinphi) / b,
              mu * u.y + a * (u.y * u.z * cosphi + u.x * sinphi) / b, mu * u.z - a * b * cosphi);
  }
  else
  {
    b = sqrt(1. - u.y * u.y);
    return v3(mu * u.x + a * (-u.x * u.y * sinphi + u.z * cosphi) / b, mu * u.y + a * b * sinphi,
              mu * u.z - a * (u.y * u.z * sinphi + u.x * cosphi) / b);
  }
}

MCW_D V3 mcw_isotropic_direction(uint64_t& s) // OMC distribution_multi.cpp:73-79
{
  double phi = mcw_uniform(0., 2.0 * MCW_PI, s);
  double mu = mcw_uniform(-1., 1., s);
  double sp, cp;
  mcw_sincos_c(phi, &sp, &cp);
  return v3(mu, sqrt(1.0 - mu * mu) * cp, sqrt(1.0 - mu * mu) * sp);
}

// ------------------------------------------------------------------ cross sections
struct Micro
{
  double total, absorption, fission, nu_fission;
  int i_grid;
  double f;
};

MCW_D int mcw_log_bin(const McwLibraryDev& L, double E) // OMC material.cpp:767-768
{
  int i = (int)(mcw_log(E / L.E_min) / L.log_spacing);
  if (i < 0)
    i = 0;
  if (i > L.M - 1)
    i = L.M - 1;
  return i;
}

// OMC nuclide.cpp:557-763 (interpolation branch) + search.h:14-20.  Bounded binary search between
// grid_index[i_log] and grid_index[i_log+1]+1, duplicate-point bump, lin-lin on two adjacent
// 32-byte rows.
MCW_D Micro mcw_micro_xs(const McwLibraryDev& L, int i_nuc, double E, int i_log)
{
  const McwNucParams& n = L.nuc[i_nuc];
  const double* __restrict__ en = L.energy + n.off;
  Micro m;
  int i_grid;
  if (E < en[0])
    i_grid = 0;
  else if (E > en[n.G - 1])
    i_grid = n.G - 2;
  else
  {
    const int32_t* gi = L.grid_index + (int64_t)i_nuc * (L.M + 1) + i_log;
    int i_low = __ldg(gi);
    int i_high = __ldg(gi + 1) + 1;
    // lower_bound_index(&energy[i_low], &energy[i_high], E)
    if (__ldg(en + i_low) == E)
      i_grid = i_low;
    else
    {
      int lo = i_low, hi = i_high; // first position in [lo, hi) with energy >= E
      while (lo < hi)
      {
        int mid = (lo + hi) >> 1;
        if (__ldg(en + mid) < E)
          lo = mid + 1;
        else
          hi = mid;
      }
      i_grid = lo - 1;
    }
  }
  double e0 = __ldg(en + i_grid), e1 = __ldg(en + i_grid + 1);
  if (e0 == e1)
  {
    ++i_grid;
    e0 = e1;
    e1 = __ldg(en + i_grid + 1);
  }
  double f = (E - e0) / (e1 - e0);
  m.i_grid = i_grid;
  m.f = f;
  const double2* row = reinterpret_cast<const double2*>(L.xs4 + 4 * (n.off + i_grid));
  double2 a0 = __ldg(row), a1 = __ldg(row + 1), b0 = __ldg(row + 2), b1 = __ldg(row + 3);
  m.total = (1.0 - f) * a0.x + f * b0.x;
  m.absorption = (1.0 - f) * a0.y + f * b0.y;
  if (n.fissionable)
  {
    m.fission = (1.0 - f) * a1.x + f * b1.x;
    m.nu_fission = (1.0 - f) * a1.y + f * b1.y;
  }
  else
  {
    m.fission = 0.0;
    m.nu_fission = 0.0;
  }
  return m;
}

// ---- thermal scattering cross sections (OMC thermal.cpp:229-243, endf.cpp:153-239, nuclide.cpp:763-791)
// lower_bound_index (OMC search.h:14-20) on x[0 .. n)
MCW_D int mcw_lower_bound_index(const double* __restrict__ x, int n, double v)
{
  if (x[0] == v)
    return 0;
  int lo = 0, hi = n;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (x[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

// Tabulated1D::operator() with one lin-lin region
MCW_D double mcw_tabulated1d(const double* __restrict__ x, const double* __restrict__ y, int n, double v)
{
  if (v < x[0])
    return y[0];
  if (v > x[n - 1])
    return y[n - 1];
  const int i = mcw_lower_bound_index(x, n, v);
  const double r = (v - x[i]) / (x[i + 1] - x[i]);
  return y[i] + r * (y[i + 1] - y[i]);
}

// S(a,b) table of material entry q at energy E, or -1 (OMC material.cpp:778-806: none assigned, or E above the table)
MCW_D int mcw_sab_of_entry(const McwLibraryDev& L, int q, double E)
{
  if (!MCW_RARE || L.n_sab == 0)
    return -1;
  const int t = L.mat_sab[q];
  if (t < 0 || E > L.sab[t].energy_max)
    return -1;
  return t;
}

// ThermalData::calculate_xs
MCW_D void mcw_sab_xs(const McwLibraryDev& L, const McwSabDev& t, double E, double& elastic, double& inelastic)
{
  const double* P = L.sab_pool;
  if (t.el_type == MCW_SAB_ELASTIC_COHERENT)
    elastic = E < P[t.el_x] ? 0.0 : P[t.el_y + mcw_lower_bound_index(P + t.el_x, t.n_el, E)] / E;
  else if (t.el_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
    elastic = mcw_tabulated1d(P + t.el_x, P + t.el_y, t.n_el, E);
  else
    elastic = 0.0;
  inelastic = mcw_tabulated1d(P + t.ixs_e, P + t.ixs, t.n_ixs, E);
}

// Nuclide::calculate_sab_xs on a freshly interpolated micro of material entry q (table i_sab >= 0): corrects
// m.total; returns thermal, thermal_elastic and the corrected micro.elastic.  Kept out of line: only libraries
// with S(a,b) tables come here, and the callers' hot paths should not carry its registers.
__device__ __noinline__ void mcw_sab_apply_ool(const McwLibraryDev* Lg, int q, int i_sab, double E, int i_grid, double f,
                                               double* out4)
{
  const McwLibraryDev& L = *Lg; // the global-memory copy (McwLibraryDev::self_dev)
  const double sab_frac = L.mat_sab_frac[q];
  double el, inel;
  mcw_sab_xs(L, L.sab[i_sab], E, el, inel);
  const double thermal = sab_frac * (el + inel);
  const McwNucParams& n = L.nuc[L.mat_nuc[q]];
  const double* ei = L.el_inel + 2 * (n.off + i_grid);
  const double free_elastic = (1.0 - f) * ei[0] + f * ei[2]; // Nuclide::calculate_elastic_xs (nuclide.cpp:517-529)
  out4[0] = thermal;
  out4[1] = sab_frac * el;                               // thermal_elastic
  out4[2] = sab_frac * free_elastic;                     // what the total loses
  out4[3] = thermal + (1.0 - sab_frac) * free_elastic;   // corrected micro.elastic
}
MCW_D void mcw_sab_apply(const McwLibraryDev& L, int q, int i_sab, double E, Micro& m, double& thermal, double& thermal_elastic,
                         double& elastic)
{
  double o[4];
  mcw_sab_apply_ool(L.self_dev, q, i_sab, E, m.i_grid, m.f, o);
  thermal = o[0];
  thermal_elastic = o[1];
  m.total = m.total + thermal - o[2];
  elastic = o[3];
}

// ---- unresolved-resonance probability tables (OMC nuclide.cpp:753-759, 795-924)
// mcw_function1d: Polynomial (OMC endf.cpp:112-121, Horner over increasing powers) or Tabulated1D with one region,
// histogram or lin-lin, constant outside its range (:156-210)
MCW_D double mcw_fn_eval(const McwLibraryDev& L, const McwFn& f, double v)
{
  const double* __restrict__ p = L.fn_pool + f.off;
  if (f.kind == MCW_FN_POLYNOMIAL)
  {
    double r = 0.0;
    for (int k = f.n - 1; k >= 0; --k)
      r = r * v + __ldg(p + k);
    return r;
  }
  const double* __restrict__ y = p + f.n;
  if (v < __ldg(p))
    return __ldg(y);
  if (v > __ldg(p + f.n - 1))
    return __ldg(y + f.n - 1);
  const int i = mcw_lower_bound_index(p, f.n, v);
  if (f.interp == 1)
    return __ldg(y + i);
  const double x0 = __ldg(p + i), x1 = __ldg(p + i + 1), y0 = __ldg(y + i), y1 = __ldg(y + i + 1);
  const double r = (v - x0) / (x1 - x0);
  return y0 + r * (y1 - y0);
}

// Nuclide::nu(E, total) (OMC nuclide.cpp:510-516): the total-nu function, else the prompt yield; scalar description: the
// polynomial nu_poly
MCW_D double mcw_nu_total(const McwLibraryDev& L, const McwNucParams& n, double E)
{
  if (MCW_RARE && n.fis_first >= 0)
    return mcw_fn_eval(L, n.total_nu.kind != MCW_FN_NONE ? n.total_nu : L.fis[n.fis_first].yield, E);
  double nu = 0.0; // Polynomial (OMC endf.cpp:112-121): Horner, increasing powers
  nu = nu * E + n.nu_poly[1];
  nu = nu * E + n.nu_poly[0];
  return nu;
}

// true when nuclide i_nuc has tables and E lies strictly inside their range
MCW_D bool mcw_urr_in_range(const McwLibraryDev& L, int i_nuc, double E)
{
  if (!MCW_RARE || L.n_urr == 0)
    return false;
  const McwNucParams& n = L.nuc[i_nuc];
  if (n.urr_n == 0)
    return false;
  const double* e = L.urr_pool + n.urr_first;
  return E > e[0] && E < e[n.urr_n - 1];
}

// Nuclide::calculate_urr_xs, lin-lin tables, on the smooth values (i_grid, f, absorption, fission) of nuclide i_nuc;
// seed = the history's URR stream.  out5 = total, absorption, fission, nu_fission, elastic.  Out of line, on the
// global-memory copy of the library (McwLibraryDev::self_dev).
__device__ __noinline__ void mcw_urr_apply_ool(const McwLibraryDev* Lg, int i_nuc, double E, uint64_t seed, int i_grid, double f_grid,
                                               double absorption, double fission_smooth, double* out5)
{
  const McwLibraryDev& L = *Lg;
  const McwNucParams& n = L.nuc[i_nuc];
  const double* en = L.urr_pool + n.urr_first;
  const double* tab = en + n.urr_n;
  const int nb = n.urr_bands;
  int i_energy = 0;
  while (E >= en[i_energy + 1])
    ++i_energy;
  uint64_t fseed = mcw_future_seed((uint64_t)n.identity, seed); // future_prn(index_, ...) (OMC nuclide.cpp:809, random_lcg.cpp:49-53)
  const double r = mcw_prn(fseed);
  const double* t0 = tab + (size_t)i_energy * 6 * nb;
  const double* t1 = t0 + 6 * nb;
  int i_low = 0;
  while (t0[i_low] <= r)
    ++i_low;
  int i_up = 0;
  while (t1[i_up] <= r)
    ++i_up;
  double f = (E - en[i_energy]) / (en[i_energy + 1] - en[i_energy]);
  double elastic = (1. - f) * t0[2 * nb + i_low] + f * t1[2 * nb + i_up];
  double fission = (1. - f) * t0[3 * nb + i_low] + f * t1[3 * nb + i_up];
  double capture = (1. - f) * t0[4 * nb + i_low] + f * t1[4 * nb + i_up];
  const double* ei = L.el_inel + 2 * (n.off + i_grid); // (elastic, inelastic) per grid point
  double inelastic = 0.;
  if (n.urr_inel_flag != -1 && i_grid - n.inel_thr >= 0)
    inelastic = (1. - f_grid) * ei[1] + f_grid * ei[3];
  if (n.urr_multiply_smooth)
  {
    const double free_elastic = (1.0 - f_grid) * ei[0] + f_grid * ei[2];
    elastic *= free_elastic;
    capture *= (absorption - fission_smooth);
    fission *= fission_smooth;
  }
  if (elastic < 0.)
    elastic = 0.;
  if (fission < 0.)
    fission = 0.;
  if (capture < 0.)
    capture = 0.;
  out5[0] = elastic + inelastic + capture + fission;
  out5[1] = capture + fission;
  out5[2] = fission;
  out5[3] = mcw_nu_total(L, n, E) * fission;
  out5[4] = elastic;
}
MCW_D void mcw_urr_apply(const McwLibraryDev& L, int i_nuc, double E, uint64_t seed, Micro& m, double& elastic)
{
  double o[5];
  mcw_urr_apply_ool(L.self_dev, i_nuc, E, seed, m.i_grid, m.f, m.absorption, m.fission, o);
  m.total = o[0];
  m.absorption = o[1];
  m.fission = o[2];
  if (L.nuc[i_nuc].fissionable)
    m.nu_fission = o[3];
  elastic = o[4];
}

// micro cross sections of material entry q: its nuclide, with the S(a,b) correction of the total where a table applies
// and, when the history's URR stream seed is given, the probability tables of a nuclide in its unresolved range
MCW_D Micro mcw_micro_xs_entry(const McwLibraryDev& L, int q, double E, int i_log, const uint64_t* seed_urr = nullptr)
{
  const int i_nuc = L.mat_nuc[q];
  Micro m = mcw_micro_xs(L, i_nuc, E, i_log);
  const int i_sab = mcw_sab_of_entry(L, q, E);
  if (i_sab >= 0)
  {
    double th, the, el;
    mcw_sab_apply(L, q, i_sab, E, m, th, the, el);
  }
  if (seed_urr && mcw_urr_in_range(L, i_nuc, E))
  {
    double el;
    mcw_urr_apply(L, i_nuc, E, *seed_urr, m, el);
  }
  return m;
}

struct Macro
{
  double total, absorption, fission, nu_fission;
};

// OMC material.cpp:748-831.  prefix (may be null) receives the running total after every nuclide: the
// partial sums sample_nuclide (OMC physics.cpp:442-465) forms at the next collision at this energy.
MCW_D Macro mcw_macro_xs(const McwLibraryDev& L, int i_mat, double E, double* __restrict__ prefix = nullptr,
                         const uint64_t* seed_urr = nullptr)
{
  Macro M;
  M.total = M.absorption = M.fission = M.nu_fission = 0.0;
  int i_log = mcw_log_bin(L, E);
  int b = L.mat_off[i_mat], e = L.mat_off[i_mat + 1];
  for (int i = b; i < e; ++i)
  {
    Micro m = mcw_micro_xs_entry(L, i, E, i_log, seed_urr);
    double d = L.mat_dens[i];
    M.total += d * m.total;
    M.absorption += d * m.absorption;
    M.fission += d * m.fission;
    M.nu_fission += d * m.nu_fission;
    if (prefix)
      prefix[i - b] = M.total;
  }
  return M;
}

// Macroscopic cross sections of every material slot at energy E, written to the history's sectors
// of McwParticlesDev::xs.  Called by the kernels that change a history's energy, so the tracking
// kernel never has to stop for a lookup when it enters another material.
MCW_D void mcw_lookup_all_slots(const McwParams& P, int64_t i, double E)
{
  double* xs_of_history = P.p.xs + (size_t)i * P.n_slots * 4;
  double* prefix = P.p.xs_prefix + (size_t)i * P.n_plan_pad;
  uint64_t seed_urr = 0;
  const bool urr = MCW_RARE && P.lib.n_urr != 0;
  if (urr)
    seed_urr = P.p.seed_urr[i]; // the history's URR stream picks the probability-table band
  for (int s = 0; s < P.n_slots; ++s)
  {
    const Macro M = mcw_macro_xs(P.lib, P.slot_mat[s], E, prefix + P.slot_plan_off[s], urr ? &seed_urr : nullptr);
    double2* o = reinterpret_cast<double2*>(xs_of_history + 4 * s);
    o[0] = make_double2(M.total, M.absorption);
    o[1] = make_double2(M.fission, M.nu_fission);
  }
}

MCW_D Macro mcw_load_xs(const double* __restrict__ xs, int n_slots, int64_t i, int slot)
{
  const double2* c = reinterpret_cast<const double2*>(xs + ((size_t)i * n_slots + slot) * 4);
  const double2 a = c[0], b = c[1];
  Macro M;
  M.total = a.x;
  M.absorption = a.y;
  M.fission = b.x;
  M.nu_fission = b.y;
  return M;
}

// ------------------------------------------------------------------ element records
template <int TYPE>
struct ElemRec
{
  using T = McwElem<TYPE>;
  const unsigned char* base;
  MCW_D ElemRec(const McwMeshDev& m, int64_t e) : base(m.rec + (size_t)e * T::STRIDE) {}
  MCW_D V3 P(int v) const
  {
    const double* c = reinterpret_cast<const double*>(base) + v * T::CPV;
    if (T::CPV == 3)
      return v3(__ldg(c), __ldg(c + 1), __ldg(c + 2));
    return v3(__ldg(c), __ldg(c + 1), 0.0);
  }
  MCW_D int nbr_code(int s) const
  {
    return __ldg(reinterpret_cast<const int32_t*>(base + T::NV * T::CPV * 8) + s);
  }
  // neighbour element id, or -(boundary id)-1 on the domain boundary
  MCW_D int nbr(int s) const
  {
    const int c = nbr_code(s);
    return c < 0 ? c : (c >> MCW_NBR_SHIFT);
  }
  MCW_D double hmax() const
  {
    return __ldg(reinterpret_cast<const double*>(base + T::NV * T::CPV * 8 + T::NS * 4));
  }
  MCW_D int slot() const // cross-section slot of the element's material (McwParams::slot_mat)
  {
    return __ldg(reinterpret_cast<const int32_t*>(base + T::NV * T::CPV * 8 + T::NS * 4 + 8));
  }
  MCW_D int block() const
  {
    return __ldg(reinterpret_cast<const int32_t*>(base + T::NV * T::CPV * 8 + T::NS * 4 + 12));
  }
};

template <int TYPE>
MCW_D int side_node(int s, int i);
template <>
MCW_D int side_node<MCW_TET4>(int s, int i)
{
  // {0,2,1},{0,1,3},{1,2,3},{2,0,3}
  const unsigned tbl = 0x0u | (2u << 2) | (1u << 4) | (0u << 6) | (1u << 8) | (3u << 10) | (1u << 12) |
                       (2u << 14) | (3u << 16) | (2u << 18) | (0u << 20) | (3u << 22);
  return (tbl >> (2 * (3 * s + i))) & 3u;
}
template <>
MCW_D int side_node<MCW_HEX8>(int s, int i)
{
  // {0,3,2,1},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}; 3 bits each, 24 entries = 72 bits
  const unsigned long long lo = 0ull | (3ull << 3) | (2ull << 6) | (1ull << 9) | (0ull << 12) | (1ull << 15) |
                                (5ull << 18) | (4ull << 21) | (1ull << 24) | (2ull << 27) | (6ull << 30) |
                                (5ull << 33) | (2ull << 36) | (3ull << 39) | (7ull << 42) | (6ull << 45);
  const unsigned hi = 3u | (0u << 3) | (4u << 6) | (7u << 9) | (4u << 12) | (5u << 15) | (6u << 18) | (7u << 21);
  int k = 4 * s + i;
  return k < 16 ? (int)((lo >> (3 * k)) & 7ull) : (int)((hi >> (3 * (k - 16))) & 7u);
}
template <>
MCW_D int side_node<MCW_QUAD4>(int s, int i)
{
  return (s + i) & 3; // {0,1},{1,2},{2,3},{3,0}
}

struct Extrema // RT include/raytracing/ElemExtrema.h
{
  int first, second;
  MCW_D Extrema() : first(MCW_INVALID), second(MCW_INVALID) {}
  MCW_D bool atVertex() const { return first != MCW_INVALID && second == MCW_INVALID; }
  MCW_D bool atEdge() const { return first != MCW_INVALID && second != MCW_INVALID; }
  MCW_D bool atExtrema() const { return first != MCW_INVALID; }
  MCW_D void invalidate() { first = second = MCW_INVALID; }
  MCW_D void setVertex(int v)
  {
    first = v;
    second = MCW_INVALID;
  }
  MCW_D void setEdge(int a, int b)
  {
    first = a;
    second = b;
  }
  MCW_D bool atEdgeOf(int a, int b) const
  {
    return atEdge() && ((first == a && second == b) || (first == b && second == a));
  }
};

// RT TraceRayTools.C:371-478
MCW_D bool mcw_intersect_triangle(V3 start, V3 dir, V3 p0, V3 p1, V3 p2, int v0, int v1, int v2,
                                  double& dist, Extrema& ext, double hmax)
{
  const double inv_hmax = 1.0 / hmax;
  V3 edge1 = (p1 - p0) * inv_hmax;
  V3 edge2 = (p2 - p0) * inv_hmax;
  V3 pvec = cross(dir, edge2);
  double det = dot(edge1, pvec);
  if (det < MCW_TRACE_TOL)
    return false;
  V3 tvec = (start - p0) * inv_hmax;
  double u = dot(tvec, pvec);
  if (u < -MCW_TRACE_TOL || u > det + MCW_TRACE_TOL)
    return false;
  V3 qvec = cross(tvec, edge1);
  double v = dot(dir, qvec);
  if (v < -MCW_TRACE_TOL || u + v > det + MCW_TRACE_TOL)
    return false;
  double possible = dot(edge2, qvec) / det;
  if (possible <= MCW_TRACE_TOL)
    return false;
  dist = possible * hmax;
  if (u < MCW_TRACE_TOL * det)
  {
    if (v < MCW_TRACE_TOL * det)
      ext.setVertex(v0);
    else if (v > (1.0 - MCW_TRACE_TOL) * det)
      ext.setVertex(v2);
    else
      ext.setEdge(v0, v2);
  }
  else if (v < MCW_TRACE_TOL * det)
  {
    if (u > (1.0 - MCW_TRACE_TOL) * det)
      ext.setVertex(v1);
    else
      ext.setEdge(v0, v1);
  }
  else if (u + v > (1.0 - MCW_TRACE_TOL) * det)
    ext.setEdge(v1, v2);
  return true;
}

// RT TraceRayTools.C:60-128
MCW_D bool mcw_line_line_2d(V3 start, V3 dir, double length, V3 v0, V3 v1, V3& ipoint, double& idist,
                            int& seg_v)
{
  V3 r = dir * length;
  V3 s = v1 - v0;
  double rxs = r.x * s.y - r.y * s.x;
  if (fabs(rxs) < MCW_TRACE_TOL)
    return false;
  V3 v0mu0 = v0 - start;
  double t = (v0mu0.x * s.y - v0mu0.y * s.x) / rxs;
  if (0 >= t + MCW_TRACE_TOL || t - MCW_TRACE_TOL > 1.0)
    return false;
  double u = (v0mu0.x * r.y - v0mu0.y * r.x) / rxs;
  if (0 < u + MCW_TRACE_TOL && u - MCW_TRACE_TOL <= 1.0)
  {
    ipoint = start + r * t;
    idist = t * length;
    if (u < MCW_TRACE_TOL)
      seg_v = 0;
    else if (u > 1.0 - MCW_TRACE_TOL)
      seg_v = 1;
    return true;
  }
  return false;
}

// RT TraceRayTools.h:558-856 sideIntersectedByLine<T>
template <int TYPE>
MCW_D bool mcw_side_hit(const McwMeshDev& mesh, const ElemRec<TYPE>& R, V3 start, V3 dir, int s,
                        V3& ipoint, double& idist, Extrema& ext, double hmax)
{
  if (TYPE == MCW_QUAD4)
  {
    int seg_v = -1;
    int a = side_node<TYPE>(s, 0), b = side_node<TYPE>(s, 1);
    bool hit = mcw_line_line_2d(start, dir, mesh.domain_max_length, R.P(a), R.P(b), ipoint, idist, seg_v);
    if (seg_v >= 0)
      ext.setVertex(seg_v == 0 ? a : b);
    return hit;
  }
  bool hit;
  if (TYPE == MCW_TET4)
  {
    int a = side_node<TYPE>(s, 2), b = side_node<TYPE>(s, 1), c = side_node<TYPE>(s, 0);
    hit = mcw_intersect_triangle(start, dir, R.P(a), R.P(b), R.P(c), a, b, c, idist, ext, hmax);
  }
  else
  {
    // RT TraceRayTools.C:480-553 intersectQuad(v00, v10, v11, v01)
    int v00 = side_node<TYPE>(s, 3), v10 = side_node<TYPE>(s, 2), v11 = side_node<TYPE>(s, 1),
        v01 = side_node<TYPE>(s, 0);
    V3 p00 = R.P(v00), p11 = R.P(v11);
    hit = mcw_intersect_triangle(start, dir, p00, R.P(v10), p11, v00, v10, v11, idist, ext, hmax);
    if (!hit)
      hit = mcw_intersect_triangle(start, dir, p11, R.P(v01), p00, v11, v01, v00, idist, ext, hmax);
    if (hit && ext.atEdgeOf(v00, v11))
      ext.invalidate();
  }
  if (hit)
    ipoint = start + idist * dir;
  return hit;
}

// RT TraceRay.C:292-493 exitsElem (no backface culling, planar sides)
template <int TYPE>
MCW_D bool mcw_exits_elem(const McwMeshDev& mesh, const ElemRec<TYPE>& R, int incoming_side, V3 start,
                          V3 dir, V3& ipoint, int& iside, Extrema& iext, double& idist)
{
  const double hmax = R.hmax();
  double best = MCW_TRACE_TOL * hmax;
  iside = MCW_INVALID;
#pragma unroll 1
  for (int s = 0; s < McwElem<TYPE>::NS; ++s)
  {
    if (s == incoming_side)
      continue;
    V3 cp = v3(0, 0, 0);
    double cd = 0;
    Extrema ce;
    if (mcw_side_hit<TYPE>(mesh, R, start, dir, s, cp, cd, ce, hmax) && cd > best)
    {
      iside = s;
      idist = cd;
      ipoint = cp;
      iext = ce;
      best = cd;
    }
  }
  return iside != MCW_INVALID;
}

// Outward unit normal at the side centroid (libMesh FE face normal; oracle Mesh::sideNormal)
template <int TYPE>
MCW_D V3 mcw_side_normal(const ElemRec<TYPE>& R, int s)
{
  if (TYPE == MCW_QUAD4)
  {
    V3 a = R.P(side_node<TYPE>(s, 0)), b = R.P(side_node<TYPE>(s, 1));
    return unit(v3(b.y - a.y, -(b.x - a.x), 0.0));
  }
  if (TYPE == MCW_TET4)
  {
    V3 a = R.P(side_node<TYPE>(s, 0)), b = R.P(side_node<TYPE>(s, 1)), c = R.P(side_node<TYPE>(s, 2));
    return unit(cross(b - a, c - a));
  }
  V3 a = R.P(side_node<TYPE>(s, 0)), b = R.P(side_node<TYPE>(s, 1)), c = R.P(side_node<TYPE>(s, 2)),
     d = R.P(side_node<TYPE>(s, 3));
  V3 dxi = (b + c) - (a + d), deta = (c + d) - (a + b);
  return unit(cross(dxi, deta));
}

template <int TYPE>
MCW_D V3 mcw_side_centroid(const ElemRec<TYPE>& R, int s)
{
  V3 c = v3(0, 0, 0);
#pragma unroll
  for (int i = 0; i < McwElem<TYPE>::NPS; ++i)
    c = c + R.P(side_node<TYPE>(s, i));
  return c / (double)McwElem<TYPE>::NPS;
}

template <int TYPE>
MCW_D bool mcw_contains_point(const ElemRec<TYPE>& R, V3 p)
{
  double tol = 1e-6 * R.hmax();
  for (int s = 0; s < McwElem<TYPE>::NS; ++s)
  {
    V3 n = mcw_side_normal<TYPE>(R, s);
    V3 d = p - mcw_side_centroid<TYPE>(R, s);
    if (McwElem<TYPE>::DIM == 2)
      d.z = 0;
    if (dot(d, n) > tol)
      return false;
  }
  return true;
}

template <int TYPE>
MCW_D bool mcw_side_contains_point(const ElemRec<TYPE>& R, int s, V3 p)
{
  double tol = 1e-6 * R.hmax();
  V3 d = p - mcw_side_centroid<TYPE>(R, s);
  if (McwElem<TYPE>::DIM == 2)
    d.z = 0;
  return fabs(dot(d, mcw_side_normal<TYPE>(R, s))) <= tol && mcw_contains_point<TYPE>(R, p);
}

// Point locator + ClaimRays tie-break (RT ClaimRays.C:104-139)
template <int TYPE>
MCW_D int mcw_locate(const McwMeshDev& m, V3 p, int64_t ray_id)
{
  double ex = m.bb_max[0] - m.bb_min[0], ey = m.bb_max[1] - m.bb_min[1], ez = m.bb_max[2] - m.bb_min[2];
  int i = 0, j = 0, k = 0;
  if (ex > 0)
  {
    i = (int)floor((p.x - m.bb_min[0]) / ex * m.gx);
    i = min(max(i, 0), m.gx - 1);
  }
  if (ey > 0)
  {
    j = (int)floor((p.y - m.bb_min[1]) / ey * m.gy);
    j = min(max(j, 0), m.gy - 1);
  }
  if (McwElem<TYPE>::DIM == 3 && ez > 0)
  {
    k = (int)floor((p.z - m.bb_min[2]) / ez * m.gz);
    k = min(max(k, 0), m.gz - 1);
  }
  size_t c = ((size_t)k * m.gy + j) * m.gx + i;
  const bool smallest = (ray_id % 2) == 0;
  int best = -1;
  for (int64_t q = m.cell_off[c]; q < m.cell_off[c + 1]; ++q)
  {
    int e = m.cell_elems[q];
    ElemRec<TYPE> R(m, e);
    if (!mcw_contains_point<TYPE>(R, p))
      continue;
    if (best < 0 || (smallest && e < best) || (!smallest && e > best))
      best = e;
  }
  return best;
}

template <int TYPE>
MCW_D int mcw_starting_incoming_side(const ElemRec<TYPE>& R, V3 p, V3 dir) // RT ClaimRays.C:141-163
{
  for (int s = 0; s < McwElem<TYPE>::NS; ++s)
    if (mcw_side_contains_point<TYPE>(R, s, p) && dot(mcw_side_normal<TYPE>(R, s), dir) < MCW_TRACE_TOL)
      return s;
  return MCW_INVALID;
}

MCW_D uint64_t mcw_hash_step(uint64_t h, int64_t elem, int side)
{
  uint64_t key = ((uint64_t)elem << 4) | (uint64_t)(side & 15);
  return (h ^ key) * 0x100000001b3ull;
}

// ------------------------------------------------------------------ warp-aggregated helpers
MCW_D unsigned long long mcw_warp_append(unsigned long long* counter, bool want)
{
  // one atomicAdd per warp: lanes that want a slot get consecutive indices
  unsigned mask = __ballot_sync(__activemask(), want);
  if (!want)
    return 0;
  int leader = __ffs(mask) - 1;
  int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == leader)
    base = atomicAdd(counter, (unsigned long long)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1));
}
