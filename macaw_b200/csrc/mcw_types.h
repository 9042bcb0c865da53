// mcw_types.h -- device data layout of libmacaw_b200 (shared by kernels and the host API).
//
// HBM layout:
//   * Library: all nuclides concatenated.  energy[sumG] (f64), xs4[sumG][4] (total, absorption,
//     fission, nu-fission: one 32-byte sector per grid point, two adjacent rows = one 64-byte read),
//     el_inel[sumG][2] (elastic, level-inelastic; collision kernel only), grid_index[n_nuc][M+1] (i32).
//   * Mesh: one fixed-stride record per element (128 B for TET4/QUAD4 = one L2 line, 256 B for HEX8)
//     holding the element's OWN copy of its vertex coordinates, one 32-bit code per side
//     (interior: neighbour id << 3 | the neighbour's side that faces back, so the incoming side of
//     the next element needs no second record; boundary: -(bnd_id)-1), the block's hmax, its
//     cross-section slot (dense index over the materials the mesh uses) and block index.  One
//     dependent load per element crossing instead of connectivity -> node-coordinate indirection.
//   * Particles: structure of arrays indexed by history slot.
#pragma once
#include <stdint.h>

#include "../../include/macaw_b200.h"

#define MCW_MAX_NUC_PER_MAT 64
#define MCW_MAX_XS_SLOTS 32   // materials actually present in the mesh
#define MCW_NBR_SHIFT 3       // side code = neighbour << 3 | back side
#define MCW_MAX_ELEMS (1ll << 28)

// mcw_function1d on the device: values in McwLibraryDev::fn_pool at `off` -- tabulated: x[n] then y[n]; polynomial: y[n]
struct McwFn
{
  int32_t kind, n, interp, off;
};

// mcw_fission_product on the device; the continuous tabular law's table sits in the tab_* pools like every other
struct McwFisProduct
{
  McwFn yield, a, b;
  double u;
  int32_t law, en_first, en_n, en_hist;
};

struct McwNucParams
{
  double awr, kT;
  double inel_E, inel_ratio;
  double nu_poly[2];
  double dyield[MCW_MAX_DELAYED_GROUPS];
  double dtheta[MCW_MAX_DELAYED_GROUPS];
  double watt_a, watt_b, q;
  int64_t off;      // offset of this nuclide in the concatenated grids
  int32_t G;
  int32_t fissionable;
  int32_t n_inel, inel_thr;
  int32_t n_precursor;
  int32_t inel_law;            // MCW_LAW_*
  // tabulated secondary distributions: first incoming energy in McwLibraryDev::tab_* and how many (0 = absent)
  int32_t el_ang_first, el_ang_n;
  int32_t inel_ang_first, inel_ang_n;
  int32_t inel_en_first, inel_en_n;
  int32_t inel_en_hist, inel_in_cm; // ContinuousTabular: histogram interpolation in incoming energy; scatter_in_cm_
  // unresolved-resonance probability tables (mcw_urr_desc): energies at McwLibraryDev::urr_pool[urr_first ..], then the
  // table [n_energy][6][n_band]; urr_n = 0: none
  int32_t urr_first, urr_n, urr_bands, urr_inel_flag, urr_multiply_smooth;
  int32_t identity; // the nuclide this data set belongs to (its index among the nuclides, OMC Nuclide::index_): what the URR
                    // band draw and the nuclide tally bins go by when a nuclide has data sets at several temperatures
  // product distributions that give energy and cosine together (inel_law >= MCW_LAW_CORRELATED): law_first is an entry
  // of the tab_* pools -- correlated: the first of one angular distribution per point of the outgoing-energy table
  // (point j of the table -> entry law_first + j); Kalbach-Mann: ONE entry whose tab_x / tab_p hold r / a per point
  int32_t law_first, nbody_n;
  double nbody_mass_ratio, nbody_q;
  // the nuclide's scattering reactions other than elastic, in OMC index_inelastic_scatter_ order: McwLibraryDev::rx
  // [rx_first, rx_first + 1 + n_rx_more); the first is the inline reaction above again (the common kernel build reads the
  // inline fields only: level law, yield 1), the others exist only in problems that run on the full build
  int32_t rx_first, n_rx_more;
  // fission neutrons as a nuclide file has them: McwLibraryDev::fis[fis_first + 0] prompt, [+ g] delayed group g, and the
  // total nu (kind 0: the prompt yield); fis_first < 0: the scalar description above.  Full kernel build only.
  int32_t fis_first, fis_pad;
  McwFn total_nu;
};

// One scattering reaction (mcw_reaction_desc, or the nuclide's inline one): what inelastic_scatter needs
struct McwRxParams
{
  double level_E, level_ratio; // MCW_LAW_LEVEL
  double yield;                // constant multiplicity of the neutron product
  double nbody_mass_ratio, nbody_q;
  int64_t xs_off;              // McwLibraryDev::rx_xs[xs_off + (i_grid - thr)]; -1 for the inline reaction (el_inel holds it)
  int32_t thr, law, in_cm;
  int32_t ang_first, ang_n;    // tab_* entries as in McwNucParams
  int32_t en_first, en_n, en_hist;
  int32_t law_first, nbody_n;
  int32_t pad[2];
  McwFn yield_fn;              // energy-dependent multiplicity (kind 0: the constant `yield`)
  McwFn a, b;                  // MCW_LAW_MAXWELL / _EVAPORATION / _WATT: theta(E) or a(E), b(E) ...
  double u;                    // ... and the restriction energy
};

// One temperature of an S(alpha, beta) table (mcw_thermal_desc): offsets into McwLibraryDev::sab_pool (doubles) /
// sab_ipool (the continuous law's distribution offsets)
struct McwSabDev
{
  double energy_max;
  int32_t inel_type, el_type, n_ixs, n_ien, n_mu, n_e_out, skewed, n_el, n_el_mu, ioff;
  int32_t ixs_e, ixs, ien, eout, pdf, cdf, mu, el_x, el_y, el_mu;
};

struct McwLibraryDev
{
  const double* energy;      // [sumG]
  const double* xs4;         // [sumG][4]
  const double* el_inel;     // [sumG][2]
  const int32_t* grid_index; // [n_nuc][M+1]
  const McwNucParams* nuc;   // [n_nuc]
  const int32_t* mat_off;    // [n_mat+1]
  const int32_t* mat_nuc;    // concatenated nuclide indices
  const double* mat_dens;
  const int32_t* mat_iso;    // isotropic-in-lab flag per material entry (OMC material.cpp p0_)
  const double* mat_e0;      // first / last grid energy of the entry's nuclide (staged lookup)
  const double* mat_e1;
  const int32_t* mat_fissionable;
  // pools of the tabulated secondary distributions of all nuclides (mcw_tabular2d): per incoming energy its
  // value, the ABSOLUTE first / one-past-last point in tab_x / tab_p / tab_c, the interpolation and the number
  // of leading discrete lines
  const double* tab_energy;
  const int32_t* tab_off; // [n_tab_energy + 1 per table, tables back to back: entry e -> points [tab_off[e], tab_end[e])]
  const int32_t* tab_end;
  const int32_t* tab_interp;
  const int32_t* tab_ndisc;
  const double *tab_x, *tab_p, *tab_c;
  const McwFisProduct* fis; // fission-neutron products (McwNucParams::fis_first)
  const double* fn_pool;    // values of every McwFn
  const McwRxParams* rx; // scattering reactions of all nuclides (McwNucParams::rx_first)
  const double* rx_xs;   // cross sections of the reactions after each nuclide's first, from their thresholds on
  // thermal scattering: tables, their pools, and per material entry the table (-1 = none) and the bound fraction
  const McwSabDev* sab;
  const double* sab_pool;
  const int32_t* sab_ipool;
  const int32_t* mat_sab;
  const double* mat_sab_frac;
  int32_t n_sab, n_urr; // number of S(a,b) tables in use / of nuclides with probability tables (0: branches off)
  const double* urr_pool;
  // This struct again, in global memory.  The kernels get McwParams by value (parameter space); an out-of-line device
  // function that took `lib` by reference would force a private copy of the whole parameter block into every thread's
  // local memory, so the rare out-of-line paths (S(a,b), nuclide tally bins) are handed this pointer instead.
  const McwLibraryDev* self_dev;
  int32_t n_nuc, n_mat, M;
  int32_t n_identities; // nuclides (OMC data::nuclides.size(): the URR stream advance); n_nuc counts every data set
  double E_min, E_max, log_spacing;
};

// element record strides (bytes) and field offsets
template <int TYPE>
struct McwElem;
template <>
struct McwElem<MCW_TET4>
{
  static constexpr int NV = 4, NS = 4, NPS = 3, DIM = 3, NE = 6, CPV = 3, STRIDE = 128;
};
template <>
struct McwElem<MCW_HEX8>
{
  static constexpr int NV = 8, NS = 6, NPS = 4, DIM = 3, NE = 12, CPV = 3, STRIDE = 256;
};
template <>
struct McwElem<MCW_QUAD4>
{
  static constexpr int NV = 4, NS = 4, NPS = 2, DIM = 2, NE = 0, CPV = 2, STRIDE = 128;
};

struct McwMeshDev
{
  const unsigned char* rec; // element records
  const int64_t* conn;      // [n_elem][NV] global node ids (rare path only)
  const int64_t* n2e_off;   // node -> elements CSR (rare path only)
  const int32_t* n2e;
  const int32_t* bnd_kind;  // [n_bnd]
  const int64_t* cell_off;  // point-location grid
  const int32_t* cell_elems;
  int32_t gx, gy, gz;
  int32_t type, n_bnd, rectangular;
  int64_t n_elem, n_nodes;
  double bb_min[3], bb_max[3];
  double domain_max_length;
};

// particle status between kernels
enum
{
  MCW_ST_FLY = 0,      // needs (xs +) tracking
  MCW_ST_COLLIDE = 1,  // stopped at a collision point, waiting for the collision kernel
  MCW_ST_FALLBACK = 2, // next step needs the general (vertex/edge/neighbour) walk
  MCW_ST_DEAD = 3
};

struct McwParticlesDev
{
  double *x, *y, *z, *ux, *uy, *uz, *E, *wgt;
  uint64_t *seed_trk, *seed_src, *seed_urr;
  double *k_col, *k_abs, *k_trk;
  int32_t *elem, *last_elem;
  int32_t *n_progeny, *n_seg, *n_coll, *n_xs;
  int8_t *inc_side, *ext_first, *ext_second, *status, *fate;
  uint64_t* hash;
  // macroscopic cross sections at the history's current energy for EVERY material slot, refreshed
  // by the kernel that changes the energy (source initialisation, collision):
  // xs[(history * n_slots + slot) * 4 + {total, absorption, fission, nu-fission}] -- one 32-byte
  // sector per (history, slot), so a material change in the tracking kernel is one sector read
  double* xs;
  // Running sums of N_i * sigma_t,i over the refresh plan (McwParams::plan_*), restarting at every slot:
  // xs_prefix[history * n_plan_pad + entry].  They are the partial sums `prob` of sample_nuclide
  // (OMC physics.cpp:442-465) -- the same additions, in the same order, that form the material's total in
  // Material::calculate_xs (OMC material.cpp:790-829) -- so the collision kernel picks the nuclide by scanning
  // them instead of looking every nuclide of the material up a second time.
  double* xs_prefix;
  // settings.record_traces == 2 only: exit point of the segment a collision interrupted (the point the reference
  // caches when the history is absorbed there, RT TraceRay caches the untruncated segment end)
  double *tx, *ty, *tz;
};

struct McwTallyDev
{
  double* value; // [bins][n_nuc_bins][n_scores]
  const double* ebins;
  const int32_t* nuclides; // nuclide bins (library index, -1 = material total); null = the single bin "total"
  int32_t n_nuc_bins;
  int32_t n_edges, n_energy, n_scores, granularity, estimator;
  int32_t scores[MCW_N_SCORES];
};

#define MCW_MAX_TALLIES 8

struct McwCounters // device-side counters, one cache line each where contended
{
  unsigned long long n_fly[2], pad0[14]; // ping-pong counts of the two fly queues
  unsigned long long n_collide, pad1[15];
  unsigned long long n_fallback, pad2[15];
  unsigned long long n_bank, pad3[15];
  unsigned long long head, pad4[15]; // next unclaimed entry of the queue the tracking kernel is draining
  unsigned long long n_segments, n_collisions, n_xs, n_nuc_lookups, n_vertex_edge, n_moved, n_failed,
      n_bank_dropped;
  unsigned long long n_refresh, n_trace, pad5[6]; // survivors of the round's collisions = entries of the energy-binned refresh queue
};

// Device-side queues of the persistent generation kernel (mcw_pump.cuh): three rings of history indices
// (-1 = empty slot) with ticket counters one cache line apart, and the count of histories still alive.
struct McwPump
{
  int32_t *q_fly, *q_collide, *q_fallback;
  unsigned long long *fly_head, *fly_tail, *collide_head, *collide_tail, *fallback_head, *fallback_tail;
  long long* alive;
  uint32_t mask; // ring capacity - 1 (power of two, > 2 x histories + every lane of the device)
};

struct McwParams
{
  McwLibraryDev lib;
  McwMeshDev mesh;
  McwParticlesDev p;
  McwTallyDev tally[MCW_MAX_TALLIES];
  const McwTallyDev* tally_dev; // the same array in global memory, for out-of-line scoring paths (see McwLibraryDev::self_dev)
  int32_t n_tallies, tallies_active;
  int32_t any_tracklength, any_collision_tally;
  // 1: the collide / fly queues are rebuilt every round by compacting the status array (history-index
  // order, so neighbouring threads touch neighbouring state); kernels then skip their queue appends
  int32_t ordered_queues;
  mcw_source_site* bank;
  int64_t bank_capacity;
  McwCounters* counters;
  int64_t n_particles_global, work_offset; // id = work_offset + slot + 1
  int64_t master_seed;
  int32_t batch, is_2d, seed_through_double, max_events;
  int32_t n_slots;   // cross-section slots = materials present in the mesh
  int32_t visit_cap; // segments one lane may walk per visit of the tracking kernel
  int32_t record_traces; // keep the per-history crossing hash (parity output)
  int32_t collision_kappa; // some collision-estimator tally scores kappa-fission (needs every fissionable nuclide)
  mcw_trace_point* trace;  // cached trace points (null unless settings.record_traces == 2)
  int64_t trace_capacity;
  int32_t slot_mat[MCW_MAX_XS_SLOTS]; // slot -> material index
  // refresh plan: the (slot, nuclide) pairs of all slots, material order inside a slot (mcw_xs.cuh);
  // n_plan = 0 when some slot's material has no nuclide (then the per-slot kernels are used)
  const int32_t* plan_nuc;
  const int32_t* plan_slot;
  const double* plan_dens;
  const double* plan_e0; // first / last grid energy of the entry's nuclide
  const double* plan_e1;
  int32_t n_plan;
  int32_t n_plan_pad;                           // row stride of xs_prefix (n_plan rounded up to 4 doubles = 32 bytes)
  int32_t slot_plan_off[MCW_MAX_XS_SLOTS + 1]; // plan entries of slot s: [slot_plan_off[s], slot_plan_off[s + 1])
  int32_t pad1;
  double z_coord, keff;
};
