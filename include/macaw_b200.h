/* macaw_b200.h -- C ABI of libmacaw_b200.so, the B200 (sm_100a) implementation of MaCaw's
 * particle-history hot path.
 *
 * This is the drop-in boundary.  In the reference the path sits behind MOOSE objects
 * (OpenMCStudy, CollisionKernel, OpenMCTally, OpenMCTallyAux) that call OpenMC C++ internals per ray
 * segment.  In GPU mode those objects stay in the input file as configuration carriers; the study
 * forwards their parameters through the calls below once, then calls mcw_run_generation() once per
 * batch instead of looping rays through TraceRay::trace().  Every entry point cites the reference
 * interface it replaces as  [file:line]  relative to the reference tree
 * (MC = MaCaw src/, RT = moose/modules/ray_tracing, OMC = openmc/).
 *
 * Conventions (mirroring openmc/include/openmc/capi.h:37-190): every call returns int, 0 = success,
 * negative = MCW_E_* and a message retrievable with mcw_last_error(); the caller owns all input
 * arrays (they are copied during the call); the library owns all device memory; one handle per
 * GPU / rank; calls on one handle come from one host thread.  No torch / C++ types cross the ABI.
 */
#ifndef MACAW_B200_H
#define MACAW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (cf. OPENMC_E_* in OMC include/openmc/capi.h:180-190) ---- */
#define MCW_OK 0
#define MCW_E_INVALID_ARGUMENT (-1)
#define MCW_E_NOT_READY (-2)   /* call order violated (e.g. run before mesh/library/materials set) */
#define MCW_E_CUDA (-3)        /* CUDA runtime error or no sm_100-class device; never falls back to CPU */
#define MCW_E_ALLOCATE (-4)
#define MCW_E_PHYSICS (-5)     /* e.g. no fission sites banked (OMC eigenvalue.cpp:107-109) */
#define MCW_E_UNSUPPORTED (-6)

/* ---- element types (libMesh ElemType subset traced by MaCaw inputs; RT TraceRayTools.h:558-856) ---- */
#define MCW_QUAD4 0
#define MCW_TET4 1
#define MCW_HEX8 2

/* ---- boundary condition kinds (RT src/raybcs/ReflectRayBC.C:39-77, KillRayBC.C) ---- */
#define MCW_BC_KILL 0
#define MCW_BC_REFLECT 1

/* ---- tally scores (OMC tally_scoring.cpp:557,581,726,760,1201), estimator fixed per tally ---- */
#define MCW_SCORE_FLUX 0
#define MCW_SCORE_TOTAL 1
#define MCW_SCORE_FISSION 2
#define MCW_SCORE_NU_FISSION 3
#define MCW_SCORE_KAPPA_FISSION 4
#define MCW_N_SCORES 5

#define MCW_ESTIMATOR_COLLISION 0   /* MC OpenMCTally.C:89-90: the only estimator the reference accepts */
#define MCW_ESTIMATOR_TRACKLENGTH 1 /* extension, spec OMC tally_scoring.cpp:2289-2352 */

/* ---- tally granularity (MC OpenMCTallyAux.C:70-215 "granularity") ---- */
#define MCW_FILTER_CELL 0     /* one bin per mesh element (cell filter, MC OpenMCTally.C:157-171) */
#define MCW_FILTER_UNIVERSE 1 /* one bin per block index (universe filter, MC OpenMCTally.C:141-156) */
#define MCW_FILTER_GLOBAL 2   /* a single spatial bin */

#define MCW_MAX_DELAYED_GROUPS 8 /* OMC particle_data.h:26 */

/* Source / fission-bank site.  Same 96-byte layout as openmc::SourceSite
 * (OMC include/openmc/particle_data.h:43-53); the 4 padding bytes after `particle` carry the mesh
 * element the site was born in (-1 = unknown, located on first use; MC relocates every batch with
 * ClaimRays, RT ClaimRays.C:104-163). */
typedef struct mcw_source_site
{
  double r[3];
  double u[3];
  double E;
  double wgt;
  int32_t delayed_group;
  int32_t surf_id;
  int32_t particle;
  int32_t elem;
  int64_t parent_id;
  int64_t progeny_id;
} mcw_source_site;

/* Distributions tabulated on a grid of incoming energies, the layout OpenMC's HDF5 files use for angular
 * distributions ("angle": `energy` + `mu` (3, N) with `offsets` / `interpolation`, OMC distribution_angle.cpp:20-63)
 * and for continuous tabular outgoing-energy laws ("continuous": `energy` + `distribution` (3, N) with `offsets` /
 * `interpolation` / `n_discrete_lines`, OMC distribution_energy.cpp:56-148): distribution k holds the points
 * [offset[k], offset[k + 1]) of x (mu or E_out), pdf and cdf.  The cdf is used as tabulated (as the reference does to
 * match ACE data) and must end at exactly 1.  n_energy = 0: no table. */
typedef struct mcw_tabular2d
{
  int32_t n_energy;
  int32_t histogram_in_energy;  /* continuous tabular only: the one incoming-energy interpolation region is histogram */
  const double* energy;         /* [n_energy] incoming energies, increasing */
  const int32_t* offset;        /* [n_energy + 1] */
  const int32_t* interpolation; /* [n_energy] 1 = histogram, 2 = linear-linear */
  const int32_t* n_discrete;    /* [n_energy] discrete lines leading each distribution (continuous tabular), or NULL */
  const double* x;
  const double* pdf;
  const double* cdf;
} mcw_tabular2d;

#define MCW_LAW_LEVEL 0              /* LevelInelastic, OMC distribution_energy.cpp:42-51 */
#define MCW_LAW_CONTINUOUS_TABULAR 1 /* ContinuousTabular, OMC distribution_energy.cpp:150-275 */
#define MCW_LAW_CORRELATED 2         /* CorrelatedAngleEnergy, OMC secondary_correlated.cpp:156-262 */
#define MCW_LAW_KALBACH_MANN 3       /* KalbachMann, OMC secondary_kalbach.cpp:117-235 */
#define MCW_LAW_NBODY 4              /* NBodyPhaseSpace, OMC secondary_nbody.cpp:25-68 */
/* evaporation-type energy laws: fission-neutron products (mcw_fission_product) and uncorrelated scattering products */
#define MCW_LAW_MAXWELL 5            /* MaxwellEnergy, OMC distribution_energy.cpp:279-291 */
#define MCW_LAW_EVAPORATION 6        /* Evaporation, OMC distribution_energy.cpp:305-322 */
#define MCW_LAW_WATT 7               /* WattEnergy, OMC distribution_energy.cpp:342-356 */

/* A function of the incoming energy as nuclide files hold them (OMC endf.h Function1D): a Polynomial (coefficients in
 * increasing powers, Horner; endf.cpp:112-121) or a Tabulated1D with one interpolation region, histogram or
 * linear-linear, constant outside its range (endf.cpp:156-210).  kind 0 = absent. */
#define MCW_FN_NONE 0
#define MCW_FN_POLYNOMIAL 1
#define MCW_FN_TABULATED 2
typedef struct mcw_function1d
{
  int32_t kind;
  int32_t n;             /* coefficients, or points */
  int32_t interpolation; /* tabulated: 1 histogram, 2 linear-linear */
  int32_t reserved0;
  const double* x;       /* tabulated: abscissae (increasing); polynomial: unused */
  const double* y;       /* tabulated: ordinates; polynomial: the coefficients */
} mcw_function1d;

/* One neutron product of the fission reaction (OMC ReactionProduct of fission_rx_[0]): the prompt neutrons (first) or one
 * delayed group.  yield = neutrons per fission as a function of the incoming energy; the outgoing energy comes from a
 * Watt spectrum with tabulated a(E), b(E), a Maxwell or an evaporation spectrum with tabulated theta(E) (in `a`) -- each
 * with its restriction energy u: E_out <= E - u -- or a continuous tabular law; the cosine is isotropic
 * (UncorrelatedAngleEnergy without an angle table, secondary_uncorrelated.cpp:54-68). */
typedef struct mcw_fission_product
{
  mcw_function1d yield;
  int32_t law; /* MCW_LAW_WATT, MCW_LAW_MAXWELL, MCW_LAW_EVAPORATION, MCW_LAW_CONTINUOUS_TABULAR */
  int32_t reserved0;
  double u;
  mcw_function1d a, b;
  mcw_tabular2d energy;
} mcw_fission_product;

/* A further scattering reaction of a nuclide besides elastic and the first inelastic one (which mcw_nuclide_desc
 * describes inline): one entry of Nuclide::index_inelastic_scatter_ (OMC nuclide.cpp:216-219 -- every non-redundant
 * reaction with is_inelastic_scatter(mt), endf.cpp:63-78: levels MT 51-91, (n,2n) MT 16 / 875-891, (n,3n), (n,n'alpha)
 * ...).  scatter() walks the list in order, adding each reaction's cross section above its threshold until the sampled
 * cutoff is passed (OMC physics.cpp:676-697), then inelastic_scatter samples that reaction's neutron product and
 * applies its yield (physics.cpp:1073-1120).  Fields as for the inline reaction. */
typedef struct mcw_reaction_desc
{
  int32_t mt;                   /* ENDF MT number (Particle::event_mt; informational) */
  int32_t threshold;            /* grid index of the threshold (xs_.threshold) */
  const double* xs;             /* [n_grid - threshold] */
  double threshold_energy;      /* MCW_LAW_LEVEL: LevelInelastic threshold_ */
  double mass_ratio;            /* MCW_LAW_LEVEL: LevelInelastic mass_ratio_ */
  /* neutron multiplicity of the product, a constant (Polynomial of order 0).  Integral (1 for levels, 2 for (n,2n)):
   * yield - 1 secondary neutrons, which MaCaw pops and drops (MC CollisionKernel.C:232-236), the history itself goes on
   * unchanged.  Not integral: the weight is multiplied by it (OMC physics.cpp:1108-1118).  0 means 1. */
  double yield;
  int32_t law;                  /* MCW_LAW_* (0 ... 7) */
  int32_t in_cm;                /* scatter_in_cm_ (level: always 1) */
  mcw_tabular2d angle;          /* angular distribution (uncorrelated laws), n_energy = 0: isotropic */
  mcw_tabular2d energy;         /* outgoing-energy table of CONTINUOUS_TABULAR, CORRELATED, KALBACH_MANN */
  const int32_t* corr_offset;   /* MCW_LAW_CORRELATED, as mcw_nuclide_desc::corr_* */
  const int32_t* corr_interp;
  const double* corr_mu;
  const double* corr_pdf;
  const double* corr_cdf;
  const double* km_r;           /* MCW_LAW_KALBACH_MANN */
  const double* km_a;
  int32_t nbody_n;              /* MCW_LAW_NBODY */
  int32_t reserved0;
  double nbody_mass_ratio;
  double nbody_q;
  /* An energy-dependent multiplicity (Polynomial / Tabulated1D yield_, OMC reaction_product.cpp:40-43) instead of the
   * constant `yield` (kind 0: the constant), and the parameters of MCW_LAW_MAXWELL / _EVAPORATION / _WATT when the
   * uncorrelated product's energy law is one of those (theta(E) in `a`; Watt a(E), b(E); restriction energy u). */
  mcw_function1d yield_fn;
  double u;
  mcw_function1d a, b;
} mcw_reaction_desc;

/* One continuous-energy nuclide at one temperature (further temperatures: mcw_add_temperature_set).  Field meaning
 * follows openmc::Nuclide (OMC src/nuclide.cpp:50-302, 309-444): `xs` is the derived 5-column table xs_[t](i, col) with
 * col = total, absorption, fission, nu-fission, photon-production, as Nuclide::create_derived sums it.  Reactions:
 * elastic (MT 2; isotropic in CM or a tabulated angular distribution; free-gas target below 400 kT), a list of further
 * scattering reactions -- the first one inline, the others as mcw_reaction_desc -- each with its law (MCW_LAW_*), system,
 * and multiplicity; radiative capture and the other disappearance reactions (in `xs` only); fission, described either by
 * scalars (polynomial total nu, constant delayed yields, Watt prompt and Maxwell delayed spectra) or by
 * mcw_fission_product entries as a nuclide file has them. */
typedef struct mcw_nuclide_desc
{
  int32_t n_grid;
  int32_t fissionable;
  double awr;
  double kT;               /* temperature of the data set, eV (OMC nuclide.cpp kTs_) */
  const double* energy;    /* [n_grid] strictly non-decreasing */
  const double* xs;        /* [n_grid][5] row-major */
  const double* elastic;   /* [n_grid] MT 2 */
  int32_t n_inelastic;     /* 0 or 1: the first inelastic reaction, described inline (further ones: more_reactions) */
  int32_t inel_threshold;  /* grid index of the reaction threshold (OMC reaction xs_.threshold) */
  const double* inelastic; /* [n_grid - inel_threshold] */
  double inel_threshold_energy; /* LevelInelastic threshold_ */
  double inel_mass_ratio;       /* LevelInelastic mass_ratio_ */
  double nu_poly[2];            /* total nu(E) = c0 + c1*E (OMC endf.cpp Polynomial) */
  int32_t n_precursor;          /* delayed groups, <= MCW_MAX_DELAYED_GROUPS */
  int32_t reserved0;
  double delayed_yield[MCW_MAX_DELAYED_GROUPS]; /* constant yield per group */
  double delayed_theta[MCW_MAX_DELAYED_GROUPS]; /* Maxwell temperature per group, eV */
  double watt_a, watt_b;                        /* prompt Watt spectrum parameters */
  double q_fission;                             /* fission Q value, eV (kappa-fission score) */
  /* Secondary distributions beyond "isotropic in the centre of mass" (all optional, zero = absent):
   * tabular angular distributions of the elastic and of the inelastic reaction (AngleDistribution::sample, OMC
   * distribution_angle.cpp:65-95, used by elastic_scatter physics.cpp:749-757 and UncorrelatedAngleEnergy::sample
   * secondary_uncorrelated.cpp:54-68), and the outgoing-energy law of the inelastic reaction. */
  mcw_tabular2d elastic_angle;
  mcw_tabular2d inelastic_angle;
  int32_t inelastic_law;                        /* MCW_LAW_* (default MCW_LAW_LEVEL) */
  int32_t inelastic_in_cm;                      /* scatter_in_cm_ of the inelastic reaction (level: always 1) */
  mcw_tabular2d inelastic_energy;               /* outgoing-energy table of CONTINUOUS_TABULAR, CORRELATED, KALBACH_MANN */
  /* MCW_LAW_CORRELATED (product distribution "correlated"; replaces the angle table): one angular Tabular per POINT
   * of inelastic_energy -- point j holds the cosines [corr_offset[j], corr_offset[j + 1]) of corr_mu / corr_pdf /
   * corr_cdf, interpolation corr_interp[j] (1 histogram, 2 lin-lin). */
  const int32_t* corr_offset;
  const int32_t* corr_interp;
  const double* corr_mu;
  const double* corr_pdf;
  const double* corr_cdf;
  /* MCW_LAW_KALBACH_MANN ("kalbach-mann"): precompound fraction r and angular slope a per point of inelastic_energy */
  const double* km_r;
  const double* km_a;
  /* MCW_LAW_NBODY ("nbody"): n_particles (3, 4 or 5), total_mass and q_value; A is the nuclide's awr */
  int32_t nbody_n;
  int32_t reserved1;
  double nbody_mass_ratio;
  double nbody_q;
  /* Neutron multiplicity of the inline inelastic reaction (see mcw_reaction_desc::yield; 0 means 1), and the further
   * scattering reactions that follow it in Nuclide::index_inelastic_scatter_ order (needs n_inelastic = 1). */
  double inelastic_yield;
  int32_t n_more_reactions;
  int32_t reserved2;
  const mcw_reaction_desc* more_reactions;
  /* Fission neutrons as a nuclide file has them, instead of the scalar description above (nu_poly, delayed_yield,
   * delayed_theta, watt_a, watt_b): n_fission_products = 1 + n_precursor products (prompt, then the delayed groups) and the
   * total nu (kind 0: the prompt yield stands in, OMC nuclide.cpp:510-516).  sample_fission_neutron (physics.cpp:1006-1071)
   * evaluates beta = sum of the delayed yields / total nu, picks the group by the yields and samples the product's law,
   * resampling energies >= the problem's maximum.  0 products: the scalar description. */
  int32_t n_fission_products;
  int32_t reserved3;
  const mcw_fission_product* fission_products;
  mcw_function1d total_nu;
  /* the inline inelastic reaction's mcw_reaction_desc::yield_fn, u, a, b */
  mcw_function1d inelastic_yield_fn;
  double inelastic_u;
  mcw_function1d inelastic_a, inelastic_b;
} mcw_nuclide_desc;

/* One material (OMC src/material.cpp: nuclide_, atom_density_ in atom/b-cm).  isotropic_in_lab
 * (may be NULL) flags the nuclides of the material's <isotropic> list: after a scatter off such a
 * nuclide the outgoing direction is resampled isotropically in the laboratory frame
 * [OMC material.cpp:255-270 p0_, physics.cpp:706-718]; MaCaw's own study test uses it for every
 * nuclide (test/tests/study/materials.xml). */
typedef struct mcw_material_desc
{
  int32_t n_nuclides;
  int32_t reserved0;
  const int32_t* nuclide;          /* indices into the library */
  const double* atom_density;      /* [n_nuclides] */
  const int32_t* isotropic_in_lab; /* [n_nuclides] 0 / 1, or NULL */
} mcw_material_desc;

/* Thermal scattering S(alpha, beta) table at one temperature: openmc::ThermalScattering / ThermalData (OMC thermal.cpp:
 * 38-282) with the angle-energy distributions of secondary_thermal.cpp.  Bound-atom cross sections replace the free-atom
 * elastic one below energy_max (Nuclide::calculate_sab_xs, OMC nuclide.cpp:763-791; Material::calculate_neutron_xs,
 * material.cpp:778-806) and sab_scatter (physics.cpp:788-802) samples the outgoing energy and cosine.
 * Inelastic: Tabulated1D cross section (lin-lin, endf.cpp:153-212) with IncoherentInelasticAE (continuous outgoing
 * energies with n_mu equiprobable cosines each, secondary_thermal.cpp:190-332) or IncoherentInelasticAEDiscrete
 * (:128-188: equiprobable or skewed outgoing energies on the cross section's own energy grid).
 * Elastic (optional): CoherentElasticXS (Bragg edges and partial sums of structure factors, endf.cpp:214-239) with
 * CoherentElasticAE (secondary_thermal.cpp:27-58), or a Tabulated1D cross section with IncoherentElasticAEDiscrete
 * (:88-126).  The Debye-Waller form (IncoherentElasticXS / IncoherentElasticAE) is not carried. */
#define MCW_SAB_INELASTIC_CONTINUOUS 0
#define MCW_SAB_INELASTIC_DISCRETE 1
#define MCW_SAB_ELASTIC_NONE 0
#define MCW_SAB_ELASTIC_COHERENT 1
#define MCW_SAB_ELASTIC_INCOHERENT_DISCRETE 2
typedef struct mcw_thermal_desc
{
  double energy_max;                /* energy_max_: the table is not used above it */
  int32_t inelastic_type;           /* MCW_SAB_INELASTIC_* */
  int32_t elastic_type;             /* MCW_SAB_ELASTIC_* */
  int32_t n_inelastic_xs;           /* inelastic cross section: points of the Tabulated1D */
  int32_t n_inelastic_energy;       /* CONTINUOUS: incident energies of the distribution */
  const double* inelastic_xs_energy;
  const double* inelastic_xs;
  const double* inelastic_energy;   /* CONTINUOUS: [n_inelastic_energy] */
  const int32_t* inelastic_offset;  /* CONTINUOUS: [n_inelastic_energy + 1]; distribution l = points [offset[l], offset[l+1]) */
  const double* inelastic_e_out;    /* CONTINUOUS: per point.  DISCRETE: [n_inelastic_xs][n_e_out] */
  const double* inelastic_pdf;      /* CONTINUOUS */
  const double* inelastic_cdf;      /* CONTINUOUS, used as tabulated */
  const double* inelastic_mu;       /* CONTINUOUS: [point][n_mu].  DISCRETE: [n_inelastic_xs][n_e_out][n_mu] */
  int32_t n_mu;
  int32_t n_e_out;                  /* DISCRETE */
  int32_t skewed;                   /* DISCRETE */
  int32_t n_elastic;                /* elastic: Bragg edges (COHERENT) or points of the Tabulated1D */
  const double* elastic_x;
  const double* elastic_y;          /* COHERENT: partial sums of structure factors; else the cross section */
  const double* elastic_mu;         /* INCOHERENT_DISCRETE: [n_elastic][n_elastic_mu] */
  int32_t n_elastic_mu;
  int32_t reserved0;
} mcw_thermal_desc;

/* Unresolved-resonance probability tables of one nuclide at its temperature: openmc::UrrData (OMC urr.cpp:7-33), used by
 * Nuclide::calculate_urr_xs (nuclide.cpp:795-924) when the energy lies strictly inside the table's range.  The band is
 * picked with one random number of the history's URR stream, future_prn(nuclide index, seed) -- the stream MaCaw carries
 * as ray aux data and advances by n_nuclides after every collision that changes the energy (OMC physics.cpp:140-144) --
 * so the cross sections of a history depend on its URR seed.  Elastic scattering off a nuclide in its unresolved range
 * takes the target at rest (physics.cpp:735-738).  Linear-linear tables only (interpolation = 2). */
typedef struct mcw_urr_desc
{
  int32_t nuclide;          /* library index */
  int32_t interpolation;    /* 2 = lin-lin; log-log (5) is not carried */
  int32_t inelastic_flag;   /* -1 = none (C_NONE), else the MT of the competing inelastic reaction (the nuclide's own: 51) */
  int32_t absorption_flag;  /* read by the reference, not used in the cross-section calculation */
  int32_t multiply_smooth;  /* 1: the table holds factors on the smooth cross sections */
  int32_t n_energy;
  int32_t n_band;
  int32_t reserved0;
  const double* energy;     /* [n_energy] */
  const double* table;      /* [n_energy][6][n_band]: cumulative probability, total, elastic, fission, (n,gamma), heating */
} mcw_urr_desc;

/* Run settings: the subset of openmc::settings MaCaw reads (OMC settings.cpp:41-116) plus the
 * MaCaw/MOOSE parameters that change the numbers (MC CollisionKernel.C:38, RT RayTracingStudy.C). */
typedef struct mcw_settings
{
  int64_t n_particles;     /* settings::n_particles (global, all ranks) */
  int64_t master_seed;     /* OMC random_lcg.cpp:9, default 1 */
  int32_t n_inactive;
  int32_t n_log_bins;      /* settings::n_log_bins, default 8000 */
  int32_t is_2d;           /* mesh dimension == 2 (MC CollisionKernel.C:47) */
  int32_t seed_through_double; /* 1 = reproduce MaCaw's double-typed ray aux seeds */
  double z_coord;          /* MC CollisionKernel.C:38 */
  double source_box[6];    /* batch-1 box source lower-left / upper-right (OMC source.cpp:143-231) */
  int32_t source_only_fissionable;
  int32_t max_events;      /* safety cap on events per history (OMC constants.h:56 MAX_EVENTS) */
  int32_t sort_xs_queue;   /* 1 = sort the cross-section queue by (material, energy bucket) */
  int32_t record_traces;   /* 1 = keep per-history crossing hashes (parity output); 2 = also cache trace points */
} mcw_settings;

/* Per-generation result (MC OpenMCStudy.C:399-401 prints keff / keff_std / k_generation;
 * counters follow RT PerProcessorRayTracingResultsVectorPostprocessor names). */
typedef struct mcw_generation_result
{
  double k_generation;     /* OMC eigenvalue.cpp:49-69 (track-length based) */
  double keff;             /* running mean over active generations, OMC eigenvalue.cpp:313-351 */
  double keff_std;
  double k_collision;      /* this generation's global tallies / total weight */
  double k_absorption;
  double k_tracklength;
  double leakage;          /* always 0: MaCaw's kill boundary ends the ray without scoring OpenMC's leakage global tally
                            * (that happens in Particle::cross_vacuum_bc, which the MOOSE path never calls) */
  double total_weight;
  int64_t n_histories;
  int64_t n_segments;      /* "intersections" */
  int64_t n_collisions;
  int64_t n_xs_lookups;    /* macroscopic lookups performed */
  int64_t n_nuclide_lookups;
  int64_t n_bank_sites;    /* fission sites banked before sampling */
  int64_t n_vertex_edge_hits; /* exits through a vertex or edge (RT VERTEX_HITS + EDGE_HITS) */
  int64_t n_moved_through_neighbors; /* RT MOVED_THROUGH_NEIGHBORS */
  int64_t n_failed_traces;    /* RT FAILED_TRACES (tolerate_failure semantics: counted and killed) */
  int64_t n_event_rounds;     /* event-loop iterations */
  int64_t n_kernel_launches;  /* kernels of this library launched in the generation */
  double ms_total;            /* device time of the generation (CUDA events) */
  double ms_xs, ms_track, ms_collide, ms_bank; /* per-kernel-family device time */
  int64_t n_bank_dropped;     /* fission sites not banked because the bank (3 x this rank's histories, OMC
                               * simulation.cpp:286) was full: OMC physics.cpp:203-209 warns "The shared fission
                               * bank is full"; here the count is returned and mcw_last_error carries the warning */
} mcw_generation_result;

/* Per-history record for parity checks (device -> host on request). */
typedef struct mcw_history_record
{
  uint64_t seed_tracking, seed_source, seed_urr; /* state at death */
  double E;
  double k_collision, k_absorption, k_tracklength;
  int32_t n_segments, n_collisions, n_progeny;
  int32_t last_elem;
  int32_t fate;        /* 0 absorbed, 1 killed on boundary, 2 failed trace, 3 event cap */
  int32_t n_xs;
  uint64_t trace_hash; /* rolling hash of (element, exit side) for every segment */
} mcw_history_record;

/* One cached trace point (RT include/raytracing/TraceData.h:15-29 TracePointData + the ray's aux data): what
 * RayTracingStudy caches per completed segment when `segments_on_cache_traces` / `aux_data_on_cache_traces` are on
 * and RayTracingExodus writes as one EDGE2 element per segment (RT src/outputs/RayTracingMeshOutput.C:252-445,
 * 447-580).  A segment in which the history scatters is restarted from the collision point and caches nothing
 * (SURVEY F7): every cached point lies on an element side.  Kept when settings.record_traces == 2. */
typedef struct mcw_trace_point
{
  int64_t ray_id;  /* RT ray id of the history: (batch - 1) * n_particles + (id - 1) [SURVEY A13] */
  int32_t seq;     /* segments the history had walked when the point was cached (orders a ray's points) */
  int32_t kind;    /* 0 start point, 1 end of a completed segment (an element side), 3 absorbed in the element (last point) */
  double point[3];
  double aux[8];   /* energy, weight, n_progeny, id, seed_tracking, seed_source, seed_URR, particle_type
                    * (registration order of MC OpenMCStudy.C:190-197), as after the segment */
} mcw_trace_point;

typedef struct mcw_context* mcw_handle;

/* Lifetime.  Replaces OpenMCStudy's ctor/dtor calls openmc_init / openmc_simulation_init /
 * openmc_finalize [MC OpenMCStudy.C:129,147,213; OMC capi.h:37,56,105].  rank/nranks describe the
 * particle sharding (OMC simulation.cpp:529-551 calculate_work). */
int mcw_create(mcw_handle* h, int device_ordinal, int rank, int nranks);
void mcw_destroy(mcw_handle h);
const char* mcw_last_error(mcw_handle h); /* OMC error.cpp:32 openmc_err_msg */
int mcw_version(void);
/* Run-time knobs that do not change any per-history result (tally / k sums move by FP64 addition
 * order only):
 *   "profile"          1 = per-kernel-family CUDA-event timing in mcw_generation_result
 *   "sync_every"       event rounds between host reads of the queue counters (default 4)
 *   "tail_threshold"   histories in flight below which k_track runs fused, collisions inline (16384)
 *   "pump_threshold"   histories in flight below which the REST of the generation runs in the persistent kernel
 *                      k_pump (tracking / collision / general-walk warps connected by device-side rings, one
 *                      launch, no host round trip); 0 = off (default).  "pump_track_warps" (1-3 of a CTA's 4 warps
 *                      track, default 2), "pump_patience" (polls a collision warp waits for a full batch, default 4)
 *   "collide_min"      fused mode: parked collisions per warp before they are processed (8)
 *   "ordered_queues"   1 = collide / fly queues rebuilt in history-index order every round (1)
 *   "xs_sort_min"      with settings.sort_xs_queue: smallest collision queue worth an energy sort
 *                      (default: max(32768, 2.4e6 / number of (material, nuclide) pairs in use))
 *   "xs_plan"          stand-alone sorted lookup: 1 = tile kernel (all windows of a pass staged at once), 0 = per nuclide
 *   "fallback_every_round"  1 = serve the general-walk queue in every round instead of after host syncs (0)
 *   "drain_lanes", "drain_min_queue", "visit_cap", "fallback_blocks"   tracking / rare-path launch shape */
int mcw_set_option(mcw_handle h, const char* name, int value);
/* Builds all device state now (otherwise done lazily by the first call that needs it). */
int mcw_prepare(mcw_handle h);
/* The CUDA stream (cudaStream_t) every kernel of this handle is launched on, so a caller can time
 * with its own events or enqueue collectives in order. */
int mcw_stream(mcw_handle h, void** stream);
int mcw_struct_sizes(int32_t* out, int n); /* sizeof of the six ABI structs, for binding self-checks */

/* Mesh.  Replaces libMesh access on the path: elem->point / neighbor_ptr / subdomain_id
 * [RT TraceRayTools.C:404-407, TraceRay.C:1433, MC CollisionKernel.C:110] and boundary info
 * [RT TraceRay.C:725-782].  conn is n_elem x (4 | 4 | 8) node ids in libMesh node order
 * (SURVEY Appendix B); block_id holds raw subdomain ids.  neighbors (n_elem x n_sides, -1 on the
 * boundary) may be NULL (computed by face matching).  side_bnd_id is n_elem x n_sides with the
 * boundary (sideset) id of every boundary side, -1 elsewhere; bnd_kind[n_bnd] gives the RayBC
 * applied on each boundary id (MCW_BC_KILL / MCW_BC_REFLECT; RT TraceRay.C:2043-2146 looks the
 * RayBCs up per boundary id the same way).  A boundary side without an id kills the ray. */
int mcw_set_mesh(mcw_handle h, int elem_type, int64_t n_nodes, const double* xyz, int64_t n_elem,
                 const int64_t* conn, const int32_t* block_id, const int64_t* neighbors,
                 const int32_t* side_bnd_id, int n_bnd, const int32_t* bnd_kind);

/* Nuclear data.  Replaces HDF5 loading in openmc_init + Nuclide::init_grid
 * [OMC nuclide.cpp:50-302,446-476; simulation.cpp:553-619]. */
int mcw_set_library(mcw_handle h, const mcw_nuclide_desc* nuclides, int n);

/* A 64-bit digest (FNV-1a) of everything an array of descriptors says -- scalars and every array they point to, in
 * declaration order, what the library would copy -- so that two hosts (an adapter and a reference run, the C++ and the
 * Python host of this repository) can check that they hand over the same data.  kind 0: mcw_nuclide_desc, 1:
 * mcw_thermal_desc, 2: mcw_urr_desc.  Host code only, no handle, no device. */
int mcw_digest(int kind, const void* descriptors, int n, uint64_t* out);

/* One more data set of nuclide `nuclide` (its index in mcw_set_library's array) at the temperature set->kT: OpenMC's
 * Nuclide::kTs_[t] with grid_[t], xs_[t], every reaction's xs_[t] and threshold [OMC nuclide.cpp:170-200,
 * reaction.cpp:25-50].  Everything temperature dependent is taken from `set` (energy, xs, elastic, the inelastic reactions'
 * cross sections and threshold indices); it must hold the same reactions as the nuclide, whose secondary distributions
 * apply at every temperature.  *entry receives the library entry of the set (>= the number of nuclides): the index
 * mcw_urr_desc::nuclide uses for that temperature's probability tables.  Materials and tally nuclide bins keep naming
 * nuclides.  At run time a history in an element of temperature T (mcw_set_temperature; MC CollisionKernel.C:109:
 * sqrtkT = sqrt(K_BOLTZMANN T)) uses, for every nuclide, the data set whose kT is nearest to sqrtkT^2 -- the first of
 * equally near ones by increasing kT -- [OMC nuclide.cpp:614-626, TemperatureMethod::NEAREST, the default]: its cross
 * sections, its kT for the free-gas target [physics.cpp:649-652], its probability tables.  The energy range of the problem
 * comes from every nuclide's lowest-temperature grid [simulation.cpp:558-566].  S(alpha, beta) tables have their own
 * temperatures: mcw_set_thermal_temperatures.  Call after mcw_set_library (which discards earlier sets) and before mcw_set_urr_tables. */
int mcw_add_temperature_set(mcw_handle h, int nuclide, const mcw_nuclide_desc* set, int* entry);

/* The checks mcw_set_library makes, without a handle and without a device (host code only): grids sorted, tables complete,
 * every law the library implements, reaction lists well formed.  Returns the code mcw_set_library would return and writes
 * its message (what mcw_last_error would say) into `message`.  For adapters that want to refuse a nuclide they cannot
 * express before any GPU work starts (OMC raises the matching fatal_error while reading the file, e.g.
 * secondary_uncorrelated.cpp:47, distribution.cpp:201-205). */
int mcw_validate_library(const mcw_nuclide_desc* nuclides, int n, char* message, int message_bytes);

/* Materials and the block -> material map of CollisionKernel's `blocks` / `materials` parameters
 * [MC CollisionKernel.C:34-37,64-69,110]. */
int mcw_set_materials(mcw_handle h, const mcw_material_desc* materials, int n,
                      const int32_t* block_ids, const int32_t* block_material, int n_blocks);

/* Thermal scattering tables of the library (data::thermal_scatt, OMC thermal.cpp:30-32) and their assignment to the
 * nuclides of a material: `<sab name= fraction=>` of materials.xml -> Material::thermal_tables_ (index_table,
 * index_nuclide, fraction; OMC material.cpp:140-160,317-360).  nuclide_position indexes the material's own nuclide
 * list.  Call after mcw_set_library / mcw_set_materials. */
int mcw_set_thermal_tables(mcw_handle h, const mcw_thermal_desc* tables, int n);
int mcw_set_material_thermal(mcw_handle h, int material, int n, const int32_t* nuclide_position, const int32_t* table,
                             const double* fraction);

/* Several temperatures of S(alpha, beta) tables (openmc::ThermalScattering::kTs_ / data_[t], OMC thermal.cpp:38-147):
 * entry k of mcw_set_thermal_tables is the temperature kT[k] of the table whose FIRST entry is same_table_as[k] (k itself
 * for a table's first or only temperature); n = the number of entries.  Materials (mcw_set_material_thermal) name first
 * entries.  A history in an element of temperature T uses, for every table, the closer of the two temperatures that bound
 * K_BOLTZMANN T -- ThermalScattering::calculate_xs with the nearest-temperature method [thermal.cpp:149-171] -- for the
 * bound cross sections and for sab_scatter.  Call after mcw_set_thermal_tables (which discards it). */
int mcw_set_thermal_temperatures(mcw_handle h, const int32_t* same_table_as, const double* kT, int n);

/* Probability tables of the nuclides that have them (`urr/<T>K` of a nuclide file, OMC nuclide.cpp:225-270).  Call after
 * mcw_set_library.  The stand-alone lookups mcw_xs_lookup / mcw_xs_lookup_device have no history and therefore no URR
 * stream: they return the smooth cross sections. */
int mcw_set_urr_tables(mcw_handle h, const mcw_urr_desc* tables, int n);

/* Coupled temperature variable of CollisionKernel in K, one value per element or one scalar (n = 1)
 * [MC CollisionKernel.C:33,109]; not set: 0 K, as in MaCaw's own study test.  Picks each nuclide's data set
 * (mcw_add_temperature_set) and each S(alpha, beta) table's temperature (mcw_set_thermal_temperatures); with one data set
 * everywhere it has no effect, like NEAREST in the reference.  The picks are part of the prepared problem: a field that
 * differs from the current one, on a library with several temperatures, prepares the problem again on the next call
 * that needs it -- which, like every change of the problem, starts the run over (source bank, k statistics, tally sums).
 * A coupled run that updates the field between time steps carries its source bank across with mcw_get_source_bank /
 * mcw_set_source_bank (host/macaw_host.hpp OpenMCStudy::setTemperature does). */
int mcw_set_temperature(mcw_handle h, const double* T, int64_t n);

int mcw_set_settings(mcw_handle h, const mcw_settings* s);

/* Tally definition.  Replaces OpenMCTally::initialize [MC OpenMCTally.C:103-225]. */
int mcw_add_tally(mcw_handle h, int id, int estimator, int granularity, const int32_t* scores,
                  int n_scores, const double* energy_bins, int n_energy_edges);

/* Nuclide bins of a tally: the `nuclides` parameter of OpenMCTally [MC OpenMCTally.C:56-58,188-204 ->
 * OMC tallies/tally.cpp:543-558 Tally::set_nuclides].  nuclides[j] = index of a library nuclide, or -1 for the
 * material total ("total"), or -2 for a bin that is allocated but never scored.  The caller spells the reference's
 * conventions out: a list of names = the list followed by -1 (OpenMCTally.C:198-202); "all" = 0 .. n_nuclides-1, -1, -2:
 * OpenMCTally.C:190-197 appends every nuclide and -1 to the tally's default {-1} WITHOUT clearing it (OpenMC's own
 * reader clears first, OMC tally.cpp:517-533), so results_ has n_nuclides + 2 nuclide bins of which
 * score_all_nuclides (OMC tally_scoring.cpp:2142-2175) fills bin i for nuclide i and bin n_nuclides for the total.  A bin whose nuclide is not in the material of the event scores nothing
 * (OMC tally_scoring.cpp:2391-2396); the others score micro * atom_density * flux (tally_scoring.cpp:516-1260).
 * Results gain a nuclide dimension: score_index = nuclide_bin * n_scores + score, as in OpenMC's results_.  Without
 * this call a tally has the single bin "total". */
int mcw_set_tally_nuclides(mcw_handle h, int id, const int32_t* nuclides, int n);

/* Source bank.  mcw_sample_initial_source replaces initialize_source [OMC source.cpp:325-338];
 * mcw_set_source_bank / mcw_get_source_bank replace openmc_source_bank [OMC bank.cpp:110-125].
 * n is this rank's share (work_per_rank). */
int mcw_sample_initial_source(mcw_handle h);
int mcw_set_source_bank(mcw_handle h, const mcw_source_site* sites, int64_t n);
int mcw_get_source_bank(mcw_handle h, mcw_source_site* sites, int64_t capacity, int64_t* n);

/* One batch = one generation.  Replaces OpenMCStudy::execute -> generateRays / trace loop /
 * postExecuteStudy on one rank [MC OpenMCStudy.C:218-402; RT TraceRay.C:999-1592;
 * MC CollisionKernel.C:98-275].  Split in two so a multi-rank caller can exchange between them:
 *   mcw_transport_generation : initialize_batch/generation, histories, local fission bank sorted by
 *                              (parent_id, progeny_id) [OMC bank.cpp:60-104]
 *   mcw_finalize_generation  : synchronize_bank sampling on this rank's slice of the global bank,
 *                              k estimators, tally accumulation [OMC eigenvalue.cpp:49-351,
 *                              simulation.cpp:334-449, tallies/tally.cpp:639-670,805-846]
 * mcw_run_generation = both, for a single rank. */
int mcw_transport_generation(mcw_handle h, int batch);
int mcw_finalize_generation(mcw_handle h, int batch, mcw_generation_result* out);
int mcw_run_generation(mcw_handle h, int batch, mcw_generation_result* out);

/* Multi-rank plumbing: raw device views so the caller can run NCCL collectives in place
 * (reference call sites: MC OpenMCStudy.C:467,477,548,559; OMC eigenvalue.cpp:60,233,279;
 * tallies/tally.cpp:761-798).  Pointers stay valid until the next mcw_set_* / destroy call. */
int mcw_bank_count(mcw_handle h, int64_t* n_local);               /* local fission sites */
int mcw_bank_device_ptr(mcw_handle h, void** sites, int64_t* n);  /* sorted local fission bank */
int mcw_set_bank_window(mcw_handle h, int64_t global_start, int64_t global_total); /* Exscan/Bcast result */
/* synchronize_bank sampling on this rank's slice [OMC eigenvalue.cpp:112-160], then the last-rank
 * trim / pad to exactly n_particles [OMC eigenvalue.cpp:184-206; MC OpenMCStudy.C:557-588]. */
int mcw_sample_bank(mcw_handle h, int batch, int64_t* n_sampled);
int mcw_finish_sampling(mcw_handle h, int64_t global_start_of_sampled, int is_last, int64_t* n_sampled);
int mcw_sampled_device_ptr(mcw_handle h, void** sites, int64_t* n); /* sites this rank sampled */
int mcw_source_device_ptr(mcw_handle h, void** sites, int64_t* capacity); /* where received sites go */
int mcw_set_source_count(mcw_handle h, int64_t n);
int mcw_reduction_device_ptr(mcw_handle h, void** doubles, int64_t* n); /* [tally VALUEs | k_col,k_abs,k_trk,leak | weight] */

/* Tally read-back.  Replaces OpenMCTallyAux::computeValue [MC OpenMCTallyAux.C:70-215]:
 * results layout = OpenMC results_(filter_bin, score, {VALUE,SUM,SUM_SQ}); filter_bin =
 * spatial_bin * n_energy_bins + energy_bin (filters sorted universe, cell, energy;
 * MC OpenMCTally.h:71-88). */
int mcw_get_tally(mcw_handle h, int id, double* sum, double* sum_sq, int64_t n, int32_t* n_realizations);
int mcw_tally_shape(mcw_handle h, int id, int64_t* n_spatial, int32_t* n_energy, int32_t* n_scores);
/* number of nuclide bins (1 unless mcw_set_tally_nuclides was called): values per filter bin = bins * n_scores */
int mcw_tally_nuclide_bins(mcw_handle h, int id, int32_t* n_nuclide_bins);

/* Parity / diagnostics output (RT TraceData.h:20-68 cached traces are the reference analogue). */
int mcw_get_history_records(mcw_handle h, mcw_history_record* rec, int64_t capacity, int64_t* n);
/* Cached trace points of the last generation, ordered by (ray_id, seq); *n = how many were cached (may exceed
 * capacity).  The device buffer holds 2048 points per history of this rank, at most 2^23; a generation that
 * caches more fails with MCW_E_ALLOCATE -- tracing is a parity / visualisation output for small runs. */
int mcw_get_traces(mcw_handle h, mcw_trace_point* out, int64_t capacity, int64_t* n);
int mcw_get_fission_bank(mcw_handle h, mcw_source_site* sites, int64_t capacity, int64_t* n);
int mcw_get_grid_index(mcw_handle h, int i_nuclide, int32_t* out, int n /* n_log_bins + 1 */);

/* Stand-alone kernels of the path, exposed for unit parity and the XS-lookup bandwidth metric.
 * mcw_xs_lookup: macroscopic cross sections for n (energy, material) pairs in HOST memory
 * [OMC material.cpp:763-831 + nuclide.cpp:557-763]; out = n x 4 (total, absorption, fission,
 * nu-fission).  mcw_xs_lookup_device: same on device pointers, returns device milliseconds per
 * launch; sorted = 0: one thread per point reading the tables directly; 1: the caller's arrays are
 * already ordered by (material, energy), shared-memory-staged kernel; 2: the library orders the
 * points itself (keys + 32-bit radix sort, inside the timed region) and runs the staged kernel,
 * results in caller order.  All three are bit-identical. */
int mcw_xs_lookup(mcw_handle h, int64_t n, const double* E, const int32_t* material, double* out);
int mcw_xs_lookup_device(mcw_handle h, int64_t n, const void* d_E, const void* d_material,
                         void* d_out, int sorted, int repeats, float* ms_per_launch);
/* mcw_trace_rays: walk n rays through the mesh with no physics until a kill boundary (the RT
 * `lots.i` tests' workload); out_distance[n], out_segments[n], out_hash[n]. */
int mcw_trace_rays(mcw_handle h, int64_t n, const double* start, const double* dir,
                   const int32_t* start_elem, const int32_t* start_side /* NULL, or -1 none / -2 find */,
                   const double* max_distance /* NULL = unbounded */, double* out_distance,
                   int32_t* out_segments, uint64_t* out_hash);
/* mcw_locate_points: starting element by the ClaimRays rule [RT ClaimRays.C:104-139]. */
int mcw_locate_points(mcw_handle h, int64_t n, const double* xyz, const int64_t* ray_id, int32_t* elem);

/* Diagnostics: the FP64 division the tile lookup inlines (the compiler's own expansion of a / b written out so that
 * several divisions overlap) next to the plain operator, for n host operand pairs; ok[i] = the expansion's range
 * checks held, i.e. q_fast[i] is used as is -- it must then equal q_plain[i] bit for bit. */
int mcw_test_division(mcw_handle h, int64_t n, const double* a, const double* b, double* q_fast, double* q_plain,
                      int32_t* ok);

/* Host-side helpers that need no GPU (work split and bank sampling plan; tested on CPU). */
int mcw_calculate_work(int64_t n_particles, int nranks, int64_t* work_index /* nranks+1 */);

#ifdef __cplusplus
}
#endif
#endif /* MACAW_B200_H */
