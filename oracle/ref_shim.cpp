// ref_shim.cpp -- extern "C" doorways into the REFERENCE's own compiled sources (openmc/src: random_lcg, random_dist,
// math_functions, distribution*, position, bank, eigenvalue, simulation, settings, particle, tallies/tally, physics, ... --
// the REF_SRCS list of oracle/Makefile -- + vendored pugixml / Faddeeva), linked into oracle/_ref/libomc_ref.so.
// Test infrastructure only: used to pin the oracle's RNG, samplers, bank / keff / tally bookkeeping and history
// initialisation bit for bit against the real reference code (ref_shim_tally.cpp and ref_shim_physics.cpp do the same for
// tally scoring and for lookup + collision physics).  Nothing here restates an algorithm; where a few lines of glue from
// a function that cannot be linked are repeated, the comment at that spot says which and why.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <cstring>

#include "openmc/bank.h"
#include "openmc/distribution_multi.h"
#include "openmc/eigenvalue.h"
#include "openmc/message_passing.h"
#include "openmc/settings.h"
#include "openmc/simulation.h"
#include "openmc/tallies/tally.h"
#include "openmc/distribution_spatial.h"
#include "openmc/particle.h"
#include "openmc/physics.h"
#include "openmc/search.h"
#include "pugixml.hpp"
#include "xtensor/xview.hpp"
#include "openmc/math_functions.h"
#include "openmc/position.h"
#include "openmc/random_dist.h"
#include "openmc/random_lcg.h"

extern "C" {
// OMC bank.cpp:60-104 sort_fission_bank + eigenvalue.cpp:71-311 synchronize_bank (single process), the
// reference's own functions on the reference's own globals: fission bank in, sorted bank and next source bank out
int64_t ref_sort_and_synchronize_bank(const void* fission_sites, int64_t n_sites, const int64_t* progeny_per_particle,
                                      int64_t n_particles, int total_gen, int current_batch, int current_gen,
                                      int gen_per_batch, void* sorted_out, void* source_out)
{
  using namespace openmc;
  static_assert(sizeof(SourceSite) == 96, "SourceSite layout (particle_data.h:43-53)");
  settings::n_particles = n_particles;
  settings::gen_per_batch = gen_per_batch;
  simulation::total_gen = total_gen;
  simulation::current_batch = current_batch;
  simulation::current_gen = current_gen;
  mpi::n_procs = 1;
  mpi::rank = 0;
  simulation::work_index = {0, n_particles};
  simulation::work_per_rank = n_particles;
  simulation::fission_bank.reserve(3 * n_particles);
  simulation::fission_bank.resize(0);
  const SourceSite* in = static_cast<const SourceSite*>(fission_sites);
  for (int64_t i = 0; i < n_sites; ++i)
    simulation::fission_bank.thread_safe_append(in[i]);
  simulation::progeny_per_particle.assign(progeny_per_particle, progeny_per_particle + n_particles);
  simulation::source_bank.resize(n_particles);
  sort_fission_bank();
  std::memcpy(sorted_out, simulation::fission_bank.data(), sizeof(SourceSite) * n_sites);
  synchronize_bank();
  std::memcpy(source_out, simulation::source_bank.data(), sizeof(SourceSite) * n_particles);
  return (int64_t)simulation::source_bank.size();
}
// OMC simulation.cpp:529-551 calculate_work: the particle split over ranks (work_index has n_procs + 1 entries)
int64_t ref_calculate_work(int64_t n_particles, int n_procs, int rank, int64_t* work_index)
{
  using namespace openmc;
  settings::n_particles = n_particles;
  mpi::n_procs = n_procs;
  mpi::rank = rank;
  calculate_work();
  for (int i = 0; i <= n_procs; ++i)
    work_index[i] = simulation::work_index[i];
  return simulation::work_per_rank;
}
// OMC simulation.cpp:451-481 initialize_history (MC OpenMCStudy.C:326, CollisionKernel.C:75) for the index_source-th site of
// this rank's source bank in batch `batch` (gen_per_batch 1): the particle's id, its four stream seeds and what from_source
// copied (E, wgt, r, u)
void ref_initialize_history(int64_t n_particles, int n_procs, int rank, int64_t master_seed, int batch, int64_t index_source,
                            const double* site_r_u_E_wgt, int64_t* id, uint64_t* seeds4, double* state8)
{
  using namespace openmc;
  settings::run_mode = RunMode::EIGENVALUE;
  settings::n_particles = n_particles;
  settings::gen_per_batch = 1;
  settings::write_all_tracks = false;
  settings::track_identifiers.clear();
  mpi::n_procs = n_procs;
  mpi::rank = rank;
  calculate_work();
  openmc_set_seed(master_seed);
  simulation::total_gen = 0;
  simulation::current_batch = batch;
  simulation::current_gen = 1;
  simulation::source_bank.assign(index_source, SourceSite{});
  SourceSite& s = simulation::source_bank[index_source - 1];
  s.r = {site_r_u_E_wgt[0], site_r_u_E_wgt[1], site_r_u_E_wgt[2]};
  s.u = {site_r_u_E_wgt[3], site_r_u_E_wgt[4], site_r_u_E_wgt[5]};
  s.E = site_r_u_E_wgt[6];
  s.wgt = site_r_u_E_wgt[7];
  s.particle = ParticleType::neutron;
  Particle p;
  initialize_history(p, index_source);
  *id = p.id();
  for (int k = 0; k < N_STREAMS; ++k)
    seeds4[k] = p.seeds(k);
  state8[0] = p.r().x, state8[1] = p.r().y, state8[2] = p.r().z;
  state8[3] = p.u().x, state8[4] = p.u().y, state8[5] = p.u().z;
  state8[6] = p.E(), state8[7] = p.wgt();
}
// OMC tallies/tally.cpp:639-670 Tally::accumulate on a reference Tally object: `values` = n_batches x n_entries
// VALUE arrays scored in successive batches; returns SUM and SUM_SQ per entry and the number of realisations
int ref_tally_accumulate(const double* values, int n_batches, int n_entries, int64_t n_particles, double* sum, double* sum_sq)
{
  using namespace openmc;
  settings::n_particles = n_particles;
  settings::gen_per_batch = 1;
  settings::reduce_tallies = true;
  settings::run_mode = RunMode::EIGENVALUE;
  mpi::master = true;
  mpi::n_procs = 1;
  Tally t(987654);
  t.results_ = xt::zeros<double>({(size_t)n_entries, (size_t)1, (size_t)3});
  for (int b = 0; b < n_batches; ++b)
  {
    for (int i = 0; i < n_entries; ++i)
      t.results_(i, 0, static_cast<int>(TallyResult::VALUE)) = values[(size_t)b * n_entries + i];
    t.accumulate();
  }
  for (int i = 0; i < n_entries; ++i)
  {
    sum[i] = t.results_(i, 0, static_cast<int>(TallyResult::SUM));
    sum_sq[i] = t.results_(i, 0, static_cast<int>(TallyResult::SUM_SQ));
  }
  return t.n_realizations_;
}
// OMC eigenvalue.cpp:313-351 calculate_average_keff on the reference's own globals: the running mean and the
// standard deviation of the mean over active generations, one call per batch (gen_per_batch = 1)
void ref_keff_statistics(const double* k_generation, int n_batches, int n_inactive, double* keff, double* keff_std)
{
  using namespace openmc;
  settings::gen_per_batch = 1;
  settings::n_inactive = n_inactive;
  settings::confidence_intervals = false;
  simulation::total_gen = 0;
  simulation::k_generation.clear();
  simulation::k_sum = {0.0, 0.0};
  simulation::keff = 1.0;
  simulation::keff_std = 0.0;
  for (int b = 1; b <= n_batches; ++b)
  {
    simulation::current_batch = b;
    simulation::current_gen = 1;
    simulation::n_realizations = b > n_inactive ? b - n_inactive - 1 : 0; // finalize_batch increments it afterwards
    simulation::k_generation.push_back(k_generation[b - 1]);
    calculate_average_keff();
    keff[b - 1] = simulation::keff;
    keff_std[b - 1] = simulation::keff_std;
  }
}
// End-of-generation bookkeeping on the reference's own Particle and globals: every history's k scores go through
// Particle::event_death (OMC particle.cpp:319-352: atomic adds into global_tally_*, in history order here), then the
// global-tally update MaCaw does (MC OpenMCStudy.C:411-428 = OMC simulation.cpp:409-425), calculate_generation_keff
// (eigenvalue.cpp:49-69) and calculate_average_keff (:313-351).  k3 = [n_generations][n_histories][collision, absorption,
// tracklength]; outputs per generation: the three global-tally VALUEs, k_generation, keff, keff_std.
void ref_generation_bookkeeping(int64_t n_particles, int n_inactive, int n_generations, int64_t n_histories, const double* k3,
                                double* gt3, double* k_generation, double* keff, double* keff_std)
{
  using namespace openmc;
  settings::run_mode = RunMode::EIGENVALUE;
  settings::n_particles = n_particles;
  settings::gen_per_batch = 1;
  settings::n_inactive = n_inactive;
  settings::confidence_intervals = false;
  mpi::n_procs = 1;
  mpi::rank = 0;
  simulation::work_index = {0, n_particles};
  simulation::progeny_per_particle.assign(n_particles, 0);
  simulation::total_gen = 0;
  simulation::k_generation.clear();
  simulation::k_sum = {0.0, 0.0};
  simulation::keff = 1.0;
  simulation::keff_std = 0.0;
  simulation::keff_generation = 0.0;
  simulation::n_realizations = 0;
  settings::reduce_tallies = true;
  model::tallies.clear();
  simulation::global_tallies = xt::zeros<double>({N_GLOBAL_TALLIES, 3});
  global_tally_collision = global_tally_absorption = global_tally_tracklength = global_tally_leakage = 0.0;
  auto& gt = simulation::global_tallies;
  for (int b = 1; b <= n_generations; ++b)
  {
    simulation::current_batch = b;
    simulation::current_gen = 1;
    // initialize_generation (simulation.cpp:401-403; MC OpenMCStudy.C:305) remembers the starting track-length VALUE; not
    // called itself because its ufs branch would drag mesh.cpp into the link
    simulation::keff_generation = gt(GlobalTally::K_TRACKLENGTH, TallyResult::VALUE);
    const double* k = k3 + (size_t)(b - 1) * n_histories * 3;
    Particle p;
    for (int64_t i = 0; i < n_histories; ++i)
    {
      p.id() = i + 1;
      p.n_progeny() = 0;
      p.keff_tally_collision() = k[3 * i];
      p.keff_tally_absorption() = k[3 * i + 1];
      p.keff_tally_tracklength() = k[3 * i + 2];
      p.event_death();
    }
    gt(GlobalTally::K_COLLISION, TallyResult::VALUE) += global_tally_collision;
    gt(GlobalTally::K_ABSORPTION, TallyResult::VALUE) += global_tally_absorption;
    gt(GlobalTally::K_TRACKLENGTH, TallyResult::VALUE) += global_tally_tracklength;
    global_tally_collision = global_tally_absorption = global_tally_tracklength = 0.0;
    gt3[3 * (b - 1)] = gt(GlobalTally::K_COLLISION, TallyResult::VALUE);
    gt3[3 * (b - 1) + 1] = gt(GlobalTally::K_ABSORPTION, TallyResult::VALUE);
    gt3[3 * (b - 1) + 2] = gt(GlobalTally::K_TRACKLENGTH, TallyResult::VALUE);
    calculate_generation_keff();
    calculate_average_keff();
    k_generation[b - 1] = simulation::k_generation.back();
    keff[b - 1] = simulation::keff;
    keff_std[b - 1] = simulation::keff_std;
    // finalize_batch (simulation.cpp:334-345): accumulate_tallies moves VALUE / total_weight into SUM, SUM_SQ and counts
    // the realisation; inactive batches are wiped
    simulation::total_weight = (double)n_histories;
    accumulate_tallies();
    if (b <= n_inactive)
    {
      xt::view(gt, xt::all()) = 0.0;
      simulation::n_realizations = 0;
    }
  }
}
// OMC physics.cpp:955-1004, linked out of the reference's physics.cpp (see oracle/fake_hdf5/hdf5.h)
void ref_sample_cxs_target_velocity(double awr, double E, const double* u, double kT, uint64_t* seed, double* v)
{
  const openmc::Direction d = openmc::sample_cxs_target_velocity(awr, E, {u[0], u[1], u[2]}, kT, seed);
  v[0] = d.x, v[1] = d.y, v[2] = d.z;
}
// OMC search.h:14-20, the primitive of Nuclide::calculate_xs' energy-grid search (nuclide.cpp:596-611)
int64_t ref_lower_bound_index(const double* first, int64_t n, double value)
{
  return (int64_t)openmc::lower_bound_index(first, first + n, value);
}
// OMC distribution_spatial.cpp SpatialBox (the <space type="box"> of settings.xml), built the way
// OpenMC builds it: from an XML node
void ref_box_sample(const double* box6, uint64_t* seed, double* xyz)
{
  char xml[512];
  std::snprintf(xml, sizeof(xml), "<space type=\"box\"><parameters>%.17g %.17g %.17g %.17g %.17g %.17g</parameters></space>", box6[0],
                box6[1], box6[2], box6[3], box6[4], box6[5]);
  pugi::xml_document doc;
  doc.load_string(xml);
  const openmc::SpatialBox box(doc.child("space"));
  const openmc::Position r = box.sample(seed);
  xyz[0] = r.x, xyz[1] = r.y, xyz[2] = r.z;
}
// OMC distribution_multi.cpp:73-79 (used by isotropic-in-lab scattering, physics.cpp:715, and the source)
void ref_isotropic_direction(double* uvw, uint64_t* seed)
{
  const openmc::Direction u = openmc::isotropic_direction(seed);
  uvw[0] = u.x, uvw[1] = u.y, uvw[2] = u.z;
}
double ref_prn(uint64_t* seed) { return openmc::prn(seed); }
double ref_future_prn(int64_t n, uint64_t seed) { return openmc::future_prn(n, seed); }
uint64_t ref_future_seed(uint64_t n, uint64_t seed) { return openmc::future_seed(n, seed); }
uint64_t ref_init_seed(int64_t id, int offset) { return openmc::init_seed(id, offset); }
void ref_init_particle_seeds(int64_t id, uint64_t* seeds) { openmc::init_particle_seeds(id, seeds); }
void ref_set_master_seed(int64_t s) { openmc::openmc_set_seed(s); }
double ref_maxwell_spectrum(double T, uint64_t* seed) { return openmc::maxwell_spectrum(T, seed); }
double ref_watt_spectrum(double a, double b, uint64_t* seed) { return openmc::watt_spectrum(a, b, seed); }
double ref_uniform(double a, double b, uint64_t* seed) { return openmc::uniform_distribution(a, b, seed); }
void ref_rotate_angle(double* uvw, double mu, const double* phi, uint64_t* seed)
{
  openmc::rotate_angle_c(uvw, mu, phi, seed);
}
}
