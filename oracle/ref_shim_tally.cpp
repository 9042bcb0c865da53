// ref_shim_tally.cpp -- extern "C" doorway into the REFERENCE's own collision-estimator scoring:
// openmc/src/tallies/tally_scoring.cpp score_collision_tally (:2354) -> score_general_ce (:516), with the reference's
// CellFilter / UniverseFilter / EnergyFilter (tallies/filter_*.cpp) doing the binning and the reference's Tally holding the
// results.  MaCaw reaches it through Particle::event_collide (src/raykernels/CollisionKernel.C:185 -> openmc/src/particle.cpp:247).  Linked into
// oracle/_ref/libomc_ref.so by oracle/Makefile; test infrastructure only.  Nothing here restates an algorithm: the shim builds
// the objects (n_cells empty cell slots, one tally) and replays (cell, universe, E_last, wgt_last, macro xs) events through it.
#include <cstdint>
#include <memory>
#include "openmc/cell.h"
#include "openmc/geometry.h"
#include "openmc/particle.h"
#include "openmc/settings.h"
#include "openmc/tallies/filter.h"
#include "openmc/tallies/filter_cell.h"
#include "openmc/tallies/filter_energy.h"
#include "openmc/tallies/filter_universe.h"
#include "openmc/tallies/tally.h"
#include "openmc/tallies/tally_scoring.h"
// distance == nullptr: collision estimator (E, wgt are the pre-collision E_last, wgt_last); otherwise track-length estimator
// (E, wgt are the particle's current values, distance[k] the segment length).
static int score_events(int n_cells, int n_universes, const double* ebins, int n_edges, int use_universe,
                        const char* const* scores, int n_scores, int64_t n_events, const int32_t* cell, const int32_t* universe,
                        const double* E_last, const double* wgt_last, const double* distance, const double* macro4,
                        double* value_out)
{
  using namespace openmc;
  model::n_coord_levels = 1;
  settings::run_CE = true;
  model::cells.clear();
  model::cells.resize(n_cells);
  model::universes.clear();
  model::universes.resize(n_universes);
  model::tally_filters.clear();
  model::filter_map.clear();
  model::tallies.clear();
  model::tally_map.clear();
  std::vector<Filter*> filters;
  if (use_universe) {
    auto* f = Filter::create<UniverseFilter>();
    std::vector<int32_t> ids(n_universes);
    for (int i = 0; i < n_universes; ++i) ids[i] = i;
    f->set_universes(ids);
    filters.push_back(f);
  } else {
    auto* f = Filter::create<CellFilter>();
    std::vector<int32_t> ids(n_cells);
    for (int i = 0; i < n_cells; ++i) ids[i] = i;
    f->set_cells(ids);
    filters.push_back(f);
  }
  if (n_edges > 0) {
    auto* f = Filter::create<EnergyFilter>();
    f->set_bins({ebins, (size_t)n_edges});
    filters.push_back(f);
  }
  Tally* t = Tally::create(1);
  t->set_filters(filters);
  std::vector<std::string> sc;
  for (int i = 0; i < n_scores; ++i) sc.push_back(scores[i]);
  t->set_scores(sc);
  t->estimator_ = distance ? TallyEstimator::TRACKLENGTH : TallyEstimator::COLLISION;
  t->set_active(true);
  t->init_results();
  t->reset();
  setup_active_tallies();
  Particle p;
  p.type() = ParticleType::neutron;
  p.n_coord() = 1;
  for (int64_t k = 0; k < n_events; ++k) {
    p.coord(0).cell = cell[k];
    p.coord(0).universe = universe[k];
    if (distance)
      p.E() = p.E_last() = E_last[k], p.wgt() = wgt_last[k]; // event_calculate_xs sets E_last = E before the advance (particle.cpp:93)
    else
      p.E_last() = E_last[k], p.wgt_last() = wgt_last[k];
    p.macro_xs().total = macro4[4 * k];
    p.macro_xs().absorption = macro4[4 * k + 1];
    p.macro_xs().fission = macro4[4 * k + 2];
    p.macro_xs().nu_fission = macro4[4 * k + 3];
    if (distance)
      score_tracklength_tally(p, distance[k]);
    else
      score_collision_tally(p);
  }
  const auto& r = t->results_;
  size_t o = 0;
  for (size_t i = 0; i < r.shape()[0]; ++i)
    for (size_t j = 0; j < r.shape()[1]; ++j)
      value_out[o++] = r(i, j, static_cast<int>(TallyResult::VALUE));
  return (int)(r.shape()[0] * r.shape()[1]);
}

extern "C" {
__attribute__((visibility("default"))) int ref_score_collision_events(
    int n_cells, int n_universes, const double* ebins, int n_edges, int use_universe, const char* const* scores, int n_scores,
    int64_t n_events, const int32_t* cell, const int32_t* universe, const double* E_last, const double* wgt_last,
    const double* macro4, double* value_out)
{
  return score_events(n_cells, n_universes, ebins, n_edges, use_universe, scores, n_scores, n_events, cell, universe, E_last,
                      wgt_last, nullptr, macro4, value_out);
}
// tally_scoring.cpp score_tracklength_tally (:2289)
__attribute__((visibility("default"))) int ref_score_tracklength_events(
    int n_cells, int n_universes, const double* ebins, int n_edges, int use_universe, const char* const* scores, int n_scores,
    int64_t n_events, const int32_t* cell, const int32_t* universe, const double* E, const double* wgt, const double* distance,
    const double* macro4, double* value_out)
{
  return score_events(n_cells, n_universes, ebins, n_edges, use_universe, scores, n_scores, n_events, cell, universe, E, wgt,
                      distance, macro4, value_out);
}
}
