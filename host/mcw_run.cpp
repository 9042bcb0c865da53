// mcw_run -- stand-alone C++ driver of the transport study over the C ABI: what `macaw-opt -i run.i`
// does for the hot path (Transient executioner, one OpenMCStudy::execute per time step;
// MC OpenMCStudy.C:366-402), on a problem file instead of a MOOSE input + Exodus mesh.
//
//   mcw_run problem.bin [--batches N] [--device D] [--nccl-id FILE] [--quiet]
//
// One process per GPU: with WORLD_SIZE > 1 in the environment (RANK, LOCAL_RANK as torchrun sets
// them) the particles are sharded and banks / tallies go through NCCL (host/nccl_comm.hpp).
// Rank 0 prints one JSON line per batch and a final line with the tally read-back through
// OpenMCTallyAux.  Doubles are printed with %.17g so a test can compare them bit for bit.
#include <cstdlib>
#include <cstring>
#include <memory>

#include "macaw_host.hpp"
#include "problem_io.hpp"
#ifdef MCW_WITH_NCCL
#include "nccl_comm.hpp"
#endif

static int env_int(const char* name, int dflt)
{
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int main(int argc, char** argv)
{
  if (argc < 2)
  {
    std::fprintf(stderr, "usage: mcw_run problem.bin [--batches N] [--device D] [--nccl-id FILE] [--quiet]\n");
    return 2;
  }
  int batches = 5, device = -1;
  bool quiet = false;
  std::string id_file = "/tmp/mcw_nccl_id";
  for (int a = 2; a < argc; ++a)
  {
    if (!std::strcmp(argv[a], "--batches") && a + 1 < argc)
      batches = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--device") && a + 1 < argc)
      device = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--nccl-id") && a + 1 < argc)
      id_file = argv[++a];
    else if (!std::strcmp(argv[a], "--quiet"))
      quiet = true;
  }
  const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1);
  if (device < 0)
    device = env_int("LOCAL_RANK", 0);
  try
  {
    using namespace macaw;
    const Problem P = ProblemFile(argv[1]).problem();
    std::unique_ptr<Comm> comm;
    if (nranks > 1)
    {
#ifdef MCW_WITH_NCCL
      comm = std::make_unique<NcclComm>(rank, nranks, device, id_file);
#else
      mooseError("mcw_run was built without NCCL (make NCCL=1)");
#endif
    }
    // the input file's [RayKernels], [UserObjects] and [AuxKernels] blocks
    CollisionKernel kernel(P.mesh, P.blocks, P.materials, P.temperature,
                           P.mesh.dimension() == 2 ? std::optional<double>(P.settings.z_coord) : std::nullopt);
    static const char* score_names[MCW_N_SCORES] = {"flux", "total", "fission", "nu-fission", "kappa-fission"};
    std::vector<OpenMCTally> tallies;
    for (const auto& t : P.tallies)
    {
      std::vector<std::string> scores, filters;
      for (int32_t s : t.scores)
        scores.push_back(score_names[s]);
      if (t.granularity == MCW_FILTER_CELL)
        filters.push_back("cell");
      else if (t.granularity == MCW_FILTER_UNIVERSE)
        filters.push_back("universe");
      if (!t.energy_bins.empty())
        filters.push_back("energy");
      tallies.emplace_back(t.id, scores, filters, t.estimator == MCW_ESTIMATOR_TRACKLENGTH ? "TRACKLENGTH" : "COLLISION",
                           "neutron", t.energy_bins);
    }
    OpenMCStudy study(P.mesh, P.library, kernel, P.settings, tallies, device, comm.get(), /*verbose=*/false);
    for (int b = 1; b <= batches; ++b)
    {
      const mcw_generation_result r = study.execute();
      if (rank == 0 && !quiet)
        std::printf("{\"batch\": %d, \"k_generation\": %.17g, \"keff\": %.17g, \"keff_std\": %.17g, \"k_collision\": %.17g, "
                    "\"k_absorption\": %.17g, \"k_tracklength\": %.17g, \"leakage\": %.17g, \"n_segments\": %lld, "
                    "\"n_collisions\": %lld, \"n_bank_sites\": %lld, \"ms\": %.3f}\n",
                    b, r.k_generation, r.keff, r.keff_std, r.k_collision, r.k_absorption, r.k_tracklength, r.leakage,
                    (long long)r.n_segments, (long long)r.n_collisions, (long long)r.n_bank_sites, r.ms_total);
    }
    if (rank == 0)
    {
      std::printf("{\"tallies\": [");
      for (size_t k = 0; k < tallies.size(); ++k)
      {
        const OpenMCTally& t = tallies[k];
        if (t.granularity() == MCW_FILTER_GLOBAL)
          continue;
        OpenMCTallyAux aux(study, t, t.granularity() == MCW_FILTER_CELL ? "cell" : "universe", t.scoreNames()[0]);
        aux.timestepSetup();
        double total = 0.0, first = aux.nSpatial() ? aux.computeValue(0) : 0.0;
        for (int64_t e = 0; e < aux.nSpatial(); ++e)
          total += aux.computeValue(e);
        std::printf("%s{\"id\": %d, \"score\": \"%s\", \"n_spatial\": %lld, \"sum\": %.17g, \"first\": %.17g}", k ? ", " : "", t.id(),
                    t.scoreNames()[0].c_str(), (long long)aux.nSpatial(), total, first);
      }
      std::printf("]}\n");
    }
  }
  catch (const macaw::Error& e)
  {
    std::fprintf(stderr, "*** ERROR *** (%d) %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
