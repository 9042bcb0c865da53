// mcw_run -- stand-alone C++ driver of the transport study over the C ABI: what `macaw-opt -i run.i`
// does for the hot path (Transient executioner, one OpenMCStudy::execute per time step;
// MC OpenMCStudy.C:366-402), on a problem file instead of a MOOSE input + Exodus mesh.
//
//   mcw_run problem.bin [--batches N] [--device D] [--nccl-id FILE] [--quiet] [--digest]
//   mcw_run --input run.i --tables lib.bin [--xml-dir DIR] [...]     the reference's input-file surface:
//            a MOOSE input (Mesh / RayKernels / RayBCs / UserObjects / AuxKernels / Executioner blocks) next
//            to OpenMC's settings.xml and materials.xml (host/input_file.hpp)
//
// One process per GPU: with WORLD_SIZE > 1 in the environment (RANK, LOCAL_RANK as torchrun sets
// them) the particles are sharded and banks / tallies go through NCCL (host/nccl_comm.hpp).
// Rank 0 prints one JSON line per batch and a final line with the tally read-back through
// OpenMCTallyAux.  Doubles are printed with %.17g so a test can compare them bit for bit.
#include <cstdlib>
#include <cstring>
#include <memory>

#include "input_file.hpp"
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
    std::fprintf(stderr, "usage: mcw_run problem.bin | --input run.i --tables lib.bin [--xml-dir DIR]\n"
                         "               [--batches N] [--device D] [--nccl-id FILE] [--quiet]\n");
    return 2;
  }
  int batches = -1, device = -1;
  bool quiet = false, check = false; // --check: parse the inputs, print what was understood, touch no GPU
  bool digest = false;               // --digest: validate the library and print mcw_digest of its descriptors, touch no GPU
  std::string id_file = "/tmp/mcw_nccl_id", problem_path, input_path, xml_dir, tables_path;
  for (int a = 1; a < argc; ++a)
  {
    if (!std::strcmp(argv[a], "--input") && a + 1 < argc)
      input_path = argv[++a];
    else if (!std::strcmp(argv[a], "--xml-dir") && a + 1 < argc)
      xml_dir = argv[++a];
    else if (!std::strcmp(argv[a], "--tables") && a + 1 < argc)
      tables_path = argv[++a];
    else if (argv[a][0] != '-')
      problem_path = argv[a];
    else if (!std::strcmp(argv[a], "--batches") && a + 1 < argc)
      batches = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--device") && a + 1 < argc)
      device = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--nccl-id") && a + 1 < argc)
      id_file = argv[++a];
    else if (!std::strcmp(argv[a], "--quiet"))
      quiet = true;
    else if (!std::strcmp(argv[a], "--check"))
      check = true;
    else if (!std::strcmp(argv[a], "--digest"))
      digest = true;
  }
  const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1);
  if (device < 0)
    device = env_int("LOCAL_RANK", 0);
  try
  {
    using namespace macaw;
    InputProblem in;
    if (!input_path.empty())
    {
      if (tables_path.empty())
        mooseError("--input needs --tables (nuclide table file: python -m macaw_b200.problem_io --library FILE)");
      if (xml_dir.empty())
      {
        const size_t slash = input_path.find_last_of('/');
        xml_dir = slash == std::string::npos ? "." : input_path.substr(0, slash);
      }
      in = readInput(input_path, xml_dir, tables_path);
      if (batches < 0)
        batches = in.batches;
    }
    else if (problem_path.empty())
      mooseError("no problem file and no --input");
    if (batches < 0)
      batches = 5;
    const Problem P = input_path.empty() ? ProblemFile(problem_path).problem() : in.problem;
    if (digest)
    {
      // what this host would hand to the library, as the library's own digest of it (host code only): equal to the digest
      // of the Python host's descriptors when both read the same problem (tests/test_cpp_host.py)
      const auto nd = P.library.nuclideDescs();
      const auto sd = P.library.temperatureSetDescs();
      std::vector<mcw_thermal_desc> td;
      for (const auto& t : P.library.thermal)
        td.push_back(t.desc());
      std::vector<mcw_urr_desc> ud;
      for (const auto& u : P.library.urr)
        ud.push_back(u.desc());
      char msg[512] = "";
      if (mcw_validate_library(nd.data(), (int)nd.size(), msg, sizeof(msg)))
        mooseError(std::string("library refused: ") + msg);
      for (const auto& d : sd)
        if (mcw_validate_library(&d, 1, msg, sizeof(msg)))
          mooseError(std::string("temperature set refused: ") + msg);
      uint64_t h[4] = {0, 0, 0, 0};
      mcw_digest(0, nd.data(), (int)nd.size(), &h[0]);
      mcw_digest(0, sd.data(), (int)sd.size(), &h[1]);
      mcw_digest(1, td.data(), (int)td.size(), &h[2]);
      mcw_digest(2, ud.data(), (int)ud.size(), &h[3]);
      std::printf("{\"nuclides\": \"%016llx\", \"temperature_sets\": \"%016llx\", \"thermal\": \"%016llx\", \"urr\": \"%016llx\"}\n",
                  (unsigned long long)h[0], (unsigned long long)h[1], (unsigned long long)h[2], (unsigned long long)h[3]);
      return 0;
    }
    if (check)
    {
      std::printf("{\"elem_type\": %d, \"n_elem\": %lld, \"n_nodes\": %lld, \"batches\": %d, \"n_particles\": %lld, \"n_inactive\": %d, "
                  "\"seed\": %lld, \"temperature\": %.17g, \"source_box\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"only_fissionable\": %d,\n",
                  P.mesh.elem_type, (long long)P.mesh.nElem(), (long long)P.mesh.nNodes(), batches, (long long)P.settings.n_particles,
                  P.settings.n_inactive, (long long)P.settings.master_seed, P.temperature[0], P.settings.source_box[0],
                  P.settings.source_box[1], P.settings.source_box[2], P.settings.source_box[3], P.settings.source_box[4],
                  P.settings.source_box[5], P.settings.source_only_fissionable);
      auto list = [](const char* name, const auto& v, const char* fmt, bool last = false) {
        std::printf(" \"%s\": [", name);
        for (size_t k = 0; k < v.size(); ++k)
        {
          std::printf(k ? ", " : "");
          std::printf(fmt, v[k]);
        }
        std::printf(last ? "]" : "],\n");
      };
      list("block_id", P.mesh.block_id, "%d");
      list("bnd_kind", P.mesh.bnd_kind, "%d");
      list("blocks", P.blocks, "%u");
      list("materials", P.materials, "%u");
      list("xyz", P.mesh.xyz, "%.17g");
      list("conn", P.mesh.conn, "%ld");
      list("side_bnd_id", P.mesh.side_bnd_id, "%d");
      std::printf(" \"library_materials\": [");
      for (size_t m = 0; m < P.library.materials.size(); ++m)
      {
        const Material& M = P.library.materials[m];
        std::printf("%s{", m ? ", " : "");
        list("nuclide", M.nuclide, "%d");
        list("atom_density", M.atom_density, "%.17g");
        list("isotropic", M.isotropic, "%d", true);
        std::printf("}");
      }
      std::printf("],\n \"tallies\": %zu, \"aux\": %zu, \"n_library_nuclides\": %zu}\n", P.tallies.size(), in.aux.size(),
                  P.library.nuclides.size());
      return 0;
    }
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
    for (size_t k = 0; k < in.tally_scores.size(); ++k) // --input: the strings as written, validated by OpenMCTally
      tallies.emplace_back(P.tallies[k].id, in.tally_scores[k], in.tally_filters[k], in.tally_estimators[k], "neutron",
                           P.tallies[k].energy_bins, in.tally_nuclides[k]);
    for (const auto& t : P.tallies)
    {
      if (!in.tally_scores.empty())
        break;
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
      if (!t.nuclides.empty())
        tallies.back().setNuclideBins(t.nuclides);
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
    if (rank == 0 && !in.aux.empty())
    {
      // the input's [AuxKernels]: one value per element (cell) or block (universe), like the AuxVariable MOOSE would fill
      std::printf("{\"aux\": [");
      for (size_t k = 0; k < in.aux.size(); ++k)
      {
        const auto& a = in.aux[k];
        const OpenMCTally* t = nullptr;
        for (const auto& c : tallies)
          if (c.id() == a.tally_id)
            t = &c;
        if (!t)
          mooseError("OpenMCTallyAux " + a.name + ": no OpenMCTally with id " + std::to_string(a.tally_id));
        OpenMCTallyAux aux(study, *t, a.granularity, a.score, a.energy_bin >= 0 ? std::optional<int>(a.energy_bin) : std::nullopt,
                           a.nuclide, &P.library);
        aux.timestepSetup();
        std::printf("%s{\"name\": \"%s\", \"tally_id\": %d, \"score\": \"%s\", \"values\": [", k ? ", " : "", a.name.c_str(),
                    a.tally_id, a.score.c_str());
        for (int64_t e = 0; e < aux.nSpatial(); ++e)
          std::printf("%s%.17g", e ? ", " : "", aux.computeValue(e));
        std::printf("]}");
      }
      std::printf("]}\n");
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
  catch (const std::exception& e) // e.g. a malformed number in an input file
  {
    std::fprintf(stderr, "*** ERROR *** %s\n", e.what());
    return 1;
  }
  return 0;
}
