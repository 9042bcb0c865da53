// macaw_host.hpp -- C++ host layer above the C ABI (include/macaw_b200.h).
//
// Mirrors, for the particle-history hot path only, the four MOOSE objects through which the
// reference exposes that path (same parameter names, same error strings, same call order):
//
//   macaw::OpenMCStudy     <- MC src/userobjects/OpenMCStudy.C:48-76 validParams, :78-204 ctor, :366-402 execute
//   macaw::CollisionKernel <- MC src/raykernels/CollisionKernel.C:24-70 (configuration carrier in GPU mode)
//   macaw::OpenMCTally     <- MC src/userobjects/OpenMCTally.C:30-101 (filters / scores / estimator)
//   macaw::OpenMCTallyAux  <- MC src/auxkernels/OpenMCTallyAux.C:26-215 (per-element read-back)
//
// MOOSE and libMesh are not in this image, so these are plain classes instead of
// registerMooseObject'ed ones; INTEGRATION.md shows the thin MOOSE adapter that forwards
// InputParameters into them.  Everything below is host plumbing: no physics, no geometry, and no CPU
// fallback -- every number comes out of libmacaw_b200.so (CUDA, sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/macaw_b200.h"

namespace macaw
{

// mooseError / paramError analogue: the reference aborts through these; here they throw.
struct Error : std::runtime_error
{
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void mooseError(const std::string& m) { throw Error(MCW_E_INVALID_ARGUMENT, m); }
inline void paramError(const std::string& param, const std::string& m) { throw Error(MCW_E_INVALID_ARGUMENT, param + ": " + m); }

// What the study reads from libMesh on the path (RT TraceRayTools.C:404-407, TraceRay.C:1433,
// MC CollisionKernel.C:110): node coordinates, connectivity in libMesh node order, subdomain ids,
// neighbours (optional), boundary ids per side and the RayBC kind on every boundary id.
struct Mesh
{
  int elem_type = MCW_HEX8;
  std::vector<double> xyz;          // n_nodes x 3
  std::vector<int64_t> conn;        // n_elem x nodes_per_elem
  std::vector<int32_t> block_id;    // n_elem
  std::vector<int64_t> neighbors;   // n_elem x n_sides or empty (library matches faces)
  std::vector<int32_t> side_bnd_id; // n_elem x n_sides, -1 = interior
  std::vector<int32_t> bnd_kind;    // per boundary id: MCW_BC_KILL / MCW_BC_REFLECT
  int nodesPerElem() const { return elem_type == MCW_HEX8 ? 8 : 4; }
  int sidesPerElem() const { return elem_type == MCW_HEX8 ? 6 : 4; }
  int dimension() const { return elem_type == MCW_QUAD4 ? 2 : 3; }
  int64_t nElem() const { return (int64_t)block_id.size(); }
  int64_t nNodes() const { return (int64_t)xyz.size() / 3; }
};

// Distributions on a grid of incoming energies (mcw_tabular2d): AngleDistribution / ContinuousTabular data
struct Tabular2D
{
  std::vector<double> energy, x, pdf, cdf;
  std::vector<int32_t> offset, interpolation, n_discrete;
  int histogram_in_energy = 0;
  mcw_tabular2d desc() const
  {
    mcw_tabular2d t{};
    t.n_energy = (int32_t)energy.size();
    if (t.n_energy == 0)
      return t;
    t.histogram_in_energy = histogram_in_energy;
    t.energy = energy.data(), t.offset = offset.data(), t.interpolation = interpolation.data();
    t.n_discrete = n_discrete.empty() ? nullptr : n_discrete.data();
    t.x = x.data(), t.pdf = pdf.data(), t.cdf = cdf.data();
    return t;
  }
};
// mcw_function1d: a Polynomial or a one-region Tabulated1D (OMC endf.h Function1D)
struct Function1D
{
  int kind = MCW_FN_NONE, interpolation = 2;
  std::vector<double> x, y;
  mcw_function1d desc() const
  {
    mcw_function1d f{};
    f.kind = kind, f.n = (int32_t)y.size(), f.interpolation = interpolation;
    f.x = x.empty() ? nullptr : x.data(), f.y = y.data();
    return f;
  }
};
// One neutron product of the fission reaction (mcw_fission_product): prompt, or a delayed group
struct FissionProduct
{
  Function1D yield, a, b;
  int law = MCW_LAW_WATT;
  double u = 0;
  Tabular2D energy;
  mcw_fission_product desc() const
  {
    mcw_fission_product f{};
    f.yield = yield.desc(), f.law = law, f.u = u, f.a = a.desc(), f.b = b.desc(), f.energy = energy.desc();
    return f;
  }
};
// A scattering reaction after a nuclide's first (mcw_reaction_desc; OMC Nuclide::index_inelastic_scatter_ order)
struct Reaction
{
  int mt = 0, threshold = 0, law = MCW_LAW_LEVEL, in_cm = 1, nbody_n = 0;
  double threshold_energy = 0, mass_ratio = 0, yield = 1, nbody_mass_ratio = 0, nbody_q = 0;
  std::vector<double> xs;
  Tabular2D angle, energy;
  std::vector<int32_t> corr_offset, corr_interp;
  std::vector<double> corr_mu, corr_pdf, corr_cdf, km_r, km_a;
  Function1D yield_fn, a, b; // energy-dependent multiplicity; parameters of the Maxwell / evaporation / Watt laws
  double u = 0;
  mcw_reaction_desc desc() const
  {
    mcw_reaction_desc r{};
    r.yield_fn = yield_fn.desc(), r.u = u, r.a = a.desc(), r.b = b.desc();
    r.mt = mt, r.threshold = threshold, r.xs = xs.data();
    r.threshold_energy = threshold_energy, r.mass_ratio = mass_ratio, r.yield = yield;
    r.law = law, r.in_cm = in_cm;
    r.angle = angle.desc(), r.energy = energy.desc();
    r.corr_offset = corr_offset.data(), r.corr_interp = corr_interp.data();
    r.corr_mu = corr_mu.data(), r.corr_pdf = corr_pdf.data(), r.corr_cdf = corr_cdf.data();
    r.km_r = km_r.data(), r.km_a = km_a.data();
    r.nbody_n = nbody_n, r.nbody_mass_ratio = nbody_mass_ratio, r.nbody_q = nbody_q;
    return r;
  }
};
struct Nuclide
{
  std::string name; // as in materials.xml (OMC data::nuclide_map); may be empty for generated libraries
  std::vector<double> energy, xs, elastic, inelastic;
  Tabular2D elastic_angle, inelastic_angle, inelastic_energy; // optional tabulated secondary distributions
  int inelastic_in_cm = 1;
  // -1: level, or continuous tabular when inelastic_energy is set; else MCW_LAW_CORRELATED / KALBACH_MANN / NBODY with
  // the angular distribution per point of inelastic_energy (correlated), r / a per point (Kalbach-Mann), or
  // d.nbody_n / nbody_mass_ratio / nbody_q (N-body)
  int inelastic_law = -1;
  std::vector<int32_t> corr_offset, corr_interp;
  std::vector<double> corr_mu, corr_pdf, corr_cdf, km_r, km_a;
  double inelastic_yield = 1.0;
  Function1D inelastic_yield_fn, inelastic_a, inelastic_b; // the inline reaction's Reaction::yield_fn, a, b, u
  double inelastic_u = 0;
  std::vector<Reaction> more_reactions;
  mutable std::vector<mcw_reaction_desc> more_descs; // filled by Library::nuclideDescs(), which hands out pointers into it
  std::vector<FissionProduct> fission_products;      // fission neutrons as a nuclide file has them (empty: the scalars in d)
  Function1D total_nu;
  mutable std::vector<mcw_fission_product> fission_descs;
  mcw_nuclide_desc d{}; // scalar fields; the pointers are filled by Library::descs()
};
struct Material
{
  std::vector<int32_t> nuclide;
  std::vector<double> atom_density;
  std::vector<int32_t> isotropic; // <isotropic> list of materials.xml as per-nuclide flags; empty = none
  // <sab> entries (OMC Material::thermal_tables_): position in `nuclide`, table index, fraction
  std::vector<int32_t> sab_position, sab_table;
  std::vector<double> sab_fraction;
};
// One temperature of an S(alpha, beta) table (mcw_thermal_desc; OMC ThermalData)
struct ThermalTable
{
  double energy_max = 0;
  int inelastic_type = 0, elastic_type = 0, n_mu = 0, n_e_out = 0, skewed = 0, n_elastic_mu = 0;
  std::vector<double> i_xs_energy, i_xs, i_energy, i_e_out, i_pdf, i_cdf, i_mu, e_x, e_y, e_mu;
  std::vector<int32_t> i_offset;
  mcw_thermal_desc desc() const
  {
    mcw_thermal_desc t{};
    t.energy_max = energy_max, t.inelastic_type = inelastic_type, t.elastic_type = elastic_type;
    t.n_inelastic_xs = (int32_t)i_xs_energy.size(), t.n_inelastic_energy = (int32_t)i_energy.size();
    t.inelastic_xs_energy = i_xs_energy.data(), t.inelastic_xs = i_xs.data(), t.inelastic_energy = i_energy.data();
    t.inelastic_offset = i_offset.data(), t.inelastic_e_out = i_e_out.data(), t.inelastic_pdf = i_pdf.data();
    t.inelastic_cdf = i_cdf.data(), t.inelastic_mu = i_mu.data();
    t.n_mu = n_mu, t.n_e_out = n_e_out, t.skewed = skewed;
    t.n_elastic = (int32_t)e_x.size(), t.elastic_x = e_x.data(), t.elastic_y = e_y.data(), t.elastic_mu = e_mu.data();
    t.n_elastic_mu = n_elastic_mu;
    return t;
  }
};
// Probability tables of one nuclide (mcw_urr_desc; OMC UrrData)
struct UrrTable
{
  int nuclide = 0, interpolation = 2, inelastic_flag = -1, absorption_flag = -1, multiply_smooth = 0, n_band = 0;
  std::vector<double> energy, table;
  mcw_urr_desc desc() const
  {
    return {nuclide, interpolation, inelastic_flag, absorption_flag, multiply_smooth, (int32_t)energy.size(), n_band, 0,
            energy.data(), table.data()};
  }
};
struct Library
{
  std::vector<Nuclide> nuclides;
  // further data sets of the nuclides at other temperatures (OMC Nuclide::kTs_ / grid_[t] / xs_[t]): (nuclide index, set);
  // set k becomes library entry nuclides.size() + k (mcw_add_temperature_set), which an UrrTable of that set names
  std::vector<std::pair<int, Nuclide>> temperature_sets;
  std::vector<Material> materials;
  std::vector<ThermalTable> thermal; // data::thermal_scatt: every temperature of every table as one entry
  // several temperatures per table (mcw_set_thermal_temperatures): per entry the first entry of its table, and its kT;
  // empty = one temperature per table
  std::vector<int32_t> thermal_same;
  std::vector<double> thermal_kT;
  std::vector<UrrTable> urr;         // urr_data_ of the nuclides that have it
  std::vector<mcw_nuclide_desc> nuclideDescs() const
  {
    std::vector<const Nuclide*> all;
    for (const auto& n : nuclides)
      all.push_back(&n);
    return descsOf(all);
  }
  std::vector<mcw_nuclide_desc> temperatureSetDescs() const
  {
    std::vector<const Nuclide*> all;
    for (const auto& s : temperature_sets)
      all.push_back(&s.second);
    return descsOf(all);
  }
  static std::vector<mcw_nuclide_desc> descsOf(const std::vector<const Nuclide*>& list)
  {
    std::vector<mcw_nuclide_desc> out;
    static const double zero = 0.0;
    for (const Nuclide* np : list)
    {
      const Nuclide& n = *np;
      mcw_nuclide_desc d = n.d;
      d.n_grid = (int32_t)n.energy.size();
      d.energy = n.energy.data();
      d.xs = n.xs.data();
      d.elastic = n.elastic.data();
      d.n_inelastic = n.inelastic.empty() ? 0 : 1;
      d.inelastic = n.inelastic.empty() ? &zero : n.inelastic.data();
      d.elastic_angle = n.elastic_angle.desc();
      d.inelastic_angle = n.inelastic_angle.desc();
      d.inelastic_energy = n.inelastic_energy.desc();
      d.inelastic_law = n.inelastic_law >= 0 ? n.inelastic_law
                                             : (n.inelastic_energy.energy.empty() ? MCW_LAW_LEVEL : MCW_LAW_CONTINUOUS_TABULAR);
      d.inelastic_in_cm = n.inelastic_in_cm;
      d.corr_offset = n.corr_offset.data(), d.corr_interp = n.corr_interp.data();
      d.corr_mu = n.corr_mu.data(), d.corr_pdf = n.corr_pdf.data(), d.corr_cdf = n.corr_cdf.data();
      d.km_r = n.km_r.data(), d.km_a = n.km_a.data();
      d.inelastic_yield = n.inelastic_yield;
      d.inelastic_yield_fn = n.inelastic_yield_fn.desc(), d.inelastic_u = n.inelastic_u;
      d.inelastic_a = n.inelastic_a.desc(), d.inelastic_b = n.inelastic_b.desc();
      n.more_descs.clear();
      for (const auto& r : n.more_reactions)
        n.more_descs.push_back(r.desc());
      d.n_more_reactions = (int32_t)n.more_descs.size();
      d.more_reactions = n.more_descs.empty() ? nullptr : n.more_descs.data();
      n.fission_descs.clear();
      for (const auto& f : n.fission_products)
        n.fission_descs.push_back(f.desc());
      d.n_fission_products = (int32_t)n.fission_descs.size();
      d.fission_products = n.fission_descs.empty() ? nullptr : n.fission_descs.data();
      d.total_nu = n.total_nu.desc();
      out.push_back(d);
    }
    return out;
  }
  std::vector<mcw_material_desc> materialDescs() const
  {
    std::vector<mcw_material_desc> out;
    for (const auto& m : materials)
      out.push_back({(int32_t)m.nuclide.size(), 0, m.nuclide.data(), m.atom_density.data(),
                     m.isotropic.size() == m.nuclide.size() ? m.isotropic.data() : nullptr});
    return out;
  }
};

// RAII over mcw_handle; check() turns the C convention (int + mcw_last_error, cf. OMC capi.h:180-190
// + error.cpp:32) into the exception the MOOSE side would raise with mooseError.
class Handle
{
public:
  Handle(int device, int rank, int nranks)
  {
    if (int rc = mcw_create(&_h, device, rank, nranks))
      throw Error(rc, "mcw_create failed: a CUDA device of compute capability 10.x (B200) is required; there is no CPU fallback");
  }
  ~Handle()
  {
    if (_h)
      mcw_destroy(_h);
  }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  mcw_handle get() const { return _h; }
  void check(int rc) const
  {
    if (rc)
      throw Error(rc, mcw_last_error(_h));
  }

private:
  mcw_handle _h = nullptr;
};

// Parameters: temperature, blocks, materials, z_coord, verbose (MC CollisionKernel.C:24-42).
class CollisionKernel
{
public:
  CollisionKernel(const Mesh& mesh, std::vector<unsigned> blocks, std::vector<unsigned> materials,
                  std::vector<double> temperature, std::optional<double> z_coord = std::nullopt, bool verbose = false)
    : _blocks(std::move(blocks)), _materials(std::move(materials)), _temperature(std::move(temperature)),
      _is_2D(mesh.dimension() == 2), _z_coord(z_coord.value_or(0.0)), _verbose(verbose)
  {
    if (_is_2D && !z_coord)
      mooseError("The z_coord parameter must be specified for 2D calculations"); // :60-61
    if (_blocks.size() != _materials.size())
      paramError("blocks", "The blocks parameter must be the same size as the materials parameter."); // :66-67
    if (_temperature.size() != 1 && (int64_t)_temperature.size() != mesh.nElem())
      paramError("temperature", "Only CONST MONOMIAL temperatures are supported."); // :52-57: one value per element
  }
  std::vector<int32_t> blockIds() const { return {_blocks.begin(), _blocks.end()}; }
  // OpenMC material ids are 1-based; model::material_map resolves them to indices (:110)
  std::vector<int32_t> materialIndices() const
  {
    std::vector<int32_t> out;
    for (unsigned m : _materials)
      out.push_back((int32_t)m - 1);
    return out;
  }
  const std::vector<double>& temperature() const { return _temperature; }
  bool is2D() const { return _is_2D; }
  double zCoord() const { return _z_coord; }
  bool verbose() const { return _verbose; }

private:
  std::vector<unsigned> _blocks, _materials;
  std::vector<double> _temperature;
  bool _is_2D;
  double _z_coord;
  bool _verbose;
};

// Parameters: id, particle_type, estimator, scores, filters, nuclides, energy_bins (MC OpenMCTally.C:30-65).
class OpenMCTally
{
public:
  OpenMCTally(int id, std::vector<std::string> scores, std::vector<std::string> filters,
              const std::string& estimator = "COLLISION", const std::string& particle_type = "neutron",
              std::vector<double> energy_bins = {}, std::vector<std::string> nuclides = {"total"})
    : _id(id), _score_names(std::move(scores)), _filters(std::move(filters)), _energy_bins(std::move(energy_bins)),
      _nuclides(std::move(nuclides))
  {
    if (_nuclides.empty())
      _nuclides = {"total"};
    if (particle_type != "neutron")
      mooseError("Only neutrons are currently supported."); // :87-88
    if (estimator == "COLLISION")
      _estimator = MCW_ESTIMATOR_COLLISION;
    else if (estimator == "TRACKLENGTH")
      _estimator = MCW_ESTIMATOR_TRACKLENGTH; // extension: the reference rejects it (:89-90), north star names it
    else
      mooseError("Only collision estimator currently supported"); // :89-90
    bool has_energy = false;
    _granularity = MCW_FILTER_GLOBAL;
    for (const auto& f : _filters)
    {
      if (f == "energy")
        has_energy = true;
      else if (f == "cell")
        _granularity = MCW_FILTER_CELL;
      else if (f == "universe")
      {
        if (_granularity != MCW_FILTER_CELL)
          _granularity = MCW_FILTER_UNIVERSE;
      }
      else if (f != "particle")
        mooseError("Unrecognized filter"); // :173-174
    }
    if (has_energy && _energy_bins.empty())
      mooseError("The energy bins parameter should be supplied if an energy filter is requested.");
    if (!has_energy)
      _energy_bins.clear();
    static const char* names[MCW_N_SCORES] = {"flux", "total", "fission", "nu-fission", "kappa-fission"};
    for (const auto& s : _score_names)
    {
      int code = -1;
      for (int k = 0; k < MCW_N_SCORES; ++k)
        if (s == names[k])
          code = k;
      if (code < 0)
        paramError("scores", "score '" + s + "' is not on the GPU path");
      _scores.push_back(code);
    }
  }
  int id() const { return _id; }
  int estimator() const { return _estimator; }
  int granularity() const { return _granularity; }
  const std::vector<int32_t>& scores() const { return _scores; }
  const std::vector<std::string>& scoreNames() const { return _score_names; }
  const std::vector<double>& energyBins() const { return _energy_bins; }

  // MC OpenMCTally.C:188-204 against the study's library.  Fills what every nuclide bin of results_ scores (library index,
  // -1 = total, -2 = allocated but never scored) and the reference's own nuclides_ vector (what OpenMCTallyAux.C:93-106
  // inspects).  `all` appends every nuclide and -1 to the tally's default {-1} without clearing it, so results_ has
  // n + 2 nuclide bins: score_all_nuclides (OMC tally_scoring.cpp:2142-2175) fills bin i for nuclide i and bin n for the
  // total.  A list of names gets "total" appended (:198-202); unknown names fail like OMC tally.cpp:552-554.
  // nuclide bins given as library indices (problem files); resolveNuclides then leaves them alone
  void setNuclideBins(std::vector<int32_t> bins)
  {
    _bins = std::move(bins);
    _ref_nuclides = _bins;
    _bins_given = true;
  }
  void resolveNuclides(const Library& library) const
  {
    if (_bins_given)
      return;
    const int n = (int)library.nuclides.size();
    _bins.clear();
    _ref_nuclides.clear();
    if (_nuclides[0] == "all")
    {
      _ref_nuclides.push_back(-1);
      for (int i = 0; i < n; ++i)
        _bins.push_back(i), _ref_nuclides.push_back(i);
      _bins.push_back(-1), _bins.push_back(-2);
      _ref_nuclides.push_back(-1);
      return;
    }
    if (_nuclides[0] == "total")
    {
      _bins = _ref_nuclides = {-1};
      return;
    }
    for (const auto& name : _nuclides)
    {
      int found = name == "total" ? -1 : -3;
      for (int i = 0; i < n && found == -3; ++i)
        if (library.nuclides[i].name == name)
          found = i;
      if (found == -3)
        mooseError("Could not find the nuclide " + name + " specified in tally " + std::to_string(_id) + " in any material");
      _bins.push_back(found);
    }
    _bins.push_back(-1);
    _ref_nuclides = _bins;
  }
  const std::vector<int32_t>& nuclideBins() const { return _bins; }
  const std::vector<int32_t>& referenceNuclides() const { return _ref_nuclides; }

private:
  int _id, _estimator = MCW_ESTIMATOR_COLLISION, _granularity;
  std::vector<std::string> _score_names, _filters;
  std::vector<int32_t> _scores;
  std::vector<double> _energy_bins;
  std::vector<std::string> _nuclides;
  mutable std::vector<int32_t> _bins = {-1}, _ref_nuclides = {-1};
  bool _bins_given = false;
};

// Communication the study needs between transport and finalize when particles are sharded over
// several GPUs (MC OpenMCStudy.C:454-607 synchronizeBanks; OMC eigenvalue.cpp:71-311).  The
// single-rank study never calls it.  host/nccl_comm.hpp implements it over NCCL.
struct Comm
{
  virtual ~Comm() = default;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  virtual std::vector<int64_t> allGather(int64_t v) = 0;                     // Exscan + Bcast of counts
  virtual void allReduceSum(double* device_buf, int64_t n, void* stream) = 0; // tallies + k estimators
  // grouped send/recv of byte ranges between device buffers
  struct Range
  {
    int peer;
    int64_t offset, bytes;
  };
  virtual void exchange(const char* send_base, const std::vector<Range>& sends, char* recv_base,
                        const std::vector<Range>& recvs, void* stream) = 0;
};

// RayTracingStudy-derived transport study.  Construction = ctor + initialSetup of the reference
// (openmc_init, material / cell set-up, openmc_simulation_init: MC OpenMCStudy.C:78-204);
// execute() = one Transient time step = one batch = one generation (:366-402).
class OpenMCStudy
{
public:
  OpenMCStudy(const Mesh& mesh, const Library& library, const CollisionKernel& kernel, mcw_settings settings,
              const std::vector<OpenMCTally>& tallies = {}, int device = 0, Comm* comm = nullptr, bool verbose = true)
    : _comm(comm), _rank(comm ? comm->rank() : 0), _nranks(comm ? comm->size() : 1), _h(device, _rank, _nranks),
      _verbose(verbose), _settings(settings)
  {
    _settings.is_2d = kernel.is2D();
    if (kernel.is2D())
      _settings.z_coord = kernel.zCoord();
    const bool have_nbr = !mesh.neighbors.empty();
    _h.check(mcw_set_mesh(_h.get(), mesh.elem_type, mesh.nNodes(), mesh.xyz.data(), mesh.nElem(), mesh.conn.data(),
                          mesh.block_id.data(), have_nbr ? mesh.neighbors.data() : nullptr, mesh.side_bnd_id.data(),
                          (int)mesh.bnd_kind.size(), mesh.bnd_kind.data()));
    auto nd = library.nuclideDescs();
    _h.check(mcw_set_library(_h.get(), nd.data(), (int)nd.size()));
    {
      // further temperatures of the nuclides (OMC kTs_): the study's coupled temperature picks the nearest per element
      auto sd = library.temperatureSetDescs();
      for (size_t k = 0; k < sd.size(); ++k)
      {
        int entry = -1;
        _h.check(mcw_add_temperature_set(_h.get(), library.temperature_sets[k].first, &sd[k], &entry));
        if (entry != (int)(nd.size() + k))
          mooseError("temperature set became library entry " + std::to_string(entry) + ", expected " + std::to_string(nd.size() + k));
      }
    }
    auto md = library.materialDescs();
    auto bid = kernel.blockIds();
    auto bm = kernel.materialIndices();
    _h.check(mcw_set_materials(_h.get(), md.data(), (int)md.size(), bid.data(), bm.data(), (int)bid.size()));
    if (!library.urr.empty())
    {
      std::vector<mcw_urr_desc> ud;
      for (const auto& u : library.urr)
        ud.push_back(u.desc());
      _h.check(mcw_set_urr_tables(_h.get(), ud.data(), (int)ud.size()));
    }
    if (!library.thermal.empty())
    {
      std::vector<mcw_thermal_desc> td;
      for (const auto& t : library.thermal)
        td.push_back(t.desc());
      _h.check(mcw_set_thermal_tables(_h.get(), td.data(), (int)td.size()));
      if (!library.thermal_same.empty())
        _h.check(mcw_set_thermal_temperatures(_h.get(), library.thermal_same.data(), library.thermal_kT.data(), (int)td.size()));
      for (size_t m = 0; m < library.materials.size(); ++m)
      {
        const Material& M = library.materials[m];
        if (!M.sab_position.empty())
          _h.check(mcw_set_material_thermal(_h.get(), (int)m, (int)M.sab_position.size(), M.sab_position.data(),
                                            M.sab_table.data(), M.sab_fraction.data()));
      }
    }
    _h.check(mcw_set_temperature(_h.get(), kernel.temperature().data(), (int64_t)kernel.temperature().size()));
    _h.check(mcw_set_settings(_h.get(), &_settings));
    for (const auto& t : tallies)
    {
      _h.check(mcw_add_tally(_h.get(), t.id(), t.estimator(), t.granularity(), t.scores().data(), (int)t.scores().size(),
                             t.energyBins().empty() ? nullptr : t.energyBins().data(), (int)t.energyBins().size()));
      t.resolveNuclides(library);
      if (t.nuclideBins().size() != 1 || t.nuclideBins()[0] != -1)
        _h.check(mcw_set_tally_nuclides(_h.get(), t.id(), t.nuclideBins().data(), (int)t.nuclideBins().size()));
    }
    _work_index.resize(_nranks + 1);
    _h.check(mcw_calculate_work(_settings.n_particles, _nranks, _work_index.data())); // OMC simulation.cpp:529-551
    _h.check(mcw_prepare(_h.get()));
    _h.check(mcw_sample_initial_source(_h.get())); // openmc_simulation_init -> initialize_source (:147)
  }

  // The coupled temperature changed (a new time step of the multiphysics solve, MC CollisionKernel.C:109 reads it every
  // segment).  With several temperatures in the library the new field re-prepares the problem, which starts the run
  // over: the source bank of this rank is carried across; k statistics and tally sums restart with the new field.
  void setTemperature(const std::vector<double>& T)
  {
    int64_t n = 0;
    std::vector<mcw_source_site> bank((size_t)std::max<int64_t>(_settings.n_particles, 1));
    _h.check(mcw_get_source_bank(_h.get(), bank.data(), (int64_t)bank.size(), &n));
    _h.check(mcw_set_temperature(_h.get(), T.data(), (int64_t)T.size()));
    if (n > 0)
      _h.check(mcw_set_source_bank(_h.get(), bank.data(), n));
  }

  // One batch.  Returns the generation result; prints the reference's line when verbose (:399-401).
  mcw_generation_result execute()
  {
    ++_batch;
    mcw_generation_result res{};
    if (_nranks == 1)
      _h.check(mcw_run_generation(_h.get(), _batch, &res));
    else
    {
      _h.check(mcw_transport_generation(_h.get(), _batch)); // generateRays() + propagation
      synchronizeBanks();                                    // postExecuteStudy()
      _h.check(mcw_finalize_generation(_h.get(), _batch, &res));
    }
    _last = res;
    if (_verbose && _rank == 0)
      std::printf("Keff %.10f (%.10f) Generation: %.10f\n", res.keff, res.keff_std, res.k_generation);
    return res;
  }

  int batch() const { return _batch; }
  int rank() const { return _rank; }
  const mcw_generation_result& last() const { return _last; }
  const Handle& handle() const { return _h; }
  int64_t localParticles() const { return _work_index[_rank + 1] - _work_index[_rank]; }

  // openmc_source_bank analogue (OMC bank.cpp:110-125)
  std::vector<mcw_source_site> sourceBank() const
  {
    int64_t n = 0;
    _h.check(mcw_get_source_bank(_h.get(), nullptr, 0, &n));
    std::vector<mcw_source_site> out((size_t)n);
    _h.check(mcw_get_source_bank(_h.get(), out.data(), n, &n));
    return out;
  }
  void setSourceBank(const std::vector<mcw_source_site>& sites)
  {
    _h.check(mcw_set_source_bank(_h.get(), sites.data(), (int64_t)sites.size()));
  }

private:
  // MC OpenMCStudy.C:454-607 with its MPI calls mapped onto Comm (see macaw_b200/dist.py for the
  // same sequence on torch.distributed): bank offsets, one fused all-reduce, sampling of this rank's
  // slice of the global bank, last-rank trim / pad, redistribution to the owners.
  void synchronizeBanks()
  {
    mcw_handle h = _h.get();
    void* stream = nullptr;
    _h.check(mcw_stream(h, &stream));
    int64_t n_local = 0;
    _h.check(mcw_bank_count(h, &n_local));
    auto counts = _comm->allGather(n_local);
    std::vector<int64_t> off(_nranks + 1, 0);
    for (int r = 0; r < _nranks; ++r)
    {
      if (counts[r] == 0)
        throw Error(MCW_E_PHYSICS, "No fission sites banked on rank " + std::to_string(r)); // OMC eigenvalue.cpp:107-109
      off[r + 1] = off[r] + counts[r];
    }
    _h.check(mcw_set_bank_window(h, off[_rank], off[_nranks]));
    void* red = nullptr;
    int64_t n_red = 0;
    _h.check(mcw_reduction_device_ptr(h, &red, &n_red));
    _comm->allReduceSum(static_cast<double*>(red), n_red, stream);
    int64_t n_s = 0;
    _h.check(mcw_sample_bank(h, _batch, &n_s));
    auto scounts = _comm->allGather(n_s);
    std::vector<int64_t> soff(_nranks + 1, 0);
    for (int r = 0; r < _nranks; ++r)
      soff[r + 1] = soff[r] + scounts[r];
    _h.check(mcw_finish_sampling(h, soff[_rank], _rank == _nranks - 1, &n_s));
    // only the last rank trims / pads and it always ends at exactly n_particles (OMC eigenvalue.cpp:184-206)
    scounts[_nranks - 1] = std::max<int64_t>(0, _work_index[_nranks] - soff[_nranks - 1]);
    for (int r = 0; r < _nranks; ++r)
      soff[r + 1] = soff[r] + scounts[r];
    void *sampled = nullptr, *source = nullptr;
    int64_t n_sampled = 0, cap = 0;
    _h.check(mcw_sampled_device_ptr(h, &sampled, &n_sampled));
    _h.check(mcw_source_device_ptr(h, &source, &cap));
    const int64_t S = (int64_t)sizeof(mcw_source_site);
    std::vector<Comm::Range> sends, recvs;
    for (int d = 0; d < _nranks; ++d) // what I hold [soff[rank], soff[rank+1]) vs what d owns
    {
      const int64_t lo = std::max(soff[_rank], _work_index[d]), hi = std::min(soff[_rank + 1], _work_index[d + 1]);
      if (hi > lo)
        sends.push_back({d, (lo - soff[_rank]) * S, (hi - lo) * S});
    }
    for (int s = 0; s < _nranks; ++s)
    {
      const int64_t lo = std::max(soff[s], _work_index[_rank]), hi = std::min(soff[s + 1], _work_index[_rank + 1]);
      if (hi > lo)
        recvs.push_back({s, (lo - _work_index[_rank]) * S, (hi - lo) * S});
    }
    _comm->exchange(static_cast<const char*>(sampled), sends, static_cast<char*>(source), recvs, stream);
    _h.check(mcw_set_source_count(h, localParticles()));
  }

  Comm* _comm;
  int _rank, _nranks;
  Handle _h;
  bool _verbose;
  mcw_settings _settings;
  std::vector<int64_t> _work_index;
  int _batch = 0;
  mcw_generation_result _last{};
};

// Parameters: tally_id, granularity, score, estimator, particle_type, energy_bin (MC OpenMCTallyAux.C:26-50).
// computeValue(elem) = results_(filter_index, score, SUM) summed over the energy bins unless
// energy_bin is given -- the raw sum over batches, not the mean (:183,201).
class OpenMCTallyAux
{
public:
  OpenMCTallyAux(const OpenMCStudy& study, const OpenMCTally& tally, const std::string& granularity, const std::string& score,
                 std::optional<int> energy_bin = std::nullopt, const std::string& nuclide = "all",
                 const Library* library = nullptr)
    : _study(study), _tally_id(tally.id()), _energy_bin(energy_bin)
  {
    if (granularity != "cell" && granularity != "universe")
      mooseError("global granularity is not supported");
    _score = -1;
    for (size_t k = 0; k < tally.scoreNames().size(); ++k)
      if (tally.scoreNames()[k] == score)
        _score = (int)k;
    if (_score < 0)
      paramError("score", "score '" + score + "' is not part of tally " + std::to_string(_tally_id));
    // MC OpenMCTallyAux.C:93-119: the nuclide bin and the score index.  The reference combines them as
    // score_bin * nuclides_.size() + nuc_bin although OpenMC stores nuclide_bin * n_scores + score_bin
    // (OMC tally_scoring.cpp:2400); the two agree when either count is 1.  Kept as the reference has it and flagged here.
    const auto& ref = tally.referenceNuclides();
    int nuc_bin = 0;
    if (ref[0] != -1)
    {
      if (nuclide == "all")
        nuc_bin = (int)ref.size() - 1;
      else
      {
        int idx = -1;
        if (library)
          for (size_t i = 0; i < library->nuclides.size(); ++i)
            if (library->nuclides[i].name == nuclide)
              idx = (int)i;
        nuc_bin = (int)(std::find(ref.begin(), ref.end(), idx) - ref.begin());
      }
    }
    _score = _score * (int)ref.size() + nuc_bin;
  }
  // fetches the tally once per step (the AuxKernel is evaluated once per element per step)
  void timestepSetup()
  {
    const Handle& h = _study.handle();
    h.check(mcw_tally_shape(h.get(), _tally_id, &_n_spatial, &_n_energy, &_n_scores));
    int32_t n_bins = 1;
    h.check(mcw_tally_nuclide_bins(h.get(), _tally_id, &n_bins));
    _n_scores *= n_bins; // OpenMC's score_index dimension: nuclide bins x scores
    _sum.assign((size_t)(_n_spatial * _n_energy * _n_scores), 0.0);
    _sum_sq = _sum;
    int32_t nr = 0;
    h.check(mcw_get_tally(h.get(), _tally_id, _sum.data(), _sum_sq.data(), (int64_t)_sum.size(), &nr));
  }
  double computeValue(int64_t spatial_bin) const
  {
    double v = 0.0;
    for (int e = 0; e < _n_energy; ++e)
      if (!_energy_bin || *_energy_bin == e)
        v += _sum[(size_t)((spatial_bin * _n_energy + e) * _n_scores + _score)];
    return v;
  }
  int64_t nSpatial() const { return _n_spatial; }

private:
  const OpenMCStudy& _study;
  int _tally_id, _score;
  std::optional<int> _energy_bin;
  int64_t _n_spatial = 0;
  int32_t _n_energy = 0, _n_scores = 0;
  std::vector<double> _sum, _sum_sq;
};

} // namespace macaw
