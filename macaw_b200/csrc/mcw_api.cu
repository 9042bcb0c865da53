// mcw_api.cu -- host side of libmacaw_b200.so: context, set-up, the event loop, C ABI.
// Declarations and the reference interfaces each entry point replaces are in include/macaw_b200.h.
// There is no CPU execution path in this file: every compute call launches the sm_100a kernels of
// mcw_kernels.cu and fails with MCW_E_CUDA if no Blackwell-class device is present.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mcw_detmath.h"
#include "mcw_kernels.h"

// launcher of the kernel build this problem needs (mcw_kernels.h): `h` must be in scope
#define MCWK(fn, ...) (h->rare ? mcw_r1::fn(__VA_ARGS__) : mcw_r0::fn(__VA_ARGS__))

namespace
{
struct HostTally
{
  int id, estimator, granularity, n_scores, n_energy;
  int scores[MCW_N_SCORES];
  std::vector<double> ebins;
  std::vector<int32_t> nuclides; // nuclide bins (empty = the single bin "total")
  int32_t* d_nuclides = nullptr;
  int n_nuc_bins() const { return nuclides.empty() ? 1 : (int)nuclides.size(); }
  int64_t n_spatial = 1, n_values = 0, offset = 0; // offset into the reduction buffer
  double *d_sum = nullptr, *d_sum_sq = nullptr, *d_ebins = nullptr;
  int n_realizations = 0;
};

struct HostThermal // one mcw_thermal_desc, copied
{
  double energy_max = 0;
  int inel_type = 0, el_type = 0, n_mu = 0, n_e_out = 0, skewed = 0, n_el_mu = 0;
  std::vector<double> ixs_e, ixs, ien, eout, pdf, cdf, mu, el_x, el_y, el_mu;
  std::vector<int32_t> ioff;
};

struct HostTable // one mcw_tabular2d, copied
{
  std::vector<double> energy, x, p, c;
  std::vector<int32_t> off, interp, ndisc;
  int hist = 0;
};

struct HostUrr // one mcw_urr_desc, copied
{
  std::vector<double> energy, table;
  int n_band = 0, inel_flag = -1, multiply_smooth = 0;
};

struct HostFn // mcw_function1d, copied
{
  int kind = MCW_FN_NONE, interp = 2;
  std::vector<double> x, y;
};

struct HostFis // mcw_fission_product, copied
{
  HostFn yield, a, b;
  int law = MCW_LAW_WATT;
  double u = 0;
  HostTable energy;
};

struct HostRx // one scattering reaction after a nuclide's first (mcw_reaction_desc), copied
{
  HostFn yield_fn, fa, fb; // energy-dependent multiplicity; parameters of the Maxwell / evaporation / Watt laws
  double fu = 0;
  McwRxParams prm;
  std::vector<double> xs;
  HostTable angle, energy, law;
};

struct HostNuclide
{
  // several temperatures: entries [0, n_identities) of Handle::nuclides are the nuclides, later ones further data sets of
  // nuclide `identity`; `sets` (on the nuclide's own entry) = all its data sets by increasing kT
  int identity = -1;
  std::vector<int> sets;
  std::vector<HostRx> more;
  HostUrr urr;
  McwNucParams prm;
  std::vector<double> energy, xs5, elastic, inelastic_full;
  std::vector<int32_t> grid_index;
  HostTable el_angle, inel_angle, inel_energy;
  HostTable law; // correlated: one angular distribution per point of inel_energy; Kalbach-Mann: one entry, x = r, p = a
  double inel_yield = 1.0;
  HostFn inel_yield_fn, inel_fa, inel_fb; // the inline reaction's mcw_reaction_desc::yield_fn, a, b, u
  double inel_fu = 0;
  std::vector<HostFis> fis; // fission neutrons as a nuclide file has them (empty: the scalar description)
  HostFn total_nu;
};

// copies and checks one table of the descriptor; returns an error text or nullptr
static const char* take_table(const mcw_tabular2d& t, HostTable& out, bool continuous)
{
  out = HostTable();
  if (t.n_energy == 0)
    return nullptr;
  if (t.n_energy < 2 || !t.energy || !t.offset || !t.interpolation || !t.x || !t.pdf || !t.cdf)
    return "tabulated distribution incomplete (needs >= 2 incoming energies and all arrays)";
  const int n = t.n_energy;
  out.energy.assign(t.energy, t.energy + n);
  out.off.assign(t.offset, t.offset + n + 1);
  out.interp.assign(t.interpolation, t.interpolation + n);
  out.ndisc.assign(n, 0);
  if (continuous && t.n_discrete)
    out.ndisc.assign(t.n_discrete, t.n_discrete + n);
  out.hist = continuous ? (t.histogram_in_energy != 0) : 0;
  for (int k = 0; k < n; ++k)
  {
    if (k && !(out.energy[k] > out.energy[k - 1]))
      return "tabulated distribution: incoming energies must increase";
    const int m = out.off[k + 1] - out.off[k];
    if (out.off[k] < 0 || m < 2 || out.ndisc[k] < 0 || out.ndisc[k] > m - 2)
      return "tabulated distribution: every distribution needs at least two continuous points";
    if (out.interp[k] != 1 && out.interp[k] != 2)
      // OMC distribution.cpp:201-205 "Only histogram and linear-linear interpolation ..."
      return "Only histogram and linear-linear interpolation for tabular distribution is supported.";
    if (t.cdf[out.off[k + 1] - 1] != 1.0)
      return "tabulated distribution: the cdf must end at exactly 1";
  }
  const int total = out.off[n];
  out.x.assign(t.x, t.x + total);
  out.p.assign(t.pdf, t.pdf + total);
  out.c.assign(t.cdf, t.cdf + total);
  return nullptr;
}

// copies and checks one function; returns an error text or nullptr
static const char* take_fn(const mcw_function1d& f, HostFn& out, bool required)
{
  out = HostFn();
  if (f.kind == MCW_FN_NONE)
    return required ? "fission product: a function it needs is missing" : nullptr;
  if ((f.kind != MCW_FN_POLYNOMIAL && f.kind != MCW_FN_TABULATED) || f.n < 1 || !f.y)
    return "function: polynomial or tabulated, with its values";
  out.kind = f.kind;
  out.y.assign(f.y, f.y + f.n);
  if (f.kind == MCW_FN_TABULATED)
  {
    if (f.n < 2 || !f.x)
      return "tabulated function: needs >= 2 points";
    if (f.interpolation != 1 && f.interpolation != 2)
      // OMC endf.cpp:190-205 also knows lin-log, log-lin and log-log
      return "tabulated function: only histogram and linear-linear interpolation are carried";
    out.interp = f.interpolation;
    out.x.assign(f.x, f.x + f.n);
    for (int k = 1; k < f.n; ++k)
      if (!(out.x[k] > out.x[k - 1]))
        return "tabulated function: abscissae must increase";
  }
  return nullptr;
}

// a scattering reaction's optional yield function and the parameter functions of the Maxwell / evaporation / Watt laws
static const char* take_law_functions(int law, const mcw_function1d& yield_fn, const mcw_function1d& a, const mcw_function1d& b,
                                      HostFn& out_yield, HostFn& out_a, HostFn& out_b)
{
  const char* bad = take_fn(yield_fn, out_yield, false);
  if (!bad && law >= MCW_LAW_MAXWELL)
    bad = take_fn(a, out_a, true);
  if (!bad && law == MCW_LAW_WATT)
    bad = take_fn(b, out_b, true);
  return bad;
}

// checks a reaction's law against its outgoing-energy table and copies what the product distributions that give energy and
// cosine together need into `out` (correlated: one angular distribution per point of the energy table; Kalbach-Mann: one
// entry, x = r, p = a); returns an error text or nullptr
static const char* take_law(int law, const HostTable& energy, const int32_t* corr_offset, const int32_t* corr_interp,
                            const double* corr_mu, const double* corr_pdf, const double* corr_cdf, const double* km_r,
                            const double* km_a, int nbody_n, HostTable& out)
{
  out = HostTable();
  if ((law == MCW_LAW_CONTINUOUS_TABULAR || law == MCW_LAW_CORRELATED || law == MCW_LAW_KALBACH_MANN) && energy.energy.empty())
    return "tabulated inelastic law without its outgoing-energy table";
  if (law == MCW_LAW_CORRELATED)
  {
    // one angular Tabular per point of the outgoing-energy table (OMC secondary_correlated.cpp:100-147)
    const int np = energy.off.back();
    if (!corr_offset || !corr_interp || !corr_mu || !corr_pdf || !corr_cdf)
      return "correlated law without its angular distributions";
    out.energy.assign(np, 0.0);
    out.off.assign(corr_offset, corr_offset + np + 1);
    out.interp.assign(corr_interp, corr_interp + np);
    out.ndisc.assign(np, 0);
    for (int k = 0; k < np; ++k)
      if (out.off[k + 1] - out.off[k] < 2 || (out.interp[k] != 1 && out.interp[k] != 2))
        return "correlated law: every angular distribution needs >= 2 points, histogram or lin-lin";
    out.x.assign(corr_mu, corr_mu + out.off[np]);
    out.p.assign(corr_pdf, corr_pdf + out.off[np]);
    out.c.assign(corr_cdf, corr_cdf + out.off[np]);
  }
  if (law == MCW_LAW_KALBACH_MANN)
  {
    const int np = energy.off.back();
    if (!km_r || !km_a)
      return "Kalbach-Mann law without r / a";
    out.energy.assign(1, 0.0);
    out.off = {0, np};
    out.interp.assign(1, 2);
    out.ndisc.assign(1, 0);
    out.x.assign(km_r, km_r + np);
    out.p.assign(km_a, km_a + np);
    out.c.assign(np, 0.0);
  }
  if (law == MCW_LAW_NBODY && (nbody_n < 3 || nbody_n > 5))
    // OMC secondary_nbody.cpp:57 "N-body phase space with >5 bodies."
    return "N-body phase space needs 3, 4 or 5 bodies";
  return nullptr;
}

template <int TYPE>
struct Topo;
template <>
struct Topo<MCW_QUAD4>
{
  static constexpr int NV = 4, NS = 4, NPS = 2;
  static const int (*side_nodes())[4]
  {
    static const int t[4][4] = {{0, 1, 0, 0}, {1, 2, 0, 0}, {2, 3, 0, 0}, {3, 0, 0, 0}};
    return t;
  }
};
template <>
struct Topo<MCW_TET4>
{
  static constexpr int NV = 4, NS = 4, NPS = 3;
  static const int (*side_nodes())[4]
  {
    static const int t[4][4] = {{0, 2, 1, 0}, {0, 1, 3, 0}, {1, 2, 3, 0}, {2, 0, 3, 0}};
    return t;
  }
};
template <>
struct Topo<MCW_HEX8>
{
  static constexpr int NV = 8, NS = 6, NPS = 4;
  static const int (*side_nodes())[4]
  {
    static const int t[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
    return t;
  }
};

struct TopoRT
{
  int nv, ns, nps, cpv, stride;
  const int (*sn)[4];
};
TopoRT topo_of(int type)
{
  if (type == MCW_QUAD4)
    return {4, 4, 2, 2, McwElem<MCW_QUAD4>::STRIDE, Topo<MCW_QUAD4>::side_nodes()};
  if (type == MCW_TET4)
    return {4, 4, 3, 3, McwElem<MCW_TET4>::STRIDE, Topo<MCW_TET4>::side_nodes()};
  return {8, 6, 4, 3, McwElem<MCW_HEX8>::STRIDE, Topo<MCW_HEX8>::side_nodes()};
}
} // namespace

struct mcw_context
{
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<void*> allocs;

  // host-side problem description
  int elem_type = -1;
  int64_t n_nodes = 0, n_elem = 0;
  std::vector<double> xyz;
  std::vector<int64_t> conn, nbr;
  std::vector<int32_t> block_raw, side_bnd, bnd_kind;
  std::vector<HostNuclide> nuclides;
  std::vector<std::vector<int32_t>> mat_nuc;
  std::vector<std::vector<double>> mat_dens;
  std::vector<std::vector<int32_t>> mat_iso;
  std::vector<HostThermal> thermal;                // S(a,b) tables (data::thermal_scatt)
  std::vector<std::vector<int32_t>> mat_sab;       // per material, per nuclide position: table or -1 (empty = none)
  std::vector<std::vector<double>> mat_sab_frac;
  std::vector<int32_t> block_ids, block_material;
  int n_identities = 0; // nuclides given to mcw_set_library; h->nuclides beyond are further temperature sets
  // several temperatures of S(a,b) tables (mcw_set_thermal_temperatures): per thermal entry the first entry of its table
  // and its kT; empty = one temperature per table
  std::vector<int32_t> thermal_same;
  std::vector<double> thermal_kT;
  std::vector<double> temperature;
  std::vector<HostTally> tallies;
  mcw_settings set{};
  bool have_mesh = false, have_lib = false, have_mats = false, have_set = false, prepared = false;
  std::vector<int64_t> work_index;
  int64_t n_local = 0;

  // device state
  McwParams P{};
  McwCounters* h_counters = nullptr; // pinned
  mcw_source_site *d_source = nullptr, *d_bank = nullptr, *d_sorted = nullptr, *d_sampled = nullptr;
  int32_t *d_q_fly[2] = {nullptr, nullptr}, *d_q_collide = nullptr, *d_q_fallback = nullptr, *d_q_sorted = nullptr;
  uint32_t *d_keys = nullptr, *d_bins = nullptr, *d_bin_offsets = nullptr; // energy binning of the refresh queue (mcw_xs.cuh)
  int32_t* d_q_hist = nullptr;
  // event rounds with fewer collisions refresh cross sections without sorting; 0 = chosen from the
  // number of (slot, nuclide) plan entries: the sort pays off sooner the more lookups it serves
  // (A/B: 39 entries from ~32k collisions, 12 entries from ~200k; profiles/ab_r01h_sort_threshold.txt)
  int64_t xs_sort_min = 0;
  int64_t *d_i64_a = nullptr, *d_i64_b = nullptr;
  void* d_temp = nullptr;
  size_t temp_bytes = 0;
  double* d_reduce = nullptr; // [tally values | k_col k_abs k_trk leak | weight]
  int64_t n_reduce = 0, n_tally_values = 0;
  double* d_box = nullptr;
  mcw_history_record* d_records = nullptr;
  int64_t n_source = 0, n_bank = 0, n_sampled = 0, bank_capacity = 0, sampled_capacity = 0;
  int64_t bank_global_start = 0, bank_global_total = -1;
  bool records_valid = false;

  // eigenvalue bookkeeping (OMC simulation.cpp:253-276)
  double keff = 1.0, keff_std = 0.0, k_sum[2] = {0, 0};
  int n_realizations = 0;
  std::vector<double> k_generation;
  double gt_sum[4] = {0, 0, 0, 0}, gt_sum_sq[4] = {0, 0, 0, 0};

  // which build of the kernels runs this problem (mcw_kernels.h): 1 when it has S(a,b) tables, tallies with nuclide
  // bins, cached trace points or tabulated secondary distributions; "force_rare" option for tests
  bool rare = false;
  int force_rare = -1;

  // options
  int profile = 0, sync_every = 1, fallback_blocks = 8, visit_cap = 1 << 30, track_max_blocks = 148, fused_max_blocks = 148;
  int drain_lanes = 28;        // exhausted tracking warps below this many histories hand them to the next round ...
  int64_t drain_min_queue = 0; // ... when the round's queue is at least this long (sweeps: profiles/ab_r01g_drain.txt)
  // below this many histories in flight the tracking kernel runs fused (collisions inline, no
  // round per collision): rounds would last as long as the slowest flight with most lanes idle
  int64_t tail_threshold = 16384; // sweeps: profiles/sweep_r02_tail_threshold.txt (49152 before the warp-shared refresh)
  int fallback_every_round = 0; // 1: serve the general-walk queue in every round (A/B, tests)
  int xs_plan = 1; // sorted refresh: stage all (slot, nuclide) windows of a tile at once (k_xs_all_slots_plan)
  int ordered_queues = 1; // rebuild the collide / fly queues in history-index order every round
  int collide_min = 8; // fused mode: a warp processes its parked collisions once this many wait
  // below this many histories in flight the rest of the generation runs in the persistent kernel (mcw_pump.cuh);
  // 0 = never (event rounds, then the fused tracking kernel below tail_threshold)
  int64_t pump_threshold = 0;
  int pump_track_warps = 2, pump_patience = 4, pump_blocks = 0;
  McwPump Q{};
  void* d_pump_ctl = nullptr;
  size_t pump_ctl_bytes = 0;
  mcw_generation_result last{};
  cudaEvent_t ev[8] = {};
  std::vector<cudaEvent_t> pool; // per-round profiling events (5 per round), reused across generations
  cudaEvent_t pool_event(size_t i)
  {
    while (pool.size() <= i)
    {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[i];
  }
};

#define CK(call)                                                                                               \
  do                                                                                                           \
  {                                                                                                            \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
    {                                                                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                             \
      return MCW_E_CUDA;                                                                                       \
    }                                                                                                          \
  } while (0)

static int fail(mcw_handle h, int code, const std::string& msg)
{
  if (h)
    h->err = msg;
  return code;
}

template <typename T>
static cudaError_t dev_alloc(mcw_handle h, T** p, size_t n)
{
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e == cudaSuccess)
  {
    h->allocs.push_back(q);
    *p = (T*)q;
  }
  return e;
}

template <typename T>
static cudaError_t dev_upload(mcw_handle h, T** p, const std::vector<T>& v)
{
  cudaError_t e = dev_alloc(h, p, v.size());
  if (e != cudaSuccess)
    return e;
  if (!v.empty())
    e = cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

static void dev_free(mcw_handle h, void* p)
{
  if (!p)
    return;
  auto it = std::find(h->allocs.begin(), h->allocs.end(), p);
  if (it != h->allocs.end())
    h->allocs.erase(it);
  cudaFree(p);
}

extern "C" {

int mcw_version(void) { return 100; }

int mcw_create(mcw_handle* out, int device_ordinal, int rank, int nranks)
{
  if (!out || nranks < 1 || rank < 0 || rank >= nranks)
    return MCW_E_INVALID_ARGUMENT;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
  {
    cudaGetLastError();
    return MCW_E_CUDA; // no GPU: there is deliberately no CPU fallback
  }
  if (device_ordinal < 0 || device_ordinal >= n_dev)
    return MCW_E_INVALID_ARGUMENT;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess)
    return MCW_E_CUDA;
  if (prop.major != 10)
    return MCW_E_CUDA; // kernels are built for sm_100a only
  if (cudaSetDevice(device_ordinal) != cudaSuccess)
    return MCW_E_CUDA;
  mcw_handle h = new mcw_context();
  h->device = device_ordinal;
  h->rank = rank;
  h->nranks = nranks;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_counters, sizeof(McwCounters)) != cudaSuccess)
  {
    delete h;
    return MCW_E_CUDA;
  }
  for (auto& e : h->ev)
    cudaEventCreate(&e);
  *out = h;
  return MCW_OK;
}

void mcw_destroy(mcw_handle h)
{
  if (!h)
    return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs)
    cudaFree(p);
  if (h->h_counters)
    cudaFreeHost(h->h_counters);
  for (auto& e : h->ev)
    if (e)
      cudaEventDestroy(e);
  for (auto& e : h->pool)
    cudaEventDestroy(e);
  if (h->stream)
    cudaStreamDestroy(h->stream);
  delete h;
}

const char* mcw_last_error(mcw_handle h) { return h ? h->err.c_str() : "invalid handle"; }

int mcw_set_option(mcw_handle h, const char* name, int value)
{
  if (!h || !name)
    return MCW_E_INVALID_ARGUMENT;
  std::string n(name);
  if (n == "profile")
    h->profile = value;
  else if (n == "sync_every")
    h->sync_every = std::max(1, value);
  else if (n == "fallback_blocks")
    h->fallback_blocks = std::max(1, value);
  else if (n == "drain_lanes")
    h->drain_lanes = std::min(32, std::max(0, value));
  else if (n == "drain_min_queue")
    h->drain_min_queue = std::max(0, value);
  else if (n == "tail_threshold")
    h->tail_threshold = std::max(0, value);
  else if (n == "fallback_every_round")
    h->fallback_every_round = value != 0;
  else if (n == "xs_plan")
    h->xs_plan = value != 0;
  else if (n == "ordered_queues")
    h->ordered_queues = value != 0;
  else if (n == "xs_sort_min")
    h->xs_sort_min = std::max(1, value);
  else if (n == "collide_min")
    h->collide_min = std::min(32, std::max(1, value));
  else if (n == "force_rare")
  {
    h->force_rare = value;
    h->prepared = false;
  }
  else if (n == "pump_threshold")
    h->pump_threshold = std::max(0, value);
  else if (n == "pump_track_warps")
    h->pump_track_warps = std::min(3, std::max(1, value));
  else if (n == "pump_patience")
    h->pump_patience = std::max(1, value);
  else if (n == "visit_cap")
  {
    h->visit_cap = value > 0 ? value : (1 << 30);
    h->P.visit_cap = h->visit_cap;
  }
  else
    return fail(h, MCW_E_INVALID_ARGUMENT, "unknown option " + n);
  return MCW_OK;
}

int mcw_stream(mcw_handle h, void** stream)
{
  if (!h || !stream)
    return MCW_E_INVALID_ARGUMENT;
  *stream = (void*)h->stream;
  return MCW_OK;
}

int mcw_calculate_work(int64_t n_particles, int nranks, int64_t* work_index)
{
  // OMC simulation.cpp:529-551
  if (nranks < 1 || !work_index || n_particles < 0)
    return MCW_E_INVALID_ARGUMENT;
  int64_t min_work = n_particles / nranks, rem = n_particles % nranks;
  work_index[0] = 0;
  for (int i = 0; i < nranks; ++i)
    work_index[i + 1] = work_index[i] + (i < rem ? min_work + 1 : min_work);
  return MCW_OK;
}

int mcw_set_mesh(mcw_handle h, int elem_type, int64_t n_nodes, const double* xyz, int64_t n_elem, const int64_t* conn,
                 const int32_t* block_id, const int64_t* neighbors, const int32_t* side_bnd_id, int n_bnd,
                 const int32_t* bnd_kind)
{
  if (!h)
    return MCW_E_INVALID_ARGUMENT;
  if (elem_type != MCW_QUAD4 && elem_type != MCW_TET4 && elem_type != MCW_HEX8)
    return fail(h, MCW_E_UNSUPPORTED, "element type not supported (QUAD4, TET4, HEX8)");
  if (!xyz || !conn || !block_id || n_nodes <= 0 || n_elem <= 0 || n_elem > 0x7fffffff)
    return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_mesh: bad arguments");
  TopoRT t = topo_of(elem_type);
  h->elem_type = elem_type;
  h->n_nodes = n_nodes;
  h->n_elem = n_elem;
  h->xyz.assign(xyz, xyz + 3 * n_nodes);
  h->conn.assign(conn, conn + n_elem * t.nv);
  h->block_raw.assign(block_id, block_id + n_elem);
  for (int64_t i = 0; i < n_elem * t.nv; ++i)
    if (conn[i] < 0 || conn[i] >= n_nodes)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_mesh: connectivity out of range");
  if (neighbors)
    h->nbr.assign(neighbors, neighbors + n_elem * t.ns);
  else
  {
    // face matching (what libMesh find_neighbors provides as elem->neighbor_ptr, RT TraceRay.C:1433)
    struct Face
    {
      int64_t k[4];
      int64_t es;
    };
    std::vector<Face> faces((size_t)n_elem * t.ns);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int s = 0; s < t.ns; ++s)
      {
        Face& f = faces[e * t.ns + s];
        f.k[0] = f.k[1] = f.k[2] = f.k[3] = -1;
        for (int i = 0; i < t.nps; ++i)
          f.k[i] = conn[e * t.nv + t.sn[s][i]];
        std::sort(f.k, f.k + t.nps);
        f.es = e * t.ns + s;
      }
    std::sort(faces.begin(), faces.end(), [](const Face& a, const Face& b) {
      for (int i = 0; i < 4; ++i)
        if (a.k[i] != b.k[i])
          return a.k[i] < b.k[i];
      return a.es < b.es;
    });
    h->nbr.assign((size_t)n_elem * t.ns, -1);
    for (size_t i = 0; i + 1 < faces.size(); ++i)
      if (!memcmp(faces[i].k, faces[i + 1].k, sizeof(faces[i].k)))
      {
        h->nbr[faces[i].es] = faces[i + 1].es / t.ns;
        h->nbr[faces[i + 1].es] = faces[i].es / t.ns;
      }
  }
  if (side_bnd_id)
    h->side_bnd.assign(side_bnd_id, side_bnd_id + n_elem * t.ns);
  else
    h->side_bnd.assign((size_t)n_elem * t.ns, -1);
  h->bnd_kind.assign(bnd_kind, bnd_kind + (bnd_kind ? n_bnd : 0));
  h->have_mesh = true;
  h->prepared = false;
  return MCW_OK;
}

} // extern "C"

// mcw_set_library's checks and copies, without a device: also behind mcw_validate_library
static int parse_library(const mcw_nuclide_desc* d, int n, std::vector<HostNuclide>& nuclides, std::string& err)
{
  auto fail = [&](void*, int code, const std::string& msg) {
    err = msg;
    return code;
  };
  void* h = nullptr;
  nuclides.clear();
  int64_t off = 0;
  for (int i = 0; i < n; ++i)
  {
    if (d[i].n_grid < 2 || !d[i].energy || !d[i].xs || !d[i].elastic)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_library: nuclide table missing");
    if (d[i].n_precursor < 0 || d[i].n_precursor > MCW_MAX_DELAYED_GROUPS)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_library: too many delayed groups");
    HostNuclide N;
    memset(&N.prm, 0, sizeof(N.prm));
    const int G = d[i].n_grid;
    N.prm.G = G;
    N.prm.off = off;
    off += G + (G & 1); // every nuclide starts at an even grid point: 16-byte aligned windows (mcw_xs.cuh)
    N.prm.fissionable = d[i].fissionable != 0;
    N.prm.awr = d[i].awr;
    N.prm.kT = d[i].kT;
    N.prm.n_inel = d[i].n_inelastic;
    N.prm.inel_thr = d[i].inel_threshold;
    N.prm.inel_E = d[i].inel_threshold_energy;
    N.prm.inel_ratio = d[i].inel_mass_ratio;
    N.prm.nu_poly[0] = d[i].nu_poly[0];
    N.prm.nu_poly[1] = d[i].nu_poly[1];
    N.prm.n_precursor = d[i].n_precursor;
    for (int g = 0; g < MCW_MAX_DELAYED_GROUPS; ++g)
    {
      N.prm.dyield[g] = d[i].delayed_yield[g];
      N.prm.dtheta[g] = d[i].delayed_theta[g];
    }
    N.prm.watt_a = d[i].watt_a;
    N.prm.watt_b = d[i].watt_b;
    N.prm.q = d[i].q_fission;
    N.energy.assign(d[i].energy, d[i].energy + G);
    for (int k = 1; k < G; ++k)
      if (N.energy[k] < N.energy[k - 1])
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_library: energy grid not sorted");
    N.xs5.assign(d[i].xs, d[i].xs + 5 * (size_t)G);
    N.elastic.assign(d[i].elastic, d[i].elastic + G);
    N.inelastic_full.assign(G, 0.0);
    if (N.prm.n_inel)
    {
      if (!d[i].inelastic || N.prm.inel_thr < 0 || N.prm.inel_thr >= G)
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_library: inelastic table missing");
      for (int k = N.prm.inel_thr; k < G; ++k)
        N.inelastic_full[k] = d[i].inelastic[k - N.prm.inel_thr];
    }
    N.prm.inel_law = d[i].inelastic_law;
    N.prm.inel_in_cm = N.prm.inel_law == MCW_LAW_LEVEL ? 1 : (d[i].inelastic_in_cm != 0);
    if (N.prm.inel_law < MCW_LAW_LEVEL || N.prm.inel_law > MCW_LAW_WATT)
      // OMC secondary_uncorrelated.cpp:47 "Energy distribution type '...' not implemented."
      return fail(h, MCW_E_UNSUPPORTED, "mcw_set_library: inelastic law not implemented (level, continuous tabular, "
                                        "correlated, Kalbach-Mann, N-body, Maxwell, evaporation, Watt)");
    const char* bad = take_table(d[i].elastic_angle, N.el_angle, false);
    if (!bad)
      bad = take_table(d[i].inelastic_angle, N.inel_angle, false);
    if (!bad)
      bad = take_table(d[i].inelastic_energy, N.inel_energy, true);
    const int law = N.prm.inel_law;
    if (!bad && N.prm.n_inel)
      bad = take_law(law, N.inel_energy, d[i].corr_offset, d[i].corr_interp, d[i].corr_mu, d[i].corr_pdf, d[i].corr_cdf, d[i].km_r,
                     d[i].km_a, d[i].nbody_n, N.law);
    N.prm.nbody_n = d[i].nbody_n;
    N.prm.nbody_mass_ratio = d[i].nbody_mass_ratio;
    N.prm.nbody_q = d[i].nbody_q;
    if (law >= MCW_LAW_CORRELATED && law <= MCW_LAW_NBODY)
      N.inel_angle = HostTable(); // the product distribution gives the cosine itself
    N.inel_yield = d[i].inelastic_yield == 0.0 ? 1.0 : d[i].inelastic_yield;
    if (!bad && N.prm.n_inel)
      bad = take_law_functions(law, d[i].inelastic_yield_fn, d[i].inelastic_a, d[i].inelastic_b, N.inel_yield_fn, N.inel_fa, N.inel_fb);
    N.inel_fu = d[i].inelastic_u;
    if (!bad && d[i].n_more_reactions > 0 && (!N.prm.n_inel || !d[i].more_reactions))
      bad = "more_reactions needs the inline inelastic reaction (n_inelastic = 1) before them";
    for (int j = 0; !bad && j < d[i].n_more_reactions; ++j)
    {
      const mcw_reaction_desc& r = d[i].more_reactions[j];
      HostRx R;
      memset(&R.prm, 0, sizeof(R.prm));
      if (r.threshold < 0 || r.threshold >= G - 1 || !r.xs)
        bad = "reaction threshold outside the grid, or its cross section missing";
      else if (r.law < MCW_LAW_LEVEL || r.law > MCW_LAW_WATT)
        bad = "reaction law not implemented (level, continuous tabular, correlated, Kalbach-Mann, N-body, Maxwell, evaporation, Watt)";
      if (bad)
        break;
      R.xs.assign(r.xs, r.xs + (G - r.threshold));
      R.prm.thr = r.threshold;
      R.prm.law = r.law;
      R.prm.in_cm = r.law == MCW_LAW_LEVEL ? 1 : (r.in_cm != 0);
      R.prm.level_E = r.threshold_energy;
      R.prm.level_ratio = r.mass_ratio;
      R.prm.yield = r.yield == 0.0 ? 1.0 : r.yield;
      R.prm.nbody_n = r.nbody_n;
      R.prm.nbody_mass_ratio = r.nbody_mass_ratio;
      R.prm.nbody_q = r.nbody_q;
      bad = take_table(r.angle, R.angle, false);
      if (!bad)
        bad = take_table(r.energy, R.energy, true);
      if (!bad)
        bad = take_law(r.law, R.energy, r.corr_offset, r.corr_interp, r.corr_mu, r.corr_pdf, r.corr_cdf, r.km_r, r.km_a, r.nbody_n,
                       R.law);
      if (r.law >= MCW_LAW_CORRELATED && r.law <= MCW_LAW_NBODY)
        R.angle = HostTable();
      if (!bad)
        bad = take_law_functions(r.law, r.yield_fn, r.a, r.b, R.yield_fn, R.fa, R.fb);
      R.fu = r.u;
      N.more.push_back(std::move(R));
    }
    if (!bad && d[i].n_fission_products > 0)
    {
      // fission neutrons as a nuclide file has them: prompt + one product per delayed group
      if (!N.prm.fissionable || !d[i].fission_products || d[i].n_fission_products != 1 + d[i].n_precursor)
        bad = "fission products: needs a fissionable nuclide and 1 + n_precursor products";
      for (int k = 0; !bad && k < d[i].n_fission_products; ++k)
      {
        const mcw_fission_product& f = d[i].fission_products[k];
        HostFis F;
        F.law = f.law;
        F.u = f.u;
        bad = take_fn(f.yield, F.yield, true);
        if (!bad && f.law == MCW_LAW_CONTINUOUS_TABULAR)
        {
          bad = take_table(f.energy, F.energy, true);
          if (!bad && F.energy.energy.empty())
            bad = "fission product: continuous tabular law without its table";
        }
        else if (!bad && (f.law == MCW_LAW_WATT || f.law == MCW_LAW_MAXWELL || f.law == MCW_LAW_EVAPORATION))
        {
          bad = take_fn(f.a, F.a, true);
          if (!bad && f.law == MCW_LAW_WATT)
            bad = take_fn(f.b, F.b, true);
        }
        else if (!bad)
          // OMC secondary_uncorrelated.cpp:47 "Energy distribution type '...' not implemented."
          bad = "fission product: energy law not implemented (Watt, Maxwell, evaporation, continuous tabular)";
        N.fis.push_back(std::move(F));
      }
      if (!bad)
        bad = take_fn(d[i].total_nu, N.total_nu, false);
    }
    if (bad)
      return fail(h, MCW_E_INVALID_ARGUMENT, std::string("mcw_set_library: ") + bad);
    nuclides.push_back(std::move(N));
  }
  return MCW_OK;
}

extern "C" {

int mcw_set_library(mcw_handle h, const mcw_nuclide_desc* d, int n)
{
  if (!h || !d || n <= 0)
    return MCW_E_INVALID_ARGUMENT;
  const int rc = parse_library(d, n, h->nuclides, h->err);
  if (rc)
  {
    h->nuclides.clear();
    h->n_identities = 0;
    return rc;
  }
  h->n_identities = n;
  for (int i = 0; i < n; ++i)
  {
    h->nuclides[i].identity = i;
    h->nuclides[i].sets.assign(1, i);
  }
  h->have_lib = true;
  h->prepared = false;
  return MCW_OK;
}

int mcw_add_temperature_set(mcw_handle h, int nuclide, const mcw_nuclide_desc* set, int* entry)
{
  if (!h || !set || !h->have_lib)
    return MCW_E_INVALID_ARGUMENT;
  if (nuclide < 0 || nuclide >= h->n_identities)
    return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_add_temperature_set: no such nuclide");
  std::vector<HostNuclide> one;
  const int rc = parse_library(set, 1, one, h->err);
  if (rc)
    return rc;
  const HostNuclide& base = h->nuclides[nuclide];
  if ((one[0].prm.n_inel != 0) != (base.prm.n_inel != 0) || one[0].more.size() != base.more.size() ||
      one[0].prm.fissionable != base.prm.fissionable || one[0].fis.size() != base.fis.size())
    // every reaction has a cross section at every temperature (OMC reaction.cpp:25-50)
    return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_add_temperature_set: the data set must hold the same reactions as the nuclide");
  for (int e : base.sets)
    if (h->nuclides[e].prm.kT == one[0].prm.kT)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_add_temperature_set: the nuclide already has a data set at this temperature");
  one[0].identity = nuclide;
  const int e = (int)h->nuclides.size();
  h->nuclides.push_back(std::move(one[0]));
  std::vector<int>& sets = h->nuclides[nuclide].sets;
  sets.push_back(e);
  std::stable_sort(sets.begin(), sets.end(), [&](int a, int b) { return h->nuclides[a].prm.kT < h->nuclides[b].prm.kT; });
  if (entry)
    *entry = e;
  h->prepared = false;
  return MCW_OK;
}

} // extern "C"

// ---- mcw_digest: FNV-1a over everything a descriptor array says, arrays included, in declaration order (host only)
namespace
{
struct Fnv
{
  uint64_t h = 1469598103934665603ull;
  void bytes(const void* p, size_t n)
  {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i)
      h = (h ^ b[i]) * 1099511628211ull;
  }
  template <typename T>
  void val(T v) { bytes(&v, sizeof(v)); }
  template <typename T>
  void arr(const T* p, size_t n)
  {
    val((uint64_t)(p ? n : 0));
    if (p && n)
      bytes(p, n * sizeof(T));
  }
  void table(const mcw_tabular2d& t)
  {
    val(t.n_energy);
    if (t.n_energy <= 0)
      return;
    val(t.histogram_in_energy);
    const size_t n = (size_t)t.n_energy, np = (size_t)t.offset[n];
    arr(t.energy, n), arr(t.offset, n + 1), arr(t.interpolation, n);
    for (size_t k = 0; k < n; ++k) // an absent n_discrete means zeros
      val((int32_t)(t.n_discrete ? t.n_discrete[k] : 0));
    arr(t.x, np), arr(t.pdf, np), arr(t.cdf, np);
  }
  void fn(const mcw_function1d& f)
  {
    val(f.kind);
    if (f.kind == MCW_FN_NONE)
      return;
    val(f.n);
    if (f.kind == MCW_FN_TABULATED)
      val(f.interpolation), arr(f.x, (size_t)f.n);
    arr(f.y, (size_t)f.n);
  }
  // the law part shared by the inline reaction and mcw_reaction_desc
  void law(int law, int in_cm, const mcw_tabular2d& angle, const mcw_tabular2d& energy, const int32_t* co, const int32_t* ci,
           const double* cm, const double* cp, const double* cc, const double* kr, const double* ka, int nbody_n,
           double nbody_mass, double nbody_q, double yield, const mcw_function1d& yfn, double u, const mcw_function1d& a,
           const mcw_function1d& b)
  {
    val(law), val((int32_t)(law == MCW_LAW_LEVEL ? 1 : (in_cm != 0)));
    const bool together = law >= MCW_LAW_CORRELATED && law <= MCW_LAW_NBODY;
    if (!together)
      table(angle);
    if (law == MCW_LAW_CONTINUOUS_TABULAR || law == MCW_LAW_CORRELATED || law == MCW_LAW_KALBACH_MANN)
      table(energy);
    const size_t np = energy.n_energy > 0 ? (size_t)energy.offset[energy.n_energy] : 0;
    if (law == MCW_LAW_CORRELATED)
      arr(co, np + 1), arr(ci, np), arr(cm, (size_t)co[np]), arr(cp, (size_t)co[np]), arr(cc, (size_t)co[np]);
    if (law == MCW_LAW_KALBACH_MANN)
      arr(kr, np), arr(ka, np);
    if (law == MCW_LAW_NBODY)
      val(nbody_n), val(nbody_mass), val(nbody_q);
    if (yfn.kind != MCW_FN_NONE)
      fn(yfn);
    else
      val(yield == 0.0 ? 1.0 : yield);
    if (law >= MCW_LAW_MAXWELL)
    {
      val(u), fn(a);
      if (law == MCW_LAW_WATT)
        fn(b);
    }
  }
};
} // namespace

extern "C" {

int mcw_digest(int kind, const void* descriptors, int n, uint64_t* out)
{
  if ((!descriptors && n > 0) || n < 0 || !out || kind < 0 || kind > 2)
    return MCW_E_INVALID_ARGUMENT;
  Fnv f;
  f.val((int32_t)kind), f.val((int32_t)n);
  for (int i = 0; i < n; ++i)
  {
    if (kind == 0)
    {
      const mcw_nuclide_desc& d = ((const mcw_nuclide_desc*)descriptors)[i];
      const size_t G = (size_t)d.n_grid;
      f.val(d.n_grid), f.val((int32_t)(d.fissionable != 0)), f.val(d.awr), f.val(d.kT);
      f.arr(d.energy, G), f.arr(d.xs, 5 * G), f.arr(d.elastic, G);
      f.table(d.elastic_angle);
      f.val((int32_t)(d.n_inelastic != 0));
      if (d.n_inelastic)
      {
        f.val(d.inel_threshold), f.arr(d.inelastic, G - (size_t)d.inel_threshold);
        if (d.inelastic_law == MCW_LAW_LEVEL)
          f.val(d.inel_threshold_energy), f.val(d.inel_mass_ratio);
        f.law(d.inelastic_law, d.inelastic_in_cm, d.inelastic_angle, d.inelastic_energy, d.corr_offset, d.corr_interp, d.corr_mu,
              d.corr_pdf, d.corr_cdf, d.km_r, d.km_a, d.nbody_n, d.nbody_mass_ratio, d.nbody_q, d.inelastic_yield,
              d.inelastic_yield_fn, d.inelastic_u, d.inelastic_a, d.inelastic_b);
        f.val(d.n_more_reactions);
        for (int j = 0; j < d.n_more_reactions; ++j)
        {
          const mcw_reaction_desc& r = d.more_reactions[j];
          f.val(r.mt), f.val(r.threshold), f.arr(r.xs, G - (size_t)r.threshold);
          if (r.law == MCW_LAW_LEVEL)
            f.val(r.threshold_energy), f.val(r.mass_ratio);
          f.law(r.law, r.in_cm, r.angle, r.energy, r.corr_offset, r.corr_interp, r.corr_mu, r.corr_pdf, r.corr_cdf, r.km_r, r.km_a,
                r.nbody_n, r.nbody_mass_ratio, r.nbody_q, r.yield, r.yield_fn, r.u, r.a, r.b);
        }
      }
      if (d.fissionable)
      {
        f.val(d.q_fission), f.val(d.n_precursor), f.val(d.n_fission_products);
        if (d.n_fission_products > 0)
        {
          f.fn(d.total_nu);
          for (int k = 0; k < d.n_fission_products; ++k)
          {
            const mcw_fission_product& p = d.fission_products[k];
            f.fn(p.yield), f.val(p.law);
            if (p.law == MCW_LAW_CONTINUOUS_TABULAR)
              f.table(p.energy);
            else
            {
              f.val(p.u), f.fn(p.a);
              if (p.law == MCW_LAW_WATT)
                f.fn(p.b);
            }
          }
        }
        else
        {
          f.val(d.nu_poly[0]), f.val(d.nu_poly[1]), f.val(d.watt_a), f.val(d.watt_b);
          for (int g = 0; g < d.n_precursor; ++g)
            f.val(d.delayed_yield[g]), f.val(d.delayed_theta[g]);
        }
      }
    }
    else if (kind == 1)
    {
      const mcw_thermal_desc& t = ((const mcw_thermal_desc*)descriptors)[i];
      f.val(t.energy_max), f.val(t.inelastic_type), f.val(t.elastic_type), f.val(t.n_mu);
      f.arr(t.inelastic_xs_energy, (size_t)t.n_inelastic_xs), f.arr(t.inelastic_xs, (size_t)t.n_inelastic_xs);
      if (t.inelastic_type == MCW_SAB_INELASTIC_CONTINUOUS)
      {
        const size_t ne = (size_t)t.n_inelastic_energy, np = (size_t)t.inelastic_offset[ne];
        f.arr(t.inelastic_energy, ne), f.arr(t.inelastic_offset, ne + 1);
        f.arr(t.inelastic_e_out, np), f.arr(t.inelastic_pdf, np), f.arr(t.inelastic_cdf, np), f.arr(t.inelastic_mu, np * (size_t)t.n_mu);
      }
      else
      {
        const size_t np = (size_t)t.n_inelastic_xs * (size_t)t.n_e_out;
        f.val(t.n_e_out), f.val((int32_t)(t.skewed != 0));
        f.arr(t.inelastic_e_out, np), f.arr(t.inelastic_mu, np * (size_t)t.n_mu);
      }
      if (t.elastic_type != MCW_SAB_ELASTIC_NONE)
      {
        f.arr(t.elastic_x, (size_t)t.n_elastic), f.arr(t.elastic_y, (size_t)t.n_elastic);
        if (t.elastic_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
          f.val(t.n_elastic_mu), f.arr(t.elastic_mu, (size_t)t.n_elastic * (size_t)t.n_elastic_mu);
      }
    }
    else
    {
      const mcw_urr_desc& u = ((const mcw_urr_desc*)descriptors)[i];
      f.val(u.nuclide), f.val(u.interpolation), f.val(u.inelastic_flag), f.val(u.absorption_flag), f.val((int32_t)(u.multiply_smooth != 0));
      f.val(u.n_energy), f.val(u.n_band);
      f.arr(u.energy, (size_t)u.n_energy), f.arr(u.table, (size_t)u.n_energy * 6 * (size_t)u.n_band);
    }
  }
  *out = f.h;
  return MCW_OK;
}

int mcw_validate_library(const mcw_nuclide_desc* d, int n, char* message, int message_bytes)
{
  if (!d || n <= 0)
    return MCW_E_INVALID_ARGUMENT;
  std::vector<HostNuclide> scratch;
  std::string err;
  const int rc = parse_library(d, n, scratch, err);
  if (message && message_bytes > 0)
  {
    strncpy(message, err.c_str(), (size_t)message_bytes - 1);
    message[message_bytes - 1] = 0;
  }
  return rc;
}


int mcw_set_materials(mcw_handle h, const mcw_material_desc* m, int n, const int32_t* block_ids,
                      const int32_t* block_material, int n_blocks)
{
  if (!h || !m || n <= 0 || !block_ids || !block_material || n_blocks <= 0)
    return MCW_E_INVALID_ARGUMENT;
  h->mat_nuc.clear();
  h->mat_dens.clear();
  h->mat_iso.clear();
  h->mat_sab.clear(); // S(a,b) assignments belong to the materials they were made for
  h->mat_sab_frac.clear();
  for (int i = 0; i < n; ++i)
  {
    if (m[i].n_nuclides <= 0 || m[i].n_nuclides > MCW_MAX_NUC_PER_MAT)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_materials: bad nuclide count");
    h->mat_nuc.emplace_back(m[i].nuclide, m[i].nuclide + m[i].n_nuclides);
    h->mat_dens.emplace_back(m[i].atom_density, m[i].atom_density + m[i].n_nuclides);
    if (m[i].isotropic_in_lab)
      h->mat_iso.emplace_back(m[i].isotropic_in_lab, m[i].isotropic_in_lab + m[i].n_nuclides);
    else
      h->mat_iso.emplace_back((size_t)m[i].n_nuclides, 0);
  }
  h->block_ids.assign(block_ids, block_ids + n_blocks);
  h->block_material.assign(block_material, block_material + n_blocks);
  for (int b = 0; b < n_blocks; ++b)
    if (block_material[b] < 0 || block_material[b] >= n)
      // MC CollisionKernel.C:84-90 "Could not find material ... in OpenMC materials."
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_materials: block mapped to an unknown material");
  h->have_mats = true;
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_thermal_tables(mcw_handle h, const mcw_thermal_desc* d, int n)
{
  if (!h || n < 0 || (n > 0 && !d))
    return MCW_E_INVALID_ARGUMENT;
  std::vector<HostThermal> out;
  for (int k = 0; k < n; ++k)
  {
    const mcw_thermal_desc& s = d[k];
    const std::string at = " (thermal table " + std::to_string(k) + ")";
    if (s.n_inelastic_xs < 2 || !s.inelastic_xs_energy || !s.inelastic_xs || s.n_mu < 1 || !s.inelastic_e_out || !s.inelastic_mu)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: inelastic data missing" + at);
    HostThermal t;
    t.energy_max = s.energy_max;
    t.inel_type = s.inelastic_type;
    t.el_type = s.elastic_type;
    t.n_mu = s.n_mu;
    t.n_e_out = s.n_e_out;
    t.skewed = s.skewed;
    t.n_el_mu = s.n_elastic_mu;
    t.ixs_e.assign(s.inelastic_xs_energy, s.inelastic_xs_energy + s.n_inelastic_xs);
    t.ixs.assign(s.inelastic_xs, s.inelastic_xs + s.n_inelastic_xs);
    if (s.inelastic_type == MCW_SAB_INELASTIC_CONTINUOUS)
    {
      if (s.n_inelastic_energy < 2 || !s.inelastic_energy || !s.inelastic_offset || !s.inelastic_pdf || !s.inelastic_cdf)
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: continuous inelastic distribution missing" + at);
      t.ien.assign(s.inelastic_energy, s.inelastic_energy + s.n_inelastic_energy);
      t.ioff.assign(s.inelastic_offset, s.inelastic_offset + s.n_inelastic_energy + 1);
      for (int l = 0; l < s.n_inelastic_energy; ++l)
        if (t.ioff[l + 1] - t.ioff[l] < 2)
          return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: a distribution needs at least two outgoing energies" + at);
      const size_t np = (size_t)t.ioff.back();
      t.eout.assign(s.inelastic_e_out, s.inelastic_e_out + np);
      t.pdf.assign(s.inelastic_pdf, s.inelastic_pdf + np);
      t.cdf.assign(s.inelastic_cdf, s.inelastic_cdf + np);
      t.mu.assign(s.inelastic_mu, s.inelastic_mu + np * s.n_mu);
    }
    else if (s.inelastic_type == MCW_SAB_INELASTIC_DISCRETE)
    {
      if (s.n_e_out < (s.skewed ? 5 : 1))
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: too few discrete outgoing energies" + at);
      const size_t np = (size_t)s.n_inelastic_xs * s.n_e_out;
      t.eout.assign(s.inelastic_e_out, s.inelastic_e_out + np);
      t.mu.assign(s.inelastic_mu, s.inelastic_mu + np * s.n_mu);
    }
    else
      return fail(h, MCW_E_UNSUPPORTED, "mcw_set_thermal_tables: unknown inelastic type" + at);
    if (s.elastic_type == MCW_SAB_ELASTIC_COHERENT || s.elastic_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
    {
      if (s.n_elastic < 2 || !s.elastic_x || !s.elastic_y)
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: elastic data missing" + at);
      t.el_x.assign(s.elastic_x, s.elastic_x + s.n_elastic);
      t.el_y.assign(s.elastic_y, s.elastic_y + s.n_elastic);
      // the samplers interpolate between the grid point below and the one ABOVE the incident energy (OMC
      // secondary_thermal.cpp:14-25,46-48): the grid has to reach energy_max or they would read past its end
      if (!(t.el_x.back() >= s.energy_max) || (s.elastic_type == MCW_SAB_ELASTIC_COHERENT && !(t.el_x.back() > s.energy_max)))
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: the elastic grid ends below energy_max" + at);
      if (s.elastic_type == MCW_SAB_ELASTIC_INCOHERENT_DISCRETE)
      {
        if (!s.elastic_mu || s.n_elastic_mu < 1)
          return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: elastic cosines missing" + at);
        t.el_mu.assign(s.elastic_mu, s.elastic_mu + (size_t)s.n_elastic * s.n_elastic_mu);
      }
    }
    else if (s.elastic_type != MCW_SAB_ELASTIC_NONE)
      // IncoherentElasticXS / IncoherentElasticAE (Debye-Waller form, OMC endf.cpp:245-258) is not carried
      return fail(h, MCW_E_UNSUPPORTED, "mcw_set_thermal_tables: unsupported elastic type" + at);
    if (!((t.inel_type == MCW_SAB_INELASTIC_CONTINUOUS ? t.ien.back() : t.ixs_e.back()) >= s.energy_max))
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_tables: the inelastic grid ends below energy_max" + at);
    out.push_back(std::move(t));
  }
  h->thermal = std::move(out);
  h->thermal_same.clear();
  h->thermal_kT.clear();
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_thermal_temperatures(mcw_handle h, const int32_t* same_table_as, const double* kT, int n)
{
  if (!h || !same_table_as || !kT || n != (int)h->thermal.size())
    return MCW_E_INVALID_ARGUMENT;
  for (int k = 0; k < n; ++k)
    if (same_table_as[k] < 0 || same_table_as[k] > k || same_table_as[same_table_as[k]] != same_table_as[k])
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_thermal_temperatures: same_table_as must name a table's first entry");
  h->thermal_same.assign(same_table_as, same_table_as + n);
  h->thermal_kT.assign(kT, kT + n);
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_urr_tables(mcw_handle h, const mcw_urr_desc* d, int n)
{
  if (!h || n < 0 || (n > 0 && !d))
    return MCW_E_INVALID_ARGUMENT;
  if (!h->have_lib)
    return fail(h, MCW_E_NOT_READY, "mcw_set_urr_tables: call mcw_set_library first");
  for (auto& nuc : h->nuclides)
    nuc.urr = HostUrr();
  for (int k = 0; k < n; ++k)
  {
    const mcw_urr_desc& u = d[k];
    if (u.nuclide < 0 || u.nuclide >= (int)h->nuclides.size() || u.n_energy < 2 || u.n_band < 1 || !u.energy || !u.table)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_urr_tables: bad table " + std::to_string(k));
    if (u.interpolation != 2)
      // log-log tables (OMC nuclide.cpp:846-884) need exp / log of the table values: not carried
      return fail(h, MCW_E_UNSUPPORTED, "mcw_set_urr_tables: only linear-linear probability tables are supported");
    HostNuclide& nuc = h->nuclides[u.nuclide];
    if (u.inelastic_flag > 0 && nuc.prm.n_inel == 0)
      // OMC nuclide.cpp:266-270
      return fail(h, MCW_E_INVALID_ARGUMENT, "Could no find inelastic reaction specified on unresolved resonance probability table.");
    for (int e = 0; e < u.n_energy; ++e)
      if (!(u.table[((size_t)e * 6 + 0) * u.n_band + u.n_band - 1] >= 1.0))
        return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_urr_tables: a cumulative probability does not reach 1");
    nuc.urr.energy.assign(u.energy, u.energy + u.n_energy);
    nuc.urr.table.assign(u.table, u.table + (size_t)u.n_energy * 6 * u.n_band);
    nuc.urr.n_band = u.n_band;
    nuc.urr.inel_flag = u.inelastic_flag > 0 ? u.inelastic_flag : -1;
    nuc.urr.multiply_smooth = u.multiply_smooth != 0;
  }
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_material_thermal(mcw_handle h, int material, int n, const int32_t* position, const int32_t* table,
                             const double* fraction)
{
  if (!h || n < 0 || (n > 0 && (!position || !table || !fraction)))
    return MCW_E_INVALID_ARGUMENT;
  if (!h->have_mats || material < 0 || material >= (int)h->mat_nuc.size())
    return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_material_thermal: unknown material (call mcw_set_materials first)");
  h->mat_sab.resize(h->mat_nuc.size());
  h->mat_sab_frac.resize(h->mat_nuc.size());
  std::vector<int32_t> tab(h->mat_nuc[material].size(), -1);
  std::vector<double> frac(h->mat_nuc[material].size(), 0.0);
  for (int k = 0; k < n; ++k)
  {
    if (position[k] < 0 || position[k] >= (int)tab.size() || table[k] < 0 || table[k] >= (int)h->thermal.size())
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_material_thermal: nuclide position or table out of range");
    tab[position[k]] = table[k];
    frac[position[k]] = fraction[k];
  }
  h->mat_sab[material] = std::move(tab);
  h->mat_sab_frac[material] = std::move(frac);
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_temperature(mcw_handle h, const double* T, int64_t n)
{
  if (!h || !T || n <= 0)
    return MCW_E_INVALID_ARGUMENT;
  // The coupled temperature picks every nuclide's nearest data set (OMC nuclide.cpp:614-626 NEAREST); with one data set
  // per nuclide it selects that one whatever it is.  The pick is made in prepare() (temperature classes, effective
  // materials), so a change of the field after prepare() prepares again.
  if (h->temperature.size() != (size_t)n || memcmp(h->temperature.data(), T, sizeof(double) * (size_t)n) != 0)
  {
    h->temperature.assign(T, T + n);
    if ((int)h->nuclides.size() > h->n_identities || !h->thermal_same.empty())
      h->prepared = false;
  }
  return MCW_OK;
}

int mcw_set_settings(mcw_handle h, const mcw_settings* s)
{
  if (!h || !s || s->n_particles <= 0)
    return MCW_E_INVALID_ARGUMENT;
  if (s->n_log_bins < 1)
    return fail(h, MCW_E_INVALID_ARGUMENT, "n_log_bins must be positive");
  h->set = *s;
  h->work_index.assign(h->nranks + 1, 0);
  mcw_calculate_work(s->n_particles, h->nranks, h->work_index.data());
  h->n_local = h->work_index[h->rank + 1] - h->work_index[h->rank];
  h->have_set = true;
  h->prepared = false;
  return MCW_OK;
}

int mcw_add_tally(mcw_handle h, int id, int estimator, int granularity, const int32_t* scores, int n_scores,
                  const double* energy_bins, int n_energy_edges)
{
  if (!h || !scores || n_scores <= 0 || n_scores > MCW_N_SCORES)
    return MCW_E_INVALID_ARGUMENT;
  if (estimator != MCW_ESTIMATOR_COLLISION && estimator != MCW_ESTIMATOR_TRACKLENGTH)
    // MC OpenMCTally.C:89-90 "Only collision estimator currently supported" (+ track-length extension)
    return fail(h, MCW_E_UNSUPPORTED, "estimator must be collision or tracklength");
  if (granularity < 0 || granularity > MCW_FILTER_GLOBAL)
    return fail(h, MCW_E_INVALID_ARGUMENT, "unknown granularity");
  if ((int)h->tallies.size() >= MCW_MAX_TALLIES)
    return fail(h, MCW_E_UNSUPPORTED, "too many tallies");
  for (auto& t : h->tallies)
    if (t.id == id)
      return fail(h, MCW_E_INVALID_ARGUMENT, "duplicate tally id");
  HostTally t;
  t.id = id;
  t.estimator = estimator;
  t.granularity = granularity;
  t.n_scores = n_scores;
  for (int i = 0; i < n_scores; ++i)
  {
    if (scores[i] < 0 || scores[i] >= MCW_N_SCORES)
      return fail(h, MCW_E_INVALID_ARGUMENT, "unknown score");
    t.scores[i] = scores[i];
  }
  t.n_energy = 1;
  if (energy_bins && n_energy_edges >= 2)
  {
    t.ebins.assign(energy_bins, energy_bins + n_energy_edges);
    t.n_energy = n_energy_edges - 1;
  }
  h->tallies.push_back(t);
  h->prepared = false;
  return MCW_OK;
}

int mcw_set_tally_nuclides(mcw_handle h, int id, const int32_t* nuclides, int n)
{
  if (!h || !nuclides || n <= 0)
    return MCW_E_INVALID_ARGUMENT;
  for (auto& t : h->tallies)
    if (t.id == id)
    {
      for (int j = 0; j < n; ++j)
        if (nuclides[j] < -2 || nuclides[j] >= (h->have_lib ? h->n_identities : (int)h->nuclides.size()))
          // OMC tally.cpp:552-554 "Could not find the nuclide ... specified in tally ... in any material"
          return fail(h, MCW_E_INVALID_ARGUMENT, "Could not find the nuclide with index " + std::to_string(nuclides[j]) +
                                                     " specified in tally " + std::to_string(id) + " in the library");
      t.nuclides.assign(nuclides, nuclides + n);
      h->prepared = false;
      return MCW_OK;
    }
  return fail(h, MCW_E_INVALID_ARGUMENT, "unknown tally id");
}

// ------------------------------------------------------------------ prepare: build device state
static int prepare(mcw_handle h)
{
  if (h->prepared)
    return MCW_OK;
  if (!h->have_mesh || !h->have_lib || !h->have_mats || !h->have_set)
    return fail(h, MCW_E_NOT_READY, "mesh, library, materials and settings must be set first");
  CK(cudaSetDevice(h->device));
  for (void* p : h->allocs)
    cudaFree(p);
  h->allocs.clear();
  McwParams& P = h->P;
  memset(&P, 0, sizeof(P));
  TopoRT t = topo_of(h->elem_type);
  const int n_nuc = (int)h->nuclides.size(), n_id = h->n_identities;
  for (auto& mn : h->mat_nuc)
    for (int v : mn)
      if (v < 0 || v >= n_id)
        return fail(h, MCW_E_INVALID_ARGUMENT, "material refers to a nuclide outside the library");
  // ---- temperatures.  Every nuclide may have several data sets; the particle's temperature -- the coupled temperature of
  // the element it is in (MC CollisionKernel.C:109: sqrtkT = sqrt(K_BOLTZMANN * T)) -- picks the nearest one per nuclide
  // (OMC nuclide.cpp:614-626).  Elements whose temperature picks the same set for every nuclide form a class; a material
  // in a class is an "effective material" whose entries name the picked data sets, and the kernels see nothing but
  // effective materials: effective index = class * n_materials + material, class 0 = the class of the first temperature
  // value (what the stand-alone lookups use).  One data set per nuclide: one class, the materials as given.
  struct
  {
    std::vector<std::vector<int32_t>> nuc, iso, sab;
    std::vector<std::vector<double>> dens, sab_frac;
    std::vector<int32_t> elem_class; // per element, empty = all elements in class 0
    int n_class = 1;
  } EM;
  {
    const int n_mat0 = (int)h->mat_nuc.size();
    bool multi = false;
    for (int i = 0; i < n_id; ++i)
      multi |= h->nuclides[i].sets.size() > 1;
    // temperatures of the S(a,b) tables: the entries of every table (named by its first entry) by increasing kT
    const int n_th = (int)h->thermal.size();
    std::vector<std::vector<int>> th_sets(n_th);
    for (int k = 0; k < (int)h->thermal_same.size(); ++k)
      th_sets[h->thermal_same[k]].push_back(k);
    for (auto& ts : th_sets)
    {
      std::stable_sort(ts.begin(), ts.end(), [&](int a, int b) { return h->thermal_kT[a] < h->thermal_kT[b]; });
      multi |= ts.size() > 1;
    }
    const int64_t nT = (int64_t)h->temperature.size();
    if (nT > 1 && nT != h->n_elem)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_set_temperature: one value or one per element");
    std::vector<std::vector<int>> class_sel; // per class, the picked entry of every nuclide
    auto pick_all = [&](double T) {
      const double sqrtkT = std::sqrt(8.617333262e-5 * T); // OMC constants.h:90 K_BOLTZMANN
      const double kT = sqrtkT * sqrtkT;
      std::vector<int> sel(n_id + n_th);
      for (int t = 0; t < n_th; ++t)
      {
        // ThermalScattering::calculate_xs (OMC thermal.cpp:149-171): the closer of the two temperatures that bound kT
        const std::vector<int>& ts = th_sets[t];
        size_t i = 0;
        if (ts.size() > 1)
        {
          while (h->thermal_kT[ts[i + 1]] < kT && i + 1 < ts.size() - 1)
            ++i;
          if (kT - h->thermal_kT[ts[i]] > h->thermal_kT[ts[i + 1]] - kT)
            ++i;
        }
        sel[n_id + t] = ts.empty() ? t : ts[i];
      }
      for (int i = 0; i < n_id; ++i)
      {
        const auto& sets = h->nuclides[i].sets;
        int i_temp = 0;
        double max_diff = INFINITY;
        for (size_t t = 0; t < sets.size(); ++t)
        {
          const double diff = std::fabs(h->nuclides[sets[t]].prm.kT - kT);
          if (diff < max_diff)
          {
            i_temp = (int)t;
            max_diff = diff;
          }
        }
        sel[i] = sets.empty() ? i : sets[i_temp];
      }
      return sel;
    };
    if (!multi)
      class_sel.push_back(pick_all(0.0));
    else
    {
      std::map<double, int> class_of_T;
      std::map<std::vector<int>, int> class_of_sel;
      auto class_for = [&](double T) {
        auto it = class_of_T.find(T);
        if (it != class_of_T.end())
          return it->second;
        std::vector<int> sel = pick_all(T);
        auto js = class_of_sel.find(sel);
        int c;
        if (js == class_of_sel.end())
        {
          c = (int)class_sel.size();
          class_of_sel[sel] = c;
          class_sel.push_back(std::move(sel));
        }
        else
          c = js->second;
        class_of_T[T] = c;
        return c;
      };
      if (nT <= 1)
        class_for(nT == 1 ? h->temperature[0] : 0.0);
      else
      {
        EM.elem_class.resize(h->n_elem);
        for (int64_t e = 0; e < h->n_elem; ++e)
          EM.elem_class[e] = class_for(h->temperature[e]);
      }
    }
    EM.n_class = (int)class_sel.size();
    for (int c = 0; c < EM.n_class; ++c)
      for (int m = 0; m < n_mat0; ++m)
      {
        std::vector<int32_t> nuc(h->mat_nuc[m].size());
        for (size_t q = 0; q < nuc.size(); ++q)
          nuc[q] = class_sel[c][h->mat_nuc[m][q]];
        EM.nuc.push_back(std::move(nuc));
        EM.dens.push_back(h->mat_dens[m]);
        EM.iso.push_back(h->mat_iso[m]);
        std::vector<int32_t> sab = (size_t)m < h->mat_sab.size() ? h->mat_sab[m] : std::vector<int32_t>();
        for (int32_t& ts : sab) // the table's temperature for this class
          if (ts >= 0 && ts < n_th)
            ts = class_sel[c][n_id + ts];
        EM.sab.push_back(std::move(sab));
        EM.sab_frac.push_back((size_t)m < h->mat_sab_frac.size() ? h->mat_sab_frac[m] : std::vector<double>());
      }
    // every data set starts at an even grid point of the concatenated grids (16-byte aligned windows, mcw_xs.cuh)
    int64_t off = 0;
    for (auto& n : h->nuclides)
    {
      n.prm.off = off;
      off += n.prm.G + (n.prm.G & 1);
    }
  }
  const int n_mat0 = (int)h->mat_nuc.size();
  const int n_mat = (int)EM.nuc.size();
  {
    // which build of the kernels this problem needs (mcw_kernels.h)
    bool rare = h->set.record_traces == 2;
    for (size_t m = 0; m < EM.sab.size(); ++m)
      for (int32_t ts : EM.sab[m])
        rare |= ts >= 0;
    for (const auto& tl : h->tallies)
      rare |= !tl.nuclides.empty();
    for (const auto& n : h->nuclides)
      rare |= !n.el_angle.energy.empty() || !n.inel_angle.energy.empty() || !n.inel_energy.energy.empty() || !n.urr.energy.empty() ||
              (n.prm.n_inel && (n.prm.inel_law >= MCW_LAW_CORRELATED || !n.more.empty() || n.inel_yield != 1.0 ||
                                n.inel_yield_fn.kind != MCW_FN_NONE)) ||
              !n.fis.empty();
    h->rare = rare || h->force_rare > 0;
  }

  // ---- library: OMC simulation.cpp:553-619 initialize_data + nuclide.cpp:446-476 init_grid
  const int M = h->set.n_log_bins;
  double E_min = 0.0, E_max = INFINITY;
  for (int i = 0; i < n_id; ++i)
  {
    // every nuclide's FIRST (lowest-temperature) grid bounds the energy range (OMC simulation.cpp:558-566 grid_[0])
    const HostNuclide& n = h->nuclides[h->nuclides[i].sets.empty() ? i : h->nuclides[i].sets[0]];
    E_min = std::max(E_min, n.energy.front());
    E_max = std::min(E_max, n.energy.back());
  }
  const double spacing = mcw_log(E_max / E_min) / M;
  std::vector<double> energy, xs4, el_inel;
  std::vector<int32_t> grid_index;
  std::vector<McwNucParams> prm;
  std::vector<double> tab_energy, tab_x, tab_p, tab_c;
  std::vector<int32_t> tab_off, tab_end, tab_interp, tab_ndisc;
  auto add_table = [&](const HostTable& t, int32_t& first, int32_t& count) {
    first = (int32_t)tab_energy.size();
    count = (int32_t)t.energy.size();
    const int32_t base = (int32_t)tab_x.size();
    for (size_t k = 0; k < t.energy.size(); ++k)
    {
      tab_energy.push_back(t.energy[k]);
      tab_off.push_back(base + t.off[k]);
      tab_end.push_back(base + t.off[k + 1]);
      tab_interp.push_back(t.interp[k]);
      tab_ndisc.push_back(t.ndisc[k]);
    }
    tab_x.insert(tab_x.end(), t.x.begin(), t.x.end());
    tab_p.insert(tab_p.end(), t.p.begin(), t.p.end());
    tab_c.insert(tab_c.end(), t.c.begin(), t.c.end());
  };
  std::vector<McwRxParams> rx;
  std::vector<double> rx_xs;
  std::vector<McwFisProduct> fis;
  std::vector<double> fn_pool;
  auto add_fn = [&](const HostFn& f) {
    McwFn o;
    o.kind = f.kind, o.n = (int32_t)f.y.size(), o.interp = f.interp, o.off = (int32_t)fn_pool.size();
    fn_pool.insert(fn_pool.end(), f.x.begin(), f.x.end()); // tabulated: x[n] then y[n]; polynomial: y[n]
    fn_pool.insert(fn_pool.end(), f.y.begin(), f.y.end());
    return o;
  };
  std::vector<double> urr_pool;
  int n_urr = 0;
  for (auto& n : h->nuclides)
  {
    // probability tables: energies, then the table
    n.prm.urr_first = (int32_t)urr_pool.size();
    n.prm.urr_n = (int32_t)n.urr.energy.size();
    n.prm.urr_bands = n.urr.n_band;
    n.prm.urr_inel_flag = n.urr.inel_flag;
    n.prm.urr_multiply_smooth = n.urr.multiply_smooth;
    n.prm.identity = n.identity;
    urr_pool.insert(urr_pool.end(), n.urr.energy.begin(), n.urr.energy.end());
    urr_pool.insert(urr_pool.end(), n.urr.table.begin(), n.urr.table.end());
    n_urr += n.prm.urr_n > 0;
  }
  for (auto& n : h->nuclides)
  {
    add_table(n.el_angle, n.prm.el_ang_first, n.prm.el_ang_n);
    add_table(n.inel_angle, n.prm.inel_ang_first, n.prm.inel_ang_n);
    add_table(n.inel_energy, n.prm.inel_en_first, n.prm.inel_en_n);
    n.prm.inel_en_hist = n.inel_energy.hist;
    {
      int32_t count;
      add_table(n.law, n.prm.law_first, count);
    }
    // fission neutrons as a nuclide file has them: functions into fn_pool, continuous tabular spectra into the tab_* pools
    n.prm.fis_first = n.fis.empty() ? -1 : (int32_t)fis.size();
    n.prm.fis_pad = 0;
    n.prm.total_nu = add_fn(n.total_nu);
    for (const HostFis& f : n.fis)
    {
      McwFisProduct F;
      memset(&F, 0, sizeof(F));
      F.yield = add_fn(f.yield), F.a = add_fn(f.a), F.b = add_fn(f.b);
      F.u = f.u;
      F.law = f.law;
      add_table(f.energy, F.en_first, F.en_n);
      F.en_hist = f.energy.hist;
      fis.push_back(F);
    }
    // the scattering reactions: the inline one again, then the further ones with their tables and cross sections
    n.prm.rx_first = (int32_t)rx.size();
    n.prm.n_rx_more = (int32_t)n.more.size();
    if (n.prm.n_inel)
    {
      McwRxParams r0;
      memset(&r0, 0, sizeof(r0));
      r0.level_E = n.prm.inel_E, r0.level_ratio = n.prm.inel_ratio, r0.yield = n.inel_yield;
      r0.nbody_mass_ratio = n.prm.nbody_mass_ratio, r0.nbody_q = n.prm.nbody_q, r0.nbody_n = n.prm.nbody_n;
      r0.xs_off = -1;
      r0.thr = n.prm.inel_thr, r0.law = n.prm.inel_law, r0.in_cm = n.prm.inel_in_cm;
      r0.ang_first = n.prm.inel_ang_first, r0.ang_n = n.prm.inel_ang_n;
      r0.en_first = n.prm.inel_en_first, r0.en_n = n.prm.inel_en_n, r0.en_hist = n.prm.inel_en_hist;
      r0.law_first = n.prm.law_first;
      r0.yield_fn = add_fn(n.inel_yield_fn), r0.a = add_fn(n.inel_fa), r0.b = add_fn(n.inel_fb), r0.u = n.inel_fu;
      rx.push_back(r0);
      for (auto& m : n.more)
      {
        int32_t count;
        add_table(m.angle, m.prm.ang_first, m.prm.ang_n);
        add_table(m.energy, m.prm.en_first, m.prm.en_n);
        m.prm.en_hist = m.energy.hist;
        add_table(m.law, m.prm.law_first, count);
        m.prm.yield_fn = add_fn(m.yield_fn), m.prm.a = add_fn(m.fa), m.prm.b = add_fn(m.fb), m.prm.u = m.fu;
        m.prm.xs_off = (int64_t)rx_xs.size();
        rx_xs.insert(rx_xs.end(), m.xs.begin(), m.xs.end());
        rx.push_back(m.prm);
      }
    }
    const int G = n.prm.G;
    energy.insert(energy.end(), n.energy.begin(), n.energy.end());
    for (int k = 0; k < G; ++k)
    {
      // a non-fissionable nuclide contributes no fission / nu-fission whatever its table holds (OMC
      // nuclide.cpp:680-689 `if (fissionable_)`): the device rows carry zeros there, so kernels may interpolate
      // all four columns unconditionally
      for (int c = 0; c < 4; ++c)
        xs4.push_back((c >= 2 && !n.prm.fissionable) ? 0.0 : n.xs5[5 * (size_t)k + c]);
      el_inel.push_back(n.elastic[k]);
      el_inel.push_back(n.inelastic_full[k]);
    }
    if (G & 1)
    {
      // padding point so the next nuclide's offset is even (never indexed by a lookup)
      energy.push_back(n.energy.back());
      xs4.insert(xs4.end(), 4, 0.0);
      el_inel.insert(el_inel.end(), 2, 0.0);
    }
    int j = 0;
    const double step = (M * spacing - 0.0) / M;
    for (int k = 0; k <= M; ++k)
    {
      double umesh = 0.0 + k * step;
      while (mcw_log(n.energy[j + 1] / E_min) <= umesh)
      {
        if (j + 2 == G)
          break;
        ++j;
      }
      grid_index.push_back(j);
    }
    prm.push_back(n.prm);
  }
  std::vector<int32_t> mat_off(1, 0), mat_nuc, mat_fis, mat_iso;
  std::vector<double> mat_dens, mat_e0, mat_e1;
  for (int m = 0; m < n_mat; ++m)
  {
    bool fis = false;
    for (size_t i = 0; i < EM.nuc[m].size(); ++i)
    {
      mat_nuc.push_back(EM.nuc[m][i]);
      mat_dens.push_back(EM.dens[m][i]);
      mat_iso.push_back(EM.iso[m][i] != 0);
      mat_e0.push_back(h->nuclides[EM.nuc[m][i]].energy.front());
      mat_e1.push_back(h->nuclides[EM.nuc[m][i]].energy.back());
      fis |= h->nuclides[EM.nuc[m][i]].prm.fissionable != 0;
    }
    mat_off.push_back((int32_t)mat_nuc.size());
    mat_fis.push_back(fis);
  }
  energy.insert(energy.end(), 2, energy.back()); // staged windows are copied in pairs: one pair of slack at the end
  xs4.insert(xs4.end(), 8, 0.0);
  double *d_energy, *d_xs4, *d_el, *d_dens;
  int32_t *d_gi, *d_moff, *d_mnuc, *d_mfis;
  McwNucParams* d_prm;
  CK(dev_upload(h, &d_energy, energy));
  CK(dev_upload(h, &d_xs4, xs4));
  CK(dev_upload(h, &d_el, el_inel));
  CK(dev_upload(h, &d_gi, grid_index));
  CK(dev_upload(h, &d_prm, prm));
  CK(dev_upload(h, &d_moff, mat_off));
  CK(dev_upload(h, &d_mnuc, mat_nuc));
  CK(dev_upload(h, &d_dens, mat_dens));
  int32_t* d_miso;
  CK(dev_upload(h, &d_miso, mat_iso));
  P.lib.mat_iso = d_miso;
  double *d_me0, *d_me1;
  CK(dev_upload(h, &d_me0, mat_e0));
  CK(dev_upload(h, &d_me1, mat_e1));
  P.lib.mat_e0 = d_me0;
  P.lib.mat_e1 = d_me1;
  CK(dev_upload(h, &d_mfis, mat_fis));
  {
    double *d_te, *d_tx, *d_tp, *d_tc;
    int32_t *d_to, *d_tn, *d_ti, *d_td;
    CK(dev_upload(h, &d_te, tab_energy));
    CK(dev_upload(h, &d_tx, tab_x));
    CK(dev_upload(h, &d_tp, tab_p));
    CK(dev_upload(h, &d_tc, tab_c));
    CK(dev_upload(h, &d_to, tab_off));
    CK(dev_upload(h, &d_tn, tab_end));
    CK(dev_upload(h, &d_ti, tab_interp));
    CK(dev_upload(h, &d_td, tab_ndisc));
    P.lib.tab_energy = d_te, P.lib.tab_x = d_tx, P.lib.tab_p = d_tp, P.lib.tab_c = d_tc;
    P.lib.tab_off = d_to, P.lib.tab_end = d_tn, P.lib.tab_interp = d_ti, P.lib.tab_ndisc = d_td;
    McwRxParams* d_rx;
    double* d_rxx;
    CK(dev_upload(h, &d_rx, rx));
    CK(dev_upload(h, &d_rxx, rx_xs));
    P.lib.rx = d_rx, P.lib.rx_xs = d_rxx;
    McwFisProduct* d_fis;
    double* d_fn;
    CK(dev_upload(h, &d_fis, fis));
    CK(dev_upload(h, &d_fn, fn_pool));
    P.lib.fis = d_fis, P.lib.fn_pool = d_fn;
  }
  {
    // thermal scattering: tables into one pool, per material entry its table and bound fraction
    std::vector<McwSabDev> sab;
    std::vector<double> pool;
    std::vector<int32_t> ipool, mat_sab;
    std::vector<double> mat_sab_frac;
    auto put = [&](const std::vector<double>& v) {
      const int32_t at = (int32_t)pool.size();
      pool.insert(pool.end(), v.begin(), v.end());
      return at;
    };
    for (const HostThermal& t : h->thermal)
    {
      McwSabDev d{};
      d.energy_max = t.energy_max;
      d.inel_type = t.inel_type, d.el_type = t.el_type;
      d.n_ixs = (int32_t)t.ixs_e.size(), d.n_ien = (int32_t)t.ien.size();
      d.n_mu = t.n_mu, d.n_e_out = t.n_e_out, d.skewed = t.skewed;
      d.n_el = (int32_t)t.el_x.size(), d.n_el_mu = t.n_el_mu;
      d.ioff = (int32_t)ipool.size();
      ipool.insert(ipool.end(), t.ioff.begin(), t.ioff.end());
      d.ixs_e = put(t.ixs_e), d.ixs = put(t.ixs), d.ien = put(t.ien), d.eout = put(t.eout), d.pdf = put(t.pdf);
      d.cdf = put(t.cdf), d.mu = put(t.mu), d.el_x = put(t.el_x), d.el_y = put(t.el_y), d.el_mu = put(t.el_mu);
      sab.push_back(d);
    }
    bool any = false;
    for (int m = 0; m < n_mat; ++m)
      for (size_t i = 0; i < EM.nuc[m].size(); ++i)
      {
        const bool have = (size_t)m < EM.sab.size() && EM.sab[m].size() == EM.nuc[m].size();
        mat_sab.push_back(have ? EM.sab[m][i] : -1);
        mat_sab_frac.push_back(have ? EM.sab_frac[m][i] : 0.0);
        if (mat_sab.back() >= (int32_t)sab.size())
          return fail(h, MCW_E_INVALID_ARGUMENT, "a material names a thermal table that is not set");
        any |= mat_sab.back() >= 0;
      }
    McwSabDev* d_sab = nullptr;
    double *d_pool = nullptr, *d_frac = nullptr;
    int32_t *d_ipool = nullptr, *d_msab = nullptr;
    if (any)
    {
      pool.push_back(0.0);
      ipool.push_back(0);
      CK(dev_upload(h, &d_sab, sab));
      CK(dev_upload(h, &d_pool, pool));
      CK(dev_upload(h, &d_ipool, ipool));
      CK(dev_upload(h, &d_msab, mat_sab));
      CK(dev_upload(h, &d_frac, mat_sab_frac));
    }
    P.lib.sab = d_sab, P.lib.sab_pool = d_pool, P.lib.sab_ipool = d_ipool;
    P.lib.mat_sab = d_msab, P.lib.mat_sab_frac = d_frac;
    P.lib.n_sab = any ? (int32_t)sab.size() : 0; // 0 switches every S(a,b) branch of the kernels off
  }
  {
    double* d_urr = nullptr;
    urr_pool.push_back(0.0);
    if (n_urr)
      CK(dev_upload(h, &d_urr, urr_pool));
    P.lib.urr_pool = d_urr;
    P.lib.n_urr = n_urr;
  }
  P.lib.energy = d_energy;
  P.lib.xs4 = d_xs4;
  P.lib.el_inel = d_el;
  P.lib.grid_index = d_gi;
  P.lib.nuc = d_prm;
  P.lib.mat_off = d_moff;
  P.lib.mat_nuc = d_mnuc;
  P.lib.mat_dens = d_dens;
  P.lib.mat_fissionable = d_mfis;
  P.lib.n_nuc = n_nuc;
  P.lib.n_identities = n_id;
  P.lib.n_mat = n_mat;
  P.lib.M = M;
  P.lib.E_min = E_min;
  P.lib.E_max = E_max;
  P.lib.log_spacing = mcw_log(E_max / E_min) / M;
  {
    // the library descriptor once more in global memory, for out-of-line device paths (McwLibraryDev::self_dev)
    McwLibraryDev* d_lib = nullptr;
    CK(dev_alloc(h, &d_lib, 1));
    P.lib.self_dev = d_lib;
    CK(cudaMemcpy(d_lib, &P.lib, sizeof(McwLibraryDev), cudaMemcpyHostToDevice));
  }

  // ---- mesh: block index / material, subdomain hmax (RT RayTracingStudy.C:497-513), records
  std::map<int32_t, int> block_index;
  for (size_t b = 0; b < h->block_ids.size(); ++b)
    block_index[h->block_ids[b]] = (int)b;
  const int n_blocks = (int)h->block_ids.size();
  std::vector<double> block_hmax(n_blocks, 0.0);
  std::vector<int32_t> eblock(h->n_elem);
  auto PX = [&](int64_t e, int v, int c) { return h->xyz[3 * h->conn[e * t.nv + v] + c]; };
  for (int64_t e = 0; e < h->n_elem; ++e)
  {
    auto it = block_index.find(h->block_raw[e]);
    if (it == block_index.end())
      return fail(h, MCW_E_INVALID_ARGUMENT, "mesh block " + std::to_string(h->block_raw[e]) + " has no material");
    eblock[e] = it->second;
    double h2 = 0;
    for (int a = 0; a < t.nv; ++a)
      for (int b = a + 1; b < t.nv; ++b)
      {
        double dx = PX(e, a, 0) - PX(e, b, 0), dy = PX(e, a, 1) - PX(e, b, 1), dz = PX(e, a, 2) - PX(e, b, 2);
        h2 = std::max(h2, dx * dx + dy * dy + dz * dz);
      }
    block_hmax[eblock[e]] = std::max(block_hmax[eblock[e]], std::sqrt(h2));
  }
  // the effective material of an element: its block's material in its temperature class
  auto emat = [&](int64_t e) { return (EM.elem_class.empty() ? 0 : EM.elem_class[e]) * n_mat0 + h->block_material[eblock[e]]; };
  const int n_bnd = (int)h->bnd_kind.size();
  if (h->n_elem > MCW_MAX_ELEMS)
    return fail(h, MCW_E_UNSUPPORTED, "more than 2^28 elements: side codes hold neighbour << 3 | back side in 32 bits");
  // cross-section slots: the materials the mesh actually uses, in material order
  std::vector<int> slot_of(n_mat, -1);
  {
    std::vector<char> used(n_mat, 0);
    for (int64_t e = 0; e < h->n_elem; ++e)
      used[emat(e)] = 1;
    int ns = 0;
    for (int m = 0; m < n_mat; ++m)
      if (used[m])
      {
        if (ns >= MCW_MAX_XS_SLOTS)
          return fail(h, MCW_E_UNSUPPORTED, "the mesh uses more than " + std::to_string(MCW_MAX_XS_SLOTS) + " materials");
        P.slot_mat[ns] = m;
        slot_of[m] = ns++;
      }
    P.n_slots = ns;
    std::vector<int32_t> pl_nuc, pl_slot;
    std::vector<double> pl_dens, pl_e0, pl_e1;
    bool every_slot = true;
    for (int sl = 0; sl < ns; ++sl)
    {
      const int m = P.slot_mat[sl];
      every_slot &= !EM.nuc[m].empty();
      for (size_t q = 0; q < EM.nuc[m].size(); ++q)
      {
        const HostNuclide& N = h->nuclides[EM.nuc[m][q]];
        pl_nuc.push_back(EM.nuc[m][q]);
        pl_slot.push_back(sl);
        pl_dens.push_back(EM.dens[m][q]);
        pl_e0.push_back(N.energy.front());
        pl_e1.push_back(N.energy.back());
      }
    }
    (void)every_slot; // mcw_set_materials rejects empty materials
    P.n_plan = (int32_t)pl_nuc.size();
    P.n_plan_pad = (P.n_plan + 3) & ~3;
    for (int sl = 0, q = 0; sl <= ns; ++sl)
    {
      P.slot_plan_off[sl] = q;
      if (sl < ns)
        q += (int)EM.nuc[P.slot_mat[sl]].size();
    }
    int32_t *d_pn, *d_ps;
    double *d_pd, *d_p0, *d_p1;
    CK(dev_upload(h, &d_pn, pl_nuc));
    CK(dev_upload(h, &d_ps, pl_slot));
    CK(dev_upload(h, &d_pd, pl_dens));
    CK(dev_upload(h, &d_p0, pl_e0));
    CK(dev_upload(h, &d_p1, pl_e1));
    P.plan_nuc = d_pn, P.plan_slot = d_ps, P.plan_dens = d_pd, P.plan_e0 = d_p0, P.plan_e1 = d_p1;
  }
  std::vector<unsigned char> rec((size_t)h->n_elem * t.stride, 0);
  for (int64_t e = 0; e < h->n_elem; ++e)
  {
    unsigned char* r = rec.data() + (size_t)e * t.stride;
    double* c = reinterpret_cast<double*>(r);
    for (int v = 0; v < t.nv; ++v)
      for (int k = 0; k < t.cpv; ++k)
        c[v * t.cpv + k] = PX(e, v, k);
    int32_t* nb = reinterpret_cast<int32_t*>(r + t.nv * t.cpv * 8);
    for (int s = 0; s < t.ns; ++s)
    {
      int64_t n = h->nbr[e * t.ns + s];
      if (n >= 0)
      {
        // RT TraceRay.C:1456-1458: the first side of the neighbour whose neighbour is this element
        int back = 0;
        for (int k = t.ns - 1; k >= 0; --k)
          if (h->nbr[n * t.ns + k] == e)
            back = k;
        nb[s] = (int32_t)((n << MCW_NBR_SHIFT) | back);
      }
      else
      {
        int32_t id = h->side_bnd[e * t.ns + s];
        if (id < 0 || id >= n_bnd)
          id = n_bnd; // no boundary condition on this side: treated as kill (RT TraceRay.C:2100-2146 errors)
        nb[s] = -1 - id;
      }
    }
    *reinterpret_cast<double*>(r + t.nv * t.cpv * 8 + t.ns * 4) = block_hmax[eblock[e]];
    *reinterpret_cast<int32_t*>(r + t.nv * t.cpv * 8 + t.ns * 4 + 8) = slot_of[emat(e)];
    *reinterpret_cast<int32_t*>(r + t.nv * t.cpv * 8 + t.ns * 4 + 12) = eblock[e];
  }
  // node -> elements CSR
  std::vector<int64_t> n2e_off(h->n_nodes + 1, 0);
  for (int64_t i = 0; i < h->n_elem * t.nv; ++i)
    n2e_off[h->conn[i] + 1]++;
  for (int64_t i = 0; i < h->n_nodes; ++i)
    n2e_off[i + 1] += n2e_off[i];
  std::vector<int32_t> n2e(n2e_off.back());
  {
    std::vector<int64_t> fill(n2e_off.begin(), n2e_off.end() - 1);
    for (int64_t e = 0; e < h->n_elem; ++e)
      for (int v = 0; v < t.nv; ++v)
        n2e[fill[h->conn[e * t.nv + v]]++] = (int32_t)e;
  }
  // bounding box, rectangular-domain flag, point-location grid
  double bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < h->n_nodes; ++i)
    for (int c = 0; c < 3; ++c)
    {
      bmin[c] = std::min(bmin[c], h->xyz[3 * i + c]);
      bmax[c] = std::max(bmax[c], h->xyz[3 * i + c]);
    }
  const int dim = h->elem_type == MCW_QUAD4 ? 2 : 3;
  double diag = std::sqrt((bmax[0] - bmin[0]) * (bmax[0] - bmin[0]) + (bmax[1] - bmin[1]) * (bmax[1] - bmin[1]) +
                          (bmax[2] - bmin[2]) * (bmax[2] - bmin[2]));
  const double dml = 1.01 * diag; // RT RayTracingStudy.C:154
  bool rectangular = true;
  {
    double tol = 1e-10 * dml;
    for (int64_t e = 0; e < h->n_elem && rectangular; ++e)
      for (int s = 0; s < t.ns && rectangular; ++s)
        if (h->nbr[e * t.ns + s] < 0)
        {
          bool on_plane = false;
          for (int d = 0; d < dim && !on_plane; ++d)
          {
            bool lo = true, hi = true;
            for (int i = 0; i < t.nps; ++i)
            {
              double c = PX(e, t.sn[s][i], d);
              lo = lo && std::fabs(c - bmin[d]) <= tol;
              hi = hi && std::fabs(c - bmax[d]) <= tol;
            }
            on_plane = lo || hi;
          }
          if (!on_plane)
            rectangular = false;
        }
  }
  int gx, gy, gz;
  {
    double n = std::max<double>(1.0, (double)h->n_elem);
    if (dim == 2)
    {
      gx = gy = std::max(1, (int)std::sqrt(n));
      gz = 1;
    }
    else
      gx = gy = gz = std::max(1, (int)std::cbrt(n));
  }
  std::vector<int64_t> cell_off((size_t)gx * gy * gz + 1, 0);
  std::vector<int32_t> cell_elems;
  {
    double ext[3] = {bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]};
    int g[3] = {gx, gy, gz};
    auto range = [&](double lo, double hi, int d, int& a, int& b) {
      if (ext[d] <= 0)
      {
        a = b = 0;
        return;
      }
      a = (int)std::floor((lo - bmin[d]) / ext[d] * g[d] - 1e-9);
      b = (int)std::floor((hi - bmin[d]) / ext[d] * g[d] + 1e-9);
      a = std::min(std::max(a, 0), g[d] - 1);
      b = std::min(std::max(b, 0), g[d] - 1);
    };
    std::vector<int64_t> cnt(cell_off.size(), 0);
    for (int pass = 0; pass < 2; ++pass)
    {
      for (int64_t e = 0; e < h->n_elem; ++e)
      {
        double lo[3], hi[3];
        for (int c = 0; c < 3; ++c)
          lo[c] = hi[c] = PX(e, 0, c);
        for (int v = 1; v < t.nv; ++v)
          for (int c = 0; c < 3; ++c)
          {
            lo[c] = std::min(lo[c], PX(e, v, c));
            hi[c] = std::max(hi[c], PX(e, v, c));
          }
        int a[3], b[3];
        for (int c = 0; c < 3; ++c)
          range(lo[c], hi[c], c, a[c], b[c]);
        for (int k = a[2]; k <= b[2]; ++k)
          for (int j = a[1]; j <= b[1]; ++j)
            for (int i = a[0]; i <= b[0]; ++i)
            {
              size_t c = ((size_t)k * gy + j) * gx + i;
              if (pass == 0)
                cnt[c + 1]++;
              else
                cell_elems[cnt[c]++] = (int32_t)e;
            }
      }
      if (pass == 0)
      {
        for (size_t c = 0; c + 1 < cnt.size(); ++c)
          cnt[c + 1] += cnt[c];
        cell_off = cnt;
        cell_elems.assign(cnt.back(), 0);
        cnt.assign(cell_off.begin(), cell_off.end());
      }
    }
  }
  unsigned char* d_rec;
  int64_t *d_conn, *d_n2e_off, *d_cell_off;
  int32_t *d_n2e, *d_bk, *d_cell_elems;
  CK(dev_upload(h, &d_rec, rec));
  CK(dev_upload(h, &d_conn, h->conn));
  CK(dev_upload(h, &d_n2e_off, n2e_off));
  CK(dev_upload(h, &d_n2e, n2e));
  CK(dev_upload(h, &d_bk, h->bnd_kind));
  CK(dev_upload(h, &d_cell_off, cell_off));
  CK(dev_upload(h, &d_cell_elems, cell_elems));
  P.mesh.rec = d_rec;
  P.mesh.conn = d_conn;
  P.mesh.n2e_off = d_n2e_off;
  P.mesh.n2e = d_n2e;
  P.mesh.bnd_kind = d_bk;
  P.mesh.cell_off = d_cell_off;
  P.mesh.cell_elems = d_cell_elems;
  P.mesh.gx = gx;
  P.mesh.gy = gy;
  P.mesh.gz = gz;
  P.mesh.type = h->elem_type;
  P.mesh.n_bnd = n_bnd;
  P.mesh.rectangular = rectangular;
  P.mesh.n_elem = h->n_elem;
  P.mesh.n_nodes = h->n_nodes;
  for (int c = 0; c < 3; ++c)
  {
    P.mesh.bb_min[c] = bmin[c];
    P.mesh.bb_max[c] = bmax[c];
  }
  P.mesh.domain_max_length = dml;

  // ---- particles, queues, banks
  const int64_t n = h->n_local;
  McwParticlesDev& p = P.p;
  double** dd[] = {&p.x, &p.y, &p.z, &p.ux, &p.uy, &p.uz, &p.E, &p.wgt, &p.k_col, &p.k_abs, &p.k_trk};
  for (auto q : dd)
    CK(dev_alloc(h, q, n));
  uint64_t** du[] = {&p.seed_trk, &p.seed_src, &p.seed_urr, &p.hash};
  for (auto q : du)
    CK(dev_alloc(h, q, n));
  int32_t** di[] = {&p.elem, &p.last_elem, &p.n_progeny, &p.n_seg, &p.n_coll, &p.n_xs};
  for (auto q : di)
    CK(dev_alloc(h, q, n));
  int8_t** db[] = {&p.inc_side, &p.ext_first, &p.ext_second, &p.status, &p.fate};
  for (auto q : db)
    CK(dev_alloc(h, q, n));
  // macroscopic cross sections of every material slot at each history's energy: 32 B per (history, slot)
  P.visit_cap = h->visit_cap;
  CK(dev_alloc(h, &p.xs, (size_t)P.n_slots * 4 * (size_t)n));
  CK(dev_alloc(h, &p.xs_prefix, (size_t)P.n_plan_pad * (size_t)n)); // 8 B per (history, plan entry)
  if (h->set.record_traces == 2)
  {
    // cached trace points (RT TraceData.h): parity / RayTracingExodus-style output, not a production mode
    double** dt[] = {&p.tx, &p.ty, &p.tz};
    for (auto q : dt)
      CK(dev_alloc(h, q, n));
    P.trace_capacity = std::min<int64_t>(2048 * n, (int64_t)1 << 23); // 104 B each: at most 0.9 GB
    CK(dev_alloc(h, &P.trace, P.trace_capacity));
  }
  {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->track_max_blocks = std::max(1, MCWK(mcwk_track_blocks_per_sm, h->elem_type, false)) * prop.multiProcessorCount;
    h->fused_max_blocks = std::max(1, MCWK(mcwk_track_blocks_per_sm, h->elem_type, true)) * prop.multiProcessorCount;
    h->pump_blocks = std::max(1, MCWK(mcwk_pump_blocks_per_sm, h->elem_type)) * prop.multiProcessorCount;
  }
  h->bank_capacity = 3 * n; // OMC simulation.cpp:286
  h->sampled_capacity = 3 * n + 16;
  CK(dev_alloc(h, &h->d_source, h->sampled_capacity));
  CK(dev_alloc(h, &h->d_bank, h->bank_capacity));
  CK(dev_alloc(h, &h->d_sorted, h->bank_capacity));
  CK(dev_alloc(h, &h->d_sampled, h->sampled_capacity));
  CK(dev_alloc(h, &h->d_q_fly[0], n));
  CK(dev_alloc(h, &h->d_q_fly[1], n));
  CK(dev_alloc(h, &h->d_q_collide, n));
  CK(dev_alloc(h, &h->d_q_fallback, n));
  CK(dev_alloc(h, &h->d_q_sorted, n + 256));
  CK(dev_alloc(h, &h->d_keys, n + 256));
  CK(dev_alloc(h, &h->d_q_hist, n + 256));
  CK(dev_alloc(h, &h->d_bins, MCWK(mcwk_refresh_bins, )));
  CK(dev_alloc(h, &h->d_bin_offsets, MCWK(mcwk_refresh_bins, )));
  CK(cudaMemset(h->d_bins, 0, MCWK(mcwk_refresh_bins, ) * sizeof(uint32_t)));
  CK(dev_alloc(h, &h->d_i64_a, std::max(h->bank_capacity, n)));
  CK(dev_alloc(h, &h->d_i64_b, std::max(h->bank_capacity, n)));
  h->temp_bytes = std::max(MCWK(mcwk_scan_temp_bytes, std::max(h->bank_capacity, n)), MCWK(mcwk_compact_temp_bytes, n));
  {
    unsigned char* tmp;
    CK(dev_alloc(h, &tmp, h->temp_bytes));
    h->d_temp = tmp;
  }
  CK(dev_alloc(h, &P.counters, 1));
  CK(cudaMemset(P.counters, 0, sizeof(McwCounters)));
  {
    // rings of the persistent generation kernel: every history is in at most one ring at a time and every
    // consumer lane holds at most one ticket, so 2 n + 2^18 slots never wrap onto a live entry
    uint64_t cap = 1;
    while (cap < 2 * (uint64_t)n + (1u << 18))
      cap <<= 1;
    McwPump& Q = h->Q;
    CK(dev_alloc(h, &Q.q_fly, cap));
    CK(dev_alloc(h, &Q.q_collide, cap));
    CK(dev_alloc(h, &Q.q_fallback, cap));
    CK(cudaMemset(Q.q_fly, 0xff, cap * 4));
    CK(cudaMemset(Q.q_collide, 0xff, cap * 4));
    CK(cudaMemset(Q.q_fallback, 0xff, cap * 4));
    unsigned long long* ctl;
    h->pump_ctl_bytes = 7 * 16 * sizeof(unsigned long long);
    CK(dev_alloc(h, &ctl, 7 * 16));
    h->d_pump_ctl = ctl;
    CK(cudaMemset(ctl, 0, h->pump_ctl_bytes));
    Q.fly_head = ctl, Q.fly_tail = ctl + 16, Q.collide_head = ctl + 32, Q.collide_tail = ctl + 48;
    Q.fallback_head = ctl + 64, Q.fallback_tail = ctl + 80;
    Q.alive = reinterpret_cast<long long*>(ctl + 96);
    Q.mask = (uint32_t)(cap - 1);
  }
  CK(dev_alloc(h, &h->d_records, n));
  CK(dev_alloc(h, &h->d_box, 6));
  P.bank = h->d_bank;
  P.bank_capacity = h->bank_capacity;

  // ---- tallies: VALUE arrays live in one buffer followed by the 5 global doubles so a single
  //      all-reduce covers everything (OMC tallies/tally.cpp:745-803 reduce_tally_results)
  int64_t total_values = 0;
  for (auto& tl : h->tallies)
  {
    tl.n_spatial =
        tl.granularity == MCW_FILTER_CELL ? h->n_elem : (tl.granularity == MCW_FILTER_UNIVERSE ? n_blocks : 1);
    tl.n_values = tl.n_spatial * tl.n_energy * tl.n_nuc_bins() * tl.n_scores;
    tl.offset = total_values;
    total_values += tl.n_values;
  }
  h->n_tally_values = total_values;
  h->n_reduce = total_values + 5;
  CK(dev_alloc(h, &h->d_reduce, h->n_reduce));
  CK(cudaMemset(h->d_reduce, 0, h->n_reduce * sizeof(double)));
  P.n_tallies = (int)h->tallies.size();
  P.any_tracklength = P.any_collision_tally = P.collision_kappa = 0;
  for (size_t i = 0; i < h->tallies.size(); ++i)
  {
    HostTally& tl = h->tallies[i];
    CK(dev_alloc(h, &tl.d_sum, tl.n_values));
    CK(dev_alloc(h, &tl.d_sum_sq, tl.n_values));
    CK(cudaMemset(tl.d_sum, 0, tl.n_values * sizeof(double)));
    CK(cudaMemset(tl.d_sum_sq, 0, tl.n_values * sizeof(double)));
    CK(dev_upload(h, &tl.d_ebins, tl.ebins));
    tl.d_nuclides = nullptr;
    if (!tl.nuclides.empty())
      CK(dev_upload(h, &tl.d_nuclides, tl.nuclides));
    tl.n_realizations = 0;
    McwTallyDev& d = P.tally[i];
    d.value = h->d_reduce + tl.offset;
    d.ebins = tl.d_ebins;
    d.nuclides = tl.d_nuclides;
    d.n_nuc_bins = tl.n_nuc_bins();
    d.n_edges = (int)tl.ebins.size();
    d.n_energy = tl.n_energy;
    d.n_scores = tl.n_scores;
    d.granularity = tl.granularity;
    d.estimator = tl.estimator;
    for (int s = 0; s < tl.n_scores; ++s)
      d.scores[s] = tl.scores[s];
    if (tl.estimator == MCW_ESTIMATOR_TRACKLENGTH)
      P.any_tracklength = 1;
    else
    {
      P.any_collision_tally = 1;
      for (int s = 0; s < tl.n_scores; ++s)
        if (tl.scores[s] == MCW_SCORE_KAPPA_FISSION)
          P.collision_kappa = 1;
    }
  }
  {
    McwTallyDev* d_tallies = nullptr;
    CK(dev_alloc(h, &d_tallies, MCW_MAX_TALLIES));
    CK(cudaMemcpy(d_tallies, P.tally, sizeof(P.tally), cudaMemcpyHostToDevice));
    P.tally_dev = d_tallies;
  }
  P.n_particles_global = h->set.n_particles;
  P.work_offset = h->work_index[h->rank];
  P.master_seed = h->set.master_seed;
  P.is_2d = h->set.is_2d;
  P.seed_through_double = h->set.seed_through_double;
  P.max_events = h->set.max_events > 0 ? h->set.max_events : 1000000;
  P.record_traces = h->set.record_traces;
  P.z_coord = h->set.z_coord;
  h->keff = 1.0;
  h->keff_std = 0.0;
  h->k_sum[0] = h->k_sum[1] = 0.0;
  h->n_realizations = 0;
  h->k_generation.clear();
  for (int i = 0; i < 4; ++i)
    h->gt_sum[i] = h->gt_sum_sq[i] = 0.0;
  h->n_source = 0;
  h->prepared = true;
  return MCW_OK;
}

int mcw_prepare(mcw_handle h)
{
  if (!h)
    return MCW_E_INVALID_ARGUMENT;
  return prepare(h);
}

// ------------------------------------------------------------------ source bank
int mcw_sample_initial_source(mcw_handle h)
{
  if (!h)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_box, h->set.source_box, 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->P.batch = 1;
  CK(cudaMemsetAsync(h->P.counters, 0, sizeof(McwCounters), h->stream));
  MCWK(mcwk_sample_source, h->P, h->d_source, h->n_local, h->d_box, h->set.source_only_fissionable, h->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_counters, h->P.counters, sizeof(McwCounters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_counters->n_failed)
    // OMC source.cpp:218-226 (EXTSRC_REJECT_THRESHOLD / _FRACTION): a box that misses the mesh or the fissionable
    // region is a fatal error in the reference, not a silently biased first generation
    return fail(h, MCW_E_PHYSICS,
                "More than 95% of external source sites sampled were rejected. Please check your external source's spatial "
                "definition. (" + std::to_string(h->h_counters->n_failed) + " of " + std::to_string(h->n_local) +
                    " sites found no position in the mesh" + (h->set.source_only_fissionable ? " in fissionable material" : "") +
                    " within 10^6 tries)");
  h->n_source = h->n_local;
  return MCW_OK;
}

int mcw_set_source_bank(mcw_handle h, const mcw_source_site* sites, int64_t n)
{
  if (!h || !sites || n < 0)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  if (n > h->n_local)
    return fail(h, MCW_E_INVALID_ARGUMENT, "source bank larger than this rank's share of n_particles");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_source, sites, n * sizeof(mcw_source_site), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n_source = n;
  return MCW_OK;
}

int mcw_get_source_bank(mcw_handle h, mcw_source_site* sites, int64_t capacity, int64_t* n)
{
  if (!h || !n)
    return MCW_E_INVALID_ARGUMENT;
  *n = h->n_source;
  if (sites && capacity > 0)
  {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(sites, h->d_source, std::min(capacity, h->n_source) * sizeof(mcw_source_site),
                  cudaMemcpyDeviceToHost));
  }
  return MCW_OK;
}

// ------------------------------------------------------------------ the event loop
int mcw_transport_generation(mcw_handle h, int batch)
{
  if (!h || batch < 1)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  if (h->n_source <= 0)
    return fail(h, MCW_E_NOT_READY, "no source bank: call mcw_sample_initial_source or mcw_set_source_bank");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  McwParams& P = h->P;
  const int64_t n = h->n_source;
  P.batch = batch;
  P.keff = h->keff;
  P.tallies_active = batch > h->set.n_inactive; // OMC simulation.cpp:318-326
  P.ordered_queues = h->ordered_queues;
  mcw_generation_result& R = h->last;
  memset(&R, 0, sizeof(R));
  CK(cudaEventRecord(h->ev[0], st));
  CK(cudaMemsetAsync(P.counters, 0, sizeof(McwCounters), st));
  CK(cudaMemsetAsync(h->d_reduce + h->n_tally_values, 0, 5 * sizeof(double), st));
  MCWK(mcwk_init_histories, P, h->d_source, n, h->d_q_fly[0], st);
  R.n_kernel_launches += 1;
  // simulation::total_weight is the sum of the SOURCE weights (OMC simulation.cpp:466 in initialize_history): reduced now,
  // before a reaction with a non-integral yield can change a history's weight
  MCWK(mcwk_reduce_sum, h->d_temp, h->temp_bytes, P.p.wgt, h->d_reduce + h->n_tally_values + 4, n, st);
  {
    unsigned long long nn = (unsigned long long)n;
    CK(cudaMemcpyAsync(&P.counters->n_fly[0], &nn, sizeof(nn), cudaMemcpyHostToDevice, st));
  }
  int cur = 0;
  int64_t upper = n;
  int64_t rounds = 0;
  float ms;
  const bool prof = h->profile != 0;
  if (prof)
    cudaEventRecord(h->ev[4], st);
  // One event round: every history in the fly queue walks elements until it collides or dies
  // (k_track; rare vertex / edge steps go through k_fallback), then every collided history is
  // processed by k_collide, which also refreshes its cross sections for all materials at the new
  // energy and puts the survivors into the other fly queue.
  // The general walk is a long single-thread step (~60 us for a TET4 vertex hit) for a handful of
  // histories per round, so its queue is only served in the round after a host sync (every round
  // once few histories are left, and in fused rounds); until then those histories wait in
  // MCW_ST_FALLBACK.  fb_pending is exact at every sync, so the loop cannot end with work waiting.
  int64_t fb_pending = 0;
  bool fb_round = true;
  while (upper > 0 || fb_pending > 0)
  {
    int32_t* q = h->d_q_fly[cur];
    unsigned long long* cnt_cur = &P.counters->n_fly[cur];
    unsigned long long* cnt_next = &P.counters->n_fly[cur ^ 1];
    if (prof)
      cudaEventRecord(h->pool_event(3 * rounds + 0), st);
    if (upper + fb_pending < h->pump_threshold && P.lib.n_urr == 0) // the pump's lanes do not carry the URR stream
    {
      // the rest of the generation in one persistent launch: tracking, collision and general-walk warps
      // connected by device-side rings (mcw_pump.cuh)
      MCWK(mcwk_pump, P, h->Q, h->d_pump_ctl, h->pump_ctl_bytes, q, cnt_cur, h->d_q_fallback, &P.counters->n_fallback, h->pump_blocks,
                h->pump_track_warps, h->pump_patience, st);
      R.n_kernel_launches += 2;
      if (prof)
      {
        cudaEventRecord(h->pool_event(3 * rounds + 1), st);
        cudaEventRecord(h->pool_event(3 * rounds + 2), st);
      }
      ++rounds;
      MCWK(mcwk_round_reset, P.counters, cur, true, st);
      R.n_kernel_launches += 1;
      break;
    }
    const bool fused = upper < h->tail_threshold;
    const int drain = (!fused && upper >= h->drain_min_queue) ? h->drain_lanes : 0;
    MCWK(mcwk_track, P, q, cnt_cur, upper, fused ? h->fused_max_blocks : h->track_max_blocks, h->d_q_fly[cur ^ 1], cnt_next, h->d_q_collide, h->d_q_fallback,
               drain, fused, h->collide_min, st);
    if (prof)
      cudaEventRecord(h->pool_event(3 * rounds + 1), st);
    const bool fb_now = fb_round || fused || h->fallback_every_round;
    // collisions of this round: at most every flight plus every general step served now
    const int64_t upper_c = upper + (fb_now ? fb_pending : 0);
    if (fb_now)
    {
      MCWK(mcwk_fallback, P, h->d_q_fallback, &P.counters->n_fallback, h->fallback_blocks, h->d_q_fly[cur ^ 1], cnt_next,
                    h->d_q_collide, st);
      R.n_kernel_launches += 1;
    }
    int32_t* qc = h->d_q_collide;
    if (P.ordered_queues)
    {
      MCWK(mcwk_compact_status, h->d_temp, h->temp_bytes, P.p.status, n, MCW_ST_COLLIDE, qc, &P.counters->n_collide, st);
      R.n_kernel_launches += 2;
    }
    // Particle::event_calculate_xs at the survivors' new energies, all material slots.  With
    // sort_xs_queue the survivors are first ordered by energy (OpenMC's TODO at OMC event.cpp:76-82)
    // and the lookup runs from shared-memory windows (mcw_xs.cuh); small queues are not worth a sort.
    const int64_t sort_min = h->xs_sort_min > 0 ? h->xs_sort_min : std::max<int64_t>(32768, 2400000 / std::max(1, P.n_plan));
    // the tile kernels stage free-atom tables only: libraries with S(a,b) tables refresh through the direct kernel
    const bool staged = h->set.sort_xs_queue != 0 && upper_c >= sort_min && P.lib.n_sab == 0 && P.lib.n_urr == 0;
    MCWK(mcwk_collide, P, qc, &P.counters->n_collide, upper_c, h->d_q_fly[cur ^ 1], cnt_next, staged ? h->d_keys : nullptr,
                 staged ? h->d_q_hist : nullptr, h->d_bins, st);
    R.n_kernel_launches += MCWK(mcwk_xs_refresh, P, qc, &P.counters->n_collide, upper_c, staged, h->d_keys, h->d_q_hist, h->d_bins,
                                           h->d_bin_offsets, h->d_q_sorted, st);
    if (P.ordered_queues)
    {
      MCWK(mcwk_compact_status, h->d_temp, h->temp_bytes, P.p.status, n, MCW_ST_FLY, h->d_q_fly[cur ^ 1], cnt_next, st);
      R.n_kernel_launches += 2;
    }
    if (prof)
      cudaEventRecord(h->pool_event(3 * rounds + 2), st);
    R.n_kernel_launches += 2; // k_track, k_collide
    ++rounds;
    const bool sync_now = (rounds % h->sync_every) == 0 || upper < 4096;
    if (fb_now)
      fb_pending = 0; // k_fallback ran after k_track: nothing waits
    if (sync_now)
    {
      CK(cudaMemcpyAsync(h->h_counters, P.counters, sizeof(McwCounters), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      upper = (int64_t)h->h_counters->n_fly[cur ^ 1];
      if (!fb_now)
        fb_pending = (int64_t)h->h_counters->n_fallback;
    }
    // reset the consumed queues for the next round (the fallback queue only when it was served)
    MCWK(mcwk_round_reset, P.counters, cur, fb_now, st);
    fb_round = sync_now;
    R.n_kernel_launches += 1;
    cur ^= 1;
    if (rounds > 100000000)
      return fail(h, MCW_E_PHYSICS, "event loop did not terminate");
  }
  R.n_event_rounds = rounds;
  if (prof)
    cudaEventRecord(h->ev[5], st);
  // ---- end of transport: counters, k accumulators, sorted bank
  CK(cudaMemcpyAsync(h->h_counters, P.counters, sizeof(McwCounters), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const McwCounters& C = *h->h_counters;
  if (h->pump_threshold > 0)
  {
    long long alive = 0;
    CK(cudaMemcpy(&alive, h->Q.alive, sizeof(alive), cudaMemcpyDeviceToHost));
    if (alive != 0)
      return fail(h, MCW_E_CUDA, "persistent generation kernel stopped with histories unfinished (alive = " +
                                     std::to_string(alive) + ")");
  }
  h->n_bank = std::min<int64_t>((int64_t)C.n_bank, h->bank_capacity);
  double* g = h->d_reduce + h->n_tally_values;
  MCWK(mcwk_reduce_sum, h->d_temp, h->temp_bytes, P.p.k_col, g + 0, n, st);
  MCWK(mcwk_reduce_sum, h->d_temp, h->temp_bytes, P.p.k_abs, g + 1, n, st);
  MCWK(mcwk_reduce_sum, h->d_temp, h->temp_bytes, P.p.k_trk, g + 2, n, st);
  MCWK(mcwk_bank_sort, h->d_temp, h->temp_bytes, P.p, n, P.work_offset, h->d_bank, h->n_bank, h->d_i64_a, h->d_i64_b,
                 h->d_sorted, st);
  R.n_kernel_launches += 4 * 2 + 4;
  CK(cudaEventRecord(h->ev[6], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
  R.ms_total = ms;
  if (prof)
  {
    cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]);
    R.ms_bank = ms;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[4]);
    R.ms_xs = ms; // source initialisation including the first all-material cross-section lookup
    for (int64_t r = 0; r < rounds; ++r)
    {
      cudaEventElapsedTime(&ms, h->pool[3 * r + 0], h->pool[3 * r + 1]);
      R.ms_track += ms;
      cudaEventElapsedTime(&ms, h->pool[3 * r + 1], h->pool[3 * r + 2]);
      R.ms_collide += ms; // general walk + (sort +) collision with the all-material lookup
    }
  }
  R.n_histories = n;
  R.n_segments = (int64_t)C.n_segments;
  R.n_collisions = (int64_t)C.n_collisions;
  R.n_bank_sites = h->n_bank;
  R.n_vertex_edge_hits = (int64_t)C.n_vertex_edge;
  R.n_moved_through_neighbors = (int64_t)C.n_moved;
  R.n_failed_traces = (int64_t)C.n_failed;
  R.n_bank_dropped = (int64_t)C.n_bank_dropped;
  if (C.n_bank_dropped)
    // OMC physics.cpp:203-209: a warning, not an error -- the generation goes on with the sites that fit
    h->err = "The shared fission bank is full. " + std::to_string(C.n_bank_dropped) +
             " fission sites created in this generation were not banked.";
  h->bank_global_start = 0;
  h->bank_global_total = h->n_bank;
  h->records_valid = true;
  h->n_sampled = 0;
  return MCW_OK;
}

int mcw_bank_count(mcw_handle h, int64_t* n_local)
{
  if (!h || !n_local)
    return MCW_E_INVALID_ARGUMENT;
  *n_local = h->n_bank;
  return MCW_OK;
}

int mcw_bank_device_ptr(mcw_handle h, void** sites, int64_t* n)
{
  if (!h || !sites || !n)
    return MCW_E_INVALID_ARGUMENT;
  *sites = h->d_sorted;
  *n = h->n_bank;
  return MCW_OK;
}

int mcw_set_bank_window(mcw_handle h, int64_t global_start, int64_t global_total)
{
  if (!h || global_start < 0 || global_total < 0)
    return MCW_E_INVALID_ARGUMENT;
  h->bank_global_start = global_start;
  h->bank_global_total = global_total;
  return MCW_OK;
}

// OMC eigenvalue.cpp:112-160 on this rank's slice of the global fission bank
int mcw_sample_bank(mcw_handle h, int batch, int64_t* n_sampled)
{
  if (!h || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  if (h->bank_global_total <= 0)
    return fail(h, MCW_E_PHYSICS, "No fission sites banked"); // OMC eigenvalue.cpp:107-109
  CK(cudaSetDevice(h->device));
  const int64_t total = h->bank_global_total, np = h->set.n_particles;
  uint64_t seed0 = MCWK(mcwk_host_init_seed, 0 + batch, 0 /*STREAM_TRACKING*/, h->set.master_seed);
  if ((total < np ? np / total + 1 : 1) * h->n_bank > h->sampled_capacity)
    return fail(h, MCW_E_ALLOCATE, "sampled-site buffer too small for this bank (too few fission sites)");
  MCWK(mcwk_sample_bank, h->d_temp, h->temp_bytes, h->d_sorted, h->n_bank, h->bank_global_start, total, np, seed0, h->d_i64_a,
                   h->d_i64_b, h->d_sampled, h->sampled_capacity, h->stream);
  int64_t last_off = 0, last_cnt = 0;
  if (h->n_bank > 0)
  {
    CK(cudaMemcpyAsync(&last_off, h->d_i64_b + h->n_bank - 1, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&last_cnt, h->d_i64_a + h->n_bank - 1, 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->n_sampled = std::min(last_off + last_cnt, h->sampled_capacity);
  h->last.n_kernel_launches += 3;
  if (n_sampled)
    *n_sampled = h->n_sampled;
  return MCW_OK;
}

// OMC eigenvalue.cpp:184-206: only the last rank trims / pads to exactly n_particles
int mcw_finish_sampling(mcw_handle h, int64_t global_start_of_sampled, int is_last, int64_t* n_sampled)
{
  if (!h || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(h->device));
  const int64_t np = h->set.n_particles;
  int64_t finish = global_start_of_sampled + h->n_sampled;
  if (is_last)
  {
    if (finish > np)
      h->n_sampled = std::max<int64_t>(0, np - global_start_of_sampled);
    else if (finish < np)
    {
      int64_t needed = np - finish;
      if (h->n_sampled + needed > h->sampled_capacity)
        return fail(h, MCW_E_ALLOCATE, "sampled-site buffer too small for padding");
      MCWK(mcwk_pad_from_bank_end, h->d_sorted, h->n_bank, h->d_sampled, h->n_sampled, needed, h->stream);
      CK(cudaStreamSynchronize(h->stream));
      h->n_sampled += needed;
      h->last.n_kernel_launches += 1;
    }
  }
  if (n_sampled)
    *n_sampled = h->n_sampled;
  return MCW_OK;
}

int mcw_sampled_device_ptr(mcw_handle h, void** sites, int64_t* n)
{
  if (!h || !sites || !n)
    return MCW_E_INVALID_ARGUMENT;
  *sites = h->d_sampled;
  *n = h->n_sampled;
  return MCW_OK;
}

int mcw_source_device_ptr(mcw_handle h, void** sites, int64_t* capacity)
{
  if (!h || !sites || !capacity)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  *sites = h->d_source;
  *capacity = h->n_local;
  return MCW_OK;
}

int mcw_set_source_count(mcw_handle h, int64_t n)
{
  if (!h || n < 0 || n > h->n_local)
    return MCW_E_INVALID_ARGUMENT;
  h->n_source = n;
  return MCW_OK;
}

int mcw_reduction_device_ptr(mcw_handle h, void** doubles, int64_t* n)
{
  if (!h || !doubles || !n)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  *doubles = h->d_reduce;
  *n = h->n_reduce;
  return MCW_OK;
}

// OMC eigenvalue.cpp:49-69,313-351; tallies/tally.cpp:639-670,805-846; simulation.cpp:334-347
int mcw_finalize_generation(mcw_handle h, int batch, mcw_generation_result* out)
{
  if (!h || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(h->device));
  double g[5];
  CK(cudaMemcpyAsync(g, h->d_reduce + h->n_tally_values, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const mcw_settings& S = h->set;
  const double w = g[4];
  double k_gen = g[2] / (double)S.n_particles;
  h->k_generation.push_back(k_gen);
  int n = batch > S.n_inactive ? h->n_realizations + 1 : 0;
  if (n <= 0)
    h->keff = k_gen;
  else
  {
    h->k_sum[0] += k_gen;
    h->k_sum[1] += k_gen * k_gen;
    h->keff = h->k_sum[0] / n;
    if (n > 1)
      h->keff_std = std::sqrt((h->k_sum[1] / n - h->keff * h->keff) / (n - 1));
  }
  h->n_realizations += 1;
  for (int i = 0; i < 4; ++i)
  {
    double val = g[i] / w;
    h->gt_sum[i] += val;
    h->gt_sum_sq[i] += val * val;
  }
  mcw_generation_result& R = h->last;
  R.k_generation = k_gen;
  R.keff = h->keff;
  R.keff_std = h->keff_std;
  R.k_collision = g[0] / w;
  R.k_absorption = g[1] / w;
  R.k_tracklength = g[2] / w;
  R.leakage = g[3] / w;
  R.total_weight = w;
  if (batch > S.n_inactive)
  {
    const double norm = 1.0 / ((double)S.n_particles * 1.0);
    for (auto& t : h->tallies)
    {
      t.n_realizations += 1;
      MCWK(mcwk_tally_accumulate, h->d_reduce + t.offset, t.d_sum, t.d_sum_sq, t.n_values, norm, h->stream);
      R.n_kernel_launches += 1;
    }
  }
  else
  {
    for (int i = 0; i < 4; ++i)
      h->gt_sum[i] = h->gt_sum_sq[i] = 0.0;
    h->n_realizations = 0;
    if (h->n_tally_values > 0)
      CK(cudaMemsetAsync(h->d_reduce, 0, h->n_tally_values * sizeof(double), h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (out)
    *out = R;
  return MCW_OK;
}

int mcw_run_generation(mcw_handle h, int batch, mcw_generation_result* out)
{
  if (!h)
    return MCW_E_INVALID_ARGUMENT;
  if (h->nranks != 1)
    return fail(h, MCW_E_UNSUPPORTED, "mcw_run_generation is single-rank; use the split calls with an exchange");
  int rc = mcw_transport_generation(h, batch);
  if (rc)
    return rc;
  rc = mcw_sample_bank(h, batch, nullptr);
  if (rc)
    return rc;
  rc = mcw_finish_sampling(h, 0, 1, nullptr);
  if (rc)
    return rc;
  std::swap(h->d_source, h->d_sampled);
  h->n_source = std::min(h->n_sampled, h->n_local);
  h->n_sampled = 0;
  return mcw_finalize_generation(h, batch, out);
}

// ------------------------------------------------------------------ read-back
int mcw_tally_shape(mcw_handle h, int id, int64_t* n_spatial, int32_t* n_energy, int32_t* n_scores)
{
  if (!h)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  for (auto& t : h->tallies)
    if (t.id == id)
    {
      if (n_spatial)
        *n_spatial = t.n_spatial;
      if (n_energy)
        *n_energy = t.n_energy;
      if (n_scores)
        *n_scores = t.n_scores;
      return MCW_OK;
    }
  return fail(h, MCW_E_INVALID_ARGUMENT, "unknown tally id");
}

int mcw_tally_nuclide_bins(mcw_handle h, int id, int32_t* n_nuclide_bins)
{
  if (!h || !n_nuclide_bins)
    return MCW_E_INVALID_ARGUMENT;
  for (auto& t : h->tallies)
    if (t.id == id)
    {
      *n_nuclide_bins = t.n_nuc_bins();
      return MCW_OK;
    }
  return fail(h, MCW_E_INVALID_ARGUMENT, "unknown tally id");
}

int mcw_get_tally(mcw_handle h, int id, double* sum, double* sum_sq, int64_t n, int32_t* n_realizations)
{
  if (!h || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(h->device));
  for (auto& t : h->tallies)
    if (t.id == id)
    {
      int64_t m = std::min(n, t.n_values);
      if (sum)
        CK(cudaMemcpy(sum, t.d_sum, m * sizeof(double), cudaMemcpyDeviceToHost));
      if (sum_sq)
        CK(cudaMemcpy(sum_sq, t.d_sum_sq, m * sizeof(double), cudaMemcpyDeviceToHost));
      if (n_realizations)
        *n_realizations = t.n_realizations;
      return MCW_OK;
    }
  return fail(h, MCW_E_INVALID_ARGUMENT, "unknown tally id");
}

int mcw_get_history_records(mcw_handle h, mcw_history_record* rec, int64_t capacity, int64_t* n)
{
  if (!h || !n || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  *n = h->records_valid ? h->last.n_histories : 0;
  if (rec && capacity > 0 && *n > 0)
  {
    CK(cudaSetDevice(h->device));
    MCWK(mcwk_fill_records, h->P.p, *n, h->d_records, h->stream);
    CK(cudaMemcpyAsync(rec, h->d_records, std::min(capacity, *n) * sizeof(mcw_history_record), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return MCW_OK;
}

int mcw_get_traces(mcw_handle h, mcw_trace_point* out, int64_t capacity, int64_t* n)
{
  if (!h || !n || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  if (!h->P.trace)
    return fail(h, MCW_E_NOT_READY, "trace points are cached only with settings.record_traces = 2");
  CK(cudaSetDevice(h->device));
  *n = (int64_t)h->h_counters->n_trace; // counters of the last generation (copied at the end of transport)
  if (*n > h->P.trace_capacity)
    return fail(h, MCW_E_ALLOCATE, "the generation cached " + std::to_string(*n) + " trace points, the buffer holds " +
                                       std::to_string(h->P.trace_capacity) + " (2048 per history, at most 2^23): trace fewer histories");
  const int64_t have = *n;
  if (out && capacity > 0 && have > 0)
  {
    std::vector<mcw_trace_point> all((size_t)have);
    CK(cudaMemcpy(all.data(), h->P.trace, have * sizeof(mcw_trace_point), cudaMemcpyDeviceToHost));
    std::sort(all.begin(), all.end(), [](const mcw_trace_point& a, const mcw_trace_point& b) {
      return a.ray_id != b.ray_id ? a.ray_id < b.ray_id : (a.seq != b.seq ? a.seq < b.seq : a.kind < b.kind);
    });
    memcpy(out, all.data(), std::min(capacity, have) * sizeof(mcw_trace_point));
  }
  return MCW_OK;
}

int mcw_get_fission_bank(mcw_handle h, mcw_source_site* sites, int64_t capacity, int64_t* n)
{
  if (!h || !n || !h->prepared)
    return MCW_E_INVALID_ARGUMENT;
  *n = h->n_bank;
  if (sites && capacity > 0 && h->n_bank > 0)
  {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(sites, h->d_sorted, std::min(capacity, h->n_bank) * sizeof(mcw_source_site),
                  cudaMemcpyDeviceToHost));
  }
  return MCW_OK;
}

int mcw_get_grid_index(mcw_handle h, int i_nuc, int32_t* out, int n)
{
  if (!h || !out)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  if (i_nuc < 0 || i_nuc >= h->P.lib.n_nuc || n != h->P.lib.M + 1)
    return MCW_E_INVALID_ARGUMENT;
  CK(cudaMemcpy(out, h->P.lib.grid_index + (int64_t)i_nuc * (h->P.lib.M + 1), n * sizeof(int32_t),
                cudaMemcpyDeviceToHost));
  return MCW_OK;
}

// ------------------------------------------------------------------ stand-alone kernels
int mcw_xs_lookup(mcw_handle h, int64_t n, const double* E, const int32_t* material, double* out)
{
  if (!h || !E || !material || !out || n < 0)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  CK(cudaSetDevice(h->device));
  for (int64_t i = 0; i < n; ++i)
    if (material[i] < 0 || material[i] >= h->P.lib.n_mat)
      return fail(h, MCW_E_INVALID_ARGUMENT, "material index out of range");
  double *dE = nullptr, *dO = nullptr;
  int32_t* dM = nullptr;
  CK(dev_alloc(h, &dE, n));
  CK(dev_alloc(h, &dM, n));
  CK(dev_alloc(h, &dO, 4 * n));
  CK(cudaMemcpyAsync(dE, E, n * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dM, material, n * 4, cudaMemcpyHostToDevice, h->stream));
  MCWK(mcwk_xs_lookup_array, h->P.lib, n, dE, dM, dO, h->stream);
  CK(cudaMemcpyAsync(out, dO, 4 * n * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  dev_free(h, dE);
  dev_free(h, dM);
  dev_free(h, dO);
  return MCW_OK;
}

int mcw_xs_lookup_device(mcw_handle h, int64_t n, const void* d_E, const void* d_material, void* d_out, int sorted,
                         int repeats, float* ms_per_launch)
{
  if (!h || !d_E || !d_material || !d_out || n < 0 || repeats < 1 || sorted < 0 || sorted > 2)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  CK(cudaSetDevice(h->device));
  const double* E = (const double*)d_E;
  const int32_t* mat = (const int32_t*)d_material;
  uint32_t *keys = nullptr, *keys_out = nullptr;
  int32_t *idx = nullptr, *idx_sorted = nullptr;
  unsigned char* temp = nullptr;
  size_t temp_bytes = 0;
  if (sorted == 2 && n > 0)
  {
    if (n > INT32_MAX)
      return fail(h, MCW_E_INVALID_ARGUMENT, "mcw_xs_lookup_device: at most 2^31-1 points per call");
    temp_bytes = MCWK(mcwk_sort_temp_bytes, n);
    CK(dev_alloc(h, &keys, n));
    CK(dev_alloc(h, &keys_out, n));
    CK(dev_alloc(h, &idx, n));
    CK(dev_alloc(h, &idx_sorted, n));
    CK(dev_alloc(h, &temp, temp_bytes));
  }
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (h->P.lib.n_sab)
    sorted = 0; // the staged kernels hold free-atom tables only (no history here: probability tables never apply)
  for (int r = 0; r < repeats; ++r)
  {
    if (sorted == 0)
      MCWK(mcwk_xs_lookup_array, h->P.lib, n, E, mat, (double*)d_out, h->stream);
    else if (sorted == 1)
      MCWK(mcwk_xs_lookup_staged, h->P.lib, n, E, mat, nullptr, (double*)d_out, h->xs_plan != 0, h->stream);
    else
    {
      MCWK(mcwk_xs_sort, n, E, mat, temp, temp_bytes, keys, keys_out, idx, idx_sorted, h->stream);
      MCWK(mcwk_xs_lookup_staged, h->P.lib, n, E, mat, idx_sorted, (double*)d_out, h->xs_plan != 0, h->stream);
    }
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  if (ms_per_launch)
    *ms_per_launch = ms / repeats;
  if (keys)
  {
    dev_free(h, keys);
    dev_free(h, keys_out);
    dev_free(h, idx);
    dev_free(h, idx_sorted);
    dev_free(h, temp);
  }
  return MCW_OK;
}

int mcw_trace_rays(mcw_handle h, int64_t n, const double* start, const double* dir, const int32_t* start_elem,
                   const int32_t* start_side, const double* max_distance, double* out_distance, int32_t* out_segments,
                   uint64_t* out_hash)
{
  if (!h || !start || !dir || !start_elem || !out_distance || !out_segments || n < 0)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  CK(cudaSetDevice(h->device));
  double *dS, *dD, *dMx = nullptr, *dDist;
  int32_t *dE, *dSd = nullptr, *dSeg, *dStat;
  uint64_t* dH;
  CK(dev_alloc(h, &dS, 3 * n));
  CK(dev_alloc(h, &dD, 3 * n));
  CK(dev_alloc(h, &dDist, n));
  CK(dev_alloc(h, &dE, n));
  CK(dev_alloc(h, &dSeg, n));
  CK(dev_alloc(h, &dStat, n));
  CK(dev_alloc(h, &dH, n));
  CK(cudaMemcpy(dS, start, 3 * n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, dir, 3 * n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dE, start_elem, n * 4, cudaMemcpyHostToDevice));
  if (start_side)
  {
    CK(dev_alloc(h, &dSd, n));
    CK(cudaMemcpy(dSd, start_side, n * 4, cudaMemcpyHostToDevice));
  }
  if (max_distance)
  {
    CK(dev_alloc(h, &dMx, n));
    CK(cudaMemcpy(dMx, max_distance, n * 8, cudaMemcpyHostToDevice));
  }
  MCWK(mcwk_trace_rays, h->P, n, dS, dD, dE, dSd, dMx, dDist, dSeg, dH, dStat, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpy(out_distance, dDist, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_segments, dSeg, n * 4, cudaMemcpyDeviceToHost));
  if (out_hash)
    CK(cudaMemcpy(out_hash, dH, n * 8, cudaMemcpyDeviceToHost));
  for (void* p : {(void*)dS, (void*)dD, (void*)dMx, (void*)dDist, (void*)dE, (void*)dSd, (void*)dSeg, (void*)dStat,
                  (void*)dH})
    dev_free(h, p);
  return MCW_OK;
}

int mcw_locate_points(mcw_handle h, int64_t n, const double* xyz, const int64_t* ray_id, int32_t* elem)
{
  if (!h || !xyz || !elem || n < 0)
    return MCW_E_INVALID_ARGUMENT;
  int rc = prepare(h);
  if (rc)
    return rc;
  CK(cudaSetDevice(h->device));
  double* dX;
  int64_t* dR = nullptr;
  int32_t* dE;
  CK(dev_alloc(h, &dX, 3 * n));
  CK(dev_alloc(h, &dE, n));
  CK(cudaMemcpy(dX, xyz, 3 * n * 8, cudaMemcpyHostToDevice));
  if (ray_id)
  {
    CK(dev_alloc(h, &dR, n));
    CK(cudaMemcpy(dR, ray_id, n * 8, cudaMemcpyHostToDevice));
  }
  MCWK(mcwk_locate_points, h->P, n, dX, dR, dE, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpy(elem, dE, n * 4, cudaMemcpyDeviceToHost));
  dev_free(h, dX);
  dev_free(h, dR);
  dev_free(h, dE);
  return MCW_OK;
}

int mcw_test_division(mcw_handle h, int64_t n, const double* a, const double* b, double* q_fast, double* q_plain, int32_t* ok)
{
  if (!h || !a || !b || !q_fast || !q_plain || !ok || n < 0)
    return MCW_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(h->device));
  double *da, *db, *df, *dp;
  int32_t* dk;
  CK(dev_alloc(h, &da, n));
  CK(dev_alloc(h, &db, n));
  CK(dev_alloc(h, &df, n));
  CK(dev_alloc(h, &dp, n));
  CK(dev_alloc(h, &dk, n));
  CK(cudaMemcpy(da, a, n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice));
  MCWK(mcwk_test_division, n, da, db, df, dp, dk, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpy(q_fast, df, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q_plain, dp, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ok, dk, n * 4, cudaMemcpyDeviceToHost));
  for (void* p : {(void*)da, (void*)db, (void*)df, (void*)dp, (void*)dk})
    dev_free(h, p);
  return MCW_OK;
}

int mcw_struct_sizes(int32_t* out, int n)
{
  const int32_t v[] = {(int32_t)sizeof(mcw_source_site), (int32_t)sizeof(mcw_nuclide_desc),
                       (int32_t)sizeof(mcw_material_desc), (int32_t)sizeof(mcw_settings),
                       (int32_t)sizeof(mcw_generation_result), (int32_t)sizeof(mcw_history_record)};
  for (int i = 0; i < n && i < 6; ++i)
    out[i] = v[i];
  return 6;
}

} // extern "C"
