// problem_io.hpp -- reader of the problem files written by macaw_b200/problem_io.py
// (magic "MCWPROB1", uint32 record count, records {char name[32]; u32 dtype; u32 reserved; u64 count; data}).
// Stands in for what the MOOSE adapter gets from libMesh + the input file; set-up only, not hot path.
#pragma once
#include <cstring>
#include <fstream>
#include <map>

#include "macaw_host.hpp"

namespace macaw
{

struct Problem
{
  Mesh mesh;
  Library library;
  std::vector<unsigned> blocks, materials;
  std::vector<double> temperature;
  mcw_settings settings{};
  struct TallyDef
  {
    int id, estimator, granularity;
    std::vector<int32_t> scores;
    std::vector<double> energy_bins;
    std::vector<int32_t> nuclides; // nuclide bins already as library indices (-1 total, -2 never scored); empty = total
  };
  std::vector<TallyDef> tallies;
};

class ProblemFile
{
public:
  explicit ProblemFile(const std::string& path)
  {
    std::ifstream f(path, std::ios::binary);
    if (!f)
      mooseError("cannot open problem file " + path);
    char magic[8];
    f.read(magic, 8);
    if (std::memcmp(magic, "MCWPROB1", 8) != 0)
      mooseError(path + " is not a MCWPROB1 problem file");
    uint32_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 4);
    for (uint32_t k = 0; k < n; ++k)
    {
      char name[33] = {0};
      uint32_t dtype = 0, reserved = 0;
      uint64_t count = 0;
      f.read(name, 32);
      f.read(reinterpret_cast<char*>(&dtype), 4);
      f.read(reinterpret_cast<char*>(&reserved), 4);
      f.read(reinterpret_cast<char*>(&count), 8);
      const size_t item = dtype == 1 ? 4 : 8;
      Rec r;
      r.dtype = (int)dtype;
      r.bytes.resize(count * item);
      f.read(r.bytes.data(), (std::streamsize)r.bytes.size());
      f.ignore((std::streamsize)((8 - r.bytes.size() % 8) % 8));
      if (!f)
        mooseError("truncated problem file " + path);
      _recs[name] = std::move(r);
    }
  }
  bool has(const std::string& n) const { return _recs.count(n) != 0; }
  template <typename T>
  std::vector<T> get(const std::string& n, int dtype) const
  {
    auto it = _recs.find(n);
    if (it == _recs.end())
      mooseError("problem file lacks record " + n);
    if (it->second.dtype != dtype)
      mooseError("record " + n + " has the wrong type");
    std::vector<T> out(it->second.bytes.size() / sizeof(T));
    std::memcpy(out.data(), it->second.bytes.data(), it->second.bytes.size());
    return out;
  }
  std::vector<double> f64(const std::string& n) const { return get<double>(n, 0); }
  std::vector<int32_t> i32(const std::string& n) const { return get<int32_t>(n, 1); }
  std::vector<int64_t> i64(const std::string& n) const { return get<int64_t>(n, 2); }

  Problem problem() const
  {
    Problem P;
    P.mesh.elem_type = i32("mesh.elem_type")[0];
    P.mesh.xyz = f64("mesh.xyz");
    P.mesh.conn = i64("mesh.conn");
    P.mesh.block_id = i32("mesh.block_id");
    if (has("mesh.neighbors"))
      P.mesh.neighbors = i64("mesh.neighbors");
    P.mesh.side_bnd_id = i32("mesh.side_bnd_id");
    P.mesh.bnd_kind = i32("mesh.bnd_kind");
    const auto counts = i32("lib.counts");
    // further temperature sets of the nuclides follow them as nuc<n>, nuc<n + 1>, ...; lib.set_of names each one's nuclide
    const std::vector<int32_t> set_of = has("lib.set_of") ? i32("lib.set_of") : std::vector<int32_t>();
    for (int k = 0; k < counts[0] + (int)set_of.size(); ++k)
    {
      const std::string p = "nuc" + std::to_string(k) + ".";
      Nuclide n;
      const auto I = i32(p + "ints");
      const auto S = f64(p + "scalars");
      n.energy = f64(p + "energy");
      n.xs = f64(p + "xs");
      n.elastic = f64(p + "elastic");
      n.inelastic = f64(p + "inelastic");
      n.d.fissionable = I[1];
      n.d.inel_threshold = I[3];
      n.d.n_precursor = I[4];
      n.d.awr = S[0], n.d.kT = S[1], n.d.inel_threshold_energy = S[2], n.d.inel_mass_ratio = S[3];
      n.d.nu_poly[0] = S[4], n.d.nu_poly[1] = S[5], n.d.watt_a = S[6], n.d.watt_b = S[7], n.d.q_fission = S[8];
      const auto dy = f64(p + "delayed_yield"), dt = f64(p + "delayed_theta");
      for (int g = 0; g < MCW_MAX_DELAYED_GROUPS; ++g)
        n.d.delayed_yield[g] = dy[g], n.d.delayed_theta[g] = dt[g];
      auto table = [&](const std::string& key, Tabular2D& t) {
        if (!has(p + key + ".e"))
          return;
        t.energy = f64(p + key + ".e");
        const size_t ne = t.energy.size();
        const auto ti = i32(p + key + ".i"); // histogram flag, in_cm flag, offsets[ne + 1], interpolation[ne], n_discrete[ne]
        t.histogram_in_energy = ti[0];
        if (key == "inel_energy")
          n.inelastic_in_cm = ti[1];
        t.offset.assign(ti.begin() + 2, ti.begin() + 2 + ne + 1);
        t.interpolation.assign(ti.begin() + 3 + ne, ti.begin() + 3 + 2 * ne);
        t.n_discrete.assign(ti.begin() + 3 + 2 * ne, ti.begin() + 3 + 3 * ne);
        const auto xpc = f64(p + key + ".xpc");
        const size_t np = xpc.size() / 3;
        t.x.assign(xpc.begin(), xpc.begin() + np);
        t.pdf.assign(xpc.begin() + np, xpc.begin() + 2 * np);
        t.cdf.assign(xpc.begin() + 2 * np, xpc.end());
      };
      table("el_angle", n.elastic_angle);
      table("inel_angle", n.inelastic_angle);
      table("inel_energy", n.inelastic_energy);
      // a function record: [kind, interpolation] + x + y (tabulated) or + coefficients (polynomial)
      auto fn = [&](const std::string& key, Function1D& f) {
          if (!has(p + key))
            return;
          const auto v = f64(p + key);
          f.kind = (int)v[0], f.interpolation = (int)v[1];
          const size_t m = v.size() - 2;
          if (f.kind == MCW_FN_TABULATED)
          {
            f.x.assign(v.begin() + 2, v.begin() + 2 + m / 2);
            f.y.assign(v.begin() + 2 + m / 2, v.end());
          }
          else
            f.y.assign(v.begin() + 2, v.end());
      };
      if (has(p + "fis.head"))
      {
        // fission neutrons as a nuclide file has them: fis.head = count, laws; fis.u; functions fis.nu (total nu),
        // fis<j>.y (yield), .a, .b; continuous tabular spectra fis<j>.en like the nuclide's tables
        const auto head = i32(p + "fis.head");
        const auto us = f64(p + "fis.u");
        fn("fis.nu", n.total_nu);
        const int keep_cm = n.inelastic_in_cm;
        for (int j = 0; j < head[0]; ++j)
        {
          const std::string q = "fis" + std::to_string(j) + ".";
          FissionProduct F;
          F.law = head[1 + j];
          F.u = us[j];
          fn(q + "y", F.yield);
          fn(q + "a", F.a);
          fn(q + "b", F.b);
          table(q + "en", F.energy);
          n.fission_products.push_back(std::move(F));
        }
        n.inelastic_in_cm = keep_cm;
      }
      if (has(p + "rx.head"))
      {
        // the scattering reactions after the first: rx<j>.i = mt, threshold, law, in_cm, n_particles; rx<j>.f =
        // threshold_energy, mass_ratio, yield, total_mass, q_value; rx<j>.xs; tables rx<j>.ang / rx<j>.en like the
        // nuclide's; rx<j>.corr.i / .f, rx<j>.km
        const auto head = f64(p + "rx.head");
        n.inelastic_yield = head[1];
        n.inelastic_u = head.size() > 2 ? head[2] : 0.0;
        fn("rx.y", n.inelastic_yield_fn);
        fn("rx.a", n.inelastic_a);
        fn("rx.b", n.inelastic_b);
        const int keep_cm = n.inelastic_in_cm;
        for (int j = 0; j < (int)head[0]; ++j)
        {
          const std::string q = "rx" + std::to_string(j) + ".";
          Reaction R;
          const auto ri = i32(p + q + "i");
          const auto rf = f64(p + q + "f");
          R.mt = ri[0], R.threshold = ri[1], R.law = ri[2], R.in_cm = ri[3], R.nbody_n = ri[4];
          R.threshold_energy = rf[0], R.mass_ratio = rf[1], R.yield = rf[2], R.nbody_mass_ratio = rf[3], R.nbody_q = rf[4];
          R.xs = f64(p + q + "xs");
          R.u = has(p + q + "u") ? f64(p + q + "u")[0] : 0.0;
          fn(q + "y", R.yield_fn);
          fn(q + "a", R.a);
          fn(q + "b", R.b);
          table(q + "ang", R.angle);
          table(q + "en", R.energy);
          const size_t np = R.energy.offset.empty() ? 0 : (size_t)R.energy.offset.back();
          if (R.law == MCW_LAW_CORRELATED)
          {
            const auto ci = i32(p + q + "corr.i");
            const auto cf = f64(p + q + "corr.f");
            R.corr_offset.assign(ci.begin(), ci.begin() + np + 1);
            R.corr_interp.assign(ci.begin() + np + 1, ci.begin() + 2 * np + 1);
            const size_t nm = (size_t)R.corr_offset.back();
            R.corr_mu.assign(cf.begin(), cf.begin() + nm);
            R.corr_pdf.assign(cf.begin() + nm, cf.begin() + 2 * nm);
            R.corr_cdf.assign(cf.begin() + 2 * nm, cf.begin() + 3 * nm);
          }
          else if (R.law == MCW_LAW_KALBACH_MANN)
          {
            const auto km = f64(p + q + "km");
            R.km_r.assign(km.begin(), km.begin() + np);
            R.km_a.assign(km.begin() + np, km.begin() + 2 * np);
          }
          n.more_reactions.push_back(std::move(R));
        }
        n.inelastic_in_cm = keep_cm;
      }
      if (has(p + "law.i"))
      {
        // law, in_cm, n_particles [, offsets[np + 1], interpolation[np]]; total_mass, q_value [, mu / pdf / cdf | r / a]
        const auto li = i32(p + "law.i");
        const auto lf = f64(p + "law.f");
        n.inelastic_law = li[0];
        n.inelastic_in_cm = li[1];
        n.d.nbody_n = li[2];
        n.d.nbody_mass_ratio = lf[0], n.d.nbody_q = lf[1];
        const size_t np = n.inelastic_energy.offset.empty() ? 0 : (size_t)n.inelastic_energy.offset.back();
        if (n.inelastic_law == MCW_LAW_CORRELATED)
        {
          n.corr_offset.assign(li.begin() + 3, li.begin() + 3 + np + 1);
          n.corr_interp.assign(li.begin() + 4 + np, li.begin() + 4 + 2 * np);
          const size_t nm = (size_t)n.corr_offset.back();
          n.corr_mu.assign(lf.begin() + 2, lf.begin() + 2 + nm);
          n.corr_pdf.assign(lf.begin() + 2 + nm, lf.begin() + 2 + 2 * nm);
          n.corr_cdf.assign(lf.begin() + 2 + 2 * nm, lf.begin() + 2 + 3 * nm);
        }
        else if (n.inelastic_law == MCW_LAW_KALBACH_MANN)
        {
          n.km_r.assign(lf.begin() + 2, lf.begin() + 2 + np);
          n.km_a.assign(lf.begin() + 2 + np, lf.begin() + 2 + 2 * np);
        }
      }
      if (k < counts[0])
        P.library.nuclides.push_back(std::move(n));
      else
        P.library.temperature_sets.push_back({set_of[k - counts[0]], std::move(n)});
    }
    for (int k = 0; k < counts[1]; ++k)
    {
      const std::string p = "mat" + std::to_string(k) + ".";
      Material M{i32(p + "nuclide"), f64(p + "density"), has(p + "isotropic") ? i32(p + "isotropic") : std::vector<int32_t>(), {}, {}, {}};
      if (has(p + "sab_i"))
      {
        const auto si = i32(p + "sab_i");
        for (size_t a = 0; a + 1 < si.size(); a += 2)
          M.sab_position.push_back(si[a]), M.sab_table.push_back(si[a + 1]);
        M.sab_fraction = f64(p + "sab_f");
      }
      P.library.materials.push_back(std::move(M));
    }
    const int n_thermal = has("lib.n_thermal") ? i32("lib.n_thermal")[0] : 0;
    for (int k = 0; k < n_thermal; ++k)
    {
      const std::string p = "sab" + std::to_string(k) + ".";
      ThermalTable t;
      const auto I = i32(p + "ints");
      t.inelastic_type = I[0], t.elastic_type = I[1], t.n_mu = I[2], t.n_e_out = I[3], t.skewed = I[4], t.n_elastic_mu = I[5];
      t.energy_max = f64(p + "emax")[0];
      t.i_xs_energy = f64(p + "i_xs_energy"), t.i_xs = f64(p + "i_xs"), t.i_energy = f64(p + "i_energy");
      t.i_e_out = f64(p + "i_e_out"), t.i_pdf = f64(p + "i_pdf"), t.i_cdf = f64(p + "i_cdf"), t.i_mu = f64(p + "i_mu");
      t.e_x = f64(p + "e_x"), t.e_y = f64(p + "e_y"), t.e_mu = f64(p + "e_mu");
      t.i_offset = i32(p + "i_offset");
      P.library.thermal.push_back(std::move(t));
    }
    if (has("lib.sab_same"))  // several temperatures per table: the first entry of each entry's table, and the kTs
      P.library.thermal_same = i32("lib.sab_same"), P.library.thermal_kT = f64("lib.sab_kT");
    const int n_urr = has("lib.n_urr") ? i32("lib.n_urr")[0] : 0;
    for (int k = 0; k < n_urr; ++k)
    {
      const std::string p = "urr" + std::to_string(k) + ".";
      const auto I = i32(p + "ints");
      UrrTable u;
      u.nuclide = I[0], u.interpolation = I[1], u.inelastic_flag = I[2], u.absorption_flag = I[3], u.multiply_smooth = I[4];
      u.n_band = I[6];
      u.energy = f64(p + "energy");
      u.table = f64(p + "table");
      P.library.urr.push_back(std::move(u));
    }
    for (int32_t b : i32("kernel.blocks"))
      P.blocks.push_back((unsigned)b);
    for (int32_t m : i32("kernel.materials"))
      P.materials.push_back((unsigned)m);
    P.temperature = f64("kernel.temperature");
    const auto si = i64("settings.ints");
    const auto sr = f64("settings.reals");
    mcw_settings& s = P.settings;
    s.n_particles = si[0], s.master_seed = si[1], s.n_inactive = (int32_t)si[2], s.n_log_bins = (int32_t)si[3];
    s.is_2d = (int32_t)si[4], s.seed_through_double = (int32_t)si[5], s.source_only_fissionable = (int32_t)si[6];
    s.max_events = (int32_t)si[7], s.sort_xs_queue = (int32_t)si[8], s.record_traces = (int32_t)si[9];
    s.z_coord = sr[0];
    for (int k = 0; k < 6; ++k)
      s.source_box[k] = sr[1 + k];
    const int nt = i32("tallies.count")[0];
    for (int k = 0; k < nt; ++k)
    {
      const std::string p = "tally" + std::to_string(k) + ".";
      const auto I = i32(p + "ints");
      P.tallies.push_back({I[0], I[1], I[2], i32(p + "scores"), f64(p + "energy_bins"),
                           has(p + "nuclides") ? i32(p + "nuclides") : std::vector<int32_t>()});
    }
    return P;
  }

private:
  struct Rec
  {
    int dtype;
    std::vector<char> bytes;
  };
  std::map<std::string, Rec> _recs;
};

} // namespace macaw
