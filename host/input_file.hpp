// input_file.hpp -- the reference's input-file surface for the hot path, without MOOSE:
//
//   * a MOOSE input file (hit syntax: [Block] / [./sub] [../] / [A/B] paths, key = value, quoted lists,
//     # comments) with the blocks MaCaw's examples use for the path:
//       [Mesh]        GeneratedMeshGenerator (+ SubdomainBoundingBoxGenerator)   examples/infinite_medium/infinite_medium.i:1-22
//                     FileMeshGenerator (Exodus-II, host/exodus_reader.hpp) + TransformGenerator TRANSLATE
//                                                                               examples/beavrs_assembly/run_assembly.i:5-16
//                     FancyExtruderGenerator (QUAD4 -> HEX8)                     examples/beavrs_assembly/run_assembly_3d.i:16-31
//       [Variables/temperature] initial_condition                               :30-34
//       [RayKernels]  CollisionKernel: temperature, blocks, materials, z_coord  MC raykernels/CollisionKernel.C:24-42
//       [RayBCs]      ReflectRayBC / KillRayBC: boundary                         RT raybcs/ReflectRayBC.C, KillRayBC.C
//       [UserObjects] OpenMCStudy; OpenMCTally: id, scores, filters, estimator, energy_bins  MC userobjects/*.C
//       [AuxKernels]  OpenMCTallyAux: tally_id, granularity, score, energy_bin  MC auxkernels/OpenMCTallyAux.C:26-50
//       [Executioner] num_steps
//   * OpenMC's settings.xml (particles, batches, inactive, seed, box source) and materials.xml
//     (material id, density, nuclide name + ao / wo, <isotropic>)               OMC settings.cpp:230-420, material.cpp:50-412
//   * a nuclide table file in place of cross_sections.xml + HDF5 (written by
//     `python -m macaw_b200.problem_io --library`): the synthetic continuous-energy tables by name.
//
// Blocks and parameters outside the list above are parsed and ignored, like MOOSE objects that are
// not on the path (Postprocessors, Outputs, ...).  Set-up code only; nothing here is on the hot path.
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <map>
#include <sstream>

#include "exodus_reader.hpp"
#include "problem_io.hpp"

namespace macaw
{

// ------------------------------------------------------------------ hit (MOOSE input) subset
struct HitNode
{
  std::string name;
  std::vector<std::pair<std::string, std::string>> params;
  std::vector<HitNode> children;
  const std::string* find(const std::string& key) const
  {
    for (const auto& kv : params)
      if (kv.first == key)
        return &kv.second;
    return nullptr;
  }
  std::string get(const std::string& key, const std::string& dflt = "") const
  {
    const std::string* v = find(key);
    return v ? *v : dflt;
  }
  const HitNode* child(const std::string& n) const
  {
    for (const auto& c : children)
      if (c.name == n)
        return &c;
    return nullptr;
  }
  HitNode& childOrNew(const std::string& n)
  {
    for (auto& c : children)
      if (c.name == n)
        return c;
    children.push_back(HitNode{n, {}, {}});
    return children.back();
  }
};

// MOOSE's `active = '...'` / `inactive = '...'` lists on a block select which sub-blocks exist
inline bool blockIsActive(const HitNode& parent, const std::string& child);

inline std::string trim(const std::string& s)
{
  const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

inline std::vector<std::string> splitWords(const std::string& s)
{
  std::vector<std::string> out;
  std::istringstream is(s);
  for (std::string w; is >> w;)
    out.push_back(w);
  return out;
}

inline bool blockIsActive(const HitNode& parent, const std::string& child)
{
  if (const std::string* a = parent.find("active"))
  {
    const auto l = splitWords(*a);
    if (std::find(l.begin(), l.end(), child) == l.end())
      return false;
  }
  if (const std::string* a = parent.find("inactive"))
  {
    const auto l = splitWords(*a);
    if (std::find(l.begin(), l.end(), child) != l.end())
      return false;
  }
  return true;
}

inline std::vector<double> toReals(const std::string& s)
{
  std::vector<double> out;
  for (const auto& w : splitWords(s))
    out.push_back(std::stod(w));
  return out;
}

inline HitNode parseHit(const std::string& text)
{
  HitNode root{"", {}, {}};
  std::vector<HitNode*> stack{&root};
  std::istringstream is(text);
  std::string line, pending_key, pending_val;
  char open_quote = 0;
  int lineno = 0;
  auto store = [&](const std::string& k, std::string v) {
    v = trim(v);
    if (v.size() >= 2 && (v.front() == '\'' || v.front() == '"') && v.back() == v.front())
      v = v.substr(1, v.size() - 2);
    stack.back()->params.emplace_back(k, trim(v));
  };
  while (std::getline(is, line))
  {
    ++lineno;
    if (open_quote) // continuation of a multi-line quoted value
    {
      const size_t close = line.find(open_quote);
      if (close == std::string::npos)
        pending_val += " " + line;
      else
      {
        pending_val += " " + line.substr(0, close + 1); // whatever follows the closing quote is a comment
        store(pending_key, pending_val);
        open_quote = 0;
      }
      continue;
    }
    // strip comments outside quotes
    {
      char q = 0;
      for (size_t i = 0; i < line.size(); ++i)
      {
        if (!q && (line[i] == '\'' || line[i] == '"'))
          q = line[i];
        else if (q && line[i] == q)
          q = 0;
        else if (!q && line[i] == '#')
        {
          line.erase(i);
          break;
        }
      }
    }
    line = trim(line);
    if (line.empty())
      continue;
    if (line.front() == '[')
    {
      const size_t e = line.find(']');
      if (e == std::string::npos)
        mooseError("input line " + std::to_string(lineno) + ": missing ]");
      std::string n = trim(line.substr(1, e - 1));
      if (n.empty() || n == "../" || n == "..")
      {
        if (stack.size() == 1)
          mooseError("input line " + std::to_string(lineno) + ": unmatched block end");
        // a [A/B] header pushed one level per path component: its [] pops them all
        HitNode* closing = stack.back();
        stack.pop_back();
        while (stack.size() > 1 && closing->name.rfind("\x01", 0) == 0)
        {
          closing->name.erase(0, 1);
          closing = stack.back();
          stack.pop_back();
        }
        continue;
      }
      if (n.rfind("./", 0) == 0)
        n = n.substr(2);
      // [A/B/C] opens nested blocks; inner ones are marked so that a single [] closes the whole path
      std::vector<std::string> parts;
      {
        std::istringstream ps(n);
        for (std::string part; std::getline(ps, part, '/');)
          if (!part.empty())
            parts.push_back(part);
      }
      for (size_t k = 0; k < parts.size(); ++k)
      {
        HitNode& c = stack.back()->childOrNew(parts[k]);
        if (k > 0)
          c.name = "\x01" + c.name;
        stack.push_back(&c);
      }
      continue;
    }
    const size_t eq = line.find('=');
    if (eq == std::string::npos)
      mooseError("input line " + std::to_string(lineno) + ": expected key = value");
    const std::string key = trim(line.substr(0, eq));
    const std::string val = trim(line.substr(eq + 1));
    if (!val.empty() && (val.front() == '\'' || val.front() == '"') && (val.size() == 1 || val.find(val.front(), 1) == std::string::npos))
    {
      open_quote = val.front();
      pending_key = key;
      pending_val = val;
      continue;
    }
    store(key, val);
  }
  if (stack.size() != 1)
    mooseError("input: block [" + stack.back()->name + "] is not closed");
  return root;
}

// ------------------------------------------------------------------ XML subset (OpenMC inputs)
struct XmlNode
{
  std::string tag, text;
  std::map<std::string, std::string> attr;
  std::vector<XmlNode> children;
  const XmlNode* child(const std::string& t) const
  {
    for (const auto& c : children)
      if (c.tag == t)
        return &c;
    return nullptr;
  }
  // OpenMC accepts a value as attribute or as child element (OMC xml_interface.cpp:17-36)
  std::string value(const std::string& name, const std::string& dflt = "") const
  {
    auto it = attr.find(name);
    if (it != attr.end())
      return it->second;
    const XmlNode* c = child(name);
    return c ? trim(c->text) : dflt;
  }
};

inline XmlNode parseXml(const std::string& s)
{
  XmlNode root;
  std::vector<XmlNode*> stack{&root};
  size_t i = 0;
  while (i < s.size())
  {
    if (s[i] != '<')
    {
      const size_t e = s.find('<', i);
      stack.back()->text += s.substr(i, e == std::string::npos ? std::string::npos : e - i);
      i = e == std::string::npos ? s.size() : e;
      continue;
    }
    if (s.compare(i, 4, "<!--") == 0)
    {
      const size_t e = s.find("-->", i);
      i = e == std::string::npos ? s.size() : e + 3;
      continue;
    }
    if (s.compare(i, 2, "<?") == 0)
    {
      const size_t e = s.find("?>", i);
      i = e == std::string::npos ? s.size() : e + 2;
      continue;
    }
    const size_t e = s.find('>', i);
    if (e == std::string::npos)
      mooseError("xml: unterminated tag");
    std::string body = s.substr(i + 1, e - i - 1);
    i = e + 1;
    if (!body.empty() && body[0] == '/')
    {
      if (stack.size() == 1)
        mooseError("xml: unmatched closing tag " + body);
      stack.pop_back();
      continue;
    }
    const bool self = !body.empty() && body.back() == '/';
    if (self)
      body.pop_back();
    XmlNode n;
    size_t p = body.find_first_of(" \t\r\n");
    n.tag = body.substr(0, p);
    while (p != std::string::npos && p < body.size())
    {
      const size_t k0 = body.find_first_not_of(" \t\r\n", p);
      if (k0 == std::string::npos)
        break;
      const size_t eq = body.find('=', k0);
      if (eq == std::string::npos)
        break;
      const size_t q0 = body.find_first_of("\"'", eq);
      if (q0 == std::string::npos)
        break;
      const size_t q1 = body.find(body[q0], q0 + 1);
      if (q1 == std::string::npos)
        mooseError("xml: unterminated attribute in <" + n.tag + ">");
      n.attr[trim(body.substr(k0, eq - k0))] = body.substr(q0 + 1, q1 - q0 - 1);
      p = q1 + 1;
    }
    stack.back()->children.push_back(std::move(n));
    if (!self)
      stack.push_back(&stack.back()->children.back());
  }
  return root;
}

inline std::string readFile(const std::string& path)
{
  std::ifstream f(path);
  if (!f)
    mooseError("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// ------------------------------------------------------------------ GeneratedMeshGenerator
// Node / element / side numbering of libMesh's build_cube as stated in-tree by MOOSE's
// DistributedRectilinearMeshGenerator.C:652-776: node id = i + (nx+1)(j + k(ny+1));
// HEX8 boundary ids back=0 bottom=1 right=2 top=3 left=4 front=5; QUAD4 bottom=0 right=1 top=2 left=3.
inline Mesh generatedMesh(int dim, int nx, int ny, int nz, const double lo[3], const double hi[3])
{
  if (dim != 2 && dim != 3)
    mooseError("GeneratedMeshGenerator: dim must be 2 or 3 on this path");
  Mesh m;
  m.elem_type = dim == 3 ? MCW_HEX8 : MCW_QUAD4;
  if (dim == 2)
    nz = 0;
  auto lin = [](double a, double b, int n, int i) { return n == 0 ? a : a + (b - a) * ((double)i / n == 1.0 ? 1.0 : (double)i / n); };
  for (int k = 0; k <= nz; ++k)
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i)
      {
        // numpy.linspace arithmetic (start + i * step, last point exact), like macaw_b200/meshgen.py
        const double sx = (hi[0] - lo[0]) / nx, sy = (hi[1] - lo[1]) / ny, sz = nz ? (hi[2] - lo[2]) / nz : 0.0;
        m.xyz.push_back(i == nx ? hi[0] : lo[0] + i * sx);
        m.xyz.push_back(j == ny ? hi[1] : lo[1] + j * sy);
        m.xyz.push_back(dim == 3 ? (k == nz ? hi[2] : lo[2] + k * sz) : 0.0);
      }
  (void)lin;
  auto nid = [&](int i, int j, int k) { return (int64_t)i + (int64_t)(nx + 1) * (j + (int64_t)k * (ny + 1)); };
  const int ns = dim == 3 ? 6 : 4;
  for (int k = 0; k < std::max(nz, 1); ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i)
      {
        if (dim == 3)
        {
          const int64_t c[8] = {nid(i, j, k),     nid(i + 1, j, k),     nid(i + 1, j + 1, k),     nid(i, j + 1, k),
                                nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)};
          m.conn.insert(m.conn.end(), c, c + 8);
          const int b[6] = {k == 0 ? 0 : -1,      j == 0 ? 1 : -1, i == nx - 1 ? 2 : -1,
                            j == ny - 1 ? 3 : -1, i == 0 ? 4 : -1, k == nz - 1 ? 5 : -1};
          m.side_bnd_id.insert(m.side_bnd_id.end(), b, b + 6);
        }
        else
        {
          const int64_t c[4] = {nid(i, j, 0), nid(i + 1, j, 0), nid(i + 1, j + 1, 0), nid(i, j + 1, 0)};
          m.conn.insert(m.conn.end(), c, c + 4);
          const int b[4] = {j == 0 ? 0 : -1, i == nx - 1 ? 1 : -1, j == ny - 1 ? 2 : -1, i == 0 ? 3 : -1};
          m.side_bnd_id.insert(m.side_bnd_id.end(), b, b + 4);
        }
        m.block_id.push_back(0);
      }
  m.bnd_kind.assign(ns, MCW_BC_KILL); // a boundary without a RayBC kills the ray (RT TraceRay.C:2100-2146 errors instead)
  return m;
}

// QUAD4 -> HEX8 extrusion with the numbering of MOOSE's FancyExtruderGenerator
// (moose/framework/src/meshgenerators/FancyExtruderGenerator.C:255-285, 444-462, 551, 604-646): node id =
// node + plane * n_nodes; element id = elem + layer * n_elem; Hex8 nodes = the quad's nodes on the lower plane,
// then on the upper one; quad side s -> hex side s + 1 with its boundary id; hex side 0 of the first layer and
// side 5 of the last one get the bottom / top boundary ids.  Plane positions accumulate height / num_layers.
inline Mesh extrudeQuads(const Mesh& q, const std::vector<int>& num_layers, const std::vector<double>& heights, const double dir[3],
                         int bottom_boundary, int top_boundary)
{
  if (q.elem_type != MCW_QUAD4)
    mooseError("FancyExtruderGenerator: only QUAD4 input meshes are extruded by this front end");
  if (num_layers.size() != heights.size() || num_layers.empty())
    mooseError("FancyExtruderGenerator: num_layers and heights must have the same, non-zero, length");
  std::vector<double> px{0.0}, py{0.0}, pz{0.0};
  for (size_t e = 0; e < num_layers.size(); ++e)
    for (int k = 0; k < num_layers[e]; ++k)
    {
      const double step = heights[e] / (double)num_layers[e];
      px.push_back(px.back() + dir[0] * step);
      py.push_back(py.back() + dir[1] * step);
      pz.push_back(pz.back() + dir[2] * step);
    }
  const int64_t L = (int64_t)px.size() - 1, nn = q.nNodes(), ne = q.nElem();
  Mesh h;
  h.elem_type = MCW_HEX8;
  for (int64_t k = 0; k <= L; ++k)
    for (int64_t i = 0; i < nn; ++i)
    {
      h.xyz.push_back(q.xyz[3 * i] + px[k]);
      h.xyz.push_back(q.xyz[3 * i + 1] + py[k]);
      h.xyz.push_back(q.xyz[3 * i + 2] + pz[k]);
    }
  int n_bnd = (int)q.bnd_kind.size();
  for (int64_t k = 0; k < L; ++k)
    for (int64_t e = 0; e < ne; ++e)
    {
      for (int up = 0; up < 2; ++up)
        for (int v = 0; v < 4; ++v)
          h.conn.push_back(q.conn[4 * e + v] + (k + up) * nn);
      h.block_id.push_back(q.block_id[e]);
      int32_t sb[6] = {-1, q.side_bnd_id[4 * e], q.side_bnd_id[4 * e + 1], q.side_bnd_id[4 * e + 2], q.side_bnd_id[4 * e + 3], -1};
      if (k == 0 && bottom_boundary >= 0)
        sb[0] = bottom_boundary;
      if (k == L - 1 && top_boundary >= 0)
        sb[5] = top_boundary;
      h.side_bnd_id.insert(h.side_bnd_id.end(), sb, sb + 6);
    }
  n_bnd = std::max(n_bnd, std::max(bottom_boundary, top_boundary) + 1);
  h.bnd_kind.assign((size_t)n_bnd, MCW_BC_KILL);
  std::copy(q.bnd_kind.begin(), q.bnd_kind.end(), h.bnd_kind.begin());
  return h;
}

inline int boundaryIdByName(int dim, const std::string& n)
{
  static const char* n3[6] = {"back", "bottom", "right", "top", "left", "front"};
  static const char* n2[4] = {"bottom", "right", "top", "left"};
  for (int k = 0; k < (dim == 3 ? 6 : 4); ++k)
    if (n == (dim == 3 ? n3[k] : n2[k]))
      return k;
  try
  {
    return std::stoi(n);
  }
  catch (...)
  {
    mooseError("unknown boundary '" + n + "'");
  }
  return -1;
}

// ------------------------------------------------------------------ nuclide table file
// Same record format as the problem files; names as one ';'-joined int32 character record.
struct NuclideTables
{
  std::vector<std::string> names;
  std::vector<Nuclide> nuclides;
  int find(const std::string& n) const
  {
    for (size_t k = 0; k < names.size(); ++k)
      if (names[k] == n)
        return (int)k;
    return -1;
  }
};

inline NuclideTables readNuclideTables(const std::string& path)
{
  ProblemFile f(path);
  NuclideTables T;
  {
    std::string joined;
    for (int32_t c : f.i32("lib.names"))
      joined.push_back((char)c);
    std::istringstream is(joined);
    for (std::string n; std::getline(is, n, ';');)
      T.names.push_back(n);
  }
  for (size_t k = 0; k < T.names.size(); ++k)
  {
    const std::string p = "nuc" + std::to_string(k) + ".";
    Nuclide n;
    const auto I = f.i32(p + "ints");
    const auto S = f.f64(p + "scalars");
    n.energy = f.f64(p + "energy");
    n.xs = f.f64(p + "xs");
    n.elastic = f.f64(p + "elastic");
    n.inelastic = f.f64(p + "inelastic");
    n.d.fissionable = I[1];
    n.d.inel_threshold = I[3];
    n.d.n_precursor = I[4];
    n.d.awr = S[0], n.d.kT = S[1], n.d.inel_threshold_energy = S[2], n.d.inel_mass_ratio = S[3];
    n.d.nu_poly[0] = S[4], n.d.nu_poly[1] = S[5], n.d.watt_a = S[6], n.d.watt_b = S[7], n.d.q_fission = S[8];
    const auto dy = f.f64(p + "delayed_yield"), dt = f.f64(p + "delayed_theta");
    for (int g = 0; g < MCW_MAX_DELAYED_GROUPS; ++g)
      n.d.delayed_yield[g] = dy[g], n.d.delayed_theta[g] = dt[g];
    T.nuclides.push_back(std::move(n));
  }
  return T;
}

// ------------------------------------------------------------------ input -> Problem
struct InputProblem
{
  Problem problem;
  int batches = 0; // [Executioner] num_steps, else settings.xml <batches>
  struct AuxDef
  {
    std::string name, granularity, score;
    int tally_id;
    int energy_bin; // -1 = all
    std::string nuclide = "all"; // OpenMCTallyAux.C:44,61-62
  };
  std::vector<AuxDef> aux;
  std::vector<std::vector<std::string>> tally_scores, tally_filters; // as written, parallel to problem.tallies
  std::vector<std::string> tally_estimators;
  std::vector<std::vector<std::string>> tally_nuclides; // OpenMCTally `nuclides` (default "total")
};

inline InputProblem readInput(const std::string& input_path, const std::string& xml_dir, const std::string& tables_path)
{
  InputProblem R;
  Problem& P = R.problem;
  const HitNode in = parseHit(readFile(input_path));

  // ---- [RayBCs] (read first: a file mesh needs the side-set ids that carry a boundary condition)
  std::vector<std::pair<int, std::vector<std::string>>> ray_bcs; // (kind, boundary names or ids)
  if (const HitNode* bc = in.child("RayBCs"))
    for (const auto& b : bc->children)
    {
      if (!blockIsActive(*bc, b.name))
        continue;
      const std::string type = b.get("type");
      const int kind = type == "ReflectRayBC" ? MCW_BC_REFLECT : type == "KillRayBC" ? MCW_BC_KILL : -1;
      if (kind < 0)
        mooseError("RayBC '" + type + "' is not on this path");
      ray_bcs.emplace_back(kind, splitWords(b.get("boundary")));
    }

  // ---- [Mesh]
  const HitNode* mesh = in.child("Mesh");
  if (!mesh)
    mooseError("the input has no [Mesh] block");
  bool have_mesh = false, file_mesh = false;
  int dim = 3;
  const size_t slash = input_path.find_last_of('/');
  const std::string input_dir = slash == std::string::npos ? "." : input_path.substr(0, slash);
  std::vector<const HitNode*> generators;
  for (const auto& g : mesh->children)
    generators.push_back(&g);
  if (mesh->find("type"))
    generators.insert(generators.begin(), mesh);
  for (const HitNode* g : generators)
  {
    const std::string type = g->get("type");
    if (type == "GeneratedMeshGenerator" || type == "GeneratedMesh")
    {
      dim = std::stoi(g->get("dim", "3"));
      const double lo[3] = {std::stod(g->get("xmin", "0")), std::stod(g->get("ymin", "0")), std::stod(g->get("zmin", "0"))};
      const double hi[3] = {std::stod(g->get("xmax", "1")), std::stod(g->get("ymax", "1")), std::stod(g->get("zmax", "1"))};
      P.mesh = generatedMesh(dim, std::stoi(g->get("nx", "1")), std::stoi(g->get("ny", "1")), std::stoi(g->get("nz", "1")), lo, hi);
      have_mesh = true;
    }
    else if (type == "FileMeshGenerator" || type == "FileMesh")
    {
      std::vector<int> reflect, kill;
      for (const auto& bc : ray_bcs)
        for (const auto& n : bc.second)
        {
          int id = -1;
          try
          {
            id = std::stoi(n);
          }
          catch (...)
          {
            mooseError("boundary '" + n + "': a file mesh's side sets are addressed by id here");
          }
          (bc.first == MCW_BC_REFLECT ? reflect : kill).push_back(id);
        }
      std::string file = g->get("file");
      if (!file.empty() && file[0] != '/')
        file = input_dir + "/" + file;
      P.mesh = readExodus(file, reflect, kill);
      dim = P.mesh.dimension();
      have_mesh = file_mesh = true;
    }
    else if (type == "TransformGenerator")
    {
      if (!have_mesh)
        mooseError("TransformGenerator before its input mesh");
      if (g->get("transform") != "TRANSLATE")
        mooseError("TransformGenerator: only TRANSLATE is supported by this front end");
      const auto v = toReals(g->get("vector_value"));
      if (v.size() != 3)
        mooseError("TransformGenerator: vector_value needs three components");
      // (a 2-D run ignores z: the shift of examples/beavrs_assembly only matters through z_coord or an extrusion)
      for (size_t i = 0; i < P.mesh.xyz.size() / 3; ++i)
        for (int d = 0; d < 3; ++d)
          P.mesh.xyz[3 * i + d] += v[d];
    }
    else if (type == "FancyExtruderGenerator" || type == "AdvancedExtruderGenerator")
    {
      if (!have_mesh)
        mooseError("FancyExtruderGenerator before its input mesh");
      std::vector<int> nl;
      for (const auto& w : splitWords(g->get("num_layers")))
        nl.push_back(std::stoi(w));
      const auto dirv = toReals(g->get("direction"));
      if (dirv.size() != 3)
        mooseError("FancyExtruderGenerator: direction needs three components");
      const double d3[3] = {dirv[0], dirv[1], dirv[2]};
      const std::vector<int32_t> kinds = P.mesh.bnd_kind;
      P.mesh = extrudeQuads(P.mesh, nl, toReals(g->get("heights")), d3, g->find("bottom_boundary") ? std::stoi(g->get("bottom_boundary")) : -1,
                            g->find("top_boundary") ? std::stoi(g->get("top_boundary")) : -1);
      dim = 3;
      // boundary ids the extruder created may carry RayBCs too
      for (const auto& bc : ray_bcs)
        for (const auto& n : bc.second)
        {
          const int id = std::stoi(n);
          if (id >= 0 && id < (int)P.mesh.bnd_kind.size())
            P.mesh.bnd_kind[id] = bc.first;
        }
    }
    else if (type == "SubdomainBoundingBoxGenerator")
    {
      if (!have_mesh)
        mooseError("SubdomainBoundingBoxGenerator before its input mesh");
      const auto a = toReals(g->get("bottom_left")), b = toReals(g->get("top_right"));
      const int id = std::stoi(g->get("block_id"));
      const int nv = P.mesh.nodesPerElem();
      for (int64_t e = 0; e < P.mesh.nElem(); ++e)
      {
        double c[3] = {0, 0, 0};
        for (int v = 0; v < nv; ++v)
          for (int d = 0; d < 3; ++d)
            c[d] += P.mesh.xyz[3 * P.mesh.conn[e * nv + v] + d];
        bool inside = true;
        for (int d = 0; d < dim; ++d)
          inside &= c[d] / nv >= a[d] && c[d] / nv <= b[d];
        if (inside)
          P.mesh.block_id[e] = id;
      }
    }
    else if (!type.empty())
      mooseError("mesh generator '" + type + "' is not supported by this front end");
  }
  if (!have_mesh)
    mooseError("[Mesh] holds no GeneratedMeshGenerator or FileMeshGenerator");

  // ---- [Variables/temperature]
  double T0 = 0.0; // MOOSE default initial condition (the study test runs with T = 0)
  for (const char* blk : {"Variables", "AuxVariables"}) // the examples declare it either way
    if (const HitNode* v = in.child(blk))
      if (const HitNode* t = v->child("temperature"))
        T0 = std::stod(t->get("initial_condition", "0"));

  // ---- [RayKernels]: the CollisionKernel
  const HitNode* ck = nullptr;
  if (const HitNode* rk = in.child("RayKernels"))
    for (const auto& k : rk->children)
      if (k.get("type") == "CollisionKernel" && blockIsActive(*rk, k.name))
        ck = &k;
  if (!ck)
    mooseError("the input has no CollisionKernel in [RayKernels]");
  for (const auto& w : splitWords(ck->get("blocks")))
    P.blocks.push_back((unsigned)std::stoul(w));
  for (const auto& w : splitWords(ck->get("materials")))
    P.materials.push_back((unsigned)std::stoul(w));
  {
    // `temperature` is a coupled variable name, or a constant (examples/beavrs_assembly/run_assembly_3d.i:62)
    const std::string t = ck->get("temperature");
    char* end = nullptr;
    const double c = std::strtod(t.c_str(), &end);
    if (!t.empty() && end && *end == '\0')
      T0 = c;
  }
  P.temperature.assign(1, T0);
  P.settings.z_coord = std::stod(ck->get("z_coord", "0"));
  if (dim == 2 && !ck->find("z_coord"))
    mooseError("The z_coord parameter must be specified for 2D calculations"); // MC CollisionKernel.C:60-61
  // blocks named by the kernel but absent from the mesh are legal (examples list "0 1 2" for a 2-block mesh)

  // ---- [RayBCs] on a generated mesh: boundaries by name or id (a file mesh got them in readExodus)
  if (!file_mesh)
    for (const auto& bc : ray_bcs)
      for (const auto& n : bc.second)
      {
        const int id = boundaryIdByName(dim, n);
        if (id < 0 || id >= (int)P.mesh.bnd_kind.size())
          mooseError("boundary '" + n + "' does not exist on the mesh");
        P.mesh.bnd_kind[id] = bc.first;
      }

  // ---- [UserObjects]: OpenMCStudy (present), OpenMCTally
  bool have_study = false;
  if (const HitNode* uo = in.child("UserObjects"))
    for (const auto& u : uo->children)
    {
      if (!blockIsActive(*uo, u.name))
        continue;
      const std::string type = u.get("type");
      if (type == "OpenMCStudy")
        have_study = true;
      else if (type == "OpenMCTally")
      {
        Problem::TallyDef t;
        t.id = std::stoi(u.get("id"));
        t.estimator = MCW_ESTIMATOR_COLLISION; // validated by macaw::OpenMCTally
        t.granularity = MCW_FILTER_GLOBAL;
        P.tallies.push_back(t);
        R.tally_scores.push_back(splitWords(u.get("scores")));
        R.tally_filters.push_back(splitWords(u.get("filters")));
        R.tally_estimators.push_back(u.get("estimator", "COLLISION"));
        R.tally_nuclides.push_back(splitWords(u.get("nuclides", "total")));
        P.tallies.back().energy_bins = toReals(u.get("energy_bins"));
      }
    }
  if (!have_study)
    mooseError("the input has no OpenMCStudy in [UserObjects]");

  // ---- [AuxKernels]: OpenMCTallyAux
  if (const HitNode* ak = in.child("AuxKernels"))
    for (const auto& a : ak->children)
      if (a.get("type") == "OpenMCTallyAux" && blockIsActive(*ak, a.name))
        R.aux.push_back({a.name, a.get("granularity"), a.get("score"), std::stoi(a.get("tally_id", "-1")),
                         a.find("energy_bin") ? std::stoi(a.get("energy_bin")) : -1, a.get("nuclide", "all")});

  // ---- settings.xml (OMC settings.cpp:230-420)
  const XmlNode sx = parseXml(readFile(xml_dir + "/settings.xml"));
  const XmlNode* st = sx.child("settings");
  if (!st)
    mooseError("settings.xml has no <settings>");
  if (st->value("run_mode", "eigenvalue") != "eigenvalue")
    mooseError("only run_mode eigenvalue is on this path");
  mcw_settings& s = P.settings;
  s.n_particles = std::stoll(st->value("particles"));
  s.n_inactive = std::stoi(st->value("inactive", "0"));
  s.master_seed = std::stoll(st->value("seed", "1"));
  s.n_log_bins = std::stoi(st->value("log_grid_bins", "8000"));
  s.is_2d = dim == 2;
  s.seed_through_double = 1;
  s.max_events = 1000000; // OMC constants.h:56
  s.source_only_fissionable = 0;
  s.record_traces = 0;
  R.batches = std::stoi(st->value("batches", "1"));
  if (const HitNode* ex = in.child("Executioner"))
    if (ex->find("num_steps"))
      R.batches = std::stoi(ex->get("num_steps"));
  const XmlNode* src = st->child("source");
  const XmlNode* space = src ? src->child("space") : nullptr;
  if (!space || (space->value("type") != "box" && space->value("type") != "fission"))
    mooseError("settings.xml: the source must be a box (or fission box) <space>");
  {
    const auto b = toReals(space->value("parameters"));
    if (b.size() != 6)
      mooseError("settings.xml: box source needs 6 parameters");
    for (int k = 0; k < 6; ++k)
      s.source_box[k] = b[k];
    s.source_only_fissionable = space->value("type") == "fission";
  }

  // ---- materials.xml + nuclide tables (OMC material.cpp:50-412)
  const NuclideTables T = readNuclideTables(tables_path);
  std::map<std::string, int> lib_index; // nuclides in order of first use (OMC data::nuclide_map)
  const XmlNode mx = parseXml(readFile(xml_dir + "/materials.xml"));
  const XmlNode* mats = mx.child("materials");
  if (!mats)
    mooseError("materials.xml has no <materials>");
  std::map<int, int> index_of_id;
  for (const auto& m : mats->children)
  {
    if (m.tag != "material")
      continue;
    Material M;
    std::vector<double> frac;
    std::vector<std::string> names;
    bool weight = false;
    for (const auto& n : m.children)
      if (n.tag == "nuclide")
      {
        const std::string name = n.value("name");
        const int t = T.find(name);
        if (t < 0)
          mooseError("nuclide " + name + " is not in the table file " + tables_path);
        if (!lib_index.count(name))
        {
          lib_index[name] = (int)P.library.nuclides.size();
          P.library.nuclides.push_back(T.nuclides[t]);
          P.library.nuclides.back().name = name;
        }
        M.nuclide.push_back(lib_index[name]);
        names.push_back(name);
        if (n.attr.count("wo"))
        {
          weight = true;
          frac.push_back(-std::stod(n.value("wo")));
        }
        else
          frac.push_back(std::stod(n.value("ao")));
      }
    if (M.nuclide.empty())
      mooseError("material without nuclides");
    // Material::normalize_density (OMC material.cpp:366-412)
    const XmlNode* dn = m.child("density");
    const std::string units = dn ? dn->value("units") : "sum";
    double density = dn && units != "sum" ? std::stod(dn->value("value")) : 0.0;
    auto awr = [&](size_t k) { return P.library.nuclides[M.nuclide[k]].d.awr; };
    if (weight)
      for (size_t k = 0; k < frac.size(); ++k)
        frac[k] = -frac[k] / awr(k);
    double sum = 0.0;
    for (double f : frac)
      sum += f;
    if (units == "sum")
      density = sum;
    else if (units == "g/cm3" || units == "g/cc")
      density = -density;
    else if (units == "kg/m3")
      density = -density * 1e-3;
    else if (units != "atom/b-cm")
      mooseError("density units '" + units + "' are not supported");
    for (double& f : frac)
      f /= sum;
    if (density < 0.0)
    {
      double sp = 0.0;
      for (size_t k = 0; k < frac.size(); ++k)
        sp += frac[k] * awr(k);
      density = -density * 0.602214076 / 1.00866491595 * (1.0 / sp); // N_AVOGADRO / MASS_NEUTRON, OMC constants.h:81,89
    }
    for (double f : frac)
      M.atom_density.push_back(f * density);
    if (const XmlNode* sab = m.child("sab"))
      std::fprintf(stderr, "note: material %s: S(alpha,beta) table %s ignored (thermal scattering data are not on this path yet)\n",
                   m.value("id").c_str(), sab->value("name").c_str());
    if (const XmlNode* iso = m.child("isotropic"))
    {
      const auto list = splitWords(iso->text);
      for (const auto& nm : names)
        M.isotropic.push_back(std::find(list.begin(), list.end(), nm) != list.end());
    }
    index_of_id[std::stoi(m.value("id"))] = (int)P.library.materials.size();
    P.library.materials.push_back(std::move(M));
  }
  // CollisionKernel's `materials` are OpenMC ids; the study resolves them through material_map (MC CollisionKernel.C:110)
  for (auto& id : P.materials)
  {
    auto it = index_of_id.find((int)id);
    if (it == index_of_id.end())
      mooseError("CollisionKernel: OpenMC material id " + std::to_string(id) + " is not in materials.xml");
    id = (unsigned)it->second + 1; // macaw::CollisionKernel::materialIndices subtracts the 1 again
  }
  return R;
}

} // namespace macaw
