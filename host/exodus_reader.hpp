// exodus_reader.hpp -- Exodus-II meshes stored as classic NetCDF (CDF-1 / CDF-2), the files MaCaw's
// FileMeshGenerator inputs point at (examples/beavrs_assembly/31enr_16_in.e), without libMesh or
// libnetcdf: a minimal reader of the classic header (dimensions, attributes, fixed-size variables,
// big-endian data) and the Exodus naming conventions on top of it (coordx/y/z or coord, connectN with
// its elem_type attribute, eb_prop1, ss_prop1, elem_ssN / side_ssN).  Mirrors macaw_b200/exodus.py.
// Exodus node / side orders of QUAD4, TETRA4 and HEX8 are the libMesh ones.  Set-up code only.
#pragma once
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>

#include "macaw_host.hpp"

namespace macaw
{

class NetcdfClassic
{
public:
  struct Var
  {
    std::vector<int64_t> shape;
    int type = 0; // 1 byte, 2 char, 3 short, 4 int, 5 float, 6 double
    uint64_t begin = 0;
    std::map<std::string, std::string> text_attrs;
    int64_t count() const
    {
      int64_t n = 1;
      for (int64_t d : shape)
        n *= d;
      return n;
    }
  };

  explicit NetcdfClassic(const std::string& path)
  {
    std::ifstream f(path, std::ios::binary);
    if (!f)
      mooseError("cannot open mesh file " + path);
    _buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (_buf.size() < 8 || std::memcmp(_buf.data(), "CDF", 3) != 0 || (_buf[3] != 1 && _buf[3] != 2))
      mooseError(path + " is not a classic NetCDF (CDF-1 / CDF-2) file; HDF5-based Exodus files need libnetcdf");
    const bool off64 = _buf[3] == 2;
    _pos = 4;
    u32(); // numrecs
    // dimensions
    std::vector<int64_t> dim_len;
    {
      const uint32_t tag = u32(), n = u32();
      if (tag != 0 && tag != 0x0A)
        mooseError("netcdf: bad dimension list");
      for (uint32_t k = 0; k < n; ++k)
      {
        const std::string name = str();
        const int64_t len = u32();
        _dims[name] = len;
        dim_len.push_back(len);
      }
    }
    attrs(nullptr); // global attributes
    {
      const uint32_t tag = u32(), n = u32();
      if (tag != 0 && tag != 0x0B)
        mooseError("netcdf: bad variable list");
      for (uint32_t k = 0; k < n; ++k)
      {
        const std::string name = str();
        Var v;
        const uint32_t nd = u32();
        for (uint32_t d = 0; d < nd; ++d)
          v.shape.push_back(dim_len.at(u32()));
        attrs(&v);
        v.type = (int)u32();
        u32(); // vsize
        v.begin = off64 ? u64() : u32();
        _vars[name] = std::move(v);
      }
    }
  }

  bool has(const std::string& n) const { return _vars.count(n) != 0; }
  int64_t dim(const std::string& n, int64_t dflt = -1) const
  {
    auto it = _dims.find(n);
    return it == _dims.end() ? dflt : it->second;
  }
  const Var& var(const std::string& n) const
  {
    auto it = _vars.find(n);
    if (it == _vars.end())
      mooseError("mesh file lacks variable " + n);
    return it->second;
  }
  std::vector<double> reals(const std::string& n) const
  {
    const Var& v = var(n);
    std::vector<double> out((size_t)v.count());
    for (size_t k = 0; k < out.size(); ++k)
    {
      if (v.type == 6)
      {
        uint64_t b = be<uint64_t>(v.begin + 8 * k);
        std::memcpy(&out[k], &b, 8);
      }
      else if (v.type == 5)
      {
        uint32_t b = be<uint32_t>(v.begin + 4 * k);
        float x;
        std::memcpy(&x, &b, 4);
        out[k] = x;
      }
      else
        mooseError("variable " + n + " is not floating point");
    }
    return out;
  }
  std::vector<int64_t> ints(const std::string& n) const
  {
    const Var& v = var(n);
    if (v.type != 4)
      mooseError("variable " + n + " is not a 32-bit integer array");
    std::vector<int64_t> out((size_t)v.count());
    for (size_t k = 0; k < out.size(); ++k)
      out[k] = (int32_t)be<uint32_t>(v.begin + 4 * k);
    return out;
  }

private:
  template <typename T>
  T be(uint64_t at) const
  {
    if (at + sizeof(T) > _buf.size())
      mooseError("netcdf: truncated file");
    T v = 0;
    for (size_t k = 0; k < sizeof(T); ++k)
      v = (T)((v << 8) | (unsigned char)_buf[at + k]);
    return v;
  }
  uint32_t u32()
  {
    const uint32_t v = be<uint32_t>(_pos);
    _pos += 4;
    return v;
  }
  uint64_t u64()
  {
    const uint64_t v = be<uint64_t>(_pos);
    _pos += 8;
    return v;
  }
  std::string str()
  {
    const uint32_t n = u32();
    std::string s(_buf.data() + _pos, _buf.data() + _pos + n);
    _pos += (n + 3) & ~3u;
    return s;
  }
  void attrs(Var* v)
  {
    const uint32_t tag = u32(), n = u32();
    if (tag != 0 && tag != 0x0C)
      mooseError("netcdf: bad attribute list");
    static const int size_of[7] = {0, 1, 1, 2, 4, 4, 8};
    for (uint32_t k = 0; k < n; ++k)
    {
      const std::string name = str();
      const uint32_t type = u32(), count = u32();
      if (type < 1 || type > 6)
        mooseError("netcdf: bad attribute type");
      const uint64_t bytes = (uint64_t)count * size_of[type];
      if (v && type == 2)
        v->text_attrs[name] = std::string(_buf.data() + _pos, _buf.data() + _pos + bytes);
      _pos += (bytes + 3) & ~(uint64_t)3;
    }
  }
  std::vector<char> _buf;
  uint64_t _pos = 0;
  std::map<std::string, int64_t> _dims;
  std::map<std::string, Var> _vars;
};

// Exodus -> Mesh.  reflect / kill: side-set ids carrying ReflectRayBC / KillRayBC; boundary sides in
// no listed set get no id (the ray is killed there).
inline Mesh readExodus(const std::string& path, const std::vector<int>& reflect, const std::vector<int>& kill)
{
  const NetcdfClassic f(path);
  Mesh m;
  const int64_t dim = f.dim("num_dim"), n_nodes = f.dim("num_nodes"), n_blk = f.dim("num_el_blk");
  if (dim < 2 || dim > 3 || n_nodes <= 0 || n_blk <= 0)
    mooseError(path + ": not an Exodus-II mesh");
  m.xyz.assign((size_t)n_nodes * 3, 0.0);
  if (f.has("coord"))
  {
    const auto c = f.reals("coord"); // [num_dim][num_nodes]
    for (int64_t d = 0; d < dim; ++d)
      for (int64_t i = 0; i < n_nodes; ++i)
        m.xyz[3 * i + d] = c[d * n_nodes + i];
  }
  else
  {
    static const char* names[3] = {"coordx", "coordy", "coordz"};
    for (int64_t d = 0; d < dim; ++d)
    {
      const auto c = f.reals(names[d]);
      for (int64_t i = 0; i < n_nodes; ++i)
        m.xyz[3 * i + d] = c[i];
    }
  }
  const auto block_ids = f.ints("eb_prop1");
  int etype = -1;
  for (int64_t b = 0; b < n_blk; ++b)
  {
    const std::string name = "connect" + std::to_string(b + 1);
    const NetcdfClassic::Var& v = f.var(name);
    std::string t = v.text_attrs.count("elem_type") ? v.text_attrs.at("elem_type") : "";
    t.erase(std::find(t.begin(), t.end(), '\0'), t.end());
    std::transform(t.begin(), t.end(), t.begin(), ::toupper);
    const int type = (t == "QUAD4" || t == "QUAD") ? MCW_QUAD4
                     : (t == "TETRA4" || t == "TETRA" || t == "TET4") ? MCW_TET4
                     : (t == "HEX8" || t == "HEX") ? MCW_HEX8 : -1;
    if (type < 0)
      mooseError("element type '" + t + "' is not traced by this path (QUAD4, TET4, HEX8)");
    if (etype >= 0 && type != etype)
      mooseError("mixed element types in one mesh");
    etype = type;
    for (int64_t node : f.ints(name))
      m.conn.push_back(node - 1);
    m.block_id.insert(m.block_id.end(), (size_t)v.shape.at(0), (int32_t)block_ids[b]);
  }
  m.elem_type = etype;
  const int ns = m.sidesPerElem();
  m.side_bnd_id.assign((size_t)m.nElem() * ns, -1);
  // a side set only marks boundary sides here; whether a side is on the boundary is decided by the
  // library's face matching, so interior members of a listed set must not be flagged: count how
  // often each face (sorted node tuple) occurs
  std::map<std::vector<int64_t>, int> face_count;
  static const int quad_sides[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
  static const int tet_sides[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};
  static const int hex_sides[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
  const int nv = m.nodesPerElem(), nps = etype == MCW_QUAD4 ? 2 : etype == MCW_TET4 ? 3 : 4;
  auto face = [&](int64_t e, int s) {
    std::vector<int64_t> key;
    for (int k = 0; k < nps; ++k)
    {
      const int local = etype == MCW_QUAD4 ? quad_sides[s][k] : etype == MCW_TET4 ? tet_sides[s][k] : hex_sides[s][k];
      key.push_back(m.conn[(size_t)e * nv + local]);
    }
    std::sort(key.begin(), key.end());
    return key;
  };
  for (int64_t e = 0; e < m.nElem(); ++e)
    for (int s = 0; s < ns; ++s)
      ++face_count[face(e, s)];
  int max_id = -1;
  std::map<int, int> kind_of;
  for (int id : reflect)
    kind_of[id] = MCW_BC_REFLECT;
  for (int id : kill)
    kind_of[id] = MCW_BC_KILL;
  for (const auto& kv : kind_of)
    max_id = std::max(max_id, kv.first);
  if (f.has("ss_prop1"))
  {
    const auto ss_ids = f.ints("ss_prop1");
    for (size_t k = 0; k < ss_ids.size(); ++k)
    {
      if (!kind_of.count((int)ss_ids[k]))
        continue;
      const auto elems = f.ints("elem_ss" + std::to_string(k + 1)), sides = f.ints("side_ss" + std::to_string(k + 1));
      for (size_t q = 0; q < elems.size(); ++q)
      {
        const int64_t e = elems[q] - 1;
        const int s = (int)sides[q] - 1;
        if (e < 0 || e >= m.nElem() || s < 0 || s >= ns)
          mooseError("side set entry outside the mesh");
        if (face_count[face(e, s)] == 1)
          m.side_bnd_id[(size_t)e * ns + s] = (int32_t)ss_ids[k];
      }
    }
  }
  m.bnd_kind.assign((size_t)(max_id + 1), MCW_BC_KILL);
  for (const auto& kv : kind_of)
    m.bnd_kind[kv.first] = kv.second;
  return m;
}

} // namespace macaw
