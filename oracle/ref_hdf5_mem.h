// ref_hdf5_mem.h -- an in-memory tree of groups / datasets / attributes that stands where an HDF5 file would, so the
// reference's own data-object constructors (Nuclide, Reaction, ReactionProduct, the angle-energy distributions, ...) can be
// run by oracle/_ref without the HDF5 library: ref_hdf5_mem.cpp implements the functions of the reference's
// include/openmc/hdf5_interface.h (which its hdf5_interface.cpp would implement over libhdf5) on top of this tree.
// Object ids (hid_t) are Node pointers.  Test infrastructure only.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace memh5
{
struct Array
{
  enum Kind { REAL, INT, TEXT } kind = REAL;
  std::vector<unsigned long long> shape; // empty: scalar
  std::vector<double> real;
  std::vector<long long> integer;
  std::string text;
};

struct Node
{
  std::string name;
  bool is_dataset = false;
  std::map<std::string, std::unique_ptr<Node>> children; // name order, like HDF5's default link index
  std::map<std::string, Array> attrs;
  Array data;

  Node* group(const std::string& name)
  {
    auto& c = children[name];
    if (!c)
    {
      c.reset(new Node);
      c->name = name;
    }
    return c.get();
  }
  Node* dataset(const std::string& name, const std::vector<double>& v, std::vector<unsigned long long> shape = {})
  {
    Node* n = group(name);
    n->is_dataset = true;
    n->data.kind = Array::REAL;
    n->data.real = v;
    n->data.shape = shape.empty() ? std::vector<unsigned long long>{v.size()} : shape;
    return n;
  }
  void attr(const std::string& name, double v) { attrs[name] = Array{Array::REAL, {}, {v}, {}, {}}; }
  void attr(const std::string& name, int v) { attrs[name] = Array{Array::INT, {}, {}, {v}, {}}; }
  void attr(const std::string& name, const std::vector<int>& v)
  {
    Array a{Array::INT, {v.size()}, {}, {}, {}};
    a.integer.assign(v.begin(), v.end());
    attrs[name] = a;
  }
  void attr(const std::string& name, const std::vector<int>& v, std::vector<unsigned long long> shape)
  {
    Array a{Array::INT, shape, {}, {}, {}};
    a.integer.assign(v.begin(), v.end());
    attrs[name] = a;
  }
  void attr(const std::string& name, const char* v) { attrs[name] = Array{Array::TEXT, {}, {}, {}, v}; }
};

inline int64_t id_of(Node* n) { return (int64_t)(intptr_t)n; }
inline Node* node_of(int64_t id) { return (Node*)(intptr_t)id; }
} // namespace memh5
