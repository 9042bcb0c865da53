// ref_hdf5_mem.cpp -- the reference's include/openmc/hdf5_interface.h implemented over the in-memory tree of
// ref_hdf5_mem.h instead of libhdf5 (its own hdf5_interface.cpp cannot be built here: no HDF5 library).  Only the read side
// exists; the writers abort (no ref_* doorway reaches a to_hdf5 path).  Test infrastructure only; nothing here restates a
// transport algorithm -- it only hands arrays to the reference's own constructors.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "openmc/hdf5_interface.h"
#include "ref_hdf5_mem.h"

using memh5::Array;
using memh5::Node;
using memh5::node_of;

namespace openmc
{
template<> const hid_t H5TypeMap<bool>::type_id = H5T_NATIVE_INT8;
template<> const hid_t H5TypeMap<int>::type_id = H5T_NATIVE_INT;
template<> const hid_t H5TypeMap<unsigned long>::type_id = H5T_NATIVE_ULONG;
template<> const hid_t H5TypeMap<unsigned long long>::type_id = H5T_NATIVE_UINT64;
template<> const hid_t H5TypeMap<unsigned int>::type_id = H5T_NATIVE_UINT;
template<> const hid_t H5TypeMap<int64_t>::type_id = H5T_NATIVE_INT64;
template<> const hid_t H5TypeMap<double>::type_id = H5T_NATIVE_DOUBLE;
template<> const hid_t H5TypeMap<char>::type_id = 10;

static void die(const std::string& what)
{
  std::fprintf(stderr, "oracle/_ref in-memory hdf5: %s\n", what.c_str());
  std::abort();
}

static const Array& attr_of(hid_t obj, const char* name)
{
  Node* n = node_of(obj);
  auto it = n->attrs.find(name);
  if (it == n->attrs.end())
    die(std::string("no attribute ") + name);
  return it->second;
}
static Node* child_of(hid_t obj, const char* name)
{
  // names may be paths ("urr/294K")
  Node* n = node_of(obj);
  std::string rest = name;
  while (!rest.empty())
  {
    const size_t slash = rest.find('/');
    const std::string head = rest.substr(0, slash);
    rest = slash == std::string::npos ? "" : rest.substr(slash + 1);
    if (head.empty())
      continue;
    auto it = n->children.find(head);
    if (it == n->children.end())
      return nullptr;
    n = it->second.get();
  }
  return n;
}
static void copy_out(const Array& a, hid_t mem_type, void* buffer)
{
  const size_t n = a.kind == Array::REAL ? a.real.size() : a.integer.size();
  for (size_t k = 0; k < n; ++k)
  {
    const double r = a.kind == Array::REAL ? a.real[k] : (double)a.integer[k];
    const long long i = a.kind == Array::REAL ? (long long)a.real[k] : a.integer[k];
    if (mem_type == H5T_NATIVE_DOUBLE)
      ((double*)buffer)[k] = r;
    else if (mem_type == H5T_NATIVE_INT)
      ((int*)buffer)[k] = (int)i;
    else if (mem_type == H5T_NATIVE_UINT)
      ((unsigned*)buffer)[k] = (unsigned)i;
    else if (mem_type == H5T_NATIVE_INT64 || mem_type == H5T_NATIVE_LONG)
      ((int64_t*)buffer)[k] = (int64_t)i;
    else if (mem_type == H5T_NATIVE_UINT64 || mem_type == H5T_NATIVE_ULONG)
      ((uint64_t*)buffer)[k] = (uint64_t)i;
    else if (mem_type == H5T_NATIVE_INT8)
      ((int8_t*)buffer)[k] = (int8_t)i;
    else
      die("unsupported memory type " + std::to_string((long long)mem_type));
  }
}

void read_attr(hid_t obj_id, const char* name, hid_t mem_type_id, void* buffer)
{
  copy_out(attr_of(obj_id, name), mem_type_id, buffer);
}
void read_dataset_lowlevel(hid_t obj_id, const char* name, hid_t mem_type_id, hid_t, bool, void* buffer)
{
  Node* n = name ? child_of(obj_id, name) : node_of(obj_id);
  if (!n || !n->is_dataset)
    die(std::string("no dataset ") + (name ? name : "(self)"));
  copy_out(n->data, mem_type_id, buffer);
}
vector<hsize_t> attribute_shape(hid_t obj_id, const char* name)
{
  const Array& a = attr_of(obj_id, name);
  return vector<hsize_t>(a.shape.begin(), a.shape.end());
}
vector<hsize_t> object_shape(hid_t obj_id)
{
  const Array& a = node_of(obj_id)->data;
  return vector<hsize_t>(a.shape.begin(), a.shape.end());
}
std::string object_name(hid_t obj_id) { return "/" + node_of(obj_id)->name; }
static vector<std::string> names(hid_t group_id, bool datasets)
{
  vector<std::string> out;
  for (const auto& kv : node_of(group_id)->children)
    if (kv.second->is_dataset == datasets)
      out.push_back(kv.first);
  return out;
}
vector<std::string> dataset_names(hid_t group_id) { return names(group_id, true); }
vector<std::string> group_names(hid_t group_id) { return names(group_id, false); }
void write_dataset_lowlevel(hid_t, int, const hsize_t*, const char*, hid_t, hid_t, bool, const void*)
{
  die("write_dataset_lowlevel: the tree is read-only");
}
void write_string(hid_t, const char*, const std::string&, bool) { die("write_string: the tree is read-only"); }
void write_attr(hid_t, int, const hsize_t*, const char*, hid_t, const void*) { die("write_attr: the tree is read-only"); }
hid_t create_group(hid_t, const std::string& name)
{
  die("create_group(" + name + "): the tree is read-only");
  return -1;
}
hid_t file_open(const std::string& filename, char, bool)
{
  die("file_open(" + filename + "): there are no files here");
  return -1;
}

extern "C" {
bool attribute_exists(hid_t obj_id, const char* name) { return node_of(obj_id)->attrs.count(name) != 0; }
size_t attribute_typesize(hid_t obj_id, const char* name) { return attr_of(obj_id, name).text.size(); }
void read_attr_string(hid_t obj_id, const char* name, size_t slen, char* buffer)
{
  const Array& a = attr_of(obj_id, name);
  std::memcpy(buffer, a.text.data(), std::min(slen, a.text.size()));
}
bool object_exists(hid_t object_id, const char* name) { return child_of(object_id, name) != nullptr; }
hid_t open_group(hid_t group_id, const char* name)
{
  Node* n = child_of(group_id, name);
  if (!n || n->is_dataset)
    die(std::string("no group ") + name);
  return memh5::id_of(n);
}
hid_t open_dataset(hid_t group_id, const char* name)
{
  Node* n = child_of(group_id, name);
  if (!n || !n->is_dataset)
    die(std::string("no dataset ") + name);
  return memh5::id_of(n);
}
void write_attr_string(hid_t, const char*, const char*) { die("write_attr_string: the tree is read-only"); }
void file_close(hid_t) {}
void close_group(hid_t) {}
void close_dataset(hid_t) {}
}
} // namespace openmc
