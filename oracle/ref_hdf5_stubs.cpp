// ref_hdf5_stubs.cpp -- what the reference's hdf5_interface.cpp would define, for the objects that are linked into
// oracle/_ref/libomc_ref.so without it (that source needs the real HDF5 library; fake_hdf5/hdf5.h only lets its
// *callers* parse).  The type-id constants are data symbols that templates in hdf5_interface.h refer to; the two
// writers are reachable only from to_hdf5 / write-results paths, which no ref_* doorway calls: they abort.
// Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "openmc/hdf5_interface.h"

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

static void no_hdf5(const char* what)
{
  std::fprintf(stderr, "oracle/_ref: %s needs the HDF5 library, which this build stubs out\n", what);
  std::abort();
}
void write_dataset_lowlevel(hid_t, int, const hsize_t*, const char*, hid_t, hid_t, bool, const void*)
{
  no_hdf5("write_dataset_lowlevel");
}
void write_string(hid_t, const char*, const std::string&, bool) { no_hdf5("write_string"); }
} // namespace openmc
