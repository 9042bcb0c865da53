// ref_gather.cpp -- oracle/_ref doorway for host/openmc_nuclide_reader.hpp (the MaCaw-side gatherLibrary() of
// INTEGRATION.md): reads the nuclide-file image the reference's own constructors just read (ref_load_library) through
// OpenMC's hdf5_interface.h, in the order of openmc::data::nuclides, and returns the library's digest (mcw_digest, passed
// in as a function pointer so that this shim does not link the CUDA library) of the gathered nuclides, further
// temperature sets and probability tables.  Test infrastructure.
#include "../host/openmc_nuclide_reader.hpp"

#include "openmc/nuclide.h"
#include "openmc/thermal.h"

#include "ref_hdf5_mem.h"

memh5::Node* ref_image_root(); // ref_shim_physics.cpp

extern "C" __attribute__((visibility("default"))) int ref_gather_digest(void* digest_fn, uint64_t* out3, char* message,
                                                                         int message_bytes, uint64_t* thermal_out,
                                                                         int32_t* thermal_same, double* thermal_kT, int thermal_cap)
{
  using digest_t = int (*)(int, const void*, int, uint64_t*);
  digest_t digest = (digest_t)digest_fn;
  try
  {
    memh5::Node* root = ref_image_root();
    std::vector<std::pair<std::string, hid_t>> groups;
    for (const auto& nuc : openmc::data::nuclides)
      groups.push_back({nuc->name_, memh5::id_of(root->children.at(nuc->name_).get())});
    const macaw::Library lib = macaw::omcfile::readLibrary(groups);
    const auto nd = lib.nuclideDescs();
    const auto sd = lib.temperatureSetDescs();
    std::vector<mcw_urr_desc> ud;
    for (const auto& u : lib.urr)
      ud.push_back(u.desc());
    digest(0, nd.data(), (int)nd.size(), &out3[0]);
    digest(0, sd.data(), (int)sd.size(), &out3[1]);
    digest(2, ud.data(), (int)ud.size(), &out3[2]);
    if (thermal_out)
    {
      // the thermal-scattering tables, in the order of data::thermal_scatt: digest of every temperature entry, and which
      // table / kT each entry is
      macaw::Library th;
      std::vector<hid_t> tg;
      for (const auto& t : openmc::data::thermal_scatt)
        tg.push_back(memh5::id_of(root->children.at(t->name_).get()));
      macaw::omcfile::readThermalTables(tg, th);
      std::vector<mcw_thermal_desc> td;
      for (const auto& t : th.thermal)
        td.push_back(t.desc());
      digest(1, td.data(), (int)td.size(), thermal_out);
      for (size_t k = 0; k < th.thermal_same.size() && (int)k < thermal_cap; ++k)
        thermal_same[k] = th.thermal_same[k], thermal_kT[k] = th.thermal_kT[k];
    }
    return 0;
  }
  catch (const std::exception& e)
  {
    std::snprintf(message, (size_t)message_bytes, "%s", e.what());
    return 1;
  }
}
