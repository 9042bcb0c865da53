// mcw_kernels.h -- host-callable launchers of the kernels in mcw_kernels.cu
//
// mcw_kernels.cu is compiled TWICE: -DMCW_RARE=0 (namespace mcw_r0) holds the path every benchmark and most problems
// take; -DMCW_RARE=1 (namespace mcw_r1) is the same code with the rarely used branches compiled in -- thermal
// scattering S(a,b), tallies with nuclide bins, cached trace points, tabulated secondary distributions.  Those
// branches cost registers, spills and instruction-cache footprint in the tracking and collision kernels even when they
// are never taken (out-of-line calls, longer live ranges), so they are kept out of the common build altogether and the
// library picks the set per problem (mcw_context::rare, decided in prepare()).  Results of the two builds are
// bit-identical wherever both apply (tests/test_gpu_parity.py runs problems through both).
#pragma once
#include <cuda_runtime.h>

#include "mcw_types.h"

namespace mcw_r0
{
#include "mcw_kernels_decl.inc"
}
namespace mcw_r1
{
#include "mcw_kernels_decl.inc"
}
