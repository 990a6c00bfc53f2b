/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/FieldParallel.hpp>: the
 * few functions the reference's edge algorithm sources name are declared with
 * the stand-in mesh in nalu/RefHarness.h. */
#ifndef NW_REF_SHIM_STK_FIELDPARALLEL_HPP
#define NW_REF_SHIM_STK_FIELDPARALLEL_HPP
#include "Types.hpp"
#endif
