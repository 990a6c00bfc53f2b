/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/Ngp.hpp>. */
#ifndef NW_REF_SHIM_STK_NGP_HPP
#define NW_REF_SHIM_STK_NGP_HPP
#include "NgpMesh.hpp"
#include "NgpField.hpp"
#endif
