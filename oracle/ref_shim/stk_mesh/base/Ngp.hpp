/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/Ngp.hpp>. */
#ifndef NW_REF_SHIM_STK_NGP_HPP
#define NW_REF_SHIM_STK_NGP_HPP
namespace stk {
namespace mesh {
class NgpMesh {};
} // namespace mesh
} // namespace stk
#endif
