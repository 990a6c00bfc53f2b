/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/NgpField.hpp>. */
#ifndef NW_REF_SHIM_STK_NGPFIELD_HPP
#define NW_REF_SHIM_STK_NGPFIELD_HPP
namespace stk {
namespace mesh {
template <class T>
class NgpField {};
} // namespace mesh
} // namespace stk
#endif
