/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/Entity.hpp>. */
#ifndef NW_REF_SHIM_STK_ENTITY_HPP
#define NW_REF_SHIM_STK_ENTITY_HPP
#include <cstdint>
namespace stk {
namespace mesh {
struct Entity
{
  uint64_t m_value = 0;
  uint64_t local_offset() const { return m_value; }
};
} // namespace mesh
} // namespace stk
#endif
