/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/FieldBase.hpp>:
 * the type names FieldTypeDef.h aliases; never instantiated here. */
#ifndef NW_REF_SHIM_STK_FIELDBASE_HPP
#define NW_REF_SHIM_STK_FIELDBASE_HPP
#include <cstdint>
#include "Entity.hpp"
namespace stk {
namespace mesh {
typedef uint64_t EntityId;
class FieldBase {};
template <class T>
class Field : public FieldBase {};
} // namespace mesh
} // namespace stk
#endif
