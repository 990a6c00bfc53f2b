/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_mesh/base/FieldBase.hpp>:
 * a named field handle (ordinal into the harness' field table). */
#ifndef NW_REF_SHIM_STK_FIELDBASE_HPP
#define NW_REF_SHIM_STK_FIELDBASE_HPP
#include <string>
#include "Types.hpp"
namespace stk {
namespace mesh {
class FieldBase
{
public:
  FieldBase() {}
  FieldBase(const std::string& n, unsigned ord, int ncomp)
    : name_(n), ordinal_(ord), ncomp_(ncomp)
  {
  }
  const std::string& name() const { return name_; }
  unsigned mesh_meta_data_ordinal() const { return ordinal_; }
  int max_size() const { return ncomp_; }
  /* states: "F" is NP1, "F_n" is N, "F_nm1" is NM1 where registered (the
   * look-up lives with the field table, nalu/RefHarness.h) */
  unsigned number_of_states() const;
  FieldBase& field_of_state(FieldState s) const;
  std::string name_;
  unsigned ordinal_ = InvalidOrdinal;
  int ncomp_ = 0;
};
template <class T>
class Field : public FieldBase
{
public:
  using FieldBase::FieldBase;
};
} // namespace mesh
} // namespace stk
#endif
