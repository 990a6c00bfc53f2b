/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for stk::mesh::NgpField: a
 * view of a plain [entity][component] array. */
#ifndef NW_REF_SHIM_STK_NGPFIELD_HPP
#define NW_REF_SHIM_STK_NGPFIELD_HPP
#include "Types.hpp"
namespace stk {
namespace mesh {
template <class T>
class NgpField
{
public:
  using value_type = T;
  NgpField() {}
  stk::topology::rank_t get_rank() const { return stk::topology::NODE_RANK; }
  NgpField(T* d, int nc) : data_(d), ncomp_(nc) {}
  T& get(const FastMeshIndex& i, int c) const
  {
    return data_[(size_t)i.bucket_ord * ncomp_ + c];
  }
  T& operator()(const FastMeshIndex& i, int c) const { return get(i, c); }
  template <class Mesh>
  T& get(const Mesh&, const Entity& e, int c) const
  {
    return data_[(size_t)e.m_value * ncomp_ + c];
  }
  void sync_to_device() const {}
  void sync_to_host() const {}
  void modify_on_device() const {}
  void modify_on_host() const {}
  void clear_sync_state() const {}
  T* data_ = nullptr;
  int ncomp_ = 0;
};
} // namespace mesh
} // namespace stk
#endif
