/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for stk::mesh::NgpMesh. */
#ifndef NW_REF_SHIM_STK_NGPMESH_HPP
#define NW_REF_SHIM_STK_NGPMESH_HPP
#include "Types.hpp"
namespace stk {
namespace mesh {
class NgpMesh
{
public:
  FastMeshIndex fast_mesh_index(const Entity& e) const
  {
    return FastMeshIndex{0u, (unsigned)e.m_value};
  }
};
} // namespace mesh
} // namespace stk
#endif
