/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for stk::mesh::NgpMesh. */
#ifndef NW_REF_SHIM_STK_NGPMESH_HPP
#define NW_REF_SHIM_STK_NGPMESH_HPP
#include "Types.hpp"
namespace stk {
namespace mesh {
/* the nodes of an entity as the loop shells hand them to a CoeffApplier */
struct ConnectedNodesShim
{
  const Entity* p = nullptr;
  unsigned n = 0;
  ConnectedNodesShim() {}
  ConnectedNodesShim(const Entity* p_, unsigned n_) : p(p_), n(n_) {}
  const Entity& operator[](unsigned i) const { return p[i]; }
  unsigned size() const { return n; }
};
class NgpMesh
{
public:
  using ConnectedNodes = ConnectedNodesShim;
  using MeshIndex = FastMeshIndex;
  Entity get_entity(stk::topology::rank_t, const FastMeshIndex& i) const
  {
    Entity e;
    e.m_value = i.bucket_ord;
    return e;
  }
  FastMeshIndex fast_mesh_index(const Entity& e) const
  {
    return FastMeshIndex{0u, (unsigned)e.m_value};
  }
};
} // namespace mesh
} // namespace stk
#endif
