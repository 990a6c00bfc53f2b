/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for the STK mesh types the
 * reference's edge algorithms name: a flat one-bucket mesh view (entity index
 * == bucket ordinal), node / edge fields as plain arrays.  Written for this
 * repo; not STK code. */
#ifndef NW_REF_SHIM_STK_TYPES_HPP
#define NW_REF_SHIM_STK_TYPES_HPP
#include <cstdint>
#include <string>
#include <vector>
#include <stk_topology/topology.hpp>
#include "Entity.hpp"
namespace stk {
namespace mesh {
typedef stk::topology::rank_t EntityRank;
typedef uint64_t EntityId;
constexpr unsigned InvalidOrdinal = ~0u;
enum FieldState {
  StateNone = 0,
  StateNew = 0,
  StateNP1 = 0,
  StateOld = 1,
  StateN = 1,
  StateNM1 = 2
};
struct FastMeshIndex
{
  unsigned bucket_id;
  unsigned bucket_ord;
};
class BulkData; /* defined with the stand-in mesh, nalu/RefHarness.h */
class MetaData;
class Part
{
};
typedef std::vector<Part*> PartVector;
class Selector
{
public:
  Selector() {}
  Selector(const Part&) {}
  Selector operator&(const Selector&) const { return *this; }
  Selector operator|(const Selector&) const { return *this; }
  Selector operator!() const { return *this; }
};
inline Selector operator&(const Part&, const Selector& s) { return s; }
inline Selector selectUnion(const PartVector&) { return Selector(); }
} // namespace mesh
} // namespace stk
#endif
