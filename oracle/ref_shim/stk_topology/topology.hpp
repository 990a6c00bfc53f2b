/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_topology/topology.hpp>:
 * the enumerators the reference's AlgTraits.h and master elements name (values
 * arbitrary) and the side counts / side topologies / node counts of the
 * linear element topologies, which the master-element constructors query. */
#ifndef NW_REF_SHIM_STK_TOPOLOGY_HPP
#define NW_REF_SHIM_STK_TOPOLOGY_HPP
namespace stk {
struct topology
{
  enum topology_t {
    INVALID_TOPOLOGY,
    NODE,
    LINE_2,
    LINE_3,
    BEAM_2,
    TRI_3,
    TRI_3_2D,
    TRI_6_2D,
    QUAD_4,
    QUAD_4_2D,
    QUAD_9,
    QUAD_9_2D,
    QUADRILATERAL_4,
    QUADRILATERAL_4_2D,
    TRIANGLE_3_2D,
    QUADRILATERAL_9,
    SHELL_TRI_3,
    SHELL_QUAD_4,
    TET_4,
    TET_10,
    PYRAMID_5,
    WEDGE_6,
    HEX_8,
    HEX_27,
    SUPEREDGE_START = 100,
    SUPERFACE_START = 200,
    SUPERELEMENT_START = 300
  };
  enum rank_t {
    BEGIN_RANK = 0,
    NODE_RANK = 0,
    EDGE_RANK = 1,
    FACE_RANK = 2,
    ELEM_RANK = 3,
    ELEMENT_RANK = 3,
    CONSTRAINT_RANK = 4,
    END_RANK = 5,
    INVALID_RANK = 256
  };
  topology_t m_value;
  constexpr topology() : m_value(INVALID_TOPOLOGY) {}
  constexpr topology(topology_t t) : m_value(t) {}
  constexpr operator topology_t() const { return m_value; }
  constexpr topology_t value() const { return m_value; }
  unsigned num_nodes() const
  {
    switch (m_value) {
    case NODE: return 1;
    case LINE_2: case BEAM_2: return 2;
    case LINE_3: case TRI_3: case TRI_3_2D: case TRIANGLE_3_2D: case SHELL_TRI_3: return 3;
    case QUAD_4: case QUAD_4_2D: case QUADRILATERAL_4: case QUADRILATERAL_4_2D:
    case SHELL_QUAD_4: case TET_4: return 4;
    case PYRAMID_5: return 5;
    case WEDGE_6: case TRI_6_2D: return 6;
    case HEX_8: return 8;
    case QUAD_9: case QUAD_9_2D: case QUADRILATERAL_9: return 9;
    case TET_10: return 10;
    case HEX_27: return 27;
    default: return 0;
    }
  }
  unsigned num_sides() const
  {
    switch (m_value) {
    case TRI_3_2D: case TRIANGLE_3_2D: case TRI_6_2D: return 3;
    case QUAD_4_2D: case QUADRILATERAL_4_2D: case QUAD_9_2D: case TET_4: case TET_10: return 4;
    case PYRAMID_5: case WEDGE_6: return 5;
    case HEX_8: case HEX_27: return 6;
    default: return 0;
    }
  }
  topology side_topology(unsigned k) const
  {
    switch (m_value) {
    case TRI_3_2D: case TRIANGLE_3_2D: case QUAD_4_2D: case QUADRILATERAL_4_2D: return LINE_2;
    case TRI_6_2D: case QUAD_9_2D: return LINE_3;
    case TET_4: return TRI_3;
    case PYRAMID_5: return k < 4 ? TRI_3 : QUAD_4;
    case WEDGE_6: return k < 3 ? QUAD_4 : TRI_3;
    case HEX_8: return QUAD_4;
    case HEX_27: return QUAD_9;
    default: return INVALID_TOPOLOGY;
    }
  }
};
} // namespace stk
#endif
