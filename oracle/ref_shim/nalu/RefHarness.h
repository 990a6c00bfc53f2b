/*
 * oracle/ref_shim/nalu/RefHarness.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-ins for the parts of nalu-wind that surround the edge algorithms, so
 * that the reference's OWN source files
 *     src/edge_kernels/{Momentum,Scalar,Continuity}EdgeSolverAlg.C
 *     src/edge_kernels/{WallDistEdgeSolverAlg,MomentumEdgePecletAlg}.C
 *     src/ngp_algorithms/{MdotEdgeAlg,NodalGradEdgeAlg}.C
 * compile unmodified, from where they lie, and their constructors and
 * execute() bodies -- the per-edge lambdas SURVEY.md 8(a) rows a1, a2, a4, a5, a6 and
 * 8(f) rows 1, 3 cite -- run here on arrays handed in from Python.
 *
 * What is the reference's: every line of those four .C files and of the
 * headers they own (edge_kernels/*.h, ngp_algorithms/MdotEdgeAlg.h,
 * PecletFunction.h/.C, EdgeKernelUtils.h, Enums.h, FieldTypeDef.h,
 * KokkosInterface.h, SimdInterface.h).
 * What is a stand-in (this file and its one-line forwarding headers, which
 * shadow the reference's headers of the same name on the include path):
 * Realm, SolutionOptions, EquationSystem, Algorithm, the loop shell
 * AssembleEdgeSolverAlgorithm::run_algorithm / nalu_ngp::run_edge_algorithm
 * (here: a serial loop over the edge list that zeroes the local block, calls
 * the reference's lambda and records the block instead of scattering it), the
 * field manager and get_field_ordinal.  Written for this repo.
 */
#ifndef NW_REF_HARNESS_H
#define NW_REF_HARNESS_H

#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mpi.h>
#include <stdexcept>
#include <string>
#include <vector>

#include <KokkosInterface.h>
#include <SimdInterface.h>
#include <Enums.h>
#include <FieldTypeDef.h>
#include <PecletFunction.h>
#include <stk_mesh/base/Types.hpp>
#include <stk_mesh/base/NgpMesh.hpp>
#include <stk_mesh/base/NgpField.hpp>
#include <stk_util/util/ReportHandler.hpp>

namespace nwref {

struct FieldRec
{
  std::string name;
  int rank; /* stk::topology::NODE_RANK / EDGE_RANK */
  int ncomp;
  double* data;
  int* idata = nullptr; /* integer fields (hypre_global_id, nalu_global_id) */
};

/* everything the stand-in Realm answers from */
struct World
{
  int ndim = 3;
  long nNodes = 0, nEdges = 0;
  const int* edgeNodes = nullptr; /* [nEdges][2] */
  std::vector<FieldRec> fields;
  std::vector<stk::mesh::FieldBase*> fieldHandles;
  std::map<std::string, double> opt; /* "alpha:velocity", "dt", ... */
  double gravity[3] = {0, 0, 0};
  int pecletForm = 0; /* 0 classic(hybridFactor = pecletA), 1 tanh(c1, c2) */
  double pecletA = 1.0, pecletB = 1.0;
  /* parallel decomposition as seen by this "rank" (one process plays one
   * rank at a time): STK identifier, owner and periodic master id per local
   * node; the hypre row range of this rank */
  int rank = 0, nranks = 1;
  const long* nodeIdentifier = nullptr; /* [nNodes], STK global ids (1-based) */
  const int* nodeOwner = nullptr;       /* [nNodes] */
  std::map<long, long> nodeOfIdentifier;
  long hypreILower = 0, hypreIUpper = 0, hypreNumNodes = 0;
  std::vector<int> hypreOffsets;
  /* node selector of run_entity_algorithm (empty: every node) */
  std::vector<char> nodeSelected;
  /* where run_algorithm records the local blocks */
  double* lhsOut = nullptr;
  double* rhsOut = nullptr;
  /* ... or, when set, hands them on (to the reference's own CoeffApplier,
   * oracle/ref_hypre_driver.cpp: ref_hypre_sweep) as the reference's shell does */
  std::function<void(long, const double*, const double*, int)> applyHook;

  static World& self()
  {
    static World w;
    return w;
  }
  double get(const std::string& key) const
  {
    auto it = opt.find(key);
    if (it == opt.end())
      throw std::runtime_error("ref harness: option not set: " + key);
    return it->second;
  }
  double get(const std::string& key, double dflt) const
  {
    auto it = opt.find(key);
    return it == opt.end() ? dflt : it->second;
  }
  bool has(const std::string& name, int rank) const
  {
    for (const auto& f : fields)
      if (f.name == name && f.rank == rank)
        return true;
    return false;
  }
  static std::string state_name(const std::string& name, int state)
  {
    return state == 0 ? name : name + (state == 1 ? "_n" : "_nm1");
  }
  unsigned ordinal(const std::string& name, int rank) const
  {
    for (size_t i = 0; i < fields.size(); ++i)
      if (fields[i].name == name && fields[i].rank == rank)
        return (unsigned)i;
    throw std::runtime_error("ref harness: no such field: " + name);
  }
};

} // namespace nwref

namespace stk {
namespace mesh {
class MetaData
{
public:
  unsigned spatial_dimension() const { return nwref::World::self().ndim; }
  Part& locally_owned_part() const
  {
    static Part p;
    return p;
  }
  Part& universal_part() const { return locally_owned_part(); }
  Part& globally_shared_part() const { return locally_owned_part(); }
  EntityRank side_rank() const { return stk::topology::FACE_RANK; }
  const std::vector<FieldBase*>& get_fields() const
  {
    return nwref::World::self().fieldHandles;
  }
  template <class T>
  Field<T>* get_field(EntityRank rank, const std::string& name) const
  {
    const auto& w = nwref::World::self();
    return static_cast<Field<T>*>(w.fieldHandles.at(w.ordinal(name, rank)));
  }
};
/* one bucket holding every entity of a rank: all nodes, or the locally-owned
 * edges (the edge list handed in is the owned one) */
class Bucket
{
public:
  typedef size_t size_type;
  Bucket(EntityRank r, size_t n) : rank_(r), n_(n) {}
  size_t size() const { return n_; }
  Entity operator[](size_t k) const
  {
    Entity e;
    e.m_value = k;
    return e;
  }
  stk::topology topology() const
  {
    return rank_ == stk::topology::EDGE_RANK ? stk::topology::LINE_2
                                             : stk::topology::NODE;
  }
  const Entity* begin_nodes(size_t k) const
  {
    const auto& w = nwref::World::self();
    if (rank_ == stk::topology::EDGE_RANK) {
      nodes_[0].m_value = (uint64_t)w.edgeNodes[2 * k];
      nodes_[1].m_value = (uint64_t)w.edgeNodes[2 * k + 1];
    } else {
      nodes_[0].m_value = k;
    }
    return nodes_;
  }
  unsigned num_nodes(size_t) const { return rank_ == stk::topology::EDGE_RANK ? 2 : 1; }
  EntityRank rank_;
  size_t n_;
  mutable Entity nodes_[2];
};
typedef std::vector<const Bucket*> BucketVector;
class BulkData
{
public:
  const MetaData& mesh_meta_data() const
  {
    static MetaData m;
    return m;
  }
  MPI_Comm parallel() const { return 0; }
  int parallel_size() const { return nwref::World::self().nranks; }
  int parallel_rank() const { return nwref::World::self().rank; }
  /* nodes and edges only: there are no elements or faces in this mesh view */
  const BucketVector& get_buckets(EntityRank rank, const Selector&) const
  {
    const auto& w = nwref::World::self();
    vec_[rank].clear();
    if (rank == stk::topology::EDGE_RANK || rank == stk::topology::NODE_RANK) {
      bucket_[rank].reset(new Bucket(
        rank, rank == stk::topology::EDGE_RANK ? w.nEdges : w.nNodes));
      vec_[rank].push_back(bucket_[rank].get());
    }
    return vec_[rank];
  }
  const Entity* begin_nodes(Entity edge) const
  {
    const auto& w = nwref::World::self();
    nodes_[0].m_value = (uint64_t)w.edgeNodes[2 * edge.m_value];
    nodes_[1].m_value = (uint64_t)w.edgeNodes[2 * edge.m_value + 1];
    return nodes_;
  }
  unsigned num_nodes(Entity) const { return 2; }
  unsigned num_elements(Entity) const { return 0; }
  const Entity* begin_elements(Entity) const { return nullptr; }
  EntityId identifier(Entity node) const
  {
    const auto& w = nwref::World::self();
    return w.nodeIdentifier ? (EntityId)w.nodeIdentifier[node.m_value]
                            : (EntityId)node.m_value + 1;
  }
  Entity get_entity(EntityRank, EntityId id) const
  {
    const auto& w = nwref::World::self();
    Entity e;
    if (!w.nodeIdentifier) {
      e.m_value = id - 1;
      return e;
    }
    auto it = w.nodeOfIdentifier.find((long)id);
    e.m_value = it == w.nodeOfIdentifier.end() ? ~uint64_t(0) : (uint64_t)it->second;
    return e;
  }
  bool is_valid(Entity e) const { return e.m_value != ~uint64_t(0); }
  int parallel_owner_rank(Entity node) const
  {
    const auto& w = nwref::World::self();
    return w.nodeOwner ? w.nodeOwner[node.m_value] : w.rank;
  }

private:
  mutable std::unique_ptr<Bucket> bucket_[6];
  mutable BucketVector vec_[6];
  mutable Entity nodes_[2];
};
class Ghosting
{
};
inline void
communicate_field_data(const Ghosting&, const std::vector<const FieldBase*>&)
{
}
inline Selector selectField(const FieldBase&) { return Selector(); }
inline double*
field_data(const FieldBase& f, Entity e)
{
  const auto& r = nwref::World::self().fields.at(f.mesh_meta_data_ordinal());
  return r.data + (size_t)e.m_value * r.ncomp;
}
inline double*
field_data(const Field<double>& f, Entity e)
{
  return field_data(static_cast<const FieldBase&>(f), e);
}
inline int*
field_data(const Field<int>& f, Entity e)
{
  const auto& r = nwref::World::self().fields.at(f.mesh_meta_data_ordinal());
  return r.idata + (size_t)e.m_value * r.ncomp;
}
/* nalu_global_id is a Field<EntityId>; kept as int here */
inline EntityId*
field_data(const Field<EntityId>& f, Entity e)
{
  static thread_local EntityId v;
  const auto& r = nwref::World::self().fields.at(f.mesh_meta_data_ordinal());
  v = (EntityId)r.idata[(size_t)e.m_value * r.ncomp];
  return &v;
}
inline void
field_fill(double v, const FieldBase& f)
{
  const auto& w = nwref::World::self();
  const auto& r = w.fields.at(f.mesh_meta_data_ordinal());
  const size_t n =
    (size_t)(r.rank == stk::topology::EDGE_RANK ? w.nEdges : w.nNodes) * r.ncomp;
  for (size_t i = 0; i < n; ++i)
    r.data[i] = v;
}
inline void
copy_owned_to_shared(const BulkData&, const std::vector<const FieldBase*>&)
{
}
} // namespace mesh
} // namespace stk

namespace stk {
namespace mesh {
inline unsigned
FieldBase::number_of_states() const
{
  const auto& w = nwref::World::self();
  const int rank = w.fields.at(ordinal_).rank;
  return 1u + (w.has(name_ + "_n", rank) ? 1u : 0u) +
         (w.has(name_ + "_nm1", rank) ? 1u : 0u);
}
inline FieldBase&
FieldBase::field_of_state(FieldState s) const
{
  const auto& w = nwref::World::self();
  const int rank = w.fields.at(ordinal_).rank;
  return *w.fieldHandles.at(w.ordinal(nwref::World::state_name(name_, s), rank));
}
} // namespace mesh
} // namespace stk

namespace sierra {
namespace nalu {

/* utils/StkHelpers.h */
inline unsigned
get_field_ordinal(
  const stk::mesh::MetaData&, const std::string& name,
  const stk::mesh::EntityRank rank = stk::topology::NODE_RANK)
{
  return nwref::World::self().ordinal(name, rank);
}
inline unsigned
get_field_ordinal(
  const stk::mesh::MetaData&, const std::string& name,
  const stk::mesh::FieldState state,
  const stk::mesh::EntityRank rank = stk::topology::NODE_RANK)
{
  return nwref::World::self().ordinal(nwref::World::state_name(name, state), rank);
}

inline unsigned
max_extent(const stk::mesh::FieldBase& field, unsigned)
{
  return (unsigned)field.max_size();
}

namespace nalu_ngp {
/* ngp_utils/NgpFieldManager.h */
class FieldManager
{
public:
  template <class T>
  stk::mesh::NgpField<T> get_field(unsigned ord) const
  {
    const auto& f = nwref::World::self().fields.at(ord);
    if constexpr (std::is_same<T, double>::value)
      return stk::mesh::NgpField<T>(f.data, f.ncomp);
    else
      return stk::mesh::NgpField<T>(reinterpret_cast<T*>(f.idata), f.ncomp);
  }
};

/* ngp_utils/NgpLoopUtils.h: EntityInfo and the edge loop shell */
template <class Mesh>
struct EntityInfo
{
  stk::mesh::FastMeshIndex meshIdx;
  stk::mesh::Entity entity;
  stk::mesh::Entity entityNodes[2];
};

template <class Mesh, class Lambda>
void
run_edge_algorithm(
  const std::string&, const Mesh&, const stk::mesh::Selector&, const Lambda& f)
{
  const auto& w = nwref::World::self();
  for (long e = 0; e < w.nEdges; ++e) {
    EntityInfo<Mesh> info;
    info.meshIdx = stk::mesh::FastMeshIndex{0u, (unsigned)e};
    info.entity.m_value = (uint64_t)e;
    info.entityNodes[0].m_value = (uint64_t)w.edgeNodes[2 * e];
    info.entityNodes[1].m_value = (uint64_t)w.edgeNodes[2 * e + 1];
    f(info);
  }
}

/* ngp_utils/NgpTypes.h, NgpLoopUtils.h: node loop */
template <class Mesh = stk::mesh::NgpMesh>
struct NGPMeshTraits
{
  using MeshIndex = stk::mesh::FastMeshIndex;
};
template <class Mesh, class Lambda>
void
run_entity_algorithm(
  const std::string&, const Mesh&, stk::topology::rank_t rank,
  const stk::mesh::Selector&, const Lambda& f)
{
  const auto& w = nwref::World::self();
  const long n = rank == stk::topology::EDGE_RANK ? w.nEdges : w.nNodes;
  const bool sel = rank == stk::topology::NODE_RANK && !w.nodeSelected.empty();
  for (long i = 0; i < n; ++i)
    if (!sel || w.nodeSelected[i])
      f(stk::mesh::FastMeshIndex{0u, (unsigned)i});
}

/* ngp_utils/NgpFieldOps.h, edge_nodal_field_updater: the reference adds with
 * Kokkos::atomic_add from a parallel loop; the serial loop here adds in edge
 * order (what the oracle does, edge_oracle.cpp orc_nodal_grad_edge) */
template <class Mesh, class Field>
struct EdgeNodalUpdaterShim
{
  struct Ops
  {
    double& ref;
    void operator=(const double& v) const { ref = v; }
    void operator+=(const double& v) const { ref += v; }
    void operator-=(const double& v) const { ref += -v; }
  };
  Ops operator()(const EntityInfo<Mesh>& einfo, unsigned ni, unsigned ic) const
  {
    const stk::mesh::FastMeshIndex i{0u, (unsigned)einfo.entityNodes[ni].m_value};
    return Ops{fld.get(i, ic)};
  }
  Field fld;
};
template <class Mesh, class Field>
EdgeNodalUpdaterShim<Mesh, Field>
edge_nodal_field_updater(const Mesh&, const Field& fld)
{
  return EdgeNodalUpdaterShim<Mesh, Field>{fld};
}

/* ngp_utils/NgpMeshInfo.h */
class MeshInfoShim
{
public:
  const stk::mesh::MetaData& meta() const { return bulk_.mesh_meta_data(); }
  const stk::mesh::NgpMesh& ngp_mesh() const { return mesh_; }
  const FieldManager& ngp_field_manager() const { return fm_; }
  stk::mesh::BulkData bulk_;
  stk::mesh::NgpMesh mesh_;
  FieldManager fm_;
};
/* ngp_utils/NgpFieldUtils.h */
template <class T = double>
inline stk::mesh::NgpField<T>
get_ngp_field(
  const MeshInfoShim& meshInfo, const std::string& fieldName,
  const stk::mesh::EntityRank& rank = stk::topology::NODE_RANK)
{
  return meshInfo.ngp_field_manager().template get_field<T>(
    nwref::World::self().ordinal(fieldName, rank));
}
} // namespace nalu_ngp

class SolutionOptions
{
public:
  bool realm_has_vof_ = false;
  bool use_balanced_buoyancy_force_ = false;
  TurbulenceModel turbulenceModel_ = TurbulenceModel::LAMINAR;
  double get_relaxation_factor(const std::string& dof) const
  {
    return nwref::World::self().get("relax:" + dof);
  }
  std::vector<double> get_gravity_vector(const unsigned nDim) const
  {
    const auto& w = nwref::World::self();
    return std::vector<double>(w.gravity, w.gravity + nDim);
  }
  std::string get_coordinates_name() const { return "coordinates"; }
};

class OversetInfo;
class OversetManager
{
public:
  stk::mesh::Ghosting* oversetGhosting_ = nullptr;
  std::vector<OversetInfo*> oversetInfoVec_;
};
class OversetInfo
{
public:
  stk::mesh::Entity orphanNode_;
  stk::mesh::Entity owningElement_;
};
class NonConformalManager
{
public:
  stk::mesh::Ghosting* nonConformalGhosting_ = nullptr;
};

class Realm
{
public:
  Realm() : solutionOptions_(&so_)
  {
    const auto& w = nwref::World::self();
    hypreILower_ = (HypreIntType)w.hypreILower;
    hypreIUpper_ = (HypreIntType)w.hypreIUpper;
    hypreNumNodes_ = (HypreIntType)w.hypreNumNodes;
    hypreOffsets_.assign(w.hypreOffsets.begin(), w.hypreOffsets.end());
    for (size_t i = 0; i < w.fields.size(); ++i) {
      if (w.fields[i].name == "hypre_global_id")
        hypreGlobalId_ = static_cast<HypreIDFieldType*>(w.fieldHandles[i]);
      if (w.fields[i].name == "nalu_global_id")
        naluGlobalId_ = static_cast<GlobalIdFieldType*>(w.fieldHandles[i]);
    }
  }
  stk::mesh::MetaData& meta_data() const
  {
    return const_cast<stk::mesh::MetaData&>(bulk_.mesh_meta_data());
  }
  const stk::mesh::BucketVector&
  get_buckets(stk::mesh::EntityRank rank, const stk::mesh::Selector& s) const
  {
    return bulk_.get_buckets(rank, s);
  }
  stk::mesh::PartVector get_slave_part_vector() const { return stk::mesh::PartVector(); }
  HypreIDFieldType* hypreGlobalId_ = nullptr;
  GlobalIdFieldType* naluGlobalId_ = nullptr;
  HypreIntType hypreILower_ = 0, hypreIUpper_ = 0, hypreNumNodes_ = 0;
  std::vector<HypreIntType> hypreOffsets_;
  OversetManager* oversetManager_ = nullptr;
  NonConformalManager* nonConformalManager_ = nullptr;
  bool isFinalOuterIter_ = false;
  double l2Scaling_ = 1.0;
  bool hasPeriodic_ = false;
  bool hasOverset_ = false;
  stk::mesh::BulkData& bulk_data() { return bulk_; }
  const stk::mesh::NgpMesh& ngp_mesh() const { return ngpMesh_; }
  const nalu_ngp::FieldManager& ngp_field_manager() const { return fm_; }
  const nalu_ngp::MeshInfoShim& mesh_info() const { return meshInfo_; }
  std::string get_coordinates_name() const { return "coordinates"; }
  bool does_mesh_move() const { return false; }
  bool has_mesh_deformation() const { return w().get("mesh_deformation", 0.0) != 0.0; }
  bool is_turbulent() const { return false; }
  double get_divU() const { return w().get("divU"); }
  double get_alpha_factor(const std::string d) const { return w().get("alpha:" + d); }
  double get_alpha_upw_factor(const std::string d) const { return w().get("alpha_upw:" + d); }
  double get_upw_factor(const std::string d) const { return w().get("upw:" + d); }
  bool primitive_uses_limiter(const std::string d) const { return w().get("limiter:" + d) != 0.0; }
  bool get_noc_usage(const std::string d) const { return w().get("noc:" + d) != 0.0; }
  double get_mdot_interp() const { return w().get("mdot_interp"); }
  double get_incompressible_solve() const { return w().get("solve_incompressible"); }
  double get_time_step() const { return w().get("dt"); }
  double get_gamma1() const { return w().get("gamma1"); }
  double get_gamma2() const { return w().get("gamma2"); }
  double get_gamma3() const { return w().get("gamma3"); }
  stk::mesh::Selector get_inactive_selector() const { return stk::mesh::Selector(); }

  SolutionOptions so_;
  SolutionOptions* solutionOptions_;

private:
  static const nwref::World& w() { return nwref::World::self(); }
  stk::mesh::BulkData bulk_;
  stk::mesh::NgpMesh ngpMesh_;
  nalu_ngp::FieldManager fm_;
  nalu_ngp::MeshInfoShim meshInfo_;
};

class LinearSystem;
class EquationSystem
{
public:
  explicit EquationSystem(int numDof) : realm_(default_realm()), numDof_(numDof) {}
  EquationSystem(int numDof, Realm& realm) : realm_(realm), numDof_(numDof) {}
  static Realm& default_realm()
  {
    static Realm r;
    return r;
  }
  /* what NGPApplyCoeff (src/SolverAlgorithm.C) asks of its equation system */
  Realm& realm_;
  LinearSystem* linsys_ = nullptr;
  bool extractDiagonal_ = false;
  bool resetOversetRows_ = true;
  /* host-side twin of extract_diagonal (src/EquationSystem.C); not on the NGP path */
  void save_diagonal_term(
    const std::vector<stk::mesh::Entity>&, const std::vector<int>&,
    const std::vector<double>&)
  {
  }
  std::string diagonalFieldName_ = "momentum_diag";
  ScalarFieldType* get_diagonal_field() const
  {
    const auto& w = nwref::World::self();
    return static_cast<ScalarFieldType*>(
      w.fieldHandles.at(w.ordinal(diagonalFieldName_, stk::topology::NODE_RANK)));
  }
  /* src/EquationSystem.C: builds the blending function the input file names */
  template <class T>
  PecletFunction<T>* ngp_create_peclet_function(const std::string&)
  {
    const auto& w = nwref::World::self();
    if (w.pecletForm == 0)
      return new ClassicPecletFunction<T>((T)5.0, (T)w.pecletA);
    return new TanhFunction<T>((T)w.pecletA, (T)w.pecletB);
  }
  int numDof_;
  std::string name_ = "EqSys";
  std::string userSuppliedName_ = "EqSys";
  int linsysWriteCounter_ = 0;
  bool firstTimeStepSolve_ = true;
};

/* Algorithm.h */
class Algorithm
{
public:
  Algorithm(Realm& realm, stk::mesh::Part* part) : realm_(realm), partVec_(1, part) {}
  virtual ~Algorithm() {}
  virtual void execute() = 0;
  Realm& realm_;
  stk::mesh::PartVector partVec_;
};

/* SharedMemData.h: the edge scratch the lambdas write */
struct SharedMemData_EdgeShim
{
  SharedMemView<double*, DeviceShmem> rhs;
  SharedMemView<double**, DeviceShmem> lhs;
};

/* AssembleEdgeSolverAlgorithm.h: same members the derived classes use; the loop
 * shell (include/AssembleEdgeSolverAlgorithm.h:47-100) as a serial loop that
 * records the local block where the reference hands it to the CoeffApplier */
class AssembleEdgeSolverAlgorithm : public Algorithm
{
public:
  using DblType = double;
  using ShmemDataType = SharedMemData_EdgeShim;

  AssembleEdgeSolverAlgorithm(
    Realm& realm, stk::mesh::Part* part, EquationSystem* eqSystem)
    : Algorithm(realm, part), eqSystem_(eqSystem), rhsSize_(2 * eqSystem->numDof_)
  {
  }
  virtual ~AssembleEdgeSolverAlgorithm() = default;

  template <typename LambdaFunction>
  void run_algorithm(stk::mesh::BulkData&, LambdaFunction lambdaFunc)
  {
    auto& w = nwref::World::self();
    const int n = rhsSize_;
    std::vector<double> l((size_t)n * n), r(n);
    ShmemDataType smdata;
    smdata.lhs = SharedMemView<double**, DeviceShmem>(l.data(), n, n);
    smdata.rhs = SharedMemView<double*, DeviceShmem>(r.data(), n);
    for (long e = 0; e < w.nEdges; ++e) {
      const stk::mesh::FastMeshIndex edge{0u, (unsigned)e};
      const stk::mesh::FastMeshIndex nodeL{0u, (unsigned)w.edgeNodes[2 * e]};
      const stk::mesh::FastMeshIndex nodeR{0u, (unsigned)w.edgeNodes[2 * e + 1]};
      set_vals(smdata.rhs, 0.0);
      set_vals(smdata.lhs, 0.0);
      lambdaFunc(smdata, edge, nodeL, nodeR);
      if (w.applyHook) {
        w.applyHook(e, l.data(), r.data(), n);
        continue;
      }
      for (int i = 0; i < n * n; ++i)
        w.lhsOut[(size_t)e * n * n + i] = l[i];
      for (int i = 0; i < n; ++i)
        w.rhsOut[(size_t)e * n + i] = r[i];
    }
  }

  EquationSystem* eqSystem_;

protected:
  static constexpr stk::mesh::EntityRank entityRank_{stk::topology::EDGE_RANK};
  static constexpr int nodesPerEntity_{2};
  static constexpr int NDimMax_{3};
  const int rhsSize_;
};

} // namespace nalu
} // namespace sierra
#endif
