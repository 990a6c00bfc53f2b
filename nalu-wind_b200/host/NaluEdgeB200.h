/*
 * NaluEdgeB200.h -- host C++ mirror of the reference's edge-assembly surface.
 *
 * Header-only shim over the C ABI (include/nalu_edge_b200.h).  It keeps the
 * reference's class and method names for the hot path so that a caller written
 * against nalu-wind's interface reads the same:
 *
 *   reference (Exawind/nalu-wind)                           here
 *   ------------------------------------------------------  -------------------
 *   Realm (mesh, fields, option getters)                     Realm
 *     get_noc_usage / get_mdot_interp / get_alpha_factor ... same names
 *   LinearSystem (include/LinearSystem.h:88-256)             LinearSystem
 *   HypreLinearSystem / HypreUVWLinearSystem                 same names
 *     buildEdgeToNodeGraph, finalizeLinearSystem, zeroSystem,
 *     loadComplete                                           same names
 *   EquationSystem::linsys_, ::name                          EquationSystem
 *   Algorithm::execute (include/Algorithm.h)                 Algorithm
 *   SolverAlgorithm::initialize_connectivity                 SolverAlgorithm
 *   AssembleEdgeSolverAlgorithm                              same name
 *   MomentumEdgeSolverAlg / ContinuityEdgeSolverAlg /
 *   ScalarEdgeSolverAlg (src/edge_kernels/ *.C)               same names
 *   MdotEdgeAlg, NodalGradEdgeAlg<Phi,Grad> + aliases,
 *   MomentumEdgePecletAlg                                    same names
 *
 * What changes for the caller: fields are registered on the Realm by name and
 * moved with Realm::upload / download instead of stk::mesh::NgpField sync calls;
 * the per-edge lambda + CoeffApplier call of run_algorithm() runs inside the
 * CUDA kernels, so there is no device functor to capture.  Errors are C++
 * exceptions (std::runtime_error), as in the reference's host code.
 */
#ifndef NALU_EDGE_B200_HOST_H
#define NALU_EDGE_B200_HOST_H

#include <map>
#include <stdexcept>
#include <memory>
#include <string>
#include <vector>

#include "nalu_edge_b200.h"

namespace sierra {
namespace nalu {

inline void
nw_check(int rc)
{
  if (rc != NW_OK)
    throw std::runtime_error(nw_last_error());
}

/* SolutionOptions defaults: src/SolutionOptions.C:32-52 */
struct SolutionOptions
{
  double hybridDefault_ = 0.0, alphaDefault_ = 0.0, alphaUpwDefault_ = 1.0,
         upwDefault_ = 1.0, relaxFactorDefault_ = 1.0;
  bool nocDefault_ = true;
  double includeDivU_ = 0.0;
  bool mdotInterpRhoUTogether_ = true;
  bool solveIncompressibleContinuity_ = false;
  /* SolutionOptions::realm_has_vof_: MomentumEdgeSolverAlg then reads the edge
   * field "mass_vof_balanced_flow_rate" as well
   * (src/edge_kernels/MomentumEdgeSolverAlg.C:52-64, 88) */
  bool realm_has_vof_ = false;
  std::map<std::string, double> hybridMap_, alphaMap_, alphaUpwMap_, upwMap_,
    relaxFactorMap_, tanhTransMap_, tanhWidthMap_;
  std::map<std::string, bool> nocMap_, limiterMap_;
  std::map<std::string, std::string> tanhFormMap_; /* "classic" | "tanh" */
  double get_relaxation_factor(const std::string& dof) const
  {
    auto it = relaxFactorMap_.find(dof);
    return it == relaxFactorMap_.end() ? relaxFactorDefault_ : it->second;
  }
};

/* minimal Realm: owns the device context, this rank's mesh partition and the
 * option getters the edge algorithms call (src/Realm.C:4121-4200, 4353-4371) */
class Realm
{
public:
  Realm(int cudaDevice, const nw_mesh_desc& desc)
  {
    nw_check(nw_ctx_create(cudaDevice, &ctx_));
    try {
      nw_check(nw_mesh_create(ctx_, &desc, &mesh_));
    } catch (...) {
      nw_ctx_destroy(ctx_);
      throw;
    }
    ndim_ = desc.ndim;
  }
  ~Realm()
  {
    nw_mesh_destroy(mesh_);
    nw_ctx_destroy(ctx_);
  }
  Realm(const Realm&) = delete;
  Realm& operator=(const Realm&) = delete;

  int spatial_dimension() const { return ndim_; }
  /* Realm::periodic_field_update (src/Realm.C:3090-3100) */
  void periodic_field_update(const std::string& field)
  {
    nw_check(nw_field_periodic_update(mesh_, field_ordinal(field)));
  }
  /* the udiag post-processing of MomentumEquationSystem::assemble_and_solve
   * (src/LowMachEquationSystem.C:2759-2821) */
  void momentum_diag_post_process(
    double dt, double gamma1, double alphaU,
    const std::string& udiag = "momentum_diag")
  {
    nw_check(nw_momentum_diag_post_process(
      mesh_, field_ordinal(udiag), field_ordinal("density"),
      field_ordinal("dual_nodal_volume"), dt, gamma1, alphaU));
  }
  nw_mesh* mesh() { return mesh_; }
  nw_ctx* ctx() { return ctx_; }

  /* field registry by the reference's field names */
  int register_field(const std::string& name, nw_entity_rank rank, int ncomp)
  {
    int id;
    nw_check(nw_field_register(mesh_, name.c_str(), rank, ncomp, &id));
    return id;
  }
  int field_ordinal(const std::string& name) const
  {
    int id;
    nw_check(nw_field_find(mesh_, name.c_str(), &id));
    return id;
  }
  void upload(const std::string& name, const double* host)
  {
    nw_check(nw_field_upload(mesh_, field_ordinal(name), host));
  }
  /* pipelined upload: copy on the copy stream now, permute in at commit */
  void stage(const std::string& name, const double* pinnedHost)
  {
    nw_check(nw_field_stage(mesh_, field_ordinal(name), pinnedHost));
  }
  void commit(const std::string& name)
  {
    nw_check(nw_field_commit(mesh_, field_ordinal(name)));
  }
  void download(const std::string& name, double* host)
  {
    nw_check(nw_field_download(mesh_, field_ordinal(name), host));
  }
  void sync() { nw_check(nw_ctx_sync(ctx_)); }

  /* time integrator scalars */
  double get_time_step() const { return dt_; }
  double get_gamma1() const { return gamma1_; }
  double get_gamma2() const { return gamma2_; }
  double get_gamma3() const { return gamma3_; }
  void set_time_step(double dt, double gamma1)
  {
    dt_ = dt;
    gamma1_ = gamma1;
  }
  /* TimeIntegrator BDF coefficients (BDF1: 1, -1, 0; BDF2: 1.5, -2, 0.5) */
  void set_bdf(double dt, double gamma1, double gamma2, double gamma3)
  {
    dt_ = dt;
    gamma1_ = gamma1;
    gamma2_ = gamma2;
    gamma3_ = gamma3;
  }
  /* field_of_state: states of field F are registered as F (NP1), F_n, F_nm1;
   * a two-state field has no F_nm1 and uses F_n (number_of_states() == 2,
   * src/node_kernels/ScalarMassBDFNodeKernel.C:36-40) */
  bool has_field(const std::string& name) const
  {
    int id;
    return nw_field_find(mesh_, name.c_str(), &id) == NW_OK;
  }
  int state_ordinal(const std::string& name, int state /* 0 NP1, 1 N, 2 NM1 */) const
  {
    if (state == 0)
      return field_ordinal(name);
    if (state == 2 && has_field(name + "_nm1"))
      return field_ordinal(name + "_nm1");
    return field_ordinal(name + "_n");
  }

  /* option getters, same names and defaults as the reference */
  bool get_noc_usage(const std::string& dof) const
  {
    return lookup(solutionOptions_.nocMap_, dof, solutionOptions_.nocDefault_);
  }
  double get_mdot_interp() const
  {
    return solutionOptions_.mdotInterpRhoUTogether_ ? 1.0 : 0.0;
  }
  double get_incompressible_solve() const
  {
    return solutionOptions_.solveIncompressibleContinuity_ ? 1.0 : 0.0;
  }
  double get_divU() const { return solutionOptions_.includeDivU_; }
  double get_hybrid_factor(const std::string& dof) const
  {
    return lookup(
      solutionOptions_.hybridMap_, dof, solutionOptions_.hybridDefault_);
  }
  double get_alpha_factor(const std::string& dof) const
  {
    return lookup(
      solutionOptions_.alphaMap_, dof, solutionOptions_.alphaDefault_);
  }
  double get_alpha_upw_factor(const std::string& dof) const
  {
    return lookup(
      solutionOptions_.alphaUpwMap_, dof, solutionOptions_.alphaUpwDefault_);
  }
  double get_upw_factor(const std::string& dof) const
  {
    return lookup(solutionOptions_.upwMap_, dof, solutionOptions_.upwDefault_);
  }
  bool primitive_uses_limiter(const std::string& dof) const
  {
    return lookup(solutionOptions_.limiterMap_, dof, false);
  }
  std::string get_tanh_functional_form(const std::string& dof) const
  {
    return lookup(
      solutionOptions_.tanhFormMap_, dof, std::string("classic"));
  }
  double get_tanh_trans(const std::string& dof) const
  {
    return lookup(solutionOptions_.tanhTransMap_, dof, 2.0);
  }
  double get_tanh_width(const std::string& dof) const
  {
    return lookup(solutionOptions_.tanhWidthMap_, dof, 4.0);
  }
  /* EquationSystem::ngp_create_peclet_function, include/EquationSystem.h:399-417 */
  nw_peclet_fn peclet_function(const std::string& dof) const
  {
    nw_peclet_fn f;
    if (get_tanh_functional_form(dof) == "classic") {
      f.form = NW_PECLET_CLASSIC;
      f.a = get_hybrid_factor(dof);
      f.b = 0.0;
    } else {
      f.form = NW_PECLET_TANH;
      f.a = get_tanh_trans(dof);
      f.b = get_tanh_width(dof);
    }
    return f;
  }

  SolutionOptions solutionOptions_;

private:
  template <class M, class V>
  static V lookup(const M& m, const std::string& k, V dflt)
  {
    auto it = m.find(k);
    return it == m.end() ? dflt : V(it->second);
  }
  nw_ctx* ctx_ = nullptr;
  nw_mesh* mesh_ = nullptr;
  int ndim_ = 3;
  double dt_ = 1.0, gamma1_ = 1.0, gamma2_ = -1.0, gamma3_ = 0.0;
};

/* include/LinearSystem.h:88-256 (assembly part) */
class LinearSystem
{
public:
  virtual ~LinearSystem()
  {
    if (ls_)
      nw_linsys_destroy(ls_);
  }
  /* the reference takes a stk::mesh::PartVector; this path has one part */
  virtual void buildEdgeToNodeGraph()
  {
    nw_check(nw_linsys_build_edge_to_node_graph(ls_));
  }
  virtual void finalizeLinearSystem() { nw_check(nw_linsys_finalize(ls_)); }
  virtual void zeroSystem() { nw_check(nw_linsys_zero(ls_)); }
  virtual void loadComplete() { nw_check(nw_linsys_load_complete(ls_)); }
  /* not in the reference: the edge algorithm is the last contribution to rows
   * shared with other ranks before loadComplete, so their exchange may start
   * from inside its execute() (nw_linsys_set_eager_exchange) */
  void eagerExchange(bool on = true)
  {
    nw_check(nw_linsys_set_eager_exchange(ls_, on ? 1 : 0));
  }
  void skipRows(const std::vector<int64_t>& rows)
  {
    nw_check(nw_linsys_set_skipped_rows(ls_, rows.data(), (int64_t)rows.size()));
  }
  /* LinearSystem::applyDirichletBCs (include/LinearSystem.h:163-168): the
   * part vector becomes the list of local nodes of those parts */
  virtual void applyDirichletBCs(
    const std::string& solutionField, const std::string& bcValuesField,
    const std::vector<int32_t>& nodes)
  {
    nw_check(nw_linsys_apply_dirichlet_bcs(
      ls_, realm_.field_ordinal(solutionField), realm_.field_ordinal(bcValuesField),
      (int64_t)nodes.size(), nodes.data()));
  }
  /* CoeffApplier::resetRows (include/LinearSystem.h:53-60) */
  virtual void resetRows(
    const std::vector<int32_t>& nodeList, const unsigned /*beginPos*/,
    const unsigned /*endPos*/, const double diag_value = 0.0,
    const double rhs_residual = 0.0)
  {
    nw_check(nw_linsys_reset_rows(
      ls_, (int64_t)nodeList.size(), nodeList.data(), diag_value, rhs_residual));
  }
  /* CoeffApplier::operator() (include/LinearSystem.h:62-70) for blocks the
   * caller computed on the device */
  void sumInto(
    int64_t numEntities, int nodesPerEntity, const int32_t* d_entityNodes,
    const double* d_lhs, const double* d_rhs)
  {
    nw_check(nw_linsys_sum_into(
      ls_, numEntities, nodesPerEntity, d_entityNodes, d_lhs, d_rhs));
  }
  unsigned numDof() const { return numDof_; }
  nw_linsys* handle() { return ls_; }
  nw_linsys_sizes sizes() const
  {
    nw_linsys_sizes s;
    nw_check(nw_linsys_get_sizes(ls_, &s));
    return s;
  }
  /* host copies in the layout handed to HYPRE_IJMatrixSetValues2 /
   * HYPRE_IJVectorSetValues (src/HypreLinearSystem.C:1572-1590, 1665-1673) */
  void copy_values(std::vector<double>& values, std::vector<double>& rhs)
  {
    const nw_linsys_sizes s = sizes();
    int64_t nx = 0;
    nw_check(nw_linsys_get_extra(ls_, &nx, nullptr, nullptr));
    values.resize(s.num_nonzeros_owned + s.num_nonzeros_shared + nx);
    rhs.resize((s.num_rows_owned + s.num_rows_shared) * s.num_rhs);
    nw_check(nw_linsys_get_values(ls_, values.data(), rhs.data()));
  }

protected:
  LinearSystem(Realm& realm, int kind, unsigned numDof)
    : realm_(realm), numDof_(numDof)
  {
    nw_check(nw_linsys_create(realm.mesh(), kind, (int)numDof, &ls_));
  }
  Realm& realm_;
  unsigned numDof_;
  nw_linsys* ls_ = nullptr;
};

class HypreLinearSystem : public LinearSystem
{
public:
  HypreLinearSystem(Realm& realm, unsigned numDof)
    : LinearSystem(realm, NW_LINSYS_HYPRE, numDof)
  {
  }
};

/* src/HypreUVWLinearSystem.C:15-32: scalar graph, nDim right-hand sides */
class HypreUVWLinearSystem : public LinearSystem
{
public:
  HypreUVWLinearSystem(Realm& realm, unsigned numDof)
    : LinearSystem(realm, NW_LINSYS_HYPRE_UVW, numDof)
  {
  }
};

class EquationSystem
{
public:
  EquationSystem(Realm& realm, const std::string& name)
    : realm_(realm), name_(name)
  {
  }
  Realm& realm_;
  std::string name_;
  LinearSystem* linsys_ = nullptr;
};

class Algorithm
{
public:
  explicit Algorithm(Realm& realm) : realm_(realm) {}
  virtual ~Algorithm() = default;
  virtual void execute() = 0;
  Realm& realm_;
};

class SolverAlgorithm : public Algorithm
{
public:
  SolverAlgorithm(Realm& realm, EquationSystem* eqSystem)
    : Algorithm(realm), eqSystem_(eqSystem)
  {
  }
  virtual void initialize_connectivity() = 0;
  EquationSystem* eqSystem_;
};

/* include/AssembleEdgeSolverAlgorithm.h + src/AssembleEdgeSolverAlgorithm.C:26-30 */
class AssembleEdgeSolverAlgorithm : public SolverAlgorithm
{
public:
  AssembleEdgeSolverAlgorithm(Realm& realm, EquationSystem* eqSystem)
    : SolverAlgorithm(realm, eqSystem)
  {
  }
  void initialize_connectivity() override
  {
    eqSystem_->linsys_->buildEdgeToNodeGraph();
  }
};

/* src/edge_kernels/ContinuityEdgeSolverAlg.C */
class ContinuityEdgeSolverAlg : public AssembleEdgeSolverAlgorithm
{
public:
  using AssembleEdgeSolverAlgorithm::AssembleEdgeSolverAlgorithm;
  void execute() override
  {
    nw_continuity_opts o;
    o.dt = realm_.get_time_step();
    o.gamma1 = realm_.get_gamma1();
    o.noc_fac = realm_.get_noc_usage("pressure") ? 1.0 : 0.0;
    o.interp_together = realm_.get_mdot_interp();
    o.solve_incompressible = realm_.get_incompressible_solve();
    nw_check(nw_assemble_continuity_edge(eqSystem_->linsys_->handle(), &o));
  }
};

/* src/edge_kernels/ScalarEdgeSolverAlg.C */
class ScalarEdgeSolverAlg : public AssembleEdgeSolverAlgorithm
{
public:
  ScalarEdgeSolverAlg(
    Realm& realm,
    EquationSystem* eqSystem,
    const std::string& scalarQ,
    const std::string& dqdx,
    const std::string& diffFluxCoeff)
    : AssembleEdgeSolverAlgorithm(realm, eqSystem),
      dofName_(scalarQ),
      dqdx_(dqdx),
      diffFluxCoeff_(diffFluxCoeff)
  {
  }
  nw_scalar_opts options() const
  {
    nw_scalar_opts o;
    o.alpha = realm_.get_alpha_factor(dofName_);
    o.alpha_upw = realm_.get_alpha_upw_factor(dofName_);
    o.ho_upwind = realm_.get_upw_factor(dofName_);
    o.relax_fac = realm_.solutionOptions_.get_relaxation_factor(dofName_);
    o.use_limiter = realm_.primitive_uses_limiter(dofName_) ? 1 : 0;
    o.eps = 1.0e-16;
    o.pf = realm_.peclet_function(dofName_);
    return o;
  }
  void execute() override
  {
    const nw_scalar_opts o = options();
    nw_check(nw_assemble_scalar_edge(
      eqSystem_->linsys_->handle(), realm_.field_ordinal(dofName_),
      realm_.field_ordinal(dqdx_), realm_.field_ordinal(diffFluxCoeff_), &o));
  }
  /* this algorithm and `other` (another scalar of the same mesh and graph,
   * assembled from the same state: SST's TKE + SDR) in one launch; same result
   * as execute() on both */
  void execute_with(ScalarEdgeSolverAlg& other)
  {
    const nw_scalar_opts oa = options(), ob = other.options();
    nw_check(nw_assemble_scalar_edge_pair(
      eqSystem_->linsys_->handle(), realm_.field_ordinal(dofName_),
      realm_.field_ordinal(dqdx_), realm_.field_ordinal(diffFluxCoeff_), &oa,
      other.eqSystem_->linsys_->handle(), realm_.field_ordinal(other.dofName_),
      realm_.field_ordinal(other.dqdx_),
      realm_.field_ordinal(other.diffFluxCoeff_), &ob));
  }

private:
  std::string dofName_, dqdx_, diffFluxCoeff_;
};

/* src/edge_kernels/MomentumEdgeSolverAlg.C */
class MomentumEdgeSolverAlg : public AssembleEdgeSolverAlgorithm
{
public:
  MomentumEdgeSolverAlg(
    Realm& realm,
    EquationSystem* eqSystem,
    const std::string& viscName = "viscosity")
    : AssembleEdgeSolverAlgorithm(realm, eqSystem), viscName_(viscName)
  {
  }
  /* NGPApplyCoeff::extract_diagonal target ("momentum_diag"), "" = off */
  std::string diagField_;
  bool fusePeclet_ = false;
  void execute() override
  {
    const std::string dof = "velocity";
    nw_momentum_opts o;
    o.include_divu = realm_.get_divU();
    o.alpha = realm_.get_alpha_factor(dof);
    o.alpha_upw = realm_.get_alpha_upw_factor(dof);
    o.ho_upwind = realm_.get_upw_factor(dof);
    o.relax_fac = realm_.solutionOptions_.get_relaxation_factor(dof);
    o.use_limiter = realm_.primitive_uses_limiter(dof) ? 1 : 0;
    o.eps = 1.0e-16;
    o.fuse_peclet = fusePeclet_ ? 1 : 0;
    o.pf = realm_.peclet_function(dof);
    o.pec_eps = 1.0e-16;
    o.diag_field = diagField_.empty() ? -1 : realm_.field_ordinal(diagField_);
    o.has_vof = realm_.solutionOptions_.realm_has_vof_ ? 1 : 0;
    nw_check(nw_assemble_momentum_edge(
      eqSystem_->linsys_->handle(), realm_.field_ordinal(viscName_), &o));
  }

private:
  std::string viscName_;
};

/* src/ngp_algorithms/MdotEdgeAlg.C */
class MdotEdgeAlg : public Algorithm
{
public:
  using Algorithm::Algorithm;
  void execute() override
  {
    nw_mdot_opts o;
    o.noc_fac = realm_.get_noc_usage("pressure") ? 1.0 : 0.0;
    o.interp_together = realm_.get_mdot_interp();
    nw_check(nw_mdot_edge(realm_.mesh(), &o));
  }
};

/* src/edge_kernels/MomentumEdgePecletAlg.C */
class MomentumEdgePecletAlg : public Algorithm
{
public:
  MomentumEdgePecletAlg(Realm& realm, const std::string& viscName = "viscosity")
    : Algorithm(realm), viscName_(viscName)
  {
  }
  void execute() override
  {
    nw_peclet_opts o;
    o.pf = realm_.peclet_function("velocity");
    o.eps = 1.0e-16;
    nw_check(nw_peclet_edge(realm_.mesh(), realm_.field_ordinal(viscName_), &o));
  }

private:
  std::string viscName_;
};

/* src/ngp_algorithms/NodalGradEdgeAlg.C + NodalGradAlgDriver.C (zero, edge
 * contributions, shared-node sum) */
struct ScalarFieldType
{
};
struct VectorFieldType
{
};
struct TensorFieldType
{
};
template <typename PhiType, typename GradPhiType>
class NodalGradEdgeAlg : public Algorithm
{
public:
  NodalGradEdgeAlg(
    Realm& realm, const std::string& phi, const std::string& gradPhi)
    : Algorithm(realm), phi_(phi), gradPhi_(gradPhi)
  {
  }
  void execute() override
  {
    nw_check(nw_nodal_grad_edge(
      realm_.mesh(), realm_.field_ordinal(phi_), realm_.field_ordinal(gradPhi_)));
  }

  /* the SST system's dkdx and dwdx drivers in one launch (same result as the
   * two execute() calls; src/ShearStressTransportEquationSystem.C:247-320) */
  static void execute_pair(
    Realm& realm, const std::string& phiA, const std::string& gradA,
    const std::string& phiB, const std::string& gradB)
  {
    nw_check(nw_nodal_grad_edge_pair(
      realm.mesh(), realm.field_ordinal(phiA), realm.field_ordinal(gradA),
      realm.field_ordinal(phiB), realm.field_ordinal(gradB)));
  }

private:
  std::string phi_, gradPhi_;
};
/* src/edge_kernels/WallDistEdgeSolverAlg.C */
class WallDistEdgeSolverAlg : public AssembleEdgeSolverAlgorithm
{
public:
  using AssembleEdgeSolverAlgorithm::AssembleEdgeSolverAlgorithm;
  void execute() override
  {
    nw_check(nw_assemble_wall_dist_edge(eqSystem_->linsys_->handle()));
  }
};

/* Node kernels (src/node_kernels/) as the reference registers them:
 *   nodeAlg.add_kernel<ScalarMassBDFNodeKernel>("turbulent_ke");
 * Time states of a field follow Realm::state_ordinal. */
struct NodeKernel
{
  virtual ~NodeKernel() = default;
  virtual void execute(Realm& realm, LinearSystem& linsys) = 0;
};

namespace detail {
inline nw_mass_bdf_opts
mass_opts(Realm& realm, const std::string& q, bool hasQ, bool momentum)
{
  nw_mass_bdf_opts o{};
  o.dt = realm.get_time_step();
  o.gamma1 = realm.get_gamma1();
  o.gamma2 = realm.get_gamma2();
  o.gamma3 = realm.get_gamma3();
  o.q_nm1 = o.q_n = o.q_np1 = -1;
  if (hasQ) {
    o.q_np1 = realm.state_ordinal(q, 0);
    o.q_n = realm.state_ordinal(q, 1);
    o.q_nm1 = realm.state_ordinal(q, 2);
  }
  o.rho_np1 = realm.state_ordinal("density", 0);
  o.rho_n = realm.state_ordinal("density", 1);
  o.rho_nm1 = realm.state_ordinal("density", 2);
  /* populate_dnv_states: a static mesh has one dual_nodal_volume state */
  const bool moving = realm.has_field("dual_nodal_volume_n");
  o.dnv_np1 = realm.state_ordinal("dual_nodal_volume", 0);
  o.dnv_n = moving ? realm.state_ordinal("dual_nodal_volume", 1) : o.dnv_np1;
  o.dnv_nm1 = moving ? realm.state_ordinal("dual_nodal_volume", 2) : o.dnv_np1;
  o.dpdx = momentum ? realm.field_ordinal("dpdx") : -1;
  return o;
}
} // namespace detail

/* src/node_kernels/ScalarMassBDFNodeKernel.C */
class ScalarMassBDFNodeKernel : public NodeKernel
{
public:
  explicit ScalarMassBDFNodeKernel(const std::string& scalarQ) : q_(scalarQ) {}
  void execute(Realm& realm, LinearSystem& linsys) override
  {
    const nw_mass_bdf_opts o = detail::mass_opts(realm, q_, true, false);
    nw_check(nw_assemble_mass_bdf_node(linsys.handle(), NW_MASS_SCALAR, &o));
  }

private:
  std::string q_;
};

/* src/node_kernels/MomentumMassBDFNodeKernel.C */
class MomentumMassBDFNodeKernel : public NodeKernel
{
public:
  void execute(Realm& realm, LinearSystem& linsys) override
  {
    const nw_mass_bdf_opts o = detail::mass_opts(realm, "velocity", true, true);
    nw_check(nw_assemble_mass_bdf_node(linsys.handle(), NW_MASS_MOMENTUM, &o));
  }
};

/* src/node_kernels/ContinuityMassBDFNodeKernel.C */
class ContinuityMassBDFNodeKernel : public NodeKernel
{
public:
  void execute(Realm& realm, LinearSystem& linsys) override
  {
    const nw_mass_bdf_opts o = detail::mass_opts(realm, "", false, false);
    nw_check(nw_assemble_mass_bdf_node(linsys.handle(), NW_MASS_CONTINUITY, &o));
  }
};

/* src/node_kernels/WallDistNodeKernel.C */
class WallDistNodeKernel : public NodeKernel
{
public:
  void execute(Realm& realm, LinearSystem& linsys) override
  {
    nw_check(nw_assemble_wall_dist_node(
      linsys.handle(), realm.field_ordinal("dual_nodal_volume")));
  }
};

/* src/AssembleNGPNodeSolverAlgorithm.C */
class AssembleNGPNodeSolverAlgorithm : public SolverAlgorithm
{
public:
  using SolverAlgorithm::SolverAlgorithm;
  /* the reference's node graph (one (row,row) entry per node) is contained in
   * the edge graph of this path */
  void initialize_connectivity() override {}
  template <typename T, class... Args>
  void add_kernel(Args&&... args)
  {
    nodeKernels_.emplace_back(new T(std::forward<Args>(args)...));
  }
  void execute() override
  {
    for (auto& k : nodeKernels_)
      k->execute(realm_, *eqSystem_->linsys_);
  }

private:
  std::vector<std::unique_ptr<NodeKernel>> nodeKernels_;
};

/* src/FixPressureAtNodeAlgorithm.C:57-121: reset the row of the reference
 * node, then sum lhs = 1, rhs = refPressure - p into it */
class FixPressureAtNodeAlgorithm : public Algorithm
{
public:
  FixPressureAtNodeAlgorithm(
    Realm& realm, EquationSystem* eqSystem, int32_t targetNode,
    double refPressure, double pressureAtNode)
    : Algorithm(realm),
      eqSystem_(eqSystem),
      targetNode_(targetNode),
      refPressure_(refPressure),
      pressureN_(pressureAtNode)
  {
  }
  /* d_scratch: device memory for one node index + two doubles (the caller's
   * device allocator; this header stays free of CUDA runtime calls) */
  void execute(int32_t* d_node, double* d_lhs, double* d_rhs)
  {
    eqSystem_->linsys_->resetRows({targetNode_}, 0, 1);
    eqSystem_->linsys_->sumInto(1, 1, d_node, d_lhs, d_rhs);
  }
  void execute() override
  {
    throw std::runtime_error(
      "FixPressureAtNodeAlgorithm: pass device scratch holding {node}, {1.0}, "
      "{refPressure - p} to execute(d_node, d_lhs, d_rhs)");
  }
  double rhs_value() const { return refPressure_ - pressureN_; }

private:
  EquationSystem* eqSystem_;
  int32_t targetNode_;
  double refPressure_, pressureN_;
};

/* src/ngp_algorithms/GeometryAlgDriver.C + GeometryInteriorAlg.C: pre_work
 * zero-fill, interior element blocks, shared-node sum of the volumes */
class GeometryAlgDriver : public Algorithm
{
public:
  using Algorithm::Algorithm;
  /* one call per element block; npe = 8 (Hex8), 4 (Tet4; 2-D mesh: Quad4),
   * 6 (Wed6) or 5 (Pyr5) */
  void register_elem_block(
    int npe, std::vector<int32_t> elemNodes, std::vector<unsigned char> owned = {})
  {
    blocks_.push_back({npe, std::move(elemNodes), std::move(owned)});
  }
  void execute() override
  {
    const int x = realm_.field_ordinal("coordinates");
    const int v = realm_.field_ordinal("dual_nodal_volume");
    const int a = realm_.field_ordinal("edge_area_vector");
    nw_check(nw_field_fill(realm_.mesh(), v, 0.0));
    nw_check(nw_field_fill(realm_.mesh(), a, 0.0));
    for (auto& b : blocks_) {
      const int64_t n = (int64_t)b.nodes.size() / b.npe;
      const unsigned char* ow = b.owned.empty() ? nullptr : b.owned.data();
      /* one GeometryInteriorAlg<AlgTraits> per block topology
       * (GeometryAlgDriver::register_elem_algorithm) */
      auto fn = nw_geometry_interior_hex8;
      if (realm_.spatial_dimension() == 2)
        fn = nw_geometry_interior_quad4;
      else if (b.npe == 4)
        fn = nw_geometry_interior_tet4;
      else if (b.npe == 6)
        fn = nw_geometry_interior_wed6;
      else if (b.npe == 5)
        fn = nw_geometry_interior_pyr5;
      else if (b.npe != 8)
        throw std::runtime_error("GeometryAlgDriver: unsupported element topology");
      nw_check(fn(realm_.mesh(), n, b.nodes.data(), ow, x, v, a));
    }
    nw_check(nw_field_parallel_sum(realm_.mesh(), v));
  }

private:
  struct Block
  {
    int npe;
    std::vector<int32_t> nodes;
    std::vector<unsigned char> owned;
  };
  std::vector<Block> blocks_;
};

using ScalarNodalGradEdgeAlg = NodalGradEdgeAlg<ScalarFieldType, VectorFieldType>;
using VectorNodalGradEdgeAlg = NodalGradEdgeAlg<VectorFieldType, TensorFieldType>;
using TensorNodalGradEdgeAlg = VectorNodalGradEdgeAlg;

} // namespace nalu
} // namespace sierra

#endif
