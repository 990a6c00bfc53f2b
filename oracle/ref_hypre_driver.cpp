/*
 * oracle/ref_hypre_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points over the reference's own src/HypreLinearSystem.C and
 * src/HypreUVWLinearSystem.C (compiled unmodified, oracle/Makefile.ref, against
 * the stand-ins of oracle/ref_shim/{,nalu,hypre}): graph construction
 * (beginLinearSystemConstruction, buildEdgeToNodeGraph, buildDirichletNodeGraph,
 * finalizeLinearSystem with the CSR / shared-row / periodic device structures),
 * zeroSystem + resetCoeffApplierData, the CoeffApplier's operator() (sort,
 * sum_into / sum_into_1DoF, the UVW variant) edge by edge, and loadComplete --
 * where the stand-in hypre IJ interface records what the reference hands to
 * HYPRE_IJMatrixSetValues2 / AddToValues2 and the IJVector calls.  One process
 * plays one MPI rank at a time (ref_world_parallel).
 */
#include <HypreLinearSystem.h>
#include <HypreUVWLinearSystem.h>
#include <edge_kernels/MomentumEdgeSolverAlg.h>
#include <edge_kernels/ScalarEdgeSolverAlg.h>
#include <edge_kernels/ContinuityEdgeSolverAlg.h>
#include <SolverAlgorithm.h>

#include <cstring>
#include <memory>
#include <string>

using namespace sierra::nalu;
using nwref::World;

namespace sierra {
namespace nalu {
/* src/LinearSystem.C needs Tpetra; the base class' few out-of-line members */
LinearSystem::LinearSystem(
  Realm& realm, const unsigned numDof, EquationSystem* eqSys,
  LinearSolver* linearSolver)
  : realm_(realm),
    eqSys_(eqSys),
    inConstruction_(false),
    numDof_(numDof),
    eqSysName_(eqSys->name_),
    linearSolver_(linearSolver),
    linearSolveIterations_(0),
    nonLinearResidual_(0.0),
    linearResidual_(0.0),
    firstNonLinearResidual_(1.0e8),
    scaledNonLinearResidual_(1.0e8),
    recomputePreconditioner_(true),
    reusePreconditioner_(false),
    provideOutput_(true)
{
}
const LinearSolverConfig&
LinearSystem::config() const
{
  return *linearSolver_->getConfig();
}
void LinearSystem::sync_field(const stk::mesh::FieldBase*) {}
bool LinearSystem::debug() { return false; }
double LinearSystem::get_timer_precond() { return 0.0; }
void LinearSystem::zero_timer_precond() {}
bool LinearSystem::useSegregatedSolver() const { return false; }
} // namespace nalu
} // namespace sierra

extern int g_nwref_rank, g_nwref_size; /* ref_driver.cpp: NaluEnv */

namespace {

std::string g_err;

template <class F>
int
guarded(F&& f)
{
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

/* reach the protected applier */
struct ProbeSys : HypreLinearSystem
{
  using HypreLinearSystem::HypreLinearSystem;
  HypreLinSysCoeffApplier* applier()
  {
    return dynamic_cast<HypreLinSysCoeffApplier*>(hostCoeffApplier.get());
  }
};
struct ProbeUvw : HypreUVWLinearSystem
{
  using HypreUVWLinearSystem::HypreUVWLinearSystem;
  HypreLinSysCoeffApplier* applier()
  {
    return dynamic_cast<HypreLinSysCoeffApplier*>(hostCoeffApplier.get());
  }
};

struct Handle
{
  Realm realm;
  EquationSystem eq;
  HypreUVWSolver solver;
  std::unique_ptr<ProbeSys> sys;
  std::unique_ptr<ProbeUvw> uvw;
  stk::mesh::Part part;
  int numDof, nrhs;
  size_t firstMatrix, firstVector;
  Handle(int isUvw, int nd)
    : eq(nd), numDof(isUvw ? 1 : nd), nrhs(isUvw ? nd : 1)
  {
    firstMatrix = nwref::Recorder::self().matrices.size();
    firstVector = nwref::Recorder::self().vectors.size();
    if (isUvw)
      uvw.reset(new ProbeUvw(realm, nd, &eq, &solver));
    else
      sys.reset(new ProbeSys(realm, nd, &eq, &solver));
  }
  HypreLinearSystem& ls() { return uvw ? static_cast<HypreLinearSystem&>(*uvw) : *sys; }
  HypreLinearSystem::HypreLinSysCoeffApplier* app()
  {
    return uvw ? uvw->applier() : sys->applier();
  }
};

} // namespace

extern "C" {

const char*
ref_hypre_last_error()
{
  return g_err.c_str();
}

/* the decomposition as one rank sees it; iupper is one past the last owned
 * node row (Realm::hypreIUpper_), offsets[nranks + 1] (Realm::hypreOffsets_) */
void
ref_world_parallel(
  int rank, int nranks, const long* nodeIdentifier, const int* nodeOwner,
  long ilower, long iupper, long numNodes, const int* offsets)
{
  auto& w = World::self();
  w.rank = rank;
  w.nranks = nranks;
  g_nwref_rank = rank;
  g_nwref_size = nranks;
  w.nodeIdentifier = nodeIdentifier;
  w.nodeOwner = nodeOwner;
  w.nodeOfIdentifier.clear();
  if (nodeIdentifier)
    for (long i = 0; i < w.nNodes; ++i)
      w.nodeOfIdentifier[nodeIdentifier[i]] = i;
  w.hypreILower = ilower;
  w.hypreIUpper = iupper;
  w.hypreNumNodes = numNodes;
  w.hypreOffsets.assign(offsets, offsets + nranks + 1);
}

/* integer nodal fields: "hypre_global_id", "nalu_global_id" */
void
ref_world_int_field(const char* name, int rank, int ncomp, int* data)
{
  auto& w = World::self();
  nwref::FieldRec r{name, rank, ncomp, nullptr};
  r.idata = data;
  w.fields.push_back(r);
  w.fieldHandles.push_back(
    new stk::mesh::Field<int>(name, (unsigned)w.fields.size() - 1, ncomp));
}

/* uvw != 0: HypreUVWLinearSystem (one graph, ndim right-hand sides) */
void*
ref_hypre_create(int uvw, int numDof)
{
  Handle* h = nullptr;
  if (guarded([&] { h = new Handle(uvw, numDof); }))
    return nullptr;
  return h;
}

void
ref_hypre_destroy(void* p)
{
  delete static_cast<Handle*>(p);
}

/* applyDirichletBCs bookkeeping: buildDirichletNodeGraph(vector<Entity>) */
int
ref_hypre_dirichlet_nodes(void* p, const int* nodes, int n)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    std::vector<stk::mesh::Entity> v(n);
    for (int i = 0; i < n; ++i)
      v[i].m_value = (uint64_t)nodes[i];
    h->ls().buildDirichletNodeGraph(v);
  });
}

int
ref_hypre_build_edge_graph_and_finalize(void* p)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    stk::mesh::PartVector parts(1, &h->part);
    h->ls().buildEdgeToNodeGraph(parts);
    h->ls().finalizeLinearSystem();
  });
}

/* out: numRowsOwned, nnzOwned, numRowsShared, nnzShared, nPeriodicRows */
int
ref_hypre_sizes(void* p, long* out)
{
  auto* h = static_cast<Handle*>(p);
  auto* a = h->app();
  out[0] = a->num_rows_owned_;
  out[1] = a->num_nonzeros_owned_;
  out[2] = a->num_rows_shared_;
  out[3] = a->num_nonzeros_shared_;
  out[4] = (long)a->periodic_bc_rows_owned_.extent(0);
  return 0;
}

/* CSR structures of the CoeffApplier and the per-nnz row ids of the system */
int
ref_hypre_graph(
  void* p, int* rowStartOwned, int* rowStartShared, int* cols, int* rows,
  int* rowIndicesShared, int* periodicRows)
{
  auto* h = static_cast<Handle*>(p);
  auto* a = h->app();
  auto& ls = h->ls();
  const long nro = a->num_rows_owned_, nrs = a->num_rows_shared_;
  const long nnz = a->num_nonzeros_owned_ + a->num_nonzeros_shared_;
  for (long i = 0; i <= nro; ++i)
    rowStartOwned[i] = (int)a->mat_row_start_owned_(i);
  if (nrs > 0)
    for (long i = 0; i <= nrs; ++i)
      rowStartShared[i] = (int)a->mat_row_start_shared_(i);
  for (long i = 0; i < nnz; ++i) {
    cols[i] = (int)a->cols_dev_(i);
    rows[i] = (int)ls.rows_host_(i);
  }
  for (long i = 0; i < nrs; ++i)
    rowIndicesShared[i] = (int)ls.row_indices_shared_host_(i);
  for (size_t i = 0; i < a->periodic_bc_rows_owned_.extent(0); ++i)
    periodicRows[i] = (int)a->periodic_bc_rows_owned_(i);
  return 0;
}

/* zeroSystem, get_coeff_applier (resetCoeffApplierData), then operator() for
 * every edge with its block lhs[e][n][n], rhs[e][n] (n = 2 numDof of the
 * calling algorithm: 2 for 1-dof systems, 2 ndim for momentum) */
int
ref_hypre_assemble(void* p, const double* lhs, const double* rhs, int n)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    auto& w = World::self();
    h->ls().zeroSystem();
    auto* dev = dynamic_cast<HypreLinearSystem::HypreLinSysCoeffApplier*>(
      h->ls().get_coeff_applier());
    std::vector<int> ids(n), perm(n);
    SharedMemView<int*, DeviceShmem> idv(ids.data(), n), pv(perm.data(), n);
    stk::mesh::Entity nodes[2];
    for (long e = 0; e < w.nEdges; ++e) {
      nodes[0].m_value = (uint64_t)w.edgeNodes[2 * e];
      nodes[1].m_value = (uint64_t)w.edgeNodes[2 * e + 1];
      stk::mesh::NgpMesh::ConnectedNodes cn(nodes, 2);
      SharedMemView<const double*, DeviceShmem> rv(rhs + (size_t)e * n, n);
      SharedMemView<const double**, DeviceShmem> lv(lhs + (size_t)e * n * n, n, n);
      (*dev)(2, cn, idv, pv, rv, lv, "ref_hypre_assemble");
    }
  });
}

/* One assembly as the reference runs it: zeroSystem, get_coeff_applier, the
 * edge algorithm's execute() whose shell hands every local block straight to
 * the reference's CoeffApplier (no recording), loadComplete.  alg: 0 momentum
 * (system created with numDof = ndim, UVW or monolithic), 1 continuity,
 * 2 scalar (q, dqdx, dflux name the fields); diagField non-empty: the
 * extract_diagonal side channel of NGPApplyCoeff.  For timing the reference's own
 * code (tools/cpu_reference_code_timing.py) and as an end-to-end check. */
int
ref_hypre_sweep(
  void* p, int alg, const char* q, const char* dqdx, const char* dflux,
  const char* diagField)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    auto& w = World::self();
    h->ls().zeroSystem();
    /* the applier the reference's loop shell calls: NGPApplyCoeff
     * (SolverAlgorithm.h:89, src/SolverAlgorithm.C:49-151) -- extract_diagonal
     * into `diagField` when given, then the linear system's CoeffApplier,
     * which its constructor fetches with get_coeff_applier() */
    h->eq.linsys_ = &h->ls();
    h->eq.extractDiagonal_ = diagField && diagField[0];
    if (h->eq.extractDiagonal_)
      h->eq.diagonalFieldName_ = diagField;
    NGPApplyCoeff applier(&h->eq);
    const int nmax = 2 * 3;
    std::vector<int> ids(nmax), perm(nmax);
    std::vector<double> lbuf(nmax * nmax), rbuf(nmax);
    stk::mesh::Entity nodes[2];
    w.applyHook = [&](long e, const double* lhs, const double* rhs, int n) {
      nodes[0].m_value = (uint64_t)w.edgeNodes[2 * e];
      nodes[1].m_value = (uint64_t)w.edgeNodes[2 * e + 1];
      stk::mesh::NgpMesh::ConnectedNodes cn(nodes, 2);
      std::memcpy(lbuf.data(), lhs, sizeof(double) * n * n);
      std::memcpy(rbuf.data(), rhs, sizeof(double) * n);
      SharedMemView<int*, DeviceShmem> idv(ids.data(), n), pv(perm.data(), n);
      SharedMemView<double*, DeviceShmem> rv(rbuf.data(), n);
      SharedMemView<double**, DeviceShmem> lv(lbuf.data(), n, n);
      applier(2, cn, idv, pv, rv, lv, "ref_hypre_sweep");
    };
    try {
      auto handle = [&](const char* nm) {
        return static_cast<stk::mesh::Field<double>*>(
          w.fieldHandles.at(w.ordinal(nm, stk::topology::NODE_RANK)));
      };
      if (alg == 0) {
        MomentumEdgeSolverAlg a(h->realm, &h->part, &h->eq);
        a.execute();
      } else if (alg == 1) {
        ContinuityEdgeSolverAlg a(h->realm, &h->part, &h->eq);
        a.execute();
      } else {
        ScalarEdgeSolverAlg a(
          h->realm, &h->part, &h->eq, handle(q), handle(dqdx), handle(dflux));
        a.execute();
      }
    } catch (...) {
      w.applyHook = nullptr;
      throw;
    }
    w.applyHook = nullptr;
    h->ls().loadComplete();
  });
}

/* CoeffApplier::resetRows (FixPressureAtNodeAlgorithm, include/LinearSystem.h:53-60)
 * on the system as it stands after ref_hypre_assemble */
int
ref_hypre_reset_rows(void* p, const int* nodes, int n, double diag, double rhs)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    std::vector<stk::mesh::Entity> v(n);
    for (int i = 0; i < n; ++i)
      v[i].m_value = (uint64_t)nodes[i];
    h->app()->resetRows((unsigned)n, v.data(), 0, h->numDof, diag, rhs);
  });
}

/* HypreLinearSystem::applyDirichletBCs over the given nodes (the part selector
 * of the reference becomes a node list) */
int
ref_hypre_apply_dirichlet(
  void* p, const char* solution, const char* bcValues, const int* nodes, int n)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] {
    auto& w = World::self();
    w.nodeSelected.assign((size_t)w.nNodes, 0);
    for (int i = 0; i < n; ++i)
      w.nodeSelected[nodes[i]] = 1;
    auto handle = [&](const char* nm) {
      return w.fieldHandles.at(w.ordinal(nm, stk::topology::NODE_RANK));
    };
    stk::mesh::PartVector parts(1, &h->part);
    try {
      h->ls().applyDirichletBCs(handle(solution), handle(bcValues), parts, 0, h->numDof);
    } catch (...) {
      w.nodeSelected.clear();
      throw;
    }
    w.nodeSelected.clear();
  });
}

/* values[nnzOwned + nnzShared]; rhs[nrhs][numRowsOwned + numRowsShared rows x
 * numDof] as the applier holds them (rhs_dev_ is LayoutLeft: one column per
 * right-hand side) */
int
ref_hypre_values(void* p, double* values, double* rhs)
{
  auto* h = static_cast<Handle*>(p);
  auto* a = h->app();
  const size_t nnz = a->values_dev_.extent(0);
  for (size_t i = 0; i < nnz; ++i)
    values[i] = a->values_dev_(i);
  const size_t nr = a->rhs_dev_.extent(0), nc = a->rhs_dev_.extent(1);
  for (size_t c = 0; c < nc; ++c)
    for (size_t r = 0; r < nr; ++r)
      rhs[c * nr + r] = a->rhs_dev_(r, c);
  return 0;
}

int
ref_hypre_rhs_shape(void* p, long* out)
{
  auto* a = static_cast<Handle*>(p)->app();
  out[0] = (long)a->rhs_dev_.extent(0);
  out[1] = (long)a->rhs_dev_.extent(1);
  out[2] = (long)a->values_dev_.extent(0);
  return 0;
}

int
ref_hypre_load_complete(void* p)
{
  auto* h = static_cast<Handle*>(p);
  return guarded([&] { h->ls().loadComplete(); });
}

/* what loadComplete handed to the IJ interface: the calls on the matrix created
 * by the last zeroSystem.  which: 0 matrix, 1.. = right-hand-side vector d */
int
ref_hypre_ij_calls(void* p, int which)
{
  auto* h = static_cast<Handle*>(p);
  auto& R = nwref::Recorder::self();
  if (which == 0)
    return (int)R.matrices.back()->calls.size();
  const size_t nv = R.vectors.size();
  /* zeroSystem creates rhs then sln (UVW: rhs[d], sln[d] pairs) */
  const size_t per = 2 * (size_t)h->nrhs;
  nwref::IJVector* v = R.vectors[nv - per + 2 * (size_t)(which - 1)];
  return (int)v->calls.size();
}

static nwref::IJCall*
ij_call(Handle* h, int which, int k)
{
  auto& R = nwref::Recorder::self();
  if (which == 0)
    return &R.matrices.back()->calls.at(k);
  const size_t nv = R.vectors.size();
  const size_t per = 2 * (size_t)h->nrhs;
  return &R.vectors[nv - per + 2 * (size_t)(which - 1)]->calls.at(k);
}

/* sizes of call k: out = nrows, nvalues, isAddTo */
int
ref_hypre_ij_call_sizes(void* p, int which, int k, long* out)
{
  auto* c = ij_call(static_cast<Handle*>(p), which, k);
  out[0] = (long)c->rows.size();
  out[1] = (long)c->values.size();
  out[2] = c->what == "AddTo";
  return 0;
}

int
ref_hypre_ij_call_get(
  void* p, int which, int k, int* ncols, int* rows, int* cols, double* values)
{
  auto* c = ij_call(static_cast<Handle*>(p), which, k);
  for (size_t i = 0; i < c->rows.size(); ++i) {
    ncols[i] = c->ncols[i];
    rows[i] = c->rows[i];
  }
  for (size_t i = 0; i < c->values.size(); ++i) {
    cols[i] = c->cols[i];
    values[i] = c->values[i];
  }
  return 0;
}

} // extern "C"
