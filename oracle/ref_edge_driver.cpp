/*
 * oracle/ref_edge_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points that construct the reference's own edge algorithm classes
 *   MomentumEdgeSolverAlg    src/edge_kernels/MomentumEdgeSolverAlg.C
 *   ScalarEdgeSolverAlg      src/edge_kernels/ScalarEdgeSolverAlg.C
 *   ContinuityEdgeSolverAlg  src/edge_kernels/ContinuityEdgeSolverAlg.C
 *   MdotEdgeAlg              src/ngp_algorithms/MdotEdgeAlg.C
 *   NodalGradEdgeAlg         src/ngp_algorithms/NodalGradEdgeAlg.C
 *   MomentumEdgePecletAlg    src/edge_kernels/MomentumEdgePecletAlg.C
 *   WallDistEdgeSolverAlg    src/edge_kernels/WallDistEdgeSolverAlg.C
 *   {Scalar,Momentum,Continuity}MassBDFNodeKernel, WallDistNodeKernel
 *                            src/node_kernels/*.C (setup + execute per node;
 *                            the shell AssembleNGPNodeSolverAlgorithm.C:85-146
 *                            is the serial loop of run_node_kernel below)
 * (compiled unmodified from /root/reference, oracle/Makefile.ref) over the
 * stand-in Realm of oracle/ref_shim/nalu/RefHarness.h and call execute():
 * the per-edge arithmetic that runs is the reference's, the local 2x2 /
 * 2 ndim x 2 ndim blocks it leaves in smdata.lhs / smdata.rhs are recorded
 * edge by edge.  Used by tests/golden/extract_reference_runs.py (fixture) and
 * tests/test_reference_edge_runs.py (live); never by the product.
 */
#include <edge_kernels/MomentumEdgeSolverAlg.h>
#include <edge_kernels/ScalarEdgeSolverAlg.h>
#include <edge_kernels/ContinuityEdgeSolverAlg.h>
#include <edge_kernels/WallDistEdgeSolverAlg.h>
#include <edge_kernels/MomentumEdgePecletAlg.h>
#include <ngp_algorithms/MdotEdgeAlg.h>
#include <ngp_algorithms/NodalGradEdgeAlg.h>
#include <node_kernels/ScalarMassBDFNodeKernel.h>
#include <node_kernels/MomentumMassBDFNodeKernel.h>
#include <node_kernels/ContinuityMassBDFNodeKernel.h>
#include <node_kernels/WallDistNodeKernel.h>

#include <cstring>
#include <string>

using namespace sierra::nalu;
using nwref::World;

extern int g_nwref_rank, g_nwref_size; /* ref_driver.cpp: NaluEnv */

namespace {
std::string g_err;
bool g_has_vof = false, g_buoyancy = false;

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

void
configure(Realm& realm)
{
  realm.so_.realm_has_vof_ = g_has_vof;
  realm.so_.use_balanced_buoyancy_force_ = g_buoyancy;
}
} // namespace

extern "C" {

const char*
ref_last_error()
{
  return g_err.c_str();
}

void
ref_world_reset(int ndim, long nNodes, long nEdges, const int* edgeNodes)
{
  auto& w = World::self();
  for (auto* h : w.fieldHandles)
    delete h;
  w = World();
  w.ndim = ndim;
  w.nNodes = nNodes;
  w.nEdges = nEdges;
  w.edgeNodes = edgeNodes;
  g_has_vof = g_buoyancy = false;
  g_nwref_rank = 0;
  g_nwref_size = 1;
}

/* rank: 0 node, 1 edge; data[entity][ncomp] stays owned by the caller */
void
ref_world_field(const char* name, int rank, int ncomp, double* data)
{
  auto& w = World::self();
  w.fields.push_back(nwref::FieldRec{name, rank, ncomp, data});
  w.fieldHandles.push_back(new stk::mesh::Field<double>(
    name, (unsigned)w.fields.size() - 1, ncomp));
}

void
ref_world_option(const char* key, double value)
{
  World::self().opt[key] = value;
}

void
ref_world_gravity(const double* g)
{
  for (int d = 0; d < 3; ++d)
    World::self().gravity[d] = g[d];
}

void
ref_world_peclet(int form, double a, double b)
{
  auto& w = World::self();
  w.pecletForm = form;
  w.pecletA = a;
  w.pecletB = b;
}

void
ref_world_flags(int has_vof, int balanced_buoyancy)
{
  g_has_vof = has_vof != 0;
  g_buoyancy = balanced_buoyancy != 0;
}

/* lhsOut[nEdges][2 ndim][2 ndim], rhsOut[nEdges][2 ndim] */
int
ref_run_momentum(double* lhsOut, double* rhsOut)
{
  return guarded([&] {
    auto& w = World::self();
    w.lhsOut = lhsOut;
    w.rhsOut = rhsOut;
    Realm realm;
    configure(realm);
    EquationSystem eq(w.ndim);
    stk::mesh::Part part;
    MomentumEdgeSolverAlg alg(realm, &part, &eq);
    alg.execute();
  });
}

/* lhsOut[nEdges][2][2], rhsOut[nEdges][2] */
int
ref_run_scalar(
  const char* q, const char* dqdx, const char* dflux, double* lhsOut,
  double* rhsOut)
{
  return guarded([&] {
    auto& w = World::self();
    w.lhsOut = lhsOut;
    w.rhsOut = rhsOut;
    Realm realm;
    configure(realm);
    EquationSystem eq(1);
    stk::mesh::Part part;
    auto handle = [&](const char* n) {
      return static_cast<stk::mesh::Field<double>*>(
        w.fieldHandles.at(w.ordinal(n, stk::topology::NODE_RANK)));
    };
    ScalarEdgeSolverAlg alg(
      realm, &part, &eq, handle(q), handle(dqdx), handle(dflux));
    alg.execute();
  });
}

int
ref_run_continuity(double* lhsOut, double* rhsOut)
{
  return guarded([&] {
    auto& w = World::self();
    w.lhsOut = lhsOut;
    w.rhsOut = rhsOut;
    Realm realm;
    configure(realm);
    EquationSystem eq(1);
    stk::mesh::Part part;
    ContinuityEdgeSolverAlg alg(realm, &part, &eq);
    alg.execute();
  });
}

/* writes the edge field "mass_flow_rate" */
int
ref_run_mdot()
{
  return guarded([&] {
    Realm realm;
    configure(realm);
    stk::mesh::Part part;
    MdotEdgeAlg alg(realm, &part);
    alg.execute();
  });
}

/* NodalGradEdgeAlg: adds into the nodal field `grad` (dim1 x ndim components;
 * the caller zero-fills it, NodalGradAlgDriver::pre_work) */
int
ref_run_nodal_grad(const char* phi, const char* grad)
{
  return guarded([&] {
    auto& w = World::self();
    Realm realm;
    configure(realm);
    stk::mesh::Part part;
    auto handle = [&](const char* n) {
      return static_cast<stk::mesh::Field<double>*>(
        w.fieldHandles.at(w.ordinal(n, stk::topology::NODE_RANK)));
    };
    ScalarNodalGradEdgeAlg alg(realm, &part, handle(phi), handle(grad));
    alg.execute();
  });
}

/* MomentumEdgePecletAlg: writes the edge fields peclet_number, peclet_factor */
int
ref_run_peclet()
{
  return guarded([&] {
    auto& w = World::self();
    Realm realm;
    configure(realm);
    EquationSystem eq(w.ndim);
    stk::mesh::Part part;
    MomentumEdgePecletAlg alg(realm, &part, &eq);
    alg.execute();
  });
}

/* WallDistEdgeSolverAlg: lhsOut[nEdges][2][2], rhsOut[nEdges][2] */
int
ref_run_wall_dist(double* lhsOut, double* rhsOut)
{
  return guarded([&] {
    auto& w = World::self();
    w.lhsOut = lhsOut;
    w.rhsOut = rhsOut;
    Realm realm;
    configure(realm);
    EquationSystem eq(1);
    stk::mesh::Part part;
    WallDistEdgeSolverAlg alg(realm, &part, &eq);
    alg.execute();
  });
}

/* node kernels: which = 0 ScalarMassBDF(q), 1 MomentumMassBDF, 2 ContinuityMassBDF,
 * 3 WallDist; lhsOut[nNodes][n][n], rhsOut[nNodes][n] with n = numDof.  Time
 * states of a field F are the registered fields F, F_n, F_nm1. */
int
ref_run_node_kernel(int which, const char* q, double* lhsOut, double* rhsOut)
{
  return guarded([&] {
    auto& w = World::self();
    Realm realm;
    configure(realm);
    const int n = which == 1 ? w.ndim : 1;
    std::vector<double> l((size_t)n * n), r(n);
    NodeKernelTraits::LhsType lhs(l.data(), n, n);
    NodeKernelTraits::RhsType rhs(r.data(), n);
    auto run = [&](NodeKernel& k) {
      k.setup(realm);
      for (long i = 0; i < w.nNodes; ++i) {
        set_vals(rhs, 0.0);
        set_vals(lhs, 0.0);
        k.execute(lhs, rhs, stk::mesh::FastMeshIndex{0u, (unsigned)i});
        for (int j = 0; j < n * n; ++j)
          lhsOut[(size_t)i * n * n + j] = l[j];
        for (int j = 0; j < n; ++j)
          rhsOut[(size_t)i * n + j] = r[j];
      }
    };
    if (which == 0) {
      auto* f = static_cast<stk::mesh::Field<double>*>(
        w.fieldHandles.at(w.ordinal(q, stk::topology::NODE_RANK)));
      ScalarMassBDFNodeKernel k(realm.bulk_data(), f);
      run(k);
    } else if (which == 1) {
      MomentumMassBDFNodeKernel k(realm.bulk_data());
      run(k);
    } else if (which == 2) {
      ContinuityMassBDFNodeKernel k(realm.bulk_data());
      run(k);
    } else {
      WallDistNodeKernel k(realm.bulk_data());
      run(k);
    }
  });
}

} // extern "C"
