/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points over a handful of the REFERENCE'S OWN source files, compiled
 * unmodified from where they lie under /root/reference (recipe:
 * oracle/Makefile.ref, output: oracle/_ref/libnalu_ref.so, git-ignored):
 *
 *   src/master_element/{Hex8,Tet4,Pyr5,Wed6,Quad42D,Tri32D}CVFEM.C, MasterElement.C
 *       -> SCV volumes, SCS area vectors, ipNodeMap, adjacentNodes: what
 *          GeometryInteriorAlg (src/ngp_algorithms/GeometryInteriorAlg.C:72-112,
 *          165-225) integrates into dual_nodal_volume / edge_area_vector
 *   src/PecletFunction.C            -> ClassicPecletFunction / TanhFunction
 *   include/edge_kernels/EdgeKernelUtils.h -> van_leer
 *
 * Kokkos, STK and MPI are not installed here; oracle/ref_shim/ stands in for
 * the few names of theirs these files touch (serial host semantics, written
 * for this repo).  The edge kernels themselves (the execute() lambdas) are
 * bound to Realm / STK fields and cannot be built this way -- for them the
 * oracle stays a restatement pinned by the reference's golds.
 *
 * Used by tests/golden/extract_reference_runs.py to produce
 * tests/golden/reference_runs.json, and by tests/test_reference_runs.py when
 * the library is present.  Nothing under nalu-wind_b200/ links or loads it.
 */
#include <master_element/MasterElement.h>
#include <master_element/Hex8CVFEM.h>
#include <master_element/Tet4CVFEM.h>
#include <master_element/Pyr5CVFEM.h>
#include <master_element/Wed6CVFEM.h>
#include <master_element/Quad42DCVFEM.h>
#include <master_element/Tri32DCVFEM.h>
#include <PecletFunction.h>
#include <edge_kernels/EdgeKernelUtils.h>
#include <NaluEnv.h>

#include <iostream>
#include <memory>
#include <vector>

using namespace sierra::nalu;

namespace {

enum Topo { HEX8 = 0, TET4 = 1, PYR5 = 2, WED6 = 3, QUAD4_2D = 4, TRI3_2D = 5 };

std::unique_ptr<MasterElement>
make_scv(int t)
{
  switch (t) {
  case HEX8: return std::make_unique<HexSCV>();
  case TET4: return std::make_unique<TetSCV>();
  case PYR5: return std::make_unique<PyrSCV>();
  case WED6: return std::make_unique<WedSCV>();
  case QUAD4_2D: return std::make_unique<Quad42DSCV>();
  case TRI3_2D: return std::make_unique<Tri32DSCV>();
  }
  return nullptr;
}

std::unique_ptr<MasterElement>
make_scs(int t)
{
  switch (t) {
  case HEX8: return std::make_unique<HexSCS>();
  case TET4: return std::make_unique<TetSCS>();
  case PYR5: return std::make_unique<PyrSCS>();
  case WED6: return std::make_unique<WedSCS>();
  case QUAD4_2D: return std::make_unique<Quad42DSCS>();
  case TRI3_2D: return std::make_unique<Tri32DSCS>();
  }
  return nullptr;
}

} // namespace

extern "C" {

/* sizes: out[0..4] = ndim, nodesPerElement, numScvIp, numScsIp */
int
ref_me_sizes(int topo, int* out)
{
  auto v = make_scv(topo);
  auto s = make_scs(topo);
  if (!v || !s)
    return 1;
  out[0] = v->nDim_;
  out[1] = v->nodesPerElement_;
  out[2] = v->num_integration_points();
  out[3] = s->num_integration_points();
  return 0;
}

/* SCV: volume[numScvIp] of one element, coords[npe][ndim]; simd != 0 takes the
 * DoubleType overload GeometryInteriorAlg uses (one lane here) */
int
ref_scv_volume(int topo, const double* coords, int simd, double* volume)
{
  auto me = make_scv(topo);
  if (!me)
    return 1;
  const int npe = me->nodesPerElement_, nd = me->nDim_;
  const int nip = me->num_integration_points();
  try {
    if (!simd) {
      std::vector<double> c(coords, coords + npe * nd);
      SharedMemView<double**> cv(c.data(), npe, nd);
      SharedMemView<double*> vv(volume, nip);
      me->determinant(cv, vv);
    } else {
      std::vector<DoubleType> c(npe * nd), v(nip);
      for (int i = 0; i < npe * nd; ++i)
        c[i] = coords[i];
      SharedMemView<DoubleType**, DeviceShmem> cv(c.data(), npe, nd);
      SharedMemView<DoubleType*, DeviceShmem> vv(v.data(), nip);
      me->determinant(cv, vv);
      for (int i = 0; i < nip; ++i)
        volume[i] = stk::simd::get_data(v[i], 0);
    }
  } catch (const std::exception& e) {
    std::cerr << "ref_scv_volume: " << e.what() << "\n";
    return 2;
  }
  return 0;
}

/* SCS: areav[numScsIp][ndim] */
int
ref_scs_areav(int topo, const double* coords, int simd, double* areav)
{
  auto me = make_scs(topo);
  if (!me)
    return 1;
  const int npe = me->nodesPerElement_, nd = me->nDim_;
  const int nip = me->num_integration_points();
  try {
    if (!simd) {
      std::vector<double> c(coords, coords + npe * nd);
      SharedMemView<double**> cv(c.data(), npe, nd);
      SharedMemView<double**> av(areav, nip, nd);
      me->determinant(cv, av);
    } else {
      std::vector<DoubleType> c(npe * nd), a(nip * nd);
      for (int i = 0; i < npe * nd; ++i)
        c[i] = coords[i];
      SharedMemView<DoubleType**, DeviceShmem> cv(c.data(), npe, nd);
      SharedMemView<DoubleType**, DeviceShmem> av(a.data(), nip, nd);
      me->determinant(cv, av);
      for (int i = 0; i < nip * nd; ++i)
        areav[i] = stk::simd::get_data(a[i], 0);
    }
  } catch (const std::exception& e) {
    std::cerr << "ref_scs_areav: " << e.what() << "\n";
    return 2;
  }
  return 0;
}

/* one element block of a mesh: conn[nElems][npe] into coords[nNodes][ndim];
 * volume[nElems][numScvIp], areav[nElems][numScsIp][ndim] (DoubleType overloads) */
int
ref_geometry_block(
  int topo, long nElems, const int* conn, const double* coords, double* volume,
  double* areav)
{
  auto v = make_scv(topo);
  auto s = make_scs(topo);
  if (!v || !s)
    return 1;
  const int npe = v->nodesPerElement_, nd = v->nDim_;
  const int nscv = v->num_integration_points();
  const int nscs = s->num_integration_points();
  std::vector<double> x(npe * nd);
  for (long e = 0; e < nElems; ++e) {
    for (int n = 0; n < npe; ++n)
      for (int d = 0; d < nd; ++d)
        x[n * nd + d] = coords[(long)conn[e * npe + n] * nd + d];
    if (int rc = ref_scv_volume(topo, x.data(), 1, volume + e * nscv))
      return rc;
    if (int rc = ref_scs_areav(topo, x.data(), 1, areav + e * nscs * nd))
      return rc;
  }
  return 0;
}

/* ipNodeMap of the SCV (numScvIp ints) and adjacentNodes of the SCS
 * (2 numScsIp ints: left, right) */
int
ref_me_maps(int topo, int* scvIpNode, int* scsLr)
{
  auto v = make_scv(topo);
  auto s = make_scs(topo);
  if (!v || !s)
    return 1;
  const int* m = v->ipNodeMap();
  for (int i = 0; i < v->num_integration_points(); ++i)
    scvIpNode[i] = m[i];
  const int* lr = s->adjacentNodes();
  for (int i = 0; i < 2 * s->num_integration_points(); ++i)
    scsLr[i] = lr[i];
  return 0;
}

double
ref_peclet_classic(double A, double hf, double pec)
{
  ClassicPecletFunction<double> f(A, hf);
  return f.execute(pec);
}

double
ref_peclet_tanh(double c1, double c2, double pec)
{
  TanhFunction<double> f(c1, c2);
  return f.execute(pec);
}

double
ref_van_leer(double dqm, double dqp, double eps)
{
  return van_leer(dqm, dqp, eps);
}

} // extern "C"

/* NaluEnv is declared by the reference (include/NaluEnv.h); its definition
 * (src/NaluEnv.C) needs MPI.  The master elements only use it to print. */
/* rank / size as the stand-in world of the other drivers sets them */
int g_nwref_rank = 0, g_nwref_size = 1;
static int nwref_rank() { return g_nwref_rank; }
static int nwref_size() { return g_nwref_size; }

namespace sierra {
namespace nalu {
NaluEnv::NaluEnv() : parallelCommunicator_(0), pSize_(1), pRank_(0) {}
NaluEnv::~NaluEnv() {}
NaluEnv&
NaluEnv::self()
{
  static NaluEnv e;
  return e;
}
namespace {
/* swallows the reference's log lines (e.g. the VOF warning of the
 * MomentumEdgeSolverAlg constructor) */
struct NullBuf : std::streambuf
{
  int overflow(int c) override { return c; }
};
std::ostream&
null_stream()
{
  static NullBuf b;
  static std::ostream s(&b);
  return s;
}
} // namespace
int NaluEnv::parallel_rank() { return nwref_rank(); }
int NaluEnv::parallel_size() { return nwref_size(); }
MPI_Comm NaluEnv::parallel_comm() { return 0; }
double NaluEnv::nalu_time() { return 0.0; }
std::ostream& NaluEnv::naluOutputP0() { return null_stream(); }
std::ostream& NaluEnv::naluOutput() { return null_stream(); }
} // namespace nalu
} // namespace sierra
