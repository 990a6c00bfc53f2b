/*
 * shim_nodes.cpp -- reads like unit_tests/node_kernels/UnitTestScalarMassBDFNodeKernel.C,
 * unit_tests/edge_kernels/UnitTestWallDistEdgeSolver.C and
 * unit_tests/ngp_algorithms/UnitTestGeometryAlg.C: the one-element hex8 mesh,
 * through the reference-named C++ classes of nalu-wind_b200/host/NaluEdgeB200.h
 * (GeometryAlgDriver, AssembleNGPNodeSolverAlgorithm::add_kernel<...>,
 * WallDistEdgeSolverAlg).  Prints what the Python test compares with the
 * reference's golden values.
 * usage: shim_nodes <cuda device>
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "NaluEdgeB200.h"

using namespace sierra::nalu;

static void
print_system(const char* tag, LinearSystem& linsys)
{
  std::vector<double> values, rhs;
  linsys.copy_values(values, rhs);
  const nw_linsys_sizes s = linsys.sizes();
  std::vector<int64_t> rows(values.size()), cols(values.size());
  nw_check(nw_linsys_get_graph(
    linsys.handle(), nullptr, nullptr, cols.data(), rows.data(), nullptr,
    nullptr));
  std::printf("%s_rhs", tag);
  for (int i = 0; i < s.num_rows_owned; ++i)
    std::printf(" %.17g", rhs[i]);
  std::printf("\n");
  for (size_t k = 0; k < values.size(); ++k)
    std::printf("%s_lhs %d %d %.17g\n", tag, (int)rows[k], (int)cols[k], values[k]);
}

int
main(int argc, char** argv)
{
  const int device = argc > 1 ? std::atoi(argv[1]) : 0;
  const double pi = std::acos(-1.0);
  std::vector<double> coords(24);
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        const int n = i + 2 * j + 4 * k;
        coords[3 * n] = i;
        coords[3 * n + 1] = j;
        coords[3 * n + 2] = k;
      }
  const int hexEdges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6},
                               {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  const int l2id[8] = {0, 1, 3, 2, 4, 5, 7, 6};
  std::vector<int32_t> edges;
  for (int e = 0; e < 12; ++e) {
    int na = l2id[hexEdges[e][0]], nb = l2id[hexEdges[e][1]];
    if (na > nb)
      std::swap(na, nb);
    edges.push_back(na);
    edges.push_back(nb);
  }
  std::vector<int64_t> hid = {0, 1, 2, 3, 4, 5, 6, 7};
  const int64_t offsets[2] = {0, 8};
  nw_mesh_desc d = {};
  d.ndim = 3;
  d.nranks = 1;
  d.n_nodes = 8;
  d.n_edges = 12;
  d.edge_nodes = edges.data();
  d.node_hypre_id = hid.data();
  d.hypre_offsets = offsets;
  d.coords = coords.data();
  d.tile_nodes = 8;
  try {
    Realm realm(device, d);
    /* TimeIntegrator: timeStepN_ = 0.1, gamma = 1, -1, 0 */
    realm.set_bdf(0.1, 1.0, -1.0, 0.0);
    /* MixtureFractionKernelHex8Mesh fields: only StateNP1 is initialised */
    std::vector<double> z(8), rho(8), zero(8, 0.0);
    for (int n = 0; n < 8; ++n) {
      z[n] = 2.0 * std::cos(pi * coords[3 * n]) * std::cos(pi * coords[3 * n + 1]) *
             std::cos(pi * coords[3 * n + 2]);
      rho[n] = 1.0 / (z[n] / 0.163 + (1.0 - z[n]) / 1.18);
    }
    for (const char* f : {"mixture_fraction", "mixture_fraction_n", "density",
                          "density_n", "dual_nodal_volume"})
      realm.register_field(f, NW_NODE, 1);
    realm.register_field("edge_area_vector", NW_EDGE, 3);
    realm.upload("mixture_fraction", z.data());
    realm.upload("mixture_fraction_n", zero.data());
    realm.upload("density", rho.data());
    realm.upload("density_n", zero.data());

    /* GeometryAlgDriver: dual nodal volumes and edge area vectors on the device */
    GeometryAlgDriver geom(realm);
    geom.register_elem_block(8, std::vector<int32_t>(l2id, l2id + 8));
    geom.execute();
    std::vector<double> dnv(8), area(36);
    realm.download("dual_nodal_volume", dnv.data());
    realm.download("edge_area_vector", area.data());
    std::printf("dnv");
    for (double v : dnv)
      std::printf(" %.17g", v);
    std::printf("\narea");
    for (double v : area)
      std::printf(" %.17g", v);
    std::printf("\n");

    /* NGP_scalar_mass_node */
    {
      HypreLinearSystem linsys(realm, 1);
      EquationSystem eqSys(realm, "MixtureFractionEQS");
      eqSys.linsys_ = &linsys;
      linsys.buildEdgeToNodeGraph();
      linsys.finalizeLinearSystem();
      AssembleNGPNodeSolverAlgorithm nodeAlg(realm, &eqSys);
      nodeAlg.add_kernel<ScalarMassBDFNodeKernel>("mixture_fraction");
      linsys.zeroSystem();
      nodeAlg.execute();
      print_system("mass", linsys);
    }
    /* NGP_wall_dist_edge + WallDistNodeKernel */
    {
      HypreLinearSystem linsys(realm, 1);
      EquationSystem eqSys(realm, "WallDistEQS");
      eqSys.linsys_ = &linsys;
      WallDistEdgeSolverAlg alg(realm, &eqSys);
      alg.initialize_connectivity();
      linsys.finalizeLinearSystem();
      AssembleNGPNodeSolverAlgorithm nodeAlg(realm, &eqSys);
      nodeAlg.add_kernel<WallDistNodeKernel>();
      linsys.zeroSystem();
      alg.execute();
      nodeAlg.execute();
      print_system("wdist", linsys);
    }
  } catch (const std::exception& ex) {
    std::printf("error: %s\n", ex.what());
    return 1;
  }
  return 0;
}
