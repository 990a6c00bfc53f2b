/*
 * shim_continuity.cpp -- reads like unit_tests/edge_kernels/UnitTestContinuityAdvEdge.C:
 * one-element hex8 mesh, trig fields (a = 0.3), dt = gamma1 = 1, assemble the
 * continuity edge system through the reference-named C++ classes
 * (nalu-wind_b200/host/NaluEdgeB200.h) and print rhs + dense lhs so that the
 * Python test can compare with the reference's golden values.
 * usage: shim_continuity <cuda device | -1>
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "NaluEdgeB200.h"

using namespace sierra::nalu;

int
main(int argc, char** argv)
{
  const int device = argc > 1 ? std::atoi(argv[1]) : 0;
  const double a = 0.3, pi = std::acos(-1.0);
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
  std::vector<double> area(36, 0.0);
  for (int e = 0; e < 12; ++e) {
    int na = l2id[hexEdges[e][0]], nb = l2id[hexEdges[e][1]];
    if (na > nb)
      std::swap(na, nb);
    edges.push_back(na);
    edges.push_back(nb);
    for (int d = 0; d < 3; ++d)
      if (coords[3 * nb + d] != coords[3 * na + d])
        area[3 * e + d] = 0.25;
  }
  std::vector<int64_t> hid = {0, 1, 2, 3, 4, 5, 6, 7};
  const int64_t offsets[2] = {0, 8};
  nw_mesh_desc d = {};
  d.ndim = 3;
  d.rank = 0;
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
    realm.set_time_step(1.0, 1.0);
    realm.solutionOptions_.mdotInterpRhoUTogether_ = true;
    std::vector<double> vel(24, 0.0), dpdx(24, 0.0), p(8), rho(8, 1.0),
      udiag(8, 1.0);
    for (int n = 0; n < 8; ++n) {
      const double x = coords[3 * n], y = coords[3 * n + 1];
      vel[3 * n] = -std::cos(a * pi * x) * std::sin(a * pi * y);
      vel[3 * n + 1] = std::sin(a * pi * x) * std::cos(a * pi * y);
      p[n] = -0.25 * (std::cos(2 * a * pi * x) + std::cos(2 * a * pi * y));
      dpdx[3 * n] = 0.5 * a * pi * std::sin(2 * a * pi * x);
      dpdx[3 * n + 1] = 0.5 * a * pi * std::sin(2 * a * pi * y);
    }
    realm.register_field("velocity", NW_NODE, 3);
    realm.register_field("dpdx", NW_NODE, 3);
    realm.register_field("pressure", NW_NODE, 1);
    realm.register_field("density", NW_NODE, 1);
    realm.register_field("momentum_diag", NW_NODE, 1);
    realm.register_field("edge_area_vector", NW_EDGE, 3);
    HypreLinearSystem linsys(realm, 1);
    EquationSystem eqSys(realm, "ContinuityEQS");
    eqSys.linsys_ = &linsys;
    ContinuityEdgeSolverAlg alg(realm, &eqSys);
    alg.initialize_connectivity();
    linsys.finalizeLinearSystem();
    if (device < 0) {
      /* host-only context: graph only; compute must refuse */
      try {
        alg.execute();
        std::printf("UNEXPECTED: execute succeeded without a device\n");
        return 2;
      } catch (const std::runtime_error& ex) {
        std::printf("refused: %s\n", ex.what());
        return 0;
      }
    }
    realm.upload("velocity", vel.data());
    realm.upload("dpdx", dpdx.data());
    realm.upload("pressure", p.data());
    realm.upload("density", rho.data());
    realm.upload("momentum_diag", udiag.data());
    realm.upload("edge_area_vector", area.data());
    realm.sync();
    linsys.zeroSystem();
    alg.execute();
    linsys.loadComplete();
    std::vector<double> values, rhs;
    linsys.copy_values(values, rhs);
    const nw_linsys_sizes s = linsys.sizes();
    std::vector<int64_t> rows(values.size()), cols(values.size());
    nw_check(nw_linsys_get_graph(
      linsys.handle(), nullptr, nullptr, cols.data(), rows.data(), nullptr,
      nullptr));
    std::printf("rhs");
    for (int i = 0; i < s.num_rows_owned; ++i)
      std::printf(" %.17g", rhs[i]);
    std::printf("\n");
    for (size_t k = 0; k < values.size(); ++k)
      std::printf("lhs %d %d %.17g\n", (int)rows[k], (int)cols[k], values[k]);
  } catch (const std::exception& ex) {
    std::printf("error: %s\n", ex.what());
    return 1;
  }
  return 0;
}
