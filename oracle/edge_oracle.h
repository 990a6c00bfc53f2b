/*
 * edge_oracle.h -- CPU oracle for the nalu-wind edge-based CVFEM assembly path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nalu-wind_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker or the timed CPU baseline.
 *
 * It is a restatement (not a copy) of the reference's Kokkos lambdas with the
 * same per-edge operation order, the same loop shell and the same
 * column-walk + add scatter.  Every function cites the reference file:line it
 * follows (paths relative to /root/reference).
 *
 * Parity pin: checked against the reference's own unit-test golden vectors
 * (tests/test_oracle_golds.py): UnitTestMomentumAdvDiffEdge.C:19-227,
 * UnitTestContinuityAdvEdge.C:18-104, UnitTestScalarAdvDiffEdge.C:24-143 (serial
 * and 2-rank CSR), UnitTestMdotAlg.C:21-78, UnitTestNodalGradAlg.C:22-123.
 * The hypre IJ hand-off layout itself has no in-tree known-answer test (only
 * regression norms), so for that layout parity is "unpinned"; see DESIGN.md.
 *
 * Field layout everywhere in this file: the reference's, i.e. one array per
 * field, entity-major / component-minor ("AoS"): f[entity*ncomp + comp].
 */
#ifndef EDGE_ORACLE_H
#define EDGE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* threads used by the edge loops (default 1 == serial, deterministic) */
void orc_set_num_threads(int n);
int orc_get_num_threads(void);

/* ---- Peclet blending function (src/PecletFunction.C:41-45, 68-71) ---- */
enum { ORC_PECLET_CLASSIC = 0, ORC_PECLET_TANH = 1 };
typedef struct {
  int form;   /* ORC_PECLET_CLASSIC: hf*Pe ...; ORC_PECLET_TANH */
  double a;   /* classic: hybrid factor hf ; tanh: c1 (transition) */
  double b;   /* classic: unused          ; tanh: c2 (width)      */
} orc_peclet;

double orc_peclet_eval(const orc_peclet* f, double pecnum);
/* include/edge_kernels/EdgeKernelUtils.h:18-24 */
double orc_van_leer(double dqm, double dqp, double eps);

/* ---- linear-system sinks ("CoeffApplier", include/LinearSystem.h:44-74) ---- */
typedef struct orc_applier orc_applier;

/* dense test sink: unit_tests/UnitTestLinearSystem.h:42-72 (keeps only (d,d)
 * component pairs).  row index = node*numDof + d. */
orc_applier* orc_applier_dense_create(int64_t n_nodes, int num_dof);
/* copies out lhs (n*n row-major) and rhs (n), n = n_nodes*num_dof */
void orc_applier_dense_get(const orc_applier*, double* lhs, double* rhs);

/* recorder: keeps every local block in call order (single-threaded runs: edge
 * order); test infrastructure for the edge-by-edge comparison with the
 * reference's own lambdas (oracle/_ref) */
orc_applier* orc_applier_record_create(void);
int64_t orc_applier_record_count(const orc_applier*, int* n);
void orc_applier_record_get(const orc_applier*, double* lhs, double* rhs);

/* hypre-IJ style CSR graph: src/HypreLinearSystem.C:97-208 (begin),
 * :211-269 (fill), :412-478 (edge graph), :999-1236 (CSR build),
 * :883-993 (rows/cols arrays). */
typedef struct orc_graph orc_graph;
orc_graph* orc_graph_create(int num_dof, int64_t i_lower, int64_t i_upper);
void orc_graph_destroy(orc_graph*);
/* skippedRows_ (Dirichlet rows), global row ids (already *numDof if numDof>1) */
void orc_graph_set_skipped(orc_graph*, const int64_t* rows, int64_t n);
/* buildEdgeToNodeGraph over the locally-owned edges in the given order.
 * node_hid = per local node hypre id, periodic slaves already carrying their
 * master's id (get_entity_hypre_id, :2460-2470). */
void orc_graph_add_edges(
  orc_graph*, int64_t n_edges, const int32_t* edge_nodes,
  const int64_t* node_hid);
/* buildNodeGraph (:286-340): one (row,row) entry per owned node. */
void orc_graph_add_nodes(
  orc_graph*, int64_t n_nodes, const int32_t* nodes, const int64_t* node_hid);
/* finalizeLinearSystem: :819-880 */
void orc_graph_finalize(orc_graph*);

enum {
  ORC_G_NUM_ROWS_OWNED = 0,
  ORC_G_NNZ_OWNED = 1,
  ORC_G_NUM_ROWS_SHARED = 2,
  ORC_G_NNZ_SHARED = 3,
  ORC_G_NUM_PERIODIC_ROWS = 4
};
int64_t orc_graph_size(const orc_graph*, int what);
enum {
  ORC_G_ROW_START_OWNED = 0,  /* int64[numRowsOwned+1]   mat_row_start_owned_ */
  ORC_G_ROW_START_SHARED = 1, /* int64[numRowsShared+1]  mat_row_start_shared_ */
  ORC_G_COLS = 2,             /* int64[nnzOwned+nnzShared] cols_host_ */
  ORC_G_ROWS = 3,             /* int64[nnzOwned+nnzShared] rows_host_ */
  ORC_G_ROW_INDICES_SHARED = 4, /* int64[numRowsShared] ascending */
  ORC_G_PERIODIC_ROWS = 5     /* int64[numPeriodic] periodic_bc_rows_owned_ */
};
void orc_graph_copy(const orc_graph*, int what, int64_t* out);

/* HypreLinSysCoeffApplier (numDof==1: sum_into_1DoF :2165-2239; numDof>1:
 * sum_into :2059-2161) or, with uvw_ndim>0, HypreUVWLinSysCoeffApplier
 * (src/HypreUVWLinearSystem.C:695-767; graph must be numDof==1, rhs has
 * uvw_ndim columns). */
orc_applier* orc_applier_hypre_create(
  const orc_graph*, const int64_t* node_hid, int64_t n_nodes, int uvw_ndim);
/* resetCoeffApplierData (:1386-1430): zero + periodic rows diag 1 / rhs 0 */
void orc_applier_hypre_reset(orc_applier*);
/* |contribution| sums (the tests' tolerance scale) on / off; off for timing runs */
void orc_applier_hypre_track_abs(orc_applier*, int on);
/* values: [nnzOwned+nnzShared]; rhs: column-major [numRows][nrhs] where
 * numRows = owned+shared, nrhs = 1 or uvw_ndim (rhs_dev_(index,d)). */
void orc_applier_hypre_get(const orc_applier*, double* values, double* rhs);
/* per-call log of the value-array positions written, for the bit-exact
 * edge->slot map comparison: slots[(call*n + ii)*n + kk] (n = nEntities*numDof
 * or nEntities for UVW), rhs_index[call*n + ii]; -1 where the row was skipped. */
void orc_applier_hypre_enable_log(orc_applier*, int64_t n_calls);
void orc_applier_hypre_get_log(
  const orc_applier*, int64_t* slots, int64_t* rhs_index);
/* magnitude sums: same scatter of |lhs|,|rhs|, used as the per-entry
 * cancellation scale for the 1e-12 tolerance. */
void orc_applier_hypre_get_abs(const orc_applier*, double* values, double* rhs);

/* CoeffApplier::operator() (include/LinearSystem.h:62-70) over a batch:
 * lhs [n_entities][n][n] row-major, rhs [n_entities][n] */
void orc_applier_apply(
  orc_applier*, int64_t n_entities, int nodes_per_entity,
  const int32_t* entity_nodes, const double* lhs, const double* rhs, int n);
/* CoeffApplier::resetRows (src/HypreLinearSystem.C:2262-2330,
 * src/HypreUVWLinearSystem.C:787-851) */
void orc_applier_hypre_reset_rows(
  orc_applier*, int64_t n_nodes, const int32_t* nodes, double diag_value,
  double rhs_residual);
/* applyDirichletBCs (src/HypreLinearSystem.C:2407-2457,
 * src/HypreUVWLinearSystem.C:377-427) */
void orc_applier_hypre_dirichlet(
  orc_applier*, int64_t n_nodes, const int32_t* nodes, const double* solution,
  const double* bc_values, int ncomp);

/* node kernels through AssembleNGPNodeSolverAlgorithm
 * (src/AssembleNGPNodeSolverAlgorithm.C:85-146); nodes = selected local nodes
 * (locally owned, not periodic slaves); fields in the reference layout */
void orc_scalar_mass_bdf_node(
  int64_t n_sel, const int32_t* nodes, const double* qNm1, const double* qN,
  const double* qNp1, const double* rhoNm1, const double* rhoN,
  const double* rhoNp1, const double* dnvNm1, const double* dnvN,
  const double* dnvNp1, double dt, double gamma1, double gamma2, double gamma3,
  orc_applier*);
void orc_momentum_mass_bdf_node(
  int ndim, int64_t n_sel, const int32_t* nodes, const double* uNm1,
  const double* uN, const double* uNp1, const double* rhoNm1,
  const double* rhoN, const double* rhoNp1, const double* dnvNm1,
  const double* dnvN, const double* dnvNp1, const double* dpdx, double dt,
  double gamma1, double gamma2, double gamma3, orc_applier*);
void orc_continuity_mass_bdf_node(
  int64_t n_sel, const int32_t* nodes, const double* rhoNm1, const double* rhoN,
  const double* rhoNp1, const double* dnvNm1, const double* dnvN,
  const double* dnvNp1, double dt, double gamma1, double gamma2, double gamma3,
  orc_applier*);

/* src/edge_kernels/WallDistEdgeSolverAlg.C:28-66 and
 * src/node_kernels/WallDistNodeKernel.C:34-43 (Poisson system of the SST
 * minimum wall distance) */
void orc_wall_dist_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* edge_area, orc_applier* sink);
void orc_wall_dist_node(
  int64_t n_sel, const int32_t* nodes, const double* dual_nodal_volume,
  orc_applier*);

/* GeometryInteriorAlg<AlgTraitsHex8> (src/ngp_algorithms/GeometryInteriorAlg.C:72-112,
 * 165-225; Hex8CVFEM.C:365-390, 567-592; Hex8GeometryFunctions.h:33-325):
 * accumulates dual_nodal_volume [n_nodes], edge_area [n_edges][3] (may be NULL),
 * writes elem_volume [n_elems] (may be NULL) */
void orc_geometry_interior_hex8(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area);

/* GeometryInteriorAlg<AlgTraitsQuad4_2D> (src/master_element/Quad42DCVFEM.C:
 * 139-200, 384-445); 2-component coords / edge_area */
void orc_geometry_interior_quad4(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area);

/* GeometryInteriorAlg<AlgTraitsTet4 / Wed6 / Pyr5> (Tet4CVFEM.C:243-343, 522-619;
 * Wed6CVFEM.C:267-369, 541-647; Pyr5CVFEM.C:348-572, 772-900); conventions of
 * orc_geometry_interior_hex8 */
void orc_geometry_interior_tet4(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area);
void orc_geometry_interior_wed6(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area);
void orc_geometry_interior_pyr5(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area);

/* MdotEdgeAlg (sink == NULL) / ContinuityEdgeSolverAlg (sink != NULL) with the
 * optional balanced-buoyancy and GCL terms (MdotEdgeAlg.C:153-163, 175-180;
 * ContinuityEdgeSolverAlg.C:147-158, 172-177) */
void orc_mdot_continuity_edge_ext(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* gpdx, const double* density,
  const double* pressure, const double* udiag, const double* edge_area,
  double noc_fac, double interp_together, int add_balanced_forcing,
  const double* gravity, const double* source, const double* source_mask,
  int needs_gcl, const double* edge_face_vel_mag, double dt, double gamma1,
  double solve_incompressible, double* mdot, orc_applier* sink);

void orc_applier_destroy(orc_applier*);

/* ---- edge algorithms ---- */

/* src/ngp_algorithms/MdotEdgeAlg.C:117-190 (no buoyancy, no GCL) */
void orc_mdot_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* gpdx, const double* density,
  const double* pressure, const double* udiag, const double* edge_area,
  double noc_fac, double interp_together, double* mdot);

/* src/edge_kernels/MomentumEdgePecletAlg.C:74-101 */
void orc_peclet_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* vrtm, const double* density, const double* viscosity,
  const orc_peclet* pf, double eps, double* pecnum, double* pecfac);

/* src/ngp_algorithms/NodalGradEdgeAlg.C:85-109 ; grad is accumulated into
 * (NodalGradAlgDriver::pre_work zeroes it, :30-36). dim2 = ndim. */
void orc_nodal_grad_edge(
  int dim1, int dim2, int64_t n_edges, const int32_t* edge_nodes,
  const double* phi, const double* edge_area, const double* dual_vol,
  double* grad);

typedef struct {
  double dt, gamma1;
  double noc_fac;
  double interp_together;
  double solve_incompressible;
} orc_continuity_opts;
/* include/AssembleEdgeSolverAlgorithm.h:72-98 +
 * src/edge_kernels/ContinuityEdgeSolverAlg.C:109-194 */
void orc_continuity_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* gpdx, const double* density,
  const double* pressure, const double* udiag, const double* edge_area,
  const orc_continuity_opts* o, orc_applier* sink);

typedef struct {
  double alpha, alpha_upw, ho_upwind, relax_fac;
  int use_limiter;
  double eps; /* 1e-16 */
  orc_peclet pf;
} orc_scalar_opts;
/* src/edge_kernels/ScalarEdgeSolverAlg.C:85-205 */
void orc_scalar_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* vrtm, const double* q, const double* dqdx,
  const double* density, const double* diff_flux_coeff,
  const double* edge_area, const double* mdot, const orc_scalar_opts* o,
  orc_applier* sink);

typedef struct {
  double include_divu, alpha, alpha_upw, ho_upwind, relax_fac;
  int use_limiter;
  double eps; /* 1e-16 */
} orc_momentum_opts;
/* src/edge_kernels/MomentumEdgeSolverAlg.C:105-312 (has_vof = 0;
 * orc_momentum_edge_vof below: has_vof = 1).
 * udiag_accum (nullable): NGPApplyCoeff::extract_diagonal,
 * src/SolverAlgorithm.C:87-105, adds lhs(i*ndim,i*ndim) per end node. */
void orc_momentum_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* dudx, const double* viscosity,
  const double* density, const double* node_mask, const double* edge_area,
  const double* mdot, const double* pecfac, const orc_momentum_opts* o,
  orc_applier* sink, double* udiag_accum);
/* the same with solutionOptions_->realm_has_vof_ set (:88, 124-125, 174-192):
 * mdot + mass_vof_balanced_flow_rate, density-jump upwinding factors */
void orc_momentum_edge_vof(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* dudx, const double* viscosity,
  const double* density, const double* node_mask, const double* edge_area,
  const double* mdot, const double* mass_vof_balanced, const double* pecfac,
  const orc_momentum_opts* o, orc_applier* sink, double* udiag_accum);

#ifdef __cplusplus
}
#endif
#endif
