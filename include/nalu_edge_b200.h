/*
 * nalu_edge_b200.h -- C ABI of the B200-native edge-based CVFEM assembly path.
 *
 * Drop-in boundary for the hot path of Exawind/nalu-wind (citations relative to
 * the nalu-wind source tree): the edge algorithms
 *   MdotEdgeAlg, NodalGradEdgeAlg, MomentumEdgePecletAlg,
 *   ContinuityEdgeSolverAlg, ScalarEdgeSolverAlg, MomentumEdgeSolverAlg
 * and the LinearSystem / CoeffApplier assembly surface of HypreLinearSystem and
 * HypreUVWLinearSystem.  Plain C, opaque handles, plain pointers and sizes; no
 * C++/torch types.  The C++ shim classes that keep the reference's names live
 * in nalu-wind_b200/host/ and call only this header.
 *
 * Conventions
 *  - every function returns nw_status (0 == NW_OK); nw_last_error() gives the
 *    message of the last failure on the calling thread (the reference throws
 *    std::runtime_error / STK_ThrowRequire on the host and never reports device
 *    errors; here host-side misuse is an error code, never a silent fallback).
 *  - there is NO CPU fallback: every compute entry point fails with
 *    NW_ERR_CUDA if no CUDA device is usable.
 *  - host arrays use the reference's field layout: entity-major,
 *    component-minor (f[entity*ncomp + comp]); nodes are indexed by the
 *    caller's local node index (0..n_nodes), edges by the caller's edge order
 *    (the order of the STK edge buckets the reference loops over).
 *  - one CUDA stream per context; calls are asynchronous on that stream unless
 *    they return data to the host.  No host-side concurrency on one context
 *    (same rule as the reference's LinearSystem).
 */
#ifndef NALU_EDGE_B200_H
#define NALU_EDGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NW_OK = 0,
  NW_ERR_ARG = 1,   /* bad argument / inconsistent sizes */
  NW_ERR_CUDA = 2,  /* CUDA runtime failure or no device */
  NW_ERR_STATE = 3, /* call out of order (e.g. assemble before finalize) */
  NW_ERR_COMM = 4,  /* NCCL failure / communicator missing */
  NW_ERR_LIMIT = 5  /* a plan limit was exceeded */
} nw_status;

typedef struct nw_ctx nw_ctx;       /* device + stream + communicator */
typedef struct nw_mesh nw_mesh;     /* one rank's mesh partition + fields */
typedef struct nw_linsys nw_linsys; /* LinearSystem: graph + values */

const char* nw_last_error(void);
/* library / ABI version: major*1000 + minor */
int nw_version(void);
/* Diagnostics (no reference counterpart; the reference relies on Kokkos
 * tools): per-phase cycle counters of the tile kernels, out[8][12], filled
 * only by the NW_PHASE_TIMING build of the library (zeros otherwise). */
int nw_debug_phase_times(unsigned long long* out, int n, int reset);
/* Measurement aid (no reference counterpart): while on, nw_linsys_load_complete,
 * nw_nodal_grad_edge's shared-node sum and nw_field_parallel_sum /
 * nw_field_copy_owned_to_shared return without exchanging anything, so that a
 * bench can time the same sweep with and without its halo exchanges
 * (bench.py north_star.exchange_share).  Results of a multi-rank assembly are
 * INCOMPLETE while it is on; never set it in a solver. */
int nw_debug_skip_exchange(nw_ctx* ctx, int on);

/* ------------------------------------------------------------------ */
/* context                                                             */
/* ------------------------------------------------------------------ */

/* replaces Kokkos::initialize + hypre_initialize for this path (nalu.C:83-86) */
int nw_ctx_create(int cuda_device, nw_ctx** out);
int nw_ctx_destroy(nw_ctx* ctx);
/* fence: Kokkos::fence() equivalent on the context's stream */
int nw_ctx_sync(nw_ctx* ctx);
/* cudaStream_t of the context (as void*) so callers can order their own work */
void* nw_ctx_stream(nw_ctx* ctx);

/* Multi-GPU: one process per GPU.  The communicator replaces the MPI
 * communicator STK/hypre use (stk::mesh::parallel_sum,
 * HYPRE_IJMatrixAssemble).  unique_id is the 128-byte ncclUniqueId, produced
 * on rank 0 by nw_comm_unique_id and distributed by the caller (MPI_Bcast,
 * torch.distributed, a file ...). */
#define NW_UNIQUE_ID_BYTES 128
int nw_comm_unique_id(void* unique_id_out);
int nw_ctx_comm_init(nw_ctx* ctx, const void* unique_id, int nranks, int rank);
/* Transport of the two halo exchanges (shared-row add of loadComplete,
 * shared-node sum of the nodal gradient).  nw_ctx_comm_init also tries to set
 * up a peer-memory mailbox over NVLink (CUDA IPC windows; NW_P2P=0 disables
 * it): the sender stores straight into the receiver's window -- from inside
 * the producing kernel (gradient kernels; edge assemblies after
 * nw_linsys_set_eager_exchange) or from a push kernel -- and a pull kernel
 * publishes / waits for the epoch flags and does the ordered add, instead of
 * pack / ncclSend / ncclRecv / unpack.  All ranks use the same transport
 * (agreed collectively); results are bit-identical between the two.  Any
 * decomposition works: every exchange signals and waits for the union of the
 * neighbour ranks of all exchange objects of the context.  A neighbour that
 * does not arrive within NW_P2P_TIMEOUT_S (default 20 s) makes the next call
 * that returns data to the host fail with NW_ERR_COMM; nothing stale is ever
 * added. */
typedef enum {
  NW_TRANSPORT_NONE = 0,       /* single rank / structure not built yet */
  NW_TRANSPORT_NCCL = 1,
  NW_TRANSPORT_PEER_MEMORY = 2
} nw_halo_transport;
int nw_ctx_peer_memory(const nw_ctx* ctx); /* 1 if the mailbox is up */
/* The receiving half of a peer-memory exchange (wait for the neighbours +
 * add) runs on a separate high-priority communication stream, so that it
 * overlaps whatever the caller enqueues next (NW_P2P_ASYNC=0: on the compute
 * stream); every nw_* call that touches the exchanged field / linear system
 * again -- nw_field_device_view and nw_linsys_device_arrays included -- is
 * ordered after it automatically.  nw_ctx_join_comm orders the context's
 * compute stream after ALL exchanges issued so far (no host
 * synchronisation); nw_ctx_sync waits for both streams. */
int nw_ctx_join_comm(nw_ctx* ctx);

/* ------------------------------------------------------------------ */
/* mesh partition                                                      */
/* ------------------------------------------------------------------ */

typedef struct {
  int32_t ndim;   /* 2 or 3 (meta.spatial_dimension()) */
  int32_t rank;   /* bulk.parallel_rank() */
  int32_t nranks; /* bulk.parallel_size() */
  int64_t n_nodes; /* local nodes: owned + shared (+ periodic slaves) */
  int64_t n_edges; /* locally-owned edges: the selector of
                      AssembleEdgeSolverAlgorithm.h:56-58 */
  /* [2*n_edges] (nodeL, nodeR) local node indices, in the reference's edge
   * order and orientation (STK: nodes of an edge in ascending global id) */
  const int32_t* edge_nodes;
  /* [n_nodes] Realm::hypreGlobalId_ with periodic slaves already carrying the
   * master's id (HypreLinearSystem::get_entity_hypre_id,
   * src/HypreLinearSystem.C:2460-2470) */
  const int64_t* node_hypre_id;
  /* [n_nodes] the node's own row id before periodic resolution (only differs
   * from node_hypre_id for periodic slaves); NULL == same as node_hypre_id */
  const int64_t* node_own_hypre_id;
  /* [nranks+1] Realm::hypreOffsets_ (src/Realm.C:3620-3626): rank r owns rows
   * [offsets[r], offsets[r+1]) */
  const int64_t* hypre_offsets;
  /* [n_nodes*ndim] coordinates; used for the locality tiling and registered
   * as the nodal field "coordinates" */
  const double* coords;
  /* target nodes per tile; 0 = library default */
  int32_t tile_nodes;
} nw_mesh_desc;

int nw_mesh_create(nw_ctx* ctx, const nw_mesh_desc* desc, nw_mesh** out);
int nw_mesh_destroy(nw_mesh* mesh);

typedef struct {
  int64_t n_nodes, n_edges;
  int64_t n_tiles;
  int64_t n_tile_edges;     /* edges incl. the copies cut edges get per tile */
  int64_t n_halo_nodes;     /* sum over tiles of staged non-owned nodes */
  int64_t max_tile_nodes, max_tile_staged, max_tile_edges, max_tile_halfedges;
  int64_t plan_bytes_device; /* integer plan arrays resident in HBM */
} nw_mesh_stats;
int nw_mesh_get_stats(const nw_mesh* mesh, nw_mesh_stats* out);

/* ---- fields (replaces the stk::mesh::NgpField the lambdas capture) ---- */

typedef enum { NW_NODE = 0, NW_EDGE = 1 } nw_entity_rank;

/* Registers (or finds) a field by the reference's field name, e.g.
 * "velocity", "pressure", "density", "dpdx", "dudx", "momentum_diag",
 * "viscosity", "dual_nodal_volume", "edge_area_vector", "mass_flow_rate",
 * "peclet_factor", "abl_wall_no_slip_wall_func_node_mask", "turbulent_ke" ...
 * Storage is zero-initialised. */
int nw_field_register(
  nw_mesh* mesh, const char* name, int entity_rank, int ncomp, int* field_id);
int nw_field_find(const nw_mesh* mesh, const char* name, int* field_id);
/* host (pageable or pinned) -> device, reference layout; asynchronous on the
 * context stream when the host buffer is pinned */
int nw_field_upload(nw_mesh* mesh, int field_id, const double* host);
/* device -> host, reference layout; synchronises the stream */
int nw_field_download(nw_mesh* mesh, int field_id, double* host);
/* Pipelined form of nw_field_upload for a caller that feeds new nodal state
 * every step from pinned host memory (the ngp_field sync_to_device of the
 * reference, include/ngp_utils/NgpFieldManager.h, stk::mesh::NgpField::
 * sync_to_device): nw_field_stage enqueues the host -> device copy on the
 * context's copy stream into a staging buffer owned by the field and returns
 * at once, so the copy overlaps whatever the compute stream is doing;
 * nw_field_commit makes the compute stream wait for that copy and permutes the
 * staged data into the field (internal SoA layout).  A second nw_field_stage
 * on the same field waits (on the copy stream) until the previous commit has
 * consumed the staging buffer.  `host` must stay valid until the matching
 * commit has been issued and must be pinned for the copy to be asynchronous. */
int nw_field_stage(nw_mesh* mesh, int field_id, const double* host);
int nw_field_commit(nw_mesh* mesh, int field_id);
/* fill every component with a constant (stk::mesh::field_fill) */
int nw_field_fill(nw_mesh* mesh, int field_id, double value);
/* Device view of the internal storage: component c of internal entity i is at
 * base[c*stride + i]; internal numbering via nw_mesh_get_node_permutation. */
int nw_field_device_view(
  nw_mesh* mesh, int field_id, double** base, int64_t* stride);
/* perm[i] = caller's local node index of internal slot i, or -1 for a padding
 * slot; n_slots >= n_nodes */
int nw_mesh_get_node_permutation(
  const nw_mesh* mesh, int64_t* n_slots, int32_t* perm /* may be NULL */);


/* ---- shared-node exchange lists (multi-rank) ----
 * A node is identified across ranks by its own hypre row id.  With a
 * communicator attached (nw_ctx_comm_init before nw_mesh_create) the lists are
 * exchanged automatically; otherwise the caller moves them with its own
 * transport: for every peer, get_send on the sharer -> set_recv on the owner,
 * then commit on every rank.  Replaces the STK comm lists behind
 * stk::mesh::parallel_sum / copy_owned_to_shared. */
int nw_mesh_halo_send_count(const nw_mesh* mesh, int peer, int64_t* n);
int nw_mesh_halo_get_send(const nw_mesh* mesh, int peer, int64_t* own_hids);
int nw_mesh_halo_set_recv(
  nw_mesh* mesh, int peer, int64_t n, const int64_t* own_hids);
int nw_mesh_halo_commit(nw_mesh* mesh);
/* stk::mesh::parallel_sum(bulk, {field}) for a nodal field: every sharer ends
 * with the sum over all ranks' copies (owner adds in ascending rank order,
 * then owner -> sharers).  Single rank: no-op. */
int nw_field_parallel_sum(nw_mesh* mesh, int field_id);
/* Realm::periodic_field_update(field, size) (src/Realm.C:3090-3100 ->
 * PeriodicManager::apply_constraints with addSlaves and setSlaves,
 * src/PeriodicManager.C:1007-1058, 1139-1190) for a nodal field: the nodes that
 * resolve to one row (node_hypre_id equal: a periodic master and its slaves)
 * end with master + sum of the slaves -- added in ascending own-id order --
 * on every copy.  No-op on a mesh without periodic aliases.  A master that
 * lives on another rank than its slaves is NW_ERR_LIMIT (the slab / block
 * decompositions of the periodic decks keep them together). */
int nw_field_periodic_update(nw_mesh* mesh, int field_id);
/* stk::mesh::copy_owned_to_shared(bulk, {field}) for a nodal field (what the
 * reference does to a solution field after every solve, src/LinearSystem.C:
 * 161-169, so that the next sweep reads current values on shared nodes): every
 * non-owned copy takes its owner's value.  Single rank: no-op. */
int nw_field_copy_owned_to_shared(nw_mesh* mesh, int field_id);
/* The consumer of extract_diagonal: the part of
 * MomentumEquationSystem::assemble_and_solve that follows the momentum
 * assembly when projected_timescale_type is momentum_diag_inv
 * (src/LowMachEquationSystem.C:2759-2821).  In the reference's order:
 * stk::mesh::parallel_sum of the shared-node copies of udiag (:2770-2774);
 * on every locally owned node that is not a periodic slave
 *     udiag = (udiag / (density * dual_nodal_volume) - gamma1/dt) * alpha_u
 *             + gamma1/dt                                     (:2776-2790)
 * (alpha_u: the velocity relaxation factor the assembly divided the diagonal
 * by); copy_owned_to_shared (:2795-2800); periodic slaves take their master's
 * value (apply_constraints with setSlaves only, :2802-2808).  All on the
 * device, no host round trip; the non-conformal / overset updates of :2809-2817
 * are out of scope.  The caller resets udiag before the assembly
 * (nw_field_fill, :2741-2751) and passes it as nw_momentum_opts.diag_field. */
int nw_momentum_diag_post_process(
  nw_mesh* mesh, int udiag_field, int density_field,
  int dual_nodal_volume_field, double dt, double gamma1, double alpha_u);
/* transport the nodal halo sum of this mesh uses (nw_halo_transport) */
int nw_mesh_halo_transport(const nw_mesh* mesh);

/* ------------------------------------------------------------------ */
/* edge algorithms without a linear system                             */
/* ------------------------------------------------------------------ */

typedef enum { NW_PECLET_CLASSIC = 0, NW_PECLET_TANH = 1 } nw_peclet_form;
typedef struct {
  int32_t form; /* nw_peclet_form (EquationSystem::ngp_create_peclet_function,
                   include/EquationSystem.h:399-417) */
  double a;     /* classic: hybrid factor; tanh: transition c1 */
  double b;     /* classic: unused;        tanh: width c2 */
} nw_peclet_fn;

typedef struct {
  double noc_fac;         /* realm.get_noc_usage("pressure") ? 1 : 0 */
  double interp_together; /* realm.get_mdot_interp() */
} nw_mdot_opts;
/* MdotEdgeAlg::execute (src/ngp_algorithms/MdotEdgeAlg.C:43-194); reads
 * coordinates, velocity, dpdx, density, pressure, momentum_diag,
 * edge_area_vector; writes the edge field mass_flow_rate. */
int nw_mdot_edge(nw_mesh* mesh, const nw_mdot_opts* opts);

typedef struct {
  nw_peclet_fn pf;
  double eps; /* MomentumEdgePecletAlg::eps_ = 1e-16 */
} nw_peclet_opts;
/* MomentumEdgePecletAlg::execute (src/edge_kernels/MomentumEdgePecletAlg.C:49-102);
 * reads coordinates, velocity, density, viscosity_field; writes peclet_factor. */
int nw_peclet_edge(
  nw_mesh* mesh, int viscosity_field, const nw_peclet_opts* opts);

/* NodalGradAlgDriver::execute for the interior edge algorithm
 * (src/ngp_algorithms/NodalGradAlgDriver.C:30-72 +
 *  NodalGradEdgeAlg.C:58-111): pre_work grad = 0 (fused); execute grad += edge
 * contributions; post_work the shared-node sum over ranks
 * (stk::mesh::parallel_sum; a multi-rank mesh needs a communicator or
 * caller-committed halo lists, else NW_ERR_STATE) and, on a mesh with periodic
 * aliases, realm_.periodic_field_update(gradPhi) (:63-65) as
 * nw_field_periodic_update does.  phi has dim1 components (1 or ndim), grad
 * dim1*ndim. */
int nw_nodal_grad_edge(nw_mesh* mesh, int phi_field, int grad_field);
/* Two scalar nodal gradients in one launch (SST: dkdx and dwdx, the
 * NodalGradEdgeAlg instances ShearStressTransportEquationSystem runs back to
 * back on unchanged inputs, src/ShearStressTransportEquationSystem.C:247-320):
 * same arithmetic and result as two nw_nodal_grad_edge calls; the (L,R)
 * records, area vectors, half-edge lists and dual volumes are staged once.
 * Both phi fields have one component, both grad fields ndim. */
int nw_nodal_grad_edge_pair(
  nw_mesh* mesh, int phi_a, int grad_a, int phi_b, int grad_b);

/* ------------------------------------------------------------------ */
/* linear system (LinearSystem / HypreLinearSystem / HypreUVWLinearSystem) */
/* ------------------------------------------------------------------ */

typedef enum {
  NW_LINSYS_HYPRE = 0,    /* HypreLinearSystem, numDof rows per node */
  NW_LINSYS_HYPRE_UVW = 1 /* HypreUVWLinearSystem: scalar graph, ndim RHS */
} nw_linsys_kind;

/* LinearSystem::create (src/LinearSystem.C:117-159) */
int nw_linsys_create(
  nw_mesh* mesh, int kind, int num_dof, nw_linsys** out);
int nw_linsys_destroy(nw_linsys* ls);
/* skippedRows_ (Dirichlet rows, global row ids); call before finalize */
int nw_linsys_set_skipped_rows(nw_linsys* ls, const int64_t* rows, int64_t n);
/* LinearSystem::buildEdgeToNodeGraph (src/HypreLinearSystem.C:412-478) over
 * the mesh's owned edges */
int nw_linsys_build_edge_to_node_graph(nw_linsys* ls);
/* LinearSystem::finalizeLinearSystem (src/HypreLinearSystem.C:819-880): CSR
 * arrays, edge->slot map, device plans */
int nw_linsys_finalize(nw_linsys* ls);

typedef struct {
  int64_t i_lower, i_upper; /* inclusive owned row range (iLower_, iUpper_) */
  int64_t num_rows_owned, num_nonzeros_owned;
  int64_t num_rows_shared, num_nonzeros_shared;
  int64_t num_periodic_rows; /* periodic_bc_rows_owned_ */
  int32_t num_rhs;           /* columns of rhs_dev_: 1, or ndim for UVW */
  int32_t block;             /* rows (== cols) of the per-edge block: 2*numDof
                                (UVW: 2) */
} nw_linsys_sizes;
int nw_linsys_get_sizes(const nw_linsys* ls, nw_linsys_sizes* out);

/* Integer structures for bit-exact comparison with the reference
 * (src/HypreLinearSystem.C:999-1236, 883-993).  Any output may be NULL. */
int nw_linsys_get_graph(
  const nw_linsys* ls,
  int64_t* mat_row_start_owned,  /* [num_rows_owned+1] */
  int64_t* mat_row_start_shared, /* [num_rows_shared+1] */
  int64_t* cols,                 /* [nnz_owned+nnz_shared] cols_host_ */
  int64_t* rows,                 /* [nnz_owned+nnz_shared] rows_host_ */
  int64_t* row_indices_shared,   /* [num_rows_shared] ascending */
  int64_t* periodic_rows_owned); /* [num_periodic_rows] */
/* edge -> CSR slot map: for edge e (caller's order) and local block entry
 * (ii,kk), slots[(e*block + ii)*block + kk] is the index into the value array
 * that HypreLinSysCoeffApplier::sum_into{,_1DoF} / the UVW applier would add
 * lhs(ii,kk) to (after its sort + column walk), or -1 when that row is
 * skipped / absent.  rhs_rows[e*block + ii] is the row of rhs_dev_. */
int nw_linsys_get_edge_slots(
  const nw_linsys* ls, int64_t* slots, int64_t* rhs_rows);

/* LinearSystem::zeroSystem + resetCoeffApplierData
 * (src/HypreLinearSystem.C:1892-1933, 1386-1430): values = 0, rhs = 0,
 * periodic-slave rows diag 1 / rhs 0. */
int nw_linsys_zero(nw_linsys* ls);

typedef enum {
  NW_SCATTER_SEGMENTED = 0, /* default: deterministic tile kernel, row-sorted
                               segmented warp-shuffle reduction, no atomics */
  NW_SCATTER_ATOMIC = 1     /* warp-aggregated fp64 atomicAdd variant */
} nw_scatter_mode;
int nw_linsys_set_scatter_mode(nw_linsys* ls, int mode);
/* 1 when a segmented edge assembly of this (finalized) system runs on the
 * deterministic tile kernels: 1-dof and UVW systems whose reduction plan fits
 * the per-tile limits; the monolithic ndim-dof system (HypreLinearSystem with
 * numDof = ndim, sum_into of the full 2 ndim x 2 ndim block,
 * src/HypreLinearSystem.C:2059-2161) when its skipped rows cover whole nodes
 * (as applyDirichletBCs lists them) -- it then walks the node graph's plan
 * and writes ndim rows per node.  0: the warp-aggregated atomic kernels. */
int nw_linsys_uses_tile_path(const nw_linsys* ls);

/* Eager exchange of the shared rows (several ranks, peer-memory transport,
 * tile path; ignored otherwise).  The reference sums a system's shared rows
 * in loadComplete (HypreLinearSystem::loadComplete -> hypre assemble,
 * src/HypreLinearSystem.C:1236-1385), after every algorithm has contributed.
 * With on != 0 the caller promises that the edge assembly
 * (nw_assemble_*_edge) is the LAST contribution to rows shared with other
 * ranks before nw_linsys_load_complete: the assembly kernel then stores the
 * shared rows straight into their owners' memory while it runs (no separate
 * send); nw_linsys_load_complete only adds what has arrived.  Until then any call
 * that adds to the system returns NW_ERR_STATE (the addition could not reach
 * the owners any more); calls that read it complete the exchange first.
 * Results are bit-identical to the default order. */
int nw_linsys_set_eager_exchange(nw_linsys* ls, int on);

typedef struct {
  double dt, gamma1;           /* tauScale = dt / gamma1 */
  double noc_fac;              /* get_noc_usage("pressure") */
  double interp_together;      /* get_mdot_interp() */
  double solve_incompressible; /* get_incompressible_solve() */
} nw_continuity_opts;
/* ContinuityEdgeSolverAlg::execute incl. the CoeffApplier scatter
 * (src/edge_kernels/ContinuityEdgeSolverAlg.C:37-195). */
int nw_assemble_continuity_edge(nw_linsys* ls, const nw_continuity_opts* opts);

/* The optional terms of MdotEdgeAlg / ContinuityEdgeSolverAlg
 * (src/ngp_algorithms/MdotEdgeAlg.C:153-163, 175-180;
 * src/edge_kernels/ContinuityEdgeSolverAlg.C:147-158, 172-177), off in the
 * ABL / airfoil decks: balanced buoyancy forcing
 * (solutionOptions_->use_balanced_buoyancy_force_: gravity vector, nodal fields
 * buoyancy_source [ndim] and buoyancy_source_mask [1]) and the GCL term of
 * deforming meshes (realm_.has_mesh_deformation(): edge field
 * edge_face_velocity_mag [1]).  The *_ext entry points take them in `extra`;
 * they use direct-gather kernels (the continuity one scatters with fp64
 * atomics through the edge->slot map and accumulates like the atomic mode),
 * the default entry points above stay on the tile kernels. */
typedef struct {
  int32_t add_balanced_forcing;
  double gravity[3];
  int32_t source_field;            /* buoyancy_source */
  int32_t source_mask_field;       /* buoyancy_source_mask */
  int32_t needs_gcl;
  int32_t edge_face_vel_mag_field; /* edge_face_velocity_mag */
} nw_mdot_extra_opts;
int nw_mdot_edge_ext(
  nw_mesh* mesh, const nw_mdot_opts* opts, const nw_mdot_extra_opts* extra);
int nw_assemble_continuity_edge_ext(
  nw_linsys* ls, const nw_continuity_opts* opts,
  const nw_mdot_extra_opts* extra);

typedef struct {
  double alpha, alpha_upw, ho_upwind, relax_fac;
  int32_t use_limiter;
  double eps; /* 1e-16 */
  nw_peclet_fn pf;
} nw_scalar_opts;
/* ScalarEdgeSolverAlg::execute (src/edge_kernels/ScalarEdgeSolverAlg.C:55-206)
 * for scalar q with gradient dqdx and diffusive-flux coefficient field. */
int nw_assemble_scalar_edge(
  nw_linsys* ls,
  int q_field,
  int dqdx_field,
  int diff_flux_coeff_field,
  const nw_scalar_opts* opts);
/* The TKE and SDR assemblies of ShearStressTransportEquationSystem::
 * solve_and_update in one launch (src/ShearStressTransportEquationSystem.C:
 * 247-320: both systems are assembled from the same state -- the solves write
 * kTmp / wTmp, update_and_clip comes after both).  Equivalent to
 * nw_assemble_scalar_edge(ls_a, ...) followed by nw_assemble_scalar_edge(ls_b,
 * ...): same arithmetic, same results bit for bit; what the two share
 * (coordinates, velocity, density, edge streams, the reduction plan) is staged
 * once per tile.  Both systems must be 1-dof hypre systems of the same mesh
 * with the same skipped rows.  The fused kernel (one 512-thread CTA per tile,
 * ~150 KB of shared memory) is opt-in, NW_SCALAR_PAIR_FUSED=1: on a B200 it
 * measured slower than the two launches (one CTA per SM leaves the staging
 * latency uncovered); by default, and whenever the fused
 * path does not apply (atomic scatter mode, accumulation onto a non-empty
 * system, tile too large), the two assemblies run one after the other. */
int nw_assemble_scalar_edge_pair(
  nw_linsys* ls_a, int q_a, int dqdx_a, int diff_flux_coeff_a,
  const nw_scalar_opts* opts_a,
  nw_linsys* ls_b, int q_b, int dqdx_b, int diff_flux_coeff_b,
  const nw_scalar_opts* opts_b);

typedef struct {
  double include_divu, alpha, alpha_upw, ho_upwind, relax_fac;
  int32_t use_limiter;
  double eps; /* 1e-16 */
  /* fuse MomentumEdgePecletAlg into the kernel instead of reading the
   * peclet_factor edge field; pf/pec_eps only used then */
  int32_t fuse_peclet;
  nw_peclet_fn pf;
  double pec_eps;
  /* NGPApplyCoeff::extract_diagonal (src/SolverAlgorithm.C:87-105): nodal
   * field that accumulates lhs(i*ndim, i*ndim) of both end nodes, or -1 */
  int32_t diag_field;
  /* solutionOptions_->realm_has_vof_ (src/edge_kernels/MomentumEdgeSolverAlg.C:
   * 52-64, 88, 124-125, 174-192): the edge mass flow is mass_flow_rate +
   * mass_vof_balanced_flow_rate (edge field of that name, 1 component, must be
   * registered) and alphaUpw, the Peclet factor and the limited extrapolation
   * are pushed to full upwinding across a density jump,
   * 1 - erf(6 |rhoL - rhoR| / min(rhoL, rhoR)).  0: off (both fields alias
   * mass_flow_rate in the reference, the factor is exactly 1). */
  int32_t has_vof;
} nw_momentum_opts;
/* MomentumEdgeSolverAlg::execute (src/edge_kernels/MomentumEdgeSolverAlg.C:70-313)
 * into a NW_LINSYS_HYPRE_UVW system (x-x entries + ndim RHS,
 * src/HypreUVWLinearSystem.C:695-767) or a monolithic NW_LINSYS_HYPRE system
 * with num_dof == ndim (src/HypreLinearSystem.C:2059-2161). */
int nw_assemble_momentum_edge(
  nw_linsys* ls, int viscosity_field, const nw_momentum_opts* opts);

/* Generic CoeffApplier entry (include/LinearSystem.h:62-70) for callers that
 * computed their own per-entity blocks on the device: n_entities entities of
 * nodes_per_entity nodes (local node indices, device pointer), lhs
 * [n_entities][n][n] row-major and rhs [n_entities][n] with
 * n = nodes_per_entity*numDof (device pointers).  Uses fp64 atomics. */
int nw_linsys_sum_into(
  nw_linsys* ls,
  int64_t n_entities,
  int nodes_per_entity,
  const int32_t* d_entity_nodes,
  const double* d_lhs,
  const double* d_rhs);

/* GeometryInteriorAlg<AlgTraitsHex8> (src/ngp_algorithms/GeometryInteriorAlg.C:
 * 72-112 dual nodal volume, 165-225 edge area vectors; HexSCV / HexSCS
 * determinants src/master_element/Hex8CVFEM.C:365-390, 567-592) on the device,
 * so that a moving-mesh run refreshes the geometric inputs of the edge
 * kernels without the host.  elem_nodes: host array [n_elems][8] of local node
 * indices in the Hex8 node order.  Volumes are accumulated from the elements
 * flagged in elem_owned (NULL: all; the reference's locally_owned selector --
 * follow with nw_field_parallel_sum), area vectors from every element given,
 * into the mesh's (locally owned) edges with the reference's sign rule (left
 * sub-control-volume node == first edge node).  Accumulates with fp64 atomics:
 * zero the fields first (GeometryAlgDriver::pre_work) with nw_field_fill.
 * Either field id may be -1.  coordinates_field: the (current) coordinates. */
int nw_geometry_interior_hex8(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field);
/* The same for 2-D Quad4 blocks (AlgTraitsQuad4_2D; Quad42DSCV / Quad42DSCS
 * determinants, src/master_element/Quad42DCVFEM.C:139-200, 384-445);
 * elem_nodes is [n_elems][4]. */
int nw_geometry_interior_quad4(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field);

/* The same for Tet4, Wed6 and Pyr5 blocks of a 3-D mesh (AlgTraitsTet4 / Wed6 /
 * Pyr5; TetSCV / TetSCS src/master_element/Tet4CVFEM.C:243-343, 522-619,
 * WedSCV / WedSCS Wed6CVFEM.C:267-369, 541-647, PyrSCV / PyrSCS
 * Pyr5CVFEM.C:348-572, 772-900 -- the pyramid has two sub-control surfaces on
 * each apex edge, scsIpEdgeOrd include/master_element/Pyr5CVFEM.h:304-305).
 * elem_nodes is [n_elems][4 | 6 | 5] in the stk / Exodus node order of the
 * topology.  A mixed mesh calls once per block into the same two fields, as
 * GeometryAlgDriver runs one GeometryInteriorAlg per topology. */
int nw_geometry_interior_tet4(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field);
int nw_geometry_interior_wed6(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field);
int nw_geometry_interior_pyr5(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field);

/* Poisson system of the SST minimum wall distance:
 * WallDistEdgeSolverAlg::execute (src/edge_kernels/WallDistEdgeSolverAlg.C:28-66,
 * lhs = asq/axdx [[+1,-1],[-1,+1]], no rhs; same tile / atomic kernels as the
 * other edge systems) and WallDistNodeKernel::execute
 * (src/node_kernels/WallDistNodeKernel.C:34-43, rhs += dual_nodal_volume on
 * owned non-slave nodes). */
int nw_assemble_wall_dist_edge(nw_linsys* ls);
int nw_assemble_wall_dist_node(nw_linsys* ls, int dual_nodal_volume_field);

/* Time-derivative node kernels through AssembleNGPNodeSolverAlgorithm
 * (src/AssembleNGPNodeSolverAlgorithm.C:85-146: every locally-owned node that
 * is not a periodic slave; 1-node block zeroed, kernel, CoeffApplier):
 *   NW_MASS_SCALAR     ScalarMassBDFNodeKernel::execute
 *                      (src/node_kernels/ScalarMassBDFNodeKernel.C:72-96)
 *   NW_MASS_MOMENTUM   MomentumMassBDFNodeKernel::execute
 *                      (src/node_kernels/MomentumMassBDFNodeKernel.C:80-107)
 *   NW_MASS_CONTINUITY ContinuityMassBDFNodeKernel::execute
 *                      (src/node_kernels/ContinuityMassBDFNodeKernel.C:63-84)
 * Accumulates into the system (call after the edge assembly, before
 * loadComplete).  Field ids name the three time states of each field; a
 * two-state field passes its N id for NM1 as the reference does. */
typedef enum {
  NW_MASS_SCALAR = 0,
  NW_MASS_MOMENTUM = 1,
  NW_MASS_CONTINUITY = 2
} nw_mass_kind;
typedef struct {
  double dt, gamma1, gamma2, gamma3;
  int32_t q_nm1, q_n, q_np1;       /* scalar (1 comp) / velocity (ndim); unused for continuity */
  int32_t rho_nm1, rho_n, rho_np1; /* density */
  int32_t dnv_nm1, dnv_n, dnv_np1; /* dual_nodal_volume */
  int32_t dpdx;                    /* momentum only */
} nw_mass_bdf_opts;
int nw_assemble_mass_bdf_node(
  nw_linsys* ls, int kind, const nw_mass_bdf_opts* opts);

/* CoeffApplier::resetRows (include/LinearSystem.h:53-60;
 * HypreLinSysCoeffApplier::reset_rows src/HypreLinearSystem.C:2262-2330,
 * HypreUVWLinSysCoeffApplier::reset_rows src/HypreUVWLinearSystem.C:787-851):
 * zero the matrix rows of the given local nodes (all dofs; owned rows and rows
 * in the shared tail), set their diagonal to diag_value and their rhs (every
 * column) to rhs_residual.  FixPressureAtNodeAlgorithm::execute
 * (src/FixPressureAtNodeAlgorithm.C:57-121) is this call with (0, 0) followed
 * by nw_linsys_sum_into of the 1x1 block lhs = 1, rhs = refPressure - p.
 * `nodes` is a host array. */
int nw_linsys_reset_rows(
  nw_linsys* ls, int64_t n_nodes, const int32_t* nodes, double diag_value,
  double rhs_residual);

/* HypreLinearSystem::applyDirichletBCs (src/HypreLinearSystem.C:2407-2457) /
 * HypreUVWLinearSystem::applyDirichletBCs (src/HypreUVWLinearSystem.C:377-427):
 * for every locally-owned node of the list and every dof d, the first entry of
 * the row (the diagonal: Dirichlet rows are the skipped, diagonal-only rows of
 * nw_linsys_set_skipped_rows) becomes 1 and rhs = bc_values(node, d) -
 * solution(node, d).  Fields are nodal with one component per dof.  `nodes` is
 * a host array of local node indices. */
int nw_linsys_apply_dirichlet_bcs(
  nw_linsys* ls, int solution_field, int bc_values_field, int64_t n_nodes,
  const int32_t* nodes);


/* ---- shared-row exchange structure (multi-rank) ----
 * The rows of the shared tail destined to one owner form one contiguous
 * segment (rows ascending, ranks own contiguous row ranges), sent in place.
 * With a communicator the structure is exchanged inside nw_linsys_finalize;
 * otherwise: send_info / get_send on the sender -> set_recv on the owner ->
 * commit on every rank.  Received (row, col) pairs that the owner's local
 * graph does not contain (the edge lives on the sender) get one slot each in a
 * COO extension behind the reference-layout arrays: values[nnz_owned +
 * nnz_shared + x], rows/cols from nw_linsys_get_extra -- the hand-off to
 * HYPRE_IJMatrixAddToValues2 then needs no off-rank rows at all. */
int nw_linsys_halo_send_info(
  const nw_linsys* ls, int peer, int64_t* n_rows, int64_t* n_vals);
int nw_linsys_halo_get_send(
  const nw_linsys* ls, int peer, int64_t* rows, int64_t* row_lens,
  int64_t* cols);
int nw_linsys_halo_set_recv(
  nw_linsys* ls, int peer, int64_t n_rows, const int64_t* rows,
  const int64_t* row_lens, const int64_t* cols);
int nw_linsys_halo_commit(nw_linsys* ls);
/* destination of every received value / rhs entry (bit-exact halo lists) */
int nw_linsys_halo_get_recv_slots(
  const nw_linsys* ls, int peer, int64_t* n_vals, int64_t* val_slots,
  int64_t* n_rows, int64_t* rhs_rows);
int nw_linsys_get_extra(
  const nw_linsys* ls, int64_t* n_extra, int64_t* rows, int64_t* cols);

/* LinearSystem::loadComplete (src/HypreLinearSystem.C:1848-1889): the
 * shared-row halo sum.  Single rank: no-op. */
/* The reference's pre-assembly dump (solver option
 * write_preassembly_matrix_files; HypreLinearSystem::hypreIJMatrixSetAddToValues
 * src/HypreLinearSystem.C:1517-1568, hypreIJVectorSetAddToValues :1625-1661,
 * HypreUVWLinearSystem.C:135-166), byte for byte, so that a site with a real
 * nalu-wind build can diff this library's assembly against its own:
 *   <eq>.IJM.<n>.mat.<rank%05d>.preassem.{i,j,v,meta}
 *   <eq>[d].IJV.<n>.rhs.<rank%05d>.preassem.{i,v,meta}   ([d] for UVW systems)
 * i/j: HypreIntType row / column ids of the owned entries followed by the
 * shared tail; v: doubles; mat meta = {globalNumRows, iLower, iUpper,
 * nnzOwned, nnzShared, nnzOwned+nnzShared}; rhs meta = {rowsOwned, rowsShared,
 * rowsOwned+rowsShared}.  Like the reference, call it before
 * nw_linsys_load_complete (the state hypre would receive).  hypre_int_bytes =
 * sizeof(HYPRE_Int) of the build to compare with (4, or 8 for bigint).
 * `directory` may be NULL (current directory).  Synchronises the stream. */
int nw_linsys_write_preassembly_files(
  nw_linsys* ls, const char* directory, const char* eq_sys_name,
  int write_counter, int hypre_int_bytes);

int nw_linsys_load_complete(nw_linsys* ls);
/* transport nw_linsys_load_complete uses (nw_halo_transport) */
int nw_linsys_halo_transport(const nw_linsys* ls);

/* Device pointers in exactly the layout the reference hands to
 * HYPRE_IJMatrixSetValues2 / AddToValues2 and HYPRE_IJVectorSetValues
 * (src/HypreLinearSystem.C:1572-1590, 1665-1673): values[nnz_owned+nnz_shared],
 * rhs column-major [(rows_owned+rows_shared) x num_rhs]. */
int nw_linsys_device_arrays(
  nw_linsys* ls, double** values, double** rhs, int64_t* rhs_stride);
/* copy to host; either may be NULL; synchronises */
int nw_linsys_get_values(nw_linsys* ls, double* values, double* rhs);
/* sum of squares of the owned part of every rhs column (device reduction,
 * deterministic); feeds the nonlinear residual norm
 * (src/HypreLinearSystem.C:2534, 2611-2638).  out[num_rhs]. */
int nw_linsys_rhs_norm2(nw_linsys* ls, double* out);
/* the same summed over all ranks (one ncclAllReduce of num_rhs doubles): the
 * squared nonlinear residual norm of the whole system.  Collective. */
int nw_linsys_rhs_norm2_global(nw_linsys* ls, double* out);

#ifdef __cplusplus
}
#endif
#endif
