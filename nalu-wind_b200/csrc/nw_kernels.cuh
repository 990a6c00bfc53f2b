/*
 * nw_kernels.cuh -- launch interface between the C-ABI layer (nw_api.cu) and
 * the sm_100a kernels (nw_kernels.cu).
 */
#ifndef NW_KERNELS_CUH
#define NW_KERNELS_CUH

#include <cuda_runtime.h>

#include "nw_types.h"

namespace nw {

constexpr int kMaxNodeComps = 20;

/* diagnostic per-phase cycle counters (filled only by the NW_PHASE_TIMING
 * build): [kernel][slot], last slot = CTA count.  Kernels: 0 continuity,
 * 1 scalar, 2 momentum, 3 mdot, 4 peclet, 5 grad scalar, 6 grad vector. */
constexpr int kPhaseKernels = 16; /* 8.. : stream kernels */
constexpr int kPhaseSlots = 12;
cudaError_t phase_times_read(unsigned long long* out, bool reset);

/* peer-memory mailbox as the kernels see it (nw_p2p in nw_internal.h) */
struct P2pDev
{
  double* const* peerWindow = nullptr;            /* [nranks] window base of rank r */
  unsigned long long* const* peerFlags = nullptr; /* [nranks] flag array of rank r */
  double* myWindow = nullptr;
  const unsigned long long* myFlags = nullptr;
  unsigned* sync = nullptr; /* [0] block counter, [1] error word */
  /* ranks signalled / waited for: the union of the neighbours of every
   * exchange object of the context (see p2p_register_peers) */
  const int32_t* peers = nullptr;
  int nPeers = 0;
  int myRank = 0;
  int64_t winOff = 0; /* parity * winDoubles */
  unsigned long long epoch = 0;
  long long timeoutCycles = 40000000000ll; /* bounded spin of the pull kernels */
};

/* component arrays of one or more nodal fields */
struct CompPtrs
{
  double* c[9];
};

/* fused push of a tile kernel (eager exchange, nw_halo.inc): the shared tail
 * is one contiguous segment per owner; segment k of the tail values / rows
 * lands in that owner's window */
struct PushSeg
{
  int64_t val0 = 0, valN = 0, valDst = 0; /* tail value range -> window offset */
  int64_t row0 = 0, rowN = 0, rowDst = 0; /* tail row range -> window offset ... */
  int64_t rowStride = 0;                  /* ... of rhs column 0; next column + rowStride */
  int32_t peer = 0;
  int32_t pad = 0;
};
struct LsPushDev
{
  const PushSeg* seg = nullptr; /* null: nothing to send */
  int enabled = 0;              /* 0: no fused push in this launch */
  int nSeg = 0;
  int nSendTiles = 0; /* tiles with hasShared */
  int64_t nnzOwned = 0, numRowsOwned = 0;
  P2pDev pp;
};

/* fused push of the gradient kernel: the nodal send list grouped by tile */
struct NodePushDev
{
  const int32_t* tilePtr = nullptr; /* [nTiles + 1], launch order; null: no fused push */
  const int32_t* slot = nullptr;    /* internal node slot of entry g */
  const int32_t* peer = nullptr;
  const int64_t* dst = nullptr;     /* window entry (x nc components) */
  int nc = 0;                       /* components per window entry */
  int nSendTiles = 0;
  P2pDev pp;
};

struct MeshPlanDev
{
  const TileHdr* tiles = nullptr;
  const int32_t* haloNodes = nullptr;
  const int32_t* haloBlock = nullptr; /* [nTiles][kHaloBlock], -1 padded */
  const uint32_t* lr = nullptr;
  const uint32_t* heNodeEll = nullptr;   /* sliced-ELL node-keyed half-edges */
  const int32_t* sliceOffNode = nullptr;
  const uint8_t* primary = nullptr;
  int nTiles = 0;
  int ndim = 3;
  int maxStaged = 0;   /* max over tiles of even(nOwnPad + nHalo) */
  int maxTileEdges = 0;
  int maxTileNodes = 0;
  int maxTileEllNode = 0;
  /* NW_DBG_SKIP (timing experiments only, results are wrong): bit 0 no halo
   * gather, bit 1 no edge physics, bit 2 no row reduction / copy-out, bit 3
   * no own-range node copies */
  int dbgSkip = 0;
};

struct LsPlanDev
{
  const LsTileHdr* tiles = nullptr;
  const EntInfo* entInfo = nullptr;
  const int32_t* entRhsRow = nullptr;
  const uint32_t* heEll = nullptr;   /* sliced-ELL row-keyed half-edges */
  const int32_t* sliceOff = nullptr;
  const int32_t* entGo = nullptr; /* value offset of each tile row */
  int maxTileNnz = 0;
  int maxTileEnts = 0;
  int maxTileEll = 0;
  double* values = nullptr;
  double* rhs = nullptr;
  int64_t rhsStride = 0; /* rows_owned + rows_shared */
  /* NGPApplyCoeff::extract_diagonal target (nodal field, internal slots) of
   * the current launch, or null */
  double* diagOut = nullptr;
  LsPushDev push;
};

/* atomic-variant slot map, per tile-edge slot */
struct AtomicMapDev
{
  const int32_t* slots = nullptr;   /* [nTileEdgeSlots][4]: LL, LR, RL, RR */
  const int32_t* rhsRows = nullptr; /* [nTileEdgeSlots][2] */
};

/* node component pointers (SoA, internal numbering) */
struct NodeComps
{
  const double* c[kMaxNodeComps];
};
struct EdgeComps
{
  const double* area[3];
  const double* mdot;
  const double* pecfac;
};

/* optional mdot / continuity terms (balanced buoyancy forcing, GCL): device
 * view of nw_mdot_extra_opts */
struct ContExtraDev
{
  int balanced = 0, gcl = 0;
  double gravity[3] = {0.0, 0.0, 0.0};
  const double* smask = nullptr;  /* nodal, 1 comp */
  const double* src[3] = {nullptr, nullptr, nullptr}; /* nodal, ndim comps */
  const double* faceVelMag = nullptr; /* edge (tile-edge slots), 1 comp */
};
/* extended variants: direct gathers, one thread per tile-edge; continuity
 * scatters through the atomic slot map (values / rhs must be zeroed or hold
 * what is to be accumulated onto) */
cudaError_t launch_mdot_ext(
  const MeshPlanDev& mp, const NodeComps& nc, const EdgeComps& ec,
  const ContExtraDev& ex, double* mdotOut, nw_mdot_opts o, cudaStream_t s);
cudaError_t launch_continuity_ext_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, const ContExtraDev& ex,
  nw_continuity_opts o, cudaStream_t s);

/* every launcher returns the cudaError_t of the launch */
cudaError_t launch_mdot_tile(
  const MeshPlanDev& mp, const NodeComps& nc, const EdgeComps& ec,
  double* mdotOut, nw_mdot_opts o, cudaStream_t s);
cudaError_t launch_peclet_tile(
  const MeshPlanDev& mp, const NodeComps& nc, double* pecfacOut,
  nw_peclet_opts o, cudaStream_t s);
cudaError_t launch_grad_tile(
  const MeshPlanDev& mp, int dim1, const NodeComps& phi, const double* dualVol,
  const EdgeComps& ec, double* const* gradOut /* dim1*ndim comps */,
  cudaStream_t s, const NodePushDev* push = nullptr);

cudaError_t launch_continuity_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, nw_continuity_opts o, cudaStream_t s);
/* WallDistEdgeSolverAlg: coordinates in nc.c[0..ndim), area in ec */
cudaError_t launch_wall_dist_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, cudaStream_t s);
cudaError_t launch_wall_dist_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, cudaStream_t s);
cudaError_t launch_wall_dist_node(
  const int64_t* rows, int64_t nRows, const double* dualVol, double* rhs,
  cudaStream_t s);
cudaError_t launch_scalar_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, nw_scalar_opts o, cudaStream_t s);
/* shared-memory check of a policy's tile kernel (0 continuity, 1 scalar,
 * 2 momentum UVW, 7 wall distance): false -> use the atomic path */
bool ls_tile_fits(const MeshPlanDev& mp, const LsPlanDev& lp, int policy);
/* two scalar systems sharing graph and plan (lpA: plan + values / rhs of
 * system A); nc: x, v, rho, then per system q, dqdx, diffFluxCoeff.
 * *launched = false (and no error): the tile does not fit one CTA's shared
 * memory, assemble the systems one by one */
cudaError_t launch_scalar_pair_tile(
  const MeshPlanDev& mp, const LsPlanDev& lpA, double* valuesB, double* rhsB,
  const NodeComps& nc, const EdgeComps& ec, nw_scalar_opts oA,
  nw_scalar_opts oB, bool* launched, cudaStream_t s);
/* monolithic ndim-dof momentum on the tile path: lp carries the node graph's
 * plan with entGo / entRhsRow = value offset / local row of each node's first
 * dof (nw_api.cu: build_mono_twin) */
cudaError_t launch_momentum_mono_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, nw_momentum_opts o, double* diagOut, cudaStream_t s);
cudaError_t launch_momentum_uvw_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, nw_momentum_opts o, double* diagOut, cudaStream_t s);

/* atomic variants (accumulate into values / rhs) */
cudaError_t launch_continuity_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, nw_continuity_opts o,
  cudaStream_t s);
cudaError_t launch_scalar_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, nw_scalar_opts o, cudaStream_t s);
cudaError_t launch_momentum_uvw_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, nw_momentum_opts o,
  double* diagOut, cudaStream_t s);
/* monolithic momentum (numDof == ndim): 2ndim x 2ndim blocks through the
 * per-edge slot map [nTileEdgeSlots][4*ndim*ndim] */
cudaError_t launch_momentum_mono_atomic(
  const MeshPlanDev& mp, const int32_t* slots, const int32_t* rhsRows,
  double* values, double* rhs, const NodeComps& nc, const EdgeComps& ec,
  nw_momentum_opts o, double* diagOut, cudaStream_t s);
cudaError_t launch_grad_atomic(
  const MeshPlanDev& mp, int dim1, const NodeComps& phi, const double* dualVol,
  const EdgeComps& ec, double* const* gradOut, cudaStream_t s);

/* utility kernels */
cudaError_t launch_node_gather(   /* AoS caller order -> SoA internal */
  const double* srcAos, int ncomp, const int32_t* nodeOfSlot, int64_t nSlots,
  double* dstSoa, cudaStream_t s);
cudaError_t launch_node_scatter(  /* SoA internal -> AoS caller order */
  const double* srcSoa, int ncomp, const int32_t* nodeOfSlot, int64_t nSlots,
  double* dstAos, cudaStream_t s);
cudaError_t launch_edge_gather(
  const double* srcAos, int ncomp, const int32_t* tileEdgeSrc, int64_t nSlots,
  double* dstSoa, cudaStream_t s);
cudaError_t launch_edge_scatter(
  const double* srcSoa, int ncomp, const int32_t* primarySlotOfEdge,
  int64_t nEdges, int64_t slotStride, double* dstAos, cudaStream_t s);
cudaError_t launch_fill(double* p, int64_t n, double v, cudaStream_t s);
cudaError_t launch_edge_sum(
  const double* a, const double* b, int64_t n, double* out, cudaStream_t s);
/* rows no tile writes: values zero (periodic rows: diagonal 1), rhs zero */
cudaError_t launch_row_init(
  const int32_t* rows, int nRows, const int64_t* rowPtr /* [R+1] */,
  const uint8_t* isPeriodic, double* values, double* rhs, int64_t rhsStride,
  int nRhs, cudaStream_t s);
/* GeometryInteriorAlg<Hex8>: elemSlots [n][8] node slots, elemEdges [n][12] =
 * 2 * primary tile-edge slot + (1: negate), -1: edge not local; accumulates
 * with fp64 atomics (as the reference does) */
cudaError_t launch_geometry_hex8(
  int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s);
/* GeometryInteriorAlg<Quad4_2D>: elemSlots [n][4], elemEdges [n][4] */
cudaError_t launch_geometry_quad4(
  int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s);
/* GeometryInteriorAlg<Tet4 / Wed6 / Pyr5> (nw_geometry.cu; topology = the
 * nw::geo::Topology of geometry_cvfem.h): elemSlots [n][npe], elemEdges [n][nScs] */
cudaError_t launch_geometry_cvfem(
  int topology, int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s);
cudaError_t launch_edge_mirror(
  const int32_t* primarySlot, const int32_t* secondSlot, int64_t nEdges,
  int ncomp, int64_t stride, double* f, cudaStream_t s);
/* mass-BDF node kernels; rows: int64[n][4] = slot, diagonal value offset, rhs
 * row, dof (-1: UVW system, all rhs columns).  q / rho / dnv: the NM1, N, NP1
 * states (SoA, stride fieldStride). */
struct MassBdfFields
{
  const double* q[3];
  const double* rho[3];
  const double* dnv[3];
  const double* dpdx;
  int64_t fieldStride;
};
cudaError_t launch_mass_bdf_node(
  int kind, int ndim, const int64_t* rows, int64_t nRows,
  const MassBdfFields& f, double dt, double gamma1, double gamma2,
  double gamma3, double* values, double* rhs, int64_t rhsStride,
  cudaStream_t s);
/* CoeffApplier::resetRows; rows: int64[nRows][4] = value offset, length,
 * diagonal position (-1: none), rhs row */
cudaError_t launch_reset_rows(
  const int64_t* rows, int64_t nRows, double diagValue, double rhsResidual,
  double* values, double* rhs, int64_t rhsStride, int nRhs, cudaStream_t s);
/* applyDirichletBCs; rows: int64[nRows][5] = value offset of the row's first
 * entry, rhs row, rhs column, field slot, field component */
cudaError_t launch_dirichlet_rows(
  const int64_t* rows, int64_t nRows, const double* solution, const double* bc,
  int64_t fieldStride, double* values, double* rhs, int64_t rhsStride,
  cudaStream_t s);
/* deterministic sum of squares of rhs[d*stride + (0..n)] for each d */
cudaError_t launch_norm2(
  const double* rhs, int64_t n, int64_t stride, int nRhs, double* partial,
  int nPartial, double* out, cudaStream_t s);
/* generic CoeffApplier entry: binary-search column lookup + atomics */
cudaError_t launch_sum_into(
  int64_t nEnt, int npe, int numDof, const int32_t* entNodes,
  const int64_t* nodeHid, const double* lhs, const double* rhsIn,
  int64_t iLower, int64_t iUpper, int64_t nRowsOwned, int64_t nnzOwned,
  const int64_t* rowStartOwned, const int64_t* rowStartShared,
  const int64_t* rowIndicesShared, int64_t nRowsShared, const int64_t* cols,
  const int64_t* skipped, int64_t nSkipped, int uvwDim, double* values,
  double* rhs, int64_t rhsStride, cudaStream_t s);
/* halo exchange helpers */
cudaError_t launch_pack(
  const double* src, const int64_t* idx, int64_t n, double* dst,
  cudaStream_t s);
cudaError_t launch_scatter_assign(
  const double* src, const int64_t* idx, int64_t n, double* dst,
  cudaStream_t s);
cudaError_t launch_p2p_push_nodal(
  const double* base, int64_t stride, int nc, const int64_t* sendIdx,
  const int32_t* sendPeer, const int64_t* sendDst, int64_t n, const P2pDev& pp,
  cudaStream_t s);
/* beside: the launch shares the GPU with compute kernels of another stream
 * (asynchronous completion) -- small grid */
cudaError_t launch_p2p_pull_nodal(
  const CompPtrs& comps /* nc component arrays, internal slots */, int nc,
  const int64_t* recvIdx, int64_t n, const P2pDev& pp, bool beside,
  cudaStream_t s, bool signal = false);
cudaError_t launch_p2p_pull_assign_nodal(
  double* base, int64_t stride, int nc, const int64_t* recvIdx,
  const unsigned char* recvIsGhost, int64_t n, const P2pDev& pp, bool beside,
  cudaStream_t s);
cudaError_t launch_p2p_push_segments(
  const double* const* segSrc, const int64_t* segStart, const int64_t* segDst,
  const int32_t* segPeer, int nSeg, int64_t total, const P2pDev& pp,
  cudaStream_t s);
cudaError_t launch_p2p_pull_accumulate2(
  const int64_t* valDst, const int64_t* valPtr, const int64_t* valPos,
  int64_t nVal, double* values, int64_t rhsOff, int64_t rhsColStride, int nR,
  const int64_t* rhsDst, const int64_t* rhsPtr, const int64_t* rhsPos,
  int64_t nRhs, double* rhs, int64_t rhsStride, const P2pDev& pp, bool beside,
  cudaStream_t s, bool signal = false);
cudaError_t launch_p2p_pull_accumulate(
  int64_t bufOff, int64_t entStride, int64_t compStride, int nc,
  const int64_t* dstIdx, const int64_t* ptr, const int64_t* pos, int64_t nDst,
  double* dst, int64_t dstCompStride, const P2pDev& pp, bool wait, bool beside,
  cudaStream_t s, bool signal = false);

/* all peers + all components in one launch; buffer element of concatenated
 * entry g, component c at buf[g * entStride + c * compStride] */
cudaError_t launch_pack_multi(
  const double* src, int64_t srcCompStride, int nc, const int64_t* idx,
  int64_t n, double* buf, int64_t entStride, int64_t compStride,
  cudaStream_t s);
cudaError_t launch_scatter_multi(
  const double* buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* idx, int64_t n, double* dst, int64_t dstCompStride,
  cudaStream_t s);
/* dst[dstIdx[u]] += sum of buf entries pos[ptr[u] .. ptr[u+1]) in that order */
cudaError_t launch_accumulate_multi(
  const double* buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* dstIdx, const int64_t* ptr, const int64_t* pos, int64_t nDst,
  double* dst, int64_t dstCompStride, cudaStream_t s);
/* periodic_field_update: groups in CSR form over slots, master first */
cudaError_t launch_periodic_update(
  double* base, int64_t stride, int nc, const int32_t* ptr,
  const int32_t* slots, int nGroups, cudaStream_t s);
/* apply_constraints with setSlaves only: slaves take the master's value */
cudaError_t launch_periodic_set(
  double* base, int64_t stride, int nc, const int32_t* ptr,
  const int32_t* slots, int nGroups, cudaStream_t s);
/* LowMach::udiag_post_processing over the selected node slots
 * (src/LowMachEquationSystem.C:2783-2790) */
cudaError_t launch_udiag_post(
  double* udiag, const double* rho, const double* dvol, const int32_t* slots,
  int64_t n, double projTimeScale, double alphaU, cudaStream_t s);
cudaError_t launch_unpack_add(
  const double* src, const int64_t* idx, int64_t n, double* dst,
  cudaStream_t s);

} // namespace nw

#endif
