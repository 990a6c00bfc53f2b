/*
 * plan.h -- host-side plan builder (pure C++17, no CUDA).
 *
 * From one rank's mesh partition it derives
 *   MeshPlan : locality tiles (recursive coordinate bisection over row groups),
 *              the internal node numbering (tile-contiguous, TMA-aligned), the
 *              per-tile halo lists, the tile-edge lists (cut edges appear in
 *              both tiles) and the node-keyed half-edge lists;
 *   Graph    : the hypre-IJ CSR graph of LinearSystem::buildEdgeToNodeGraph +
 *              finalizeLinearSystem (src/HypreLinearSystem.C:412-478,
 *              999-1236, 883-993) and the integer edge->CSR-slot map that
 *              replaces the per-entry column walk of sum_into (:2165-2239);
 *   LsPlan   : the per-tile row staging layout, row-keyed half-edge lists and
 *              copy-out runs used by the deterministic tile kernels.
 * The graph / slot map / halo lists are exported through the C ABI and compared
 * bit-for-bit with the CPU oracle in tests/.
 */
#ifndef NW_PLAN_H
#define NW_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

#include "nw_types.h"

namespace nw {

struct MeshInput
{
  int ndim = 3;
  int rank = 0, nranks = 1;
  int64_t nNodes = 0, nEdges = 0;
  const int32_t* edgeNodes = nullptr;
  const int64_t* nodeHid = nullptr;
  const int64_t* nodeOwnHid = nullptr; /* may be null */
  const int64_t* hypreOffsets = nullptr;
  const double* coords = nullptr;
  int tileNodes = 0;
};

struct MeshPlan
{
  int ndim = 3;
  int rank = 0, nranks = 1;
  int64_t nNodes = 0, nEdges = 0;
  int64_t iLowerNode = 0, iUpperNode = -1; /* owned node-row range, inclusive */
  std::vector<int64_t> hypreOffsets;
  std::vector<int32_t> edgeNodes; /* copy, caller order */
  std::vector<int64_t> nodeHid;   /* copy, resolved */

  /* internal numbering */
  int64_t nSlots = 0;               /* >= nNodes, even-aligned tile starts */
  std::vector<int32_t> slotOfNode;  /* [nNodes] */
  std::vector<int32_t> nodeOfSlot;  /* [nSlots], -1 = padding */
  std::vector<int32_t> tileOfNode;  /* [nNodes] */

  int64_t nTiles = 0;
  std::vector<TileHdr> tiles;
  std::vector<int32_t> haloNodes; /* internal slots, per tile ascending */
  /* the first kHaloBlock halo slots of every tile at a fixed stride (entry
   * t * kHaloBlock + k, -1 = none): thread k of the tile's CTA loads its
   * halo index together with the tile header, one dependent load less in the
   * staging chain */
  std::vector<int32_t> haloBlock;

  /* tile-edge arrays (slot index space, even-aligned tile starts) */
  int64_t nTileEdgeSlots = 0;
  std::vector<int32_t> tileEdgeSrc;  /* [slots] caller edge index, -1 = pad */
  std::vector<uint32_t> lr;          /* [slots] localL | localR<<16 */
  std::vector<uint8_t> tileEdgePrimary; /* [slots] 1 on the copy in tile(L) */
  std::vector<int32_t> primarySlotOfEdge; /* [nEdges] */
  std::vector<int32_t> secondSlotOfEdge;  /* [nEdges] copy in tile(R) or -1 */

  /* node-keyed half-edges: flat sorted list + sliced-ELL form (nw_types.h) */
  std::vector<uint32_t> heNode;
  std::vector<int32_t> warpSplitNode; /* per tile kMaxWarps+1 */
  std::vector<uint32_t> heNodeEll;
  std::vector<int32_t> sliceOffNode;
  int64_t maxTileEllNode = 0;

  int64_t maxTileNodes = 0, maxTileStaged = 0, maxTileEdges = 0,
          maxTileHalf = 0;
  int64_t totalHalo = 0;
};

/* throws std::runtime_error on invalid input or exceeded limits */
void build_mesh_plan(const MeshInput& in, MeshPlan& out);

struct Graph
{
  int numDof = 1;
  int kind = NW_LINSYS_HYPRE;
  int ndim = 3;
  int64_t iLower = 0, iUpper = -1; /* inclusive, scaled by numDof */
  int64_t numRowsOwned = 0, nnzOwned = 0, numRowsShared = 0, nnzShared = 0;
  std::vector<int64_t> rowStartOwned, rowStartShared;
  std::vector<int64_t> cols, rows;
  std::vector<int64_t> rowIndicesShared;
  std::vector<int64_t> periodicRowsOwned;
  std::vector<int64_t> skippedRows; /* sorted */

  int block = 2; /* rows of the per-edge block */
  /* edge->slot map in caller edge order: [nEdges][block][block], -1 = none */
  std::vector<int64_t> edgeSlots;
  std::vector<int64_t> edgeRhsRows; /* [nEdges][block] */

  /* unified local row space: owned rows then shared rows */
  int64_t numRowsLocal() const { return numRowsOwned + numRowsShared; }
  /* value offset of local row r */
  int64_t rowPtr(int64_t r) const
  {
    return r < numRowsOwned ? rowStartOwned[r]
                            : nnzOwned + rowStartShared[r - numRowsOwned];
  }
  int64_t rowLen(int64_t r) const
  {
    return r < numRowsOwned
             ? rowStartOwned[r + 1] - rowStartOwned[r]
             : rowStartShared[r - numRowsOwned + 1] -
                 rowStartShared[r - numRowsOwned];
  }
  /* local row of a global row id, or -1 (skipped shared rows are absent) */
  int64_t localRow(int64_t hid) const;
};

void build_graph(
  const MeshPlan& mp,
  int kind,
  int numDof,
  const std::vector<int64_t>& skippedRows,
  Graph& g);

struct LsPlan
{
  std::vector<LsTileHdr> tiles;
  std::vector<EntInfo> entInfo;
  std::vector<int32_t> entRhsRow; /* local row (index into rhs) per tile ent */
  std::vector<int32_t> entGo;     /* offset of the row's first value in values[] */
  std::vector<uint32_t> he;       /* flat, sorted by (row, k, edge) */
  std::vector<int32_t> warpSplit;
  std::vector<uint32_t> heEll;    /* sliced-ELL form read by the kernels */
  std::vector<int32_t> sliceOff;
  std::vector<Run> runs;
  /* local rows no tile writes (Dirichlet / periodic-slave / untouched rows):
   * zeroed (periodic: diag 1) by the row-init kernel */
  std::vector<int32_t> uncoveredRows;
  int64_t maxTileNnz = 0, maxTileEnts = 0, maxTileHalf = 0, maxTileEll = 0,
          maxTileRuns = 0;
  bool usable = true;       /* false: tile path impossible, use atomics */
  std::string whyNot;
};

/* 1-dof graphs only (scalar / continuity / UVW momentum) */
void build_ls_plan(const MeshPlan& mp, const Graph& g, LsPlan& lp);

/* sliced-ELL transposition of one tile's sorted half-edge list (nEnts
 * entities): appends the records to ell (padding records are 0) and the
 * nSlices+1 slice offsets (relative to the tile's block) to sliceOff; returns
 * the number of records appended (a multiple of 32) */
int64_t append_sliced_ell(
  const uint32_t* he, int n, int nEnts, std::vector<uint32_t>& ell,
  std::vector<int32_t>& sliceOff);

/* balanced split of a sorted half-edge list among warps at entity boundaries */
void split_half_edges(
  const uint32_t* he, int n, int nWarps, int32_t* split /* nWarps+1 */);

} // namespace nw

#endif
