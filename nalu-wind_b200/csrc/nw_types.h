/*
 * nw_types.h -- internal plain-data types shared by the host plan builder, the
 * CUDA kernels and the CPU plan emulator.
 */
#ifndef NW_TYPES_H
#define NW_TYPES_H

#include <stdint.h>

#include "nalu_edge_b200.h"

namespace nw {

/* ---- per-tile limits baked into the packed encodings ---- */
constexpr int kMaxTileEdges = 4096;     /* 12 bits in a half-edge record */
constexpr int kMaxTileEnts = 1024;      /* 10 bits */
constexpr int kMaxRowNnz = 128;         /* 7 bits: slot of an entry in its row */
constexpr int kMaxTileStaged = 65535;   /* 16 bits per end in an lr record */
constexpr int kMaxWarps = 8;            /* CTA = 256 threads */
constexpr int kTileThreads = 256;
constexpr int kHaloBlock = 256;         /* halo slots per tile stored at a fixed stride */

/* half-edge record (32 bit):
 *   [0,12)  tile-edge index
 *   [12]    side: 0 = this entity is the edge's L node, 1 = R node
 *   [13,20) k: position of the off-diagonal entry inside the entity's row
 *   [20,30) entity-local index (row or node within the tile)
 *   [30]    dup: an earlier half-edge of this row hits the same (row, k) slot
 *           (periodic aliases); the first one of a group is not marked, so
 *           "dup ? accumulate : store" writes every staged slot without a
 *           zero fill
 *   [31]    valid
 * The lists are kept in two forms: the flat row-sorted list (host side: plan
 * checks, CPU walk-through) and a sliced-ELL transposition for the device --
 * slices of 32 entities, slice s stored as W_s x 32 records so that lane i of
 * a warp reads the w-th half-edge of entity 32 s + i with one coalesced
 * load; sliceOff[s] is the record offset of slice s inside the tile's block
 * (sliceOff[s+1] - sliceOff[s] = 32 W_s). */
constexpr uint32_t kHeValid = 0x80000000u;
constexpr uint32_t kHeDup = 0x40000000u;
inline uint32_t
he_pack(uint32_t edge, uint32_t side, uint32_t k, uint32_t ent, bool dup)
{
  return edge | (side << 12) | (k << 13) | (ent << 20) | (dup ? kHeDup : 0u) |
         kHeValid;
}
#if defined(__CUDACC__)
#define NW_HDI __host__ __device__ __forceinline__
#else
#define NW_HDI inline
#endif
NW_HDI uint32_t he_edge(uint32_t h) { return h & 0xfffu; }
NW_HDI uint32_t he_side(uint32_t h) { return (h >> 12) & 1u; }
NW_HDI uint32_t he_k(uint32_t h) { return (h >> 13) & 0x7fu; }
NW_HDI uint32_t he_ent(uint32_t h) { return (h >> 20) & 0x3ffu; }

/* Tile header: mesh part (node / edge staging). 64 bytes. */
struct TileHdr
{
  int32_t node0;      /* first internal node slot (even) */
  int32_t nOwn;       /* owned nodes (slots node0 .. node0+nOwn) */
  int32_t nOwnPad;    /* nOwn rounded up to even (TMA: 16-byte granules) */
  int32_t nHalo;      /* staged non-owned nodes */
  int32_t haloPtr;    /* offset into haloNodes[] */
  int32_t edge0;      /* first tile-edge slot (even) */
  int32_t nEdges;     /* tile-edges (internal + cut) */
  int32_t nHalfNode;  /* node-keyed half-edges (gradient) */
  int32_t hePtrNode;  /* offset into heNode[] (flat list, host side) */
  int32_t warpPtrNode;/* offset into warp split table (kMaxWarps+1 entries) */
  int32_t ellPtrNode; /* offset into heNodeEll[] (multiple of 32) */
  int32_t ellLenNode; /* records incl. padding (multiple of 32) */
  int32_t slicePtrNode; /* offset into sliceOffNode[]: nSlices+1 entries */
  int32_t pad[3];
};

/* Tile header: linear-system part. 64 bytes. */
struct LsTileHdr
{
  int32_t nEnts;      /* rows handled by this tile */
  int32_t entPtr;     /* offset into entInfo[] / entRhsRow[] */
  int32_t nnz;        /* staged matrix values for this tile */
  int32_t nHalf;      /* row-keyed half-edges */
  int32_t hePtr;      /* offset into he[] */
  int32_t warpPtr;    /* offset into warp split table */
  int32_t runPtr;     /* offset into runs[] */
  int32_t nRuns;
  int32_t ellPtr;     /* offset into heEll[] (multiple of 32) */
  int32_t ellLen;     /* records incl. padding (multiple of 32) */
  int32_t slicePtr;   /* offset into sliceOff[]: nSlices+1 entries */
  int32_t hasShared;  /* 1: some row of this tile lies in the shared tail
                         (another rank owns it): fused push of the tile kernel */
  int32_t pad[4];
};

/* per tile row: where its values sit in the staging buffer */
struct EntInfo
{
  uint16_t base; /* staging offset of the row's first value */
  uint8_t diagK; /* position of the diagonal inside the row */
  uint8_t nnz;   /* row length (<= kMaxRowNnz) */
};

/* contiguous copy-out segment: staging[so .. so+len) -> values[go ..] */
struct Run
{
  int64_t go;
  int32_t so;
  int32_t len;
};

} // namespace nw

#endif
