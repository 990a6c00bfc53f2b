/*
 * nw_internal.h -- handle definitions behind the opaque C-ABI types.
 */
#ifndef NW_INTERNAL_H
#define NW_INTERNAL_H

#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "nw_kernels.cuh"
#include "plan.h"

namespace nw {

void set_error(const std::string& m);

/* device buffer with RAII */
struct DevBuf
{
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes)
  {
    o.p = nullptr;
    o.bytes = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t alloc(size_t n)
  {
    release();
    if (n == 0)
      n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess)
      bytes = n;
    return e;
  }
  template <class T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};

/* NCCL, bound at run time (dlopen) so the library loads on hosts without it */
struct NcclApi;
struct Comm
{
  void* comm = nullptr; /* ncclComm_t */
  int nranks = 1, rank = 0;
};

} // namespace nw

/* Peer-memory mailbox over NVLink (one process per GPU, CUDA IPC): every rank
 * owns a double-buffered receive window and one flag word per peer; a halo
 * exchange is then two kernels -- push (pack + remote stores into the peers'
 * windows + release-store of the epoch into their flag words) and pull
 * (acquire-wait on the own flag words + ordered accumulate) -- instead of
 * pack / ncclSend+ncclRecv / unpack.  Set up once at nw_ctx_comm_init; any
 * failure (IPC not permitted, window too small) leaves the NCCL path in use on
 * ALL ranks (the decision is agreed with an all-reduce). */
struct nw_p2p
{
  bool ok = false;
  int64_t winDoubles = 0; /* doubles per window slot (three slots) */
  nw::DevBuf window;      /* [3][winDoubles] */
  nw::DevBuf flags;       /* unsigned long long [nranks]: epoch written by rank r */
  nw::DevBuf sync;        /* [0] block counter, [1] timeout/error word */
  std::vector<void*> mappedWindow, mappedFlags; /* per rank, opened IPC handles */
  nw::DevBuf dPeerWindow, dPeerFlags; /* device arrays of those pointers */
  unsigned long long epoch = 0;
  /* Flags are exchanged with the UNION of the neighbours of every exchange
   * object registered on this context (p2p_register_peers): an exchange
   * signals and waits for all of them whether or not it carries data for
   * them, so the epoch / window-parity protocol holds for objects with
   * different neighbour sets (general RCB / graph partitions). */
  std::vector<int32_t> unionPeers; /* ascending */
  nw::DevBuf dUnionPeers;          /* int32 [nranks] */
  long long timeoutCycles = 40000000000ll; /* NW_P2P_TIMEOUT_S, default 20 s */
  unsigned* hErr = nullptr; /* pinned host copy of sync[1] (p2p_queue_error_read) */
  /* asynchronous completion (default, NW_P2P_ASYNC=0 switches it off): pushes
   * run on the compute stream (fused into the producing kernel, or a push
   * kernel), pulls on `commStream` beside whatever the compute stream does
   * next.  An object with a pull in flight carries its completion event; every
   * later use of the object first makes the compute stream wait for it, and so
   * does every later push kernel (p2p_next has the protocol; a kernel with a
   * fused push waits for the pull two exchanges back only). */
  bool async = false;
  cudaStream_t commStream = nullptr;
  cudaEvent_t pushDone = nullptr; /* compute stream: producer + push issued */
  cudaEvent_t lastPull = nullptr; /* completion event of the latest pull (not owned) */
  /* "the pull of epoch e is done", slot e mod 3 (p2p_next: a kernel with a
   * fused push waits for the pull two exchanges back only) */
  cudaEvent_t pullRing[3] = {nullptr, nullptr, nullptr};
  unsigned long long pullRingEpoch[3] = {0, 0, 0};
  /* a linear system whose shared rows were pushed from its assembly call
   * (eager exchange) and not yet pulled: the next exchange of the context
   * completes it first (the window protocol wants pull(e) before push(e+1)) */
  struct nw_linsys* pendingEager = nullptr;
};

struct nw_ctx
{
  int device = -1; /* < 0: host-only context (plan building, no compute) */
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr; /* nw_field_stage: H2D beside the compute */
  nw::Comm comm;
  nw_p2p p2p;
  bool skipExchange = false; /* nw_debug_skip_exchange (measurement aid) */
};

struct nw_field_t
{
  std::string name;
  int rank = NW_NODE;
  int ncomp = 1;
  int64_t stride = 0; /* entities incl. padding */
  nw::DevBuf buf;
  /* nw_field_stage / nw_field_commit */
  nw::DevBuf staging;
  cudaEvent_t staged = nullptr;   /* copy stream: staging holds the new data */
  cudaEvent_t consumed = nullptr; /* compute stream: staging may be reused */
  bool stagePending = false;
  /* peer-memory halo sum in flight on the communication stream */
  cudaEvent_t pullDone = nullptr;
  bool pullPending = false;
  ~nw_field_t()
  {
    if (pullDone)
      cudaEventDestroy(pullDone);
    if (staged)
      cudaEventDestroy(staged);
    if (consumed)
      cudaEventDestroy(consumed);
  }
};

/* owner-side accumulation plan: distinct destinations, each with the buffer
 * positions that add into it, in ascending (peer, entry) order */
struct nw_accum_plan
{
  int64_t nDst = 0;
  nw::DevBuf dDst, dPtr, dPos; /* int64 */
};

/* neighbour exchange lists of one mesh (nodal fields) */
struct nw_node_halo
{
  bool built = false;
  std::vector<int> peers; /* ascending rank */
  /* as a sharer: my non-owned nodes grouped by owner */
  std::vector<std::vector<int32_t>> ghostSlots; /* per peer: internal slots */
  std::vector<std::vector<int64_t>> ghostHids;  /* per peer: row ids sent */
  /* as the owner: my owned nodes other ranks hold copies of */
  std::vector<std::vector<int32_t>> ownedSlots; /* per peer */
  std::vector<nw::DevBuf> dGhostIdx, dOwnedIdx; /* int64 slot lists */
  /* all peers concatenated (ascending peer): one launch per exchange phase */
  nw::DevBuf dGhostAll, dOwnedAll;
  std::vector<int64_t> ghostOff, ownedOff; /* per peer offsets, +1 total */
  nw_accum_plan ownedAccum;
  nw::DevBuf sendBuf, recvBuf;
  /* peer-memory path.  p2pMode 1 (every shared node has exactly two sharers):
   * per peer send [ghosts owned by it | my owned nodes it ghosts], receive the
   * mirror, both sides add (one epoch).  p2pMode 2 (any number of sharers):
   * ghosts -> owner with the ordered accumulate, then owner -> ghosts (two
   * epochs), same summation order as the NCCL path. */
  bool p2p = false;
  int p2pMode = 0;
  int64_t nSendA = 0, nSendB = 0, nRecvB = 0;
  nw::DevBuf dSendIdxA, dSendDstA, dSendIdxB, dSendDstB, dRecvIdxB; /* int64 */
  nw::DevBuf dSendPeerA, dSendPeerB;                                /* int32 */
  nw::DevBuf dRecvAllGhost;                                         /* uint8, all 1 */
  int64_t nSendP2p = 0, nRecvP2p = 0;
  nw::DevBuf dSendIdx, dSendDst, dRecvIdx; /* int64 */
  nw::DevBuf dSendPeer, dPeerList;         /* int32 */
  nw::DevBuf dRecvIsGhost;                 /* uint8: receive entry is a ghost of mine */
  /* fused push of the gradient kernels (NodePushDev): the first send list of
   * the mode grouped by tile */
  int nPushTiles = 0;
  nw::DevBuf dPushTilePtr, dPushSlot, dPushPeer, dPushDst;
};

struct nw_mesh;
/* a nodal halo sum between its two halves (nw_halo.inc) */
struct NodeHaloSum
{
  nw_mesh* mesh = nullptr;
  nw_field_t* f = nullptr;
  nw::P2pDev pp;
  int mode = 0; /* 0: nothing sent yet (NCCL path / exchange skipped) */
  bool signal = false; /* the push was fused into the producing kernel: the
                          first pull kernel publishes the epoch */
};

struct nw_ls_shared;
struct nw_mesh
{
  nw_ctx* ctx = nullptr;
  nw::MeshPlan plan;
  nw::MeshPlanDev dev;
  nw::DevBuf dTiles, dHalo, dHaloBlock, dLr, dHeNode, dWarpNode, dPrimary, dNodeOfSlot,
    dTileEdgeSrc, dPrimarySlot, dSecondSlot;
  nw::DevBuf scratch; /* staging for field upload / download */
  /* mass_flow_rate + mass_vof_balanced_flow_rate over the tile-edge slots: the
   * mdot stream of the VOF momentum kernels (nw_momentum_opts::has_vof) */
  nw::DevBuf dVofMdot;
  std::vector<std::unique_ptr<nw_field_t>> fields;
  std::map<std::string, int> fieldByName;
  nw_node_halo halo;
  /* own row id - iLowerNode -> local node, -1 = none (multi-rank meshes only) */
  std::vector<int32_t> ownedNodeOfHid;
  /* GeometryInteriorAlg tables of the last element block of each topology
   * (hex8, quad4, tet4, wed6, pyr5; cached: a moving mesh calls every step
   * with the same connectivity) */
  struct GeoCache
  {
    int64_t nElems = -1;
    uint64_t hash = 0;
    nw::DevBuf dElemSlots, dElemEdges, dOwned;
    bool hasOwned = false;
  } geo[5];
  /* periodic row groups (nodes sharing one node_hypre_id, more than one
   * member): CSR over internal slots, master first, slaves in ascending own id;
   * perMasterMissing: a group without its master on this rank exists */
  std::vector<int32_t> perPtr, perSlots;
  bool perMasterMissing = false;
  bool perUploaded = false;
  nw::DevBuf dPerPtr, dPerSlots;
  /* node-kernel selector: locally owned and not a periodic slave */
  std::vector<uint8_t> nodeKernelActive;
  /* slots of the selected nodes on the device (udiag post-processing) */
  nw::DevBuf dActiveSlots;
  int64_t nActiveSlots = -1;
  int64_t planBytes = 0;
  /* finalized graphs / plans of this mesh's linear systems (see nw_ls_shared) */
  std::vector<std::shared_ptr<nw_ls_shared>> lsCache;
};

enum nw_ls_state { NW_LS_UNSET = 0, NW_LS_LAZY_ZERO = 1, NW_LS_ACCUM = 2 };

/* graph + reduction plan of a linear system.  Systems of one mesh with the same
 * dofs per node and the same skipped rows (continuity, TKE, SDR, the UVW
 * momentum system: all node-row graphs of the same edge list) share one
 * instance, host and device side -- HypreLinearSystem builds the same graph
 * once per equation system (src/HypreLinearSystem.C:412-478). */
struct nw_ls_shared
{
  int numDof = 1;
  std::vector<int64_t> skipped; /* sorted, unique: the cache key */
  nw::Graph g;
  nw::LsPlan lp;
  bool uploaded = false;
  nw::DevBuf dLsTiles, dEntInfo, dEntRhsRow, dHe, dWarp, dRuns;
  nw::DevBuf dUncovered, dUncoveredPeriodic, dRowPtr;
  nw::DevBuf dPeriodicRows;
};

struct nw_linsys
{
  nw_mesh* mesh = nullptr;
  std::shared_ptr<nw_ls_shared> sh;
  int kind = NW_LINSYS_HYPRE;
  int numDof = 1;
  int nRhs = 1;
  bool graphBuilt = false, finalized = false;
  std::vector<int64_t> skipped;
  nw::LsPlanDev dev;
  int mode = NW_SCATTER_SEGMENTED;
  int state = NW_LS_UNSET;

  nw::DevBuf dValues, dRhs;
  /* atomic variant maps (lazy) */
  bool atomicBuilt = false;
  nw::DevBuf dASlots, dARhsRows;
  /* device copy of the graph for nw_linsys_sum_into (lazy) */
  bool devGraphBuilt = false;
  nw::DevBuf dRowStartOwned, dRowStartShared, dRowIndicesShared, dCols,
    dSkipped, dNodeHid;
  nw::DevBuf dNormPartial, dNormOut;
  /* node-kernel scatter table (lazy): int64[n][4] = slot, diagonal value
   * offset, rhs row, dof (-1: UVW, all rhs columns) */
  bool nodeRowsBuilt = false;
  int64_t nNodeRows = 0;
  nw::DevBuf dNodeRows;

  /* shared-row halo (multi-rank) */
  struct Peer
  {
    int rank = -1;
    /* send: contiguous segment of my shared tail */
    int64_t sendRow0 = 0, sendRows = 0, sendVal0 = 0, sendVals = 0;
    /* recv: structure received at finalize, destination slots in my arrays */
    std::vector<int64_t> recvRows, recvRowLens, recvCols;
    std::vector<int64_t> recvValSlot, recvRhsRow;
    nw::DevBuf dRecvValSlot, dRecvRhsRow;
    int64_t recvVals = 0, recvNRows = 0;
  };
  std::vector<Peer> peers;
  bool haloBuilt = false;
  nw::DevBuf haloRecv;
  /* receive layout: [all peers' values | rhs column 0 rows | column 1 ...] */
  int64_t recvValTotal = 0, recvRowTotal = 0;
  nw_accum_plan valAccum, rhsAccum;
  /* peer-memory path: my tail segments -> the owners' windows */
  bool p2p = false;
  cudaEvent_t pullDone = nullptr; /* shared-row add in flight (communication stream) */
  bool pullPending = false;
  std::vector<int64_t> p2pPeerInfo; /* per peer: voff, roff, valTotal, rowTotal */
  const double* p2pBuiltFor = nullptr; /* dev.values the segment table was built for */
  int p2pNSeg = 0;
  int64_t p2pTotal = 0;
  nw::DevBuf dSegSrc, dSegStart, dSegDst, dSegPeer, dPeerList;
  /* columns received for owned rows that the local graph does not have:
   * (row, col) pairs appended after the reference-layout arrays */
  int64_t nExtra = 0;
  std::vector<int64_t> extraRows, extraCols;
  /* monolithic ndim-dof system on the tile path: the node graph's plan (the
   * 1-dof twin, shared with the scalar systems of the mesh) plus, per tile
   * row, the value offset and the local row of the node's first dof */
  std::shared_ptr<nw_ls_shared> twin;
  bool monoOk = false;
  std::vector<int32_t> monoGo, monoRow, monoUncovered;
  std::vector<uint8_t> monoUncoveredPer;
  nw::DevBuf dMonoGo, dMonoRow, dMonoUncovered, dMonoUncoveredPer;
  /* eager exchange (nw_linsys_set_eager_exchange): the tile assembly runs the
   * tiles that own shared or receiving rows first, pushes the shared tail and
   * assembles the interior tiles behind the push; load_complete only pulls.
   * eagerState: 0 nothing outstanding, 1 pushed (pull outstanding), 2 pulled
   * by a later exchange of the context (load_complete is a no-op) */
  bool eager = false;
  bool eagerFused = false; /* the outstanding push was done by the tile kernel */
  int eagerState = 0;
  nw::P2pDev eagerPP;
  int nSendTiles = 0;      /* tiles with rows of the shared tail */
  bool fusedPushOk = false;
  int nPushSeg = 0;
  bool p2pNothingToSend = false; /* no shared rows at all on this rank */
  nw::DevBuf dPushSeg;     /* PushSeg per owner */
};

#endif
