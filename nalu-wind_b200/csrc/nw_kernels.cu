/*
 * nw_kernels.cu -- hand-written sm_100a kernels of the edge assembly path.
 *
 * Tile kernels (default, deterministic):  one CTA per locality tile.
 *   stage : the tile's own node range of every SoA field component is pulled
 *           into shared memory by TMA bulk copies (cp.async.bulk + mbarrier
 *           complete_tx); the halo nodes of the tile are gathered by the
 *           threads (they hit L2: neighbouring tiles run close in time);
 *   phase1: one thread per tile-edge evaluates the edge physics from shared
 *           memory (edge data is streamed coalesced from HBM) and leaves a
 *           compact per-edge result in shared memory;
 *   phase2: row-sorted reduction, one thread per matrix row: the row's
 *           half-edges come from the sliced-ELL list (one coalesced record per
 *           lane and step, staged by a TMA bulk copy), diagonal and rhs are
 *           summed in registers in list order (deterministic), every
 *           off-diagonal is stored once into the row staging buffer.  (Round-1
 *           profile: the segmented warp-shuffle scan this replaces was 45 % of
 *           all issued instructions, profiles/r01a_*; it survives only as the
 *           warp aggregation of the atomic comparison variant);
 *   phase3: each warp copies the staging range of the 32 rows it has just
 *           reduced to the CSR value array, element-wise (coalesced wherever
 *           the rows are consecutive) -- each matrix
 *           value and rhs entry is written exactly once, no atomics, no
 *           zero-fill pass (replaces resetCoeffApplierData's deep_copy(0) and
 *           the column walk + atomic_add of sum_into,
 *           src/HypreLinearSystem.C:1386-1430, 2165-2239).
 * Atomic kernels (comparison variant): same tiles, direct global gathers,
 *   warp-aggregated fp64 atomicAdd through a precomputed integer slot map.
 */
#include "nw_kernels.cuh"

#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "edge_physics.h"

namespace nw {

namespace {

constexpr unsigned kFull = 0xffffffffu;

/* ------------------------------------------------------------------ */
/*  optional per-phase cycle accounting (diagnostic build only:        */
/*  make prof -> libnalu_edge_b200_prof.so, tools/phase_times.py)      */
/* ------------------------------------------------------------------ */
#ifdef NW_PHASE_TIMING
__device__ unsigned long long g_phase[kPhaseKernels][kPhaseSlots];
__device__ __forceinline__ long long
phase_clock()
{
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
struct PhaseTimer
{
  int kid;
  int slot = 0;
  bool me;
  long long last;
  __device__ __forceinline__ explicit PhaseTimer(int k, int who = 0)
    : kid(k), me((int)threadIdx.x == who)
  {
    last = phase_clock();
  }
  /* the observing thread: add the cycles since the previous mark to the next slot */
  __device__ __forceinline__ void mark()
  {
    if (me) {
      const long long t = phase_clock();
      atomicAdd(&g_phase[kid][slot], (unsigned long long)(t - last));
      last = t;
    }
    ++slot;
  }
  /* persistent kernels: restart the slot sequence for the next tile */
  __device__ __forceinline__ void lap()
  {
    slot = 0;
    if (me)
      atomicAdd(&g_phase[kid][kPhaseSlots - 1], 1ull);
  }
  __device__ __forceinline__ void done()
  {
    if (me)
      atomicAdd(&g_phase[kid][kPhaseSlots - 1], 1ull);
  }
};
#define NW_PT_BEGIN(k) PhaseTimer pt_(k)
#define NW_PT_BEGIN_AT(k, who) PhaseTimer pt_(k, who)
#define NW_PT_MARK() pt_.mark()
#define NW_PT_LAP() pt_.lap()
#define NW_PT_END() pt_.done()
#else
#define NW_PT_BEGIN(k) (void)(k)
#define NW_PT_BEGIN_AT(k, who) (void)(k)
#define NW_PT_MARK()
#define NW_PT_LAP()
#define NW_PT_END()
#endif

/* ------------------------------------------------------------------ */
/*  TMA bulk copy + mbarrier (PTX)                                     */
/* ------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t
smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
  /* make the initialised barrier visible to the async (TMA) proxy */
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile(
    "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
    "r"(bytes)
    : "memory");
}

__device__ __forceinline__ void
mbar_wait(uint64_t* bar, uint32_t parity)
{
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  } while (!done);
}

/* 1-D bulk copy global -> shared, completion signalled on an mbarrier.
 * bytes % 16 == 0, both addresses 16-byte aligned.  SASS: UBLKCP. */
__device__ __forceinline__ void
tma_load_1d(void* dstSmem, const void* srcGmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
    "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)),
    "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar))
    : "memory");
}

__device__ __forceinline__ void
cp_async4(void* dstSmem, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dstSmem)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void
cp_async8(void* dstSmem, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dstSmem)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void
cp_async16(void* dstSmem, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dstSmem)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void
cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void
cp_async_wait_all()
{
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ uint32_t
round16(uint32_t bytes)
{
  return (bytes + 15u) & ~15u;
}

/* Issue of a tile's bulk copies: copy q is issued by lane 0 of warp
 * q mod nWarps.  A bulk copy costs its issuing WARP ~100 cycles (UBLKCP is a
 * uniform-datapath instruction: lanes of one warp take turns, profiles/
 * r02c_phase_cycles_tile192.txt: 28 copies from one warp = 2.4-3 k cycles per
 * tile, whichever lanes issue them), while the TMA unit accepts copies from
 * different warps back to back (r02b_microbench_stage_layouts.txt, V4). */
template <class F>
__device__ __forceinline__ void
issue_spread(int nCopies, F&& copy, int mode = 0)
{
  if (mode & 32) { /* experiment: all copies from thread 0 */
    if (threadIdx.x == 0)
      for (int q = 0; q < nCopies; ++q)
        copy(q);
    return;
  }
  if ((threadIdx.x & 31) == 0)
    for (int q = threadIdx.x >> 5; q < nCopies; q += blockDim.x >> 5)
      copy(q);
}

/* Stage NC node components of one tile: own range by TMA, halo by gather.
 * s_node[c*stride + i]; i < nOwnPad own, nOwnPad + k halo k.  The caller has
 * initialised `bar` (count 1), posted the expected bytes and synchronised the
 * CTA.  Copy q < NC of the tile's copy table. */
template <int NC>
__device__ __forceinline__ void
node_copy(
  int q, double* s_node, int stride, const NodeComps& nc, const TileHdr& h,
  uint64_t* bar, bool skip = false)
{
  const uint32_t bytes = skip ? 0u : (uint32_t)h.nOwnPad * 8u;
  if (bytes)
    tma_load_1d(s_node + q * stride, nc.c[q] + h.node0, bytes, bar);
}
__device__ __forceinline__ uint32_t
node_copy_bytes(int ncomp, const TileHdr& h, bool skip = false)
{
  return skip ? 0u : (uint32_t)h.nOwnPad * 8u * (uint32_t)ncomp;
}

/* halo nodes: asynchronous 8-byte copies global -> shared (LDGSTS: one
 * load/store-unit operation per value instead of a load and a store, no
 * register round trip); the issuing thread waits for its own copies with
 * stage_halo_wait() before the CTA barrier */
/* the calling thread's halo slot of this tile from the fixed-stride block
 * (issue it right after the header load: both travel together) */
__device__ __forceinline__ int32_t
halo_index_early(const MeshPlanDev& mp, int tile)
{
  return (int)threadIdx.x < kHaloBlock
           ? __ldg(mp.haloBlock + (size_t)tile * kHaloBlock + threadIdx.x)
           : -1;
}
__device__ __forceinline__ int32_t
halo_index_early(const MeshPlanDev& mp)
{
  return halo_index_early(mp, (int)blockIdx.x);
}

/* ---- fused push (eager exchange): a boundary tile stores what other ranks
 * need straight into their windows and is done -- no fence, no counter: a
 * system fence per tile holds the tile's SM slot for an NVLink round trip
 * (measured at 512^3 over 8 GPUs, 14.6 k boundary tiles per rank:
 * +1.5 ms per sweep, profiles/r02v_bench_n8_fence_per_tile.json).  The
 * stores have landed when the kernel has completed; the epoch is published
 * by the first block of the pull kernel that follows on the stream
 * (p2p_signal_then_wait). ---- */
__device__ __forceinline__ const PushSeg*
push_seg_of_value(const LsPushDev& pd, int64_t k /* index in the shared tail */)
{
  for (int i = 0; i < pd.nSeg; ++i)
    if (k >= pd.seg[i].val0 && k < pd.seg[i].val0 + pd.seg[i].valN)
      return pd.seg + i;
  return nullptr;
}
__device__ __forceinline__ const PushSeg*
push_seg_of_row(const LsPushDev& pd, int64_t r /* row index in the shared tail */)
{
  for (int i = 0; i < pd.nSeg; ++i)
    if (r >= pd.seg[i].row0 && r < pd.seg[i].row0 + pd.seg[i].rowN)
      return pd.seg + i;
  return nullptr;
}


template <int NC>
__device__ __forceinline__ void
stage_halo_gather(
  double* s_node,
  int stride,
  const NodeComps& nc,
  const TileHdr& h,
  const int32_t* __restrict__ haloNodes,
  int32_t g0 /* halo_index_early() */)
{
  /* halo node k < kHaloBlock by thread k, its index is already here */
  if ((int)threadIdx.x < h.nHalo && (int)threadIdx.x < kHaloBlock) {
    const int k = threadIdx.x;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      cp_async8(s_node + c * stride + h.nOwnPad + k, nc.c[c] + g0);
  }
  /* the rest (more halo nodes than threads, or a tile with more than
   * kHaloBlock of them): through the list */
  const int32_t* halo = haloNodes + h.haloPtr;
  const int first = (int)blockDim.x < kHaloBlock ? (int)blockDim.x : kHaloBlock;
  for (int k = first + threadIdx.x; k < h.nHalo; k += blockDim.x) {
    const int32_t g = __ldg(halo + k);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      cp_async8(s_node + c * stride + h.nOwnPad + k, nc.c[c] + g);
  }
  cp_async_commit();
}
__device__ __forceinline__ void
stage_halo_wait()
{
  cp_async_wait_all();
}

/* bytes the edge-stream bulk copies of one tile deliver: the packed (L,R)
 * records plus `ncomp` double components (tile-edge runs start on a multiple
 * of 4 records and the arrays are padded, so the rounded sizes stay in bounds) */
__device__ __forceinline__ uint32_t
edge_stream_bytes(const TileHdr& h, int ncomp)
{
  return round16((uint32_t)h.nEdges * 4u) +
         (uint32_t)ncomp * round16((uint32_t)h.nEdges * 8u);
}

/* edge input component c of a kernel: area[0..ND), then mdot, then pecfac
 * (constant indices only: a dynamic index into the by-value kernel parameter
 * would copy the struct to local memory) */
template <int ND>
struct EdgeCompSel
{
  const EdgeComps& ec;
  __device__ __forceinline__ const double* operator()(int c) const
  {
    return c == 0 ? ec.area[0]
           : c == 1 ? ec.area[1]
           : (ND == 3 && c == 2) ? ec.area[2]
           : c == ND ? ec.mdot
                     : ec.pecfac;
  }
};

/* edge-stream copy q of a tile: q == 0 the packed (L,R) records, q - 1 = c the
 * edge component c, which lands at s_edge[c * estride + j] */
template <class F>
__device__ __forceinline__ void
edge_copy(
  int q, uint32_t* s_lr, double* s_edge, int estride, const MeshPlanDev& mp,
  const TileHdr& h, const F& comps, uint64_t* bar)
{
  if (h.nEdges == 0)
    return;
  if (q == 0)
    tma_load_1d(s_lr, mp.lr + h.edge0, round16((uint32_t)h.nEdges * 4u), bar);
  else
    tma_load_1d(
      s_edge + (q - 1) * estride, comps(q - 1) + h.edge0,
      round16((uint32_t)h.nEdges * 8u), bar);
}

template <int NV>
__device__ __forceinline__ void
seg_scan(double (&v)[NV], uint32_t key, int lane)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t k2 = __shfl_up_sync(kFull, key, d);
    const bool take = (lane >= d) && (k2 == key);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double t = __shfl_up_sync(kFull, v[i], d);
      if (take)
        v[i] += t;
    }
  }
}

__device__ __forceinline__ bool
seg_tail(uint32_t key, int lane)
{
  const uint32_t kn = __shfl_down_sync(kFull, key, 1);
  return (lane == 31) || (kn != key);
}

__device__ __forceinline__ int
even_up_i(int v)
{
  return (v + 1) & ~1;
}

/* ------------------------------------------------------------------ */
/*  physics policies                                                   */
/* ------------------------------------------------------------------ */

/* A policy provides
 *   NC    staged node components            NRES  doubles per edge result
 *   NR    rhs columns                       Opts  option struct
 *   compute(ld, l, r, av, mdot, pecfac, o, res)   phase 1
 *   contrib(side, res, diag, off, rhs)            phase 2
 * Node component order for each policy is fixed here and mirrored by
 * nw_api.cu when it fills NodeComps. */

template <int ND>
struct ContinuityP
{
  static constexpr int kND = ND;
  static constexpr int NC = 3 * ND + 3; /* x, u, dpdx, rho, p, udiag */
  static constexpr int kPhaseId = 0;
  static constexpr int kMinBlocks = 3; /* <= 85 registers: 3 CTAs per SM */
  static constexpr int kStreamCtas = 3;
  static constexpr int NRES = 2;
  static constexpr int NR = 1;
  static constexpr bool kNeedsMdot = false;
  static constexpr bool kNeedsPec = false;
  using Opts = nw_continuity_opts;
  template <class LD>
  __device__ __forceinline__ static void
  load(const LD& ld, int i, ContNode<ND>& n)
  {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      n.x[d] = ld(d, i);
      n.u[d] = ld(ND + d, i);
      n.g[d] = ld(2 * ND + d, i);
    }
    n.rho = ld(3 * ND, i);
    n.p = ld(3 * ND + 1, i);
    n.ud = ld(3 * ND + 2, i);
  }
  using Node = ContNode<ND>;
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double, double,
    const Opts& o, double* res)
  {
    continuity_edge<ND>(L, R, av, o, res[0], res[1]);
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double m, double pf,
    const Opts& o, double* res)
  {
    Node L, R;
    load(ld, l, L);
    load(ld, r, R);
    compute_n(L, R, av, m, pf, o, res);
  }
  /* contribution of edge j to the row of its `side` node, read from the
   * phase-1 results s_res[k * rs + j] */
  __device__ __forceinline__ static void contrib(
    uint32_t side, const double* s_res, int rs, int j, double& diag,
    double& off, double* rhs)
  {
    const double f = s_res[j], m = s_res[rs + j];
    diag = -f;
    off = f;
    rhs[0] = side ? m : -m;
  }
  __device__ __forceinline__ static void block(
    const double* res, double& LL, double& LR, double& RL, double& RR,
    double* flux)
  {
    LL = -res[0];
    LR = res[0];
    RL = res[0];
    RR = -res[0];
    flux[0] = res[1];
  }
};

/* WallDistEdgeSolverAlg (src/edge_kernels/WallDistEdgeSolverAlg.C:28-66):
 * lhsfac = asq / axdx, lhs = [[+f, -f], [-f, +f]], no rhs.  Stages the
 * coordinates only. */
struct nw_wall_dist_opts_
{
  int unused;
};
template <int ND>
struct WallDistP
{
  static constexpr int kND = ND;
  static constexpr bool kPairLanes = false;
  static constexpr int NC = ND; /* x */
  static constexpr int kPhaseId = 7;
  static constexpr int kMinBlocks = 3;
  static constexpr int kStreamCtas = 3;
  static constexpr int NRES = 1;
  static constexpr int NR = 1;
  static constexpr bool kNeedsMdot = false;
  static constexpr bool kNeedsPec = false;
  using Opts = nw_wall_dist_opts_;
  struct Node
  {
    double x[ND];
  };
  template <class LD>
  __device__ __forceinline__ static void load(const LD& ld, int i, Node& n)
  {
#pragma unroll
    for (int d = 0; d < ND; ++d)
      n.x[d] = ld(d, i);
  }
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double, double, const Opts&,
    double* res)
  {
    double asq = 0.0, axdx = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const double dxj = R.x[d] - L.x[d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
    }
    /* an exact divide: this kernel is nowhere near any arithmetic limit */
    res[0] = asq / axdx;
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double m, double pf,
    const Opts& o, double* res)
  {
    Node L, R;
    load(ld, l, L);
    load(ld, r, R);
    compute_n(L, R, av, m, pf, o, res);
  }
  __device__ __forceinline__ static void contrib(
    uint32_t, const double* s_res, int, int j, double& diag, double& off,
    double* rhs)
  {
    const double f = s_res[j];
    diag = f;
    off = -f;
    rhs[0] = 0.0;
  }
  __device__ __forceinline__ static void block(
    const double* res, double& LL, double& LR, double& RL, double& RR,
    double* flux)
  {
    LL = res[0];
    LR = -res[0];
    RL = -res[0];
    RR = res[0];
    flux[0] = 0.0;
  }
};

template <int ND>
struct ScalarP
{
  static constexpr int kND = ND;
  static constexpr int NC = 3 * ND + 3; /* x, vrtm, dqdx, q, rho, dflux */
  static constexpr int kPhaseId = 1;
  static constexpr int kMinBlocks = 2;
  static constexpr int kStreamCtas = 2;
  static constexpr int NRES = 5;
  static constexpr int NR = 1;
  static constexpr bool kNeedsMdot = true;
  static constexpr bool kNeedsPec = false;
  using Opts = nw_scalar_opts;
  template <class LD>
  __device__ __forceinline__ static void
  load(const LD& ld, int i, ScalNode<ND>& n)
  {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      n.x[d] = ld(d, i);
      n.v[d] = ld(ND + d, i);
      n.dq[d] = ld(2 * ND + d, i);
    }
    n.q = ld(3 * ND, i);
    n.rho = ld(3 * ND + 1, i);
    n.mu = ld(3 * ND + 2, i);
  }
  using Node = ScalNode<ND>;
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double mdot, double,
    const Opts& o, double* res)
  {
    scalar_edge<ND>(L, R, av, mdot, o, res, res[4]);
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double mdot, double pf,
    const Opts& o, double* res)
  {
    Node L, R;
    load(ld, l, L);
    load(ld, r, R);
    compute_n(L, R, av, mdot, pf, o, res);
  }
  __device__ __forceinline__ static void contrib(
    uint32_t side, const double* s_res, int rs, int j, double& diag,
    double& off, double* rhs)
  {
    /* L row: (a00, a01, -flux); R row: (a11, a10, +flux) */
    const double* p = s_res + j;
    diag = p[side ? 3 * rs : 0];
    off = p[side ? 2 * rs : rs];
    const double f = p[4 * rs];
    rhs[0] = side ? f : -f;
  }
  __device__ __forceinline__ static void block(
    const double* res, double& LL, double& LR, double& RL, double& RR,
    double* flux)
  {
    LL = res[0];
    LR = res[1];
    RL = res[2];
    RR = res[3];
    flux[0] = res[4];
  }
};

template <int ND>
struct MomentumUvwP
{
  /* x, u, dudx, visc, rho, mask */
  static constexpr int kND = ND;
  static constexpr int NC = 2 * ND + ND * ND + 3;
  static constexpr int kPhaseId = 2;
  static constexpr int kMinBlocks = 2; /* 128 registers x 256 threads x 2 */
  static constexpr int kStreamCtas = 2;
  static constexpr int NRES = 4 + ND;
  static constexpr int NR = ND;
  static constexpr bool kNeedsMdot = true;
  static constexpr bool kNeedsPec = true;
  using Opts = nw_momentum_opts;
  template <class LD>
  __device__ __forceinline__ static void
  load(const LD& ld, int i, MomNode<ND>& n)
  {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      n.x[d] = ld(d, i);
      n.u[d] = ld(ND + d, i);
    }
#pragma unroll
    for (int d = 0; d < ND * ND; ++d)
      n.g[d] = ld(2 * ND + d, i);
    n.mu = ld(2 * ND + ND * ND, i);
    n.rho = ld(2 * ND + ND * ND + 1, i);
    n.mask = ld(2 * ND + ND * ND + 2, i);
  }
  using Node = MomNode<ND>;
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Node L, R;
    load(ld, l, L);
    load(ld, r, R);
    compute_n(L, R, av, mdot, pecfac, o, res);
  }
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    compute_n_t<false>(L, R, av, mdot, pecfac, o, res);
  }
  /* VOF: the realm_has_vof_ branch of the physics (MomentumUvwVofP) */
  template <bool VOF>
  __device__ __forceinline__ static void compute_n_t(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    if (o.fuse_peclet) {
      /* MomentumEdgePecletAlg fused (src/edge_kernels/MomentumEdgePecletAlg.C:74-101) */
      PecNode<ND> pl, pr;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        pl.x[d] = L.x[d];
        pr.x[d] = R.x[d];
        pl.v[d] = L.u[d];
        pr.v[d] = R.u[d];
      }
      pl.rho = L.rho;
      pr.rho = R.rho;
      pl.mu = L.mu;
      pr.mu = R.mu;
      pecfac = peclet_eval(o.pf, peclet_number<ND>(pl, pr, o.pec_eps));
    }
    MomResult<ND> m;
    if (VOF)
      momentum_edge_vof<ND>(L, R, av, mdot, pecfac, o, m);
    else
      momentum_edge<ND>(L, R, av, mdot, pecfac, o, m);
    /* HypreUVWLinSysCoeffApplier keeps only the x-x entries
     * (src/HypreUVWLinearSystem.C:738) */
    momentum_block_entry<ND>(
      m, av, o.relax_fac, 0, 0, res[0], res[1], res[2], res[3]);
#pragma unroll
    for (int d = 0; d < ND; ++d)
      res[4 + d] = m.flux[d];
  }
  __device__ __forceinline__ static void contrib(
    uint32_t side, const double* s_res, int rs, int j, double& diag,
    double& off, double* rhs)
  {
    const double* p = s_res + j;
    diag = p[side ? 3 * rs : 0];
    off = p[side ? 2 * rs : rs];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const double f = p[(4 + d) * rs];
      rhs[d] = side ? f : -f;
    }
  }
  __device__ __forceinline__ static void block(
    const double* res, double& LL, double& LR, double& RL, double& RR,
    double* flux)
  {
    LL = res[0];
    LR = res[1];
    RL = res[2];
    RR = res[3];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      flux[d] = res[4 + d];
  }
};

/* shared-memory loader */
struct SmemLd
{
  const double* s;
  int stride;
  __device__ __forceinline__ double operator()(int c, int i) const
  {
    return s[c * stride + i];
  }
};
/* global loader through the read-only path */
struct GmemLd
{
  const NodeComps* nc;
  __device__ __forceinline__ double operator()(int c, int i) const
  {
    return __ldg(nc->c[c] + i);
  }
};

/* ------------------------------------------------------------------ */
/*  linear-system tile kernel                                          */
/* ------------------------------------------------------------------ */

/* shared-memory carve-up of ls_tile_kernel, shared by the kernel and the
 * host-side size computation (all region sizes are multiples of 16 bytes) */
/* Monolithic 3-dof momentum (HypreLinearSystem with numDof = ndim: the full
 * 2ND x 2ND block of MomentumEdgeSolverAlg.C:275-310 through sum_into,
 * src/HypreLinearSystem.C:2059-2161).  Same staging and physics as the UVW
 * policy; an edge keeps the four same-component scalars, the flux, and the
 * three factors of the cross-component term -viscIp a_i a_j / axdx, from which
 * the row walk forms the ND x ND blocks (momentum_block_entry's expression and
 * order).  The reduction runs on the NODE graph's plan (the 1-dof twin of the
 * 3-dof graph: row 3r+i of node r holds the entries of the node row, three
 * columns each) and writes straight to the CSR arrays -- no staging. */
template <int ND>
struct MomentumMonoP : MomentumUvwP<ND>
{
  using Base = MomentumUvwP<ND>;
  using Opts = nw_momentum_opts;
  using Node = MomNode<ND>;
  static constexpr int kMinBlocks = 1;
  static constexpr int NRES = 4 + ND + 2 + ND;
  static constexpr int kSame = 0, kFlux = 4, kVisc = 4 + ND, kInv = 5 + ND, kArea = 6 + ND;
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    compute_n_t<false>(L, R, av, mdot, pecfac, o, res);
  }
  template <bool VOF>
  __device__ __forceinline__ static void compute_n_t(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    if (o.fuse_peclet) {
      PecNode<ND> pl, pr;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        pl.x[d] = L.x[d];
        pr.x[d] = R.x[d];
        pl.v[d] = L.u[d];
        pr.v[d] = R.u[d];
      }
      pl.rho = L.rho;
      pr.rho = R.rho;
      pl.mu = L.mu;
      pr.mu = R.mu;
      pecfac = peclet_eval(o.pf, peclet_number<ND>(pl, pr, o.pec_eps));
    }
    MomResult<ND> m;
    if (VOF)
      momentum_edge_vof<ND>(L, R, av, mdot, pecfac, o, m);
    else
      momentum_edge<ND>(L, R, av, mdot, pecfac, o, m);
    res[0] = m.sLL;
    res[1] = m.sLR;
    res[2] = m.sRL;
    res[3] = m.sRR;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      res[kFlux + d] = m.flux[d];
      res[kArea + d] = av[d];
    }
    res[kVisc] = m.viscIp;
    res[kInv] = m.inv_axdx;
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Node L, R;
    Base::load(ld, l, L);
    Base::load(ld, r, R);
    compute_n(L, R, av, mdot, pecfac, o, res);
  }
  /* entry (0,0) of the diagonal block of this half-edge (extract_diagonal) */
  __device__ __forceinline__ static double diag00(
    uint32_t side, const double* s_res, int rs, int j, const Opts& o)
  {
    const double* p = s_res + j;
    const double ns = -p[kVisc * rs] * p[kArea * rs] * p[kArea * rs] * p[kInv * rs];
    return p[side ? 3 * rs : 0] - ns * nw_rcp(o.relax_fac);
  }
};
template <class P>
struct IsMonoPolicy
{
  static constexpr bool value = false;
};
template <int ND>
struct IsMonoPolicy<MomentumMonoP<ND>>
{
  static constexpr bool value = true;
};

/* realm_has_vof_ (src/edge_kernels/MomentumEdgeSolverAlg.C:88, 174-192): the
 * same kernels with the VOF branch of the physics compiled in, as policies of
 * their own so that the kernels of every other deck stay as they are.  The
 * edge stream they read as mdot is mass_flow_rate +
 * mass_vof_balanced_flow_rate (edge_sum_kernel, nw_assemble_momentum_edge). */
template <int ND>
struct MomentumUvwVofP : MomentumUvwP<ND>
{
  using Base = MomentumUvwP<ND>;
  using Opts = nw_momentum_opts;
  using Node = MomNode<ND>;
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Base::template compute_n_t<true>(L, R, av, mdot, pecfac, o, res);
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Node L, R;
    Base::load(ld, l, L);
    Base::load(ld, r, R);
    compute_n(L, R, av, mdot, pecfac, o, res);
  }
};
template <int ND>
struct MomentumMonoVofP : MomentumMonoP<ND>
{
  using Base = MomentumMonoP<ND>;
  using Opts = nw_momentum_opts;
  using Node = MomNode<ND>;
  __device__ __forceinline__ static void compute_n(
    const Node& L, const Node& R, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Base::template compute_n_t<true>(L, R, av, mdot, pecfac, o, res);
  }
  template <class LD>
  __device__ __forceinline__ static void compute(
    const LD& ld, int l, int r, const double* av, double mdot, double pecfac,
    const Opts& o, double* res)
  {
    Node L, R;
    Base::load(ld, l, L);
    Base::load(ld, r, R);
    compute_n(L, R, av, mdot, pecfac, o, res);
  }
};
template <int ND>
struct IsMonoPolicy<MomentumMonoVofP<ND>>
{
  static constexpr bool value = true;
};

template <class P>
struct LsSmem
{
  int nodeRegion; /* doubles: node stage, reused as row staging after phase 1 */
  int rowRegion;  /* doubles of the row staging inside it */
  int resStride;  /* doubles per result component */
  int valsLen;    /* staged matrix values (doubles) + their int32 deltas */
  int ellLen;     /* uint32 records */
  int entLen;     /* EntInfo / rhs-row / value-offset records */
  int lrLen;      /* packed (L,R) records */
  /* the edge inputs (area, mdot, pecfac) are bulk-copied into the result
   * region: a thread has read edge j's inputs before it writes edge j's
   * results, so the two share storage */
  static constexpr int NIN_MAX =
    P::kND + (P::kNeedsMdot ? 1 : 0) + (P::kNeedsPec ? 1 : 0);
  static constexpr int NEDGE = P::NRES > NIN_MAX ? P::NRES : NIN_MAX;
  __host__ __device__ LsSmem(const MeshPlanDev& mp, const LsPlanDev& lp)
  {
    resStride = (mp.maxTileEdges + 1) & ~1;
    lrLen = (mp.maxTileEdges + 3) & ~3;
    valsLen = (lp.maxTileNnz + 3) & ~3;
    /* row staging: values (8 B) + value-offset deltas (4 B), and behind it the
     * node-keyed half-edge list of extract_diagonal (4 B records); the
     * monolithic policy stages ND x ND values per node-row entry and copies
     * out node by node (no deltas) */
    rowRegion = IsMonoPolicy<P>::value ? valsLen * P::kND * P::kND
                                       : valsLen + valsLen / 2;
    const int rowAll = rowRegion + (mp.maxTileEllNode + 1) / 2 + 2;
    const int stage = P::NC * mp.maxStaged;
    nodeRegion = stage > rowAll ? stage : rowAll;
    nodeRegion = (nodeRegion + 1) & ~1;
    ellLen = lp.maxTileEll;
    entLen = (lp.maxTileEnts + 3) & ~3;
  }
  __host__ __device__ size_t bytes() const
  {
    return sizeof(double) * ((size_t)nodeRegion + (size_t)NEDGE * resStride) +
           4u * (size_t)ellLen + 12u * (size_t)entLen + 4u * (size_t)lrLen;
  }
};

template <class P, int ND, int MINB = P::kMinBlocks, int THREADS = kTileThreads>
__global__ void __launch_bounds__(THREADS, MINB) ls_tile_kernel(
  const MeshPlanDev mp,
  const LsPlanDev lp,
  const NodeComps nc,
  const EdgeComps ec,
  const typename P::Opts o)
{
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar[3];
  /* slice offsets of the tile's sliced-ELL lists (fetched during the stage: a
   * global load at the top of phase 2 would sit on the critical path) */
  __shared__ int32_t s_slice[kMaxTileEnts / 32 + 2];
  __shared__ int32_t s_sliceN[kMaxTileEnts / 32 + 2]; /* node-keyed (diagOut) */
  NW_PT_BEGIN(P::kPhaseId);

  const int tile = (int)blockIdx.x;
  const TileHdr h = mp.tiles[tile];
  const LsTileHdr lh = lp.tiles[tile];
  const int32_t g0 = halo_index_early(mp, tile);
  const LsSmem<P> L(mp, lp);
  const int stride = even_up_i(h.nOwnPad + h.nHalo);

  double* s_node = smem;
  double* s_res = s_node + L.nodeRegion;
  double* s_vals = s_node; /* row staging: the node stage is dead by then */
  int32_t* s_delta = reinterpret_cast<int32_t*>(s_vals + L.valsLen);
  uint32_t* s_ell =
    reinterpret_cast<uint32_t*>(s_res + LsSmem<P>::NEDGE * L.resStride);
  EntInfo* s_ent = reinterpret_cast<EntInfo*>(s_ell + L.ellLen);
  int32_t* s_row = reinterpret_cast<int32_t*>(s_ent + L.entLen);
  int32_t* s_go = s_row + L.entLen;
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(s_go + L.entLen);

  /* edge input streams of this policy: area, [mdot], [pecfac] */
  constexpr int kMdot = ND;
  constexpr int kPec = ND + 1;
  const bool hasPec = P::kNeedsPec && ec.pecfac != nullptr;
  const int nin = ND + (P::kNeedsMdot ? 1 : 0) + (hasPec ? 1 : 0);
  const EdgeCompSel<ND> ecomp{ec};

  /* ---- stage: every contiguous per-tile stream is a TMA bulk copy ---- */
  const uint32_t bEll = (uint32_t)lh.ellLen * 4u;
  const uint32_t bEnt = round16((uint32_t)lh.nEnts * 4u);
  const bool skipOwn = (mp.dbgSkip & 8) != 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    mbar_expect_tx(
      &bar[0], node_copy_bytes(P::NC, h, skipOwn) + edge_stream_bytes(h, nin));
    mbar_expect_tx(&bar[1], bEll + 3u * bEnt);
  }
  __syncthreads();
  NW_PT_MARK(); /* 0: headers + barrier init */
  /* copy table: node components, (L,R) records + edge components, then the
   * reduction plan of phases 2-3 (half-edge records, row layout, rhs rows,
   * value offsets; second barrier: needed only after phase 1, so its latency
   * hides behind the physics) */
  issue_spread(P::NC + 1 + nin + 4, [&](int q) {
    if (q < P::NC)
      node_copy<P::NC>(q, s_node, stride, nc, h, &bar[0], skipOwn);
    else if (q <= P::NC + nin)
      edge_copy(q - P::NC, s_lr, s_res, L.resStride, mp, h, ecomp, &bar[0]);
    else {
      const int r = q - (P::NC + 1 + nin);
      if (r == 0 && bEll)
        tma_load_1d(s_ell, lp.heEll + lh.ellPtr, bEll, &bar[1]);
      else if (r == 1 && bEnt)
        tma_load_1d(s_ent, lp.entInfo + lh.entPtr, bEnt, &bar[1]);
      else if (r == 2 && bEnt)
        tma_load_1d(s_row, lp.entRhsRow + lh.entPtr, bEnt, &bar[1]);
      else if (r == 3 && bEnt)
        tma_load_1d(s_go, lp.entGo + lh.entPtr, bEnt, &bar[1]);
    }
  }, mp.dbgSkip);
  NW_PT_MARK(); /* 1: TMA issue */
  {
    const int nSl = (lh.nEnts + 31) >> 5;
    if ((int)threadIdx.x <= nSl)
      s_slice[threadIdx.x] = __ldg(lp.sliceOff + lh.slicePtr + threadIdx.x);
    const int nSlN = (h.nOwn + 31) >> 5;
    if (lp.diagOut && (int)threadIdx.x <= nSlN)
      s_sliceN[threadIdx.x] =
        __ldg(mp.sliceOffNode + h.slicePtrNode + threadIdx.x);
  }
  if (!(mp.dbgSkip & 1))
    stage_halo_gather<P::NC>(s_node, stride, nc, h, mp.haloNodes, g0);
  NW_PT_MARK(); /* 2: halo gather */
  stage_halo_wait();
  mbar_wait(&bar[0], 0);
  __syncthreads();
  NW_PT_MARK(); /* 3: stage wait */

  /* ---- phase 1: per-edge physics, entirely out of shared memory ---- */
  {
    const SmemLd ld{s_node, stride};
    for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
      const uint32_t v = s_lr[j];
      const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
      double av[ND];
#pragma unroll
      for (int d = 0; d < ND; ++d)
        av[d] = s_res[d * L.resStride + j];
      double mdot = 0.0, pecfac = 0.0;
      if (P::kNeedsMdot)
        mdot = s_res[kMdot * L.resStride + j];
      if (hasPec)
        pecfac = s_res[kPec * L.resStride + j];
      double res[P::NRES];
      if (mp.dbgSkip & 2) {
#pragma unroll
        for (int k = 0; k < P::NRES; ++k)
          res[k] = av[0] + (double)(l + r);
      } else
        P::compute(ld, l, r, av, mdot, pecfac, o, res);
      /* results overwrite this edge's own inputs (all read above) */
#pragma unroll
      for (int k = 0; k < P::NRES; ++k)
        s_res[k * L.resStride + j] = res[k];
    }
  }
  NW_PT_MARK(); /* 4: phase 1 */
  mbar_wait(&bar[1], 0);
  __syncthreads();
  NW_PT_MARK(); /* 5: phase-1 barrier */
  /* extract_diagonal: the node stage is dead now; the tile's node-keyed
   * half-edge list (the gradient kernel's list) lands behind the row staging
   * while phases 2-3 run */
  uint32_t* s_ellN = reinterpret_cast<uint32_t*>(s_node + L.rowRegion);
  if (lp.diagOut && threadIdx.x == 0) {
    const uint32_t bN = (uint32_t)h.ellLenNode * 4u;
    /* the copy overwrites shared memory the threads have just read */
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar[2], bN);
    if (bN)
      tma_load_1d(s_ellN, mp.heNodeEll + h.ellPtrNode, bN, &bar[2]);
  }

  /* ---- phases 2+3, warp by warp: one thread per row reduces the row's
   * half-edges in list order, then the warp copies the staging range of its
   * 32 rows to the CSR arrays (only a warp-level sync in between).  The walk
   * is blocked by four list steps: the records, then the edge results they
   * point to, are loaded before anything of the block is stored, so that the
   * shared-memory latencies of a block overlap (round-1 phase cycles: this
   * phase was a chain of dependent 30-cycle loads, 20 % of a CTA's life). ---- */
  if constexpr (IsMonoPolicy<P>::value) {
    /* ---- monolithic rows: one thread per node walks the node's half-edges
     * and writes its ND rows (ND x ND block per neighbour) in place; s_go is
     * the value offset of row ND r, s_row its local row ---- */
    const double invRelax = nw_rcp(o.relax_fac);
    const int rs = L.resStride;
    const int lane = threadIdx.x & 31;
    for (int row0 = (int)threadIdx.x - lane; row0 < lh.nEnts; row0 += blockDim.x) {
      const int row = row0 + lane;
      if (row < lh.nEnts) {
        const int sl = row0 >> 5;
        const int o0 = s_slice[sl], o1 = s_slice[sl + 1];
        const uint32_t* hp = s_ell + o0 + lane;
        const int W = (o1 - o0) >> 5;
        const EntInfo ei = s_ent[row];
        const int rowLen = ND * (int)ei.nnz;
        double* vbase = s_vals + ND * ND * (int)ei.base; /* staged: ND rows */
        double dg[ND][ND], rhs[ND];
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          rhs[i] = 0.0;
#pragma unroll
          for (int c = 0; c < ND; ++c)
            dg[i][c] = 0.0;
        }
        for (int w = 0; w < W; ++w) {
          const uint32_t hv = hp[w * 32];
          if (!(hv & kHeValid))
            continue;
          const double* p = s_res + he_edge(hv);
          const bool side = he_side(hv) != 0;
          const double sXX = p[side ? 3 * rs : 0];
          const double sXY = p[side ? 2 * rs : rs];
          const double visc = p[P::kVisc * rs], inv = p[P::kInv * rs];
          double a[ND];
#pragma unroll
          for (int d = 0; d < ND; ++d)
            a[d] = p[(P::kArea + d) * rs];
          double* dst = vbase + ND * (int)he_k(hv);
          const bool dup = (hv & kHeDup) != 0;
#pragma unroll
          for (int i = 0; i < ND; ++i) {
            const double f = p[(P::kFlux + i) * rs];
            rhs[i] += side ? f : -f;
#pragma unroll
            for (int c = 0; c < ND; ++c) {
              /* momentum_block_entry (edge_physics.h) */
              const double ns = -visc * a[i] * a[c] * inv;
              const double s = (i == c) ? 1.0 : 0.0;
              dg[i][c] += s * sXX - ns * invRelax;
              double off = s * sXY + ns;
              double* q = dst + i * rowLen + c;
              if (dup)
                off += *q;
              *q = off;
            }
          }
        }
        double* dd = vbase + ND * (int)ei.diagK;
        const int64_t grow = s_row[row];
#pragma unroll
        for (int i = 0; i < ND; ++i) {
#pragma unroll
          for (int c = 0; c < ND; ++c)
            dd[i * rowLen + c] = dg[i][c];
          lp.rhs[grow + i] = rhs[i];
        }
      }
      __syncwarp();
      /* copy-out: the ND rows of a node are contiguous in the CSR arrays */
      const int last = min(row0 + 31, lh.nEnts - 1);
      for (int r = row0; r <= last; ++r) {
        const EntInfo er = s_ent[r];
        const int n = ND * ND * (int)er.nnz;
        const double* src = s_vals + ND * ND * (int)er.base;
        double* dstg = lp.values + s_go[r];
        for (int t = lane; t < n; t += 32)
          dstg[t] = src[t];
      }
    }
  } else {
  /* eager exchange: this tile holds rows of the shared tail */
  const bool pushing = lp.push.seg != nullptr && lh.hasShared != 0;
  if (!(mp.dbgSkip & 4)) {
    const int lane = threadIdx.x & 31;
    for (int row0 = (int)threadIdx.x - lane; row0 < lh.nEnts;
         row0 += blockDim.x) {
      const int row = row0 + lane;
      if (row < lh.nEnts) {
        const int sl = row0 >> 5;
        const int o0 = s_slice[sl], o1 = s_slice[sl + 1];
        const uint32_t* hp = s_ell + o0 + lane;
        const int W = (o1 - o0) >> 5;
        const EntInfo ei = s_ent[row];
        double* vrow = s_vals + ei.base;
        double diag = 0.0;
        double rhs[P::NR];
#pragma unroll
        for (int d = 0; d < P::NR; ++d)
          rhs[d] = 0.0;
        constexpr int kBlk = 4;
        for (int w0 = 0; w0 < W; w0 += kBlk) {
          uint32_t hv[kBlk];
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            hv[u] = (w0 + u < W) ? hp[(w0 + u) * 32] : 0u;
          double dg[kBlk], off[kBlk], rr[kBlk][P::NR];
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            if (hv[u] & kHeValid)
              P::contrib(
                he_side(hv[u]), s_res, L.resStride, (int)he_edge(hv[u]), dg[u],
                off[u], rr[u]);
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            if (hv[u] & kHeValid) {
              diag += dg[u];
#pragma unroll
              for (int d = 0; d < P::NR; ++d)
                rhs[d] += rr[u][d];
              double* dst = vrow + he_k(hv[u]);
              if (hv[u] & kHeDup)
                off[u] += *dst;
              *dst = off[u];
            }
        }
        vrow[ei.diagK] = diag;
        /* value offset of every staged element of this row */
        const int32_t delta = s_go[row] - (int32_t)ei.base;
        for (int k = 0; k < (int)ei.nnz; ++k)
          s_delta[ei.base + k] = delta;
        const int64_t grow = s_row[row];
#pragma unroll
        for (int d = 0; d < P::NR; ++d)
          lp.rhs[(int64_t)d * lp.rhsStride + grow] = rhs[d];
        if (pushing && grow >= lp.push.numRowsOwned) {
          /* a row another rank owns: its rhs goes to the owner's window too */
          const int64_t tr = grow - lp.push.numRowsOwned;
          if (const PushSeg* sg = push_seg_of_row(lp.push, tr)) {
            double* w = lp.push.pp.peerWindow[sg->peer] + lp.push.pp.winOff +
                        sg->rowDst + (tr - sg->row0);
#pragma unroll
            for (int d = 0; d < P::NR; ++d)
              w[(int64_t)d * sg->rowStride] = rhs[d];
          }
        }
      }
      __syncwarp();
      /* copy-out of this warp's rows: every value written exactly once */
      {
        const int last = min(row0 + 31, lh.nEnts - 1);
        const EntInfo e0 = s_ent[row0], e1 = s_ent[last];
        const int end = (int)e1.base + (int)e1.nnz;
#pragma unroll 4
        for (int e = (int)e0.base + lane; e < end; e += 32)
          lp.values[e + s_delta[e]] = s_vals[e];
        if (pushing)
          for (int e = (int)e0.base + lane; e < end; e += 32) {
            const int64_t tk = (int64_t)e + s_delta[e] - lp.push.nnzOwned;
            if (tk >= 0)
              if (const PushSeg* sg = push_seg_of_value(lp.push, tk))
                lp.push.pp.peerWindow[sg->peer]
                                     [lp.push.pp.winOff + sg->valDst + (tk - sg->val0)] =
                  s_vals[e];
          }
      }
    }
  }
  } /* !mono */
  NW_PT_MARK(); /* 6: phases 2+3 */
  if (lp.diagOut) {
    /* NGPApplyCoeff::extract_diagonal (src/SolverAlgorithm.C:87-105):
     * diagField(node) += lhs(ix, ix) for both nodes of every edge, i.e. per
     * node the sum of its half-edges' diagonal-block entries -- node by node
     * (periodic slaves and Dirichlet nodes included: the reference extracts
     * before the row is skipped or redirected), summed in list order, one
     * writer per node instead of the reference's atomic_add */
    mbar_wait(&bar[2], 0);
    for (int i = threadIdx.x; i < h.nOwn; i += blockDim.x) {
      const int sl = i >> 5;
      const int o0 = s_sliceN[sl], o1 = s_sliceN[sl + 1];
      const uint32_t* hp = s_ellN + o0 + (i & 31);
      const int W = (o1 - o0) >> 5;
      double acc = 0.0;
      for (int w = 0; w < W; ++w) {
        const uint32_t hv = hp[w * 32];
        if (hv & kHeValid) {
          if constexpr (IsMonoPolicy<P>::value) {
            acc += P::diag00(he_side(hv), s_res, L.resStride, (int)he_edge(hv), o);
          } else {
            double dg, off, rr[P::NR];
            P::contrib(
              he_side(hv), s_res, L.resStride, (int)he_edge(hv), dg, off, rr);
            acc += dg;
          }
        }
      }
      lp.diagOut[h.node0 + i] += acc;
    }
  }
  NW_PT_END();
}

/* ------------------------------------------------------------------ */
/*  pipelined linear-system kernel: warp-specialised, persistent        */
/* ------------------------------------------------------------------ */

/* profiles/r02d_ablation_timings.txt: in ls_tile_kernel the stage, the physics
 * and the row reduction of a tile add up -- each is a chain of latencies and
 * an SM holds only two or three tiles.  Here one persistent CTA per SM keeps
 * three tiles in flight with different warps:
 *
 *   memory warps (kPipeMemWarps):  reduce tile k (phases 2-3 of
 *       ls_tile_kernel: row walk, copy-out), then stage tile k+2 into the
 *       slot tile k has just left (bulk copies + asynchronous halo gather:
 *       issue only, the data lands while they reduce tile k+1);
 *   compute warps (the rest):  the edge physics of tile k+1.
 *
 * Two slots (node stage, edge inputs / results, reduction plan); per slot an
 * mbarrier `full` (bulk-copy bytes + one asynchronous arrival per memory
 * thread for its cp.async gathers) and an mbarrier `done` (one arrival per
 * compute thread after its results are in shared memory).  The FP64 pipe
 * works while the two memory phases of the neighbouring tiles run.  Same plan
 * data, same arithmetic, same order of additions as ls_tile_kernel: results
 * are bit-identical (tests/test_gpu_parity.py::test_pipe_kernel_*). */
constexpr int kPipeThreads = 512;
constexpr int kPipeHdrRing = 4;

__device__ __forceinline__ void
mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
/* arrival that fires when all prior cp.async of the calling thread have
 * landed; does not change the barrier's pending count (counted at init) */
__device__ __forceinline__ void
mbar_arrive_cp_async(uint64_t* bar)
{
  asm volatile(
    "cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
    : "memory");
}
template <class P>
struct PipeSmem
{
  static constexpr int NIN_MAX = LsSmem<P>::NIN_MAX;
  static constexpr int NEDGE = LsSmem<P>::NEDGE;
  int nodeLen, resStride, lrLen, valsLen, ellLen, entLen;
  size_t nodeBytes, rslotBytes, rowBytes;
  __host__ __device__ PipeSmem(const MeshPlanDev& mp, const LsPlanDev& lp)
  {
    nodeLen = (P::NC * mp.maxStaged + 1) & ~1;
    resStride = (mp.maxTileEdges + 1) & ~1;
    lrLen = (mp.maxTileEdges + 3) & ~3;
    valsLen = (lp.maxTileNnz + 3) & ~3;
    ellLen = (lp.maxTileEll + 3) & ~3;
    entLen = (lp.maxTileEnts + 3) & ~3;
    nodeBytes = sizeof(double) * (size_t)nodeLen;
    rslotBytes = sizeof(double) * (size_t)NEDGE * resStride + 4u * (size_t)lrLen +
                 4u * (size_t)ellLen + 12u * (size_t)entLen;
    rslotBytes = (rslotBytes + 15) & ~size_t(15);
    rowBytes = 8u * (size_t)valsLen + 4u * (size_t)valsLen;
  }
  /* two node slots, three result / plan slots, one row staging */
  __host__ __device__ size_t bytes() const
  {
    return 2 * nodeBytes + 3 * rslotBytes + rowBytes;
  }
};

/* Roles (warp index): 0 .. NRED-1 reducers, then NSTG stagers, the rest compute.
 *   stagers : per tile k: wait until the physics of tile k is done (its node
 *             slot is free) and the reduction of tile k-1 is done (its result
 *             slot is free), then stage tile k+2 -- all bulk copies, the halo
 *             gather (cp.async), slice offsets, header ring.  Issue only.
 *   reducers: per tile k: wait for its results, row walk + copy-out.
 *   compute : 32-edge units are dealt round-robin to the compute warps ACROSS
 *             tiles (a warp that is done with its units of tile k goes on to
 *             tile k+1 at once), so a tile whose edge count is not a multiple
 *             of the compute threads costs no idle round.
 * mbarriers per slot (three): full (bulk-copy bytes + stager lanes' cp.async
 * arrivals + the stager's own arrive), done (one arrival per compute warp),
 * red (one arrival per reducer warp). */
template <class P, int ND, int NRED, int NSTG>
__global__ void __launch_bounds__(kPipeThreads, 1) ls_pipe_kernel(
  const MeshPlanDev mp,
  const LsPlanDev lp,
  const NodeComps nc,
  const EdgeComps ec,
  const typename P::Opts o)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t barFull[3], barDone[3], barRed[3];
  __shared__ __align__(16) TileHdr s_hdr[kPipeHdrRing];
  __shared__ __align__(16) LsTileHdr s_lhdr[kPipeHdrRing];
  __shared__ int32_t s_slice[3][kMaxTileEnts / 32 + 2];

  constexpr int kRedThreads = NRED * 32;
  constexpr int kCmpWarps = kPipeThreads / 32 - NRED - NSTG;
  constexpr int kStgThreads = NSTG * 32;
  using S = PipeSmem<P>;
  const S L(mp, lp);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int K = (mp.nTiles - (int)blockIdx.x + G - 1) / G; /* my tiles */
  auto tile_of = [&](int k) { return (int)blockIdx.x + k * G; };

  /* tile k: node slot k & 1 (free again when its physics is done), result /
   * plan slot k % 3 (edge inputs -> edge results -> read by the reduction) */
  auto s_node_of = [&](int k) {
    return reinterpret_cast<double*>(smem_raw + (size_t)(k & 1) * L.nodeBytes);
  };
  auto s_res_of = [&](int k) {
    return reinterpret_cast<double*>(
      smem_raw + 2 * L.nodeBytes + (size_t)(k % 3) * L.rslotBytes);
  };
  auto s_lr_of = [&](int k) {
    return reinterpret_cast<uint32_t*>(s_res_of(k) + S::NEDGE * L.resStride);
  };
  auto s_ell_of = [&](int k) { return s_lr_of(k) + L.lrLen; };
  auto s_ent_of = [&](int k) {
    return reinterpret_cast<EntInfo*>(s_ell_of(k) + L.ellLen);
  };
  double* s_vals =
    reinterpret_cast<double*>(smem_raw + 2 * L.nodeBytes + 3 * L.rslotBytes);
  int32_t* s_delta = reinterpret_cast<int32_t*>(s_vals + L.valsLen);

  constexpr int kMdot = ND;
  constexpr int kPec = ND + 1;
  const bool hasPec = P::kNeedsPec && ec.pecfac != nullptr;
  const int nin = ND + (P::kNeedsMdot ? 1 : 0) + (hasPec ? 1 : 0);
  const EdgeCompSel<ND> ecomp{ec};

  if (tid == 0) {
    for (int b = 0; b < 3; ++b) {
      mbar_init(&barFull[b], 1 + kStgThreads);
      mbar_init(&barDone[b], kCmpWarps);
      mbar_init(&barRed[b], NRED);
    }
  }
  __syncthreads();

  if (warp >= NRED && warp < NRED + NSTG) {
    /* ============================== stagers ============================= */
    const int sw = warp - NRED;        /* stager warp index */
    const int st = sw * 32 + lane;     /* stager thread index */
    auto stagers_sync = [&]() {
      if (NSTG > 1)
        asm volatile("bar.sync 2, %0;" ::"n"(kStgThreads) : "memory");
      else
        __syncwarp();
    };
    /* headers of tile k -> ring slot k % kPipeHdrRing (lane i: word i of the
     * 64-byte TileHdr, lanes 16..31: the LsTileHdr) */
    auto hdr_load = [&](int k) -> int32_t {
      if (k >= K || sw != 0)
        return 0;
      return lane < 16
               ? __ldg(reinterpret_cast<const int32_t*>(mp.tiles + tile_of(k)) + lane)
               : __ldg(reinterpret_cast<const int32_t*>(lp.tiles + tile_of(k)) + (lane - 16));
    };
    auto hdr_store = [&](int k, int32_t w) {
      if (k >= K || sw != 0)
        return;
      if (lane < 16)
        reinterpret_cast<int32_t*>(&s_hdr[k % kPipeHdrRing])[lane] = w;
      else
        reinterpret_cast<int32_t*>(&s_lhdr[k % kPipeHdrRing])[lane - 16] = w;
    };
    /* stage tile k (its headers are in the ring, stored by this warp) */
    auto stage = [&](int k) {
      const TileHdr h = s_hdr[k % kPipeHdrRing];
      const LsTileHdr lh = s_lhdr[k % kPipeHdrRing];
      const int stride = even_up_i(h.nOwnPad + h.nHalo);
      double* s_node = s_node_of(k);
      double* s_res = s_res_of(k);
      const uint32_t bEll = (uint32_t)lh.ellLen * 4u;
      const uint32_t bEnt = round16((uint32_t)lh.nEnts * 4u);
      EntInfo* s_ent = s_ent_of(k);
      int32_t* s_row = reinterpret_cast<int32_t*>(s_ent + L.entLen);
      int32_t* s_go = s_row + L.entLen;
      uint64_t* bar = &barFull[k % 3];
      /* halo gather first (the longest chain: index -> data): lanes take
       * halo nodes q = lane, lane + 32, ...; indices from the fixed-stride
       * block, beyond it from the list */
      {
        const int32_t* blk = mp.haloBlock + (size_t)tile_of(k) * kHaloBlock;
        const int32_t* halo = mp.haloNodes + h.haloPtr;
        for (int q0 = 0; q0 < h.nHalo; q0 += 4 * kStgThreads) {
          int32_t g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * kStgThreads + st;
            g[u] = q < h.nHalo ? (q < kHaloBlock ? __ldg(blk + q) : __ldg(halo + q))
                               : -1;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * kStgThreads + st;
            if (q < h.nHalo) {
#pragma unroll
              for (int c = 0; c < P::NC; ++c)
                cp_async8(s_node + c * stride + h.nOwnPad + q, nc.c[c] + g[u]);
            }
          }
        }
      }
      /* slice offsets of the row-keyed list */
      if (sw == 0) {
        if (lane <= ((lh.nEnts + 31) >> 5))
          s_slice[k % 3][lane] = __ldg(lp.sliceOff + lh.slicePtr + lane);
        if (lane + 32 <= ((lh.nEnts + 31) >> 5))
          s_slice[k % 3][lane + 32] = __ldg(lp.sliceOff + lh.slicePtr + lane + 32);
      }
      /* bulk copies: copy q from lane 0 of stager warp q mod NSTG */
      if (lane == 0) {
        for (int q = sw; q < P::NC + 1 + nin + 4; q += NSTG) {
          if (q < P::NC)
            node_copy<P::NC>(q, s_node, stride, nc, h, bar);
          else if (q <= P::NC + nin)
            edge_copy(q - P::NC, s_lr_of(k), s_res, L.resStride, mp, h, ecomp, bar);
          else {
            const int r = q - (P::NC + 1 + nin);
            if (r == 0 && bEll)
              tma_load_1d(s_ell_of(k), lp.heEll + lh.ellPtr, bEll, bar);
            else if (r == 1 && bEnt)
              tma_load_1d(s_ent, lp.entInfo + lh.entPtr, bEnt, bar);
            else if (r == 2 && bEnt)
              tma_load_1d(s_row, lp.entRhsRow + lh.entPtr, bEnt, bar);
            else if (r == 3 && bEnt)
              tma_load_1d(s_go, lp.entGo + lh.entPtr, bEnt, bar);
          }
        }
      }
      /* every lane: asynchronous arrival for its gathers; then, after the
       * warp's plain stores (header ring, slice offsets), lane 0 posts the
       * byte count with a release-arrive */
      mbar_arrive_cp_async(bar);
      __syncwarp();
      if (st == 0)
        mbar_expect_tx(
          bar, node_copy_bytes(P::NC, h) + edge_stream_bytes(h, nin) + bEll +
                 3u * bEnt);
    };

    {
      const int32_t w0 = hdr_load(0), w1 = hdr_load(1), w2 = hdr_load(2);
      hdr_store(0, w0);
      hdr_store(1, w1);
      hdr_store(2, w2);
      stagers_sync();
    }
    for (int k = 0; k < 2 && k < K; ++k)
      stage(k);
    for (int k = 0; k + 2 < K; ++k) {
      const int32_t wNext = hdr_load(k + 3);
      /* node slot of tile k+2 = that of tile k: free when its physics is done */
      mbar_wait(&barDone[k % 3], (uint32_t)(k / 3) & 1u);
      /* result slot of tile k+2 = that of tile k-1: free when it is reduced */
      if (k >= 1)
        mbar_wait(&barRed[(k - 1) % 3], (uint32_t)((k - 1) / 3) & 1u);
      /* the copies overwrite shared memory read through the generic proxy */
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      stage(k + 2);
      hdr_store(k + 3, wNext);
      stagers_sync();
    }
  } else if (warp < NRED) {
    /* ============================= reducers ============================= */
    for (int k = 0; k < K; ++k) {
      const uint32_t par = (uint32_t)(k / 3) & 1u;
      mbar_wait(&barFull[k % 3], par); /* plan, headers, slice offsets */
      mbar_wait(&barDone[k % 3], par); /* edge results */
      const LsTileHdr lh = s_lhdr[k % kPipeHdrRing];
      const double* s_res = s_res_of(k);
      const uint32_t* s_ell = s_ell_of(k);
      const EntInfo* s_ent = s_ent_of(k);
      const int32_t* s_row = reinterpret_cast<const int32_t*>(s_ent + L.entLen);
      const int32_t* s_go = s_row + L.entLen;
      const int32_t* sliceOff = s_slice[k % 3];
      for (int row0 = tid - lane; row0 < lh.nEnts; row0 += kRedThreads) {
        const int row = row0 + lane;
        if (row < lh.nEnts) {
          const int s = row0 >> 5;
          const int o0 = sliceOff[s], o1 = sliceOff[s + 1];
          const uint32_t* hp = s_ell + o0 + lane;
          const int W = (o1 - o0) >> 5;
          const EntInfo ei = s_ent[row];
          double* vrow = s_vals + ei.base;
          double diag = 0.0;
          double rhs[P::NR];
#pragma unroll
          for (int d = 0; d < P::NR; ++d)
            rhs[d] = 0.0;
          constexpr int kBlk = 4;
          for (int w0 = 0; w0 < W; w0 += kBlk) {
            uint32_t hv[kBlk];
#pragma unroll
            for (int u = 0; u < kBlk; ++u)
              hv[u] = (w0 + u < W) ? hp[(w0 + u) * 32] : 0u;
            double dg[kBlk], off[kBlk], rr[kBlk][P::NR];
#pragma unroll
            for (int u = 0; u < kBlk; ++u)
              if (hv[u] & kHeValid)
                P::contrib(
                  he_side(hv[u]), s_res, L.resStride, (int)he_edge(hv[u]), dg[u],
                  off[u], rr[u]);
#pragma unroll
            for (int u = 0; u < kBlk; ++u)
              if (hv[u] & kHeValid) {
                diag += dg[u];
#pragma unroll
                for (int d = 0; d < P::NR; ++d)
                  rhs[d] += rr[u][d];
                double* dst = vrow + he_k(hv[u]);
                if (hv[u] & kHeDup)
                  off[u] += *dst;
                *dst = off[u];
              }
          }
          vrow[ei.diagK] = diag;
          const int32_t delta = s_go[row] - (int32_t)ei.base;
          for (int q = 0; q < (int)ei.nnz; ++q)
            s_delta[ei.base + q] = delta;
          const int64_t grow = s_row[row];
#pragma unroll
          for (int d = 0; d < P::NR; ++d)
            lp.rhs[(int64_t)d * lp.rhsStride + grow] = rhs[d];
        }
        __syncwarp();
        {
          const int last = min(row0 + 31, lh.nEnts - 1);
          const EntInfo e0 = s_ent[row0], e1 = s_ent[last];
          const int end = (int)e1.base + (int)e1.nnz;
#pragma unroll 4
          for (int e = (int)e0.base + lane; e < end; e += 32)
            lp.values[e + s_delta[e]] = s_vals[e];
        }
        __syncwarp();
      }
      /* this warp has left the result slot; the row staging is shared by the
       * reducer warps: all of them are through before the next tile's rows
       * are written */
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&barRed[k % 3]);
      asm volatile("bar.sync 1, %0;" ::"n"(kRedThreads) : "memory");
    }
  } else {
    /* ============================== compute ============================= */
    const int cw = warp - NRED - NSTG; /* 0 .. kCmpWarps-1 */
    int g0 = 0;                     /* units dealt before this tile, mod kCmpWarps */
    for (int k = 0; k < K; ++k) {
      const uint32_t par = (uint32_t)(k / 3) & 1u;
      mbar_wait(&barFull[k % 3], par);
      const TileHdr h = s_hdr[k % kPipeHdrRing];
      const int stride = even_up_i(h.nOwnPad + h.nHalo);
      double* s_res = s_res_of(k);
      const uint32_t* s_lr = s_lr_of(k);
      const SmemLd ld{s_node_of(k), stride};
      const int nUnits = (h.nEdges + 31) >> 5;
      /* my first unit of this tile: (g0 + u) % kCmpWarps == cw */
      int u = cw - g0;
      if (u < 0)
        u += kCmpWarps;
      for (; u < nUnits; u += kCmpWarps) {
        const int j = u * 32 + lane;
        if (j < h.nEdges) {
          const uint32_t v = s_lr[j];
          const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
          double av[ND];
#pragma unroll
          for (int d = 0; d < ND; ++d)
            av[d] = s_res[d * L.resStride + j];
          double mdot = 0.0, pecfac = 0.0;
          if (P::kNeedsMdot)
            mdot = s_res[kMdot * L.resStride + j];
          if (hasPec)
            pecfac = s_res[kPec * L.resStride + j];
          double res[P::NRES];
          P::compute(ld, l, r, av, mdot, pecfac, o, res);
#pragma unroll
          for (int q = 0; q < P::NRES; ++q)
            s_res[q * L.resStride + j] = res[q];
        }
      }
      g0 = (g0 + nUnits) % kCmpWarps;
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&barDone[k % 3]);
    }
  }
}

/* ------------------------------------------------------------------ */
/*  two scalar systems of one graph in one launch (SST: TKE + SDR)      */
/* ------------------------------------------------------------------ */

/* ShearStressTransportEquationSystem assembles the TKE and the SDR system
 * from the same state (src/ShearStressTransportEquationSystem.C:247-320: both
 * assemble_and_solve calls precede update_and_clip), with the same
 * ScalarEdgeSolverAlg arithmetic (src/edge_kernels/ScalarEdgeSolverAlg.C:
 * 55-206) on the same mesh and the same graph.  This kernel stages what the
 * two assemblies share -- coordinates, velocity, density, the (L,R) records,
 * area vectors, mass flow rates and the whole reduction plan -- once per
 * tile: 17 node components instead of 2 x 12, one set of edge / plan streams
 * instead of two, one CTA start-up instead of two.  One CTA of 512 threads
 * per tile; a work item of phase 1 is (edge, system), of phase 2 (row,
 * system); every value and rhs entry of both systems is written once, in
 * the same order of additions as ls_tile_kernel<ScalarP> (bit-identical
 * results, tests/test_gpu_parity.py::test_scalar_pair_equals_two_calls).
 * Staged node components: x[ND], v[ND], rho, then per system q, dqdx[ND],
 * diffFluxCoeff. */
constexpr int kPairThreads = 512;

template <int ND>
struct PairSmem
{
  static constexpr int NC = 2 * ND + 1 + 2 * (ND + 2);
  static constexpr int NIN = ND + 1; /* area, mdot */
  static constexpr int NRES = 5;     /* per system: a00 a01 a10 a11 flux */
  int nodeRegion, resStride, valsLen, ellLen, entLen, lrLen;
  __host__ __device__ PairSmem(const MeshPlanDev& mp, const LsPlanDev& lp)
  {
    resStride = (mp.maxTileEdges + 1) & ~1;
    lrLen = (mp.maxTileEdges + 3) & ~3;
    valsLen = (lp.maxTileNnz + 3) & ~3;
    /* row staging: values of both systems (8 B each) + deltas (4 B) */
    const int rowRegion = 2 * valsLen + valsLen / 2;
    const int stage = NC * mp.maxStaged;
    nodeRegion = stage > rowRegion ? stage : rowRegion;
    ellLen = lp.maxTileEll;
    entLen = (lp.maxTileEnts + 3) & ~3;
  }
  __host__ __device__ size_t bytes() const
  {
    return sizeof(double) *
             ((size_t)nodeRegion + (size_t)(NIN + 2 * NRES) * resStride) +
           4u * (size_t)ellLen + 12u * (size_t)entLen + 4u * (size_t)lrLen;
  }
};

template <int ND>
__global__ void __launch_bounds__(kPairThreads, 1) scalar_pair_tile_kernel(
  const MeshPlanDev mp,
  const LsPlanDev lp, /* plan; values / rhs of system A */
  double* __restrict__ valuesB,
  double* __restrict__ rhsB,
  const NodeComps nc,
  const EdgeComps ec,
  const nw_scalar_opts oA,
  const nw_scalar_opts oB)
{
  using S = PairSmem<ND>;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int32_t s_slice[kMaxTileEnts / 32 + 2];

  const TileHdr h = mp.tiles[blockIdx.x];
  const LsTileHdr lh = lp.tiles[blockIdx.x];
  const int32_t g0 = halo_index_early(mp);
  const S L(mp, lp);
  const int stride = even_up_i(h.nOwnPad + h.nHalo);
  const int rs = L.resStride;

  double* s_node = smem;
  double* s_in = s_node + L.nodeRegion;  /* area[ND], mdot */
  double* s_res = s_in + S::NIN * rs;    /* [system][5][edge] */
  double* s_vals = s_node;               /* row staging of both systems */
  int32_t* s_delta = reinterpret_cast<int32_t*>(s_vals + 2 * L.valsLen);
  uint32_t* s_ell = reinterpret_cast<uint32_t*>(s_res + 2 * S::NRES * rs);
  EntInfo* s_ent = reinterpret_cast<EntInfo*>(s_ell + L.ellLen);
  int32_t* s_row = reinterpret_cast<int32_t*>(s_ent + L.entLen);
  int32_t* s_go = s_row + L.entLen;
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(s_go + L.entLen);

  const EdgeCompSel<ND> ecomp{ec};
  const uint32_t bEll = (uint32_t)lh.ellLen * 4u;
  const uint32_t bEnt = round16((uint32_t)lh.nEnts * 4u);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_expect_tx(
      &bar[0], node_copy_bytes(S::NC, h) + edge_stream_bytes(h, S::NIN));
    mbar_expect_tx(&bar[1], bEll + 3u * bEnt);
  }
  __syncthreads();
  issue_spread(S::NC + 1 + S::NIN + 4, [&](int q) {
    if (q < S::NC)
      node_copy<S::NC>(q, s_node, stride, nc, h, &bar[0]);
    else if (q <= S::NC + S::NIN)
      edge_copy(q - S::NC, s_lr, s_in, rs, mp, h, ecomp, &bar[0]);
    else {
      const int r = q - (S::NC + 1 + S::NIN);
      if (r == 0 && bEll)
        tma_load_1d(s_ell, lp.heEll + lh.ellPtr, bEll, &bar[1]);
      else if (r == 1 && bEnt)
        tma_load_1d(s_ent, lp.entInfo + lh.entPtr, bEnt, &bar[1]);
      else if (r == 2 && bEnt)
        tma_load_1d(s_row, lp.entRhsRow + lh.entPtr, bEnt, &bar[1]);
      else if (r == 3 && bEnt)
        tma_load_1d(s_go, lp.entGo + lh.entPtr, bEnt, &bar[1]);
    }
  });
  {
    const int nSl = (lh.nEnts + 31) >> 5;
    if ((int)threadIdx.x <= nSl)
      s_slice[threadIdx.x] = __ldg(lp.sliceOff + lh.slicePtr + threadIdx.x);
  }
  stage_halo_gather<S::NC>(s_node, stride, nc, h, mp.haloNodes, g0);
  stage_halo_wait();
  mbar_wait(&bar[0], 0);
  __syncthreads();

  /* ---- phase 1: work item = (edge j, system sys) ---- */
  for (int it = threadIdx.x; it < 2 * h.nEdges; it += blockDim.x) {
    const int sys = it >= h.nEdges ? 1 : 0;
    const int j = it - sys * h.nEdges;
    const uint32_t v = s_lr[j];
    const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
    double av[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      av[d] = s_in[d * rs + j];
    const double mdot = s_in[ND * rs + j];
    const int cq = 2 * ND + 1 + sys * (ND + 2); /* first component of the system */
    ScalNode<ND> Ln, Rn;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      Ln.x[d] = s_node[d * stride + l];
      Rn.x[d] = s_node[d * stride + r];
      Ln.v[d] = s_node[(ND + d) * stride + l];
      Rn.v[d] = s_node[(ND + d) * stride + r];
      Ln.dq[d] = s_node[(cq + 1 + d) * stride + l];
      Rn.dq[d] = s_node[(cq + 1 + d) * stride + r];
    }
    Ln.rho = s_node[2 * ND * stride + l];
    Rn.rho = s_node[2 * ND * stride + r];
    Ln.q = s_node[cq * stride + l];
    Rn.q = s_node[cq * stride + r];
    Ln.mu = s_node[(cq + 1 + ND) * stride + l];
    Rn.mu = s_node[(cq + 1 + ND) * stride + r];
    double res[S::NRES];
    scalar_edge<ND>(Ln, Rn, av, mdot, sys ? oB : oA, res, res[4]);
    double* out = s_res + sys * S::NRES * rs + j;
#pragma unroll
    for (int k = 0; k < S::NRES; ++k)
      out[k * rs] = res[k];
  }
  mbar_wait(&bar[1], 0);
  __syncthreads();

  /* ---- phases 2+3: work item = (32-row slice, system), warp by warp ---- */
  {
    const int lane = threadIdx.x & 31;
    const int nWarps = blockDim.x >> 5;
    const int nSl = (lh.nEnts + 31) >> 5;
    for (int p = threadIdx.x >> 5; p < 2 * nSl; p += nWarps) {
      const int sys = p >= nSl ? 1 : 0;
      const int sl = p - sys * nSl;
      const int row0 = sl << 5, row = row0 + lane;
      const double* sres = s_res + sys * S::NRES * rs;
      double* svals = s_vals + sys * L.valsLen;
      double* gvals = sys ? valuesB : lp.values;
      double* grhs = sys ? rhsB : lp.rhs;
      if (row < lh.nEnts) {
        const int o0 = s_slice[sl], o1 = s_slice[sl + 1];
        const uint32_t* hp = s_ell + o0 + lane;
        const int W = (o1 - o0) >> 5;
        const EntInfo ei = s_ent[row];
        double* vrow = svals + ei.base;
        double diag = 0.0, rhs = 0.0;
        constexpr int kBlk = 4;
        for (int w0 = 0; w0 < W; w0 += kBlk) {
          uint32_t hv[kBlk];
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            hv[u] = (w0 + u < W) ? hp[(w0 + u) * 32] : 0u;
          double dg[kBlk], off[kBlk], rr[kBlk][1];
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            if (hv[u] & kHeValid)
              ScalarP<ND>::contrib(
                he_side(hv[u]), sres, rs, (int)he_edge(hv[u]), dg[u], off[u],
                rr[u]);
#pragma unroll
          for (int u = 0; u < kBlk; ++u)
            if (hv[u] & kHeValid) {
              diag += dg[u];
              rhs += rr[u][0];
              double* dst = vrow + he_k(hv[u]);
              if (hv[u] & kHeDup)
                off[u] += *dst;
              *dst = off[u];
            }
        }
        vrow[ei.diagK] = diag;
        /* both systems write the same deltas (same graph) */
        const int32_t delta = s_go[row] - (int32_t)ei.base;
        for (int k = 0; k < (int)ei.nnz; ++k)
          s_delta[ei.base + k] = delta;
        grhs[s_row[row]] = rhs;
      }
      __syncwarp();
      {
        const int last = min(row0 + 31, lh.nEnts - 1);
        const EntInfo e0 = s_ent[row0], e1 = s_ent[last];
        const int end = (int)e1.base + (int)e1.nnz;
#pragma unroll 4
        for (int e = (int)e0.base + lane; e < end; e += 32)
          gvals[e + s_delta[e]] = svals[e];
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/*  linear-system atomic kernel (comparison variant)                   */
/* ------------------------------------------------------------------ */

__device__ __forceinline__ int
global_slot(const TileHdr& h, const int32_t* haloNodes, int local)
{
  return local < h.nOwnPad ? h.node0 + local
                           : __ldg(haloNodes + h.haloPtr + (local - h.nOwnPad));
}

template <class P, int ND>
__global__ void __launch_bounds__(kTileThreads) ls_atomic_kernel(
  const MeshPlanDev mp,
  const LsPlanDev lp,
  const AtomicMapDev am,
  const NodeComps nc,
  const EdgeComps ec,
  const typename P::Opts o,
  double* diagOut)
{
  const TileHdr h = mp.tiles[blockIdx.x];
  const GmemLd ld{&nc};
  const int lane = threadIdx.x & 31;
  const int nIter = (h.nEdges + blockDim.x - 1) / blockDim.x;
  for (int it = 0; it < nIter; ++it) {
    const int j = it * blockDim.x + threadIdx.x;
    const int64_t es = (int64_t)h.edge0 + j;
    const bool valid = (j < h.nEdges) && mp.primary[es];
    double LL = 0, LR = 0, RL = 0, RR = 0;
    double flux[P::NR];
#pragma unroll
    for (int d = 0; d < P::NR; ++d)
      flux[d] = 0.0;
    int sLL = -1, sLR = -1, sRL = -1, sRR = -1, rowL = -1, rowR = -1;
    int gl = 0, gr = 0;
    if (valid) {
      const uint32_t v = __ldg(mp.lr + es);
      gl = global_slot(h, mp.haloNodes, (int)(v & 0xffffu));
      gr = global_slot(h, mp.haloNodes, (int)(v >> 16));
      double av[ND];
#pragma unroll
      for (int d = 0; d < ND; ++d)
        av[d] = __ldg(ec.area[d] + es);
      double mdot = 0.0, pecfac = 0.0;
      if (P::kNeedsMdot)
        mdot = __ldg(ec.mdot + es);
      if (P::kNeedsPec && ec.pecfac)
        pecfac = __ldg(ec.pecfac + es);
      double res[P::NRES];
      P::compute(ld, gl, gr, av, mdot, pecfac, o, res);
      P::block(res, LL, LR, RL, RR, flux);
      const int4 s4 = __ldg(reinterpret_cast<const int4*>(am.slots) + es);
      sLL = s4.x;
      sLR = s4.y;
      sRL = s4.z;
      sRR = s4.w;
      const int2 r2 = __ldg(reinterpret_cast<const int2*>(am.rhsRows) + es);
      rowL = r2.x;
      rowR = r2.y;
    }
    /* warp aggregation of the L-row sums: tile-edges are sorted by L node, so
     * lanes hitting the same diagonal slot are adjacent */
    double acc[1 + P::NR];
    acc[0] = LL;
#pragma unroll
    for (int d = 0; d < P::NR; ++d)
      acc[1 + d] = -flux[d];
    const uint32_t key = (uint32_t)sLL;
    seg_scan<1 + P::NR>(acc, key, lane);
    const bool tail = seg_tail(key, lane);
    if (valid) {
      if (tail && sLL >= 0) {
        atomicAdd(lp.values + sLL, acc[0]);
#pragma unroll
        for (int d = 0; d < P::NR; ++d)
          atomicAdd(lp.rhs + (int64_t)d * lp.rhsStride + rowL, acc[1 + d]);
      }
      if (sLR >= 0)
        atomicAdd(lp.values + sLR, LR);
      if (sRL >= 0)
        atomicAdd(lp.values + sRL, RL);
      if (sRR >= 0) {
        atomicAdd(lp.values + sRR, RR);
#pragma unroll
        for (int d = 0; d < P::NR; ++d)
          atomicAdd(lp.rhs + (int64_t)d * lp.rhsStride + rowR, flux[d]);
      }
      if (diagOut) {
        /* NGPApplyCoeff::extract_diagonal (src/SolverAlgorithm.C:87-105) */
        atomicAdd(diagOut + gl, LL);
        atomicAdd(diagOut + gr, RR);
      }
    }
  }
}

/* monolithic momentum: full 2ND x 2ND block through the slot map */
template <int ND, bool VOF = false>
__global__ void __launch_bounds__(kTileThreads) momentum_mono_atomic_kernel(
  const MeshPlanDev mp,
  const int32_t* __restrict__ slots,
  const int32_t* __restrict__ rhsRows,
  double* values,
  double* rhs,
  const NodeComps nc,
  const EdgeComps ec,
  const nw_momentum_opts o,
  double* diagOut)
{
  using P = MomentumUvwP<ND>;
  constexpr int NB = 2 * ND;
  const TileHdr h = mp.tiles[blockIdx.x];
  const GmemLd ld{&nc};
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const int64_t es = (int64_t)h.edge0 + j;
    if (!mp.primary[es])
      continue;
    const uint32_t v = __ldg(mp.lr + es);
    const int gl = global_slot(h, mp.haloNodes, (int)(v & 0xffffu));
    const int gr = global_slot(h, mp.haloNodes, (int)(v >> 16));
    double av[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      av[d] = __ldg(ec.area[d] + es);
    const double mdot = __ldg(ec.mdot + es);
    double pecfac = ec.pecfac ? __ldg(ec.pecfac + es) : 0.0;
    MomNode<ND> L, R;
    P::load(ld, gl, L);
    P::load(ld, gr, R);
    if (o.fuse_peclet) {
      PecNode<ND> pl, pr;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        pl.x[d] = L.x[d];
        pr.x[d] = R.x[d];
        pl.v[d] = L.u[d];
        pr.v[d] = R.u[d];
      }
      pl.rho = L.rho;
      pr.rho = R.rho;
      pl.mu = L.mu;
      pr.mu = R.mu;
      pecfac = peclet_eval(o.pf, peclet_number<ND>(pl, pr, o.pec_eps));
    }
    MomResult<ND> m;
    if (VOF)
      momentum_edge_vof<ND>(L, R, av, mdot, pecfac, o, m);
    else
      momentum_edge<ND>(L, R, av, mdot, pecfac, o, m);
    const int32_t* sl = slots + es * (NB * NB);
    const int32_t* rr = rhsRows + es * NB;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
#pragma unroll
      for (int jj = 0; jj < ND; ++jj) {
        double LL, LR, RL, RR;
        momentum_block_entry<ND>(m, av, o.relax_fac, i, jj, LL, LR, RL, RR);
        const int a = sl[i * NB + jj], b = sl[i * NB + ND + jj];
        const int c = sl[(ND + i) * NB + jj], d = sl[(ND + i) * NB + ND + jj];
        if (a >= 0)
          atomicAdd(values + a, LL);
        if (b >= 0)
          atomicAdd(values + b, LR);
        if (c >= 0)
          atomicAdd(values + c, RL);
        if (d >= 0)
          atomicAdd(values + d, RR);
        if (diagOut && i == 0 && jj == 0) {
          atomicAdd(diagOut + gl, LL);
          atomicAdd(diagOut + gr, RR);
        }
      }
      if (rr[i] >= 0)
        atomicAdd(rhs + rr[i], -m.flux[i]);
      if (rr[ND + i] >= 0)
        atomicAdd(rhs + rr[ND + i], m.flux[i]);
    }
  }
}

/* ------------------------------------------------------------------ */
/*  mdot / peclet tile kernels (no reduction)                          */
/* ------------------------------------------------------------------ */

template <int ND>
__global__ void __launch_bounds__(kTileThreads) mdot_tile_kernel(
  const MeshPlanDev mp,
  const NodeComps nc,
  const EdgeComps ec,
  double* __restrict__ mdotOut,
  const nw_mdot_opts o)
{
  using P = ContinuityP<ND>;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar;
  NW_PT_BEGIN(3);
  const TileHdr h = mp.tiles[blockIdx.x];
  const int32_t g0 = halo_index_early(mp);
  const int stride = even_up_i(h.nOwnPad + h.nHalo);
  const int estride = even_up_i(mp.maxTileEdges);
  double* s_area = smem + (size_t)P::NC * mp.maxStaged;
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(s_area + ND * estride);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_expect_tx(&bar, node_copy_bytes(P::NC, h) + edge_stream_bytes(h, ND));
  }
  const EdgeCompSel<ND> ecomp{ec};
  __syncthreads();
  NW_PT_MARK();
  issue_spread(P::NC + 1 + ND, [&](int q) {
    if (q < P::NC)
      node_copy<P::NC>(q, smem, stride, nc, h, &bar);
    else
      edge_copy(q - P::NC, s_lr, s_area, estride, mp, h, ecomp, &bar);
  });
  NW_PT_MARK();
  stage_halo_gather<P::NC>(smem, stride, nc, h, mp.haloNodes, g0);
  NW_PT_MARK();
  stage_halo_wait();
  mbar_wait(&bar, 0);
  __syncthreads();
  NW_PT_MARK();
  const SmemLd ld{smem, stride};
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const uint32_t v = s_lr[j];
    const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
    double av[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      av[d] = s_area[d * estride + j];
    ContNode<ND> L, R;
    P::load(ld, l, L);
    P::load(ld, r, R);
    const MdotCore<ND> c =
      mdot_core<ND>(L, R, av, o.noc_fac, o.interp_together);
    mdotOut[h.edge0 + j] = c.tmdot;
  }
  NW_PT_MARK();
  NW_PT_END();
}

template <int ND>
__global__ void __launch_bounds__(kTileThreads) peclet_tile_kernel(
  const MeshPlanDev mp,
  const NodeComps nc, /* x, v, rho, mu */
  double* __restrict__ pecfacOut,
  const nw_peclet_opts o)
{
  constexpr int NC = 2 * ND + 2;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar;
  NW_PT_BEGIN(4);
  const TileHdr h = mp.tiles[blockIdx.x];
  const int32_t g0 = halo_index_early(mp);
  const int stride = even_up_i(h.nOwnPad + h.nHalo);
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(smem + (size_t)NC * mp.maxStaged);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_expect_tx(&bar, node_copy_bytes(NC, h) + edge_stream_bytes(h, 0));
  }
  const EdgeComps noEdge{};
  const EdgeCompSel<ND> ecomp{noEdge};
  __syncthreads();
  NW_PT_MARK();
  issue_spread(NC + 1, [&](int q) {
    if (q < NC)
      node_copy<NC>(q, smem, stride, nc, h, &bar);
    else
      edge_copy(0, s_lr, (double*)nullptr, 0, mp, h, ecomp, &bar);
  });
  NW_PT_MARK();
  stage_halo_gather<NC>(smem, stride, nc, h, mp.haloNodes, g0);
  NW_PT_MARK();
  stage_halo_wait();
  mbar_wait(&bar, 0);
  __syncthreads();
  NW_PT_MARK();
  const SmemLd ld{smem, stride};
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const uint32_t v = s_lr[j];
    const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
    PecNode<ND> L, R;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      L.x[d] = ld(d, l);
      R.x[d] = ld(d, r);
      L.v[d] = ld(ND + d, l);
      R.v[d] = ld(ND + d, r);
    }
    L.rho = ld(2 * ND, l);
    R.rho = ld(2 * ND, r);
    L.mu = ld(2 * ND + 1, l);
    R.mu = ld(2 * ND + 1, r);
    pecfacOut[h.edge0 + j] = peclet_eval(o.pf, peclet_number<ND>(L, R, o.eps));
  }
  NW_PT_MARK();
  NW_PT_END();
}

/* ------------------------------------------------------------------ */
/*  mdot / continuity with the optional terms (balanced forcing, GCL)  */
/* ------------------------------------------------------------------ */

template <int ND>
__device__ __forceinline__ ContExtra<ND>
load_cont_extra(const ContExtraDev& ex, int gl, int gr, int64_t es)
{
  ContExtra<ND> x;
  x.balanced = ex.balanced != 0;
  x.gcl = ex.gcl != 0;
  x.smaskL = x.smaskR = 0.0;
  x.faceVelMag = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    x.gravity[d] = ex.gravity[d];
    x.srcL[d] = x.srcR[d] = 0.0;
  }
  if (x.balanced) {
    x.smaskL = __ldg(ex.smask + gl);
    x.smaskR = __ldg(ex.smask + gr);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      x.srcL[d] = __ldg(ex.src[d] + gl);
      x.srcR[d] = __ldg(ex.src[d] + gr);
    }
  }
  if (x.gcl)
    x.faceVelMag = __ldg(ex.faceVelMag + es);
  return x;
}

/* MdotEdgeAlg with the optional terms: every tile-edge slot (both copies of a
 * cut edge) is written, like mdot_tile_kernel */
template <int ND>
__global__ void __launch_bounds__(kTileThreads) mdot_ext_kernel(
  const MeshPlanDev mp, const NodeComps nc, const EdgeComps ec,
  const ContExtraDev ex, double* __restrict__ mdotOut, const nw_mdot_opts o)
{
  using P = ContinuityP<ND>;
  const TileHdr h = mp.tiles[blockIdx.x];
  const GmemLd ld{&nc};
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const int64_t es = (int64_t)h.edge0 + j;
    const uint32_t v = __ldg(mp.lr + es);
    const int gl = global_slot(h, mp.haloNodes, (int)(v & 0xffffu));
    const int gr = global_slot(h, mp.haloNodes, (int)(v >> 16));
    double av[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      av[d] = __ldg(ec.area[d] + es);
    ContNode<ND> L, R;
    P::load(ld, gl, L);
    P::load(ld, gr, R);
    const ContExtra<ND> x = load_cont_extra<ND>(ex, gl, gr, es);
    mdotOut[es] = mdot_core_ext<ND>(L, R, x, av, o.noc_fac, o.interp_together).tmdot;
  }
}

/* ContinuityEdgeSolverAlg with the optional terms, atomic scatter through the
 * slot map (primary copies only) */
template <int ND>
__global__ void __launch_bounds__(kTileThreads) continuity_ext_atomic_kernel(
  const MeshPlanDev mp, const LsPlanDev lp, const AtomicMapDev am,
  const NodeComps nc, const EdgeComps ec, const ContExtraDev ex,
  const nw_continuity_opts o)
{
  using P = ContinuityP<ND>;
  const TileHdr h = mp.tiles[blockIdx.x];
  const GmemLd ld{&nc};
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const int64_t es = (int64_t)h.edge0 + j;
    if (!mp.primary[es])
      continue;
    const uint32_t v = __ldg(mp.lr + es);
    const int gl = global_slot(h, mp.haloNodes, (int)(v & 0xffffu));
    const int gr = global_slot(h, mp.haloNodes, (int)(v >> 16));
    double av[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      av[d] = __ldg(ec.area[d] + es);
    ContNode<ND> L, R;
    P::load(ld, gl, L);
    P::load(ld, gr, R);
    const ContExtra<ND> x = load_cont_extra<ND>(ex, gl, gr, es);
    const MdotCore<ND> c =
      mdot_core_ext<ND>(L, R, x, av, o.noc_fac, o.interp_together);
    /* scaling as continuity_edge */
    const double solveInc = o.solve_incompressible;
    const double denScale = nw_rcp(c.rhoIp) * solveInc + (1.0 - solveInc);
    const double invTauScale = o.gamma1 * nw_rcp(o.dt);
    double tmdot = c.tmdot;
    tmdot *= invTauScale;
    tmdot *= denScale;
    const double lhsfac = -c.asq_inv_axdx * c.projTimeScale * denScale * invTauScale;
    const int4 s4 = __ldg(reinterpret_cast<const int4*>(am.slots) + es);
    const int2 r2 = __ldg(reinterpret_cast<const int2*>(am.rhsRows) + es);
    if (s4.x >= 0) {
      atomicAdd(lp.values + s4.x, -lhsfac);
      atomicAdd(lp.rhs + r2.x, -tmdot);
    }
    if (s4.y >= 0)
      atomicAdd(lp.values + s4.y, lhsfac);
    if (s4.z >= 0)
      atomicAdd(lp.values + s4.z, lhsfac);
    if (s4.w >= 0) {
      atomicAdd(lp.values + s4.w, -lhsfac);
      atomicAdd(lp.rhs + r2.y, tmdot);
    }
  }
}

/* ------------------------------------------------------------------ */
/*  nodal gradient                                                     */
/* ------------------------------------------------------------------ */

using GradOut = CompPtrs;

/* NodalGradEdgeAlg (src/ngp_algorithms/NodalGradEdgeAlg.C:85-109) with the
 * zero-fill of NodalGradAlgDriver::pre_work fused: one thread per owned node
 * walks the node's half-edges (sliced-ELL list, TMA-staged), sums in
 * registers in list order and writes the node's gradient once, coalesced. */
template <int D1, int ND>
__global__ void __launch_bounds__(kTileThreads, D1 == 1 ? 8 : 5) grad_tile_kernel(
  const MeshPlanDev mp,
  const NodeComps phi,
  const double* __restrict__ dualVol,
  const EdgeComps ec,
  const GradOut out,
  const NodePushDev push)
{
  constexpr int NV = D1 * ND;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int32_t s_slice[kMaxTileEnts / 32 + 2];
  NW_PT_BEGIN(D1 == 1 ? 5 : 6);
  const int tile = (int)blockIdx.x;
  const TileHdr h = mp.tiles[tile];
  const int32_t g0 = halo_index_early(mp, tile);
  const int stride = even_up_i(h.nOwnPad + h.nHalo);
  const int estride = even_up_i(mp.maxTileEdges);
  double* s_phi = smem;
  double* s_area = smem + (size_t)D1 * mp.maxStaged;
  uint32_t* s_ell = reinterpret_cast<uint32_t*>(s_area + ND * estride);
  uint32_t* s_lr = s_ell + mp.maxTileEllNode;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  const EdgeCompSel<ND> ecomp{ec};
  const uint32_t bEll = (uint32_t)h.ellLenNode * 4u;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], node_copy_bytes(D1, h) + edge_stream_bytes(h, ND));
    mbar_expect_tx(&bar[1], bEll);
  }
  __syncthreads();
  NW_PT_MARK();
  issue_spread(D1 + 1 + ND + 1, [&](int q) {
    if (q < D1)
      node_copy<D1>(q, s_phi, stride, phi, h, &bar[0]);
    else if (q <= D1 + ND)
      edge_copy(q - D1, s_lr, s_area, estride, mp, h, ecomp, &bar[0]);
    else if (bEll)
      tma_load_1d(s_ell, mp.heNodeEll + h.ellPtrNode, bEll, &bar[1]);
  });
  NW_PT_MARK();
  {
    const int nSl = (h.nOwn + 31) >> 5;
    if ((int)threadIdx.x <= nSl)
      s_slice[threadIdx.x] = __ldg(mp.sliceOffNode + h.slicePtrNode + threadIdx.x);
  }
  stage_halo_gather<D1>(s_phi, stride, phi, h, mp.haloNodes, g0);
  NW_PT_MARK();
  stage_halo_wait();
  mbar_wait(&bar[0], 0);
  mbar_wait(&bar[1], 0);
  __syncthreads();
  NW_PT_MARK();

  for (int i = threadIdx.x; i < h.nOwn; i += blockDim.x) {
    const int sl = i >> 5;
    const int o0 = s_slice[sl], o1 = s_slice[sl + 1];
    const uint32_t* hp = s_ell + o0 + (i & 31);
    const int W = (o1 - o0) >> 5;
    /* issued now, needed after the walk */
    const double vol = __ldg(dualVol + h.node0 + i);
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k)
      acc[k] = 0.0;
    for (int w = 0; w < W; ++w) {
      const uint32_t hv = hp[w * 32];
      if (hv & kHeValid) {
        const int j = (int)he_edge(hv);
        const uint32_t v = s_lr[j];
        const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
        /* L node: += a_j phiIp ; R node: -= a_j phiIp (NodalGradEdgeAlg.C:100-106) */
        const double sgn = he_side(hv) ? -1.0 : 1.0;
        double av[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
          av[d] = sgn * s_area[d * estride + j];
#pragma unroll
        for (int c = 0; c < D1; ++c) {
          const double phiIp =
            0.5 * (s_phi[c * stride + l] + s_phi[c * stride + r]);
#pragma unroll
          for (int d = 0; d < ND; ++d)
            acc[c * ND + d] += av[d] * phiIp;
        }
      }
    }
    /* the reference divides every edge term by the dual volume before summing;
     * scaling the node's sum once differs by rounding only (a few ulp) and
     * saves a divide per half-edge */
    const double invVol = nw_rcp(vol);
#pragma unroll
    for (int k = 0; k < NV; ++k)
      out.c[k][h.node0 + i] = acc[k] * invVol;
  }
  NW_PT_MARK();
  if (push.tilePtr) {
    /* eager exchange: the partial sums of this tile's shared nodes go straight
     * to the other sharers' windows (the values were written by this CTA: the
     * barrier makes them visible to all of its threads) */
    const int p0 = __ldg(push.tilePtr + tile), p1 = __ldg(push.tilePtr + tile + 1);
    if (p1 > p0) { /* uniform over the CTA */
      __syncthreads();
      for (int g = p0 + (int)threadIdx.x; g < p1; g += blockDim.x) {
        const int slot = __ldg(push.slot + g);
        double* w = push.pp.peerWindow[__ldg(push.peer + g)] + push.pp.winOff +
                    __ldg(push.dst + g) * NV;
#pragma unroll
        for (int c = 0; c < NV; ++c) /* constant indices: no local copy of `out` */
          w[c] = out.c[c][slot];
      }
    }
  }
  NW_PT_END();
}

/* atomic comparison variant: grad must be zeroed beforehand */
template <int D1, int ND>
__global__ void __launch_bounds__(kTileThreads) grad_atomic_kernel(
  const MeshPlanDev mp,
  const NodeComps phi,
  const double* __restrict__ dualVol,
  const EdgeComps ec,
  const GradOut out)
{
  const TileHdr h = mp.tiles[blockIdx.x];
  for (int j = threadIdx.x; j < h.nEdges; j += blockDim.x) {
    const int64_t es = (int64_t)h.edge0 + j;
    if (!mp.primary[es])
      continue;
    const uint32_t v = __ldg(mp.lr + es);
    const int gl = global_slot(h, mp.haloNodes, (int)(v & 0xffffu));
    const int gr = global_slot(h, mp.haloNodes, (int)(v >> 16));
    const double invVolL = 1.0 / __ldg(dualVol + gl);
    const double invVolR = 1.0 / __ldg(dualVol + gr);
#pragma unroll
    for (int i = 0; i < D1; ++i) {
      const double phiIp = 0.5 * (__ldg(phi.c[i] + gl) + __ldg(phi.c[i] + gr));
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double ajPhiIp = __ldg(ec.area[d] + es) * phiIp;
        atomicAdd(out.c[i * ND + d] + gl, ajPhiIp * invVolL);
        atomicAdd(out.c[i * ND + d] + gr, -(ajPhiIp * invVolR));
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/*  stream kernels: persistent CTAs, software-pipelined tile staging   */
/* ------------------------------------------------------------------ */

/* Round-1c per-phase cycle accounting (profiles/r01c_phase_cycles_tile192.txt)
 * showed the one-CTA-per-tile kernels spend 40-55 % of a CTA's life in the
 * stage: header read -> ~100 cycles of issue latency per TMA bulk copy (28 of
 * them for momentum, serialised in the SM's TMA unit whichever warp issues
 * them, profiles/r01d_*) -> dependent halo gather (three DRAM round trips).
 * The stream kernels keep 256-thread CTAs resident for the whole launch (two
 * or three per SM, so that one CTA's barriers and row reduction overlap the
 * other's physics) and run a software pipeline over the tiles
 * t_k = blockIdx.x + k * gridDim.x :
 *
 *   after the phase-1 barrier of tile k (node stage dead) the CTA issues
 *     - the node data of t_{k+1}: own range as 16-byte cp.async.cg (LDGSTS,
 *       8 cycles of issue per warp-op, spread over all warps), halo nodes as
 *       8-byte cp.async through the staged halo list;
 *     - the reduction plan of t_{k+1} into the other plan buffer;
 *     - the halo slot list of t_{k+2} and the 64-byte headers of t_{k+3}
 *       into small rings,
 *   and after the phase-2 barrier (edge results dead) the edge streams of
 *   t_{k+1}.  Every dependent address (header -> list -> halo data) is in
 *   shared memory a full tile time before it is needed.
 *
 * Same plan data, same physics, same write-once row ownership as the tile
 * kernels; the row reduction uses four lanes per row (partial sums combined by
 * a fixed xor-shuffle tree, so the result is still deterministic). */

constexpr int kStreamThreads = 256;
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kHdrRing = 4;
constexpr int kListRing = 2;

/* one warp copies `bytes` (multiple of 16, both ends 16-byte aligned) */
__device__ __forceinline__ void
warp_copy16(void* dstSmem, const void* src, uint32_t bytes, int lane)
{
  char* d = static_cast<char*>(dstSmem);
  const char* s = static_cast<const char*>(src);
  for (uint32_t q = (uint32_t)lane * 16u; q < bytes; q += 32u * 16u)
    cp_async16(d + q, s + q);
}

/* static shared state of a stream kernel */
struct StreamShared
{
  TileHdr hdr[kHdrRing];
  LsTileHdr lhdr[kHdrRing];
};

/* headers of this CTA's k-th tile -> ring slot k % kHdrRing (threads 0..7) */
template <bool LS>
__device__ __forceinline__ void
stream_issue_hdr(
  StreamShared& ss, const TileHdr* tiles, const LsTileHdr* ltiles, int k, int tile)
{
  const int t = threadIdx.x;
  if (t < 4)
    cp_async16(
      reinterpret_cast<char*>(&ss.hdr[k % kHdrRing]) + 16 * t,
      reinterpret_cast<const char*>(tiles + tile) + 16 * t);
  else if (LS && t < 8)
    cp_async16(
      reinterpret_cast<char*>(&ss.lhdr[k % kHdrRing]) + 16 * (t - 4),
      reinterpret_cast<const char*>(ltiles + tile) + 16 * (t - 4));
}

/* halo slot list of a tile -> ring slot */
__device__ __forceinline__ void
stream_issue_list(
  int32_t* s_list, const TileHdr& h, const int32_t* __restrict__ haloNodes)
{
  for (int i = threadIdx.x; i < h.nHalo; i += kStreamThreads)
    cp_async4(s_list + i, haloNodes + h.haloPtr + i);
}

/* node data of a tile: comp c of the own range by warp c % nWarps (16-byte
 * copies), halo nodes by per-thread 8-byte async copies through the staged
 * halo list */
template <int NC>
__device__ __forceinline__ void
stream_issue_nodes(
  double* s_node, int stride, const NodeComps& nc, const TileHdr& h,
  const int32_t* s_list, int warp, int lane)
{
  const uint32_t bNode = (uint32_t)h.nOwnPad * 8u;
  for (int c = warp; c < NC; c += kStreamWarps)
    warp_copy16(s_node + c * stride, nc.c[c] + h.node0, bNode, lane);
  for (int i = threadIdx.x; i < h.nHalo; i += kStreamThreads) {
    const int32_t g = s_list[i];
    double* dst = s_node + h.nOwnPad + i;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      cp_async8(dst + c * stride, nc.c[c] + g);
  }
}

/* shared-memory carve-up of ls_stream_kernel (sizes in bytes, 16-aligned) */
template <class P>
struct LsStreamSmem
{
  static constexpr int NEDGE = LsSmem<P>::NEDGE;
  int stride, resStride, lrLen, valsLen, ellLen, entLen, listLen, sliceLen;
  size_t nodeBytes, resBytes, planBytes, valsBytes, listBytes;
  __host__ __device__ LsStreamSmem(const MeshPlanDev& mp, const LsPlanDev& lp)
  {
    stride = mp.maxStaged;
    resStride = (mp.maxTileEdges + 1) & ~1;
    lrLen = (mp.maxTileEdges + 3) & ~3;
    valsLen = (lp.maxTileNnz + 3) & ~3;
    ellLen = (lp.maxTileEll + 3) & ~3;
    entLen = (lp.maxTileEnts + 3) & ~3;
    listLen = (mp.maxStaged + 3) & ~3;
    sliceLen = ((lp.maxTileEnts + 31) / 32 + 1 + 3) & ~3;
    nodeBytes = sizeof(double) * (size_t)P::NC * stride;
    resBytes = sizeof(double) * (size_t)NEDGE * resStride + 4u * (size_t)lrLen;
    planBytes = 4u * (size_t)ellLen + 12u * (size_t)entLen + 4u * (size_t)sliceLen;
    valsBytes = 12u * (size_t)valsLen;
    listBytes = 4u * (size_t)listLen;
  }
  __host__ __device__ size_t bytes() const
  {
    return nodeBytes + resBytes + 2 * planBytes + valsBytes + kListRing * listBytes;
  }
};

template <class P, int ND, int MINB>
__global__ void __launch_bounds__(kStreamThreads, MINB) ls_stream_kernel(
  const MeshPlanDev mp,
  const LsPlanDev lp,
  const NodeComps nc,
  const EdgeComps ec,
  const typename P::Opts o)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) StreamShared ss;
  using SM = LsStreamSmem<P>;
  const SM L(mp, lp);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int nMine = (mp.nTiles - (int)blockIdx.x + G - 1) / G;
  NW_PT_BEGIN_AT(8 + P::kPhaseId, 32 * 5);

  /* every region is an offset from the dynamic shared base, so that all
   * accesses stay shared-space LDS/STS */
  double* s_node = reinterpret_cast<double*>(smem_raw);
  double* s_res = reinterpret_cast<double*>(smem_raw + L.nodeBytes);
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(s_res + SM::NEDGE * L.resStride);
  unsigned char* s_planBase = smem_raw + L.nodeBytes + L.resBytes;
  double* s_vals = reinterpret_cast<double*>(s_planBase + 2 * L.planBytes);
  int32_t* s_delta = reinterpret_cast<int32_t*>(s_vals + L.valsLen);
  int32_t* s_listRing =
    reinterpret_cast<int32_t*>(s_planBase + 2 * L.planBytes + L.valsBytes);
  const int stride = L.stride;
  auto s_ell_of = [&](int b) {
    return reinterpret_cast<uint32_t*>(s_planBase + (size_t)b * L.planBytes);
  };

  /* edge input streams of this policy: area, [mdot], [pecfac] */
  constexpr int kMdot = ND;
  constexpr int kPec = ND + (P::kNeedsMdot ? 1 : 0);
  const bool hasPec = P::kNeedsPec && ec.pecfac != nullptr;
  const int nin = kPec + (hasPec ? 1 : 0);

  auto tile_of = [&](int k) { return (int)blockIdx.x + k * G; };

  /* node data + reduction plan of the k-th tile (headers and halo list of
   * that tile are already visible in their rings) */
  auto issue_nodes_plan = [&](int k) {
    const TileHdr& h = ss.hdr[k % kHdrRing];
    const LsTileHdr& lh = ss.lhdr[k % kHdrRing];
    stream_issue_nodes<P::NC>(
      s_node, stride, nc, h, s_listRing + (k % kListRing) * L.listLen, warp,
      lane);
    uint32_t* s_ell = s_ell_of(k & 1);
    EntInfo* s_ent = reinterpret_cast<EntInfo*>(s_ell + L.ellLen);
    int32_t* s_row = reinterpret_cast<int32_t*>(s_ent + L.entLen);
    int32_t* s_go = s_row + L.entLen;
    int32_t* s_slice = s_go + L.entLen;
    const uint32_t bEnt = round16((uint32_t)lh.nEnts * 4u);
    /* the node copies went to warps 0 .. NC-1 (mod nWarps): start the plan
     * copies where they stopped */
    const int w0 = P::NC % kStreamWarps;
    if (warp == (w0 + 0) % kStreamWarps)
      warp_copy16(s_ell, lp.heEll + lh.ellPtr, (uint32_t)lh.ellLen * 4u, lane);
    if (warp == (w0 + 1) % kStreamWarps)
      warp_copy16(s_ent, lp.entInfo + lh.entPtr, bEnt, lane);
    if (warp == (w0 + 2) % kStreamWarps)
      warp_copy16(s_row, lp.entRhsRow + lh.entPtr, bEnt, lane);
    if (warp == (w0 + 3) % kStreamWarps) {
      warp_copy16(s_go, lp.entGo + lh.entPtr, bEnt, lane);
      const int nSl = (lh.nEnts + 31) / 32 + 1;
      for (int i = lane; i < nSl; i += 32)
        cp_async4(s_slice + i, lp.sliceOff + lh.slicePtr + i);
    }
  };
  /* edge streams of the k-th tile: packed (L,R) records + input components */
  auto issue_edges = [&](int k) {
    const TileHdr& h = ss.hdr[k % kHdrRing];
    const uint32_t bEdge = round16((uint32_t)h.nEdges * 8u);
    for (int i = warp; i < nin + 1; i += kStreamWarps) {
      if (i == nin) {
        warp_copy16(s_lr, mp.lr + h.edge0, round16((uint32_t)h.nEdges * 4u), lane);
      } else {
        const double* src =
          i < ND ? ec.area[i < ND ? i : 0]
                 : (P::kNeedsMdot && i == kMdot ? ec.mdot : ec.pecfac);
        warp_copy16(s_res + i * L.resStride, src + h.edge0, bEdge, lane);
      }
    }
  };
  auto issue_list = [&](int k) {
    stream_issue_list(
      s_listRing + (k % kListRing) * L.listLen, ss.hdr[k % kHdrRing],
      mp.haloNodes);
  };

  /* ---- prologue: fill the rings ---- */
  for (int k = 0; k < 3 && k < nMine; ++k)
    stream_issue_hdr<true>(ss, mp.tiles, lp.tiles, k, tile_of(k));
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (int k = 0; k < 2 && k < nMine; ++k)
    issue_list(k);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  if (nMine > 0) {
    issue_nodes_plan(0);
    issue_edges(0);
  }
  cp_async_commit();

  for (int k = 0; k < nMine; ++k) {
    NW_PT_MARK(); /* 0: loop overhead (first lap: prologue) */
    /* everything issued for tile k (and the ring entries issued during the
     * previous iteration) has landed */
    cp_async_wait_all();
    NW_PT_MARK(); /* 1: own cp.async landed */
    __syncthreads();
    NW_PT_MARK(); /* 2: top barrier */

    const TileHdr& h = ss.hdr[k % kHdrRing];
    const LsTileHdr& lh = ss.lhdr[k % kHdrRing];
    const uint32_t* s_ell = s_ell_of(k & 1);
    const EntInfo* s_ent = reinterpret_cast<const EntInfo*>(s_ell + L.ellLen);
    const int32_t* s_row = reinterpret_cast<const int32_t*>(s_ent + L.entLen);
    const int32_t* s_go = s_row + L.entLen;
    const int32_t* s_slice = s_go + L.entLen;

    /* row staging starts from zero: a slot no local half-edge touches (a
     * column that only other algorithms fill) is written as 0 */
    for (int e = tid; e < lh.nnz; e += kStreamThreads)
      s_vals[e] = 0.0;

    /* ---- phase 1: per-edge physics out of shared memory ---- */
    {
      const SmemLd ld{s_node, stride};
      for (int j = tid; j < h.nEdges; j += kStreamThreads) {
        const uint32_t v = s_lr[j];
        const int l = (int)(v & 0xffffu), r = (int)(v >> 16);
        double av[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
          av[d] = s_res[d * L.resStride + j];
        double mdot = 0.0, pecfac = 0.0;
        if (P::kNeedsMdot)
          mdot = s_res[kMdot * L.resStride + j];
        if (hasPec)
          pecfac = s_res[kPec * L.resStride + j];
        double res[P::NRES];
        P::compute(ld, l, r, av, mdot, pecfac, o, res);
#pragma unroll
        for (int q = 0; q < P::NRES; ++q)
          s_res[q * L.resStride + j] = res[q];
      }
    }
    NW_PT_MARK(); /* 3: zero + phase 1 */
    __syncthreads();
    NW_PT_MARK(); /* 4: phase-1 barrier */

    /* the node stage is dead: start pulling the next tile in */
    if (k + 3 < nMine)
      stream_issue_hdr<true>(ss, mp.tiles, lp.tiles, k + 3, tile_of(k + 3));
    if (k + 2 < nMine)
      issue_list(k + 2);
    if (k + 1 < nMine)
      issue_nodes_plan(k + 1);
    cp_async_commit();
    NW_PT_MARK(); /* 5: issue nodes + plan of the next tile */

    /* ---- phase 2: four lanes per row ---- */
    {
      const int sub = lane & 3;
      for (int rbase = warp * 8; rbase < lh.nEnts; rbase += kStreamWarps * 8) {
        const int row = rbase + (lane >> 2);
        const bool active = row < lh.nEnts;
        double diag = 0.0;
        double rhs[P::NR];
#pragma unroll
        for (int d = 0; d < P::NR; ++d)
          rhs[d] = 0.0;
        EntInfo ei{0, 0, 0};
        const uint32_t* hp = s_ell;
        int W = 0;
        bool sawDup = false;
        if (active) {
          const int sl = row >> 5;
          const int o0 = s_slice[sl], o1 = s_slice[sl + 1];
          hp = s_ell + o0 + (row & 31);
          W = (o1 - o0) >> 5;
          ei = s_ent[row];
          double* vrow = s_vals + ei.base;
          for (int w = sub; w < W; w += 4) {
            const uint32_t hv = hp[w * 32];
            if (hv & kHeValid) {
              double dg, off, rr[P::NR];
              P::contrib(
                he_side(hv), s_res, L.resStride, (int)he_edge(hv), dg, off, rr);
              diag += dg;
#pragma unroll
              for (int d = 0; d < P::NR; ++d)
                rhs[d] += rr[d];
              if (hv & kHeDup)
                sawDup = true;
              else
                vrow[he_k(hv)] = off;
            }
          }
        }
        /* fixed combination tree over the four lanes of the row */
        diag += __shfl_xor_sync(kFull, diag, 1);
        diag += __shfl_xor_sync(kFull, diag, 2);
#pragma unroll
        for (int d = 0; d < P::NR; ++d) {
          rhs[d] += __shfl_xor_sync(kFull, rhs[d], 1);
          rhs[d] += __shfl_xor_sync(kFull, rhs[d], 2);
        }
        const bool anyDup = __any_sync(kFull, sawDup);
        if (anyDup) {
          /* periodic aliases: later members of a (row, column) group are added
           * in list order by one lane, after the group's first store */
          __syncwarp();
          if (active && sub == 0) {
            double* vrow = s_vals + ei.base;
            for (int w = 0; w < W; ++w) {
              const uint32_t hv = hp[w * 32];
              if ((hv & kHeValid) && (hv & kHeDup)) {
                double dg, off, rr[P::NR];
                P::contrib(
                  he_side(hv), s_res, L.resStride, (int)he_edge(hv), dg, off, rr);
                vrow[he_k(hv)] += off;
              }
            }
          }
        }
        if (active) {
          if (sub == 0)
            s_vals[ei.base + ei.diagK] = diag;
          const int32_t delta = s_go[row] - (int32_t)ei.base;
          for (int q = sub; q < (int)ei.nnz; q += 4)
            s_delta[ei.base + q] = delta;
          if (sub < P::NR) {
            double mine = rhs[0];
#pragma unroll
            for (int d = 1; d < P::NR; ++d)
              if (sub == d)
                mine = rhs[d];
            lp.rhs[(int64_t)sub * lp.rhsStride + s_row[row]] = mine;
          }
        }
      }
    }
    NW_PT_MARK(); /* 6: phase 2 */
    __syncthreads();
    NW_PT_MARK(); /* 7: phase-2 barrier */

    /* the edge results are dead: pull the next tile's edge streams */
    if (k + 1 < nMine)
      issue_edges(k + 1);
    cp_async_commit();
    NW_PT_MARK(); /* 8: issue edge streams of the next tile */

    /* ---- phase 3: coalesced copy-out, every value written exactly once ---- */
    for (int e = tid; e < lh.nnz; e += kStreamThreads)
      lp.values[e + s_delta[e]] = s_vals[e];
    NW_PT_MARK(); /* 9: phase 3 */
    NW_PT_LAP();
  }
}


/* ------------------------------------------------------------------ */
/*  utility kernels                                                    */
/* ------------------------------------------------------------------ */

__global__ void
node_gather_kernel(
  const double* __restrict__ src,
  int ncomp,
  const int32_t* __restrict__ nodeOfSlot,
  int64_t nSlots,
  double* __restrict__ dst)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSlots)
    return;
  const int32_t n = nodeOfSlot[i];
  for (int c = 0; c < ncomp; ++c)
    dst[(int64_t)c * nSlots + i] = (n >= 0) ? src[(int64_t)n * ncomp + c] : 0.0;
}

__global__ void
node_scatter_kernel(
  const double* __restrict__ src,
  int ncomp,
  const int32_t* __restrict__ nodeOfSlot,
  int64_t nSlots,
  double* __restrict__ dst)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSlots)
    return;
  const int32_t n = nodeOfSlot[i];
  if (n < 0)
    return;
  for (int c = 0; c < ncomp; ++c)
    dst[(int64_t)n * ncomp + c] = src[(int64_t)c * nSlots + i];
}

__global__ void
edge_scatter_kernel(
  const double* __restrict__ src,
  int ncomp,
  const int32_t* __restrict__ primarySlot,
  int64_t nEdges,
  int64_t slotStride,
  double* __restrict__ dst)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEdges)
    return;
  const int32_t s = primarySlot[e];
  for (int c = 0; c < ncomp; ++c)
    dst[e * ncomp + c] = src[(int64_t)c * slotStride + s];
}

__global__ void
fill_kernel(double* p, int64_t n, double v)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    p[i] = v;
}

__global__ void
row_init_kernel(
  const int32_t* __restrict__ rows,
  int nRows,
  const int64_t* __restrict__ rowPtr,
  const uint8_t* __restrict__ isPeriodic,
  double* values,
  double* rhs,
  int64_t rhsStride,
  int nRhs)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nRows)
    return;
  const int32_t r = rows[i];
  const int64_t a = rowPtr[r], b = rowPtr[r + 1];
  for (int64_t k = a; k < b; ++k)
    values[k] = 0.0;
  if (isPeriodic[i])
    values[a] = 1.0;
  for (int d = 0; d < nRhs; ++d)
    rhs[(int64_t)d * rhsStride + r] = 0.0;
}

/* ---- GeometryInteriorAlg<Hex8>: dual nodal volumes and edge area vectors ----
 * (src/ngp_algorithms/GeometryInteriorAlg.C:72-112, 165-225).  One thread per
 * element.  The 27 points of the 8-hex subdivision (corners, edge / face /
 * body mid points, Hex8GeometryFunctions.h:258-325) are averages of corner
 * subsets, kept as bit masks; sub-control-volume volumes use Grandy's 24
 * triangle formula (:83-158), sub-control-surface areas the 4-triangle fan
 * about the facet mid point (:33-81), both in the reference's summation order.
 */
__constant__ unsigned char kHexSubMask[27] = {
  0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, /* corners 0..7 */
  0x03, 0x06, 0x0c, 0x09, 0x0f,                   /* face 0: edges + centre */
  0x30, 0x60, 0xc0, 0x90, 0xf0,                   /* face 1 */
  0x22, 0x11, 0x33,                               /* face 2 */
  0x88, 0x44, 0xcc,                               /* face 3 */
  0x66, 0x99,                                     /* faces 4, 5 */
  0xff};                                          /* centroid */
__constant__ unsigned char kHexScvTable[8][8] = {
  {0, 8, 12, 11, 19, 20, 26, 25},  {8, 1, 9, 12, 20, 18, 24, 26},
  {12, 9, 2, 10, 26, 24, 22, 23},  {11, 12, 10, 3, 25, 26, 23, 21},
  {19, 20, 26, 25, 4, 13, 17, 16}, {20, 18, 24, 26, 13, 5, 14, 17},
  {26, 24, 22, 23, 17, 14, 6, 15}, {25, 26, 23, 21, 16, 17, 15, 7}};
__constant__ unsigned char kHexScsTable[12][4] = {
  {20, 8, 12, 26},  {24, 9, 12, 26},  {10, 12, 26, 23}, {11, 25, 26, 12},
  {13, 20, 26, 17}, {17, 14, 24, 26}, {17, 15, 23, 26}, {16, 17, 26, 25},
  {19, 20, 26, 25}, {20, 18, 24, 26}, {22, 23, 26, 24}, {21, 25, 26, 23}};
__constant__ unsigned char kGrandyFace[6][4] = {
  {0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 3, 7}};
__constant__ unsigned char kGrandyTri[24][3] = {
  {0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},
  {4, 9, 7},  {4, 5, 9},  {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10},
  {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11}, {6, 12, 2}, {5, 12, 6},
  {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};

__global__ void __launch_bounds__(128) geometry_hex8_kernel(
  int64_t nElems, const int32_t* __restrict__ elemSlots /* [n][8] */,
  const int32_t* __restrict__ elemEdges /* [n][12]: 2*slot + negate, -1 none */,
  const unsigned char* __restrict__ owned, const double* __restrict__ x,
  int64_t xStride, double* dualVol, double* area, int64_t areaStride)
{
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nElems)
    return;
  double v[27][3];
  {
    double c[8][3];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int64_t sl = elemSlots[8 * el + n];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        c[n][d] = x[(int64_t)d * xStride + sl];
    }
    for (int p = 0; p < 27; ++p) {
      const unsigned m = kHexSubMask[p];
      const int cnt = __popc(m);
      /* the reference sums the subset in ascending corner order and scales by
       * 1, 0.5, 0.25 or 0.125 (exact powers of two) */
      const double w = cnt == 1 ? 1.0 : (cnt == 2 ? 0.5 : (cnt == 4 ? 0.25 : 0.125));
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double acc = 0.0;
        bool first = true;
        for (int n = 0; n < 8; ++n)
          if (m & (1u << n)) {
            acc = first ? c[n][d] : acc + c[n][d];
            first = false;
          }
        v[p][d] = cnt == 1 ? acc : w * acc;
      }
    }
    /* edge mid point 11 is (c3 + c0) and 16 is (c7 + c4) in the reference:
     * addition is commutative, the bits are the same */
  }
  if (dualVol && (!owned || owned[el])) {
    for (int ip = 0; ip < 8; ++ip) {
      double cv[14][3];
      for (int n = 0; n < 8; ++n)
        for (int d = 0; d < 3; ++d)
          cv[n][d] = v[kHexScvTable[ip][n]][d];
      for (int k = 0; k < 6; ++k)
        for (int d = 0; d < 3; ++d)
          cv[8 + k][d] =
            0.25 * (cv[kGrandyFace[k][0]][d] + cv[kGrandyFace[k][1]][d] +
                    cv[kGrandyFace[k][2]][d] + cv[kGrandyFace[k][3]][d]);
      double vol = 0.0;
      for (int k = 0; k < 24; ++k) {
        const int p = kGrandyTri[k][0], q = kGrandyTri[k][1], r = kGrandyTri[k][2];
        const double m0 = cv[p][0] + cv[q][0] + cv[r][0];
        const double m1 = cv[p][1] + cv[q][1] + cv[r][1];
        const double m2 = cv[p][2] + cv[q][2] + cv[r][2];
        const double d0 = (cv[q][1] - cv[p][1]) * (cv[r][2] - cv[p][2]) -
                          (cv[r][1] - cv[p][1]) * (cv[q][2] - cv[p][2]);
        const double d1 = (cv[r][0] - cv[p][0]) * (cv[q][2] - cv[p][2]) -
                          (cv[q][0] - cv[p][0]) * (cv[r][2] - cv[p][2]);
        const double d2 = (cv[q][0] - cv[p][0]) * (cv[r][1] - cv[p][1]) -
                          (cv[r][0] - cv[p][0]) * (cv[q][1] - cv[p][1]);
        vol += m0 * d0 + m1 * d1 + m2 * d2;
      }
      vol /= 18.0;
      atomicAdd(dualVol + elemSlots[8 * el + ip], vol);
    }
  }
  if (!area)
    return;
  for (int ip = 0; ip < 12; ++ip) {
    const int32_t code = elemEdges[12 * el + ip];
    if (code < 0)
      continue;
    const double* q0 = v[kHexScsTable[ip][0]];
    double xm[3], r1[3], a[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < 3; ++d) {
      xm[d] = 0.25 * (v[kHexScsTable[ip][0]][d] + v[kHexScsTable[ip][1]][d] +
                      v[kHexScsTable[ip][2]][d] + v[kHexScsTable[ip][3]][d]);
      r1[d] = q0[d] - xm[d];
    }
    for (int it = 0; it < 4; ++it) {
      const double* qt = v[kHexScsTable[ip][(it + 1) & 3]];
      const double r2[3] = {qt[0] - xm[0], qt[1] - xm[1], qt[2] - xm[2]};
      a[0] += r1[1] * r2[2] - r2[1] * r1[2];
      a[1] += r1[2] * r2[0] - r2[2] * r1[0];
      a[2] += r1[0] * r2[1] - r2[0] * r1[1];
      r1[0] = r2[0];
      r1[1] = r2[1];
      r1[2] = r2[2];
    }
    const double sg = (code & 1) ? -0.5 : 0.5;
    const int64_t slot = code >> 1;
    for (int d = 0; d < 3; ++d)
      atomicAdd(area + (int64_t)d * areaStride + slot, a[d] * sg);
  }
}

/* GeometryInteriorAlg<Quad4_2D> (src/master_element/Quad42DCVFEM.C:139-200,
 * 384-445): four 2x2-Gauss sub-control-volume areas, four sub-control-surface
 * normals from the element centre to the face mid points */
__global__ void __launch_bounds__(128) geometry_quad4_kernel(
  int64_t nElems, const int32_t* __restrict__ elemSlots /* [n][4] */,
  const int32_t* __restrict__ elemEdges /* [n][4] */,
  const unsigned char* __restrict__ owned, const double* __restrict__ x,
  int64_t xStride, double* dualVol, double* area, int64_t areaStride)
{
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nElems)
    return;
  double cx[4], cy[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int64_t sl = elemSlots[4 * el + n];
    cx[n] = x[sl];
    cy[n] = x[xStride + sl];
  }
  if (dualVol && (!owned || owned[el])) {
    const double gp[2] = {-0.144337567, 0.144337567};
    const double cv[2] = {-0.25, 0.25};
    /* corner k of the parent square: xi sign pattern (-,+,+,-), eta (-,-,+,+) */
    const int sx[4] = {0, 1, 1, 0}, sy[4] = {0, 0, 1, 1};
    for (int ki = 0; ki < 4; ++ki) {
      double vol = 0.0;
      for (int kq = 0; kq < 4; ++kq) {
        const double xi = cv[sx[ki]] + gp[sx[kq]];
        const double eta = cv[sy[ki]] + gp[sy[kq]];
        const double d0[4] = {-(0.5 - eta), (0.5 - eta), (0.5 + eta), -(0.5 + eta)};
        const double d1[4] = {-(0.5 - xi), -(0.5 + xi), (0.5 + xi), (0.5 - xi)};
        double xs1 = 0.0, xs2 = 0.0, ys1 = 0.0, ys2 = 0.0;
#pragma unroll
        for (int kn = 0; kn < 4; ++kn) {
          xs1 += d0[kn] * cx[kn];
          xs2 += d1[kn] * cx[kn];
          ys1 += d0[kn] * cy[kn];
          ys2 += d1[kn] * cy[kn];
        }
        vol += (xs1 * ys2 - ys1 * xs2) * 0.0625;
      }
      atomicAdd(dualVol + elemSlots[4 * el + ki], vol);
    }
  }
  if (!area)
    return;
  const double x1 = (cx[0] + cx[1] + cx[2] + cx[3]) * 0.25;
  const double y1 = (cy[0] + cy[1] + cy[2] + cy[3]) * 0.25;
  for (int f = 0; f < 4; ++f) {
    const int32_t code = elemEdges[4 * el + f];
    if (code < 0)
      continue;
    const int a = f, b = (f + 1) & 3;
    const double x2 = (cx[a] + cx[b]) * 0.5, y2 = (cy[a] + cy[b]) * 0.5;
    /* surfaces 0-2: (-(dy), dx); surface 3 points from node 0 to node 3 */
    const double ax = f < 3 ? -(y2 - y1) : (y2 - y1);
    const double ay = f < 3 ? (x2 - x1) : -(x2 - x1);
    const double sg = (code & 1) ? -1.0 : 1.0;
    const int64_t slot = code >> 1;
    atomicAdd(area + slot, ax * sg);
    atomicAdd(area + areaStride + slot, ay * sg);
  }
}

/* a cut edge has a second tile-edge slot: keep it equal to the primary one */
__global__ void
edge_mirror_kernel(
  const int32_t* __restrict__ primarySlot, const int32_t* __restrict__ secondSlot,
  int64_t nEdges, int ncomp, int64_t stride, double* f)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEdges || secondSlot[e] < 0)
    return;
  for (int c = 0; c < ncomp; ++c)
    f[(int64_t)c * stride + secondSlot[e]] = f[(int64_t)c * stride + primarySlot[e]];
}

/* time-derivative node kernels (src/node_kernels/{Scalar,Momentum,Continuity}
 * MassBDFNodeKernel.C): one thread per (node, row), plain read-modify-write --
 * every selected node owns its rows */
template <int KIND>
__global__ void
mass_bdf_node_kernel(
  int ndim, const int64_t* __restrict__ rows, int64_t nRows,
  const MassBdfFields f, double dt, double gamma1, double gamma2,
  double gamma3, double* values, double* rhs, int64_t rhsStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nRows)
    return;
  const int64_t slot = rows[4 * t], diagOff = rows[4 * t + 1],
                rr = rows[4 * t + 2], dof = rows[4 * t + 3];
  const double rhoNm1 = f.rho[0][slot], rhoN = f.rho[1][slot],
               rhoNp1 = f.rho[2][slot];
  const double dnvNm1 = f.dnv[0][slot], dnvN = f.dnv[1][slot],
               dnvNp1 = f.dnv[2][slot];
  if (KIND == NW_MASS_SCALAR) {
    const double qNm1 = f.q[0][slot], qN = f.q[1][slot], qNp1 = f.q[2][slot];
    const double lhsTime = gamma1 * rhoNp1 * dnvNp1 / dt;
    rhs[rr] -= (gamma1 * rhoNp1 * qNp1 * dnvNp1 + gamma2 * qN * rhoN * dnvN +
                gamma3 * qNm1 * rhoNm1 * dnvNm1) /
               dt;
    values[diagOff] += lhsTime;
  } else if (KIND == NW_MASS_MOMENTUM) {
    const double lhsfac = gamma1 * rhoNp1 * dnvNp1 / dt;
    const int i0 = dof < 0 ? 0 : (int)dof, i1 = dof < 0 ? ndim : (int)dof + 1;
    for (int i = i0; i < i1; ++i) {
      const int64_t o = (int64_t)i * f.fieldStride + slot;
      const double uNm1 = f.q[0][o], uN = f.q[1][o], uNp1 = f.q[2][o];
      const double dpdx = f.dpdx[o];
      const int64_t ri = dof < 0 ? (int64_t)i * rhsStride + rr : rr;
      rhs[ri] += -(gamma1 * rhoNp1 * uNp1 * dnvNp1 + gamma2 * rhoN * uN * dnvN +
                   gamma3 * rhoNm1 * uNm1 * dnvNm1) /
                   dt -
                 dpdx * dnvNp1;
    }
    /* lhs(i,i) += lhsfac; the UVW system keeps the x-x entry only */
    values[diagOff] += lhsfac;
  } else {
    rhs[rr] -= (gamma1 * rhoNp1 * dnvNp1 + gamma2 * rhoN * dnvN +
                gamma3 * rhoNm1 * dnvNm1) /
               dt * (gamma1 / dt);
  }
}

/* WallDistNodeKernel (src/node_kernels/WallDistNodeKernel.C:34-43): rhs += V */
__global__ void
wall_dist_node_kernel(
  const int64_t* __restrict__ rows, int64_t nRows,
  const double* __restrict__ dualVol, double* rhs)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nRows)
    rhs[rows[4 * t + 2]] += dualVol[rows[4 * t]];
}

/* CoeffApplier::resetRows: rows[t] = (value offset, length, diagonal position
 * or -1, rhs row): zero the row, diagonal = diagValue, every rhs column =
 * rhsResidual (src/HypreLinearSystem.C:2262-2315) */
__global__ void
reset_rows_kernel(
  const int64_t* __restrict__ rows, int64_t nRows, double diagValue,
  double rhsResidual, double* values, double* rhs, int64_t rhsStride, int nRhs)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nRows)
    return;
  const int64_t a = rows[4 * t], len = rows[4 * t + 1], dk = rows[4 * t + 2],
                rr = rows[4 * t + 3];
  for (int64_t k = 0; k < len; ++k)
    values[a + k] = (k == dk) ? diagValue : 0.0;
  for (int d = 0; d < nRhs; ++d)
    rhs[(int64_t)d * rhsStride + rr] = rhsResidual;
}

/* applyDirichletBCs (src/HypreLinearSystem.C:2446-2456): rows[t] = (value
 * offset of the row's first entry, rhs row, rhs column, field slot, field
 * component): first entry = 1, rhs = bc - solution */
__global__ void
dirichlet_rows_kernel(
  const int64_t* __restrict__ rows, int64_t nRows,
  const double* __restrict__ solution, const double* __restrict__ bc,
  int64_t fieldStride, double* values, double* rhs, int64_t rhsStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nRows)
    return;
  const int64_t a = rows[5 * t], rr = rows[5 * t + 1], col = rows[5 * t + 2],
                slot = rows[5 * t + 3], comp = rows[5 * t + 4];
  values[a] = 1.0;
  rhs[col * rhsStride + rr] =
    bc[comp * fieldStride + slot] - solution[comp * fieldStride + slot];
}

/* fixed-shape two-level reduction: deterministic for a given (n, nPartial) */
__global__ void __launch_bounds__(256) norm2_partial_kernel(
  const double* __restrict__ rhs, int64_t n, int64_t stride, double* partial)
{
  __shared__ double sm[256];
  const int d = blockIdx.y;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t a = (int64_t)blockIdx.x * per;
  const int64_t b = a + per < n ? a + per : n;
  double s = 0.0;
  for (int64_t i = a + threadIdx.x; i < b; i += 256) {
    const double v = rhs[(int64_t)d * stride + i];
    s += v * v;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w)
      sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    partial[(int64_t)d * gridDim.x + blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256)
norm2_final_kernel(const double* __restrict__ partial, int nPartial, double* out)
{
  __shared__ double sm[256];
  const int d = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < nPartial; i += 256)
    s += partial[(int64_t)d * nPartial + i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w)
      sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    out[d] = sm[0];
}

__device__ __forceinline__ int64_t
lower_bound_dev(const int64_t* a, int64_t n, int64_t v)
{
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* Generic CoeffApplier::operator() (include/LinearSystem.h:62-70): one thread
 * per entity; the reference's sort + linear column walk becomes a binary
 * search per entry (same destination, see tests/test_graph_parity.py). */
__global__ void
sum_into_kernel(
  int64_t nEnt,
  int npe,
  int numDof,
  const int32_t* __restrict__ entNodes,
  const int64_t* __restrict__ nodeHid,
  const double* __restrict__ lhs,
  const double* __restrict__ rhsIn,
  int64_t iLower,
  int64_t iUpper,
  int64_t nRowsOwned,
  int64_t nnzOwned,
  const int64_t* __restrict__ rowStartOwned,
  const int64_t* __restrict__ rowStartShared,
  const int64_t* __restrict__ rowIndicesShared,
  int64_t nRowsShared,
  const int64_t* __restrict__ cols,
  const int64_t* __restrict__ skipped,
  int64_t nSkipped,
  int uvwDim,
  double* values,
  double* rhs,
  int64_t rhsStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nEnt)
    return;
  const int ld = uvwDim > 0 ? uvwDim : numDof; /* dofs per node in lhs/rhs */
  const int n = npe * ld;
  const double* L = lhs + t * (int64_t)n * n;
  const double* R = rhsIn + t * (int64_t)n;
  for (int i = 0; i < npe; ++i) {
    const int64_t hid = nodeHid[entNodes[t * npe + i]];
    const int64_t first = hid * numDof;
    if (nSkipped) {
      const int64_t p = lower_bound_dev(skipped, nSkipped, first);
      if (p < nSkipped && skipped[p] == first)
        continue;
    }
    for (int d = 0; d < numDof; ++d) {
      const int64_t row = first + d;
      int64_t base, len, lrow;
      if (row >= iLower && row <= iUpper) {
        lrow = row - iLower;
        base = rowStartOwned[lrow];
        len = rowStartOwned[lrow + 1] - base;
      } else {
        const int64_t p = lower_bound_dev(rowIndicesShared, nRowsShared, row);
        if (p >= nRowsShared || rowIndicesShared[p] != row)
          continue;
        lrow = nRowsOwned + p;
        base = nnzOwned + rowStartShared[p];
        len = rowStartShared[p + 1] - rowStartShared[p];
      }
      const int ii = uvwDim > 0 ? i * uvwDim : i * numDof + d;
      for (int k = 0; k < npe; ++k) {
        const int64_t hk = nodeHid[entNodes[t * npe + k]];
        for (int dd = 0; dd < numDof; ++dd) {
          const int64_t col = hk * numDof + dd;
          const int64_t p = lower_bound_dev(cols + base, len, col);
          if (p < len && cols[base + p] == col) {
            const int kk = uvwDim > 0 ? k * uvwDim : k * numDof + dd;
            atomicAdd(values + base + p, L[ii * n + kk]);
          }
        }
      }
      if (uvwDim > 0) {
        for (int q = 0; q < uvwDim; ++q)
          atomicAdd(rhs + (int64_t)q * rhsStride + lrow, R[ii + q]);
      } else
        atomicAdd(rhs + lrow, R[ii]);
    }
  }
}

__global__ void
pack_kernel(
  const double* __restrict__ src,
  const int64_t* __restrict__ idx,
  int64_t n,
  double* __restrict__ dst)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dst[i] = src[idx[i]];
}

/* idx entries are unique within one call (one slot per received entry of one
 * neighbour), so a plain read-modify-write is race-free and the sum order is
 * the fixed neighbour order of the caller */
__global__ void
unpack_add_kernel(
  const double* __restrict__ src,
  const int64_t* __restrict__ idx,
  int64_t n,
  double* dst)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] >= 0)
    dst[idx[i]] += src[i];
}

__global__ void
scatter_assign_kernel(
  const double* __restrict__ src,
  const int64_t* __restrict__ idx,
  int64_t n,
  double* dst)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] >= 0)
    dst[idx[i]] = src[i];
}

/* halo exchange, all peers and all components in one launch.
 * buffer element of concatenated entry g, component c:
 *   buf[g * entStride + c * compStride] */
__global__ void
pack_multi_kernel(
  const double* __restrict__ src, int64_t srcCompStride, int nc,
  const int64_t* __restrict__ idx, int64_t n, double* __restrict__ buf,
  int64_t entStride, int64_t compStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nc)
    return;
  const int64_t g = t / nc;
  const int c = (int)(t - g * nc);
  buf[g * entStride + c * compStride] = src[(int64_t)c * srcCompStride + idx[g]];
}

__global__ void
scatter_multi_kernel(
  const double* __restrict__ buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* __restrict__ idx, int64_t n, double* __restrict__ dst,
  int64_t dstCompStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nc)
    return;
  const int64_t g = t / nc;
  const int c = (int)(t - g * nc);
  if (idx[g] >= 0)
    dst[(int64_t)c * dstCompStride + idx[g]] = buf[g * entStride + c * compStride];
}

/* owner-side sum: destination u receives the buffer entries pos[ptr[u] ..
 * ptr[u+1]) in that (ascending peer) order -- race-free when several peers
 * share one destination, and a fixed summation order */
__global__ void
accumulate_multi_kernel(
  const double* __restrict__ buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* __restrict__ dstIdx, const int64_t* __restrict__ ptr,
  const int64_t* __restrict__ pos, int64_t nDst, double* dst,
  int64_t dstCompStride)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nDst * nc)
    return;
  const int64_t u = t / nc;
  const int c = (int)(t - u * nc);
  double* d = dst + (int64_t)c * dstCompStride + dstIdx[u];
  double v = *d;
  for (int64_t q = ptr[u]; q < ptr[u + 1]; ++q)
    v += buf[pos[q] * entStride + c * compStride];
  *d = v;
}

/* ---- peer-memory halo exchange (NVLink, CUDA IPC windows) ---- */

__device__ __forceinline__ void
st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long
ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

/* end of a push kernel: when the last block has made its remote stores
 * visible, publish `epoch` in every peer's flag word for this rank */
__device__ __forceinline__ void
p2p_signal(const P2pDev& pp)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(pp.sync, 1u) + 1u;
    if (done == gridDim.x) {
      pp.sync[0] = 0u; /* next launch on this stream starts from zero */
      __threadfence_system();
      for (int i = 0; i < pp.nPeers; ++i)
        st_release_sys(pp.peerFlags[pp.peers[i]] + pp.myRank, pp.epoch);
    }
  }
}

/* start of a pull kernel: every peer's push of this epoch has landed in the
 * own window.  Bounded spin (pp.timeoutCycles): a peer that never arrives sets
 * the error word and the WHOLE kernel returns without touching the window --
 * nothing stale is ever accumulated; the host reports NW_ERR_COMM at its next
 * synchronisation point (p2p_check_error).  Returns false on timeout (or when
 * an earlier exchange of this context already failed). */
__device__ __forceinline__ bool
p2p_wait(const P2pDev& pp)
{
  int bad = 0;
  if ((int)threadIdx.x < pp.nPeers) {
    const unsigned long long* f = pp.myFlags + pp.peers[threadIdx.x];
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < pp.epoch) {
      if (clock64() - t0 > pp.timeoutCycles) {
        atomicExch(pp.sync + 1, 1u);
        bad = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  if (threadIdx.x == 0 && *(volatile unsigned*)(pp.sync + 1) != 0u)
    bad = 1; /* the data of a failed exchange is never consumed later either */
  return __syncthreads_or(bad) == 0;
}

/* pull kernel behind a kernel with a fused push: that kernel has completed,
 * its remote stores have landed -- the first block publishes the epoch, then
 * everybody waits for the neighbours' */
__device__ __forceinline__ bool
p2p_signal_then_wait(const P2pDev& pp, int signal)
{
  if (signal && blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    for (int i = 0; i < pp.nPeers; ++i)
      st_release_sys(pp.peerFlags[pp.peers[i]] + pp.myRank, pp.epoch);
  }
  return p2p_wait(pp);
}

/* nodal push: entry g of the concatenated send list, component c ->
 * window of rank sendPeer[g] at entry sendDst[g] */
__global__ void __launch_bounds__(256) p2p_push_nodal_kernel(
  const double* __restrict__ base, int64_t stride, int nc,
  const int64_t* __restrict__ sendIdx, const int32_t* __restrict__ sendPeer,
  const int64_t* __restrict__ sendDst, int64_t n, const P2pDev pp)
{
  const int64_t total = n * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = t / nc;
    const int c = (int)(t - g * nc);
    pp.peerWindow[sendPeer[g]][pp.winOff + sendDst[g] * nc + c] =
      base[(int64_t)c * stride + sendIdx[g]];
  }
  p2p_signal(pp);
}

/* nodal pull: own partial + the other sharer's partial (two sharers per node:
 * a + b on one side, b + a on the other -- the same bits) */
__global__ void __launch_bounds__(256) p2p_pull_nodal_kernel(
  const CompPtrs comps, int nc, const int64_t* __restrict__ recvIdx,
  int64_t n, const P2pDev pp, int signal)
{
  if (!p2p_signal_then_wait(pp, signal))
    return;
  const double* win = pp.myWindow + pp.winOff;
  const int64_t total = n * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = t / nc;
    const int c = (int)(t - g * nc);
    double* d = comps.c[c] + recvIdx[g];
    *d += __ldcg(win + t);
  }
}

/* copy_owned_to_shared: same push; the receiver overwrites its ghost copies
 * (entries flagged in recvIsGhost) with the owner's values */
__global__ void __launch_bounds__(256) p2p_pull_assign_nodal_kernel(
  double* base, int64_t stride, int nc, const int64_t* __restrict__ recvIdx,
  const unsigned char* __restrict__ recvIsGhost, int64_t n, const P2pDev pp)
{
  if (!p2p_wait(pp))
    return;
  const double* win = pp.myWindow + pp.winOff;
  const int64_t total = n * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = t / nc;
    if (!recvIsGhost[g])
      continue;
    const int c = (int)(t - g * nc);
    base[(int64_t)c * stride + recvIdx[g]] = __ldcg(win + t);
  }
}

/* linear-system push: contiguous tail segments -> the owners' windows */
__global__ void __launch_bounds__(256) p2p_push_segments_kernel(
  const double* const* __restrict__ segSrc, const int64_t* __restrict__ segStart,
  const int64_t* __restrict__ segDst, const int32_t* __restrict__ segPeer,
  int nSeg, const P2pDev pp)
{
  const int64_t total = segStart[nSeg];
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int sg = 0;
    while (sg + 1 < nSeg && segStart[sg + 1] <= t)
      ++sg;
    const int64_t k = t - segStart[sg];
    pp.peerWindow[segPeer[sg]][pp.winOff + segDst[sg] + k] = segSrc[sg][k];
  }
  p2p_signal(pp);
}

/* linear-system pull: accumulate_multi out of the own window, after the wait */
__global__ void __launch_bounds__(256) p2p_pull_accumulate_kernel(
  int64_t bufOff, int64_t entStride, int64_t compStride, int nc,
  const int64_t* __restrict__ dstIdx, const int64_t* __restrict__ ptr,
  const int64_t* __restrict__ pos, int64_t nDst, double* dst,
  int64_t dstCompStride, const P2pDev pp, int wait, int signal)
{
  if (wait) {
    if (!p2p_signal_then_wait(pp, signal))
      return;
  } else if (*(volatile unsigned*)(pp.sync + 1) != 0u) {
    return; /* the launch that waited for this epoch timed out */
  }
  const double* buf = pp.myWindow + pp.winOff + bufOff;
  const int64_t total = nDst * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = t / nc;
    const int c = (int)(t - u * nc);
    double* d = dst + (int64_t)c * dstCompStride + dstIdx[u];
    double v = *d;
    for (int64_t q = ptr[u]; q < ptr[u + 1]; ++q)
      v += __ldcg(buf + pos[q] * entStride + c * compStride);
    *d = v;
  }
}

/* both accumulate plans of a linear system in one launch: the matrix values
 * (window offset 0) and the rhs columns (window offset rhsOff, column c at
 * + c * rhsColStride) -- same order of additions as two
 * p2p_pull_accumulate_kernel launches */
__global__ void __launch_bounds__(256) p2p_pull_accumulate2_kernel(
  const int64_t* __restrict__ valDst, const int64_t* __restrict__ valPtr,
  const int64_t* __restrict__ valPos, int64_t nVal, double* values,
  int64_t rhsOff, int64_t rhsColStride, int nR,
  const int64_t* __restrict__ rhsDst, const int64_t* __restrict__ rhsPtr,
  const int64_t* __restrict__ rhsPos, int64_t nRhs, double* rhs,
  int64_t rhsStride, const P2pDev pp, int signal)
{
  if (!p2p_signal_then_wait(pp, signal))
    return;
  const double* buf = pp.myWindow + pp.winOff;
  const int64_t total = nVal + nRhs * nR;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    if (t < nVal) {
      double* d = values + valDst[t];
      double v = *d;
      for (int64_t q = valPtr[t]; q < valPtr[t + 1]; ++q)
        v += __ldcg(buf + valPos[q]);
      *d = v;
    } else {
      const int64_t t2 = t - nVal;
      const int64_t u = t2 / nR;
      const int c = (int)(t2 - u * nR);
      double* d = rhs + (int64_t)c * rhsStride + rhsDst[u];
      double v = *d;
      for (int64_t q = rhsPtr[u]; q < rhsPtr[u + 1]; ++q)
        v += __ldcg(buf + rhsOff + rhsPos[q] + (int64_t)c * rhsColStride);
      *d = v;
    }
  }
}

template <class K>
cudaError_t
set_smem(K kernel, size_t bytes)
{
  /* static shared memory counts against the 48 KB default too */
  if (bytes > 40 * 1024)
    return cudaFuncSetAttribute(
      kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

/* node stage (nodeComps) + edge components + packed (L,R) records */
inline size_t
edge_kernel_smem(const MeshPlanDev& mp, int nodeComps, int edgeComps)
{
  const size_t estride = (size_t)((mp.maxTileEdges + 1) & ~1);
  return sizeof(double) * ((size_t)nodeComps * mp.maxStaged + edgeComps * estride) +
         4u * (size_t)((mp.maxTileEdges + 3) & ~3);
}

template <class P>
size_t
ls_tile_smem(const MeshPlanDev& mp, const LsPlanDev& lp)
{
  return LsSmem<P>(mp, lp).bytes();
}

/* SM count of the current device (persistent grids) */
inline int
sm_count()
{
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0)
      n = 148;
  }
  return n;
}

/* tuning switches (diagnostic; the defaults are the product path):
 *   NW_STREAM=1        use the persistent stream kernels (comparison variant)
 *   NW_STREAM_CTAS=n   resident stream CTAs per SM (1 or 2) */
inline int
env_int(const char* name, int dflt)
{
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
/* NW_DBG_SKIP: timing experiments (see MeshPlanDev::dbgSkip) */
template <class K>
inline MeshPlanDev
with_pf(const MeshPlanDev& mp, K, int, size_t)
{
  static const int dbgEnv = env_int("NW_DBG_SKIP", 0);
  MeshPlanDev m = mp;
  m.dbgSkip = dbgEnv;
  return m;
}

inline bool
stream_enabled()
{
  static const int on = env_int("NW_STREAM", 0);
  return on != 0;
}

template <class P, int ND>
cudaError_t launch_ls_stream(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, const typename P::Opts& o, cudaStream_t s,
  bool* launched);

template <class P, int ND>
cudaError_t
launch_ls_tile(
  const MeshPlanDev& mp,
  const LsPlanDev& lpIn,
  const NodeComps& nc,
  const EdgeComps& ec,
  const typename P::Opts& o,
  cudaStream_t s,
  double* diagOut = nullptr)
{
  LsPlanDev lp = lpIn;
  lp.diagOut = diagOut;
  bool launched = false;
  cudaError_t e = cudaSuccess;
  /* the stream variant has no extract_diagonal pass; neither experimental
   * variant knows the fused push */
  if (!diagOut && !lp.push.enabled)
    e = launch_ls_stream<P, ND>(mp, lp, nc, ec, o, s, &launched);
  if (e != cudaSuccess || launched)
    return e;
  /* NW_PIPE (experiment): the warp-specialised persistent
   * kernel, when its slots fit one CTA's shared memory (tiles of <= ~136 nodes
   * for momentum on a hex mesh) and no extract_diagonal pass is asked for.
   * Bit-identical to the tile kernel, measured slower: not the default. */
  const int pipeEnv = env_int("NW_PIPE", 0); /* read per call: tests toggle it */
  if (pipeEnv && !diagOut && !lp.push.enabled && mp.nTiles > 0) {
    const size_t pb = PipeSmem<P>(mp, lp).bytes();
    if (pb + 2048 <= 227 * 1024) {
      const int grid = std::min(mp.nTiles, sm_count());
      /* NW_PIPE=<stagers><reducers>: 32 (also for any other non-zero value), 43 */
#define NW_PIPE_LAUNCH(NR, NS)                                                   \
  do {                                                                          \
    e = set_smem(ls_pipe_kernel<P, ND, NR, NS>, pb);                            \
    if (e != cudaSuccess)                                                       \
      return e;                                                                 \
    ls_pipe_kernel<P, ND, NR, NS><<<grid, kPipeThreads, pb, s>>>(mp, lp, nc, ec, o); \
  } while (0)
      if (pipeEnv == 43)
        NW_PIPE_LAUNCH(3, 4);
      else
        NW_PIPE_LAUNCH(2, 3);
#undef NW_PIPE_LAUNCH
      return cudaGetLastError();
    }
  }
  const size_t bytes = ls_tile_smem<P>(mp, lp);
  if (bytes > 227 * 1024)
    return cudaErrorInvalidConfiguration;
  e = set_smem(ls_tile_kernel<P, ND>, bytes);
  if (e != cudaSuccess)
    return e;
  if (mp.nTiles == 0)
    return cudaSuccess;
  ls_tile_kernel<P, ND><<<mp.nTiles, kTileThreads, bytes, s>>>(
    with_pf(mp, ls_tile_kernel<P, ND>, kTileThreads, bytes), lp, nc, ec, o);
  return cudaGetLastError();
}

/* the tile kernel only (no comparison variants): the VOF policies */
template <class P, int ND>
cudaError_t
launch_ls_tile_plain(
  const MeshPlanDev& mp,
  const LsPlanDev& lpIn,
  const NodeComps& nc,
  const EdgeComps& ec,
  const typename P::Opts& o,
  cudaStream_t s,
  double* diagOut)
{
  LsPlanDev lp = lpIn;
  lp.diagOut = diagOut;
  const size_t bytes = ls_tile_smem<P>(mp, lp);
  if (bytes > 227 * 1024)
    return cudaErrorInvalidConfiguration;
  cudaError_t e = set_smem(ls_tile_kernel<P, ND>, bytes);
  if (e != cudaSuccess)
    return e;
  if (mp.nTiles == 0)
    return cudaSuccess;
  ls_tile_kernel<P, ND><<<mp.nTiles, kTileThreads, bytes, s>>>(
    with_pf(mp, ls_tile_kernel<P, ND>, kTileThreads, bytes), lp, nc, ec, o);
  return cudaGetLastError();
}

template <class P, int ND>
cudaError_t
launch_ls_stream(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const NodeComps& nc,
  const EdgeComps& ec,
  const typename P::Opts& o,
  cudaStream_t s,
  bool* launched)
{
  *launched = false;
  if (!stream_enabled() || mp.nTiles == 0)
    return cudaSuccess;
  const size_t bytes = LsStreamSmem<P>(mp, lp).bytes();
  static const int ctasEnv = env_int("NW_STREAM_CTAS", 0);
  int ctas = ctasEnv > 0 ? ctasEnv : P::kStreamCtas;
  /* resident CTAs the shared memory allows (1 KB reserved per CTA) */
  while (ctas > 1 && (size_t)ctas * (bytes + 1024 + sizeof(StreamShared)) > 228 * 1024)
    --ctas;
  if (bytes + 1024 + sizeof(StreamShared) > 227 * 1024)
    return cudaSuccess; /* tile too large for the stream layout: tile kernel */
  const int grid = std::min(mp.nTiles, sm_count() * ctas);
  cudaError_t e;
  if (ctas >= 3) {
    if ((e = set_smem(ls_stream_kernel<P, ND, 3>, bytes)) != cudaSuccess)
      return e;
    ls_stream_kernel<P, ND, 3>
      <<<grid, kStreamThreads, bytes, s>>>(mp, lp, nc, ec, o);
  } else if (ctas == 2) {
    if ((e = set_smem(ls_stream_kernel<P, ND, 2>, bytes)) != cudaSuccess)
      return e;
    ls_stream_kernel<P, ND, 2>
      <<<grid, kStreamThreads, bytes, s>>>(mp, lp, nc, ec, o);
  } else {
    if ((e = set_smem(ls_stream_kernel<P, ND, 1>, bytes)) != cudaSuccess)
      return e;
    ls_stream_kernel<P, ND, 1>
      <<<grid, kStreamThreads, bytes, s>>>(mp, lp, nc, ec, o);
  }
  *launched = true;
  return cudaGetLastError();
}

template <class P, int ND>
cudaError_t
launch_ls_atomic(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const AtomicMapDev& am,
  const NodeComps& nc,
  const EdgeComps& ec,
  const typename P::Opts& o,
  double* diagOut,
  cudaStream_t s)
{
  ls_atomic_kernel<P, ND>
    <<<mp.nTiles, kTileThreads, 0, s>>>(mp, lp, am, nc, ec, o, diagOut);
  return cudaGetLastError();
}

inline int
blocks_for(int64_t n, int bs)
{
  return (int)((n + bs - 1) / bs);
}

} // namespace

/* ------------------------------------------------------------------ */
/*  launchers                                                          */
/* ------------------------------------------------------------------ */

cudaError_t
phase_times_read(unsigned long long* out, bool reset)
{
#ifdef NW_PHASE_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(out, g_phase, sizeof(g_phase));
  if (e != cudaSuccess || !reset)
    return e;
  static const unsigned long long zero[kPhaseKernels][kPhaseSlots] = {};
  return cudaMemcpyToSymbol(g_phase, zero, sizeof(zero));
#else
  (void)reset;
  for (int i = 0; i < kPhaseKernels * kPhaseSlots; ++i)
    out[i] = 0;
  return cudaSuccess;
#endif
}

cudaError_t
launch_mdot_tile(
  const MeshPlanDev& mp,
  const NodeComps& nc,
  const EdgeComps& ec,
  double* mdotOut,
  nw_mdot_opts o,
  cudaStream_t s)
{
  cudaError_t e;
  if (mp.ndim == 3) {
    const size_t bytes = edge_kernel_smem(mp, ContinuityP<3>::NC, 3);
    if ((e = set_smem(mdot_tile_kernel<3>, bytes)) != cudaSuccess)
      return e;
    mdot_tile_kernel<3><<<mp.nTiles, kTileThreads, bytes, s>>>(
      with_pf(mp, mdot_tile_kernel<3>, kTileThreads, bytes), nc, ec, mdotOut, o);
  } else {
    const size_t bytes = edge_kernel_smem(mp, ContinuityP<2>::NC, 2);
    if ((e = set_smem(mdot_tile_kernel<2>, bytes)) != cudaSuccess)
      return e;
    mdot_tile_kernel<2><<<mp.nTiles, kTileThreads, bytes, s>>>(
      with_pf(mp, mdot_tile_kernel<2>, kTileThreads, bytes), nc, ec, mdotOut, o);
  }
  return cudaGetLastError();
}

cudaError_t
launch_mdot_ext(
  const MeshPlanDev& mp, const NodeComps& nc, const EdgeComps& ec,
  const ContExtraDev& ex, double* mdotOut, nw_mdot_opts o, cudaStream_t s)
{
  if (mp.ndim == 3)
    mdot_ext_kernel<3><<<mp.nTiles, kTileThreads, 0, s>>>(mp, nc, ec, ex, mdotOut, o);
  else
    mdot_ext_kernel<2><<<mp.nTiles, kTileThreads, 0, s>>>(mp, nc, ec, ex, mdotOut, o);
  return cudaGetLastError();
}

cudaError_t
launch_continuity_ext_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, const ContExtraDev& ex,
  nw_continuity_opts o, cudaStream_t s)
{
  if (mp.ndim == 3)
    continuity_ext_atomic_kernel<3>
      <<<mp.nTiles, kTileThreads, 0, s>>>(mp, lp, am, nc, ec, ex, o);
  else
    continuity_ext_atomic_kernel<2>
      <<<mp.nTiles, kTileThreads, 0, s>>>(mp, lp, am, nc, ec, ex, o);
  return cudaGetLastError();
}

cudaError_t
launch_peclet_tile(
  const MeshPlanDev& mp,
  const NodeComps& nc,
  double* pecfacOut,
  nw_peclet_opts o,
  cudaStream_t s)
{
  cudaError_t e;
  if (mp.ndim == 3) {
    const size_t bytes = edge_kernel_smem(mp, 8, 0);
    if ((e = set_smem(peclet_tile_kernel<3>, bytes)) != cudaSuccess)
      return e;
    peclet_tile_kernel<3><<<mp.nTiles, kTileThreads, bytes, s>>>(
      with_pf(mp, peclet_tile_kernel<3>, kTileThreads, bytes), nc, pecfacOut, o);
  } else {
    const size_t bytes = edge_kernel_smem(mp, 6, 0);
    if ((e = set_smem(peclet_tile_kernel<2>, bytes)) != cudaSuccess)
      return e;
    peclet_tile_kernel<2><<<mp.nTiles, kTileThreads, bytes, s>>>(
      with_pf(mp, peclet_tile_kernel<2>, kTileThreads, bytes), nc, pecfacOut, o);
  }
  return cudaGetLastError();
}

namespace {
template <int D1, int ND>
cudaError_t
launch_grad_tile_t(
  const MeshPlanDev& mp,
  const NodeComps& phi,
  const double* dualVol,
  const EdgeComps& ec,
  double* const* gradOut,
  cudaStream_t s,
  const NodePushDev* push)
{
  GradOut go;
  for (int k = 0; k < D1 * ND; ++k)
    go.c[k] = gradOut[k];
  NodePushDev pd;
  if (push) {
    if (push->nc != D1 * ND)
      return cudaErrorInvalidValue;
    pd = *push;
  }
  const size_t bytes = edge_kernel_smem(mp, D1, ND) + 4u * (size_t)mp.maxTileEllNode;
  cudaError_t e = set_smem(grad_tile_kernel<D1, ND>, bytes);
  if (e != cudaSuccess)
    return e;
  if (mp.nTiles == 0)
    return cudaSuccess;
  grad_tile_kernel<D1, ND><<<mp.nTiles, kTileThreads, bytes, s>>>(
    with_pf(mp, grad_tile_kernel<D1, ND>, kTileThreads, bytes), phi, dualVol, ec,
    go, pd);
  return cudaGetLastError();
}
template <int D1, int ND>
cudaError_t
launch_grad_atomic_t(
  const MeshPlanDev& mp,
  const NodeComps& phi,
  const double* dualVol,
  const EdgeComps& ec,
  double* const* gradOut,
  cudaStream_t s)
{
  GradOut go;
  for (int k = 0; k < D1 * ND; ++k)
    go.c[k] = gradOut[k];
  grad_atomic_kernel<D1, ND>
    <<<mp.nTiles, kTileThreads, 0, s>>>(mp, phi, dualVol, ec, go);
  return cudaGetLastError();
}
} // namespace

cudaError_t
launch_grad_tile(
  const MeshPlanDev& mp,
  int dim1,
  const NodeComps& phi,
  const double* dualVol,
  const EdgeComps& ec,
  double* const* gradOut,
  cudaStream_t s,
  const NodePushDev* push)
{
  /* dim1 == 2 on a 3-D mesh: two scalar fields at once (nw_nodal_grad_edge_pair) */
  if (mp.ndim == 3)
    return dim1 == 1   ? launch_grad_tile_t<1, 3>(mp, phi, dualVol, ec, gradOut, s, push)
           : dim1 == 2 ? launch_grad_tile_t<2, 3>(mp, phi, dualVol, ec, gradOut, s, push)
                       : launch_grad_tile_t<3, 3>(mp, phi, dualVol, ec, gradOut, s, push);
  return dim1 == 1 ? launch_grad_tile_t<1, 2>(mp, phi, dualVol, ec, gradOut, s, push)
                   : launch_grad_tile_t<2, 2>(mp, phi, dualVol, ec, gradOut, s, push);
}

cudaError_t
launch_grad_atomic(
  const MeshPlanDev& mp,
  int dim1,
  const NodeComps& phi,
  const double* dualVol,
  const EdgeComps& ec,
  double* const* gradOut,
  cudaStream_t s)
{
  if (mp.ndim == 3)
    return dim1 == 1
             ? launch_grad_atomic_t<1, 3>(mp, phi, dualVol, ec, gradOut, s)
             : launch_grad_atomic_t<3, 3>(mp, phi, dualVol, ec, gradOut, s);
  return dim1 == 1
           ? launch_grad_atomic_t<1, 2>(mp, phi, dualVol, ec, gradOut, s)
           : launch_grad_atomic_t<2, 2>(mp, phi, dualVol, ec, gradOut, s);
}

cudaError_t
launch_continuity_tile(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_continuity_opts o,
  cudaStream_t s)
{
  return mp.ndim == 3
           ? launch_ls_tile<ContinuityP<3>, 3>(mp, lp, nc, ec, o, s)
           : launch_ls_tile<ContinuityP<2>, 2>(mp, lp, nc, ec, o, s);
}

cudaError_t
launch_wall_dist_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, cudaStream_t s)
{
  const nw_wall_dist_opts_ o{0};
  return mp.ndim == 3 ? launch_ls_tile<WallDistP<3>, 3>(mp, lp, nc, ec, o, s)
                      : launch_ls_tile<WallDistP<2>, 2>(mp, lp, nc, ec, o, s);
}

cudaError_t
launch_wall_dist_atomic(
  const MeshPlanDev& mp, const LsPlanDev& lp, const AtomicMapDev& am,
  const NodeComps& nc, const EdgeComps& ec, cudaStream_t s)
{
  const nw_wall_dist_opts_ o{0};
  return mp.ndim == 3
           ? launch_ls_atomic<WallDistP<3>, 3>(mp, lp, am, nc, ec, o, nullptr, s)
           : launch_ls_atomic<WallDistP<2>, 2>(mp, lp, am, nc, ec, o, nullptr, s);
}

cudaError_t
launch_scalar_tile(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_scalar_opts o,
  cudaStream_t s)
{
  return mp.ndim == 3 ? launch_ls_tile<ScalarP<3>, 3>(mp, lp, nc, ec, o, s)
                      : launch_ls_tile<ScalarP<2>, 2>(mp, lp, nc, ec, o, s);
}

/* does the tile kernel of a policy fit one CTA's shared memory on this mesh /
 * graph?  (0 continuity, 1 scalar, 2 momentum UVW, 7 wall distance; the
 * kPhaseId of the policies) */
bool
ls_tile_fits(const MeshPlanDev& mp, const LsPlanDev& lp, int policy)
{
  size_t b = 0;
  const bool d3 = mp.ndim == 3;
  switch (policy) {
  case 0:
    b = d3 ? ls_tile_smem<ContinuityP<3>>(mp, lp) : ls_tile_smem<ContinuityP<2>>(mp, lp);
    break;
  case 1:
    b = d3 ? ls_tile_smem<ScalarP<3>>(mp, lp) : ls_tile_smem<ScalarP<2>>(mp, lp);
    break;
  case 2:
    b = d3 ? ls_tile_smem<MomentumUvwP<3>>(mp, lp) : ls_tile_smem<MomentumUvwP<2>>(mp, lp);
    break;
  case 3:
    b = d3 ? ls_tile_smem<MomentumMonoP<3>>(mp, lp) : ls_tile_smem<MomentumMonoP<2>>(mp, lp);
    break;
  default:
    b = d3 ? ls_tile_smem<WallDistP<3>>(mp, lp) : ls_tile_smem<WallDistP<2>>(mp, lp);
  }
  return b <= 227 * 1024;
}

cudaError_t
launch_scalar_pair_tile(
  const MeshPlanDev& mp, const LsPlanDev& lpA, double* valuesB, double* rhsB,
  const NodeComps& nc, const EdgeComps& ec, nw_scalar_opts oA,
  nw_scalar_opts oB, bool* launched, cudaStream_t s)
{
  *launched = false;
  cudaError_t e;
  if (mp.ndim == 3) {
    const size_t bytes = PairSmem<3>(mp, lpA).bytes();
    if (bytes > 226 * 1024)
      return cudaSuccess; /* the caller assembles the systems one by one */
    if ((e = set_smem(scalar_pair_tile_kernel<3>, bytes)) != cudaSuccess)
      return e;
    scalar_pair_tile_kernel<3><<<mp.nTiles, kPairThreads, bytes, s>>>(
      mp, lpA, valuesB, rhsB, nc, ec, oA, oB);
  } else {
    const size_t bytes = PairSmem<2>(mp, lpA).bytes();
    if (bytes > 226 * 1024)
      return cudaSuccess;
    if ((e = set_smem(scalar_pair_tile_kernel<2>, bytes)) != cudaSuccess)
      return e;
    scalar_pair_tile_kernel<2><<<mp.nTiles, kPairThreads, bytes, s>>>(
      mp, lpA, valuesB, rhsB, nc, ec, oA, oB);
  }
  *launched = true;
  return cudaGetLastError();
}

cudaError_t
launch_momentum_uvw_tile(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_momentum_opts o,
  double* diagOut,
  cudaStream_t s)
{
  if (o.has_vof)
    return mp.ndim == 3
             ? launch_ls_tile_plain<MomentumUvwVofP<3>, 3>(mp, lp, nc, ec, o, s, diagOut)
             : launch_ls_tile_plain<MomentumUvwVofP<2>, 2>(mp, lp, nc, ec, o, s, diagOut);
  return mp.ndim == 3
           ? launch_ls_tile<MomentumUvwP<3>, 3>(mp, lp, nc, ec, o, s, diagOut)
           : launch_ls_tile<MomentumUvwP<2>, 2>(mp, lp, nc, ec, o, s, diagOut);
}

namespace {
constexpr int kMonoThreads = 512; /* one CTA per SM (145 KB of shared memory) */
template <int ND, class P = MomentumMonoP<ND>>
cudaError_t
launch_momentum_mono_tile_t(
  const MeshPlanDev& mp, const LsPlanDev& lpIn, const NodeComps& nc,
  const EdgeComps& ec, const nw_momentum_opts& o, double* diagOut, cudaStream_t s)
{
  LsPlanDev lp = lpIn;
  lp.diagOut = diagOut;
  lp.push = LsPushDev(); /* shared rows of a monolithic system go through load_complete */
  const size_t bytes = ls_tile_smem<P>(mp, lp);
  if (bytes > 227 * 1024)
    return cudaErrorInvalidConfiguration;
  cudaError_t e = set_smem(ls_tile_kernel<P, ND, 1, kMonoThreads>, bytes);
  if (e != cudaSuccess)
    return e;
  if (mp.nTiles == 0)
    return cudaSuccess;
  ls_tile_kernel<P, ND, 1, kMonoThreads><<<mp.nTiles, kMonoThreads, bytes, s>>>(
    with_pf(mp, ls_tile_kernel<P, ND, 1, kMonoThreads>, kMonoThreads, bytes), lp,
    nc, ec, o);
  return cudaGetLastError();
}
} // namespace

cudaError_t
launch_momentum_mono_tile(
  const MeshPlanDev& mp, const LsPlanDev& lp, const NodeComps& nc,
  const EdgeComps& ec, nw_momentum_opts o, double* diagOut, cudaStream_t s)
{
  if (o.has_vof)
    return mp.ndim == 3
             ? launch_momentum_mono_tile_t<3, MomentumMonoVofP<3>>(mp, lp, nc, ec, o, diagOut, s)
             : launch_momentum_mono_tile_t<2, MomentumMonoVofP<2>>(mp, lp, nc, ec, o, diagOut, s);
  return mp.ndim == 3
           ? launch_momentum_mono_tile_t<3>(mp, lp, nc, ec, o, diagOut, s)
           : launch_momentum_mono_tile_t<2>(mp, lp, nc, ec, o, diagOut, s);
}

cudaError_t
launch_continuity_atomic(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const AtomicMapDev& am,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_continuity_opts o,
  cudaStream_t s)
{
  return mp.ndim == 3
           ? launch_ls_atomic<ContinuityP<3>, 3>(mp, lp, am, nc, ec, o, nullptr, s)
           : launch_ls_atomic<ContinuityP<2>, 2>(mp, lp, am, nc, ec, o, nullptr, s);
}

cudaError_t
launch_scalar_atomic(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const AtomicMapDev& am,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_scalar_opts o,
  cudaStream_t s)
{
  return mp.ndim == 3
           ? launch_ls_atomic<ScalarP<3>, 3>(mp, lp, am, nc, ec, o, nullptr, s)
           : launch_ls_atomic<ScalarP<2>, 2>(mp, lp, am, nc, ec, o, nullptr, s);
}

cudaError_t
launch_momentum_uvw_atomic(
  const MeshPlanDev& mp,
  const LsPlanDev& lp,
  const AtomicMapDev& am,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_momentum_opts o,
  double* diagOut,
  cudaStream_t s)
{
  if (o.has_vof)
    return mp.ndim == 3
             ? launch_ls_atomic<MomentumUvwVofP<3>, 3>(mp, lp, am, nc, ec, o, diagOut, s)
             : launch_ls_atomic<MomentumUvwVofP<2>, 2>(mp, lp, am, nc, ec, o, diagOut, s);
  return mp.ndim == 3
           ? launch_ls_atomic<MomentumUvwP<3>, 3>(mp, lp, am, nc, ec, o, diagOut, s)
           : launch_ls_atomic<MomentumUvwP<2>, 2>(mp, lp, am, nc, ec, o, diagOut, s);
}

cudaError_t
launch_momentum_mono_atomic(
  const MeshPlanDev& mp,
  const int32_t* slots,
  const int32_t* rhsRows,
  double* values,
  double* rhs,
  const NodeComps& nc,
  const EdgeComps& ec,
  nw_momentum_opts o,
  double* diagOut,
  cudaStream_t s)
{
  if (o.has_vof) {
    if (mp.ndim == 3)
      momentum_mono_atomic_kernel<3, true><<<mp.nTiles, kTileThreads, 0, s>>>(
        mp, slots, rhsRows, values, rhs, nc, ec, o, diagOut);
    else
      momentum_mono_atomic_kernel<2, true><<<mp.nTiles, kTileThreads, 0, s>>>(
        mp, slots, rhsRows, values, rhs, nc, ec, o, diagOut);
    return cudaGetLastError();
  }
  if (mp.ndim == 3)
    momentum_mono_atomic_kernel<3><<<mp.nTiles, kTileThreads, 0, s>>>(
      mp, slots, rhsRows, values, rhs, nc, ec, o, diagOut);
  else
    momentum_mono_atomic_kernel<2><<<mp.nTiles, kTileThreads, 0, s>>>(
      mp, slots, rhsRows, values, rhs, nc, ec, o, diagOut);
  return cudaGetLastError();
}

cudaError_t
launch_node_gather(
  const double* srcAos,
  int ncomp,
  const int32_t* nodeOfSlot,
  int64_t nSlots,
  double* dstSoa,
  cudaStream_t s)
{
  node_gather_kernel<<<blocks_for(nSlots, 256), 256, 0, s>>>(
    srcAos, ncomp, nodeOfSlot, nSlots, dstSoa);
  return cudaGetLastError();
}

cudaError_t
launch_node_scatter(
  const double* srcSoa,
  int ncomp,
  const int32_t* nodeOfSlot,
  int64_t nSlots,
  double* dstAos,
  cudaStream_t s)
{
  node_scatter_kernel<<<blocks_for(nSlots, 256), 256, 0, s>>>(
    srcSoa, ncomp, nodeOfSlot, nSlots, dstAos);
  return cudaGetLastError();
}

cudaError_t
launch_edge_gather(
  const double* srcAos,
  int ncomp,
  const int32_t* tileEdgeSrc,
  int64_t nSlots,
  double* dstSoa,
  cudaStream_t s)
{
  /* same access pattern as the node gather: slot -> source entity */
  node_gather_kernel<<<blocks_for(nSlots, 256), 256, 0, s>>>(
    srcAos, ncomp, tileEdgeSrc, nSlots, dstSoa);
  return cudaGetLastError();
}

cudaError_t
launch_edge_scatter(
  const double* srcSoa,
  int ncomp,
  const int32_t* primarySlotOfEdge,
  int64_t nEdges,
  int64_t slotStride,
  double* dstAos,
  cudaStream_t s)
{
  if (nEdges == 0)
    return cudaSuccess;
  edge_scatter_kernel<<<blocks_for(nEdges, 256), 256, 0, s>>>(
    srcSoa, ncomp, primarySlotOfEdge, nEdges, slotStride, dstAos);
  return cudaGetLastError();
}

/* out[i] = a[i] + b[i]: the VOF momentum kernels' edge mass flow,
 * massFlowRate + has_vof * massVofBalancedFlowRate with has_vof == 1.0
 * (src/edge_kernels/MomentumEdgeSolverAlg.C:124-125; 1.0 * b is b exactly) */
__global__ void
edge_sum_kernel(
  const double* __restrict__ a, const double* __restrict__ b, int64_t n,
  double* __restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = a[i] + b[i];
}

cudaError_t
launch_edge_sum(
  const double* a, const double* b, int64_t n, double* out, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  edge_sum_kernel<<<blocks_for(n, 256), 256, 0, s>>>(a, b, n, out);
  return cudaGetLastError();
}

cudaError_t
launch_fill(double* p, int64_t n, double v, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  fill_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}

cudaError_t
launch_row_init(
  const int32_t* rows,
  int nRows,
  const int64_t* rowPtr,
  const uint8_t* isPeriodic,
  double* values,
  double* rhs,
  int64_t rhsStride,
  int nRhs,
  cudaStream_t s)
{
  if (nRows == 0)
    return cudaSuccess;
  row_init_kernel<<<blocks_for(nRows, 128), 128, 0, s>>>(
    rows, nRows, rowPtr, isPeriodic, values, rhs, rhsStride, nRhs);
  return cudaGetLastError();
}

cudaError_t
launch_geometry_hex8(
  int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s)
{
  if (nElems == 0)
    return cudaSuccess;
  geometry_hex8_kernel<<<blocks_for(nElems, 128), 128, 0, s>>>(
    nElems, elemSlots, elemEdges, owned, x, xStride, dualVol, area, areaStride);
  return cudaGetLastError();
}

cudaError_t
launch_geometry_quad4(
  int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s)
{
  if (nElems == 0)
    return cudaSuccess;
  geometry_quad4_kernel<<<blocks_for(nElems, 128), 128, 0, s>>>(
    nElems, elemSlots, elemEdges, owned, x, xStride, dualVol, area, areaStride);
  return cudaGetLastError();
}

cudaError_t
launch_edge_mirror(
  const int32_t* primarySlot, const int32_t* secondSlot, int64_t nEdges,
  int ncomp, int64_t stride, double* f, cudaStream_t s)
{
  if (nEdges == 0)
    return cudaSuccess;
  edge_mirror_kernel<<<blocks_for(nEdges, 256), 256, 0, s>>>(
    primarySlot, secondSlot, nEdges, ncomp, stride, f);
  return cudaGetLastError();
}

cudaError_t
launch_mass_bdf_node(
  int kind, int ndim, const int64_t* rows, int64_t nRows,
  const MassBdfFields& f, double dt, double gamma1, double gamma2,
  double gamma3, double* values, double* rhs, int64_t rhsStride, cudaStream_t s)
{
  if (nRows == 0)
    return cudaSuccess;
  const int nb = blocks_for(nRows, 256);
  if (kind == NW_MASS_SCALAR)
    mass_bdf_node_kernel<NW_MASS_SCALAR><<<nb, 256, 0, s>>>(
      ndim, rows, nRows, f, dt, gamma1, gamma2, gamma3, values, rhs, rhsStride);
  else if (kind == NW_MASS_MOMENTUM)
    mass_bdf_node_kernel<NW_MASS_MOMENTUM><<<nb, 256, 0, s>>>(
      ndim, rows, nRows, f, dt, gamma1, gamma2, gamma3, values, rhs, rhsStride);
  else
    mass_bdf_node_kernel<NW_MASS_CONTINUITY><<<nb, 256, 0, s>>>(
      ndim, rows, nRows, f, dt, gamma1, gamma2, gamma3, values, rhs, rhsStride);
  return cudaGetLastError();
}

cudaError_t
launch_wall_dist_node(
  const int64_t* rows, int64_t nRows, const double* dualVol, double* rhs,
  cudaStream_t s)
{
  if (nRows == 0)
    return cudaSuccess;
  wall_dist_node_kernel<<<blocks_for(nRows, 256), 256, 0, s>>>(
    rows, nRows, dualVol, rhs);
  return cudaGetLastError();
}

cudaError_t
launch_reset_rows(
  const int64_t* rows, int64_t nRows, double diagValue, double rhsResidual,
  double* values, double* rhs, int64_t rhsStride, int nRhs, cudaStream_t s)
{
  if (nRows == 0)
    return cudaSuccess;
  reset_rows_kernel<<<blocks_for(nRows, 128), 128, 0, s>>>(
    rows, nRows, diagValue, rhsResidual, values, rhs, rhsStride, nRhs);
  return cudaGetLastError();
}

cudaError_t
launch_dirichlet_rows(
  const int64_t* rows, int64_t nRows, const double* solution, const double* bc,
  int64_t fieldStride, double* values, double* rhs, int64_t rhsStride,
  cudaStream_t s)
{
  if (nRows == 0)
    return cudaSuccess;
  dirichlet_rows_kernel<<<blocks_for(nRows, 128), 128, 0, s>>>(
    rows, nRows, solution, bc, fieldStride, values, rhs, rhsStride);
  return cudaGetLastError();
}

cudaError_t
launch_norm2(
  const double* rhs,
  int64_t n,
  int64_t stride,
  int nRhs,
  double* partial,
  int nPartial,
  double* out,
  cudaStream_t s)
{
  dim3 grid(nPartial, nRhs);
  norm2_partial_kernel<<<grid, 256, 0, s>>>(rhs, n, stride, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  norm2_final_kernel<<<nRhs, 256, 0, s>>>(partial, nPartial, out);
  return cudaGetLastError();
}

cudaError_t
launch_sum_into(
  int64_t nEnt,
  int npe,
  int numDof,
  const int32_t* entNodes,
  const int64_t* nodeHid,
  const double* lhs,
  const double* rhsIn,
  int64_t iLower,
  int64_t iUpper,
  int64_t nRowsOwned,
  int64_t nnzOwned,
  const int64_t* rowStartOwned,
  const int64_t* rowStartShared,
  const int64_t* rowIndicesShared,
  int64_t nRowsShared,
  const int64_t* cols,
  const int64_t* skipped,
  int64_t nSkipped,
  int uvwDim,
  double* values,
  double* rhs,
  int64_t rhsStride,
  cudaStream_t s)
{
  if (nEnt == 0)
    return cudaSuccess;
  sum_into_kernel<<<blocks_for(nEnt, 128), 128, 0, s>>>(
    nEnt, npe, numDof, entNodes, nodeHid, lhs, rhsIn, iLower, iUpper,
    nRowsOwned, nnzOwned, rowStartOwned, rowStartShared, rowIndicesShared,
    nRowsShared, cols, skipped, nSkipped, uvwDim, values, rhs, rhsStride);
  return cudaGetLastError();
}

cudaError_t
launch_pack(
  const double* src, const int64_t* idx, int64_t n, double* dst, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  pack_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, idx, n, dst);
  return cudaGetLastError();
}

cudaError_t
launch_scatter_assign(
  const double* src, const int64_t* idx, int64_t n, double* dst, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  scatter_assign_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, idx, n, dst);
  return cudaGetLastError();
}

cudaError_t
launch_pack_multi(
  const double* src, int64_t srcCompStride, int nc, const int64_t* idx,
  int64_t n, double* buf, int64_t entStride, int64_t compStride, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  pack_multi_kernel<<<blocks_for(n * nc, 256), 256, 0, s>>>(
    src, srcCompStride, nc, idx, n, buf, entStride, compStride);
  return cudaGetLastError();
}

cudaError_t
launch_scatter_multi(
  const double* buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* idx, int64_t n, double* dst, int64_t dstCompStride,
  cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  scatter_multi_kernel<<<blocks_for(n * nc, 256), 256, 0, s>>>(
    buf, entStride, compStride, nc, idx, n, dst, dstCompStride);
  return cudaGetLastError();
}

cudaError_t
launch_accumulate_multi(
  const double* buf, int64_t entStride, int64_t compStride, int nc,
  const int64_t* dstIdx, const int64_t* ptr, const int64_t* pos, int64_t nDst,
  double* dst, int64_t dstCompStride, cudaStream_t s)
{
  if (nDst == 0)
    return cudaSuccess;
  accumulate_multi_kernel<<<blocks_for(nDst * nc, 256), 256, 0, s>>>(
    buf, entStride, compStride, nc, dstIdx, ptr, pos, nDst, dst, dstCompStride);
  return cudaGetLastError();
}

namespace {
inline int
p2p_grid(int64_t elems)
{
  const int64_t b = (elems + 255) / 256;
  const int64_t cap = 2 * (int64_t)sm_count();
  return (int)std::max<int64_t>(1, std::min(b, cap));
}
/* pull kernels run on the communication stream beside the compute kernels of
 * the main stream and spin until the neighbour's epoch has arrived: keep them
 * small so that they hold few SM resources while they wait -- but not so
 * small that a large exchange (512^3 over 8 GPUs: 5 M values per system)
 * becomes a millisecond of serial work the next-but-one kernel has to wait
 * for: at most ~8 grid-stride iterations per thread */
inline int
p2p_pull_grid(int64_t elems, bool beside)
{
  if (!beside)
    return p2p_grid(elems);
  const int64_t b = (elems + 255) / 256;
  const int64_t want = std::max<int64_t>(32, (b + 7) / 8);
  return (int)std::max<int64_t>(
    1, std::min<int64_t>(std::min(b, want), 2 * (int64_t)sm_count()));
}
} // namespace

cudaError_t
launch_p2p_push_nodal(
  const double* base, int64_t stride, int nc, const int64_t* sendIdx,
  const int32_t* sendPeer, const int64_t* sendDst, int64_t n, const P2pDev& pp,
  cudaStream_t s)
{
  p2p_push_nodal_kernel<<<p2p_grid(n * nc), 256, 0, s>>>(
    base, stride, nc, sendIdx, sendPeer, sendDst, n, pp);
  return cudaGetLastError();
}

cudaError_t
launch_p2p_pull_nodal(
  const CompPtrs& comps, int nc, const int64_t* recvIdx, int64_t n,
  const P2pDev& pp, bool beside, cudaStream_t s, bool signal)
{
  p2p_pull_nodal_kernel<<<p2p_pull_grid(n * nc, beside), 256, 0, s>>>(
    comps, nc, recvIdx, n, pp, signal ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t
launch_p2p_pull_assign_nodal(
  double* base, int64_t stride, int nc, const int64_t* recvIdx,
  const unsigned char* recvIsGhost, int64_t n, const P2pDev& pp, bool beside,
  cudaStream_t s)
{
  p2p_pull_assign_nodal_kernel<<<p2p_pull_grid(n * nc, beside), 256, 0, s>>>(
    base, stride, nc, recvIdx, recvIsGhost, n, pp);
  return cudaGetLastError();
}

cudaError_t
launch_p2p_push_segments(
  const double* const* segSrc, const int64_t* segStart, const int64_t* segDst,
  const int32_t* segPeer, int nSeg, int64_t total, const P2pDev& pp,
  cudaStream_t s)
{
  p2p_push_segments_kernel<<<p2p_grid(total), 256, 0, s>>>(
    segSrc, segStart, segDst, segPeer, nSeg, pp);
  return cudaGetLastError();
}

cudaError_t
launch_p2p_pull_accumulate(
  int64_t bufOff, int64_t entStride, int64_t compStride, int nc,
  const int64_t* dstIdx, const int64_t* ptr, const int64_t* pos, int64_t nDst,
  double* dst, int64_t dstCompStride, const P2pDev& pp, bool wait, bool beside,
  cudaStream_t s, bool signal)
{
  p2p_pull_accumulate_kernel<<<p2p_pull_grid(nDst * nc, beside), 256, 0, s>>>(
    bufOff, entStride, compStride, nc, dstIdx, ptr, pos, nDst, dst,
    dstCompStride, pp, wait ? 1 : 0, signal ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t
launch_p2p_pull_accumulate2(
  const int64_t* valDst, const int64_t* valPtr, const int64_t* valPos,
  int64_t nVal, double* values, int64_t rhsOff, int64_t rhsColStride, int nR,
  const int64_t* rhsDst, const int64_t* rhsPtr, const int64_t* rhsPos,
  int64_t nRhs, double* rhs, int64_t rhsStride, const P2pDev& pp, bool beside,
  cudaStream_t s, bool signal)
{
  p2p_pull_accumulate2_kernel<<<p2p_pull_grid(nVal + nRhs * nR, beside), 256, 0, s>>>(
    valDst, valPtr, valPos, nVal, values, rhsOff, rhsColStride, nR, rhsDst,
    rhsPtr, rhsPos, nRhs, rhs, rhsStride, pp, signal ? 1 : 0);
  return cudaGetLastError();
}

namespace {
/* PeriodicManager add_slave_to_master + set_slave_to_master for one field:
 * thread per (group, component) */
__global__ void __launch_bounds__(256) periodic_update_kernel(
  double* base, int64_t stride, int nc, const int32_t* __restrict__ ptr,
  const int32_t* __restrict__ slots, int nGroups)
{
  const int64_t total = (int64_t)nGroups * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(t / nc);
    const int c = (int)(t - (int64_t)g * nc);
    double* f = base + (int64_t)c * stride;
    const int a = ptr[g], b = ptr[g + 1];
    double v = f[slots[a]];
    for (int q = a + 1; q < b; ++q)
      v += f[slots[q]];
    for (int q = a; q < b; ++q)
      f[slots[q]] = v;
  }
}
} // namespace

cudaError_t
launch_periodic_update(
  double* base, int64_t stride, int nc, const int32_t* ptr,
  const int32_t* slots, int nGroups, cudaStream_t s)
{
  if (nGroups == 0)
    return cudaSuccess;
  periodic_update_kernel<<<blocks_for((int64_t)nGroups * nc, 256), 256, 0, s>>>(
    base, stride, nc, ptr, slots, nGroups);
  return cudaGetLastError();
}

namespace {
/* LowMach::udiag_post_processing (src/LowMachEquationSystem.C:2783-2790):
 * thread per selected node (locally owned, not a periodic slave); arithmetic
 * in edge_physics.h (udiag_post_value) */
__global__ void __launch_bounds__(256) udiag_post_kernel(
  double* udiag, const double* __restrict__ rho,
  const double* __restrict__ dvol, const int32_t* __restrict__ slots, int64_t n,
  double projTimeScale, double alphaU)
{
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t k = slots[t];
    udiag[k] = udiag_post_value(udiag[k], rho[k], dvol[k], projTimeScale, alphaU);
  }
}

/* PeriodicManager::apply_constraints with setSlaves only
 * (src/LowMachEquationSystem.C:2802-2808): every slave takes its master's
 * value; thread per (group, component) */
__global__ void __launch_bounds__(256) periodic_set_kernel(
  double* base, int64_t stride, int nc, const int32_t* __restrict__ ptr,
  const int32_t* __restrict__ slots, int nGroups)
{
  const int64_t total = (int64_t)nGroups * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(t / nc);
    const int c = (int)(t - (int64_t)g * nc);
    double* f = base + (int64_t)c * stride;
    const int a = ptr[g], b = ptr[g + 1];
    const double v = f[slots[a]];
    for (int q = a + 1; q < b; ++q)
      f[slots[q]] = v;
  }
}
} // namespace

cudaError_t
launch_udiag_post(
  double* udiag, const double* rho, const double* dvol, const int32_t* slots,
  int64_t n, double projTimeScale, double alphaU, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  udiag_post_kernel<<<blocks_for(n, 256), 256, 0, s>>>(
    udiag, rho, dvol, slots, n, projTimeScale, alphaU);
  return cudaGetLastError();
}

cudaError_t
launch_periodic_set(
  double* base, int64_t stride, int nc, const int32_t* ptr,
  const int32_t* slots, int nGroups, cudaStream_t s)
{
  if (nGroups == 0)
    return cudaSuccess;
  periodic_set_kernel<<<blocks_for((int64_t)nGroups * nc, 256), 256, 0, s>>>(
    base, stride, nc, ptr, slots, nGroups);
  return cudaGetLastError();
}

cudaError_t
launch_unpack_add(
  const double* src, const int64_t* idx, int64_t n, double* dst, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  unpack_add_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, idx, n, dst);
  return cudaGetLastError();
}

} // namespace nw
