/*
 * stage_layouts.cu -- microbenchmark of the tile kernels' STAGING phase under
 * different nodal layouts (design aid for the layout change planned in
 * DESIGN.md section 6; not part of the product library).
 *
 * A synthetic schedule that looks like the momentum kernel at 128^3: tiles of
 * 192 own nodes (a contiguous slot range) + ~200 halo nodes taken as runs from
 * six neighbouring tiles, 18 fp64 components per node, 256-thread CTAs, two
 * CTAs per SM (shared memory padded to force that).  Every variant stages one
 * tile's node data into shared memory, touches it (a checksum, so nothing is
 * optimised away) and moves on; the kernel reports the average cycles per tile
 * spent in "issue own-range copies", "halo gather" and "wait".
 *
 *   V0  SoA [18][N]            : 18 bulk copies + 18 x 8-byte loads per halo node   (today)
 *   V1  per-field AoS 3|3|9|1|1|1 : 6 bulk copies + 18 x 8-byte loads, contiguous per field
 *   V2  one record [N][18]     : 1 bulk copy + 9 x 16-byte loads per halo node
 *   V3  SoA + 2-D tensor map   : 1 tensor copy (18 rows x 192 columns) + V0's halo gather
 *   V4  record [N][18], halo by one 144-byte bulk copy per halo node
 *
 * Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
 *              -o stage_layouts tools/microbench/stage_layouts.cu -lcuda
 * Run:    ./stage_layouts [n_nodes_millions=2.1] [reps=20]
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__,                   \
                   cudaGetErrorString(e_));                                    \
      std::exit(1);                                                            \
    }                                                                          \
  } while (0)

constexpr int kComps = 18;
constexpr int kOwn = 192;     /* own nodes per tile (multiple of 2) */
constexpr int kHaloMax = 224; /* halo entries per tile, upper bound */
constexpr int kThreads = 256;

struct TileHdr
{
  int32_t ownBegin; /* first own slot */
  int32_t nHalo;
  int64_t haloPtr; /* into the halo list */
};

__device__ __forceinline__ uint32_t
smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void
mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void
mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
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
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  } while (!done);
}
__device__ __forceinline__ void
tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
      smem_u32(dst)),
    "l"(src), "r"(bytes), "r"(smem_u32(bar))
    : "memory");
}
__device__ __forceinline__ void
tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
    "%3}], [%4];" ::"r"(smem_u32(dst)),
    "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
    : "memory");
}

/* cycles: [0] issue, [1] gather, [2] wait, [3] consume, [4] tiles */
__device__ unsigned long long g_cycles[8];

template <int V>
__global__ void __launch_bounds__(kThreads, 2) stage_kernel(
  const TileHdr* __restrict__ tiles, int nTiles, const int32_t* __restrict__ halo,
  const double* __restrict__ data, int64_t nNodes, const CUtensorMap* __restrict__ tmap,
  double* sink)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s = reinterpret_cast<double*>(smem_raw); /* [kComps][kOwn + kHaloMax] or records */
  __shared__ uint64_t bar;
  constexpr int kStage = kOwn + kHaloMax;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    mbar_init(&bar, 1);
  __syncthreads();
  uint32_t parity = 0;
  double acc = 0.0;
  long long cIssue = 0, cGather = 0, cWait = 0, cUse = 0;
  int done = 0;
  /* field layout of V1: component offsets of the six fields */
  const int fOff[7] = {0, 3, 6, 15, 16, 17, 18};
  for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
    const TileHdr h = tiles[t];
    long long t0 = clock64();
    if (warp == 0) {
      if (lane == 0) {
        if (V == 0) {
          mbar_expect_tx(&bar, kComps * kOwn * 8);
          for (int c = 0; c < kComps; ++c)
            tma_load_1d(s + c * kStage, data + (int64_t)c * nNodes + h.ownBegin, kOwn * 8, &bar);
        } else if (V == 1) {
          mbar_expect_tx(&bar, kComps * kOwn * 8);
          for (int f = 0; f < 6; ++f) {
            const int nc = fOff[f + 1] - fOff[f];
            tma_load_1d(
              s + fOff[f] * kStage, data + (int64_t)fOff[f] * nNodes + (int64_t)h.ownBegin * nc,
              kOwn * nc * 8, &bar);
          }
        } else if (V == 2 || V == 4) {
          mbar_expect_tx(
            &bar, kComps * kOwn * 8 + (V == 4 ? (uint32_t)h.nHalo * kComps * 8 : 0u));
          tma_load_1d(s, data + (int64_t)h.ownBegin * kComps, kOwn * kComps * 8, &bar);
        } else {
          mbar_expect_tx(&bar, kComps * kOwn * 8);
          tma_load_2d(s, tmap, h.ownBegin, 0, &bar);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    /* halo gather: warps 1..7 (V4: every lane issues bulk copies) */
    if (V == 0 || V == 3) {
      /* V3's tensor box is [18][kOwn] dense: halo rows go behind it */
      double* hs = V == 0 ? s : s + kComps * kOwn;
      const int hstride = V == 0 ? kStage : kHaloMax;
      const int hbase = V == 0 ? kOwn : 0;
      if (warp > 0)
        for (int k = tid - 32; k < h.nHalo; k += kThreads - 32) {
          const int64_t n = halo[h.haloPtr + k];
#pragma unroll
          for (int c = 0; c < kComps; ++c)
            hs[c * hstride + hbase + k] = __ldg(data + (int64_t)c * nNodes + n);
        }
    } else if (V == 1) {
      if (warp > 0)
        for (int k = tid - 32; k < h.nHalo; k += kThreads - 32) {
          const int64_t n = halo[h.haloPtr + k];
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            const int nc = fOff[f + 1] - fOff[f];
            const double* src = data + (int64_t)fOff[f] * nNodes + n * nc;
            double* dst = s + fOff[f] * kStage + (kOwn + k) * nc;
            for (int c = 0; c < nc; ++c)
              dst[c] = __ldg(src + c);
          }
        }
    } else if (V == 2) {
      /* 9 lanes per halo node, one 16-byte load each */
      if (warp > 0)
        for (int q = tid - 32; q < h.nHalo * 9; q += kThreads - 32) {
          const int k = q / 9, j = q - 9 * k;
          const int64_t n = halo[h.haloPtr + k];
          const double2 v = __ldg(reinterpret_cast<const double2*>(data + n * kComps) + j);
          reinterpret_cast<double2*>(s + (int64_t)(kOwn + k) * kComps)[j] = v;
        }
    } else {
      for (int k = tid; k < h.nHalo; k += kThreads) {
        const int64_t n = halo[h.haloPtr + k];
        tma_load_1d(s + (int64_t)(kOwn + k) * kComps, data + n * kComps, kComps * 8, &bar);
      }
    }
    long long t2 = clock64();
    mbar_wait(&bar, parity);
    parity ^= 1;
    __syncthreads();
    long long t3 = clock64();
    /* consume: every staged value once */
    const int total = kComps * (kOwn + h.nHalo);
    if (V == 0) {
      for (int c = 0; c < kComps; ++c)
        for (int i = tid; i < kOwn + h.nHalo; i += kThreads)
          acc += s[c * kStage + i];
    } else if (V == 3) {
      for (int i = tid; i < kComps * kOwn; i += kThreads)
        acc += s[i];
      for (int c = 0; c < kComps; ++c)
        for (int i = tid; i < h.nHalo; i += kThreads)
          acc += s[kComps * kOwn + c * kHaloMax + i];
    } else if (V == 1) {
      for (int f = 0; f < 6; ++f) {
        const int nc = fOff[f + 1] - fOff[f];
        for (int i = tid; i < nc * (kOwn + h.nHalo); i += kThreads)
          acc += s[fOff[f] * kStage + i];
      }
    } else {
      for (int i = tid; i < total; i += kThreads)
        acc += s[i];
    }
    __syncthreads();
    long long t4 = clock64();
    cIssue += t1 - t0;
    cGather += t2 - t1;
    cWait += t3 - t2;
    cUse += t4 - t3;
    ++done;
  }
  if (acc == 12345.678)
    sink[0] = acc;
  if (tid == 32) { /* a gathering warp's view */
    atomicAdd(&g_cycles[1], (unsigned long long)cGather);
    atomicAdd(&g_cycles[2], (unsigned long long)cWait);
    atomicAdd(&g_cycles[3], (unsigned long long)cUse);
    atomicAdd(&g_cycles[4], (unsigned long long)done);
  }
  if (tid == 0)
    atomicAdd(&g_cycles[0], (unsigned long long)cIssue);
}

typedef CUresult (*EncodeTiled)(
  CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
static void
run(
  const char* name, const TileHdr* dTiles, int nTiles, const int32_t* dHalo, const double* dData,
  int64_t nNodes, const CUtensorMap* dMap, double* dSink, int reps, int smCount, double usefulMB)
{
  const size_t smem = 100 * 1024; /* two CTAs per SM, as the momentum kernel */
  CK(cudaFuncSetAttribute(stage_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = 2 * smCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w)
    stage_kernel<V><<<grid, kThreads, smem>>>(dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink);
  CK(cudaDeviceSynchronize());
  unsigned long long zero[8] = {0};
  CK(cudaMemcpyToSymbol(g_cycles, zero, sizeof(zero)));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r)
    stage_kernel<V><<<grid, kThreads, smem>>>(dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long cyc[8];
  CK(cudaMemcpyFromSymbol(cyc, g_cycles, sizeof(cyc)));
  const double tiles = (double)cyc[4];
  std::printf(
    "%-34s %8.3f ms/launch  %7.1f GB/s useful | cycles per tile: issue %6.0f gather %6.0f wait "
    "%6.0f consume %6.0f\n",
    name, ms / reps, usefulMB / 1e3 / (ms / reps / 1e3) , cyc[0] / tiles, cyc[1] / tiles,
    cyc[2] / tiles, cyc[3] / tiles);
}

int
main(int argc, char** argv)
{
  const double mn = argc > 1 ? std::atof(argv[1]) : 2.1;
  const int reps = argc > 2 ? std::atoi(argv[2]) : 20;
  const int nTiles = (int)(mn * 1e6 / kOwn);
  const int64_t nNodes = (int64_t)nTiles * kOwn;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  std::printf("%s, %d SMs; %d tiles of %d nodes (%lld nodes), %d comps\n", prop.name,
              prop.multiProcessorCount, nTiles, kOwn, (long long)nNodes, kComps);
  /* halo: runs of consecutive slots from six neighbouring tiles (a face of a
   * 6 x 6 x 5.3 brick is 30..36 nodes: runs of 6), ~200 per tile */
  std::mt19937 rng(20261017);
  std::vector<TileHdr> tiles(nTiles);
  std::vector<int32_t> halo;
  const int nb[6] = {-1, 1, -24, 24, -580, 580}; /* tile-index offsets of the neighbours */
  for (int t = 0; t < nTiles; ++t) {
    tiles[t].ownBegin = t * kOwn;
    tiles[t].haloPtr = (int64_t)halo.size();
    std::vector<int32_t> h;
    for (int f = 0; f < 6; ++f) {
      int u = t + nb[f];
      if (u < 0 || u >= nTiles)
        continue;
      for (int r = 0; r < 6; ++r) { /* six runs of 5..6 nodes */
        const int start = (int)(rng() % (kOwn - 6));
        const int len = 5 + (int)(rng() % 2);
        for (int k = 0; k < len; ++k)
          h.push_back(u * kOwn + start + k);
      }
    }
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    if ((int)h.size() > kHaloMax)
      h.resize(kHaloMax);
    tiles[t].nHalo = (int32_t)h.size();
    halo.insert(halo.end(), h.begin(), h.end());
  }
  const double usefulMB = ((double)nNodes + (double)halo.size()) * kComps * 8 / 1e6;
  std::printf("halo entries per node %.2f; useful bytes per launch %.1f MB\n",
              (double)halo.size() / nNodes, usefulMB);
  TileHdr* dTiles;
  int32_t* dHalo;
  double *dData, *dSink;
  CK(cudaMalloc(&dTiles, sizeof(TileHdr) * nTiles));
  CK(cudaMalloc(&dHalo, sizeof(int32_t) * halo.size()));
  CK(cudaMalloc(&dData, sizeof(double) * nNodes * kComps));
  CK(cudaMalloc(&dSink, 64));
  CK(cudaMemcpy(dTiles, tiles.data(), sizeof(TileHdr) * nTiles, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dHalo, halo.data(), sizeof(int32_t) * halo.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dData, 0, sizeof(double) * nNodes * kComps));
  /* 2-D tensor map over the SoA array: rows = components (stride nNodes * 8 B),
   * box = [18 rows][192 columns] */
  CUtensorMap hMap;
  CUtensorMap* dMap = nullptr;
  bool haveMap = false;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) ==
          cudaSuccess &&
        fn && qres == cudaDriverEntryPointSuccess) {
      const cuuint64_t dims[2] = {(cuuint64_t)nNodes, (cuuint64_t)kComps};
      const cuuint64_t strides[1] = {(cuuint64_t)nNodes * 8};
      const cuuint32_t box[2] = {(cuuint32_t)kOwn, (cuuint32_t)kComps};
      const cuuint32_t estr[2] = {1, 1};
      CUresult r = ((EncodeTiled)fn)(
        &hMap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dData, dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) {
        CK(cudaMalloc(&dMap, sizeof(CUtensorMap)));
        CK(cudaMemcpy(dMap, &hMap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        haveMap = true;
      } else
        std::printf("cuTensorMapEncodeTiled failed (%d): V3 skipped\n", (int)r);
    }
  }
  const int sm = prop.multiProcessorCount;
  run<0>("V0 SoA, 18 bulk + 8 B gather", dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink, reps, sm, usefulMB);
  run<1>("V1 per-field AoS, 6 bulk", dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink, reps, sm, usefulMB);
  run<2>("V2 record, 1 bulk + 16 B gather", dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink, reps, sm, usefulMB);
  if (haveMap)
    run<3>("V3 SoA, 1 tensor copy + 8 B gather", dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink, reps, sm, usefulMB);
  run<4>("V4 record, halo by 144 B bulk copies", dTiles, nTiles, dHalo, dData, nNodes, dMap, dSink, reps, sm, usefulMB);
  return 0;
}
