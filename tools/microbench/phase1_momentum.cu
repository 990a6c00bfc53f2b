/*
 * phase1_momentum.cu -- microbenchmark of the momentum tile kernel's PHASE 1
 * alone (per-edge physics from shared-memory node data into the shared-memory
 * result array; no global traffic inside the timed loop).  Design aid for
 * is the phase bound by resident warps (latency) or by
 * instruction issue, and does instruction-level parallelism (two edges per
 * thread per iteration) help?  Uses the product's physics header unchanged.
 *
 * A synthetic tile: 176 staged nodes x 18 components (SoA, as today), 512
 * tile edges with random end nodes (two per thread), Peclet fused as in the
 * momentum kernel's default configuration; sized so that 1, 2 or 3 CTAs fit.
 *
 *   ILP1 / ILP2          : one / two edges per thread per loop iteration
 *   CTAs per SM 1, 2, 3  : forced with dynamic shared memory padding
 *
 * Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
 *              -Iinclude -Inalu-wind_b200/csrc -o phase1_momentum \
 *              tools/microbench/phase1_momentum.cu
 * Run:    ./phase1_momentum [passes=200]
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "edge_physics.h"

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__,                   \
                   cudaGetErrorString(e_));                                    \
      std::exit(1);                                                            \
    }                                                                          \
  } while (0)

using namespace nw;

constexpr int ND = 3;
constexpr int NC = 2 * ND + ND * ND + 3; /* 18 */
constexpr int kNodes = 176; /* sized so that three CTAs still fit one SM */
constexpr int kEdges = 512; /* two edges per thread */
constexpr int kThreads = 256;
constexpr int NRES = 4 + ND;

struct TileData
{
  double node[NC * kNodes];
  double area[ND * kEdges];
  double mdot[kEdges];
  uint32_t lr[kEdges]; /* l | r << 16 */
};

__device__ __forceinline__ void
edge_compute(
  const double* s_node, const double* s_area, const double* s_mdot, const uint32_t* s_lr,
  int j, const nw_momentum_opts& o, double* res)
{
  const uint32_t p = s_lr[j];
  const int l = p & 0xffff, r = p >> 16;
  MomNode<ND> L, R;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    L.x[d] = s_node[d * kNodes + l];
    R.x[d] = s_node[d * kNodes + r];
    L.u[d] = s_node[(ND + d) * kNodes + l];
    R.u[d] = s_node[(ND + d) * kNodes + r];
  }
#pragma unroll
  for (int d = 0; d < ND * ND; ++d) {
    L.g[d] = s_node[(2 * ND + d) * kNodes + l];
    R.g[d] = s_node[(2 * ND + d) * kNodes + r];
  }
  L.mu = s_node[(2 * ND + ND * ND) * kNodes + l];
  R.mu = s_node[(2 * ND + ND * ND) * kNodes + r];
  L.rho = s_node[(2 * ND + ND * ND + 1) * kNodes + l];
  R.rho = s_node[(2 * ND + ND * ND + 1) * kNodes + r];
  L.mask = s_node[(2 * ND + ND * ND + 2) * kNodes + l];
  R.mask = s_node[(2 * ND + ND * ND + 2) * kNodes + r];
  double av[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d)
    av[d] = s_area[d * kEdges + j];
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
  const double pecfac = peclet_eval(o.pf, peclet_number<ND>(pl, pr, o.pec_eps));
  MomResult<ND> m;
  momentum_edge<ND>(L, R, av, s_mdot[j], pecfac, o, m);
  momentum_block_entry<ND>(m, av, o.relax_fac, 0, 0, res[0], res[1], res[2], res[3]);
#pragma unroll
  for (int d = 0; d < ND; ++d)
    res[4 + d] = m.flux[d];
}

__device__ unsigned long long g_cyc[2];

template <int ILP, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
phase1_kernel(const TileData* __restrict__ td, nw_momentum_opts o, int passes, double* sink)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_node = reinterpret_cast<double*>(smem_raw);
  double* s_area = s_node + NC * kNodes;
  double* s_mdot = s_area + ND * kEdges;
  double* s_res = s_mdot + kEdges;
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(s_res + NRES * kEdges);
  const int tid = threadIdx.x;
  for (int i = tid; i < NC * kNodes; i += kThreads)
    s_node[i] = td->node[i];
  for (int i = tid; i < ND * kEdges; i += kThreads)
    s_area[i] = td->area[i];
  for (int i = tid; i < kEdges; i += kThreads) {
    s_mdot[i] = td->mdot[i];
    s_lr[i] = td->lr[i];
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    if (ILP == 1) {
      for (int j = tid; j < kEdges; j += kThreads) {
        double res[NRES];
        edge_compute(s_node, s_area, s_mdot, s_lr, j, o, res);
#pragma unroll
        for (int k = 0; k < NRES; ++k)
          s_res[k * kEdges + j] = res[k];
      }
    } else {
      for (int j = tid; j < kEdges; j += 2 * kThreads) {
        const int j2 = j + kThreads;
        double ra[NRES], rb[NRES];
        edge_compute(s_node, s_area, s_mdot, s_lr, j, o, ra);
        edge_compute(s_node, s_area, s_mdot, s_lr, j2 < kEdges ? j2 : j, o, rb);
#pragma unroll
        for (int k = 0; k < NRES; ++k) {
          s_res[k * kEdges + j] = ra[k];
          if (j2 < kEdges)
            s_res[k * kEdges + j2] = rb[k];
        }
      }
    }
    __syncthreads(); /* the product has a block barrier after phase 1 */
  }
  const long long t1 = clock64();
  if (tid == 0) {
    atomicAdd(&g_cyc[0], (unsigned long long)(t1 - t0));
    atomicAdd(&g_cyc[1], 1ull);
    if (s_res[3] == 12345.678)
      sink[0] = s_res[5];
  }
}

template <int ILP, int MINB>
static void
run(const char* name, const TileData* dTile, const nw_momentum_opts& o, int passes, int sm,
    double* dSink, double clockGHz)
{
  /* shared memory: data + padding so that exactly MINB CTAs fit an SM */
  const size_t need = sizeof(double) * (NC * kNodes + ND * kEdges + kEdges + NRES * kEdges) +
                      sizeof(uint32_t) * kEdges;
  const size_t pad = (size_t)(220 * 1024) / MINB;
  const size_t smem = pad > need ? pad : need;
  CK(cudaFuncSetAttribute(phase1_kernel<ILP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, phase1_kernel<ILP, MINB>));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phase1_kernel<ILP, MINB>, kThreads, smem));
  const int grid = sm * MINB;
  phase1_kernel<ILP, MINB><<<grid, kThreads, smem>>>(dTile, o, 4, dSink);
  CK(cudaDeviceSynchronize());
  unsigned long long zero[2] = {0, 0};
  CK(cudaMemcpyToSymbol(g_cyc, zero, sizeof(zero)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  phase1_kernel<ILP, MINB><<<grid, kThreads, smem>>>(dTile, o, passes, dSink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long cyc[2];
  CK(cudaMemcpyFromSymbol(cyc, g_cyc, sizeof(cyc)));
  const double perPass = (double)cyc[0] / cyc[1] / passes;
  const double gedges = (double)grid * passes * kEdges / (ms * 1e-3) / 1e9;
  std::printf(
    "%-22s regs %3d, CTAs/SM %d (occupancy query %d): %7.0f cycles per tile pass per CTA, "
    "%6.2f Gedges/s whole chip (physics only)\n",
    name, fa.numRegs, MINB, occ, perPass, gedges);
  (void)clockGHz;
}

int
main(int argc, char** argv)
{
  const int passes = argc > 1 ? std::atoi(argv[1]) : 200;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  std::printf("%s, %d SMs; tile: %d nodes, %d edges, %d passes\n", prop.name,
              prop.multiProcessorCount, kNodes, kEdges, passes);
  std::mt19937 rng(20261017);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<TileData> h(1);
  TileData& t = h[0];
  /* nodes on a jittered 8 x 8 x 3 lattice (spacing 1), smooth + noisy fields */
  for (int n = 0; n < kNodes; ++n) {
    const int i = n % 8, j = (n / 8) % 8, k = n / 64;
    const double x[3] = {i + 0.1 * U(rng), j + 0.1 * U(rng), k + 0.1 * U(rng)};
    for (int d = 0; d < 3; ++d) {
      t.node[d * kNodes + n] = x[d];
      t.node[(3 + d) * kNodes + n] = 7.0 + std::sin(0.3 * x[d]) + 0.05 * U(rng);
    }
    for (int d = 0; d < 9; ++d)
      t.node[(6 + d) * kNodes + n] = 0.3 * std::cos(0.2 * x[d % 3]) + 0.01 * U(rng);
    t.node[15 * kNodes + n] = 1.2e-5 * (1.0 + 0.1 * U(rng));
    t.node[16 * kNodes + n] = 1.178 * (1.0 + 0.01 * U(rng));
    t.node[17 * kNodes + n] = 1.0;
  }
  for (int e = 0; e < kEdges; ++e) {
    int l = (int)(rng() % kNodes), r = (int)(rng() % kNodes);
    if (r == l)
      r = (l + 1) % kNodes;
    t.lr[e] = (uint32_t)l | ((uint32_t)r << 16);
    double dx[3], len2 = 0.0;
    for (int d = 0; d < 3; ++d) {
      dx[d] = t.node[d * kNodes + r] - t.node[d * kNodes + l];
      len2 += dx[d] * dx[d];
    }
    for (int d = 0; d < 3; ++d)
      t.area[d * kEdges + e] = 0.3 * dx[d] / std::sqrt(len2) + 0.02 * (U(rng) - 0.5);
    t.mdot[e] = 2.0 * (U(rng) - 0.4);
  }
  TileData* dTile;
  double* dSink;
  CK(cudaMalloc(&dTile, sizeof(TileData)));
  CK(cudaMalloc(&dSink, 64));
  CK(cudaMemcpy(dTile, h.data(), sizeof(TileData), cudaMemcpyHostToDevice));
  nw_momentum_opts o{};
  o.include_divu = 0.0;
  o.alpha = 0.0;
  o.alpha_upw = 1.0;
  o.ho_upwind = 1.0;
  o.relax_fac = 0.7;
  o.use_limiter = 1;
  o.eps = 1e-16;
  o.fuse_peclet = 1;
  o.pf.form = NW_PECLET_CLASSIC;
  o.pf.a = 1.0;
  o.pf.b = 0.0;
  o.pec_eps = 1e-16;
  o.diag_field = -1;
  const int sm = prop.multiProcessorCount;
  const double ghz = prop.clockRate * 1e-6;
  run<1, 1>("ILP1, 8 warps/SM", dTile, o, passes, sm, dSink, ghz);
  run<1, 2>("ILP1, 16 warps/SM", dTile, o, passes, sm, dSink, ghz);
  run<1, 3>("ILP1, 24 warps/SM", dTile, o, passes, sm, dSink, ghz);
  run<2, 1>("ILP2, 8 warps/SM", dTile, o, passes, sm, dSink, ghz);
  run<2, 2>("ILP2, 16 warps/SM", dTile, o, passes, sm, dSink, ghz);
  std::printf("product kernel at 128^3: 6.39 M edges x 1.18 tile-edge records in 0.479 ms = "
              "15.7 G records/s through all phases\n");
  return 0;
}
