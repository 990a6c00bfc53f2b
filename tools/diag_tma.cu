// standalone check of cp.async.bulk + mbarrier on this GPU (diagnostic only)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const double* src, double* dst, int n)
{
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bytes = n * 8;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(s32(sm)), "l"(src + (size_t)blockIdx.x * n), "r"(bytes), "r"(s32(&bar)) : "memory");
  }
  uint32_t done; long spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done) : "r"(s32(&bar)), "r"(0) : "memory");
    if (++spins > 20000000) { if (threadIdx.x==0) printf("block %d: TMA wait timed out\n", blockIdx.x); return; }
  } while (!done);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[(size_t)blockIdx.x * n + i] = sm[i] * 2.0;
}
int main()
{
  const int n = 512, nb = 64;
  double *a, *b; cudaMalloc(&a, n*nb*8); cudaMalloc(&b, n*nb*8);
  double* h = new double[n*nb]; for (int i = 0; i < n*nb; ++i) h[i] = i;
  cudaMemcpy(a, h, n*nb*8, cudaMemcpyHostToDevice);
  k<<<nb, 256, n*8>>>(a, b, n);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, b, n*nb*8, cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < n*nb; ++i) if (h[i] != 2.0*i) ++bad;
  printf("tma diag: %d mismatches\n", bad);
  return bad != 0;
}
