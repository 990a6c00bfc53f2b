/*
 * nw_geometry.cu -- GeometryInteriorAlg for Tet4 / Wed6 / Pyr5 element blocks
 * (src/ngp_algorithms/GeometryInteriorAlg.C:72-112, 165-225): dual nodal
 * volumes and edge area vectors from the current coordinates, on the device, so
 * a moving-mesh case refreshes the edge kernels' geometry inputs without the
 * host.  The per-element arithmetic is geometry_cvfem.h; Hex8 and Quad4 have
 * their own kernels in nw_kernels.cu.
 *
 * One thread per element, results accumulated with fp64 atomics exactly as the
 * reference does (Kokkos::atomic_add into the node / edge fields): an element
 * block is visited once per mesh motion step, the traffic is the connectivity
 * (npe + nScs int32 per element) plus scattered 8-byte reads and atomics -- the
 * kernel is a producer outside the per-iteration sweep, not a roofline kernel.
 */
#include "geometry_cvfem.h"
#include "nw_kernels.cuh"

namespace nw {

namespace {

template <int T>
__global__ void __launch_bounds__(128) geometry_cvfem_kernel(
  int64_t nElems, const int32_t* __restrict__ elemSlots /* [n][npe] node slots */,
  const int32_t* __restrict__ elemEdges /* [n][nScs]: 2 * slot + negate, -1: none */,
  const unsigned char* __restrict__ owned, const double* __restrict__ x,
  int64_t xStride, double* dualVol, double* area, int64_t areaStride)
{
  constexpr int npe = geo::Traits<T>::npe;
  constexpr int nScv = geo::Traits<T>::nScv;
  constexpr int nScs = geo::Traits<T>::nScs;
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nElems)
    return;
  double c[npe][3];
  double v[geo::Traits<T>::nSub][3];
  for (int n = 0; n < npe; ++n) {
    const int64_t sl = elemSlots[(int64_t)npe * el + n];
    for (int d = 0; d < 3; ++d)
      c[n][d] = x[(int64_t)d * xStride + sl];
  }
  geo::sub_points<T>(c, v);
  if (dualVol && (!owned || owned[el])) {
    for (int ip = 0; ip < nScv; ++ip)
      atomicAdd(dualVol + elemSlots[(int64_t)npe * el + ip], geo::scv_volume<T>(ip, v));
  }
  if (!area)
    return;
  for (int ip = 0; ip < nScs; ++ip) {
    const int32_t code = elemEdges[(int64_t)nScs * el + ip];
    if (code < 0)
      continue;
    double a[3];
    geo::scs_area<T>(ip, v, a);
    const double sg = (code & 1) ? -1.0 : 1.0;
    const int64_t slot = code >> 1;
    for (int d = 0; d < 3; ++d)
      atomicAdd(area + (int64_t)d * areaStride + slot, a[d] * sg);
  }
}

} // namespace

cudaError_t
launch_geometry_cvfem(
  int topology, int64_t nElems, const int32_t* elemSlots, const int32_t* elemEdges,
  const unsigned char* owned, const double* x, int64_t xStride, double* dualVol,
  double* area, int64_t areaStride, cudaStream_t s)
{
  if (nElems <= 0)
    return cudaSuccess;
  const unsigned blocks = (unsigned)((nElems + 127) / 128);
  switch (topology) {
  case geo::TET4:
    geometry_cvfem_kernel<geo::TET4><<<blocks, 128, 0, s>>>(
      nElems, elemSlots, elemEdges, owned, x, xStride, dualVol, area, areaStride);
    break;
  case geo::WED6:
    geometry_cvfem_kernel<geo::WED6><<<blocks, 128, 0, s>>>(
      nElems, elemSlots, elemEdges, owned, x, xStride, dualVol, area, areaStride);
    break;
  case geo::PYR5:
    geometry_cvfem_kernel<geo::PYR5><<<blocks, 128, 0, s>>>(
      nElems, elemSlots, elemEdges, owned, x, xStride, dualVol, area, areaStride);
    break;
  default:
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

} // namespace nw
