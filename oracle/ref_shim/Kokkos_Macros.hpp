/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <Kokkos_Macros.hpp> so
 * that a few self-contained source files of the reference (master-element
 * geometry, Peclet functions, the van Leer limiter) compile here, unmodified
 * and from where they lie under /root/reference, with plain g++ -- Kokkos, STK
 * and MPI are not installed in this image.  Written from scratch: it provides
 * only the names those files use, with serial host semantics.  See
 * oracle/Makefile.ref and oracle/ref_driver.cpp. */
#ifndef NW_REF_SHIM_KOKKOS_MACROS_HPP
#define NW_REF_SHIM_KOKKOS_MACROS_HPP
#define KOKKOS_FUNCTION
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_DEFAULTED_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [ =, *this ]
#define KOKKOS_RESTRICT __restrict__
#endif
