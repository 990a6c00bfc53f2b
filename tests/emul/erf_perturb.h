/* TEST INFRASTRUCTURE (tests/emul, libnw_emul_erf.so only): replaces erf by a
 * version whose result is moved by a pseudo-random -2 ... +2 ulp -- the error
 * bound of CUDA's erf(double) -- so that the CPU test-suite can size the
 * tolerance of the one branch of the path (realm_has_vof_) that calls a
 * transcendental function the device does not evaluate bit for bit like the
 * host.  Force-included ahead of the product header (-include). */
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstring>
static inline double
nw_test_erf(double x)
{
  double r = ::erf(x);
  uint64_t b;
  std::memcpy(&b, &x, 8);
  b ^= b >> 29;
  b *= 0x9E3779B97F4A7C15ull;
  b ^= b >> 32;
  const int k = (int)(b % 5) - 2;
  for (int i = 0; i < (k < 0 ? -k : k); ++i)
    r = std::nextafter(r, k < 0 ? -2.0 : 2.0);
  return r;
}
#define erf nw_test_erf
