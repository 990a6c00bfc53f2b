/* oracle/ref_shim/nalu: TEST INFRASTRUCTURE.  Shadows the reference's ngp_utils/NgpFieldUtils.h on
 * the include path of the edge-algorithm build (oracle/Makefile.ref); the
 * stand-ins live in RefHarness.h. */
#ifndef NW_REF_SHADOW_NGP_UTILS_NGPFIELDUTILS_H
#define NW_REF_SHADOW_NGP_UTILS_NGPFIELDUTILS_H
#include <RefHarness.h>
#endif
