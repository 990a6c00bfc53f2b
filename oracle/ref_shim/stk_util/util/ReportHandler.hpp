/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for STK's ReportHandler
 * macros (throw std::runtime_error on a failed requirement). */
#ifndef NW_REF_SHIM_REPORTHANDLER_HPP
#define NW_REF_SHIM_REPORTHANDLER_HPP
#include <sstream>
#include <stdexcept>
#define NW_SHIM_THROW(msg)                                                     \
  do {                                                                         \
    std::ostringstream nw_os_;                                                 \
    nw_os_ << msg;                                                             \
    throw std::runtime_error(nw_os_.str());                                    \
  } while (0)
#define STK_ThrowRequire(c)                                                    \
  do {                                                                         \
    if (!(c))                                                                  \
      NW_SHIM_THROW("requirement failed: " #c);                                \
  } while (0)
#define STK_ThrowRequireMsg(c, m)                                              \
  do {                                                                         \
    if (!(c))                                                                  \
      NW_SHIM_THROW(m);                                                        \
  } while (0)
#define STK_ThrowAssert(c) STK_ThrowRequire(c)
#define STK_ThrowAssertMsg(c, m) STK_ThrowRequireMsg(c, m)
#define STK_ThrowErrorMsg(m) NW_SHIM_THROW(m)
#define STK_ThrowErrorMsgIf(c, m)                                              \
  do {                                                                         \
    if (c)                                                                     \
      NW_SHIM_THROW(m);                                                        \
  } while (0)
#define STK_NGP_ThrowRequire(c) STK_ThrowRequire(c)
#define STK_NGP_ThrowRequireMsg(c, m) STK_ThrowRequireMsg(c, m)
#define STK_NGP_ThrowAssert(c) STK_ThrowRequire(c)
#define STK_NGP_ThrowAssertMsg(c, m) STK_ThrowRequireMsg(c, m)
#define STK_NGP_ThrowErrorMsg(m) NW_SHIM_THROW(m)
#define STK_NGP_ThrowErrorMsgIf(c, m) STK_ThrowErrorMsgIf(c, m)
#endif
