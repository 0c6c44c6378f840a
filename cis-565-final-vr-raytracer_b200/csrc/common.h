// common.h — error plumbing shared by the host-side translation units of libeidola.so.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include "eidola.h"

namespace eid {

// thread-local message returned by eid_last_error()
inline std::string& lastError() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  lastError() = buf;
  return code;
}

struct Error {   // internal exception type; never crosses the C-ABI
  int code;
  std::string msg;
};
[[noreturn]] inline void raise(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

}  // namespace eid

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CUDA_CHECK(x)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) eid::raise(EID_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#endif

// wraps a C-ABI body: converts internal exceptions into status codes + eid_last_error()
#define EID_TRY try {
#define EID_CATCH                                                                   \
  }                                                                                 \
  catch (const eid::Error& e) { eid::lastError() = e.msg; return e.code; }          \
  catch (const std::exception& e) { eid::lastError() = e.what(); return EID_ERR_INVALID; } \
  catch (...) { eid::lastError() = "unknown error"; return EID_ERR_INVALID; }
