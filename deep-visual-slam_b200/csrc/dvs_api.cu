// Library-level entry points of libdvsloss.so (version, error strings).
#include "dvs_host.h"

namespace dvs {
int& last_cuda_error() {
  static thread_local int e = 0;
  return e;
}
}  // namespace dvs

extern "C" int dvs_version(void) { return 200; }   // 0.2.0: round-2 ABI (input formats, pose parameters, decoder kernels)

extern "C" int dvs_last_cuda_error(void) { return dvs::last_cuda_error(); }

extern "C" const char* dvs_error_string(int code) {
  switch (code) {
    case DVS_OK: return "ok";
    case DVS_EINVAL: return "invalid argument (shape, null pointer, unsupported N or S)";
    case DVS_ECUDA: return "CUDA runtime error (see dvs_last_cuda_error)";
    case DVS_EWORKSPACE: return "workspace pointer null or not 256-byte aligned";
  }
  return "unknown error code";
}
