// Host-side helpers shared by the translation units of libdvsloss.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/dvsloss.h"

namespace dvs {
// cudaError_t of the last failing CUDA call made by this library on the calling host thread.
int& last_cuda_error();
struct FusedParams;
// dvs_pair.cu: the two-source tile kernel, dispatched on (want_grad, io_flags)
cudaError_t launch_pair_kernel(const FusedParams& p, int nblk, cudaStream_t st);
}  // namespace dvs

#define DVS_CUDA_TRY(expr)                          \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      ::dvs::last_cuda_error() = (int)_e;           \
      return DVS_ECUDA;                             \
    }                                               \
  } while (0)
