// Host-side helpers shared by the C-ABI translation units: thread-local error text, CUDA error
// propagation and the sm_100 architecture gate.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "../../include/llamarec_b200.h"

namespace lrb {

char* last_error_buffer();          // defined in api.cu (thread-local, 512 bytes)
int set_error(int code, const char* fmt, ...);
int check_arch();                   // LRB_OK on sm_100, LRB_ERR_ARCH otherwise (cached per device)
int device_sm_count();              // number of SMs of the current device (cached per device)

#define LRB_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::lrb::set_error(static_cast<int>(_e), "%s failed: %s", #expr,             \
                              cudaGetErrorString(_e));                                  \
  } while (0)

#define LRB_REQUIRE(cond, ...)                                                          \
  do {                                                                                  \
    if (!(cond)) return ::lrb::set_error(LRB_ERR_BAD_ARG, __VA_ARGS__);                 \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace lrb
