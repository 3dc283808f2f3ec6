// Host-side helpers shared by all translation units of libxmc.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/xmc.h"

namespace xmc {

typedef __nv_bfloat16 bf16;

void set_cuda_error(cudaError_t e);
int num_sms();
int grid_sms();  // SMs the persistent GEMM kernels may occupy (xmc_set_sm_limit)

#define XMC_CUDA_CHECK(expr)                 \
  do {                                       \
    cudaError_t _e = (expr);                 \
    if (_e != cudaSuccess) {                 \
      ::xmc::set_cuda_error(_e);             \
      return XMC_ECUDA;                      \
    }                                        \
  } while (0)

#define XMC_LAUNCH_CHECK() XMC_CUDA_CHECK(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace xmc
