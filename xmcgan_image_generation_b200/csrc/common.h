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

// Programmatic dependent launch (sm_90+): the kernel is launched with the stream-serialisation attribute, so its thread
// blocks may become resident — and run their prologue up to pdl_wait() — while the previous kernel of the stream is
// still draining; pdl_wait() (griddepcontrol.wait) returns once that kernel has completed and its writes are visible.
// Used for the tensor-core kernels (see pdl.cuh for the measurement). Captured into a CUDA graph the attribute becomes
// a programmatic edge. XMC_PDL=0 (read once per process) launches plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace xmc
