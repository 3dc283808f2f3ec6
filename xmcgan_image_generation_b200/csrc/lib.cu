// Library-level plumbing of libxmc.so: error strings, device queries.
#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"

namespace xmc {

static thread_local char g_err[256] = "";

void set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  (void)cudaGetLastError();  // clear the sticky-less error state
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs).
int num_sms() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> sms[kMaxDev];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool cached = dev >= 0 && dev < kMaxDev;
  if (cached) {
    const int v = sms[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  if (cached) sms[dev].store(v, std::memory_order_relaxed);
  return v;
}

// Number of SMs the persistent tensor-core kernels may occupy (xmc_set_sm_limit): all of them unless the caller has
// reserved some for a collective kernel that runs concurrently. Work decomposition (tile widths, K splits) is always
// planned on the full SM count, so results do not depend on the limit.
static std::atomic<int> g_sm_limit{0};
int grid_sms() {
  const int n = num_sms(), l = g_sm_limit.load(std::memory_order_relaxed);
  return (l > 0 && l < n) ? l : n;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("XMC_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

}  // namespace xmc

extern "C" int xmc_set_sm_limit(int sms) {
  xmc::g_sm_limit.store(sms > 0 ? sms : 0, std::memory_order_relaxed);
  return XMC_OK;
}

extern "C" const char* xmc_strerror(int code) {
  switch (code) {
    case XMC_OK: return "ok";
    case XMC_EINVAL: return "invalid descriptor or unsupported shape";
    case XMC_ECUDA: return "CUDA call failed (see xmc_last_cuda_error)";
    case XMC_EALIGN: return "pointer or pitch not 16-byte aligned";
    default: return "unknown error";
  }
}

extern "C" const char* xmc_last_cuda_error(void) { return xmc::g_err; }
extern "C" int xmc_version(void) { return 200; }
extern "C" int xmc_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(XmcConvDesc);
    case 1: return (int)sizeof(XmcWgradDesc);
    case 2: return (int)sizeof(XmcBnDesc);
    case 3: return (int)sizeof(XmcPrepEntry);
    case 4: return (int)sizeof(XmcSnEntry);
    default: return -1;
  }
}
extern "C" int xmc_num_sms(void) { return xmc::num_sms(); }
