// C-ABI glue: error string, version, device queries.  See include/stllm_b200.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_set>

#include "common.cuh"

static thread_local char g_last_error[1024] = "";

void stllm_set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" const char* stllm_last_error(void) { return g_last_error; }
extern "C" int stllm_abi_version(void) { return 1; }
extern "C" int stllm_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return STLLM_ERR_CUDA;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return STLLM_ERR_CUDA;
  return n;
}

// STLLM_PDL=0 disables programmatic dependent launch (default: on)
bool stllm_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STLLM_PDL");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// STLLM_CARVEOUT=1: see common.cuh
bool stllm_carveout_pin_once(const void* fn) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("STLLM_CARVEOUT");
    enabled = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled) return false;
  static std::mutex mu;
  static std::unordered_set<const void*> seen;
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert(fn).second;
}
