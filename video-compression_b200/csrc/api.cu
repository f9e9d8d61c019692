// Error handling, version and device queries of the b200vc C-ABI (include/b200vc.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b200vc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return B200VC_ECUDA;
  }
  return B200VC_OK;
}

int sm_count() {
  // Per-device cache; immutable after first query (no mutable global state that affects results).
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200vc

extern "C" {

int b200vc_version(void) { return B200VC_VERSION; }

const char* b200vc_last_error(void) { return b200vc::g_err; }

int b200vc_sm_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    b200vc::set_error("b200vc_sm_count: no CUDA device");
    return B200VC_ECUDA;
  }
  return b200vc::sm_count();
}

int b200vc_reduce_blocks(int64_t elems_per_sample) {
  // Deterministic function of the size only (NOT of the device), so that partial-sum shapes -- and thus
  // the fp64 totals -- are identical on every GPU of a sharded run.  Capped at 148 SMs x 8 resident CTAs.
  if (elems_per_sample <= 0) return 1;
  int64_t b = (elems_per_sample + 1023) / 1024;  // 256 threads x one 128-bit access each
  if (b > 1184) b = 1184;
  if (b < 1) b = 1;
  return (int)b;
}
}
