// Library-level state: error string, launch counter, device properties, GEMM selection.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace tnpy {

static thread_local char g_error[512] = "";
std::atomic<int64_t> g_launch_count{0};
static std::atomic<int> g_gemm_algo{TNPY_GEMM_AUTO};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

int current_gemm_algo() { return g_gemm_algo.load(std::memory_order_relaxed); }

__device__ double g_one_storage[2] = {1.0, 1.0};
const double* device_one() {
  static const double* ptr = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_one_storage) == cudaSuccess) ptr = static_cast<const double*>(p);
  });
  return ptr;
}

}  // namespace tnpy

extern "C" int tnpy_version(void) { return 100; }
extern "C" const char* tnpy_last_error(void) { return tnpy::g_error; }
extern "C" int64_t tnpy_launch_count(void) { return tnpy::g_launch_count.load(); }
extern "C" int tnpy_set_gemm_algo(int algo) {
  if (algo < TNPY_GEMM_AUTO || algo > TNPY_GEMM_OZAKI) {
    tnpy::set_error("tnpy_set_gemm_algo: unknown algo %d", algo);
    return TNPY_EINVAL;
  }
  tnpy::g_gemm_algo.store(algo);
  return TNPY_OK;
}
