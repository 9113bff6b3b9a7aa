// Library-level state: error string, launch counter, device properties, GEMM selection.
#include <stdarg.h>

#include <mutex>
#include <set>
#include <utility>

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

constexpr int kMaxDevices = 64;

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  const int dev = current_device();
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool first_time_on_device(const void* key) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> seen;
  std::lock_guard<std::mutex> lock(mu);
  return seen.emplace(key, current_device()).second;
}

namespace {
struct ThreadPinned {
  void* host = nullptr;
  ThreadPinned() {
    if (cudaMallocHost(&host, 256) != cudaSuccess) host = nullptr;
  }
  ~ThreadPinned() {
    if (host) cudaFreeHost(host);
  }
};
}  // namespace
void* thread_pinned_scratch() {
  static thread_local ThreadPinned p;
  return p.host;
}

int current_gemm_algo() { return g_gemm_algo.load(std::memory_order_relaxed); }

__device__ double g_one_storage[2] = {1.0, 1.0};
const double* device_one() {
  // the address of a __device__ symbol differs per device
  static std::atomic<const double*> ptr[kMaxDevices];
  const int dev = current_device();
  const double* p = ptr[dev].load(std::memory_order_acquire);
  if (!p) {
    void* q = nullptr;
    if (cudaGetSymbolAddress(&q, g_one_storage) == cudaSuccess) p = static_cast<const double*>(q);
    ptr[dev].store(p, std::memory_order_release);
  }
  return p;
}

}  // namespace tnpy

extern "C" int tnpy_version(void) { return 100; }
extern "C" const char* tnpy_last_error(void) { return tnpy::g_error; }
extern "C" int64_t tnpy_launch_count(void) { return tnpy::g_launch_count.load(); }
extern "C" int tnpy_set_gemm_algo(int algo) {
  if (algo < TNPY_GEMM_AUTO || algo > TNPY_GEMM_FP64) {
    tnpy::set_error("tnpy_set_gemm_algo: unknown algo %d", algo);
    return TNPY_EINVAL;
  }
  tnpy::g_gemm_algo.store(algo);
  return TNPY_OK;
}
