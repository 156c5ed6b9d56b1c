// Error reporting, launch accounting and ABI version of libhrnb.so.
#include <atomic>
#include <cstdio>
#include <cstring>
#include "common.h"

namespace hrnb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int fail_cuda(cudaError_t e, const char* where) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return HRNB_ECUDA;
}
int check_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, kernel);
  return HRNB_OK;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static unsigned long long* g_hang_host = nullptr;
constexpr int kHangWords = 2 + 2 * 255;
unsigned long long* hang_buffer_device_ptr() {
  if (g_hang_host == nullptr) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, kHangWords * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    std::memset(h, 0, kHangWords * sizeof(unsigned long long));
    g_hang_host = (unsigned long long*)h;
  }
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, g_hang_host, 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (unsigned long long*)d;
}
}  // namespace hrnb

// Arm the hang diagnostics on the current device (call once per device, outside stream capture): kernels that hit the
// mbarrier time-out then leave (kernel, CTA, warp, barrier) records in host memory before they trap.
extern "C" int hrnb_hang_init(void) {
  int rc = hrnb::bind_hang_buffer_conv();
  if (rc) return rc;
  return hrnb::bind_hang_buffer_wgrad();
}
// Copy up to n 64-bit words of the record buffer (word 0 != 0: something timed out; records from word 2, two words each,
// see ptx.cuh).  Plain host memory: readable after the CUDA context died.  Returns the words copied.
extern "C" int hrnb_hang_report(uint64_t* out, int n) {
  if (!out || n <= 0 || hrnb::g_hang_host == nullptr) return 0;
  if (n > hrnb::kHangWords) n = hrnb::kHangWords;
  std::memcpy(out, hrnb::g_hang_host, (size_t)n * sizeof(uint64_t));
  return n;
}

extern "C" const char* hrnb_last_error(void) { return hrnb::g_err; }
extern "C" int hrnb_abi_version(void) { return HRNB_ABI_VERSION; }
extern "C" int64_t hrnb_launch_count(void) { return hrnb::g_launches.load(); }
