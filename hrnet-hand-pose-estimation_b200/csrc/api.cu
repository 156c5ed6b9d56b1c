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
}  // namespace hrnb

extern "C" const char* hrnb_last_error(void) { return hrnb::g_err; }
extern "C" int hrnb_abi_version(void) { return HRNB_ABI_VERSION; }
extern "C" int64_t hrnb_launch_count(void) { return hrnb::g_launches.load(); }
