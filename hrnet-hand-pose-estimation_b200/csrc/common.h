// Host-side helpers shared by every translation unit of libhrnb.so: thread-local error string,
// launch accounting, CUDA error mapping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hrnb.h"

namespace hrnb {
int fail(int code, const char* msg);                 // records msg, returns code
int fail_cuda(cudaError_t e, const char* where);     // records "where: <cuda error>", returns HRNB_ECUDA
int check_launch(const char* kernel);                // cudaGetLastError() -> HRNB_OK / HRNB_ECUDA
void count_launch();
}  // namespace hrnb
