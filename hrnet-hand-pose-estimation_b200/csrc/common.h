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
extern int g_debug[8];                               // hrnb_debug_set knobs (conv_tc.cu); [4] != 0: PDL for the elementwise / wgrad kernels

// kernel launch with the programmatic-dependent-launch attribute when knob 4 is set (the kernel must call pdl_enter())
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_debug[4] ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace hrnb
