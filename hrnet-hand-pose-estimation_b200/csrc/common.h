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
unsigned long long* hang_buffer_device_ptr();        // mapped host buffer of the mbarrier time-out records (api.cu), or nullptr
int bind_hang_buffer_conv();                         // per translation unit: point its g_hang_buf at the buffer
int bind_hang_buffer_wgrad();
// Every kernel that allocates tensor memory asks for at least this much dynamic shared memory, so that no two of them are
// ever co-resident on one SM (2 x 116 KB > 227 KB): a conv CTA may hold all 512 TMEM columns, and a second CTA blocked in
// tcgen05.alloc next to it is the one kind of cross-kernel coupling these kernels could have (the TMEM deadlock chain under
// programmatic dependent launch described in conv_tc.cu was of that kind).  No measurable cost at batch 64 [28.73 vs 28.75
// ms/step]; it did NOT cure the round-1 mbarrier time-out (that was the producer's prefetch wait, conv_tc.cu).
// hrnb_debug_set(6, 1) / HRNB_TMEM_SHARE=1 turns the padding off.
constexpr long long kTmemExclusiveSmem = 116 * 1024;
extern int g_debug[16];                               // hrnb_debug_set knobs (conv_tc.cu); [4] != 0: PDL for the elementwise / wgrad kernels

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
