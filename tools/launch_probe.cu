// Launch-floor probe (debug tool): back-to-back latency of kernels that only do the conv kernel's prologue/teardown.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/launch_probe.bin tools/launch_probe.cu
#include <cstdio>
#include "../hrnet-hand-pose-estimation_b200/csrc/ptx.cuh"
using namespace hrnb;

// mode bits: 1 = TMEM alloc/dealloc, 2 = mbarrier init + fence, 4 = global load before sync, 8 = PDL instructions
__global__ void __launch_bounds__(320, 2) k_probe(int mode, int cols, const float* g, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 256);
  float* bs = reinterpret_cast<float*>(smem + 512);
  const int warp = threadIdx.x >> 5;
  if (mode & 8) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((mode & 2) && threadIdx.x == 0) {
    for (int i = 0; i < 24; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if ((mode & 1) && warp == 1) { tmem_alloc(tptr, cols); tmem_relinquish(); }
  if ((mode & 4) && warp >= 2) bs[threadIdx.x] = g[threadIdx.x];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (mode & 8) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0 && (mode & 4)) out[0] = bs[64];
  tc_fence_before_sync();
  __syncthreads();
  if ((mode & 1) && warp == 1) { tc_fence_after_sync(); tmem_dealloc(*tptr, cols); }
}

int main() {
  float *g, *out;
  cudaMalloc(&g, 4096); cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 400;
  for (int smem : {8 * 1024, 100 * 1024}) {
    for (int grid : {148, 296}) {
      for (int mode : {0, 1, 2, 4, 7, 8, 15}) {
        for (int pdl = 0; pdl < 2; ++pdl) {
          if (pdl && !(mode & 8)) continue;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = attr; cfg.numAttrs = pdl;
          int cols = 256;
          for (int i = 0; i < 20; ++i) cudaLaunchKernelEx(&cfg, k_probe, mode, cols, (const float*)g, out);
          cudaDeviceSynchronize();
          cudaEventRecord(e0);
          for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, k_probe, mode, cols, (const float*)g, out);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          printf("smem %3dKB grid %3d mode %2d pdl %d : %.2f us/launch %s\n", smem / 1024, grid, mode, pdl, ms * 1e3 / reps,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
    }
  }
  return 0;
}
