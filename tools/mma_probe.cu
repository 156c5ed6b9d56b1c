// tcgen05.mma issue-rate probe (debug tool): cycles per 128xNx16 bf16 MMA for different smem layouts and
// numbers of independent accumulators.  The issue loop is fully unrolled with compile-time offsets so that
// it measures the tensor pipe, not scalar address arithmetic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe.bin tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../hrnet-hand-pose-estimation_b200/csrc/ptx.cuh"
using namespace hrnb;

// MODE 0: no swizzle, plane layout (LBO = 4096, SBO = 128); MODE 1: SW128 (SBO = 1024); MODE 2: no swizzle, start + 16 B
template <int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, long long* out, unsigned a_lbo_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  if (__shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) == 0) {
    const uint32_t a0 = smem_u32(smem) + (MODE == 2 ? 16u : 0u), b0 = smem_u32(smem + 96 * 1024);
    const uint32_t idesc = make_idesc_bf16_m128((uint32_t)N);
    const uint32_t layout_hi = (MODE == 1 ? (2u << 29) : 0u) | (1u << 14);
    const uint32_t a_hi = layout_hi | ((MODE == 1 ? 1024u : 128u) >> 4);
    const uint32_t b_hi = a_hi;
    const uint32_t a_lo = (a0 >> 4) | ((MODE == 1 ? 1u : (a_lbo_bytes >> 4)) << 16);
    const uint32_t b_lo = (b0 >> 4) | ((MODE == 1 ? 1u : (((uint32_t)N * 16u) >> 4)) << 16);
    const uint32_t a_step = MODE == 1 ? 2u : (2u * a_lbo_bytes >> 4), b_step = MODE == 1 ? 2u : ((2u * (uint32_t)N * 16u) >> 4);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < reps; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + (u & 3) * a_step);
          const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + (u & 3) * b_step);
          if (elect_one_sync()) umma_bf16_ss(tb + (uint32_t)((u % NACC) * N), ad, bd, idesc, (i > 0 || u >= NACC) ? 1u : 0u);
        }
      }
      __syncwarp();
      if (elect_one_sync()) umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      long long t1 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 0) out[rep] = t1 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int MODE, int NACC>
void run(const char* name, long long* d, unsigned lbo = 4096) {
  cudaFuncSetAttribute(probe<MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 4000;
  for (int N : {32, 64, 128, 256}) {
    if (NACC * N > 512) continue;
    probe<MODE, NACC><<<148, 128, 160 * 1024>>>(N, reps, d, lbo);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3] = {0, 0, 0};
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%-28s lbo=%5u nacc=%d N=%3d  cycles/MMA %.1f  ideal %.0f  %s\n", name, lbo, NACC, N, (double)h[2] / reps, N / 2.0,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (unsigned lbo : {4096u, 9280u, 9344u, 10304u, 2048u, 4160u, 2368u}) { run<0, 4>("noswz plane", d, lbo); run<2, 4>("noswz plane +16B", d, lbo); }
  return 0;
}
