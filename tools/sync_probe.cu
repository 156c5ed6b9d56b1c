// Sync-cost probe (debug tool): what do tcgen05.commit / mbarrier hand-offs cost around small MMA batches?
#include <cstdio>
#include "../hrnet-hand-pose-estimation_b200/csrc/ptx.cuh"
using namespace hrnb;

// mode 0: batches of NB MMAs (N=64), each followed by commit to a ring barrier, never waited except at the end
// mode 1: same, but wait for each batch's commit before the next batch (round-trip latency)
// mode 2: like 0 but additionally a try_wait on an already-complete barrier + fence before each batch
// mode 3: like 2, plus a 2-party ping-pong with another warp per batch (consumer arrives on `empty` after seeing full)
__global__ void __launch_bounds__(128, 1) probe(int mode, int NB, int batches, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint64_t done_bar, ready[8], empty[8];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&ready[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int N = 64;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const uint32_t idesc = make_idesc_bf16_m128((uint32_t)N);
    const uint32_t hi = (1u << 14) | (128u >> 4);
    const uint32_t a_lo = (a0 >> 4) | ((4096u >> 4) << 16), b_lo = (b0 >> 4) | (((uint32_t)N * 16u >> 4) << 16);
    long long t0 = clock64();
    int stage = 0, phase = 0;
    for (int b = 0; b < batches; ++b) {
      if (mode == 2) { mbar_wait(&done_bar, 1); tc_fence_after_sync(); }           // already complete (parity 1 of a fresh barrier)
      if (mode == 3) { mbar_wait(&ready[stage], phase); tc_fence_after_sync(); }    // producer warp filled the stage
      for (int i = 0; i < NB; ++i) {
        const uint64_t ad = ((uint64_t)hi << 32) | (a_lo + (i & 3) * 512u);
        const uint64_t bd = ((uint64_t)hi << 32) | (b_lo + (i & 3) * 128u);
        if (elect_one_sync()) umma_bf16_ss(tb + (uint32_t)((i & 3) * N), ad, bd, idesc, (b > 0 || i > 3) ? 1u : 0u);
      }
      __syncwarp();
      if (mode == 3) {
        if (elect_one_sync()) umma_commit(&empty[stage]);
        if (++stage == 8) { stage = 0; phase ^= 1; }
      } else {
        if (elect_one_sync()) umma_commit(&bars[b & 15]);
        if (mode == 1) { mbar_wait(&bars[b & 15], (b >> 4) & 1); tc_fence_after_sync(); }
      }
    }
    __syncwarp();
    if (elect_one_sync()) umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  } else if (warp == 1 && mode == 3) {
    int stage = 0, phase = 0;
    for (int b = 0; b < batches; ++b) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one_sync()) mbar_arrive(&ready[stage]);
      __syncwarp();
      if (++stage == 8) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int batches = 512;
  for (int mode = 0; mode < 4; ++mode)
    for (int NB : {1, 4, 8, 16}) {
      probe<<<148, 128, 160 * 1024>>>(mode, NB, batches, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("mode %d  NB=%2d : %.0f cycles/batch  (%.1f per MMA; ideal %d per MMA) %s\n", mode, NB, (double)h / batches,
             (double)h / batches / NB, 48, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
