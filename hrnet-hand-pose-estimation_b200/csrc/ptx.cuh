// Thin inline-PTX wrappers for the sm_100a features the HRNet kernels use:
// mbarrier, 1-D bulk async copy (TMA engine, UBLKCP), cp.async, tcgen05 (alloc / mma / commit / ld)
// and the proxy / thread-sync fences between them.  No CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hrnb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One lane of the (converged) warp returns true.  The MMA / TMA issue loops are executed by the WHOLE warp with
// warp-uniform operands and only the async instruction itself is predicated on this: that keeps descriptors in
// uniform registers (UTCHMMA / UBLKCP take UR operands; a loop run by a single lane pays ~5 R2UR per MMA).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %1;\n\t"
      "@%%px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// Programmatic dependent launch: let the next kernel of the stream be scheduled now (its blocks become resident and
// wait), then wait until the previous kernel has completed and its writes are visible.  Both are no-ops for a kernel
// that was launched without the programmatic attribute / has no such dependent.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier.init visible to the async proxy (TMA / tcgen05.commit arrive on them)
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Hang diagnostics (debug build only: nvcc -DHRNB_HANG_RECORDS): a mapped HOST buffer (hrnb_hang_init) that survives the
// trap.  Every warp that times out leaves a record: word 0 = blockDim.x << 48 | gridDim.x << 32 | blockIdx.x << 8 | warp,
// word 1 = barrier smem address << 8 | parity (blockDim.x tells the kernel: 608 conv, 736 conv gather, 192 wgrad; the
// address tells the barrier).  Not in the default build: the extra operand set-up in front of every wait was measurable in
// the wait-heavy wgrad kernel.
static __device__ unsigned long long* g_hang_buf = nullptr;   // one copy per translation unit, bound by hrnb_hang_init
constexpr int kHangSlots = 255;
__device__ __forceinline__ void mbar_timeout(uint32_t bar_addr, uint32_t parity) {
#ifdef HRNB_HANG_RECORDS
  unsigned long long* b = g_hang_buf;
  if (b != nullptr && (threadIdx.x & 31) == 0) {
    const unsigned warp = threadIdx.x >> 5;
    const unsigned slot = (blockIdx.x * 23u + warp) % kHangSlots;
    b[2 + 2 * slot] = ((unsigned long long)blockDim.x << 48) | ((unsigned long long)gridDim.x << 32) |
                      ((unsigned long long)blockIdx.x << 8) | warp;
    b[3 + 2 * slot] = ((unsigned long long)bar_addr << 8) | parity;
    b[0] = 1ull;
    __threadfence_system();
  }
  const uint64_t t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < 20000000ull) {}   // 20 ms: let the other stuck warps leave their records too
#endif
  __trap();
}
// Bounded wait: a protocol bug must trap (launch error) instead of hanging the GPU box.  The guard counts failed probes -
// nothing else may sit in or in front of this loop (a %globaltimer read per probe put microseconds on every pipeline
// hand-off).  A failed probe that suspends takes ~4 us [measured], so 2^22 probes bound a wait by ~16 s; probes that
// return at once by >= 0.2 s - no legitimate wait on this path is longer than milliseconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifdef HRNB_HANG_RECORDS
    if (++spins > (1u << 20)) mbar_timeout(smem_u32(bar), parity);
#else
    if (++spins > (1u << 22)) __trap();
#endif
  }
}

// ----------------------------------------------------------------------------------------------
// async copies
// ----------------------------------------------------------------------------------------------
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar` as tx bytes.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Ampere-style 16-byte cp.async (LDGSTS) used by the gather producer.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to async-proxy readers (tcgen05.mma operand fetch)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_NONE ("interleaved" canonical layout):
//   core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes (row r at +16*r);
//   LBO = byte distance between the two core matrices adjacent in K inside one K=16 MMA,
//   SBO = byte distance between core matrices adjacent in M (or N).
// Bit layout follows cute::UMMA::SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset [49,52)=0, layout_type [61,64)=0).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

// Instruction descriptor for kind::f16, A=B=bf16 (K-major both), D=fp32, M=128, N=n (multiple of 16).
// Bit layout follows cute::UMMA::InstrDescriptor.
__device__ __forceinline__ uint32_t make_idesc_bf16_m128(uint32_t n) {
  return (1u << 4)      // c_format  = F32
         | (1u << 7)    // a_format  = BF16
         | (1u << 10)   // b_format  = BF16
         | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 -> packed bf16x2 with ReLU fused into the conversion (lo = a, hi = b)
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed fp32 pairs (sm_100: FADD2 / FFMA2 issue two fp32 operations per instruction)
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace hrnb
