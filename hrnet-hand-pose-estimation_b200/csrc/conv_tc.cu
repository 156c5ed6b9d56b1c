// Implicit-GEMM convolution on tcgen05 / TMEM for the HRNet hot path (sm_100a only).
//
// Replaces the nn.Conv2d -> BatchNorm2d(eval) -> (+residual) -> ReLU chains of
// lib/models/pose_hrnet.py:28-98 (BasicBlock / Bottleneck), :187-242 (fuse convs), :335-350 (head),
// :419-458 (transitions).
//
// GEMM view:   D[M = positions, N = cout] = A[M, K = taps*cin] * W[N, K]^T,  bf16 x bf16 -> fp32 in TMEM.
//
// Activations live in HBM in the PF8 layout: channel planes of 8, each plane a flat array
// of 16-byte positions over the zero-padded image grid (shared pad row/column), batch folded in.
// That makes a 3x3/stride-1 conv a *shifted flat GEMM*: tap (r,s) of output position p reads input position
// p + (r-1)*Wp + (s-1).  One CTA therefore loads a single contiguous halo range of every input plane with
// 1-D bulk-async copies (TMA engine) and feeds all nine taps to the tensor core from that one smem copy by
// offsetting the UMMA shared-memory descriptor by `shift*16` bytes (SWIZZLE_NONE K-major canonical layout:
// rows are 16 bytes apart, so any row shift is a legal descriptor start).  Stride-2 convs cannot be
// expressed as a flat shift; they use the GATHER variant whose producer warps build each tap's A tile with
// 16-byte cp.async.
//
// Warp roles (one CTA = MB x 128 output positions x BN output channels):
//   warp 0      : bulk-copy producer (A halo chunks, packed weight tiles)           [1 elected lane]
//   warp 1      : TMEM allocator + tcgen05.mma issuer                                [1 elected lane]
//   warps 2..5  : epilogue  (tcgen05.ld -> +bias -> +residual -> ReLU -> bf16 -> coalesced 16 B stores)
//   warps 6..9  : (GATHER only) cp.async gather producers, one thread per A row
#include "ptx.cuh"
#include "common.h"

namespace hrnb {

struct ConvK {
  const __nv_bfloat16* in;
  long long in_ps;
  const __nv_bfloat16* wpk;
  const float* bias;
  const __nv_bfloat16* res;
  long long res_ps;
  void* out;
  long long out_ps;
  int P, Hp, Wp;          // output positions / padded dims
  int in_Hp, in_Wp;       // input padded dims
  int H, W;               // output real dims
  int nchunks, KC, taps, stride;
  int BN, MB, SA, SB;
  int halo;               // A rows per plane per stage
  int cout, flags;
  int tmem_cols;
  int desc_swap;          // debug: exchange LBO/SBO roles (hrnb_debug_set(0, 1))
  unsigned a_stage_bytes, b_stage_bytes;
};

constexpr int kBarBytes = 256;    // mbarriers + tmem ptr
constexpr int kBiasBytes = 1024;  // up to 256 fp32
constexpr int kSmemHeader = kBarBytes + kBiasBytes;
constexpr int kMaxSA = 4, kMaxSB = 8;

template <bool GATHER>
__global__ void __launch_bounds__(GATHER ? 320 : 192, 1) conv_tc_kernel(const ConvK k) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_a = full_a + kMaxSA;
  uint64_t* full_b = empty_a + kMaxSA;
  uint64_t* empty_b = full_b + kMaxSB;
  uint64_t* accum_full = empty_b + kMaxSB;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(accum_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem + kBarBytes);
  uint8_t* a_ring = smem + kSmemHeader;
  uint8_t* b_ring = a_ring + (size_t)k.SA * k.a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntile = blockIdx.y;
  const int mblk0 = blockIdx.x * k.MB;  // first 128-row block of this CTA

  if (threadIdx.x == 0) {
    for (int i = 0; i < k.SA; ++i) {
      mbar_init(&full_a[i], GATHER ? 128u : 1u);
      mbar_init(&empty_a[i], 1u);
    }
    for (int i = 0; i < k.SB; ++i) {
      mbar_init(&full_b[i], 1u);
      mbar_init(&empty_b[i], 1u);
    }
    mbar_init(accum_full, 1u);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, (uint32_t)k.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 6) {
    for (int i = threadIdx.x - 64; i < k.BN; i += 128) bias_s[i] = k.bias[ntile * k.BN + i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int steps_per_chunk = k.taps;
  const int rowsA = GATHER ? 128 * k.MB : k.halo;  // rows per plane in an A stage

  if (warp == 0) {
    // =============================== bulk-copy producer ===============================
    if (lane == 0) {
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
      const long long pstart =
          (long long)mblk0 * 128 - (k.taps == 9 ? (k.Wp + 1) : 0);  // first halo position (may be < 0: guard)
      const __nv_bfloat16* wsrc =
          k.wpk + (size_t)ntile * k.nchunks * k.taps * (size_t)(k.b_stage_bytes / 2);
      for (int c = 0; c < k.nchunks; ++c) {
        if (!GATHER) {
          mbar_wait(&empty_a[a_stage], a_phase ^ 1);
          mbar_arrive_expect_tx(&full_a[a_stage], k.a_stage_bytes);
          uint8_t* dst = a_ring + (size_t)a_stage * k.a_stage_bytes;
          const uint32_t plane_bytes = (uint32_t)k.halo * 16u;
          for (int j = 0; j < k.KC; ++j) {
            const __nv_bfloat16* src = k.in + ((long long)(c * k.KC + j) * k.in_ps + pstart) * 8;
            bulk_g2s(dst + (size_t)j * plane_bytes, src, plane_bytes, &full_a[a_stage]);
          }
          if (++a_stage == k.SA) { a_stage = 0; a_phase ^= 1; }
        }
        for (int t = 0; t < steps_per_chunk; ++t) {
          mbar_wait(&empty_b[b_stage], b_phase ^ 1);
          mbar_arrive_expect_tx(&full_b[b_stage], k.b_stage_bytes);
          bulk_g2s(b_ring + (size_t)b_stage * k.b_stage_bytes, wsrc, k.b_stage_bytes, &full_b[b_stage]);
          wsrc += k.b_stage_bytes / 2;
          if (++b_stage == k.SB) { b_stage = 0; b_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_m128((uint32_t)k.BN);
      const uint32_t a_lbo = (uint32_t)rowsA * 16u;
      const uint32_t b_lbo = (uint32_t)k.BN * 16u;
      const uint32_t a_ring_addr = smem_u32(a_ring);
      const uint32_t b_ring_addr = smem_u32(b_ring);
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
      uint32_t accumulate = 0;
      for (int c = 0; c < k.nchunks; ++c) {
        if (!GATHER) {
          mbar_wait(&full_a[a_stage], a_phase);
          tc_fence_after_sync();
        }
        for (int t = 0; t < steps_per_chunk; ++t) {
          if (GATHER) {
            mbar_wait(&full_a[a_stage], a_phase);
          }
          mbar_wait(&full_b[b_stage], b_phase);
          tc_fence_after_sync();
          const uint32_t shift = (!GATHER && k.taps == 9) ? (uint32_t)((t / 3) * k.Wp + (t % 3)) : 0u;
          const uint32_t a_base = a_ring_addr + (uint32_t)a_stage * k.a_stage_bytes + shift * 16u;
          const uint32_t b_base = b_ring_addr + (uint32_t)b_stage * k.b_stage_bytes;
          for (int mb = 0; mb < k.MB; ++mb) {
            uint32_t acc = accumulate;
            for (int j = 0; j < k.KC / 2; ++j) {
              const uint32_t a_addr = a_base + (uint32_t)mb * 2048u + (uint32_t)j * 2u * a_lbo;
              const uint32_t b_addr = b_base + (uint32_t)j * 2u * b_lbo;
              const uint64_t adesc = k.desc_swap ? make_kmajor_desc(a_addr, 128u, a_lbo) : make_kmajor_desc(a_addr, a_lbo, 128u);
              const uint64_t bdesc = k.desc_swap ? make_kmajor_desc(b_addr, 128u, b_lbo) : make_kmajor_desc(b_addr, b_lbo, 128u);
              umma_bf16_ss(tmem_base + (uint32_t)(mb * k.BN), adesc, bdesc, idesc, acc);
              acc = 1u;
            }
          }
          accumulate = 1u;
          umma_commit(&empty_b[b_stage]);
          if (++b_stage == k.SB) { b_stage = 0; b_phase ^= 1; }
          if (GATHER) {
            umma_commit(&empty_a[a_stage]);
            if (++a_stage == k.SA) { a_stage = 0; a_phase ^= 1; }
          }
        }
        if (!GATHER) {
          umma_commit(&empty_a[a_stage]);
          if (++a_stage == k.SA) { a_stage = 0; a_phase ^= 1; }
        }
      }
      umma_commit(accum_full);
    }
  } else if (warp < 6) {
    // =============================== epilogue ===============================
    mbar_wait(accum_full, 0);
    tc_fence_after_sync();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool relu = (k.flags & HRNB_CONV_RELU) != 0;
    const bool nchw = (k.flags & HRNB_CONV_OUT_NCHW) != 0;
    for (int mb = 0; mb < k.MB; ++mb) {
      const long long p = (long long)(mblk0 + mb) * 128 + q * 32 + lane;
      const bool valid = p < k.P;
      const int px = (int)(p % k.Wp);
      const int rowi = (int)(p / k.Wp);
      const int py = rowi % k.Hp;
      const int n = rowi / k.Hp;
      const bool real = valid && px > 0 && py > 0;
      for (int g = 0; g < k.BN / 16; ++g) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * k.BN + g * 16), v);
        tmem_ld_wait();
        const int cb = ntile * k.BN + g * 16;
        if (nchw) {
          if (real) {
            float* o = reinterpret_cast<float*>(k.out);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = cb + i;
              if (c < k.cout) {
                float x = __uint_as_float(v[i]) + bias_s[g * 16 + i];
                if (relu) x = fmaxf(x, 0.f);
                o[(((long long)n * k.cout + c) * k.H + (py - 1)) * k.W + (px - 1)] = x;
              }
            }
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[h * 8 + i]) + bias_s[g * 16 + h * 8 + i];
            const long long plane = (long long)(cb / 8 + h);
            if (k.res != nullptr && real) {
              const uint4 r = ldg_nc_v4(k.res + (plane * k.res_ps + p) * 8);
              x[0] += bf16_lo(r.x); x[1] += bf16_hi(r.x);
              x[2] += bf16_lo(r.y); x[3] += bf16_hi(r.y);
              x[4] += bf16_lo(r.z); x[5] += bf16_hi(r.z);
              x[6] += bf16_lo(r.w); x[7] += bf16_hi(r.w);
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
            }
            uint4 o;
            if (real) {
              o.x = pack_bf16x2(x[0], x[1]);
              o.y = pack_bf16x2(x[2], x[3]);
              o.z = pack_bf16x2(x[4], x[5]);
              o.w = pack_bf16x2(x[6], x[7]);
            } else {
              o = make_uint4(0u, 0u, 0u, 0u);  // keep the shared zero padding intact
            }
            if (valid) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(k.out) + (plane * k.out_ps + p) * 8) = o;
            }
          }
        }
      }
    }
  } else {
    // =============================== gather producers (GATHER only) ===============================
    if (GATHER) {
      const int g = threadIdx.x - 192;  // A row handled by this thread (per M block)
      long long pin[4];
      for (int mb = 0; mb < k.MB; ++mb) {
        const long long p = (long long)(mblk0 + mb) * 128 + g;
        const int px = (int)(p % k.Wp);
        const int rowi = (int)(p / k.Wp);
        const int py = rowi % k.Hp;
        const int n = rowi / k.Hp;
        const bool real = p < k.P && px > 0 && py > 0;
        // input position of tap (0,0) (3x3) for this output pixel; invalid rows read position -1.. (guard = 0)
        pin[mb] = real ? ((long long)n * k.in_Hp + (long long)(py - 1) * k.stride) * k.in_Wp +
                             (long long)(px - 1) * k.stride
                       : -1;
      }
      constexpr int LAG = 2;
      int it = 0;
      const int total = k.nchunks * k.taps;
      for (int c = 0; c < k.nchunks; ++c) {
        for (int t = 0; t < k.taps; ++t, ++it) {
          const int stage = it % k.SA;
          const int phase = (it / k.SA) & 1;
          mbar_wait(&empty_a[stage], phase ^ 1);
          const int r = (k.taps == 9) ? t / 3 : 1;
          const int s = (k.taps == 9) ? t % 3 : 1;
          const long long toff = (long long)r * k.in_Wp + s;
          uint8_t* dst = a_ring + (size_t)stage * k.a_stage_bytes;
          for (int mb = 0; mb < k.MB; ++mb) {
            const long long src_pos = pin[mb] < 0 ? -1 : pin[mb] + toff;
            for (int j = 0; j < k.KC; ++j) {
              cp_async16(dst + ((size_t)j * rowsA + mb * 128 + g) * 16,
                         k.in + ((long long)(c * k.KC + j) * k.in_ps + src_pos) * 8);
            }
          }
          cp_async_commit();
          if (it >= LAG) {
            cp_async_wait<LAG>();
            fence_proxy_async_smem();
            mbar_arrive(&full_a[(it - LAG) % k.SA]);
          }
        }
      }
      // drain
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int d = (total > LAG ? total - LAG : 0); d < total; ++d) mbar_arrive(&full_a[d % k.SA]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)k.tmem_cols);
  }
}

int g_debug[8] = {0, 0, 0, 0, 0, 0, 0, 0};

static int next_pow2_cols(int c) {
  int r = 32;
  while (r < c) r <<= 1;
  return r;
}

// Derive launch geometry; returns smem bytes or <0.
static long long derive(const hrnb_conv_params* p, ConvK* k) {
  if (!p || !p->in || !p->wpk || !p->bias || !p->out) return fail(HRNB_EINVAL, "conv: null pointer");
  const bool gather = (p->flags & HRNB_CONV_GATHER) != 0;
  if (p->taps != 1 && p->taps != 9) return fail(HRNB_EINVAL, "conv: taps must be 1 or 9");
  if (p->stride != 1 && p->stride != 2) return fail(HRNB_EINVAL, "conv: stride must be 1 or 2");
  if (p->stride == 2 && !gather) return fail(HRNB_EINVAL, "conv: stride 2 needs HRNB_CONV_GATHER");
  if (p->cin % 16 || p->cin <= 0) return fail(HRNB_EINVAL, "conv: cin must be a positive multiple of 16");
  if (p->KC <= 0 || (p->KC & 1) || (p->cin / 8) % p->KC) return fail(HRNB_EINVAL, "conv: KC must be even and divide cin/8");
  if (p->BN < 16 || p->BN > 256 || p->BN % 16) return fail(HRNB_EINVAL, "conv: BN must be a multiple of 16 in [16,256]");
  if (p->MB != 1 && p->MB != 2 && p->MB != 4) return fail(HRNB_EINVAL, "conv: MB must be 1, 2 or 4");
  if (p->MB * p->BN > 512) return fail(HRNB_EINVAL, "conv: MB*BN exceeds 512 TMEM columns");
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->cout <= 0) return fail(HRNB_EINVAL, "conv: bad geometry");
  if (p->in_H != p->H * p->stride || p->in_W != p->W * p->stride) return fail(HRNB_EINVAL, "conv: in_H/in_W must equal H*stride/W*stride");
  const bool nchw = (p->flags & HRNB_CONV_OUT_NCHW) != 0;
  if (!nchw && (p->cout % p->BN)) return fail(HRNB_EINVAL, "conv: PF8 output needs cout % BN == 0");
  k->in = (const __nv_bfloat16*)p->in;
  k->in_ps = p->in_ps;
  k->wpk = (const __nv_bfloat16*)p->wpk;
  k->bias = p->bias;
  k->res = (const __nv_bfloat16*)p->res;
  k->res_ps = p->res_ps;
  k->out = p->out;
  k->out_ps = p->out_ps;
  k->Hp = p->H + 1;
  k->Wp = p->W + 1;
  k->in_Hp = p->in_H + 1;
  k->in_Wp = p->in_W + 1;
  k->H = p->H;
  k->W = p->W;
  const long long P = (long long)p->N * k->Hp * k->Wp;
  if (P > 0x7fffffffLL) return fail(HRNB_EINVAL, "conv: too many positions");
  k->P = (int)P;
  k->KC = p->KC;
  k->nchunks = (p->cin / 8) / p->KC;
  k->taps = p->taps;
  k->stride = p->stride;
  k->BN = p->BN;
  k->MB = p->MB;
  k->cout = p->cout;
  k->flags = p->flags;
  k->tmem_cols = next_pow2_cols(p->MB * p->BN);
  k->desc_swap = g_debug[0];
  k->b_stage_bytes = (unsigned)(p->KC * p->BN * 16);
  if (gather) {
    k->halo = 128 * p->MB;
    k->a_stage_bytes = (unsigned)(p->KC * 128 * p->MB * 16);
    k->SA = 4;
  } else {
    k->halo = 128 * p->MB + (p->taps == 9 ? 2 * (k->Wp + 1) : 0);
    k->a_stage_bytes = (unsigned)(p->KC * k->halo * 16);
    k->SA = k->nchunks > 1 ? 2 : 1;
  }
  const long long limit = 200 * 1024;
  int SB = 6;
  const int total_b = k->nchunks * k->taps;
  if (SB > total_b) SB = total_b;
  while (SB > 1 && kSmemHeader + (long long)k->SA * k->a_stage_bytes + (long long)SB * k->b_stage_bytes > limit) --SB;
  if (!gather && k->SA == 2 &&
      kSmemHeader + (long long)k->SA * k->a_stage_bytes + (long long)SB * k->b_stage_bytes > limit) {
    k->SA = 1;
  }
  if (gather && kSmemHeader + (long long)k->SA * k->a_stage_bytes + (long long)SB * k->b_stage_bytes > limit) {
    k->SA = 3;  // the gather producer runs LAG = 2 stages ahead: 3 is the minimum ring depth
  }
  const long long smem = kSmemHeader + (long long)k->SA * k->a_stage_bytes + (long long)SB * k->b_stage_bytes;
  if (smem > 227 * 1024) return fail(HRNB_EINVAL, "conv: tile does not fit in shared memory");
  k->SB = SB;
  return smem;
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_debug_set(int key, int value) {
  if (key < 0 || key >= 8) return HRNB_EINVAL;
  g_debug[key] = value;
  return HRNB_OK;
}

extern "C" int64_t hrnb_conv_smem_bytes(const hrnb_conv_params* p) {
  ConvK k;
  return derive(p, &k);
}

extern "C" int hrnb_conv(const hrnb_conv_params* p, void* stream) {
  ConvK k;
  const long long smem = derive(p, &k);
  if (smem < 0) return (int)smem;
  const bool gather = (p->flags & HRNB_CONV_GATHER) != 0;
  const int mblocks = (k.P + 127) / 128;
  dim3 grid((mblocks + k.MB - 1) / k.MB, (p->cout + p->BN - 1) / p->BN);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev][gather ? 1 : 0]) {
    cudaError_t e = gather ? cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                           : cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "conv: cudaFuncSetAttribute");
    attr_set[dev][gather ? 1 : 0] = true;
  }
  if (gather)
    conv_tc_kernel<true><<<grid, 320, (size_t)smem, st>>>(k);
  else
    conv_tc_kernel<false><<<grid, 192, (size_t)smem, st>>>(k);
  count_launch();
  return check_launch("conv_tc_kernel");
}
