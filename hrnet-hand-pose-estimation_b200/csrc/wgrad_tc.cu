// Weight-gradient of the HRNet convolutions on tcgen05 / TMEM (sm_100a only).
//
// Replaces the autograd backward of nn.Conv2d w.r.t. its weight (the reference trains through
// loss.backward(), lib/core/function.py:101-106; conv layers lib/models/pose_hrnet.py:28-98,187-242,335-350,419-458).
//
// GEMM view per tap t:   dW_t[co, ci] = sum over positions p of  dY[p, co] * X[p + dpos_t, ci]
// i.e. M = cout, N = cin, K = positions.  In the PF8 layout a plane is [position][8 channels] with
// 16 bytes per position, which is exactly the SWIZZLE_NONE *MN-major* canonical UMMA operand layout
// (8 channels contiguous, K rows 16 bytes apart, 8-row K blocks 128 bytes apart, 8-channel MN blocks one plane
// apart), so both operands are streamed from HBM by 1-D bulk copies with no transposition, and the tap shift is
// again just a descriptor start offset (dpos_t * 16 bytes) into one halo copy of X.  Zero padding positions of dY
// contribute nothing, zero padding of X supplies the conv's implicit zeros.
//
// Work item (one CTA) = (128-channel cout tile, NT-channel cin tile, group of TG taps, K split); the accumulators of
// the TG taps sit side by side in TMEM (TG*NT <= 512 fp32 columns).  The CTA streams its position range through an
// S-stage smem ring (KP positions per stage), issues (KP/16)*TG MMAs per stage and finally adds its partial
// dW to global memory with coalesced fp32 reductions (red.global.add): lane = cout channel, so a warp covers 32
// consecutive floats of the [tap][cin][cout] gradient layout.
//
// Warp roles: warp 0 bulk-copy producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 epilogue.
#include <atomic>
#include "ptx.cuh"
#include "common.h"

namespace hrnb {

struct WgradK {
  const __nv_bfloat16* dy;
  long long dy_ps;
  const __nv_bfloat16* x;
  long long x_ps;
  float* dw;
  int P, cout, cin, dy_planes;
  int ntap;
  int tap_boff[9];   // lead + dpos of tap t (rows inside a B stage)
  int tap_id[9];     // index of tap t in the gradient layout
  int n_cot, NT, n_cit, TG, n_tg, KP, S, lead, haloB, ksplit, nchunks;
  unsigned a_stage_bytes, b_stage_bytes;
  int tmem_cols;
  // M-stacked mode (3x3 convs with cout <= 64): the M = 128 rows hold `stack` copies of dY shifted by s0, s0+1, ... positions
  // (copy i = rows [i*cpl*8, (i+1)*cpl*8)), so ONE MMA per kernel row r yields the taps (r, s0..s0+stack-1):
  //   dW[r,s][co,ci] = sum_q dY[q - s, co] * X[q + (r-1)*Wp - 1, ci]            (q = p + s)
  int stack, s0, cpl;   // stack == 0: plain mode
  // tap groups (one per CTA): taps [grp_t0[g], grp_t0[g] + grp_tn[g]) read the X tensor at x + grp_x_off[g] elements -
  // uniform groups of TG taps of one tensor, or one group per input phase of a stride-2 conv (hrnb_wgrad_params.tap_src)
  int grp_t0[9], grp_tn[9];
  long long grp_x_off[9];
};

constexpr int kWgThreads = 192;

// one stage (KP positions = ksteps K-steps of 16) for TN taps: per K step TN MMAs into TN accumulators, straight-line
template <int TN>
__device__ __forceinline__ void issue_chunk(uint32_t d_base, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo_stage, uint32_t b_hi,
                                            const uint32_t (&boff)[9], uint32_t NT, uint32_t idesc, uint32_t accumulate,
                                            int ksteps) {
  uint32_t acc = accumulate;
#pragma unroll 2
  for (int kk = 0; kk < ksteps; ++kk) {
    if (elect_one_sync()) {
      const uint64_t adesc = ((uint64_t)a_hi << 32) | a_lo;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        const uint64_t bdesc = ((uint64_t)b_hi << 32) | (b_lo_stage + boff[t]);
        umma_bf16_ss(d_base + (uint32_t)t * NT, adesc, bdesc, idesc, acc);
      }
    }
    acc = 1u;
    a_lo += 16u;         // 16 positions * 16 bytes, in 16-byte units
    b_lo_stage += 16u;
  }
}
constexpr int kWgMaxS = 6;
constexpr int kWgHeader = 256;

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgradK k) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kWgMaxS;
  uint64_t* tmem_full = empty + kWgMaxS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
  uint8_t* a_ring = smem + kWgHeader;
  uint8_t* b_ring = a_ring + (size_t)k.S * k.a_stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < k.S; ++i) {
      mbar_init(&full[i], 1u);
      mbar_init(&empty[i], 1u);
    }
    mbar_init(tmem_full, 1u);
    fence_mbar_init();
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 1) {
    asm volatile("griddepcontrol.wait;" ::: "memory");    // never hold TMEM while waiting for the previous kernel (see conv_tc.cu)
    tmem_alloc(tmem_ptr_smem, (uint32_t)k.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_enter();     // everything above (barriers, TMEM) overlaps the previous kernel's tail; global memory only from here on

  // work item -> (cout tile, cin tile, tap group, K split)
  int r = blockIdx.x;
  const int cot = r % k.n_cot; r /= k.n_cot;
  const int cit = r % k.n_cit; r /= k.n_cit;
  const int tg = r % k.n_tg;
  const int ks = r / k.n_tg;
  const int c0 = (int)((long long)ks * k.nchunks / k.ksplit);
  const int c1 = (int)((long long)(ks + 1) * k.nchunks / k.ksplit);
  const int t0 = k.grp_t0[tg];
  const int tn = k.grp_tn[tg];
  const __nv_bfloat16* const xsrc = k.x + k.grp_x_off[tg];
  const int mt = (k.dy_planes - cot * 16 < 16) ? (k.dy_planes - cot * 16) : 16;
  const int nplanes_b = k.NT / 8;

  if (c1 > c0) {
    if (warp == 0) {
      // =============================== producer ===============================
      int stage = 0, phase = 0;
      const uint32_t a_plane_bytes = (uint32_t)k.KP * 16u, b_plane_bytes = (uint32_t)k.haloB * 16u;
      const uint32_t tx = (uint32_t)mt * a_plane_bytes + (uint32_t)nplanes_b * b_plane_bytes;
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[stage], tx);
          uint8_t* a_dst = a_ring + (size_t)stage * k.a_stage_bytes;
          uint8_t* b_dst = b_ring + (size_t)stage * k.b_stage_bytes;
          const long long pa = (long long)c * k.KP;
          if (k.stack > 0) {
            for (int j = 0; j < mt; ++j)      // plane j of the stage = channel plane j % cpl of copy j / cpl
              bulk_g2s(a_dst + (size_t)j * a_plane_bytes,
                       k.dy + ((long long)(j % k.cpl) * k.dy_ps + pa - (k.s0 + j / k.cpl)) * 8, a_plane_bytes, &full[stage]);
          } else {
            for (int j = 0; j < mt; ++j)
              bulk_g2s(a_dst + (size_t)j * a_plane_bytes, k.dy + ((long long)(cot * 16 + j) * k.dy_ps + pa) * 8,
                       a_plane_bytes, &full[stage]);
          }
          for (int j = 0; j < nplanes_b; ++j)
            bulk_g2s(b_dst + (size_t)j * b_plane_bytes,
                     xsrc + ((long long)(cit * nplanes_b + j) * k.x_ps + pa - k.lead) * 8, b_plane_bytes, &full[stage]);
        }
        __syncwarp();
        if (++stage == k.S) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer ===============================
      // both operands MN-major, SWIZZLE_NONE: LBO = 128 B (next 8 positions), SBO = one plane of the stage
      const uint32_t idesc = make_idesc_bf16_m128((uint32_t)k.NT) | (1u << 15) | (1u << 16);
      const uint32_t a_hi = (uint32_t)k.KP | (1u << 14);       // SBO = KP*16 bytes (in 16-byte units), version 1
      const uint32_t b_hi = (uint32_t)k.haloB | (1u << 14);
      const uint32_t lbo = 8u << 16;                           // 128 bytes
      const uint32_t a_lo_ring = (smem_u32(a_ring) >> 4) | lbo;
      const uint32_t b_lo_ring = (smem_u32(b_ring) >> 4) | lbo;
      const uint32_t a_stage16 = k.a_stage_bytes >> 4, b_stage16 = k.b_stage_bytes >> 4;
      const int ksteps = k.KP / 16;
      int stage = 0, phase = 0;
      uint32_t accumulate = 0;
      // tap offsets of this CTA's group in uniform registers; the per-K-step issue is straight-line code (TN MMAs
      // inside one elect block): with a runtime tap loop the issuing thread, not the tensor pipe, set the pace
      // [measured: 74 us -> see profiles/ for the 32-channel 64x64 layers]
      uint32_t boff[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) boff[t] = (uint32_t)k.tap_boff[(t0 + t < 9) ? t0 + t : 8];
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after_sync();
        const uint32_t a_lo = a_lo_ring + (uint32_t)stage * a_stage16;
        const uint32_t b_lo_stage = b_lo_ring + (uint32_t)stage * b_stage16;
        switch (tn) {
          case 1: issue_chunk<1>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 2: issue_chunk<2>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 3: issue_chunk<3>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 4: issue_chunk<4>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 5: issue_chunk<5>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 6: issue_chunk<6>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 7: issue_chunk<7>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          case 8: issue_chunk<8>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
          default: issue_chunk<9>(tmem_base, a_lo, a_hi, b_lo_stage, b_hi, boff, (uint32_t)k.NT, idesc, accumulate, ksteps); break;
        }
        accumulate = 1u;
        __syncwarp();
        if (elect_one_sync()) umma_commit(&empty[stage]);
        if (++stage == k.S) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
      if (elect_one_sync()) umma_commit(tmem_full);
      __syncwarp();
    } else {
      // =============================== epilogue ===============================
      const int q = warp & 3;
      int co = cot * 128 + q * 32 + lane;
      int tap_add = 0;
      bool row_ok = co < k.cout;
      if (k.stack > 0) {                       // row -> (copy, channel)
        const int row = q * 32 + lane, per = k.cpl * 8;
        const int copy = row / per;
        co = row - copy * per;
        tap_add = k.s0 + copy;
        row_ok = copy < k.stack && co < k.cout;
      }
      mbar_wait(tmem_full, 0);
      tc_fence_after_sync();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
      const int groups = k.NT / 16;
      for (int t = 0; t < tn; ++t) {
        const long long row0 = (long long)(k.tap_id[t0 + t] + tap_add) * k.cin + (long long)cit * k.NT;
        for (int g = 0; g < groups; ++g) {
          uint32_t v[16];
          tmem_ld16(t_base + (uint32_t)(t * k.NT + g * 16), v);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = k.dw + (row0 + g * 16) * k.cout + co;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + (long long)i * k.cout, __uint_as_float(v[i]));
          }
        }
      }
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

// stacked[0] = copies of dY stacked in M (0 = plain mode), stacked[1] = first column shift s0 of this launch
static long long derive_wgrad(const hrnb_wgrad_params* p, WgradK* k, int* grid, const int* stacked = nullptr) {
  if (!p || !p->dy || !p->x || !p->dw) return fail(HRNB_EINVAL, "wgrad: null pointer");
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->cout <= 0 || p->cin <= 0) return fail(HRNB_EINVAL, "wgrad: bad geometry");
  if (p->cin % 16) return fail(HRNB_EINVAL, "wgrad: cin must be a multiple of 16");
  if (p->ntap < 1 || p->ntap > 9) return fail(HRNB_EINVAL, "wgrad: ntap must be in [1, 9]");
  const int Wp = p->W + 1, Hp = p->H + 1;
  const long long P = (long long)p->N * Hp * Wp;
  if (P > 0x7fffffffLL) return fail(HRNB_EINVAL, "wgrad: too many positions");
  int lo = 0, hi = 0;
  for (int t = 0; t < p->ntap; ++t) {
    if (p->tap_dpos[t] < lo) lo = p->tap_dpos[t];
    if (p->tap_dpos[t] > hi) hi = p->tap_dpos[t];
    if (p->tap_id[t] < 0) return fail(HRNB_EINVAL, "wgrad: negative tap_id");
  }
  if (-lo > HRNB_GUARD_LEAD(Wp) || hi > HRNB_GUARD_LEAD(Wp)) return fail(HRNB_EINVAL, "wgrad: tap offset exceeds the PF8 guard band");
  k->dy = (const __nv_bfloat16*)p->dy;
  k->dy_ps = p->dy_ps;
  k->x = (const __nv_bfloat16*)p->x;
  k->x_ps = p->x_ps;
  k->dw = p->dw;
  k->P = (int)P;
  k->cout = p->cout;
  k->cin = p->cin;
  k->dy_planes = (p->cout + 7) / 8;
  k->ntap = p->ntap;
  k->lead = -lo;
  k->stack = k->s0 = k->cpl = 0;
  for (int t = 0; t < 9; ++t) {
    k->tap_boff[t] = t < p->ntap ? k->lead + p->tap_dpos[t] : 0;
    k->tap_id[t] = t < p->ntap ? p->tap_id[t] : 0;
  }
  int extra_k = 0;
  if (stacked && stacked[0] > 0) {
    // "taps" of this launch are the three kernel rows: B offset (r-1)*Wp - 1, gradient index r*3 (+ column shift in the epilogue)
    k->stack = stacked[0];
    k->s0 = stacked[1];
    k->cpl = (p->cout + 7) / 8;
    k->dy_planes = k->stack * k->cpl;
    k->ntap = 3;
    lo = -(Wp + 1);
    hi = Wp - 1;
    k->lead = -lo;
    for (int r = 0; r < 3; ++r) {
      k->tap_boff[r] = k->lead + (r - 1) * Wp - 1;
      k->tap_id[r] = r * 3;
    }
    extra_k = 2;      // q = p + s runs two positions past P
  }
  k->n_cot = (k->dy_planes + 15) / 16;
  // cin tile: the widest multiple of 16 dividing cin (<= 256 for single-tap layers, <= 128 otherwise so that
  // several taps share one pass over dY)
  int NT = p->NT;
  if (NT <= 0) {
    const int cap = p->ntap == 1 ? 256 : 128;
    for (NT = cap; NT >= 16; NT -= 16)
      if (p->cin % NT == 0) break;
  }
  if (NT < 16 || NT > 256 || NT % 16 || p->cin % NT) return fail(HRNB_EINVAL, "wgrad: NT must be a multiple of 16 (<= 256) dividing cin");
  k->NT = NT;
  k->n_cit = p->cin / NT;
  // taps per CTA: three (one kernel row) for 3x3 convs - balanced groups, and fewer K splits (= fewer fp32 reductions
  // into dW) than with all nine taps per CTA [measured on B200, tools/wgrad_bench.py: 43 -> 37 us (32 ch @ 64x64),
  // 29.5 -> 20.5 us (64 ch @ 32x32), 68.6 -> 47 us (64 ch @ 64x64) at batch 64]
  int TG = p->TG > 0 ? p->TG : (p->ntap == 9 ? 3 : 512 / NT);
  if (k->stack > 0) TG = 3;
  if (TG * NT > 512) TG = 512 / NT;
  if (TG > k->ntap) TG = k->ntap;
  if (TG < 1 || TG * NT > 512) return fail(HRNB_EINVAL, "wgrad: TG*NT exceeds 512 TMEM columns");
  k->TG = TG;
  k->n_tg = (k->ntap + TG - 1) / TG;
  for (int g = 0; g < 9; ++g) {
    k->grp_t0[g] = g * TG;
    k->grp_tn[g] = (k->ntap - g * TG < TG) ? (k->ntap - g * TG > 0 ? k->ntap - g * TG : 0) : TG;
    k->grp_x_off[g] = 0;
  }
  if (p->x_src_stride != 0 && k->stack == 0) {
    // one tap group per source tensor: consecutive taps with the same tap_src
    int ng = 0, maxn = 0;
    for (int t = 0; t < p->ntap;) {
      int e = t;
      while (e < p->ntap && p->tap_src[e] == p->tap_src[t]) ++e;
      if (ng >= 9 || p->tap_src[t] < 0 || p->tap_src[t] > 3) return fail(HRNB_EINVAL, "wgrad: bad tap_src");
      for (int g = 0; g < ng; ++g)
        if (k->grp_x_off[g] == (long long)p->tap_src[t] * p->x_src_stride) return fail(HRNB_EINVAL, "wgrad: taps of one source must be consecutive");
      k->grp_t0[ng] = t;
      k->grp_tn[ng] = e - t;
      k->grp_x_off[ng] = (long long)p->tap_src[t] * p->x_src_stride;
      if (e - t > maxn) maxn = e - t;
      ++ng;
      t = e;
    }
    if (maxn * NT > 512) return fail(HRNB_EINVAL, "wgrad: taps of one source x NT exceed 512 TMEM columns (pass a smaller NT)");
    k->TG = maxn;
    k->n_tg = ng;
    TG = maxn;
  }
  int cols = 32;
  while (cols < TG * NT) cols <<= 1;
  k->tmem_cols = cols;
  const int mt_max = k->dy_planes < 16 ? k->dy_planes : 16;
  const int halo_extra = hi - lo;
  auto stage_bytes = [&](int kp, unsigned* a, unsigned* b) {
    *a = (unsigned)(mt_max * kp * 16);
    *b = (unsigned)((NT / 8) * (kp + halo_extra) * 16);
  };
  const long long budget = 200 * 1024;
  int KP = p->KP, S = 0;
  unsigned ab = 0, bb = 0;
  if (KP > 0) {
    if (KP % 16 || KP > 512) return fail(HRNB_EINVAL, "wgrad: KP must be a multiple of 16, <= 512");
    stage_bytes(KP, &ab, &bb);
    S = (int)((budget - kWgHeader - 16 * KP * 16) / (ab + bb));
  } else {
    for (KP = 256; KP >= 32; KP >>= 1) {
      stage_bytes(KP, &ab, &bb);
      S = (int)((budget - kWgHeader - 16 * KP * 16) / (ab + bb));
      if (S >= 3) break;
    }
    if (KP < 32) { KP = 32; stage_bytes(KP, &ab, &bb); S = (int)((budget - kWgHeader - 16 * KP * 16) / (ab + bb)); }
  }
  if (S < 2) return fail(HRNB_EINVAL, "wgrad: stage does not fit in shared memory");
  if (S > kWgMaxS) S = kWgMaxS;
  k->KP = KP;
  k->S = S;
  k->haloB = KP + halo_extra;
  k->a_stage_bytes = ab;
  k->b_stage_bytes = bb;
  k->nchunks = (int)((P + extra_k + KP - 1) / KP);
  if (HRNB_GUARD_TAIL(Wp) < KP + hi + extra_k) return fail(HRNB_EINVAL, "wgrad: KP exceeds the PF8 tail guard");
  const int base_items = k->n_cot * k->n_cit * k->n_tg;
  int ksplit = p->ksplit;
  if (ksplit <= 0) {
    ksplit = 148 / base_items;       // one wave of CTAs: rounding up to 152 items costs a whole second wave
    if (ksplit < 1) ksplit = 1;
  }
  if (ksplit > k->nchunks) ksplit = k->nchunks;
  k->ksplit = ksplit;
  *grid = base_items * ksplit;
  // the MMA always reads 16 A planes (M = 128): planes beyond `mt` fall into the following bytes of the ring, so the
  // allocation keeps one full 16-plane stage of slack after the last A stage (contents irrelevant: rows >= cout are
  // never stored)
  long long smem = kWgHeader + (long long)S * (ab + bb);
  const long long need = kWgHeader + (long long)(S - 1) * ab + 16LL * KP * 16;
  if (smem < need) smem = need;
  if (smem > 227 * 1024) return fail(HRNB_EINVAL, "wgrad: shared memory budget exceeded");
  return smem;
}

int bind_hang_buffer_wgrad() {
  unsigned long long* d = hang_buffer_device_ptr();
  if (d == nullptr) return fail(HRNB_ECUDA, "hang buffer: cudaHostAlloc failed");
  cudaError_t e = cudaMemcpyToSymbol(g_hang_buf, &d, sizeof(d));
  return e == cudaSuccess ? HRNB_OK : fail_cuda(e, "hang buffer: cudaMemcpyToSymbol");
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int64_t hrnb_wgrad_smem_bytes(const hrnb_wgrad_params* p) {
  WgradK k;
  int grid = 0;
  return derive_wgrad(p, &k, &grid);
}

// standard 3x3 stride-1 tap table (tap t = (r,s) reads p + (r-1)*Wp + (s-1), gradient slot t)?
static bool is_standard_3x3(const hrnb_wgrad_params* p) {
  if (!p || p->ntap != 9) return false;
  const int Wp = p->W + 1;
  for (int t = 0; t < 9; ++t)
    if (p->tap_dpos[t] != (t / 3 - 1) * Wp + (t % 3 - 1) || p->tap_id[t] != t) return false;
  return true;
}

static int launch_wgrad(const WgradK& k, int grid, long long smem, cudaStream_t st);

extern "C" int hrnb_wgrad(const hrnb_wgrad_params* p, void* stream) {
  WgradK k;
  int grid = 0;
  // M-stacked form for thin 3x3 layers (cout <= 40: all three column shifts fit in M = 128): 3x fewer MMAs and a single pass
  // over dY / X [measured, batch 64: 36.9 -> 22.6 us for 32 ch @ 64x64; with only two copies per launch (cout 48 / 64) the
  // second launch costs more than it saves: 20.5 -> 31.7 us, so those keep the plain form]
  if (is_standard_3x3(p) && p->cout <= 40 && p->TG == 0 && hrnb::g_debug[5] == 0) {
    const int cpl = (p->cout + 7) / 8;
    const int per = 16 / cpl >= 3 ? 3 : 16 / cpl;        // copies per launch
    for (int s0 = 0; s0 < 3; s0 += per) {
      const int st2[2] = {s0 + per <= 3 ? per : 3 - s0, s0};
      const long long smem = derive_wgrad(p, &k, &grid, st2);
      if (smem < 0) return (int)smem;
      const int rc = launch_wgrad(k, grid, smem, (cudaStream_t)stream);
      if (rc) return rc;
    }
    return HRNB_OK;
  }
  const long long smem = derive_wgrad(p, &k, &grid);
  if (smem < 0) return (int)smem;
  return launch_wgrad(k, grid, smem, (cudaStream_t)stream);
}

static int launch_wgrad(const WgradK& k, int grid, long long smem, cudaStream_t stream) {
  static std::atomic<unsigned char> attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "wgrad: cudaFuncSetAttribute");
    attr_set[dev].store(1, std::memory_order_release);
  }
  if (smem < kTmemExclusiveSmem && hrnb::g_debug[6] == 0) smem = kTmemExclusiveSmem;   // one TMEM-holding CTA per SM (common.h)
  launch_pdl(wgrad_tc_kernel, dim3((unsigned)grid), dim3(kWgThreads), (size_t)smem, (cudaStream_t)stream, k);
  count_launch();
  return check_launch("wgrad_tc_kernel");
}
