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
// p + (r-1)*Wp + (s-1).  One tile therefore loads a single contiguous halo range of every input plane with
// 1-D bulk-async copies (TMA engine) and feeds all nine taps to the tensor core from that one smem copy by
// offsetting the UMMA shared-memory descriptor by `shift*16` bytes (SWIZZLE_NONE K-major canonical layout:
// rows are 16 bytes apart, so any row shift is a legal descriptor start).  Stride-2 convs cannot be
// expressed as a flat shift; they use the GATHER variant whose producer warps build each tap's A tile with
// 16-byte cp.async.
//
// The kernel is PERSISTENT: grid = min(tiles, SMs x CTAs/SM); every CTA walks tiles blockIdx.x, +gridDim.x, ...
// A tile = MB x 128 output positions x BN output channels.  Three decoupled pipelines, all mbarrier based:
//   smem rings   (producer  -> MMA)      : A halo chunks [SA stages], packed weight tiles [SB stages]
//   TMEM stages  (MMA       -> epilogue) : two accumulator buffers of MB*BN fp32 columns
// so the loads of tile i+1 and the epilogue of tile i-1 overlap the MMAs of tile i.
//
// Warp roles (16 epilogue warps in the default build, HRNB_EPI_WARPS):
//   warp 0       : bulk-copy producer (A halo chunks, packed weight tiles)           [1 elected lane]
//   warp 1       : TMEM allocator + tcgen05.mma issuer                                [1 elected lane]
//   warps 2..17  : epilogue  (tcgen05.ld -> +bias -> +residual (+fuse sources) -> ReLU -> bf16 -> coalesced 16 B stores);
//                  residuals are requested before the accumulator is waited for
//   warp 18      : second tcgen05.mma issuer on alternate tiles (k.dual: resident-weight flat-shift launches), else idle
//   warps 19..22 : (GATHER only) cp.async gather producers, one thread per A row
#include <atomic>
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
  int nsrc;               // 1, or 4 = stride-2 conv over a phase-split input (4 half-resolution PF8 tensors)
  long long src_stride;   // elements between consecutive phase tensors of the input
  long long out_phase_stride;  // elements between phase tensors of the output (HRNB_CONV_OUT_PHASES), else 0
  __nv_bfloat16* out2;    // optional second, phase-split copy of the output (consumed by stride-2 convs)
  long long out2_ps, out2_phase_stride;
  int oHp2, oWp2;         // padded dims of the output phase grid
  int lead;               // halo rows in front of the tile's first position
  int lag;                // gather producer: stages issued ahead of the one being published (SA - 2)
  int halo;               // A rows per plane per stage
  int cout, flags;
  int tmem_cols;
  int n_tiles, num_tiles; // N tiles, total tiles (m groups x N tiles)
  int nbias;              // n_tiles * BN
  long long* trace;       // debug: per-role timeline of CTA 0 (hrnb_debug_trace)
  int dbg;                // debug bitmask (hrnb_debug_set(3, m)): 1 no residual loads, 2 no stores, 4 no TMEM loads
  unsigned a_stage_bytes, b_stage_bytes;
  int tap_off[9];         // flat-shift path: row offset of tap t inside an A stage (source * KC * halo + lead + dpos)
  float* stats_sums;      // STATS variant: per-channel (sum, sum of squares) of the conv output over the real positions
  float* stats_ws;        // STATS variant: ticket counter + per-CTA partial sums (hrnb_conv_stats_ws_floats)
  // LEAN epilogue: exact division by multiplication for n < 2^31, d >= 2: n / d == __umulhi(n, m) >> s (see fast_magic)
  unsigned mWp, sWp, mHp, sHp, mNt, sNt;
  // fuse-layer sum in the epilogue: extra PF8 sources read with nearest up-sampling (include/hrnb.h: nfuse)
  int nfuse;
  int fuse_shift[3];
  const __nv_bfloat16* fuse_src[3];
  long long fuse_ps[3];
  int in_up_shift;        // gather 1x1: input read at (y >> s, x >> s)
  int wres;               // weights resident in shared memory for the whole CTA (one K chunk, one N tile)
  int dual;               // two MMA-issuing warps on alternate tiles (resident-weight flat-shift launches, SA >= 3)
  int twin;               // host only: launch the 8-epilogue-warp instantiation with two CTAs per SM
  // grouped launch (hrnb_conv_params.ngroup): tile -> (M group, conv g, N tile); per conv: taps, tap offsets, weights, output
  int ngroup, n_tiles_all;
  int grp_taps[4];
  int grp_tap_off[4][4];
  const __nv_bfloat16* grp_wpk[4];
  unsigned grp_b_bytes[4];
  long long grp_out_stride, grp_res_stride;
};

// floor(n / d) for n < 2^31 and the divisor behind (m, s): m = ceil(2^(31 + c) / d), s = c - 1, c = ceil(log2 d) >= 1.
// Exact: n * m / 2^(31+c) = n/d + n*e/(d * 2^(31+c)) with e = m*d - 2^(31+c) < d <= 2^c, and n*e < 2^(31+c).
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned m, unsigned s) { return __umulhi(n, m) >> s; }

// debug timeline: slot = role*64 + 2*tile_iter + {0,1}; written by one lane of CTA 0 only when tracing is on
#define HRNB_TRACE(role, iter, ev)                                                                   \
  do {                                                                                                \
    if (k.trace != nullptr && blockIdx.x == 0 && lane == 0 && (iter) < 16)                            \
      k.trace[(role) * 64 + 2 * (iter) + (ev)] = clock64();                                           \
  } while (0)

constexpr int kBarBytes = 384;    // mbarriers (2*kMaxSA + 2*kMaxSB + 4) + tmem ptr + the STATS ticket at byte 320
constexpr int kBiasBytes = 3072;  // up to 768 fp32 (whole padded bias vector)
constexpr int kSmemHeader = kBarBytes + kBiasBytes;
constexpr int kMaxSA = 8, kMaxSB = 8;   // kMaxSB: also the most K chunks a resident weight slab may have

#ifndef HRNB_EPI_WARPS
#define HRNB_EPI_WARPS 16
#endif
constexpr int kEpiWarps = HRNB_EPI_WARPS;          // warps per CTA draining TMEM (a multiple of 4: k per TMEM lane quarter)
constexpr int kCtasPerSm = kEpiWarps <= 8 ? 2 : 1;  // 8 epilogue warps leave registers for two co-resident CTAs
constexpr int kThreadsFS = 96 + 32 * kEpiWarps;    // producer + MMA + epilogue + second MMA issuer (dual-issue launches)
constexpr int kThreadsGather = kThreadsFS + 128;   // + gather producers

// KSTEPS = KC/2 (K=16 MMA steps per tap and chunk) is a template parameter so that the MMA issue loop is straight-line code:
// the issuing thread can only run about one MMA ahead of the tensor pipe, so every scalar instruction and branch between two
// tcgen05.mma is tensor-pipe idle time for the thin (N = 32 / 64) layers [measured: 106 cycles/MMA with a runtime loop].
//
// STATS (training path, BN == cout <= 64): the epilogue also accumulates the BatchNorm batch statistics of the tensor it
// writes - sum and sum of squares per output channel over the real positions - so that the separate pass over the conv
// output (bn_stats_kernel) disappears.  Every epilogue thread owns one TMEM lane (row) and, because BN/16 divides the four
// warps of a lane quarter, ONE fixed 16-column group for all its tiles: 32 running sums live in registers for the whole
// persistent loop (two FMAs per element), are reduced once per CTA (transposed butterfly over the lanes, warps in order)
// and stored as the CTA's partial; the last CTA to arrive (ticket) adds the partials in CTA order.  Tile -> CTA mapping
// and every summation order are static, so the statistics are bit-reproducible run to run.
constexpr int kStatsSlot = 128;   // floats per CTA partial: [column group (<= 4)][sum x16 | sumsq x16]
constexpr int kStatsMaxCtas = 320;
// LEAN (flat-shift path, PF8 output, no phase-split output copy - i.e. almost every launch of the network): a specialised
// epilogue.  The general one below spends ~1,470 instructions per warp and tile on the thin layers (SASS: 8 runtime integer
// divisions per tile for the row geometry, phase-offset arithmetic and three store variants per 16-column group, cursor
// bookkeeping of the residual ring); with 16 epilogue warps on 4 schedulers that is ~5,900 issue cycles per tile against
// ~2,900 tensor-pipe cycles, and conv_bench's debug masks showed the 32-channel 64x64 layers bound by exactly that - not by
// HBM (no operand loads at all: 62 vs 66 us) and not by TMEM reads.  The lean form: row geometry by multiply-high (fast_div),
// per-warp item table (TMEM column, output / residual / bias offsets) hoisted out of the persistent loop, at most four items
// per warp and tile fully unrolled with all residual loads issued before the accumulator wait, packed fp32x2 adds.
// PH (lean flat-shift launches whose output also / only exists in the phase-split form a following stride-2 conv reads):
// kept out of the plain variant - the thin layers are bound by the epilogue's issue slots, every instruction there counts
// GRP (grouped launch, hrnb_conv_params.ngroup > 1): its per-tile conv lookup likewise lives in its own instantiation - inside
// the plain variant it cost 2.8 % (batch 256) / 7 % (batch 64) of the whole inference pass [in-trip A/B of library builds]
// EW (epilogue warps): 16, or 8 in the TWO-CTAS-PER-SM instantiation of the plain lean variant used by the thin resident-weight
// layers at large batch (hrnb_conv: `twin`): those layers are bound by HBM latency per SM (one halo + residual stream per CTA),
// two co-resident CTAs keep twice the loads in flight [libhrnb_epi8.so experiment: -16 % on the 32-channel 64x64 layers, while
// every HBM-bound layer with many epilogue items LOSES with 8 warps - hence per launch, not per build]
template <bool GATHER, bool NCHW, int KSTEPS, bool STATS = false, bool LEAN = false, bool PH = false, bool GRP = false, int EW = kEpiWarps>
__global__ void __launch_bounds__(GATHER ? kThreadsGather : 96 + 32 * EW, GATHER ? 1 : (EW <= 8 ? 2 : 1)) conv_tc_kernel(const ConvK k) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_a = full_a + kMaxSA;
  uint64_t* full_b = empty_a + kMaxSA;
  uint64_t* empty_b = full_b + kMaxSB;
  uint64_t* tmem_full = empty_b + kMaxSB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + kBarBytes);
  uint8_t* a_ring = smem + kSmemHeader;
  uint8_t* b_ring = a_ring + (size_t)k.SA * k.a_stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // Programmatic dependent launch: let the next kernel of the stream start its prologue now; this kernel's own
  // prologue (barrier init, TMEM allocation, bias staging - weights only) overlaps the predecessor's tail.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int i = 0; i < k.SA; ++i) {
      mbar_init(&full_a[i], GATHER ? 128u : 1u);
      mbar_init(&empty_a[i], 1u);
    }
    for (int i = 0; i < k.SB; ++i) {
      mbar_init(&full_b[i], 1u);
      mbar_init(&empty_b[i], 1u);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1u);
      mbar_init(&tmem_empty[i], (uint32_t)EW);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    // TMEM must NOT be held while waiting for the previous kernel: a programmatically launched dependent that
    // allocates first and then waits can starve a co-resident CTA of the kernel it depends on (CTA A holds 256
    // columns, CTA B of the next kernel blocks on 512, CTA C of the kernel after that gets the free 256 and waits for B
    // -> B never gets 512) [observed as rare mbarrier time-outs with three co-resident conv CTAs / several streams].
    asm volatile("griddepcontrol.wait;" ::: "memory");
    tmem_alloc(tmem_ptr_smem, (uint32_t)k.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 2 + EW) {
    for (int i = threadIdx.x - 64; i < k.nbias; i += 32 * EW) bias_s[i] = k.bias[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Activations written by earlier kernels of the stream may only be touched after griddepcontrol.wait; each role
  // issues it itself so that the producer can fetch the first WEIGHT stage (never written by a kernel) before it.

  const int rowsA = GATHER ? 128 * k.MB : k.halo;  // rows per plane in an A stage
  const int acc_cols = k.MB * k.BN;                // fp32 columns of one accumulator stage

  if (warp == 0) {
    // =============================== bulk-copy producer ===============================
    // executed by the whole warp (uniform operands); one elected lane issues the copies
    {
      int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0;
      bool b_prefetched = false;
      // RESIDENT WEIGHTS: a layer whose whole weight tensor is one stage (one K chunk, one N tile - the 32- and 64-channel
      // 3x3 layers) loads it once per CTA instead of once per tile: -30 % shared-memory fill traffic on the thinnest layers
      // and one ring stage instead of three, which is what lets two CTAs share an SM there (hrnb_conv: per_sm)
      const bool wres = k.wres != 0;
      if ((int)blockIdx.x < k.num_tiles && !(k.dbg & 16)) {   // first weight stage of the first tile, ahead of the wait
        const int na0 = (int)blockIdx.x % k.n_tiles_all;
        const int g0 = GRP ? na0 / k.n_tiles : 0;
        const int nt0 = na0 - g0 * k.n_tiles;
        const unsigned bytes0 = GRP ? k.grp_b_bytes[g0] : k.b_stage_bytes;
        const __nv_bfloat16* w0 = GRP ? k.grp_wpk[g0] : k.wpk;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_b[0], bytes0);
          bulk_g2s(b_ring, w0 + (size_t)nt0 * k.nchunks * (bytes0 / 2), bytes0, &full_b[0]);
        }
        __syncwarp();
        b_prefetched = true;
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int tile = blockIdx.x; tile < k.num_tiles; tile += gridDim.x) {
        const int mg = tile / k.n_tiles_all, nall = tile - mg * k.n_tiles_all;
        const int grp = GRP ? nall / k.n_tiles : 0;
        const int ntile = nall - grp * k.n_tiles;
        const long long pstart = (long long)mg * k.MB * 128 - k.lead;  // first halo position (may be < 0: guard band)
        const unsigned bbytes = GRP ? k.grp_b_bytes[grp] : k.b_stage_bytes;
        const size_t b_elems = bbytes / 2;
        const __nv_bfloat16* wsrc = (GRP ? k.grp_wpk[grp] : k.wpk) + (size_t)ntile * k.nchunks * b_elems;
        const int pit = (tile - (int)blockIdx.x) / (int)gridDim.x;
        HRNB_TRACE(0, pit, 0);
        for (int c = 0; c < k.nchunks; ++c) {
          if (!GATHER) {
            mbar_wait(&empty_a[a_stage], a_phase ^ 1);
            if (k.dbg & 8) {
              if (elect_one_sync()) mbar_arrive(&full_a[a_stage]);
            } else if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full_a[a_stage], k.a_stage_bytes);
              uint8_t* dst = a_ring + (size_t)a_stage * k.a_stage_bytes;
              const uint32_t plane_bytes = (uint32_t)k.halo * 16u;
              for (int sidx = 0; sidx < k.nsrc; ++sidx) {       // stage layout [source][plane][row]
                for (int j = 0; j < k.KC; ++j) {
                  const __nv_bfloat16* src =
                      k.in + sidx * k.src_stride + ((long long)(c * k.KC + j) * k.in_ps + pstart) * 8;
                  bulk_g2s(dst + (size_t)(sidx * k.KC + j) * plane_bytes, src, plane_bytes, &full_a[a_stage]);
                }
              }
            }
            __syncwarp();
            if (++a_stage == k.SA) { a_stage = 0; a_phase ^= 1; }
          }
          // one weight stage per K chunk: all taps of the chunk in a single bulk copy (one hand-off per chunk)
          if (wres) b_stage = c;     // resident slab: chunk c lives in stage c for the whole CTA
          if (b_prefetched) {
            // Stage 0 of the first tile is already in flight - and there is NOTHING to wait for: waiting here for the
            // "previous" phase of empty_b[0] (parity 1) is only a no-op while that barrier is still in phase 0.  The MMA
            // warp can consume the prefetched stage and commit empty_b[0] before this warp gets here (it only needs the
            // A stage issued a few instructions above; an instruction-cache miss of this warp under multi-kernel
            // co-residency is enough), the barrier flips to phase 1 and the wait then blocks until the stage is released
            // a SECOND time - never, in a CTA with one tile and <= SB chunks.  That was the device-side mbarrier
            // time-out of round 1 (programmatic dependent launch / extra streams; profiles/r1_hang_records_*.txt: only
            // producer warps of 1-tile conv CTAs stuck on empty_b[0], parity 1).
            b_prefetched = false;
          } else if (!wres || tile == (int)blockIdx.x) {
            mbar_wait(&empty_b[b_stage], b_phase ^ 1);
            if (k.dbg & 16) {
              if (elect_one_sync()) mbar_arrive(&full_b[b_stage]);
            } else if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full_b[b_stage], bbytes);
              bulk_g2s(b_ring + (size_t)b_stage * k.b_stage_bytes, wsrc, bbytes, &full_b[b_stage]);
            }
          }
          __syncwarp();
          wsrc += b_elems;
          if (!wres && ++b_stage == k.SB) { b_stage = 0; b_phase ^= 1; }
          if (c == k.nchunks - 1) HRNB_TRACE(0, pit, 1);
        }
      }
    }
  } else if (warp == 1 || (!GATHER && warp == 2 + EW)) {
    // =============================== MMA issuer(s) ===============================
    // Executed by the whole warp so that descriptors stay in uniform registers; one elected lane issues.
    // Per-MMA scalar work is two 32-bit adds: the 64-bit descriptors are (constant hi word, running lo word)
    // [measured: with descriptors rebuilt from scratch per MMA the issue loop, not the tensor pipe, set the pace].
    //
    // DUAL ISSUE (k.dual: resident-weight layers, i.e. the thin 32-/64-channel convs): a second warp issues the MMAs of every
    // other tile into the other accumulator stage.  One issuing thread runs only about one MMA ahead of the tensor pipe and
    // stalls ~400 cycles per tile on the barrier hand-offs; with two independent issue streams the pipe stays fed while one
    // of them waits (the effect two co-resident CTAs have - measured -16 % on the 32-channel layers - without a second copy
    // of the weights and the halo ring in shared memory).  tcgen05.commit tracks the executing thread's own MMAs, so each
    // issuer releases exactly the stages it consumed.
    const int mw = warp == 1 ? 0 : 1;
    const int nmw = k.dual ? 2 : 1;
    if (mw < nmw) {
      const uint32_t idesc = make_idesc_bf16_m128((uint32_t)k.BN);
      const uint32_t a_lbo16 = (uint32_t)rowsA;            // LBO in 16-byte units (rows * 16 B)
      const uint32_t b_lbo16 = (uint32_t)k.BN;
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1, SWIZZLE_NONE
      const uint32_t a_lo_ring = (smem_u32(a_ring) >> 4) | (a_lbo16 << 16);
      const uint32_t b_lo_ring = (smem_u32(b_ring) >> 4) | (b_lbo16 << 16);
      const uint32_t a_stage16 = k.a_stage_bytes >> 4, b_stage16 = k.b_stage_bytes >> 4;
      const uint32_t a_jstep = 2u * a_lbo16, b_jstep = 2u * b_lbo16;   // K advance of 16 elements = two planes
      const uint32_t b_tap16 = (uint32_t)(k.KC * k.BN);                 // one tap's weight tile in 16-byte units
      int a_stage = mw, a_phase = 0, b_stage = 0, b_phase = 0;   // dual: one A stage per tile (nchunks == 1), issuer mw starts at stage mw
      const bool wres = k.wres != 0;   // resident weights (see the producer)
      for (int it = mw, tile = blockIdx.x + mw * gridDim.x; tile < k.num_tiles; tile += nmw * gridDim.x, it += nmw) {
        const int as = it & 1, aph = (it >> 1) & 1;
        HRNB_TRACE(1, it, 0);
        mbar_wait(&tmem_empty[as], aph ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after_sync();
        HRNB_TRACE(2, it, 0);
        const uint32_t d_base = tmem_base + (uint32_t)(as * acc_cols);
        uint32_t accumulate = 0;
        int grp = 0, ntaps = k.taps;
        if constexpr (GRP) {
          grp = (tile % k.n_tiles_all) / k.n_tiles;
          ntaps = k.grp_taps[grp];
        }
        for (int c = 0; c < k.nchunks; ++c) {
          // one wait per chunk for the A halo (flat-shift) and for the weights of all taps: every wait / commit
          // stalls the tensor pipe (~100-140 cycles each, measured), so hand-offs are per chunk, not per tap
          if (!GATHER) mbar_wait(&full_a[a_stage], a_phase);
          if (wres) b_stage = c;     // resident slab: chunk c lives in stage c, filled once (first tile of this CTA)
          if (!wres || it == mw) mbar_wait(&full_b[b_stage], b_phase);
          tc_fence_after_sync();
          if (c == 0) HRNB_TRACE(2, it, 1);
          uint32_t b_lo_tap = b_lo_ring + (uint32_t)b_stage * b_stage16;
          for (int t = 0; t < ntaps; ++t) {
            if (GATHER) {
              mbar_wait(&full_a[a_stage], a_phase);
              tc_fence_after_sync();
            }
            // tap t reads input position p + dpos[t] of source src[t]: a row offset into the halo stage (host table)
            const uint32_t toff = GATHER ? 0u : (uint32_t)(GRP ? k.grp_tap_off[grp][t & 3] : k.tap_off[t]);
            const uint32_t a_lo_tap = a_lo_ring + (uint32_t)a_stage * a_stage16 + toff;
            uint32_t d = d_base;
            uint32_t a_lo_mb = a_lo_tap;
            for (int mb = 0; mb < k.MB; ++mb) {
              uint64_t adesc = ((uint64_t)desc_hi << 32) | a_lo_mb;
              uint64_t bdesc = ((uint64_t)desc_hi << 32) | b_lo_tap;
              if (elect_one_sync()) {
                umma_bf16_ss(d, adesc, bdesc, idesc, accumulate);
#pragma unroll
                for (int j = 1; j < KSTEPS; ++j) {
                  adesc += a_jstep;
                  bdesc += b_jstep;
                  umma_bf16_ss(d, adesc, bdesc, idesc, 1u);
                }
              }
              d += (uint32_t)k.BN;
              a_lo_mb += 128u;   // next 128-row block: 128 rows * 16 B
            }
            accumulate = 1u;
            b_lo_tap += b_tap16;
            if (GATHER) {
              __syncwarp();
              if (elect_one_sync()) umma_commit(&empty_a[a_stage]);
              if (++a_stage == k.SA) { a_stage = 0; a_phase ^= 1; }
            }
          }
          __syncwarp();
          if (elect_one_sync()) {
            if (!wres) umma_commit(&empty_b[b_stage]);
            if (!GATHER) umma_commit(&empty_a[a_stage]);
          }
          if (!wres && ++b_stage == k.SB) { b_stage = 0; b_phase ^= 1; }
          if (!GATHER) {
            a_stage += nmw;
            if (a_stage >= k.SA) { a_stage -= k.SA; a_phase ^= 1; }
          }
        }
        HRNB_TRACE(1, it, 1);
        if (elect_one_sync()) umma_commit(&tmem_full[as]);
        __syncwarp();
      }
    }
  } else if (warp < 2 + EW) {
    // =============================== epilogue ===============================
    // The epilogue is HBM-latency bound (residual reads), so residuals are prefetched PD 16-channel groups ahead
    // into a register ring; the first PD groups of a tile are requested BEFORE waiting for its accumulator.
    constexpr int PD = 4;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int cs = (warp - 2) >> 2;     // which share of the column groups this warp takes
    constexpr int CS = EW / 4;   // warps per quarter
    const bool relu = (k.flags & HRNB_CONV_RELU) != 0;
    constexpr bool nchw = NCHW;
    const bool has_res = !STATS && k.res != nullptr && !(k.dbg & 1);
    const int groups = k.BN / 16;  // 16-column groups per M block
    // STATS: running sums of this thread's row over its column group, as packed fp32 pairs (FADD2 / FFMA2):
    // st2[j] = sums of columns (2j, 2j+1), st2[8 + j] = their sums of squares
    unsigned long long st2[STATS ? 16 : 1];
#pragma unroll
    for (int i = 0; i < (STATS ? 16 : 1); ++i) st2[i] = 0ull;
    const int E = k.MB * groups;   // groups per tile; this warp takes e = cs, cs + CS, ...
    int it = 0;
    if constexpr (LEAN) {
      // ---- per-warp item table: item u = (M block, 16-column group) number cs + u*CS of a tile; MB*BN <= 256 => <= 4 items
      constexpr int MAXI = 4;
      uint32_t tcol[MAXI];          // TMEM column of the item inside an accumulator stage
      long long ooff[MAXI], roff[MAXI];   // element offsets of the item's first plane from the tile's base pointers
      int boff[MAXI], imb[MAXI];
      const long long out_g = 2 * k.out_ps * 8, res_g = 2 * k.res_ps * 8;   // two planes per 16-channel group
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        const int e = cs + u * CS;
        int mb = 0, g = e;
        while (g >= groups) { g -= groups; ++mb; }
        imb[u] = mb;
        tcol[u] = (uint32_t)(mb * k.BN + g * 16);
        ooff[u] = (long long)g * out_g + (long long)mb * 1024;
        roff[u] = (long long)g * res_g + (long long)mb * 1024;
        boff[u] = g * 16;
      }
      __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(k.out);
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      // phase-split form of the output (input of a following 3x3 stride-2 conv): either the only output (OUT_PHASES) or a
      // second copy next to the plain PF8 tensor (out2); padding of the phase tensors is never written and stays zero
      const bool primary = PH ? k.out_phase_stride == 0 : true;
      __nv_bfloat16* const ph_base = PH ? (primary ? k.out2 : outp) : nullptr;
      const long long ph_ps = primary ? k.out2_ps : k.out_ps;
      const unsigned ph_stride16 = (unsigned)((primary ? k.out2_phase_stride : k.out_phase_stride) >> 3);
      // FUSE_AFTER_RELU: the phase copy receives the unit's own output ReLU(acc + bias + res); the fuse sources are added to
      // THAT and a second ReLU gives the primary output - the last conv of branch 0 hosting fuse output 0
      const bool fuse_after = PH && (k.flags & HRNB_CONV_FUSE_AFTER_RELU) != 0;
      for (int tile = blockIdx.x; tile < k.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1, aph = (it >> 1) & 1;
        const unsigned mg = k.n_tiles_all == 1 ? (unsigned)tile : fast_div((unsigned)tile, k.mNt, k.sNt);
        int ntile = tile - (int)mg * k.n_tiles_all;
        long long gout = 0, gres = 0;      // grouped launch: this tile's conv writes / accumulates into its own tensor
        if constexpr (GRP) {
          const int grp = ntile / k.n_tiles;
          ntile -= grp * k.n_tiles;
          gout = (long long)grp * k.grp_out_stride;
          gres = (long long)grp * k.grp_res_stride;
        }
        const unsigned p0 = mg * (unsigned)(k.MB * 128) + (unsigned)(q * 32 + lane);
        unsigned validm = 0, realm = 0;
        unsigned phoff[4] = {0u, 0u, 0u, 0u};   // phase-split copy / output: position of the row in 16-byte units, phase offset included
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          if (mb < k.MB) {
            const unsigned p = p0 + (unsigned)(mb * 128);
            const unsigned rowi = fast_div(p, k.mWp, k.sWp);
            const unsigned px = p - rowi * (unsigned)k.Wp;
            const unsigned n = fast_div(rowi, k.mHp, k.sHp);
            const unsigned py = rowi - n * (unsigned)k.Hp;
            if (p < (unsigned)k.P) {
              validm |= 1u << mb;
              if (px != 0u && py != 0u) realm |= 1u << mb;
            }
            if constexpr (PH) {
              const unsigned x = px - 1u, y = py - 1u;   // garbage on padding rows, which are never stored to the phases
              phoff[mb] = ((y & 1u) * 2u + (x & 1u)) * ph_stride16 + (n * (unsigned)k.oHp2 + (y >> 1) + 1u) * (unsigned)k.oWp2 + (x >> 1) + 1u;
            }
          }
        }
        const int plane0 = ntile * (k.BN / 8);
        __nv_bfloat16* const obase = outp + gout + ((long long)plane0 * k.out_ps + p0) * 8;
        const __nv_bfloat16* const rbase = k.res + gres + ((long long)plane0 * k.res_ps + p0) * 8;
        const float* const bias_t = bias_s + ntile * k.BN;
        // With 16 epilogue warps a warp owns <= 4 items of a tile (ROUNDS == 1); with 8 warps (two CTAs per SM) up to 8: a second
        // round over items 4..7, whose table entries are recomputed instead of being held in registers.
        constexpr int ROUNDS = (16 + MAXI * CS - 1) / (MAXI * CS);
        const uint32_t t_base = lane_base + (uint32_t)(as * acc_cols);
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
        if (rd > 0 && cs + rd * MAXI * CS >= E) break;
        uint32_t tcol_r[MAXI];
        long long ooff_r[MAXI], roff_r[MAXI];
        int boff_r[MAXI], imb_r[MAXI];
#pragma unroll
        for (int u = 0; u < MAXI; ++u) {
          if (rd == 0) {
            tcol_r[u] = tcol[u]; ooff_r[u] = ooff[u]; roff_r[u] = roff[u]; boff_r[u] = boff[u]; imb_r[u] = imb[u];
          } else {
            const int e = cs + (u + rd * MAXI) * CS;
            int mb = 0, g = e;
            while (g >= groups) { g -= groups; ++mb; }
            imb_r[u] = mb;
            tcol_r[u] = (uint32_t)(mb * k.BN + g * 16);
            ooff_r[u] = (long long)g * out_g + (long long)mb * 1024;
            roff_r[u] = (long long)g * res_g + (long long)mb * 1024;
            boff_r[u] = g * 16;
          }
        }
        // residuals of all items of this warp: requested before the accumulator is waited for
        uint4 rb[MAXI][2];
#pragma unroll
        for (int u = 0; u < MAXI; ++u) {
          rb[u][0] = rb[u][1] = make_uint4(0u, 0u, 0u, 0u);
          if (has_res && cs + (u + rd * MAXI) * CS < E && ((realm >> imb_r[u]) & 1u)) {
            rb[u][0] = ldg_nc_v4(rbase + roff_r[u]);
            rb[u][1] = ldg_nc_v4(rbase + roff_r[u] + k.res_ps * 8);
          }
        }
        if (rd == 0) {
          mbar_wait(&tmem_full[as], aph);
          tc_fence_after_sync();
        }
#pragma unroll
        for (int u = 0; u < MAXI; ++u) {
          if (cs + (u + rd * MAXI) * CS < E) {
            uint32_t v[16];
            tmem_ld16(t_base + tcol_r[u], v);
            tmem_ld_wait();
            const bool valid = (validm >> imb_r[u]) & 1u, real = (realm >> imb_r[u]) & 1u;
            unsigned long long x2[8];      // 8 packed column pairs
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = *reinterpret_cast<const float4*>(&bias_t[boff_r[u] + 4 * i]);
              x2[2 * i] = add_f32x2(pack_f32x2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), pack_f32x2(b4.x, b4.y));
              x2[2 * i + 1] = add_f32x2(pack_f32x2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), pack_f32x2(b4.z, b4.w));
            }
            if constexpr (STATS) {
              if (real) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  st2[j] = add_f32x2(st2[j], x2[j]);
                  st2[8 + j] = fma_f32x2(x2[j], x2[j], st2[8 + j]);
                }
              }
            }
            // residual first (the unit's own output), then - before or after its ReLU - the fuse sources
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint4 r = rb[u][h];
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) x2[h * 4 + j] = add_f32x2(x2[h * 4 + j], pack_f32x2(bf16_lo(rw[j]), bf16_hi(rw[j])));
            }
            uint32_t own[PH ? 8 : 1];   // FUSE_AFTER_RELU: the unit's own output as bf16 pairs (goes to the phase copy)
            if constexpr (PH) if (fuse_after) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float lo, hi;
                unpack_f32x2(x2[j], lo, hi);
                if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                own[j] = pack_bf16x2(lo, hi);
                x2[j] = pack_f32x2(lo, hi);
              }
            }
            if (k.nfuse != 0 && real) {
              // fuse-layer sum: the other branches' contributions, nearest up-sampled on the fly (tiny, L2-resident sources)
              const unsigned p = p0 + (unsigned)(imb_r[u] * 128);
              const unsigned rowi = fast_div(p, k.mWp, k.sWp);
              const unsigned n = fast_div(rowi, k.mHp, k.sHp);
              const int x = (int)(p - rowi * (unsigned)k.Wp) - 1, y = (int)(rowi - n * (unsigned)k.Hp) - 1;
              for (int f = 0; f < k.nfuse; ++f) {
                const int sh = k.fuse_shift[f];
                const long long sp = ((long long)n * ((k.H >> sh) + 1) + (y >> sh) + 1) * ((k.W >> sh) + 1) + (x >> sh) + 1;
                const __nv_bfloat16* fs = k.fuse_src[f] + ((long long)(plane0 + 2 * (boff_r[u] >> 4)) * k.fuse_ps[f] + sp) * 8;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint4 r = ldg_nc_v4(fs + (long long)h * k.fuse_ps[f] * 8);
                  const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) x2[h * 4 + j] = add_f32x2(x2[h * 4 + j], pack_f32x2(bf16_lo(rw[j]), bf16_hi(rw[j])));
                }
              }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t ow[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float lo, hi;
                unpack_f32x2(x2[h * 4 + j], lo, hi);
                ow[j] = relu ? pack_bf16x2_relu(lo, hi) : pack_bf16x2(lo, hi);
              }
              uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              if (!real) o = make_uint4(0u, 0u, 0u, 0u);      // keep the shared zero padding intact
              if (primary && valid) *reinterpret_cast<uint4*>(obase + ooff_r[u] + (long long)h * k.out_ps * 8) = o;
              if constexpr (PH) if (real) {
                if (fuse_after) o = make_uint4(own[h * 4], own[h * 4 + 1], own[h * 4 + 2], own[h * 4 + 3]);
                const int im = imb_r[u];   // selected without dynamic indexing (keeps phoff in registers)
                const unsigned po = im == 0 ? phoff[0] : (im == 1 ? phoff[1] : (im == 2 ? phoff[2] : phoff[3]));
                *reinterpret_cast<uint4*>(ph_base + ((long long)(plane0 + (boff_r[u] >> 3) + h) * ph_ps + po) * 8) = o;
              }
            }
          }
        }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
      }
    } else
    for (int tile = blockIdx.x; tile < k.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1, aph = (it >> 1) & 1;
      const int mg = tile / k.n_tiles, ntile = tile - mg * k.n_tiles;
      const int p0 = mg * k.MB * 128 + q * 32 + lane;  // row of M block 0 (P < 2^31)
      unsigned validm = 0, realm = 0;
      for (int mb = 0; mb < k.MB; ++mb) {
        const int p = p0 + mb * 128;
        const int rowi = p / k.Wp;
        const int px = p - rowi * k.Wp;
        const int py = rowi % k.Hp;
        if (p < k.P) {
          validm |= 1u << mb;
          if (px > 0 && py > 0) realm |= 1u << mb;
        }
      }
      const int plane0 = ntile * (k.BN / 8);
      const __nv_bfloat16* res_base = has_res ? k.res + ((long long)plane0 * k.res_ps + p0) * 8 : nullptr;
      const long long res_gstep = 2 * k.res_ps * 8;  // two planes per 16-channel group
      uint4 rb[PD][2];
      int pmb = 0, pg = cs;  // prefetch cursor (mb, group)
      while (pg >= groups) { pg -= groups; ++pmb; }
      auto prefetch = [&](uint4(&dst)[2]) {
        if (has_res && pmb < k.MB && ((realm >> pmb) & 1u)) {
          const __nv_bfloat16* src = res_base + (long long)pg * res_gstep + (long long)pmb * 1024;
          dst[0] = ldg_nc_v4(src);
          dst[1] = ldg_nc_v4(src + k.res_ps * 8);
        } else {
          dst[0] = dst[1] = make_uint4(0u, 0u, 0u, 0u);
        }
        pg += CS;
        while (pg >= groups) { pg -= groups; ++pmb; }
      };
#pragma unroll
      for (int u = 0; u < PD; ++u) prefetch(rb[u]);
      if (warp == 2) HRNB_TRACE(3, it, 0);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after_sync();
      if (warp == 2) HRNB_TRACE(3, it, 1);
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_cols);
      int cmb = 0, cg = cs;  // consume cursor
      while (cg >= groups) { cg -= groups; ++cmb; }
      for (int e0 = cs; e0 < ((k.dbg & 64) ? 0 : E); e0 += PD * CS) {
#pragma unroll
        for (int u = 0; u < PD; ++u) {
          if (e0 + u * CS < E) {
            const bool valid = (validm >> cmb) & 1u, real = (realm >> cmb) & 1u;
            const int p = p0 + cmb * 128;
            long long ph_off = 0;
            const long long ph_stride = k.out2 ? k.out2_phase_stride : k.out_phase_stride;
            if (ph_stride != 0 && real) {
              const int rowi = p / k.Wp;
              const int x = p - rowi * k.Wp - 1;
              const int n = rowi / k.Hp;
              const int y = rowi - n * k.Hp - 1;
              ph_off = (long long)((y & 1) * 2 + (x & 1)) * ph_stride +
                       ((long long)(n * k.oHp2 + (y >> 1) + 1) * k.oWp2 + (x >> 1) + 1) * 8;
            }
            uint32_t v[16];
            if (!(k.dbg & 4)) {
              tmem_ld16(t_base + (uint32_t)(cmb * k.BN + cg * 16), v);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
            const int cb = ntile * k.BN + cg * 16;
            float bsv[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[cb + 4 * i]);
              bsv[4 * i] = b4.x; bsv[4 * i + 1] = b4.y; bsv[4 * i + 2] = b4.z; bsv[4 * i + 3] = b4.w;
            }
            if constexpr (nchw) {
              if (real) {
                float* o = reinterpret_cast<float*>(k.out);
                const int rowi = p / k.Wp;
                const int px = p - rowi * k.Wp;
                const int n = rowi / k.Hp;
                const int py = rowi - n * k.Hp;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int c = cb + i;
                  if (c < k.cout) {
                    float x = __uint_as_float(v[i]) + bsv[i];
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
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[h * 8 + i]) + bsv[h * 8 + i];
                if constexpr (STATS) {
                  if (real) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const unsigned long long xx = pack_f32x2(x[2 * j], x[2 * j + 1]);
                      st2[h * 4 + j] = add_f32x2(st2[h * 4 + j], xx);
                      st2[8 + h * 4 + j] = fma_f32x2(xx, xx, st2[8 + h * 4 + j]);
                    }
                  }
                }
                const uint4 r = rb[u][h];  // zeros when there is no residual / padding row
                x[0] += bf16_lo(r.x); x[1] += bf16_hi(r.x);
                x[2] += bf16_lo(r.y); x[3] += bf16_hi(r.y);
                x[4] += bf16_lo(r.z); x[5] += bf16_hi(r.z);
                x[6] += bf16_lo(r.w); x[7] += bf16_hi(r.w);
                uint4 o;
                if (relu) {   // ReLU is folded into the bf16x2 conversion
                  o.x = pack_bf16x2_relu(x[0], x[1]);
                  o.y = pack_bf16x2_relu(x[2], x[3]);
                  o.z = pack_bf16x2_relu(x[4], x[5]);
                  o.w = pack_bf16x2_relu(x[6], x[7]);
                } else {
                  o.x = pack_bf16x2(x[0], x[1]);
                  o.y = pack_bf16x2(x[2], x[3]);
                  o.z = pack_bf16x2(x[4], x[5]);
                  o.w = pack_bf16x2(x[6], x[7]);
                }
                if (!real) o = make_uint4(0u, 0u, 0u, 0u);  // keep the shared zero padding intact
                if (valid && !(k.dbg & 2)) {
                  const long long plane = (long long)(cb / 8 + h);
                  if (k.out_phase_stride == 0) {
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(k.out) + (plane * k.out_ps + p) * 8) = o;
                    if (k.out2 != nullptr && real)
                      *reinterpret_cast<uint4*>(k.out2 + ph_off + plane * k.out2_ps * 8) = o;
                  } else if (real) {
                    // write the output as 4 half-resolution phase tensors (input of a following stride-2 conv);
                    // their padding is never written and stays zero
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(k.out) + ph_off + plane * k.out_ps * 8) = o;
                  }
                }
              }
            }
            prefetch(rb[u]);  // refill this ring slot with the group PD ahead
            cg += CS;
            while (cg >= groups) { cg -= groups; ++cmb; }
          }
        }
      }
      if (warp == 2) HRNB_TRACE(4, it, 0);
      // hand the accumulator stage back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    if constexpr (STATS) {
      float st[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        unpack_f32x2(st2[j], st[2 * j], st[2 * j + 1]);
        unpack_f32x2(st2[8 + j], st[16 + 2 * j], st[16 + 2 * j + 1]);
      }
      // transposed butterfly: 32 values x 32 lanes -> lane l holds the warp total of value l (fixed order)
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send = upper ? st[i] : st[i + off];
          const float keep = upper ? st[i + off] : st[i];
          st[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      bias_s[256 + (warp - 2) * 32 + lane] = st[0];   // bias uses <= 64 floats in this variant; 16 warps x 32 floats
    }
  } else {
    // =============================== gather producers (GATHER only) ===============================
    if (GATHER && warp >= 3 + EW) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      const int g = threadIdx.x - (96 + 32 * EW);  // A row handled by this thread (per M block)
      const int LAG = k.lag;
      auto wait_lag = [&]() {   // cp.async.wait_group needs an immediate
        switch (LAG) {
          case 1: cp_async_wait<1>(); break;
          case 2: cp_async_wait<2>(); break;
          case 3: cp_async_wait<3>(); break;
          case 4: cp_async_wait<4>(); break;
          case 5: cp_async_wait<5>(); break;
          default: cp_async_wait<6>(); break;
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < k.num_tiles; tile += gridDim.x) {
        const int mg = tile / k.n_tiles;
        long long pin[4];
        for (int mb = 0; mb < k.MB; ++mb) {
          const long long p = ((long long)mg * k.MB + mb) * 128 + g;
          const int px = (int)(p % k.Wp);
          const int rowi = (int)(p / k.Wp);
          const int py = rowi % k.Hp;
          const int n = rowi / k.Hp;
          const bool real = p < k.P && px > 0 && py > 0;
          // input position of tap (0,0) for this output pixel; padding rows read position -1 (guard band = 0)
          pin[mb] = real ? ((long long)n * k.in_Hp + (long long)(py - 1) * k.stride) * k.in_Wp +
                               (long long)(px - 1) * k.stride
                         : -1;
          // 1x1 conv on a nearest-up-sampled input (in_up_shift): output pixel (y, x) reads input pixel (y >> s, x >> s)
          if (k.in_up_shift != 0 && real)
            pin[mb] = ((long long)n * k.in_Hp + ((py - 1) >> k.in_up_shift)) * k.in_Wp + ((px - 1) >> k.in_up_shift);
        }
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
              wait_lag();
              fence_proxy_async_smem();
              mbar_arrive(&full_a[(it - LAG) % k.SA]);
            }
          }
        }
      }
      // drain the last LAG stages
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int d = (it > LAG ? it - LAG : 0); d < it; ++d) mbar_arrive(&full_a[d % k.SA]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)k.tmem_cols);
  }
  if constexpr (STATS) {
    const int groups = k.BN / 16, nval = groups * 32;
    const int t = threadIdx.x;
    unsigned* counter = reinterpret_cast<unsigned*>(k.stats_ws);
    float* partials = k.stats_ws + 32;
    // the ticket lives in the spare bytes of the barrier block (36 mbarriers + the TMEM pointer end at byte 292 of 384):
    // a static __shared__ variable would push dynamic + static shared memory over the 227 KB opt-in limit
    volatile unsigned& ticket_s = *reinterpret_cast<volatile unsigned*>(smem + 320);
    if (t < nval) {
      // column group g is drained by the warps with cs % groups == g (cs = epilogue warp index / 4), all four lane quarters
      const int g = t >> 5, i = t & 31;
      float s = 0.f;
      for (int cs = g; cs < EW / 4; cs += groups) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s += bias_s[256 + (cs * 4 + j) * 32 + i];
      }
      partials[(size_t)blockIdx.x * kStatsSlot + t] = s;
      __threadfence();
    }
    __syncthreads();
    if (t == 0) ticket_s = atomicAdd(counter, 1u);
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {   // last CTA: every partial is visible
      __threadfence();
      float* red2 = bias_s + 256;      // [4 slices][128]
      const int idx = t & 127, slice = t >> 7;
      if (slice < 4) {
        float s = 0.f;
        if (idx < nval) {
#pragma unroll 8
          for (unsigned b = slice; b < gridDim.x; b += 4) s += __ldcg(partials + (size_t)b * kStatsSlot + idx);
        }
        red2[slice * 128 + idx] = s;
      }
      __syncthreads();
      if (t < nval) {
        const float tot = (red2[t] + red2[128 + t]) + (red2[256 + t] + red2[384 + t]);
        const int g = t >> 5, i = t & 31;
        const int ch = g * 16 + (i & 15);
        if (ch < k.cout) k.stats_sums[2 * ch + (i >> 4)] = tot;
      }
      if (t == 0) *counter = 0u;
    }
  }
}

int g_debug[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
long long* g_trace = nullptr;

// (m, s) of fast_div for divisor d >= 2
static void fast_magic(unsigned d, unsigned* m, unsigned* s) {
  unsigned c = 1;
  while ((1ull << c) < d) ++c;
  const unsigned long long num = 1ull << (31 + c);
  *m = (unsigned)((num + d - 1) / d);
  *s = c - 1;
}

static int next_pow2_cols(int c) {
  int r = 32;
  while (r < c) r <<= 1;
  return r;
}

struct Launch {
  long long smem;
  int grid;
};

// Derive launch geometry; returns smem bytes or <0.
static long long derive(const hrnb_conv_params* p, ConvK* k) {
  if (!p || !p->in || !p->wpk || !p->bias || !p->out) return fail(HRNB_EINVAL, "conv: null pointer");
  const bool gather = (p->flags & HRNB_CONV_GATHER) != 0;
  const bool custom = p->ntap_custom > 0;
  if (custom) {
    if (gather || p->stride != 1 || p->taps != p->ntap_custom || p->ntap_custom > 9 || (p->flags & HRNB_CONV_IN_PHASES))
      return fail(HRNB_EINVAL, "conv: custom taps need the flat-shift path, stride 1 and taps == ntap_custom <= 9");
  } else if (p->taps != 1 && p->taps != 9) return fail(HRNB_EINVAL, "conv: taps must be 1 or 9");
  if (p->stride != 1 && p->stride != 2) return fail(HRNB_EINVAL, "conv: stride must be 1 or 2");
  const bool phases_in = (p->flags & HRNB_CONV_IN_PHASES) != 0;
  if (p->stride == 2 && !gather && !phases_in) return fail(HRNB_EINVAL, "conv: stride 2 needs HRNB_CONV_GATHER or HRNB_CONV_IN_PHASES");
  if (phases_in && (gather || p->stride != 2 || p->taps != 9)) return fail(HRNB_EINVAL, "conv: HRNB_CONV_IN_PHASES is for 3x3 stride-2 flat-shift convs");
  if ((p->flags & HRNB_CONV_OUT_PHASES) && ((p->flags & HRNB_CONV_OUT_NCHW) || (p->H & 1) || (p->W & 1)))
    return fail(HRNB_EINVAL, "conv: HRNB_CONV_OUT_PHASES needs PF8 output with even H and W");
  if (p->cin % 16 || p->cin <= 0) return fail(HRNB_EINVAL, "conv: cin must be a positive multiple of 16");
  if (p->KC <= 0 || (p->KC & 1) || (p->cin / 8) % p->KC) return fail(HRNB_EINVAL, "conv: KC must be even and divide cin/8");
  if (p->BN < 16 || p->BN > 256 || p->BN % 16) return fail(HRNB_EINVAL, "conv: BN must be a multiple of 16 in [16,256]");
  if (p->MB != 1 && p->MB != 2 && p->MB != 4) return fail(HRNB_EINVAL, "conv: MB must be 1, 2 or 4");
  if (p->MB * p->BN > 256) return fail(HRNB_EINVAL, "conv: MB*BN exceeds 256 (two accumulator stages in 512 TMEM columns)");
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->cout <= 0) return fail(HRNB_EINVAL, "conv: bad geometry");
  if (p->in_up_shift != 0) {
    if (!gather || p->taps != 1 || p->stride != 1 || p->in_up_shift < 0 || p->in_up_shift > 3 ||
        p->in_H != (p->H >> p->in_up_shift) || p->in_W != (p->W >> p->in_up_shift) || (p->H & ((1 << p->in_up_shift) - 1)) ||
        (p->W & ((1 << p->in_up_shift) - 1)))
      return fail(HRNB_EINVAL, "conv: in_up_shift needs a gathered 1x1 stride-1 conv with in_H/in_W == H/W >> in_up_shift");
  } else if (p->in_H != p->H * p->stride || p->in_W != p->W * p->stride) return fail(HRNB_EINVAL, "conv: in_H/in_W must equal H*stride/W*stride");
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
  k->n_tiles = (p->cout + p->BN - 1) / p->BN;
  k->nbias = k->n_tiles * p->BN;
  k->dbg = g_debug[3];
  k->trace = g_trace;
  if (k->nbias * 4 > kBiasBytes) return fail(HRNB_EINVAL, "conv: more than 768 (padded) output channels");
  const int mblocks = (k->P + 127) / 128;
  k->ngroup = p->ngroup > 1 ? p->ngroup : 1;
  if (k->ngroup > 4) return fail(HRNB_EINVAL, "conv: ngroup must be <= 4");
  if (k->ngroup > 1 && (gather || !custom || phases_in || nchw || p->out2 || p->stats_sums || p->nfuse > 0 ||
                        (p->flags & HRNB_CONV_OUT_PHASES)))
    return fail(HRNB_EINVAL, "conv: a grouped launch needs the flat-shift path with custom taps and a plain PF8 output");
  k->n_tiles_all = k->n_tiles * k->ngroup;
  k->num_tiles = ((mblocks + p->MB - 1) / p->MB) * k->n_tiles_all;
  k->grp_out_stride = p->grp_out_stride;
  k->grp_res_stride = p->grp_res_stride;
  k->tmem_cols = next_pow2_cols(2 * p->MB * p->BN);
  k->src_stride = p->in_phase_stride;
  for (int t = 0; t < 9; ++t) k->tap_off[t] = 0;
  k->out_phase_stride = (p->flags & HRNB_CONV_OUT_PHASES) ? p->out_phase_stride : 0;
  k->out2 = (__nv_bfloat16*)p->out2;
  k->out2_ps = p->out2_ps;
  k->out2_phase_stride = p->out2_phase_stride;
  if (p->out2 && ((p->flags & (HRNB_CONV_OUT_NCHW | HRNB_CONV_OUT_PHASES)) || (p->H & 1) || (p->W & 1) || p->out2_phase_stride <= 0))
    return fail(HRNB_EINVAL, "conv: out2 needs a plain PF8 primary output with even H and W");
  k->oHp2 = p->H / 2 + 1;
  k->oWp2 = p->W / 2 + 1;
  k->stats_sums = p->stats_sums;
  k->stats_ws = p->stats_ws;
  k->nfuse = p->nfuse;
  k->in_up_shift = p->in_up_shift;
  if (p->nfuse < 0 || p->nfuse > 3) return fail(HRNB_EINVAL, "conv: nfuse must be 0..3");
  for (int f = 0; f < 3; ++f) {
    k->fuse_src[f] = f < p->nfuse ? (const __nv_bfloat16*)p->fuse_src[f] : nullptr;
    k->fuse_ps[f] = p->fuse_ps[f];
    k->fuse_shift[f] = p->fuse_shift[f];
    if (f < p->nfuse) {
      const int sh = p->fuse_shift[f];
      if (!p->fuse_src[f] || sh < 0 || sh > 3 || (p->H & ((1 << sh) - 1)) || (p->W & ((1 << sh) - 1)))
        return fail(HRNB_EINVAL, "conv: bad fuse source (NULL, or shift not in 0..3 / not dividing H and W)");
    }
  }
  if (p->nfuse > 0 && ((p->flags & (HRNB_CONV_OUT_NCHW | HRNB_CONV_OUT_PHASES)) || p->stats_sums))
    return fail(HRNB_EINVAL, "conv: fuse sources need a PF8 primary output without fused statistics");
  if ((p->flags & HRNB_CONV_FUSE_AFTER_RELU) && (p->nfuse == 0 || (p->flags & HRNB_CONV_GATHER)))
    return fail(HRNB_EINVAL, "conv: HRNB_CONV_FUSE_AFTER_RELU needs fuse sources on the flat-shift path");
  fast_magic((unsigned)k->Wp, &k->mWp, &k->sWp);
  fast_magic((unsigned)k->Hp, &k->mHp, &k->sHp);
  k->mNt = 0; k->sNt = 0;
  if (k->n_tiles_all > 1) fast_magic((unsigned)k->n_tiles_all, &k->mNt, &k->sNt);
  if (p->stats_sums != nullptr) {
    const int groups = p->BN / 16;
    if (!p->stats_ws || gather || nchw || k->n_tiles != 1 || (groups != 1 && groups != 2 && groups != 4) || p->res != nullptr ||
        (p->flags & HRNB_CONV_RELU))
      return fail(HRNB_EINVAL, "conv: fused BatchNorm statistics need the flat-shift PF8 path, BN == cout in {16, 32, 64}, no residual, no ReLU");
  }
  k->lead = 0;
  if (phases_in && p->in_phase_stride <= 0) return fail(HRNB_EINVAL, "conv: in_phase_stride missing");
  if ((p->flags & HRNB_CONV_OUT_PHASES) && p->out_phase_stride <= 0) return fail(HRNB_EINVAL, "conv: out_phase_stride missing");
  k->b_stage_bytes = (unsigned)(p->taps * p->KC * p->BN * 16);   // all taps of one K chunk
  // tap table: tap t reads source tsrc[t] at position p + tdpos[t]
  int tsrc[9] = {0}, tdpos[9] = {0};
  k->nsrc = 1;
  if (custom) {
    for (int t = 0; t < p->taps; ++t) {
      tsrc[t] = p->tap_src[t];
      tdpos[t] = p->tap_dpos[t];
      if (tsrc[t] < 0 || tsrc[t] > 3) return fail(HRNB_EINVAL, "conv: tap_src out of range");
      if (tsrc[t] + 1 > k->nsrc) k->nsrc = tsrc[t] + 1;
    }
    if (k->nsrc > 1 && p->in_phase_stride <= 0) return fail(HRNB_EINVAL, "conv: in_phase_stride missing");
  } else if (phases_in) {     // tap (r,s) reads phase (r != 1, s != 1) at (dy,dx) = (-(r == 0), -(s == 0)) on the half grid
    k->nsrc = 4;
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t % 3;
      tsrc[t] = (r != 1) * 2 + (s != 1);
      tdpos[t] = -(r == 0) * k->Wp - (s == 0);
    }
  } else if (p->taps == 9) {
    for (int t = 0; t < 9; ++t) tdpos[t] = (t / 3 - 1) * k->Wp + (t % 3 - 1);
  }
  if (gather) {
    k->halo = 128 * p->MB;
    k->a_stage_bytes = (unsigned)(p->KC * 128 * p->MB * 16);
    k->SA = 8;   // deep ring: the gather is latency bound, the producer runs SA - 2 taps ahead
  } else {
    int lo = 0, hi = 0;
    for (int t = 0; t < p->taps; ++t) {
      if (tdpos[t] < lo) lo = tdpos[t];
      if (tdpos[t] > hi) hi = tdpos[t];
    }
    int max_taps = p->taps;
    if (k->ngroup > 1) {      // one halo serves every conv of the group: union of their tap ranges
      for (int g = 0; g < k->ngroup; ++g) {
        if (p->grp_ntap[g] < 1 || p->grp_ntap[g] > 4 || !p->grp_wpk[g]) return fail(HRNB_EINVAL, "conv: grouped launch: 1..4 taps and packed weights per conv");
        if (p->grp_ntap[g] > max_taps) max_taps = p->grp_ntap[g];
        for (int t = 0; t < p->grp_ntap[g]; ++t) {
          if (p->grp_tap_dpos[g][t] < lo) lo = p->grp_tap_dpos[g][t];
          if (p->grp_tap_dpos[g][t] > hi) hi = p->grp_tap_dpos[g][t];
        }
      }
      k->b_stage_bytes = (unsigned)(max_taps * p->KC * p->BN * 16);   // ring stage = the widest conv of the group
    }
    if (-lo > HRNB_GUARD_LEAD(k->Wp) || hi > HRNB_GUARD_LEAD(k->Wp)) return fail(HRNB_EINVAL, "conv: tap offset exceeds the PF8 guard band");
    k->lead = -lo;
    k->halo = 128 * p->MB + hi - lo;
    for (int t = 0; t < p->taps; ++t) k->tap_off[t] = tsrc[t] * p->KC * k->halo + k->lead + tdpos[t];
    for (int g = 0; g < 4; ++g) {
      k->grp_taps[g] = g < k->ngroup && k->ngroup > 1 ? p->grp_ntap[g] : 0;
      k->grp_wpk[g] = g < k->ngroup && k->ngroup > 1 ? (const __nv_bfloat16*)p->grp_wpk[g] : nullptr;
      k->grp_b_bytes[g] = (unsigned)(k->grp_taps[g] * p->KC * p->BN * 16);
      for (int t = 0; t < 4; ++t) k->grp_tap_off[g][t] = (k->ngroup > 1 && g < k->ngroup && t < p->grp_ntap[g]) ? k->lead + p->grp_tap_dpos[g][t] : 0;
    }
    k->a_stage_bytes = (unsigned)(k->nsrc * p->KC * k->halo * 16);
    k->SA = 2;  // next tile / next chunk is prefetched while the current one is multiplied
  }
  // resident weights: the weight slab of ONE N tile stays in shared memory for the life of the CTA.  Default: single-chunk,
  // single-N-tile layers (+0.5 ... 1 % in-trip).  OPT-IN (hrnb_debug_set(10, 1) / HRNB_SLAB=1): multi-chunk slabs, one ring stage
  // per K chunk, and several N tiles with the grid a multiple of n_tiles so that every CTA keeps to one N tile (the 480-channel
  // head conv as 3 x 160).  Measured SLOWER in-trip: the head conv 787 vs 629 us at batch 256 - with 206 KB of the slab + two
  // 24 KB halo stages only 49 KB of activations are in flight per SM, and the loads, not the re-streamed weights, set the pace.
  k->wres = 0;
  if (k->ngroup == 1 && k->nchunks <= kMaxSB && k->n_tiles <= 148 &&
      (k->nchunks == 1 ? k->n_tiles == 1
                       : !gather && g_debug[10] != 0 && kSmemHeader + 2LL * k->a_stage_bytes + (long long)k->nchunks * k->b_stage_bytes <= 216 * 1024))
    k->wres = 1;
  // dual issue: two tiles in flight need two A stages, a third one is the prefetch; hrnb_debug_set(8, 1) turns it off (A/B)
  k->dual = 0;
  if (k->wres && k->nchunks == 1 && !gather && g_debug[8] == 0 &&
      kSmemHeader + 3LL * k->a_stage_bytes + (long long)k->b_stage_bytes <= 200 * 1024) {
    k->dual = 1;
    k->SA = 3;
    if (kSmemHeader + 4LL * k->a_stage_bytes + (long long)k->b_stage_bytes <= 200 * 1024) k->SA = 4;   // deeper halo prefetch
  }
  // twin: two CTAs per SM (8 epilogue warps each, single issuer, two halo stages) for plain lean launches of resident-weight
  // layers with enough tiles.  OPT-IN (hrnb_debug_set(9, n) / HRNB_TWIN_MIN=n: at least n tiles per CTA slot): next to the dual
  // issuers and the 4-stage halo ring it measured no gain in-trip (batch 256: 21.8 - 22.0 k vs 22.1 - 22.2 k images/s)
  k->twin = 0;
  if (g_debug[9] > 0) {
    const long long min_tiles = (long long)g_debug[9] * 2 * 148;
    const bool plain = !gather && !nchw && p->out2 == nullptr && !(p->flags & HRNB_CONV_OUT_PHASES) && p->stats_sums == nullptr &&
                       k->ngroup == 1 && g_debug[3] == 0 && g_debug[0] == 0;
    if (k->wres && k->nchunks == 1 && plain && k->tmem_cols <= 256 && k->num_tiles >= min_tiles &&
        kSmemHeader + 2LL * k->a_stage_bytes + (long long)k->b_stage_bytes <= 110 * 1024) {
      k->twin = 1;
      k->dual = 0;
      k->SA = 2;
    }
  }
  const long long limit = 200 * 1024;
  int SB = k->wres ? k->nchunks : 3;   // resident weights: one stage per K chunk
  auto total = [&](int sa, int sb) { return kSmemHeader + (long long)sa * k->a_stage_bytes + (long long)sb * k->b_stage_bytes; };
  while (!k->wres && SB > 2 && total(k->SA, SB) > limit) --SB;
  while (gather && k->SA > 3 && total(k->SA, SB) > limit) --k->SA;
  k->lag = gather ? k->SA - 2 : 0;
  const long long smem = total(k->SA, SB);
  if (smem > 227 * 1024) return fail(HRNB_EINVAL, "conv: tile does not fit in shared memory (reduce KC, BN or MB)");
  k->SB = SB;
  return smem;
}

int bind_hang_buffer_conv() {
  unsigned long long* d = hang_buffer_device_ptr();
  if (d == nullptr) return fail(HRNB_ECUDA, "hang buffer: cudaHostAlloc failed");
  cudaError_t e = cudaMemcpyToSymbol(g_hang_buf, &d, sizeof(d));
  return e == cudaSuccess ? HRNB_OK : fail_cuda(e, "hang buffer: cudaMemcpyToSymbol");
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return HRNB_EINVAL;
  g_debug[key] = value;
  return HRNB_OK;
}

extern "C" int hrnb_debug_trace(void* dev_buf_5x64_i64) {
  g_trace = (long long*)dev_buf_5x64_i64;
  return HRNB_OK;
}

extern "C" int64_t hrnb_conv_stats_ws_floats(void) { return 32 + (int64_t)kStatsMaxCtas * kStatsSlot; }

extern "C" int64_t hrnb_conv_smem_bytes(const hrnb_conv_params* p) {
  ConvK k;
  return derive(p, &k);
}

extern "C" int hrnb_conv(const hrnb_conv_params* p, void* stream) {
  ConvK k;
  const long long smem = derive(p, &k);
  if (smem < 0) return (int)smem;
  const bool gather = (p->flags & HRNB_CONV_GATHER) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  // per-device launch state: written once per (device, variant), read on every launch, possibly from one host thread per
  // GPU (nn.DataParallel calls forward that way, tools/train.py:254) -> atomics; cudaFuncSetAttribute itself is idempotent
  static std::atomic<unsigned char> attr_set[64][72] = {};
  static std::atomic<int> sm_count[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  const bool nchw_out = (p->flags & HRNB_CONV_OUT_NCHW) != 0;
  if (gather && nchw_out) return fail(HRNB_EINVAL, "conv: NCHW output is not available for the gather variant");
  const bool stats = p->stats_sums != nullptr;
  const int ks = p->KC / 2;
  int ksi = -1;
  const void* fn = nullptr;
  // the lean epilogue covers the flat-shift path with a plain PF8 output (no phase-split form or copy)
  const bool phases = k.out_phase_stride != 0 || k.out2 != nullptr;
  // the lean epilogue: PF8 output; the phase-split form / copy only on the flat-shift path without fused statistics
  const bool lean_ok = !nchw_out && (!phases || (!gather && !stats));
  if ((p->flags & HRNB_CONV_FUSE_AFTER_RELU) && (k.out2 == nullptr || !lean_ok))
    return fail(HRNB_EINVAL, "conv: HRNB_CONV_FUSE_AFTER_RELU needs out2 on the flat-shift path");
  if (k.nfuse > 0 && !lean_ok) return fail(HRNB_EINVAL, "conv: fuse sources need the lean epilogue (plain PF8 output)");
  if (k.ngroup > 1 && (!lean_ok || g_debug[3] != 0 || g_debug[0] != 0)) return fail(HRNB_EINVAL, "conv: a grouped launch needs the lean epilogue");
  const bool lean = lean_ok && ((g_debug[3] == 0 && g_debug[0] == 0) || k.nfuse > 0);
  const bool twin = k.twin != 0 && lean;
#define HRNB_PICK(KS, IDX)                                                                                   \
  if (ks == KS) {                                                                                            \
    ksi = IDX;                                                                                               \
    fn = gather ? (lean ? (const void*)conv_tc_kernel<true, false, KS, false, true>                          \
                        : (const void*)conv_tc_kernel<true, false, KS>)                                      \
                : (nchw_out ? (const void*)conv_tc_kernel<false, true, KS>                                   \
                            : (lean ? (stats ? (const void*)conv_tc_kernel<false, false, KS, true, true>     \
                                             : (phases ? (const void*)conv_tc_kernel<false, false, KS, false, true, true>   \
                                                       : (k.ngroup > 1 ? (const void*)conv_tc_kernel<false, false, KS, false, true, false, true>   \
                                                                       : (const void*)conv_tc_kernel<false, false, KS, false, true>)))   \
                                    : (stats ? (const void*)conv_tc_kernel<false, false, KS, true>           \
                                             : (const void*)conv_tc_kernel<false, false, KS>)));             \
  }
  HRNB_PICK(1, 0) HRNB_PICK(2, 1) HRNB_PICK(3, 2) HRNB_PICK(4, 3) HRNB_PICK(6, 4) HRNB_PICK(8, 5) HRNB_PICK(16, 6)
#undef HRNB_PICK
  // two CTAs per SM with 8 epilogue warps each: plain lean launches of resident-weight layers (single issuer, two halo stages
  // per CTA: derive() sized them for that) with at least g_debug[9] (default 4) tiles per CTA slot
  if (twin) {
#define HRNB_TWIN(KS) if (ks == KS) fn = (const void*)conv_tc_kernel<false, false, KS, false, true, false, false, 8>;
    HRNB_TWIN(1) HRNB_TWIN(2) HRNB_TWIN(3) HRNB_TWIN(4) HRNB_TWIN(6) HRNB_TWIN(8) HRNB_TWIN(16)
#undef HRNB_TWIN
  }
  if (!fn) return fail(HRNB_EINVAL, "conv: KC must be one of 2, 4, 6, 8, 12, 16, 32");
  const int variant = twin ? 64 + ksi : ksi * 9 + (gather ? (lean ? 6 : 2) : (nchw_out ? 1 : (lean && phases ? 5 : (k.ngroup > 1 ? 8 : (stats ? 3 : 0) + (lean ? 4 : 0)))));
  if (!attr_set[dev][variant].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "conv: cudaFuncSetAttribute");
    attr_set[dev][variant].store(1, std::memory_order_release);
  }
  int nsm = sm_count[dev].load(std::memory_order_relaxed);
  if (nsm == 0) {
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    sm_count[dev].store(nsm, std::memory_order_relaxed);
  }
  // persistent grid: one or two CTAs per SM (two when shared memory and TMEM columns allow it)
  int per_sm = ((kCtasPerSm == 2 || twin) && !gather && smem <= 110 * 1024 && 2 * k.tmem_cols <= 512) ? 2 : 1;
  if (g_debug[1] > 0) per_sm = g_debug[1] == 1 ? 1 : per_sm;   // debug: force one CTA per SM
  int grid = nsm * per_sm;
  if (grid > k.num_tiles) grid = k.num_tiles;
  if (k.wres && k.n_tiles_all > 1 && grid >= k.n_tiles_all) grid -= grid % k.n_tiles_all;   // every CTA keeps to one N tile
  if (stats && k.BN / 16 > kEpiWarps / 4 && k.BN / 16 != 1)
    return fail(HRNB_EINVAL, "conv: fused statistics need one fixed column group per epilogue warp (BN/16 <= epilogue warps / 4)");
  if (stats && grid > kStatsMaxCtas) return fail(HRNB_EINVAL, "conv: fused statistics support at most 320 CTAs");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(gather ? (unsigned)kThreadsGather : (twin ? 96u + 32u * 8u : (unsigned)kThreadsFS));
  cfg.dynamicSmemBytes = (size_t)(per_sm == 1 && smem < kTmemExclusiveSmem && g_debug[6] == 0 ? kTmemExclusiveSmem : smem);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_debug[2] || (p->flags & HRNB_CONV_NO_PDL)) ? 0 : 1;   // debug knob 2 / flag: no programmatic dependent launch
  void* kargs[1] = {(void*)&k};
  cudaError_t le = cudaLaunchKernelExC(&cfg, fn, kargs);
  count_launch();
  if (le != cudaSuccess) return fail_cuda(le, "conv_tc_kernel launch");
  return check_launch("conv_tc_kernel");
}
