/* hrnb.h — C ABI of libhrnb.so, the sm_100a implementation of the HRNet hand-pose hot path.
 *
 * The reference (ZJULiHongxin/HRNet-Hand-Pose-Estimation) has no C/FFI boundary on this path: every
 * function below replaces a PyTorch/cuDNN/numpy call made from the reference's Python files. The
 * "replaces" notes cite reference file:line (relative to the reference's repository root).
 *
 * Conventions
 *   - every function returns 0 on success or a negative HRNB_E* code; hrnb_last_error() gives a
 *     thread-local message. Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer.
 *   - calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no allocation.
 *   - activations between convolutions use the PF8 layout: for a logical tensor
 *     [N, C, H, W]   Hp = H+1, Wp = W+1, P = N*Hp*Wp positions, p = (n*Hp + py)*Wp + px with real pixels
 *     at py in [1,H], px in [1,W] (row 0 / column 0 are shared zero padding), C/8 planes of
 *     [P][8] bf16 (16 bytes per position), plane stride `ps` positions, with zero guard bands of at
 *     least HRNB_GUARD_LEAD(Wp) positions before p=0 and HRNB_GUARD_TAIL(Wp) after p=P-1.
 */
#ifndef HRNB_H_
#define HRNB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRNB_OK 0
#define HRNB_EINVAL (-1)  /* bad argument / unsupported shape */
#define HRNB_ECUDA (-2)   /* CUDA runtime error (message in hrnb_last_error) */

#define HRNB_ABI_VERSION 5

/* guard bands (in positions) a PF8 plane must carry around [0, P) */
#define HRNB_GUARD_LEAD(Wp) ((((Wp) + 2) + 7) / 8 * 8)
#define HRNB_GUARD_TAIL(Wp) (((((Wp) + 2) + 7) / 8 * 8) + 512)

/* ---- convolution (implicit GEMM on tcgen05 / TMEM) ------------------------------------------ */

enum {
  HRNB_CONV_RELU = 1,       /* ReLU after bias (+ residual)                                   */
  HRNB_CONV_OUT_NCHW = 2,   /* write fp32 NCHW [N, cout_real, H, W] instead of PF8 bf16       */
  HRNB_CONV_GATHER = 4,     /* per-tap gathered A operand (any stride); otherwise flat-shift      */
  HRNB_CONV_IN_PHASES = 8,  /* 3x3 stride-2 conv whose input is given as 4 phase tensors (see below) */
  HRNB_CONV_OUT_PHASES = 16, /* write the PF8 output as 4 phase tensors for a following stride-2 conv */
  HRNB_CONV_NO_PDL = 32,    /* launch without programmatic dependent launch (plain stream order)      */
  HRNB_CONV_FUSE_AFTER_RELU = 64 /* with nfuse > 0: out2 (phase copy) = ReLU(acc + bias + res), the unit's own output;
                                    out = ReLU(that + fuse sources) - the last conv of a branch hosting its fuse output */
};

/* One conv + folded-BN bias (+ residual) (+ ReLU) launch.
 * Replaces nn.Conv2d -> nn.BatchNorm2d(eval) -> [+= residual] -> nn.ReLU chains:
 *   lib/models/pose_hrnet.py:43-59 (BasicBlock), :80-100 (Bottleneck), :419-458 (transition),
 *   :187-242 (fuse-layer convs), :335-350 (last_layer).
 * Weights are pre-packed by hrnb_pack_conv_weights (BN scale folded in, bf16). */
typedef struct hrnb_conv_params {
  const void* in;        /* PF8 bf16 input, p = 0 of plane 0                                    */
  int64_t in_ps;         /* input plane stride, positions                                       */
  const void* wpk;       /* packed weights [ntile][chunk][tap][KC][BN][8] bf16                  */
  const float* bias;     /* [n_tiles*BN] fp32 (folded BN shift / conv bias), zero padded        */
  const void* res;       /* optional PF8 residual (same geometry as out) or NULL                */
  int64_t res_ps;
  void* out;             /* PF8 bf16 output (or fp32 NCHW when HRNB_CONV_OUT_NCHW)              */
  int64_t out_ps;
  int32_t N, H, W;       /* OUTPUT batch / height / width                                       */
  int32_t in_H, in_W;    /* input height / width (== H, W for stride 1; 2H, 2W for stride 2)    */
  int32_t cin;           /* input channels, multiple of 16                                      */
  int32_t cout;          /* real output channels                                                */
  int32_t taps;          /* 1 (1x1) or 9 (3x3, pad 1)                                           */
  int32_t stride;        /* 1 or 2 (2 requires HRNB_CONV_GATHER)                                */
  int32_t KC;            /* planes (8-channel groups) per K chunk: even, divides cin/8          */
  int32_t BN;            /* N tile: multiple of 16, <= 256                                      */
  int32_t MB;            /* 128-row M blocks per CTA (1, 2 or 4); MB*BN <= 512                  */
  int32_t flags;
  /* Phase-split tensors: a [N,C,H,W] tensor stored as 4 PF8 tensors of [N,C,H/2,W/2], phase (a,b) holding
   * X[2y+a, 2x+b], tensor index 2a+b, consecutive tensors `*_phase_stride` ELEMENTS apart, same plane stride.
   * A 3x3 stride-2 pad-1 conv then is a flat-shift conv over the 4 phases (tap (r,s) reads phase (r!=1, s!=1)
   * at offset (-(r==0), -(s==0))), so it runs on the bulk-copy path with ~1.3x instead of 18x input reads. */
  int64_t in_phase_stride;   /* HRNB_CONV_IN_PHASES: `in` is phase (0,0); in_ps is the phases' plane stride */
  int64_t out_phase_stride;  /* HRNB_CONV_OUT_PHASES: `out` is phase (0,0); out_ps the phases' plane stride  */
  void* out2;                /* optional: ALSO write the output phase-split here (NULL = off)               */
  int64_t out2_ps;
  int64_t out2_phase_stride;
  /* Custom tap table (flat-shift path only; ntap_custom == 0 = the standard 1x1 / 3x3 geometry above).
   * The conv becomes out[p] = sum_t W_t * in_{tap_src[t]}[p + tap_dpos[t]] on the OUTPUT position grid
   * (stride must be 1, taps == ntap_custom, 1..9); source s is the tensor `in + s*in_phase_stride`.
   * This is how the data-gradient of a 3x3 stride-2 conv is expressed per input phase (transposed weights
   * packed with hrnb_pack_conv_weights_batch). */
  int32_t ntap_custom;
  int32_t tap_src[9];
  int32_t tap_dpos[9];
  /* Fused BatchNorm batch statistics (training path; NULL = off).  The epilogue also writes
   * stats_sums[2*c] = sum_p out[p,c], stats_sums[2*c+1] = sum_p out[p,c]^2 over the real positions, from the fp32
   * accumulators, in a fixed summation order (bit-reproducible) - the input of hrnb_bn_apply / hrnb_bn_bwd_*, i.e. it
   * replaces the hrnb_bn_stats pass of nn.BatchNorm2d(train) lib/models/pose_hrnet.py:36.  Needs the flat-shift PF8
   * path with BN == cout in {16, 32, 64}, no residual, no ReLU.  stats_ws: hrnb_conv_stats_ws_floats() floats, zeroed
   * once, not shared between launches that may run concurrently. */
  float* stats_sums;
  float* stats_ws;
  /* Fuse-layer sum inside the epilogue (inference; replaces the summation loop + nn.Upsample(nearest) + ReLU of
   * HighResolutionModule.forward, lib/models/pose_hrnet.py:199-207,257-266): after bias and `res`, the epilogue adds nfuse
   * (0..3) further PF8 tensors, source f given on the grid [N, H >> fuse_shift[f], W >> fuse_shift[f]] and read with nearest
   * up-sampling (fuse_shift 0: same grid, e.g. the output of another stride-2 chain).  Flat-shift or gather path with a
   * plain PF8 output; no fused statistics. */
  int32_t nfuse;
  int32_t fuse_shift[3];
  const void* fuse_src[3];
  int64_t fuse_ps[3];
  /* HRNB_CONV_GATHER 1x1 convs only: the input tensor lives on the grid [N, H >> in_up_shift, W >> in_up_shift] and is read
   * with nearest up-sampling, i.e. the conv runs on the up-sampled grid (1x1 conv and nearest up-sampling commute).  This is
   * how the fuse sum of the highest-resolution branch, which has no stride-2 chain of its own, gets a host convolution.
   * in_H / in_W are then H >> in_up_shift / W >> in_up_shift. */
  int32_t in_up_shift;
  /* ABI 5: GROUPED launch (flat-shift path, lean PF8 epilogue): ngroup (0 / 1 = off, else 2..4) convolutions over the SAME
   * input with the same cin / cout / tile shape, each with its own custom tap table (<= 4 taps, source 0), packed weights
   * and output (residual) tensor `out + g*grp_out_stride` (`res + g*grp_res_stride`) elements - the data gradient of a
   * 3x3 stride-2 conv, whose four input phases each receive 1, 2, 2 and 4 taps, in ONE launch instead of four.  The bias
   * vector is shared.  taps / ntap_custom / tap_* / wpk describe group 0 as for a single launch. */
  int32_t ngroup;
  int32_t grp_ntap[4];
  int32_t grp_tap_dpos[4][4];
  const void* grp_wpk[4];
  int64_t grp_out_stride;
  int64_t grp_res_stride;
} hrnb_conv_params;

int hrnb_conv(const hrnb_conv_params* p, void* stream);

/* dynamic shared memory (bytes) hrnb_conv will request for these parameters; < 0 on invalid params */
int64_t hrnb_conv_smem_bytes(const hrnb_conv_params* p);

/* floats of workspace a launch with stats_sums != NULL needs (first 32 floats = ticket counter, zero-initialised) */
int64_t hrnb_conv_stats_ws_floats(void);

/* Pack OIHW fp32 conv weights (device) into the tile order hrnb_conv streams, folding a per-output
 * channel scale (BN gamma/sqrt(var+eps), or NULL for 1).  out must hold
 * n_tiles*BN * taps*cin bf16 elements, n_tiles = ceil(cout/BN).  bias_out[n_tiles*BN] =
 * shift (NULL -> 0), zero padded. Replaces nothing in the reference (it is the inference-time BN
 * fold implied by lib/models/pose_hrnet.py:36,47 in eval mode). */
int hrnb_pack_conv_weights(const float* w_oihw, const float* scale, const float* shift, int32_t cout,
                           int32_t cin, int32_t taps, int32_t KC, int32_t BN, void* wpk_out,
                           float* bias_out, void* stream);

/* ---- stem conv1: 3 -> 64, 3x3 stride 2 pad 1, NCHW fp32 in -> PF8 bf16 out -------------------- */
/* Replaces conv1+bn1+relu, lib/models/pose_hrnet.py:283-285,512-514. w: [64][27] fp32 (scale folded),
 * bias[64]. */
int hrnb_stem_conv1(const float* x_nchw, const float* w, const float* bias, void* out, int64_t out_ps,
                    int32_t N, int32_t in_H, int32_t in_W, void* stream);

/* NCHW fp32 [N,3,H,W] -> PF8 bf16 with 32 channels at (H/2, W/2): channel k = ci*9 + r*3 + s holds the input value
 * tap (r,s) of the 3x3 stride-2 pad-1 stem conv would read (27..31 = 0), so that conv1 (lib/models/pose_hrnet.py:283)
 * becomes a 1x1 conv on the tensor pipe. */
int hrnb_stem_im2col(const float* x_nchw, void* out, int64_t out_ps, int32_t N, int32_t in_H, int32_t in_W,
                     void* stream);

/* PF8 [N,C,H,W] -> 4 phase tensors [N,C,H/2,W/2] (dst = phase (0,0), phase_stride elements apart); the phases'
 * padding must be zero on entry and is not written. */
int hrnb_phase_split(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                     int64_t dst_ps, int64_t phase_stride, void* stream);

/* ---- fuse-layer sum: out = ReLU(sum_k src_k[nearest-upsample by 2^shift_k]) -------------------- */
/* Replaces HighResolutionModule.forward's summation loop + nn.Upsample(nearest) + ReLU,
 * lib/models/pose_hrnet.py:199-207,257-266. All tensors PF8; src k has spatial (H>>shift, W>>shift). */
typedef struct hrnb_fuse_params {
  const void* src[4];
  int64_t src_ps[4];
  int32_t shift[4];
  int32_t nsrc;
  void* out;
  int64_t out_ps;
  int32_t N, H, W, C;
  int32_t relu;
} hrnb_fuse_params;
int hrnb_fuse_sum(const hrnb_fuse_params* p, void* stream);

/* ---- head: bilinear upsample of one branch into the concat buffer (PF8 -> PF8) ---------------- */
/* Replaces F.interpolate(..., mode='bilinear', align_corners=True) lib/models/pose_hrnet_softmax.py:500-502
 * (align_corners=1) and F.upsample(..., mode='bilinear') lib/models/pose_hrnet.py:561-563
 * (align_corners=0), plus the torch.cat at :504 / :565 (dst is the concat buffer's plane offset). */
int hrnb_bilinear_up(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t sH, int32_t sW, void* dst,
                     int64_t dst_ps, int32_t dH, int32_t dW, int32_t align_corners, void* stream);

/* ---- layout conversion ------------------------------------------------------------------------ */
int hrnb_pf8_to_nchw_f32(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W, float* dst,
                         void* stream);
int hrnb_nchw_f32_to_pf8(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, int64_t dst_ps,
                         void* stream);

/* ---- heatmap decode ---------------------------------------------------------------------------- */
/* argmax decode. row_stride_mode 0: x = idx % w, y = idx / w (core/inference.py:18-46, get_max_preds);
 * 1: x = idx % h, y = idx / h (utils/heatmap_decoding.py:102-107, the reference's H-as-stride form).
 * mask_nonpositive != 0 zeroes coords whose max <= 0 (core/inference.py:41-45). First maximum wins.
 * preds [B*J*2] fp32 (x, y), maxvals [B*J] fp32 (may be NULL), idx_out [B*J] int64 (may be NULL). */
int hrnb_decode_argmax(const float* hm, int32_t BJ, int32_t h, int32_t w, int32_t row_stride_mode,
                       int32_t mask_nonpositive, float* preds, float* maxvals, int64_t* idx_out, void* stream);

/* Spatial softmax of logits*temp over h*w (lib/models/pose_hrnet_softmax.py:521-524) fused with the
 * integral soft-argmax E[x], E[y] on pixel grids (kornia spatial_expectation2d(normalized_coordinates=False)
 * called at lib/utils/heatmap_decoding.py:100).  heat_out (NULL to skip) receives the softmax map,
 * coords [BJ*2] (NULL to skip).  temp_dev points at the scalar temperature on the device. */
int hrnb_softmax_softargmax(const float* logits, const float* temp_dev, int32_t BJ, int32_t h, int32_t w,
                            float* heat_out, float* coords, void* stream);
/* soft-argmax alone on an already normalised map (no renormalisation, as kornia). */
int hrnb_softargmax(const float* hm, int32_t BJ, int32_t h, int32_t w, float* coords, void* stream);
/* backward of hrnb_softmax_softargmax: given d_heat (NULL = 0) and d_coords (NULL = 0) produce d_logits and
 * accumulate d_temp (one float, pre-zeroed by caller; NULL to skip). */
int hrnb_softmax_softargmax_bwd(const float* logits, const float* temp_dev, const float* heat, const float* d_heat,
                                const float* d_coords, int32_t BJ, int32_t h, int32_t w, float* d_logits,
                                float* d_temp, void* stream);

/* core/inference.py:49-85 get_final_preds: argmax (masked) + optional quarter-pixel shift + inverse affine
 * (utils/transforms.py:50-96 with rot = 0: x' = (x - w/2) * (scale_x*200/w) + cx, likewise y with the SAME
 * factor scale_x*200/w since the 3-point affine is a similarity built from src_w/dst_w only).
 * center, scale: [B*2] fp32.  preds [B*J*2], maxvals [B*J]. */
int hrnb_final_preds(const float* hm, int32_t B, int32_t J, int32_t h, int32_t w, const float* center,
                     const float* scale, int32_t post_process, float* preds, float* maxvals, void* stream);

/* ---- losses ------------------------------------------------------------------------------------ */
/* HeatmapLoss (core/loss.py:15-28): sum over h*w of (pred-gt)^2 (mode 0) or |pred-gt| (mode 1), mean over BJ.
 * loss: one float (written, not accumulated).  d_pred (NULL to skip) = dloss/dpred * (*grad_scale_dev or 1). */
int hrnb_loss_heatmap(const float* pred, const float* gt, int32_t BJ, int32_t hw, int32_t mode, float* loss,
                      float* d_pred, const float* grad_scale_dev, float* partial_ws, void* stream);
/* JointsMSELoss (core/loss.py:30-50): sum(||pred-gt||_2 * vis) / max(1, sum(vis)), or sum(||.||)/J when vis
 * is NULL.  d_pred (NULL to skip) [B*J*2]. */
int hrnb_loss_pose2d(const float* pred, const float* gt, const float* vis, int32_t B, int32_t J, float* loss,
                     float* d_pred, void* stream);


/* =============================== training path ================================================= */
/* The reference trains through autograd (loss.backward() + optimizer.step(), lib/core/function.py:101-106).
 * The functions below are the hand-written forward (batch-statistics BN) and backward of the same layers. */

/* ---- weight gradient of a conv (tcgen05, both operands straight from PF8) ---------------------- */
/* dw[tap_id[t]][ci][co] += sum_p dy[p][co] * x[p + tap_dpos[t]][ci]   for t < ntap  (fp32, accumulated with
 * red.global.add: zero dw before the first launch of a step).  The gradient layout is [taps_total][cin][cout];
 * hrnb_adam_step / hrnb_grad_to_natural read it through hrnb_param_seg.  dy: PF8 with ceil(cout/8) planes on the
 * conv's OUTPUT grid [N,H,W] (zero padding / guards as always); x: PF8 input on the SAME grid: the conv input itself
 * for stride 1 (3x3: tap_dpos = (r-1)*(W+1) + (s-1)), one phase tensor of it for stride-2 convs (tap (r,s) of a
 * 3x3 stride-2 conv reads phase (r != 1, s != 1) at dpos = -(r == 0)*(W+1) - (s == 0): one launch per phase, or one launch
 * for all phases with tap_src / x_src_stride).
 * Replaces the weight-gradient half of nn.Conv2d's backward for lib/models/pose_hrnet.py:28-98,187-242,335-350,419-458. */
typedef struct hrnb_wgrad_params {
  const void* dy;
  int64_t dy_ps;
  const void* x;
  int64_t x_ps;
  float* dw;
  int32_t N, H, W;
  int32_t cin;           /* multiple of 16 (gradient rows per tap)                                  */
  int32_t cout;          /* real output channels (row length of dw)                                 */
  int32_t ntap;
  int32_t tap_dpos[9];
  int32_t tap_id[9];
  int32_t NT, TG, KP, ksplit; /* tile overrides: cin tile, taps per CTA, positions per stage, K splits; 0 = auto */
  /* ABI 5: taps over several input tensors in ONE launch (3x3 stride-2 conv over a phase-split input: all four phases).
   * Tap t reads the tensor at x + tap_src[t] * x_src_stride elements (plane stride x_ps for every source); taps of one source
   * must be consecutive in the list and form one CTA tap group (<= 4 taps per source).  x_src_stride == 0: single source. */
  int32_t tap_src[9];
  int32_t pad_;
  int64_t x_src_stride;
} hrnb_wgrad_params;
int hrnb_wgrad(const hrnb_wgrad_params* p, void* stream);
int64_t hrnb_wgrad_smem_bytes(const hrnb_wgrad_params* p);

/* ---- batched weight packing (forward and data-gradient orientation) ---------------------------- */
/* One job = one packed copy of one OIHW fp32 weight tensor w[cout][cin][taps_total].  The logical conv it feeds has
 * lcout x lcin channels (lcin a multiple of 16, zero padded) and ntap taps, logical tap t = source tap tap_ids[t];
 * transpose != 0 swaps the channel roles (logical cout = source cin): with tap_ids reversed this is the
 * data-gradient conv of a stride-1 3x3 conv.  Output order as hrnb_pack_conv_weights ([ntile][chunk][tap][KC][BN][8]).
 * `jobs` and `block_job` (job index of every 256-thread block; job j owns blocks [block0, block0 + ceil(pairs/256)) with
 * pairs = ntiles*BN*lcin: one thread packs all ntap taps of one (output channel, input channel) pair)
 * live in DEVICE memory.  All pointers inside a job are device pointers. */
typedef struct hrnb_pack_job {
  const void* w;
  const void* scale;     /* per source-cout scale or NULL                                            */
  const void* shift;     /* bias source [lcout] or NULL (-> zeros)                                   */
  void* wpk_out;
  void* bias_out;        /* [ntiles*BN] fp32 or NULL                                                 */
  int32_t cout, cin, taps_total;
  int32_t transpose;
  int32_t lcout, lcin;
  int32_t ntap;
  int32_t tap_ids[9];
  int32_t KC, BN;
  int32_t block0;
  int32_t pad_;
} hrnb_pack_job;
int hrnb_pack_conv_weights_batch(const hrnb_pack_job* jobs_dev, const int32_t* block_job_dev, int32_t nblocks,
                                 void* stream);

/* ---- BatchNorm2d, train mode (momentum / eps as nn.BatchNorm2d) -------------------------------- */
/* Reductions over positions are DETERMINISTIC: blocks store partial sums into the workspace `ws` and the last block
 * of a channel plane adds them in block order (no fp32 atomics, whose ordering noise a deep BatchNorm network
 * amplifies chaotically).  ws: hrnb_reduce_ws_floats() floats, zero-initialised once by the caller; calls sharing a
 * workspace must be ordered on one stream. */
int64_t hrnb_reduce_ws_floats(void);
/* sums[c][0] = sum x, sums[c][1] = sum x^2 over the N*H*W real positions of PF8 tensor c. */
int hrnb_bn_stats(const void* c, int64_t c_ps, int32_t N, int32_t C, int32_t H, int32_t W, float* sums, float* ws,
                  void* stream);
/* out[c] = sum over positions, c < C (bias gradient of the BN-less final conv). */
int hrnb_channel_sum(const void* c, int64_t c_ps, int32_t N, int32_t C, int32_t H, int32_t W, float* out, float* ws,
                     void* stream);

typedef struct hrnb_bn_params {
  const void* c;          /* conv output (pre-BN), PF8                                               */
  int64_t c_ps;
  const float* sums;      /* [C][2] from hrnb_bn_stats                                                */
  const float* gamma;     /* [C]                                                                      */
  const float* beta;      /* [C]                                                                      */
  const void* res;        /* optional residual added after the normalisation, PF8 (NULL = none)       */
  int64_t res_ps;
  void* out;              /* PF8: [relu](bn(c) + res), zeros at padding                               */
  int64_t out_ps;
  float* running_mean;    /* updated in place with the batch mean / unbiased variance (NULL = skip)   */
  float* running_var;
  int32_t N, C, H, W;
  int32_t relu;
  float eps, momentum;
  /* ABI 5: optional second, phase-split copy of `out` (the input format of a following 3x3 stride-2 conv, see
   * hrnb_conv_params): saves the hrnb_phase_split pass over the unit output.  NULL = off; H and W must be even. */
  void* out2;
  int64_t out2_ps;
  int64_t out2_phase_stride;
} hrnb_bn_params;
int hrnb_bn_apply(const hrnb_bn_params* p, void* stream);
/* Horizontally batched form: statistics (written to p[j].sums) + normalisation of n <= 4 independent tensors (the
 * branches of a HighResolutionModule at the same depth) in two launches instead of 2n.  Bit j of have_stats_mask set:
 * p[j].sums already holds the statistics (hrnb_conv_params.stats_sums), tensor j is left out of the statistics launch. */
int hrnb_bn_forward_batch(const hrnb_bn_params* p, int32_t n, float* ws, int32_t have_stats_mask, void* stream);

typedef struct hrnb_bn_bwd_params {
  const void* dy;         /* gradient of the unit output, PF8                                         */
  int64_t dy_ps;
  const void* y;          /* unit output (ReLU mask y > 0); NULL when relu == 0                       */
  int64_t y_ps;
  const void* c;          /* conv output saved by the forward pass                                    */
  int64_t c_ps;
  const float* sums;      /* forward statistics [C][2]                                                */
  const float* gamma;
  float* dsums;           /* [C][2] written by hrnb_bn_bwd_reduce (sum g, sum g*xhat)                */
  float* ws;              /* reduction workspace (hrnb_reduce_ws_floats)                              */
  void* dc;               /* gradient w.r.t. the conv output (may alias dy)                           */
  int64_t dc_ps;
  void* dres;             /* gradient buffer of the residual input or NULL                            */
  int64_t dres_ps;
  int32_t dres_mode;      /* 1 = write g, 2 = accumulate g   (g = dy * relu mask)                     */
  float* dgamma;          /* [C] written (sum g*xhat)                                                 */
  float* dbeta;           /* [C] written (sum g)                                                      */
  int32_t N, C, H, W;
  int32_t relu;
  float eps;
} hrnb_bn_bwd_params;
int hrnb_bn_bwd_reduce(const hrnb_bn_bwd_params* p, void* stream);
int hrnb_bn_bwd_apply(const hrnb_bn_bwd_params* p, void* stream);
/* Horizontally batched hrnb_bn_bwd_reduce + hrnb_bn_bwd_apply over n <= 4 independent units (two launches); all
 * p[j].ws must be the same workspace. */
int hrnb_bn_backward_batch(const hrnb_bn_bwd_params* p, int32_t n, void* stream);

/* ---- backward of hrnb_fuse_sum / hrnb_bilinear_up / hrnb_phase_split ---------------------------- */
/* dsrc[q] (=|+=) sum over the 2^shift x 2^shift block of dy * (y > 0) (mode 1 write, 2 accumulate); dy, y on the
 * fused output grid [N,H,W], dsrc on [N, H>>shift, W>>shift]. */
int hrnb_fuse_sum_bwd(const void* dy, int64_t dy_ps, const void* y, int64_t y_ps, void* dsrc, int64_t dsrc_ps, int32_t N,
                      int32_t H, int32_t W, int32_t C, int32_t shift, int32_t relu, int32_t mode, void* stream);
/* ABI 5: the gradients of ALL n <= 4 sources of one fuse output in one launch (source j: dsrc[j] on the grid
 * [N, H >> shift[j], W >> shift[j]], mode[j] 1 = write, 2 = accumulate); dsrc / dsrc_ps / shift / mode are HOST arrays. */
int hrnb_fuse_sum_bwd_batch(const void* dy, int64_t dy_ps, const void* y, int64_t y_ps, int32_t n, void* const* dsrc,
                            const int64_t* dsrc_ps, const int32_t* shift, const int32_t* mode, int32_t N, int32_t H, int32_t W,
                            int32_t C, int32_t relu, void* stream);
int hrnb_bilinear_up_bwd(const void* d_dst, int64_t d_dst_ps, int32_t N, int32_t C, int32_t dH, int32_t dW, void* d_src,
                         int64_t d_src_ps, int32_t sH, int32_t sW, int32_t align_corners, int32_t mode, void* stream);
int hrnb_phase_merge(const void* src_phase00, int64_t src_ps, int64_t phase_stride, void* dst, int64_t dst_ps, int32_t N,
                     int32_t C, int32_t H, int32_t W, int32_t mode, void* stream);

/* ---- fused Adam over flat fp32 parameter / moment buffers (lib/utils/utils.py:71-92: torch.optim.Adam, L2 decay) */
typedef struct hrnb_param_seg {
  int64_t p_off;          /* offset of the tensor in the flat param / m / v buffers (natural layout)  */
  int64_t g_off;          /* offset of its gradient in the flat gradient buffer                       */
  int32_t numel;
  int32_t cout, cin, cin_g, taps; /* conv weights: natural [cout][cin][taps], gradient [taps][cin_g][cout]; taps == 0: 1-D */
  int32_t block0;         /* first 1024-element block of this tensor in the block map                 */
  int32_t frozen;         /* requires_grad == False                                                   */
  int32_t pad_;
} hrnb_param_seg;
/* hyper_dev = [lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, grad_scale] (device floats) */
int hrnb_adam_step(float* params, float* m, float* v, const float* grads, const hrnb_param_seg* segs_dev,
                   const int32_t* block_seg_dev, int32_t nblocks, const float* hyper_dev, void* stream);
/* step += 1 and refresh bias_corr1/2 in hyper_dev */
int hrnb_adam_tick(float* hyper_dev, int64_t* step_dev, void* stream);
/* gradient buffer ([tap][cin][cout] conv layout) -> natural parameter layout (for .grad of the nn.Module) */
int hrnb_grad_to_natural(const float* grads, float* out, const hrnb_param_seg* segs_dev, const int32_t* block_seg_dev,
                         int32_t nblocks, void* stream);

/* ---- algebraic (DLT) triangulation: SURVEY §8 row (f), staged in round 1 ------------------------- */
/* Replaces the per-joint loop of AlgebraicTriangulationNet.forward (lib/models/triangulation.py:258-261) over
 * DLT_sii_pytorch (lib/utils/misc.py:64-97) + homogeneous_to_euclidean (lib/utils/misc.py:28-35) with one launch.
 * points [B][V][J][2] image coordinates, proj [B][V][3][4] projection matrices, bk0 [J][B][4] unit start vectors of the
 * shifted inverse iteration (the reference draws torch.rand(B,4,1) per joint on the host and normalises it),
 * out [B][J][3] euclidean 3-D joints; all fp32 device pointers. V >= 2, iterations >= 1 (reference: 2). */
int hrnb_triangulate_dlt(const float* points, const float* proj, const float* bk0, int32_t B, int32_t V, int32_t J,
                         int32_t iterations, float* out, void* stream);
/* Adjoint of hrnb_triangulate_dlt w.r.t. the 2-D points (the reference's DLT_sii_pytorch is a differentiable torch graph;
 * lib/core/function.py train3D back-propagates the 3-D joint loss through it into the backbone).  d_out [B][J][3] ->
 * d_points [B][V][J][2] (written).  The projection matrices are data (no gradient).  iterations <= 4. */
int hrnb_triangulate_dlt_bwd(const float* points, const float* proj, const float* bk0, const float* d_out, int32_t B,
                             int32_t V, int32_t J, int32_t iterations, float* d_points, void* stream);

/* ---- loop glue on the device: SURVEY §8 rows (f1) confidence head, (f3) targets + normalisation, (f4) flip test ------ */
/* Ground-truth heat maps, replaces HeatmapGenerator.__call__ (lib/dataset/target_generators/target_generators.py:28-53) run per
 * sample in the data-loader workers: joints [BJ][joint_stride] fp32 rows (u, v[, visible]) in heat-map pixels, out [BJ][h][w]
 * fp32.  x = int(u), y = int(v); joints with visible <= 0 or outside the map give an all-zero map; otherwise the
 * (6 sigma + 3)^2 patch around (x, y) holds exp(-d^2 / (2 sigma^2)) evaluated in float64 (as numpy) and the rest is zero. */
int hrnb_gen_heatmaps(const float* joints, int32_t joint_stride, int32_t BJ, int32_t h, int32_t w, float sigma, float* out,
                      void* stream);
/* uint8 NHWC image [N][in_H][in_W][3] (device) -> the stem's im2col slab (as hrnb_stem_im2col) with torchvision's ToTensor
 * (/255) and Normalize(mean, std) applied on the fly (lib/dataset/transforms/build.py:82-85): the fp32 NCHW image tensor
 * the reference uploads (4x the bytes) never exists.  mean3_host / std3_host: 3 floats each in HOST memory. */
int hrnb_stem_im2col_u8(const uint8_t* img_nhwc, const float* mean3_host, const float* std3_host, void* out, int64_t out_ps,
                        int32_t N, int32_t in_H, int32_t in_W, void* stream);
/* Flip test merge, replaces flip_back (lib/utils/transforms.py:16-30: reverse x, swap the matched joint pairs) + the
 * SHIFT_HEATMAP one-pixel shift + the average of lib/core/function.py:681-701, which the reference does through a
 * device->host->device round trip in numpy.  out = 0.5 * (hm + shift(flip_back(hm_flipped))), or shift(flip_back(hm_flipped))
 * alone when hm is NULL; perm_dev [J] int32 = the joint permutation of the pair swaps; all maps [B][J][h][w] fp32. */
int hrnb_flip_merge(const float* hm, const float* hm_flipped, const int32_t* perm_dev, int32_t B, int32_t J, int32_t h, int32_t w,
                    int32_t shift, float* out, void* stream);
/* GlobalAveragePoolingHead (lib/models/pose_hrnet_volumetric.py:22-56): nn.MaxPool2d(2) + ReLU on PF8 (the conv + BatchNorm
 * in front of it is an hrnb_conv launch), and mean over positions + Linear/ReLU + Linear/ReLU + Linear/Sigmoid -> out [N][NC].
 * Linear weights row-major [out][in] fp32 as nn.Linear stores them. */
int hrnb_maxpool2_relu(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, int64_t dst_ps,
                       void* stream);
int hrnb_gap_mlp(const void* x, int64_t x_ps, int32_t N, int32_t C, int32_t H, int32_t W, const float* w1, const float* b1,
                 int32_t H1, const float* w2, const float* b2, int32_t H2, const float* w3, const float* b3, int32_t NC,
                 float* out, void* stream);

/* Cross-view fusion: out = a * x + b * y over n fp32 values - Aggregation.fuse_with_weights
 * (lib/models/multiview_pose_hrnet.py:51-55) after the ChannelWiseFC GEMMs (:15-29), which run as hrnb_conv launches with
 * the flattened heat map as the channel axis. */
int hrnb_axpby(float a, const float* x, float b, const float* y, float* out, int64_t n, void* stream);

/* ---- misc --------------------------------------------------------------------------------------- */
const char* hrnb_last_error(void);
int hrnb_abi_version(void);
/* number of kernel launches issued through this library by the calling process (for bench gpu_launches) */
int64_t hrnb_launch_count(void);
/* Hang diagnostics.  Every mbarrier wait of the tcgen05 kernels is bounded (4 s): a protocol bug traps instead of hanging
 * the device.  hrnb_hang_init() (once per device, outside stream capture) arms a mapped host buffer into which the warps
 * that time out write (kernel, CTA, warp, barrier) records before trapping; hrnb_hang_report() copies up to n 64-bit words
 * of it (word 0 != 0: a time-out happened; records from word 2, two words each: blockDim.x << 48 | gridDim.x << 32 |
 * blockIdx.x << 8 | warp, barrier shared-memory address << 8 | parity) and works after the context died. */
int hrnb_hang_init(void);
int hrnb_hang_report(uint64_t* out_host, int n);
/* debug knobs (key 0: exchange the LBO/SBO roles of the UMMA descriptors); not part of the product API */
int hrnb_debug_set(int key, int value);
/* debug: device buffer of 5*64 int64 receiving clock64 timestamps of CTA 0's roles (NULL = off) */
int hrnb_debug_trace(void* dev_buf_5x64_i64);

#ifdef __cplusplus
}
#endif
#endif /* HRNB_H_ */
