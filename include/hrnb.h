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

#define HRNB_ABI_VERSION 1

/* guard bands (in positions) a PF8 plane must carry around [0, P) */
#define HRNB_GUARD_LEAD(Wp) ((((Wp) + 2) + 7) / 8 * 8)
#define HRNB_GUARD_TAIL(Wp) (((((Wp) + 2) + 7) / 8 * 8) + 512)

/* ---- convolution (implicit GEMM on tcgen05 / TMEM) ------------------------------------------ */

enum {
  HRNB_CONV_RELU = 1,       /* ReLU after bias (+ residual)                                   */
  HRNB_CONV_OUT_NCHW = 2,   /* write fp32 NCHW [N, cout_real, H, W] instead of PF8 bf16       */
  HRNB_CONV_GATHER = 4,     /* per-tap gathered A operand (any stride); otherwise flat-shift      */
  HRNB_CONV_IN_PHASES = 8,  /* 3x3 stride-2 conv whose input is given as 4 phase tensors (see below) */
  HRNB_CONV_OUT_PHASES = 16 /* write the PF8 output as 4 phase tensors for a following stride-2 conv */
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
} hrnb_conv_params;

int hrnb_conv(const hrnb_conv_params* p, void* stream);

/* dynamic shared memory (bytes) hrnb_conv will request for these parameters; < 0 on invalid params */
int64_t hrnb_conv_smem_bytes(const hrnb_conv_params* p);

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

/* ---- misc --------------------------------------------------------------------------------------- */
const char* hrnb_last_error(void);
int hrnb_abi_version(void);
/* number of kernel launches issued through this library by the calling process (for bench gpu_launches) */
int64_t hrnb_launch_count(void);
/* debug knobs (key 0: exchange the LBO/SBO roles of the UMMA descriptors); not part of the product API */
int hrnb_debug_set(int key, int value);
/* debug: device buffer of 5*64 int64 receiving clock64 timestamps of CTA 0's roles (NULL = off) */
int hrnb_debug_trace(void* dev_buf_5x64_i64);

#ifdef __cplusplus
}
#endif
#endif /* HRNB_H_ */
