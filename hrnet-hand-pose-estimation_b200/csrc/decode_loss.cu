// Heat-map decode and loss kernels: single-pass, one CTA per (image, joint) map, float4 coalesced loads,
// warp-shuffle + shared-memory reductions.  HBM-bound; algorithmic bytes per map = h*w*4 read
// (+ h*w*4 written when the softmax map is materialised).
//
// Reference semantics (file:line relative to the reference repo):
//   argmax            lib/core/inference.py:18-46  /  lib/utils/heatmap_decoding.py:102-107
//   softmax           lib/models/pose_hrnet_softmax.py:521-524
//   soft-argmax       kornia spatial_expectation2d(normalized_coordinates=False), lib/utils/heatmap_decoding.py:100
//   final preds       lib/core/inference.py:49-85 + lib/utils/transforms.py:50-96
//   HeatmapLoss       lib/core/loss.py:15-28        JointsMSELoss  lib/core/loss.py:30-50
#include <math.h>
#include "ptx.cuh"
#include "common.h"

namespace hrnb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- block reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// deterministic (fixed-order) block sum, result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int i = 1; i < kWarps; ++i) t = fmaxf(t, red[i]);
  return t;
}

// first-maximum argmax of one map; every thread returns the block result
__device__ __forceinline__ void block_argmax(const float* __restrict__ m, int hw, float& best_v, int& best_i) {
  __shared__ float sv[kWarps];
  __shared__ int si[kWarps];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const int nvec = hw >> 2;
  const float4* m4 = reinterpret_cast<const float4*>(m);
  // four independent 16-byte loads in flight per thread before the first comparison (a 64 x 64 map is exactly one round)
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * kThreads) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kThreads;
      v[j] = i < nvec ? __ldg(m4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = (i0 + j * kThreads) << 2;
      // strictly-greater keeps the first occurrence inside a thread (indices visited in increasing order)
      if (v[j].x > bv) { bv = v[j].x; bi = b; }
      if (v[j].y > bv) { bv = v[j].y; bi = b + 1; }
      if (v[j].z > bv) { bv = v[j].z; bi = b + 2; }
      if (v[j].w > bv) { bv = v[j].w; bi = b + 3; }
    }
  }
  for (int i = (nvec << 2) + threadIdx.x; i < hw; i += kThreads) {
    const float v = __ldg(m + i);
    if (v > bv) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  bv = sv[0]; bi = si[0];
#pragma unroll
  for (int i = 1; i < kWarps; ++i) {
    if (sv[i] > bv || (sv[i] == bv && si[i] < bi)) { bv = sv[i]; bi = si[i]; }
  }
  if (bi == 0x7fffffff) bi = 0;  // all -inf / NaN map: numpy would still answer 0 for all -inf
  best_v = bv;
  best_i = bi;
}

// ---- D1 / D2: argmax decode -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) decode_argmax_kernel(const float* __restrict__ hm, int hw, int stride_div,
                                                                int mask_nonpositive, float* __restrict__ preds,
                                                                float* __restrict__ maxvals,
                                                                long long* __restrict__ idx_out) {
  const int map = blockIdx.x;
  float bv; int bi;
  block_argmax(hm + (long long)map * hw, hw, bv, bi);
  if (threadIdx.x == 0) {
    float x = (float)(bi % stride_div);
    float y = (float)(bi / stride_div);
    if (mask_nonpositive && !(bv > 0.f)) { x = 0.f; y = 0.f; }
    preds[2 * map] = x;
    preds[2 * map + 1] = y;
    if (maxvals) maxvals[map] = bv;
    if (idx_out) idx_out[map] = bi;
  }
}

// ---- D4: final preds (argmax + quarter-pixel shift + inverse similarity) ---------------------------
__global__ void __launch_bounds__(kThreads) final_preds_kernel(const float* __restrict__ hm, int J, int h, int w,
                                                              const float* __restrict__ center,
                                                              const float* __restrict__ scale, int post_process,
                                                              float* __restrict__ preds, float* __restrict__ maxvals) {
  const int map = blockIdx.x;
  const int b = map / J;
  const float* m = hm + (long long)map * h * w;
  float bv; int bi;
  block_argmax(m, h * w, bv, bi);
  if (threadIdx.x == 0) {
    float cx = (float)(bi % w), cy = (float)(bi / w);
    if (!(bv > 0.f)) { cx = 0.f; cy = 0.f; }
    if (post_process) {
      const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
      if (1 < px && px < w - 1 && 1 < py && py < h - 1) {
        const float dx = m[py * w + px + 1] - m[py * w + px - 1];
        const float dy = m[(py + 1) * w + px] - m[(py - 1) * w + px];
        cx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
        cy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
      }
    }
    // inverse of the 3-point affine with rot = 0: similarity with factor (scale_x*200)/w around the centres
    const double f = ((double)scale[2 * b] * 200.0) / (double)w;
    preds[2 * map] = (float)(((double)cx - 0.5 * (double)w) * f + (double)center[2 * b]);
    preds[2 * map + 1] = (float)(((double)cy - 0.5 * (double)h) * f + (double)center[2 * b + 1]);
    maxvals[map] = bv;
  }
}

// ---- M7 + D3: softmax(logits*temp) + soft-argmax ----------------------------------------------------
// dynamic smem: hw floats (the map is read from HBM exactly once)
__global__ void __launch_bounds__(kThreads) softmax_softargmax_kernel(const float* __restrict__ logits,
                                                                     const float* __restrict__ temp_dev, int h, int w,
                                                                     float* __restrict__ heat_out,
                                                                     float* __restrict__ coords) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[kWarps];
  const int hw = h * w;
  const int map = blockIdx.x;
  const float temp = temp_dev ? __ldg(temp_dev) : 1.f;
  const float* src = logits + (long long)map * hw;
  float mx = -INFINITY;
  const int nvec = hw >> 2;
  for (int i = threadIdx.x; i < nvec; i += kThreads) {
    float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    v.x *= temp; v.y *= temp; v.z *= temp; v.w *= temp;
    reinterpret_cast<float4*>(sm)[i] = v;
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int i = (nvec << 2) + threadIdx.x; i < hw; i += kThreads) {
    const float v = __ldg(src + i) * temp;
    sm[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < hw; i += kThreads) {
    const float e = expf(sm[i] - mx);
    sm[i] = e;
    s += e;
  }
  s = block_sum(s, red);
  float sx = 0.f, sy = 0.f;
  float* dst = heat_out ? heat_out + (long long)map * hw : nullptr;
  for (int i = threadIdx.x; i < hw; i += kThreads) {
    const float p = sm[i] / s;
    if (dst) dst[i] = p;
    const int y = i / w, x = i - y * w;
    sx = fmaf(p, (float)x, sx);
    sy = fmaf(p, (float)y, sy);
  }
  if (coords) {
    sx = block_sum(sx, red);
    sy = block_sum(sy, red);
    if (threadIdx.x == 0) { coords[2 * map] = sx; coords[2 * map + 1] = sy; }
  }
}

// Register form of the kernel above for maps of up to NV * 1024 elements with w % 4 == 0 (64 x 64: NV = 4, 96 x 72: NV = 7): every
// thread keeps its NV float4 in registers across the three phases (max, exp + sum, normalise + expectation), so the map is read
// from HBM once, written once and never staged in shared memory; exp through ex2.approx (__expf, 2 ulp), one reciprocal per map.
template <int NV>
__global__ void __launch_bounds__(kThreads) softmax_softargmax_reg_kernel(const float* __restrict__ logits,
                                                                         const float* __restrict__ temp_dev, int h, int w,
                                                                         float* __restrict__ heat_out,
                                                                         float* __restrict__ coords) {
  __shared__ float red[kWarps];
  const int hw = h * w, nvec = hw >> 2;
  const int map = blockIdx.x;
  const float temp = temp_dev ? __ldg(temp_dev) : 1.f;
  const float4* src = reinterpret_cast<const float4*>(logits + (long long)map * hw);
  float4 v[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      float4 t = __ldg(src + i);
      t.x *= temp; t.y *= temp; t.z *= temp; t.w *= temp;
      v[j] = t;
      mx = fmaxf(fmaxf(mx, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
    } else {
      v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x = __expf(v[j].x - mx); v[j].y = __expf(v[j].y - mx); v[j].z = __expf(v[j].z - mx); v[j].w = __expf(v[j].w - mx);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
  float sx = 0.f, sy = 0.f;
  float4* dst = heat_out ? reinterpret_cast<float4*>(heat_out + (long long)map * hw) : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      const float4 p = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
      if (dst) dst[i] = p;
      const int e = i << 2;
      const int y = e / w, x = e - y * w;          // w % 4 == 0: the four elements share a row
      const float fx = (float)x;
      sx += (p.x * fx + p.y * (fx + 1.f)) + (p.z * (fx + 2.f) + p.w * (fx + 3.f));
      sy = fmaf((p.x + p.y) + (p.z + p.w), (float)y, sy);
    }
  }
  if (coords) {
    sx = block_sum(sx, red);
    sy = block_sum(sy, red);
    if (threadIdx.x == 0) { coords[2 * map] = sx; coords[2 * map + 1] = sy; }
  }
}

__global__ void __launch_bounds__(kThreads) softargmax_kernel(const float* __restrict__ hm, int h, int w,
                                                             float* __restrict__ coords) {
  __shared__ float red[kWarps];
  const int hw = h * w;
  const int map = blockIdx.x;
  const float* src = hm + (long long)map * hw;
  float sx = 0.f, sy = 0.f;
  if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 16-byte loads, four in flight per thread; one row/column split per float4 (its elements share a row)
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const int nvec = hw >> 2;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * kThreads) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kThreads;
        v[j] = i < nvec ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = (i0 + j * kThreads) << 2;
        const int y = e / w, x = e - y * w;
        const float fx = (float)x;
        sx += (v[j].x * fx + v[j].y * (fx + 1.f)) + (v[j].z * (fx + 2.f) + v[j].w * (fx + 3.f));
        sy = fmaf((v[j].x + v[j].y) + (v[j].z + v[j].w), (float)y, sy);
      }
    }
  } else {
    for (int i = threadIdx.x; i < hw; i += kThreads) {
      const float p = __ldg(src + i);
      const int y = i / w, x = i - y * w;
      sx = fmaf(p, (float)x, sx);
      sy = fmaf(p, (float)y, sy);
    }
  }
  sx = block_sum(sx, red);
  sy = block_sum(sy, red);
  if (threadIdx.x == 0) { coords[2 * map] = sx; coords[2 * map + 1] = sy; }
}

// backward: G_i = d_heat_i + dx*x_i + dy*y_i ; dz_i = p_i (G_i - sum_j p_j G_j) ; d_logit = temp*dz ;
// d_temp += sum_i dz_i * logit_i
__global__ void __launch_bounds__(kThreads) softmax_softargmax_bwd_kernel(
    const float* __restrict__ logits, const float* __restrict__ temp_dev, const float* __restrict__ heat,
    const float* __restrict__ d_heat, const float* __restrict__ d_coords, int h, int w, float* __restrict__ d_logits,
    float* __restrict__ d_temp) {
  extern __shared__ __align__(16) float sm[];  // G
  __shared__ float red[kWarps];
  const int hw = h * w;
  const int map = blockIdx.x;
  const long long off = (long long)map * hw;
  const float temp = temp_dev ? __ldg(temp_dev) : 1.f;
  const float dx = d_coords ? d_coords[2 * map] : 0.f;
  const float dy = d_coords ? d_coords[2 * map + 1] : 0.f;
  float dot = 0.f;
  for (int i = threadIdx.x; i < hw; i += kThreads) {
    const int y = i / w, x = i - y * w;
    float G = dx * (float)x + dy * (float)y;
    if (d_heat) G += __ldg(d_heat + off + i);
    sm[i] = G;
    dot = fmaf(__ldg(heat + off + i), G, dot);
  }
  dot = block_sum(dot, red);
  float dt = 0.f;
  for (int i = threadIdx.x; i < hw; i += kThreads) {
    const float dz = __ldg(heat + off + i) * (sm[i] - dot);
    d_logits[off + i] = temp * dz;
    dt = fmaf(dz, __ldg(logits + off + i), dt);
  }
  if (d_temp) {
    dt = block_sum(dt, red);
    if (threadIdx.x == 0) atomicAdd(d_temp, dt);
  }
}

// ---- L1: heat-map loss ------------------------------------------------------------------------------
// stage 1: per-CTA partial sums (+ optional gradient), stage 2: fixed-order final sum -> deterministic.
__global__ void __launch_bounds__(kThreads) heatmap_loss_partial_kernel(const float* __restrict__ pred,
                                                                       const float* __restrict__ gt, long long n,
                                                                       int mode, float inv_bj,
                                                                       const float* __restrict__ grad_scale_dev,
                                                                       float* __restrict__ d_pred,
                                                                       float* __restrict__ partial) {
  __shared__ float red[kWarps];
  const float gs = (grad_scale_dev ? __ldg(grad_scale_dev) : 1.f) * inv_bj;
  float acc = 0.f;
  const long long nvec = n >> 2;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i0 = (long long)blockIdx.x * kThreads + threadIdx.x; i0 < nvec; i0 += 2 * stride) {
    // two independent (pred, gt) pairs in flight per thread
    float4 a[2], b[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long i = i0 + j * stride;
      const bool in = i < nvec;
      a[j] = in ? __ldg(reinterpret_cast<const float4*>(pred) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[j] = in ? __ldg(reinterpret_cast<const float4*>(gt) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long i = i0 + j * stride;
      if (i >= nvec) break;
      const float d0 = a[j].x - b[j].x, d1 = a[j].y - b[j].y, d2 = a[j].z - b[j].z, d3 = a[j].w - b[j].w;
      if (mode == 0) {
        acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        if (d_pred) reinterpret_cast<float4*>(d_pred)[i] = make_float4(2.f * d0 * gs, 2.f * d1 * gs, 2.f * d2 * gs, 2.f * d3 * gs);
      } else {
        acc += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
        if (d_pred) {
          auto sg = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
          reinterpret_cast<float4*>(d_pred)[i] = make_float4(sg(d0) * gs, sg(d1) * gs, sg(d2) * gs, sg(d3) * gs);
        }
      }
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (nvec << 2) + threadIdx.x; i < n; i += kThreads) {
      const float d = pred[i] - gt[i];
      acc += mode == 0 ? d * d : fabsf(d);
      if (d_pred) d_pred[i] = mode == 0 ? 2.f * d * gs : (d > 0.f ? gs : (d < 0.f ? -gs : 0.f));
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(kThreads) heatmap_loss_final_kernel(const float* __restrict__ partial, int nparts,
                                                                     float inv_bj, float* __restrict__ loss) {
  __shared__ float red[kWarps];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += kThreads) acc += partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *loss = acc * inv_bj;
}

// ---- L2: pose2d loss --------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pose2d_loss_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ gt,
                                                              const float* __restrict__ vis, int n, int J,
                                                              float* __restrict__ loss, float* __restrict__ d_pred) {
  __shared__ float red[kWarps];
  float acc = 0.f, vs = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float dx = pred[2 * i] - gt[2 * i], dy = pred[2 * i + 1] - gt[2 * i + 1];
    const float d = sqrtf(dx * dx + dy * dy);
    const float v = vis ? vis[i] : 1.f;
    acc += d * v;
    vs += v;
  }
  acc = block_sum(acc, red);
  vs = block_sum(vs, red);
  const float denom = vis ? fmaxf(1.f, vs) : (float)J;
  if (threadIdx.x == 0) *loss = acc / denom;
  if (d_pred) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const float dx = pred[2 * i] - gt[2 * i], dy = pred[2 * i + 1] - gt[2 * i + 1];
      const float d = sqrtf(dx * dx + dy * dy);
      const float v = vis ? vis[i] : 1.f;
      const float c = d > 0.f ? v / (d * denom) : 0.f;  // torch.norm backward yields 0 at zero distance
      d_pred[2 * i] = dx * c;
      d_pred[2 * i + 1] = dy * c;
    }
  }
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_decode_argmax(const float* hm, int32_t BJ, int32_t h, int32_t w, int32_t row_stride_mode,
                                  int32_t mask_nonpositive, float* preds, float* maxvals, int64_t* idx_out,
                                  void* stream) {
  if (!hm || !preds) return fail(HRNB_EINVAL, "decode_argmax: null pointer");
  if (BJ < 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "decode_argmax: bad shape");
  if (BJ == 0) return HRNB_OK;
  if ((reinterpret_cast<uintptr_t>(hm) & 15) || ((h * w) & 3)) return fail(HRNB_EINVAL, "decode_argmax: maps must be 16-byte aligned with h*w % 4 == 0");
  decode_argmax_kernel<<<BJ, kThreads, 0, (cudaStream_t)stream>>>(hm, h * w, row_stride_mode ? h : w, mask_nonpositive,
                                                                 preds, maxvals, (long long*)idx_out);
  count_launch();
  return check_launch("decode_argmax_kernel");
}

extern "C" int hrnb_final_preds(const float* hm, int32_t B, int32_t J, int32_t h, int32_t w, const float* center,
                                const float* scale, int32_t post_process, float* preds, float* maxvals, void* stream) {
  if (!hm || !center || !scale || !preds || !maxvals) return fail(HRNB_EINVAL, "final_preds: null pointer");
  if (B < 0 || J <= 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "final_preds: bad shape");
  if (B == 0) return HRNB_OK;
  if ((reinterpret_cast<uintptr_t>(hm) & 15) || ((h * w) & 3)) return fail(HRNB_EINVAL, "final_preds: maps must be 16-byte aligned with h*w % 4 == 0");
  final_preds_kernel<<<B * J, kThreads, 0, (cudaStream_t)stream>>>(hm, J, h, w, center, scale, post_process, preds, maxvals);
  count_launch();
  return check_launch("final_preds_kernel");
}

static int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
  }
  return HRNB_OK;
}

extern "C" int hrnb_softmax_softargmax(const float* logits, const float* temp_dev, int32_t BJ, int32_t h, int32_t w,
                                       float* heat_out, float* coords, void* stream) {
  if (!logits) return fail(HRNB_EINVAL, "softmax_softargmax: null pointer");
  if (BJ < 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "softmax_softargmax: bad shape");
  if (BJ == 0) return HRNB_OK;
  const size_t smem = (size_t)h * w * sizeof(float);
  if (smem > 200 * 1024) return fail(HRNB_EINVAL, "softmax_softargmax: map too large for shared memory");
  if ((reinterpret_cast<uintptr_t>(logits) & 15) || ((h * w) & 3)) return fail(HRNB_EINVAL, "softmax_softargmax: maps must be 16-byte aligned with h*w % 4 == 0");
  const int nvec = (h * w) >> 2;
  if ((w & 3) == 0 && nvec <= 8 * kThreads && (heat_out == nullptr || (reinterpret_cast<uintptr_t>(heat_out) & 15) == 0)) {
    // register-resident form: 64 x 64 maps (4 float4 per thread), 96 x 72 (7), up to 8192 elements (8)
    const int nv = (nvec + kThreads - 1) / kThreads;
    cudaStream_t st = (cudaStream_t)stream;
    if (nv <= 4) softmax_softargmax_reg_kernel<4><<<BJ, kThreads, 0, st>>>(logits, temp_dev, h, w, heat_out, coords);
    else if (nv <= 7) softmax_softargmax_reg_kernel<7><<<BJ, kThreads, 0, st>>>(logits, temp_dev, h, w, heat_out, coords);
    else softmax_softargmax_reg_kernel<8><<<BJ, kThreads, 0, st>>>(logits, temp_dev, h, w, heat_out, coords);
    count_launch();
    return check_launch("softmax_softargmax_reg_kernel");
  }
  int rc = set_smem((const void*)softmax_softargmax_kernel, smem);
  if (rc) return rc;
  softmax_softargmax_kernel<<<BJ, kThreads, smem, (cudaStream_t)stream>>>(logits, temp_dev, h, w, heat_out, coords);
  count_launch();
  return check_launch("softmax_softargmax_kernel");
}

extern "C" int hrnb_softargmax(const float* hm, int32_t BJ, int32_t h, int32_t w, float* coords, void* stream) {
  if (!hm || !coords) return fail(HRNB_EINVAL, "softargmax: null pointer");
  if (BJ < 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "softargmax: bad shape");
  if (BJ == 0) return HRNB_OK;
  softargmax_kernel<<<BJ, kThreads, 0, (cudaStream_t)stream>>>(hm, h, w, coords);
  count_launch();
  return check_launch("softargmax_kernel");
}

extern "C" int hrnb_softmax_softargmax_bwd(const float* logits, const float* temp_dev, const float* heat,
                                           const float* d_heat, const float* d_coords, int32_t BJ, int32_t h, int32_t w,
                                           float* d_logits, float* d_temp, void* stream) {
  if (!logits || !heat || !d_logits) return fail(HRNB_EINVAL, "softmax_bwd: null pointer");
  if (BJ < 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "softmax_bwd: bad shape");
  if (BJ == 0) return HRNB_OK;
  const size_t smem = (size_t)h * w * sizeof(float);
  if (smem > 200 * 1024) return fail(HRNB_EINVAL, "softmax_bwd: map too large for shared memory");
  int rc = set_smem((const void*)softmax_softargmax_bwd_kernel, smem);
  if (rc) return rc;
  softmax_softargmax_bwd_kernel<<<BJ, kThreads, smem, (cudaStream_t)stream>>>(logits, temp_dev, heat, d_heat, d_coords, h,
                                                                             w, d_logits, d_temp);
  count_launch();
  return check_launch("softmax_softargmax_bwd_kernel");
}

extern "C" int hrnb_loss_heatmap(const float* pred, const float* gt, int32_t BJ, int32_t hw, int32_t mode, float* loss,
                                 float* d_pred, const float* grad_scale_dev, float* partial_ws, void* stream) {
  if (!pred || !gt || !loss || !partial_ws) return fail(HRNB_EINVAL, "loss_heatmap: null pointer");
  if (BJ <= 0 || hw <= 0 || (mode != 0 && mode != 1)) return fail(HRNB_EINVAL, "loss_heatmap: bad shape/mode");
  if ((reinterpret_cast<uintptr_t>(pred) & 15) || (reinterpret_cast<uintptr_t>(gt) & 15) ||
      (d_pred && (reinterpret_cast<uintptr_t>(d_pred) & 15)))
    return fail(HRNB_EINVAL, "loss_heatmap: buffers must be 16-byte aligned");
  const long long n = (long long)BJ * hw;
  long long blocks = (n / 4 + kThreads * 4 - 1) / (kThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 1024) blocks = 1024;  // partial_ws must hold 1024 floats
  const float inv_bj = 1.f / (float)BJ;
  heatmap_loss_partial_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(pred, gt, n, mode, inv_bj,
                                                                                       grad_scale_dev, d_pred, partial_ws);
  count_launch();
  int rc = check_launch("heatmap_loss_partial_kernel");
  if (rc) return rc;
  heatmap_loss_final_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(partial_ws, (int)blocks, inv_bj, loss);
  count_launch();
  return check_launch("heatmap_loss_final_kernel");
}

extern "C" int hrnb_loss_pose2d(const float* pred, const float* gt, const float* vis, int32_t B, int32_t J, float* loss,
                                float* d_pred, void* stream) {
  if (!pred || !gt || !loss) return fail(HRNB_EINVAL, "loss_pose2d: null pointer");
  if (B <= 0 || J <= 0) return fail(HRNB_EINVAL, "loss_pose2d: bad shape");
  pose2d_loss_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(pred, gt, vis, B * J, J, loss, d_pred);
  count_launch();
  return check_launch("pose2d_loss_kernel");
}
