// Loop glue of the hot path's callers: small HBM/latency-bound kernels that keep the data path on
// the device between the network and the host loop.
//
//   gen_heatmaps        ground-truth Gaussian heat maps from joint coordinates   lib/dataset/target_generators/target_generators.py:15-53
//   stem_im2col_u8      ToTensor + Normalize folded into the stem's im2col        lib/dataset/transforms/build.py:82-85
//   flip_merge          flip_back + SHIFT_HEATMAP + average of the flip test      lib/utils/transforms.py:16-30, lib/core/function.py:681-701
//   maxpool2_relu, gap_mlp   GlobalAveragePoolingHead (confidence head)           lib/models/pose_hrnet_volumetric.py:22-56
#include "ptx.cuh"
#include "common.h"
#include "geo.cuh"

namespace hrnb {

// ------------------------------------------------------------------------------------------------
// HeatmapGenerator.__call__: one block per (sample, joint) map, threads sweep the h*w pixels (coalesced fp32 stores).
// Semantics kept bit for bit: x = int(u), y = int(v) (truncation), joint skipped when not visible (pt[2] <= 0) or outside
// the map; the (6 sigma + 3)^2 patch [x - 3 sigma - 1, x + 3 sigma + 2) (numpy round = half to even) holds
// exp(-((px - x)^2 + (py - y)^2) / (2 sigma^2)) evaluated in float64 and stored as float32; everything else is zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gen_heatmaps_kernel(const float* __restrict__ joints, int stride, int BJ, int h, int w,
                                                          float sigma, float* __restrict__ out) {
  const int m = blockIdx.x;
  if (m >= BJ) return;
  const float* pt = joints + (long long)m * stride;
  const bool vis = stride < 3 || pt[2] > 0.f;
  const int x = (int)pt[0], y = (int)pt[1];          // int(): truncation toward zero, as Python's
  const bool on = vis && x >= 0 && y >= 0 && x < w && y < h;
  const double s3 = 3.0 * (double)sigma;
  const int ulx = (int)rint((double)x - s3 - 1.0), uly = (int)rint((double)y - s3 - 1.0);
  const int brx = (int)rint((double)x + s3 + 2.0), bry = (int)rint((double)y + s3 + 2.0);
  const double inv = 1.0 / (2.0 * (double)sigma * (double)sigma);
  float* o = out + (long long)m * h * w;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int py = i / w, px = i - py * w;
    float v = 0.f;
    if (on && px >= ulx && px < brx && py >= uly && py < bry) {
      // index into the reference's precomputed patch g: g[i] is centred at 3 sigma + 1 (== px - x for integer sigma)
      const double dx = (double)(px - ulx) - (s3 + 1.0), dy = (double)(py - uly) - (s3 + 1.0);
      v = (float)exp(-(dx * dx + dy * dy) * inv);
    }
    o[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// uint8 NHWC image -> PF8 im2col slab of the stem conv (see stem_im2col_kernel in elementwise.cu) with
// ToTensor (/255) and Normalize((v - mean[c]) / std[c]) applied on the fly; zero padding stays zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_im2col_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                            long long out_ps, Geo g, int inH, int inW, float m0, float m1,
                                                            float m2, float s0, float s1, float s2) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P) return;
  const Pos q = decode_pos(g, p);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  if (q.px > 0 && q.py > 0) {
    const float mean[3] = {m0, m1, m2}, istd[3] = {1.f / s0, 1.f / s1, 1.f / s2};
    const int iy0 = (q.py - 1) * 2 - 1, ix0 = (q.px - 1) * 2 - 1;
    const uint8_t* base = img + (long long)q.n * inH * inW * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = iy0 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ix0 + s;
        if (iy >= 0 && iy < inH && ix >= 0 && ix < inW) {
          const uint8_t* px3 = base + ((long long)iy * inW + ix) * 3;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
            v[ci * 9 + r * 3 + s] = ((float)px3[ci] / 255.f - mean[ci]) * istd[ci];   // ToTensor then Normalize (fp32, torchvision order)
        }
      }
    }
  }
#pragma unroll
  for (int pl = 0; pl < 4; ++pl) {
    uint4 o;
    o.x = pack_bf16x2(v[pl * 8 + 0], v[pl * 8 + 1]); o.y = pack_bf16x2(v[pl * 8 + 2], v[pl * 8 + 3]);
    o.z = pack_bf16x2(v[pl * 8 + 4], v[pl * 8 + 5]); o.w = pack_bf16x2(v[pl * 8 + 6], v[pl * 8 + 7]);
    *reinterpret_cast<uint4*>(out + ((long long)pl * out_ps + p) * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// flip test: out[b, j, y, x] = 0.5 * (hm[b, j, y, x] + flipped[b, perm[j], y, xs]) with xs = w-1-x, or w-x when the flipped
// map is shifted one pixel to the right (SHIFT_HEATMAP; column 0 of the shifted map keeps its unshifted value).
// perm = the joint permutation flip_back's pair swaps produce.  float4 per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flip_merge_kernel(const float* __restrict__ hm, const float* __restrict__ flipped,
                                                        const int* __restrict__ perm, int J, int h, int w, int shift,
                                                        long long total, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % w);
  const long long r = i / w;
  const int y = (int)(r % h);
  const long long bj = r / h;
  const int j = (int)(bj % J);
  const long long b = bj / J;
  int xs = w - 1 - x;
  if (shift && x > 0) xs = w - x;          // shifted[..., x] = flipped_back[..., x - 1] for x >= 1
  const float f = flipped[((b * J + perm[j]) * h + y) * w + xs];
  out[i] = hm != nullptr ? (hm[i] + f) * 0.5f : f;          // hm == NULL: flip_back (+ shift) alone
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2) + ReLU on PF8: thread = (output position, plane)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_relu_kernel(const __nv_bfloat16* __restrict__ src, long long src_ps, Geo sg,
                                                           __nv_bfloat16* __restrict__ dst, long long dst_ps, Geo dg) {
  const int plane = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= dg.P) return;
  const Pos q = decode_pos(dg, p);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (q.px > 0 && q.py > 0) {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = 0.f;                      // ReLU folded into the running maximum
    const long long base = ((long long)q.n * sg.Hp + (2 * (q.py - 1) + 1)) * sg.Wp + (2 * (q.px - 1) + 1);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float a[8];
        unpack8(ldg_nc_v4(src + ((long long)plane * src_ps + base + dy * sg.Wp + dx) * 8), a);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], a[i]);
      }
    o = pack8(m);
  }
  *reinterpret_cast<uint4*>(dst + ((long long)plane * dst_ps + p) * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// global average pool + Linear/ReLU + Linear/ReLU + Linear/Sigmoid, one block per sample (fp32; weights stream through L2)
// ------------------------------------------------------------------------------------------------
struct GapMlpK {
  const __nv_bfloat16* x;
  long long x_ps;
  Geo g;
  int C, H1, H2, NC;
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  float* out;
};
__global__ void __launch_bounds__(256) gap_mlp_kernel(const GapMlpK k) {
  __shared__ float a[1024], bvec[1024];
  const int n = blockIdx.x, t = threadIdx.x;
  const float inv = 1.f / (float)(k.g.H * k.g.W);
  for (int c = t; c < k.C; c += blockDim.x) {
    const __nv_bfloat16* plane = k.x + (long long)(c >> 3) * k.x_ps * 8 + (c & 7);
    float s = 0.f;
    for (int y = 1; y <= k.g.H; ++y) {
      const long long row = ((long long)n * k.g.Hp + y) * k.g.Wp;
      for (int x = 1; x <= k.g.W; ++x) s += __bfloat162float(plane[(row + x) * 8]);
    }
    a[c] = s * inv;
  }
  __syncthreads();
  for (int o = t; o < k.H1; o += blockDim.x) {
    float s = k.b1[o];
    const float* w = k.w1 + (long long)o * k.C;
    for (int i = 0; i < k.C; ++i) s = fmaf(w[i], a[i], s);
    bvec[o] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int o = t; o < k.H2; o += blockDim.x) {
    float s = k.b2[o];
    const float* w = k.w2 + (long long)o * k.H1;
    for (int i = 0; i < k.H1; ++i) s = fmaf(w[i], bvec[i], s);
    a[o] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int o = t; o < k.NC; o += blockDim.x) {
    float s = k.b3[o];
    const float* w = k.w3 + (long long)o * k.H2;
    for (int i = 0; i < k.H2; ++i) s = fmaf(w[i], a[i], s);
    k.out[(long long)n * k.NC + o] = 1.f / (1.f + __expf(-s));
  }
}

// out = a * x + b * y over n floats (float4 main loop): the weighted view fusion of Aggregation.fuse_with_weights
__global__ void __launch_bounds__(256) axpby_kernel(float a, const float* __restrict__ x, float b, const float* __restrict__ y,
                                                   float* __restrict__ out, long long n) {
  const long long nvec = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(x) + i), v = __ldg(reinterpret_cast<const float4*>(y) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z, a * u.w + b * v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) out[i] = a * x[i] + b * y[i];
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_axpby(float a, const float* x, float b, const float* y, float* out, int64_t n, void* stream) {
  if (!x || !y || !out || n < 0) return fail(HRNB_EINVAL, "axpby: bad params");
  if (n == 0) return HRNB_OK;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(HRNB_EINVAL, "axpby: buffers must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  axpby_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, x, b, y, out, (long long)n);
  count_launch();
  return check_launch("axpby_kernel");
}

extern "C" int hrnb_gen_heatmaps(const float* joints, int32_t joint_stride, int32_t BJ, int32_t h, int32_t w, float sigma,
                                 float* out, void* stream) {
  if (!joints || !out) return fail(HRNB_EINVAL, "gen_heatmaps: null pointer");
  if (BJ < 0 || h <= 0 || w <= 0 || sigma <= 0.f || (joint_stride != 2 && joint_stride != 3))
    return fail(HRNB_EINVAL, "gen_heatmaps: bad shape (joint rows are (u, v) or (u, v, visible))");
  if (BJ == 0) return HRNB_OK;
  gen_heatmaps_kernel<<<(unsigned)BJ, 256, 0, (cudaStream_t)stream>>>(joints, joint_stride, BJ, h, w, sigma, out);
  count_launch();
  return check_launch("gen_heatmaps_kernel");
}

extern "C" int hrnb_stem_im2col_u8(const uint8_t* img_nhwc, const float* mean3_host, const float* std3_host, void* out,
                                   int64_t out_ps, int32_t N, int32_t in_H, int32_t in_W, void* stream) {
  if (!img_nhwc || !mean3_host || !std3_host || !out) return fail(HRNB_EINVAL, "stem_im2col_u8: null pointer");
  if (N <= 0 || in_H <= 0 || in_W <= 0 || (in_H & 1) || (in_W & 1)) return fail(HRNB_EINVAL, "stem_im2col_u8: bad shape");
  const Geo g = make_geo(N, in_H / 2, in_W / 2);
  stem_im2col_u8_kernel<<<(unsigned)((g.P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      img_nhwc, (__nv_bfloat16*)out, out_ps, g, in_H, in_W, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0],
      std3_host[1], std3_host[2]);
  count_launch();
  return check_launch("stem_im2col_u8_kernel");
}

extern "C" int hrnb_flip_merge(const float* hm, const float* hm_flipped, const int32_t* perm_dev, int32_t B, int32_t J,
                               int32_t h, int32_t w, int32_t shift, float* out, void* stream) {
  if (!hm_flipped || !perm_dev || !out) return fail(HRNB_EINVAL, "flip_merge: null pointer");
  if (B < 0 || J <= 0 || h <= 0 || w <= 0) return fail(HRNB_EINVAL, "flip_merge: bad shape");
  const long long total = (long long)B * J * h * w;
  if (total == 0) return HRNB_OK;
  flip_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hm, hm_flipped, perm_dev, J, h, w, shift,
                                                                                       total, out);
  count_launch();
  return check_launch("flip_merge_kernel");
}

extern "C" int hrnb_maxpool2_relu(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                                  int64_t dst_ps, void* stream) {
  if (!src || !dst || C % 8 || N <= 0 || (H & 1) || (W & 1) || H <= 0 || W <= 0) return fail(HRNB_EINVAL, "maxpool2_relu: bad params");
  const Geo sg = make_geo(N, H, W), dg = make_geo(N, H / 2, W / 2);
  dim3 grid((unsigned)((dg.P + 255) / 256), C / 8);
  maxpool2_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, src_ps, sg, (__nv_bfloat16*)dst, dst_ps, dg);
  count_launch();
  return check_launch("maxpool2_relu_kernel");
}

extern "C" int hrnb_gap_mlp(const void* x, int64_t x_ps, int32_t N, int32_t C, int32_t H, int32_t W, const float* w1,
                            const float* b1, int32_t H1, const float* w2, const float* b2, int32_t H2, const float* w3,
                            const float* b3, int32_t NC, float* out, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !out) return fail(HRNB_EINVAL, "gap_mlp: null pointer");
  if (N <= 0 || C % 8 || C > 1024 || H1 > 1024 || H2 > 1024 || H1 <= 0 || H2 <= 0 || NC <= 0) return fail(HRNB_EINVAL, "gap_mlp: sizes up to 1024");
  GapMlpK k;
  k.x = (const __nv_bfloat16*)x; k.x_ps = x_ps; k.g = make_geo(N, H, W);
  k.C = C; k.H1 = H1; k.H2 = H2; k.NC = NC;
  k.w1 = w1; k.b1 = b1; k.w2 = w2; k.b2 = b2; k.w3 = w3; k.b3 = b3; k.out = out;
  gap_mlp_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(k);
  count_launch();
  return check_launch("gap_mlp_kernel");
}
