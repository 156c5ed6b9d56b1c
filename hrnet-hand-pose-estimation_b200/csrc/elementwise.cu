// HBM-bound companions of the tensor-core conv: weight packing, the 3-channel stem conv, the fuse-layer
// sum with nearest up-sampling, the head's bilinear up-sample/concat and layout conversion.
// All activations are PF8 (see include/hrnb.h); every thread moves one 16-byte position of one plane,
// so consecutive lanes touch consecutive 16-byte words (fully coalesced).
#include "ptx.cuh"
#include "common.h"
#include "geo.cuh"

namespace hrnb {

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [ntile][chunk][tap][KC][BN][8] bf16, scale folded
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                    const float* __restrict__ shift, int cout, int cin, int taps, int KC, int BN,
                                    int ntiles, __nv_bfloat16* __restrict__ out, float* __restrict__ bias_out) {
  const long long total = (long long)ntiles * BN * taps * cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)ntiles * BN) {
    bias_out[i] = (i < cout && shift) ? shift[i] : 0.f;
  }
  if (i >= total) return;
  // decode i = ((((nt*nch + c)*taps + t)*KC + j)*BN + n)*8 + e
  long long r = i;
  const int e = (int)(r % 8); r /= 8;
  const int n = (int)(r % BN); r /= BN;
  const int j = (int)(r % KC); r /= KC;
  const int t = (int)(r % taps); r /= taps;
  const int nch = (cin / 8) / KC;
  const int c = (int)(r % nch); r /= nch;
  const int nt = (int)r;
  const int co = nt * BN + n;
  const int ci = (c * KC + j) * 8 + e;
  float v = 0.f;
  if (co < cout) {
    v = w[((long long)co * cin + ci) * taps + t];
    if (scale) v *= scale[co];
  }
  out[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// stem conv1: [N,3,inH,inW] fp32 NCHW -> PF8 bf16 64 channels at (inH/2, inW/2); 3x3 s2 p1 + bias + ReLU
// thread = (position, plane of 8 channels)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                        long long out_ps, Geo g, int inH, int inW) {
  __shared__ __align__(16) float ws[27 * 64];  // [k = ci*9 + r*3 + s][co]
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int co = i % 64, kk = i / 64;
    ws[i] = w[co * 27 + kk];
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int plane = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P) return;
  const Pos q_ = decode_pos(g, p);      // 32-bit divisions whenever P fits (geo.cuh)
  const int px = q_.px, py = q_.py, n = q_.n;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (px > 0 && py > 0) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bs[plane * 8 + i];
    const int iy0 = (py - 1) * 2 - 1, ix0 = (px - 1) * 2 - 1;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* xc = x + ((long long)n * 3 + ci) * inH * inW;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = iy0 + r;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = ix0 + s;
          float v = 0.f;
          if (iy >= 0 && iy < inH && ix >= 0 && ix < inW) v = __ldg(xc + (long long)iy * inW + ix);
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[(ci * 9 + r * 3 + s) * 64 + plane * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[(ci * 9 + r * 3 + s) * 64 + plane * 8 + 4]);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  }
  *reinterpret_cast<uint4*>(out + ((long long)plane * out_ps + p) * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// stem im2col: [N,3,inH,inW] fp32 NCHW -> PF8 bf16, 32 channels (27 taps + 5 zeros) at (inH/2, inW/2)
// thread = output position; 27 scalar loads (L1 resident), four 16-byte stores
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         long long out_ps, Geo g, int inH, int inW) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P) return;
  const Pos q_ = decode_pos(g, p);      // 32-bit divisions whenever P fits (geo.cuh)
  const int px = q_.px, py = q_.py, n = q_.n;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  if (px > 0 && py > 0) {
    const int iy0 = (py - 1) * 2 - 1, ix0 = (px - 1) * 2 - 1;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* xc = x + ((long long)n * 3 + ci) * inH * inW;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = iy0 + r;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = ix0 + s;
          if (iy >= 0 && iy < inH && ix >= 0 && ix < inW) v[ci * 9 + r * 3 + s] = __ldg(xc + (long long)iy * inW + ix);
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
// phase split: PF8 [N,C,H,W] -> 4 x PF8 [N,C,H/2,W/2]; thread = (plane, source position)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) phase_split_kernel(const __nv_bfloat16* __restrict__ src, long long src_ps, Geo g,
                                                         __nv_bfloat16* __restrict__ dst, long long dst_ps,
                                                         long long phase_stride) {
  pdl_enter();
  const int plane = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P) return;
  const Pos q_ = decode_pos(g, p);      // 32-bit divisions whenever P fits (geo.cuh)
  const int px = q_.px, py = q_.py, n = q_.n;
  if (px == 0 || py == 0) return;
  const int y = py - 1, x = px - 1;
  const uint4 v = *reinterpret_cast<const uint4*>(src + ((long long)plane * src_ps + p) * 8);
  const int Hp2 = g.H / 2 + 1, Wp2 = g.W / 2 + 1;
  const long long q = ((long long)n * Hp2 + (y >> 1) + 1) * Wp2 + (x >> 1) + 1;
  *reinterpret_cast<uint4*>(dst + (long long)((y & 1) * 2 + (x & 1)) * phase_stride + ((long long)plane * dst_ps + q) * 8) = v;
}

// ------------------------------------------------------------------------------------------------
// fuse sum
// ------------------------------------------------------------------------------------------------
struct FuseK {
  const __nv_bfloat16* src[4];
  long long src_ps[4];
  int shift[4];
  int nsrc;
  __nv_bfloat16* out;
  long long out_ps;
  Geo g;
  int relu;
};

__device__ __forceinline__ void acc8(float (&a)[8], const uint4 r) {
  a[0] += bf16_lo(r.x); a[1] += bf16_hi(r.x); a[2] += bf16_lo(r.y); a[3] += bf16_hi(r.y);
  a[4] += bf16_lo(r.z); a[5] += bf16_hi(r.z); a[6] += bf16_lo(r.w); a[7] += bf16_hi(r.w);
}

__global__ void __launch_bounds__(256) fuse_sum_kernel(const FuseK k) {
  pdl_enter();
  const int plane = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= k.g.P) return;
  const Pos q_ = decode_pos(k.g, p);      // 32-bit divisions whenever P fits (geo.cuh)
  const int px = q_.px, py = q_.py, n = q_.n;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (px > 0 && py > 0) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // summation order j = 0..nsrc-1 as in the reference loop (pose_hrnet.py:260-265)
    for (int s = 0; s < k.nsrc; ++s) {
      const int sh = k.shift[s];
      const int sHp = (k.g.H >> sh) + 1, sWp = (k.g.W >> sh) + 1;
      const long long sp = ((long long)n * sHp + (((py - 1) >> sh) + 1)) * sWp + (((px - 1) >> sh) + 1);
      const uint4 r = *reinterpret_cast<const uint4*>(k.src[s] + ((long long)plane * k.src_ps[s] + sp) * 8);
      acc8(a, r);
    }
    if (k.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], 0.f);
    }
    o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]);
    o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
  }
  *reinterpret_cast<uint4*>(k.out + ((long long)plane * k.out_ps + p) * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// bilinear up-sample (PyTorch upsample_bilinear2d index rules, fp32 index math) PF8 -> PF8
// ------------------------------------------------------------------------------------------------
constexpr int kBilPlanes = 4;   // channel planes per thread: the index / weight math of a position is shared by all of them
__global__ void __launch_bounds__(256) bilinear_kernel(const __nv_bfloat16* __restrict__ src, long long src_ps, Geo sg,
                                                      __nv_bfloat16* __restrict__ dst, long long dst_ps, Geo dg,
                                                      int align, int planes) {
  const int plane0 = blockIdx.y * kBilPlanes;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= dg.P) return;
  const Pos q_ = decode_pos(dg, p);
  const int px = q_.px, py = q_.py, n = q_.n;
  const bool real = px > 0 && py > 0;
  int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
  float ly = 0.f, lx = 0.f;
  if (real) {
    bil_index(py - 1, sg.H, dg.H, align != 0, y0, y1, ly);
    bil_index(px - 1, sg.W, dg.W, align != 0, x0, x1, lx);
  }
  const float hy = 1.f - ly, hx = 1.f - lx;
  const long long r0 = ((long long)n * sg.Hp + y0 + 1) * sg.Wp, r1 = ((long long)n * sg.Hp + y1 + 1) * sg.Wp;
  const long long o00 = (r0 + x0 + 1) * 8, o01 = (r0 + x1 + 1) * 8, o10 = (r1 + x0 + 1) * 8, o11 = (r1 + x1 + 1) * 8;
#pragma unroll
  for (int j = 0; j < kBilPlanes; ++j) {
    const int plane = plane0 + j;
    if (plane >= planes) break;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (real) {
      const __nv_bfloat16* base = src + (long long)plane * src_ps * 8;
      const uint4 v00 = *reinterpret_cast<const uint4*>(base + o00);
      const uint4 v01 = *reinterpret_cast<const uint4*>(base + o01);
      const uint4 v10 = *reinterpret_cast<const uint4*>(base + o10);
      const uint4 v11 = *reinterpret_cast<const uint4*>(base + o11);
      const uint32_t* a = &v00.x; const uint32_t* b = &v01.x; const uint32_t* c = &v10.x; const uint32_t* d = &v11.x;
      uint32_t ow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // same association as ATen: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11)
        const float lo = hy * (hx * bf16_lo(a[i]) + lx * bf16_lo(b[i])) + ly * (hx * bf16_lo(c[i]) + lx * bf16_lo(d[i]));
        const float hi = hy * (hx * bf16_hi(a[i]) + lx * bf16_hi(b[i])) + ly * (hx * bf16_hi(c[i]) + lx * bf16_hi(d[i]));
        ow[i] = pack_bf16x2(lo, hi);
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    *reinterpret_cast<uint4*>(dst + ((long long)plane * dst_ps + p) * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------
// PF8 -> NCHW fp32. thread = (n, c, y, x) with x fastest: 4-byte coalesced writes, 2-byte strided reads
// served from L1/L2 (the 8 channel-threads of a plane hit the same 16-byte word).
__global__ void __launch_bounds__(256) pf8_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long long src_ps, Geo g,
                                                         int C, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * C * g.H * g.W;
  if (i >= total) return;
  const int x = (int)(i % g.W);
  long long r = i / g.W;
  const int y = (int)(r % g.H); r /= g.H;
  const int c = (int)(r % C);
  const int n = (int)(r / C);
  const long long p = ((long long)n * g.Hp + y + 1) * g.Wp + x + 1;
  dst[i] = __bfloat162float(src[((long long)(c >> 3) * src_ps + p) * 8 + (c & 7)]);
}

// NCHW fp32 -> PF8 (C padded up to a multiple of 8 with zeros); thread = (plane, position)
__global__ void __launch_bounds__(256) nchw_to_pf8_kernel(const float* __restrict__ src, Geo g, int C,
                                                         __nv_bfloat16* __restrict__ dst, long long dst_ps) {
  const int plane = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P) return;
  const Pos q_ = decode_pos(g, p);      // 32-bit divisions whenever P fits (geo.cuh)
  const int px = q_.px, py = q_.py, n = q_.n;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (px > 0 && py > 0) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = plane * 8 + e;
      v[e] = c < C ? src[(((long long)n * C + c) * g.H + (py - 1)) * g.W + (px - 1)] : 0.f;
    }
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  }
  *reinterpret_cast<uint4*>(dst + ((long long)plane * dst_ps + p) * 8) = o;
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_pack_conv_weights(const float* w, const float* scale, const float* shift, int32_t cout, int32_t cin,
                                      int32_t taps, int32_t KC, int32_t BN, void* wpk_out, float* bias_out,
                                      void* stream) {
  if (!w || !wpk_out || !bias_out) return fail(HRNB_EINVAL, "pack: null pointer");
  if (cin % 16 || KC <= 0 || (cin / 8) % KC || BN % 16 || BN <= 0 || (taps != 1 && taps != 9))
    return fail(HRNB_EINVAL, "pack: bad geometry");
  const int ntiles = (cout + BN - 1) / BN;
  const long long total = (long long)ntiles * BN * taps * cin;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  pack_weights_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(w, scale, shift, cout, cin, taps, KC, BN,
                                                                             ntiles, (__nv_bfloat16*)wpk_out, bias_out);
  count_launch();
  return check_launch("pack_weights_kernel");
}

extern "C" int hrnb_stem_conv1(const float* x, const float* w, const float* bias, void* out, int64_t out_ps, int32_t N,
                               int32_t in_H, int32_t in_W, void* stream) {
  if (!x || !w || !bias || !out) return fail(HRNB_EINVAL, "stem: null pointer");
  if (N <= 0 || in_H <= 0 || in_W <= 0 || (in_H & 1) || (in_W & 1)) return fail(HRNB_EINVAL, "stem: H, W must be even");
  const Geo g = make_geo(N, in_H / 2, in_W / 2);
  dim3 grid((unsigned)((g.P + 255) / 256), 8);
  stem_conv1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, (__nv_bfloat16*)out, out_ps, g, in_H, in_W);
  count_launch();
  return check_launch("stem_conv1_kernel");
}

extern "C" int hrnb_stem_im2col(const float* x, void* out, int64_t out_ps, int32_t N, int32_t in_H, int32_t in_W,
                                void* stream) {
  if (!x || !out) return fail(HRNB_EINVAL, "stem_im2col: null pointer");
  if (N <= 0 || in_H <= 0 || in_W <= 0 || (in_H & 1) || (in_W & 1)) return fail(HRNB_EINVAL, "stem_im2col: H, W must be even");
  const Geo g = make_geo(N, in_H / 2, in_W / 2);
  stem_im2col_kernel<<<(unsigned)((g.P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, out_ps, g,
                                                                                      in_H, in_W);
  count_launch();
  return check_launch("stem_im2col_kernel");
}

extern "C" int hrnb_phase_split(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                                int64_t dst_ps, int64_t phase_stride, void* stream) {
  if (!src || !dst || C % 8 || (H & 1) || (W & 1) || phase_stride <= 0) return fail(HRNB_EINVAL, "phase_split: bad params");
  const Geo g = make_geo(N, H, W);
  dim3 grid((unsigned)((g.P + 255) / 256), C / 8);
  launch_pdl(phase_split_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, (long long)src_ps, g,
             (__nv_bfloat16*)dst, (long long)dst_ps, (long long)phase_stride);
  count_launch();
  return check_launch("phase_split_kernel");
}

extern "C" int hrnb_fuse_sum(const hrnb_fuse_params* p, void* stream) {
  if (!p || !p->out || p->nsrc < 1 || p->nsrc > 4) return fail(HRNB_EINVAL, "fuse: bad params");
  if (p->C % 8) return fail(HRNB_EINVAL, "fuse: C must be a multiple of 8");
  FuseK k;
  for (int i = 0; i < 4; ++i) {
    k.src[i] = i < p->nsrc ? (const __nv_bfloat16*)p->src[i] : nullptr;
    k.src_ps[i] = p->src_ps[i];
    k.shift[i] = p->shift[i];
    if (i < p->nsrc) {
      if (!p->src[i]) return fail(HRNB_EINVAL, "fuse: null source");
      if (p->shift[i] < 0 || p->shift[i] > 3 || (p->H % (1 << p->shift[i])) || (p->W % (1 << p->shift[i])))
        return fail(HRNB_EINVAL, "fuse: bad shift");
    }
  }
  k.nsrc = p->nsrc;
  k.out = (__nv_bfloat16*)p->out;
  k.out_ps = p->out_ps;
  k.g = make_geo(p->N, p->H, p->W);
  k.relu = p->relu;
  dim3 grid((unsigned)((k.g.P + 255) / 256), p->C / 8);
  launch_pdl(fuse_sum_kernel, grid, dim3(256), 0, (cudaStream_t)stream, k);
  count_launch();
  return check_launch("fuse_sum_kernel");
}

extern "C" int hrnb_bilinear_up(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t sH, int32_t sW, void* dst,
                                int64_t dst_ps, int32_t dH, int32_t dW, int32_t align_corners, void* stream) {
  if (!src || !dst || C % 8) return fail(HRNB_EINVAL, "bilinear: bad params");
  const Geo sg = make_geo(N, sH, sW), dg = make_geo(N, dH, dW);
  dim3 grid((unsigned)((dg.P + 255) / 256), (C / 8 + kBilPlanes - 1) / kBilPlanes);
  bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, src_ps, sg, (__nv_bfloat16*)dst,
                                                           dst_ps, dg, align_corners, C / 8);
  count_launch();
  return check_launch("bilinear_kernel");
}

extern "C" int hrnb_pf8_to_nchw_f32(const void* src, int64_t src_ps, int32_t N, int32_t C, int32_t H, int32_t W,
                                    float* dst, void* stream) {
  if (!src || !dst) return fail(HRNB_EINVAL, "pf8_to_nchw: null pointer");
  const Geo g = make_geo(N, H, W);
  const long long total = (long long)N * C * H * W;
  pf8_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src,
                                                                                        src_ps, g, C, dst);
  count_launch();
  return check_launch("pf8_to_nchw_kernel");
}

extern "C" int hrnb_nchw_f32_to_pf8(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                                    int64_t dst_ps, void* stream) {
  if (!src || !dst) return fail(HRNB_EINVAL, "nchw_to_pf8: null pointer");
  const Geo g = make_geo(N, H, W);
  dim3 grid((unsigned)((g.P + 255) / 256), (C + 7) / 8);
  nchw_to_pf8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, g, C, (__nv_bfloat16*)dst, dst_ps);
  count_launch();
  return check_launch("nchw_to_pf8_kernel");
}
