// Padded-flat geometry of a PF8 tensor + small bf16x8 helpers shared by the elementwise kernels.
#pragma once
#include "ptx.cuh"

namespace hrnb {

struct Geo {  // padded-flat geometry of one PF8 tensor
  int N, H, W, Hp, Wp;
  long long P;
};
__host__ __device__ inline Geo make_geo(int N, int H, int W) {
  Geo g;
  g.N = N; g.H = H; g.W = W; g.Hp = H + 1; g.Wp = W + 1;
  g.P = (long long)N * g.Hp * g.Wp;
  return g;
}


// position p -> (n, py, px); real pixels have py > 0 and px > 0
struct Pos {
  int n, py, px;
};
__device__ __forceinline__ Pos decode_pos(const Geo& g, long long p) {
  Pos r;
  if (g.P <= 0xffffffffLL) {   // uniform branch: 32-bit unsigned divisions (the 64-bit ones cost ~4x the instructions)
    const unsigned up = (unsigned)p;
    const unsigned rowi = up / (unsigned)g.Wp;
    r.px = (int)(up - rowi * (unsigned)g.Wp);
    const unsigned n = rowi / (unsigned)g.Hp;
    r.py = (int)(rowi - n * (unsigned)g.Hp);
    r.n = (int)n;
    return r;
  }
  r.px = (int)(p % g.Wp);
  const long long rowi = p / g.Wp;
  r.py = (int)(rowi % g.Hp);
  r.n = (int)(rowi / g.Hp);
  return r;
}

__device__ __forceinline__ void unpack8(const uint4 r, float (&a)[8]) {
  a[0] = bf16_lo(r.x); a[1] = bf16_hi(r.x); a[2] = bf16_lo(r.y); a[3] = bf16_hi(r.y);
  a[4] = bf16_lo(r.z); a[5] = bf16_hi(r.z); a[6] = bf16_lo(r.w); a[7] = bf16_hi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  uint4 o;
  o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]);
  o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
  return o;
}

// PyTorch upsample_bilinear2d source index rule (fp32 index math), shared by forward and backward
__device__ __forceinline__ void bil_index(int d, int in, int out, bool align, int& i0, int& i1, float& l1) {
  float src;
  if (align) {
    const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * (float)d;
  } else {
    const float sc = (float)in / (float)out;
    src = sc * ((float)d + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - (float)i0;
}

}  // namespace hrnb
