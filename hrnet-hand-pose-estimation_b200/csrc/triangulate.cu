// Algebraic (DLT) triangulation of J joints from V views - SURVEY §8 row (f), BASELINE configs[3].
//
// Replaces the per-joint Python loop of AlgebraicTriangulationNet.forward (lib/models/triangulation.py:258-261) over
// DLT_sii_pytorch (lib/utils/misc.py:64-97): per (sample, joint) build the 2V x 4 system
//     rows  u * P[v][2] - P[v][0],   v * P[v][2] - P[v][1]
// form A^T A + 1e-3 I (fp32, as the reference's `.float()`), run `iterations` steps of shifted inverse iteration from a
// given unit start vector (the reference draws it with torch.rand on the host; the Python mirror does the same and passes
// it in), and return -b_k de-homogenised (lib/utils/misc.py:28-35).  The reference spends 21 x ~12 tiny torch launches per
// batch on this; here it is ONE launch, one thread per (sample, joint): 4x4 LU with partial pivoting in registers (the
// factorisation torch.linalg.solve / LAPACK gesv performs).  Latency-, not bandwidth-bound: B*J*(8V + 28) bytes.
//
// Forward verified on B200 against the reference golden (tests/test_gpu_head_backward_dlt.py; tests/golden/triangulation.npz).
// The reference function is differentiable (train3D back-propagates the 3-D loss through it into the backbone):
// triangulate_dlt_bwd_kernel is the hand-written adjoint w.r.t. the 2-D points - shifted inverse iteration re-run in
// registers, then the chain  X = b[:3]/b[3]  <-  b = y/|y|  <-  y = B^-1 b_prev  (B symmetric: the adjoint solve uses the
// same matrix)  <-  B = A^T A + eps I  <-  rows of A linear in (u, v).
#include "common.h"

namespace hrnb {

// solve M x = b for a 4x4 system, LU with partial pivoting (M, b are overwritten); all loops unroll into registers
__device__ __forceinline__ void solve4(float (&M)[4][4], float (&b)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    float best = fabsf(M[c][c]);
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const float a = fabsf(M[r][c]);
      if (a > best) { best = a; piv = r; }
    }
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {          // swap rows c <-> piv without dynamic register indexing
      if (r == piv) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float t = M[c][k]; M[c][k] = M[r][k]; M[r][k] = t; }
        const float t = b[c]; b[c] = b[r]; b[r] = t;
      }
    }
    const float inv = 1.f / M[c][c];
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      const float f = M[r][c] * inv;
#pragma unroll
      for (int k = c + 1; k < 4; ++k) M[r][k] = fmaf(-f, M[c][k], M[r][k]);
      b[r] = fmaf(-f, b[c], b[r]);
    }
  }
#pragma unroll
  for (int c = 3; c >= 0; --c) {
    float s = b[c];
#pragma unroll
    for (int k = c + 1; k < 4; ++k) s = fmaf(-M[c][k], b[k], s);
    b[c] = s / M[c][c];
  }
}

// points [B][V][J][2], proj [B][V][3][4], bk0 [J][B][4] (unit vectors), out [B][J][3]
__global__ void __launch_bounds__(128) triangulate_dlt_kernel(const float* __restrict__ points, const float* __restrict__ proj,
                                                             const float* __restrict__ bk0, int B, int V, int J, int iterations,
                                                             float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * J) return;
  const int b = idx / J, j = idx - b * J;
  float AtA[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) AtA[r][c] = 0.f;
  for (int v = 0; v < V; ++v) {
    const float* P = proj + ((long long)b * V + v) * 12;
    const float* uv = points + (((long long)b * V + v) * J + j) * 2;
    const float u = uv[0], w = uv[1];
    float r0[4], r1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float p2 = __ldg(P + 8 + k);
      r0[k] = u * p2 - __ldg(P + k);
      r1[k] = w * p2 - __ldg(P + 4 + k);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) AtA[r][c] = fmaf(r1[r], r1[c], fmaf(r0[r], r0[c], AtA[r][c]));
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) AtA[d][d] += 0.001f;
  float bk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) bk[k] = bk0[((long long)j * B + b) * 4 + k];
  for (int it = 0; it < iterations; ++it) {
    float M[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) M[r][c] = AtA[r][c];
    solve4(M, bk);
    const float inv = rsqrtf(bk[0] * bk[0] + bk[1] * bk[1] + bk[2] * bk[2] + bk[3] * bk[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) bk[k] *= inv;
  }
  const float inv_w = 1.f / bk[3];       // (-b) / (-b_w): the sign cancels in the de-homogenisation
  float* o = out + (long long)idx * 3;
  o[0] = bk[0] * inv_w;
  o[1] = bk[1] * inv_w;
  o[2] = bk[2] * inv_w;
}

constexpr int kMaxDltIters = 4;

// d_out [B][J][3] -> d_points [B][V][J][2]  (projection matrices are data: no gradient, as in the reference's use)
__global__ void __launch_bounds__(128) triangulate_dlt_bwd_kernel(const float* __restrict__ points, const float* __restrict__ proj,
                                                                 const float* __restrict__ bk0, const float* __restrict__ d_out,
                                                                 int B, int V, int J, int iterations, float* __restrict__ d_points) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * J) return;
  const int b = idx / J, j = idx - b * J;
  float AtA[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) AtA[r][c] = 0.f;
  for (int v = 0; v < V; ++v) {
    const float* P = proj + ((long long)b * V + v) * 12;
    const float* uv = points + (((long long)b * V + v) * J + j) * 2;
    const float u = uv[0], w = uv[1];
    float r0[4], r1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float p2 = __ldg(P + 8 + k);
      r0[k] = u * p2 - __ldg(P + k);
      r1[k] = w * p2 - __ldg(P + 4 + k);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) AtA[r][c] = fmaf(r1[r], r1[c], fmaf(r0[r], r0[c], AtA[r][c]));
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) AtA[d][d] += 0.001f;
  // forward pass, keeping every iterate: y_k = B^-1 b_{k-1}, b_k = y_k / |y_k|
  float ys[kMaxDltIters][4], bs[kMaxDltIters][4], invn[kMaxDltIters];
  float bk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) bk[k] = bk0[((long long)j * B + b) * 4 + k];
#pragma unroll
  for (int it = 0; it < kMaxDltIters; ++it) {
    if (it < iterations) {
      float M[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[r][c] = AtA[r][c];
      solve4(M, bk);
      const float inv = rsqrtf(bk[0] * bk[0] + bk[1] * bk[1] + bk[2] * bk[2] + bk[3] * bk[3]);
      invn[it] = inv;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ys[it][k] = bk[k]; bk[k] *= inv; bs[it][k] = bk[k]; }
    }
  }
  // X_i = b_i / b_3
  const float* g = d_out + (long long)idx * 3;
  const float inv_w = 1.f / bk[3];
  float db[4];
  db[0] = g[0] * inv_w; db[1] = g[1] * inv_w; db[2] = g[2] * inv_w;
  db[3] = -(g[0] * bk[0] + g[1] * bk[1] + g[2] * bk[2]) * inv_w * inv_w;
  float dB[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) dB[r][c] = 0.f;
#pragma unroll
  for (int it = kMaxDltIters - 1; it >= 0; --it) {
    if (it < iterations) {
      const float dot = bs[it][0] * db[0] + bs[it][1] * db[1] + bs[it][2] * db[2] + bs[it][3] * db[3];
      float lam[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lam[k] = (db[k] - bs[it][k] * dot) * invn[it];      // d y_k
      float M[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[r][c] = AtA[r][c];
      solve4(M, lam);                                                                  // lam = B^-T d y_k  (B symmetric)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dB[r][c] = fmaf(-lam[r], ys[it][c], dB[r][c]);
#pragma unroll
      for (int k = 0; k < 4; ++k) db[k] = lam[k];                                      // d b_{k-1}
    }
  }
  // B = sum_rows a a^T + eps I  ->  d a = (dB + dB^T) a ;  a = (u or v) * P[2] - P[0 or 1]
  for (int v = 0; v < V; ++v) {
    const float* P = proj + ((long long)b * V + v) * 12;
    const float* uv = points + (((long long)b * V + v) * J + j) * 2;
    const float u = uv[0], w = uv[1];
    float r0[4], r1[4], p2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      p2[k] = __ldg(P + 8 + k);
      r0[k] = u * p2[k] - __ldg(P + k);
      r1[k] = w * p2[k] - __ldg(P + 4 + k);
    }
    float du = 0.f, dw = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float s_ = dB[r][c] + dB[c][r];
        a0 = fmaf(s_, r0[c], a0);
        a1 = fmaf(s_, r1[c], a1);
      }
      du = fmaf(a0, p2[r], du);
      dw = fmaf(a1, p2[r], dw);
    }
    float* o = d_points + (((long long)b * V + v) * J + j) * 2;
    o[0] = du;
    o[1] = dw;
  }
}

}  // namespace hrnb

using namespace hrnb;

extern "C" int hrnb_triangulate_dlt_bwd(const float* points, const float* proj, const float* bk0, const float* d_out, int32_t B,
                                        int32_t V, int32_t J, int32_t iterations, float* d_points, void* stream) {
  if (!points || !proj || !bk0 || !d_out || !d_points) return fail(HRNB_EINVAL, "triangulate_dlt_bwd: null pointer");
  if (B <= 0 || V < 2 || J <= 0 || iterations < 1 || iterations > kMaxDltIters || (long long)B * J > 0x7fffffffLL)
    return fail(HRNB_EINVAL, "triangulate_dlt_bwd: need B, J >= 1, V >= 2 views, 1 <= iterations <= 4");
  const int n = B * J;
  triangulate_dlt_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(points, proj, bk0, d_out, B, V, J,
                                                                                            iterations, d_points);
  count_launch();
  return check_launch("triangulate_dlt_bwd_kernel");
}

extern "C" int hrnb_triangulate_dlt(const float* points, const float* proj, const float* bk0, int32_t B, int32_t V, int32_t J,
                                    int32_t iterations, float* out, void* stream) {
  if (!points || !proj || !bk0 || !out) return fail(HRNB_EINVAL, "triangulate_dlt: null pointer");
  if (B <= 0 || V < 2 || J <= 0 || iterations < 1 || (long long)B * J > 0x7fffffffLL)
    return fail(HRNB_EINVAL, "triangulate_dlt: need B, J >= 1, V >= 2 views, iterations >= 1");
  const int n = B * J;
  triangulate_dlt_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(points, proj, bk0, B, V, J, iterations, out);
  count_launch();
  return check_launch("triangulate_dlt_kernel");
}
