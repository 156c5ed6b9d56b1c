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
// STAGED at the end of round 1 (written without GPU access): parity against the oracle / golden fixture is tested in
// tests/test_zz_staged_gpu.py; the oracle itself is pinned to the unmodified reference (tests/golden/triangulation.npz).
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

}  // namespace hrnb

using namespace hrnb;

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
