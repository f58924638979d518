// Register-tiled fp32 SIMT GEMM shared by the codec kernels (snac.cu, mimi.cu): fp32-exact building block for the
// vocoder stages whose results must stay within 1e-3 of the reference's fp32 convolutions.
#pragma once
#include "common.cuh"

namespace vb {

// ------------------------------------------------------------------------------------------
// register-tiled fp32 GEMM  C[M][N] = A[M][K] * B[K][N] with functor-defined B fetch and C store.  The N axis
// is the batch FOLDED with the position range: column g = b * nr + (n - n_lo), so the early decoder stages
// (a dozen positions per window) still fill whole tiles and every weight tile is read once per 64 columns of
// the whole batch rather than once per window.  Tile (16*TM) x 64, BK = 16, 256 threads, TM x 4 outputs per
// thread (TM = 8: rows ty*4.. and 64+ty*4..), next tile prefetched into registers while the current one is
// multiplied.  Each output is still one sequential FMA chain over k, so results do not depend on the tiling.
// ------------------------------------------------------------------------------------------
constexpr int GB_N = 64, GB_K = 16;

template <int TM, class BLoad, class CStore>
__device__ __forceinline__ void gemm_tile_f32(const float* __restrict__ A, int M, int K, int lda, int n_lo, int nr,
                                              int n_total, BLoad bload, CStore cstore) {
  constexpr int TILE_M = 16 * TM, AV = TM / 4;       // AV float4 of A per thread per k-step
  __shared__ float As[GB_K][TILE_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TILE_M, g0 = blockIdx.x * GB_N;
  // this thread's 4 columns (the same 4 for the B fetch and the C store)
  int cb[4], cn[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = g0 + tx * 4 + j;
    cb[j] = g < n_total ? g / nr : -1;
    cn[j] = n_lo + g - max(cb[j], 0) * nr;
  }
  const int am = tid >> 2, akq = (tid & 3) * 4, bk = tid >> 4;
  float4 pa[AV];
  float pb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < AV; ++h) {
      const int m = m0 + am + 64 * h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const float* src = A + static_cast<size_t>(m) * lda + k0 + akq;
        if (k0 + akq + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (k0 + akq + 0 < K) v.x = src[0];
          if (k0 + akq + 1 < K) v.y = src[1];
          if (k0 + akq + 2 < K) v.z = src[2];
        }
      }
      pa[h] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) pb[j] = (k0 + bk < K && cb[j] >= 0) ? bload(k0 + bk, cb[j], cn[j]) : 0.f;
  };
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
    for (int h = 0; h < AV; ++h) {
      As[akq + 0][am + 64 * h] = pa[h].x; As[akq + 1][am + 64 * h] = pa[h].y;
      As[akq + 2][am + 64 * h] = pa[h].z; As[akq + 3][am + 64 * h] = pa[h].w;
    }
    *reinterpret_cast<float4*>(&Bs[bk][tx * 4]) = make_float4(pb[0], pb[1], pb[2], pb[3]);
    __syncthreads();
    if (k0 + GB_K < K) fetch(k0 + GB_K);
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int h = 0; h < AV; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4 + 64 * h]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[h * 4 + i][j] += av[i] * bv[j];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int h = 0; h < AV; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + 64 * h + i;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m < M && cb[j] >= 0) cstore(m, cb[j], cn[j], acc[h * 4 + i][j]);
    }
}


}  // namespace vb
