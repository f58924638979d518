// SNAC decoder stages (vox_serve/tokenizer/snac.py:119-267, 297-357), fp32, activations [B][C][T] with T
// contiguous.  Weight-norm is folded on the host at load time (snac.py:244-249 re-derives it every
// forward).  Every Snake activation is fused into the epilogue of the kernel that produces its input
// (or, for the residual stream, into the depthwise conv's tile load), so the decoder is
//   from_codes -> dwconv7 -> pwconv(+snake) -> 4 x [ convtr -> pwconv(noise) -> 3 x (dwconv7(snake in,
//   snake out) -> pwconv(+residual[, +snake])) ] -> final conv7 + tanh.
// The two GEMM-shaped stages (1x1 convs, transposed convs) share one register-tiled fp32 kernel.
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

__device__ __forceinline__ float snake_f(float x, float alpha) {
  const float s = sinf(alpha * x);
  return x + (1.0f / (alpha + 1e-9f)) * s * s;
}

// ------------------------------------------------------------------------------------------
// RVQ from_codes (snac.py:297-301, 350-357)
// ------------------------------------------------------------------------------------------
__global__ void from_codes_kernel(float* __restrict__ z, const int32_t* __restrict__ c0,
                                  const int32_t* __restrict__ c1, const int32_t* __restrict__ c2,
                                  const float* __restrict__ cb, const float* __restrict__ pw,
                                  const float* __restrict__ pb, int C, int T, int cb_size, int cb_dim, int s0,
                                  int s1, int s2) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int32_t* codes[3] = {c0, c1, c2};
  const int strides[3] = {s0, s1, s2};
  __shared__ float e[3][16];
  if (threadIdx.x < 3 * cb_dim) {
    const int i = threadIdx.x / cb_dim, d = threadIdx.x % cb_dim;
    const int tl = T / strides[i];
    int code = codes[i][b * tl + t / strides[i]];
    code = min(max(code, 0), cb_size - 1);
    e[i][d] = cb[(static_cast<size_t>(i) * cb_size + code) * cb_dim + d];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float a = pb[i * C + c];
      for (int d = 0; d < cb_dim; ++d) a += pw[(static_cast<size_t>(i) * C + c) * cb_dim + d] * e[i][d];
      acc += a;
    }
    z[(static_cast<size_t>(b) * C + c) * T + t] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// depthwise conv k=7, dilation d, "same" padding; optional Snake on the input tile and on the output
// ------------------------------------------------------------------------------------------
constexpr int DW_TILE = 256;
__global__ void __launch_bounds__(DW_TILE) dwconv7_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ alpha_in,
                                                          const float* __restrict__ alpha_out, int C, int T, int dil,
                                                          int t_lo, int t_hi) {
  extern __shared__ float tile[];  // DW_TILE + 6*dil
  const int c = blockIdx.y, b = blockIdx.z, t0 = t_lo + blockIdx.x * DW_TILE;
  const float* xr = x + (static_cast<size_t>(b) * C + c) * T;
  const int halo = 3 * dil, n = DW_TILE + 2 * halo;
  const float ain = alpha_in ? alpha_in[c] : 0.f;
  for (int i = threadIdx.x; i < n; i += DW_TILE) {
    const int t = t0 - halo + i;
    float v = 0.f;
    if (t >= 0 && t < T) {
      v = xr[t];
      if (alpha_in) v = snake_f(v, ain);
    }
    tile[i] = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t < t_hi) {
    float acc = bias ? bias[c] : 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc += w[c * 7 + k] * tile[threadIdx.x + k * dil];
    if (alpha_out) acc = snake_f(acc, alpha_out[c]);
    y[(static_cast<size_t>(b) * C + c) * T + t] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// register-tiled fp32 GEMM  C[M][N] = A[M][K] * B[K][N], per batch item, with functor-defined B fetch
// and C store.  64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------
constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

template <class BLoad, class CStore>
__device__ __forceinline__ void gemm_tile_f32(const float* __restrict__ A, int M, int K, int lda, int n_lo, int n_hi,
                                              BLoad bload, CStore cstore) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GB_M, n0 = n_lo + blockIdx.x * GB_N;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GB_K) {
    // A tile: 64 x 16, K contiguous in memory -> 4 floats per thread
    {
      const int m = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + m < M) {
        const float* src = A + static_cast<size_t>(m0 + m) * lda + k0 + kq;
        if (k0 + kq + 3 < K) v = *reinterpret_cast<const float4*>(src);
        else {
          if (k0 + kq + 0 < K) v.x = src[0];
          if (k0 + kq + 1 < K) v.y = src[1];
          if (k0 + kq + 2 < K) v.z = src[2];
        }
      }
      As[kq + 0][m] = v.x; As[kq + 1][m] = v.y; As[kq + 2][m] = v.z; As[kq + 3][m] = v.w;
    }
    // B tile: 16 x 64, N contiguous
    {
      const int k = tid >> 4, nq = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + nq + j;
        Bs[k][nq + j] = (k0 + k < K && n < n_hi) ? bload(k0 + k, n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < n_hi) cstore(m, n, acc[i][j]);
    }
  }
}

// pointwise conv; epilogue 0 plain, 1 + resid, 2 noise block (x + noise * Wx); optional Snake after
__global__ void __launch_bounds__(256) pwconv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     const float* __restrict__ resid, const float* __restrict__ noise,
                                                     const float* __restrict__ alpha_out, int epi, int Cin, int Cout,
                                                     int T, int t_lo, int t_hi) {
  const int b = blockIdx.z;
  const float* xb = x + static_cast<size_t>(b) * Cin * T;
  float* yb = y + static_cast<size_t>(b) * Cout * T;
  const float* rb = resid ? resid + static_cast<size_t>(b) * Cout * T : nullptr;
  const float* nb = noise ? noise + static_cast<size_t>(b) * T : nullptr;
  gemm_tile_f32(
      w, Cout, Cin, Cin, t_lo, t_hi, [&](int k, int n) { return xb[static_cast<size_t>(k) * T + n]; },
      [&](int m, int n, float v) {
        if (bias) v += bias[m];
        if (epi == 1) v += rb[static_cast<size_t>(m) * T + n];
        else if (epi == 2) v = xb[static_cast<size_t>(m) * T + n] + nb[n] * v;
        if (alpha_out) v = snake_f(v, alpha_out[m]);
        yb[static_cast<size_t>(m) * T + n] = v;
      });
}

// transposed conv, kernel 2s, stride s, padding ceil(s/2), output_padding s%2 (snac.py:222-231).
// wp: host-repacked [s][Cout][2*Cin]: wp[r][co][tap*Cin + ci] = W[ci][co][r + tap*s].
// For phase r = blockIdx.z % s the GEMM N index is q = ti (input position); output to = ti*s + r - pad.
__global__ void __launch_bounds__(256) convtr_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ wp, const float* __restrict__ bias,
                                                     const float* __restrict__ alpha_out, int Cin, int Cout, int T,
                                                     int s, int pad, int n_lo, int n_hi) {
  const int r = blockIdx.z % s, b = blockIdx.z / s;
  const int Tout = T * s;
  const float* xb = x + static_cast<size_t>(b) * Cin * T;
  float* yb = y + static_cast<size_t>(b) * Cout * Tout;
  const float* wr = wp + static_cast<size_t>(r) * Cout * 2 * Cin;
  // ti ranges over [0, T]: ti = T only receives the tap-1 term (x[T-1])
  gemm_tile_f32(
      wr, Cout, 2 * Cin, 2 * Cin, n_lo, n_hi,
      [&](int k, int n) {
        const int tap = k >= Cin, ci = k - tap * Cin, ti = n - tap;
        return (ti >= 0 && ti < T) ? xb[static_cast<size_t>(ci) * T + ti] : 0.f;
      },
      [&](int m, int n, float v) {
        const int to = n * s + r - pad;
        if (to < 0 || to >= Tout) return;
        if (bias) v += bias[m];
        if (alpha_out) v = snake_f(v, alpha_out[m]);
        yb[static_cast<size_t>(m) * Tout + to] = v;
      });
}

// final: tanh(conv_k7(x) + bias), Cout = 1; x already carries the last Snake
__global__ void __launch_bounds__(256) final_conv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ alpha_in, int C, int T, int t0,
                                                         int t1) {
  extern __shared__ float ws[];  // C*7 weights (+ C alphas)
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) ws[i] = w[i];
  if (alpha_in)
    for (int i = threadIdx.x; i < C; i += blockDim.x) ws[C * 7 + i] = alpha_in[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t1) return;
  const float* xb = x + static_cast<size_t>(b) * C * T;
  float acc = bias[0];
  for (int c = 0; c < C; ++c) {
    const float* xr = xb + static_cast<size_t>(c) * T;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int tt = t + k - 3;
      if (tt >= 0 && tt < T) {
        float v = xr[tt];
        if (alpha_in) v = snake_f(v, ws[C * 7 + c]);
        acc += ws[c * 7 + k] * v;
      }
    }
  }
  y[static_cast<size_t>(b) * (t1 - t0) + (t - t0)] = tanhf(acc);
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_snac_from_codes(float* d_z, const int32_t* d_codes0, const int32_t* d_codes1, const int32_t* d_codes2,
                       const float* d_codebooks, const float* d_proj_w, const float* d_proj_b, int B, int C, int T,
                       int cb_size, int cb_dim, int stride0, int stride1, int stride2, void* stream) {
  VB_CHECK_ARG(d_z && d_codes0 && d_codes1 && d_codes2 && d_codebooks && d_proj_w && d_proj_b,
               "vb_snac_from_codes: null pointer");
  VB_CHECK_ARG(cb_dim <= 16 && T % stride0 == 0 && T % stride1 == 0 && T % stride2 == 0,
               "vb_snac_from_codes: bad dims");
  if (B <= 0) return 0;
  from_codes_kernel<<<dim3(T, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_z, d_codes0, d_codes1, d_codes2, d_codebooks, d_proj_w, d_proj_b, C, T, cb_size, cb_dim, stride0, stride1,
      stride2);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_dwconv7(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_alpha_in,
                    const float* d_alpha_out, int B, int C, int T, int dilation, int t_lo, int t_hi, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w, "vb_snac_dwconv7: null pointer");
  VB_CHECK_ARG(dilation >= 1 && dilation <= 64, "vb_snac_dwconv7: dilation %d", dilation);
  VB_CHECK_ARG(0 <= t_lo && t_lo <= t_hi && t_hi <= T, "vb_snac_dwconv7: range [%d, %d) outside [0, %d)", t_lo, t_hi, T);
  if (B <= 0 || t_lo == t_hi) return 0;
  const size_t smem = (DW_TILE + 6 * dilation) * sizeof(float);
  dwconv7_kernel<<<dim3((t_hi - t_lo + DW_TILE - 1) / DW_TILE, C, B), DW_TILE, smem, static_cast<cudaStream_t>(stream)>>>(
      d_y, d_x, d_w, d_bias, d_alpha_in, d_alpha_out, C, T, dilation, t_lo, t_hi);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_pwconv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_resid,
                   const float* d_noise, const float* d_alpha_out, int epilogue, int B, int Cin, int Cout, int T,
                   int t_lo, int t_hi, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w, "vb_snac_pwconv: null pointer");
  VB_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "vb_snac_pwconv: epilogue %d", epilogue);
  VB_CHECK_ARG(epilogue != 1 || d_resid, "vb_snac_pwconv: residual epilogue needs d_resid");
  VB_CHECK_ARG(epilogue != 2 || (d_noise && Cin == Cout), "vb_snac_pwconv: noise epilogue needs noise and Cin == Cout");
  VB_CHECK_ARG(Cin % 4 == 0, "vb_snac_pwconv: Cin must be a multiple of 4");
  VB_CHECK_ARG(0 <= t_lo && t_lo <= t_hi && t_hi <= T, "vb_snac_pwconv: range [%d, %d) outside [0, %d)", t_lo, t_hi, T);
  if (B <= 0 || t_lo == t_hi) return 0;
  pwconv_kernel<<<dim3((t_hi - t_lo + GB_N - 1) / GB_N, (Cout + GB_M - 1) / GB_M, B), 256, 0,
                  static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w, d_bias, d_resid, d_noise, d_alpha_out, epilogue,
                                                       Cin, Cout, T, t_lo, t_hi);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias,
                   const float* d_alpha_out, int B, int Cin, int Cout, int T, int stride, int o_lo, int o_hi,
                   void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_packed, "vb_snac_convtr: null pointer");
  VB_CHECK_ARG(stride >= 1 && Cin % 2 == 0, "vb_snac_convtr: bad dims");
  VB_CHECK_ARG(0 <= o_lo && o_lo <= o_hi && o_hi <= T * stride, "vb_snac_convtr: output range [%d, %d) outside [0, %d)",
               o_lo, o_hi, T * stride);
  if (B <= 0 || o_lo == o_hi) return 0;
  const int pad = (stride + 1) / 2;
  // input positions n in [0, T] whose outputs n*stride + r - pad (r < stride) fall inside [o_lo, o_hi)
  int n_lo = (o_lo + pad - (stride - 1)) / stride;
  if (o_lo + pad - (stride - 1) < 0) n_lo = 0;
  int n_hi = (o_hi - 1 + pad) / stride + 1;
  if (n_hi > T + 1) n_hi = T + 1;
  convtr_kernel<<<dim3((n_hi - n_lo + GB_N - 1) / GB_N, (Cout + GB_M - 1) / GB_M, B * stride), 256, 0,
                  static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w_packed, d_bias, d_alpha_out, Cin, Cout, T, stride,
                                                       pad, n_lo, n_hi);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_final(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_alpha_in,
                  int B, int C, int T, int t0, int t1, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w && d_bias, "vb_snac_final: null pointer");
  VB_CHECK_ARG(0 <= t0 && t0 < t1 && t1 <= T, "vb_snac_final: bad output range [%d, %d) of %d", t0, t1, T);
  if (B <= 0) return 0;
  const size_t smem = static_cast<size_t>(C) * 8 * sizeof(float);
  final_conv_kernel<<<dim3((t1 - t0 + 255) / 256, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      d_y, d_x, d_w, d_bias, d_alpha_in, C, T, t0, t1);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
