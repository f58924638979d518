// SNAC decoder stages (vox_serve/tokenizer/snac.py:119-267, 297-357), fp32, activations [B][C][T] with T
// contiguous.  Weight-norm is folded on the host at load time (snac.py:244-249 re-derives it every
// forward).  Every Snake activation is fused into the epilogue of the kernel that produces its input
// (or, for the residual stream, into the depthwise conv's tile load), so the decoder is
//   from_codes -> dwconv7 -> pwconv(+snake) -> 4 x [ convtr -> pwconv(noise) -> 3 x (dwconv7(snake in,
//   snake out) -> pwconv(+residual[, +snake])) ] -> final conv7 + tanh.
// The two GEMM-shaped stages (1x1 convs, transposed convs) share one register-tiled fp32 kernel here; layers with at
// least 128 input channels (a multiple of 32) run on the tensor cores instead (snac_mma.cu, tf32 hi/lo split), the
// caller chooses per layer (vox_serve_b200/tokenizer/snac.py).
#include "../../include/vb_api.h"
#include "common.cuh"
#include "simt_gemm.cuh"

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
constexpr int DW_THREADS = 256;
// OPT outputs per thread (tile = OPT * 256 positions of one channel of one window): the long late stages use 4 so
// a launch is a few thousand CTAs instead of tens of thousands of 256-output ones (CTA turnover, not arithmetic or
// bytes, was what the 30-40 us of those launches paid for)
template <int OPT>
__global__ void __launch_bounds__(DW_THREADS) dwconv7_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             const float* __restrict__ alpha_in,
                                                             const float* __restrict__ alpha_out, int C, int T, int dil,
                                                             int t_lo, int t_hi) {
  extern __shared__ float tile[];  // OPT * DW_THREADS + 6*dil
  constexpr int TILE = OPT * DW_THREADS;
  const int c = blockIdx.y, b = blockIdx.z, t0 = t_lo + blockIdx.x * TILE;
  const float* xr = x + (static_cast<size_t>(b) * C + c) * T;
  const int halo = 3 * dil, n = min(TILE, t_hi - t0) + 2 * halo;
  const float ain = alpha_in ? alpha_in[c] : 0.f;
  for (int i = threadIdx.x; i < n; i += DW_THREADS) {
    const int t = t0 - halo + i;
    float v = 0.f;
    if (t >= 0 && t < T) {
      v = xr[t];
      if (alpha_in) v = snake_f(v, ain);
    }
    tile[i] = v;
  }
  __syncthreads();
  float wk[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) wk[k] = w[c * 7 + k];
  const float bv = bias ? bias[c] : 0.f, aout = alpha_out ? alpha_out[c] : 0.f;
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const int i = o * DW_THREADS + threadIdx.x, t = t0 + i;
    if (t < t_hi) {
      float acc = bv;
#pragma unroll
      for (int k = 0; k < 7; ++k) acc += wk[k] * tile[i + k * dil];
      if (alpha_out) acc = snake_f(acc, aout);
      y[(static_cast<size_t>(b) * C + c) * T + t] = acc;
    }
  }
}

// pointwise conv; epilogue 0 plain, 1 + resid, 2 noise block (x + noise * Wx); optional Snake after
template <int TM>
__global__ void __launch_bounds__(256) pwconv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     const float* __restrict__ resid, const float* __restrict__ noise,
                                                     const float* __restrict__ alpha_out, int epi, int B, int Cin,
                                                     int Cout, int T, int t_lo, int t_hi) {
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * T;
  gemm_tile_f32<TM>(
      w, Cout, Cin, Cin, t_lo, t_hi - t_lo, B * (t_hi - t_lo),
      [=](int k, int b, int n) { return x[b * xs + static_cast<size_t>(k) * T + n]; },
      [=](int m, int b, int n, float v) {
        const size_t o = b * ys + static_cast<size_t>(m) * T + n;
        if (bias) v += bias[m];
        if (epi == 1) v += resid[o];
        else if (epi == 2) v = x[o] + noise[static_cast<size_t>(b) * T + n] * v;
        if (alpha_out) v = snake_f(v, alpha_out[m]);
        y[o] = v;
      });
}

// transposed conv, kernel 2s, stride s, padding ceil(s/2), output_padding s%2 (snac.py:222-231).
// wp: host-repacked [s][Cout][2*Cin]: wp[r][co][tap*Cin + ci] = W[ci][co][r + tap*s].
// For phase r = blockIdx.z the GEMM column is the input position ti; output to = ti*s + r - pad.
template <int TM>
__global__ void __launch_bounds__(256) convtr_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ wp, const float* __restrict__ bias,
                                                     const float* __restrict__ alpha_out, int B, int Cin, int Cout,
                                                     int T, int s, int pad, int n_lo, int n_hi) {
  const int r = blockIdx.z;
  const int Tout = T * s;
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * Tout;
  const float* wr = wp + static_cast<size_t>(r) * Cout * 2 * Cin;
  // ti ranges over [0, T]: ti = T only receives the tap-1 term (x[T-1])
  gemm_tile_f32<TM>(
      wr, Cout, 2 * Cin, 2 * Cin, n_lo, n_hi - n_lo, B * (n_hi - n_lo),
      [=](int k, int b, int n) {
        const int tap = k >= Cin, ci = k - tap * Cin, ti = n - tap;
        return (ti >= 0 && ti < T) ? x[b * xs + static_cast<size_t>(ci) * T + ti] : 0.f;
      },
      [=](int m, int b, int n, float v) {
        const int to = n * s + r - pad;
        if (to < 0 || to >= Tout) return;
        if (bias) v += bias[m];
        if (alpha_out) v = snake_f(v, alpha_out[m]);
        y[b * ys + static_cast<size_t>(m) * Tout + to] = v;
      });
}

// final: tanh(conv_k7(x) + bias), Cout = 1; x already carries the last Snake unless alpha_in is given.  A CTA owns
// 256 outputs; channels are staged eight at a time as [8][256 + 6] tiles so every input element is loaded (and
// Snake-activated) once instead of once per tap.
constexpr int FC_CH = 8;
__global__ void __launch_bounds__(256) final_conv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ alpha_in, int C, int T, int t0,
                                                         int t1) {
  extern __shared__ float ws[];  // C*7 weights, C alphas, FC_CH * 262 tile
  float* al = ws + C * 7;
  float* tile = al + C;
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) al[i] = alpha_in ? alpha_in[i] : 0.f;
  const int b = blockIdx.y;
  const int tb = t0 + blockIdx.x * 256, t = tb + threadIdx.x;
  const float* xb = x + static_cast<size_t>(b) * C * T;
  float acc = bias[0];
  for (int c0 = 0; c0 < C; c0 += FC_CH) {
    __syncthreads();             // weights staged (first pass) / previous tile consumed
    for (int i = threadIdx.x; i < FC_CH * 262; i += 256) {
      const int cc = i / 262, j = i - cc * 262, tt = tb - 3 + j;
      float v = 0.f;
      if (c0 + cc < C && tt >= 0 && tt < T) {
        v = xb[static_cast<size_t>(c0 + cc) * T + tt];
        if (alpha_in) v = snake_f(v, al[c0 + cc]);
      }
      tile[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < FC_CH; ++cc) {
      if (c0 + cc < C) {
#pragma unroll
        for (int k = 0; k < 7; ++k) acc += ws[(c0 + cc) * 7 + k] * tile[cc * 262 + threadIdx.x + k];
      }
    }
  }
  if (t < t1) y[static_cast<size_t>(b) * (t1 - t0) + (t - t0)] = tanhf(acc);
}

}  // namespace vb

using namespace vb;

static int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}
static bool gemm_wide_tiles(int nx, int M, int nz) {
  if (M < 128) return false;
  return static_cast<long long>(nx) * ((M + 127) / 128) * nz >= 2LL * device_sm_count();
}

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
  const int range = t_hi - t_lo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (range >= 4 * DW_THREADS) {
    const size_t smem = (4 * DW_THREADS + 6 * dilation) * sizeof(float);
    dwconv7_kernel<4><<<dim3((range + 4 * DW_THREADS - 1) / (4 * DW_THREADS), C, B), DW_THREADS, smem, st>>>(
        d_y, d_x, d_w, d_bias, d_alpha_in, d_alpha_out, C, T, dilation, t_lo, t_hi);
  } else {
    const size_t smem = (DW_THREADS + 6 * dilation) * sizeof(float);
    dwconv7_kernel<1><<<dim3((range + DW_THREADS - 1) / DW_THREADS, C, B), DW_THREADS, smem, st>>>(
        d_y, d_x, d_w, d_bias, d_alpha_in, d_alpha_out, C, T, dilation, t_lo, t_hi);
  }
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
  const int cols = B * (t_hi - t_lo), nx = (cols + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 128-row tiles when they still give every SM two CTAs, else 64-row tiles
  if (gemm_wide_tiles(nx, Cout, 1))
    pwconv_kernel<8><<<dim3(nx, (Cout + 127) / 128, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_noise,
                                                                      d_alpha_out, epilogue, B, Cin, Cout, T, t_lo, t_hi);
  else
    pwconv_kernel<4><<<dim3(nx, (Cout + 63) / 64, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_noise,
                                                                    d_alpha_out, epilogue, B, Cin, Cout, T, t_lo, t_hi);
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
  const int cols = B * (n_hi - n_lo), nx = (cols + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gemm_wide_tiles(nx, Cout, stride))
    convtr_kernel<8><<<dim3(nx, (Cout + 127) / 128, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, d_alpha_out,
                                                                           B, Cin, Cout, T, stride, pad, n_lo, n_hi);
  else
    convtr_kernel<4><<<dim3(nx, (Cout + 63) / 64, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, d_alpha_out, B,
                                                                         Cin, Cout, T, stride, pad, n_lo, n_hi);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_final(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_alpha_in,
                  int B, int C, int T, int t0, int t1, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w && d_bias, "vb_snac_final: null pointer");
  VB_CHECK_ARG(0 <= t0 && t0 < t1 && t1 <= T, "vb_snac_final: bad output range [%d, %d) of %d", t0, t1, T);
  if (B <= 0) return 0;
  const size_t smem = (static_cast<size_t>(C) * 8 + FC_CH * 262) * sizeof(float);
  final_conv_kernel<<<dim3((t1 - t0 + 255) / 256, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      d_y, d_x, d_w, d_bias, d_alpha_in, C, T, t0, t1);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
