// Mimi decode stages (vox_serve/tokenizer/mimi.py:2993-3018: split RVQ decode -> learnt x2 upsampling -> 8-layer causal
// transformer -> SEANet decoder), fp32, activations [B][C][T] with T contiguous ("conv layout": the transformer's
// linears are 1x1 convolutions in it, so nothing is ever transposed).  Every chunk is decoded with zero left context,
// exactly as the reference's stateless StreamingConv1d / StreamingConvTranspose1d do (mimi.py:2116-2148, 2192-2215).
//
// All dense contractions (causal Conv1d k7 / k3 / k1, the transformer's linears, the causal ConvTranspose1d stack) go
// through the shared register-tiled fp32 GEMM tile (simt_gemm.cuh) with the convolution expressed in the B-operand
// fetch -- no im2col buffer -- and ELU / LayerScale / GELU / residual adds fused into fetch and store:
//   * conv:    A = W[Cout][Cin * k] (PyTorch layout, flattened), B(kk, b, t) = act(x[b][kk / k][t - (k-1 - kk % k) * dil])
//   * convtr:  one GEMM per output phase r < s with A = Wp[r][Cout][2 Cin], B(kk, b, n) = act(x[b][ci][n - tap]),
//              output t = n s + r (kernel 2 s: two taps per phase), rightmost K - S outputs never produced.
// This first version is SIMT fp32 (the tf32 hi/lo tcgen05 path of snac_mma.cu is the next step for the wide layers).
#include "../../include/vb_api.h"
#include "common.cuh"
#include "simt_gemm.cuh"

namespace vb {

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

enum MimiEpi { ME_PLAIN = 0, ME_RESID = 1, ME_SCALE_RESID = 2, ME_GELU = 3 };

// z[b][d][t] = sum_{k = k0 .. k1-1} emb[k][codes[b][k][t]][d], in ascending k (ResidualVectorQuantization.decode,
// mimi.py:482-490); emb = embedding_sum / clamp(cluster_usage, eps), prepared once on the host.  codes clamp to the table.
__global__ void __launch_bounds__(256) mimi_codes_sum_kernel(float* __restrict__ z, const long long* __restrict__ codes,
                                                             const float* __restrict__ emb, int K_total, int k0, int k1,
                                                             int bins, int D, int T) {
  const int t = blockIdx.x, b = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) {
      long long c = codes[(static_cast<size_t>(b) * K_total + k) * T + t];
      c = c < 0 ? 0 : (c >= bins ? bins - 1 : c);
      acc += emb[(static_cast<size_t>(k) * bins + c) * D + d];
    }
    z[(static_cast<size_t>(b) * D + d) * T + t] = acc;
  }
}

// causal Conv1d (kernel ksize, dilation dil, zero left context) / Linear in conv layout (ksize = 1)
template <int TM>
__global__ void __launch_bounds__(256) mimi_conv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ resid, const float* __restrict__ scale,
                                                        int epi, int elu_in, int B, int Cin, int Cout, int T, int ksize,
                                                        int dil) {
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * T;
  gemm_tile_f32<TM>(
      w, Cout, Cin * ksize, Cin * ksize, 0, T, B * T,
      [=](int kk, int b, int n) {
        const int ci = kk / ksize, j = kk - ci * ksize;
        const int t = n - (ksize - 1 - j) * dil;
        if (t < 0) return 0.f;
        const float v = x[b * xs + static_cast<size_t>(ci) * T + t];
        return elu_in ? elu_f(v) : v;
      },
      [=](int m, int b, int n, float v) {
        const size_t o = b * ys + static_cast<size_t>(m) * T + n;
        if (bias) v += bias[m];
        if (epi == ME_RESID) v = resid[o] + v;
        else if (epi == ME_SCALE_RESID) v = resid[o] + scale[m] * v;
        else if (epi == ME_GELU) v = gelu_f(v);
        y[o] = v;
      });
}

// causal ConvTranspose1d, kernel 2 s, stride s: y[b][co][n s + r] = bias + sum_ci x[ci][n] W[ci][co][r] + x[ci][n-1] W[ci][co][r+s]
// wp: host-repacked [s][Cout][2 Cin]: wp[r][co][tap * Cin + ci] = W[ci][co][r + tap * s].  Tout = T s.
template <int TM>
__global__ void __launch_bounds__(256) mimi_convtr_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                          const float* __restrict__ wp, const float* __restrict__ bias,
                                                          int elu_in, int B, int Cin, int Cout, int T, int s) {
  const int r = blockIdx.z;
  const int Tout = T * s;
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * Tout;
  const float* wr = wp + static_cast<size_t>(r) * Cout * 2 * Cin;
  gemm_tile_f32<TM>(
      wr, Cout, 2 * Cin, 2 * Cin, 0, T, B * T,
      [=](int k, int b, int n) {
        const int tap = k >= Cin, ci = k - tap * Cin, ti = n - tap;
        if (ti < 0) return 0.f;
        const float v = x[b * xs + static_cast<size_t>(ci) * T + ti];
        return elu_in ? elu_f(v) : v;
      },
      [=](int m, int b, int n, float v) {
        if (bias) v += bias[m];
        y[b * ys + static_cast<size_t>(m) * Tout + n * s + r] = v;
      });
}

// channel-wise (groups = C) causal ConvTranspose1d, kernel 2 s, stride s, no bias: ConvTrUpsample1d (mimi.py:2272-2323)
__global__ void __launch_bounds__(256) mimi_upsample_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                            const float* __restrict__ w, int C, int T, int s) {
  const int c = blockIdx.x, b = blockIdx.y;
  const float* xr = x + (static_cast<size_t>(b) * C + c) * T;
  float* yr = y + (static_cast<size_t>(b) * C + c) * T * s;
  const float* wc = w + static_cast<size_t>(c) * 2 * s;
  for (int to = threadIdx.x; to < T * s; to += blockDim.x) {
    const int n = to / s, r = to - n * s;
    float v = xr[n] * wc[r];
    if (n > 0) v += xr[n - 1] * wc[r + s];
    yr[to] = v;
  }
}

// LayerNorm over the channel axis of [B][C][T] (nn.LayerNorm(C, eps) applied to the [B, T, C] view): one warp per (b, t)
__global__ void __launch_bounds__(256) mimi_layernorm_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int C, int T, float eps) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const float* xb = x + static_cast<size_t>(b) * C * T + t;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xb[static_cast<size_t>(c) * T];
  const float mean = warp_sum(s) / static_cast<float>(C);
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xb[static_cast<size_t>(c) * T] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / static_cast<float>(C) + eps);
  float* yb = y + static_cast<size_t>(b) * C * T + t;
  for (int c = lane; c < C; c += 32) yb[static_cast<size_t>(c) * T] = (xb[static_cast<size_t>(c) * T] - mean) * rstd * w[c] + bias[c];
}

// causal self-attention of one (head, batch item) over a chunk: qkv [B][3 C][T] with channel = p * C + h * D + d (the
// "(p h d)" packing of in_proj, mimi.py:1519-1523), interleaved-pair RoPE at offset 0 (apply_rope :874-930), softmax in
// fp32, out [B][C][T].  T <= 64, D <= 128: a chunk is 2 x detokenize_interval = 20 positions.
__global__ void __launch_bounds__(256) mimi_attention_kernel(float* __restrict__ out, const float* __restrict__ qkv, int C,
                                                             int H, int T, float max_period) {
  extern __shared__ float sm[];       // q[D][T], k[D][T], v[D][T], p[T][T]
  const int h = blockIdx.x, b = blockIdx.y, D = C / H;
  float* q = sm;
  float* k = q + D * T;
  float* v = k + D * T;
  float* p = v + D * T;
  const float* base = qkv + static_cast<size_t>(b) * 3 * C * T;
  // load with RoPE: pair j = (2j, 2j+1), angle t * exp(-ln(P) * 2 j / D)
  for (int i = threadIdx.x; i < (D / 2) * T; i += blockDim.x) {
    const int j = i / T, t = i - j * T;
    const float freq = expf(static_cast<float>(j) * (-logf(max_period) * 2.f / static_cast<float>(D)));
    float sn, cs;
    sincosf(static_cast<float>(t) * freq, &sn, &cs);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* src = base + (static_cast<size_t>(which) * C + h * D + 2 * j) * T + t;
      const float xr = src[0], xi = src[T];
      float* dst = (which == 0 ? q : k) + (2 * j) * T + t;
      dst[0] = xr * cs - xi * sn;
      dst[T] = xr * sn + xi * cs;
    }
  }
  for (int i = threadIdx.x; i < D * T; i += blockDim.x) v[i] = base[(static_cast<size_t>(2) * C + h * D) * T + i];
  __syncthreads();
  const float scale = rsqrtf(static_cast<float>(D));
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int tq = i / T, tk = i - tq * T;
    float s = -INFINITY;
    if (tk <= tq) {
      s = 0.f;
      for (int d = 0; d < D; ++d) s += q[d * T + tq] * k[d * T + tk];
      s *= scale;
    }
    p[i] = s;
  }
  __syncthreads();
  for (int tq = threadIdx.x; tq < T; tq += blockDim.x) {
    float mx = -INFINITY;
    for (int tk = 0; tk <= tq; ++tk) mx = fmaxf(mx, p[tq * T + tk]);
    float den = 0.f;
    for (int tk = 0; tk <= tq; ++tk) {
      const float e = expf(p[tq * T + tk] - mx);
      p[tq * T + tk] = e;
      den += e;
    }
    const float inv = 1.f / den;
    for (int tk = 0; tk < T; ++tk) p[tq * T + tk] = tk <= tq ? p[tq * T + tk] * inv : 0.f;
  }
  __syncthreads();
  float* ob = out + (static_cast<size_t>(b) * C + h * D) * T;
  for (int i = threadIdx.x; i < D * T; i += blockDim.x) {
    const int d = i / T, tq = i - d * T;
    float acc = 0.f;
    for (int tk = 0; tk <= tq; ++tk) acc += p[tq * T + tk] * v[d * T + tk];
    ob[i] = acc;
  }
}

static int mimi_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}
static bool mimi_wide(int nx, int M, int nz) {
  return M >= 128 && static_cast<long long>(nx) * ((M + 127) / 128) * nz >= 2LL * mimi_sm_count();
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_mimi_codes_sum(float* d_z, const int64_t* d_codes, const float* d_emb, int B, int K_total, int k0, int k1, int bins,
                      int D, int T, void* stream) {
  VB_CHECK_ARG(d_z && d_codes && d_emb, "vb_mimi_codes_sum: null pointer");
  VB_CHECK_ARG(0 <= k0 && k0 < k1 && k1 <= K_total && bins > 0 && D > 0, "vb_mimi_codes_sum: bad codebook range");
  if (B <= 0 || T <= 0) return 0;
  mimi_codes_sum_kernel<<<dim3(T, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_z, reinterpret_cast<const long long*>(d_codes), d_emb, K_total, k0, k1, bins, D, T);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_mimi_conv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_resid,
                 const float* d_scale, int epilogue, int elu_in, int B, int Cin, int Cout, int T, int ksize, int dilation,
                 void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w, "vb_mimi_conv: null pointer");
  VB_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "vb_mimi_conv: epilogue %d", epilogue);
  VB_CHECK_ARG((epilogue != ME_RESID && epilogue != ME_SCALE_RESID) || d_resid, "vb_mimi_conv: residual epilogue needs d_resid");
  VB_CHECK_ARG(epilogue != ME_SCALE_RESID || d_scale, "vb_mimi_conv: LayerScale epilogue needs d_scale");
  VB_CHECK_ARG(ksize >= 1 && dilation >= 1 && (Cin * ksize) % 4 == 0, "vb_mimi_conv: Cin * ksize must be a multiple of 4");
  if (B <= 0 || T <= 0) return 0;
  const int nx = (B * T + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mimi_wide(nx, Cout, 1))
    mimi_conv_kernel<8><<<dim3(nx, (Cout + 127) / 128, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_scale, epilogue,
                                                                         elu_in, B, Cin, Cout, T, ksize, dilation);
  else
    mimi_conv_kernel<4><<<dim3(nx, (Cout + 63) / 64, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_scale, epilogue,
                                                                       elu_in, B, Cin, Cout, T, ksize, dilation);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_mimi_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias, int elu_in, int B, int Cin,
                   int Cout, int T, int stride, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_packed, "vb_mimi_convtr: null pointer");
  VB_CHECK_ARG(stride >= 1 && Cin % 2 == 0, "vb_mimi_convtr: bad dims");
  if (B <= 0 || T <= 0) return 0;
  const int nx = (B * T + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mimi_wide(nx, Cout, stride))
    mimi_convtr_kernel<8><<<dim3(nx, (Cout + 127) / 128, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, elu_in, B, Cin,
                                                                                Cout, T, stride);
  else
    mimi_convtr_kernel<4><<<dim3(nx, (Cout + 63) / 64, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, elu_in, B, Cin,
                                                                              Cout, T, stride);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_mimi_upsample(float* d_y, const float* d_x, const float* d_w, int B, int C, int T, int stride, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w && stride >= 1, "vb_mimi_upsample: bad arguments");
  if (B <= 0 || T <= 0) return 0;
  mimi_upsample_kernel<<<dim3(C, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w, C, T, stride);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_mimi_layernorm(float* d_y, const float* d_x, const float* d_w, const float* d_bias, int B, int C, int T, float eps,
                      void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w && d_bias, "vb_mimi_layernorm: null pointer");
  if (B <= 0 || T <= 0) return 0;
  mimi_layernorm_kernel<<<dim3((T + 7) / 8, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w, d_bias, C, T, eps);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_mimi_attention(float* d_out, const float* d_qkv, int B, int C, int H, int T, float max_period, void* stream) {
  VB_CHECK_ARG(d_out && d_qkv, "vb_mimi_attention: null pointer");
  VB_CHECK_ARG(H > 0 && C % H == 0 && (C / H) % 2 == 0 && C / H <= 128 && T >= 1 && T <= 64,
               "vb_mimi_attention: head_dim %d (even, <= 128) / T %d (<= 64) unsupported", H ? C / H : 0, T);
  if (B <= 0) return 0;
  const int D = C / H;
  const size_t smem = (static_cast<size_t>(3) * D * T + static_cast<size_t>(T) * T) * sizeof(float);
  if (smem > 48 * 1024)
    VB_CHECK_CUDA(cudaFuncSetAttribute(mimi_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  mimi_attention_kernel<<<dim3(H, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(d_out, d_qkv, C, H, T, max_period);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
