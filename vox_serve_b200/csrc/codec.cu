// Streaming codec-decoder stages with per-request caches: the Qwen3-TTS 12 Hz decoder
// (vox_serve/tokenizer/qwen3_codec.py:1541-1667: forward_chunk; :239-340 causal conv cache, :359-397 transposed-conv cache,
// :434-468 ConvNeXt, :573-655 sliding-window attention cache, :1004-1018 SnakeBeta).  fp32, activations [B][C][T] with T
// contiguous, like mimi.cu; every dense contraction goes through the register-tiled fp32 GEMM tile (simt_gemm.cuh) with the
// convolution, its LEFT CONTEXT (the cache of the previous chunk instead of zeros) and the input activation (SnakeBeta /
// ELU) expressed in the B-operand fetch.  The caches hold ACTIVATED inputs, exactly as the reference's do (its activation
// runs before forward_chunk copies the tail into the cache), so context reads are not activated again; vb_codec_cache_update
// applies the activation when it refreshes a cache from the raw chunk.
#include "../../include/vb_api.h"
#include "common.cuh"
#include "simt_gemm.cuh"

namespace vb {

enum CodecAct { CA_NONE = 0, CA_ELU = 1, CA_SNAKE = 2 };
enum CodecEpi { CE_PLAIN = 0, CE_RESID = 1, CE_SCALE_RESID = 2, CE_GELU = 3, CE_SILU = 4, CE_MUL = 5, CE_CLAMP = 6 };

// SnakeBeta with host-prepared per-channel a = exp(alpha), ib = 1 / (exp(beta) + 1e-9):  x + ib * sin^2(a x)
__device__ __forceinline__ float codec_act(float v, int act, const float* __restrict__ a, const float* __restrict__ ib, int c) {
  if (act == CA_SNAKE) {
    const float s = sinf(v * a[c]);
    return v + ib[c] * (s * s);
  }
  if (act == CA_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

// causal Conv1d (kernel ksize, dilation dil): input position t - (ksize-1-j) dil; negative positions come from
// ctx [B][Cin][pad] (pad = (ksize-1) dil; NULL = zeros), which already holds activated values
template <int TM>
__global__ void __launch_bounds__(256) codec_conv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ resid, const float* __restrict__ scale,
                                                         const float* __restrict__ ctx, const float* __restrict__ act_a,
                                                         const float* __restrict__ act_ib, int epi, int act, int B, int Cin,
                                                         int Cout, int T, int ksize, int dil) {
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * T;
  const int pad = (ksize - 1) * dil;
  gemm_tile_f32<TM>(
      w, Cout, Cin * ksize, Cin * ksize, 0, T, B * T,
      [=](int kk, int b, int n) {
        const int ci = kk / ksize, j = kk - ci * ksize;
        const int t = n - (ksize - 1 - j) * dil;
        if (t < 0) return ctx ? ctx[(static_cast<size_t>(b) * Cin + ci) * pad + (pad + t)] : 0.f;
        return codec_act(x[b * xs + static_cast<size_t>(ci) * T + t], act, act_a, act_ib, ci);
      },
      [=](int m, int b, int n, float v) {
        const size_t o = b * ys + static_cast<size_t>(m) * T + n;
        if (bias) v += bias[m];
        switch (epi) {
          case CE_RESID: v = resid[o] + v; break;
          case CE_SCALE_RESID: v = resid[o] + scale[m] * v; break;
          case CE_GELU: v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); break;
          case CE_SILU: v = v / (1.f + expf(-v)); break;
          case CE_MUL: v = resid[o] * v; break;
          case CE_CLAMP: v = fminf(1.f, fmaxf(-1.f, v)); break;
          default: break;
        }
        y[o] = v;
      });
}

// causal ConvTranspose1d, kernel 2 s, stride s:  y[n s + r] = bias + act(x[n]) W[r] + prev W[r + s], prev = act(x[n-1]) or, for
// n = 0, ctx [B][Cin][1] (NULL = 0).  wp: [s][Cout][2 Cin] packed per output phase (tap 0 | tap 1)
template <int TM>
__global__ void __launch_bounds__(256) codec_convtr_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                           const float* __restrict__ wp, const float* __restrict__ bias,
                                                           const float* __restrict__ ctx, const float* __restrict__ act_a,
                                                           const float* __restrict__ act_ib, int act, int B, int Cin,
                                                           int Cout, int T, int s) {
  const int r = blockIdx.z, Tout = T * s;
  const size_t xs = static_cast<size_t>(Cin) * T, ys = static_cast<size_t>(Cout) * Tout;
  const float* wr = wp + static_cast<size_t>(r) * Cout * 2 * Cin;
  gemm_tile_f32<TM>(
      wr, Cout, 2 * Cin, 2 * Cin, 0, T, B * T,
      [=](int k, int b, int n) {
        const int tap = k >= Cin, ci = k - tap * Cin, ti = n - tap;
        if (ti < 0) return ctx ? ctx[static_cast<size_t>(b) * Cin + ci] : 0.f;
        return codec_act(x[b * xs + static_cast<size_t>(ci) * T + ti], act, act_a, act_ib, ci);
      },
      [=](int m, int b, int n, float v) {
        if (bias) v += bias[m];
        y[b * ys + static_cast<size_t>(m) * Tout + n * s + r] = v;
      });
}

// cache [B][C][pad] <- the last `pad` ACTIVATED inputs after appending the chunk x [B][C][L]
// (CausalConvNet.forward_chunk :318-325; pad = 1: the transposed convolution's one-sample cache :391).  One thread per (b, c);
// the shift runs front to back so every old value is read before its slot is overwritten.
__global__ void codec_cache_update_kernel(float* __restrict__ cache, const float* __restrict__ x,
                                          const float* __restrict__ act_a, const float* __restrict__ act_ib, int act, int B,
                                          int C, int L, int pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C;
  float* cc = cache + static_cast<size_t>(i) * pad;
  const float* xr = x + static_cast<size_t>(i) * L;
  if (L >= pad) {
    for (int p = 0; p < pad; ++p) cc[p] = codec_act(xr[L - pad + p], act, act_a, act_ib, c);
  } else {
    const int keep = pad - L;
    for (int p = 0; p < keep; ++p) cc[p] = cc[p + L];
    for (int p = 0; p < L; ++p) cc[keep + p] = codec_act(xr[p], act, act_a, act_ib, c);
  }
}

// y = act(x) elementwise over [B][C][T]: the tensor-core convolutions take pre-activated input (a k-tap convolution would
// otherwise evaluate SnakeBeta k times per element in its operand loaders, which made them loader-bound)
__global__ void __launch_bounds__(256) codec_activate_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                             const float* __restrict__ act_a, const float* __restrict__ act_ib,
                                                             int act, int C, int T) {
  const int c = blockIdx.x % C;
  const size_t row = static_cast<size_t>(blockIdx.x) * T;
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < T; t += gridDim.y * blockDim.x)
    y[row + t] = codec_act(x[row + t], act, act_a, act_ib, c);
}

// depthwise causal Conv1d (groups = C, kernel ksize, dilation 1) with left context: the ConvNeXt block's dwconv
__global__ void __launch_bounds__(256) codec_dwconv_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ ctx, int C, int T, int ksize) {
  const int c = blockIdx.x, b = blockIdx.y, pad = ksize - 1;
  const float* xr = x + (static_cast<size_t>(b) * C + c) * T;
  const float* cr = ctx ? ctx + (static_cast<size_t>(b) * C + c) * pad : nullptr;
  const float* wc = w + static_cast<size_t>(c) * ksize;
  float* yr = y + (static_cast<size_t>(b) * C + c) * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float acc = bias ? bias[c] : 0.f;
    for (int j = 0; j < ksize; ++j) {
      const int ti = t - (pad - j);
      const float v = ti >= 0 ? xr[ti] : (cr ? cr[pad + ti] : 0.f);
      acc += v * wc[j];
    }
    yr[t] = acc;
  }
}

// RMSNorm over the channel axis of [B][C][T]: w * (x * rsqrt(mean(x^2) + eps)) (Qwen3TTSTokenizerV2DecoderRMSNorm :713-718);
// one warp per (b, t)
__global__ void __launch_bounds__(256) codec_rmsnorm_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                            const float* __restrict__ w, int C, int T, float eps) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const float* xb = x + static_cast<size_t>(b) * C * T + t;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = xb[static_cast<size_t>(c) * T];
    s += v * v;
  }
  const float r = rsqrtf(warp_sum(s) / static_cast<float>(C) + eps);
  float* yb = y + static_cast<size_t>(b) * C * T + t;
  for (int c = lane; c < C; c += 32) yb[static_cast<size_t>(c) * T] = w[c] * (xb[static_cast<size_t>(c) * T] * r);
}

// One (kv head, batch item) of DecoderAttention.forward_chunk (:573-655): rotate-half RoPE of the chunk's q / k at positions
// pos0[b] + t, the head's cache slice [W][2 D] (K | V per slot) shifted left by T with the new K | V appended, then every
// query head of the group attends to ALL W slots (slots never written hold zeros and take softmax mass like any other key --
// the reference's zero-initialised cache) under the chunk-causal mask  j <= W - T + t.
// qkv [B][(H + 2 Hkv) D][T] (q heads | k heads | v heads); out [B][H D][T]; T < W.  cache: [Hkv][W][2 D] per batch item,
// items cache_batch_stride elements apart (one layer of the reference's [B][layers][Hkv][W][2 D] tensor).
__global__ void __launch_bounds__(256) codec_attn_chunk_kernel(float* __restrict__ out, const float* __restrict__ qkv,
                                                               float* __restrict__ cache, long long cache_batch_stride,
                                                               const long long* __restrict__ pos0, int H, int Hkv, int D,
                                                               int T, int W, float theta) {
  extern __shared__ float sm[];        // kv[W][2 D] | q[T][D] | p[T][W]
  const int hk = blockIdx.x, b = blockIdx.y, G = H / Hkv, half = D / 2;
  float* kv = sm;
  float* q = kv + W * 2 * D;
  float* p = q + T * D;
  float* cslice = cache + static_cast<size_t>(b) * cache_batch_stride + static_cast<size_t>(hk) * W * 2 * D;
  const size_t rows = static_cast<size_t>(H + 2 * Hkv) * D;
  const float* base = qkv + static_cast<size_t>(b) * rows * T;
  const float* kb = base + (static_cast<size_t>(H + hk) * D) * T;
  const float* vb_ = base + (static_cast<size_t>(H + Hkv + hk) * D) * T;
  const float p0 = static_cast<float>(pos0[b]);
  // shifted cache
  for (int i = threadIdx.x; i < (W - T) * 2 * D; i += blockDim.x) kv[i] = cslice[i + T * 2 * D];
  // new K (rotated) | V
  for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
    const int t = i / D, d = i - t * D;
    const int j = d < half ? d : d - half;
    const float inv = powf(theta, -static_cast<float>(2 * j) / static_cast<float>(D));
    float sn, cs;
    sincosf((p0 + static_cast<float>(t)) * inv, &sn, &cs);
    const float xk = kb[static_cast<size_t>(d) * T + t];
    const float pk = d < half ? -kb[static_cast<size_t>(d + half) * T + t] : kb[static_cast<size_t>(d - half) * T + t];
    kv[(W - T + t) * 2 * D + d] = xk * cs + pk * sn;
    kv[(W - T + t) * 2 * D + D + d] = vb_[static_cast<size_t>(d) * T + t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W * 2 * D; i += blockDim.x) cslice[i] = kv[i];
  const float scale = rsqrtf(static_cast<float>(D));
  for (int g = 0; g < G; ++g) {
    const int h = hk * G + g;
    const float* qb = base + (static_cast<size_t>(h) * D) * T;
    __syncthreads();
    for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
      const int t = i / D, d = i - t * D;
      const int j = d < half ? d : d - half;
      const float inv = powf(theta, -static_cast<float>(2 * j) / static_cast<float>(D));
      float sn, cs;
      sincosf((p0 + static_cast<float>(t)) * inv, &sn, &cs);
      const float xq = qb[static_cast<size_t>(d) * T + t];
      const float pq = d < half ? -qb[static_cast<size_t>(d + half) * T + t] : qb[static_cast<size_t>(d - half) * T + t];
      q[t * D + d] = xq * cs + pq * sn;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * W; i += blockDim.x) {
      const int t = i / W, j = i - t * W;
      float s = -INFINITY;
      if (j <= W - T + t) {
        s = 0.f;
        for (int d = 0; d < D; ++d) s += q[t * D + d] * kv[j * 2 * D + d];
        s *= scale;
      }
      p[i] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int last = W - T + t;
      float mx = -INFINITY;
      for (int j = 0; j <= last; ++j) mx = fmaxf(mx, p[t * W + j]);
      float den = 0.f;
      for (int j = 0; j <= last; ++j) {
        const float e = expf(p[t * W + j] - mx);
        p[t * W + j] = e;
        den += e;
      }
      const float inv = 1.f / den;
      for (int j = 0; j <= last; ++j) p[t * W + j] *= inv;
    }
    __syncthreads();
    float* ob = out + (static_cast<size_t>(b) * H + h) * D * T;
    for (int i = threadIdx.x; i < D * T; i += blockDim.x) {
      const int d = i / T, t = i - d * T;
      float acc = 0.f;
      for (int j = 0; j <= W - T + t; ++j) acc += p[t * W + j] * kv[j * 2 * D + D + d];
      ob[i] = acc;
    }
  }
}

static int codec_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}
static bool codec_wide(int nx, int M, int nz) {
  return M >= 128 && static_cast<long long>(nx) * ((M + 127) / 128) * nz >= 2LL * codec_sm_count();
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_codec_conv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_resid,
                  const float* d_scale, const float* d_ctx, const float* d_act_a, const float* d_act_ib, int epilogue,
                  int act_in, int B, int Cin, int Cout, int T, int ksize, int dilation, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w, "vb_codec_conv: null pointer");
  VB_CHECK_ARG(epilogue >= 0 && epilogue <= 6 && act_in >= 0 && act_in <= 2, "vb_codec_conv: epilogue %d / activation %d", epilogue,
               act_in);
  VB_CHECK_ARG((epilogue != CE_RESID && epilogue != CE_SCALE_RESID && epilogue != CE_MUL) || d_resid,
               "vb_codec_conv: this epilogue needs d_resid");
  VB_CHECK_ARG(epilogue != CE_SCALE_RESID || d_scale, "vb_codec_conv: LayerScale epilogue needs d_scale");
  VB_CHECK_ARG(act_in != CA_SNAKE || (d_act_a && d_act_ib), "vb_codec_conv: SnakeBeta needs its two per-channel tables");
  VB_CHECK_ARG(ksize >= 1 && dilation >= 1 && (Cin * ksize) % 4 == 0, "vb_codec_conv: Cin * ksize must be a multiple of 4");
  if (B <= 0 || T <= 0) return 0;
  const int nx = (B * T + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (codec_wide(nx, Cout, 1))
    codec_conv_kernel<8><<<dim3(nx, (Cout + 127) / 128, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_scale, d_ctx, d_act_a,
                                                                          d_act_ib, epilogue, act_in, B, Cin, Cout, T, ksize,
                                                                          dilation);
  else
    codec_conv_kernel<4><<<dim3(nx, (Cout + 63) / 64, 1), 256, 0, st>>>(d_y, d_x, d_w, d_bias, d_resid, d_scale, d_ctx, d_act_a,
                                                                        d_act_ib, epilogue, act_in, B, Cin, Cout, T, ksize,
                                                                        dilation);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias, const float* d_ctx,
                    const float* d_act_a, const float* d_act_ib, int act_in, int B, int Cin, int Cout, int T, int stride,
                    void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_packed, "vb_codec_convtr: null pointer");
  VB_CHECK_ARG(stride >= 1 && Cin % 2 == 0 && act_in >= 0 && act_in <= 2, "vb_codec_convtr: bad arguments");
  VB_CHECK_ARG(act_in != CA_SNAKE || (d_act_a && d_act_ib), "vb_codec_convtr: SnakeBeta needs its two per-channel tables");
  if (B <= 0 || T <= 0) return 0;
  const int nx = (B * T + GB_N - 1) / GB_N;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (codec_wide(nx, Cout, stride))
    codec_convtr_kernel<8><<<dim3(nx, (Cout + 127) / 128, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, d_ctx, d_act_a,
                                                                                 d_act_ib, act_in, B, Cin, Cout, T, stride);
  else
    codec_convtr_kernel<4><<<dim3(nx, (Cout + 63) / 64, stride), 256, 0, st>>>(d_y, d_x, d_w_packed, d_bias, d_ctx, d_act_a,
                                                                               d_act_ib, act_in, B, Cin, Cout, T, stride);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_cache_update(float* d_cache, const float* d_x, const float* d_act_a, const float* d_act_ib, int act_in, int B,
                          int C, int L, int pad, void* stream) {
  VB_CHECK_ARG(d_cache && d_x && pad >= 1 && L >= 1, "vb_codec_cache_update: bad arguments");
  VB_CHECK_ARG(act_in != CA_SNAKE || (d_act_a && d_act_ib), "vb_codec_cache_update: SnakeBeta needs its two per-channel tables");
  if (B <= 0 || C <= 0) return 0;
  codec_cache_update_kernel<<<(B * C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_cache, d_x, d_act_a, d_act_ib,
                                                                                              act_in, B, C, L, pad);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_activate(float* d_y, const float* d_x, const float* d_act_a, const float* d_act_ib, int act_in, int B, int C, int T,
                      void* stream) {
  VB_CHECK_ARG(d_y && d_x && act_in >= 0 && act_in <= 2, "vb_codec_activate: bad arguments");
  VB_CHECK_ARG(act_in != CA_SNAKE || (d_act_a && d_act_ib), "vb_codec_activate: SnakeBeta needs its two per-channel tables");
  if (B <= 0 || C <= 0 || T <= 0) return 0;
  const int gy = T >= 4096 ? 8 : (T >= 1024 ? 2 : 1);
  codec_activate_kernel<<<dim3(B * C, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_act_a, d_act_ib, act_in, C, T);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_dwconv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_ctx, int B, int C,
                    int T, int ksize, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w && ksize >= 1, "vb_codec_dwconv: bad arguments");
  if (B <= 0 || T <= 0) return 0;
  codec_dwconv_kernel<<<dim3(C, B), 64, 0, static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w, d_bias, d_ctx, C, T, ksize);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_rmsnorm(float* d_y, const float* d_x, const float* d_w, int B, int C, int T, float eps, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w, "vb_codec_rmsnorm: null pointer");
  if (B <= 0 || T <= 0) return 0;
  codec_rmsnorm_kernel<<<dim3((T + 7) / 8, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_w, C, T, eps);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_codec_attn_chunk(float* d_out, const float* d_qkv, float* d_cache, int64_t cache_batch_stride, const int64_t* d_pos0,
                        int B, int H, int Hkv, int D, int T, int W, float theta, void* stream) {
  VB_CHECK_ARG(d_out && d_qkv && d_cache && d_pos0, "vb_codec_attn_chunk: null pointer");
  VB_CHECK_ARG(cache_batch_stride >= static_cast<int64_t>(Hkv) * W * 2 * D, "vb_codec_attn_chunk: cache batch stride too small");
  VB_CHECK_ARG(H > 0 && Hkv > 0 && H % Hkv == 0 && D % 2 == 0 && T >= 1 && T < W,
               "vb_codec_attn_chunk: needs H %% Hkv == 0, even head_dim and a chunk shorter than the window (T %d, W %d)", T, W);
  const size_t smem = (static_cast<size_t>(W) * 2 * D + static_cast<size_t>(T) * D + static_cast<size_t>(T) * W) * sizeof(float);
  VB_CHECK_ARG(smem <= VB_MAX_DYN_SMEM, "vb_codec_attn_chunk: %zu bytes of shared memory", smem);
  if (B <= 0) return 0;
  if (smem > 48 * 1024)
    VB_CHECK_CUDA(cudaFuncSetAttribute(codec_attn_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  codec_attn_chunk_kernel<<<dim3(Hkv, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      d_out, d_qkv, d_cache, static_cast<long long>(cache_batch_stride), reinterpret_cast<const long long*>(d_pos0), H, Hkv, D, T,
      W, theta);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
