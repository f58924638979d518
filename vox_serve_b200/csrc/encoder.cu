// Elementwise / reduction stages of the GLM-4-Voice speech tokenizer (Whisper-style VQ encoder,
// vox_serve/encoder/glm.py:84-323) -- the STS prompt side.  The dense contractions run on gemm_bf16_kernel (convs as
// GEMMs over overlapping-row tensor maps, projections, the codebook product) and the block-causal attention on
// paged_prefill_attn_kernel (per-row key bounds); what is left is LayerNorm, GELU, layout, pooling and the arg-min.
// All activations are token-major [T][C] bf16, rounded where the reference's bf16 modules round.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 4
#include "common.cuh"

namespace vb {

constexpr int ENC_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < ENC_THREADS / 32; ++w) t += red[w];
  return t;
}

// h = bf16(h + delta) (if delta) ; y = bf16((h - mean) * rstd * w + b) (if y)      glm.py:195-214 (nn.LayerNorm)
__global__ void __launch_bounds__(ENC_THREADS) add_layernorm_kernel(__nv_bfloat16* __restrict__ y,
                                                                    __nv_bfloat16* __restrict__ h,
                                                                    const __nv_bfloat16* __restrict__ delta,
                                                                    const __nv_bfloat16* __restrict__ w,
                                                                    const __nv_bfloat16* __restrict__ b, int dim,
                                                                    float eps, int y_xt_tile) {
  pdl_sync();
  __shared__ float red[ENC_THREADS / 32];
  __nv_bfloat16* hr = h + static_cast<size_t>(blockIdx.x) * dim;
  const __nv_bfloat16* dr = delta ? delta + static_cast<size_t>(blockIdx.x) * dim : nullptr;
  float s = 0.f;
  for (int i = threadIdx.x; i < dim; i += ENC_THREADS) {
    float v = __bfloat162float(hr[i]);
    if (dr) {
      v = round_bf16(v + __bfloat162float(dr[i]));
      hr[i] = __float2bfloat16_rn(v);
    }
    s += v;
  }
  if (!y) return;
  const float mean = block_sum(s, red) / dim;
  float q = 0.f;
  for (int i = threadIdx.x; i < dim; i += ENC_THREADS) {      // each thread re-reads the elements it wrote itself
    const float v = __bfloat162float(hr[i]) - mean;
    q += v * v;
  }
  const float rstd = rsqrtf(block_sum(q, red) / dim + eps);
  // row-major, or the tiled activation layout the consuming projection streams with linear bulk copies
  const int row = blockIdx.x, num_kb = (dim + 63) >> 6;
  for (int i = threadIdx.x; i < dim; i += ENC_THREADS) {
    const size_t o = y_xt_tile ? xt_index(row, i, y_xt_tile, num_kb) : static_cast<size_t>(row) * dim + i;
    y[o] = __float2bfloat16_rn((__bfloat162float(hr[i]) - mean) * rstd * __bfloat162float(w[i]) + __bfloat162float(b[i]));
  }
}

// y = bf16(gelu(x)) (exact erf form, nn.GELU / F.gelu default) ; then y = bf16(y + add) if add     glm.py:288-295
__global__ void __launch_bounds__(ENC_THREADS) gelu_add_kernel(__nv_bfloat16* __restrict__ y,
                                                               const __nv_bfloat16* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ add, long long n,
                                                               int dim, int y_xt_tile) {
  pdl_sync();
  const long long stride = static_cast<long long>(gridDim.x) * ENC_THREADS;
  const int num_kb = (dim + 63) >> 6;
  for (long long i = static_cast<long long>(blockIdx.x) * ENC_THREADS + threadIdx.x; i < n; i += stride) {
    const float v = __bfloat162float(x[i]);
    float g = round_bf16(0.5f * v * (1.f + erff(v * 0.70710678118654752440f)));
    if (add) g = round_bf16(g + __bfloat162float(add[i]));
    size_t o = static_cast<size_t>(i);
    if (y_xt_tile) {
      const int row = static_cast<int>(i / dim);
      o = xt_index(row, static_cast<int>(i - static_cast<long long>(row) * dim), y_xt_tile, num_kb);
    }
    y[o] = __float2bfloat16_rn(g);
  }
}

// out[pad + t][c] = in[c][t], rows [0, pad) zero: channels-first features -> token-major rows with the causal
// convolution's left padding in place (glm.py:84-107)
__global__ void __launch_bounds__(ENC_THREADS) chw_to_rows_kernel(__nv_bfloat16* __restrict__ out,
                                                                  const __nv_bfloat16* __restrict__ in, int C, int T,
                                                                  int pad) {
  pdl_sync();
  __shared__ __nv_bfloat16 tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 rows of 32 per pass
  for (int j = ty; j < 32; j += ENC_THREADS / 32) {
    const int c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < C && t < T) ? in[static_cast<size_t>(c) * T + t] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += ENC_THREADS / 32) {
    const int t = t0 + j, c = c0 + tx;
    if (t < T && c < C) out[static_cast<size_t>(pad + t) * C + c] = tile[tx][j];
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < pad * 32; i += ENC_THREADS) {
      const int r = i / 32, c = c0 + (i & 31);
      if (c < C) out[static_cast<size_t>(r) * C + c] = __float2bfloat16_rn(0.f);
    }
}

// out[t][c] = bf16(sum_{j<k} in[t*k + j][c] / k), rows beyond T count as zeros (F.pad + AvgPool1d, glm.py:303-313)
__global__ void __launch_bounds__(ENC_THREADS) avgpool_rows_kernel(__nv_bfloat16* __restrict__ out,
                                                                   const __nv_bfloat16* __restrict__ in, int T, int D,
                                                                   int k) {
  pdl_sync();
  const int t = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += ENC_THREADS) {
    float s = 0.f;
    for (int j = 0; j < k; ++j) {
      const int r = t * k + j;
      if (r < T) s += __bfloat162float(in[static_cast<size_t>(r) * D + c]);
    }
    out[static_cast<size_t>(t) * D + c] = __float2bfloat16_rn(s / k);
  }
}

// ids[t] = first arg-min over n of bf16(bf16(c2[n] + x2[t]) - 2 * acc[t][n]), x2[t] = bf16(sum_d bf16(x[t][d]^2))
// (vector_quantize, glm.py:247-258: the distances are a bf16 tensor on the serving path, ties go to the first index)
__global__ void __launch_bounds__(ENC_THREADS) vq_argmin_kernel(int64_t* __restrict__ ids, const float* __restrict__ acc,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ c2, int N, int D) {
  pdl_sync();
  __shared__ float red[ENC_THREADS / 32];
  __shared__ float s_val[ENC_THREADS / 32];
  __shared__ int s_idx[ENC_THREADS / 32];
  const int t = blockIdx.x;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += ENC_THREADS) {
    const float v = __bfloat162float(x[static_cast<size_t>(t) * D + d]);
    s += round_bf16(v * v);
  }
  const float x2 = round_bf16(block_sum(s, red));
  float best = INFINITY;
  int bi = 0x7fffffff;
  const float* ar = acc + static_cast<size_t>(t) * N;
  for (int n = threadIdx.x; n < N; n += ENC_THREADS) {
    const float base = round_bf16(__bfloat162float(c2[n]) + x2);
    const float d = round_bf16(base - 2.f * ar[n]);
    if (d < best) { best = d; bi = n; }          // n ascends per thread: the first minimum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < ENC_THREADS / 32; ++w)
      if (s_val[w] < best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
    ids[t] = bi;
  }
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_add_layernorm(void* d_y, void* d_h, const void* d_delta, const void* d_w, const void* d_b, int rows, int dim,
                     float eps, int y_xt_tile, void* stream) {
  VB_CHECK_ARG(d_h && (d_y || d_delta), "vb_add_layernorm: nothing to do");
  VB_CHECK_ARG(!d_y || (d_w && d_b), "vb_add_layernorm: the norm output needs weight and bias");
  if (rows <= 0) return 0;
  VB_LAUNCH_PDL(add_layernorm_kernel, rows, ENC_THREADS, 0, stream, static_cast<__nv_bfloat16*>(d_y),
                static_cast<__nv_bfloat16*>(d_h), static_cast<const __nv_bfloat16*>(d_delta),
                static_cast<const __nv_bfloat16*>(d_w), static_cast<const __nv_bfloat16*>(d_b), dim, eps, y_xt_tile);
  return 0;
}

int vb_gelu_add(void* d_y, const void* d_x, const void* d_add, int64_t n, int dim, int y_xt_tile, void* stream) {
  VB_CHECK_ARG(d_y && d_x, "vb_gelu_add: null pointer");
  VB_CHECK_ARG(!y_xt_tile || (dim > 0 && n % dim == 0 && d_y != d_x), "vb_gelu_add: the tiled output needs the row width and its own buffer");
  if (n <= 0) return 0;
  long long blocks = (n + ENC_THREADS - 1) / ENC_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  VB_LAUNCH_PDL(gelu_add_kernel, static_cast<unsigned>(blocks), ENC_THREADS, 0, stream, static_cast<__nv_bfloat16*>(d_y),
                static_cast<const __nv_bfloat16*>(d_x), static_cast<const __nv_bfloat16*>(d_add),
                static_cast<long long>(n), dim > 0 ? dim : 1, y_xt_tile);
  return 0;
}

int vb_chw_to_rows(void* d_out, const void* d_in, int C, int T, int pad, void* stream) {
  VB_CHECK_ARG(d_out && d_in && C > 0 && pad >= 0, "vb_chw_to_rows: bad arguments");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(chw_to_rows_kernel, dim3((T + 31) / 32, (C + 31) / 32), ENC_THREADS, 0, stream,
                static_cast<__nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(d_in), C, T, pad);
  return 0;
}

int vb_avgpool_rows(void* d_out, const void* d_in, int T, int D, int k, void* stream) {
  VB_CHECK_ARG(d_out && d_in && k > 0, "vb_avgpool_rows: bad arguments");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(avgpool_rows_kernel, (T + k - 1) / k, ENC_THREADS, 0, stream, static_cast<__nv_bfloat16*>(d_out),
                static_cast<const __nv_bfloat16*>(d_in), T, D, k);
  return 0;
}

int vb_vq_argmin(int64_t* d_ids, const float* d_acc, const void* d_x, const void* d_c2, int T, int N, int D,
                 void* stream) {
  VB_CHECK_ARG(d_ids && d_acc && d_x && d_c2 && N > 0, "vb_vq_argmin: bad arguments");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(vq_argmin_kernel, T, ENC_THREADS, 0, stream, d_ids, d_acc, static_cast<const __nv_bfloat16*>(d_x),
                static_cast<const __nv_bfloat16*>(d_c2), N, D);
  return 0;
}

}  // extern "C"
