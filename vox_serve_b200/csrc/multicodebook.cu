// Multi-codebook glue for the depth-transformer models (CSM, Qwen3-TTS): what the reference adapters do with a handful
// of torch ops per frame between the backbone and the depth decoder (vox_serve/model/csm.py:637-663, 665-769;
// vox_serve/worker/cuda_graph_worker.py:1058-1160), as single launches that read their indices from device memory, so
// that a whole frame -- backbone step, codebook-0 sample, 31 depth steps with their samples -- is one CUDA graph.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 8
#include "common.cuh"

namespace vb {

// out[t][:] = bf16( sum_c mask(t, c) * table(c)[ids(t, c) + offset(c)] ), fp32 accumulation in column order, one
// rounding at the end: `(embeds * masks[:, :, None]).sum(dim=1)` of csm.py:647-654 (torch sums bf16 in fp32).
// Columns c < n_cols_a come from table_a with row offset (col0 + c) * col_offset (CsmBackboneModelEmbeddings:
// ids + arange(N) * vocab, csm.py:158-168); columns >= n_cols_a from table_b without offset (the text stream).
// ids are int64, element (t, c) at ids[t * ld_t + c * ld_c]; masks (uint8, same strides as [T][C] row-major) may be NULL.
__global__ void __launch_bounds__(256) multi_embed_sum_kernel(__nv_bfloat16* __restrict__ out, int ld_out,
                                                              const long long* __restrict__ ids, long long ld_t,
                                                              long long ld_c, const uint8_t* __restrict__ mask,
                                                              const __nv_bfloat16* __restrict__ table_a, long long rows_a,
                                                              long long col_offset, int col0, int n_cols_a,
                                                              const __nv_bfloat16* __restrict__ table_b, long long rows_b,
                                                              int C, int dim, int round_each) {
  pdl_sync();
  extern __shared__ __align__(16) long long s_row[];   // [C (+1: the compiler reads pairs)] source row of every column, -1 = masked off
  const size_t t = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    long long r = -1;
    if (!mask || mask[t * C + c]) {
      const long long id = ids[t * ld_t + c * ld_c];
      if (c < n_cols_a) {
        r = id + static_cast<long long>(col0 + c) * col_offset;
        r = r < 0 ? 0 : (r >= rows_a ? rows_a - 1 : r);
      } else {
        r = id < 0 ? 0 : (id >= rows_b ? rows_b - 1 : id);
        r = -2 - r;                                // table_b rows are encoded as <= -2
      }
    }
    s_row[c] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long r = s_row[c];
      if (r == -1) continue;
      const __nv_bfloat16* src = r >= 0 ? table_a + static_cast<size_t>(r) * dim : table_b + static_cast<size_t>(-2 - r) * dim;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += bf16_lo(u[j]);
        acc[2 * j + 1] += bf16_hi(u[j]);
      }
      if (round_each) {        // a chain of bf16 `+=` (qwen3_tts.py:2002) rounds after every term
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = round_bf16(acc[j]);
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out + t * ld_out)[i] = o;
  }
}

// Qwen3-TTS talker input row (vox_serve/model/qwen3_tts.py:1835-1853):
//   out[t] = bf16( (needs_codec[t] ? bf16(text[t] + codec[cb0[t]]) : text[t]) + features[t] )
// text rows come with stride ld_text (0 = one row broadcast: decode rows all carry text_projection(embed(tts_pad)));
// cb0 int64 with element stride ld_id; needs_codec NULL = all rows; features NULL = zeros.
__global__ void __launch_bounds__(256) talker_embed_kernel(__nv_bfloat16* __restrict__ out, int ld_out,
                                                           const __nv_bfloat16* __restrict__ text, long long ld_text,
                                                           const __nv_bfloat16* __restrict__ codec, long long codec_rows,
                                                           const long long* __restrict__ cb0, long long ld_id,
                                                           const uint8_t* __restrict__ needs_codec,
                                                           const __nv_bfloat16* __restrict__ features, long long ld_feat,
                                                           int dim) {
  pdl_sync();
  const size_t t = blockIdx.x;
  const bool use_c = !needs_codec || needs_codec[t];
  long long id = use_c ? cb0[t * ld_id] : 0;
  id = id < 0 ? 0 : (id >= codec_rows ? codec_rows - 1 : id);
  const __nv_bfloat16* tx = text + t * ld_text;
  const __nv_bfloat16* cx = codec + static_cast<size_t>(id) * dim;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = __bfloat162float(tx[i]);
    if (use_c) v = round_bf16(v + __bfloat162float(cx[i]));
    if (features) v = round_bf16(v + __bfloat162float(features[t * ld_feat + i]));
    out[t * ld_out + i] = __float2bfloat16_rn(v);
  }
}

// rows of two [n, dim] matrices interleaved: out[2 r] = a[r], out[2 r + 1] = b[r]  -- the depth decoder's 2-row prefill
// input [backbone hidden state, embed(codebook 0)] per request (csm.py:700-701: torch.cat along a new dim)
__global__ void __launch_bounds__(256) interleave_rows_kernel(uint4* __restrict__ out, const uint4* __restrict__ a,
                                                              const uint4* __restrict__ b, int vec_per_row) {
  pdl_sync();
  const size_t r = blockIdx.x;
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) {
    out[(2 * r) * vec_per_row + i] = a[r * vec_per_row + i];
    out[(2 * r + 1) * vec_per_row + i] = b[r * vec_per_row + i];
  }
}

// dst[r * ld_dst + c] = src[c * ld_src + r]: the frame buffer kept codebook-major on the device ([C][B], one contiguous
// row per sampler call) -> the request-major [B][C] ids the host appends to lm_output_tokens
__global__ void transpose_i64_kernel(long long* __restrict__ dst, const long long* __restrict__ src, int B, int C,
                                     int ld_dst, int ld_src) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int r = i / C, c = i - r * C;
  dst[static_cast<size_t>(r) * ld_dst + c] = src[static_cast<size_t>(c) * ld_src + r];
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_multi_embed_sum(void* d_out, int ld_out, const int64_t* d_ids, int64_t ld_t, int64_t ld_c, const uint8_t* d_mask,
                       const void* d_table_a, int64_t rows_a, int64_t col_offset, int col0, int n_cols_a,
                       const void* d_table_b, int64_t rows_b, int T, int C, int dim, int round_each, void* stream) {
  VB_CHECK_ARG(d_out && d_ids && (d_table_a || n_cols_a == 0) && (d_table_b || n_cols_a >= C),
               "vb_multi_embed_sum: null pointer");
  VB_CHECK_ARG(dim > 0 && dim % 8 == 0 && ld_out >= dim && C > 0 && C <= 1024 && n_cols_a >= 0 && n_cols_a <= C,
               "vb_multi_embed_sum: bad shape (dim %d, C %d, n_cols_a %d)", dim, C, n_cols_a);
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(multi_embed_sum_kernel, T, 256, static_cast<size_t>((C + 2) & ~1) * sizeof(long long), stream,
                static_cast<__nv_bfloat16*>(d_out), ld_out, reinterpret_cast<const long long*>(d_ids),
                static_cast<long long>(ld_t), static_cast<long long>(ld_c), d_mask,
                static_cast<const __nv_bfloat16*>(d_table_a), static_cast<long long>(rows_a),
                static_cast<long long>(col_offset), col0, n_cols_a, static_cast<const __nv_bfloat16*>(d_table_b),
                static_cast<long long>(rows_b), C, dim, round_each);
  return 0;
}

int vb_talker_embed(void* d_out, int ld_out, const void* d_text, int64_t ld_text, const void* d_codec, int64_t codec_rows,
                    const int64_t* d_cb0, int64_t ld_id, const uint8_t* d_needs_codec, const void* d_features,
                    int64_t ld_feat, int T, int dim, void* stream) {
  VB_CHECK_ARG(d_out && d_text && d_codec && d_cb0 && dim > 0 && ld_out >= dim, "vb_talker_embed: bad arguments");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(talker_embed_kernel, T, 256, 0, stream, static_cast<__nv_bfloat16*>(d_out), ld_out,
                static_cast<const __nv_bfloat16*>(d_text), static_cast<long long>(ld_text),
                static_cast<const __nv_bfloat16*>(d_codec), static_cast<long long>(codec_rows),
                reinterpret_cast<const long long*>(d_cb0), static_cast<long long>(ld_id), d_needs_codec,
                static_cast<const __nv_bfloat16*>(d_features), static_cast<long long>(ld_feat), dim);
  return 0;
}

int vb_interleave_rows(void* d_out, const void* d_a, const void* d_b, int n, int row_bytes, void* stream) {
  VB_CHECK_ARG(d_out && d_a && d_b && row_bytes % 16 == 0, "vb_interleave_rows: bad arguments");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(interleave_rows_kernel, n, 256, 0, stream, static_cast<uint4*>(d_out), static_cast<const uint4*>(d_a),
                static_cast<const uint4*>(d_b), row_bytes / 16);
  return 0;
}

int vb_transpose_i64(int64_t* d_dst, const int64_t* d_src, int B, int C, int ld_dst, int ld_src, void* stream) {
  VB_CHECK_ARG(d_dst && d_src && ld_dst >= C && ld_src >= B, "vb_transpose_i64: bad arguments");
  if (B <= 0 || C <= 0) return 0;
  VB_LAUNCH_PDL(transpose_i64_kernel, (B * C + 255) / 256, 256, 0, stream, reinterpret_cast<long long*>(d_dst),
                reinterpret_cast<const long long*>(d_src), B, C, ld_dst, ld_src);
  return 0;
}

}  // extern "C"
