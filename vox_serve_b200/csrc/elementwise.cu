// Row-wise / bookkeeping kernels of the decode step: RMSNorm, RoPE, page-table plan, KV append,
// split-K reduction fused with residual + RMSNorm, fused QKV tail (reduce -> RoPE -> cache scatter),
// embedding / row gathers, PCM16, Orpheus window de-interleave.  All HBM/latency-bound CUDA-core work.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 4
#include "common.cuh"

namespace vb {

// ------------------------------------------------------------------------------------------
// RMSNorm (flashinfer_utils.py:251-267; arithmetic order of norm.cuh:64-101: (x * rcp) * w)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarp) ? red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) rmsnorm_kernel(__nv_bfloat16* __restrict__ out,
                                                      const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ w, int dim, float eps,
                                                      int xt_tile) {
  pdl_sync();
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * dim;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
    if (i + 8 <= dim) {
      uint4 v = *reinterpret_cast<const uint4*>(xr + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
        ss += a * a + b * b;
      }
    } else {
      for (int j = i; j < dim; ++j) {
        float a = __bfloat162float(xr[j]);
        ss += a * a;
      }
    }
  }
  ss = block_sum(ss, red);
  const float rcp = rsqrtf(ss / static_cast<float>(dim) + eps);
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = __bfloat162float(xr[i]) * rcp * __bfloat162float(w[i]);
    const size_t oi = xt_tile ? xt_index(static_cast<int>(row), i, xt_tile, (dim + 63) >> 6) : row * dim + i;
    out[oi] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// RoPE (flashinfer_utils.py:270-324; pos_enc.cuh:129-147, 594-617)
// freq[e], e < rotary_dim, is the per-element frequency (host: vox_serve_b200.flashinfer_utils).
// ------------------------------------------------------------------------------------------
__global__ void rope_freq_kernel(float* freq, int rotary_dim, int interleave, float rope_rcp_scale,
                                 float rope_rcp_theta, float smooth_a, float smooth_b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rotary_dim) return;
  float e = interleave ? float(2 * (i / 2)) / float(rotary_dim) : float(2 * (i % (rotary_dim / 2))) / float(rotary_dim);
  float f = powf(rope_rcp_theta, e);
  float smooth = fminf(1.f, fmaxf(0.f, f * smooth_a + smooth_b));
  freq[i] = (1.f - smooth) * (f * rope_rcp_scale) + smooth * f;
}

__device__ __forceinline__ int rope_partner(int e, int rotary_dim, int interleave, float& sign) {
  if (interleave) {
    sign = (e & 1) ? 1.f : -1.f;
    return e ^ 1;
  }
  const int half = rotary_dim >> 1;
  sign = (e < half) ? -1.f : 1.f;
  return (e < half) ? e + half : e - half;
}

// one block per token; inputs staged in shared memory first so q_out / k_out may alias q / k
__global__ void __launch_bounds__(256) rope_kernel(__nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out,
                                                   const __nv_bfloat16* q, const __nv_bfloat16* k,
                                                   const int32_t* __restrict__ pos, const float* __restrict__ freq,
                                                   int n_q, int n_kv, int D, int rotary_dim, int interleave) {
  pdl_sync();
  extern __shared__ float sm[];  // cos[rotary_dim], sin[rotary_dim], values[(n_q+n_kv)*D]
  float* cs = sm;
  float* val = sm + 2 * rotary_dim;
  const size_t t = blockIdx.x;
  const float p = static_cast<float>(pos[t]);
  for (int e = threadIdx.x; e < rotary_dim; e += blockDim.x) {
    float s, c;
    sincosf(p * freq[e], &s, &c);
    cs[e] = c;
    cs[rotary_dim + e] = s;
  }
  const int nq_el = n_q * D, total = (n_q + n_kv) * D;
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    val[i] = __bfloat162float(i < nq_el ? q[t * nq_el + i] : k[t * (total - nq_el) + (i - nq_el)]);
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int h = i / D, e = i - h * D;
    float v = val[i];
    if (e < rotary_dim) {
      float sign;
      const int pe = rope_partner(e, rotary_dim, interleave, sign);
      v = v * cs[e] + sign * val[h * D + pe] * cs[rotary_dim + e];
    }
    if (i < nq_el) q_out[t * nq_el + i] = __float2bfloat16_rn(v);
    else k_out[t * (total - nq_el) + (i - nq_el)] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// plan: page table -> per-row attention / append metadata  (flashinfer_utils.py:86-124, 217-225)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) plan_rows_kernel(const int32_t* __restrict__ qo_indptr,
                                                         const int32_t* __restrict__ kv_indptr,
                                                         const int32_t* __restrict__ kv_indices,
                                                         const int32_t* __restrict__ last_page_len,
                                                         const int32_t* __restrict__ kv_len_in, int n_req,
                                                         int n_rows_padded, int page_size, int chunk,
                                                         int32_t* __restrict__ row_req, int32_t* __restrict__ row_kvlen,
                                                         int32_t* __restrict__ row_page, int32_t* __restrict__ row_slot,
                                                         int32_t* __restrict__ row_chunk_start,
                                                         int32_t* __restrict__ row_pagebase,
                                                         int32_t* __restrict__ row_old) {
  pdl_sync();
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  const int tid = threadIdx.x;
  // pass 1: per request, fill its rows
  for (int r = tid; r < n_req; r += blockDim.x) {
    const int row0 = qo_indptr ? qo_indptr[r] : r;
    const int n_new = qo_indptr ? qo_indptr[r + 1] - row0 : 1;
    const int p0 = kv_indptr[r];
    const int kv_len = kv_len_in ? kv_len_in[r] : (kv_indptr[r + 1] - p0 - 1) * page_size + last_page_len[r];
    for (int j = 0; j < n_new; ++j) {
      const int row = row0 + j;
      if (row >= n_rows_padded) break;
      const int g = kv_len - n_new + j;
      row_req[row] = r;
      row_pagebase[row] = p0;
      row_old[row] = kv_len - n_new;
      row_kvlen[row] = g + 1;
      row_page[row] = kv_indices[p0 + g / page_size];
      row_slot[row] = g % page_size;
    }
  }
  const int n_valid = qo_indptr ? qo_indptr[n_req] : n_req;
  for (int row = n_valid + tid; row < n_rows_padded; row += blockDim.x) {
    row_req[row] = -1;
    row_pagebase[row] = 0;
    row_old[row] = 0;
    row_kvlen[row] = 0;
    row_page[row] = -1;
    row_slot[row] = 0;
  }
  if (tid == 0) carry = 0;
  __syncthreads();
  // pass 2: exclusive scan of ceil(kvlen / chunk) over rows, 1024 rows per sweep
  for (int base = 0; base < n_rows_padded; base += blockDim.x) {
    const int row = base + tid;
    const int c = (row < n_rows_padded) ? (row_kvlen[row] + chunk - 1) / chunk : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += n;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      int v = warp_tot[tid], w = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, w, o);
        if (tid >= o) w += n;
      }
      warp_tot[tid] = w - v;  // exclusive
    }
    __syncthreads();
    const int excl = carry + warp_tot[tid >> 5] + incl - c;
    if (row < n_rows_padded) row_chunk_start[row] = excl;
    __syncthreads();
    if (tid == blockDim.x - 1) carry = excl + c;
    __syncthreads();
  }
  if (tid == 0) row_chunk_start[n_rows_padded] = carry;
}

// ------------------------------------------------------------------------------------------
// KV append (flashinfer_utils.py:144-145, 243-244)
// ------------------------------------------------------------------------------------------
__global__ void kv_append_kernel(__nv_bfloat16* __restrict__ kv, const __nv_bfloat16* __restrict__ k,
                                 const __nv_bfloat16* __restrict__ v, const int32_t* __restrict__ row_page,
                                 const int32_t* __restrict__ row_slot, int page_size, int row_elems /*n_kv*D*/) {
  pdl_sync();
  const size_t t = blockIdx.x;
  const int page = row_page[t];
  if (page < 0) return;
  const size_t slab = static_cast<size_t>(page_size) * row_elems;
  __nv_bfloat16* kd = kv + (static_cast<size_t>(page) * 2) * slab + static_cast<size_t>(row_slot[t]) * row_elems;
  __nv_bfloat16* vd = kd + slab;
  const uint4* ks = reinterpret_cast<const uint4*>(k + t * row_elems);
  const uint4* vs = reinterpret_cast<const uint4*>(v + t * row_elems);
  for (int i = threadIdx.x; i < row_elems / 8; i += blockDim.x) {
    reinterpret_cast<uint4*>(kd)[i] = ks[i];
    reinterpret_cast<uint4*>(vd)[i] = vs[i];
  }
}

// ------------------------------------------------------------------------------------------
// split-K reduce + residual + RMSNorm (orpheus.py:125-151 rounding points)
// ------------------------------------------------------------------------------------------
// One row is split over the CTAs of a thread-block cluster (blockIdx.x = part): every CTA reduces N / parts
// columns with all partial loads in flight at once, the row's sum of squares is exchanged through distributed
// shared memory, and each CTA normalises its own columns.  (One CTA per row left 116 SMs idle and serialised
// ~70 KB of dependent loads per CTA: 10 us for 400 KB of traffic.)
constexpr int RR_THREADS = 256;
constexpr int RR_MAX_ITER = 4;      // columns per CTA <= RR_THREADS * 4 * RR_MAX_ITER
__global__ void __launch_bounds__(RR_THREADS) reduce_residual_rmsnorm_kernel(
    __nv_bfloat16* __restrict__ hidden_out, __nv_bfloat16* __restrict__ normed_out,
    const float* __restrict__ partials, int split_k, const __nv_bfloat16* __restrict__ residual,
    const __nv_bfloat16* __restrict__ norm_w, int T, int N, float eps, int xt_tile) {
  __shared__ float red[32];
  __shared__ float part_ss[8];        // one slot per CTA of the cluster, written by the peers
  const int parts = gridDim.x, part = blockIdx.x;
  const size_t t = blockIdx.y;
  const int cols = N / parts, c0 = part * cols;
  const size_t plane = static_cast<size_t>(T) * N;
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(5, split_k) : -1;
  // the norm weights are parameters: fetched before the dependency resolves
  uint2 w2[RR_MAX_ITER];
#pragma unroll
  for (int it = 0; it < RR_MAX_ITER; ++it) {
    const int n = c0 + (it * RR_THREADS + threadIdx.x) * 4;
    w2[it] = (normed_out && n < c0 + cols) ? __ldg(reinterpret_cast<const uint2*>(norm_w + n)) : make_uint2(0u, 0u);
  }
  pdl_sync();
  float h[RR_MAX_ITER][4];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < RR_MAX_ITER; ++it) {
    const int n = c0 + (it * RR_THREADS + threadIdx.x) * 4;
    if (n < c0 + cols) {
      const float* src = partials + t * N + n;
      float4 acc = *reinterpret_cast<const float4*>(src);
#pragma unroll 8
      for (int s = 1; s < split_k; ++s) {
        const float4 p4 = *reinterpret_cast<const float4*>(src + s * plane);
        acc.x += p4.x; acc.y += p4.y; acc.z += p4.z; acc.w += p4.w;
      }
      h[it][0] = round_bf16(acc.x); h[it][1] = round_bf16(acc.y);
      h[it][2] = round_bf16(acc.z); h[it][3] = round_bf16(acc.w);
      if (residual) {
        const uint2 r = *reinterpret_cast<const uint2*>(residual + t * N + n);
        h[it][0] = round_bf16(bf16_lo(r.x) + h[it][0]);
        h[it][1] = round_bf16(bf16_hi(r.x) + h[it][1]);
        h[it][2] = round_bf16(bf16_lo(r.y) + h[it][2]);
        h[it][3] = round_bf16(bf16_hi(r.y) + h[it][3]);
      }
      if (hidden_out) {
        uint2 o;
        o.x = pack_bf16(h[it][0], h[it][1]);
        o.y = pack_bf16(h[it][2], h[it][3]);
        *reinterpret_cast<uint2*>(hidden_out + t * N + n) = o;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ss += h[it][j] * h[it][j];
    }
  }
  if (!normed_out) { trace_end(tr); return; }      // uniform across the cluster: no barrier is pending
  ss = block_sum(ss, red);
  float total = ss;
  if (parts > 1) {
    // push this CTA's sum into slot `part` of every peer, one barrier, then every CTA adds its own eight slots in
    // the same order: nobody reads remote memory after the barrier, so nobody has to wait before exiting
    if (threadIdx.x < parts) {
      uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(&part_ss[part])), raddr;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(threadIdx.x));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(ss) : "memory");
    }
    cluster_sync_all();
    total = 0.f;
    for (int r = 0; r < parts; ++r) total += part_ss[r];
  }
  const float rcp = rsqrtf(total / static_cast<float>(N) + eps);
#pragma unroll
  for (int it = 0; it < RR_MAX_ITER; ++it) {
    const int n = c0 + (it * RR_THREADS + threadIdx.x) * 4;
    if (n < c0 + cols) {
      uint2 o;
      o.x = pack_bf16(h[it][0] * rcp * bf16_lo(w2[it].x), h[it][1] * rcp * bf16_hi(w2[it].x));
      o.y = pack_bf16(h[it][2] * rcp * bf16_lo(w2[it].y), h[it][3] * rcp * bf16_hi(w2[it].y));
      const size_t oi = xt_tile ? xt_index(static_cast<int>(t), n, xt_tile, (N + 63) >> 6) : t * N + n;
      *reinterpret_cast<uint2*>(normed_out + oi) = o;
    }
  }
  trace_end(tr);
}

// ------------------------------------------------------------------------------------------
// fused QKV tail: reduce partials -> bf16, RoPE(q, k), q out, K/V scatter into the layer cache.
// Heads are independent, so a row is split over gridDim.x CTAs by groups of heads (blockIdx.x).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qkv_rope_append_kernel(
    __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ kv, const float* __restrict__ partials,
    int split_k, const int32_t* __restrict__ pos, const float* __restrict__ freq,
    const int32_t* __restrict__ row_page, const int32_t* __restrict__ row_slot, int T, int n_q, int n_kv, int D,
    int page_size, int rotary_dim, int interleave, int heads_per_cta, const __nv_bfloat16* __restrict__ q_norm_w,
    const __nv_bfloat16* __restrict__ k_norm_w, float norm_eps, const __nv_bfloat16* __restrict__ qkv_bias) {
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(6, split_k) : -1;
  pdl_sync();
  if (tr >= 0) trace_mark(24);
  extern __shared__ float sm[];  // [2*rotary_dim cos/sin][heads_per_cta * D values]
  float* cs = sm;
  float* val = sm + 2 * rotary_dim;
  const size_t t = blockIdx.y;
  const int n_heads = n_q + 2 * n_kv;
  const int h_lo = blockIdx.x * heads_per_cta, h_hi = min(n_heads, h_lo + heads_per_cta);
  if (h_lo >= h_hi) { trace_end(tr); return; }
  const int W = n_heads * D, w_lo = h_lo * D, w_n = (h_hi - h_lo) * D;
  const size_t plane = static_cast<size_t>(T) * W;
  const float p = static_cast<float>(pos[t]);
  if (h_lo < n_q + n_kv) {       // groups made only of V heads need no rotation table
    for (int e = threadIdx.x; e < rotary_dim; e += blockDim.x) {
      float s, c;
      sincosf(p * freq[e], &s, &c);
      cs[e] = c;
      cs[rotary_dim + e] = s;
    }
  }
  for (int n = threadIdx.x * 4; n < w_n; n += blockDim.x * 4) {
    const float* src = partials + t * W + w_lo + n;
    float4 acc = *reinterpret_cast<const float4*>(src);
#pragma unroll 4
    for (int s = 1; s < split_k; ++s) {
      const float4 q4 = *reinterpret_cast<const float4*>(src + s * plane);
      acc.x += q4.x; acc.y += q4.y; acc.z += q4.z; acc.w += q4.w;
    }
    if (qkv_bias) {      // nn.Linear with bias: added in fp32 before the single bf16 rounding (cosyvoice2.py:139-143)
      const uint2 b2 = __ldg(reinterpret_cast<const uint2*>(qkv_bias + w_lo + n));
      acc.x += bf16_lo(b2.x); acc.y += bf16_hi(b2.x); acc.z += bf16_lo(b2.y); acc.w += bf16_hi(b2.y);
    }
    val[n] = round_bf16(acc.x);
    val[n + 1] = round_bf16(acc.y);
    val[n + 2] = round_bf16(acc.z);
    val[n + 3] = round_bf16(acc.w);
  }
  __syncthreads();
  if (q_norm_w) {
    // per-head RMSNorm of every q and k head before the rotation (Qwen3: q_norm / k_norm over head_dim,
    // vox_serve/model/qwen3_tts.py:603-604, 620-625): fp32 statistics over the bf16-rounded projection output, result
    // rounded to bf16 like flashinfer.norm.rmsnorm; one warp per head
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int hl = warp; hl < h_hi - h_lo; hl += n_warps) {
      const int h = h_lo + hl;
      if (h >= n_q + n_kv) continue;                      // V heads are not normalised
      const __nv_bfloat16* nw = h < n_q ? q_norm_w : k_norm_w;
      float ss = 0.f;
      for (int e = lane; e < D; e += 32) ss += val[hl * D + e] * val[hl * D + e];
      ss = warp_sum(ss);
      const float rcp = rsqrtf(ss / static_cast<float>(D) + norm_eps);
      for (int e = lane; e < D; e += 32) val[hl * D + e] = round_bf16((val[hl * D + e] * rcp) * __bfloat162float(nw[e]));
    }
    __syncthreads();
  }
  const int page = row_page[t];
  const size_t row_elems = static_cast<size_t>(n_kv) * D;
  const size_t slab = static_cast<size_t>(page_size) * row_elems;
  __nv_bfloat16* kd = (page >= 0) ? kv + (static_cast<size_t>(page) * 2) * slab + static_cast<size_t>(row_slot[t]) * row_elems
                                   : nullptr;
  for (int il = threadIdx.x * 2; il < w_n; il += blockDim.x * 2) {
    const int hl = il / D, e = il - hl * D;
    const int h = h_lo + hl, i = w_lo + il;
    float v0 = val[il], v1 = val[il + 1];
    if (h < n_q + n_kv && e < rotary_dim) {
      float s0, s1;
      const int p0 = rope_partner(e, rotary_dim, interleave, s0);
      const int p1 = rope_partner(e + 1, rotary_dim, interleave, s1);
      v0 = v0 * cs[e] + s0 * val[hl * D + p0] * cs[rotary_dim + e];
      v1 = v1 * cs[e + 1] + s1 * val[hl * D + p1] * cs[rotary_dim + e + 1];
    }
    const uint32_t packed = pack_bf16(v0, v1);
    if (h < n_q) {
      *reinterpret_cast<uint32_t*>(q_out + t * static_cast<size_t>(n_q) * D + i) = packed;
    } else if (kd) {
      const int j = i - n_q * D;  // offset inside [k heads | v heads]
      if (j < static_cast<int>(row_elems))
        *reinterpret_cast<uint32_t*>(kd + j) = packed;
      else
        *reinterpret_cast<uint32_t*>(kd + slab + (j - row_elems)) = packed;
    }
  }
  trace_end(tr);
}

__global__ void embedding_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ table,
                                 const int32_t* __restrict__ ids, int dim, int vocab) {
  pdl_sync();
  const size_t t = blockIdx.x;
  int id = ids[t];
  id = min(max(id, 0), vocab - 1);
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * dim);
  uint4* dst = reinterpret_cast<uint4*>(out + t * dim);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) dst[i] = src[i];
}

__global__ void gather_rows_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                   const int32_t* __restrict__ idx, int row_bytes, int idx_offset) {
  pdl_sync();
  const size_t i = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(in + static_cast<size_t>(idx[i] + idx_offset) * row_bytes);
  uint4* dst = reinterpret_cast<uint4*>(out + i * row_bytes);
  for (int j = threadIdx.x; j < row_bytes / 16; j += blockDim.x) dst[j] = src[j];
}

__global__ void pcm16_kernel(int16_t* __restrict__ out, const float* __restrict__ a, int64_t n) {
  pdl_sync();
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<int16_t>(__float2int_rz(a[i] * 32767.0f));
}

__global__ void orpheus_window_codes_kernel(int32_t* __restrict__ c0, int32_t* __restrict__ c1,
                                            int32_t* __restrict__ c2, const int64_t* __restrict__ ids, int B,
                                            int base) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over B*28
  if (i >= B * 28) return;
  const int b = i / 28, r = i % 28, f = r / 7, p = r % 7;
  long long v = (ids[i] - base) % 4096;
  if (v < 0) v += 4096;  // python modulo
  const int code = static_cast<int>(v);
  if (p == 0) c0[b * 4 + f] = code;
  else if (p == 1) c1[b * 8 + f * 2 + 0] = code;
  else if (p == 4) c1[b * 8 + f * 2 + 1] = code;
  else if (p == 2) c2[b * 16 + f * 4 + 0] = code;
  else if (p == 3) c2[b * 16 + f * 4 + 1] = code;
  else if (p == 5) c2[b * 16 + f * 4 + 2] = code;
  else c2[b * 16 + f * 4 + 3] = code;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_rmsnorm(void* d_out, const void* d_x, const void* d_weight, int rows, int dim, float eps, int xt_tile,
               void* stream) {
  VB_CHECK_ARG(d_out && d_x && d_weight, "vb_rmsnorm: null pointer");
  VB_CHECK_ARG(dim > 0 && dim % 8 == 0, "vb_rmsnorm: dim %d must be a positive multiple of 8", dim);
  if (rows <= 0) return 0;
  VB_LAUNCH_PDL(rmsnorm_kernel, rows, 256, 0, stream, static_cast<__nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(d_x), static_cast<const __nv_bfloat16*>(d_weight), dim, eps, xt_tile);
  return 0;
}

int vb_rope_freqs(float* d_freq, int rotary_dim, int interleave, float rope_scale, float rope_theta, int llama31,
                  float low_freq_factor, float high_freq_factor, float old_context_len, void* stream) {
  VB_CHECK_ARG(d_freq && rotary_dim > 0 && rotary_dim % 2 == 0, "vb_rope_freqs: bad arguments");
  float a = 0.f, b = 0.f;
  if (llama31) {
    a = old_context_len / (2.f * 3.14159265358979323846f * high_freq_factor -
                           2.f * 3.14159265358979323846f * low_freq_factor);
    b = -1.0f / (high_freq_factor / low_freq_factor - 1.0f);
  }
  VB_LAUNCH_PLAIN(rope_freq_kernel, (rotary_dim + 127) / 128, 128, 0, stream, d_freq, rotary_dim, interleave, 1.0f / rope_scale, 1.0f / rope_theta, a, b);
  return 0;
}

int vb_rope(void* d_q_out, void* d_k_out, const void* d_q, const void* d_k, const int32_t* d_pos,
            const float* d_freq, int T, int n_q, int n_kv, int head_dim, int rotary_dim, int interleave,
            void* stream) {
  VB_CHECK_ARG(d_q_out && d_k_out && d_q && d_k && d_pos && d_freq, "vb_rope: null pointer");
  VB_CHECK_ARG(rotary_dim > 0 && rotary_dim <= head_dim && rotary_dim % 2 == 0, "vb_rope: rotary_dim %d invalid",
               rotary_dim);
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(rope_kernel, T, 256, (2 * rotary_dim + (n_q + n_kv) * head_dim) * sizeof(float), stream, static_cast<__nv_bfloat16*>(d_q_out), static_cast<__nv_bfloat16*>(d_k_out), static_cast<const __nv_bfloat16*>(d_q), static_cast<const __nv_bfloat16*>(d_k), d_pos, d_freq, n_q, n_kv, head_dim, rotary_dim, interleave);
  return 0;
}

int vb_plan_rows(const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                 const int32_t* d_last_page_len, const int32_t* d_kv_len, int n_req, int n_rows_padded, int page_size,
                 int chunk_tokens,
                 int32_t* d_row_req, int32_t* d_row_kvlen, int32_t* d_row_page, int32_t* d_row_slot,
                 int32_t* d_row_chunk_start, int32_t* d_row_pagebase, int32_t* d_row_old, void* stream) {
  VB_CHECK_ARG(d_kv_indptr && d_kv_indices && (d_last_page_len || d_kv_len) && d_row_req && d_row_kvlen &&
                   d_row_page && d_row_slot && d_row_chunk_start && d_row_pagebase && d_row_old,
               "vb_plan_rows: null pointer");
  VB_CHECK_ARG(page_size > 0 && chunk_tokens > 0 && page_size % chunk_tokens == 0,
               "vb_plan_rows: chunk_tokens %d must divide page_size %d", chunk_tokens, page_size);
  VB_CHECK_ARG(n_req >= 0 && n_rows_padded >= 0, "vb_plan_rows: negative sizes");
  VB_LAUNCH_PDL(plan_rows_kernel, 1, 1024, 0, stream, d_qo_indptr, d_kv_indptr, d_kv_indices, d_last_page_len, d_kv_len, n_req, n_rows_padded, page_size, chunk_tokens, d_row_req, d_row_kvlen, d_row_page, d_row_slot, d_row_chunk_start, d_row_pagebase, d_row_old);
  return 0;
}

int vb_kv_append(void* d_layer_kv, const void* d_k, const void* d_v, const int32_t* d_row_page,
                 const int32_t* d_row_slot, int T, int page_size, int n_kv, int head_dim, void* stream) {
  VB_CHECK_ARG(d_layer_kv && d_k && d_v && d_row_page && d_row_slot, "vb_kv_append: null pointer");
  VB_CHECK_ARG((n_kv * head_dim) % 8 == 0, "vb_kv_append: n_kv*head_dim must be a multiple of 8");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(kv_append_kernel, T, 128, 0, stream, static_cast<__nv_bfloat16*>(d_layer_kv), static_cast<const __nv_bfloat16*>(d_k), static_cast<const __nv_bfloat16*>(d_v), d_row_page, d_row_slot, page_size, n_kv * head_dim);
  return 0;
}

int vb_reduce_residual_rmsnorm(void* d_hidden_out, void* d_normed_out, const float* d_partials, int split_k,
                               const void* d_residual, const void* d_norm_weight, int T, int N, float eps,
                               int normed_xt_tile, void* stream) {
  VB_CHECK_ARG(d_partials && split_k >= 1, "vb_reduce_residual_rmsnorm: bad partials");
  VB_CHECK_ARG(N % 4 == 0, "vb_reduce_residual_rmsnorm: N %d must be a multiple of 4", N);
  VB_CHECK_ARG(!d_normed_out || d_norm_weight, "vb_reduce_residual_rmsnorm: norm output needs a weight");
  if (T <= 0) return 0;
  // widest cluster (<= 8, the portable limit) that divides the row into float4-aligned parts; few rows -> more parts
  int parts = 1;
  const int want = T <= 64 ? 8 : (T <= 256 ? 4 : 1);
  for (int c = want; c >= 1; c >>= 1)
    if (N % (4 * c) == 0) { parts = c; break; }
  while (N / parts > RR_THREADS * 4 * RR_MAX_ITER) {
    VB_CHECK_ARG(parts < 8 && N % (8 * parts) == 0, "vb_reduce_residual_rmsnorm: N %d too wide", N);
    parts *= 2;
  }
  VB_LAUNCH_PDL_CLUSTER(reduce_residual_rmsnorm_kernel, dim3(parts, T), RR_THREADS, 0, stream, parts,
                        static_cast<__nv_bfloat16*>(d_hidden_out), static_cast<__nv_bfloat16*>(d_normed_out),
                        d_partials, split_k, static_cast<const __nv_bfloat16*>(d_residual),
                        static_cast<const __nv_bfloat16*>(d_norm_weight), T, N, eps, normed_xt_tile);
  return 0;
}

int vb_qkv_rope_append(void* d_q_out, void* d_layer_kv, const float* d_partials, int split_k, const int32_t* d_pos,
                       const float* d_freq, const int32_t* d_row_page, const int32_t* d_row_slot, int T, int n_q,
                       int n_kv, int head_dim, int page_size, int rotary_dim, int interleave, const void* d_q_norm,
                       const void* d_k_norm, float norm_eps, const void* d_qkv_bias, void* stream) {
  VB_CHECK_ARG((d_q_norm == nullptr) == (d_k_norm == nullptr), "vb_qkv_rope_append: q_norm and k_norm come together");
  VB_CHECK_ARG(d_q_out && d_layer_kv && d_partials && d_pos && d_freq && d_row_page && d_row_slot,
               "vb_qkv_rope_append: null pointer");
  VB_CHECK_ARG(head_dim % 4 == 0 && rotary_dim % 2 == 0 && rotary_dim <= head_dim, "vb_qkv_rope_append: bad dims");
  if (T <= 0) return 0;
  const int n_heads = n_q + 2 * n_kv;
  // few rows (decode): spread a row's heads over several CTAs; many rows (prefill): one CTA per row is enough
  int heads_per_cta = T <= 64 ? 4 : (T <= 256 ? 8 : n_heads);
  while (static_cast<size_t>(2 * rotary_dim + heads_per_cta * head_dim) * sizeof(float) > 96 * 1024 && heads_per_cta > 1)
    heads_per_cta /= 2;
  const int parts = (n_heads + heads_per_cta - 1) / heads_per_cta;
  const size_t smem = (2 * static_cast<size_t>(rotary_dim) + static_cast<size_t>(heads_per_cta) * head_dim) *
                      sizeof(float);
  if (smem > 48 * 1024)
    VB_CHECK_CUDA(cudaFuncSetAttribute(qkv_rope_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       VB_MAX_DYN_SMEM));
  VB_LAUNCH_PDL(qkv_rope_append_kernel, dim3(parts, T), 256, smem, stream, static_cast<__nv_bfloat16*>(d_q_out),
                static_cast<__nv_bfloat16*>(d_layer_kv), d_partials, split_k, d_pos, d_freq, d_row_page, d_row_slot, T,
                n_q, n_kv, head_dim, page_size, rotary_dim, interleave, heads_per_cta,
                static_cast<const __nv_bfloat16*>(d_q_norm), static_cast<const __nv_bfloat16*>(d_k_norm), norm_eps,
                static_cast<const __nv_bfloat16*>(d_qkv_bias));
  return 0;
}

int vb_embedding(void* d_out, const void* d_table, const int32_t* d_ids, int T, int dim, int vocab, void* stream) {
  VB_CHECK_ARG(d_out && d_table && d_ids && dim % 8 == 0, "vb_embedding: bad arguments");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(embedding_kernel, T, 128, 0, stream, static_cast<__nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(d_table), d_ids, dim, vocab);
  return 0;
}

int vb_gather_rows(void* d_out, const void* d_in, const int32_t* d_idx, int n, int row_bytes, int idx_offset,
                   void* stream) {
  VB_CHECK_ARG(d_out && d_in && d_idx && row_bytes % 16 == 0, "vb_gather_rows: bad arguments");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(gather_rows_kernel, n, 256, 0, stream, static_cast<uint8_t*>(d_out), static_cast<const uint8_t*>(d_in), d_idx, row_bytes, idx_offset);
  return 0;
}

int vb_pcm16(int16_t* d_out, const float* d_audio, int64_t n, void* stream) {
  VB_CHECK_ARG(d_out && d_audio, "vb_pcm16: null pointer");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(pcm16_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, stream, d_out, d_audio, n);
  return 0;
}

int vb_orpheus_window_codes(int32_t* d_c0, int32_t* d_c1, int32_t* d_c2, const int64_t* d_ids, int B,
                            int audio_id_base, void* stream) {
  VB_CHECK_ARG(d_c0 && d_c1 && d_c2 && d_ids, "vb_orpheus_window_codes: null pointer");
  if (B <= 0) return 0;
  VB_LAUNCH_PDL(orpheus_window_codes_kernel, (B * 28 + 127) / 128, 128, 0, stream, d_c0, d_c1, d_c2, d_ids, B, audio_id_base);
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// slot-resident decode state: no host work between CUDA-graph replays
// ------------------------------------------------------------------------------------------
namespace vb {
// state = {kv_len[B], position[B]} advance by one token per request (worker/base.py:312-325 does this on
// the host); active[b] == 0 freezes a slot.
__global__ void decode_advance_kernel(int32_t* kv_len, int32_t* pos, const int32_t* active, int B) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && (!active || active[b])) {
    kv_len[b] += 1;
    pos[b] += 1;
  }
}
// sampled ids of batch row b (int64) -> slot s = slots[b] (identity when null): next input id, token history
// ring history[s][n_out[s] % cap], ++n_out[s]   (orpheus.py:447-448, 456-458 keep these in Python lists)
__global__ void token_feedback_kernel(const int64_t* ids, const int32_t* slots, int32_t* next_input,
                                      int32_t* history, int32_t* n_out, int B, int cap, int skip_token) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int s = slots ? slots[b] : b;
  const int v = static_cast<int>(ids[b]);
  next_input[s] = v;
  // the stop id is fed back (a scheduler running one step ahead issues one more LM step) but never becomes an audio
  // token: the host pops it from lm_output_audio_tokens (orpheus.py:461-463), so the ring must not hold it either
  if (v == skip_token) return;
  const int n = n_out[s];
  if (history) history[static_cast<size_t>(s) * cap + (n % cap)] = v;
  n_out[s] = n + 1;
}
// next step's input ids gathered by slot: ids[b] = next_input[slots[b]]
__global__ void gather_i32_kernel(int32_t* out, const int32_t* src, const int32_t* idx, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx ? idx[i] : i];
}
// step input ids: decode rows take the id sampled last step for their slot, prefill rows the uploaded prompt id
__global__ void build_input_ids_kernel(int32_t* out, const int32_t* host_ids, const int32_t* next_input,
                                       const int32_t* row_slot, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = row_slot[i];
  out[i] = s >= 0 ? next_input[s] : host_ids[i];
}
// newest full window per slot: first[i] = n_out[slot[i]] - window (device-resident loop: the host never
// sees the token counters between replays)
__global__ void latest_window_kernel(int32_t* first, const int32_t* n_out, const int32_t* slot, int n, int window) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) first[i] = max(0, n_out[slot ? slot[i] : i] - window);
}
// windows[i][j] = history[slot[i]][(first[i] + min(j, n_valid[i] - 1)) % cap], j < win : the detokenize window
// of cuda_graph_worker.py:1176-1190 (a short last window repeats its final token, :1183-1185).
__global__ void gather_windows_kernel(int64_t* windows, const int32_t* history, const int32_t* slot,
                                      const int32_t* first, const int32_t* n_valid, int n, int cap, int win) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * win) return;
  const int w = i / win, j = i - w * win;
  const int nv = n_valid ? n_valid[w] : win;
  const int jj = j < nv ? j : nv - 1;
  windows[i] = history[static_cast<size_t>(slot[w]) * cap + ((first[w] + jj) % cap)];
}
}  // namespace vb

extern "C" {
int vb_decode_advance(int32_t* d_kv_len, int32_t* d_pos, const int32_t* d_active, int B, void* stream) {
  VB_CHECK_ARG(d_kv_len && d_pos, "vb_decode_advance: null pointer");
  if (B <= 0) return 0;
  VB_LAUNCH_PDL(vb::decode_advance_kernel, (B + 127) / 128, 128, 0, stream, d_kv_len, d_pos, d_active, B);
  return 0;
}
int vb_token_feedback(const int64_t* d_ids, const int32_t* d_slots, int32_t* d_next_input, int32_t* d_history,
                      int32_t* d_n_out, int B, int history_cap, int skip_token, void* stream) {
  VB_CHECK_ARG(d_ids && d_next_input && d_n_out, "vb_token_feedback: null pointer");
  VB_CHECK_ARG(history_cap > 0, "vb_token_feedback: history_cap must be positive");
  if (B <= 0) return 0;
  VB_LAUNCH_PDL(vb::token_feedback_kernel, (B + 127) / 128, 128, 0, stream, d_ids, d_slots, d_next_input, d_history, d_n_out, B, history_cap,
                skip_token);
  return 0;
}
int vb_gather_i32(int32_t* d_out, const int32_t* d_src, const int32_t* d_idx, int n, void* stream) {
  VB_CHECK_ARG(d_out && d_src, "vb_gather_i32: null pointer");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(vb::gather_i32_kernel, (n + 127) / 128, 128, 0, stream, d_out, d_src, d_idx, n);
  return 0;
}
int vb_build_input_ids(int32_t* d_out, const int32_t* d_host_ids, const int32_t* d_next_input,
                       const int32_t* d_row_slot, int n, void* stream) {
  VB_CHECK_ARG(d_out && d_host_ids && d_next_input && d_row_slot, "vb_build_input_ids: null pointer");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(vb::build_input_ids_kernel, (n + 127) / 128, 128, 0, stream, d_out, d_host_ids, d_next_input, d_row_slot, n);
  return 0;
}
int vb_latest_window(int32_t* d_first, const int32_t* d_n_out, const int32_t* d_slot, int n, int window,
                     void* stream) {
  VB_CHECK_ARG(d_first && d_n_out, "vb_latest_window: null pointer");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(vb::latest_window_kernel, (n + 127) / 128, 128, 0, stream, d_first, d_n_out, d_slot, n, window);
  return 0;
}
int vb_gather_windows(int64_t* d_windows, const int32_t* d_history, const int32_t* d_slot, const int32_t* d_first,
                      const int32_t* d_n_valid, int n, int history_cap, int window, void* stream) {
  VB_CHECK_ARG(d_windows && d_history && d_slot && d_first, "vb_gather_windows: null pointer");
  if (n <= 0) return 0;
  VB_LAUNCH_PDL(vb::gather_windows_kernel, (n * window + 127) / 128, 128, 0, stream, d_windows, d_history, d_slot, d_first, d_n_valid, n, history_cap, window);
  return 0;
}
}  // extern "C"

VB_DEFINE_TRACE_SETTER(elementwise)
