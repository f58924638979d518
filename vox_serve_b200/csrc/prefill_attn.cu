// Tiled paged attention for prefill-shaped steps (many query rows per request).
// Replaces FlashInferPrefillWrapper.run (vox_serve/flashinfer_utils.py:68-80, 132; causal=True) for steps whose
// rows are mostly prompt rows; vb_paged_attn (attn.cu) treats every query row as its own KV stream, which is right
// for decode (one row per request) but re-reads a request's K/V once per prompt row.
//
// Structure (FlashAttention-2 on mma.sync tensor-core tiles, paged K/V):
//  * a Q TILE is up to TQ = 16 * slabs consecutive prompt rows of ONE request (tiles start at the request's first
//    row, so a tile never mixes requests); a CTA owns (Q tile, kv head, chunk of GC grouped query heads).  Tile b
//    is resolved on the device from qo_indptr with a block-wide scan -- no host-side tile list, CUDA-graph safe;
//    the grid is the upper bound ceil(rows / TQ) + requests and surplus CTAs exit;
//  * warp w = (query head gq = w % GC of the group, 16-row slab w / GC) keeps its 16 x D Q fragment, fp32 output
//    accumulators and online-softmax state in registers;
//  * K and V tiles of KT tokens of the CTA's kv head (KT * D * 2 bytes each; a tile never crosses a page) travel
//    through a two-stage cp.async ring, rows skewed by 16 bytes so ldmatrix reads are bank-conflict free; every
//    warp of the CTA shares them, i.e. a request's K/V is read once per TQ rows x GC heads instead of once per row;
//  * S = Q K^T and O += P V on mma.sync.m16n8k16 (bf16 in, fp32 accumulate); exp2 via ex2.approx, P rounded to
//    bf16 and the denominator summed from the rounded P (FlashInfer FA2 numerics, as attn.cu);
//  * the mask is a per-row key bound (row_kvlen of the plan: causal prefix for LM prompts; any non-decreasing
//    bound, e.g. the block-causal mask of a Whisper-style encoder, works the same way).  KV tiles beyond a warp's
//    largest bound are skipped by that warp.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 1
#include "common.cuh"

namespace vb {

struct PrefillAttnParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* q;
  const __nv_bfloat16* kv;       // whole cache [slabs][2][page_size][n_kv][D]
  const int32_t* qo_indptr;      // [n_req + 1]
  const int32_t* kv_indptr;      // [n_req + 1]
  const int32_t* kv_indices;
  const int32_t* row_kvlen;      // [n_rows] keys visible to the row
  int slab_base;
  int n_req, n_rows, n_q, n_kv, G, GC, slabs, page_size;
  int out_xt_tile;               // 0: out is [row][n_q][D]; else the tiled XT layout of [row][n_q * D]
  float scale_log2;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float pf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ size_t pf_out_index(const PrefillAttnParams& p, int row, int hq, int d, int D) {
  if (p.out_xt_tile == 0) return (static_cast<size_t>(row) * p.n_q + hq) * D + d;
  return xt_index(row, hq * D + d, p.out_xt_tile, (p.n_q * D + 63) >> 6);
}

template <int D, int KT>
__global__ void __launch_bounds__(256, 1) paged_prefill_attn_kernel(const PrefillAttnParams p) {
  constexpr int KS = D / 16;          // k-steps of Q K^T
  constexpr int ND = D / 8;           // 8-wide output column tiles
  constexpr int NT = KT / 8;          // 8-token column tiles of S
  constexpr int RS = D * 2 + 16;      // shared-memory row stride (bytes): 16-byte skew
  constexpr int CPR = D * 2 / 16;     // 16-byte chunks per K/V row
  __shared__ __align__(16) uint8_t kvbuf[2][2][KT * RS];   // [stage][K|V]
  __shared__ int s_warp[8];
  __shared__ int s_tile[3];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NTHR = blockDim.x, NWARP = NTHR >> 5;
  const int TQ = p.slabs * 16;

  pdl_sync();

  // rows beyond the last request (CUDA-graph padding) produce zeros
  if (blockIdx.y == 0 && blockIdx.z == 0) {
    const int n_valid = p.qo_indptr[p.n_req];
    for (int row = n_valid + blockIdx.x; row < p.n_rows; row += gridDim.x)
      for (int i = tid; i < p.n_q * D / 2; i += NTHR)
        *reinterpret_cast<uint32_t*>(p.out + pf_out_index(p, row, (2 * i) / D, (2 * i) % D, D)) = 0u;
  }

  // ---- which (request, rows) is tile blockIdx.x?  prefix sum of ceil(n_new / TQ) over the requests ----
  if (tid == 0) s_tile[0] = -1;
  int base = 0;
  for (int r0 = 0; r0 < p.n_req; r0 += NTHR) {
    __syncthreads();
    const int r = r0 + tid;
    int nt = 0, row0 = 0, nnew = 0;
    if (r < p.n_req) {
      row0 = p.qo_indptr[r];
      nnew = p.qo_indptr[r + 1] - row0;
      nt = (nnew + TQ - 1) / TQ;
    }
    int incl = nt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < NWARP; ++w) {
      const int v = s_warp[w];
      if (w < warp) woff += v;
      tot += v;
    }
    const int excl = base + woff + incl - nt;
    const int b = static_cast<int>(blockIdx.x);
    if (nt > 0 && b >= excl && b < excl + nt) {
      s_tile[0] = r;
      s_tile[1] = row0 + (b - excl) * TQ;
      s_tile[2] = min(TQ, nnew - (b - excl) * TQ);
    }
    base += tot;
  }
  __syncthreads();
  const int req = s_tile[0];
  if (req < 0) return;
  const int row0 = s_tile[1], nrows = s_tile[2];

  const int hk = blockIdx.y;
  const int gq = warp % p.GC, slab = warp / p.GC;
  const int hq = hk * p.G + static_cast<int>(blockIdx.z) * p.GC + gq;
  const int pbase = p.kv_indptr[req];
  const int g = lane >> 2, qd = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const int rA = slab * 16 + g, rB = rA + 8;            // this lane's two rows of the tile (C-fragment rows)
  const bool okA = rA < nrows, okB = rB < nrows;
  const int klA = okA ? p.row_kvlen[row0 + rA] : 0;
  const int klB = okB ? p.row_kvlen[row0 + rB] : 0;
  int wmax = max(klA, klB);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  __syncthreads();                    // s_warp is reused
  if (lane == 0) s_warp[warp] = wmax;
  __syncthreads();
  int kmax = 0;
  for (int w = 0; w < NWARP; ++w) kmax = max(kmax, s_warp[w]);
  const int n_tiles = (kmax + KT - 1) / KT;

  // ---- Q fragments (A operand, rows = prompt rows of this slab, k = head dim) ----
  uint32_t qa[KS][4];
  {
    const __nv_bfloat16* qA = p.q + (static_cast<size_t>(row0 + (okA ? rA : 0)) * p.n_q + hq) * D;
    const __nv_bfloat16* qB = p.q + (static_cast<size_t>(row0 + (okB ? rB : 0)) * p.n_q + hq) * D;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int d0 = ks * 16 + qd * 2;
      qa[ks][0] = okA ? *reinterpret_cast<const uint32_t*>(qA + d0) : 0u;
      qa[ks][1] = okB ? *reinterpret_cast<const uint32_t*>(qB + d0) : 0u;
      qa[ks][2] = okA ? *reinterpret_cast<const uint32_t*>(qA + d0 + 8) : 0u;
      qa[ks][3] = okB ? *reinterpret_cast<const uint32_t*>(qB + d0 + 8) : 0u;
    }
  }

  const size_t page_elems = static_cast<size_t>(p.page_size) * p.n_kv * D;
  auto load_tile = [&](int t, int st) {
    const int tok0 = t * KT;
    const int page = __ldg(&p.kv_indices[pbase + tok0 / p.page_size]);
    const int slot0 = tok0 % p.page_size;
    const __nv_bfloat16* src = p.kv + static_cast<size_t>(p.slab_base + page) * 2 * page_elems +
                               (static_cast<size_t>(slot0) * p.n_kv + hk) * D;
    for (int i = tid; i < 2 * KT * CPR; i += NTHR) {
      const int kvsel = i / (KT * CPR);
      const int j = i - kvsel * (KT * CPR);
      const int tok = j / CPR, c = j - tok * CPR;
      cp_async16(smem_u32(&kvbuf[st][kvsel][tok * RS + c * 16]),
                 src + kvsel * page_elems + static_cast<size_t>(tok) * p.n_kv * D + c * 8);
    }
  };

  float O[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) O[nd][0] = O[nd][1] = O[nd][2] = O[nd][3] = 0.f;
  float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;

  if (n_tiles > 0) {
    load_tile(0, 0);
    cp_async_commit();
  }
  for (int t = 0; t < n_tiles; ++t) {
    const int st = t & 1;
    if (t + 1 < n_tiles) {
      load_tile(t + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int tok0 = t * KT;
    if (tok0 < wmax) {
      const uint32_t kbase = smem_u32(&kvbuf[st][0][0]);
      const uint32_t vbase = smem_u32(&kvbuf[st][1][0]);
      // ---- S = Q K^T : 16 rows x KT tokens ----
      float S[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
        const uint32_t rb = kbase + (nt * 8 + r8) * RS + ((mi >> 1) * 16 + (mi & 1) * 8) * 2;
#pragma unroll
        for (int k2 = 0; k2 < KS / 2; ++k2) {
          uint32_t b[4];
          ldmatrix_x4(b, rb + k2 * 64);
          mma_bf16_16816(S[nt], qa[2 * k2], b[0], b[1]);
          mma_bf16_16816(S[nt], qa[2 * k2 + 1], b[2], b[3]);
        }
      }
      // ---- mask, running max, P = exp2(S - m) rounded to bf16 ----
      float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int tk = tok0 + nt * 8 + qd * 2;
        S[nt][0] = (tk < klA) ? S[nt][0] * p.scale_log2 : -INFINITY;
        S[nt][1] = (tk + 1 < klA) ? S[nt][1] * p.scale_log2 : -INFINITY;
        S[nt][2] = (tk < klB) ? S[nt][2] * p.scale_log2 : -INFINITY;
        S[nt][3] = (tk + 1 < klB) ? S[nt][3] * p.scale_log2 : -INFINITY;
        mxA = fmaxf(mxA, fmaxf(S[nt][0], S[nt][1]));
        mxB = fmaxf(mxB, fmaxf(S[nt][2], S[nt][3]));
      }
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, o));
        mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, o));
      }
      const float nA = fmaxf(mA, mxA), nB = fmaxf(mB, mxB);
      const float uA = (nA == -INFINITY) ? 0.f : nA, uB = (nB == -INFINITY) ? 0.f : nB;   // all masked so far
      const float alA = pf_ex2(mA - uA), alB = pf_ex2(mB - uB);                           // ex2(-inf) = 0
      mA = nA; mB = nB;
      float sumA = 0.f, sumB = 0.f;
      uint32_t pa[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float p0 = round_bf16(pf_ex2(S[nt][0] - uA)), p1 = round_bf16(pf_ex2(S[nt][1] - uA));
        const float p2 = round_bf16(pf_ex2(S[nt][2] - uB)), p3 = round_bf16(pf_ex2(S[nt][3] - uB));
        sumA += p0 + p1;
        sumB += p2 + p3;
        pa[nt][0] = pack_bf16(p0, p1);
        pa[nt][1] = pack_bf16(p2, p3);
      }
      lA = lA * alA + sumA;
      lB = lB * alB + sumB;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        O[nd][0] *= alA; O[nd][1] *= alA;
        O[nd][2] *= alB; O[nd][3] *= alB;
      }
      // ---- O += P V ----
#pragma unroll
      for (int kt = 0; kt < KT / 16; ++kt) {
        const uint32_t a[4] = {pa[2 * kt][0], pa[2 * kt][1], pa[2 * kt + 1][0], pa[2 * kt + 1][1]};
        const uint32_t rb = vbase + (kt * 16 + (mi & 1) * 8 + r8) * RS + ((mi >> 1) * 8) * 2;
#pragma unroll
        for (int n2 = 0; n2 < ND / 2; ++n2) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, rb + n2 * 32);
          mma_bf16_16816(O[2 * n2], a, b[0], b[1]);
          mma_bf16_16816(O[2 * n2 + 1], a, b[2], b[3]);
        }
      }
    }
    __syncthreads();     // stage st is overwritten by the load of tile t + 2
  }

  // ---- normalise and store ----
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    lA += __shfl_xor_sync(0xffffffffu, lA, o);
    lB += __shfl_xor_sync(0xffffffffu, lB, o);
  }
  const float iA = lA > 0.f ? 1.f / lA : 0.f, iB = lB > 0.f ? 1.f / lB : 0.f;
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int d = nd * 8 + qd * 2;
    if (okA)
      *reinterpret_cast<uint32_t*>(p.out + pf_out_index(p, row0 + rA, hq, d, D)) = pack_bf16(O[nd][0] * iA, O[nd][1] * iA);
    if (okB)
      *reinterpret_cast<uint32_t*>(p.out + pf_out_index(p, row0 + rB, hq, d, D)) = pack_bf16(O[nd][2] * iB, O[nd][3] * iB);
  }
}

template <int D, int KT>
static int launch_prefill_attn(const PrefillAttnParams& p, dim3 grid, int threads, cudaStream_t stream) {
  auto kern = paged_prefill_attn_kernel<D, KT>;
  VB_LAUNCH_PDL(kern, grid, threads, 0, stream, p);
  return 0;
}

// grouped query heads one CTA takes (its warps = GC x slabs <= 8) and 16-row slabs per Q tile
static void prefill_attn_shape(int G, int* gc, int* slabs) {
  int c = G <= 8 ? G : 8;
  while (G % c != 0) --c;
  *gc = c;
  *slabs = 8 / c < 1 ? 1 : 8 / c;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_prefill_attn_tile_rows(int n_q, int n_kv) {
  if (n_kv < 1 || n_q < n_kv || n_q % n_kv != 0) return -1;
  int gc, slabs;
  prefill_attn_shape(n_q / n_kv, &gc, &slabs);
  return 16 * slabs;
}

int vb_paged_prefill_attn(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                          const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                          const int32_t* d_row_kvlen, int n_req, int n_rows, int n_q, int n_kv, int head_dim,
                          int page_size, float sm_scale, int out_xt_tile, void* stream) {
  VB_CHECK_ARG(d_out && d_q && d_kv && d_qo_indptr && d_kv_indptr && d_kv_indices && d_row_kvlen,
               "vb_paged_prefill_attn: null pointer");
  VB_CHECK_ARG(n_kv > 0 && n_q % n_kv == 0, "vb_paged_prefill_attn: %d query heads / %d kv heads", n_q, n_kv);
  VB_CHECK_ARG(head_dim == 64 || head_dim == 128, "vb_paged_prefill_attn: head_dim %d unsupported (64, 128)", head_dim);
  VB_CHECK_ARG(page_size >= 16 && page_size % 16 == 0, "vb_paged_prefill_attn: page_size %d must be a multiple of 16",
               page_size);
  if (n_rows <= 0 || n_req <= 0) return 0;
  PrefillAttnParams p;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.kv = static_cast<const __nv_bfloat16*>(d_kv);
  p.qo_indptr = d_qo_indptr;
  p.kv_indptr = d_kv_indptr;
  p.kv_indices = d_kv_indices;
  p.row_kvlen = d_row_kvlen;
  p.slab_base = static_cast<int>(slab_base);
  p.n_req = n_req; p.n_rows = n_rows; p.n_q = n_q; p.n_kv = n_kv; p.G = n_q / n_kv; p.page_size = page_size;
  prefill_attn_shape(p.G, &p.GC, &p.slabs);
  p.out_xt_tile = out_xt_tile;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  const int tq = 16 * p.slabs;
  const dim3 grid((n_rows + tq - 1) / tq + n_req, n_kv, p.G / p.GC);
  const int threads = 32 * p.GC * p.slabs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool kt32 = page_size % 32 == 0;
  if (head_dim == 128)
    return kt32 ? launch_prefill_attn<128, 32>(p, grid, threads, st) : launch_prefill_attn<128, 16>(p, grid, threads, st);
  return kt32 ? launch_prefill_attn<64, 32>(p, grid, threads, st) : launch_prefill_attn<64, 16>(p, grid, threads, st);
}

}  // extern "C"
