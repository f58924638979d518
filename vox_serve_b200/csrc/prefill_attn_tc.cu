// Tiled paged prefill attention on the 5th-generation tensor cores (tcgen05 + TMEM): the FlashAttention schedule of
// prefill_attn.cu with both contractions issued as tcgen05.mma from shared-memory operand tiles and their accumulators
// in tensor memory.  Same inputs, same result contract (FlashInferPrefillWrapper.run, flashinfer_utils.py:68-80, 132).
//
//  * a CTA owns (Q tile, kv head).  The Q tile is M = 128 accumulator rows = TQ = 128 / G consecutive prompt rows of ONE
//    request x the G grouped query heads of the kv head (row r = token r / G, head r % G): the GQA group rides in the MMA's
//    M dimension, so a K/V tile is fetched once for all of them;
//  * operand tiles are the 128-byte-swizzled K-major images the projection kernel uses (xt_index): Q [128 rows][D],
//    K [128 tokens][D] (B operand of S = Q K^T), P [128 rows][128 tokens] bf16 (A operand of O = P V) and V TRANSPOSED
//    [D dims][128 tokens] (B operand, K-major = tokens contiguous).  K arrives with cp.async straight into its swizzled
//    place; V arrives row-major in a staging buffer (cp.async) and is transposed shared -> shared;
//  * S (128 x 128 fp32) and the per-tile product P V (128 x D fp32) live in TMEM (256 columns); thread r of the CTA's 128
//    threads owns accumulator row r = TMEM lane r: it reads its S row with tcgen05.ld, keeps the online-softmax state
//    and the running output row in registers (no shuffles: a row is one thread), writes its P row as the next MMA's
//    operand; exp2 via ex2.approx, P rounded to bf16, denominator from the rounded P (as the other attention kernels);
//  * one elected thread issues the MMAs; completion comes back through mbarriers (tcgen05.commit); the K tile of the next
//    step is requested as soon as S is complete and its V rows as soon as the transpose has read the staging buffer, so
//    the loads run under the softmax and the second MMA.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 1
#include "common.cuh"

namespace vb {

struct PrefillTcParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* q;
  const __nv_bfloat16* kv;       // whole cache [slabs][2][page_size][n_kv][D]
  const int32_t* qo_indptr;      // [n_req + 1]
  const int32_t* kv_indptr;      // [n_req + 1]
  const int32_t* kv_indices;
  const int32_t* row_kvlen;      // [n_rows] keys visible to the row
  int slab_base;
  int n_req, n_rows, n_q, n_kv, G, TQ, page_size;
  int out_xt_tile;
  float scale_log2;
};

constexpr int PTC_THREADS = 128;

__device__ __forceinline__ void ptc_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ptc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ptc_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float ptc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// KT = tokens per K/V tile: 128 at head_dim 64, 64 at head_dim 128 -- either way a CTA stays below half an SM's shared
// memory (98 KB) and half its tensor memory (256 columns), so TWO CTAs share an SM and one's loads / softmax run under
// the other's MMAs (the phases inside a CTA are serial).
template <int D, int KT>
struct PtcLayout {
  static constexpr int Q = 0;                               // [D/64][128][64] bf16
  static constexpr int K = Q + 128 * D * 2;                 // [D/64][KT][64]
  static constexpr int VT = K + KT * D * 2;                 // [KT/64][D][64]
  static constexpr int P = VT + KT * D * 2;                 // [KT/64][128][64]
  static constexpr int VRAW = P + 128 * KT * 2;             // [KT][D * 2 + 16 bytes]
  static constexpr int VRS = D * 2 + 16;
  static constexpr int BAR = VRAW + KT * VRS;               // bar_s, bar_o, tmem slot
  static constexpr int TOTAL = BAR + 64 + 1024;             // + alignment slack
};

template <int D, int KT>
__global__ void __launch_bounds__(PTC_THREADS, 1) paged_prefill_attn_tc_kernel(const PrefillTcParams p) {
  using L = PtcLayout<D, KT>;
  constexpr int PTC_KT = KT;
  constexpr int KBD = D / 64;          // 64-element k-blocks of the head dim
  constexpr int KBT = PTC_KT / 64;     // ... of a token tile
  constexpr int CPR = D * 2 / 16;      // 16-byte chunks per K / V row
  constexpr int TPT = PTC_THREADS / PTC_KT;      // threads that share one token's K / V row (1 or 2)
  constexpr int CPT = CPR / TPT;                 // 16-byte chunks of the row per thread
  extern __shared__ uint8_t ptc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ptc_smem_raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + L::Q);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + L::K);
  __nv_bfloat16* sVT = reinterpret_cast<__nv_bfloat16*>(smem + L::VT);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(smem + L::P);
  uint8_t* sVraw = smem + L::VRAW;
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 2);
  __shared__ int s_warp[PTC_THREADS / 32];
  __shared__ int s_tile[3];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    s_tile[0] = -1;
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_sync();

  // rows beyond the last request (CUDA-graph padding) produce zeros
  if (blockIdx.y == 0) {
    const int n_valid = p.qo_indptr[p.n_req];
    for (int row = n_valid + blockIdx.x; row < p.n_rows; row += gridDim.x)
      for (int i = tid; i < p.n_q * D / 8; i += PTC_THREADS) {
        const int e = i * 8;
        const size_t o = p.out_xt_tile ? xt_index(row, e, p.out_xt_tile, (p.n_q * D + 63) >> 6)
                                       : static_cast<size_t>(row) * p.n_q * D + e;
        *reinterpret_cast<uint4*>(p.out + o) = make_uint4(0u, 0u, 0u, 0u);
      }
  }

  // ---- which (request, rows) is tile blockIdx.x?  prefix sum of ceil(n_new / TQ) over the requests ----
  const int TQ = p.TQ;
  int base = 0;
  for (int r0 = 0; r0 < p.n_req; r0 += PTC_THREADS) {
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
    for (int w = 0; w < PTC_THREADS / 32; ++w) {
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
  if (req >= 0) {
    const int row0 = s_tile[1], nrows = s_tile[2];
    const int hk = blockIdx.y;
    const int pbase = p.kv_indptr[req];
    // this thread's accumulator row: prompt row t of the tile, grouped head g
    const int t = tid / p.G, g = tid - t * p.G;
    const bool row_ok = t < nrows;
    const int kl = row_ok ? p.row_kvlen[row0 + t] : 0;
    // the tile's largest bound (the bounds do not decrease inside a request: the last row's)
    const int kmax = p.row_kvlen[row0 + nrows - 1];
    const int n_tiles = (kmax + PTC_KT - 1) / PTC_KT;
    const size_t page_elems = static_cast<size_t>(p.page_size) * p.n_kv * D;

    // ---- Q tile: row r = tid -> swizzled K-major image ----
    {
      const __nv_bfloat16* src = p.q + (static_cast<size_t>(row0 + (row_ok ? t : 0)) * p.n_q + hk * p.G + g) * D;
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok) v = *reinterpret_cast<const uint4*>(src + c * 8);
        *reinterpret_cast<uint4*>(sQ + xt_index(tid, c * 8, 128, KBD)) = v;
      }
    }
    // ---- K / V loads: thread j fetches token tok0 + j of the CTA's kv head ----
    auto token_src = [&](int tok) -> const __nv_bfloat16* {
      const int page = __ldg(&p.kv_indices[pbase + tok / p.page_size]);
      return p.kv + static_cast<size_t>(p.slab_base + page) * 2 * page_elems +
             (static_cast<size_t>(tok % p.page_size) * p.n_kv + hk) * D;
    };
    const int ltok = tid / TPT, lc0 = (tid % TPT) * CPT;      // this thread's token of a tile and its first chunk
    auto load_k = [&](int tile) {
      const int tok = tile * PTC_KT + ltok;
      if (tok < kmax) {
        const __nv_bfloat16* src = token_src(tok);
#pragma unroll
        for (int c = lc0; c < lc0 + CPT; ++c) ptc_cp_async16(smem_u32(sK + xt_index(ltok, c * 8, PTC_KT, KBD)), src + c * 8);
      } else {
#pragma unroll
        for (int c = lc0; c < lc0 + CPT; ++c)
          *reinterpret_cast<uint4*>(sK + xt_index(ltok, c * 8, PTC_KT, KBD)) = make_uint4(0u, 0u, 0u, 0u);
      }
      ptc_cp_commit();
    };
    auto load_v = [&](int tile) {
      const int tok = tile * PTC_KT + ltok;
      uint8_t* dst = sVraw + ltok * L::VRS;
      if (tok < kmax) {
        const __nv_bfloat16* src = token_src(tok) + page_elems;
#pragma unroll
        for (int c = lc0; c < lc0 + CPT; ++c) ptc_cp_async16(smem_u32(dst + c * 16), src + c * 8);
      } else {
#pragma unroll
        for (int c = lc0; c < lc0 + CPT; ++c) *reinterpret_cast<uint4*>(dst + c * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      ptc_cp_commit();
    };

    float O[D];
#pragma unroll
    for (int i = 0; i < D; ++i) O[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t idesc_s = umma_idesc(128, PTC_KT, 1u);
    const uint32_t idesc_o = umma_idesc(128, D, 1u);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);   // this warp's TMEM lane quarter

    if (n_tiles > 0) {
      load_k(0);
      load_v(0);
    }
    uint32_t ph = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int tok0 = tile * PTC_KT;
      // ---- K (and, for tile 0, Q) in place -> S = Q K^T ----
      // tile 0: groups were committed K then V -- the K group is enough here; later tiles: V (requested first) and K
      if (tile == 0) ptc_cp_wait<1>(); else ptc_cp_wait<0>();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KBD; ++kb) {
          const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(sQ) + kb * (128 * 128));
          const uint64_t b_desc = umma_desc_sw128_kmajor(smem_u32(sK) + kb * (PTC_KT * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_s);
      }
      // ---- meanwhile: V rows -> transposed, swizzled operand tile ----
      ptc_cp_wait<0>();
      __syncthreads();               // every thread's V row has landed in the staging buffer
      {
        // this thread's token of the tile: its (part of the) row, 8 dims at a time, scattered to rows d of the [D][KT] tile
        const uint8_t* srow = sVraw + ltok * L::VRS;
#pragma unroll 4
        for (int c = lc0; c < lc0 + CPT; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(srow + c * 16);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint16_t h = static_cast<uint16_t>((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
            *reinterpret_cast<uint16_t*>(sVT + xt_index(c * 8 + e, ltok, D, KBT)) = h;
          }
        }
      }
      __syncthreads();               // the staging buffer has been read by everybody
      if (tile + 1 < n_tiles) load_v(tile + 1);
      // ---- S complete: this thread's row ----
      mbar_wait(bar_s, ph);
      tc_fence_after();
      if (tile + 1 < n_tiles) load_k(tile + 1);      // the first MMA has read K: fetch the next tile under the softmax
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < PTC_KT / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(t_lane + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (tok0 + c * 16 + i < kl) mx = fmaxf(mx, __uint_as_float(v[i]) * p.scale_log2);
      }
      const float m_new = fmaxf(m_run, mx);
      const float u = (m_new == -INFINITY) ? 0.f : m_new;       // every key masked so far
      const float alpha = ptc_ex2(m_run - u);                    // ex2(-inf) = 0
      m_run = m_new;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < PTC_KT / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(t_lane + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          uint32_t pk[4];
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            const int i = i8 * 8 + i2 * 2, j = tok0 + c * 16 + i;
            const float p0 = (j < kl) ? round_bf16(ptc_ex2(__uint_as_float(v[i]) * p.scale_log2 - u)) : 0.f;
            const float p1 = (j + 1 < kl) ? round_bf16(ptc_ex2(__uint_as_float(v[i + 1]) * p.scale_log2 - u)) : 0.f;
            lsum += p0 + p1;
            pk[i2] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(sP + xt_index(tid, c * 16 + i8 * 8, 128, KBT)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      l_run = l_run * alpha + lsum;
      // ---- O_tile = P V ----
      fence_proxy_async();           // P and V^T were written with ordinary stores
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KBT; ++kb) {
          const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(sP) + kb * (128 * 128));
          const uint64_t b_desc = umma_desc_sw128_kmajor(smem_u32(sVT) + kb * (D * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + 128, a_desc + 2 * k, b_desc + 2 * k, idesc_o, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
      mbar_wait(bar_o, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(t_lane + 128 + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) O[c * 16 + i] = O[c * 16 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();             // TMEM reads done before the next tile's MMAs overwrite S / O_tile
      ph ^= 1;
    }

    // ---- normalise and store this thread's row ----
    if (row_ok) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int hq = hk * p.G + g;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(O[c * 8 + 2 * i] * inv, O[c * 8 + 2 * i + 1] * inv);
        const size_t o = p.out_xt_tile ? xt_index(row0 + t, hq * D + c * 8, p.out_xt_tile, (p.n_q * D + 63) >> 6)
                                       : (static_cast<size_t>(row0 + t) * p.n_q + hq) * D + c * 8;
        *reinterpret_cast<uint4*>(p.out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int D, int KT>
static int launch_prefill_tc(const PrefillTcParams& p, dim3 grid, cudaStream_t stream) {
  auto kern = paged_prefill_attn_tc_kernel<D, KT>;
  // (a fixed size per instantiation -- safe under CUDA-graph replay -- and below the limit together with the few
  // bytes of static shared memory)
  constexpr int smem_bytes = PtcLayout<D, KT>::TOTAL;
  VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  VB_LAUNCH_PDL(kern, grid, PTC_THREADS, smem_bytes, stream, p);
  return 0;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_prefill_attn_tc_tile_rows(int n_q, int n_kv) {
  if (n_kv < 1 || n_q < n_kv || n_q % n_kv != 0 || n_q / n_kv > 128) return -1;
  return 128 / (n_q / n_kv);
}

int vb_paged_prefill_attn_tc(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                             const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                             const int32_t* d_row_kvlen, int n_req, int n_rows, int n_q, int n_kv, int head_dim,
                             int page_size, float sm_scale, int out_xt_tile, void* stream) {
  VB_CHECK_ARG(d_out && d_q && d_kv && d_qo_indptr && d_kv_indptr && d_kv_indices && d_row_kvlen,
               "vb_paged_prefill_attn_tc: null pointer");
  VB_CHECK_ARG(n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= 128, "vb_paged_prefill_attn_tc: %d query heads / %d kv heads",
               n_q, n_kv);
  VB_CHECK_ARG(head_dim == 64 || head_dim == 128, "vb_paged_prefill_attn_tc: head_dim %d unsupported (64, 128)", head_dim);
  VB_CHECK_ARG(page_size >= 1, "vb_paged_prefill_attn_tc: page_size %d", page_size);
  if (n_rows <= 0 || n_req <= 0) return 0;
  PrefillTcParams p;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.kv = static_cast<const __nv_bfloat16*>(d_kv);
  p.qo_indptr = d_qo_indptr;
  p.kv_indptr = d_kv_indptr;
  p.kv_indices = d_kv_indices;
  p.row_kvlen = d_row_kvlen;
  p.slab_base = static_cast<int>(slab_base);
  p.n_req = n_req; p.n_rows = n_rows; p.n_q = n_q; p.n_kv = n_kv; p.G = n_q / n_kv; p.page_size = page_size;
  p.TQ = 128 / p.G;
  p.out_xt_tile = out_xt_tile;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  const dim3 grid((n_rows + p.TQ - 1) / p.TQ + n_req, n_kv, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return head_dim == 128 ? launch_prefill_tc<128, 64>(p, grid, st) : launch_prefill_tc<64, 128>(p, grid, st);
}

}  // extern "C"
