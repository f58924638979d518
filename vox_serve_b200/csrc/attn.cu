// Paged GQA attention for decode and (ragged, causal) prefill rows.
// Replaces FlashInferDecodeWrapper.run / FlashInferPrefillWrapper.run (vox_serve/flashinfer_utils.py:132,
// 228-230) behind the plan produced by vb_plan_rows.
//
// Work item = (query row, chunk of CHUNK tokens of that row's KV, kv head).  A persistent grid walks the
// item list.  Warp roles per CTA: one PRODUCER warp resolves items (one 32-byte record per (row, chunk),
// written by the plan kernel) and TMA-loads the chunk's K and V tiles (5-D tensor map over the whole
// cache, 128B-swizzled [CHUNK x 64-dim] boxes) plus the group's Q rows (bulk copy) into a STAGES-deep
// shared-memory ring behind full/empty mbarriers; CHUNK/16 CONSUMER warps each own 16 tokens:
// S = Q K^T and O = P V run on mma.sync m16n8k16 with the G grouped query heads as the 16-row operand
// (the tile is read once for the whole GQA group), softmax max/sum use quad shuffles, fp32 throughout,
// P rounded to bf16 for the PV product.  Rows spanning several chunks leave (m, l, O) partials; the last
// CTA to finish a (row, head) merges them in chunk order (deterministic) and restores the counter to 0.
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

struct AttnParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* q;
  const int32_t* row_kvlen;
  const int32_t* row_chunk_start;
  const int4* rc_meta;  // 2 x int4 per (row, chunk)
  int32_t* counters;    // [n_rows * n_kv]
  float* part_ml;       // [chunks][n_kv][G][2]
  float* part_o;        // [chunks][n_kv][G][D]
  int slab_base;
  int n_rows, n_q, n_kv, G, page_size, max_chunks;
  float scale_log2;
};

struct ItemMeta {
  int row, token0, page, kvlen, n_chunks, first_rc, rc, h;
};

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int D, int CHUNK, int STAGES>
struct AttnSmem {
  static constexpr int NW = CHUNK / 16;              // consumer warps
  static constexpr int TILE = CHUNK * D * 2;         // bytes of one K (or V) tile
  static constexpr int QROW = D * 2 + 16;            // padded Q row (bank-conflict-free fragment loads)
  static constexpr int QBYTES = 16 * QROW;           // up to 16 grouped heads
  static constexpr int STAGE = (2 * TILE + QBYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_META = STAGES * STAGE;    // ItemMeta[STAGES]
  static constexpr int OFF_BAR = OFF_META + STAGES * 32;      // full[STAGES], empty[STAGES]
  static constexpr int OFF_WRED = OFF_BAR + 2 * STAGES * 8;   // float[2][NW][16]
  static constexpr int OFF_FLAG = OFF_WRED + 2 * NW * 16 * 4;
  static constexpr int OFF_ORED = (OFF_FLAG + 16 + 127) / 128 * 128;  // float[NW][G][D]
  static int bytes(int G) { return OFF_ORED + NW * G * D * 4 + 1024 /*alignment slack*/; }
};

template <int D, int CHUNK, int STAGES, bool HI>
__global__ void __launch_bounds__(CHUNK * 2 + 32) paged_attn_kernel(const AttnParams p,
                                                                    const __grid_constant__ CUtensorMap kv_map) {
  using L = AttnSmem<D, CHUNK, STAGES>;
  constexpr int NW = L::NW;
  constexpr int NC = NW * 32;           // consumer threads
  constexpr int NH = D / 64;            // 64-dim half tiles per row
  constexpr int HALF = CHUNK * 128;     // bytes of one half tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  ItemMeta* meta = reinterpret_cast<ItemMeta*>(smem + L::OFF_META);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  float* wmax = reinterpret_cast<float*>(smem + L::OFF_WRED);
  float* wsum = wmax + NW * 16;
  int* flag = reinterpret_cast<int*>(smem + L::OFF_FLAG);
  float* ored = reinterpret_cast<float*>(smem + L::OFF_ORED);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G;
  const int n_rc = p.row_chunk_start[p.n_rows];
  if (n_rc > p.max_chunks) __trap();   // plan overflowed the workspace the caller sized
  const int n_items = n_rc * p.n_kv;

  if (tid == 0) {
    prefetch_tmap(&kv_map);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == NW) {
    // ===================== producer warp =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      int item = blockIdx.x;
      int4 a = make_int4(0, 0, 0, 0), b = a;
      if (item < n_items) {
        const int rc = item / p.n_kv;
        a = __ldg(&p.rc_meta[rc * 2]);
        b = __ldg(&p.rc_meta[rc * 2 + 1]);
      }
      while (item < n_items) {
        const int rc = item / p.n_kv, h = item - rc * p.n_kv;
        const int nxt = item + gridDim.x;
        int4 na = a, nb = b;
        if (nxt < n_items) {   // prefetch the next record while this item's loads are issued
          const int nrc = nxt / p.n_kv;
          na = __ldg(&p.rc_meta[nrc * 2]);
          nb = __ldg(&p.rc_meta[nrc * 2 + 1]);
        }
        mbar_wait(&empty[stage], ph ^ 1);
        ItemMeta m;
        m.row = a.x; m.token0 = a.y; m.page = a.z; m.kvlen = a.w;
        m.n_chunks = b.x; m.first_rc = b.y; m.rc = rc; m.h = h;
        meta[stage] = m;
        uint8_t* dst = smem + stage * L::STAGE;
        mbar_arrive_expect_tx(&full[stage], 2 * L::TILE + G * D * 2);
        const int slot0 = m.token0 % p.page_size;
#pragma unroll
        for (int kv = 0; kv < 2; ++kv)
#pragma unroll
          for (int hh = 0; hh < NH; ++hh)
            tma_load_5d(dst + kv * L::TILE + hh * HALF, &kv_map, &full[stage], hh * 64, h, slot0, kv,
                        p.slab_base + m.page);
        const __nv_bfloat16* qrow = p.q + (static_cast<size_t>(m.row) * p.n_q + h * G) * D;
        for (int g = 0; g < G; ++g)
          bulk_copy_g2s(dst + 2 * L::TILE + g * L::QROW, qrow + g * D, D * 2, &full[stage]);
        a = na; b = nb;
        item = nxt;
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  // padded / empty rows produce zeros
  for (int row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
    if (p.row_kvlen[row] == 0) {
      uint32_t* o = reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(row) * p.n_q * D);
      for (int i = tid; i < p.n_q * D / 2; i += NC) o[i] = 0u;
    }
  }
  auto csync = []() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); };

  const int r0 = lane >> 2;        // head row of c[0], c[1]; r0 + 8 for c[2], c[3]
  const int cq = (lane & 3) * 2;   // column pair inside an 8-wide n-tile
  const bool v0 = r0 < G, v1 = HI && (r0 + 8 < G);

  int stage = 0;
  uint32_t ph = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    mbar_wait(&full[stage], ph);
    const ItemMeta m = meta[stage];
    const uint32_t kbase = smem_u32(smem + stage * L::STAGE);
    const uint32_t vbase = kbase + L::TILE;
    const uint8_t* qs = smem + stage * L::STAGE + 2 * L::TILE;
    const int tokw = warp * 16;

    // ---- S = Q K^T : 16 heads x 16 tokens per warp ----
    float S[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    {
      const int mi = lane >> 3, r = lane & 7;
      const int tok = tokw + (mi >> 1) * 8 + r;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        uint32_t qa[4];
        const int d0 = (ks * 16 + cq) * 2;
        qa[0] = v0 ? *reinterpret_cast<const uint32_t*>(qs + r0 * L::QROW + d0) : 0u;
        qa[1] = v1 ? *reinterpret_cast<const uint32_t*>(qs + (r0 + 8) * L::QROW + d0) : 0u;
        qa[2] = v0 ? *reinterpret_cast<const uint32_t*>(qs + r0 * L::QROW + d0 + 16) : 0u;
        qa[3] = v1 ? *reinterpret_cast<const uint32_t*>(qs + (r0 + 8) * L::QROW + d0 + 16) : 0u;
        const int c16 = ks * 2 + (mi & 1);
        const uint32_t addr = kbase + (c16 >> 3) * HALF + tok * 128 + (((c16 & 7) ^ (tok & 7)) << 4);
        uint32_t bfr[4];
        ldmatrix_x4(bfr, addr);
        mma_bf16_16816(S[0], qa, bfr[0], bfr[1]);
        mma_bf16_16816(S[1], qa, bfr[2], bfr[3]);
      }
    }
    // ---- mask + row max ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tok = m.token0 + tokw + nt * 8 + cq + (j & 1);
        const float s = (tok < m.kvlen) ? S[nt][j] * p.scale_log2 : -INFINITY;
        S[nt][j] = s;
        if (j < 2) mx0 = fmaxf(mx0, s); else mx1 = fmaxf(mx1, s);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    if (HI) {
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    }
    if ((lane & 3) == 0) {
      wmax[warp * 16 + r0] = mx0;
      wmax[warp * 16 + r0 + 8] = mx1;
    }
    csync();
    float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      M0 = fmaxf(M0, wmax[w * 16 + r0]);
      if (HI) M1 = fmaxf(M1, wmax[w * 16 + r0 + 8]);
    }
    if (!HI || M1 == -INFINITY) M1 = 0.f;   // unused head rows
    if (M0 == -INFINITY) M0 = 0.f;
    // ---- P = exp2(S - M), rounded to bf16; row sums of the rounded values ----
    uint32_t pa[4];
    float l0 = 0.f, l1 = 0.f;
    {
      float e[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = round_bf16(exp2f(S[nt][j] - (j < 2 ? M0 : M1)));
          e[nt][j] = v;
          if (j < 2) l0 += v; else l1 += v;
        }
      pa[0] = pack_bf16(e[0][0], e[0][1]);
      pa[1] = pack_bf16(e[0][2], e[0][3]);
      pa[2] = pack_bf16(e[1][0], e[1][1]);
      pa[3] = pack_bf16(e[1][2], e[1][3]);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    if (HI) {
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    }
    if ((lane & 3) == 0) {
      wsum[warp * 16 + r0] = l0;
      wsum[warp * 16 + r0 + 8] = l1;
    }
    // ---- O = P V : 16 heads x D dims over this warp's 16 tokens ----
    float O[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) O[nt][j] = 0.f;
    {
      const int mi = lane >> 3, r = lane & 7;
      const int tok = tokw + (mi & 1) * 8 + r;
#pragma unroll
      for (int dn = 0; dn < D / 16; ++dn) {
        const int c16 = dn * 2 + (mi >> 1);
        const uint32_t addr = vbase + (c16 >> 3) * HALF + tok * 128 + (((c16 & 7) ^ (tok & 7)) << 4);
        uint32_t bfr[4];
        ldmatrix_x4_trans(bfr, addr);
        mma_bf16_16816(O[2 * dn], pa, bfr[0], bfr[1]);
        mma_bf16_16816(O[2 * dn + 1], pa, bfr[2], bfr[3]);
      }
    }
    // this warp is done with the stage's tiles: hand the slot back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == STAGES) { stage = 0; ph ^= 1; }

    // ---- cross-warp reduction of O through shared memory ----
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
      const int d = nt * 8 + cq;
      if (v0) *reinterpret_cast<float2*>(&ored[(warp * G + r0) * D + d]) = make_float2(O[nt][0], O[nt][1]);
      if (v1) *reinterpret_cast<float2*>(&ored[(warp * G + r0 + 8) * D + d]) = make_float2(O[nt][2], O[nt][3]);
    }
    csync();

    const bool single = (m.n_chunks == 1);
    const size_t pbase = static_cast<size_t>(m.rc) * p.n_kv + m.h;
    for (int i = tid; i < G * D; i += NC) {
      const int g = i / D;
      float o = 0.f, l = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        o += ored[w * G * D + i];
        l += wsum[w * 16 + g];
      }
      if (single) {
        p.out[(static_cast<size_t>(m.row) * p.n_q + m.h * G) * D + i] = __float2bfloat16_rn(o / l);
      } else {
        p.part_o[pbase * G * D + i] = o;
        if (i - g * D == 0) {
          float mm = -INFINITY;
#pragma unroll
          for (int w = 0; w < NW; ++w) mm = fmaxf(mm, wmax[w * 16 + g]);
          p.part_ml[(pbase * G + g) * 2 + 0] = mm;
          p.part_ml[(pbase * G + g) * 2 + 1] = l;
        }
      }
    }
    if (!single) {
      __threadfence();
      csync();
      if (tid == 0) {
        const int old = atomicAdd(&p.counters[m.row * p.n_kv + m.h], 1);
        const int last = (old == m.n_chunks - 1);
        if (last) p.counters[m.row * p.n_kv + m.h] = 0;
        *flag = last;
      }
      csync();
      if (*flag) {
        __threadfence();
        const size_t first = static_cast<size_t>(m.first_rc) * p.n_kv + m.h;   // chunk 0 of this row
        const size_t cstride = static_cast<size_t>(p.n_kv);
        for (int i = tid; i < G * D; i += NC) {
          const int g = i / D;
          float Mx = -INFINITY;
          for (int c = 0; c < m.n_chunks; ++c)
            Mx = fmaxf(Mx, __ldcg(&p.part_ml[((first + c * cstride) * G + g) * 2]));
          float num = 0.f, den = 0.f;
          for (int c = 0; c < m.n_chunks; ++c) {
            const size_t pb = first + c * cstride;
            const float sc = exp2f(__ldcg(&p.part_ml[(pb * G + g) * 2]) - Mx);
            den += sc * __ldcg(&p.part_ml[(pb * G + g) * 2 + 1]);
            num += sc * __ldcg(&p.part_o[pb * G * D + i]);
          }
          p.out[(static_cast<size_t>(m.row) * p.n_q + m.h * G) * D + i] = __float2bfloat16_rn(num / den);
        }
      }
    }
    csync();   // wmax / wsum / ored / flag reusable
  }
}

template <int D, int CHUNK, int STAGES, bool HI>
static int launch_attn(const AttnParams& p, const CUtensorMap* map, int grid, cudaStream_t stream) {
  using L = AttnSmem<D, CHUNK, STAGES>;
  const int smem = L::bytes(p.G);
  auto kern = paged_attn_kernel<D, CHUNK, STAGES, HI>;
  VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  kern<<<grid, CHUNK * 2 + 32, smem, stream>>>(p, *map);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // namespace vb

using namespace vb;

extern "C" {

size_t vb_paged_attn_workspace_bytes(int max_rows, int max_chunks_total, int n_q, int n_kv, int head_dim) {
  const size_t G = static_cast<size_t>(n_q / (n_kv > 0 ? n_kv : 1));
  size_t counters = (static_cast<size_t>(max_rows) * n_kv * sizeof(int32_t) + 255) / 256 * 256;
  size_t ml = (static_cast<size_t>(max_chunks_total) * n_kv * G * 2 * sizeof(float) + 255) / 256 * 256;
  size_t po = static_cast<size_t>(max_chunks_total) * n_kv * G * head_dim * sizeof(float);
  return counters + ml + po;
}

int vb_paged_attn(void* d_out, const void* d_q, const void* kv_map, int64_t slab_base,
                  const int32_t* d_row_kvlen, const int32_t* d_row_chunk_start, const int32_t* d_rc_meta,
                  int n_rows, int max_chunks_total, int n_q, int n_kv, int head_dim, int page_size,
                  int chunk_tokens, float sm_scale, void* d_workspace, size_t workspace_bytes, int grid_ctas,
                  void* stream) {
  VB_CHECK_ARG(d_out && d_q && kv_map && d_row_kvlen && d_row_chunk_start && d_rc_meta && d_workspace,
               "vb_paged_attn: null pointer");
  VB_CHECK_ARG(n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= 16, "vb_paged_attn: GQA group %d/%d unsupported (<=16)",
               n_q, n_kv);
  VB_CHECK_ARG(head_dim == 64 || head_dim == 128, "vb_paged_attn: head_dim %d unsupported (64, 128)", head_dim);
  VB_CHECK_ARG(chunk_tokens == 16 || chunk_tokens == 32 || chunk_tokens == 64,
               "vb_paged_attn: chunk_tokens %d unsupported (16, 32, 64)", chunk_tokens);
  VB_CHECK_ARG(page_size % chunk_tokens == 0, "vb_paged_attn: chunk_tokens must divide page_size");
  VB_CHECK_ARG(workspace_bytes >= vb_paged_attn_workspace_bytes(n_rows, max_chunks_total, n_q, n_kv, head_dim),
               "vb_paged_attn: workspace too small");
  VB_CHECK_ARG(grid_ctas > 0, "vb_paged_attn: grid_ctas must be positive");
  VB_CHECK_ARG((reinterpret_cast<uintptr_t>(d_rc_meta) & 31) == 0, "vb_paged_attn: rc_meta must be 32-byte aligned");
  if (n_rows <= 0) return 0;
  const int G = n_q / n_kv;
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.row_kvlen = d_row_kvlen;
  p.row_chunk_start = d_row_chunk_start;
  p.rc_meta = reinterpret_cast<const int4*>(d_rc_meta);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const size_t counters = (static_cast<size_t>(n_rows) * n_kv * sizeof(int32_t) + 255) / 256 * 256;
  const size_t ml = (static_cast<size_t>(max_chunks_total) * n_kv * G * 2 * sizeof(float) + 255) / 256 * 256;
  p.counters = reinterpret_cast<int32_t*>(ws);
  p.part_ml = reinterpret_cast<float*>(ws + counters);
  p.part_o = reinterpret_cast<float*>(ws + counters + ml);
  p.slab_base = static_cast<int>(slab_base);
  p.n_rows = n_rows; p.n_q = n_q; p.n_kv = n_kv; p.G = G; p.page_size = page_size;
  p.max_chunks = max_chunks_total;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  const CUtensorMap* map = static_cast<const CUtensorMap*>(kv_map);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool hi = G > 8;
#define VB_ATTN_CASE(DD, CC, SS)                                                  \
  if (head_dim == DD && chunk_tokens == CC)                                       \
    return hi ? launch_attn<DD, CC, SS, true>(p, map, grid_ctas, st)              \
              : launch_attn<DD, CC, SS, false>(p, map, grid_ctas, st);
  VB_ATTN_CASE(128, 64, 3)
  VB_ATTN_CASE(128, 32, 4)
  VB_ATTN_CASE(128, 16, 4)
  VB_ATTN_CASE(64, 64, 4)
  VB_ATTN_CASE(64, 32, 4)
  VB_ATTN_CASE(64, 16, 4)
#undef VB_ATTN_CASE
  vb::set_error("vb_paged_attn: no kernel instance");
  return -1;
}

}  // extern "C"
