// Paged GQA attention for decode and (ragged, causal) prefill rows.
// Replaces FlashInferDecodeWrapper.run / FlashInferPrefillWrapper.run (vox_serve/flashinfer_utils.py:132,
// 228-230) behind the plan produced by vb_plan_rows.
//
// HBM-bound by design (decode reads every K/V byte once, ~6 flop per byte).  Structure:
//  * a TILE is TOK consecutive tokens of one query row's KV with ALL kv heads: one contiguous
//    TOK * n_kv * D * 2-byte run of the page for K and one for V (64 KiB together for Orpheus), so HBM sees
//    long sequential bursts; tiles are linearised (row, tile) and cut into gridDim.x equal contiguous ranges
//    ("stream-K" over KV): every CTA streams the same number of tiles whatever the mix of sequence lengths;
//  * one PRODUCER warp per CTA walks its range -- 32 tiles of metadata resolved at a time, one lane each, from
//    shared-memory copies of the plan -- and feeds a ring of HALF-stages (the K rows of a tile, then its V rows:
//    32 KiB + skew each) behind full/empty mbarriers with the TMA unit's LINEAR bulk copies
//    (cp.async.bulk.shared.global, L2 evict_first: a layer's KV is not read again before the next step): one
//    copy per token row (n_kv * D * 2 bytes, all heads), issued by the warp's lanes in parallel.  Three
//    half-stages (~100 KiB) keep a CTA below half an SM's shared memory, so the CTA of a NEIGHBOURING kernel of
//    the programmatic-dependent-launch chain (the next layer's attention in a back-to-back run; the O projection's
//    weight prefetch or the QKV tail in a decode step) is resident beside it: its launch latency, plan look-up
//    and the KV of tokens that were already in the cache stream in while this kernel is still running.  Rows land with a 16-byte
//    skew (row stride n_kv * D * 2 + 16) so that the consumers' ldmatrix reads of 8 consecutive tokens are
//    bank-conflict free without a swizzle.  (A tiled tensor-map box must keep a 128-byte inner extent under the
//    128B swizzle; measured, the TMA unit then spends ~19 cycles per 128-byte row and caps the kernel below
//    2 TB/s -- 2 KiB linear rows are 16x fewer requests.)  The row's Q (all heads) is bulk-copied into a double
//    buffer at each segment start;
//  * consumer warp w owns kv head (w % n_kv) and 16-token slab (w / n_kv) of every tile and keeps PRIVATE
//    online-softmax state across tiles.  Tokens are the MMA M dimension: S^T = K Q^T (mma.sync m16n8k16, the G
//    grouped query heads are the N columns), P^T is transposed in registers with movmatrix, O^T += V^T P^T;
//    exp2 via ex2.approx, P rounded to bf16 and the denominator summed from the rounded P (FlashInfer FA2
//    numerics).  No block-wide synchronisation per tile -- only at the end of a segment;
//  * a row that spans several CTAs leaves one partial per CTA; the last CTA to arrive (one atomic per segment)
//    merges them in CTA order (deterministic) and restores the arrival counter to 0.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 1
#include "common.cuh"

namespace vb {

constexpr int ATTN_MAX_STAGES = 8;     // half-stages (K rows or V rows of one tile)
constexpr int ATTN_QBUF = 1;           // consumers copy their Q fragments to registers at the segment start

struct AttnParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* q;
  const int32_t* row_kvlen;
  const int32_t* row_chunk_start;
  const int32_t* row_pagebase;   // first entry of the row's request in kv_indices
  const int32_t* row_old;        // tokens of the row's request that were in the cache before this step
  const int32_t* kv_indices;
  const __nv_bfloat16* kv;       // whole cache [slabs][2][page_size][n_kv][D]
  int32_t* counters;             // [n_rows][n_kv]
  float* part_ml;                // [grid * 2][n_q][2]
  float* part_o;                 // [grid * 2][n_q][D]
  int slab_base;
  int n_rows, n_q, n_kv, G, page_size, tok, stages;
  int out_xt_tile;               // 0: out is [row][n_q][D]; else the tiled XT(out_xt_tile) layout of [row][n_q * D]
  float scale_log2;
};

struct TileMeta {   // 32 bytes, one per ring stage
  int row, token0, kvlen, flags;   // flags: 1 = segment start, 2 = segment end, 4 = segment covers the whole row
  int slot, cta_a, n_parts, qslot;
};

// output element (row, head hq, dim d): row-major or the tiled activation layout the O projection streams
__device__ __forceinline__ size_t attn_out_index(const AttnParams& p, int row, int hq, int d, int D) {
  if (p.out_xt_tile == 0) return (static_cast<size_t>(row) * p.n_q + hq) * D + d;
  return xt_index(row, hq * D + d, p.out_xt_tile, (p.n_q * D + 63) >> 6);
}

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                   uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// shared-memory carve-up (host and device agree through this one function)
struct AttnLayout {
  int stage_bytes, off_q, off_meta, off_bar, off_wml, off_flag, off_ored, off_plan, total;
  __host__ __device__ AttnLayout(int D, int tok, int n_kv, int n_q, int G, int stages, int n_rows) {
    stage_bytes = tok * (n_kv * D * 2 + 16);                    // one half-stage: K (or V) rows, each skewed by 16 bytes
    off_q = stages * stage_bytes;
    off_meta = off_q + ATTN_QBUF * n_q * D * 2;
    off_bar = off_meta + ATTN_MAX_STAGES * 32;                  // full[8], empty[8], qfull, qempty
    off_wml = off_bar + (2 * ATTN_MAX_STAGES + 2 * ATTN_QBUF) * 8;   // float[n_warps][16][2]
    const int n_warps = n_kv * (tok / 16);
    // (the block-level slab merge only exists when a tile holds several 16-token slabs, i.e. n_kv < 8)
    const int slabbed = tok > 16 ? 1 : 0;
    off_flag = off_wml + slabbed * n_warps * 16 * 2 * 4;
    off_ored = (off_flag + 16 + 127) / 128 * 128;               // float[n_warps][G][D]
    off_plan = off_ored + slabbed * n_warps * G * D * 4;        // int[4][n_rows + 1]
    total = off_plan + 4 * (n_rows + 1) * 4 + 128;              // + alignment slack
  }
};

// MINB = CTAs per SM the register allocation is sized for: 1 -> up to 168 registers, no spills (deep ring, the SM to
// itself); 2 -> 96 registers (some spills) so that a second CTA -- this kernel's or a neighbouring kernel's -- fits.
template <int D, bool HI, int MINB>
__global__ void __launch_bounds__(288, MINB) paged_attn_kernel(const AttnParams p) {
  constexpr int KS = D / 16;            // k-steps of K Q^T
  constexpr int MT = D / 16;            // m-tiles of V^T P^T
  constexpr int NT = HI ? 2 : 1;        // 8-head column tiles of the GQA group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int G = p.G, TOK = p.tok, STAGES = p.stages;
  const AttnLayout L(D, TOK, p.n_kv, p.n_q, G, STAGES, p.n_rows);
  const int NW = p.n_kv * (TOK / 16);   // consumer warps
  const int NC = NW * 32;
  const int ROWB = p.n_kv * D * 2;      // bytes of one token row (all heads) in the cache
  const int RS = ROWB + 16;             // its stride in shared memory (16-byte skew)
  TileMeta* meta = reinterpret_cast<TileMeta*>(smem + L.off_meta);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* empty = full + ATTN_MAX_STAGES;
  uint64_t* qfull = empty + ATTN_MAX_STAGES;
  uint64_t* qempty = qfull + ATTN_QBUF;
  float* wml = reinterpret_cast<float*>(smem + L.off_wml);
  int* flag = reinterpret_cast<int*>(smem + L.off_flag);
  float* ored = reinterpret_cast<float*>(smem + L.off_ored);
  int* s_cstart = reinterpret_cast<int*>(smem + L.off_plan);   // [n_rows + 1]
  int* s_kvlen = s_cstart + p.n_rows + 1;                      // [n_rows]
  int* s_pbase = s_kvlen + p.n_rows + 1;                       // [n_rows]
  int* s_old = s_pbase + p.n_rows + 1;                         // [n_rows]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tr = (tid == 0 && trace_block0()) ? trace_begin(2) : -1;
  unsigned long long* fine = (lane == 0) ? trace_fine_base() : nullptr;    // (dev) per-tile marks, block 0

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    for (int s = 0; s < ATTN_QBUF; ++s) {
      mbar_init(&qfull[s], 1);
      mbar_init(&qempty[s], NW);
    }
    fence_barrier_init();
  }
  for (int i = tid; i <= p.n_rows; i += blockDim.x) {
    s_cstart[i] = p.row_chunk_start[i];
    if (i < p.n_rows) {
      s_kvlen[i] = p.row_kvlen[i];
      s_pbase[i] = p.row_pagebase[i];
      s_old[i] = p.row_old[i];
    }
  }
  __syncthreads();
  const int total = s_cstart[p.n_rows];
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int c0 = min(total, static_cast<int>(blockIdx.x) * per);
  const int c1 = min(total, c0 + per);

  if (warp == NW) {
    // ===================== producer warp (stays converged; lane 0 issues) =====================
    // PDL: the plan (>= 2 kernels upstream) and the KV of earlier steps are final when this kernel starts, so
    // tiles that hold only old tokens are requested before the grid dependency resolves; the first tile that
    // needs the predecessor's output (the new tokens' K/V, or any Q) waits for it.
    int stage = 0, nseg = 0;
    uint32_t ph = 0;
    bool dep_ok = false;
    int issued = 0, n_pend = 0, pend_row0 = 0, pend_row1 = 0;
    const uint64_t pol_kv = policy_evict_first();
    auto issue_q = [&](int row, int seg) {
      mbar_wait(&qempty[0], (seg & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qfull[0], p.n_q * D * 2);
        bulk_copy_g2s(smem + L.off_q, p.q + static_cast<size_t>(row) * p.n_q * D, p.n_q * D * 2, &qfull[0]);
      }
      __syncwarp();
    };
    for (int base = c0; base < c1; base += 32) {
      // ---- each lane resolves one tile of the next 32 ----
      const int Lx = base + lane;
      int m_row = 0, m_tok = 0, m_kvlen = 0, m_flags = 0, m_slot = 0, m_ctaa = 0, m_parts = 1, m_page = 0, m_old = 0;
      if (Lx < c1) {
        int lo = 0, hi = p.n_rows;      // largest row with cstart[row] <= Lx (rows with no tiles are skipped)
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_cstart[mid] <= Lx) lo = mid; else hi = mid;
        }
        const int row = lo, Ls = s_cstart[row], Le = s_cstart[row + 1];
        const int seg0 = max(Ls, c0), seg1 = min(Le, c1);
        m_row = row; m_tok = (Lx - Ls) * TOK; m_kvlen = s_kvlen[row]; m_old = s_old[row];
        m_flags = (Lx == seg0 ? 1 : 0) | (Lx == seg1 - 1 ? 2 : 0) | ((Ls >= c0 && Le <= c1) ? 4 : 0);
        m_slot = (Ls > c0) ? 1 : 0;
        m_ctaa = Ls / per;
        m_parts = (Le - 1) / per - m_ctaa + 1;
        m_page = __ldg(&p.kv_indices[s_pbase[row] + m_tok / p.page_size]);
      }
      const int n_here = min(32, c1 - base);
      for (int k = 0; k < n_here; ++k) {
        TileMeta m;
        m.row = __shfl_sync(0xffffffffu, m_row, k);
        m.token0 = __shfl_sync(0xffffffffu, m_tok, k);
        m.kvlen = __shfl_sync(0xffffffffu, m_kvlen, k);
        m.flags = __shfl_sync(0xffffffffu, m_flags, k);
        m.slot = __shfl_sync(0xffffffffu, m_slot, k);
        m.cta_a = __shfl_sync(0xffffffffu, m_ctaa, k);
        m.n_parts = __shfl_sync(0xffffffffu, m_parts, k);
        const int page = __shfl_sync(0xffffffffu, m_page, k);
        const int old = __shfl_sync(0xffffffffu, m_old, k);
        const bool seg_start = (m.flags & 1) != 0;
        auto resolve_dep = [&]() {
          pdl_wait();
          pdl_trigger();
          dep_ok = true;
          for (int i = 0; i < n_pend; ++i) issue_q(i == 0 ? pend_row0 : pend_row1, i);
          n_pend = 0;
        };
        if (!dep_ok && (m.token0 + TOK > old || (seg_start && n_pend >= ATTN_QBUF))) resolve_dep();
        if (seg_start) {
          // the row's Q (all heads) goes into the double buffer; its slot was freed two segments ago
          m.qslot = nseg;
          if (dep_ok) {
            issue_q(m.row, nseg);
          } else {
            if (n_pend == 0) pend_row0 = m.row; else pend_row1 = m.row;
            ++n_pend;
          }
          ++nseg;
        } else {
          m.qslot = nseg - 1;
        }
        {
          const int slot0 = m.token0 % p.page_size;
          const size_t page_elems = static_cast<size_t>(p.page_size) * p.n_kv * D;
          const __nv_bfloat16* src0 = p.kv + static_cast<size_t>(p.slab_base + page) * 2 * page_elems +
                                      static_cast<size_t>(slot0) * p.n_kv * D;
          // two half-stages per tile: the K rows, then the V rows (TOK row copies each, spread over the lanes)
#pragma unroll
          for (int kvsel = 0; kvsel < 2; ++kvsel) {
            // a half-stage that is being re-used is freed by the consumers, who need Q, i.e. the predecessor kernel
            if (!dep_ok && 2 * issued + kvsel >= STAGES) resolve_dep();
            mbar_wait(&empty[stage], ph ^ 1);
            if (kvsel == 0) trace_fine(fine, 0, base + k - c0);
            if (lane == 0) {
              if (kvsel == 0) meta[stage] = m;
              mbar_arrive_expect_tx(&full[stage], TOK * ROWB);
            }
            __syncwarp();
            uint8_t* dst = smem + stage * L.stage_bytes;
            for (int t = lane; t < TOK; t += 32)
              bulk_copy_g2s_hint(dst + t * RS, src0 + kvsel * page_elems + static_cast<size_t>(t) * p.n_kv * D, ROWB,
                                 &full[stage], pol_kv);
            __syncwarp();
            if (++stage == STAGES) { stage = 0; ph ^= 1; }
          }
          ++issued;
        }
      }
    }
    if (!dep_ok) {
      pdl_wait();
      pdl_trigger();
      for (int i = 0; i < n_pend; ++i) issue_q(i == 0 ? pend_row0 : pend_row1, i);
    }
    return;
  }
  if (warp > NW) return;

  // ===================== consumer warps =====================
  // padded / empty rows produce zeros
  for (int row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
    if (s_kvlen[row] == 0) {
      for (int i = tid; i < p.n_q * D / 2; i += NC)
        *reinterpret_cast<uint32_t*>(p.out + attn_out_index(p, row, (2 * i) / D, (2 * i) % D, D)) = 0u;
    }
  }
  auto csync = [NC]() { asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory"); };

  const int g = lane >> 2;         // C-fragment row (token g, g + 8) / B-fragment column (head g)
  const int qd = lane & 3;         // C-fragment column pair (heads 2qd, 2qd + 1)
  const int head = warp % p.n_kv;  // kv head of this warp
  const int slab = warp / p.n_kv;  // 16-token slab of the tile
  const int hoff = head * D * 2;   // byte offset of this warp's head inside a token row
  const int mi = lane >> 3, r8 = lane & 7;

  uint32_t qb[NT][KS][2];
  float O[MT][NT][4];
  float mrun[NT][2], lrun[NT][2];   // per head column (2qd, 2qd+1); l: this lane's share until the segment end

  int stage = 0;
  uint32_t ph = 0;
  for (int Lx = c0; Lx < c1; ++Lx) {
    mbar_wait(&full[stage], ph);
    if (warp == 0) trace_fine(fine, 1, Lx - c0);
    if (tr >= 0 && Lx == c0) trace_mark(22);          // first tile landed
    const TileMeta m = meta[stage];
    const uint32_t kbase = smem_u32(smem + stage * L.stage_bytes);

    if (m.flags & 1) {
      // ---- new segment: reset the running state, pull this head group's Q fragments into registers ----
      mbar_wait(&qfull[0], m.qslot & 1);
      const uint8_t* qsm = smem + L.off_q + head * G * D * 2;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int hq = nt * 8 + g;       // B fragment: column n = head within the group
        const bool ok = hq < G;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int d0 = (ks * 16 + qd * 2) * 2;
          qb[nt][ks][0] = ok ? *reinterpret_cast<const uint32_t*>(qsm + hq * (D * 2) + d0) : 0u;
          qb[nt][ks][1] = ok ? *reinterpret_cast<const uint32_t*>(qsm + hq * (D * 2) + d0 + 16) : 0u;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&qempty[0]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) O[mt][nt][j] = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mrun[nt][0] = mrun[nt][1] = -INFINITY;
        lrun[nt][0] = lrun[nt][1] = 0.f;
      }
    }

    // ---- S^T = K Q^T : this warp's 16 tokens x the group's heads ----
    float S[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
    {
      const int trow = slab * 16 + r8 + (mi & 1) * 8;
      const uint32_t rbase = kbase + trow * RS + hoff + (mi >> 1) * 16;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a, rbase + ks * 32);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(S[nt], a, qb[nt][ks][0], qb[nt][ks][1]);
      }
    }
    // this warp is done with the K half-stage
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == STAGES) { stage = 0; ph ^= 1; }
    // ---- mask, running max, P = exp2(S - m) rounded to bf16, sums of the rounded values ----
    const int tok_lo = m.token0 + slab * 16 + g;
    const bool in_lo = tok_lo < m.kvlen, in_hi = tok_lo + 8 < m.kvlen;
    uint32_t pb[NT][2];
    float alpha[NT][2];
    bool rescale = false;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float s0 = in_lo ? S[nt][0] * p.scale_log2 : -INFINITY, s1 = in_lo ? S[nt][1] * p.scale_log2 : -INFINITY;
      float s2 = in_hi ? S[nt][2] * p.scale_log2 : -INFINITY, s3 = in_hi ? S[nt][3] * p.scale_log2 : -INFINITY;
      float mx0 = fmaxf(s0, s2), mx1 = fmaxf(s1, s3);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
      }
      const float n0 = fmaxf(mrun[nt][0], mx0), n1 = fmaxf(mrun[nt][1], mx1);
      const float u0 = (n0 == -INFINITY) ? 0.f : n0, u1 = (n1 == -INFINITY) ? 0.f : n1;   // all masked so far
      alpha[nt][0] = ex2_approx(mrun[nt][0] - u0);                                         // ex2(-inf) = 0
      alpha[nt][1] = ex2_approx(mrun[nt][1] - u1);
      rescale |= (n0 != mrun[nt][0]) | (n1 != mrun[nt][1]);
      mrun[nt][0] = n0; mrun[nt][1] = n1;
      const float p0 = round_bf16(ex2_approx(s0 - u0)), p1 = round_bf16(ex2_approx(s1 - u1));
      const float p2 = round_bf16(ex2_approx(s2 - u0)), p3 = round_bf16(ex2_approx(s3 - u1));
      lrun[nt][0] = lrun[nt][0] * alpha[nt][0] + (p0 + p2);
      lrun[nt][1] = lrun[nt][1] * alpha[nt][1] + (p1 + p3);
      // C layout (token g | g+8, heads 2qd, 2qd+1) -> B layout (tokens 2qd, 2qd+1 | +8, head g)
      pb[nt][0] = movmatrix_trans(pack_bf16(p0, p1));
      pb[nt][1] = movmatrix_trans(pack_bf16(p2, p3));
    }
    rescale = __any_sync(0xffffffffu, rescale);
    // ---- O^T = O^T * alpha + V^T P^T ----
    mbar_wait(&full[stage], ph);
    {
      const uint32_t vbase = smem_u32(smem + stage * L.stage_bytes);
      const int trow = slab * 16 + r8 + (mi >> 1) * 8;
      const uint32_t rbase = vbase + trow * RS + hoff + (mi & 1) * 16;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, rbase + mt * 32);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float(&o)[4] = O[mt][nt];
          if (rescale) {
            o[0] *= alpha[nt][0]; o[1] *= alpha[nt][1];
            o[2] *= alpha[nt][0]; o[3] *= alpha[nt][1];
          }
          mma_bf16_16816(o, a, pb[nt][0], pb[nt][1]);
        }
      }
    }
    // this warp is done with the V half-stage: hand the slot back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (warp == 0) trace_fine(fine, 2, Lx - c0);
    if (++stage == STAGES) { stage = 0; ph ^= 1; }

    if (!(m.flags & 2)) continue;
    if (tr >= 0 && Lx == c1 - 1) trace_mark(23);      // last tile consumed: what follows is the segment tail / merge

    // ================= end of segment =================
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        lrun[nt][0] += __shfl_xor_sync(0xffffffffu, lrun[nt][0], o);
        lrun[nt][1] += __shfl_xor_sync(0xffffffffu, lrun[nt][1], o);
      }
    if (TOK == 16) {
      // ---- one slab per tile: this warp owns its kv head outright, no block-level step at all ----
      const bool complete = (m.flags & 4) != 0;
      const size_t pslot = static_cast<size_t>(blockIdx.x) * 2 + m.slot;
      const int hq0 = head * G;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int h0 = nt * 8 + qd * 2;
        if (complete) {
          const float i0 = 1.f / lrun[nt][0], i1 = 1.f / lrun[nt][1];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int d = mt * 16 + g;
            if (h0 < G) {
              p.out[attn_out_index(p, m.row, hq0 + h0, d, D)] = __float2bfloat16_rn(O[mt][nt][0] * i0);
              p.out[attn_out_index(p, m.row, hq0 + h0, d + 8, D)] = __float2bfloat16_rn(O[mt][nt][2] * i0);
            }
            if (h0 + 1 < G) {
              p.out[attn_out_index(p, m.row, hq0 + h0 + 1, d, D)] = __float2bfloat16_rn(O[mt][nt][1] * i1);
              p.out[attn_out_index(p, m.row, hq0 + h0 + 1, d + 8, D)] = __float2bfloat16_rn(O[mt][nt][3] * i1);
            }
          }
        } else {
          float* o0 = p.part_o + (pslot * p.n_q + hq0 + h0) * D;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int d = mt * 16 + g;
            if (h0 < G) { o0[d] = O[mt][nt][0]; o0[d + 8] = O[mt][nt][2]; }
            if (h0 + 1 < G) { o0[D + d] = O[mt][nt][1]; o0[D + d + 8] = O[mt][nt][3]; }
          }
          if (g == 0) {
            float* ml = p.part_ml + (pslot * p.n_q + hq0 + h0) * 2;
            if (h0 < G) { ml[0] = mrun[nt][0]; ml[1] = lrun[nt][0]; }
            if (h0 + 1 < G) { ml[2] = mrun[nt][1]; ml[3] = lrun[nt][1]; }
          }
        }
      }
      if (!complete) {
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          __threadfence();           // orders the whole warp's partial (fences are cumulative) before the arrival
          const int old = atomicAdd(&p.counters[m.row * p.n_kv + head], 1);
          last = (old == m.n_parts - 1);
          if (last) {
            p.counters[m.row * p.n_kv + head] = 0;
            __threadfence();
          }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          // merge this head group's partials in CTA order (slot 1 only if the row starts inside CTA cta_a's range).
          // All CTAs reach this point at about the same time, so the merge is pure tail latency: every load of a
          // step is issued before the first one is used (two L2 round trips per four parts, not two per part).
          const int Ls = s_cstart[m.row];
          const int n = m.n_parts;
          auto pslot_of = [&](int c) {
            return static_cast<size_t>(m.cta_a + c) * 2 + ((c == 0 && Ls > m.cta_a * per) ? 1 : 0);
          };
          for (int g0 = 0; g0 < G; g0 += 4) {
            float Mx[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) Mx[hh] = -INFINITY;
            for (int cb = 0; cb < n; cb += 32) {
              const int c = cb + lane;
              float mm[4];
#pragma unroll
              for (int hh = 0; hh < 4; ++hh)
                mm[hh] = (c < n && g0 + hh < G) ? __ldcg(&p.part_ml[(pslot_of(c) * p.n_q + hq0 + g0 + hh) * 2]) : -INFINITY;
#pragma unroll
              for (int hh = 0; hh < 4; ++hh) Mx[hh] = fmaxf(Mx[hh], warp_max(mm[hh]));
            }
            float acc[4][D / 32], den[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              den[hh] = 0.f;
#pragma unroll
              for (int j = 0; j < D / 32; ++j) acc[hh][j] = 0.f;
            }
            // (two parts per round keeps the kernel under the 112 registers that let two CTAs share an SM)
            constexpr int MP = 2;
            for (int cb = 0; cb < n; cb += MP) {
              float mm[MP][4], ll[MP][4], v[MP][4][D / 32];
#pragma unroll
              for (int k = 0; k < MP; ++k) {
                const bool okc = cb + k < n;
                const size_t ps = pslot_of(okc ? cb + k : 0);
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const bool ok = okc && g0 + hh < G;
                  const size_t hb = ps * p.n_q + hq0 + g0 + hh;
                  mm[k][hh] = ok ? __ldcg(&p.part_ml[hb * 2]) : -INFINITY;
                  ll[k][hh] = ok ? __ldcg(&p.part_ml[hb * 2 + 1]) : 0.f;
#pragma unroll
                  for (int j = 0; j < D / 32; ++j) v[k][hh][j] = ok ? __ldcg(&p.part_o[hb * D + j * 32 + lane]) : 0.f;
                }
              }
#pragma unroll
              for (int k = 0; k < MP; ++k)
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const float sc = (mm[k][hh] == -INFINITY) ? 0.f : ex2_approx(mm[k][hh] - Mx[hh]);
                  den[hh] += sc * ll[k][hh];
#pragma unroll
                  for (int j = 0; j < D / 32; ++j) acc[hh][j] += sc * v[k][hh][j];
                }
            }
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              if (g0 + hh < G) {
                const float inv = 1.f / den[hh];
#pragma unroll
                for (int j = 0; j < D / 32; ++j)
                  p.out[attn_out_index(p, m.row, hq0 + g0 + hh, j * 32 + lane, D)] = __float2bfloat16_rn(acc[hh][j] * inv);
              }
            }
          }
        }
      }
      continue;
    }

    // ---- several slabs per tile (n_kv < 8): merge the slabs of each head through shared memory ----
    // stage (m, l, O) of this warp's head group in shared memory: wml[warp][head 0..15][2], ored[warp][G][D]
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int h0 = nt * 8 + qd * 2;
      if (g == 0) {
        wml[(warp * 16 + h0) * 2 + 0] = mrun[nt][0];
        wml[(warp * 16 + h0) * 2 + 1] = lrun[nt][0];
        wml[(warp * 16 + h0 + 1) * 2 + 0] = mrun[nt][1];
        wml[(warp * 16 + h0 + 1) * 2 + 1] = lrun[nt][1];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int d = mt * 16 + g;
        if (h0 < G) {
          ored[(warp * G + h0) * D + d] = O[mt][nt][0];
          ored[(warp * G + h0) * D + d + 8] = O[mt][nt][2];
        }
        if (h0 + 1 < G) {
          ored[(warp * G + h0 + 1) * D + d] = O[mt][nt][1];
          ored[(warp * G + h0 + 1) * D + d + 8] = O[mt][nt][3];
        }
      }
    }
    csync();

    const bool complete = (m.flags & 4) != 0;
    const size_t pslot = static_cast<size_t>(blockIdx.x) * 2 + m.slot;
    const int slabs = TOK / 16;
    for (int i = tid; i < p.n_q * D; i += NC) {
      const int hq = i / D, d = i - hq * D;
      const int hk = hq / G, gg = hq - hk * G;
      float Mg = -INFINITY;
      for (int s = 0; s < slabs; ++s) Mg = fmaxf(Mg, wml[((s * p.n_kv + hk) * 16 + gg) * 2]);
      float o = 0.f, l = 0.f;
      for (int s = 0; s < slabs; ++s) {
        const int w = s * p.n_kv + hk;
        const float mw = wml[(w * 16 + gg) * 2];
        const float sc = (mw == -INFINITY) ? 0.f : ex2_approx(mw - Mg);
        o += ored[(w * G + gg) * D + d] * sc;
        l += wml[(w * 16 + gg) * 2 + 1] * sc;
      }
      if (complete) {
        p.out[attn_out_index(p, m.row, hq, d, D)] = __float2bfloat16_rn(o / l);
      } else {
        p.part_o[pslot * p.n_q * D + i] = o;
        if (d == 0) {
          p.part_ml[(pslot * p.n_q + hq) * 2 + 0] = Mg;
          p.part_ml[(pslot * p.n_q + hq) * 2 + 1] = l;
        }
      }
    }
    if (!complete) {
      csync();                       // every partial of this CTA is written ...
      if (tid == 0) {
        __threadfence();             // ... and ordered before the arrival (fences are cumulative)
        const int old = atomicAdd(&p.counters[m.row * p.n_kv], 1);
        const int last = (old == m.n_parts - 1);
        if (last) {
          p.counters[m.row * p.n_kv] = 0;
          __threadfence();
        }
        *flag = last;
      }
      csync();
      if (*flag) {
        // partial of CTA c for this row: slot 1 only if the row starts inside CTA cta_a's range
        const int Ls = s_cstart[m.row];
        for (int i = tid; i < p.n_q * D; i += NC) {
          const int hq = i / D;
          float Mx = -INFINITY;
          for (int c = 0; c < m.n_parts; ++c) {
            const size_t ps = static_cast<size_t>(m.cta_a + c) * 2 + ((c == 0 && Ls > m.cta_a * per) ? 1 : 0);
            Mx = fmaxf(Mx, __ldcg(&p.part_ml[(ps * p.n_q + hq) * 2]));
          }
          float num = 0.f, den = 0.f;
          for (int c = 0; c < m.n_parts; ++c) {
            const size_t ps = static_cast<size_t>(m.cta_a + c) * 2 + ((c == 0 && Ls > m.cta_a * per) ? 1 : 0);
            const float mm = __ldcg(&p.part_ml[(ps * p.n_q + hq) * 2]);
            const float sc = (mm == -INFINITY) ? 0.f : ex2_approx(mm - Mx);
            den += sc * __ldcg(&p.part_ml[(ps * p.n_q + hq) * 2 + 1]);
            num += sc * __ldcg(&p.part_o[ps * p.n_q * D + i]);
          }
          p.out[attn_out_index(p, m.row, hq, i - hq * D, D)] = __float2bfloat16_rn(num / den);
        }
      }
    }
    csync();   // wml / ored / flag reusable
  }
  trace_end(tr);
}

template <int D, bool HI, int MINB>
static int launch_attn(const AttnParams& p, int grid, cudaStream_t stream) {
  const AttnLayout L(D, p.tok, p.n_kv, p.n_q, p.G, p.stages, p.n_rows);
  VB_CHECK_ARG(L.total <= VB_MAX_DYN_SMEM, "vb_paged_attn: %d rows / %d-token tiles need %d bytes of shared memory",
               p.n_rows, p.tok, L.total);
  auto kern = paged_attn_kernel<D, HI, MINB>;
  VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  const int threads = (p.n_kv * (p.tok / 16) + 1) * 32;
  VB_LAUNCH_PDL(kern, grid, threads, L.total, stream, p);
  return 0;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_attn_tile_tokens(int page_size, int n_kv) {
  if (page_size < 16 || page_size % 16 != 0 || n_kv < 1 || n_kv > 8) return -1;
  int tok = 16 * (8 / n_kv);            // 8 consumer warps: one per (kv head, 16-token slab)
  while (tok > 16 && (tok > page_size || page_size % tok != 0)) tok -= 16;
  return tok;
}

size_t vb_paged_attn_workspace_bytes(int max_rows, int max_grid_ctas, int n_q, int n_kv, int head_dim) {
  const size_t ml = (static_cast<size_t>(max_grid_ctas) * 2 * n_q * 2 * sizeof(float) + 255) / 256 * 256;
  const size_t po = (static_cast<size_t>(max_grid_ctas) * 2 * n_q * head_dim * sizeof(float) + 255) / 256 * 256;
  const size_t counters = (static_cast<size_t>(max_rows) * n_kv * sizeof(int32_t) + 255) / 256 * 256;
  return ml + po + counters;
}

int vb_paged_attn(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                  const int32_t* d_row_kvlen, const int32_t* d_row_chunk_start, const int32_t* d_row_pagebase,
                  const int32_t* d_row_old, const int32_t* d_kv_indices, int n_rows, int n_q, int n_kv, int head_dim, int page_size,
                  int chunk_tokens, float sm_scale, void* d_workspace, size_t workspace_bytes, int grid_ctas,
                  int ws_grid_ctas, int out_xt_tile, void* stream) {
  VB_CHECK_ARG(d_out && d_q && d_kv && d_row_kvlen && d_row_chunk_start && d_row_pagebase && d_row_old && d_kv_indices &&
                   d_workspace,
               "vb_paged_attn: null pointer");
  VB_CHECK_ARG(n_kv > 0 && n_kv <= 8 && n_q % n_kv == 0 && n_q / n_kv <= 16,
               "vb_paged_attn: %d query / %d kv heads unsupported (kv heads <= 8, group <= 16)", n_q, n_kv);
  VB_CHECK_ARG(head_dim == 64 || head_dim == 128, "vb_paged_attn: head_dim %d unsupported (64, 128)", head_dim);
  VB_CHECK_ARG(chunk_tokens == vb_attn_tile_tokens(page_size, n_kv),
               "vb_paged_attn: chunk_tokens %d != vb_attn_tile_tokens(%d, %d)", chunk_tokens, page_size, n_kv);
  VB_CHECK_ARG(grid_ctas > 0 && grid_ctas <= ws_grid_ctas, "vb_paged_attn: grid_ctas %d outside (0, %d]", grid_ctas,
               ws_grid_ctas);
  VB_CHECK_ARG(workspace_bytes >= vb_paged_attn_workspace_bytes(n_rows, ws_grid_ctas, n_q, n_kv, head_dim),
               "vb_paged_attn: workspace too small for %d rows x %d CTAs", n_rows, ws_grid_ctas);
  if (n_rows <= 0) return 0;
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.row_kvlen = d_row_kvlen;
  p.row_chunk_start = d_row_chunk_start;
  p.row_pagebase = d_row_pagebase;
  p.row_old = d_row_old;
  p.kv_indices = d_kv_indices;
  p.kv = static_cast<const __nv_bfloat16*>(d_kv);
  // layout fixed by ws_grid_ctas (what the caller sized the buffer for), not by this launch: counters come last
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const size_t ml = (static_cast<size_t>(ws_grid_ctas) * 2 * n_q * 2 * sizeof(float) + 255) / 256 * 256;
  const size_t po = (static_cast<size_t>(ws_grid_ctas) * 2 * n_q * head_dim * sizeof(float) + 255) / 256 * 256;
  p.part_ml = reinterpret_cast<float*>(ws);
  p.part_o = reinterpret_cast<float*>(ws + ml);
  p.counters = reinterpret_cast<int32_t*>(ws + ml + po);
  p.slab_base = static_cast<int>(slab_base);
  p.n_rows = n_rows; p.n_q = n_q; p.n_kv = n_kv; p.G = n_q / n_kv; p.page_size = page_size;
  p.tok = chunk_tokens;
  p.out_xt_tile = out_xt_tile;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  // deepest ring of half-stages that fits the budget next to the fixed buffers.  Measured on B200 (28 launches back to
  // back, kv 200 / 500 / 900): a bulk-copied tile lands ~3 us after it was requested once HBM is loaded, so the stream
  // needs ~190 KiB in flight per SM; three half-stages (two CTAs per SM, VB_ATTN_SMEM_KB=111) reach only 27 GB/s per
  // SM and lose more in the stream than co-residency with the neighbouring kernel gains (42 vs 20 us per launch).
  // Default: the whole SM, six half-stages.
  static int budget = -1;
  if (budget < 0) {
    const char* e = getenv("VB_ATTN_SMEM_KB");
    const int kb = e ? atoi(e) : 224;
    budget = (kb >= 64 && kb <= 227 ? kb : 224) * 1024;
  }
  int stages = ATTN_MAX_STAGES;
  while (stages > 2 && AttnLayout(head_dim, p.tok, n_kv, n_q, p.G, stages, n_rows).total > budget) --stages;
  VB_CHECK_ARG(AttnLayout(head_dim, p.tok, n_kv, n_q, p.G, stages, n_rows).total <= VB_MAX_DYN_SMEM,
               "vb_paged_attn: %d rows need more shared memory than an SM has", n_rows);
  p.stages = stages;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool hi = p.G > 8;
  const bool two = 2 * (AttnLayout(head_dim, p.tok, n_kv, n_q, p.G, stages, n_rows).total + 1024) <= 232448;
  if (two) {
    if (head_dim == 128) return hi ? launch_attn<128, true, 2>(p, grid_ctas, st) : launch_attn<128, false, 2>(p, grid_ctas, st);
    return hi ? launch_attn<64, true, 2>(p, grid_ctas, st) : launch_attn<64, false, 2>(p, grid_ctas, st);
  }
  if (head_dim == 128) return hi ? launch_attn<128, true, 1>(p, grid_ctas, st) : launch_attn<128, false, 1>(p, grid_ctas, st);
  return hi ? launch_attn<64, true, 1>(p, grid_ctas, st) : launch_attn<64, false, 1>(p, grid_ctas, st);
}

}  // extern "C"

VB_DEFINE_TRACE_SETTER(attn)
