// Dense projections Y[T][N] = X[T][K] * W[N][K]^T for the LM decode / prefill step on tcgen05.
// Replaces the nn.Linear calls of vox_serve/model/orpheus.py:41-47, 68-79, 91-93, 110, 197, 219 and, in the
// fused decode modes, everything the reference runs between them (orpheus.py:81-151): the RMSNorm in front of
// the QKV / gate-up projections, RoPE + KV append behind the QKV projection, SiLU(gate)*up, and the split-K
// reduction + residual add behind the O / down projections.
//
// At decode time T <= 64, so the problem is weight streaming (HBM-bound).  A tile of `tile_rows` (<= 128,
// multiple of 8) weight rows is the UMMA "A" operand (M = 128: rows past tile_rows are stale shared memory whose
// accumulator lanes are never read), the token rows are the "B" operand (UMMA N = t_tile), both K-major bf16,
// fp32 accumulators in TMEM.  tile_rows is a launch parameter so that every projection can be cut into ~148
// equal CTAs whatever its N.  Warp roles: warp 0 = TMA producer (64-wide K blocks into a 128B-swizzled ring; the
// weight tiles of the first ring pass are requested BEFORE the programmatic-dependent-launch wait, i.e. while
// the previous kernel is still running -- the ring is sized so that two CTAs fit one SM, which is what lets a
// dependent kernel sit next to its predecessor), warp 1 = single-thread tcgen05.mma issuer, warps 2..5 =
// epilogue (tcgen05.ld, each warp owns its 32-lane TMEM quarter) and, in the norm-fused modes, the B-operand
// producers: they read the bf16 hidden state, apply rsqrt(mean square) * weight, round to bf16 exactly where
// flashinfer.norm.rmsnorm does, and write the swizzled K-major tile the MMA expects.
//
// Split-K (blockIdx.y) spreads skinny problems over all SMs.  mode 1 leaves fp32 partial planes for a separate
// consumer (prefill path); modes 3 / 4 finish in the kernel: every CTA parks its partial tile in an L2-resident
// workspace, the last CTA of a tile to arrive (one atomic per CTA) sums the partials in split order --
// deterministic -- and runs the epilogue.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 2
#include "common.cuh"

namespace vb {

constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_EPI_THREADS = 128;

enum GemmMode {
  GM_BF16 = 0,      // Y bf16 [T][ldy]
  GM_PARTIAL = 1,   // Y fp32 [split][T][ldy]
  GM_SILU = 2,      // tile = h gate rows then the h matching up rows; Y bf16 [T][ldy] = silu(gate) * up
  GM_RESID = 3,     // split-K reduce -> bf16 -> + residual -> bf16 hidden [T][N]; per-tile sums of squares
  GM_ROPE = 4,      // split-K reduce -> bf16 -> RoPE (q, k heads) -> q out / K,V scatter into the page
};

struct GemmParams {
  void* y;
  const uint8_t* w_tiles;   // weights re-tiled by vb_pack_weight_tiles: [n_tile][k_block][tile_rows x 128 B, swizzled]
  const uint8_t* x_tiles;   // activations in the tiled XT(t_tile) layout (common.cuh), or nullptr: x_map (row-major)
  int y_tiled;              // mode 2: write Y in the XT(t_tile) layout (it is the next projection's activation)
  int T, N, K, ldy, mode, split_k, t_tile, stages, tmem_cols, tile_rows, n_out, pad;
  // B operand = rmsnorm(x) * w formed in the kernel from the raw x tile TMA delivered (norm != 0)
  int norm;
  const float* n_ssq;       // [n_ssq_parts][T] partial sums of squares of the x rows
  const __nv_bfloat16* n_w;
  int n_ssq_parts;
  float n_eps;
  // mode 3
  const __nv_bfloat16* residual;
  float* ssq_out;           // [tiles][T]
  __nv_bfloat16* y_tiles;   // optional second copy of the new hidden rows in the XT(t_tile) layout (next GEMM's input)
  // mode 4 (tile = one head, tile_rows = head_dim)
  __nv_bfloat16* q_out;
  __nv_bfloat16* kv;        // layer cache [pages][2][page_size][n_kv][D]
  const float* rope_cs;     // [T][2][D] cos | sin
  const int32_t* row_page;
  const int32_t* row_slot;
  int n_q, n_kv, page_size;
  // mode 0: optional bias [N] added in fp32 before the single bf16 rounding (nn.Linear with bias)
  const __nv_bfloat16* bias;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_ptr, uint32_t cta_rank) {
  uint32_t laddr = smem_u32(local_ptr), raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(cta_rank));
  return raddr;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

constexpr int GEMM_MAX_SPLIT_CLUSTER = 8;
constexpr int GEMM_EPI_CHUNK = 8;     // tokens an epilogue pass of the reducing modes handles at once

__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_bf16_kernel(const GemmParams p,
                                                                    const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage = p.tile_rows * 128;
  const int b_stage = p.t_tile * 128;
  const int stage_bytes = a_stage + b_stage;
  // (pad: the MMA always reads 128 A rows = 16 KiB from a stage's base; with short tiles the last stage's read
  // runs past the ring, so the allocation continues that far)
  uint8_t* tail = smem + p.stages * stage_bytes + p.pad;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);          // TMA bytes of the stage have landed
  uint64_t* empty = full + p.stages;                           // the MMAs have read the stage
  uint64_t* cfull = empty + p.stages;                          // (norm) the B tile has been normalised in place
  uint64_t* tmem_full = cfull + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* ssq_sm = reinterpret_cast<float*>(tail + 512);        // [4][GEMM_EPI_CHUNK]
  __nv_bfloat16* w_sm = reinterpret_cast<__nv_bfloat16*>(tail + 1024);   // (norm) the norm weight vector [K]
  // after the main loop the ring is free: partial tile [t_tile][128] fp32, then the epilogue exchange buffer
  float* part_sm = reinterpret_cast<float*>(smem);
  float* xchg = reinterpret_cast<float*>(smem + p.t_tile * 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, split = blockIdx.y, t_blk = blockIdx.z;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int kb0 = static_cast<int>(static_cast<long long>(split) * num_kb / p.split_k);
  const int kb1 = static_cast<int>(static_cast<long long>(split + 1) * num_kb / p.split_k);
  const bool bnorm = p.norm != 0;
  const bool reducing = p.mode == GM_RESID || p.mode == GM_ROPE;
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(1, p.mode) : -1;
  // (dev) per-stage marks of the gate/up launches: producer and MMA threads of block 0
  unsigned long long* fine = (lane == 0 && warp <= 1 && p.mode == GM_SILU) ? trace_fine_base() : nullptr;

  if (warp == 0 && lane == 0) {
    if (!p.x_tiles) prefetch_tmap(&x_map);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&cfull[s], GEMM_EPI_THREADS / 32);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // operands of the reducing epilogues that do not depend on the accumulator: fetched early, used after the sum
  float pre_a[GEMM_EPI_CHUNK], pre_b[GEMM_EPI_CHUNK];
  int pre_pg[GEMM_EPI_CHUNK], pre_sl[GEMM_EPI_CHUNK];

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();   // weights are streamed once per step
      const uint64_t pol_x = policy_evict_last();    // activations are re-read by every CTA
      // The weights do not depend on the previous kernel: fill the whole ring with weight tiles while it is
      // still running, then wait for it and add the activation tiles to the same stages.
      // A (tile, k-block) weight tile is one contiguous, pre-swizzled a_stage-byte run in HBM and this CTA's
      // k-blocks follow each other: a single linear bulk copy per stage, sequential DRAM bursts.
      const uint8_t* wsrc = p.w_tiles + (static_cast<size_t>(n_tile) * num_kb + kb0) * a_stage;
      const int npre = min(p.stages, kb1 - kb0);
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(&full[i], stage_bytes);
        bulk_g2s_hint(smem + i * stage_bytes, wsrc + static_cast<size_t>(i) * a_stage, a_stage, &full[i], pol_w);
      }
      pdl_wait();
      pdl_trigger();
      trace_fine(fine, 1, 0);
      if (tr >= 0) trace_mark(20, p.mode);
      // activations: one linear bulk copy per stage from the tiled layout, else a tensor-map box (t_tile row
      // requests: ~0.3 us of TMA issue time per stage, which then bounds the whole stream at ~4 TB/s)
      const uint8_t* xsrc = p.x_tiles ? p.x_tiles + static_cast<size_t>(t_blk) * num_kb * b_stage : nullptr;
      for (int i = 0; i < npre; ++i) {
        if (xsrc)
          bulk_g2s_hint(smem + i * stage_bytes + a_stage, xsrc + static_cast<size_t>(kb0 + i) * b_stage, b_stage, &full[i],
                        pol_x);
        else
          tma_load_2d_hint(smem + i * stage_bytes + a_stage, &x_map, &full[i], (kb0 + i) * GEMM_BLOCK_K,
                           t_blk * p.t_tile, pol_x);
      }
      int s = npre == p.stages ? 0 : npre;
      uint32_t ph = npre == p.stages ? 1 : 0;
      for (int kb = kb0 + npre; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        trace_fine(fine, 0, kb - kb0);
        uint8_t* a = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        bulk_g2s_hint(a, wsrc + static_cast<size_t>(kb - kb0) * a_stage, a_stage, &full[s], pol_w);
        if (xsrc) bulk_g2s_hint(a + a_stage, xsrc + static_cast<size_t>(kb) * b_stage, b_stage, &full[s], pol_x);
        else tma_load_2d_hint(a + a_stage, &x_map, &full[s], kb * GEMM_BLOCK_K, t_blk * p.t_tile, pol_x);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, p.t_tile, 1u);
      uint64_t* ready = bnorm ? cfull : full;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&ready[s], ph);
        trace_fine(fine, 3, kb - kb0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint64_t a_desc = umma_desc_sw128_kmajor(a_addr);
        const uint64_t b_desc = umma_desc_sw128_kmajor(a_addr + a_stage);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
          // +32 bytes per 16-element K step inside the 128-byte swizzle atom (address field is >>4)
          umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the smem stage once these MMAs have read it
        trace_fine(fine, 4, kb - kb0);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    const int et = threadIdx.x - 64;              // 0..127
    if (bnorm) {
      // ============ B-operand finishers: raw x tile -> rmsnorm(x) * w in bf16, in place ============
      // the norm weights are parameters: staged in shared memory before the dependency wait
      for (int i = et * 8; i < p.K; i += GEMM_EPI_THREADS * 8)
        *reinterpret_cast<uint4*>(w_sm + i) = __ldg(reinterpret_cast<const uint4*>(p.n_w + i));
      pdl_wait();
      // All four warps finish each tile together (tpt threads share a token, each owns `chunks` 16-byte chunks of the
      // 64-wide block): the ring is only ~5 stages deep here, so what matters is how LONG a stage waits for its
      // activation tile, not how many tiles are in flight (one warp per stage measured 26 us for the gate/up
      // projection against 20 us this way; the 10-slot persistent chain is where warp-per-stage pays).
      const int tpt = GEMM_EPI_THREADS / p.t_tile;            // 8, 4, 2   (t_tile 16, 32, 64)
      const int chunks = 8 / tpt;                             // 1, 2, 4
      const int t = et / tpt, c0 = (et % tpt) * chunks;
      float rstd = 0.f;
      if (t < p.T) {
        // per-tile partial sums of squares, 8 loads in flight at a time
        float ss = 0.f;
        for (int i0 = 0; i0 < p.n_ssq_parts; i0 += 8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[i] = (i0 + i < p.n_ssq_parts) ? __ldcg(&p.n_ssq[static_cast<size_t>(i0 + i) * p.T + t]) : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) ss += v[i];
        }
        rstd = rsqrtf(ss / static_cast<float>(p.K) + p.n_eps);
      }
      epi_bar();                                              // w_sm complete
      const uint32_t row_off = static_cast<uint32_t>(t) * 128u;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        uint8_t* b = smem + s * stage_bytes + a_stage + row_off;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < chunks) {
            uint4* px = reinterpret_cast<uint4*>(b + (((c0 + c) ^ (t & 7)) << 4));
            const uint4 xv = *px;
            const uint4 wv = *reinterpret_cast<const uint4*>(w_sm + kb * GEMM_BLOCK_K + (c0 + c) * 8);
            const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
            const uint32_t ws_[4] = {wv.x, wv.y, wv.z, wv.w};
            uint32_t r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              r[j] = pack_bf16(bf16_lo(xs[j]) * rstd * bf16_lo(ws_[j]), bf16_hi(xs[j]) * rstd * bf16_hi(ws_[j]));
            *px = make_uint4(r[0], r[1], r[2], r[3]);
          }
        }
        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfull[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    if (reducing) {
      const int quarter_ = warp & 3, row_ = quarter_ * 32 + lane;
      const int S_ = p.split_k;
      const int rank_ = S_ > 1 ? static_cast<int>(cluster_ctarank()) : 0;
      if (p.mode == GM_RESID) {
        if (!bnorm) pdl_wait();           // the residual may come from the kernel right in front of this one
        const int n_ = n_tile * p.tile_rows + row_;
#pragma unroll
        for (int i = 0; i < GEMM_EPI_CHUNK; ++i) {
          const int t = rank_ + i * S_;
          pre_a[i] = 0.f;
          if (p.residual && t < p.T && row_ < p.tile_rows && n_ < p.N) {
            const unsigned short rb =
                __ldcg(reinterpret_cast<const unsigned short*>(p.residual) + static_cast<size_t>(t) * p.N + n_);
            pre_a[i] = __uint_as_float(static_cast<uint32_t>(rb) << 16);
          }
        }
      } else {
        const int D_ = p.tile_rows;
        const bool rot = n_tile < p.n_q + p.n_kv;
#pragma unroll
        for (int i = 0; i < GEMM_EPI_CHUNK; ++i) {
          const int t = rank_ + i * S_;
          const bool ok = t < p.T && row_ < D_;
          pre_a[i] = (ok && rot) ? p.rope_cs[static_cast<size_t>(t) * 2 * D_ + row_] : 1.f;
          pre_b[i] = (ok && rot) ? p.rope_cs[static_cast<size_t>(t) * 2 * D_ + D_ + row_] : 0.f;
          pre_pg[i] = (ok && n_tile >= p.n_q) ? p.row_page[t] : -1;
          pre_sl[i] = (ok && n_tile >= p.n_q) ? p.row_slot[t] : 0;
        }
      }
    }
    // ================= epilogue: TMEM -> registers -> global =================
    // (stores happen after the accumulator is complete, i.e. after activation tiles that were only requested
    // once the previous kernel had finished: no explicit wait needed on this path)
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int row = quarter * 32 + lane;          // accumulator row = weight row inside the tile
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (et == 0 && trace_block0()) trace_mark(21, p.mode);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int t_base = t_blk * p.t_tile;
    if (p.mode == GM_SILU) {
      // Rows are packed per warp: lanes 0..15 hold 16 gate rows, lanes 16..31 the 16 matching up rows (see
      // ops.interleave_gate_up), so no shared-memory exchange and no barrier: each of the four epilogue warps runs
      // on its own.  Both half-warps work: lanes 0..15 finish the even tokens of their output column, lanes 16..31
      // the odd ones (gate and up cross over with one shuffle each).  The warp is alone on its scheduler, so the
      // arithmetic is kept short: the store index is split into a per-thread part and a 3-instruction per-token part.
      const int h = p.tile_rows >> 1;
      const bool hi = lane >= 16;
      const int n_out = n_tile * h + quarter * 16 + (lane & 15);
      const bool live = row < p.tile_rows && n_out < p.n_out;
      __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y);
      // tiled output: element (t, n) of token block t_blk sits at col_base + tt * 64 + ((c ^ (tt & 7)) << 3)
      const int yc = (n_out >> 3) & 7;
      const size_t col_base = p.y_tiled
          ? ((static_cast<size_t>(t_blk) * ((p.n_out + 63) >> 6) + (n_out >> 6)) * p.t_tile) * 64 + (n_out & 7)
          : static_cast<size_t>(t_base) * p.ldy + n_out;
      for (int c0 = 0; c0 < p.t_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // this lane finishes token c0 + 2 i + hi: gate from the low half-warp, up from the high one
          const float mine = __uint_as_float(v[2 * i + (hi ? 1 : 0)]);
          const float send = __uint_as_float(v[2 * i + (hi ? 0 : 1)]);       // what the partner lane needs from here
          const float other = __shfl_xor_sync(0xffffffffu, send, 16);
          const float g = round_bf16(hi ? other : mine);
          const float up = round_bf16(hi ? mine : other);
          const float s = round_bf16(g / (1.0f + expf(-g)));
          o[i] = s * up;
        }
        if (live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int tt = c0 + 2 * i + (hi ? 1 : 0);
            if (t_base + tt < p.T) {
              const size_t yi = p.y_tiled ? col_base + static_cast<size_t>(tt) * 64 + ((yc ^ (tt & 7)) << 3)
                                          : col_base + static_cast<size_t>(tt) * p.ldy;
              y[yi] = __float2bfloat16_rn(o[i]);
            }
          }
        }
      }
    } else if (!reducing) {
      const int n = n_tile * p.tile_rows + row;
      const bool valid = row < p.tile_rows && n < p.N;
      for (int c0 = 0; c0 < p.t_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          if (p.mode == GM_BF16) {
            __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y);
            const float bv = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = t_base + c0 + j;
              if (t < p.T) y[static_cast<size_t>(t) * p.ldy + n] = __float2bfloat16_rn(__uint_as_float(v[j]) + bv);
            }
          } else {
            float* y = static_cast<float*>(p.y) + static_cast<size_t>(split) * p.T * p.ldy;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = t_base + c0 + j;
              if (t < p.T) y[static_cast<size_t>(t) * p.ldy + n] = __uint_as_float(v[j]);
            }
          }
        }
      }
    } else {
      // ---- modes 3, 4, step 1: park this CTA's partial tile in its own shared memory: part_sm[t][row] ----
      for (int c0 = 0; c0 < p.t_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) part_sm[(c0 + j) * 128 + row] = __uint_as_float(v[j]);
      }
    }
    tc_fence_before();
  }

  if (reducing) {
    // ---- modes 3, 4, step 2: the split_k CTAs of a tile form a thread-block cluster; CTA r finishes the
    // tokens t = r, r + S, ...: sums the S partials through distributed shared memory in split order
    // (deterministic) and runs the fused tail on them ----
    const int S = p.split_k;
    if (S > 1) cluster_sync_all(); else __syncthreads();
    if (warp >= 2) {
      const int quarter = warp & 3, row = quarter * 32 + lane, et = threadIdx.x - 64;
      const int rank = S > 1 ? static_cast<int>(cluster_ctarank()) : 0;
      uint32_t peer[GEMM_MAX_SPLIT_CLUSTER];
#pragma unroll
      for (int s = 0; s < GEMM_MAX_SPLIT_CLUSTER; ++s)
        peer[s] = (S > 1 && s < S) ? dsmem_addr(part_sm, s) : smem_u32(part_sm);
      const int n = n_tile * p.tile_rows + row;
      const bool valid = row < p.tile_rows && n < p.N;
      const int n_mine = (p.T > rank) ? (p.T - rank + S - 1) / S : 0;   // tokens rank, rank + S, ... < T
      for (int i0 = 0; i0 < n_mine; i0 += GEMM_EPI_CHUNK) {
        float a[GEMM_EPI_CHUNK];
#pragma unroll
        for (int i = 0; i < GEMM_EPI_CHUNK; ++i) a[i] = 0.f;
#pragma unroll
        for (int s = 0; s < GEMM_MAX_SPLIT_CLUSTER; ++s) {
          if (s < S) {
#pragma unroll
            for (int i = 0; i < GEMM_EPI_CHUNK; ++i) {
              const int t = rank + (i0 + i) * S;
              const float v = (i0 + i < n_mine) ? ld_dsmem(peer[s] + static_cast<uint32_t>(t * 128 + row) * 4u) : 0.f;
              a[i] = (s == 0) ? v : a[i] + v;
            }
          }
        }
        if (p.mode == GM_RESID) {
          __nv_bfloat16* hid = static_cast<__nv_bfloat16*>(p.y);
          const int nkb_y = (p.N + 63) >> 6;
#pragma unroll
          for (int i = 0; i < GEMM_EPI_CHUNK; ++i) {
            const int t = rank + (i0 + i) * S;
            float sq = 0.f;
            if (valid && i0 + i < n_mine) {
              const size_t idx = static_cast<size_t>(t) * p.N + n;
              float hv = round_bf16(a[i]);
              if (p.residual) {
                float rv = pre_a[i];
                if (i0 > 0) rv = __bfloat162float(p.residual[idx]);     // beyond the prefetched chunk
                hv = round_bf16(rv + hv);
              }
              const __nv_bfloat16 hb = __float2bfloat16_rn(hv);
              hid[idx] = hb;
              if (p.y_tiles) p.y_tiles[xt_index(t, n, p.t_tile, nkb_y)] = hb;
              sq = hv * hv;
            }
            sq = warp_sum(sq);
            if (lane == 0) ssq_sm[quarter * GEMM_EPI_CHUNK + i] = sq;
          }
          epi_bar();
          if (et < GEMM_EPI_CHUNK && i0 + et < n_mine && p.ssq_out)
            p.ssq_out[static_cast<size_t>(n_tile) * p.T + rank + (i0 + et) * S] =
                (ssq_sm[et] + ssq_sm[GEMM_EPI_CHUNK + et]) + (ssq_sm[2 * GEMM_EPI_CHUNK + et] + ssq_sm[3 * GEMM_EPI_CHUNK + et]);
          epi_bar();
        } else {
          // GM_ROPE: the tile is head `n_tile`, the row is the element e of that head
          constexpr int ldx = GEMM_EPI_CHUNK + 1;
          const int D = p.tile_rows, half = D >> 1;
          const int head = n_tile;
          const bool rot = head < p.n_q + p.n_kv;       // V heads are not rotated
#pragma unroll
          for (int i = 0; i < GEMM_EPI_CHUNK; ++i) xchg[row * ldx + i] = round_bf16(a[i]);
          epi_bar();
          if (row < D) {
            const int prow = row < half ? row + half : row - half;
            const float sign = row < half ? -1.f : 1.f;
#pragma unroll
            for (int i = 0; i < GEMM_EPI_CHUNK; ++i) {
              const int t = rank + (i0 + i) * S;
              if (i0 + i < n_mine) {
                float v = xchg[row * ldx + i];
                if (rot) {
                  float cv = pre_a[i], sv = pre_b[i];
                  if (i0 > 0) {                  // beyond the prefetched chunk
                    const float* cs = p.rope_cs + static_cast<size_t>(t) * 2 * D;
                    cv = cs[row]; sv = cs[D + row];
                  }
                  v = v * cv + sign * xchg[prow * ldx + i] * sv;
                }
                const __nv_bfloat16 o = __float2bfloat16_rn(v);
                if (head < p.n_q) {
                  p.q_out[(static_cast<size_t>(t) * p.n_q + head) * D + row] = o;
                } else {
                  const int page = i0 > 0 ? p.row_page[t] : pre_pg[i];
                  if (page >= 0) {
                    const size_t row_elems = static_cast<size_t>(p.n_kv) * D;
                    const size_t slab = static_cast<size_t>(p.page_size) * row_elems;
                    const int hk = head - p.n_q;               // 0 .. 2 n_kv - 1: k heads then v heads
                    const int is_v = hk >= p.n_kv ? 1 : 0;
                    const int slot_ = i0 > 0 ? p.row_slot[t] : pre_sl[i];
                    p.kv[(static_cast<size_t>(page) * 2 + is_v) * slab + static_cast<size_t>(slot_) * row_elems +
                         static_cast<size_t>(hk - is_v * p.n_kv) * D + row] = o;
                  }
                }
              }
            }
          }
          epi_bar();
        }
      }
    }
    // nobody leaves (or frees its shared memory) while a peer may still be reading its partial tile
    if (S > 1) cluster_sync_all(); else __syncthreads();
  } else {
    __syncthreads();
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  trace_end(tr);
}

// nn.Linear weight [N][K] (leading dimension ldw) -> [n_tile][k_block][tile_rows][64] bf16 with the 128-byte
// swizzle of a K-major UMMA operand applied (16-byte chunk c of row r sits at chunk c ^ (r & 7)); rows past N and
// columns past K are zero.  One thread per 16-byte chunk.
__global__ void __launch_bounds__(256) pack_weight_tiles_kernel(uint4* __restrict__ dst,
                                                                const __nv_bfloat16* __restrict__ w, int N, int K,
                                                                long long ldw, int tile_rows, int num_kb) {
  const long long tile_kb = blockIdx.x;                       // n_tile * num_kb + kb
  const int n_tile = static_cast<int>(tile_kb / num_kb), kb = static_cast<int>(tile_kb % num_kb);
  for (int i = threadIdx.x; i < tile_rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int n = n_tile * tile_rows + r, k = kb * GEMM_BLOCK_K + c * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < N && k + 8 <= K) {
      v = *reinterpret_cast<const uint4*>(w + static_cast<long long>(n) * ldw + k);
    } else if (n < N && k < K) {
      __nv_bfloat16 tmp[8];
      for (int j = 0; j < 8; ++j) tmp[j] = (k + j < K) ? w[static_cast<long long>(n) * ldw + k + j] : __float2bfloat16(0.f);
      v = *reinterpret_cast<uint4*>(tmp);
    }
    dst[(tile_kb * tile_rows + r) * 8 + (c ^ (r & 7))] = v;
  }
}

// cos / sin of pos[t] * freq[e] for every row of the step (the same for all layers): [T][2][D]
__global__ void rope_table_kernel(float* __restrict__ cs, const int32_t* __restrict__ pos,
                                  const float* __restrict__ freq, int D) {
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(10) : -1;
  pdl_sync();
  const size_t t = blockIdx.x;
  const float ps = static_cast<float>(pos[t]);
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    float s, c;
    sincosf(ps * freq[e], &s, &c);
    cs[t * 2 * D + e] = c;
    cs[t * 2 * D + D + e] = s;
  }
  trace_end(tr);
}

// rows of x -> one sum of squares each (the ssq input of the norm-fused projections for a hidden state that
// did not come out of a GM_RESID projection: the embedding output)
__global__ void __launch_bounds__(256) row_ssq_kernel(float* __restrict__ ssq, const __nv_bfloat16* __restrict__ x,
                                                      int dim) {
  pdl_sync();
  __shared__ float red[8];
  const size_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * dim;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
      ss += a * a + b * b;
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    ssq[row] = tot;
  }
}

static int g_gemm_smem_budget = -1;
static int g_gemm_gu_kb = -1;

static int gemm_smem_budget() {
  if (g_gemm_smem_budget < 0) {
    // default: two CTAs per SM (a programmatic dependent can sit beside its predecessor and prefetch weights)
    int kb = 104;
    if (const char* e = getenv("VB_GEMM_SMEM_KB")) {
      const int v = atoi(e);
      if (v >= 48 && v <= 220) kb = v;
    }
    g_gemm_smem_budget = kb * 1024;
  }
  return g_gemm_smem_budget;
}

static int launch_gemm(GemmParams& p, const void* w_tiles, const void* x_map, cudaStream_t stream,
                       const void* x_tiles = nullptr) {
  static const CUtensorMap dummy_map = {};
  p.w_tiles = static_cast<const uint8_t*>(w_tiles);
  p.x_tiles = static_cast<const uint8_t*>(x_tiles);
  if (!x_map) x_map = &dummy_map;      // never dereferenced by the kernel when x_tiles is set
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  VB_CHECK_ARG(p.T > 0 && p.N > 0 && p.K > 0, "gemm: empty problem T=%d N=%d K=%d", p.T, p.N, p.K);
  VB_CHECK_ARG(p.K % 8 == 0, "gemm: K %d must be a multiple of 8", p.K);
  VB_CHECK_ARG(p.tile_rows >= 8 && p.tile_rows <= 128 && p.tile_rows % 8 == 0,
               "gemm: tile_rows %d must be a multiple of 8 in [8, 128]", p.tile_rows);
  VB_CHECK_ARG(p.split_k >= 1 && p.split_k <= num_kb, "gemm: split_k %d outside [1, %d]", p.split_k, num_kb);
  const bool reducing = p.mode == GM_RESID || p.mode == GM_ROPE;
  VB_CHECK_ARG(!reducing || p.split_k <= GEMM_MAX_SPLIT_CLUSTER,
               "gemm: split_k %d exceeds the cluster size limit %d of the fused modes", p.split_k, GEMM_MAX_SPLIT_CLUSTER);
  int cols = 32;
  while (cols < p.t_tile) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = p.tile_rows * 128 + p.t_tile * 128;
  const int extra = 1024 /*barriers, ssq*/ + (p.norm ? (p.K * 2 + 1023) / 1024 * 1024 : 0) + 1024 /*alignment slack*/;
  // after the main loop the ring holds the partial tile (t_tile x 128 fp32) and the 128 x 17 exchange buffer
  const int min_ring = p.t_tile * 512 + 128 * 17 * 4;
  int budget = gemm_smem_budget();
  if (p.mode == GM_SILU && !p.norm) {
    // the gate/up stream is the long one (48 stages per CTA) and it follows a small kernel: a 10-stage ring on the
    // whole SM beats co-residency here (forward 2.68 -> 2.64 ms; VB_GEMM_SMEM_KB_GU overrides)
    if (g_gemm_gu_kb < 0) { const char* e = getenv("VB_GEMM_SMEM_KB_GU"); g_gemm_gu_kb = e ? atoi(e) : 200; }
    // more tiles than SMs (GLM-4-Voice: 286): two CTAs per SM in one wave beat 1.9 waves of whole-SM CTAs
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int n_tiles = (p.N + p.tile_rows - 1) / p.tile_rows;
    if (g_gemm_gu_kb >= 48 && g_gemm_gu_kb <= 220 && (n_tiles <= sms || sms <= 0)) budget = g_gemm_gu_kb * 1024;
  }
  int stages = (budget - extra) / stage_bytes;
  if (stages > 12) stages = 12;
  if (stages > num_kb) stages = num_kb;
  if (stages < 2) stages = 2;
  while (stages * stage_bytes < min_ring) ++stages;
  p.pad = stage_bytes < 16384 ? 16384 - stage_bytes : 0;
  const int smem = stages * stage_bytes + p.pad + extra;
  VB_CHECK_ARG(stages <= 20 && smem <= VB_MAX_DYN_SMEM, "gemm: tile too large for shared memory (%d bytes)", smem);
  p.stages = stages;
  VB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  dim3 grid((p.N + p.tile_rows - 1) / p.tile_rows, p.split_k, (p.T + p.t_tile - 1) / p.t_tile);
  // the fused modes reduce split-K through distributed shared memory: cluster = the split_k CTAs of a tile
  const dim3 cluster(1, reducing ? p.split_k : 1, 1);
  VB_CHECK_CUDA(launch_kernel_cluster3(gemm_bf16_kernel, grid, dim3(GEMM_THREADS), smem, stream, true, cluster, p,
                                       *static_cast<const CUtensorMap*>(x_map)));
  return 0;
}

static int fused_t_tile(int T) { return T <= 16 ? 16 : (T <= 32 ? 32 : 64); }

}  // namespace vb

using namespace vb;

extern "C" {

size_t vb_weight_tiles_bytes(int N, int K, int tile_rows) {
  if (N <= 0 || K <= 0 || tile_rows < 8 || tile_rows > 128 || tile_rows % 8) return 0;
  const size_t tiles = (static_cast<size_t>(N) + tile_rows - 1) / tile_rows, num_kb = (static_cast<size_t>(K) + 63) / 64;
  return tiles * num_kb * tile_rows * 128;
}

int vb_pack_weight_tiles(void* d_dst, const void* d_w, int N, int K, int64_t ldw, int tile_rows, void* stream) {
  VB_CHECK_ARG(d_dst && d_w, "vb_pack_weight_tiles: null pointer");
  VB_CHECK_ARG(vb_weight_tiles_bytes(N, K, tile_rows) > 0, "vb_pack_weight_tiles: bad shape N=%d K=%d tile_rows=%d", N, K,
               tile_rows);
  VB_CHECK_ARG(ldw % 8 == 0, "vb_pack_weight_tiles: leading dimension must be a multiple of 8 elements");
  const int tiles = (N + tile_rows - 1) / tile_rows, num_kb = (K + 63) / 64;
  VB_LAUNCH_PLAIN(pack_weight_tiles_kernel, static_cast<unsigned>(static_cast<long long>(tiles) * num_kb), 256, 0, stream,
                  static_cast<uint4*>(d_dst), static_cast<const __nv_bfloat16*>(d_w), N, K, static_cast<long long>(ldw),
                  tile_rows, num_kb);
  return 0;
}

int vb_set_gemm_smem_kb(int ring_kb, int gate_up_ring_kb) {
  VB_CHECK_ARG((ring_kb == 0 || (ring_kb >= 48 && ring_kb <= 220)) && (gate_up_ring_kb == 0 || (gate_up_ring_kb >= 48 && gate_up_ring_kb <= 220)),
               "vb_set_gemm_smem_kb: budgets are 0 (keep) or 48..220 KiB");
  if (ring_kb) g_gemm_smem_budget = ring_kb * 1024;
  if (gate_up_ring_kb) g_gemm_gu_kb = gate_up_ring_kb;
  return 0;
}

int vb_gemm_t_tile(int T) {
  // decode-sized steps: 16 / 32 / 64 (what the fused modes and the tiled activation layout use); beyond that the
  // token tile grows in steps of 16 up to the 256 columns one MMA takes
  if (T <= 16) return 16;
  if (T <= 32) return 32;
  if (T <= 64) return 64;
  if (T >= 256) return 256;
  return (T + 15) / 16 * 16;
}

int vb_gemm_bf16(void* d_y, const void* d_w_tiles, const void* x_map, const void* d_x_tiles, int T, int N, int K,
                 int ldy, int mode, int split_k, int tile_rows, int n_out, int y_tiled, const void* d_bias, void* stream) {
  VB_CHECK_ARG(d_y && d_w_tiles && (x_map || d_x_tiles), "vb_gemm_bf16: null pointer");
  VB_CHECK_ARG(!d_bias || mode == 0, "vb_gemm_bf16: bias is a mode 0 option");
  VB_CHECK_ARG(!y_tiled || mode == 2, "vb_gemm_bf16: tiled output is a mode 2 option");
  VB_CHECK_ARG(mode >= 0 && mode <= 2, "vb_gemm_bf16: mode %d", mode);
  VB_CHECK_ARG(mode == 1 || split_k == 1, "vb_gemm_bf16: split_k > 1 needs mode 1 (fp32 partials)");
  VB_CHECK_ARG(mode != 2 || (tile_rows % 32 == 0 && N % tile_rows == 0),
               "vb_gemm_bf16: mode 2 needs tile_rows %% 32 == 0 and N %% tile_rows == 0 (16 gate + 16 up rows per warp)");
  GemmParams p = {};
  p.y = d_y; p.T = T; p.N = N; p.K = K; p.ldy = ldy; p.mode = mode; p.split_k = split_k;
  p.tile_rows = tile_rows > 0 ? tile_rows : 128;
  p.n_out = n_out > 0 ? n_out : (mode == 2 ? N / 2 : N);
  p.t_tile = vb_gemm_t_tile(T);
  p.y_tiled = y_tiled;
  p.bias = static_cast<const __nv_bfloat16*>(d_bias);
  return launch_gemm(p, d_w_tiles, x_map, static_cast<cudaStream_t>(stream), d_x_tiles);
}

size_t vb_norm_lmhead_workspace_bytes(int T, int K) {
  if (T <= 0 || K <= 0) return 0;
  const int t_tile = vb_gemm_t_tile(T);
  return static_cast<size_t>((T + t_tile - 1) / t_tile) * ((K + 63) / 64) * t_tile * 64 * sizeof(__nv_bfloat16);
}

// logits = lm_head(rmsnorm(hidden) * norm_weight) (+ bias): the normed rows go straight into the tiled activation
// layout the projection streams (no row-major round trip), the projection is a programmatic dependent launch of the
// norm and has its first ring pass of lm_head tiles in flight before the normed rows exist.
int vb_norm_lmhead(void* d_logits, const void* d_hidden, const void* d_norm_weight, float eps, const void* d_w_tiles,
                   const void* d_bias, int T, int N, int K, int ldy, int tile_rows, void* d_workspace,
                   size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(d_logits && d_hidden && d_norm_weight && d_w_tiles && d_workspace, "vb_norm_lmhead: null pointer");
  VB_CHECK_ARG(workspace_bytes >= vb_norm_lmhead_workspace_bytes(T, K), "vb_norm_lmhead: workspace too small");
  if (T <= 0) return 0;
  const int rc = vb_rmsnorm(d_workspace, d_hidden, d_norm_weight, T, K, eps, vb_gemm_t_tile(T), stream);
  if (rc != 0) return rc;
  return vb_gemm_bf16(d_logits, d_w_tiles, nullptr, d_workspace, T, N, K, ldy, 0, 1, tile_rows, 0, 0, d_bias, stream);
}

int vb_proj_residual(void* d_hidden_out, void* d_hidden_tiles_out, float* d_ssq_out, const void* d_w_tiles,
                     const void* x_map, const void* d_x_tiles, const void* d_residual, int T, int N, int K, int split_k,
                     int tile_rows, void* stream) {
  VB_CHECK_ARG(d_hidden_out && d_w_tiles && (x_map || d_x_tiles), "vb_proj_residual: null pointer");
  VB_CHECK_ARG(T > 0 && T <= 64, "vb_proj_residual: T %d outside (0, 64] (decode-sized batches only)", T);
  GemmParams p = {};
  p.y = d_hidden_out; p.T = T; p.N = N; p.K = K; p.ldy = N; p.mode = GM_RESID; p.split_k = split_k;
  p.tile_rows = tile_rows > 0 ? tile_rows : 128;
  p.n_out = N;
  p.t_tile = fused_t_tile(T);
  p.residual = static_cast<const __nv_bfloat16*>(d_residual);
  p.ssq_out = d_ssq_out;
  p.y_tiles = static_cast<__nv_bfloat16*>(d_hidden_tiles_out);
  return launch_gemm(p, d_w_tiles, x_map, static_cast<cudaStream_t>(stream), d_x_tiles);
}

int vb_proj_norm_gateup_silu(void* d_act_out, const void* d_w_tiles, const void* x_map, const void* d_x_tiles,
                             const float* d_ssq, int n_ssq_parts, const void* d_norm_weight, float eps, int T,
                             int N_packed, int K, int tile_rows, int n_out, int y_tiled, void* stream) {
  VB_CHECK_ARG(d_act_out && d_w_tiles && (x_map || d_x_tiles) && d_ssq && d_norm_weight,
               "vb_proj_norm_gateup_silu: null pointer");
  VB_CHECK_ARG(T > 0 && T <= 64, "vb_proj_norm_gateup_silu: T %d outside (0, 64]", T);
  VB_CHECK_ARG(tile_rows % 32 == 0 && N_packed % tile_rows == 0, "vb_proj_norm_gateup_silu: bad tile_rows %d", tile_rows);
  VB_CHECK_ARG(K % 64 == 0, "vb_proj_norm_gateup_silu: K %d must be a multiple of 64", K);
  GemmParams p = {};
  p.y = d_act_out; p.T = T; p.N = N_packed; p.K = K; p.ldy = n_out; p.mode = GM_SILU; p.split_k = 1;
  p.tile_rows = tile_rows; p.n_out = n_out;
  p.t_tile = fused_t_tile(T);
  p.norm = 1; p.n_ssq = d_ssq; p.n_ssq_parts = n_ssq_parts;
  p.n_w = static_cast<const __nv_bfloat16*>(d_norm_weight); p.n_eps = eps;
  p.y_tiled = y_tiled;
  return launch_gemm(p, d_w_tiles, x_map, static_cast<cudaStream_t>(stream), d_x_tiles);
}

int vb_proj_norm_qkv_rope_append(void* d_q_out, void* d_layer_kv, const void* d_w_tiles, const void* x_map,
                                 const void* d_x_tiles, const float* d_ssq, int n_ssq_parts, const void* d_norm_weight, float eps,
                                 const float* d_rope_cs, const int32_t* d_row_page, const int32_t* d_row_slot, int T,
                                 int K, int n_q, int n_kv, int head_dim, int page_size, int split_k, void* stream) {
  VB_CHECK_ARG(d_q_out && d_layer_kv && d_w_tiles && (x_map || d_x_tiles) && d_ssq && d_norm_weight && d_rope_cs && d_row_page &&
                   d_row_slot,
               "vb_proj_norm_qkv_rope_append: null pointer");
  VB_CHECK_ARG(T > 0 && T <= 64, "vb_proj_norm_qkv_rope_append: T %d outside (0, 64]", T);
  VB_CHECK_ARG(head_dim == 64 || head_dim == 128, "vb_proj_norm_qkv_rope_append: head_dim %d (64 or 128)", head_dim);
  VB_CHECK_ARG(K % 64 == 0, "vb_proj_norm_qkv_rope_append: K %d must be a multiple of 64", K);
  GemmParams p = {};
  const int tiles = n_q + 2 * n_kv;
  p.T = T; p.N = tiles * head_dim; p.K = K; p.ldy = 0; p.mode = GM_ROPE; p.split_k = split_k;
  p.tile_rows = head_dim; p.n_out = p.N;
  p.t_tile = fused_t_tile(T);
  p.norm = 1; p.n_ssq = d_ssq; p.n_ssq_parts = n_ssq_parts;
  p.n_w = static_cast<const __nv_bfloat16*>(d_norm_weight); p.n_eps = eps;
  p.q_out = static_cast<__nv_bfloat16*>(d_q_out); p.kv = static_cast<__nv_bfloat16*>(d_layer_kv);
  p.rope_cs = d_rope_cs; p.row_page = d_row_page; p.row_slot = d_row_slot;
  p.n_q = n_q; p.n_kv = n_kv; p.page_size = page_size;
  return launch_gemm(p, d_w_tiles, x_map, static_cast<cudaStream_t>(stream), d_x_tiles);
}

int vb_rope_table(float* d_cs, const int32_t* d_pos, const float* d_freq, int T, int head_dim, void* stream) {
  VB_CHECK_ARG(d_cs && d_pos && d_freq, "vb_rope_table: null pointer");
  if (T <= 0) return 0;
  VB_LAUNCH_PDL(rope_table_kernel, T, 128, 0, stream, d_cs, d_pos, d_freq, head_dim);
  return 0;
}

int vb_row_ssq(float* d_ssq, const void* d_x, int rows, int dim, void* stream) {
  VB_CHECK_ARG(d_ssq && d_x, "vb_row_ssq: null pointer");
  VB_CHECK_ARG(dim > 0 && dim % 8 == 0, "vb_row_ssq: dim %d must be a positive multiple of 8", dim);
  if (rows <= 0) return 0;
  VB_LAUNCH_PDL(row_ssq_kernel, rows, 256, 0, stream, d_ssq, static_cast<const __nv_bfloat16*>(d_x), dim);
  return 0;
}

}  // extern "C"

VB_DEFINE_TRACE_SETTER(gemm)
