// Dense projections Y[T][N] = X[T][K] * W[N][K]^T for the LM decode / prefill step on tcgen05.
// Replaces the nn.Linear calls of vox_serve/model/orpheus.py:41-47, 68-79, 91-93, 110, 197, 219.
//
// At decode time T <= 32, so the problem is weight streaming (HBM-bound).  The weight matrix is the
// 128-row UMMA "A" operand (M = 128 output features per CTA), the token rows are the "B" operand
// (UMMA N = T rounded up to 16, <= 256), both K-major bf16, fp32 accumulators in TMEM.  Warp roles:
// warp 0 = TMA producer (64-wide K blocks into a multi-stage 128B-swizzled ring), warp 1 = single-thread
// tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld, each warp owns its 32-lane TMEM quarter).
// Split-K (blockIdx.y) spreads skinny problems over all SMs; partials are written as fp32 planes and
// summed in split order by the fused consumers in elementwise.cu (deterministic).
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_A_STAGE = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KiB
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  void* y;
  const void* next_w;       // weights of the projection that runs after this one: pulled into L2 meanwhile
  size_t next_bytes;
  int T, N, K, ldy, mode, split_k, t_tile, stages, tmem_cols;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_kernel(const GemmParams p,
                                                                    const __grid_constant__ CUtensorMap w_map,
                                                                    const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage = p.t_tile * 128;
  const int stage_bytes = GEMM_A_STAGE + b_stage;
  uint8_t* tail = smem + p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* xchg = reinterpret_cast<float*>(tail + 1024);  // mode 2: [64][t_tile + 1]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, split = blockIdx.y, t_blk = blockIdx.z;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int kb0 = static_cast<int>(static_cast<long long>(split) * num_kb / p.split_k);
  const int kb1 = static_cast<int>(static_cast<long long>(split + 1) * num_kb / p.split_k);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&w_map);
    prefetch_tmap(&x_map);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
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

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();   // weights are streamed once per step
      const uint64_t pol_x = policy_evict_last();    // activations are re-read by every CTA
      // The weights do not depend on the previous kernel: fill the whole ring with weight tiles while it is
      // still running, then wait for it and add the activation tiles to the same stages.
      const int npre = min(p.stages, kb1 - kb0);
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(&full[i], stage_bytes);
        tma_load_2d_hint(smem + i * stage_bytes, &w_map, &full[i], (kb0 + i) * GEMM_BLOCK_K, n_tile * GEMM_BLOCK_M,
                         pol_w);
      }
      pdl_wait();
      pdl_trigger();
      for (int i = 0; i < npre; ++i)
        tma_load_2d_hint(smem + i * stage_bytes + GEMM_A_STAGE, &x_map, &full[i], (kb0 + i) * GEMM_BLOCK_K,
                         t_blk * p.t_tile, pol_x);
      int s = npre == p.stages ? 0 : npre;
      uint32_t ph = npre == p.stages ? 1 : 0;
      for (int kb = kb0 + npre; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_2d_hint(a, &w_map, &full[s], kb * GEMM_BLOCK_K, n_tile * GEMM_BLOCK_M, pol_w);
        tma_load_2d_hint(a + GEMM_A_STAGE, &x_map, &full[s], kb * GEMM_BLOCK_K, t_blk * p.t_tile, pol_x);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(GEMM_BLOCK_M, p.t_tile, 1u);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint64_t a_desc = umma_desc_sw128_kmajor(a_addr);
        const uint64_t b_desc = umma_desc_sw128_kmajor(a_addr + GEMM_A_STAGE);
#pragma unroll
        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
          // +32 bytes per 16-element K step inside the 128-byte swizzle atom (address field is >>4)
          umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the smem stage once these MMAs have read it
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ================= epilogue: TMEM -> registers -> global =================
    // while the accumulator is being produced these warps are idle: warp 2 pulls this CTA's share of the NEXT
    // projection's weights into L2 (weights never depend on a predecessor, so this is legal before any wait)
    if (warp == 2)
      prefetch_l2_slice(p.next_w, p.next_bytes, blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z),
                        gridDim.x * gridDim.y * gridDim.z, lane, 32);
    // (stores happen after the accumulator is complete, i.e. after activation tiles that were only requested
    // once the previous kernel had finished: no explicit wait needed on this path)
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int row = quarter * 32 + lane;          // accumulator row = output feature inside the tile
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int t_base = t_blk * p.t_tile;
    if (p.mode == 2) {
      const bool is_up = row >= 64;
      const int ldx = p.t_tile + 1;
      for (int c0 = 0; c0 < p.t_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        if (is_up) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xchg[(row - 64) * ldx + c0 + j] = round_bf16(__uint_as_float(v[j]));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (!is_up) {
          const int n_out = n_tile * 64 + row;
          if (n_out < p.N / 2) {
            __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = t_base + c0 + j;
              if (t < p.T) {
                const float g = round_bf16(__uint_as_float(v[j]));
                const float s = round_bf16(g / (1.0f + expf(-g)));
                y[static_cast<size_t>(t) * p.ldy + n_out] = __float2bfloat16_rn(s * xchg[row * ldx + c0 + j]);
              }
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    } else {
      const int n = n_tile * GEMM_BLOCK_M + row;
      for (int c0 = 0; c0 < p.t_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        if (n < p.N) {
          if (p.mode == 0) {
            __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = t_base + c0 + j;
              if (t < p.T) y[static_cast<size_t>(t) * p.ldy + n] = __float2bfloat16_rn(__uint_as_float(v[j]));
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
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_gemm_t_tile(int T) {
  if (T <= 0) return 16;
  if (T >= 256) return 256;
  return (T + 15) / 16 * 16;
}

int vb_gemm_bf16(void* d_y, const void* w_map, const void* x_map, int T, int N, int K, int ldy, int mode,
                 int split_k, const void* d_prefetch, size_t prefetch_bytes, void* stream) {
  VB_CHECK_ARG(d_y && w_map && x_map, "vb_gemm_bf16: null pointer");
  VB_CHECK_ARG(T > 0 && N > 0 && K > 0, "vb_gemm_bf16: empty problem T=%d N=%d K=%d", T, N, K);
  VB_CHECK_ARG(mode >= 0 && mode <= 2, "vb_gemm_bf16: mode %d", mode);
  VB_CHECK_ARG(K % 8 == 0, "vb_gemm_bf16: K %d must be a multiple of 8", K);
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  VB_CHECK_ARG(split_k >= 1 && split_k <= num_kb, "vb_gemm_bf16: split_k %d outside [1, %d]", split_k, num_kb);
  VB_CHECK_ARG(mode == 1 || split_k == 1, "vb_gemm_bf16: split_k > 1 needs mode 1 (fp32 partials)");
  VB_CHECK_ARG(mode != 2 || N % 128 == 0, "vb_gemm_bf16: mode 2 needs N %% 128 == 0 (64 gate + 64 up rows per tile)");
  GemmParams p;
  p.y = d_y; p.next_w = d_prefetch; p.next_bytes = d_prefetch ? prefetch_bytes : 0; p.T = T; p.N = N; p.K = K; p.ldy = ldy; p.mode = mode; p.split_k = split_k;
  p.t_tile = vb_gemm_t_tile(T);
  int cols = 32;
  while (cols < p.t_tile) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = GEMM_A_STAGE + p.t_tile * 128;
  const int extra = 1024 + (mode == 2 ? 64 * (p.t_tile + 1) * 4 : 0) + 1024 /*align slack*/;
  int stages = (200 * 1024 - extra) / stage_bytes;
  if (stages > 10) stages = 10;
  if (stages > num_kb) stages = num_kb < 2 ? 2 : num_kb;
  VB_CHECK_ARG(stages >= 2, "vb_gemm_bf16: tile too large for shared memory");
  p.stages = stages;
  const int smem = stages * stage_bytes + extra;
  VB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  dim3 grid((N + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, split_k, (T + p.t_tile - 1) / p.t_tile);
  VB_LAUNCH_PDL(gemm_bf16_kernel, grid, GEMM_THREADS, smem, stream, p, *static_cast<const CUtensorMap*>(w_map), *static_cast<const CUtensorMap*>(x_map));
  return 0;
}

}  // extern "C"
