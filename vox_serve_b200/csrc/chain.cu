// Persistent projection chain for decode-sized steps: up to four dependent projections of the transformer layer
//     O-proj + residual  ->  RMSNorm + gate/up + SiLU*up  ->  down-proj + residual  ->  RMSNorm + QKV + RoPE + KV append
// (vox_serve/model/orpheus.py:81-151, the part of the layer between two attention calls) as ONE launch of at most
// one CTA per SM.  What a chain of separate kernels cannot do, this one does: the weight stream never stops at a
// phase boundary.  Weights depend on nothing, so the producer warp keeps filling the shared-memory ring with the
// tiles of the NEXT phase while the current phase is still being reduced, normalised, written back and
// synchronised; the boundary costs (grid-wide dependency, activation reload, epilogue) are paid from the ~180 KiB
// of weights already on chip instead of from an idle HBM.
//
// Roles (224 threads): warp 0 = weight producer (linear bulk copies of pre-tiled, pre-swizzled weight tiles, see
// vb_pack_weight_tiles), warp 6 = activation loader (TMA tiles of the phase's input, issued once the phase it
// depends on is complete grid-wide), warp 1 = tcgen05.mma issuer (accumulators in TMEM), warps 2-5 = B-operand
// finishers (RMSNorm applied in place on the raw activation tile) and epilogue.
// Phases are separated by a grid-wide arrival counter in global memory (all CTAs are co-resident: grid <= SMs,
// one CTA per SM); split-K partial tiles are exchanged through an L2-resident workspace, every CTA of a tile
// finishing the tokens t = split, split + S, ... in split order (deterministic).
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

constexpr int CH_BLOCK_K = 64;
constexpr int CH_THREADS = 224;
constexpr int CH_EPI_THREADS = 128;
constexpr int CH_MAX_PHASES = 4;
constexpr int CH_CHUNK = 8;
constexpr int CH_A_BYTES = 16384;        // every ring slot reserves a full 128-row operand (the MMA reads M = 128)

enum { CK_RESID = 0, CK_SILU = 1, CK_ROPE = 2 };

struct ChainPhase {
  const uint8_t* w_tiles;
  void* y;
  const float* n_ssq;
  const __nv_bfloat16* n_w;
  const __nv_bfloat16* residual;
  float* ssq_out;
  __nv_bfloat16* kv;
  int kind, N, K, tile_rows, n_tiles, split_k, n_out, n_ssq_parts;
  float n_eps;
  int pad_;
};
struct ChainParams {
  ChainPhase ph[CH_MAX_PHASES];
  float* ws;               // split-K partial tiles [item][t_tile][128] fp32
  unsigned int* flags;     // [CH_MAX_PHASES][1 + max_tiles]: grid arrival counter, then per-tile counters; zero at launch
  const float* rope_cs;
  const int32_t* row_page;
  const int32_t* row_slot;
  int n_phases, T, t_tile, stages, max_tiles, n_q, n_kv, page_size, norm_k;
};

__device__ __forceinline__ void ch_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned target) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    if (ld_acquire_gpu(p) >= target) return;
    __nanosleep(20);
  }
  __trap();
}
__device__ __forceinline__ void ch_epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct ChItem {
  bool has;
  int tile, split, kb0, kb1, num_kb;
};
__device__ __forceinline__ ChItem ch_item(const ChainPhase& ph) {
  ChItem it;
  const int items = ph.n_tiles * ph.split_k;
  it.has = static_cast<int>(blockIdx.x) < items;
  it.tile = blockIdx.x / ph.split_k;
  it.split = blockIdx.x % ph.split_k;
  it.num_kb = (ph.K + CH_BLOCK_K - 1) / CH_BLOCK_K;
  it.kb0 = static_cast<int>(static_cast<long long>(it.split) * it.num_kb / ph.split_k);
  it.kb1 = static_cast<int>(static_cast<long long>(it.split + 1) * it.num_kb / ph.split_k);
  return it;
}

__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const __grid_constant__ ChainParams P,
                                                              const __grid_constant__ CUtensorMap xm0,
                                                              const __grid_constant__ CUtensorMap xm1,
                                                              const __grid_constant__ CUtensorMap xm2,
                                                              const __grid_constant__ CUtensorMap xm3) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slot_bytes = CH_A_BYTES + P.t_tile * 128;
  uint8_t* tail = smem + P.stages * slot_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);           // weight + activation bytes of the slot have landed
  uint64_t* empty = full + P.stages;                            // the MMAs have read the slot
  uint64_t* cfull = empty + P.stages;                           // the B tile is final (normalised where needed)
  uint64_t* tmem_full = cfull + P.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* ssq_sm = reinterpret_cast<float*>(tail + 768);         // [4][CH_CHUNK]
  float* xchg = reinterpret_cast<float*>(tail + 1024);          // [128][17]
  __nv_bfloat16* w_sm = reinterpret_cast<__nv_bfloat16*>(tail + 1024 + 128 * 17 * 4);   // [2][norm_k]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = 1 + P.max_tiles;
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(3, P.n_phases) : -1;
  const CUtensorMap* xmaps[CH_MAX_PHASES] = {&xm0, &xm1, &xm2, &xm3};

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 2);                 // weight producer + activation loader (each with its byte count)
      mbar_init(&empty[s], 1);
      mbar_init(&cfull[s], CH_EPI_THREADS / 32);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= weight producer: runs ahead across phase boundaries =================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();
      uint32_t g = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        const ChItem it = ch_item(ph);
        if (!it.has) continue;
        const uint32_t a_bytes = ph.tile_rows * 128;
        const uint8_t* src = ph.w_tiles + (static_cast<size_t>(it.tile) * it.num_kb + it.kb0) * a_bytes;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          const uint32_t slot = g % P.stages, par = (g / P.stages) & 1;
          mbar_wait(&empty[slot], par ^ 1);
          trace_fine(0, g);
          mbar_arrive_expect_tx(&full[slot], a_bytes);
          ch_bulk_g2s(smem + slot * slot_bytes, src + static_cast<size_t>(kb - it.kb0) * a_bytes, a_bytes, &full[slot],
                      pol_w);
        }
      }
    }
  } else if (warp == 6) {
    // ================= activation loader: waits for the producing phase, then feeds the same slots =================
    if (lane == 0) {
      const uint64_t pol_x = policy_evict_last();
      const uint32_t b_bytes = P.t_tile * 128;
      uint32_t g = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        if (p == 0) {
          prefetch_tmap(xmaps[0]);
          pdl_wait();           // the kernel in front of the chain (attention / embedding) is complete
          pdl_trigger();
        } else {
          spin_until(&P.flags[(p - 1) * stride], gridDim.x);      // every CTA is past phase p - 1
          asm volatile("fence.proxy.async;" ::: "memory");        // their generic-proxy writes -> our TMA reads
        }
        if (trace_block0()) trace_mark(30 + p);
        const ChItem it = ch_item(ph);
        if (!it.has) continue;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          const uint32_t slot = g % P.stages, par = (g / P.stages) & 1;
          mbar_wait(&empty[slot], par ^ 1);
          trace_fine(1, g);
          mbar_arrive_expect_tx(&full[slot], b_bytes);
          tma_load_2d_hint(smem + slot * slot_bytes + CH_A_BYTES, xmaps[p], &full[slot], kb * CH_BLOCK_K, 0, pol_x);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, P.t_tile, 1u);
      uint32_t g = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChItem it = ch_item(P.ph[p]);
        if (!it.has) continue;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          const uint32_t slot = g % P.stages, par = (g / P.stages) & 1;
          mbar_wait(&cfull[slot], par);
          trace_fine(3, g);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + slot * slot_bytes);
          const uint64_t a_desc = umma_desc_sw128_kmajor(a_addr);
          const uint64_t b_desc = umma_desc_sw128_kmajor(a_addr + CH_A_BYTES);
#pragma unroll
          for (int k = 0; k < CH_BLOCK_K / 16; ++k)
            umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[slot]);
          trace_fine(4, g);
        }
        umma_commit(tmem_full);
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // ================= B-operand finishers + epilogue =================
    const int et = threadIdx.x - 64;
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int tpt = CH_EPI_THREADS / P.t_tile;              // threads per token: 8, 4, 2
    const int chunks = 8 / tpt;
    const int ct = et / tpt, cc0 = (et % tpt) * chunks;     // converter token and first 16-byte chunk
    const uint32_t crow_off = static_cast<uint32_t>(ct) * 128u;
    uint32_t g = 0, n_done = 0, n_norm = 0;
    for (int p = 0; p < P.n_phases; ++p) {
      const ChainPhase& ph = P.ph[p];
      const ChItem it = ch_item(ph);
      unsigned* gbar = &P.flags[p * stride];
      if (!it.has) {
        if (et == 0) red_release_gpu(gbar, 1u);
        continue;
      }
      const bool norm = ph.kind != CK_RESID;
      const __nv_bfloat16* wbuf = w_sm + (n_norm & 1) * P.norm_k;
      if (norm) {
        // norm weights are parameters: staged before the phase's input exists
        __nv_bfloat16* wdst = w_sm + (n_norm & 1) * P.norm_k;
        for (int i = et * 8; i < ph.K; i += CH_EPI_THREADS * 8)
          *reinterpret_cast<uint4*>(wdst + i) = __ldg(reinterpret_cast<const uint4*>(ph.n_w + i));
        ++n_norm;
        ch_epi_bar();
      }
      float rstd = 0.f;
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
        const uint32_t slot = g % P.stages, par = (g / P.stages) & 1;
        mbar_wait(&full[slot], par);
        if (et == 0) trace_fine(2, g);
        if (norm) {
          if (kb == it.kb0 && ct < P.T) {
            // the phase's input is complete grid-wide (its tiles have just landed): so are its row statistics
            float ss = 0.f;
            for (int i = 0; i < ph.n_ssq_parts; ++i) ss += __ldcg(&ph.n_ssq[static_cast<size_t>(i) * P.T + ct]);
            rstd = rsqrtf(ss / static_cast<float>(ph.K) + ph.n_eps);
          }
          uint8_t* b = smem + slot * slot_bytes + CH_A_BYTES + crow_off;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < chunks) {
              uint4* px = reinterpret_cast<uint4*>(b + (((cc0 + c) ^ (ct & 7)) << 4));
              const uint4 xv = *px;
              const uint4 wv = *reinterpret_cast<const uint4*>(wbuf + kb * CH_BLOCK_K + (cc0 + c) * 8);
              const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
              const uint32_t ws_[4] = {wv.x, wv.y, wv.z, wv.w};
              uint32_t r[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                r[j] = pack_bf16(bf16_lo(xs[j]) * rstd * bf16_lo(ws_[j]), bf16_hi(xs[j]) * rstd * bf16_hi(ws_[j]));
              *px = make_uint4(r[0], r[1], r[2], r[3]);
            }
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfull[slot]);
        if (et == 0) trace_fine(5, g);
      }
      // ---- accumulator complete ----
      mbar_wait(tmem_full, n_done & 1);
      ++n_done;
      tc_fence_after();
      if (et == 0 && trace_block0()) trace_mark(40 + p);
      if (ph.kind == CK_SILU) {
        const int h = ph.tile_rows >> 1;
        const bool is_gate = row < h, is_up = row >= h && row < 2 * h;
        constexpr int ldx = 17;
        for (int c0 = 0; c0 < P.t_tile; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
          if (is_up) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xchg[(row - h) * ldx + j] = round_bf16(__uint_as_float(v[j]));
          }
          ch_epi_bar();
          if (is_gate) {
            const int n_out = it.tile * h + row;
            if (n_out < ph.n_out) {
              __nv_bfloat16* y = static_cast<__nv_bfloat16*>(ph.y);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int t = c0 + j;
                if (t < P.T) {
                  const float gv = round_bf16(__uint_as_float(v[j]));
                  const float sv = round_bf16(gv / (1.0f + expf(-gv)));
                  y[static_cast<size_t>(t) * ph.n_out + n_out] = __float2bfloat16_rn(sv * xchg[row * ldx + j]);
                }
              }
            }
          }
          ch_epi_bar();
        }
      } else {
        // ---- split-K: park the partial tile in the L2-resident workspace, meet the other splits of the tile ----
        const int S = ph.split_k;
        float* mine = P.ws + static_cast<size_t>(blockIdx.x) * P.t_tile * 128;
        for (int c0 = 0; c0 < P.t_tile; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) __stcg(&mine[(c0 + j) * 128 + row], __uint_as_float(v[j]));
        }
        if (S > 1) {
          __threadfence();
          ch_epi_bar();
          if (et == 0) {
            unsigned* tflag = &P.flags[p * stride + 1 + it.tile];
            red_release_gpu(tflag, 1u);
            spin_until(tflag, static_cast<unsigned>(S));
          }
        }
        ch_epi_bar();
        const float* part = P.ws + static_cast<size_t>(it.tile) * S * P.t_tile * 128;
        const int rank = it.split;
        const int n = it.tile * ph.tile_rows + row;
        const bool valid = row < ph.tile_rows && n < ph.N;
        const int n_mine = (P.T > rank) ? (P.T - rank + S - 1) / S : 0;
        for (int i0 = 0; i0 < n_mine; i0 += CH_CHUNK) {
          float a[CH_CHUNK];
#pragma unroll
          for (int i = 0; i < CH_CHUNK; ++i) a[i] = 0.f;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < S) {
#pragma unroll
              for (int i = 0; i < CH_CHUNK; ++i) {
                const int t = rank + (i0 + i) * S;
                const float v = (i0 + i < n_mine)
                                    ? __ldcg(&part[(static_cast<size_t>(s) * P.t_tile + t) * 128 + row]) : 0.f;
                a[i] = (s == 0) ? v : a[i] + v;
              }
            }
          }
          if (ph.kind == CK_RESID) {
            __nv_bfloat16* hid = static_cast<__nv_bfloat16*>(ph.y);
#pragma unroll
            for (int i = 0; i < CH_CHUNK; ++i) {
              const int t = rank + (i0 + i) * S;
              float sq = 0.f;
              if (valid && i0 + i < n_mine) {
                const size_t idx = static_cast<size_t>(t) * ph.N + n;
                float hv = round_bf16(a[i]);
                if (ph.residual) {
                  const unsigned short rb = __ldcg(reinterpret_cast<const unsigned short*>(ph.residual) + idx);
                  hv = round_bf16(__uint_as_float(static_cast<uint32_t>(rb) << 16) + hv);
                }
                hid[idx] = __float2bfloat16_rn(hv);
                sq = hv * hv;
              }
              sq = warp_sum(sq);
              if (lane == 0) ssq_sm[quarter * CH_CHUNK + i] = sq;
            }
            ch_epi_bar();
            if (et < CH_CHUNK && i0 + et < n_mine && ph.ssq_out)
              ph.ssq_out[static_cast<size_t>(it.tile) * P.T + rank + (i0 + et) * S] =
                  (ssq_sm[et] + ssq_sm[CH_CHUNK + et]) + (ssq_sm[2 * CH_CHUNK + et] + ssq_sm[3 * CH_CHUNK + et]);
            ch_epi_bar();
          } else {
            // CK_ROPE: the tile is one head, the row is the element e of that head
            constexpr int ldx = CH_CHUNK + 1;
            const int D = ph.tile_rows, half = D >> 1;
            const int head = it.tile;
            const bool rot = head < P.n_q + P.n_kv;
#pragma unroll
            for (int i = 0; i < CH_CHUNK; ++i) xchg[row * ldx + i] = round_bf16(a[i]);
            ch_epi_bar();
            if (row < D) {
              const int prow = row < half ? row + half : row - half;
              const float sign = row < half ? -1.f : 1.f;
#pragma unroll
              for (int i = 0; i < CH_CHUNK; ++i) {
                const int t = rank + (i0 + i) * S;
                if (i0 + i < n_mine) {
                  float v = xchg[row * ldx + i];
                  if (rot) {
                    const float* cs = P.rope_cs + static_cast<size_t>(t) * 2 * D;
                    v = v * cs[row] + sign * xchg[prow * ldx + i] * cs[D + row];
                  }
                  const __nv_bfloat16 o = __float2bfloat16_rn(v);
                  if (head < P.n_q) {
                    static_cast<__nv_bfloat16*>(ph.y)[(static_cast<size_t>(t) * P.n_q + head) * D + row] = o;
                  } else {
                    const int page = P.row_page[t];
                    if (page >= 0) {
                      const size_t row_elems = static_cast<size_t>(P.n_kv) * D;
                      const size_t slab = static_cast<size_t>(P.page_size) * row_elems;
                      const int hk = head - P.n_q;
                      const int is_v = hk >= P.n_kv ? 1 : 0;
                      ph.kv[(static_cast<size_t>(page) * 2 + is_v) * slab + static_cast<size_t>(P.row_slot[t]) * row_elems +
                            static_cast<size_t>(hk - is_v * P.n_kv) * D + row] = o;
                    }
                  }
                }
              }
            }
            ch_epi_bar();
          }
        }
      }
      // ---- this CTA's share of the phase is in global memory: arrive at the grid-wide counter ----
      tc_fence_before();
      __threadfence();
      ch_epi_bar();
      if (et == 0) red_release_gpu(gbar, 1u);
      if (et == 0 && trace_block0()) trace_mark(50 + p);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
  trace_end(tr);
}

}  // namespace vb

using namespace vb;

extern "C" {

size_t vb_decode_chain_workspace_bytes(int max_items) {
  return static_cast<size_t>(max_items > 0 ? max_items : 0) * 64 * 128 * sizeof(float) + 256;
}

size_t vb_decode_chain_flags_bytes(int max_tiles) {
  return static_cast<size_t>(CH_MAX_PHASES) * (1 + (max_tiles > 0 ? max_tiles : 0)) * sizeof(unsigned int);
}

int vb_decode_chain(const vb_chain_phase* phases, int n_phases, int T, const float* d_rope_cs,
                    const int32_t* d_row_page, const int32_t* d_row_slot, int n_q, int n_kv, int page_size,
                    void* d_workspace, size_t workspace_bytes, void* d_flags, size_t flags_bytes, int max_tiles,
                    void* stream) {
  VB_CHECK_ARG(phases && d_workspace && d_flags, "vb_decode_chain: null pointer");
  VB_CHECK_ARG(n_phases >= 1 && n_phases <= CH_MAX_PHASES, "vb_decode_chain: %d phases (1..%d)", n_phases, CH_MAX_PHASES);
  VB_CHECK_ARG(T > 0 && T <= 64, "vb_decode_chain: T %d outside (0, 64] (decode-sized batches only)", T);
  VB_CHECK_ARG(flags_bytes >= vb_decode_chain_flags_bytes(max_tiles), "vb_decode_chain: flag buffer too small");
  int sms = 0, dev = 0;
  VB_CHECK_CUDA(cudaGetDevice(&dev));
  VB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ChainParams P = {};
  P.n_phases = n_phases; P.T = T; P.t_tile = T <= 16 ? 16 : (T <= 32 ? 32 : 64);
  P.ws = static_cast<float*>(d_workspace); P.flags = static_cast<unsigned int*>(d_flags);
  P.rope_cs = d_rope_cs; P.row_page = d_row_page; P.row_slot = d_row_slot;
  P.n_q = n_q; P.n_kv = n_kv; P.page_size = page_size; P.max_tiles = max_tiles;
  const CUtensorMap* maps[CH_MAX_PHASES] = {nullptr, nullptr, nullptr, nullptr};
  int grid = 1, norm_k = 0;
  for (int i = 0; i < n_phases; ++i) {
    const vb_chain_phase& in = phases[i];
    ChainPhase& ph = P.ph[i];
    VB_CHECK_ARG(in.w_tiles && in.x_map && in.out, "vb_decode_chain: phase %d: null pointer", i);
    VB_CHECK_ARG(in.kind >= 0 && in.kind <= 2, "vb_decode_chain: phase %d: kind %d", i, in.kind);
    VB_CHECK_ARG(in.tile_rows >= 8 && in.tile_rows <= 128 && in.tile_rows % 8 == 0,
                 "vb_decode_chain: phase %d: tile_rows %d", i, in.tile_rows);
    VB_CHECK_ARG(in.K > 0 && in.K % 8 == 0 && in.N > 0, "vb_decode_chain: phase %d: N %d K %d", i, in.N, in.K);
    const int num_kb = (in.K + CH_BLOCK_K - 1) / CH_BLOCK_K;
    VB_CHECK_ARG(in.split_k >= 1 && in.split_k <= 8 && in.split_k <= num_kb, "vb_decode_chain: phase %d: split_k %d", i,
                 in.split_k);
    ph.kind = in.kind; ph.N = in.N; ph.K = in.K; ph.tile_rows = in.tile_rows; ph.split_k = in.split_k;
    ph.n_tiles = (in.N + in.tile_rows - 1) / in.tile_rows;
    ph.n_out = in.n_out > 0 ? in.n_out : in.N;
    ph.w_tiles = static_cast<const uint8_t*>(in.w_tiles);
    ph.y = in.out;
    ph.residual = static_cast<const __nv_bfloat16*>(in.residual);
    ph.ssq_out = in.ssq_out;
    ph.n_ssq = in.ssq_in; ph.n_ssq_parts = in.n_ssq_parts; ph.n_w = static_cast<const __nv_bfloat16*>(in.norm_weight);
    ph.n_eps = in.eps;
    ph.kv = static_cast<__nv_bfloat16*>(in.layer_kv);
    if (in.kind != CK_RESID) {
      VB_CHECK_ARG(in.ssq_in && in.norm_weight && in.n_ssq_parts > 0 && in.K % 64 == 0,
                   "vb_decode_chain: phase %d: norm inputs missing or K %% 64 != 0", i);
      VB_CHECK_ARG(in.split_k == 1 || in.kind == CK_ROPE, "vb_decode_chain: phase %d: the gate/up phase cannot split K", i);
      norm_k = in.K > norm_k ? in.K : norm_k;
    }
    if (in.kind == CK_SILU)
      VB_CHECK_ARG(in.tile_rows % 16 == 0 && in.N % in.tile_rows == 0, "vb_decode_chain: phase %d: gate/up tile_rows %d", i,
                   in.tile_rows);
    if (in.kind == CK_ROPE) {
      VB_CHECK_ARG(in.layer_kv && d_rope_cs && d_row_page && d_row_slot && (in.tile_rows == 64 || in.tile_rows == 128) &&
                       in.N == (n_q + 2 * n_kv) * in.tile_rows,
                   "vb_decode_chain: phase %d: QKV phase needs the cache, the rope table and one head per tile", i);
    }
    VB_CHECK_ARG(ph.n_tiles <= max_tiles, "vb_decode_chain: phase %d: %d tiles exceed max_tiles %d", i, ph.n_tiles, max_tiles);
    const int items = ph.n_tiles * ph.split_k;
    VB_CHECK_ARG(items <= sms, "vb_decode_chain: phase %d: %d work items exceed the %d SMs (all CTAs must be co-resident)",
                 i, items, sms);
    VB_CHECK_ARG(workspace_bytes >= vb_decode_chain_workspace_bytes(items), "vb_decode_chain: workspace too small");
    grid = items > grid ? items : grid;
    maps[i] = static_cast<const CUtensorMap*>(in.x_map);
  }
  for (int i = n_phases; i < CH_MAX_PHASES; ++i) maps[i] = maps[0];
  P.norm_k = (norm_k + 63) / 64 * 64;
  const int slot_bytes = CH_A_BYTES + P.t_tile * 128;
  const int fixed = 1024 /*barriers, ssq*/ + 128 * 17 * 4 + 2 * P.norm_k * 2 + 1024 /*alignment*/;
  int stages = (VB_MAX_DYN_SMEM - fixed) / slot_bytes;
  if (const char* e = getenv("VB_CHAIN_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v < stages) stages = v;
  }
  if (stages > 24) stages = 24;       // barrier block: 3 * 24 * 8 + 16 < 768
  VB_CHECK_ARG(stages >= 2, "vb_decode_chain: shared memory too small for a 2-slot ring");
  P.stages = stages;
  const int smem = stages * slot_bytes + fixed;
  VB_CHECK_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  VB_LAUNCH_PDL(chain_kernel, grid, CH_THREADS, smem, stream, P, *maps[0], *maps[1], *maps[2], *maps[3]);
  return 0;
}

}  // extern "C"

VB_DEFINE_TRACE_SETTER(chain)
