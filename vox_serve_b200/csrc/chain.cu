// Persistent projection chain for decode-sized steps: up to four dependent projections of the transformer layer
//     O-proj + residual  ->  RMSNorm + gate/up + SiLU*up  ->  down-proj + residual  ->  RMSNorm + QKV + RoPE + KV append
// (vox_serve/model/orpheus.py:81-151, the part of the layer between two attention calls) as ONE launch of at most
// one CTA per SM.  What a chain of separate kernels cannot do, this one does: the weight stream never stops at a
// phase boundary.  Weights depend on nothing, so the producer warp keeps filling the shared-memory ring with the
// tiles of the NEXT phase while the current phase is still being reduced, normalised, written back and
// synchronised; the boundary costs (grid-wide dependency, activation reload, epilogue) are paid from the ~180 KiB
// of weights already on chip instead of from an idle HBM.
//
// Roles (192 threads): warp 0 = weight producer (linear bulk copies of pre-tiled, pre-swizzled weight tiles, see
// vb_pack_weight_tiles -- the only user of the TMA unit, whose issue rate of ~48 GB/s per SM is what bounds the
// stream), warp 1 = tcgen05.mma issuer (accumulators in TMEM), warps 2-5 = activation loaders / finishers and
// epilogue: they pull the phase's input tile K-block by K-block with 16-byte cp.async copies (LSU path, four
// K-blocks in flight, swizzled on the way in), apply the RMSNorm in place where the phase has one, and hand
// the tile to the MMA.
// Phases are separated by a grid-wide arrival counter in global memory (all CTAs are co-resident: grid <= SMs,
// one CTA per SM); split-K partial tiles are exchanged through an L2-resident workspace, every CTA of a tile
// finishing the tokens t = split, split + S, ... in split order (deterministic).
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

constexpr int CH_BLOCK_K = 64;
constexpr int CH_THREADS = 192;
constexpr int CH_XDEPTH = 2;             // activation tiles in flight per finisher warp (cp.async groups)
constexpr int CH_EPI_THREADS = 128;
constexpr int CH_MAX_PHASES = 4;
constexpr int CH_CHUNK = 8;
constexpr int CH_A_BYTES = 16384;        // every ring slot reserves a full 128-row operand (the MMA reads M = 128)

enum { CK_RESID = 0, CK_SILU = 1, CK_ROPE = 2 };

struct ChainPhase {
  const uint8_t* w_tiles;
  const __nv_bfloat16* x;   // phase input [T][K] row-major (leading dimension ldx)
  void* y;
  const float* n_ssq;
  const __nv_bfloat16* n_w;
  const __nv_bfloat16* residual;
  float* ssq_out;
  __nv_bfloat16* kv;
  int kind, N, K, tile_rows, n_tiles, split_k, n_out, n_ssq_parts;
  float n_eps;
  int ldx;
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  // src-size 0 = zero fill (rows past T, columns past K)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(pred ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int T_TILE>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int slot_bytes = CH_A_BYTES + T_TILE * 128;
  uint8_t* tail = smem + P.stages * slot_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);           // the weight tile of the slot has landed
  uint64_t* empty = full + P.stages;                            // the MMAs have read the slot
  uint64_t* cfull = empty + P.stages;                           // the activation tile is in place (normalised)
  uint64_t* tmem_full = cfull + P.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* ssq_sm = reinterpret_cast<float*>(tail + 768);         // [4][CH_CHUNK]
  float* xchg = reinterpret_cast<float*>(tail + 1024);          // [128][17]
  __nv_bfloat16* w_sm = reinterpret_cast<__nv_bfloat16*>(tail + 1024 + 128 * 17 * 4);   // [2][norm_k]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = 1 + P.max_tiles;
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(3, P.n_phases) : -1;
  // (dev) per-slot marks by one thread of each role
  unsigned long long* fine = (lane == 0 && (warp <= 2)) ? trace_fine_base() : nullptr;   // finisher marks: warp 2 only

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&cfull[s], 1);
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
      uint32_t g = 0, slot = 0, par = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        const ChItem it = ch_item(ph);
        if (!it.has) continue;
        const uint32_t a_bytes = ph.tile_rows * 128;
        const uint8_t* src = ph.w_tiles + (static_cast<size_t>(it.tile) * it.num_kb + it.kb0) * a_bytes;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          mbar_wait(&empty[slot], par ^ 1);
          trace_fine(fine, 0, g);
          mbar_arrive_expect_tx(&full[slot], a_bytes);
          ch_bulk_g2s(smem + slot * slot_bytes, src + static_cast<size_t>(kb - it.kb0) * a_bytes, a_bytes, &full[slot],
                      pol_w);
          if (++slot == static_cast<uint32_t>(P.stages)) { slot = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, T_TILE, 1u);
      uint32_t g = 0, slot = 0, par = 0;
      for (int p = 0; p < P.n_phases; ++p) {
        const ChItem it = ch_item(P.ph[p]);
        if (!it.has) continue;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          mbar_wait(&cfull[slot], par);     // weights landed AND activation tile final (the finisher checked both)
          trace_fine(fine, 3, g);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + slot * slot_bytes);
          const uint64_t a_desc = umma_desc_sw128_kmajor(a_addr);
          const uint64_t b_desc = umma_desc_sw128_kmajor(a_addr + CH_A_BYTES);
#pragma unroll
          for (int k = 0; k < CH_BLOCK_K / 16; ++k)
            umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[slot]);
          trace_fine(fine, 4, g);
          if (++slot == static_cast<uint32_t>(P.stages)) { slot = 0; par ^= 1; }
        }
        umma_commit(tmem_full);
      }
    }
  } else {
    // ================= activation loaders / finishers + epilogue =================
    const int et = threadIdx.x - 64;
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // activation tiles: finisher warp cw owns the ring slots g = cw (mod 4); inside a tile a lane owns cpl
    // consecutive 16-byte chunks of the row-major (token, chunk) order -- whole 128-byte token rows for t_tile 32
    const int cw = warp - 2;
    constexpr int cpl = T_TILE / 4;                         // chunks per lane: 4, 8, 16
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
      }
      // ---- the phase's input must be complete grid-wide ----
      if (p == 0) {
        pdl_wait();             // the kernel in front of the chain (attention / embedding) is complete
        if (et == 0) pdl_trigger();
      } else {
        if (et == 0) spin_until(&P.flags[(p - 1) * stride], gridDim.x);
      }
      trace_fine(fine, 8, p * 8 + 6);
      ch_epi_bar();             // (also: w_sm staged by all)
      if (et == 0 && trace_block0()) trace_mark(30 + p);
      const uint32_t g0 = g, g1 = g + static_cast<uint32_t>(it.kb1 - it.kb0);
      auto issue = [&](uint32_t gg) {
        const uint32_t slot = gg % P.stages, par = (gg / P.stages) & 1;
        const int kb = it.kb0 + static_cast<int>(gg - g0);
        mbar_wait(&empty[slot], par ^ 1);
        const uint32_t b = smem_u32(smem + slot * slot_bytes + CH_A_BYTES);
#pragma unroll
        for (int ci = 0; ci < cpl; ++ci) {
          const int idx = lane * cpl + ci, t = idx >> 3, c = idx & 7;
          const int k = kb * CH_BLOCK_K + c * 8;
          const bool ok = t < P.T && k < ph.K;
          cp_async16(b + t * 128 + ((c ^ (t & 7)) << 4), ph.x + (ok ? static_cast<size_t>(t) * ph.ldx + k : 0), ok);
        }
      };
      const uint32_t my0 = g0 + ((static_cast<uint32_t>(cw) - g0) & 3u);
#pragma unroll
      for (int i = 0; i < CH_XDEPTH; ++i) {
        if (my0 + 4u * i < g1) issue(my0 + 4u * i);
        cp_async_commit();
      }
      float rstd0 = 0.f, rstd1 = 0.f;
      float res[CH_CHUNK];
      if (norm) {
        // row statistics of the input: per-tile partial sums written by the producing phase (all loads in flight)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int t = (lane * cpl) / 8 + h2;
          float ss = 0.f;
          if (t < P.T && (h2 == 0 || cpl > 8)) {
            for (int i0 = 0; i0 < ph.n_ssq_parts; i0 += 8) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[i] = (i0 + i < ph.n_ssq_parts) ? __ldcg(&ph.n_ssq[static_cast<size_t>(i0 + i) * P.T + t]) : 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) ss += v[i];
            }
            const float r = rsqrtf(ss / static_cast<float>(ph.K) + ph.n_eps);
            if (h2 == 0) rstd0 = r; else rstd1 = r;
          }
        }
      } else {
        // residual values of the tokens this CTA will finish: fetched now, used after the reduction
        const int n = it.tile * ph.tile_rows + row;
#pragma unroll
        for (int i = 0; i < CH_CHUNK; ++i) {
          const int t = it.split + i * ph.split_k;
          res[i] = 0.f;
          if (ph.residual && t < P.T && row < ph.tile_rows && n < ph.N) {
            const unsigned short rb =
                __ldcg(reinterpret_cast<const unsigned short*>(ph.residual) + static_cast<size_t>(t) * ph.N + n);
            res[i] = __uint_as_float(static_cast<uint32_t>(rb) << 16);
          }
        }
      }
      for (uint32_t gg = my0; gg < g1; gg += 4u) {
        const uint32_t slot = gg % P.stages, par = (gg / P.stages) & 1;
        const int kb = it.kb0 + static_cast<int>(gg - g0);
        cp_async_wait<CH_XDEPTH - 1>();      // this lane's chunks of the tile are in shared memory
        trace_fine(fine, 2, gg);
        if (norm) {
          uint8_t* b = smem + slot * slot_bytes + CH_A_BYTES;
#pragma unroll
          for (int ci = 0; ci < cpl; ++ci) {
            const int idx = lane * cpl + ci, t = idx >> 3, c = idx & 7;
            const float rstd = (ci < 8) ? rstd0 : rstd1;
            uint4* px = reinterpret_cast<uint4*>(b + t * 128 + ((c ^ (t & 7)) << 4));
            const uint4 xv = *px;
            const uint4 wv = *reinterpret_cast<const uint4*>(wbuf + kb * CH_BLOCK_K + c * 8);
            const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
            const uint32_t ws_[4] = {wv.x, wv.y, wv.z, wv.w};
            uint32_t r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              r[j] = pack_bf16(bf16_lo(xs[j]) * rstd * bf16_lo(ws_[j]), bf16_hi(xs[j]) * rstd * bf16_hi(ws_[j]));
            *px = make_uint4(r[0], r[1], r[2], r[3]);
          }
        }
        trace_fine(fine, 1, gg);
        fence_proxy_async();      // generic-proxy writes (cp.async, st.shared) -> the tensor core's async-proxy reads
        if (lane == 0) mbar_wait(&full[slot], par);      // the slot's weight tile has landed too
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfull[slot]);
        trace_fine(fine, 5, gg);
        if (gg + 4u * CH_XDEPTH < g1) issue(gg + 4u * CH_XDEPTH);
        trace_fine(fine, 7, gg);
        cp_async_commit();
      }
      g = g1;
      // ---- accumulator complete ----
      mbar_wait(tmem_full, n_done & 1);
      ++n_done;
      tc_fence_after();
      if (et == 0 && trace_block0()) trace_mark(40 + p);
      trace_fine(fine, 8, p * 8 + 0);
      if (ph.kind == CK_SILU) {
        // rows packed per warp: lanes 0..15 gate, lanes 16..31 the matching up rows (ops.interleave_gate_up); lanes
        // 0..15 finish the even tokens of their output column, lanes 16..31 the odd ones
        const int h = ph.tile_rows >> 1;
        const bool hi = lane >= 16;
        const int n_out = it.tile * h + quarter * 16 + (lane & 15);
        const bool live_o = row < ph.tile_rows && n_out < ph.n_out;
        __nv_bfloat16* y = static_cast<__nv_bfloat16*>(ph.y) + n_out;
        for (int c0 = 0; c0 < T_TILE; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float mine = __uint_as_float(v[2 * i + (hi ? 1 : 0)]);
            const float send = __uint_as_float(v[2 * i + (hi ? 0 : 1)]);
            const float other = __shfl_xor_sync(0xffffffffu, send, 16);
            const float gv = round_bf16(hi ? other : mine);
            const float up = round_bf16(hi ? mine : other);
            const float sv = round_bf16(gv / (1.0f + expf(-gv)));
            o[i] = sv * up;
          }
          if (live_o) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int t = c0 + 2 * i + (hi ? 1 : 0);
              if (t < P.T) y[static_cast<size_t>(t) * ph.n_out] = __float2bfloat16_rn(o[i]);
            }
          }
        }
      } else {
        // ---- split-K: park the partial tile in the L2-resident workspace, meet the other splits of the tile ----
        const int S = ph.split_k;
        float* mine = P.ws + static_cast<size_t>(blockIdx.x) * T_TILE * 128;
        for (int c0 = 0; c0 < T_TILE; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) __stcg(&mine[(c0 + j) * 128 + row], __uint_as_float(v[j]));
        }
        trace_fine(fine, 8, p * 8 + 1);
        ch_epi_bar();             // every thread's partial stores happen-before thread 0's release below
        if (S > 1 && et == 0) {
          unsigned* tflag = &P.flags[p * stride + 1 + it.tile];
          red_release_gpu(tflag, 1u);
          trace_fine(fine, 8, p * 8 + 2);
          spin_until(tflag, static_cast<unsigned>(S));
          trace_fine(fine, 8, p * 8 + 3);
        }
        ch_epi_bar();
        const float* part = P.ws + static_cast<size_t>(it.tile) * S * T_TILE * 128;
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
                                    ? __ldcg(&part[(static_cast<size_t>(s) * T_TILE + t) * 128 + row]) : 0.f;
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
                  float rv = res[i];
                  if (i0 > 0) {       // beyond the prefetched chunk (more than CH_CHUNK tokens per CTA)
                    const unsigned short rb = __ldcg(reinterpret_cast<const unsigned short*>(ph.residual) + idx);
                    rv = __uint_as_float(static_cast<uint32_t>(rb) << 16);
                  }
                  hv = round_bf16(rv + hv);
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
            // rotation table and page slots of this chunk's tokens: independent loads, issued together
            float cv[CH_CHUNK], sv[CH_CHUNK];
            int pg[CH_CHUNK], sl[CH_CHUNK];
#pragma unroll
            for (int i = 0; i < CH_CHUNK; ++i) {
              const int t = rank + (i0 + i) * S;
              const bool ok = i0 + i < n_mine && row < D;
              cv[i] = (ok && rot) ? P.rope_cs[static_cast<size_t>(t) * 2 * D + row] : 1.f;
              sv[i] = (ok && rot) ? P.rope_cs[static_cast<size_t>(t) * 2 * D + D + row] : 0.f;
              pg[i] = (ok && head >= P.n_q) ? P.row_page[t] : -1;
              sl[i] = (ok && head >= P.n_q) ? P.row_slot[t] : 0;
            }
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
                  if (rot) v = v * cv[i] + sign * xchg[prow * ldx + i] * sv[i];
                  const __nv_bfloat16 o = __float2bfloat16_rn(v);
                  if (head < P.n_q) {
                    static_cast<__nv_bfloat16*>(ph.y)[(static_cast<size_t>(t) * P.n_q + head) * D + row] = o;
                  } else if (pg[i] >= 0) {
                    const size_t row_elems = static_cast<size_t>(P.n_kv) * D;
                    const size_t slab = static_cast<size_t>(P.page_size) * row_elems;
                    const int hk = head - P.n_q;
                    const int is_v = hk >= P.n_kv ? 1 : 0;
                    ph.kv[(static_cast<size_t>(pg[i]) * 2 + is_v) * slab + static_cast<size_t>(sl[i]) * row_elems +
                          static_cast<size_t>(hk - is_v * P.n_kv) * D + row] = o;
                  }
                }
              }
            }
            ch_epi_bar();
          }
        }
      }
      // ---- this CTA's share of the phase is in global memory: arrive at the grid-wide counter ----
      trace_fine(fine, 8, p * 8 + 4);
      tc_fence_before();
      ch_epi_bar();               // all epilogue stores happen-before the release (cumulative)
      if (et == 0) red_release_gpu(gbar, 1u);
      trace_fine(fine, 8, p * 8 + 5);
      if (et == 0 && trace_block0()) trace_mark(50 + p);
    }
    cp_async_wait<0>();
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
  int grid = 1, norm_k = 0;
  for (int i = 0; i < n_phases; ++i) {
    const vb_chain_phase& in = phases[i];
    ChainPhase& ph = P.ph[i];
    VB_CHECK_ARG(in.w_tiles && in.x && in.out, "vb_decode_chain: phase %d: null pointer", i);
    VB_CHECK_ARG(in.ldx >= in.K && in.ldx % 8 == 0, "vb_decode_chain: phase %d: ldx %d (>= K, multiple of 8)", i, static_cast<int>(in.ldx));
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
    ph.x = static_cast<const __nv_bfloat16*>(in.x); ph.ldx = static_cast<int>(in.ldx);
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
      VB_CHECK_ARG(in.tile_rows % 32 == 0 && in.N % in.tile_rows == 0, "vb_decode_chain: phase %d: gate/up tile_rows %d", i,
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
  }
  P.norm_k = (norm_k + 63) / 64 * 64;
  const int slot_bytes = CH_A_BYTES + P.t_tile * 128;
  const int fixed = 1024 /*barriers, ssq*/ + 128 * 17 * 4 + 2 * P.norm_k * 2 + 1024 /*alignment*/;
  int stages = (VB_MAX_DYN_SMEM - fixed) / slot_bytes;
  if (const char* e = getenv("VB_CHAIN_STAGES")) {
    const int v = atoi(e);
    if (v >= 5 && v < stages) stages = v;
  }
  if (stages > 24) stages = 24;       // barrier block: 3 * 24 * 8 + 16 < 768
  VB_CHECK_ARG(stages >= 5, "vb_decode_chain: shared memory too small for the ring (%d slots)", stages);
  P.stages = stages;
  const int smem = stages * slot_bytes + fixed;
  auto kern = P.t_tile == 16 ? chain_kernel<16> : (P.t_tile == 32 ? chain_kernel<32> : chain_kernel<64>);
  VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_MAX_DYN_SMEM));
  VB_LAUNCH_PDL(kern, grid, CH_THREADS, smem, stream, P);
  return 0;
}

}  // extern "C"

VB_DEFINE_TRACE_SETTER(chain)
