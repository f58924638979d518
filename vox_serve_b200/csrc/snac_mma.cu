// SNAC decoder, the two GEMM-shaped stages (1x1 convs, phase-decomposed transposed convs; snac.py:170-231) on the
// tcgen05 tensor cores with fp32-grade results: every fp32 operand is split into a tf32-exact high part and the
// remainder (x = hi + lo, hi = x with the 13 low mantissa bits cleared), and the product is accumulated as
//   lo_w * hi_x + hi_w * lo_x + hi_w * hi_x        (kind::tf32, fp32 accumulators in TMEM)
// which leaves a relative error of ~2^-21 per product -- the waveform stays inside the 1e-3 the contract allows by
// three orders of magnitude (the parity test holds it to 2e-4 absolute against the fp32 oracle).
//
//   D[m][col] = sum_k W[m][k] * X[k][col],  m = output channel, col = (window b, position n) FOLDED, so a dozen
//   positions per window still fill 128-column tiles and a weight tile is fetched once per 128 columns of the batch.
//
// A (weights): packed once at load by vb_snac_pack_tf32x3 as [phase][m_tile][k_block][hi|lo][128 rows][32 fp32],
//   K-major rows of 128 bytes with the 128-byte UMMA swizzle applied -> one contiguous 32 KiB bulk copy per stage.
// B (activations [B][C][T], T contiguous): eight loader warps read 32 channels x 128 columns per stage with
//   coalesced loads (a warp = 32 consecutive positions of one channel), split hi / lo in registers and store both
//   tiles K-major + swizzled; two groups of four warps alternate stages so one group's L2 round trip hides behind
//   the other's stores.  For the transposed conv the second half of K reads the same rows shifted by one position.
// One MMA thread issues 12 x (128 x 128 x 8) per stage.  Epilogue: TMEM -> registers -> per-warp shared-memory
// transpose -> lanes along the position axis, so bias / residual / noise / Snake and the store are coalesced.
#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {

constexpr int SM_THREADS = 320;          // warp 0 weight producer, warp 1 MMA issuer, warps 2..9 loaders + epilogue
constexpr int SM_BLOCK_K = 32;           // fp32 elements per 128-byte row
constexpr int SM_TILE_N = 128;           // columns per CTA (UMMA N)
constexpr int SM_A_BYTES = 2 * 128 * 128;        // hi + lo weight tiles
constexpr int SM_B_BYTES = 2 * SM_TILE_N * 128;  // hi + lo activation tiles
constexpr int SM_STAGE = SM_A_BYTES + SM_B_BYTES;
constexpr int SM_STAGES = 3;
constexpr int SM_SMEM = SM_STAGES * SM_STAGE + 1024 /*barriers*/ + 1024 /*alignment*/;

struct SnacGemmParams {
  const uint8_t* w_tiles;
  const float* x;
  float* y;
  const float* bias;
  const float* resid;
  const float* noise;
  const float* alpha_out;
  int kind;      // 0 pointwise conv, 1 transposed conv (grid.z = phase), 2 causal conv with ksize taps (K = ksize * Cin, tap-major)
  int epi;       // 0 plain, 1 + resid, 2 noise block, 3 resid + scale[m] v, 4 GELU, 5 SiLU, 6 resid * v, 7 clamp(-1, 1)
  int B, Cin, Cout, T, K;
  int n_lo, nr, n_total;
  int s, pad;
  // streaming codecs (codec.cu): left context instead of zeros, input activation in the B-operand fetch
  const float* ctx;      // kind 2: [B][Cin][(ksize-1) dil]; kind 1: [B][Cin][1]; NULL = zeros.  Holds ACTIVATED values
  const float* act_a;    // SnakeBeta tables per input channel (act_in == 2)
  const float* act_ib;
  const float* scale;    // epi 3
  int act_in;            // 0 none, 1 ELU, 2 SnakeBeta
  int ksize, dil;        // kind 2
};

__device__ __forceinline__ float snake_tc(float x, float alpha) {
  const float s = sinf(alpha * x);
  return x + (1.0f / (alpha + 1e-9f)) * s * s;
}

__device__ __forceinline__ void snac_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(SM_THREADS, 1) snac_gemm_tf32x3_kernel(const SnacGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM_STAGES * SM_STAGE);   // weights landed + B tile written
  uint64_t* empty = full + SM_STAGES;                                           // the MMAs have read the stage
  uint64_t* tmem_full = empty + SM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * SM_TILE_N, m_tile = blockIdx.y, phase = blockIdx.z;
  const int tr = (threadIdx.x == 0 && trace_block0()) ? trace_begin(60, p.K / SM_BLOCK_K) : -1;
  const int num_kb = p.K / SM_BLOCK_K;
  const int m_tiles = (p.Cout + 127) >> 7;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SM_STAGES; ++s) {
      mbar_init(&full[s], 1 + 4);      // the producer's expect_tx arrival + one arrival per loader warp of a group
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, SM_TILE_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr >= 0) trace_mark(61);

  if (warp == 0) {
    // ================= weight producer: one 32 KiB bulk copy per stage =================
    if (lane == 0) {
      const uint8_t* wsrc = p.w_tiles + (static_cast<size_t>(phase) * m_tiles + m_tile) * num_kb * SM_A_BYTES;
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % SM_STAGES;
        if (it >= SM_STAGES) mbar_wait(&empty[s], ((it / SM_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], SM_A_BYTES);
        snac_bulk_g2s(smem + s * SM_STAGE, wsrc + static_cast<size_t>(it) * SM_A_BYTES, SM_A_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, SM_TILE_N, 2u);      // tf32 operands
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % SM_STAGES;
        mbar_wait(&full[s], (it / SM_STAGES) & 1);
        if (it == 0 && trace_block0()) trace_mark(62);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + s * SM_STAGE);
        const uint64_t a_hi = umma_desc_sw128_kmajor(base), a_lo = umma_desc_sw128_kmajor(base + 128 * 128);
        const uint64_t b_hi = umma_desc_sw128_kmajor(base + SM_A_BYTES);
        const uint64_t b_lo = umma_desc_sw128_kmajor(base + SM_A_BYTES + SM_TILE_N * 128);
#pragma unroll
        for (int k = 0; k < SM_BLOCK_K / 8; ++k) {
          // +32 bytes per 8-element K step inside the 128-byte swizzle atom (address field is >> 4); small terms first
          umma_tf32(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
          umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
      if (trace_block0()) trace_mark(63);
    }
  } else {
    // ================= activation loaders (two groups of four warps, alternating stages) =================
    const int lw = warp - 2, group = lw >> 2;
    const int j = (lw & 3) * 32 + lane;                    // column of the tile this thread fills
    {
      const int gcol = col0 + j;
      const bool colok = gcol < p.n_total;
      const int b = colok ? gcol / p.nr : 0;
      const int n = p.n_lo + gcol - b * p.nr;
      const float* xb = p.x + static_cast<size_t>(b) * p.Cin * p.T;
      const uint32_t row_off = static_cast<uint32_t>(j) * 128u;
      const int sw = j & 7;
      // software-pipelined: the 32 loads of this group's NEXT stage are in flight while the current one is split
      // and stored
      const int ctx_pad = p.kind == 2 ? (p.ksize - 1) * p.dil : 1;
      auto fetch = [&](int it, float (&v)[SM_BLOCK_K]) {
        const int kbase = it * SM_BLOCK_K;
        const int tap = p.kind == 2 ? kbase / p.Cin : ((p.kind == 1 && kbase >= p.Cin) ? 1 : 0);
        const int cb = kbase - tap * p.Cin;                 // first input channel of this k-block
        const int ti = n - (p.kind == 2 ? (p.ksize - 1 - tap) * p.dil : tap);
        const bool ok = it < num_kb && colok && ti >= 0 && ti < p.T;
        if (ok || !(p.ctx && it < num_kb && colok && ti < 0)) {
          const float* src = xb + static_cast<size_t>(cb) * p.T + ti;
#pragma unroll
          for (int kk = 0; kk < SM_BLOCK_K; ++kk) v[kk] = ok ? __ldg(src + static_cast<size_t>(kk) * p.T) : 0.f;
          if (ok && p.act_in == 2) {
#pragma unroll
            for (int kk = 0; kk < SM_BLOCK_K; ++kk) {
              const float sn = sinf(v[kk] * __ldg(&p.act_a[cb + kk]));
              v[kk] += __ldg(&p.act_ib[cb + kk]) * (sn * sn);
            }
          } else if (ok && p.act_in == 1) {
#pragma unroll
            for (int kk = 0; kk < SM_BLOCK_K; ++kk) v[kk] = v[kk] > 0.f ? v[kk] : expm1f(v[kk]);
          }
        } else {                                            // left context of the previous chunk (already activated)
          const float* src = p.ctx + (static_cast<size_t>(b) * p.Cin + cb) * ctx_pad + (ctx_pad + ti);
#pragma unroll
          for (int kk = 0; kk < SM_BLOCK_K; ++kk) v[kk] = __ldg(src + static_cast<size_t>(kk) * ctx_pad);
        }
      };
      auto store = [&](int it, const float (&v)[SM_BLOCK_K]) {
        const int s = it % SM_STAGES;
        if (it >= SM_STAGES) mbar_wait(&empty[s], ((it / SM_STAGES) & 1) ^ 1);
        uint8_t* bh = smem + s * SM_STAGE + SM_A_BYTES + row_off;
        uint8_t* bl = bh + SM_TILE_N * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = __uint_as_float(__float_as_uint(v[c * 4 + e]) & 0xffffe000u);
            lo[e] = v[c * 4 + e] - hi[e];
          }
          const uint32_t off = static_cast<uint32_t>((c ^ sw) << 4);
          *reinterpret_cast<float4*>(bh + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(bl + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      };
      float va[SM_BLOCK_K], vb_[SM_BLOCK_K];
      fetch(group, va);
      for (int it = group; it < num_kb; it += 4) {
        fetch(it + 2, vb_);
        store(it, va);
        if (it + 2 < num_kb) {
          fetch(it + 4, va);
          store(it + 2, vb_);
        }
      }
    }
    // ================= epilogue: TMEM -> registers -> per-warp transpose -> coalesced global =================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int half = lw >> 2;                     // columns [64*half, +64) of the tile
    float* st = reinterpret_cast<float*>(smem) + lw * (32 * 33);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int m_base = m_tile * 128 + quarter * 32;
    const int rows = max(0, min(32, p.Cout - m_base));
    const int Tout = p.kind == 1 ? p.T * p.s : p.T;
    const size_t ys = static_cast<size_t>(p.Cout) * Tout;
    // Rolled on purpose: unrolled over 2 chunks x 32 rows (with Snake inlined) this was ~80 KB of straight-line code
    // that every warp ran once -- instruction fetch, not arithmetic or memory, made the epilogue 10-37 us.  Rows go in
    // batches of 16: the batch's residual / skip operands are requested together (lane = column, coalesced), then
    // the batch is finished and stored.
    const float bias_l = (p.bias && lane < rows) ? __ldg(&p.bias[m_base + lane]) : 0.f;
    const float alpha_l = (p.alpha_out && lane < rows) ? __ldg(&p.alpha_out[m_base + lane]) : 1.f;
    const float scale_l = (p.epi == 3 && lane < rows) ? __ldg(&p.scale[m_base + lane]) : 1.f;
    const float* e = (p.epi == 1 || p.epi == 3 || p.epi == 6) ? p.resid : p.x;
    const bool need_ext = p.epi == 1 || p.epi == 2 || p.epi == 3 || p.epi == 6;
    mbar_wait(tmem_full, 0);          // all MMAs done: the accumulator is complete and the ring is idle
    tc_fence_after();
    if (warp == 2 && lane == 0 && trace_block0()) trace_mark(64);
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      const int c0 = half * 64 + ch * 32;
      if (col0 + c0 >= p.n_total) break;            // warp-uniform
      const int gcol = col0 + c0 + lane;
      const bool colok = gcol < p.n_total;
      const int b = colok ? gcol / p.nr : 0;
      const int n = p.n_lo + gcol - b * p.nr;
      int to = n;
      bool ok = colok;
      if (p.kind == 1) {
        to = n * p.s + phase - p.pad;
        ok = ok && to >= 0 && to < Tout;
      }
      const size_t obase = static_cast<size_t>(b) * ys + static_cast<size_t>(m_base) * Tout + to;
      const float nz = (p.epi == 2 && ok) ? p.noise[static_cast<size_t>(b) * p.T + n] : 0.f;
      uint32_t v[32];
      tmem_ld_32x32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) st[lane * 33 + c] = __uint_as_float(v[c]);
      __syncwarp();
#pragma unroll 1
      for (int r0 = 0; r0 < rows; r0 += 16) {
        float ext[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          ext[i] = (need_ext && ok && r0 + i < rows) ? __ldcg(e + obase + static_cast<size_t>(r0 + i) * Tout) : 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int r = r0 + i;
          const float bias_r = __shfl_sync(0xffffffffu, bias_l, r & 31);
          const float alpha_r = __shfl_sync(0xffffffffu, alpha_l, r & 31);
          const float scale_r = __shfl_sync(0xffffffffu, scale_l, r & 31);
          if (ok && r < rows) {
            float val = st[r * 33 + lane] + bias_r;
            switch (p.epi) {
              case 1: val += ext[i]; break;
              case 2: val = ext[i] + nz * val; break;
              case 3: val = ext[i] + scale_r * val; break;
              case 4: val = 0.5f * val * (1.f + erff(val * 0.70710678118654752f)); break;
              case 5: val = val / (1.f + expf(-val)); break;
              case 6: val = ext[i] * val; break;
              case 7: val = fminf(1.f, fmaxf(-1.f, val)); break;
              default: break;
            }
            if (p.alpha_out) val = snake_tc(val, alpha_r);
            p.y[obase + static_cast<size_t>(r) * Tout] = val;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, SM_TILE_N);
  trace_end(tr);
}

// fp32 weights [phases][M][K] -> [phase][m_tile][k_block][hi|lo][128][32], swizzled; rows past M are zero
__global__ void __launch_bounds__(256) snac_pack_tf32x3_kernel(float4* __restrict__ dst, const float* __restrict__ w,
                                                               int M, int K, int m_tiles, int num_kb) {
  const long long tile = blockIdx.x;            // (phase * m_tiles + m_tile) * num_kb + kb
  const int kb = static_cast<int>(tile % num_kb);
  const int m_tile = static_cast<int>((tile / num_kb) % m_tiles);
  const int phase = static_cast<int>(tile / num_kb / m_tiles);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int m = m_tile * 128 + r, k = kb * SM_BLOCK_K + c * 4;
    float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < M) {
      const float4 v = *reinterpret_cast<const float4*>(w + (static_cast<size_t>(phase) * M + m) * K + k);
      const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __uint_as_float(__float_as_uint(a[e]) & 0xffffe000u);
        lo[e] = a[e] - hi[e];
      }
    }
    const size_t base = static_cast<size_t>(tile) * (2 * 128 * 8);
    dst[base + r * 8 + (c ^ (r & 7))] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    dst[base + 128 * 8 + r * 8 + (c ^ (r & 7))] = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static int launch_snac_gemm(const SnacGemmParams& p, int phases, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(snac_gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
    attr_set = true;
  }
  const dim3 grid((p.n_total + SM_TILE_N - 1) / SM_TILE_N, (p.Cout + 127) / 128, phases);
  snac_gemm_tf32x3_kernel<<<grid, SM_THREADS, SM_SMEM, stream>>>(p);
  VB_CHECK_LAUNCH();
  return 0;
}

}  // namespace vb
VB_DEFINE_TRACE_SETTER(snac)

using namespace vb;

extern "C" {

long long vb_snac_tf32x3_bytes(int phases, int M, int K) {
  if (phases <= 0 || M <= 0 || K <= 0 || K % SM_BLOCK_K != 0) return -1;
  return static_cast<long long>(phases) * ((M + 127) / 128) * (K / SM_BLOCK_K) * SM_A_BYTES;
}

int vb_snac_pack_tf32x3(void* d_dst, const float* d_w, int phases, int M, int K, void* stream) {
  VB_CHECK_ARG(d_dst && d_w, "vb_snac_pack_tf32x3: null pointer");
  VB_CHECK_ARG(phases >= 1 && M >= 1 && K >= SM_BLOCK_K && K % SM_BLOCK_K == 0,
               "vb_snac_pack_tf32x3: K = %d must be a positive multiple of %d", K, SM_BLOCK_K);
  const int m_tiles = (M + 127) / 128, num_kb = K / SM_BLOCK_K;
  snac_pack_tf32x3_kernel<<<static_cast<unsigned>(static_cast<long long>(phases) * m_tiles * num_kb), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(static_cast<float4*>(d_dst), d_w, M, K, m_tiles, num_kb);
  VB_CHECK_LAUNCH();
  return 0;
}

int vb_snac_pwconv_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_resid,
                      const float* d_noise, const float* d_alpha_out, int epilogue, int B, int Cin, int Cout, int T,
                      int t_lo, int t_hi, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_tiles, "vb_snac_pwconv_tc: null pointer");
  VB_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "vb_snac_pwconv_tc: epilogue %d", epilogue);
  VB_CHECK_ARG(epilogue != 1 || d_resid, "vb_snac_pwconv_tc: residual epilogue needs d_resid");
  VB_CHECK_ARG(epilogue != 2 || (d_noise && Cin == Cout), "vb_snac_pwconv_tc: noise epilogue needs noise and Cin == Cout");
  VB_CHECK_ARG(Cin % SM_BLOCK_K == 0, "vb_snac_pwconv_tc: Cin must be a multiple of %d", SM_BLOCK_K);
  VB_CHECK_ARG(0 <= t_lo && t_lo <= t_hi && t_hi <= T, "vb_snac_pwconv_tc: range [%d, %d) outside [0, %d)", t_lo, t_hi, T);
  if (B <= 0 || t_lo == t_hi) return 0;
  SnacGemmParams p = {};
  p.w_tiles = static_cast<const uint8_t*>(d_w_tiles);
  p.x = d_x; p.y = d_y; p.bias = d_bias; p.resid = d_resid; p.noise = d_noise; p.alpha_out = d_alpha_out;
  p.kind = 0; p.epi = epilogue; p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.K = Cin;
  p.n_lo = t_lo; p.nr = t_hi - t_lo; p.n_total = B * (t_hi - t_lo); p.s = 1; p.pad = 0;
  return launch_snac_gemm(p, 1, static_cast<cudaStream_t>(stream));
}

int vb_snac_convtr_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias,
                      const float* d_alpha_out, int B, int Cin, int Cout, int T, int stride, int o_lo, int o_hi,
                      void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_tiles, "vb_snac_convtr_tc: null pointer");
  VB_CHECK_ARG(stride >= 1 && Cin % SM_BLOCK_K == 0, "vb_snac_convtr_tc: Cin must be a multiple of %d", SM_BLOCK_K);
  VB_CHECK_ARG(0 <= o_lo && o_lo <= o_hi && o_hi <= T * stride,
               "vb_snac_convtr_tc: output range [%d, %d) outside [0, %d)", o_lo, o_hi, T * stride);
  if (B <= 0 || o_lo == o_hi) return 0;
  const int pad = (stride + 1) / 2;
  // input positions n in [0, T] whose outputs n*stride + r - pad (r < stride) fall inside [o_lo, o_hi)
  int n_lo = (o_lo + pad - (stride - 1)) / stride;
  if (o_lo + pad - (stride - 1) < 0) n_lo = 0;
  int n_hi = (o_hi - 1 + pad) / stride + 1;
  if (n_hi > T + 1) n_hi = T + 1;
  SnacGemmParams p = {};
  p.w_tiles = static_cast<const uint8_t*>(d_w_tiles);
  p.x = d_x; p.y = d_y; p.bias = d_bias; p.alpha_out = d_alpha_out;
  p.kind = 1; p.epi = 0; p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.K = 2 * Cin;
  p.n_lo = n_lo; p.nr = n_hi - n_lo; p.n_total = B * (n_hi - n_lo); p.s = stride; p.pad = pad;
  return launch_snac_gemm(p, stride, static_cast<cudaStream_t>(stream));
}

// ---- streaming codec stages on the same kernel (codec.cu holds the fp32 SIMT forms with identical arguments) ----
int vb_codec_conv_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_resid,
                     const float* d_scale, const float* d_ctx, const float* d_act_a, const float* d_act_ib, int epilogue,
                     int act_in, int B, int Cin, int Cout, int T, int ksize, int dilation, void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_tiles, "vb_codec_conv_tc: null pointer");
  VB_CHECK_ARG(epilogue >= 0 && epilogue <= 6 && act_in >= 0 && act_in <= 2, "vb_codec_conv_tc: epilogue %d / activation %d",
               epilogue, act_in);
  VB_CHECK_ARG((epilogue != 1 && epilogue != 2 && epilogue != 5) || d_resid, "vb_codec_conv_tc: this epilogue needs d_resid");
  VB_CHECK_ARG(epilogue != 2 || d_scale, "vb_codec_conv_tc: LayerScale epilogue needs d_scale");
  VB_CHECK_ARG(act_in != 2 || (d_act_a && d_act_ib), "vb_codec_conv_tc: SnakeBeta needs its two per-channel tables");
  VB_CHECK_ARG(ksize >= 1 && dilation >= 1 && Cin % SM_BLOCK_K == 0, "vb_codec_conv_tc: Cin must be a multiple of %d", SM_BLOCK_K);
  if (B <= 0 || T <= 0) return 0;
  // codec.cu epilogue codes -> this kernel's: 0 plain, 1 resid, 2 scale_resid, 3 gelu, 4 silu, 5 mul, 6 clamp
  static const int epi_map[7] = {0, 1, 3, 4, 5, 6, 7};
  SnacGemmParams p = {};
  p.w_tiles = static_cast<const uint8_t*>(d_w_tiles);
  p.x = d_x; p.y = d_y; p.bias = d_bias; p.resid = d_resid; p.scale = d_scale; p.ctx = d_ctx; p.act_a = d_act_a; p.act_ib = d_act_ib;
  p.kind = ksize == 1 ? 0 : 2; p.epi = epi_map[epilogue]; p.act_in = act_in; p.ksize = ksize; p.dil = dilation;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.K = Cin * ksize;
  p.n_lo = 0; p.nr = T; p.n_total = B * T; p.s = 1; p.pad = 0;
  return launch_snac_gemm(p, 1, static_cast<cudaStream_t>(stream));
}

int vb_codec_convtr_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_ctx,
                       const float* d_act_a, const float* d_act_ib, int act_in, int B, int Cin, int Cout, int T, int stride,
                       void* stream) {
  VB_CHECK_ARG(d_y && d_x && d_w_tiles, "vb_codec_convtr_tc: null pointer");
  VB_CHECK_ARG(stride >= 1 && Cin % SM_BLOCK_K == 0 && act_in >= 0 && act_in <= 2, "vb_codec_convtr_tc: Cin must be a multiple of %d",
               SM_BLOCK_K);
  VB_CHECK_ARG(act_in != 2 || (d_act_a && d_act_ib), "vb_codec_convtr_tc: SnakeBeta needs its two per-channel tables");
  if (B <= 0 || T <= 0) return 0;
  SnacGemmParams p = {};
  p.w_tiles = static_cast<const uint8_t*>(d_w_tiles);
  p.x = d_x; p.y = d_y; p.bias = d_bias; p.ctx = d_ctx; p.act_a = d_act_a; p.act_ib = d_act_ib;
  p.kind = 1; p.epi = 0; p.act_in = act_in; p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.K = 2 * Cin;
  p.n_lo = 0; p.nr = T; p.n_total = B * T; p.s = stride; p.pad = 0;       // causal: output n s + r, inputs n and n - 1
  return launch_snac_gemm(p, stride, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
