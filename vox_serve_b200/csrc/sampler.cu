// Fused multi-codebook sampler: repetition penalty -> temperature -> softmax -> top-k / top-p / min-p
// filter -> draw, plus the repetition-cache update.  Replaces vox_serve/sampling.py (Sampler.run_sampling,
// apply_repetition_penalty, update_repetition_penalty_cache) and the FlashInfer sampling kernels it calls.
//
// Exact + deterministic by construction:
//   * logits are bf16, so after penalty and temperature every token's value is one of 65536 bf16 codes (a
//     monotone 16-bit key) and every token with the same code has the same bf16-rounded probability;
//   * one thread-block cluster per row keeps the whole row on chip after a single pass (sample_kernel below):
//     integer masses, fixed reduction trees, no atomics -- the result does not depend on thread scheduling;
//   * greedy = largest key, smallest index on ties: exactly torch.argmax on the penalised bf16 logits.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 16
#include "common.cuh"

namespace vb {

struct SampleParams {
  const __nv_bfloat16* logits;
  const uint8_t* rep_cache;    // [B][W][C][V] or null
  const int32_t* cache_rows;   // optional: batch row b reads / marks cache row cache_rows[b] (slot-resident caches)
  int rows, vocab, ld;
  int W, C_cache, C_logits;
  float penalty, temperature;
  int strategy, top_k;
  float top_p, min_p;
  int mask_token;
  uint64_t seed, offset;
  unsigned long long* rng_state;   // optional device {seed, offset, arrivals}: offset advances by one per call
  int64_t* out;
};

__device__ __forceinline__ uint32_t bf16_key(float v) {  // monotone: larger value <-> larger key
  const uint32_t b = __float_as_uint(v) >> 16;
  return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  const uint32_t b = (k & 0x8000u) ? (k & 0x7fffu) : (~k & 0xffffu);
  return __uint_as_float(b << 16);
}

// penalised (and, for stochastic strategies, temperature-scaled) bf16 value of token i of `row`
__device__ __forceinline__ float token_value(const SampleParams& p, int row, int i, bool scaled) {
  float l = __bfloat162float(p.logits[static_cast<size_t>(row) * p.ld + i]);
  if (p.rep_cache) {
    const int b0 = row / p.C_logits;
    const int b = p.cache_rows ? p.cache_rows[b0] : b0;
    const int c = (p.C_logits == 1 && p.C_cache != 1) ? 0 : row % p.C_logits;   // sampling.py:140-141
    bool seen = false;
    for (int w = 0; w < p.W; ++w)
      seen |= p.rep_cache[((static_cast<size_t>(b) * p.W + w) * p.C_cache + c) * p.vocab + i] != 0;
    if (seen) l = (l > 0.f) ? round_bf16(l / p.penalty) : round_bf16(l * p.penalty);   // sampling.py:143-144
  }
  if (i == p.mask_token) l = -INFINITY;
  if (scaled) l = round_bf16(l / p.temperature);
  return l;
}

// ---- Philox4x32-10 ----
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t offset, uint32_t stream, uint32_t sub) {
  uint32_t c[4] = {static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32), stream, sub};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint64_t bits = (static_cast<uint64_t>(c[0]) << 21) ^ (static_cast<uint64_t>(c[1]) >> 11);  // 53 bits
  return static_cast<double>(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}

// Block-wide exclusive scan of one u64 per thread, thread order; returns exclusive prefix, *total = sum.
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* sh,
                                                              unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < nw ? sh[lane] : 0ull, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    sh[lane] = wi - w;
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  const unsigned long long res = sh[warp] + incl - v;
  *total = sh[32];
  __syncthreads();
  return res;
}

// ---------------------------------------------------------------------------------------------------------
// One thread-block CLUSTER per row.  The row's tokens are spread over S CTAs x 1024 threads x (2 * PAIRS) tokens;
// after one pass over the logits (+ repetition cache) every token lives on chip -- its 16-bit value key in a
// register, its fixed-point probability mass in shared memory -- and everything else is on-chip reductions:
// block tree -> one DSMEM store per peer CTA -> one cluster barrier.  No histogram, no atomics, no second
// pass over HBM/L2.
//   max key -> [top-k: 16-step bisection on counts] -> softmax denominator (fp32, fixed summation tree) ->
//   mass q_i = floor(bf16(p_i) * 2^32) -> [min-p: direct filter] -> draw by inverse CDF in (cta, thread, slot)
//   order; top-p by REJECTION: the drawn token j is in the nucleus iff mass{p > p_j} < top_p, otherwise
//   everything with p <= p_j is discarded and the draw repeats on what is left (acceptance >= top_p per round;
//   conditional on acceptance the draw is the renormalised nucleus -- the scheme of FlashInfer's sorting-free
//   sampler, with exact integer masses instead of float partial sums, hence scheduling-independent).
// ---------------------------------------------------------------------------------------------------------
constexpr int SMP_THREADS = 1024;
constexpr int SMP_PAIRS = 20;            // 40 tokens per thread: 40960 tokens per CTA, 327680 per 8-CTA cluster
constexpr int SMP_MAX_CLUSTER = 8;

__device__ __forceinline__ void st_dsmem_u64(unsigned long long* local_ptr, uint32_t cta_rank, unsigned long long v) {
  uint32_t laddr = smem_u32(local_ptr), raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(cta_rank));
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(raddr), "l"(v) : "memory");
}

struct SmpCtx {
  int S, rank, buf;
  unsigned long long (*xch)[SMP_MAX_CLUSTER];   // [2][8] exchange slots, written by the peers
  unsigned long long* sh;                       // [33]
  __device__ __forceinline__ void publish(unsigned long long v) {   // one warp: lane r stores into CTA r
    const int lane = threadIdx.x & 31;
    if (S == 1) {
      if (lane == 0) xch[buf][0] = v;
    } else if (lane < S) {
      st_dsmem_u64(&xch[buf][rank], lane, v);
    }
  }
  __device__ __forceinline__ void sync() {
    if (S == 1) __syncthreads(); else cluster_sync_all();
  }
};
enum { SMP_SUM = 0, SMP_MAX = 1, SMP_MIN = 2 };
template <int OP>
__device__ __forceinline__ unsigned long long smp_op(unsigned long long a, unsigned long long b) {
  return OP == SMP_SUM ? a + b : (OP == SMP_MAX ? (a > b ? a : b) : (a < b ? a : b));
}
// cluster-wide reduction of one u64 per thread; every thread of every CTA gets the result
template <int OP>
__device__ __forceinline__ unsigned long long smp_reduce(unsigned long long v, SmpCtx& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = smp_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) c.sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long x = c.sh[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = smp_op<OP>(x, __shfl_xor_sync(0xffffffffu, x, o));
    c.publish(x);
  }
  c.sync();
  unsigned long long r = c.xch[c.buf][0];
  for (int i = 1; i < c.S; ++i) r = smp_op<OP>(r, c.xch[c.buf][i]);
  c.buf ^= 1;
  return r;
}
// fp32 sum with a fixed tree (xor butterfly inside the warp and across warps, CTA partials in rank order)
__device__ __forceinline__ float smp_sum_f32(float v, SmpCtx& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  float* shf = reinterpret_cast<float*>(c.sh);
  if (lane == 0) shf[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const float x = warp_sum(shf[lane]);
    c.publish(static_cast<unsigned long long>(__float_as_uint(x)));
  }
  c.sync();
  float r = 0.f;
  for (int i = 0; i < c.S; ++i) r += __uint_as_float(static_cast<uint32_t>(c.xch[c.buf][i]));
  c.buf ^= 1;
  return r;
}

// The kernel is instruction-bound (the whole row is on chip): the softmax uses the hardware exponential
// (ex2.approx, 2 ulp) -- a sampler's masses do not need expf's last bit.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float smp_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(SMP_THREADS, 1) sample_kernel(const SampleParams p, const int npairs) {
  extern __shared__ uint32_t q_sm[];                 // [2 * npairs][1024] masses (Q32)
  __shared__ unsigned long long xch[2][SMP_MAX_CLUSTER];
  __shared__ unsigned long long pick[2][2];
  __shared__ unsigned long long sh[34];
  SmpCtx c;
  c.S = gridDim.x; c.rank = blockIdx.x; c.buf = 0; c.xch = xch; c.sh = sh;
  const int S = c.S, rank = c.rank, row = blockIdx.y, tid = threadIdx.x;
  const bool scaled = p.strategy != 0;
  const int tr = (tid == 0 && trace_block0()) ? trace_begin(4) : -1;
  pdl_sync();

  // ---- one pass over the row: keys (0 = no token / NaN) ----
  // The pass is load-latency-bound (40 tokens per thread, 64 registers per thread), so the loads are batched: ten
  // token PAIRS at a time -- one 32-bit load for the two logits, one 16-bit load per window slot for their
  // repetition flags, all twenty-odd requests in flight together -- and only then finished (penalty, mask,
  // temperature: same operations and rounding points as token_value()).  Rows whose layout does not allow the paired
  // loads (odd leading dimension / vocabulary, unaligned base) take the scalar path.
  uint32_t kk[SMP_PAIRS];
  const __nv_bfloat16* lrow = p.logits + static_cast<size_t>(row) * p.ld;
  const uint8_t* rep0 = nullptr;
  size_t rep_wstride = 0;
  if (p.rep_cache) {
    const int b0 = row / p.C_logits;
    const int b = p.cache_rows ? p.cache_rows[b0] : b0;
    const int cc = (p.C_logits == 1 && p.C_cache != 1) ? 0 : row % p.C_logits;   // sampling.py:140-141
    rep_wstride = static_cast<size_t>(p.C_cache) * p.vocab;
    rep0 = p.rep_cache + (static_cast<size_t>(b) * p.W * p.C_cache + cc) * p.vocab;
  }
  const bool paired = ((p.ld | p.vocab) & 1) == 0 && (reinterpret_cast<uintptr_t>(lrow) & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(rep0) | rep_wstride) & 1) == 0;
  auto finish = [&](float l, bool seen, int i) -> uint32_t {
    if (seen) l = (l > 0.f) ? round_bf16(l / p.penalty) : round_bf16(l * p.penalty);   // sampling.py:143-144
    if (i == p.mask_token) l = -INFINITY;
    if (scaled) l = round_bf16(l / p.temperature);
    return (l == l) ? bf16_key(l) : 0u;
  };
  constexpr int SMP_BATCH = 5;
#pragma unroll
  for (int j0 = 0; j0 < SMP_PAIRS; j0 += SMP_BATCH) {
    uint32_t lw[SMP_BATCH], sb[SMP_BATCH];
#pragma unroll
    for (int jj = 0; jj < SMP_BATCH; ++jj) {
      const int j = j0 + jj;
      const int i0 = 2 * ((j * S + rank) * SMP_THREADS + tid);
      lw[jj] = 0u;
      sb[jj] = 0u;
      if (j < npairs && i0 < p.vocab) {
        if (paired) {
          lw[jj] = __ldg(reinterpret_cast<const uint32_t*>(lrow + i0));
          for (int w = 0; w < p.W; ++w)
            if (rep0) sb[jj] |= __ldg(reinterpret_cast<const unsigned short*>(rep0 + w * rep_wstride + i0));
        } else {
          lw[jj] = __ldg(reinterpret_cast<const unsigned short*>(lrow + i0));
          if (i0 + 1 < p.vocab)
            lw[jj] |= static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(lrow + i0 + 1))) << 16;
          for (int w = 0; w < p.W; ++w) {
            if (rep0) {
              sb[jj] |= __ldg(rep0 + w * rep_wstride + i0);
              if (i0 + 1 < p.vocab) sb[jj] |= static_cast<uint32_t>(__ldg(rep0 + w * rep_wstride + i0 + 1)) << 8;
            }
          }
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < SMP_BATCH; ++jj) {
      const int j = j0 + jj;
      const int i0 = 2 * ((j * S + rank) * SMP_THREADS + tid);
      uint32_t lo = 0u, hi = 0u;
      if (j < npairs && i0 < p.vocab) {
        lo = finish(bf16_lo(lw[jj]), (sb[jj] & 0xffu) != 0u, i0);
        if (i0 + 1 < p.vocab) hi = finish(bf16_hi(lw[jj]), (sb[jj] & 0xff00u) != 0u, i0 + 1);
      }
      kk[j] = lo | (hi << 16);
    }
  }
  auto tok_index = [&](int j, int half) { return 2 * ((j * S + rank) * SMP_THREADS + tid) + half; };

  // ---- max key ----
  uint32_t kloc = 0u;
#pragma unroll
  for (int j = 0; j < SMP_PAIRS; ++j) kloc = max(kloc, max(kk[j] & 0xffffu, kk[j] >> 16));
  const uint32_t kmax = static_cast<uint32_t>(smp_reduce<SMP_MAX>(kloc, c));
  // smallest index holding the max key: the greedy answer and the fallback of degenerate rows
  auto first_max_index = [&]() -> unsigned long long {
    unsigned long long best = 0xffffffffull;
#pragma unroll
    for (int j = SMP_PAIRS - 1; j >= 0; --j) {
      if ((kk[j] >> 16) == kmax) best = static_cast<unsigned long long>(tok_index(j, 1));
      if ((kk[j] & 0xffffu) == kmax) best = static_cast<unsigned long long>(tok_index(j, 0));
    }
    const unsigned long long r = smp_reduce<SMP_MIN>(kmax ? best : 0xffffffffull, c);
    return r == 0xffffffffull ? 0ull : r;
  };
  if (p.strategy == 0) {
    const unsigned long long idx = first_max_index();
    if (rank == 0 && tid == 0) p.out[row] = static_cast<int64_t>(idx);
    return;
  }

  // ---- top-k threshold: smallest key k with #{key_i > k} < top_k (ties at the k-th value are kept) ----
  uint32_t k_floor = 0u;
  if (p.strategy == 1 || p.strategy == 3) {
    uint32_t lo = 0u, hi = 65535u;          // invariant: count(> hi) < top_k; count(> lo - 1) >= top_k or lo == 0
    // bisection over [0, 65535] for the smallest k with count(> k) < top_k
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      uint32_t cnt = 0u;
#pragma unroll
      for (int j = 0; j < SMP_PAIRS; ++j) cnt += ((kk[j] & 0xffffu) > mid ? 1u : 0u) + ((kk[j] >> 16) > mid ? 1u : 0u);
      const unsigned long long tot = smp_reduce<SMP_SUM>(cnt, c);
      if (tot < static_cast<unsigned long long>(p.top_k)) hi = mid; else lo = mid + 1;
    }
    k_floor = lo;
  }
  const uint32_t z_floor = (p.strategy == 3) ? k_floor : 0u;   // top-k-first: the softmax only sees the top-k

  // ---- softmax denominator (fp32) ----
  const float xmax = key_value(kmax);
  float zpart = 0.f;
#pragma unroll
  for (int j = 0; j < SMP_PAIRS; ++j) {
    const uint32_t a = kk[j] & 0xffffu, b = kk[j] >> 16;
    if (a != 0u && a >= z_floor) zpart += smp_ex2((key_value(a) - xmax) * kLog2e);
    if (b != 0u && b >= z_floor) zpart += smp_ex2((key_value(b) - xmax) * kLog2e);
  }
  const float Z = smp_sum_f32(zpart, c);
  const float log2Z = log2f(Z);

  // ---- masses: bf16-rounded softmax output as Q32 fixed point; filtered tokens get 0 ----
  const uint32_t keep_floor = (p.strategy == 1 || p.strategy == 3) ? k_floor : 0u;
  const float minp_thr = (p.strategy == 4) ? p.min_p * round_bf16(smp_ex2(-log2Z)) : 0.f;
  unsigned long long msum = 0ull;
#pragma unroll
  for (int j = 0; j < SMP_PAIRS; ++j) {
    if (j < npairs) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t k = h ? (kk[j] >> 16) : (kk[j] & 0xffffu);
        uint32_t q = 0u;
        if (k != 0u && k >= z_floor && k >= keep_floor) {
          // softmax output e^(x - max) / Z as ONE exponential: 2^((x - max) log2 e - log2 Z)
          const float pr = round_bf16(smp_ex2((key_value(k) - xmax) * kLog2e - log2Z));
          if (pr >= minp_thr) q = __float2uint_rz(pr * 4294967296.f);   // saturates at 2^32 - 1
        }
        q_sm[(2 * j + h) * SMP_THREADS + tid] = q;
        msum += q;
      }
    }
  }
  unsigned long long Mc = smp_reduce<SMP_SUM>(msum, c);   // mass of the current candidate set
  const uint64_t seed = p.rng_state ? p.rng_state[0] : p.seed;
  const uint64_t offset = p.rng_state ? p.rng_state[1] : p.offset;
  if (Mc == 0ull || !(Z > 0.f)) {
    // degenerate row (everything masked / non-finite): argmax, like the reference's fallthrough
    const unsigned long long idx = first_max_index();
    if (rank == 0 && tid == 0) p.out[row] = static_cast<int64_t>(idx);
  } else {
    const bool nucleus = (p.strategy == 2 || p.strategy == 3);
    unsigned long long P = static_cast<unsigned long long>(static_cast<double>(p.top_p) * 4294967296.0);
    if (P == 0ull) P = 1ull;
    uint32_t pivot = 0u;                     // candidates: key > pivot
    int pb = 0;
    unsigned long long result = 0ull;
    for (uint32_t round = 0; round < 64u; ++round) {
      // inverse CDF over the candidates in (cta, thread, slot) order
      unsigned long long mine = 0ull;
#pragma unroll
      for (int j = 0; j < SMP_PAIRS; ++j) {
        if (j < npairs) {
          if ((kk[j] & 0xffffu) > pivot) mine += q_sm[(2 * j) * SMP_THREADS + tid];
          if ((kk[j] >> 16) > pivot) mine += q_sm[(2 * j + 1) * SMP_THREADS + tid];
        }
      }
      unsigned long long cta_total;
      const unsigned long long excl = block_excl_scan(mine, sh, &cta_total);
      if (tid < 32) c.publish(cta_total);
      c.sync();
      unsigned long long base = 0ull;
      for (int r = 0; r < rank; ++r) base += c.xch[c.buf][r];
      c.buf ^= 1;
      const double u = philox_uniform(seed, offset, static_cast<uint32_t>(row), round);
      unsigned long long target = static_cast<unsigned long long>(u * static_cast<double>(Mc));
      if (target >= Mc) target = Mc - 1;
      const unsigned long long lo = base + excl;
      if (mine != 0ull && lo <= target && target < lo + mine) {
        unsigned long long run = lo;
        unsigned long long found = 0ull;
        bool done = false;
#pragma unroll
        for (int j = 0; j < SMP_PAIRS; ++j) {
          if (j < npairs) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t k = h ? (kk[j] >> 16) : (kk[j] & 0xffffu);
              if (!done && k > pivot) {
                const unsigned long long q = q_sm[(2 * j + h) * SMP_THREADS + tid];
                if (target < run + q) {
                  found = (static_cast<unsigned long long>(k) << 32) | static_cast<uint32_t>(tok_index(j, h));
                  done = true;
                }
                run += q;
              }
            }
          }
        }
        if (S == 1) pick[pb][0] = found;
        else for (int r = 0; r < S; ++r) st_dsmem_u64(&pick[pb][0], r, found);
      }
      c.sync();
      result = pick[pb][0];
      pb ^= 1;
      if (!nucleus) break;
      const uint32_t kj = static_cast<uint32_t>(result >> 32);
      unsigned long long above = 0ull;
#pragma unroll
      for (int j = 0; j < SMP_PAIRS; ++j) {
        if (j < npairs) {
          if ((kk[j] & 0xffffu) > kj) above += q_sm[(2 * j) * SMP_THREADS + tid];
          if ((kk[j] >> 16) > kj) above += q_sm[(2 * j + 1) * SMP_THREADS + tid];
        }
      }
      above = smp_reduce<SMP_SUM>(above, c);
      if (above < P) break;                  // j is inside the nucleus: accept
      pivot = kj;                            // nothing with p <= p_j is: drop it and redraw
      Mc = above;
    }
    if (rank == 0 && tid == 0) p.out[row] = static_cast<int64_t>(result & 0xffffffffull);
  }
  // advance the device RNG offset once per call: the last row to get here (every row read the offset above,
  // before its final cluster barrier)
  if (p.rng_state && rank == 0 && tid == 0) {
    const unsigned long long old = atomicAdd(&p.rng_state[2], 1ull);
    if (old == static_cast<unsigned long long>(gridDim.y) - 1ull) {
      p.rng_state[2] = 0ull;
      p.rng_state[1] = offset + 1ull;
    }
  }
  trace_end(tr);
}

// Standard normal draws (SNAC NoiseBlock input, vox_serve/tokenizer/snac.py:208 uses torch.randn): one Philox4x32-10
// block per 4 outputs, counter = (offset, quad index), two Box-Muller pairs.  With a device state {seed, offset,
// arrivals} the offset advances by one per launch (last block to arrive), so CUDA-graph replays keep drawing.
__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ out, long long n, unsigned long long seed_,
                                                    unsigned long long offset_, unsigned long long* state) {
  pdl_sync();
  const unsigned long long seed = state ? state[0] : seed_, offset = state ? state[1] : offset_;
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q * 4 < n) {
    uint32_t c[4] = {static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32), static_cast<uint32_t>(q),
                     static_cast<uint32_t>(static_cast<unsigned long long>(q) >> 32) ^ 0x5eed0000u};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // u1 in (0, 1], u2 in [0, 1): 2^-32 granularity (torch's curand path uses the same resolution)
      const float u1 = (static_cast<float>(c[2 * h]) + 1.0f) * 2.3283064365386963e-10f;
      const float u2 = static_cast<float>(c[2 * h + 1]) * 2.3283064365386963e-10f;
      const float rad = sqrtf(-2.0f * logf(u1));
      float s, co;
      sincospif(2.0f * u2, &s, &co);
      z[2 * h] = rad * co;
      z[2 * h + 1] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = z[j];
  }
  if (state) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long old = atomicAdd(&state[2], 1ull);
      if (old == static_cast<unsigned long long>(gridDim.x) - 1ull) {
        state[2] = 0ull;
        state[1] = offset + 1ull;
      }
    }
  }
}

__global__ void penalty_kernel(__nv_bfloat16* out, const SampleParams p) {
  const int row = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.vocab; i += gridDim.x * blockDim.x)
    out[static_cast<size_t>(row) * p.vocab + i] = __float2bfloat16_rn(token_value(p, row, i, false));
}

// window > 1: cache[:, :-1] = cache[:, 1:]; cache[:, -1] = 0   (sampling.py:166-168)
__global__ void rep_shift_kernel(uint8_t* cache, const int32_t* cache_rows, int B, int W, size_t plane /*C*V*/) {
  pdl_sync();
  const size_t n = static_cast<size_t>(B) * plane;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b0 = i / plane, r = i - b0 * plane;
    const size_t b = cache_rows ? static_cast<size_t>(cache_rows[b0]) : b0;
    uint8_t* base = cache + b * W * plane + r;
    for (int w = 0; w + 1 < W; ++w) base[w * plane] = base[(w + 1) * plane];
    base[(W - 1) * plane] = 0;
  }
}
// cache[b, w(s), c(s), ids[b', c']] = 1 for every b and every (b', c')   (sampling.py:169-178)
__global__ void rep_mark_kernel(uint8_t* cache, const int32_t* cache_rows, const int64_t* ids, int B, int W, int C,
                                int V, int C_ids, int windowed) {
  pdl_sync();
  const int n_ids = B * C_ids;
  const bool cb0_only = (C_ids == 1 && C != 1);
  const int w_lo = windowed ? W - 1 : 0, w_hi = W;
  const int c_lo = 0, c_hi = cb0_only ? 1 : C;
  const int per_b = (w_hi - w_lo) * (c_hi - c_lo) * n_ids;
  const long long total = static_cast<long long>(B) * per_b;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b0 = static_cast<int>(t / per_b);
    int r = static_cast<int>(t - static_cast<long long>(b0) * per_b);
    const int b = cache_rows ? cache_rows[b0] : b0;
    const int j = r % n_ids; r /= n_ids;
    const int c = c_lo + r % (c_hi - c_lo); r /= (c_hi - c_lo);
    const int w = w_lo + r;
    long long id = ids[j];
    if (id < 0) id += V;
    if (id >= 0 && id < V) cache[((static_cast<size_t>(b) * W + w) * C + c) * V + id] = 1;
  }
}

}  // namespace vb

using namespace vb;

extern "C" {

size_t vb_sample_workspace_bytes(int rows, int vocab) {
  (void)vocab; (void)rows;
  return 256;   // the sampler keeps the row on chip; the workspace is kept in the ABI for callers that size it
}

static int fill_params(SampleParams& p, int64_t* d_out_ids, const void* d_logits, int rows, int vocab, int ld,
                       const uint8_t* d_rep_cache, int W, int C_cache, int C_logits, float penalty, int strategy,
                       int top_k, float top_p, float min_p, float temperature, uint64_t seed, uint64_t offset,
                       int mask_token, void* ws, uint64_t* rng_state = nullptr, const int32_t* cache_rows = nullptr) {
  p.logits = static_cast<const __nv_bfloat16*>(d_logits);
  p.rep_cache = d_rep_cache;
  p.cache_rows = cache_rows;
  p.rows = rows; p.vocab = vocab; p.ld = ld;
  p.W = W; p.C_cache = C_cache; p.C_logits = C_logits > 0 ? C_logits : 1;
  p.penalty = penalty; p.temperature = temperature;
  p.strategy = strategy; p.top_k = top_k; p.top_p = top_p; p.min_p = min_p;
  p.mask_token = mask_token; p.seed = seed; p.offset = offset;
  p.rng_state = reinterpret_cast<unsigned long long*>(rng_state);
  (void)ws;
  p.out = d_out_ids;
  return 0;
}

int vb_sample(int64_t* d_out_ids, const void* d_logits, int rows, int vocab, int ld_logits,
              const uint8_t* d_rep_cache, const int32_t* d_cache_rows, int rep_window_slots, int rep_codebooks,
              int logit_codebooks,
              float penalty, int strategy, int top_k, float top_p, float min_p, float temperature, uint64_t seed,
              uint64_t offset, uint64_t* d_rng_state, int mask_token, void* d_workspace, size_t workspace_bytes,
              void* stream) {
  VB_CHECK_ARG(d_out_ids && d_logits && d_workspace, "vb_sample: null pointer");
  VB_CHECK_ARG(strategy >= 0 && strategy <= 4, "vb_sample: strategy %d", strategy);
  VB_CHECK_ARG(workspace_bytes >= vb_sample_workspace_bytes(rows, vocab), "vb_sample: workspace too small");
  VB_CHECK_ARG(strategy == 0 || temperature > 0.f, "vb_sample: temperature must be > 0 for stochastic sampling");
  VB_CHECK_ARG(!(strategy == 1 || strategy == 3) || top_k > 0, "vb_sample: top_k must be > 0");
  VB_CHECK_ARG(!d_rep_cache || (rep_window_slots > 0 && rep_codebooks > 0), "vb_sample: bad repetition cache dims");
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SampleParams p;
  fill_params(p, d_out_ids, d_logits, rows, vocab, ld_logits, d_rep_cache, rep_window_slots, rep_codebooks,
              logit_codebooks, penalty, strategy, top_k, top_p, min_p, temperature, seed, offset, mask_token,
              d_workspace, d_rng_state, d_cache_rows);
  // smallest cluster (1, 2, 4 or 8 CTAs) whose threads can hold the row on chip
  int S = 1;
  while (S < SMP_MAX_CLUSTER && static_cast<long long>(S) * SMP_THREADS * 2 * SMP_PAIRS < vocab) S <<= 1;
  VB_CHECK_ARG(static_cast<long long>(S) * SMP_THREADS * 2 * SMP_PAIRS >= vocab,
               "vb_sample: vocab %d exceeds the %d tokens one cluster holds on chip", vocab,
               SMP_MAX_CLUSTER * SMP_THREADS * 2 * SMP_PAIRS);
  VB_CHECK_ARG(rows <= 65535, "vb_sample: %d rows exceed the grid limit (65535)", rows);
  const int npairs = (vocab + S * SMP_THREADS * 2 - 1) / (S * SMP_THREADS * 2);
  const size_t smem = static_cast<size_t>(2 * npairs) * SMP_THREADS * sizeof(uint32_t);
  // (static shared memory counts against the same 227 KiB: leave room for it)
  VB_CHECK_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     2 * SMP_PAIRS * SMP_THREADS * static_cast<int>(sizeof(uint32_t))));
  VB_LAUNCH_PDL_CLUSTER(sample_kernel, dim3(S, rows), SMP_THREADS, smem, st, S, p, npairs);
  return 0;
}

int vb_randn(float* d_out, int64_t n, uint64_t seed, uint64_t offset, uint64_t* d_rng_state, void* stream) {
  VB_CHECK_ARG(d_out, "vb_randn: null pointer");
  VB_CHECK_ARG(n >= 0 && n < (static_cast<int64_t>(1) << 40), "vb_randn: bad element count");
  if (n == 0) return 0;
  const long long quads = (n + 3) / 4;
  const unsigned blocks = static_cast<unsigned>((quads + 255) / 256);
  VB_LAUNCH_PDL(randn_kernel, blocks, 256, 0, stream, d_out, static_cast<long long>(n), seed, offset,
                reinterpret_cast<unsigned long long*>(d_rng_state));
  return 0;
}

int vb_apply_repetition_penalty(void* d_out, const void* d_logits, const uint8_t* d_rep_cache,
                                int rep_window_slots, int rep_codebooks, int logit_codebooks, float penalty,
                                int rows, int vocab, void* stream) {
  VB_CHECK_ARG(d_out && d_logits && d_rep_cache, "vb_apply_repetition_penalty: null pointer");
  if (rows <= 0) return 0;
  SampleParams p;
  fill_params(p, nullptr, d_logits, rows, vocab, vocab, d_rep_cache, rep_window_slots, rep_codebooks,
              logit_codebooks, penalty, 0, 0, 0.f, 0.f, 1.f, 0, 0, -1, nullptr);
  const int gx = max(1, min(64, (vocab + 255) / 256));
  VB_LAUNCH_PLAIN(penalty_kernel, dim3(gx, rows), 256, 0, stream, static_cast<__nv_bfloat16*>(d_out), p);
  return 0;
}

int vb_update_repetition_cache(uint8_t* d_cache, const int32_t* d_cache_rows, const int64_t* d_ids, int B, int W,
                               int C, int V, int C_ids, int window, void* stream) {
  VB_CHECK_ARG(d_cache && d_ids, "vb_update_repetition_cache: null pointer");
  VB_CHECK_ARG(B > 0 && W > 0 && C > 0 && V > 0 && C_ids > 0, "vb_update_repetition_cache: bad dims");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int windowed = window > 1;
  if (windowed) {
    const size_t plane = static_cast<size_t>(C) * V;
    size_t nb = (static_cast<size_t>(B) * plane + 255) / 256; if (nb > 4096) nb = 4096; const unsigned blocks = static_cast<unsigned>(nb);
    VB_LAUNCH_PDL(rep_shift_kernel, blocks, 256, 0, st, d_cache, d_cache_rows, B, W, plane);
  }
  VB_LAUNCH_PDL(rep_mark_kernel, 64, 256, 0, st, d_cache, d_cache_rows, d_ids, B, W, C, V, C_ids, windowed);
  return 0;
}

}  // extern "C"

VB_DEFINE_TRACE_SETTER(sampler)
