// Fused multi-codebook sampler: repetition penalty -> temperature -> softmax -> top-k / top-p / min-p
// filter -> draw, plus the repetition-cache update.  Replaces vox_serve/sampling.py (Sampler.run_sampling,
// apply_repetition_penalty, update_repetition_penalty_cache) and the FlashInfer sampling kernels it calls.
//
// Exact + deterministic by construction:
//   * logits are bf16, so after penalty and temperature every token's value is one of 65536 bf16 codes and
//     every token with the same code has the same (bf16-rounded) probability.  One pass turns a row into a
//     65536-bin integer histogram (monotone 16-bit key) with L2 reductions spread over the whole GPU; one
//     CTA per row then scans the histogram from the top: max, softmax denominator, filter thresholds and
//     the sampled (key, rank) pair.  All sums are integer counts times a per-key fixed-point (Q40) mass, so
//     the result is independent of thread scheduling; a final pass resolves (key, rank) to the token index.
//   * greedy is a packed 64-bit atomicMax over (key, ~index): largest value, smallest index on ties, exactly
//     torch.argmax on the penalised bf16 logits.
#include "../../include/vb_api.h"
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
  unsigned long long* rng_state;   // optional device {seed, offset}: offset advances by one per vb_sample call
  unsigned long long* packed;  // [rows] greedy result
  uint32_t* hist;              // [rows][65536]
  uint32_t* pick;              // [rows][2]  (key, rank)
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

__global__ void __launch_bounds__(256) argmax_kernel(const SampleParams p) {
  const int row = blockIdx.y;
  unsigned long long best = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.vocab; i += gridDim.x * blockDim.x) {
    const float v = token_value(p, row, i, false);
    if (v != v) continue;  // NaN never wins (torch.argmax would propagate; logits are finite here)
    const unsigned long long pk =
        (static_cast<unsigned long long>(bf16_key(v)) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
    best = pk > best ? pk : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(&p.packed[row], best);
}

__global__ void unpack_argmax_kernel(int64_t* out, const unsigned long long* packed, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) out[r] = static_cast<int64_t>(~static_cast<uint32_t>(packed[r] & 0xffffffffull));
}

__global__ void __launch_bounds__(256) hist_kernel(const SampleParams p) {
  const int row = blockIdx.y;
  uint32_t* h = p.hist + static_cast<size_t>(row) * 65536;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.vocab; i += gridDim.x * blockDim.x) {
    const float v = token_value(p, row, i, true);
    atomicAdd(&h[bf16_key(v)], 1u);
  }
}

// ---- Philox4x32-10 ----
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t offset, uint32_t stream) {
  uint32_t c[4] = {static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32), stream, 0u};
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

constexpr int SCAN_THREADS = 1024;
constexpr int KEYS_PER_THREAD = 65536 / SCAN_THREADS;   // 64
constexpr double Q40 = 1099511627776.0;                 // 2^40

// One CTA per row.  Thread t owns keys [hi - 63, hi], hi = 65535 - 64 t  (descending value order).
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const SampleParams p) {
  pdl_sync();
  __shared__ unsigned long long sh[33];
  __shared__ float sh_f[32];
  __shared__ int sh_i[4];
  __shared__ unsigned long long sh_u[4];
  const int row = blockIdx.x, tid = threadIdx.x;
  const uint32_t* h = p.hist + static_cast<size_t>(row) * 65536;
  const int khi = 65535 - KEYS_PER_THREAD * tid;

  // ---- max key (skip NaN codes: keys above +inf / below -inf never produced by finite logits) ----
  int kmax = -1;
  for (int j = 0; j < KEYS_PER_THREAD; ++j) {
    if (h[khi - j] != 0u) { kmax = khi - j; break; }
  }
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if ((tid & 31) == 0) sh_f[tid >> 5] = __int_as_float(kmax);
  __syncthreads();
  if (tid < 32) {
    int v = __float_as_int(sh_f[tid]);
    v = __reduce_max_sync(0xffffffffu, v);
    if (tid == 0) sh_i[0] = v;
  }
  __syncthreads();
  kmax = sh_i[0];
  const float xmax = key_value(kmax);

  // pass A: un-normalised weights.  strategy 3 first restricts to the top-k keys.
  auto weight = [&](int k) { return expf(key_value(k) - xmax); };   // fp32 softmax numerator
  unsigned long long total, excl;

  int k_floor = 0;      // keys below k_floor are filtered out before the softmax (top-k-first, strategy 3)
  if (p.strategy == 3 || p.strategy == 1) {
    unsigned long long c = 0;
    for (int j = 0; j < KEYS_PER_THREAD; ++j) c += h[khi - j];
    excl = block_excl_scan(c, sh, &total);
    // first key (descending) where cumulative count reaches top_k
    if (tid == 0) sh_i[1] = 0;
    __syncthreads();
    if (excl < static_cast<unsigned long long>(p.top_k) && excl + c >= static_cast<unsigned long long>(p.top_k)) {
      unsigned long long run = excl;
      for (int j = 0; j < KEYS_PER_THREAD; ++j) {
        run += h[khi - j];
        if (run >= static_cast<unsigned long long>(p.top_k)) { sh_i[1] = khi - j; break; }
      }
    }
    __syncthreads();
    k_floor = sh_i[1];     // 0 when the row has fewer than top_k tokens: keep everything
  }
  // softmax denominator over the keys the reference's softmax sees
  const int z_floor = (p.strategy == 3) ? k_floor : 0;
  float zpart = 0.f;
  for (int j = 0; j < KEYS_PER_THREAD; ++j) {
    const int k = khi - j;
    const uint32_t c = h[k];
    if (c != 0u && k >= z_floor) zpart += static_cast<float>(c) * weight(k);
  }
  // deterministic tree: warp shuffle then warp 0
  zpart = warp_sum(zpart);
  if ((tid & 31) == 0) sh_f[tid >> 5] = zpart;
  __syncthreads();
  if (tid < 32) {
    float v = sh_f[tid];
    v = warp_sum(v);
    if (tid == 0) sh_f[0] = v;
  }
  __syncthreads();
  const float Z = sh_f[0];
  __syncthreads();
  auto prob_q40 = [&](int k) -> unsigned long long {   // bf16-rounded softmax output as Q40 fixed point
    const float pr = round_bf16(weight(k) / Z);
    return static_cast<unsigned long long>(static_cast<double>(pr) * Q40);
  };

  // ---- filter: smallest kept key k_keep ----
  int k_keep = (p.strategy == 1 || p.strategy == 3) ? k_floor : 0;
  if (p.strategy == 2 || p.strategy == 3) {
    // keep key k while mass(keys > k) < top_p   (FlashInfer top-p: ties at the boundary all kept)
    unsigned long long mpart = 0;
    for (int j = 0; j < KEYS_PER_THREAD; ++j) {
      const int k = khi - j;
      const uint32_t c = h[k];
      if (c != 0u && k >= z_floor) mpart += static_cast<unsigned long long>(c) * prob_q40(k);
    }
    excl = block_excl_scan(mpart, sh, &total);
    const unsigned long long P = static_cast<unsigned long long>(static_cast<double>(p.top_p) * Q40);
    if (tid == 0) sh_i[2] = z_floor;
    __syncthreads();
    if (excl < P && excl + mpart >= P) {
      unsigned long long run = excl;
      for (int j = 0; j < KEYS_PER_THREAD; ++j) {
        const int k = khi - j;
        const uint32_t c = h[k];
        if (c != 0u && k >= z_floor) {
          run += static_cast<unsigned long long>(c) * prob_q40(k);
          if (run >= P) { sh_i[2] = k; break; }
        }
      }
    }
    __syncthreads();
    k_keep = max(k_keep, sh_i[2]);
  } else if (p.strategy == 4) {
    const float pmax = round_bf16(weight(kmax) / Z);
    int mine = 65536;
    for (int j = 0; j < KEYS_PER_THREAD; ++j) {
      const int k = khi - j;
      if (h[k] != 0u && round_bf16(weight(k) / Z) >= p.min_p * pmax) mine = k;   // keeps the smallest passing key
    }
    mine = __reduce_min_sync(0xffffffffu, mine);
    if ((tid & 31) == 0) sh_f[tid >> 5] = __int_as_float(mine);
    __syncthreads();
    if (tid < 32) {
      int v = __float_as_int(sh_f[tid]);
      v = __reduce_min_sync(0xffffffffu, v);
      if (tid == 0) sh_i[2] = v;
    }
    __syncthreads();
    k_keep = sh_i[2];
  }

  // ---- draw from the kept keys, proportional to probability ----
  unsigned long long wpart = 0;
  for (int j = 0; j < KEYS_PER_THREAD; ++j) {
    const int k = khi - j;
    const uint32_t c = h[k];
    if (c != 0u && k >= k_keep) wpart += static_cast<unsigned long long>(c) * prob_q40(k);
  }
  excl = block_excl_scan(wpart, sh, &total);
  if (tid == 0) {
    const uint64_t seed = p.rng_state ? p.rng_state[0] : p.seed;
    const uint64_t offset = p.rng_state ? p.rng_state[1] : p.offset;
    const double u = philox_uniform(seed, offset, static_cast<uint32_t>(row));
    unsigned long long target = static_cast<unsigned long long>(u * static_cast<double>(total));
    if (target >= total) target = total ? total - 1 : 0;
    sh_u[0] = target;
    p.pick[row * 2 + 0] = static_cast<uint32_t>(kmax);   // fallback: degenerate mass -> argmax key, rank 0
    p.pick[row * 2 + 1] = 0u;
  }
  __syncthreads();
  const unsigned long long target = sh_u[0];
  if (wpart != 0ull && excl <= target && target < excl + wpart) {
    unsigned long long run = excl;
    for (int j = 0; j < KEYS_PER_THREAD; ++j) {
      const int k = khi - j;
      const uint32_t c = h[k];
      if (c != 0u && k >= k_keep) {
        const unsigned long long q = prob_q40(k);
        const unsigned long long m = static_cast<unsigned long long>(c) * q;
        if (target < run + m) {
          unsigned long long r = q ? (target - run) / q : 0ull;
          if (r >= c) r = c - 1;
          p.pick[row * 2 + 0] = static_cast<uint32_t>(k);
          p.pick[row * 2 + 1] = static_cast<uint32_t>(r);
          break;
        }
        run += m;
      }
    }
  }
}

// resolve (key, rank) -> index of the rank-th token (index order) whose value code equals key
__global__ void __launch_bounds__(1024) resolve_kernel(const SampleParams p) {
  pdl_sync();
  __shared__ unsigned long long sh[33];
  const int row = blockIdx.x, tid = threadIdx.x;
  const uint32_t key = p.pick[row * 2], rank = p.pick[row * 2 + 1];
  const int per = (p.vocab + blockDim.x - 1) / blockDim.x;
  const int i0 = tid * per, i1 = min(p.vocab, i0 + per);
  unsigned long long c = 0;
  for (int i = i0; i < i1; ++i) c += (bf16_key(token_value(p, row, i, true)) == key) ? 1ull : 0ull;
  unsigned long long total;
  const unsigned long long excl = block_excl_scan(c, sh, &total);
  if (tid == 0 && total == 0ull) p.out[row] = 0;
  if (tid == 0 && row == 0 && p.rng_state) p.rng_state[1] += 1ull;   // all scan CTAs have finished (stream order)
  if (c != 0ull && excl <= rank && rank < excl + c) {
    unsigned long long run = excl;
    for (int i = i0; i < i1; ++i) {
      if (bf16_key(token_value(p, row, i, true)) == key) {
        if (run == rank) { p.out[row] = i; break; }
        ++run;
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
  (void)vocab;
  return static_cast<size_t>(rows) * (65536 * sizeof(uint32_t) + 64);
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
  uint8_t* w = static_cast<uint8_t*>(ws);
  p.hist = reinterpret_cast<uint32_t*>(w);
  p.packed = w ? reinterpret_cast<unsigned long long*>(w + static_cast<size_t>(rows) * 65536 * 4) : nullptr;
  p.pick = w ? reinterpret_cast<uint32_t*>(w + static_cast<size_t>(rows) * 65536 * 4 + static_cast<size_t>(rows) * 8)
             : nullptr;
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
  const int gx = max(1, min(32, (vocab + 256 * 8 - 1) / (256 * 8)));
  if (strategy == 0) {
    VB_CHECK_CUDA(cudaMemsetAsync(p.packed, 0, static_cast<size_t>(rows) * 8, st));
    VB_LAUNCH_PLAIN(argmax_kernel, dim3(gx, rows), 256, 0, st, p);
    VB_LAUNCH_PDL(unpack_argmax_kernel, (rows + 127) / 128, 128, 0, st, d_out_ids, p.packed, rows);
    return 0;
  }
  VB_CHECK_CUDA(cudaMemsetAsync(p.hist, 0, static_cast<size_t>(rows) * 65536 * 4, st));
  VB_LAUNCH_PLAIN(hist_kernel, dim3(gx, rows), 256, 0, st, p);
  VB_LAUNCH_PDL(scan_kernel, rows, SCAN_THREADS, 0, st, p);
  VB_LAUNCH_PDL(resolve_kernel, rows, 1024, 0, st, p);
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
