// Dev micro-benchmark (not part of the library): how fast can the SMs READ a large HBM-resident buffer with
//   (a) cp.async.bulk linear copies into a shared-memory ring (what the GEMM / attention producers do),
//   (b) plain LDG.128 from many warps (what a copy kernel does)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bw stream_bw.cu && ./stream_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// each CTA streams its contiguous share; `slots` x `slot_bytes` ring; `split` bulk copies per slot issued by `split` lanes
__device__ __forceinline__ void bulk_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__global__ void __launch_bounds__(64) bulk_stream(const uint8_t* src, size_t per_cta, int slots, int slot_bytes, int split, unsigned long long* sink, int hint, int delay) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = (uint64_t*)(sm + (size_t)slots * slot_bytes);
  uint64_t* empty = full + slots;
  if (threadIdx.x == 0) { for (int s = 0; s < slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / slot_bytes);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int i = 0; i < n; ++i) {
      const int s = i % slots; const uint32_t par = (i / slots) & 1;
      mbar_wait(&empty[s], par ^ 1);
      if (lane == 0) mbar_expect(&full[s], slot_bytes);
      __syncwarp();
      const int piece = slot_bytes / split;
      if (lane < split) {
        if (hint) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); bulk_hint(sm + (size_t)s * slot_bytes + lane * piece, base + (size_t)i * slot_bytes + lane * piece, piece, &full[s], pol); }
        else bulk(sm + (size_t)s * slot_bytes + lane * piece, base + (size_t)i * slot_bytes + lane * piece, piece, &full[s]);
      }
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (int i = 0; i < n; ++i) {
      const int s = i % slots; const uint32_t par = (i / slots) & 1;
      mbar_wait(&full[s], par);
      acc += *(volatile unsigned long long*)(sm + (size_t)s * slot_bytes);
      if (delay) __nanosleep(delay);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

__global__ void __launch_bounds__(1024) ldg_stream(const uint4* src, size_t n_vec, unsigned long long* sink) {
  unsigned long long acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  if (acc == 0x1234567) *sink = acc;
}

int main() {
  const size_t bytes = (size_t)2 << 30;   // 2 GiB >> L2
  uint8_t* d; unsigned long long* sink;
  cudaMalloc(&d, bytes); cudaMalloc(&sink, 8); cudaMemset(d, 1, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("SMs %d\n", sms);
  struct Cfg { int ctas_per_sm, slots, slot_kb, split, hint, delay; };
  Cfg cfgs[] = {{1, 5, 16, 1, 0, 0}, {1, 10, 16, 1, 0, 0}, {1, 10, 14, 1, 0, 0}, {1, 10, 14, 1, 1, 0}, {1, 10, 16, 1, 1, 0}, {1, 10, 14, 1, 1, 200}, {1, 10, 14, 1, 1, 400},
                {1, 10, 14, 1, 0, 400}, {1, 3, 64, 32, 0, 0}, {1, 3, 64, 32, 1, 0}, {1, 6, 32, 16, 0, 0}, {1, 6, 32, 2, 0, 0}};
  for (auto c : cfgs) {
    const int grid = sms * c.ctas_per_sm;
    const int slot_bytes = c.slot_kb * 1024;
    size_t per = bytes / grid / slot_bytes * slot_bytes;
    const int smem = c.slots * slot_bytes + 2 * c.slots * 8 + 64;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      bulk_stream<<<grid, 64, smem>>>(d, per, c.slots, slot_bytes, c.split, sink, c.hint, c.delay);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("bulk  ctas/sm %d slots %2d x %2d KB split %2d hint %d consumer delay %3d ns (in flight %4d KB/SM): %7.1f GB/s  %s\n", c.ctas_per_sm, c.slots, c.slot_kb, c.split, c.hint, c.delay,
           c.ctas_per_sm * c.slots * c.slot_kb, (double)per * grid / best / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  for (int bps = 1; bps <= 2; ++bps) {
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      ldg_stream<<<sms * bps, 1024>>>((const uint4*)d, bytes / 16, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("ldg   %d x 1024 threads/SM, 4 x LDG.128 in flight per thread: %7.1f GB/s\n", bps, (double)bytes / best / 1e6);
  }
  return 0;
}
