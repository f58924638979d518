// Dev micro-benchmark (not part of the library): can the idle HBM time of latency-bound kernels be used to pull the
// NEXT kernel's weights into the 126 MB L2?
//   A. bulk-copy stream of an L2-resident buffer vs an HBM-resident one (per-SM ring, what the GEMM producer does)
//   B. kernel P prefetches X MB into L2 (cp.async.bulk.prefetch.L2 or per-thread prefetch.global.L2), kernel S then
//      streams the same bytes: time of S cold / after P, and of P itself
//   C. does a prefetched range survive a 100 MB evict_first stream over another buffer?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_prefetch l2_prefetch.cu && ./l2_prefetch
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
__device__ __forceinline__ void bulk_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// every CTA streams its contiguous share `reps` times through a ring of `slots` x `slot_bytes`
__global__ void __launch_bounds__(64) bulk_stream(const uint8_t* src, size_t per_cta, int slots, int slot_bytes, int reps,
                                                  unsigned long long* sink, int hint) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = (uint64_t*)(sm + (size_t)slots * slot_bytes);
  uint64_t* empty = full + slots;
  if (threadIdx.x == 0) { for (int s = 0; s < slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  const int n1 = (int)(per_cta / slot_bytes), n = n1 * reps;
  if (threadIdx.x == 0) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int i = 0; i < n; ++i) {
      const int s = i % slots; const uint32_t par = (i / slots) & 1;
      mbar_wait(&empty[s], par ^ 1);
      mbar_expect(&full[s], slot_bytes);
      if (hint) bulk_hint(sm + (size_t)s * slot_bytes, base + (size_t)(i % n1) * slot_bytes, slot_bytes, &full[s], pol);
      else bulk(sm + (size_t)s * slot_bytes, base + (size_t)(i % n1) * slot_bytes, slot_bytes, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (int i = 0; i < n; ++i) {
      const int s = i % slots; const uint32_t par = (i / slots) & 1;
      mbar_wait(&full[s], par);
      acc += *(volatile unsigned long long*)(sm + (size_t)s * slot_bytes);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

// mode 0: cp.async.bulk.prefetch.L2 in `piece`-byte pieces (one thread per piece); mode 1: prefetch.global.L2 per 128 B line;
// mode 2: prefetch.global.L2::evict_last per 128 B line
__global__ void __launch_bounds__(256) prefetch_kernel(const uint8_t* src, size_t bytes, int mode, int piece) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (size_t off = tid * piece; off < bytes; off += nth * piece)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((uint64_t)(src + off)), "r"((uint32_t)piece) : "memory");
  } else if (mode == 1) {
    for (size_t off = tid * 128; off < bytes; off += nth * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"((uint64_t)(src + off)) : "memory");
  } else {
    for (size_t off = tid * 128; off < bytes; off += nth * 128)
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"((uint64_t)(src + off)) : "memory");
  }
}

__global__ void spin_kernel(long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
}

__global__ void flush_kernel(uint4* p, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_uint4(1, 2, 3, 4);
}

static int sms;
static cudaEvent_t e0, e1, e2;
static uint8_t* flushbuf;
static unsigned long long* sink;

static void flush() { flush_kernel<<<sms * 4, 1024>>>((uint4*)flushbuf, ((size_t)512 << 20) / 16); }

static float stream_ms(const uint8_t* buf, size_t bytes, int reps, int hint, int slots = 10, int slot_kb = 16) {
  const int slot_bytes = slot_kb * 1024;
  const size_t per = bytes / sms / slot_bytes * slot_bytes;
  const int smem = slots * slot_bytes + 2 * slots * 8 + 64;
  cudaEventRecord(e0);
  bulk_stream<<<sms, 64, smem>>>(buf, per, slots, slot_bytes, reps, sink, hint);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  const size_t big = (size_t)1 << 30;
  uint8_t *x, *y;
  cudaMalloc(&x, big); cudaMalloc(&y, big); cudaMalloc(&flushbuf, (size_t)512 << 20); cudaMalloc(&sink, 8);
  cudaMemset(x, 1, big); cudaMemset(y, 2, big);
  cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("SMs %d\n", sms);

  // ---- A: L2-resident vs HBM-resident stream ----
  for (int mb : {8, 16, 32, 64, 96}) {
    const size_t bytes = (size_t)mb << 20;
    for (int hint = 0; hint < 2; ++hint) {
      flush();
      stream_ms(x, bytes, 1, 0);   // warm the range into L2 (no evict_first)
      float best = 1e9;
      for (int r = 0; r < 3; ++r) { float ms = stream_ms(x, bytes, 8, hint); if (ms < best) best = ms; }
      const int slot_bytes = 16 * 1024;
      const size_t per = bytes / sms / slot_bytes * slot_bytes;
      printf("A  %3d MB re-read x8 hint %d: %7.1f GB/s\n", mb, hint, (double)per * sms * 8 / best / 1e6);
    }
  }
  {
    float best = 1e9;
    for (int r = 0; r < 3; ++r) { float ms = stream_ms(x, big, 1, 1); if (ms < best) best = ms; }
    printf("A  1 GiB HBM stream (evict_first): %7.1f GB/s\n", (double)(big / sms / 16384 * 16384) * sms / best / 1e6);
  }

  // ---- B: prefetch then stream ----
  for (int mb : {20, 50, 100}) {
    const size_t bytes = (size_t)mb << 20;
    flush(); cudaDeviceSynchronize();
    float cold = stream_ms(x, bytes, 1, 1);
    printf("B  %3d MB cold stream: %7.2f us (%6.1f GB/s)\n", mb, cold * 1e3, bytes / cold / 1e6);
    struct M { int mode, piece, grid; const char* name; };
    M ms_[] = {{0, 8192, 148, "bulk.prefetch 8K x148 CTAs"}, {0, 32768, 148, "bulk.prefetch 32K x148"}, {0, 8192, 16, "bulk.prefetch 8K x16 CTAs"},
               {1, 0, 148, "prefetch.global.L2 x148"}, {1, 0, 592, "prefetch.global.L2 x592"}, {2, 0, 148, "prefetch.L2::evict_last x148"}};
    for (auto m : ms_) {
      for (int wait_us : {0, 30}) {
        flush(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        prefetch_kernel<<<m.grid, 256>>>(x, bytes, m.mode, m.piece);
        cudaEventRecord(e1);
        if (wait_us) spin_kernel<<<1, 1>>>(wait_us * 1000LL);
        cudaEventRecord(e2);
        cudaEventSynchronize(e2);
        float tp, tw; cudaEventElapsedTime(&tp, e0, e1); cudaEventElapsedTime(&tw, e1, e2);
        float warm = stream_ms(x, bytes, 1, 1);
        printf("B  %3d MB %-30s: prefetch kernel %7.2f us, gap %5.1f us, then stream %7.2f us (%6.1f GB/s)  err %s\n", mb, m.name, tp * 1e3,
               tw * 1e3, warm * 1e3, bytes / warm / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }

  // ---- C: survival of a prefetched range under an evict_first stream of another buffer ----
  for (int keep_mb : {20, 50}) {
    for (int other_mb : {50, 100, 200}) {
      for (int mode : {1, 2}) {
        const size_t kb = (size_t)keep_mb << 20, ob = (size_t)other_mb << 20;
        flush(); cudaDeviceSynchronize();
        prefetch_kernel<<<148, 256>>>(x, kb, mode, 0);
        spin_kernel<<<1, 1>>>(40 * 1000LL);
        float o = stream_ms(y, ob, 1, 1);
        float k = stream_ms(x, kb, 1, 1);
        printf("C  keep %3d MB (mode %d), stream other %3d MB evict_first (%6.1f GB/s), then keep-range stream %7.2f us (%6.1f GB/s)\n", keep_mb, mode,
               other_mb, ob / o / 1e6, k * 1e3, kb / k / 1e6);
      }
    }
  }

  // ---- D: back-to-back small streams (emulates a chain of projection launches): cold vs each prefetched by the previous ----
  {
    const int n = 8; const size_t each = (size_t)20 << 20;
    flush(); cudaDeviceSynchronize();
    const int slot_bytes = 16384, slots = 10, smem = slots * slot_bytes + 2 * slots * 8 + 64;
    const size_t per = each / sms / slot_bytes * slot_bytes;
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i) bulk_stream<<<sms, 64, smem>>>(x + i * each, per, slots, slot_bytes, 1, sink, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t0; cudaEventElapsedTime(&t0, e0, e1);
    flush(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n) prefetch_kernel<<<148, 256>>>(x + (i + 1) * each, each, 1, 0);
      bulk_stream<<<sms, 64, smem>>>(x + i * each, per, slots, slot_bytes, 1, sink, 1);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t1; cudaEventElapsedTime(&t1, e0, e1);
    printf("D  8 x 20 MB streams back to back: plain %7.2f us (%6.1f GB/s), with next-range prefetch kernels in between %7.2f us (%6.1f GB/s)\n",
           t0 * 1e3, n * each / t0 / 1e6, t1 * 1e3, n * each / t1 / 1e6);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
