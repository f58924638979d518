// Whole-page gather / scatter between the paged KV cache and a contiguous staging buffer: the device half of the
// prefill-KV hand-off between replicas ("[L, n_pages, 2, page, Hkv, D] slices over NVLink").
// The reference has no such path (one replica serves a request for its lifetime, vox_serve/launch.py:471-474);
// north_star names it as the one optional NCCL transfer.  Pure HBM copy: 16-byte vectors, four independent loads in
// flight per thread, a grid of a few CTAs per SM.
#include "../../include/vb_api.h"
#define VB_PDL_FAMILY 4
#include "common.cuh"

namespace vb {

// staging: [layer][i][page_vec]; cache: [layer][pages_per_layer][page_vec]  (page_vec = 16-byte vectors per page)
__global__ void __launch_bounds__(256) copy_pages_kernel(uint4* __restrict__ cache, uint4* __restrict__ staging,
                                                         const int32_t* __restrict__ page_ids, int n_pages,
                                                         int n_layers, long long pages_per_layer, long long page_vec,
                                                         int to_cache) {
  pdl_sync();
  const long long total = static_cast<long long>(n_layers) * n_pages * page_vec;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; idx < total; idx += 4 * stride) {
    uint4 v[4];
    long long src_off[4], dst_off[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = idx + u * stride;
      ok[u] = j < total;
      const long long jj = ok[u] ? j : 0;
      const long long lp = jj / page_vec, vi = jj - lp * page_vec;
      const int l = static_cast<int>(lp / n_pages), i = static_cast<int>(lp - static_cast<long long>(l) * n_pages);
      const long long c = (static_cast<long long>(l) * pages_per_layer + page_ids[i]) * page_vec + vi;
      src_off[u] = to_cache ? jj : c;
      dst_off[u] = to_cache ? c : jj;
    }
    const uint4* src = to_cache ? staging : cache;
    uint4* dst = to_cache ? cache : staging;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) v[u] = src[src_off[u]];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) dst[dst_off[u]] = v[u];
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb_copy_pages(void* d_cache, void* d_staging, const int32_t* d_page_ids, int n_pages, int n_layers,
                             int64_t pages_per_layer, int64_t page_bytes, int to_cache, void* stream) {
  VB_CHECK_ARG(d_cache && d_staging && d_page_ids, "vb_copy_pages: null pointer");
  VB_CHECK_ARG(page_bytes > 0 && page_bytes % 16 == 0, "vb_copy_pages: page_bytes %lld must be a multiple of 16",
               static_cast<long long>(page_bytes));
  VB_CHECK_ARG(n_layers > 0 && pages_per_layer > 0, "vb_copy_pages: bad cache shape");
  if (n_pages <= 0) return 0;
  const long long page_vec = page_bytes / 16;
  const long long total = static_cast<long long>(n_layers) * n_pages * page_vec;
  int sms = 148, dummy = 0;
  vb_device_info(&sms, &dummy);
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 8LL * sms) blocks = 8LL * sms;
  VB_LAUNCH_PDL(copy_pages_kernel, static_cast<unsigned>(blocks), 256, 0, stream, static_cast<uint4*>(d_cache),
                static_cast<uint4*>(d_staging), d_page_ids, n_pages, n_layers, static_cast<long long>(pages_per_layer),
                page_vec, to_cache);
  return 0;
}
