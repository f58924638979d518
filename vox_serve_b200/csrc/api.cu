// Error plumbing, device queries and host-side TMA descriptor encoding for libvoxb200.so.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cudaTypedefs.h>

#include "../../include/vb_api.h"
#include "common.cuh"

namespace vb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1, g_pdl_off = 0;
bool pdl_enabled(int family) {
  if (g_pdl < 0) {
    const char* e = getenv("VB_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
    if (const char* m = getenv("VB_PDL_OFF")) g_pdl_off = atoi(m);
  }
  return g_pdl != 0 && (family & g_pdl_off) == 0;
}

// cuTensorMapEncodeTiled comes from the driver; resolve it at run time so the library links without
// libcuda (the authoring container has no driver).
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}
cudaError_t trace_set_elementwise(unsigned long long*);
cudaError_t trace_set_attn(unsigned long long*);
cudaError_t trace_set_gemm(unsigned long long*);
cudaError_t trace_set_sampler(unsigned long long*);
cudaError_t trace_set_snac(unsigned long long*);
}  // namespace vb

extern "C" {

int vb_set_trace(void* d_buffer) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buffer);
  VB_CHECK_CUDA(vb::trace_set_elementwise(p));
  VB_CHECK_CUDA(vb::trace_set_attn(p));
  VB_CHECK_CUDA(vb::trace_set_gemm(p));
  VB_CHECK_CUDA(vb::trace_set_sampler(p));
  VB_CHECK_CUDA(vb::trace_set_snac(p));
  return 0;
}

const char* vb_last_error(void) { return vb::g_err; }
int vb_version(void) { return 100; }
int vb_set_pdl(int enabled) {
  const int old = vb::pdl_enabled(0) ? 1 : 0;
  vb::g_pdl = enabled ? 1 : 0;
  return old;
}

int vb_device_info(int* sm_count, int* max_smem_optin) {
  int dev = 0;
  VB_CHECK_CUDA(cudaGetDevice(&dev));
  if (sm_count) VB_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (max_smem_optin)
    VB_CHECK_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return 0;
}

int vb_tensor_map_kv(void* out_map, const void* d_kv, int64_t n_slabs, int page_size, int n_kv_heads,
                     int head_dim, int box_tokens) {
  VB_CHECK_ARG(out_map && d_kv, "vb_tensor_map_kv: null pointer");
  VB_CHECK_ARG(head_dim % 64 == 0, "vb_tensor_map_kv: head_dim %d must be a multiple of 64", head_dim);
  VB_CHECK_ARG(box_tokens >= 1 && box_tokens <= 256 && page_size % box_tokens == 0,
               "vb_tensor_map_kv: box_tokens %d must divide page_size %d", box_tokens, page_size);
  auto fn = vb::encode_fn();
  VB_CHECK_ARG(fn, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const cuuint64_t D = head_dim, H = n_kv_heads, P = page_size;
  // dimension order (dim, token, head, k|v, page): a box of all heads lands in shared memory as [head][token][64],
  // i.e. every head's rows are contiguous (what the attention consumers' ldmatrix pattern needs) while the
  // global footprint of the box is one contiguous run of the page
  cuuint64_t dims[5] = {D, P, H, 2, static_cast<cuuint64_t>(n_slabs)};
  cuuint64_t strides[4] = {H * D * 2, D * 2, P * H * D * 2, 2 * P * H * D * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(box_tokens), static_cast<cuuint32_t>(H), 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                  const_cast<void*>(d_kv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(kv) failed: CUresult %d", static_cast<int>(r));
  return 0;
}

int vb_tensor_map_2d_bf16(void* out_map, const void* d_base, int64_t rows, int64_t cols, int64_t ld,
                          int box_rows) {
  VB_CHECK_ARG(out_map && d_base, "vb_tensor_map_2d_bf16: null pointer");
  VB_CHECK_ARG(ld % 8 == 0, "vb_tensor_map_2d_bf16: leading dimension %lld must be a multiple of 8 elements",
               static_cast<long long>(ld));
  VB_CHECK_ARG(box_rows >= 1 && box_rows <= 256, "vb_tensor_map_2d_bf16: box_rows %d out of range", box_rows);
  auto fn = vb::encode_fn();
  VB_CHECK_ARG(fn, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(d_base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: CUresult %d", static_cast<int>(r));
  return 0;
}

}  // extern "C"
