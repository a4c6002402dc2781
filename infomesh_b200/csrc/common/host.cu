#include <cstdlib>
#include "host.h"

#include <stdlib.h>

#include <mutex>

namespace im {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

static PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes, TmapSwizzle swizzle) {
  auto encode = resolve_encode();
  if (!encode) return set_error("make_tmap_2d", "cuTensorMapEncodeTiled not available");
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle == TMAP_SW_128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == TMAP_SW_64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == TMAP_SW_32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = encode(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "CUresult=%d rows=%llu cols=%llu stride=%llu box=[%u,%u] base=%p", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_rows,
             box_cols, base);
    return set_error("cuTensorMapEncodeTiled", msg);
  }
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int debug_poison_slots() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("INFOMESH_B200_POISON_SLOTS");
    on = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 1 : 0;
  }
  return on;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("INFOMESH_B200_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

}  // namespace im

IM_API void im_set_pdl(int on) { im::g_pdl = on ? 1 : 0; }

IM_API const char* im_last_error() { return im::last_error_buf(); }
IM_API int im_sm_count() { return im::sm_count(); }
IM_API int im_abi_version() { return 1; }
