// Host-side helpers shared by all launchers: error handling, TMA tensor-map construction
// (driver entry point resolved at run time so the library links without libcuda on the
// CPU-only build box), and the exported-symbol macro.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define IM_API extern "C" __attribute__((visibility("default")))

namespace im {

// Last error text, readable from Python through im_last_error().
char* last_error_buf();

inline int set_error(const char* what, const char* detail) {
  snprintf(last_error_buf(), 512, "%s: %s", what, detail ? detail : "");
  return -1;
}

#define IM_CUDA_OK(expr)                                                 \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return im::set_error(#expr, cudaGetErrorString(_e)); \
  } while (0)

#define IM_LAUNCH_OK(name)                                                        \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) return im::set_error(name, cudaGetErrorString(_e));    \
  } while (0)

enum TmapSwizzle { TMAP_SW_NONE = 0, TMAP_SW_32 = 1, TMAP_SW_64 = 2, TMAP_SW_128 = 3 };

// 2-D row-major tensor [rows, cols] of `elem_bytes` elements with row pitch `row_stride_bytes`;
// box = [box_rows, box_cols].  Out-of-bounds elements are zero-filled on load.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes, TmapSwizzle swizzle);

int sm_count();

// Debug build of the peer channels, switched at run time: INFOMESH_B200_POISON_SLOTS=1 makes every exchange consumer
// overwrite the receive slots it has consumed with a poison pattern and trap when it ever READS that pattern -- which can
// only happen if an arrival counter reported a delivery that has not landed (a flag-protocol or memory-ordering bug the
// executable protocol model in parallel/sim.py cannot see).  Costs one extra store per consumed element; off by default.
int debug_poison_slots();

// Programmatic dependent launch (PDL): the kernel may start while its stream predecessor is still draining; it runs
// its prologue (barrier init, TMEM alloc, descriptor prefetch) and blocks at pdl_wait() (ptx.cuh) until the predecessor
// has completed and flushed.  Captured into CUDA graphs as programmatic dependency edges.  INFOMESH_B200_PDL=0
// turns it off (plain stream-ordered launches).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

// launch_pdl with a thread-block cluster of `cluster_x` consecutive CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

}  // namespace im
