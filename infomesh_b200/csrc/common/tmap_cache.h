// Small process-wide cache of TMA tensor maps so steady-state launches (static buffers under
// CUDA graphs, model weights) never re-encode on the host.
#pragma once
#include <mutex>
#include <unordered_map>

#include "host.h"

namespace im {

struct TmapKey {
  const void* base;
  uint64_t rows, cols, stride;
  uint32_t box_rows, box_cols;
  int elem_bytes, swizzle;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && stride == o.stride && box_rows == o.box_rows &&
           box_cols == o.box_cols && elem_bytes == o.elem_bytes && swizzle == o.swizzle;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 31 + k.stride * 131 + k.box_rows * 17 + k.box_cols * 7 + k.elem_bytes * 3 + k.swizzle);
    return static_cast<size_t>(h);
  }
};

inline int get_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                       uint32_t box_rows, uint32_t box_cols, int elem_bytes, TmapSwizzle swizzle) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{base, rows, cols, row_stride_bytes, box_rows, box_cols, elem_bytes, static_cast<int>(swizzle)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  int rc = make_tmap_2d(out, base, rows, cols, row_stride_bytes, box_rows, box_cols, elem_bytes, swizzle);
  if (rc != 0) return rc;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

}  // namespace im
