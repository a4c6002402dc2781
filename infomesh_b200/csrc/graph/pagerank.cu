// K13 — domain-level PageRank power iteration on the GPU (offline; reference infomesh/index/link_graph.py:206-235 runs
// it as Python dict loops).  Edges are COO (src, dst, share) with share = w / out_weight(src) precomputed once;
// one step is an edge-parallel scatter  nxt[dst] += damping * score[src] * share  onto a vector pre-filled with
// (1 - damping) / n, followed by a fused L1-delta reduction so the host checks convergence with one scalar read.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/host.h"

namespace im {

__global__ void __launch_bounds__(256)
pagerank_scatter_kernel(const int* __restrict__ src, const int* __restrict__ dst, const float* __restrict__ share,
                        long long n_edges, const float* __restrict__ score, float* __restrict__ nxt, float damping) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += stride)
    atomicAdd(nxt + dst[e], damping * score[src[e]] * share[e]);
}

__global__ void __launch_bounds__(256)
l1_delta_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += fabsf(a[i] - b[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

}  // namespace im

// nxt must be pre-filled with (1 - damping) / n; delta_out (one float) must be zero.
IM_API int im_pagerank_step(const int* src, const int* dst, const float* share, long long n_edges, int n_nodes,
                            const float* score, float* nxt, float damping, float* delta_out, void* stream) {
  using namespace im;
  if (n_nodes <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n_edges > 0) {
    long long blocks = (n_edges + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    pagerank_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(src, dst, share, n_edges, score, nxt, damping);
    IM_LAUNCH_OK("pagerank_scatter_kernel");
  }
  if (delta_out != nullptr) {
    int blocks = (n_nodes + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    l1_delta_kernel<<<blocks, 256, 0, s>>>(score, nxt, n_nodes, delta_out);
    IM_LAUNCH_OK("l1_delta_kernel");
  }
  return 0;
}
