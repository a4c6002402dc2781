// Peer-memory communication backend (SURVEY C-0): a symmetric heap built on CUDA IPC and the small device-side
// primitives that move data through it.
//
//   * im_symm_alloc / im_symm_open / im_symm_close / im_symm_free — every rank cudaMallocs one slab, exports an
//     IPC handle, and maps every peer's slab; the same offset addresses "the same object" on every rank.
//   * p2p_allgather_kernel — push-based all-gather of a small block (top-k lists, logits, fingerprints): each CTA
//     stores its slice into slot[rank] of every peer with 16-byte NVLink writes, publishes an arrival count with
//     red.release.sys, waits for every peer's count, then copies the gathered slots to a private output.  One
//     launch, no host involvement, CUDA-graph safe (the step counter lives in device memory).
//   * p2p_barrier_kernel — device barrier across ranks.
//
// Flag protocol: every channel owns a private {step, done} pair in local memory and `world` arrival counters in the
// symmetric heap.  Counters only grow: a channel whose producers arrive `c` times per use is complete for use s when
// flag >= (s + 1) * c.  The consuming kernel advances `step` itself (last CTA out), so a channel may be used any
// number of times per engine step, or skipped, without a host-side epoch.  Payload buffers are double-buffered on
// (s & 1); a peer cannot run two uses ahead because finishing use s + 1 needs this rank's use s + 1 arrivals, which
// are stream-ordered after this rank's use s reads.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace im {

constexpr int kAgThreads = 256;

__device__ __forceinline__ void flag_wait(const uint32_t* flag, uint32_t target, const char* what, int who) {
  uint32_t spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(flag) - target) < 0) {
    if (++spins > IM_WAIT_LIMIT) {
      printf("[infomesh_b200] %s timeout (peer %d, have %u want %u)\n", what, who, ld_acquire_sys(flag), target);
      __trap();
    }
    __nanosleep(20);
  }
}

// Last CTA out advances the channel: every CTA has read state[0] before it can have counted itself done.
__device__ __forceinline__ void channel_advance(uint32_t* state) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1u) == gridDim.x * gridDim.y - 1u) {
      state[1] = 0u;
      state[0] += 1u;
      __threadfence();
    }
  }
}

// src: this rank's block (`bytes`, multiple of 16).  peer_buf[p]: base of rank p's receive area
// [2][world][bytes]; peer_flag[p]: rank p's counters [world].  out: [world][bytes] private copy.
__global__ void __launch_bounds__(kAgThreads)
p2p_allgather_kernel(const uint8_t* __restrict__ src, size_t bytes, uint8_t* const* __restrict__ peer_buf,
                     uint32_t* const* __restrict__ peer_flag, uint32_t* __restrict__ state, int world, int rank,
                     uint8_t* __restrict__ out, int debug_poison) {
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(state);
  const size_t par_off = static_cast<size_t>(step & 1u) * world * bytes;
  const size_t n16 = bytes / 16;
  const size_t per_cta = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per_cta, hi = min(n16, lo + per_cta);
  // ---- push my slice to every peer (own rank included: the local copy takes the same path)
  for (int pp = 0; pp < world; ++pp) {
    const int p = (rank + pp) % world;   // stagger destinations so ranks do not all hit peer 0 first
    uint4* dst = reinterpret_cast<uint4*>(peer_buf[p] + par_off + static_cast<size_t>(rank) * bytes);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = s4[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < static_cast<unsigned>(world)) {
    uint32_t* f = peer_flag[threadIdx.x] + rank;
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
  }
  // ---- wait until every rank's gridDim.x CTAs have arrived here, then copy out my slice of every slot
  const uint32_t target = (step + 1u) * gridDim.x;
  if (threadIdx.x < static_cast<unsigned>(world)) flag_wait(peer_flag[rank] + threadIdx.x, target, "p2p_allgather", threadIdx.x);
  __syncthreads();
  const uint8_t* mine = peer_buf[rank] + par_off;
  for (int p = 0; p < world; ++p) {
    const uint4* s4 = reinterpret_cast<const uint4*>(mine + static_cast<size_t>(p) * bytes);
    uint4* d4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * bytes);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const uint4 w = s4[i];
      if (debug_poison) {   // INFOMESH_B200_POISON_SLOTS=1: consumed slots are poisoned; meeting poison = flag said "arrived" too early
        if (w.x == 0xdeadbeefu && w.y == 0xdeadbeefu && w.z == 0xdeadbeefu && w.w == 0xdeadbeefu) {
          printf("[infomesh_b200] p2p_allgather: consumed a POISONED slot (step %u, peer %d, word %llu)\n", step, p, (unsigned long long)i);
          __trap();
        }
        const_cast<uint4*>(s4)[i] = make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu);
      }
      d4[i] = w;
    }
  }
  channel_advance(state);
}

__global__ void p2p_barrier_kernel(uint32_t* const* __restrict__ peer_flag, uint32_t* __restrict__ state, int world,
                                   int rank) {
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(state);
  if (threadIdx.x < static_cast<unsigned>(world)) {
    __threadfence_system();
    uint32_t* f = peer_flag[threadIdx.x] + rank;
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
    flag_wait(peer_flag[rank] + threadIdx.x, step + 1u, "p2p_barrier", threadIdx.x);
  }
  channel_advance(state);
}

}  // namespace im

using namespace im;

// ---------------------------------------------------------------- symmetric heap (CUDA IPC)
IM_API int im_symm_alloc(size_t bytes, void** ptr_out, uint8_t* handle_out /* 64 bytes */) {
  void* p = nullptr;
  IM_CUDA_OK(cudaMalloc(&p, bytes));
  IM_CUDA_OK(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_error("cudaIpcGetMemHandle", cudaGetErrorString(e));
  }
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return 0;
}

IM_API int im_symm_open(const uint8_t* handle, void** ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  IM_CUDA_OK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

IM_API int im_symm_close(void* ptr) {
  IM_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

IM_API int im_symm_free(void* ptr) {
  IM_CUDA_OK(cudaFree(ptr));
  return 0;
}

IM_API int im_can_access_peer(int dev, int peer) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, dev, peer) != cudaSuccess) return 0;
  return ok;
}

// ---------------------------------------------------------------- device primitives
IM_API int im_p2p_allgather(const void* src, size_t bytes, void* const* peer_buf, uint32_t* const* peer_flag,
                            uint32_t* state, int world, int rank, void* out, int ctas, void* stream) {
  if (bytes == 0 || (bytes % 16) != 0) return set_error("im_p2p_allgather", "block size must be a non-zero multiple of 16 bytes");
  if (world < 1 || world > 32) return set_error("im_p2p_allgather", "world must be 1..32");
  const size_t n16 = bytes / 16;
  int grid = ctas > 0 ? ctas : static_cast<int>((n16 + kAgThreads * 4 - 1) / (kAgThreads * 4));
  if (grid < 1) grid = 1;
  if (grid > 32) grid = 32;   // every CTA spins on peers: stay far below one wave so all are co-resident
  p2p_allgather_kernel<<<grid, kAgThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), bytes, reinterpret_cast<uint8_t* const*>(peer_buf), peer_flag, state, world,
      rank, static_cast<uint8_t*>(out), debug_poison_slots());
  IM_LAUNCH_OK("p2p_allgather_kernel");
  return grid;
}

IM_API int im_p2p_barrier(uint32_t* const* peer_flag, uint32_t* state, int world, int rank, void* stream) {
  if (world < 1 || world > 32) return set_error("im_p2p_barrier", "world must be 1..32");
  p2p_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(peer_flag, state, world, rank);
  IM_LAUNCH_OK("p2p_barrier_kernel");
  return 0;
}

