// K3 — sharded exact dense retrieval: similarity GEMM with the top-k fused into the epilogue.
//
// Replaces ChromaDB/hnswlib ANN search (reference infomesh/index/vector_store.py:187-254,
// `collection.query(query_embeddings, n_results)`): the whole shard D[n_docs, dim] (bf16, rows
// L2-normalised so dot == cosine) is streamed once from HBM by TMA, multiplied against the resident
// query tile Q[<=128, dim] on tcgen05 (queries = MMA M rows -> TMEM lanes, documents = MMA N columns),
// and every epilogue thread keeps the running top-K of ITS query row straight out of TMEM.  The
// B x n_docs score matrix is never materialised; HBM traffic is exactly one pass over the shard.
//
//   warp 0      TMA producer  (Q once, then 128-doc x 64-dim tiles through an 8-deep ring)
//   warp 1      MMA issuer    (tcgen05.mma 128x128x16, TMEM accumulator double-buffered)
//   warps 2..5  epilogue      (tcgen05.ld -> per-thread sorted top-K in registers)
//
// Each persistent CTA writes one candidate list per query; topk_merge (below) reduces the per-CTA
// lists, and the same kernel merges the per-GPU lists after the NVLink exchange (K4).
#include <math_constants.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "../common/tmap_cache.h"

namespace im {

constexpr int kSimBM = 128;  // query rows (zero padded by TMA)
constexpr int kSimBN = 128;  // documents per tile
constexpr int kSimBK = 64;
constexpr int kSimThreads = 192;
constexpr int kSimTileBytes = kSimBN * kSimBK * 2;  // 16 KB
constexpr int kSimMaxSmem = 227 * 1024;

template <int KTOP>
struct TopK {
  float v[KTOP];
  int id[KTOP];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KTOP; ++i) {
      v[i] = -CUDART_INF_F;
      id[i] = -1;
    }
  }
  __device__ __forceinline__ float vmin() const { return v[KTOP - 1]; }
  // sorted descending; ties keep the earlier (lower id) entry because callers scan ids ascending
  __device__ __forceinline__ void insert(float s, int d) {
    v[KTOP - 1] = s;
    id[KTOP - 1] = d;
#pragma unroll
    for (int i = KTOP - 1; i > 0; --i) {
      const bool sw = v[i] > v[i - 1];
      const float tv = v[i];
      const int ti = id[i];
      v[i] = sw ? v[i - 1] : v[i];
      id[i] = sw ? id[i - 1] : id[i];
      v[i - 1] = sw ? tv : v[i - 1];
      id[i - 1] = sw ? ti : id[i - 1];
    }
  }
};

template <int KTOP>
__global__ void __launch_bounds__(kSimThreads, 1)
sim_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, int nq,
                int n_docs, int dim, int stages, const uint8_t* __restrict__ alive, float* __restrict__ out_scores,
                int* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int num_kb = dim / kSimBK;
  uint8_t* smem_q = smem;                                   // num_kb tiles of [128 x 64] bf16
  uint8_t* smem_d = smem + num_kb * (kSimBM * kSimBK * 2);  // ring of [128 docs x 64] tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_d + stages * kSimTileBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* q_bar = empty_bar + stages;
  uint64_t* tmem_full = q_bar + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const uint32_t warp = warp_id(), lane = lane_id();
  const int num_tiles = (n_docs + kSimBN - 1) / kSimBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * kSimBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // queries: resident for the whole kernel (re-read by every MMA)
      mbar_expect_tx(q_bar, num_kb * kSimBM * kSimBK * 2);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(smem_q + kb * (kSimBM * kSimBK * 2), &tmap_q, q_bar, kb * kSimBK, 0);
      const uint64_t pol = l2_policy_evict_first();  // the shard is streamed exactly once
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kSimTileBytes);
          tma_load_2d_hint(smem_d + stage * kSimTileBytes, &tmap_d, &full_bar[stage], kb * kSimBK, t * kSimBN, pol);
          if (++stage == (uint32_t)stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kSimBM, kSimBN);
      mbar_wait(q_bar, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t q0 = smem_u32(smem_q);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kSimBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a0 = q0 + kb * (kSimBM * kSimBK * 2);
          const uint32_t b0 = smem_u32(smem_d + stage * kSimTileBytes);
#pragma unroll
          for (int k = 0; k < kSimBK / 16; ++k)
            umma_bf16(d_tmem, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == (uint32_t)stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const uint32_t quad = warp & 3u;
    const int qrow = static_cast<int>(quad * 32u + lane);
    const bool warp_has_queries = static_cast<int>(quad * 32u) < nq;
    TopK<KTOP> top;
    top.init();
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp_has_queries) {
        const int doc_base = t * kSimBN;
#pragma unroll 1
        for (int c = 0; c < kSimBN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * kSimBN + c, v);
          tmem_ld_wait();
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          const bool row_ok = qrow < nq;
          // warp-uniform fast reject: no lane (query) can improve its top-K from these 32 documents
          if (__any_sync(0xffffffffu, row_ok && mx > top.vmin())) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = __uint_as_float(v[i]);
              const int d = doc_base + c + i;
              bool hit = row_ok && s > top.vmin() && d < n_docs;
              if (hit && alive != nullptr) hit = alive[d] != 0;
              // the (predicated, fully unrolled) insertion only runs for documents some lane actually wants
              if (__any_sync(0xffffffffu, hit)) {
                if (hit) top.insert(s, d);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (qrow < nq) {
      float* os = out_scores + (static_cast<size_t>(blockIdx.x) * nq + qrow) * KTOP;
      int* oi = out_ids + (static_cast<size_t>(blockIdx.x) * nq + qrow) * KTOP;
#pragma unroll
      for (int i = 0; i < KTOP; ++i) {
        os[i] = top.v[i];
        oi[i] = top.id[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * kSimBN);
}

// ------------------------------------------------------------------------------------------------
// K4 — merge P sorted candidate lists per query into one top-K (score desc, id asc), optionally
// pushing the result into every peer's gather slot over NVLink (flag-signalled), optionally waiting
// for all peers' slots first.  One CTA per query.
// ------------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;

__global__ void __launch_bounds__(kMergeThreads)
topk_merge_kernel(const float* __restrict__ cand_scores, const int64_t* __restrict__ cand_ids64,
                  const int* __restrict__ cand_ids32, int P, int nq, int k_in, int k_out, int64_t id_offset,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                  // fused exchange (all optional)
                  float* const* peer_scores, int64_t* const* peer_ids, uint32_t* const* peer_flags, int world, int rank,
                  const uint32_t* wait_flags, uint32_t epoch) {
  extern __shared__ uint8_t msm[];
  const int q = blockIdx.x;
  const int n = P * k_in;
  float* s_sc = reinterpret_cast<float*>(msm);
  int64_t* s_id = reinterpret_cast<int64_t*>(msm + ((static_cast<size_t>(n) * 4 + 7) & ~size_t(7)));
  __shared__ float red_v[kMergeThreads / 32];
  __shared__ int64_t red_i[kMergeThreads / 32];
  __shared__ int red_p[kMergeThreads / 32];
  __shared__ int win_pos;

  if (wait_flags != nullptr) {
    if (threadIdx.x < static_cast<unsigned>(world)) {
      uint32_t spins = 0;
      while (ld_acquire_sys(wait_flags + threadIdx.x) < epoch) {
        if (++spins > IM_WAIT_LIMIT) {
          printf("[infomesh_b200] topk_merge flag timeout peer=%d\n", (int)threadIdx.x);
          __trap();
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / k_in, j = i % k_in;
    const size_t src = (static_cast<size_t>(p) * nq + q) * k_in + j;
    float sc = cand_scores[src];
    int64_t id = cand_ids64 ? cand_ids64[src] : static_cast<int64_t>(cand_ids32[src]);
    if (id < 0) sc = -CUDART_INF_F;
    else if (cand_ids64 == nullptr) id += id_offset;
    s_sc[i] = sc;
    s_id[i] = id;
  }
  __syncthreads();
  for (int r = 0; r < k_out; ++r) {
    float bv = -CUDART_INF_F;
    int64_t bi = INT64_MAX;
    int bp = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float sc = s_sc[i];
      const int64_t id = s_id[i];
      if (id >= 0 && (sc > bv || (sc == bv && id < bi))) {
        bv = sc;
        bi = id;
        bp = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const int op = __shfl_xor_sync(0xffffffffu, bp, off);
      if (op >= 0 && (bp < 0 || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
        bp = op;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      red_v[threadIdx.x >> 5] = bv;
      red_i[threadIdx.x >> 5] = bi;
      red_p[threadIdx.x >> 5] = bp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kMergeThreads / 32; ++w) {
        if (red_p[w] >= 0 && (bp < 0 || red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi))) {
          bv = red_v[w];
          bi = red_i[w];
          bp = red_p[w];
        }
      }
      win_pos = bp;
      const float osc = bp >= 0 ? bv : -CUDART_INF_F;
      const int64_t oid = bp >= 0 ? bi : -1;
      if (out_scores != nullptr) {
        out_scores[static_cast<size_t>(q) * k_out + r] = osc;
        out_ids[static_cast<size_t>(q) * k_out + r] = oid;
      }
      if (peer_scores != nullptr) {
        for (int p = 0; p < world; ++p) {
          const size_t dst = (static_cast<size_t>(rank) * nq + q) * k_out + r;
          peer_scores[p][dst] = osc;
          peer_ids[p][dst] = oid;
        }
      }
      if (bp >= 0) {
        // also drop duplicates of the winning id (same document reported by two lists)
        s_id[bp] = -1;
      }
    }
    __syncthreads();
    // duplicate suppression across lists: any other candidate with the same id is retired
    if (win_pos >= 0) {
      const int64_t wid = out_ids != nullptr ? out_ids[static_cast<size_t>(q) * k_out + r] : -2;
      if (wid >= 0)
        for (int i = threadIdx.x; i < n; i += blockDim.x)
          if (s_id[i] == wid) s_id[i] = -1;
    }
    __syncthreads();
  }
  if (peer_flags != nullptr) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      // one arrival per query block; the consumer waits for nq arrivals per epoch (cumulative counter)
      for (int p = 0; p < world; ++p) {
        uint32_t* f = peer_flags[p] + rank;
        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
      }
    }
  }
}

}  // namespace im

// Per-CTA candidate lists: out_scores/out_ids are [grid, nq, ktop]; returns grid (CTA count) or <0.
IM_API int im_sim_topk(const void* Q, const void* D, int nq, int n_docs, int dim, int ldq, int ldd, int ktop,
                       const uint8_t* alive, float* out_scores, int* out_ids, int max_ctas, void* stream) {
  using namespace im;
  if (nq < 1 || nq > kSimBM) return set_error("im_sim_topk", "nq must be in [1,128]");
  if (dim % kSimBK != 0 || dim > 512) return set_error("im_sim_topk", "dim must be a multiple of 64 and <= 512");
  if (ktop != 16 && ktop != 32) return set_error("im_sim_topk", "ktop must be 16 or 32");
  const int num_kb = dim / kSimBK;
  const int q_bytes = num_kb * kSimBM * kSimBK * 2;
  int stages = (kSimMaxSmem - 1024 - 512 - q_bytes) / kSimTileBytes;
  if (stages > 12) stages = 12;
  if (stages < 2) return set_error("im_sim_topk", "not enough shared memory for the document ring");
  const int smem_bytes = q_bytes + stages * kSimTileBytes + 1024 + 512;
  const int num_tiles = (n_docs + kSimBN - 1) / kSimBN;
  int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  CUtensorMap tq, td;
  if (get_tmap_2d(&tq, Q, nq, dim, static_cast<uint64_t>(ldq) * 2, kSimBM, kSimBK, 2, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&td, D, n_docs, dim, static_cast<uint64_t>(ldd) * 2, kSimBN, kSimBK, 2, TMAP_SW_128)) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ktop == 16) {
    IM_CUDA_OK(cudaFuncSetAttribute(sim_topk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    sim_topk_kernel<16><<<grid, kSimThreads, smem_bytes, s>>>(tq, td, nq, n_docs, dim, stages, alive, out_scores,
                                                                 out_ids);
  } else {
    IM_CUDA_OK(cudaFuncSetAttribute(sim_topk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    sim_topk_kernel<32><<<grid, kSimThreads, smem_bytes, s>>>(tq, td, nq, n_docs, dim, stages, alive, out_scores,
                                                                 out_ids);
  }
  IM_LAUNCH_OK("sim_topk_kernel");
  return grid;
}

IM_API int im_topk_merge(const float* cand_scores, const int64_t* cand_ids64, const int* cand_ids32, int P, int nq,
                         int k_in, int k_out, int64_t id_offset, float* out_scores, int64_t* out_ids,
                         float* const* peer_scores, int64_t* const* peer_ids, uint32_t* const* peer_flags, int world,
                         int rank, const uint32_t* wait_flags, uint32_t epoch, void* stream) {
  using namespace im;
  if (nq <= 0) return 0;
  const size_t n = static_cast<size_t>(P) * k_in;
  const size_t smem = ((n * 4 + 7) & ~size_t(7)) + n * 8;
  if (smem > 200 * 1024) return set_error("im_topk_merge", "too many candidates per query");
  if (smem > 48 * 1024)
    IM_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<nq, kMergeThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      cand_scores, cand_ids64, cand_ids32, P, nq, k_in, k_out, id_offset, out_scores, out_ids, peer_scores, peer_ids,
      peer_flags, world, rank, wait_flags, epoch);
  IM_LAUNCH_OK("topk_merge_kernel");
  return 0;
}
