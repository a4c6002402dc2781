// K3 — sharded exact dense retrieval: similarity GEMM with the top-k fused into the epilogue.
//
// Replaces ChromaDB/hnswlib ANN search (reference infomesh/index/vector_store.py:187-254,
// `collection.query(query_embeddings, n_results)`): the whole shard D[n_docs, dim] (bf16, rows
// L2-normalised so dot == cosine) is streamed once from HBM by TMA, multiplied against the resident
// query tile Q[<=128, dim] on tcgen05 (documents = MMA M rows -> TMEM lanes, queries = MMA N columns);
// every epilogue thread owns one document row, compares its scores against per-query thresholds in smem and,
// on the rare hit, the warp cooperatively inserts into the CTA's sorted per-query list (lane i = entry i).  The
// B x n_docs score matrix is never materialised; HBM traffic is exactly one pass over the shard.
//
//   warp 0      TMA producer  (Q once, then 128-doc x 64-dim tiles through an 8-deep ring)
//   warp 1      MMA issuer    (tcgen05.mma 128x128x16, TMEM accumulator double-buffered)
//   warps 2..5  epilogue      (tcgen05.ld -> threshold filter -> warp-cooperative sorted insert in smem)
//
// Each persistent CTA writes one candidate list per query; topk_merge (below) reduces the per-CTA
// lists, and the same kernel merges the per-GPU lists after the NVLink exchange (K4).
#include <math_constants.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "../common/tmap_cache.h"

namespace im {

constexpr int kSimBM = 128;  // documents per tile (MMA M rows -> TMEM lanes)
constexpr int kSimBK = 64;
constexpr int kSimThreads = 192;        // NPAD == 16: producer + MMA + 4 epilogue warps
constexpr int kSimThreadsWide = 320;    // NPAD >= 32: 8 epilogue warps, two per TMEM quadrant, each owning half of the query columns
// Why 8: with one epilogue warp per scheduler the scan is bound by the epilogue's dependent-instruction latency (ncu: 16 % issue
// utilisation, ~9 cycles per issued instruction, ~700 instructions per warp and tile), not by HBM -- an fp8 shard streamed no
// faster than a bf16 one.  Two warps per scheduler halve the per-warp work and hide each other's latency.
template <int NPAD> constexpr int sim_threads() { return NPAD >= 32 ? kSimThreadsWide : kSimThreads; }
constexpr int kSimTileBytes = kSimBM * kSimBK * 2;  // 16 KB
constexpr int kSimMaxSmem = 227 * 1024;

// Warp-cooperative insert of (s, d) into THIS WARP's sorted list of query j (score desc, id asc): lane i holds
// entry i.  Lists are private to a warp (no locks); `thr[j]` is a CTA-shared lower bound (max of the warps'
// K-th values, updated with a benign race) used only to reject candidates early.
__device__ __forceinline__ void list_insert(volatile float* lv, volatile int* li, volatile float* thr, int K, int j,
                                            float s, int d, uint32_t lane) {
  const bool in = static_cast<int>(lane) < K;
  float v = in ? lv[j * K + lane] : -CUDART_INF_F;
  int id = in ? li[j * K + lane] : 0x7fffffff;
  const bool before = in && ((v > s) || (v == s && id >= 0 && id < d));
  const int pos = __popc(__ballot_sync(0xffffffffu, before));
  if (pos < K) {
    const float vu = __shfl_up_sync(0xffffffffu, v, 1);
    const int iu = __shfl_up_sync(0xffffffffu, id, 1);
    if (static_cast<int>(lane) == pos) {
      v = s;
      id = d;
    } else if (static_cast<int>(lane) > pos) {
      v = vu;
      id = iu;
    }
    if (in && static_cast<int>(lane) >= pos) {
      lv[j * K + lane] = v;
      li[j * K + lane] = id;
    }
    if (static_cast<int>(lane) == K - 1 && v > thr[j]) thr[j] = v;
  }
  __syncwarp();
}

// NPAD = queries padded to the MMA N granule (16); documents are the MMA M dimension so that every epilogue
// thread owns ONE document row and compares its NPAD scores against per-query thresholds held in smem.
//
// F8: the shard is e4m3 with one fp32 scale per document row (half the HBM bytes of bf16 -- the pass is HBM-bound, so
// about twice the documents per second); queries arrive as e4m3 + one scale each (emitted by pool_norm).  A k-block is
// still 128 bytes of K (128 elements instead of 64), the MMA is kind::f8f6f4 (K = 32 per instruction, same 32-byte
// descriptor step), and the epilogue turns the raw accumulator into the true score acc * d_scale[doc] * q_scale[j]
// before the threshold filter, so lists, thresholds and the merge are unchanged.
template <int NPAD, bool F8>
__global__ void __launch_bounds__(sim_threads<NPAD>(), 1)
sim_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, int nq,
                int n_docs, int dim, int stages, int ktop, const uint8_t* __restrict__ alive,
                float* __restrict__ out_scores, int* __restrict__ out_ids, const float* __restrict__ thr_init,
                int thr_stride, const float* __restrict__ d_scale, const float* __restrict__ q_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kBKe = F8 ? 2 * kSimBK : kSimBK;   // K elements per 128-byte k-block
  const int num_kb = dim / kBKe;
  constexpr int kQTileBytes = NPAD * kSimBK * 2;
  uint8_t* smem_q = smem;                          // num_kb tiles of [NPAD x 64] bf16 (resident)
  uint8_t* smem_d = smem + num_kb * kQTileBytes;   // ring of [128 docs x 64] tiles; NPAD*128 B keeps 1024 alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_d + stages * kSimTileBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* q_bar = empty_bar + stages;
  uint64_t* tmem_full = q_bar + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* thr = reinterpret_cast<float*>(tmem_slot + 2);   // [NPAD] CTA-wide lower bound of the K-th best per query
  float* qsc = thr + NPAD;                                  // [NPAD] per-query dequantisation scale (F8)
  float* thr_cmp = F8 ? qsc + NPAD : thr;                   // [NPAD] what the filter compares against: thr (bf16) or thr / q_scale (F8)
  float* list_v = reinterpret_cast<float*>(qsc + 2 * NPAD); // [4 quadrants][nq][ktop]
  int* list_i = reinterpret_cast<int*>(list_v + 4 * nq * ktop);

  constexpr uint32_t kTmemCols = (2 * NPAD) < 32 ? 32 : (2 * NPAD);
  constexpr int kThreads = sim_threads<NPAD>();
  constexpr int kEpiWarps = (kThreads - 64) / 32;        // 4 or 8
  constexpr int kColsPerWarp = NPAD / (kEpiWarps / 4);   // query columns owned by one epilogue warp: NPAD or NPAD / 2
  const uint32_t warp = warp_id(), lane = lane_id();
  const int num_tiles = (n_docs + kSimBM - 1) / kSimBM;

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
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // thr_init[j]: a known lower bound of query j's K-th best score (the K-th best of a sample, see ops/search.py).
  // One ulp below it, so the sampled document that defines the bound still passes the strict `>` filter.
  for (int i = threadIdx.x; i < NPAD; i += kThreads) {
    float t0 = -CUDART_INF_F;
    if (thr_init != nullptr && i < nq) {
      const float b = thr_init[static_cast<size_t>(i) * thr_stride];
      if (b > -CUDART_INF_F && b == b) t0 = __uint_as_float(b > 0.f ? __float_as_uint(b) - 1u : (b < 0.f ? __float_as_uint(b) + 1u : 0x80000001u));
    }
    thr[i] = t0;
    const float qs_i = (F8 && i < nq) ? q_scale[i] : 1.0f;
    qsc[i] = qs_i;
    if constexpr (F8) thr_cmp[i] = t0 / qs_i;      // -inf stays -inf (scales are positive powers of two)
  }
  for (int i = threadIdx.x; i < 4 * nq * ktop; i += kThreads) {
    list_v[i] = -CUDART_INF_F;
    list_i[i] = -1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_bar, num_kb * kQTileBytes);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem_q + kb * kQTileBytes, &tmap_q, q_bar, kb * kBKe, 0);
      const uint64_t pol = l2_policy_evict_first();  // the shard is streamed exactly once
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kSimTileBytes);
          tma_load_2d_hint(smem_d + stage * kSimTileBytes, &tmap_d, &full_bar[stage], kb * kBKe, t * kSimBM, pol);
          if (++stage == (uint32_t)stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = F8 ? umma_idesc_f8(kSimBM, NPAD) : umma_idesc_f16(kSimBM, NPAD);
      mbar_wait(q_bar, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t q0 = smem_u32(smem_q);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NPAD;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem_d + stage * kSimTileBytes);  // documents: A operand
          const uint32_t b0 = q0 + kb * kQTileBytes;                      // queries:   B operand
#pragma unroll
          for (int k = 0; k < kSimBK / 16; ++k) {   // 4 x 32 bytes of K: 16 bf16 or 32 e4m3 elements per MMA
            if constexpr (F8)
              umma_f8(d_tmem, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(d_tmem, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          }
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
  } else if (ktop == 0) {
    // ---- threshold pre-pass: per-thread running maximum of every query column this warp owns, no candidate lists.  Each
    // epilogue warp ends up with the best score of the documents it saw (a distinct document per quadrant and query), which
    // is all sample_threshold() needs; branch-free FMNMX instead of the insert path, so the pass stays HBM-bound.
    const uint32_t quad = warp & 3u;
    const int col0 = static_cast<int>((warp - 2u) >> 2) * kColsPerWarp;      // first query column of this warp
    constexpr int kChunk = kColsPerWarp < 32 ? 16 : 32;
    float rmax[kColsPerWarp];
#pragma unroll
    for (int i = 0; i < kColsPerWarp; ++i) rmax[i] = -CUDART_INF_F;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int doc = t * kSimBM + static_cast<int>(quad * 32u + lane);
      bool doc_ok = doc < n_docs;
      if (doc_ok && alive != nullptr) doc_ok = alive[doc] != 0;
      const float ds = (F8 && doc < n_docs) ? d_scale[doc] : 1.0f;
#pragma unroll
      for (int c = 0; c < kColsPerWarp; c += kChunk) {
        uint32_t v[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * NPAD + col0 + c, v);
        else tmem_ld_32x32b_x16(tmem_base + ((quad * 32u) << 16) + acc * NPAD + col0 + c, v);
        tmem_ld_wait();
        if (doc_ok) {
#pragma unroll
          for (int i = 0; i < kChunk; ++i) {
            float x = __uint_as_float(v[i]);
            if constexpr (F8) x *= ds;          // the per-query scale is a positive constant per column: applied once, at the end
            rmax[c + i] = fmaxf(rmax[c + i], x);
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
    // one list entry per (quadrant, query): out is [grid * 4, nq, 1]; the id only has to be unique and non-negative
    const int slot = blockIdx.x * 4 + static_cast<int>(quad);
#pragma unroll
    for (int j = 0; j < kColsPerWarp; ++j) {
      float m = rmax[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const int jq = col0 + j;
      if (lane == 0 && jq < nq) {
        if constexpr (F8) m *= qsc[jq];
        out_scores[static_cast<size_t>(slot) * nq + jq] = m;
        out_ids[static_cast<size_t>(slot) * nq + jq] = m > -CUDART_INF_F ? slot : -1;
      }
    }
  } else {
    const uint32_t quad = warp & 3u;
    const int col0 = static_cast<int>((warp - 2u) >> 2) * kColsPerWarp;
    float* my_v = list_v + quad * nq * ktop;     // the two warps of a quadrant own disjoint query ranges of the same list block
    int* my_i = list_i + quad * nq * ktop;
    uint32_t acc = 0, acc_phase = 0;
    constexpr int kChunk = kColsPerWarp < 32 ? 16 : 32;  // TMEM columns per load
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int doc = t * kSimBM + static_cast<int>(quad * 32u + lane);
      bool doc_ok = doc < n_docs;
      if (doc_ok && alive != nullptr) doc_ok = alive[doc] != 0;
      const float ds = (F8 && doc < n_docs) ? d_scale[doc] : 1.0f;
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += kChunk) {
        uint32_t v[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * NPAD + col0 + c, v);
        else tmem_ld_32x32b_x16(tmem_base + ((quad * 32u) << 16) + acc * NPAD + col0 + c, v);
        tmem_ld_wait();
        if constexpr (F8) {   // raw accumulator x document scale; the per-query scale lives in the threshold (thr / q_scale)
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * ds);
        }
#pragma unroll
        for (int g = 0; g < kChunk; g += 8) {
          if (col0 + c + g >= nq) break;  // warp-uniform: padded query columns
          // F8: thr_cmp = thr / q_scale, kept beside thr (list_insert updates thr; the compare copy is refreshed on a hit)
          const float4 t0 = *reinterpret_cast<const float4*>(thr_cmp + col0 + c + g);
          const float4 t1 = *reinterpret_cast<const float4*>(thr_cmp + col0 + c + g + 4);
          const float th[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
          uint32_t mask = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) mask |= (__uint_as_float(v[g + q]) > th[q]) ? (1u << q) : 0u;
          if (!doc_ok) mask = 0;
          if (__any_sync(0xffffffffu, mask != 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = col0 + c + g + q;
              uint32_t hits = __ballot_sync(0xffffffffu, (mask >> q) & 1u);
              if (j >= nq) hits = 0;
              while (hits) {
                const int src = __ffs(hits) - 1;
                hits &= hits - 1;
                float sc = __shfl_sync(0xffffffffu, __uint_as_float(v[g + q]), src);
                if constexpr (F8) sc *= qsc[j];                       // true score only for the rare candidate
                const int d = __shfl_sync(0xffffffffu, doc, src);
                if (sc > *reinterpret_cast<volatile float*>(thr + j)) {
                  list_insert(my_v, my_i, thr, ktop, j, sc, d, lane);
                  if constexpr (F8) {
                    if (lane == 0) thr_cmp[j] = *reinterpret_cast<volatile float*>(thr + j) / qsc[j];
                    __syncwarp();
                  }
                }
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
    // all epilogue warps are done inserting -> fold the per-quadrant lists into quadrant 0's (epilogue warp e owns the
    // queries j == e mod kEpiWarps), then publish one candidate list per query for this CTA
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    for (int j = static_cast<int>(warp) - 2; j < nq; j += kEpiWarps) {
      for (int w = 1; w < 4; ++w) {
        for (int e = 0; e < ktop; ++e) {
          const float sc = reinterpret_cast<volatile float*>(list_v)[(w * nq + j) * ktop + e];
          const int d = reinterpret_cast<volatile int*>(list_i)[(w * nq + j) * ktop + e];
          if (d < 0) break;
          const float kth = reinterpret_cast<volatile float*>(list_v)[j * ktop + ktop - 1];
          if (sc < kth) break;  // sources are sorted: nothing further can enter
          list_insert(list_v, list_i, thr, ktop, j, sc, d, lane);
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    const int et = static_cast<int>(threadIdx.x) - 64;
    for (int i = et; i < nq * ktop; i += kEpiWarps * 32) {
      out_scores[static_cast<size_t>(blockIdx.x) * nq * ktop + i] = list_v[i];
      out_ids[static_cast<size_t>(blockIdx.x) * nq * ktop + i] = list_i[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// K4 — merge P sorted candidate lists per query into one top-K (score desc, id asc), optionally
// pushing the result into every peer's gather slot over NVLink (flag-signalled), optionally waiting
// for all peers' slots first.  One CTA per query.
// ------------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;
constexpr int64_t kPoisonId = -0x6b6b6b6b6b6b6b6bLL;   // what a consumed receive slot holds in the debug build of the channel

__global__ void __launch_bounds__(kMergeThreads)
topk_merge_kernel(const float* __restrict__ cand_scores, const int64_t* __restrict__ cand_ids64,
                  const int* __restrict__ cand_ids32, int P, int nq, int k_in, int k_out, int64_t id_offset,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                  // fused exchange (all optional)
                  float* const* peer_scores, int64_t* const* peer_ids, uint32_t* const* peer_flags, int world, int rank,
                  const uint32_t* wait_flags, uint32_t* state, uint32_t* status, uint32_t wait_limit, int debug_poison) {
  extern __shared__ uint8_t msm[];
  const int q = blockIdx.x;
  const int n = P * k_in;
  float* s_sc = reinterpret_cast<float*>(msm);
  int64_t* s_id = reinterpret_cast<int64_t*>(msm + ((static_cast<size_t>(n) * 4 + 7) & ~size_t(7)));
  __shared__ float red_v[kMergeThreads / 32];
  __shared__ int64_t red_i[kMergeThreads / 32];
  __shared__ int red_p[kMergeThreads / 32];
  __shared__ int win_pos;
  __shared__ int64_t win_id;

  // Exchange protocol (see comm/symm.cu): receive areas are [2][world][nq][k] double-buffered on step parity,
  // arrival counters are cumulative: one arrival per query block per step.
  const uint32_t step = state != nullptr ? *reinterpret_cast<volatile uint32_t*>(state) : 0u;
  __shared__ uint32_t dead_mask;  // bit p: shard p did not deliver in time -> its list is ignored ("search is never blocked")
  if (threadIdx.x == 0) dead_mask = status != nullptr ? *reinterpret_cast<volatile uint32_t*>(status) : 0u;
  __syncthreads();
  if (wait_flags != nullptr) {
    if (threadIdx.x < static_cast<unsigned>(P) && !((dead_mask >> threadIdx.x) & 1u)) {
      const uint32_t target = (step + 1u) * static_cast<uint32_t>(nq);
      const uint32_t limit = wait_limit != 0u ? wait_limit : IM_WAIT_LIMIT;
      uint32_t spins = 0;
      while (static_cast<int32_t>(ld_acquire_sys(wait_flags + threadIdx.x) - target) < 0) {
        if (++spins > limit) {
          if (status == nullptr) {
            printf("[infomesh_b200] topk_merge flag timeout peer=%d\n", (int)threadIdx.x);
            __trap();
          }
          // degraded mode: remember the silent shard (sticky until the host clears the word) and answer without it
          atomicOr(&dead_mask, 1u << threadIdx.x);
          atomicOr(status, 1u << threadIdx.x);
          break;
        }
        __nanosleep(20);
      }
    }
    __syncthreads();
    const size_t par = static_cast<size_t>(step & 1u) * P * nq * k_in;
    cand_scores += par;
    if (cand_ids64 != nullptr) cand_ids64 += par;
    if (cand_ids32 != nullptr) cand_ids32 += par;
  }
  // ---- load + compact: with the threshold filter most per-CTA lists are (nearly) empty, so only live candidates are kept
  // (positions [0, n_live) of the smem arrays); the selection rounds below then touch tens of entries instead of P * k_in.
  __shared__ int n_live_s;
  __shared__ int warp_cnt[kMergeThreads / 32];
  __shared__ float w_sc[32];          // winners, written out (locally and to every peer) after the last round
  __shared__ int64_t w_id[32];
  if (threadIdx.x == 0) n_live_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kMergeThreads) {
    const int i = base + static_cast<int>(threadIdx.x);
    float sc = -CUDART_INF_F;
    int64_t id = -1;
    if (i < n) {
      const int p = i / k_in, j = i % k_in;
      const size_t src = (static_cast<size_t>(p) * nq + q) * k_in + j;
      sc = cand_scores[src];
      id = cand_ids64 ? cand_ids64[src] : static_cast<int64_t>(cand_ids32[src]);
      if (debug_poison && wait_flags != nullptr && cand_ids64 != nullptr && !((dead_mask >> p) & 1u)) {
        // debug build of the channel (INFOMESH_B200_POISON_SLOTS=1): every slot is poisoned once consumed, so reading poison
        // means the arrival counter said "delivered" for a slot no producer has rewritten -- a protocol or ordering bug
        if (id == kPoisonId) {
          printf("[infomesh_b200] topk exchange: consumed a POISONED slot (step %u, peer %d, query %d, entry %d)\n", step, p, q, j);
          __trap();
        }
        const_cast<int64_t*>(cand_ids64)[src] = kPoisonId;
        const_cast<float*>(cand_scores)[src] = __uint_as_float(0x7fc0deadu);
      }
      if (wait_flags != nullptr && ((dead_mask >> p) & 1u)) id = -1;
      if (id >= 0 && cand_ids64 == nullptr) id += id_offset;
    }
    const bool live = id >= 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, live);
    const int wid = static_cast<int>(threadIdx.x >> 5), ln = static_cast<int>(threadIdx.x & 31);
    if (ln == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int off = n_live_s;
    for (int w = 0; w < wid; ++w) off += warp_cnt[w];
    if (live) {
      const int pos = off + __popc(bal & ((1u << ln) - 1u));
      s_sc[pos] = sc;
      s_id[pos] = id;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < kMergeThreads / 32; ++w) tot += warp_cnt[w];
      n_live_s += tot;
    }
    __syncthreads();
  }
  const int n_live = n_live_s;
  for (int r = 0; r < k_out; ++r) {
    float bv = -CUDART_INF_F;
    int64_t bi = INT64_MAX;
    int bp = -1;
    for (int i = threadIdx.x; i < n_live; i += blockDim.x) {
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
      win_id = bp >= 0 ? bi : -1;
      w_sc[r] = bp >= 0 ? bv : -CUDART_INF_F;
      w_id[r] = bp >= 0 ? bi : -1;
      if (bp >= 0) s_id[bp] = -1;
    }
    __syncthreads();
    // duplicate suppression across lists: any other candidate with the same id is retired
    if (win_pos >= 0) {
      const int64_t wid = win_id;
      for (int i = threadIdx.x; i < n_live; i += blockDim.x)
        if (s_id[i] == wid) s_id[i] = -1;
    }
    __syncthreads();
    if (win_pos < 0) {        // candidates exhausted: the remaining ranks are empty
      for (int rr = r + 1 + static_cast<int>(threadIdx.x); rr < k_out; rr += blockDim.x) {
        w_sc[rr] = -CUDART_INF_F;
        w_id[rr] = -1;
      }
      __syncthreads();
      break;
    }
  }
  // ---- publish: the local result and, for the fused exchange, slot[rank] of every peer's receive area -- all threads store
  for (int idx = threadIdx.x; idx < k_out * (1 + (peer_scores != nullptr ? world : 0)); idx += blockDim.x) {
    const int r = idx % k_out, dst_rank = idx / k_out - 1;       // -1: local output
    if (dst_rank < 0) {
      if (out_scores != nullptr) {
        out_scores[static_cast<size_t>(q) * k_out + r] = w_sc[r];
        out_ids[static_cast<size_t>(q) * k_out + r] = w_id[r];
      }
    } else {
      const size_t dst = ((static_cast<size_t>(step & 1u) * world + rank) * nq + q) * k_out + r;
      peer_scores[dst_rank][dst] = w_sc[r];
      peer_ids[dst_rank][dst] = w_id[r];
    }
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
  if (wait_flags != nullptr && state != nullptr) {
    // consumer side: the last query block out advances the channel (see comm/symm.cu)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(state + 1, 1u) == gridDim.x - 1u) {
        state[1] = 0u;
        state[0] += 1u;
        __threadfence();
      }
    }
  }
}

}  // namespace im

// Per-CTA candidate lists: out_scores/out_ids are [grid, nq, ktop]; returns grid (CTA count) or <0.
template <int NPAD, bool F8 = false>
static int launch_sim(const void* Q, const void* D, int nq, int n_docs, int dim, int ldq, int ldd, int ktop,
                      const uint8_t* alive, float* out_scores, int* out_ids, int max_ctas, const float* thr_init,
                      int thr_stride, cudaStream_t s, const float* d_scale = nullptr, const float* q_scale = nullptr) {
  using namespace im;
  constexpr int kEB = F8 ? 1 : 2;                  // bytes per element
  constexpr int kBKe = 128 / kEB;                  // elements per 128-byte k-block
  const int num_kb = dim / kBKe;
  const int q_bytes = num_kb * NPAD * kSimBK * 2;
  const int misc = 1024 /*align*/ + 512 /*barriers*/ + NPAD * 12 + 4 * nq * ktop * 8;
  int stages = (kSimMaxSmem - misc - q_bytes) / kSimTileBytes;
  if (stages > 12) stages = 12;
  if (stages < 2) return set_error("im_sim_topk", "not enough shared memory for the document ring");
  const int smem_bytes = q_bytes + stages * kSimTileBytes + misc;
  const int num_tiles = (n_docs + kSimBM - 1) / kSimBM;
  int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  CUtensorMap tq, td;
  if (get_tmap_2d(&tq, Q, nq, dim, static_cast<uint64_t>(ldq) * kEB, NPAD, kBKe, kEB, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&td, D, n_docs, dim, static_cast<uint64_t>(ldd) * kEB, kSimBM, kBKe, kEB, TMAP_SW_128)) return -1;
  IM_CUDA_OK(cudaFuncSetAttribute(sim_topk_kernel<NPAD, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  sim_topk_kernel<NPAD, F8><<<grid, sim_threads<NPAD>(), smem_bytes, s>>>(tq, td, nq, n_docs, dim, stages, ktop, alive, out_scores,
                                                                   out_ids, thr_init, thr_stride, d_scale, q_scale);
  IM_LAUNCH_OK("sim_topk_kernel");
  return grid;
}

IM_API int im_sim_topk(const void* Q, const void* D, int nq, int n_docs, int dim, int ldq, int ldd, int ktop,
                       const uint8_t* alive, float* out_scores, int* out_ids, int max_ctas, const float* thr_init,
                       int thr_stride, void* stream) {
  using namespace im;
  if (nq < 1 || nq > 128) return set_error("im_sim_topk", "nq must be in [1,128]");
  if (dim % kSimBK != 0 || dim > 512) return set_error("im_sim_topk", "dim must be a multiple of 64 and <= 512");
  if (ktop < 0 || ktop > 32) return set_error("im_sim_topk", "ktop must be in [0,32] (0 = per-warp maxima only)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (nq <= 16) return launch_sim<16>(Q, D, nq, n_docs, dim, ldq, ldd, ktop, alive, out_scores, out_ids, max_ctas, thr_init, thr_stride, s);
  if (nq <= 32) return launch_sim<32>(Q, D, nq, n_docs, dim, ldq, ldd, ktop, alive, out_scores, out_ids, max_ctas, thr_init, thr_stride, s);
  if (nq <= 64) return launch_sim<64>(Q, D, nq, n_docs, dim, ldq, ldd, ktop, alive, out_scores, out_ids, max_ctas, thr_init, thr_stride, s);
  return launch_sim<128>(Q, D, nq, n_docs, dim, ldq, ldd, ktop, alive, out_scores, out_ids, max_ctas, thr_init, thr_stride, s);
}

// e4m3 shard + per-row scales (see the F8 note above the kernel).  Q8 / D8: row-major e4m3 bytes, ld in elements.
IM_API int im_sim_topk_f8(const void* Q8, const float* q_scale, const void* D8, const float* d_scale, int nq, int n_docs,
                          int dim, int ldq, int ldd, int ktop, const uint8_t* alive, float* out_scores, int* out_ids,
                          int max_ctas, const float* thr_init, int thr_stride, void* stream) {
  using namespace im;
  if (nq < 1 || nq > 128) return set_error("im_sim_topk_f8", "nq must be in [1,128]");
  if (dim % 128 != 0 || dim > 1024) return set_error("im_sim_topk_f8", "dim must be a multiple of 128 and <= 1024");
  if (ldq % 16 != 0 || ldd % 16 != 0) return set_error("im_sim_topk_f8", "row strides must be multiples of 16 bytes");
  if (ktop < 0 || ktop > 32) return set_error("im_sim_topk_f8", "ktop must be in [0,32]");
  if (q_scale == nullptr || d_scale == nullptr) return set_error("im_sim_topk_f8", "scales are required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define IM_SIM_F8(NP) launch_sim<NP, true>(Q8, D8, nq, n_docs, dim, ldq, ldd, ktop, alive, out_scores, out_ids, max_ctas, thr_init, thr_stride, s, d_scale, q_scale)
  if (nq <= 16) return IM_SIM_F8(16);
  if (nq <= 32) return IM_SIM_F8(32);
  if (nq <= 64) return IM_SIM_F8(64);
  return IM_SIM_F8(128);
#undef IM_SIM_F8
}

// Exact re-scoring of a candidate list against the bf16 rows (fp8 retrieval over-fetches, this restores the fp32-accurate
// order): one CTA per query, warp w scores candidate w (<= 32 candidates), warp 0 then ranks them (score desc, id asc)
// and writes the best k_out.  cand ids are LOCAL rows (int64, -1 = empty); out ids = row + id_offset.
namespace im {
__global__ void __launch_bounds__(1024)
rescore_topk_kernel(const __nv_bfloat16* __restrict__ Q, int ldq, const __nv_bfloat16* __restrict__ D, int ldd, int dim,
                    const int64_t* __restrict__ cand, int n_cand, int k_out, int64_t id_offset,
                    float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ float sc[32];
  __shared__ long long id[32];
  const int q = blockIdx.x;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (static_cast<int>(warp) < n_cand) {
    const int64_t row = cand[static_cast<size_t>(q) * n_cand + warp];
    float acc = 0.f;
    if (row >= 0) {
      const uint4* qv = reinterpret_cast<const uint4*>(Q + static_cast<size_t>(q) * ldq);
      const uint4* dv = reinterpret_cast<const uint4*>(D + static_cast<size_t>(row) * ldd);
      for (int c = lane; c < dim / 8; c += 32) {
        const uint4 a = qv[c], b = dv[c];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = unpack_bf16x2(aw[i]), y = unpack_bf16x2(bw[i]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) {
      sc[warp] = row >= 0 ? acc : -CUDART_INF_F;
      id[warp] = row;
    }
  } else if (warp < 32 && lane == 0) {
    sc[warp] = -CUDART_INF_F;
    id[warp] = -1;
  }
  __syncthreads();
  if (warp == 0) {
    const float s = sc[lane];
    const long long d = id[lane];
    int rank = 0;
    for (int j = 0; j < 32; ++j) {
      const float o = sc[j];
      const long long od = id[j];
      rank += (od >= 0 && (o > s || (o == s && (od < d || (od == d && j < static_cast<int>(lane)))))) ? 1 : 0;
    }
    if (d < 0) rank = 32;   // empties go last (and are rewritten below)
    if (rank < k_out) {
      out_scores[static_cast<size_t>(q) * k_out + rank] = s;
      out_ids[static_cast<size_t>(q) * k_out + rank] = d + id_offset;
    }
    const int live = __popc(__ballot_sync(0xffffffffu, d >= 0));
    for (int r = live + static_cast<int>(lane); r < k_out; r += 32) {
      out_scores[static_cast<size_t>(q) * k_out + r] = -CUDART_INF_F;
      out_ids[static_cast<size_t>(q) * k_out + r] = -1;
    }
  }
}
}  // namespace im

IM_API int im_rescore_topk(const void* Q, int ldq, const void* D, int ldd, int dim, const int64_t* cand, int nq, int n_cand,
                           int k_out, int64_t id_offset, float* out_scores, int64_t* out_ids, void* stream) {
  using namespace im;
  if (nq <= 0) return 0;
  if (n_cand < 1 || n_cand > 32 || k_out < 1 || k_out > 32) return set_error("im_rescore_topk", "at most 32 candidates");
  if (dim % 8 != 0 || ldq % 8 != 0 || ldd % 8 != 0) return set_error("im_rescore_topk", "rows must be 16-byte aligned");
  rescore_topk_kernel<<<nq, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)D, ldd, dim, cand, n_cand, k_out, id_offset, out_scores, out_ids);
  IM_LAUNCH_OK("rescore_topk_kernel");
  return 0;
}

IM_API int im_topk_merge(const float* cand_scores, const int64_t* cand_ids64, const int* cand_ids32, int P, int nq,
                         int k_in, int k_out, int64_t id_offset, float* out_scores, int64_t* out_ids,
                         float* const* peer_scores, int64_t* const* peer_ids, uint32_t* const* peer_flags, int world,
                         int rank, const uint32_t* wait_flags, uint32_t* state, void* stream, uint32_t* status,
                         unsigned wait_limit) {
  using namespace im;
  if (nq <= 0) return 0;
  if (k_out < 1 || k_out > 32) return set_error("im_topk_merge", "k_out must be in [1,32]");
  const size_t n = static_cast<size_t>(P) * k_in;
  const size_t smem = ((n * 4 + 7) & ~size_t(7)) + n * 8;
  if (smem > 200 * 1024) return set_error("im_topk_merge", "too many candidates per query");
  if (smem > 48 * 1024)
    IM_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<nq, kMergeThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      cand_scores, cand_ids64, cand_ids32, P, nq, k_in, k_out, id_offset, out_scores, out_ids, peer_scores, peer_ids,
      peer_flags, world, rank, wait_flags, state, status, wait_limit, debug_poison_slots());
  IM_LAUNCH_OK("topk_merge_kernel");
  return 0;
}
