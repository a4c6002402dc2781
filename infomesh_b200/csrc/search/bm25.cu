// K2 — BM25 posting-list intersection + scoring + top-k over an HBM-resident CSR inverted index.
//
// Replaces the SQLite FTS5 `MATCH ... ORDER BY bm25()` hot loop (reference infomesh/index/local_store.py:316-332,
// called from infomesh/search/query.py:124-133): implicit AND of the query terms, Okapi BM25 with FTS5's
// constants (k1 = 1.2, b = 0.75, idf = ln((N - n + 0.5) / (n + 0.5)) floored at 1e-6, score reported positive).
//
// Layout (per shard): post_off[V+1] (int64), post_doc[nnz] (int32, ascending inside a term), post_tf[nnz] (uint8),
// doc_norm[n_docs] = k1 * (1 - b + b * len / avg_len)  (fp32, pre-computed), idf[V] (fp32, global statistics).
//
// Dense terms (document frequency >= n_docs / 32, a few hundred under Zipf) additionally get a DIRECT MAP: one tf byte per
// document (`dense_tf[slot][doc]`, 0 = absent), so probing them during an intersection is a single load instead of a
// ~22-step binary search through a multi-million-entry list -- this is what keeps queries that contain common words
// (the reference's FTS5 handles them with its doclist index) from degenerating.
//
// One query per blockIdx.y.  The shortest posting list drives; each lane takes one driver posting, gallops
// (binary search) through the other lists, and survivors are pushed into a WARP-DISTRIBUTED sorted top-32
// (lane i holds entry i: insertion = ballot + shuffle, no shared memory).  Per-warp lists are merged by
// topk_merge_kernel (sim_topk.cu), which is also where the per-shard lists meet after the NVLink exchange.
#include <math_constants.h>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace im {

constexpr int kBm25MaxTerms = 16;
constexpr int kBm25Threads = 128;
constexpr float kBm25K1 = 1.2f;

struct WarpTop32 {
  float v;
  int id;
  __device__ __forceinline__ void init() {
    v = -CUDART_INF_F;
    id = -1;
  }
  __device__ __forceinline__ float kth() const { return __shfl_sync(0xffffffffu, v, 31); }
  // warp-uniform (s, d)
  __device__ __forceinline__ void insert(float s, int d, uint32_t lane) {
    const bool before = (v > s) || (v == s && id >= 0 && id < d);
    const int pos = __popc(__ballot_sync(0xffffffffu, before));
    if (pos >= 32) return;
    const float vu = __shfl_up_sync(0xffffffffu, v, 1);
    const int iu = __shfl_up_sync(0xffffffffu, id, 1);
    if (static_cast<int>(lane) == pos) {
      v = s;
      id = d;
    } else if (static_cast<int>(lane) > pos) {
      v = vu;
      id = iu;
    }
  }
};

// lower_bound of `key` in docs[lo, hi); returns index or -1 when absent
__device__ __forceinline__ long long find_doc(const int* __restrict__ docs, long long lo, long long hi, int key) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    const int v = __ldg(docs + mid);
    if (v < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBm25Threads)
bm25_and_topk_kernel(const long long* __restrict__ post_off, const int* __restrict__ post_doc,
                     const uint8_t* __restrict__ post_tf, const float* __restrict__ doc_norm,
                     const float* __restrict__ idf, const uint8_t* __restrict__ alive,
                     const int* __restrict__ q_terms, int max_terms, int vocab, float* __restrict__ out_scores,
                     int* __restrict__ out_ids, int* __restrict__ out_counts, const int* __restrict__ dense_slot,
                     const uint8_t* __restrict__ dense_tf, long long n_docs) {
  const int q = blockIdx.y;
  const uint32_t lane = threadIdx.x & 31;
  const int warp_in_q = blockIdx.x * (kBm25Threads / 32) + (threadIdx.x >> 5);
  const int warps_per_q = gridDim.x * (kBm25Threads / 32);

  // query terms: -1 padded; any out-of-vocabulary term makes the AND empty
  int terms[kBm25MaxTerms];
  const uint8_t* t_map[kBm25MaxTerms];   // direct tf map of a dense term, else null
  long long t_lo[kBm25MaxTerms], t_hi[kBm25MaxTerms];
  int nt = 0;
  bool empty = false;
  for (int i = 0; i < max_terms && i < kBm25MaxTerms; ++i) {
    const int t = q_terms[q * max_terms + i];
    if (t == -1) continue;
    if (t < 0 || t >= vocab) {
      empty = true;
      continue;
    }
    bool dup = false;
    for (int k = 0; k < nt; ++k) dup |= (terms[k] == t);
    if (dup) continue;
    terms[nt] = t;
    t_lo[nt] = post_off[t];
    t_hi[nt] = post_off[t + 1];
    const int slot = dense_slot != nullptr ? dense_slot[t] : -1;
    t_map[nt] = slot >= 0 ? dense_tf + static_cast<size_t>(slot) * n_docs : nullptr;
    if (t_hi[nt] == t_lo[nt]) empty = true;
    ++nt;
  }
  WarpTop32 top;
  top.init();
  int matched = 0;
  if (nt > 0 && !empty) {
    int drv = 0;
    for (int i = 1; i < nt; ++i)
      if (t_hi[i] - t_lo[i] < t_hi[drv] - t_lo[drv]) drv = i;
    const long long d_lo = t_lo[drv], d_len = t_hi[drv] - t_lo[drv];
    const float idf_drv = idf[terms[drv]];
    for (long long base = static_cast<long long>(warp_in_q) * 32; base < d_len;
         base += static_cast<long long>(warps_per_q) * 32) {
      const long long idx = base + lane;
      bool hit = idx < d_len;
      int doc = -1;
      float score = 0.f;
      if (hit) {
        doc = __ldg(post_doc + d_lo + idx);
        if (alive != nullptr && alive[doc] == 0) hit = false;
      }
      float norm = 0.f;
      if (hit) {
        norm = __ldg(doc_norm + doc);
        const float tf = static_cast<float>(__ldg(post_tf + d_lo + idx));
        score = idf_drv * tf * (kBm25K1 + 1.0f) / (tf + norm);
      }
      for (int i = 0; i < nt; ++i) {
        if (i == drv) continue;
        if (hit) {
          float tf = 0.f;
          if (t_map[i] != nullptr) {
            tf = static_cast<float>(__ldg(t_map[i] + doc));          // dense term: one byte, 0 = absent
          } else {
            const long long pos = find_doc(post_doc, t_lo[i], t_hi[i], doc);
            if (pos < t_hi[i] && __ldg(post_doc + pos) == doc) tf = static_cast<float>(__ldg(post_tf + pos));
          }
          if (tf > 0.f) score += idf[terms[i]] * tf * (kBm25K1 + 1.0f) / (tf + norm);
          else hit = false;
        }
      }
      matched += hit ? 1 : 0;
      const float kth = top.kth();  // all lanes take part in the shuffle BEFORE the (short-circuiting) predicate
      uint32_t cand = __ballot_sync(0xffffffffu, hit && score > kth);
      while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        const float s = __shfl_sync(0xffffffffu, score, src);
        const int d = __shfl_sync(0xffffffffu, doc, src);
        top.insert(s, d, lane);
      }
    }
  }
  const size_t o = (static_cast<size_t>(warp_in_q) * gridDim.y + q) * 32 + lane;  // [P = warps][nq][32]
  out_scores[o] = top.v;
  out_ids[o] = top.id;
  if (out_counts != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) matched += __shfl_xor_sync(0xffffffffu, matched, off);
    if (lane == 0 && matched) atomicAdd(out_counts + q, matched);
  }
}

// ------------------------------------------------------------------------------------------------
// K11 — passage selection: for (query, document) pairs pick the passage maximising
//   coverage + 0.1 * density,   coverage = |Q ∩ P| / |Q|,  density = (#tokens of P that are query terms) / |P|
// (reference infomesh/search/passage.py:143-180, select_best_passage :183-227).  Documents are pre-tokenised
// term-id streams resident in HBM with passage boundaries; one CTA per pair, one warp per passage.
// ------------------------------------------------------------------------------------------------
constexpr int kPassMaxQ = 32;

__global__ void __launch_bounds__(128)
passage_score_kernel(const int* __restrict__ tok, const long long* __restrict__ pass_off,
                     const long long* __restrict__ doc_pass_off, const int* __restrict__ pair_doc,
                     const int* __restrict__ pair_query, const int* __restrict__ q_terms, int max_q,
                     float* __restrict__ out_score, int* __restrict__ out_passage) {
  __shared__ float best_s[4];
  __shared__ int best_p[4];
  const int pair = blockIdx.x;
  const int doc = pair_doc[pair], qi = pair_query[pair];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lane j holds query term j (unique, -1 padded)
  const int my_term = (static_cast<int>(lane) < max_q && lane < kPassMaxQ) ? q_terms[qi * max_q + lane] : -1;
  const uint32_t q_mask = __ballot_sync(0xffffffffu, my_term >= 0);
  const int nq_terms = __popc(q_mask);
  float bs = -1.f;
  int bp = -1;
  if (doc >= 0 && nq_terms > 0) {
    const long long p0 = doc_pass_off[doc], p1 = doc_pass_off[doc + 1];
    for (long long p = p0 + warp; p < p1; p += 4) {
      const long long a = pass_off[p], b = pass_off[p + 1];
      uint32_t present = 0;
      int hits = 0;
      for (long long base = a; base < b; base += 32) {  // warp-uniform trip count (shuffles inside)
        const long long i = base + lane;
        const int t = i < b ? __ldg(tok + i) : -2;
        uint32_t m = 0;
        for (int j = 0; j < 32; ++j) {
          if (!((q_mask >> j) & 1u)) continue;
          const int qt = __shfl_sync(0xffffffffu, my_term, j);  // broadcast query term j
          m |= (qt == t) ? (1u << j) : 0u;
        }
        present |= m;
        hits += m ? 1 : 0;
      }
      present = __reduce_or_sync(0xffffffffu, present);
      hits = __reduce_add_sync(0xffffffffu, hits);
      const float len = static_cast<float>(b - a);
      const float s = len > 0.f ? static_cast<float>(__popc(present)) / nq_terms + 0.1f * hits / len : 0.f;
      const int pl = static_cast<int>(p - p0);
      if (s > bs || (s == bs && pl < bp)) {
        bs = s;
        bp = pl;
      }
    }
  }
  if (lane == 0) {
    best_s[warp] = bs;
    best_p[warp] = bp;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w)
      if (best_p[w] >= 0 && (best_s[w] > bs || (best_s[w] == bs && best_p[w] < bp) || bp < 0)) {
        bs = best_s[w];
        bp = best_p[w];
      }
    out_score[pair] = bp >= 0 ? bs : 0.f;
    out_passage[pair] = bp;
  }
}

}  // namespace im

// out lists: [P][nq][32] with P = blocks_per_query * 4 warps; returns P or <0
static int bm25_topk_impl(const long long* post_off, const int* post_doc, const uint8_t* post_tf, const float* doc_norm,
                          const float* idf, const uint8_t* alive, const int* q_terms, int nq, int max_terms, int vocab,
                          int blocks_per_query, float* out_scores, int* out_ids, int* out_counts, void* stream,
                          const int* dense_slot, const uint8_t* dense_tf, long long n_docs) {
  using namespace im;
  if (nq <= 0) return 0;
  if (max_terms > kBm25MaxTerms) return set_error("im_bm25_topk", "at most 16 query terms");
  if (blocks_per_query < 1) blocks_per_query = 1;
  dim3 grid(blocks_per_query, nq);
  bm25_and_topk_kernel<<<grid, kBm25Threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      post_off, post_doc, post_tf, doc_norm, idf, alive, q_terms, max_terms, vocab, out_scores, out_ids, out_counts, dense_slot,
      dense_tf, n_docs);
  IM_LAUNCH_OK("bm25_and_topk_kernel");
  return blocks_per_query * (kBm25Threads / 32);
}
IM_API int im_bm25_topk(const long long* post_off, const int* post_doc, const uint8_t* post_tf, const float* doc_norm,
                        const float* idf, const uint8_t* alive, const int* q_terms, int nq, int max_terms, int vocab,
                        int blocks_per_query, float* out_scores, int* out_ids, int* out_counts, void* stream) {
  return bm25_topk_impl(post_off, post_doc, post_tf, doc_norm, idf, alive, q_terms, nq, max_terms, vocab, blocks_per_query,
                        out_scores, out_ids, out_counts, stream, nullptr, nullptr, 0);
}
// Same with direct tf maps for dense terms: dense_slot[vocab] (-1 = sparse), dense_tf[n_slots][n_docs].
IM_API int im_bm25_topk_dense(const long long* post_off, const int* post_doc, const uint8_t* post_tf, const float* doc_norm,
                              const float* idf, const uint8_t* alive, const int* q_terms, int nq, int max_terms, int vocab,
                              int blocks_per_query, float* out_scores, int* out_ids, int* out_counts, const int* dense_slot,
                              const uint8_t* dense_tf, long long n_docs, void* stream) {
  return bm25_topk_impl(post_off, post_doc, post_tf, doc_norm, idf, alive, q_terms, nq, max_terms, vocab, blocks_per_query,
                        out_scores, out_ids, out_counts, stream, dense_slot, dense_tf, n_docs);
}

IM_API int im_passage_score(const int* tok, const long long* pass_off, const long long* doc_pass_off,
                            const int* pair_doc, const int* pair_query, const int* q_terms, int n_pairs, int max_q,
                            float* out_score, int* out_passage, void* stream) {
  using namespace im;
  if (n_pairs <= 0) return 0;
  if (max_q > kPassMaxQ) return set_error("im_passage_score", "at most 32 query terms");
  passage_score_kernel<<<n_pairs, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tok, pass_off, doc_pass_off, pair_doc, pair_query, q_terms, max_q, out_score, out_passage);
  IM_LAUNCH_OK("passage_score_kernel");
  return 0;
}
