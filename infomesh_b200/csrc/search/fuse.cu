// K4/K12 tail of the hybrid pipeline (all tiny, latency-bound kernels that keep the query on the GPU):
//   rrf_fuse      Reciprocal Rank Fusion of the BM25 and dense candidate lists (reference
//                 infomesh/search/merge.py:37-133: RRF(d) = sum_s w_s / (60 + rank_s(d)), keyed by document),
//                 one warp per query, output sorted by (rrf desc, id asc)
//   build_pairs   assemble cross-encoder inputs  <s> query </s></s> passage </s>  for (query, candidate) pairs,
//                 reading passage tokens from the owning shard (local or peer-mapped HBM over NVLink)
//   rerank_select per query: order candidates by reranker logit (desc) and emit the final top-k
//                 (reference infomesh/search/reranker.py:86-163 returns a permutation; failures keep input order —
//                 here a NaN logit sorts last so the fused order is preserved for it)
#include <math_constants.h>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace im {

constexpr int kFuseMax = 32;  // candidates per source list

__global__ void __launch_bounds__(128)
rrf_fuse_kernel(const int64_t* __restrict__ ids_a, const int64_t* __restrict__ ids_b, int ka, int kb, int nq,
                float rrf_k, float wa, float wb, int k_out, float* __restrict__ out_scores,
                int64_t* __restrict__ out_ids, int* __restrict__ out_src) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (q >= nq) return;
  // lane i holds entry i of each list (rank = i + 1)
  const int64_t ia = (static_cast<int>(lane) < ka) ? ids_a[static_cast<size_t>(q) * ka + lane] : -1;
  const int64_t ib = (static_cast<int>(lane) < kb) ? ids_b[static_cast<size_t>(q) * kb + lane] : -1;
  float sa = ia >= 0 ? wa / (rrf_k + lane + 1.0f) : 0.f;  // contribution of list a for entry `lane`
  float sb = ib >= 0 ? wb / (rrf_k + lane + 1.0f) : 0.f;
  int src_a = ia >= 0 ? 1 : 0;  // bit0 = list a (fts), bit1 = list b (vector)
  bool b_dup = false;
  // match list b entries against list a: a gets b's contribution, the b entry is retired
  for (int j = 0; j < kb; ++j) {
    const int64_t idj = __shfl_sync(0xffffffffu, ib, j);
    const float sj = __shfl_sync(0xffffffffu, sb, j);
    const bool m = (idj >= 0) && (ia == idj);
    if (m) {
      sa += sj;
      src_a |= 2;
    }
    if (__any_sync(0xffffffffu, m) && static_cast<int>(lane) == j) b_dup = true;
  }
  // 64 candidates: (sa, ia) and (sb, ib) unless dup / empty
  const float ca = ia >= 0 ? sa : -CUDART_INF_F;
  const float cb = (ib >= 0 && !b_dup) ? sb : -CUDART_INF_F;
  // rank by counting candidates that sort before mine (score desc, id asc)
  int ra = 0, rb = 0;
  for (int j = 0; j < 32; ++j) {
    const float oa = __shfl_sync(0xffffffffu, ca, j), ob = __shfl_sync(0xffffffffu, cb, j);
    const int64_t oia = __shfl_sync(0xffffffffu, ia, j), oib = __shfl_sync(0xffffffffu, ib, j);
    ra += (oa > ca || (oa == ca && oia < ia)) ? 1 : 0;
    ra += (ob > ca || (ob == ca && oib < ia)) ? 1 : 0;
    rb += (oa > cb || (oa == cb && oia < ib)) ? 1 : 0;
    rb += (ob > cb || (ob == cb && oib < ib)) ? 1 : 0;
  }
  if (ca > -CUDART_INF_F && ra < k_out) {
    out_scores[static_cast<size_t>(q) * k_out + ra] = ca;
    out_ids[static_cast<size_t>(q) * k_out + ra] = ia;
    if (out_src) out_src[static_cast<size_t>(q) * k_out + ra] = src_a;
  }
  if (cb > -CUDART_INF_F && rb < k_out) {
    out_scores[static_cast<size_t>(q) * k_out + rb] = cb;
    out_ids[static_cast<size_t>(q) * k_out + rb] = ib;
    if (out_src) out_src[static_cast<size_t>(q) * k_out + rb] = 2;
  }
  // pad the tail when fewer than k_out distinct candidates exist
  const int n_valid = __popc(__ballot_sync(0xffffffffu, ca > -CUDART_INF_F)) +
                      __popc(__ballot_sync(0xffffffffu, cb > -CUDART_INF_F));
  for (int r = n_valid + lane; r < k_out; r += 32) {
    out_scores[static_cast<size_t>(q) * k_out + r] = -CUDART_INF_F;
    out_ids[static_cast<size_t>(q) * k_out + r] = -1;
    if (out_src) out_src[static_cast<size_t>(q) * k_out + r] = 0;
  }
}

// One CTA per (query, candidate) pair.  Documents are sharded contiguously: owner = id / docs_per_shard.
__global__ void __launch_bounds__(128)
build_pairs_kernel(const int* __restrict__ q_tok, const int* __restrict__ q_len, int max_q_len,
                   const int64_t* __restrict__ cand_ids, int n_cand, const int* const* __restrict__ shard_tok,
                   const int* const* __restrict__ shard_len, int64_t docs_per_shard, int passage_len, int seq_len,
                   int bos, int eos, int pad, int* __restrict__ out_ids, int* __restrict__ out_lens) {
  const int pair = blockIdx.x;
  const int q = pair / n_cand;
  const int64_t doc = cand_ids[pair];
  int* out = out_ids + static_cast<size_t>(pair) * seq_len;
  const int ql = min(q_len[q], min(max_q_len, seq_len / 2 - 2));
  int pl = 0;
  const int* ptok = nullptr;
  if (doc >= 0) {
    const int64_t owner = doc / docs_per_shard, local = doc - owner * docs_per_shard;
    ptok = shard_tok[owner] + local * passage_len;
    pl = min(shard_len[owner][local], passage_len);
  }
  pl = min(pl, seq_len - ql - 4);
  if (pl < 0) pl = 0;
  const int total = ql + pl + 4;  // <s> q </s> </s> p </s>
  for (int i = threadIdx.x; i < seq_len; i += blockDim.x) {
    int t;
    if (i == 0) t = bos;
    else if (i <= ql) t = q_tok[static_cast<size_t>(q) * max_q_len + (i - 1)];
    else if (i == ql + 1 || i == ql + 2) t = eos;
    else if (i < ql + 3 + pl) t = ptok[i - (ql + 3)];
    else if (i == ql + 3 + pl) t = eos;
    else t = pad;
    out[i] = t;
  }
  if (threadIdx.x == 0) out_lens[pair] = doc >= 0 ? total : 1;
}

// One warp per query: candidates (<= 32) ordered by logit desc (NaN / missing last, ties keep fused order).
__global__ void __launch_bounds__(128)
rerank_select_kernel(const float* __restrict__ logits, const int64_t* __restrict__ cand_ids, int n_cand, int nq,
                     int k_out, float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (q >= nq) return;
  const bool has = static_cast<int>(lane) < n_cand;
  const int64_t id = has ? cand_ids[static_cast<size_t>(q) * n_cand + lane] : -1;
  float s = has ? logits[static_cast<size_t>(q) * n_cand + lane] : -CUDART_INF_F;
  if (id < 0 || !(s == s)) s = -CUDART_INF_F;
  int rank = 0;
  for (int j = 0; j < 32; ++j) {
    const float o = __shfl_sync(0xffffffffu, s, j);
    rank += (o > s || (o == s && j < static_cast<int>(lane))) ? 1 : 0;
  }
  if (has && rank < k_out) {
    out_scores[static_cast<size_t>(q) * k_out + rank] = s;
    out_ids[static_cast<size_t>(q) * k_out + rank] = id;
  }
  for (int r = n_cand + lane; r < k_out; r += 32) {
    out_scores[static_cast<size_t>(q) * k_out + r] = -CUDART_INF_F;
    out_ids[static_cast<size_t>(q) * k_out + r] = -1;
  }
}


// K12 — six-signal rank fuse (reference infomesh/index/ranking.py:104-148,171-238): one warp per query, one candidate
// per lane (<= 32).  score = w0 * bm25 / (bm25 + max_bm25_of_query) + w1 * max(0.05, 2^(-age / half_life)) + w2 * trust
//                          + w3 * authority + w4 * title_match + w5 * url_path; candidates come out sorted by score
// (ties keep retrieval order).  Per-document signals (crawl time, trust, authority) are gathered by document row;
// the query-dependent bonuses are optional per-pair arrays.
__global__ void __launch_bounds__(128)
rank_fuse_kernel(const float* __restrict__ bm25, const int64_t* __restrict__ rows, int n_cand, int nq, int64_t row_base,
                 const float* __restrict__ crawled_at, const float* __restrict__ trust, const float* __restrict__ authority,
                 const float* __restrict__ title_match, const float* __restrict__ url_path, float now, float half_life,
                 float min_fresh, float default_trust, const float* __restrict__ w, int k_out, float* __restrict__ out_scores,
                 int64_t* __restrict__ out_rows, float* __restrict__ out_signals) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (q >= nq) return;
  const bool has = static_cast<int>(lane) < n_cand;
  const size_t at = static_cast<size_t>(q) * n_cand + lane;
  const int64_t row = has ? rows[at] : -1;
  const bool live = has && row >= 0;
  const float raw = live ? bm25[at] : 0.f;
  float top = raw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, o));
  if (!(top > 0.f)) top = 1.f;
  const int64_t local = row - row_base;
  const float sb = raw > 0.f ? raw / (raw + top) : 0.f;
  const float age = live && crawled_at != nullptr ? fmaxf(0.f, now - crawled_at[local]) : 0.f;
  const float sf = fmaxf(min_fresh, exp2f(-age / half_life));
  const float st = live && trust != nullptr ? trust[local] : default_trust;
  const float sa = live && authority != nullptr ? authority[local] : 0.f;
  const float stt = live && title_match != nullptr ? title_match[at] : 0.f;
  const float su = live && url_path != nullptr ? url_path[at] : 0.f;
  float s = w[0] * sb + w[1] * sf + w[2] * st + w[3] * sa + w[4] * stt + w[5] * su;
  if (!live) s = -CUDART_INF_F;
  int rank = 0;
  for (int j = 0; j < 32; ++j) {
    const float o = __shfl_sync(0xffffffffu, s, j);
    rank += (o > s || (o == s && j < static_cast<int>(lane))) ? 1 : 0;
  }
  if (has && rank < k_out) {
    out_scores[static_cast<size_t>(q) * k_out + rank] = s;
    out_rows[static_cast<size_t>(q) * k_out + rank] = live ? row : -1;
    if (out_signals != nullptr) {
      float* o = out_signals + (static_cast<size_t>(q) * k_out + rank) * 6;
      o[0] = sb; o[1] = sf; o[2] = st; o[3] = sa; o[4] = stt; o[5] = su;
    }
  }
  for (int r = n_cand + lane; r < k_out; r += 32) {
    out_scores[static_cast<size_t>(q) * k_out + r] = -CUDART_INF_F;
    out_rows[static_cast<size_t>(q) * k_out + r] = -1;
  }
}

}  // namespace im

IM_API int im_rrf_fuse(const int64_t* ids_a, const int64_t* ids_b, int ka, int kb, int nq, float rrf_k, float wa,
                       float wb, int k_out, float* out_scores, int64_t* out_ids, int* out_src, void* stream) {
  using namespace im;
  if (nq <= 0) return 0;
  if (ka > kFuseMax || kb > kFuseMax || k_out > 2 * kFuseMax) return set_error("im_rrf_fuse", "lists limited to 32 entries");
  rrf_fuse_kernel<<<(nq + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids_a, ids_b, ka, kb, nq, rrf_k, wa, wb, k_out, out_scores, out_ids, out_src);
  IM_LAUNCH_OK("rrf_fuse_kernel");
  return 0;
}

IM_API int im_build_pairs(const int* q_tok, const int* q_len, int max_q_len, const int64_t* cand_ids, int nq, int n_cand,
                          const int* const* shard_tok, const int* const* shard_len, long long docs_per_shard,
                          int passage_len, int seq_len, int bos, int eos, int pad, int* out_ids, int* out_lens,
                          void* stream) {
  using namespace im;
  if (nq * n_cand <= 0) return 0;
  build_pairs_kernel<<<nq * n_cand, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q_tok, q_len, max_q_len, cand_ids, n_cand, shard_tok, shard_len, docs_per_shard, passage_len, seq_len, bos, eos,
      pad, out_ids, out_lens);
  IM_LAUNCH_OK("build_pairs_kernel");
  return 0;
}

IM_API int im_rerank_select(const float* logits, const int64_t* cand_ids, int n_cand, int nq, int k_out,
                            float* out_scores, int64_t* out_ids, void* stream) {
  using namespace im;
  if (nq <= 0) return 0;
  if (n_cand > 32) return set_error("im_rerank_select", "at most 32 candidates per query");
  rerank_select_kernel<<<(nq + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, cand_ids, n_cand, nq,
                                                                                        k_out, out_scores, out_ids);
  IM_LAUNCH_OK("rerank_select_kernel");
  return 0;
}

IM_API int im_rank_fuse(const float* bm25, const int64_t* rows, int n_cand, int nq, long long row_base, const float* crawled_at,
                        const float* trust, const float* authority, const float* title_match, const float* url_path, float now,
                        float half_life, float min_fresh, float default_trust, const float* weights, int k_out, float* out_scores,
                        int64_t* out_rows, float* out_signals, void* stream) {
  using namespace im;
  if (nq <= 0) return 0;
  if (n_cand > 32 || k_out > 32) return set_error("im_rank_fuse", "at most 32 candidates per query");
  rank_fuse_kernel<<<(nq + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      bm25, rows, n_cand, nq, row_base, crawled_at, trust, authority, title_match, url_path, now, half_life, min_fresh, default_trust,
      weights, k_out, out_scores, out_rows, out_signals);
  IM_LAUNCH_OK("rank_fuse_kernel");
  return 0;
}
