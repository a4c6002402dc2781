// K7 — flash-style attention forward on tcgen05 (encoder bidirectional, causal, cross, T5 relative bias).
//
// One CTA = (128 query rows, one head, one sequence).  Per 128-key chunk:
//   MMA1  S[128q x 128k] = Q · K^T          tcgen05.mma, operands TMA-loaded K-major, S lives in TMEM
//   softmax (thread t owns query row t = TMEM lane t: no cross-lane reductions), two passes over TMEM:
//         row max, then p = exp2(s - m) packed to bf16 into a 128B-swizzled K-major smem tile
//   MMA2  O_c[128q x HD] = P · V_c          V tile consumed MN-major straight from its natural [keys, HD] layout
//   merge O_c into fp32 register accumulators with the usual online-softmax rescale
// K_{c+1} is prefetched as soon as MMA1(c) retires and V_{c+1} as soon as MMA2(c) retires, so a single K and V
// buffer (80 KB smem per CTA, 256 TMEM columns) lets two CTAs share an SM and overlap each other's
// softmax with tensor-core work.  This is the "summarizer's attention" hot path of the north-star; the reference
// has no attention code of its own (it calls sentence-transformers / an HTTP LLM: infomesh/index/vector_store.py:124,
// infomesh/summarizer/engine.py:126-141).
#include <math_constants.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "../common/tmap_cache.h"

namespace im {

struct AttnParams {
  __nv_bfloat16* out;     // [B*Sq, ldo], head h at column h*HD
  int ldo;
  int Sq, Sk;             // padded per-sequence lengths (rows per batch element in the q / kv buffers)
  const int* kv_lens;     // [B] valid keys per sequence (null => Sk)
  int causal;             // key j visible to query i iff j <= i + causal_offset
  int causal_offset;
  float scale_log2;       // softmax scale * log2(e)
  const float* rel_bias;  // [nH, Sq + Sk - 1] additive bias indexed by (j - i) + (Sq - 1), or null (already * log2e)
  int alias_p;            // single-chunk, HD == 64: P overwrites the (dead) Q+K tiles -> 48 KB smem, 4 CTAs / SM
  const int* cu_seqlens;  // [B+1] packed (unpadded) batch: sequence b owns rows [cu[b], cu[b+1]) of q / k / v / out
  // MXFP8 output (single-chunk kernel, head_dim 64): the normalised O row is quantised in the epilogue -- each thread
  // owns exactly one 32-column block -- and written as e4m3 bytes + ue8m0 scales in the SFA chunk layout of the
  // out-projection GEMM (csrc/gemm/gemm_mxf8.cu), so no bf16 context tensor and no quantiser kernel exist on that path.
  // Context parallelism (multi-chunk kernel): the key / value sequence is sharded over `cp_world` ranks (cp_sk_local keys
  // per rank and sequence, a multiple of 128); this rank's queries sit at global positions [cp_q_pos0, +Sq).  The kernel
  // pulls every rank's K / V tiles itself -- TMA loads through tensor maps over the peers' symmetric-heap buffers -- in
  // an all-pairs schedule that starts with its own shard (NVSwitch makes every peer one hop, so no ring is needed).
  int cp_world, cp_rank, cp_sk_local, cp_q_pos0;
  uint8_t* out_q;         // [rows, ld_outq] e4m3 bytes or null
  uint8_t* out_sf;        // [ceil(rows/128)][n_kb][512] scale chunks
  int ld_outq, n_kb;
};

struct CpMaps {           // K / V tensor maps of every context-parallel rank (unused slots repeat slot 0)
  CUtensorMap k[8];
  CUtensorMap v[8];
};

constexpr int kAttnThreads = 128;
constexpr int kAttnBQ = 128;
constexpr int kAttnBKV = 128;
constexpr float kNegBig = -1.0e30f;

template <int HD>
struct AttnCfg {
  static constexpr int kRowBytes = HD * 2;                 // 128 (SW128) or 64 (SW64)
  static constexpr int kTileBytes = kAttnBQ * kRowBytes;   // Q / K / V tile
  static constexpr int kPBytes = kAttnBQ * kAttnBKV * 2;   // 32 KB, two [128 x 64] SW128 k-blocks
  static constexpr int kTmemCols = 128;                    // S: 128 columns; O_c reuses S[0, HD) once P left TMEM
  static constexpr int kGroupBytes = 8 * kRowBytes;        // 8-row swizzle group (SBO)
};

template <int HD>
__device__ __forceinline__ uint64_t desc_k_major(uint32_t saddr) {
  return HD == 64 ? umma_desc_k_sw128(saddr) : umma_desc_k_sw64(saddr);
}
template <int HD>
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t saddr) {
  return HD == 64 ? umma_desc_mn_sw128(saddr, 16) : umma_desc_mn_sw64(saddr, 16);
}

template <int HD, bool CP>
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, const __grid_constant__ CpMaps cpm) {
  using Cfg = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kTileBytes;
  uint8_t* sV = sK + Cfg::kTileBytes;
  // offsets stay multiples of 1024 (tile bytes are 8 KB or 16 KB)
  uint8_t* sP = p.alias_p ? sQ : sV + Cfg::kTileBytes;
  uint64_t* q_bar = reinterpret_cast<uint64_t*>(sV + Cfg::kTileBytes + (p.alias_p ? 0 : Cfg::kPBytes));
  uint64_t* k_bar = q_bar + 1;
  uint64_t* v_bar = q_bar + 2;
  uint64_t* s_bar = q_bar + 3;
  uint64_t* o_bar = q_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 5);
  float* s_bias = reinterpret_cast<float*>(q_bar + 6);

  const int tid = threadIdx.x;
  const uint32_t warp = warp_id();
  const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  pdl_trigger();
  pdl_wait();  // (kv_lens / rel_bias are read right below: no prologue to overlap in this kernel, only launch latency)
  const int q_row0 = b * p.Sq + qb * kAttnBQ;  // first query row of this tile in the q buffer
  const int kv_row0 = b * p.Sk;
  const int q_loc = qb * kAttnBQ + tid;         // query row inside this rank's [Sq] shard (store guard)
  const int q_idx = q_loc + (CP ? p.cp_q_pos0 : 0);   // position of this thread's query within the whole sequence
  const int cpr = CP ? p.cp_sk_local / kAttnBKV : 0;   // key chunks per context-parallel rank

  int kv_limit = p.kv_lens ? min(p.kv_lens[b], p.Sk) : p.Sk;
  if (p.causal) kv_limit = min(kv_limit, qb * kAttnBQ + kAttnBQ + p.causal_offset);
  const int num_chunks = max(0, (kv_limit + kAttnBKV - 1) / kAttnBKV);
  // iteration c works on global key chunk cg(c): context parallelism rotates the order so every rank starts on its own
  // shard and the ranks never all pull from the same peer at once
  const int rot = (CP && num_chunks > 0) ? (p.cp_rank * cpr) % num_chunks : 0;
  auto chunk_of = [&](int c) { return CP ? (c + rot) % num_chunks : c; };
  auto load_k = [&](int c) {
    const int cg = chunk_of(c);
    if constexpr (CP) tma_load_2d(sK, &cpm.k[cg / cpr], k_bar, head * HD, b * p.cp_sk_local + (cg % cpr) * kAttnBKV);
    else tma_load_2d(sK, &tmap_k, k_bar, head * HD, kv_row0 + cg * kAttnBKV);
  };
  auto load_v = [&](int c) {
    const int cg = chunk_of(c);
    if constexpr (CP) tma_load_2d(sV, &cpm.v[cg / cpr], v_bar, head * HD, b * p.cp_sk_local + (cg % cpr) * kAttnBKV);
    else tma_load_2d(sV, &tmap_v, v_bar, head * HD, kv_row0 + cg * kAttnBKV);
  };

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_bar, 1);
    mbar_init(k_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (p.rel_bias != nullptr) {
    const int nb = (CP ? p.Sk : p.Sq) + p.Sk - 1;
    for (int i = tid; i < nb; i += kAttnThreads) s_bias[i] = p.rel_bias[static_cast<size_t>(head) * nb + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // 128 columns
  const uint32_t tmem_o = tmem_base;        // HD columns, aliasing S (S is dead once P sits in smem)
  const uint32_t lane_base = (warp * 32u) << 16;

  if (tid == 0 && num_chunks > 0) {
    mbar_expect_tx(q_bar, Cfg::kTileBytes);
    tma_load_2d(sQ, &tmap_q, q_bar, head * HD, q_row0);
    mbar_expect_tx(k_bar, Cfg::kTileBytes);
    load_k(0);
    mbar_expect_tx(v_bar, Cfg::kTileBytes);
    load_v(0);
  }

  float o_acc[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
  float m_run = kNegBig, l_run = 0.f;
  const int kv_len = p.kv_lens ? min(p.kv_lens[b], p.Sk) : p.Sk;

  for (int c = 0; c < num_chunks; ++c) {
    const uint32_t ph = c & 1;
    // ---------------- MMA1: S = Q K^T ----------------
    if (tid == 0) {
      if (c == 0) mbar_wait(q_bar, 0);
      mbar_wait(k_bar, ph);
      tc_fence_after();
      constexpr uint32_t idesc1 = umma_idesc_f16(kAttnBQ, kAttnBKV);
      const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem_s, desc_k_major<HD>(a0 + k * 32), desc_k_major<HD>(b0 + k * 32), idesc1, k != 0 ? 1u : 0u);
      umma_commit(s_bar);
    }
    __syncwarp();
    mbar_wait(s_bar, ph);
    tc_fence_after();
    if (tid == 0 && c + 1 < num_chunks) {  // K buffer is free: prefetch the next chunk under the softmax
      mbar_expect_tx(k_bar, Cfg::kTileBytes);
      load_k(c + 1);
    }
    __syncwarp();

    // ---------------- softmax over this thread's row ----------------
    const int key0 = chunk_of(c) * kAttnBKV;
    int vis_end = kv_len;  // keys [0, vis_end) visible
    if (p.causal) vis_end = min(vis_end, q_idx + p.causal_offset + 1);
    // bias table index = (j - i) + (S_q_total - 1); with context parallelism the table spans the WHOLE sequence (Sk == S)
    const int sq_tot = CP ? p.Sk : p.Sq;
    const float* bias_row = p.rel_bias ? s_bias + (sq_tot - 1 - min(q_idx, sq_tot - 1)) : nullptr;  // index by key j
    float m_c = kNegBig;
#pragma unroll 1
    for (int cc = 0; cc < kAttnBKV; cc += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_s + lane_base + cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = key0 + cc + i;
        float s = __uint_as_float(v[i]) * p.scale_log2;
        if (bias_row) s += bias_row[min(j, p.Sk - 1)];
        s = (j < vis_end) ? s : kNegBig;
        m_c = fmaxf(m_c, s);
      }
    }
    float l_c = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < kAttnBKV; cc += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_s + lane_base + cc, v);
      tmem_ld_wait();
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float e[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = key0 + cc + i + u;
          float s = __uint_as_float(v[i + u]) * p.scale_log2;
          if (bias_row) s += bias_row[min(j, p.Sk - 1)];
          e[u] = (j < vis_end) ? exp2f(s - m_c) : 0.f;
        }
        // accumulate the bf16-rounded value so numerator and denominator agree
        const uint32_t pk = pack_bf16x2(e[0], e[1]);
        const float2 r = unpack_bf16x2(pk);
        l_c += r.x + r.y;
        packed[i >> 1] = pk;
      }
      const int kb = cc >> 6;               // which [128 x 64] k-block
      const int ch0 = (cc & 63) >> 3;       // first 16-byte chunk inside the 128-byte row
      uint8_t* prow = sP + kb * (kAttnBQ * 128) + tid * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int ch = (ch0 + q4) ^ (tid & 7);
        *reinterpret_cast<uint4*>(prow + (ch << 4)) =
            make_uint4(packed[q4 * 4 + 0], packed[q4 * 4 + 1], packed[q4 * 4 + 2], packed[q4 * 4 + 3]);
      }
    }
    fence_proxy_async_smem();  // st.shared P -> visible to tcgen05.mma (async proxy)
    tc_fence_before();
    __syncthreads();

    // ---------------- MMA2: O_c = P V_c ----------------
    if (tid == 0) {
      mbar_wait(v_bar, ph);
      tc_fence_after();
      constexpr uint32_t idesc2 = umma_idesc_f16(kAttnBQ, HD, 1, false, true);
      const uint32_t a0 = smem_u32(sP), b0 = smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < kAttnBKV / 16; ++ks) {
        const uint32_t a_addr = a0 + (ks >> 2) * (kAttnBQ * 128) + (ks & 3) * 32;
        const uint32_t b_addr = b0 + ks * 2 * Cfg::kGroupBytes;  // 16 keys = two 8-row swizzle groups
        umma_bf16(tmem_o, umma_desc_k_sw128(a_addr), desc_mn_major<HD>(b_addr), idesc2, ks != 0 ? 1u : 0u);
      }
      umma_commit(o_bar);
    }
    __syncwarp();
    mbar_wait(o_bar, ph);
    tc_fence_after();
    if (tid == 0 && c + 1 < num_chunks) {  // V buffer is free
      mbar_expect_tx(v_bar, Cfg::kTileBytes);
      load_v(c + 1);
    }
    __syncwarp();

    // ---------------- merge into the running accumulators ----------------
    const float m_new = fmaxf(m_run, m_c);
    const float alpha = exp2f(m_run - m_new);
    const float beta = exp2f(m_c - m_new);
    l_run = l_run * alpha + l_c * beta;
    m_run = m_new;
#pragma unroll
    for (int cc = 0; cc < HD; cc += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_o + lane_base + cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[cc + i] = o_acc[cc + i] * alpha + __uint_as_float(v[i]) * beta;
    }
    tc_fence_before();
    __syncthreads();  // everyone is done with S / O / P before the next chunk's MMAs overwrite them
  }

  // ---------------- normalise + store ----------------
  if (q_loc < p.Sq) {
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = p.out + static_cast<size_t>(q_row0 + tid) * p.ldo + head * HD;
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
      uint4 q;
      q.x = pack_bf16x2(o_acc[i + 0] * inv, o_acc[i + 1] * inv);
      q.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
      q.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
      q.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
      *reinterpret_cast<uint4*>(orow + i) = q;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Single-chunk fast path (Sk <= 128, HD == 64, no relative bias): the cross-encoder / BERT-encoder shape that is
// ~11 % of the serving step.  Same math as attn_fwd_kernel, restructured for latency:
//   * 256 threads: TWO threads per query row (warps w and w+4 share TMEM lane quadrant w%4), each owning 64 of the
//     128 key columns, so the serial per-row softmax chain is half as long and 32 warps/SM hide TMEM/MUFU latency
//   * row max of the RAW scores (scale > 0), then one FFMA + EX2 per element; partner halves exchange max / sum
//     through smem with a 64-thread named barrier
//   * both MMAs are issued by warp 0 as fused elect.sync blocks (back-to-back UTCHMMA)
//   * P overwrites the dead Q+K tiles and the normalised O tile overwrites the dead V tile (48 KB smem, 4 CTAs/SM);
//     O leaves through one coalesced TMA store instead of 128-byte-per-thread row stores
// ------------------------------------------------------------------------------------------------
constexpr int kAttn1Threads = 256;

template <int HD>
__global__ void __launch_bounds__(kAttn1Threads, HD == 64 ? 4 : 3)
attn_fwd_1chunk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                       const __grid_constant__ CUtensorMap tmap_q32, const __grid_constant__ CUtensorMap tmap_k32,
                       const __grid_constant__ CUtensorMap tmap_v32, const AttnParams p, int tma_out, int sub_boxes) {
  using Cfg = AttnCfg<HD>;
  constexpr int kRow = Cfg::kRowBytes;  // bytes per Q/K/V/O row: 128 (SW128 tiles) or 64 (SW64 tiles)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kTileBytes;
  uint8_t* sV = sK + Cfg::kTileBytes;
  // P: [2 k-blocks][128 rows][128 B].  head_dim 64: over the dead Q + K tiles (2 x 16 KB); head_dim 32: its own 32 KB
  uint8_t* sP = HD == 64 ? sQ : sV + Cfg::kTileBytes;
  uint8_t* sO = sV;  // [128 rows][kRow] over the dead V tile
  uint64_t* qk_bar = reinterpret_cast<uint64_t*>(sV + Cfg::kTileBytes + (HD == 64 ? 0 : Cfg::kPBytes));
  uint64_t* v_bar = qk_bar + 1;
  uint64_t* s_bar = qk_bar + 2;
  uint64_t* o_bar = qk_bar + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qk_bar + 4);
  float* s_m = reinterpret_cast<float*>(qk_bar + 6);  // [2 halves][128 rows] partial row max (raw score units)
  float* s_l = s_m + 2 * kAttnBQ;                     // [2 halves][128 rows] partial row sums

  const int tid = threadIdx.x;
  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t quad = warp & 3u, half = warp >> 2;
  const int row = static_cast<int>(quad * 32u + lane);  // query row inside the tile == TMEM lane
  const int head = blockIdx.y, b = blockIdx.z;
  pdl_trigger();

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    if (tma_out) tma_prefetch_desc(&tmap_o);
    if (sub_boxes) {
      tma_prefetch_desc(&tmap_q32);
      tma_prefetch_desc(&tmap_k32);
      tma_prefetch_desc(&tmap_v32);
    }
    mbar_init(qk_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (quad * 32u) << 16;

  pdl_wait();  // prologue above overlapped the producer GEMM's tail; q/k/v, kv_lens and cu_seqlens are its outputs
  int q_row0 = b * p.Sq, kv_row0 = b * p.Sk, q_len = p.Sq;
  int kv_len = p.kv_lens ? min(p.kv_lens[b], p.Sk) : p.Sk;
  if (p.cu_seqlens != nullptr) {  // packed self-attention: the 128-row TMA boxes run into the following sequences,
    q_row0 = kv_row0 = p.cu_seqlens[b];  // whose keys are masked and whose query rows are never stored
    q_len = kv_len = min(p.cu_seqlens[b + 1] - q_row0, kAttnBKV);
  }
  int vis_end = kv_len;
  if (p.causal) vis_end = min(vis_end, row + p.causal_offset + 1);

  if (warp == 0) {
    if (lane == 0) {
      const int qg = (min(q_len, kAttnBQ) + 31) >> 5, kg = (kv_len + 31) >> 5;  // 32-row groups that hold real rows
      if (sub_boxes && (qg < 4 || kg < 4)) {
        // short sequence: fetch only the 32-row groups that contain its rows (the 32-row boxes land at the same smem
        // offsets a 128-row box would use; the swizzle pattern repeats every 8 rows).  Whatever stale data sits in the
        // groups that are not loaded is masked (K), zeroed (V, below) or belongs to rows that are never stored (Q).
        constexpr int kGroup = 32 * kRow;
        mbar_expect_tx(qk_bar, (qg + kg) * kGroup);
        for (int g = 0; g < qg; ++g) tma_load_2d(sQ + g * kGroup, &tmap_q32, qk_bar, head * HD, q_row0 + 32 * g);
        for (int g = 0; g < kg; ++g) tma_load_2d(sK + g * kGroup, &tmap_k32, qk_bar, head * HD, kv_row0 + 32 * g);
        mbar_expect_tx(v_bar, kg * kGroup);
        for (int g = 0; g < kg; ++g) tma_load_2d(sV + g * kGroup, &tmap_v32, v_bar, head * HD, kv_row0 + 32 * g);
      } else {
        mbar_expect_tx(qk_bar, 2 * Cfg::kTileBytes);
        tma_load_2d(sQ, &tmap_q, qk_bar, head * HD, q_row0);
        tma_load_2d(sK, &tmap_k, qk_bar, head * HD, kv_row0);
        mbar_expect_tx(v_bar, Cfg::kTileBytes);
        tma_load_2d(sV, &tmap_v, v_bar, head * HD, kv_row0);
      }
    }
    __syncwarp();
    // ---------------- MMA1: S[128q x 128k] = Q K^T (one 64-wide K block) ----------------
    mbar_wait(qk_bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc1 = umma_idesc_f16(kAttnBQ, kAttnBKV);
    if constexpr (HD == 64)
      umma_bf16_kblock64_warp(tmem_base, umma_desc_k_sw128(smem_u32(sQ)), umma_desc_k_sw128(smem_u32(sK)), idesc1, 0u, s_bar);
    else
      umma_bf16_kblock32_warp(tmem_base, umma_desc_k_sw64(smem_u32(sQ)), umma_desc_k_sw64(smem_u32(sK)), idesc1, 0u, s_bar);
  }
  mbar_wait(s_bar, 0);
  tc_fence_after();

  // ---------------- softmax: this thread owns row `row` and two 32-key chunks: [32*half, +32) and [64 + 32*half, +32) ------
  // The two warps of a quadrant used to split the keys as [0,64) / [64,128).  Packed sequences are ~70 keys long, so the
  // second warp had almost nothing to do while the first evaluated 64 exponentials per thread; interleaving the 32-key
  // chunks balances them (38 vs 32 live keys at length 70) and shortens the softmax critical path accordingly.
  //
  // Short (packed) sequences leave most of the 128 x 128 score tile masked; ncu showed this kernel issue-bound (69 % issue
  // slots, MUFU.EX2 + F2FP on the XU pipe at its limit), so masked work is skipped, not computed-and-discarded:
  //   * a warp whose 32 query rows all lie beyond q_len does no softmax at all (its P rows only feed O rows nobody stores);
  //   * 32-key chunks beyond the visible keys are not read from TMEM, and 8-key groups beyond them get P = 0 without
  //     evaluating exp2.  The chunk-level skip must be warp-uniform (tcgen05.ld is .sync.aligned): causal masks vary per row,
  //     so it is disabled for them while the per-group skip still applies per thread.
  const bool warp_live = static_cast<int>(quad * 32u) < q_len;
  float m_raw = kNegBig;
  if (warp_live) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int k0 = 64 * c + 32 * static_cast<int>(half);       // first key of this chunk
      const int n_vis = max(0, min(32, vis_end - k0));            // visible keys in it
      if (!p.causal && n_vis == 0) continue;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + k0, v);
      tmem_ld_wait();
      if (n_vis == 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m_raw = fmaxf(m_raw, __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) m_raw = fmaxf(m_raw, (i < n_vis) ? __uint_as_float(v[i]) : kNegBig);
      }
    }
  }
  s_m[half * kAttnBQ + row] = m_raw;
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // the two warps of this quadrant
  m_raw = fmaxf(m_raw, s_m[(half ^ 1u) * kAttnBQ + row]);
  const float m_s = m_raw * p.scale_log2;  // scale > 0: max of scaled == scaled max
  float l_half = 0.f;
  if (warp_live) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int k0 = 64 * c + 32 * static_cast<int>(half);
      const int n_vis = max(0, min(32, vis_end - k0));
      const bool any = p.causal || n_vis > 0;
      uint32_t v[32];
      if (any) {
        tmem_ld_32x32b_x32(tmem_base + lane_base + k0, v);
        tmem_ld_wait();
      }
      // P k-block c holds keys [64c, 64c + 64) as 8 16-byte chunks per row; this thread's 32 keys are chunks 4*half .. +3
      uint8_t* prow = sP + c * (kAttnBQ * 128) + row * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int g0 = 8 * q4;                          // first key of this 8-wide group inside the chunk
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);          // masked keys: P = 0 (V rows behind the sequence are zeroed below)
        if (any && g0 < n_vis) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = exp2f(fmaf(__uint_as_float(v[g0 + i]), p.scale_log2, -m_s));
            if (g0 + 8 > n_vis) e[i] = (g0 + i < n_vis) ? e[i] : 0.f;
            l_half += e[i];
          }
          pk = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        }
        const int ch = (4 * static_cast<int>(half) + q4) ^ (row & 7);
        *reinterpret_cast<uint4*>(prow + (ch << 4)) = pk;
      }
    }
  }
  s_l[half * kAttnBQ + row] = l_half;
  if (kv_len < kAttnBKV) {
    // P is exactly 0 for keys >= kv_len, but 0 * NaN = NaN: the V rows behind a short sequence may be another
    // sequence's data (finite) or, at the tail of a packed batch, never-written memory.  Zero them in smem.
    mbar_wait(v_bar, 0);
    const int n16 = (kAttnBKV - kv_len) * (kRow / 16);
    uint4* bad = reinterpret_cast<uint4*>(sV + kv_len * kRow);
    for (int c = tid; c < n16; c += kAttn1Threads) bad[c] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();  // st.shared P (and V fix-up) -> visible to tcgen05.mma (async proxy)
  tc_fence_before();
  __syncthreads();           // all S reads done (O aliases S), P complete, partial sums published

  // ---------------- MMA2: O[128q x 64] = P V (8 K=16 steps over the 128 keys) ----------------
  if (warp == 0) {
    mbar_wait(v_bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc2 = umma_idesc_f16(kAttnBQ, HD, 1, false, true);
    const uint64_t pa = umma_desc_k_sw128(smem_u32(sP));
    const uint64_t vb = desc_mn_major<HD>(smem_u32(sV));
    // A: +32 B per step inside a k-block, second k-block 16 KB further; B (MN-major): 16 keys = two 8-row swizzle
    // groups = 2 KB (head_dim 64) or 1 KB (head_dim 32) per step
    constexpr uint32_t kVStep = (2 * Cfg::kGroupBytes) >> 4;
    umma_bf16_x4_warp(tmem_base, pa, vb, 2u, kVStep, idesc2, 0u);
    umma_bf16_x4_warp(tmem_base, pa + ((kAttnBQ * 128) >> 4), vb + 4u * kVStep, 2u, kVStep, idesc2, 1u);
    umma_commit_warp(o_bar);
  }
  const float l_tot = l_half + s_l[(half ^ 1u) * kAttnBQ + row];
  const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
  mbar_wait(o_bar, 0);
  tc_fence_after();

  // ---------------- normalise + store: this thread owns O columns [HD/2 * half, HD/2 * half + HD/2) of its row ------
  {
    constexpr int kCols = HD / 2;       // 32 or 16 fp32 columns per thread
    uint32_t v[kCols];
    if constexpr (HD == 64) tmem_ld_32x32b_x32(tmem_base + lane_base + half * kCols, v);
    else tmem_ld_32x32b_x16(tmem_base + lane_base + half * kCols, v);
    tmem_ld_wait();
    if constexpr (HD == 64) {
      if (p.out_q != nullptr) {
        // ---- MXFP8 epilogue: this thread's 32 columns are one scale block of the out-projection's K dimension ----
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(__uint_as_float(v[i])));
        float sinv;
        const uint32_t e = ue8m0_from_amax(amax * inv, sinv);
        sinv *= inv;
        if (row < q_len) {
          const int grow = q_row0 + row;
          const int col = head * HD + static_cast<int>(half) * 32;
          uint4* dst = reinterpret_cast<uint4*>(p.out_q + static_cast<size_t>(grow) * p.ld_outq + col);
          uint4 q0, q1;
#define IM_F(i) (__uint_as_float(v[i]) * sinv)
          q0.x = pack_e4m3x4(IM_F(0), IM_F(1), IM_F(2), IM_F(3));
          q0.y = pack_e4m3x4(IM_F(4), IM_F(5), IM_F(6), IM_F(7));
          q0.z = pack_e4m3x4(IM_F(8), IM_F(9), IM_F(10), IM_F(11));
          q0.w = pack_e4m3x4(IM_F(12), IM_F(13), IM_F(14), IM_F(15));
          q1.x = pack_e4m3x4(IM_F(16), IM_F(17), IM_F(18), IM_F(19));
          q1.y = pack_e4m3x4(IM_F(20), IM_F(21), IM_F(22), IM_F(23));
          q1.z = pack_e4m3x4(IM_F(24), IM_F(25), IM_F(26), IM_F(27));
          q1.w = pack_e4m3x4(IM_F(28), IM_F(29), IM_F(30), IM_F(31));
#undef IM_F
          dst[0] = q0;
          dst[1] = q1;
          p.out_sf[(static_cast<size_t>(grow >> 7) * p.n_kb + (col >> 7)) * 512 + (grow & 31) * 16 + ((grow & 127) >> 5) * 4 +
                   ((col & 127) >> 5)] = static_cast<uint8_t>(e);
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
        return;
      }
    }
    uint4 q[kCols / 8];
#pragma unroll
    for (int j = 0; j < kCols / 8; ++j) {
      q[j].x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
      q[j].y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
      q[j].z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
      q[j].w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
    }
    // stage the normalised [128 x HD] tile in smem over the dead V tile: 128B-swizzled rows for head_dim 64 (the layout
    // the TMA store expects), plain rows for head_dim 32
#pragma unroll
    for (int j = 0; j < kCols / 8; ++j) {
      int ch = static_cast<int>(half) * (kCols / 8) + j;
      if constexpr (HD == 64) ch ^= (row & 7);
      *reinterpret_cast<uint4*>(sO + row * kRow + (ch << 4)) = q[j];
    }
  }
  tc_fence_before();
  if (tma_out) fence_proxy_async_smem();
  __syncthreads();
  if (tma_out) {
    if (tid == 0) {
      tma_store_2d(&tmap_o, sO, head * HD, q_row0);
      tma_store_commit();
      tma_store_wait_read<0>();  // smem must stay valid until the bulk store has read it
    }
  } else {
    // short / packed sequences: only rows < q_len exist.  kRow/16 lanes move one row, so a warp writes 4 (or 8) full
    // rows per instruction instead of 32 scattered 16-byte pieces.
    constexpr int kLanes = kRow / 16;
    const int c = tid & (kLanes - 1);
    for (int r = tid / kLanes; r < q_len; r += kAttn1Threads / kLanes) {
      const int sc = HD == 64 ? (c ^ (r & 7)) : c;
      const uint4 val = *reinterpret_cast<const uint4*>(sO + r * kRow + (sc << 4));
      *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(q_row0 + r) * p.ldo + head * HD + c * 8) = val;
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// Single-query decode attention against a KV cache (T5 decoder step, CLS-only last cross-encoder layer): one warp per
// (batch, head), ONE KEY PER LANE.  Each lane keeps the query (HD registers), computes whole dot products for keys
// lane, lane+32, ... with its own online-softmax state and a private HD-wide accumulator; the 32 partial states are
// merged once at the end (log-sum-exp rescale + warp reductions).  No per-key shuffles -- the earlier
// one-key-per-iteration version spent ~100 us on 256 keys in its dependent shuffle chain.
// q: [B, nH*HD]; k/v cache: [B, S_max, nH*HD]; keys [0, kv_len[b]) visible.
template <int HD>
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ kc,
                   const __nv_bfloat16* __restrict__ vc, int ld_kv, int s_max, const int* __restrict__ kv_lens,
                   int kv_len_all, float scale_log2, const float* __restrict__ rel_bias, int bias_len, int q_pos,
                   int n_heads, int n_pairs, __nv_bfloat16* __restrict__ out, int ldo,
                   const int* __restrict__ seq_start, const int* __restrict__ step_dev) {
  const int pair = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const int b = pair / n_heads, h = pair % n_heads;
  // step_dev: decode step t kept on the device (CUDA-graph replay of one decoder step): keys [0, t], query position t
  if (step_dev != nullptr) {
    q_pos = *step_dev;
    kv_len_all = q_pos + 1;
  }
  const int len = kv_lens ? kv_lens[b] : kv_len_all;
  float qv[HD];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + static_cast<size_t>(b) * ldq + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = __ldg(qp + c);
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      qv[8 * c + 0] = f0.x; qv[8 * c + 1] = f0.y; qv[8 * c + 2] = f1.x; qv[8 * c + 3] = f1.y;
      qv[8 * c + 4] = f2.x; qv[8 * c + 5] = f2.y; qv[8 * c + 6] = f3.x; qv[8 * c + 7] = f3.y;
    }
  }
  float m = kNegBig, l = 0.f, acc[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) acc[i] = 0.f;
  // seq_start: keys of sequence b begin at row seq_start[b] of a packed (unpadded) K/V buffer instead of b * s_max
  const size_t row0 = seq_start != nullptr ? static_cast<size_t>(seq_start[b]) : static_cast<size_t>(b) * s_max;
  const __nv_bfloat16* kb = kc + row0 * ld_kv + h * HD;
  const __nv_bfloat16* vb = vc + row0 * ld_kv + h * HD;
  for (int j = lane; j < len; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + static_cast<size_t>(j) * ld_kv);
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = kr[c];
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      d = fmaf(qv[8 * c + 0], f0.x, d); d = fmaf(qv[8 * c + 1], f0.y, d);
      d = fmaf(qv[8 * c + 2], f1.x, d); d = fmaf(qv[8 * c + 3], f1.y, d);
      d = fmaf(qv[8 * c + 4], f2.x, d); d = fmaf(qv[8 * c + 5], f2.y, d);
      d = fmaf(qv[8 * c + 6], f3.x, d); d = fmaf(qv[8 * c + 7], f3.y, d);
    }
    float sc = d * scale_log2;
    if (rel_bias) {
      int bi = j - q_pos + (bias_len - 1) / 2;  // table centred on relative position 0
      bi = max(0, min(bias_len - 1, bi));
      sc += rel_bias[static_cast<size_t>(h) * bias_len + bi];
    }
    const float m_new = fmaxf(m, sc);
    const float a = exp2f(m - m_new), pj = exp2f(sc - m_new);
    l = fmaf(l, a, pj);
    m = m_new;
    const uint4* vr = reinterpret_cast<const uint4*>(vb + static_cast<size_t>(j) * ld_kv);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = vr[c];
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      acc[8 * c + 0] = fmaf(acc[8 * c + 0], a, pj * f0.x); acc[8 * c + 1] = fmaf(acc[8 * c + 1], a, pj * f0.y);
      acc[8 * c + 2] = fmaf(acc[8 * c + 2], a, pj * f1.x); acc[8 * c + 3] = fmaf(acc[8 * c + 3], a, pj * f1.y);
      acc[8 * c + 4] = fmaf(acc[8 * c + 4], a, pj * f2.x); acc[8 * c + 5] = fmaf(acc[8 * c + 5], a, pj * f2.y);
      acc[8 * c + 6] = fmaf(acc[8 * c + 6], a, pj * f3.x); acc[8 * c + 7] = fmaf(acc[8 * c + 7], a, pj * f3.y);
    }
  }
  // merge the 32 per-lane softmax states
  float m_all = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
  const float f = exp2f(m - m_all);   // lanes that saw no key: m = kNegBig -> 0
  float l_all = l * f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l_all += __shfl_xor_sync(0xffffffffu, l_all, o);
  const float inv = l_all > 0.f ? 1.f / l_all : 0.f;
  constexpr int PER = HD / 32;  // output dims written per lane (1 or 2)
  float mine[PER];
#pragma unroll
  for (int i = 0; i < HD; ++i) {
    float t = acc[i] * f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (i / PER == lane) mine[i % PER] = t;
  }
#pragma unroll
  for (int i = 0; i < PER; ++i)
    out[static_cast<size_t>(b) * ldo + h * HD + lane * PER + i] = __float2bfloat16(mine[i] * inv);
}

// Decoder KV-cache append with the step on the device: row *step_dev of sequence b in kc / vc <- the K / V slices of
// this step's fused QKV row (one block per sequence).  Lets one decoder step replay from a CUDA graph.
__global__ void __launch_bounds__(128)
kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int inner, __nv_bfloat16* __restrict__ kc,
                 __nv_bfloat16* __restrict__ vc, int s_max, const int* __restrict__ step_dev) {
  const int b = blockIdx.x, t = *step_dev;
  if (t < 0 || t >= s_max) return;
  const uint4* src_k = reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(b) * ld_qkv + inner);
  const uint4* src_v = reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(b) * ld_qkv + 2 * inner);
  uint4* dst_k = reinterpret_cast<uint4*>(kc + (static_cast<size_t>(b) * s_max + t) * inner);
  uint4* dst_v = reinterpret_cast<uint4*>(vc + (static_cast<size_t>(b) * s_max + t) * inner);
  for (int c = threadIdx.x; c < inner / 8; c += blockDim.x) {
    dst_k[c] = src_k[c];
    dst_v[c] = src_v[c];
  }
}

}  // namespace im

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* out, int B, int n_heads, int head_dim, int Sq,
                         int Sk, int ldq, int ldk, int ldv, int ldo, const int* kv_lens, int causal, int causal_offset,
                         float scale, const float* rel_bias_log2, void* stream, const int* cu_seqlens, void* out_q,
                         int ld_outq, void* out_sf, int n_kb) {
  using namespace im;
  if (out_q != nullptr && !(head_dim == 64 && Sk <= kAttnBKV && Sq <= kAttnBQ && rel_bias_log2 == nullptr && scale > 0.f &&
                            (ld_outq % 16) == 0))
    return set_error("im_attn_fwd_mx", "MXFP8 output: single-chunk path only (head_dim 64, S <= 128, no bias)");
  if (B <= 0 || Sq <= 0 || Sk <= 0) return 0;
  if (cu_seqlens != nullptr && !(Sq == Sk && Sk <= kAttnBKV && rel_bias_log2 == nullptr && !causal && scale > 0.f))
    return set_error("im_attn_fwd", "packed (cu_seqlens) attention: self-attention, max_seqlen <= 128, no bias");
  if (head_dim != 64 && head_dim != 32) return set_error("im_attn_fwd", "head_dim must be 32 or 64");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return set_error("im_attn_fwd", "row pitches must be multiples of 8");
  const TmapSwizzle sw = head_dim == 64 ? TMAP_SW_128 : TMAP_SW_64;
  CUtensorMap tq, tk, tv;
  const uint64_t cols = static_cast<uint64_t>(n_heads) * head_dim;
  if (get_tmap_2d(&tq, q, static_cast<uint64_t>(B) * Sq, cols, static_cast<uint64_t>(ldq) * 2, kAttnBQ, head_dim, 2, sw))
    return -1;
  if (get_tmap_2d(&tk, k, static_cast<uint64_t>(B) * Sk, cols, static_cast<uint64_t>(ldk) * 2, kAttnBKV, head_dim, 2, sw))
    return -1;
  if (get_tmap_2d(&tv, v, static_cast<uint64_t>(B) * Sk, cols, static_cast<uint64_t>(ldv) * 2, kAttnBKV, head_dim, 2, sw))
    return -1;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.Sq = Sq;
  p.Sk = Sk;
  p.kv_lens = kv_lens;
  p.causal = causal;
  p.causal_offset = causal_offset;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rel_bias = rel_bias_log2;
  p.alias_p = (head_dim == 64 && Sk <= kAttnBKV) ? 1 : 0;
  p.cu_seqlens = cu_seqlens;
  p.cp_world = 1;
  p.cp_rank = p.cp_sk_local = p.cp_q_pos0 = 0;
  CpMaps no_cp;
  for (int i = 0; i < 8; ++i) no_cp.k[i] = no_cp.v[i] = tk;
  p.out_q = reinterpret_cast<uint8_t*>(out_q);
  p.out_sf = reinterpret_cast<uint8_t*>(out_sf);
  p.ld_outq = ld_outq;
  p.n_kb = n_kb;
  const int bias_bytes = rel_bias_log2 ? (Sq + Sk) * 4 : 0;
  dim3 grid((Sq + kAttnBQ - 1) / kAttnBQ, n_heads, B);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (Sk <= kAttnBKV && Sq <= kAttnBQ && rel_bias_log2 == nullptr && scale > 0.f) {
    // single-chunk fast path; head_dim 64: O goes out through TMA when a 128-row box cannot spill into the next sequence
    const int tma_out = (head_dim == 64 && Sq == kAttnBQ && cu_seqlens == nullptr && out_q == nullptr) ? 1 : 0;
    CUtensorMap to = tq;
    if (tma_out &&
        get_tmap_2d(&to, out, static_cast<uint64_t>(B) * Sq, cols, static_cast<uint64_t>(ldo) * 2, kAttnBQ, head_dim, 2, sw))
      return -1;
    // 32-row boxes for sequences that do not fill the 128-row tile (kv_lens / cu_seqlens known only on the device)
    const int sub_boxes = (kv_lens != nullptr || cu_seqlens != nullptr || Sq < kAttnBQ) ? 1 : 0;
    CUtensorMap tq32 = tq, tk32 = tk, tv32 = tv;
    if (sub_boxes) {
      if (get_tmap_2d(&tq32, q, static_cast<uint64_t>(B) * Sq, cols, static_cast<uint64_t>(ldq) * 2, 32, head_dim, 2, sw)) return -1;
      if (get_tmap_2d(&tk32, k, static_cast<uint64_t>(B) * Sk, cols, static_cast<uint64_t>(ldk) * 2, 32, head_dim, 2, sw)) return -1;
      if (get_tmap_2d(&tv32, v, static_cast<uint64_t>(B) * Sk, cols, static_cast<uint64_t>(ldv) * 2, 32, head_dim, 2, sw)) return -1;
    }
    if (head_dim == 64) {
      const int smem = 3 * AttnCfg<64>::kTileBytes + 1024 + 64 + 4 * kAttnBQ * 4;
      IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_1chunk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      IM_CUDA_OK(launch_pdl(attn_fwd_1chunk_kernel<64>, dim3(1, n_heads, B), dim3(kAttn1Threads), smem, s, tq, tk, tv, to, tq32,
                            tk32, tv32, p, tma_out, sub_boxes));
    } else {
      const int smem = 3 * AttnCfg<32>::kTileBytes + AttnCfg<32>::kPBytes + 1024 + 64 + 4 * kAttnBQ * 4;
      IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_1chunk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      IM_CUDA_OK(launch_pdl(attn_fwd_1chunk_kernel<32>, dim3(1, n_heads, B), dim3(kAttn1Threads), smem, s, tq, tk, tv, to, tq32,
                            tk32, tv32, p, tma_out, sub_boxes));
    }
  } else if (head_dim == 64) {
    const int smem = 3 * AttnCfg<64>::kTileBytes + (p.alias_p ? 0 : AttnCfg<64>::kPBytes) + 1024 + 64 + bias_bytes;
    IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IM_CUDA_OK(launch_pdl(attn_fwd_kernel<64, false>, grid, dim3(kAttnThreads), smem, s, tq, tk, tv, p, no_cp));
  } else {
    const int smem = 3 * AttnCfg<32>::kTileBytes + AttnCfg<32>::kPBytes + 1024 + 64 + bias_bytes;
    IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IM_CUDA_OK(launch_pdl(attn_fwd_kernel<32, false>, grid, dim3(kAttnThreads), smem, s, tq, tk, tv, p, no_cp));
  }
  IM_LAUNCH_OK("attn_fwd_kernel");
  return 0;
}

// q: [B*Sq, ldq] with head h at column h*HD (pass base pointer already offset to the Q block); k, v likewise.
IM_API int im_attn_fwd(const void* q, const void* k, const void* v, void* out, int B, int n_heads, int head_dim, int Sq,
                       int Sk, int ldq, int ldk, int ldv, int ldo, const int* kv_lens, int causal, int causal_offset,
                       float scale, const float* rel_bias_log2, void* stream, const int* cu_seqlens) {
  return attn_fwd_impl(q, k, v, out, B, n_heads, head_dim, Sq, Sk, ldq, ldk, ldv, ldo, kv_lens, causal, causal_offset, scale,
                       rel_bias_log2, stream, cu_seqlens, nullptr, 0, nullptr, 0);
}
// Same, with the context written as MXFP8 (e4m3 bytes [rows, ld_outq] + SFA scale chunks with n_kb k-blocks per row block).
IM_API int im_attn_fwd_mx(const void* q, const void* k, const void* v, void* out_q, int ld_outq, void* out_sf, int n_kb, int B,
                          int n_heads, int head_dim, int Sq, int Sk, int ldq, int ldk, int ldv, const int* kv_lens, float scale,
                          void* stream, const int* cu_seqlens) {
  return attn_fwd_impl(q, k, v, nullptr, B, n_heads, head_dim, Sq, Sk, ldq, ldk, ldv, 8, kv_lens, 0, 0, scale, nullptr, stream,
                       cu_seqlens, out_q, ld_outq, out_sf, n_kb);
}

// Context-parallel attention: this rank's queries q [B * Sq_local, ldq] against the keys / values of ALL ranks.  k_ptrs /
// v_ptrs: host arrays of `world` device pointers to every rank's K / V buffer ([B * sk_local, ldk], peer-mapped through
// the symmetric heap).  The kernel pulls the tiles itself (TMA over NVLink); the caller only has to make sure every
// rank's K / V are complete (one heap barrier after the QKV projection).  kv_lens are GLOBAL key counts per sequence;
// rel_bias_log2 is the [heads, 2 * S_total - 1] table of the whole sequence.
IM_API int im_attn_fwd_cp(const void* q, const void* const* k_ptrs, const void* const* v_ptrs, void* out, int B, int n_heads,
                          int head_dim, int sq_local, int sk_local, int world, int rank, int ldq, int ldk, int ldv, int ldo,
                          const int* kv_lens, float scale, const float* rel_bias_log2, void* stream) {
  using namespace im;
  if (B <= 0 || sq_local <= 0 || sk_local <= 0) return 0;
  if (world < 1 || world > 8) return set_error("im_attn_fwd_cp", "context-parallel world must be 1..8");
  if (sk_local % kAttnBKV) return set_error("im_attn_fwd_cp", "keys per rank must be a multiple of 128");
  if (head_dim != 64 && head_dim != 32) return set_error("im_attn_fwd_cp", "head_dim must be 32 or 64");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return set_error("im_attn_fwd_cp", "row pitches must be multiples of 8");
  const TmapSwizzle sw = head_dim == 64 ? TMAP_SW_128 : TMAP_SW_64;
  const uint64_t cols = static_cast<uint64_t>(n_heads) * head_dim;
  CUtensorMap tq;
  if (get_tmap_2d(&tq, q, static_cast<uint64_t>(B) * sq_local, cols, static_cast<uint64_t>(ldq) * 2, kAttnBQ, head_dim, 2, sw)) return -1;
  CpMaps cpm;
  for (int r = 0; r < 8; ++r) {
    const int src = r < world ? r : 0;
    if (get_tmap_2d(&cpm.k[r], k_ptrs[src], static_cast<uint64_t>(B) * sk_local, cols, static_cast<uint64_t>(ldk) * 2, kAttnBKV, head_dim, 2, sw)) return -1;
    if (get_tmap_2d(&cpm.v[r], v_ptrs[src], static_cast<uint64_t>(B) * sk_local, cols, static_cast<uint64_t>(ldv) * 2, kAttnBKV, head_dim, 2, sw)) return -1;
  }
  const int sk_total = sk_local * world;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.Sq = sq_local;
  p.Sk = sk_total;
  p.kv_lens = kv_lens;
  p.causal = 0;
  p.causal_offset = 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rel_bias = rel_bias_log2;
  p.alias_p = 0;
  p.cu_seqlens = nullptr;
  p.cp_world = world;
  p.cp_rank = rank;
  p.cp_sk_local = sk_local;
  p.cp_q_pos0 = rank * sq_local;
  p.out_q = nullptr;
  p.out_sf = nullptr;
  p.ld_outq = p.n_kb = 0;
  const int bias_bytes = rel_bias_log2 ? 2 * sk_total * 4 : 0;
  dim3 grid((sq_local + kAttnBQ - 1) / kAttnBQ, n_heads, B);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) {
    const int smem = 3 * AttnCfg<64>::kTileBytes + AttnCfg<64>::kPBytes + 1024 + 64 + bias_bytes;
    IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IM_CUDA_OK(launch_pdl(attn_fwd_kernel<64, true>, grid, dim3(kAttnThreads), smem, s, tq, cpm.k[0], cpm.v[0], p, cpm));
  } else {
    const int smem = 3 * AttnCfg<32>::kTileBytes + AttnCfg<32>::kPBytes + 1024 + 64 + bias_bytes;
    IM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IM_CUDA_OK(launch_pdl(attn_fwd_kernel<32, true>, grid, dim3(kAttnThreads), smem, s, tq, cpm.k[0], cpm.v[0], p, cpm));
  }
  IM_LAUNCH_OK("attn_fwd_kernel<cp>");
  return 0;
}

IM_API int im_attn_decode(const void* q, int ldq, const void* kc, const void* vc, int ld_kv, int s_max,
                          const int* kv_lens, int kv_len_all, float scale, const float* rel_bias_log2, int bias_len,
                          int q_pos, int B, int n_heads, int head_dim, void* out, int ldo, void* stream,
                          const int* seq_start, const int* step_dev) {
  using namespace im;
  const int n_pairs = B * n_heads;
  if (n_pairs <= 0) return 0;
  const float sl2 = scale * 1.4426950408889634f;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (n_pairs + 3) / 4;
  if (head_dim == 64)
    attn_decode_kernel<64><<<grid, 128, 0, s>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kc,
                                               (const __nv_bfloat16*)vc, ld_kv, s_max, kv_lens, kv_len_all, sl2,
                                               rel_bias_log2, bias_len, q_pos, n_heads, n_pairs, (__nv_bfloat16*)out, ldo, seq_start, step_dev);
  else if (head_dim == 32)
    attn_decode_kernel<32><<<grid, 128, 0, s>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kc,
                                               (const __nv_bfloat16*)vc, ld_kv, s_max, kv_lens, kv_len_all, sl2,
                                               rel_bias_log2, bias_len, q_pos, n_heads, n_pairs, (__nv_bfloat16*)out, ldo, seq_start, step_dev);
  else
    return set_error("im_attn_decode", "head_dim must be 32 or 64");
  IM_LAUNCH_OK("attn_decode_kernel");
  return 0;
}

IM_API int im_kv_append(const void* qkv, int ld_qkv, int inner, void* kc, void* vc, int s_max, int B, const int* step_dev,
                        void* stream) {
  using namespace im;
  if (B <= 0) return 0;
  if (inner % 8 || ld_qkv % 8) return set_error("im_kv_append", "inner and the qkv pitch must be multiples of 8");
  kv_append_kernel<<<B, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const __nv_bfloat16*)qkv, ld_qkv, inner,
                                                                         (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, s_max,
                                                                         step_dev);
  IM_LAUNCH_OK("kv_append_kernel");
  return 0;
}
