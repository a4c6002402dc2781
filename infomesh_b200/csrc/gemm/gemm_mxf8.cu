// K6 (block-scaled) — persistent warp-specialised MXFP8 GEMM for sm_100a.
//
//   C[M,N] = epilogue( (A ⊙ SFA)[M,K] · (B ⊙ SFB)[N,K]^T )
//
// A, B: e4m3 bytes, K contiguous.  SFA / SFB: one ue8m0 scale per 32 K-elements per row, stored in 512-byte chunks that
// are already in the shape `tcgen05.cp.32x128b.warpx4` wants (see ptx.cuh): chunk (rb, kb) covers rows [128 rb, +128) and
// K elements [128 kb, +128); byte (r % 32) * 16 + (r / 32) * 4 + (k / 32) % 4.  SFA is [ceil(M/128)][K/128] chunks,
// SFB is [K/128][ceil(N/128) (+1 pad)] chunks (k-block major so the two chunks a 192-row tile straddles are contiguous).
//
// The MMAs are `tcgen05.mma.kind::mxf8f6f4.block_scale` 128 x 192 x 32 with fp32 accumulators in TMEM; scale chunks
// travel global -> smem with `cp.async.bulk` on the same mbarrier as the TMA operand tiles and smem -> TMEM with
// `tcgen05.cp` issued by the MMA thread right in front of the four MMAs of the k-block.  BN = 192 because two 256-column
// accumulators would fill all 512 TMEM columns and leave none for the scales: 2 x 192 + 4 (SFA) + 8 (SFB) = 396.
// On odd N tiles the 192 rows start 64 rows into a 128-row scale chunk, so the SFB operand address is shifted by two
// TMEM columns (the same trick CUTLASS' sm100 block-scaled collective uses for CtaN = 192).
//
// A-resident variant (ARES, K <= 768): at K = 768 a 128 x 192 tile needs 240 KB of operands for ~2300 tensor cycles, and
// 148 SMs asking for that saturate the ~12 TB/s the L2 can deliver to TMA before the tensor pipes are busy.  The
// activation block [128 rows x K] (<= 96 KB) is therefore loaded ONCE per work unit and stays in shared memory while
// the unit's G consecutive N tiles stream only their weight tiles through a 3-stage ring (-27..37 % operand traffic).
// Each A k-block has its own full / empty barrier pair, so the next unit's A[kb] is fetched as soon as the last N tile
// of the current unit has consumed A[kb] -- no bubble between units.
//
// Epilogues (12 warps = 3 per TMEM lane quadrant, 64 columns each; both 32-column TMEM loads of a warp are issued before
// the first is consumed -- at K = 768 a tile is only 6 k-blocks (~2300 tensor cycles), so the epilogue's latency chain,
// not its instruction count, decides whether the tensor pipe stays busy; TMEM -> registers -> swizzled smem -> TMA store):
//   out_mx = 0: bf16  C = act(acc + bias) + residual
//   out_mx = 1: MXFP8 C = quantise(act(acc + bias)): per row and 32 columns amax -> ue8m0 -> e4m3, scales written in the
//               SFA chunk layout of the NEXT GEMM (so FFN-up -> GELU -> FFN-down never materialises bf16 activations and
//               no standalone quantiser kernel runs).
// Replaces the fp32/cuBLAS linear layers behind infomesh/index/vector_store.py:104-125 and the external LLM reranker
// (infomesh/search/reranker.py:124-159).
#include "../common/host.h"
#include "../common/ptx.cuh"
#include "../common/tmap_cache.h"

namespace im {

struct MxEpilogue {
  const float* bias;        // [N] or null
  uint8_t* c_sf;            // out_mx: scale chunks of the output, [ceil(M/128)][N/128][512]
  const uint8_t* sfa;       // [ceil(M/128)][K/128][512]
  const uint8_t* sfb;       // [K/128][n_chunks_b][512]
  int n_chunks_b;           // chunks per k-block in sfb (>= ceil(N/128) + 1 so an odd last tile may over-read one chunk)
  int act;                  // 0 none, 1 gelu(erf), 2 relu, 3 gelu(tanh), 4 tanh
  int out_mx;
  int has_res;
  const int* m_dev;         // optional device-side row count (varlen batches under CUDA graphs)
  int n_per_unit;           // G: consecutive N tiles per work unit (1 unless the A-resident variant is used)
  // fused LayerNorm variant (LNF): C = LN(acc + bias + residual) in bf16 AND its MXFP8 copy (q_out + c_sf)
  const float* ln_g;        // [N]
  const float* ln_b;        // [N] or null
  float ln_eps;
  uint8_t* q_out;           // [M, ldq] e4m3 bytes of the normalised rows
  int ldq;
};

constexpr int kMxBM = 128;
constexpr int kMxBN = 192;
constexpr int kMxBK = 128;                 // fp8 elements (= bytes) per k-block
constexpr int kMxStages = 4;
constexpr int kMxEpiWarps = 12;
constexpr int kMxThreads = 64 + 32 * kMxEpiWarps;
constexpr int kMxABBytes = (kMxBM + kMxBN) * kMxBK;           // 40960: multiple of 1024 (swizzle-128 alignment)
constexpr int kMxSfBytes = 512 + 1024;                        // SFA chunk + two SFB chunks
constexpr int kMxStoreTile = 2048;                            // [32 rows x 32 bf16] or [32 rows x 32 B] staging tile
constexpr int kMxChunks = 2;                                  // 32-column chunks per epilogue warp and tile (3 warps x 64 = 192)
constexpr int kMxStoreBytes = kMxEpiWarps * kMxChunks * kMxStoreTile;
constexpr int kMxTmemCols = 512;
constexpr int kMxSfaCol = 2 * kMxBN;                          // 384
constexpr int kMxSfbCol = kMxSfaCol + 4;                      // 388
constexpr int kMxBarBytes = 512;
constexpr int kMxSmemBytes = kMxStages * (kMxABBytes + kMxSfBytes) + kMxStoreBytes + 1024 + kMxBarBytes;
// fused-LayerNorm variant: a cluster of N / 192 CTAs owns one 128-row block; 3 operand stages make room for the row statistics
constexpr int kLnfStages = 3;
constexpr int kLnfMaxCtas = 4;                                // N <= 768
constexpr int kLnfColWarps = kMxEpiWarps / 4;                 // 3 epilogue warps (64 columns each) per TMEM lane quadrant
constexpr int kLnfStatBytes = 2 * kLnfMaxCtas * kLnfColWarps * kMxBM * 8;     // [buf][src cta][col warp][row] (sum, sum sq)
constexpr int kLnfSfRing = ((kLnfStages * kMxSfBytes + 1023) / 1024) * 1024;
constexpr int kLnfColBytes = 3 * kMxBN * 4;                   // bias | gamma | beta of this CTA's 192 columns
constexpr int kLnfSmemBytes = kLnfStages * kMxABBytes + kLnfSfRing + kMxStoreBytes + 1024 + kMxBarBytes + kLnfStatBytes + kLnfColBytes;
// A-resident variant
constexpr int kAresMaxKb = 6;                                 // K <= 768
constexpr int kAresBStages = 3;
constexpr int kAresABytes = kMxBM * kMxBK;                    // 16384 per k-block
constexpr int kAresBBytes = kMxBN * kMxBK;                    // 24576 per stage
constexpr int kAresSmemBytes = kAresMaxKb * (kAresABytes + 512) + kAresBStages * (kAresBBytes + 1024) + kMxStoreBytes + 1024 +
                               kMxBarBytes;

__device__ __forceinline__ void mx_act32(float (&f)[32], int act) {
  switch (act) {
    case 1: {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t g = gelu_erf2(pk2(f[2 * i], f[2 * i + 1]));
        upk2(g, f[2 * i], f[2 * i + 1]);
      }
    } break;
    case 2:
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
      break;
    case 3:
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = gelu_tanh(f[i]);
      break;
    case 4:
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = tanh_approx(f[i]);
      break;
    default:
      break;
  }
}

template <bool ARES, bool LNF = false>
__global__ void __launch_bounds__(kMxThreads, 1)
gemm_mxf8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r,
                 const MxEpilogue ep, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  // ring layout: [stage][A 16K | B 24K] ... [stage][SFA 512 | SFB 1024];  A-resident: A[kb] x6, B[stage] x3, SFA[kb] x6, SFB[stage] x3
  uint8_t* sA = smem;
  uint8_t* sB = ARES ? smem + kAresMaxKb * kAresABytes : smem + kMxBM * kMxBK;
  static_assert(!(ARES && LNF), "the fused-LayerNorm epilogue runs on the operand-ring kernel");
  constexpr int kRing = LNF ? kLnfStages : kMxStages;
  uint8_t* sSFA = ARES ? sB + kAresBStages * kAresBBytes : smem + kRing * kMxABBytes;
  uint8_t* sSFB = ARES ? sSFA + kAresMaxKb * 512 : sSFA + 512;
  uint8_t* smem_store = ARES ? sSFB + kAresBStages * 1024                                   // 1024-aligned in every layout
                             : sSFA + (LNF ? kLnfSfRing : kMxStages * kMxSfBytes);
  constexpr uint32_t kAStride = ARES ? kAresABytes : kMxABBytes;      // bytes between consecutive A slots
  constexpr uint32_t kBStride = ARES ? kAresBBytes : kMxABBytes;
  constexpr uint32_t kSfaStride = ARES ? 512 : kMxSfBytes;
  constexpr uint32_t kSfbStride = ARES ? 1024 : kMxSfBytes;
  constexpr uint32_t kBStages = ARES ? kAresBStages : kRing;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_store + kMxStoreBytes);
  uint64_t* empty_bar = full_bar + kMxStages;
  uint64_t* tmem_full = empty_bar + kMxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;      // [12] residual tiles of one epilogue warp
  uint64_t* a_full = res_bar + kMxEpiWarps;   // [6] A-resident: k-block kb of the unit's activation block has landed
  uint64_t* a_empty = a_full + kAresMaxKb;    // [6] ... and has been consumed by the unit's last N tile
  uint64_t* stat_bar = a_empty + kAresMaxKb;  // [2] LNF: every epilogue warp of every CTA of the cluster has published its row sums
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_bar + 2);
  float2* stats = reinterpret_cast<float2*>(smem_store + kMxStoreBytes + kMxBarBytes);   // LNF only
  float* s_cols = reinterpret_cast<float*>(smem_store + kMxStoreBytes + kMxBarBytes + kLnfStatBytes);   // LNF: [3][192]

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kMxEpiWarps);
    }
    for (int i = 0; i < kMxEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < kAresMaxKb; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    if constexpr (LNF) {
      for (int i = 0; i < 2; ++i) mbar_init(&stat_bar[i], 1);     // one local expect_tx arrival; the data arrive as transactions
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kMxTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (LNF) cluster_sync_all();     // no peer may arrive on a barrier that is not initialised yet
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();
  if (ep.m_dev != nullptr) M = min(M, max(0, *ep.m_dev));
  const int num_m = (M + kMxBM - 1) / kMxBM;
  const int num_n = (N + kMxBN - 1) / kMxBN;
  const int num_k = K / kMxBK;
  // work units: (m block, group of G consecutive N tiles); G = 1 makes a unit one tile in row-major tile order
  const int G = ARES ? max(1, ep.n_per_unit) : 1;
  const int n_groups = (num_n + G - 1) / G;
  const int num_units = num_m * n_groups;

  if (warp == 0) {
    // ------------------------------- producer: TMA tiles + bulk scale chunks -------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, unit_i = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++unit_i) {
        const int m_blk = u / n_groups, n0 = (u % n_groups) * G, n1 = min(num_n, n0 + G);
        for (int n_blk = n0; n_blk < n1; ++n_blk) {
          const int chunk_b = (n_blk * kMxBN) >> 7;     // first 128-row scale chunk the tile touches
          for (int kb = 0; kb < num_k; ++kb) {
            if constexpr (ARES) {
              if (n_blk == n0) {    // the unit's activation k-block: once, into its own slot
                mbar_wait(&a_empty[kb], (unit_i & 1u) ^ 1u);
                mbar_expect_tx(&a_full[kb], kAresABytes + 512);
                tma_load_2d(sA + kb * kAStride, &tmap_a, &a_full[kb], kb * kMxBK, m_blk * kMxBM);
                bulk_load(sSFA + kb * kSfaStride, ep.sfa + (static_cast<size_t>(m_blk) * num_k + kb) * 512, 512, &a_full[kb]);
              }
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], ARES ? kAresBBytes + 1024 : kMxABBytes + kMxSfBytes);
            if constexpr (!ARES) {
              tma_load_2d(sA + stage * kAStride, &tmap_a, &full_bar[stage], kb * kMxBK, m_blk * kMxBM);
              bulk_load(sSFA + stage * kSfaStride, ep.sfa + (static_cast<size_t>(m_blk) * num_k + kb) * 512, 512, &full_bar[stage]);
            }
            tma_load_2d(sB + stage * kBStride, &tmap_b, &full_bar[stage], kb * kMxBK, n_blk * kMxBN);
            bulk_load(sSFB + stage * kSfbStride, ep.sfb + (static_cast<size_t>(kb) * ep.n_chunks_b + chunk_b) * 512, 1024,
                      &full_bar[stage]);
            if (++stage == kBStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (whole warp, one elected lane issues) ----------------------
    const uint32_t idesc = umma_idesc_mxf8(kMxBM, kMxBN);
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(sA));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(sB));
    const uint64_t sfa_desc0 = umma_desc_sf_chunk(smem_u32(sSFA));
    const uint64_t sfb_desc0 = umma_desc_sf_chunk(smem_u32(sSFB));
    const uint32_t t_sfa = tmem_base + kMxSfaCol;
    const uint32_t t_sfb_cp = tmem_base + kMxSfbCol;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, unit_i = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++unit_i) {
      const int n0 = (u % n_groups) * G, n1 = min(num_n, n0 + G);
      for (int n_blk = n0; n_blk < n1; ++n_blk) {
        // odd tiles start 64 rows into their first scale chunk: skip two TMEM columns of SFB
        const uint32_t t_sfb = t_sfb_cp + (((n_blk * kMxBN) & 127) >> 5);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMxBN;
        for (int kb = 0; kb < num_k; ++kb) {
          if constexpr (ARES) {
            if (n_blk == n0) mbar_wait(&a_full[kb], unit_i & 1u);
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_slot = ARES ? static_cast<uint32_t>(kb) : stage;
          umma_mxf8_kblock128_warp(d_tmem, a_desc0 + static_cast<uint64_t>(a_slot * (kAStride >> 4)),
                                   b_desc0 + static_cast<uint64_t>(stage * (kBStride >> 4)), idesc, t_sfa, t_sfb_cp, t_sfb,
                                   sfa_desc0 + static_cast<uint64_t>(a_slot * (kSfaStride >> 4)),
                                   sfb_desc0 + static_cast<uint64_t>(stage * (kSfbStride >> 4)), kb != 0 ? 1u : 0u, &empty_bar[stage]);
          if constexpr (ARES) {
            if (n_blk == n1 - 1) umma_commit_warp(&a_empty[kb]);   // last N tile of the unit: A[kb] may be overwritten
          }
          if (++stage == kBStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps --------------------------------------------------------
    const uint32_t quad = warp & 3u;
    const uint32_t ew = warp - 2;                              // 0..11
    const int c_lo = static_cast<int>(ew >> 2) * (32 * kMxChunks);  // this warp's 64 columns of the tile
    uint8_t* my_store = smem_store + ew * kMxChunks * kMxStoreTile;
    const int n_kb_out = N >> 7;                               // k-blocks of the NEXT GEMM (out_mx)
    if constexpr (LNF) {
      // this CTA always owns the same 192 columns (its rank in the cluster): bias / gamma / beta live in shared memory
      const int c0 = static_cast<int>(cluster_ctarank()) * kMxBN;
      for (int i = static_cast<int>(threadIdx.x) - 64; i < kMxBN; i += 32 * kMxEpiWarps) {
        s_cols[i] = ep.bias != nullptr ? __ldg(ep.bias + c0 + i) : 0.f;
        s_cols[kMxBN + i] = __ldg(ep.ln_g + c0 + i);
        s_cols[2 * kMxBN + i] = ep.ln_b != nullptr ? __ldg(ep.ln_b + c0 + i) : 0.f;
      }
      named_bar_sync(1, 32 * kMxEpiWarps);
    }
    uint32_t acc = 0, acc_phase = 0, tile_cnt = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x)
    for (int m_blk = u / n_groups, n_blk = (u % n_groups) * G, n_end = min(num_n, n_blk + G); n_blk < n_end; ++n_blk, ++tile_cnt) {
      const int tile_row0 = m_blk * kMxBM + static_cast<int>(quad * 32u);
      const int col_base = n_blk * kMxBN + c_lo;
      // the staging tiles are free once the previous tile's stores have been read out of smem
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      if (ep.has_res && lane == 0) {
        mbar_expect_tx(&res_bar[ew], kMxChunks * kMxStoreTile);
#pragma unroll
        for (int j = 0; j < kMxChunks; ++j)
          tma_load_2d(my_store + j * kMxStoreTile, &tmap_r, &res_bar[ew], col_base + 32 * j, tile_row0);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // both accumulator chunks in flight before either is touched
      uint32_t v[kMxChunks][32];
#pragma unroll
      for (int j = 0; j < kMxChunks; ++j)
        tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * kMxBN + c_lo + 32 * j, v[j]);
      tmem_ld_wait();
      // TMEM reads done -> the MMA warp may start the tile after next on this accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (ep.has_res) mbar_wait(&res_bar[ew], tile_cnt & 1u);
      if constexpr (LNF) {
        // ---- fused LayerNorm: the cluster's N / 192 CTAs hold one full row block between them ----
        // pass 1: y = acc + bias + residual stays in registers; per-row (sum, sum of squares) of this warp's 64 columns go to
        // every CTA of the cluster with st.async (the 8 bytes complete 8 bytes of transaction count on the receiver's
        // mbarrier when they land: no fence, no separate signal)
        const uint32_t sw = (lane >> 1) & 3u;
        const uint32_t n_cta = cluster_nctarank(), me = cluster_ctarank();
        const uint32_t buf = tile_cnt & 1u;
        if (ew == 0 && lane == 0) mbar_expect_tx(&stat_bar[buf], n_cta * kMxEpiWarps * 32u * 8u);
        uint64_t acc1 = pk2(0.f, 0.f), acc2 = pk2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kMxChunks; ++j) {
          const uint8_t* tile = my_store + j * kMxStoreTile;
          const float* bias_s = s_cols + c_lo + 32 * j;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 8 * q4);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 8 * q4 + 4);
            uint64_t y2[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
            if (ep.has_res) {
              const uint4 rq = *reinterpret_cast<const uint4*>(tile + lane * 64 + ((static_cast<uint32_t>(q4) ^ sw) << 4));
              const float2 a = unpack_bf16x2(rq.x), b = unpack_bf16x2(rq.y), cc = unpack_bf16x2(rq.z), d = unpack_bf16x2(rq.w);
              y2[0] = add2(y2[0], pk2(a.x, a.y));
              y2[1] = add2(y2[1], pk2(b.x, b.y));
              y2[2] = add2(y2[2], pk2(cc.x, cc.y));
              y2[3] = add2(y2[3], pk2(d.x, d.y));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t yv = add2(y2[i], pk2u(v[j][8 * q4 + 2 * i], v[j][8 * q4 + 2 * i + 1]));
              float lo, hi;
              upk2(yv, lo, hi);
              v[j][8 * q4 + 2 * i] = __float_as_uint(lo);
              v[j][8 * q4 + 2 * i + 1] = __float_as_uint(hi);
              acc1 = add2(acc1, yv);
              acc2 = fma2(yv, yv, acc2);
            }
          }
        }
        float s1, s1b, s2, s2b;
        upk2(acc1, s1, s1b);
        upk2(acc2, s2, s2b);
        s1 += s1b;
        s2 += s2b;
        const uint32_t row_in_blk = quad * 32u + lane;
        const uint32_t mine = smem_u32(stats + ((buf * kLnfMaxCtas + me) * kLnfColWarps + (ew >> 2)) * kMxBM + row_in_blk);
        const uint32_t bar = smem_u32(&stat_bar[buf]);
        for (uint32_t p = 0; p < n_cta; ++p) st_async_f32x2(mapa_shared(mine, p), s1, s2, mapa_shared(bar, p));
        mbar_wait(&stat_bar[buf], (tile_cnt >> 1) & 1u);
        float S1 = 0.f, S2 = 0.f;
        for (uint32_t p = 0; p < n_cta; ++p)
#pragma unroll
          for (int cw = 0; cw < kLnfColWarps; ++cw) {
            const float2 t = stats[((buf * kLnfMaxCtas + p) * kLnfColWarps + cw) * kMxBM + row_in_blk];
            S1 += t.x;
            S2 += t.y;
          }
        const float inv_n = 1.0f / static_cast<float>(N);
        const float mean = S1 * inv_n;
        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, S2 * inv_n), 0.f) + ep.ln_eps);
        const uint64_t nmean2 = pk2(-mean, -mean), rstd2 = pk2(rstd, rstd);
        // pass 2: normalise, write the bf16 tile through the staging buffer (TMA store) and the e4m3 copy + its scale bytes directly
        const bool row_ok = tile_row0 + static_cast<int>(lane) < M;
#pragma unroll
        for (int j = 0; j < kMxChunks; ++j) {
          const int col0 = col_base + 32 * j;
          uint8_t* tile = my_store + j * kMxStoreTile;
          const float* g_s = s_cols + kMxBN + c_lo + 32 * j;
          const float* b_s = s_cols + 2 * kMxBN + c_lo + 32 * j;
          float f[32];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 g4 = *reinterpret_cast<const float4*>(g_s + 4 * q4);
            const float4 b4 = *reinterpret_cast<const float4*>(b_s + 4 * q4);
            const uint64_t d0 = mul2(add2(pk2u(v[j][4 * q4 + 0], v[j][4 * q4 + 1]), nmean2), rstd2);
            const uint64_t d1 = mul2(add2(pk2u(v[j][4 * q4 + 2], v[j][4 * q4 + 3]), nmean2), rstd2);
            upk2(fma2(d0, pk2(g4.x, g4.y), pk2(b4.x, b4.y)), f[4 * q4 + 0], f[4 * q4 + 1]);
            upk2(fma2(d1, pk2(g4.z, g4.w), pk2(b4.z, b4.w)), f[4 * q4 + 2], f[4 * q4 + 3]);
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 q;
            q.x = pack_bf16x2(f[8 * q4 + 0], f[8 * q4 + 1]);
            q.y = pack_bf16x2(f[8 * q4 + 2], f[8 * q4 + 3]);
            q.z = pack_bf16x2(f[8 * q4 + 4], f[8 * q4 + 5]);
            q.w = pack_bf16x2(f[8 * q4 + 6], f[8 * q4 + 7]);
            *reinterpret_cast<uint4*>(tile + lane * 64 + ((static_cast<uint32_t>(q4) ^ sw) << 4)) = q;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmap_c, tile, col0, tile_row0);
          float amax = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(f[i]));
          float inv;
          const uint32_t e = ue8m0_from_amax(amax, inv);
          const uint64_t inv2 = pk2(inv, inv);
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a0, a1, a2, a3;
            upk2(mul2(pk2(f[4 * i], f[4 * i + 1]), inv2), a0, a1);
            upk2(mul2(pk2(f[4 * i + 2], f[4 * i + 3]), inv2), a2, a3);
            w8[i] = pack_e4m3x4(a0, a1, a2, a3);
          }
          {
            // a row's 32 e4m3 bytes are one sector: lane pairs swap halves so that EACH store instruction writes whole sectors
            // (even lane: low half of its own row, then low half of the odd lane's row; odd lane: the two high halves)
            const bool odd = (lane & 1u) != 0u;
            uint32_t mine_[4], got[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              mine_[i] = odd ? w8[i] : w8[4 + i];                 // the half this lane gives away
              got[i] = __shfl_xor_sync(0xffffffffu, mine_[i], 1);
            }
            const int r_even = tile_row0 + static_cast<int>(lane & ~1u);
            uint8_t* base = ep.q_out + col0 + (odd ? 16 : 0);
            const uint4 first = odd ? make_uint4(got[0], got[1], got[2], got[3]) : make_uint4(w8[0], w8[1], w8[2], w8[3]);
            const uint4 second = odd ? make_uint4(w8[4], w8[5], w8[6], w8[7]) : make_uint4(got[0], got[1], got[2], got[3]);
            if (r_even < M) *reinterpret_cast<uint4*>(base + static_cast<size_t>(r_even) * ep.ldq) = first;           // row 2i
            if (r_even + 1 < M) *reinterpret_cast<uint4*>(base + static_cast<size_t>(r_even + 1) * ep.ldq) = second;  // row 2i+1
          }
          if (row_ok) {
            ep.c_sf[(static_cast<size_t>(m_blk) * n_kb_out + (col0 >> 7)) * 512 + lane * 16 + quad * 4 + ((col0 & 127) >> 5)] =
                static_cast<uint8_t>(e);
          }
        }
      } else {
#pragma unroll
      for (int j = 0; j < kMxChunks; ++j) {
        const int col0 = col_base + 32 * j;
        if (col0 >= N) continue;
        float f[32];
        if (ep.bias != nullptr) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q4);
            f[4 * q4 + 0] = __uint_as_float(v[j][4 * q4 + 0]) + b4.x;
            f[4 * q4 + 1] = __uint_as_float(v[j][4 * q4 + 1]) + b4.y;
            f[4 * q4 + 2] = __uint_as_float(v[j][4 * q4 + 2]) + b4.z;
            f[4 * q4 + 3] = __uint_as_float(v[j][4 * q4 + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[j][i]);
        }
        mx_act32(f, ep.act);
        uint8_t* tile = my_store + j * kMxStoreTile;
        if (ep.out_mx) {
          float amax = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(f[i]));
          float inv;
          const uint32_t e = ue8m0_from_amax(amax, inv);
          const uint64_t inv2 = pk2(inv, inv);
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a0, a1, a2, a3;
            upk2(mul2(pk2(f[4 * i], f[4 * i + 1]), inv2), a0, a1);
            upk2(mul2(pk2(f[4 * i + 2], f[4 * i + 3]), inv2), a2, a3);
            w8[i] = pack_e4m3x4(a0, a1, a2, a3);
          }
          const uint4 q0 = make_uint4(w8[0], w8[1], w8[2], w8[3]), q1 = make_uint4(w8[4], w8[5], w8[6], w8[7]);
          // 32-byte rows: swap the two 16-byte halves on every other group of four lanes so a quarter-warp's stores spread
          // over all eight 16-byte bank groups ... the TMA store below expects plain rows, so the swap is undone by
          // construction: (half ^ x) selects the slot, x ? the other vector : this one selects the data
          const uint32_t x = (lane >> 2) & 1u;
          *reinterpret_cast<uint4*>(tile + lane * 32 + ((0u ^ x) << 4)) = x ? q1 : q0;
          *reinterpret_cast<uint4*>(tile + lane * 32 + ((1u ^ x) << 4)) = x ? q0 : q1;
          // scale byte of (row, 32-column block) in the consumer GEMM's SFA chunk layout
          if (tile_row0 + static_cast<int>(lane) < M) {
            uint8_t* sf = ep.c_sf + (static_cast<size_t>(m_blk) * n_kb_out + (col0 >> 7)) * 512 + lane * 16 + quad * 4 +
                          ((col0 & 127) >> 5);
            *sf = static_cast<uint8_t>(e);
          }
        } else {
          // bf16 tile: 64-byte rows, TMA SWIZZLE_64B (16-byte chunk index ^= (row >> 1) & 3)
          const uint32_t sw = (lane >> 1) & 3u;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4* slot = reinterpret_cast<uint4*>(tile + lane * 64 + ((static_cast<uint32_t>(q4) ^ sw) << 4));
            if (ep.has_res) {
              const uint4 rq = *slot;
              const float2 a = unpack_bf16x2(rq.x), b = unpack_bf16x2(rq.y), cc = unpack_bf16x2(rq.z), d = unpack_bf16x2(rq.w);
              f[8 * q4 + 0] += a.x; f[8 * q4 + 1] += a.y; f[8 * q4 + 2] += b.x; f[8 * q4 + 3] += b.y;
              f[8 * q4 + 4] += cc.x; f[8 * q4 + 5] += cc.y; f[8 * q4 + 6] += d.x; f[8 * q4 + 7] += d.y;
            }
            uint4 q;
            q.x = pack_bf16x2(f[8 * q4 + 0], f[8 * q4 + 1]);
            q.y = pack_bf16x2(f[8 * q4 + 2], f[8 * q4 + 3]);
            q.z = pack_bf16x2(f[8 * q4 + 4], f[8 * q4 + 5]);
            q.w = pack_bf16x2(f[8 * q4 + 6], f[8 * q4 + 7]);
            *slot = q;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&tmap_c, tile, col0, tile_row0);
      }
      }
      if (lane == 0) tma_store_commit();   // one bulk group per tile
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (LNF) cluster_sync_all();     // a CTA's statistics slots and barriers stay valid until every peer is done
  if (warp == 2) tmem_dealloc(tmem_base, kMxTmemCols);
}

}  // namespace im

// C = act(A·B^T + bias) (+ residual) with MXFP8 operands; out_mx selects bf16 or MXFP8 output (see file header).
IM_API int im_gemm_mxf8(const void* A, const void* SFA, const void* B, const void* SFB, int n_chunks_b, void* C, void* C_sf,
                        const float* bias, const void* residual, int M, int N, int K, int lda, int ldb, int ldc, int ldr,
                        int act, int out_mx, const int* m_dev, int max_ctas, void* stream, int mode) {
  // mode: 0 = pick the variant (A-resident when K <= 768 and it pays), 1 = always the operand-ring kernel, 2 = force the
  // A-resident kernel (tests / A-B timing).  Returns < 0 on error, else the N-tile group size used (0 = ring kernel).
  using namespace im;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % kMxBK) return set_error("im_gemm_mxf8", "K must be a multiple of 128");
  if (N % 32) return set_error("im_gemm_mxf8", "N must be a multiple of 32");
  if ((lda % 16) || (ldb % 16)) return set_error("im_gemm_mxf8", "operand row pitches must be multiples of 16 bytes");
  if (out_mx && (N % 128)) return set_error("im_gemm_mxf8", "MXFP8 output needs N % 128 == 0");
  if (out_mx && residual != nullptr) return set_error("im_gemm_mxf8", "residual is only fused into the bf16 epilogue");
  if (n_chunks_b < ((N + kMxBN - 1) / kMxBN * kMxBN + 127) / 128)
    return set_error("im_gemm_mxf8", "SFB needs ceil(round_up(N,192)/128) chunks per k-block");
  static bool configured = false;
  if (!configured) {
    IM_CUDA_OK(cudaFuncSetAttribute(gemm_mxf8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmemBytes));
    IM_CUDA_OK(cudaFuncSetAttribute(gemm_mxf8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAresSmemBytes));
    configured = true;
  }
  CUtensorMap ta, tb, tc, tr;
  if (get_tmap_2d(&ta, A, M, K, static_cast<uint64_t>(lda), kMxBM, kMxBK, 1, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&tb, B, N, K, static_cast<uint64_t>(ldb), kMxBN, kMxBK, 1, TMAP_SW_128)) return -1;
  if (out_mx) {
    if (get_tmap_2d(&tc, C, M, N, static_cast<uint64_t>(ldc), 32, 32, 1, TMAP_SW_NONE)) return -1;
  } else {
    if (get_tmap_2d(&tc, C, M, N, static_cast<uint64_t>(ldc) * 2, 32, 32, 2, TMAP_SW_64)) return -1;
  }
  tr = tc;
  if (residual != nullptr && get_tmap_2d(&tr, residual, M, N, static_cast<uint64_t>(ldr) * 2, 32, 32, 2, TMAP_SW_64)) return -1;
  MxEpilogue ep;
  ep.bias = bias;
  ep.c_sf = reinterpret_cast<uint8_t*>(C_sf);
  ep.sfa = reinterpret_cast<const uint8_t*>(SFA);
  ep.sfb = reinterpret_cast<const uint8_t*>(SFB);
  ep.n_chunks_b = n_chunks_b;
  ep.act = act;
  ep.out_mx = out_mx;
  ep.has_res = residual != nullptr ? 1 : 0;
  ep.m_dev = m_dev;
  ep.ln_g = nullptr;
  ep.ln_b = nullptr;
  ep.ln_eps = 0.f;
  ep.q_out = nullptr;
  ep.ldq = 0;
  const int num_m = (M + kMxBM - 1) / kMxBM, num_n = (N + kMxBN - 1) / kMxBN;
  // A-resident variant when the activation block fits (K <= 768) and there is enough parallelism left after grouping N
  // tiles: halve the group until the units cover the SMs at least twice (G = 1 degenerates to the ring kernel's order)
  int G = 1;
  // Measured on B200 (profiles/mx_gemm_check_r2c.log): the kernel is bound by shared-memory bandwidth (TMA fill + UMMA
  // operand reads + epilogue staging ~ 576 KB per 128x192x768 tile against 128 B/clk), and with only 3 B stages left the
  // A-resident variant loses more to L2 latency than it saves in fill traffic (QKV 200 vs 174 us, FFN-up 230 vs 242 us).
  // It therefore runs only on request (mode 2); the operand ring is the default.
  bool ares = (K / kMxBK) <= kAresMaxKb && num_n > 1 && mode == 2;
  if (ares) {
    G = num_n;
    while (G > 1 && static_cast<long long>(num_m) * ((num_n + G - 1) / G) < 2LL * sm_count()) G = (G + 1) / 2;
    if (G <= 1 && mode != 2) ares = false;
    if (G < 1) G = 1;
  }
  ep.n_per_unit = G;
  const long long units = static_cast<long long>(num_m) * (ares ? (num_n + G - 1) / G : num_n);
  int grid = units < sm_count() ? static_cast<int>(units) : sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ares)
    IM_CUDA_OK(launch_pdl(gemm_mxf8_kernel<true>, dim3(grid), dim3(kMxThreads), kAresSmemBytes, st, ta, tb, tc, tr, ep, M, N, K));
  else
    IM_CUDA_OK(launch_pdl(gemm_mxf8_kernel<false>, dim3(grid), dim3(kMxThreads), kMxSmemBytes, st, ta, tb, tc, tr, ep, M, N, K));
  IM_LAUNCH_OK("gemm_mxf8_kernel");
  return ares ? G : 0;
}

// C = LayerNorm(A·B^T + bias + residual) in bf16, plus the MXFP8 copy of C (Cq + Cq_sf: the next block-scaled GEMM's A
// operand), in ONE kernel: N / 192 CTAs (2 or 4) form a thread-block cluster that owns a 128-row block, exchange per-row
// (sum, sum of squares) through distributed shared memory and normalise their own 192 columns.  Replaces the
// GEMM -> sum_ln_mx pair of the transformer's attention-output and FFN-down projections (one HBM round trip of the
// [M, N] activations and one launch less per projection).
IM_API int im_gemm_mxf8_ln(const void* A, const void* SFA, const void* B, const void* SFB, int n_chunks_b, void* C, void* Cq,
                           void* Cq_sf, const float* bias, const void* residual, const float* gamma, const float* beta, float eps,
                           int M, int N, int K, int lda, int ldb, int ldc, int ldr, int ldq, const int* m_dev, int max_ctas,
                           void* stream) {
  using namespace im;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % kMxBK) return set_error("im_gemm_mxf8_ln", "K must be a multiple of 128");
  if (N % kMxBN || N / kMxBN < 2 || N / kMxBN > kLnfMaxCtas || (N / kMxBN) & (N / kMxBN - 1))
    return set_error("im_gemm_mxf8_ln", "N must be 384 or 768 (a cluster of 2 or 4 CTAs owns a row block)");
  if ((lda % 16) || (ldb % 16) || (ldq % 16)) return set_error("im_gemm_mxf8_ln", "row pitches must be multiples of 16 bytes");
  if (gamma == nullptr || Cq == nullptr || Cq_sf == nullptr) return set_error("im_gemm_mxf8_ln", "gamma and the MXFP8 output are required");
  if (n_chunks_b < (N + 127) / 128) return set_error("im_gemm_mxf8_ln", "SFB needs ceil(N/128) chunks per k-block");
  const int n_cta = N / kMxBN;
  static bool configured = false;
  if (!configured) {
    IM_CUDA_OK(cudaFuncSetAttribute(gemm_mxf8_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnfSmemBytes));
    configured = true;
  }
  CUtensorMap ta, tb, tc, tr;
  if (get_tmap_2d(&ta, A, M, K, static_cast<uint64_t>(lda), kMxBM, kMxBK, 1, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&tb, B, N, K, static_cast<uint64_t>(ldb), kMxBN, kMxBK, 1, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&tc, C, M, N, static_cast<uint64_t>(ldc) * 2, 32, 32, 2, TMAP_SW_64)) return -1;
  tr = tc;
  if (residual != nullptr && get_tmap_2d(&tr, residual, M, N, static_cast<uint64_t>(ldr) * 2, 32, 32, 2, TMAP_SW_64)) return -1;
  MxEpilogue ep;
  ep.bias = bias;
  ep.c_sf = reinterpret_cast<uint8_t*>(Cq_sf);
  ep.sfa = reinterpret_cast<const uint8_t*>(SFA);
  ep.sfb = reinterpret_cast<const uint8_t*>(SFB);
  ep.n_chunks_b = n_chunks_b;
  ep.act = 0;
  ep.out_mx = 0;
  ep.has_res = residual != nullptr ? 1 : 0;
  ep.m_dev = m_dev;
  ep.n_per_unit = 1;
  ep.ln_g = gamma;
  ep.ln_b = beta;
  ep.ln_eps = eps;
  ep.q_out = reinterpret_cast<uint8_t*>(Cq);
  ep.ldq = ldq;
  const int num_m = (M + kMxBM - 1) / kMxBM;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // one cluster per row block in flight; the grid is a whole number of clusters, at most as many as can be co-resident
  static int max_clusters[kLnfMaxCtas + 1] = {0, 0, 0, 0, 0};
  if (max_clusters[n_cta] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / n_cta * n_cta));
    cfg.blockDim = dim3(kMxThreads);
    cfg.dynamicSmemBytes = kLnfSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(n_cta);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_mxf8_kernel<false, true>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = sm_count() / n_cta;
    }
    max_clusters[n_cta] = n;
  }
  int clusters = num_m < max_clusters[n_cta] ? num_m : max_clusters[n_cta];
  if (max_ctas > 0 && clusters * n_cta > max_ctas) clusters = max_ctas / n_cta;
  if (clusters < 1) clusters = 1;
  IM_CUDA_OK(launch_pdl_cluster(gemm_mxf8_kernel<false, true>, dim3(static_cast<unsigned>(clusters * n_cta)), dim3(kMxThreads),
                                kLnfSmemBytes, st, static_cast<unsigned>(n_cta), ta, tb, tc, tr, ep, M, N, K));
  IM_LAUNCH_OK("gemm_mxf8_kernel<ln>");
  return clusters;
}
