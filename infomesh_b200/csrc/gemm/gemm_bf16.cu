// K6/K8 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] · B[N,K]^T )        A, B bf16 row-major (K contiguous, "TN")
//
// Replaces the encoder / reranker / summariser linear layers that the reference reaches through
// sentence-transformers / an external LLM server (reference infomesh/index/vector_store.py:104-125,
// infomesh/summarizer/engine.py:126-141).  Structure (one CTA per SM, 10 warps):
//   warp 0      TMA producer      cp.async.bulk.tensor 128B-swizzled A/B tiles -> smem ring (mbarrier tx)
//   warp 1      MMA issuer        one thread issues tcgen05.mma.kind::f16 128xBNx16, accumulators in TMEM
//   warps 2..9  epilogue          tcgen05.ld 32 lanes x 32 cols -> bias / activation / residual -> bf16 stores
// The TMEM accumulator is double buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Fused-collective hooks (used by parallel/tp.py):
//   * a_ready flags: the producer acquires a per-row-block flag before loading A (all-gather -> GEMM:
//     comm CTAs / peers publish row blocks as they land),
//   * peer_c: the epilogue pushes each output row block straight into the owning rank's receive slot over
//     NVLink (GEMM -> reduce-scatter) and bumps a per-row-block arrival counter there.
#include "../common/host.h"
#include "../common/ptx.cuh"
#include "../common/tmap_cache.h"

namespace im {

struct GemmEpilogue {
  void* c;                  // output (bf16 or fp32), row pitch ldc elements
  const float* bias;        // [N] or null
  const __nv_bfloat16* residual;  // [M, ldr] or null
  int ldc, ldr;
  int act;                  // 0 none, 1 gelu(erf), 2 relu, 3 gelu(tanh), 4 tanh
  int tma_store;            // bf16 output staged through swizzled smem and written by TMA (coalesced)
  int out_fp32;
  float alpha;
  // --- fused reduce-scatter push (null => local store) ---
  void* const* peer_c;      // [tp] receive-buffer base of every rank: [tp_src][rows_per_rank][ldc]
  uint32_t* const* peer_flags;  // [tp] arrival counters of every rank: [tp_src][rows_per_rank/128]
  int rank, rows_per_rank;
  // --- fused all-gather wait (null => no wait) ---
  const uint32_t* a_ready;  // [ceil(M/128)] cumulative per-row arrival counters (one bump per landed row)
  uint32_t* a_state;        // {use, done}: block m is loadable once a_ready[m] >= (use + 1) * rows_in_block(m);
                            // the last CTA out advances `use` (see comm/symm.cu for the protocol)
  int m_rotate;             // first row block processed (so the local shard goes first)
  int fp8;                  // A and B are e4m3 bytes (kind::f8f6f4, 128 elements per K block); C / bias / residual as usual
  const float* row_scale;   // optional per-row (per-token) dequantisation scale multiplied into alpha
  const int* m_dev;         // optional device-side row count (unpadded / varlen batches inside a CUDA graph): rows
                            // beyond min(M, *m_dev) are neither loaded nor computed
};

struct PeerMaps {
  CUtensorMap m[8];  // store maps over every rank's reduce-scatter receive area [tp * rows_per_rank, N]
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;   // two per TMEM lane quadrant: each takes half of the tile's columns, and the pair hides
                               // each other's TMEM-load / MUFU / bias-load latency on the shared scheduler
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = (kBM + BN) * kBK * 2;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;  // 256 or 512: both powers of two
  static constexpr int kStoreBytes = kEpiWarps * 4096;  // one [32 rows x 64 bf16] staging tile per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Activation over a 32-value register chunk.  The switch sits OUTSIDE the element loops on purpose: a per-element
// if-chain gets if-converted by ptxas and then evaluates every activation for every element.
__device__ __forceinline__ void apply_act32(float (&f)[32], int act) {
  switch (act) {
    case 1:
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
      break;
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

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r,
                    const __grid_constant__ PeerMaps tmap_peers, const GemmEpilogue ep, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_store = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-aligned (stage bytes are multiples of 1 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_store + Cfg::kStoreBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;  // [8 epilogue warps] residual-tile arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  pdl_trigger();  // the next kernel may start its prologue now; it cannot touch our data before we complete

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();  // ---- everything above overlapped the previous kernel's tail; from here on we read its output ----
  if (ep.m_dev != nullptr) M = min(M, max(0, *ep.m_dev));
  const int num_m = (M + kBM - 1) / kBM;
  const int num_n = (N + BN - 1) / BN;
  const int kbk = ep.fp8 ? 2 * kBK : kBK;  // elements per 128-byte K block
  const int num_k = (K + kbk - 1) / kbk;
  const int num_tiles = num_m * num_n;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t a_use = ep.a_state != nullptr ? *reinterpret_cast<volatile uint32_t*>(ep.a_state) : 0u;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk = t / num_n, n_blk = t % num_n;
        m_blk = (m_blk + ep.m_rotate) % num_m;
        if (ep.a_ready != nullptr) {
          const uint32_t target = (a_use + 1u) * static_cast<uint32_t>(min(kBM, M - m_blk * kBM));
          uint32_t spins = 0;
          while (static_cast<int32_t>(ld_acquire_sys(ep.a_ready + m_blk) - target) < 0) {
            if (++spins > IM_WAIT_LIMIT) {
              printf("[infomesh_b200] gemm a_ready timeout m_blk=%d\n", m_blk);
              __trap();
            }
            __nanosleep(20);
          }
        }
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kBM * kBK * 2;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kbk, m_blk * kBM);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kbk, n_blk * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // The whole warp runs the loop (warp-uniform control flow); one elected lane issues (see umma_bf16_kblock64_warp).
    {
      const uint32_t idesc = ep.fp8 ? umma_idesc_f8(kBM, BN) : umma_idesc_f16(kBM, BN);
      static_assert(kBK == 64, "umma_bf16_kblock64_warp issues exactly one 64-wide K block");
      // stage s operands live at smem + s * kStageBytes (A) / + kBM*kBK*2 (B): descriptors differ only in the
      // address field (16-byte units), so they are built once and stepped with integer adds
      const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem));
      const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem) + kBM * kBK * 2);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t soff = static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
          if (ep.fp8)
            umma_f8_kblock128_warp(d_tmem, a_desc0 + soff, b_desc0 + soff, idesc, kb != 0 ? 1u : 0u, &empty_bar[stage]);
          else
            umma_bf16_kblock64_warp(d_tmem, a_desc0 + soff, b_desc0 + soff, idesc, kb != 0 ? 1u : 0u, &empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_warp(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps ------------------------------
    const uint32_t quad = warp & 3u;  // TMEM lane quadrant this warp may touch
    const uint32_t ew = warp - 2;     // 0..7
    const int c_lo = static_cast<int>(ew >> 2) * (BN / 2), c_hi = c_lo + BN / 2;  // this warp's half of the tile columns
    const uint32_t row_in_tile = quad * 32u + lane;
    uint32_t acc = 0, acc_phase = 0, store_cnt = 0;
    constexpr int kChunks = BN / 128;     // bulk-store groups this warp commits per tile
    uint32_t* pending_flag = nullptr;     // fused reduce-scatter: arrival counter of the previous tile (lane 0)
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk = t / num_n, n_blk = t % num_n;
      m_blk = (m_blk + ep.m_rotate) % num_m;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * kBM + static_cast<int>(row_in_tile);
      const bool row_ok = row < M;
      const float alpha_row = (ep.row_scale != nullptr && row_ok) ? ep.alpha * ep.row_scale[row] : ep.alpha;
      // destination row pointer (local C, or the owner's receive slot for fused reduce-scatter)
      uint8_t* c_row = nullptr;
      int owner = 0;
      if (ep.peer_c != nullptr) {
        owner = (m_blk * kBM) / ep.rows_per_rank;
        const int local_row = row - owner * ep.rows_per_rank;
        uint8_t* base = reinterpret_cast<uint8_t*>(ep.peer_c[owner]);
        c_row = base + (static_cast<size_t>(ep.rank) * ep.rows_per_rank + local_row) * ep.ldc * (ep.out_fp32 ? 4 : 2);
      } else {
        c_row = reinterpret_cast<uint8_t*>(ep.c) + static_cast<size_t>(row) * ep.ldc * (ep.out_fp32 ? 4 : 2);
      }
      const __nv_bfloat16* r_row = ep.residual ? ep.residual + static_cast<size_t>(row) * ep.ldr : nullptr;
      if (ep.tma_store) {
        // ---- coalesced path: TMEM -> regs -> 128B-swizzled smem tile [32 rows x 64 cols] -> TMA store ----
        const int tile_row0 = m_blk * kBM + static_cast<int>(quad * 32u);
        // fused reduce-scatter: same staged tiles, but the TMA store targets the owner's receive slot over NVLink
        const CUtensorMap* store_map = ep.peer_c != nullptr ? &tmap_peers.m[owner] : &tmap_c;
        const int store_row0 = ep.peer_c != nullptr ? (ep.rank - owner) * ep.rows_per_rank + tile_row0 : tile_row0;
        int n_commit = 0;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 64) {
          const int col0 = n_blk * BN + c;
          if (col0 >= N) break;
          ++n_commit;
          uint8_t* stage = smem_store + ew * 4096;
          if (lane == 0) tma_store_wait_read<0>();  // the previous store from this buffer has been read out
          __syncwarp();
          if (ep.residual != nullptr) {
            if (lane == 0) {
              mbar_expect_tx(&res_bar[ew], 4096);
              tma_load_2d(stage, &tmap_r, &res_bar[ew], col0, tile_row0);
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * BN + c + half * 32, v);
            tmem_ld_wait();
            const int cb = col0 + half * 32;
            uint64_t g[16];  // 32 fp32 values as 16 packed pairs (FFMA2 path)
            if (ep.bias != nullptr && cb + 32 <= N) {
              // common case: one FFMA2 per element PAIR (alpha * acc + bias)
              const uint64_t al2 = pk2(alpha_row, alpha_row);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + cb) + j);
                g[2 * j] = fma2(pk2u(v[4 * j], v[4 * j + 1]), al2, pk2(b4.x, b4.y));
                g[2 * j + 1] = fma2(pk2u(v[4 * j + 2], v[4 * j + 3]), al2, pk2(b4.z, b4.w));
              }
            } else {
              float f[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * alpha_row;
              if (ep.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (cb + i < N) f[i] += __ldg(ep.bias + cb + i);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] = pk2(f[2 * i], f[2 * i + 1]);
            }
            if (ep.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] = gelu_erf2(g[i]);
            } else if (ep.act != 0) {
              float f[32];
#pragma unroll
              for (int i = 0; i < 16; ++i) upk2(g[i], f[2 * i], f[2 * i + 1]);
              apply_act32(f, ep.act);
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] = pk2(f[2 * i], f[2 * i + 1]);
            }
            if (ep.residual != nullptr && half == 0) mbar_wait(&res_bar[ew], store_cnt & 1u);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int ch = (half * 4 + q4) ^ static_cast<int>(lane & 7u);
              uint4* slot = reinterpret_cast<uint4*>(stage + lane * 128 + (ch << 4));
              if (ep.residual != nullptr) {
                const uint4 rq = *slot;
                const float2 a = unpack_bf16x2(rq.x), b = unpack_bf16x2(rq.y), cc = unpack_bf16x2(rq.z), d = unpack_bf16x2(rq.w);
                g[q4 * 4 + 0] = add2(g[q4 * 4 + 0], pk2(a.x, a.y));
                g[q4 * 4 + 1] = add2(g[q4 * 4 + 1], pk2(b.x, b.y));
                g[q4 * 4 + 2] = add2(g[q4 * 4 + 2], pk2(cc.x, cc.y));
                g[q4 * 4 + 3] = add2(g[q4 * 4 + 3], pk2(d.x, d.y));
              }
              uint4 q;
              q.x = pack_bf16x2_pair(g[q4 * 4 + 0]);
              q.y = pack_bf16x2_pair(g[q4 * 4 + 1]);
              q.z = pack_bf16x2_pair(g[q4 * 4 + 2]);
              q.w = pack_bf16x2_pair(g[q4 * 4 + 3]);
              *slot = q;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(store_map, stage, col0, store_row0);
            tma_store_commit();
          }
          ++store_cnt;
        }
        // always kChunks groups per tile so wait_group<kChunks> below means "the previous tile has landed"
        if (ep.peer_c != nullptr && lane == 0)
          for (; n_commit < kChunks; ++n_commit) tma_store_commit();
      } else
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((quad * 32u) << 16) + acc * BN + c, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c;
        if (!row_ok || col0 >= N) continue;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * alpha_row;
        const bool full_chunk = (col0 + 32 <= N);
        if (ep.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full_chunk || col0 + i < N) f[i] += __ldg(ep.bias + col0 + i);
        }
        apply_act32(f, ep.act);
        if (r_row != nullptr) {
          if (full_chunk) {
            const uint4* rp = reinterpret_cast<const uint4*>(r_row + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q = __ldg(rp + j);
              float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), cc = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
              f[j * 8 + 0] += a.x; f[j * 8 + 1] += a.y; f[j * 8 + 2] += b.x; f[j * 8 + 3] += b.y;
              f[j * 8 + 4] += cc.x; f[j * 8 + 5] += cc.y; f[j * 8 + 6] += d.x; f[j * 8 + 7] += d.y;
            }
          } else {
            for (int i = 0; i < 32; ++i)
              if (col0 + i < N) f[i] += __bfloat162float(r_row[col0 + i]);
          }
        }
        if (ep.out_fp32) {
          float* out = reinterpret_cast<float*>(c_row) + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(out)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            for (int i = 0; i < 32; ++i)
              if (col0 + i < N) out[i] = f[i];
          }
        } else {
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(c_row) + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q;
              q.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              q.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              q.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              q.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              reinterpret_cast<uint4*>(out)[j] = q;
            }
          } else {
            for (int i = 0; i < 32; ++i)
              if (col0 + i < N) out[i] = __float2bfloat16(f[i]);
          }
        }
      }
      // TMEM reads done -> hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty[acc]);
        if (ep.peer_c != nullptr) {
          const int blk_local = (m_blk * kBM - owner * ep.rows_per_rank) / kBM;
          uint32_t* flag = ep.peer_flags[owner] + static_cast<size_t>(ep.rank) * (ep.rows_per_rank / kBM) + blk_local;
          if (ep.tma_store) {
            // deferred arrival: the counter of tile t-1 is bumped once its bulk stores have COMPLETED (not merely been
            // read out of smem), which by now costs nothing -- tile t's stores are the only groups still in flight.
            // red.release.sys orders the completed async-proxy writes before the bump for the acquiring consumer.
            if (pending_flag != nullptr) {
              tma_store_wait_all<kChunks>();
              asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(pending_flag) : "memory");
            }
            pending_flag = flag;
          } else {
            asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (ep.tma_store && lane == 0) {
      tma_store_wait_all<0>();
      if (pending_flag != nullptr)
        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(pending_flag) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (ep.a_state != nullptr && threadIdx.x == 0) {  // last CTA out advances the all-gather channel
    __threadfence();
    if (atomicAdd(ep.a_state + 1, 1u) == gridDim.x - 1u) {
      ep.a_state[1] = 0u;
      ep.a_state[0] += 1u;
      __threadfence();
    }
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                       const PeerMaps& tp, const GemmEpilogue& ep, int M, int N, int K, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    IM_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  IM_CUDA_OK(launch_pdl(gemm_bf16_tn_kernel<BN>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, tc, tr,
                        tp, ep, M, N, K));
  IM_LAUNCH_OK("gemm_bf16_tn_kernel");
  return 0;
}

}  // namespace im

// C[M,N] = act(alpha * A[M,K] B[N,K]^T + bias) + residual.  bn = 0 picks the tile width.
IM_API int im_gemm_bf16_tn(const void* A, const void* B, void* C, const float* bias, const void* residual, int M, int N,
                           int K, int lda, int ldb, int ldc, int ldr, int act, int out_fp32, float alpha, int bn,
                           void* const* peer_c, uint32_t* const* peer_flags, int rank, int rows_per_rank,
                           const uint32_t* a_ready, uint32_t* a_state, int m_rotate, int max_ctas, void* stream,
                           const void* const* peer_c_host, int world, const int* m_dev, int fp8,
                           const float* row_scale) {
  using namespace im;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (residual != nullptr && (ldr % 8))) return set_error("im_gemm_bf16_tn", "leading dims must be multiples of 8");
  if (peer_c != nullptr && (rows_per_rank % kBM) != 0)
    return set_error("im_gemm_bf16_tn", "rows_per_rank must be a multiple of 128 for the fused reduce-scatter");
  if (bn == 0) {
    const int tiles256 = ((M + kBM - 1) / kBM) * ((N + 255) / 256);
    bn = (N % 256 == 0 && tiles256 >= sm_count()) ? 256 : 128;
  }
  CUtensorMap ta, tb;
  const int eb = fp8 ? 1 : 2;                 // operand element bytes; a K block is always 128 bytes wide
  if (fp8 && ((lda % 16) || (ldb % 16))) return set_error("im_gemm_bf16_tn", "fp8 operands: leading dims must be multiples of 16");
  if (get_tmap_2d(&ta, A, M, K, static_cast<uint64_t>(lda) * eb, kBM, 128 / eb, eb, TMAP_SW_128)) return -1;
  if (get_tmap_2d(&tb, B, N, K, static_cast<uint64_t>(ldb) * eb, bn, 128 / eb, eb, TMAP_SW_128)) return -1;
  GemmEpilogue ep;
  ep.c = C;
  ep.bias = bias;
  ep.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  ep.ldc = ldc;
  ep.ldr = ldr;
  ep.act = act;
  ep.out_fp32 = out_fp32;
  ep.alpha = alpha;
  ep.peer_c = peer_c;
  ep.peer_flags = peer_flags;
  ep.rank = rank;
  ep.rows_per_rank = rows_per_rank > 0 ? rows_per_rank : M;
  ep.a_ready = a_ready;
  ep.a_state = a_state;
  const int num_m = (M + kBM - 1) / kBM;
  ep.m_rotate = num_m > 0 ? ((m_rotate % num_m) + num_m) % num_m : 0;
  ep.m_dev = m_dev;
  ep.fp8 = fp8;
  ep.row_scale = row_scale;
  // coalesced TMA-store epilogue for plain local bf16 outputs
  CUtensorMap tc = ta, tr = ta;
  PeerMaps tp;
  for (auto& m : tp.m) m = ta;
  const bool peer_tma = peer_c != nullptr && peer_c_host != nullptr && !out_fp32 && residual == nullptr && world >= 1 && world <= 8;
  if (peer_tma) {
    // one store map per owner: its receive area seen as [tp * rows_per_rank, N] with row pitch ldc
    for (int p = 0; p < world; ++p)
      if (get_tmap_2d(&tp.m[p], peer_c_host[p], static_cast<uint64_t>(world) * rows_per_rank, N, static_cast<uint64_t>(ldc) * 2, 32,
                      64, 2, TMAP_SW_128))
        return -1;
  }
  ep.tma_store = (!out_fp32 && (peer_c == nullptr || peer_tma)) ? 1 : 0;
  if (ep.tma_store && peer_c == nullptr) {
    if (get_tmap_2d(&tc, C, M, N, static_cast<uint64_t>(ldc) * 2, 32, 64, 2, TMAP_SW_128)) return -1;
    if (residual != nullptr &&
        get_tmap_2d(&tr, residual, M, N, static_cast<uint64_t>(ldr) * 2, 32, 64, 2, TMAP_SW_128))
      return -1;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 256) return launch_gemm<256>(ta, tb, tc, tr, tp, ep, M, N, K, max_ctas, s);
  return launch_gemm<128>(ta, tb, tc, tr, tp, ep, M, N, K, max_ctas, s);
}
