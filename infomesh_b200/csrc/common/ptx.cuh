// infomesh_b200 — sm_100a PTX wrappers shared by every tensor-core kernel.
//
// Everything here is Blackwell-only: mbarrier/TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld / fences) and the UMMA shared-memory + instruction
// descriptors.  Bit layouts follow the PTX ISA "tcgen05" chapter (the same fields
// CUTLASS' cute/arch/mma_sm100_desc.hpp names SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace im {

// ---------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

#ifndef IM_WAIT_LIMIT
// Bounded spin: a lost arrive traps (kernel error) instead of hanging the GPU.
#define IM_WAIT_LIMIT (1u << 24)
#endif

// ---------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > IM_WAIT_LIMIT) {
      printf("[infomesh_b200] mbarrier wait timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------
// Thread-block clusters: distributed shared memory + remote mbarrier arrivals (gemm_mxf8.cu, fused LayerNorm epilogue)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// all threads of every CTA in the cluster (kernel start: barriers initialised before any peer arrives; kernel end: nobody
// exits while a peer may still touch its shared memory)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// 8 bytes into a peer CTA's shared memory, completing 8 bytes of transaction count on THAT CTA's mbarrier when they land:
// data and signal travel together, no fence on the sender (the receiver set the expected byte count with expect_tx)
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr),
               "f"(a), "f"(b), "r"(cluster_bar_addr)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
// one arrival on a (possibly remote) mbarrier; release at cluster scope publishes this thread's -- and, after a __syncwarp,
// its warp's -- earlier st.shared::cluster writes to whoever acquires the phase
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > IM_WAIT_LIMIT) {
      printf("[infomesh_b200] cluster mbarrier wait timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------
// Programmatic dependent launch: pdl_trigger() lets the NEXT kernel in the stream start its prologue while this one
// still runs; pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are visible.  Both are no-ops
// when the kernel was launched without the programmatic-serialization attribute (host.h launch_pdl).
// Rule: nothing produced by an earlier kernel may be read, and nothing global may be written, before pdl_wait().
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — tensor maps are built on the host (tmap.h)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tile load: c0 = innermost (contiguous) coordinate, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Same with an L2 cache-policy hint (createpolicy value).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// 2-D tile store smem -> global (bulk async group); rows/cols outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, loads, fences
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One 64-wide K block (4 x K=16 bf16 MMAs) + the commit that frees its smem stage, as ONE asm statement, called by
// ALL 32 lanes of the (converged) MMA warp; `elect.sync` inside picks the issuing lane.
//
// Why not `if (lane == 0) { tcgen05.mma ... }`: in divergent code ptxas must assume any subset of lanes is active, so
// it wraps every tcgen05 instruction in an ELECT / vote / BRA.U.ANY serialisation loop and re-materialises the
// uniform-register descriptors each time (~25 SASS instructions per MMA; the issue loop then takes as long as the
// MMAs themselves and the tensor pipe idles ~45 %).  With warp-uniform control flow and an elect.sync predicate the
// MMAs compile to back-to-back `@UP UTCHMMA`.  The descriptors' 14-bit address field is stepped in place
// (+2 = 32 bytes per K=16 step inside the 128B swizzle atom).
__device__ __forceinline__ void umma_bf16_kblock64_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                        uint32_t accumulate_first, uint64_t* commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, pe;\n"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 q, %4, %4;\n"
      "add.s64 a1, %1, 2;\n"
      "add.s64 b1, %2, 2;\n"
      "add.s64 a2, %1, 4;\n"
      "add.s64 b2, %2, 4;\n"
      "add.s64 a3, %1, 6;\n"
      "add.s64 b3, %2, 6;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, q;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_u32(commit_bar))
      : "memory");
}
// fp8 (e4m3 / e5m2) flavour of umma_bf16_kblock64_warp: one 128-BYTE K block = 4 x (K = 32 fp8) kind::f8f6f4 MMAs.
// Same smem tiles (128-byte swizzled rows) and the same +32-byte descriptor step as the bf16 version -- only the
// instruction kind, the instruction descriptor and the number of elements per row differ.
__device__ __forceinline__ void umma_f8_kblock128_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate_first, uint64_t* commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, pe;\n"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 q, %4, %4;\n"
      "add.s64 a1, %1, 2;\n"
      "add.s64 b1, %2, 2;\n"
      "add.s64 a2, %1, 4;\n"
      "add.s64 b2, %2, 4;\n"
      "add.s64 a3, %1, 6;\n"
      "add.s64 b3, %2, 6;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], a1, b1, %3, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], a2, b2, %3, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], a3, b3, %3, q;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_u32(commit_bar))
      : "memory");
}
// 32-wide K block (2 x K=16 bf16 MMAs, e.g. head_dim 32) + commit; same calling convention as umma_bf16_kblock64_warp.
__device__ __forceinline__ void umma_bf16_kblock32_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                        uint32_t accumulate_first, uint64_t* commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, pe;\n"
      ".reg .b64 a1, b1;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 q, %4, %4;\n"
      "add.s64 a1, %1, 2;\n"
      "add.s64 b1, %2, 2;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, q;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_u32(commit_bar))
      : "memory");
}
// Four bf16 MMAs with register descriptor strides (address-field units of 16 bytes), issued by one elected lane of a
// converged warp; no commit.  Used where the K steps are not contiguous inside one swizzle atom (P.V in attention).
__device__ __forceinline__ void umma_bf16_x4_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step,
                                                  uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, pe;\n"
      ".reg .b64 sa, sb, a1, a2, a3, b1, b2, b3;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.eq.b32 q, %6, %6;\n"
      "cvt.u64.u32 sa, %3;\n"
      "cvt.u64.u32 sb, %4;\n"
      "add.s64 a1, %1, sa;\n"
      "add.s64 b1, %2, sb;\n"
      "add.s64 a2, a1, sa;\n"
      "add.s64 b2, b1, sb;\n"
      "add.s64 a3, a2, sa;\n"
      "add.s64 b3, b2, sb;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, p;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %5, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %5, q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %5, q;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// tcgen05.commit from one elected lane of a converged warp
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// fp8 (e4m3/e5m2) dense MMA, K = 32 per instruction.
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make the mbarrier track completion of all prior tcgen05.mma of this thread.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: warp w reads lanes [32*(w%4), +32); thread = one lane (row), 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM address: bits[31:16] = lane, bits[15:0] = column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// ---------------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------------
enum : uint32_t { kSwizzleNone = 0, kSwizzle128 = 2, kSwizzle64 = 4, kSwizzle32 = 6 };

// Shared-memory matrix descriptor.  start address / LBO / SBO are in 16-byte units.
//   K-major, swizzled : rows of <swizzle-bytes>; 8-row groups are SBO bytes apart; LBO unused (=1).
//   MN-major, swizzled: <swizzle-bytes> contiguous along MN, 8 K-rows per group, SBO between K groups,
//                       LBO between MN chunks.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}
// K-major operand tile [rows, 64 bf16] with 128B swizzle (what TMA box {64, rows} SWIZZLE_128B writes).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return umma_smem_desc(saddr, 16, 1024, kSwizzle128);
}
// K-major operand tile [rows, 32 bf16] with 64B swizzle.
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
  return umma_smem_desc(saddr, 16, 512, kSwizzle64);
}
// MN-major operand: [K rows, 64 bf16 along MN] 128B swizzle (a V tile [keys, head_dim=64]).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return umma_smem_desc(saddr, lbo_bytes, 1024, kSwizzle128);
}
// MN-major operand: [K rows, 32 bf16 along MN] 64B swizzle (a V tile [keys, head_dim=32]).
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(uint32_t saddr, uint32_t lbo_bytes) {
  return umma_smem_desc(saddr, lbo_bytes, 512, kSwizzle64);
}

// Instruction descriptor, kind::f16 with fp32 accumulate.  fmt: 0 = f16, 1 = bf16.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N, uint32_t fmt = 1, bool a_mn_major = false,
                                                      bool b_mn_major = false) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// kind::f8f6f4, e4m3 x e4m3 -> fp32 (format code 0 = e4m3, 1 = e5m2)
__host__ __device__ constexpr uint32_t umma_idesc_f8(uint32_t M, uint32_t N, uint32_t afmt = 0, uint32_t bfmt = 0) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}


// ---------------------------------------------------------------------------------
// Block-scaled fp8 (MX): kind::mxf8f6f4.block_scale — e4m3 operands with one ue8m0 scale per 32 K-elements per row.
// Scale factors live in TMEM: for 128 rows x (4 scale bytes = one 128-element K block), lane (r % 32) of EVERY lane
// quadrant holds, in column (r / 32), the 32-bit word {sf(k0), sf(k1), sf(k2), sf(k3)} of row r.  That is exactly what
// `tcgen05.cp.32x128b.warpx4` produces from a 512-byte shared-memory chunk laid out as 32 rows x 16 bytes
// (byte offset (r % 32) * 16 + (r / 32) * 4 + k): each row's 16 bytes become 4 TMEM columns, broadcast to the 4
// quadrants.  The MMA picks byte k of the word through the a_sf_id / b_sf_id fields of the instruction descriptor
// (mirrored in bits 31:30 of the TMEM address operand, as CUTLASS' make_runtime_instr_desc_block_scaled does).
// ---------------------------------------------------------------------------------
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes: multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory descriptor of one 512-byte scale chunk (32 rows x 16 bytes, no swizzle: 8-row core matrices 128 B apart)
__device__ __forceinline__ uint64_t umma_desc_sf_chunk(uint32_t saddr) { return umma_smem_desc(saddr, 0, 128, kSwizzleNone); }

// kind::mxf8f6f4.block_scale instruction descriptor (no c_format field: the accumulator is always fp32).
// scale_format bit 23 = 1 (ue8m0); a/b format 0 = e4m3; sf ids are OR-ed in per MMA: a_sf_id << 29 | b_sf_id << 4.
__host__ __device__ constexpr uint32_t umma_idesc_mxf8(uint32_t M, uint32_t N, uint32_t afmt = 0, uint32_t bfmt = 0) {
  return (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}

// One 128-element K block of a block-scaled GEMM, as ONE asm statement executed by all 32 lanes of the converged MMA
// warp (see umma_bf16_kblock64_warp for why): three scale-chunk copies smem -> TMEM (SFA, and the two chunks covering a
// 192-row B tile), then 4 x (K = 32) MMAs with sf id 0..3, then the commit that frees the smem stage.  tcgen05.cp and
// tcgen05.mma execute in issue order, so the copies of the NEXT k-block cannot overtake these MMAs and one TMEM scale
// buffer is enough.
__device__ __forceinline__ void umma_mxf8_kblock128_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                         uint32_t tmem_sfa, uint32_t tmem_sfb_cp, uint32_t tmem_sfb,
                                                         uint64_t sfa_desc, uint64_t sfb_desc, uint32_t accumulate_first,
                                                         uint64_t* commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, pe;\n"
      ".reg .b64 a1, a2, a3, b1, b2, b3, sfb1;\n"
      ".reg .b32 i1, i2, i3, fa1, fa2, fa3, fb1, fb2, fb3, cpb1;\n"
      "setp.ne.b32 p, %9, 0;\n"
      "setp.eq.b32 q, %9, %9;\n"
      "add.s64 a1, %1, 2;\n"
      "add.s64 b1, %2, 2;\n"
      "add.s64 a2, %1, 4;\n"
      "add.s64 b2, %2, 4;\n"
      "add.s64 a3, %1, 6;\n"
      "add.s64 b3, %2, 6;\n"
      "add.s64 sfb1, %8, 32;\n"               // second 512-byte chunk of the B scales (address field is in 16-byte units)
      "add.u32 cpb1, %5, 4;\n"
      "or.b32 i1, %3, 0x20000010;\n"          // a_sf_id = b_sf_id = 1
      "or.b32 i2, %3, 0x40000020;\n"          // 2
      "or.b32 i3, %3, 0x60000030;\n"          // 3
      "or.b32 fa1, %4, 0x40000000;\n"
      "or.b32 fa2, %4, 0x80000000;\n"
      "or.b32 fa3, %4, 0xC0000000;\n"
      "or.b32 fb1, %6, 0x40000000;\n"
      "or.b32 fb2, %6, 0x80000000;\n"
      "or.b32 fb3, %6, 0xC0000000;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.cp.cta_group::1.32x128b.warpx4 [%4], %7;\n"
      "@pe tcgen05.cp.cta_group::1.32x128b.warpx4 [%5], %8;\n"
      "@pe tcgen05.cp.cta_group::1.32x128b.warpx4 [cpb1], sfb1;\n"
      "@pe tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%4], [%6], p;\n"
      "@pe tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], a1, b1, i1, [fa1], [fb1], q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], a2, b2, i2, [fa2], [fb2], q;\n"
      "@pe tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], a3, b3, i3, [fa3], [fb3], q;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(tmem_sfa), "r"(tmem_sfb_cp), "r"(tmem_sfb), "l"(sfa_desc), "l"(sfb_desc),
      "r"(accumulate_first), "r"(smem_u32(commit_bar))
      : "memory");
}

// ue8m0 block scale for an absolute maximum: the smallest power of two s with amax / s <= 448 (e4m3 max), as the biased
// exponent byte, and 1/s as a float.  amax == 0 maps to exponent 1 (any finite scale works for an all-zero block).
__device__ __forceinline__ uint32_t ue8m0_from_amax(float amax, float& inv_scale) {
  const float t = amax * (1.0f / 448.0f);
  uint32_t e = (__float_as_uint(t) + 0x7FFFFFu) >> 23;   // ceil to the next power of two
  e = e < 1u ? 1u : (e > 253u ? 253u : e);
  inv_scale = __uint_as_float((254u - e) << 23);          // 2^(127 - e)
  return e;
}
// four floats -> four e4m3 bytes (little-endian: f0 in the low byte), round-to-nearest, saturating
__device__ __forceinline__ uint32_t pack_e4m3x4(float f0, float f1, float f2, float f3) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(f1), "f"(f0));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(f3), "f"(f2));
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}

// ---------------------------------------------------------------------------------
// small numeric helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 + MUFU approx error, far below bf16 resolution):
// one MUFU.RCP + one MUFU.EX2 + 7 FMA-class ops
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);
  const float r = fmaf(-p * t, e, 1.0f);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf_precise(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }
// erf-GELU through ONE MUFU: erf(x / sqrt 2) ~= tanh(x * (a + b x^2 + c x^4)) with (a, b, c) fitted by minimax against
// 0.5 x (1 + erf(x / sqrt 2)) on [-8, 8]: max |error| 2.5e-5 (+ tanh.approx ~5e-4 relative on the tanh term) — below
// half a bf16 ulp of the result everywhere.  5 FMA-class ops + 1 MUFU.TANH per element instead of 12 + 2 MUFU, which
// is what keeps the FFN-up epilogue (N = 3072, only 12 k-blocks of MMA per tile) under the MMA time of its tile.
__device__ __forceinline__ float gelu_erf(float x) {
  const float x2 = fminf(x * x, 40.0f);   // beyond |x| ~ 6.3 the tanh is saturated; the clamp keeps the quartic term from flipping the sign
  float p = fmaf(-3.51516790e-4f, x2, 3.70056460e-2f);
  p = fmaf(p, x2, 7.97507884e-1f);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(p * x), hx);
}
// ---- packed fp32x2 math (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 lanes per instruction).  A pair lives in one
// 64-bit register; pk2 / upk2 are register renames, not instructions.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_pair(uint64_t v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
// gelu_erf on a pair: 6 packed FMA-class instructions + 2 FMNMX + 2 MUFU.TANH for two elements
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  float a, b;
  upk2(mul2(x, x), a, b);
  const uint64_t x2 = pk2(fminf(a, 40.0f), fminf(b, 40.0f));
  uint64_t p = fma2(pk2(-3.51516790e-4f, -3.51516790e-4f), x2, pk2(3.70056460e-2f, 3.70056460e-2f));
  p = fma2(p, x2, pk2(7.97507884e-1f, 7.97507884e-1f));
  upk2(mul2(p, x), a, b);
  const uint64_t th = pk2(tanh_approx(a), tanh_approx(b));
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  return fma2(hx, th, hx);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanh_approx(k0 * fmaf(k1 * x * x, x, x)));
}

// ---- NVLS multicast (multimem.*) on addresses of a cuMulticast mapping: one instruction acts on every GPU's copy
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_red_add_release(uint32_t* mc_addr, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
// 8 bf16 sums (fp32 accumulation inside the switch) of the same 16 bytes on every GPU of the multicast group
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const void* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}


// system-scope release / acquire on flags living in (possibly peer) global memory
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace im
