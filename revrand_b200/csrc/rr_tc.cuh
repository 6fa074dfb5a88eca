// tcgen05 / TMEM / mbarrier building blocks for sm_100a (inline PTX).
//
// Layout conventions used by every tensor-core kernel in this library:
//  * Shared-memory operand tiles are K-major, 128-byte swizzled: a tile of R
//    rows (an M or N index) by 64 fp16 (one 128-byte line per row), 1024-byte
//    aligned.  Element (r, c) lives at byte
//        r*128 + (((c>>3) ^ (r&7)) << 4) + (c&7)*2
//    i.e. the 16-byte chunk index is XORed with the row index mod 8 -- the
//    same image TMA SWIZZLE_128B would produce.  8-row groups are 1024 bytes
//    apart (the descriptor's stride byte offset).
//  * One tcgen05.mma (kind::f16) consumes K = 16 elements = 32 bytes of each
//    row; successive K steps advance the descriptor start address by 32 bytes.
//  * Accumulators are fp32 in TMEM: row i of D -> TMEM lane i, column j ->
//    TMEM column base + j.  tcgen05.ld.32x32b lets warp w read lanes
//    32*(w%4) .. 32*(w%4)+31, one lane per thread.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace rr {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- proxies / fences -------------------------------------------------------
// Generic-proxy shared-memory writes (st.shared) -> visible to the async proxy
// (tensor core operand reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// One thread of a fully converged warp (elect.sync): unlike `lane == 0` the
// compiler knows the guarded region runs in exactly one thread, so tcgen05.mma
// operands stay in uniform registers instead of going through an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop per instruction (measured: ~100
// cycles of issue per MMA with `lane == 0`, which starved the tensor pipe).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMEM allocation (one full warp executes these) -------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile(
      "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
          smem_u32(smem_dst)),
      "r"(ncols)
      : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr),
               "r"(ncols)
               : "memory");
}

// ---- descriptors ------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::
// SmemDescriptor bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) with 2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;   // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16, fp16 A/B (both K-major), fp32 D.
// (cute::UMMA::InstrDescriptor: c_format [4,6)=1, a_format [7,10)=0,
// b_format [10,13)=0, a_major bit15=0, b_major bit16=0, N>>3 [17,23),
// M>>4 [24,29).)
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a,
                                            uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all previously issued MMAs of this thread have
// completed (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
          "r"(smem_u32(bar))
      : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns per warp.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, "
      "%30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a
// swizzled K-major tile.
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Split 8 fp32 values into fp16 hi and lo parts (x ~= hi + lo, |err| <= 2^-25
// for |x| <= 1) and pack each into one 16-byte vector.
__device__ __forceinline__ void split8(const float* x, uint4* hi, uint4* lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    float2 back = __half22float2(h[i]);
    l[i] = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
  }
  *hi = *reinterpret_cast<uint4*>(h);
  *lo = *reinterpret_cast<uint4*>(l);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ============================================================================
// CTA-pair (cta_group::2) variants.  A cluster of two CTAs shares one MMA:
// M = 256 (128 accumulator lanes in each CTA's TMEM), the B operand's N rows
// are split half/half between the two CTAs' shared memories, the leader CTA
// (cluster rank 0) issues, and completion is multicast to both CTAs' barriers.
// ============================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Address of the same shared-memory location in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Arrive on a barrier given by its cluster address (possibly in the peer CTA).
// Default (CTA-scope release) semantics on purpose: the data this arrive
// publishes is shared memory consumed by the tensor core through the async
// proxy, ordered by the fence.proxy.async that precedes it; a cluster-scope
// release/acquire makes ptxas emit MEMBAR.GPU / CCTL.IVALL around every
// arrive and inside every try_wait spin (measured: 40% of all stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// try_wait; -DRR_MBAR_SUSPEND_HINT adds a suspend-time hint (the warp is parked
// until the phase completes or the hint, in ns, expires).  Measured on the fused
// value pass: polls are more than half of all issued instructions without the
// hint, yet neither form changes the kernel time, and a microbenchmark with four
// polling warps next to sixteen working ones shows no slowdown either.
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#ifdef RR_MBAR_SUSPEND_HINT
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_cl(bar, parity)) {
#ifdef RR_MBAR_BACKOFF
    __nanosleep(RR_MBAR_BACKOFF);   // experiment: fewer polls, less issue power
#endif
  }
}

__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile(
      "tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
          smem_u32(smem_dst)),
      "r"(ncols)
      : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr),
               "r"(ncols)
               : "memory");
}

// Instruction descriptor with explicit operand format (0 = F16, 1 = BF16,
// 2 = TF32), both operands K-major, fp32 accumulator.
__device__ __forceinline__ uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t desc_a,
                                             uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32_ss(uint32_t tmem_d, uint64_t desc_a,
                                              uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
// once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// TMEM -> registers without the trailing wait (issue several, then wait once).
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, "
      "%30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns per warp, no wait.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// TMEM -> registers: 32 lanes x 8 consecutive fp32 columns per warp, no wait.
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}


// ============================================================================
// kind::i8 (signed 8-bit operands, int32 accumulators) and the 64-byte swizzle.
//
// An i8 operand tile is K-major with 64-byte rows (K = 64 elements) in the
// SWIZZLE_64B layout: byte k of row r lives at
//     r*64 + (((k>>4) ^ ((r>>1)&3)) << 4) + (k&15)
// (cute Swizzle<2,4,3>: address bits [7,9) are XORed into bits [4,6)); 8-row
// groups are 512 bytes apart (stride byte offset) and tiles are 512-byte
// aligned.  One tcgen05.mma.kind::i8 consumes K = 32 bytes of every row; the
// second K step of a 64-byte row advances the descriptor start by 32 bytes.
// ============================================================================
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // LBO: unused for swizzled K-major
  d |= (uint64_t)(512u >> 4) << 32;    // SBO: 8 rows * 64 B
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;              // SWIZZLE_64B
  return d;
}
__device__ __host__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t chunk) {
  return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}
// Instruction descriptor for kind::i8: signed 8-bit A and B (both K-major),
// int32 accumulator (c_format = 2), no saturation.
__device__ __forceinline__ uint32_t make_idesc_i8(uint32_t M, uint32_t N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8_ss(uint32_t tmem_d, uint64_t desc_a,
                                           uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_i8_ss(uint32_t tmem_d, uint64_t desc_a,
                                            uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- bulk (TMA, non-tensor) global -> shared copies ---------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// TMEM -> registers: 32 lanes x 16 consecutive int32 columns per warp, no wait.
__device__ __forceinline__ void tmem_ld16i_nowait(uint32_t taddr, int* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
        "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

}  // namespace tc
}  // namespace rr
