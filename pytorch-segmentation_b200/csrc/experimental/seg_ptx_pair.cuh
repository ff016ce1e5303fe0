// seg_ptx_pair.cuh — inline-PTX wrappers for the CTA-pair (cta_group::2) forms of the primitives in seg_ptx.cuh, for the
// 256-row pair kernel planned in DESIGN.md §9.1.  NOT part of libseg_b200.so yet: nothing includes this header from the
// product; pair_ptx_check.cu instantiates every wrapper so that `tools/check_pair_ptx.sh` proves ptxas (sm_100a) accepts
// each instruction form.  Protocol (cross-checked against the vendored CUTLASS: cute/arch/copy_sm100_tma.hpp,
// cute/arch/tmem_allocator_sm100.hpp, cutlass/arch/barrier.h, cutlass/pipeline/sm100_pipeline.hpp):
//
//   cluster (2,1,1); CTA rank 0 = leader.  One UMMA computes a 256 x N tile: accumulator rows 0..127 live in CTA0's
//   TMEM, rows 128..255 in CTA1's, same column range; A rows and the two halves of B are read from BOTH CTAs' shared
//   memory at the SAME offsets.  Only the leader's MMA thread issues tcgen05.mma.cta_group::2.
//
//   full[stage]   (used in CTA0 only) init count 1.  CTA0's producer: arrive.expect_tx(bytes of BOTH CTAs).  BOTH
//                 producers issue cp.async.bulk.tensor...cta_group::2 with the barrier operand = own address & PEER_MASK
//                 (bit 24 of a shared::cluster address selects the CTA of the pair; clearing it targets CTA0).
//   empty[stage]  (one per CTA) init count 1.  Each producer waits on its OWN empty barrier; the leader releases both
//                 with tcgen05.commit.cta_group::2 ... .multicast::cluster [empty], 0b11.
//   tfull[buf]    (one per CTA) count 1; leader: commit multicast 0b11; each CTA's epilogue warps wait locally and read
//                 their own 128 TMEM lanes.
//   tempty[buf]   (CTA0) count = 2 x epilogue warps; every epilogue warp arrives with mbarrier.arrive.shared::cluster on
//                 (address & PEER_MASK).
//   TMEM          tcgen05.alloc.cta_group::2 by the warp with the same index in BOTH CTAs, same shared-memory result
//                 slot offset; relinquish / dealloc in the cta_group::2 form, dealloc only after a cluster barrier.
//   cluster barrier after mbarrier init (remote arrives need initialised barriers) and before exit (no CTA may leave while
//   its peer can still signal into its shared memory).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace seg {
namespace ptx2 {

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // cute::Sm100MmaPeerBitMask

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}

// ---- TMA loads whose transaction bytes land on CTA0's barrier ----
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// One CTA loads, BOTH CTAs of the mask receive (same shared-memory offset, each CTA's own barrier gets the bytes): the
// alternative design that keeps cta_group::1 MMAs and only shares the weight tile between two M tiles.
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// ---- TMEM (pair allocation) ----
__device__ __forceinline__ void tmem2_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- MMA (leader thread only) ----
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `cta_mask` once the leader's previously issued MMAs have completed
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}
// the cta_group::1 form with multicast (weight-tile sharing design: each CTA's MMAs release the slot in both CTAs)
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}
// epilogue -> leader: "this accumulator buffer is drained", always lands on CTA0's barrier
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}

// instruction descriptor: bf16 x bf16 -> fp32 with M = 256 (pair) and N = n (<= 256)
__host__ __device__ constexpr uint32_t make_idesc_bf16_m256(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

}  // namespace ptx2
}  // namespace seg
